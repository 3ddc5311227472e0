#!/usr/bin/env python
"""North-star check at per-GPU scale: S concurrent 200 MS/s IQ streams, each channelized into 1000 tones (cfg2
parameters), fed in REAL TIME from pinned host buffers (one 1e6-sample transport buffer every 5 ms per stream, as a
USRP would deliver them), through the pipelined host API (gsdr_rx_submit / gsdr_rx_wait), one host thread and one
demodulator instance per stream (the reference's threading model, cpp/USRP_server_link_threads.cpp:605-702).
64 streams on an 8-GPU box = 8 streams per GPU; streams are independent, so one GPU is the unit of proof.

Reports per stream: packets processed, worst and mean completion latency (arrival -> result in host memory), and
whether the stream ever fell behind (a packet submitted later than the next packet's arrival time).
Usage (GPU box): python tools/realtime_streams.py --streams 8 --seconds 4 [--sc16]"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definition: cfg2)
import gpu_sdr_b200 as g  # noqa: E402


def stream_worker(idx, p, args, t_start, result):
    rx = g.RX_buffer_demodulator(p)
    depth = 3
    L = p.buffer_len
    src = bench.synth_buffers(p, 2, seed=1000 + idx)
    if args.sc16:
        hin = [g.pinned_empty(L // 2).view(np.int16) for _ in range(depth)]
        for k in range(depth):
            hin[k][:] = np.clip(np.round(src[k % 2].view(np.float32) * 32767.0), -32768, 32767).astype(np.int16)
        submit = rx.submit_sc16
    else:
        hin = [g.pinned_empty(L) for _ in range(depth)]
        for k in range(depth):
            hin[k][:] = src[k % 2]
        submit = rx.submit
    hout = [g.pinned_empty(rx.max_output()) for _ in range(depth)]
    period = L / float(p.rate)  # 5 ms
    n_packets = int(args.seconds / period)
    lat, late, pending = [], 0, []
    result["barrier"].wait()   # every stream is set up
    result["ready"].wait()     # the main thread has published the common start time
    t0 = t_start[0]
    for k in range(n_packets):
        arrival = t0 + (k + 1) * period  # the buffer is complete when its last sample has arrived
        now = time.perf_counter()
        if now < arrival:
            time.sleep(arrival - now)  # no spinning: eight spinning threads would fight over the GIL
        elif now > arrival + period:
            late += 1
        if args.blocking:  # the reference's call pattern: one blocking process() per packet; latency = arrival -> result
            t, _ = submit(hin[k % depth], hout[k % depth])
            rx.wait(t)
            lat.append(time.perf_counter() - arrival)
            continue
        if len(pending) >= depth - 1:
            t, arr = pending.pop(0)
            rx.wait(t)
            lat.append(time.perf_counter() - arr)
        t, _ = submit(hin[k % depth], hout[k % depth])
        pending.append((t, arrival))
    for t, arr in pending:
        rx.wait(t)
        lat.append(time.perf_counter() - arr)
    rx.close()
    result[idx] = {"packets": n_packets, "late_submits": late, "latency_ms_mean": 1e3 * float(np.mean(lat)),
                   "latency_ms_max": 1e3 * float(np.max(lat))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--sc16", action="store_true")
    ap.add_argument("--blocking", action="store_true", help="wait for each packet right after submitting it (true latency)")
    args = ap.parse_args()
    p = bench.workload_param()
    result = {"ready": threading.Event(), "barrier": threading.Barrier(args.streams + 1)}
    t_start = [0.0]
    th = [threading.Thread(target=stream_worker, args=(i, p, args, t_start, result)) for i in range(args.streams)]
    for t in th:
        t.start()
    result["barrier"].wait()  # instance creation (twiddles, layouts, pinned buffers) is not part of the real-time run
    t_start[0] = time.perf_counter() + 0.05
    result["ready"].set()
    for t in th:
        t.join()
    wall = time.perf_counter() - t_start[0]
    per = [result[i] for i in range(args.streams)]
    total_samples = sum(r["packets"] for r in per) * p.buffer_len
    out = {"test": "real-time multi-stream channelizer (cfg2 per stream)", "streams": args.streams, "input": "sc16" if args.sc16 else "fc32", "call": "blocking" if args.blocking else "pipelined depth 3 (latency includes two packet periods of queueing by design)",
           "stream_rate_MSps": p.rate / 1e6, "seconds": args.seconds, "wall_s": wall,
           "aggregate_input_MSps": total_samples / wall / 1e6, "required_MSps": args.streams * p.rate / 1e6,
           "late_submits_total": sum(r["late_submits"] for r in per),
           "latency_ms_mean": float(np.mean([r["latency_ms_mean"] for r in per])),
           "latency_ms_max": float(np.max([r["latency_ms_max"] for r in per])),
           "sustained": all(r["late_submits"] == 0 for r in per)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
