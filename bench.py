#!/usr/bin/env python
"""bench.py -- throughput of the RX hot path (TONES-mode polyphase channelizer) on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one rank per
GPU under torchrun for N>1, ONE JSON line on rank 0.

Workload = BASELINE.json configs[4] built from configs[1]: 64 concurrent 200 MS/s IQ streams, each
channelized by the 2048-channel, 4-tap PFB into 1000 selected tones (pf_average=4), transport buffers
of 1e6 complex samples, sharded by stream over the N GPUs (64/N streams per GPU, stream s -> rank
s mod N, no data-path collective).  One step = 8 N transport buffers (packet periods) of every
stream, i.e. 512e6 input samples per GPU at every N ("weak": per-GPU work is fixed).

  value    inputs already resident in HBM (two alternating 512 MB batches per GPU, i.e. larger than
           the 126 MB L2, so no step can be served from cache); ONE group launch per step; CUDA
           events on the launching stream.
  e2e      the same streams through the host-fed call (gsdr_rx_group_submit / _wait): pinned HOST
           buffers in, pinned host buffers out, one packet period (one buffer of every stream) per
           call, H2D and D2H inside the timed region.  Reported with the plain-cudaMemcpyAsync
           ceiling of the platform measured at the same N (`e2e.pcie`).
  roofline algorithmic bytes (8 + 8*T/N per input sample) / event-timed launch duration vs the
           measured HBM copy peak in MEASURED_PEAKS.json.
  modes    (N=1) every other BASELINE configuration, device-resident, with its own roofline
           fraction: cfg2 single stream, single-buffer launches, cfg1 DIRECT, cfg3 CHIRP, TX chirp,
           cfg4 full duplex.
  cpu_baseline  the NumPy/SciPy port of the same chain (oracle/cpu_port.py) on the host cores, on a
           bounded sample, rank 0 at N=1 only.

`--impl reference` runs the reference's own RX_buffer_demodulator (unmodified sources compiled for
sm_100a into oracle/_ref/libgsdr_ref.so) on the same workload, one instance per stream and one
process per GPU: `value` is its kernels alone (CUDA events on its stream around process(), its two
copies timed separately and subtracted), `e2e` its blocking process() from one worker thread per
stream (its threading model) with pinned host buffers.  The reference has no CPU DSP path, so its
arm runs on the GPU as it does in production.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE, NFFT, PTAPS, NTONES, BUFLEN = 200_000_000, 2048, 4, 1000, 1_000_000
BYTES_PER_SAMPLE = 8.0 + 8.0 * NTONES / NFFT  # 11.90625 B per input sample
TOTAL_STREAMS = 64                            # BASELINE.json configs[4]
SEED = 1337
METRIC = "demodulated input MS/s (1000-tone PFB, whole job)"


def workload_param(stream: int = 0):
    """cfg2 parameters; every stream has its own 1000-tone list (seeded by the global stream number)."""
    import gpu_sdr_b200 as g
    rng = np.random.default_rng(SEED + 7919 * stream)
    ks = rng.choice(np.arange(-NFFT // 2 + 1, NFFT // 2), size=NTONES, replace=False)
    freq = [int(k * (RATE / NFFT)) for k in ks]
    return g.param(mode="RX", rate=RATE, fft_tones=NFFT, pf_average=PTAPS, buffer_len=BUFLEN, decim=0, freq=freq,
                   wave_type=[g.TONES] * NTONES, ampl=[1.0 / NTONES] * NTONES)


def synth_buffers(p, n_distinct, seed):
    """n_distinct transport buffers of tones + noise on the sc16 grid (float32 exact)."""
    rng = np.random.default_rng(seed)
    f = np.array(p.freq[:32], dtype=np.int64)  # 32 of the tones carry power; the rest see noise
    L = int(p.buffer_len)
    out = []
    for b in range(n_distinct):
        n = np.arange(b * L, (b + 1) * L, dtype=np.int64)
        x = np.zeros(L, dtype=np.complex64)
        for fi in f:
            ph = ((fi * n) % p.rate).astype(np.float32) * np.float32(2 * np.pi / p.rate)
            x += (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64) * np.float32(1.0 / 64)
        x += (rng.standard_normal(L, dtype=np.float32) + 1j * rng.standard_normal(L, dtype=np.float32)) * np.float32(1e-3)
        x = (np.round(x.real * 32768) + 1j * np.round(x.imag * 32768)).astype(np.complex64) / np.float32(32768)
        out.append(x.astype(np.complex64))
    return out


SAMPLES_PER_GPU_STEP = 512_000_000            # per-GPU work of one step: streams per GPU x buffers per stream x 1e6


def layout(world: int, streams_total: int = TOTAL_STREAMS):
    """(streams per GPU, buffers per stream per step): every GPU processes 512e6 samples per step at every N -- 64 streams x 8
    buffers on one GPU, 8 streams x 64 buffers on each of eight.  (Short launches cost roofline: every stream boundary inside
    a launch is an extra pipeline fill and drain on some CTA.  64 streams x 1 buffer per launch, the real-time shape on ONE
    GPU, runs at 0.60 of the HBM roofline, 8 streams x 8 buffers at 0.68, this shape at 0.71)"""
    s = max(1, streams_total // world)
    return s, max(1, SAMPLES_PER_GPU_STEP // (s * BUFLEN))


def shared_config(world, S, B):
    """`config` is identical in our arm and the reference arm: it names the workload, nothing else."""
    return {"workload": f"cfg5 (BASELINE configs[4] on configs[1] parameters): {S * world} concurrent 200 MS/s IQ streams x 1000-tone PFB "
                        f"(N={NFFT} channels, P={PTAPS} taps, pf_average={PTAPS}), sharded by stream over {world} GPU(s) = {S} streams per GPU; "
                        f"one step = {B} transport buffer(s) of 1e6 samples per stream = {S * B}e6 input samples per GPU",
            "streams_total": S * world, "streams_per_gpu": S, "buffers_per_stream_per_step": B, "buffer_len": BUFLEN,
            "l2": f"inputs larger than L2: two alternating {S * B * 8} MB batches per GPU"}


# ---- distributed plumbing ------------------------------------------------------------------------
class Dist:
    def __init__(self, n_gpus, force_gloo=False):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            backend = "nccl" if (torch.cuda.is_available() and not force_gloo) else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
                os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
                dist.init_process_group(backend, device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend)
            self.dev = torch.device("cuda", self.local_rank) if backend == "nccl" else torch.device("cpu")

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def _reduce(self, v: float, op) -> float:
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.MAX) if self.dist else v

    def min(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.MIN) if self.dist else v

    def sum(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.SUM) if self.dist else v

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def shard_streams(n_streams: int, rank: int, world: int):
    """stream s -> rank s mod world."""
    return [s for s in range(n_streams) if s % world == rank]


# ---- clocks --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8 or not (t_begin - 0.05 <= ts <= t_end + 0.15):
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu summary, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "pfb_traffic.json")))
    except Exception:
        return None


# ---- CPU baseline --------------------------------------------------------------------------------
def cpu_baseline(p, budget_s=12.0):
    from oracle import cpu_port
    port = cpu_port.PFBPort(p.rate, p.fft_tones, p.pf_average, p.buffer_len, p.freq)
    bufs = synth_buffers(p, 2, SEED + 7)
    port.process(bufs[0])  # warm-up (FFT plan, page faults)
    n, t0 = 0, time.perf_counter()
    while True:
        port.process(bufs[n % 2])
        n += 1
        el = time.perf_counter() - t0
        if (el >= budget_s and n >= 8) or n >= 4000:
            break
    return {"value": n * BUFLEN / el / 1e6, "unit": "MS/s", "cores": cpu_port.cores(), "kind": "port",
            "sample": f"{n} transport buffers of 1e6 samples of one stream ({el:.1f} s) through oracle/cpu_port.PFBPort "
                      f"(complex64 NumPy + scipy.fft workers={cpu_port.cores()})"}


# ---- per-mode device-resident figures (N=1) --------------------------------------------------------
def _noise(n, seed=1):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n, dtype=np.float32) * 0.1 + 1j * rng.standard_normal(n, dtype=np.float32) * 0.1).astype(np.complex64)


def _rx_mode(g, dev, peak, name, p, n_buf, bytes_per_sample, steps, bytes_note=None):
    """Device-resident rate of one RX configuration: n_buf consecutive buffers per launch, two alternating input batches."""
    rx = g.RX_buffer_demodulator(p, device=dev)
    L = int(p.buffer_len)
    base = _noise(L)
    n_alt = 2 if n_buf * L * 8 > 200e6 else max(2, int(300e6 // (n_buf * L * 8)) + 1)   # rotate through > L2 worth of inputs
    ins = []
    for h in range(n_alt):
        d = g.DeviceBuffer(n_buf * L, device=dev)
        for b in range(n_buf):
            d.upload(np.roll(base, 7 * (b + h)), offset=b * L)
        ins.append(d)
    out = g.DeviceBuffer(rx.max_output_batch(n_buf), device=dev)
    for i in range(3):
        rx.process_device(ins[i % n_alt].ptr, n_buf, out.ptr)
    rx.sync()
    l0 = rx.launch_count()
    rx.timer_start()
    for i in range(steps):
        rx.process_device(ins[i % n_alt].ptr, n_buf, out.ptr)
    ms = rx.timer_stop() / steps
    gbs = n_buf * L * bytes_per_sample / (ms * 1e-3) / 1e9
    res = {"mode": name, "kernel": rx.kernel_name(), "value": n_buf * L / (ms * 1e-3) / 1e6, "unit": "MS/s", "buffers_per_launch": n_buf,
           "ms_per_step": ms, "steps": steps, "launches_per_step": (rx.launch_count() - l0) / steps,
           "bytes_per_sample": bytes_per_sample, "achieved_GBps": gbs, "frac": gbs / peak}
    if bytes_note:
        res["bytes_note"] = bytes_note
    rx.close()
    for d in ins:
        d.free()
    out.free()
    return res


def direct_param(g, rate=100_000_000, T=16, decim=100, f=4, L=BUFLEN):
    """cfg1 (same builder as tests/common.py: T distinct integer tones in (-rate/2, rate/2), a quarter of them negative)."""
    rng = np.random.default_rng(SEED + 1)
    freq = []
    while len(freq) < T:
        v = int(rng.integers(-rate // 2 + 1, rate // 2))
        if len(freq) < max(T // 4, 1):
            v = -abs(v) - 1
        if v not in freq:
            freq.append(v)
    return g.param(mode="RX", rate=rate, decim=decim, pf_average=f, buffer_len=L, freq=freq, wave_type=[g.DIRECT] * T, ampl=[1.0 / T] * T,
                   data_mem_mult=max(int(np.ceil(T / max(decim, 1))), 1))


def chirp_param(g, mode="RX", ampl=0.5):
    """cfg3 (get_VNA workload): 100 MHz span, 1e5 points, 1 s, decim 1 -> 2000 samples per point."""
    return g.param(mode=mode, rate=RATE, decim=1, buffer_len=BUFLEN, freq=[-50_000_000], chirp_f=[50_000_000], swipe_s=[100_000],
                   chirp_t=[1.0], wave_type=[g.CHIRP], ampl=[ampl])


def run_modes(g, dev, peak):
    modes = []
    p2 = workload_param(0)
    modes.append(_rx_mode(g, dev, peak, "cfg2: one stream, 64 buffers per launch (the round-1 headline)", p2, 64, BYTES_PER_SAMPLE, 20))
    modes.append(_rx_mode(g, dev, peak, "cfg2: one stream, ONE 1e6-sample buffer per launch (real-time shape), back to back", p2, 1,
                          BYTES_PER_SAMPLE, 200))
    pd = direct_param(g)
    bd = 8 + 8 * 16 / 100
    modes.append(_rx_mode(g, dev, peak, "cfg1: DIRECT T=16 decim=100 pf_average=4, rate 1e8, 64 buffers per launch", pd, 64, bd, 10))
    modes.append(_rx_mode(g, dev, peak, "cfg1: DIRECT, ONE 1e6-sample buffer per launch, back to back", pd, 1, bd, 100))
    pc = chirp_param(g)
    ppt = 2000
    bc = 8 * (1 - (ppt // 10) / ppt) + 8 / ppt
    modes.append(_rx_mode(g, dev, peak, "cfg3: CHIRP lock-in, 1e5 points over 100 MHz, ppt=2000, 64 buffers per launch", pc, 64, bc, 20,
                          bytes_note="the lock-in profile is zero for the first ppt/10 samples of every point (make_flat_window), which the "
                                     "kernel neither reads nor demodulates: 0.9 x 8 B read + 8/ppt B written per input sample"))
    # TX chirp synthesis
    pt = chirp_param(g, mode="TX")
    tx = g.TX_buffer_generator(pt, device=dev)
    n_buf = 64
    d = g.DeviceBuffer(n_buf * BUFLEN, device=dev)
    for _ in range(3):
        tx.get_device(d.ptr, n_buf)
    tx.sync()
    tx.timer_start()
    for _ in range(20):
        tx.get_device(d.ptr, n_buf)
    ms = tx.timer_stop() / 20
    n = n_buf * BUFLEN
    modes.append({"mode": "TX CHIRP synthesis, 64 buffers per launch", "kernel": "chirp_gen_kernel", "value": n / (ms * 1e-3) / 1e6,
                  "unit": "MS/s (output)", "ms_per_step": ms, "steps": 20, "bytes_per_sample": 8.0, "achieved_GBps": n * 8 / (ms * 1e-3) / 1e9,
                  "frac": n * 8 / (ms * 1e-3) / 1e9 / peak})
    tx.close()
    d.free()
    modes.append(full_duplex_mode(g, dev, peak))
    return modes


def full_duplex_mode(g, dev, peak):
    """cfg4: 1000-tone TX buffer synthesis + RX at 200 MS/s on two front-ends (A, B) of one GPU, full size.
    TX: TX_buffer_generator TONES, T=1000, rate 2e8 (one period of 2e8 samples + buffer_len wrap, built once; get() is
    pointer arithmetic, cpp/USRP_buffer_generator.cpp:60-99,226-229).  RX: each front-end channelizes the looped-back TX
    waveform into its 1000 tones (--sw_loop identity): device-resident as one group launch over both front-ends, and
    end to end with two worker threads calling tx.get() + the blocking rx.process() like TXRX::tx_single_link /
    rx_single_link (cpp/USRP_server_link_threads.cpp:542-702)."""
    res = {"mode": "cfg4: full duplex, TX TONES T=1000 synthesis + RX PFB on two front-ends at 200 MS/s"}
    ps = [workload_param(1000 + fe) for fe in range(2)]
    t0 = time.perf_counter()
    txs = []
    for p in ps:
        q = g.param(mode="TX", rate=p.rate, buffer_len=p.buffer_len, freq=p.freq, ampl=p.ampl, wave_type=p.wave_type)
        txs.append(g.TX_buffer_generator(q, device=dev))
    res["tx_init_s_per_front_end"] = (time.perf_counter() - t0) / 2
    res["tx_kernel"] = "tones_synth_kernel (integer-phase GEMM over 2e8 samples x 1000 tones)"
    res["tx_tone_samples_per_s"] = RATE * NTONES / res["tx_init_s_per_front_end"]
    rxs = [g.RX_buffer_demodulator(p, device=dev) for p in ps]
    grp = g.RxGroup(rxs)
    nb = 32
    ins = [[g.DeviceBuffer(nb * BUFLEN, device=dev) for _ in range(2)] for _ in range(2)]
    for fe in range(2):
        for h in range(2):
            txs[fe].get_device(ins[fe][h].ptr, nb)   # consecutive TX buffers, as a loop-back cable would deliver them
        txs[fe].sync()
    outs = [g.DeviceBuffer(rx.max_output_batch(nb), device=dev) for rx in rxs]
    for i in range(3):
        grp.process_device([ins[0][i & 1].ptr, ins[1][i & 1].ptr], nb, [o.ptr for o in outs])
    grp.sync()
    grp.timer_start()
    steps = 20
    for i in range(steps):
        grp.process_device([ins[0][i & 1].ptr, ins[1][i & 1].ptr], nb, [o.ptr for o in outs])
    ms = grp.timer_stop() / steps
    n = 2 * nb * BUFLEN
    gbs = n * BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9
    res.update({"kernel": rxs[0].kernel_name(), "value": n / (ms * 1e-3) / 1e6, "unit": "MS/s (RX input, both front-ends)",
                "ms_per_step": ms, "steps": steps, "bytes_per_sample": BYTES_PER_SAMPLE, "achieved_GBps": gbs, "frac": gbs / peak})
    # known answer of the loop: tones on bin centres with amplitude 1/T -> every selected bin is ~1/T in every frame
    tot, lens = grp.process_device([ins[0][0].ptr, ins[1][0].ptr], 1, [o.ptr for o in outs])
    grp.sync()
    y = outs[0].download(int(lens[0][0]))
    res["loopback_bin_amplitude_times_T"] = float(np.median(np.abs(y))) * NTONES
    grp.close()
    for fe in range(2):
        for h in range(2):
            ins[fe][h].free()
    for o in outs:
        o.free()
    # end to end: worker thread per front-end, tx.get() -> blocking rx.process() on pinned buffers
    for rx in rxs:
        rx.reset()
    houts = [g.pinned_empty(rx.max_output()) for rx in rxs]
    n_pk = 64

    def worker(fe, n_packets):
        for _ in range(n_packets):
            buf = txs[fe].get()
            rxs[fe].process(buf, houts[fe])

    for fe in range(2):
        worker(fe, 4)
    th = [threading.Thread(target=worker, args=(fe, n_pk)) for fe in range(2)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    res["e2e"] = {"value": 2 * n_pk * BUFLEN / dt / 1e6, "unit": "MS/s", "api": "TX_buffer_generator::get + RX_buffer_demodulator::process, "
                  "one thread per front-end", "required_MSps": 2 * RATE / 1e6}
    for o in houts:
        g.pinned_free(o)
    for rx in rxs:
        rx.close()
    for tx in txs:
        tx.close()
    return res


# ---- our arm -------------------------------------------------------------------------------------
def run_ours(args, dd: Dist):
    import gpu_sdr_b200 as g
    lib = g.load()
    ndev = lib.gsdr_device_count()
    if ndev <= 0:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (" + g._lib.last_error() + ")")
    dev = dd.local_rank % ndev
    S, B = layout(dd.world, args.streams_total)
    if args.buffers:
        B = args.buffers
    my_streams = shard_streams(S * dd.world, dd.rank, dd.world)
    assert len(my_streams) == S
    params = [workload_param(s) for s in my_streams]
    rxs = [g.RX_buffer_demodulator(p, device=dev) for p in params]
    kernel_name = rxs[0].kernel_name()
    group = g.RxGroup(rxs)
    # device-resident input: two alternating batches per stream built from 4 distinct buffers
    distinct = synth_buffers(params[0], 4, SEED)
    ins = []
    for s in range(S):
        pair = []
        for half in range(2):
            d = g.DeviceBuffer(B * BUFLEN, device=dev)
            for b in range(B):
                d.upload(distinct[(b + half + s) % 4], offset=b * BUFLEN)
            pair.append(d)
        ins.append(pair)
    outs = [g.DeviceBuffer(rx.max_output_batch(B), device=dev) for rx in rxs]
    # the step is one C-ABI call on pre-built pointer arrays: at N = 8 the ranks share the host's cores, and a step that
    # spends longer in Python than the launch takes on the GPU would be timing the host
    in_arrs = [(C.c_void_p * S)(*[ins[s][h].ptr for s in range(S)]) for h in range(2)]
    out_arr = (C.c_void_p * S)(*[o.ptr for o in outs])
    lens_arr = (C.c_int * (S * B))()

    def step(i):
        if lib.gsdr_rx_group_process_device(group._h, in_arrs[i & 1], B, out_arr, lens_arr) < 0:
            raise SystemExit("gsdr_rx_group_process_device: " + g._lib.last_error())

    for i in range(args.warmup):
        step(i)
    group.sync()
    sampler = ClockSampler(dev)
    time.sleep(0.25)
    # The sampler's start-up leaves the GPU idle for a quarter of a second and the clocks fall back: with a small K the
    # timed region would sit on the ramp (K=3: 0.189 ms per step against 0.167 at K=50).  ~20 ms of the same work, untimed,
    # brings the clocks back before the barrier; the timed region below is still exactly K steps.
    ramp_steps = max(0, 128 - args.warmup)
    for i in range(ramp_steps):
        step(i)
    group.sync()
    dd.barrier()
    group.sync()
    l0 = group.launch_count()
    t_begin = time.time()
    group.timer_start()
    for i in range(args.steps):
        step(args.warmup + i)
    ms_total = group.timer_stop()
    dd.barrier()
    group.sync()
    l1 = group.launch_count()
    # keep the same load running until the clock sampler has seen it for ~1.2 s (100 ms period)
    extra_i = 0
    while time.time() - t_begin < 1.2 and not args.profile:
        step(extra_i)
        extra_i += 1
        if extra_i % 16 == 0:
            group.sync()
    group.sync()
    clocks = sampler.stop(t_begin, time.time())
    clocks["window"] = "timed region + same load continued to 1.2 s"

    ms_step = dd.max(ms_total / args.steps)
    samples_step = dd.sum(float(S * B * BUFLEN))
    value = samples_step / (ms_step * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    achieved = (S * B * BUFLEN * BYTES_PER_SAMPLE) / ((ms_total / args.steps) * 1e-3) / 1e9  # this rank's GB/s
    achieved_min = dd.min(achieved)
    traffic = ncu_traffic()
    for s in range(S):
        for d in ins[s]:
            d.free()
    for o in outs:
        o.free()

    base = {"metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": dd.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(dd.world, S, B), "kernel": kernel_name, "clock_ramp_steps": ramp_steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_min_over_ranks": achieved_min / peak,
                         "traffic": ((traffic or {}).get("dram_bytes_per_sample") or 0) * S * B * BUFLEN or None,
                         "traffic_note": (traffic or {}).get("source"),
                         "peak_source": peak_src, "bytes_per_sample": BYTES_PER_SAMPLE, "samples_per_launch": S * B * BUFLEN,
                         "launch_ms": ms_total / args.steps, "kernel": kernel_name},
            "gpu_launches": int(l1 - l0), "clocks": clocks}
    if args.profile:  # launch-list / ncu runs: device-resident leg only
        base["profile_run"] = True
        group.close()
        for r in rxs:
            r.close()
        return base

    # ---- e2e: host buffers through the host-fed group call, one packet period (one buffer per stream) per call -------------
    for r in rxs:
        r.reset()
    ring = 3
    numa_node = int(lib.gsdr_device_numa_node(dev))
    print(f"[bench rank {dd.rank}] GPU {dev}: NUMA node {numa_node}; pinned buffers allocated node-local when the kernel allows",
          file=sys.stderr, flush=True)
    hin = [[g.pinned_empty(BUFLEN) for _ in range(S)] for _ in range(ring)]
    hout = [[g.pinned_empty(rx.max_output()) for rx in rxs] for _ in range(ring)]
    for k in range(ring):
        for s in range(S):
            hin[k][s][:] = distinct[(k + s) % 4]
    ptrs = [group.pointer_arrays(hin[k], hout[k]) for k in range(ring)]
    e2e_steps = max(2, min(args.steps, 8))

    def e2e_run(n_periods, arrays, sc16=False):
        """n_periods packet periods back to back, at most ring-1 in flight: a continuous stream, drained at the end."""
        tickets, d2h = [], 0
        for b in range(n_periods):
            if len(tickets) >= ring - 1:
                group.wait(tickets.pop(0))
            t, lens = group.submit(*arrays[b % ring], sc16=sc16)
            d2h += 8 * sum(lens)
            tickets.append(t)
        for t in tickets:
            group.wait(t)
        return d2h

    def timed_e2e(arrays, sc16=False):
        e2e_run(max(B, ring, 32), arrays, sc16)  # warm-up (the measured form decides within 4 + 16 periods, a few more if it restarts)
        dd.barrier()
        t0 = time.perf_counter()
        d2h = e2e_run(e2e_steps * B, arrays, sc16)   # e2e_steps steps of B periods each, pipelined across step boundaries
        sec = dd.max((time.perf_counter() - t0) / e2e_steps)
        return dd.sum(float(S * B * BUFLEN)) / sec / 1e6, d2h // e2e_steps, sec

    forms = {}
    d2h_bytes = 0
    form_modes = {"copied": 0, "zero_copy": 1, "copy_in_store_out": 2, "measured": 3}
    for form, mode in form_modes.items():
        lib.gsdr_rx_group_set_zero_copy(group._h, mode)
        for r in rxs:
            r.reset()
        v, d2h_bytes, sec = timed_e2e(ptrs)
        lf = int(lib.gsdr_rx_group_last_form(group._h))
        forms[form] = {"value": v, "unit": "MS/s", "s_per_step": sec, "inputs_read_in_place": bool(lf & 1), "outputs_written_in_place": bool(lf & 2)}
        if form == "measured":   # opt-in (GSDR_GROUP_ZEROCOPY=3): zero-copy and copied timed against each other once, the clearly faster kept
            kept = int(lib.gsdr_rx_group_auto_choice(group._h, 0))
            forms[form]["kept"] = {0: "copied", 1: "zero_copy"}.get(kept, "undecided")
            forms[form]["kept_by_rank_min_max"] = [int(dd.min(float(kept))), int(dd.max(float(kept)))]
    default_form = args.e2e_form
    lib.gsdr_rx_group_set_zero_copy(group._h, form_modes[default_form])
    e2e_val = forms[default_form]["value"]
    # sc16 ingest: the wire format crosses PCIe, conversion on the GPU.  Reported beside the
    # fc32 figure, never instead of it: the reference's interface is fc32.
    for r in rxs:
        r.reset()
    hraw = [[g.pinned_empty(BUFLEN // 2).view(np.int16) for _ in range(S)] for _ in range(ring)]
    for k in range(ring):
        for s in range(S):
            v = hin[k][s].view(np.float32) * np.float32(32767.0)
            hraw[k][s][:] = np.clip(np.round(v), -32768, 32767).astype(np.int16)
    ptrs16 = [group.pointer_arrays(hraw[k], hout[k]) for k in range(ring)]
    lib.gsdr_rx_group_set_zero_copy(group._h, form_modes["zero_copy"])   # the library's default: int16 read in place by the channelizer
    sc16_val, _, _ = timed_e2e(ptrs16, sc16=True)
    lib.gsdr_rx_group_set_zero_copy(group._h, form_modes[default_form])
    # the unchanged blocking drop-in call, RX_buffer_demodulator::process: (a) one stream alone, (b) every stream of this GPU
    # from its own worker thread at once (the reference's threading model, cpp/USRP_server_link_threads.cpp:605-702)
    for r in rxs:
        r.reset()
    n_blk = 32
    for b in range(4):
        rxs[0].process(hin[b % ring][0], hout[b % ring][0])
    t0 = time.perf_counter()
    for b in range(n_blk):
        rxs[0].process(hin[b % ring][0], hout[b % ring][0])
    blocking_one = n_blk * BUFLEN / (time.perf_counter() - t0) / 1e6

    def blk_worker(s, n):
        for b in range(n):
            rxs[s].process(hin[b % ring][s], hout[b % ring][s])

    n_thr = max(4, 2 * B)
    dd.barrier()
    th = [threading.Thread(target=blk_worker, args=(s, n_thr)) for s in range(S)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    blk_sec = dd.max(time.perf_counter() - t0)
    blocking_threads = dd.sum(float(S * n_thr * BUFLEN)) / blk_sec / 1e6
    # what plain cudaMemcpyAsync gives on this platform at this N (every rank at once): the ceiling of any host-fed figure
    pcie = None
    ceilings = {}
    h2d_need = BUFLEN * 8.0
    d2h_need = d2h_bytes / max(1, S * B)   # bytes down per buffer
    for nq in (1, 4):
        # allocate first, then barrier before every timed pass: all ranks pull on the host at the same moment, for ~0.3 s a pass
        probe = lib.gsdr_pcie_probe_create(dev, BUFLEN * 8, NTONES * 488 * 8, nq, ring * S)
        ok = dd.min(1.0 if probe else 0.0) > 0
        gbs = [0.0, 0.0, 0.0, 0.0]
        reps = 0
        if ok:
            w = (C.c_double * 2)()
            dd.barrier()
            ok = lib.gsdr_pcie_probe_run(probe, 1, 1, 64, w) == 0 and w[0] > 0
            reps = int(dd.min(float(max(32, min(4096, int(0.3 * w[0] * 1e9 / h2d_need)))) if ok else 0.0))
            for up, dn, iu, idn in ((1, 0, 0, None), (0, 1, None, 1), (1, 1, 2, 3)):
                best = None   # two passes of each kind, the better kept (a ceiling: transients only ever read low)
                for _ in range(2):
                    dd.barrier()
                    if reps and lib.gsdr_pcie_probe_run(probe, up, dn, reps, w) == 0:
                        if best is None or w[0] + w[1] > best[0] + best[1]:
                            best = (w[0], w[1])
                    else:
                        ok = False
                if best is not None:
                    if iu is not None:
                        gbs[iu] = best[0]
                    if idn is not None:
                        gbs[idn] = best[1]
            dd.barrier()
        if probe:
            lib.gsdr_pcie_probe_destroy(probe)
        ok = dd.min(1.0 if ok else 0.0) > 0
        t_buf = max(h2d_need / (gbs[2] * 1e9), d2h_need / (gbs[3] * 1e9)) if ok and gbs[2] > 0 and gbs[3] > 0 else float("inf")
        ceil_total = dd.sum(BUFLEN / t_buf / 1e6)
        ceilings[nq] = {"copy_queues_per_direction": nq, "buffers_per_pass": reps, "h2d_alone_GBps": gbs[0], "d2h_alone_GBps": gbs[1],
                        "h2d_duplex_GBps": gbs[2], "d2h_duplex_GBps": gbs[3], "min_over_ranks_h2d_duplex_GBps": dd.min(gbs[2]),
                        "sum_over_ranks_duplex_GBps": dd.sum(gbs[2] + gbs[3]), "e2e_ceiling_MSps": ceil_total if ok else None}
    best = max((c for c in ceilings.values() if c["e2e_ceiling_MSps"]), key=lambda c: c["e2e_ceiling_MSps"], default=None)
    if best:
        pcie = dict(best)
        pcie.update({"rank": 0, "all": list(ceilings.values()),
                     "how": "gsdr_pcie_probe: one cudaMemcpyAsync per 8 MB buffer up and per 3.9 MB buffer down, the e2e loop's own pinned-buffer footprint, "
                            "1 and 4 copy queues per direction, both rates over the same interval, two passes of ~0.3 s each (the better kept), every rank "
                            "started at a barrier; "
                            "the better of the two",
                     "e2e_frac_of_ceiling": e2e_val / best["e2e_ceiling_MSps"]})

    res = dict(base)
    res["e2e"] = {"value": e2e_val, "unit": "MS/s", "h2d_bytes_per_step": S * B * BUFLEN * 8, "d2h_bytes_per_step": d2h_bytes,
                  "api": f"gsdr_rx_group_submit/gsdr_rx_group_wait: {S} pinned host buffers in and {S} out per call (one packet period), "
                         f"depth-3 pipeline, {default_form} form", "steps": e2e_steps, "forms": forms, "default_form": default_form,
                  "real_time_need_MSps": S * dd.world * RATE / 1e6, "pcie": pcie,
                  "blocking_process_value": blocking_one,
                  "blocking_process_all_streams_threads": {"value": blocking_threads, "unit": "MS/s", "threads_per_gpu": S,
                                                           "api": "gsdr_rx_process (RX_buffer_demodulator::process), one worker thread per stream"},
                  "pinned_numa_node_rank0": numa_node,
                  "sc16_ingest": {"value": sc16_val, "unit": "MS/s", "h2d_bytes_per_step": S * B * BUFLEN * 4,
                                  "api": "gsdr_rx_group_submit_sc16 (int16 I/Q in, converted inside the channelizer), zero_copy form"}}
    group.close()
    for r in rxs:
        r.close()
    for k in range(ring):
        for a in hin[k] + hout[k]:
            g.pinned_free(a)
        for a in hraw[k]:
            g.pinned_free(a.view(np.complex64))
    if dd.rank == 0 and dd.world == 1 and not args.no_modes:
        res["modes"] = run_modes(g, dev, peak)
    if dd.rank == 0 and dd.world == 1 and not args.no_cpu:
        res["cpu_baseline"] = cpu_baseline(params[0])
    return res


# ---- reference arm -------------------------------------------------------------------------------
def run_reference(args, dd: Dist):
    import importlib.util  # test infrastructure: the reference's own object code behind tests/common.py
    spec = importlib.util.spec_from_file_location("gsdr_tests_common", os.path.join(ROOT, "tests", "common.py"))
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    RefRX, ref_lib = tc.RefRX, tc.ref_lib
    S, B = layout(dd.world, args.streams_total)
    if args.buffers:
        B = args.buffers
    my_streams = shard_streams(S * dd.world, dd.rank, dd.world)
    params = [workload_param(s) for s in my_streams]
    lib = ref_lib()
    have = lib is not None and lib.gsdr_ref_device_count() > 0
    if not (dd.min(1.0 if have else 0.0) > 0):
        # reference object library not built / no GPU: the oracle port is the only runnable restatement (rank 0, one stream)
        if dd.rank != 0:
            return None
        cpu = cpu_baseline(params[0], budget_s=8.0)
        return {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "MS/s", "n_gpus": dd.world, "steps": args.steps,
                "warmup": args.warmup, "higher_is_better": True, "ms_per_step": S * B * BUFLEN / (cpu["value"] * 1e6) * 1e3,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": shared_config(dd.world, S, B),
                "impl_detail": "oracle/_ref unavailable: CPU port of the chain (oracle/cpu_port.py)", "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    refs = [RefRX(p) for p in params]
    distinct = synth_buffers(params[0], 4, SEED)
    ring = 2
    hin, keep = [], []
    for s in range(S):
        for k in range(ring):
            ptr = lib.gsdr_ref_host_alloc(BUFLEN * 8)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(2 * BUFLEN,)).view(np.complex64)
            a[:] = distinct[(k + s) % 4]
            hin.append(ptr)
            keep.append(a)
    cap = NTONES * lib.gsdr_ref_rx_batching(refs[0].h)
    houts = [lib.gsdr_ref_host_alloc(cap * 8) for _ in range(S)]
    hs = (C.c_void_p * S)(*[r.h for r in refs])
    ins_arr = (C.c_void_p * (S * ring))(*hin)
    outs_arr = (C.c_void_p * S)(*houts)

    # e2e: the reference's threading model, one worker thread per front-end calling the blocking process()
    for _ in range(args.warmup):
        lib.gsdr_ref_rx_multi_process_timed(hs, S, ins_arr, ring, outs_arr, B)
    dd.barrier()
    t = [lib.gsdr_ref_rx_multi_process_timed(hs, S, ins_arr, ring, outs_arr, B) for _ in range(args.steps)]
    e2e_sec = dd.max(sum(t) / len(t))
    e2e_val = dd.sum(float(S * B * BUFLEN)) / e2e_sec / 1e6
    # kernels alone: CUDA events on each instance's stream around process(), its two copies timed alone and subtracted;
    # streams one after the other, so the figure is the GPU time of the reference's launches for one step
    tot_ms, cop_ms = C.c_double(0), C.c_double(0)
    kern_ms_steps = []
    n_kernel_steps = args.warmup + args.steps
    for it in range(n_kernel_steps):
        k_ms = 0.0
        for s in range(S):
            sub = (C.c_void_p * ring)(*hin[s * ring:(s + 1) * ring])
            rc = lib.gsdr_ref_rx_process_split_timed(refs[s].h, sub, ring, houts[s], B, BUFLEN * 8, cap * 8, C.byref(tot_ms), C.byref(cop_ms))
            if rc != 0:
                raise SystemExit("reference split timing failed")
            k_ms += max(tot_ms.value - cop_ms.value, 0.0)
        if it >= args.warmup:
            kern_ms_steps.append(k_ms)
    ms_step = dd.max(sum(kern_ms_steps) / len(kern_ms_steps))
    val = dd.sum(float(S * B * BUFLEN)) / (ms_step * 1e-3) / 1e6
    for r in refs:
        r.close()
    if dd.rank != 0:
        return None
    cpu = cpu_baseline(params[0], budget_s=8.0) if (dd.world == 1 and not args.no_cpu) else None
    res = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MS/s", "n_gpus": dd.world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": shared_config(dd.world, S, B),
           "impl_detail": {"reference": "unmodified cpp/kernels.cu + cpp/USRP_demodulator.cpp compiled for sm_100a (oracle/_ref/libgsdr_ref.so), one "
                                        "RX_buffer_demodulator per stream, one process per GPU; the reference has no CPU DSP path, so this arm runs "
                                        "its CUDA path on the same B200(s)",
                           "value": "kernels only: CUDA events on the instance's stream around RX_buffer_demodulator::process minus its two copies "
                                    "(timed alone on the same stream), summed over the GPU's streams (cpp/USRP_demodulator.cpp:486-565: "
                                    "polyphase_filter + cufftExecC2C + move_buffer + tone_select)",
                           "e2e": "blocking process() with pinned host buffers from one worker thread per stream "
                                  "(cpp/USRP_server_link_threads.cpp:605-702)"},
           "e2e": {"value": e2e_val, "unit": "MS/s", "h2d_bytes_per_step": S * B * BUFLEN * 8, "d2h_bytes_per_step": S * B * cap * 8,
                   "api": "RX_buffer_demodulator::process, one thread per stream", "steps": args.steps},
           "gpu_launches": 0}
    if cpu:
        res["cpu_baseline"] = cpu
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams-total", type=int, default=TOTAL_STREAMS, help="IQ streams of the whole job (cfg5: 64)")
    ap.add_argument("--buffers", type=int, default=0, help="transport buffers per stream per step (default: 512 / streams per GPU)")
    ap.add_argument("--e2e-form", default="zero_copy", choices=["measured", "zero_copy", "copied", "copy_in_store_out"],
                    help="form of the host-fed call behind e2e.value (zero_copy is the library's default; measured = "
                         "GSDR_GROUP_ZEROCOPY=3, the group times zero_copy against copied once and keeps the clearly faster; every form "
                         "is reported under e2e.forms)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-modes", action="store_true", help="skip the per-configuration `modes` legs (N=1)")
    ap.add_argument("--profile", action="store_true", help="device-resident leg only (for ncu launch lists / captures)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    if args.impl == "reference" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # the reference never selects a device (it runs on the current one): give every rank its own GPU before CUDA starts
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        os.environ["CUDA_VISIBLE_DEVICES"] = vis[lr % len(vis)] if vis else str(lr)
    # stdout carries exactly ONE JSON line: whatever a library prints to file descriptor 1 (NCCL's version banner does,
    # whatever NCCL_DEBUG_FILE says) is sent to stderr, and the result goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    dd = Dist(args.gpus, force_gloo=(args.impl == "reference"))
    try:
        res = run_reference(args, dd) if args.impl == "reference" else run_ours(args, dd)
        if dd.rank == 0 and res is not None:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(res) + "\n").encode())
    finally:
        dd.close()


if __name__ == "__main__":
    main()
