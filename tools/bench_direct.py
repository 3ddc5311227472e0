#!/usr/bin/env python
"""cfg1 DIRECT (T=16, decim=100, pf_average=4, rate 1e8) device-resident rate and accuracy per kernel variant
(GSDR_DIRECT_VARIANT = i8 | tc | fp32), 64 buffers and ONE buffer per launch.  One JSON line per case.
Usage (GPU box): python tools/bench_direct.py > gpurun_out/direct.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_sdr_b200 as g  # noqa: E402
from common import direct_param, orc, tone_stream  # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0

cases = [("cfg1 T=16 decim=100 f=4", dict(), 64), ("cfg1, one buffer per launch", dict(), 1), ("T=64 decim=100 f=4", dict(T=64), 16),
         ("T=1000 decim=1000 f=4", dict(T=1000, decim=1000), 2)]
for variant in (sys.argv[1:] or ["i8", "tc", "fp32"]):
    os.environ["GSDR_DIRECT_VARIANT"] = variant
    for name, kw, n_buf in cases:
        if variant == "fp32" and kw.get("T", 16) > 64:
            continue
        p = direct_param(**kw)
        L, T, M = p.buffer_len, len(p.freq), p.decim
        rx = g.RX_buffer_demodulator(p)
        bufs = [tone_stream(p.rate, p.freq, p.ampl, i * L, L) for i in range(2)]
        n_alt = 2 if n_buf * L * 8 > 200e6 else 24
        ins = []
        for h in range(n_alt):
            d = g.DeviceBuffer(n_buf * L)
            for b in range(n_buf):
                d.upload(bufs[(b + h) & 1], offset=b * L)
            ins.append(d)
        out = g.DeviceBuffer(rx.max_output_batch(n_buf))
        for i in range(3):
            rx.process_device(ins[i % n_alt].ptr, n_buf, out.ptr)
        rx.sync()
        steps = 10 if n_buf > 1 else 100
        rx.timer_start()
        for i in range(steps):
            rx.process_device(ins[i % n_alt].ptr, n_buf, out.ptr)
        ms = rx.timer_stop() / steps
        # accuracy on the first two buffers of a fresh stream
        rx.reset()
        host_out = g.pinned_empty(rx.max_output())
        o = orc.DirectDemodulator(p.rate, p.freq, M, p.pf_average, L)
        num = den = 0.0
        for x in bufs:
            n = rx.process(x, host_out)
            want = o.process(x)
            num += float(np.sum(np.abs(host_out[:n].astype(np.complex128) - want) ** 2))
            den += float(np.sum(np.abs(want) ** 2))
        bps = 8 + 8 * T / M
        gs = n_buf * L / (ms * 1e-3) / 1e9
        print(json.dumps({"case": name, "variant": variant, "kernel": rx.kernel_name(), "buffers_per_launch": n_buf, "ms": ms,
                          "input_GSps": gs, "hbm_frac": gs * bps / PEAK, "rel_l2_vs_fp64": float(np.sqrt(num / den)),
                          "fp32_equiv_TFLOPs": gs * 1e9 * 16 * T * 2 / 1e12}), flush=True)
        rx.close()
        for d in ins:
            d.free()
        out.free()
