#!/usr/bin/env python
"""First-light check of the tensor-core DIRECT kernel on a GPU box: relative L2 error against the fp64 oracle for a
few shapes (tc and fp32 variants) and device-resident throughput at cfg1.  Prints one JSON line per measurement."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_sdr_b200 as g  # noqa: E402
from common import direct_param, orc, rx_run, tone_stream  # noqa: E402


def err_case(tag, p, nbuf=2):
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(nbuf)]
    res = {"case": tag}
    for variant in ("fp32", "tc"):
        os.environ["GSDR_DIRECT_VARIANT"] = variant
        try:
            ours = rx_run(p, bufs)
            o = orc.DirectDemodulator(p.rate, p.freq, p.decim, p.pf_average, p.buffer_len)
            res[variant] = max(orc.rel_l2(a, o.process(x)) for a, x in zip(ours, bufs))
        except Exception as e:  # noqa: BLE001
            res[variant] = "ERROR " + str(e)[:200]
    print(json.dumps(res), flush=True)


def speed(tag, p, n_buf, steps=20):
    rng = np.random.default_rng(1)
    L = p.buffer_len
    base = (rng.standard_normal(L, dtype=np.float32) * 0.1 + 1j * rng.standard_normal(L, dtype=np.float32) * 0.1).astype(np.complex64)
    for variant in ("fp32", "tc"):
        os.environ["GSDR_DIRECT_VARIANT"] = variant
        rx = g.RX_buffer_demodulator(p)
        ins = []
        for h in range(2):
            d = g.DeviceBuffer(n_buf * L)
            for b in range(n_buf):
                d.upload(np.roll(base, 7 * (b + h)), offset=b * L)
            ins.append(d)
        out = g.DeviceBuffer(rx.max_output_batch(n_buf))
        for i in range(3):
            rx.process_device(ins[i & 1].ptr, n_buf, out.ptr)
        rx.sync()
        rx.timer_start()
        for i in range(steps):
            rx.process_device(ins[i & 1].ptr, n_buf, out.ptr)
        ms = rx.timer_stop() / steps
        T = len(p.freq)
        gs = n_buf * L / (ms * 1e-3) / 1e9
        print(json.dumps({"speed": tag, "variant": variant, "kernel": rx.kernel_name(), "n_buf": n_buf, "ms": ms, "input_GSps": gs,
                          "fp32_equiv_TFLOPs": gs * 1e9 * 16 * T * p.pf_average / 4 * 2 / 1e12}), flush=True)
        rx.close()
        for d in ins:
            d.free()
        out.free()


if __name__ == "__main__":
    err_case("small T=4 M=20 f=4", direct_param(rate=1_000_000, T=4, decim=20, f=4, L=40_000))
    err_case("cfg1", direct_param())
    err_case("f=1 T=70 M=25", direct_param(rate=10_000_000, T=70, decim=25, f=1, L=50_000))
    err_case("f=8 T=5 M=10", direct_param(rate=1_000_000, T=5, decim=10, f=8, L=50_000))
    err_case("f=2 T=33 M=50", direct_param(rate=10_000_000, T=33, decim=50, f=2, L=50_000))
    if len(sys.argv) > 1 and sys.argv[1] == "errors":
        sys.exit(0)
    speed("cfg1 T=16 M=100 f=4", direct_param(), 16)
    speed("cfg1 single buffer", direct_param(), 1, steps=50)
    speed("T=1000 M=1000 f=4", direct_param(T=1000, decim=1000), 2, steps=3)
    speed("T=64 M=100 f=4", direct_param(T=64), 8, steps=5)
