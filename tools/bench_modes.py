#!/usr/bin/env python
"""Device-resident throughput of every kernel on the path (not only the headline PFB): one JSON line
per mode with input MS/s, algorithmic GB/s and the fraction of the measured HBM copy peak.
CUDA events on the launching stream, 3 warm-up passes, inputs larger than L2 where the mode is
memory-bound.  Usage (GPU box): python tools/bench_modes.py > gpurun_out/modes.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_sdr_b200 as g  # noqa: E402
from common import chirp_param, direct_param, pfb_param  # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def noise(n, seed=1):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n, dtype=np.float32) * 0.1 + 1j * rng.standard_normal(n, dtype=np.float32) * 0.1).astype(np.complex64)


def run_rx(name, p, n_buf, bytes_per_sample, steps=20, extra=None):
    rx = g.RX_buffer_demodulator(p)
    L = p.buffer_len
    base = noise(L)
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
    l0 = rx.launch_count()
    rx.timer_start()
    for i in range(steps):
        rx.process_device(ins[i & 1].ptr, n_buf, out.ptr)
    ms = rx.timer_stop() / steps
    res = {"mode": name, "kernel": rx.kernel_name(), "samples_per_step": n_buf * L, "ms_per_step": ms,
           "input_MSps": n_buf * L / (ms * 1e-3) / 1e6, "bytes_per_sample": bytes_per_sample,
           "algorithmic_GBps": n_buf * L * bytes_per_sample / (ms * 1e-3) / 1e9,
           "launches_per_step": (rx.launch_count() - l0) / steps}
    res["hbm_frac_of_measured"] = res["algorithmic_GBps"] / PEAK
    if extra:
        res.update(extra(res))
    rx.close()
    for d in ins:
        d.free()
    out.free()
    print(json.dumps(res), flush=True)


def main():
    # cfg2 headline and its multi-P / generic relatives
    run_rx("cfg2 TONES N=2048 P=4 T=1000 (fused)", pfb_param(), 64, 8 + 8 * 1000 / 2048)
    run_rx("TONES N=2048 P=2 T=100 (fused)", pfb_param(P=2, T=100), 64, 8 + 8 * 100 / 2048)
    run_rx("NOISE N=2048 P=4 full spectrum (fused)", g.param(rate=200_000_000, fft_tones=2048, pf_average=4, buffer_len=1_000_000,
                                                              freq=[0], wave_type=[g.NOISE], ampl=[1.0]), 32, 16.0)
    # channel counts without a fused kernel (the pyUSRP client sets fft_tones = its decimation factor): the filter bank as a
    # GEMM on the tensor cores (default) against the CUDA-core FIR + DFT pair (GSDR_PFB_VARIANT=generic)
    for variant in ("tc", "generic"):
        if variant == "generic":
            os.environ["GSDR_PFB_VARIANT"] = "generic"
        run_rx(f"TONES N=100 P=4 T=16 [{variant}]", pfb_param(rate=100_000_000, N=100, P=4, T=16), 8, 8 + 8 * 16 / 100, steps=5,
               extra=lambda r: {"fp32_equiv_TFLOPs": r["input_MSps"] * 1e6 * 8 * 4 * 16 / 1e12})
        run_rx(f"TONES N=1000 P=4 T=100 [{variant}]", pfb_param(rate=100_000_000, N=1000, P=4, T=100), 8, 8 + 8 * 100 / 1000, steps=5,
               extra=lambda r: {"fp32_equiv_TFLOPs": r["input_MSps"] * 1e6 * 8 * 4 * 100 / 1e12})
    os.environ.pop("GSDR_PFB_VARIANT", None)
    # cfg1 DIRECT: 16*T real FMA per input sample.  Default = tensor-core kernel (tcgen05, 3xTF32 split GEMM) where the
    # shape fills the GPU and decim <= 128; GSDR_DIRECT_VARIANT=fp32 forces the CUDA-core kernel for comparison.
    T = 16
    flops = lambda t: (lambda r: {"fp32_equiv_TFLOPs": r["input_MSps"] * 1e6 * 16 * t * 2 / 1e12})  # noqa: E731
    for variant in ("tc", "fp32"):
        os.environ["GSDR_DIRECT_VARIANT"] = variant
        run_rx(f"cfg1 DIRECT T=16 decim=100 f=4 [{variant}]", direct_param(), 16, 8 + 8 * T / 100, extra=flops(16))
        run_rx(f"cfg1 DIRECT, one 1e6-sample buffer per launch [{variant}]", direct_param(), 1, 8 + 8 * T / 100, steps=50, extra=flops(16))
        run_rx(f"DIRECT T=64 decim=100 f=4 [{variant}]", direct_param(T=64), 8, 8 + 8 * 64 / 100, steps=5, extra=flops(64))
        run_rx(f"DIRECT T=1000 decim=1000 f=4 [{variant}]", direct_param(T=1000, decim=1000), 2, 8 + 8.0, steps=3, extra=flops(1000))
    os.environ.pop("GSDR_DIRECT_VARIANT", None)
    run_rx("DIRECT T=16 decim=0 (mix only)", direct_param(decim=0, f=1), 4, 8 + 8 * T, steps=5)
    # cfg3 CHIRP
    run_rx("cfg3 CHIRP lock-in ppt=2000", chirp_param(), 64, 8 + 8 / 2000)
    run_rx("CHIRP decim=0 (demod only)", chirp_param(decim=0), 32, 16.0)
    run_rx("CHIRP true chirp length=1 ppt=200", chirp_param(steps=0, t=0.001, decim=200), 64, 8 + 8 / 200)
    # TX
    p = chirp_param(ampl=0.5)
    p.mode = "TX"
    tx = g.TX_buffer_generator(p)
    n_buf = 64
    d = g.DeviceBuffer(n_buf * p.buffer_len)
    for _ in range(3):
        tx.get_device(d.ptr, n_buf)
    tx.sync()
    tx.timer_start()
    for _ in range(20):
        tx.get_device(d.ptr, n_buf)
    ms = tx.timer_stop() / 20
    n = n_buf * p.buffer_len
    print(json.dumps({"mode": "TX CHIRP synthesis", "kernel": "chirp_gen_kernel", "samples_per_step": n, "ms_per_step": ms,
                      "output_MSps": n / (ms * 1e-3) / 1e6, "bytes_per_sample": 8.0, "algorithmic_GBps": n * 8 / (ms * 1e-3) / 1e9,
                      "hbm_frac_of_measured": n * 8 / (ms * 1e-3) / 1e9 / PEAK}), flush=True)
    tx.close()
    d.free()
    import time
    for rate, T in ((20_000_000, 1000), (200_000_000, 16)):
        rng = np.random.default_rng(3)
        freq = [int(v) for v in rng.choice(np.arange(-rate // 2 + 1, rate // 2), size=T, replace=False)]
        pt = g.param(mode="TX", rate=rate, buffer_len=1_000_000, freq=freq, ampl=[1.0 / T] * T, wave_type=[g.TONES] * T)
        t0 = time.perf_counter()
        tx = g.TX_buffer_generator(pt)
        dt = time.perf_counter() - t0
        print(json.dumps({"mode": f"TX TONES period synthesis rate={rate} T={T} (init, incl. pinned alloc + D2H)",
                          "kernel": "tones_synth_kernel", "seconds": dt, "period_samples": rate,
                          "tone_samples_per_s": rate * T / dt}), flush=True)
        tx.close()
    # multi-stream group: 8 streams x 8 buffers in ONE launch
    ps = [pfb_param(seed=100 + s) for s in range(8)]
    rxs = [g.RX_buffer_demodulator(p) for p in ps]
    grp = g.RxGroup(rxs)
    L, nb = 1_000_000, 8
    base = noise(L)
    ins, outs = [], []
    for s in range(8):
        dd = g.DeviceBuffer(nb * L)
        for b in range(nb):
            dd.upload(np.roll(base, 11 * (b + s)), offset=b * L)
        ins.append(dd)
        outs.append(g.DeviceBuffer(rxs[s].max_output_batch(nb)))
    for _ in range(3):
        grp.process_device([x.ptr for x in ins], nb, [x.ptr for x in outs])
    grp.sync()
    grp.timer_start()
    for _ in range(20):
        grp.process_device([x.ptr for x in ins], nb, [x.ptr for x in outs])
    ms = grp.timer_stop() / 20
    n = 8 * nb * L
    bps = 8 + 8 * 1000 / 2048
    print(json.dumps({"mode": "cfg5-style group: 8 streams x 8 buffers, one launch", "kernel": rxs[0].kernel_name(),
                      "samples_per_step": n, "ms_per_step": ms, "input_MSps": n / (ms * 1e-3) / 1e6,
                      "algorithmic_GBps": n * bps / (ms * 1e-3) / 1e9, "hbm_frac_of_measured": n * bps / (ms * 1e-3) / 1e9 / PEAK}), flush=True)


if __name__ == "__main__":
    main()
