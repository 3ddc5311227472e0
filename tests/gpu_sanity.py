"""Quick on-GPU sanity run (not a pytest file): ours vs the fp64 oracle vs the reference library for
every mode, printing relative L2 errors.  `python tests/gpu_sanity.py [--ref]`."""
import sys
import time

import numpy as np

from common import (TOL, RefRX, RefTX, chirp_param, direct_param, g, orc, pfb_param, ref_lib, rx_run, tone_stream)

use_ref = "--ref" in sys.argv and ref_lib() is not None
worst = 0.0


def report(name, a, b, tol=TOL):
    global worst
    if len(a) != len(b):
        print(f"{name}: LENGTH MISMATCH {len(a)} vs {len(b)}")
        worst = 1.0
        return
    e = orc.rel_l2(a, b)
    worst = max(worst, e if np.isfinite(e) else 1.0)
    print(f"{name}: n={len(a)} rel_l2={e:.3e} {'OK' if e <= tol else 'FAIL'}")


def run_pfb(N, P, T, L, nbuf, rate=200_000_000):
    p = pfb_param(rate=rate, N=N, P=P, T=T, L=L)
    bufs = [tone_stream(rate, p.freq, p.ampl, i * L, L) for i in range(nbuf)]
    t0 = time.time()
    ours = rx_run(p, bufs)
    t1 = time.time()
    o = orc.PFBDemodulator(rate, N, P, L, p.freq)
    ref = RefRX(p) if use_ref else None
    for i, x in enumerate(bufs):
        want = o.process(x)
        report(f"pfb N={N} P={P} T={len(p.freq)} buf{i} ours-vs-oracle", ours[i], want)
        if ref:
            r = ref.process(x, len(p.freq) * o.batching)
            report(f"pfb N={N} P={P} buf{i} ref-vs-oracle", r, want)
    if ref:
        ref.close()
    print(f"  (ours {t1 - t0:.2f}s)")


def run_direct(T, decim, f, L, nbuf, rate=100_000_000):
    p = direct_param(rate=rate, T=T, decim=decim, f=f, L=L)
    bufs = [tone_stream(rate, p.freq, p.ampl, i * L, L) for i in range(nbuf)]
    ours = rx_run(p, bufs)
    o = orc.DirectDemodulator(rate, p.freq, decim, f, L)
    ref = RefRX(p) if use_ref else None
    for i, x in enumerate(bufs):
        want = o.process(x)
        report(f"direct T={T} decim={decim} buf{i} ours-vs-oracle", ours[i], want)
        if ref:
            r = ref.process(x, len(want) + 16)
            sk = (f - 1) * T if (i == 0 and decim > 0) else 0  # reference FIR tail starts uninitialised
            report(f"direct T={T} decim={decim} buf{i} ref-vs-oracle", r[sk:], want[sk:])
    if ref:
        ref.close()


def run_chirp(decim, L, nbuf, steps=100_000, t=1.0, rate=200_000_000):
    p = chirp_param(rate=rate, steps=steps, t=t, decim=decim, L=L)
    gen = orc.ChirpGenerator(rate, p.freq[0], p.chirp_f[0], steps, t, 1.0, L)
    rng = np.random.default_rng(5)
    bufs = []
    for i in range(nbuf):
        s21 = 0.5 * np.exp(2j * np.pi * 0.1 * i)
        bufs.append((gen.get() * s21 + 1e-3 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))).astype(np.complex64))
    ours = rx_run(p, bufs)
    o = orc.ChirpDemodulator(rate, p.freq[0], p.chirp_f[0], steps, t, decim, L)
    ref = RefRX(p) if use_ref else None
    for i, x in enumerate(bufs):
        want = o.process(x)
        report(f"chirp decim={decim} steps={steps} buf{i} ours-vs-oracle", ours[i], want)
        if ref:
            r = ref.process(x, L)
            report(f"chirp decim={decim} buf{i} ref-vs-oracle", r, want)
    if ref:
        ref.close()


def run_tx():
    rate, L = 1_000_000, 50_000
    freq = [1000, -2500, 333_333, -499_999, 77]
    ampl = [0.2, 0.1, 0.3, 0.15, 0.05]
    p = g.param(mode="TX", rate=rate, buffer_len=L, freq=freq, ampl=ampl, wave_type=[g.TONES] * 5)
    tx = g.TX_buffer_generator(p)
    o = orc.ToneGenerator(rate, freq, ampl, L)
    ref = RefTX(p) if use_ref else None
    for i in range(3):
        a = tx.get().copy()
        want = o.get()
        report(f"tx tones buf{i} ours-vs-oracle", a, want)
        if ref:
            report(f"tx tones buf{i} ref-vs-oracle", ref.get(), want)
    tx.close()
    if ref:
        ref.close()
    pc = chirp_param(rate=200_000_000, steps=1000, t=0.01, L=100_000, ampl=0.7)
    pc.mode = "TX"
    tx = g.TX_buffer_generator(pc)
    o = orc.ChirpGenerator(pc.rate, pc.freq[0], pc.chirp_f[0], 1000, 0.01, 0.7, 100_000)
    ref = RefTX(pc) if use_ref else None
    buf = g.pinned_empty(100_000)
    for i in range(3):
        a = tx.get(buf).copy()
        want = o.get()
        report(f"tx chirp buf{i} ours-vs-oracle", a, want)
        if ref:
            report(f"tx chirp buf{i} ref-vs-oracle", ref.get(), want)
    tx.close()
    if ref:
        ref.close()


def run_probes():
    global worst
    p = orc.chirp_params(200_000_000, -50_000_000, 50_000_000, 100_000, 1.0)
    for last in (0, 123_456_789, 199_999_000):
        a = g.hostlogic.probe_chirp_index(g.hostlogic.chirp_params(200_000_000, -50_000_000, 50_000_000, 100_000, 1.0), last, 100_000)
        b = orc.chirp_index(last, 100_000, p)
        ok = np.array_equal(a, b)
        worst = max(worst, 0.0 if ok else 1.0)
        print(f"chirp index probe last={last}: {'bit-exact' if ok else 'MISMATCH'}")
    for tf in (12_345_677, -49_999_999, 1, -1):
        a = g.hostlogic.probe_direct_phase(tf, 100_000_000, 99_000_000, 5, 200_000)
        b = orc.direct_phase(tf, 0, 100_000_000, 99_000_000, 5, 200_000)
        ok = np.array_equal(a, b)
        worst = max(worst, 0.0 if ok else 1.0)
        print(f"direct phase probe tf={tf}: {'bit-exact' if ok else 'MISMATCH'}")


if __name__ == "__main__":
    print("devices:", g.load().gsdr_device_count(), "ref:", use_ref)
    run_probes()
    run_pfb(2048, 4, 1000, 1_000_000, 3)
    run_pfb(2048, 2, 17, 100_000, 3)
    run_pfb(100, 3, 7, 50_000, 3, rate=1_000_000)
    run_pfb(64, 4, 8, 20_000, 3, rate=1_000_000)
    run_direct(16, 100, 4, 1_000_000, 3)
    run_direct(3, 0, 1, 50_000, 2)
    run_direct(5, 10, 8, 50_000, 3, rate=1_000_000)
    run_chirp(1, 1_000_000, 3)
    run_chirp(0, 100_000, 2)
    run_chirp(3, 100_000, 4, steps=1000, t=0.01)
    run_chirp(200, 100_000, 3, steps=0, t=0.001)
    run_tx()
    print("WORST", worst)
    sys.exit(0 if worst <= TOL else 1)
