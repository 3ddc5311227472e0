#!/usr/bin/env python
"""Small tensor-core-kernel cases for compute-sanitizer (memcheck): DIRECT shapes with history rows, tone-group tails,
segments, the register (non-TMA) operand path, and the filter-bank-as-GEMM form with a ragged carry-over."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("GSDR_DIRECT_VARIANT", "tc")   # GSDR_DIRECT_VARIANT=i8 / GSDR_PFB_VARIANT=i8: the integer kernel
from common import direct_param, orc, pfb_param, rx_run, tone_stream  # noqa: E402


def direct(T, decim, f, L, rate, nbuf=2):
    p = direct_param(rate=rate, T=T, decim=decim, f=f, L=L)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * L, L) for i in range(nbuf)]
    ours = rx_run(p, bufs)
    o = orc.DirectDemodulator(p.rate, p.freq, p.decim, p.pf_average, L)
    return max(orc.rel_l2(a, o.process(x)) for a, x in zip(ours, bufs))


def pfb(N, P, T, L, rate, nbuf=3):
    p = pfb_param(rate=rate, N=N, P=P, T=T, L=L)
    bufs = [tone_stream(rate, p.freq, p.ampl, i * L, L) for i in range(nbuf)]
    ours = rx_run(p, bufs)
    o = orc.PFBDemodulator(rate, N, P, L, p.freq)
    return max(orc.rel_l2(a, o.process(x)) for a, x in zip(ours, bufs))


if __name__ == "__main__":
    print("direct 16x100x4", direct(16, 100, 4, 40_000, 10_000_000), flush=True)
    print("direct 5x10x8", direct(5, 10, 8, 20_000, 1_000_000), flush=True)
    print("direct 33x50x2", direct(33, 50, 2, 20_000, 10_000_000), flush=True)
    print("direct 3x13x1 (odd decim: register path)", direct(3, 13, 1, 26_000, 1_000_000), flush=True)
    print("direct 2x300x4 (3 segments)", direct(2, 300, 4, 60_000, 10_000_000), flush=True)
    print("pfb 100x4x16", pfb(100, 4, 16, 30_000, 10_000_000), flush=True)
    print("pfb 1000x2x40", pfb(1000, 2, 40, 60_000, 100_000_000), flush=True)
