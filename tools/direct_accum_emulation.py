"""CPU emulation (NumPy, every operation rounded to fp32) of the accumulation ORDER of the DIRECT FIR on the cfg1 signal:
how much of the distance to the fp64 oracle is the order in which the 400 taps of one output are summed.  No GPU needed.

    python tools/direct_accum_emulation.py

Measured here (8 of the 16 cfg1 tones, two 2e5-sample buffers; LO from fp64 rounded to fp32, products in fp32):
    one 400-term chain                         3.5e-7   (our fp32 kernel on the GPU: 4.2e-7, tensor-core kernel 5.1e-7)
    4 chains of 100 taps, then 3 adds          1.2e-7   (= the reference on the GPU: cuBLAS Cgemm per 100 taps + 4 Caxpy)
    25 chains of 16 taps, added in order       9.8e-8
    25 chains of 16 taps, added pairwise       5.8e-8   (the floor set by rounding the LO and the products)
So the reference's smaller error is its blocking, not its fp64 LO, and a blocked sum puts these kernels below it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import direct_param, orc, tone_stream  # noqa: E402

c64 = np.complex64


def chain(W, h):
    acc = np.zeros(W.shape[0], dtype=c64)
    for k in range(W.shape[1]):
        acc = (acc + W[:, k] * h[k]).astype(c64)
    return acc


def add_in_order(parts):
    t = parts[0]
    for q in parts[1:]:
        t = (t + q).astype(c64)
    return t


def add_pairwise(parts):
    while len(parts) > 1:
        parts = [(parts[i] + parts[i + 1]).astype(c64) if i + 1 < len(parts) else parts[i] for i in range(0, len(parts), 2)]
    return parts[0]


def main():
    p = direct_param(L=200_000)
    L, M, f, R = p.buffer_len, p.decim, p.pf_average, p.rate
    x = np.concatenate([tone_stream(R, p.freq, p.ampl, i * L, L) for i in range(2)])
    h = orc.make_sinc_window(M * f, float(np.float32(0.75 / (2 * M)))).astype(np.float32)
    n = np.arange(x.size, dtype=np.int64)
    names = ["one 400-term chain", "4 x 100 taps, in order", "25 x 16 taps, in order", "25 x 16 taps, pairwise"]
    num, den = dict.fromkeys(names, 0.0), 0.0
    for tf in p.freq[:8]:
        lo64 = np.exp(-2j * np.pi * ((tf * n) % R) / R)
        mixed32 = (x * lo64.astype(c64)).astype(c64)
        nout = (x.size - M * f) // M + 1
        idx = np.arange(nout)[:, None] * M + np.arange(M * f)[None, :]
        W32 = mixed32[idx]
        ref = (x.astype(np.complex128) * lo64)[idx] @ h.astype(np.float64)

        def blocks(B):
            return [chain(W32[:, s:s + B], h[s:s + B]) for s in range(0, M * f, B)]

        got = [chain(W32, h), add_in_order(blocks(100)), add_in_order(blocks(16)), add_pairwise(blocks(16))]
        den += float(np.sum(np.abs(ref) ** 2))
        for k, v in zip(names, got):
            num[k] += float(np.sum(np.abs(v - ref) ** 2))
    for k in names:
        print(f"{k:28s} {np.sqrt(num[k] / den):.3e}")


if __name__ == "__main__":
    main()
