"""CPU emulation of the arithmetic of direct_fir_tc_kernel on the cfg1 signal (no GPU): which term of the tensor-core
scheme sets its distance to the fp64 oracle, and what each candidate change buys.  Model of tcgen05.mma kind::tf32:
operands truncated to TF32 (10 explicit mantissa bits, low 13 bits dropped), products exact, the 8 products of a k-step
summed exactly and added to the fp32 accumulator with truncation toward zero (measured on B200: ~1.6e-7 relative per
accumulated k-step for same-sign terms).  Everything outside the MMA (splits, folds, the F-term sum, the LO rotation) in
fp32 round-to-nearest, as the kernel does it.

    python tools/direct_tc_emulation.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import direct_param, orc, tone_stream  # noqa: E402

f32 = np.float32


def tf32_rn(x):
    b = x.astype(f32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(f32)


def tf32_trunc(x):
    return (x.astype(f32).view(np.uint32) & np.uint32(0xffffe000)).view(f32)


def trunc32(x64):
    """float64 -> float32 toward zero."""
    y = x64.astype(f32)
    over = np.abs(y.astype(np.float64)) > np.abs(x64)
    y[over] = np.nextafter(y[over], f32(0))
    return y


def mma_chain(A_parts, B_parts, ksteps, trunc_acc=True):
    """sum over listed (A, B) operand pairs of A @ B, accumulated k-step by k-step (8 reals per k-step).
    A: [rows, K] float32 (already TF32-representable or truncated here), B: [K, N]."""
    rows, K = A_parts[0][0].shape
    N = A_parts[0][1].shape[1]
    acc = np.zeros((rows, N), dtype=f32)
    for k0 in range(0, K, 8):
        s = np.zeros((rows, N), dtype=np.float64)
        for A, B in A_parts:
            s += tf32_trunc(A[:, k0:k0 + 8]).astype(np.float64) @ tf32_trunc(B[k0:k0 + 8]).astype(np.float64)
        t = acc.astype(np.float64) + s
        acc = trunc32(t) if trunc_acc else t.astype(f32)
    return acc


def run(variant, x, p, tones):
    L, M, F, R = p.buffer_len, p.decim, p.pf_average, p.rate
    h = orc.make_sinc_window(M * F, float(np.float32(0.75 / (2 * M)))).astype(np.float64)
    nrows = x.size // M
    A = np.empty((nrows, 2 * M), dtype=f32)
    A[:, 0::2] = x.real.reshape(nrows, M)
    A[:, 1::2] = x.imag.reshape(nrows, M)
    n_out = nrows - F + 1
    num = den = 0.0
    for tf in tones:
        m = np.arange(M * F, dtype=np.int64)
        g = h * np.exp(-2j * np.pi * ((tf * m) % R) / R)
        g = (g.real.astype(f32) + 1j * g.imag.astype(f32))
        # exact reference from the same fp32 operands' fp64 values would hide the operand rounding: use the true chain
        n = np.arange(x.size, dtype=np.int64)
        lo64 = np.exp(-2j * np.pi * ((tf * n) % R) / R)
        idx = np.arange(n_out)[:, None] * M + m[None, :]
        ref = (x.astype(np.complex128) * lo64)[idx] @ h
        Z = np.zeros((nrows, F), dtype=np.complex64)
        seg_k = variant["seg_kblocks"] * 32   # reals per accumulation segment
        for i in range(F):
            gi = g[i * M:(i + 1) * M]
            B = np.empty((2 * M, 2), dtype=f32)     # columns: Re, Im of the complex product
            B[0::2, 0] = gi.real
            B[1::2, 0] = -gi.imag
            B[0::2, 1] = gi.imag
            B[1::2, 1] = gi.real
            Ahi, Bhi = tf32_rn(A), tf32_rn(B)
            Alo, Blo = (A - Ahi).astype(f32), (B - Bhi).astype(f32)
            if variant.get("round_lo"):
                Alo, Blo = tf32_rn(Alo), tf32_rn(Blo)
            tot = None
            for s0 in range(0, 2 * M, seg_k):
                sl = slice(s0, min(2 * M, s0 + seg_k))
                main = mma_chain([(Ahi[:, sl], Bhi[sl])], None, None)
                pairs = [(Alo[:, sl], Bhi[sl]), (Ahi[:, sl], Blo[sl])]
                if variant.get("lolo"):
                    pairs.append((Alo[:, sl], Blo[sl]))
                corr = mma_chain(pairs, None, None)
                part = (main + corr).astype(f32)
                tot = part if tot is None else (tot + part).astype(f32)
            Z[:, i] = tot[:, 0] + 1j * tot[:, 1]
        y = Z[0:n_out, 0]
        for i in range(1, F):
            y = (y + Z[i:i + n_out, i]).astype(np.complex64)
        n0 = (np.arange(n_out, dtype=np.int64) * M) % R
        rot = np.exp(-2j * np.pi * ((tf * n0) % R) / R).astype(np.complex64)
        y = (y * rot).astype(np.complex64)
        num += float(np.sum(np.abs(y - ref) ** 2))
        den += float(np.sum(np.abs(ref) ** 2))
    return np.sqrt(num / den)


def main():
    p = direct_param(L=100_000)
    x = np.concatenate([tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(2)])
    tones = p.freq[:4]
    variants = [
        ("as built: seg 8 K blocks, lo truncated by the MMA, lo*lo dropped", dict(seg_kblocks=8)),
        ("seg 1", dict(seg_kblocks=1)),
        ("seg 8 + lo rounded to TF32 before the MMA", dict(seg_kblocks=8, round_lo=True)),
        ("seg 2 + lo rounded", dict(seg_kblocks=2, round_lo=True)),
        ("seg 1 + lo rounded", dict(seg_kblocks=1, round_lo=True)),
        ("seg 1 + lo rounded + lo*lo", dict(seg_kblocks=1, round_lo=True, lolo=True)),
        ("seg 8 + lo rounded + lo*lo", dict(seg_kblocks=8, round_lo=True, lolo=True)),
    ]
    for name, v in variants:
        print(f"{name:70s} {run(v, x, p, tones):.3e}", flush=True)


if __name__ == "__main__":
    main()
