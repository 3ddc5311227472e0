"""TEST / BASELINE INFRASTRUCTURE ONLY -- multi-threaded float32 NumPy/SciPy transcription of the
hot path for the *reported CPU baseline* (BASELINE.md 2b).  The reference has no native CPU DSP
path (its only CPU DSP is the 10-line NumPy DDC in scripts/raw_data_analisys.py:56-66), so this is
a "port": the same chains as oracle/gsdr_oracle.py, but complex64 and with scipy.fft workers so it
uses the host cores the way a competent CPU implementation would.  Checked against the fp64
oracle in tests/test_oracle.py.  Never imported by the product.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.fft

from . import gsdr_oracle as orc


def cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class PFBPort:
    """TONES chain: frames = x.reshape(-1, N); P-tap weighted sum; scipy.fft (all cores); bin gather."""

    def __init__(self, rate, fft_tones, pf_average, buffer_len, freq):
        self.N, self.P, self.L = int(fft_tones), int(pf_average), int(buffer_len)
        self.window = orc.make_sinc_window(self.N * self.P, float(np.float32(1.0 / (2 * self.N)))).reshape(self.P, self.N)
        self.bins = orc.tone_bins(rate, self.N, freq) % self.N
        self.helper = orc.BufferHelper(self.N, self.L, self.P, len(freq))
        self.raw = np.zeros(self.N * orc.pfb_batching(self.L, self.N, self.P), dtype=np.complex64)
        self.workers = cores()

    def process(self, x):
        h, N, P = self.helper, self.N, self.P
        self.raw[h.new_0:h.new_0 + self.L] = x
        cb = h.current_batch
        rows = self.raw[: (cb + P - 1) * N].reshape(cb + P - 1, N)
        z = rows[0:cb] * self.window[0]
        for i in range(1, P):
            z += rows[i:i + cb] * self.window[i]
        spec = scipy.fft.fft(z, axis=1, workers=self.workers)
        out = spec[:, self.bins].reshape(-1)
        self.raw[: h.spare_samples] = self.raw[h.spare_begin:h.spare_begin + h.spare_samples].copy()
        h.update()
        return out


class DirectPort:
    """DIRECT chain: x * exp(-2 pi j ((f n) mod R)/R), (nb, M) @ (M, f) polyphase FIR, overlap-add."""

    def __init__(self, rate, freq, decim, pf_average, buffer_len):
        self.o = orc.DirectDemodulator(rate, freq, decim, pf_average, buffer_len)

    def process(self, x):
        o = self.o
        n = (np.arange(o.L, dtype=np.int64) + o.index) % o.R
        out = np.empty((o.T, o.nb), dtype=np.complex64)
        taps = o.taps32.reshape(o.f, o.M)
        for c, tf in enumerate(o.freq):
            ph = np.fmod(tf * n, o.R).astype(np.float32) * np.float32(-2.0 * np.pi / o.R)
            d = (x * (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)).reshape(o.nb, o.M)
            trapz = d @ taps.T
            for i in range(o.f):
                o.dout[c, o.f - 1 - i:o.f - 1 - i + o.nb] += trapz[:, i]
            out[c] = o.dout[c, : o.nb]
            tail = o.dout[c, o.nb:o.nb + o.f - 1].copy()
            o.dout[c, : o.f - 1] = tail
            o.dout[c, o.f - 1:o.f - 1 + o.nb] = 0
        o.index = (o.index + o.L) % o.R
        return out.T.reshape(-1)
