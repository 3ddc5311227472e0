"""Shared helpers for the parity tests: seeded synthetic IQ, parameter builders for
the five BASELINE.json configurations (scaled where stated), and loaders for the oracle and the
reference library."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import gpu_sdr_b200 as g  # noqa: E402
from oracle import gsdr_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libgsdr_ref.so")
SEED = 1337  # also the reference's only RNG seed (cpp/kernels.cu:319)
TOL = 1e-5   # BASELINE.json north_star: relative L2 <= 1e-5 vs the fp64 transcription


def has_gpu() -> bool:
    try:
        return g.load().gsdr_device_count() > 0
    except Exception:
        return False


def quantize_iq(x: np.ndarray) -> np.ndarray:
    """Round to the sc16 grid the USRP wire format has (k/32768): exact in float32 and compact
    to store (int16 pairs) in the golden fixtures."""
    re = np.clip(np.round(x.real * 32768.0), -32768, 32767)
    im = np.clip(np.round(x.imag * 32768.0), -32768, 32767)
    return ((re + 1j * im) / 32768.0).astype(np.complex64)


def iq_to_int16(x: np.ndarray) -> np.ndarray:
    out = np.empty((x.size, 2), dtype=np.int16)
    out[:, 0] = np.round(x.real * 32768.0)
    out[:, 1] = np.round(x.imag * 32768.0)
    return out


def int16_to_iq(a: np.ndarray) -> np.ndarray:
    return ((a[:, 0].astype(np.float32) + 1j * a[:, 1].astype(np.float32)) / np.float32(32768.0)).astype(np.complex64)


def tone_stream(rate, freq, ampl, n0, n, noise=1e-3, seed=SEED):
    """x[n] = sum_t a_t exp(2 pi j f_t n / rate) + complex gaussian noise, samples n0..n0+n."""
    idx = np.arange(n0, n0 + n, dtype=np.int64)
    x = np.zeros(n, dtype=np.complex128)
    for f, a in zip(freq, ampl):
        ph = (int(f) * idx) % int(rate)
        x += a * np.exp(2j * np.pi * ph / rate)
    rng = np.random.default_rng([seed, n0])
    x += noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return quantize_iq(x)


def pfb_param(rate=200_000_000, N=2048, P=4, T=1000, L=1_000_000, seed=SEED):
    """cfg2: T distinct tones on bin centres k*rate/N, int()-truncated like the client does
    (pyUSRP/USRP_files.py:614), so the reference's ceil-style bin rule is exercised."""
    rng = np.random.default_rng(seed)
    ks = rng.choice(np.arange(-N // 2 + 1, N // 2), size=min(T, N - 1), replace=False)
    freq = [int(k * (rate / N)) for k in ks]
    p = g.param(mode="RX", rate=rate, fft_tones=N, pf_average=P, buffer_len=L, decim=0,
                freq=freq, wave_type=[g.TONES] * len(freq), ampl=[1.0 / len(freq)] * len(freq))
    return p


def direct_param(rate=100_000_000, T=16, decim=100, f=4, L=1_000_000, seed=SEED):
    """cfg1: T distinct integer tones in (-rate/2, rate/2), at least a quarter negative."""
    rng = np.random.default_rng(seed + 1)
    freq = []
    while len(freq) < T:
        v = int(rng.integers(-rate // 2 + 1, rate // 2))
        if len(freq) < max(T // 4, 1):
            v = -abs(v) - 1
        if v not in freq:
            freq.append(v)
    return g.param(mode="RX", rate=rate, decim=decim, pf_average=f, buffer_len=L, freq=freq,
                   wave_type=[g.DIRECT] * T, ampl=[1.0 / T] * T, data_mem_mult=max(int(np.ceil(T / max(decim, 1))), 1))


def chirp_param(rate=200_000_000, f0=-50_000_000, f1=50_000_000, steps=100_000, t=1.0, decim=1, L=1_000_000, ampl=0.5):
    """cfg3 (get_VNA workload)."""
    return g.param(mode="RX", rate=rate, decim=decim, buffer_len=L, freq=[f0], chirp_f=[f1], swipe_s=[steps],
                   chirp_t=[t], wave_type=[g.CHIRP], ampl=[ampl])


def c_param(p):
    return p.to_c()


# ---- reference library (oracle/_ref) -------------------------------------------------------------
_ref = None


def ref_lib():
    """The reference's own object code behind oracle/ref_harness.cu, or None when not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        lib = C.CDLL(REF_SO)
        lib.gsdr_ref_rx_create.restype = C.c_void_p
        lib.gsdr_ref_rx_create.argtypes = [C.c_void_p]
        lib.gsdr_ref_rx_process.restype = C.c_int
        lib.gsdr_ref_rx_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gsdr_ref_rx_process_timed.restype = C.c_double
        lib.gsdr_ref_rx_process_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        lib.gsdr_ref_rx_multi_process_timed.restype = C.c_double
        lib.gsdr_ref_rx_multi_process_timed.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                                        C.POINTER(C.c_void_p), C.c_int]
        lib.gsdr_ref_rx_process_split_timed.restype = C.c_int
        lib.gsdr_ref_rx_process_split_timed.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_size_t,
                                                        C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.gsdr_ref_rx_close.argtypes = [C.c_void_p]
        lib.gsdr_ref_rx_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.gsdr_ref_rx_batching.argtypes = [C.c_void_p]
        lib.gsdr_ref_rx_window.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.gsdr_ref_rx_chirp_param.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsdr_ref_tx_create.restype = C.c_void_p
        lib.gsdr_ref_tx_create.argtypes = [C.c_void_p]
        lib.gsdr_ref_tx_get.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsdr_ref_tx_chirp_param.argtypes = [C.c_void_p, C.c_void_p]
        lib.gsdr_ref_tx_close.argtypes = [C.c_void_p]
        lib.gsdr_ref_make_sinc_window.argtypes = [C.c_int, C.c_float, C.c_void_p]
        lib.gsdr_ref_make_flat_window.argtypes = [C.c_int, C.c_int, C.c_void_p]
        lib.gsdr_ref_buffer_helper_seq.argtypes = [C.c_int] * 5 + [C.c_void_p]
        lib.gsdr_ref_vna_helper_seq.argtypes = [C.c_int] * 3 + [C.c_void_p]
        lib.gsdr_ref_host_alloc.restype = C.c_void_p
        lib.gsdr_ref_host_alloc.argtypes = [C.c_size_t]
        lib.gsdr_ref_host_free.argtypes = [C.c_void_p]
        _ref = lib
    return _ref


class RefRX:
    """Drives the reference's RX_buffer_demodulator (unmodified) through the harness."""

    def __init__(self, p):
        self.lib = ref_lib()
        self.cp, self.keep = p.to_c()
        self.h = self.lib.gsdr_ref_rx_create(C.byref(self.cp))
        self.p = p

    def process(self, x: np.ndarray, cap: int) -> np.ndarray:
        xin = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros(cap, dtype=np.complex64)
        n = self.lib.gsdr_ref_rx_process(self.h, xin.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def bins(self, T):
        b = np.empty(T, dtype=np.int32)
        self.lib.gsdr_ref_rx_bins(self.h, b.ctypes.data_as(C.c_void_p), T)
        return b

    def close(self):
        if self.h:
            self.lib.gsdr_ref_rx_close(self.h)
            self.h = None


class RefTX:
    def __init__(self, p):
        self.lib = ref_lib()
        self.cp, self.keep = p.to_c()
        self.h = self.lib.gsdr_ref_tx_create(C.byref(self.cp))
        self.L = int(p.buffer_len)

    def get(self) -> np.ndarray:
        out = np.zeros(self.L, dtype=np.complex64)
        self.lib.gsdr_ref_tx_get(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def close(self):
        if self.h:
            self.lib.gsdr_ref_tx_close(self.h)
            self.h = None


def rx_run(p, buffers, device=0):
    """Run our RX_buffer_demodulator over a list of input buffers; returns list of outputs."""
    rx = g.RX_buffer_demodulator(p, device=device)
    out = g.pinned_empty(rx.max_output())
    res = []
    try:
        for x in buffers:
            xin = np.ascontiguousarray(x, dtype=np.complex64)
            n = rx.process(xin, out)
            res.append(out[:n].copy())
    finally:
        rx.close()
        g.pinned_free(out)
    return res
