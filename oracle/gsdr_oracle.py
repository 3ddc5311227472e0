"""TEST INFRASTRUCTURE ONLY -- fp64 NumPy restatement ("oracle") of GPU_SDR's RX/TX DSP chains.

Never imported by the product package ``gpu_sdr_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / reported baseline.

The integer bookkeeping and the float32 tap builders live in ``gsdr_oracle.c`` (plain C, same
libm and C conversion semantics as the reference's host code); this file restates the streaming
floating-point chains on top of them in complex128 ("fp64 transcription", the tolerance
reference named by BASELINE.json: relative L2 <= 1e-5).  Paths cited are relative to
``/root/reference``.

Parity pin: checked against the reference's own object code (``oracle/_ref/libgsdr_ref.so``,
the unmodified reference sources compiled for sm_100a, run on a B200) through the committed
fixtures under ``tests/golden/`` (made by ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgsdr_oracle.so")


def build(force: bool = False) -> str:
    """Compile gsdr_oracle.c with gcc (no fast-math, no FMA contraction: the reference's taps
    are built by nvcc's default host flags, cpp/kernels.cu is not compiled with -march=native)."""
    src = os.path.join(_HERE, "gsdr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", _SO, src, "-lm"]
        )
    return _SO


_lib = None


class _BH(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "n_tones", "eff_length", "buffer_len", "average", "n_eff_tones",
        "new_0", "copy_size", "current_batch", "spare_samples", "spare_begin")]


class _VH(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("valid_size", "new0", "total_len", "spare_begin", "ppt", "buffer_len")]


class ChirpParam(C.Structure):
    _fields_ = [("num_steps", C.c_uint64), ("length", C.c_uint64), ("chirpness", C.c_uint32), ("f0", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_pfb_batching.restype = C.c_int
    return _lib


# ----------------------------------------------------------------------------------------------
# taps (float32-exact) and integer helpers
# ----------------------------------------------------------------------------------------------
def make_sinc_window(length: int, fc: float) -> np.ndarray:
    """cpp/kernels.cu:258-310. Returns the float32 real taps (imag part is identically 0)."""
    out = np.empty(length, dtype=np.float32)
    lib().orc_make_sinc_window(C.c_int(length), C.c_float(fc), out.ctypes.data_as(C.c_void_p))
    return out


def make_flat_window(length: int, side: int) -> np.ndarray:
    """cpp/kernels.cu:208-253."""
    out = np.empty(length, dtype=np.float32)
    lib().orc_make_flat_window(C.c_int(length), C.c_int(side), out.ctypes.data_as(C.c_void_p))
    return out


class BufferHelper:
    """cpp/USRP_server_memory_management.cpp:104-156."""

    FIELDS = ("eff_length", "new_0", "copy_size", "current_batch", "spare_samples", "spare_begin")

    def __init__(self, n_tones, buffer_len, average, n_eff_tones):
        self._h = _BH()
        lib().orc_buffer_helper_init(C.byref(self._h), n_tones, buffer_len, average, n_eff_tones)

    def update(self):
        lib().orc_buffer_helper_update(C.byref(self._h))

    def __getattr__(self, k):
        return getattr(self._h, k)

    def state(self):
        return tuple(getattr(self._h, k) for k in self.FIELDS)


class VNAHelper:
    """cpp/USRP_server_memory_management.cpp:30-56."""

    FIELDS = ("valid_size", "new0", "total_len", "spare_begin")

    def __init__(self, ppt, buffer_len):
        self._h = _VH()
        lib().orc_vna_helper_init(C.byref(self._h), ppt, buffer_len)

    def update(self):
        lib().orc_vna_helper_update(C.byref(self._h))

    def __getattr__(self, k):
        return getattr(self._h, k)

    def state(self):
        return tuple(getattr(self._h, k) for k in self.FIELDS)


def pfb_batching(buffer_len, fft_tones, pf_average) -> int:
    """cpp/USRP_demodulator.cpp:706."""
    return int(lib().orc_pfb_batching(buffer_len, fft_tones, pf_average))


def tone_bins(rate, fft_tones, freq) -> np.ndarray:
    """cpp/USRP_demodulator.cpp:722-734 (last match wins; -1 = never matched)."""
    f = np.ascontiguousarray(freq, dtype=np.int32)
    out = np.empty(len(f), dtype=np.int32)
    lib().orc_tone_bins(int(rate), int(fft_tones), f.ctypes.data_as(C.c_void_p), len(f), out.ctypes.data_as(C.c_void_p))
    return out


def chirp_params(rate, freq0, chirp_f0, swipe_s0, chirp_t0, tx=False) -> ChirpParam:
    """cpp/USRP_demodulator.cpp:192-214 / cpp/USRP_buffer_generator.cpp:114-137."""
    p = ChirpParam()
    lib().orc_chirp_params(int(rate), int(freq0), int(chirp_f0), int(swipe_s0), C.c_float(chirp_t0), int(bool(tx)), C.byref(p))
    return p


def chirp_index(last_index, n, p: ChirpParam) -> np.ndarray:
    """cpp/kernels.cu:401-419: the int32 phase index of each sample."""
    out = np.empty(n, dtype=np.int32)
    lib().orc_chirp_index(C.c_uint64(last_index), C.c_uint32(n), C.byref(p), out.ctypes.data_as(C.c_void_p))
    return out


def direct_phase(tf, tp, wavetablelen, index_counter, n0, n) -> np.ndarray:
    """cpp/kernels.cu:63-68: signed int64 LO phase (in units of 1/wavetablelen turns)."""
    out = np.empty(n, dtype=np.int64)
    lib().orc_direct_phase(int(tf), int(tp), int(wavetablelen), C.c_uint64(index_counter), C.c_uint64(n0), C.c_uint64(n),
                           out.ctypes.data_as(C.c_void_p))
    return out


# ----------------------------------------------------------------------------------------------
# RX: TONES (polyphase filter bank), cpp/USRP_demodulator.cpp:486-565
# ----------------------------------------------------------------------------------------------
class PFBDemodulator:
    """process_pfb restated: upload at new_0, P-tap polyphase front end (cpp/kernels.cu:474-516),
    N-point unnormalised forward FFT per frame (cufftExecC2C FORWARD, USRP_demodulator.cpp:501),
    tone_select (cpp/kernels.cu:531-554), carry-over via buffer_helper + move_buffer."""

    def __init__(self, rate, fft_tones, pf_average, buffer_len, freq):
        self.N, self.P, self.L = int(fft_tones), int(pf_average), int(buffer_len)
        self.T = len(freq)
        fcut = np.float32(1.0 / (2 * self.N))  # USRP_demodulator.cpp:131 (double -> float member)
        self.window32 = make_sinc_window(self.N * self.P, float(fcut))
        self.window = self.window32.astype(np.float64).reshape(self.P, self.N)
        self.bins = tone_bins(rate, self.N, freq)
        self.batching = pfb_batching(self.L, self.N, self.P)
        self.helper = BufferHelper(self.N, self.L, self.P, self.T)
        self.raw = np.zeros(self.N * self.batching, dtype=np.complex128)

    def process(self, x):
        h, N, P = self.helper, self.N, self.P
        self.raw[h.new_0:h.new_0 + self.L] = np.asarray(x, dtype=np.complex128)
        cb = h.current_batch
        rows = self.raw[: (cb + P - 1) * N].reshape(cb + P - 1, N)
        z = np.zeros((cb, N), dtype=np.complex128)
        for i in range(P):  # y[b,k] = sum_i x[(b+i)N+k] * w[iN+k]
            z += rows[i:i + cb] * self.window[i]
        spec = np.fft.fft(z, axis=1)
        out = spec[:, self.bins % N].reshape(-1)  # out[b*T+u]
        # move_buffer: spare samples to the head (USRP_demodulator.cpp:504-509)
        self.raw[: h.spare_samples] = self.raw[h.spare_begin:h.spare_begin + h.spare_samples].copy()
        h.update()
        return out


# ----------------------------------------------------------------------------------------------
# RX: DIRECT (DDC + decimating FIR), cpp/USRP_demodulator.cpp:400-464
# ----------------------------------------------------------------------------------------------
class DirectDemodulator:
    """direct_demodulator_integer (cpp/kernels.cu:45-86) + FIR::run_fir per tone (cpp/fir.cu:44-88)
    + transpose to sample-major (USRP_demodulator.cpp:422-455).  FIR tail starts at zero (the
    reference's cudaMemset(&_dout...) at fir.cu:26 is a no-op on a fresh cudaMalloc)."""

    def __init__(self, rate, freq, decim, pf_average, buffer_len):
        self.R, self.M, self.f, self.L = int(rate), int(decim), int(pf_average), int(buffer_len)
        self.freq = [int(v) for v in freq]
        self.T = len(self.freq)
        self.index = 0
        if self.M > 0:
            assert self.L % self.M == 0  # fir.cu:20
            fc = np.float32(0.75 / (self.M * 2))  # USRP_demodulator.cpp:99 (double -> float arg)
            self.taps32 = make_sinc_window(self.M * self.f, float(fc))
            self.taps = self.taps32.astype(np.float64).reshape(self.f, self.M)
            self.nb = self.L // self.M
            self.dout = np.zeros((self.T, self.nb + self.f - 1), dtype=np.complex128)

    def mix(self, x):
        x = np.asarray(x, dtype=np.complex128)
        d = np.empty((self.T, self.L), dtype=np.complex128)
        for c, tf in enumerate(self.freq):
            ph = direct_phase(tf, 0, self.R, self.index, 0, self.L)
            d[c] = x * np.exp(-2j * np.pi * (ph.astype(np.float64) / self.R))
        return d

    def process(self, x):
        d = self.mix(x)
        self.index = (self.index + self.L) % self.R  # USRP_demodulator.cpp:437-440
        if self.M <= 0:
            return d.T.reshape(-1)  # out[n*T+ch]
        out = np.empty((self.T, self.nb), dtype=np.complex128)
        for c in range(self.T):
            trapz = d[c].reshape(self.nb, self.M) @ self.taps.T  # [b, j] = sum_k d[bM+k] h[jM+k]
            for i in range(self.f):  # dout[f-1-i+b] += trapz[i][b]
                self.dout[c, self.f - 1 - i:self.f - 1 - i + self.nb] += trapz[:, i]
            out[c] = self.dout[c, : self.nb]
            rem = self.f - 1
            tail = self.dout[c, self.nb:self.nb + rem].copy()
            self.dout[c, :rem] = tail
            self.dout[c, rem:rem + self.nb] = 0
        return out.T.reshape(-1)  # out[p*T+ch]


# ----------------------------------------------------------------------------------------------
# RX: CHIRP (VNA lock-in), cpp/USRP_demodulator.cpp:342-397
# ----------------------------------------------------------------------------------------------
def chirp_phasor(index):
    """(sin(pi*theta), -cos(pi*theta)) with theta = index / 2147483647.5 (cpp/kernels.cu:421-422)."""
    th = index.astype(np.float64) / 2147483647.5
    return np.sin(np.pi * th) - 1j * np.cos(np.pi * th)


class ChirpDemodulator:
    def __init__(self, rate, freq0, chirp_f0, swipe_s0, chirp_t0, decim, buffer_len):
        self.L = int(buffer_len)
        self.p = chirp_params(rate, freq0, chirp_f0, swipe_s0, chirp_t0, tx=False)
        self.last_index = 0
        self.decim = int(decim)
        self.spare_size = 0
        if self.decim > 0:
            self.ppt = int(self.p.length) * self.decim  # USRP_demodulator.cpp:231
            self.helper = VNAHelper(self.ppt, self.L)
            self.profile32 = make_flat_window(self.ppt, self.ppt // 10)  # :246
            self.profile = self.profile32.astype(np.float64)
            # the reference's device buffer is 3*buffer_len (USRP_demodulator.cpp:225); it overruns it when
            # ppt > 2*buffer_len.  The restatement keeps enough room so long integrations stay defined.
            self.buf = np.zeros(max(3 * self.L, self.ppt + self.L), dtype=np.complex128)

    def period(self):
        return int(self.p.num_steps) * int(self.p.length)

    def process(self, x):
        x = np.asarray(x, dtype=np.complex128)
        idx = chirp_index(self.last_index, self.L, self.p)
        dem = x * np.conj(chirp_phasor(idx))  # out = in * (sin + j cos), kernels.cu:424-425
        self.last_index = (self.last_index + self.L) % self.period()
        if self.decim <= 0:
            return dem
        h = self.helper
        self.buf[self.spare_size:self.spare_size + self.L] = dem
        valid = h.valid_size
        out = self.buf[: valid * self.ppt].reshape(valid, self.ppt) @ self.profile  # cublas_decim
        self.spare_size = h.new0
        if self.spare_size > 0:
            self.buf[: h.new0] = self.buf[h.spare_begin:h.spare_begin + h.new0].copy()
        h.update()
        return out


# ----------------------------------------------------------------------------------------------
# TX: CHIRP and TONES, cpp/USRP_buffer_generator.cpp
# ----------------------------------------------------------------------------------------------
class ChirpGenerator:
    """get_from_chirp (:208-221) + chirp_gen (cpp/kernels.cu:335-372)."""

    def __init__(self, rate, freq0, chirp_f0, swipe_s0, chirp_t0, ampl0, buffer_len):
        self.L = int(buffer_len)
        self.p = chirp_params(rate, freq0, chirp_f0, swipe_s0, chirp_t0, tx=True)
        self.scale = float(np.float32(ampl0))
        self.last_index = 0

    def get(self):
        idx = chirp_index(self.last_index, self.L, self.p)
        self.last_index = (self.last_index + self.L) % (int(self.p.num_steps) * int(self.p.length))
        return chirp_phasor(idx) * self.scale


def tone_period(rate, freq, ampl) -> np.ndarray:
    """tone_gen (cpp/kernels.cu:589-684): unnormalised inverse DFT of a spectrum with
    base[f>0 ? f : rate+f].x = ampl (assignment: the last duplicate wins; f == 0 indexes
    base[rate], one past the end, so that tone is dropped).  Evaluated in closed form."""
    R = int(rate)
    bins = {}
    for f, a in zip(freq, ampl):
        f = int(f)
        k = f if f > 0 else R + f
        if 0 <= k < R:
            bins[k] = float(np.float32(a))
    n = np.arange(R, dtype=np.int64)
    x = np.zeros(R, dtype=np.complex128)
    for k, a in bins.items():
        x += a * np.exp(2j * np.pi * (((k * n) % R).astype(np.float64) / R))
    return x


class ToneGenerator:
    """TX TONES branch (cpp/USRP_buffer_generator.cpp:60-99) + get_from_tones (:226-229)."""

    def __init__(self, rate, freq, ampl, buffer_len):
        self.L, R = int(buffer_len), int(rate)
        base = tone_period(rate, freq, ampl)
        self.period_len = R
        if self.L > R:
            ratio = int(np.ceil(np.float32(self.L) / np.float32(R)))
            self.period_len = ratio * R
            base = np.tile(base, ratio)
        self.base = np.concatenate([base, base[: self.L]])
        self.last = 0

    def get(self):
        out = self.base[self.last:self.last + self.L]
        self.last = (self.last + self.L) % self.period_len
        return out


def rel_l2(a, b) -> float:
    """||a-b||_2 / ||b||_2 over complex vectors (b = fp64 oracle)."""
    a = np.asarray(a, dtype=np.complex128).reshape(-1)
    b = np.asarray(b, dtype=np.complex128).reshape(-1)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


# ---- client-side noise spectra: the specification a GPU version will be checked against ------------
def welch_psd_real(x, fs, nperseg):
    """One-sided power spectral density of a real series by Welch's method exactly as the client asks SciPy for it
    (pyUSRP/USRP_noise.py:698-699: `signal.welch(x, nperseg=welch, fs=sampling_rate, detrend='linear', scaling='density')`),
    written out: periodic Hann window, 50 % overlap (`noverlap = nperseg // 2`), no padding (a trailing partial segment is
    dropped), a least-squares straight line removed from every segment before windowing, |rfft|^2 / (fs * sum(w^2)), doubled
    except at DC and (nperseg even) at Nyquist, mean over segments.  fp64.  Returns (freqs, psd)."""
    x = np.asarray(x, dtype=np.float64)
    nperseg = int(min(nperseg, x.size))
    step = nperseg - nperseg // 2
    n = np.arange(nperseg, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)
    nseg = (x.size - nperseg) // step + 1
    t = n - n.mean()
    tt = float(np.dot(t, t))
    acc = np.zeros(nperseg // 2 + 1, dtype=np.float64)
    for s in range(nseg):
        seg = x[s * step:s * step + nperseg]
        m = seg.mean()
        slope = float(np.dot(t, seg - m)) / tt if tt > 0 else 0.0
        X = np.fft.rfft((seg - m - slope * t) * w)
        acc += X.real ** 2 + X.imag ** 2
    psd = acc / nseg / (fs * float(np.dot(w, w)))
    if nperseg % 2:
        psd[1:] *= 2.0
    else:
        psd[1:-1] *= 2.0
    return np.fft.rfftfreq(nperseg, 1.0 / fs), psd


def spec_from_samples(samples, sampling_rate=1.0, welch=None, dbc=False, rotate=True, clip_samples=False):
    """pyUSRP/USRP_noise.py:655-703 restated: rotate the IQ plane so the mean is real and positive, optionally scale to the
    carrier (dBc) and remove it, clip `clip_samples` at both ends, then the Welch PSD of the real and of the imaginary part with
    `nperseg = int(L / welch)` (L = the unclipped length; `welch=None` -> one segment of L).  Returns
    (freqs, 10 log10 PSD(real part), 10 log10 PSD(imaginary part)) -- the order the reference returns, not the one its docstring says."""
    z = np.asarray(samples, dtype=np.complex128)
    L = z.size
    nperseg = L if welch is None else int(L / welch)
    lo, hi = (0, L) if not clip_samples else (int(clip_samples), int(L - clip_samples))
    if rotate:
        m = z.mean()
        z = z * (abs(m) / m)
    if dbc:
        z = z / z.mean()
        z = z - z.mean()
    f, re = welch_psd_real(z[lo:hi].real, sampling_rate, nperseg)
    _, im = welch_psd_real(z[lo:hi].imag, sampling_rate, nperseg)
    with np.errstate(divide="ignore"):
        return f, 10.0 * np.log10(re), 10.0 * np.log10(im)
