"""Python access to the host-side routines of libgsdr that define results (taps, carry-over
bookkeeping, tone->bin map, chirp quantisation) and to the integer-phase probes.  All of it runs
inside libgsdr.so; nothing is re-implemented in Python."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BufferHelper, ChirpParam, VnaHelper, check


def make_sinc_window(length: int, fc: float) -> np.ndarray:
    out = np.empty(int(length), dtype=np.float32)
    check(_lib.load().gsdr_make_sinc_window(int(length), float(fc), out.ctypes.data_as(C.c_void_p)), "gsdr_make_sinc_window")
    return out


def make_flat_window(length: int, side: int) -> np.ndarray:
    out = np.empty(int(length), dtype=np.float32)
    check(_lib.load().gsdr_make_flat_window(int(length), int(side), out.ctypes.data_as(C.c_void_p)), "gsdr_make_flat_window")
    return out


def pfb_batching(buffer_len, fft_tones, pf_average) -> int:
    return int(_lib.load().gsdr_pfb_batching(int(buffer_len), int(fft_tones), int(pf_average)))


def tone_bins(rate, fft_tones, freq) -> np.ndarray:
    f = np.ascontiguousarray(freq, dtype=np.int32)
    out = np.empty(len(f), dtype=np.int32)
    check(_lib.load().gsdr_tone_bins(int(rate), int(fft_tones), f.ctypes.data_as(C.c_void_p), len(f), out.ctypes.data_as(C.c_void_p)),
          "gsdr_tone_bins")
    return out


def pfb_gather_layout(bins, n_tones=None) -> np.ndarray:
    """Slot (0..15) of every one of the 2048 bins inside its shared-memory row; bins=None: all bins in order."""
    out = np.empty(2048, dtype=np.uint8)
    if bins is None:
        n, ptr = int(n_tones if n_tones is not None else 2048), None
    else:
        b = np.ascontiguousarray(bins, dtype=np.int32)
        n, ptr = len(b), b.ctypes.data_as(C.c_void_p)
    check(_lib.load().gsdr_pfb_gather_layout(ptr, n, out.ctypes.data_as(C.c_void_p)), "gsdr_pfb_gather_layout")
    return out


BH_FIELDS = ("eff_length", "new_0", "copy_size", "current_batch", "spare_samples", "spare_begin")
VH_FIELDS = ("valid_size", "new0", "total_len", "spare_begin")


def buffer_helper_sequence(n_tones, buffer_len, average, n_eff_tones, n) -> np.ndarray:
    """State after construction and after each of n-1 updates, rows of BH_FIELDS."""
    h = BufferHelper()
    lib = _lib.load()
    lib.gsdr_buffer_helper_init(C.byref(h), int(n_tones), int(buffer_len), int(average), int(n_eff_tones))
    rows = []
    for _ in range(n):
        rows.append([getattr(h, k) for k in BH_FIELDS])
        lib.gsdr_buffer_helper_update(C.byref(h))
    return np.array(rows, dtype=np.int32)


def vna_helper_sequence(ppt, buffer_len, n) -> np.ndarray:
    h = VnaHelper()
    lib = _lib.load()
    lib.gsdr_vna_helper_init(C.byref(h), int(ppt), int(buffer_len))
    rows = []
    for _ in range(n):
        rows.append([getattr(h, k) for k in VH_FIELDS])
        lib.gsdr_vna_helper_update(C.byref(h))
    return np.array(rows, dtype=np.int32)


def chirp_params(rate, freq0, chirp_f0, swipe_s0, chirp_t0, tx=False) -> ChirpParam:
    p = ChirpParam()
    check(_lib.load().gsdr_chirp_params(int(rate), int(freq0), int(chirp_f0), int(swipe_s0), float(chirp_t0), int(bool(tx)), C.byref(p)),
          "gsdr_chirp_params")
    return p


def probe_chirp_index(p: ChirpParam, last_index: int, n: int, device: int = 0) -> np.ndarray:
    out = np.empty(int(n), dtype=np.int32)
    check(_lib.load().gsdr_probe_chirp_index(int(device), C.byref(p), int(last_index), int(n), out.ctypes.data_as(C.c_void_p)),
          "gsdr_probe_chirp_index")
    return out


def probe_direct_phase(tone_freq, rate, index_counter, n0, n, device: int = 0) -> np.ndarray:
    out = np.empty(int(n), dtype=np.int64)
    check(_lib.load().gsdr_probe_direct_phase(int(device), int(tone_freq), int(rate), int(index_counter), int(n0), int(n),
                                              out.ctypes.data_as(C.c_void_p)), "gsdr_probe_direct_phase")
    return out


def probe_direct_tile_phase(tone_freq, rate, pos0, row0, M, n_rows=128, device: int = 0):
    """(unreduced integer phase, 32-bit phase word) of the rows of one tile, from the device functions the tensor-core DIRECT
    kernels' epilogues call (gsdr_probe_direct_tile_phase)."""
    ph = np.empty(int(n_rows), dtype=np.int64)
    wd = np.empty(int(n_rows), dtype=np.uint32)
    check(_lib.load().gsdr_probe_direct_tile_phase(int(device), int(tone_freq), int(rate), int(pos0), int(row0), int(M), int(n_rows),
                                                   ph.ctypes.data_as(C.c_void_p), wd.ctypes.data_as(C.c_void_p)),
          "gsdr_probe_direct_tile_phase")
    return ph, wd


def spec_from_samples(samples, sampling_rate=1.0, welch=None, dbc=False, rotate=True, clip_samples=False, device=0):
    """pyUSRP/USRP_noise.py:655-703 on the GPU (gsdr_spec_from_samples): returns (freqs, 10 log10 PSD of the real part,
    10 log10 PSD of the imaginary part), the order the reference returns."""
    import ctypes as C
    lib = _lib.load()
    z = np.ascontiguousarray(samples, dtype=np.complex64)
    clip = int(clip_samples) if clip_samples else 0
    w = 0 if welch is None else int(welch)
    nf = int(lib.gsdr_spec_n_freq(z.size, w, clip))
    if nf <= 0:
        raise _lib.GsdrError("spec_from_samples: empty record")
    f = np.empty(nf, dtype=np.float64)
    re = np.empty(nf, dtype=np.float32)
    im = np.empty(nf, dtype=np.float32)
    _lib.check(lib.gsdr_spec_from_samples(int(device), z.ctypes.data_as(C.c_void_p), z.size, float(sampling_rate), w, int(bool(dbc)),
                                          int(bool(rotate)), clip, f.ctypes.data_as(C.c_void_p), re.ctypes.data_as(C.c_void_p),
                                          im.ctypes.data_as(C.c_void_p)), "gsdr_spec_from_samples")
    return f, re, im
