"""The drop-in boundary: libgsdr.so loads, exports every symbol include/gsdr.h declares, and its
DSP entry points fail loudly (never fall back) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, g, has_gpu
from gpu_sdr_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsdr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsdr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_struct_layouts():
    assert C.sizeof(_lib.Float2) == 8
    assert C.sizeof(_lib.ChirpParam) == 24
    assert C.sizeof(_lib.BufferHelper) == 40 and C.sizeof(_lib.VnaHelper) == 24
    assert C.sizeof(_lib.CParam) == 8 + 5 * 8 + 6 * 16


def test_no_library_means_loud_failure(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libgsdr.so"))
    with pytest.raises(_lib.GsdrError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_does_not_use_oracle_torch_or_fft_libraries():
    """The product path may not route through oracle/, PyTorch or cuFFT/cuBLAS."""
    pkg = os.path.join(ROOT, "gpu_sdr_b200")
    bad = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|import\s+torch|#include\s*[<\"]cufft|#include\s*[<\"]cublas|cufft\w+\s*\(|cublas\w+\s*\()")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if not f.endswith(".py"):  # comments cite the reference's library calls; code must not make them
                    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
                    src = re.sub(r"//[^\n]*", "", src)
                assert not bad.search(src), (f, bad.search(src).group(0))
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert "-lcufft" not in mk and "-lcublas" not in mk


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_rx_tx_create_fail_loudly_without_device():
    p = g.param(rate=1_000_000, fft_tones=64, pf_average=4, buffer_len=10_000, freq=[1000], wave_type=[g.TONES], ampl=[1.0])
    with pytest.raises(g.GsdrError, match="no CUDA device"):
        g.RX_buffer_demodulator(p)
    with pytest.raises(g.GsdrError, match="no CUDA device"):
        g.TX_buffer_generator(p)


def test_param_roundtrip_to_c():
    p = g.param(rate=5, fft_tones=7, decim=3, pf_average=2, buffer_len=11, freq=[1, -2], ampl=[0.5, 0.25],
                wave_type=[g.DIRECT, g.DIRECT], chirp_t=[1.5], chirp_f=[9], swipe_s=[4])
    c, keep = p.to_c()
    assert (c.rate, c.fft_tones, c.decim, c.pf_average, c.buffer_len) == (5, 7, 3, 2, 11)
    assert [c.freq[i] for i in range(c.n_freq)] == [1, -2]
    assert [c.wave_type[i] for i in range(c.n_wave_type)] == [6, 6]
    assert np.isclose(c.chirp_t[0], 1.5) and c.chirp_f[0] == 9 and c.swipe_s[0] == 4
