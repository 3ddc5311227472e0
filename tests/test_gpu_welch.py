"""Client-side Welch spectra on the GPU (gsdr_spec_from_samples) against the fp64 restatement of
pyUSRP/USRP_noise.py:655-703 (oracle.spec_from_samples, itself pinned on scipy.signal.welch in tests/test_oracle.py).
Tolerance: relative L2 <= 1e-4 on the linear densities (a float32 Bluestein chain of up to 2^21 points), 0.01 dB on every
bin that carries power."""
import numpy as np
import pytest

from common import g, orc

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def _signal(n, seed, fs=1.0e6):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    z = (0.8 + 0.3j) * (1.0 + 0.01 * np.sin(2 * np.pi * 1234.5 * t)) * np.exp(1j * 0.002 * np.sin(2 * np.pi * 777.0 * t))
    z = z + 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) + 1e-6 * t * fs / n * (1 - 2j)
    return z.astype(np.complex64)


@pytest.mark.parametrize("n,welch,dbc,rotate,clip", [
    (100_000, 10, False, True, False),        # nperseg 10000 (not a power of two): Bluestein over 2^15
    (100_000, None, False, True, False),      # one segment of the whole record
    (131_072, 8, True, True, False),          # nperseg 16384: the power-of-two path, dBc
    (99_991, 7, True, False, 1000),           # prime length, clipped, odd nperseg 14284 -> even; no rotation
    (50_001, 3, False, True, 2500),           # odd nperseg (16667): no Nyquist bin
    (1_000_000, 20, True, True, 10_000),      # a full 1e6-sample channel as calculate_noise sees it
])
def test_spec_from_samples_matches_the_fp64_restatement(n, welch, dbc, rotate, clip):
    fs = 1.0e6
    z = _signal(n, seed=n % 97)
    f, re, im = g.hostlogic.spec_from_samples(z, sampling_rate=fs, welch=welch, dbc=dbc, rotate=rotate, clip_samples=clip)
    fo, reo, imo = orc.spec_from_samples(z, sampling_rate=fs, welch=welch, dbc=dbc, rotate=rotate, clip_samples=clip)
    assert len(f) == len(fo) and np.allclose(f, fo, rtol=1e-12, atol=0)
    for got_db, want_db in ((re, reo), (im, imo)):
        got, want = 10.0 ** (got_db.astype(np.float64) / 10.0), 10.0 ** (want_db / 10.0)
        assert np.sqrt(np.sum((got - want) ** 2) / np.sum(want ** 2)) <= 1e-4
        strong = want > 1e-7 * want.max()
        strong[:2] = False                     # the detrended DC neighbourhood is rounding noise in both
        assert np.max(np.abs(got_db[strong] - want_db[strong])) <= 0.01
