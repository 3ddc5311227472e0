"""TX twin: multi-tone period buffer and chirp synthesis vs the fp64 oracle."""
import numpy as np
import pytest

from common import TOL, chirp_param, g, orc

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


@pytest.mark.parametrize("rate,L,freq,ampl", [
    (1_000_000, 50_000, [1000, -2500, 333_333, -499_999, 77], [0.2, 0.1, 0.3, 0.15, 0.05]),
    (100_000, 10_000, [1000, -2500, 33_333, -49_999, 77, 1000, 0], [0.2, 0.1, 0.3, 0.15, 0.05, 0.25, 0.4]),
    (30_000, 50_000, [1000, -2500], [0.5, 0.25]),
    (8_193, 1_000, [1, -1, 4096], [0.3, 0.3, 0.3]),
])
def test_tones_vs_oracle(rate, L, freq, ampl):
    p = g.param(mode="TX", rate=rate, buffer_len=L, freq=freq, ampl=ampl, wave_type=[g.TONES] * len(freq))
    tx = g.TX_buffer_generator(p)
    assert not tx.dynamic_buffer()
    o = orc.ToneGenerator(rate, freq, ampl, L)
    for _ in range(2 * (rate // L + 2)):  # walk through the period wrap more than once
        a = tx.get()
        want = o.get()
        assert orc.rel_l2(a, want) <= TOL
    tx.close()


def test_many_tones_period():
    """300 tones at 1/300 amplitude on a 2e6-sample period (scaled cfg4 TX)."""
    rate, L, T = 2_000_000, 50_000, 300
    rng = np.random.default_rng(3)
    freq = [int(v) for v in rng.choice(np.arange(-rate // 2 + 1, rate // 2), size=T, replace=False)]
    p = g.param(mode="TX", rate=rate, buffer_len=L, freq=freq, ampl=[1.0 / T] * T, wave_type=[g.TONES] * T)
    tx = g.TX_buffer_generator(p)
    got = np.concatenate([tx.get().copy() for _ in range(2)])
    tx.close()
    n = np.arange(2 * L, dtype=np.int64)
    want = np.zeros(2 * L, dtype=np.complex128)
    for f in freq:
        want += np.exp(2j * np.pi * (((f % rate) * n) % rate) / rate) / T
    assert orc.rel_l2(got, want) <= TOL


@pytest.mark.parametrize("steps,t,L,f0,f1", [(1000, 0.01, 100_000, -50_000_000, 50_000_000), (0, 0.001, 50_000, 10_000_000, -30_000_000),
                                            (0, 0.0000001, 20_000, 0, 1_000_000), (100_000, 1.0, 1_000_000, -50_000_000, 50_000_000)])
def test_chirp_vs_oracle(steps, t, L, f0, f1):
    p = chirp_param(f0=f0, f1=f1, steps=steps, t=t, L=L, ampl=0.7)
    p.mode = "TX"
    tx = g.TX_buffer_generator(p)
    assert tx.dynamic_buffer()
    cp, op = tx.chirp_param(), orc.chirp_params(p.rate, f0, f1, steps, t, tx=True)
    assert (cp.num_steps, cp.length, cp.chirpness, cp.f0) == (op.num_steps, op.length, op.chirpness, op.f0)
    o = orc.ChirpGenerator(p.rate, f0, f1, steps, t, 0.7, L)
    buf = g.pinned_empty(L)
    for _ in range(4):
        a = tx.get(buf)
        assert a.ctypes.data == buf.ctypes.data
        assert orc.rel_l2(a, o.get()) <= TOL
    tx.close()


def test_device_resident_tx_matches_host_path():
    p = chirp_param(steps=1000, t=0.01, L=50_000, ampl=0.5)
    a, b = g.TX_buffer_generator(p), g.TX_buffer_generator(p)
    buf = g.pinned_empty(50_000)
    host = np.concatenate([a.get(buf).copy() for _ in range(3)])
    d = g.DeviceBuffer(150_000)
    b.get_device(d.ptr, 3)
    b.sync()
    assert np.array_equal(d.download(), host)
    a.close()
    b.close()


def test_unsupported_tx_types():
    for w in (g.NODSP, g.SWONLY, g.RAMP, g.DIRECT):
        with pytest.raises(g.GsdrError):
            g.TX_buffer_generator(g.param(rate=1_000_000, buffer_len=10_000, freq=[1], ampl=[1], wave_type=[w]))
