"""VNA chirp demodulator + lock-in and the integer phase: CUDA path vs oracle.  The int32 chirp
phase index and the signed int64 DIRECT LO phase must be BIT-EXACT; float outputs within 1e-5."""
import numpy as np
import pytest

from common import TOL, chirp_param, g, orc, quantize_iq, rx_run

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def chirp_inputs(p, nbuf, seed=5):
    gen = orc.ChirpGenerator(p.rate, p.freq[0], p.chirp_f[0], p.swipe_s[0], p.chirp_t[0], 1.0, p.buffer_len)
    rng = np.random.default_rng(seed)
    L = p.buffer_len
    return [quantize_iq(gen.get() * (0.5 * np.exp(2j * np.pi * 0.1 * i)) + 1e-3 * (rng.standard_normal(L) + 1j * rng.standard_normal(L)))
            for i in range(nbuf)]


def run_case(p, nbuf):
    bufs = chirp_inputs(p, nbuf)
    ours = rx_run(p, bufs)
    o = orc.ChirpDemodulator(p.rate, p.freq[0], p.chirp_f[0], p.swipe_s[0], p.chirp_t[0], p.decim, p.buffer_len)
    worst = 0.0
    for a, x in zip(ours, bufs):
        want = o.process(x)
        assert len(a) == len(want)  # VNA_decimator_helper bookkeeping: exact
        if len(want):
            worst = max(worst, orc.rel_l2(a, want))
    assert worst <= TOL, worst
    return ours


def test_cfg3_full_size_lockin():
    """cfg3: rate 2e8, -50..+50 MHz, 1e5 points over 1 s, decim=1 -> ppt=2000, 500 outputs/buffer."""
    p = chirp_param()
    rx = g.RX_buffer_demodulator(p)
    cp = rx.chirp_param()
    rx.close()
    assert (cp.num_steps, cp.length, cp.chirpness, cp.f0) == (100000, 2000, 21475, -1073741823)
    out = run_case(p, 4)
    assert all(len(o) == 500 for o in out)


@pytest.mark.parametrize("steps,t,decim,L", [(1000, 0.01, 3, 100_000), (1000, 0.01, 1, 70_001), (0, 0.001, 200, 100_000),
                                             (10, 0.02, 1, 50_000), (50_000, 0.0005, 7, 60_000), (1000, 0.01, 0, 100_000),
                                             (100, 0.001, 1, 50_000), (7, 0.0001, 1, 20_000)])
def test_lockin_and_full_vs_oracle(steps, t, decim, L):
    run_case(chirp_param(steps=steps, t=t, decim=decim, L=L), 5)


def test_down_chirp_wraps_like_the_reference():
    """chirp_f < freq: chirpness is a negative double cast to unsigned (wraps on x86)."""
    p = chirp_param(f0=30_000_000, f1=-40_000_000, steps=2000, t=0.005, decim=2, L=80_000)
    run_case(p, 4)


def test_known_answer_loopback_gives_scale():
    """sw-loop identity: TX chirp -> RX chirp demod -> lock-in == scale + 0j (profile sums to 1)."""
    p = chirp_param(steps=1000, t=0.01, decim=1, L=100_000)
    gen = orc.ChirpGenerator(p.rate, p.freq[0], p.chirp_f[0], 1000, 0.01, 0.7, p.buffer_len)
    bufs = [gen.get().astype(np.complex64) for _ in range(3)]
    for o in rx_run(p, bufs):
        assert np.allclose(o, 0.7, atol=2e-6)


@pytest.mark.parametrize("args", [(200_000_000, -50_000_000, 50_000_000, 100_000, 1.0), (200_000_000, 10_000_000, -30_000_000, 0, 0.001),
                                  (100_000_000, 1, 2, 3, 0.5), (200_000_000, -99_999_999, 99_999_999, 7, 0.0001)])
@pytest.mark.parametrize("last", [0, 1, 123_456_789, 199_999_000])
def test_chirp_phase_index_is_bit_exact(args, last):
    cp = g.hostlogic.chirp_params(*args)
    op = orc.chirp_params(*args)
    assert (cp.num_steps, cp.length, cp.chirpness, cp.f0) == (op.num_steps, op.length, op.chirpness, op.f0)
    period = int(cp.num_steps) * int(cp.length)
    got = g.hostlogic.probe_chirp_index(cp, last % period, 200_000)
    assert np.array_equal(got, orc.chirp_index(last % period, 200_000, op))


@pytest.mark.parametrize("tf", [12_345_677, -49_999_999, 1, -1, 0, 50_000_000])
@pytest.mark.parametrize("counter", [0, 99_000_000, 12_345])
def test_direct_lo_phase_is_bit_exact(tf, counter):
    rate = 100_000_000
    got = g.hostlogic.probe_direct_phase(tf, rate, counter, 7, 300_000)
    assert np.array_equal(got, orc.direct_phase(tf, 0, rate, counter, 7, 300_000))


@pytest.mark.parametrize("tf", [12_345_677, -49_999_999, 1, -1, 0, 50_000_000, 99_999_999])
@pytest.mark.parametrize("pos0,row0,M", [(0, 0, 100), (99_999_900, 125, 100), (3_000_000_000_123, 7_812_375, 100), (12_345, 640, 2048),
                                         (199_999_999, 1, 16_384)])
def test_direct_tile_lo_phase_is_bit_exact(tf, pos0, row0, M):
    """The tensor-core DIRECT kernels (tc and i8 epilogues) form the LO phase per tile as base + row * step: the integer must be
    congruent to the reference's (f * n) % rate (cpp/kernels.cu:59-75) and the phase word must be that integer times 2^32 / rate,
    rounded: checked here against Python integers (the word within one unit of the exactly rounded value: the product with the
    rounded double 2^32 / rate carries 2^-13 of a unit of error, and exact on all but a few rows)."""
    rate = 200_000_000
    ph, wd = g.hostlogic.probe_direct_tile_phase(tf, rate, pos0, row0, M)
    exact_ph = [((tf % rate) * ((pos0 + (row0 + r) * M) % rate)) % rate for r in range(128)]
    assert [int(p) % rate for p in ph] == exact_ph
    assert all(0 <= int(p) < 129 * rate for p in ph)
    n_exact = 0
    for r in range(128):
        num = int(ph[r]) * (1 << 32)
        q, rem = divmod(num, rate)
        lo, hi = q % (1 << 32), (q + 1) % (1 << 32)
        w = int(wd[r])
        assert w in (lo, hi) or (rem == 0 and w == lo)
        nearest = lo if 2 * rem < rate else hi if 2 * rem > rate else (lo if q % 2 == 0 else hi)
        n_exact += w == nearest
    assert n_exact >= 126


def test_device_batch_equals_sequential():
    p = chirp_param(steps=1000, t=0.01, decim=3, L=50_000)
    bufs = chirp_inputs(p, 5)
    seq = rx_run(p, bufs)
    rx = g.RX_buffer_demodulator(p)
    din = g.DeviceBuffer(5 * p.buffer_len)
    din.upload(np.concatenate(bufs))
    dout = g.DeviceBuffer(rx.max_output_batch(5))
    tot, lens = rx.process_device(din.ptr, 5, dout.ptr)
    rx.sync()
    assert lens == [len(s) for s in seq]
    assert np.array_equal(dout.download(tot), np.concatenate(seq))
    rx.close()
