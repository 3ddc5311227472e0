"""DIRECT mode (multi-tone DDC + Hamming-sinc FIR + decimation): CUDA path vs the fp64 oracle."""
import numpy as np
import pytest

from common import TOL, direct_param, g, orc, rx_run, tone_stream

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def run_case(p, nbuf, noise=1e-3):
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len, noise=noise) for i in range(nbuf)]
    ours = rx_run(p, bufs)
    o = orc.DirectDemodulator(p.rate, p.freq, p.decim, p.pf_average, p.buffer_len)
    worst = 0.0
    for a, x in zip(ours, bufs):
        want = o.process(x)
        assert len(a) == len(want)
        worst = max(worst, orc.rel_l2(a, want))
    assert worst <= TOL, worst
    return ours


def test_cfg1_full_size():
    """cfg1: rate 1e8, 16 tones (4 negative), decim 100, 400 taps, 1e6-sample buffers."""
    p = direct_param()
    rx = g.RX_buffer_demodulator(p)
    taps = rx.taps()
    rx.close()
    assert len(taps) == 400 and int(np.argmax(taps)) == 199 and abs(float(taps.sum()) - 1) < 1e-6
    assert np.array_equal(taps, orc.make_sinc_window(400, float(np.float32(0.75 / 200))))
    out = run_case(p, 3)
    assert all(len(o) == 160_000 for o in out)


@pytest.mark.parametrize("T,decim,f,L,rate", [(5, 10, 8, 50_000, 1_000_000), (1, 1000, 4, 100_000, 100_000_000),
                                              (33, 50, 2, 50_000, 10_000_000), (4, 7, 1, 70_000, 1_000_000),
                                              (7, 2, 16, 20_000, 1_000_000), (2, 5000, 4, 100_000, 200_000_000)])
def test_decimating_vs_oracle(T, decim, f, L, rate):
    run_case(direct_param(rate=rate, T=T, decim=decim, f=f, L=L), 3)


@pytest.mark.parametrize("T", [1, 3, 300])
def test_no_decimation_vs_oracle(T):
    run_case(direct_param(rate=1_000_000, T=T, decim=0, f=1, L=20_000), 2)


def test_known_answer_single_tone_goes_to_dc():
    """sw-loop identity: a tone demodulated at its own frequency settles to DC = amplitude
    (FIR taps are sum-normalised)."""
    rate, L, f = 100_000_000, 100_000, -12_345_677
    p = g.param(rate=rate, decim=100, pf_average=4, buffer_len=L, freq=[f], wave_type=[g.DIRECT], ampl=[1.0])
    n = np.arange(2 * L, dtype=np.int64)
    x = (0.3 * np.exp(2j * np.pi * ((f * n) % rate) / rate)).astype(np.complex64)
    out = rx_run(p, [x[:L], x[L:]])
    assert np.allclose(out[1], 0.3, atol=3e-6)
    assert np.allclose(out[0][4:], 0.3, atol=3e-6)


def test_phase_continuity_across_rate_wrap():
    """index_counter wraps modulo rate (cpp/USRP_demodulator.cpp:437-440): 30 buffers of 50k at 1 MS/s."""
    run_case(direct_param(rate=1_000_000, T=3, decim=10, f=4, L=50_000), 30, noise=0.0)


def test_device_batch_equals_sequential():
    p = direct_param(rate=10_000_000, T=6, decim=20, f=4, L=40_000)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(4)]
    seq = rx_run(p, bufs)
    rx = g.RX_buffer_demodulator(p)
    din = g.DeviceBuffer(4 * p.buffer_len)
    din.upload(np.concatenate(bufs))
    dout = g.DeviceBuffer(rx.max_output_batch(4))
    tot, lens = rx.process_device(din.ptr, 4, dout.ptr)
    rx.sync()
    assert lens == [len(s) for s in seq]
    assert np.array_equal(dout.download(tot), np.concatenate(seq))
    rx.close()


def test_buffer_len_must_be_multiple_of_decim():
    with pytest.raises(g.GsdrError, match="multiple of decim"):
        g.RX_buffer_demodulator(direct_param(rate=1_000_000, T=2, decim=7, f=4, L=50_000))
