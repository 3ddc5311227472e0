"""Transport-buffer size limits.  The server clamps `buffer_len` to [5e4, 6e6] (chk_param,
cpp/USRP_JSON_interpreter.cpp:268-438); both ends of that range go through every RX mode here, against the
fp64 oracle (float outputs: relative L2 <= 1e-5; valid lengths exact), two consecutive buffers each so the
carry-over at these sizes is exercised as well."""
import numpy as np
import pytest

from common import TOL, chirp_param, direct_param, g, orc, pfb_param, quantize_iq, rx_run, tone_stream

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]

L_MIN, L_MAX = 50_000, 6_000_000


def compare(ours, oracle, bufs):
    worst = 0.0
    for a, x in zip(ours, bufs):
        want = oracle.process(x)
        assert len(a) == len(want)
        if len(want):
            worst = max(worst, orc.rel_l2(a, want))
    assert worst <= TOL, worst


@pytest.mark.parametrize("L", [L_MIN, L_MAX])
def test_tones_at_buffer_len_limits(L):
    p = pfb_param(L=L)
    bufs = [tone_stream(p.rate, p.freq[:4], p.ampl[:4], i * L, L) for i in range(2)]
    rx = g.RX_buffer_demodulator(p)
    assert "pfb_fused" in rx.kernel_name()
    assert rx.max_output() >= 1000 * (L // 2048 + 1)
    rx.close()
    compare(rx_run(p, bufs), orc.PFBDemodulator(p.rate, 2048, 4, L, p.freq), bufs)


@pytest.mark.parametrize("L", [L_MIN, L_MAX])
def test_direct_at_buffer_len_limits(L):
    p = direct_param(T=4, L=L)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * L, L) for i in range(2)]
    ours = rx_run(p, bufs)
    assert all(len(o) == 4 * L // 100 for o in ours)
    compare(ours, orc.DirectDemodulator(p.rate, p.freq, p.decim, p.pf_average, L), bufs)


@pytest.mark.parametrize("L", [L_MIN, L_MAX])
def test_chirp_lockin_at_buffer_len_limits(L):
    p = chirp_param(L=L)
    gen = orc.ChirpGenerator(p.rate, p.freq[0], p.chirp_f[0], p.swipe_s[0], p.chirp_t[0], 1.0, L)
    rng = np.random.default_rng(11)
    bufs = [quantize_iq(gen.get() * 0.5 + 1e-3 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))) for _ in range(2)]
    ours = rx_run(p, bufs)
    assert all(len(o) == L // 2000 for o in ours)
    compare(ours, orc.ChirpDemodulator(p.rate, p.freq[0], p.chirp_f[0], p.swipe_s[0], p.chirp_t[0], p.decim, L), bufs)
