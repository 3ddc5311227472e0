"""TONES-mode PFB channelizer: CUDA path (through the C-ABI) vs the fp64 oracle, vs cuFFT, and
size-independent properties at the full cfg2 size.  Float tolerance: relative L2 <= 1e-5
(BASELINE.json north_star); valid lengths and bins are integers and must match exactly."""
import numpy as np
import pytest

from common import TOL, g, orc, pfb_param, rx_run, tone_stream

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def check_against_oracle(p, bufs, expect_kernel=None):
    rx = g.RX_buffer_demodulator(p)
    if expect_kernel:
        assert expect_kernel in rx.kernel_name(), rx.kernel_name()
    o = orc.PFBDemodulator(p.rate, p.fft_tones, p.pf_average, p.buffer_len, p.freq)
    assert np.array_equal(rx.bins(), np.where(o.bins < 0, 0, o.bins))
    assert np.array_equal(rx.taps(), o.window32)
    out = g.pinned_empty(rx.max_output())
    worst = 0.0
    for x in bufs:
        n = rx.process(x, out)
        want = o.process(x)
        assert n == len(want)  # integer bookkeeping: exact
        worst = max(worst, orc.rel_l2(out[:n], want))
    rx.close()
    assert worst <= TOL, worst
    return worst


@pytest.mark.parametrize("P", [1, 2, 3, 4])
@pytest.mark.parametrize("T", [1, 17, 1000])
def test_fused_2048_vs_oracle(P, T):
    p = pfb_param(N=2048, P=P, T=T, L=200_000)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(4)]
    check_against_oracle(p, bufs, "pfb_fused")


def test_cfg2_full_size_eight_buffers():
    """cfg2 exactly: 1e6-sample buffers, 8 of them so the 6-state carry-over cycle is covered."""
    p = pfb_param()
    rx = g.RX_buffer_demodulator(p)
    o = orc.PFBDemodulator(p.rate, 2048, 4, 1_000_000, p.freq)
    out = g.pinned_empty(rx.max_output())
    lens = []
    for i in range(8):
        x = tone_stream(p.rate, p.freq[:64], p.ampl[:64], i * 1_000_000, 1_000_000)
        n = rx.process(x, out)
        want = o.process(x)
        lens.append(n)
        assert n == len(want)
        assert orc.rel_l2(out[:n], want) <= TOL
    assert lens[:5] == [485000, 488000, 488000, 489000, 488000]
    rx.close()


@pytest.mark.parametrize("N,P,T,L,rate", [(100, 3, 7, 50_000, 1_000_000), (64, 4, 8, 20_000, 1_000_000),
                                          (10, 1, 3, 50_000, 1_000_000), (2048, 8, 33, 100_000, 200_000_000),
                                          (4096, 4, 50, 100_000, 200_000_000), (1000, 2, 100, 60_000, 100_000_000),
                                          (3001, 2, 5, 50_000, 1_000_000)])
def test_generic_path_vs_oracle(N, P, T, L, rate):
    p = pfb_param(rate=rate, N=N, P=P, T=T, L=L)
    bufs = [tone_stream(rate, p.freq, p.ampl, i * L, L) for i in range(3)]
    # pf_average in {1,2,4,8}: the filter bank runs as a GEMM on the tensor cores; otherwise the CUDA-core FIR + DFT pair
    check_against_oracle(p, bufs, "direct_fir_i8_kernel" if P in (1, 2, 4, 8) else "generic")


@pytest.mark.parametrize("N,P,T,L,rate", [(64, 4, 8, 20_000, 1_000_000), (1000, 2, 100, 60_000, 100_000_000)])
def test_generic_cuda_core_pair_still_matches(N, P, T, L, rate, monkeypatch):
    monkeypatch.setenv("GSDR_PFB_VARIANT", "generic")
    p = pfb_param(rate=rate, N=N, P=P, T=T, L=L)
    bufs = [tone_stream(rate, p.freq, p.ampl, i * L, L) for i in range(3)]
    check_against_oracle(p, bufs, "generic")


def test_odd_buffer_length_uses_unaligned_path():
    p = pfb_param(N=2048, P=4, T=40, L=100_001)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(5)]
    check_against_oracle(p, bufs, "pfb_fused")


def test_buffer_shorter_than_filter_span():
    """L < P*N: some calls produce zero frames; the carry-over must keep accumulating."""
    p = pfb_param(N=2048, P=4, T=9, L=3000)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * 3000, 3000) for i in range(9)]
    check_against_oracle(p, bufs)


def test_known_answer_bin_centre_tone():
    """A tone exactly on a bin centre comes out as its amplitude in every frame (taps sum to 1)."""
    rate, N, L = 200_000_000, 2048, 100_000
    rate, f = 204_800_000, 37 * 100_000  # bin width exactly 100 kHz
    p = g.param(rate=rate, fft_tones=N, pf_average=4, buffer_len=L, freq=[f, -f], wave_type=[g.TONES] * 2, ampl=[1, 1])
    n = np.arange(L, dtype=np.int64)
    x = (0.25 * np.exp(2j * np.pi * ((f * n) % rate) / rate)).astype(np.complex64)
    out = rx_run(p, [x])[0].reshape(-1, 2)
    assert np.allclose(np.abs(out[:, 0]), 0.25, rtol=2e-5)
    assert np.all(np.abs(out[:, 1]) < 1e-3)  # far-away bin: only filter leakage


def test_scaling_by_power_of_two_is_bit_exact():
    """Linearity property at the full cfg2 size: every stage is linear, so x/4 -> out/4 exactly."""
    p = pfb_param()
    x = tone_stream(p.rate, p.freq[:16], p.ampl[:16], 0, p.buffer_len)
    a = rx_run(p, [x])[0]
    b = rx_run(p, [(x * np.float32(0.25)).astype(np.complex64)])[0]
    assert np.array_equal(a * np.float32(0.25), b)


def test_device_batch_equals_sequential_calls():
    """process_device over n buffers == n process() calls, bit for bit, including valid lengths."""
    p = pfb_param(N=2048, P=4, T=100, L=150_000)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(5)]
    seq = rx_run(p, bufs)
    rx = g.RX_buffer_demodulator(p)
    din = g.DeviceBuffer(5 * p.buffer_len)
    din.upload(np.concatenate(bufs))
    dout = g.DeviceBuffer(rx.max_output_batch(5))
    tot, lens = rx.process_device(din.ptr, 5, dout.ptr)
    rx.sync()
    got = dout.download(tot)
    assert lens == [len(s) for s in seq]
    assert np.array_equal(got, np.concatenate(seq))
    # two more single-buffer calls continue the stream seamlessly
    more = [tone_stream(p.rate, p.freq, p.ampl, (5 + i) * p.buffer_len, p.buffer_len) for i in range(2)]
    out = g.pinned_empty(rx.max_output())
    o = orc.PFBDemodulator(p.rate, 2048, 4, p.buffer_len, p.freq)
    for x in bufs:
        o.process(x)
    for x in more:
        n = rx.process(x, out)
        want = o.process(x)
        assert n == len(want) and orc.rel_l2(out[:n], want) <= TOL
    rx.close()


def test_group_launch_matches_individual_streams():
    ps = [pfb_param(N=2048, P=4, T=t, L=120_000, seed=s) for t, s in ((100, 1), (1000, 2), (7, 3))]
    bufs = [[tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len, seed=10 + j) for i in range(3)]
            for j, p in enumerate(ps)]
    want = [np.concatenate(rx_run(p, b)) for p, b in zip(ps, bufs)]
    rxs = [g.RX_buffer_demodulator(p) for p in ps]
    grp = g.RxGroup(rxs)
    dins = []
    for b in bufs:
        d = g.DeviceBuffer(3 * 120_000)
        d.upload(np.concatenate(b))
        dins.append(d)
    douts = [g.DeviceBuffer(r.max_output_batch(3)) for r in rxs]
    tot, lens = grp.process_device([d.ptr for d in dins], 3, [d.ptr for d in douts])
    grp.sync()
    for j in range(3):
        got = douts[j].download(int(lens[j].sum()))
        assert np.array_equal(got, want[j])
    assert tot == sum(len(w) for w in want)
    grp.close()
    for r in rxs:
        r.close()


def test_noise_mode_full_spectrum():
    """NOISE (decim=0): every bin in natural order, copy_size = N*current_batch."""
    rate, N, L = 200_000_000, 2048, 100_000
    p = g.param(rate=rate, fft_tones=N, pf_average=4, buffer_len=L, freq=[1_000_000], wave_type=[g.NOISE], ampl=[1.0])
    x = tone_stream(rate, [1_000_000, -30_000_000], [0.3, 0.2], 0, L)
    out = rx_run(p, [x])[0]
    o = orc.PFBDemodulator(rate, N, 4, L, list(range(0)))
    o.bins = np.arange(N, dtype=np.int32)
    o.T = N
    want = o.process(x)
    assert len(out) == len(want) == N * 45
    assert orc.rel_l2(out, want) <= TOL


def test_in_kernel_fft_agrees_with_cufft():
    """north_star: the in-shared-memory radix FFT is checked against cuFFT (via torch.fft on the GPU)."""
    torch = pytest.importorskip("torch")
    p = pfb_param(N=2048, P=4, T=2048 - 1, L=100_000)
    x = tone_stream(p.rate, p.freq[:50], p.ampl[:50], 0, p.buffer_len)
    rx = g.RX_buffer_demodulator(p)
    bins, taps = rx.bins(), rx.taps().reshape(4, 2048)
    rx.close()
    ours = rx_run(p, [x])[0].reshape(-1, len(bins))
    cb = ours.shape[0]
    rows = x[: (cb + 3) * 2048].reshape(cb + 3, 2048)
    z = sum(rows[i:i + cb] * taps[i] for i in range(4)).astype(np.complex64)
    spec = torch.fft.fft(torch.from_numpy(z).cuda(), dim=1).cpu().numpy()
    assert orc.rel_l2(ours, spec[:, bins]) <= TOL


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(g.GsdrError, match="Mixed RX"):
        g.RX_buffer_demodulator(g.param(rate=1_000_000, fft_tones=64, buffer_len=10_000, freq=[1, 2], wave_type=[g.TONES, g.DIRECT]))
    with pytest.raises(g.GsdrError, match="Multiple chirp"):
        g.RX_buffer_demodulator(g.param(rate=1_000_000, buffer_len=10_000, freq=[1, 2], wave_type=[g.CHIRP, g.CHIRP],
                                        chirp_t=[1, 1], chirp_f=[2, 2], swipe_s=[3, 3]))
    with pytest.raises(g.GsdrError):
        g.RX_buffer_demodulator(g.param(rate=1_000_000, buffer_len=10_000, freq=[1], wave_type=[g.RAMP]))


def test_nodsp_passthrough():
    p = g.param(rate=1_000_000, buffer_len=10_000, wave_type=[])
    x = tone_stream(1_000_000, [1000], [0.5], 0, 10_000)
    assert np.array_equal(rx_run(p, [x])[0], x)


@pytest.mark.parametrize("decim,L", [(4, 100_000), (7, 60_000), (100, 100_000)])
def test_noise_mode_spectral_decimation(decim, L):
    """NOISE with decim > 0 (process_pfb_spec + decimate_spectra, cpp/USRP_demodulator.cpp:568-649): the mean of every
    `decim` consecutive spectra, groups running across buffer boundaries.  The reference's own version accumulates
    into a never-zeroed buffer with float atomics, so the check is against the function it is meant to
    compute, evaluated in fp64 on the oracle's spectra."""
    rate, N = 200_000_000, 2048
    p = g.param(rate=rate, fft_tones=N, pf_average=4, buffer_len=L, decim=decim, freq=[1_000_000], wave_type=[g.NOISE], ampl=[1.0])
    # tones on bin centres (multiples of 4 bins are integer Hz): their phase is the same in every frame, so the
    # coherent mean does not cancel and the float32 comparison is meaningful for large `decim` too
    tones = [12 * 97_656.25, -300 * 97_656.25, 800 * 97_656.25]
    bufs = [tone_stream(rate, [int(t) for t in tones], [0.3, 0.2, 0.1], i * L, L) for i in range(5)]
    outs = rx_run(p, bufs)
    o = orc.PFBDemodulator(rate, N, 4, L, [])
    o.bins = np.arange(N, dtype=np.int32)
    o.T = N
    spectra = [o.process(x).reshape(-1, N) for x in bufs]
    # per-buffer valid lengths: groups that complete inside each buffer
    carried, want_lens = 0, []
    for s in spectra:
        gb = (carried + len(s)) // decim
        carried = carried + len(s) - gb * decim
        want_lens.append(gb * N)
    assert [len(x) for x in outs] == want_lens
    allspec = np.concatenate(spectra)
    ng = len(allspec) // decim
    want = allspec[: ng * decim].reshape(ng, decim, N).mean(axis=1).reshape(-1)
    got = np.concatenate(outs)
    assert len(got) == len(want) and len(got) > 0
    assert orc.rel_l2(got, want) <= TOL
    # the result must not depend on how the stream was cut into calls: one device-resident batch of all buffers
    rx = g.RX_buffer_demodulator(p)
    d_in = g.DeviceBuffer(len(bufs) * L)
    for i, x in enumerate(bufs):
        d_in.upload(x, offset=i * L)
    d_out = g.DeviceBuffer(rx.max_output_batch(len(bufs)))
    total, lens = rx.process_device(d_in.ptr, len(bufs), d_out.ptr)
    rx.sync()
    batch = d_out.download(total)
    rx.close()
    assert list(lens) == want_lens
    assert np.array_equal(batch.view(np.uint32), got.view(np.uint32))


def test_several_tiles_per_cta_same_results():
    """A launch over more streams than SMs makes every persistent CTA walk several tiles (frame counters, ring slots
    and spectrum-tile parities run on across tile boundaries, tone stores are drained per tile).  The same path is
    forced here with a capped grid (GSDR_PFB_MAX_GRID, read once per process -> subprocess) on a group of 3 streams with
    different tone lists, and must reproduce the full-grid results bit for bit."""
    import os
    import subprocess
    import sys
    import tempfile
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from common import g, pfb_param, tone_stream
L, nb = 150_000, 3
ps = [pfb_param(N=2048, P=4, T=t, L=L, seed=s) for t, s in ((1000, 1), (37, 2), (513, 3))]
rxs = [g.RX_buffer_demodulator(p) for p in ps]
grp = g.RxGroup(rxs)
ins, outs = [], []
for p in ps:
    d = g.DeviceBuffer(nb * L)
    for b in range(nb):
        d.upload(tone_stream(p.rate, p.freq[:20], p.ampl[:20], b * L, L), offset=b * L)
    ins.append(d)
    outs.append(g.DeviceBuffer(rxs[0].max_output_batch(nb) * 2))
res = []
for rep in range(2):  # second call: carried-over history
    tot, lens = grp.process_device([d.ptr for d in ins], nb, [o.ptr for o in outs])
    grp.sync()
    res += [o.download(int(l.sum())) for o, l in zip(outs, lens)]
np.savez(sys.argv[1], *res)
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        files = []
        for cap in ("0", "5"):
            f = os.path.join(td, f"out_{cap}.npz")
            env = dict(os.environ, GSDR_PFB_MAX_GRID=cap)
            subprocess.run([sys.executable, "-c", code, f], check=True, env=env, timeout=300)
            files.append(np.load(f))
        a, b = files
        assert len(a.files) == len(b.files) == 6
        for k in a.files:
            assert a[k].size > 0 and np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


@pytest.mark.parametrize("decim,L,T", [(4, 100_000, 50), (7, 60_000, 1000), (1, 60_000, 20), (100, 100_000, 3)])
def test_tones_mode_post_pfb_decimation(decim, L, T):
    """TONES with decim > 0 (decimate_pfb + accumulate_ffts, cpp/USRP_demodulator.cpp:520-545,
    cpp/kernels.cu:754-790).  The reference's kernel indexes input[j * (offset % nfft)] instead of frame j and counts
    floor(current_batch / decim) outputs per buffer without carrying the remainder, so parity is defined against the
    function it is meant to compute: the mean of every `decim` consecutive frames per selected tone, groups running
    across buffer boundaries, evaluated in fp64 on the oracle's channelizer output.  decim = 1 is the identity."""
    p = pfb_param(T=T, L=L)
    p.decim = decim
    # the power sits on selected tones (bin centres, < 1 Hz off after int() truncation: the same phase in every frame,
    # so the coherent mean does not cancel and the float32 comparison is meaningful for large `decim` too)
    lit = list(p.freq[:min(T, 6)])
    bufs = [tone_stream(p.rate, lit, [0.15] * len(lit), i * L, L) for i in range(5)]
    outs = rx_run(p, bufs)
    o = orc.PFBDemodulator(p.rate, p.fft_tones, p.pf_average, L, p.freq)
    frames = [o.process(x).reshape(-1, T) for x in bufs]
    carried, want_lens = 0, []
    for f in frames:
        gb = (carried + len(f)) // decim
        carried = carried + len(f) - gb * decim
        want_lens.append(gb * T)
    assert [len(x) for x in outs] == want_lens
    allf = np.concatenate(frames)
    ng = len(allf) // decim
    want = allf[: ng * decim].reshape(ng, decim, T).mean(axis=1).reshape(-1)
    got = np.concatenate(outs)
    assert len(got) == len(want) and len(got) > 0
    assert orc.rel_l2(got, want) <= TOL
    # independent of how the stream is cut into calls
    rx = g.RX_buffer_demodulator(p)
    d_in = g.DeviceBuffer(len(bufs) * L)
    for i, x in enumerate(bufs):
        d_in.upload(x, offset=i * L)
    d_out = g.DeviceBuffer(rx.max_output_batch(len(bufs)))
    total, lens = rx.process_device(d_in.ptr, len(bufs), d_out.ptr)
    rx.sync()
    batch = d_out.download(total)
    rx.close()
    assert list(lens) == want_lens
    assert np.array_equal(batch.view(np.uint32), got.view(np.uint32))
