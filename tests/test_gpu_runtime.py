"""Pinned pool, pipelined submit/wait, replay source, full-duplex loop (cfg4, scaled)."""
import threading

import numpy as np
import pytest

from common import TOL, g, orc, pfb_param, rx_run, tone_stream

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def test_pool_get_trash_close():
    pool = g.preallocator(10_000, 6)
    assert pool.size() == 5  # the reference pre-fills pipe_size-1 buffers
    got = [pool.get() for _ in range(5)]
    assert len({b.ctypes.data for b in got}) == 5
    for b in got:
        b[:] = 1 + 2j
    extra = pool.get()  # pool dry: the grower thread supplies one more, like the reference's filler
    assert pool.size() >= 6
    for b in got + [extra]:
        pool.trash(b)
    assert pool.available() == pool.size()
    pool.close()


def test_pool_blocks_until_trash_when_growth_disabled():
    pool = g.preallocator(1000, 3, prefill_init=False)
    a, b = pool.get(), pool.get()
    res = []
    t = threading.Thread(target=lambda: res.append(pool.get()))
    t.start()
    t.join(0.3)
    assert t.is_alive()
    pool.trash(a)
    t.join(2)
    assert not t.is_alive() and res[0].ctypes.data == a.ctypes.data
    pool.trash(b)
    pool.trash(res[0])
    pool.close()


def test_pipelined_submit_wait_matches_blocking():
    p = pfb_param(N=2048, P=4, T=200, L=200_000)
    bufs = [tone_stream(p.rate, p.freq, p.ampl, i * p.buffer_len, p.buffer_len) for i in range(7)]
    want = rx_run(p, bufs)
    rx = g.RX_buffer_demodulator(p)
    pool_in, pool_out = g.preallocator(p.buffer_len, 8), g.preallocator(rx.max_output(), 8)
    pending, got = [], []
    for x in bufs:
        hin, hout = pool_in.get(), pool_out.get()
        hin[:] = x
        t, n = rx.submit(hin, hout)
        pending.append((t, n, hin, hout))
        if len(pending) == 3:
            t0, n0, i0, o0 = pending.pop(0)
            rx.wait(t0)
            got.append(o0[:n0].copy())
            pool_in.trash(i0)
            pool_out.trash(o0)
    for t0, n0, i0, o0 in pending:
        rx.wait(t0)
        got.append(o0[:n0].copy())
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    rx.close()
    pool_in.close()
    pool_out.close()


def test_replay_loopback_tones_known_answer():
    """Replay source in TX_LOOP mode == --sw_loop: RX sees exactly the TX waveform; with tones on
    bin centres every selected bin is the tone amplitude in every frame."""
    rate, N, L, T = 204_800_000 // 100, 2048, 100_000, 16  # 2.048 MS/s: bin width 1 kHz exactly
    ks = [3, -7, 100, -1000, 511, 77, -300, 900, 12, -13, 640, -641, 1, -1, 250, -999]
    freq = [k * 1000 for k in ks]
    p = g.param(mode="RX", rate=rate, fft_tones=N, pf_average=4, buffer_len=L, freq=freq, ampl=[1.0 / T] * T, wave_type=[g.TONES] * T)
    pool = g.preallocator(L, 8)
    src = g.ReplaySource(p, pool, kind=g.ReplaySource.TX_LOOP, front_end_code="B")
    rx = g.RX_buffer_demodulator(p)
    out = g.pinned_empty(rx.max_output())
    for k in range(4):
        pkt, buf = src.next()
        assert (pkt.packet_number, pkt.length, pkt.errors, pkt.channels, pkt.front_end_code) == (k + 1, L, 0, T, b"B")
        n = rx.process(buf, out)
        pool.trash(buf)
        y = out[:n].reshape(-1, T)
        assert np.allclose(np.abs(y), 1.0 / T, rtol=2e-2)  # other tones leak in at the prototype filter's stop-band level
    rx.close()
    src.close()
    pool.close()


def test_replay_with_noise_is_deterministic_and_parity_holds():
    p = pfb_param(rate=2_000_000, N=2048, P=4, T=50, L=60_000)
    outs = []
    for _ in range(2):
        pool = g.preallocator(p.buffer_len, 4)
        src = g.ReplaySource(p, pool, kind=g.ReplaySource.TONES_NOISE, noise_sigma=1e-3, seed=99)
        bufs = []
        for _k in range(3):
            _pkt, b = src.next()
            bufs.append(b[: p.buffer_len].copy())
            pool.trash(b)
        src.close()
        pool.close()
        outs.append(bufs)
    assert all(np.array_equal(a, b) for a, b in zip(*outs))
    o = orc.PFBDemodulator(p.rate, 2048, 4, p.buffer_len, p.freq)
    for a, x in zip(rx_run(p, outs[0]), outs[0]):
        assert orc.rel_l2(a, o.process(x)) <= TOL


def test_full_duplex_two_front_ends_concurrently():
    """cfg4 (scaled): two front-ends, each TX TONES -> loopback -> RX, driven from two threads on one GPU."""
    rate, L, T = 2_048_000, 100_000, 100
    results = {}

    def front_end(code, seed):
        rng = np.random.default_rng(seed)
        ks = rng.choice(np.arange(-1000, 1000), size=T, replace=False)
        freq = [int(k) * 1000 for k in ks]
        p = g.param(mode="RX", rate=rate, fft_tones=2048, pf_average=4, buffer_len=L, freq=freq, ampl=[1.0 / T] * T,
                    wave_type=[g.TONES] * T)
        tx = g.TX_buffer_generator(p)
        rx = g.RX_buffer_demodulator(p)
        out = g.pinned_empty(rx.max_output())
        ok = True
        for _ in range(5):
            n = rx.process(tx.get(), out)
            ok = ok and np.allclose(np.abs(out[:n]), 1.0 / T, rtol=5e-2)
        rx.close()
        tx.close()
        results[code] = ok

    th = [threading.Thread(target=front_end, args=(c, s)) for c, s in (("A", 1), ("B", 2))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert results == {"A": True, "B": True}


@pytest.mark.parametrize("mode", ["tones", "direct", "chirp", "nodsp"])
def test_sc16_ingest_equals_host_side_conversion(mode):
    """gsdr_rx_process_sc16 (wire-format int16 I/Q, converted on the GPU) must equal gsdr_rx_process on the buffer a
    host would have produced with UHD's sc16 -> fc32 rule, (float)v * (1/32767), bit for bit -- and the fp64 oracle."""
    from common import chirp_param, direct_param
    L = 100_003 if mode == "nodsp" else 100_000  # the ragged tail of the conversion kernel is exercised by nodsp only
    if mode == "tones":
        p = pfb_param(N=2048, P=4, T=300, L=L)
        o = orc.PFBDemodulator(p.rate, p.fft_tones, p.pf_average, L, p.freq)
    elif mode == "direct":
        p = direct_param(rate=10_000_000, T=5, decim=50, f=4, L=L)
        o = orc.DirectDemodulator(p.rate, p.freq, p.decim, p.pf_average, L)
    elif mode == "chirp":
        p = chirp_param(steps=1000, t=0.01, decim=2, L=L)
        o = orc.ChirpDemodulator(p.rate, p.freq[0], p.chirp_f[0], p.swipe_s[0], p.chirp_t[0], p.decim, L)
    else:
        p = g.param(rate=1_000_000, buffer_len=L, wave_type=[g.NODSP], freq=[0], ampl=[1.0])
        o = None
    rng = np.random.default_rng(99)
    raws = [rng.integers(-20000, 20000, size=2 * L, dtype=np.int16) for _ in range(3)]
    as_float = [(r.astype(np.float32) * np.float32(1.0 / 32767.0)).view(np.complex64) for r in raws]
    want = rx_run(p, as_float)
    rx = g.RX_buffer_demodulator(p)
    out = g.pinned_empty(rx.max_output())
    pinned_raw = g.pinned_empty((L + 1) // 2 + 1).view(np.int16)
    try:
        for r, w, xf in zip(raws, want, as_float):
            pinned_raw[:2 * L] = r
            n = rx.process_sc16(pinned_raw, out)
            assert n == len(w)
            assert np.array_equal(out[:n].view(np.uint32), w.view(np.uint32))
            if o is not None:
                assert orc.rel_l2(out[:n], o.process(xf)) <= TOL
    finally:
        rx.close()


def test_sc16_pipelined_submit_wait():
    p = pfb_param(N=2048, P=4, T=100, L=150_000)
    L = p.buffer_len
    rng = np.random.default_rng(5)
    raws = [rng.integers(-8000, 8000, size=2 * L, dtype=np.int16) for _ in range(6)]
    want = rx_run(p, [(r.astype(np.float32) * np.float32(1.0 / 32767.0)).view(np.complex64) for r in raws])
    rx = g.RX_buffer_demodulator(p)
    ins = [g.pinned_empty(L // 2).view(np.int16) for _ in range(3)]
    outs = [g.pinned_empty(rx.max_output()) for _ in range(3)]
    pending, got = [], []
    for i, r in enumerate(raws):
        if len(pending) == 3:
            t0, n0, k0 = pending.pop(0)
            rx.wait(t0)
            got.append(outs[k0][:n0].copy())
        k = i % 3
        ins[k][:] = r
        t, n = rx.submit_sc16(ins[k], outs[k])
        pending.append((t, n, k))
    for t0, n0, k0 in pending:
        rx.wait(t0)
        got.append(outs[k0][:n0].copy())
    rx.close()
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_packet_frame_in_place_is_one_contiguous_wire_frame():
    """gsdr_packet_frame writes the 21-byte header into the pool headroom directly in front of the payload: the frame
    the client expects (header dtype + length complex64, pyUSRP/USRP_connections.py:814-970) without a staging copy."""
    import ctypes as C
    from gpu_sdr_b200 import _lib
    lib = _lib.load()
    pool = g.preallocator(5000, 3)
    buf = pool.get()
    rng = np.random.default_rng(1)
    buf[:] = (rng.standard_normal(5000) + 1j * rng.standard_normal(5000)).astype(np.complex64)
    pkt = _lib.RxPacket()
    pkt.buffer = buf.ctypes.data
    pkt.usrp_number, pkt.front_end_code, pkt.packet_number, pkt.length, pkt.errors, pkt.channels = 0, b"D", 7, 4000, 0, 8
    frame, nbytes = C.c_void_p(), C.c_size_t()
    assert lib.gsdr_packet_frame(C.byref(pkt), C.byref(frame), C.byref(nbytes)) == 0
    assert nbytes.value == 21 + 8 * 4000 and frame.value == buf.ctypes.data - 21
    wire = C.string_at(frame.value, nbytes.value)
    header_type = np.dtype([("usrp_number", np.int32), ("front_end_code", np.dtype("|S1")), ("packet_number", np.int32),
                            ("length", np.int32), ("errors", np.int32), ("channels", np.int32)])
    h = np.frombuffer(wire[:21], dtype=header_type)[0]
    assert h["length"] == 4000 and h["front_end_code"] == b"D" and h["channels"] == 8 and h["packet_number"] == 7
    data = np.frombuffer(wire[21:], dtype=np.complex64)
    assert np.array_equal(data, buf[:4000])
    # client-side de-interleave (pyUSRP/USRP_connections.py:157)
    assert np.reshape(data, (4000 // 8, 8)).T.shape == (8, 500)
    pool.trash(buf)
    pool.close()


@pytest.mark.parametrize("L,T,mode", [(1_000_000, 1000, "tones"), (200_000, 17, "tones"), (131_073, 40, "tones"), (200_000, 0, "noise")])
def test_blocking_process_zero_copy_equals_copied(L, T, mode, monkeypatch):
    """Blocking process() on pinned buffers runs ONE kernel that reads the caller's input and writes the caller's output in
    place over PCIe (gsdr_rx_process, zero-copy form); with GSDR_PROCESS_ZEROCOPY=0, or pageable buffers, the same call
    uploads / launches / downloads in chunks.  Same kernel and frames: outputs and valid lengths must be bit-identical,
    carry-over included (5 consecutive buffers), and the input must not be modified."""
    if mode == "noise":
        p = g.param(mode="RX", rate=200_000_000, fft_tones=2048, pf_average=4, buffer_len=L, decim=0, freq=[0], wave_type=[g.NOISE], ampl=[1.0])
    else:
        p = pfb_param(N=2048, P=4, T=T, L=L)
    rng = np.random.default_rng(21)
    bufs = [(0.1 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))).astype(np.complex64) for _ in range(5)]

    def run(pinned_in):
        rx = g.RX_buffer_demodulator(p)
        pool_in = g.preallocator(L, 4)           # pool buffers: interior pointers of a pinned allocation (headroom in front)
        out = g.pinned_empty(rx.max_output())
        res, launches = [], []
        for x in bufs:
            hin = pool_in.get() if pinned_in else x.copy()
            hin[:] = x
            l0 = rx.launch_count()
            n = rx.process(hin, out)
            launches.append(rx.launch_count() - l0)
            assert np.array_equal(hin, x)
            res.append(out[:n].copy())
            if pinned_in:
                pool_in.trash(hin)
        rx.close()
        pool_in.close()
        g.pinned_free(out)
        return res, launches

    zc, zc_launches = run(True)
    assert zc_launches == [1] * 5, zc_launches      # one launch per buffer, carry-over copy included
    pageable, _ = run(False)                        # pageable input: falls back to the copied form
    monkeypatch.setenv("GSDR_PROCESS_ZEROCOPY", "0")
    copied, copied_launches = run(True)
    assert max(copied_launches) >= 2
    for a, b, c in zip(zc, copied, pageable):
        assert len(a) == len(b) == len(c) and len(a) > 0
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_diagnostic_flag_dumps_the_polyphase_window(tmp_path, monkeypatch):
    """init_diagnostic = true: the reference writes the polyphase window as float2 into USRP_polyphase_filter_window.dat
    (make_sinc_window(..., diagnostic, ...), cpp/kernels.cu:290-296, called from cpp/USRP_demodulator.cpp:134)."""
    monkeypatch.chdir(tmp_path)
    p = pfb_param(N=2048, P=4, T=5, L=100_000)
    rx = g.RX_buffer_demodulator(p, True)
    taps = rx.taps()
    rx.close()
    dump = np.fromfile(tmp_path / "USRP_polyphase_filter_window.dat", dtype=np.complex64)
    assert dump.size == 2048 * 4 and np.array_equal(dump.real, taps) and not dump.imag.any()


def test_fallback_direct_kernel_on_a_second_device_after_the_first():
    """direct_fir_kernel<staged> (pf_average > 8) raises its dynamic shared-memory limit per device: a demodulator on GPU 1
    must work after one on GPU 0 has run (the limit used to be remembered process-wide)."""
    lib = g.load()
    if lib.gsdr_device_count() < 2:
        pytest.skip("needs two GPUs")
    from common import direct_param
    p = direct_param(rate=1_000_000, T=3, decim=100, f=9, L=50_000)
    x = tone_stream(p.rate, p.freq, p.ampl, 0, p.buffer_len)
    o = orc.DirectDemodulator(p.rate, p.freq, p.decim, p.pf_average, p.buffer_len)
    want = o.process(x)
    for dev in (0, 1):
        got = rx_run(p, [x], device=dev)[0]
        assert orc.rel_l2(got, want) <= TOL


def test_pcie_probe_reports_both_directions():
    """gsdr_pcie_probe_*: the cudaMemcpyAsync ceiling bench.py reports its host-fed figures against."""
    import ctypes as C
    lib = g.load()
    probe = lib.gsdr_pcie_probe_create(0, 8 << 20, 4 << 20, 2, 6)
    assert probe
    try:
        w = (C.c_double * 2)()
        assert lib.gsdr_pcie_probe_run(probe, 1, 0, 8, w) == 0 and w[0] > 1.0 and w[1] == 0.0
        assert lib.gsdr_pcie_probe_run(probe, 0, 1, 8, w) == 0 and w[0] == 0.0 and w[1] > 1.0
        assert lib.gsdr_pcie_probe_run(probe, 1, 1, 8, w) == 0 and w[0] > 1.0 and w[1] > 0.5
        assert lib.gsdr_pcie_probe_run(probe, 0, 0, 8, w) != 0   # nothing to time
    finally:
        lib.gsdr_pcie_probe_destroy(probe)
    assert not lib.gsdr_pcie_probe_create(0, 0, 0, 1, 1)          # no bytes either way
    four = (C.c_double * 4)()
    assert lib.gsdr_pcie_copy_ceiling_streams(0, 1 << 20, 1 << 20, 4, 2, four) == 0 and min(four) > 0.5
