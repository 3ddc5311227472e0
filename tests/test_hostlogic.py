"""Host-side logic that defines results: product (libgsdr.so) vs the C oracle vs the golden
fixtures produced by the reference's own object code -- all BIT-EXACT.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from common import GOLDEN, g, orc, ref_lib

HL = g.hostlogic


def golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet (tests/golden/make_golden.py)")
    return np.load(path)


def test_survey_worked_values():
    # ceil-style tone->bin rule
    assert list(HL.tone_bins(200_000_000, 2048, [292968, -292968, 1, -1, 0])) == [3, 2046, 1, 0, 0]
    # row a9: cfg2 carry-over sequence (eff, cb, spare, spare_begin)
    seq = HL.buffer_helper_sequence(2048, 1_000_000, 4, 1000, 5)
    cols = [HL.BH_FIELDS.index(k) for k in ("eff_length", "current_batch", "spare_samples", "spare_begin")]
    assert seq[:, cols].tolist() == [[1000000, 485, 6720, 993280], [1006720, 488, 7296, 999424], [1007296, 488, 7872, 999424],
                                     [1007872, 489, 6400, 1001472], [1006400, 488, 6976, 999424]]
    assert HL.pfb_batching(1_000_000, 2048, 4) == 498
    # row a10: cfg3 chirp quantisation; client-side restatement pyUSRP/USRP_VNA.py:740
    cp = HL.chirp_params(200_000_000, -50_000_000, 50_000_000, 100_000, 1.0)
    assert (cp.num_steps, cp.length, cp.chirpness, cp.f0) == (100000, 2000, 21475, -1073741823)
    assert cp.chirpness == int((2**32 - 1) * (50e6 - -50e6) / (100000 - 1) / 200e6)
    # row a3/KAT 3: 400-tap cfg1 filter
    taps = HL.make_sinc_window(400, float(np.float32(0.75 / 200)))
    assert int(np.argmax(taps)) == 199 and abs(float(taps.astype(np.float64).sum()) - 1.0) < 1e-6
    # make_flat_window zeroes only the LEADING side taps
    flat = HL.make_flat_window(2000, 200)
    assert np.all(flat[:200] == 0) and np.all(flat[200:] == np.float32(1.0) / np.float32(1800.0))


@pytest.mark.parametrize("L,fc", [(8192, 1.0 / 4096), (400, 0.75 / 200), (401, 0.01), (256, 1.0 / 128), (300, 1.0 / 200),
                                  (80, 0.75 / 20), (7, 0.2), (2, 0.25), (16384, 1.0 / 8192), (20000, 0.75 / 10000)])
def test_sinc_window_product_equals_oracle(L, fc):
    fc = float(np.float32(fc))
    assert np.array_equal(HL.make_sinc_window(L, fc), orc.make_sinc_window(L, fc))


@pytest.mark.parametrize("L,side", [(2000, 200), (6000, 600), (200, 20), (7, 0), (1, 0), (33, 3), (10, 9)])
def test_flat_window_product_equals_oracle(L, side):
    assert np.array_equal(HL.make_flat_window(L, side), orc.make_flat_window(L, side))


def test_helpers_product_equals_oracle_randomised():
    rng = np.random.default_rng(7)
    for _ in range(60):
        N, P = int(rng.integers(2, 5000)), int(rng.integers(1, 9))
        L, T = int(rng.integers(max(N, 1000), 2_000_000)), int(rng.integers(1, 50))
        h = orc.BufferHelper(N, L, P, T)
        rows = []
        for _i in range(12):
            rows.append(h.state())
            h.update()
        assert np.array_equal(HL.buffer_helper_sequence(N, L, P, T, 12), np.array(rows))
        assert HL.pfb_batching(L, N, P) == orc.pfb_batching(L, N, P)
        ppt = int(rng.integers(1, 3_000_000))
        v = orc.VNAHelper(ppt, L)
        rows = []
        for _i in range(12):
            rows.append(v.state())
            v.update()
        assert np.array_equal(HL.vna_helper_sequence(ppt, L, 12), np.array(rows))


def test_tone_bins_product_equals_oracle_randomised():
    rng = np.random.default_rng(8)
    for rate, N in ((200_000_000, 2048), (100_000_000, 1000), (1_000_000, 100), (10_000_000, 4096), (250_000_000, 10), (999_983, 77)):
        freq = rng.integers(-rate // 2 - 10, rate // 2 + 10, size=300).astype(np.int32)
        bs = rate / N
        freq[:50] = (rng.integers(-N // 2, N // 2, size=50) * bs).astype(np.int32)  # truncated bin centres
        assert np.array_equal(HL.tone_bins(rate, N, freq), orc.tone_bins(rate, N, freq))


def test_chirp_params_product_equals_oracle_randomised():
    rng = np.random.default_rng(9)
    for _ in range(300):
        rate = int(rng.choice([1_000_000, 100_000_000, 200_000_000, 250_000_000]))
        f0, f1 = int(rng.integers(-rate // 2, rate // 2)), int(rng.integers(-rate // 2, rate // 2))
        steps = int(rng.choice([0, 1, 2, 10, 1000, 100_000, 1_000_000]))
        t = float(np.float32(rng.choice([1e-7, 1e-4, 1e-3, 0.01, 0.5, 1.0, 3.0])))
        for tx in (False, True):
            a, b = HL.chirp_params(rate, f0, f1, steps, t, tx), orc.chirp_params(rate, f0, f1, steps, t, tx)
            assert (a.num_steps, a.length, a.chirpness, a.f0) == (b.num_steps, b.length, b.chirpness, b.f0)


# ---- against the reference itself (fixtures made by its object code; live library when built) ----
def test_golden_windows_and_helpers():
    z = golden("host_logic.npz")
    n = 0
    for key in z.files:
        kind, *args = key.split("_")
        if kind == "sinc":
            L, fc = int(args[0]), float(args[1])
            assert np.array_equal(HL.make_sinc_window(L, fc), z[key]), key
            assert np.array_equal(orc.make_sinc_window(L, fc), z[key]), key
        elif kind == "flat":
            L, side = int(args[0]), int(args[1])
            assert np.array_equal(HL.make_flat_window(L, side), z[key]), key
            assert np.array_equal(orc.make_flat_window(L, side), z[key]), key
        elif kind == "bh":
            N, L, P, T = map(int, args)
            assert np.array_equal(HL.buffer_helper_sequence(N, L, P, T, 16), z[key]), key
        elif kind == "vna":
            ppt, L = map(int, args)
            assert np.array_equal(HL.vna_helper_sequence(ppt, L, 16), z[key]), key
        n += 1
    assert n >= 20


def test_golden_tone_bins():
    z = golden("tone_bins.npz")
    for tag, rate, N in (("edge", 200_000_000, 2048), ("full", 200_000_000, 2048), ("odd", 1_000_000, 100)):
        f, want = z[tag + "_freq"], z[tag + "_bins"]
        assert np.array_equal(HL.tone_bins(rate, N, f), want), tag
        assert np.array_equal(orc.tone_bins(rate, N, f), want), tag


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_object_code_host_side():
    """The reference's own make_sinc_window / buffer_helper / VNA_decimator_helper run on the CPU."""
    lib = ref_lib()
    for L, fc in [(8192, 1.0 / 4096), (400, 0.75 / 200), (401, 0.01), (300, 1.0 / 200), (5000, 0.75 / 2500)]:
        fc = float(np.float32(fc))
        w = np.empty(L, dtype=np.float32)
        lib.gsdr_ref_make_sinc_window(L, C.c_float(fc), w.ctypes.data_as(C.c_void_p))
        assert np.array_equal(HL.make_sinc_window(L, fc), w) and np.array_equal(orc.make_sinc_window(L, fc), w)
    rng = np.random.default_rng(10)
    for _ in range(40):
        N, P = int(rng.integers(2, 5000)), int(rng.integers(1, 9))
        L, T = int(rng.integers(max(N, 1000), 2_000_000)), int(rng.integers(1, 50))
        out = np.empty((12, 6), dtype=np.int32)
        lib.gsdr_ref_buffer_helper_seq(N, L, P, T, 12, out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(HL.buffer_helper_sequence(N, L, P, T, 12), out)
        ppt = int(rng.integers(1, 3_000_000))
        out = np.empty((12, 4), dtype=np.int32)
        lib.gsdr_ref_vna_helper_seq(ppt, L, 12, out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(HL.vna_helper_sequence(ppt, L, 12), out)


@pytest.mark.parametrize("T,seed,order", [(1000, 1, "random"), (1000, 2, "sorted"), (1024, 3, "random"), (37, 4, "random"),
                                         (2048, 5, "all"), (600, 6, "dups"), (1, 7, "random")])
def test_pfb_gather_layout_is_conflict_free(T, seed, order):
    """Bipartite edge colouring behind the fused kernel's tone gather: every 16-bin row holds a permutation
    of the 16 slots, and every run of 16 consecutive tones reads 16 different slots (= bank pairs)."""
    from gpu_sdr_b200 import hostlogic as hl
    rng = np.random.default_rng(seed)
    if order == "all":
        bins = None
        sel = np.arange(2048)
    else:
        sel = rng.choice(2048, size=T, replace=False)
        if order == "sorted":
            sel = np.sort(sel)
        if order == "dups":
            sel[100:130] = sel[0:30]  # repeated tones: first occurrence defines the slot
        bins = sel.astype(np.int32)
    pos = hl.pfb_gather_layout(bins, T)
    assert pos.shape == (2048,) and pos.max() < 16
    b = np.arange(2048)
    rows = (b & 7) * 16 + (b >> 7)
    for r in range(128):
        assert sorted(pos[rows == r]) == list(range(16))
    clashes = 0
    for h in range(0, len(sel), 16):
        grp = np.unique(sel[h:h + 16])  # duplicates inside one read are a broadcast
        clashes += len(grp) - len(set(pos[grp]))
    # repeated tones are pinned first and everything else is coloured around them; a clash is only possible
    # next to a pinned bin and should not occur on a list like this one
    assert clashes == 0, clashes


def test_packet_header_matches_the_client_dtype():
    """The 21-byte data-socket header (Sync_server::format_net_buffer, cpp/USRP_server_network.cpp:164-191) as the
    client parses it (pyUSRP/USRP_low_level.py:63-70: a packed numpy dtype)."""
    import ctypes as C
    from gpu_sdr_b200 import _lib
    header_type = np.dtype([("usrp_number", np.int32), ("front_end_code", np.dtype("|S1")), ("packet_number", np.int32),
                            ("length", np.int32), ("errors", np.int32), ("channels", np.int32)])
    assert header_type.itemsize == 21
    lib = _lib.load()
    pkt = _lib.RxPacket()
    pkt.usrp_number, pkt.front_end_code, pkt.packet_number = 3, b"B", 123456
    pkt.length, pkt.errors, pkt.channels = 488000, -2, 1000
    raw = np.zeros(21, dtype=np.uint8)
    assert lib.gsdr_packet_header_write(C.byref(pkt), raw.ctypes.data_as(C.c_void_p)) == 21
    h = np.frombuffer(raw.tobytes(), dtype=header_type)[0]
    assert (h["usrp_number"], h["front_end_code"], h["packet_number"], h["length"], h["errors"], h["channels"]) == \
        (3, b"B", 123456, 488000, -2, 1000)
    back = _lib.RxPacket()
    assert lib.gsdr_packet_header_read(raw.ctypes.data_as(C.c_void_p), C.byref(back)) == 21
    assert (back.usrp_number, back.front_end_code, back.packet_number, back.length, back.errors, back.channels) == \
        (3, b"B", 123456, 488000, -2, 1000)


@pytest.mark.parametrize("frames,grid", [([488] * 64, 148), ([3904] * 8, 148), ([488], 148), ([5, 0, 700, 3, 0, 1200], 148),
                                         ([488] * 3, 5), ([1] * 200, 148), ([97, 488, 2929], 148), ([3904] * 64, 148), ([488] * 2, 148)])
def test_multi_stream_tile_partition(frames, grid):
    """gsdr_pfb_partition (the cut of a gsdr_rx_group launch): every frame of every stream exactly once and in order, tiles never
    span streams, each stream's last tile flagged (it carries the carry-over copy), CTA ranges contiguous, costs balanced."""
    import ctypes as C
    lib = g.load()
    nf = np.array(frames, dtype=np.int32)
    cap = 2 * grid + 2 * len(frames) + 8
    tiles = np.zeros((cap, 4), dtype=np.int32)
    cb = np.zeros(grid + 1, dtype=np.int32)
    n = lib.gsdr_pfb_partition(nf.ctypes.data_as(C.c_void_p), len(frames), grid, tiles.ctypes.data_as(C.c_void_p), cap, cb.ctypes.data_as(C.c_void_p))
    assert n > 0
    tiles = tiles[:n]
    assert cb[0] == 0 and cb[grid] == n and np.all(np.diff(cb) >= 0)
    nxt = {j: 0 for j, f in enumerate(frames) if f > 0}
    order = []
    for job, fa, fb, flags in tiles:
        assert fa == nxt[job] and fb > fa and fb <= frames[job]
        nxt[job] = fb
        assert (flags & 1) == (1 if fb == frames[job] else 0)
        order.append(job)
    assert all(nxt[j] == frames[j] for j in nxt) and order == sorted(order)
    cost = [sum(24 + (tiles[t][2] - tiles[t][1]) for t in range(cb[c], cb[c + 1])) for c in range(grid)]
    busy = [c for c in cost if c > 0]
    if sum(frames) >= 40 * grid:   # enough work per CTA for the balance to mean something
        assert len(busy) == grid and max(busy) <= 1.12 * (sum(busy) / len(busy)), (max(busy), sum(busy) / len(busy))


def _simulate_form(busy, copied, zero_copy, lag=2):
    lib = g.load()
    b = np.asarray(busy, dtype=np.uint8)
    out = np.zeros(len(b), dtype=np.int8)
    kept = lib.gsdr_group_form_simulate(b.ctypes.data_as(C.c_void_p), len(b), float(copied), float(zero_copy), int(lag),
                                        out.ctypes.data_as(C.c_void_p))
    return kept, out.tolist()


def test_group_form_paced_caller_stays_zero_copy():
    """gsdr_rx_group_submit's default form (GroupAutoForm): a caller whose previous period is always done never measures."""
    kept, forms = _simulate_form([0] * 100, copied=1e-3, zero_copy=2e-3)
    assert kept == -1 and forms == [1] * 100


@pytest.mark.parametrize("copied,zero_copy,want", [(1.0e-3, 1.3e-3, 0), (1.3e-3, 1.0e-3, 1), (1.0e-3, 1.0e-3, 1), (0.98e-3, 1.0e-3, 1),
                                                   (0.96e-3, 1.0e-3, 0)])
def test_group_form_saturated_caller_measures_and_keeps_the_faster(copied, zero_copy, want):
    busy = [0] + [1] * 60                      # the first submit has no predecessor
    kept, forms = _simulate_form(busy, copied, zero_copy)
    assert kept == want
    # submit 0 plus three busy ones, then (from the fourth busy submit) 8 zero-copy and 8 copied periods, then the form kept
    assert forms[:12] == [1] * 12 and forms[12:20] == [0] * 8 and forms[20:] == [want] * (len(busy) - 20)


def test_group_form_drained_pipeline_abandons_the_measurement():
    busy = [0] + [1] * 14 + [0] + [1] * 40     # the pipeline drains in the middle of the copied block
    kept, forms = _simulate_form(busy, copied=1.0e-3, zero_copy=2.0e-3)
    assert forms[12:15] == [0] * 3 and forms[15:19] == [1] * 4          # back to zero-copy, counting busy submits again
    assert forms[19:27] == [1] * 8 and forms[27:35] == [0] * 8          # a fresh measurement
    assert kept == 0 and forms[35:] == [0] * (len(busy) - 35)


def test_group_form_without_wait_reports_keeps_zero_copy():
    kept, forms = _simulate_form([0] + [1] * 40, copied=1e-3, zero_copy=2e-3, lag=1000)   # the caller never waits in time
    assert kept == 1 and forms[20:] == [1] * (len(forms) - 20)
