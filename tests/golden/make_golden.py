"""Generates the golden fixtures under tests/golden/ by running the REFERENCE's own object code
(oracle/_ref/libgsdr_ref.so: the unmodified /root/reference sources compiled for sm_100a) on a
B200.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

and copy gpurun_out/golden/*.npz into tests/golden/.  Inputs are stored as int16 IQ pairs (the
sc16 grid, exact in float32), reference outputs as complex64, parameters as JSON.  These files
pin the CPU oracle (oracle/) and are the reference side of the GPU parity tests; nothing here is
product code.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from common import RefRX, RefTX, g, iq_to_int16, orc, ref_lib, tone_stream  # noqa: E402


def pjson(p):
    return json.dumps({k: getattr(p, k) for k in (
        "rate", "fft_tones", "pf_average", "buffer_len", "decim", "freq", "wave_type", "ampl", "chirp_t", "chirp_f", "swipe_s")})


def rx_case(out_dir, name, p, bufs, cap):
    ref = RefRX(p)
    outs = [ref.process(x, cap) for x in bufs]
    extra = {}
    if p.wave_type[0] == g.TONES:
        extra["bins"] = ref.bins(len(p.freq))
        extra["batching"] = np.int32(ref.lib.gsdr_ref_rx_batching(ref.h))
        w = np.empty(p.fft_tones * p.pf_average, dtype=np.float32)
        ref.lib.gsdr_ref_rx_window(ref.h, w.ctypes.data_as(C.c_void_p), w.size)
        extra["window"] = w
    if p.wave_type[0] == g.CHIRP:
        cp = orc.ChirpParam()
        ref.lib.gsdr_ref_rx_chirp_param(ref.h, C.byref(cp))
        extra["chirp_param"] = np.array([cp.num_steps, cp.length, cp.chirpness, cp.f0], dtype=np.int64)
    ref.close()
    np.savez_compressed(
        os.path.join(out_dir, name + ".npz"), param=pjson(p), n_buffers=len(bufs),
        inputs=np.stack([iq_to_int16(x) for x in bufs]), lengths=np.array([len(o) for o in outs], dtype=np.int64),
        outputs=np.concatenate(outs) if outs else np.zeros(0, np.complex64), **extra)
    print(name, [len(o) for o in outs])


def tx_case(out_dir, name, p, n):
    ref = RefTX(p)
    outs = [ref.get() for _ in range(n)]
    extra = {}
    if p.wave_type[0] == g.CHIRP:
        cp = orc.ChirpParam()
        ref.lib.gsdr_ref_tx_chirp_param(ref.h, C.byref(cp))
        extra["chirp_param"] = np.array([cp.num_steps, cp.length, cp.chirpness, cp.f0], dtype=np.int64)
    ref.close()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), param=pjson(p), n_buffers=n, outputs=np.stack(outs), **extra)
    print(name, outs[0][:2])


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    lib = ref_lib()
    assert lib is not None and lib.gsdr_ref_device_count() > 0, "needs oracle/_ref/libgsdr_ref.so and a GPU"

    # ---- taps and integer helpers -----------------------------------------------------------------
    host = {}
    for (L, fc) in [(8192, np.float32(1.0 / 4096)), (400, np.float32(0.75 / 200)), (401, np.float32(0.01)), (256, np.float32(1.0 / 128)),
                    (300, np.float32(1.0 / 200)), (80, np.float32(0.75 / 20)), (7, np.float32(0.2))]:
        w = np.empty(L, dtype=np.float32)
        lib.gsdr_ref_make_sinc_window(L, C.c_float(fc), w.ctypes.data_as(C.c_void_p))
        host[f"sinc_{L}_{float(fc)!r}"] = w
    for (L, side) in [(2000, 200), (6000, 600), (200, 20), (7, 0), (1, 0), (33, 3)]:
        w = np.empty(L, dtype=np.float32)
        assert lib.gsdr_ref_make_flat_window(L, side, w.ctypes.data_as(C.c_void_p)) == L
        host[f"flat_{L}_{side}"] = w
    for (N, L, P, T) in [(2048, 1_000_000, 4, 1000), (100, 50_000, 3, 7), (64, 20_000, 4, 8), (10, 1_000_000, 1, 3), (4096, 5_999_999, 8, 5)]:
        out = np.empty((16, 6), dtype=np.int32)
        lib.gsdr_ref_buffer_helper_seq(N, L, P, T, 16, out.ctypes.data_as(C.c_void_p))
        host[f"bh_{N}_{L}_{P}_{T}"] = out
    for (ppt, L) in [(2000, 1_000_000), (6007, 1_000_000), (3, 100_000), (200, 100_000), (1_500_000, 1_000_000), (1, 50_000)]:
        out = np.empty((16, 4), dtype=np.int32)
        lib.gsdr_ref_vna_helper_seq(ppt, L, 16, out.ctypes.data_as(C.c_void_p))
        host[f"vna_{ppt}_{L}"] = out
    np.savez_compressed(os.path.join(out_dir, "host_logic.npz"), **host)

    # ---- tone -> bin map on the full cfg2 tone set plus edge cases ----------------------------------
    rate, N = 200_000_000, 2048
    edge = [292968, -292968, 1, -1, 0, 97656, 97657, 195312, 195313, -97656, -97657, 99_999_999, -99_999_999, 48828, -48829]
    p = g.param(rate=rate, fft_tones=N, pf_average=1, buffer_len=50_000, freq=edge, wave_type=[g.TONES] * len(edge))
    r = RefRX(p)
    bins_edge = r.bins(len(edge))
    r.close()
    rng = np.random.default_rng(1337)
    ks = rng.choice(np.arange(-N // 2 + 1, N // 2), size=1000, replace=False)
    full = [int(k * (rate / N)) for k in ks]
    p = g.param(rate=rate, fft_tones=N, pf_average=1, buffer_len=50_000, freq=full, wave_type=[g.TONES] * len(full))
    r = RefRX(p)
    bins_full = r.bins(len(full))
    r.close()
    odd = [12345, -4321, 499_999, -499_999, 3, 250_000]
    p = g.param(rate=1_000_000, fft_tones=100, pf_average=1, buffer_len=50_000, freq=odd, wave_type=[g.TONES] * len(odd))
    r = RefRX(p)
    bins_odd = r.bins(len(odd))
    r.close()
    np.savez_compressed(os.path.join(out_dir, "tone_bins.npz"), edge_freq=np.array(edge), edge_bins=bins_edge,
                        full_freq=np.array(full), full_bins=bins_full, odd_freq=np.array(odd), odd_bins=bins_odd)

    # ---- RX chains ---------------------------------------------------------------------------------
    def tones_case(name, rate, N, P, T, L, nbuf, seed):
        rng = np.random.default_rng(seed)
        ks = rng.choice(np.arange(-N // 2 + 1, N // 2), size=T, replace=False)
        freq = [int(k * (rate / N)) for k in ks]
        p = g.param(mode="RX", rate=rate, fft_tones=N, pf_average=P, buffer_len=L, freq=freq, wave_type=[g.TONES] * T,
                    ampl=[1.0 / T] * T)
        bufs = [tone_stream(rate, freq, p.ampl, i * L, L, seed=seed) for i in range(nbuf)]
        rx_case(out_dir, name, p, bufs, T * orc.pfb_batching(L, N, P))

    tones_case("rx_tones_n64", 1_000_000, 64, 4, 8, 10_000, 4, 11)
    tones_case("rx_tones_n100", 1_000_000, 100, 3, 7, 10_000, 4, 12)
    tones_case("rx_tones_n2048", 200_000_000, 2048, 4, 24, 40_000, 3, 13)

    def direct_case(name, rate, freq, decim, f, L, nbuf, seed):
        T = len(freq)
        p = g.param(mode="RX", rate=rate, decim=decim, pf_average=f, buffer_len=L, freq=freq, wave_type=[g.DIRECT] * T,
                    ampl=[1.0 / T] * T)
        bufs = [tone_stream(rate, freq, p.ampl, i * L, L, seed=seed) for i in range(nbuf)]
        rx_case(out_dir, name, p, bufs, T * L // max(decim, 1) + 16)

    direct_case("rx_direct_decim10", 1_000_000, [12345, -4321, 400_001, -499_999, 7], 10, 8, 10_000, 3, 21)
    direct_case("rx_direct_decim100", 100_000_000, [12_345_677, -49_999_999, 1, -25_000_000], 100, 4, 10_000, 3, 22)
    direct_case("rx_direct_nodecim", 1_000_000, [12345, -4321, 499_999], 0, 1, 4096, 2, 23)

    def chirp_case(name, rate, f0, f1, steps, t, decim, L, nbuf, seed):
        p = g.param(mode="RX", rate=rate, decim=decim, buffer_len=L, freq=[f0], chirp_f=[f1], swipe_s=[steps], chirp_t=[t],
                    wave_type=[g.CHIRP], ampl=[1.0])
        gen = orc.ChirpGenerator(rate, f0, f1, steps, t, 1.0, L)
        rng = np.random.default_rng(seed)
        bufs = []
        for i in range(nbuf):
            s21 = 0.5 * np.exp(2j * np.pi * 0.07 * i)
            from common import quantize_iq
            bufs.append(quantize_iq(gen.get() * s21 + 1e-3 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))))
        rx_case(out_dir, name, p, bufs, L + 16)

    chirp_case("rx_chirp_lockin", 200_000_000, -50_000_000, 50_000_000, 1000, 0.01, 3, 10_000, 4, 31)
    chirp_case("rx_chirp_full", 200_000_000, -50_000_000, 50_000_000, 1000, 0.01, 0, 10_000, 2, 32)
    chirp_case("rx_chirp_true", 200_000_000, 10_000_000, -30_000_000, 0, 0.001, 200, 10_000, 3, 33)

    # ---- TX ----------------------------------------------------------------------------------------
    p = g.param(mode="TX", rate=100_000, buffer_len=10_000, freq=[1000, -2500, 33_333, -49_999, 77, 1000],
                ampl=[0.2, 0.1, 0.3, 0.15, 0.05, 0.25], wave_type=[g.TONES] * 6)  # duplicate 1000: last wins
    # (freq == 0 is left out on purpose: the reference then writes base_vector[rate], one element past its malloc)
    tx_case(out_dir, "tx_tones", p, 12)  # 12 x 10000 wraps the 100000-sample period
    p = g.param(mode="TX", rate=30_000, buffer_len=50_000, freq=[1000, -2500], ampl=[0.5, 0.25], wave_type=[g.TONES] * 2)
    tx_case(out_dir, "tx_tones_long", p, 3)  # buffer_len > rate: replicated period
    p = g.param(mode="TX", rate=200_000_000, buffer_len=10_000, freq=[-50_000_000], chirp_f=[50_000_000], swipe_s=[1000],
                chirp_t=[0.01], wave_type=[g.CHIRP], ampl=[0.7])
    tx_case(out_dir, "tx_chirp", p, 3)
    p = g.param(mode="TX", rate=200_000_000, buffer_len=10_000, freq=[10_000_000], chirp_f=[-30_000_000], swipe_s=[0],
                chirp_t=[0.001], wave_type=[g.CHIRP], ampl=[1.0])
    tx_case(out_dir, "tx_chirp_true", p, 3)
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join("gpurun_out", "golden"))
