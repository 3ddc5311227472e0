"""JSON parameter fields and enum spellings stay those of the reference
(cpp/USRP_JSON_interpreter.cpp:28-251, cpp/USRP_server_settings.cpp:7-54, chk_param clamps :268-438)."""
import json

from common import g
from gpu_sdr_b200 import params


def test_w_type_order_and_strings():
    assert params.W_TYPES == ("TONES", "CHIRP", "NOISE", "RAMP", "NODSP", "SWONLY", "DIRECT")
    for i, s in enumerate(params.W_TYPES):
        assert params.w_type_to_str(i) == s
    for s in ("TONES", "CHIRP", "NOISE", "NODSP", "SWONLY", "DIRECT"):
        assert params.w_type_to_str(params.string_to_w_type(s)) == s
    # the reference's string_to_w_type has no RAMP branch and defaults to NODSP
    assert params.string_to_w_type("RAMP") == g.NODSP
    assert params.string_to_w_type("bogus") == g.NODSP


def test_json_antenna_object():
    cmd = {"A_RX2": {"mode": "RX", "rf": 3e8, "tuning_mode": 0, "rate": 2e8, "decim": 0, "fft_tones": 2048, "pf_average": 4,
                     "samples": 2e9, "buffer_len": 1e6, "burst_off": 0, "burst_on": 0, "bw": 4e8, "delay": 1, "gain": 0,
                     "freq": [1e6, -2e6], "ampl": [0.5, 0.5], "wave_type": ["TONES", "TONES"], "chirp_t": [0, 0],
                     "chirp_f": [0, 0], "swipe_s": [0, 0], "data_mem_mult": 1}, "device": 0}
    p = g.param.from_json(json.dumps(cmd), "A_RX2")
    assert p.rate == 200_000_000 and p.fft_tones == 2048 and p.pf_average == 4 and p.buffer_len == 1_000_000
    assert p.freq == [1_000_000, -2_000_000] and p.wave_type == [g.TONES, g.TONES]
    assert not p.dynamic_buffer()


def test_chk_param_clamps():
    p = g.param.from_json_obj({"rate": 1e6, "fft_tones": 1, "pf_average": 0, "buffer_len": 10, "wave_type": ["CHIRP"]})
    assert p.fft_tones == 2 and p.pf_average == 1 and p.buffer_len == params.DEFAULT_BUFFER_LEN
    assert p.dynamic_buffer()
    p = g.param.from_json_obj({"buffer_len": 7e6})
    assert p.buffer_len == params.DEFAULT_BUFFER_LEN
