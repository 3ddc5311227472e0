"""``param``: the per-antenna parameter block, mirroring the reference's struct field for field
(headers/USRP_server_settings.hpp:130-167) and its JSON spelling (cpp/USRP_JSON_interpreter.cpp:28-251,
client defaults pyUSRP/USRP_files.py:449-479).  Only the DSP-layer fields reach the C-ABI."""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass, field
from typing import List

from ._lib import CParam

# enum w_type { TONES, CHIRP, NOISE, RAMP, NODSP, SWONLY, DIRECT } (headers/USRP_server_settings.hpp:113)
W_TYPES = ("TONES", "CHIRP", "NOISE", "RAMP", "NODSP", "SWONLY", "DIRECT")
TONES, CHIRP, NOISE, RAMP, NODSP, SWONLY, DIRECT = range(7)

MIN_USEFULL_BUFFER = 50000      # headers/USRP_server_settings.hpp:98-102
MAX_USEFULL_BUFFER = 6000000
DEFAULT_BUFFER_LEN = 1000000


def string_to_w_type(s: str) -> int:
    """cpp/USRP_server_settings.cpp:38-54: unknown strings (and "RAMP") map to NODSP."""
    return {"NODSP": NODSP, "CHIRP": CHIRP, "NOISE": NOISE, "TONES": TONES, "SWONLY": SWONLY, "DIRECT": DIRECT}.get(s, NODSP)


def w_type_to_str(w: int) -> str:
    return W_TYPES[w] if 0 <= w < len(W_TYPES) else "UNINIT"


@dataclass
class param:
    mode: str = "OFF"
    rate: int = 0
    gain: int = 0
    bw: int = 0
    rf: float = 0.0
    samples: int = 0
    delay: float = 0.0
    burst_on: float = 0.0
    burst_off: float = 0.0
    buffer_len: int = DEFAULT_BUFFER_LEN
    tuning_mode: int = 0
    freq: List[int] = field(default_factory=list)
    wave_type: List[int] = field(default_factory=list)
    ampl: List[float] = field(default_factory=list)
    decim: int = 0
    chirp_t: List[float] = field(default_factory=list)
    chirp_f: List[int] = field(default_factory=list)
    swipe_s: List[int] = field(default_factory=list)
    data_mem_mult: int = 1
    fft_tones: int = 0
    pf_average: int = 1

    def dynamic_buffer(self) -> bool:
        """cpp/USRP_server_settings.cpp:98-102."""
        return any(w != TONES for w in self.wave_type)

    @classmethod
    def from_json_obj(cls, obj: dict) -> "param":
        """One antenna object (A_TXRX / A_RX2 / B_TXRX / B_RX2) of the client's JSON command.
        Applies the same clamps as chk_param (cpp/USRP_JSON_interpreter.cpp:268-438) for the
        fields the DSP layer reads."""
        p = cls()
        p.mode = str(obj.get("mode", "OFF"))
        for k in ("rate", "gain", "bw", "samples", "buffer_len", "decim", "fft_tones", "pf_average", "data_mem_mult",
                  "tuning_mode"):
            if k in obj:
                setattr(p, k, int(float(obj[k])))
        for k in ("rf", "delay", "burst_on", "burst_off"):
            if k in obj:
                setattr(p, k, float(obj[k]))
        p.freq = [int(float(v)) for v in obj.get("freq", [])]
        p.chirp_f = [int(float(v)) for v in obj.get("chirp_f", [])]
        p.swipe_s = [int(float(v)) for v in obj.get("swipe_s", [])]
        p.ampl = [float(v) for v in obj.get("ampl", [])]
        p.chirp_t = [float(v) for v in obj.get("chirp_t", [])]
        p.wave_type = [string_to_w_type(str(v)) for v in obj.get("wave_type", [])]
        if p.pf_average < 1:
            p.pf_average = 1
        if p.fft_tones < 2:
            p.fft_tones = 2
        if p.buffer_len < MIN_USEFULL_BUFFER or p.buffer_len > MAX_USEFULL_BUFFER:
            p.buffer_len = DEFAULT_BUFFER_LEN
        return p

    @classmethod
    def from_json(cls, text: str, antenna: str) -> "param":
        return cls.from_json_obj(json.loads(text)[antenna])

    def to_c(self):
        """Returns (CParam, keepalive) -- keep `keepalive` referenced while the CParam is in use."""
        def arr(ctype, vals):
            a = (ctype * max(len(vals), 1))(*vals)
            return a, len(vals)

        freq, nf = arr(C.c_int32, [int(v) for v in self.freq])
        ampl, na = arr(C.c_float, [float(v) for v in self.ampl])
        wt, nw = arr(C.c_int32, [int(v) for v in self.wave_type])
        ct, nct = arr(C.c_float, [float(v) for v in self.chirp_t])
        cf, ncf = arr(C.c_int32, [int(v) for v in self.chirp_f])
        ss, nss = arr(C.c_int32, [int(v) for v in self.swipe_s])
        cp = CParam(
            rate=int(self.rate), fft_tones=int(self.fft_tones), decim=int(self.decim), pf_average=int(self.pf_average),
            buffer_len=int(self.buffer_len), data_mem_mult=int(self.data_mem_mult), samples=int(self.samples),
            freq=C.cast(freq, C.POINTER(C.c_int32)), n_freq=nf, ampl=C.cast(ampl, C.POINTER(C.c_float)), n_ampl=na,
            wave_type=C.cast(wt, C.POINTER(C.c_int32)), n_wave_type=nw,
            chirp_t=C.cast(ct, C.POINTER(C.c_float)), n_chirp_t=nct,
            chirp_f=C.cast(cf, C.POINTER(C.c_int32)), n_chirp_f=ncf,
            swipe_s=C.cast(ss, C.POINTER(C.c_int32)), n_swipe_s=nss)
        return cp, (freq, ampl, wt, ct, cf, ss)
