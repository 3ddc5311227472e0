"""Per-role cycle counters of direct_fir_i8_kernel on cfg1 (64 buffers per launch and one buffer per launch).
GSDR_DIRECT_I8_DEBUG=1 python tools/direct_i8_debug.py   (prints to stderr; synchronises every launch)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["GSDR_DIRECT_VARIANT"] = "i8"
os.environ["GSDR_DIRECT_I8_DEBUG"] = "1"
import gpu_sdr_b200 as g  # noqa: E402
from common import direct_param  # noqa: E402

for kw, n_buf in ((dict(), 64), (dict(), 1), (dict(T=64), 16)):
    p = direct_param(**kw)
    L = p.buffer_len
    rng = np.random.default_rng(1)
    x = (0.1 * (rng.standard_normal(L) + 1j * rng.standard_normal(L))).astype(np.complex64)
    rx = g.RX_buffer_demodulator(p)
    d = g.DeviceBuffer(n_buf * L)
    for b in range(n_buf):
        d.upload(x, offset=b * L)
    out = g.DeviceBuffer(rx.max_output_batch(n_buf))
    print(f"--- T={len(p.freq)} decim={p.decim} buffers={n_buf}", file=sys.stderr, flush=True)
    for _ in range(3):
        rx.process_device(d.ptr, n_buf, out.ptr)
        rx.sync()
    rx.close()
