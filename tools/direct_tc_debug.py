#!/usr/bin/env python
"""Per-role wait/run cycles of the tensor-core DIRECT kernel (GSDR_DIRECT_TC_DEBUG=1 makes every launch print them)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["GSDR_DIRECT_VARIANT"] = "tc"
import gpu_sdr_b200 as g  # noqa: E402
from common import direct_param  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
p = direct_param()
rx = g.RX_buffer_demodulator(p)
rng = np.random.default_rng(0)
x = (rng.standard_normal(p.buffer_len) + 1j * rng.standard_normal(p.buffer_len)).astype(np.complex64) * 0.1
d = g.DeviceBuffer(nb * p.buffer_len)
for b in range(nb):
    d.upload(x, offset=b * p.buffer_len)
out = g.DeviceBuffer(rx.max_output_batch(nb))
for i in range(3):
    rx.process_device(d.ptr, nb, out.ptr)
rx.sync()
os.environ["GSDR_DIRECT_TC_DEBUG"] = "1"
for i in range(3):
    rx.process_device(d.ptr, nb, out.ptr)
rx.sync()
