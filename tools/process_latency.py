"""Blocking process() time per transport buffer for cfg2 (TONES, N=2048, P=4, T=1000, buffer_len 1e6), pinned host in/out.
One JSON line; run once per environment setting (the switches are read when the library first needs them), e.g.
    GSDR_PROCESS_ZEROCOPY=1 GSDR_PFB_MIN_TILE=16 python tools/process_latency.py
Also checks the outputs of the run against the default path's (bit-identical: same kernel, same frames)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_sdr_b200 as g  # noqa: E402
from common import pfb_param  # noqa: E402

L = int(os.environ.get("LAT_BUFLEN", 1_000_000))
REPS = int(os.environ.get("LAT_REPS", 300))
p = pfb_param(L=L)
rx = g.RX_buffer_demodulator(p)
rng = np.random.default_rng(3)
hin = [g.pinned_empty(L) for _ in range(4)]
for h in hin:
    h[:] = (rng.standard_normal(L) + 1j * rng.standard_normal(L)).astype(np.complex64) * 0.1
hout = [g.pinned_empty(rx.max_output()) for _ in range(4)]
for i in range(20):
    rx.process(hin[i % 4], hout[i % 4])
t0 = time.perf_counter()
for i in range(REPS):
    rx.process(hin[i % 4], hout[i % 4])
dt = (time.perf_counter() - t0) / REPS
rx.reset()
chk = 0.0
for i in range(6):
    n = rx.process(hin[i % 4], hout[i % 4])
    chk += float(np.abs(hout[i % 4][:n]).astype(np.float64).sum()) * (i + 1)
rx.close()
print(json.dumps({"zerocopy": os.environ.get("GSDR_PROCESS_ZEROCOPY", "1 (default)"), "min_tile": os.environ.get("GSDR_PFB_MIN_TILE", "auto"), "ptr_cache": os.environ.get("GSDR_PROCESS_PTRCACHE", "0"),
                  "chunks": os.environ.get("GSDR_PROCESS_CHUNKS", "default"), "buffer_len": L, "us_per_buffer": dt * 1e6,
                  "MS_per_s": L / dt / 1e6, "checksum": repr(chk)}))
