import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import gpu_sdr_b200 as g
from common import direct_param
p = direct_param(L=1_000_000)
rx = g.RX_buffer_demodulator(p)
nb = 8
rng = np.random.default_rng(0)
x = (rng.standard_normal(p.buffer_len) + 1j*rng.standard_normal(p.buffer_len)).astype(np.complex64)*0.1
d = g.DeviceBuffer(nb*p.buffer_len)
for b in range(nb): d.upload(x, offset=b*p.buffer_len)
out = g.DeviceBuffer(rx.max_output_batch(nb))
for i in range(4):
    rx.process_device(d.ptr, nb, out.ptr)
rx.sync()
print(rx.kernel_name())
