"""gpu_sdr_b200 -- B200-native (sm_100a) RX/TX readout DSP path of GPU_SDR.

The product is the C++/CUDA shared library ``libgsdr.so`` (C-ABI in ``include/gsdr.h``); this
package is its thin host-side mirror of the reference's buffer-wrapper interface
(``RX_buffer_demodulator``, ``TX_buffer_generator``, ``preallocator``, ``param``).  No PyTorch, no
CPU fallback: importing works anywhere the library is built, computing needs a CUDA device.
"""
from ._lib import GsdrError, LIB_PATH, load  # noqa: F401
from .params import (param, W_TYPES, TONES, CHIRP, NOISE, RAMP, NODSP, SWONLY, DIRECT,  # noqa: F401
                     string_to_w_type, w_type_to_str)
from .demodulator import (RX_buffer_demodulator, TX_buffer_generator, preallocator, ReplaySource, RxGroup,  # noqa: F401
                          DeviceBuffer, pinned_empty, pinned_free)
from . import hostlogic  # noqa: F401

__all__ = ["param", "RX_buffer_demodulator", "TX_buffer_generator", "preallocator", "ReplaySource", "RxGroup",
           "DeviceBuffer", "pinned_empty", "pinned_free", "hostlogic", "GsdrError", "load"]
