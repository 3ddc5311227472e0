"""Host-side mirror of the reference's buffer-wrapper classes, over the C-ABI.

Same names, argument meaning and error behaviour as the reference so that callers (and the
parity tests) read like the reference's own call sites:

    RX_buffer_demodulator(param, diagnostic=False).process(in, out) -> valid length ; .close()
        headers/USRP_demodulator.hpp:13-33, called from cpp/USRP_server_link_threads.cpp:121,666,475
    TX_buffer_generator(param).get(buf) -> buffer ; .close()
        headers/USRP_buffer_generator.hpp:49-68, called from cpp/USRP_server_link_threads.cpp:191,584,510
    preallocator(vector_size, pipe_size).get() / .trash(buf) / .close()
        headers/USRP_server_memory_management.hpp:103-273

Buffers are numpy complex64 arrays (== float2).  Pinned buffers come from ``preallocator`` or
``pinned_empty``; pageable arrays work too (CUDA stages them) but do not reach PCIe speed.
Unsupported configurations raise ``GsdrError`` with the reference's message (the reference
prints it and calls exit(-1)).  There is no CPU path: without libgsdr.so / a CUDA device every
constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import GsdrError, check
from .params import param


def _ptr(a: np.ndarray):
    if a.dtype != np.complex64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("buffers must be C-contiguous complex64 arrays")
    return a.ctypes.data_as(C.c_void_p)


def _ptr_i16(a: np.ndarray):
    if a.dtype != np.int16 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("sc16 buffers must be C-contiguous int16 arrays (interleaved I/Q)")
    return a.ctypes.data_as(C.c_void_p)


class _Pinned:
    """Owner of one cudaMallocHost block exposed as a complex64 numpy array."""

    def __init__(self, n):
        self.lib = _lib.load()
        self.ptr = self.lib.gsdr_host_alloc(int(n) * 8)
        if not self.ptr:
            raise GsdrError("gsdr_host_alloc: " + _lib.last_error())
        self.n = int(n)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.gsdr_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


_pinned_owners = {}


def pinned_empty(n: int) -> np.ndarray:
    """complex64[n] in pinned host memory (kept alive until ``pinned_free`` or interpreter exit)."""
    own = _Pinned(n)
    a = np.ctypeslib.as_array(C.cast(own.ptr, C.POINTER(C.c_float)), shape=(2 * own.n,)).view(np.complex64)
    _pinned_owners[a.ctypes.data] = own
    return a


def pinned_free(a: np.ndarray) -> None:
    _pinned_owners.pop(a.ctypes.data, None)


class RX_buffer_demodulator:
    def __init__(self, init_parameters: param, init_diagnostic: bool = False, device: int = 0):
        self.lib = _lib.load()
        self.parameters = init_parameters
        self._c, self._keep = init_parameters.to_c()
        self._h = self.lib.gsdr_rx_create(C.byref(self._c), int(device), int(bool(init_diagnostic)))
        if not self._h:
            raise GsdrError(_lib.last_error())
        self.device = int(device)
        self.fcut = float(self.lib.gsdr_rx_fcut(self._h))
        self.channels = int(self.lib.gsdr_rx_channels(self._h))

    # -- the reference interface -----------------------------------------------------------------
    def process(self, input_buffer: np.ndarray, output_buffer: np.ndarray) -> int:
        """Blocking; returns the number of valid complex64 in output_buffer (sample-major)."""
        if input_buffer.size < self.parameters.buffer_len:
            raise ValueError("input buffer shorter than buffer_len")
        if output_buffer.size < self.max_output():
            raise ValueError(f"output buffer holds {output_buffer.size} < {self.max_output()} samples")
        return check(self.lib.gsdr_rx_process(self._h, _ptr(input_buffer), _ptr(output_buffer)), "gsdr_rx_process")

    def close(self) -> None:
        if self._h:
            self.lib.gsdr_rx_destroy(self._h)
            self._h = None

    # -- pipelined / device-resident extensions ----------------------------------------------------
    def submit(self, input_buffer, output_buffer):
        n = C.c_int(0)
        t = check(self.lib.gsdr_rx_submit(self._h, _ptr(input_buffer), _ptr(output_buffer), C.byref(n)), "gsdr_rx_submit")
        return t, n.value

    def wait(self, ticket: int) -> None:
        check(self.lib.gsdr_rx_wait(self._h, int(ticket)), "gsdr_rx_wait")

    # -- sc16 ingest: the USRP wire format goes to the GPU as is --------
    def process_sc16(self, iq: np.ndarray, output_buffer: np.ndarray) -> int:
        """iq: int16 array of 2*buffer_len interleaved I/Q.  Same result as process(iq_as_float / 32767)."""
        if iq.dtype != np.int16 or iq.size < 2 * self.parameters.buffer_len:
            raise ValueError("iq must be int16 with 2*buffer_len elements")
        if output_buffer.size < self.max_output():
            raise ValueError(f"output buffer holds {output_buffer.size} < {self.max_output()} samples")
        return check(self.lib.gsdr_rx_process_sc16(self._h, _ptr_i16(iq), _ptr(output_buffer)), "gsdr_rx_process_sc16")

    def submit_sc16(self, iq, output_buffer):
        n = C.c_int(0)
        t = check(self.lib.gsdr_rx_submit_sc16(self._h, _ptr_i16(iq), _ptr(output_buffer), C.byref(n)), "gsdr_rx_submit_sc16")
        return t, n.value

    def process_device(self, in_dev: int, n_buffers: int, out_dev: int):
        lens = (C.c_int * n_buffers)()
        tot = check(self.lib.gsdr_rx_process_device(self._h, C.c_void_p(in_dev), n_buffers, C.c_void_p(out_dev), lens),
                    "gsdr_rx_process_device")
        return int(tot), list(lens)

    def sync(self):
        check(self.lib.gsdr_rx_sync(self._h), "gsdr_rx_sync")

    def reset(self):
        check(self.lib.gsdr_rx_reset(self._h), "gsdr_rx_reset")

    def timer_start(self):
        check(self.lib.gsdr_rx_timer_start(self._h), "gsdr_rx_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self.lib.gsdr_rx_timer_stop(self._h, C.byref(ms)), "gsdr_rx_timer_stop")
        return ms.value

    def max_output(self) -> int:
        return int(self.lib.gsdr_rx_max_output(self._h))

    def max_output_batch(self, n: int) -> int:
        return int(self.lib.gsdr_rx_max_output_batch(self._h, int(n)))

    def launch_count(self) -> int:
        return int(self.lib.gsdr_rx_launch_count(self._h))

    def kernel_name(self) -> str:
        return self.lib.gsdr_rx_kernel_name(self._h).decode()

    def taps(self) -> np.ndarray:
        n = self.lib.gsdr_rx_get_taps(self._h, None, 0)
        out = np.empty(max(n, 0), dtype=np.float32)
        if n > 0:
            self.lib.gsdr_rx_get_taps(self._h, out.ctypes.data_as(C.c_void_p), n)
        return out

    def bins(self) -> np.ndarray:
        n = self.lib.gsdr_rx_get_bins(self._h, None, 0)
        out = np.empty(max(n, 0), dtype=np.int32)
        if n > 0:
            self.lib.gsdr_rx_get_bins(self._h, out.ctypes.data_as(C.c_void_p), n)
        return out

    def chirp_param(self) -> _lib.ChirpParam:
        cp = _lib.ChirpParam()
        check(self.lib.gsdr_rx_chirp_param(self._h, C.byref(cp)), "gsdr_rx_chirp_param")
        return cp

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RxGroup:
    """One persistent launch over many TONES streams (gsdr_rx_group_*)."""

    def __init__(self, members):
        self.lib = _lib.load()
        self.members = list(members)
        arr = (C.c_void_p * len(self.members))(*[m._h for m in self.members])
        self._h = self.lib.gsdr_rx_group_create(arr, len(self.members))
        if not self._h:
            raise GsdrError(_lib.last_error())

    def process_device(self, in_devs, n_buffers, out_devs):
        n = len(self.members)
        ins = (C.c_void_p * n)(*in_devs)
        outs = (C.c_void_p * n)(*out_devs)
        lens = (C.c_int * (n * n_buffers))()
        tot = check(self.lib.gsdr_rx_group_process_device(self._h, ins, n_buffers, outs, lens), "gsdr_rx_group_process_device")
        return int(tot), np.array(lens, dtype=np.int64).reshape(n, n_buffers)

    # -- host-fed, one packet period per call (cfg5: many streams per GPU) -------------------------
    def pointer_arrays(self, ins, outs):
        """Pre-built ctypes pointer arrays for submit(): build once per ring position, reuse every period."""
        n = len(self.members)
        if len(ins) != n or len(outs) != n:
            raise ValueError("one input and one output buffer per member")
        for m, o in zip(self.members, outs):
            if o.size < m.max_output():
                raise ValueError("output buffer shorter than the member's max_output()")
        ia = (C.c_void_p * n)(*[a.ctypes.data for a in ins])
        oa = (C.c_void_p * n)(*[_ptr(a).value for a in outs])
        return ia, oa

    def submit(self, ins, outs=None, sc16=False):
        """ins/outs: lists of per-member buffers, or the pair returned by pointer_arrays().  -> (ticket, valid lengths)"""
        ia, oa = (ins, outs) if isinstance(ins, C.Array) else self.pointer_arrays(ins, outs)
        lens = (C.c_int * len(self.members))()
        fn = self.lib.gsdr_rx_group_submit_sc16 if sc16 else self.lib.gsdr_rx_group_submit
        t = check(fn(self._h, ia, oa, lens), "gsdr_rx_group_submit")
        return t, list(lens)

    def wait(self, ticket: int) -> None:
        check(self.lib.gsdr_rx_group_wait(self._h, int(ticket)), "gsdr_rx_group_wait")

    def process(self, ins, outs=None, sc16=False):
        t, lens = self.submit(ins, outs, sc16=sc16)
        self.wait(t)
        return lens

    def input_consumed(self, ticket: int) -> bool:
        return bool(check(self.lib.gsdr_rx_group_input_consumed(self._h, int(ticket)), "gsdr_rx_group_input_consumed"))

    def zero_copy(self) -> bool:
        return bool(self.lib.gsdr_rx_group_zero_copy(self._h))

    def set_form(self, mode: int) -> None:
        """0 copied both ways, 1 zero-copy both ways (default), 2 copy engine in / kernel stores out, 3 measured: a caller that
        keeps the pipeline full gets zero-copy and copied timed against each other once, the clearly faster one is kept."""
        check(self.lib.gsdr_rx_group_set_zero_copy(self._h, int(mode)), "gsdr_rx_group_set_zero_copy")

    def auto_choice(self, sc16: bool = False) -> int:
        """-1 while the measured form (mode 3) has not decided, else the form it kept (0 copied, 1 zero-copy)."""
        return int(self.lib.gsdr_rx_group_auto_choice(self._h, 1 if sc16 else 0))

    def last_form(self) -> int:
        return int(self.lib.gsdr_rx_group_last_form(self._h))

    def sync(self):
        check(self.lib.gsdr_rx_group_sync(self._h), "gsdr_rx_group_sync")

    def timer_start(self):
        check(self.lib.gsdr_rx_group_timer_start(self._h), "gsdr_rx_group_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self.lib.gsdr_rx_group_timer_stop(self._h, C.byref(ms)), "gsdr_rx_group_timer_stop")
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.gsdr_rx_group_launch_count(self._h))

    def close(self):
        if self._h:
            self.lib.gsdr_rx_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TX_buffer_generator:
    def __init__(self, init_parameters: param, device: int = 0):
        self.lib = _lib.load()
        self.parameters = init_parameters
        self.buffer_len = int(init_parameters.buffer_len)
        self._c, self._keep = init_parameters.to_c()
        self._h = self.lib.gsdr_tx_create(C.byref(self._c), int(device))
        if not self._h:
            raise GsdrError(_lib.last_error())
        self.device = int(device)

    def get(self, buffer: np.ndarray | None = None) -> np.ndarray:
        """TONES: returns a view into the generator-owned period buffer (``buffer`` is ignored,
        like the reference re-pointing ``*in``).  CHIRP: fills and returns ``buffer``."""
        p = C.c_void_p(buffer.ctypes.data if buffer is not None else None)
        if buffer is not None:
            _ptr(buffer)
        check(self.lib.gsdr_tx_get(self._h, C.byref(p)), "gsdr_tx_get")
        if buffer is not None and p.value == buffer.ctypes.data:
            return buffer
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(2 * self.buffer_len,)).view(np.complex64)

    def get_device(self, out_dev: int, n_buffers: int) -> None:
        check(self.lib.gsdr_tx_get_device(self._h, C.c_void_p(out_dev), int(n_buffers)), "gsdr_tx_get_device")

    def sync(self):
        check(self.lib.gsdr_tx_sync(self._h), "gsdr_tx_sync")

    def dynamic_buffer(self) -> bool:
        return bool(self.lib.gsdr_tx_dynamic_buffer(self._h))

    def launch_count(self) -> int:
        return int(self.lib.gsdr_tx_launch_count(self._h))

    def timer_start(self):
        check(self.lib.gsdr_tx_timer_start(self._h), "gsdr_tx_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self.lib.gsdr_tx_timer_stop(self._h, C.byref(ms)), "gsdr_tx_timer_stop")
        return ms.value

    def chirp_param(self) -> _lib.ChirpParam:
        cp = _lib.ChirpParam()
        check(self.lib.gsdr_tx_chirp_param(self._h, C.byref(cp)), "gsdr_tx_chirp_param")
        return cp

    def close(self) -> None:
        if self._h:
            self.lib.gsdr_tx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class preallocator:
    """Pinned buffer pool with the reference's get/trash/close contract."""

    def __init__(self, init_vector_size: int, init_pipe_size: int, prefill_init: bool = True):
        self.lib = _lib.load()
        self.vector_size = int(init_vector_size)
        self._h = self.lib.gsdr_pool_create(self.vector_size, int(init_pipe_size), int(bool(prefill_init)))
        if not self._h:
            raise GsdrError(_lib.last_error())

    def get(self) -> np.ndarray:
        p = self.lib.gsdr_pool_get(self._h)
        if not p:
            raise GsdrError("pool closed")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(2 * self.vector_size,)).view(np.complex64)

    def trash(self, buf: np.ndarray) -> None:
        self.lib.gsdr_pool_trash(self._h, C.c_void_p(buf.ctypes.data))

    def available(self) -> int:
        return int(self.lib.gsdr_pool_available(self._h))

    def size(self) -> int:
        return int(self.lib.gsdr_pool_size(self._h))

    def close(self) -> None:
        if self._h:
            self.lib.gsdr_pool_close(self._h)
            self._h = None


class ReplaySource:
    """Hardware-free IQ source (gsdr_replay_*): yields RX_wrapper-like packets from a pool."""

    TONES_NOISE, TX_LOOP = 0, 1

    def __init__(self, p: param, pool: preallocator, kind: int = 0, noise_sigma: float = 0.0, seed: int = 1337,
                 front_end_code: str = "B", rate_limit_msps: float = 0.0, device: int = 0):
        self.lib = _lib.load()
        self.pool = pool
        self.L = int(p.buffer_len)
        self._c, self._keep = p.to_c()
        self._h = self.lib.gsdr_replay_create(C.byref(self._c), int(kind), float(noise_sigma), int(seed), pool._h,
                                              front_end_code.encode()[:1], float(rate_limit_msps), int(device))
        if not self._h:
            raise GsdrError(_lib.last_error())

    def next(self):
        pkt = _lib.RxPacket()
        check(self.lib.gsdr_replay_next(self._h, C.byref(pkt)), "gsdr_replay_next")
        buf = np.ctypeslib.as_array(C.cast(pkt.buffer, C.POINTER(C.c_float)), shape=(2 * self.pool.vector_size,)).view(np.complex64)
        return pkt, buf

    def close(self):
        if self._h:
            self.lib.gsdr_replay_destroy(self._h)
            self._h = None


class DeviceBuffer:
    """Raw device allocation through the C-ABI (bench / device-resident tests)."""

    def __init__(self, n_complex: int, device: int = 0):
        self.lib = _lib.load()
        self.device = int(device)
        self.n = int(n_complex)
        self.ptr = self.lib.gsdr_dev_alloc(self.device, self.n * 8)
        if not self.ptr:
            raise GsdrError(_lib.last_error())

    def upload(self, a: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(a, dtype=np.complex64)
        check(self.lib.gsdr_memcpy_h2d(self.device, C.c_void_p(self.ptr + 8 * offset), a.ctypes.data_as(C.c_void_p), a.nbytes),
              "gsdr_memcpy_h2d")

    def download(self, n: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.n - offset if n is None else int(n)
        out = np.empty(n, dtype=np.complex64)
        if n:
            check(self.lib.gsdr_memcpy_d2h(self.device, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr + 8 * offset), out.nbytes),
                  "gsdr_memcpy_d2h")
        return out

    def free(self):
        if self.ptr:
            self.lib.gsdr_dev_free(self.device, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
