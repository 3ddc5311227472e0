"""ctypes binding of libgsdr.so (the C-ABI declared in include/gsdr.h).

The library is the product; this module only loads it and declares signatures.  There is no
fallback of any kind: if the shared library is missing the import raises, and every DSP entry
point fails with the library's own error string when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSDR_LIB_PATH", os.path.join(_HERE, "libgsdr.so"))  # override: kernel experiments only


class GsdrError(RuntimeError):
    pass


class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class CParam(C.Structure):
    """gsdr_param (include/gsdr.h) == POD flattening of the reference's ``param``."""

    _fields_ = [
        ("rate", C.c_int32), ("fft_tones", C.c_int32),
        ("decim", C.c_uint64), ("pf_average", C.c_uint64), ("buffer_len", C.c_uint64),
        ("data_mem_mult", C.c_uint64), ("samples", C.c_uint64),
        ("freq", C.POINTER(C.c_int32)), ("n_freq", C.c_uint64),
        ("ampl", C.POINTER(C.c_float)), ("n_ampl", C.c_uint64),
        ("wave_type", C.POINTER(C.c_int32)), ("n_wave_type", C.c_uint64),
        ("chirp_t", C.POINTER(C.c_float)), ("n_chirp_t", C.c_uint64),
        ("chirp_f", C.POINTER(C.c_int32)), ("n_chirp_f", C.c_uint64),
        ("swipe_s", C.POINTER(C.c_int32)), ("n_swipe_s", C.c_uint64),
    ]


class BufferHelper(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "n_tones", "eff_length", "buffer_len", "average", "n_eff_tones",
        "new_0", "copy_size", "current_batch", "spare_samples", "spare_begin")]


class VnaHelper(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("valid_size", "new0", "total_len", "spare_begin", "ppt", "buffer_len")]


class ChirpParam(C.Structure):
    _fields_ = [("num_steps", C.c_uint64), ("length", C.c_uint64), ("chirpness", C.c_uint32), ("f0", C.c_int32)]


class RxPacket(C.Structure):
    """gsdr_rx_packet == RX_wrapper (headers/USRP_server_settings.hpp:216-224)."""

    _fields_ = [("buffer", C.c_void_p), ("usrp_number", C.c_int32), ("front_end_code", C.c_char),
                ("packet_number", C.c_int32), ("length", C.c_int32), ("errors", C.c_int32), ("channels", C.c_int32)]


# name -> (restype, argtypes); every symbol include/gsdr.h declares
SIGNATURES = {
    "gsdr_last_error": (C.c_char_p, []),
    "gsdr_version": (C.c_char_p, []),
    "gsdr_device_count": (C.c_int, []),
    "gsdr_sm_count": (C.c_int, [C.c_int]),
    "gsdr_rx_create": (C.c_void_p, [C.POINTER(CParam), C.c_int, C.c_int]),
    "gsdr_rx_destroy": (None, [C.c_void_p]),
    "gsdr_rx_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsdr_rx_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "gsdr_rx_submit_sc16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "gsdr_rx_process_sc16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsdr_rx_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "gsdr_rx_input_consumed": (C.c_int, [C.c_void_p, C.c_int]),
    "gsdr_rx_pipeline_depth": (C.c_int, [C.c_void_p]),
    "gsdr_rx_process_device": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "gsdr_rx_sync": (C.c_int, [C.c_void_p]),
    "gsdr_rx_reset": (C.c_int, [C.c_void_p]),
    "gsdr_rx_channels": (C.c_int, [C.c_void_p]),
    "gsdr_rx_mode": (C.c_int, [C.c_void_p]),
    "gsdr_rx_max_output": (C.c_size_t, [C.c_void_p]),
    "gsdr_rx_max_output_batch": (C.c_size_t, [C.c_void_p, C.c_int]),
    "gsdr_rx_fcut": (C.c_float, [C.c_void_p]),
    "gsdr_rx_launch_count": (C.c_uint64, [C.c_void_p]),
    "gsdr_rx_timer_start": (C.c_int, [C.c_void_p]),
    "gsdr_rx_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "gsdr_rx_get_taps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "gsdr_rx_get_bins": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "gsdr_rx_chirp_param": (C.c_int, [C.c_void_p, C.POINTER(ChirpParam)]),
    "gsdr_rx_kernel_name": (C.c_char_p, [C.c_void_p]),
    "gsdr_rx_group_create": (C.c_void_p, [C.POINTER(C.c_void_p), C.c_int]),
    "gsdr_rx_group_destroy": (None, [C.c_void_p]),
    "gsdr_rx_group_process_device": (C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p),
                                                 C.POINTER(C.c_int)]),
    "gsdr_rx_group_submit": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "gsdr_rx_group_submit_sc16": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "gsdr_rx_group_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "gsdr_rx_group_input_consumed": (C.c_int, [C.c_void_p, C.c_int]),
    "gsdr_rx_group_process": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "gsdr_rx_group_pipeline_depth": (C.c_int, [C.c_void_p]),
    "gsdr_rx_group_members": (C.c_int, [C.c_void_p]),
    "gsdr_rx_group_zero_copy": (C.c_int, [C.c_void_p]),
    "gsdr_rx_group_set_zero_copy": (C.c_int, [C.c_void_p, C.c_int]),
    "gsdr_rx_group_auto_choice": (C.c_int, [C.c_void_p, C.c_int]),
    "gsdr_rx_group_last_form": (C.c_int, [C.c_void_p]),
    "gsdr_rx_group_sync": (C.c_int, [C.c_void_p]),
    "gsdr_rx_group_timer_start": (C.c_int, [C.c_void_p]),
    "gsdr_rx_group_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "gsdr_rx_group_launch_count": (C.c_uint64, [C.c_void_p]),
    "gsdr_tx_create": (C.c_void_p, [C.POINTER(CParam), C.c_int]),
    "gsdr_tx_destroy": (None, [C.c_void_p]),
    "gsdr_tx_get": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gsdr_tx_get_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gsdr_tx_sync": (C.c_int, [C.c_void_p]),
    "gsdr_tx_dynamic_buffer": (C.c_int, [C.c_void_p]),
    "gsdr_tx_buffer_len": (C.c_int, [C.c_void_p]),
    "gsdr_tx_launch_count": (C.c_uint64, [C.c_void_p]),
    "gsdr_tx_timer_start": (C.c_int, [C.c_void_p]),
    "gsdr_tx_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "gsdr_tx_chirp_param": (C.c_int, [C.c_void_p, C.POINTER(ChirpParam)]),
    "gsdr_make_sinc_window": (C.c_int, [C.c_int, C.c_float, C.c_void_p]),
    "gsdr_make_flat_window": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "gsdr_pfb_batching": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "gsdr_tone_bins": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gsdr_pfb_gather_layout": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gsdr_pfb_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gsdr_group_form_simulate": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "gsdr_buffer_helper_init": (None, [C.POINTER(BufferHelper), C.c_int, C.c_int, C.c_int, C.c_int]),
    "gsdr_buffer_helper_update": (None, [C.POINTER(BufferHelper)]),
    "gsdr_vna_helper_init": (None, [C.POINTER(VnaHelper), C.c_int, C.c_int]),
    "gsdr_vna_helper_update": (None, [C.POINTER(VnaHelper)]),
    "gsdr_chirp_params": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.POINTER(ChirpParam)]),
    "gsdr_probe_chirp_index": (C.c_int, [C.c_int, C.POINTER(ChirpParam), C.c_uint64, C.c_uint32, C.c_void_p]),
    "gsdr_probe_direct_phase": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "gsdr_probe_direct_tile_phase": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gsdr_spec_n_freq": (C.c_longlong, [C.c_size_t, C.c_int, C.c_size_t]),
    "gsdr_spec_from_samples": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "gsdr_pool_create": (C.c_void_p, [C.c_size_t, C.c_int, C.c_int]),
    "gsdr_pool_get": (C.c_void_p, [C.c_void_p]),
    "gsdr_pool_trash": (None, [C.c_void_p, C.c_void_p]),
    "gsdr_pool_close": (None, [C.c_void_p]),
    "gsdr_pool_available": (C.c_int, [C.c_void_p]),
    "gsdr_pool_size": (C.c_int, [C.c_void_p]),
    "gsdr_pcie_copy_ceiling": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "gsdr_pcie_copy_ceiling_streams": (C.c_int, [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gsdr_pcie_probe_create": (C.c_void_p, [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int]),
    "gsdr_pcie_probe_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gsdr_pcie_probe_destroy": (None, [C.c_void_p]),
    "gsdr_host_alloc": (C.c_void_p, [C.c_size_t]),
    "gsdr_device_numa_node": (C.c_int, [C.c_int]),
    "gsdr_host_free": (None, [C.c_void_p]),
    "gsdr_dev_alloc": (C.c_void_p, [C.c_int, C.c_size_t]),
    "gsdr_dev_free": (None, [C.c_int, C.c_void_p]),
    "gsdr_memcpy_h2d": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "gsdr_memcpy_d2h": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "gsdr_dev_memset": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_size_t]),
    "gsdr_device_synchronize": (C.c_int, [C.c_int]),
    "gsdr_replay_create": (C.c_void_p, [C.POINTER(CParam), C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_char,
                                        C.c_double, C.c_int]),
    "gsdr_replay_next": (C.c_int, [C.c_void_p, C.POINTER(RxPacket)]),
    "gsdr_packet_header_write": (C.c_int, [C.POINTER(RxPacket), C.c_void_p]),
    "gsdr_packet_header_read": (C.c_int, [C.c_void_p, C.POINTER(RxPacket)]),
    "gsdr_packet_frame": (C.c_int, [C.POINTER(RxPacket), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "gsdr_replay_destroy": (None, [C.c_void_p]),
    "gsdr_replay_packets": (C.c_uint64, [C.c_void_p]),
}

_lib = None


def load():
    """Load libgsdr.so and attach signatures.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GsdrError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                "gpu_sdr_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().gsdr_last_error().decode(errors="replace")


def check(rc, what):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise GsdrError(f"{what}: {last_error()}")
    return rc
