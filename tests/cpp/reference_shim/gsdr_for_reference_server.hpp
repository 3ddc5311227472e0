// How the reference's server is built against libgsdr: this header goes FIRST in every translation unit that uses the
// buffer-wrapper classes (g++ -include gsdr_for_reference_server.hpp ...).  It takes the reference's own
// headers/USRP_server_settings.hpp (param, w_type, RX_wrapper, the queue typedefs) and then claims the include guards of
// the four headers whose classes now live behind the C-ABI, so the reference's own copies become empty:
//
//   headers/USRP_server_memory_management.hpp  (USRP_MEMORY_INCLUDED)       preallocator<T>, buffer_helper,
//                                                                            VNA_decimator_helper, threading_condition
//   headers/USRP_demodulator.hpp               (USRP_DEMODULATOR_INCLUDED)   RX_buffer_demodulator
//   headers/USRP_buffer_generator.hpp          (USRP_BUFFER_GEN_INCLUDED)    TX_buffer_generator
//   headers/kernels.cuh                        (GPU_KERNELS_INCUDED_h)       nothing left to declare
//
// tests/test_reference_sources_compile.py compiles the reference's UNCHANGED cpp/USRP_server_link_threads.cpp this way and
// checks that the object file binds to gsdr_rx_process / gsdr_tx_get / gsdr_pool_* and to none of the reference's kernels.
#pragma once
#include "USRP_server_settings.hpp"
#include "USRP_server_diagnostic.hpp"
#define USRP_MEMORY_INCLUDED
#define USRP_DEMODULATOR_INCLUDED
#define USRP_BUFFER_GEN_INCLUDED
#define GPU_KERNELS_INCUDED_h
#define GSDR_COMPAT_REFERENCE_SETTINGS
#include "gsdr_compat.hpp"
