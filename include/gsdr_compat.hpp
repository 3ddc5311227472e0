// gsdr_compat.hpp -- header-only C++ shims with the reference's exact class names and signatures
// over the C-ABI (gsdr.h), so that the reference's callers (cpp/USRP_server_link_threads.cpp:
// 109-226 construct, :584 get, :666 process, :475-521 close) compile against libgsdr.so unchanged.
//
//   reference declaration                                        shim below
//   enum w_type / ant_mode   headers/USRP_server_settings.hpp:113,122    same enumerators, same order
//   struct param             headers/USRP_server_settings.hpp:130-167    same fields + methods
//   struct RX_wrapper        headers/USRP_server_settings.hpp:216-224    same fields
//   preallocator<T>          headers/USRP_server_memory_management.hpp:103-273   get/trash/close
//   buffer_helper, VNA_decimator_helper   headers/USRP_server_memory_management.hpp:22-101
//   RX_buffer_demodulator    headers/USRP_demodulator.hpp:13-33
//   TX_buffer_generator      headers/USRP_buffer_generator.hpp:49-68
//
// Error behaviour follows the reference: an unsupported configuration prints the error and calls
// exit(-1) from the constructor; process()/get() cannot fail (a CUDA error aborts the same way).
// No CUDA headers are needed to use this file.
#pragma once
#include <cstddef>
#include <cstdio>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "gsdr.h"

// GSDR_COMPAT_REFERENCE_SETTINGS: the including translation unit has the reference's own headers/USRP_server_settings.hpp
// (float2 from CUDA, w_type, ant_mode, param, RX_wrapper, the queue typedefs) -- the way the reference's server is built
// against this library (tests/cpp/reference_shim/).  Otherwise the same types are defined here, field for field.
#ifndef GSDR_COMPAT_REFERENCE_SETTINGS
#if !defined(__VECTOR_TYPES_H__) && !defined(GSDR_COMPAT_HAVE_FLOAT2)
#define GSDR_COMPAT_HAVE_FLOAT2
struct float2 { float x, y; };
#endif

enum w_type { TONES, CHIRP, NOISE, RAMP, NODSP, SWONLY, DIRECT };
enum ant_mode { TX, RX, OFF };

struct param {
    ant_mode mode = OFF;
    int rate = 0, gain = 0, bw = 0;
    size_t tone = 0;
    size_t samples = 0;
    double delay = 0;
    float burst_on = 0, burst_off = 0;
    size_t buffer_len = 0;
    bool tuning_mode = false;
    std::vector<int> freq;
    std::vector<w_type> wave_type;
    std::vector<float> ampl;
    size_t decim = 0;
    std::vector<float> chirp_t;
    std::vector<int> chirp_f;
    std::vector<int> swipe_s;
    size_t data_mem_mult = 1;
    int fft_tones = 0;
    size_t pf_average = 1;

    int get_output_buffer_size() { return (int)((buffer_len + (decim ? decim : 1) - 1) / (decim ? decim : 1)) * (int)wave_type.size(); }
    bool dynamic_buffer() {
        for (size_t i = 0; i < wave_type.size(); i++)
            if (wave_type[i] != TONES) return true;
        return false;
    }
};

struct RX_wrapper {
    float2* buffer;
    int usrp_number;
    char front_end_code;
    int packet_number;
    int length;
    int errors;
    int channels;
};
#endif  // GSDR_COMPAT_REFERENCE_SETTINGS
static_assert(sizeof(float2) == sizeof(gsdr_float2), "float2 layout");

namespace gsdr_compat {
struct flat_param {  // keeps the int32 copy of wave_type alive next to the POD block
    gsdr_param c{};
    std::vector<int32_t> wt;
    explicit flat_param(const param* p) {
        wt.assign(p->wave_type.begin(), p->wave_type.end());
        c.rate = p->rate;
        c.fft_tones = p->fft_tones;
        c.decim = p->decim;
        c.pf_average = p->pf_average;
        c.buffer_len = p->buffer_len;
        c.data_mem_mult = p->data_mem_mult;
        c.samples = p->samples;
        c.freq = p->freq.data();        c.n_freq = p->freq.size();
        c.ampl = p->ampl.data();        c.n_ampl = p->ampl.size();
        c.wave_type = wt.data();        c.n_wave_type = wt.size();
        c.chirp_t = p->chirp_t.data();  c.n_chirp_t = p->chirp_t.size();
        c.chirp_f = p->chirp_f.data();  c.n_chirp_f = p->chirp_f.size();
        c.swipe_s = p->swipe_s.data();  c.n_swipe_s = p->swipe_s.size();
    }
};
[[noreturn]] inline void die(const char* what) {
    std::fprintf(stderr, "\n\033[1;31mERROR\033[0m: %s\n", what);  // print_error's format, then the reference's exit(-1)
    std::exit(-1);
}
}  // namespace gsdr_compat

// headers/USRP_server_memory_management.hpp:9-19 (the link threads wait on it before joining)
class threading_condition {
  public:
    threading_condition() : ready(false) {}
    void wait() {
        std::unique_lock<std::mutex> lock(ready_mutex);
        while (!ready) ready_cond.wait(lock);
    }
    void release() {
        {
            std::unique_lock<std::mutex> lock(ready_mutex);
            ready = true;
        }
        ready_cond.notify_all();
    }
    void rearm() {
        {
            std::unique_lock<std::mutex> lock(ready_mutex);
            ready = false;
        }
        ready_cond.notify_all();
    }

  private:
    std::condition_variable ready_cond;
    std::mutex ready_mutex;
    bool ready;
};

template <typename vector_type>
class preallocator {
  public:
    int vector_size, pipe_size, wait_on_full = 5;
    bool prefil;
    preallocator(int init_vector_size, int init_pipe_size, bool prefill_init = true, int /*core*/ = -1)
        : vector_size(init_vector_size), pipe_size(init_pipe_size), prefil(prefill_init) {
        static_assert(sizeof(vector_type) == sizeof(gsdr_float2), "the pool holds float2 buffers");
        pool_ = gsdr_pool_create((size_t)init_vector_size, init_pipe_size, prefill_init ? 1 : 0);
        if (!pool_) gsdr_compat::die("Memory manager cannot allocate pinned host memory!");
    }
    vector_type* get() { return reinterpret_cast<vector_type*>(gsdr_pool_get(pool_)); }
    void trash(vector_type* v) { gsdr_pool_trash(pool_, reinterpret_cast<gsdr_float2*>(v)); }
    void close() {
        gsdr_pool_close(pool_);
        pool_ = nullptr;
    }
    gsdr_pool* handle() { return pool_; }

  private:
    gsdr_pool* pool_ = nullptr;
};

class buffer_helper {
  public:
    int n_tones, eff_length, buffer_len, average, n_eff_tones, new_0, copy_size, current_batch, spare_samples, spare_begin;
    buffer_helper(int _n_tones, int _buffer_len, int _average, int _n_eff_tones) {
        gsdr_buffer_helper_init(&h_, _n_tones, _buffer_len, _average, _n_eff_tones);
        sync();
    }
    void update() {
        gsdr_buffer_helper_update(&h_);
        sync();
    }

  private:
    gsdr_buffer_helper h_;
    void sync() {
        n_tones = h_.n_tones; eff_length = h_.eff_length; buffer_len = h_.buffer_len; average = h_.average;
        n_eff_tones = h_.n_eff_tones; new_0 = h_.new_0; copy_size = h_.copy_size; current_batch = h_.current_batch;
        spare_samples = h_.spare_samples; spare_begin = h_.spare_begin;
    }
};

class VNA_decimator_helper {
  public:
    int valid_size, new0, total_len, spare_begin;
    VNA_decimator_helper(int init_ppt, int init_buffer_len) {
        gsdr_vna_helper_init(&h_, init_ppt, init_buffer_len);
        sync();
    }
    void update() {
        gsdr_vna_helper_update(&h_);
        sync();
    }

  private:
    gsdr_vna_helper h_;
    void sync() { valid_size = h_.valid_size; new0 = h_.new0; total_len = h_.total_len; spare_begin = h_.spare_begin; }
};

class RX_buffer_demodulator {
  public:
    param* parameters;
    float fcut;
    RX_buffer_demodulator(param* init_parameters, bool init_diagnostic = false, int device = 0) : parameters(init_parameters) {
        gsdr_compat::flat_param fp(init_parameters);
        h_ = gsdr_rx_create(&fp.c, device, init_diagnostic ? 1 : 0);
        if (!h_) gsdr_compat::die(gsdr_last_error());
        fcut = gsdr_rx_fcut(h_);
    }
    // `*in`: buffer_len float2 of pinned host memory (not modified); `*out`: pinned host buffer of
    // buffer_len*max(data_mem_mult,1) float2.  Returns the valid float2 count (sample-major).
    int process(float2** __restrict__ in, float2** __restrict__ out) {
        const int n = gsdr_rx_process(h_, reinterpret_cast<const gsdr_float2*>(*in), reinterpret_cast<gsdr_float2*>(*out));
        if (n < 0) gsdr_compat::die(gsdr_last_error());
        return n;
    }
    void close() {
        gsdr_rx_destroy(h_);
        h_ = nullptr;
    }
    gsdr_rx* handle() { return h_; }

  private:
    gsdr_rx* h_ = nullptr;
};

class TX_buffer_generator {
  public:
    int buffer_len;
    param* parameters;
    explicit TX_buffer_generator(param* init_parameters, int device = 0) : parameters(init_parameters) {
        buffer_len = (int)init_parameters->buffer_len;
        gsdr_compat::flat_param fp(init_parameters);
        h_ = gsdr_tx_create(&fp.c, device);
        if (!h_) gsdr_compat::die(gsdr_last_error());
    }
    // TONES: re-points *in into generator-owned pinned memory; CHIRP: fills the caller's *in.
    void get(float2** __restrict__ in) {
        gsdr_float2* p = reinterpret_cast<gsdr_float2*>(*in);
        if (gsdr_tx_get(h_, &p)) gsdr_compat::die(gsdr_last_error());
        *in = reinterpret_cast<float2*>(p);
    }
    void close() {
        gsdr_tx_destroy(h_);
        h_ = nullptr;
    }
    gsdr_tx* handle() { return h_; }

  private:
    gsdr_tx* h_ = nullptr;
};
