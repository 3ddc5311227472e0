// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI harness around the reference's OWN, UNMODIFIED
// host classes and kernels, compiled from where they lie under /root/reference into
// oracle/_ref/libgsdr_ref.so (see oracle/Makefile, target `ref`).  It lets tests/ and
// bench.py --impl reference drive RX_buffer_demodulator / TX_buffer_generator exactly as
// TXRX::rx_single_link / tx_single_link do (cpp/USRP_server_link_threads.cpp:584,666), with the
// same POD parameter block as the product's C-ABI (include/gsdr.h).
//
// No DSP is implemented here.  The only definitions are (a) the glue that fills a reference
// `param` from a gsdr_param, (b) trivial bodies for three non-DSP symbols whose own translation
// units need UHD/Boost proper (thread priority/naming, param::dynamic_buffer), and (c) read-out
// of private members (taps, bins, chirp parameters) for the parity tests.
#define private public  // test-only peek at RX_buffer_demodulator / TX_buffer_generator internals
#include "USRP_demodulator.hpp"
#include "USRP_buffer_generator.hpp"
#undef private

#include <chrono>
#include <new>
#include <thread>
#include <vector>
#include "../include/gsdr.h"

// ---- non-DSP symbols normally provided by USRP_server_settings.cpp / USRP_server_diagnostic.cpp
void Thread_Prioriry(boost::thread&, int, int) {}
void SetThreadName(boost::thread*, const char*) {}
void set_this_thread_name(std::string) {}
void print_chirp_params(std::string, chirp_parameter) {}
bool param::dynamic_buffer() {  // cpp/USRP_server_settings.cpp:98-102
    bool dynamic = false;
    for (size_t i = 0; i < wave_type.size(); i++) if (wave_type[i] != TONES) dynamic = true;
    return dynamic;
}
int param::get_output_buffer_size() { return std::ceil((float)buffer_len / (float)decim) * wave_type.size(); }

namespace {
param* make_param(const gsdr_param* g) {
    param* p = new param();
    p->mode = RX;
    p->rate = g->rate;
    p->gain = 0;
    p->bw = 0;
    p->tone = 0;
    p->samples = g->samples;
    p->delay = 0;
    p->burst_on = 0;
    p->burst_off = 0;
    p->buffer_len = g->buffer_len;
    p->tuning_mode = 0;
    p->decim = g->decim;
    p->data_mem_mult = g->data_mem_mult;
    p->fft_tones = g->fft_tones;
    p->pf_average = g->pf_average;
    p->freq.assign(g->freq, g->freq + g->n_freq);
    p->ampl.assign(g->ampl, g->ampl + g->n_ampl);
    for (uint64_t i = 0; i < g->n_wave_type; i++) p->wave_type.push_back((w_type)g->wave_type[i]);
    p->chirp_t.assign(g->chirp_t, g->chirp_t + g->n_chirp_t);
    p->chirp_f.assign(g->chirp_f, g->chirp_f + g->n_chirp_f);
    p->swipe_s.assign(g->swipe_s, g->swipe_s + g->n_swipe_s);
    return p;
}
struct ref_rx { param* p; RX_buffer_demodulator* d; };
struct ref_tx { param* p; TX_buffer_generator* g; };
}  // namespace

extern "C" {

int gsdr_ref_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void* gsdr_ref_rx_create(const gsdr_param* g) {
    ref_rx* r = new ref_rx();
    r->p = make_param(g);
    // The class leaves several members uninitialised (e.g. `handle`, which close_pfb/close_chirp
    // pass to cublasDestroy even when no handle was created).  Constructing into zeroed storage
    // gives them the value a fresh heap page has, so close() is deterministic in a long test run.
    void* mem = calloc(1, sizeof(RX_buffer_demodulator));
    r->d = new (mem) RX_buffer_demodulator(r->p, false);  // cpp/USRP_server_link_threads.cpp:121
    cudaDeviceSynchronize();
    return r;
}
int gsdr_ref_rx_process(void* h, gsdr_float2* in, gsdr_float2* out) {
    ref_rx* r = (ref_rx*)h;
    float2* i = (float2*)in;
    float2* o = (float2*)out;
    return r->d->process(&i, &o);  // cpp/USRP_server_link_threads.cpp:666
}
// K back-to-back process() calls on the same host buffers, wall-clock timed (process() is
// synchronous: it ends with cudaStreamSynchronize).  Returns seconds.
double gsdr_ref_rx_process_timed(void* h, gsdr_float2* in, gsdr_float2* out, int k, int* last_len) {
    ref_rx* r = (ref_rx*)h;
    float2* i = (float2*)in;
    float2* o = (float2*)out;
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    int len = 0;
    for (int n = 0; n < k; n++) len = r->d->process(&i, &o);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    if (last_len) *last_len = len;
    return std::chrono::duration<double>(t1 - t0).count();
}
// The reference's threading model for several front-ends: one worker thread per RX_buffer_demodulator, each calling the
// blocking process() on its own packets (TXRX::rx_single_link, cpp/USRP_server_link_threads.cpp:605-702).  n instances,
// thread i makes k calls cycling through its `ring` input buffers ins[i*ring .. i*ring+ring).  Wall-clock seconds.
double gsdr_ref_rx_multi_process_timed(void** hs, int n, gsdr_float2** ins, int ring, gsdr_float2** outs, int k) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceSynchronize();
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++)
        th.emplace_back([=] {
            cudaSetDevice(dev);
            ref_rx* r = (ref_rx*)hs[i];
            for (int c = 0; c < k; c++) {
                float2* in = (float2*)ins[(size_t)i * ring + (c % ring)];
                float2* out = (float2*)outs[i];
                r->d->process(&in, &out);
            }
        });
    for (auto& t : th) t.join();
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
// Device time of k process() calls split into "everything the call enqueues on its stream" and "its two copies": CUDA events
// on the instance's own stream around the unmodified process() (which ends with cudaStreamSynchronize, so the stream is idle
// at both events), then the same two cudaMemcpyAsync (in_bytes up, out_bytes down, same pinned host buffers) alone on that
// stream.  kernels = total - copies: the reference's launch sequence (cpp/USRP_demodulator.cpp:486-565 for TONES) without
// H2D / D2H, as SURVEY.md section 8(d) asks.
int gsdr_ref_rx_process_split_timed(void* h, gsdr_float2** ins, int ring, gsdr_float2* out, int k, size_t in_bytes, size_t out_bytes,
                                    double* total_ms, double* copy_ms) {
    ref_rx* r = (ref_rx*)h;
    cudaStream_t st = r->d->internal_stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    void *d_a = nullptr, *d_b = nullptr;
    if (cudaMalloc(&d_a, in_bytes ? in_bytes : 1) != cudaSuccess || cudaMalloc(&d_b, out_bytes ? out_bytes : 1) != cudaSuccess) return -1;
    double tot = 0.0, cop = 0.0;
    float ms = 0.f;
    for (int c = 0; c < k; c++) {
        float2* in = (float2*)ins[c % ring];
        float2* o = (float2*)out;
        cudaEventRecord(e0, st);
        r->d->process(&in, &o);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        tot += ms;
        cudaEventRecord(e0, st);
        cudaMemcpyAsync(d_a, in, in_bytes, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(o, d_b, out_bytes, cudaMemcpyDeviceToHost, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        cop += ms;
    }
    cudaFree(d_a);
    cudaFree(d_b);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (total_ms) *total_ms = tot;
    if (copy_ms) *copy_ms = cop;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
void gsdr_ref_rx_close(void* h) {
    ref_rx* r = (ref_rx*)h;
    r->d->close();
    r->d->~RX_buffer_demodulator();
    free(r->d);
    delete r->p;
    delete r;
}
int gsdr_ref_rx_channels(void* h) { return (int)((ref_rx*)h)->d->parameters->wave_type.size(); }
// private read-outs for parity tests
int gsdr_ref_rx_bins(void* h, int32_t* bins, int cap) {
    ref_rx* r = (ref_rx*)h;
    int n = r->d->h_param.eff_n_tones;
    if (n > cap) n = cap;
    cudaMemcpy(bins, r->d->h_param.tones, n * sizeof(int), cudaMemcpyDeviceToHost);
    return n;
}
int gsdr_ref_rx_batching(void* h) { return ((ref_rx*)h)->d->batching; }
int gsdr_ref_rx_window(void* h, float* taps, int cap) {  // PFB window (device float2 -> real part)
    ref_rx* r = (ref_rx*)h;
    int n = r->d->h_param.n_tones * r->d->h_param.average_buffer;
    if (n > cap) n = cap;
    std::vector<float2> tmp(n);
    cudaMemcpy(tmp.data(), r->d->window, n * sizeof(float2), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; i++) taps[i] = tmp[i].x;
    return n;
}
int gsdr_ref_rx_chirp_param(void* h, gsdr_chirp_param* out) {
    ref_rx* r = (ref_rx*)h;
    out->num_steps = r->d->h_parameter.num_steps;
    out->length = r->d->h_parameter.length;
    out->chirpness = r->d->h_parameter.chirpness;
    out->f0 = r->d->h_parameter.f0;
    return 0;
}

void* gsdr_ref_tx_create(const gsdr_param* g) {
    ref_tx* t = new ref_tx();
    t->p = make_param(g);
    t->p->mode = TX;
    void* mem = calloc(1, sizeof(TX_buffer_generator));
    t->g = new (mem) TX_buffer_generator(t->p);  // cpp/USRP_server_link_threads.cpp:191
    cudaDeviceSynchronize();
    return t;
}
// Copies the next buffer_len samples into `out` (TONES re-points; CHIRP fills the buffer).
int gsdr_ref_tx_get(void* h, gsdr_float2* out) {
    ref_tx* t = (ref_tx*)h;
    float2* ptr = (float2*)out;
    t->g->get(&ptr);  // cpp/USRP_server_link_threads.cpp:584
    if (ptr != (float2*)out) memcpy(out, ptr, sizeof(float2) * t->g->buffer_len);
    return t->g->buffer_len;
}
int gsdr_ref_tx_chirp_param(void* h, gsdr_chirp_param* out) {
    ref_tx* t = (ref_tx*)h;
    out->num_steps = t->g->h_parameter.num_steps;
    out->length = t->g->h_parameter.length;
    out->chirpness = t->g->h_parameter.chirpness;
    out->f0 = t->g->h_parameter.f0;
    return 0;
}
void gsdr_ref_tx_close(void* h) {
    ref_tx* t = (ref_tx*)h;
    t->g->close();
    t->g->~TX_buffer_generator();
    free(t->g);
    delete t->p;
    delete t;
}

// Host-only: no CUDA call is made with host_ret=true (cpp/kernels.cu:297-308).
int gsdr_ref_make_sinc_window(int length, float fc, float* taps) {
    float2* w = make_sinc_window(length, fc, false, true);
    for (int i = 0; i < length; i++) taps[i] = w[i].x;
    free(w);
    return length;
}
// Needs a GPU (the reference returns a device pointer).
int gsdr_ref_make_flat_window(int length, int side, float* taps) {
    float2* d = make_flat_window(length, side, false);
    std::vector<float2> tmp(length);
    if (cudaMemcpy(tmp.data(), d, length * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int i = 0; i < length; i++) taps[i] = tmp[i].x;
    cudaFree(d);
    return length;
}
// buffer_helper state after construction and after each of n-1 updates:
// rows of (eff_length,new_0,copy_size,current_batch,spare_samples,spare_begin).  Host-only.
void gsdr_ref_buffer_helper_seq(int n_tones, int buffer_len, int average, int n_eff, int n, int* out) {
    buffer_helper h(n_tones, buffer_len, average, n_eff);
    for (int i = 0; i < n; i++) {
        int* o = out + 6 * i;
        o[0] = h.eff_length; o[1] = h.new_0; o[2] = h.copy_size;
        o[3] = h.current_batch; o[4] = h.spare_samples; o[5] = h.spare_begin;
        h.update();
    }
}
// VNA_decimator_helper rows of (valid_size,new0,total_len,spare_begin).  Host-only.
void gsdr_ref_vna_helper_seq(int ppt, int buffer_len, int n, int* out) {
    VNA_decimator_helper h(ppt, buffer_len);
    for (int i = 0; i < n; i++) {
        int* o = out + 4 * i;
        o[0] = h.valid_size; o[1] = h.new0; o[2] = h.total_len; o[3] = h.spare_begin;
        h.update();
    }
}
void* gsdr_ref_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void gsdr_ref_host_free(void* p) { cudaFreeHost(p); }

}  // extern "C"
