// Shared declarations for libgsdr (B200 / sm_100a readout DSP path).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gsdr.h"

namespace gsdr {

// thread-local error string behind gsdr_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

// "Done once per device": cudaFuncSetAttribute applies to the current device only, so a process that runs demodulators on
// several GPUs has to raise a kernel's dynamic shared-memory limit on each of them; worker threads of different instances
// may get here concurrently (a repeated set is harmless, a missing one fails the launch).
class DeviceOnce {
    std::atomic<unsigned long long> mask_{0};

  public:
    // device to prepare, or -1 when the current device already is (devices >= 64 are prepared on every call)
    int pending() const {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return 0;
        if (d < 0 || d >= 64) return d < 0 ? 0 : d;
        return ((mask_.load(std::memory_order_acquire) >> d) & 1ull) ? -1 : d;
    }
    void done(int d) {
        if (d >= 0 && d < 64) mask_.fetch_or(1ull << d, std::memory_order_release);
    }
};

#define GSDR_CUDA_OK(expr)                                                                 \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::gsdr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                              __FILE__, __LINE__);                                         \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

#define GSDR_CUDA_OK_NULL(expr)                                                            \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::gsdr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                              __FILE__, __LINE__);                                         \
            return nullptr;                                                                \
        }                                                                                  \
    } while (0)

// ---- host runtime helpers (host.cu) -----------------------------------------------------------
// Device-side alias of [host, host+bytes) when it is pinned and mapped (a kernel may read / write it in place), else nullptr.
// Library allocations are answered from a registry; foreign pointers are asked of the driver (cached only on request).
void* host_alias_of(const void* host, size_t bytes);
void host_registry_add(const void* p, size_t bytes);
void host_registry_remove(const void* p);
// cudaSetDevice only when the calling thread's current device differs (cudaGetDevice is a thread-local read)
int ensure_device(int device);

// ---- host logic that defines results (hostlogic.cpp) ----------------------------------------
void make_sinc_window(int length, float fc, float* out);
void make_flat_window(int length, int side, float* out);
int pfb_batching(int buffer_len, int fft_tones, int pf_average);
void tone_bins(int rate, int fft_tones, const int32_t* freq, int n, int32_t* bins);
void pfb_gather_coloring(const int32_t* bins /* nullptr: all bins in order */, int T, uint8_t* perm /* [2048] */);
void buffer_helper_init(gsdr_buffer_helper* h, int n_tones, int buffer_len, int average, int n_eff_tones);
void buffer_helper_update(gsdr_buffer_helper* h);
void vna_helper_init(gsdr_vna_helper* h, int ppt, int buffer_len);
void vna_helper_update(gsdr_vna_helper* h);
void chirp_params(int rate, int freq0, int chirp_f0, int swipe_s0, float chirp_t0, bool tx, gsdr_chirp_param* out);

// A view of a stream window that is the concatenation of the carried-over samples of the
// previous call (`hist`, n_hist float2) and the new samples (`in`, n_in float2).  This is what
// replaces the reference's move_buffer + upload-at-new_0 (cpp/USRP_demodulator.cpp:491-509).
struct Window {
    const float2* hist;
    const float2* in;
    long long n_hist;
    long long n_in;
};

// ---- PFB (TONES / NOISE) launches (pfb_kernels.cu) -------------------------------------------
struct PfbJob {          // one stream's share of a launch
    Window win;
    const float* taps;   // [P*N] real taps (device)
    const int* bins;     // [T] selected FFT bins (device); nullptr => all bins in order (T == N)
    float2* out;         // [n_frames*T] sample-major
    int first_frame;     // frame index within the window (0 = window start)
    int n_frames;
    int N, P, T;
    const unsigned char* xperm = nullptr;  // [2048] in-row position of each bin (pfb_gather_coloring); nullptr => k2
    // Carry-over of a single-stream launch: the last tail_n samples of the window go to tail_dst (the other half of the
    // history ping-pong).  The warp-specialised fused kernel does this copy itself (its last CTA has the short remainder
    // tile), which saves the separate window_tail_kernel launch; every other path copies after its kernels.
    float2* tail_dst = nullptr;
    long long tail_n = 0;
    // Launch hint (host side only): lower bound on the frames per tile.  Set when the window lives in pinned HOST memory
    // (zero-copy blocking call): reads over PCIe are not kept in L2, so the P-1 halo rows at every tile start are paid in
    // full and few long tiles beat one short tile per SM.
    int min_tile = 0;
    // sc16 ingest fused into the warp-specialised kernel: win.in then points at interleaved int16 I/Q pairs (n_in samples of
    // 4 bytes) and the producers convert on the fly, (float)v * (1/32767); the history is fc32 as always.
    int in_sc16 = 0;
};
// One tile of a multi-stream launch of the warp-specialised fused kernel: frames [fa, fb) of job `job`, relative to the job's
// first_frame.  The host cuts the concatenated frame sequence of all jobs into equal shares, one per CTA, and splits a share
// where it crosses a stream boundary, so every CTA carries the same number of frames whatever the number of streams
// (uniform per-job tiles leave up to a quarter of the SMs idle at 64 streams).  flags bit 0: this is the job's last tile --
// its CTA also copies the job's carry-over tail (PfbJob::tail_dst).
struct PfbTile {
    int job, fa, fb, flags;
};
// The host side of that cut (hostlogic.cpp; gsdr_pfb_partition exposes it to the CPU tests): equal COST per CTA, where
// every tile a CTA starts costs kPfbTileCost frames' worth of pipeline fill, drain and per-stream constants.
constexpr long long kPfbTileCost = 24;
void pfb_partition(const int* n_frames, int n_jobs, int grid, std::vector<PfbTile>& tiles, std::vector<int>& cta_begin);

// Form of the host-fed group call (rx.cu, gsdr_rx_group_submit) decided by measurement -- mode 3, opt-in -- (hostlogic.cpp;
// gsdr_group_form_simulate exposes it to the CPU tests).  Which form moves a packet period faster depends on the platform (who
// else pulls on the host's memory, PCIe topology): measured on B200 boxes with the pipeline kept full, the zero-copy form is
// level at 1 GPU, 16 % ahead at 2, 6 % behind at 4 and 2.5 % behind at 8.  It is always the lower-latency form (no separate copy
// phases), so a caller that feeds packets at their own pace -- the previous period done before the next is submitted --
// simply gets it.  Only when the caller keeps the pipeline full (kBusy submits in a row that found the previous period still in
// flight: throughput is what counts) the form is MEASURED: kBlock periods zero-copy, kBlock copied, the time from submit to
// the return of the wait averaged over each block (its first kSkip periods, which queue behind the other form, left out),
// the copied form kept from then on if it is at least 3 % faster, else zero-copy.  A submit that finds the pipeline drained
// abandons the measurement.  Both forms give
// bit-identical results, so the switches are invisible in the data.
struct GroupAutoForm {
    static constexpr int kBlock = 8, kSkip = 3, kBusy = 4;
    static constexpr double kMargin = 0.97;   // the copied form must take less than this fraction of the zero-copy time
    int choice = -1;            // -1 undecided, else 0 (copied) or 1 (zero-copy)
    bool measuring = false;
    int busy_run = 0;           // consecutive submits that found the previous period in flight
    int cal_submits = 0;        // submits since the measurement began
    double sum[2] = {0.0, 0.0};
    int cnt[2] = {0, 0};
    // One submit: `busy` = the previous period is still in flight.  Returns the form for this period (0 copied, 1 zero-copy);
    // *block = 0 / 1 when the period is a measured one (report its submit-to-wait time with on_wait), else -1.
    int on_submit(bool busy, int* block);
    void on_wait(int block, double seconds);
};
// Device scratch a multi-stream launch needs for its job table, tile list and per-CTA tile ranges.
size_t pfb_table_bytes(int n_jobs, int sm_count);
// Returns the number of kernel launches issued (>0) or -1.  `workspace` is device scratch of at
// least pfb_workspace_bytes() for the generic path (may be null for the fused path).
bool pfb_fused_supported(int N, int P, int T, const Window& w);
bool pfb_fused_sc16_available();   // the default (warp-specialised) kernel reads sc16 windows; the lock-step cross-check kernel does not
const char* pfb_kernel_name(int N, int P, int T);
size_t pfb_workspace_bytes(int N, int P, int max_frames);
int pfb_launch(const PfbJob* jobs_host, int n_jobs, void* jobs_dev_scratch, void* workspace, const float2* twiddle_dev,
               int sm_count, cudaStream_t stream);
// NOISE + decim: average every d consecutive N-bin spectra; `acc` (2*N float2, device) carries the open group.
int spectra_decimate_launch(const float2* spec, long long n_frames, int N, int d, int carried, float2* acc, float2* out,
                            long long n_groups, int sm_count, cudaStream_t stream);
// Copies the last n_tail samples of the window into dst (device), double-buffer safe.
int window_tail_copy(const Window& w, long long n_tail, float2* dst, cudaStream_t stream);
int window_tail_copy_multi(const Window* wins, const long long* n_tail, float2* const* dst, int n, void* scratch,
                           cudaStream_t stream);
size_t window_tail_multi_scratch_bytes(int n);

// ---- CHIRP launches (chirp_kernels.cu) --------------------------------------------------------
struct ChirpDev {   // kernel-side copy of gsdr_chirp_param plus derived 32-bit constants
    unsigned long long period;   // num_steps*length
    unsigned long long length;
    unsigned int chirpness;
    int f0;
    unsigned long long num_steps;  // period / length
};
int chirp_demod_launch(const Window& w, unsigned long long pos0 /* chirp position of window sample 0 */,
                       const ChirpDev& cp, const float* profile, int side /* >=0: flat window, skip [0,side) */,
                       float flat_weight, int ppt, long long n_out, float2* out, float2* partial, int sm_count,
                       cudaStream_t stream);
int chirp_demod_full_launch(const float2* in, long long n, unsigned long long pos0, const ChirpDev& cp, float2* out,
                            cudaStream_t stream);
int chirp_gen_launch(float2* out, long long n, unsigned long long pos0, const ChirpDev& cp, float scale, cudaStream_t stream);
int chirp_index_probe_launch(int* out, unsigned int n, unsigned long long pos0, const ChirpDev& cp, cudaStream_t stream);
size_t chirp_partial_count(int ppt, long long n_out, int sm_count);

// ---- DIRECT launches (direct_kernels.cu) ------------------------------------------------------
int direct_fir_launch(const Window& w, const float2* g /* [T][ntaps] */, const int* freq_dev, int T, int M, int ntaps,
                      int rate, long long pos0 /* stream position (mod rate) of window sample 0 */, long long n_out,
                      float2* out, cudaStream_t stream);
// tcgen05 / TMEM version of direct_fir_launch (direct_tc_kernels.cu): 3xTF32 split GEMM, pf_average in {1,2,4,8}
bool direct_fir_tc_supported(int T, int M, int ntaps, long long n_out);
bool direct_fir_tc_preferred(int T, int M, int ntaps, long long n_out);
// rotate = 0: no LO rotation of the outputs (freq_dev unused) -- the form the generic-size polyphase channelizer takes.
// allow_tma = false: the window rows come through plain loads instead of TMA boxes (GSDR_DIRECT_TC_TMA=0 at create; windows
// in pinned host memory unless GSDR_DIRECT_TC_HOST_TMA=1).
int direct_fir_tc_launch(const Window& w, const float2* g, const int* freq_dev, int T, int M, int ntaps, int rate, long long pos0,
                         long long n_out, float2* out, int sm_count, cudaStream_t stream, int rotate = 1, bool allow_tma = true);
// exact integer version (direct_i8_kernels.cu): kind::i8 MMAs on 24-bit fixed-point operands cut into three signed digits,
// int32 accumulators, 64-bit combine -- no accumulation error at all.  The filter bank is prepared once per demodulator.
struct DirectI8Bank {
    void* d_bank = nullptr;        // digit planes of the filter bank, rows of 128 bytes (see direct_i8_bank_create)
    float* d_inv_sb = nullptr;     // [tone groups * TG] 1 / (fixed-point scale of the tone's taps)
    int KQ = 0, tone_groups = 0;   // quads of K blocks per tile, tone groups
    float* d_row_amax = nullptr;   // scratch: largest |sample| per row tile (launches with more than two tone groups)
    size_t row_amax_cap = 0;
    alignas(64) unsigned char tmap_b[128];   // CUtensorMap of the bank
};
bool direct_fir_i8_supported(int T, int M, int ntaps, long long n_out);
bool direct_fir_i8_preferred(int T, int M, int ntaps, long long n_out);
int direct_i8_bank_create(const double* g /* [T][ntaps] (re, im) */, int T, int M, int ntaps, DirectI8Bank* bank);
void direct_i8_bank_destroy(DirectI8Bank* bank);
int direct_fir_i8_launch(DirectI8Bank& bank, const Window& w, const int* freq_dev, int T, int M, int ntaps, int rate, long long pos0,
                         long long n_out, float2* out, int sm_count, cudaStream_t stream, int rotate = 1, bool allow_tma = true);
int direct_mix_launch(const float2* in, long long n, const int* freq_dev, int T, int rate, long long pos0, float2* out,
                      cudaStream_t stream);
int direct_tile_phase_probe_launch(long long* phase, unsigned int* word, int n_rows, int tone_freq, int rate, long long pos0,
                                   long long row0, int M, cudaStream_t stream);
int direct_phase_probe_launch(long long* out, unsigned int n, int tone_freq, int rate, unsigned long long index_counter,
                              unsigned long long n0, cudaStream_t stream);

// ---- TX tones synthesis (tones_kernels.cu) ----------------------------------------------------
int tones_synth_launch(float2* out, long long n0, long long n, const int* bins_dev, const float* ampl_dev, int T, int rate,
                       cudaStream_t stream);

}  // namespace gsdr
