// Host-side logic of the path that *defines results*: FIR/PFB tap builders, carry-over
// bookkeeping, tone->bin mapping and chirp parameter quantisation.  These must agree with the
// reference bit for bit (taps: float32; the rest: integers), so each routine documents the
// reference lines whose arithmetic it reproduces (paths relative to the reference tree).
// Built WITHOUT fast-math / FMA contraction (see Makefile) for the same reason.
#include <cmath>
#include <limits>

#include "common.hpp"

namespace gsdr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static const float kPiF = 3.14159265358979f;  // headers/kernels.cuh:34 (float literal)

// Hamming-windowed sinc prototype, sum-normalised.  cpp/kernels.cu:258-310.
//  - centre index is (length-1)/2 in INTEGER arithmetic (even lengths are off-centre);
//  - sinc and cosine are evaluated in float32 (sinf/cosf), the Hamming factor in double;
//  - the normalising sum is accumulated in float32 in index order.
void make_sinc_window(int length, float fc, float* out) {
    const int centre = (length - 1) / 2;
    float sum = 0.f;
    for (int i = 0; i < length; ++i) {
        const int d = i - centre;
        float tap = 2.f * fc;
        if (d != 0) {
            const float arg = 2.f * kPiF * fc * d;
            tap = (2.f * fc) * sinf(arg) / arg;
        }
        const float c = cosf(2.f * kPiF * i / (length - 1));
        tap = static_cast<float>(static_cast<double>(tap) * (0.54 - 0.46 * static_cast<double>(c)));
        out[i] = tap;
        sum += tap;
    }
    for (int i = 0; i < length; ++i) out[i] /= sum;
}

// Lock-in profile.  cpp/kernels.cu:208-253: only the first `side` taps end up zero (the loop
// that clears the trailing `side` taps runs before the fill loop, which overwrites them).
void make_flat_window(int length, int side, float* out) {
    float sum = 0.f;
    for (int i = 0; i < length; ++i) out[i] = 0.f;
    for (int i = side; i < length; ++i) {
        out[i] = 1.f;
        sum += out[i];
    }
    for (int i = 0; i < length; ++i) out[i] /= sum;
}

// cpp/USRP_demodulator.cpp:706
int pfb_batching(int buffer_len, int fft_tones, int pf_average) {
    return static_cast<int>(std::ceil(static_cast<float>(buffer_len) / static_cast<float>(fft_tones)) + pf_average + 5);
}

// cpp/USRP_demodulator.cpp:722-734.  The reference scans the whole bin axis for every tone and
// keeps the LAST axis point whose open interval (axis-bs, axis+bs) holds the tone, which makes
// the rule ceil(f/bs) rather than round(f/bs); the double comparisons are kept literally so the
// edge cases (exact multiples, truncated negative tones) fall the same way.  -1 = no match.
void tone_bins(int rate, int fft_tones, const int32_t* freq, int n, int32_t* bins) {
    const double bs = static_cast<double>(rate) / static_cast<double>(fft_tones);
    const int half = fft_tones / 2;
    for (int u = 0; u < n; ++u) {
        int32_t hit = -1;
        for (int i = 0; i < fft_tones; ++i) {
            const double axis = i * bs - bs * half;
            if (freq[u] < axis + bs && freq[u] > axis - bs) hit = (i + half) % fft_tones;
        }
        bins[u] = hit;
    }
}

// cpp/USRP_server_memory_management.cpp:104-156
static int frames_available(const gsdr_buffer_helper* h) {
    int frames = 0;
    for (int off = 0; off + h->average * h->n_tones < h->eff_length; off += h->n_tones) ++frames;
    return frames;
}
void buffer_helper_init(gsdr_buffer_helper* h, int n_tones, int buffer_len, int average, int n_eff_tones) {
    h->n_tones = n_tones;
    h->buffer_len = buffer_len;
    h->average = average;
    h->n_eff_tones = n_eff_tones;
    h->eff_length = buffer_len;
    h->new_0 = 0;
    h->current_batch = frames_available(h);
    h->spare_samples = h->eff_length - h->current_batch * n_tones;
    h->spare_begin = h->eff_length - h->spare_samples;
    h->copy_size = n_eff_tones * h->current_batch;
}
void buffer_helper_update(gsdr_buffer_helper* h) {
    h->new_0 = h->spare_samples;
    h->eff_length = h->spare_samples + h->buffer_len;
    h->current_batch = frames_available(h);
    h->copy_size = h->n_eff_tones * h->current_batch;
    h->spare_samples = h->eff_length - h->current_batch * h->n_tones;
    h->spare_begin = h->eff_length - h->spare_samples;
}

// cpp/USRP_server_memory_management.cpp:30-56
static void vna_refresh(gsdr_vna_helper* h) {
    h->valid_size = h->total_len / h->ppt;
    h->new0 = h->total_len - h->ppt * h->valid_size;
    h->spare_begin = h->total_len - h->new0;
}
void vna_helper_init(gsdr_vna_helper* h, int ppt, int buffer_len) {
    h->ppt = ppt;
    h->buffer_len = buffer_len;
    h->total_len = buffer_len;
    vna_refresh(h);
}
void vna_helper_update(gsdr_vna_helper* h) {
    h->total_len = h->buffer_len + h->new0;
    vna_refresh(h);
}

// Conversions with the wrap behaviour the reference gets from gcc/x86-64: double -> unsigned int
// goes through a 64-bit truncation (so a negative chirpness wraps modulo 2^32), double -> int
// out of range gives INT_MIN.
static uint32_t wrap_u32(double v) {
    if (!(v > -9.2e18 && v < 9.2e18)) return 0u;
    return static_cast<uint32_t>(static_cast<uint64_t>(static_cast<int64_t>(v)));
}
static int32_t trunc_i32(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return std::numeric_limits<int32_t>::min();
    return static_cast<int32_t>(v);
}

// RX: cpp/USRP_demodulator.cpp:192-214.  TX: cpp/USRP_buffer_generator.cpp:114-137 (the TX side
// additionally resets num_steps when a step would be shorter than one sample).
void chirp_params(int rate, int freq0, int chirp_f0, int swipe_s0, float chirp_t0, bool tx, gsdr_chirp_param* out) {
    uint64_t steps = static_cast<uint64_t>(static_cast<int64_t>(swipe_s0));
    if (steps < 1) steps = static_cast<uint64_t>(chirp_t0 * static_cast<float>(rate));
    uint64_t len = static_cast<uint64_t>(chirp_t0 * static_cast<float>(rate) / static_cast<float>(steps));
    if (len < 1) {
        len = 1;
        if (tx) steps = static_cast<uint64_t>(chirp_t0 * static_cast<float>(rate));
    }
    const double full = std::pow(2, 32) - 1;
    out->num_steps = steps;
    out->length = len;
    out->chirpness = wrap_u32((full * (chirp_f0 - freq0) / (static_cast<double>(steps) - 1.)) / static_cast<double>(rate));
    out->f0 = trunc_i32(full * (static_cast<double>(freq0) / static_cast<double>(rate)));
}

}  // namespace gsdr
