/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the integer / tap-building parts
 * of zjc263/GPU_SDR's RX/TX DSP path.  Never linked into, imported by or called from the
 * product (gpu_sdr_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker.
 *
 * Plain C (gcc + libm) on purpose: the reference builds its FIR/PFB taps in *host* code with
 * float32 sinf/cosf and float accumulation (cpp/kernels.cu:208-310), so a bit-exact restatement
 * has to go through the same libm, and the integer bookkeeping relies on C conversion / `%`
 * semantics (sign of the remainder, double->unsigned wrap) that are easiest to keep exact in C.
 * The streaming floating-point chains (mix, FIR, PFB, FFT, lock-in) are restated in fp64 NumPy
 * in oracle/gsdr_oracle.py on top of these functions.
 *
 * Parity pin: every function here is checked against the reference's own object code
 * (oracle/_ref/libgsdr_ref.so, built from the unmodified sources under /root/reference) by
 * tests/test_oracle_vs_ref.py when that library is present, and against the committed
 * fixtures under tests/golden/ (generated from that library by tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI_F 3.14159265358979f /* headers/kernels.cuh:34 `pi_f` (a float literal) */

/* cpp/kernels.cu:258-310 make_sinc_window: Hamming-windowed sinc, float32 arithmetic where the
 * reference has it (sinf/cosf resolve to the float overloads in nvcc host code), double for the
 * `0.54-0.46*cos()` factor, float running sum, float divide.  Only .x is returned (.y == 0). */
void orc_make_sinc_window(int length, float fc, float *out) {
    float scale = 0;
    for (int i = 0; i < length; i++) {
        int sinc_index = i - (length - 1) / 2; /* integer division: asymmetric for even length */
        float v;
        if (sinc_index != 0) {
            float a = 2.f * ORC_PI_F * fc * sinc_index;
            v = (2.f * fc) * sinf(a) / a;
        } else {
            v = (2.f * fc);
        }
        /* h_win[i].x *= (0.54-0.46*cos(2.f*pi_f*i/(length-1))): float cosf, double combine */
        float c = cosf(2.f * ORC_PI_F * i / (length - 1));
        v = (float)((double)v * (0.54 - 0.46 * (double)c));
        out[i] = v;
        scale += v;
    }
    for (int i = 0; i < length; i++) out[i] /= scale;
}

/* cpp/kernels.cu:208-253 make_flat_window: the leading `side` taps are zero; the trailing
 * zeroing loop (:223-226) is overwritten by the fill loop (:227-233) which runs i+side up to
 * length-1, so every tap from `side` on is 1/(length-side). */
void orc_make_flat_window(int length, int side, float *out) {
    float scale = 0;
    for (int i = 0; i < side; i++) out[i] = 0;
    for (int i = length - side; i < length; i++) if (i >= 0) out[i] = 0;
    for (int i = 0; i < length - side; i++) {
        out[i + side] = 1.f;
        scale += out[i + side];
    }
    for (int i = 0; i < length; i++) out[i] /= scale;
}

/* cpp/USRP_server_memory_management.cpp:104-156 buffer_helper (PFB carry-over bookkeeping). */
typedef struct {
    int n_tones, eff_length, buffer_len, average, n_eff_tones;
    int new_0, copy_size, current_batch, spare_samples, spare_begin;
} orc_buffer_helper;

static int orc_bh_simulate(const orc_buffer_helper *h) { /* :145-156 */
    int offset = 0, batching = 0;
    while (offset + h->average * h->n_tones < h->eff_length) {
        offset += h->n_tones;
        batching++;
    }
    return batching;
}

void orc_buffer_helper_init(orc_buffer_helper *h, int n_tones, int buffer_len, int average, int n_eff_tones) {
    h->n_tones = n_tones; /* :104-124 */
    h->buffer_len = buffer_len;
    h->average = average;
    h->n_eff_tones = n_eff_tones;
    h->eff_length = buffer_len;
    h->current_batch = orc_bh_simulate(h);
    h->spare_samples = h->eff_length - h->current_batch * n_tones;
    h->spare_begin = h->eff_length - h->spare_samples;
    h->new_0 = 0;
    h->copy_size = n_eff_tones * h->current_batch;
}

void orc_buffer_helper_update(orc_buffer_helper *h) { /* :126-142 */
    h->new_0 = h->spare_samples;
    h->eff_length = h->spare_samples + h->buffer_len;
    h->current_batch = orc_bh_simulate(h);
    h->copy_size = h->n_eff_tones * h->current_batch;
    h->spare_samples = h->eff_length - h->current_batch * h->n_tones;
    h->spare_begin = h->eff_length - h->spare_samples;
}

/* cpp/USRP_server_memory_management.cpp:30-56 VNA_decimator_helper. */
typedef struct { int valid_size, new0, total_len, spare_begin, ppt, buffer_len; } orc_vna_helper;

void orc_vna_helper_init(orc_vna_helper *h, int ppt, int buffer_len) {
    h->ppt = ppt;
    h->buffer_len = buffer_len;
    h->total_len = buffer_len;
    h->valid_size = h->total_len / ppt;
    h->new0 = h->total_len - ppt * h->valid_size;
    h->spare_begin = h->total_len - h->new0;
}

void orc_vna_helper_update(orc_vna_helper *h) {
    h->total_len = h->buffer_len + h->new0;
    h->valid_size = h->total_len / h->ppt;
    h->new0 = h->total_len - h->ppt * h->valid_size;
    h->spare_begin = h->total_len - h->new0;
}

/* cpp/USRP_demodulator.cpp:706 batching = ceil((float)L/(float)N) + P + 5 */
int orc_pfb_batching(int buffer_len, int fft_tones, int pf_average) {
    return (int)(ceilf((float)buffer_len / (float)fft_tones) + pf_average + 5);
}

/* cpp/USRP_demodulator.cpp:722-734 tone -> FFT bin.  Literal restatement of the double loop:
 * the LAST bin-axis point whose open interval (axis-bs, axis+bs) contains the tone wins.
 * Unmatched tones are left at -1 here (uninitialised malloc in the reference). */
void orc_tone_bins(int rate, int fft_tones, const int *freq, int n, int *bins) {
    double bin_size = (double)rate / (double)fft_tones;
    for (int u = 0; u < n; u++) bins[u] = -1;
    for (int i = 0; i < fft_tones; i++) {
        double axis = i * bin_size - bin_size * (fft_tones / 2);
        for (int u = 0; u < n; u++) {
            if ((freq[u] < axis + bin_size) && (freq[u] > axis - bin_size))
                bins[u] = (i + (fft_tones / 2)) % fft_tones;
        }
    }
}

/* cpp/USRP_demodulator.cpp:192-214 (RX) and cpp/USRP_buffer_generator.cpp:114-137 (TX):
 * chirp parameters.  `tx` selects the TX variant, which also resets num_steps when length<1.
 * double -> unsigned int of a negative value is UB in C; gcc/x86-64 converts through a 64-bit
 * cvttsd2si and keeps the low 32 bits, which is what the int64 detour below does portably.
 * double -> int out of range yields INT_MIN on x86 (cvttsd2si "indefinite"). */
typedef struct { uint64_t num_steps, length; uint32_t chirpness; int32_t f0; } orc_chirp_param;

static uint32_t orc_d2u32(double v) {
    if (!(v > -9.2e18 && v < 9.2e18)) return 0u;
    return (uint32_t)(uint64_t)(int64_t)v;
}
static int32_t orc_d2i32(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return (int32_t)v;
}

void orc_chirp_params(int rate, int freq0, int chirp_f0, int swipe_s0, float chirp_t0, int tx, orc_chirp_param *p) {
    p->num_steps = (uint64_t)(int64_t)swipe_s0; /* int -> unsigned long */
    if (p->num_steps < 1) p->num_steps = (uint64_t)(chirp_t0 * (float)rate);
    /* float * int / unsigned long: all float arithmetic, then truncation */
    p->length = (uint64_t)(chirp_t0 * (float)rate / (float)p->num_steps);
    if (p->length < 1) {
        p->length = 1;
        if (tx) p->num_steps = (uint64_t)(chirp_t0 * (float)rate);
    }
    double c = ((pow(2, 32) - 1) * (chirp_f0 - freq0) / ((double)p->num_steps - 1.)) / (double)rate;
    p->chirpness = orc_d2u32(c);
    p->f0 = orc_d2i32((pow(2, 32) - 1) * ((double)freq0 / (double)rate));
}

/* cpp/kernels.cu:401-419 (chirp_demodulator) == :348-365 (chirp_gen): the int32 phase index of
 * sample `offset` of a buffer that starts at stream position last_index.  All arithmetic is
 * unsigned 64-bit, truncated to int on assignment. */
void orc_chirp_index(uint64_t last_index, uint32_t n, const orc_chirp_param *p, int32_t *index) {
    for (uint32_t offset = 0; offset < n; offset++) {
        uint64_t effective_index = (last_index + offset) % (p->num_steps * p->length);
        uint64_t frequency_index = effective_index / p->length;
        uint64_t q_phase = (frequency_index / 2) * (frequency_index + 1) + (frequency_index % 2) * ((frequency_index + 1) / 2);
        uint64_t phase_correction = (uint64_t)p->chirpness * (p->length * q_phase);
        uint64_t f0u = (uint64_t)(int64_t)p->f0; /* int promoted to unsigned long */
        index[offset] = (int32_t)(uint32_t)(effective_index * (f0u + frequency_index * (uint64_t)p->chirpness) - phase_correction);
    }
}

/* cpp/kernels.cu:63-68 direct_demodulator_integer: my_phase = tp + (tf*ii) % wavetablelen with
 * ii = (n + index_counter) % wavetablelen, in (signed) long long -- C remainder keeps the sign
 * of the dividend, so negative tones give my_phase in (-R, 0]. */
void orc_direct_phase(int tf, int tp, int wavetablelen, uint64_t index_counter, uint64_t n0, uint64_t n, int64_t *phase) {
    for (uint64_t i = 0; i < n; i++) {
        long long ii = (long long)((n0 + i + index_counter) % (uint64_t)wavetablelen);
        phase[i] = (long long)tp + ((long long)tf * ii) % (long long)wavetablelen;
    }
}
