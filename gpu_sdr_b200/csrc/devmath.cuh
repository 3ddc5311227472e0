// Device-side helpers shared by the DSP kernels: window access, complex arithmetic, and a
// float32 sincos driven by an INTEGER phase so that range reduction is exact.
#pragma once
#include <cuda_runtime.h>

#include "common.hpp"

namespace gsdr {

__device__ __forceinline__ float2 dev_win_at(const Window& w, long long s) {
    if (s < w.n_hist) return w.hist[s];
    s -= w.n_hist;
    if (s < w.n_in) return w.in[s];
    return make_float2(0.f, 0.f);
}

__device__ __forceinline__ float2 dev_cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// sin and cos of 2*pi*phase/2^32 for a 32-bit phase word.
// Quadrant (top two bits, rounded) is peeled off in integers, so the polynomial only ever sees
// |x| <= pi/4 and the int->float conversion error is <= 2^-24 * pi/4 ~ 4.7e-8 rad.  Polynomials
// are the classic single-precision minimax kernels (error ~1 ulp on the reduced range).
__device__ __forceinline__ void sincos_phase32(unsigned int phase, float& s, float& c) {
    const unsigned int u = phase + 0x20000000u;
    const unsigned int q = u >> 30;
    const int rem = (int)(u & 0x3FFFFFFFu) - 0x20000000;       // [-2^29, 2^29)
    const float x = (float)rem * 1.4629180792671596e-9f;       // 2*pi / 2^32
    const float z = x * x;
    float sp = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    const float sn = fmaf(x * z, sp, x);
    float cp = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    const float cs = fmaf(z * z, cp, fmaf(-0.5f, z, 1.0f));
    // rotate by q quarter turns
    const float s1 = (q & 1u) ? cs : sn;
    const float c1 = (q & 1u) ? sn : cs;
    s = (q & 2u) ? -s1 : s1;
    c = ((q + 1u) & 2u) ? -c1 : c1;
}

// ---- chirp integer phase (cpp/kernels.cu:401-419 == :348-365) ---------------------------------
// index = eff*(f0 + k*chirpness) - chirpness*(length*T(k)), T(k) = k(k+1)/2, evaluated by the
// reference in unsigned 64-bit and truncated to int32.  Truncation commutes with + - *, so the
// same bits come out of 32-bit arithmetic as long as T(k) itself is formed exactly (64-bit).
__device__ __forceinline__ unsigned int chirp_index_word(unsigned long long eff, unsigned long long k, const ChirpDev& cp) {
    const unsigned long long tri = (k & 1ull) ? ((k + 1ull) >> 1) * k : (k >> 1) * (k + 1ull);  // exact for k < 2^32
    const unsigned int step = (unsigned int)cp.f0 + (unsigned int)k * cp.chirpness;
    const unsigned int corr = cp.chirpness * ((unsigned int)cp.length * (unsigned int)tri);
    return (unsigned int)eff * step - corr;
}

// The reference feeds (double)index / 2147483647.5 to sinpi/cospi: one turn is 2^32-1 counts, not
// 2^32.  Treating the int32 index as a 2^32-per-turn phase word is off by |index| * 2.3e-10
// relative, at most 1.5e-9 rad -- forty times below float32 resolution of the result.
// Returns chirp = (sin(pi*theta), -cos(pi*theta)).
__device__ __forceinline__ float2 chirp_phasor(unsigned int index_word) {
    float s, c;
    sincos_phase32(index_word, s, c);
    return make_float2(s, -c);
}

// Walks chirp positions pos, pos+stride, pos+2*stride ... : inside a frequency step the index is a
// phase accumulator (one 32-bit multiply-add).  Crossing a step boundary costs one 32-bit division and a
// re-evaluation of the closed form -- no 64-bit division -- so short steps (a true chirp has length 1: every
// advance crosses) stay cheap.  64-bit divisions only run in seek() and when the sweep wraps around.
struct ChirpWalker {
    unsigned long long k;        // step index
    unsigned int r, len32;       // offset inside the step, step length (clamped to 2^31-1)
    unsigned int idx, step;
    bool huge;                   // step length does not fit 31 bits: the walk never leaves the step in practice
    unsigned long long eff0, adv;  // only used when `huge`
    __device__ __forceinline__ void seek(unsigned long long pos, const ChirpDev& cp) {
        const unsigned long long eff = pos % cp.period;
        k = eff / cp.length;
        huge = cp.length > 0x7fffffffull;
        r = huge ? 0u : (unsigned int)(eff - k * cp.length);
        len32 = huge ? 0x7fffffffu : (unsigned int)cp.length;
        eff0 = eff;
        adv = 0;
        idx = chirp_index_word(eff, k, cp);
        step = (unsigned int)cp.f0 + (unsigned int)k * cp.chirpness;
    }
    __device__ __forceinline__ void advance(unsigned int stride, const ChirpDev& cp) {
        if (huge) {  // rare configuration (a step longer than 2^31 samples): exact but slow
            adv += stride;
            seek(eff0 + adv, cp);
            return;
        }
        r += stride;
        if (r < len32) {
            idx += stride * step;
            return;
        }
        const unsigned int dk = (len32 == 1u) ? r : r / len32;  // a true chirp (one sample per step) needs no division
        r -= dk * len32;
        k += dk;
        if (k >= cp.num_steps) k %= cp.num_steps;
        // chirp_index_word only uses the low 32 bits of the position
        idx = chirp_index_word((unsigned long long)((unsigned int)k * len32 + r), k, cp);
        step = (unsigned int)cp.f0 + (unsigned int)k * cp.chirpness;
    }
};

}  // namespace gsdr
