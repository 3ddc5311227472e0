// Packed complex arithmetic for sm_100a: a complex float lives in one 64-bit register pair and the
// component-wise operations map to the Blackwell packed-fp32 instructions (FADD2 / FMUL2 / FFMA2 in
// SASS, `add/mul/fma.rn.f32x2` in PTX).  One issue slot does the work of two: the FFT butterflies of
// the channelizer (cpp/USRP_demodulator.cpp:501 -- cufftExecC2C in the reference) and its polyphase
// FIR (cpp/kernels.cu:495-506) are issue-bound on CUDA cores, not FMA-pipe bound.
//
// Every packed op is the same IEEE fp32 operation per component as its scalar twin.
#pragma once
#include <cuda_runtime.h>

namespace gsdr {

typedef unsigned long long c2;  // {re (low word), im (high word)}

__device__ __forceinline__ c2 c2_pack(float x, float y) {
    c2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void c2_unpack(c2 a, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
}
__device__ __forceinline__ c2 c2_from(float2 a) { return c2_pack(a.x, a.y); }
__device__ __forceinline__ float2 c2_to(c2 a) {
    float2 r;
    c2_unpack(a, r.x, r.y);
    return r;
}
__device__ __forceinline__ c2 c2_add(c2 a, c2 b) {
    c2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c2 c2_sub(c2 a, c2 b) {
    c2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c2 c2_mul(c2 a, c2 b) {
    c2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c2 c2_fma(c2 a, c2 b, c2 c) {
    c2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// scale both components by a real scalar (ptxas folds the {s, s} pair into the broadcast operand form)
// The {s, s} pair is formed inside the asm statement so that loop-invariant code motion cannot hoist it
// into a live register pair (which would double the register cost of every loop-invariant scalar).
__device__ __forceinline__ c2 c2_scale(c2 a, float s) {
    c2 r;
    asm("{\n\t.reg .b64 ss;\n\tmov.b64 ss, {%2, %2};\n\tmul.rn.f32x2 %0, %1, ss;\n\t}" : "=l"(r) : "l"(a), "f"(s));
    return r;
}
__device__ __forceinline__ c2 c2_fma_s(c2 a, float s, c2 c) {
    c2 r;
    asm("{\n\t.reg .b64 ss;\n\tmov.b64 ss, {%2, %2};\n\tfma.rn.f32x2 %0, %1, ss, %3;\n\t}" : "=l"(r) : "l"(a), "f"(s), "l"(c));
    return r;
}

// Multiplication by -j / +j: a swap of the halves plus one sign flip.  ptxas folds the pair below into OPERAND
// MODIFIERS of the consuming packed instruction (SASS `R.F32x2.LO_HI.NP`), so quarter-turn rotations are free.
__device__ __forceinline__ c2 c2_mj(c2 a) {
    float x, y;
    c2_unpack(a, x, y);
    return c2_pack(y, -x);
}
__device__ __forceinline__ c2 c2_pj(c2 a) {
    float x, y;
    c2_unpack(a, x, y);
    return c2_pack(-y, x);
}
// Packed instructions execute on the fmaheavy pipe only, two cycles each; scalar FADD/FFMA can also use fmalite.
// An all-packed FFT therefore trades issue slots for fmaheavy cycles.  GSDR_PK_MODE picks the mix (measured on the
// fused PFB kernel, bench.py --profile): bit 0 = butterflies fully packed (8 instructions instead of 6 packed + 4
// scalar), bit 1 = twiddles fully packed (2 instructions instead of 1 packed + 2 scalar).
#ifndef GSDR_PK_MODE
#define GSDR_PK_MODE 0
#endif
// a * (bx + j by)
__device__ __forceinline__ c2 c2_cmul(c2 a, float bx, float by) {
#if GSDR_PK_MODE & 2
    return c2_fma_s(c2_pj(a), by, c2_scale(a, bx));
#elif GSDR_PK_MODE & 4
    float ax, ay;
    c2_unpack(a, ax, ay);
    return c2_pack(fmaf(-ay, by, ax * bx), fmaf(ax, by, ay * bx));
#else
    float ax, ay, mx, my;
    c2_unpack(a, ax, ay);
    c2_unpack(c2_scale(a, bx), mx, my);
    return c2_pack(fmaf(-ay, by, mx), fmaf(ax, by, my));
#endif
}
// a * sqrt(1/2) (1 - j)  =  W8^1
__device__ __forceinline__ c2 c2_mul_w8_1(c2 a) {
#if GSDR_PK_MODE & 2
    return c2_fma_s(c2_mj(a), 0.70710678118654752f, c2_scale(a, 0.70710678118654752f));
#else
    float mx, my;
    c2_unpack(c2_scale(a, 0.70710678118654752f), mx, my);
    return c2_pack(mx + my, my - mx);
#endif
}
// a * sqrt(1/2) (-1 - j)  =  W8^3
__device__ __forceinline__ c2 c2_mul_w8_3(c2 a) {
#if GSDR_PK_MODE & 2
    return c2_fma_s(c2_mj(a), 0.70710678118654752f, c2_scale(a, -0.70710678118654752f));
#else
    float mx, my;
    c2_unpack(c2_scale(a, 0.70710678118654752f), mx, my);
    return c2_pack(my - mx, -mx - my);
#endif
}

// t + (-j) d and t - (-j) d
__device__ __forceinline__ void c2_rot_pair(c2 t, c2 d, c2& plus, c2& minus) {
#if GSDR_PK_MODE & 1
    const c2 r = c2_mj(d);
    plus = c2_add(t, r);
    minus = c2_sub(t, r);
#else
    float tx, ty, dx, dy;
    c2_unpack(t, tx, ty);
    c2_unpack(d, dx, dy);
    plus = c2_pack(tx + dy, ty - dx);
    minus = c2_pack(tx - dy, ty + dx);
#endif
}

// 4-point forward DFT in place, natural order out.
__device__ __forceinline__ void c2_fft4(c2& a0, c2& a1, c2& a2, c2& a3) {
    const c2 t0 = c2_add(a0, a2), t1 = c2_sub(a0, a2), t2 = c2_add(a1, a3), d = c2_sub(a1, a3);
#if GSDR_PK_MODE & 16
    float ux, uy, vx, vy;
    c2_unpack(t0, ux, uy);
    c2_unpack(t2, vx, vy);
    a0 = c2_pack(ux + vx, uy + vy);
    a2 = c2_pack(ux - vx, uy - vy);
#else
    a0 = c2_add(t0, t2);
    a2 = c2_sub(t0, t2);
#endif
    c2_rot_pair(t1, d, a1, a3);
}
// same with a2 pre-multiplied by -j (folds the W16^4 twiddle of the 16-point transform)
__device__ __forceinline__ void c2_fft4_a2mj(c2& a0, c2& a1, c2& a2, c2& a3) {
    c2 t0, t1;
    c2_rot_pair(a0, a2, t0, t1);
    const c2 t2 = c2_add(a1, a3), d = c2_sub(a1, a3);
    a0 = c2_add(t0, t2);
    a2 = c2_sub(t0, t2);
    c2_rot_pair(t1, d, a1, a3);
}

#define GSDR_PK_C1 0.92387953251128674f /* cos(pi/8) */
#define GSDR_PK_S1 0.38268343236508977f /* sin(pi/8) */

// 16-point forward DFT in two halves.  Input v[n] natural; output X[ka + 4 kb] lands in v[4 ka + kb].
// After the first half every input has been consumed (a caller may release the buffer it was read from).
__device__ __forceinline__ void c2_fft16_first(c2 (&v)[16]) {
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) c2_fft4(v[nb], v[4 + nb], v[8 + nb], v[12 + nb]);
}
__device__ __forceinline__ void c2_fft16_second(c2 (&v)[16]) {
    // v[4 ka + nb] *= W16^(nb ka); W16^4 on v[10] is folded into the third second-level butterfly
    v[5] = c2_cmul(v[5], GSDR_PK_C1, -GSDR_PK_S1);    // W^1
    v[6] = c2_mul_w8_1(v[6]);                          // W^2
    v[7] = c2_cmul(v[7], GSDR_PK_S1, -GSDR_PK_C1);    // W^3
    v[9] = c2_mul_w8_1(v[9]);                          // W^2
    v[11] = c2_mul_w8_3(v[11]);                        // W^6
    v[13] = c2_cmul(v[13], GSDR_PK_S1, -GSDR_PK_C1);  // W^3
    v[14] = c2_mul_w8_3(v[14]);                        // W^6
    v[15] = c2_cmul(v[15], -GSDR_PK_C1, GSDR_PK_S1);  // W^9
    c2_fft4(v[0], v[1], v[2], v[3]);
    c2_fft4(v[4], v[5], v[6], v[7]);
    c2_fft4_a2mj(v[8], v[9], v[10], v[11]);
    c2_fft4(v[12], v[13], v[14], v[15]);
}
__device__ __forceinline__ void c2_fft16(c2 (&v)[16]) {
    c2_fft16_first(v);
    c2_fft16_second(v);
}

// 8-point forward DFT.  Input v[n] natural (n = 2 na + nb); output X[ka + 4 kb] in v[2 ka + kb].
__device__ __forceinline__ void c2_fft8(c2 (&v)[8]) {
    c2_fft4(v[0], v[2], v[4], v[6]);
    c2_fft4(v[1], v[3], v[5], v[7]);
    v[3] = c2_mul_w8_1(v[3]);
    v[7] = c2_mul_w8_3(v[7]);
    {
        const c2 a = v[0], b = v[1];
        v[0] = c2_add(a, b);
        v[1] = c2_sub(a, b);
    }
    {
        const c2 a = v[2], b = v[3];
        v[2] = c2_add(a, b);
        v[3] = c2_sub(a, b);
    }
    c2_rot_pair(v[4], v[5], v[4], v[5]);  // odd input carries W8^2 = -j
    {
        const c2 a = v[6], b = v[7];
        v[6] = c2_add(a, b);
        v[7] = c2_sub(a, b);
    }
}

__device__ __forceinline__ c2 lds_c2(const float2* p) { return *reinterpret_cast<const c2*>(p); }
__device__ __forceinline__ void sts_c2(float2* p, c2 v) { *reinterpret_cast<c2*>(p) = v; }

}  // namespace gsdr
