#pragma once
// TONES-mode polyphase-filter-bank channelizer for sm_100a: the fused kernels (shared by pfb_kernels.cu, which holds the
// lock-step kernel's launcher, the generic path and pfb_launch, and by pfb_wsp_p{1,2,3,4}.cu, one translation unit per tap
// count of the warp-specialised kernel so that its 32 instantiations compile in parallel).
//
// Replaces the reference's four-kernel chain (cpp/USRP_demodulator.cpp:486-565):
//   polyphase_filter (cpp/kernels.cu:474-516) -> cufftExecC2C(FORWARD) -> tone_select
//   (cpp/kernels.cu:531-554) -> move_buffer (cpp/kernels.cu:444-459)
// with ONE fused persistent kernel: P-tap polyphase FIR in registers (sliding row window),
// in-shared-memory 2048-point FFT (16 x 16 x 8, three register passes, two exchanges), and a
// gather of the selected bins straight into the sample-major output.  The 8 B/sample
// intermediate the reference writes and re-reads twice never leaves the SM.
//
//   y[b][k]   = sum_{i<P} x[(b+i)N + k] * w[iN + k]         (real taps)
//   Y[b][k']  = sum_k y[b][k] exp(-2 pi j k k'/N)           (unnormalised, forward)
//   out[b*T+u] = Y[b][bins[u]]
//
// A generic two-kernel path (FIR, then a direct DFT of the selected bins only) serves every
// (N, P, T) the fused kernel is not instantiated for; it is also CUDA -- there is no CPU path.
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>
#include <utility>

#include <vector>

#include "common.hpp"
#include "packed_f32x2.cuh"

namespace gsdr {

namespace {

// --------------------------------------------------------------------------------------------
// window access (carry-over samples ++ new samples)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 win_at(const Window& w, long long s) {
    if (s < w.n_hist) return w.hist[s];
    s -= w.n_hist;
    if (s < w.n_in) return w.in[s];
    return make_float2(0.f, 0.f);
}

// sc16 ingest fused into the channelizer: the `in` segment then holds interleaved int16 I/Q pairs (the USRP wire format)
// and every sample is converted on the fly exactly as UHD's fc32 conversion does, (float)v * (1 / 32767)
// (cpp/USRP_hardware_manager.cpp:764-820 asks UHD for fc32); the carried-over history is always fc32.
constexpr float kSc16ToFloat = 1.0f / 32767.0f;
template <bool SC16>
__device__ __forceinline__ float2 win_at_t(const Window& w, long long s) {
    if (!SC16) return win_at(w, s);
    if (s < w.n_hist) return w.hist[s];
    s -= w.n_hist;
    if (s < w.n_in) {
        const short2 v = reinterpret_cast<const short2*>(w.in)[s];
        return make_float2((float)v.x * kSc16ToFloat, (float)v.y * kSc16ToFloat);
    }
    return make_float2(0.f, 0.f);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -j (forward-transform quarter turn)
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }

// 4-point forward DFT in place, natural order out.
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mj(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

#define GSDR_C1 0.92387953251128674f  /* cos(pi/8) */
#define GSDR_S1 0.38268343236508977f  /* sin(pi/8) */
#define GSDR_R2 0.70710678118654752f  /* sqrt(1/2) */

// 16-point forward DFT in registers.  Input v[n] natural; output X[ka + 4 kb] lands in v[4 ka + kb].
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) fft4(v[nb], v[4 + nb], v[8 + nb], v[12 + nb]);
    // v[4 ka + nb] *= W16^(nb ka)
    v[5] = cmul(v[5], make_float2(GSDR_C1, -GSDR_S1));                            // W^1
    v[6] = make_float2((v[6].x + v[6].y) * GSDR_R2, (v[6].y - v[6].x) * GSDR_R2);   // W^2
    v[7] = cmul(v[7], make_float2(GSDR_S1, -GSDR_C1));                            // W^3
    v[9] = make_float2((v[9].x + v[9].y) * GSDR_R2, (v[9].y - v[9].x) * GSDR_R2);   // W^2
    v[10] = mul_mj(v[10]);                                                        // W^4
    v[11] = make_float2((v[11].y - v[11].x) * GSDR_R2, -(v[11].x + v[11].y) * GSDR_R2);  // W^6
    v[13] = cmul(v[13], make_float2(GSDR_S1, -GSDR_C1));                          // W^3
    v[14] = make_float2((v[14].y - v[14].x) * GSDR_R2, -(v[14].x + v[14].y) * GSDR_R2);  // W^6
    v[15] = cmul(v[15], make_float2(-GSDR_C1, GSDR_S1));                          // W^9
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) fft4(v[4 * ka], v[4 * ka + 1], v[4 * ka + 2], v[4 * ka + 3]);
}

// 8-point forward DFT in registers.  Input v[n] natural (n = 2 na + nb); output X[ka + 4 kb] in
// v[2 ka + kb].
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
    fft4(v[0], v[2], v[4], v[6]);
    fft4(v[1], v[3], v[5], v[7]);
    // odd branch twiddles W8^ka on v[2 ka + 1]
    v[3] = make_float2((v[3].x + v[3].y) * GSDR_R2, (v[3].y - v[3].x) * GSDR_R2);    // W8^1
    v[5] = mul_mj(v[5]);                                                           // W8^2
    v[7] = make_float2((v[7].y - v[7].x) * GSDR_R2, -(v[7].x + v[7].y) * GSDR_R2);   // W8^3
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        const float2 a = v[2 * ka], b = v[2 * ka + 1];
        v[2 * ka] = cadd(a, b);
        v[2 * ka + 1] = csub(a, b);
    }
}

__device__ __forceinline__ void team_barrier(int team, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float4 ldg_f4(const float2* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// --------------------------------------------------------------------------------------------
// fused kernel, N = 2048
// --------------------------------------------------------------------------------------------
constexpr int FN = 2048;        // channels
constexpr int FTEAMS = 4;       // frames in flight per CTA (one 128-thread team each)
constexpr int FTEAM_THREADS = 128;
constexpr int FTHREADS = FTEAMS * FTEAM_THREADS;
constexpr int FBUF = 2064;      // float2 per exchange buffer (16*129)
constexpr int FTW1 = 16 * 128;  // W_2048^(l k1) laid out [k1][l]
constexpr int FTW2 = 8 * 16;    // W_128^(n3 k2) laid out [n3][k2]

struct FusedSmem {
    float2 a[FTEAMS][FBUF];
    float2 b[FTEAMS][FBUF];
    float2 tw1[FTW1];
    float2 tw2[FTW2];
};

// Load the two float4 (4 samples) of window row `row` that thread t owns.  Uniform fast path when
// the row lies entirely in the `in` segment at a 16-byte aligned offset.
template <bool kFast>
__device__ __forceinline__ void load_row(const Window& w, long long row, int t, float4 (&dst)[2]) {
    const long long s0 = row * FN;
    if (kFast) {
        const float2* p = w.in + (s0 - w.n_hist);
        dst[0] = ldg_f4(p + 2 * t);
        dst[1] = ldg_f4(p + 2 * (t + FTHREADS));
    } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const long long s = s0 + 2 * (t + c * FTHREADS);
            const float2 e0 = win_at(w, s), e1 = win_at(w, s + 1);
            dst[c] = make_float4(e0.x, e0.y, e1.x, e1.y);
        }
    }
}

template <int P>
__global__ void __launch_bounds__(FTHREADS, 1)
pfb_fused_2048_kernel(const PfbJob single, const PfbJob* __restrict__ table, const int* __restrict__ tile_begin, int n_jobs,
                      int iters_per_tile, int total_tiles, const float2* __restrict__ tw_global) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
    const int t = threadIdx.x;
    const int team = t >> 7;
    const int tid = t & 127;

    for (int i = t; i < FTW1 + FTW2; i += FTHREADS) {
        if (i < FTW1) sm.tw1[i] = tw_global[i];
        else sm.tw2[i - FTW1] = tw_global[i];
    }

    int loaded_job = -1;
    float w[P][2][2];       // taps of this thread's 4 columns
    unsigned int bp[8];     // the (up to 16) selected bins this thread gathers, two per register

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        // ---- locate the job (stream) this tile belongs to --------------------------------------
        int j = 0;
        if (table != nullptr) {
            while (j + 1 < n_jobs && tile >= tile_begin[j + 1]) ++j;
        }
        const PfbJob job = (table != nullptr) ? table[j] : single;
        const int tile_in_job = tile - ((table != nullptr) ? tile_begin[j] : 0);
        if (j != loaded_job) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int u0 = tid + FTEAM_THREADS * (2 * jj), u1 = u0 + FTEAM_THREADS;
                const unsigned int b0 = u0 < job.T ? (job.bins ? (unsigned int)__ldg(job.bins + u0) : (unsigned int)u0) : 0u;
                const unsigned int b1 = u1 < job.T ? (job.bins ? (unsigned int)__ldg(job.bins + u1) : (unsigned int)u1) : 0u;
                bp[jj] = (b0 & 0xffffu) | (b1 << 16);
            }
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float2 tw = __ldg(reinterpret_cast<const float2*>(job.taps + i * FN + 2 * (t + c * FTHREADS)));
                    w[i][c][0] = tw.x;
                    w[i][c][1] = tw.y;
                }
            loaded_job = j;
        }
        const Window win = job.win;
        const int it0 = tile_in_job * iters_per_tile;
        const int job_iters = (job.n_frames + FTEAMS - 1) / FTEAMS;
        const int n_it = min(iters_per_tile, job_iters - it0);
        const long long frame0 = (long long)job.first_frame + (long long)it0 * FTEAMS;  // window frame index
        const long long last_frame = (long long)job.first_frame + job.n_frames;        // exclusive
        // rows whose 2048 samples all come from `in`, 16-byte aligned, and exist
        const bool aligned = ((win.n_hist & 1) == 0) && ((reinterpret_cast<uintptr_t>(win.in) & 15) == 0);
        const long long fast_lo = (win.n_hist + FN - 1) / FN;             // first row fully inside `in`
        const long long fast_hi = (win.n_hist + win.n_in) / FN;           // first row not fully present

        float4 r[FTEAMS + P - 1][2];
        // rows [row_lo, row_lo+FTEAMS) -> r[P-1 ..]; vector path when they lie inside `in`, aligned
        auto load_new_rows = [&](long long row_lo) {
            if (aligned && row_lo >= fast_lo && row_lo + FTEAMS <= fast_hi) {
#pragma unroll
                for (int q = 0; q < FTEAMS; ++q) load_row<true>(win, row_lo + q, t, r[P - 1 + q]);
            } else {
#pragma unroll
                for (int q = 0; q < FTEAMS; ++q) load_row<false>(win, row_lo + q, t, r[P - 1 + q]);
            }
        };
        // pull rows [row_lo, row_lo+FTEAMS) towards L2 (one request per 128-byte line)
        auto prefetch_rows = [&](long long row_lo) {
            if ((t & 7) != 0) return;
            const long long s = (row_lo * FN - win.n_hist) + 2 * (long long)t;
            if (s < 0) return;
#pragma unroll
            for (int q = 0; q < FTEAMS; ++q)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const long long sq = s + (long long)q * FN + 2 * c * FTHREADS;
                    if (sq + 16 <= win.n_in) asm volatile("prefetch.global.L2 [%0];" ::"l"(win.in + sq));
                }
        };
        // prologue: the P-1 rows older than the first new row, then iteration 0's new rows
#pragma unroll
        for (int i = 0; i < P - 1; ++i) {
            const long long row = frame0 + i;
            if (aligned && row >= fast_lo && row < fast_hi) load_row<true>(win, row, t, r[i]);
            else load_row<false>(win, row, t, r[i]);
        }
        load_new_rows(frame0 + P - 1);
        if (n_it > 1) prefetch_rows(frame0 + P - 1 + FTEAMS);

        for (int it = 0; it < n_it; ++it) {
            const long long fbase = frame0 + (long long)it * FTEAMS;
            // ---- polyphase FIR: frame (fbase+q) = sum_i row[q+i] * w_i, straight into team q's buffer
#pragma unroll
            for (int q = 0; q < FTEAMS; ++q) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const float4 x = r[q + i][c];
                        acc.x = fmaf(x.x, w[i][c][0], acc.x);
                        acc.y = fmaf(x.y, w[i][c][0], acc.y);
                        acc.z = fmaf(x.z, w[i][c][1], acc.z);
                        acc.w = fmaf(x.w, w[i][c][1], acc.w);
                    }
                    *reinterpret_cast<float4*>(&sm.a[q][2 * (t + c * FTHREADS)]) = acc;
                }
            }
            // slide the row window
#pragma unroll
            for (int i = 0; i < P - 1; ++i) {
                r[i][0] = r[i + FTEAMS][0];
                r[i][1] = r[i + FTEAMS][1];
            }
            __syncthreads();
            // ---- one 2048-point FFT per team ------------------------------------------------------
            const long long frame = fbase + team;
            if (frame < last_frame) {
                float2* A = sm.a[team];
                float2* B = sm.b[team];
                float2 v[16];
                {   // stage 1: radix-16 over n1 (stride 128); thread l = n2*8+n3
                    const int l = tid;
#pragma unroll
                    for (int n1 = 0; n1 < 16; ++n1) v[n1] = A[l + 128 * n1];
                    fft16(v);
#pragma unroll
                    for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                        for (int kb = 0; kb < 4; ++kb) {
                            const int k1 = ka + 4 * kb;
                            float2 x = v[4 * ka + kb];
                            if (k1 != 0) x = cmul(x, sm.tw1[k1 * 128 + l]);
                            B[k1 * 129 + l] = x;
                        }
                }
                team_barrier(team, FTEAM_THREADS);
                {   // stage 2: radix-16 over n2; thread = n3*16 + k1
                    const int n3 = tid >> 4, k1 = tid & 15;
#pragma unroll
                    for (int n2 = 0; n2 < 16; ++n2) v[n2] = B[k1 * 129 + n2 * 8 + n3];
                    fft16(v);
#pragma unroll
                    for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                        for (int kb = 0; kb < 4; ++kb) {
                            const int k2 = ka + 4 * kb;
                            float2 x = v[4 * ka + kb];
                            if (k2 != 0) x = cmul(x, sm.tw2[n3 * 16 + k2]);
                            A[n3 * 256 + k2 * 16 + k1] = x;
                        }
                }
                team_barrier(team, FTEAM_THREADS);
                // stage 3: radix-8 over n3, in place; g = k1 + 16 k2, output bin = g + 256 k3
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int g = tid + 128 * h;
                    float2 u[8];
#pragma unroll
                    for (int n3 = 0; n3 < 8; ++n3) u[n3] = A[n3 * 256 + g];
                    fft8(u);
#pragma unroll
                    for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) A[(ka + 4 * kb) * 256 + g] = u[2 * ka + kb];
                }
            }
            // The FFT registers are dead from here on: issue the next iteration's row loads now so
            // their latency hides under the barrier, the gather and the next FIR's first FMAs.
            if (it + 1 < n_it) {
                load_new_rows(fbase + FTEAMS + P - 1);
                if (it + 2 < n_it) prefetch_rows(fbase + 2 * FTEAMS + P - 1);
            }
            if (frame < last_frame) {
                float2* A = sm.a[team];
                team_barrier(team, FTEAM_THREADS);
                // tone selection: coalesced sample-major store, 8 independent gathers in flight
                float2* o = job.out + (frame - job.first_frame) * (long long)job.T;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h * 8 * FTEAM_THREADS < job.T) {
                        float2 val[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const int j2 = h * 8 + jj;
                            const unsigned int bin = (bp[j2 >> 1] >> (16 * (j2 & 1))) & 0xffffu;
                            val[jj] = A[bin];
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const int u = tid + FTEAM_THREADS * (h * 8 + jj);
                            if (u < job.T) o[u] = val[jj];
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// --------------------------------------------------------------------------------------------
// warp-specialised fused kernel, N = 2048 = 8 x 16 x 16: shared definitions
//
//   FRONT (warps 0-7, 256 threads): thread l owns the 8 polyphase columns l + 256 j: P-tap FIR, first FFT
//     stage (radix-8 over j), stage-1 twiddles W_2048^(l k1), hand-over through a ring of exchange buffers.
//   BACK (2 teams of 128 threads, alternate frames): stage 2 (radix-16, twiddles W_256^(n3 k2)), exchange,
//     stage 3 (radix-16).
// The producers run ahead through a 4-deep ring guarded by named barriers, so row loads, FIR, FFT passes and
// output stores of different frames overlap on the SM.
// --------------------------------------------------------------------------------------------
constexpr int WS_FRONT = 256;
constexpr int WS_TEAM = 128;
constexpr int WS_TEAMS = 2;
constexpr int WS_THREADS = WS_FRONT + WS_TEAMS * WS_TEAM;  // 512
constexpr int WS_TW1 = 8 * 256;     // W_2048^(l k1) laid out [k1][l]
constexpr int WS_TW2 = 16 * 16;     // W_256^(n3 k2) laid out [n3][k2]


__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
constexpr int WS_PC = WS_FRONT + WS_TEAM;        // participants of a full/empty barrier

// tile -> (job, frame range); identical in every warp role so the frame counters stay in step.
// Two forms: arithmetic (one stream, or the lock-step kernel's uniform tiles) and an explicit table built by the host for
// multi-stream launches (PfbTile: balanced over the CTAs, tiles never span two streams).
struct WsTile {
    int job;
    long long fa, fb;  // window frame indices [fa, fb)
};
__device__ __forceinline__ WsTile ws_locate(int tile, const PfbJob& single, const PfbJob* table, const int* tile_begin,
                                            int n_jobs, int frames_per_tile, PfbJob& job) {
    int j = 0;
    if (table != nullptr) {
        while (j + 1 < n_jobs && tile >= tile_begin[j + 1]) ++j;
        job = table[j];
    } else {
        job = single;
    }
    const int tile_in_job = tile - ((table != nullptr) ? tile_begin[j] : 0);
    WsTile r;
    r.job = j;
    r.fa = (long long)job.first_frame + (long long)tile_in_job * frames_per_tile;
    const long long end = (long long)job.first_frame + job.n_frames;
    r.fb = r.fa + frames_per_tile < end ? r.fa + frames_per_tile : end;
    return r;
}
__device__ __forceinline__ WsTile ws_locate_table(const PfbTile* tiles, int ti, const PfbJob* table, PfbJob& job) {
    const PfbTile tt = tiles[ti];
    job = table[tt.job];
    WsTile r;
    r.job = tt.job;
    r.fa = (long long)job.first_frame + tt.fa;
    r.fb = (long long)job.first_frame + tt.fb;
    return r;
}

// --------------------------------------------------------------------------------------------
// packed warp-specialised fused kernel, N = 2048 = 8 x 16 x 16   (the product default)
//
// Same role split as the kernel above, re-derived for the issue and barrier budget:
//   * all complex arithmetic runs on the packed fp32 pipe (FFMA2 / FADD2 / FMUL2, packed_f32x2.cuh):
//     the FIR is P packed FMAs per sample with the real tap broadcast, a radix-4 butterfly is 6 packed
//     + 4 scalar adds, a twiddle is 1 packed multiply + 2 FMAs;
//   * FRONT keeps all seven stage-1 twiddles W_2048^(l k1) in registers (no per-frame products);
//   * BACK maps one 256-point sub-transform (fixed k1) onto one HALF-WARP: lane = n3 for stage 2 and
//     lane = k2 for stage 3, so the stage-2 -> stage-3 exchange is a 16x16 transpose private to the
//     half-warp (padded tile, __syncwarp) instead of a team-wide buffer behind two named barriers;
//   * the finished spectrum is written once to a double-buffered [k1][k3][k2] tile, so a team needs a
//     single named barrier per frame (before the tone gather).
// --------------------------------------------------------------------------------------------
#ifndef GSDR_WP_D1
#define GSDR_WP_D1 4
#endif
#ifndef GSDR_WP_XB
#define GSDR_WP_XB 2
#endif
constexpr int WP_D1 = GSDR_WP_D1;  // exchange-1 ring depth
constexpr int WP_E1 = 8 * 258;   // [k1][n2*16+n3], k1 stride 258 float2
constexpr int WP_H = 16 * 17;    // one 16x16 transpose tile, row stride 17 float2
constexpr int WP_X = 8 * 256;    // one frame's spectrum: rows (k1, k3) of 16 bins, in-row order from PfbJob::xperm
constexpr int WP_THREADS = WS_FRONT + WS_TEAMS * WS_TEAM;  // 512
constexpr int WP_XB = GSDR_WP_XB;   // spectrum tiles per team (a team may finish two frames before a gather frees one)
#ifndef GSDR_WP_LA
#define GSDR_WP_LA 1
#endif
// 1: the head frames of a stream's window (those that read carried-over history rows) take the "mixed" span body instead
// of the general one.  Off: a third instantiation of the span body cost the single-stream kernel 4 - 6 % (register
// allocation of the producers' loop), more than the four head frames per stream are worth.
#ifndef GSDR_WP_MIXED_HEAD
#define GSDR_WP_MIXED_HEAD 0
#endif
#ifndef GSDR_WP_HOIST
#define GSDR_WP_HOIST 0
#endif
constexpr int WP_LA = GSDR_WP_LA;
constexpr bool WP_HOIST = GSDR_WP_HOIST != 0;
// Register split (setmaxnreg).  The pool is what the CTA got at launch, 512 threads x 128 registers; the
// split must not exceed it (an `inc` the pool cannot satisfy never returns).  ptxas rounds the launch
// allocation of a setmaxnreg kernel DOWN to a multiple of 32 registers per thread, so 512 threads is the
// only CTA shape that owns the whole register file.
#ifndef GSDR_WP_FRONT_REGS
#define GSDR_WP_FRONT_REGS 160
#endif
constexpr int WP_FRONT_REGS = GSDR_WP_FRONT_REGS;
constexpr int WP_BACK_REGS = 256 - WP_FRONT_REGS;
static_assert(WP_THREADS == 512 && WS_FRONT * WP_FRONT_REGS + WS_TEAMS * WS_TEAM * WP_BACK_REGS <= WP_THREADS * 128,
              "setmaxnreg split must fit the launch-time register pool");

struct WpSmem {
    float2 x[WS_TEAMS][WP_XB][WP_X];  // finished spectra, [team][tile]
    unsigned long long xfull[WS_TEAMS][WP_XB];   // mbarrier: the team's 4 warps have written the tile
    unsigned long long xempty[WS_TEAMS][WP_XB];  // mbarrier: the 8 producer warps have gathered from it
    float2 e1[WP_D1][WP_E1];
    float2 h[WS_TEAMS * (WS_TEAM / 32)][2][WP_H];
};
// named barriers of the exchange-1 ring; each has 384 participants (the 256 producers and one 128-thread team)
constexpr int WP_BAR_FULL = 1;                    // + slot : FRONT arrives, the frame's BACK team waits
constexpr int WP_BAR_EMPTY = 1 + WP_D1;           // + slot : BACK team arrives, FRONT waits

// mbarriers guard the spectrum tiles: unlike a named barrier they can be POLLED, so the producers gather a
// finished frame whenever one is ready instead of at a fixed distance behind production.
__device__ __forceinline__ void mbar_init(unsigned int addr, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned int addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// One arrival for the whole warp: 32 lanes arriving on one address would serialise into 32 shared-memory
// atomics.  __syncwarp orders the lanes' earlier shared-memory accesses before the elected lane's release.
// (Predicated inside the asm statement: a C-level `if (lane == 0)` is a branch, and ptxas then spills around it.)
__device__ __forceinline__ void mbar_arrive_warp(unsigned int addr) {
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(addr),
        "r"(threadIdx.x & 31u)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned int addr, unsigned int parity) {
    unsigned int ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned int addr, unsigned int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(addr),
        "r"(parity)
        : "memory");
}

template <typename F, int... I>
__device__ __forceinline__ void for_each_index(std::integer_sequence<int, I...>, F&& fn) {
    (fn(std::integral_constant<int, I>{}), ...);
}

// LA = input rows in flight per producer thread (register look-ahead, in frames); HOIST = keep all seven
// stage-1 twiddles in registers instead of three plus per-frame products.
// The round-1 form of the kernel below, kept verbatim for the single-stream fc32 launch from device memory (the headline
// path): the generalised kernel (tile table, sc16, mixed head rows) loses 4 - 6 % there to ptxas's register allocation of the
// producers' loop (362 against 385 GS/s on the same B200, A/B'd through GSDR_LIB_PATH), so that launch keeps this one.
// LA = input rows in flight per producer thread (register look-ahead, in frames); HOIST = keep all seven
// stage-1 twiddles in registers instead of three plus per-frame products.
template <int P, int LA, bool HOIST>
__global__ void __launch_bounds__(WP_THREADS, 1)
pfb_fused_wsp1_2048_kernel(const PfbJob single, const PfbJob* __restrict__ table, const int* __restrict__ tile_begin, int n_jobs,
                          int frames_per_tile, int total_tiles, const float2* __restrict__ tw_global) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WpSmem& sm = *reinterpret_cast<WpSmem*>(smem_raw);
    const int t = threadIdx.x;
    unsigned int f = 0;  // frames this CTA has started, counted identically by every role
    PfbJob job;
    const unsigned int xfull0 = (unsigned int)__cvta_generic_to_shared(&sm.xfull[0][0]);
    const unsigned int xempty0 = (unsigned int)__cvta_generic_to_shared(&sm.xempty[0][0]);
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < WS_TEAMS * WP_XB; ++i) {
            mbar_init(xfull0 + 8 * i, WS_TEAM / 32);   // one arrival per warp (see mbar_arrive_warp)
            mbar_init(xempty0 + 8 * i, WS_FRONT / 32);
        }
    }
    // carry-over of a single-stream launch (what move_buffer does in the reference, cpp/kernels.cu:444-470): the CTA with
    // the last, shorter tile copies the window's tail into the other history buffer -- input only, nothing here reads it
    if (n_jobs == 1 && single.tail_dst != nullptr && blockIdx.x == gridDim.x - 1) {
        const long long first = single.win.n_hist + single.win.n_in - single.tail_n;
        for (long long i = t; i < single.tail_n; i += WS_THREADS) single.tail_dst[i] = win_at(single.win, first + i);
    }
    __syncthreads();
    // frame number n (per CTA) -> team n & 1, that team's frame c = n >> 1, tile c % WP_XB, use c / WP_XB of the tile

    if (t < WS_FRONT) {
        // ======================================= FRONT ===========================================
        // Transposed-form FIR: each arriving input row updates the P frames it contributes to, so a thread
        // holds P accumulators (the finished one IS the FFT input) and only the rows still in flight.  The
        // per-frame FMA chain is the reference's (i = 0..P-1, cpp/kernels.cu:495-506), bit for bit.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WP_FRONT_REGS));
        const int l = t;
        constexpr int NTW = HOIST ? 8 : 5;
        float2 tw[NTW];  // W_2048^(l k1): k1 = 1..7, or k1 = 1, 2, 4 (slots 1, 2, 4)
        if (HOIST) {
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) tw[k1 % NTW] = __ldg(&tw_global[k1 * 256 + l]);
        } else {
            tw[1] = __ldg(&tw_global[1 * 256 + l]);
            tw[2] = __ldg(&tw_global[2 * 256 + l]);
            tw[4] = __ldg(&tw_global[4 * 256 + l]);
        }
        float w[P][8];
        constexpr int U = (P % LA == 0) ? P : P * LA;  // unroll period: accumulator and landing roles repeat
        c2 acc[P][8];
        c2 land[LA][8];
        // Tone selection (tone_select of the reference, cpp/kernels.cu:531-554) also lives here: the producers
        // have issue slots to spare, the FFT teams do not.  Thread l stores out[frame*T + l + 256 j], j < 8.
        constexpr int NU = FN / WS_FRONT;
        unsigned int bp[NU / 2];  // byte offsets (inside a spectrum tile) of the bins this thread gathers, two per register
        int nv = 0;               // how many of the NU output slots exist (u < T)
        const unsigned int x_base = (unsigned int)__cvta_generic_to_shared(&sm.x[0][0][0]);
        int loaded_job = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const WsTile tl = ws_locate(tile, single, table, tile_begin, n_jobs, frames_per_tile, job);
            if (tl.job != loaded_job) {
#pragma unroll
                for (int i = 0; i < P; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[i][j] = __ldg(job.taps + i * FN + l + 256 * j);
#pragma unroll
                for (int jj = 0; jj < NU / 2; ++jj) {
                    unsigned int pk = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int u = l + WS_FRONT * (2 * jj + h);
                        unsigned int bin = 0;
                        if (u < job.T) bin = job.bins ? (unsigned int)__ldg(job.bins + u) : (unsigned int)u;
                        bin &= (FN - 1);
                        const unsigned int pos = job.xperm ? (unsigned int)__ldg(job.xperm + bin) : ((bin >> 3) & 15u);
                        const unsigned int idx = (bin & 7u) * 256u + (bin >> 7) * 16u + (pos & 15u);
                        pk |= (idx * 8u) << (16 * h);
                    }
                    bp[jj] = pk;
                }
                nv = (job.T - l + WS_FRONT - 1) / WS_FRONT;
                nv = nv < 0 ? 0 : (nv > NU ? NU : nv);
                loaded_job = tl.job;
            }
            const unsigned int f_tile0 = f;  // CTA frame counter of the tile's first frame
            unsigned int g = f;              // next frame whose tones are still to be stored
            // Gather the selected bins of frame n (CTA numbering) from its team's spectrum tile and store them
            // sample-major (coalesced 8-byte stores).  `block` = wait for the tile; otherwise poll once and
            // return false when the team has not finished the frame yet.
            auto gather = [&](const unsigned int n, const bool block) -> bool {
                const unsigned int qq = n & 1u, cc = n >> 1;
                const unsigned int tile_i = qq * WP_XB + (cc % WP_XB), par = (cc / WP_XB) & 1u;
                if (block) {
                    mbar_wait(xfull0 + 8 * tile_i, par);
                } else {
                    // warp-uniform decision (lanes can observe the phase flip at different times)
                    if (!__all_sync(0xffffffffu, mbar_test(xfull0 + 8 * tile_i, par))) return false;
                }
                const unsigned int xa = x_base + tile_i * (unsigned int)(WP_X * sizeof(float2));
                c2* o = reinterpret_cast<c2*>(job.out) + ((tl.fa - job.first_frame) + (long long)(n - f_tile0)) * (long long)job.T + l;
#pragma unroll
                for (int h = 0; h < NU / 4; ++h) {
                    if (h * 4 < nv) {
                        c2 val[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j2 = h * 4 + jj;
                            const unsigned int off = (j2 & 1) ? (bp[j2 >> 1] >> 16) : (bp[j2 >> 1] & 0xffffu);
                            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(val[jj]) : "r"(xa + off) : "memory");
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (h * 4 + jj < nv) o[WS_FRONT * (h * 4 + jj)] = val[jj];
                    }
                }
                mbar_arrive_warp(xempty0 + 8 * tile_i);  // after the stores: they have consumed the gathered values
                return true;
            };
            const Window win = job.win;
            const long long fast_lo = (win.n_hist + FN - 1) / FN;    // first row fully inside `in`
            const long long fast_hi = (win.n_hist + win.n_in) / FN;  // first row not fully present
            // The span body is instantiated twice: kFast when every row it touches lies fully inside the
            // `in` segment (plain coalesced 8-byte loads, no per-row range logic), general otherwise.
            auto run_tile = [&](auto fast_tag, const long long fa, const long long fb) {
                constexpr bool kFast = decltype(fast_tag)::value;
                const long long last_row = fb + P - 1;      // rows this span needs: [fa, last_row)
                const long long n_steps = last_row - fa;    // one step per input row
                auto load_row8 = [&](long long row, c2 (&dst)[8]) {
                    if (kFast) {
                        // volatile: the loads stay where the pipeline puts them (after the FIR that frees their
                        // landing registers) instead of being hoisted into extra registers by the scheduler
                        const c2* p = reinterpret_cast<const c2*>(win.in + (row * FN - win.n_hist) + l);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(dst[j]) : "l"(p + 256 * j) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) dst[j] = c2_from(win_at(win, row * FN + l + 256 * j));
                    }
                };
                // warp 0 pulls a whole row (128 lines of 128 bytes) towards L2 several frames ahead
                auto prefetch_row = [&](long long row) {
                    if (kFast && l < 32 && row < last_row) {
                        const float2* p = win.in + (row * FN - win.n_hist) + 16 * l;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 512 * jj));
                    }
                };
#pragma unroll
                for (int i = 0; i < LA; ++i)
                    if (fa + i < last_row) load_row8(fa + i, land[i]);
#pragma unroll
                for (int i = LA; i < LA + 4; ++i) prefetch_row(fa + i);

                // one step: row (fa + s) arrives, frame (fa + s - P + 1) completes
                auto step = [&](const long long s, auto u_tag, const bool guarded, const bool emit) {
                    constexpr int u = decltype(u_tag)::value;
                    c2(&x)[8] = land[u % LA];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
#if GSDR_PK_MODE & 8
                        float xr, xi;
                        c2_unpack(x[j], xr, xi);
#pragma unroll
                        for (int i = P - 1; i >= 1; --i) {
                            float ar, ai;
                            c2_unpack(acc[(u + P * U - i) % P][j], ar, ai);
                            acc[(u + P * U - i) % P][j] = c2_pack(fmaf(xr, w[i][j], ar), fmaf(xi, w[i][j], ai));
                        }
                        acc[u % P][j] = c2_pack(xr * w[0][j], xi * w[0][j]);
#else
#pragma unroll
                        for (int i = P - 1; i >= 1; --i) acc[(u + P * U - i) % P][j] = c2_fma_s(x[j], w[i][j], acc[(u + P * U - i) % P][j]);
                        acc[u % P][j] = c2_scale(x[j], w[0][j]);
#endif
                    }
                    if (!guarded || s + LA < n_steps) load_row8(fa + s + LA, x);
                    prefetch_row(fa + s + LA + 4);
                    if (!emit) return;
                    // ---- FFT stage 1: radix-8 over j, then twiddle; X[ka + 4 kb] sits in z[2 ka + kb]
                    c2 z[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] = acc[(u + 1) % P][j];
                    c2_fft8(z);
                    const int slot = f % WP_D1;
                    if (f >= WP_D1) bar_sync(WP_BAR_EMPTY + slot, WS_PC);
                    float2* E = sm.e1[slot] + l;
                    sts_c2(E + 0 * 258, z[0]);
                    if (HOIST) {
                        sts_c2(E + 1 * 258, c2_cmul(z[2], tw[1 % NTW].x, tw[1 % NTW].y));
                        sts_c2(E + 2 * 258, c2_cmul(z[4], tw[2 % NTW].x, tw[2 % NTW].y));
                        sts_c2(E + 3 * 258, c2_cmul(z[6], tw[3 % NTW].x, tw[3 % NTW].y));
                        sts_c2(E + 4 * 258, c2_cmul(z[1], tw[4 % NTW].x, tw[4 % NTW].y));
                        sts_c2(E + 5 * 258, c2_cmul(z[3], tw[5 % NTW].x, tw[5 % NTW].y));
                        sts_c2(E + 6 * 258, c2_cmul(z[5], tw[6 % NTW].x, tw[6 % NTW].y));
                        sts_c2(E + 7 * 258, c2_cmul(z[7], tw[7 % NTW].x, tw[7 % NTW].y));
                    } else {
                        const float2 t1 = tw[1], t2 = tw[2], t4 = tw[4];
                        const float2 t3 = cmul(t1, t2), t5 = cmul(t1, t4), t6 = cmul(t2, t4);
                        const float2 t7 = cmul(t3, t4);
                        sts_c2(E + 1 * 258, c2_cmul(z[2], t1.x, t1.y));
                        sts_c2(E + 2 * 258, c2_cmul(z[4], t2.x, t2.y));
                        sts_c2(E + 3 * 258, c2_cmul(z[6], t3.x, t3.y));
                        sts_c2(E + 4 * 258, c2_cmul(z[1], t4.x, t4.y));
                        sts_c2(E + 5 * 258, c2_cmul(z[3], t5.x, t5.y));
                        sts_c2(E + 6 * 258, c2_cmul(z[5], t6.x, t6.y));
                        sts_c2(E + 7 * 258, c2_cmul(z[7], t7.x, t7.y));
                    }
                    bar_arrive(WP_BAR_FULL + slot, WS_PC);
                    ++f;
                    // One poll per produced frame keeps the stores a frame or two behind the teams.  (Polling once
                    // per unrolled group instead -- smaller hot code -- measured 8 % slower: the stores then come in
                    // bursts and the teams wait for their tiles.)
                    if (g + 1 < f && gather(g, false)) ++g;
                };
                auto guarded_group = [&](const long long s0) {
                    for_each_index(std::make_integer_sequence<int, U>{}, [&](auto u_tag) {
                        const long long ss = s0 + decltype(u_tag)::value;
                        if (ss < n_steps) step(ss, u_tag, true, ss >= P - 1);
                    });
                };
                // head: the first P-1 rows only prime the accumulators
                guarded_group(0);
                long long s = U;
                // steady state: whole groups whose frames all emit and whose look-ahead rows all exist
                for (; s + U - 1 + LA < n_steps; s += U)
                    for_each_index(std::make_integer_sequence<int, U>{}, [&](auto u_tag) {
                        step(s + decltype(u_tag)::value, u_tag, false, true);
                    });
                // tail
                for (; s < n_steps; s += U) guarded_group(s);
            };
            // frames whose P rows all lie inside `in` take the fast body; the few that touch the carried-over
            // history (head of a window) or its ragged end take the general one
            long long f0 = tl.fa > fast_lo ? tl.fa : fast_lo;
            long long f1 = tl.fb < fast_hi - P + 1 ? tl.fb : fast_hi - P + 1;
            if (f0 > tl.fb) f0 = tl.fb;
            if (f1 < f0) f1 = f0;
            if (tl.fa < f0) run_tile(std::false_type{}, tl.fa, f0);
            if (f0 < f1) run_tile(std::true_type{}, f0, f1);
            if (f1 < tl.fb) run_tile(std::false_type{}, f1, tl.fb);
            // drain: what the polls have not stored yet
            for (; g < f; ++g) gather(g, true);
        }
    } else {
        // ======================================== BACK ============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WP_BACK_REGS));
        const int q = (t - WS_FRONT) / WS_TEAM;       // team: frames with (f & 1) == q
        const int tid = (t - WS_FRONT) % WS_TEAM;
        const int lane16 = tid & 15;                  // n3 in stage 2, k2 in stage 3
        const int k1 = tid >> 4;                      // the 256-point sub-transform of this half-warp
        float2 tw2[16];  // W_256^(n3 k2)
#pragma unroll
        for (int k2 = 1; k2 < 16; ++k2) tw2[k2] = __ldg(&tw_global[WS_TW1 + lane16 * 16 + k2]);
        float2* H = sm.h[(t - WS_FRONT) >> 5][(tid >> 4) & 1];
        unsigned int c = 0;  // frames this team has finished
        const unsigned int xw0 = (unsigned int)__cvta_generic_to_shared(&sm.x[q][0][0]) + (unsigned int)(k1 * 256 * sizeof(float2));
        unsigned int xo[4];  // byte offset inside row (k1, k3) of this thread's bin k1 + 8 k2 + 128 k3, one byte per k3
        int loaded_job = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const WsTile tl = ws_locate(tile, single, table, tile_begin, n_jobs, frames_per_tile, job);
            if (tl.job != loaded_job) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned int pk = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int bin = k1 + 8 * lane16 + 128 * (4 * i + b);
                        const unsigned int pos = job.xperm ? (unsigned int)__ldg(job.xperm + bin) : (unsigned int)lane16;
                        pk |= ((pos & 15u) * 8u) << (8 * b);
                    }
                    xo[i] = pk;
                }
                loaded_job = tl.job;
            }
            const int nf = (int)(tl.fb - tl.fa);
            for (int i = (int)((q - f) & 1u); i < nf; i += 2) {
                const int slot = (f + i) % WP_D1;
                c2 v[16];
                // ---- stage 2: radix-16 over n2; lane = n3
                bar_sync(WP_BAR_FULL + slot, WS_PC);
                {
                    const float2* E = sm.e1[slot] + k1 * 258 + lane16;
#pragma unroll
                    for (int n2 = 0; n2 < 16; ++n2) v[n2] = lds_c2(E + n2 * 16);
                }
                c2_fft16_first(v);  // consumes every loaded value: the slot can go back to the producers
                bar_arrive(WP_BAR_EMPTY + slot, WS_PC);
                c2_fft16_second(v);
#pragma unroll
                for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const int k2 = ka + 4 * kb;
                        c2 x = v[4 * ka + kb];
                        if (k2 != 0) x = c2_cmul(x, tw2[k2].x, tw2[k2].y);
                        sts_c2(H + k2 * 17 + lane16, x);
                    }
                __syncwarp();
                // ---- stage 3: radix-16 over n3; lane = k2; bin = k1 + 8 k2 + 128 k3
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = lds_c2(H + lane16 * 17 + m);
                __syncwarp();
                c2_fft16(v);
                const unsigned int tile_i = q * WP_XB + (c % WP_XB), use = c / WP_XB;
                if (use > 0) mbar_wait(xempty0 + 8 * tile_i, (use - 1) & 1u);  // the tile's previous frame has been gathered
                const unsigned int xw = xw0 + (c % WP_XB) * (unsigned int)(WP_X * sizeof(float2));
#pragma unroll
                for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                    {
                        const int k3 = ka + 4 * kb;
                        const unsigned int off = __byte_perm(xo[k3 >> 2], 0u, 0x4440u + (k3 & 3));
                        asm volatile("st.shared.b64 [%0], %1;" ::"r"(xw + off + (unsigned int)(k3 * 16 * sizeof(float2))), "l"(v[4 * ka + kb])
                                     : "memory");
                    }
                mbar_arrive_warp(xfull0 + 8 * tile_i);
                ++c;
            }
            f += nf;
        }
    }
}

// TABLE: multi-stream launch driven by the host-built tile list (the single-stream arguments are unused and vice versa: two
// instantiations, so that neither carries the other's live registers through the producers' loop)
template <int P, int LA, bool HOIST, bool SC16, bool TABLE>
__global__ void __launch_bounds__(WP_THREADS, 1)
pfb_fused_wsp_2048_kernel(const PfbJob single, const PfbJob* __restrict__ table, const int* __restrict__ tile_begin, int n_jobs,
                          int frames_per_tile, int total_tiles, const float2* __restrict__ tw_global,
                          const PfbTile* __restrict__ tiles, const int* __restrict__ cta_begin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WpSmem& sm = *reinterpret_cast<WpSmem*>(smem_raw);
    const int t = threadIdx.x;
    unsigned int f = 0;  // frames this CTA has started, counted identically by every role
    PfbJob job;
    const unsigned int xfull0 = (unsigned int)__cvta_generic_to_shared(&sm.xfull[0][0]);
    const unsigned int xempty0 = (unsigned int)__cvta_generic_to_shared(&sm.xempty[0][0]);
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < WS_TEAMS * WP_XB; ++i) {
            mbar_init(xfull0 + 8 * i, WS_TEAM / 32);   // one arrival per warp (see mbar_arrive_warp)
            mbar_init(xempty0 + 8 * i, WS_FRONT / 32);
        }
    }
    // This CTA's tiles: [t_lo, t_hi) step t_step.  Table form (multi-stream): a contiguous run of the host-built tile list;
    // arithmetic form: tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    const int t_lo = TABLE ? cta_begin[blockIdx.x] : (int)blockIdx.x;
    const int t_hi = TABLE ? cta_begin[blockIdx.x + 1] : total_tiles;
    const int t_step = TABLE ? 1 : (int)gridDim.x;
    // carry-over (what move_buffer does in the reference, cpp/kernels.cu:444-470): the CTA that owns a stream's last tile
    // copies the window's tail into the other history buffer -- input only, nothing in this launch reads it
    if (TABLE) {
        for (int ti = t_lo; ti < t_hi; ++ti) {
            const PfbTile tt = tiles[ti];
            if (tt.flags & 1) {
                const PfbJob jb = table[tt.job];
                if (jb.tail_dst != nullptr) {
                    const long long first = jb.win.n_hist + jb.win.n_in - jb.tail_n;
                    for (long long i = t; i < jb.tail_n; i += WS_THREADS) jb.tail_dst[i] = win_at_t<SC16>(jb.win, first + i);
                }
            }
        }
    } else if (n_jobs == 1 && single.tail_dst != nullptr && blockIdx.x == gridDim.x - 1) {
        const long long first = single.win.n_hist + single.win.n_in - single.tail_n;
        for (long long i = t; i < single.tail_n; i += WS_THREADS) single.tail_dst[i] = win_at_t<SC16>(single.win, first + i);
    }
    __syncthreads();
    // frame number n (per CTA) -> team n & 1, that team's frame c = n >> 1, tile c % WP_XB, use c / WP_XB of the tile

    if (t < WS_FRONT) {
        // ======================================= FRONT ===========================================
        // Transposed-form FIR: each arriving input row updates the P frames it contributes to, so a thread
        // holds P accumulators (the finished one IS the FFT input) and only the rows still in flight.  The
        // per-frame FMA chain is the reference's (i = 0..P-1, cpp/kernels.cu:495-506), bit for bit.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WP_FRONT_REGS));
        const int l = t;
        constexpr int NTW = HOIST ? 8 : 5;
        float2 tw[NTW];  // W_2048^(l k1): k1 = 1..7, or k1 = 1, 2, 4 (slots 1, 2, 4)
        if (HOIST) {
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) tw[k1 % NTW] = __ldg(&tw_global[k1 * 256 + l]);
        } else {
            tw[1] = __ldg(&tw_global[1 * 256 + l]);
            tw[2] = __ldg(&tw_global[2 * 256 + l]);
            tw[4] = __ldg(&tw_global[4 * 256 + l]);
        }
        float w[P][8];
        constexpr int U = (P % LA == 0) ? P : P * LA;  // unroll period: accumulator and landing roles repeat
        c2 acc[P][8];
        c2 land[LA][8];
        // Tone selection (tone_select of the reference, cpp/kernels.cu:531-554) also lives here: the producers
        // have issue slots to spare, the FFT teams do not.  Thread l stores out[frame*T + l + 256 j], j < 8.
        constexpr int NU = FN / WS_FRONT;
        unsigned int bp[NU / 2];  // byte offsets (inside a spectrum tile) of the bins this thread gathers, two per register
        int nv = 0;               // how many of the NU output slots exist (u < T)
        const unsigned int x_base = (unsigned int)__cvta_generic_to_shared(&sm.x[0][0][0]);
        int loaded_job = -1;
        for (int tile = t_lo; tile < t_hi; tile += t_step) {
            const WsTile tl = TABLE ? ws_locate_table(tiles, tile, table, job)
                                    : ws_locate(tile, single, nullptr, nullptr, 1, frames_per_tile, job);
            if (tl.job != loaded_job) {
#pragma unroll
                for (int i = 0; i < P; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[i][j] = __ldg(job.taps + i * FN + l + 256 * j);
#pragma unroll
                for (int jj = 0; jj < NU / 2; ++jj) {
                    unsigned int pk = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int u = l + WS_FRONT * (2 * jj + h);
                        unsigned int bin = 0;
                        if (u < job.T) bin = job.bins ? (unsigned int)__ldg(job.bins + u) : (unsigned int)u;
                        bin &= (FN - 1);
                        const unsigned int pos = job.xperm ? (unsigned int)__ldg(job.xperm + bin) : ((bin >> 3) & 15u);
                        const unsigned int idx = (bin & 7u) * 256u + (bin >> 7) * 16u + (pos & 15u);
                        pk |= (idx * 8u) << (16 * h);
                    }
                    bp[jj] = pk;
                }
                nv = (job.T - l + WS_FRONT - 1) / WS_FRONT;
                nv = nv < 0 ? 0 : (nv > NU ? NU : nv);
                loaded_job = tl.job;
            }
            const unsigned int f_tile0 = f;  // CTA frame counter of the tile's first frame
            unsigned int g = f;              // next frame whose tones are still to be stored
            // Gather the selected bins of frame n (CTA numbering) from its team's spectrum tile and store them
            // sample-major (coalesced 8-byte stores).  `block` = wait for the tile; otherwise poll once and
            // return false when the team has not finished the frame yet.
            auto gather = [&](const unsigned int n, const bool block) -> bool {
                const unsigned int qq = n & 1u, cc = n >> 1;
                const unsigned int tile_i = qq * WP_XB + (cc % WP_XB), par = (cc / WP_XB) & 1u;
                if (block) {
                    mbar_wait(xfull0 + 8 * tile_i, par);
                } else {
                    // warp-uniform decision (lanes can observe the phase flip at different times)
                    if (!__all_sync(0xffffffffu, mbar_test(xfull0 + 8 * tile_i, par))) return false;
                }
                const unsigned int xa = x_base + tile_i * (unsigned int)(WP_X * sizeof(float2));
                c2* o = reinterpret_cast<c2*>(job.out) + ((tl.fa - job.first_frame) + (long long)(n - f_tile0)) * (long long)job.T + l;
#pragma unroll
                for (int h = 0; h < NU / 4; ++h) {
                    if (h * 4 < nv) {
                        c2 val[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j2 = h * 4 + jj;
                            const unsigned int off = (j2 & 1) ? (bp[j2 >> 1] >> 16) : (bp[j2 >> 1] & 0xffffu);
                            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(val[jj]) : "r"(xa + off) : "memory");
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (h * 4 + jj < nv) o[WS_FRONT * (h * 4 + jj)] = val[jj];
                    }
                }
                mbar_arrive_warp(xempty0 + 8 * tile_i);  // after the stores: they have consumed the gathered values
                return true;
            };
            const Window win = job.win;
            const long long fast_lo = (win.n_hist + FN - 1) / FN;    // first row fully inside `in`
            const long long fast_hi = (win.n_hist + win.n_in) / FN;  // first row not fully present
            const long long hist_full = win.n_hist / FN;             // rows [0, hist_full) lie fully inside the carried-over history
            // The span body is instantiated three times: kMode 1 when every row it touches lies fully inside the `in` segment
            // (plain coalesced 8-byte loads, no per-row range logic); kMode 2 for the tile at the head of a stream's window,
            // whose first rows are carried-over history -- a CTA-uniform choice of the base pointer per row, only the one
            // row that straddles history and new samples is read element by element (a multi-stream launch has such a tile
            // per stream: splitting it into a general and a fast span restarted the FIR pipeline twice per stream);
            // kMode 0 otherwise (ragged window end).
            auto run_tile = [&](auto mode_tag, const long long fa, const long long fb) {
                constexpr int kMode = decltype(mode_tag)::value;
                constexpr bool kFast = kMode == 1;
                const long long last_row = fb + P - 1;      // rows this span needs: [fa, last_row)
                const long long n_steps = last_row - fa;    // one step per input row
                auto load_row8 = [&](long long row, c2 (&dst)[8]) {
                    if (kFast || (kMode == 2 && (row >= fast_lo || row < hist_full))) {
                        // volatile: the loads stay where the pipeline puts them (after the FIR that frees their
                        // landing registers) instead of being hoisted into extra registers by the scheduler
                        const bool from_in = kFast || row >= fast_lo;
                        if (SC16 && from_in) {   // wire format: 4 bytes per sample, kept raw in the landing register until the FIR uses it
                            const unsigned int* p = reinterpret_cast<const unsigned int*>(win.in) + (row * FN - win.n_hist) + l;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                unsigned int raw;
                                asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(raw) : "l"(p + 256 * j) : "memory");
                                dst[j] = (c2)raw;   // converted where the FIR consumes it (step knows which rows are raw)
                            }
                        } else {
                            const float2* base = from_in ? win.in + (row * FN - win.n_hist) : win.hist + row * FN;
                            const c2* p = reinterpret_cast<const c2*>(base + l);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(dst[j]) : "l"(p + 256 * j) : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) dst[j] = c2_from(win_at_t<SC16>(win, row * FN + l + 256 * j));
                    }
                };
                // warp 0 pulls a whole row (128 lines of 128 bytes) towards L2 several frames ahead
                auto prefetch_row = [&](long long row) {
                    if ((kFast || (kMode == 2 && row >= fast_lo)) && l < 32 && row < last_row) {
                        if (SC16) {   // 8 KB per row: 64 lines
                            const short2* p = reinterpret_cast<const short2*>(win.in) + (row * FN - win.n_hist) + 32 * l;
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 1024 * jj));
                        } else {
                            const float2* p = win.in + (row * FN - win.n_hist) + 16 * l;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 512 * jj));
                        }
                    }
                };
#pragma unroll
                for (int i = 0; i < LA; ++i)
                    if (fa + i < last_row) load_row8(fa + i, land[i]);
#pragma unroll
                for (int i = LA; i < LA + 4; ++i) prefetch_row(fa + i);

                // one step: row (fa + s) arrives, frame (fa + s - P + 1) completes
                auto step = [&](const long long s, auto u_tag, const bool guarded, const bool emit) {
                    constexpr int u = decltype(u_tag)::value;
                    c2(&x)[8] = land[u % LA];
                    if (SC16 && kMode != 0 && (kFast || fa + s >= fast_lo)) {   // this row came in as raw int16 pairs
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const unsigned int raw = (unsigned int)x[j];
                            x[j] = c2_pack((float)(short)(raw & 0xffffu) * kSc16ToFloat, (float)(short)(raw >> 16) * kSc16ToFloat);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
#if GSDR_PK_MODE & 8
                        float xr, xi;
                        c2_unpack(x[j], xr, xi);
#pragma unroll
                        for (int i = P - 1; i >= 1; --i) {
                            float ar, ai;
                            c2_unpack(acc[(u + P * U - i) % P][j], ar, ai);
                            acc[(u + P * U - i) % P][j] = c2_pack(fmaf(xr, w[i][j], ar), fmaf(xi, w[i][j], ai));
                        }
                        acc[u % P][j] = c2_pack(xr * w[0][j], xi * w[0][j]);
#else
#pragma unroll
                        for (int i = P - 1; i >= 1; --i) acc[(u + P * U - i) % P][j] = c2_fma_s(x[j], w[i][j], acc[(u + P * U - i) % P][j]);
                        acc[u % P][j] = c2_scale(x[j], w[0][j]);
#endif
                    }
                    if (!guarded || s + LA < n_steps) load_row8(fa + s + LA, x);
                    prefetch_row(fa + s + LA + 4);
                    if (!emit) return;
                    // ---- FFT stage 1: radix-8 over j, then twiddle; X[ka + 4 kb] sits in z[2 ka + kb]
                    c2 z[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] = acc[(u + 1) % P][j];
                    c2_fft8(z);
                    const int slot = f % WP_D1;
                    if (f >= WP_D1) bar_sync(WP_BAR_EMPTY + slot, WS_PC);
                    float2* E = sm.e1[slot] + l;
                    sts_c2(E + 0 * 258, z[0]);
                    if (HOIST) {
                        sts_c2(E + 1 * 258, c2_cmul(z[2], tw[1 % NTW].x, tw[1 % NTW].y));
                        sts_c2(E + 2 * 258, c2_cmul(z[4], tw[2 % NTW].x, tw[2 % NTW].y));
                        sts_c2(E + 3 * 258, c2_cmul(z[6], tw[3 % NTW].x, tw[3 % NTW].y));
                        sts_c2(E + 4 * 258, c2_cmul(z[1], tw[4 % NTW].x, tw[4 % NTW].y));
                        sts_c2(E + 5 * 258, c2_cmul(z[3], tw[5 % NTW].x, tw[5 % NTW].y));
                        sts_c2(E + 6 * 258, c2_cmul(z[5], tw[6 % NTW].x, tw[6 % NTW].y));
                        sts_c2(E + 7 * 258, c2_cmul(z[7], tw[7 % NTW].x, tw[7 % NTW].y));
                    } else {
                        const float2 t1 = tw[1], t2 = tw[2], t4 = tw[4];
                        const float2 t3 = cmul(t1, t2), t5 = cmul(t1, t4), t6 = cmul(t2, t4);
                        const float2 t7 = cmul(t3, t4);
                        sts_c2(E + 1 * 258, c2_cmul(z[2], t1.x, t1.y));
                        sts_c2(E + 2 * 258, c2_cmul(z[4], t2.x, t2.y));
                        sts_c2(E + 3 * 258, c2_cmul(z[6], t3.x, t3.y));
                        sts_c2(E + 4 * 258, c2_cmul(z[1], t4.x, t4.y));
                        sts_c2(E + 5 * 258, c2_cmul(z[3], t5.x, t5.y));
                        sts_c2(E + 6 * 258, c2_cmul(z[5], t6.x, t6.y));
                        sts_c2(E + 7 * 258, c2_cmul(z[7], t7.x, t7.y));
                    }
                    bar_arrive(WP_BAR_FULL + slot, WS_PC);
                    ++f;
                    // One poll per produced frame keeps the stores a frame or two behind the teams.  (Polling once
                    // per unrolled group instead -- smaller hot code -- measured 8 % slower: the stores then come in
                    // bursts and the teams wait for their tiles.)
                    if (g + 1 < f && gather(g, false)) ++g;
                };
                auto guarded_group = [&](const long long s0) {
                    for_each_index(std::make_integer_sequence<int, U>{}, [&](auto u_tag) {
                        const long long ss = s0 + decltype(u_tag)::value;
                        if (ss < n_steps) step(ss, u_tag, true, ss >= P - 1);
                    });
                };
                // head: the first P-1 rows only prime the accumulators
                guarded_group(0);
                long long s = U;
                // steady state: whole groups whose frames all emit and whose look-ahead rows all exist
                for (; s + U - 1 + LA < n_steps; s += U)
                    for_each_index(std::make_integer_sequence<int, U>{}, [&](auto u_tag) {
                        step(s + decltype(u_tag)::value, u_tag, false, true);
                    });
                // tail
                for (; s < n_steps; s += U) guarded_group(s);
            };
            // frames whose P rows all lie inside `in` take the fast body; the few at the head of a stream's window that still
            // read carried-over history rows take the mixed one (its per-row pointer choice costs the producers ~15 % per
            // frame, measured on a whole tile: kept to the frames that need it); frames past the last complete row (never
            // produced by buffer_helper, kept for safety) the general one
            long long f1 = tl.fb < fast_hi - P + 1 ? tl.fb : fast_hi - P + 1;
            if (f1 < tl.fa) f1 = tl.fa;
            long long f0 = tl.fa > fast_lo ? tl.fa : fast_lo;   // first frame whose rows all lie in `in`
            if (f0 > f1) f0 = f1;
#if GSDR_WP_MIXED_HEAD
            if (tl.fa < f0) run_tile(std::integral_constant<int, 2>{}, tl.fa, f0);
#else
            if (tl.fa < f0) run_tile(std::integral_constant<int, 0>{}, tl.fa, f0);
#endif
            if (f0 < f1) run_tile(std::integral_constant<int, 1>{}, f0, f1);
            if (f1 < tl.fb) run_tile(std::integral_constant<int, 0>{}, f1, tl.fb);
            // drain: what the polls have not stored yet
            for (; g < f; ++g) gather(g, true);
        }
    } else {
        // ======================================== BACK ============================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WP_BACK_REGS));
        const int q = (t - WS_FRONT) / WS_TEAM;       // team: frames with (f & 1) == q
        const int tid = (t - WS_FRONT) % WS_TEAM;
        const int lane16 = tid & 15;                  // n3 in stage 2, k2 in stage 3
        const int k1 = tid >> 4;                      // the 256-point sub-transform of this half-warp
        float2 tw2[16];  // W_256^(n3 k2)
#pragma unroll
        for (int k2 = 1; k2 < 16; ++k2) tw2[k2] = __ldg(&tw_global[WS_TW1 + lane16 * 16 + k2]);
        float2* H = sm.h[(t - WS_FRONT) >> 5][(tid >> 4) & 1];
        unsigned int c = 0;  // frames this team has finished
        const unsigned int xw0 = (unsigned int)__cvta_generic_to_shared(&sm.x[q][0][0]) + (unsigned int)(k1 * 256 * sizeof(float2));
        unsigned int xo[4];  // byte offset inside row (k1, k3) of this thread's bin k1 + 8 k2 + 128 k3, one byte per k3
        int loaded_job = -1;
        for (int tile = t_lo; tile < t_hi; tile += t_step) {
            const WsTile tl = TABLE ? ws_locate_table(tiles, tile, table, job)
                                    : ws_locate(tile, single, nullptr, nullptr, 1, frames_per_tile, job);
            if (tl.job != loaded_job) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned int pk = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int bin = k1 + 8 * lane16 + 128 * (4 * i + b);
                        const unsigned int pos = job.xperm ? (unsigned int)__ldg(job.xperm + bin) : (unsigned int)lane16;
                        pk |= ((pos & 15u) * 8u) << (8 * b);
                    }
                    xo[i] = pk;
                }
                loaded_job = tl.job;
            }
            const int nf = (int)(tl.fb - tl.fa);
            for (int i = (int)((q - f) & 1u); i < nf; i += 2) {
                const int slot = (f + i) % WP_D1;
                c2 v[16];
                // ---- stage 2: radix-16 over n2; lane = n3
                bar_sync(WP_BAR_FULL + slot, WS_PC);
                {
                    const float2* E = sm.e1[slot] + k1 * 258 + lane16;
#pragma unroll
                    for (int n2 = 0; n2 < 16; ++n2) v[n2] = lds_c2(E + n2 * 16);
                }
                c2_fft16_first(v);  // consumes every loaded value: the slot can go back to the producers
                bar_arrive(WP_BAR_EMPTY + slot, WS_PC);
                c2_fft16_second(v);
#pragma unroll
                for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        const int k2 = ka + 4 * kb;
                        c2 x = v[4 * ka + kb];
                        if (k2 != 0) x = c2_cmul(x, tw2[k2].x, tw2[k2].y);
                        sts_c2(H + k2 * 17 + lane16, x);
                    }
                __syncwarp();
                // ---- stage 3: radix-16 over n3; lane = k2; bin = k1 + 8 k2 + 128 k3
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = lds_c2(H + lane16 * 17 + m);
                __syncwarp();
                c2_fft16(v);
                const unsigned int tile_i = q * WP_XB + (c % WP_XB), use = c / WP_XB;
                if (use > 0) mbar_wait(xempty0 + 8 * tile_i, (use - 1) & 1u);  // the tile's previous frame has been gathered
                const unsigned int xw = xw0 + (c % WP_XB) * (unsigned int)(WP_X * sizeof(float2));
#pragma unroll
                for (int ka = 0; ka < 4; ++ka)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                    {
                        const int k3 = ka + 4 * kb;
                        const unsigned int off = __byte_perm(xo[k3 >> 2], 0u, 0x4440u + (k3 & 3));
                        asm volatile("st.shared.b64 [%0], %1;" ::"r"(xw + off + (unsigned int)(k3 * 16 * sizeof(float2))), "l"(v[4 * ka + kb])
                                     : "memory");
                    }
                mbar_arrive_warp(xfull0 + 8 * tile_i);
                ++c;
            }
            f += nf;
        }
    }
}


}  // namespace (anonymous)

// launcher of the warp-specialised kernel; explicitly instantiated for P = 1..4 in pfb_wsp_p<P>.cu
template <int P>
int pfb_launch_ws(const PfbJob* jobs, int n_jobs, void* scratch, const float2* tw, int sm_count, cudaStream_t stream) {
    static DeviceOnce attr_once;
    constexpr size_t smem_bytes = sizeof(WpSmem);
    // Windows in pinned host memory (zero-copy calls) are latency-bound on PCIe reads: two input rows in flight per producer
    // thread instead of one (LA = 2) doubles the bytes a CTA keeps outstanding.
    bool host_window = false, sc16 = false;
    for (int j = 0; j < n_jobs; ++j) {
        host_window = host_window || jobs[j].min_tile > 0;
        sc16 = sc16 || jobs[j].in_sc16 != 0;
        if ((jobs[j].in_sc16 != 0) != (jobs[0].in_sc16 != 0)) {
            set_error("pfb_launch: fc32 and sc16 windows cannot share a launch");
            return -1;
        }
    }
    const bool table_mode = n_jobs > 1;
    auto pick = [&](auto table_tag) {
        constexpr bool TB = decltype(table_tag)::value;
        return sc16 ? (host_window ? pfb_fused_wsp_2048_kernel<P, 2, WP_HOIST, true, TB> : pfb_fused_wsp_2048_kernel<P, WP_LA, WP_HOIST, true, TB>)
                    : (host_window ? pfb_fused_wsp_2048_kernel<P, 2, WP_HOIST, false, TB> : pfb_fused_wsp_2048_kernel<P, WP_LA, WP_HOIST, false, TB>);
    };
    auto kernel = table_mode ? pick(std::true_type{}) : pick(std::false_type{});
    const bool classic = !table_mode && !sc16 && !host_window;   // the headline launch: the round-1 kernel (see its comment)
    if (const int dev = attr_once.pending(); dev >= 0) {
        auto raise = [&](auto k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes); };
        GSDR_CUDA_OK(raise(pfb_fused_wsp1_2048_kernel<P, WP_LA, WP_HOIST>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, WP_LA, WP_HOIST, false, false>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, 2, WP_HOIST, false, false>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, WP_LA, WP_HOIST, true, false>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, 2, WP_HOIST, true, false>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, WP_LA, WP_HOIST, false, true>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, 2, WP_HOIST, false, true>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, WP_LA, WP_HOIST, true, true>));
        GSDR_CUDA_OK(raise(pfb_fused_wsp_2048_kernel<P, 2, WP_HOIST, true, true>));
        attr_once.done(dev);
    }
    long long total_frames = 0;
    for (int j = 0; j < n_jobs; ++j) total_frames += jobs[j].n_frames;
    if (total_frames == 0) return 0;
    int frames_per_tile = (int)((total_frames + sm_count - 1) / sm_count);
    if (frames_per_tile < 1) frames_per_tile = 1;
    {   // experiment hook: a lower bound on the tile length (fewer, longer tiles: less halo re-read per frame)
        static int min_tile = -1;
        if (min_tile < 0) {
            const char* e = getenv("GSDR_PFB_MIN_TILE");
            min_tile = e ? atoi(e) : 0;
        }
        if (min_tile > frames_per_tile) frames_per_tile = min_tile;
        if (min_tile == 0 && n_jobs == 1 && jobs[0].min_tile > frames_per_tile) frames_per_tile = jobs[0].min_tile;
    }
    int grid_cap = 0;
    {   // test hook: a smaller grid makes every CTA walk several tiles; results must not change
        static int cap = -1;
        if (cap < 0) {
            const char* e = getenv("GSDR_PFB_MAX_GRID");
            cap = e ? atoi(e) : 0;
        }
        grid_cap = cap;
    }
    if (n_jobs > 1) {
        // ---- multi-stream launch: balanced explicit tile list ------------------------------------------------------------
        if (!scratch) {
            set_error("pfb_launch: multi-stream launch needs a job table buffer");
            return -1;
        }
        int grid = (int)(total_frames < sm_count ? total_frames : sm_count);
        // windows in pinned HOST memory (zero-copy group submit): reads over PCIe are not kept in L2, so the P-1 halo rows of
        // every tile cross the link again -- fewer, longer tiles (PfbJob::min_tile frames), never fewer than six CTAs
        int want_tile = 0;
        for (int j = 0; j < n_jobs; ++j) want_tile = jobs[j].min_tile > want_tile ? jobs[j].min_tile : want_tile;
        if (want_tile > 0) {
            long long g2 = total_frames / want_tile;
            g2 = g2 < 6 ? 6 : g2;
            if (g2 < grid) grid = (int)g2;
        }
        if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
        std::vector<PfbTile> tiles;
        std::vector<int> cta_begin;
        {
            std::vector<int> nf(n_jobs);
            for (int k = 0; k < n_jobs; ++k) nf[k] = jobs[k].n_frames;
            pfb_partition(nf.data(), n_jobs, grid, tiles, cta_begin);
        }
        // streams without frames in this launch still carry their tail: give them an empty tile on the last CTA
        for (int k = 0; k < n_jobs; ++k)
            if (jobs[k].n_frames == 0 && jobs[k].tail_dst && jobs[k].tail_n > 0) tiles.push_back(PfbTile{k, 0, 0, 1}), cta_begin[grid] = (int)tiles.size();
        const size_t off_tiles = (sizeof(PfbJob) * n_jobs + 15) & ~size_t(15);
        const size_t off_cta = off_tiles + sizeof(PfbTile) * tiles.size();
        const size_t bytes = off_cta + sizeof(int) * cta_begin.size();
        if (bytes > pfb_table_bytes(n_jobs, sm_count)) {
            set_error("pfb_launch: internal: tile table larger than its buffer");
            return -1;
        }
        std::vector<unsigned char> blob(bytes);
        memcpy(blob.data(), jobs, sizeof(PfbJob) * n_jobs);
        memcpy(blob.data() + off_tiles, tiles.data(), sizeof(PfbTile) * tiles.size());
        memcpy(blob.data() + off_cta, cta_begin.data(), sizeof(int) * cta_begin.size());
        unsigned char* base = static_cast<unsigned char*>(scratch);
        // < 64 KB from pageable memory: the driver embeds the bytes in the command stream, the call does not wait for the GPU
        GSDR_CUDA_OK(cudaMemcpyAsync(base, blob.data(), bytes, cudaMemcpyHostToDevice, stream));
        kernel<<<grid, WS_THREADS, smem_bytes, stream>>>(jobs[0], reinterpret_cast<const PfbJob*>(base), nullptr, n_jobs, 0, (int)tiles.size(), tw,
                                                         reinterpret_cast<const PfbTile*>(base + off_tiles),
                                                         reinterpret_cast<const int*>(base + off_cta));
        GSDR_CUDA_OK(cudaGetLastError());
        return 1;
    }
    const int total_tiles = (int)((total_frames + frames_per_tile - 1) / frames_per_tile);
    int grid = total_tiles < sm_count ? total_tiles : sm_count;
    if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
    if (classic)
        pfb_fused_wsp1_2048_kernel<P, WP_LA, WP_HOIST><<<grid, WS_THREADS, smem_bytes, stream>>>(jobs[0], nullptr, nullptr, 1, frames_per_tile, total_tiles, tw);
    else
        kernel<<<grid, WS_THREADS, smem_bytes, stream>>>(jobs[0], nullptr, nullptr, 1, frames_per_tile, total_tiles, tw, nullptr, nullptr);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}


}  // namespace gsdr
