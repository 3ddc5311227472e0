// TONES-mode polyphase-filter-bank channelizer for sm_100a.
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

#include "common.hpp"

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
    unsigned short bins[FN];
    int job_index;
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
    float w[P][2][2];  // taps of this thread's 4 columns

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        // ---- locate the job (stream) this tile belongs to --------------------------------------
        int j = 0;
        if (table != nullptr) {
            while (j + 1 < n_jobs && tile >= tile_begin[j + 1]) ++j;
        }
        const PfbJob job = (table != nullptr) ? table[j] : single;
        const int tile_in_job = tile - ((table != nullptr) ? tile_begin[j] : 0);
        if (j != loaded_job) {
            __syncthreads();  // previous tile's gather may still read sm.bins
            for (int u = t; u < job.T; u += FTHREADS) sm.bins[u] = job.bins ? (unsigned short)job.bins[u] : (unsigned short)u;
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
        // prologue: the P-1 rows older than the first new row of iteration 0
#pragma unroll
        for (int i = 0; i < P - 1; ++i) {
            const long long row = frame0 + i;
            if (aligned && row >= fast_lo && row < fast_hi) load_row<true>(win, row, t, r[i]);
            else load_row<false>(win, row, t, r[i]);
        }

        for (int it = 0; it < n_it; ++it) {
            const long long fbase = frame0 + (long long)it * FTEAMS;
            // ---- new rows of this iteration -----------------------------------------------------
            {
                const long long row_lo = fbase + P - 1, row_hi = row_lo + FTEAMS;
                if (aligned && row_lo >= fast_lo && row_hi <= fast_hi) {
#pragma unroll
                    for (int q = 0; q < FTEAMS; ++q) load_row<true>(win, row_lo + q, t, r[P - 1 + q]);
                } else {
#pragma unroll
                    for (int q = 0; q < FTEAMS; ++q) load_row<false>(win, row_lo + q, t, r[P - 1 + q]);
                }
                // pull the next iteration's rows towards L2 while the FFTs run
                if (it + 1 < n_it && (t & 7) == 0) {
                    const long long s = (row_hi * FN - win.n_hist) + 2 * (long long)t;
                    if (s >= 0) {
#pragma unroll
                        for (int q = 0; q < FTEAMS; ++q) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const long long sq = s + (long long)q * FN + 2 * c * FTHREADS;
                                if (sq + 16 <= win.n_in)
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(win.in + sq));
                            }
                        }
                    }
                }
            }
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
                team_barrier(team, FTEAM_THREADS);
                // tone selection: coalesced sample-major store
                float2* o = job.out + (frame - job.first_frame) * (long long)job.T;
                for (int u = tid; u < job.T; u += FTEAM_THREADS) o[u] = A[sm.bins[u]];
            }
            __syncthreads();
        }
    }
}

// --------------------------------------------------------------------------------------------
// generic path: any N, P, T
// --------------------------------------------------------------------------------------------
__global__ void pfb_fir_generic_kernel(const PfbJob job, float2* __restrict__ z) {
    const long long total = (long long)job.n_frames * job.N;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const long long f = o / job.N;
        const int k = (int)(o - f * job.N);
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < job.P; ++i) {
            const float2 x = win_at(job.win, (job.first_frame + f + i) * (long long)job.N + k);
            const float wt = job.taps[i * job.N + k];
            acc.x = fmaf(x.x, wt, acc.x);
            acc.y = fmaf(x.y, wt, acc.y);
        }
        z[o] = acc;
    }
}

constexpr int GEN_CHUNK = 2048;   // z samples staged in shared memory per pass
constexpr int GEN_THREADS = 128;  // tones per block

// out[f][u] = sum_k z[f][k] W_N^(k bin_u); W table in global memory (built in double on the host).
__global__ void __launch_bounds__(GEN_THREADS)
pfb_dft_generic_kernel(const PfbJob job, const float2* __restrict__ z, const float2* __restrict__ wtab) {
    __shared__ float2 zs[GEN_CHUNK];
    const int f = blockIdx.x;
    const int u = blockIdx.y * GEN_THREADS + threadIdx.x;
    const int N = job.N;
    const bool live = u < job.T;
    int bin = 0;
    if (live) bin = job.bins ? job.bins[u] : u;
    bin = ((bin % N) + N) % N;
    float2 total = make_float2(0.f, 0.f);
    float2 comp = make_float2(0.f, 0.f);  // Kahan compensation of the chunk-level sum
    int idx = 0;                          // (k * bin) mod N
    for (int k0 = 0; k0 < N; k0 += GEN_CHUNK) {
        const int len = min(GEN_CHUNK, N - k0);
        __syncthreads();
        for (int k = threadIdx.x; k < len; k += GEN_THREADS) zs[k] = z[(long long)f * N + k0 + k];
        __syncthreads();
        if (live) {
            for (int kb = 0; kb < len; kb += 32) {
                float2 part = make_float2(0.f, 0.f);
                const int ke = min(kb + 32, len);
                for (int k = kb; k < ke; ++k) {
                    const float2 wv = __ldg(&wtab[idx]);
                    const float2 x = zs[k];
                    part.x = fmaf(x.x, wv.x, fmaf(-x.y, wv.y, part.x));
                    part.y = fmaf(x.x, wv.y, fmaf(x.y, wv.x, part.y));
                    idx += bin;
                    if (idx >= N) idx -= N;
                }
                // compensated accumulate of the 32-term partial
                float yx = part.x - comp.x, yy = part.y - comp.y;
                float tx = total.x + yx, ty = total.y + yy;
                comp.x = (tx - total.x) - yx;
                comp.y = (ty - total.y) - yy;
                total.x = tx;
                total.y = ty;
            }
        }
    }
    if (live) job.out[(long long)f * job.T + u] = total;
}

__global__ void window_tail_kernel(const Window w, long long first, long long n, float2* __restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = win_at(w, first + i);
}

}  // namespace

bool pfb_fused_supported(int N, int P, int T, const Window&) {
    return N == FN && P >= 1 && P <= 4 && T >= 1 && T <= FN;
}

const char* pfb_kernel_name(int N, int P, int T) {
    Window w{};
    if (pfb_fused_supported(N, P, T, w)) {
        switch (P) {
            case 1: return "pfb_fused_2048_kernel<1>";
            case 2: return "pfb_fused_2048_kernel<2>";
            case 3: return "pfb_fused_2048_kernel<3>";
            default: return "pfb_fused_2048_kernel<4>";
        }
    }
    return "pfb_fir_generic_kernel+pfb_dft_generic_kernel";
}

size_t pfb_workspace_bytes(int N, int P, int max_frames) {
    (void)P;
    return (size_t)max_frames * (size_t)N * sizeof(float2);
}

template <int P>
static int launch_fused(const PfbJob* jobs, int n_jobs, void* scratch, const float2* tw, int sm_count, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        GSDR_CUDA_OK(cudaFuncSetAttribute(pfb_fused_2048_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(FusedSmem)));
        attr_set = true;
    }
    long long total_iters = 0;
    for (int j = 0; j < n_jobs; ++j) total_iters += (jobs[j].n_frames + FTEAMS - 1) / FTEAMS;
    if (total_iters == 0) return 0;
    // one tile per CTA and wave where possible; the P-1 halo rows re-read per tile stay negligible
    int iters_per_tile = (int)((total_iters + sm_count - 1) / sm_count);
    if (iters_per_tile < 1) iters_per_tile = 1;
    std::vector<int> tile_begin(n_jobs + 1, 0);
    for (int j = 0; j < n_jobs; ++j) {
        const int it = (jobs[j].n_frames + FTEAMS - 1) / FTEAMS;
        tile_begin[j + 1] = tile_begin[j] + (it + iters_per_tile - 1) / iters_per_tile;
    }
    const int total_tiles = tile_begin[n_jobs];
    const int grid = total_tiles < sm_count ? total_tiles : sm_count;
    const PfbJob* table = nullptr;
    const int* tb = nullptr;
    if (n_jobs > 1) {
        if (!scratch) {
            set_error("pfb_launch: multi-stream launch needs a job table buffer");
            return -1;
        }
        unsigned char* base = static_cast<unsigned char*>(scratch);
        GSDR_CUDA_OK(cudaMemcpyAsync(base, jobs, sizeof(PfbJob) * n_jobs, cudaMemcpyHostToDevice, stream));
        const size_t off = (sizeof(PfbJob) * n_jobs + 15) & ~size_t(15);
        GSDR_CUDA_OK(cudaMemcpyAsync(base + off, tile_begin.data(), sizeof(int) * (n_jobs + 1), cudaMemcpyHostToDevice, stream));
        table = reinterpret_cast<const PfbJob*>(base);
        tb = reinterpret_cast<const int*>(base + off);
    }
    pfb_fused_2048_kernel<P><<<grid, FTHREADS, sizeof(FusedSmem), stream>>>(jobs[0], table, tb, n_jobs, iters_per_tile,
                                                                             total_tiles, tw);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

int pfb_launch(const PfbJob* jobs, int n_jobs, void* scratch, void* workspace, const float2* tw, int sm_count,
               cudaStream_t stream) {
    if (n_jobs <= 0) return 0;
    bool fused = true;
    for (int j = 0; j < n_jobs; ++j)
        fused = fused && pfb_fused_supported(jobs[j].N, jobs[j].P, jobs[j].T, jobs[j].win) && jobs[j].P == jobs[0].P;
    if (fused) {
        switch (jobs[0].P) {
            case 1: return launch_fused<1>(jobs, n_jobs, scratch, tw, sm_count, stream);
            case 2: return launch_fused<2>(jobs, n_jobs, scratch, tw, sm_count, stream);
            case 3: return launch_fused<3>(jobs, n_jobs, scratch, tw, sm_count, stream);
            default: return launch_fused<4>(jobs, n_jobs, scratch, tw, sm_count, stream);
        }
    }
    int launches = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const PfbJob& job = jobs[j];
        if (job.n_frames <= 0) continue;
        if (!workspace || !tw) {
            set_error("pfb_launch: generic path needs workspace and W_N table");
            return -1;
        }
        float2* z = static_cast<float2*>(workspace);
        const long long total = (long long)job.n_frames * job.N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sm_count * 16) blocks = sm_count * 16;
        pfb_fir_generic_kernel<<<blocks, 256, 0, stream>>>(job, z);
        dim3 grid(job.n_frames, (job.T + GEN_THREADS - 1) / GEN_THREADS);
        pfb_dft_generic_kernel<<<grid, GEN_THREADS, 0, stream>>>(job, z, tw);
        GSDR_CUDA_OK(cudaGetLastError());
        launches += 2;
    }
    return launches;
}

int window_tail_copy(const Window& w, long long n_tail, float2* dst, cudaStream_t stream) {
    if (n_tail <= 0) return 0;
    const long long first = w.n_hist + w.n_in - n_tail;
    int blocks = (int)((n_tail + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    window_tail_kernel<<<blocks, 256, 0, stream>>>(w, first, n_tail, dst);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

}  // namespace gsdr
