// TX multi-tone period-buffer synthesis for sm_100a.
//
// The reference builds the TONES buffer with one `rate`-point inverse cuFFT of a spectrum that
// has base[f>0 ? f : rate+f].x = ampl (cpp/kernels.cu:589-684), i.e.
//     x[n] = sum_t a_t exp(+2 pi j (k_t n mod R)/R),   n in [0, R),  k_t in [0, R)
// Here the same sum is evaluated directly as a complex GEMM whose operands are generated on the
// fly from INTEGER phases, so no 1.6 GB in-place FFT (2e8 points at 200 MS/s) is needed and the
// result does not depend on an FFT library's rounding:
//     n = n1*8192 + j*128 + i         (n1: tile, j < 64, i < 128)
//     x[n] = sum_t ( a_t W^{k_t n1 8192} * W^{k_t j 128} ) * W^{k_t i},   W = exp(2 pi j / R)
// A[j][t] (64 x T) is a per-tile coarse phasor times a per-tone table E2, B[t][i] (T x 128) is a
// per-tone table E3; both tables are built once in double precision by tones_tables_kernel.
#include "devmath.cuh"

namespace gsdr {
namespace {

constexpr int TJ = 64, TI = 128, TILE = TJ * TI, KC = 8;

// e^{+2 pi j ph / R}
__device__ __forceinline__ float2 up_phasor_f64(long long ph, int R) {
    double s, c;
    sincospi(2.0 * (double)ph / (double)R, &s, &c);
    return make_float2((float)c, (float)s);
}
__device__ __forceinline__ long long mulmod(long long a, long long b, long long R) {  // a,b in [0,R), R < 2^31
    return (a * b) % R;
}

// E2[t][j] = W^{k_t * (j*128 mod R)}, E3[t][i] = W^{k_t * i}
__global__ void tones_tables_kernel(const int* __restrict__ bins, int T, int R, float2* __restrict__ e2, float2* __restrict__ e3) {
    const int total = T * (TJ + TI);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int t = e / (TJ + TI), r = e - t * (TJ + TI);
        const long long k = bins[t];
        if (r < TJ) e2[t * TJ + r] = up_phasor_f64(mulmod(k, ((long long)r * TI) % R, R), R);
        else e3[t * TI + (r - TJ)] = up_phasor_f64(mulmod(k, (r - TJ) % R, R), R);
    }
}

// 256 threads: ty = tid/16 owns rows j = 4*ty..4*ty+3, tx = tid%16 owns columns i = tx + 16*b.
__global__ void __launch_bounds__(256)
tones_synth_kernel(float2* __restrict__ out, long long n0, long long n, const int* __restrict__ bins,
                   const float* __restrict__ ampl, const float2* __restrict__ e2, const float2* __restrict__ e3, int T, int R) {
    __shared__ float2 As[KC][TJ];
    __shared__ float2 Bs[KC][TI];
    __shared__ float2 E1[KC];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long tile = blockIdx.x;
    const long long base = n0 + tile * TILE;  // absolute sample index (mod R handled in phases)
    float2 acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = make_float2(0.f, 0.f);

    for (int t0 = 0; t0 < T; t0 += KC) {
        __syncthreads();
        if (tid < KC) {
            const int t = t0 + tid;
            float2 v = make_float2(0.f, 0.f);
            if (t < T) {
                const float2 p = up_phasor_f64(mulmod(bins[t], base % R, R), R);
                v = make_float2(p.x * ampl[t], p.y * ampl[t]);
            }
            E1[tid] = v;
        }
        __syncthreads();
        for (int e = tid; e < KC * TJ; e += 256) {
            const int kk = e / TJ, j = e - kk * TJ, t = t0 + kk;
            As[kk][j] = (t < T) ? dev_cmul(E1[kk], e2[t * TJ + j]) : make_float2(0.f, 0.f);
        }
        for (int e = tid; e < KC * TI; e += 256) {
            const int kk = e / TI, i = e - kk * TI, t = t0 + kk;
            Bs[kk][i] = (t < T) ? e3[t * TI + i] : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float2 av[4], bv[8];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[kk][4 * ty + a];
#pragma unroll
            for (int b = 0; b < 8; ++b) bv[b] = Bs[kk][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    acc[a][b].x = fmaf(av[a].x, bv[b].x, fmaf(-av[a].y, bv[b].y, acc[a][b].x));
                    acc[a][b].y = fmaf(av[a].x, bv[b].y, fmaf(av[a].y, bv[b].x, acc[a][b].y));
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const long long s = tile * TILE + (long long)(4 * ty + a) * TI + tx + 16 * b;
            if (s < n) out[s] = acc[a][b];
        }
}

}  // namespace

// Synthesises out[s] = x[(n0+s) mod R] for s in [0,n).  bins_dev: k_t in [0,R); duplicates must be
// resolved by the caller (the reference's spectrum assignment keeps the last one).
int tones_synth_launch(float2* out, long long n0, long long n, const int* bins_dev, const float* ampl_dev, int T, int rate,
                       cudaStream_t stream) {
    if (n <= 0) return 0;
    if (T <= 0) {
        GSDR_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float2) * n, stream));
        return 0;
    }
    float2* tables = nullptr;
    GSDR_CUDA_OK(cudaMallocAsync(&tables, sizeof(float2) * (size_t)T * (TJ + TI), stream));
    float2* e2 = tables;
    float2* e3 = tables + (size_t)T * TJ;
    int tb = (T * (TJ + TI) + 255) / 256;
    if (tb > 1024) tb = 1024;
    tones_tables_kernel<<<tb, 256, 0, stream>>>(bins_dev, T, rate, e2, e3);
    const long long tiles = (n + TILE - 1) / TILE;
    tones_synth_kernel<<<(unsigned)tiles, 256, 0, stream>>>(out, n0 % rate, n, bins_dev, ampl_dev, e2, e3, T, rate);
    GSDR_CUDA_OK(cudaGetLastError());
    GSDR_CUDA_OK(cudaFreeAsync(tables, stream));
    return 2;
}

}  // namespace gsdr
