// DIRECT-mode demodulator (multi-tone DDC + windowed-sinc FIR + decimation) for sm_100a.
//
// Replaces, per buffer, direct_demodulator_integer (cpp/kernels.cu:45-86: an fp64 sincospi and two
// 64-bit remainders per tone-sample, writing a T x L tone-major intermediate), T x FIR::run_fir
// (cpp/fir.cu:44-88: cublasCgemm + f x cublasCaxpy + 2 memcpy + memset per tone) and the
// cublasCgeam transpose (cpp/USRP_demodulator.cpp:422-433) with ONE kernel.
//
// The LO phase is an integer: phase(n) = (tf * n) mod R turns/R.  Because it is exactly additive,
//     y[p,ch] = sum_m h[m] x[n0+m] e^{-j theta_ch (n0+m)}
//             = e^{-j theta_ch n0} * sum_m ( h[m] e^{-j theta_ch m} ) x[n0+m],   n0 = (p-f+1) M
// i.e. a bank of per-tone complex FIRs g_ch[m] (built once, in double, from the same integer
// phases) followed by one rotation per OUTPUT sample.  No trig and no 64-bit remainder per input
// sample remain, and nothing but the decimated, sample-major result is written.
#include <cmath>

#include "devmath.cuh"
#include "direct_common.cuh"
#include "packed_f32x2.cuh"

namespace gsdr {
namespace {

constexpr int TP = 4;   // outputs per warp tile
constexpr int TC = 8;   // tones per warp tile
constexpr int WARPS = 8;

// kStaged: the block's input span is staged in shared memory once and reused by every tone group;
// otherwise (very long filters) the taps stream straight from global memory through L1/L2.
template <bool kStaged>
__global__ void __launch_bounds__(WARPS * 32, 2)
direct_fir_kernel(const Window w, const float2* __restrict__ g, const int* __restrict__ freq, int T, int M, int ntaps,
                  int rate, long long pos0, long long n_out, int PB, float2* __restrict__ out) {
    extern __shared__ __align__(16) float2 xs[];  // (PB-1)*M + ntaps samples
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p0 = (long long)blockIdx.x * PB;
    if (kStaged) {
        const int span = (PB - 1) * M + ntaps;
        for (int i = threadIdx.x; i < span; i += blockDim.x) xs[i] = dev_win_at(w, p0 * M + i);
        __syncthreads();
    }

    const int tone_groups = (T + TC - 1) / TC;
    const int out_groups = PB / TP;
    const double inv_R = 1.0 / (double)rate;
    for (int unit = warp + WARPS * blockIdx.y; unit < out_groups * tone_groups; unit += WARPS * gridDim.y) {
        const int og = unit % out_groups, tg = unit / out_groups;
        float2 acc[TP][TC];
#pragma unroll
        for (int a = 0; a < TP; ++a)
#pragma unroll
            for (int b = 0; b < TC; ++b) acc[a][b] = make_float2(0.f, 0.f);
        const float2* gbase[TC];
#pragma unroll
        for (int b = 0; b < TC; ++b) gbase[b] = g + (long long)min(tg * TC + b, T - 1) * ntaps;
        const float2* xbase = xs + (og * TP) * M;
        const long long gpos = (p0 + og * TP) * M;
        for (int m = lane; m < ntaps; m += 32) {
            float2 xv[TP], gv[TC];
#pragma unroll
            for (int a = 0; a < TP; ++a) xv[a] = kStaged ? xbase[a * M + m] : dev_win_at(w, gpos + (long long)a * M + m);
#pragma unroll
            for (int b = 0; b < TC; ++b) gv[b] = __ldg(gbase[b] + m);
#pragma unroll
            for (int a = 0; a < TP; ++a)
#pragma unroll
                for (int b = 0; b < TC; ++b) {
                    acc[a][b].x = fmaf(gv[b].x, xv[a].x, fmaf(-gv[b].y, xv[a].y, acc[a][b].x));
                    acc[a][b].y = fmaf(gv[b].x, xv[a].y, fmaf(gv[b].y, xv[a].x, acc[a][b].y));
                }
        }
        // warp-shuffle accumulation of the 16 partial sums
#pragma unroll
        for (int a = 0; a < TP; ++a)
#pragma unroll
            for (int b = 0; b < TC; ++b)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    acc[a][b].x += __shfl_xor_sync(0xffffffffu, acc[a][b].x, o);
                    acc[a][b].y += __shfl_xor_sync(0xffffffffu, acc[a][b].y, o);
                }
        // each lane finishes one (output, tone) of the 4 x 8 tile: rotate by the LO phase of its first tap and store
        float2 mine = make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < TP; ++a)
#pragma unroll
            for (int b = 0; b < TC; ++b)
                if (lane == a * TC + b) mine = acc[a][b];
        if (lane < TP * TC) {
            const int a = lane / TC, b = lane % TC;
            const long long p = p0 + og * TP + a;
            const int ch = tg * TC + b;
            if (p < n_out && ch < T) {
                // stream position of the first tap of output p: pos0 + p*M (pos0 already includes -(f-1)M)
                long long n0 = (pos0 + p * (long long)M) % rate;
                if (n0 < 0) n0 += rate;
                long long ph = direct_phase_signed(freq[ch], (unsigned long long)n0, rate);
                if (ph < 0) ph += rate;  // same residue class; the reference keeps the sign
                out[p * T + ch] = dev_cmul(mine, lo_phasor(ph, inv_R));
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// Register-tiled version (the default): lane = output, 16 tones per thread.
//
// The filter is M x f taps: y[p] = sum_{i<f} sum_{k<M} x[(p+i) M + k] g[i M + k].  A block owns 32*NW consecutive
// outputs and one group of 16 tones and walks the taps in chunks of KC columns k: the chunk's input columns of all
// PB+f-1 rows go to shared memory with an ODD row stride (KC+1), so the 32 lanes -- 32 consecutive outputs, M samples
// apart -- read 32 different bank pairs; the chunk's taps go to shared memory tone-fastest, so one tap of two tones is
// ONE broadcast 16-byte load.  Per tap and thread: 1 + 8 shared loads feed 32 packed FMAs (acc += x*gr + (j x)*gi),
// i.e. the kernel is bound by the fp32 pipe (16 T real FMAs per input sample), not by loads or shuffles; there is no
// cross-lane reduction at all.  The previous kernel (lanes stride the taps, 320 shuffles per 4 x 8 tile) stays as the
// fallback for filters with more than D_FMAX blocks.
// --------------------------------------------------------------------------------------------
constexpr int D_TC = 16;    // tones per thread
constexpr int D_KC = 32;    // tap columns per staged chunk
constexpr int D_GST = 18;   // float2 stride of one staged tap (16 tones + 2 pad: 16-byte aligned, spreads the staging stores)
constexpr int D_FMAX = 8;

template <int NW>
__global__ void __launch_bounds__(32 * NW, 4)
direct_fir_tiled_kernel(const Window w, const float2* __restrict__ g, const int* __restrict__ freq, int T, int M, int f, int rate,
                        long long pos0, long long n_out, float2* __restrict__ out) {
    extern __shared__ __align__(16) float2 dsm[];
    constexpr int PB = 32 * NW;
    const int rows = PB + f - 1;
    float2* xs = dsm;                              // [rows][D_KC + 1]
    float2* gs = dsm + rows * (D_KC + 1);          // [f][D_KC][D_GST]
    if ((reinterpret_cast<uintptr_t>(gs) & 15) != 0) gs += 1;  // 16-byte alignment for the broadcast loads
    const int tid = threadIdx.x;
    const long long p0 = (long long)blockIdx.x * PB;
    const int ch0 = blockIdx.y * D_TC;
    const int ntaps = f * M;

    // Blocked summation: every (tap chunk, FIR block) pair -- at most D_KC taps -- is its own fp32 chain `acc`, folded into
    // `tot` when it ends.  One chain over all f*M taps (400 at cfg1) leaves 3 - 4e-7 relative L2 against fp64; the reference's
    // cuBLAS Cgemm per block of M taps plus f axpy (cpp/fir.cu:48-57) 1.2e-7; chains of <= 32 taps sit below that
    // (tools/direct_accum_emulation.py).
    c2 acc[D_TC], tot[D_TC];
#pragma unroll
    for (int b = 0; b < D_TC; ++b) tot[b] = c2_pack(0.f, 0.f);

    // equal chunks (M = 100 -> 4 x 25 rather than 32 + 32 + 32 + 4: every chunk pays the same staging round trip)
    const int n_chunks = (M + D_KC - 1) / D_KC;
    const int kstep = (M + n_chunks - 1) / n_chunks;
    for (int k0 = 0; k0 < M; k0 += kstep) {
        const int kc = min(kstep, M - k0);
        __syncthreads();
        // Staging: eight independent loads in flight per thread before the first store (one load -> one store per
        // iteration serialises ~50 global-memory latencies per chunk).  `inside`: the block's whole input span lies in
        // the `in` segment, plain pointer arithmetic instead of the history/zero-padding logic.
        {
            const int cnt = rows * kc;
            const long long span_lo = p0 * (long long)M, span_hi = (p0 + rows) * (long long)M;
            const bool inside = span_lo >= w.n_hist && span_hi <= w.n_hist + w.n_in;
            const float2* xin = w.in + (span_lo - w.n_hist) + k0;
            for (int base = 0; base < cnt; base += PB * 8) {
                float2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + tid + u * PB;
                    const int r = e / kc, kk = e - r * kc;
                    v[u] = make_float2(0.f, 0.f);
                    if (e < cnt) v[u] = inside ? __ldg(xin + (long long)r * M + kk) : dev_win_at(w, span_lo + (long long)r * M + k0 + kk);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + tid + u * PB;
                    const int r = e / kc, kk = e - r * kc;
                    if (e < cnt) xs[r * (D_KC + 1) + kk] = v[u];
                }
            }
            const int gcnt = f * D_TC * kc;
            for (int base = 0; base < gcnt; base += PB * 8) {
                float2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + tid + u * PB;
                    const int kk = e % kc, ch = (e / kc) % D_TC, i = e / (kc * D_TC);
                    v[u] = make_float2(0.f, 0.f);
                    if (e < gcnt && ch0 + ch < T) v[u] = __ldg(g + (long long)(ch0 + ch) * ntaps + i * M + k0 + kk);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + tid + u * PB;
                    const int kk = e % kc, ch = (e / kc) % D_TC, i = e / (kc * D_TC);
                    if (e < gcnt) gs[(i * D_KC + kk) * D_GST + ch] = v[u];
                }
            }
        }
        __syncthreads();
        for (int i = 0; i < f; ++i) {
            const float2* xr = xs + (tid + i) * (D_KC + 1);
            const float4* gr = reinterpret_cast<const float4*>(gs + i * D_KC * D_GST);
#pragma unroll
            for (int b = 0; b < D_TC; ++b) acc[b] = c2_pack(0.f, 0.f);
#pragma unroll 4
            for (int kk = 0; kk < kc; ++kk) {
                const float2 xv = xr[kk];
                const c2 x = c2_pack(xv.x, xv.y), jx = c2_pack(-xv.y, xv.x);
#pragma unroll
                for (int b2 = 0; b2 < D_TC / 2; ++b2) {
                    const float4 gg = gr[kk * (D_GST / 2) + b2];  // taps of two tones, same address in every lane
                    acc[2 * b2] = c2_fma_s(x, gg.x, acc[2 * b2]);
                    acc[2 * b2] = c2_fma_s(jx, gg.y, acc[2 * b2]);
                    acc[2 * b2 + 1] = c2_fma_s(x, gg.z, acc[2 * b2 + 1]);
                    acc[2 * b2 + 1] = c2_fma_s(jx, gg.w, acc[2 * b2 + 1]);
                }
            }
#pragma unroll
            for (int b = 0; b < D_TC; ++b) tot[b] = c2_add(tot[b], acc[b]);
        }
    }
    // rotate by the LO phase of the output's first tap (integer phase, cpp/kernels.cu:59-75) and store sample-major
    const long long p = p0 + tid;
    if (p < n_out) {
        const double inv_R = 1.0 / (double)rate;
        long long n0 = (pos0 + p * (long long)M) % rate;
        if (n0 < 0) n0 += rate;
#pragma unroll
        for (int b = 0; b < D_TC; ++b) {
            const int ch = ch0 + b;
            if (ch < T) {
                long long ph = direct_phase_signed(freq[ch], (unsigned long long)n0, rate);
                if (ph < 0) ph += rate;  // same residue class; the reference keeps the sign
                out[p * T + ch] = dev_cmul(c2_to(tot[b]), lo_phasor(ph, inv_R));
            }
        }
    }
}

// decim == 0: pure mixing, out[n*T + ch] = x[n] e^{-j theta_ch n} (cpp/USRP_demodulator.cpp:442-457).
// A block takes MIX_S consecutive samples and all tones.  Per tone, once per block: tf = freq mod R (positive residue) and
// the integer phase of the block's first sample, base = (tf * n0) mod R, in 64-bit integers.  Sample i of the block then
// has phase base + i*tf < 1025 R < 2^53, formed exactly by one DFMA, and the 32-bit phase word round(phase * 2^32 / R)
// mod 2^32 is the low mantissa word of a second DFMA with the constant 1.5 * 2^52: no remainder, no division and no
// int <-> double conversion per output.  Warps walk samples, lanes walk tones: stores are tone-fastest, coalesced.
constexpr int MIX_S_MAX = 1024;  // samples per block: about 8192 outputs (32 per thread), between 32 and 1024 samples
__global__ void __launch_bounds__(256)
direct_mix_kernel(const float2* __restrict__ in, long long n, const int* __restrict__ freq, int T, int rate, long long pos0,
                  int MIX_S, float2* __restrict__ out) {
    __shared__ float2 xs[MIX_S_MAX];
    __shared__ double2 bs[256];   // (base phase, phase step per sample) of the tones of the current group of 256
    const long long s0 = (long long)blockIdx.x * MIX_S;
    for (int i = threadIdx.x; i < MIX_S; i += blockDim.x) xs[i] = (s0 + i < n) ? in[s0 + i] : make_float2(0.f, 0.f);
    const double word_per_phase = 4294967296.0 / (double)rate;
    const long long nb = (pos0 + s0) % rate;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_here = (int)min((long long)MIX_S, n - s0);
    for (int c0 = 0; c0 < T; c0 += 256) {
        __syncthreads();
        const int ch = c0 + threadIdx.x;
        if (ch < T) {
            long long tf = (long long)freq[ch] % rate;
            if (tf < 0) tf += rate;   // same residue class as the reference's signed remainder (cpp/kernels.cu:59-75)
            const unsigned long long base = ((unsigned long long)tf * (unsigned long long)nb) % (unsigned long long)rate;
            bs[threadIdx.x] = make_double2((double)base, (double)tf);
        }
        __syncthreads();
        const int tones_here = min(256, T - c0);
        // few tones: a warp takes 32 / tp2 samples at once (tp2 = tones rounded up to a power of two), so that all lanes work
        int sh = 5;
        while (sh > 0 && (1 << (sh - 1)) >= tones_here) --sh;
        const int tp2 = 1 << sh, spw = 32 >> sh;
        for (int i0 = warp * spw; i0 < n_here; i0 += 8 * spw) {
            const int i = i0 + (lane >> sh);
            if (i >= n_here) continue;
            const float2 x = xs[i];
            const double di = (double)i;
            float2* row = out + (s0 + i) * T + c0;
            for (int cl = lane & (tp2 - 1); cl < tones_here; cl += tp2) {
                const double2 b = bs[cl];
                const double ph = fma(di, b.y, b.x);
                const unsigned int word = (unsigned int)__double2loint(fma(ph, word_per_phase, 6755399441055744.0));
                float sn, cs;
                sincos_phase32(word, sn, cs);
                row[cl] = dev_cmul(x, make_float2(cs, -sn));
            }
        }
    }
}

__global__ void direct_phase_probe_kernel(long long* __restrict__ out, unsigned int n, int tf, int rate,
                                          unsigned long long index_counter, unsigned long long n0) {
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = direct_phase_signed(tf, n0 + i + index_counter, rate);
}

// The tile form of the LO phase (direct_common.cuh: lo_phase_tile / lo_phase_row / lo_phase_word), as the epilogues of
// direct_fir_tc_kernel and direct_fir_i8_kernel evaluate it: row r of a tile -> (unreduced integer phase, 32-bit phase word).
__global__ void direct_tile_phase_probe_kernel(long long* __restrict__ phase, unsigned int* __restrict__ word, int n_rows, int tf,
                                               int rate, long long pos0, long long row0, int M) {
    const double2 bs = lo_phase_tile(tf, rate, pos0, row0, M);
    const double word_per_phase = 4294967296.0 / (double)rate;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        phase[r] = (long long)lo_phase_row(bs, (double)r);
        word[r] = lo_phase_word(bs, (double)r, word_per_phase);
    }
}

}  // namespace

int direct_tile_phase_probe_launch(long long* phase, unsigned int* word, int n_rows, int tone_freq, int rate, long long pos0,
                                   long long row0, int M, cudaStream_t stream) {
    if (n_rows <= 0) return 0;
    direct_tile_phase_probe_kernel<<<1, 128, 0, stream>>>(phase, word, n_rows, tone_freq, rate, pos0, row0, M);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

int direct_fir_launch(const Window& w, const float2* g, const int* freq_dev, int T, int M, int ntaps, int rate,
                      long long pos0, long long n_out, float2* out, cudaStream_t stream) {
    if (n_out <= 0) return 0;
    const int f = M > 0 ? ntaps / M : 0;
    if (M >= 1 && f >= 1 && f <= D_FMAX && f * M == ntaps) {
        // register-tiled kernel; 64-output blocks when 128-output blocks would leave SMs idle
        const int tone_groups = (T + D_TC - 1) / D_TC;
        // 128- or 64-output blocks, whichever fills its waves better (a 64-output block needs 36 KB of shared memory:
        // 6 per SM; a 128-output block 53 KB: 4 per SM).  Ties go to the larger block (taps are staged once per block).
        auto wave_eff = [&](int pb, int per_sm) {
            const double tiles = (double)((n_out + pb - 1) / pb) * tone_groups, slots = 148.0 * per_sm;
            return tiles / (std::ceil(tiles / slots) * slots);
        };
        const bool small = wave_eff(64, 6) > wave_eff(128, 4) + 0.02;
        const int PB2 = small ? 64 : 128;
        const size_t smem = ((size_t)(PB2 + f - 1) * (D_KC + 1) + (size_t)f * D_KC * D_GST + 2) * sizeof(float2);
        dim3 grid((unsigned)((n_out + PB2 - 1) / PB2), (unsigned)tone_groups);
        static DeviceOnce attr_once;
        if (const int dev = attr_once.pending(); dev >= 0) {
            const int cap = (int)(((size_t)(128 + D_FMAX - 1) * (D_KC + 1) + (size_t)D_FMAX * D_KC * D_GST + 2) * sizeof(float2));
            GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_tiled_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            attr_once.done(dev);
        }
        if (small)
            direct_fir_tiled_kernel<2><<<grid, 64, smem, stream>>>(w, g, freq_dev, T, M, f, rate, pos0, n_out, out);
        else
            direct_fir_tiled_kernel<4><<<grid, 128, smem, stream>>>(w, g, freq_dev, T, M, f, rate, pos0, n_out, out);
        GSDR_CUDA_OK(cudaGetLastError());
        return 1;
    }
    // outputs per block: as many as fit 96 KB of staged input, multiple of TP, at most 64
    int PB = 64;
    while (PB > TP && ((size_t)(PB - 1) * M + ntaps) * sizeof(float2) > 96 * 1024) PB -= TP;
    size_t smem = ((size_t)(PB - 1) * M + ntaps) * sizeof(float2);
    const bool staged = smem <= 96 * 1024;
    if (!staged) smem = 0;
    static DeviceOnce fallback_once;   // per device (the attribute is): the largest staging area this launcher ever asks for
    if (const int dev = fallback_once.pending(); dev >= 0) {
        GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        fallback_once.done(dev);
    }
    // Split a block's (output group, tone group) units over gridDim.y only when there are too few output
    // blocks to fill the GPU: every y-slice re-stages the same input span.
    const int units = (PB / TP) * ((T + TC - 1) / TC);
    const long long xblocks = (n_out + PB - 1) / PB;
    int gy = 1;
    while (xblocks * gy < 2 * 148 && gy * WARPS < units && gy < 16) gy *= 2;
    dim3 grid((unsigned)xblocks, gy);
    if (staged)
        direct_fir_kernel<true><<<grid, WARPS * 32, smem, stream>>>(w, g, freq_dev, T, M, ntaps, rate, pos0, n_out, PB, out);
    else
        direct_fir_kernel<false><<<grid, WARPS * 32, 0, stream>>>(w, g, freq_dev, T, M, ntaps, rate, pos0, n_out, PB, out);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

int direct_mix_launch(const float2* in, long long n, const int* freq_dev, int T, int rate, long long pos0, float2* out,
                      cudaStream_t stream) {
    if (n <= 0) return 0;
    int S = 8192 / (T > 0 ? T : 1);
    S = S < 32 ? 32 : (S > MIX_S_MAX ? MIX_S_MAX : S & ~31);
    direct_mix_kernel<<<(unsigned)((n + S - 1) / S), 256, 0, stream>>>(in, n, freq_dev, T, rate, pos0, S, out);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

int direct_phase_probe_launch(long long* out, unsigned int n, int tone_freq, int rate, unsigned long long index_counter,
                              unsigned long long n0, cudaStream_t stream) {
    if (n == 0) return 0;
    direct_phase_probe_kernel<<<64, 128, 0, stream>>>(out, n, tone_freq, rate, index_counter, n0);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

}  // namespace gsdr
