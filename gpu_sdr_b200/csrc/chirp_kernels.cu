// VNA chirp / swipe path for sm_100a.
//
// RX: replaces chirp_demodulator (cpp/kernels.cu:389-427) + cublas_decim (cpp/kernels.cu:852-872,
// a cublasCgemv against the flat lock-in profile) + move_buffer with ONE fused pass:
//     S[j] = sum_{i<ppt} in[j*ppt+i] * conj(chirp(pos0 + j*ppt + i)) * profile[i]
// so the 8 B/sample demodulated intermediate is never written.  The chirp phase is an integer
// accumulator: within a frequency step the int32 phase index advances by a constant, so the
// per-sample cost is one integer add plus an exact-range-reduced float32 sincos.
// TX: chirp_gen (cpp/kernels.cu:335-372) is the same phase walk with a store instead of a load.
#include <type_traits>

#include "devmath.cuh"

namespace gsdr {
namespace {

constexpr int SEG = 8192;  // samples per partial sum when ppt is large


// One warp per (output j, segment): lanes stride the samples, xor-shuffle reduction.
// profile == nullptr selects the flat lock-in window the reference always uses (make_flat_window:
// weight 0 for i < side, one constant after): the zero-weight head is skipped -- it is neither read
// nor demodulated -- and the constant is applied once to the sum.
__global__ void __launch_bounds__(256)
chirp_lockin_warp_kernel(const Window w, unsigned long long pos0, const ChirpDev cp, const float* __restrict__ profile,
                         int side, float flat_weight, int ppt, long long n_out, int n_seg, float2* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long units = n_out * n_seg;
    const bool flat = profile == nullptr;
    for (long long unit = warp0; unit < units; unit += n_warps) {
        const long long j = unit / n_seg;
        const int sg = (int)(unit - j * n_seg);
        int i0 = sg * SEG;
        const int i1 = min(ppt, i0 + SEG);
        if (flat && i0 < side) i0 = side;
        float2 acc = make_float2(0.f, 0.f);
        if (i0 + lane < i1) {
            const long long s0 = j * (long long)ppt;  // window index of the output's first sample
            ChirpWalker cw;
            cw.seek(pos0 + (unsigned long long)(s0 + i0 + lane), cp);
            // every sample of this unit inside the `in` segment: plain pointer walk
            const bool direct = (s0 + i0 >= w.n_hist) && (s0 + i1 <= w.n_hist + w.n_in);
            if (direct && flat) {
                const float2* p = w.in + (s0 - w.n_hist);
                int i = i0 + lane;
                // Eight, then four independent 256-byte warp loads in flight per iteration (one load per iteration
                // leaves the SM with too few bytes outstanding to cover HBM latency); same per-lane summation order.
                auto burst = [&](auto width_tag) {
                    constexpr int W = decltype(width_tag)::value;
                    for (; i + 32 * (W - 1) < i1; i += 32 * W) {
                        float2 x[W];
#pragma unroll
                        for (int u = 0; u < W; ++u) x[u] = __ldg(p + i + 32 * u);
#pragma unroll
                        for (int u = 0; u < W; ++u) {
                            const float2 ch = chirp_phasor(cw.idx);
                            acc.x = fmaf(ch.x, x[u].x, fmaf(ch.y, x[u].y, acc.x));
                            acc.y = fmaf(ch.x, x[u].y, fmaf(-ch.y, x[u].x, acc.y));
                            cw.advance(32u, cp);
                        }
                    }
                };
                burst(std::integral_constant<int, 8>{});
                burst(std::integral_constant<int, 4>{});
                for (; i < i1; i += 32) {
                    const float2 x = __ldg(p + i);
                    const float2 ch = chirp_phasor(cw.idx);
                    // out = in * conj(chirp): (cx*ix + cy*iy, cx*iy - cy*ix), cpp/kernels.cu:424-425
                    acc.x = fmaf(ch.x, x.x, fmaf(ch.y, x.y, acc.x));
                    acc.y = fmaf(ch.x, x.y, fmaf(-ch.y, x.x, acc.y));
                    cw.advance(32u, cp);
                }
            } else {
                for (int i = i0 + lane; i < i1; i += 32) {
                    const float2 x = dev_win_at(w, s0 + i);
                    const float2 ch = chirp_phasor(cw.idx);
                    if (flat) {  // same arithmetic as the pointer walk: results do not depend on the path taken
                        acc.x = fmaf(ch.x, x.x, fmaf(ch.y, x.y, acc.x));
                        acc.y = fmaf(ch.x, x.y, fmaf(-ch.y, x.x, acc.y));
                    } else {
                        const float pw = __ldg(&profile[i]);
                        acc.x = fmaf(fmaf(ch.x, x.x, ch.y * x.y), pw, acc.x);
                        acc.y = fmaf(fmaf(ch.x, x.y, -ch.y * x.x), pw, acc.y);
                    }
                    cw.advance(32u, cp);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (flat) {
            acc.x *= flat_weight;
            acc.y *= flat_weight;
        }
        if (lane == 0) dst[unit] = acc;
    }
}

// Short integration windows (ppt < 32): one thread per output.
__global__ void __launch_bounds__(256)
chirp_lockin_thread_kernel(const Window w, unsigned long long pos0, const ChirpDev cp, const float* __restrict__ profile,
                           int ppt, long long n_out, float2* __restrict__ out) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_out; j += (long long)gridDim.x * blockDim.x) {
        ChirpWalker cw;
        cw.seek(pos0 + (unsigned long long)(j * ppt), cp);
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < ppt; ++i) {
            const float2 x = dev_win_at(w, j * (long long)ppt + i);
            const float2 ch = chirp_phasor(cw.idx);
            const float pw = __ldg(&profile[i]);
            acc.x = fmaf(fmaf(ch.x, x.x, ch.y * x.y), pw, acc.x);
            acc.y = fmaf(fmaf(ch.x, x.y, -ch.y * x.x), pw, acc.y);
            cw.advance(1u, cp);
        }
        out[j] = acc;
    }
}

__global__ void chirp_lockin_finalize_kernel(const float2* __restrict__ partial, int n_seg, long long n_out, float2* __restrict__ out) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n_out; j += (long long)gridDim.x * blockDim.x) {
        float2 acc = make_float2(0.f, 0.f);
        for (int s = 0; s < n_seg; ++s) {
            const float2 p = partial[j * n_seg + s];
            acc.x += p.x;
            acc.y += p.y;
        }
        out[j] = acc;
    }
}

// decim == 0: plain demodulation, one output per input (cpp/USRP_demodulator.cpp:384-391).
// A warp owns a tile of 32 x RUN consecutive samples, lanes stride it: every load and store of the warp is one
// contiguous 256-byte line pair, all RUN loads are issued before the first is used.
constexpr int RUN = 8;
__global__ void __launch_bounds__(256)
chirp_demod_full_kernel(const float2* __restrict__ in, long long n, unsigned long long pos0, const ChirpDev cp,
                        float2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int TILE = 32 * RUN;
    const long long tiles = (n + TILE - 1) / TILE;
    for (long long tile = warp0; tile < tiles; tile += n_warps) {
        const long long s0 = tile * TILE + lane;
        if (s0 >= n) continue;
        float2 x[RUN];
#pragma unroll
        for (int i = 0; i < RUN; ++i) x[i] = (s0 + 32 * i < n) ? __ldg(in + s0 + 32 * i) : make_float2(0.f, 0.f);
        ChirpWalker cw;
        cw.seek(pos0 + (unsigned long long)s0, cp);
#pragma unroll
        for (int i = 0; i < RUN; ++i) {
            const float2 ch = chirp_phasor(cw.idx);
            if (s0 + 32 * i < n) out[s0 + 32 * i] = make_float2(fmaf(ch.x, x[i].x, ch.y * x[i].y), fmaf(ch.x, x[i].y, -ch.y * x[i].x));
            cw.advance(32u, cp);
        }
    }
}

// TX: out = scale * (sin(pi theta), -cos(pi theta)), cpp/kernels.cu:367-368.
__global__ void __launch_bounds__(256)
chirp_gen_kernel(float2* __restrict__ out, long long n, unsigned long long pos0, const ChirpDev cp, float scale) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int TILE = 32 * 16;  // samples per warp tile
    const long long tiles = (n + TILE - 1) / TILE;
    for (long long tile = warp0; tile < tiles; tile += n_warps) {
        const long long s0 = tile * TILE + lane;
        if (s0 >= n) continue;
        ChirpWalker cw;
        cw.seek(pos0 + (unsigned long long)s0, cp);
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const long long s = s0 + 32 * i;
            if (s < n) {
                const float2 ch = chirp_phasor(cw.idx);
                out[s] = make_float2(ch.x * scale, ch.y * scale);
                cw.advance(32u, cp);
            }
        }
    }
}

__global__ void chirp_index_probe_kernel(int* __restrict__ out, unsigned int n, unsigned long long pos0, const ChirpDev cp) {
    // two ways on purpose: a fresh seek per element and an incremental walk must agree
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int stride = gridDim.x * blockDim.x;
    if (t >= n) return;
    ChirpWalker cw;
    cw.seek(pos0 + t, cp);
    for (unsigned int i = t; i < n; i += stride) {
        out[i] = (int)cw.idx;
        cw.advance(stride, cp);
    }
}

}  // namespace

size_t chirp_partial_count(int ppt, long long n_out, int) {
    const int n_seg = (ppt + SEG - 1) / SEG;
    return n_seg > 1 ? (size_t)n_out * n_seg : 0;
}

int chirp_demod_launch(const Window& w, unsigned long long pos0, const ChirpDev& cp, const float* profile, int side,
                       float flat_weight, int ppt, long long n_out, float2* out, float2* partial, int sm_count,
                       cudaStream_t stream) {
    if (n_out <= 0) return 0;
    int launches = 0;
    if (ppt < 32) {
        int blocks = (int)((n_out + 255) / 256);
        if (blocks > sm_count * 8) blocks = sm_count * 8;
        chirp_lockin_thread_kernel<<<blocks, 256, 0, stream>>>(w, pos0, cp, profile, ppt, n_out, out);
        launches = 1;
    } else {
        const int n_seg = (ppt + SEG - 1) / SEG;
        const long long units = n_out * n_seg;
        long long blocks = (units + 7) / 8;  // 8 warps per block
        if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
        if (n_seg > 1 && !partial) {
            set_error("chirp_demod_launch: partial buffer missing");
            return -1;
        }
        // the flat window is applied in-kernel (profile pointer not needed); an explicit profile array is only
        // used when a caller supplies a non-flat one
        chirp_lockin_warp_kernel<<<(int)blocks, 256, 0, stream>>>(w, pos0, cp, side >= 0 ? nullptr : profile, side, flat_weight, ppt,
                                                                   n_out, n_seg, n_seg > 1 ? partial : out);
        launches = 1;
        if (n_seg > 1) {
            int fb = (int)((n_out + 255) / 256);
            if (fb > 1024) fb = 1024;
            chirp_lockin_finalize_kernel<<<fb, 256, 0, stream>>>(partial, n_seg, n_out, out);
            launches = 2;
        }
    }
    GSDR_CUDA_OK(cudaGetLastError());
    return launches;
}

int chirp_demod_full_launch(const float2* in, long long n, unsigned long long pos0, const ChirpDev& cp, float2* out,
                            cudaStream_t stream) {
    if (n <= 0) return 0;
    long long blocks = ((n + 32 * RUN - 1) / (32 * RUN) + 7) / 8;  // 8 warps per block, one tile per warp and pass
    if (blocks > 148 * 16) blocks = 148 * 16;
    chirp_demod_full_kernel<<<(int)blocks, 256, 0, stream>>>(in, n, pos0, cp, out);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

int chirp_gen_launch(float2* out, long long n, unsigned long long pos0, const ChirpDev& cp, float scale, cudaStream_t stream) {
    if (n <= 0) return 0;
    long long blocks = ((n + 511) / 512 + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    chirp_gen_kernel<<<(int)blocks, 256, 0, stream>>>(out, n, pos0, cp, scale);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

int chirp_index_probe_launch(int* out, unsigned int n, unsigned long long pos0, const ChirpDev& cp, cudaStream_t stream) {
    if (n == 0) return 0;
    chirp_index_probe_kernel<<<64, 128, 0, stream>>>(out, n, pos0, cp);
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}

}  // namespace gsdr
