// TONES-mode polyphase-filter-bank channelizer for sm_100a: launch logic, the generic (any N, P, T) path, spectral
// decimation and carry-over copies.  The fused kernels live in pfb_fused.cuh.
#include "pfb_fused.cuh"

namespace gsdr {

extern template int pfb_launch_ws<1>(const PfbJob*, int, void*, const float2*, int, cudaStream_t);
extern template int pfb_launch_ws<2>(const PfbJob*, int, void*, const float2*, int, cudaStream_t);
extern template int pfb_launch_ws<3>(const PfbJob*, int, void*, const float2*, int, cudaStream_t);
extern template int pfb_launch_ws<4>(const PfbJob*, int, void*, const float2*, int, cudaStream_t);

namespace {

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

// NOISE mode with decim > 0 (process_pfb_spec + decimate_spectra, cpp/USRP_demodulator.cpp:568-649,
// cpp/kernels.cu:704-749): every `d` consecutive spectra are averaged bin by bin.  The reference accumulates with
// float atomics into a buffer it never zeroes and scales on another stream; what is built here is the function it
// is meant to compute, out[g][k] = (1/d) sum_{j<d} X[g d + j][k], with the running sum carried across calls so group
// boundaries do not depend on how the stream was cut into buffers.  One thread per (group, bin), sequential in j:
// coalesced across bins, deterministic, bit-identical for any buffering.
__global__ void __launch_bounds__(256)
spectra_decimate_kernel(const float2* __restrict__ spec, long long n_frames, int N, int d, int carried /* frames already in acc */,
                        float2* __restrict__ acc /* [N] running sum of the open group */, float2* __restrict__ out, long long n_groups) {
    const float inv_d = 1.0f / (float)d;
    // units 0..n_groups-1 close a group; unit n_groups folds the remaining frames into acc
    const long long units = (n_groups + 1) * (long long)N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < units; e += (long long)gridDim.x * blockDim.x) {
        const long long grp = e / N;
        const int k = (int)(e - grp * N);
        long long f0 = grp * d - carried, f1 = f0 + d;
        if (grp == n_groups) f1 = n_frames;
        float2 sum = make_float2(0.f, 0.f);
        if (f0 < 0) {  // the open group starts in an earlier call
            f0 = 0;
            if (carried > 0) sum = acc[k];
        }
        for (long long fr = f0; fr < f1; ++fr) {
            const float2 v = spec[fr * N + k];
            sum.x += v.x;
            sum.y += v.y;
        }
        if (grp < n_groups) out[grp * N + k] = make_float2(sum.x * inv_d, sum.y * inv_d);
        else acc[N + k] = sum;  // staged: group 0's threads may still be reading acc[k]; spectra_commit_kernel moves it
    }
}
__global__ void spectra_commit_kernel(float2* __restrict__ acc, int N) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) acc[k] = acc[N + k];
}

struct TailJob {
    Window w;
    long long first, n;
    float2* dst;
};
__global__ void window_tail_multi_kernel(const TailJob* __restrict__ jobs) {
    const TailJob tj = jobs[blockIdx.y];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tj.n; i += (long long)gridDim.x * blockDim.x)
        tj.dst[i] = win_at(tj.w, tj.first + i);
}

__global__ void window_tail_kernel(const Window w, long long first, long long n, float2* __restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = win_at(w, first + i);
}

}  // namespace

bool pfb_fused_supported(int N, int P, int T, const Window&) {
    return N == FN && P >= 1 && P <= 4 && T >= 1 && T <= FN;
}

static int pfb_variant();
const char* pfb_kernel_name(int N, int P, int T) {
    Window w{};
    if (pfb_fused_supported(N, P, T, w) && pfb_variant() == 1) {
        switch (P) {
            case 1: return "pfb_fused_wsp_2048_kernel<1>";
            case 2: return "pfb_fused_wsp_2048_kernel<2>";
            case 3: return "pfb_fused_wsp_2048_kernel<3>";
            default: return "pfb_fused_wsp_2048_kernel<4>";
        }
    }
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
    static DeviceOnce attr_once;
    if (const int dev = attr_once.pending(); dev >= 0) {
        GSDR_CUDA_OK(cudaFuncSetAttribute(pfb_fused_2048_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(FusedSmem)));
        attr_once.done(dev);
    }
    long long total_iters = 0;
    for (int j = 0; j < n_jobs; ++j) total_iters += (jobs[j].n_frames + FTEAMS - 1) / FTEAMS;
    if (total_iters == 0) return 0;
    // one tile per CTA and wave where possible; the P-1 halo rows re-read per tile stay negligible
    int iters_per_tile = (int)((total_iters + sm_count - 1) / sm_count);
    if (iters_per_tile < 1) iters_per_tile = 1;
    std::vector<int> tile_begin(n_jobs + 1, 0);
    for (;; ++iters_per_tile) {  // keep the launch to one wave (see launch_ws)
        for (int j = 0; j < n_jobs; ++j) {
            const int it = (jobs[j].n_frames + FTEAMS - 1) / FTEAMS;
            tile_begin[j + 1] = tile_begin[j] + (it + iters_per_tile - 1) / iters_per_tile;
        }
        if (tile_begin[n_jobs] <= sm_count || n_jobs > sm_count) break;
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

size_t pfb_table_bytes(int n_jobs, int sm_count) {
    return ((sizeof(PfbJob) * (size_t)n_jobs + 15) & ~size_t(15)) + sizeof(PfbTile) * (2 * (size_t)sm_count + 2 * (size_t)n_jobs + 8) +
           sizeof(int) * ((size_t)sm_count + 2) + 256;
}

// 0 = lock-step kernel (kept as an independent second implementation for cross-checks),
// 1 = packed warp-specialised kernel (default).  GSDR_PFB_VARIANT=lockstep selects 0.
static int pfb_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GSDR_PFB_VARIANT");
        v = (e && e[0] == 'l') ? 0 : 1;
    }
    return v;
}

bool pfb_fused_sc16_available() { return pfb_variant() == 1; }

int pfb_fused_twiddle_count() { return FTW1 + FTW2 + WS_TW1 + WS_TW2; }

int pfb_launch(const PfbJob* jobs, int n_jobs, void* scratch, void* workspace, const float2* tw, int sm_count,
               cudaStream_t stream) {
    if (n_jobs <= 0) return 0;
    bool fused = true;
    for (int j = 0; j < n_jobs; ++j)
        fused = fused && pfb_fused_supported(jobs[j].N, jobs[j].P, jobs[j].T, jobs[j].win) && jobs[j].P == jobs[0].P;
    // carry-over requested through the job (single stream): done by the warp-specialised kernel itself, else copied here
    auto tail_after = [&](int nl, bool kernel_did_it) {
        if (nl < 0 || n_jobs != 1 || !jobs[0].tail_dst || jobs[0].tail_n <= 0 || kernel_did_it) return nl;
        const int tl = window_tail_copy(jobs[0].win, jobs[0].tail_n, jobs[0].tail_dst, stream);
        return tl < 0 ? -1 : nl + tl;
    };
    if (fused && pfb_variant() == 1) {
        const float2* tws = tw + FTW1 + FTW2;
        int nl;
        switch (jobs[0].P) {
            case 1: nl = pfb_launch_ws<1>(jobs, n_jobs, scratch, tws, sm_count, stream); break;
            case 2: nl = pfb_launch_ws<2>(jobs, n_jobs, scratch, tws, sm_count, stream); break;
            case 3: nl = pfb_launch_ws<3>(jobs, n_jobs, scratch, tws, sm_count, stream); break;
            default: nl = pfb_launch_ws<4>(jobs, n_jobs, scratch, tws, sm_count, stream); break;
        }
        return tail_after(nl, nl > 0);   // nl == 0: no frames, no launch
    }
    for (int j = 0; j < n_jobs; ++j)
        if (jobs[j].in_sc16) {
            set_error("pfb_launch: sc16 windows are read by the warp-specialised fused kernel only");
            return -1;
        }
    if (fused) {
        int nl;
        switch (jobs[0].P) {
            case 1: nl = launch_fused<1>(jobs, n_jobs, scratch, tw, sm_count, stream); break;
            case 2: nl = launch_fused<2>(jobs, n_jobs, scratch, tw, sm_count, stream); break;
            case 3: nl = launch_fused<3>(jobs, n_jobs, scratch, tw, sm_count, stream); break;
            default: nl = launch_fused<4>(jobs, n_jobs, scratch, tw, sm_count, stream); break;
        }
        return tail_after(nl, false);
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
    return tail_after(launches, false);
}

// acc: device buffer of 2*N float2 (running sum + staging).  Returns launches or -1.
int spectra_decimate_launch(const float2* spec, long long n_frames, int N, int d, int carried, float2* acc, float2* out,
                            long long n_groups, int sm_count, cudaStream_t stream) {
    if (n_frames <= 0) return 0;
    const long long units = (n_groups + 1) * (long long)N;
    long long blocks = (units + 255) / 256;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    spectra_decimate_kernel<<<(int)blocks, 256, 0, stream>>>(spec, n_frames, N, d, carried, acc, out, n_groups);
    spectra_commit_kernel<<<(N + 255) / 256, 256, 0, stream>>>(acc, N);
    GSDR_CUDA_OK(cudaGetLastError());
    return 2;
}

// One launch for the carry-over of every stream of a group.  `scratch` must hold n TailJob (device).
int window_tail_copy_multi(const Window* wins, const long long* n_tail, float2* const* dst, int n, void* scratch,
                           cudaStream_t stream) {
    if (n <= 0) return 0;
    std::vector<TailJob> tj(n);
    long long longest = 0;
    for (int i = 0; i < n; ++i) {
        tj[i] = TailJob{wins[i], wins[i].n_hist + wins[i].n_in - n_tail[i], n_tail[i], dst[i]};
        if (n_tail[i] > longest) longest = n_tail[i];
    }
    if (longest <= 0) return 0;
    GSDR_CUDA_OK(cudaMemcpyAsync(scratch, tj.data(), sizeof(TailJob) * n, cudaMemcpyHostToDevice, stream));
    int bx = (int)((longest + 255) / 256);
    if (bx > 64) bx = 64;
    window_tail_multi_kernel<<<dim3(bx, n), 256, 0, stream>>>(static_cast<const TailJob*>(scratch));
    GSDR_CUDA_OK(cudaGetLastError());
    return 1;
}
size_t window_tail_multi_scratch_bytes(int n) { return sizeof(TailJob) * (size_t)n; }

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
