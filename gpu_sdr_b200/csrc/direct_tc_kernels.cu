// DIRECT-mode demodulator (and generic-size channelizer) on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same function as direct_fir_tiled_kernel (direct_kernels.cu), i.e. the reference's
// direct_demodulator_integer + T x FIR::run_fir + cublasCgeam (cpp/kernels.cu:45-86, cpp/fir.cu:44-88,
// cpp/USRP_demodulator.cpp:400-464), as a GEMM.  The filter has F blocks of M taps and the decimation is M, so
//     y[p,t] = rot(p,t) * sum_{i<F} Z_i[p+i, t],      Z_i[r, t] = sum_{k<M} w[r M + k] g_t[i M + k]
// where w is the stream window cut into rows of M samples -- rows do NOT overlap, the window IS the A operand:
//     A [rows x 2M]      = the input samples, re/im interleaved, exactly as they lie in memory,
//     B [2M x F*TG*2]    = (g_r, g_i ; -g_i, g_r) per (FIR block i, tone t): complex multiply as a real GEMM,
//     D [rows x F*TG*2]  = all F partial filters Z_i of TG tones at once (N = 128 columns for every supported F).
// The F-term sum over neighbouring ROWS (the reference's overlap-add, cpp/fir.cu:55-69) is a lane shift in the
// epilogue, followed by the per-output LO rotation from the integer phase (cpp/kernels.cu:59-75).
//
// Precision: the tolerance is 1e-5 relative L2 against fp64 and one TF32 product has a 2^-11 relative error, so both
// operands are split x = hi + lo (hi = x rounded to TF32, lo = x - hi, exact) and three MMAs are issued per k-step:
// hi*hi into the `main` TMEM accumulator, lo*hi + hi*lo into a second one (`corr`: kept apart so that the small
// terms do not add truncating accumulation steps to the long chain); the dropped lo*lo term is 2^-22 relative.
//
// The tensor core's accumulate truncates (about 1.6e-7 relative per accumulated k-step when all terms have one sign), so
// a chain is cut after at most 32 k-steps (8 K blocks of 16 taps): each such SEGMENT is a separate accumulation and the
// epilogue adds the segments of a tile in fp32 -- any decimation runs here, not only decim <= 128.
//
// The same kernel with rotate = 0 is the polyphase channelizer for every channel count that has no fused FFT kernel: a
// PFB followed by a DFT of which T bins are kept is DIRECT with decim = N, F = P and taps w[m] e^{-2 pi j m bin / N}
// (rx.cu, init_pfb).
//
// What bounds the kernel is shared-memory bandwidth (128 B/clk/SM), which the operand fetch of tcgen05.mma shares
// with TMA writes and LDS/STS.  So the A operand (the samples) is fed from TENSOR MEMORY, written there by the
// producers with tcgen05.st straight from registers: it never exists in shared memory in operand form, and per
// 16-tap K block shared memory carries 16 KB (TMA landing) + 16 KB (one read of it) + 32 KB (B tiles written) +
// 48 KB (B read by the MMAs) instead of 176 KB.
//
// Warp roles (one persistent CTA per SM, static tile scheduler over (row tile, tone group)):
//   warps 0-3   epilogue: fold corr into main in TMEM (frees corr for the next tile), tcgen05.ld, row shift-and-add,
//               LO rotation, sample-major store
//   warp  4     TMEM allocation + single-thread tcgen05.mma issue (kind::tf32, M=128, N=128, K=8, A from TMEM)
//   warps 5-12  operand producers: thread = one window row (TMEM lane) and one half of its K block; 4 x LDS.128 of the
//               row from the TMA landing slot, hi/lo split, tcgen05.st; the tone filters come through registers (L2
//               hits) and are split into the 128-byte-swizzled K-major B tiles; two operand stages
//   warp  13    TMA issue: one cp.async.bulk.tensor (128 rows x 32 floats, SWIZZLE_128B) per K block into a ring of
//               four landing slots, each box prefetched into L2 eight K blocks earlier
// TMEM columns: main x 2 (the epilogue of tile n overlaps the MMAs of tile n+1), corr x 1, A (hi, lo) x 2 stages.
// From the second K block of a tile on, A_hi [B_hi | B_lo] is ONE N = 256 instruction into [main | corr] (the B tiles
// are laid out in the order of the accumulators), so a K block is 8 MMAs, not 12: the issuing thread spends ~60 cycles
// per instruction, about what a 128 x 128 x 8 MMA takes to execute.
#include <cuda.h>

#include <cstdlib>

#include "devmath.cuh"
#include "direct_common.cuh"

namespace gsdr {
namespace {

constexpr int TC_ROWS = 128;                       // window rows per tile = UMMA M
constexpr int TC_N = 128;                          // accumulator columns = F * TG * 2
constexpr int TC_KC = 16;                          // complex taps per K block: 32 floats = one 128-byte swizzle row
constexpr int TC_RAW = 4;                          // TMA landing slots
constexpr int TC_GROUPS = 2;                       // producer groups == operand stages, alternate K blocks
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_GROUP_WARPS = 4;                  // one warp per TMEM lane quarter
constexpr int TC_GROUP_THREADS = 32 * TC_GROUP_WARPS;
constexpr int TC_LOAD_WARPS = TC_GROUPS * TC_GROUP_WARPS;
constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 1 + TC_LOAD_WARPS + 1);
constexpr int TC_TMA_WARP = TC_EPI_WARPS + 1 + TC_LOAD_WARPS;
constexpr int TC_RAW_BYTES = TC_ROWS * 128;        // one landing slot: 128 rows x 128 bytes
constexpr int TC_B_BYTES = 2 * TC_N * 128;         // B_hi and B_lo tiles of one stage (128 rows x 128 bytes each)
constexpr int TC_XCH_FLOATS = 7 * 7 * 16;          // per epilogue warp: (F-1) lanes x (F-1) blocks x chunk, F <= 8
constexpr int TC_TMEM_COLS = 512;
constexpr unsigned int TC_COL_CORR = 128;          // main(0) at 0, corr at 128, main(1) at 256: [main(0) | corr] and [corr | main(1)] are
                                                   // both 256 contiguous columns, the D operand of one N = 256 MMA
constexpr unsigned int TC_COL_A = 384;             // A stage g: hi at +64 g, lo at +64 g + 32
constexpr int TC_L2_AHEAD = 8;                     // K blocks between the L2 prefetch of a box and its TMA load
constexpr int TC_B_PER_THREAD = (TC_N / 2) * TC_KC / (32 * TC_LOAD_WARPS);   // 4 filter taps
constexpr int TC_SEG_MAX = 8;                      // K blocks per accumulation segment: at most 32 k-steps per tensor-core chain
constexpr int TC_HIST_MAX = 7 * 128;               // carried-over samples kept in shared memory (the rest is read from global)
static_assert((TC_EPI_WARPS + 1) % 4 == 1 && TC_GROUP_WARPS == 4, "producer warp w owns TMEM lane quarter w % 4");

struct TcShared {
    unsigned long long raw_full[TC_RAW];     // TMA -> producers: the window rows of a K block have landed
    unsigned long long raw_empty[TC_RAW];    // producers -> TMA
    unsigned long long full[TC_GROUPS];      // producers -> MMA: A in TMEM and B in shared memory are ready
    unsigned long long empty[TC_GROUPS];     // MMA (tcgen05.commit) -> producers
    unsigned long long tmem_full[2];         // MMA -> epilogue: main(as) and corr hold the tile
    unsigned long long tmem_empty[2];        // epilogue -> MMA: main(as) has been read
    unsigned long long corr_empty;           // epilogue -> MMA: corr has been folded into main
    unsigned int tmem_base;
    unsigned int pad;
    double2 ph[64];          // per tone of the group: (LO phase of the tile's first row, phase step per row), integers < rate
    alignas(16) float xch[TC_EPI_WARPS][TC_XCH_FLOATS];
    alignas(16) float2 hist[TC_HIST_MAX];    // copy of the carried-over samples (rows the TMA tensor does not hold)
    float total[TC_N * TC_ROWS];             // [column][row]: sum of a tile's finished segments (decim > 128 only)
};
constexpr size_t TC_SMEM_BYTES = 1024 + (size_t)TC_RAW * TC_RAW_BYTES + (size_t)TC_GROUPS * TC_B_BYTES + sizeof(TcShared);

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned int addr, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned int addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// Bounded wait: a protocol error traps (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned int addr, unsigned int parity) {
    long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
        unsigned int done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins == 64) t0 = clock64();
        if (spins > 64 && (spins & 255u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned int addr, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// 2-D tiled TMA load: box (32 floats x 128 rows) at (c0 = float column, c1 = row), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(unsigned int dst, const CUtensorMap* map, int c0, int c1, unsigned int mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<unsigned long long>(map)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
// debug builds of the schedule: cycles spent in a wait, accumulated per role (dbg == nullptr in production)
__device__ __forceinline__ void mbar_wait_t(unsigned int addr, unsigned int parity, long long& acc, bool on) {
    if (!on) {
        mbar_wait(addr, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(addr, parity);
    acc += clock64() - t0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned int mbar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
}
// cute::UMMA::InstrDescriptor: D = f32 (1 @4), A = B = tf32 (2 @7, 2 @10), K-major both, N>>3 @17, M>>4 @24
constexpr unsigned int tc_idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(TC_ROWS >> 4) << 24); }
constexpr unsigned int TC_IDESC = tc_idesc(TC_N);
constexpr unsigned int TC_IDESC_WIDE = tc_idesc(2 * TC_N);

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns of tf32.  The accumulate flag is a compile-time constant
// (the predicate folds away): the single issuing thread is the kernel's critical path, ~10 instructions per MMA count.
template <bool kAccumulate>
__device__ __forceinline__ void tc_mma_tf32_ts(unsigned int d_tmem, unsigned int a_tmem, unsigned long long b_desc, unsigned int idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(kAccumulate ? 1 : 0)
        : "memory");
}
// one k-step, wide form: [main | corr] += A_hi [B_hi | B_lo] (d_wide / b_wide in the accumulators' order); corr += A_lo B_hi
__device__ __forceinline__ void tc_kstep_wide(unsigned int d_wide, unsigned int d_corr, unsigned int a_hi, unsigned long long b_wide,
                                              unsigned long long b_hi, int ks) {
    const unsigned long long adv = (unsigned long long)(2 * ks);
    tc_mma_tf32_ts<true>(d_wide, a_hi + 8u * ks, b_wide + adv, TC_IDESC_WIDE);
    tc_mma_tf32_ts<true>(d_corr, a_hi + 32u + 8u * ks, b_hi + adv, TC_IDESC);
}
// one k-step: main (+)= A_hi B_hi; small (+)= A_lo B_hi; small += A_hi B_lo
template <bool kAccMain, bool kAccSmall>
__device__ __forceinline__ void tc_kstep(unsigned int d_main, unsigned int d_small, unsigned int a_hi, unsigned long long b_hi,
                                         unsigned long long b_lo, int ks) {
    const unsigned long long adv = (unsigned long long)(2 * ks);   // 32 bytes >> 4 per k-step
    tc_mma_tf32_ts<kAccMain>(d_main, a_hi + 8u * ks, b_hi + adv, TC_IDESC);
    tc_mma_tf32_ts<kAccSmall>(d_small, a_hi + 32u + 8u * ks, b_hi + adv, TC_IDESC);
    tc_mma_tf32_ts<true>(d_small, a_hi + 8u * ks, b_lo + adv, TC_IDESC);
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) = 1, SBO>>4 [32,46) = 64, version [46,48) = 1,
// layout [61,64) = 2 (SWIZZLE_128B)).
__device__ __forceinline__ unsigned long long tc_smem_desc(unsigned int saddr) {
    const unsigned int lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const unsigned int hi = 64u | (1u << 14) | (2u << 29);
    return ((unsigned long long)hi << 32) | lo;
}
template <int CW>
__device__ __forceinline__ void tmem_ld(unsigned int taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld<32>(unsigned int taddr, float* v) {
    unsigned int* r = reinterpret_cast<unsigned int*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(unsigned int taddr, float* v) {
    unsigned int* r = reinterpret_cast<unsigned int*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned int taddr, const float* v) {
    const unsigned int* r = reinterpret_cast<const unsigned int*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st32(unsigned int taddr, const float* v) {
    tmem_st16(taddr, v);
    tmem_st16(taddr + 16u, v + 16);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory"); }

// x = hi + lo with hi = x rounded to TF32 (10 explicit mantissa bits, half away from zero: one integer add and a
// mask on the sign-magnitude bits); lo = x - hi is exact in fp32.  The tensor core drops the low 13 bits of lo, an
// error of at most 2^-21 |x| with the sign of lo, i.e. unbiased.  (cvt.rna.tf32.f32 costs six instructions on sm_100a.)
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// byte offset of (row, complex tap kc) inside a 128-row K-major SWIZZLE_128B operand tile
__device__ __forceinline__ unsigned int tc_swz(unsigned int row, unsigned int kc) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((((kc >> 1) ^ row) & 7u) << 4) + ((kc & 1u) << 3);
}

struct TcTile {
    long long row0;   // first window row of the tile
    int ch0;          // first tone of the group
};

// F = FIR blocks (pf_average); TG = 64 / F tones per group; outputs per tile = 128 - (F - 1)
template <int F, bool rotate>
__global__ void __launch_bounds__(TC_THREADS, 1)
direct_fir_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int use_tma, const long long hist_rows,
                     const Window w, const float2* __restrict__ g, const int* __restrict__ freq, int T, int M, int rate,
                     long long pos0, long long n_out, int n_row_tiles, int n_tone_groups, const int seg_shift, float2* __restrict__ out,
                     long long* __restrict__ dbg) {
    constexpr int TG = 64 / F;
    constexpr int RB = TC_ROWS - (F - 1);
    constexpr int TCW = TG < 16 ? TG : 16;     // tones per epilogue chunk
    constexpr int CW = 2 * TCW;                // accumulator columns per chunk and FIR block
    constexpr int NCHUNK = TG / TCW;

    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the address space: LDS/STS, not generic)
    unsigned char* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char* smem_b = smem + (size_t)TC_RAW * TC_RAW_BYTES;
    TcShared* sh = reinterpret_cast<TcShared*>(smem_b + (size_t)TC_GROUPS * TC_B_BYTES);
    const unsigned int smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntaps = F * M;
    const int KB = (M + TC_KC - 1) / TC_KC;           // K blocks per tile
    const int ksteps_total = (M + 3) / 4;             // MMA k-steps (4 complex taps = 8 tf32) per tile
    const int TC_SEG = 1 << seg_shift;                // K blocks per accumulation segment (1, 2, 4 or 8)
    const int NSEG = (KB + TC_SEG - 1) >> seg_shift;  // accumulation segments per tile (1 for decim <= 128 at the default 8)
    const int n_tiles = n_row_tiles * n_tone_groups;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const bool timed = dbg != nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_RAW; ++s) {
            mbar_init(smem_u32(&sh->raw_full[s]), 1);
            mbar_init(smem_u32(&sh->raw_empty[s]), 32 * TC_LOAD_WARPS);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&sh->full[a]), 32 * TC_LOAD_WARPS);
            mbar_init(smem_u32(&sh->empty[a]), 1);
            mbar_init(smem_u32(&sh->tmem_full[a]), 1);
            mbar_init(smem_u32(&sh->tmem_empty[a]), 32 * TC_EPI_WARPS);
        }
        mbar_init(smem_u32(&sh->corr_empty), 32 * TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the carried-over samples: the first `hist_rows` window rows hold samples of the previous call; the TMA tensor
    // starts at the first row that lies entirely in the `in` segment

    if ((int)blockIdx.x < n_tone_groups)   // only the CTAs that own a tile of the first row tile meet history rows
        for (int i = threadIdx.x; i < TC_HIST_MAX && i < w.n_hist; i += TC_THREADS) sh->hist[i] = w.hist[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned int tmem_base = sh->tmem_base;
    if (warp < TC_EPI_WARPS) {   // corr starts at zero: tiles only ever accumulate into it
        float z[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) z[e] = 0.f;
        for (int j = 0; j < TC_N / 32; ++j) tmem_st32(tmem_base + ((unsigned int)(warp * 32) << 16) + TC_COL_CORR + 32u * j, z);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    auto tile_of = [&](int n) {
        const int id = (int)blockIdx.x + n * (int)gridDim.x;
        TcTile t;
        t.row0 = (long long)(id / n_tone_groups) * RB;
        t.ch0 = (id % n_tone_groups) * TG;
        return t;
    };

    if (warp < TC_EPI_WARPS) {
        // ======================================= EPILOGUE =======================================
        const int row_in_tile = warp * 32 + lane;
        const double row_d = (double)row_in_tile;
        const double word_per_phase = 4294967296.0 / (double)rate;
        float* xw = sh->xch[warp];
        const float* xn = sh->xch[(warp + 1) & 3];
        long long w_acc0 = 0, e_fold = 0, e_gather = 0, e_rot = 0;
        const long long t_role0 = clock64();
        for (int n = 0; n < my_tiles; ++n) {
            const TcTile tl = tile_of(n);
            // LO phase (cpp/kernels.cu:59-75) of the tile's first output for each tone of the group, and its step per
            // output, as exact integers; row r of the tile then has phase base + r * step < 128 * rate < 2^53, formed
            // exactly by one DFMA.
            if (rotate && threadIdx.x < TG) {
                const int ch = tl.ch0 + (int)threadIdx.x;
                const double2 bs = ch < T ? lo_phase_tile(freq[ch], rate, pos0, tl.row0, M) : make_double2(0.0, 0.0);
                sh->ph[threadIdx.x] = bs;
            }
            // The tensor core's accumulate truncates, so a chain is cut after TC_SEG K blocks (at most 32 k-steps): every segment
            // is a separate accumulation, handed over like a tile (as = segment counter & 1), and the segments of a tile
            // are added here in fp32 (round to nearest), through a [column][row] array in shared memory.
            unsigned int as = 0, t_main = 0;
            for (int sgi = 0; sgi < NSEG; ++sgi) {
                const int c = n * NSEG + sgi;
                as = c & 1;
                mbar_wait_t(smem_u32(&sh->tmem_full[as]), (unsigned)(c >> 1) & 1u, w_acc0, timed);
                tc_fence_after();
                t_main = tmem_base + ((unsigned int)(warp * 32) << 16) + as * 256u;
                const long long te0 = timed ? clock64() : 0;
                const bool first = sgi == 0, last = sgi == NSEG - 1;
                if (KB > 1) {
                    // main += corr (+ earlier segments), in place, and corr = 0: corr is free for the next segment / tile
                    // long before this one has been written out
                    const unsigned int t_corr = tmem_base + ((unsigned int)(warp * 32) << 16) + TC_COL_CORR;
#pragma unroll 1
                    for (int j = 0; j < TC_N / 32; ++j) {
                        float a[32], b[32];
                        tmem_ld<32>(t_main + 32u * j, a);
                        tmem_ld<32>(t_corr + 32u * j, b);
                        tmem_ld_wait();
                        float* tot = sh->total + (32 * j) * TC_ROWS + row_in_tile;
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            a[e] += b[e], b[e] = 0.f;
                            if (!first) a[e] += tot[e * TC_ROWS];
                            if (!last) tot[e * TC_ROWS] = a[e];
                        }
                        if (last) tmem_st32(t_main + 32u * j, a);
                        tmem_st32(t_corr + 32u * j, b);   // the next segment accumulates into corr from its second K block on
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sh->corr_empty));
                }
                if (timed) e_fold += clock64() - te0;
                if (!last) {   // main(as) has been read: the MMA warp may start segment c + 2 in it
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sh->tmem_empty[as]));
                }
            }
            const long long p = tl.row0 + row_in_tile;
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                const long long tc0 = timed ? clock64() : 0;
                float y[CW], za[CW];
                tmem_ld<CW>(t_main + (unsigned)(c * CW), y);
                tmem_ld_wait();
#pragma unroll
                for (int i = 1; i < F; ++i) {
                    const unsigned int col = (unsigned)((i * TG + c * TCW) * 2);
                    tmem_ld<CW>(t_main + col, za);
                    tmem_ld_wait();
                    if (lane < F - 1) {
                        float4* dst = reinterpret_cast<float4*>(xw + (lane * (F - 1) + (i - 1)) * CW);
#pragma unroll
                        for (int j = 0; j < CW / 4; ++j) dst[j] = make_float4(za[4 * j], za[4 * j + 1], za[4 * j + 2], za[4 * j + 3]);
                    }
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const float v = __shfl_down_sync(0xffffffffu, za[j], i);
                        if (lane + i < 32) y[j] += v;
                    }
                }
                // every accumulator column of the tile is in registers: the MMA warp may start tile n + 2 in this stage
                if (c == NCHUNK - 1) {
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sh->tmem_empty[as]));
                }
                if (F > 1) {
                    epi_bar();
#pragma unroll
                    for (int i = 1; i < F; ++i) {
                        if (lane + i >= 32) {   // rows of the next warp (meaningless for the last warp: those outputs belong to the next tile)
                            const float4* src = reinterpret_cast<const float4*>(xn + ((lane + i - 32) * (F - 1) + (i - 1)) * CW);
#pragma unroll
                            for (int j = 0; j < CW / 4; ++j) {
                                const float4 v = src[j];
                                y[4 * j] += v.x, y[4 * j + 1] += v.y, y[4 * j + 2] += v.z, y[4 * j + 3] += v.w;
                            }
                        }
                    }
                } else if (c == 0) {
                    epi_bar();   // publishes sh->ph
                }
                const long long tc1 = timed ? clock64() : 0;
                e_gather += tc1 - tc0;
                if (row_in_tile < RB && p < n_out) {
                    // phase word = round(phase * 2^32 / rate) mod 2^32: the low mantissa bits of phase * c + 1.5 * 2^52
                    float2 o[TCW];
#pragma unroll
                    for (int t = 0; t < TCW; ++t) {
                        o[t] = make_float2(y[2 * t], y[2 * t + 1]);
                        if (rotate) {   // the channelizer form (pfb as GEMM) has no LO: whole turns per row
                            const unsigned int word = lo_phase_word(sh->ph[c * TCW + t], row_d, word_per_phase);
                            float sn, cs;
                            sincos_phase32(word, sn, cs);
                            o[t] = dev_cmul(o[t], make_float2(cs, -sn));
                        }
                    }
                    float2* dst = out + p * T + tl.ch0 + c * TCW;
                    const int n_valid = T - (tl.ch0 + c * TCW);   // tones of this chunk that exist
                    if (n_valid >= TCW && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                        for (int t = 0; t < TCW; t += 2) *reinterpret_cast<float4*>(dst + t) = make_float4(o[t].x, o[t].y, o[t + 1].x, o[t + 1].y);
                    } else {
#pragma unroll
                        for (int t = 0; t < TCW; ++t)
                            if (t < n_valid) dst[t] = o[t];
                    }
                }
                if (timed) e_rot += clock64() - tc1;
                if (F > 1) epi_bar();   // the exchange buffer (and sh->ph after the last chunk) may be rewritten
            }
            if (F == 1) epi_bar();
        }
        if (dbg && threadIdx.x == 0) {
            atomicAdd((unsigned long long*)&dbg[0], (unsigned long long)w_acc0);                      // epilogue: wait for accumulators
            atomicAdd((unsigned long long*)&dbg[1], (unsigned long long)(clock64() - t_role0));       // epilogue: role time
            atomicAdd((unsigned long long*)&dbg[11], (unsigned long long)e_fold);                     // epilogue: corr fold
            atomicAdd((unsigned long long*)&dbg[12], (unsigned long long)e_gather);                   // epilogue: TMEM loads, row shift, exchange
            atomicAdd((unsigned long long*)&dbg[13], (unsigned long long)e_rot);                      // epilogue: LO rotation + store
        }
    } else if (warp == TC_EPI_WARPS) {
        // ======================================= MMA ISSUE =======================================
        // Stage g = iteration & 1 (the producer groups alternate).  The first K block of a tile sends its two
        // small products into `main` (the chain is short and small there), so that the epilogue has a whole K
        // block of time to fold the previous tile's corr before this tile overwrites it.
        int it = 0;
        long long w_acc0 = 0, w_acc1 = 0, w_issue = 0;
        const long long t_role0 = clock64();
        int c = 0;   // accumulation segment counter: accumulator stage and barrier parities follow it
        for (int n = 0; n < my_tiles; ++n) {
            for (int sgi = 0; sgi < NSEG; ++sgi, ++c) {
                const unsigned int as = c & 1;
                mbar_wait_t(smem_u32(&sh->tmem_empty[as]), ((unsigned)(c >> 1) & 1u) ^ 1u, w_acc1, timed);
                tc_fence_after();
                const unsigned int d_main = tmem_base + as * 256u, d_corr = tmem_base + TC_COL_CORR;
                const unsigned int d_wide = tmem_base + as * 128u;   // [main(0) | corr] or [corr | main(1)]
                const int kb0 = sgi << seg_shift, kb1 = min(KB, kb0 + TC_SEG);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int gst = it & 1;
                    mbar_wait_t(smem_u32(&sh->full[gst]), (unsigned)(it >> 1) & 1u, w_acc0, timed);
                    if (kb == kb0 + 1) mbar_wait_t(smem_u32(&sh->corr_empty), (unsigned)(c & 1) ^ 1u, w_acc1, timed);
                    tc_fence_after();
                    if (lane == 0) {
                        const long long ti0 = timed ? clock64() : 0;
                        const unsigned int a_hi = tmem_base + TC_COL_A + 64u * gst;
                        const unsigned int b0 = smem_u32(smem_b) + (unsigned)gst * TC_B_BYTES;
                        // B tiles of the stage lie in the order of the accumulators: [B_hi | B_lo] for as = 0, [B_lo | B_hi] for as = 1
                        const unsigned long long b_x = tc_smem_desc(b0), b_y = tc_smem_desc(b0 + TC_B_BYTES / 2);
                        const unsigned long long b_hi = as ? b_y : b_x, b_lo = as ? b_x : b_y;
                        const int ks_n = min(4, ksteps_total - 4 * kb);
                        if (kb == kb0) {        // first K block of a segment: every product into main, which the first MMA initialises
                            tc_kstep<false, true>(d_main, d_main, a_hi, b_hi, b_lo, 0);
                            for (int ks = 1; ks < ks_n; ++ks) tc_kstep<true, true>(d_main, d_main, a_hi, b_hi, b_lo, ks);
                        } else if (ks_n == 4) { // corr was zeroed by the epilogue's fold
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) tc_kstep_wide(d_wide, d_corr, a_hi, b_x, b_hi, ks);
                        } else {
                            for (int ks = 0; ks < ks_n; ++ks) tc_kstep_wide(d_wide, d_corr, a_hi, b_x, b_hi, ks);
                        }
                        tc_commit(smem_u32(&sh->empty[gst]));
                        if (kb == kb1 - 1) tc_commit(smem_u32(&sh->tmem_full[as]));
                        if (timed) w_issue += clock64() - ti0;
                    }
                    __syncwarp();
                }
            }
        }
        if (dbg && lane == 0) {
            atomicAdd((unsigned long long*)&dbg[2], (unsigned long long)w_acc0);                      // MMA: wait for operands
            atomicAdd((unsigned long long*)&dbg[3], (unsigned long long)w_acc1);                      // MMA: wait for a free accumulator
            atomicAdd((unsigned long long*)&dbg[4], (unsigned long long)(clock64() - t_role0));       // MMA: role time
            atomicAdd((unsigned long long*)&dbg[10], (unsigned long long)w_issue);                    // MMA: inside the issue block
        }
    } else if (warp < TC_TMA_WARP) {
        // ======================================= OPERAND PRODUCERS =======================================
        // Eight warps, every K block: warp w owns TMEM lane quarter w % 4 (thread = one window row); the two warps of a
        // quarter take one half (16 floats) of the row's K block each.  Stage = iteration & 1.
        const int half = (warp - (TC_EPI_WARPS + 1)) / TC_GROUP_WARPS;    // which 16 floats of the 32-float K block
        const int q = warp & 3;                                           // TMEM lane quarter this warp may access
        const int row = 32 * q + lane;                                    // window row of the tile == TMEM lane
        const int lt = (int)threadIdx.x - 32 * (TC_EPI_WARPS + 1);        // 0..255
        const unsigned int kc = (unsigned)lt & 15u, sub = (unsigned)lt >> 4;   // tap inside the K block, (block,tone) slot
        const unsigned int b_off_re = tc_swz(2u * sub, kc);         // + u * 4096   (row n = 2*(u*16 + sub) + {0,1})
        const unsigned int b_off_im = tc_swz(2u * sub + 1u, kc);
        const unsigned int a_rowoff = (unsigned)(row >> 3) * 1024u + (unsigned)(row & 7) * 128u;
        const unsigned int a_tmem0 = tmem_base + ((unsigned int)(32 * q) << 16) + TC_COL_A + 16u * half;
        // (block i, tone t) of this thread's four filter taps: slot = u * 16 + sub, i = slot / TG, t = slot % TG
        int g_off[TC_B_PER_THREAD], g_t[TC_B_PER_THREAD];
#pragma unroll
        for (int u = 0; u < TC_B_PER_THREAD; ++u) {
            const int slot = u * 16 + (int)sub;
            g_t[u] = slot % TG;
            g_off[u] = (slot / TG) * M + g_t[u] * ntaps;
        }
        auto issue_loads = [&](const TcTile& tl, int kb, float2* vb) {
            const int k = kb * TC_KC + (int)kc;
            const float2* gk = g + (long long)tl.ch0 * ntaps + k;
#pragma unroll
            for (int u = 0; u < TC_B_PER_THREAD; ++u) {
                vb[u] = make_float2(0.f, 0.f);
                if (k < M && tl.ch0 + g_t[u] < T) vb[u] = __ldg(gk + g_off[u]);
            }
        };

        int n_cur = 0, kb_cur = 0, it = 0;
        TcTile t_cur = tile_of(0);
        long long w_acc0 = 0, p_a = 0, p_b = 0, p_stw = 0, p_fence = 0;
        const long long t_role0 = clock64();
        // One K block: prefetch the next block's filter taps into `nb`, convert this block from `cb`.  Called with the
        // two buffers swapped on alternate iterations: a register copy of `nb` would wait for the loads to land.
        auto k_block = [&](float2* cb, float2* nb) {
            int n_nxt = n_cur, kb_nxt = kb_cur + 1;
            if (kb_nxt == KB) kb_nxt = 0, ++n_nxt;
            const TcTile t_nxt = (n_nxt == n_cur) ? t_cur : tile_of(n_nxt);
            if (n_nxt < my_tiles) issue_loads(t_nxt, kb_nxt, nb);
            const int st = it & 1, r = it & (TC_RAW - 1);
            mbar_wait_t(smem_u32(&sh->empty[st]), ((unsigned)(it >> 1) & 1u) ^ 1u, w_acc0, timed);   // the MMAs of this stage's previous use are done
            mbar_wait_t(smem_u32(&sh->raw_full[r]), (unsigned)(it >> 2) & 1u, w_acc0, timed);
            tc_fence_after();
            const long long tp0 = timed ? clock64() : 0;
            const unsigned char* raw = smem + (size_t)r * TC_RAW_BYTES;
            const bool from_smem = use_tma && t_cur.row0 + row >= hist_rows;   // else: history row, or no TMA at all
            const long long s_row = (t_cur.row0 + row) * (long long)M + kb_cur * TC_KC + 8 * half;
            {
                float v[16], hi[16], lo[16];
                if (from_smem) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 x = *reinterpret_cast<const float4*>(raw + a_rowoff + ((unsigned)((4 * half + c4) ^ (row & 7)) << 4));
                        v[4 * c4] = x.x, v[4 * c4 + 1] = x.y, v[4 * c4 + 2] = x.z, v[4 * c4 + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = kb_cur * TC_KC + 8 * half + j;
                        float2 a = make_float2(0.f, 0.f);
                        if (k < M) a = (use_tma && s_row + j < w.n_hist && s_row + j < TC_HIST_MAX) ? sh->hist[s_row + j] : dev_win_at(w, s_row + j);
                        v[2 * j] = a.x, v[2 * j + 1] = a.y;
                    }
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) tf32_split(v[e], hi[e], lo[e]);
                __syncwarp();   // tcgen05.st is warp-collective: reconverge after the per-lane source selection
                tmem_st16(a_tmem0 + 64u * st, hi);
                tmem_st16(a_tmem0 + 64u * st + 32u, lo);
            }
            const int b_hi_off = ((n_cur * NSEG + (kb_cur >> seg_shift)) & 1) ? TC_B_BYTES / 2 : 0;   // accumulator order of this K block's segment: [main(0) | corr], [corr | main(1)]
            const long long tp1 = timed ? clock64() : 0;
            mbar_arrive(smem_u32(&sh->raw_empty[r]));   // this thread's reads of the landing slot are complete (values consumed)
            unsigned char* const bst = smem_b + (size_t)st * TC_B_BYTES;
#pragma unroll
            for (int u = 0; u < TC_B_PER_THREAD; ++u) {
                float2 hi, lo;
                tf32_split(cb[u].x, hi.x, lo.x);
                tf32_split(cb[u].y, hi.y, lo.y);
                unsigned char* bh = bst + b_hi_off + u * 4096;
                unsigned char* bl = bst + (TC_B_BYTES / 2 - b_hi_off) + u * 4096;
                *reinterpret_cast<float2*>(bh + b_off_re) = make_float2(hi.x, -hi.y);   // Re(x g): x_r g_r - x_i g_i
                *reinterpret_cast<float2*>(bh + b_off_im) = make_float2(hi.y, hi.x);    // Im(x g): x_r g_i + x_i g_r
                *reinterpret_cast<float2*>(bl + b_off_re) = make_float2(lo.x, -lo.y);
                *reinterpret_cast<float2*>(bl + b_off_im) = make_float2(lo.y, lo.x);
            }
            const long long tp2 = timed ? clock64() : 0;
            tmem_st_wait();
            const long long tp3 = timed ? clock64() : 0;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(smem_u32(&sh->full[st]));
            if (timed) p_a += tp1 - tp0, p_b += tp2 - tp1, p_stw += tp3 - tp2, p_fence += clock64() - tp3;
            ++it;
            n_cur = n_nxt, kb_cur = kb_nxt, t_cur = t_nxt;
        };
        float2 xb0[TC_B_PER_THREAD], xb1[TC_B_PER_THREAD];
        if (my_tiles > 0) issue_loads(t_cur, 0, xb0);
        while (n_cur < my_tiles) {
            k_block(xb0, xb1);
            if (n_cur >= my_tiles) break;
            k_block(xb1, xb0);
        }
        if (dbg && lt == 0) {
            atomicAdd((unsigned long long*)&dbg[5], (unsigned long long)w_acc0);                      // producers: wait for raw rows / a free stage
            atomicAdd((unsigned long long*)&dbg[6], (unsigned long long)(clock64() - t_role0));       // producers: role time
            atomicAdd((unsigned long long*)&dbg[14], (unsigned long long)p_a);
            atomicAdd((unsigned long long*)&dbg[15], (unsigned long long)p_b);
            atomicAdd((unsigned long long*)&dbg[16], (unsigned long long)p_stw);
            atomicAdd((unsigned long long*)&dbg[17], (unsigned long long)p_fence);
        }
    } else if (warp == TC_TMA_WARP) {
        // ======================================= TMA ISSUE =======================================
        // One lane.  Besides the load of the K block into its landing slot, the box TC_L2_AHEAD iterations further on
        // is prefetched into L2, so that the load proper finds its rows there.  History rows have negative tensor
        // coordinates: TMA zero-fills them and the producers take them from sh->hist.
        if (lane == 0) {
            int s = 0;
            unsigned int ph = 0;
            int n_pf = 0, kb_pf = 0;   // (tile, K block) of the next L2 prefetch
            long long w_acc0 = 0;
            const long long t_role0 = clock64();
            TcTile t_pf = tile_of(0);
            auto prefetch_next = [&]() {
                if (n_pf >= my_tiles) return;
                if (use_tma) tma_prefetch_2d(&tmap, kb_pf * 2 * TC_KC, (int)(t_pf.row0 - hist_rows));
                if (++kb_pf == KB) {
                    kb_pf = 0;
                    ++n_pf;
                    t_pf = tile_of(n_pf);
                }
            };
            for (int j = 0; j < TC_L2_AHEAD; ++j) prefetch_next();
            for (int n = 0; n < my_tiles; ++n) {
                const TcTile tl = tile_of(n);
                for (int kb = 0; kb < KB; ++kb) {
                    prefetch_next();
                    mbar_wait_t(smem_u32(&sh->raw_empty[s]), ph ^ 1u, w_acc0, timed);
                    const unsigned int bar = smem_u32(&sh->raw_full[s]);
                    if (use_tma) {
                        mbar_arrive_expect_tx(bar, TC_RAW_BYTES);
                        tma_load_2d(smem_base + (unsigned)s * TC_RAW_BYTES, &tmap, kb * 2 * TC_KC, (int)(tl.row0 - hist_rows), bar);
                    } else {
                        mbar_arrive(bar);
                    }
                    if (++s == TC_RAW) s = 0, ph ^= 1u;
                }
            }
            if (dbg) {
                atomicAdd((unsigned long long*)&dbg[7], (unsigned long long)w_acc0);                  // TMA: wait for a free slot
                atomicAdd((unsigned long long*)&dbg[8], (unsigned long long)(clock64() - t_role0));   // TMA: role time
                atomicAdd((unsigned long long*)&dbg[9], (unsigned long long)(my_tiles * KB));         // stage iterations
            }
        }
    }

    // teardown: every MMA has completed before the last tmem_full arrival, every tcgen05.ld before this barrier
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == TC_EPI_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime (libgsdr.so does not link libcuda)
TcEncodeTiledFn tc_encode_fn() {
    static TcEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TcEncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// The part of the `in` segment that starts on a window-row boundary as a 2-D tensor [rows][2 M floats] (row pitch M * 8
// bytes), box = 128 rows x 32 floats, 128-byte swizzle: exactly the K-major operand tile of one K block.  Needs a
// 16-byte-aligned base and pitch.  *hist_rows = window rows that hold carried-over samples (ceil(n_hist / M)).
bool tc_make_tensor_map(const Window& w, int M, CUtensorMap* map, long long* hist_rows, bool allow_tma) {
    const long long hr = (w.n_hist + M - 1) / M, off0 = hr * M - w.n_hist;
    *hist_rows = hr;
    if (!allow_tma || M < TC_KC || (M & 1) || w.n_in - off0 < M) return false;
    const float2* base = w.in + off0;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    TcEncodeTiledFn enc = tc_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)(2 * M), (cuuint64_t)((w.n_in - off0) / M)};
    const cuuint64_t strides[1] = {(cuuint64_t)M * 8};
    const cuuint32_t box[2] = {2 * TC_KC, TC_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int F>
int tc_launch(const Window& w, const float2* g, const int* freq_dev, int T, int M, int rate, long long pos0, long long n_out,
              float2* out, int sm_count, cudaStream_t stream, int rotate, bool allow_tma) {
    constexpr int TG = 64 / F, RB = TC_ROWS - (F - 1);
    const int row_tiles = (int)((n_out + RB - 1) / RB), tone_groups = (T + TG - 1) / TG;
    static DeviceOnce attr_once;
    if (const int dev = attr_once.pending(); dev >= 0) {
        GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_tc_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_tc_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        attr_once.done(dev);
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    long long hist_rows = 0;
    const int use_tma = tc_make_tensor_map(w, M, &map, &hist_rows, allow_tma) ? 1 : 0;
    const long long tiles = (long long)row_tiles * tone_groups;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    // GSDR_DIRECT_TC_DEBUG=1: per-role wait / run cycles of this launch on stderr (synchronises; schedule tuning only)
    const char* de = getenv("GSDR_DIRECT_TC_DEBUG");
    long long* dbg = nullptr;
    if (de && de[0] == '1') {
        GSDR_CUDA_OK(cudaMalloc(&dbg, 24 * sizeof(long long)));
        GSDR_CUDA_OK(cudaMemsetAsync(dbg, 0, 24 * sizeof(long long), stream));
    }
    // GSDR_DIRECT_TC_SEG = 1, 2, 4 or 8: K blocks (of 16 taps) per accumulation chain.  Shorter chains cost folds and the 12-MMA first
    // K block of every segment, and buy accuracy: the chain's truncating accumulate is the kernel's largest error term.
    int seg_shift = 3;   // 8 K blocks = TC_SEG_MAX
    if (const char* se = getenv("GSDR_DIRECT_TC_SEG")) {
        const int v = atoi(se);
        if (v == 1 || v == 2 || v == 4 || v == 8) seg_shift = v == 1 ? 0 : v == 2 ? 1 : v == 4 ? 2 : 3;
    }
    if (rotate)
        direct_fir_tc_kernel<F, true><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(map, use_tma, hist_rows, w, g, freq_dev, T, M, rate, pos0,
                                                                                   n_out, row_tiles, tone_groups, seg_shift, out, dbg);
    else
        direct_fir_tc_kernel<F, false><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(map, use_tma, hist_rows, w, g, freq_dev, T, M, rate, pos0,
                                                                                    n_out, row_tiles, tone_groups, seg_shift, out, dbg);
    GSDR_CUDA_OK(cudaGetLastError());
    if (dbg) {
        long long h[24];
        GSDR_CUDA_OK(cudaStreamSynchronize(stream));
        GSDR_CUDA_OK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        const double it = h[9] > 0 ? (double)h[9] : 1.0;
        fprintf(stderr,
                "[direct_tc] grid %d tiles %lld use_tma %d | per stage iteration (cycles): epi wait %.0f of %.0f | mma wait-operands %.0f "
                "wait-acc %.0f issue %.0f of %.0f | producers wait %.0f of %.0f | tma wait %.0f of %.0f | epi fold %.0f gather %.0f rotate+store %.0f | prod A %.0f B %.0f st-wait %.0f fence+arrive %.0f\n",
                grid, tiles, use_tma, h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[10] / it, h[4] / it, h[5] / it, h[6] / it, h[7] / it, h[8] / it,
                h[11] / it, h[12] / it, h[13] / it, h[14] / it, h[15] / it, h[16] / it, h[17] / it);
    }
    return 1;
}

}  // namespace

// F in {1, 2, 4, 8} so that F * TG * 2 = 128 accumulator columns; the LO phase arithmetic needs 128 * rate < 2^53.
// Any decimation: the tensor-core accumulation truncates (measured on B200: about 1.6e-7 relative per accumulated k-step
// when every term has the same sign, 4e-5 for one chain over decim = 1000), so a chain is cut after 32 k-steps (128
// taps) and the segments are added in fp32 by the epilogue.
bool direct_fir_tc_supported(int T, int M, int ntaps, long long n_out) {
    if (M < 1 || T < 1 || n_out < 1 || ntaps % M != 0) return false;
    const int f = ntaps / M;
    if (!(f == 1 || f == 2 || f == 4 || f == 8)) return false;
    const long long rb = TC_ROWS - (f - 1), tg = 64 / f;
    const long long tiles = ((n_out + rb - 1) / rb) * ((T + tg - 1) / tg);
    return tiles < (1ll << 30);
}

// Worth it when there are enough (row tile, tone group) tiles to occupy a good part of the GPU; small problems
// (few outputs per buffer, i.e. very large decimation) stay on the fp32 kernel.
bool direct_fir_tc_preferred(int T, int M, int ntaps, long long n_out) {
    if (!direct_fir_tc_supported(T, M, ntaps, n_out)) return false;
    const int f = ntaps / M;
    const long long rb = TC_ROWS - (f - 1), tg = 64 / f;
    const long long tiles = ((n_out + rb - 1) / rb) * ((T + tg - 1) / tg);
    return tiles >= 64;
}

int direct_fir_tc_launch(const Window& w, const float2* g, const int* freq_dev, int T, int M, int ntaps, int rate, long long pos0,
                         long long n_out, float2* out, int sm_count, cudaStream_t stream, int rotate, bool allow_tma) {
    if (n_out <= 0) return 0;
    if (!direct_fir_tc_supported(T, M, ntaps, n_out)) {
        set_error("direct_fir_tc_launch: unsupported shape (T=%d M=%d ntaps=%d)", T, M, ntaps);
        return -1;
    }
    switch (ntaps / M) {
        case 1: return tc_launch<1>(w, g, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
        case 2: return tc_launch<2>(w, g, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
        case 4: return tc_launch<4>(w, g, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
        default: return tc_launch<8>(w, g, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
    }
}

}  // namespace gsdr
