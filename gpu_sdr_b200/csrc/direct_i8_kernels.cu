// DIRECT-mode demodulator (and generic-size channelizer) as an EXACT integer GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, int32 accumulators in TMEM), sm_100a only.
//
// Same function as direct_fir_tc_kernel (direct_tc_kernels.cu) and direct_fir_tiled_kernel (direct_kernels.cu), i.e. the
// reference's direct_demodulator_integer + T x FIR::run_fir + cublasCgeam (cpp/kernels.cu:45-86, cpp/fir.cu:44-88,
// cpp/USRP_demodulator.cpp:400-464):
//     y[p,t] = rot(p,t) * sum_{i<F} Z_i[p+i, t],      Z_i[r, t] = sum_{k<M} w[r M + k] g_t[i M + k]
// with the window cut into rows of M samples (A [rows x 2M], the samples as they lie in memory) and the per-tone complex
// filters as the real matrix B [2M x F*TG*2] = (g_r, g_i ; -g_i, g_r).
//
// Why integers.  The TF32 path (three split products, fp32 accumulators) is bound in accuracy by the tensor core's
// truncating accumulate: 5.1e-7 relative L2 against fp64 on cfg1, where the reference's cuBLAS chain has 1.2e-7
// (tools/direct_tc_emulation.py reproduces both).  Here both operands are put on a 24-bit fixed-point grid and cut into
// three signed 8-bit digits,
//     round(x * sA) = a1 2^16 + a2 2^8 + a3,     round(g * sB) = b1 2^16 + b2 2^8 + b3,     a_i, b_j in [-128, 127],
// and the six digit products of weight >= 2^16 are accumulated by kind::i8 MMAs in int32 -- exactly, whatever the chain
// length -- in three accumulators (weights 2^32, 2^24, 2^16).  The epilogue combines them in 64-bit integers, adds the F
// row-shifted blocks (the reference's overlap-add, cpp/fir.cu:55-69) still in integers, and converts ONCE to float.  What
// is left is the quantisation of the operands (2^-24 of the tile's largest sample, of each tone's largest tap) and two fp32
// roundings of the result.  Six int8 MMAs of K = 32 per 16-tap K block cost half the tensor time of the TF32 scheme's
// twelve, and the operands are 3 bytes per real instead of 8.
//
//   sB  per tone, fixed when the demodulator is created (the bank is built on the host from the double-precision taps and
//       lies in HBM as ready-made K-major tiles: the producers never touch the filters, TMA brings them);
//   sA  per tile (128 window rows), from the largest |sample| of the tile, found one tile ahead: the scan-issue warp lands the
//       NEXT tile's K-block boxes in two slots of their own (the same boxes the operand path asks for one tile later, by then
//       from L2: the window crosses HBM once) and the producers fold them into a running maximum between two K blocks.
//
// Warp roles (one persistent CTA of 20 warps per SM, static scheduler over (row tile, tone group)):
//   warps 0-3, 4-7    two epilogue teams, units of 8 tones dealt alternately: tcgen05.ld of the three accumulators, row
//                     shift-and-add in int32, 64-bit combine, int -> float, LO rotation from the integer phase
//                     (cpp/kernels.cu:59-75), sample-major store.  The accumulators are single-buffered (3 x 128 + 96 operand
//                     columns of the 512): the teams read them side by side so that the MMA warp gets them back early.
//   warps 8-15        operand producers: thread = one window row (TMEM lane) and one half of its K block: 4 x LDS.128 from the
//                     TMA landing slot, scale, round, digits, tcgen05.st -- the A operand only ever exists in tensor memory
//   warp  16          TMEM allocation + single-thread tcgen05.mma issue: per K block two N = 256 and two N = 128 MMAs
//   warp  17          TMA issue: one box (128 rows x 32 floats, SWIZZLE_128B) per K block for A, three boxes (128 x 128 bytes)
//                     per four K blocks for the digit planes of B
//   warp  18          scan-box issue (above); warp 19 idles
#include <cuda.h>

#include <cmath>
#include <cstdlib>

#include "devmath.cuh"
#include "direct_common.cuh"

namespace gsdr {
namespace {

constexpr int I8_ROWS = 128;                 // window rows per tile = UMMA M
constexpr int I8_N = 128;                    // accumulator columns = F * TG * 2
constexpr int I8_KC = 16;                    // complex taps per K block: 32 reals = one kind::i8 MMA (K = 32)
constexpr int I8_RAW = 4;                    // TMA landing slots for A
#ifndef GSDR_I8_AST
#define GSDR_I8_AST 4
#endif
constexpr int I8_AST = GSDR_I8_AST;          // A operand stages in TMEM (24 columns each; a fifth fits and changes nothing)
constexpr int I8_BST = 2;                    // B operand stages in shared memory (each: 3 digit planes x 4 K blocks)
// Twenty warps (96 registers each: ptxas keeps every role inside the launch allocation, with or without setmaxnreg):
//   warps 0-3, 4-7    two epilogue teams (units of 8 tones dealt alternately)
//   warps 8-11, 12-15 operand producers (first / second half of a K block)
//   warp 16 MMA issue, 17 TMA issue, 18 scan-box issue, 19 idle
constexpr int I8_EPI_TEAMS = 2;
constexpr int I8_EPI_WARPS = 4 * I8_EPI_TEAMS;
constexpr int I8_PROD_WARPS = 8;
constexpr int I8_PROD_WARP0 = I8_EPI_WARPS;
constexpr int I8_MMA_WARP = I8_PROD_WARP0 + I8_PROD_WARPS;
constexpr int I8_TMA_WARP = I8_MMA_WARP + 1;
constexpr int I8_SCAN_WARP = I8_MMA_WARP + 2;   // issues the scan boxes (see the producers)
constexpr int I8_THREADS = 32 * (I8_MMA_WARP + 4);
constexpr int I8_RAW_BYTES = I8_ROWS * 128;  // one landing slot: 128 rows x 128 bytes
constexpr int I8_B_PLANE = I8_N * 128;       // one digit plane: 128 columns (rows of the K-major tile) x 128 bytes (4 K blocks)
constexpr int I8_B_STAGE = 3 * I8_B_PLANE;
constexpr unsigned int I8_COL_A = 384;       // accumulators: weight 2^32 at column 0, 2^24 at 128, 2^16 at 256; A stage s at 384 + 24 s
constexpr int I8_TMEM_COLS = 512;
// exchange buffer per epilogue warp: (F-1) lanes x (F-1) blocks x 3 accumulators x 16 columns
__host__ __device__ constexpr int i8_xch_ints(int F) { return F > 1 ? (F - 1) * (F - 1) * 3 * 16 : 16; }
// landing slots (one A-sized box each) for the scan of the NEXT tile's largest sample; F = 8 has no room for them (its exchange
// buffer is 37 KB) and keeps the scan on global loads
// epilogue warps with work: with one unit per tile (F = 8) the second team idles
__host__ __device__ constexpr int i8_xch_warps(int F) { return 64 / F / 8 >= 2 ? 8 : 4; }
__host__ __device__ constexpr int i8_scan_slots(int F) { return F <= 4 ? 2 : 0; }
constexpr int I8_HIST_MAX = 7 * 128;
constexpr float I8_FULL_SCALE = 8355000.0f;  // |digits| <= 127 * (2^16 + 2^8 + 1) = 8355711
static_assert(I8_PROD_WARP0 % 4 == 0, "warp w owns TMEM lane quarter w % 4");
static_assert(I8_COL_A + 24 * I8_AST <= I8_TMEM_COLS && (I8_RAW & (I8_RAW - 1)) == 0, "tensor-memory columns / landing-slot ring");

template <int XCH, int XW>
struct I8SharedT {
    unsigned long long raw_full[I8_RAW], raw_empty[I8_RAW];   // TMA <-> producers: window rows of a K block
    unsigned long long a_full[I8_AST], a_empty[I8_AST];       // producers <-> MMA: A digits in TMEM
    unsigned long long b_full[I8_BST], b_empty[I8_BST];       // TMA <-> MMA: digit planes of B
    unsigned long long tmem_full, tmem_empty;                 // MMA <-> epilogue: the accumulators
    unsigned long long scan_full[4], scan_empty[4];           // scan-issue warp <-> producers: boxes of the next tile
    unsigned int tmem_base;
    unsigned int amax_bits[4];
    double2 ph[64];          // per tone of the group: (LO phase of the tile's first row, phase step per row), integers < rate
    float inv_sb[64];        // per tone of the group: 1 / sB
    alignas(16) int xch[XW][XCH];   // one per epilogue warp that has units to do
    alignas(16) float2 hist[I8_HIST_MAX];
};
constexpr size_t i8_smem_bytes(int F) {
    return 1024 + (size_t)(I8_RAW + i8_scan_slots(F)) * I8_RAW_BYTES + (size_t)I8_BST * I8_B_STAGE +
           (F == 1 ? sizeof(I8SharedT<i8_xch_ints(1), i8_xch_warps(1)>) : F == 2 ? sizeof(I8SharedT<i8_xch_ints(2), i8_xch_warps(2)>) :
            F == 4 ? sizeof(I8SharedT<i8_xch_ints(4), i8_xch_warps(4)>) : sizeof(I8SharedT<i8_xch_ints(8), i8_xch_warps(8)>));
}
static_assert(i8_smem_bytes(1) <= 232448 && i8_smem_bytes(2) <= 232448 && i8_smem_bytes(4) <= 232448 && i8_smem_bytes(8) <= 232448, "shared memory per CTA");

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned int addr, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned int addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// (a suspend-time hint on try_wait was measured: the waiting warps free issue slots but wake late, 232 -> 215 GS/s on cfg1)
// Bounded wait: a protocol error traps (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned int addr, unsigned int parity) {
    // A tight poll loop (the waiting warps share issue slots with the working ones: the round-2 profile showed 17 instructions per
    // poll, a fifth of everything issued); the clock is read once per 4096 polls only.
    long long t0 = 0;
    for (unsigned int rounds = 0;; ++rounds) {
        unsigned int done;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\tmov.u32 n, 4096;\n"
            "WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t@p bra DONE_%=;\n\tsub.u32 n, n, 1;\n\tsetp.ne.u32 q, n, 0;\n\t@q bra WAIT_%=;\n"
            "DONE_%=:\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (rounds == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// schedule tuning (GSDR_DIRECT_I8_DEBUG=1): cycles spent in a wait, accumulated per role
__device__ __forceinline__ void mbar_wait_t(unsigned int addr, unsigned int parity, long long& acc, bool on) {
    if (!on) {
        mbar_wait(addr, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(addr, parity);
    acc += clock64() - t0;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned int addr, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned int dst, const CUtensorMap* map, int c0, int c1, unsigned int mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned int mbar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
}
// cute::UMMA::InstrDescriptor: D = s32 (2 @4), A = B = signed int8 (1 @7, 1 @10), K-major both, N>>3 @17, M>>4 @24
constexpr unsigned int i8_idesc(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(I8_ROWS >> 4) << 24); }
constexpr unsigned int I8_IDESC = i8_idesc(I8_N);
constexpr unsigned int I8_IDESC_WIDE = i8_idesc(2 * I8_N);
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns (32 int8 per lane)
__device__ __forceinline__ void mma_i8_ts(unsigned int d_tmem, unsigned int a_tmem, unsigned long long b_desc, unsigned int idesc,
                                          unsigned int accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) = 1, SBO>>4 [32,46) = 64, version [46,48) = 1,
// layout [61,64) = 2 (SWIZZLE_128B)).
__device__ __forceinline__ unsigned long long i8_smem_desc(unsigned int saddr) {
    const unsigned int lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const unsigned int hi = 64u | (1u << 14) | (2u << 29);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ void tmem_ld16(unsigned int taddr, int* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st4(unsigned int taddr, unsigned int a, unsigned int b, unsigned int c, unsigned int d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int team) {   // one named barrier per epilogue team (2 belongs to the producers)
    if (team == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 3, 128;" ::: "memory");
}

// round_to_nearest_float(s1 2^16 + s2 2^8 + s3): the exact 64-bit sum, cut into a high part (arithmetic shift by 21: an int32
// here, |sum| < 2^52) and a 21-bit low part, both converted exactly; the one rounding is the FFMA's.
__device__ __forceinline__ float i8_combine(int s1, int s2, int s3) {
    const long long v = ((long long)s1 << 16) + ((long long)s2 << 8) + (long long)s3;
    const int hi = (int)(v >> 21);
    const int lo = (int)((unsigned int)v & 0x1FFFFFu);
    const float flo = __int_as_float(0x4B000000 | lo) - 8388608.0f;   // exact: lo < 2^21
    return fmaf((float)hi, 2097152.0f, flo);
}

struct I8Tile {
    long long row0;   // first window row of the tile
    int ch0;          // first tone of the group
    int tg;           // tone group index
};

// F = FIR blocks (pf_average); TG = 64 / F tones per group; outputs per tile = 128 - (F - 1)
template <int F, bool rotate>
__global__ void __launch_bounds__(I8_THREADS, 1)
direct_fir_i8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const int use_tma,
                     const long long hist_rows, const Window w, const float* __restrict__ inv_sb, const int* __restrict__ freq, int T,
                     int M, int rate, long long pos0, long long n_out, int n_row_tiles, int n_tone_groups, int KQ,
                     const unsigned int a_order, const float* __restrict__ amax_rows, float2* __restrict__ out, long long* __restrict__ dbg) {
    constexpr int TG = 64 / F;
    constexpr int RB = I8_ROWS - (F - 1);
    constexpr int NUNIT = TG / 8;              // epilogue units of 8 tones (16 accumulator columns) per tile
    constexpr int NSCAN = i8_scan_slots(F);
    using I8Shared = I8SharedT<i8_xch_ints(F), i8_xch_warps(F)>;
    static_assert(TG % 8 == 0 && NSCAN <= 4, "tile shape");

    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char* smem = i8_smem_raw + ((1024u - (smem_u32(i8_smem_raw) & 1023u)) & 1023u);
    unsigned char* smem_b = smem + (size_t)I8_RAW * I8_RAW_BYTES;
    unsigned char* smem_scan = smem_b + (size_t)I8_BST * I8_B_STAGE;
    I8Shared* sh = reinterpret_cast<I8Shared*>(smem_scan + (size_t)NSCAN * I8_RAW_BYTES);
    const unsigned int smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = (M + I8_KC - 1) / I8_KC;           // K blocks per tile
    const int n_tiles = n_row_tiles * n_tone_groups;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const bool timed = dbg != nullptr;
    long long wt0 = 0, wt1 = 0, wt2 = 0;
    const long long t_role0 = clock64();

    if (threadIdx.x == 0) {
        for (int s = 0; s < I8_RAW; ++s) {
            mbar_init(smem_u32(&sh->raw_full[s]), 1);
            mbar_init(smem_u32(&sh->raw_empty[s]), 32 * I8_PROD_WARPS);
        }
        for (int s = 0; s < I8_AST; ++s) {
            mbar_init(smem_u32(&sh->a_full[s]), 32 * I8_PROD_WARPS);
            mbar_init(smem_u32(&sh->a_empty[s]), 1);
        }
        for (int s = 0; s < I8_BST; ++s) {
            mbar_init(smem_u32(&sh->b_full[s]), 1);
            mbar_init(smem_u32(&sh->b_empty[s]), 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(smem_u32(&sh->scan_full[s]), 1);
            mbar_init(smem_u32(&sh->scan_empty[s]), 32 * I8_PROD_WARPS);
        }
        mbar_init(smem_u32(&sh->tmem_full), 1);
        mbar_init(smem_u32(&sh->tmem_empty), 32 * I8_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == I8_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                     "r"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if ((int)blockIdx.x < n_tone_groups)   // only the CTAs that own a tile of the first row tile meet history rows
        for (int i = threadIdx.x; i < I8_HIST_MAX && i < w.n_hist; i += I8_THREADS) sh->hist[i] = w.hist[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned int tmem_base = sh->tmem_base;

    auto tile_of = [&](int n) {
        const int id = (int)blockIdx.x + n * (int)gridDim.x;
        I8Tile t;
        t.row0 = (long long)(id / n_tone_groups) * RB;
        t.tg = id % n_tone_groups;
        t.ch0 = t.tg * TG;
        return t;
    };

    // Tile n + 1 of this CTA is scanned out of TMA-landed boxes (instead of with global loads) when every one of its rows lies in
    // the tensor map: the producers and the scan-issue warp take the same decision.
    auto scan_by_tma = [&](int n_next) -> bool {
        return NSCAN > 0 && use_tma && amax_rows == nullptr && n_next < my_tiles && tile_of(n_next).row0 >= hist_rows;
    };

    if (warp < I8_EPI_WARPS) {
        // ======================================= EPILOGUE =======================================
        const int team = warp >> 2, wq = warp & 3;            // team: units u = team, team + 2, ...; wq: TMEM lane quarter
        const int u_last = NUNIT - 1 - ((NUNIT - 1 - team) & 1);   // this team's last unit (< team: it has none)
        const bool u_has = u_last >= team;
        const int et = (int)threadIdx.x & 127;                // 0..127 inside the team
        const int row_in_tile = wq * 32 + lane;
        const double row_d = (double)row_in_tile;
        const double word_per_phase = 4294967296.0 / (double)rate;
        int* xw = sh->xch[u_has ? warp : 0];   // a team without units never touches them
        const int* xn = sh->xch[u_has ? 4 * team + ((wq + 1) & 3) : 0];
        const unsigned int lane_base = tmem_base + ((unsigned int)(wq * 32) << 16);
        for (int n = 0; n < my_tiles; ++n) {
            const I8Tile tl = tile_of(n);
            if (et < TG && ((et >> 3) & 1) == team) {   // the tones of this team's units
                const int ch = tl.ch0 + et;
                double2 bs = make_double2(0.0, 0.0);
                float isb = 0.f;
                if (ch < T) {
                    isb = inv_sb[ch];
                    if (rotate) bs = lo_phase_tile(freq[ch], rate, pos0, tl.row0, M);
                }
                sh->ph[et] = bs;
                sh->inv_sb[et] = isb;
            }
            mbar_wait_t(smem_u32(&sh->tmem_full), (unsigned)n & 1u, wt0, timed);
            tc_fence_after();
            // the tile's scale was fixed by the producers before they cut its first K block (they ran ahead of these MMAs)
            const float amax = __uint_as_float(sh->amax_bits[n & 3]);
            const float inv_sa = amax * (1.0f / I8_FULL_SCALE);
            const long long p = tl.row0 + row_in_tile;
            if (u_last < team) {   // nothing to read: the accumulators are free as far as this team is concerned
                tc_fence_before();
                mbar_arrive(smem_u32(&sh->tmem_empty));
            }
#pragma unroll 1
            for (int u = team; u < NUNIT; u += I8_EPI_TEAMS) {
                // One unit = 8 tones = 16 accumulator columns.  The F blocks are added with their row shift (the reference's
                // overlap-add, cpp/fir.cu:55-69) per accumulator, in int32 -- |D| <= 3 * 2 M * 2^14, times F, stays below 2^31 --
                // so the 64-bit combine D1 2^16 + D2 2^8 + D3 runs once per output and not once per block.
                int S[3][16];
                const long long te0 = timed ? clock64() : 0;
                // (measured and dropped: x8 loads software-pipelined against the shift-add of the previous one, 223 -> 216 GS/s; a
                // fifth operand stage: no change)
#pragma unroll
                for (int a = 0; a < 3; ++a) {   // one accumulator at a time: 16 sums + 16 fresh values live
#pragma unroll
                    for (int i = 0; i < F; ++i) {
                        const unsigned int col = 128u * a + (unsigned)((i * TG + u * 8) * 2);
                        int d[16];
                        tmem_ld16(lane_base + col, d);
                        tmem_ld_wait();
                        if (i == 0) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) S[a][j] = d[j];
                        } else {
                            if (lane < F - 1) {   // rows the previous warp needs: its lanes 32 - i .. 31 read them after the barrier
                                int4* dst = reinterpret_cast<int4*>(xw + (((lane * (F - 1) + (i - 1)) * 3) + a) * 16);
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_int4(d[4 * j4], d[4 * j4 + 1], d[4 * j4 + 2], d[4 * j4 + 3]);
                            }
                            const bool here = lane + i < 32;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int o = __shfl_down_sync(0xffffffffu, d[j], i);
                                if (here) S[a][j] += o;
                            }
                        }
                    }
                }
                // every accumulator column this team reads is in registers: the MMA warp may start the next tile
                if (u == u_last) {
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sh->tmem_empty));
                }
                const long long te1 = timed ? clock64() : 0;
                epi_bar(team);   // exchange buffers written; sh->ph / sh->inv_sb published
                if (timed) wt1 += te1 - te0, wt2 += clock64() - te1;
                if (F > 1) {
#pragma unroll
                    for (int i = 1; i < F; ++i) {
                        if (lane + i >= 32) {   // rows of the next warp (meaningless for the last warp: those outputs belong to the next tile)
                            const int4* src = reinterpret_cast<const int4*>(xn + (((lane + i - 32) * (F - 1) + (i - 1)) * 3) * 16);
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int j4 = 0; j4 < 4; ++j4) {
                                    const int4 v = src[4 * a + j4];
                                    S[a][4 * j4] += v.x, S[a][4 * j4 + 1] += v.y, S[a][4 * j4 + 2] += v.z, S[a][4 * j4 + 3] += v.w;
                                }
                        }
                    }
                }
                if (row_in_tile < RB && p < n_out) {
                    float2 o[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float sc = inv_sa * sh->inv_sb[u * 8 + t] * 65536.0f;   // the sum is in units of 2^-16 of a digit-1 product
                        o[t] = make_float2(i8_combine(S[0][2 * t], S[1][2 * t], S[2][2 * t]) * sc, i8_combine(S[0][2 * t + 1], S[1][2 * t + 1], S[2][2 * t + 1]) * sc);
                        if (rotate) {   // the channelizer form (pfb as GEMM) has no LO: whole turns per row
                            const unsigned int word = lo_phase_word(sh->ph[u * 8 + t], row_d, word_per_phase);
                            float sn, cs;
                            sincos_phase32(word, sn, cs);
                            o[t] = dev_cmul(o[t], make_float2(cs, -sn));
                        }
                    }
                    float2* dst = out + p * T + tl.ch0 + u * 8;
                    const int n_valid = T - (tl.ch0 + u * 8);   // tones of this unit that exist
                    if (n_valid >= 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                        for (int t = 0; t < 8; t += 2) *reinterpret_cast<float4*>(dst + t) = make_float4(o[t].x, o[t].y, o[t + 1].x, o[t + 1].y);
                    } else {
#pragma unroll
                        for (int t = 0; t < 8; ++t)
                            if (t < n_valid) dst[t] = o[t];
                    }
                }
                epi_bar(team);   // the exchange buffer (and sh->ph after the last unit) may be rewritten
            }
        }
        if (dbg && threadIdx.x == 0) {
            atomicAdd((unsigned long long*)&dbg[0], (unsigned long long)wt0);                     // epilogue: wait for accumulators
            atomicAdd((unsigned long long*)&dbg[1], (unsigned long long)wt1);                     // epilogue: TMEM loads + row shift-add
            atomicAdd((unsigned long long*)&dbg[14], (unsigned long long)wt2);                    // epilogue: first barrier of each unit
            atomicAdd((unsigned long long*)&dbg[2], (unsigned long long)(clock64() - t_role0));   // epilogue: role time
            atomicAdd((unsigned long long*)&dbg[15], (unsigned long long)my_tiles);
        }
    } else if (warp >= I8_PROD_WARP0 && warp < I8_MMA_WARP) {
        // ======================================= OPERAND PRODUCERS =======================================
        const int half = (warp - I8_PROD_WARP0) / 4;         // which 16 reals of the 32-real K block
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int row = 32 * q + lane;                       // window row of the tile == TMEM lane
        const unsigned int a_rowoff = (unsigned)(row >> 3) * 1024u + (unsigned)(row & 7) * 128u;
        const unsigned int a_tmem0 = tmem_base + ((unsigned int)(32 * q) << 16) + I8_COL_A + 4u * half;
        // byte order of the four K elements inside a 32-bit TMEM column: a_order = 0 -> K ascending with byte significance
        const unsigned int sel_pair = a_order ? 0x0004u : 0x0040u;   // (x.b0, y.b0) -> low half-word
        const unsigned int sel_quad = a_order ? 0x1054u : 0x5410u;
        int it = 0, sit = 0;   // K blocks cut, scan boxes consumed
        long long wt3 = 0, wt4 = 0, wt5 = 0;
        const int pt = (int)threadIdx.x - 32 * I8_PROD_WARP0;   // 0..255
        // The fixed-point scale of a tile comes from its largest |sample|.  The producers find it themselves, one tile ahead and
        // spread over the K blocks of the tile they are cutting.  Normal case (scan_by_tma): the scan-issue warp lands the NEXT
        // tile's K-block boxes in their own slots -- the same boxes the operand path will ask for one tile later, by then from
        // L2 -- and after cutting a K block each producer folds 64 bytes of a landed box into its running maximum: four LDS.128,
        // no load latency on the producers' critical path.  Otherwise (history rows, no tensor map, F = 8): the 256 threads
        // sweep the next tile with coalesced 8-byte global loads, eight per thread issued when a K block starts and folded in
        // when it is done.  8 loads x KB K blocks x 256 threads >= 128 M samples.
        const int scan_n = I8_ROWS * M;                              // samples per tile
        auto scan_src = [&](const I8Tile& t, bool& inside) -> long long {
            const long long t0 = t.row0 * (long long)M, t1 = t0 + (long long)I8_ROWS * M;
            inside = t0 >= w.n_hist && t1 <= w.n_hist + w.n_in;      // uniform over the CTA
            return t0;
        };
        auto publish_scale = [&](int n_next, float m) {              // all 256 producers, between two tiles
            const int slot = n_next & 3;
            if (amax_rows != nullptr) {                              // many tone groups share a row tile: a pre-pass has scanned it
                if (pt == 0) sh->amax_bits[slot] = __float_as_uint(amax_rows[tile_of(n_next).row0 / RB]);
            } else {
                const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));   // non-negative floats order like integers
                if (lane == 0) atomicMax(&sh->amax_bits[slot], wm);
            }
            asm volatile("bar.sync 2, %0;" ::"n"(32 * I8_PROD_WARPS) : "memory");
            if (pt == 0) sh->amax_bits[(n_next + 1) & 3] = 0u;       // the slot of the tile after: its last reader finished long ago
        };
        if (my_tiles > 0) {   // the first tile: scan it now
            if (pt < 4) sh->amax_bits[pt] = 0u;
            asm volatile("bar.sync 2, %0;" ::"n"(32 * I8_PROD_WARPS) : "memory");
            float m = 0.f;
            if (amax_rows == nullptr) {
                bool inside;
                const long long s0 = scan_src(tile_of(0), inside);
                for (int i = pt; i < scan_n; i += 32 * I8_PROD_WARPS) {
                    const float2 v = dev_win_at(w, s0 + i);
                    m = fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y)));
                }
            }
            publish_scale(0, m);
        }
        for (int n = 0; n < my_tiles; ++n) {
            const I8Tile tl = tile_of(n);
            const float amax = __uint_as_float(sh->amax_bits[n & 3]);
            const float sa = amax > 0.f ? I8_FULL_SCALE / amax : 0.f;
            const bool nx_tma = scan_by_tma(n + 1);
            const bool scan_next = amax_rows == nullptr && n + 1 < my_tiles && !nx_tma;
            bool nx_inside = false;
            const long long nx_s0 = scan_next ? scan_src(tile_of(n + 1), nx_inside) : 0;
            const float2* nx_p = w.in + (nx_s0 - w.n_hist);
            // byte distance from tile n+1 to tile n+2 of this CTA (same tone group order: consecutive tiles of a CTA are gridDim.x apart)
            const long long nx_stride_bytes = (n + 2 < my_tiles) ? (tile_of(n + 2).row0 - tile_of(n + 1).row0) * (long long)M * 8 : -1;
            const char* nx_end = reinterpret_cast<const char*>(w.in + w.n_in);
            float nx_m = 0.f;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int st = it % I8_AST, r = it & (I8_RAW - 1);
                const unsigned int st_par = (unsigned)(it / I8_AST) & 1u;
                float2 sv[8];
                if (scan_next) {   // eight loads of the next tile's samples, consumed when this K block is done
                    if (nx_inside) {   // straight-line: indices clamped into the tile (a repeated sample does not change a maximum)
                        // and the tile after that one: its lines are pulled towards L2 now (one 128-byte line per thread and K
                        // block), so that the loads above find them there one tile period from now instead of in HBM
                        {
                            const long long line = (long long)pt + 32LL * I8_PROD_WARPS * kb;
                            const char* q = reinterpret_cast<const char*>(nx_p) + nx_stride_bytes + line * 128;
                            if (nx_stride_bytes > 0 && line * 128 < (long long)scan_n * 8 && q + 128 <= nx_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                        }
#pragma unroll
                        for (int u8 = 0; u8 < 8; ++u8) {
                            const int idx = min(pt + 32 * I8_PROD_WARPS * (kb * 8 + u8), scan_n - 1);
                            unsigned long long raw;
                            asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(raw) : "l"(nx_p + idx));
                            sv[u8] = make_float2(__uint_as_float((unsigned int)raw), __uint_as_float((unsigned int)(raw >> 32)));
                        }
                    } else {
#pragma unroll
                        for (int u8 = 0; u8 < 8; ++u8) {
                            const int idx = pt + 32 * I8_PROD_WARPS * (kb * 8 + u8);
                            sv[u8] = idx < scan_n ? dev_win_at(w, nx_s0 + idx) : make_float2(0.f, 0.f);
                        }
                    }
                }
                mbar_wait_t(smem_u32(&sh->a_empty[st]), st_par ^ 1u, wt1, timed);   // the MMAs of this stage's previous use are done
                mbar_wait_t(smem_u32(&sh->raw_full[r]), (unsigned)(it >> 2) & 1u, wt2, timed);
                tc_fence_after();
                const long long tp0 = timed ? clock64() : 0;
                const unsigned char* raw = smem + (size_t)r * I8_RAW_BYTES;
                const bool from_smem = use_tma && tl.row0 + row >= hist_rows;   // else: history row, or no TMA at all
                const long long s_row = (tl.row0 + row) * (long long)M + kb * I8_KC + 8 * half;
                float v[16];
                if (from_smem) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 x = *reinterpret_cast<const float4*>(raw + a_rowoff + ((unsigned)((4 * half + c4) ^ (row & 7)) << 4));
                        v[4 * c4] = x.x, v[4 * c4 + 1] = x.y, v[4 * c4 + 2] = x.z, v[4 * c4 + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = kb * I8_KC + 8 * half + j;
                        float2 a = make_float2(0.f, 0.f);
                        if (k < M) a = (use_tma && s_row + j < w.n_hist && s_row + j < I8_HIST_MAX) ? sh->hist[s_row + j] : dev_win_at(w, s_row + j);
                        v[2 * j] = a.x, v[2 * j + 1] = a.y;
                    }
                }
                mbar_arrive(smem_u32(&sh->raw_empty[r]));   // this thread's reads of the landing slot are complete (values in registers)
                // fixed point, three signed digits: q = d1 2^16 + d2 2^8 + d3, each the low byte taken as signed with the carry moved
                // up: d3 = byte 0 of q, d2 = byte 1 of q + 128, d1 = byte 2 of q + 128 + 128 * 256 (the two carries folded into one add)
                unsigned int w1[4], w2[4], w3[4];
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    int q0[4], t1[4], t2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        q0[e] = __float2int_rn(v[4 * g4 + e] * sa);
                        t1[e] = q0[e] + 128;
                        t2[e] = q0[e] + 32896;
                    }
                    w3[g4] = __byte_perm(__byte_perm(q0[0], q0[1], sel_pair), __byte_perm(q0[2], q0[3], sel_pair), sel_quad);
                    w2[g4] = __byte_perm(__byte_perm(t1[0], t1[1], sel_pair + 0x11u), __byte_perm(t1[2], t1[3], sel_pair + 0x11u), sel_quad);
                    w1[g4] = __byte_perm(__byte_perm(t2[0], t2[1], sel_pair + 0x22u), __byte_perm(t2[2], t2[3], sel_pair + 0x22u), sel_quad);
                }
                __syncwarp();   // tcgen05.st is warp-collective: reconverge after the per-lane source selection
                const long long tp1 = timed ? clock64() : 0;
                const unsigned int ta = a_tmem0 + 24u * st;
                tmem_st4(ta, w1[0], w1[1], w1[2], w1[3]);
                tmem_st4(ta + 8u, w2[0], w2[1], w2[2], w2[3]);
                tmem_st4(ta + 16u, w3[0], w3[1], w3[2], w3[3]);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&sh->a_full[st]));
                if (timed) wt4 += tp1 - tp0, wt5 += clock64() - tp1;
                if (scan_next) {
#pragma unroll
                    for (int u8 = 0; u8 < 8; ++u8) nx_m = fmaxf(nx_m, fmaxf(fabsf(sv[u8].x), fabsf(sv[u8].y)));
                }
                if (NSCAN > 0 && nx_tma) {   // this thread's 64 bytes of the next tile's box kb (any order: it is a maximum)
                    const int ss = sit % (NSCAN > 0 ? NSCAN : 1);
                    mbar_wait_t(smem_u32(&sh->scan_full[ss]), (unsigned)(sit / (NSCAN > 0 ? NSCAN : 1)) & 1u, wt3, timed);
                    const float4* bx = reinterpret_cast<const float4*>(smem_scan + (size_t)ss * I8_RAW_BYTES) + pt;
                    float4 x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = bx[256 * j];
                    mbar_arrive(smem_u32(&sh->scan_empty[ss]));
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        nx_m = fmaxf(fmaxf(nx_m, fmaxf(fabsf(x[j].x), fabsf(x[j].y))), fmaxf(fabsf(x[j].z), fabsf(x[j].w)));
                    ++sit;
                }
            }
            if (n + 1 < my_tiles) {
                const long long tsc = timed ? clock64() : 0;
                publish_scale(n + 1, nx_m);
                if (timed) wt0 += clock64() - tsc;
            }
        }
        if (dbg && (int)threadIdx.x == 32 * I8_PROD_WARP0) {
            atomicAdd((unsigned long long*)&dbg[7], (unsigned long long)wt0);                     // producers: scale hand-over between tiles
            atomicAdd((unsigned long long*)&dbg[8], (unsigned long long)wt1);                     // producers: wait for a free A stage
            atomicAdd((unsigned long long*)&dbg[9], (unsigned long long)wt2);                     // producers: wait for the TMA rows
            atomicAdd((unsigned long long*)&dbg[10], (unsigned long long)(clock64() - t_role0));  // producers: role time
            atomicAdd((unsigned long long*)&dbg[16], (unsigned long long)wt3);                    // producers: wait for a scan box
            atomicAdd((unsigned long long*)&dbg[18], (unsigned long long)wt4);                    // producers: rows out of shared memory + digits
            atomicAdd((unsigned long long*)&dbg[19], (unsigned long long)wt5);                    // producers: tcgen05.st + wait + hand-over
        }
    } else {
        // three single-thread issue roles and an idle warp
        if (warp == I8_MMA_WARP) {
        // ======================================= MMA ISSUE =======================================
        int it = 0, bc = 0;   // K blocks issued (A stage = it & 3), B stages consumed (stage = bc & 1)
        long long wt3 = 0;
        const unsigned int d1 = tmem_base, d2 = tmem_base + 128u, d3 = tmem_base + 256u;
        for (int n = 0; n < my_tiles; ++n) {
            mbar_wait_t(smem_u32(&sh->tmem_empty), ((unsigned)n & 1u) ^ 1u, wt0, timed);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int st = it % I8_AST, bs = bc & (I8_BST - 1);
                if ((kb & 3) == 0) mbar_wait_t(smem_u32(&sh->b_full[bs]), (unsigned)(bc >> 1) & 1u, wt1, timed);
                mbar_wait_t(smem_u32(&sh->a_full[st]), (unsigned)(it / I8_AST) & 1u, wt2, timed);
                tc_fence_after();
                const long long tm0 = timed ? clock64() : 0;
                if (lane == 0) {
                    const unsigned int a1 = tmem_base + I8_COL_A + 24u * st, a2 = a1 + 8u, a3 = a1 + 16u;
                    const unsigned int b0 = smem_u32(smem_b) + (unsigned)bs * I8_B_STAGE;
                    const unsigned long long adv = (unsigned long long)(2 * (kb & 3));   // 32 bytes >> 4 per K block inside the 128-byte row
                    const unsigned long long B1 = i8_smem_desc(b0) + adv, B3 = i8_smem_desc(b0 + 2 * I8_B_PLANE) + adv;
                    const unsigned int acc = kb > 0 ? 1u : 0u;
                    mma_i8_ts(d3, a1, B3, I8_IDESC, acc);          // 2^16: a1 b3
                    mma_i8_ts(d1, a1, B1, I8_IDESC_WIDE, acc);     // 2^32: a1 b1 | 2^24: a1 b2      ([B1 | B2] -> [D1 | D2])
                    mma_i8_ts(d2, a2, B1, I8_IDESC_WIDE, 1u);      // 2^24: a2 b1 | 2^16: a2 b2      ([B1 | B2] -> [D2 | D3])
                    mma_i8_ts(d3, a3, B1, I8_IDESC, 1u);           // 2^16: a3 b1
                    tc_commit(smem_u32(&sh->a_empty[st]));
                    if ((kb & 3) == 3 || kb == KB - 1) tc_commit(smem_u32(&sh->b_empty[bs]));
                    if (kb == KB - 1) tc_commit(smem_u32(&sh->tmem_full));
                }
                __syncwarp();
                if (timed) wt3 += clock64() - tm0;
                if ((kb & 3) == 3 || kb == KB - 1) ++bc;
            }
        }
        if (dbg && lane == 0) {
            atomicAdd((unsigned long long*)&dbg[3], (unsigned long long)wt0);                     // MMA: wait for free accumulators
            atomicAdd((unsigned long long*)&dbg[4], (unsigned long long)wt1);                     // MMA: wait for B planes
            atomicAdd((unsigned long long*)&dbg[5], (unsigned long long)wt2);                     // MMA: wait for A digits
            atomicAdd((unsigned long long*)&dbg[6], (unsigned long long)(clock64() - t_role0));   // MMA: role time
            atomicAdd((unsigned long long*)&dbg[17], (unsigned long long)wt3);                    // MMA: the issue block (4 MMAs + commits)
        }
        } else if (warp == I8_TMA_WARP) {
        // ======================================= TMA ISSUE =======================================
        if (lane == 0) {
            int it = 0, bc = 0;
            for (int n = 0; n < my_tiles; ++n) {
                const I8Tile tl = tile_of(n);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    if ((kb & 3) == 0) {   // the three digit planes of the next four K blocks of this tone group
                        const int bs = bc & (I8_BST - 1);
                        mbar_wait_t(smem_u32(&sh->b_empty[bs]), ((unsigned)(bc >> 1) & 1u) ^ 1u, wt0, timed);
                        const unsigned int bar = smem_u32(&sh->b_full[bs]);
                        mbar_arrive_expect_tx(bar, I8_B_STAGE);
                        const int row_b = ((tl.tg * KQ + (kb >> 2)) * 3) * I8_N;
                        const unsigned int dst = smem_u32(smem_b) + (unsigned)bs * I8_B_STAGE;
                        for (int pl = 0; pl < 3; ++pl) tma_load_2d(dst + (unsigned)pl * I8_B_PLANE, &tmap_b, 0, row_b + pl * I8_N, bar);
                        ++bc;
                    }
                    const int s = it & (I8_RAW - 1);
                    mbar_wait_t(smem_u32(&sh->raw_empty[s]), ((unsigned)(it >> 2) & 1u) ^ 1u, wt1, timed);
                    const unsigned int bar = smem_u32(&sh->raw_full[s]);
                    if (use_tma) {
                        mbar_arrive_expect_tx(bar, I8_RAW_BYTES);
                        tma_load_2d(smem_base + (unsigned)s * I8_RAW_BYTES, &tmap_a, kb * 2 * I8_KC, (int)(tl.row0 - hist_rows), bar);
                    } else {
                        mbar_arrive(bar);
                    }
                }
            }
            if (dbg) {
                atomicAdd((unsigned long long*)&dbg[11], (unsigned long long)wt0);                     // TMA: wait for a free B stage
                atomicAdd((unsigned long long*)&dbg[12], (unsigned long long)wt1);                     // TMA: wait for a free landing slot
                atomicAdd((unsigned long long*)&dbg[13], (unsigned long long)(clock64() - t_role0));   // TMA: role time
            }
        }
        } else if (warp == I8_SCAN_WARP) {
        // ======================================= SCAN BOX ISSUE =======================================
        if (NSCAN > 0 && lane == 0) {
            int sit = 0;
            for (int n = 0; n + 1 < my_tiles; ++n) {
                if (!scan_by_tma(n + 1)) continue;
                const I8Tile nx = tile_of(n + 1);
                for (int kb = 0; kb < KB; ++kb, ++sit) {
                    const int ss = sit % (NSCAN > 0 ? NSCAN : 1);
                    mbar_wait(smem_u32(&sh->scan_empty[ss]), ((unsigned)(sit / (NSCAN > 0 ? NSCAN : 1)) & 1u) ^ 1u);
                    const unsigned int bar = smem_u32(&sh->scan_full[ss]);
                    mbar_arrive_expect_tx(bar, I8_RAW_BYTES);
                    tma_load_2d(smem_u32(smem_scan) + (unsigned)ss * I8_RAW_BYTES, &tmap_a, kb * 2 * I8_KC, (int)(nx.row0 - hist_rows), bar);
                }
            }
        }
        }
    }

    // teardown: every MMA has completed before the last tmem_full arrival, every tcgen05.ld before this barrier
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == I8_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

// Largest |sample| of every row tile (rows [t RB, t RB + 128) of M samples), one block per tile: used when several tone groups
// share each row tile, so that the scan is done once and not once per (row tile, tone group).
__global__ void __launch_bounds__(256) direct_i8_amax_kernel(const Window w, int M, int RB, float* __restrict__ amax_rows) {
    __shared__ unsigned int best;
    if (threadIdx.x == 0) best = 0u;
    __syncthreads();
    const long long s0 = (long long)blockIdx.x * RB * M, s1 = s0 + (long long)I8_ROWS * M;
    float m = 0.f;
    for (long long s = s0 + threadIdx.x; s < s1; s += 256 * 8) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (s + u * 256 < s1) ? dev_win_at(w, s + u * 256) : make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) m = fmaxf(m, fmaxf(fabsf(v[u].x), fabsf(v[u].y)));
    }
    const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0) atomicMax(&best, wm);
    __syncthreads();
    if (threadIdx.x == 0) amax_rows[blockIdx.x] = __uint_as_float(best);
}

typedef CUresult (*I8EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
I8EncodeTiledFn i8_encode_fn() {
    static I8EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<I8EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// The part of the `in` segment that starts on a window-row boundary as a 2-D tensor [rows][2 M floats] (row pitch M * 8
// bytes), box = 128 rows x 32 floats, 128-byte swizzle.  *hist_rows = window rows that hold carried-over samples.
bool i8_make_tensor_map_a(const Window& w, int M, CUtensorMap* map, long long* hist_rows, bool allow_tma) {
    const long long hr = (w.n_hist + M - 1) / M, off0 = hr * M - w.n_hist;
    *hist_rows = hr;
    if (!allow_tma || M < I8_KC || (M & 1) || w.n_in - off0 < M) return false;
    const float2* base = w.in + off0;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    I8EncodeTiledFn enc = i8_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)(2 * M), (cuuint64_t)((w.n_in - off0) / M)};
    const cuuint64_t strides[1] = {(cuuint64_t)M * 8};
    const cuuint32_t box[2] = {2 * I8_KC, I8_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int F>
int i8_launch(DirectI8Bank& bank, const Window& w, const int* freq_dev, int T, int M, int rate, long long pos0, long long n_out,
              float2* out, int sm_count, cudaStream_t stream, int rotate, bool allow_tma) {
    constexpr int TG = 64 / F, RB = I8_ROWS - (F - 1);
    const int row_tiles = (int)((n_out + RB - 1) / RB), tone_groups = (T + TG - 1) / TG;
    static DeviceOnce attr_once;
    if (const int dev = attr_once.pending(); dev >= 0) {
        GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_i8_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem_bytes(F)));
        GSDR_CUDA_OK(cudaFuncSetAttribute(direct_fir_i8_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem_bytes(F)));
        attr_once.done(dev);
    }
    CUtensorMap map_a;
    memset(&map_a, 0, sizeof(map_a));
    long long hist_rows = 0;
    const int use_tma = i8_make_tensor_map_a(w, M, &map_a, &hist_rows, allow_tma) ? 1 : 0;
    CUtensorMap map_b;
    static_assert(sizeof(map_b) == sizeof(bank.tmap_b), "tensor map storage");
    memcpy(&map_b, bank.tmap_b, sizeof(map_b));
    const long long tiles = (long long)row_tiles * tone_groups;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    int launches = 1;
    const float* amax_rows = nullptr;
    if (tone_groups > 2) {   // the in-kernel scan would read every row tile once per tone group
        if ((size_t)row_tiles > bank.row_amax_cap) {
            if (bank.d_row_amax) cudaFree(bank.d_row_amax);
            bank.d_row_amax = nullptr;
            bank.row_amax_cap = 0;
            GSDR_CUDA_OK(cudaMalloc(&bank.d_row_amax, sizeof(float) * (size_t)row_tiles));
            bank.row_amax_cap = (size_t)row_tiles;
        }
        direct_i8_amax_kernel<<<row_tiles, 256, 0, stream>>>(w, M, RB, bank.d_row_amax);
        GSDR_CUDA_OK(cudaGetLastError());
        amax_rows = bank.d_row_amax;
        launches = 2;
    }
    // GSDR_DIRECT_I8_DEBUG=1: per-role wait / run cycles of this launch on stderr (synchronises; schedule tuning only)
    static const bool debug = [] {
        const char* e = getenv("GSDR_DIRECT_I8_DEBUG");
        return e && e[0] == '1';
    }();
    long long* dbg = nullptr;
    if (debug) {
        GSDR_CUDA_OK(cudaMalloc(&dbg, 20 * sizeof(long long)));
        GSDR_CUDA_OK(cudaMemsetAsync(dbg, 0, 20 * sizeof(long long), stream));
    }
    static const unsigned int a_order = [] {
        const char* e = getenv("GSDR_I8_AORDER");
        return (e && e[0] == '1') ? 1u : 0u;
    }();
    if (rotate)
        direct_fir_i8_kernel<F, true><<<grid, I8_THREADS, i8_smem_bytes(F), stream>>>(map_a, map_b, use_tma, hist_rows, w, bank.d_inv_sb, freq_dev, T, M,
                                                                                   rate, pos0, n_out, row_tiles, tone_groups, bank.KQ, a_order, amax_rows, out, dbg);
    else
        direct_fir_i8_kernel<F, false><<<grid, I8_THREADS, i8_smem_bytes(F), stream>>>(map_a, map_b, use_tma, hist_rows, w, bank.d_inv_sb, freq_dev, T, M,
                                                                                    rate, pos0, n_out, row_tiles, tone_groups, bank.KQ, a_order, amax_rows, out, dbg);
    GSDR_CUDA_OK(cudaGetLastError());
    if (dbg) {
        long long h[20];
        GSDR_CUDA_OK(cudaStreamSynchronize(stream));
        GSDR_CUDA_OK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        const double nt = h[15] > 0 ? (double)h[15] : 1.0;   // tiles
        fprintf(stderr,
                "[direct_i8] grid %d tiles %lld KB %d use_tma %d | cycles per tile: epilogue wait-acc %.0f loads+shift %.0f bar %.0f of %.0f | mma wait-acc %.0f wait-B %.0f "
                "wait-A %.0f issue %.0f of %.0f | producers scale-handover %.0f wait-stage %.0f wait-rows %.0f wait-scan %.0f digits %.0f store %.0f of %.0f | tma wait-B %.0f wait-slot %.0f of %.0f\n",
                grid, tiles, (M + I8_KC - 1) / I8_KC, use_tma, h[0] / nt, h[1] / nt, h[14] / nt, h[2] / nt, h[3] / nt, h[4] / nt, h[5] / nt, h[17] / nt, h[6] / nt, h[7] / nt, h[8] / nt,
                h[9] / nt, h[16] / nt, h[18] / nt, h[19] / nt, h[10] / nt, h[11] / nt, h[12] / nt, h[13] / nt);
    }
    return launches;
}

}  // namespace

// F in {1, 2, 4, 8} so that F * TG * 2 = 128 accumulator columns; the LO phase arithmetic needs 128 * rate < 2^53; the int32
// accumulators hold three digit products of at most 2^14 per real tap: 3 * 2 M * 2^14 < 2^31.
bool direct_fir_i8_supported(int T, int M, int ntaps, long long n_out) {
    if (M < 1 || T < 1 || n_out < 1 || ntaps % M != 0 || M > 16384) return false;
    const int f = ntaps / M;
    if (!(f == 1 || f == 2 || f == 4 || f == 8)) return false;
    const long long rb = I8_ROWS - (f - 1), tg = 64 / f;
    const long long tiles = ((n_out + rb - 1) / rb) * ((T + tg - 1) / tg);
    const long long kq = ((M + I8_KC - 1) / I8_KC + 3) / 4;
    const long long bank_bytes = ((T + tg - 1) / tg) * kq * I8_B_STAGE;
    return tiles < (1ll << 30) && bank_bytes <= (512ll << 20);
}

// Worth it when there are enough (row tile, tone group) tiles to occupy a good part of the GPU.
bool direct_fir_i8_preferred(int T, int M, int ntaps, long long n_out) {
    if (!direct_fir_i8_supported(T, M, ntaps, n_out)) return false;
    const int f = ntaps / M;
    const long long rb = I8_ROWS - (f - 1), tg = 64 / f;
    const long long tiles = ((n_out + rb - 1) / rb) * ((T + tg - 1) / tg);
    return tiles >= 64;
}

// The filter bank as digit planes.  g = T x ntaps complex taps in double (re, im interleaved).  Layout in HBM: rows of 128
// bytes; row ((tg KQ + kq) 3 + plane) 128 + n holds, for accumulator column n = (block i, tone t, re/im) of tone group tg, the
// digits of plane `plane` (0: 2^16, 1: 2^8, 2: 2^0) for K blocks 4 kq .. 4 kq + 3 (32 bytes each: reals 2k, 2k+1 of taps
// 16 kb + k).  TMA (SWIZZLE_128B) turns 128 such rows into one K-major operand tile.
int direct_i8_bank_create(const double* g, int T, int M, int ntaps, DirectI8Bank* bank) {
    memset(bank, 0, sizeof(*bank));
    if (!direct_fir_i8_supported(T, M, ntaps, 1)) {
        set_error("direct_i8_bank_create: unsupported shape (T=%d M=%d ntaps=%d)", T, M, ntaps);
        return -1;
    }
    const int F = ntaps / M, TG = 64 / F, tone_groups = (T + TG - 1) / TG;
    const int KB = (M + I8_KC - 1) / I8_KC, KQ = (KB + 3) / 4;
    const size_t rows = (size_t)tone_groups * KQ * 3 * I8_N;
    std::vector<signed char> host(rows * 128, 0);
    std::vector<float> inv_sb((size_t)tone_groups * TG, 0.f);
    for (int t = 0; t < T; ++t) {
        double mx = 0.0;
        for (int m = 0; m < 2 * ntaps; ++m) mx = std::fmax(mx, std::fabs(g[(size_t)t * 2 * ntaps + m]));
        const double sb = mx > 0.0 ? (double)I8_FULL_SCALE / mx : 0.0;
        inv_sb[t] = mx > 0.0 ? (float)(1.0 / sb) : 0.f;
        const int tg = t / TG, tt = t % TG;
        for (int i = 0; i < F; ++i)
            for (int k = 0; k < M; ++k) {
                const double gr = g[((size_t)t * ntaps + (size_t)i * M + k) * 2], gi = g[((size_t)t * ntaps + (size_t)i * M + k) * 2 + 1];
                // column Re: (x_r, x_i) . (g_r, -g_i); column Im: (x_r, x_i) . (g_i, g_r)
                const double col_re[2] = {gr, -gi}, col_im[2] = {gi, gr};
                const int kb = k / I8_KC, kk = k % I8_KC, kq = kb / 4, kj = kb % 4;
                for (int c = 0; c < 2; ++c) {
                    const int ncol = (i * TG + tt) * 2 + c;
                    for (int e = 0; e < 2; ++e) {
                        const long long q0 = llrint((c == 0 ? col_re[e] : col_im[e]) * sb);
                        const long long q1 = (q0 + 128) >> 8, q2 = (q1 + 128) >> 8;
                        const signed char dg[3] = {(signed char)(q2 & 0xff), (signed char)(q1 & 0xff), (signed char)(q0 & 0xff)};
                        for (int pl = 0; pl < 3; ++pl) {
                            const size_t row = ((size_t)(tg * KQ + kq) * 3 + pl) * I8_N + ncol;
                            host[row * 128 + kj * 32 + 2 * kk + e] = dg[pl];
                        }
                    }
                }
            }
    }
    void* d = nullptr;
    float* d_inv = nullptr;
    GSDR_CUDA_OK(cudaMalloc(&d, host.size()));
    if (cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess || cudaMalloc(&d_inv, sizeof(float) * inv_sb.size()) != cudaSuccess ||
        cudaMemcpy(d_inv, inv_sb.data(), sizeof(float) * inv_sb.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("direct_i8_bank_create: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(d);
        if (d_inv) cudaFree(d_inv);
        return -1;
    }
    I8EncodeTiledFn enc = i8_encode_fn();
    CUtensorMap map;
    const cuuint64_t dims[2] = {128, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {128, I8_N};
    const cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        set_error("direct_i8_bank_create: cuTensorMapEncodeTiled failed for the filter bank");
        cudaFree(d);
        cudaFree(d_inv);
        return -1;
    }
    static_assert(sizeof(map) == sizeof(bank->tmap_b), "tensor map storage");
    memcpy(bank->tmap_b, &map, sizeof(map));
    bank->d_bank = d;
    bank->d_inv_sb = d_inv;
    bank->KQ = KQ;
    bank->tone_groups = tone_groups;
    return 0;
}

void direct_i8_bank_destroy(DirectI8Bank* bank) {
    if (bank->d_row_amax) cudaFree(bank->d_row_amax);
    if (bank->d_bank) cudaFree(bank->d_bank);
    if (bank->d_inv_sb) cudaFree(bank->d_inv_sb);
    memset(bank, 0, sizeof(*bank));
}

int direct_fir_i8_launch(DirectI8Bank& bank, const Window& w, const int* freq_dev, int T, int M, int ntaps, int rate, long long pos0,
                         long long n_out, float2* out, int sm_count, cudaStream_t stream, int rotate, bool allow_tma) {
    if (n_out <= 0) return 0;
    if (!bank.d_bank || !direct_fir_i8_supported(T, M, ntaps, n_out)) {
        set_error("direct_fir_i8_launch: unsupported shape (T=%d M=%d ntaps=%d) or missing bank", T, M, ntaps);
        return -1;
    }
    switch (ntaps / M) {
        case 1: return i8_launch<1>(bank, w, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
        case 2: return i8_launch<2>(bank, w, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
        case 4: return i8_launch<4>(bank, w, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
        default: return i8_launch<8>(bank, w, freq_dev, T, M, rate, pos0, n_out, out, sm_count, stream, rotate, allow_tma);
    }
}

}  // namespace gsdr
