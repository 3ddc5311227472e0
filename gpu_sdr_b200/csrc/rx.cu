// gsdr_rx: host side of the RX demodulator (the RX_buffer_demodulator replacement).
//
// Reference call structure being replaced (cpp/USRP_demodulator.cpp):
//   ctor :7-327 (mode dispatch, taps, plans, helpers)   process_* :335-649   close_* :466-698
// One instance == one IQ stream.  Differences in *how*, not *what*:
//   - three streams per instance (copy-in | compute | copy-out) and a ring of pipeline slots, so
//     H2D of buffer k+1 overlaps the kernels of buffer k and the D2H of buffer k-1; the blocking
//     process() is submit()+wait();
//   - the carry-over between buffers (PFB spare samples, FIR history, chirp remainder) lives in a
//     small ping-pong device buffer and the kernels read "history ++ new samples" directly: there
//     is no move_buffer and no upload-at-offset;
//   - valid lengths come from the same integer helpers as the reference, evaluated on the host
//     before the GPU runs, and only the valid part is copied back.
#include <chrono>
#include <cmath>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "common.hpp"

using namespace gsdr;

namespace {

struct Slot {
    float2* d_in = nullptr;
    float2* d_out = nullptr;
    short2* d_raw = nullptr;  // sc16 ingest: the wire-format copy of the buffer (allocated on first use)
    cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
    bool used = false;
};

// sc16 -> fc32 on the device.  With a USRP the wire format is sc16 and UHD converts to fc32 on the host CPU (the
// reference asks for stream_args_t("fc32"), cpp/USRP_hardware_manager.cpp:764-820; "The UHD libraries use the CPU to
// convert the data", server_docs/01_installation.md:9) with the scale factor 1/32767.  Doing it here halves the bytes
// that cross PCIe and the pinned-pool footprint.  4 samples per thread: one 16-byte load, two 16-byte stores.
constexpr float kSc16Scale = 1.0f / 32767.0f;
__global__ void __launch_bounds__(256)
sc16_to_fc32_kernel(const short2* __restrict__ in, long long n, float2* __restrict__ out) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(in) + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float4 o[2];
        float* of = reinterpret_cast<float*>(o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            of[2 * k] = (float)(short)(w[k] & 0xffff) * kSc16Scale;
            of[2 * k + 1] = (float)(short)(w[k] >> 16) * kSc16Scale;
        }
        reinterpret_cast<float4*>(out)[2 * i] = o[0];
        reinterpret_cast<float4*>(out)[2 * i + 1] = o[1];
    }
    // ragged tail (buffer_len not a multiple of 4)
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const short2 v = in[i];
        out[i] = make_float2((float)v.x * kSc16Scale, (float)v.y * kSc16Scale);
    }
}

}  // namespace

struct gsdr_rx {
    int device = 0, sm_count = 148;
    int mode = GSDR_NODSP;
    bool diagnostic = false;
    // parameters (copied)
    int rate = 0, N = 0;
    long long L = 0, decim = 0, P = 0;
    std::vector<int32_t> freq;
    int T = 0;  // channels == wave_type.size()
    float fcut = 0.f;

    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::vector<Slot> slots;
    uint64_t tickets = 0;
    uint64_t launches = 0;
    size_t max_out = 0;
    const char* kernel_name = "none";
    std::string kernel_name_buf;

    // carry-over (ping-pong)
    float2* hist[2] = {nullptr, nullptr};
    int hist_cur = 0;
    long long n_hist = 0, hist_cap = 0;

    // TONES / NOISE
    std::vector<float> taps_host;
    std::vector<int32_t> bins_host;
    float* d_taps = nullptr;
    int* d_bins = nullptr;
    unsigned char* d_xperm = nullptr;  // fused kernel: in-row placement that makes the tone gather conflict-free
    float2* d_tw = nullptr;
    void* d_work = nullptr;
    size_t work_bytes = 0;
    // NOISE with decim > 0: spectra of the current call, running sum of the open group, frames already in it
    // blocking process(): chunk events (H2D of chunk c done / kernels of chunk c done)
    std::vector<cudaEvent_t> ev_chunk_in, ev_chunk_comp;
    float2* d_spec = nullptr;
    size_t spec_bytes = 0;
    float2* d_spec_acc = nullptr;
    int spec_carried = 0;
    bool post_decim = false;  // NOISE with decim > 0, TONES with decim > 1
    bool pfb_tc = false;      // channelizer size without a fused kernel: the filter bank as a GEMM on the tensor cores
    int batching = 0, T_sel = 0;
    bool fused = false;
    bool zc_enabled = true;    // GSDR_PROCESS_ZEROCOPY (read once, at create): blocking process() on pinned buffers is one launch
    bool host_window = false;  // this call's input window is pinned host memory read in place (zero-copy blocking call)
    bool tc_tma = true, tc_host_tma = false;  // GSDR_DIRECT_TC_TMA=0 / GSDR_DIRECT_TC_HOST_TMA=1, read at create
    bool in_sc16 = false;      // this call's input window holds int16 I/Q pairs, converted inside the fused channelizer
    gsdr_buffer_helper bh{};

    // CHIRP
    gsdr_chirp_param cpar{};
    ChirpDev cdev{};
    unsigned long long last_index = 0;
    int ppt = 0;
    gsdr_vna_helper vh{};
    float* d_profile = nullptr;
    float2* d_partial = nullptr;
    size_t partial_cap = 0;

    // DIRECT
    float2* d_g = nullptr;
    int* d_freq = nullptr;
    int ntaps = 0;
    long long index_counter = 0;
    bool direct_tc = false;
    // exact integer tensor-core path (direct_i8_kernels.cu), DIRECT and the generic-size channelizer
    bool direct_i8 = false, pfb_i8 = false;
    DirectI8Bank i8bank;
};

namespace {

int set_dev(const gsdr_rx* rx) { return ensure_device(rx->device); }
bool tc_tma_allowed(const gsdr_rx* rx) { return rx->tc_tma && (!rx->host_window || rx->tc_host_tma); }

// Stream state that a call advances.  enqueue_compute works on the live fields and puts this snapshot back when anything
// fails after the helpers have moved, so a failed call (cudaMalloc, a launch) leaves the demodulator where it was.
struct RxState {
    gsdr_buffer_helper bh;
    gsdr_vna_helper vh;
    int spec_carried, hist_cur;
    long long n_hist, index_counter;
    unsigned long long last_index;
};
RxState save_state(const gsdr_rx* rx) {
    return RxState{rx->bh, rx->vh, rx->spec_carried, rx->hist_cur, rx->n_hist, rx->index_counter, rx->last_index};
}
void restore_state(gsdr_rx* rx, const RxState& st) {
    rx->bh = st.bh, rx->vh = st.vh, rx->spec_carried = st.spec_carried, rx->hist_cur = st.hist_cur;
    rx->n_hist = st.n_hist, rx->index_counter = st.index_counter, rx->last_index = st.last_index;
}

template <class Tp>
int dev_upload(Tp** dst, const Tp* src, size_t n) {
    GSDR_CUDA_OK(cudaMalloc(dst, sizeof(Tp) * (n ? n : 1)));
    if (n) GSDR_CUDA_OK(cudaMemcpy(*dst, src, sizeof(Tp) * n, cudaMemcpyHostToDevice));
    return 0;
}

// twiddle tables of the fused 2048-point kernels.  Lock-step kernel (16x16x8): [16][128] W_2048^(l k1)
// then [8][16] W_128^(n3 k2).  Warp-specialised kernel (8x16x16): [8][256] W_2048^(l k1) then
// [16][16] W_256^(n3 k2).  All built in double.
std::vector<float2> fused_twiddles() {
    std::vector<float2> tw(16 * 128 + 8 * 16 + 8 * 256 + 16 * 16);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 0; k1 < 16; ++k1)
        for (int l = 0; l < 128; ++l) {
            const double a = -two_pi * (double)((l * k1) % 2048) / 2048.0;
            tw[k1 * 128 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int n3 = 0; n3 < 8; ++n3)
        for (int k2 = 0; k2 < 16; ++k2) {
            const double a = -two_pi * (double)((n3 * k2) % 128) / 128.0;
            tw[16 * 128 + n3 * 16 + k2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    const int base = 16 * 128 + 8 * 16;
    for (int k1 = 0; k1 < 8; ++k1)
        for (int l = 0; l < 256; ++l) {
            const double a = -two_pi * (double)((l * k1) % 2048) / 2048.0;
            tw[base + k1 * 256 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int n3 = 0; n3 < 16; ++n3)
        for (int k2 = 0; k2 < 16; ++k2) {
            const double a = -two_pi * (double)((n3 * k2) % 256) / 256.0;
            tw[base + 8 * 256 + n3 * 16 + k2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    return tw;
}

std::vector<float2> generic_twiddles(int N) {
    std::vector<float2> tw(N);
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < N; ++i) {
        const double a = -two_pi * (double)i / (double)N;
        tw[i] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    return tw;
}

int init_pfb(gsdr_rx* rx, bool all_bins) {
    const int N = rx->N, P = (int)rx->P;
    if (N < 2 || P < 1) {
        set_error("TONES/NOISE: fft_tones=%d pf_average=%d is not a valid filter bank", N, P);
        return -1;
    }
    rx->fcut = (float)(1. / (2 * N));  // cpp/USRP_demodulator.cpp:131
    rx->taps_host.resize((size_t)N * P);
    make_sinc_window(N * P, rx->fcut, rx->taps_host.data());
    if (rx->diagnostic) {
        // init_diagnostic = true: the reference dumps the polyphase window as float2 (imaginary parts zero) into the working
        // directory (make_sinc_window(..., diagnostic, ...), cpp/kernels.cu:290-296; call at cpp/USRP_demodulator.cpp:134)
        fprintf(stderr, "gsdr_rx: warning: Demodulator diagnostic enabled.\n");
        if (FILE* f = fopen("USRP_polyphase_filter_window.dat", "wb")) {
            std::vector<float2> wz(rx->taps_host.size());
            for (size_t i = 0; i < wz.size(); ++i) wz[i] = make_float2(rx->taps_host[i], 0.f);
            fwrite(wz.data(), sizeof(float2), wz.size(), f);
            fclose(f);
        }
    }
    rx->batching = pfb_batching((int)rx->L, N, P);
    if (all_bins) {
        rx->T_sel = N;  // full spectrum, bins in natural order (cpp/USRP_demodulator.cpp:301)
    } else {
        rx->T_sel = rx->T;
        rx->bins_host.resize(rx->T);
        tone_bins(rx->rate, N, rx->freq.data(), rx->T, rx->bins_host.data());
        for (auto& b : rx->bins_host)
            if (b < 0) b = 0;  // the reference leaves unmatched tones uninitialised; pin them to DC
        if (dev_upload(&rx->d_bins, rx->bins_host.data(), rx->bins_host.size())) return -1;
    }
    if (dev_upload(&rx->d_taps, rx->taps_host.data(), rx->taps_host.size())) return -1;
    buffer_helper_init(&rx->bh, N, (int)rx->L, P, rx->T_sel);
    Window probe{};
    rx->fused = pfb_fused_supported(N, P, rx->T_sel, probe);
    rx->kernel_name = pfb_kernel_name(N, P, rx->T_sel);
    if (rx->fused) {
        auto tw = fused_twiddles();
        if (dev_upload(&rx->d_tw, tw.data(), tw.size())) return -1;
        std::vector<uint8_t> perm(2048);
        pfb_gather_coloring(all_bins ? nullptr : rx->bins_host.data(), rx->T_sel, perm.data());
        if (dev_upload(&rx->d_xperm, perm.data(), perm.size())) return -1;
    } else {
        auto tw = generic_twiddles(N);
        if (dev_upload(&rx->d_tw, tw.data(), tw.size())) return -1;
        // Any channel count the fused kernel does not cover (the pyUSRP client sets fft_tones to its decimation factor,
        // pyUSRP/USRP_noise.py:509-521): the channelizer IS a bank of decimating complex FIRs,
        //     out[b][u] = sum_{i<P} sum_{k<N} x[(b+i)N + k] * ( w[iN+k] e^{-2 pi j k bin_u / N} ),
        // i.e. DIRECT with decim = N, pf_average = P, per-tone taps g_u[m] = w[m] e^{-2 pi j (m bin_u mod N)/N} and no LO
        // rotation (whole turns per frame) -- direct_fir_tc_kernel runs it as a 3xTF32 GEMM on the tensor cores instead of
        // an O(N T) DFT on the CUDA cores.  GSDR_PFB_VARIANT=generic keeps the CUDA-core pair (cross-checks).
        // GSDR_PFB_VARIANT: i8 (default: exact integer MMAs) | tc (3xTF32) | generic (CUDA-core FIR + DFT pair, cross-checks)
        const char* e = getenv("GSDR_PFB_VARIANT");
        const bool want_generic = e && !strcmp(e, "generic"), want_tc = e && !strcmp(e, "tc");
        const size_t n_g = (size_t)rx->T_sel * N * P;
        if (!want_generic && n_g <= (size_t)32 << 20 && direct_fir_tc_supported(rx->T_sel, N, N * P, 1)) {
            std::vector<double> gd(2 * n_g);
            const double two_pi = 6.283185307179586476925286766559;
            for (int u = 0; u < rx->T_sel; ++u) {
                const long long bin = all_bins ? u : rx->bins_host[u];
                for (int m = 0; m < N * P; ++m) {
                    const double a = -two_pi * (double)((bin * m) % N) / (double)N;
                    const double h = rx->taps_host[m];
                    gd[2 * ((size_t)u * N * P + m)] = h * std::cos(a);
                    gd[2 * ((size_t)u * N * P + m) + 1] = h * std::sin(a);
                }
            }
            if (!want_tc && direct_fir_i8_supported(rx->T_sel, N, N * P, 1)) {
                if (direct_i8_bank_create(gd.data(), rx->T_sel, N, N * P, &rx->i8bank)) return -1;
                rx->pfb_tc = rx->pfb_i8 = true;   // pfb_tc: "the channelizer runs as a decimating filter bank GEMM"
                rx->kernel_name = "direct_fir_i8_kernel (filter bank as GEMM)";
            } else {
                std::vector<float2> g(n_g);
                for (size_t k = 0; k < n_g; ++k) g[k] = make_float2((float)gd[2 * k], (float)gd[2 * k + 1]);
                if (dev_upload(&rx->d_g, g.data(), g.size())) return -1;
                rx->pfb_tc = true;
                rx->kernel_name = "direct_fir_tc_kernel (filter bank as GEMM)";
            }
        }
    }
    rx->hist_cap = (long long)N * P + 16;
    rx->max_out = (size_t)rx->T_sel * rx->batching;
    return 0;
}

int init_chirp(gsdr_rx* rx, const gsdr_param* p) {
    if (p->n_chirp_t < 1 || p->n_chirp_f < 1 || p->n_swipe_s < 1 || p->n_freq < 1) {
        set_error("CHIRP: chirp_t/chirp_f/swipe_s/freq must each hold one value");
        return -1;
    }
    chirp_params(rx->rate, p->freq[0], p->chirp_f[0], p->swipe_s[0], p->chirp_t[0], false, &rx->cpar);
    if (rx->cpar.num_steps * rx->cpar.length == 0) {
        set_error("CHIRP: empty sweep (num_steps*length == 0)");
        return -1;
    }
    rx->cdev.period = rx->cpar.num_steps * rx->cpar.length;
    rx->cdev.length = rx->cpar.length;
    rx->cdev.chirpness = rx->cpar.chirpness;
    rx->cdev.f0 = rx->cpar.f0;
    rx->cdev.num_steps = rx->cpar.num_steps;
    rx->last_index = 0;
    rx->kernel_name = "chirp_demod_full_kernel";
    rx->max_out = (size_t)rx->L;
    if (rx->decim > 0) {
        const long long ppt = (long long)rx->cpar.length * rx->decim;  // cpp/USRP_demodulator.cpp:231
        if (ppt <= 0 || ppt > 0x7fffffffLL) {
            set_error("CHIRP: points per tone %lld out of range", ppt);
            return -1;
        }
        rx->ppt = (int)ppt;
        vna_helper_init(&rx->vh, rx->ppt, (int)rx->L);
        std::vector<float> prof(rx->ppt);
        make_flat_window(rx->ppt, rx->ppt / 10, prof.data());  // :246
        if (dev_upload(&rx->d_profile, prof.data(), prof.size())) return -1;
        rx->taps_host = prof;
        rx->hist_cap = rx->ppt + 16;
        rx->max_out = (size_t)(rx->L / rx->ppt + 2);
        rx->kernel_name = rx->ppt < 32 ? "chirp_lockin_thread_kernel" : "chirp_lockin_warp_kernel";
    }
    return 0;
}

// mean of `decim` consecutive channelizer rows after the fused kernel (NOISE: every bin; TONES: the selected tones)
int init_post_decim(gsdr_rx* rx) {
    const size_t n = sizeof(float2) * 2 * (size_t)rx->T_sel;
    if (cudaMalloc(&rx->d_spec_acc, n) != cudaSuccess || cudaMemset(rx->d_spec_acc, 0, n) != cudaSuccess) {
        set_error("cudaMalloc for the post-PFB accumulator failed");
        return -1;
    }
    rx->post_decim = true;
    rx->kernel_name_buf = std::string(rx->kernel_name) + " + spectra_decimate_kernel";
    rx->kernel_name = rx->kernel_name_buf.c_str();
    return 0;
}

int init_direct(gsdr_rx* rx) {
    const int T = rx->T;
    if ((int)rx->freq.size() < T) {
        set_error("DIRECT: %d wave types but only %zu frequencies", T, rx->freq.size());
        return -1;
    }
    if (dev_upload(&rx->d_freq, rx->freq.data(), (size_t)T)) return -1;
    rx->index_counter = 0;
    rx->kernel_name = "direct_mix_kernel";
    rx->max_out = (size_t)rx->L * T;
    if (rx->decim > 0) {
        const long long M = rx->decim, f = rx->P;
        if (rx->L % M != 0) {  // cpp/fir.cu:20 assert(nt % M == 0)
            set_error("DIRECT: buffer_len %lld is not a multiple of decim %lld", rx->L, M);
            return -1;
        }
        if (f < 1) {
            set_error("DIRECT: pf_average must be >= 1");
            return -1;
        }
        rx->ntaps = (int)(M * f);
        rx->taps_host.resize(rx->ntaps);
        make_sinc_window(rx->ntaps, (float)(0.75 / (M * 2)), rx->taps_host.data());  // cpp/USRP_demodulator.cpp:99
        // g[ch][m] = h[m] * exp(-2 pi j ((tf*m) mod R)/R): per-tone complex FIR, built in double
        std::vector<double> gd(2 * (size_t)T * rx->ntaps);
        std::vector<float2> g((size_t)T * rx->ntaps);
        const double two_pi = 6.283185307179586476925286766559;
        for (int ch = 0; ch < T; ++ch)
            for (int m = 0; m < rx->ntaps; ++m) {
                long long ph = ((long long)rx->freq[ch] * (long long)(m % rx->rate)) % rx->rate;
                const double a = -two_pi * (double)ph / (double)rx->rate;
                const double h = rx->taps_host[m];
                const size_t k = (size_t)ch * rx->ntaps + m;
                gd[2 * k] = h * std::cos(a), gd[2 * k + 1] = h * std::sin(a);
                g[k] = make_float2((float)gd[2 * k], (float)gd[2 * k + 1]);
            }
        if (dev_upload(&rx->d_g, g.data(), g.size())) return -1;
        rx->hist_cap = (f - 1) * M + 16;
        rx->n_hist = (f - 1) * M;  // FIR history starts as zeros (cpp/fir.cu:23-26)
        rx->max_out = (size_t)(rx->L / M) * T;
        rx->kernel_name = (rx->P >= 1 && rx->P <= 8) ? "direct_fir_tiled_kernel" : "direct_fir_kernel";
        // Tensor-core paths: the exact integer GEMM (direct_i8_kernels.cu) is the default when the shape fills the GPU;
        // GSDR_DIRECT_VARIANT = i8 | tc (3xTF32, direct_tc_kernels.cu) | fp32 (CUDA cores) forces one.
        const long long n_out1 = rx->L / M;
        const char* e = getenv("GSDR_DIRECT_VARIANT");
        rx->direct_i8 = direct_fir_i8_preferred(T, (int)M, rx->ntaps, n_out1);
        if (e && !strcmp(e, "i8")) rx->direct_i8 = direct_fir_i8_supported(T, (int)M, rx->ntaps, n_out1);
        if (e && !strcmp(e, "tc")) rx->direct_i8 = false, rx->direct_tc = direct_fir_tc_supported(T, (int)M, rx->ntaps, n_out1);
        if (e && !strcmp(e, "fp32")) rx->direct_i8 = rx->direct_tc = false;
        if (rx->direct_i8) {
            if (direct_i8_bank_create(gd.data(), T, (int)M, rx->ntaps, &rx->i8bank)) return -1;
            rx->kernel_name = "direct_fir_i8_kernel";
        }
        if (rx->direct_tc) rx->kernel_name = "direct_fir_tc_kernel";
    }
    return 0;
}

// Enqueue the compute for n_buf consecutive buffers at d_in on s_comp.  Fills lens (per buffer) and
// returns the total valid float2 count.
long long enqueue_compute_impl(gsdr_rx* rx, const float2* d_in, int n_buf, float2* d_out, int* lens);
long long enqueue_compute(gsdr_rx* rx, const float2* d_in, int n_buf, float2* d_out, int* lens) {
    const RxState before = save_state(rx);
    const long long total = enqueue_compute_impl(rx, d_in, n_buf, d_out, lens);
    if (total < 0) restore_state(rx, before);
    return total;
}
long long enqueue_compute_impl(gsdr_rx* rx, const float2* d_in, int n_buf, float2* d_out, int* lens) {
    cudaStream_t st = rx->s_comp;
    const long long L = rx->L;
    Window w{rx->hist[rx->hist_cur], d_in, rx->n_hist, L * n_buf};
    long long total = 0;
    switch (rx->mode) {
        case GSDR_TONES:
        case GSDR_NOISE: {
            long long frames = 0;
            // NOISE: spectral decimation (decimate_spectra); TONES: the post-PFB decimator (decimate_pfb, cpp/USRP_demodulator.cpp:520-545).
            // Both are the mean of `decim` consecutive rows of the channelizer output; row width = T_sel.
            const bool spec_decim = rx->post_decim;
            long long groups = 0;
            int carried = rx->spec_carried;
            for (int b = 0; b < n_buf; ++b) {
                int v = rx->T_sel * rx->bh.current_batch;
                if (spec_decim) {  // complete groups of `decim` spectra that end inside this buffer
                    const long long gb = ((long long)carried + rx->bh.current_batch) / rx->decim;
                    carried = (int)((long long)carried + rx->bh.current_batch - gb * rx->decim);
                    groups += gb;
                    v = (int)(gb * rx->T_sel);
                }
                if (lens) lens[b] = v;
                frames += rx->bh.current_batch;
                total += v;
                buffer_helper_update(&rx->bh);
            }
            float2* const final_out = d_out;
            if (spec_decim) {
                const size_t need = sizeof(float2) * (size_t)(frames > 0 ? frames : 1) * rx->T_sel;
                if (need > rx->spec_bytes) {
                    if (rx->d_spec) cudaFree(rx->d_spec);
                    rx->d_spec = nullptr;
                    if (cudaMalloc(&rx->d_spec, need) != cudaSuccess) {
                        set_error("cudaMalloc(%zu) for the spectra buffer failed", need);
                        return -1;
                    }
                    rx->spec_bytes = need;
                }
                d_out = rx->d_spec;
            }
            const long long tail = rx->bh.new_0;  // spare samples to carry into the next call
            if (w.n_hist + w.n_in - frames * rx->N != tail) {
                set_error("internal: PFB carry-over mismatch (%lld vs %lld)", w.n_hist + w.n_in - frames * rx->N, tail);
                return -1;
            }
            PfbJob job{w, rx->d_taps, rx->d_bins, d_out, 0, (int)frames, rx->N, (int)rx->P, rx->T_sel, rx->d_xperm};
            if (!rx->pfb_tc) {   // the channelizer launch carries the carry-over copy (fused kernel: inside the kernel)
                job.tail_dst = rx->hist[rx->hist_cur ^ 1];
                job.tail_n = tail;
            }
            job.in_sc16 = rx->in_sc16 ? 1 : 0;
            if (rx->host_window) {
                // Tile length when the window is read over PCIe: the largest power of two that still leaves six CTAs, at most
                // 128 frames.  Measured (tools/process_latency.py, B200, us per buffer): 488 frames (buffer_len 1e6): 4 (one
                // tile per SM) 328, 8: 261, 16: 227, 32: 205, 48: 199, 61: 206-208, 64: 195, 80: 199, 96: 216, 128: 257;
                // 2929 frames (6e6): 64: 1044, 128: 1018, 256: 1018; 97 frames (2e5): 13: 73, 16: 65, 32: 84.
                long long tile = 1;
                while (tile * 2 * 6 <= frames && tile < 128) tile *= 2;
                job.min_tile = (int)tile;
            }
            if (!rx->fused && !rx->pfb_tc) {
                const size_t need = pfb_workspace_bytes(rx->N, (int)rx->P, (int)frames);
                if (need > rx->work_bytes) {
                    if (rx->d_work) cudaFree(rx->d_work);
                    rx->d_work = nullptr;
                    if (cudaMalloc(&rx->d_work, need) != cudaSuccess) {
                        set_error("cudaMalloc(%zu) for the PFB workspace failed", need);
                        return -1;
                    }
                    rx->work_bytes = need;
                }
            }
            const int nl = rx->pfb_i8   ? direct_fir_i8_launch(rx->i8bank, w, nullptr, rx->T_sel, rx->N, rx->N * (int)rx->P, 1, 0, frames, d_out,
                                                               rx->sm_count, st, /*rotate=*/0, tc_tma_allowed(rx))
                           : rx->pfb_tc ? direct_fir_tc_launch(w, rx->d_g, nullptr, rx->T_sel, rx->N, rx->N * (int)rx->P, 1, 0, frames, d_out,
                                                               rx->sm_count, st, /*rotate=*/0, tc_tma_allowed(rx))
                                        : pfb_launch(&job, 1, nullptr, rx->d_work, rx->d_tw, rx->sm_count, st);
            if (nl < 0) return -1;
            rx->launches += nl;
            if (spec_decim) {
                const int dl = spectra_decimate_launch(rx->d_spec, frames, rx->T_sel, rx->decim, rx->spec_carried, rx->d_spec_acc, final_out,
                                                       groups, rx->sm_count, st);
                if (dl < 0) return -1;
                rx->launches += dl;
                rx->spec_carried = carried;
            }
            if (rx->pfb_tc) {
                const int tl = window_tail_copy(w, tail, rx->hist[rx->hist_cur ^ 1], st);
                if (tl < 0) return -1;
                rx->launches += tl;
            }
            rx->hist_cur ^= 1;
            rx->n_hist = tail;
            break;
        }
        case GSDR_CHIRP: {
            if (rx->decim <= 0) {
                const int nl = chirp_demod_full_launch(d_in, L * n_buf, rx->last_index, rx->cdev, d_out, st);
                if (nl < 0) return -1;
                rx->launches += nl;
                for (int b = 0; b < n_buf; ++b)
                    if (lens) lens[b] = (int)L;
                total = L * n_buf;
            } else {
                long long n_out = 0, tail = 0;
                for (int b = 0; b < n_buf; ++b) {
                    if (lens) lens[b] = rx->vh.valid_size;
                    n_out += rx->vh.valid_size;
                    tail = rx->vh.new0;  // remainder of THIS buffer (the reference reads new0 before update())
                    vna_helper_update(&rx->vh);
                }
                if (w.n_hist + w.n_in - n_out * rx->ppt != tail) {
                    set_error("internal: chirp carry-over mismatch");
                    return -1;
                }
                const size_t pc = chirp_partial_count(rx->ppt, n_out, rx->sm_count);
                if (pc > rx->partial_cap) {
                    if (rx->d_partial) cudaFree(rx->d_partial);
                    rx->d_partial = nullptr;
                    if (cudaMalloc(&rx->d_partial, pc * sizeof(float2)) != cudaSuccess) {
                        set_error("cudaMalloc for chirp partial sums failed");
                        return -1;
                    }
                    rx->partial_cap = pc;
                }
                // chirp position of window sample 0 = position of the first new sample minus the carry
                const unsigned long long per = rx->cdev.period;
                const unsigned long long pos0 = (rx->last_index + per - ((unsigned long long)w.n_hist % per)) % per;
                const int side = rx->ppt / 10;  // make_flat_window(ppt, ppt/10): taps [0,side) are zero, the rest equal
                const int nl = chirp_demod_launch(w, pos0, rx->cdev, rx->d_profile, side, rx->taps_host[rx->ppt - 1], rx->ppt,
                                                  n_out, d_out, rx->d_partial, rx->sm_count, st);
                if (nl < 0) return -1;
                rx->launches += nl;
                const int tl = window_tail_copy(w, tail, rx->hist[rx->hist_cur ^ 1], st);
                if (tl < 0) return -1;
                rx->launches += tl;
                rx->hist_cur ^= 1;
                rx->n_hist = tail;
                total = n_out;
            }
            rx->last_index = (rx->last_index + (unsigned long long)(L * n_buf)) % rx->cdev.period;
            break;
        }
        case GSDR_DIRECT: {
            if (rx->decim <= 0) {
                const int nl = direct_mix_launch(d_in, L * n_buf, rx->d_freq, rx->T, rx->rate, rx->index_counter, d_out, st);
                if (nl < 0) return -1;
                rx->launches += nl;
                for (int b = 0; b < n_buf; ++b)
                    if (lens) lens[b] = (int)(L * rx->T);
                total = L * n_buf * rx->T;
            } else {
                const long long M = rx->decim;
                const long long n_out = (L / M) * n_buf;
                long long pos0 = (rx->index_counter - w.n_hist) % rx->rate;
                if (pos0 < 0) pos0 += rx->rate;
                const int nl = rx->direct_i8 ? direct_fir_i8_launch(rx->i8bank, w, rx->d_freq, rx->T, (int)M, rx->ntaps, rx->rate, pos0,
                                                                    n_out, d_out, rx->sm_count, st, /*rotate=*/1, tc_tma_allowed(rx))
                               : rx->direct_tc ? direct_fir_tc_launch(w, rx->d_g, rx->d_freq, rx->T, (int)M, rx->ntaps, rx->rate, pos0,
                                                                    n_out, d_out, rx->sm_count, st, /*rotate=*/1, tc_tma_allowed(rx))
                                             : direct_fir_launch(w, rx->d_g, rx->d_freq, rx->T, (int)M, rx->ntaps, rx->rate, pos0, n_out,
                                                                 d_out, st);
                if (nl < 0) return -1;
                rx->launches += nl;
                const long long tail = w.n_hist;  // (f-1)*M samples, constant
                const int tl = window_tail_copy(w, tail, rx->hist[rx->hist_cur ^ 1], st);
                if (tl < 0) return -1;
                rx->launches += tl;
                rx->hist_cur ^= 1;
                for (int b = 0; b < n_buf; ++b)
                    if (lens) lens[b] = (int)((L / M) * rx->T);
                total = n_out * rx->T;
            }
            rx->index_counter = (rx->index_counter + L * n_buf) % rx->rate;  // cpp/USRP_demodulator.cpp:437-440
            break;
        }
        default:
            set_error("mode %d has no device compute", rx->mode);
            return -1;
    }
    return total;
}

void free_all(gsdr_rx* rx) {
    cudaSetDevice(rx->device);
    for (auto& s : rx->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_raw) cudaFree(s.d_raw);
        if (s.in_done) cudaEventDestroy(s.in_done);
        if (s.comp_done) cudaEventDestroy(s.comp_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    for (int i = 0; i < 2; ++i)
        if (rx->hist[i]) cudaFree(rx->hist[i]);
    if (rx->d_taps) cudaFree(rx->d_taps);
    if (rx->d_bins) cudaFree(rx->d_bins);
    if (rx->d_xperm) cudaFree(rx->d_xperm);
    if (rx->d_tw) cudaFree(rx->d_tw);
    if (rx->d_work) cudaFree(rx->d_work);
    for (auto e : rx->ev_chunk_in) cudaEventDestroy(e);
    for (auto e : rx->ev_chunk_comp) cudaEventDestroy(e);
    if (rx->d_spec) cudaFree(rx->d_spec);
    if (rx->d_spec_acc) cudaFree(rx->d_spec_acc);
    if (rx->d_profile) cudaFree(rx->d_profile);
    if (rx->d_partial) cudaFree(rx->d_partial);
    if (rx->d_g) cudaFree(rx->d_g);
    direct_i8_bank_destroy(&rx->i8bank);
    if (rx->d_freq) cudaFree(rx->d_freq);
    if (rx->t0) cudaEventDestroy(rx->t0);
    if (rx->t1) cudaEventDestroy(rx->t1);
    if (rx->s_in) cudaStreamDestroy(rx->s_in);
    if (rx->s_comp) cudaStreamDestroy(rx->s_comp);
    if (rx->s_out) cudaStreamDestroy(rx->s_out);
}

constexpr int kDepth = 3;

}  // namespace

extern "C" {

gsdr_rx* gsdr_rx_create(const gsdr_param* p, int device, int diagnostic) {
    if (!p) {
        set_error("gsdr_rx_create: null parameters");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("gsdr_rx_create: no CUDA device available (this library has no CPU path)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        set_error("gsdr_rx_create: device %d out of range (0..%d)", device, ndev - 1);
        return nullptr;
    }
    // mode dispatch: cpp/USRP_demodulator.cpp:15-39
    int mode = GSDR_NODSP;
    if (p->n_wave_type > 0) mode = p->wave_type[0];
    int chirps = 0;
    for (uint64_t i = 0; i < p->n_wave_type; ++i) {
        if (p->wave_type[i] != mode) {
            set_error("Mixed RX buffer demodulation has been requested. This feature is not implemented yet.");
            return nullptr;
        }
        if (p->wave_type[i] == GSDR_CHIRP) ++chirps;
    }
    if (chirps > 1) {
        set_error("Multiple chirp RX buffer demodulation has been requested. This feature is not implemented yet.");
        return nullptr;
    }
    if (p->buffer_len == 0 || p->buffer_len > 0x7fffffffULL) {
        set_error("gsdr_rx_create: buffer_len %llu out of range", (unsigned long long)p->buffer_len);
        return nullptr;
    }
    std::unique_ptr<gsdr_rx> rx(new gsdr_rx());
    rx->device = device;
    rx->mode = mode;
    rx->diagnostic = diagnostic != 0;
    {   // process-wide switches are read here, once, never on the per-buffer path
        const char* zc = getenv("GSDR_PROCESS_ZEROCOPY");
        rx->zc_enabled = !(zc && zc[0] == '0');
        const char* tm = getenv("GSDR_DIRECT_TC_TMA");        // =0: every tile through the register path (tests)
        rx->tc_tma = !(tm && tm[0] == '0');
        const char* ht = getenv("GSDR_DIRECT_TC_HOST_TMA");   // =1: TMA boxes fetched from pinned host memory
        rx->tc_host_tma = ht && ht[0] == '1';
    }
    rx->rate = p->rate;
    rx->N = p->fft_tones;
    rx->L = (long long)p->buffer_len;
    rx->decim = (long long)p->decim;
    rx->P = (long long)p->pf_average;
    rx->T = (int)p->n_wave_type;
    rx->freq.assign(p->freq, p->freq + p->n_freq);
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) rx->sm_count = prop.multiProcessorCount;

    int rc = 0;
    switch (mode) {
        case GSDR_TONES:
            if ((int)rx->freq.size() < rx->T) {
                set_error("TONES: %d wave types but only %zu frequencies", rx->T, rx->freq.size());
                return nullptr;
            }
            rc = init_pfb(rx.get(), false);
            // Post-PFB decimator (decimate_pfb / accumulate_ffts, cpp/kernels.cu:754-790): the reference's kernel reads
            // input[j * (offset % nfft)] -- not the j-th frame -- so its output is not a function worth reproducing.  Built
            // here: what it is meant to be, out[g][t] = mean_{j<decim} X[g decim + j][bin_t], groups running across
            // buffer boundaries (decim == 1 is the identity).
            if (rc == 0 && rx->decim > 1) rc = init_post_decim(rx.get());
            break;
        case GSDR_NOISE:
            rc = init_pfb(rx.get(), true);
            // spectral decimation (cpp/USRP_demodulator.cpp:596-624): running sum + staging row
            if (rc == 0 && rx->decim > 0) rc = init_post_decim(rx.get());
            break;
        case GSDR_CHIRP:
            rc = init_chirp(rx.get(), p);
            break;
        case GSDR_DIRECT:
            if (rx->rate <= 0) {
                set_error("DIRECT: rate must be positive");
                return nullptr;
            }
            rc = init_direct(rx.get());
            break;
        case GSDR_NODSP:
            rx->max_out = (size_t)rx->L;
            rx->kernel_name = "none (host copy)";
            break;
        default:
            set_error("Void demodulation operation has not been implemented yet!");
            return nullptr;
    }
    if (rc) {
        free_all(rx.get());
        return nullptr;
    }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const int prio = (mode == GSDR_CHIRP) ? hi : lo;  // cpp/USRP_demodulator.cpp:42-44,186-189
    bool ok = cudaStreamCreateWithPriority(&rx->s_in, cudaStreamNonBlocking, prio) == cudaSuccess &&
              cudaStreamCreateWithPriority(&rx->s_comp, cudaStreamNonBlocking, prio) == cudaSuccess &&
              cudaStreamCreateWithPriority(&rx->s_out, cudaStreamNonBlocking, prio) == cudaSuccess &&
              cudaEventCreate(&rx->t0) == cudaSuccess && cudaEventCreate(&rx->t1) == cudaSuccess;
    if (ok && rx->hist_cap > 0) {
        for (int i = 0; i < 2 && ok; ++i) {
            ok = cudaMalloc(&rx->hist[i], sizeof(float2) * rx->hist_cap) == cudaSuccess &&
                 cudaMemset(rx->hist[i], 0, sizeof(float2) * rx->hist_cap) == cudaSuccess;
        }
    }
    if (ok && mode != GSDR_NODSP) {
        rx->slots.resize(kDepth);
        for (auto& s : rx->slots) {
            ok = ok && cudaMalloc(&s.d_in, sizeof(float2) * rx->L) == cudaSuccess &&
                 cudaMalloc(&s.d_out, sizeof(float2) * (rx->max_out ? rx->max_out : 1)) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s.comp_done, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming) == cudaSuccess;
        }
    }
    if (!ok || cudaDeviceSynchronize() != cudaSuccess) {
        set_error("gsdr_rx_create: CUDA resource allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        free_all(rx.get());
        return nullptr;
    }
    return rx.release();
}

void gsdr_rx_destroy(gsdr_rx* rx) {
    if (!rx) return;
    cudaSetDevice(rx->device);
    if (rx->s_comp) cudaStreamSynchronize(rx->s_comp);
    if (rx->s_out) cudaStreamSynchronize(rx->s_out);
    if (rx->s_in) cudaStreamSynchronize(rx->s_in);
    free_all(rx);
    delete rx;
}

int gsdr_rx_submit(gsdr_rx* rx, const gsdr_float2* in, gsdr_float2* out, int* valid_len) {
    if (!rx || !in || !out) {
        set_error("gsdr_rx_submit: null argument");
        return -1;
    }
    if (set_dev(rx)) return -1;
    if (rx->mode == GSDR_NODSP) {  // cpp/USRP_demodulator.cpp:335-339
        std::memcpy(out, in, sizeof(float2) * rx->L);
        if (valid_len) *valid_len = (int)rx->L;
        return (int)(rx->tickets++ % 0x40000000u);
    }
    const int ticket = (int)(rx->tickets++ % 0x40000000u);
    Slot& s = rx->slots[(size_t)ticket % rx->slots.size()];
    if (s.used) {
        GSDR_CUDA_OK(cudaEventSynchronize(s.out_done));  // the slot's previous occupant must be fully done
    }
    GSDR_CUDA_OK(cudaMemcpyAsync(s.d_in, in, sizeof(float2) * rx->L, cudaMemcpyHostToDevice, rx->s_in));
    GSDR_CUDA_OK(cudaEventRecord(s.in_done, rx->s_in));
    GSDR_CUDA_OK(cudaStreamWaitEvent(rx->s_comp, s.in_done, 0));
    int len = 0;
    const long long total = enqueue_compute(rx, s.d_in, 1, s.d_out, &len);
    if (total < 0) return -1;
    GSDR_CUDA_OK(cudaEventRecord(s.comp_done, rx->s_comp));
    GSDR_CUDA_OK(cudaStreamWaitEvent(rx->s_out, s.comp_done, 0));
    if (total > 0)
        GSDR_CUDA_OK(cudaMemcpyAsync(out, s.d_out, sizeof(float2) * total, cudaMemcpyDeviceToHost, rx->s_out));
    GSDR_CUDA_OK(cudaEventRecord(s.out_done, rx->s_out));
    s.used = true;
    if (valid_len) *valid_len = len;
    return ticket;
}

int gsdr_rx_wait(gsdr_rx* rx, int ticket) {
    if (!rx) return -1;
    if (rx->mode == GSDR_NODSP) return 0;
    if (set_dev(rx)) return -1;
    Slot& s = rx->slots[(uint64_t)ticket % rx->slots.size()];
    if (!s.used) return 0;
    GSDR_CUDA_OK(cudaEventSynchronize(s.out_done));
    return 0;
}

int gsdr_rx_input_consumed(gsdr_rx* rx, int ticket) {
    if (!rx) return -1;
    if (rx->mode == GSDR_NODSP) return 1;
    Slot& s = rx->slots[(uint64_t)ticket % rx->slots.size()];
    if (!s.used) return 1;
    return cudaEventQuery(s.in_done) == cudaSuccess ? 1 : 0;
}

int gsdr_rx_pipeline_depth(const gsdr_rx* rx) { return rx ? (int)(rx->slots.empty() ? 1 : rx->slots.size()) : 0; }

// Blocking call, fused filter bank: the buffer is uploaded in kChunks pieces and every piece's frames are launched
// and downloaded as soon as their samples are on the device, so H2D, kernels and D2H of ONE buffer overlap.  The
// reference's process_pfb (cpp/USRP_demodulator.cpp:486-565) and a naive port serialise the three (8 MB up, kernels,
// 3.9 MB down); this makes the unchanged blocking drop-in call run at the speed of the slower PCIe direction.
// Results are identical to the one-launch path: same frames, same kernel, same carry-over.
static int process_pfb_chunked_impl(gsdr_rx* rx, const gsdr_float2* in, gsdr_float2* out);
static int process_pfb_chunked(gsdr_rx* rx, const gsdr_float2* in, gsdr_float2* out) {
    const RxState before = save_state(rx);
    const int len = process_pfb_chunked_impl(rx, in, out);
    if (len < 0) restore_state(rx, before);
    return len;
}
static int process_pfb_chunked_impl(gsdr_rx* rx, const gsdr_float2* in, gsdr_float2* out) {
    // Upload chunks per call.  More chunks shorten the tail (the last chunk's kernel + download) but cost ~6 API calls each,
    // and on this pool's hosts those dominate: measured per 1e6-sample buffer (8 MB up, 3.9 MB down) 1 chunk 259 us, 2 chunks
    // 228 us, 4 chunks 242 us, 8 chunks 272 us, 16 chunks 316 us.  GSDR_PROCESS_CHUNKS overrides.  (Storing the tones straight
    // into the mapped pinned output buffer instead of copying them was tried: same times -- the call is bound by the upload
    // plus the cross-stream hand-offs, not by the number of API calls.)
    static const int kChunks = [] {
        const char* e = getenv("GSDR_PROCESS_CHUNKS");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 32) ? v : 2;
    }();
    const long long L = rx->L, N = rx->N, P = rx->P;
    Slot& s = rx->slots[0];
    if (rx->ev_chunk_in.empty()) {
        rx->ev_chunk_in.resize(kChunks);
        rx->ev_chunk_comp.resize(kChunks);
        for (int c = 0; c < kChunks; ++c) {
            GSDR_CUDA_OK(cudaEventCreateWithFlags(&rx->ev_chunk_in[c], cudaEventDisableTiming));
            GSDR_CUDA_OK(cudaEventCreateWithFlags(&rx->ev_chunk_comp[c], cudaEventDisableTiming));
        }
    }
    const Window w{rx->hist[rx->hist_cur], s.d_in, rx->n_hist, L};
    const long long frames = rx->bh.current_batch;
    const int len = rx->T_sel * rx->bh.current_batch;
    buffer_helper_update(&rx->bh);
    const long long tail = rx->bh.new_0;
    if (w.n_hist + w.n_in - frames * N != tail) {
        set_error("internal: PFB carry-over mismatch (%lld vs %lld)", w.n_hist + w.n_in - frames * N, tail);
        return -1;
    }
    const long long Lc = ((L / kChunks) + 1) & ~1LL;  // even: chunk starts stay 16-byte aligned
    long long done = 0;
    for (int c = 0; c < kChunks; ++c) {
        const long long off = c * Lc;
        if (off >= L) break;
        const long long n = (c == kChunks - 1 || off + Lc > L) ? L - off : Lc;
        GSDR_CUDA_OK(cudaMemcpyAsync(s.d_in + off, reinterpret_cast<const float2*>(in) + off, sizeof(float2) * n, cudaMemcpyHostToDevice,
                                     rx->s_in));
        GSDR_CUDA_OK(cudaEventRecord(rx->ev_chunk_in[c], rx->s_in));
        const long long avail = rx->n_hist + off + n;  // window samples on the device once this chunk has landed
        long long fa = avail / N - (P - 1);
        if (fa < 0) fa = 0;
        if (fa > frames || off + n >= L) fa = frames;
        GSDR_CUDA_OK(cudaStreamWaitEvent(rx->s_comp, rx->ev_chunk_in[c], 0));
        if (fa > done) {
            PfbJob job{w, rx->d_taps, rx->d_bins, s.d_out + done * rx->T_sel, (int)done, (int)(fa - done), rx->N, (int)rx->P, rx->T_sel,
                       rx->d_xperm};
            const int nl = pfb_launch(&job, 1, nullptr, nullptr, rx->d_tw, rx->sm_count, rx->s_comp);
            if (nl < 0) return -1;
            rx->launches += nl;
            GSDR_CUDA_OK(cudaEventRecord(rx->ev_chunk_comp[c], rx->s_comp));
            GSDR_CUDA_OK(cudaStreamWaitEvent(rx->s_out, rx->ev_chunk_comp[c], 0));
            GSDR_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<float2*>(out) + done * rx->T_sel, s.d_out + done * rx->T_sel,
                                         sizeof(float2) * (size_t)(fa - done) * rx->T_sel, cudaMemcpyDeviceToHost, rx->s_out));
            done = fa;
        }
    }
    const int tl = window_tail_copy(w, tail, rx->hist[rx->hist_cur ^ 1], rx->s_comp);  // after the last chunk's wait on s_comp
    if (tl < 0) return -1;
    rx->launches += tl;
    rx->hist_cur ^= 1;
    rx->n_hist = tail;
    GSDR_CUDA_OK(cudaEventRecord(s.in_done, rx->s_in));
    GSDR_CUDA_OK(cudaEventRecord(s.comp_done, rx->s_comp));
    GSDR_CUDA_OK(cudaEventRecord(s.out_done, rx->s_out));
    rx->tickets++;
    GSDR_CUDA_OK(cudaEventSynchronize(s.out_done));
    GSDR_CUDA_OK(cudaEventSynchronize(s.comp_done));  // carry-over written, input fully consumed
    return len;
}

// Blocking call, zero-copy form (the default when both buffers are pinned and mapped; GSDR_PROCESS_ZEROCOPY=0 at create turns
// it off): the kernels' loads read the caller's pinned input buffer over PCIe and their stores write the result into the
// caller's pinned output buffer, so both PCIe directions run concurrently under the launch itself and no copy-engine call or
// cross-stream hand-off is made.  Same kernels, same frames, same carry-over as every other entry point (outputs
// bit-identical).  Fused channelizer, measured per 1e6-sample buffer on B200 (tools/process_latency.py,
// profiles/r1_process_latency.jsonl): chunked copies 229 us; zero-copy with one 4-frame tile per SM 328 us (reads from host
// memory are not kept in L2, so the three halo rows of every tile cross PCIe again: 1.75x the bytes); with 8 / 16 / 32 / 64 /
// 128-frame tiles 261 / 227 / 205 / 195 / 257 us (128: four CTAs no longer keep the read pipe full).
static int process_zerocopy(gsdr_rx* rx, const float2* in_alias, float2* out_alias) {
    int len = 0;
    rx->host_window = true;
    const long long total = enqueue_compute(rx, in_alias, 1, out_alias, &len);
    rx->host_window = false;
    if (total < 0) return -1;
    rx->tickets++;
    GSDR_CUDA_OK(cudaStreamSynchronize(rx->s_comp));
    return len;
}

int gsdr_rx_process(gsdr_rx* rx, const gsdr_float2* in, gsdr_float2* out) {
    if (!rx || !in || !out) {
        set_error("gsdr_rx_process: null argument");
        return -1;
    }
    if (rx->mode == GSDR_NODSP || rx->slots.empty()) {
        int len = 0;
        const int ticket = gsdr_rx_submit(rx, in, out, &len);
        return ticket < 0 ? -1 : len;
    }
    if (set_dev(rx)) return -1;
    // keep the ticket/slot rotation of submit() intact: the blocking paths leave every slot idle
    for (auto& sl : rx->slots)
        if (sl.used) {
            GSDR_CUDA_OK(cudaEventSynchronize(sl.out_done));
            sl.used = false;
        }
    const bool pfb = (rx->mode == GSDR_TONES || rx->mode == GSDR_NOISE);
    const bool pfb_fused_call = pfb && rx->fused && rx->L >= 64LL * rx->N && !rx->post_decim;
    // one-launch form: the fused channelizer (measured faster than the copied form), DIRECT and CHIRP (one pass over the input,
    // so reading it over PCIe costs what the upload would, and the kernels, the download and ~10 API calls overlap with it)
    if (rx->zc_enabled && (pfb_fused_call || rx->mode == GSDR_DIRECT || rx->mode == GSDR_CHIRP)) {
        void* ia = host_alias_of(in, sizeof(float2) * (size_t)rx->L);
        void* oa = ia ? host_alias_of(out, sizeof(float2) * (rx->max_out ? rx->max_out : 1)) : nullptr;
        if (ia && oa) return process_zerocopy(rx, static_cast<const float2*>(ia), static_cast<float2*>(oa));
    }
    if (pfb_fused_call) return process_pfb_chunked(rx, in, out);
    int len = 0;
    const int ticket = gsdr_rx_submit(rx, in, out, &len);
    if (ticket < 0) return -1;
    if (gsdr_rx_wait(rx, ticket)) return -1;
    return len;
}

int gsdr_rx_submit_sc16(gsdr_rx* rx, const int16_t* in_iq, gsdr_float2* out, int* valid_len) {
    if (!rx || !in_iq || !out) {
        set_error("gsdr_rx_submit_sc16: null argument");
        return -1;
    }
    if (set_dev(rx)) return -1;
    if (rx->mode == GSDR_NODSP) {  // no GPU work in this mode: convert while copying
        float* o = reinterpret_cast<float*>(out);
        for (long long i = 0; i < 2 * rx->L; ++i) o[i] = (float)in_iq[i] * kSc16Scale;
        if (valid_len) *valid_len = (int)rx->L;
        return (int)(rx->tickets++ % 0x40000000u);
    }
    const int ticket = (int)(rx->tickets++ % 0x40000000u);
    Slot& s = rx->slots[(size_t)ticket % rx->slots.size()];
    if (s.used) GSDR_CUDA_OK(cudaEventSynchronize(s.out_done));
    if (!s.d_raw) GSDR_CUDA_OK(cudaMalloc(&s.d_raw, sizeof(short2) * rx->L));
    GSDR_CUDA_OK(cudaMemcpyAsync(s.d_raw, in_iq, sizeof(short2) * rx->L, cudaMemcpyHostToDevice, rx->s_in));
    GSDR_CUDA_OK(cudaEventRecord(s.in_done, rx->s_in));
    GSDR_CUDA_OK(cudaStreamWaitEvent(rx->s_comp, s.in_done, 0));
    // the fused channelizer reads the wire format itself; every other kernel gets an fc32 copy first
    const bool fused_sc16 = (rx->mode == GSDR_TONES || rx->mode == GSDR_NOISE) && rx->fused && pfb_fused_sc16_available();
    if (!fused_sc16) {
        long long blocks = ((rx->L >> 2) + 255) / 256;
        if (blocks < 1) blocks = 1;
        if (blocks > (long long)rx->sm_count * 8) blocks = (long long)rx->sm_count * 8;
        sc16_to_fc32_kernel<<<(int)blocks, 256, 0, rx->s_comp>>>(s.d_raw, rx->L, s.d_in);
        GSDR_CUDA_OK(cudaGetLastError());
        rx->launches++;
    }
    int len = 0;
    rx->in_sc16 = fused_sc16;
    const long long total = enqueue_compute(rx, fused_sc16 ? reinterpret_cast<const float2*>(s.d_raw) : s.d_in, 1, s.d_out, &len);
    rx->in_sc16 = false;
    if (total < 0) return -1;
    GSDR_CUDA_OK(cudaEventRecord(s.comp_done, rx->s_comp));
    GSDR_CUDA_OK(cudaStreamWaitEvent(rx->s_out, s.comp_done, 0));
    if (total > 0)
        GSDR_CUDA_OK(cudaMemcpyAsync(out, s.d_out, sizeof(float2) * total, cudaMemcpyDeviceToHost, rx->s_out));
    GSDR_CUDA_OK(cudaEventRecord(s.out_done, rx->s_out));
    s.used = true;
    if (valid_len) *valid_len = len;
    return ticket;
}

int gsdr_rx_process_sc16(gsdr_rx* rx, const int16_t* in_iq, gsdr_float2* out) {
    // pinned buffers + fused channelizer: one launch reads the wire-format buffer and writes the tones in place over PCIe
    if (rx && in_iq && out && rx->zc_enabled && !rx->slots.empty() && (rx->mode == GSDR_TONES || rx->mode == GSDR_NOISE) && rx->fused &&
        !rx->post_decim && rx->L >= 64LL * rx->N && pfb_fused_sc16_available()) {
        if (set_dev(rx)) return -1;
        void* ia = host_alias_of(in_iq, sizeof(short2) * (size_t)rx->L);
        void* oa = ia ? host_alias_of(out, sizeof(float2) * (rx->max_out ? rx->max_out : 1)) : nullptr;
        if (ia && oa) {
            for (auto& sl : rx->slots)
                if (sl.used) {
                    GSDR_CUDA_OK(cudaEventSynchronize(sl.out_done));
                    sl.used = false;
                }
            rx->in_sc16 = true;
            const int n = process_zerocopy(rx, static_cast<const float2*>(ia), static_cast<float2*>(oa));
            rx->in_sc16 = false;
            return n;
        }
    }
    int len = 0;
    const int ticket = gsdr_rx_submit_sc16(rx, in_iq, out, &len);
    if (ticket < 0) return -1;
    if (gsdr_rx_wait(rx, ticket)) return -1;
    return len;
}

int64_t gsdr_rx_process_device(gsdr_rx* rx, const gsdr_float2* in_dev, int n_buffers, gsdr_float2* out_dev, int* valid_lens) {
    if (!rx || !in_dev || !out_dev || n_buffers <= 0) {
        set_error("gsdr_rx_process_device: bad argument");
        return -1;
    }
    if (set_dev(rx)) return -1;
    if (rx->mode == GSDR_NODSP) {
        GSDR_CUDA_OK(cudaMemcpyAsync(out_dev, in_dev, sizeof(float2) * rx->L * n_buffers, cudaMemcpyDeviceToDevice, rx->s_comp));
        for (int b = 0; b < n_buffers; ++b)
            if (valid_lens) valid_lens[b] = (int)rx->L;
        return rx->L * n_buffers;
    }
    return enqueue_compute(rx, reinterpret_cast<const float2*>(in_dev), n_buffers, reinterpret_cast<float2*>(out_dev), valid_lens);
}

int gsdr_rx_sync(gsdr_rx* rx) {
    if (!rx) return -1;
    if (set_dev(rx)) return -1;
    GSDR_CUDA_OK(cudaStreamSynchronize(rx->s_in));
    GSDR_CUDA_OK(cudaStreamSynchronize(rx->s_comp));
    GSDR_CUDA_OK(cudaStreamSynchronize(rx->s_out));
    return 0;
}

int gsdr_rx_reset(gsdr_rx* rx) {
    if (!rx) return -1;
    if (gsdr_rx_sync(rx)) return -1;
    rx->hist_cur = 0;
    for (int i = 0; i < 2; ++i)
        if (rx->hist[i]) GSDR_CUDA_OK(cudaMemset(rx->hist[i], 0, sizeof(float2) * rx->hist_cap));
    rx->n_hist = 0;
    switch (rx->mode) {
        case GSDR_TONES:
        case GSDR_NOISE:
            buffer_helper_init(&rx->bh, rx->N, (int)rx->L, (int)rx->P, rx->T_sel);
            rx->spec_carried = 0;
            if (rx->d_spec_acc) GSDR_CUDA_OK(cudaMemset(rx->d_spec_acc, 0, sizeof(float2) * 2 * rx->T_sel));
            break;
        case GSDR_CHIRP:
            rx->last_index = 0;
            if (rx->decim > 0) vna_helper_init(&rx->vh, rx->ppt, (int)rx->L);
            break;
        case GSDR_DIRECT:
            rx->index_counter = 0;
            if (rx->decim > 0) rx->n_hist = (rx->P - 1) * rx->decim;
            break;
        default:
            break;
    }
    return 0;
}

int gsdr_rx_channels(const gsdr_rx* rx) { return rx ? rx->T : 0; }
int gsdr_rx_mode(const gsdr_rx* rx) { return rx ? rx->mode : -1; }
size_t gsdr_rx_max_output(const gsdr_rx* rx) { return rx ? rx->max_out : 0; }
size_t gsdr_rx_max_output_batch(const gsdr_rx* rx, int n) {
    if (!rx || n <= 0) return 0;
    if (rx->mode == GSDR_CHIRP && rx->decim > 0) return (size_t)((rx->L * n) / rx->ppt + 2);
    return rx->max_out * (size_t)n;
}
float gsdr_rx_fcut(const gsdr_rx* rx) { return rx ? rx->fcut : 0.f; }
uint64_t gsdr_rx_launch_count(const gsdr_rx* rx) { return rx ? rx->launches : 0; }
const char* gsdr_rx_kernel_name(const gsdr_rx* rx) { return rx ? rx->kernel_name : ""; }

int gsdr_rx_chirp_param(const gsdr_rx* rx, gsdr_chirp_param* out) {
    if (!rx || !out || rx->mode != GSDR_CHIRP) return -1;
    *out = rx->cpar;
    return 0;
}

int gsdr_rx_timer_start(gsdr_rx* rx) {
    if (!rx) return -1;
    if (set_dev(rx)) return -1;
    GSDR_CUDA_OK(cudaEventRecord(rx->t0, rx->s_comp));
    return 0;
}
int gsdr_rx_timer_stop(gsdr_rx* rx, float* ms) {
    if (!rx) return -1;
    if (set_dev(rx)) return -1;
    GSDR_CUDA_OK(cudaEventRecord(rx->t1, rx->s_comp));
    GSDR_CUDA_OK(cudaEventSynchronize(rx->t1));
    float v = 0.f;
    GSDR_CUDA_OK(cudaEventElapsedTime(&v, rx->t0, rx->t1));
    if (ms) *ms = v;
    return 0;
}

int gsdr_rx_get_taps(const gsdr_rx* rx, float* taps, size_t cap) {
    if (!rx) return -1;
    const size_t n = rx->taps_host.size();
    if (taps) std::memcpy(taps, rx->taps_host.data(), sizeof(float) * (n < cap ? n : cap));
    return (int)n;
}
int gsdr_rx_get_bins(const gsdr_rx* rx, int32_t* bins, size_t cap) {
    if (!rx) return -1;
    const size_t n = rx->bins_host.size();
    if (bins) std::memcpy(bins, rx->bins_host.data(), sizeof(int32_t) * (n < cap ? n : cap));
    return (int)n;
}

// ---- multi-stream group: one persistent launch over every member's frames ----------------------
// cfg5 as the reference would run it is one RX_buffer_demodulator per front-end, each fed by its own link thread calling
// the blocking process() (cpp/USRP_server_link_threads.cpp:605-702, call at :666).  With S streams on a GPU that is S
// uploads, S x 4 kernel launches and S downloads per packet period, each launch a fraction of a wave.  A group takes one
// buffer of every member stream per call and issues ONE channelizer launch whose tile list spans all of them:
//   device-resident  gsdr_rx_group_process_device   (inputs already in HBM)
//   host-fed         gsdr_rx_group_submit / _wait   (pinned host buffers in and out, pipelined, fc32 or sc16)
// Host-fed, copied form: one cudaMemcpyAsync per stream buffer on the copy-in stream into a contiguous staging area, the
// launch on the compute stream, one cudaMemcpyAsync per stream on the copy-out stream; three slots, so the upload of period
// k+1, the launch of period k and the download of period k-1 overlap.  Zero-copy form (every buffer pinned and mapped): the
// launch reads the S host buffers and writes the S host outputs in place -- one table upload and one launch per period.
namespace {
struct GroupSlot {
    float2* d_in = nullptr;    // [S][L]           (copied form)
    float2* d_out = nullptr;   // [sum max_out_i]  (copied form)
    short2* d_raw = nullptr;   // [S][L] sc16 staging, allocated on first use
    cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
    bool used = false;
};
// up to 64 raw sc16 buffers per conversion launch (kernel parameter block)
struct Sc16Batch {
    const short2* src[64];
    float2* dst[64];
};
__global__ void __launch_bounds__(256) sc16_to_fc32_batch_kernel(const Sc16Batch b, long long n) {
    const short2* __restrict__ in = b.src[blockIdx.y];
    float2* __restrict__ out = b.dst[blockIdx.y];
    const long long n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(in) + i);
            const int w[4] = {v.x, v.y, v.z, v.w};
            float4 o[2];
            float* of = reinterpret_cast<float*>(o);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                of[2 * k] = (float)(short)(w[k] & 0xffff) * kSc16Scale;
                of[2 * k + 1] = (float)(short)(w[k] >> 16) * kSc16Scale;
            }
            reinterpret_cast<float4*>(out)[2 * i] = o[0];
            reinterpret_cast<float4*>(out)[2 * i + 1] = o[1];
        }
    }
    for (long long i = (vec ? (n4 << 2) : 0) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const short2 v = in[i];
        out[i] = make_float2((float)v.x * kSc16Scale, (float)v.y * kSc16Scale);
    }
}
}  // namespace

struct gsdr_rx_group {
    std::vector<gsdr_rx*> members;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    void* d_table = nullptr;
    void* d_tail = nullptr;
    uint64_t launches = 0;
    int device = 0;
    // host-fed path
    std::vector<GroupSlot> slots;
    std::vector<size_t> out_off;   // member i's outputs start at d_out + out_off[i]
    size_t out_total = 0;
    long long L = 0;
    uint64_t tickets = 0;
    // Copied form: the per-stream copies of a period are spread round-robin over n_copy streams per direction (s_in / s_out
    // are streams 0).  One copy engine queue is latency-bound when eight GPUs pull on the same host memory (20 GB/s per GPU
    // measured at N = 8 against 52 alone); several queues keep more reads in flight.  GSDR_GROUP_COPY_STREAMS (default 4).
    std::vector<cudaStream_t> x_in, x_out;          // extra streams 1 .. n_copy-1
    std::vector<cudaEvent_t> j_in, j_out, f_out;    // join events (extra stream -> stream 0) and the fork event for the downloads
    int n_copy = 1;
    int zc_mode = 1;               // GSDR_GROUP_ZEROCOPY (read at create): 0 copied, 1 zero-copy both ways (default), 2 copy in / kernel
                                   // stores out, 3 = measured: see GroupAutoForm
    GroupAutoForm auto_form[2];    // [0] fc32 input, [1] sc16 input (common.hpp)
    struct AutoStamp {
        std::chrono::steady_clock::time_point t;
        int kind = -1, block = -1;  // block -1: not a measured period
        uint64_t ticket = ~0ull;
    } auto_stamp[8];
    bool last_zero_copy = false;
    int last_form = 0;             // bit 0: inputs read in place, bit 1: outputs written in place
};

namespace {
void group_free(gsdr_rx_group* g) {
    cudaSetDevice(g->device);
    for (auto& s : g->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_raw) cudaFree(s.d_raw);
        if (s.in_done) cudaEventDestroy(s.in_done);
        if (s.comp_done) cudaEventDestroy(s.comp_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    if (g->d_table) cudaFree(g->d_table);
    if (g->d_tail) cudaFree(g->d_tail);
    if (g->t0) cudaEventDestroy(g->t0);
    if (g->t1) cudaEventDestroy(g->t1);
    if (g->stream) cudaStreamDestroy(g->stream);
    if (g->s_in) cudaStreamDestroy(g->s_in);
    if (g->s_out) cudaStreamDestroy(g->s_out);
    for (auto st : g->x_in) cudaStreamDestroy(st);
    for (auto st : g->x_out) cudaStreamDestroy(st);
    for (auto e : g->j_in) cudaEventDestroy(e);
    for (auto e : g->j_out) cudaEventDestroy(e);
    for (auto e : g->f_out) cudaEventDestroy(e);
}

// One channelizer launch over n_buffers consecutive buffers of every member (in[i] / out[i]: device pointers or device
// aliases of pinned host buffers).  Member state is advanced on copies and committed only after every launch succeeded.
int64_t group_enqueue(gsdr_rx_group* g, const float2* const* in, int n_buffers, float2* const* out, int* valid_lens, bool host_window,
                      bool in_sc16 = false) {
    const int n = (int)g->members.size();
    std::vector<PfbJob> jobs(n);
    std::vector<long long> tails(n);
    std::vector<gsdr_buffer_helper> bhs(n);
    int64_t total = 0;
    long long frames_all = 0;
    for (int i = 0; i < n; ++i) {
        gsdr_rx* rx = g->members[i];
        bhs[i] = rx->bh;
        Window w{rx->hist[rx->hist_cur], in[i], rx->n_hist, rx->L * n_buffers};
        long long frames = 0;
        for (int b = 0; b < n_buffers; ++b) {
            const int v = rx->T_sel * bhs[i].current_batch;
            if (valid_lens) valid_lens[(size_t)i * n_buffers + b] = v;
            frames += bhs[i].current_batch;
            total += v;
            buffer_helper_update(&bhs[i]);
        }
        tails[i] = bhs[i].new_0;
        if (w.n_hist + w.n_in - frames * rx->N != tails[i]) {
            set_error("internal: PFB carry-over mismatch in group member %d", i);
            return -1;
        }
        frames_all += frames;
        jobs[i] = PfbJob{w, rx->d_taps, rx->d_bins, out[i], 0, (int)frames, rx->N, (int)rx->P, rx->T_sel, rx->d_xperm};
        jobs[i].tail_dst = rx->hist[rx->hist_cur ^ 1];   // the launch carries every member's carry-over copy
        jobs[i].tail_n = tails[i];
        if (host_window) jobs[i].min_tile = 64;          // reads over PCIe: long tiles (see gsdr_rx_process, zero-copy form)
        jobs[i].in_sc16 = in_sc16 ? 1 : 0;
    }
    int nl = pfb_launch(jobs.data(), n, g->d_table, nullptr, g->members[0]->d_tw, g->members[0]->sm_count, g->stream);
    if (nl < 0) return -1;
    if (n == 1 || frames_all == 0) {
        // a single member goes through the one-stream launch, which does its tail itself (pfb_launch); no frames at all
        // (buffers shorter than a frame): nothing was launched, only the carry-over grows
        if (frames_all == 0) {
            std::vector<Window> wins(n);
            std::vector<float2*> dsts(n);
            for (int i = 0; i < n; ++i) wins[i] = jobs[i].win, dsts[i] = jobs[i].tail_dst;
            const int tl = window_tail_copy_multi(wins.data(), tails.data(), dsts.data(), n, g->d_tail, g->stream);
            if (tl < 0) return -1;
            nl += tl;
        }
    }
    g->launches += nl;
    for (int i = 0; i < n; ++i) {
        gsdr_rx* rx = g->members[i];
        rx->bh = bhs[i];
        rx->hist_cur ^= 1;
        rx->n_hist = tails[i];
    }
    return total;
}
}  // namespace

gsdr_rx_group* gsdr_rx_group_create(gsdr_rx** members, int n) {
    if (!members || n <= 0) {
        set_error("gsdr_rx_group_create: no members");
        return nullptr;
    }
    for (int i = 0; i < n; ++i) {
        gsdr_rx* m = members[i];
        if (!m || (m->mode != GSDR_TONES && m->mode != GSDR_NOISE) || !m->fused || m->device != members[0]->device ||
            m->N != members[0]->N || m->P != members[0]->P || m->post_decim || m->L != members[0]->L) {
            set_error("gsdr_rx_group_create: member %d is not a fused TONES/NOISE stream compatible with member 0", i);
            return nullptr;
        }
        for (int k = 0; k < i; ++k)
            if (members[k] == m) {
                set_error("gsdr_rx_group_create: member %d appears twice", i);
                return nullptr;
            }
    }
    gsdr_rx_group* g = new gsdr_rx_group();
    g->members.assign(members, members + n);
    g->device = members[0]->device;
    g->L = members[0]->L;
    {
        const char* zc = getenv("GSDR_GROUP_ZEROCOPY");
        g->zc_mode = (zc && zc[0] >= '0' && zc[0] <= '3') ? zc[0] - '0' : 1;
        const char* cs = getenv("GSDR_GROUP_COPY_STREAMS");
        const int k = cs ? atoi(cs) : 4;
        g->n_copy = k < 1 ? 1 : (k > 8 ? 8 : k);
    }
    g->out_off.resize(n);
    for (int i = 0; i < n; ++i) {
        g->out_off[i] = g->out_total;
        g->out_total += (members[i]->max_out + 1) & ~size_t(1);   // 16-byte aligned starts
    }
    int lo = 0, hi = 0;
    bool ok = cudaSetDevice(g->device) == cudaSuccess && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&g->stream, cudaStreamNonBlocking, lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&g->s_in, cudaStreamNonBlocking, lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&g->s_out, cudaStreamNonBlocking, lo) == cudaSuccess &&
              cudaEventCreate(&g->t0) == cudaSuccess && cudaEventCreate(&g->t1) == cudaSuccess &&
              cudaMalloc(&g->d_table, pfb_table_bytes(n, members[0]->sm_count)) == cudaSuccess &&
              cudaMalloc(&g->d_tail, window_tail_multi_scratch_bytes(n)) == cudaSuccess;
    for (int k = 1; k < g->n_copy && ok; ++k) {
        cudaStream_t a = nullptr, b = nullptr;
        cudaEvent_t ea = nullptr, eb = nullptr;
        ok = cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, lo) == cudaSuccess && cudaStreamCreateWithPriority(&b, cudaStreamNonBlocking, lo) == cudaSuccess &&
             cudaEventCreateWithFlags(&ea, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&eb, cudaEventDisableTiming) == cudaSuccess;
        if (a) g->x_in.push_back(a);
        if (b) g->x_out.push_back(b);
        if (ea) g->j_in.push_back(ea);
        if (eb) g->j_out.push_back(eb);
    }
    if (!ok) {
        set_error("gsdr_rx_group_create: CUDA resource allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        group_free(g);
        delete g;
        return nullptr;
    }
    return g;
}

void gsdr_rx_group_destroy(gsdr_rx_group* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->s_in) cudaStreamSynchronize(g->s_in);
    if (g->stream) cudaStreamSynchronize(g->stream);
    if (g->s_out) cudaStreamSynchronize(g->s_out);
    group_free(g);
    delete g;
}

int64_t gsdr_rx_group_process_device(gsdr_rx_group* g, const gsdr_float2* const* in_dev, int n_buffers,
                                     gsdr_float2* const* out_dev, int* valid_lens) {
    if (!g || !in_dev || !out_dev || n_buffers <= 0) {
        set_error("gsdr_rx_group_process_device: bad argument");
        return -1;
    }
    if (ensure_device(g->device)) return -1;
    return group_enqueue(g, reinterpret_cast<const float2* const*>(in_dev), n_buffers, reinterpret_cast<float2* const*>(out_dev), valid_lens,
                         false);
}

// kDepth pipeline slots, allocated on the first host-fed call (a device-resident-only group never pays for them)
static int group_slots_ready(gsdr_rx_group* g, bool need_d_in, bool need_d_out, bool need_d_raw) {
    if (g->slots.empty()) {
        g->slots.resize(kDepth);
        for (auto& s : g->slots) {
            GSDR_CUDA_OK(cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
            GSDR_CUDA_OK(cudaEventCreateWithFlags(&s.comp_done, cudaEventDisableTiming));
            GSDR_CUDA_OK(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
        }
    }
    const size_t S = g->members.size();
    for (auto& s : g->slots) {
        if (need_d_in && !s.d_in) GSDR_CUDA_OK(cudaMalloc(&s.d_in, sizeof(float2) * S * (size_t)g->L));
        if (need_d_out && !s.d_out) GSDR_CUDA_OK(cudaMalloc(&s.d_out, sizeof(float2) * (g->out_total ? g->out_total : 1)));
        if (need_d_raw && !s.d_raw) GSDR_CUDA_OK(cudaMalloc(&s.d_raw, sizeof(short2) * S * (size_t)g->L));
    }
    return 0;
}

static int group_submit_any(gsdr_rx_group* g, const void* const* in_host, bool sc16, gsdr_float2* const* out_host, int* valid_lens) {
    if (!g || !in_host || !out_host) {
        set_error("gsdr_rx_group_submit: null argument");
        return -1;
    }
    if (ensure_device(g->device)) return -1;
    // a measured period is timed from HERE: the copied form spends a few hundred microseconds of host time enqueuing its
    // 2 S copies, during which the period is already "in flight" for the caller; stamping after the enqueues made the copied
    // form look 3 - 5 % faster than it is
    const auto t_enter = std::chrono::steady_clock::now();
    const int S = (int)g->members.size();
    const size_t in_bytes = (sc16 ? sizeof(short2) : sizeof(float2)) * (size_t)g->L;
    for (int i = 0; i < S; ++i)
        if (!in_host[i] || !out_host[i]) {
            set_error("gsdr_rx_group_submit: null buffer for member %d", i);
            return -1;
        }
    // Each direction on its own: zin = the launch reads the host input buffers in place, zout = it writes the host output
    // buffers in place; both need every buffer of the period pinned and mapped.  g->zc_mode: 0 copied both ways, 1 zero-copy
    // both ways, 2 copy engine in / kernel stores out.
    std::vector<const void*> ia(S, nullptr);
    std::vector<float2*> oa(S, nullptr);
    int mode = g->zc_mode, auto_block = -1;
    if (mode == 3) {   // measured form (GroupAutoForm, common.hpp)
        bool busy = false;   // is the previous period still in flight?
        if (g->auto_form[sc16 ? 1 : 0].choice < 0 && g->tickets > 0 && !g->slots.empty()) {
            GroupSlot& prev = g->slots[(size_t)((g->tickets - 1) % 0x40000000u) % g->slots.size()];
            if (prev.used && cudaEventQuery(prev.out_done) == cudaErrorNotReady) busy = true;
            cudaGetLastError();   // "not ready" is an answer, not an error to be found by the next launch check
        }
        GroupAutoForm& af = g->auto_form[sc16 ? 1 : 0];
        const int before = af.choice;
        mode = af.on_submit(busy, &auto_block);
        if (before < 0 && af.choice >= 0) {
            static const bool debug = [] {
                const char* e = getenv("GSDR_GROUP_AUTO_DEBUG");
                return e && e[0] == '1';
            }();
            if (debug)
                fprintf(stderr, "[gsdr_rx_group] measured form (%s input): zero-copy %.3f ms over %d periods, copied %.3f ms over %d -> %s\n",
                        sc16 ? "sc16" : "fc32", af.cnt[0] ? 1e3 * af.sum[0] / af.cnt[0] : 0.0, af.cnt[0], af.cnt[1] ? 1e3 * af.sum[1] / af.cnt[1] : 0.0,
                        af.cnt[1], af.choice ? "zero-copy" : "copied");
        }
    }
    bool zin = mode == 1, zout = mode != 0;
    for (int i = 0; i < S && zin; ++i) zin = (ia[i] = host_alias_of(in_host[i], in_bytes)) != nullptr;
    for (int i = 0; i < S && zout; ++i)
        zout = (oa[i] = static_cast<float2*>(host_alias_of(out_host[i], sizeof(float2) * g->members[i]->max_out))) != nullptr;
    // sc16: the fused channelizer converts the wire format itself (no fc32 copy exists); the lock-step cross-check kernel needs one
    const bool fused_sc16 = sc16 && pfb_fused_sc16_available();
    const bool convert = sc16 && !fused_sc16;
    // a measurement of the forms needs the staging of the copied form a few periods from now: allocate it as the measurement
    // begins (inside the periods it leaves out), not in the middle of a measured block
    const bool all_staging = g->zc_mode == 3 && g->auto_form[sc16 ? 1 : 0].measuring;
    if (group_slots_ready(g, (!zin && !sc16) || convert || (all_staging && !sc16), !zout || all_staging, sc16 && (!zin || all_staging))) return -1;
    const int ticket = (int)(g->tickets % 0x40000000u);
    GroupSlot& s = g->slots[(size_t)ticket % g->slots.size()];
    if (s.used) GSDR_CUDA_OK(cudaEventSynchronize(s.out_done));   // the slot's previous period is fully done
    std::vector<const float2*> kin(S);
    std::vector<float2*> kout(S);
    for (int i = 0; i < S; ++i) {
        if (!zin) {   // one cudaMemcpyAsync per stream buffer
            void* dst = sc16 ? static_cast<void*>(s.d_raw + (size_t)i * g->L) : static_cast<void*>(s.d_in + (size_t)i * g->L);
            const int k = i % g->n_copy;
            GSDR_CUDA_OK(cudaMemcpyAsync(dst, in_host[i], in_bytes, cudaMemcpyHostToDevice, k == 0 ? g->s_in : g->x_in[k - 1]));
        }
        if (convert) kin[i] = s.d_in + (size_t)i * g->L;
        else if (zin) kin[i] = static_cast<const float2*>(ia[i]);
        else kin[i] = sc16 ? reinterpret_cast<const float2*>(s.d_raw + (size_t)i * g->L) : s.d_in + (size_t)i * g->L;
        kout[i] = zout ? oa[i] : s.d_out + g->out_off[i];
    }
    if (!zin)   // the extra upload queues join stream 0 (before the conversion launch, when there is one)
        for (size_t k = 0; k < g->x_in.size(); ++k) {
            GSDR_CUDA_OK(cudaEventRecord(g->j_in[k], g->x_in[k]));
            GSDR_CUDA_OK(cudaStreamWaitEvent(g->s_in, g->j_in[k], 0));
        }
    if (convert) {
        // wire format -> fc32 on the copy-in stream, one launch per 64 streams (reading the host buffers in place when zin)
        long long blocks = ((g->L >> 2) + 255) / 256;
        blocks = blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks);
        for (int i0 = 0; i0 < S; i0 += 64) {
            Sc16Batch b{};
            const int nb = S - i0 < 64 ? S - i0 : 64;
            for (int k = 0; k < nb; ++k) {
                b.src[k] = zin ? static_cast<const short2*>(ia[i0 + k]) : s.d_raw + (size_t)(i0 + k) * g->L;
                b.dst[k] = s.d_in + (size_t)(i0 + k) * g->L;
            }
            sc16_to_fc32_batch_kernel<<<dim3((unsigned)blocks, (unsigned)nb), 256, 0, g->s_in>>>(b, g->L);
            GSDR_CUDA_OK(cudaGetLastError());
            g->launches++;
        }
    }
    const bool staged = !zin || convert;   // the inputs reach the device on the copy-in stream
    if (staged) {
        GSDR_CUDA_OK(cudaEventRecord(s.in_done, g->s_in));   // the inputs are on the device: the host buffers are free
        GSDR_CUDA_OK(cudaStreamWaitEvent(g->stream, s.in_done, 0));
    }
    std::vector<int> lens(S);
    const int64_t total = group_enqueue(g, kin.data(), 1, kout.data(), lens.data(), zin && !convert, fused_sc16);
    if (total < 0) return -1;
    GSDR_CUDA_OK(cudaEventRecord(s.comp_done, g->stream));
    if (!staged) GSDR_CUDA_OK(cudaEventRecord(s.in_done, g->stream));   // the launch itself consumes the inputs
    if (!zout) {
        GSDR_CUDA_OK(cudaStreamWaitEvent(g->s_out, s.comp_done, 0));
        for (auto st : g->x_out) GSDR_CUDA_OK(cudaStreamWaitEvent(st, s.comp_done, 0));
        for (int i = 0; i < S; ++i)
            if (lens[i] > 0) {
                const int k = i % g->n_copy;
                GSDR_CUDA_OK(cudaMemcpyAsync(out_host[i], kout[i], sizeof(float2) * (size_t)lens[i], cudaMemcpyDeviceToHost,
                                             k == 0 ? g->s_out : g->x_out[k - 1]));
            }
        for (size_t k = 0; k < g->x_out.size(); ++k) {
            GSDR_CUDA_OK(cudaEventRecord(g->j_out[k], g->x_out[k]));
            GSDR_CUDA_OK(cudaStreamWaitEvent(g->s_out, g->j_out[k], 0));
        }
        GSDR_CUDA_OK(cudaEventRecord(s.out_done, g->s_out));
    } else {
        GSDR_CUDA_OK(cudaEventRecord(s.out_done, g->stream));
    }
    s.used = true;
    g->last_zero_copy = zin && zout;
    g->last_form = (zin ? 1 : 0) | (zout ? 2 : 0);
    if (g->zc_mode == 3) {
        auto& st = g->auto_stamp[g->tickets % 8];
        st.t = t_enter;
        st.kind = sc16 ? 1 : 0;
        st.block = auto_block;
        st.ticket = g->tickets;
    }
    g->tickets++;
    if (valid_lens)
        for (int i = 0; i < S; ++i) valid_lens[i] = lens[i];
    return ticket;
}

int gsdr_rx_group_submit(gsdr_rx_group* g, const gsdr_float2* const* in_host, gsdr_float2* const* out_host, int* valid_lens) {
    return group_submit_any(g, reinterpret_cast<const void* const*>(in_host), false, out_host, valid_lens);
}
int gsdr_rx_group_submit_sc16(gsdr_rx_group* g, const int16_t* const* in_iq, gsdr_float2* const* out_host, int* valid_lens) {
    return group_submit_any(g, reinterpret_cast<const void* const*>(in_iq), true, out_host, valid_lens);
}
int gsdr_rx_group_wait(gsdr_rx_group* g, int ticket) {
    if (!g) return -1;
    if (g->slots.empty()) return 0;
    if (ensure_device(g->device)) return -1;
    GroupSlot& s = g->slots[(uint64_t)ticket % g->slots.size()];
    if (!s.used) return 0;
    GSDR_CUDA_OK(cudaEventSynchronize(s.out_done));
    if (g->zc_mode == 3) {   // a measured period: time from its submit to now
        auto& st = g->auto_stamp[(uint64_t)ticket % 8];
        if (st.block >= 0 && st.kind >= 0 && (st.ticket % 0x40000000u) == (uint64_t)ticket) {
            g->auto_form[st.kind].on_wait(st.block, std::chrono::duration<double>(std::chrono::steady_clock::now() - st.t).count());
            st.block = -1;
        }
    }
    return 0;
}
int gsdr_rx_group_input_consumed(gsdr_rx_group* g, int ticket) {
    if (!g) return -1;
    if (g->slots.empty()) return 1;
    GroupSlot& s = g->slots[(uint64_t)ticket % g->slots.size()];
    if (!s.used) return 1;
    return cudaEventQuery(s.in_done) == cudaSuccess ? 1 : 0;
}
int gsdr_rx_group_process(gsdr_rx_group* g, const gsdr_float2* const* in_host, gsdr_float2* const* out_host, int* valid_lens) {
    const int t = gsdr_rx_group_submit(g, in_host, out_host, valid_lens);
    if (t < 0) return -1;
    return gsdr_rx_group_wait(g, t);
}
int gsdr_rx_group_pipeline_depth(const gsdr_rx_group* g) { return g ? kDepth : 0; }
int gsdr_rx_group_members(const gsdr_rx_group* g) { return g ? (int)g->members.size() : 0; }
int gsdr_rx_group_zero_copy(const gsdr_rx_group* g) { return g && g->last_zero_copy ? 1 : 0; }
int gsdr_rx_group_set_zero_copy(gsdr_rx_group* g, int mode) {
    if (!g || mode < 0 || mode > 3) return -1;
    g->zc_mode = mode;
    if (mode == 3) {   // measure again
        g->auto_form[0] = GroupAutoForm();
        g->auto_form[1] = GroupAutoForm();
        for (auto& st : g->auto_stamp) st.block = -1;
    }
    return 0;
}
int gsdr_rx_group_auto_choice(const gsdr_rx_group* g, int sc16) { return g ? g->auto_form[sc16 ? 1 : 0].choice : -1; }
int gsdr_rx_group_last_form(const gsdr_rx_group* g) { return g ? g->last_form : 0; }

int gsdr_rx_group_sync(gsdr_rx_group* g) {
    if (!g) return -1;
    if (ensure_device(g->device)) return -1;
    GSDR_CUDA_OK(cudaStreamSynchronize(g->s_in));
    GSDR_CUDA_OK(cudaStreamSynchronize(g->stream));
    GSDR_CUDA_OK(cudaStreamSynchronize(g->s_out));
    return 0;
}
int gsdr_rx_group_timer_start(gsdr_rx_group* g) {
    if (!g) return -1;
    if (ensure_device(g->device)) return -1;
    GSDR_CUDA_OK(cudaEventRecord(g->t0, g->stream));
    return 0;
}
int gsdr_rx_group_timer_stop(gsdr_rx_group* g, float* ms) {
    if (!g) return -1;
    if (ensure_device(g->device)) return -1;
    GSDR_CUDA_OK(cudaEventRecord(g->t1, g->stream));
    GSDR_CUDA_OK(cudaEventSynchronize(g->t1));
    float v = 0.f;
    GSDR_CUDA_OK(cudaEventElapsedTime(&v, g->t0, g->t1));
    if (ms) *ms = v;
    return 0;
}
uint64_t gsdr_rx_group_launch_count(const gsdr_rx_group* g) { return g ? g->launches : 0; }

}  // extern "C"
