// gsdr_tx: host side of the TX buffer generator (the TX_buffer_generator replacement).
//
// Reference: cpp/USRP_buffer_generator.cpp.  TONES (:60-99) builds one period of the multi-tone
// waveform, replicates it if buffer_len > rate, appends buffer_len wrap-around samples and then
// serves get() by pointer arithmetic (:226-229); CHIRP (:101-157, :208-221) synthesises every
// buffer on the GPU and copies it to the caller's buffer.  Same contract here; the period buffer
// is synthesised by the integer-phase GEMM kernel in tones_kernels.cu instead of a rate-point
// inverse cuFFT, and lives in pinned host memory (plus a device copy for device-resident use).
#include <cmath>
#include <map>
#include <memory>

#include "common.hpp"

using namespace gsdr;

struct gsdr_tx {
    int device = 0;
    int mode = GSDR_TONES;
    int rate = 0;
    long long L = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint64_t launches = 0;
    // TONES
    float2* h_base = nullptr;
    float2* d_base = nullptr;
    size_t period_len = 0, last_sample = 0;
    // CHIRP
    gsdr_chirp_param cpar{};
    ChirpDev cdev{};
    float scale = 1.f;
    unsigned long long last_index = 0;
    float2* d_buf = nullptr;
};

namespace {

void tx_free(gsdr_tx* tx) {
    cudaSetDevice(tx->device);
    if (tx->h_base) {
        host_registry_remove(tx->h_base);
        cudaFreeHost(tx->h_base);
    }
    if (tx->d_base) cudaFree(tx->d_base);
    if (tx->d_buf) cudaFree(tx->d_buf);
    if (tx->t0) cudaEventDestroy(tx->t0);
    if (tx->t1) cudaEventDestroy(tx->t1);
    if (tx->stream) cudaStreamDestroy(tx->stream);
}

int init_tones(gsdr_tx* tx, const gsdr_param* p) {
    const long long R = tx->rate;
    if (R <= 0) {
        set_error("TX TONES: rate must be positive");
        return -1;
    }
    const size_t T = (size_t)p->n_wave_type;
    if (p->n_freq < T || p->n_ampl < T) {
        set_error("TX TONES: %zu tones but %llu freq / %llu ampl", T, (unsigned long long)p->n_freq, (unsigned long long)p->n_ampl);
        return -1;
    }
    // spectrum slot of each tone: f>0 ? f : R+f (cpp/kernels.cu:626-631).  A later tone on the same
    // slot overwrites an earlier one; slot R (f == 0) and anything outside [0,R) is outside the
    // reference's buffer and contributes nothing.
    std::map<int, float> slots;
    for (size_t i = 0; i < T; ++i) {
        const long long f = p->freq[i];
        const long long k = f > 0 ? f : R + f;
        if (k >= 0 && k < R) slots[(int)k] = p->ampl[i];
        else if (f == 0)   // the reference writes spectrum slot R, one past its buffer (cpp/kernels.cu:626-631): a DC tone is silent
            fprintf(stderr, "gsdr_tx: warning: TX tone %zu has freq == 0 Hz; like the reference it contributes nothing to the buffer\n", i);
    }
    std::vector<int> bins;
    std::vector<float> ampl;
    for (auto& kv : slots) {
        bins.push_back(kv.first);
        ampl.push_back(kv.second);
    }
    int* d_bins = nullptr;
    float* d_ampl = nullptr;
    const size_t nt = bins.size();
    GSDR_CUDA_OK(cudaMalloc(&d_bins, sizeof(int) * (nt ? nt : 1)));
    GSDR_CUDA_OK(cudaMalloc(&d_ampl, sizeof(float) * (nt ? nt : 1)));
    if (nt) {
        GSDR_CUDA_OK(cudaMemcpy(d_bins, bins.data(), sizeof(int) * nt, cudaMemcpyHostToDevice));
        GSDR_CUDA_OK(cudaMemcpy(d_ampl, ampl.data(), sizeof(float) * nt, cudaMemcpyHostToDevice));
    }
    tx->period_len = (size_t)R;
    size_t ratio = 1;
    if (tx->L > R) {  // cpp/USRP_buffer_generator.cpp:80-92
        ratio = (size_t)std::ceil((float)tx->L / (float)R);
        tx->period_len = ratio * (size_t)R;
    }
    const size_t total = tx->period_len + (size_t)tx->L;
    GSDR_CUDA_OK(cudaMalloc(&tx->d_base, sizeof(float2) * total));
    const int nl = tones_synth_launch(tx->d_base, 0, R, d_bins, d_ampl, (int)nt, (int)R, tx->stream);
    if (nl < 0) return -1;
    tx->launches += nl;
    for (size_t j = 1; j < ratio; ++j)
        GSDR_CUDA_OK(cudaMemcpyAsync(tx->d_base + j * R, tx->d_base, sizeof(float2) * R, cudaMemcpyDeviceToDevice, tx->stream));
    // wrap-around tail: the first buffer_len samples again (:95-97)
    GSDR_CUDA_OK(cudaMemcpyAsync(tx->d_base + tx->period_len, tx->d_base, sizeof(float2) * tx->L, cudaMemcpyDeviceToDevice,
                                 tx->stream));
    GSDR_CUDA_OK(cudaMallocHost(&tx->h_base, sizeof(float2) * total));
    host_registry_add(tx->h_base, sizeof(float2) * total);   // get() hands out pointers into it: RX may read them in place
    GSDR_CUDA_OK(cudaMemcpyAsync(tx->h_base, tx->d_base, sizeof(float2) * total, cudaMemcpyDeviceToHost, tx->stream));
    GSDR_CUDA_OK(cudaStreamSynchronize(tx->stream));
    cudaFree(d_bins);
    cudaFree(d_ampl);
    tx->last_sample = 0;
    return 0;
}

int init_chirp(gsdr_tx* tx, const gsdr_param* p) {
    if (p->n_chirp_t < 1 || p->n_chirp_f < 1 || p->n_swipe_s < 1 || p->n_freq < 1 || p->n_ampl < 1) {
        set_error("TX CHIRP: chirp_t/chirp_f/swipe_s/freq/ampl must each hold one value");
        return -1;
    }
    chirp_params(tx->rate, p->freq[0], p->chirp_f[0], p->swipe_s[0], p->chirp_t[0], true, &tx->cpar);
    if (tx->cpar.num_steps * tx->cpar.length == 0) {
        set_error("TX CHIRP: empty sweep (num_steps*length == 0)");
        return -1;
    }
    tx->cdev.period = tx->cpar.num_steps * tx->cpar.length;
    tx->cdev.length = tx->cpar.length;
    tx->cdev.chirpness = tx->cpar.chirpness;
    tx->cdev.f0 = tx->cpar.f0;
    tx->cdev.num_steps = tx->cpar.num_steps;
    tx->scale = p->ampl[0];
    tx->last_index = 0;
    GSDR_CUDA_OK(cudaMalloc(&tx->d_buf, sizeof(float2) * tx->L));
    return 0;
}

}  // namespace

extern "C" {

gsdr_tx* gsdr_tx_create(const gsdr_param* p, int device) {
    if (!p || p->n_wave_type == 0) {
        set_error("gsdr_tx_create: no wave type given");
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("gsdr_tx_create: no CUDA device available (this library has no CPU path)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        set_error("gsdr_tx_create: device %d out of range", device);
        return nullptr;
    }
    const int mode = p->wave_type[0];
    int chirps = 0;
    for (uint64_t i = 0; i < p->n_wave_type; ++i) {
        if (p->wave_type[i] != mode) {
            set_error("Mixed TX buffer generation has been requested. This feature is not implemented yet.");
            return nullptr;
        }
        if (p->wave_type[i] == GSDR_CHIRP) ++chirps;
    }
    if (chirps > 1) {
        set_error("Multiple chirp TX buffer generation has been requested. This feature is not implemented yet.");
        return nullptr;
    }
    if (mode != GSDR_TONES && mode != GSDR_NOISE && mode != GSDR_CHIRP) {
        set_error("TX wave type %d is not implemented (reference: NODSP/SWONLY/RAMP/DIRECT exit)", mode);
        return nullptr;
    }
    if (p->buffer_len == 0 || p->buffer_len > 0x7fffffffULL) {
        set_error("gsdr_tx_create: buffer_len out of range");
        return nullptr;
    }
    std::unique_ptr<gsdr_tx> tx(new gsdr_tx());
    tx->device = device;
    tx->mode = (mode == GSDR_CHIRP) ? GSDR_CHIRP : GSDR_TONES;  // NOISE falls through to TONES (:56-60)
    tx->rate = p->rate;
    tx->L = (long long)p->buffer_len;
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&tx->stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreate(&tx->t0) != cudaSuccess || cudaEventCreate(&tx->t1) != cudaSuccess) {
        set_error("gsdr_tx_create: stream/event creation failed");
        tx_free(tx.get());
        return nullptr;
    }
    const int rc = tx->mode == GSDR_CHIRP ? init_chirp(tx.get(), p) : init_tones(tx.get(), p);
    if (rc) {
        tx_free(tx.get());
        return nullptr;
    }
    return tx.release();
}

void gsdr_tx_destroy(gsdr_tx* tx) {
    if (!tx) return;
    cudaSetDevice(tx->device);
    if (tx->stream) cudaStreamSynchronize(tx->stream);
    tx_free(tx);
    delete tx;
}

int gsdr_tx_get(gsdr_tx* tx, gsdr_float2** io) {
    if (!tx || !io) {
        set_error("gsdr_tx_get: null argument");
        return -1;
    }
    if (tx->mode == GSDR_TONES) {  // cpp/USRP_buffer_generator.cpp:226-229
        *io = reinterpret_cast<gsdr_float2*>(tx->h_base + tx->last_sample);
        tx->last_sample = (tx->last_sample + (size_t)tx->L) % tx->period_len;
        return 0;
    }
    if (!*io) {
        set_error("gsdr_tx_get: CHIRP needs a caller buffer");
        return -1;
    }
    GSDR_CUDA_OK(cudaSetDevice(tx->device));
    const int nl = chirp_gen_launch(tx->d_buf, tx->L, tx->last_index, tx->cdev, tx->scale, tx->stream);
    if (nl < 0) return -1;
    tx->launches += nl;
    tx->last_index = (tx->last_index + (unsigned long long)tx->L) % tx->cdev.period;
    GSDR_CUDA_OK(cudaMemcpyAsync(*io, tx->d_buf, sizeof(float2) * tx->L, cudaMemcpyDeviceToHost, tx->stream));
    GSDR_CUDA_OK(cudaStreamSynchronize(tx->stream));
    return 0;
}

int gsdr_tx_get_device(gsdr_tx* tx, gsdr_float2* out_dev, int n_buffers) {
    if (!tx || !out_dev || n_buffers <= 0) {
        set_error("gsdr_tx_get_device: bad argument");
        return -1;
    }
    GSDR_CUDA_OK(cudaSetDevice(tx->device));
    float2* out = reinterpret_cast<float2*>(out_dev);
    if (tx->mode == GSDR_TONES) {
        for (int b = 0; b < n_buffers; ++b) {
            GSDR_CUDA_OK(cudaMemcpyAsync(out + (size_t)b * tx->L, tx->d_base + tx->last_sample, sizeof(float2) * tx->L,
                                         cudaMemcpyDeviceToDevice, tx->stream));
            tx->last_sample = (tx->last_sample + (size_t)tx->L) % tx->period_len;
        }
        return 0;
    }
    const int nl = chirp_gen_launch(out, tx->L * n_buffers, tx->last_index, tx->cdev, tx->scale, tx->stream);
    if (nl < 0) return -1;
    tx->launches += nl;
    tx->last_index = (tx->last_index + (unsigned long long)(tx->L * n_buffers)) % tx->cdev.period;
    return 0;
}

int gsdr_tx_sync(gsdr_tx* tx) {
    if (!tx) return -1;
    GSDR_CUDA_OK(cudaSetDevice(tx->device));
    GSDR_CUDA_OK(cudaStreamSynchronize(tx->stream));
    return 0;
}
int gsdr_tx_dynamic_buffer(const gsdr_tx* tx) { return tx && tx->mode != GSDR_TONES ? 1 : 0; }
int gsdr_tx_buffer_len(const gsdr_tx* tx) { return tx ? (int)tx->L : 0; }
uint64_t gsdr_tx_launch_count(const gsdr_tx* tx) { return tx ? tx->launches : 0; }
int gsdr_tx_timer_start(gsdr_tx* tx) {
    if (!tx) return -1;
    GSDR_CUDA_OK(cudaSetDevice(tx->device));
    GSDR_CUDA_OK(cudaEventRecord(tx->t0, tx->stream));
    return 0;
}
int gsdr_tx_timer_stop(gsdr_tx* tx, float* ms) {
    if (!tx) return -1;
    GSDR_CUDA_OK(cudaSetDevice(tx->device));
    GSDR_CUDA_OK(cudaEventRecord(tx->t1, tx->stream));
    GSDR_CUDA_OK(cudaEventSynchronize(tx->t1));
    float v = 0.f;
    GSDR_CUDA_OK(cudaEventElapsedTime(&v, tx->t0, tx->t1));
    if (ms) *ms = v;
    return 0;
}
int gsdr_tx_chirp_param(const gsdr_tx* tx, gsdr_chirp_param* out) {
    if (!tx || !out || tx->mode != GSDR_CHIRP) return -1;
    *out = tx->cpar;
    return 0;
}

}  // extern "C"
