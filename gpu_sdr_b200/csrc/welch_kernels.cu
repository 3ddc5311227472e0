// Client-side Welch spectra on the GPU.
//
// Reference: pyUSRP/USRP_noise.py:655-703 (spec_from_samples) -- what get_noise / calculate_noise run per channel on the
// client, one scipy.signal.welch per IQ component through joblib:
//     rotate the IQ plane so that the mean is real and positive, optionally scale to the carrier and remove it (dBc),
//     clip both ends, then signal.welch(x.real) and signal.welch(x.imag) with nperseg = int(L / welch), the default Hann
//     window (periodic), 50 % overlap, detrend='linear', scaling='density', one-sided; returned as 10 log10.
// nperseg is whatever L / welch gives (arbitrary, not a power of two), so the transform is Bluestein's chirp-z form of the
// DFT over a power-of-two Stockham FFT, with every chirp and twiddle phase formed in integers (k^2 mod 2W, exact for any
// length -- the same idea as the demodulators' integer phase accumulators).  One complex transform per segment serves both
// components: X_re[k] = (Z[k] + conj Z[W-k]) / 2, X_im[k] = (Z[k] - conj Z[W-k]) / 2j.  Segment sums for the linear
// detrend are accumulated in double, deterministically (two-stage, no atomics).
#include <cmath>
#include <vector>

#include "devmath.cuh"

namespace gsdr {
namespace {

constexpr int W_CHUNK = 4096;   // samples per partial sum of the detrend pass

__device__ __forceinline__ float2 phasor_turns64(unsigned long long num, unsigned long long den) {
    // e^{-2 pi j num / den}, num < den < 2^40: phase word = round(num 2^32 / den) via double (53 bits cover the quotient)
    const unsigned int word = (unsigned int)(unsigned long long)((double)num * (4294967296.0 / (double)den) + 0.5);
    float s, c;
    sincos_phase32(word, s, c);
    return make_float2(c, -s);
}

// ---- stage 1: sum z over everything (for the rotation / dBc constants) ----------------------------------------------
__global__ void __launch_bounds__(256) welch_sum_kernel(const float2* __restrict__ z, long long n, double2* __restrict__ partial) {
    __shared__ double sx[256], sy[256];
    double ax = 0.0, ay = 0.0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        const float2 v = z[i];
        ax += v.x, ay += v.y;
    }
    sx[threadIdx.x] = ax, sy[threadIdx.x] = ay;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sx[threadIdx.x] += sx[threadIdx.x + o], sy[threadIdx.x] += sy[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(sx[0], sy[0]);
}

struct WelchXform {   // x = z * c1 + c0 (rotation, dBc scaling and carrier removal folded into one affine map)
    double c1x, c1y, c0;
};
__device__ __forceinline__ double2 welch_x(const float2 v, const WelchXform& t) {
    return make_double2((double)v.x * t.c1x - (double)v.y * t.c1y + t.c0, (double)v.x * t.c1y + (double)v.y * t.c1x);
}

// ---- stage 2: per (segment, chunk): sum x and sum i x, in double --------------------------------------------------------
__global__ void __launch_bounds__(256) welch_detrend_partial_kernel(const float2* __restrict__ z, long long lo, long long step, int W, int n_chunks,
                                                                     WelchXform t, double4* __restrict__ partial) {
    __shared__ double s0x[256], s0y[256], s1x[256], s1y[256];
    const int seg = blockIdx.y, ch = blockIdx.x;
    const float2* x = z + lo + (long long)seg * step;
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
    const int i_end = min(W, (ch + 1) * W_CHUNK);
    for (int i = ch * W_CHUNK + threadIdx.x; i < i_end; i += 256) {
        const double2 v = welch_x(x[i], t);
        a0 += v.x, a1 += v.y, b0 += (double)i * v.x, b1 += (double)i * v.y;
    }
    s0x[threadIdx.x] = a0, s0y[threadIdx.x] = a1, s1x[threadIdx.x] = b0, s1y[threadIdx.x] = b1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s0x[threadIdx.x] += s0x[threadIdx.x + o], s0y[threadIdx.x] += s0y[threadIdx.x + o];
            s1x[threadIdx.x] += s1x[threadIdx.x + o], s1y[threadIdx.x] += s1y[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)seg * n_chunks + ch] = make_double4(s0x[0], s0y[0], s1x[0], s1y[0]);
}
// per segment: intercept a and slope b of the least-squares line through each component (scipy detrend type='linear')
__global__ void welch_detrend_fit_kernel(const double4* __restrict__ partial, int n_chunks, int W, int n_seg, double4* __restrict__ fit) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    double s0x = 0, s0y = 0, s1x = 0, s1y = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const double4 p = partial[(size_t)seg * n_chunks + c];
        s0x += p.x, s0y += p.y, s1x += p.z, s1y += p.w;
    }
    const double n = (double)W, ibar = (n - 1.0) * 0.5;
    const double sxx = n * (n * n - 1.0) / 12.0;   // sum (i - ibar)^2
    const double bx = sxx > 0 ? (s1x - ibar * s0x) / sxx : 0.0, by = sxx > 0 ? (s1y - ibar * s0y) / sxx : 0.0;
    fit[seg] = make_double4(s0x / n - bx * ibar, s0y / n - by * ibar, bx, by);   // (a_re, a_im, b_re, b_im)
}

// ---- stage 3: detrend, window, pre-chirp, zero-pad: the Bluestein input ------------------------------------------------
// direct == 1 (W a power of two): no chirp, the FFT itself is the transform.
__global__ void __launch_bounds__(256) welch_prepare_kernel(const float2* __restrict__ z, long long lo, long long step, int W, int M2, int direct,
                                                            WelchXform t, const double4* __restrict__ fit, float2* __restrict__ out) {
    const int seg = blockIdx.y;
    const float2* x = z + lo + (long long)seg * step;
    const double4 f = fit[seg];
    float2* o = out + (size_t)seg * M2;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < M2; i += 256 * gridDim.x) {
        float2 r = make_float2(0.f, 0.f);
        if (i < W) {
            const double2 v = welch_x(x[i], t);
            const float2 hw = phasor_turns64((unsigned long long)i, (unsigned long long)W);        // cos(2 pi i / W) in .x
            const double win = 0.5 - 0.5 * (double)hw.x;                                             // periodic Hann
            const float2 d = make_float2((float)((v.x - f.x - f.z * (double)i) * win), (float)((v.y - f.y - f.w * (double)i) * win));
            if (direct) {
                r = d;
            } else {
                const unsigned long long q = ((unsigned long long)i * (unsigned long long)i) % (2ull * (unsigned long long)W);
                r = dev_cmul(d, phasor_turns64(q, 2ull * (unsigned long long)W));                    // e^{-j pi i^2 / W}
            }
        }
        o[i] = r;
    }
}
// the chirp filter h[i] = e^{+j pi i^2 / W} for |i| < W, wrapped into M2 points
__global__ void __launch_bounds__(256) welch_chirp_filter_kernel(int W, int M2, float2* __restrict__ h) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < M2; i += 256 * gridDim.x) {
        const int d = i < W ? i : (M2 - i < W ? M2 - i : -1);
        float2 r = make_float2(0.f, 0.f);
        if (d >= 0) {
            const unsigned long long q = ((unsigned long long)d * (unsigned long long)d) % (2ull * (unsigned long long)W);
            const float2 p = phasor_turns64(q, 2ull * (unsigned long long)W);
            r = make_float2(p.x, -p.y);
        }
        h[i] = r;
    }
}

// ---- power-of-two FFT: Stockham autosort, radix 2, one pass per launch, batched ------------------------------------------
// pass with stride s (1, 2, 4, ...): y[q + s (2 p + r)] from x[q + s (p + r m)], m = n / 2: out-of-place, no bit reversal.
__global__ void __launch_bounds__(256) welch_fft_pass_kernel(const float2* __restrict__ x, float2* __restrict__ y, int n, int s, int inverse) {
    const int m = n >> 1;
    const float2* xb = x + (size_t)blockIdx.y * n;
    float2* yb = y + (size_t)blockIdx.y * n;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < m; i += 256 * gridDim.x) {
        const int p = i / s, q = i - p * s;                 // p < m / s
        const float2 a = xb[q + s * p], b = xb[q + s * (p + m / s)];
        // twiddle e^{-+2 pi j p s / n}
        float2 w = phasor_turns64((unsigned long long)p * (unsigned long long)s, (unsigned long long)n);
        if (inverse) w.y = -w.y;
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        yb[q + s * (2 * p)] = make_float2(a.x + b.x, a.y + b.y);
        yb[q + s * (2 * p + 1)] = dev_cmul(d, w);
    }
}
__global__ void __launch_bounds__(256) welch_pointwise_mul_kernel(float2* __restrict__ a, const float2* __restrict__ h, int M2) {
    float2* ab = a + (size_t)blockIdx.y * M2;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < M2; i += 256 * gridDim.x) ab[i] = dev_cmul(ab[i], h[i]);
}

// ---- stage 5: spectra of the two components from the complex transform, summed over the segments ---------------------
// Bluestein: Z[k] = e^{-j pi k^2 / W} * conv[k] / M2 (unnormalised inverse FFT); direct: Z[k] = fft[k].
__global__ void __launch_bounds__(256) welch_psd_kernel(const float2* __restrict__ conv, int W, int M2, int direct, int n_seg, double scale,
                                                        float* __restrict__ re_db, float* __restrict__ im_db) {
    const int n_freq = W / 2 + 1;
    const float inv = direct ? 1.0f : 1.0f / (float)M2;
    for (int k = blockIdx.x * 256 + threadIdx.x; k < n_freq; k += 256 * gridDim.x) {
        const int k2 = (W - k) % W;
        float2 c1 = make_float2(1.f, 0.f), c2 = c1;
        if (!direct) {
            c1 = phasor_turns64(((unsigned long long)k * (unsigned long long)k) % (2ull * W), 2ull * W);
            c2 = phasor_turns64(((unsigned long long)k2 * (unsigned long long)k2) % (2ull * W), 2ull * W);
        }
        double pr = 0.0, pi = 0.0;
        for (int sgm = 0; sgm < n_seg; ++sgm) {
            const float2* cb = conv + (size_t)sgm * M2;
            float2 a = dev_cmul(cb[k], c1), b = dev_cmul(cb[k2], c2);
            a.x *= inv, a.y *= inv, b.x *= inv, b.y *= inv;
            const float xr = 0.5f * (a.x + b.x), xi = 0.5f * (a.y - b.y);    // (Z[k] + conj Z[W-k]) / 2
            const float yr = 0.5f * (a.y + b.y), yi = -0.5f * (a.x - b.x);   // (Z[k] - conj Z[W-k]) / 2j
            pr += (double)xr * xr + (double)xi * xi;
            pi += (double)yr * yr + (double)yi * yi;
        }
        const bool edge = k == 0 || (2 * k == W);   // one-sided density: everything but DC and Nyquist counts twice
        const double f = scale * (edge ? 1.0 : 2.0) / (double)n_seg;
        re_db[k] = (float)(10.0 * log10(pr * f));
        im_db[k] = (float)(10.0 * log10(pi * f));
    }
}

int grid_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace

// nperseg and the number of one-sided frequencies for a record of n samples (the reference's int(L / welch); scipy clamps
// nperseg to the clipped length)
long long welch_nperseg(long long n, int welch, long long clip) {
    long long W = welch > 0 ? n / welch : n;
    const long long n_c = n - 2 * clip;
    if (W > n_c) W = n_c;
    return W;
}

int welch_spectra(const float2* d_z, long long n, double fs, int welch, int dbc, int rotate, long long clip, float* d_re_db, float* d_im_db,
                  cudaStream_t st) {
    if (n <= 0 || clip < 0 || 2 * clip >= n || fs <= 0.0) {
        set_error("welch_spectra: bad argument (n=%lld clip=%lld)", n, clip);
        return -1;
    }
    const long long lo = clip, n_c = n - 2 * clip;
    const long long W = welch_nperseg(n, welch, clip);
    if (W < 2 || W > (1ll << 26)) {
        set_error("welch_spectra: nperseg %lld out of range", W);
        return -1;
    }
    const long long noverlap = W / 2, step = W - noverlap;
    const int n_seg = (int)((n_c - noverlap) / step);
    const bool direct = (W & (W - 1)) == 0;
    long long M2 = W;
    if (!direct) {
        M2 = 1;
        while (M2 < 2 * W - 1) M2 <<= 1;
    }
    // ---- rotation / dBc constants from the mean of the whole record (the reference takes the mean before clipping)
    WelchXform t{1.0, 0.0, 0.0};
    double2* d_part = nullptr;
    const int nb = 256;
    if (rotate || dbc) {
        GSDR_CUDA_OK(cudaMalloc(&d_part, sizeof(double2) * nb));
        welch_sum_kernel<<<nb, 256, 0, st>>>(d_z, n, d_part);
        std::vector<double2> hp(nb);
        GSDR_CUDA_OK(cudaMemcpyAsync(hp.data(), d_part, sizeof(double2) * nb, cudaMemcpyDeviceToHost, st));
        GSDR_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(d_part);
        double mx = 0, my = 0;
        for (auto& p : hp) mx += p.x, my += p.y;
        mx /= (double)n, my /= (double)n;
        const double m2 = mx * mx + my * my, mabs = std::sqrt(m2);
        if (m2 > 0) {
            if (dbc) {   // z / mean - 1 (with or without the rotation: the rotation cancels in z / mean)
                t.c1x = mx / m2, t.c1y = -my / m2, t.c0 = -1.0;
            } else {     // z * |m| / m
                t.c1x = mx / mabs, t.c1y = -my / mabs;
            }
        }
    }
    const int n_chunks = (int)((W + W_CHUNK - 1) / W_CHUNK);
    double4 *d_partial = nullptr, *d_fit = nullptr;
    float2 *d_a = nullptr, *d_b = nullptr, *d_h = nullptr;
    int rc = -1;
    do {
        if (cudaMalloc(&d_partial, sizeof(double4) * (size_t)n_seg * n_chunks) != cudaSuccess || cudaMalloc(&d_fit, sizeof(double4) * n_seg) != cudaSuccess ||
            cudaMalloc(&d_a, sizeof(float2) * (size_t)n_seg * M2) != cudaSuccess || cudaMalloc(&d_b, sizeof(float2) * (size_t)n_seg * M2) != cudaSuccess ||
            (!direct && cudaMalloc(&d_h, sizeof(float2) * 2 * (size_t)M2) != cudaSuccess)) {
            set_error("welch_spectra: cudaMalloc failed (%d segments of %lld points)", n_seg, M2);
            break;
        }
        welch_detrend_partial_kernel<<<dim3(n_chunks, n_seg), 256, 0, st>>>(d_z, lo, step, (int)W, n_chunks, t, d_partial);
        welch_detrend_fit_kernel<<<(n_seg + 63) / 64, 64, 0, st>>>(d_partial, n_chunks, (int)W, n_seg, d_fit);
        welch_prepare_kernel<<<dim3(grid_for(M2), n_seg), 256, 0, st>>>(d_z, lo, step, (int)W, (int)M2, direct ? 1 : 0, t, d_fit, d_a);
        auto fft = [&](float2*& src, float2*& dst, int batch, int inverse) {
            for (long long s = 1; s < M2; s <<= 1) {
                welch_fft_pass_kernel<<<dim3(grid_for(M2 / 2), batch), 256, 0, st>>>(src, dst, (int)M2, (int)s, inverse);
                std::swap(src, dst);
            }
        };
        float2 *src = d_a, *dst = d_b;
        fft(src, dst, n_seg, 0);
        if (!direct) {
            float2 *hs = d_h, *hd = d_h + M2;
            welch_chirp_filter_kernel<<<grid_for(M2), 256, 0, st>>>((int)W, (int)M2, hs);
            fft(hs, hd, 1, 0);
            welch_pointwise_mul_kernel<<<dim3(grid_for(M2), n_seg), 256, 0, st>>>(src, hs, (int)M2);
            fft(src, dst, n_seg, 1);
        }
        // density scaling: 1 / (fs * sum w^2); periodic Hann: sum w^2 = 3 W / 8 (W >= 3)
        double sw2 = 0.375 * (double)W;
        if (W < 3) {
            sw2 = 0.0;
            for (int i = 0; i < W; ++i) {
                const double wv = 0.5 - 0.5 * std::cos(6.283185307179586 * i / (double)W);
                sw2 += wv * wv;
            }
        }
        welch_psd_kernel<<<grid_for(W / 2 + 1), 256, 0, st>>>(src, (int)W, (int)M2, direct ? 1 : 0, n_seg, 1.0 / (fs * sw2), d_re_db, d_im_db);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("welch_spectra: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = (int)(W / 2 + 1);
    } while (false);
    cudaFree(d_partial);
    cudaFree(d_fit);
    cudaFree(d_a);
    cudaFree(d_b);
    if (d_h) cudaFree(d_h);
    return rc;
}

}  // namespace gsdr

using namespace gsdr;

extern "C" {

long long gsdr_spec_n_freq(size_t n, int welch, size_t clip_samples) {
    if (n == 0 || 2 * clip_samples >= n) return -1;
    return welch_nperseg((long long)n, welch, (long long)clip_samples) / 2 + 1;
}

int gsdr_spec_from_samples(int device, const gsdr_float2* samples, size_t n, double sampling_rate, int welch, int dbc, int rotate,
                           size_t clip_samples, double* freqs, float* re_db, float* im_db) {
    if (!samples || !re_db || !im_db) {
        set_error("gsdr_spec_from_samples: null argument");
        return -1;
    }
    const long long nf = gsdr_spec_n_freq(n, welch, clip_samples);
    if (nf <= 0) {
        set_error("gsdr_spec_from_samples: empty record");
        return -1;
    }
    if (ensure_device(device)) return -1;
    float2* d_z = nullptr;
    float *d_re = nullptr, *d_im = nullptr;
    cudaStream_t st = nullptr;
    int rc = -1;
    do {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&d_z, sizeof(float2) * n) != cudaSuccess ||
            cudaMalloc(&d_re, sizeof(float) * nf) != cudaSuccess || cudaMalloc(&d_im, sizeof(float) * nf) != cudaSuccess ||
            cudaMemcpyAsync(d_z, samples, sizeof(float2) * n, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            set_error("gsdr_spec_from_samples: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = welch_spectra(d_z, (long long)n, sampling_rate, welch, dbc, rotate, (long long)clip_samples, d_re, d_im, st);
        if (rc < 0) break;
        if (cudaMemcpyAsync(re_db, d_re, sizeof(float) * nf, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(im_db, d_im, sizeof(float) * nf, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("gsdr_spec_from_samples: %s", cudaGetErrorString(cudaGetLastError()));
            rc = -1;
            break;
        }
        if (freqs) {   // numpy.fft.rfftfreq(nperseg, 1 / fs)
            const long long W = welch_nperseg((long long)n, welch, (long long)clip_samples);
            for (long long k = 0; k < nf; ++k) freqs[k] = (double)k * sampling_rate / (double)W;
        }
    } while (false);
    if (d_z) cudaFree(d_z);
    if (d_re) cudaFree(d_re);
    if (d_im) cudaFree(d_im);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
