// PCIe read probe: how fast can a kernel pull pinned HOST memory into the GPU, by access method?
// (a) copy engine (cudaMemcpyAsync)   (b) coalesced 8-byte loads   (c) coalesced 16-byte loads
// (d) cp.async.bulk (TMA 1-D bulk copy, CHUNK bytes per request) into shared memory
// Decides whether the zero-copy blocking call (gsdr_rx_process) should fetch its input rows with bulk copies.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pcie_read_probe tools/pcie_read_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void ld8_kernel(const unsigned long long* __restrict__ p, size_t n, unsigned long long* sink) {
    unsigned long long acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long v;
        asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p + i));
        acc ^= v;
    }
    if (acc == 0x1234567) *sink = acc;
}
__global__ void ld16_kernel(const uint4* __restrict__ p, size_t n, unsigned long long* sink) {
    unsigned int acc = 0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {   // four requests in flight per thread
        uint4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
        acc ^= a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n; i += stride) acc ^= __ldg(p + i).x;
    if (acc == 0x1234567) *sink = acc;
}
template <int CHUNK, int DEPTH>
__global__ void bulk_kernel(const unsigned char* __restrict__ p, size_t bytes, unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar[DEPTH];
    const size_t n_chunks = bytes / CHUNK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) {
            unsigned int a = (unsigned int)__cvta_generic_to_shared(&bar[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        unsigned int acc = 0;
        size_t issued = 0, done = 0;
        const size_t first = blockIdx.x, step = gridDim.x;
        size_t next = first;
        unsigned int phase[DEPTH] = {0};
        auto issue = [&](size_t c, int s) {
            unsigned int a = (unsigned int)__cvta_generic_to_shared(&bar[s]);
            unsigned int d = (unsigned int)__cvta_generic_to_shared(sm + (size_t)s * CHUNK);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                         "l"(p + c * CHUNK), "r"(CHUNK), "r"(a)
                         : "memory");
        };
        for (int s = 0; s < DEPTH && next < n_chunks; ++s, next += step, ++issued) issue(next, s);
        while (done < issued) {
            const int s = (int)(done % DEPTH);
            unsigned int a = (unsigned int)__cvta_generic_to_shared(&bar[s]);
            unsigned int ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                             : "=r"(ok) : "r"(a), "r"(phase[s]) : "memory");
            phase[s] ^= 1u;
            acc ^= *reinterpret_cast<volatile unsigned int*>(sm + (size_t)s * CHUNK);
            ++done;
            if (next < n_chunks) {
                issue(next, s);
                next += step;
                ++issued;
            }
        }
        if (acc == 0x1234567) *sink = acc;
    }
}

int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 64) << 20;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    void* h = nullptr;
    void* d = nullptr;
    unsigned long long* sink = nullptr;
    CK(cudaMallocHost(&h, bytes));
    memset(h, 1, bytes);
    CK(cudaMalloc(&d, bytes));
    CK(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    auto timeit = [&](const char* name, auto fn) {
        fn();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) fn();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"method\": \"%s\", \"GBps\": %.2f}\n", name, (double)bytes * reps / (ms * 1e6));
        fflush(stdout);
    };
    timeit("copy engine, one cudaMemcpyAsync", [&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0); });
    timeit("copy engine, 8 MB per cudaMemcpyAsync", [&] {
        for (size_t o = 0; o < bytes; o += (8u << 20)) cudaMemcpyAsync((char*)d + o, (char*)h + o, 8u << 20, cudaMemcpyHostToDevice, 0);
    });
    for (int ctas : {6, 16, 32, 148}) {
        char nm[96];
        snprintf(nm, sizeof nm, "ld.global.b64 coalesced, %d CTAs x 256 thr", ctas);
        timeit(nm, [&] { ld8_kernel<<<ctas, 256>>>((const unsigned long long*)h, bytes / 8, sink); });
        snprintf(nm, sizeof nm, "ld.global.v4 x4 in flight, %d CTAs x 256 thr", ctas);
        timeit(nm, [&] { ld16_kernel<<<ctas, 256>>>((const uint4*)h, bytes / 16, sink); });
    }
    (void)sms;
    CK(cudaFuncSetAttribute(bulk_kernel<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
    CK(cudaFuncSetAttribute(bulk_kernel<4096, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));
    CK(cudaFuncSetAttribute(bulk_kernel<65536, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 * 3));
    for (int ctas : {4, 8, 16, 32, 148}) {
        char nm[96];
        snprintf(nm, sizeof nm, "cp.async.bulk 16 KB x4 in flight, %d CTAs", ctas);
        timeit(nm, [&] { bulk_kernel<16384, 4><<<ctas, 32, 16384 * 4>>>((const unsigned char*)h, bytes, sink); });
        snprintf(nm, sizeof nm, "cp.async.bulk 4 KB x8 in flight, %d CTAs", ctas);
        timeit(nm, [&] { bulk_kernel<4096, 8><<<ctas, 32, 4096 * 8>>>((const unsigned char*)h, bytes, sink); });
        snprintf(nm, sizeof nm, "cp.async.bulk 64 KB x3 in flight, %d CTAs", ctas);
        timeit(nm, [&] { bulk_kernel<65536, 3><<<ctas, 32, 65536 * 3>>>((const unsigned char*)h, bytes, sink); });
    }
    return 0;
}
