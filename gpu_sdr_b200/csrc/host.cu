// Host-side runtime pieces of the path: the pinned buffer pool (preallocator<float2> contract),
// the hardware-free replay source, and the small C-ABI utilities (errors, memory, probes).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <map>
#include <random>
#include <shared_mutex>
#include <thread>
#include <unordered_map>

#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>

#include "common.hpp"

using namespace gsdr;

// ------------------------------------------------------------------------------------------------
// NUMA-local pinned memory.  On an 8-GPU box half of the GPUs hang off the other CPU socket: a pinned
// buffer that lives on the wrong node is DMA-ed across the socket interconnect, which all of them then
// share.  The pool and gsdr_host_alloc therefore ask the kernel (set_mempolicy, MPOL_PREFERRED) for
// pages on the node of the GPU that is current when the buffer is allocated; the reference pins its
// worker threads to cores instead (cpp/USRP_server_link_threads.cpp:264,299) and allocates wherever
// they run.  Everything degrades to a plain cudaMallocHost when sysfs or the syscall say no.
// ------------------------------------------------------------------------------------------------
namespace {
int numa_node_of_device_uncached(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* c = bus; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
int numa_node_of_device(int device) {
    static std::mutex m;
    static int cache[64];
    static bool known[64] = {false};
    if (device < 0 || device >= 64) return -1;
    std::lock_guard<std::mutex> lk(m);
    if (!known[device]) cache[device] = numa_node_of_device_uncached(device), known[device] = true;
    return cache[device];
}
// cudaMallocHost with the calling thread's memory policy pointed at the current GPU's node for the duration.  The caller's
// own policy (numactl --membind / --interleave, or a set_mempolicy of the application -- the caller may be the reference
// server's link thread) is read first and put back exactly as it was.
cudaError_t pinned_alloc_local(void** p, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1, cudaGetLastError();
    static const bool numa_off = [] {
        const char* e = getenv("GSDR_NUMA_LOCAL");
        return e && e[0] == '0';
    }();
    const int node = numa_off ? -1 : numa_node_of_device(dev);
    bool bound = false;
    int old_mode = 0;
    unsigned long old_mask[16] = {0};   // 1024 nodes
#if defined(SYS_set_mempolicy) && defined(SYS_get_mempolicy)
    if (node >= 0 && node < 64 &&
        syscall(SYS_get_mempolicy, &old_mode, old_mask, sizeof(old_mask) * 8, nullptr, 0) == 0) {
        unsigned long mask = 1ul << node;
        bound = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, sizeof(mask) * 8 + 1) == 0;
    }
#endif
    const cudaError_t rc = cudaMallocHost(p, bytes);
#if defined(SYS_set_mempolicy) && defined(SYS_get_mempolicy)
    if (bound) {
        bool any = false;
        for (unsigned long m : old_mask) any = any || m != 0;
        if (syscall(SYS_set_mempolicy, old_mode, any ? old_mask : nullptr, any ? sizeof(old_mask) * 8 : 0) != 0)
            syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
    }
#endif
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Which host pointers can a kernel read / write in place?  Every pinned block this library hands out (pool buffers,
// gsdr_host_alloc, the TX period buffer) is entered here with its device-side alias, so the zero-copy entry points
// answer "is this buffer pinned and mapped" with one ordered-map lookup and no CUDA call -- that is the reference's flow
// (buffers come from the preallocator) and the bench's.  Pointers the library did not allocate are asked of the driver
// (cudaPointerGetAttributes, ~1 us) on every call: remembering the answer would be wrong the day such a buffer is freed and
// its address reused by pageable memory.  GSDR_PROCESS_PTRCACHE=1 remembers it anyway, for callers that guarantee a buffer
// keeps its nature while the library is in use.
// ------------------------------------------------------------------------------------------------
struct HostRange {
    size_t bytes;
    intptr_t delta;   // device alias = host address + delta (0 under unified addressing)
};
std::shared_mutex g_reg_mutex;
std::map<uintptr_t, HostRange> g_ranges;                 // keyed by base address
std::unordered_map<uintptr_t, void*> g_foreign;          // pointer -> alias (nullptr: pageable / not mapped)
}  // namespace

namespace gsdr {
void host_registry_add(const void* p, size_t bytes) {
    if (!p || !bytes) return;
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    std::unique_lock<std::shared_mutex> lk(g_reg_mutex);
    g_ranges[reinterpret_cast<uintptr_t>(p)] = HostRange{bytes, (intptr_t)(reinterpret_cast<uintptr_t>(d) - reinterpret_cast<uintptr_t>(p))};
}
void host_registry_remove(const void* p) {
    if (!p) return;
    std::unique_lock<std::shared_mutex> lk(g_reg_mutex);
    g_ranges.erase(reinterpret_cast<uintptr_t>(p));
    g_foreign.clear();   // addresses may be reused by memory of another kind
}
void* host_alias_of(const void* host, size_t bytes) {
    static const bool cache_on = [] {
        const char* e = getenv("GSDR_PROCESS_PTRCACHE");
        return e && e[0] == '1';
    }();
    const uintptr_t a = reinterpret_cast<uintptr_t>(host);
    {
        std::shared_lock<std::shared_mutex> lk(g_reg_mutex);
        auto it = g_ranges.upper_bound(a);
        if (it != g_ranges.begin()) {
            --it;
            if (a >= it->first && a + bytes <= it->first + it->second.bytes) return reinterpret_cast<void*>(a + it->second.delta);
        }
        if (cache_on) {
            auto f = g_foreign.find(a);
            if (f != g_foreign.end()) return f->second;
        }
    }
    cudaPointerAttributes at{};
    void* alias = nullptr;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) cudaGetLastError();
    else if (at.type == cudaMemoryTypeHost) alias = at.devicePointer;
    if (cache_on) {
        std::unique_lock<std::shared_mutex> lk(g_reg_mutex);
        if (g_foreign.size() < 65536) g_foreign[a] = alias;
    }
    return alias;
}
int ensure_device(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return 0;
    GSDR_CUDA_OK(cudaSetDevice(device));
    return 0;
}
}  // namespace gsdr

namespace {
}  // namespace

// ------------------------------------------------------------------------------------------------
// Pinned pool.  Reference: template preallocator<T>, headers/USRP_server_memory_management.hpp:
// 103-273 -- cudaMallocHost buffers behind two lock-free queues, a filler thread that pre-fills
// pipe_size-1 buffers and grows the pool when fewer than 10 % are free, a recycler thread, and
// get()/trash()/close().  Same contract; one mutex-protected free list, one grower thread (so the
// real-time get() never pays a cudaMallocHost unless the pool is completely dry), and trash()
// recycles immediately instead of after the reference's 3 ms sleep per buffer.
// ------------------------------------------------------------------------------------------------
struct gsdr_pool {
    size_t vector_size = 0;
    int pipe_size = 0;
    int device = 0;  // GPU that was current at creation: pinned pages go to its NUMA node
    bool grow = true;
    std::mutex m;
    std::condition_variable cv_free, cv_grow;
    std::deque<gsdr_float2*> free_list;
    std::vector<gsdr_float2*> all;
    std::thread grower;
    bool closing = false;
    int waiters = 0;   // threads inside gsdr_pool_get: close() does not free the pool under them
    std::condition_variable cv_idle;

    // Every buffer has GSDR_POOL_HEADROOM bytes in front of it, so the data-socket header can be written directly
    // before the payload and the whole frame leaves with one send() (gsdr_packet_frame).
    gsdr_float2* alloc_one() {
        void* p = nullptr;
        const size_t bytes = vector_size * sizeof(gsdr_float2) + GSDR_POOL_HEADROOM;
        if (pinned_alloc_local(&p, bytes) != cudaSuccess) return nullptr;
        host_registry_add(p, bytes);
        return reinterpret_cast<gsdr_float2*>(static_cast<char*>(p) + GSDR_POOL_HEADROOM);
    }
    static void free_one(gsdr_float2* b) {
        void* p = reinterpret_cast<char*>(b) - GSDR_POOL_HEADROOM;
        host_registry_remove(p);
        cudaFreeHost(p);
    }
    void grow_loop() {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(m);
        while (!closing) {
            cv_grow.wait(lk, [&] { return closing || (grow && (double)free_list.size() < pipe_size / 10.); });
            if (closing) break;
            lk.unlock();
            gsdr_float2* b = alloc_one();
            lk.lock();
            if (!b) break;
            all.push_back(b);
            free_list.push_back(b);
            ++pipe_size;
            cv_free.notify_one();
        }
    }
};

extern "C" {

gsdr_pool* gsdr_pool_create(size_t vector_size, int pipe_size, int prefill) {
    if (vector_size == 0 || pipe_size < 2) {
        set_error("gsdr_pool_create: vector_size=%zu pipe_size=%d", vector_size, pipe_size);
        return nullptr;
    }
    std::unique_ptr<gsdr_pool> pool(new gsdr_pool());
    pool->vector_size = vector_size;
    pool->pipe_size = pipe_size;
    pool->grow = prefill != 0;
    if (cudaGetDevice(&pool->device) != cudaSuccess) pool->device = 0, cudaGetLastError();
    for (int i = 0; i < pipe_size - 1; ++i) {  // reference pre-fills pipe_size-1 buffers (:211-227)
        gsdr_float2* b = pool->alloc_one();
        if (!b) {
            set_error("Memory manager cannot allocate pinned host memory!");
            for (auto* q : pool->all) gsdr_pool::free_one(q);
            return nullptr;
        }
        pool->all.push_back(b);
        pool->free_list.push_back(b);
    }
    gsdr_pool* raw = pool.release();
    raw->grower = std::thread([raw] { raw->grow_loop(); });
    return raw;
}

gsdr_float2* gsdr_pool_get(gsdr_pool* pool) {
    if (!pool) return nullptr;
    std::unique_lock<std::mutex> lk(pool->m);
    if (pool->closing) return nullptr;
    ++pool->waiters;
    if ((double)pool->free_list.size() - 1 < pool->pipe_size / 10.) pool->cv_grow.notify_one();
    pool->cv_free.wait(lk, [&] { return pool->closing || !pool->free_list.empty(); });
    gsdr_float2* b = nullptr;
    if (!pool->closing && !pool->free_list.empty()) {
        b = pool->free_list.front();
        pool->free_list.pop_front();
    }
    if (--pool->waiters == 0 && pool->closing) pool->cv_idle.notify_all();
    return b;
}

void gsdr_pool_trash(gsdr_pool* pool, gsdr_float2* buf) {
    if (!pool || !buf) return;
    {
        std::lock_guard<std::mutex> lk(pool->m);
        pool->free_list.push_back(buf);
    }
    pool->cv_free.notify_one();
}

void gsdr_pool_close(gsdr_pool* pool) {
    if (!pool) return;
    {
        std::lock_guard<std::mutex> lk(pool->m);
        pool->closing = true;
    }
    pool->cv_grow.notify_all();
    pool->cv_free.notify_all();
    if (pool->grower.joinable()) pool->grower.join();
    {   // getters woken by `closing` still hold or re-acquire the mutex: let them leave before the pool goes away
        std::unique_lock<std::mutex> lk(pool->m);
        pool->cv_idle.wait(lk, [&] { return pool->waiters == 0; });
    }
    for (auto* b : pool->all) gsdr_pool::free_one(b);
    delete pool;
}

int gsdr_pool_available(const gsdr_pool* pool) {
    if (!pool) return 0;
    std::lock_guard<std::mutex> lk(const_cast<gsdr_pool*>(pool)->m);
    return (int)pool->free_list.size();
}
int gsdr_pool_size(const gsdr_pool* pool) {
    if (!pool) return 0;
    std::lock_guard<std::mutex> lk(const_cast<gsdr_pool*>(pool)->m);
    return (int)pool->all.size();
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Replay source.  Model: hardware_manager::software_rx_thread (cpp/USRP_hardware_manager.cpp:
// 1331-1395): take a buffer from the RX pool, fill it, wrap it in an RX_wrapper with a running
// packet_number, errors = 0 and the software front-end code ('B'/'D').  The samples come from a
// gsdr_tx of the same parameters (the --sw_loop identity: RX sees exactly the TX waveform),
// optionally plus deterministic complex Gaussian noise.
// ------------------------------------------------------------------------------------------------
struct gsdr_replay {
    gsdr_tx* tx = nullptr;
    gsdr_pool* pool = nullptr;
    long long L = 0;
    int channels = 0;
    char code = 'B';
    float sigma = 0.f;
    std::vector<gsdr_float2> noise;  // bank of L + kBankExtra samples
    uint64_t packets = 0;
    double rate_limit = 0.0;  // samples per second, 0 = unlimited
    std::chrono::steady_clock::time_point start;
    std::vector<gsdr_float2> chirp_tmp;
    static constexpr size_t kBankExtra = 65536;
};

extern "C" {

gsdr_replay* gsdr_replay_create(const gsdr_param* p, int kind, float noise_sigma, uint64_t seed, gsdr_pool* pool,
                                char front_end_code, double rate_limit_msps, int device) {
    if (!p || !pool) {
        set_error("gsdr_replay_create: null argument");
        return nullptr;
    }
    if (pool->vector_size < p->buffer_len) {
        set_error("gsdr_replay_create: pool buffers (%zu) shorter than buffer_len (%llu)", pool->vector_size,
                  (unsigned long long)p->buffer_len);
        return nullptr;
    }
    // the source waveform: TONES for every multi-tone RX mode, CHIRP for chirp
    gsdr_param q = *p;
    std::vector<int32_t> wt(p->n_wave_type ? p->n_wave_type : 1,
                            (p->n_wave_type && p->wave_type[0] == GSDR_CHIRP) ? GSDR_CHIRP : GSDR_TONES);
    std::vector<float> ampl;
    q.wave_type = wt.data();
    q.n_wave_type = wt.size();
    if (p->n_ampl < wt.size()) {  // RX parameter blocks often carry no amplitudes: use 1/T each
        ampl.assign(wt.size(), 1.0f / (float)wt.size());
        q.ampl = ampl.data();
        q.n_ampl = ampl.size();
    }
    std::unique_ptr<gsdr_replay> r(new gsdr_replay());
    r->tx = gsdr_tx_create(&q, device);
    if (!r->tx) return nullptr;
    r->pool = pool;
    r->L = (long long)p->buffer_len;
    r->channels = (int)p->n_wave_type;
    r->code = front_end_code;
    r->sigma = (kind == GSDR_REPLAY_TX_LOOP) ? 0.f : noise_sigma;
    r->rate_limit = rate_limit_msps > 0 ? rate_limit_msps * 1e6 : 0.0;
    if (r->sigma > 0.f) {
        std::mt19937_64 gen(seed);
        std::normal_distribution<float> nd(0.f, r->sigma);
        r->noise.resize((size_t)r->L + gsdr_replay::kBankExtra);
        for (auto& v : r->noise) {
            v.x = nd(gen);
            v.y = nd(gen);
        }
    }
    r->start = std::chrono::steady_clock::now();
    return r.release();
}

int gsdr_replay_next(gsdr_replay* r, gsdr_rx_packet* pkt) {
    if (!r || !pkt) return -1;
    if (r->rate_limit > 0) {  // pace like a radio delivering rate_limit samples per second
        const double due = (double)((r->packets + 1) * (uint64_t)r->L) / r->rate_limit;
        std::this_thread::sleep_until(r->start + std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                                                      std::chrono::duration<double>(due)));
    }
    gsdr_float2* buf = gsdr_pool_get(r->pool);
    if (!buf) {
        set_error("gsdr_replay_next: pool closed");
        return -1;
    }
    gsdr_float2* src = buf;
    if (gsdr_tx_get(r->tx, &src)) {  // TONES re-points src; CHIRP fills buf
        gsdr_pool_trash(r->pool, buf);
        return -1;
    }
    if (src != buf) std::memcpy(buf, src, sizeof(gsdr_float2) * r->L);
    if (r->sigma > 0.f) {
        const gsdr_float2* nz = r->noise.data() + (r->packets * 7919ull) % gsdr_replay::kBankExtra;
        for (long long i = 0; i < r->L; ++i) {
            buf[i].x += nz[i].x;
            buf[i].y += nz[i].y;
        }
    }
    ++r->packets;
    pkt->buffer = buf;
    pkt->usrp_number = 0;
    pkt->front_end_code = r->code;
    pkt->packet_number = (int32_t)r->packets;
    pkt->length = (int32_t)r->L;
    pkt->errors = 0;
    pkt->channels = r->channels;
    return 0;
}

void gsdr_replay_destroy(gsdr_replay* r) {
    if (!r) return;
    gsdr_tx_destroy(r->tx);
    delete r;
}
uint64_t gsdr_replay_packets(const gsdr_replay* r) { return r ? r->packets : 0; }

// ------------------------------------------------------------------------------------------------
// misc C-ABI
// ------------------------------------------------------------------------------------------------
const char* gsdr_last_error(void) { return get_error(); }
const char* gsdr_version(void) { return "gsdr-b200 0.1 (sm_100a)"; }
int gsdr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int gsdr_sm_count(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    return prop.multiProcessorCount;
}

int gsdr_make_sinc_window(int length, float fc, float* out) {
    if (length < 1 || !out) return -1;
    make_sinc_window(length, fc, out);
    return length;
}
int gsdr_make_flat_window(int length, int side, float* out) {
    if (length < 1 || side < 0 || side > length || !out) return -1;
    make_flat_window(length, side, out);
    return length;
}
int gsdr_pfb_batching(int buffer_len, int fft_tones, int pf_average) { return pfb_batching(buffer_len, fft_tones, pf_average); }
int gsdr_tone_bins(int rate, int fft_tones, const int32_t* freq, int n, int32_t* bins) {
    if (fft_tones < 1 || n < 0 || !freq || !bins) return -1;
    tone_bins(rate, fft_tones, freq, n, bins);
    return n;
}
int gsdr_pfb_partition(const int32_t* n_frames, int n_jobs, int grid, int32_t* tiles_out, int cap_tiles, int32_t* cta_begin_out) {
    if (!n_frames || n_jobs <= 0 || grid <= 0 || !tiles_out || !cta_begin_out) return -1;
    std::vector<PfbTile> tiles;
    std::vector<int> cb;
    pfb_partition(n_frames, n_jobs, grid, tiles, cb);
    if ((int)tiles.size() > cap_tiles) return -1;
    for (size_t i = 0; i < tiles.size(); ++i) {
        tiles_out[4 * i] = tiles[i].job, tiles_out[4 * i + 1] = tiles[i].fa, tiles_out[4 * i + 2] = tiles[i].fb, tiles_out[4 * i + 3] = tiles[i].flags;
    }
    for (int c = 0; c <= grid; ++c) cta_begin_out[c] = cb[c];
    return (int)tiles.size();
}
int gsdr_group_form_simulate(const uint8_t* busy, int n_submits, double seconds_copied, double seconds_zero_copy, int wait_lag,
                             int8_t* form_out) {
    if (!busy || !form_out || n_submits < 0 || wait_lag < 0) return -2;
    GroupAutoForm af;
    std::vector<int> blocks((size_t)n_submits, -1);
    std::vector<int> forms((size_t)n_submits, 1);
    for (int i = 0; i < n_submits; ++i) {
        if (i - wait_lag >= 0 && blocks[i - wait_lag] >= 0)   // the wait on an earlier ticket returns before this submit
            af.on_wait(blocks[i - wait_lag], forms[i - wait_lag] ? seconds_zero_copy : seconds_copied);
        forms[i] = af.on_submit(busy[i] != 0, &blocks[i]);
        form_out[i] = (int8_t)forms[i];
    }
    return af.choice;
}
int gsdr_pfb_gather_layout(const int32_t* bins, int n_tones, uint8_t* pos_out) {
    if (n_tones < 0 || n_tones > 2048 || !pos_out) return -1;
    pfb_gather_coloring(bins, n_tones, pos_out);
    return n_tones;
}
void gsdr_buffer_helper_init(gsdr_buffer_helper* h, int n_tones, int buffer_len, int average, int n_eff) {
    buffer_helper_init(h, n_tones, buffer_len, average, n_eff);
}
void gsdr_buffer_helper_update(gsdr_buffer_helper* h) { buffer_helper_update(h); }
void gsdr_vna_helper_init(gsdr_vna_helper* h, int ppt, int buffer_len) { vna_helper_init(h, ppt, buffer_len); }
void gsdr_vna_helper_update(gsdr_vna_helper* h) { vna_helper_update(h); }
int gsdr_chirp_params(int rate, int freq0, int chirp_f0, int swipe_s0, float chirp_t0, int tx, gsdr_chirp_param* out) {
    if (!out) return -1;
    chirp_params(rate, freq0, chirp_f0, swipe_s0, chirp_t0, tx != 0, out);
    return 0;
}

int gsdr_probe_chirp_index(int device, const gsdr_chirp_param* p, uint64_t last_index, uint32_t n, int32_t* out_host) {
    if (!p || !out_host || p->num_steps * p->length == 0) {
        set_error("gsdr_probe_chirp_index: bad argument");
        return -1;
    }
    GSDR_CUDA_OK(cudaSetDevice(device));
    ChirpDev cd{p->num_steps * p->length, p->length, p->chirpness, p->f0, p->num_steps};
    int* d = nullptr;
    GSDR_CUDA_OK(cudaMalloc(&d, sizeof(int) * (n ? n : 1)));
    const int rc = chirp_index_probe_launch(d, n, last_index, cd, 0);
    if (rc >= 0) {
        if (cudaMemcpy(out_host, d, sizeof(int) * n, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("gsdr_probe_chirp_index: %s", cudaGetErrorString(cudaGetLastError()));
            cudaFree(d);
            return -1;
        }
    }
    cudaFree(d);
    return rc < 0 ? -1 : 0;
}

int gsdr_probe_direct_phase(int device, int tone_freq, int rate, uint64_t index_counter, uint64_t n0, uint32_t n,
                            int64_t* out_host) {
    if (!out_host || rate <= 0) {
        set_error("gsdr_probe_direct_phase: bad argument");
        return -1;
    }
    GSDR_CUDA_OK(cudaSetDevice(device));
    long long* d = nullptr;
    GSDR_CUDA_OK(cudaMalloc(&d, sizeof(long long) * (n ? n : 1)));
    const int rc = direct_phase_probe_launch(d, n, tone_freq, rate, index_counter, n0, 0);
    if (rc >= 0) {
        if (cudaMemcpy(out_host, d, sizeof(long long) * n, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("gsdr_probe_direct_phase: %s", cudaGetErrorString(cudaGetLastError()));
            cudaFree(d);
            return -1;
        }
    }
    cudaFree(d);
    return rc < 0 ? -1 : 0;
}

int gsdr_probe_direct_tile_phase(int device, int tone_freq, int rate, int64_t pos0, int64_t row0, int M, int n_rows,
                                 int64_t* phase_out_host, uint32_t* word_out_host) {
    if (!phase_out_host || !word_out_host || rate <= 0 || M <= 0 || n_rows <= 0 || n_rows > 128) {
        set_error("gsdr_probe_direct_tile_phase: bad argument");
        return -1;
    }
    GSDR_CUDA_OK(cudaSetDevice(device));
    long long* d_ph = nullptr;
    unsigned int* d_w = nullptr;
    GSDR_CUDA_OK(cudaMalloc(&d_ph, sizeof(long long) * n_rows));
    if (cudaMalloc(&d_w, sizeof(unsigned int) * n_rows) != cudaSuccess) {
        cudaFree(d_ph);
        set_error("gsdr_probe_direct_tile_phase: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    int rc = direct_tile_phase_probe_launch(d_ph, d_w, n_rows, tone_freq, rate, pos0, row0, M, 0);
    if (rc >= 0 && (cudaMemcpy(phase_out_host, d_ph, sizeof(long long) * n_rows, cudaMemcpyDeviceToHost) != cudaSuccess ||
                    cudaMemcpy(word_out_host, d_w, sizeof(unsigned int) * n_rows, cudaMemcpyDeviceToHost) != cudaSuccess)) {
        set_error("gsdr_probe_direct_tile_phase: %s", cudaGetErrorString(cudaGetLastError()));
        rc = -1;
    }
    cudaFree(d_ph);
    cudaFree(d_w);
    return rc < 0 ? -1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Data-socket framing.  Reference: Sync_server::format_net_buffer (cpp/USRP_server_network.cpp:164-191) packs
// usrp_number | front_end_code | packet_number | length | errors | channels (4+1+4+4+4+4 = 21 bytes, host byte order,
// no padding) and then memcpy's `length` float2 behind it into a malloc'd staging buffer for every packet
// (cpp/USRP_server_network.cpp:187,215).  The client reads it with the packed dtype of pyUSRP/USRP_low_level.py:63-70.
// ------------------------------------------------------------------------------------------------
int gsdr_packet_header_write(const gsdr_rx_packet* pkt, uint8_t* dst) {
    if (!pkt || !dst) return -1;
    std::memcpy(dst + 0, &pkt->usrp_number, 4);
    std::memcpy(dst + 4, &pkt->front_end_code, 1);
    std::memcpy(dst + 5, &pkt->packet_number, 4);
    std::memcpy(dst + 9, &pkt->length, 4);
    std::memcpy(dst + 13, &pkt->errors, 4);
    std::memcpy(dst + 17, &pkt->channels, 4);
    return GSDR_PACKET_HEADER_BYTES;
}
int gsdr_packet_header_read(const uint8_t* src, gsdr_rx_packet* pkt) {
    if (!pkt || !src) return -1;
    std::memcpy(&pkt->usrp_number, src + 0, 4);
    std::memcpy(&pkt->front_end_code, src + 4, 1);
    std::memcpy(&pkt->packet_number, src + 5, 4);
    std::memcpy(&pkt->length, src + 9, 4);
    std::memcpy(&pkt->errors, src + 13, 4);
    std::memcpy(&pkt->channels, src + 17, 4);
    return GSDR_PACKET_HEADER_BYTES;
}
int gsdr_packet_frame(const gsdr_rx_packet* pkt, const void** frame, size_t* frame_bytes) {
    if (!pkt || !pkt->buffer || !frame || !frame_bytes || pkt->length < 0) {
        set_error("gsdr_packet_frame: bad argument");
        return -1;
    }
    uint8_t* hdr = reinterpret_cast<uint8_t*>(pkt->buffer) - GSDR_PACKET_HEADER_BYTES;  // inside the pool headroom
    gsdr_packet_header_write(pkt, hdr);
    *frame = hdr;
    *frame_bytes = GSDR_PACKET_HEADER_BYTES + sizeof(gsdr_float2) * (size_t)pkt->length;
    return 0;
}

// What plain cudaMemcpyAsync between pinned host memory and `device` delivers on this platform: the ceiling every host-fed
// figure is measured against.  One cudaMemcpyAsync per buffer (h2d_bytes up / d2h_bytes down, `reps` buffers each way), three
// passes: upload alone, download alone, both directions at once on two streams.  out_gbs = {h2d, d2h, duplex h2d, duplex d2h}
// in 1e9 bytes/s.  Ranks of a multi-GPU job call it between barriers, so the numbers include what the host shares out.
// The probe is an object so that several processes (one per GPU) can do their allocations first, meet at a barrier, and then run
// each timed pass at the same moment: pinning memory takes longer than a short pass, and ranks that drift apart measure an
// uncontended link.
struct gsdr_pcie_probe {
    int device = 0, n_streams = 1, n_bufs = 0;
    size_t h2d_bytes = 0, d2h_bytes = 0;
    std::vector<void*> h_in, h_out, d_in, d_out;
    cudaStream_t s_up[8] = {nullptr}, s_dn[8] = {nullptr};
};
void gsdr_pcie_probe_destroy(gsdr_pcie_probe* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (void* q : p->h_in) cudaFreeHost(q);
    for (void* q : p->h_out) cudaFreeHost(q);
    for (void* q : p->d_in) cudaFree(q);
    for (void* q : p->d_out) cudaFree(q);
    for (int k = 0; k < 8; ++k) {
        if (p->s_up[k]) cudaStreamDestroy(p->s_up[k]);
        if (p->s_dn[k]) cudaStreamDestroy(p->s_dn[k]);
    }
    delete p;
}
gsdr_pcie_probe* gsdr_pcie_probe_create(int device, size_t h2d_bytes, size_t d2h_bytes, int n_streams, int n_bufs) {
    if (n_streams < 1 || n_streams > 8 || n_bufs < 1 || n_bufs > 4096 || (h2d_bytes == 0 && d2h_bytes == 0)) {
        set_error("gsdr_pcie_probe_create: bad argument");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("gsdr_pcie_probe_create: cudaSetDevice(%d): %s", device, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    auto* p = new gsdr_pcie_probe;
    p->device = device, p->n_streams = n_streams, p->n_bufs = n_bufs, p->h2d_bytes = h2d_bytes, p->d2h_bytes = d2h_bytes;
    bool ok = true;
    for (int k = 0; k < n_streams && ok; ++k)
        ok = cudaStreamCreateWithFlags(&p->s_up[k], cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&p->s_dn[k], cudaStreamNonBlocking) == cudaSuccess;
    // device side: a few buffers are enough (HBM is never the limit here); host side: n_bufs distinct pinned buffers each way
    const int n_dev = n_bufs < 8 ? n_bufs : 8;
    for (int i = 0; i < n_bufs && ok; ++i) {
        void *hi = nullptr, *ho = nullptr;
        ok = pinned_alloc_local(&hi, h2d_bytes ? h2d_bytes : 1) == cudaSuccess;
        if (ok) p->h_in.push_back(hi), memset(hi, 1, h2d_bytes ? h2d_bytes : 1);
        ok = ok && pinned_alloc_local(&ho, d2h_bytes ? d2h_bytes : 1) == cudaSuccess;
        if (ok) p->h_out.push_back(ho), memset(ho, 0, d2h_bytes ? d2h_bytes : 1);
    }
    for (int i = 0; i < n_dev && ok; ++i) {
        void *di = nullptr, *dout = nullptr;
        ok = cudaMalloc(&di, h2d_bytes ? h2d_bytes : 1) == cudaSuccess;
        if (ok) p->d_in.push_back(di);
        ok = ok && cudaMalloc(&dout, d2h_bytes ? d2h_bytes : 1) == cudaSuccess;
        if (ok) p->d_out.push_back(dout);
    }
    if (!ok) {
        set_error("gsdr_pcie_probe_create: %s", cudaGetErrorString(cudaGetLastError()));
        gsdr_pcie_probe_destroy(p);
        return nullptr;
    }
    return p;
}
// One timed pass: `reps` buffers in each enabled direction, one cudaMemcpyAsync per buffer, dealt round-robin over the queues and
// over the host buffers.  Both rates are taken over the SAME interval (until the last copy of either direction has finished), so
// with both directions on they are the rate of (upload, download) PAIRS in the workload's own byte ratio.  out_gbs[2] = {h2d, d2h}
// in 1e9 bytes/s (0 for a direction that was off).
int gsdr_pcie_probe_run(gsdr_pcie_probe* p, int up, int dn, int reps, double* out_gbs) {
    if (!p || !out_gbs || reps < 1 || (!up && !dn)) {
        set_error("gsdr_pcie_probe_run: bad argument");
        return -1;
    }
    GSDR_CUDA_OK(cudaSetDevice(p->device));
    up = up && p->h2d_bytes, dn = dn && p->d2h_bytes;
    const int n_dev = (int)p->d_in.size();
    for (int k = 0; k < p->n_streams; ++k) {
        GSDR_CUDA_OK(cudaStreamSynchronize(p->s_up[k]));
        GSDR_CUDA_OK(cudaStreamSynchronize(p->s_dn[k]));
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) {
        if (up) GSDR_CUDA_OK(cudaMemcpyAsync(p->d_in[i % n_dev], p->h_in[i % p->n_bufs], p->h2d_bytes, cudaMemcpyHostToDevice, p->s_up[i % p->n_streams]));
        if (dn) GSDR_CUDA_OK(cudaMemcpyAsync(p->h_out[i % p->n_bufs], p->d_out[i % n_dev], p->d2h_bytes, cudaMemcpyDeviceToHost, p->s_dn[i % p->n_streams]));
    }
    for (int k = 0; k < p->n_streams; ++k) {
        GSDR_CUDA_OK(cudaStreamSynchronize(p->s_up[k]));
        GSDR_CUDA_OK(cudaStreamSynchronize(p->s_dn[k]));
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out_gbs[0] = up && sec > 0 ? (double)p->h2d_bytes * reps / (sec * 1e9) : 0.0;
    out_gbs[1] = dn && sec > 0 ? (double)p->d2h_bytes * reps / (sec * 1e9) : 0.0;
    return 0;
}
// Single-process convenience forms: out_gbs[4] = {h2d alone, d2h alone, h2d and d2h with both running}.
int gsdr_pcie_copy_ceiling_streams(int device, size_t h2d_bytes, size_t d2h_bytes, int reps, int n_streams, double* out_gbs) {
    if (!out_gbs || reps < 1) {
        set_error("gsdr_pcie_copy_ceiling: bad argument");
        return -1;
    }
    gsdr_pcie_probe* p = gsdr_pcie_probe_create(device, h2d_bytes, d2h_bytes, n_streams, 8);
    if (!p) return -1;
    double w[2], a[2] = {0, 0}, b[2] = {0, 0}, c[2] = {0, 0};
    int rc = gsdr_pcie_probe_run(p, 1, 1, 2 * n_streams, w);   // warm-up
    if (!rc && h2d_bytes) rc = gsdr_pcie_probe_run(p, 1, 0, reps, a);
    if (!rc && d2h_bytes) rc = gsdr_pcie_probe_run(p, 0, 1, reps, b);
    if (!rc) rc = gsdr_pcie_probe_run(p, 1, 1, reps, c);
    gsdr_pcie_probe_destroy(p);
    out_gbs[0] = a[0], out_gbs[1] = b[1], out_gbs[2] = c[0], out_gbs[3] = c[1];
    return rc;
}
int gsdr_pcie_copy_ceiling(int device, size_t h2d_bytes, size_t d2h_bytes, int reps, double* out_gbs) {
    return gsdr_pcie_copy_ceiling_streams(device, h2d_bytes, d2h_bytes, reps, 1, out_gbs);
}

void* gsdr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (pinned_alloc_local(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_error("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    host_registry_add(p, bytes ? bytes : 1);
    return p;
}
int gsdr_device_numa_node(int device) { return numa_node_of_device(device); }
void gsdr_host_free(void* p) {
    if (!p) return;
    host_registry_remove(p);
    cudaFreeHost(p);
}
void* gsdr_dev_alloc(int device, size_t bytes) {
    void* p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_error("cudaMalloc(%zu) on device %d: %s", bytes, device, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void gsdr_dev_free(int device, void* p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}
int gsdr_memcpy_h2d(int device, void* dst, const void* src, size_t bytes) {
    GSDR_CUDA_OK(cudaSetDevice(device));
    GSDR_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int gsdr_memcpy_d2h(int device, void* dst, const void* src, size_t bytes) {
    GSDR_CUDA_OK(cudaSetDevice(device));
    GSDR_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int gsdr_dev_memset(int device, void* dst, int value, size_t bytes) {
    GSDR_CUDA_OK(cudaSetDevice(device));
    GSDR_CUDA_OK(cudaMemset(dst, value, bytes));
    return 0;
}
int gsdr_device_synchronize(int device) {
    GSDR_CUDA_OK(cudaSetDevice(device));
    GSDR_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
