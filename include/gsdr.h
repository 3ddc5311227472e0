/*
 * gsdr.h -- C-ABI of the B200-native RX/TX readout DSP path (libgsdr.so).
 *
 * This is the drop-in boundary for GPU_SDR's per-buffer hot path.  Every entry point names the
 * reference interface it replaces (paths relative to the reference tree):
 *
 *   RX_buffer_demodulator(param*, bool)           headers/USRP_demodulator.hpp:24   -> gsdr_rx_create
 *   int RX_buffer_demodulator::process(in, out)   headers/USRP_demodulator.hpp:29   -> gsdr_rx_process
 *   void RX_buffer_demodulator::close()           headers/USRP_demodulator.hpp:32   -> gsdr_rx_destroy
 *   TX_buffer_generator(param*)                   headers/USRP_buffer_generator.hpp:58 -> gsdr_tx_create
 *   void TX_buffer_generator::get(float2**)       headers/USRP_buffer_generator.hpp:61 -> gsdr_tx_get
 *   void TX_buffer_generator::close()             headers/USRP_buffer_generator.hpp:64 -> gsdr_tx_destroy
 *   preallocator<float2>{get,trash,close}         headers/USRP_server_memory_management.hpp:103-273 -> gsdr_pool_*
 *   struct param                                  headers/USRP_server_settings.hpp:130-167 -> gsdr_param
 *   enum w_type                                   headers/USRP_server_settings.hpp:113 -> gsdr_w_type
 *   make_sinc_window / make_flat_window           cpp/kernels.cu:258-310, 208-253 -> gsdr_make_*_window
 *   buffer_helper / VNA_decimator_helper          cpp/USRP_server_memory_management.cpp:104-156, 30-56
 *                                                                                 -> gsdr_buffer_helper_*, gsdr_vna_helper_*
 *   upload_multitone_parameters (tone -> bin)     cpp/USRP_demodulator.cpp:702-768 -> gsdr_tone_bins, gsdr_pfb_batching
 *   hardware_manager::software_rx_thread          cpp/USRP_hardware_manager.cpp:1331-1395 -> gsdr_replay_*
 *
 * The C++ classes with the reference's exact names and signatures are header-only shims over
 * this ABI (include/gsdr_compat.hpp), so cpp/USRP_server_link_threads.cpp-style callers compile
 * unchanged.  Plain pointers and sizes only; no C++/CUDA/torch types cross the boundary.
 *
 * Conventions: float2 == {float re, im} (8 bytes).  All functions return 0 / non-NULL on
 * success unless stated; on failure gsdr_last_error() describes the error (thread-local).
 * There is NO CPU fallback: every DSP entry point needs a CUDA device and fails loudly without.
 */
#ifndef GSDR_H_INCLUDED
#define GSDR_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsdr_float2 { float x, y; } gsdr_float2;

/* same numeric order as the reference's `enum w_type` */
typedef enum gsdr_w_type {
    GSDR_TONES = 0, GSDR_CHIRP = 1, GSDR_NOISE = 2, GSDR_RAMP = 3,
    GSDR_NODSP = 4, GSDR_SWONLY = 5, GSDR_DIRECT = 6
} gsdr_w_type;

/* POD flattening of the reference's `param` (only the fields the DSP layer reads). */
typedef struct gsdr_param {
    int32_t rate;            /* param::rate  (samples per second, also the DIRECT wavetable length) */
    int32_t fft_tones;       /* param::fft_tones (PFB channels N) */
    uint64_t decim;          /* param::decim */
    uint64_t pf_average;     /* param::pf_average (PFB taps per channel P / DIRECT taps per decimation f) */
    uint64_t buffer_len;     /* param::buffer_len (float2 per transport buffer) */
    uint64_t data_mem_mult;  /* param::data_mem_mult (output pool multiplier; informational) */
    uint64_t samples;        /* param::samples (informational) */
    const int32_t *freq;      uint64_t n_freq;       /* param::freq      [Hz] */
    const float *ampl;        uint64_t n_ampl;       /* param::ampl */
    const int32_t *wave_type; uint64_t n_wave_type;  /* param::wave_type (gsdr_w_type values) */
    const float *chirp_t;     uint64_t n_chirp_t;    /* param::chirp_t   [s] */
    const int32_t *chirp_f;   uint64_t n_chirp_f;    /* param::chirp_f   [Hz] */
    const int32_t *swipe_s;   uint64_t n_swipe_s;    /* param::swipe_s */
} gsdr_param;

const char *gsdr_last_error(void);
const char *gsdr_version(void);
int gsdr_device_count(void);      /* <=0: no usable CUDA device */
int gsdr_sm_count(int device);

/* ------------------------------------------------------------------------------------------
 * RX demodulator (one instance per RF front-end / IQ stream, like RX_buffer_demodulator)
 * ------------------------------------------------------------------------------------------ */
typedef struct gsdr_rx gsdr_rx;
struct gsdr_chirp_param;

/* Unsupported configurations (mixed wave types, >1 chirp, unknown type) return NULL with an
 * error string; the compat shim turns that into the reference's print_error + exit(-1). */
gsdr_rx *gsdr_rx_create(const gsdr_param *p, int device, int diagnostic);
void gsdr_rx_destroy(gsdr_rx *rx);

/* Blocking drop-in for RX_buffer_demodulator::process: `in` = buffer_len float2 in host memory
 * (pinned for full speed; not modified), `out` = host buffer of at least gsdr_rx_max_output()
 * float2.  Returns the number of valid float2 written (all channels, sample-major
 * out[t*channels + c]) or <0 on error.  On return `in` has been fully consumed (the caller may
 * recycle it at once, cpp/USRP_server_link_threads.cpp:669) and out[0..ret) is complete.
 * With both buffers pinned and mapped (pool buffers, gsdr_host_alloc, cudaMallocHost) the fused
 * channelizer (TONES / NOISE, 2048 channels), DIRECT and CHIRP run as kernel launches that read `in`
 * and write `out` in place over PCIe; otherwise the data is copied up and down.  Same results either
 * way (GSDR_PROCESS_ZEROCOPY=0 when the instance is created forces the copied form).
 * Buffers allocated by this library (gsdr_pool, gsdr_host_alloc) are recognised without a driver
 * call; any other pointer is asked of the driver on every call (GSDR_PROCESS_PTRCACHE=1 remembers
 * the answer per address, for callers whose buffers keep their nature while the library is in use). */
int gsdr_rx_process(gsdr_rx *rx, const gsdr_float2 *in, gsdr_float2 *out);

/* Pipelined variant: submit returns immediately after enqueuing H2D | kernel | D2H on the
 * instance's copy-in / compute / copy-out streams and reports the valid length (known on the host
 * from the integer helpers before the GPU runs).  `in` must stay valid until the matching wait
 * (or until gsdr_rx_input_consumed(ticket) returns 1); up to gsdr_rx_pipeline_depth() tickets may
 * be outstanding.  wait blocks until `out` holds the result. */
int gsdr_rx_submit(gsdr_rx *rx, const gsdr_float2 *in, gsdr_float2 *out, int *valid_len);
int gsdr_rx_wait(gsdr_rx *rx, int ticket);
int gsdr_rx_input_consumed(gsdr_rx *rx, int ticket);
int gsdr_rx_pipeline_depth(const gsdr_rx *rx);

/* sc16 ingest.  `in_iq` = buffer_len interleaved int16 I/Q pairs, the USRP wire
 * format; the sc16 -> fc32 conversion that UHD does on the host CPU for the reference
 * (stream_args_t("fc32"), cpp/USRP_hardware_manager.cpp:764-820; scale 1/32767) runs on the GPU instead, so half
 * the bytes cross PCIe.  Results equal gsdr_rx_process on (float)iq * (1.0f/32767.0f) bit for bit.  Same ticket /
 * wait protocol and output contract as gsdr_rx_submit / gsdr_rx_process. */
int gsdr_rx_submit_sc16(gsdr_rx *rx, const int16_t *in_iq, gsdr_float2 *out, int *valid_len);
int gsdr_rx_process_sc16(gsdr_rx *rx, const int16_t *in_iq, gsdr_float2 *out);

/* Device-resident batched variant (inputs already in HBM): `in_dev` holds n_buffers consecutive
 * transport buffers of this stream (n_buffers*buffer_len float2, contiguous), `out_dev` receives
 * the concatenated valid outputs; valid_lens[i] gets buffer i's valid float2 count (host array,
 * may be NULL).  Asynchronous on the instance's compute stream; returns total valid float2 or <0.
 * State (LO phase, carry-over, FIR tail) advances exactly as n_buffers gsdr_rx_process calls. */
int64_t gsdr_rx_process_device(gsdr_rx *rx, const gsdr_float2 *in_dev, int n_buffers,
                               gsdr_float2 *out_dev, int *valid_lens);
int gsdr_rx_sync(gsdr_rx *rx);
/* Re-arm the stream state (LO phase, carry-over, FIR tail) as freshly constructed. */
int gsdr_rx_reset(gsdr_rx *rx);

int gsdr_rx_channels(const gsdr_rx *rx);            /* == param::wave_type.size() */
int gsdr_rx_mode(const gsdr_rx *rx);                /* gsdr_w_type */
size_t gsdr_rx_max_output(const gsdr_rx *rx);       /* upper bound of one process() return */
size_t gsdr_rx_max_output_batch(const gsdr_rx *rx, int n_buffers);
float gsdr_rx_fcut(const gsdr_rx *rx);              /* RX_buffer_demodulator::fcut */
/* Kernel launches issued so far by this instance (for bench.py's gpu_launches claim). */
uint64_t gsdr_rx_launch_count(const gsdr_rx *rx);
/* CUDA-event timing on the instance's compute stream: start/stop bracket enqueued work. */
int gsdr_rx_timer_start(gsdr_rx *rx);
int gsdr_rx_timer_stop(gsdr_rx *rx, float *elapsed_ms); /* synchronises */
/* Introspection for tests: copy the tap vector / tone bins the instance uses. */
int gsdr_rx_get_taps(const gsdr_rx *rx, float *taps, size_t cap);   /* returns tap count */
int gsdr_rx_get_bins(const gsdr_rx *rx, int32_t *bins, size_t cap); /* returns bin count */
int gsdr_rx_chirp_param(const gsdr_rx *rx, struct gsdr_chirp_param *out);  /* CHIRP instances only */
/* Which kernel serves this instance: e.g. "pfb_fused_wsp_2048_kernel<4>", "direct_fir_tc_kernel". */
const char *gsdr_rx_kernel_name(const gsdr_rx *rx);

/* A group launches ONE persistent kernel over all member streams' frames (TONES / NOISE members on the fused
 * 2048-channel kernel, same buffer_len): the multi-stream path for many concurrent IQ streams per GPU.  It stands where
 * the reference runs one RX_buffer_demodulator per front-end, each fed by its own link thread through the blocking
 * process() (cpp/USRP_server_link_threads.cpp:605-702, call at :666).  Results and state are exactly those of calling
 * gsdr_rx_process / gsdr_rx_process_device on each member.  Members stay owned by the caller and must outlive the group;
 * while a group is in use its members must not be driven through their own entry points.
 *
 * Device-resident: in_dev[i] / out_dev[i] are per-member device pointers, n_buffers consecutive buffers each. */
typedef struct gsdr_rx_group gsdr_rx_group;
gsdr_rx_group *gsdr_rx_group_create(gsdr_rx **members, int n_members);
void gsdr_rx_group_destroy(gsdr_rx_group *g);
int64_t gsdr_rx_group_process_device(gsdr_rx_group *g, const gsdr_float2 *const *in_dev, int n_buffers,
                                     gsdr_float2 *const *out_dev, int *valid_lens /* [member][buffer] */);
/* Host-fed, one packet period per call: in_host[i] = member i's next transport buffer (buffer_len float2, or
 * buffer_len interleaved int16 I/Q pairs for _sc16) and out_host[i] its output buffer (>= gsdr_rx_max_output(member i)
 * float2), both in host memory; valid_lens[i] receives member i's valid float2 count (known at submit time from the
 * integer helpers).  submit returns a ticket after enqueuing the period; up to gsdr_rx_group_pipeline_depth() tickets
 * may be outstanding; wait blocks until every out_host[i] of that ticket is complete; input buffers may be recycled once
 * gsdr_rx_group_input_consumed(ticket) returns 1 (always true after wait).
 * Pinned, mapped buffers (gsdr_pool / gsdr_host_alloc / cudaMallocHost): ONE launch per period reads the S input
 * buffers and writes the S outputs in place over PCIe (GSDR_GROUP_ZEROCOPY=0 at create selects the copied form, =2 the
 * mixed one: copy engine in, kernel stores out).
 * Otherwise: one cudaMemcpyAsync per stream buffer up, one launch, one cudaMemcpyAsync per stream down, on three
 * streams so that consecutive periods overlap.  gsdr_rx_group_process = submit + wait. */
int gsdr_rx_group_submit(gsdr_rx_group *g, const gsdr_float2 *const *in_host, gsdr_float2 *const *out_host, int *valid_lens);
int gsdr_rx_group_submit_sc16(gsdr_rx_group *g, const int16_t *const *in_iq, gsdr_float2 *const *out_host, int *valid_lens);
int gsdr_rx_group_wait(gsdr_rx_group *g, int ticket);
int gsdr_rx_group_input_consumed(gsdr_rx_group *g, int ticket);
int gsdr_rx_group_process(gsdr_rx_group *g, const gsdr_float2 *const *in_host, gsdr_float2 *const *out_host, int *valid_lens);
int gsdr_rx_group_pipeline_depth(const gsdr_rx_group *g);
int gsdr_rx_group_members(const gsdr_rx_group *g);
int gsdr_rx_group_zero_copy(const gsdr_rx_group *g);   /* 1 when the last submit took the zero-copy form */
/* Form of the host-fed call from the next submit on: 0 = copied both ways (one cudaMemcpyAsync per buffer), 1 = zero-copy both
 * ways (default: the launch reads and writes the pinned host buffers in place; the lower-latency form), 2 = inputs by the copy
 * engine, outputs written in place by the kernel's stores, 3 = measured (opt-in): zero-copy as long as the caller feeds packets
 * at their own pace; a caller that keeps the pipeline full (four submits in a row that found the previous period still in
 * flight) gets 8 periods zero-copy and 8 copied timed against each other, and the copied form from then on if it came out at
 * least 3 % faster.  Which form wins depends on how many GPUs share the host (measured: level at 1, zero-copy +16 % at 2,
 * copied +6 % at 4 and +2.5 % at 8); the results are bit-identical.  The short blocks overrate the copied form by a few per
 * cent, which is why mode 3 is not the default.  Forms that need pinned buffers fall back per direction
 * when a buffer of the period is pageable.
 * gsdr_rx_group_auto_choice: -1 while mode 3 has not decided (or another mode is set), else the form it kept (0 or 1),
 * separately for fc32 (sc16 = 0) and sc16 input. */
int gsdr_rx_group_set_zero_copy(gsdr_rx_group *g, int mode);
int gsdr_rx_group_auto_choice(const gsdr_rx_group *g, int sc16);
int gsdr_rx_group_last_form(const gsdr_rx_group *g);   /* of the last submit: bit 0 inputs read in place, bit 1 outputs written in place */
int gsdr_rx_group_sync(gsdr_rx_group *g);
int gsdr_rx_group_timer_start(gsdr_rx_group *g);
int gsdr_rx_group_timer_stop(gsdr_rx_group *g, float *elapsed_ms);
uint64_t gsdr_rx_group_launch_count(const gsdr_rx_group *g);

/* ------------------------------------------------------------------------------------------
 * TX buffer generator
 * ------------------------------------------------------------------------------------------ */
typedef struct gsdr_tx gsdr_tx;
gsdr_tx *gsdr_tx_create(const gsdr_param *p, int device);
void gsdr_tx_destroy(gsdr_tx *tx);
/* Drop-in for TX_buffer_generator::get(float2**): TONES re-points *io into generator-owned
 * pinned host memory (the period buffer); CHIRP fills the caller's buffer *io with buffer_len
 * float2 (blocking). */
int gsdr_tx_get(gsdr_tx *tx, gsdr_float2 **io);
/* Device-resident synthesis of n_buffers consecutive buffers into out_dev (async). */
int gsdr_tx_get_device(gsdr_tx *tx, gsdr_float2 *out_dev, int n_buffers);
int gsdr_tx_sync(gsdr_tx *tx);
int gsdr_tx_dynamic_buffer(const gsdr_tx *tx);   /* param::dynamic_buffer(): 0 for TONES */
int gsdr_tx_buffer_len(const gsdr_tx *tx);
uint64_t gsdr_tx_launch_count(const gsdr_tx *tx);
int gsdr_tx_timer_start(gsdr_tx *tx);
int gsdr_tx_timer_stop(gsdr_tx *tx, float *elapsed_ms);
int gsdr_tx_chirp_param(const gsdr_tx *tx, struct gsdr_chirp_param *out);   /* CHIRP instances only */

/* ------------------------------------------------------------------------------------------
 * Host-side pieces of the path that define results (bit-exact restatements; no GPU needed)
 * ------------------------------------------------------------------------------------------ */
int gsdr_make_sinc_window(int length, float fc, float *taps_out);          /* real taps; imag == 0 */
int gsdr_make_flat_window(int length, int side, float *taps_out);
int gsdr_pfb_batching(int buffer_len, int fft_tones, int pf_average);
int gsdr_tone_bins(int rate, int fft_tones, const int32_t *freq, int n, int32_t *bins_out); /* -1: unmatched */
/* Shared-memory placement the fused 2048-channel kernel uses so that tone_select (cpp/kernels.cu:531-554)
 * reads its bins without bank conflicts: pos_out[bin] = slot 0..15 of the bin inside its 16-bin row
 * (bin & 7, bin >> 7).  bins == NULL means all 2048 bins in order (NOISE).  Diagnostic/test hook. */
int gsdr_pfb_gather_layout(const int32_t *bins, int n_tones, uint8_t *pos_out /* [2048] */);

/* How a multi-stream launch of the fused channelizer (gsdr_rx_group_*) is cut: tiles (job, first frame, end frame, flags:
 * bit 0 = last tile of its stream) never span two streams, CTA c works through tiles [cta_begin[c], cta_begin[c+1]), and
 * every CTA carries the same cost (frames + 24 per tile started).  tiles_out: 4 int32 per tile; cta_begin_out: grid + 1
 * entries.  Returns the number of tiles or -1.  Diagnostic/test hook. */
int gsdr_pfb_partition(const int32_t *n_frames, int n_jobs, int grid, int32_t *tiles_out, int cap_tiles, int32_t *cta_begin_out);

/* The decision logic behind the default form of gsdr_rx_group_submit (mode 3 of gsdr_rx_group_set_zero_copy), run on a
 * scripted sequence: busy[i] = submit i finds the previous period still in flight; a period takes seconds_copied /
 * seconds_zero_copy from submit to the return of its wait, and that wait returns before submit i + wait_lag.  form_out[i] =
 * the form submit i would use (0 copied, 1 zero-copy).  Returns the form kept (0 / 1), -1 while undecided, -2 on a bad
 * argument.  Diagnostic/test hook, no GPU involved. */
int gsdr_group_form_simulate(const uint8_t *busy, int n_submits, double seconds_copied, double seconds_zero_copy, int wait_lag,
                             int8_t *form_out);

typedef struct gsdr_buffer_helper {
    int n_tones, eff_length, buffer_len, average, n_eff_tones;
    int new_0, copy_size, current_batch, spare_samples, spare_begin;
} gsdr_buffer_helper;
void gsdr_buffer_helper_init(gsdr_buffer_helper *h, int n_tones, int buffer_len, int average, int n_eff_tones);
void gsdr_buffer_helper_update(gsdr_buffer_helper *h);

typedef struct gsdr_vna_helper { int valid_size, new0, total_len, spare_begin, ppt, buffer_len; } gsdr_vna_helper;
void gsdr_vna_helper_init(gsdr_vna_helper *h, int ppt, int buffer_len);
void gsdr_vna_helper_update(gsdr_vna_helper *h);

typedef struct gsdr_chirp_param { uint64_t num_steps, length; uint32_t chirpness; int32_t f0; } gsdr_chirp_param;
int gsdr_chirp_params(int rate, int freq0, int chirp_f0, int swipe_s0, float chirp_t0, int tx, gsdr_chirp_param *out);

/* Integer-phase probes: run the SAME device functions the production kernels use and return the
 * raw integers, so tests can require bit-exactness against the reference arithmetic. */
int gsdr_probe_chirp_index(int device, const gsdr_chirp_param *p, uint64_t last_index, uint32_t n, int32_t *index_out_host);
int gsdr_probe_direct_phase(int device, int tone_freq, int rate, uint64_t index_counter, uint64_t n0, uint32_t n,
                            int64_t *phase_out_host);
/* The tile form the tensor-core DIRECT kernels use: for rows r < n_rows (<= 128) of the tile whose first window row is row0
 * (window rows of M samples, stream position pos0 at row 0): phase_out[r] = the unreduced integer phase base + r * step
 * (congruent to f * (pos0 + (row0 + r) * M) modulo rate), word_out[r] = the 32-bit phase word handed to the sin/cos. */
int gsdr_probe_direct_tile_phase(int device, int tone_freq, int rate, int64_t pos0, int64_t row0, int M, int n_rows,
                                 int64_t *phase_out_host, uint32_t *word_out_host);

/* ------------------------------------------------------------------------------------------
 * Client-side Welch spectra: pyUSRP/USRP_noise.py:655-703, spec_from_samples --
 * rotate the IQ plane onto the real axis, optionally scale to and remove the carrier (dBc), clip `clip_samples` at both
 * ends, then the one-sided Welch density of the real and of the imaginary part (nperseg = n / welch, periodic Hann, 50 %
 * overlap, linear detrend), as 10 log10.  `samples`: n float2 in host memory.  Outputs: gsdr_spec_n_freq() values each;
 * `freqs` may be NULL.  Returns the number of frequencies or <0.  welch <= 0: one segment of the whole record.
 * ------------------------------------------------------------------------------------------ */
long long gsdr_spec_n_freq(size_t n, int welch, size_t clip_samples);
int gsdr_spec_from_samples(int device, const gsdr_float2 *samples, size_t n, double sampling_rate, int welch, int dbc, int rotate,
                           size_t clip_samples, double *freqs, float *re_db, float *im_db);

/* ------------------------------------------------------------------------------------------
 * Memory: pinned host pool with the preallocator<float2> contract, plus raw helpers
 * ------------------------------------------------------------------------------------------ */
typedef struct gsdr_pool gsdr_pool;
gsdr_pool *gsdr_pool_create(size_t vector_size /* float2 per buffer */, int pipe_size, int prefill);
gsdr_float2 *gsdr_pool_get(gsdr_pool *pool);              /* blocks until a buffer is free (grows if allowed) */
void gsdr_pool_trash(gsdr_pool *pool, gsdr_float2 *buf);  /* recycle */
void gsdr_pool_close(gsdr_pool *pool);
int gsdr_pool_available(const gsdr_pool *pool);
int gsdr_pool_size(const gsdr_pool *pool);

/* Plain cudaMemcpyAsync ceiling between pinned host memory and `device` (one call per buffer): what every host-fed figure is
 * measured against.  The probe is an object so that the ranks of a multi-GPU job can allocate first, meet at a barrier, and start
 * every timed pass at the same moment (the host's memory and PCIe root are shared).  n_streams (1..8) copy queues per direction,
 * n_bufs distinct pinned host buffers per direction, buffers dealt round-robin.  gsdr_pcie_probe_run: `reps` buffers in each
 * enabled direction; out_gbs[2] = {h2d, d2h} in 1e9 bytes/s, both over the same interval. */
typedef struct gsdr_pcie_probe gsdr_pcie_probe;
gsdr_pcie_probe *gsdr_pcie_probe_create(int device, size_t h2d_bytes, size_t d2h_bytes, int n_streams, int n_bufs);
int gsdr_pcie_probe_run(gsdr_pcie_probe *probe, int up, int dn, int reps, double *out_gbs);
void gsdr_pcie_probe_destroy(gsdr_pcie_probe *probe);
/* single-process convenience forms: out_gbs[4] = {h2d alone, d2h alone, h2d and d2h with both running} */
int gsdr_pcie_copy_ceiling(int device, size_t h2d_bytes, size_t d2h_bytes, int reps, double *out_gbs);
int gsdr_pcie_copy_ceiling_streams(int device, size_t h2d_bytes, size_t d2h_bytes, int reps, int n_streams, double *out_gbs);
void *gsdr_host_alloc(size_t bytes);   /* cudaMallocHost, on the NUMA node of the current GPU when the kernel allows (GSDR_NUMA_LOCAL=0: off) */
int gsdr_device_numa_node(int device); /* /sys/bus/pci/devices/<bus id>/numa_node of the GPU, -1 when unknown */
void gsdr_host_free(void *p);
void *gsdr_dev_alloc(int device, size_t bytes);
void gsdr_dev_free(int device, void *p);
int gsdr_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes);
int gsdr_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes);
int gsdr_dev_memset(int device, void *dst_dev, int value, size_t bytes);
int gsdr_device_synchronize(int device);

/* ------------------------------------------------------------------------------------------
 * Hardware-free replay source (stands where the USRP / --sw_loop thread stands)
 * ------------------------------------------------------------------------------------------ */
typedef struct gsdr_replay gsdr_replay;
typedef struct gsdr_rx_packet {   /* == RX_wrapper, headers/USRP_server_settings.hpp:216-224 */
    gsdr_float2 *buffer; int32_t usrp_number; char front_end_code; int32_t packet_number;
    int32_t length; int32_t errors; int32_t channels;
} gsdr_rx_packet;
/* Data-socket framing.  The wire frame of Sync_server::format_net_buffer
 * (cpp/USRP_server_network.cpp:164-191; client side pyUSRP/USRP_low_level.py:63-70) is a packed 21-byte header
 * followed by `length` float2.  header_write/read are pure.  gsdr_packet_frame builds the frame IN PLACE: every
 * gsdr_pool buffer is preceded by GSDR_POOL_HEADROOM writable bytes, the header goes directly in front of the payload
 * and *frame / *frame_bytes describe one contiguous region to hand to send() -- no per-packet staging copy
 * (the reference memcpy's every payload into a malloc'd buffer, cpp/USRP_server_network.cpp:187,215).
 * pkt->buffer must come from gsdr_pool_get. */
#define GSDR_PACKET_HEADER_BYTES 21
#define GSDR_POOL_HEADROOM 32
int gsdr_packet_header_write(const gsdr_rx_packet *pkt, uint8_t *dst21);
int gsdr_packet_header_read(const uint8_t *src21, gsdr_rx_packet *pkt);
int gsdr_packet_frame(const gsdr_rx_packet *pkt, const void **frame, size_t *frame_bytes);

typedef enum gsdr_replay_kind {
    GSDR_REPLAY_TONES_NOISE = 0,  /* sum of param tones (ampl) + complex gaussian noise of given sigma */
    GSDR_REPLAY_TX_LOOP = 1       /* loop back a gsdr_tx of the same param (the --sw_loop identity) */
} gsdr_replay_kind;
/* Produces packets of p->buffer_len float2 in pinned pool memory, deterministic in `seed`.
 * rate_limit_msps <= 0: as fast as possible; otherwise paced to that sample rate (real time). */
gsdr_replay *gsdr_replay_create(const gsdr_param *p, int kind, float noise_sigma, uint64_t seed,
                                gsdr_pool *pool, char front_end_code, double rate_limit_msps, int device);
int gsdr_replay_next(gsdr_replay *r, gsdr_rx_packet *pkt);  /* 0 ok; fills pkt (buffer from the pool) */
void gsdr_replay_destroy(gsdr_replay *r);
uint64_t gsdr_replay_packets(const gsdr_replay *r);   /* packets produced so far */

#ifdef __cplusplus
}
#endif
#endif /* GSDR_H_INCLUDED */
