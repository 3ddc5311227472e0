// Host-side logic of the path that *defines results*: FIR/PFB tap builders, carry-over
// bookkeeping, tone->bin mapping and chirp parameter quantisation.  These must agree with the
// reference bit for bit (taps: float32; the rest: integers), so each routine documents the
// reference lines whose arithmetic it reproduces (paths relative to the reference tree).
// Built WITHOUT fast-math / FMA contraction (see Makefile) for the same reason.
#include <algorithm>
#include <cmath>
#include <limits>

#include "common.hpp"

namespace gsdr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static const float kPiF = 3.14159265358979f;  // headers/kernels.cuh:34 (float literal)

// Hamming-windowed sinc prototype, sum-normalised.  cpp/kernels.cu:258-310.
//  - centre index is (length-1)/2 in INTEGER arithmetic (even lengths are off-centre);
//  - sinc and cosine are evaluated in float32 (sinf/cosf), the Hamming factor in double;
//  - the normalising sum is accumulated in float32 in index order.
void make_sinc_window(int length, float fc, float* out) {
    const int centre = (length - 1) / 2;
    float sum = 0.f;
    for (int i = 0; i < length; ++i) {
        const int d = i - centre;
        float tap = 2.f * fc;
        if (d != 0) {
            const float arg = 2.f * kPiF * fc * d;
            tap = (2.f * fc) * sinf(arg) / arg;
        }
        const float c = cosf(2.f * kPiF * i / (length - 1));
        tap = static_cast<float>(static_cast<double>(tap) * (0.54 - 0.46 * static_cast<double>(c)));
        out[i] = tap;
        sum += tap;
    }
    for (int i = 0; i < length; ++i) out[i] /= sum;
}

// Lock-in profile.  cpp/kernels.cu:208-253: only the first `side` taps end up zero (the loop
// that clears the trailing `side` taps runs before the fill loop, which overwrites them).
void make_flat_window(int length, int side, float* out) {
    float sum = 0.f;
    for (int i = 0; i < length; ++i) out[i] = 0.f;
    for (int i = side; i < length; ++i) {
        out[i] = 1.f;
        sum += out[i];
    }
    for (int i = 0; i < length; ++i) out[i] /= sum;
}

// cpp/USRP_demodulator.cpp:706
int pfb_batching(int buffer_len, int fft_tones, int pf_average) {
    return static_cast<int>(std::ceil(static_cast<float>(buffer_len) / static_cast<float>(fft_tones)) + pf_average + 5);
}

// cpp/USRP_demodulator.cpp:722-734.  The reference scans the whole bin axis for every tone and
// keeps the LAST axis point whose open interval (axis-bs, axis+bs) holds the tone, which makes
// the rule ceil(f/bs) rather than round(f/bs); the double comparisons are kept literally so the
// edge cases (exact multiples, truncated negative tones) fall the same way.  -1 = no match.
void tone_bins(int rate, int fft_tones, const int32_t* freq, int n, int32_t* bins) {
    const double bs = static_cast<double>(rate) / static_cast<double>(fft_tones);
    const int half = fft_tones / 2;
    for (int u = 0; u < n; ++u) {
        int32_t hit = -1;
        for (int i = 0; i < fft_tones; ++i) {
            const double axis = i * bs - bs * half;
            if (freq[u] < axis + bs && freq[u] > axis - bs) hit = (i + half) % fft_tones;
        }
        bins[u] = hit;
    }
}

// cpp/USRP_server_memory_management.cpp:104-156
static int frames_available(const gsdr_buffer_helper* h) {
    int frames = 0;
    for (int off = 0; off + h->average * h->n_tones < h->eff_length; off += h->n_tones) ++frames;
    return frames;
}
void buffer_helper_init(gsdr_buffer_helper* h, int n_tones, int buffer_len, int average, int n_eff_tones) {
    h->n_tones = n_tones;
    h->buffer_len = buffer_len;
    h->average = average;
    h->n_eff_tones = n_eff_tones;
    h->eff_length = buffer_len;
    h->new_0 = 0;
    h->current_batch = frames_available(h);
    h->spare_samples = h->eff_length - h->current_batch * n_tones;
    h->spare_begin = h->eff_length - h->spare_samples;
    h->copy_size = n_eff_tones * h->current_batch;
}
void buffer_helper_update(gsdr_buffer_helper* h) {
    h->new_0 = h->spare_samples;
    h->eff_length = h->spare_samples + h->buffer_len;
    h->current_batch = frames_available(h);
    h->copy_size = h->n_eff_tones * h->current_batch;
    h->spare_samples = h->eff_length - h->current_batch * h->n_tones;
    h->spare_begin = h->eff_length - h->spare_samples;
}

// cpp/USRP_server_memory_management.cpp:30-56
static void vna_refresh(gsdr_vna_helper* h) {
    h->valid_size = h->total_len / h->ppt;
    h->new0 = h->total_len - h->ppt * h->valid_size;
    h->spare_begin = h->total_len - h->new0;
}
void vna_helper_init(gsdr_vna_helper* h, int ppt, int buffer_len) {
    h->ppt = ppt;
    h->buffer_len = buffer_len;
    h->total_len = buffer_len;
    vna_refresh(h);
}
void vna_helper_update(gsdr_vna_helper* h) {
    h->total_len = h->buffer_len + h->new0;
    vna_refresh(h);
}

// Conversions with the wrap behaviour the reference gets from gcc/x86-64: double -> unsigned int
// goes through a 64-bit truncation (so a negative chirpness wraps modulo 2^32), double -> int
// out of range gives INT_MIN.
static uint32_t wrap_u32(double v) {
    if (!(v > -9.2e18 && v < 9.2e18)) return 0u;
    return static_cast<uint32_t>(static_cast<uint64_t>(static_cast<int64_t>(v)));
}
static int32_t trunc_i32(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return std::numeric_limits<int32_t>::min();
    return static_cast<int32_t>(v);
}

// RX: cpp/USRP_demodulator.cpp:192-214.  TX: cpp/USRP_buffer_generator.cpp:114-137 (the TX side
// additionally resets num_steps when a step would be shorter than one sample).
void chirp_params(int rate, int freq0, int chirp_f0, int swipe_s0, float chirp_t0, bool tx, gsdr_chirp_param* out) {
    uint64_t steps = static_cast<uint64_t>(static_cast<int64_t>(swipe_s0));
    if (steps < 1) steps = static_cast<uint64_t>(chirp_t0 * static_cast<float>(rate));
    uint64_t len = static_cast<uint64_t>(chirp_t0 * static_cast<float>(rate) / static_cast<float>(steps));
    if (len < 1) {
        len = 1;
        if (tx) steps = static_cast<uint64_t>(chirp_t0 * static_cast<float>(rate));
    }
    const double full = std::pow(2, 32) - 1;
    out->num_steps = steps;
    out->length = len;
    out->chirpness = wrap_u32((full * (chirp_f0 - freq0) / (static_cast<double>(steps) - 1.)) / static_cast<double>(rate));
    out->f0 = trunc_i32(full * (static_cast<double>(freq0) / static_cast<double>(rate)));
}


// Bank-conflict-free tone gather for the fused 2048-channel kernel.
//
// The kernel parks a frame's spectrum in shared memory as rows of 16 bins: row (k1, k3) holds the bins
// k1 + 8 k2 + 128 k3 (k2 = 0..15), 128 bytes, one 8-byte bank pair per bin.  Inside a row the FFT team may
// place the 16 bins in ANY order without hurting its own (half-warp, one row per store) accesses.  The
// producers then read the selected bins 16 at a time (tones 16 h .. 16 h + 15 per half-warp) -- in the
// caller's tone order, i.e. scattered rows.  That read is conflict-free iff its 16 bins sit in 16 different
// bank pairs.  Rows on one side, half-warp reads on the other, selected bins as edges: a bipartite multigraph
// of maximum degree 16, and by Koenig's theorem it always has a proper edge colouring with 16 colours
// (colour = bank pair).  Built with alternating-path recolouring; O(T * 16).
// perm[bin] = position (0..15) of the bin inside its row.  Duplicate tones (same bin twice) share the first
// occurrence's position; they are read by broadcast or, at worst, with a conflict -- never wrongly.
void pfb_gather_coloring(const int32_t* bins, int T, uint8_t* perm /* [2048] */) {
    constexpr int NB = 2048, NL = 128, NC = 16, FIXED = 1 << 29;
    const int nR = (T + 15) / 16;
    struct Edge { int l, r, bin, color; };
    std::vector<Edge> edges;
    edges.reserve(T);
    std::vector<std::vector<int>> reads(NB);  // the 16-tone reads each bin takes part in
    for (int u = 0; u < T; ++u) {
        const int bin = (bins ? bins[u] : u) & (NB - 1);
        if (reads[bin].empty() || reads[bin].back() != (u >> 4)) reads[bin].push_back(u >> 4);
    }
    std::vector<int> colL(NL * NC, -1), colR((size_t)(nR > 0 ? nR : 1) * NC, -1);
    std::vector<int> bin_color(NB, -1);
    // A tone list may name a bin more than once (two tones in one channel).  Such a bin is read by several
    // half-warps but lives in one slot: give it a slot that is free in its row and in all of its reads FIRST,
    // while almost everything is free, and never recolour it afterwards.
    for (int bin = 0; bin < NB; ++bin) {
        if (reads[bin].size() < 2) continue;
        const int x = (bin & 7) * 16 + (bin >> 7);
        int best = -1, best_clash = 1 << 30;
        for (int cdt = 0; cdt < NC; ++cdt) {
            if (colL[x * NC + cdt] >= 0) continue;
            int clash = 0;
            for (int y : reads[bin]) clash += colR[y * NC + cdt] >= 0;
            if (clash < best_clash) best_clash = clash, best = cdt;
        }
        bin_color[bin] = best;
        colL[x * NC + best] = FIXED;
        for (int y : reads[bin])
            if (colR[y * NC + best] < 0) colR[y * NC + best] = FIXED;
    }
    for (int bin = 0; bin < NB; ++bin)
        if (reads[bin].size() == 1) edges.push_back(Edge{(bin & 7) * 16 + (bin >> 7), reads[bin][0], bin, -1});
    // order of the caller's tone list (deterministic, and the test reproduces it)
    std::sort(edges.begin(), edges.end(), [](const Edge& p, const Edge& q) { return p.r != q.r ? p.r < q.r : p.bin < q.bin; });
    std::vector<int> path;
    for (int e = 0; e < (int)edges.size(); ++e) {
        const int x = edges[e].l, y = edges[e].r;
        int chosen = -1;
        // a colour free on both sides needs no work
        for (int a = 0; a < NC && chosen < 0; ++a)
            if (colL[x * NC + a] < 0 && colR[y * NC + a] < 0) chosen = a;
        // otherwise: a free at the row, b free at the read; swap a <-> b along the alternating path that starts at
        // the read with colour a.  Bipartite + a free at x => the path never returns to x.  A path through a
        // pinned (duplicate) bin is not flipped; another (a, b) pair is tried.
        for (int a = 0; a < NC && chosen < 0; ++a) {
            if (colL[x * NC + a] >= 0) continue;
            for (int b = 0; b < NC && chosen < 0; ++b) {
                if (colR[y * NC + b] >= 0) continue;
                path.clear();
                int c = a, node = y;
                bool at_right = true, ok = true;
                for (;;) {
                    const int pe = at_right ? colR[node * NC + c] : colL[node * NC + c];
                    if (pe < 0) break;
                    if (pe >= FIXED) { ok = false; break; }
                    path.push_back(pe);
                    node = at_right ? edges[pe].l : edges[pe].r;
                    at_right = !at_right;
                    c = (c == a) ? b : a;
                }
                if (!ok) continue;
                for (int pe : path) {
                    colL[edges[pe].l * NC + edges[pe].color] = -1;
                    colR[edges[pe].r * NC + edges[pe].color] = -1;
                }
                for (int pe : path) {
                    edges[pe].color = (edges[pe].color == a) ? b : a;
                    colL[edges[pe].l * NC + edges[pe].color] = pe;
                    colR[edges[pe].r * NC + edges[pe].color] = pe;
                }
                chosen = a;
            }
        }
        if (chosen < 0) {  // only reachable around pinned bins: accept one conflict rather than fail
            for (int a = 0; a < NC && chosen < 0; ++a)
                if (colL[x * NC + a] < 0) chosen = a;
        }
        edges[e].color = chosen;
        colL[x * NC + chosen] = e;
        if (colR[y * NC + chosen] < 0) colR[y * NC + chosen] = e;
    }
    for (const Edge& ed : edges) bin_color[ed.bin] = ed.color;
    // unselected bins take the positions their row has left
    for (int l = 0; l < NL; ++l) {
        const int k1 = l / 16, k3 = l % 16;
        int next = 0;
        for (int k2 = 0; k2 < 16; ++k2) {
            const int bin = k1 + 8 * k2 + 128 * k3;
            if (bin_color[bin] >= 0) continue;
            while (colL[l * NC + next] >= 0) ++next;
            bin_color[bin] = next;
            colL[l * NC + next] = FIXED;
        }
    }
    for (int bin = 0; bin < NB; ++bin) perm[bin] = (uint8_t)bin_color[bin];
}

// Multi-stream launch of the fused channelizer: cut the concatenated frame sequence of all streams into tiles (never
// spanning two streams) and hand every CTA a contiguous run of tiles of equal COST.  A tile costs its frames plus
// kPfbTileCost: the pipeline fill and drain and the reload of the stream's constants, measured at ~26 frames' worth
// (64 streams x 1 buffer in one launch: 189 us against 168 us for one stream with the same frames per CTA).  Splitting by
// frames alone put the stream boundaries' extra tiles on the critical path (8 streams x 8 buffers: 0.61 of the HBM roofline
// against 0.68 for one stream).
int GroupAutoForm::on_submit(bool busy, int* block) {
    *block = -1;
    if (choice >= 0) return choice;
    if (!busy) {
        *this = GroupAutoForm();   // paced caller (or pipeline drained): zero-copy, nothing to measure
        return 1;
    }
    if (!measuring && ++busy_run >= kBusy) measuring = true, cal_submits = 0;
    if (!measuring) return 1;
    if (cal_submits >= 2 * kBlock) {
        // every measured period has been submitted: decide with what the waits have reported (block 0 ran zero-copy, block 1
        // copied).  Zero-copy is given up only for a clear win of the copied form (kMargin: five periods per block do not
        // resolve a per cent or two, and zero-copy is the lower-latency form); without a report from both blocks it stays.
        const double m0 = cnt[0] ? sum[0] / cnt[0] : 0.0, m1 = cnt[1] ? sum[1] / cnt[1] : 0.0;
        choice = (cnt[0] && cnt[1] && m1 < kMargin * m0) ? 0 : 1;
        measuring = false;
        return choice;
    }
    const int b = cal_submits / kBlock;
    if (cal_submits % kBlock >= kSkip) *block = b;
    cal_submits++;
    return b == 0 ? 1 : 0;
}
void GroupAutoForm::on_wait(int block, double seconds) {
    if (!measuring || block < 0 || block > 1) return;
    sum[block] += seconds;
    cnt[block]++;
}

void pfb_partition(const int* n_frames, int n_jobs, int grid, std::vector<PfbTile>& tiles, std::vector<int>& cta_begin) {
    tiles.clear();
    cta_begin.assign((size_t)(grid > 0 ? grid : 0) + 1, 0);
    long long total = 0;
    int live_jobs = 0;
    for (int k = 0; k < n_jobs; ++k)
        if (n_frames[k] > 0) total += n_frames[k], ++live_jobs;
    if (grid <= 0 || total == 0) return;
    int j = 0;
    while (n_frames[j] <= 0) ++j;
    long long job_lo = 0, lo = 0;   // global index of job j's first frame, next unassigned frame
    int boundaries_left = live_jobs - 1;
    for (int c = 0; c < grid; ++c) {
        cta_begin[c] = (int)tiles.size();
        if (lo >= total) continue;
        const int ctas_left = grid - c;
        // what is left, shared equally: every remaining CTA starts one tile, every remaining stream boundary adds one
        long long budget = (total - lo + kPfbTileCost * (ctas_left + boundaries_left) + ctas_left - 1) / ctas_left;
        const bool last = c == grid - 1;
        for (;;) {
            const long long job_hi = job_lo + n_frames[j];
            long long take = last ? job_hi - lo : budget - kPfbTileCost;
            if (take > job_hi - lo) take = job_hi - lo;
            if (take < 1) take = 1;
            // no slivers: a remainder shorter than a tile's own cost rides along
            if (!last && job_hi - lo - take > 0 && job_hi - lo - take < kPfbTileCost) take = job_hi - lo;
            const long long e = lo + take;
            tiles.push_back(PfbTile{j, (int)(lo - job_lo), (int)(e - job_lo), e == job_hi ? 1 : 0});
            budget -= kPfbTileCost + take;
            lo = e;
            if (lo >= total) break;
            if (e == job_hi) {
                job_lo = job_hi;
                ++j;
                while (n_frames[j] <= 0) ++j;
                --boundaries_left;
            }
            if (!last && budget < 2 * kPfbTileCost) break;   // not worth starting another tile here
        }
    }
    cta_begin[grid] = (int)tiles.size();
}

}  // namespace gsdr
