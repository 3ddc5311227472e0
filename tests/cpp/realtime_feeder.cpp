// Real-time check of the north-star target at per-GPU scale, from C++ (no Python threads, no GIL): S concurrent 200 MS/s IQ
// streams, each channelized into 1000 tones (cfg2 parameters), delivered in REAL TIME -- one 1e6-sample transport buffer
// every 5 ms per stream, as a USRP would -- from pinned pool buffers, through the library's public C-ABI.
//
//   mode "threads"  the reference's threading model: one worker thread and one demodulator per stream, each calling the
//                   blocking drop-in gsdr_rx_process() on its packet as it arrives (TXRX::rx_single_link,
//                   cpp/USRP_server_link_threads.cpp:605-702).
//   mode "group"    one feeder: every packet period it hands one buffer of every stream to gsdr_rx_group_submit() (ONE launch
//                   per period) and waits for it.
//
// A packet is LATE when its result is not complete in host memory before the NEXT packet of its stream has fully arrived
// (arrival + one period).  Reported: late packets, mean / max latency from arrival to result, sustained yes/no.
// 64 streams on an 8-GPU box = 8 streams per GPU; streams are independent, so one GPU is the unit of proof.
//
// Build: g++ -O2 -std=c++17 -I include tests/cpp/realtime_feeder.cpp -o tests/cpp/bin/realtime_feeder gpu_sdr_b200/libgsdr.so -pthread -Wl,-rpath,'$ORIGIN/../../../gpu_sdr_b200'
// Run:   tests/cpp/bin/realtime_feeder <streams> <seconds> <threads|group> [sc16]
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "gsdr.h"

using clk = std::chrono::steady_clock;

static const int RATE = 200000000, NFFT = 2048, PTAPS = 4, NTONES = 1000, BUFLEN = 1000000;

struct StreamParam {
    std::vector<int32_t> freq, wt;
    std::vector<float> ampl;
    gsdr_param p{};
    explicit StreamParam(int stream) {
        std::mt19937 gen(1337 + 7919 * stream);
        std::vector<int> ks;
        for (int k = -NFFT / 2 + 1; k < NFFT / 2; ++k) ks.push_back(k);
        std::shuffle(ks.begin(), ks.end(), gen);
        for (int i = 0; i < NTONES; ++i) freq.push_back((int32_t)((double)ks[i] * ((double)RATE / NFFT)));
        wt.assign(NTONES, GSDR_TONES);
        ampl.assign(NTONES, 1.0f / NTONES);
        p.rate = RATE;
        p.fft_tones = NFFT;
        p.pf_average = PTAPS;
        p.buffer_len = BUFLEN;
        p.decim = 0;
        p.freq = freq.data(), p.n_freq = freq.size();
        p.ampl = ampl.data(), p.n_ampl = ampl.size();
        p.wave_type = wt.data(), p.n_wave_type = wt.size();
    }
};

static void fill(gsdr_float2* b, int seed) {
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> u(-0.1f, 0.1f);
    for (int i = 0; i < BUFLEN; ++i) b[i].x = u(gen), b[i].y = u(gen);
}

struct Stats {
    long packets = 0, late = 0;
    double lat_sum = 0, lat_max = 0;
    void add(double lat, double period) {
        ++packets;
        lat_sum += lat;
        lat_max = std::max(lat_max, lat);
        if (lat > period) ++late;
    }
};

int main(int argc, char** argv) {
    const int S = argc > 1 ? atoi(argv[1]) : 8;
    const double seconds = argc > 2 ? atof(argv[2]) : 4.0;
    const std::string mode = argc > 3 ? argv[3] : "threads";
    const bool sc16 = argc > 4 && !strcmp(argv[4], "sc16");
    if (gsdr_device_count() <= 0) {
        fprintf(stderr, "no CUDA device: %s\n", gsdr_last_error());
        return 2;
    }
    const double period = (double)BUFLEN / RATE;   // 5 ms
    const long n_packets = (long)(seconds / period);
    const int ring = 3;
    std::vector<StreamParam*> sp;
    std::vector<gsdr_rx*> rx;
    for (int s = 0; s < S; ++s) {
        sp.push_back(new StreamParam(s));
        gsdr_rx* r = gsdr_rx_create(&sp.back()->p, 0, 0);
        if (!r) {
            fprintf(stderr, "gsdr_rx_create: %s\n", gsdr_last_error());
            return 2;
        }
        rx.push_back(r);
    }
    // pinned pools, like the link threads' preallocators (cpp/USRP_server_link_threads.cpp:114,150)
    gsdr_pool* pin = gsdr_pool_create(BUFLEN, S * ring + 2, 0);
    gsdr_pool* pout = gsdr_pool_create(gsdr_rx_max_output(rx[0]), S * ring + 2, 0);
    if (!pin || !pout) {
        fprintf(stderr, "pool: %s\n", gsdr_last_error());
        return 2;
    }
    std::vector<std::vector<gsdr_float2*>> hin(S), hout(S);
    for (int s = 0; s < S; ++s)
        for (int k = 0; k < ring; ++k) {
            hin[s].push_back(gsdr_pool_get(pin));
            hout[s].push_back(gsdr_pool_get(pout));
            if (k == 0 && s < 2) fill(hin[s][0], 100 + s);
            else memcpy(hin[s][k], hin[s & 1][0], sizeof(gsdr_float2) * BUFLEN);
            // sc16 runs reinterpret the first 4 MB of the same bytes as int16 I/Q: any bit pattern is a valid sample
        }
    std::vector<Stats> st(S);
    // warm-up outside the timed run (twiddles, layouts, first launches)
    for (int s = 0; s < S; ++s) {
        if (sc16) gsdr_rx_process_sc16(rx[s], reinterpret_cast<const int16_t*>(hin[s][0]), hout[s][0]);
        else gsdr_rx_process(rx[s], hin[s][0], hout[s][0]);
    }
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    clk::time_point t0 = clk::now() + std::chrono::milliseconds(50);   // stream start (re-armed after the group warm-up)
    auto arrival = [&](long k) { return t0 + std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>((k + 1) * period)); };
    double wall = 0;
    if (mode == "threads") {
        std::vector<std::thread> th;
        for (int s = 0; s < S; ++s)
            th.emplace_back([&, s] {
                for (long k = 0; k < n_packets; ++k) {
                    const clk::time_point arr = arrival(k);   // the buffer is complete when its last sample has arrived
                    std::this_thread::sleep_until(arr);
                    int n;
                    if (sc16) n = gsdr_rx_process_sc16(rx[s], reinterpret_cast<const int16_t*>(hin[s][k % ring]), hout[s][k % ring]);
                    else n = gsdr_rx_process(rx[s], hin[s][k % ring], hout[s][k % ring]);
                    if (n < 0) {
                        fprintf(stderr, "process: %s\n", gsdr_last_error());
                        exit(3);
                    }
                    st[s].add(secs(arr, clk::now()), period);
                }
            });
        for (auto& t : th) t.join();
        wall = secs(t0, clk::now());
    } else {
        gsdr_rx_group* g = gsdr_rx_group_create(rx.data(), S);
        if (!g) {
            fprintf(stderr, "group: %s\n", gsdr_last_error());
            return 2;
        }
        std::vector<const gsdr_float2*> in(S);
        std::vector<gsdr_float2*> out(S);
        std::vector<int> lens(S);
        {   // one untimed period, like the per-stream warm-up above: first launch of the group form, staging buffers
            for (int s = 0; s < S; ++s) in[s] = hin[s][0], out[s] = hout[s][0];
            const int t = sc16 ? gsdr_rx_group_submit_sc16(g, reinterpret_cast<const int16_t* const*>(in.data()), out.data(), lens.data())
                               : gsdr_rx_group_submit(g, in.data(), out.data(), lens.data());
            if (t < 0 || gsdr_rx_group_wait(g, t)) {
                fprintf(stderr, "group warm-up: %s\n", gsdr_last_error());
                return 3;
            }
        }
        t0 = clk::now() + std::chrono::milliseconds(50);
        for (long k = 0; k < n_packets; ++k) {
            const clk::time_point arr = arrival(k);
            std::this_thread::sleep_until(arr);
            for (int s = 0; s < S; ++s) in[s] = hin[s][k % ring], out[s] = hout[s][k % ring];
            const int t = sc16 ? gsdr_rx_group_submit_sc16(g, reinterpret_cast<const int16_t* const*>(in.data()), out.data(), lens.data())
                               : gsdr_rx_group_submit(g, in.data(), out.data(), lens.data());
            if (t < 0 || gsdr_rx_group_wait(g, t)) {
                fprintf(stderr, "group submit/wait: %s\n", gsdr_last_error());
                return 3;
            }
            const double lat = secs(arr, clk::now());   // every stream's result of this period is in host memory
            for (int s = 0; s < S; ++s) st[s].add(lat, period);
        }
        wall = secs(t0, clk::now());
        gsdr_rx_group_destroy(g);
    }
    long packets = 0, late = 0;
    double lat_sum = 0, lat_max = 0;
    for (auto& x : st) packets += x.packets, late += x.late, lat_sum += x.lat_sum, lat_max = std::max(lat_max, x.lat_max);
    printf("{\"test\": \"real-time multi-stream channelizer (cfg2 per stream), C++ feeder\", \"streams\": %d, \"mode\": \"%s\", \"input\": \"%s\", "
           "\"stream_rate_MSps\": %.1f, \"seconds\": %.2f, \"wall_s\": %.3f, \"packets\": %ld, \"late_packets\": %ld, "
           "\"late_definition\": \"result not in host memory before the stream's next packet has arrived (latency > one 5 ms period)\", "
           "\"latency_ms_mean\": %.3f, \"latency_ms_max\": %.3f, \"aggregate_input_MSps\": %.1f, \"required_MSps\": %.1f, \"sustained\": %s}\n",
           S, mode.c_str(), sc16 ? "sc16" : "fc32", RATE / 1e6, seconds, wall, packets, late, 1e3 * lat_sum / std::max(1L, packets), 1e3 * lat_max,
           (double)packets * BUFLEN / wall / 1e6, (double)S * RATE / 1e6, late == 0 ? "true" : "false");
    for (auto r : rx) gsdr_rx_destroy(r);
    gsdr_pool_close(pin);
    gsdr_pool_close(pout);
    return 0;
}
