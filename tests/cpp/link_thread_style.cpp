// A caller written the way the reference's TXRX worker threads are (cpp/USRP_server_link_threads.cpp:
// tx_single_link :542-602, rx_single_link :605-702): pools from preallocator<float2>, a
// TX_buffer_generator feeding a (software-loopback) queue, an RX_buffer_demodulator draining it.
// It compiles with a plain C++ compiler against include/gsdr_compat.hpp + libgsdr.so; run on a GPU
// it prints "ok" when every demodulated frame equals the tone amplitudes (sw-loop identity).
#include <cmath>
#include <cstdio>
#include <deque>

#include "gsdr_compat.hpp"

int main() {
    const int rate = 2048000, N = 2048, L = 100000, T = 8;
    param rxp, txp;
    rxp.mode = RX;
    rxp.rate = rate;
    rxp.fft_tones = N;
    rxp.pf_average = 4;
    rxp.buffer_len = L;
    rxp.decim = 0;
    for (int t = 0; t < T; t++) {
        rxp.freq.push_back((t % 2 ? -1 : 1) * (37 + 11 * t) * 1000);
        rxp.ampl.push_back(1.f / T);
        rxp.wave_type.push_back(TONES);
    }
    txp = rxp;
    txp.mode = TX;

    preallocator<float2>* rx_memory = new preallocator<float2>(L, 8);
    preallocator<float2>* rx_output_memory = new preallocator<float2>(L * 1, 8);
    TX_buffer_generator* generator = new TX_buffer_generator(&txp);
    RX_buffer_demodulator* demodulator = new RX_buffer_demodulator(&rxp);

    std::deque<RX_wrapper> Rx_queue;
    bool ok = true;
    for (int pkt = 0; pkt < 6; pkt++) {
        // tx_single_link: TONES is not a dynamic buffer, get() re-points the pointer
        float2* tx_vector = nullptr;
        if (txp.dynamic_buffer()) tx_vector = rx_memory->get();
        generator->get(&tx_vector);
        // software_rx_thread: copy into a pool buffer and wrap it
        RX_wrapper w;
        w.buffer = rx_memory->get();
        for (int i = 0; i < L; i++) w.buffer[i] = tx_vector[i];
        w.packet_number = pkt + 1;
        w.length = L;
        w.errors = 0;
        w.front_end_code = 'B';
        Rx_queue.push_back(w);
        // rx_single_link
        RX_wrapper rx_buffer = Rx_queue.front();
        Rx_queue.pop_front();
        rx_buffer.channels = demodulator->parameters->wave_type.size();
        float2* output_buffer = rx_output_memory->get();
        rx_buffer.length = demodulator->process(&rx_buffer.buffer, &output_buffer);
        rx_memory->trash(rx_buffer.buffer);
        rx_buffer.buffer = output_buffer;
        for (int i = 0; i < rx_buffer.length; i++) {
            const float mag = std::sqrt(output_buffer[i].x * output_buffer[i].x + output_buffer[i].y * output_buffer[i].y);
            if (std::fabs(mag - 1.f / T) > 1e-4f / T) ok = false;
        }
        if (rx_buffer.length % rx_buffer.channels != 0 || rx_buffer.length == 0) ok = false;
        rx_output_memory->trash(output_buffer);
    }
    demodulator->close();
    generator->close();
    rx_memory->close();
    rx_output_memory->close();
    std::printf(ok ? "ok\n" : "MISMATCH\n");
    return ok ? 0 : 1;
}
