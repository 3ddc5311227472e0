// Integer LO phase and phasor shared by the DIRECT kernels (direct_kernels.cu, direct_tc_kernels.cu).
#pragma once
#include "devmath.cuh"

namespace gsdr {

// signed LO phase exactly as the reference forms it (C remainder keeps the dividend's sign)
__device__ __forceinline__ long long direct_phase_signed(long long tf, unsigned long long stream_index, long long R) {
    const long long ii = (long long)(stream_index % (unsigned long long)R);
    return (tf * ii) % R;
}

// e^{-2 pi j ph / R} for ph in [0, R): double divide, 32-bit phase word, exact-reduction sincos
__device__ __forceinline__ float2 lo_phasor(long long ph, double inv_R) {
    const double turns = (double)ph * inv_R;                       // [0,1)
    const unsigned int word = (unsigned int)(long long)(turns * 4294967296.0);
    float s, c;
    sincos_phase32(word, s, c);
    return make_float2(c, -s);
}

}  // namespace gsdr
