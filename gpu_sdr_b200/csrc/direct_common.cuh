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

// Tile form used by the tensor-core kernels (direct_tc_kernels.cu, direct_i8_kernels.cu).  The reference's LO phase of the
// output whose window starts at stream sample n is (f * n) % rate in integers (cpp/kernels.cu:59-75).  For the rows of one
// tile, n = pos0 + (row0 + r) * M, so phase(r) = base + r * step (mod rate) with base and step reduced exactly in 64-bit
// integers; base + r * step < 129 * rate < 2^53 is then formed exactly by ONE DFMA, and a second DFMA against 2^32 / rate with
// the 1.5 * 2^52 magic addend leaves the 32-bit phase word (whole turns drop out mod 2^32) in the low half of the result.
__device__ __forceinline__ double2 lo_phase_tile(int freq, int rate, long long pos0, long long row0, int M) {
    long long tf = (long long)freq % rate;
    if (tf < 0) tf += rate;   // same residue class as the reference's signed remainder
    long long n0 = (pos0 + row0 * (long long)M) % rate;
    if (n0 < 0) n0 += rate;
    const double base = (double)(((unsigned long long)tf * (unsigned long long)n0) % (unsigned long long)rate);
    const double step = (double)(((unsigned long long)tf * (unsigned long long)(M % rate)) % (unsigned long long)rate);
    return make_double2(base, step);
}
__device__ __forceinline__ double lo_phase_row(double2 bs, double row) { return fma(row, bs.y, bs.x); }
__device__ __forceinline__ unsigned int lo_phase_word(double2 bs, double row, double word_per_phase) {
    return (unsigned int)__double2loint(fma(lo_phase_row(bs, row), word_per_phase, 6755399441055744.0));
}

}  // namespace gsdr
