// ga_frontend.cuh -- 8-bit IQ front-end on the GPU: what the reference does in MATLAB before gps_test
// can run on an SDR capture (proc_rtl_bin_for_gps.m:31-47 for rtl-sdr uint8, proc_hackrf_bin_for_gps.m:7-19
// for HackRF int8):   y = I + jQ (uint8: minus 128) ;  y = y - mean(y)  (mean over the WHOLE capture) ;
// r[n] = real( y[n] * exp(j*2*pi*fc*n/fs) ) ;  bit = (1 - sign(r))/2 ;  fwrite(...,'ubit1') (LSB first).
// Double precision like MATLAB; the sums for the mean are exact integers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ga {

// pass 1: exact integer sums of I and Q (format 0: uint8 offset 128, format 1: int8)
__global__ void iq8_sum_kernel(const unsigned char *__restrict__ iq, size_t n_samples, int format,
                               long long *__restrict__ sums /* [2] */)
{
    long long si = 0, sq = 0;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_samples; n += (size_t)gridDim.x * blockDim.x) {
        const int a = iq[2 * n], b = iq[2 * n + 1];
        si += format == 0 ? a - 128 : (int)(signed char)a;
        sq += format == 0 ? b - 128 : (int)(signed char)b;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        si += __shfl_down_sync(0xffffffffu, si, off);
        sq += __shfl_down_sync(0xffffffffu, sq, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)&sums[0], (unsigned long long)si);
        atomicAdd((unsigned long long *)&sums[1], (unsigned long long)sq);
    }
}

// pass 2: one thread per output byte (8 samples).  n0 = index of the first sample of this buffer in the
// capture (the phase runs over the whole capture, :42 "(0:(length(y)-1))").
__global__ void iq8_to_bits_kernel(const unsigned char *__restrict__ iq, size_t n_samples, size_t n0, int format,
                                   double mean_i, double mean_q, double fc, double fs, unsigned char *__restrict__ bits)
{
    const size_t byte = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (byte * 8 >= n_samples) return;
    const double w = 2.0 * 3.141592653589793 * fc, inv_fs = 1.0 / fs;     // 2.*pi.*fc ... .*(1./fs), evaluated left to right
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t n = byte * 8 + k;
        if (n >= n_samples) break;
        const int a = iq[2 * n], b = iq[2 * n + 1];
        const double yi = (format == 0 ? a - 128 : (int)(signed char)a) - mean_i;
        const double yq = (format == 0 ? b - 128 : (int)(signed char)b) - mean_q;
        double sn, cs;
        sincos((w * (double)(n0 + n)) * inv_fs, &sn, &cs);
        const double r = yi * cs - yq * sn;
        out |= (r < 0.0 ? 1u : 0u) << k;                                     // (1-sign(r))/2, LSB first
    }
    bits[byte] = (unsigned char)out;
}

// pass 2 with an exactly periodic phase: when fc/fs = p/q in lowest terms with a small q the phasor of sample n is
// table[(p*n) mod q] (double cos/sin made on the host in long double) -- no double-precision sincos of a large
// argument per sample.  The mathematically exact phase; the MATLAB expression differs from it by its own rounding of
// 2*pi*fc*n/fs (~1e-10 rad at n ~ 1e8), which matters only where |r| is that small.
__global__ void iq8_to_bits_table_kernel(const unsigned char *__restrict__ iq, size_t n_samples, size_t n0, int format,
                                         double mean_i, double mean_q, const double2 *__restrict__ table,
                                         unsigned long long p, unsigned long long q, unsigned char *__restrict__ bits)
{
    const size_t byte = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (byte * 8 >= n_samples) return;
    unsigned long long k = (((n0 + byte * 8) % q) * p) % q;            // p, q < 2^31: no overflow
    unsigned out = 0;
    // 16 input bytes per thread: one 128-bit load when the whole group is inside the buffer
    unsigned char raw[16];
    if (byte * 8 + 8 <= n_samples) {
        const uint4 v = *reinterpret_cast<const uint4 *>(iq + 16 * byte);
        *reinterpret_cast<uint4 *>(raw) = v;
    } else {
        for (int i = 0; i < 16; i++) raw[i] = (byte * 16 + i < 2 * n_samples) ? iq[16 * byte + i] : 0;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (byte * 8 + j >= n_samples) break;
        const int a = raw[2 * j], b = raw[2 * j + 1];
        const double yi = (format == 0 ? a - 128 : (int)(signed char)a) - mean_i;
        const double yq = (format == 0 ? b - 128 : (int)(signed char)b) - mean_q;
        const double2 cs = table[k];
        const double r = yi * cs.x - yq * cs.y;
        out |= (r < 0.0 ? 1u : 0u) << j;
        k += p; if (k >= q) k -= q;
    }
    bits[byte] = (unsigned char)out;
}

// ---------------------------------------------------------------------------------------------------------------
// The reverse direction: 1-bit real IF -> interleaved int8 IQ at baseband for HackRF replay, what the reference's
// c/conv_1bit_bin_to_hackrf_bin.cpp:29-86 does:  I = A*Bipolar(bit ^ lo_sin[int(phase)]), Q = A*Bipolar(bit ^
// lo_cos[int(phase)]) with lo_sin = {1,1,0,0}, lo_cos = {1,0,0,1} (:30-31), Bipolar(1) = -A, A = 30 (:17-19); the float
// phase NCO (+= (float)(4*FC/FS), wrap at 4, :33,:79-80) runs on over the WHOLE file.  The float recurrence is run on
// the host (it is eventually periodic: pre-period mu, period lambda, a few million steps at most) and shipped as a
// table lo[k] = lo_sin | lo_cos << 1; sample i uses k = i (i < mu) or mu + (i - mu) mod lambda.
// One thread per input byte: 8 samples -> 16 output bytes, one 128-bit store.
// ---------------------------------------------------------------------------------------------------------------
__global__ void bits_to_iq8_kernel(const unsigned char *__restrict__ bits, size_t n_bytes, size_t first_sample,
                                   const unsigned char *__restrict__ lo, unsigned long long mu, unsigned long long lambda,
                                   int amp, uint4 *__restrict__ out)
{
    for (size_t byte = (size_t)blockIdx.x * blockDim.x + threadIdx.x; byte < n_bytes; byte += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long i0 = first_sample + 8ull * byte;
        unsigned long long k = i0 < mu ? i0 : mu + (i0 - mu) % lambda;
        const unsigned b = bits[byte];
        unsigned w[4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned l = lo[k];
            const unsigned bit = (b >> j) & 1u;                                   // LSB first (:66-67)
            const int vi = (bit ^ (l & 1u)) ? -amp : amp, vq = (bit ^ (l >> 1)) ? -amp : amp;
            const unsigned pair = ((unsigned)vi & 0xFFu) | (((unsigned)vq & 0xFFu) << 8);
            if (j & 1) w[j >> 1] |= pair << 16; else w[j >> 1] = pair;
            k++; if (k == mu + lambda) k = mu;
        }
        out[byte] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace ga
