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

}  // namespace ga
