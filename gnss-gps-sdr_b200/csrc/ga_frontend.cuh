// ga_frontend.cuh -- 8-bit IQ front-end on the GPU: what the reference does in MATLAB before gps_test
// can run on an SDR capture (proc_rtl_bin_for_gps.m:31-47 for rtl-sdr uint8, proc_hackrf_bin_for_gps.m:7-19
// for HackRF int8):   y = I + jQ (uint8: minus 128) ;  y = y - mean(y)  (mean over the WHOLE capture) ;
// r[n] = real( y[n] * exp(j*2*pi*fc*n/fs) ) ;  bit = (1 - sign(r))/2 ;  fwrite(...,'ubit1') (LSB first).
// Double precision like MATLAB; the sums for the mean are exact integers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ga_frontend_math.h"

namespace ga {

// pass 1: exact integer sums of I and Q (format 0: uint8 offset 128, format 1: int8).  The stream is read with 128-bit
// loads (iq 16-byte aligned: eight samples per load) and the byte lanes are summed with dp4a -- mask 0x00010001 picks the
// two I bytes of a 32-bit word, 0x01000100 the two Q bytes; the 32-bit partial sums are flushed to 64 bits every 1024
// loads (at most 1024 * 8 * 255 < 2^31).  The offset 128 of the unsigned format comes off once at the end.
__global__ void iq8_sum_kernel(const unsigned char *__restrict__ iq, size_t n_samples, int format,
                               long long *__restrict__ sums /* [2] */)
{
    long long si = 0, sq = 0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(iq) & 15) == 0) {
        const size_t n16 = n_samples / 8;                                 // whole groups of 8 samples = 16 bytes
        const uint4 *v = reinterpret_cast<const uint4 *>(iq);
        int ai = 0, aq = 0, since = 0;
        for (size_t g = tid; g < n16; g += nthr) {
            const uint4 x = __ldg(v + g);
            if (format == 0) {
                ai = (int)__dp4a(x.x, 0x00010001u, (unsigned)ai); aq = (int)__dp4a(x.x, 0x01000100u, (unsigned)aq);
                ai = (int)__dp4a(x.y, 0x00010001u, (unsigned)ai); aq = (int)__dp4a(x.y, 0x01000100u, (unsigned)aq);
                ai = (int)__dp4a(x.z, 0x00010001u, (unsigned)ai); aq = (int)__dp4a(x.z, 0x01000100u, (unsigned)aq);
                ai = (int)__dp4a(x.w, 0x00010001u, (unsigned)ai); aq = (int)__dp4a(x.w, 0x01000100u, (unsigned)aq);
                ai -= 8 * 128; aq -= 8 * 128;
            } else {
                ai = __dp4a((int)x.x, 0x00010001, ai); aq = __dp4a((int)x.x, 0x01000100, aq);
                ai = __dp4a((int)x.y, 0x00010001, ai); aq = __dp4a((int)x.y, 0x01000100, aq);
                ai = __dp4a((int)x.z, 0x00010001, ai); aq = __dp4a((int)x.z, 0x01000100, aq);
                ai = __dp4a((int)x.w, 0x00010001, ai); aq = __dp4a((int)x.w, 0x01000100, aq);
            }
            if (++since == 1024) { si += ai; sq += aq; ai = aq = since = 0; }
        }
        si += ai; sq += aq;
        for (size_t n = n16 * 8 + tid; n < n_samples; n += nthr) {         // the last n_samples % 8 samples
            const int a = iq[2 * n], b = iq[2 * n + 1];
            si += format == 0 ? a - 128 : (int)(signed char)a;
            sq += format == 0 ? b - 128 : (int)(signed char)b;
        }
    } else {
        for (size_t n = tid; n < n_samples; n += nthr) {
            const int a = iq[2 * n], b = iq[2 * n + 1];
            si += format == 0 ? a - 128 : (int)(signed char)a;
            sq += format == 0 ? b - 128 : (int)(signed char)b;
        }
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
        const int fx = format == 0 ? 0 : 0x80;                               // int8 -> biased unsigned
        double sn, cs;
        sincos((w * (double)(n0 + n)) * inv_fs, &sn, &cs);
        const double r = iq8_r(iq[2 * n] ^ fx, iq[2 * n + 1] ^ fx, mean_i, mean_q, cs, sn);
        out |= (r < 0.0 ? 1u : 0u) << k;                                     // (1-sign(r))/2, LSB first
    }
    bits[byte] = (unsigned char)out;
}

// pass 2 with an exactly periodic phase: when fc/fs = p/q in lowest terms with a small q the phasor of sample n is
// table[(p*n) mod q] (double cos/sin made on the host in long double) -- no double-precision sincos of a large
// argument per sample.  The mathematically exact phase; the MATLAB expression differs from it by its own rounding of
// 2*pi*fc*n/fs (~1e-10 rad at n ~ 1e8), which matters only where |r| is that small.
// One thread per FOUR output bytes: 32 samples = four 128-bit loads in, one 32-bit store out (d_bits is 4-byte aligned:
// it comes from cudaMalloc or from the caller's device buffer at a multiple of 32 samples); the phasor table sits in
// shared memory when it is small (the usual front-end ratios give q of a few hundred).
__global__ void iq8_to_bits_table_kernel(const unsigned char *__restrict__ iq, size_t n_samples, size_t n0, int format,
                                         double mean_i, double mean_q, const double2 *__restrict__ table,
                                         unsigned long long p, unsigned long long q, unsigned char *__restrict__ bits)
{
    extern __shared__ double2 tab_s[];
    const bool in_smem = q <= 2048;
    if (in_smem) {
        for (unsigned k = threadIdx.x; k < (unsigned)q; k += blockDim.x) tab_s[k] = table[k];
        __syncthreads();
    }
    const double2 *tab = in_smem ? tab_s : table;
    const size_t word = (size_t)blockIdx.x * blockDim.x + threadIdx.x;    // output word = samples 32*word .. 32*word+31
    if (word * 32 >= n_samples) return;
    unsigned long long k = (((n0 + word * 32) % q) * p) % q;               // p, q < 2^31: no overflow
    const bool whole = word * 32 + 32 <= n_samples && (reinterpret_cast<uintptr_t>(bits) & 3) == 0;
    unsigned out = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {                                          // output byte g of the word
        const size_t byte = word * 4 + g;
        if (byte * 8 >= n_samples) break;
        unsigned char raw[16];
        if (byte * 8 + 8 <= n_samples) {
            *reinterpret_cast<uint4 *>(raw) = __ldg(reinterpret_cast<const uint4 *>(iq + 16 * byte));
        } else {
            for (int i = 0; i < 16; i++) raw[i] = (byte * 16 + i < 2 * n_samples) ? iq[16 * byte + i] : 0;
        }
        unsigned ob = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (byte * 8 + j >= n_samples) break;
            const int fx = format == 0 ? 0 : 0x80;
            const double2 cs = tab[k];
            const double r = iq8_r(raw[2 * j] ^ fx, raw[2 * j + 1] ^ fx, mean_i, mean_q, cs.x, cs.y);
            ob |= (r < 0.0 ? 1u : 0u) << j;
            k += p; if (k >= q) k -= q;
        }
        if (whole) out |= ob << (8 * g);
        else bits[byte] = (unsigned char)ob;
    }
    if (whole) reinterpret_cast<unsigned *>(bits)[word] = out;
}

// pass 2 as a table of thresholds (ga_frontend_math.h): for the usual front-end ratios (fc/fs = p/q, q <= IQ8_THR_MAX_Q)
// the sign of r for a sample (I, Q, phase k) is  (I - thr[Q][k]) >> 31  -- an exact restatement of the double expression,
// because the table is MADE by evaluating it for all 256 x 256 x q cases.  No floating point in the hot loop.
//   iq8_thr_build_kernel   q blocks x 256 threads: entry (Q = thread, k = block) from the integer sums of pass 1 (the mean
//                          never visits the host: the three kernels of a conversion run back to back on the stream)
//   iq8_to_bits_thr_kernel persistent CTAs with the table in shared memory; one thread per output byte = 8 samples = one
//                          128-bit load; the number of working threads is a multiple of q, so the eight phase indices of a
//                          thread are the same in every trip of its grid-stride loop (kept as eight row offsets in registers);
//                          per sample: two byte extracts, one multiply-add (row address), one shared load, one subtraction,
//                          one funnel shift.  Four loads in flight per thread.
#define IQ8_THR_MAX_Q 227                       // 256 rows x (q | 1) entries x 4 B <= 227 KB of shared memory
#define IQ8_THR_THREADS 1024
__global__ void iq8_thr_build_kernel(const long long *__restrict__ sums, size_t n_total, const double2 *__restrict__ tab,
                                     unsigned pitch, unsigned *__restrict__ thr)
{
    const unsigned k = blockIdx.x, qu = threadIdx.x;
    const double mean_i = (double)sums[0] / (double)n_total, mean_q = (double)sums[1] / (double)n_total;
    const double2 cs = tab[k];
    thr[qu * pitch + k] = iq8_thr_entry((int)qu, mean_i, mean_q, cs.x, cs.y);
}

__global__ void __launch_bounds__(IQ8_THR_THREADS, 1)
iq8_to_bits_thr_kernel(const uint4 *__restrict__ iq, size_t n_samples, size_t n0, unsigned fx /* 0 or 0x80808080 */,
                       const unsigned *__restrict__ thr, unsigned p, unsigned q, unsigned pitch, unsigned n_active,
                       const long long *__restrict__ sums, size_t n_total, const double2 *__restrict__ tab,
                       unsigned char *__restrict__ bits)
{
    extern __shared__ uint4 thr_s4[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(thr);
        for (unsigned i = threadIdx.x; i < 64u * pitch; i += blockDim.x) thr_s4[i] = __ldg(src + i);     // 256 * pitch * 4 B
    }
    __syncthreads();
    const tabref_t tab_s = (tabref_t)__cvta_generic_to_shared(thr_s4);
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, pitch_b = 4u * pitch;
    const size_t n_bytes = n_samples / 8;
    if (tid < n_active) iq8_thr_thread(tid, n_active, iq, n_bytes, n0, fx, tab_s, pitch_b, p, q, bits);
    if (tid == 0 && (n_samples & 7)) {                  // the last n_samples % 8 samples: the double expression itself
        const double mean_i = (double)sums[0] / (double)n_total, mean_q = (double)sums[1] / (double)n_total;
        const unsigned char *raw = reinterpret_cast<const unsigned char *>(iq);
        unsigned ob = 0;
        for (size_t n = n_bytes * 8; n < n_samples; n++) {
            const double2 cs = tab[(((n0 + n) % q) * p) % q];
            const int f = (int)(fx & 0x80u);
            ob |= (iq8_r(raw[2 * n] ^ f, raw[2 * n + 1] ^ f, mean_i, mean_q, cs.x, cs.y) < 0.0 ? 1u : 0u) << (n & 7);
        }
        bits[n_bytes] = (unsigned char)ob;
    }
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
// (The LO table is allocated with 16 spare bytes: the eight codes of a thread are cut out of three aligned 32-bit loads
// with two funnel shifts instead of eight byte loads; the group that wraps around the period takes the byte path.)
__global__ void bits_to_iq8_kernel(const unsigned char *__restrict__ bits, size_t n_bytes, size_t first_sample,
                                   const unsigned char *__restrict__ lo, unsigned long long mu, unsigned long long lambda,
                                   int amp, uint4 *__restrict__ out)
{
    for (size_t byte = (size_t)blockIdx.x * blockDim.x + threadIdx.x; byte < n_bytes; byte += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long i0 = first_sample + 8ull * byte;
        unsigned long long k = i0 < mu ? i0 : mu + (i0 - mu) % lambda;
        const unsigned b = bits[byte];
        unsigned l03, l47;                                                        // LO codes of samples 0..3 and 4..7, one per byte
        if (k + 8 <= mu + lambda) {
            const unsigned *w = reinterpret_cast<const unsigned *>(lo + (k & ~3ull));
            const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), sh = 8u * (unsigned)(k & 3);
            l03 = __funnelshift_r(w0, w1, sh); l47 = __funnelshift_r(w1, w2, sh);
        } else {
            l03 = l47 = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned l = lo[k];
                if (j < 4) l03 |= l << (8 * j); else l47 |= l << (8 * (j - 4));
                k++; if (k == mu + lambda) k = mu;
            }
        }
        unsigned w[4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned l = ((j < 4 ? l03 : l47) >> (8 * (j & 3))) & 3u;
            const unsigned bit = (b >> j) & 1u;                                   // LSB first (:66-67)
            const int vi = (bit ^ (l & 1u)) ? -amp : amp, vq = (bit ^ (l >> 1)) ? -amp : amp;
            const unsigned pair = ((unsigned)vi & 0xFFu) | (((unsigned)vq & 0xFFu) << 8);
            if (j & 1) w[j >> 1] |= pair << 16; else w[j >> 1] = pair;
        }
        out[byte] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// The same conversion with the per-thread work cut from ~200 to ~40 instructions (the byte-per-sample kernel above is
// instruction bound: a 64-bit modulo per 16 output bytes and sixteen compare/selects): the position in the LO cycle
// advances incrementally along the grid-stride loop (32-bit; one 64-bit modulo per THREAD), and the eight samples of a
// byte are expanded with byte-lane arithmetic and four byte permutes (conv_expand4, ga_frontend_math.h).
__global__ void bits_to_iq8_v2_kernel(const unsigned char *__restrict__ bits, size_t n_bytes, size_t first_sample,
                                      const unsigned char *__restrict__ lo, unsigned long long mu, unsigned long long lambda,
                                      int amp, uint4 *__restrict__ out)
{
    conv_v2_thread((size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x, bits, n_bytes, first_sample, lo, mu, lambda, amp, out);
}

}  // namespace ga
