// ga_frontend_math.h -- the per-sample arithmetic of the stream converters (ga_frontend.cuh), written
// __host__ __device__ so that tests/emu can replay it bit for bit on the CPU.
//
//  (1) 8-bit IQ -> 1 bit (proc_rtl_bin_for_gps.m:36-44 / proc_hackrf_bin_for_gps.m:13-16): the sign of
//          r = (I - mean_I)*cos - (Q - mean_Q)*sin                       (double, unfused like MATLAB)
//      as a THRESHOLD on I.  For a fixed phasor (cos, sin) and a fixed Q every rounding step of r is a monotone
//      function of I, so {I : r < 0} is a prefix (cos > 0) or a suffix (cos < 0) of the 256 byte values, or all /
//      nothing (cos = 0).  One 32-bit table entry per (Q, phase) therefore decides a sample with ONE integer
//      subtraction -- exactly, not approximately: the entry is made by evaluating the double expression itself.
//  (2) 1 bit -> int8 IQ (c/conv_1bit_bin_to_hackrf_bin.cpp:62-77): eight samples of one input byte expanded with
//      byte-lane integer arithmetic and four byte permutes instead of sixteen compare/selects.
#pragma once
#include "ga_common.h"

namespace ga {

#if defined(__CUDACC__)
typedef uint4 u32x4;
#else
struct alignas(16) u32x4 { unsigned x, y, z, w; };
#endif
GA_HD u32x4 ld_u32x4(const u32x4 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
GA_HD unsigned ld_u32(const unsigned *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// byte permute (PRMT): result byte i = byte (s >> 4i) & 7 of the eight bytes y:x
GA_HD unsigned ga_byte_perm(unsigned x, unsigned y, unsigned s)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, s);
#else
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// ---- (1) ---------------------------------------------------------------------------------------------------
// `iu`, `qu`: the sample bytes as biased unsigned values (value = u - 128): the raw byte of the uint8 format
// (proc_rtl_bin_for_gps.m:34 "y - 128"), raw ^ 0x80 for the int8 format.
GA_HD double iq8_r(int iu, int qu, double mean_i, double mean_q, double cs, double sn)
{
    const double yi = (double)(iu - 128) - mean_i, yq = (double)(qu - 128) - mean_q;
#if defined(__CUDA_ARCH__)
    return __dsub_rn(__dmul_rn(yi, cs), __dmul_rn(yq, sn));            // never contracted into an FMA
#else
    volatile double a = yi * cs, b = yq * sn;                           // (volatile: no host-side contraction either)
    return a - b;
#endif
}

// Table entry of (qu, phasor):  X | flip << 31, where bit(iu) = [iu < X] ^ flip.
//   cos >= 0: r is non-decreasing in iu, the negatives are iu < cnt            -> X = cnt,       flip = 0
//   cos <  0: r is non-increasing,       the negatives are iu >= 256 - cnt     -> X = 256 - cnt, flip = 1
// Subtracting flip << 31 toggles bit 31 of a difference, hence:  bit = (iu - entry) >> 31  (iq8_bit).
GA_HD unsigned iq8_thr_entry(int qu, double mean_i, double mean_q, double cs, double sn)
{
    int cnt = 0;
    for (int iu = 0; iu < 256; iu++) cnt += iq8_r(iu, qu, mean_i, mean_q, cs, sn) < 0.0 ? 1 : 0;
    return cs < 0.0 ? (unsigned)(256 - cnt) | 0x80000000u : (unsigned)cnt;
}

GA_HD unsigned iq8_bit(unsigned iu, unsigned entry) { return (iu - entry) >> 31; }

// Eight samples (one 128-bit load: I0 Q0 I1 Q1 ... as biased unsigned bytes) -> one output byte, LSB first (fwrite 'ubit1').
// `tab` = threshold table in BYTE addressing: row of Q at qu * pitch_b, `koff_b[j]` = 4 * phase index of sample j.
// Each decision is shifted in at the bottom (funnel shift takes bit 31 of the difference), the byte is bit-reversed at the end.
GA_HD unsigned ga_shift_in_msb(unsigned acc, unsigned v)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(v, acc, 1);
#else
    return (acc << 1) | (v >> 31);
#endif
}
GA_HD unsigned ga_brev8(unsigned acc)          // bit j of the result = bit 7-j of acc
{
#if defined(__CUDA_ARCH__)
    return __brev(acc) >> 24;
#else
    unsigned r = 0;
    for (int j = 0; j < 8; j++) r |= ((acc >> (7 - j)) & 1u) << j;
    return r;
#endif
}
// The table is addressed through a `tabref_t`: on the device a 32-bit shared-window address (so that row offset and
// base are ONE register and the lookup is a multiply-add plus ld.shared), on the host a plain address.
#if defined(__CUDA_ARCH__)
typedef unsigned tabref_t;
GA_HD unsigned tab_ld(tabref_t a) { unsigned v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#else
typedef uintptr_t tabref_t;
GA_HD unsigned tab_ld(tabref_t a) { return *reinterpret_cast<const unsigned *>(a); }
#endif
// `krow[j]` = table base + 4 * phase index of sample j;  `pitch_b` = bytes per row of Q.
GA_HD unsigned iq8_thr_byte(const unsigned w[4], unsigned pitch_b, const tabref_t krow[8])
{
    unsigned acc = 0;
    GA_UNROLL
    for (int m = 0; m < 4; m++) {            // word m = I Q I Q of samples 2m, 2m+1; byte extracts are single permutes
        const unsigned e0 = tab_ld(krow[2 * m] + ga_byte_perm(w[m], 0u, 0x4441u) * pitch_b);
        acc = ga_shift_in_msb(acc, (w[m] & 0xFFu) - e0);
        const unsigned e1 = tab_ld(krow[2 * m + 1] + (w[m] >> 24) * pitch_b);
        acc = ga_shift_in_msb(acc, ga_byte_perm(w[m], 0u, 0x4442u) - e1);
    }
    return ga_brev8(acc);
}

// What ONE thread of iq8_to_bits_thr_kernel does after the table is in shared memory.  `n_active` working threads, a
// multiple of q: thread `tid` owns the output bytes tid, tid + n_active, ... and 8 * n_active * p = 0 (mod q), so its eight
// phase indices are loop invariants.  Four 128-bit loads are issued before the first is used.
GA_HD void iq8_thr_thread(unsigned tid, unsigned n_active, const u32x4 *iq, size_t n_bytes, size_t n0, unsigned fx,
                          tabref_t tab, unsigned pitch_b, unsigned p, unsigned q, unsigned char *bits)
{
    tabref_t krow[8];
    unsigned k = (unsigned)((((n0 + 8ull * tid) % q) * p) % q);                 // p < q < 2^31
    GA_UNROLL
    for (int j = 0; j < 8; j++) { krow[j] = tab + 4u * k; k += p; if (k >= q) k -= q; }
    for (size_t byte = tid; byte < n_bytes; byte += 4 * (size_t)n_active) {
        u32x4 x[4];
        GA_UNROLL
        for (int u = 0; u < 4; u++)
            if (byte + u * (size_t)n_active < n_bytes) x[u] = ld_u32x4(iq + byte + u * (size_t)n_active);
        GA_UNROLL
        for (int u = 0; u < 4; u++) {
            if (byte + u * (size_t)n_active >= n_bytes) break;
            const unsigned w[4] = {x[u].x ^ fx, x[u].y ^ fx, x[u].z ^ fx, x[u].w ^ fx};
            bits[byte + u * (size_t)n_active] = (unsigned char)iq8_thr_byte(w, pitch_b, krow);
        }
    }
}

// row pitch (entries) of the threshold table [256][pitch]: odd, so that the rows of different Q spread over the banks
GA_HD unsigned iq8_thr_pitch(unsigned q) { return q | 1u; }

// ---- (2) ---------------------------------------------------------------------------------------------------
// Four samples: `b4` = their data bits (bit j = sample j, LSB first, :66-67), `lo4` = their LO codes, one per byte
// (lo_sin | lo_cos << 1, :30-31), `amps` = A | (-A & 0xFF) << 8.  Returns the two output words (I0 Q0 I1 Q1, I2 Q2 I3 Q3):
// I = A*Bipolar(bit ^ lo_sin), Q = A*Bipolar(bit ^ lo_cos) (:68-71).
GA_HD void conv_expand4(unsigned b4, unsigned lo4, unsigned amps, unsigned &w0, unsigned &w1)
{
    const unsigned s = ((b4 & 0xFu) * 0x00204081u) & 0x01010101u;       // bit j -> byte j (no two partial products meet)
    const unsigned d = lo4 ^ (s * 3u);                                  // per byte: (bit ^ lo_sin) | (bit ^ lo_cos) << 1
    const unsigned t = (d & 0x01010101u) | ((d & 0x02020202u) << 3);    // one selector nibble per output byte: 0 -> +A, 1 -> -A
    w0 = ga_byte_perm(amps, 0u, t & 0xFFFFu);
    w1 = ga_byte_perm(amps, 0u, t >> 16);
}

GA_HD unsigned ga_funnel_r(unsigned lo, unsigned hi, unsigned sh)      // (hi:lo) >> sh, sh in {0, 8, 16, 24}
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// What ONE thread of bits_to_iq8_v2_kernel does: input bytes `byte`, `byte + stride`, ...; the position k in the LO table
// (pre-period mu, period lambda, both <= 2^27; 16 zero bytes of slack behind the table) advances by (8 * stride) mod lambda
// per trip once the stream is past the pre-period.
GA_HD void conv_v2_thread(size_t byte, size_t stride, const unsigned char *bits, size_t n_bytes, size_t first_sample,
                          const unsigned char *lo, unsigned long long mu, unsigned long long lambda, int amp, u32x4 *out)
{
    if (byte >= n_bytes) return;
    const unsigned end = (unsigned)(mu + lambda), lam = (unsigned)lambda;
    const unsigned kstep = (unsigned)((8ull * stride) % lambda);
    const unsigned amps = ((unsigned)amp & 0xFFu) | (((unsigned)(-amp) & 0xFFu) << 8);
    unsigned long long i0 = first_sample + 8ull * byte;
    unsigned k = (unsigned)(i0 < mu ? i0 : mu + (i0 - mu) % lambda);
    for (; byte < n_bytes; byte += stride) {
        const unsigned b = bits[byte];
        unsigned l03, l47;                                                          // LO codes of samples 0..3 and 4..7, one per byte
        if (k + 8 <= end) {
            const unsigned *w = reinterpret_cast<const unsigned *>(lo + (k & ~3u));
            const unsigned w0 = ld_u32(w), w1 = ld_u32(w + 1), w2 = ld_u32(w + 2), sh = 8u * (k & 3u);
            l03 = ga_funnel_r(w0, w1, sh); l47 = ga_funnel_r(w1, w2, sh);
        } else {                                                                    // the group that wraps around the period
            l03 = l47 = 0;
            unsigned kk = k;
            GA_UNROLL
            for (int j = 0; j < 8; j++) {
                const unsigned l = lo[kk];
                if (j < 4) l03 |= l << (8 * j); else l47 |= l << (8 * (j - 4));
                kk++; if (kk == end) kk = (unsigned)mu;
            }
        }
        u32x4 o;
        conv_expand4(b, l03, amps, o.x, o.y);
        conv_expand4(b >> 4, l47, amps, o.z, o.w);
        out[byte] = o;
        // next trip: 8 * stride samples further on
        if (i0 >= mu) { k += kstep; if (k >= end) k -= lam; }
        else { const unsigned long long i1 = i0 + 8ull * stride; k = (unsigned)(i1 < mu ? i1 : mu + (i1 - mu) % lambda); }
        i0 += 8ull * stride;
    }
}

}  // namespace ga
