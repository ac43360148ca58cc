// ga_radix.h -- register-resident radix-R DFT butterflies (R = 2,4,5 primitives,
// composites by Good-Thomas prime-factor mapping (no internal twiddles) or
// Cooley-Tukey with compile-time twiddles).  DIR = +1: exp(+2*pi*i*nk/R)
// (backward, what Correlate()'s rev_plan computes, c/search_offline.cpp:79,187);
// DIR = -1: forward (fwd_plan, :78,:105,:161).
//
// All index maps are compile-time; with full unrolling every x[] element is a
// register and the maps cost nothing.
#pragma once
#include "ga_common.h"
#include <type_traits>

namespace ga {

// ---- compile-time trigonometry (so composite twiddles are immediates) --------
constexpr double cx_pi = 3.14159265358979323846264338327950288;
constexpr double cx_sin_series(double x)
{   // |x| <= pi/2 after reduction; 15 terms are far below double epsilon
    double term = x, sum = x, x2 = x * x;
    for (int k = 1; k < 16; k++) { term *= -x2 / ((2.0 * k) * (2.0 * k + 1.0)); sum += term; }
    return sum;
}
// sin(2*pi*p/q), cos(2*pi*p/q) for integers, exact symmetries first
constexpr double cx_sin2pi(int p, int q)
{
    p %= q; if (p < 0) p += q;
    if (2 * p > q) return -cx_sin2pi(q - p, q);            // second half-turn
    if (4 * p > q) return cx_sin2pi(q - 2 * p, 2 * q);     // sin(pi - x) = sin x
    return cx_sin_series(2.0 * cx_pi * p / q);
}
constexpr double cx_cos2pi(int p, int q) { return cx_sin2pi(4 * p + q, 4 * q); }  // cos x = sin(x + pi/2)

constexpr int cx_inv_mod(int a, int m)
{
    a %= m;
    for (int x = 1; x < m; x++) if ((a * x) % m == 1) return x;
    return 1;
}

// compile-time loop: the body receives std::integral_constant<int, I>, so that
// twiddles can be formed as constexpr immediates
template <int I, int NITER, class F> GA_HD void static_for(F &&f)
{
    if constexpr (I < NITER) { f(std::integral_constant<int, I>{}); static_for<I + 1, NITER>(f); }
}

// ---- primitives --------------------------------------------------------------
#ifndef GA_R4_SCALAR
#define GA_R4_SCALAR 0
#endif
#ifndef GA_R5_SCALAR
#define GA_R5_SCALAR 0
#endif
// Scalar FADD pairs instead of FADD2 (ga_common.h: same bits, FMA-lite pipe, one more issue slot), per call site, measured on
// a B200 with byte-identical peak records (tools/gpu_r2y.sh, gpu_r2z.sh, gpu_r2zz.sh; profiles/r02_scalar_adds.txt):
//   prime radices (GRID kernels, FMA pipe 71 % busy): input and output sums scalar  +0.9 ... +2.2 %  -> default
//   radix 4 (inside 16 / 20 / 24): REF +0.3 %, GRID -1 %;   radix 5 (inside 20 / 25): REF -0.2 ... -0.4 %   -> packed
#ifndef GA_RP_SCALAR        // prime radices: bit 0 = the input sums / differences scalar, bit 1 = the output sums
#define GA_RP_SCALAR 3
#endif
template <int R, int DIR> struct Radix;

template <int DIR> struct Radix<1, DIR> { static GA_HD void run(cf (&)[1]) {} };

template <int DIR> struct Radix<2, DIR> {
    static GA_HD void run(cf (&x)[2])
    {
        cf a = x[0], b = x[1];
        x[0] = cadd(a, b); x[1] = csub(a, b);
    }
};

template <int DIR> struct Radix<4, DIR> {
    static GA_HD void run(cf (&x)[4])
    {
#if GA_R4_SCALAR        // scalar adds (FMA-lite pipe, beside the packed instructions) instead of FADD2
        const cf s0 = cadd_s(x[0], x[2]), d0 = csub_s(x[0], x[2]);
        const cf s1 = cadd_s(x[1], x[3]), d1 = csub_s(x[1], x[3]);
        x[0] = cadd_s(s0, s1); x[1] = cadd_i_s<DIR>(d0, d1);
        x[2] = csub_s(s0, s1); x[3] = csub_i_s<DIR>(d0, d1);
#else
        const cf s0 = cadd(x[0], x[2]), d0 = csub(x[0], x[2]);
        const cf s1 = cadd(x[1], x[3]), d1 = csub(x[1], x[3]);
        x[0] = cadd(s0, s1); x[1] = cadd_i<DIR>(d0, d1);     // d0 + DIR*i*d1
        x[2] = csub(s0, s1); x[3] = csub_i<DIR>(d0, d1);
#endif
    }
};

template <int DIR> struct Radix<5, DIR> {
    static GA_HD void run(cf (&x)[5])
    {
        constexpr float c1 = (float)cx_cos2pi(1, 5), c2 = (float)cx_cos2pi(2, 5);
        constexpr float s1 = (float)cx_sin2pi(1, 5), s2 = (float)cx_sin2pi(2, 5);
#if GA_R5_SCALAR == 1  // the input sums / differences scalar
        const cf t1 = cadd_s(x[1], x[4]), t3 = csub_s(x[1], x[4]);
        const cf t2 = cadd_s(x[2], x[3]), t4 = csub_s(x[2], x[3]);
#else
        const cf t1 = cadd(x[1], x[4]), t3 = csub(x[1], x[4]);
        const cf t2 = cadd(x[2], x[3]), t4 = csub(x[2], x[3]);
#endif
        const cf b1 = caxpy(caxpy(x[0], t1, c1), t2, c2);
        const cf b2 = caxpy(caxpy(x[0], t1, c2), t2, c1);
        const cf d1 = caxpy(cscale(t3, s1), t4, s2);         // X1 = b1 + DIR*i*d1
        const cf d2 = caxpy(cscale(t3, s2), t4, -s1);
#if GA_R5_SCALAR == 2  // the output sums scalar
        x[0] = cadd_s(x[0], cadd_s(t1, t2));
        x[1] = cadd_i_s<DIR>(b1, d1); x[4] = csub_i_s<DIR>(b1, d1);
        x[2] = cadd_i_s<DIR>(b2, d2); x[3] = csub_i_s<DIR>(b2, d2);
#else
        x[0] = cadd(x[0], cadd(t1, t2));
        x[1] = cadd_i<DIR>(b1, d1); x[4] = csub_i<DIR>(b1, d1);
        x[2] = cadd_i<DIR>(b2, d2); x[3] = csub_i<DIR>(b2, d2);
#endif
    }
};

// ---- Good-Thomas: R = R1*R2, gcd(R1,R2) = 1, no twiddles ----------------------
template <int R1, int R2, int DIR> struct PFA {
    static constexpr int R = R1 * R2;
    static GA_HD void run(cf (&x)[R1 * R2])
    {
        constexpr int A = R2 * cx_inv_mod(R2, R1);   // k = (A*k1 + B*k2) mod R
        constexpr int B = R1 * cx_inv_mod(R1, R2);
        cf t[R2][R1];
        GA_UNROLL
        for (int n2 = 0; n2 < R2; n2++) {
            cf col[R1];
            GA_UNROLL
            for (int n1 = 0; n1 < R1; n1++) col[n1] = x[(R2 * n1 + R1 * n2) % R];
            Radix<R1, DIR>::run(col);
            GA_UNROLL
            for (int k1 = 0; k1 < R1; k1++) t[n2][k1] = col[k1];
        }
        GA_UNROLL
        for (int k1 = 0; k1 < R1; k1++) {
            cf row[R2];
            GA_UNROLL
            for (int n2 = 0; n2 < R2; n2++) row[n2] = t[n2][k1];
            Radix<R2, DIR>::run(row);
            GA_UNROLL
            for (int k2 = 0; k2 < R2; k2++) x[(A * k1 + B * k2) % R] = row[k2];
        }
    }
};

// ---- Cooley-Tukey: R = R1*R2 with compile-time twiddles -----------------------

template <int R1, int R2, int DIR> struct CT {
    static constexpr int R = R1 * R2;
    static GA_HD void run(cf (&x)[R1 * R2])
    {
        cf t[R2][R1];
        static_for<0, R2>([&](auto n2c) {
            constexpr int n2 = decltype(n2c)::value;
            cf col[R1];
            GA_UNROLL
            for (int n1 = 0; n1 < R1; n1++) col[n1] = x[R2 * n1 + n2];
            Radix<R1, DIR>::run(col);
            static_for<0, R1>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                if constexpr (n2 * k1 != 0) {
                    constexpr float wr = (float)cx_cos2pi(n2 * k1, R);
                    constexpr float wi = (float)(DIR * cx_sin2pi(n2 * k1, R));
                    col[k1] = cmul(col[k1], mk(wr, wi));
                }
                t[n2][k1] = col[k1];
            });
        });
        GA_UNROLL
        for (int k1 = 0; k1 < R1; k1++) {
            cf row[R2];
            GA_UNROLL
            for (int n2 = 0; n2 < R2; n2++) row[n2] = t[n2][k1];
            Radix<R2, DIR>::run(row);
            GA_UNROLL
            for (int k2 = 0; k2 < R2; k2++) x[k1 + R1 * k2] = row[k2];
        }
    }
};

// ---- odd prime radix: direct evaluation with the conjugate-pair symmetry -------
// X_j = x0 + sum_{k=1..H} [ cos(2 pi jk/P) (x_k + x_{P-k}) + DIR*i*sin(2 pi jk/P) (x_k - x_{P-k}) ],  H = (P-1)/2,
// X_{P-j} = the same with the sign of the sine part flipped.  All cosines / sines are compile-time
// immediates (on sm_100a FFMA2 takes a broadcast 32-bit immediate, so a term costs one instruction
// per pair and no register); 2*H*H + O(P) packed operations per butterfly.  Used for the radices the
// 1 ms GPS block lengths need: 5456 = 16*11*31, 8184 = 24*11*31, 2800 = 16*25*7.
template <int P, int DIR> struct RadixPrime {
    static constexpr int H = (P - 1) / 2;
    // streaming form: emit(std::integral_constant<int, j>, X_j) is called as soon as an output is complete
    // (order 0, 1, P-1, 2, P-2, ...), so a consumer that reduces or stores the outputs never holds all P
    // of them next to the 2*H sums and differences.
    template <class Emit> static GA_HD void run_emit(const cf (&x)[P], Emit &&emit)
    {
        cf s[H], d[H];
        GA_UNROLL
        for (int k = 1; k <= H; k++) {
            if (GA_RP_SCALAR & 1) { s[k - 1] = cadd_s(x[k], x[P - k]); d[k - 1] = csub_s(x[k], x[P - k]); }
            else { s[k - 1] = cadd(x[k], x[P - k]); d[k - 1] = csub(x[k], x[P - k]); }
        }
        const cf x0 = x[0];
        {   // X_0: pairwise tree
            cf t[H];
            GA_UNROLL
            for (int k = 0; k < H; k++) t[k] = s[k];
            GA_UNROLL
            for (int n = H; n > 1; n = (n + 1) / 2) {
                GA_UNROLL
                for (int k = 0; k < n / 2; k++) t[k] = cadd(t[k], t[n - 1 - k]);
            }
            emit(std::integral_constant<int, 0>{}, cadd(x0, t[0]));
        }
        static_for<1, H + 1>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            cf a = x0, b = mk(0.0f, 0.0f);
            static_for<1, H + 1>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr float c = (float)cx_cos2pi(j * k, P), sn = (float)cx_sin2pi(j * k, P);
                a = caxpy(a, s[k - 1], c);
                b = (k == 1) ? cscale(d[0], sn) : caxpy(b, d[k - 1], sn);
            });
            if (GA_RP_SCALAR & 2) {
                emit(std::integral_constant<int, j>{}, cadd_i_s<DIR>(a, b));
                emit(std::integral_constant<int, P - j>{}, csub_i_s<DIR>(a, b));
            } else {
                emit(std::integral_constant<int, j>{}, cadd_i<DIR>(a, b));      // a + DIR*i*b
                emit(std::integral_constant<int, P - j>{}, csub_i<DIR>(a, b));
            }
        });
    }
    static GA_HD void run(cf (&x)[P])
    {
        cf y[P];
        run_emit(x, [&](auto ic, cf v) { y[decltype(ic)::value] = v; });
        GA_UNROLL
        for (int k = 0; k < P; k++) x[k] = y[k];
    }
};
template <int DIR> struct Radix<3, DIR>  { static GA_HD void run(cf (&x)[3])  { RadixPrime<3, DIR>::run(x); } };
template <int DIR> struct Radix<7, DIR>  { static GA_HD void run(cf (&x)[7])  { RadixPrime<7, DIR>::run(x); } };
template <int DIR> struct Radix<11, DIR> { static GA_HD void run(cf (&x)[11]) { RadixPrime<11, DIR>::run(x); } };
template <int DIR> struct Radix<31, DIR> { static GA_HD void run(cf (&x)[31]) { RadixPrime<31, DIR>::run(x); } };

template <int DIR> struct Radix<8, DIR>  { static GA_HD void run(cf (&x)[8])  { CT<2, 4, DIR>::run(x); } };
template <int DIR> struct Radix<10, DIR> { static GA_HD void run(cf (&x)[10]) { PFA<2, 5, DIR>::run(x); } };
template <int DIR> struct Radix<16, DIR> { static GA_HD void run(cf (&x)[16]) { CT<4, 4, DIR>::run(x); } };
template <int DIR> struct Radix<20, DIR> { static GA_HD void run(cf (&x)[20]) { PFA<4, 5, DIR>::run(x); } };
template <int DIR> struct Radix<25, DIR> { static GA_HD void run(cf (&x)[25]) { CT<5, 5, DIR>::run(x); } };
template <int DIR> struct Radix<24, DIR> { static GA_HD void run(cf (&x)[24]) { PFA<3, 8, DIR>::run(x); } };

// radix-R butterfly with streamed outputs: emit(std::integral_constant<int, k>, X_k).  Prime radices
// stream for real (see RadixPrime); the others run in place and then emit in index order.
constexpr bool cx_streams(int r) { return r == 3 || r == 7 || r == 11 || r == 31; }
template <int R, int DIR, class Emit> GA_HD void radix_emit(cf (&p)[R], Emit &&emit)
{
    if constexpr (cx_streams(R)) RadixPrime<R, DIR>::run_emit(p, emit);
    else {
        Radix<R, DIR>::run(p);
        static_for<0, R>([&](auto ic) { emit(ic, p[decltype(ic)::value]); });
    }
}

// powers of a unit complex number: w[k] = base^k, k < R, by a balanced product
// tree (depth ~log2 R, so rounding stays at a few ulp).  w[0] is 1.
template <int R> GA_HD void unit_powers(cf base, cf (&w)[R])
{
    w[0] = mk(1.0f, 0.0f);
    if (R > 1) w[1] = base;
    GA_UNROLL
    for (int k = 2; k < R; k++) {
        const int h = k / 2;
        w[k] = (k & 1) ? cmul(w[h], w[h + 1]) : csqr(w[h]);
    }
}

}  // namespace ga
