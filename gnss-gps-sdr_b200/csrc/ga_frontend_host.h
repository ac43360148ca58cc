// ga_frontend_host.h -- host-side preparation of the stream converters' tables (ga_frontend.cuh), plain C++ so that
// tests/emu can run it on the CPU box:
//   small_rational   shift_hz / fs as a fraction p/q in lowest terms (continued fractions): the phase of sample n is then
//                    EXACTLY 2*pi*((p*n) mod q)/q -- a table of q phasors instead of a sincos of a large argument
//   conv_lo_cycle    the float phase NCO of c/conv_1bit_bin_to_hackrf_bin.cpp:33,79-80 as a table: the recurrence is a
//                    deterministic map on floats in [0,4), hence eventually periodic (pre-period mu, period lambda)
#pragma once
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace ga {

inline bool small_rational(double x, unsigned long long &p, unsigned long long &q)
{
    // x = p/q with q <= 2^20, by continued fractions; accepted when |x - p/q| < 1e-15 * max(1, |x|)
    if (!(x >= 0) || !(x < 1e6)) return false;
    double a = x;
    unsigned long long p0 = 0, q0 = 1, p1 = 1, q1 = 0;
    for (int it = 0; it < 40; it++) {
        const double fl = floor(a);
        const unsigned long long ai = (unsigned long long)fl;
        const unsigned long long p2 = ai * p1 + p0, q2 = ai * q1 + q0;
        if (q2 > (1ull << 20)) break;
        p0 = p1; q0 = q1; p1 = p2; q1 = q2;
        if (fabs((double)p1 / (double)q1 - x) <= 1e-15 * std::max(1.0, x)) { p = p1 % q1; q = q1; return true; }
        const double fr = a - fl;
        if (fr < 1e-18) break;
        a = 1.0 / fr;
    }
    return false;
}

struct LoCycle { std::vector<unsigned char> tab; unsigned long long mu, lambda; };
// The phase NCO of :33,:79-80 is a deterministic map on floats in [0,4): eventually periodic.  Brent's cycle search
// (bounded), then the table of int(phase) -> lo_sin | lo_cos << 1 for the pre-period and one period.
inline bool conv_lo_cycle(double fc, double fs, unsigned long long n_needed, LoCycle &c)
{
    static const int lo_sin[4] = {1, 1, 0, 0}, lo_cos[4] = {1, 0, 0, 1};                  // :30-31
    const float rate = (float)(4 * fc / fs);                                               // :33
    auto step = [rate](float p) { p += rate; if (p >= 4) p -= 4; return p; };             // :79-80
    const unsigned long long LIMIT = 1ull << 27;
    unsigned long long power = 1, lam = 1;
    float t = 0, hh = step(0.0f);
    bool found = true;
    while (memcmp(&t, &hh, sizeof t) != 0) {
        if (power == lam) { t = hh; power *= 2; lam = 0; }
        hh = step(hh); lam++;
        if (lam > LIMIT) { found = false; break; }
    }
    unsigned long long mu = 0, len;
    if (found) {
        t = 0; hh = 0;
        for (unsigned long long i = 0; i < lam; i++) hh = step(hh);
        while (memcmp(&t, &hh, sizeof t) != 0) { t = step(t); hh = step(hh); mu++; if (mu > LIMIT) { found = false; break; } }
    }
    if (found) len = mu + lam;
    else { if (n_needed > LIMIT) return false; mu = 0; lam = n_needed; len = n_needed; }   // no short cycle: the whole sequence
    if (!(rate >= 0) || !(rate < 4)) return false;                                         // int(phase) would leave the 4-entry tables
    c.tab.resize(len);
    float ph = 0;
    for (unsigned long long i = 0; i < len; i++) {
        const int k = (int)ph;
        c.tab[i] = (unsigned char)(lo_sin[k & 3] | (lo_cos[k & 3] << 1));
        ph = step(ph);
    }
    c.mu = mu; c.lambda = lam;
    return true;
}

}  // namespace ga
