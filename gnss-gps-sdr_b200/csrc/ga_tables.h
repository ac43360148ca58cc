// ga_tables.h -- host-side construction of the constant tables the kernels read.
// Values are computed in long double and rounded once to float.
#pragma once
#include <vector>
#include <cmath>
#include "ga_common.h"

namespace ga {

// tw[t] = exp(+2*pi*i*t/n)
inline std::vector<cf> make_tw(int n)
{
    std::vector<cf> t((size_t)n);
    for (int i = 0; i < n; i++) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)i / (long double)n;
        t[(size_t)i] = mk((float)cosl(a), (float)sinl(a));
    }
    return t;
}

// ktab[s*RC + w] = exp(+2*pi*i*s*w/(N1*RC)): last factor of the pruned backward transform
template <class G> inline std::vector<cf> make_ktab()
{
    std::vector<cf> t((size_t)G::N1 * G::RC);
    for (int s = 0; s < G::N1; s++)
        for (int w = 0; w < G::RC; w++) {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)((s * w) % (G::N1 * G::RC)) / (long double)(G::N1 * G::RC);
            t[(size_t)s * G::RC + w] = mk((float)cosl(a), (float)sinl(a));
        }
    return t;
}

// ktab of output segment m (lags [m*N2, (m+1)*N2) of the N-point backward transform): the term of sub-sequence s
// carries the extra factor exp(+2*pi*i*s*m/N1), computed here in long double and rounded once
template <class G> inline std::vector<cf> make_ktab_seg(int m)
{
    std::vector<cf> t((size_t)G::N1 * G::RC);
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    for (int s = 0; s < G::N1; s++)
        for (int w = 0; w < G::RC; w++) {
            const long double a = two_pi * (long double)((s * w) % (G::N1 * G::RC)) / (long double)(G::N1 * G::RC) +
                                  two_pi * (long double)((s * m) % G::N1) / (long double)G::N1;
            t[(size_t)s * G::RC + w] = mk((float)cosl(a), (float)sinl(a));
        }
    return t;
}

// k1tab[s*N1 + n1] = exp(-2*pi*i*n1*s/N1): first (radix-N1) stage of the forward transform
template <class G> inline std::vector<cf> make_k1tab()
{
    std::vector<cf> t((size_t)G::N1 * G::N1);
    for (int s = 0; s < G::N1; s++)
        for (int n1 = 0; n1 < G::N1; n1++) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)((s * n1) % G::N1) / (long double)G::N1;
            t[(size_t)s * G::N1 + n1] = mk((float)cosl(a), (float)sinl(a));
        }
    return t;
}

}  // namespace ga
