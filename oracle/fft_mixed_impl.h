/*
 * oracle/fft_mixed_impl.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Mixed-radix Stockham autosort complex FFT, included twice (REAL=float /
 * REAL=double) by fft_mixed.c.  This is our own CPU FFT: it stands in for the
 * FFTW codelets the reference links against (c/search_offline.cpp:78-79,105,
 * 161,187 -- libfftw3f is not vendored and not installed here), and it is the
 * FFT of the C restatement in gpsacq_oracle.c.  Unnormalised, sign = -1
 * forward / +1 backward, any length (radices 4,2,5,3 fast paths, generic
 * O(r^2) butterfly for other primes such as 11 and 31).
 *
 * Required macros: REAL, SUF(name).
 */

/* SUF(cpx) comes from fft_mixed.h */
struct SUF(fft_plan) {
    int n, sign, npass;
    int radix[32];
    SUF(cpx) *tw;      /* tw[t] = exp(sign*2*pi*i*t/n), t < n      */
    SUF(cpx) *buf0;    /* ping                                       */
    SUF(cpx) *buf1;    /* pong                                       */
};

static void SUF(fft_factor)(SUF(fft_plan) *p)
{
    int n = p->n, np = 0;
    while (n % 4 == 0) { p->radix[np++] = 4; n /= 4; }
    while (n % 2 == 0) { p->radix[np++] = 2; n /= 2; }
    while (n % 5 == 0) { p->radix[np++] = 5; n /= 5; }
    while (n % 3 == 0) { p->radix[np++] = 3; n /= 3; }
    for (int f = 7; n > 1; f += 2)
        while (n % f == 0) { p->radix[np++] = f; n /= f; }
    p->npass = np;
}

SUF(fft_plan) *SUF(fft_plan_create)(int n, int sign)
{
    SUF(fft_plan) *p = (SUF(fft_plan) *)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->n = n;
    p->sign = sign < 0 ? -1 : +1;
    SUF(fft_factor)(p);
    p->tw   = (SUF(cpx) *)malloc(sizeof(SUF(cpx)) * (size_t)n);
    p->buf0 = (SUF(cpx) *)malloc(sizeof(SUF(cpx)) * (size_t)n);
    p->buf1 = (SUF(cpx) *)malloc(sizeof(SUF(cpx)) * (size_t)n);
    if (!p->tw || !p->buf0 || !p->buf1) return NULL;
    for (int t = 0; t < n; t++) {
        /* exact octant reduction is not needed: long double keeps the table
         * correctly rounded for REAL=double as well */
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)t / (long double)n;
        p->tw[t].re = (REAL)cosl(a);
        p->tw[t].im = (REAL)(p->sign * sinl(a));
    }
    return p;
}

void SUF(fft_plan_destroy)(SUF(fft_plan) *p)
{
    if (!p) return;
    free(p->tw); free(p->buf0); free(p->buf1); free(p);
}

/* one Stockham pass: in -> out, radix r, Ns = product of earlier radices */
static void SUF(fft_pass)(const SUF(fft_plan) *p, const SUF(cpx) *in, SUF(cpx) *out, int r, int Ns)
{
    const int n = p->n, nb = n / r;          /* butterflies in this pass      */
    const int tstep = n / (Ns * r);          /* twiddle index step            */
    const REAL sg = (REAL)p->sign;
    const SUF(cpx) *tw = p->tw;

    if (r == 4) {
        for (int j0 = 0; j0 < nb; j0 += Ns) {
            for (int k = 0; k < Ns; k++) {
                const int j = j0 + k;
                SUF(cpx) a = in[j], b = in[j + nb], c = in[j + 2 * nb], d = in[j + 3 * nb];
                if (k) {
                    SUF(cpx) w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep], t;
                    t.re = b.re * w1.re - b.im * w1.im; t.im = b.re * w1.im + b.im * w1.re; b = t;
                    t.re = c.re * w2.re - c.im * w2.im; t.im = c.re * w2.im + c.im * w2.re; c = t;
                    t.re = d.re * w3.re - d.im * w3.im; t.im = d.re * w3.im + d.im * w3.re; d = t;
                }
                REAL s0r = a.re + c.re, s0i = a.im + c.im, d0r = a.re - c.re, d0i = a.im - c.im;
                REAL s1r = b.re + d.re, s1i = b.im + d.im, d1r = b.re - d.re, d1i = b.im - d.im;
                /* (sign*i)*(d1) = (-sg*d1i, sg*d1r) */
                REAL rr = -sg * d1i, ri = sg * d1r;
                SUF(cpx) *o = out + (size_t)j0 * 4 + k;
                o[0].re      = s0r + s1r; o[0].im      = s0i + s1i;
                o[Ns].re     = d0r + rr;  o[Ns].im     = d0i + ri;
                o[2 * Ns].re = s0r - s1r; o[2 * Ns].im = s0i - s1i;
                o[3 * Ns].re = d0r - rr;  o[3 * Ns].im = d0i - ri;
            }
        }
        return;
    }
    if (r == 2) {
        for (int j0 = 0; j0 < nb; j0 += Ns) {
            for (int k = 0; k < Ns; k++) {
                const int j = j0 + k;
                SUF(cpx) a = in[j], b = in[j + nb];
                if (k) {
                    SUF(cpx) w1 = tw[k * tstep], t;
                    t.re = b.re * w1.re - b.im * w1.im; t.im = b.re * w1.im + b.im * w1.re; b = t;
                }
                SUF(cpx) *o = out + (size_t)j0 * 2 + k;
                o[0].re  = a.re + b.re; o[0].im  = a.im + b.im;
                o[Ns].re = a.re - b.re; o[Ns].im = a.im - b.im;
            }
        }
        return;
    }
    if (r == 5) {
        const REAL c1 = (REAL)0.30901699437494742410L, c2 = (REAL)-0.80901699437494742410L;
        const REAL s1 = (REAL)0.95105651629515357212L * sg, s2 = (REAL)0.58778525229247312917L * sg;
        for (int j0 = 0; j0 < nb; j0 += Ns) {
            for (int k = 0; k < Ns; k++) {
                const int j = j0 + k;
                SUF(cpx) x0 = in[j], x1 = in[j + nb], x2 = in[j + 2 * nb], x3 = in[j + 3 * nb], x4 = in[j + 4 * nb];
                if (k) {
                    SUF(cpx) w, t;
                    w = tw[k * tstep];     t.re = x1.re * w.re - x1.im * w.im; t.im = x1.re * w.im + x1.im * w.re; x1 = t;
                    w = tw[2 * k * tstep]; t.re = x2.re * w.re - x2.im * w.im; t.im = x2.re * w.im + x2.im * w.re; x2 = t;
                    w = tw[3 * k * tstep]; t.re = x3.re * w.re - x3.im * w.im; t.im = x3.re * w.im + x3.im * w.re; x3 = t;
                    w = tw[4 * k * tstep]; t.re = x4.re * w.re - x4.im * w.im; t.im = x4.re * w.im + x4.im * w.re; x4 = t;
                }
                REAL t1r = x1.re + x4.re, t1i = x1.im + x4.im, t3r = x1.re - x4.re, t3i = x1.im - x4.im;
                REAL t2r = x2.re + x3.re, t2i = x2.im + x3.im, t4r = x2.re - x3.re, t4i = x2.im - x3.im;
                REAL b1r = x0.re + c1 * t1r + c2 * t2r, b1i = x0.im + c1 * t1i + c2 * t2i;
                REAL b2r = x0.re + c2 * t1r + c1 * t2r, b2i = x0.im + c2 * t1i + c1 * t2i;
                REAL d1r = s1 * t3r + s2 * t4r, d1i = s1 * t3i + s2 * t4i;
                REAL d2r = s2 * t3r - s1 * t4r, d2i = s2 * t3i - s1 * t4i;
                SUF(cpx) *o = out + (size_t)j0 * 5 + k;
                o[0].re      = x0.re + t1r + t2r; o[0].im      = x0.im + t1i + t2i;
                o[Ns].re     = b1r - d1i;         o[Ns].im     = b1i + d1r;   /* b1 + i*d1 */
                o[4 * Ns].re = b1r + d1i;         o[4 * Ns].im = b1i - d1r;
                o[2 * Ns].re = b2r - d2i;         o[2 * Ns].im = b2i + d2r;
                o[3 * Ns].re = b2r + d2i;         o[3 * Ns].im = b2i - d2r;
            }
        }
        return;
    }
    /* generic radix (3, 7, 11, 31, ...): direct r x r DFT on the twiddled inputs */
    {
        SUF(cpx) v[64], y[64];
        const int rstep = n / r;             /* tw index step for the r-point DFT */
        for (int j0 = 0; j0 < nb; j0 += Ns) {
            for (int k = 0; k < Ns; k++) {
                const int j = j0 + k;
                for (int m = 0; m < r; m++) {
                    SUF(cpx) x = in[j + m * nb];
                    if (k && m) {
                        SUF(cpx) w = tw[(size_t)k * m * tstep], t;
                        t.re = x.re * w.re - x.im * w.im; t.im = x.re * w.im + x.im * w.re; x = t;
                    }
                    v[m] = x;
                }
                for (int q = 0; q < r; q++) {
                    REAL ar = 0, ai = 0;
                    for (int m = 0; m < r; m++) {
                        SUF(cpx) w = tw[(size_t)((q * m) % r) * rstep];
                        ar += v[m].re * w.re - v[m].im * w.im;
                        ai += v[m].re * w.im + v[m].im * w.re;
                    }
                    y[q].re = ar; y[q].im = ai;
                }
                SUF(cpx) *o = out + (size_t)j0 * r + k;
                for (int q = 0; q < r; q++) o[q * Ns] = y[q];
            }
        }
    }
}

/* out may alias in */
void SUF(fft_execute)(const SUF(fft_plan) *p, const SUF(cpx) *in, SUF(cpx) *out)
{
    const SUF(cpx) *src = in;
    SUF(cpx) *dst = p->buf0;
    int Ns = 1;
    if (p->npass == 0) { if (out != in) memcpy(out, in, sizeof(SUF(cpx)) * (size_t)p->n); return; }
    for (int s = 0; s < p->npass; s++) {
        const int last = (s == p->npass - 1);
        /* the last pass may write straight into `out` unless it is also its source */
        SUF(cpx) *d = (last && (const SUF(cpx) *)out != src) ? out : dst;
        SUF(fft_pass)(p, src, d, p->radix[s], Ns);
        Ns *= p->radix[s];
        src = d;
        dst = (d == p->buf0) ? p->buf1 : p->buf0;
    }
    if (src != out) memcpy(out, src, sizeof(SUF(cpx)) * (size_t)p->n);
}
