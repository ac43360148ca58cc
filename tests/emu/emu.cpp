// tests/emu/emu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU replay of the per-thread code of the CUDA kernels (csrc/ga_fft3.h is
// written __host__ __device__): every "thread" of a pass is run in a loop, a
// pass boundary is where the kernel has a __syncthreads().  This lets the index
// algebra, twiddles and butterflies be checked against numpy in the GPU-less
// development container.  It is never part of the product: the product path is
// the nvcc build of the same headers and fails loudly without a GPU.
#include <vector>
#include <cstring>
#include "ga_fft3.h"
#include "ga_tables.h"
#include "ga_pfa.h"
#include "ga_frontend_math.h"
#include "ga_frontend_host.h"

using namespace ga;

template <class G, int ORI>
static void emu_cell_sub(int s, int s_next, const cf *xd_blk, const cf *cext_sv, int dop, const cf *tw, const cf *ktab,
                         cf *sm, cf *acc)
{
    // what cell_rot_step<ORI> does between its barriers, thread by thread
    constexpr int NW = G::RC;
    constexpr int NORI = (ORI + 1) % 3;
    for (int j = 0; j < G::NB; j++) passB<G, +1, ORI>(j, s, tw, sm);
    int sp = 0, eoff = 0;
    if (s_next >= 0) cell_sub_offsets<G>(s_next, dop, sp, eoff);
    for (int j = 0; j < G::NC; j++) {          // pass C then, with NO barrier, pass A of the next sub-sequence
        cf a[NW];
        memcpy(a, &acc[(size_t)j * NW], sizeof a);
        cell_passC_acc<G, NW, ORI>(j, sm, ktab + (size_t)s * G::RC, a);
        memcpy(&acc[(size_t)j * NW], a, sizeof a);
        if (s_next >= 0)
            cell_passA<G, NORI>(j, s_next, xd_blk + (size_t)s_next * G::N2, cext_sv + (size_t)sp * 2 * G::N2 + eoff, tw, sm);
    }
}

template <class G>
static void emu_cell_t(const cf *xd_blk, const cf *cext_sv, int dop, int wlen, cf *y, float *best, int *besti, float *sum, int seg = -1)
{
    constexpr int NW = G::RC;
    // seg >= 0: output segment `seg` of the window (cell_kernel_tm<..., SEG = true>): its own ktab, window shortened by seg*N2
    std::vector<cf> tw = make_tw(G::N), ktab = seg >= 0 ? make_ktab_seg<G>(seg) : make_ktab<G>();
    if (seg > 0) wlen -= seg * G::N2;
    std::vector<cf> sm((size_t)G::SMEM_ELEMS);
    std::vector<cf> acc((size_t)G::NC * NW, mk(0, 0));
    if constexpr (G::ROT) {
        int sp, eoff;
        cell_sub_offsets<G>(0, dop, sp, eoff);
        for (int j = 0; j < G::NA; j++) cell_passA<G, 0>(j, 0, xd_blk, cext_sv + (size_t)sp * 2 * G::N2 + eoff, tw.data(), sm.data());
        for (int s = 0; s < G::N1; s++) {
            const int nxt = s + 1 < G::N1 ? s + 1 : -1;
            switch (s % 3) {
            case 0: emu_cell_sub<G, 0>(s, nxt, xd_blk, cext_sv, dop, tw.data(), ktab.data(), sm.data(), acc.data()); break;
            case 1: emu_cell_sub<G, 1>(s, nxt, xd_blk, cext_sv, dop, tw.data(), ktab.data(), sm.data(), acc.data()); break;
            default: emu_cell_sub<G, 2>(s, nxt, xd_blk, cext_sv, dop, tw.data(), ktab.data(), sm.data(), acc.data()); break;
            }
        }
    } else {
        for (int s = 0; s < G::N1; s++) {
            int sp, eoff;
            cell_sub_offsets<G>(s, dop, sp, eoff);
            const cf *xs = xd_blk + (size_t)s * G::N2;
            const cf *cs = cext_sv + (size_t)sp * 2 * G::N2 + eoff;
            for (int j = 0; j < G::NA; j++) cell_passA<G>(j, s, xs, cs, tw.data(), sm.data());
            for (int j = 0; j < G::NB; j++) passB<G, +1>(j, s, tw.data(), sm.data());
            for (int j = 0; j < G::NC; j++) {
                cf a[NW];
                memcpy(a, &acc[(size_t)j * NW], sizeof a);
                cell_passC_acc<G, NW>(j, sm.data(), ktab.data() + (size_t)s * G::RC, a);
                memcpy(&acc[(size_t)j * NW], a, sizeof a);
            }
        }
    }
    float b = 0, sm_ = 0; int bi = 0;
    for (int j = 0; j < G::NC; j++) {
        cf a[NW];
        memcpy(a, &acc[(size_t)j * NW], sizeof a);
        const int u = j / G::RB, v = j - u * G::RB, tau0 = u + G::RA * v;
        for (int w = 0; w < NW; w++) y[tau0 + G::OUT_STRIDE * w] = a[w];
        cell_peak_thread<G, NW>(a, tau0, wlen, b, bi, sm_);
    }
    *best = b; *besti = bi; *sum = sm_;
}

struct TimeSrc { const cf *x; cf operator()(int n) const { return x[n]; } };

template <class G>
static void emu_fwd_t(const cf *x, int s, cf *out)
{
    std::vector<cf> tw = make_tw(G::N), k1 = make_k1tab<G>();
    std::vector<cf> sm((size_t)G::SMEM_ELEMS);
    TimeSrc src{x};
    for (int j = 0; j < G::NA; j++) fwd_passA<G>(j, s, src, k1.data() + (size_t)s * G::N1, tw.data(), sm.data());
    for (int j = 0; j < G::NB; j++) passB<G, -1>(j, 0, tw.data(), sm.data());
    for (int j = 0; j < G::NC; j++) {
        cf p[G::RC];
        const int tau0 = passC<G, -1>(j, sm.data(), p);
        for (int w = 0; w < G::RC; w++) out[tau0 + G::OUT_STRIDE * w] = p[w];
    }
}

// Pass-A input of Sample() from packed bits, sample by sample (what BitSrc does in fwd_kernel, c/search_offline.cpp:143-153)
// or through the sample-group tables (fwd_lut_entry / fwd_lut_gather, ga_fft3.h): out[a*NA + j] = the value pass A feeds
// its radix-RA butterfly with, before the n2*s twiddle.
struct BitSrcHost {
    const unsigned char *chunk, *lo;
    cf operator()(int n) const
    {
        const int bit = (chunk[n >> 3] >> (n & 7)) & 1, l = lo[n];
        return mk((bit ^ (l & 1)) ? -1.0f : 1.0f, (bit ^ (l >> 1)) ? -1.0f : 1.0f);
    }
};
template <class G>
static void emu_gather_t(const unsigned char *chunk, const unsigned char *lo, int s, int use_lut, cf *out)
{
    std::vector<cf> k1 = make_k1tab<G>();
    const cf *k1s = k1.data() + (size_t)s * G::N1;
    if (!use_lut) {
        BitSrcHost src{chunk, lo};
        for (int n2 = 0; n2 < G::N2; n2++) {
            cf z = src(n2);
            for (int n1 = 1; n1 < G::N1; n1++) cfma(z, src(G::N2 * n1 + n2), k1s[n1]);      // as fwd_passA
            out[n2] = z;
        }
        return;
    }
    typedef FwdLut<G> L;
    std::vector<cf> lut((size_t)L::NG * L::ENTRIES);
    for (int e = 0; e < L::NG * L::ENTRIES; e++) lut[e] = fwd_lut_entry<G>(e, k1s);
    for (int n2 = 0; n2 < G::N2; n2++) {
        unsigned lm = 0;
        for (int n1 = 0; n1 < G::N1; n1++) lm |= (unsigned)(lo[(size_t)G::N2 * n1 + n2] & 3) << (2 * n1);   // as create_impl (gpsacq.cu)
        out[n2] = fwd_lut_gather<G>(chunk, n2, lm, lut.data());
    }
}

// ---- native W-point prime-factor transforms (ga_pfa.h), thread by thread like pfa_fwd_kernel / pfa_cell_kernel
template <class G>
static void emu_pfa_fwd_t(const cf *x, int conj, cf *out)
{
    typedef typename G::Fwd F;
    std::vector<cf> sm((size_t)F::SMEM_ELEMS);
    TimeSrc src{x};
    for (int j = 0; j < F::NA; j++) pfa_fwd_passA<F>(j, src, sm.data());
    for (int j = 0; j < F::NB; j++) pfa_passB<F, -1>(j, sm.data());
    for (int j = 0; j < F::NC; j++) {
        if (conj) pfa_fwd_passC_store<F, true>(j, sm.data(), out);
        else pfa_fwd_passC_store<F, false>(j, sm.data(), out);
    }
}

template <class G>
static void emu_pfa_cell_t(const cf *xs, const cf *cs_in, int q, cf *y, float *best, int *besti, float *sum, int *n_slow)
{
    // what pfa_rotate_replicas_kernel stores for this q: C rotated by -q
    const PfaRot rot = pfa_rotation<G>(-q);
    std::vector<cf> crot((size_t)G::W);
    for (int m = 0; m < G::W; m++) {
        int a = m / G::NA; const int j = m - a * G::NA;
        a += rot.da; if (a >= G::RA) a -= G::RA;
        crot[(size_t)m] = cs_in[a * G::NA + pfa_rot_col<G>(j, rot)];
    }
    const cf *cs = crot.data();
    std::vector<cf> sm((size_t)G::SMEM_ELEMS);
    for (int j = 0; j < G::NA; j++) pfa_cell_passA<G>(j, xs, cs, sm.data());
    for (int j = 0; j < G::NB; j++) pfa_passB<G, +1>(j, sm.data());
    float b = 0, s_ = 0; int bi = 0;
    for (int j = 0; j < G::NC; j++) {
        cf p[G::RC];
        const int t0 = pfa_passC_t0<G>(j);
        PfaPeak<G> pk;
        pk.init(t0);
        pfa_passC<G, +1>(j, sm.data(), [&](auto wc, cf v) {
            constexpr int w = decltype(wc)::value;
            p[w] = v;
            pk.template put<w>(fmaf(v.x, v.x, v.y * v.y));
        });
        for (int w = 0; w < G::RC; w++) y[pfa_lag<G>(t0, w)] = p[w];
        if (!pk.tie) pk.merge(b, bi, s_);
        else {                                  // what pfa_passC_exact does
            ++*n_slow;
            PfaPeakExact<G> pe;
            pe.init(t0);
            pfa_passC<G, +1>(j, sm.data(), [&](auto wc, cf v) { pe.template put<decltype(wc)::value>(fmaf(v.x, v.x, v.y * v.y)); });
            pe.merge(b, bi, s_);
        }
    }
    *best = b; *besti = bi; *sum = s_;
}

typedef PGeom<16, 11, 31> P5456;
typedef PGeom<24, 11, 31> P8184;
typedef PGeom<16, 25, 7> P2800;

typedef Geom<5, 20, 20, 20, true> G8000;
typedef Geom<4, 25, 20, 20> G10000;
typedef Geom<10, 20, 20, 10> G4000;
typedef Geom<2, 20, 20, 16> H6400;      // GRID: zero-padded embedding, L = 12800
typedef Geom<1, 16, 16, 16> X4096;      // GRID: exact-length twiddled transform, L = W = 4096

extern "C" {

int emu_geom(int id, int *n1, int *n2)
{
    switch (id) {
    case 0: *n1 = G8000::N1; *n2 = G8000::N2; return 0;
    case 1: *n1 = G10000::N1; *n2 = G10000::N2; return 0;
    case 2: *n1 = G4000::N1; *n2 = G4000::N2; return 0;
    case 3: *n1 = H6400::N1; *n2 = H6400::N2; return 0;
    case 4: *n1 = X4096::N1; *n2 = X4096::N2; return 0;
    }
    return -1;
}

// xd_blk: conj(X) decimated [N1][N2]; cext_sv: [N1][2*N2]; y: N2 complex outputs
int emu_cell(int id, const float *xd_blk, const float *cext_sv, int dop, int wlen, float *y, float *best, int *besti, float *sum)
{
    switch (id) {
    case 0: emu_cell_t<G8000>((const cf *)xd_blk, (const cf *)cext_sv, dop, wlen, (cf *)y, best, besti, sum); return 0;
    case 1: emu_cell_t<G10000>((const cf *)xd_blk, (const cf *)cext_sv, dop, wlen, (cf *)y, best, besti, sum); return 0;
    case 2: emu_cell_t<G4000>((const cf *)xd_blk, (const cf *)cext_sv, dop, wlen, (cf *)y, best, besti, sum); return 0;
    case 3: emu_cell_t<H6400>((const cf *)xd_blk, (const cf *)cext_sv, dop, wlen, (cf *)y, best, besti, sum); return 0;
    case 4: emu_cell_t<X4096>((const cf *)xd_blk, (const cf *)cext_sv, dop, wlen, (cf *)y, best, besti, sum); return 0;
    }
    return -1;
}

// x: W complex time samples; out: W spectral values in the cell's (a,b,c)-linear order (conjugated if conj)
// one output segment of a window longer than N2 (REF mode above 10 MHz): geometry 4 x 10000 only
int emu_cell_seg(int seg, const float *xd_blk, const float *cext_sv, int dop, int wlen, float *y, float *best, int *besti, float *sum)
{
    emu_cell_t<Geom<4, 25, 20, 20> >((const cf *)xd_blk, (const cf *)cext_sv, dop, wlen, (cf *)y, best, besti, sum, seg);
    *besti += seg * 10000;
    return 0;
}

int emu_pfa_fwd(int w, const float *x, int conj, float *out)
{
    switch (w) {
    case 5456: emu_pfa_fwd_t<P5456>((const cf *)x, conj, (cf *)out); return 0;
    case 8184: emu_pfa_fwd_t<P8184>((const cf *)x, conj, (cf *)out); return 0;
    case 2800: emu_pfa_fwd_t<P2800>((const cf *)x, conj, (cf *)out); return 0;
    }
    return -1;
}

// (a,b,c)-linear position m -> spectral index k (Good's map), for the tests
int emu_pfa_order(int w, int *k_of_m)
{
    auto fill = [&](auto g) {
        typedef decltype(g) G;
        for (int a = 0; a < G::RA; a++)
            for (int b = 0; b < G::RB; b++)
                for (int c = 0; c < G::RC; c++)
                    k_of_m[(a * G::RB + b) * G::RC + c] = (int)(((long long)a * G::KA + (long long)b * G::KB + (long long)c * G::KC) % G::W);
    };
    switch (w) {
    case 5456: fill(P5456{}); return 0;
    case 8184: fill(P8184{}); return 0;
    case 2800: fill(P2800{}); return 0;
    }
    return -1;
}

// xs = conj(X), cs = C, both in (a,b,c)-linear order; the replica operand is rotated by -q spectral bins first
// (the cell of Doppler bin r + R*q, ga_pfa.h); y: W outputs in natural lag order
int emu_pfa_cell(int w, const float *xs, const float *cs, int q, float *y, float *best, int *besti, float *sum, int *n_slow)
{
    *n_slow = 0;
    switch (w) {
    case 5456: emu_pfa_cell_t<P5456>((const cf *)xs, (const cf *)cs, q, (cf *)y, best, besti, sum, n_slow); return 0;
    case 8184: emu_pfa_cell_t<P8184>((const cf *)xs, (const cf *)cs, q, (cf *)y, best, besti, sum, n_slow); return 0;
    case 2800: emu_pfa_cell_t<P2800>((const cf *)xs, (const cf *)cs, q, (cf *)y, best, besti, sum, n_slow); return 0;
    }
    return -1;
}

// x: N complex time samples; out: N2 values X[N1*q+s]
int emu_gather(int id, const unsigned char *chunk, const unsigned char *lo, int s, int use_lut, float *out)
{
    switch (id) {
    case 0: emu_gather_t<G8000>(chunk, lo, s, use_lut, (cf *)out); return 0;
    case 1: emu_gather_t<G10000>(chunk, lo, s, use_lut, (cf *)out); return 0;
    case 2: emu_gather_t<G4000>(chunk, lo, s, use_lut, (cf *)out); return 0;
    }
    return -1;
}

int emu_fwd(int id, const float *x, int s, float *out)
{
    switch (id) {
    case 0: emu_fwd_t<G8000>((const cf *)x, s, (cf *)out); return 0;
    case 1: emu_fwd_t<G10000>((const cf *)x, s, (cf *)out); return 0;
    case 2: emu_fwd_t<G4000>((const cf *)x, s, (cf *)out); return 0;
    case 3: emu_fwd_t<H6400>((const cf *)x, s, (cf *)out); return 0;
    case 4: emu_fwd_t<X4096>((const cf *)x, s, (cf *)out); return 0;
    }
    return -1;
}

// ---- stream converters (csrc/ga_frontend.cuh): the kernels' thread functions, every thread in a loop ----------------
// iq8 -> bits through the threshold table (iq8_thr_build_kernel + iq8_to_bits_thr_kernel); n_threads = threads of the launch
int emu_iq8_thr(const unsigned char *iq, size_t n_samples, size_t n0, int fmt_s8, long long sum_i, long long sum_q, size_t n_total,
                const double *tab, unsigned p, unsigned q, unsigned n_threads, unsigned char *bits)
{
    const unsigned pitch = iq8_thr_pitch(q), n_active = n_threads / q * q, fx = fmt_s8 ? 0x80808080u : 0u;
    if (!n_active) return -1;
    const double mean_i = (double)sum_i / (double)n_total, mean_q = (double)sum_q / (double)n_total;
    std::vector<unsigned> thr((size_t)256 * pitch);
    for (unsigned k = 0; k < q; k++)                       // iq8_thr_build_kernel: block k, thread qu
        for (unsigned qu = 0; qu < 256; qu++) thr[(size_t)qu * pitch + k] = iq8_thr_entry((int)qu, mean_i, mean_q, tab[2 * k], tab[2 * k + 1]);
    const size_t n_bytes = n_samples / 8;
    std::vector<u32x4> in(n_bytes + 1);
    memcpy(in.data(), iq, 16 * n_bytes);
    for (unsigned tid = 0; tid < n_active; tid++)
        iq8_thr_thread(tid, n_active, in.data(), n_bytes, n0, fx, (tabref_t)thr.data(), 4u * pitch, p, q, bits);
    if (n_samples & 7) {
        unsigned ob = 0;
        for (size_t n = n_bytes * 8; n < n_samples; n++) {
            const size_t k = (((n0 + n) % q) * p) % q;
            const int f = (int)(fx & 0x80u);
            ob |= (iq8_r(iq[2 * n] ^ f, iq[2 * n + 1] ^ f, mean_i, mean_q, tab[2 * k], tab[2 * k + 1]) < 0.0 ? 1u : 0u) << (n & 7);
        }
        bits[n_bytes] = (unsigned char)ob;
    }
    return 0;
}

// the same decision sample by sample from the double expression (what iq8_to_bits_table_kernel evaluates)
int emu_iq8_direct(const unsigned char *iq, size_t n_samples, size_t n0, int fmt_s8, long long sum_i, long long sum_q, size_t n_total,
                   const double *tab, unsigned p, unsigned q, unsigned char *bits)
{
    const double mean_i = (double)sum_i / (double)n_total, mean_q = (double)sum_q / (double)n_total;
    const int f = fmt_s8 ? 0x80 : 0;
    memset(bits, 0, (n_samples + 7) / 8);
    for (size_t n = 0; n < n_samples; n++) {
        const size_t k = (((n0 + n) % q) * p) % q;
        if (iq8_r(iq[2 * n] ^ f, iq[2 * n + 1] ^ f, mean_i, mean_q, tab[2 * k], tab[2 * k + 1]) < 0.0) bits[n / 8] |= (unsigned char)(1u << (n & 7));
    }
    return 0;
}

// bits -> iq8, second version (bits_to_iq8_v2_kernel); lo: mu + lambda codes followed by 16 zero bytes
int emu_conv_v2(const unsigned char *bits, size_t n_bytes, size_t first_sample, const unsigned char *lo, unsigned long long mu,
                unsigned long long lambda, int amp, size_t n_threads, unsigned char *out)
{
    std::vector<u32x4> o(n_bytes + 1);
    std::vector<unsigned> lo_al((size_t)(mu + lambda + 16 + 3) / 4);          // 4-byte aligned like the cudaMalloc'ed table
    memcpy(lo_al.data(), lo, (size_t)(mu + lambda + 16));
    for (size_t tid = 0; tid < n_threads; tid++)
        conv_v2_thread(tid, n_threads, bits, n_bytes, first_sample, (const unsigned char *)lo_al.data(), mu, lambda, amp, o.data());
    memcpy(out, o.data(), 16 * n_bytes);
    return 0;
}

// ---- host-side table preparation of the converters (csrc/ga_frontend_host.h) ---------------------------------------------
int emu_small_rational(double x, unsigned long long *p, unsigned long long *q) { return small_rational(x, *p, *q) ? 1 : 0; }

// conv_lo_cycle's table applied to samples first .. first + n - 1 (k = i below mu, else mu + (i - mu) mod lambda), next to the
// float recurrence of c/conv_1bit_bin_to_hackrf_bin.cpp:33,79-80 run sample by sample from 0 (`direct`); returns 0 when the
// cycle search gives up
int emu_conv_lo_cycle(double fc, double fs, unsigned long long first, unsigned long long n, unsigned long long *mu,
                      unsigned long long *lambda, unsigned char *via_table, unsigned char *direct)
{
    LoCycle c;
    if (!conv_lo_cycle(fc, fs, first + n, c)) return 0;
    *mu = c.mu; *lambda = c.lambda;
    static const int lo_sin[4] = {1, 1, 0, 0}, lo_cos[4] = {1, 0, 0, 1};
    const float rate = (float)(4 * fc / fs);
    float ph = 0;
    for (unsigned long long i = 0; i < first + n; i++) {
        if (i >= first) {
            const unsigned long long k = i < c.mu ? i : c.mu + (i - c.mu) % c.lambda;
            via_table[i - first] = c.tab[k];
            const int ip = (int)ph;
            direct[i - first] = (unsigned char)(lo_sin[ip] | (lo_cos[ip] << 1));
        }
        ph += rate;
        if (ph >= 4) ph -= 4;
    }
    return 1;
}

}
