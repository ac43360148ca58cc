// ga_fft3.h -- the N = N1 * (RA*RB*RC) transform used by every kernel of the
// acquisition engine, written per butterfly ("what does thread j do in pass X")
// so that the same code runs inside the CUDA kernels and in the CPU replay under
// tests/emu.
//
// Index algebra (DESIGN.md "FFT decomposition").  N2 = RA*RB*RC, N = N1*N2.
// A length-N2 sequence lives in shared memory as a 3-D array P[a][b][c],
// i = a*RB*RC + b*RC + c, at padded address a*SA + b*SB + c.  Three in-place
// passes transform one axis each; thread-private radix butterflies, twiddles on
// the outputs:
//
//   pass A  thread (b,c), j = b*RC+c :  t1[u,b,c] = tw^((N1*j + s)*u)      * sum_a w_RA^(au) P[a,b,c]
//   pass B  thread (u,c)             :  t2[u,v,c] = tw^((N1*c + s)*RA*v)   * sum_b w_RB^(bv) t1[u,b,c]
//   pass C  thread (u,v)             :  G [u,v,w] =                          sum_c w_RC^(cw) t2[u,v,c]
//
// with tw = exp(DIR*2*pi*i/N).  G[u,v,w] is output index tau = u + RA*v + RA*RB*w.
// With s = 0 this is a plain N2-point DFT.  With s = 0..N1-1 and the final
// factor exp(DIR*2*pi*i*s*w/(N1*RC)) it is the s-th term of the OUTPUT-PRUNED
// N-point DFT  y[tau] = sum_s sum_i P[N1*i+s] w_N^((N1*i+s)*tau), tau < N2,
// i.e. the reference's 40000-point backward FFT (c/search_offline.cpp:187)
// evaluated only where Correlate() looks (i < FS/1000, :190).
#pragma once
#include "ga_radix.h"

namespace ga {

constexpr int cmax3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

// Shared-memory layouts.  A logical element (a,b,c) sits at sa*a + sb*b + sc*c where
// (sa,sb,sc) is a rotation of the stride set {Q, P, 1} (P, Q odd so that a warp's 64-bit
// accesses spread over the banks whichever stride is lane-fastest):
//   ORI 0: (Q,P,1)   ORI 1: (1,Q,P)   ORI 2: (P,1,Q)
// In orientation k+1 the pass-A pencil of thread t covers exactly the addresses of the
// pass-C pencil of thread t in orientation k, so with ROT geometries (RA=RB=RC) consecutive
// sub-sequences alternate orientation and need NO barrier between pass C of one and pass A
// of the next (the same thread reads, then overwrites, its own 20 words).
template <int N1_, int RA_, int RB_, int RC_, bool ROT_ = false>
struct Geom {
    static constexpr int N1 = N1_, RA = RA_, RB = RB_, RC = RC_;
    static constexpr bool ROT = ROT_;
    static_assert(!ROT_ || (RA_ == RB_ && RB_ == RC_), "rotating layouts need equal radices");
    static constexpr int N2 = RA * RB * RC, N = N1 * N2;
    static constexpr int RMAX = cmax3(RA, RB, RC);
    static constexpr int P = ROT ? (RMAX | 1) : (RC | 1);
    static constexpr int Q = ROT ? ((P * RMAX) | 1) : RB * P;
    static constexpr int SMEM_ELEMS = ROT ? Q * RMAX : RA * Q;
    static constexpr int NORI = ROT ? 3 : 1;
    static constexpr int NA = RB * RC, NB = RA * RC, NC = RA * RB;   // butterflies per pass
    static constexpr int OUT_STRIDE = RA * RB;                        // tau step between a thread's outputs
    template <int ORI> static constexpr int sa() { return ORI == 0 ? Q : ORI == 1 ? 1 : P; }
    template <int ORI> static constexpr int sb() { return ORI == 0 ? P : ORI == 1 ? Q : 1; }
    template <int ORI> static constexpr int sc() { return ORI == 0 ? 1 : ORI == 1 ? P : Q; }
};

template <int DIR> GA_HD cf tw_load(const cf *tw, int idx)
{
    cf w = ldg(tw + idx);
    return DIR > 0 ? w : cconj(w);
}

// ---- pass A tail: butterfly along a, twiddle, store -----------------------------
template <class G, int DIR, int ORI = 0>
GA_HD void passA_finish(cf (&p)[G::RA], int j, int s_tw, const cf *tw, cf *sm)
{
    Radix<G::RA, DIR>::run(p);
    cf w[G::RA];
    unit_powers<G::RA>(tw_load<DIR>(tw, G::N1 * j + s_tw), w);
    const int b = j / G::RC, c = j - b * G::RC;
    cf *dst = sm + G::template sb<ORI>() * b + G::template sc<ORI>() * c;
    dst[0] = p[0];
    GA_UNROLL
    for (int u = 1; u < G::RA; u++) dst[u * G::template sa<ORI>()] = cmul(p[u], w[u]);
}

// ---- pass B: in place along b -----------------------------------------------------
template <class G, int DIR, int ORI = 0>
GA_HD void passB(int j2, int s_tw, const cf *tw, cf *sm)
{
    // lane-fastest index = the one with the small stride in this orientation
    int u, c;
    if (ORI == 0) { u = j2 / G::RC; c = j2 - u * G::RC; }
    else          { c = j2 / G::RA; u = j2 - c * G::RA; }
    cf *col = sm + G::template sa<ORI>() * u + G::template sc<ORI>() * c;
    cf p[G::RB];
    GA_UNROLL
    for (int b = 0; b < G::RB; b++) p[b] = col[b * G::template sb<ORI>()];
    Radix<G::RB, DIR>::run(p);
    cf w[G::RB];
    unit_powers<G::RB>(tw_load<DIR>(tw, (G::N1 * c + s_tw) * G::RA), w);
    col[0] = p[0];
    GA_UNROLL
    for (int v = 1; v < G::RB; v++) col[v * G::template sb<ORI>()] = cmul(p[v], w[v]);
}

// ---- pass C: along c, results stay in registers ---------------------------------
// returns tau0 = u + RA*v; p[w] is output tau0 + RA*RB*w
template <class G, int DIR, int ORI = 0>
GA_HD int passC(int j3, const cf *sm, cf (&p)[G::RC])
{
    const int u = j3 / G::RB, v = j3 - u * G::RB;
    const cf *row = sm + G::template sa<ORI>() * u + G::template sb<ORI>() * v;
    GA_UNROLL
    for (int c = 0; c < G::RC; c++) p[c] = row[c * G::template sc<ORI>()];
    Radix<G::RC, DIR>::run(p);
    return u + G::RA * v;
}

// =====================================================================================
// Cell = one (block, PRN, Doppler bin): shifted conj-multiply + pruned backward FFT.
// =====================================================================================
// For sub-sequence s of the product spectrum, prod[N1*i+s] = conj(X[N1*i+s]) *
// C[(N1*i+s-dop) mod N] (c/search_offline.cpp:181-185).  With s-dop = N1*e + sp
// (0 <= sp < N1) the code operand is the sp-th decimated sub-sequence of C rotated
// by e.  Replica spectra are stored decimated and doubled (Cext[sp][t], t < 2*N2,
// = C[N1*(t mod N2)+sp]) so the rotation is a plain offset.
template <class G>
GA_HD void cell_sub_offsets(int s, int dop, int &sp, int &eoff)
{
    const int v = s - dop;
    int q = v / G::N1, r = v - q * G::N1;
    if (r < 0) { r += G::N1; q -= 1; }       // floor division
    sp = r;
    eoff = q % G::N2; if (eoff < 0) eoff += G::N2;
}

// Which pass-A butterfly runs on lane-slot jj (0 <= jj < NA) of the CTA.  The tile pitch of a b-row is RC+1 (odd, for
// passes B and C), so 16 lanes that straddle two b-rows store to 16 slots with one slot skipped in the middle: the
// first and the last of them share a bank pair and the 64-bit store takes two wavefronts instead of one (20 % of all
// shared-memory wavefronts of the kernel).  For RB = RC = 20 the 400 butterflies are therefore dealt in half-warps that
// never do that: 20 half-warps (b, c = 0..15) and 5 half-warps of the left-over columns c = 16..19 of four rows whose
// bank offsets are 4 apart (rows g, g+4, g+8, g+12; the last group takes rows 16..19 and keeps a 3-slot overlap).
// The operand loads stay whole 32-byte sectors (16 or 4 consecutive columns).
template <class G> GA_HD int passA_slot_to_j(int jj)
{
    if (G::RB == 20 && G::RC == 20) {
        const int h = jj >> 4, i = jj & 15;
        if (h < 20) return 20 * h + i;
        const int g = h - 20, r = i >> 2, row = g < 4 ? g + 4 * r : 16 + r;
        return 20 * row + 16 + (i & 3);
    }
    return jj;
}

// pass A of a cell: thread j of NA.  xs = conj(X) sub-sequence s (N2 values),
// cs = Cext[sv][sp] + eoff.
template <class G, int ORI = 0>
GA_HD void cell_passA(int j, int s, const cf *xs, const cf *cs, const cf *tw, cf *sm)
{
    cf p[G::RA];
    GA_UNROLL
    for (int a = 0; a < G::RA; a++) {
        const cf x = ldg(xs + a * G::NA + j), c = ldg(cs + a * G::NA + j);
        p[a] = cmul(x, c);
    }
    passA_finish<G, +1, ORI>(p, j, s, tw, sm);
}

// pass C of a cell with accumulation of term s into the thread's outputs.
// ks = ktab + s*RC, ktab[s*RC+w] = exp(+2*pi*i*s*w/(N1*RC)).
template <class G, int NW, int ORI = 0>
GA_HD void cell_passC_acc(int j3, const cf *sm, const cf *ks, cf (&acc)[NW])
{
    cf p[G::RC];
    passC<G, +1, ORI>(j3, sm, p);
#ifdef GA_EXPERIMENT_FEWACC   // timing experiment only (wrong results): 2 accumulators instead of NW
    GA_UNROLL
    for (int w = 0; w < NW; w++) cfma(acc[w & 1], p[w], ks[w]);
#else
    GA_UNROLL
    for (int w = 0; w < NW; w++) cfma(acc[w], p[w], ks[w]);
#endif
}

// peak/sum over one thread's outputs (c/search_offline.cpp:190-194): power,
// first-maximum (lowest tau wins ties), running sum.  wlen = search window.
template <class G, int NW>
GA_HD void cell_peak_thread(const cf (&acc)[NW], int tau0, int wlen, float &best, int &besti, float &sum)
{
    GA_UNROLL
    for (int w = 0; w < NW; w++) {
        const int tau = tau0 + G::OUT_STRIDE * w;
        if (tau < wlen) {
            const float pwr = fmaf(acc[w].x, acc[w].x, acc[w].y * acc[w].y);
            if (pwr > best || (pwr == best && tau < besti)) { best = pwr; besti = tau; }
            sum += pwr;
        }
    }
}

// =====================================================================================
// Forward transform with decimated output:  X[N1*q+s], q < N2, for one s.
//   X[N1*q+s] = sum_{n2<N2} w_N2^(-n2*q) * [ w_N^(-n2*s) * sum_{n1<N1} x[N2*n1+n2] w_N1^(-n1*s) ]
// The bracket is the pass-A input; the outer sum is the plain (s_tw = 0) N2-point
// forward transform.  Used for the sample blocks (Sample(), c/search_offline.cpp:161)
// and the replicas (SearchInit(), :105).
// =====================================================================================
// src(n) returns time sample n as cf.  k1s = k1tab + s*N1, k1tab[s*N1+n1] = exp(-2*pi*i*n1*s/N1).
template <class G, class Src>
GA_HD void fwd_passA(int j, int s, const Src &src, const cf *k1s, const cf *tw, cf *sm)
{
    cf p[G::RA];
    GA_UNROLL
    for (int a = 0; a < G::RA; a++) {
        const int n2 = a * G::NA + j;
        cf z = src(n2);                                   // n1 = 0 term, K1 = 1
        for (int n1 = 1; n1 < G::N1; n1++) cfma(z, src(G::N2 * n1 + n2), k1s[n1]);
        p[a] = (s == 0) ? z : cmul(z, tw_load<-1>(tw, n2 * s));
    }
    passA_finish<G, -1>(p, j, 0, tw, sm);
}

// ---- the same pass-A input by table ---------------------------------------------------------------
// In Sample() (c/search_offline.cpp:143-153) a time sample is (+-1, +-1): one data bit XOR the LO's cos / sin bit.  The
// radix-N1 gather z = x[n2] + sum_{n1>=1} x[N2*n1 + n2] * K1[s][n1] of fwd_passA therefore takes one of 4^N1 values per
// sub-sequence s.  FwdLut<G>: the values of groups of GS <= 5 samples (N1 = 10: two groups, z = z_0 + z_1), built in
// the operation order of fwd_passA -- for N1 <= 5 the table value IS what fwd_passA computes, bit for bit.
// Index: 2 bits per sample, bit 0 = "real part negative" = data ^ lo_cos, bit 1 = "imaginary part negative" = data ^ lo_sin.
template <class G> struct FwdLut {
    static constexpr int GS = G::N1 <= 5 ? G::N1 : 5, NG = (G::N1 + GS - 1) / GS;
    static constexpr int ENTRIES = 1 << (2 * GS);
    static constexpr int BYTES = NG * ENTRIES * (int)sizeof(cf);
    static_assert(G::N2 % 8 == 0, "the samples of a group must sit at the same bit of their bytes");
};
template <class G> GA_HD cf fwd_lut_entry(int e, const cf *k1s)
{
    typedef FwdLut<G> L;
    const int g = e / L::ENTRIES, code = e - g * L::ENTRIES, first = g * L::GS;
    const int last = first + L::GS < G::N1 ? first + L::GS : G::N1;
    const int c0 = code & 3;
    const cf x0 = mk((c0 & 1) ? -1.0f : 1.0f, (c0 & 2) ? -1.0f : 1.0f);
    cf z = first == 0 ? x0 : cmul(x0, k1s[first]);
    for (int n1 = first + 1; n1 < last; n1++) {
        const int c = (code >> (2 * (n1 - first))) & 3;
        cfma(z, mk((c & 1) ? -1.0f : 1.0f, (c & 2) ? -1.0f : 1.0f), k1s[n1]);
    }
    return z;
}
// lomask_n2 = sum_n1 (lo[N2*n1 + n2] & 3) << 2*n1, lo[n] = lo_cos bit | lo_sin bit << 1 at sample n
template <class G> GA_HD cf fwd_lut_gather(const unsigned char *chunk, int n2, unsigned lomask_n2, const cf *lut)
{
    typedef FwdLut<G> L;
    const int sh = n2 & 7;
    const unsigned char *b = chunk + (n2 >> 3);
    unsigned idx = 0;
    GA_UNROLL
    for (int n1 = 0; n1 < G::N1; n1++) idx |= ((((unsigned)b[n1 * (G::N2 / 8)] >> sh) & 1u) * 3u) << (2 * n1);
    idx ^= lomask_n2;
    cf z = lut[idx & (L::ENTRIES - 1)];
    if (L::NG > 1) z = cadd(z, lut[L::ENTRIES + (idx >> (2 * L::GS))]);
    return z;
}

}  // namespace ga
