// ga_pfa.h -- native W-point transforms for GRID mode (1 ms blocks): three-factor Good-Thomas
// prime-factor algorithm, W = RA*RB*RC with pairwise coprime factors -- no twiddle factors at all.
// Written per butterfly like ga_fft3.h so that tests/emu replays it on the CPU.
//
// The 1 ms GPS block lengths are not 2/5-smooth: 5456 = 16*11*31, 8184 = 24*11*31, 2800 = 16*25*7.
// ga_grid.cuh evaluates their circular correlation through a zero-padded 2/5-smooth transform of
// more than twice the length; here the W-point DFT is computed directly:
//
//   spectral index  k  <->  (a,b,c):  k = (a*W/RA + b*W/RB + c*W/RC) mod W      (Good's map)
//   lag / time idx  t  <->  (u,v,w):  u = t mod RA, v = t mod RB, w = t mod RC   (CRT map)
//   w_W^(k*t) = w_RA^(a*u) * w_RB^(b*v) * w_RC^(c*w)   =>   a plain 3-D DFT over (a,b,c).
//
// Spectra live in HBM in (a,b,c)-linear order (index a*RB*RC + b*RC + c), which is exactly what pass A
// of the cell streams with coalesced float2 loads; the element-wise product conj(X)*C does not care
// about the order as long as both operands share it.  The forward transform uses the ROTATED
// factorisation (FA,FB,FC) = (RB,RC,RA): its pass-C thread (u,v) = the cell's (b,c) holds all RA
// values of the cell's a axis, so its stores are coalesced in the same (a,b,c)-linear order.
#pragma once
#include "ga_fft3.h"

namespace ga {

constexpr bool cx_coprime(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a == 1; }

// PGeom<RA,RB,RC>: Geom<1,...> gives the padded shared-memory strides and the butterfly counts
template <int RA_, int RB_, int RC_>
struct PGeom : Geom<1, RA_, RB_, RC_> {
    typedef Geom<1, RA_, RB_, RC_> B;
    static constexpr int W = RA_ * RB_ * RC_;
    static_assert(cx_coprime(RA_, RB_) && cx_coprime(RA_, RC_) && cx_coprime(RB_, RC_), "Good-Thomas needs coprime factors");
    // CRT idempotents: EA = 1 mod RA, 0 mod RB and RC, ...   t = (u*EA + v*EB + w*EC) mod W
    static constexpr int EA = (W / RA_) * cx_inv_mod((W / RA_) % RA_, RA_);
    static constexpr int EB = (W / RB_) * cx_inv_mod((W / RB_) % RB_, RB_);
    static constexpr int EC = (W / RC_) * cx_inv_mod((W / RC_) % RC_, RC_);
    // Good's map strides: k = (a*KA + b*KB + c*KC) mod W
    static constexpr int KA = W / RA_, KB = W / RB_, KC = W / RC_;
    // spectral index k -> k + 1 moves (a,b,c) by (IA,IB,IC) per axis, cyclically:  a = (k mod RA) * KA^-1 mod RA, ...
    static constexpr int IA = cx_inv_mod(KA % RA_, RA_), IB = cx_inv_mod(KB % RB_, RB_), IC = cx_inv_mod(KC % RC_, RC_);
    typedef PGeom<RB_, RC_, RA_> Fwd;     // the forward transform's factorisation (see header)
};

// Rotation of a spectrum stored in (a,b,c)-linear order by q spectral bins (S'[k] = S[k+q]): every axis shifts
// cyclically.  Used to share forward transforms between Doppler bins that are a whole DFT bin apart:
// exp(-j2pi(d + R)n/M) = exp(-j2pi d n/M) * exp(-j2pi n/W) when M = R*W, i.e. X_{d+R}[k] = X_d[k+1].  With
// d = r + R*q:  conj(X_d[k]) C[k] = conj(X_r[k+q]) C[k]  is a circular shift (by q) of  conj(X_r[k]) C[k-q],  and a
// circular shift of the product spectrum only puts a unit-modulus ramp on the lags: |y|^2 is unchanged.  So the
// cell multiplies the UNROTATED block spectrum X_r with a replica spectrum rotated by -q, and the rotated replicas
// are made once at create time (pfa_rotate_replicas_kernel) -- the hot loop pays nothing.
struct PfaRot { int da, db, dc; };
template <class G>
GA_HD PfaRot pfa_rotation(int q)
{
    PfaRot r;
    int qa = q % G::RA, qb = q % G::RB, qc = q % G::RC;
    if (qa < 0) qa += G::RA;
    if (qb < 0) qb += G::RB;
    if (qc < 0) qc += G::RC;
    r.da = (qa * G::IA) % G::RA; r.db = (qb * G::IB) % G::RB; r.dc = (qc * G::IC) % G::RC;
    return r;
}
// position (within a row of NA) that thread j = b*RC + c reads the rotated operand from
template <class G>
GA_HD int pfa_rot_col(int j, const PfaRot &r)
{
    int b = j / G::RC, c = j - b * G::RC;
    b += r.db; if (b >= G::RB) b -= G::RB;
    c += r.dc; if (c >= G::RC) c -= G::RC;
    return b * G::RC + c;
}

// ---- passes, in place in shared memory; DIR = +1 backward, -1 forward ------------------------
// pass A of a cell: thread j = b*RC + c of NA.  xs = conj(X) of the block (W values, (a,b,c) order),
// cs = replica spectrum in the same order.  prod = conj(X)*C  (c/search_offline.cpp:183-184).
template <class G>
GA_HD void pfa_cell_passA(int j, const cf *xs, const cf *cs, cf *sm)
{
    cf p[G::RA];
    GA_UNROLL
    for (int a = 0; a < G::RA; a++) p[a] = cmul(ldg(xs + a * G::NA + j), ldg(cs + a * G::NA + j));
    const int b = j / G::RC, c = j - b * G::RC;
    cf *dst = sm + G::template sb<0>() * b + c;
    radix_emit<G::RA, +1>(p, [&](auto uc, cf v) { dst[decltype(uc)::value * G::template sa<0>()] = v; });
}

// the same in two steps, for software pipelining: raw operand rows into registers, then product + butterfly
template <class G>
GA_HD void pfa_cell_loadA(int j, const cf *xs, const cf *cs, cf (&xv)[G::RA], cf (&cv)[G::RA])
{
    GA_UNROLL
    for (int a = 0; a < G::RA; a++) { xv[a] = ldg(xs + a * G::NA + j); cv[a] = ldg(cs + a * G::NA + j); }
}
template <class G>
GA_HD void pfa_cell_passA_regs(int j, const cf (&xv)[G::RA], const cf (&cv)[G::RA], cf *sm)
{
    cf p[G::RA];
    GA_UNROLL
    for (int a = 0; a < G::RA; a++) p[a] = cmul(xv[a], cv[a]);
    const int b = j / G::RC, c = j - b * G::RC;
    cf *dst = sm + G::template sb<0>() * b + c;
    radix_emit<G::RA, +1>(p, [&](auto uc, cf v) { dst[decltype(uc)::value * G::template sa<0>()] = v; });
}

template <class G, int DIR>
GA_HD void pfa_passB(int j2, cf *sm)
{
    const int u = j2 / G::RC, c = j2 - u * G::RC;
    cf *col = sm + G::template sa<0>() * u + c;
    cf p[G::RB];
    GA_UNROLL
    for (int b = 0; b < G::RB; b++) p[b] = col[b * G::template sb<0>()];
    radix_emit<G::RB, DIR>(p, [&](auto vc, cf v) { col[decltype(vc)::value * G::template sb<0>()] = v; });
}

// pass C: thread j3 = u*RB + v; emit(std::integral_constant<int, w>, value at (u,v,w)).
// pfa_passC_t0 = (u*EA + v*EB) mod W is the lag of w = 0; the lag of w is pfa_lag(t0, w) = (t0 + w*EC) mod W.
template <class G>
GA_HD int pfa_passC_t0(int j3)
{
    const int u = j3 / G::RB, v = j3 - u * G::RB;
    return (int)(((long long)u * G::EA + (long long)v * G::EB) % G::W);
}
template <class G, int DIR, class Emit>
GA_HD void pfa_passC(int j3, const cf *sm, Emit &&emit)
{
    const int u = j3 / G::RB, v = j3 - u * G::RB;
    const cf *row = sm + G::template sa<0>() * u + G::template sb<0>() * v;
    cf p[G::RC];
    GA_UNROLL
    for (int c = 0; c < G::RC; c++) p[c] = row[c];
    radix_emit<G::RC, DIR>(p, emit);
}

// lag of output w of a pass-C thread whose w = 0 lag is t0
template <class G>
GA_HD int pfa_lag(int t0, int w)
{
    const int t = t0 + (int)(((long long)w * G::EC) % G::W);
    return t >= G::W ? t - G::W : t;
}

// statistics of one pass-C butterfly (c/search_offline.cpp:190-194: power, FIRST maximum, sum).  The
// running maximum is tracked by output number w (an immediate in the unrolled code) with a strict
// compare, branch-free; lags are only formed when the butterfly's maximum is merged into the thread's
// (best, besti).  `tie` records whether some power was ever bit-equal to the running maximum: only then
// could "first maximum wins" depend on the order the outputs were produced in, and the caller redoes
// that butterfly with PfaPeakExact (practically never: exact float ties need degenerate input).
template <class G>
struct PfaPeak {
    float bpw; int bw, t0; float sum; bool tie;
    GA_HD void init(int t0_) { bpw = -1.0f; bw = 0; t0 = t0_; sum = 0.0f; tie = false; }
    template <int W_> GA_HD void put(float pw)
    {
        tie = tie || (pw == bpw);
        const bool gt = pw > bpw;
        bpw = gt ? pw : bpw;
        bw = gt ? W_ : bw;
        sum += pw;
    }
    GA_HD void merge(float &best, int &besti, float &tot) const
    {
        const int tau = pfa_lag<G>(t0, bw);
        if (bpw > best || (bpw == best && tau < besti)) { best = bpw; besti = tau; }
        tot += sum;
    }
};
// the same with the lag compared on every tie (slow path; identical sum order)
template <class G>
struct PfaPeakExact {
    float bpw; int btau, t0; float sum;
    GA_HD void init(int t0_) { bpw = -1.0f; btau = 0; t0 = t0_; sum = 0.0f; }
    template <int W_> GA_HD void put(float pw)
    {
        const int tau = pfa_lag<G>(t0, W_);
        if (pw > bpw || (pw == bpw && tau < btau)) { bpw = pw; btau = tau; }
        sum += pw;
    }
    GA_HD void merge(float &best, int &besti, float &tot) const
    {
        if (bpw > best || (bpw == best && btau < besti)) { best = bpw; besti = btau; }
        tot += sum;
    }
};

// ---- forward transform (DIR = -1) with factorisation F = G::Fwd -----------------------------
// pass A of the forward transform: thread j = b*FC + c; input sample index n = CRT(a,b,c) for F.
// src(n) returns time sample n.
template <class F, class Src>
GA_HD void pfa_fwd_passA(int j, const Src &src, cf *sm)
{
    const int b = j / F::RC, c = j - b * F::RC;
    int n = (int)(((long long)b * F::EB + (long long)c * F::EC) % F::W);
    cf p[F::RA];
    GA_UNROLL
    for (int a = 0; a < F::RA; a++) {
        p[a] = src(n);
        n += F::EA % F::W; if (n >= F::W) n -= F::W;
    }
    cf *dst = sm + F::template sb<0>() * b + c;
    radix_emit<F::RA, -1>(p, [&](auto uc, cf v) { dst[decltype(uc)::value * F::template sa<0>()] = v; });
}

// pass C of the forward transform + store: thread j3 = (u,v) of F holds X at Good's index
// k = u*KA_F + v*KB_F + w*KC_F; in the cell's (a,b,c) = (w,u,v) order that is element w*NC_F + j3.
template <class F, bool CONJ>
GA_HD void pfa_fwd_passC_store(int j3, const cf *sm, cf *out)
{
    pfa_passC<F, -1>(j3, sm, [&](auto wc, cf v) { out[decltype(wc)::value * F::NC + j3] = CONJ ? cconj(v) : v; });
}

}  // namespace ga
