// ga_experiments.cuh -- variants of the REF hot kernel that were built, measured on B200 and found SLOWER than
// cell_kernel_tm (ga_kernels.cuh).  They stay selectable (GA_CELL_TMEM=0, GA_CELL_PIPE=1, GA_G8000_ROT=1 in
// gpsacq.cu) so the measurements in DESIGN.md section 4 can be repeated; the default build never instantiates them.
//
//   cell_kernel       accumulators in registers (spills: 77 M local-memory instructions per launch)
//   cell_kernel_tm2   operand prefetch behind the barriers (register pressure -> spills, slower)
//   cell_kernel_rot   rotating shared-memory orientations, two barriers per sub-sequence (-12 %: I-cache)
#pragma once

namespace ga {


// ---------------------------------------------------------------------------------
// The hot kernel.  Persistent CTAs; each loop iteration is one (block, Doppler) cell:
//   for s < N1:  pass A (global loads of conj(X)_s and rotated C_sp, multiply, radix-RA, twiddle) -> smem
//                pass B (radix-RB in place) ; pass C (radix-RC) -> += into register accumulators
//   |acc|^2, first-max / sum over tau < W, warp-shuffle + smem reduction, one 16-byte record out.
// Nothing but the operands is read from and nothing but the record is written to global memory.
// ---------------------------------------------------------------------------------
template <class G, int T, int NW, int MINB, int GID>
__global__ void __launch_bounds__(T, MINB) cell_kernel(const cf *__restrict__ xd, const cf *__restrict__ cext,
                                                       const int *__restrict__ sv_of_block,
                                                       const cf *__restrict__ tw,
                                                       int n_cells, int n_dop, int dmax, int wlen,
                                                       CellStat *__restrict__ cells)
{
    constexpr int ITA = cdiv(G::NA, T), ITB = cdiv(G::NB, T), ITC = cdiv(G::NC, T);
    constexpr int NWARP = cdiv(T, 32);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    __shared__ float red_best[NWARP], red_sum[NWARP];
    __shared__ int red_idx[NWARP];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int blk = cell / n_dop, dop = cell - blk * n_dop - dmax;
        const int sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
        const cf *xb = xd + (size_t)blk * G::N;
        const cf *cb = cext + (size_t)sv * (2 * G::N);

        cf acc[ITC][NW];
#pragma unroll
        for (int it = 0; it < ITC; it++)
#pragma unroll
            for (int w = 0; w < NW; w++) acc[it][w] = mk(0.0f, 0.0f);

        for (int s = 0; s < G::N1; s++) {
            int sp, eoff;
            cell_sub_offsets<G>(s, dop, sp, eoff);
            const cf *xs = xb + (size_t)s * G::N2;
            const cf *cs = cb + (size_t)sp * (2 * G::N2) + eoff;
#pragma unroll
            for (int it = 0; it < ITA; it++) {
                const int j = tid + it * T;
                if (ITA * T == G::NA || j < G::NA) cell_passA<G>(j, s, xs, cs, tw, sm);
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITB; it++) {
                const int j = tid + it * T;
                if (ITB * T == G::NB || j < G::NB) passB<G, +1>(j, s, tw, sm);
            }
            __syncthreads();
            const cf *ks = c_ktab[GID] + s * G::RC;
#pragma unroll
            for (int it = 0; it < ITC; it++) {
                const int j = tid + it * T;
                if (ITC * T == G::NC || j < G::NC) cell_passC_acc<G, NW>(j, sm, ks, acc[it]);
            }
            __syncthreads();      // smem is rewritten by the next sub-sequence's pass A
        }

        float best = 0.0f, sum = 0.0f;
        int besti = 0;
#pragma unroll
        for (int it = 0; it < ITC; it++) {
            const int j = tid + it * T;
            if (ITC * T == G::NC || j < G::NC) {
                const int u = j / G::RB, v = j - u * G::RB;
                cell_peak_thread<G, NW>(acc[it], u + G::RA * v, wlen, best, besti, sum);
            }
        }
        // warp-shuffle reduction; ties go to the lower index = "first maximum wins" (:192)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, besti, off);
            const float os = __shfl_down_sync(0xffffffffu, sum, off);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            sum += os;
        }
        if (lane == 0) { red_best[wid] = best; red_idx[wid] = besti; red_sum[wid] = sum; }
        __syncthreads();
        if (wid == 0) {
            best = lane < NWARP ? red_best[lane] : 0.0f;
            besti = lane < NWARP ? red_idx[lane] : 0x7fffffff;
            sum = lane < NWARP ? red_sum[lane] : 0.0f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, besti, off);
                const float os = __shfl_down_sync(0xffffffffu, sum, off);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                sum += os;
            }
            if (lane == 0) {
                CellStat r; r.max_pwr = best; r.tot_pwr = sum; r.max_idx = besti; r.pad = 0;
                cells[cell] = r;
            }
        }
        // red_* are rewritten only after the next cell's __syncthreads()s: no extra barrier needed
    }
}


// ---------------------------------------------------------------------------------
// Software-pipelined variant of the TMEM kernel (one 32-butterfly task per warp).
//
// ncu on cell_kernel_tm: the largest stall is pass A waiting for its 40 operand loads (L2 hits,
// ~600+ cycles), followed by the barrier waits those slow warps cause.  Here the operands of the
// NEXT sub-sequence (or of the next cell's first one) are fetched while the current one computes:
// at the three points where a thread has just stored its results and its registers are free --
// right before each block barrier -- it issues the loads of 8 / 8 / 4 of the 20 rows, and right after
// the barrier it multiplies them (conj(X)*C) and parks the products in tensor memory (16 rows; the
// last 4 stay in registers for the pass A that follows immediately).  Load latency overlaps the
// barrier wait; pass A itself starts from TMEM (12-cycle latency) instead of L2.
//   TMEM columns per thread: [0,28) output accumulators, [32,64) prefetched products.
// ---------------------------------------------------------------------------------
template <class G, int T, int NW, int GID>
__global__ void __launch_bounds__(T, TM_MINB) cell_kernel_tm2(const cf *__restrict__ xd, const cf *__restrict__ cext,
                                                        const int *__restrict__ sv_of_block, const cf *__restrict__ tw,
                                                        int n_cells, int n_dop, int dmax, int wlen, CellStat *__restrict__ cells)
{
    static_assert(T % 32 == 0, "whole warps only");
    constexpr int NWARP = T / 32;
    constexpr int NTA = cdiv(G::NA, 32);
    static_assert(G::NA == G::NB && G::NB == G::NC && NTA <= NWARP, "one task per warp, same thread map in all passes");
    static_assert(G::RA == 20 && 2 * NW <= 32, "row chunks 8/8/4 and the TMEM map below assume radix-20 pass A");
    constexpr uint32_t COL_PROD = 32, COL_SLOT = 64;
    constexpr uint32_t TM_COLS = pow2_at_least(COL_SLOT * cdiv(NWARP, 4));
    static_assert(TM_COLS * TM_MINB <= 512, "TMEM columns");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    __shared__ float red_best[NWARP], red_sum[NWARP];
    __shared__ int red_idx[NWARP];
    __shared__ uint32_t tm_base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_base_s)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_base_s;
    const uint32_t tm_mine = tm_base + ((32u * (uint32_t)(wid & 3)) << 16) + (uint32_t)(wid >> 2) * COL_SLOT;

    const bool warp_on = wid < NTA;                 // warps beyond the last task only take part in barriers
    const int j = wid * 32 + lane;
    const bool act = warp_on && j < G::NA;
    const int jc = j < G::NA ? j : G::NA - 1;       // lanes past the end shadow the last butterfly (no stores)

    // operand streams of a (cell, s): xs = conj(X) sub-sequence, cs = rotated replica sub-sequence
    auto operands = [&](int cell, int s, const cf *&xs, const cf *&cs) {
        const int blk = cell / n_dop, dop = cell - blk * n_dop - dmax;
        const int sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
        int sp, eoff;
        cell_sub_offsets<G>(s, dop, sp, eoff);
        xs = xd + (size_t)blk * G::N + (size_t)s * G::N2 + jc;
        cs = cext + (size_t)sv * (2 * G::N) + (size_t)sp * (2 * G::N2) + eoff + jc;
    };

    int cell = blockIdx.x, s = 0;
    if (cell >= n_cells) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
        return;
    }
    const cf *xs, *cs;
    operands(cell, 0, xs, cs);
    cf tail[4];                                     // products of rows 16..19 of the upcoming pass A
    if (warp_on) {                                  // prologue: everything for the very first sub-sequence
#pragma unroll
        for (int c8 = 0; c8 < 2; c8++) {
            float a[16];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const cf p = cmul(ldg(xs + (c8 * 8 + r) * G::NA), ldg(cs + (c8 * 8 + r) * G::NA));
                a[2 * r] = p.x; a[2 * r + 1] = p.y;
            }
            tm_move<16, false>(tm_mine + COL_PROD + 16 * c8, a);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) tail[r] = cmul(ldg(xs + (16 + r) * G::NA), ldg(cs + (16 + r) * G::NA));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    float best = 0.0f, sum = 0.0f;
    int besti = 0;

    for (;;) {
        // ---- who is next (for the prefetch) -------------------------------------------------------
        int ncell = cell, ns = s + 1;
        if (ns == G::N1) { ns = 0; ncell = cell + gridDim.x; }
        const bool has_next = ncell < n_cells;
        const cf *nxs = xs, *ncs = cs;
        if (has_next) operands(ncell, ns, nxs, ncs);

        // ---- pass A: 16 products from TMEM + 4 from registers -> radix-20 -> twiddle -> smem ------
        cf ld0[8], ld1[8];
        if (warp_on) {
            cf p[G::RA];
            {
                float a[32];
                tm_move<32, true>(tm_mine + COL_PROD, a);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int r = 0; r < 16; r++) p[r] = mk(a[2 * r], a[2 * r + 1]);
            }
#pragma unroll
            for (int r = 0; r < 4; r++) p[16 + r] = tail[r];
            if (act) passA_finish<G, +1>(p, j, s, tw, sm);
            if (has_next) {                          // registers are free now: fetch rows 0..7 of the next one
#pragma unroll
                for (int r = 0; r < 4; r++) { ld0[r] = ldg(nxs + r * G::NA); ld1[r] = ldg(ncs + r * G::NA); }
#pragma unroll
                for (int r = 4; r < 8; r++) { ld0[r] = ldg(nxs + r * G::NA); ld1[r] = ldg(ncs + r * G::NA); }
            }
        }
        __syncthreads();
        if (warp_on) {
            if (has_next) {
                float a[16];
#pragma unroll
                for (int r = 0; r < 8; r++) { const cf q = cmul(ld0[r], ld1[r]); a[2 * r] = q.x; a[2 * r + 1] = q.y; }
                tm_move<16, false>(tm_mine + COL_PROD, a);
            }
            // ---- pass B ---------------------------------------------------------------------------
            if (act) passB<G, +1>(j, s, tw, sm);
            if (has_next) {                          // rows 8..15
#pragma unroll
                for (int r = 0; r < 8; r++) { ld0[r] = ldg(nxs + (8 + r) * G::NA); ld1[r] = ldg(ncs + (8 + r) * G::NA); }
            }
        }
        __syncthreads();
        if (warp_on) {
            if (has_next) {
                float a[16];
#pragma unroll
                for (int r = 0; r < 8; r++) { const cf q = cmul(ld0[r], ld1[r]); a[2 * r] = q.x; a[2 * r + 1] = q.y; }
                tm_move<16, false>(tm_mine + COL_PROD + 16, a);
            }
            // ---- pass C + accumulation in TMEM ----------------------------------------------------
            const cf *ks = c_ktab[GID] + s * G::RC;
            cf p[G::RC];
            const int tau0 = passC<G, +1>(jc, sm, p);
            float a[2 * NW];
            if (s == 0) {
#pragma unroll
                for (int w = 0; w < NW; w++) { a[2 * w] = p[w].x; a[2 * w + 1] = p[w].y; }
            } else {
                tm_move<2 * NW, true>(tm_mine, a);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    cf t = mk(a[2 * w], a[2 * w + 1]);
                    cfma(t, p[w], ks[w]);
                    a[2 * w] = t.x; a[2 * w + 1] = t.y;
                }
            }
            if (s < G::N1 - 1) {
                tm_move<2 * NW, false>(tm_mine, a);
            } else if (act) {
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const int tau = tau0 + G::OUT_STRIDE * w;
                    if (tau < wlen) {
                        const float pwr = fmaf(a[2 * w], a[2 * w], a[2 * w + 1] * a[2 * w + 1]);
                        if (pwr > best || (pwr == best && tau < besti)) { best = pwr; besti = tau; }
                        sum += pwr;
                    }
                }
            }
            if (has_next) {                          // rows 16..19 stay in registers for the coming pass A
#pragma unroll
                for (int r = 0; r < 4; r++) { ld0[r] = ldg(nxs + (16 + r) * G::NA); ld1[r] = ldg(ncs + (16 + r) * G::NA); }
            }
        }

        if (s == G::N1 - 1) {
            // ---- the cell is complete: reduce and write its record ----------------------------------
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, besti, off);
                const float os = __shfl_down_sync(0xffffffffu, sum, off);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                sum += os;
            }
            if (lane == 0) { red_best[wid] = best; red_idx[wid] = besti; red_sum[wid] = sum; }
            __syncthreads();
            if (wid == 0) {
                best = lane < NWARP ? red_best[lane] : 0.0f;
                besti = lane < NWARP ? red_idx[lane] : 0x7fffffff;
                sum = lane < NWARP ? red_sum[lane] : 0.0f;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ob = __shfl_down_sync(0xffffffffu, best, off);
                    const int oi = __shfl_down_sync(0xffffffffu, besti, off);
                    const float os = __shfl_down_sync(0xffffffffu, sum, off);
                    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                    sum += os;
                }
                if (lane == 0) {
                    CellStat r; r.max_pwr = best; r.tot_pwr = sum; r.max_idx = besti; r.pad = 0;
                    cells[cell] = r;
                }
            }
            best = 0.0f; sum = 0.0f; besti = 0;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        __syncthreads();                             // smem is rewritten by the next pass A
        if (!has_next) break;
        if (warp_on) {
#pragma unroll
            for (int r = 0; r < 4; r++) tail[r] = cmul(ld0[r], ld1[r]);
        }
        cell = ncell; s = ns; xs = nxs; cs = ncs;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------
// Rotating-layout variant of the hot kernel (geometries with RA=RB=RC, e.g. 5 x 20^3).
// Consecutive sub-sequences use shared-memory orientations 0,1,2,0,... (ga_fft3.h): the
// pass-A pencil a thread writes for sub-sequence g+1 is the pass-C pencil it has just read
// for sub-sequence g, so only TWO barriers per sub-sequence remain (before and after pass
// B) and the operand loads of the next sub-sequence are free to overlap pass C.
// The orientation keeps rotating across cells, so there is no barrier between cells either.
// ---------------------------------------------------------------------------------
template <class G, int T, int NW>
struct CellState {
    static constexpr int ITC = cdiv(G::NC, T);
    cf acc[ITC][NW];
    const cf *xb, *cb;
    int dop, cell;
};

template <class G, int T, int NW, int GID, int ORI>
__device__ __forceinline__ bool cell_rot_step(CellState<G, T, NW> &st, int s, cf *sm, const cf *__restrict__ xd,
                                              const cf *__restrict__ cext, const int *__restrict__ sv_of_block,
                                              const cf *__restrict__ tw, int n_cells, int n_dop, int dmax, int wlen,
                                              CellStat *__restrict__ cells, float *red_best, float *red_sum, int *red_idx)
{
    constexpr int ITA = cdiv(G::NA, T), ITB = cdiv(G::NB, T), ITC = cdiv(G::NC, T);
    constexpr int NWARP = cdiv(T, 32);
    constexpr int NORI = (ORI + 1) % 3;
    static_assert(ITA == ITC && G::NA == G::NC, "pass A and pass C must share the thread map");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITB; it++) {
        const int j = tid + it * T;
        if (ITB * T == G::NB || j < G::NB) passB<G, +1, ORI>(j, s, tw, sm);
    }
    __syncthreads();
    const cf *ks = c_ktab[GID] + s * G::RC;
#pragma unroll
    for (int it = 0; it < ITC; it++) {
        const int j = tid + it * T;
        if (ITC * T == G::NC || j < G::NC) cell_passC_acc<G, NW, ORI>(j, sm, ks, st.acc[it]);
    }

    int s_next = s + 1;
    if (s == G::N1 - 1) {
        // ---- the cell is complete: |.|^2, first-max / sum, reduce, write the record --------
        float best = 0.0f, sum = 0.0f;
        int besti = 0;
#pragma unroll
        for (int it = 0; it < ITC; it++) {
            const int j = tid + it * T;
            if (ITC * T == G::NC || j < G::NC) {
                const int u = j / G::RB, v = j - u * G::RB;
                cell_peak_thread<G, NW>(st.acc[it], u + G::RA * v, wlen, best, besti, sum);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, besti, off);
            const float os = __shfl_down_sync(0xffffffffu, sum, off);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            sum += os;
        }
        if (lane == 0) { red_best[wid] = best; red_idx[wid] = besti; red_sum[wid] = sum; }
        __syncthreads();
        if (wid == 0) {
            best = lane < NWARP ? red_best[lane] : 0.0f;
            besti = lane < NWARP ? red_idx[lane] : 0x7fffffff;
            sum = lane < NWARP ? red_sum[lane] : 0.0f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, besti, off);
                const float os = __shfl_down_sync(0xffffffffu, sum, off);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                sum += os;
            }
            if (lane == 0) {
                CellStat r; r.max_pwr = best; r.tot_pwr = sum; r.max_idx = besti; r.pad = 0;
                cells[st.cell] = r;
            }
        }
        // ---- next cell of this persistent CTA -------------------------------------------------
        st.cell += gridDim.x;
        if (st.cell >= n_cells) return false;
        const int blk = st.cell / n_dop;
        st.dop = st.cell - blk * n_dop - dmax;
        const int sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
        st.xb = xd + (size_t)blk * G::N;
        st.cb = cext + (size_t)sv * (2 * G::N);
#pragma unroll
        for (int it = 0; it < ITC; it++)
#pragma unroll
            for (int w = 0; w < NW; w++) st.acc[it][w] = mk(0.0f, 0.0f);
        s_next = 0;
    }
    // ---- pass A of the next sub-sequence, into the rows this thread has just consumed ---------
    int sp, eoff;
    cell_sub_offsets<G>(s_next, st.dop, sp, eoff);
    const cf *xs = st.xb + (size_t)s_next * G::N2;
    const cf *cs = st.cb + (size_t)sp * (2 * G::N2) + eoff;
#pragma unroll
    for (int it = 0; it < ITA; it++) {
        const int j = tid + it * T;
        if (ITA * T == G::NA || j < G::NA) cell_passA<G, NORI>(j, s_next, xs, cs, tw, sm);
    }
    return true;
}

template <class G, int T, int NW, int MAXREG, int GID>
__global__ void __launch_bounds__(T) __maxnreg__(MAXREG)
cell_kernel_rot(const cf *__restrict__ xd, const cf *__restrict__ cext, const int *__restrict__ sv_of_block,
                const cf *__restrict__ tw, int n_cells, int n_dop, int dmax, int wlen, CellStat *__restrict__ cells)
{
    static_assert(G::ROT, "rotating-layout geometry required");
    constexpr int ITA = cdiv(G::NA, T), ITC = cdiv(G::NC, T);
    constexpr int NWARP = cdiv(T, 32);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    __shared__ float red_best[NWARP], red_sum[NWARP];
    __shared__ int red_idx[NWARP];
    const int tid = threadIdx.x;

    CellState<G, T, NW> st;
    st.cell = blockIdx.x;
    if (st.cell >= n_cells) return;
    {
        const int blk = st.cell / n_dop;
        st.dop = st.cell - blk * n_dop - dmax;
        const int sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
        st.xb = xd + (size_t)blk * G::N;
        st.cb = cext + (size_t)sv * (2 * G::N);
    }
#pragma unroll
    for (int it = 0; it < ITC; it++)
#pragma unroll
        for (int w = 0; w < NW; w++) st.acc[it][w] = mk(0.0f, 0.0f);
    {
        int sp, eoff;
        cell_sub_offsets<G>(0, st.dop, sp, eoff);
        const cf *cs = st.cb + (size_t)sp * (2 * G::N2) + eoff;
#pragma unroll
        for (int it = 0; it < ITA; it++) {
            const int j = tid + it * T;
            if (ITA * T == G::NA || j < G::NA) cell_passA<G, 0>(j, 0, st.xb, cs, tw, sm);
        }
    }
    int s = 0, ori = 0;
    for (;;) {
        bool more;
        if (ori == 0)
            more = cell_rot_step<G, T, NW, GID, 0>(st, s, sm, xd, cext, sv_of_block, tw, n_cells, n_dop, dmax, wlen, cells, red_best, red_sum, red_idx);
        else if (ori == 1)
            more = cell_rot_step<G, T, NW, GID, 1>(st, s, sm, xd, cext, sv_of_block, tw, n_cells, n_dop, dmax, wlen, cells, red_best, red_sum, red_idx);
        else
            more = cell_rot_step<G, T, NW, GID, 2>(st, s, sm, xd, cext, sv_of_block, tw, n_cells, n_dop, dmax, wlen, cells, red_best, red_sum, red_idx);
        if (!more) break;
        s = (s + 1 == G::N1) ? 0 : s + 1;
        ori = (ori + 1 == 3) ? 0 : ori + 1;
    }
}


}  // namespace ga
