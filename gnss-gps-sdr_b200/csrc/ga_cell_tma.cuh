// ga_cell_tma.cuh -- the REF hot kernel (Correlate() inner loop, c/search_offline.cpp:181-194) with its operands staged
// by the Tensor Memory Accelerator.
//
// Same arithmetic, same summation order and same records as cell_kernel_tm (ga_kernels.cuh); what changes is how the two
// operand streams of pass A -- the block spectrum sub-sequence conj(X)_s and the rotated replica sub-sequence C_sp of
// c/search_offline.cpp:181-185 -- reach the butterflies:
//
//   * one PRODUCER warp per CTA walks the CTA's work (cell -> sub-sequence -> pass-A task) ahead of the math and issues,
//     per task of 32 butterflies, two cp.async.bulk.tensor.2d loads (SASS UTMALDG): a {32 columns x RA rows} box of the
//     block spectrum and the same box of the replica spectrum, into a ring of NSTAGE shared-memory stages; completion
//     is signalled (complete_tx::bytes) on an mbarrier per TASK NUMBER, not per stage: the consumer warps wait for
//     different tasks at the same time, and with a barrier per stage a warp waiting for the stage's NEXT lap would read
//     the one-bit phase parity of the lap before as "done".  A task's barrier completes once per sub-sequence, and nobody
//     waits for sub-sequence i before every task of i-1 has been consumed (block barriers in between);
//   * NCW CONSUMER warps wait on the stage's "full" barrier, read their 2 x RA operands from shared memory (conflict-free:
//     lane = column), hand the stage back through its "empty" barrier and run multiply + radix-RA butterfly + twiddle
//     into the tile; passes B and C, the tensor-memory accumulators and the statistics are those of cell_kernel_tm.
//     The consumer warps synchronise among themselves with a named barrier (bar.sync 1, 32*NCW); the producer never
//     joins it, so it runs up to NSTAGE tasks ahead across pass, sub-sequence and cell boundaries: the L2 latency of
//     the operand fetch is off the butterflies' critical path and costs no issue slots or registers.
//
// Replica spectra for this kernel live in a HALO layout (replica_halo_kernel): sub-sequence (sv, sp) is an RA x NA
// matrix (row a = elements a*NA .. a*NA+NA-1, exactly the pass-A view) whose rows are extended by H elements on both
// sides, circularly.  The Doppler rotation (i - dop) mod N of :182 is then a column offset |e| <= H of the TMA box --
// no modulo, one descriptor for all cells.  A box must start on a 16-byte boundary (an odd column of 8-byte elements
// raises "illegal instruction"), so the replica box is 34 columns wide, starts at the even column below and the
// consumers skip its first column when the offset is odd.
#pragma once
#include <cuda.h>
#include "ga_kernels.cuh"

namespace ga {

// {c0 (column), c1 (row)} box of a 2-D tensor -> shared memory, completion bytes on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
template <int NTHREADS> __device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

// halo[(sv*N1 + sp)*RA + a][i] = C_sp[(a*NA + i - H) mod N2],  0 <= i < NA + 2H   (cext is the doubled layout: [sv][sp][2*N2])
template <class G>
__global__ void replica_halo_kernel(const cf *__restrict__ cext, int halo, cf *__restrict__ out)
{
    const int row = blockIdx.x, pitch = G::NA + 2 * halo;
    const int a = row % G::RA, sub = row / G::RA;                  // sub = sv*N1 + sp
    const cf *src = cext + (size_t)sub * (2 * G::N2);
    for (int i = threadIdx.x; i < pitch; i += blockDim.x) {
        int t = a * G::NA + i - halo;
        t %= G::N2; if (t < 0) t += G::N2;
        out[(size_t)row * pitch + i] = src[t];
    }
}

#ifndef GA_TMA_XLDG
#define GA_TMA_XLDG 0                      // 1: only the replica operand is staged by TMA, the block spectrum is read with __ldg
#endif
constexpr bool TMA_XLDG = GA_TMA_XLDG != 0;
constexpr int TMA_NSTAGE = TMA_XLDG ? 8 : 4;
constexpr int TMA_BOX_COLS = 32;           // block-spectrum box: one column per lane
constexpr int TMA_BOX_COLS_C = 34;         // replica box: + the alignment column + one more to keep rows 16-byte multiples

template <class G> struct TmaCellShape {
    static constexpr int X_ELEMS = TMA_XLDG ? 0 : G::RA * TMA_BOX_COLS, C_ELEMS = G::RA * TMA_BOX_COLS_C;
    static constexpr int TX_BYTES = (X_ELEMS + C_ELEMS) * (int)sizeof(cf);      // bytes the two loads of a stage deliver
    static constexpr int STAGE_BYTES = (TX_BYTES + 127) / 128 * 128;            // TMA destinations are 128-byte aligned
    static constexpr int STAGE_ELEMS = STAGE_BYTES / (int)sizeof(cf);
    static constexpr int TILE_BYTES = ((G::SMEM_ELEMS * (int)sizeof(cf) + 1023) / 1024) * 1024;   // ring starts 1 KB aligned
    static constexpr int SMEM_BYTES = TILE_BYTES + TMA_NSTAGE * STAGE_BYTES + 1024;            // + slack to align the base
};

// NCW consumer warps + 1 producer warp; 2 CTAs per SM.
template <class G, int NCW, int NW, int GID>
__global__ void __launch_bounds__(32 * (NCW + 1), 2)
cell_kernel_tma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c,
                const int *__restrict__ sv_of_block, const cf *__restrict__ tw,
                int n_cells, int n_dop, int dmax, int wlen, int halo, int xrow0, CellStat *__restrict__ cells, int blk0,
                int *__restrict__ sched, const cf *__restrict__ xd = nullptr)
{
    constexpr int NCT = 32 * NCW;                                   // consumer threads
    constexpr int NTA = cdiv(G::NA, 32), NTB = cdiv(G::NB, 32), NTC = cdiv(G::NC, 32);
    constexpr int ITA = cdiv(NTA, NCW), ITB = cdiv(NTB, NCW), ITC = cdiv(NTC, NCW);
    constexpr uint32_t COLS_THREAD = ITC * 2 * NW;
    constexpr uint32_t COL_SLOT = (COLS_THREAD + 7u) & ~7u;
    constexpr uint32_t TM_COLS = pow2_at_least(COL_SLOT * cdiv(NCW, 4));
    static_assert(TM_COLS * 2 <= 512, "two CTAs per SM must fit in the 512 TMEM columns");
    static_assert(G::N1 >= 2, "the ticket is drawn in sub-sequence 0 and published in sub-sequence 1");
    using Shape = TmaCellShape<G>;

    extern __shared__ unsigned char smem_dyn[];
    // dynamic shared memory starts 16-byte aligned at least; the TMA destinations want 128: align the base by hand
    unsigned char *smem_raw = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    cf *ring = reinterpret_cast<cf *>(smem_raw + Shape::TILE_BYTES);
    __shared__ __align__(8) unsigned long long bar_full[cdiv(G::NA, 32)], bar_empty[TMA_NSTAGE], bar_tick;
    __shared__ float red_best[NCW], red_sum[NCW];
    __shared__ int red_idx[NCW];
    __shared__ uint32_t tm_base_s;
    __shared__ int next_cell_s[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 32 * NCW) {            // (a lane of the producer warp: warp 0 must reach the warp-collective tcgen05.alloc converged)
        for (int i = 0; i < NTA; i++) mbar_init(smem_u32(&bar_full[i]), 1);
        for (int i = 0; i < TMA_NSTAGE; i++) mbar_init(smem_u32(&bar_empty[i]), 1);
        mbar_init(smem_u32(&bar_tick), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_base_s)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (wid == NCW) {
        // ------------------------------------------------------------------ producer: one lane feeds the ring
        if (lane == 0) {
#ifndef GA_TMA_NO_PREFETCH
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
#endif
            uint32_t q = 0, n = 0, sub = 0;                          // stage, cell and sub-sequence sequence numbers
            for (int cell = blockIdx.x; cell < n_cells; n++) {
                const int blk = cell / n_dop, dop = cell - blk * n_dop - dmax;
                const int sv = (sv_of_block ? sv_of_block[blk] : blk + blk0) & 31;
                for (int s = 0; s < G::N1; s++, sub++) {
                    const int v = s - dop;
                    int e = v / G::N1, sp = v - e * G::N1;
                    if (sp < 0) { sp += G::N1; e -= 1; }            // floor division: s - dop = N1*e + sp, |e| <= halo
                    const int xrow = xrow0 + (blk * G::N1 + s) * G::RA, crow = (sv * G::N1 + sp) * G::RA, ccol = (halo + e) & ~1;
                    for (int k = 0; k < NTA; k++, q++) {
                        // the last task's boxes are pulled back inside the row (no out-of-bounds fill): columns NA-32..NA-1
                        const int col = k * TMA_BOX_COLS < G::NA - TMA_BOX_COLS ? k * TMA_BOX_COLS : G::NA - TMA_BOX_COLS;
                        const uint32_t st = q % TMA_NSTAGE, ph = (q / TMA_NSTAGE) & 1;
                        mbar_wait(smem_u32(&bar_empty[st]), ph ^ 1);         // first lap passes at once
                        const uint32_t full = smem_u32(&bar_full[k]), dst = smem_u32(ring + (size_t)st * Shape::STAGE_ELEMS);
                        mbar_arrive_expect_tx(full, Shape::TX_BYTES);
                        if (!TMA_XLDG) tma_load_2d(dst, &map_x, full, col, xrow);
                        tma_load_2d(dst + Shape::X_ELEMS * (uint32_t)sizeof(cf), &map_c, full, ccol + col, crow);
                    }
                }
                mbar_wait(smem_u32(&bar_tick), n & 1);               // the consumers' ticket for the next cell
                cell = next_cell_s[n & 1];
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const uint32_t tm_base = tm_base_s;
    const uint32_t tm_mine = tm_base + ((32u * (uint32_t)(wid & 3)) << 16) + (uint32_t)(wid >> 2) * COL_SLOT;
    const int vw = wid;
    uint32_t q0 = 0, n = 0, sub = 0;                                 // stage sequence number of task 0 of the current pass A; cell, sub-sequence
    int ticket = 0;
    for (int cell = blockIdx.x; cell < n_cells; n++) {
        float best = 0.0f, sum = 0.0f;
        int besti = 0;
        const int blk = cell / n_dop, dop = cell - blk * n_dop - dmax;
        for (int s = 0; s < G::N1; s++, sub++) {
            const cf *xg = xd + ((size_t)blk * G::N1 + s) * G::N2;      // (TMA_XLDG) conj(X) sub-sequence s of this chunk
            // the replica box starts at the even column at or below halo + e, e = floor((s - dop) / N1)
            const int odd = (halo + (s - dop + G::N1 * (halo + 1)) / G::N1 - (halo + 1)) & 1;
            // the next cell's ticket: drawn at the very start of a cell, published one sub-sequence later (the latency of
            // the atomic is hidden), read by the producer long before it has run out of tasks of this cell
            if (tid == 0) {
                if (s == 0) ticket = (int)gridDim.x + atomicAdd(sched, 1);
                if (s == 1) { next_cell_s[n & 1] = ticket; mbar_arrive(smem_u32(&bar_tick)); }
            }
            for_tasks<ITA>([&](int it) {
                const int task = vw + it * NCW;
                if (task < NTA) {                                    // warp-uniform
                    const uint32_t st = (q0 + (uint32_t)task) % TMA_NSTAGE;
                    mbar_wait(smem_u32(&bar_full[task]), sub & 1);
                    const cf *xs = ring + (size_t)st * Shape::STAGE_ELEMS + lane, *cs = xs + Shape::X_ELEMS + odd;
                    // butterfly of this lane: column task*32 + lane, except in the last task, whose box covers the LAST 32
                    // columns of the row (its leading lanes repeat columns of the task before and stay idle)
                    const int col0 = task * 32 < G::NA - 32 ? task * 32 : G::NA - 32, j = col0 + lane;
                    cf p[G::RA];
#pragma unroll
                    for (int a = 0; a < G::RA; a++)
                        p[a] = cmul(TMA_XLDG ? ldg(xg + a * G::NA + j) : xs[a * TMA_BOX_COLS], cs[a * TMA_BOX_COLS_C]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_empty[st]));   // operands are in registers: the stage is free
                    if (j >= task * 32) passA_finish<G, +1, 0>(p, j, s, tw, sm);
                }
            });
            q0 += NTA;
            consumer_sync<NCT>();
            for_tasks<ITB>([&](int it) {
                const int j = (vw + it * NCW) * 32 + lane;
                if (j < G::NB) passB<G, +1>(j, s, tw, sm);
            });
            consumer_sync<NCT>();
            const cf *ks = c_ktab[GID] + s * G::RC;
            for_tasks<ITC>([&](int it) {
                const int task = vw + it * NCW;
                if (task < NTC) {
                    const int j = task * 32 + lane;
                    const bool act = j < G::NC;
                    const int jc = act ? j : G::NC - 1;
                    cf p[G::RC];
                    const int tau0 = passC<G, +1>(jc, sm, p);
                    const uint32_t col = tm_mine + (uint32_t)(it * 2 * NW);
                    auto accumulate = [&](float (&a)[2 * NW]) {
                        if (s == 0) {
#pragma unroll
                            for (int w = 0; w < NW; w++) { a[2 * w] = p[w].x; a[2 * w + 1] = p[w].y; }
                        } else {
                            tm_move<2 * NW, true>(col, a);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int w = 0; w < NW; w++) {
                                cf t = mk(a[2 * w], a[2 * w + 1]);
                                cfma(t, p[w], ks[w]);
                                a[2 * w] = t.x; a[2 * w + 1] = t.y;
                            }
                        }
                    };
                    if (s < G::N1 - 1) {
                        float a[2 * NW];
                        accumulate(a);
                        tm_move<2 * NW, false>(col, a);
                    } else {
                        float a[2 * NW];
                        accumulate(a);
                        if (act) {
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
                    }
                }
            });
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            consumer_sync<NCT>();      // the tile is rewritten by the next sub-sequence's pass A
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
        consumer_sync<NCT>();
        const int next_cell = next_cell_s[n & 1];         // entry n&1 is rewritten two cells from now
        if (wid == 0) {
            best = lane < NCW ? red_best[lane] : 0.0f;
            besti = lane < NCW ? red_idx[lane] : 0x7fffffff;
            sum = lane < NCW ? red_sum[lane] : 0.0f;
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
        cell = next_cell;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    consumer_sync<NCT>();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
    if (tid == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
}

}  // namespace ga
