// ga_pfa.cuh -- GRID mode kernels on the native W-point prime-factor transforms of ga_pfa.h
// (W = 5456, 8184, 2800: the 1 ms block lengths of BASELINE.json configs[1..4]).  Same semantics as
// ga_grid.cuh (the reference's Correlate(), c/search_offline.cpp:169-201, is the
// template for the conj side and the |.|^2 / first-max / sum statistics) but no zero-padded embedding:
// a cell reads exactly the 2*W*8 algorithmic bytes and does one W-point backward transform per block.
#pragma once
#include "ga_grid.cuh"
#include "ga_pfa.h"

// Task loops of pfa_cell_kernel's passes B and C fully unrolled (pass A stays a rolled loop: its global operand loads would be
// hoisted and spill).  Measured, ticket scheduler, M corr/s C1 / C2 / C3 / C4: rolled 61.6 / 38.8 / 171.9 / 43.4,
// pass C unrolled 62.8 / 40.1 / 182.2 / 43.7, passes B and C 63.2 / 39.0 / 184.8 / 42.5 -- so pass C everywhere and pass B
// except for W = 8184 (its 9 radix-31 tasks over 4 warps).
#ifndef PFA_UNROLL_B
#define PFA_UNROLL_B 1
#endif
#ifndef PFA_UNROLL_C
#define PFA_UNROLL_C 1
#endif
#ifndef PFA_PIPE_A
#define PFA_PIPE_A 0     // measured: software-pipelined pass A is 1-4 % SLOWER (the SM already overlaps the loads across its 12-16 warps)
#endif

namespace ga {

// time sample n of (block, Doppler bin d), 32-bit index arithmetic (|d|*W < 2^31 is checked at create)
struct GridSrc32 {
    const unsigned char *chunk, *lo;
    const cf *wipe;
    int d, m;
    __device__ __forceinline__ cf operator()(int n) const
    {
        const int bit = (chunk[n >> 3] >> (n & 7)) & 1, l = lo[n];
        const float xr = (bit ^ (l & 1)) ? -1.0f : 1.0f, xi = (bit ^ (l >> 1)) ? -1.0f : 1.0f;
        int k = (d * n) % m;
        if (k < 0) k += m;
        const cf ph = wipe[k];
        return mk(__fadd_rn(xr * ph.x, -(xi * ph.y)), __fadd_rn(xr * ph.y, xi * ph.x));      // as GridSrc
    }
};

// forward transform of every (block, Doppler bin) [MODE 0: conj(X)] or of the 32 one-period replicas
// [MODE 1: C]; one CTA per item, output in the cell's (a,b,c)-linear order.  F = G::Fwd.
template <class G, int T, int MODE>
__global__ void __launch_bounds__(T) pfa_fwd_kernel(const unsigned char *__restrict__ bits, int block_bytes,
                                                    const unsigned char *__restrict__ lo, const cf *__restrict__ wipe,
                                                    int n_dop, int dmax, int m, const float *__restrict__ code_w,
                                                    cf *__restrict__ out)
{
    typedef typename G::Fwd F;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    const int item = blockIdx.x;
    if (MODE == 0) {
        const int blk = item / n_dop, di = item - blk * n_dop;
        GridSrc32 src{bits + (size_t)blk * block_bytes, lo, wipe, di - dmax, m};
        for (int j = threadIdx.x; j < F::NA; j += T) pfa_fwd_passA<F>(j, src, sm);
    } else {
        RealSrc src{code_w + (size_t)item * G::W};
        for (int j = threadIdx.x; j < F::NA; j += T) pfa_fwd_passA<F>(j, src, sm);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < F::NB; j += T) pfa_passB<F, -1>(j, sm);
    __syncthreads();
    cf *dst = out + (size_t)item * G::W;
    for (int j = threadIdx.x; j < F::NC; j += T) pfa_fwd_passC_store<F, MODE == 0>(j, sm, dst);
}

// rotated copies of the 32 replica spectra: out[(q - q_min)*32 + prn][pos(a,b,c)] = C_prn[k(a,b,c) - q]
template <class G>
__global__ void pfa_rotate_replicas_kernel(const cf *__restrict__ crep, int q_min, cf *__restrict__ out)
{
    const int q = q_min + (int)blockIdx.y, prn = blockIdx.z;
    const PfaRot r = pfa_rotation<G>(-q);
    const cf *src = crep + (size_t)prn * G::W;
    cf *dst = out + ((size_t)blockIdx.y * 32 + prn) * G::W;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < G::W; m += gridDim.x * blockDim.x) {
        int a = m / G::NA;
        const int j = m - a * G::NA;
        a += r.da; if (a >= G::RA) a -= G::RA;
        dst[m] = src[a * G::NA + pfa_rot_col<G>(j, r)];
    }
}

// slow path of the K = 1 statistics: redo one pass-C butterfly (its inputs are still in shared memory)
// comparing lags on every tie.  Kept out of line: it is practically never executed.
// (running statistics go in and come back BY VALUE: reference parameters of a non-inlined function put best/besti/sum
// into a stack frame that every butterfly then stored to -- the "local memory" traffic ncu showed in round 1)
struct PfaAcc { float best; int besti; float sum; };
template <class G>
__device__ __noinline__ PfaAcc pfa_passC_exact(int jc, const cf *sm, int t0, float best, int besti, float sum)
{
    PfaPeakExact<G> pk;
    pk.init(t0);
    pfa_passC<G, +1>(jc, sm, [&](auto wc, cf v) { pk.template put<decltype(wc)::value>(fmaf(v.x, v.x, v.y * v.y)); });
    pk.merge(best, besti, sum);
    PfaAcc r; r.best = best; r.besti = besti; r.sum = sum;
    return r;
}

// The native GRID cell kernel.  cell = (acquisition, Doppler bin, PRN), PRN fastest (the 32 cells that
// share a block spectrum run together).  Per 1 ms block: pass A (coalesced loads of conj(X) and C,
// multiply, radix-RA) -> smem, pass B in place, pass C -> |y|^2.  K = 1: statistics straight from the
// registers.  K > 1 (MULTI): the power of each lag is summed over the K blocks in tensor memory
// (tcgen05.ld/st, as in cell_kernel_tm) and the statistics are taken on the sum.
template <class G, int T, int MINB, bool MULTI>
__global__ void __launch_bounds__(T, MINB) pfa_cell_kernel(const cf *__restrict__ xg, const cf *__restrict__ crep,
                                                           int n_cells, int n_dop, int dmax, int n_base, int q_min, int kblocks,
                                                           CellStat *__restrict__ cells, int *__restrict__ sched = nullptr)
{
    static_assert(T % 32 == 0, "whole warps only");
    constexpr int NWARP = T / 32;
    constexpr int NTA = cdiv(G::NA, 32), NTB = cdiv(G::NB, 32), NTC = cdiv(G::NC, 32);
    constexpr int ITA = cdiv(NTA, NWARP), ITB = cdiv(NTB, NWARP), ITC = cdiv(NTC, NWARP);
    constexpr bool PIPE_A = PFA_PIPE_A != 0;
    constexpr int UNROLL_B = (PFA_UNROLL_B && G::W != 8184) ? ITB : 1, UNROLL_C = PFA_UNROLL_C ? ITC : 1;
    constexpr int NWP = (G::RC + 1) & ~1;                            // power accumulators per butterfly, even
    constexpr uint32_t COL_SLOT = (uint32_t)((ITC * NWP + 7) & ~7);
    constexpr uint32_t TM_COLS = pow2_at_least(COL_SLOT * cdiv(NWARP, 4));
    static_assert(!MULTI || TM_COLS * MINB <= 512, "TMEM columns");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    __shared__ float red_best[NWARP], red_sum[NWARP];
    __shared__ int red_idx[NWARP];
    __shared__ uint32_t tm_base_s;
    __shared__ int next_cell_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    uint32_t tm_base = 0, tm_mine = 0;
    if (MULTI) {
        if (wid == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_base_s)), "r"(TM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tm_base = tm_base_s;
        tm_mine = tm_base + ((32u * (uint32_t)(wid & 3)) << 16) + (uint32_t)(wid >> 2) * COL_SLOT;
    }

    // cells in ascending order from a device-wide ticket counter, as in cell_kernel_tm (ga_kernels.cuh): the 32 PRN cells
    // that share a block spectrum run at the same time on neighbouring CTAs, and a launch ends with one cell of imbalance
    int next_cell = 0;
    for (int cell = blockIdx.x; cell < n_cells; cell = next_cell) {
        const int acq = cell / (n_dop * 32), r = cell - acq * (n_dop * 32);
        const int di = r >> 5, prn = r & 31;
        const cf *cs = crep + (size_t)prn * G::W;
        float best = 0.0f, sum = 0.0f;
        int besti = 0;
        // n_base = R > 0: only the bins 0..R-1 of each block were transformed (Doppler bins R apart are one DFT bin
        // apart): bin d = r + R*q reads block spectrum r and the replica spectrum rotated by -q (ga_pfa.h);
        // n_base = 0: one block spectrum per bin, unrotated replicas
        int xsel = di, xstride = n_dop;
        if (n_base > 0) {
            const int d = di - dmax;
            int rr = d % n_base;
            if (rr < 0) rr += n_base;
            xsel = rr; xstride = n_base;
            cs = crep + ((size_t)((d - rr) / n_base - q_min) * 32 + prn) * G::W;
        }

        for (int k = 0; k < kblocks; k++) {
            const cf *xs = xg + ((size_t)(acq * kblocks + k) * xstride + xsel) * G::W;
            if (k == kblocks - 1 && tid == 0) next_cell = sched ? (int)gridDim.x + atomicAdd(sched, 1) : cell + (int)gridDim.x;
            if (PIPE_A && ITA > 1) {
                // software-pipelined pass A: the operand rows of the warp's NEXT task are in flight (registers)
                // while the current task multiplies and runs its butterfly -- one exposed L2 round trip per
                // pass instead of one per task
                cf xa[G::RA], ca[G::RA], xb[G::RA], cb[G::RA];
                const int j0 = wid * 32 + lane;
                if (j0 < G::NA) pfa_cell_loadA<G>(j0, xs, cs, xa, ca);
#pragma unroll 1
                for (int it = 0; it < ITA; it += 2) {
                    const int ja = (wid + it * NWARP) * 32 + lane, jb = ja + NWARP * 32, jn = jb + NWARP * 32;
                    if (it + 1 < ITA && jb < G::NA) pfa_cell_loadA<G>(jb, xs, cs, xb, cb);
                    if (ja < G::NA) pfa_cell_passA_regs<G>(ja, xa, ca, sm);
                    if (it + 1 < ITA) {
                        if (it + 2 < ITA && jn < G::NA) pfa_cell_loadA<G>(jn, xs, cs, xa, ca);
                        if (jb < G::NA) pfa_cell_passA_regs<G>(jb, xb, cb, sm);
                    }
                }
            } else {
#pragma unroll 1
                for (int it = 0; it < ITA; it++) {
                    const int j = (wid + it * NWARP) * 32 + lane;
                    if (j < G::NA) pfa_cell_passA<G>(j, xs, cs, sm);
                }
            }
            __syncthreads();
#pragma unroll UNROLL_B
            for (int it = 0; it < ITB; it++) {
                const int j = (wid + it * NWARP) * 32 + lane;
                if (j < G::NB) pfa_passB<G, +1>(j, sm);
            }
            __syncthreads();
#pragma unroll UNROLL_C
            for (int it = 0; it < ITC; it++) {
                const int task = wid + it * NWARP;
                if (task < NTC) {                                   // warp-uniform
                    const int j = task * 32 + lane;
                    const bool act = j < G::NC;
                    const int jc = act ? j : G::NC - 1;             // idle lanes shadow the last butterfly (TMEM ops are warp-wide)
                    const int t0 = pfa_passC_t0<G>(jc);
                    if (!MULTI) {
                        // K = 1: power and statistics straight from the butterfly's output stream
                        PfaPeak<G> pk;
                        pk.init(t0);
                        pfa_passC<G, +1>(jc, sm, [&](auto wc, cf v) { pk.template put<decltype(wc)::value>(fmaf(v.x, v.x, v.y * v.y)); });
                        if (act) {
                            if (!pk.tie) pk.merge(best, besti, sum);
                            else { const PfaAcc r = pfa_passC_exact<G>(jc, sm, t0, best, besti, sum); best = r.best; besti = r.besti; sum = r.sum; }
                        }
                    } else {
                        float pw[NWP];
                        if (NWP > G::RC) pw[NWP - 1] = 0.0f;
                        pfa_passC<G, +1>(jc, sm, [&](auto wc, cf v) { pw[decltype(wc)::value] = fmaf(v.x, v.x, v.y * v.y); });
                        const uint32_t pcol = tm_mine + (uint32_t)(it * NWP);
                        if (k > 0) {
                            float old[NWP];
                            tm_move<NWP, true>(pcol, old);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int w = 0; w < G::RC; w++) pw[w] = old[w] + pw[w];       // blocks in ascending order (App. E)
                        }
                        if (k < kblocks - 1) {
                            tm_move<NWP, false>(pcol, pw);
                        } else if (act) {
                            PfaPeak<G> pk;
                            pk.init(t0);
                            static_for<0, G::RC>([&](auto wc) { pk.template put<decltype(wc)::value>(pw[decltype(wc)::value]); });
                            if (!pk.tie) pk.merge(best, besti, sum);
                            else {
                                PfaPeakExact<G> pe;
                                pe.init(t0);
                                static_for<0, G::RC>([&](auto wc) { pe.template put<decltype(wc)::value>(pw[decltype(wc)::value]); });
                                pe.merge(best, besti, sum);
                            }
                        }
                    }
                }
            }
            if (k == kblocks - 1) {
                // the cell is complete: warp-shuffle reduction (ties to the lower lag = first maximum, :192),
                // per-warp partials to smem BEFORE the barrier that also frees smem for the next pass A
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ob = __shfl_down_sync(0xffffffffu, best, off);
                    const int oi = __shfl_down_sync(0xffffffffu, besti, off);
                    const float os = __shfl_down_sync(0xffffffffu, sum, off);
                    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
                    sum += os;
                }
                if (lane == 0) { red_best[wid] = best; red_idx[wid] = besti; red_sum[wid] = sum; }
                if (tid == 0) next_cell_s = next_cell;
            }
            if (MULTI) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            __syncthreads();                                        // smem is rewritten by the next pass A
        }
        next_cell = next_cell_s;        // thread 0 rewrites it only after two more block barriers
        if (wid == 0) {
            // warp 0 finishes the record while the other warps start the next cell (they touch red_* again
            // only after the next cell's barriers, which warp 0 takes part in)
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
                CellStat rec; rec.max_pwr = best; rec.tot_pwr = sum; rec.max_idx = besti; rec.pad = 0;
                cells[((size_t)acq * 32 + prn) * n_dop + di] = rec;
            }
        }
        // red_* are rewritten only after the next cell's barriers
    }
    if (MULTI) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
    }
    if (sched && tid == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }     // last CTA out rewinds
}

}  // namespace ga
