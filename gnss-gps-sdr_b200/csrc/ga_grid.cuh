// ga_grid.cuh -- GRID mode kernels: 1 ms coherent blocks, explicit Doppler grid with time-domain
// carrier wipe-off, K-block non-coherent summation (BASELINE.json configs[1..4]; semantics defined in
// SURVEY.md App. E -- the reference has no such mode, its Correlate() (c/search_offline.cpp:169-201) is
// the template: same conj side, same |.|^2 / first-max / sum / snr statistics).
//
// A 1 ms block has W = FS/1000 samples (5456 = 2^4*11*31, 8184 = 2^3*3*11*31, 2800 = 2^4*5^2*7).  The
// W-point circular correlation is evaluated EXACTLY (not approximately) through the transform family
// of ga_fft3.h by linear-correlation embedding: the block is zero-padded to L = 2*N2 >= 2W, the replica
// holds two code periods (c[n mod W], n < 2W) and zeros, so lags 0..W-1 of the L-point circular
// correlation equal the W-point circular ones.  L = 16000 / 20000 / 8000 = N1*N2 with N1 = 2.
#pragma once
#include "ga_kernels.cuh"

namespace ga {

// time sample n of (block, Doppler bin d): XOR mix like Sample() (:143-153) times the wipe-off phasor
// exp(-j*2*pi*d*step*n/FS) = wipe[(d*n) mod M], M = FS/step (an integer); zero for n >= W.
struct GridSrc {
    const unsigned char *chunk, *lo;
    const cf *wipe;
    int w, d, m;
    __device__ __forceinline__ cf operator()(int n) const
    {
        if (n >= w) return mk(0.0f, 0.0f);
        const int bit = (chunk[n >> 3] >> (n & 7)) & 1, l = lo[n];
        const float xr = (bit ^ (l & 1)) ? -1.0f : 1.0f, xi = (bit ^ (l >> 1)) ? -1.0f : 1.0f;
        long long k = ((long long)d * n) % m;
        if (k < 0) k += m;
        const cf ph = wipe[k];
        // (xr + j*xi)*(pr + j*pi) with xr, xi = +-1: products exact, one rounding per component
        return mk(__fadd_rn(xr * ph.x, -(xi * ph.y)), __fadd_rn(xr * ph.y, xi * ph.x));
    }
};

// forward transform of every (block, Doppler bin): grid = n_blocks * n_dop * N1 CTAs; output conj(X),
// decimated: xg[(block*n_dop + di)][s][q]
template <class G, int T, int GID>
__global__ void __launch_bounds__(T) fwd_grid_kernel(const unsigned char *__restrict__ bits, int block_bytes,
                                                     const unsigned char *__restrict__ lo, const cf *__restrict__ wipe,
                                                     int wlen, int n_dop, int dmax, int m, const cf *__restrict__ tw,
                                                     cf *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    const int item = blockIdx.x / G::N1, s = blockIdx.x - item * G::N1;
    const int blk = item / n_dop, di = item - blk * n_dop;
    const cf *k1s = c_k1tab[GID] + s * G::N1;
    GridSrc src{bits + (size_t)blk * block_bytes, lo, wipe, wlen, di - dmax, m};
    for (int j = threadIdx.x; j < G::NA; j += T) fwd_passA<G>(j, s, src, k1s, tw, sm);
    __syncthreads();
    for (int j = threadIdx.x; j < G::NB; j += T) passB<G, -1>(j, 0, tw, sm);
    __syncthreads();
    for (int j = threadIdx.x; j < G::NC; j += T) {
        cf p[G::RC];
        const int tau0 = passC<G, -1>(j, sm, p);
        cf *dst = out + ((size_t)item * G::N1 + s) * G::N2 + tau0;
#pragma unroll
        for (int w = 0; w < G::RC; w++) dst[G::OUT_STRIDE * w] = cconj(p[w]);
    }
}

// replica for GRID mode: two periods of the W-sample code, scaled by W/L so that powers carry the
// W-point unnormalised-FFT scale of the definition, zero up to L.  in: [32][w], out: [32][L]
__global__ void grid_replica_extend_kernel(const float *__restrict__ code_w, int w, int L, float scale, float *__restrict__ out)
{
    const int sv = blockIdx.y;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x)
        out[(size_t)sv * L + n] = (n < 2 * w) ? code_w[(size_t)sv * w + (n >= w ? n - w : n)] * scale : 0.0f;
}

// The GRID cell kernel.  cell = (acquisition a, Doppler bin di, PRN p), p fastest so that the 32 cells
// sharing one block spectrum run together.  For each of the K blocks of the acquisition: N1
// sub-sequence transforms accumulate y in tensor memory; |y|^2 is added to a power accumulator (also
// TMEM) across blocks; the statistics are taken on the summed power.
template <class G, int T, int NW, int GID>
__global__ void __launch_bounds__(T, TM_MINB) grid_cell_kernel(const cf *__restrict__ xg, const cf *__restrict__ cext,
                                                               const cf *__restrict__ tw, int n_cells, int n_dop, int kblocks,
                                                               int wlen, int dmax, int n_base, CellStat *__restrict__ cells,
                                                               int *__restrict__ sched = nullptr)
{
    static_assert(T % 32 == 0, "tcgen05.ld/st are warp-collective: whole warps only");
    constexpr int NWARP = T / 32;
    constexpr int NTA = cdiv(G::NA, 32), NTB = cdiv(G::NB, 32), NTC = cdiv(G::NC, 32);
    constexpr int ITA = cdiv(NTA, NWARP), ITB = cdiv(NTB, NWARP), ITC = cdiv(NTC, NWARP);
    constexpr int NWP = NW + (NW & 1);                             // power accumulators, padded to an even count
    constexpr uint32_t COLS_Y = ITC * 2 * NW, COLS_P = ITC * NWP;
    constexpr uint32_t COL_SLOT = (COLS_Y + COLS_P + 7u) & ~7u;
    constexpr uint32_t TM_COLS = pow2_at_least(COL_SLOT * cdiv(NWARP, 4));
#ifndef GA_NO_TM_ASSERT
    static_assert(TM_COLS * TM_MINB <= 512 || G::SMEM_ELEMS * sizeof(cf) * TM_MINB > 227 * 1024, "TMEM columns");
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    __shared__ float red_best[NWARP], red_sum[NWARP];
    __shared__ int red_idx[NWARP];
    __shared__ uint32_t tm_base_s;
    __shared__ int next_cell_s;
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

    // cells in ascending order from the device-wide ticket counter, as in cell_kernel_tm / pfa_cell_kernel
    int next_cell = 0;
    for (int cell = blockIdx.x; cell < n_cells; cell = next_cell) {
        const int acq = cell / (n_dop * 32), r = cell - acq * (n_dop * 32);
        const int di = r >> 5, prn = r & 31;
        const cf *cb = cext + (size_t)prn * (2 * G::N);
        float best = 0.0f, sum = 0.0f;
        int besti = 0;
        // exact-length transforms (N1 = 1, L = W) with n_base = R > 0: only the bins 0..R-1 of each block were
        // transformed; bin d = r + R*q reads block spectrum r and the replica spectrum rotated by -q, which in the
        // doubled natural-order layout is a pointer offset (see ga_pfa.h for why the powers are the same)
        int xsel = di, xstride = n_dop;
        if (G::N1 == 1 && n_base > 0) {
            const int d = di - dmax;
            int rr = d % n_base;
            if (rr < 0) rr += n_base;
            int off = -((d - rr) / n_base) % G::N2;
            if (off < 0) off += G::N2;
            xsel = rr; xstride = n_base;
            cb += off;
        }

        for (int k = 0; k < kblocks; k++) {
            const cf *xb = xg + ((size_t)(acq * kblocks + k) * xstride + xsel) * G::N;
            if (k == kblocks - 1 && tid == 0) next_cell = sched ? (int)gridDim.x + atomicAdd(sched, 1) : cell + (int)gridDim.x;
            for (int s = 0; s < G::N1; s++) {
                const cf *xs = xb + (size_t)s * G::N2;
                const cf *cs = cb + (size_t)s * (2 * G::N2);
#pragma unroll
                for (int it = 0; it < ITA; it++) {
                    const int j = (wid + it * NWARP) * 32 + lane;
                    if (j < G::NA) cell_passA<G>(j, s, xs, cs, tw, sm);
                }
                __syncthreads();
#pragma unroll
                for (int it = 0; it < ITB; it++) {
                    const int j = (wid + it * NWARP) * 32 + lane;
                    if (j < G::NB) passB<G, +1>(j, s, tw, sm);
                }
                __syncthreads();
                const cf *ks = c_ktab[GID] + s * G::RC;
#pragma unroll
                for (int it = 0; it < ITC; it++) {
                    const int task = wid + it * NWARP;
                    if (task < NTC) {          // warp-uniform
                        const int j = task * 32 + lane;
                        const bool act = j < G::NC;
                        const int jc = act ? j : G::NC - 1;
                        cf p[G::RC];
                        const int tau0 = passC<G, +1>(jc, sm, p);
                        float a[2 * NW];
                        const uint32_t col = tm_mine + (uint32_t)(it * 2 * NW);
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
                        if (s < G::N1 - 1) {
                            tm_move<2 * NW, false>(col, a);
                        } else {
                            // y of this block is complete: power, non-coherent sum over the K blocks
                            float pw[NWP];
                            const uint32_t pcol = tm_mine + COLS_Y + (uint32_t)(it * NWP);
#pragma unroll
                            for (int w = 0; w < NW; w++) pw[w] = fmaf(a[2 * w], a[2 * w], a[2 * w + 1] * a[2 * w + 1]);
                            if (NW & 1) pw[NW] = 0.0f;
                            if (k > 0) {
                                float old[NWP];
                                tm_move<NWP, true>(pcol, old);
                                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                                for (int w = 0; w < NW; w++) pw[w] = old[w] + pw[w];      // b ascending (App. E)
                            }
                            if (k < kblocks - 1) {
                                tm_move<NWP, false>(pcol, pw);
                            } else if (act) {
#pragma unroll
                                for (int w = 0; w < NW; w++) {
                                    const int tau = tau0 + G::OUT_STRIDE * w;
                                    if (tau < wlen) {
                                        if (pw[w] > best || (pw[w] == best && tau < besti)) { best = pw[w]; besti = tau; }
                                        sum += pw[w];
                                    }
                                }
                            }
                        }
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                __syncthreads();
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
        if (tid == 0) next_cell_s = next_cell;
        __syncthreads();
        next_cell = next_cell_s;          // thread 0 rewrites it only after the barriers of the next cell
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
                CellStat rec; rec.max_pwr = best; rec.tot_pwr = sum; rec.max_idx = besti; rec.pad = 0;
                cells[((size_t)acq * 32 + prn) * n_dop + di] = rec;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
    if (sched && tid == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }     // last CTA out rewinds
}

}  // namespace ga
