// ga_kernels.cuh -- sm_100a kernels of the GPS L1 C/A acquisition engine.
//
//   replica_time_kernel   C/A code LFSR + code-NCO blend  (SearchInit(), c/search_offline.cpp:81-103)
//   fwd_kernel<BitSrc>    1-bit unpack + XOR mix + forward FFT (Sample(), :135-161)
//   fwd_kernel<RealSrc>   forward FFT of the replicas     (SearchInit(), :105-106)
//   cell_kernel           shifted conj-multiply + output-pruned backward FFT + |.|^2 +
//                         max/argmax/sum  (Correlate() inner loop, :181-194)  <-- the hot kernel
//   best_kernel           snr = max/(tot/W), best over Doppler (Correlate(), :196-200)
//
// Data layout in HBM: spectra are stored DECIMATED by N1: sub-sequence s
// holds X[N1*q+s], q < N2, contiguously.  Block spectra are stored conjugated
// (the product needs conj(data), :183-184); replica spectra are stored twice in
// a row (2*N2 per sub-sequence) so that the Doppler rotation (i-dop) mod N is a
// pointer offset.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ga_fft3.h"

namespace ga {

struct CellStat { float max_pwr, tot_pwr; int max_idx, pad; };            // == gpsacq_cell
struct Peak { float snr, max_pwr, tot_pwr; int lo_shift, ca_shift, sv, flags, reserved; };  // == gpsacq_peak

constexpr int KTAB_MAX = 256;      // N1*RC <= 250 for the geometries below
constexpr int K1TAB_MAX = 128;     // N1*N1 <= 100
constexpr int NGEOM = 12;      // 3 REF geometries (N = 40000) + 4 GRID embeddings (N1 = 2) + 5 exact-length GRID geometries (N1 = 1)
__constant__ cf c_ktab[NGEOM][KTAB_MAX];
__constant__ cf c_k1tab[NGEOM][K1TAB_MAX];

// ---------------------------------------------------------------------------------
// C/A replica in the time domain.  One CTA per PRN.  chip_idx/blendA/blendB are the
// code-NCO tables (functions of FS only) built on the host with the reference's
// float recurrence; taps are the G2 tap pair of the PRN (c/search_offline.cpp:20-53).
// ---------------------------------------------------------------------------------
struct SatTaps { unsigned char t0[32], t1[32]; };

__global__ void replica_time_kernel(SatTaps taps, const unsigned short *__restrict__ chip_idx,
                                    const float *__restrict__ blend_a, const float *__restrict__ blend_b,
                                    int n, float *__restrict__ out /* [32][n] */)
{
    __shared__ float chips[1024];
    const int sv = blockIdx.x;
    if (threadIdx.x == 0) {
        // G1 = x^10+x^3+1, G2 = x^10+x^9+x^8+x^6+x^3+x^2+1, all ones (c/cacode.h:15-28).
        // bit k-1 of g holds stage k; the new bit enters stage 1.
        unsigned g1 = 0x3FF, g2 = 0x3FF;
        const int t0 = taps.t0[sv], t1 = taps.t1[sv];
        for (int i = 0; i < 1023; i++) {
            const unsigned chip = ((g1 >> 9) ^ (g2 >> (t0 - 1)) ^ (g2 >> (t1 - 1))) & 1u;   // cacode.h:19-21
            chips[i] = chip ? -1.0f : 1.0f;                                              // Bipolar(), :68-70
            const unsigned f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
            const unsigned f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
            g1 = ((g1 << 1) | f1) & 0x3FF;
            g2 = ((g2 << 1) | f2) & 0x3FF;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int k = chip_idx[i];
        const float cur = chips[k], nxt = chips[k + 1 == 1023 ? 0 : k + 1];
        // chip*(1.0-ca_phase) rounded to float, + ca_phase*next  (:97-98); no FMA contraction
        out[(size_t)sv * n + i] = __fadd_rn(__fmul_rn(cur, blend_a[i]), __fmul_rn(blend_b[i], nxt));
    }
}

// ---------------------------------------------------------------------------------
// Forward transform, decimated output.  grid = n_items * N1 CTAs (item, s).
// ---------------------------------------------------------------------------------
struct BitSrc {       // Sample(): bit unpack LSB-first + XOR with the quadrature LO (:143-153)
    const unsigned char *chunk, *lo;      // lo[n] = lo_cos bit | lo_sin bit << 1 at sample n
    __device__ __forceinline__ cf operator()(int n) const
    {
        const int bit = (chunk[n >> 3] >> (n & 7)) & 1, l = lo[n];
        return mk((bit ^ (l & 1)) ? -1.0f : 1.0f, (bit ^ (l >> 1)) ? -1.0f : 1.0f);
    }
};
struct RealSrc {      // SearchInit(): real replica, imaginary part 0 (:101-102)
    const float *x;
    __device__ __forceinline__ cf operator()(int n) const { return mk(x[n], 0.0f); }
};

// MODE 0: blocks -> conj(X) into xd[item][s][q].   MODE 1: replicas -> cext[item][s][q] and [q+N2].
// MODE 0 stores through shared memory: a pass-C thread holds outputs q = tau0 + RA*RB*w, 160 bytes apart from its
// neighbour's -- written straight to HBM every 8-byte store would touch its own 32-byte sector.  Each thread puts
// its RC outputs back into the row it has just read (nobody else touches that row), and after a barrier the CTA
// streams the N2 values out in natural order with coalesced float2 stores (the read side walks shared memory with
// the odd stride Q: conflict-free).
#ifndef FWD_MINB
#define FWD_MINB 2     // measured: 2 CTAs/SM at 126 registers beat 3 at 80 (spills)
#endif
template <class G, int T, int MODE, int GID>
__global__ void __launch_bounds__(T, MODE == 0 ? FWD_MINB : 1) fwd_kernel(const unsigned char *__restrict__ bits, int chunk_bytes,
                                                const unsigned char *__restrict__ lo,
                                                const float *__restrict__ repl_time,
                                                const cf *__restrict__ tw, cf *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    const int item = blockIdx.x / G::N1, s = blockIdx.x - item * G::N1;
    const cf *k1s = c_k1tab[GID] + s * G::N1;

    if (MODE == 0) {
        BitSrc src{bits + (size_t)item * chunk_bytes, lo};
        for (int j = threadIdx.x; j < G::NA; j += T) fwd_passA<G>(j, s, src, k1s, tw, sm);
    } else {
        RealSrc src{repl_time + (size_t)item * G::N};
        for (int j = threadIdx.x; j < G::NA; j += T) fwd_passA<G>(j, s, src, k1s, tw, sm);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < G::NB; j += T) passB<G, -1>(j, 0, tw, sm);
    __syncthreads();
    if (MODE == 0) {
        for (int j = threadIdx.x; j < G::NC; j += T) {
            cf p[G::RC];
            passC<G, -1>(j, sm, p);
            const int u = j / G::RB, v = j - u * G::RB;
            cf *row = sm + G::template sa<0>() * u + G::template sb<0>() * v;
#pragma unroll
            for (int w = 0; w < G::RC; w++) row[w] = cconj(p[w]);
        }
        __syncthreads();
        cf *dst = out + ((size_t)item * G::N1 + s) * G::N2;
        for (int q = threadIdx.x; q < G::N2; q += T) {           // q = u + RA*v + RA*RB*w
            const int w = q / G::OUT_STRIDE, r = q - w * G::OUT_STRIDE;
            const int v = r / G::RA, u = r - v * G::RA;
            dst[q] = sm[G::template sa<0>() * u + G::template sb<0>() * v + w];
        }
    } else {
        for (int j = threadIdx.x; j < G::NC; j += T) {
            cf p[G::RC];
            const int tau0 = passC<G, -1>(j, sm, p);
            cf *dst = out + ((size_t)item * G::N1 + s) * (2 * G::N2) + tau0;
#pragma unroll
            for (int w = 0; w < G::RC; w++) { dst[G::OUT_STRIDE * w] = p[w]; dst[G::N2 + G::OUT_STRIDE * w] = p[w]; }
        }
    }
}

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
// Tensor-memory variant of the hot kernel: the per-thread output accumulators (NW complex
// values per butterfly, 56 floats per thread for 5 x 20^3) live in Blackwell TMEM instead of
// registers.  TMEM is 128 lanes x 512 columns x 32 bit per SM; with the 32x32b access shape
// thread i of warp w owns lane 32*(w%4)+i, so a column range is private per-thread storage:
// tcgen05.ld / tcgen05.st (SASS LDTM / STTM) move it to and from registers.  No tensor-core
// math is involved -- TMEM is used as a 2nd register file so that the radix-20 butterflies
// keep the whole 128-register budget (no local-memory spills, which the register version pays
// with ~150 M L2 write sectors per launch).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> struct TmVec;   // N 32-bit columns per thread
template <> struct TmVec<2> {
    static __device__ __forceinline__ void ld(uint32_t a, float *v)
    {
        uint32_t r0, r1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a) : "memory");
        v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
    }
    static __device__ __forceinline__ void st(uint32_t a, const float *v)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(a), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])) : "memory");
    }
};
template <> struct TmVec<4> {
    static __device__ __forceinline__ void ld(uint32_t a, float *v)
    {
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
        v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    }
    static __device__ __forceinline__ void st(uint32_t a, const float *v)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])),
                     "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
    }
};
template <> struct TmVec<8> {
    static __device__ __forceinline__ void ld(uint32_t a, float *v)
    {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
    }
    static __device__ __forceinline__ void st(uint32_t a, const float *v)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(__float_as_uint(v[0])),
                     "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                     "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
    }
};
// move NF floats (NF even) between registers and consecutive TMEM columns in x8/x4/x2 pieces
template <int NF, bool LOAD> __device__ __forceinline__ void tm_move(uint32_t a, float *v)
{
    if constexpr (NF >= 8) { if (LOAD) TmVec<8>::ld(a, v); else TmVec<8>::st(a, v); tm_move<NF - 8, LOAD>(a + 8, v + 8); }
    else if constexpr (NF >= 4) { if (LOAD) TmVec<4>::ld(a, v); else TmVec<4>::st(a, v); tm_move<NF - 4, LOAD>(a + 4, v + 4); }
    else if constexpr (NF >= 2) { if (LOAD) TmVec<2>::ld(a, v); else TmVec<2>::st(a, v); tm_move<NF - 2, LOAD>(a + 2, v + 2); }
}
constexpr uint32_t pow2_at_least(uint32_t x, uint32_t p = 32) { return p >= x ? p : pow2_at_least(x, p * 2); }
#ifndef TM_MINB
#define TM_MINB 2     // CTAs per SM the TMEM kernel is sized for
#endif

// task loops of cell_kernel_tm: fully unrolled (one task per warp in the default 448-thread shape); CTA shapes with
// several tasks per warp (experiments) must not unroll, or the compiler hoists every task's loads and spills
#ifdef GA_IT_UNROLL1
#define GA_IT_PRAGMA _Pragma("unroll 1")
#else
#define GA_IT_PRAGMA _Pragma("unroll")
#endif
template <class G, int T, int NW, int GID>
__global__ void __launch_bounds__(T, TM_MINB) cell_kernel_tm(const cf *__restrict__ xd, const cf *__restrict__ cext,
                                                       const int *__restrict__ sv_of_block, const cf *__restrict__ tw,
                                                       int n_cells, int n_dop, int dmax, int wlen, CellStat *__restrict__ cells)
{
    static_assert(T % 32 == 0, "tcgen05.ld/st are warp-collective: whole warps only");
    constexpr int NWARP = T / 32;
    // Work is handed out in warp-tasks of 32 butterflies (task k = butterflies 32k..32k+31), warp w runs
    // tasks w, w+NWARP, ...  (Rotating the map between the two CTAs of an SM to even out the load of the
    // four sub-partitions was measured: no gain, and it made the summation order depend on the CTA.)
    constexpr int NTA = cdiv(G::NA, 32), NTB = cdiv(G::NB, 32), NTC = cdiv(G::NC, 32);
    constexpr int ITA = cdiv(NTA, NWARP), ITB = cdiv(NTB, NWARP), ITC = cdiv(NTC, NWARP);
    constexpr uint32_t COLS_THREAD = ITC * 2 * NW;                 // accumulator floats per thread
    constexpr uint32_t COL_SLOT = (COLS_THREAD + 7u) & ~7u;        // column range of one warp "row" (4 warps share the lanes)
    constexpr uint32_t TM_COLS = pow2_at_least(COL_SLOT * cdiv(NWARP, 4));
#ifndef GA_NO_TM_ASSERT
    static_assert(TM_COLS * TM_MINB <= 512 || G::SMEM_ELEMS * sizeof(cf) * TM_MINB > 227 * 1024,
                  "TM_MINB CTAs per SM must fit in the 512 TMEM columns");
#endif
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
    // this warp's private window: lanes 32*(wid%4).., columns (wid/4)*COL_SLOT..
    const uint32_t tm_mine = tm_base + ((32u * (uint32_t)(wid & 3)) << 16) + (uint32_t)(wid >> 2) * COL_SLOT;
    const int vw = wid;

    for (int cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int blk = cell / n_dop, dop = cell - blk * n_dop - dmax;
        const int sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
        const cf *xb = xd + (size_t)blk * G::N;
        const cf *cb = cext + (size_t)sv * (2 * G::N);
        float best = 0.0f, sum = 0.0f;
        int besti = 0;

        for (int s = 0; s < G::N1; s++) {
            int sp, eoff;
            cell_sub_offsets<G>(s, dop, sp, eoff);
            const cf *xs = xb + (size_t)s * G::N2;
            const cf *cs = cb + (size_t)sp * (2 * G::N2) + eoff;
GA_IT_PRAGMA
            for (int it = 0; it < ITA; it++) {
                const int j = (vw + it * NWARP) * 32 + lane;
                if (j < G::NA) cell_passA<G>(j, s, xs, cs, tw, sm);
            }
            __syncthreads();
GA_IT_PRAGMA
            for (int it = 0; it < ITB; it++) {
                const int j = (vw + it * NWARP) * 32 + lane;
                if (j < G::NB) passB<G, +1>(j, s, tw, sm);
            }
            __syncthreads();
            const cf *ks = c_ktab[GID] + s * G::RC;
GA_IT_PRAGMA
            for (int it = 0; it < ITC; it++) {
                const int task = vw + it * NWARP;
                if (task < NTC) {          // warp-uniform
                    // every lane runs the butterfly (lanes past the end redo the last one) so that the
                    // warp-collective TMEM accesses below are never under divergence
                    const int j = task * 32 + lane;
                    const bool act = j < G::NC;
                    const int jc = act ? j : G::NC - 1;
                    cf p[G::RC];
                    const int tau0 = passC<G, +1>(jc, sm, p);
                    float a[2 * NW];
                    const uint32_t col = tm_mine + (uint32_t)(it * 2 * NW);
                    if (s == 0) {
#pragma unroll
                        for (int w = 0; w < NW; w++) { a[2 * w] = p[w].x; a[2 * w + 1] = p[w].y; }   // ktab[0][w] = 1
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
                    } else if (act) {
                        // last sub-sequence: the outputs are complete -> power, first max, sum (:190-194)
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
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            __syncthreads();      // smem is rewritten by the next sub-sequence's pass A
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
                cells[cell] = r;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
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

// ---------------------------------------------------------------------------------
// snr per Doppler bin and best over Doppler, one thread per block (chunk).
// ave_pwr = tot_pwr/W ; snr = max_pwr/ave_pwr ; strictly-greater scan in ascending dop
// from max_snr = 0 (c/search_offline.cpp:173,196-198); detection rule snr >= 25 (:248).
// ---------------------------------------------------------------------------------
__global__ void best_kernel(const CellStat *__restrict__ cells, const int *__restrict__ sv_of_block,
                            int n_blocks, int n_dop, int dmax, int wlen, Peak *__restrict__ peaks)
{
    // one warp per chunk: lanes take Doppler bins k, k+32, ... in ascending order, then a shuffle
    // reduction that prefers the higher snr and, on equal snr, the LOWER bin -- the same winner as the
    // reference's ascending strictly-greater scan.
    const int blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blk >= n_blocks) return;
    const CellStat *c = cells + (size_t)blk * n_dop;
    float snr = 0.0f;        // max_snr starts at 0: bins with snr <= 0 (or NaN) never win (:173,:198)
    int k = 0x7fffffff;
    for (int i = lane; i < n_dop; i += 32) {
        const float ave = __fdiv_rn(c[i].tot_pwr, (float)wlen);
        const float v = __fdiv_rn(c[i].max_pwr, ave);
        if (v > snr) { snr = v; k = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float os = __shfl_down_sync(0xffffffffu, snr, off);
        const int ok = __shfl_down_sync(0xffffffffu, k, off);
        if (os > snr || (os == snr && ok < k)) { snr = os; k = ok; }
    }
    if (lane == 0) {
        Peak p;
        p.snr = snr; p.max_pwr = 0.0f; p.tot_pwr = 0.0f; p.lo_shift = 0; p.ca_shift = 0;
        if (k != 0x7fffffff) {
            p.lo_shift = k - dmax; p.ca_shift = c[k].max_idx; p.max_pwr = c[k].max_pwr; p.tot_pwr = c[k].tot_pwr;
        }
        p.sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
        p.flags = (snr < 25.0f) ? 0 : 1;
        p.reserved = 0;
        peaks[blk] = p;
    }
}

// natural-order readback helpers for the parity probes -------------------------------
// out[k] = conj(xd[s][q]) with k = N1*q+s  (MODE 0)  or cext[s][q] (MODE 1, stride 2*N2)
__global__ void undecimate_kernel(const cf *__restrict__ in, int n1, int n2, int mode, cf *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n1 * n2) return;
    const int q = k / n1, s = k - q * n1;
    if (mode == 0) out[k] = cconj(in[(size_t)s * n2 + q]);
    else out[k] = in[(size_t)s * 2 * n2 + q];
}

}  // namespace ga
