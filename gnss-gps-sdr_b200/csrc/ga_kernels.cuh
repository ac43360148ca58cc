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
// REF mode, search windows longer than N2 (sampling rates above 10 MHz): output segment m = lags [m*N2, (m+1)*N2)
// of the N-point backward transform is the same sum of sub-sequence transforms with the extra factor
// exp(2*pi*i*s*m/N1) on term s:  y[tau + N2*m] = sum_s w_N^(s*tau) * w_N1^(s*m) * IFFT_N2(P_s)[tau].
// c_ktab_seg[m][s*RC + w] = ktab[s*RC + w] * exp(2*pi*i*s*m/N1)   (geometry 4 x 10000 only)
constexpr int MAX_SEG = 4;
__constant__ cf c_ktab_seg[MAX_SEG][KTAB_MAX];

// ---- mbarrier / bulk-copy helpers (TMA): used by the forward kernel (1-D bulk copy of a chunk's bits) and by the TMA-staged
// cell kernel (ga_cell_tma.cuh: 2-D tensor loads) -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// n bytes (a multiple of 16, source and destination 16-byte aligned) global -> shared, completion bytes on `bar` (SASS UBLKCP)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

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
//
// MODE 0, pass A input by table (lomask != nullptr).  The radix-N1 gather of Sample() -- z = x[n2] + sum_n1 x[N2*n1 + n2] *
// K1[s][n1] with x = (+-1, +-1) from one data bit XOR the two LO bits (c/search_offline.cpp:143-153) -- has only 4^N1
// possible values per sub-sequence s: the CTA builds them once in shared memory (same operation order as fwd_passA, so
// the same bits) and a sample group costs N1 byte loads, a few shifts and ONE table load instead of N1 unpack/select/
// multiply-add chains.  The LO bits of the N1 samples of a group come packed from lomask[n2] (2 bits per sample), the
// data bit is spread over both and XORed in.  N1 = 10 uses two groups of five (1024-entry tables).
#ifndef GA_FWD_BULK
#define GA_FWD_BULK 1       // forward kernel: stage the chunk's bits in shared memory with one TMA bulk copy (0: global byte loads)
#endif
constexpr int FWD_BITS_SMEM = 5120;      // a chunk: 10 packets of 512 bytes (c/search_offline.cpp:129,135-141)
#ifndef FWD_MINB
#define FWD_MINB 2     // measured: 2 CTAs/SM at 126 registers beat 3 at 80 (spills)
#endif
template <class G, int T, int MODE, int GID>
__global__ void __launch_bounds__(T, MODE == 0 ? FWD_MINB : 1) fwd_kernel(const unsigned char *__restrict__ bits, int chunk_bytes,
                                                const unsigned char *__restrict__ lo,
                                                const float *__restrict__ repl_time,
                                                const cf *__restrict__ tw, cf *__restrict__ out,
                                                const unsigned int *__restrict__ lomask = nullptr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    const int item = blockIdx.x / G::N1, s = blockIdx.x - item * G::N1;
    const cf *k1s = c_k1tab[GID] + s * G::N1;

    if (MODE == 0 && lomask != nullptr) {
        typedef FwdLut<G> L;
        cf *lut = sm + G::SMEM_ELEMS;
        // the chunk's packed bits (5120 bytes) are staged in shared memory by ONE bulk copy of the TMA engine
        // (cp.async.bulk, SASS UBLKCP), issued before the tables are built and awaited after: pass A then reads its
        // N1 bytes per sample group from shared memory instead of 100 global byte loads per butterfly
        unsigned char *bits_s = reinterpret_cast<unsigned char *>(lut + L::NG * L::ENTRIES);
        __shared__ __align__(8) unsigned long long bits_bar;
        const unsigned char *chunk_g = bits + (size_t)item * chunk_bytes;
        const bool staged = GA_FWD_BULK && (reinterpret_cast<uintptr_t>(chunk_g) & 15) == 0 && (chunk_bytes & 15) == 0 && chunk_bytes <= FWD_BITS_SMEM;
        if (staged && threadIdx.x == 0) {
            mbar_init(smem_u32(&bits_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(smem_u32(&bits_bar), (uint32_t)chunk_bytes);
            bulk_load_1d(smem_u32(bits_s), chunk_g, (uint32_t)chunk_bytes, smem_u32(&bits_bar));
        }
        for (int e = threadIdx.x; e < L::NG * L::ENTRIES; e += T) lut[e] = fwd_lut_entry<G>(e, k1s);
        __syncthreads();
        if (staged) mbar_wait(smem_u32(&bits_bar), 0);
        const unsigned char *chunk = staged ? bits_s : chunk_g;
        for (int j = threadIdx.x; j < G::NA; j += T) {
            cf p[G::RA];
#pragma unroll
            for (int a = 0; a < G::RA; a++) {
                const int n2 = a * G::NA + j;
                const cf z = fwd_lut_gather<G>(chunk, n2, lomask[n2], lut);
                p[a] = (s == 0) ? z : cmul(z, tw_load<-1>(tw, n2 * s));
            }
            passA_finish<G, -1>(p, j, 0, tw, sm);
        }
    } else if (MODE == 0) {
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

// (the register-accumulator, software-pipelined and rotating-layout variants of the hot kernel -- measured, slower --
// are not part of the product: tools/experiments/ga_experiments.cuh)

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

// Task loop of a warp (task = 32 butterflies): unrolled when a warp has at most two tasks per pass; with more (the
// 128-thread shape) NOT unrolled, or the compiler hoists every task's operand loads and spills.
#ifndef GA_UNROLL_BC
#define GA_UNROLL_BC 1      // the task loops of passes B and C (shared-memory operands) fully unrolled whatever their length:
                            // +2.2 % with three tasks per warp (160-thread CTAs; 128 registers, no spills)
#endif
#ifndef GA_PEAK_BRANCHY
#define GA_PEAK_BRANCHY 1   // peak tracking of pass C with data-dependent branches; 0: branch-free selects (measured 0.2 % slower)
#endif
#ifndef GA_UNROLL_A
#define GA_UNROLL_A 0       // the same for pass A (global operands)
#endif
template <int IT, bool FORCE = false, class F> __device__ __forceinline__ void for_tasks(F &&f)
{
    if constexpr (IT <= 2 || FORCE) {
#pragma unroll
        for (int it = 0; it < IT; it++) f(it);
    } else {
#pragma unroll 1
        for (int it = 0; it < IT; it++) f(it);
    }
}
// CTAs per SM a cell-kernel shape is sized for: 128-thread CTAs run three per SM (201 KB of shared memory, 136
// registers), the 256/448-thread shapes two
constexpr int cell_minb(int threads) { return threads <= 160 ? 3 : 2; }

// (400 butterflies per pass are 13 warp-tasks, so one of a CTA's four warps runs 4 tasks per pass and the others 3.
// Warps sit on sub-partition (hardware warp slot mod 4); the hardware hands co-resident CTAs staggered warp slots
// (0-3 / 5,6,7,4 / 10,11,8,9 -- tools/ubench/warpmap.cu), so the three heavy warps of an SM already land on three
// different sub-partitions: 10/10/10/9 tasks per pass round.  Rotating the warp -> task map per CTA on top of that
// was measured at -10 %: it re-aligns the heavy warps onto one sub-partition.)

// (Measured in round 2 and dropped, one B200, 128-chunk launches: pass-B twiddles from a per-sub-sequence table in shared
// memory instead of the product tree: 8.40 -> 7.56 M corr/s (19 more LDS per butterfly cost more than the 38 packed
// multiplies they save); operands of a warp's first pass-A task loaded before the barrier that ends the previous pass C
// (167 registers): 8.40 -> 8.31 M corr/s.)
#ifndef GA_PERM_A
#define GA_PERM_A 0         // 1: pass-A butterflies dealt in half-warps that never straddle a padded tile row (no store conflicts, but the
                            // operand loads fall apart into 128- and 32-byte pieces: measured 14 % SLOWER, round 2) -- off
#endif
// SEG: the work item is (chunk, Doppler bin, output segment) -- nseg segments of N2 lags each cover a window of up to
// N samples; the per-segment records are merged by merge_seg_kernel.  SEG = false is the plain kernel (one segment).
template <class G, int T, int NW, int GID, bool SEG = false>
__global__ void __launch_bounds__(T, cell_minb(T)) cell_kernel_tm(const cf *__restrict__ xd, const cf *__restrict__ cext,
                                                       const int *__restrict__ sv_of_block, const cf *__restrict__ tw,
                                                       int n_cells, int n_dop, int dmax, int wlen, CellStat *__restrict__ cells,
                                                       int nseg = 1, int blk0 = 0, int *__restrict__ sched = nullptr)
{
    static_assert(T % 32 == 0, "tcgen05.ld/st are warp-collective: whole warps only");
    constexpr int NWARP = T / 32;
    // Work is handed out in warp-tasks of 32 butterflies (task k = butterflies 32k..32k+31), warp w runs
    // tasks w, w+NWARP, ...  (Rotating the map between the two CTAs of an SM to even out the load of the
    // four sub-partitions was measured: no gain, and it made the summation order depend on the CTA.)
    constexpr int NTA = cdiv(G::NA, 32), NTB = cdiv(G::NB, 32), NTC = cdiv(G::NC, 32);
    constexpr int ITA = cdiv(NTA, NWARP), ITB = cdiv(NTB, NWARP), ITC = cdiv(NTC, NWARP);
    constexpr bool SPLIT_C = T <= 256;                             // see pass C: only where 128 registers are available (2 x 256 threads)
    constexpr uint32_t COLS_THREAD = ITC * 2 * NW;                 // accumulator floats per thread
    constexpr uint32_t COL_SLOT = (COLS_THREAD + 7u) & ~7u;        // column range of one warp "row" (4 warps share the lanes)
    // Five warps (160 threads): the fifth warp shares the TMEM lanes of warp 0 and would need a second column slot -- 176
    // columns, 256 allocated, too many for three CTAs per SM.  It runs at most ITC - 1 tasks per pass, though: its first
    // task's accumulators go into the free tail of the 128-column allocation and the second into a separate 32-column
    // allocation (160 columns per CTA, 480 per SM).
    constexpr bool FIVE = NWARP == 5 && ITC == 3 && NTC <= 4 + 2 * NWARP && COL_SLOT + 2 * NW <= 128 && 2 * NW <= 32;
    constexpr uint32_t TM_COLS = FIVE ? 128u : pow2_at_least(COL_SLOT * cdiv(NWARP, 4));
    constexpr uint32_t TM_COLS_ALL = FIVE ? 160u : TM_COLS;
#ifndef GA_NO_TM_ASSERT
    static_assert(TM_COLS_ALL * cell_minb(T) <= 512 || G::SMEM_ELEMS * sizeof(cf) * cell_minb(T) > 227 * 1024,
                  "cell_minb(T) CTAs per SM must fit in the 512 TMEM columns");
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cf *sm = reinterpret_cast<cf *>(smem_raw);
    __shared__ float red_best[NWARP], red_sum[NWARP];
    __shared__ int red_idx[NWARP];
    __shared__ uint32_t tm_base_s, tm_base2_s;
    __shared__ int next_cell_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_base_s)), "r"(TM_COLS) : "memory");
        if (FIVE) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm_base2_s)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_base_s;
    // this warp's private window: lanes 32*(wid%4).., columns (wid/4)*COL_SLOT..
    const uint32_t tm_mine = tm_base + ((32u * (uint32_t)(wid & 3)) << 16) + (uint32_t)(wid >> 2) * COL_SLOT;
    const uint32_t tm_second = FIVE ? tm_base2_s : 0u;      // fifth warp, second task (lanes 0..31 of the 32-column allocation)
    // Five warps: the hardware puts warps 0 and 4 of a CTA on the same SM sub-partition, and the three co-resident CTAs on
    // three different ones (tools/ubench/warpmap.cu with 160 threads).  The two warps that get only two tasks per pass are
    // therefore warps 0 and 4: every CTA loads its sub-partitions 4/3/3/3 and the SM 10/10/10/9 (warps 3 and 4: 10/10/11/8).
#ifdef GA_VW_OLD
    const int vw = wid;
#else
    const int vw = FIVE ? (wid + 4) % 5 : wid;
#endif

    // Cells are handed out in ascending order by a device-wide ticket counter (sched[0]; the first gridDim.x tickets are
    // the CTA numbers): whichever CTA finishes takes the next cell, so the CTAs of a launch work on a narrow window of
    // neighbouring (chunk, Doppler bin) cells however long the launch is -- a block spectrum is fetched from HBM once
    // and shared through L2 by the 73 cells that use it -- and a launch ends with at most one cell of imbalance.
    // The ticket is drawn by thread 0 at the start of the last sub-sequence (latency hidden) and published through
    // shared memory at the barrier of the per-cell reduction.  The last CTA to leave rewinds the counter.
    int next_cell = 0;
    for (int cell = blockIdx.x; cell < n_cells; cell = next_cell) {
        int item = cell, seg = 0;
        if (SEG) { seg = cell % nseg; item = cell / nseg; }
        const int blk = item / n_dop, dop = item - blk * n_dop - dmax;
        const int wl = SEG ? wlen - seg * G::N2 : wlen;                 // lags of this segment inside the search window
        const int sv = (sv_of_block ? sv_of_block[blk] : blk + blk0) & 31;     // caller-supplied device maps are not trusted: see best_kernel
        const cf *xb = xd + (size_t)blk * G::N;
        const cf *cb = cext + (size_t)sv * (2 * G::N);
        float best = 0.0f, sum = 0.0f;
        int besti = 0;
        for (int s = 0; s < G::N1; s++) {
            int sp, eoff;
            cell_sub_offsets<G>(s, dop, sp, eoff);
            const cf *xs = xb + (size_t)s * G::N2;
            const cf *cs = cb + (size_t)sp * (2 * G::N2) + eoff;
            if (s == G::N1 - 1 && tid == 0) next_cell = sched ? (int)gridDim.x + atomicAdd(sched, 1) : cell + (int)gridDim.x;
            for_tasks<ITA, GA_UNROLL_A != 0>([&](int it) {
                const int jj = (vw + it * NWARP) * 32 + lane;
                if (jj < G::NA) cell_passA<G>(GA_PERM_A ? passA_slot_to_j<G>(jj) : jj, s, xs, cs, tw, sm);
            });
            __syncthreads();
            for_tasks<ITB, GA_UNROLL_BC != 0>([&](int it) {
                const int j = (vw + it * NWARP) * 32 + lane;
                if (j < G::NB) passB<G, +1>(j, s, tw, sm);
            });
            __syncthreads();
            const cf *ks = (SEG ? c_ktab_seg[seg] : c_ktab[GID]) + s * G::RC;
            for_tasks<ITC, GA_UNROLL_BC != 0>([&](int it) {
                const int task = vw + it * NWARP;
                if (task < NTC) {          // warp-uniform
                    // every lane runs the butterfly (lanes past the end redo the last one) so that the
                    // warp-collective TMEM accesses below are never under divergence
                    const int j = task * 32 + lane;
                    const bool act = j < G::NC;
                    const int jc = act ? j : G::NC - 1;
                    cf p[G::RC];
                    const int tau0 = passC<G, +1>(jc, sm, p);
                    const uint32_t col = (FIVE && wid == 4 && it == 1) ? tm_second : tm_mine + (uint32_t)(it * 2 * NW);
                    // acc = TMEM accumulators + p * ktab[s]
                    auto accumulate = [&](float (&a)[2 * NW]) {
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
                    };
                    // last sub-sequence: the outputs are complete -> power, first max, sum (:190-194)
                    auto peak = [&](const float (&a)[2 * NW]) {
#if GA_PEAK_BRANCHY
#pragma unroll
                        for (int w = 0; w < NW; w++) {
                            const int tau = tau0 + G::OUT_STRIDE * w;
                            if (tau < wl) {
                                const float pwr = fmaf(a[2 * w], a[2 * w], a[2 * w + 1] * a[2 * w + 1]);
                                if (pwr > best || (pwr == best && tau < besti)) { best = pwr; besti = tau; }
                                sum += pwr;
                            }
                        }
#else
                        // branch-free: same comparisons, same summation order; a lag outside the window takes part with
                        // power -1 (never a maximum: powers are >= 0 and best starts at 0) and adds nothing to the sum
#pragma unroll
                        for (int w = 0; w < NW; w++) {
                            const int tau = tau0 + G::OUT_STRIDE * w;
                            const bool in = tau < wl;
                            const float pwr = fmaf(a[2 * w], a[2 * w], a[2 * w + 1] * a[2 * w + 1]);
                            const float cand = in ? pwr : -1.0f;
                            const bool better = cand > best || (cand == best && tau < besti);
                            best = better ? cand : best;
                            besti = better ? tau : besti;
                            sum = in ? sum + pwr : sum;
                        }
#endif
                    };
                    if constexpr (SPLIT_C) {
                        // store path and power path as separate code: their registers are allocated independently and
                        // the compiler needs no copies between them (-5 % instructions; needs register headroom)
                        if (s < G::N1 - 1) {
                            float a[2 * NW];
                            accumulate(a);
                            tm_move<2 * NW, false>(col, a);
                        } else {
                            float a[2 * NW];
                            accumulate(a);
                            if (act) peak(a);
                        }
                    } else {
                        float a[2 * NW];
                        accumulate(a);
                        if (s < G::N1 - 1) tm_move<2 * NW, false>(col, a);
                        else if (act) peak(a);
                    }
                }
            });
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            // the tile is rewritten by the next sub-sequence's pass A; after the LAST sub-sequence the barrier of the
            // per-cell reduction below does that job (one barrier less per cell)
#ifdef GA_NO_BARRIER_MERGE
            __syncthreads();
#else
            if (s < G::N1 - 1) __syncthreads();
#endif
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
        next_cell = next_cell_s;          // rewritten by thread 0 only after the 3*N1 barriers of the next cell
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
                CellStat r; r.max_pwr = best; r.tot_pwr = sum; r.max_idx = besti + (SEG ? seg * G::N2 : 0); r.pad = 0;
                cells[cell] = r;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(TM_COLS) : "memory");
        if (FIVE) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_second), "r"(32u) : "memory");
    }
    // every CTA has drawn its last ticket (one past the end) before it gets here: the last one to arrive rewinds
    if (sched && tid == 0 && atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
}

// per-segment records [pair][nseg] -> one record per (chunk, Doppler bin): the reference's single scan over i < W
// (c/search_offline.cpp:190-194) visits the segments in ascending order, so the first maximum is the lowest segment's on
// equal power and the power sum adds the segment sums in that order.
__global__ void merge_seg_kernel(const CellStat *__restrict__ seg_cells, int n_pairs, int nseg, CellStat *__restrict__ cells)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    CellStat r = seg_cells[(size_t)i * nseg];
    for (int m = 1; m < nseg; m++) {
        const CellStat c = seg_cells[(size_t)i * nseg + m];
        if (c.max_pwr > r.max_pwr) { r.max_pwr = c.max_pwr; r.max_idx = c.max_idx; }
        r.tot_pwr += c.tot_pwr;
    }
    cells[i] = r;
}

// ---------------------------------------------------------------------------------
// snr per Doppler bin and best over Doppler, one thread per block (chunk).
// ave_pwr = tot_pwr/W ; snr = max_pwr/ave_pwr ; strictly-greater scan in ascending dop
// from max_snr = 0 (c/search_offline.cpp:173,196-198); detection rule snr >= 25 (:248).
// ---------------------------------------------------------------------------------
__global__ void best_kernel(const CellStat *__restrict__ cells, const int *__restrict__ sv_of_block,
                            int n_blocks, int n_dop, int dmax, int wlen, Peak *__restrict__ peaks, int blk0 = 0)
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
        p.sv = sv_of_block ? sv_of_block[blk] : ((blk + blk0) & 31);
        p.flags = (snr < 25.0f) ? 0 : 1;
        if (p.sv < 0 || p.sv > 31) p.flags |= (int)0x80000000u;    // device-side PRN map entry out of range: searched as sv & 31
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

