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
#include "ga_fft3.h"

namespace ga {

struct CellStat { float max_pwr, tot_pwr; int max_idx, pad; };            // == gpsacq_cell
struct Peak { float snr, max_pwr, tot_pwr; int lo_shift, ca_shift, sv, flags, reserved; };  // == gpsacq_peak

constexpr int KTAB_MAX = 256;      // N1*RC <= 250 for the geometries below
constexpr int K1TAB_MAX = 128;     // N1*N1 <= 100
constexpr int NGEOM = 3;
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
template <class G, int T, int MODE, int GID>
__global__ void __launch_bounds__(T) fwd_kernel(const unsigned char *__restrict__ bits, int chunk_bytes,
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
    for (int j = threadIdx.x; j < G::NC; j += T) {
        cf p[G::RC];
        const int tau0 = passC<G, -1>(j, sm, p);
        if (MODE == 0) {
            cf *dst = out + ((size_t)item * G::N1 + s) * G::N2 + tau0;
#pragma unroll
            for (int w = 0; w < G::RC; w++) dst[G::OUT_STRIDE * w] = cconj(p[w]);
        } else {
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
// snr per Doppler bin and best over Doppler, one thread per block (chunk).
// ave_pwr = tot_pwr/W ; snr = max_pwr/ave_pwr ; strictly-greater scan in ascending dop
// from max_snr = 0 (c/search_offline.cpp:173,196-198); detection rule snr >= 25 (:248).
// ---------------------------------------------------------------------------------
__global__ void best_kernel(const CellStat *__restrict__ cells, const int *__restrict__ sv_of_block,
                            int n_blocks, int n_dop, int dmax, int wlen, Peak *__restrict__ peaks)
{
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= n_blocks) return;
    const CellStat *c = cells + (size_t)blk * n_dop;
    Peak p; p.snr = 0.0f; p.max_pwr = 0.0f; p.tot_pwr = 0.0f; p.lo_shift = 0; p.ca_shift = 0;
    float max_snr = 0.0f;
    for (int k = 0; k < n_dop; k++) {
        const float ave = __fdiv_rn(c[k].tot_pwr, (float)wlen);
        const float snr = __fdiv_rn(c[k].max_pwr, ave);
        if (snr > max_snr) {
            max_snr = snr; p.lo_shift = k - dmax; p.ca_shift = c[k].max_idx;
            p.max_pwr = c[k].max_pwr; p.tot_pwr = c[k].tot_pwr;
        }
    }
    p.snr = max_snr;
    p.sv = sv_of_block ? sv_of_block[blk] : (blk & 31);
    p.flags = (max_snr < 25.0f) ? 0 : 1;
    p.reserved = 0;
    peaks[blk] = p;
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
