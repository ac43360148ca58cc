// gpsacq.cu -- C ABI (include/gpsacq.h) of the B200 GPS L1 C/A acquisition engine.
//
// Owns all device state (the reference keeps the same things in file statics,
// c/search_offline.cpp:55-64): twiddle table, LO table, code-NCO tables, the 32
// replica spectra, and per-batch workspaces.  No CPU compute path exists here: if
// CUDA is unavailable gpsacq_create() fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>
#include <mutex>

#include "../../include/gpsacq.h"

#ifndef CELL_MINB
#define CELL_MINB 2        // CTAs per SM the cell kernels are sized for (registers, smem carveout, TMEM columns)
#endif
#define TM_MINB CELL_MINB
#include "ga_kernels.cuh"
#include "ga_cell_tma.cuh"
#include "ga_grid.cuh"
#include "ga_pfa.cuh"
#include "ga_frontend.cuh"
#include "ga_siggen.cuh"
#include "ga_tables.h"
#include "ga_frontend_host.h"

using namespace ga;

static_assert(sizeof(gpsacq_peak) == sizeof(Peak) && sizeof(gpsacq_peak) == 32, "peak record layout");
static_assert(sizeof(gpsacq_cell) == sizeof(CellStat) && sizeof(gpsacq_cell) == 16, "cell record layout");

// ---- geometries: N = 40000 = N1 * (RA*RB*RC), N2 >= W --------------------------------
typedef Geom<10, 20, 20, 10> G4000;    // W <= 4000   (FS <= 4 MHz, e.g. rtl-sdr 2.8 MHz)
typedef Geom<5, 20, 20, 20> G8000;     // W <= 8000   (FS <= 8 MHz, e.g. 5.456 MHz)
typedef Geom<4, 25, 20, 20> G10000;    // W <= 10000  (FS <= 10 MHz, e.g. 8.184 MHz, 10 MHz)
enum { GID_4000 = 0, GID_8000 = 1, GID_10000 = 2 };
// GRID mode: L = 2*N2 >= 2W (linear-correlation embedding of the W-point circular correlation)
typedef Geom<2, 20, 20, 10> H4000;     // L = 8000,  W <= 4000
typedef Geom<2, 20, 20, 20> H8000;     // L = 16000, W <= 8000
typedef Geom<2, 25, 20, 20> H10000;    // L = 20000, W <= 10000
typedef Geom<2, 20, 20, 16> H6400;     // L = 12800, W <= 6400 (e.g. 5.456 MHz: 20 % less work than L = 16000)
enum { HID_4000 = 3, HID_8000 = 4, HID_10000 = 5, HID_6400 = 6 };
// GRID mode, native W-point prime-factor transforms (ga_pfa.h) for the 1 ms block lengths of the named
// sampling rates; any other W goes through the embedding above.
// pass order: which factor runs in pass A (global loads + product), B, C (power + statistics);
// PFA_ORDER_x permutes the base triple (f0,f1,f2): 0 = (f0,f1,f2), 1 = (f2,f0,f1), 2 = (f1,f2,f0),
// 3 = (f2,f1,f0), 4 = (f0,f2,f1), 5 = (f1,f0,f2)
template <int F0, int F1, int F2, int ORDER> struct PermGeom;
template <int F0, int F1, int F2> struct PermGeom<F0, F1, F2, 0> { typedef PGeom<F0, F1, F2> type; };
template <int F0, int F1, int F2> struct PermGeom<F0, F1, F2, 1> { typedef PGeom<F2, F0, F1> type; };
template <int F0, int F1, int F2> struct PermGeom<F0, F1, F2, 2> { typedef PGeom<F1, F2, F0> type; };
template <int F0, int F1, int F2> struct PermGeom<F0, F1, F2, 3> { typedef PGeom<F2, F1, F0> type; };
template <int F0, int F1, int F2> struct PermGeom<F0, F1, F2, 4> { typedef PGeom<F0, F2, F1> type; };
template <int F0, int F1, int F2> struct PermGeom<F0, F1, F2, 5> { typedef PGeom<F1, F0, F2> type; };
#ifndef PFA_ORDER_5456
#define PFA_ORDER_5456 2
#endif
#ifndef PFA_ORDER_8184
#define PFA_ORDER_8184 2
#endif
#ifndef PFA_ORDER_2800
#define PFA_ORDER_2800 2
#endif
typedef PermGeom<16, 11, 31, PFA_ORDER_5456>::type P5456;    // fs = 5.456 MHz
typedef PermGeom<24, 11, 31, PFA_ORDER_8184>::type P8184;    // fs = 8.184 MHz
typedef PermGeom<16, 25, 7, PFA_ORDER_2800>::type P2800;      // fs = 2.8 MHz
enum { PID_5456 = 20, PID_8184 = 21, PID_2800 = 22 };
// GRID mode, exact-length transforms for 2/5-smooth block lengths: the same twiddled three-pass transform as the
// embedding, but N1 = 1 and L = W (no zero padding, one code period) -- e.g. the receiver's own FS = 10 MHz
// (c/gps.h:24), 8 / 4 MHz, and the power-of-two SDR rates 4.096 / 2.048 MHz.
typedef Geom<1, 25, 20, 20> X10000;
typedef Geom<1, 20, 20, 20> X8000;
typedef Geom<1, 20, 20, 10> X4000;
typedef Geom<1, 16, 16, 16> X4096;
typedef Geom<1, 16, 16, 8> X2048;
enum { XID_10000 = 7, XID_8000 = 8, XID_4000 = 9, XID_4096 = 10, XID_2048 = 11 };
#ifndef PFA_T_5456
#define PFA_T_5456 128      // 4 warps (a multiple of the 4 sub-partitions keeps the per-thread register budget whole)
#endif
#ifndef PFA_B_5456
#define PFA_B_5456 4
#endif
#ifndef PFA_T_8184
#define PFA_T_8184 128
#endif
#ifndef PFA_B_8184
#define PFA_B_8184 3
#endif
#ifndef PFA_T_8184_K1
#define PFA_T_8184_K1 PFA_T_8184      // K = 1 handles (no tensor memory): their own CTA shape
#endif
#ifndef PFA_T_5456_K1
#define PFA_T_5456_K1 PFA_T_5456
#endif
#ifndef PFA_B_5456_K1
#define PFA_B_5456_K1 PFA_B_5456
#endif
#ifndef PFA_T_2800
#define PFA_T_2800 128
#endif
#ifndef PFA_B_2800
#define PFA_B_2800 5
#endif

#ifndef CELL_T_4000
#define CELL_T_4000 128     // 3 CTAs/SM like the benchmark geometry (+3 % over 256 x 2)
#endif
#ifndef CELL_T_8000
#define CELL_T_8000 160     // 5 warps x 3 CTAs/SM: the 13 tasks of a pass are dealt 3/3/3/2/2 (three rounds; four warps need four),
                            // 118 registers, the fifth warp's accumulators in a second small TMEM allocation (ga_kernels.cuh).
                            // Round 2, ticket scheduler, one B200: 160 x 3 9.00 M corr/s, 128 x 3 8.91, 256 x 2 8.91, 224 x 2 8.67,
                            // 448 x 2 8.53, 192 x 2 8.14
#endif
#ifndef CELL_T_8000_WIDE
#define CELL_T_8000_WIDE 256  // search windows above 5600 samples (20 accumulators per butterfly: too many TMEM columns for three CTAs per SM):
                              // 256 x 2, 119 registers; the 448 x 2 shape of round 1 spilled (FS = 7 MHz: 6.83 -> 8.49 M corr/s)
#endif
#ifndef CELL_T_10000
#define CELL_T_10000 256
#endif
#ifndef FWD_T
#define FWD_T 256
#endif

static const double kCPS = 1.023e6;    // chip rate, c/gps_offline.h:30

// PRN -> G2 taps (c/search_offline.cpp:20-53; IS-GPS-200 Table 3-Ia), index = PRN-1
static const unsigned char kTaps[32][2] = {
    {2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3},
    {3, 4}, {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7},
    {5, 8}, {6, 9}, {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7},
    {3, 8}, {4, 9},
};

static thread_local std::string g_create_error;

// Every entry point that selects a device restores the caller's current device on return: a host application
// (torch, another CUDA library on the same thread) keeps launching where it was.
struct DeviceGuard {
    int prev;
    DeviceGuard() : prev(-1) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct gpsacq {
    gpsacq_cfg cfg;
    int gid, n, n1, n2, w, dmax, ndop, chunk_bytes, chunk_samples, cap, device, sm_count;
    int cell_ctas, cell_threads, cell_smem, cell_nw;
    double *d_iq_tab; unsigned long long iq_tab_p, iq_tab_q; bool iq_tab_neg;     // 8-bit front-end: phasor table of the last shift
    unsigned *d_iq_thr; bool iq_thr_attr;                                          // ... and its threshold table (ga_frontend_math.h)
    int sub_blocks, last_launch_blocks; // REF: chunks per kernel launch (a batch is cut into launches whose block spectra stay in L2)
    int nseg;                          // REF: output segments of N2 lags per cell (1 unless W > N2, i.e. FS > 10 MHz)
    CellStat *d_cells_seg;
    cf *d_chalo; int halo; bool use_tma; CUtensorMap map_x, map_c;    // REF, TMA-staged cell kernel (ga_cell_tma.cuh)
    int *d_sched;                      // REF: ticket counter of the cell kernel's work queue (ga_kernels.cuh) + exit counter
    int mode, kblocks, wipe_m, block_bytes, cap_acq;
    int q_min, n_q;                    // GRID, native path: replica spectra are stored rotated by -q for q_min <= q < q_min + n_q
    cf *d_crot;
    int n_base;                        // GRID, native path: R = 1000/step base spectra per block shared by all bins (0 = one per bin)
    int dmax_full, dop_first;          // GRID shard: h->dmax is dmax_full - dop_first, so that bin = index - h->dmax is absolute
    double step;
    cf *d_wipe, *d_xg;
    float *d_code_w;
    cudaStream_t own_stream, stream, copy_stream;
    cudaEvent_t ev[4], ev_copy[8], ev_done;
    bool have_batch;
    size_t last_blocks;
    // device
    cf *d_tw;
    unsigned char *d_lo;
    unsigned int *d_lomask;            // REF forward kernel: LO bits of the N1 samples of a radix-N1 group, 2 bits each (ga_kernels.cuh)
    unsigned short *d_chip_idx;
    float *d_blend_a, *d_blend_b, *d_repl_time;
    cf *d_cext, *d_xd, *d_nat;
    unsigned char *d_bits;
    int *d_sv;
    CellStat *d_cells;
    Peak *d_peaks;
    // pinned host staging
    unsigned char *h_bits;
    int *h_sv;
    Peak *h_peaks;
    std::string err;
};

#define CUDA_TRY(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (h)->err = b_;                                                                         \
            return GPSACQ_ECUDA;                                                                   \
        }                                                                                          \
    } while (0)

// ---- kernel dispatch ------------------------------------------------------------------
// (the measured-and-slower variants of the hot kernel -- register accumulators, software-pipelined operand prefetch,
// rotating shared-memory layouts -- are kept outside the product under tools/experiments/)
template <class G, int T, int NW, int GID, bool SEG = false> struct CellKernel {
    static auto get() { return cell_kernel_tm<G, T, NW, GID, SEG>; }
};

template <class G, int T, int NW, int GID, bool SEG = false>
static int launch_cells_t(gpsacq *h, size_t n_blocks, const int *d_sv, size_t off, int blk0)
{
    // off: first workspace slot (block spectra / cell records) of this (sub-)batch
    const int n_pairs = (int)(n_blocks * (size_t)h->ndop), n_cells = n_pairs * (SEG ? h->nseg : 1);
    const int grid = std::min(n_cells, h->cell_ctas);
    CellStat *out = SEG ? h->d_cells_seg + off * (size_t)h->ndop * h->nseg : h->d_cells + off * (size_t)h->ndop;
    CellKernel<G, T, NW, GID, SEG>::get()<<<grid, T, h->cell_smem, h->stream>>>(
        h->d_xd + off * (size_t)h->n, h->d_cext, d_sv, h->d_tw, n_cells, h->ndop, h->dmax, h->w, out, h->nseg, blk0, h->d_sched);
    CUDA_TRY(h, cudaGetLastError());
    if (SEG) {
        merge_seg_kernel<<<(n_pairs + 255) / 256, 256, 0, h->stream>>>(out, n_pairs, h->nseg, h->d_cells + off * (size_t)h->ndop);
        CUDA_TRY(h, cudaGetLastError());
    }
    return 0;
}

template <class G, int T, int NW, int GID, bool SEG = false>
static int setup_cells_t(gpsacq *h)
{
    auto kern = CellKernel<G, T, NW, GID, SEG>::get();
    h->cell_smem = (int)(G::SMEM_ELEMS * sizeof(cf));
    h->cell_threads = T;
    h->cell_nw = NW;
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cell_smem));
    // ask for a shared-memory carveout that fits two CTAs (the rest stays L1 for the operand loads)
    constexpr int MINB = cell_minb(T);
    const int want = MINB * (h->cell_smem + 2048);
    int pct = (int)((want * 100LL + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 88) pct = 100;
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    int per_sm = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, h->cell_smem));
    if (per_sm < 1) { h->err = "cell kernel does not fit on an SM"; return GPSACQ_ECUDA; }
    // The occupancy query is conservative for kernels that allocate tensor memory (it cannot see the
    // column count, a run-time operand of tcgen05.alloc, and reports one CTA per SM).  Registers and
    // shared memory are sized for cell_minb(T) CTAs (__launch_bounds__, carveout above) and each CTA
    // allocates at most 512 / cell_minb(T) TMEM columns, so that many CTAs are resident.
    if (per_sm < MINB) per_sm = MINB;
    h->cell_ctas = per_sm * h->sm_count;
    return 0;
}

template <class G, int MODE, int GID>
static int launch_fwd_t(gpsacq *h, size_t n_items, const unsigned char *d_bits, cf *out)
{
    auto kern = fwd_kernel<G, FWD_T, MODE, GID>;
    // opted in once by setup_fwd_t(); MODE 0 carries the sample-group tables behind the tile
    const int smem = (int)(G::SMEM_ELEMS * sizeof(cf)) + (MODE == 0 && h->d_lomask ? FwdLut<G>::BYTES + FWD_BITS_SMEM : 0);
    kern<<<(unsigned)(n_items * G::N1), FWD_T, smem, h->stream>>>(d_bits, h->chunk_bytes, h->d_lo, h->d_repl_time,
                                                                  h->d_tw, out, MODE == 0 ? h->d_lomask : nullptr);
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

template <class G, int GID> static int setup_fwd_t(gpsacq *h)
{
    const int smem = (int)(G::SMEM_ELEMS * sizeof(cf));
    CUDA_TRY(h, cudaFuncSetAttribute(fwd_kernel<G, FWD_T, 0, GID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + FwdLut<G>::BYTES + FWD_BITS_SMEM));
    CUDA_TRY(h, cudaFuncSetAttribute(fwd_kernel<G, FWD_T, 1, GID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return 0;
}

// ---- TMA-staged REF cell kernel (ga_cell_tma.cuh): tensor maps, halo replica layout, launch --------------------------
#ifndef CELL_TMA_NCW
#define CELL_TMA_NCW 7          // consumer warps per CTA (+ 1 producer warp), 2 CTAs per SM
#endif
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D tensor of complex64 elements (8 bytes, described to the TMA as FLOAT64): rows x cols, box = box_cols x box_rows
static int make_map_2d(gpsacq *h, CUtensorMap *map, void *base, size_t cols, size_t rows, int box_cols, int box_rows)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { h->err = "cuTensorMapEncodeTiled is not available in this driver"; return GPSACQ_ECUDA; }
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)(cols * sizeof(cf))};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        h->err = b; return GPSACQ_ECUDA;
    }
    return 0;
}

template <class G, int NW, int GID> static int setup_cells_tma_t(gpsacq *h)
{
    auto kern = cell_kernel_tma<G, CELL_TMA_NCW, NW, GID>;
    h->halo = h->dmax / G::N1 + 2;
    const size_t pitch = (size_t)G::NA + 2 * (size_t)h->halo, rows = (size_t)32 * G::N1 * G::RA;
    CUDA_TRY(h, cudaMalloc(&h->d_chalo, rows * pitch * sizeof(cf)));
    int rc = make_map_2d(h, &h->map_x, h->d_xd, G::NA, (size_t)h->cap * G::N1 * G::RA, TMA_BOX_COLS, G::RA);
    if (!rc) rc = make_map_2d(h, &h->map_c, h->d_chalo, pitch, rows, TMA_BOX_COLS_C, G::RA);
    if (rc) return rc;
    h->cell_smem = TmaCellShape<G>::SMEM_BYTES;
    h->cell_threads = 32 * (CELL_TMA_NCW + 1);
    h->cell_nw = NW;
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cell_smem));
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->cell_ctas = 2 * h->sm_count;        // registers (__launch_bounds__), shared memory and TMEM columns are sized for two CTAs per SM
    h->use_tma = true;
    return 0;
}

template <class G> static int build_halo_t(gpsacq *h)
{
    replica_halo_kernel<G><<<32 * G::N1 * G::RA, 128, 0, h->stream>>>(h->d_cext, h->halo, h->d_chalo);
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

template <class G, int NW, int GID>
static int launch_cells_tma_t(gpsacq *h, size_t n_blocks, const int *d_sv, size_t off, int blk0)
{
    const int n_cells = (int)(n_blocks * (size_t)h->ndop);
    const int grid = std::min(n_cells, h->cell_ctas);
    cell_kernel_tma<G, CELL_TMA_NCW, NW, GID><<<grid, h->cell_threads, h->cell_smem, h->stream>>>(
        h->map_x, h->map_c, d_sv, h->d_tw, n_cells, h->ndop, h->dmax, h->w, h->halo, (int)(off * G::N1 * G::RA),
        h->d_cells + off * (size_t)h->ndop, blk0, h->d_sched, h->d_xd + off * (size_t)h->n);
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

static int setup_cells(gpsacq *h)
{
    {
        const int rc = h->gid == GID_4000 ? setup_fwd_t<G4000, GID_4000>(h)
                     : h->gid == GID_8000 ? setup_fwd_t<G8000, GID_8000>(h) : setup_fwd_t<G10000, GID_10000>(h);
        if (rc) return rc;
    }
    switch (h->gid) {
    case GID_4000:
        return h->w <= 7 * G4000::OUT_STRIDE ? setup_cells_t<G4000, CELL_T_4000, 7, GID_4000>(h)
                                             : setup_cells_t<G4000, CELL_T_4000, 10, GID_4000>(h);
    case GID_8000:
        // the benchmark geometry (W <= 5600): GPSACQ_CELL_TMA=1 selects the kernel whose operands are staged by TMA
        // (ga_cell_tma.cuh; measured 13 % slower than the __ldg kernel)
        if (h->w <= 14 * G8000::OUT_STRIDE && h->d_sched && getenv("GPSACQ_CELL_TMA") && atoi(getenv("GPSACQ_CELL_TMA")) > 0)
            return setup_cells_tma_t<G8000, 14, GID_8000>(h);
        return h->w <= 14 * G8000::OUT_STRIDE ? setup_cells_t<G8000, CELL_T_8000, 14, GID_8000>(h)
                                              : setup_cells_t<G8000, CELL_T_8000_WIDE, 20, GID_8000>(h);
    default:
        if (h->nseg > 1) return setup_cells_t<G10000, CELL_T_10000, 20, GID_10000, true>(h);
        return h->w <= 17 * G10000::OUT_STRIDE ? setup_cells_t<G10000, CELL_T_10000, 17, GID_10000>(h)
                                               : setup_cells_t<G10000, CELL_T_10000, 20, GID_10000>(h);
    }
}

static int launch_cells(gpsacq *h, size_t n_blocks, const int *d_sv, size_t off, int blk0)
{
    if (h->use_tma) return launch_cells_tma_t<G8000, 14, GID_8000>(h, n_blocks, d_sv, off, blk0);
    switch (h->gid) {
    case GID_4000:
        return h->cell_nw == 7 ? launch_cells_t<G4000, CELL_T_4000, 7, GID_4000>(h, n_blocks, d_sv, off, blk0)
                               : launch_cells_t<G4000, CELL_T_4000, 10, GID_4000>(h, n_blocks, d_sv, off, blk0);
    case GID_8000:
        return h->cell_nw == 14 ? launch_cells_t<G8000, CELL_T_8000, 14, GID_8000>(h, n_blocks, d_sv, off, blk0)
                                : launch_cells_t<G8000, CELL_T_8000_WIDE, 20, GID_8000>(h, n_blocks, d_sv, off, blk0);
    default:
        if (h->nseg > 1) return launch_cells_t<G10000, CELL_T_10000, 20, GID_10000, true>(h, n_blocks, d_sv, off, blk0);
        return h->cell_nw == 17 ? launch_cells_t<G10000, CELL_T_10000, 17, GID_10000>(h, n_blocks, d_sv, off, blk0)
                                : launch_cells_t<G10000, CELL_T_10000, 20, GID_10000>(h, n_blocks, d_sv, off, blk0);
    }
}

static int launch_fwd(gpsacq *h, int mode, size_t n_items, const unsigned char *d_bits, cf *out)
{
    switch (h->gid) {
    case GID_4000:
        return mode == 0 ? launch_fwd_t<G4000, 0, GID_4000>(h, n_items, d_bits, out) : launch_fwd_t<G4000, 1, GID_4000>(h, n_items, d_bits, out);
    case GID_8000:
        return mode == 0 ? launch_fwd_t<G8000, 0, GID_8000>(h, n_items, d_bits, out) : launch_fwd_t<G8000, 1, GID_8000>(h, n_items, d_bits, out);
    default:
        return mode == 0 ? launch_fwd_t<G10000, 0, GID_10000>(h, n_items, d_bits, out) : launch_fwd_t<G10000, 1, GID_10000>(h, n_items, d_bits, out);
    }
}

template <class G, int GID> static int upload_const_t(gpsacq *h)
{
    std::vector<cf> kt = make_ktab<G>(), k1 = make_k1tab<G>();
    CUDA_TRY(h, cudaMemcpyToSymbol(c_ktab, kt.data(), kt.size() * sizeof(cf), (size_t)GID * KTAB_MAX * sizeof(cf)));
    CUDA_TRY(h, cudaMemcpyToSymbol(c_k1tab, k1.data(), k1.size() * sizeof(cf), (size_t)GID * K1TAB_MAX * sizeof(cf)));
    return 0;
}

// ---- host-side sequential NCOs (must follow the reference's float recurrences) --------
// Code NCO, SearchInit() c/search_offline.cpp:76,84-99.  For sample i: chip index before
// the update, and the blend weights (1,0) or ((float)(1.0-ca_phase), ca_phase) on a
// chip-edge crossing.
static void build_code_nco(double fs, int n, std::vector<unsigned short> &idx, std::vector<float> &a, std::vector<float> &b)
{
    const float ca_rate = (float)(kCPS / fs);
    float ca_phase = 0;
    int chip = 0;
    idx.resize(n); a.resize(n); b.resize(n);
    for (int i = 0; i < n; i++) {
        idx[i] = (unsigned short)chip;
        ca_phase += ca_rate;
        if (ca_phase >= 1.0) {
            ca_phase -= 1.0;
            chip = chip + 1 == 1023 ? 0 : chip + 1;
            a[i] = (float)(1.0 - (double)ca_phase);
            b[i] = ca_phase;
        } else {
            a[i] = 1.0f; b[i] = 0.0f;
        }
    }
}

// LO NCO, Sample() c/search_offline.cpp:124-127,131,152-156: per-sample phase index
// int(lo_phase), packed as lo_cos bit | lo_sin bit << 1.
static void build_lo_table(double fc, double fs, int n, std::vector<unsigned char> &lo)
{
    static const int lo_sin[4] = {1, 1, 0, 0}, lo_cos[4] = {0, 1, 1, 0};
    const float lo_rate = (float)(4 * fc / fs);
    float lo_phase = 0;
    lo.resize(n);
    for (int i = 0; i < n; i++) {
        const int k = (int)lo_phase;
        lo[i] = (unsigned char)(lo_cos[k & 3] | (lo_sin[k & 3] << 1));
        lo_phase += lo_rate;
        if (lo_phase >= 4) lo_phase -= 4;
    }
}

static void free_all(gpsacq *h)
{
    if (!h) return;
    cudaFree(h->d_tw); cudaFree(h->d_lo); cudaFree(h->d_lomask); cudaFree(h->d_chip_idx); cudaFree(h->d_blend_a); cudaFree(h->d_blend_b);
    cudaFree(h->d_repl_time); cudaFree(h->d_cext); cudaFree(h->d_xd); cudaFree(h->d_nat); cudaFree(h->d_bits);
    cudaFree(h->d_crot); cudaFree(h->d_iq_tab); cudaFree(h->d_iq_thr); cudaFree(h->d_cells_seg); cudaFree(h->d_sched); cudaFree(h->d_chalo); cudaFree(h->d_sv); cudaFree(h->d_wipe); cudaFree(h->d_xg); cudaFree(h->d_code_w); cudaFree(h->d_cells); cudaFree(h->d_peaks);
    cudaFreeHost(h->h_bits); cudaFreeHost(h->h_sv); cudaFreeHost(h->h_peaks);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 8; i++) if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
}

// device, streams and events of a handle (both modes).  Fails loudly when there is no usable sm_100 device:
// the library has no CPU path.
static int open_device(gpsacq *h)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        h->err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libgpsacq has no CPU fallback";
        return GPSACQ_ECUDA;
    }
    if (h->cfg.device >= 0) { CUDA_TRY(h, cudaSetDevice(h->cfg.device)); }
    CUDA_TRY(h, cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) { h->err = "libgpsacq is built for sm_100a (B200) only"; return GPSACQ_ECUDA; }
    h->sm_count = prop.multiProcessorCount;
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    for (int i = 0; i < 4; i++) CUDA_TRY(h, cudaEventCreate(&h->ev[i]));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    return GPSACQ_OK;
}

static int create_impl(gpsacq *h)
{
    const gpsacq_cfg &c = h->cfg;
    if (!(c.fs > 0) || !(c.fc >= 0) || !(c.max_fo >= 0)) { h->err = "fc, fs, max_fo must be positive"; return GPSACQ_EINVAL; }
    h->n = c.fft_len ? c.fft_len : GPSACQ_FFT_LEN;
    if (h->n != GPSACQ_FFT_LEN) { h->err = "only fft_len = 40000 (FFT_LEN, c/gps_offline.h:15) is supported"; return GPSACQ_EINVAL; }
    if (!(c.fs / 1000.0 <= (double)h->n) || !(c.max_fo * (double)h->n / c.fs < (double)(1 << 24))) {
        h->err = "FS/1000 must not exceed fft_len and max_fo*fft_len/FS must be a sane bin count"; return GPSACQ_EINVAL;
    }
    h->w = (int)ceil(c.fs / 1000.0);                          // for (i=0; i<FS/1000; i++)  (:190)
    h->dmax = (int)(c.max_fo * (double)h->n / c.fs);          // C truncation, (:176)
    h->ndop = 2 * h->dmax + 1;
    h->chunk_samples = ((h->n + 4095) / 4096) * 4096;         // whole 512-byte packets (:129,:135-141)
    h->chunk_bytes = h->chunk_samples / 8;
    h->cap = c.max_blocks > 0 ? c.max_blocks : 512;
    {   // chunks per launch.  With the ticket scheduler of the cell kernel (ga_kernels.cuh) the CTAs of a launch work on a
        // narrow window of neighbouring cells whatever its length, so a whole batch goes out as ONE launch triple
        // (measured: 8.90 M corr/s, against 8.74 M cut into 128-chunk launches).  Round-robin cells (GPSACQ_STATIC_SCHED)
        // drift apart in long launches and want 128-chunk launches whose block spectra stay in L2.
        const char *e = getenv("GPSACQ_SUB_BLOCKS");
        h->sub_blocks = (e && atoi(e) > 0) ? atoi(e) : (getenv("GPSACQ_STATIC_SCHED") ? 128 : h->cap);
    }
    if (h->w <= G4000::N2) { h->gid = GID_4000; h->n1 = G4000::N1; h->n2 = G4000::N2; }
    else if (h->w <= G8000::N2) { h->gid = GID_8000; h->n1 = G8000::N1; h->n2 = G8000::N2; }
    else { h->gid = GID_10000; h->n1 = G10000::N1; h->n2 = G10000::N2; }
    // FS above 10 MHz (the reference takes any FS with FS/1000 <= FFT_LEN, :190): the window is covered by
    // nseg = ceil(W / 10000) output segments of the 4 x 10000 geometry (ga_kernels.cuh, c_ktab_seg)
    h->nseg = (h->w + h->n2 - 1) / h->n2;
    if (h->nseg > MAX_SEG) { h->err = "search window longer than fft_len"; return GPSACQ_EINVAL; }
    if (h->dmax >= h->n2) { h->err = "max_fo too large for this fft_len"; return GPSACQ_EINVAL; }

    { const int rc_dev = open_device(h); if (rc_dev) return rc_dev; }

    const size_t n = (size_t)h->n, cap = (size_t)h->cap;
    CUDA_TRY(h, cudaMalloc(&h->d_tw, n * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_lo, (size_t)h->chunk_samples));
    CUDA_TRY(h, cudaMalloc(&h->d_chip_idx, n * sizeof(unsigned short)));
    CUDA_TRY(h, cudaMalloc(&h->d_blend_a, n * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_blend_b, n * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_repl_time, 32 * n * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_cext, 32 * 2 * n * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_xd, cap * n * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_nat, n * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_bits, cap * (size_t)h->chunk_bytes));
    CUDA_TRY(h, cudaMalloc(&h->d_sv, cap * sizeof(int)));
    CUDA_TRY(h, cudaMalloc(&h->d_cells, cap * (size_t)h->ndop * sizeof(CellStat)));
    if (h->nseg > 1) CUDA_TRY(h, cudaMalloc(&h->d_cells_seg, cap * (size_t)h->ndop * h->nseg * sizeof(CellStat)));
    CUDA_TRY(h, cudaMalloc(&h->d_peaks, cap * sizeof(Peak)));
    if (!getenv("GPSACQ_STATIC_SCHED")) {       // (A/B knob: round-robin cells instead of the ticket queue)
        CUDA_TRY(h, cudaMalloc(&h->d_sched, 2 * sizeof(int)));
        CUDA_TRY(h, cudaMemset(h->d_sched, 0, 2 * sizeof(int)));
    }
    CUDA_TRY(h, cudaMallocHost(&h->h_bits, cap * (size_t)h->chunk_bytes));
    CUDA_TRY(h, cudaMallocHost(&h->h_sv, cap * sizeof(int)));
    CUDA_TRY(h, cudaMallocHost(&h->h_peaks, cap * sizeof(Peak)));

    // constant tables
    std::vector<cf> tw = make_tw(h->n);
    CUDA_TRY(h, cudaMemcpy(h->d_tw, tw.data(), n * sizeof(cf), cudaMemcpyHostToDevice));
    int rc = h->gid == GID_4000 ? upload_const_t<G4000, GID_4000>(h)
           : h->gid == GID_8000 ? upload_const_t<G8000, GID_8000>(h) : upload_const_t<G10000, GID_10000>(h);
    if (rc) return rc;
    for (int m = 0; h->nseg > 1 && m < h->nseg; m++) {
        std::vector<cf> km = make_ktab_seg<G10000>(m);
        CUDA_TRY(h, cudaMemcpyToSymbol(c_ktab_seg, km.data(), km.size() * sizeof(cf), (size_t)m * KTAB_MAX * sizeof(cf)));
    }
    std::vector<unsigned char> lo;
    build_lo_table(c.fc, c.fs, h->chunk_samples, lo);
    CUDA_TRY(h, cudaMemcpy(h->d_lo, lo.data(), lo.size(), cudaMemcpyHostToDevice));
    if (!getenv("GPSACQ_FWD_NOLUT")) {      // (A/B knob: per-sample unpack + multiply-add instead of the group tables)
        std::vector<unsigned int> lm((size_t)h->n2, 0u);
        for (int n2 = 0; n2 < h->n2; n2++)
            for (int n1 = 0; n1 < h->n1; n1++) lm[n2] |= (unsigned)(lo[(size_t)h->n2 * n1 + n2] & 3) << (2 * n1);
        CUDA_TRY(h, cudaMalloc(&h->d_lomask, lm.size() * sizeof(unsigned int)));
        CUDA_TRY(h, cudaMemcpy(h->d_lomask, lm.data(), lm.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    }
    std::vector<unsigned short> ci; std::vector<float> ba, bb;
    build_code_nco(c.fs_replica > 0 ? c.fs_replica : c.fs, h->n, ci, ba, bb);     // SearchInit()-time FS (:76)
    CUDA_TRY(h, cudaMemcpy(h->d_chip_idx, ci.data(), n * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_blend_a, ba.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_blend_b, bb.data(), n * sizeof(float), cudaMemcpyHostToDevice));

    rc = setup_cells(h);
    if (rc) return rc;

    // replicas: time domain, then forward FFT into the decimated + doubled layout
    SatTaps taps;
    for (int sv = 0; sv < 32; sv++) { taps.t0[sv] = kTaps[sv][0]; taps.t1[sv] = kTaps[sv][1]; }
    replica_time_kernel<<<32, 256, 0, h->stream>>>(taps, h->d_chip_idx, h->d_blend_a, h->d_blend_b, h->n, h->d_repl_time);
    CUDA_TRY(h, cudaGetLastError());
    rc = launch_fwd(h, 1, 32, nullptr, h->d_cext);
    if (rc) return rc;
    if (h->use_tma) { rc = build_halo_t<G8000>(h); if (rc) return rc; }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}


// =========================================================================================
// GRID mode (include/gpsacq.h GPSACQ_MODE_GRID; semantics in SURVEY.md App. E)
// =========================================================================================
template <class H, int T, int NW, int HID> struct GridOps {
    static int setup(gpsacq *h)
    {
        auto kern = grid_cell_kernel<H, T, NW, HID>;
        h->cell_smem = (int)(H::SMEM_ELEMS * sizeof(cf));
        h->cell_threads = T; h->cell_nw = NW;
        CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cell_smem));
        const int want = CELL_MINB * (h->cell_smem + 2048);
        int pct = (int)((want * 100LL + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        h->cell_ctas = CELL_MINB * h->sm_count;      // TMEM kernels: see setup_cells_t
        auto fk = fwd_grid_kernel<H, FWD_T, HID>;
        CUDA_TRY(h, cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cell_smem));
        CUDA_TRY(h, cudaFuncSetAttribute(fwd_kernel<H, FWD_T, 1, HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cell_smem));
        return 0;
    }
    static int fwd_blocks(gpsacq *h, size_t n_blocks, const unsigned char *d_bits)
    {
        const int items = h->n_base > 0 ? h->n_base : h->ndop, item_dmax = h->n_base > 0 ? 0 : h->dmax;     // see PfaOps::fwd_blocks
        fwd_grid_kernel<H, FWD_T, HID><<<(unsigned)(n_blocks * items * H::N1), FWD_T, h->cell_smem, h->stream>>>(
            d_bits, h->block_bytes, h->d_lo, h->d_wipe, h->w, items, item_dmax, h->wipe_m, h->d_tw, h->d_xg);
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
    static int cells(gpsacq *h, size_t n_acq)
    {
        const int n_cells = (int)(n_acq * 32 * (size_t)h->ndop);
        const int grid = std::min(n_cells, h->cell_ctas);
        grid_cell_kernel<H, T, NW, HID><<<grid, T, h->cell_smem, h->stream>>>(h->d_xg, h->d_cext, h->d_tw, n_cells, h->ndop,
                                                                            h->kblocks, h->w, h->dmax, h->n_base, h->d_cells, h->d_sched);
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
    static int replicas(gpsacq *h) { return launch_fwd_t<H, 1, HID>(h, 32, nullptr, h->d_cext); }
    static int rotate(gpsacq *) { return 0; }
    static int consts(gpsacq *h) { return upload_const_t<H, HID>(h); }
};

template <class G, int T, int MINB, bool MULTI> struct PfaOps {
    static int setup(gpsacq *h)
    {
        auto kern = pfa_cell_kernel<G, T, MINB, MULTI>;
        h->cell_smem = (int)(G::SMEM_ELEMS * sizeof(cf));
        h->cell_threads = T; h->cell_nw = G::RC;
        CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cell_smem));
        const int want = MINB * (h->cell_smem + 2048);
        int pct = (int)((want * 100LL + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 88) pct = 100;          // carveout steps are coarse: ask for everything when close
        CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        int per_sm = 0;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, h->cell_smem));
        if (per_sm < 1) { h->err = "GRID cell kernel does not fit on an SM"; return GPSACQ_ECUDA; }
        if (MULTI && per_sm < MINB) per_sm = MINB;      // TMEM kernels: the occupancy query is conservative (setup_cells_t)
        h->cell_ctas = per_sm * h->sm_count;
        typedef typename G::Fwd F;
        const int fsm = (int)(F::SMEM_ELEMS * sizeof(cf));
        CUDA_TRY(h, cudaFuncSetAttribute(pfa_fwd_kernel<G, FWD_T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm));
        CUDA_TRY(h, cudaFuncSetAttribute(pfa_fwd_kernel<G, FWD_T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm));
        return 0;
    }
    static int fwd_blocks(gpsacq *h, size_t n_blocks, const unsigned char *d_bits)
    {
        typedef typename G::Fwd F;
        // one transform per (block, Doppler bin) -- or, when bins repeat every n_base = R bins up to a spectral
        // rotation, only for the bins 0..R-1 (pfa_cell_kernel rotates)
        const int items = h->n_base > 0 ? h->n_base : h->ndop, item_dmax = h->n_base > 0 ? 0 : h->dmax;
        pfa_fwd_kernel<G, FWD_T, 0><<<(unsigned)(n_blocks * items), FWD_T, F::SMEM_ELEMS * sizeof(cf), h->stream>>>(
            d_bits, h->block_bytes, h->d_lo, h->d_wipe, items, item_dmax, h->wipe_m, nullptr, h->d_xg);
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
    static int cells(gpsacq *h, size_t n_acq)
    {
        const int n_cells = (int)(n_acq * 32 * (size_t)h->ndop);
        const int grid = std::min(n_cells, h->cell_ctas);
        pfa_cell_kernel<G, T, MINB, MULTI><<<grid, T, h->cell_smem, h->stream>>>(h->d_xg, h->d_crot ? h->d_crot : h->d_cext, n_cells, h->ndop,
                                                                                 h->dmax, h->n_base, h->q_min, h->kblocks, h->d_cells, h->d_sched);
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
    static int replicas(gpsacq *h)
    {
        typedef typename G::Fwd F;
        pfa_fwd_kernel<G, FWD_T, 1><<<32, FWD_T, F::SMEM_ELEMS * sizeof(cf), h->stream>>>(
            nullptr, 0, nullptr, nullptr, 1, 0, 1, h->d_code_w, h->d_cext);
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
    static int consts(gpsacq *) { return 0; }            // no twiddle tables: Good-Thomas
    static int rotate(gpsacq *h)                         // the replica spectra rotated by -q, q_min <= q < q_min + n_q
    {
        pfa_rotate_replicas_kernel<G><<<dim3(8, (unsigned)h->n_q, 32), 256, 0, h->stream>>>(h->d_cext, h->q_min, h->d_crot);
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
};

#define PFA_DISPATCH(h, CALL)                                                                               \
    (h->gid == PID_5456   ? (h->kblocks > 1 ? PfaOps<P5456, PFA_T_5456, PFA_B_5456, true>::CALL : PfaOps<P5456, PFA_T_5456_K1, PFA_B_5456_K1, false>::CALL) \
     : h->gid == PID_8184 ? (h->kblocks > 1 ? PfaOps<P8184, PFA_T_8184, PFA_B_8184, true>::CALL : PfaOps<P8184, PFA_T_8184_K1, PFA_B_8184, false>::CALL) \
                          : (h->kblocks > 1 ? PfaOps<P2800, PFA_T_2800, PFA_B_2800, true>::CALL : PfaOps<P2800, PFA_T_2800, PFA_B_2800, false>::CALL))

#define GRID_DISPATCH(h, CALL)                                                                              \
    (h->gid >= PID_5456    ? PFA_DISPATCH(h, CALL)                                                          \
     : h->gid == XID_10000 ? GridOps<X10000, 256, 20, XID_10000>::CALL                                      \
     : h->gid == XID_8000  ? GridOps<X8000, 448, 20, XID_8000>::CALL                                        \
     : h->gid == XID_4000  ? GridOps<X4000, 256, 10, XID_4000>::CALL                                        \
     : h->gid == XID_4096  ? GridOps<X4096, 256, 16, XID_4096>::CALL                                        \
     : h->gid == XID_2048  ? GridOps<X2048, 128, 8, XID_2048>::CALL                                         \
     : h->gid == HID_4000    ? (h->cell_nw == 7 ? GridOps<H4000, 256, 7, HID_4000>::CALL : GridOps<H4000, 256, 10, HID_4000>::CALL)      \
     : h->gid == HID_6400  ? (h->cell_nw == 14 ? GridOps<H6400, 448, 14, HID_6400>::CALL : GridOps<H6400, 448, 16, HID_6400>::CALL)    \
     : h->gid == HID_8000  ? (h->cell_nw == 14 ? GridOps<H8000, 448, 14, HID_8000>::CALL : GridOps<H8000, 448, 20, HID_8000>::CALL)    \
                           : (h->cell_nw == 17 ? GridOps<H10000, 256, 17, HID_10000>::CALL : GridOps<H10000, 256, 20, HID_10000>::CALL))

// native transform available for this block length?  (GPSACQ_GRID_EMBED=1 forces the embedding, for A/B runs)
static int pfa_gid_for(int w, int dmax)
{
    const char *e = getenv("GPSACQ_GRID_EMBED");
    if (e && *e && *e != '0') return -1;
    if ((long long)dmax * w >= (1LL << 31)) return -1;          // GridSrc32 index arithmetic
    return w == P5456::W ? PID_5456 : w == P8184::W ? PID_8184 : w == P2800::W ? PID_2800 : -1;
}

static int create_grid(gpsacq *h)
{
    const gpsacq_cfg &c = h->cfg;
    if (!(c.fs > 0) || !(c.fc >= 0) || !(c.max_fo >= 0) || !(c.doppler_step > 0)) { h->err = "GRID mode needs fs, doppler_step > 0"; return GPSACQ_EINVAL; }
    const double wd = c.fs / 1000.0, md = c.fs / c.doppler_step;
    if (!(wd < 1e6) || !(md < (double)(1 << 28)) || !(c.max_fo / c.doppler_step < (double)(1 << 24))) {
        h->err = "GRID mode: FS/1000, FS/doppler_step or max_fo/doppler_step out of range"; return GPSACQ_EINVAL;
    }
    h->w = (int)llround(wd);
    h->wipe_m = (int)llround(md);
    if (fabs(wd - h->w) > 1e-9 || h->w % 8) { h->err = "GRID mode needs FS/1000 to be an integer multiple of 8 samples"; return GPSACQ_EINVAL; }
    if (fabs(md - h->wipe_m) > 1e-6 * md) { h->err = "GRID mode needs FS/doppler_step to be an integer"; return GPSACQ_EINVAL; }
    h->kblocks = c.noncoh_blocks > 0 ? c.noncoh_blocks : 1;
    h->step = c.doppler_step;
    h->dmax_full = (int)floor(c.max_fo / c.doppler_step + 1e-9);
    const int ndop_full = 2 * h->dmax_full + 1;
    h->dop_first = 0;
    h->ndop = ndop_full;
    if (c.dop_count != 0) {          // a shard of the Doppler grid
        if (c.dop_first < 0 || c.dop_count < 0 || (long long)c.dop_first + c.dop_count > ndop_full) { h->err = "dop_first/dop_count outside the Doppler grid"; return GPSACQ_EINVAL; }
        h->dop_first = c.dop_first;
        h->ndop = c.dop_count;
    }
    h->dmax = h->dmax_full - h->dop_first;      // kernels form the absolute bin as index - h->dmax
    h->block_bytes = h->w / 8;
    h->chunk_bytes = h->block_bytes * h->kblocks;
    h->chunk_samples = h->w;
    const int pgid = pfa_gid_for(h->w, h->dmax_full);
    const char *force_embed = getenv("GPSACQ_GRID_EMBED");
    const bool exact_ok = !(force_embed && *force_embed && *force_embed != '0');
    if (pgid >= 0) { h->gid = pgid; h->n1 = 1; h->n2 = h->w; h->cell_nw = 0; }
    else if (exact_ok && h->w == X10000::N2) { h->gid = XID_10000; h->n1 = 1; h->n2 = h->w; h->cell_nw = X10000::RC; }
    else if (exact_ok && h->w == X8000::N2) { h->gid = XID_8000; h->n1 = 1; h->n2 = h->w; h->cell_nw = X8000::RC; }
    else if (exact_ok && h->w == X4000::N2) { h->gid = XID_4000; h->n1 = 1; h->n2 = h->w; h->cell_nw = X4000::RC; }
    else if (exact_ok && h->w == X4096::N2) { h->gid = XID_4096; h->n1 = 1; h->n2 = h->w; h->cell_nw = X4096::RC; }
    else if (exact_ok && h->w == X2048::N2) { h->gid = XID_2048; h->n1 = 1; h->n2 = h->w; h->cell_nw = X2048::RC; }
    else if (h->w <= H4000::N2) { h->gid = HID_4000; h->n1 = 2; h->n2 = H4000::N2; h->cell_nw = h->w <= 7 * H4000::OUT_STRIDE ? 7 : 10; }
    else if (h->w <= H6400::N2) { h->gid = HID_6400; h->n1 = 2; h->n2 = H6400::N2; h->cell_nw = h->w <= 14 * H6400::OUT_STRIDE ? 14 : 16; }
    else if (h->w <= H8000::N2) { h->gid = HID_8000; h->n1 = 2; h->n2 = H8000::N2; h->cell_nw = h->w <= 14 * H8000::OUT_STRIDE ? 14 : 20; }
    else if (h->w <= H10000::N2) { h->gid = HID_10000; h->n1 = 2; h->n2 = H10000::N2; h->cell_nw = h->w <= 17 * H10000::OUT_STRIDE ? 17 : 20; }
    else { h->err = "sampling rates above 10 MHz are not supported yet"; return GPSACQ_EINVAL; }
    h->n = h->n1 * h->n2;                                      // L

    { const int rc_dev = open_device(h); if (rc_dev) return rc_dev; }

    // Doppler bins R = 1000/step apart differ by exactly one DFT bin of the 1 ms block (M = R*W): the native path
    // transforms only the bins 0..R-1 of each block and rotates (GPSACQ_GRID_NOSHARE=1 turns this off, for A/B runs).
    // Decided from the FULL grid, not from this handle's shard: a shard of <= R bins must use the same arithmetic as
    // the unsharded handle, or near-tie winners could differ between a sharded and an unsharded search.
    h->n_base = 0;
    {
        const char *ns = getenv("GPSACQ_GRID_NOSHARE");
        const int r = h->w > 0 ? h->wipe_m / h->w : 0;
        if ((h->gid >= PID_5456 || h->n1 == 1) && !(ns && *ns && *ns != '0') && r >= 1 && (long long)r * h->w == h->wipe_m && r < ndop_full) h->n_base = r;
    }
    // batch capacity: keep the block spectra of one batch under ~3 GB
    const size_t per_acq = (size_t)h->kblocks * (h->n_base > 0 ? h->n_base : h->ndop) * h->n * sizeof(cf);
    size_t cap = (size_t)3 << 30;
    cap = std::max<size_t>(1, cap / per_acq);
    if (c.max_blocks > 0) cap = std::min<size_t>(cap, (size_t)c.max_blocks);
    cap = std::min<size_t>(cap, 4096);
    h->cap_acq = (int)cap;
    h->cap = h->cap_acq;

    const size_t L = (size_t)h->n, W = (size_t)h->w;
    CUDA_TRY(h, cudaMalloc(&h->d_tw, L * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_wipe, (size_t)h->wipe_m * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_lo, W));
    CUDA_TRY(h, cudaMalloc(&h->d_chip_idx, W * sizeof(unsigned short)));
    CUDA_TRY(h, cudaMalloc(&h->d_blend_a, W * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_blend_b, W * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_code_w, 32 * W * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_repl_time, 32 * L * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&h->d_cext, 32 * 2 * L * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_xg, cap * per_acq));
    CUDA_TRY(h, cudaMalloc(&h->d_nat, L * sizeof(cf)));
    CUDA_TRY(h, cudaMalloc(&h->d_bits, cap * (size_t)h->chunk_bytes));
    CUDA_TRY(h, cudaMalloc(&h->d_cells, cap * 32 * (size_t)h->ndop * sizeof(CellStat)));
    if (!getenv("GPSACQ_STATIC_SCHED")) {       // ticket counter of pfa_cell_kernel's work queue (A/B knob as in REF mode)
        CUDA_TRY(h, cudaMalloc(&h->d_sched, 2 * sizeof(int)));
        CUDA_TRY(h, cudaMemset(h->d_sched, 0, 2 * sizeof(int)));
    }
    CUDA_TRY(h, cudaMalloc(&h->d_peaks, cap * 32 * sizeof(Peak)));
    CUDA_TRY(h, cudaMallocHost(&h->h_bits, cap * (size_t)h->chunk_bytes));
    CUDA_TRY(h, cudaMallocHost(&h->h_peaks, cap * 32 * sizeof(Peak)));

    std::vector<cf> tw = make_tw(h->n);
    CUDA_TRY(h, cudaMemcpy(h->d_tw, tw.data(), L * sizeof(cf), cudaMemcpyHostToDevice));
    std::vector<cf> wipe = make_tw(h->wipe_m);                 // exp(+2 pi i k/M); the wipe-off uses the conjugate
    for (auto &v : wipe) v.y = -v.y;
    CUDA_TRY(h, cudaMemcpy(h->d_wipe, wipe.data(), wipe.size() * sizeof(cf), cudaMemcpyHostToDevice));
    int rc = GRID_DISPATCH(h, consts(h));
    if (rc) return rc;
    std::vector<unsigned char> lo;
    build_lo_table(c.fc, c.fs, h->w, lo);                      // phase restarts at every block (App. E)
    CUDA_TRY(h, cudaMemcpy(h->d_lo, lo.data(), lo.size(), cudaMemcpyHostToDevice));
    std::vector<unsigned short> ci; std::vector<float> ba, bb;
    build_code_nco(c.fs, h->w, ci, ba, bb);                    // one code period, same NCO + blend as SearchInit()
    CUDA_TRY(h, cudaMemcpy(h->d_chip_idx, ci.data(), W * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_blend_a, ba.data(), W * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_blend_b, bb.data(), W * sizeof(float), cudaMemcpyHostToDevice));
    rc = GRID_DISPATCH(h, setup(h));
    if (rc) return rc;

    SatTaps taps;
    for (int sv = 0; sv < 32; sv++) { taps.t0[sv] = kTaps[sv][0]; taps.t1[sv] = kTaps[sv][1]; }
    replica_time_kernel<<<32, 256, 0, h->stream>>>(taps, h->d_chip_idx, h->d_blend_a, h->d_blend_b, h->w, h->d_code_w);
    CUDA_TRY(h, cudaGetLastError());
    if (h->gid < PID_5456) {       // embedding: two code periods, zero-padded to L (native transforms read d_code_w directly)
        grid_replica_extend_kernel<<<dim3(16, 32), 256, 0, h->stream>>>(h->d_code_w, h->w, h->n, (float)((double)h->w / (double)h->n), h->d_repl_time);
        CUDA_TRY(h, cudaGetLastError());
    }
    rc = GRID_DISPATCH(h, replicas(h));
    if (rc) return rc;
    if (h->n_base > 0 && h->gid >= PID_5456) {
        // bins of this handle: d in [-dmax, -dmax + ndop); q = floor(d / R)
        const int d_lo = -h->dmax, d_hi = -h->dmax + h->ndop - 1;
        auto fdiv = [](int a, int b) { int q = a / b; return (a % b != 0 && (a < 0)) ? q - 1 : q; };
        h->q_min = fdiv(d_lo, h->n_base);
        h->n_q = fdiv(d_hi, h->n_base) - h->q_min + 1;
        CUDA_TRY(h, cudaMalloc(&h->d_crot, (size_t)h->n_q * 32 * W * sizeof(cf)));
        rc = GRID_DISPATCH(h, rotate(h));
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

static int acquire_device_impl(gpsacq *h, const uint8_t *d_bits, size_t n_acq, gpsacq_peak *d_out)
{
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventRecord(h->ev[0], h->stream));
    int rc = GRID_DISPATCH(h, fwd_blocks(h, n_acq * (size_t)h->kblocks, d_bits));
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev[1], h->stream));
    rc = GRID_DISPATCH(h, cells(h, n_acq));
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev[2], h->stream));
    const size_t nrec = n_acq * 32;
    best_kernel<<<(unsigned)((nrec + 3) / 4), 128, 0, h->stream>>>(h->d_cells, nullptr, (int)nrec, h->ndop, h->dmax, h->w, (Peak *)d_out);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
    h->have_batch = true;
    h->last_blocks = nrec;
    return GPSACQ_OK;
}

static int iq8_convert_piece(gpsacq *h, const unsigned char *d_iq, size_t n, size_t n0, int format, const long long *d_sums,
                             size_t n_total, double shift_hz, double fs, unsigned char *d_bits);

// ---- ABI -------------------------------------------------------------------------------
extern "C" {

int gpsacq_create(const gpsacq_cfg *cfg, gpsacq_t **out)
{
    DeviceGuard guard;
    if (!cfg || !out) { g_create_error = "null argument"; return GPSACQ_EINVAL; }
    *out = nullptr;
    gpsacq *h = new (std::nothrow) gpsacq();
    if (!h) { g_create_error = "out of host memory"; return GPSACQ_ENOMEM; }
    h->cfg = *cfg;
    h->mode = cfg->mode;
    if (cfg->mode != GPSACQ_MODE_REF && cfg->mode != GPSACQ_MODE_GRID) { g_create_error = "unknown mode"; delete h; return GPSACQ_EINVAL; }
    int rc = cfg->mode == GPSACQ_MODE_GRID ? create_grid(h) : create_impl(h);
    if (rc) {
        g_create_error = h->err;
        free_all(h);
        delete h;
        return rc;
    }
    *out = h;
    return GPSACQ_OK;
}

void gpsacq_destroy(gpsacq_t *h)
{
    DeviceGuard guard;
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_all(h);
    delete h;
}

const char *gpsacq_last_error(const gpsacq_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int gpsacq_get_info(const gpsacq_t *h, gpsacq_info *info)
{
    if (!h || !info) return GPSACQ_EINVAL;
    memset(info, 0, sizeof *info);
    info->abi_version = GPSACQ_ABI_VERSION;
    info->fft_len = h->n; info->n1 = h->n1; info->n2 = h->n2;
    info->window = h->w; info->dmax = h->mode == GPSACQ_MODE_GRID ? h->dmax_full : h->dmax; info->n_doppler = h->ndop;
    info->dop_first = h->mode == GPSACQ_MODE_GRID ? h->dop_first : 0;
    info->n_doppler_full = h->mode == GPSACQ_MODE_GRID ? 2 * h->dmax_full + 1 : h->ndop;
    info->chunk_bytes = h->chunk_bytes; info->max_blocks = h->cap;
    info->device = h->device; info->sm_count = h->sm_count;
    info->cell_ctas = h->cell_ctas; info->cell_threads = h->cell_threads; info->cell_smem_bytes = h->cell_smem;
    info->bytes_per_corr = 2LL * (h->mode == GPSACQ_MODE_GRID ? h->w : h->n) * 8 + 16;
    info->blocks_per_launch = h->mode == GPSACQ_MODE_GRID ? 0 : (h->last_launch_blocks > 0 ? h->last_launch_blocks : h->sub_blocks);
    info->mode = h->mode; info->noncoh_blocks = h->mode == GPSACQ_MODE_GRID ? h->kblocks : 1;
    info->block_bytes = h->mode == GPSACQ_MODE_GRID ? h->block_bytes : h->chunk_bytes;
    info->max_acq = h->mode == GPSACQ_MODE_GRID ? h->cap_acq : 0;
    info->doppler_step = h->mode == GPSACQ_MODE_GRID ? h->step : h->cfg.fs / h->n;
    return GPSACQ_OK;
}

int gpsacq_set_stream(gpsacq_t *h, void *cuda_stream)
{
    if (!h) return GPSACQ_EINVAL;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return GPSACQ_OK;
}

int gpsacq_synchronize(gpsacq_t *h)
{
    if (!h) return GPSACQ_EINVAL;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return GPSACQ_OK;
}

// one (sub-)batch on the handle's stream; workspace slots [off, off + n_blocks)
// (`timed`: record the stage events -- the last slice of a sliced host batch, see gpsacq_stage_times)
// The batch is cut into launches of sub_blocks chunks (forward transform, cells, best-over-Doppler each): the block
// spectra a forward launch writes (320 KB per chunk) are then still in the 126 MB L2 when the cell launch reads them,
// instead of making a round trip through HBM as they do when thousands of chunks are transformed first.
// Stage events bracket the LAST launch triple.
static int search_device_impl(gpsacq *h, const uint8_t *d_bits, size_t n_blocks, const int32_t *d_sv, gpsacq_peak *d_out, size_t off,
                              bool last)
{
    const size_t sub = (size_t)h->sub_blocks;
    for (size_t done = 0; done < n_blocks; done += sub) {
        const size_t n = std::min(sub, n_blocks - done), o = off + done;
        const bool timed = last && done + n == n_blocks;
        const int32_t *sv = d_sv ? d_sv + done : nullptr;
        if (timed) CUDA_TRY(h, cudaEventRecord(h->ev[0], h->stream));
        int rc = launch_fwd(h, 0, n, d_bits + done * (size_t)h->chunk_bytes, h->d_xd + o * (size_t)h->n);
        if (rc) return rc;
        if (timed) CUDA_TRY(h, cudaEventRecord(h->ev[1], h->stream));
        rc = launch_cells(h, n, sv, o, (int)done);
        if (rc) return rc;
        if (timed) CUDA_TRY(h, cudaEventRecord(h->ev[2], h->stream));
        best_kernel<<<(unsigned)((n + 3) / 4), 128, 0, h->stream>>>(h->d_cells + o * (size_t)h->ndop, sv, (int)n, h->ndop, h->dmax, h->w, (Peak *)d_out + done, (int)done);
        CUDA_TRY(h, cudaGetLastError());
        if (timed) CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
        h->last_launch_blocks = (int)n;
    }
    return GPSACQ_OK;
}

int gpsacq_search_blocks_device(gpsacq_t *h, const uint8_t *d_bits, size_t n_blocks, const int32_t *d_sv, gpsacq_peak *d_out)
{
    DeviceGuard guard;
    if (!h || !d_bits || !d_out) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_REF) { h->err = "gpsacq_search_blocks needs a GPSACQ_MODE_REF handle"; return GPSACQ_EINVAL; }
    if (n_blocks == 0) return GPSACQ_OK;
    if (n_blocks > (size_t)h->cap) { h->err = "n_blocks exceeds max_blocks"; return GPSACQ_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int rc = search_device_impl(h, d_bits, n_blocks, d_sv, d_out, 0, true);
    if (rc) return rc;
    h->have_batch = true;
    h->last_blocks = n_blocks;
    return GPSACQ_OK;
}

// Host-buffer search.  A batch is cut into up to 4 slices: while the GPU searches slice i the host stages slice
// i+1 into pinned memory and the copy engine moves it (copy stream + events), so that only the first slice's
// transfer is exposed.  Results come back with one device->host copy per batch.
int gpsacq_search_blocks(gpsacq_t *h, const uint8_t *bits, size_t n_blocks, const int32_t *sv_of_block, gpsacq_peak *out)
{
    DeviceGuard guard;
    if (!h || (!bits && n_blocks) || (!out && n_blocks)) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_REF) { h->err = "gpsacq_search_blocks needs a GPSACQ_MODE_REF handle"; return GPSACQ_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t cb = (size_t)h->chunk_bytes;
    for (size_t done = 0; done < n_blocks;) {
        const size_t nb = std::min((size_t)h->cap, n_blocks - done);
        for (size_t b = 0; b < nb; b++) {
            const int sv = sv_of_block ? sv_of_block[done + b] : (int)((done + b) % GPSACQ_NUM_SATS);
            if (sv < 0 || sv >= GPSACQ_NUM_SATS) { h->err = "sv_of_block entry out of range 0..31"; return GPSACQ_EINVAL; }
            h->h_sv[b] = sv;
        }
        // Slices: 1/8, 1/8, 1/4, 1/2 of the batch -- the first one small, so that the GPU starts early and only its
        // staging + transfer is exposed.  A caller buffer that is page-locked already (cudaHostAlloc / cudaHostRegister /
        // torch pin_memory) is handed to the copy engine as it is; pageable memory goes through the pinned staging buffer.
        size_t cut[5] = {0, nb / 8, nb / 4, nb / 2, nb};
        size_t n_slices = 4;
        if (nb < 256) { n_slices = nb >= 64 ? 2 : 1; cut[1] = n_slices == 2 ? nb / 2 : nb; cut[2] = nb; }
        bool pinned = false;
        {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, bits + done * cb) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
            else cudaGetLastError();          // (plain malloc memory makes older runtimes return an error: not pinned)
        }
        // the previous batch's kernels may still read d_bits / d_sv: the copy stream waits for them
        CUDA_TRY(h, cudaEventRecord(h->ev_done, h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done, 0));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_sv, h->h_sv, nb * sizeof(int), cudaMemcpyHostToDevice, h->copy_stream));
        for (size_t i = 0; i < n_slices; i++) {
            const size_t lo = cut[i], n = cut[i + 1] - lo;
            if (n == 0) continue;
            const unsigned char *src = bits + (done + lo) * cb;
            if (!pinned) { memcpy(h->h_bits + lo * cb, src, n * cb); src = h->h_bits + lo * cb; }
            CUDA_TRY(h, cudaMemcpyAsync(h->d_bits + lo * cb, src, n * cb, cudaMemcpyHostToDevice, h->copy_stream));
            CUDA_TRY(h, cudaEventRecord(h->ev_copy[i], h->copy_stream));
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_copy[i], 0));
            const int rc = search_device_impl(h, h->d_bits + lo * cb, n, h->d_sv + lo, (gpsacq_peak *)(h->d_peaks + lo), lo, i + 1 == n_slices);
            if (rc) {           // nothing may still read the pinned staging buffers when the caller sees the error
                cudaStreamSynchronize(h->copy_stream);
                cudaStreamSynchronize(h->stream);
                return rc;
            }
        }
        h->have_batch = true;
        h->last_blocks = nb;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_peaks, h->d_peaks, nb * sizeof(Peak), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        memcpy(out + done, h->h_peaks, nb * sizeof(Peak));
        done += nb;
    }
    return GPSACQ_OK;
}

int gpsacq_acquire_device(gpsacq_t *h, const uint8_t *d_bits, size_t n_acq, gpsacq_peak *d_out)
{
    DeviceGuard guard;
    if (!h || !d_bits || !d_out) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_GRID) { h->err = "gpsacq_acquire needs a GPSACQ_MODE_GRID handle"; return GPSACQ_EINVAL; }
    if (n_acq == 0) return GPSACQ_OK;
    if (n_acq > (size_t)h->cap_acq) { h->err = "n_acq exceeds the batch capacity (gpsacq_info.max_acq)"; return GPSACQ_EINVAL; }
    return acquire_device_impl(h, d_bits, n_acq, d_out);
}

int gpsacq_acquire(gpsacq_t *h, const uint8_t *bits, size_t n_acq, gpsacq_peak *out)
{
    DeviceGuard guard;
    if (!h || (!bits && n_acq) || (!out && n_acq)) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_GRID) { h->err = "gpsacq_acquire needs a GPSACQ_MODE_GRID handle"; return GPSACQ_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    for (size_t done = 0; done < n_acq;) {
        const size_t na = std::min((size_t)h->cap_acq, n_acq - done);
        memcpy(h->h_bits, bits + done * (size_t)h->chunk_bytes, na * (size_t)h->chunk_bytes);
        CUDA_TRY(h, cudaMemcpyAsync(h->d_bits, h->h_bits, na * (size_t)h->chunk_bytes, cudaMemcpyHostToDevice, h->stream));
        int rc = acquire_device_impl(h, h->d_bits, na, (gpsacq_peak *)h->d_peaks);
        if (rc) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_peaks, h->d_peaks, na * 32 * sizeof(Peak), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        memcpy(out + done * 32, h->h_peaks, na * 32 * sizeof(Peak));
        done += na;
    }
    return GPSACQ_OK;
}

int gpsacq_iq8_to_bits(gpsacq_t *h, const void *iq, size_t n_samples, int format, double shift_hz, double fs, uint8_t *bits_out)
{
    DeviceGuard guard;
    if (!h || (!iq && n_samples) || (!bits_out && n_samples) || (format != GPSACQ_IQ_U8 && format != GPSACQ_IQ_S8) || !(fs > 0))
        return GPSACQ_EINVAL;
    if (n_samples == 0) return GPSACQ_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t piece = (size_t)1 << 26;                        // complex samples per device buffer (128 MB of IQ)
    const size_t cap = std::min(piece, n_samples);
    unsigned char *d_iq = nullptr, *d_bits = nullptr;
    long long *d_sums = nullptr;
    int rc = GPSACQ_OK;
    do {
        if (cudaMalloc(&d_iq, 2 * cap) != cudaSuccess || cudaMalloc(&d_bits, (cap + 7) / 8) != cudaSuccess ||
            cudaMalloc(&d_sums, 2 * sizeof(long long)) != cudaSuccess) { h->err = "front-end: device allocation failed"; rc = GPSACQ_ENOMEM; break; }
        if (cudaMemsetAsync(d_sums, 0, 2 * sizeof(long long), h->stream) != cudaSuccess) { rc = GPSACQ_ECUDA; break; }
        // pass 1: mean over the whole capture (exact integer sums)
        for (size_t done = 0; done < n_samples && rc == GPSACQ_OK; done += cap) {
            const size_t n = std::min(cap, n_samples - done);
            if (cudaMemcpyAsync(d_iq, (const unsigned char *)iq + 2 * done, 2 * n, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { rc = GPSACQ_ECUDA; break; }
            iq8_sum_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(d_iq, n, format, d_sums);
            if (n_samples > cap && cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = GPSACQ_ECUDA; break; }
        }
        if (rc) break;
        // pass 2: shift, real part, sign, pack (the mean = d_sums / n_samples, over the whole capture)
        for (size_t done = 0; done < n_samples; done += cap) {
            const size_t n = std::min(cap, n_samples - done);
            if (n_samples > cap || done > 0)
                if (cudaMemcpyAsync(d_iq, (const unsigned char *)iq + 2 * done, 2 * n, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { rc = GPSACQ_ECUDA; break; }
            const size_t nbytes = (n + 7) / 8;
            if (iq8_convert_piece(h, d_iq, n, done, format, d_sums, n_samples, shift_hz, fs, d_bits) != GPSACQ_OK) { rc = GPSACQ_ECUDA; break; }
            if (cudaMemcpyAsync(bits_out + done / 8, d_bits, nbytes, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = GPSACQ_ECUDA; break; }
        }
    } while (0);
    if (rc == GPSACQ_ECUDA) h->err = std::string("front-end: ") + cudaGetErrorString(cudaGetLastError());
    cudaFree(d_iq); cudaFree(d_bits); cudaFree(d_sums);
    return rc;
}

}  // extern "C" (re-opened below)

// ---- 8-bit IQ front-end with device buffers; exactly periodic phase table when fc/fs is a small rational ----------
// Stream converters, second version (threshold-table iq8 -> bits, permute-based bits -> iq8; ga_frontend.cuh):
// GPSACQ_FRONTEND_V2=0/1 overrides the built-in default, for A/B runs.
#ifndef GA_FRONTEND_V2_DEFAULT
#define GA_FRONTEND_V2_DEFAULT 1
#endif
static bool frontend_v2()
{
    const char *e = getenv("GPSACQ_FRONTEND_V2");
    return (e && *e) ? atoi(e) != 0 : GA_FRONTEND_V2_DEFAULT != 0;
}

// One piece of pass 2.  d_sums = the integer sums of pass 1 (device), n_total = samples they were taken over.
static int iq8_convert_piece(gpsacq *h, const unsigned char *d_iq, size_t n, size_t n0, int format, const long long *d_sums,
                             size_t n_total, double shift_hz, double fs, unsigned char *d_bits)
{
    const size_t nbytes = (n + 7) / 8;
    unsigned long long p = 0, q = 0;
    const bool periodic = small_rational(fabs(shift_hz) / fs, p, q) && ((uintptr_t)d_iq % 16) == 0;
    if (periodic && (h->iq_tab_q != q || h->iq_tab_p != p || h->iq_tab_neg != (shift_hz < 0) || !h->d_iq_tab)) {
        std::vector<double> tab(2 * q);
        for (unsigned long long k = 0; k < q; k++) {
            const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)q;
            tab[2 * k] = (double)cosl(a); tab[2 * k + 1] = (double)(shift_hz < 0 ? -sinl(a) : sinl(a));
        }
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_iq_tab); h->d_iq_tab = nullptr;
        cudaFree(h->d_iq_thr); h->d_iq_thr = nullptr;
        CUDA_TRY(h, cudaMalloc(&h->d_iq_tab, 2 * q * sizeof(double)));
        CUDA_TRY(h, cudaMemcpy(h->d_iq_tab, tab.data(), 2 * q * sizeof(double), cudaMemcpyHostToDevice));
        if (q <= IQ8_THR_MAX_Q) CUDA_TRY(h, cudaMalloc(&h->d_iq_thr, (size_t)256 * iq8_thr_pitch((unsigned)q) * sizeof(unsigned)));
        h->iq_tab_p = p; h->iq_tab_q = q; h->iq_tab_neg = shift_hz < 0;
    }
    if (periodic && q <= IQ8_THR_MAX_Q && (n0 % 8) == 0 && frontend_v2()) {
        // thresholds from the sums (on the device: no host round trip between the passes), then the integer-only pass 2
        const unsigned pitch = iq8_thr_pitch((unsigned)q);
        const size_t smem = (size_t)256 * pitch * sizeof(unsigned);
        if (!h->iq_thr_attr) {
            CUDA_TRY(h, cudaFuncSetAttribute(iq8_to_bits_thr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             256 * (int)iq8_thr_pitch(IQ8_THR_MAX_Q) * (int)sizeof(unsigned)));
            h->iq_thr_attr = true;
        }
        iq8_thr_build_kernel<<<(unsigned)q, 256, 0, h->stream>>>(d_sums, n_total, (const double2 *)h->d_iq_tab, pitch, h->d_iq_thr);
        CUDA_TRY(h, cudaGetLastError());
        const unsigned threads = (unsigned)h->sm_count * IQ8_THR_THREADS, n_active = threads / (unsigned)q * (unsigned)q;
        iq8_to_bits_thr_kernel<<<h->sm_count, IQ8_THR_THREADS, smem, h->stream>>>(
            (const uint4 *)d_iq, n, n0, format == GPSACQ_IQ_S8 ? 0x80808080u : 0u, h->d_iq_thr, (unsigned)p, (unsigned)q, pitch, n_active,
            d_sums, n_total, (const double2 *)h->d_iq_tab, d_bits);
        CUDA_TRY(h, cudaGetLastError());
        return GPSACQ_OK;
    }
    // the double-precision kernels take the mean as an argument
    long long sums[2];
    CUDA_TRY(h, cudaMemcpyAsync(sums, d_sums, sizeof sums, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const double mi = (double)sums[0] / (double)n_total, mq = (double)sums[1] / (double)n_total;
    if (periodic) {
        const size_t nwords = (nbytes + 3) / 4;                  // one thread per 32 samples
        const size_t tab_smem = q <= 2048 ? (size_t)q * sizeof(double2) : 0;
        iq8_to_bits_table_kernel<<<(unsigned)((nwords + 255) / 256), 256, tab_smem, h->stream>>>(d_iq, n, n0, format, mi, mq, (const double2 *)h->d_iq_tab, p, q, d_bits);
    } else {
        iq8_to_bits_kernel<<<(unsigned)((nbytes + 255) / 256), 256, 0, h->stream>>>(d_iq, n, n0, format, mi, mq, shift_hz, fs, d_bits);
    }
    CUDA_TRY(h, cudaGetLastError());
    return GPSACQ_OK;
}

extern "C" int gpsacq_iq8_to_bits_device(gpsacq_t *h, const void *d_iq, size_t n_samples, int format, double shift_hz, double fs,
                                         uint8_t *d_bits_out, void *d_sums)
{
    DeviceGuard guard;
    if (!h || (!d_iq && n_samples) || (!d_bits_out && n_samples) || !d_sums || (format != GPSACQ_IQ_U8 && format != GPSACQ_IQ_S8) || !(fs > 0))
        return GPSACQ_EINVAL;
    if (n_samples == 0) return GPSACQ_OK;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemsetAsync(d_sums, 0, 2 * sizeof(long long), h->stream));
    iq8_sum_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>((const unsigned char *)d_iq, n_samples, format, (long long *)d_sums);
    CUDA_TRY(h, cudaGetLastError());
    return iq8_convert_piece(h, (const unsigned char *)d_iq, n_samples, 0, format, (const long long *)d_sums, n_samples, shift_hz, fs, d_bits_out);
}

// ---- the reverse converter: 1-bit real IF -> int8 IQ (c/conv_1bit_bin_to_hackrf_bin.cpp:29-86) -------------------------
struct ConvState { int device; double fc, fs; unsigned char *d_lo; LoCycle cyc; };
static ConvState g_conv = {-1, 0, 0, nullptr, {}};
static std::mutex g_conv_mutex;        // the cached LO table is process-wide: table (re)build + launch are one critical section

static int conv_prepare(int device, double fc, double fs, unsigned long long n_needed)
{
    if (g_conv.d_lo && g_conv.device == device && g_conv.fc == fc && g_conv.fs == fs && g_conv.cyc.mu + g_conv.cyc.lambda >= std::min<unsigned long long>(n_needed, g_conv.cyc.mu + g_conv.cyc.lambda)) return GPSACQ_OK;
    if (!(fs > 0) || !(fc >= 0) || !std::isfinite(fc) || !std::isfinite(fs)) { g_create_error = "bits_to_iq8: bad fc / fs"; return GPSACQ_EINVAL; }
    LoCycle c;
    if (!conv_lo_cycle(fc, fs, n_needed, c)) { g_create_error = "bits_to_iq8: 4*fc/fs must lie in [0,4) and the input be shorter than 2^27 samples when the LO recurrence has no short cycle"; return GPSACQ_EINVAL; }
    if (g_conv.d_lo) { cudaFree(g_conv.d_lo); g_conv.d_lo = nullptr; }
    if (cudaMalloc(&g_conv.d_lo, c.tab.size() + 16) != cudaSuccess || cudaMemset(g_conv.d_lo, 0, c.tab.size() + 16) != cudaSuccess || cudaMemcpy(g_conv.d_lo, c.tab.data(), c.tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        g_create_error = "bits_to_iq8: LO table upload failed"; return GPSACQ_ECUDA;
    }
    g_conv.device = device; g_conv.fc = fc; g_conv.fs = fs; g_conv.cyc = std::move(c);
    return GPSACQ_OK;
}

extern "C" int gpsacq_bits_to_iq8_device(int device, const uint8_t *d_bits, size_t n_bytes, size_t first_sample, double fc, double fs,
                                         int amplitude, int8_t *d_iq_out, void *cuda_stream)
{
    DeviceGuard guard;
    if ((!d_bits && n_bytes) || (!d_iq_out && n_bytes) || amplitude < 0 || amplitude > 127) { g_create_error = "bits_to_iq8: bad arguments"; return GPSACQ_EINVAL; }
    if (n_bytes == 0) return GPSACQ_OK;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_create_error = "bits_to_iq8: cudaSetDevice failed (no CPU fallback)"; return GPSACQ_ECUDA; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_create_error = "bits_to_iq8: no CUDA device (no CPU fallback)"; return GPSACQ_ECUDA; }
    std::lock_guard<std::mutex> lock(g_conv_mutex);
    const int rc = conv_prepare(dev, fc, fs, first_sample + 8ull * n_bytes);
    if (rc) return rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t want = (n_bytes + 255) / 256;
    const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)sms * 16);
    if (frontend_v2())
        bits_to_iq8_v2_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_bits, n_bytes, first_sample, g_conv.d_lo, g_conv.cyc.mu, g_conv.cyc.lambda, amplitude, (uint4 *)d_iq_out);
    else
        bits_to_iq8_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_bits, n_bytes, first_sample, g_conv.d_lo, g_conv.cyc.mu, g_conv.cyc.lambda, amplitude, (uint4 *)d_iq_out);
    if (cudaGetLastError() != cudaSuccess) { g_create_error = "bits_to_iq8: kernel launch failed"; return GPSACQ_ECUDA; }
    return GPSACQ_OK;
}

extern "C" int gpsacq_bits_to_iq8(int device, const uint8_t *bits, size_t n_bytes, size_t first_sample, double fc, double fs, int amplitude,
                                  int8_t *iq_out)
{
    DeviceGuard guard;
    if ((!bits && n_bytes) || (!iq_out && n_bytes)) { g_create_error = "bits_to_iq8: bad arguments"; return GPSACQ_EINVAL; }
    if (n_bytes == 0) return GPSACQ_OK;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_create_error = "bits_to_iq8: cudaSetDevice failed (no CPU fallback)"; return GPSACQ_ECUDA; }
    // two pipeline slots: while slot k's output drains to the host, slot k^1 is converted
    const size_t piece = (size_t)4 << 20;                 // 4 MiB of bits -> 64 MiB of IQ per slot
    unsigned char *d_in[2] = {nullptr, nullptr}, *h_in[2] = {nullptr, nullptr};
    int8_t *d_out[2] = {nullptr, nullptr}, *h_out[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    int rc = GPSACQ_OK;
    for (int k = 0; k < 2 && rc == GPSACQ_OK; k++) {
        if (cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&d_in[k], piece) != cudaSuccess ||
            cudaMalloc(&d_out[k], 16 * piece) != cudaSuccess || cudaMallocHost(&h_in[k], piece) != cudaSuccess ||
            cudaMallocHost(&h_out[k], 16 * piece) != cudaSuccess) { g_create_error = "bits_to_iq8: allocation failed"; rc = GPSACQ_ENOMEM; }
    }
    size_t pending_off[2] = {0, 0}, pending_n[2] = {0, 0};
    size_t done = 0;
    for (int it = 0; rc == GPSACQ_OK && (done < n_bytes || pending_n[0] || pending_n[1]); it++) {
        const int k = it & 1;
        if (pending_n[k]) {                                // drain what this slot produced two iterations ago
            if (cudaStreamSynchronize(st[k]) != cudaSuccess) { g_create_error = "bits_to_iq8: stream failed"; rc = GPSACQ_ECUDA; break; }
            memcpy(iq_out + 16 * pending_off[k], h_out[k], 16 * pending_n[k]);
            pending_n[k] = 0;
        }
        if (done < n_bytes) {
            const size_t n = std::min(piece, n_bytes - done);
            memcpy(h_in[k], bits + done, n);
            if (cudaMemcpyAsync(d_in[k], h_in[k], n, cudaMemcpyHostToDevice, st[k]) != cudaSuccess) { g_create_error = "bits_to_iq8: H2D failed"; rc = GPSACQ_ECUDA; break; }
            rc = gpsacq_bits_to_iq8_device(-1, d_in[k], n, first_sample + 8 * done, fc, fs, amplitude, d_out[k], st[k]);
            if (rc) break;
            if (cudaMemcpyAsync(h_out[k], d_out[k], 16 * n, cudaMemcpyDeviceToHost, st[k]) != cudaSuccess) { g_create_error = "bits_to_iq8: D2H failed"; rc = GPSACQ_ECUDA; break; }
            pending_off[k] = done; pending_n[k] = n;
            done += n;
        }
    }
    for (int k = 0; k < 2; k++) {
        if (st[k]) { cudaStreamSynchronize(st[k]); cudaStreamDestroy(st[k]); }
        cudaFree(d_in[k]); cudaFree(d_out[k]); cudaFreeHost(h_in[k]); cudaFreeHost(h_out[k]);
    }
    return rc;
}

// ---- gps_sig_gen.m, literally --------------------------------------------------------------------------------------
extern "C" int gpsacq_sig_gen_literal(int device, int prn, const uint8_t *nav_bits01, int n_nav_bits, uint8_t *bits_out, void *d_bits_out)
{
    DeviceGuard guard;
    if (prn < 1 || prn > 32 || !nav_bits01 || n_nav_bits < 1 || n_nav_bits > (1 << 20) || (!bits_out && !d_bits_out)) {
        g_create_error = "gpsacq_sig_gen_literal: bad arguments"; return GPSACQ_EINVAL;
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_create_error = "gpsacq_sig_gen_literal: cudaSetDevice failed (no CPU fallback)"; return GPSACQ_ECUDA; }
    const long long n_data = (long long)n_nav_bits * SIGLIT_PER_BIT, n_out = n_data + SIGLIT_TAPS - 1;
    const size_t nbytes = (size_t)((n_out + 7) / 8);
    const double ca_rate = 1.023e6 * SIGLIT_OV, fc = ca_rate / 4;                          // gps_sig_gen.m:8-9,14,34
    const double two_pi_fc = (2.0 * 3.141592653589793) * fc, inv_rate = 1.0 / ca_rate;     // 2.*pi.*fc ... .*(1./ca_rate), left to right
    unsigned char *d_nav = nullptr, *d = (unsigned char *)d_bits_out;
    bool own = false;
    cudaError_t e = cudaMalloc(&d_nav, (size_t)n_nav_bits);
    if (e == cudaSuccess) e = cudaMemcpy(d_nav, nav_bits01, (size_t)n_nav_bits, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !d) { e = cudaMalloc(&d, nbytes); own = e == cudaSuccess; }
    if (e == cudaSuccess) {
        sig_gen_literal_kernel<<<(unsigned)((nbytes + 255) / 256), 256>>>(kTaps[prn - 1][0], kTaps[prn - 1][1], d_nav, n_data, n_out, two_pi_fc, inv_rate, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && bits_out) e = cudaMemcpy(bits_out, d, nbytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_nav);
    if (own) cudaFree(d);
    if (e != cudaSuccess) { g_create_error = std::string("gpsacq_sig_gen_literal: ") + cudaGetErrorString(e); return GPSACQ_ECUDA; }
    return GPSACQ_OK;
}

extern "C" {

int gpsacq_synth_capture(int device, double fs, double fc, const gpsacq_sat *sats, int n_sats, double noise_sigma,
                         double nav_bps, uint64_t seed, size_t n_samples, uint8_t *bits_out, void *d_bits_out)
{
    DeviceGuard guard;
    if (!(fs > 0) || !std::isfinite(fs) || !std::isfinite(fc) || !std::isfinite(noise_sigma) || !std::isfinite(nav_bps) ||
        n_sats < 0 || n_sats > SYNTH_MAX_SATS || (n_sats && !sats) || (!bits_out && !d_bits_out)) {
        g_create_error = "gpsacq_synth_capture: bad arguments (at most 16 satellites)"; return GPSACQ_EINVAL;
    }
    SynthParams p;
    memset(&p, 0, sizeof p);
    p.n_sats = n_sats; p.fs = fs; p.fc = fc; p.sigma = noise_sigma; p.nav_bps = nav_bps > 0 ? nav_bps : 50.0; p.seed = seed;
    for (int i = 0; i < n_sats; i++) {
        if (sats[i].prn < 1 || sats[i].prn > 32) { g_create_error = "gpsacq_synth_capture: prn out of range 1..32"; return GPSACQ_EINVAL; }
        if (!std::isfinite(sats[i].amp) || !std::isfinite(sats[i].doppler_hz) || !std::isfinite(sats[i].carrier_phase_cycles) ||
            !(sats[i].code_phase_chips >= 0.0 && sats[i].code_phase_chips < 1023.0) || !(fabs(sats[i].doppler_hz) < 1e7)) {
            g_create_error = "gpsacq_synth_capture: amp / doppler_hz / carrier phase must be finite and code_phase_chips in [0, 1023)"; return GPSACQ_EINVAL;
        }
        p.sat[i].prn = sats[i].prn; p.sat[i].t0 = kTaps[sats[i].prn - 1][0]; p.sat[i].t1 = kTaps[sats[i].prn - 1][1];
        p.sat[i].amp = sats[i].amp; p.sat[i].doppler_hz = sats[i].doppler_hz;
        p.sat[i].code_phase_chips = sats[i].code_phase_chips; p.sat[i].carrier_phase_cycles = sats[i].carrier_phase_cycles;
    }
    if (n_samples == 0) return GPSACQ_OK;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { g_create_error = "gpsacq_synth_capture: cudaSetDevice failed (no CPU fallback)"; return GPSACQ_ECUDA; }
    const size_t nbytes = (n_samples + 7) / 8;
    unsigned char *d = (unsigned char *)d_bits_out;
    bool own = false;
    if (!d) { if (cudaMalloc(&d, nbytes) != cudaSuccess) { g_create_error = "gpsacq_synth_capture: cudaMalloc failed"; return GPSACQ_ECUDA; } own = true; }
    synth_bits_kernel<<<(unsigned)((nbytes + 255) / 256), 256>>>(p, n_samples, 0, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && bits_out) e = cudaMemcpy(bits_out, d, nbytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (own) cudaFree(d);
    if (e != cudaSuccess) { g_create_error = std::string("gpsacq_synth_capture: ") + cudaGetErrorString(e); return GPSACQ_ECUDA; }
    return GPSACQ_OK;
}

int gpsacq_stage_times(gpsacq_t *h, float ms[4])
{
    if (!h || !ms) return GPSACQ_EINVAL;
    if (!h->have_batch) { h->err = "no batch processed yet"; return GPSACQ_ESTATE; }
    CUDA_TRY(h, cudaEventSynchronize(h->ev[3]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[0], h->ev[0], h->ev[1]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[1], h->ev[1], h->ev[2]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[2], h->ev[2], h->ev[3]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[3], h->ev[0], h->ev[3]));
    return GPSACQ_OK;
}

int gpsacq_get_replica_time(gpsacq_t *h, int sv, float *out)
{
    DeviceGuard guard;
    if (!h || !out || sv < 0 || sv >= 32) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_REF) { h->err = "probe needs a GPSACQ_MODE_REF handle"; return GPSACQ_EINVAL; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(out, h->d_repl_time + (size_t)sv * h->n, (size_t)h->n * sizeof(float), cudaMemcpyDeviceToHost));
    return GPSACQ_OK;
}

static int read_natural(gpsacq *h, const cf *src, int mode, float *out)
{
    CUDA_TRY(h, cudaSetDevice(h->device));
    undecimate_kernel<<<(h->n + 255) / 256, 256, 0, h->stream>>>(src, h->n1, h->n2, mode, h->d_nat);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(out, h->d_nat, (size_t)h->n * sizeof(cf), cudaMemcpyDeviceToHost));
    return GPSACQ_OK;
}

int gpsacq_get_replica_spectrum(gpsacq_t *h, int sv, float *out)
{
    DeviceGuard guard;
    if (!h || !out || sv < 0 || sv >= 32) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_REF) { h->err = "probe needs a GPSACQ_MODE_REF handle"; return GPSACQ_EINVAL; }
    return read_natural(h, h->d_cext + (size_t)sv * 2 * h->n, 1, out);
}

int gpsacq_get_block_spectrum(gpsacq_t *h, size_t blk, float *out)
{
    DeviceGuard guard;
    if (!h || !out) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_REF) { h->err = "probe needs a GPSACQ_MODE_REF handle"; return GPSACQ_EINVAL; }
    if (!h->have_batch || blk >= h->last_blocks) { h->err = "block index outside the last batch"; return GPSACQ_ESTATE; }
    return read_natural(h, h->d_xd + blk * (size_t)h->n, 0, out);
}

int gpsacq_get_cell_stats(gpsacq_t *h, size_t blk, gpsacq_cell *out)
{
    DeviceGuard guard;
    if (!h || !out) return GPSACQ_EINVAL;
    if (!h->have_batch || blk >= h->last_blocks) { h->err = "block index outside the last batch"; return GPSACQ_ESTATE; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(out, h->d_cells + blk * (size_t)h->ndop, (size_t)h->ndop * sizeof(CellStat), cudaMemcpyDeviceToHost));
    return GPSACQ_OK;
}

}  // extern "C"

// =========================================================================================
// Several GPUs in one process: contiguous chunk ranges per device + one ncclAllGather of peak records
// =========================================================================================
#include <dlfcn.h>
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
struct NcclApi {
    void *lib;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*AllGather)(const void *, void *, size_t, int /*ncclDataType_t*/, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};
static bool nccl_load(NcclApi &a)
{
    memset(&a, 0, sizeof a);
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
    if (!a.lib) return false;
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
    a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.AllGather;
    return a.ok;
}

struct gpsacq_group {
    std::vector<gpsacq *> eng;
    std::vector<int> dev;
    std::vector<ncclComm_t> comm;
    std::vector<Peak *> d_all;        // per device: n_gpus * cap records
    std::vector<std::vector<int> > sv;
    NcclApi nccl;
    bool use_nccl;
    int cap;                          // per-device share of a batch
    std::vector<Peak> h_all;
    std::string err;
};

// ONE collective per batch: every device ends up with every device's (padded) records; device 0's copy goes to the
// host.  Every NCCL / CUDA return code is checked; the first failure is reported (the group is still closed).
static int group_allgather(gpsacq_group *g)
{
    const size_t ng = g->eng.size();
    ncclResult_t first = 0;
    const char *what = nullptr;
    ncclResult_t r = g->nccl.GroupStart();
    if (r != 0) { first = r; what = "ncclGroupStart"; }
    for (size_t d = 0; d < ng && first == 0; d++) {
        if (cudaSetDevice(g->dev[d]) != cudaSuccess) { g->err = "cudaSetDevice failed before ncclAllGather"; first = -1; what = "cudaSetDevice"; break; }
        r = g->nccl.AllGather(g->eng[d]->d_peaks, g->d_all[d], (size_t)g->cap * sizeof(Peak), 0 /*ncclInt8*/, g->comm[d], g->eng[d]->stream);
        if (r != 0) { first = r; what = "ncclAllGather"; }
    }
    r = g->nccl.GroupEnd();
    if (r != 0 && first == 0) { first = r; what = "ncclGroupEnd"; }
    if (first != 0) {
        if (first > 0) g->err = std::string(what) + ": " + (g->nccl.GetErrorString ? g->nccl.GetErrorString(first) : "error");
        return GPSACQ_ECUDA;
    }
    if (cudaSetDevice(g->dev[0]) != cudaSuccess) { g->err = "cudaSetDevice failed"; return GPSACQ_ECUDA; }
    if (cudaMemcpyAsync(g->h_all.data(), g->d_all[0], ng * (size_t)g->cap * sizeof(Peak), cudaMemcpyDeviceToHost, g->eng[0]->stream) != cudaSuccess) { g->err = "D2H failed"; return GPSACQ_ECUDA; }
    for (size_t d = 0; d < ng; d++) {
        if (cudaSetDevice(g->dev[d]) != cudaSuccess || cudaStreamSynchronize(g->eng[d]->stream) != cudaSuccess) { g->err = "stream sync failed after ncclAllGather"; return GPSACQ_ECUDA; }
    }
    return GPSACQ_OK;
}

extern "C" {

const char *gpsacq_group_last_error(const gpsacq_group_t *g) { return g ? g->err.c_str() : g_create_error.c_str(); }
const char *gpsacq_group_gather_kind(const gpsacq_group_t *g) { return (g && g->use_nccl) ? "nccl" : "host"; }
gpsacq_t *gpsacq_group_engine(gpsacq_group_t *g, int i) { return (g && i >= 0 && i < (int)g->eng.size()) ? g->eng[i] : nullptr; }

void gpsacq_group_destroy(gpsacq_group_t *g)
{
    DeviceGuard guard;
    if (!g) return;
    for (size_t i = 0; i < g->eng.size(); i++) {
        if (cudaSetDevice(g->dev[i]) != cudaSuccess) fprintf(stderr, "libgpsacq: cudaSetDevice(%d) failed while destroying a group\n", g->dev[i]);
        if (i < g->d_all.size()) cudaFree(g->d_all[i]);
        if (g->use_nccl && i < g->comm.size() && g->comm[i]) g->nccl.CommDestroy(g->comm[i]);
        gpsacq_destroy(g->eng[i]);
    }
    delete g;
}

int gpsacq_group_create(const gpsacq_cfg *cfg, int n_gpus, const int32_t *devices, int use_nccl, gpsacq_group_t **out)
{
    DeviceGuard guard;
    if (!cfg || !out || n_gpus < 1 || (cfg->mode != GPSACQ_MODE_REF && cfg->mode != GPSACQ_MODE_GRID)) { g_create_error = "gpsacq_group_create: bad arguments"; return GPSACQ_EINVAL; }
    int ndop_full = 0;
    if (cfg->mode == GPSACQ_MODE_GRID) {
        if (!(cfg->doppler_step > 0) || cfg->dop_count != 0) { g_create_error = "gpsacq_group_create: GRID needs doppler_step > 0 and an unsharded cfg"; return GPSACQ_EINVAL; }
        ndop_full = 2 * (int)floor(cfg->max_fo / cfg->doppler_step + 1e-9) + 1;
        if (n_gpus > ndop_full) { g_create_error = "gpsacq_group_create: more GPUs than Doppler bins"; return GPSACQ_EINVAL; }
    }
    *out = nullptr;
    gpsacq_group *g = new (std::nothrow) gpsacq_group();
    if (!g) return GPSACQ_ENOMEM;
    g->use_nccl = false;
    for (int i = 0; i < n_gpus; i++) {
        gpsacq_cfg c = *cfg;
        c.device = devices ? devices[i] : i;
        if (cfg->mode == GPSACQ_MODE_GRID) {      // contiguous, balanced bin ranges in ascending order
            c.dop_first = (int)((long long)ndop_full * i / n_gpus);
            c.dop_count = (int)((long long)ndop_full * (i + 1) / n_gpus) - c.dop_first;
        }
        gpsacq *h = nullptr;
        const int rc = gpsacq_create(&c, &h);
        if (rc) { gpsacq_group_destroy(g); return rc; }
        g->eng.push_back(h);
        g->dev.push_back(h->device);
    }
    g->cap = g->eng[0]->cap;
    if (cfg->mode == GPSACQ_MODE_GRID) {          // records per device and batch: 32 per acquisition
        int cap_acq = g->eng[0]->cap_acq;
        for (gpsacq *h : g->eng) cap_acq = std::min(cap_acq, h->cap_acq);
        g->cap = cap_acq * 32;
    }
    g->sv.resize(n_gpus);
    g->d_all.assign(n_gpus, nullptr);
    for (int i = 0; i < n_gpus; i++) {
        if (cudaSetDevice(g->dev[i]) != cudaSuccess) { g_create_error = "group: cudaSetDevice failed"; gpsacq_group_destroy(g); return GPSACQ_ECUDA; }
        if (cudaMalloc(&g->d_all[i], (size_t)n_gpus * g->cap * sizeof(Peak)) != cudaSuccess) { g_create_error = "group: cudaMalloc failed"; gpsacq_group_destroy(g); return GPSACQ_ENOMEM; }
    }
    g->h_all.resize((size_t)n_gpus * g->cap);
    if (use_nccl && n_gpus > 1) {
        // NCCL was asked for: a fallback to the host gather is never silent (stderr + gpsacq_group_gather_kind()),
        // and GPSACQ_REQUIRE_NCCL=1 turns it into an error
        std::string why;
        if (!nccl_load(g->nccl)) why = "libnccl.so.2 could not be loaded";
        else {
            g->comm.assign(n_gpus, nullptr);
            const ncclResult_t r = g->nccl.CommInitAll(g->comm.data(), n_gpus, g->dev.data());
            if (r == 0) g->use_nccl = true;
            else { why = std::string("ncclCommInitAll: ") + (g->nccl.GetErrorString ? g->nccl.GetErrorString(r) : "error"); g->comm.clear(); }
        }
        if (!g->use_nccl) {
            const char *req = getenv("GPSACQ_REQUIRE_NCCL");
            if (req && *req && *req != '0') { g_create_error = "group: NCCL required but unavailable: " + why; gpsacq_group_destroy(g); return GPSACQ_ECUDA; }
            fprintf(stderr, "libgpsacq: NCCL peak gather unavailable (%s); gathering the peak records through host memory\n", why.c_str());
        }
    }
    *out = g;
    return GPSACQ_OK;
}

int gpsacq_group_search_blocks(gpsacq_group_t *g, const uint8_t *bits, size_t n_blocks, gpsacq_peak *out)
{
    DeviceGuard guard;
    if (!g || (!bits && n_blocks) || (!out && n_blocks)) return GPSACQ_EINVAL;
    const size_t ng = g->eng.size(), per_batch = (size_t)g->cap * ng;
    bool pinned = false;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, bits) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    for (size_t done = 0; done < n_blocks;) {
        const size_t nb = std::min(per_batch, n_blocks - done);
        // contiguous, balanced ranges of this batch; chunk b of the stream keeps PRN b mod 32
        std::vector<size_t> lo(ng + 1);
        for (size_t d = 0; d <= ng; d++) lo[d] = nb * d / ng;
        for (size_t d = 0; d < ng; d++) {
            gpsacq *h = g->eng[d];
            const size_t n = lo[d + 1] - lo[d];
            if (cudaSetDevice(h->device) != cudaSuccess) { g->err = "cudaSetDevice failed"; return GPSACQ_ECUDA; }
            if (n == 0) continue;
            // a page-locked caller buffer goes straight to the copy engine (every device starts at once); pageable memory
            // is staged through the engine's pinned buffer, device after device
            const unsigned char *src = bits + (done + lo[d]) * (size_t)h->chunk_bytes;
            if (!pinned) { memcpy(h->h_bits, src, n * (size_t)h->chunk_bytes); src = h->h_bits; }
            for (size_t b = 0; b < n; b++) h->h_sv[b] = (int)((done + lo[d] + b) % GPSACQ_NUM_SATS);
            if (cudaMemcpyAsync(h->d_bits, src, n * (size_t)h->chunk_bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                cudaMemcpyAsync(h->d_sv, h->h_sv, n * sizeof(int), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { g->err = "H2D copy failed"; return GPSACQ_ECUDA; }
            const int rc = gpsacq_search_blocks_device(h, h->d_bits, n, h->d_sv, (gpsacq_peak *)h->d_peaks);
            if (rc) { g->err = h->err; return rc; }
        }
        if (g->use_nccl) {
            const int rc = group_allgather(g);
            if (rc) return rc;
            for (size_t d = 0; d < ng; d++)
                memcpy(out + done + lo[d], g->h_all.data() + d * (size_t)g->cap, (lo[d + 1] - lo[d]) * sizeof(Peak));
        } else {
            for (size_t d = 0; d < ng; d++) {
                gpsacq *h = g->eng[d];
                const size_t n = lo[d + 1] - lo[d];
                if (cudaSetDevice(h->device) != cudaSuccess) { g->err = "cudaSetDevice failed"; return GPSACQ_ECUDA; }
                if (n && cudaMemcpyAsync(h->h_peaks, h->d_peaks, n * sizeof(Peak), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { g->err = "D2H failed"; return GPSACQ_ECUDA; }
            }
            for (size_t d = 0; d < ng; d++) {
                gpsacq *h = g->eng[d];
                if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) { g->err = "stream sync failed"; return GPSACQ_ECUDA; }
                memcpy(out + done + lo[d], h->h_peaks, (lo[d + 1] - lo[d]) * sizeof(Peak));
            }
        }
        done += nb;
    }
    return GPSACQ_OK;
}

int gpsacq_group_acquire(gpsacq_group_t *g, const uint8_t *bits, size_t n_acq, gpsacq_peak *out)
{
    DeviceGuard guard;
    if (!g || (!bits && n_acq) || (!out && n_acq)) return GPSACQ_EINVAL;
    if (g->eng[0]->mode != GPSACQ_MODE_GRID) { g->err = "gpsacq_group_acquire needs a GPSACQ_MODE_GRID group"; return GPSACQ_EINVAL; }
    const size_t ng = g->eng.size(), cap_acq = (size_t)g->cap / 32, acq_bytes = (size_t)g->eng[0]->chunk_bytes;
    for (size_t done = 0; done < n_acq;) {
        const size_t na = std::min(cap_acq, n_acq - done), nrec = na * 32;
        // every device gets the whole (tiny) input and searches its own Doppler bins
        for (size_t d = 0; d < ng; d++) {
            gpsacq *h = g->eng[d];
            if (cudaSetDevice(h->device) != cudaSuccess) { g->err = "cudaSetDevice failed"; return GPSACQ_ECUDA; }
            memcpy(h->h_bits, bits + done * acq_bytes, na * acq_bytes);
            if (cudaMemcpyAsync(h->d_bits, h->h_bits, na * acq_bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { g->err = "H2D copy failed"; return GPSACQ_ECUDA; }
            const int rc = acquire_device_impl(h, h->d_bits, na, (gpsacq_peak *)h->d_peaks);
            if (rc) { g->err = h->err; return rc; }
        }
        if (g->use_nccl) {
            const int rc = group_allgather(g);
            if (rc) return rc;
        } else {
            for (size_t d = 0; d < ng; d++) {
                gpsacq *h = g->eng[d];
                if (cudaSetDevice(h->device) != cudaSuccess) { g->err = "cudaSetDevice failed"; return GPSACQ_ECUDA; }
                if (cudaMemcpyAsync(h->h_peaks, h->d_peaks, nrec * sizeof(Peak), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { g->err = "D2H failed"; return GPSACQ_ECUDA; }
            }
            for (size_t d = 0; d < ng; d++) {
                gpsacq *h = g->eng[d];
                if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) { g->err = "stream sync failed"; return GPSACQ_ECUDA; }
                memcpy(g->h_all.data() + d * (size_t)g->cap, h->h_peaks, nrec * sizeof(Peak));
            }
        }
        // merge: devices hold ascending bin ranges, so "strictly greater" keeps the lower bin on equal snr
        for (size_t r = 0; r < nrec; r++) {
            Peak best = g->h_all[r];
            for (size_t d = 1; d < ng; d++) {
                const Peak &p = g->h_all[d * (size_t)g->cap + r];
                if (p.snr > best.snr) best = p;
            }
            memcpy(out + done * 32 + r, &best, sizeof(Peak));
        }
        done += na;
    }
    return GPSACQ_OK;
}

}  // extern "C" (group)

#include "ga_service.h"
