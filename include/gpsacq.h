/*
 * include/gpsacq.h -- C ABI of libgpsacq.so, the B200 (sm_100a) GPS L1 C/A
 * acquisition engine.
 *
 * This is the drop-in boundary for the reference's offline search path
 * (JiaoXianjun/GNSS-GPS-SDR, c/search_offline.cpp).  The reference exposes that
 * path as five C++-linkage functions plus three caller-defined globals
 * (c/gps_offline.h:23-25, :87-91):
 *
 *     extern double FC, FS, max_fo;
 *     int  SearchInit();  void SearchFree();  void SearchTask(char *file);
 *     void SearchEnable(int sv);  int SearchCode(int sv, int g1);
 *
 * and keeps all state in file statics (code[32][40000], fwd_buf, rev_buf, two
 * FFTW plans; c/search_offline.cpp:55-64).  The host C++ in
 * gnss-gps-sdr_b200/c/ re-implements those five symbols on top of the entry
 * points below; nothing here uses C++ or torch types, so cgo / JNI / ctypes
 * bindings work the same way (INTEGRATION.md shows them).
 *
 * What each entry point replaces:
 *   gpsacq_create            SearchInit()  c/search_offline.cpp:74-110  (replica generation + 32 forward
 *                                          FFTs; also reads FC/FS/max_fo like Sample()/Correlate() do,
 *                                          :76,:127,:176,:190) -- but on the device, into HBM.
 *   gpsacq_destroy           SearchFree()  :114-117
 *   gpsacq_search_blocks     the body of SearchTask()'s satellite loop, :239-257:
 *                            Sample() :121-165 (unpack, XOR mix, forward FFT) and
 *                            Correlate() :169-201 (Doppler loop, shifted conj-multiply,
 *                            backward FFT, |.|^2, peak/mean, best over Doppler) for a batch
 *                            of 5120-byte chunks.  REF semantics: chunk b is searched for
 *                            PRN (b mod 32)+1, or sv_of_block[b]+1 when given.
 *   gpsacq_search_blocks_device   same with device-resident input/output (no copies).
 *   gpsacq_get_*             test probes: read back what the reference holds in code[sv],
 *                            fwd_buf and the per-Doppler (max_pwr, max_pwr_i, tot_pwr) locals.
 *
 * Error convention: every int function returns 0 on success, a negative
 * GPSACQ_E* code otherwise; gpsacq_last_error() gives the text.  There is NO
 * CPU fallback: without a usable CUDA device gpsacq_create() fails with
 * GPSACQ_ECUDA.
 *
 * Threading: one host thread per handle at a time (the reference is not
 * re-entrant at all).  Several handles (e.g. one per GPU) may be used
 * concurrently from different threads.
 *
 * Current CUDA device: every entry point that has to select a device (create, search, acquire,
 * probes, the group calls) restores the calling thread's current device before it returns.
 */
#ifndef GPSACQ_H
#define GPSACQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSACQ_ABI_VERSION 4

#define GPSACQ_OK        0
#define GPSACQ_EINVAL   (-1)   /* bad argument / unsupported configuration        */
#define GPSACQ_ECUDA    (-2)   /* CUDA runtime error or no device                 */
#define GPSACQ_ENOMEM   (-3)   /* host or device allocation failed                */
#define GPSACQ_ESTATE   (-4)   /* probe called before any batch was processed     */

#define GPSACQ_NUM_SATS 32     /* NUM_SATS, c/gps_offline.h:16                     */
#define GPSACQ_FFT_LEN  40000  /* FFT_LEN,  c/gps_offline.h:15                     */

typedef struct gpsacq gpsacq_t;

typedef struct gpsacq_cfg {
    double  fc;          /* carrier at IF, Hz        (extern double FC,     c/gps_offline.h:23) */
    double  fs;          /* sampling rate, Hz        (extern double FS,     :24)               */
    double  max_fo;      /* Doppler search half-span (extern double max_fo, :25)               */
    int32_t fft_len;     /* 0 = GPSACQ_FFT_LEN; only 40000 is supported                        */
    int32_t device;      /* CUDA device ordinal; -1 = current device                           */
    int32_t max_blocks;  /* REF: batch capacity in chunks, 0 = 512 (16 runs of 32 PRNs);
                            GRID: batch capacity in acquisitions, 0 = automatic                    */
    int32_t mode;        /* GPSACQ_MODE_REF (0, the reference's semantics) or GPSACQ_MODE_GRID      */
    double  doppler_step;/* GRID only: Doppler bin spacing in Hz; FS/doppler_step must be an integer */
    int32_t noncoh_blocks;/* GRID only: K = number of 1 ms blocks summed non-coherently (>= 1)       */
    int32_t dop_first;   /* GRID only: search just the bins [dop_first, dop_first + dop_count) of the     */
    int32_t dop_count;   /*   n_doppler-bin grid (index 0 = bin -dmax).  dop_count = 0: the whole grid.    */
    int32_t reserved;    /*   This is how the (PRN x Doppler) grid of ONE acquisition is sharded over GPUs: */
                         /*   records keep absolute bin numbers, so shards merge by max snr / lower bin.    */
    double  fs_replica;  /* REF only, 0 = fs: sampling rate the C/A replicas are generated for.  The reference reads
                            FS once in SearchInit() for the code NCO (c/search_offline.cpp:76) but FC, FS and max_fo
                            again on every Sample()/Correlate() (:127,:176,:190); a caller that changes the globals
                            between SearchInit() and SearchTask() gets replicas at the old FS searched on the new grid.
                            The C++ host passes the SearchInit()-time FS here when it rebuilds the engine for changed
                            globals.                                                                              */
} gpsacq_cfg;

#define GPSACQ_MODE_REF  0   /* N = 40000 coherent window, bins of FS/N, one chunk per PRN
                                (c/search_offline.cpp as it is)                                     */
#define GPSACQ_MODE_GRID 1   /* generalised grid of BASELINE.json configs[1..4] (NOT in the reference;
                                definition in SURVEY.md App. E / DESIGN.md section 10): 1 ms coherent
                                blocks of W = FS/1000 samples shared by all 32 PRNs, Doppler bins
                                d*doppler_step for |d*doppler_step| <= max_fo with time-domain
                                wipe-off, W-point circular correlation, K-block non-coherent sum    */

/* One record per searched chunk = what Correlate() returns plus its inputs to the
 * snr division (c/search_offline.cpp:196-200). */
typedef struct gpsacq_peak {
    float   snr;         /* max over Doppler of max_pwr/(tot_pwr/W); 0 if none > 0 (:173,:198)  */
    float   max_pwr;     /* of the winning Doppler bin                                         */
    float   tot_pwr;     /* of the winning Doppler bin                                         */
    int32_t lo_shift;    /* winning Doppler bin, -dmax..+dmax  (*max_snr_dop, :198)            */
    int32_t ca_shift;    /* code phase in samples, 0..W-1      (*max_snr_i,   :198)            */
    int32_t sv;          /* 0-based satellite index (PRN-1) this chunk was searched for        */
    int32_t flags;       /* bit0: snr >= 25 (the SearchTask() detection rule, :248);
                            bit31: the device-side PRN map entry was outside 0..31 (searched as sv & 31) */
    int32_t reserved;
} gpsacq_peak;

/* Per-(chunk, Doppler bin) statistics: Correlate()'s max_pwr, max_pwr_i, tot_pwr (:177-194). */
typedef struct gpsacq_cell {
    float   max_pwr;
    float   tot_pwr;
    int32_t max_idx;
    int32_t reserved;
} gpsacq_cell;

typedef struct gpsacq_info {
    int32_t abi_version;
    int32_t fft_len;       /* N                                                   */
    int32_t n1, n2;        /* N = n1*n2: n1 decimated sub-sequences of n2 points   */
    int32_t window;        /* W = ceil(FS/1000) code phases searched (:190)        */
    int32_t dmax;          /* Doppler bins -dmax..+dmax (:176)                     */
    int32_t n_doppler;     /* 2*dmax+1                                             */
    int32_t chunk_bytes;   /* bytes consumed per chunk (5120 for N=40000, :129-141)*/
    int32_t max_blocks;    /* batch capacity                                       */
    int32_t device;        /* CUDA ordinal in use                                  */
    int32_t sm_count;
    int32_t cell_ctas;     /* persistent CTAs of the cell kernel                   */
    int32_t cell_threads;
    int32_t cell_smem_bytes;
    int64_t bytes_per_corr;/* algorithmic bytes per (PRN,Doppler) correlation: 2*N*8+16
                              (REF: N = fft_len; GRID: N = window, per coherent block)   */
    int32_t mode;          /* GPSACQ_MODE_*                                        */
    int32_t noncoh_blocks; /* K (1 in REF mode)                                    */
    int32_t block_bytes;   /* GRID: bytes per 1 ms block = window/8                */
    int32_t max_acq;       /* GRID: acquisitions per internal batch                */
    double  doppler_step;  /* Hz between Doppler bins (REF: FS/fft_len)            */
    int32_t dop_first;     /* GRID: first bin of this handle's shard (0 = bin -dmax); n_doppler = bins in the shard */
    int32_t n_doppler_full;/* bins of the whole grid, 2*dmax+1                                    */
    int32_t blocks_per_launch; /* REF: chunks per kernel launch of the most recent batch (a batch is cut into launches
                              of GPSACQ_SUB_BLOCKS = 128 chunks so that the block spectra stay in L2); 0 in GRID mode */
    int32_t reserved;
} gpsacq_info;

/* Kernels launched per launch triple (forward transform, cells, best over Doppler); a REF batch of n chunks is
 * ceil(n / blocks_per_launch) triples (for bench.py's gpu_launches). */
#define GPSACQ_LAUNCHES_PER_BATCH 3

int  gpsacq_create(const gpsacq_cfg *cfg, gpsacq_t **out);
void gpsacq_destroy(gpsacq_t *h);
const char *gpsacq_last_error(const gpsacq_t *h);   /* h may be NULL: error of the last failed create */
int  gpsacq_get_info(const gpsacq_t *h, gpsacq_info *info);

/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work of this
 * handle; NULL restores the handle's own stream. */
int  gpsacq_set_stream(gpsacq_t *h, void *cuda_stream);
int  gpsacq_synchronize(gpsacq_t *h);

/* Host-buffer search: packed_bits = n_blocks * chunk_bytes bytes of 1-bit samples,
 * LSB first (c/search_offline.cpp:143-146).  sv_of_block may be NULL (REF rule:
 * b mod 32).  out receives n_blocks records.  Any n_blocks; batches internally.
 * Includes the host->device and device->host copies and a stream synchronise. */
int  gpsacq_search_blocks(gpsacq_t *h, const uint8_t *packed_bits, size_t n_blocks,
                          const int32_t *sv_of_block, gpsacq_peak *out);

/* Device-buffer search, asynchronous on the handle's stream.  n_blocks <= max_blocks.
 * d_sv_of_block may be NULL. */
int  gpsacq_search_blocks_device(gpsacq_t *h, const uint8_t *d_packed_bits, size_t n_blocks,
                                 const int32_t *d_sv_of_block, gpsacq_peak *d_out);

/* GRID mode: n_acq acquisitions, each over noncoh_blocks consecutive 1 ms blocks of packed_bits
 * (n_acq * noncoh_blocks * block_bytes bytes, LSB first).  All 32 PRNs are searched on the same
 * blocks; out receives 32 records per acquisition (PRN order), lo_shift = Doppler bin index d
 * (Doppler = d * doppler_step Hz), ca_shift = code phase in samples.  Host buffers / device buffers. */
int  gpsacq_acquire(gpsacq_t *h, const uint8_t *packed_bits, size_t n_acq, gpsacq_peak *out);
int  gpsacq_acquire_device(gpsacq_t *h, const uint8_t *d_packed_bits, size_t n_acq, gpsacq_peak *d_out);

/* 8-bit IQ front-end (replaces the MATLAB step of the reference's SDR workflows,
 * proc_rtl_bin_for_gps.m:31-47 / proc_hackrf_bin_for_gps.m:7-19): interleaved I,Q bytes ->
 * remove the capture's mean -> shift up by shift_hz (the IF that gps_test is then told, e.g. 0.62e6)
 * -> take the real part -> 1 bit per sample, packed LSB first, ready for gpsacq_search_blocks() /
 * gpsacq_acquire().  format: GPSACQ_IQ_U8 (rtl-sdr, offset 128) or GPSACQ_IQ_S8 (HackRF).  n_samples
 * complex samples in, ceil(n_samples/8) bytes out.  Host buffers; runs on the handle's device. */
#define GPSACQ_IQ_U8 0
#define GPSACQ_IQ_S8 1
int  gpsacq_iq8_to_bits(gpsacq_t *h, const void *iq, size_t n_samples, int format, double shift_hz,
                        double fs, uint8_t *bits_out);

/* The same front-end with device-resident buffers, asynchronous on the handle's stream (d_iq: 2*n_samples bytes,
 * d_bits_out: ceil(n_samples/8) bytes, d_sums: 16 bytes of scratch).  For captures that fit in device memory.
 * When shift_hz/fs is a fraction p/q with q <= 227 (0.62/2.8, 2.6/10, n/4 ...) and d_iq is 16-byte aligned, the three
 * kernels (mean, threshold table, conversion) run back to back without a host round trip; other ratios read the mean
 * back (one stream synchronisation inside the call).  The bits are the same either way. */
int  gpsacq_iq8_to_bits_device(gpsacq_t *h, const void *d_iq, size_t n_samples, int format, double shift_hz,
                               double fs, uint8_t *d_bits_out, void *d_sums);

/* The reverse converter (c/conv_1bit_bin_to_hackrf_bin.cpp:29-86): packed 1-bit real-IF samples -> interleaved int8
 * I,Q at baseband for HackRF replay.  I = A*Bipolar(bit ^ lo_sin[int(phase)]), Q = A*Bipolar(bit ^ lo_cos[int(phase)]),
 * lo_sin = {1,1,0,0}, lo_cos = {1,0,0,1}, Bipolar(1) = -A (the reference's A is 30); the float phase NCO advances by
 * (float)(4*fc/fs) per sample, wraps at 4 and runs on over the WHOLE input (:33,:79-80).  n_bytes input bytes ->
 * 16*n_bytes output bytes.  first_sample: index of the input's first sample in the stream (0 for a whole file), so a
 * long file can be converted in pieces.  Host buffers (pipelined H2D / kernel / D2H); `_device`: device buffers,
 * asynchronous on `cuda_stream` (NULL: the default stream). */
int  gpsacq_bits_to_iq8(int device, const uint8_t *packed_bits, size_t n_bytes, size_t first_sample, double fc, double fs,
                        int amplitude, int8_t *iq_out);
int  gpsacq_bits_to_iq8_device(int device, const uint8_t *d_packed_bits, size_t n_bytes, size_t first_sample, double fc,
                               double fs, int amplitude, int8_t *d_iq_out, void *cuda_stream);

/* ---- several GPUs in one process ------------------------------------------------------------------
 * One engine per device.  REF mode: a batch's chunks are split into contiguous ranges (chunk b keeps PRN
 * b mod 32) and every device searches its range.  GRID mode: the Doppler bins of the grid are split into
 * contiguous ranges (cfg.dop_first/dop_count per device), every device searches all 32 PRNs of every
 * acquisition over its bins, and the per-device winners are merged (higher snr; equal snr -> lower bin,
 * the reference's ascending strictly-greater scan, c/search_offline.cpp:198).  Either way the 32-byte peak
 * records are exchanged with ONE collective per batch: ncclAllGather over NVLink (NCCL is dlopen'ed --
 * libnccl.so.2 -- so the library has no link-time dependency on it; if it cannot be loaded, or
 * use_nccl = 0, the records are gathered through the host).  No collective touches the data path.
 * gpsacq_group_gather_kind() says which gather is active. */
typedef struct gpsacq_group gpsacq_group_t;
int  gpsacq_group_create(const gpsacq_cfg *cfg, int n_gpus, const int32_t *devices /* NULL: 0..n-1 */,
                         int use_nccl, gpsacq_group_t **out);
void gpsacq_group_destroy(gpsacq_group_t *g);
int  gpsacq_group_search_blocks(gpsacq_group_t *g, const uint8_t *packed_bits, size_t n_blocks, gpsacq_peak *out);
/* GRID group: like gpsacq_acquire(); out receives 32 records per acquisition, identical to a single-GPU handle's. */
int  gpsacq_group_acquire(gpsacq_group_t *g, const uint8_t *packed_bits, size_t n_acq, gpsacq_peak *out);
const char *gpsacq_group_gather_kind(const gpsacq_group_t *g);   /* "nccl" or "host" */
const char *gpsacq_group_last_error(const gpsacq_group_t *g);
gpsacq_t *gpsacq_group_engine(gpsacq_group_t *g, int i);         /* engine of the i-th device (for info/probes) */

/* ---- acquisition -> tracking hand-off --------------------------------------
 * What CHANNEL::Start() (c/channel.cpp:134-171) derives from an acquisition record before it programs a tracking
 * channel: Doppler from the bin shift, 32-bit carrier / code NCO rate words, the code phase corrected for the code
 * creep since the sample was taken, the code-generator pause, and the Gold-code tap word (T1<<4)+T2 that
 * SearchTask() passes to ChanStart() (c/search.cpp:236-237).  Doppler = lo_shift*bin_num/bin_den, evaluated in that
 * order like :147: (bin_num, bin_den) = (FS, FFT_LEN) in REF mode, (doppler_step, 1) in GRID mode. */
typedef struct gpsacq_handoff {
    double   lo_dop_hz;    /* lo_shift*FS/FFT_LEN                        (:147) */
    double   ca_dop_hz;    /* lo_dop/L1*CPS                              (:148) */
    uint32_t lo_rate;      /* (FC+lo_dop)/FS*2^32                        (:151) */
    uint32_t ca_rate;      /* (CPS+ca_dop)/FS*2^32                       (:152) */
    int32_t  ca_shift;     /* + nearbyint(ca_dop*secs*FS/CPS)            (:161) */
    uint32_t ca_pause;     /* (2W - ca_shift) % W, W = samples per ms    (:164) */
    int32_t  taps;         /* (T1<<4)+T2                                        */
    int32_t  sv;
} gpsacq_handoff;
int  gpsacq_handoff_compute(const gpsacq_peak *p, double fc, double fs, double bin_num, double bin_den,
                            double secs_since_sample, gpsacq_handoff *out);

/* ---- streaming / re-acquisition service loop -------------------------------
 * The receiver's SearchTask() (c/search.cpp:214-239) over a stream of chunks: round-robin over the SVs that are
 * not being tracked (Busy[], SearchEnable()), ONE fresh chunk per searched SV, nothing sampled while all
 * num_chans channels are busy (ChanReset()), detection (snr >= 25) marks the SV busy and starts a channel.
 * Chunks are searched in GPU batches, speculatively (see csrc/ga_service.h); the event sequence is exactly the
 * one-chunk-at-a-time loop's.  REF-mode handle; the handle must outlive the service. */
typedef struct gpsacq_service gpsacq_service_t;
typedef struct gpsacq_event {
    int64_t        chunk_index;  /* which Sample() of the stream (0-based, counted over all feed() calls) */
    int32_t        sv;           /* 0-based SV that was detected                                          */
    int32_t        ch;           /* channel it was handed to (lowest free one, ChanReset())                 */
    gpsacq_peak    peak;         /* the acquisition record                                                 */
    gpsacq_handoff start;        /* CHANNEL::Start() values, secs_since_sample = one chunk                 */
} gpsacq_event;
int  gpsacq_service_create(gpsacq_t *h, int num_chans /* 0 = NUM_CHANS = 12 */, int max_rounds_per_batch /* 0 = handle capacity */,
                           gpsacq_service_t **out);
void gpsacq_service_destroy(gpsacq_service_t *s);
/* Consume chunks from `chunks` (n_chunks * chunk_bytes bytes) until they run out, every SV is busy, every channel
 * is busy or max_events events were produced.  *consumed = chunks taken (feed the rest again later). */
int  gpsacq_service_feed(gpsacq_service_t *s, const uint8_t *chunks, size_t n_chunks, size_t *consumed,
                         gpsacq_event *events, size_t max_events, size_t *n_events);
int  gpsacq_service_enable(gpsacq_service_t *s, int sv);        /* SearchEnable(sv), c/search.cpp:207-209          */
int  gpsacq_service_signal_lost(gpsacq_service_t *s, int ch);   /* CHANNEL::SignalLost() frees the channel (c/channel.cpp:245-249);
                                                                   call gpsacq_service_enable(sv) too, as :252 does */
int  gpsacq_service_state(const gpsacq_service_t *s, uint32_t *busy_svs, uint32_t *busy_chans, int64_t *chunks_seen);
const char *gpsacq_service_last_error(const gpsacq_service_t *s);

/* Synthetic 1-bit IF capture on the GPU (what gps_sig_gen.m + cacode.m produce, generalised: several
 * satellites, Doppler, code phase, noise).  n_samples real IF samples at fs around IF fc -> packed bits,
 * LSB first, into bits_out (host, ceil(n/8) bytes) and/or d_bits_out (device); either may be NULL.
 * Counter-based randomness: the same (seed, sample index) always gives the same noise / NAV bit. */
typedef struct gpsacq_sat {
    int32_t prn;                 /* 1..32 */
    int32_t reserved;
    double  amp;                 /* carrier amplitude relative to noise_sigma = 1 */
    double  doppler_hz;
    double  code_phase_chips;    /* 0..1023 at sample 0 */
    double  carrier_phase_cycles;
} gpsacq_sat;
int  gpsacq_synth_capture(int device, double fs, double fc, const gpsacq_sat *sats, int n_sats, double noise_sigma,
                          double nav_bps, uint64_t seed, size_t n_samples, uint8_t *bits_out, void *d_bits_out);

/* gps_sig_gen.m, literally (gps_sig_gen.m:8-41 with cacode.m): one satellite, chips zero-stuffed x8 to 8.184 Msps, 20
 * code periods per NAV bit, 49-tap rcosine(1,8) FIR, carrier at fs/4, sign, 'ubit1' -- the chain that wrote the
 * reference's bundled gps_sig_tmp.bin (PRN 8).  nav_bits01: n_nav_bits values 0/1 (the script draws them with rand;
 * data = 1 - 2*bit).  Writes ceil((n_nav_bits*163680 + 48)/8) bytes to bits_out (host) and/or d_bits_out (device).
 * Double precision with MATLAB's operation order: with the file's own NAV bits the output equals the file bit for bit. */
int  gpsacq_sig_gen_literal(int device, int prn, const uint8_t *nav_bits01, int n_nav_bits, uint8_t *bits_out, void *d_bits_out);

/* Elapsed milliseconds of the stages of the most recent batch, measured with CUDA
 * events on the launching stream: [0] unpack+mix+forward FFT kernel, [1] cell kernel
 * (conj-multiply + backward FFT + peak), [2] best-over-Doppler kernel, [3] whole batch.
 * (gpsacq_search_blocks() cuts large host batches into slices to overlap transfers with compute: the times
 * are then those of the LAST slice.)  Synchronises the stream. */
int  gpsacq_stage_times(gpsacq_t *h, float ms[4]);

/* ---- parity probes (copy device state to host; synchronise) -------------------- */
int  gpsacq_get_replica_time(gpsacq_t *h, int sv, float *out /* fft_len floats */);
int  gpsacq_get_replica_spectrum(gpsacq_t *h, int sv, float *out /* 2*fft_len: re,im of C[k] */);
int  gpsacq_get_block_spectrum(gpsacq_t *h, size_t block_in_last_batch, float *out /* 2*fft_len: X[k] */);
int  gpsacq_get_cell_stats(gpsacq_t *h, size_t block_in_last_batch, gpsacq_cell *out /* n_doppler */);

#ifdef __cplusplus
}
#endif
#endif /* GPSACQ_H */
