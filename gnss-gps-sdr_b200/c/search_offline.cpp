// search_offline.cpp -- SearchInit / SearchFree / SearchTask / SearchEnable / SearchCode on
// top of the libgpsacq C ABI (include/gpsacq.h): the drop-in replacement of the reference's
// c/search_offline.cpp for the offline acquisition path.  All signal processing happens in
// hand-written sm_100a kernels; this file is file traversal, batching and the stdout report.
//
// Behaviour kept from the reference (line numbers in /root/reference/c/search_offline.cpp):
//   * one "run" = 32 consecutive 5120-byte chunks, chunk k of a run searched for PRN k+1 (:239-246)
//   * a run that hits end of file is discarded after printing "run out of file!" (:241-244,:260-262)
//   * detection rule snr >= 25, hits kept in PRN order, 0-based sv printed (:248-257,:264-287)
//   * "can not open file!" on fopen failure (:224-228)
// Optional environment knobs (absent = reference behaviour):
//   GPSACQ_DEVICE=<ordinal>      first CUDA device to use (default 0)
//   GPSACQ_GPUS=<n>              shard each batch's chunks over n GPUs (default 1); the 32-byte peak
//                                records come back with one ncclAllGather per batch (GPSACQ_GATHER=host
//                                gathers through host memory instead)
//   GPSACQ_RUNS_PER_BATCH=<r>    runs handed to the GPU per call (default 16)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "gps_offline.h"
#include "cacode.h"
#include "../../include/gpsacq.h"

namespace {

struct SatTap { int prn, t0, t1; };
// G2 tap pairs of PRN 1..32 (IS-GPS-200 Table 3-Ia; same table as c/search_offline.cpp:20-53)
const SatTap kSats[NUM_SATS] = {
    {1, 2, 6},   {2, 3, 7},   {3, 4, 8},   {4, 5, 9},   {5, 1, 9},   {6, 2, 10},  {7, 1, 8},   {8, 2, 9},
    {9, 3, 10},  {10, 2, 3},  {11, 3, 4},  {12, 5, 6},  {13, 6, 7},  {14, 7, 8},  {15, 8, 9},  {16, 9, 10},
    {17, 1, 4},  {18, 2, 5},  {19, 3, 6},  {20, 4, 7},  {21, 5, 8},  {22, 6, 9},  {23, 1, 3},  {24, 4, 6},
    {25, 5, 7},  {26, 6, 8},  {27, 7, 9},  {28, 8, 10}, {29, 1, 6},  {30, 2, 7},  {31, 3, 8},  {32, 4, 9},
};

bool g_busy[NUM_SATS];
gpsacq_group_t *g_group = NULL;           // one engine per GPU + the peak-record gather (NCCL or host)
int g_chunk_bytes = 0;
int g_runs_per_batch = 16;
// The reference reads FS in SearchInit() for the replicas (:76) and FC / FS / max_fo again on every Sample() and
// Correlate() (:127,:176,:190).  The engine bakes them into device tables, so they are remembered here and the engine
// is rebuilt when SearchTask() finds the globals changed (replicas stay at the SearchInit()-time FS, like there).
double g_fs_init = 0, g_fc_used = 0, g_fs_used = 0, g_maxfo_used = 0;

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

void print_run(int run_count, const gpsacq_peak *pk)
{
    int hit[NUM_SATS], nh = 0;
    for (int sv = 0; sv < NUM_SATS; sv++) if (!(pk[sv].snr < 25)) hit[nh++] = sv;
    printf("%2d satellite: ", run_count);
    for (int i = 0; i < nh; i++) printf("%5d ", hit[i]);
    printf("\n%2d SNR(>=25): ", run_count);
    for (int i = 0; i < nh; i++) printf("%5.1f ", pk[hit[i]].snr);
    printf("\n%2d  lo_shift: ", run_count);
    for (int i = 0; i < nh; i++) printf("%5d ", pk[hit[i]].lo_shift);
    printf("\n%2d  ca_shift: ", run_count);
    for (int i = 0; i < nh; i++) printf("%5d ", pk[hit[i]].ca_shift);
    printf("\n");
    for (int sv = 0; sv < NUM_SATS; sv++) printf("%2.0f ", pk[sv].snr);
    printf("\n\n");
}

}  // namespace

static int build_engine(double fs_replica)
{
    if (g_group) gpsacq_group_destroy(g_group);
    g_group = NULL;
    const int first = env_int("GPSACQ_DEVICE", 0);
    int ngpu = env_int("GPSACQ_GPUS", 1);
    if (ngpu < 1) ngpu = 1;
    g_runs_per_batch = env_int("GPSACQ_RUNS_PER_BATCH", 16);
    if (g_runs_per_batch < 1) g_runs_per_batch = 1;
    const char *gather = getenv("GPSACQ_GATHER");
    gpsacq_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.fc = FC; cfg.fs = FS; cfg.max_fo = max_fo;
    cfg.fs_replica = fs_replica;
    cfg.fft_len = FFT_LEN;
    cfg.mode = GPSACQ_MODE_REF;
    cfg.max_blocks = g_runs_per_batch * NUM_SATS;      // per GPU
    std::vector<int32_t> devs(ngpu);
    for (int g = 0; g < ngpu; g++) devs[g] = first + g;
    const int rc = gpsacq_group_create(&cfg, ngpu, devs.data(), !(gather && strcmp(gather, "host") == 0), &g_group);
    if (rc != GPSACQ_OK) {
        fprintf(stderr, "gpsacq_group_create: %s\n", gpsacq_last_error(NULL));
        g_group = NULL;
        return rc;
    }
    g_runs_per_batch *= ngpu;                           // a host batch feeds every GPU a full share
    gpsacq_info info;
    gpsacq_get_info(gpsacq_group_engine(g_group, 0), &info);
    g_chunk_bytes = info.chunk_bytes;
    g_fc_used = FC; g_fs_used = FS; g_maxfo_used = max_fo;
    return 0;
}

int SearchInit()
{
    g_fs_init = FS;
    return build_engine(0.0);
}

void SearchFree()
{
    if (g_group) gpsacq_group_destroy(g_group);
    g_group = NULL;
}

void SearchEnable(int sv)
{
    if (sv >= 0 && sv < NUM_SATS) g_busy[sv] = false;
}

// The reference walks the LFSR until GetG1() == g1 and never returns for a value the register
// cannot take (0, or anything above 10 bits); this version gives -1 after one full period.
int SearchCode(int sv, int g1)
{
    if (sv < 0 || sv >= NUM_SATS) return -1;
    CACODE ca(kSats[sv].t0, kSats[sv].t1);
    for (int chips = 0; chips < 1023; chips++, ca.Clock())
        if (ca.GetG1() == (unsigned)g1) return chips;
    return -1;
}

void SearchTask(char *filename_1bit_bin)
{
    FILE *fp = fopen(filename_1bit_bin, "rb");
    if (!fp) { printf("can not open file!\n"); return; }
    if (!g_group) { fclose(fp); fprintf(stderr, "SearchTask: SearchInit() has not succeeded\n"); return; }
    if (FC != g_fc_used || FS != g_fs_used || max_fo != g_maxfo_used) {      // globals changed since the engine was built
        const int rc = build_engine(g_fs_init);
        if (rc) { fclose(fp); fprintf(stderr, "SearchTask: engine rebuild for changed FC/FS/max_fo failed (%d)\n", rc); return; }
    }

    const size_t run_bytes = (size_t)NUM_SATS * g_chunk_bytes;
    std::vector<unsigned char> buf(run_bytes * g_runs_per_batch);
    std::vector<gpsacq_peak> peaks((size_t)NUM_SATS * g_runs_per_batch);
    int run_count = 0;
    for (;;) {
        const size_t got = fread(buf.data(), 1, buf.size(), fp);
        const size_t full_runs = got / run_bytes;
        if (full_runs) {
            const int rc = gpsacq_group_search_blocks(g_group, buf.data(), full_runs * NUM_SATS, peaks.data());
            if (rc) { fprintf(stderr, "gpsacq_group_search_blocks: %s\n", gpsacq_group_last_error(g_group)); break; }
            for (size_t r = 0; r < full_runs; r++) print_run(run_count++, &peaks[r * NUM_SATS]);
        }
        if (got < buf.size()) { printf("run out of file!\n"); break; }
    }
    fclose(fp);
}
