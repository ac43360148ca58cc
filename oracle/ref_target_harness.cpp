/*
 * oracle/ref_target_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED on-target search loop and channel hand-off of the reference
 *     /root/reference/c/search.cpp   (SearchInit / Sample / Correlate / SearchTask, :76-239)
 *     /root/reference/c/channel.cpp  (ChanTask / ChanReset / ChanStart / CHANNEL::Start / Service / SignalLost, :134-254, :383-418)
 * on a chunk stream held in memory, so that the two "next" rows of SURVEY.md section 8(f) -- the acquisition ->
 * tracking hand-off (f3) and the re-acquisition service loop (f4) -- are pinned by the reference's own code and not by
 * a restatement.  Both files are compiled where they lie (oracle/Makefile, output in oracle/_ref/); nothing of them is
 * copied here.  What this file provides is the part of the receiver they link against and that needs the FPGA:
 *
 *   spi_set / spi_get      c/spi.cpp       -> a log of every command the reference sends, and CmdGetSamples served from the
 *                                            chunk stream (512-byte packets, 10 per Sample(), like the FPGA sampler)
 *   NextTask / TimerWait / Microseconds / CreateTask-like scheduling   c/coroutines.cpp -> the same round-robin of
 *                                            cooperative tasks (1 search task + NUM_CHANS channel tasks, main.cpp:67-68) on
 *                                            ucontext stacks, with a SIMULATED clock that advances a fixed step per yield
 *   Ephemeris[] / EPHEM::Subframe          c/ephemeris.cpp -> empty (no NAV data is ever decoded: every channel loses its
 *                                            signal after Acquisition() + Tracking()'s 20 s watchdog, :196,:205-233)
 *
 * ChanStart() is intercepted with the linker's --wrap so that its arguments (ch, sv, t_sample, taps, lo_shift, ca_shift)
 * are logged before the real function runs.  One run per process (the reference keeps its state in file statics).
 */
#include <ucontext.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "gps.h"          /* resolved through -I/root/reference/c */
#include "spi.h"
#include "ephemeris.h"

EPHEM Ephemeris[NUM_SATS];
void EPHEM::Subframe(char *) {}

namespace {

enum { KIND_START = 0, KIND_SPI = 1 };
struct Rec { int32_t kind, chunk, a, b, c, d, e; uint32_t u; };   /* 8 x 32 bit */

const int N_TASKS = 1 + NUM_CHANS;
const size_t STACK = 1 << 20;
ucontext_t g_main, g_task[N_TASKS];
char *g_stack[N_TASKS];
int g_cur = 0;
unsigned g_clock_us = 1000000, g_us_per_yield = 1000;
const uint8_t *g_feed = nullptr;
size_t g_feed_len = 0, g_feed_pos = 0;
int g_samples_triggered = 0;
std::vector<Rec> g_log;

void task_entry(int id)
{
    if (id == 0) SearchTask();        /* never returns */
    else ChanTask();                  /* never returns; takes channel number id-1 from its own static counter */
}

}  // namespace

/* ---- c/coroutines.cpp equivalents -------------------------------------------------------------- */
void NextTask()
{
    g_clock_us += g_us_per_yield;
    const int prev = g_cur;
    g_cur = (g_cur + 1) % N_TASKS;
    swapcontext(&g_task[prev], &g_task[g_cur]);
}

unsigned Microseconds(void) { return g_clock_us; }

void TimerWait(unsigned ms)          /* same loop as c/coroutines.cpp:47-54 on the simulated clock */
{
    const unsigned finish = Microseconds() + 1000 * ms;
    for (;;) {
        NextTask();
        const int diff = (int)(finish - Microseconds());
        if (diff <= 0) break;
    }
}

/* ---- c/spi.cpp equivalents ---------------------------------------------------------------------- */
void spi_set(SPI_CMD cmd, uint16_t wparam, uint32_t lparam)
{
    if (cmd == CmdSample) g_samples_triggered++;
    Rec r = {KIND_SPI, g_samples_triggered - 1, (int32_t)cmd, (int32_t)wparam, 0, 0, 0, lparam};
    g_log.push_back(r);
}

void spi_get(SPI_CMD cmd, SPI_MISO *rx, int bytes, uint16_t)
{
    if (cmd == CmdGetSamples) {
        if (g_feed_pos + (size_t)bytes > g_feed_len) setcontext(&g_main);     /* stream exhausted: the run ends here */
        memcpy(rx->byte, g_feed + g_feed_pos, (size_t)bytes);
        g_feed_pos += (size_t)bytes;
    } else {
        memset(rx->byte, 0, (size_t)bytes);                                    /* channel state upload: no signal, no NAV bits */
    }
}

void spi_hog(SPI_CMD, SPI_MISO *rx, int bytes) { memset(rx->byte, 0, (size_t)bytes); }

/* ---- ChanStart(), logged (ld --wrap=_Z9ChanStartiiiiii) ------------------------------------------- */
extern "C" void __real__Z9ChanStartiiiiii(int, int, int, int, int, int);
extern "C" void __wrap__Z9ChanStartiiiiii(int ch, int sv, int t_sample, int taps, int lo_shift, int ca_shift)
{
    Rec r = {KIND_START, g_samples_triggered - 1, ch, sv, taps, lo_shift, ca_shift, g_clock_us - (unsigned)t_sample};
    g_log.push_back(r);
    __real__Z9ChanStartiiiiii(ch, sv, t_sample, taps, lo_shift, ca_shift);
}

extern "C" {

double reft_fc(void) { return FC; }
double reft_fs(void) { return FS; }
int reft_num_chans(void) { return NUM_CHANS; }

/* Run SearchInit() and then the receiver's tasks over `bits` until the sampler runs dry.  Returns the number of log
 * records; up to max_rec of them are copied to out (8 int32 each):
 *   kind 0 (ChanStart):  chunk, ch, sv, taps, lo_shift, ca_shift, microseconds since t_sample
 *   kind 1 (spi_set):    chunk, cmd, wparam, -, -, -, lparam
 * chunk = index of the most recently triggered Sample() (0-based). */
int reft_run(const uint8_t *bits, size_t n_bytes, unsigned us_per_yield, int32_t *out, int max_rec)
{
    static bool used = false;
    if (used) return -1;
    used = true;
    g_feed = bits; g_feed_len = n_bytes; g_feed_pos = 0; g_us_per_yield = us_per_yield;
    if (SearchInit() != 0) return -2;
    volatile bool started = false;
    getcontext(&g_main);
    if (!started) {
        started = true;
        for (int i = 0; i < N_TASKS; i++) {
            g_stack[i] = (char *)malloc(STACK);
            getcontext(&g_task[i]);
            g_task[i].uc_stack.ss_sp = g_stack[i];
            g_task[i].uc_stack.ss_size = STACK;
            g_task[i].uc_link = &g_main;
            makecontext(&g_task[i], (void (*)())task_entry, 1, i);
        }
        g_cur = 0;
        setcontext(&g_task[0]);
    }
    /* back here when spi_get() found the stream exhausted */
    const int n = (int)g_log.size();
    for (int i = 0; i < n && i < max_rec; i++) memcpy(out + 8 * i, &g_log[i], sizeof(Rec));
    return n;
}

}
