// ga_service.h -- host side of the two consumers of the acquisition records:
//
//   gpsacq_handoff_compute   CHANNEL::Start(), c/channel.cpp:134-171 -- Doppler from the FFT bin shift, carrier and
//                            code NCO rates, code creep since the sample, code-generator pause, Gold-code taps
//                            ((T1<<4)+T2, the ChanStart() argument of c/search.cpp:236-237)
//   gpsacq_service_*         the receiver's SearchTask() loop, c/search.cpp:214-239: round-robin over the 32 SVs,
//                            skip the ones already tracked (Busy[], :55 / SearchEnable :207-209), one fresh chunk
//                            per searched SV, wait while all NUM_CHANS channels are busy (ChanReset, c/channel.cpp:
//                            396-404), on snr >= 25 mark the SV busy and start a channel; CHANNEL::SignalLost()
//                            (c/channel.cpp:245-254) frees the channel and re-enables the SV.
//
// The reference runs that loop one chunk at a time.  Here chunks are searched in batches on the GPU, so the loop is
// run SPECULATIVELY: the chunk -> SV assignment of the next batch is laid out as if nothing in it were detected;
// results are then scanned in order and at the first detection everything after it is thrown away (a detection
// removes the SV from all later rounds, which shifts every later assignment) and re-searched in the next batch.
// The batch size adapts: one round after a detection, doubling while nothing is found.  The event sequence is
// exactly the sequential loop's (tests/test_gpu_service.py checks it against the oracle's one-chunk-at-a-time loop
// and against itself for different batch limits).
//
// Included at the end of gpsacq.cu (host code only).
#pragma once

struct gpsacq_service {
    gpsacq *h;
    int num_chans, max_rounds;
    bool busy[GPSACQ_NUM_SATS];
    unsigned chan_busy;                 // BusyFlags
    int next_sv;                        // position of the round-robin cursor
    int rounds;                         // current speculation depth, in rounds of 32
    long long chunks_seen;              // chunks consumed since creation (= Sample() calls of the sequential loop)
    std::vector<int32_t> sv;
    std::vector<gpsacq_peak> peaks;
    std::string err;
};

extern "C" {

int gpsacq_handoff_compute(const gpsacq_peak *p, double fc, double fs, double bin_num, double bin_den,
                           double secs_since_sample, gpsacq_handoff *out)
{
    if (!p || !out || !(fs > 0) || !(bin_den != 0) || p->sv < 0 || p->sv >= GPSACQ_NUM_SATS) return GPSACQ_EINVAL;
    const double L1 = 1575.42e6;                                   // c/gps.h:22
    memset(out, 0, sizeof *out);
    // Estimate Doppler from FFT bin shift: lo_shift*FS/FFT_LEN, evaluated left to right (c/channel.cpp:147-148)
    const double lo_dop = p->lo_shift * bin_num / bin_den;
    const double ca_dop = lo_dop / L1 * kCPS;
    // NCO rates (:151-152)
    out->lo_rate = (uint32_t)((fc + lo_dop) / fs * pow(2, 32));
    out->ca_rate = (uint32_t)((kCPS + ca_dop) / fs * pow(2, 32));
    // Code creep due to code rate Doppler (:158-161)
    int ca_shift = p->ca_shift;
    ca_shift += (int)nearbyint(ca_dop * secs_since_sample * fs / kCPS);
    // Align code generator by pausing NCO (:164; 20000 / 10000 there are 2 and 1 code periods at FS = 10 MHz)
    const int w = (int)ceil(fs / 1000.0);
    out->ca_pause = (uint32_t)((2 * w - ca_shift) % w);
    out->ca_shift = ca_shift;
    out->lo_dop_hz = lo_dop;
    out->ca_dop_hz = ca_dop;
    out->taps = (kTaps[p->sv][0] << 4) + kTaps[p->sv][1];          // c/search.cpp:236-237
    out->sv = p->sv;
    return GPSACQ_OK;
}

int gpsacq_service_create(gpsacq_t *h, int num_chans, int max_rounds_per_batch, gpsacq_service_t **out)
{
    if (!h || !out) return GPSACQ_EINVAL;
    if (h->mode != GPSACQ_MODE_REF) { h->err = "the search service needs a GPSACQ_MODE_REF handle"; return GPSACQ_EINVAL; }
    gpsacq_service *s = new (std::nothrow) gpsacq_service();
    if (!s) return GPSACQ_ENOMEM;
    s->h = h;
    s->num_chans = num_chans > 0 ? std::min(num_chans, 32) : 12;    // NUM_CHANS, c/gps.h:17
    const int cap_rounds = std::max(1, h->cap / GPSACQ_NUM_SATS);
    s->max_rounds = max_rounds_per_batch > 0 ? std::min(max_rounds_per_batch, cap_rounds) : cap_rounds;
    memset(s->busy, 0, sizeof s->busy);
    s->chan_busy = 0; s->next_sv = 0; s->rounds = 1; s->chunks_seen = 0;
    *out = s;
    return GPSACQ_OK;
}

void gpsacq_service_destroy(gpsacq_service_t *s) { delete s; }
const char *gpsacq_service_last_error(const gpsacq_service_t *s) { return s ? s->err.c_str() : ""; }

int gpsacq_service_enable(gpsacq_service_t *s, int sv)            // SearchEnable(), c/search.cpp:207-209
{
    if (!s || sv < 0 || sv >= GPSACQ_NUM_SATS) return GPSACQ_EINVAL;
    s->busy[sv] = false;
    return GPSACQ_OK;
}

int gpsacq_service_signal_lost(gpsacq_service_t *s, int ch)       // CHANNEL::SignalLost(), c/channel.cpp:245-254
{
    if (!s || ch < 0 || ch >= s->num_chans) return GPSACQ_EINVAL;
    s->chan_busy &= ~(1u << ch);
    return GPSACQ_OK;
}

int gpsacq_service_state(const gpsacq_service_t *s, uint32_t *busy_svs, uint32_t *busy_chans, int64_t *chunks_seen)
{
    if (!s) return GPSACQ_EINVAL;
    uint32_t m = 0;
    for (int i = 0; i < GPSACQ_NUM_SATS; i++) m |= s->busy[i] ? (1u << i) : 0u;
    if (busy_svs) *busy_svs = m;
    if (busy_chans) *busy_chans = s->chan_busy;
    if (chunks_seen) *chunks_seen = s->chunks_seen;
    return GPSACQ_OK;
}

int gpsacq_service_feed(gpsacq_service_t *s, const uint8_t *chunks, size_t n_chunks, size_t *consumed,
                        gpsacq_event *events, size_t max_events, size_t *n_events)
{
    if (!s || (!chunks && n_chunks) || !consumed || !n_events || (!events && max_events)) return GPSACQ_EINVAL;
    gpsacq *h = s->h;
    *consumed = 0; *n_events = 0;
    while (*consumed < n_chunks && *n_events < max_events) {
        // while((ch=ChanReset())<0) NextTask();  -- all channels busy: nothing is sampled (c/search.cpp:223-224)
        int ch = -1;
        for (int c = 0; c < s->num_chans; c++) if (!(s->chan_busy & (1u << c))) { ch = c; break; }
        if (ch < 0) break;
        int n_free = 0;
        for (int i = 0; i < GPSACQ_NUM_SATS; i++) n_free += s->busy[i] ? 0 : 1;
        if (n_free == 0) break;                                    // every SV is being tracked: the loop spins without sampling
        // lay out the next batch as the sequential loop would consume it if nothing in it were detected
        const size_t want = std::min(n_chunks - *consumed, (size_t)s->rounds * (size_t)n_free);
        s->sv.resize(want);
        int cur = s->next_sv;
        for (size_t i = 0; i < want; i++) {
            while (s->busy[cur]) cur = (cur + 1) % GPSACQ_NUM_SATS;
            s->sv[i] = cur;
            cur = (cur + 1) % GPSACQ_NUM_SATS;
        }
        s->peaks.resize(want);
        const int rc = gpsacq_search_blocks(h, chunks + *consumed * (size_t)h->chunk_bytes, want, s->sv.data(), s->peaks.data());
        if (rc) { s->err = h->err; return rc; }
        size_t used = want;
        bool hit = false;
        for (size_t i = 0; i < want; i++) {
            const gpsacq_peak &p = s->peaks[i];
            if (p.snr < 25.0f) continue;                           // if (snr<25) continue;  (:232-233)
            // Busy[sv] = true; ChanStart(ch, sv, t_sample, taps, lo_shift, ca_shift);  (:235-237)
            gpsacq_event &e = events[(*n_events)++];
            memset(&e, 0, sizeof e);
            e.chunk_index = s->chunks_seen + (long long)i;
            e.sv = p.sv; e.ch = ch; e.peak = p;
            // the tracking channel starts on the first sample after the chunk: secs = chunk duration
            gpsacq_handoff_compute(&p, h->cfg.fc, h->cfg.fs, h->cfg.fs, (double)h->n, (double)h->chunk_samples / h->cfg.fs, &e.start);
            s->busy[p.sv] = true;
            s->chan_busy |= 1u << ch;
            used = i + 1;                                          // everything after it was laid out with a stale Busy[]
            hit = true;
            break;
        }
        s->next_sv = (s->sv[used - 1] + 1) % GPSACQ_NUM_SATS;
        s->chunks_seen += (long long)used;
        *consumed += used;
        s->rounds = hit ? 1 : std::min(s->rounds * 2, s->max_rounds);
    }
    return GPSACQ_OK;
}

}  // extern "C"
