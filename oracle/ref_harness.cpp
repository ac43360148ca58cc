/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Wraps the UNMODIFIED reference translation unit (/root/reference/c/
 * search_offline.cpp, pulled in with #include so that its file-static
 * Sample()/Correlate()/code[]/fwd_buf[] are reachable) behind a few extern "C"
 * entry points, so that tests can drive the real reference block by block and
 * read its intermediate spectra.  Nothing of the reference is copied into this
 * repository: the include path points into /root/reference at build time and
 * the output goes to oracle/_ref/ (git-ignored).
 *
 * Build facts that are part of the oracle's definition:
 * compile with -fno-toplevel-reorder so the reference's 960-element overrun of
 * fwd_buf (search_offline.cpp:135-153) lands in rev_buf, where it is harmless.
 */
#include <stdint.h>
#include <string.h>

double FC, FS, max_fo;                 /* the caller-defined globals, c/gps_offline.h:23-25 */

#include "search_offline.cpp"          /* resolved through -I/root/reference/c */

extern "C" {

const char *oracle_fft_backend(void);

int ref_init(double fc, double fs, double maxfo)
{
    FC = fc; FS = fs; max_fo = maxfo;
    return SearchInit();
}

const char *ref_fft_backend(void) { return oracle_fft_backend(); }
int ref_fft_len(void) { return FFT_LEN; }

/* replica spectrum of PRN sv+1 as the reference holds it: 2*FFT_LEN floats */
void ref_get_code(int sv, float *out) { memcpy(out, code[sv], sizeof code[sv]); }

/* Sample() on one 5120-byte chunk held in memory; copies the spectrum out */
int ref_sample(const uint8_t *chunk, size_t nbytes, float *spectrum_out)
{
    FILE *fp = fmemopen((void *)chunk, nbytes, "rb");
    if (!fp) return -1;
    int rc = Sample(fp);
    fclose(fp);
    if (rc == 0 && spectrum_out) memcpy(spectrum_out, fwd_buf, sizeof fwd_buf);
    return rc;
}

/* The body of SearchTask()'s sv loop for n_blocks consecutive chunks:
 * Sample() then Correlate(sv) with sv = sv_of_block[b] (or b % 32). */
int ref_search_blocks(const uint8_t *bits, size_t nbytes, int n_blocks, const int32_t *sv_of_block,
                      float *snr, int32_t *lo_shift, int32_t *ca_shift)
{
    FILE *fp = fmemopen((void *)bits, nbytes, "rb");
    if (!fp) return -1;
    for (int b = 0; b < n_blocks; b++) {
        if (Sample(fp)) { fclose(fp); return b; }
        int lo = 0, ca = 0;
        int sv = sv_of_block ? sv_of_block[b] : b % NUM_SATS;
        snr[b] = Correlate(sv, &lo, &ca);
        lo_shift[b] = lo; ca_shift[b] = ca;
    }
    fclose(fp);
    return n_blocks;
}

void ref_search_task(const char *filename) { SearchTask((char *)filename); }
int ref_search_code(int sv, int g1) { return SearchCode(sv, g1); }

}
