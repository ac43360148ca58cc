/*
 * oracle/shim/fftw3_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Backend of the fftw3.h stand-in: lets the UNMODIFIED reference
 * (c/search_offline.cpp) link and run without libfftw3f.  Plans wrap our own
 * float mixed-radix FFT (oracle/fft_mixed.c).  If the environment variable
 * ORACLE_FFT=mkl is set and MKL's DFTI entry points can be found (they are
 * exported by torch's libtorch_cpu.so; path in ORACLE_MKL_LIB), MKL is used
 * instead -- that is the strongest CPU FFT on the box and is what the
 * "reference" CPU baseline should be timed with when available.
 */
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include "fftw3.h"
#include "../fft_mixed.h"

struct oracle_fftwf_plan_s {
    int n, sign;
    fftwf_complex *in, *out;
    fft_plan_f32 *own;
    void *dfti;                 /* DFTI_DESCRIPTOR_HANDLE when MKL backend  */
};

/* ---- optional MKL DFTI backend, resolved at run time ------------------- */
typedef long (*dfti_create_t)(void **, int, long);   /* DftiCreateDescriptor_s_1d(handle*, domain, n) */
typedef long (*dfti_commit_t)(void *);
typedef long (*dfti_compute_t)(void *, void *, ...);
typedef long (*dfti_free_t)(void **);
static dfti_create_t  p_create;
static dfti_commit_t  p_commit;
static dfti_compute_t p_fwd, p_bwd;
static dfti_free_t    p_free;
static int mkl_state;           /* 0 = untried, 1 = ok, -1 = unavailable   */

enum { DFTI_COMPLEX = 32 };

static int mkl_try(void)
{
    if (mkl_state) return mkl_state > 0;
    mkl_state = -1;
    const char *want = getenv("ORACLE_FFT");
    if (!want || strcmp(want, "mkl") != 0) return 0;
    const char *lib = getenv("ORACLE_MKL_LIB");
    void *h = dlopen(lib ? lib : "libtorch_cpu.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return 0;
    p_create = (dfti_create_t)dlsym(h, "DftiCreateDescriptor_s_1d");
    p_commit = (dfti_commit_t)dlsym(h, "DftiCommitDescriptor");
    p_fwd    = (dfti_compute_t)dlsym(h, "DftiComputeForward");
    p_bwd    = (dfti_compute_t)dlsym(h, "DftiComputeBackward");
    p_free   = (dfti_free_t)dlsym(h, "DftiFreeDescriptor");
    if (p_create && p_commit && p_fwd && p_bwd && p_free) mkl_state = 1;
    return mkl_state > 0;
}

const char *oracle_fft_backend(void)
{
    return mkl_try() ? "mkl-dfti" : "builtin-f32";
}

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->n = n; p->sign = sign; p->in = in; p->out = out;
    if (mkl_try() && in == out) {
        void *d = NULL;
        if (p_create(&d, DFTI_COMPLEX, (long)n) == 0 && p_commit(d) == 0) {
            p->dfti = d;            /* default placement is in-place          */
            return p;
        }
    }
    p->own = fft_plan_create_f32(n, sign);
    if (!p->own) { free(p); return NULL; }
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    if (p->dfti) {
        if (p->sign < 0) p_fwd(p->dfti, (void *)p->in);
        else             p_bwd(p->dfti, (void *)p->in);
        return;
    }
    fft_execute_f32(p->own, (const cpx_f32 *)p->in, (cpx_f32 *)p->out);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    if (p->dfti) p_free(&p->dfti);
    if (p->own) fft_plan_destroy_f32(p->own);
    free(p);
}
