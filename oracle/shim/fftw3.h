/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Stand-in for the single-precision FFTW3 API, so that the reference's
 * c/search_offline.cpp (which does `#include <fftw3.h>`, search_offline.cpp:8)
 * compiles UNMODIFIED in a container that has no libfftw3f.  Only the five
 * symbols the reference touches are provided (call sites: search_offline.cpp
 * :78,:79 plan, :105,:161,:187 execute, :115,:116 destroy).
 *
 * Semantics kept: c2c, 1-D, unnormalised, in-place capable, FFTW_FORWARD = -1
 * exponent sign, FFTW_BACKWARD = +1.  The backend is oracle/shim/fftw3_shim.c
 * (our own mixed-radix Stockham FFT, or MKL DFTI when ORACLE_FFT=mkl).
 *
 * The reference relies on the real fftw3.h pulling in <stdio.h> (it uses
 * FILE/fopen/printf without including it), so this header does too.
 */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H

#include <stdio.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct oracle_fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out,
                             int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

/* not part of FFTW: reports which backend the shim is using ("builtin-f32",
 * "mkl-dfti") so that CPU-baseline numbers can say what they were timed on. */
const char *oracle_fft_backend(void);

#ifdef __cplusplus
}
#endif
#endif
