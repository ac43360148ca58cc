/*
 * oracle/fft_mixed.h -- TEST INFRASTRUCTURE ONLY (see fft_mixed_impl.h).
 * Float and double instances of the mixed-radix Stockham FFT.
 */
#ifndef ORACLE_FFT_MIXED_H
#define ORACLE_FFT_MIXED_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; }  cpx_f32;
typedef struct { double re, im; } cpx_f64;
typedef struct fft_plan_f32 fft_plan_f32;
typedef struct fft_plan_f64 fft_plan_f64;

fft_plan_f32 *fft_plan_create_f32(int n, int sign);
void fft_plan_destroy_f32(fft_plan_f32 *p);
void fft_execute_f32(const fft_plan_f32 *p, const cpx_f32 *in, cpx_f32 *out);

fft_plan_f64 *fft_plan_create_f64(int n, int sign);
void fft_plan_destroy_f64(fft_plan_f64 *p);
void fft_execute_f64(const fft_plan_f64 *p, const cpx_f64 *in, cpx_f64 *out);

#ifdef __cplusplus
}
#endif
#endif
