/*
 * oracle/fft_mixed.c -- TEST INFRASTRUCTURE ONLY.
 * Instantiates fft_mixed_impl.h for float and double.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fft_mixed.h"

#define REAL float
#define SUF(x) x##_f32
#include "fft_mixed_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF(x) x##_f64
#include "fft_mixed_impl.h"
#undef REAL
#undef SUF
