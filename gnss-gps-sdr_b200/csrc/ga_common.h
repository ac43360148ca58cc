// ga_common.h -- shared host/device helpers for the B200 GPS acquisition engine.
//
// Everything in this header compiles both under nvcc (device code for sm_100a)
// and under a plain host C++17 compiler; the host build exists only so that
// tests/emu can replay the exact per-thread butterfly / index arithmetic on the
// CPU box (there is no GPU in the development container).  The product path is
// the nvcc build.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define GA_HD __host__ __device__ __forceinline__
#define GA_D __device__ __forceinline__
#define GA_UNROLL _Pragma("unroll")
#else
#define GA_HD inline
#define GA_D inline
#define GA_UNROLL
#endif

namespace ga {

#if defined(__CUDACC__)
typedef float2 cf;
#else
struct alignas(8) cf { float x, y; };
#endif

GA_HD cf mk(float x, float y) { cf r; r.x = x; r.y = y; return r; }

// Complex arithmetic on (re, im) pairs.  On sm_100a every operation below is written with the
// packed FP32x2 intrinsics (__fadd2_rn / __fmul2_rn / __ffma2_rn -> SASS FADD2 / FMUL2 / FFMA2,
// new on Blackwell): one issue slot does both components, the half-swap and the sign flip of a
// multiplication by +-i are operand modifiers (.LO_HI, .NP), so a complex add/sub -- with or
// without a factor i -- is ONE instruction and a complex multiply TWO.  The host build (CPU
// replay in tests/emu) uses the plain scalar formulas.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
#define GA_PACKED 1
GA_HD cf cadd(cf a, cf b) { return __fadd2_rn(a, b); }
GA_HD cf csub(cf a, cf b) { return __fadd2_rn(a, mk(-b.x, -b.y)); }
// The same sums as two scalar FADDs (identical bits; __fadd_rn is never contracted).  A packed instruction holds the
// FMA-heavy pipe for two cycles and a scalar FADD issues beside it for free (tools/ubench/fp32x2.cu on a B200:
// FFMA2 2.1 cycles per warp-instruction per sub-partition, FFMA2 + FADD 1:1 1.04 -- profiles/r02_fp32x2.txt), at the
// price of a second issue slot: the butterflies pick per call site (GA_R4_SCALAR, GA_R5_SCALAR).
GA_HD cf cadd_s(cf a, cf b) { return mk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
GA_HD cf csub_s(cf a, cf b) { return mk(__fadd_rn(a.x, -b.x), __fadd_rn(a.y, -b.y)); }
template <int DIR> GA_HD cf cadd_i_s(cf a, cf b) { return DIR > 0 ? mk(__fadd_rn(a.x, -b.y), __fadd_rn(a.y, b.x)) : mk(__fadd_rn(a.x, b.y), __fadd_rn(a.y, -b.x)); }
template <int DIR> GA_HD cf csub_i_s(cf a, cf b) { return DIR > 0 ? mk(__fadd_rn(a.x, b.y), __fadd_rn(a.y, -b.x)) : mk(__fadd_rn(a.x, -b.y), __fadd_rn(a.y, b.x)); }
GA_HD cf cmul(cf a, cf b) { return __ffma2_rn(mk(a.y, a.x), mk(-b.y, b.y), __fmul2_rn(a, mk(b.x, b.x))); }
GA_HD cf csqr(cf a) { return cmul(a, a); }
GA_HD cf cscale(cf a, float s) { return __fmul2_rn(a, mk(s, s)); }
// acc + a*s  (s real)
GA_HD cf caxpy(cf acc, cf a, float s) { return __ffma2_rn(a, mk(s, s), acc); }
// a + DIR*i*b   and   a - DIR*i*b
template <int DIR> GA_HD cf cadd_i(cf a, cf b) { return DIR > 0 ? __fadd2_rn(a, mk(-b.y, b.x)) : __fadd2_rn(a, mk(b.y, -b.x)); }
template <int DIR> GA_HD cf csub_i(cf a, cf b) { return DIR > 0 ? __fadd2_rn(a, mk(b.y, -b.x)) : __fadd2_rn(a, mk(-b.y, b.x)); }
// acc += a*b
GA_HD void cfma(cf &acc, cf a, cf b) { acc = __ffma2_rn(mk(a.y, a.x), mk(-b.y, b.y), __ffma2_rn(a, mk(b.x, b.x), acc)); }
#else
#define GA_PACKED 0
GA_HD cf cadd(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
GA_HD cf csub(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
GA_HD cf cadd_s(cf a, cf b) { return cadd(a, b); }
GA_HD cf csub_s(cf a, cf b) { return csub(a, b); }
GA_HD cf cmul(cf a, cf b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GA_HD cf csqr(cf a) { return mk(a.x * a.x - a.y * a.y, 2.0f * a.x * a.y); }
GA_HD cf cscale(cf a, float s) { return mk(a.x * s, a.y * s); }
GA_HD cf caxpy(cf acc, cf a, float s) { return mk(fmaf(a.x, s, acc.x), fmaf(a.y, s, acc.y)); }
template <int DIR> GA_HD cf cadd_i(cf a, cf b) { return DIR > 0 ? mk(a.x - b.y, a.y + b.x) : mk(a.x + b.y, a.y - b.x); }
template <int DIR> GA_HD cf csub_i(cf a, cf b) { return DIR > 0 ? mk(a.x + b.y, a.y - b.x) : mk(a.x - b.y, a.y + b.x); }
template <int DIR> GA_HD cf cadd_i_s(cf a, cf b) { return cadd_i<DIR>(a, b); }
template <int DIR> GA_HD cf csub_i_s(cf a, cf b) { return csub_i<DIR>(a, b); }
GA_HD void cfma(cf &acc, cf a, cf b)
{
    acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
#endif
GA_HD cf cconj(cf a) { return mk(a.x, -a.y); }

// read-only global load
GA_HD cf ldg(const cf *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace ga
