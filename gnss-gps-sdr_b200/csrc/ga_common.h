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
GA_HD cf cadd(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
GA_HD cf csub(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
GA_HD cf cmul(cf a, cf b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GA_HD cf csqr(cf a) { return mk(a.x * a.x - a.y * a.y, 2.0f * a.x * a.y); }
GA_HD cf cconj(cf a) { return mk(a.x, -a.y); }
GA_HD cf cscale(cf a, float s) { return mk(a.x * s, a.y * s); }
// multiply by DIR*i  (DIR=+1: i*a ; DIR=-1: -i*a)
template <int DIR> GA_HD cf cmul_i(cf a) { return DIR > 0 ? mk(-a.y, a.x) : mk(a.y, -a.x); }
// acc += a*b
GA_HD void cfma(cf &acc, cf a, cf b)
{
    acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}

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
