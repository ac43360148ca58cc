// tools/ubench/fp32x2.cu -- micro-benchmark of the packed FP32x2 instructions the engine is built on (FFMA2 / FADD2),
// to be run on a B200:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu && ./fp32x2
//
// Questions it answers: how many cycles one FFMA2 holds the FMA pipe, its dependent-issue
// latency, whether scalar FFMA/FADD can issue next to it (fmalite pipe), and how many independent chains per warp /
// warps per sub-partition are needed to saturate the pipe.  Prints cycles per instruction per sub-partition.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// MODE 0: FFMA2 only; 1: FADD2 only; 2: scalar FFMA only; 3: FFMA2 + scalar FFMA interleaved 1:1; 4: FFMA2 + scalar FADD 1:1
template <int CHAINS, int MODE>
__global__ void kern(float2 *out, int iters, long long *cycles)
{
    float2 a[CHAINS];
    float s[CHAINS];
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-9f, -1e-9f);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = make_float2(threadIdx.x + i, 1.0f + i); s[i] = 0.5f + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0 || MODE >= 3) a[i] = __ffma2_rn(a[i], m, c);
            if (MODE == 1) a[i] = __fadd2_rn(a[i], c);
            if (MODE == 2 || MODE == 3) s[i] = fmaf(s[i], 1.0000001f, 1e-9f);
            if (MODE == 4) s[i] = s[i] + 1e-9f;
        }
    }
    const long long t1 = clock64();
    float2 r = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { r.x += a[i].x + s[i]; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS, int MODE>
static void run(const char *what, int warps_per_smsp, float2 *d_out, long long *d_cyc)
{
    const int iters = 4096, threads = 128 * warps_per_smsp;     // 4 sub-partitions per SM, one CTA on one SM
    kern<CHAINS, MODE><<<1, threads>>>(d_out, iters, d_cyc);
    CHECK(cudaDeviceSynchronize());
    long long cyc = 0;
    CHECK(cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    const int per_iter = CHAINS * ((MODE >= 3) ? 2 : 1);
    const double inst_per_smsp = (double)iters * per_iter * warps_per_smsp;
    printf("%-28s chains %2d  warps/SMSP %d : %6.3f cycles per warp-instruction per sub-partition\n", what, CHAINS, warps_per_smsp,
           (double)cyc / inst_per_smsp);
}

int main()
{
    float2 *d_out; long long *d_cyc;
    CHECK(cudaMalloc(&d_out, 1024 * sizeof(float2)));
    CHECK(cudaMalloc(&d_cyc, sizeof(long long)));
    for (int w = 1; w <= 8; w *= 2) {
        run<1, 0>("FFMA2 (dependent chain)", w, d_out, d_cyc);
        run<4, 0>("FFMA2", w, d_out, d_cyc);
        run<8, 0>("FFMA2", w, d_out, d_cyc);
        run<8, 1>("FADD2", w, d_out, d_cyc);
        run<8, 2>("FFMA (scalar)", w, d_out, d_cyc);
        run<8, 3>("FFMA2 + FFMA 1:1", w, d_out, d_cyc);
        run<8, 4>("FFMA2 + FADD 1:1", w, d_out, d_cyc);
    }
    return 0;
}
