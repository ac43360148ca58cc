// warpmap.cu -- which SM sub-partition do the warps of co-resident CTAs land on?
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/ubench/warpmap.cu -o build/warpmap && build/warpmap
// (argv[1] = threads per CTA, default 128; up to 8 warps are printed)
// Launches 3 x 148 CTAs of 128 threads with 67 KB of dynamic shared memory each (the REF cell kernel's shape), keeps
// them resident for a while and prints, for a few SMs, the arrival slot of every CTA and the hardware warp ids
// (%warpid) of its four warps.  %warpid mod 4 is the sub-partition.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ unsigned slots[1024];
__global__ void k(unsigned *out, long long spin)
{
    extern __shared__ unsigned char sm[];
    __shared__ unsigned slot;
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if (threadIdx.x == 0) slot = atomicAdd(&slots[smid], 1u);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        unsigned *o = out + (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4;
        o[0] = smid; o[1] = slot; o[2] = wid; o[3] = threadIdx.x >> 5;
    }
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { sm[threadIdx.x] += 1; }
}
int main(int argc, char **argv)
{
    const int grid = 444, smem = 67360, T = argc > 1 ? atoi(argv[1]) : 128, NW = T / 32;
    unsigned *d, *h = new unsigned[grid * 32];
    cudaMalloc(&d, grid * 32 * sizeof(unsigned));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) {
        k<<<grid, T, smem>>>(d, 2000000);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, grid * 32 * sizeof(unsigned), cudaMemcpyDeviceToHost);
        printf("launch %d (%s)\n", rep, cudaGetErrorString(cudaGetLastError()));
        int hist[4][4] = {{0}};      // [slot%4... ] (warp-in-CTA 0's sub-partition by slot)
        for (int b = 0; b < grid; b++) {
            const unsigned *o = h + b * 32;
            if (o[0] < 3) {
                printf("  sm %u cta %3d slot %u  warpid:", o[0], b, o[1]);
                for (int w = 0; w < NW; w++) printf(" %2u(sp%u)", o[w * 4 + 2], o[w * 4 + 2] & 3);
                printf("\n");
            }
            hist[o[1] % 4][o[2] & 3]++;
        }
        for (int s = 0; s < 4; s++) printf("  slot%%4=%d: warp 0 on sub-partition 0..3: %d %d %d %d\n", s, hist[s][0], hist[s][1], hist[s][2], hist[s][3]);
    }
    return 0;
}
