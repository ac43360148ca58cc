// Micro-test: which (box width, start column, row pitch) does a 2-D TMA load of 8-byte elements accept?
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_align tma_align.cu ; ./tma_align BOX_COLS COL PITCH [ROW]
// One combination per process (a faulting load poisons the context).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, int col, int row, int box_cols, int box_rows, double *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(box_cols * box_rows * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(b), "r"(col), "r"(row) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b) : "memory");
    const double *s = reinterpret_cast<const double *>(smem);
    for (int i = threadIdx.x; i < box_cols * box_rows; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char **argv)
{
    const int box_cols = atoi(argv[1]), col = atoi(argv[2]), pitch = atoi(argv[3]), row = argc > 4 ? atoi(argv[4]) : 0, rows = 64, box_rows = 20;
    std::vector<double> h((size_t)rows * pitch);
    for (int r = 0; r < rows; r++) for (int c = 0; c < pitch; c++) h[(size_t)r * pitch + c] = r * 1000.0 + c;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, box_cols * box_rows * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch * 8};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %d col %d pitch %d row %d: encode failed %d\n", box_cols, col, pitch, row, (int)r); return 0; }
    probe<<<1, 128, box_cols * box_rows * 8>>>(map, col, row, box_cols, box_rows, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %d col %d pitch %d row %d: %s\n", box_cols, col, pitch, row, cudaGetErrorString(e)); return 0; }
    std::vector<double> g(box_cols * box_rows);
    cudaMemcpy(g.data(), o, g.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < box_rows; rr++) for (int c = 0; c < box_cols; c++) {
        const double want = (col + c < pitch) ? (row + rr) * 1000.0 + col + c : 0.0;
        if (g[rr * box_cols + c] != want) bad++;
    }
    printf("box %d col %d pitch %d row %d: ok, %d mismatches\n", box_cols, col, pitch, row, bad);
    return 0;
}
