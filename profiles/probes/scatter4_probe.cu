// Probe of the Blackwell TMA tile::scatter4 store (UTMASTG.2D.SCATTER4): which tensor-map box it wants, what it writes,
// and how many 4-row x 128-byte (or 256-byte) operations per microsecond one SM / the chip sustains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scatter4_probe scatter4_probe.cu && ./scatter4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <stdint.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_scatter(const __grid_constant__ CUtensorMap tm, const int *rows, int n_ops, int cols_per_op, int col0, int n_issuers) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float *f = reinterpret_cast<float *>(sm);
    for (int i = threadIdx.x; i < 4 * cols_per_op; i += blockDim.x) f[i] = 1000.0f * blockIdx.x + i;   // row r of the box: f[r*cols + c]
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < n_issuers) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        for (int o = threadIdx.x >> 5; o < n_ops; o += n_issuers) {
            const int *r = rows + ((size_t)blockIdx.x * n_ops + o) * 4;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                         ::"l"(&tm), "r"(col0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(s) : "memory");
            if ((o & 31) == 31) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const long long N = 2097152;
    const int C = 64;
    float *d;
    cudaMalloc(&d, N * C * 4);
    for (int variant = 0; variant < 2; ++variant) {
        const int box_rows = 1;                          // (a 4-row box raises "illegal instruction")
        const int cols = variant ? 64 : 32;              // 256-byte or 128-byte rows
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N}, strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d: box {%d,%d}: encode -> %d\n", variant, cols, box_rows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaMemset(d, 0, N * C * 4);
        // correctness: one block, one op, rows 5, 100, 7, 2000000
        int hrows[4] = {5, 100, 7, 2000000}, *drows;
        cudaMalloc(&drows, 16);
        cudaMemcpy(drows, hrows, 16, cudaMemcpyHostToDevice);
        k_scatter<<<1, 256, 4 * cols * 4 + 1024>>>(tm, drows, 1, cols, 0, 1);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  launch: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> h(C);
        for (int i = 0; i < 4; ++i) {
            cudaMemcpy(h.data(), d + (size_t)hrows[i] * C, C * 4, cudaMemcpyDeviceToHost);
            printf("  row %d: [0]=%g [1]=%g [%d]=%g [%d]=%g\n", hrows[i], h[0], h[1], cols - 1, h[cols - 1], C - 1, h[C - 1]);
        }
        cudaMemcpy(h.data(), d + (size_t)6 * C, C * 4, cudaMemcpyDeviceToHost);
        printf("  row 6 (untouched?): [0]=%g\n", h[0]);
        // throughput: 148 blocks x 4096 ops each to pseudo-random rows
        const int nb = 148, nops = 4096;
        std::vector<int> hr((size_t)nb * nops * 4);
        uint64_t s = 12345;
        for (auto &v : hr) { s = s * 6364136223846793005ull + 1442695040888963407ull; v = (int)((s >> 33) % N); }
        int *dr;
        cudaMalloc(&dr, hr.size() * 4);
        cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 4; ++it) {
            const int ni = 1 << it;
            printf("  issuers %d:", ni);
            cudaEventRecord(e0);
            k_scatter<<<nb, 256, 4 * cols * 4 + 1024>>>(tm, dr, nops, cols, 0, ni);
            cudaEventRecord(e1);
            e = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("  throughput: %s  %.3f ms for %d x %d ops -> %.2f ops/us/SM, %.1f GB/s chip\n", cudaGetErrorString(e), ms, nb, nops,
                   nops / (ms * 1e3), (double)nb * nops * 4 * cols * 4 / (ms * 1e-3) / 1e9);
        }
        cudaFree(dr); cudaFree(drows);
    }
    return 0;
}
