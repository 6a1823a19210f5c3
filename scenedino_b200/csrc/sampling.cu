// NeRFRenderer sampling kernels (renderer/nerf.py:121-228, 490, 522).
//
// All results here are compared BIT-FOR-BIT with oracle/sd_oracle.c, so every fp32 operation is an
// explicitly rounded intrinsic in the reference's eager-op order and the CDF accumulates
// sequentially in double (torch CPU cumsum semantics), exactly like the oracle.
#include "common.cuh"
#include "launch.h"

namespace sd {

// ---- a-1: sample_coarse (nerf.py:121-141) -------------------------------------------------------
__global__ void __launch_bounds__(256) sample_coarse_kernel(const float *__restrict__ rays, long long R, int r_dim,
                                                            const float *__restrict__ u,
                                                            const float *__restrict__ lin, int Kc, int lindisp,
                                                            float step, float *__restrict__ z) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * Kc) return;
    const long long r = i / Kc;
    const int k = (int)(i - r * Kc);
    const float near = __ldg(rays + r * r_dim + 6), far = __ldg(rays + r * r_dim + 7);
    const float t = __fadd_rn(__ldg(lin + k), __fmul_rn(__ldg(u + i), step));
    z[i] = depth_from_t(near, far, t, lindisp);
}

// CDF of (w + 1e-5) / sum, built by one lane: sequential double accumulation, fp32 per element.
__device__ __forceinline__ void build_cdf(const float *__restrict__ w, int K, float *cdf) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += (double)__fadd_rn(w[k], 1e-5f);
    const float sf = __double2float_rn(s);
    double acc = 0.0;
    cdf[0] = 0.0f;
    for (int k = 0; k < K; ++k) {
        const float pdf = __fdiv_rn(__fadd_rn(w[k], 1e-5f), sf);
        acc += (double)pdf;
        cdf[k + 1] = __double2float_rn(acc);
    }
}

// torch.searchsorted(cdf, u, right=True): number of entries <= u
__device__ __forceinline__ int upper_bound(const float *cdf, int n, float u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- a-2: sample_fine (nerf.py:181-212); one warp per ray ----------------------------------------
__global__ void __launch_bounds__(128) sample_fine_kernel(const float *__restrict__ rays, long long R, int r_dim,
                                                          const float *__restrict__ weights, int Kc,
                                                          const float *__restrict__ u0,
                                                          const float *__restrict__ u1, int Kf, int lindisp,
                                                          float *__restrict__ z, int *__restrict__ inds) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= R) return;
    float *cdf = smem + warp * (2 * Kc + 2);
    float *w = cdf + Kc + 1;
    for (int k = lane; k < Kc; k += 32) w[k] = __ldg(weights + r * Kc + k);
    __syncwarp();
    if (lane == 0) build_cdf(w, Kc, cdf);
    __syncwarp();
    const float near = __ldg(rays + r * r_dim + 6), far = __ldg(rays + r * r_dim + 7);
    for (int j = lane; j < Kf; j += 32) {
        int ind = upper_bound(cdf, Kc + 1, __ldg(u0 + r * Kf + j)) - 1;
        ind = ind < 0 ? 0 : ind;
        if (inds) inds[r * Kf + j] = ind;
        const float t = __fdiv_rn(__fadd_rn((float)ind, __ldg(u1 + r * Kf + j)), (float)Kc);
        z[r * Kf + j] = depth_from_t(near, far, t, lindisp);
    }
}

// ---- a-3: sample_fine_depth (nerf.py:214-228) -----------------------------------------------------
__global__ void __launch_bounds__(256) sample_fine_depth_kernel(const float *__restrict__ rays, long long R,
                                                                int r_dim, const float *__restrict__ depth,
                                                                const float *__restrict__ noise, int Kfd,
                                                                float depth_std, float *__restrict__ z) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * Kfd) return;
    const long long r = i / Kfd;
    const float near = __ldg(rays + r * r_dim + 6), far = __ldg(rays + r * r_dim + 7);
    float v = __fadd_rn(__ldg(depth + r), __fmul_rn(__ldg(noise + i), depth_std));
    v = v < far ? v : far;
    v = v > near ? v : near;
    z[i] = v;
}

// ---- a-4: sample_coarse_from_dist (nerf.py:143-179); one warp per ray ----------------------------
__global__ void __launch_bounds__(128) sample_from_dist_kernel(long long R, const float *__restrict__ weights,
                                                               const float *__restrict__ z_samp, int Kp,
                                                               const float *__restrict__ u0,
                                                               const float *__restrict__ u1, int Kc, int lindisp,
                                                               float *__restrict__ z, int *__restrict__ inds) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= R) return;
    float *cdf = smem + warp * (4 * Kp + 4);
    float *w = cdf + Kp + 1;
    float *zz = w + Kp;
    float *bord = zz + Kp;  // Kp + 1
    for (int k = lane; k < Kp; k += 32) {
        w[k] = __ldg(weights + r * Kp + k);
        const float zv = __ldg(z_samp + r * Kp + k);
        zz[k] = lindisp ? __fdiv_rn(1.0f, zv) : zv;
    }
    __syncwarp();
    if (lane == 0) build_cdf(w, Kp, cdf);
    for (int k = lane; k <= Kp; k += 32) {
        float b;
        if (k == 0) b = zz[0];
        else if (k == Kp) b = zz[Kp - 1];
        else b = __fmul_rn(0.5f, __fadd_rn(zz[k], zz[k - 1]));
        bord[k] = b;
    }
    __syncwarp();
    for (int j = lane; j < Kc; j += 32) {
        int id = upper_bound(cdf, Kp + 1, __ldg(u0 + r * Kc + j)) - 1;
        id = id < 0 ? 0 : id;
        id = id > Kc - 1 ? Kc - 1 : id;  // clamp(0, num_samples - 1), nerf.py:157
        id = id > Kp - 1 ? Kp - 1 : id;  // memory safety only
        if (inds) inds[r * Kc + j] = id;
        const float t = __ldg(u1 + r * Kc + j);
        const float v = __fadd_rn(__fmul_rn(bord[id], __fsub_rn(1.0f, t)), __fmul_rn(bord[id + 1], t));
        z[r * Kc + j] = lindisp ? __fdiv_rn(1.0f, v) : v;
    }
}

// ---- a-5: row sort (nerf.py:490,522); one warp per row, bitonic network in shared memory ----------
__global__ void __launch_bounds__(128) sort_rows_kernel(float *__restrict__ z, long long R, int K, int P) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= R) return;
    float *v = smem + warp * P;
    for (int k = lane; k < P; k += 32) v[k] = k < K ? z[r * K + k] : __int_as_float(0x7f800000);
    __syncwarp();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < P / 2; i += 32) {
                const int lo = 2 * i - (i & (stride - 1));  // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const float a = v[lo], b = v[hi];
                if ((a > b) == up) { v[lo] = b; v[hi] = a; }
            }
            __syncwarp();
        }
    }
    for (int k = lane; k < K; k += 32) z[r * K + k] = v[k];
}

// ---- a-2 + a-3 + a-5 in one launch: the fine pass's depths (nerf.py:511-529) --------------------------------------
// One warp per ray: importance samples from the coarse weights (sample_fine), samples around the expected depth
// (sample_fine_depth), concatenation with the coarse depths and the ascending sort -- the same arithmetic as the four
// separate entry points, bit for bit, without the three intermediate tensors and four launches.
__global__ void __launch_bounds__(128) fine_merge_kernel(const float *__restrict__ rays, long long R, int r_dim,
                                                         const float *__restrict__ weights, const float *__restrict__ z_coarse,
                                                         const float *__restrict__ depth, int Kc, const float *__restrict__ u0,
                                                         const float *__restrict__ u1, int Kfi, const float *__restrict__ noise,
                                                         int Kfd, float depth_std, int lindisp, int P, float *__restrict__ z_all) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= R) return;
    const int K = Kc + Kfi + Kfd;
    float *cdf = smem + warp * (2 * Kc + 2 + P);
    float *w = cdf + Kc + 1;
    float *v = w + Kc + 1;
    for (int k = lane; k < Kc; k += 32) {
        w[k] = __ldg(weights + r * Kc + k);
        v[k] = __ldg(z_coarse + r * Kc + k);
    }
    for (int k = K + lane; k < P; k += 32) v[k] = __int_as_float(0x7f800000);
    __syncwarp();
    const float near = __ldg(rays + r * r_dim + 6), far = __ldg(rays + r * r_dim + 7);
    if (Kfi > 0) {
        if (lane == 0) build_cdf(w, Kc, cdf);
        __syncwarp();
        for (int j = lane; j < Kfi; j += 32) {
            int ind = upper_bound(cdf, Kc + 1, __ldg(u0 + r * Kfi + j)) - 1;
            ind = ind < 0 ? 0 : ind;
            const float t = __fdiv_rn(__fadd_rn((float)ind, __ldg(u1 + r * Kfi + j)), (float)Kc);
            v[Kc + j] = depth_from_t(near, far, t, lindisp);
        }
    }
    if (Kfd > 0) {
        const float d = __ldg(depth + r);
        for (int j = lane; j < Kfd; j += 32) {
            float x = __fadd_rn(d, __fmul_rn(__ldg(noise + r * Kfd + j), depth_std));
            x = x < far ? x : far;
            x = x > near ? x : near;
            v[Kc + Kfi + j] = x;
        }
    }
    __syncwarp();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < P / 2; i += 32) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const float a = v[lo], b2 = v[hi];
                if ((a > b2) == up) { v[lo] = b2; v[hi] = a; }
            }
            __syncwarp();
        }
    }
    for (int k = lane; k < K; k += 32) z_all[r * K + k] = v[k];
}

int launch_fine_merge(const float *rays, long long R, int r_dim, const float *weights, const float *z_coarse, const float *depth,
                      int Kc, const float *u0, const float *u1, int Kfi, const float *noise, int Kfd, float depth_std,
                      int lindisp, float *z_all, cudaStream_t st) {
    const int K = Kc + Kfi + Kfd;
    SD_REQUIRE(R >= 0 && r_dim >= 8 && Kc > 0 && Kfi >= 0 && Kfd >= 0 && Kfi + Kfd > 0, "fine pass: bad sample counts");
    if (R == 0) return SD_OK;
    SD_REQUIRE(rays && weights && z_coarse && z_all && (Kfi == 0 || (u0 && u1)) && (Kfd == 0 || (depth && noise)), "fine pass: null pointer");
    int P = 2;
    while (P < K) P <<= 1;
    const size_t smem = 4 * (size_t)(2 * Kc + 2 + P) * sizeof(float);
    SD_REQUIRE(smem <= 48 * 1024, "fine pass: %d + %d samples per ray do not fit the merge kernel's shared memory", Kc, Kfi + Kfd);
    fine_merge_kernel<<<(unsigned)((R + 3) / 4), 128, smem, st>>>(rays, R, r_dim, weights, z_coarse, depth, Kc, u0, u1, Kfi, noise,
                                                                  Kfd, depth_std, lindisp, P, z_all);
    SD_LAUNCH_OK("fine_merge_kernel");
    return SD_OK;
}

}  // namespace sd

using namespace sd;

extern "C" int sd_sample_coarse(const float *rays, long long R, int r_dim, const float *u, const float *lin,
                                int Kc, int lindisp, float *z, void *stream) {
    SD_REQUIRE(R >= 0 && r_dim >= 8 && Kc > 0, "sd_sample_coarse: bad shape");
    if (R == 0) return SD_OK;
    SD_REQUIRE(rays && u && lin && z, "sd_sample_coarse: null pointer");
    const long long n = R * Kc;
    const float step = (float)(1.0 / (double)Kc);
    sample_coarse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays, R, r_dim, u, lin, Kc,
                                                                                         lindisp, step, z);
    SD_LAUNCH_OK("sample_coarse_kernel");
    return SD_OK;
}

extern "C" int sd_sample_fine(const float *rays, long long R, int r_dim, const float *weights, int Kc,
                              const float *u0, const float *u1, int Kf, int lindisp, float *z, int *inds,
                              void *stream) {
    // 4 rays per block, (2 Kc + 2) floats each, inside the 48 KB a kernel gets without opting in
    SD_REQUIRE(R >= 0 && r_dim >= 8 && Kc > 0 && Kf > 0, "sd_sample_fine: bad shape");
    SD_REQUIRE(Kc <= 1534, "sd_sample_fine: at most 1534 coarse samples per ray (got %d)", Kc);
    if (R == 0) return SD_OK;
    SD_REQUIRE(rays && weights && u0 && u1 && z, "sd_sample_fine: null pointer");
    const size_t smem = 4 * (2 * (size_t)Kc + 2) * sizeof(float);
    sample_fine_kernel<<<(unsigned)((R + 3) / 4), 128, smem, (cudaStream_t)stream>>>(rays, R, r_dim, weights, Kc, u0,
                                                                                      u1, Kf, lindisp, z, inds);
    SD_LAUNCH_OK("sample_fine_kernel");
    return SD_OK;
}

extern "C" int sd_sample_fine_depth(const float *rays, long long R, int r_dim, const float *depth,
                                    const float *noise, int Kfd, float depth_std, float *z, void *stream) {
    SD_REQUIRE(R >= 0 && r_dim >= 8 && Kfd > 0, "sd_sample_fine_depth: bad shape");
    if (R == 0) return SD_OK;
    SD_REQUIRE(rays && depth && noise && z, "sd_sample_fine_depth: null pointer");
    const long long n = R * Kfd;
    sample_fine_depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays, R, r_dim, depth,
                                                                                             noise, Kfd, depth_std, z);
    SD_LAUNCH_OK("sample_fine_depth_kernel");
    return SD_OK;
}

extern "C" int sd_sample_coarse_from_dist(long long R, const float *weights, const float *z_samp, int Kp,
                                          const float *u0, const float *u1, int Kc, int lindisp, float *z,
                                          int *inds, void *stream) {
    // 4 rays per block, (4 Kp + 4) floats each, inside the 48 KB a kernel gets without opting in
    SD_REQUIRE(R >= 0 && Kp > 0 && Kc > 0, "sd_sample_coarse_from_dist: bad shape");
    SD_REQUIRE(Kp <= 767, "sd_sample_coarse_from_dist: at most 767 proposal samples per ray (got %d)", Kp);
    if (R == 0) return SD_OK;
    SD_REQUIRE(weights && z_samp && u0 && u1 && z, "sd_sample_coarse_from_dist: null pointer");
    const size_t smem = 4 * (4 * (size_t)Kp + 4) * sizeof(float);
    sample_from_dist_kernel<<<(unsigned)((R + 3) / 4), 128, smem, (cudaStream_t)stream>>>(R, weights, z_samp, Kp, u0,
                                                                                           u1, Kc, lindisp, z, inds);
    SD_LAUNCH_OK("sample_from_dist_kernel");
    return SD_OK;
}

extern "C" int sd_sort_rows(float *z, long long R, int K, void *stream) {
    SD_REQUIRE(R >= 0 && K > 0 && K <= 1024, "sd_sort_rows: K must be in [1,1024]");
    if (R == 0 || K == 1) return SD_OK;
    SD_REQUIRE(z, "sd_sort_rows: null pointer");
    int P = 2;
    while (P < K) P <<= 1;
    sort_rows_kernel<<<(unsigned)((R + 3) / 4), 128, 4 * (size_t)P * sizeof(float), (cudaStream_t)stream>>>(z, R, K, P);
    SD_LAUNCH_OK("sort_rows_kernel");
    return SD_OK;
}
