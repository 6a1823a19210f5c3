// Rays of whole views on the device: util.unproj_map (common/util.py:113-158), util.gen_rays
// (common/util.py:253-285) and the ray half of ImageRaySampler.sample (common/ray_sampler.py:439-486) as one kernel.
//
// A view is described by 100 bytes (pose, intrinsics, size); the [V*H*W][11] ray tensor the renderer reads is produced
// at store bandwidth instead of by a Python loop over batch elements with a dozen eager ops each.  Every fp32 operation
// is an explicitly rounded intrinsic in the order torch's kernels apply them (oracle/sd_oracle.c: sdo_gen_rays, pinned
// bit-for-bit to the reference by tests/golden/rays.npz): linspace = fma(step, i, start) below the midpoint and
// fma(-step, n - 1 - i, end) above it, |u|^2 = fma(uy, uy, ux * ux) + 1, rotation = three rounded products added left
// to right.
#include "common.cuh"

namespace sd {

constexpr int RAYS_PER_BLOCK = 256;
constexpr int RAY_DIM = 11;

struct RayGrid {
    float xs, xe, xstep, ys, ye, ystep, x_shift, y_shift, z_near, z_far;
    int V, H, W, norm_dir;
};

__device__ __forceinline__ float linspace_at(float start, float end, float step, int n, int i) {
    return i < n / 2 ? __fmaf_rn(step, (float)i, start) : __fmaf_rn(-step, (float)(n - 1 - i), end);
}

__global__ void __launch_bounds__(RAYS_PER_BLOCK) gen_rays_kernel(RayGrid g, const float *__restrict__ c2w,
                                                                  const float *__restrict__ proj,
                                                                  const float *__restrict__ frame_ids,
                                                                  float *__restrict__ rays) {
    __shared__ __align__(16) float s_ray[RAYS_PER_BLOCK * RAY_DIM];      // stride 11 words: conflict-free
    const long long R = (long long)g.V * g.H * g.W;
    const long long base = (long long)blockIdx.x * RAYS_PER_BLOCK;
    const long long r = base + threadIdx.x;
    if (r < R) {
        const int hw = g.H * g.W;
        const int v = (int)(r / hw);
        const int pix = (int)(r - (long long)v * hw);
        const int i = pix / g.W, j = pix - i * g.W;
        const float *P = c2w + 16 * v, *K = proj + 9 * v;
        const float fx = __ldg(K), fy = __ldg(K + 4), cx = __ldg(K + 2), cy = __ldg(K + 5);
        float x = linspace_at(g.xs, g.xe, g.xstep, g.W, j);
        float y = linspace_at(g.ys, g.ye, g.ystep, g.H, i);
        if (g.x_shift != 0.0f) x = __fadd_rn(x, g.x_shift);
        if (g.y_shift != 0.0f) y = __fadd_rn(y, g.y_shift);
        float ux = __fdiv_rn(__fsub_rn(x, cx), fx), uy = __fdiv_rn(__fsub_rn(y, cy), fy), uz = 1.0f;
        if (g.norm_dir) {
            float n2 = __fmul_rn(ux, ux);
            n2 = __fmaf_rn(uy, uy, n2);
            n2 = __fadd_rn(n2, 1.0f);
            const float nrm = __fsqrt_rn(n2);
            ux = __fdiv_rn(ux, nrm); uy = __fdiv_rn(uy, nrm); uz = __fdiv_rn(1.0f, nrm);
        }
        float *o = s_ray + threadIdx.x * RAY_DIM;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            o[a] = __ldg(P + 4 * a + 3);
            float acc = __fmul_rn(__ldg(P + 4 * a), ux);
            acc = __fadd_rn(acc, __fmul_rn(__ldg(P + 4 * a + 1), uy));
            acc = __fadd_rn(acc, __fmul_rn(__ldg(P + 4 * a + 2), uz));
            o[3 + a] = acc;
        }
        o[6] = g.z_near; o[7] = g.z_far;
        o[8] = frame_ids ? __ldg(frame_ids + v) : (float)v;
        o[9] = x; o[10] = y;
    }
    __syncthreads();
    // the block's rays are contiguous in the output: 16-byte coalesced stores (block bases are multiples of 11 KB)
    const long long left = R - base;
    const int nfl = (int)(left < RAYS_PER_BLOCK ? left : RAYS_PER_BLOCK) * RAY_DIM;
    float *dst = rays + base * RAY_DIM;
    const int nv4 = nfl >> 2;
    for (int k = threadIdx.x; k < nv4; k += RAYS_PER_BLOCK)
        reinterpret_cast<float4 *>(dst)[k] = reinterpret_cast<const float4 *>(s_ray)[k];
    for (int k = (nv4 << 2) + threadIdx.x; k < nfl; k += RAYS_PER_BLOCK) dst[k] = s_ray[k];
}

// ---- voxel centres of an SSC grid in the camera frame (sscbench/point_utils.py:46-67 generate_point_grid) ---------------
// centre = origin + size * idx + size * 0.5 per axis (TSDFVolume.vox2world, sscbench/fusion.py:205-219; evaluated in double,
// left to right, rounded to fp32 once -- see the kernel), then rigid_transform (fusion.py:407-411) with the
// calibration's float64 matrix: a double-precision dot product rounded once to fp32.  Flattened 'ij' order (x slowest).
struct VoxGrid {
    float ox, oy, oz;
    double vs;
    int ny, nz, x0;
    long long n;
    double T[12];
};

__global__ void __launch_bounds__(256) gen_voxel_grid_kernel(VoxGrid g, float *__restrict__ xyz) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= g.n) return;
    const int iz = (int)(i % g.nz);
    const long long t = i / g.nz;
    const int iy = (int)(t % g.ny), ix = (int)(t / g.ny) + g.x0;
    // TSDFVolume.vox2world (sscbench/fusion.py:205-219) is compiled by numba: fp32 origin and fp32 voxel index, but the voxel
    // size and the 0.5 offset are Python floats, so  origin + size * idx + size * 0.5  is evaluated in DOUBLE, left to right,
    // and rounded to fp32 once when it is stored
    const double half = __dmul_rn(g.vs, 0.5);
    const float px = __double2float_rn(__dadd_rn(__dadd_rn((double)g.ox, __dmul_rn(g.vs, (double)ix)), half));
    const float py = __double2float_rn(__dadd_rn(__dadd_rn((double)g.oy, __dmul_rn(g.vs, (double)iy)), half));
    const float pz = __double2float_rn(__dadd_rn(__dadd_rn((double)g.oz, __dmul_rn(g.vs, (double)iz)), half));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double v = fma(g.T[4 * r], (double)px, fma(g.T[4 * r + 1], (double)py, fma(g.T[4 * r + 2], (double)pz, g.T[4 * r + 3])));
        xyz[3 * i + r] = (float)v;
    }
}

}  // namespace sd

using namespace sd;

extern "C" int sd_gen_voxel_grid(const float *origin, double voxel_size, int nx, int ny, int nz, int x0, int x1,
                                 const double *T_host, float *xyz, void *stream) {
    SD_REQUIRE(origin && T_host, "sd_gen_voxel_grid: origin / T are host pointers and must not be NULL");
    SD_REQUIRE(nx > 0 && ny > 0 && nz > 0 && 0 <= x0 && x0 <= x1 && x1 <= nx, "sd_gen_voxel_grid: bad grid (%d x %d x %d, slab %d..%d)", nx, ny, nz, x0, x1);
    VoxGrid g;
    g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2]; g.vs = voxel_size;
    g.ny = ny; g.nz = nz; g.x0 = x0;
    g.n = (long long)(x1 - x0) * ny * nz;
    for (int i = 0; i < 12; ++i) g.T[i] = T_host[i];
    if (g.n == 0) return SD_OK;
    SD_REQUIRE(xyz, "sd_gen_voxel_grid: null output");
    gen_voxel_grid_kernel<<<(unsigned)((g.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, xyz);
    SD_LAUNCH_OK("gen_voxel_grid_kernel");
    return SD_OK;
}

extern "C" int sd_gen_rays(const float *c2w, const float *proj, const float *frame_ids, int V, int H, int W,
                           float z_near, float z_far, int norm_dir, float x_shift, float y_shift, float *rays,
                           void *stream) {
    SD_REQUIRE(V >= 0 && H >= 0 && W >= 0, "sd_gen_rays: negative size");
    const long long R = (long long)V * H * W;
    if (R == 0) return SD_OK;
    SD_REQUIRE(H >= 2 && W >= 2, "sd_gen_rays: an image needs at least 2 x 2 pixels (linspace step), got %d x %d", H, W);
    SD_REQUIRE(R / RAYS_PER_BLOCK < 0x7fffffffLL, "sd_gen_rays: too many rays");
    SD_REQUIRE(c2w && proj && rays, "sd_gen_rays: null pointer");
    SD_REQUIRE((reinterpret_cast<uintptr_t>(rays) & 15) == 0, "sd_gen_rays: rays must be 16-byte aligned");
    RayGrid g;
    // python doubles in the reference: -1 + .5 * (2 / W), rounded to fp32 when linspace takes them (util.py:147-150)
    g.xs = (float)(-1.0 + 0.5 * (2.0 / W)); g.xe = (float)(1.0 - 0.5 * (2.0 / W));
    g.ys = (float)(-1.0 + 0.5 * (2.0 / H)); g.ye = (float)(1.0 - 0.5 * (2.0 / H));
    g.xstep = (g.xe - g.xs) / (float)(W - 1);
    g.ystep = (g.ye - g.ys) / (float)(H - 1);
    g.x_shift = x_shift; g.y_shift = y_shift; g.z_near = z_near; g.z_far = z_far;
    g.V = V; g.H = H; g.W = W; g.norm_dir = norm_dir;
    const unsigned grid = (unsigned)((R + RAYS_PER_BLOCK - 1) / RAYS_PER_BLOCK);
    gen_rays_kernel<<<grid, RAYS_PER_BLOCK, 0, (cudaStream_t)stream>>>(g, c2w, proj, frame_ids, rays);
    SD_LAUNCH_OK("gen_rays_kernel");
    return SD_OK;
}
