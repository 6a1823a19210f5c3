// NeRFRenderer.composite core (renderer/nerf.py:246-249, 376-405, 418-421), unfused form:
// reads per-sample sigma / features / colours from HBM and writes per-ray results.
//
// One warp per ray.  Samples are walked in chunks of 32 (one lane per sample): alpha per lane, the
// exclusive transmittance product as a warp-level multiplicative scan (__shfl_up_sync) carried
// across chunks, then the weighted sums with lanes striding the channel dimension so every
// feature row is read as one coalesced segment.
#include "common.cuh"

namespace sd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inclusive multiplicative scan over the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= n;
    }
    return v;
}

// alpha of one sample (nerf.py:376-381)
__device__ __forceinline__ float sample_alpha(float delta, float sigma, bool last, int hard_alpha_cap) {
    const float sg = sigma > 0.0f ? sigma : (sigma != sigma ? sigma : 0.0f);
    float a = 1.0f - expf(-fabsf(delta) * sg);
    if (hard_alpha_cap && last) a = 1.0f;
    return a;
}

template <int DMAX_PER_LANE>
__global__ void __launch_bounds__(128) composite_kernel(const float *__restrict__ z, const float *__restrict__ sigma,
                                                        const float *__restrict__ feat, const float *__restrict__ rgb,
                                                        long long R, int K, int D, int Crgb, int hard_alpha_cap,
                                                        int white_bkgd, float *__restrict__ weights,
                                                        float *__restrict__ alphas, float *__restrict__ depth,
                                                        float *__restrict__ dino, float *__restrict__ rgb_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= R) return;
    const float *zr = z + r * K, *sr = sigma + r * K;
    float T = 1.0f;           // transmittance in front of the current chunk
    float dsum = 0.0f, wsum = 0.0f;
    float facc[DMAX_PER_LANE];
#pragma unroll
    for (int i = 0; i < DMAX_PER_LANE; ++i) facc[i] = 0.0f;
    float cacc = 0.0f;        // lane c < Crgb accumulates colour channel c (Crgb <= 32)
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const bool act = k < K;
        const float zk = act ? __ldg(zr + k) : 0.0f;
        const float zn = (k + 1 < K) ? __ldg(zr + k + 1) : 0.0f;
        const float delta = (k + 1 < K) ? zn - zk : 1e10f;
        const float a = act ? sample_alpha(delta, __ldg(sr + k), k == K - 1, hard_alpha_cap) : 0.0f;
        const float shifted = act ? (1.0f - a) + 1e-10f : 1.0f;
        const float incl = warp_scan_mul(shifted, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = a * (T * excl);
        T = T * __shfl_sync(0xffffffffu, incl, 31);
        if (act) {
            if (weights) weights[r * K + k] = w;
            if (alphas) alphas[r * K + k] = a;
        }
        dsum += act ? w * zk : 0.0f;
        wsum += act ? w : 0.0f;
        const int n = min(32, K - k0);
        for (int j = 0; j < n; ++j) {
            const float wj = __shfl_sync(0xffffffffu, w, j);
            const size_t row = (size_t)(r * K + k0 + j);
            if (feat) {
                const float *f = feat + row * D;
#pragma unroll
                for (int i = 0; i < DMAX_PER_LANE; ++i) {
                    const int c = lane + 32 * i;
                    if (c < D) facc[i] = fmaf(__ldg(f + c), wj, facc[i]);
                }
            }
            if (rgb && lane < Crgb) cacc = fmaf(wj, __ldg(rgb + row * Crgb + lane), cacc);
        }
    }
    dsum = warp_sum(dsum);
    wsum = warp_sum(wsum);
    if (lane == 0 && depth) depth[r] = dsum;
    if (dino) {
#pragma unroll
        for (int i = 0; i < DMAX_PER_LANE; ++i) {
            const int c = lane + 32 * i;
            if (c < D) dino[r * D + c] = facc[i];
        }
    }
    if (rgb_out && lane < Crgb) rgb_out[r * Crgb + lane] = white_bkgd ? cacc + 1.0f - wsum : cacc;
}

}  // namespace sd

extern "C" int sd_composite(const float *z, const float *sigma, const float *feat, const float *rgb,
                            long long R, int K, int D, int Crgb, const sd_render_cfg *cfg, float *weights,
                            float *alphas, float *depth, float *dino, float *rgb_out, void *stream) {
    SD_REQUIRE(cfg, "sd_composite: null pointer");
    SD_REQUIRE(R >= 0 && K > 0, "sd_composite: bad shape");
    if (R == 0) return SD_OK;
    SD_REQUIRE(z && sigma, "sd_composite: null pointer");
    SD_REQUIRE(D >= 0 && D <= 1024, "sd_composite: D must be <= 1024 (got %d)", D);
    SD_REQUIRE(Crgb >= 0 && Crgb <= 32, "sd_composite: at most 10 colour views (Crgb=%d)", Crgb);
    SD_REQUIRE(!(dino && D > 0) || feat, "sd_composite: dino requested without feat");
    SD_REQUIRE(!(rgb_out && Crgb > 0) || rgb, "sd_composite: rgb_out requested without rgb");
    if (R == 0) return SD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((R + 3) / 4);
    const float *f = (D > 0 && dino) ? feat : nullptr;
    const float *c = (Crgb > 0 && rgb_out) ? rgb : nullptr;
#define SD_COMPOSITE(NPL)                                                                                     \
    sd::composite_kernel<NPL><<<grid, 128, 0, st>>>(z, sigma, f, c, R, K, D, Crgb, cfg->hard_alpha_cap,     \
                                                    cfg->white_bkgd, weights, alphas, depth, dino, rgb_out)
    if (D <= 64) SD_COMPOSITE(2);
    else if (D <= 256) SD_COMPOSITE(8);
    else SD_COMPOSITE(32);
#undef SD_COMPOSITE
    SD_LAUNCH_OK("composite_kernel");
    return SD_OK;
}
