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

// ---- backward of the composite (what autograd derives from nerf.py:376-421) ---------------------------------
// With s_k = 1 - a_k + 1e-10, T_k = prod_{i<k} s_i, w_k = a_k T_k and the outputs depth = sum w z, dino = sum w f,
// rgb = sum w c (+ 1 - sum w), weights = w, alphas = a:
//   G_k      = g_weights[k] + g_depth z_k + <g_dino, f_k> + <g_rgb, c_k> - white * sum(g_rgb)        (dL/dw_k)
//   dL/df_k  = w_k g_dino,   dL/dc_k = w_k g_rgb
//   dL/da_k  = G_k T_k + g_alphas[k] - (sum_{j>k} G_j w_j) / s_k          (torch.cumprod's backward divides the same way)
//   dL/dsig_k = dL/da_k |delta_k| exp(-|delta_k| sig_k)   for sigma_k > 0 and unless hard_alpha_cap pins the last alpha to 1
// One warp per ray, K <= 32 * NCH samples held in registers (lane = sample within a chunk).
__device__ __forceinline__ float warp_scan_add_down(float v, int lane) {   // inclusive suffix sum over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += n;
    }
    return v;
}

template <int NCH>
__global__ void __launch_bounds__(128) composite_bwd_kernel(const float *__restrict__ z, const float *__restrict__ sigma,
                                                            const float *__restrict__ feat, const float *__restrict__ rgb,
                                                            long long R, int K, int D, int Crgb, int hard_alpha_cap, int white_bkgd,
                                                            const float *__restrict__ g_depth, const float *__restrict__ g_dino,
                                                            const float *__restrict__ g_rgb_out, const float *__restrict__ g_weights,
                                                            const float *__restrict__ g_alphas, float *__restrict__ g_sigma,
                                                            float *__restrict__ g_feat, float *__restrict__ g_rgb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 4 + warp;
    if (r >= R) return;
    const float *zr = z + r * K, *sr = sigma + r * K;
    float a[NCH], Tk[NCH], w[NCH], G[NCH], dl[NCH];
    float T = 1.0f;
    const float gd = g_depth ? __ldg(g_depth + r) : 0.0f;
    float gsum_rgb = 0.0f;                                   // sum_c g_rgb[c] (white background term)
    const float grgb_l = (g_rgb_out && lane < Crgb) ? __ldg(g_rgb_out + r * Crgb + lane) : 0.0f;
    if (white_bkgd) gsum_rgb = warp_sum(grgb_l);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int k = c * 32 + lane;
        const bool act = k < K;
        const float zk = act ? __ldg(zr + k) : 0.0f;
        const float zn = (k + 1 < K) ? __ldg(zr + k + 1) : 0.0f;
        dl[c] = (k + 1 < K) ? fabsf(zn - zk) : 1e10f;
        a[c] = act ? sample_alpha(dl[c], __ldg(sr + k), k == K - 1, hard_alpha_cap) : 0.0f;
        const float shifted = act ? (1.0f - a[c]) + 1e-10f : 1.0f;
        const float incl = warp_scan_mul(shifted, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        Tk[c] = T * excl;
        w[c] = a[c] * Tk[c];
        T = T * __shfl_sync(0xffffffffu, incl, 31);
        G[c] = act ? ((g_weights ? __ldg(g_weights + r * K + k) : 0.0f) + gd * zk - gsum_rgb) : 0.0f;
    }
    // <g_dino, f_k>, <g_rgb, c_k> per sample (lanes stride the channels), and the gradients of f / c
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int n = min(32, K - c * 32);
        for (int j = 0; j < n; ++j) {
            const size_t row = (size_t)(r * K + c * 32 + j);
            const float wj = __shfl_sync(0xffffffffu, w[c], j);
            float part = 0.0f;
            if (g_dino && feat) {
                for (int ch = lane; ch < D; ch += 32) {
                    const float g = __ldg(g_dino + r * D + ch);
                    part = fmaf(g, __ldg(feat + row * D + ch), part);
                    if (g_feat) g_feat[row * D + ch] = wj * g;
                }
            } else if (g_feat) {
                for (int ch = lane; ch < D; ch += 32) g_feat[row * D + ch] = 0.0f;
            }
            if (lane < Crgb) {
                if (g_rgb_out && rgb) part = fmaf(grgb_l, __ldg(rgb + row * Crgb + lane), part);
                if (g_rgb) g_rgb[row * Crgb + lane] = wj * grgb_l;
            }
            part = warp_sum(part);
            if (lane == j) G[c] += part;
        }
    }
    // suffix sums of G w, last chunk first
    float carry = 0.0f;                                       // sum over the chunks behind the current one
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) {
        const int k = c * 32 + lane;
        const bool act = k < K;
        const float gw = act ? G[c] * w[c] : 0.0f;
        const float incl = warp_scan_add_down(gw, lane);
        const float S = incl - gw + carry;                    // sum_{j>k} G_j w_j
        carry += __shfl_sync(0xffffffffu, incl, 0);
        if (act && g_sigma) {
            const float s_k = (1.0f - a[c]) + 1e-10f;
            float da = G[c] * Tk[c] + (g_alphas ? __ldg(g_alphas + r * K + k) : 0.0f) - S / s_k;
            const float sg = __ldg(sr + k);
            const bool pinned = hard_alpha_cap && k == K - 1;
            g_sigma[r * K + k] = (sg > 0.0f && !pinned) ? da * dl[c] * expf(-dl[c] * sg) : 0.0f;   // d/dsig (1 - exp(-|delta| sig))
        }
    }
}

}  // namespace sd

extern "C" int sd_composite_bwd(const float *z, const float *sigma, const float *feat, const float *rgb, long long R, int K, int D,
                                int Crgb, const sd_render_cfg *cfg, const float *g_depth, const float *g_dino,
                                const float *g_rgb_out, const float *g_weights, const float *g_alphas, float *g_sigma,
                                float *g_feat, float *g_rgb, void *stream) {
    SD_REQUIRE(cfg, "sd_composite_bwd: null pointer");
    SD_REQUIRE(R >= 0 && K > 0 && K <= 256, "sd_composite_bwd: 1 <= K <= 256 samples per ray (got %d)", K);
    if (R == 0) return SD_OK;
    SD_REQUIRE(z && sigma, "sd_composite_bwd: null pointer");
    SD_REQUIRE(D >= 0 && D <= 1024 && Crgb >= 0 && Crgb <= 32, "sd_composite_bwd: D <= 1024, Crgb <= 32");
    SD_REQUIRE(!(g_dino && D > 0) || feat, "sd_composite_bwd: g_dino without feat");
    SD_REQUIRE(!(g_rgb_out && Crgb > 0) || rgb, "sd_composite_bwd: g_rgb_out without rgb");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((R + 3) / 4);
#define SD_COMPOSITE_BWD(NCH)                                                                                                  \
    sd::composite_bwd_kernel<NCH><<<grid, 128, 0, st>>>(z, sigma, feat, rgb, R, K, D, Crgb, cfg->hard_alpha_cap, cfg->white_bkgd, \
                                                        g_depth, g_dino, g_rgb_out, g_weights, g_alphas, g_sigma, g_feat, g_rgb)
    if (K <= 32) SD_COMPOSITE_BWD(1);
    else if (K <= 64) SD_COMPOSITE_BWD(2);
    else if (K <= 128) SD_COMPOSITE_BWD(4);
    else SD_COMPOSITE_BWD(8);
#undef SD_COMPOSITE_BWD
    SD_LAUNCH_OK("composite_bwd_kernel");
    return SD_OK;
}

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
