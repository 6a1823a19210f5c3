// Pre-projection of the encoder feature map through the first layer of the head (sd_field_project).
//
//   BTSNet.sample_features + ResnetFC.lin_in   models/bts.py:299-328, models/prediction_heads/resnetfc.py:162-163
//
// Layer 1 of the head is linear and so is the bilinear gather (F.grid_sample, bts.py:300): for a sample with taps
// w_t on texels t,   W_in[:, :C] . (sum_t w_t F[t])  =  sum_t w_t (W_in[:, :C] . F[t]).   The map is therefore pushed
// through W_in[:, :C] ONCE per encode -- P[t] = W_feat . F[t], [Hf*Wf, 128] fp16 channels-last -- and the per-sample
// path (field_bin.cu) interpolates the 128 hidden pre-activations instead of the 256 features: half the bytes per
// texel and 4/5 of the layer-1 contraction gone from the per-sample path.  The learn_empty replacement
// (bts.py:311-319) is linear too: W_feat . empty_feature rides in a spare column of the code block.
//
// Blob written by sd_field_project ("projected scene", offsets in launch.h: PROJ_*):
//   [0, 32768)      fp16 UMMA K-major SWIZZLE_128B image of the 128 x 128 identity (two K chunks): layer-1 "weights" of the
//                   projected features for the gather kernel (field_tc.cu), which blends 128 projected channels
//   [32768, 49152)  the same kind of image [128 hidden][64] of the code block of W_in: columns 0..38 positional code,
//                   39..44 coordinate hi/lo split, 45..46 bias (as in sd_mlp_pack), 47 = W_feat . empty_feature (zero
//                   unless learn_empty)
//   [49152, 50176)  W_feat . empty_feature as 128 halves (the gather kernel's replacement row)
//   [50176, ...)    P: [Hf*Wf][128] fp16
//
// The GEMM: persistent CTAs, 128 texels per tile; A = 128 texel rows x 256 channels of the channels-last fp16 map,
// brought in by TMA (four 64-channel boxes, SWIZZLE_128B = the UMMA K-major layout); B = the four feature chunks of
// the packed W_in image, resident in shared memory; D fp32 in TMEM, double buffered; the epilogue converts to fp16
// and each warp writes its 32 rows (8 KB contiguous) with one bulk shared->global copy.  HBM-bound:
// Hf*Wf*(512 + 256) B.
#include <cuda.h>

#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {

// ---- tensor maps (driver entry point through the runtime: the library does not link libcuda) -------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_f16(void *tmap_out, const void *base, int rank, const unsigned long long *dims,
                  const unsigned long long *strides_bytes, const unsigned int *box) {
    return make_tmap(tmap_out, base, 2, rank, dims, strides_bytes, box);
}

int make_tmap(void *tmap_out, const void *base, int elem_bytes, int rank, const unsigned long long *dims,
              const unsigned long long *strides_bytes, const unsigned int *box) {
    SD_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "make_tmap: fp16 or fp32 elements");
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SD_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        SD_REQUIRE(p && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    const CUresult r = fn(reinterpret_cast<CUtensorMap *>(tmap_out),
                          elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                          const_cast<void *>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return SD_OK;
}

namespace pj {
using namespace tcx;

constexpr int TM = 128;
constexpr int CHUNK = 16384;
constexpr int A_BYTES = 4 * CHUNK;            // C = 256 channels = 4 K chunks of 64
constexpr int OFF_W = 0;
constexpr int OFF_A = A_BYTES;
constexpr int OFF_STAGE = OFF_A + 2 * A_BYTES;
constexpr int OFF_BAR = OFF_STAGE + TM * 256;
enum { BAR_FULL = 0, BAR_EMPTY = 2, BAR_DFULL = 4, BAR_DEMPTY = 6, BAR_WLOAD = 8, NBAR = 9 };
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_ALLOC = OFF_TMEM + 16 + 1024;
constexpr int NTHREADS = 192;                 // warps 0-3 epilogue, 4 TMA, 5 MMA
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");

struct Params {
    CUtensorMap tmap;                         // [n_texels][256] fp16, box 64 x 128
    const unsigned char *w1_img;              // K-major SW128 image of W_in (sd_mlp_pack), chunks 0..3 = features
    __half *P;                                // [n_texels][128]
    long long n_texels, n_tiles;
};

__global__ void __launch_bounds__(NTHREADS, 1) featmap_project_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;   // (warp index the compiler knows to be uniform)
    const uint32_t bar0 = sm_u + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR(BAR_FULL + s), 1); mbar_init(BAR(BAR_EMPTY + s), 1);
            mbar_init(BAR(BAR_DFULL + s), 1); mbar_init(BAR(BAR_DEMPTY + s), 4);
        }
        mbar_init(BAR(BAR_WLOAD), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(BAR(BAR_WLOAD), A_BYTES);
        for (int c = 0; c < 4; ++c) bulk_g2s(sm_u + OFF_W + c * CHUNK, P.w1_img + (size_t)c * CHUNK, CHUNK, BAR(BAR_WLOAD));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);
    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_tiles = P.n_tiles > first ? (P.n_tiles - first + stride - 1) / stride : 0;

    if (warp < 4) {
        // ---- epilogue: D (fp32, TMEM) -> fp16 rows -> shared -> one 8 KB bulk store per warp ----------------
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        unsigned char *stage = sm + OFF_STAGE + warp * 8192;
        for (long long j = 0; j < my_tiles; ++j) {
            const long long tile = first + j * stride;
            const int b = (int)(j & 1);
            mbar_wait(BAR(BAR_DFULL + b), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (lane == 0) bulk_wait_read<0>();            // the previous tile's store has read the staging rows
            __syncwarp();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t vr[64];
                tmem_ld32_issue(t_lane + b * 128 + h * 64, vr);
                tmem_ld32_issue(t_lane + b * 128 + h * 64 + 32, vr + 32);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint4 o;
                    o.x = pack_h2(__uint_as_float(vr[8 * q + 0]), __uint_as_float(vr[8 * q + 1]));
                    o.y = pack_h2(__uint_as_float(vr[8 * q + 2]), __uint_as_float(vr[8 * q + 3]));
                    o.z = pack_h2(__uint_as_float(vr[8 * q + 4]), __uint_as_float(vr[8 * q + 5]));
                    o.w = pack_h2(__uint_as_float(vr[8 * q + 6]), __uint_as_float(vr[8 * q + 7]));
                    *reinterpret_cast<uint4 *>(stage + lane * 256 + h * 128 + q * 16) = o;
                }
            }
            tc_fence_before();
            mbar_arrive_warp(BAR(BAR_DEMPTY + b));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const long long r0 = tile * TM + warp * 32;
                const long long rows = P.n_texels - r0 < 32 ? P.n_texels - r0 : 32;
                if (rows > 0) bulk_s2g(P.P + r0 * 128, smem_u32(stage), (uint32_t)rows * 256u);
                bulk_commit();
            }
        }
        if (lane == 0) bulk_wait<0>();
    } else if (warp == 4) {
        // (whole warp in lockstep; copies inside an elect_one() branch, MMAs / commits in their elected forms: tc_common.cuh)
        if (lane == 0) tma_prefetch_desc(&P.tmap);
        __syncwarp();
        for (long long j = 0; j < my_tiles; ++j) {
            const long long tile = first + j * stride;
            const int s = (int)(j & 1);
            mbar_wait_warp(BAR(BAR_EMPTY + s), (uint32_t)(((j >> 1) & 1) ^ 1));
            if (elect_one()) {
                mbar_expect_tx(BAR(BAR_FULL + s), A_BYTES);
                for (int c = 0; c < 4; ++c)
                    tma_load_2d(sm_u + OFF_A + s * A_BYTES + c * CHUNK, &P.tmap, c * 64, (int)(tile * TM), BAR(BAR_FULL + s));
            }
            __syncwarp();
        }
    } else {
        {
            mbar_wait_warp(BAR(BAR_WLOAD), 0);
            const uint32_t idesc = umma_idesc(TM, 128);
            for (long long j = 0; j < my_tiles; ++j) {
                const int s = (int)(j & 1);
                mbar_wait_warp(BAR(BAR_FULL + s), (uint32_t)((j >> 1) & 1));
                mbar_wait_warp(BAR(BAR_DEMPTY + s), (uint32_t)(((j >> 1) & 1) ^ 1));
                tc_fence_after();
                if (elect_one()) {              // ONE elected thread issues: a single-thread region (tc_common.cuh)
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        umma(tmem_base + s * 128, umma_desc(sm_u + OFF_A + s * A_BYTES + (k >> 2) * CHUNK + (k & 3) * 32),
                             umma_desc(sm_u + OFF_W + (k >> 2) * CHUNK + (k & 3) * 32), idesc, k != 0);
                    umma_commit(BAR(BAR_EMPTY + s));
                    umma_commit(BAR(BAR_DFULL + s));
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// code block of W_in with the projected empty feature in column 47 (one block of 128 threads: thread = hidden unit)
__global__ void __launch_bounds__(128) proj_code_image_kernel(const unsigned char *__restrict__ w1_img, int C,
                                                              const float *__restrict__ empty_feature, int learn_empty,
                                                              unsigned char *__restrict__ wc_img) {
    const int n = threadIdx.x;
    const __half *w = reinterpret_cast<const __half *>(w1_img);
    float pe = 0.0f;
    if (learn_empty)
        for (int c = 0; c < C; ++c)   // the operands the tensor cores would see: half(W) * half(empty), fp32 accumulation
            pe = fmaf(__half2float(w[umma_sw128_offset(n, c, 128) / 2]), __half2float(__float2half_rn(__ldg(empty_feature + c))), pe);
    __half *o = reinterpret_cast<__half *>(wc_img + PROJ_OFF_CODE);
    for (int k = 0; k < 64; ++k) {
        __half v = __float2half_rn(0.0f);
        if (k < 47) v = w[umma_sw128_offset(n, C + k, 128) / 2];
        else if (k == 47) v = __float2half_rn(pe);
        o[umma_sw128_offset(n, k, 128) / 2] = v;
    }
    __half *id = reinterpret_cast<__half *>(wc_img + PROJ_OFF_IDENT);
    for (int k = 0; k < 128; ++k) id[umma_sw128_offset(n, k, 128) / 2] = __float2half_rn(k == n ? 1.0f : 0.0f);
    reinterpret_cast<__half *>(wc_img + PROJ_OFF_EMPTY)[n] = __float2half_rn(pe);
}


// ---- rel-1e-4 variant (sd_field_project_x3): P = W_in[:, :C] . F in fp32 on the CUDA cores, stored as an fp16 (hi, lo) pair
// of maps.  Once per encode; a 64-texel x 128-unit block tile, K in slices of 32 channels through shared memory, each
// thread 4 texels x 8 units.  (fp32 FFMA like the reference's own F.linear; the sum runs over the channels in order.)
constexpr int X3_TEX = 64, X3_KS = 32;
__global__ void __launch_bounds__(256) proj_x3_kernel(const float *__restrict__ F, const float *__restrict__ w_in_t, int C,
                                                      long long n_texels, __half *__restrict__ P_hi, __half *__restrict__ P_lo) {
    __shared__ float sF[X3_TEX][X3_KS + 1];
    __shared__ __align__(16) float sW[X3_KS][128];
    const long long t0 = (long long)blockIdx.x * X3_TEX;
    const int tid = threadIdx.x, ug = tid & 15, tg = tid >> 4;        // units 8 ug .. +7, texels 4 tg .. +3
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;
    for (int k0 = 0; k0 < C; k0 += X3_KS) {
        for (int i = tid; i < X3_TEX * X3_KS; i += 256) {
            const int tx = i / X3_KS, k = i - tx * X3_KS;
            sF[tx][k] = (t0 + tx < n_texels && k0 + k < C) ? __ldg(F + (size_t)(t0 + tx) * C + k0 + k) : 0.0f;
        }
        for (int i = tid; i < X3_KS * 128; i += 256) {
            const int k = i >> 7, n = i & 127;
            sW[k][n] = k0 + k < C ? __ldg(w_in_t + (size_t)(k0 + k) * 128 + n) : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < X3_KS; ++k) {
            const float4 w0 = *reinterpret_cast<const float4 *>(&sW[k][8 * ug]), w1 = *reinterpret_cast<const float4 *>(&sW[k][8 * ug + 4]);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float f = sF[4 * tg + a][k];
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(f, w[b], acc[a][b]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const long long t = t0 + 4 * tg + a;
        if (t >= n_texels) continue;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const __half2 h = __floats2half2_rn(acc[a][2 * b], acc[a][2 * b + 1]);
            const float2 hf = __half22float2(h);
            hi[b] = tcx::as_u32(h);
            lo[b] = tcx::pack_h2(acc[a][2 * b] - hf.x, acc[a][2 * b + 1] - hf.y);
        }
        *reinterpret_cast<uint4 *>(P_hi + (size_t)t * 128 + 8 * ug) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(P_lo + (size_t)t * 128 + 8 * ug) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// (hi, lo) UMMA images of the code block of W_in (column 39: W_feat . empty_feature) and of W_out (feature rows first, the
// density row last, as field_bin.cu expects them); one block of 128 threads
__global__ void __launch_bounds__(128) x3_weights_kernel(const float *__restrict__ w_in_t, const float *__restrict__ w_out_t, int C,
                                                         int d_out, int d_out_pad, int n2, const float *__restrict__ empty_feature,
                                                         int learn_empty, unsigned char *__restrict__ out) {
    const int n = threadIdx.x;
    auto put = [](unsigned char *img_hi, unsigned char *img_lo, size_t off, float v) {
        const __half h = __float2half_rn(v);
        reinterpret_cast<__half *>(img_hi)[off / 2] = h;
        reinterpret_cast<__half *>(img_lo)[off / 2] = __float2half_rn(v - __half2float(h));
    };
    float pe = 0.0f;
    if (learn_empty)
        for (int c = 0; c < C; ++c) pe = fmaf(__ldg(w_in_t + (size_t)c * 128 + n), __ldg(empty_feature + c), pe);
    unsigned char *wc_hi = out + PROJX_OFF_WC, *wc_lo = wc_hi + 16384;
    for (int k = 0; k < 64; ++k)
        put(wc_hi, wc_lo, umma_sw128_offset(n, k, 128), k < 39 ? __ldg(w_in_t + (size_t)(C + k) * 128 + n) : (k == 39 ? pe : 0.0f));
    // W_out: image row r = feature r (nn.Linear row 1 + r) for r < d_out - 1, the density (row 0) at r = d_out - 1; K = hidden unit
    unsigned char *w2_hi = out + PROJX_OFF_W2, *w2_lo = w2_hi + 2 * (size_t)n2 * 128;
    for (int r = 0; r < n2; ++r) {
        const int o = r < d_out - 1 ? r + 1 : (r == d_out - 1 ? 0 : -1);
        const float v = o >= 0 ? __ldg(w_out_t + (size_t)n * d_out_pad + o) : 0.0f;
        put(w2_hi, w2_lo, (size_t)(n >> 6) * n2 * 128 + umma_sw128_offset(r, n & 63, n2), v);
    }
}

}  // namespace pj
}  // namespace sd

using namespace sd;

extern "C" size_t sd_field_project_x3_bytes(const sd_scene *scene) {
    if (!scene || scene->Hf <= 0 || scene->Wf <= 0) return 0;
    return (size_t)PROJX_OFF_MAP + 2 * (size_t)scene->Hf * scene->Wf * 128 * sizeof(__half);
}

extern "C" int sd_field_project_x3(const sd_scene *scene, const sd_mlp *mlp, void *proj, size_t proj_bytes, void *stream) {
    SD_REQUIRE(scene && mlp && proj, "sd_field_project_x3: null pointer");
    SD_REQUIRE(scene->feat && scene->feat_dtype == SD_F32 && scene->C == 256 && scene->nv_f == 1,
               "sd_field_project_x3: needs the fp32 channels-last map with C = 256 and one encoder view");
    SD_REQUIRE(mlp->packed && mlp->d_hidden == 128 && mlp->d_in == scene->C + 39 && scene->include_input && scene->num_freqs == 6 &&
                   mlp->d_out >= 2 && mlp->d_out <= 65,
               "sd_field_project_x3: head must be packed, d_hidden = 128, d_in = C + 39, d_out <= 65 (got d_in=%d d_out=%d)", mlp->d_in, mlp->d_out);
    SD_REQUIRE(!scene->learn_empty || scene->empty_feature, "sd_field_project_x3: learn_empty without empty_feature");
    SD_REQUIRE(((uintptr_t)proj & 1023) == 0, "sd_field_project_x3: proj must be 1024-byte aligned");
    const size_t need = sd_field_project_x3_bytes(scene);
    if (proj_bytes < need) {
        set_error("sd_field_project_x3: %zu B needed, %zu B given", need, proj_bytes);
        return SD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    unsigned char *out = reinterpret_cast<unsigned char *>(proj);
    const float *w_in_t = reinterpret_cast<const float *>(blob + L.off_w_in_t);
    const float *w_out_t = reinterpret_cast<const float *>(blob + L.off_w_out_t);
    const int n2 = (mlp->d_out + 15) / 16 * 16;
    SD_CUDA_OK(cudaMemsetAsync(out, 0, PROJX_OFF_MAP, st));
    pj::x3_weights_kernel<<<1, 128, 0, st>>>(w_in_t, w_out_t, scene->C, mlp->d_out, L.d_out_pad, n2, scene->empty_feature,
                                             scene->learn_empty, out);
    SD_LAUNCH_OK("x3_weights_kernel");
    const long long n_texels = (long long)scene->Hf * scene->Wf;
    __half *P_hi = reinterpret_cast<__half *>(out + PROJX_OFF_MAP);
    pj::proj_x3_kernel<<<(unsigned)((n_texels + pj::X3_TEX - 1) / pj::X3_TEX), 256, 0, st>>>(
        reinterpret_cast<const float *>(scene->feat), w_in_t, scene->C, n_texels, P_hi, P_hi + (size_t)n_texels * 128);
    SD_LAUNCH_OK("proj_x3_kernel");
    return SD_OK;
}

extern "C" size_t sd_field_project_bytes(const sd_scene *scene) {
    if (!scene || scene->Hf <= 0 || scene->Wf <= 0) return 0;
    return (size_t)PROJ_OFF_MAP + (size_t)scene->Hf * scene->Wf * 128 * sizeof(__half);
}

extern "C" int sd_field_project(const sd_scene *scene, const sd_mlp *mlp, void *proj, size_t proj_bytes, void *stream) {
    SD_REQUIRE(scene && mlp && proj, "sd_field_project: null pointer");
    SD_REQUIRE(scene->feat && scene->feat_dtype == SD_F16 && scene->C == 256 && scene->nv_f == 1,
               "sd_field_project: needs the fp16 channels-last map with C = 256 and one encoder view");
    SD_REQUIRE(mlp->packed && mlp->d_hidden == 128 && mlp->d_in == scene->C + 39 && scene->include_input && scene->num_freqs == 6,
               "sd_field_project: head must be packed, d_hidden = 128, d_in = C + 39 (got d_in=%d)", mlp->d_in);
    SD_REQUIRE(!scene->learn_empty || scene->empty_feature, "sd_field_project: learn_empty without empty_feature");
    SD_REQUIRE(((uintptr_t)proj & 1023) == 0 && ((uintptr_t)scene->feat & 15) == 0, "sd_field_project: proj must be 1024-byte aligned");
    const size_t need = sd_field_project_bytes(scene);
    if (proj_bytes < need) {
        set_error("sd_field_project: %zu B needed, %zu B given", need, proj_bytes);
        return SD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    unsigned char *out = reinterpret_cast<unsigned char *>(proj);
    pj::proj_code_image_kernel<<<1, 128, 0, st>>>(blob + L.off_w_in_h, scene->C, scene->empty_feature, scene->learn_empty, out);
    SD_LAUNCH_OK("proj_code_image_kernel");

    pj::Params P = {};
    P.n_texels = (long long)scene->Hf * scene->Wf;
    P.n_tiles = (P.n_texels + pj::TM - 1) / pj::TM;
    P.w1_img = blob + L.off_w_in_h;
    P.P = reinterpret_cast<__half *>(out + PROJ_OFF_MAP);
    const unsigned long long dims[2] = {256ull, (unsigned long long)P.n_texels}, strides[1] = {512ull};
    const unsigned int box[2] = {64u, (unsigned)pj::TM};
    int rc = make_tmap_f16(&P.tmap, scene->feat, 2, dims, strides, box);
    if (rc) return rc;
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(pj::featmap_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pj::SMEM_ALLOC));
    }
    const unsigned grid = (unsigned)(P.n_tiles < sm_count ? P.n_tiles : sm_count);
    pj::featmap_project_kernel<<<grid, pj::NTHREADS, pj::SMEM_ALLOC, st>>>(P);
    SD_LAUNCH_OK("featmap_project_kernel");
    return SD_OK;
}
