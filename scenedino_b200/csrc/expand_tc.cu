// MlpDimReduction.transform_expand (backbones/dino/dim_reduction.py:22-25) on the tensor cores:
//   out[N, 768] = normalize(W2 relu(W1 f + b1) + b2),  f [N, 64] fp32, F.normalize eps 1e-12.
//
// It follows every voxel query of the SSC path (models/bts.py:585) and every rendered feature image
// (demo_utils/utils.py:229); on CUDA cores (launch_mlp_simt) the 2 097 152 voxels of the SSC grid took 126 ms, 430 x the
// field query that produces their 64-d input.  Here it is bound by the 3 KB per row it has to write (6.4 GB per grid).
//
// Persistent, one CTA per SM, 192 threads, 128-row tiles:
//   warps 0-3  read the tile's fp32 rows, write the fp16 A operand (K-major SWIZZLE_128B), then run both epilogues:
//              (1) D1 + b1 -> ReLU -> fp16 -> the same TMEM columns (A operand of layer 2, TS-form MMA);
//              (2) the 768 outputs come 128 columns at a time (N = 128 MMAs into two alternating accumulators).  A
//              thread owns one row, so the row's sum of squares is a register: the chunks are computed TWICE per
//              tile -- pass 0 accumulates |v|^2, pass 1 scales and stores -- instead of parking 3 KB per row
//              somewhere (the tensor pipe has the time: 0.2 MFLOP per row against 3 KB of HBM writes).  Stores go
//              through a padded per-warp staging buffer so that every STG.128 of a warp writes 512 contiguous bytes;
//   warp 4     streams the 32 KB operand image of each 128-column block of W2 (196 KB in all: L2-resident) into a
//              two-slot ring by bulk copies;
//   warp 5     issues the MMAs (one thread): 4 x (128x128x16) for layer 1, 8 per chunk step for layer 2.
// Tiles are processed one after the other (no overlap between the epilogue of tile t and layer 1 of tile t+1): the
// layer-2 chunk steps overlap with epilogue 2 through the two accumulators, which is where the time goes.
#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {
namespace ex {
using namespace tcx;

constexpr int TM = 128;
constexpr int NW2 = 2;                        // W2 ring slots
constexpr int W2_BYTES = 32768;               // one 128-output block: [2 K blocks][128 rows][128 B]
constexpr int OFF_W1 = 0;                     // [128 hidden][64 k] fp16, 16 KB
constexpr int OFF_A = 16384;                  // [128 rows][64 k] fp16, 16 KB
constexpr int OFF_W2 = 32768;
constexpr int STAGE_ROW = 528;                // 512 B of a row chunk + 16 B: rows of a warp land in different banks
constexpr int STAGE_WARP = 32 * STAGE_ROW;
constexpr int OFF_STAGE = OFF_W2 + NW2 * W2_BYTES;
constexpr int OFF_B1 = OFF_STAGE + 4 * STAGE_WARP;
constexpr int OFF_B2 = OFF_B1 + 512;
constexpr int MAX_DOUT = 1024;
constexpr int OFF_BAR = OFF_B2 + MAX_DOUT * 4;
enum { BAR_WLOAD = 0, BAR_A_FULL, BAR_D1_FULL, BAR_H_FULL, BAR_W_FULL, BAR_W_EMPTY = BAR_W_FULL + NW2,
       BAR_D2_FULL = BAR_W_EMPTY + NW2, BAR_D2_FREE = BAR_D2_FULL + 2, NBAR = BAR_D2_FREE + 2 };
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_ALLOC = OFF_TMEM + 16 + 1024;
constexpr int NTHREADS = 192;
constexpr int TMEM_COLS = 512;                // D1 / H at 0..127, the two layer-2 accumulators at 128 and 256
constexpr int D2_COL = 128;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(STAGE_ROW % 16 == 0, "staging rows hold 16-byte accesses");

struct Params {
    int head2;                                // 1: second layer only, on explicit hidden rows (launch_head2)
    const float *h;                           // head2: [N][128] fp32 hidden rows (per-ray sums of the hidden-composite render)
    const float *scale;                       // head2: [N] factor of the bias (sum of the compositing weights) or NULL = 1
    const float *f;                           // [N][64]
    float *out;                               // [N][d_out]
    const unsigned char *w1_img;              // K-major SW128 image of W1 (first K block of the blob's W_in image)
    const unsigned char *w2_img;              // [d_out / 128][W2_BYTES]
    const float *b1, *b2;
    long long N, n_tiles;
    int d_out, nch;
};

__global__ void __launch_bounds__(NTHREADS, 1) expand_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;   // (warp index the compiler knows to be uniform)
    const uint32_t bar0 = sm_u + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    if (tid == 0) {
        mbar_init(BAR(BAR_WLOAD), 1);
        mbar_init(BAR(BAR_A_FULL), 4);
        mbar_init(BAR(BAR_D1_FULL), 1);
        mbar_init(BAR(BAR_H_FULL), 4);
        for (int s = 0; s < NW2; ++s) { mbar_init(BAR(BAR_W_FULL + s), 1); mbar_init(BAR(BAR_W_EMPTY + s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(BAR(BAR_D2_FULL + s), 1); mbar_init(BAR(BAR_D2_FREE + s), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float *s_b1 = reinterpret_cast<float *>(sm + OFF_B1), *s_b2 = reinterpret_cast<float *>(sm + OFF_B2);
    for (int i = tid; i < 128; i += NTHREADS) s_b1[i] = __ldg(P.b1 + i);
    for (int i = tid; i < P.nch * 128; i += NTHREADS) s_b2[i] = i < P.d_out ? __ldg(P.b2 + i) : 0.0f;
    __syncthreads();
    if (tid == 0 && !P.head2) {
        mbar_expect_tx(BAR(BAR_WLOAD), 16384);
        bulk_g2s(sm_u + OFF_W1, P.w1_img, 16384, BAR(BAR_WLOAD));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);
    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_tiles = P.n_tiles > first ? (P.n_tiles - first + stride - 1) / stride : 0;
    // chunk steps per tile: pass 0 (norm) and pass 1 (store); head2 has no normalisation: one pass
    const int nsteps = P.head2 ? P.nch : 2 * P.nch;
    const int first_store = P.head2 ? 0 : P.nch;

    if (warp < 4) {
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int r_tile = warp * 32 + lane;
        unsigned char *a_row = sm + OFF_A + r_tile * 128;
        unsigned char *stage = sm + OFF_STAGE + warp * STAGE_WARP;
        const uint32_t stage_u = smem_u32(stage);
        long long g = 0;                      // chunk steps so far (accumulator = g & 1)
        for (long long j = 0; j < my_tiles; ++j) {
            const long long tile = first + j * stride;
            const long long row = tile * TM + r_tile;
            float bscale = 1.0f;
            if (P.head2) {
                // ---- hidden rows given: fp32 -> fp16 pairs straight into the TMEM columns layer 2 reads as its A operand
                const uint4 *src = reinterpret_cast<const uint4 *>(P.h + row * 128);
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb) {
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        uint4 a = make_uint4(0u, 0u, 0u, 0u);
                        if (row < P.N) a = __ldg(src + kb * 8 + e);
                        pk[2 * e] = pack_h2(__uint_as_float(a.x), __uint_as_float(a.y));
                        pk[2 * e + 1] = pack_h2(__uint_as_float(a.z), __uint_as_float(a.w));
                    }
                    tmem_st16(t_lane + kb * 16, pk);
                }
                if (P.scale && row < P.N) bscale = __ldg(P.scale + row);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_H_FULL));
            } else {
            // ---- A operand: this thread's row, fp32 -> fp16, 16-byte chunk q at position q ^ (row & 7) ----------
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(P.f + row * 64);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (row < P.N) {
                        const uint4 a = __ldg(src + 2 * q), b = __ldg(src + 2 * q + 1);
                        o.x = pack_h2(__uint_as_float(a.x), __uint_as_float(a.y));
                        o.y = pack_h2(__uint_as_float(a.z), __uint_as_float(a.w));
                        o.z = pack_h2(__uint_as_float(b.x), __uint_as_float(b.y));
                        o.w = pack_h2(__uint_as_float(b.z), __uint_as_float(b.w));
                    }
                    *reinterpret_cast<uint4 *>(a_row + ((q ^ (r_tile & 7)) << 4)) = o;
                }
            }
            fence_proxy_async();              // generic-proxy writes -> visible to the MMA's async-proxy reads
            mbar_arrive_warp(BAR(BAR_A_FULL));
            // ---- epilogue 1: hidden = relu(D1 + b1) as fp16 pairs over the columns already read ------------------
            mbar_wait(BAR(BAR_D1_FULL), (uint32_t)(j & 1));
            tc_fence_after();
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
                uint32_t vr[32];
                tmem_ld32_issue(t_lane + kb * 32, vr);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    pk[e] = pack_h2_relu(__uint_as_float(vr[2 * e]) + s_b1[kb * 32 + 2 * e],
                                         __uint_as_float(vr[2 * e + 1]) + s_b1[kb * 32 + 2 * e + 1]);
                tmem_st16(t_lane + kb * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(BAR(BAR_H_FULL));
            }
            // ---- epilogue 2: two passes over the output chunks (head2: one) -------------------------------------
            float ss = 0.0f, inv = 1.0f;
            for (int s = 0; s < nsteps; ++s, ++g) {
                const int b = (int)(g & 1);
                const int c = s < P.nch ? s : s - P.nch;
                const bool store = s >= first_store;
                if (s == P.nch) {
                    const float nrm = sqrtf(ss);
                    inv = 1.0f / (nrm > 1e-12f ? nrm : 1e-12f);
                }
                mbar_wait(BAR(BAR_D2_FULL + b), (uint32_t)((g >> 1) & 1));
                tc_fence_after();
                if (store) __syncwarp();      // the copy-out of the previous chunk has read the staging rows
                {   // the whole 128-column chunk of this row in one batch of TMEM loads (one wait instead of four)
                    uint32_t vr[128];
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) tmem_ld32_issue(t_lane + D2_COL + b * 128 + qd * 32, vr + qd * 32);
                    tmem_ld_wait();
                    const float *bias = s_b2 + c * 128;
                    if (!store) {
#pragma unroll
                        for (int e = 0; e < 128; ++e) {
                            const float v = __uint_as_float(vr[e]) + bias[e];
                            ss = fmaf(v, v, ss);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            float4 o;
                            o.x = fmaf(bias[4 * e + 0], bscale, __uint_as_float(vr[4 * e + 0])) * inv;
                            o.y = fmaf(bias[4 * e + 1], bscale, __uint_as_float(vr[4 * e + 1])) * inv;
                            o.z = fmaf(bias[4 * e + 2], bscale, __uint_as_float(vr[4 * e + 2])) * inv;
                            o.w = fmaf(bias[4 * e + 3], bscale, __uint_as_float(vr[4 * e + 3])) * inv;
                            *reinterpret_cast<float4 *>(stage + lane * STAGE_ROW + e * 16) = o;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_D2_FREE + b));       // (its __syncwarp also orders the staging writes)
                if (store) {
                    // copy-out: row rr of the warp is one 512-byte STG.128 of the whole warp
                    const long long row0 = tile * TM + warp * 32;
                    float *dst = P.out + row0 * P.d_out + c * 128 + lane * 4;
                    const int cols = P.d_out - (c * 128 + lane * 4);      // columns of this lane's quad inside the row
#pragma unroll 1
                    for (int r0 = 0; r0 < 32; r0 += 8) {
                        float4 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = lds128_ordered(stage_u + (r0 + i) * STAGE_ROW + lane * 16);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (row0 + r0 + i >= P.N) continue;
                            float *d = dst + (long long)(r0 + i) * P.d_out;
                            if (cols >= 4 && !(P.d_out & 3)) stg128_ordered(d, v[i]);
                            else {                                   // ragged last chunk / rows that are not 16-byte aligned
                                if (cols > 0) d[0] = v[i].x;
                                if (cols > 1) d[1] = v[i].y;
                                if (cols > 2) d[2] = v[i].z;
                                if (cols > 3) d[3] = v[i].w;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // (whole warp in lockstep; copies inside an elect_one() branch, MMAs / commits in their elected forms: tc_common.cuh)
        long long g = 0;
        for (long long j = 0; j < my_tiles; ++j)
            for (int s = 0; s < nsteps; ++s, ++g) {
                const int slot = (int)(g % NW2);
                const int c = s < P.nch ? s : s - P.nch;
                mbar_wait_warp(BAR(BAR_W_EMPTY + slot), (uint32_t)(((g / NW2) & 1) ^ 1));
                if (elect_one()) {
                    mbar_expect_tx(BAR(BAR_W_FULL + slot), W2_BYTES);
                    const unsigned char *src = P.w2_img + (size_t)c * W2_BYTES;
                    bulk_g2s(sm_u + OFF_W2 + slot * W2_BYTES, src, 16384, BAR(BAR_W_FULL + slot));
                    bulk_g2s(sm_u + OFF_W2 + slot * W2_BYTES + 16384, src + 16384, 16384, BAR(BAR_W_FULL + slot));
                }
                __syncwarp();
            }
    } else {
        {
            if (!P.head2) mbar_wait_warp(BAR(BAR_WLOAD), 0);
            const uint32_t idesc = umma_idesc(TM, 128);
            long long g = 0;
            for (long long j = 0; j < my_tiles; ++j) {
                if (!P.head2) {
                    mbar_wait_warp(BAR(BAR_A_FULL), (uint32_t)(j & 1));
                    tc_fence_after();
                    if (elect_one()) {          // ONE elected thread issues: a single-thread region (tc_common.cuh)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma(tmem_base, umma_desc(sm_u + OFF_A + k * 32), umma_desc(sm_u + OFF_W1 + k * 32), idesc, k != 0);
                        umma_commit(BAR(BAR_D1_FULL));
                    }
                    __syncwarp();
                }
                mbar_wait_warp(BAR(BAR_H_FULL), (uint32_t)(j & 1));
                tc_fence_after();
                for (int s = 0; s < nsteps; ++s, ++g) {
                    const int slot = (int)(g % NW2), b = (int)(g & 1);
                    mbar_wait_warp(BAR(BAR_W_FULL + slot), (uint32_t)((g / NW2) & 1));
                    mbar_wait_warp(BAR(BAR_D2_FREE + b), (uint32_t)(((g >> 1) & 1) ^ 1));
                    tc_fence_after();
                    const uint32_t w2 = sm_u + OFF_W2 + slot * W2_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)          // K = 16 per instruction = 8 packed columns of the hidden tile
                            umma_ts(tmem_base + D2_COL + b * 128, tmem_base + k * 8, umma_desc(w2 + (k >> 2) * 16384 + (k & 3) * 32),
                                    idesc, k != 0);
                        umma_commit(BAR(BAR_W_EMPTY + slot));
                        umma_commit(BAR(BAR_D2_FULL + b));
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace ex

bool expand_tc_supported(const sd_mlp *mlp) {
    return mlp && mlp->packed && mlp->d_in == 64 && mlp->d_hidden == 128 && mlp->d_out >= 128 && mlp->d_out % 128 == 0 &&
           mlp->d_out <= ex::MAX_DOUT;
}

int launch_expand_tc(const sd_mlp *mlp, const float *f, long long N, float *out, cudaStream_t st) {
    SD_REQUIRE(expand_tc_supported(mlp), "expand_dim on tensor cores: needs a packed 64 -> 128 -> (multiple of 128, <= 1024) head");
    SD_REQUIRE(N >= 0, "sd_expand_dim: bad N");
    if (N == 0) return SD_OK;
    SD_REQUIRE(f && out, "sd_expand_dim: null pointer");
    SD_REQUIRE(((uintptr_t)f & 15) == 0 && ((uintptr_t)out & 15) == 0, "sd_expand_dim: f and out must be 16-byte aligned");
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    SD_REQUIRE(L.off_x_w2 != 0, "sd_expand_dim: the blob carries no expand images");
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    SD_REQUIRE(((uintptr_t)blob & 15) == 0, "mlp: packed blob must be 16-byte aligned");
    ex::Params P = {};
    P.f = f; P.out = out;
    P.w1_img = blob + L.off_w_in_h;
    P.w2_img = blob + L.off_x_w2;
    P.b1 = reinterpret_cast<const float *>(blob + L.off_b_in);
    P.b2 = reinterpret_cast<const float *>(blob + L.off_b_out);
    P.N = N; P.n_tiles = (N + ex::TM - 1) / ex::TM;
    P.d_out = mlp->d_out; P.nch = mlp->d_out / 128;
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(ex::expand_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ex::SMEM_ALLOC));
    }
    const unsigned grid = (unsigned)(P.n_tiles < sm_count ? P.n_tiles : sm_count);
    profile_before(st);
    ex::expand_tc_kernel<<<grid, ex::NTHREADS, ex::SMEM_ALLOC, st>>>(P);
    profile_after(st);
    SD_LAUNCH_OK("expand_tc_kernel");
    return SD_OK;
}

int launch_head2(const sd_mlp *mlp, const float *h, const float *scale, long long R, float *out, cudaStream_t st) {
    SD_REQUIRE(mlp && mlp->packed && mlp->d_hidden == 128 && mlp->d_out >= 2 && mlp->d_out - 1 <= ex::MAX_DOUT,
               "head2: needs a packed head with d_hidden = 128 and at most %d feature outputs", ex::MAX_DOUT);
    if (R == 0) return SD_OK;
    SD_REQUIRE(R > 0 && h && out, "head2: null pointer");
    SD_REQUIRE(((uintptr_t)h & 15) == 0 && ((uintptr_t)out & 15) == 0, "head2: h and out must be 16-byte aligned");
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    SD_REQUIRE(L.off_w_feat_blk != 0, "head2: the blob carries no feature-row images");
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    SD_REQUIRE(((uintptr_t)blob & 15) == 0, "mlp: packed blob must be 16-byte aligned");
    ex::Params P = {};
    P.head2 = 1; P.h = h; P.scale = scale; P.out = out;
    P.w2_img = blob + L.off_w_feat_blk;
    P.b1 = reinterpret_cast<const float *>(blob + L.off_b_in);               // (unused)
    P.b2 = reinterpret_cast<const float *>(blob + L.off_b_out) + 1;          // biases of the feature rows W_out[1:]
    P.N = R; P.n_tiles = (R + ex::TM - 1) / ex::TM;
    P.d_out = mlp->d_out - 1; P.nch = (P.d_out + 127) / 128;
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(ex::expand_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ex::SMEM_ALLOC));
    }
    const unsigned grid = (unsigned)(P.n_tiles < sm_count ? P.n_tiles : sm_count);
    ex::expand_tc_kernel<<<grid, ex::NTHREADS, ex::SMEM_ALLOC, st>>>(P);
    SD_LAUNCH_OK("expand_tc_kernel");
    return SD_OK;
}

}  // namespace sd
