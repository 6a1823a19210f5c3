// Fused tcgen05 implementation of the field query and of one render pass (sd_precision SD_MLP_F16_TC).
//
//   BTSNet.forward      models/bts.py:476-595     (projection, mask, gather, code, head, softplus, colours)
//   NeRFRenderer.composite renderer/nerf.py:230-449 (points along rays, the above, alpha compositing)
//   ResnetFC.forward    resnetfc.py:135-203        (rows mode: the head alone, unit test of the MMA path)
//
// One persistent CTA per SM walks 128-row tiles (rows = points, or samples of 128/K whole rays).
// Warp roles (16 warps):
//   warps 0-3   epilogue: layer-1 accumulator -> ReLU -> fp16 hidden tile written back to TMEM (the A operand of
//               layer 2, tcgen05.st); second epilogue: density + softplus, features out through shared memory
//               as coalesced stores, or composited (segmented transmittance scan + butterfly weighted sums)
//   warp  4     tcgen05.mma issuer (one lane) + TMEM owner
//   warps 5-8   one thread per row: ray point, projection, frustum mask, bilinear tap, colour
//               lookup, positional code -> code chunk of the A operand
//   warps 9-15  gather ((chunk, 16-row group) units dealt round-robin): 8 lanes per row read 4 texels x 64 channels (one 128-byte line each, 128-bit
//               loads) of the channels-last fp16 map, blend with packed HFMA2, store into the A operand
// A operand: 5 K-chunks of [128 rows x 64] fp16, K-major SWIZZLE_128B, one mbarrier pair per chunk, so
// layer-1 MMAs start while later chunks of the same tile are still being gathered and the gather of
// tile t+1 overlaps the MMAs/epilogues of tile t.  W_in / W_out live in shared memory for the whole
// kernel (loaded once with cp.async.bulk), accumulators in TMEM (layer 1 double-buffered).
#include <cstdlib>

#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {
namespace tc {

constexpr int TM = 128;                      // rows per tile = UMMA M
constexpr int MAX_CHUNKS = 5;                // K chunks of 64 (C = 256 -> 4 feature chunks + 1 code chunk)
constexpr int CHUNK_BYTES = TM * 128;        // 16 KB: [128 rows][64 fp16]
constexpr int N_EPI_WARPS = 4, N_PT_WARPS = 4, N_GA_WARPS = 7;   // 16 warps -> 128 registers per thread
constexpr int WARP_MMA = N_EPI_WARPS;
constexpr int WARP_PT0 = WARP_MMA + 1;
constexpr int WARP_GA0 = WARP_PT0 + N_PT_WARPS;
constexpr int NTHREADS = (N_EPI_WARPS + 1 + N_PT_WARPS + N_GA_WARPS) * 32;   // 512
constexpr int NGEO = 3;
constexpr int TMEM_COLS = 512;
constexpr int D2_COL = 256;                  // layer-1 accumulators at columns 0 and 128, layer 2 (<= 80 columns) at 256
constexpr int H_COL = 384;                   // fp16 hidden tile (A operand of layer 2): 64 columns, two halves per column
constexpr int D3X_COL = 336, D3F_COL = 448;  // composite on the tensor cores: per-ray sums of (depth, sum w, colours) / of the 64 features
// hidden-composite mode (cmma == 2): layer 2 is the density row alone (16 columns at D2_COL), the composite sums the 128
// ReLU'd hidden units per ray (the feature rows of W_out are applied to the per-ray sums afterwards: both are linear)
constexpr int D3X_COL2 = 272, H_COL2 = 288, D3H_COL2 = 352;
constexpr int PART_STRIDE = 80;              // per (warp, segment) composite partial: 64 feat + depth + wsum + 12 rgb
constexpr int MAX_NVC_TC = 4;
constexpr int W2_BYTES = 2 * 80 * 128;       // [2 K blocks][<= 80 rows][128 B]

enum { MODE_POINTS = 0, MODE_RENDER = 1, MODE_ROWS = 2 };

struct Geo {                                 // per-row hand-off from the point warps, structure of arrays (3072 B)
    uint32_t o[TM];                          // byte offset of the north-west texel of the (clamped) 2x2 footprint
    uint32_t w01[TM], w23[TM];               // bilinear weights as halves (nw, ne), (sw, se); zero for unused rows
    int flags[TM];                           // bit0 out of frustum, bit3 row valid
    float z[TM];                             // sample depth (render mode)
    int grow[TM];                            // global row (point / sample index) this tile row stands for, -1 if unused
};

// shared-memory map, offsets from a 1024-byte aligned base
constexpr int OFF_W1 = 0;
constexpr int OFF_RING = OFF_W1 + MAX_CHUNKS * CHUNK_BYTES;
constexpr int OFF_H = OFF_RING + MAX_CHUNKS * CHUNK_BYTES;
constexpr int OFF_W2 = OFF_H + 2 * CHUNK_BYTES;
constexpr int OFF_GEO = OFF_W2 + W2_BYTES;
constexpr int OFF_EMPTY = OFF_GEO + NGEO * (int)sizeof(Geo);   // fp16 empty_feature [256]
constexpr int OFF_CAM = OFF_EMPTY + 512;     // 21 floats per camera, 1 + 4 cameras
constexpr int OFF_PART = OFF_CAM + 448;
constexpr int OFF_TAILS = OFF_PART + 4 * 2 * PART_STRIDE * 4;
constexpr int OFF_BAR = OFF_TAILS + 32;
constexpr int NBAR = 2 * MAX_CHUNKS + 2 + 1 + 1 + 2 * NGEO + 1 + 3;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(OFF_BAR % 8 == 0 && OFF_PART % 16 == 0 && OFF_GEO % 16 == 0 && OFF_W2 % 1024 == 0, "alignment");

enum { BAR_FULL = 0, BAR_EMPTY = MAX_CHUNKS, BAR_D1 = 2 * MAX_CHUNKS, BAR_H = BAR_D1 + 2, BAR_D2 = BAR_H + 1,
       BAR_GEO_FULL = BAR_D2 + 1, BAR_GEO_EMPTY = BAR_GEO_FULL + NGEO, BAR_WLOAD = BAR_GEO_EMPTY + NGEO,
       BAR_B3 = BAR_WLOAD + 1, BAR_D3 = BAR_B3 + 1, BAR_D3_READ = BAR_D3 + 1 };
// Composite on the tensor cores (render mode on a projected scene, where layer 1 needs 3 of the 5 operand chunks):
//   out[ray][col] = sum_row Wt[ray][row] * V[row][col],  V = (64 features | depth, 1, colours) of the tile's rows as fp16
// Wt (A operand, K-major) lives in the two spare chunks of the W_in image, V (B operand, MN-major like the boxes of
// field_bin.cu) in the output staging area, which render mode does not use.
constexpr int OFF_A3 = OFF_W1 + 3 * CHUNK_BYTES;
constexpr int OFF_B3 = OFF_H;
constexpr int OFF_X3 = OFF_H + CHUNK_BYTES;          // (depth, 1, colours, depth lo) rows: B operand of the small composite MMA
// cmma == 2: V = the tile's ReLU'd hidden rows as fp16, B operand [128 rows (K)][128 units (N)], MN-major, two 64-unit
// halves of 16 KB in the two ring chunks a projected scene leaves unused; single buffered (the first epilogue of tile
// j waits for the composite of tile j-1)
constexpr int OFF_V = OFF_RING + 3 * CHUNK_BYTES;
// per-sample colours of a tile as packed halves (12 values + padding = 32 B per row), handed from the point warps to the
// epilogue next to the Geo slots: the composite takes them as fp16 anyway, so they never make a round trip through HBM
constexpr int COL_SLOT = TM * 32;
constexpr int OFF_COL1 = OFF_RING + 3 * CHUNK_BYTES, OFF_COL2 = OFF_H;   // cmma == 1 / cmma == 2
static_assert(NGEO * COL_SLOT <= CHUNK_BYTES, "colour ring fits into one spare chunk");

struct Params {
    int mode;
    int dbg;                   // SD_TC_DEBUG ablation mask (timing experiments only; results are wrong when non-zero)
    FieldParams fp;
    PointSrc src;
    const unsigned int *perm;  // point mode: tile row r of tile t stands for point perm[128 t + r] (texel-binned order) or NULL
    long long n_units;        // points / rays / rows
    int K;                    // rows per unit (1 unless render)
    int upt;                  // units per tile
    long long n_tiles;
    int nch;                  // K chunks of layer 1
    int cmma;                 // render mode on a projected scene, composite on the tensor cores: 1 = of the <= 64 features,
                              // 2 = of the 128 hidden units (any D; per-ray sums to hsum / wsum, W_out applied by the head2 kernel)
    int sig_col;              // column of the density in the layer-2 accumulator
    float *hsum, *wsum;       // cmma == 2: [rays][128] per-ray sums of w * relu(hidden), [rays] sums of w
    int n2;                   // layer-2 N: D feature rows + the density row, padded to 16
    int D;                    // feature outputs = d_out - 1
    const unsigned char *w1_img, *w2_img;
    const __half *empty_h;     // projected scene: W_feat . empty_feature as halves (else NULL: fp.empty_feature is used)
    const float *b_out;
    // rows mode
    const float *x_rows;
    int d_in;
    float *out_rows;
    // per-row outputs (any may be NULL)
    float *sigma, *dino, *rgb, *invalid;
    unsigned char *invalid_feat;
    // per-ray outputs (render)
    sd_render_cfg cfg;
    float *depth, *dino_ray, *rgb_ray, *weights, *alphas;
};

__device__ long long g_trace[4 * 64 * 8];   // [role][tile][event] clock64 stamps of CTA 0 (SD_TC_DEBUG & 8192)

using namespace tcx;

#define SD_TRACE(role, j, ev)                                                                    \
    do {                                                                                         \
        if ((P.dbg & 8192) && blockIdx.x == 0 && (j) < 64 && (threadIdx.x & 31) == 0)             \
            g_trace[((role) * 64 + (int)(j)) * 8 + (ev)] = clock64();                            \
    } while (0)

// ---- the kernel -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) field_tc_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the shared array itself: going through
    // an integer would make every shared access of the kernel a generic LD/ST
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;   // (warp index the compiler knows to be uniform)
    const uint32_t bar0 = sm_u + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    float *s_cam = reinterpret_cast<float *>(sm + OFF_CAM);
    Geo *s_geo = reinterpret_cast<Geo *>(sm + OFF_GEO);
    const bool field = P.mode != MODE_ROWS;
    const bool render = P.mode == MODE_RENDER;

    // ---- one-time setup ------------------------------------------------------------------------------
    if (tid == 0) {
        for (int c = 0; c < MAX_CHUNKS; ++c) {
            mbar_init(BAR(BAR_FULL + c), (field && c == P.nch - 1) ? N_PT_WARPS : N_GA_WARPS);   // [0]: all gathered chunks of a tile
            mbar_init(BAR(BAR_EMPTY + c), 1);
        }
        mbar_init(BAR(BAR_D1), 1); mbar_init(BAR(BAR_D1 + 1), 1);
        mbar_init(BAR(BAR_H), N_EPI_WARPS);
        mbar_init(BAR(BAR_D2), 1);
        for (int s = 0; s < NGEO; ++s) {
            mbar_init(BAR(BAR_GEO_FULL + s), N_PT_WARPS);
            mbar_init(BAR(BAR_GEO_EMPTY + s), N_GA_WARPS + N_EPI_WARPS);
        }
        mbar_init(BAR(BAR_WLOAD), 1);
        mbar_init(BAR(BAR_B3), N_EPI_WARPS); mbar_init(BAR(BAR_D3), 1); mbar_init(BAR(BAR_D3_READ), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (field) {
        for (int i = tid; i < 21 * (1 + P.fp.nv_c); i += NTHREADS) {
            const int c = i / 21, e = i - 21 * c;
            const float *K = c == 0 ? P.fp.K_f : P.fp.K_c + 9 * (c - 1);
            const float *W = c == 0 ? P.fp.w2c_f : P.fp.w2c_c + 16 * (c - 1);
            s_cam[i] = e < 9 ? __ldg(K + e) : __ldg(W + (e - 9));
        }
        __half *s_empty = reinterpret_cast<__half *>(sm + OFF_EMPTY);
        for (int i = tid; i < 256; i += NTHREADS) {
            __half v = __float2half_rn(0.0f);
            if (P.fp.learn_empty && i < P.fp.C) v = P.empty_h ? P.empty_h[i] : __float2half_rn(__ldg(P.fp.empty_feature + i));
            s_empty[i] = v;
        }
    }
    if (P.cmma) {   // rows of the weight operand beyond the rays of a tile are never written: keep them zero
        for (int i = tid; i < 2 * CHUNK_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm + OFF_A3)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
        // the partial-sum area is free in this mode: keep the output bias of the features there
        if (P.cmma == 1 && tid < 64) reinterpret_cast<float *>(sm + OFF_PART)[tid] = tid < P.D ? __ldg(P.b_out + 1 + tid) : 0.0f;
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();   // barrier inits visible before anyone (including the bulk copies) uses them
    if (tid == 0) {    // weights: global -> shared through the bulk-copy engine (async proxy, as UMMA reads them)
        const uint32_t w1b = (uint32_t)P.nch * CHUNK_BYTES, w2b = 2u * (uint32_t)P.n2 * 128u;
        mbar_expect_tx(BAR(BAR_WLOAD), w1b + w2b);
        for (int c = 0; c < P.nch; ++c)
            bulk_g2s(sm_u + OFF_W1 + c * CHUNK_BYTES, P.w1_img + (size_t)c * CHUNK_BYTES, CHUNK_BYTES, BAR(BAR_WLOAD));
        bulk_g2s(sm_u + OFF_W2, P.w2_img, w2b, BAR(BAR_WLOAD));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);
    const uint32_t h_col = P.cmma == 2 ? H_COL2 : H_COL, d3x_col = P.cmma == 2 ? D3X_COL2 : D3X_COL;
    const uint32_t d3f_col = P.cmma == 2 ? D3H_COL2 : D3F_COL;
    const int off_col = P.cmma == 2 ? OFF_COL2 : OFF_COL1;

    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_tiles = P.n_tiles > first ? (P.n_tiles - first + stride - 1) / stride : 0;

    if (warp < N_EPI_WARPS) {
        // =================================== EPILOGUE ================================================
        // Per tile: (1) layer-1 accumulator -> ReLU -> fp16 hidden tile (the bias came out of the MMA);
        // (2) layer-2 accumulator -> density (+softplus) and features; the features leave the SM through
        // the warp's own 8 KB of the hidden tile as fully coalesced 512-byte stores, or are composited.
        const int row = tid;                                     // TMEM lane == tile row
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int K = P.K, D = P.D;
        const int k_i = row % K;
        const bool row_used = row < P.upt * K;
        float *s_part = reinterpret_cast<float *>(sm + OFF_PART);
        float *s_tails = reinterpret_cast<float *>(sm + OFF_TAILS);
        const float bo_sigma = __ldg(P.b_out);
        float bo4[4];                                            // output bias of the 4 columns this lane stores
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int c = 4 * (lane & 15) + e; bo4[e] = c < D ? __ldg(P.b_out + 1 + c) : 0.0f; }
        unsigned char *stage0 = sm + OFF_H + warp * 4096, *stage1 = stage0 + CHUNK_BYTES;   // this warp's rows of H

        float z_keep = 0.0f, zn_keep = 0.0f;
        long long grow_keep = -1;
        uint32_t crgb_h[6] = {0u, 0u, 0u, 0u, 0u, 0u};           // this row's colours as packed halves (cmma)
        // composite on the tensor cores: the per-ray sums are read back and written out by point warp 3 (cm_output)
        long long cm_n = 0;          // composites handed to the MMA issuer so far
        for (long long j = 0; j <= my_tiles; ++j) {
            if (j > 0) {
                // ---------------- second epilogue of tile j-1 -----------------------------------------
                // per-sample colours of this row (written by the point warps a tile ago): in flight during the wait
                const int nrgb = 3 * P.fp.nv_c;
                float crgb[3 * MAX_NVC_TC];
#pragma unroll
                for (int c = 0; c < 3 * MAX_NVC_TC; ++c)   // (cmma: the colours came through shared memory as halves, crgb_h)
                    crgb[c] = (render && !P.cmma && grow_keep >= 0 && c < nrgb && P.rgb) ? __ldcg(P.rgb + grow_keep * nrgb + c) : 0.0f;
                mbar_wait(BAR(BAR_D2), (uint32_t)((j - 1) & 1));
                tc_fence_after();
                if (warp == 0) SD_TRACE(0, j, 0);
                const long long tile = first + (j - 1) * stride;
                uint32_t vr[64], sr;
                if (P.cmma != 2) {
                    tmem_ld32_issue(t_lane + D2_COL, vr);
                    tmem_ld32_issue(t_lane + D2_COL + 32, vr + 32);
                }
                tmem_ld1_issue(t_lane + D2_COL + P.sig_col, sr);                  // density column (behind the features)
                tmem_ld_wait();
                float v[64];
#pragma unroll
                for (int c = 0; c < 64; ++c) v[c] = P.cmma != 2 ? __uint_as_float(vr[c]) : 0.0f;
                const float sig = __uint_as_float(sr) + bo_sigma;
                const float sg = P.mode == MODE_ROWS ? sig : softplus(sig);
                const bool ok = grow_keep >= 0;
                if (ok && P.sigma) P.sigma[grow_keep] = sg;
                if (P.mode == MODE_ROWS) {
                    if (ok) {
                        float *o = P.out_rows + grow_keep * (D + 1);
                        o[0] = sg;
#pragma unroll
                        for (int c = 0; c < 64; ++c)
                            if (c < D) o[1 + c] = v[c] + __ldg(P.b_out + 1 + c);
                    }
                } else if (!render) {
                    if (P.dino && D == 64) {
                        // transpose through shared memory: lane = row writes its 16 chunks (XOR-swizzled, conflict
                        // free), then 16 lanes read one row back and the warp stores two whole rows per request
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            *reinterpret_cast<float4 *>((q < 8 ? stage0 : stage1) + lane * 128 + (((q & 7) ^ (lane & 7)) << 4)) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        __syncwarp();
                        (void)tile;
                        const int c16 = lane & 15;
                        const unsigned char *src = (c16 < 8 ? stage0 : stage1);
                        const int g32 = (int)grow_keep;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int r = 2 * i + (lane >> 4);
                            float4 x = *reinterpret_cast<const float4 *>(src + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
                            x.x += bo4[0]; x.y += bo4[1]; x.z += bo4[2]; x.w += bo4[3];
                            const int dst = __shfl_sync(0xffffffffu, g32, r);
                            if (dst >= 0) *reinterpret_cast<float4 *>(P.dino + (long long)dst * 64 + c16 * 4) = x;
                        }
                        __syncwarp();
                    } else if (P.dino && ok) {
                        float *o = P.dino + grow_keep * D;
#pragma unroll
                        for (int c = 0; c < 64; ++c)
                            if (c < D) o[c] = v[c] + __ldg(P.b_out + 1 + c);
                    }
                } else {
                    // ---- alpha, exclusive transmittance (segmented product scan over the tile), weight
                    //      (nerf.py:376-389) --------------------------------------------------------------
                    const bool last = k_i == K - 1;
                    const float delta = last ? 1e10f : zn_keep - z_keep;
                    float alpha = 0.0f;
                    if (ok) {
                        alpha = 1.0f - expf(-fabsf(delta) * fmaxf(sg, 0.0f));
                        if (P.cfg.hard_alpha_cap && last) alpha = 1.0f;
                    }
                    float incl = ok ? (1.0f - alpha) + 1e-10f : 1.0f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float n = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o && k_i >= o) incl *= n;
                    }
                    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                    if (lane == 0 || k_i == 0) excl = 1.0f;
                    if (lane == 31) s_tails[(int)(j & 1) * 4 + warp] = incl;
                    named_bar_sync(1, N_EPI_WARPS * 32);
                    float carry = 1.0f;
                    for (int prev = k_i - lane, w = warp - 1; prev > 0 && w >= 0; prev -= 32, --w) carry *= s_tails[(int)(j & 1) * 4 + w];
                    const float wgt = ok ? alpha * (excl * carry) : 0.0f;
                    if (ok) {
                        if (P.weights) P.weights[grow_keep] = wgt;
                        if (P.alphas) P.alphas[grow_keep] = alpha;
                    }
                    // ---- weighted sums over the rows of each ray (nerf.py:393-405) ----------------------
                    const int ua = (warp * 32) / K, ub = (warp * 32 + 31) / K;   // rays this warp's rows belong to
                    const int my_u = row / K;
                    if (P.cmma) {
                        // ---- composite on the tensor cores: this row's values and weight go into the operands of
                        //      out[ray][:] = sum_row Wt[ray][row] * V[row][:]; the MMA issuer does the rest ---------
                        if (warp == 0) SD_TRACE(0, j, 5);
                        // the previous composite has consumed the operands and its sums have been read from TMEM
                        if (cm_n > 0) mbar_wait(BAR(BAR_D3_READ), (uint32_t)((cm_n - 1) & 1));
                        if (warp == 0) SD_TRACE(0, j, 6);
                        if (P.cmma == 1) {
                            unsigned char *brow = sm + OFF_B3 + row * 128;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {              // 64 features (bias added at the output)
                                uint4 o;
                                o.x = ok ? pack_h2(v[8 * q + 0], v[8 * q + 1]) : 0u; o.y = ok ? pack_h2(v[8 * q + 2], v[8 * q + 3]) : 0u;
                                o.z = ok ? pack_h2(v[8 * q + 4], v[8 * q + 5]) : 0u; o.w = ok ? pack_h2(v[8 * q + 6], v[8 * q + 7]) : 0u;
                                *reinterpret_cast<uint4 *>(brow + ((q ^ (row & 7)) << 4)) = o;
                            }
                        }
                        {   // depth (hi), 1 (sum of weights), 12 colours, depth (lo): the depth keeps ~21 bits through fp16
                            const float z_hi = __half2float(__float2half_rn(z_keep));
                            unsigned char *xrow = sm + OFF_X3 + row * 128;
                            *reinterpret_cast<uint4 *>(xrow + ((0 ^ (row & 7)) << 4)) =
                                ok ? make_uint4(pack_h2(z_hi, 1.0f), crgb_h[0], crgb_h[1], crgb_h[2]) : make_uint4(0u, 0u, 0u, 0u);
                            *reinterpret_cast<uint4 *>(xrow + ((1 ^ (row & 7)) << 4)) =
                                ok ? make_uint4(crgb_h[3], crgb_h[4], crgb_h[5], pack_h2(z_keep - z_hi, 0.0f)) : make_uint4(0u, 0u, 0u, 0u);
                        }
                        const unsigned short wh = __half_as_ushort(__float2half_rn(wgt));
                        for (int u = 0; u < P.upt; ++u)                // weight matrix: column = this row, row = ray of the tile
                            *reinterpret_cast<unsigned short *>(sm + OFF_A3 + ((row >> 6) * TM + u) * 128 +
                                                                ((((row & 63) >> 3) ^ (u & 7)) << 4) + (row & 7) * 2) =
                                (u == my_u) ? wh : (unsigned short)0;
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive_warp(BAR(BAR_B3));
                        if (warp == 0) SD_TRACE(0, j, 7);
                        ++cm_n;                                        // (s_tails is double buffered: no barrier needed here)
                    } else {
#pragma unroll 1
                    for (int seg = 0; seg < 2; ++seg) {
                        const int u = seg == 0 ? ua : ub;
                        if (seg == 1 && ub == ua) break;
                        const float ws = (my_u == u) ? wgt : 0.0f;
                        float *pp = s_part + (warp * 2 + seg) * PART_STRIDE;
                        // butterfly over 32 columns at a time: after 5 exchange steps lane l holds column l
                        // summed over the 32 rows of the warp
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            float a[32];
#pragma unroll
                            for (int c = 0; c < 32; ++c) a[c] = ws * v[hh * 32 + c];
#pragma unroll
                            for (int step = 0; step < 5; ++step) {
                                const int m = 16 >> step;
                                const bool up = lane & m;
#pragma unroll
                                for (int c = 0; c < m; ++c) {
                                    const float send = up ? a[c] : a[c + m];
                                    const float keep = up ? a[c + m] : a[c];
                                    a[c] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                                }
                            }
                            pp[hh * 32 + lane] = a[0];
                        }
                        float sd_ = ws * z_keep, sw_ = ws;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            sd_ += __shfl_xor_sync(0xffffffffu, sd_, o);
                            sw_ += __shfl_xor_sync(0xffffffffu, sw_, o);
                        }
                        if (lane == 0) { pp[64] = sd_; pp[65] = sw_; }
#pragma unroll
                        for (int c = 0; c < 3 * MAX_NVC_TC; ++c) {
                            float sc = ws * crgb[c];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
                            if (lane == 0) pp[66 + c] = sc;
                        }
                    }
                    named_bar_sync(1, N_EPI_WARPS * 32);
                    {   // warp u finishes ray u of the tile: fixed summation order over the warps it spans
                        const int u = warp;
                        const long long ray = tile * P.upt + u;
                        if (u < P.upt && ray < P.n_units) {
                            const int w_lo = (u * K) / 32, w_hi = (u * K + K - 1) / 32;
                            float acc[3] = {0.f, 0.f, 0.f};
                            for (int w = w_lo; w <= w_hi; ++w) {
                                const int seg = ((w * 32) / K == u) ? 0 : 1;
                                const float *pp = s_part + (w * 2 + seg) * PART_STRIDE;
                                acc[0] += pp[lane]; acc[1] += pp[lane + 32];
                                if (lane + 64 < 78) acc[2] += pp[lane + 64];
                            }
                            const float wsum = __shfl_sync(0xffffffffu, acc[2], 1);
                            if (P.dino_ray) {   // sum_k w_k (f_k + b) = sum_k w_k f_k + b * sum_k w_k
                                if (lane < D) P.dino_ray[ray * D + lane] = fmaf(__ldg(P.b_out + 1 + lane), wsum, acc[0]);
                                if (lane + 32 < D) P.dino_ray[ray * D + lane + 32] = fmaf(__ldg(P.b_out + 33 + lane), wsum, acc[1]);
                            }
                            if (lane == 0 && P.depth) P.depth[ray] = acc[2];
                            if (lane >= 2 && lane < 2 + nrgb && P.rgb_ray)
                                P.rgb_ray[ray * nrgb + (lane - 2)] = P.cfg.white_bkgd ? acc[2] + 1.0f - wsum : acc[2];
                        }
                    }
                    named_bar_sync(1, N_EPI_WARPS * 32);   // partials / tails consumed before the next tile overwrites them
                    }
                }
            }
            if (warp == 0) SD_TRACE(0, j, 1);
            if (j == my_tiles) break;
            // ---------------- first epilogue of tile j ----------------------------------------------------
            const long long tile = first + j * stride;
            const long long unit = tile * P.upt + row / K;
            const bool ok = row_used && unit < P.n_units;
            grow_keep = ok ? unit * K + k_i : -1;
            const int slot = (int)(j % NGEO);
            if (field) {
                mbar_wait(BAR(BAR_GEO_FULL + slot), (uint32_t)((j / NGEO) & 1));
                grow_keep = s_geo[slot].grow[row];
                if (render) {
                    z_keep = s_geo[slot].z[row];
                    zn_keep = row + 1 < TM ? s_geo[slot].z[row + 1] : 0.0f;
                    if (P.cmma) {
                        const uint4 *cr = reinterpret_cast<const uint4 *>(sm + off_col + slot * COL_SLOT + row * 32);
                        const uint4 c0 = cr[0], c1 = cr[1];
                        crgb_h[0] = c0.x; crgb_h[1] = c0.y; crgb_h[2] = c0.z; crgb_h[3] = c0.w; crgb_h[4] = c1.x; crgb_h[5] = c1.y;
                    }
                }
                mbar_arrive_warp(BAR(BAR_GEO_EMPTY + slot));
            }
            // cmma == 2: V is single buffered -- the composite MMAs of the previous tile must have consumed it.  (Their
            // completion, BAR_D3, not the read-out of their sums, BAR_D3_READ: the issuer reads the sums out only behind
            // layer 2 of THIS tile, which waits for this epilogue -- that wait would close a cycle.)
            if (P.cmma == 2 && cm_n > 0) mbar_wait(BAR(BAR_D3), (uint32_t)((cm_n - 1) & 1));
            if (warp == 0) SD_TRACE(0, j, 2);
            const int b = (int)(j & 1);
            mbar_wait(BAR(BAR_D1 + b), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (warp == 0) SD_TRACE(0, j, 3);
#pragma unroll 1
            for (int kb = 0; kb < 2; ++kb) {                     // 64 hidden units -> 32 packed columns of the layer-2 A operand
                uint32_t vr[64], pk[32];
                tmem_ld32_issue(t_lane + b * 128 + kb * 64, vr);
                tmem_ld32_issue(t_lane + b * 128 + kb * 64 + 32, vr + 32);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    pk[e] = pack_h2(fmaxf(__uint_as_float(vr[2 * e]), 0.0f), fmaxf(__uint_as_float(vr[2 * e + 1]), 0.0f));
                tmem_st32(t_lane + h_col + kb * 32, pk);
                if (P.cmma == 2) {                               // the same 64 units as one 128-byte row of V's half kb
                    unsigned char *vrow = sm + OFF_V + kb * CHUNK_BYTES + row * 128;
                    const bool live = grow_keep >= 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<uint4 *>(vrow + ((q ^ (row & 7)) << 4)) =
                            live ? make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(BAR(BAR_H));
            if (warp == 0) SD_TRACE(0, j, 4);
        }
    } else if (warp == WARP_MMA) {
        // =================================== MMA ISSUER ===============================================
        // The whole warp runs the issue loop in lockstep and one elected lane issues (tc_common.cuh, "elected" forms: under
        // `if (lane == 0)` every tcgen05 instruction was wrapped in an ELECT loop, ~90 cycles each).  In render mode on a projected scene the whole warp (its TMEM lanes are 0..31 = the rays
        // of a tile) also reads the per-ray sums the composite MMAs leave in TMEM and writes them out.
        {
            mbar_wait_warp(BAR(BAR_WLOAD), 0);
            const uint32_t idesc1 = umma_idesc(TM, 128), idesc2 = umma_idesc(TM, P.n2);
            auto layer2 = [&](long long jj) {
                mbar_wait_warp(BAR(BAR_H), (uint32_t)(jj & 1));
                tc_fence_after();
                SD_TRACE(1, jj + 1, 6);
                if (elect_one()) {          // ONE elected thread issues: a single-thread region (tc_common.cuh)
#pragma unroll
                    for (int k = 0; k < 8; ++k)          // K = 16 per instruction = 8 packed columns of the hidden tile
                        umma_ts(tmem_base + D2_COL, tmem_base + h_col + k * 8,
                                umma_desc(sm_u + OFF_W2 + (k >> 2) * P.n2 * 128 + (k & 3) * 32), idesc2, k != 0);
                    umma_commit(BAR(BAR_D2));
                }
                __syncwarp();
            };
            const uint32_t idesc3f = umma_idesc(TM, P.cmma == 2 ? 128 : 64) | UMMA_B_MN_MAJOR, idesc3x = umma_idesc(TM, 16) | UMMA_B_MN_MAJOR;
            const uint32_t off_v3 = P.cmma == 2 ? OFF_V : OFF_B3;
            auto composite = [&](long long jj) {   // per-ray sums of tile jj: Wt [128 x 128 rows] . V [128 rows x (64 | 128 | 16)]
                mbar_wait_warp(BAR(BAR_B3), (uint32_t)(jj & 1));
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t a = umma_desc(sm_u + OFF_A3 + (k >> 2) * CHUNK_BYTES + (k & 3) * 32);
                        umma(tmem_base + d3f_col, a, umma_desc_mn(sm_u + off_v3 + k * 2048, CHUNK_BYTES, 1024), idesc3f, k != 0);
                        umma(tmem_base + d3x_col, a, umma_desc_mn(sm_u + OFF_X3 + k * 2048, CHUNK_BYTES, 1024), idesc3x, k != 0);
                    }
                    umma_commit(BAR(BAR_D3));
                }
                __syncwarp();
            };
            auto cm_output = [&](long long jj) {
                const float *s_bo = reinterpret_cast<const float *>(sm + OFF_PART);
                const int D = P.D;
                mbar_wait_warp(BAR(BAR_D3), (uint32_t)(jj & 1));
                tc_fence_after();
                uint32_t xr[16], fr[64];
                const int lane_ = lane;
                const long long ray = (first + jj * stride) * P.upt + lane_;   // TMEM lane = ray of the tile
                tmem_ld16_issue(tmem_base + d3x_col, xr);
                if (P.cmma == 2) {
                    // the 128 hidden sums of ray `lane` leave in two halves of 64 columns (a ray's row is 512 contiguous bytes)
#pragma unroll 1
                    for (int hq = 0; hq < 2; ++hq) {
                        tmem_ld32_issue(tmem_base + d3f_col + hq * 64, fr);
                        tmem_ld32_issue(tmem_base + d3f_col + hq * 64 + 32, fr + 32);
                        tmem_ld_wait();
                        if (lane_ < P.upt && ray < P.n_units) {
                            float4 *o = reinterpret_cast<float4 *>(P.hsum + ray * 128 + hq * 64);
#pragma unroll
                            for (int q = 0; q < 16; ++q)
                                o[q] = make_float4(__uint_as_float(fr[4 * q]), __uint_as_float(fr[4 * q + 1]),
                                                   __uint_as_float(fr[4 * q + 2]), __uint_as_float(fr[4 * q + 3]));
                        }
                    }
                } else {
                    tmem_ld32_issue(tmem_base + d3f_col, fr);
                    tmem_ld32_issue(tmem_base + d3f_col + 32, fr + 32);
                    tmem_ld_wait();
                }
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_D3_READ));                       // the accumulators may be overwritten
                if (lane_ < P.upt && ray < P.n_units) {
                    const float wsum = __uint_as_float(xr[1]);
                    const int nrgb = 3 * P.fp.nv_c;
                    if (P.depth) P.depth[ray] = __uint_as_float(xr[0]) + __uint_as_float(xr[14]);
                    if (P.cmma == 2 && P.wsum) P.wsum[ray] = wsum;
                    if (P.rgb_ray)
                        for (int c = 0; c < nrgb; ++c)
                            P.rgb_ray[ray * nrgb + c] = P.cfg.white_bkgd ? __uint_as_float(xr[2 + c]) + 1.0f - wsum : __uint_as_float(xr[2 + c]);
                    if (P.dino_ray && P.cmma == 1) {   // sum_k w_k (f_k + b) = sum_k w_k f_k + b * sum_k w_k
                        if (D == 64) {
                            float4 *o = reinterpret_cast<float4 *>(P.dino_ray + ray * 64);
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float4 b = reinterpret_cast<const float4 *>(s_bo)[q];
                                o[q] = make_float4(fmaf(b.x, wsum, __uint_as_float(fr[4 * q])), fmaf(b.y, wsum, __uint_as_float(fr[4 * q + 1])),
                                                   fmaf(b.z, wsum, __uint_as_float(fr[4 * q + 2])), fmaf(b.w, wsum, __uint_as_float(fr[4 * q + 3])));
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 64; ++c)
                                if (c < D) P.dino_ray[ray * D + c] = fmaf(s_bo[c], wsum, __uint_as_float(fr[c]));
                        }
                    }
                }
            };
            for (long long j = 0; j < my_tiles; ++j) {
                {
                    const uint32_t d1 = tmem_base + (uint32_t)(j & 1) * 128u;
                    // the gather warps fence and arrive once per tile (FULL[0]) for all the chunks they fill -- every
                    // fence.proxy.async / mbarrier round trip queues behind their loads in the LSU --, the point warps
                    // arrive on FULL[nch-1] for the code chunk; the stages are still released one by one (EMPTY[c])
                    const int ngath = field ? P.nch - 1 : P.nch;
                    for (int c = 0; c < P.nch; ++c) {
                        if (c == 0 || c == ngath) {
                            mbar_wait_warp(BAR(BAR_FULL + c), (uint32_t)(j & 1));
                            tc_fence_after();
                            SD_TRACE(1, j, c == 0 ? 0 : 4);
                        }
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma(d1, umma_desc(sm_u + OFF_RING + c * CHUNK_BYTES + k * 32),
                                     umma_desc(sm_u + OFF_W1 + c * CHUNK_BYTES + k * 32), idesc1, (c | k) != 0);
                            umma_commit(BAR(BAR_EMPTY + c));
                            if (c == P.nch - 1) umma_commit(BAR(BAR_D1 + (int)(j & 1)));
                        }
                        __syncwarp();
                    }
                    SD_TRACE(1, j, 5);
                    if (P.cmma && j > 1) composite(j - 2);     // its operands were written a tile ago
                    if (j > 0) layer2(j - 1);
                    SD_TRACE(1, j, 7);
                }
                __syncwarp();
                if (P.cmma && j > 1) cm_output(j - 2);
            }
            if (P.cmma && my_tiles > 1) composite(my_tiles - 2);
            if (my_tiles > 0) layer2(my_tiles - 1);
            __syncwarp();
            if (P.cmma && my_tiles > 1) cm_output(my_tiles - 2);
            if (P.cmma && my_tiles > 0) {
                composite(my_tiles - 1);
                __syncwarp();
                cm_output(my_tiles - 1);
            }
        }
    } else if (warp < WARP_GA0) {
        // =================================== POINT WARPS ================================================
        if (field) {
            const int row = tid - WARP_PT0 * 32;
            const int K = P.K, k_i = row % K;
            const bool row_used = row < P.upt * K;
            const int nv_c = P.fp.nv_c;
            // inputs of a tile row (global row index, 3-D point, sample depth), fetched one tile ahead so that the
            // dependent perm -> xyz (or ray + z) loads are off the critical path
            // (the loaded values are kept RAW -- origin, direction, depth -- and turned into the point when the tile is
            // processed: forming o + z d here made the loop wait for these loads right where they were issued)
            struct RowIn { long long grow; float px, py, pz, dx, dy, dz, zs; };
            auto fetch = [&](long long jj) {
                RowIn r;
                r.grow = -1; r.px = r.py = r.pz = r.dx = r.dy = r.dz = r.zs = 0.0f;
                if (jj >= my_tiles) return r;
                const long long unit = (first + jj * stride) * P.upt + row / K;
                if (!(row_used && unit < P.n_units)) return r;
                r.grow = P.perm ? (long long)__ldg(P.perm + unit) : unit * K + k_i;
                if (!P.src.xyz) {
                    const float *ry = P.src.rays + (r.grow / P.src.K) * P.src.r_dim;
                    r.zs = __ldg(P.src.z + r.grow);
                    r.px = __ldg(ry + 0); r.py = __ldg(ry + 1); r.pz = __ldg(ry + 2);
                    r.dx = __ldg(ry + 3); r.dy = __ldg(ry + 4); r.dz = __ldg(ry + 5);
                } else {
                    r.px = __ldg(P.src.xyz + 3 * r.grow); r.py = __ldg(P.src.xyz + 3 * r.grow + 1); r.pz = __ldg(P.src.xyz + 3 * r.grow + 2);
                }
                return r;
            };
            const bool from_rays = !P.src.xyz;
            RowIn nxt = fetch(0);
            for (long long j = 0; j < my_tiles; ++j) {
                const RowIn cur = nxt;
                nxt = fetch(j + 1);
                const bool ok = cur.grow >= 0;
                const long long grow = cur.grow;
                const int slot = (int)(j % NGEO);
                mbar_wait(BAR(BAR_GEO_EMPTY + slot), (uint32_t)(((j / NGEO) & 1) ^ 1));
                if (warp == WARP_PT0) SD_TRACE(2, j, 0);
                float x = 0.f, y = 0.f, zp = 0.f;
                const float zs = cur.zs;
                int flags = 0, off = 0;
                Tap t = {};
                const bool ring_col = render && P.cmma;              // colours go to the epilogue through shared memory
                float call[3 * MAX_NVC_TC];
#pragma unroll
                for (int c = 0; c < 3 * MAX_NVC_TC; ++c) call[c] = 0.0f;
                if (ok) {
                    const float px = from_rays ? ray_point(cur.px, cur.dx, zs) : cur.px, py = from_rays ? ray_point(cur.py, cur.dy, zs) : cur.py,
                                pz = from_rays ? ray_point(cur.pz, cur.dz, zs) : cur.pz;
                    float zc;
                    bool inv;
                    project_point(s_cam, s_cam + 9, px, py, pz, x, y, zc, inv);
                    x = clamp_keep_nan(x, -2.0f, 2.0f);
                    y = clamp_keep_nan(y, -2.0f, 2.0f);
                    zp = znorm(zc, P.fp.enc);
                    t = bilinear_tap(x, y, P.fp.Hf, P.fp.Wf);
                    // keep the 2x2 footprint inside the map so the gather can use fixed +1 texel / +1 row offsets:
                    // at the last column / row the out-of-range taps have weight zero, so shifting the base by one
                    // and moving the weights over is exact
                    clamp_footprint(t, P.fp.Hf, P.fp.Wf);
                    off = t.y0 * P.fp.Wf + t.x0;
                    flags = (inv ? 1 : 0) | 8;
                    if (P.invalid_feat) P.invalid_feat[grow] = inv ? 1 : 0;
                    const bool want_col = P.rgb || (ring_col && P.rgb_ray);
                    if (nv_c > 0 && (want_col || P.invalid)) {
#pragma unroll
                        for (int v = 0; v < MAX_NVC_TC; ++v) {
                            if (v >= nv_c) break;
                            float cx, cy, cz;
                            bool cinv;
                            const float *c = s_cam + 21 * (1 + v);
                            project_point(c, c + 9, px, py, pz, cx, cy, cz, cinv);
                            if (want_col) {
                                float c3[3];
                                sample_color(P.fp.rgb + (size_t)v * 3 * P.fp.Hc * P.fp.Wc, P.fp.Hc, P.fp.Wc, cx, cy, c3);
                                call[3 * v] = c3[0]; call[3 * v + 1] = c3[1]; call[3 * v + 2] = c3[2];
                                if (P.rgb) {
                                    float *o = P.rgb + (size_t)grow * 3 * nv_c + 3 * v;
                                    o[0] = c3[0]; o[1] = c3[1]; o[2] = c3[2];
                                }
                            }
                            if (P.invalid) P.invalid[(size_t)grow * nv_c + v] = (cinv || inv) ? 1.0f : 0.0f;
                        }
                    }
                }
                if (ring_col) {
                    uint4 *cr = reinterpret_cast<uint4 *>(sm + off_col + slot * COL_SLOT + row * 32);
                    cr[0] = make_uint4(pack_h2(call[0], call[1]), pack_h2(call[2], call[3]), pack_h2(call[4], call[5]), pack_h2(call[6], call[7]));
                    cr[1] = make_uint4(pack_h2(call[8], call[9]), pack_h2(call[10], call[11]), 0u, 0u);
                }
                Geo &g = s_geo[slot];
                {
                    const uint32_t tex = (uint32_t)P.fp.C * 2u;                       // bytes per texel row
                    const uint32_t o_nw = (uint32_t)off * tex;
                    const bool plain = ok && !(P.fp.learn_empty && (flags & 1));
                    g.o[row] = o_nw;
                    g.w01[row] = plain ? pack_h2(t.wnw, t.wne) : 0u;
                    g.w23[row] = plain ? pack_h2(t.wsw, t.wse) : 0u;
                    g.flags[row] = flags; g.z[row] = zs; g.grow[row] = (int)grow;
                }
                mbar_arrive_warp(BAR(BAR_GEO_FULL + slot));
                if (warp == WARP_PT0) SD_TRACE(2, j, 1);
                // ---- positional code -> last K chunk (positional_encoding.py:68-80; sin/cos of 1.5*2^k*v by
                //      angle doubling from one accurate sincosf per coordinate) ------------------------------
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) pk[i] = 0u;
                if (ok) {
                    float code[48];
                    code[0] = x; code[1] = y; code[2] = zp;
                    code[45] = 1.0f; code[46] = 1.0f; code[47] = 0.0f;   // constant-1 columns: layer-1 bias (hi, lo) comes out of the MMA
                    // hi/lo split of the raw coordinates into the K padding (see mlp_pack_kernel)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float hi = __half2float(__float2half_rn(code[d]));
                        code[39 + d] = code[d] - hi;
                        code[42 + d] = hi;
                    }
                    float s[3], c[3];
                    const float a0 = __fmul_rn(x, P.fp.enc.freq_factor), a1 = __fmul_rn(y, P.fp.enc.freq_factor), a2 = __fmul_rn(zp, P.fp.enc.freq_factor);
                    if (fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fabsf(a2)) <= 3.2f) {     // the usual case: hardware sin / cos (as field_bin.cu)
                        s[0] = __sinf(a0); c[0] = __cosf(a0); s[1] = __sinf(a1); c[1] = __cosf(a1); s[2] = __sinf(a2); c[2] = __cosf(a2);
                    } else {                                                          // next to / behind the camera: |z'| is large
                        sincosf(a0, &s[0], &c[0]); sincosf(a1, &s[1], &c[1]); sincosf(a2, &s[2], &c[2]);
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            code[3 + 6 * k + d] = s[d];
                            code[3 + 6 * k + 3 + d] = c[d];
                            const float s2 = 2.0f * s[d] * c[d], c2 = fmaf(-2.0f * s[d], s[d], 1.0f);
                            s[d] = s2; c[d] = c2;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 24; ++i) pk[i] = pack_h2(code[2 * i], code[2 * i + 1]);
                }
                const int cchunk = P.nch - 1;
                if (warp == WARP_PT0) SD_TRACE(2, j, 2);
                mbar_wait(BAR(BAR_EMPTY + cchunk), (uint32_t)((j & 1) ^ 1));
                if (warp == WARP_PT0) SD_TRACE(2, j, 3);
                unsigned char *arow = sm + OFF_RING + cchunk * CHUNK_BYTES + row * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<uint4 *>(arow + ((q ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                fence_proxy_async();
                mbar_arrive_warp(BAR(BAR_FULL + cchunk));
                if (warp == WARP_PT0) SD_TRACE(2, j, 4);
            }
        }
    } else {
        // =================================== GATHER WARPS ================================================
        const int gw = warp - WARP_GA0;                 // units u = 8*chunk + row group, u % N_GA_WARPS == gw
        const int sub = lane & 7, grp = lane >> 3;
        const int nfeat = field ? P.nch - 1 : P.nch;
        const int C = P.fp.C;
        const int n_units = nfeat * 8;                 // (chunk, 16-row group) units of a tile; mine: gw, gw+7, ...
        const uint32_t tex = (uint32_t)C * 2u, trow = (uint32_t)P.fp.Wf * tex;
        for (long long j = 0; j < my_tiles; ++j) {
            const long long tile = first + j * stride;
            const int slot = (int)(j % NGEO);
            if (field) {
                mbar_wait(BAR(BAR_GEO_FULL + slot), (uint32_t)((j / NGEO) & 1));
                if (gw == 0) SD_TRACE(3, j, 0);
                const Geo &g = s_geo[slot];
                const unsigned char *fmap = reinterpret_cast<const unsigned char *>(P.fp.feat) + sub * 16;
                const __half *s_empty = reinterpret_cast<const __half *>(sm + OFF_EMPTY) + sub * 8;
                // A unit is two halves of 8 rows (2 x 4 rows per request).  Software pipeline: the 8 LDG.128 of the
                // next half are issued before the current half is blended, so a load batch is always in flight
                // behind the HFMA2 work; loads do not need the ring stage, only the stores wait for it.
                uint4 rawA[2][4], rawB[2][4];
                uint2 wA[2], wB[2];
                auto issue = [&](int u, int half, uint4 (&raw)[2][4], uint2 (&w)[2]) {
                    const int c = u >> 3, rg = u & 7;
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int p = rg * 16 + (half * 2 + it) * 4 + grp;
                        const unsigned char *q = fmap + c * 128 + g.o[p];
                        w[it] = make_uint2(g.w01[p], g.w23[p]);
                        raw[it][0] = ldg128(q); raw[it][1] = ldg128(q + tex);
                        raw[it][2] = ldg128(q + trow); raw[it][3] = ldg128(q + trow + tex);
                    }
                };
                auto blend = [&](int u, int half, uint4 (&raw)[2][4], uint2 (&w)[2]) {
                    const int c = u >> 3, rg = u & 7;
                    unsigned char *stage = sm + OFF_RING + c * CHUNK_BYTES;
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int p = rg * 16 + (half * 2 + it) * 4 + grp;
                        uint32_t w0u = __byte_perm(w[it].x, 0, 0x1010), w1u = __byte_perm(w[it].x, 0, 0x3232);
                        const uint32_t w2u = __byte_perm(w[it].y, 0, 0x1010), w3u = __byte_perm(w[it].y, 0, 0x3232);
                        if (P.fp.learn_empty && (g.flags[p] & 9) == 9) {      // bts.py:311-319
                            raw[it][0] = *reinterpret_cast<const uint4 *>(s_empty + c * 64);
                            w0u = as_u32(__float2half2_rn(1.0f));
                        }
                        const __half2 w0 = as_h2(w0u), w1 = as_h2(w1u), w2 = as_h2(w2u), w3 = as_h2(w3u);
                        const uint32_t *a = reinterpret_cast<const uint32_t *>(&raw[it][0]);
                        const uint32_t *b = reinterpret_cast<const uint32_t *>(&raw[it][1]);
                        const uint32_t *cc = reinterpret_cast<const uint32_t *>(&raw[it][2]);
                        const uint32_t *d = reinterpret_cast<const uint32_t *>(&raw[it][3]);
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {   // two channels per HFMA2, taps in the reference order nw, ne, sw, se
                            __half2 acc = __hmul2(w0, as_h2(a[e]));
                            acc = __hfma2(w1, as_h2(b[e]), acc);
                            acc = __hfma2(w2, as_h2(cc[e]), acc);
                            acc = __hfma2(w3, as_h2(d[e]), acc);
                            pk[e] = as_u32(acc);
                        }
                        *reinterpret_cast<uint4 *>(stage + p * 128 + ((sub ^ (p & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                };
                int cur_c = -1;
                if (gw < n_units) issue(gw, 0, rawA, wA);
                for (int u = gw; u < n_units; u += N_GA_WARPS) {
                    const int c = u >> 3;
                    issue(u, 1, rawB, wB);
                    if (c != cur_c) {                       // first unit of mine in this chunk: the stage must be free
                        mbar_wait(BAR(BAR_EMPTY + c), (uint32_t)((j & 1) ^ 1));
                        cur_c = c;
                        if (gw == 0) SD_TRACE(3, j, 1 + c);
                    }
                    blend(u, 0, rawA, wA);
                    if (u + N_GA_WARPS < n_units) issue(u + N_GA_WARPS, 0, rawA, wA);
                    blend(u, 1, rawB, wB);
                }
                fence_proxy_async();
                mbar_arrive_warp(BAR(BAR_FULL + 0));
                mbar_arrive_warp(BAR(BAR_GEO_EMPTY + slot));
                if (gw == 0) SD_TRACE(3, j, 5);
            } else {
                // rows mode: x [N, d_in] fp32 -> fp16 chunk c (columns 64c .. 64c+63)
                for (int c = 0; c < nfeat; ++c) {
                    mbar_wait(BAR(BAR_EMPTY + c), (uint32_t)((j & 1) ^ 1));
                    unsigned char *stage = sm + OFF_RING + c * CHUNK_BYTES;
                    for (int rg = (gw + N_GA_WARPS - (8 * c) % N_GA_WARPS) % N_GA_WARPS; rg < 8; rg += N_GA_WARPS) {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int p = rg * 16 + it * 4 + grp;
                            const long long r = tile * TM + p;
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int col = c * 64 + sub * 8 + e;
                                f[e] = (r < P.n_units && col < P.d_in) ? __ldg(P.x_rows + r * P.d_in + col)
                                       : ((col == P.d_in + 6 || col == P.d_in + 7) ? 1.0f : 0.0f);   // bias columns
                            }
                            *reinterpret_cast<uint4 *>(stage + p * 128 + ((sub ^ (p & 7)) << 4)) =
                                make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
                        }
                    }
                }
                fence_proxy_async();
                mbar_arrive_warp(BAR(BAR_FULL + 0));
            }
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc

// debug: copies the clock64 trace of the last launch made with SD_TC_DEBUG & 8192 (not part of the public header)
extern "C" int sd_debug_read_trace(long long *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tc::g_trace, sizeof(long long) * 4 * 64 * 8));
    return SD_OK;
}

// ---- host side ------------------------------------------------------------------------------------------
constexpr int TC_MAX_D = 64;      // feature outputs the kernel's own layer 2 produces; above: hidden-composite render (any D)
static bool tc_head_ok(const sd_mlp *mlp, bool any_d = false) {
    return mlp && mlp->packed && mlp->d_hidden == 128 && mlp->d_out >= 2 && (any_d || mlp->d_out - 1 <= TC_MAX_D) && mlp->d_in >= 1 &&
           mlp->d_in + 8 <= 64 * tc::MAX_CHUNKS;
}

static bool tc_scene_ok(const sd_scene *s, const sd_mlp *mlp, bool any_d) {
    return s && (s->feat_dtype == SD_F16 || s->feat_proj) && s->C == 256 && s->Hf >= 2 && s->Wf >= 2 && s->nv_f == 1 && s->include_input &&
           s->num_freqs == 6 && s->nv_c <= tc::MAX_NVC_TC && tc_head_ok(mlp, any_d) && mlp->d_in == s->C + 39;
}

// Render mode.  A head with more than 64 feature outputs (the 768-d DINO variant) needs a projected scene: the composite
// then sums the 128 hidden units per ray and the feature rows of W_out are applied to the sums (tc_render_mode() == 2).
bool tc_supported(const sd_scene *scene, const sd_mlp *mlp, int K) {
    const bool big = mlp && mlp->d_out - 1 > TC_MAX_D;
    return tc_scene_ok(scene, mlp, big && scene && scene->feat_proj) && K >= 32 && K <= tc::TM;
}

// 0: composite in the epilogue (shuffles; unprojected scene), 1: tensor-core composite of the features, 2: of the hidden units
int tc_render_mode(const sd_scene *scene, const sd_mlp *mlp) {
    if (!scene || !mlp || !scene->feat_proj) return 0;
    if (mlp->d_out - 1 > TC_MAX_D) return 2;
    static const int forced = [] { const char *e = getenv("SD_TC_HCOMP"); return e ? atoi(e) : -1; }();
    return forced == 1 ? 2 : 1;
}

static int tc_launch(tc::Params &P, const sd_mlp *mlp, cudaStream_t st, const void *proj = nullptr) {
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    SD_REQUIRE(((uintptr_t)blob & 15) == 0, "mlp: packed blob must be 16-byte aligned");
    P.w1_img = blob + L.off_w_in_h;
    P.nch = L.d_in_pad / 64;
    if (proj) {   // projected scene: [identity (2 chunks) | code block] instead of W_in, 128 projected channels per texel
        const unsigned char *pb = reinterpret_cast<const unsigned char *>(proj);
        SD_REQUIRE(((uintptr_t)pb & 15) == 0, "projected scene: blob must be 16-byte aligned");
        P.w1_img = pb + PROJ_OFF_IDENT;        // identity chunks and the code chunk are contiguous
        P.empty_h = reinterpret_cast<const __half *>(pb + PROJ_OFF_EMPTY);
        P.fp.feat = pb + PROJ_OFF_MAP;
        P.fp.feat_f16 = 1;
        P.fp.C = 128;
        P.nch = 3;
    }
    P.w2_img = blob + L.off_w_out_h;
    P.b_out = reinterpret_cast<const float *>(blob + L.off_b_out);
    P.dbg = debug_mask();
    P.D = mlp->d_out - 1;
    if (!(proj && P.mode == tc::MODE_RENDER)) P.cmma = 0;
    P.n2 = (mlp->d_out + 15) / 16 * 16;
    P.sig_col = P.D;
    if (P.cmma == 2) {     // layer 2 = the density row alone
        SD_REQUIRE(L.off_w_sig_h != 0 && P.hsum && P.wsum, "hidden-composite render: needs the density image and the per-ray scratch");
        P.w2_img = blob + L.off_w_sig_h;
        P.n2 = 16;
        P.sig_col = 0;
    } else {
        SD_REQUIRE(P.D <= TC_MAX_D, "SD_MLP_F16_TC: more than %d feature outputs need the projected render path", TC_MAX_D);
    }
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(tc::field_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC));
    }
    const unsigned grid = (unsigned)(P.n_tiles < sm_count ? P.n_tiles : sm_count);
    profile_before(st);
    tc::field_tc_kernel<<<grid, tc::NTHREADS, tc::SMEM_ALLOC, st>>>(P);
    profile_after(st);
    SD_LAUNCH_OK("field_tc_kernel");
    return SD_OK;
}

int launch_field_tc(const FieldParams &fp, const PointSrc &src, long long N, const sd_mlp *mlp, const TcRender *render,
                    const TcOut &out, cudaStream_t st, const unsigned int *perm, const void *proj) {
    if (N == 0) return SD_OK;
    SD_REQUIRE(N < (1ll << 31), "SD_MLP_F16_TC: at most 2^31 - 1 rows per call (got %lld)", N);
    SD_REQUIRE(tc_head_ok(mlp, render && render->cmma == 2),
               "SD_MLP_F16_TC: head must be d_in <= 312, d_hidden = 128, 2 <= d_out <= 65 (any d_out for renders of a projected scene) and packed");
    SD_REQUIRE(fp.feat_f16 || proj, "SD_MLP_F16_TC: the feature map must be packed as fp16 (sd_featmap_pack with SD_F16)");
    SD_REQUIRE(fp.C == 256 && fp.code_dim == 39 && fp.enc.include_input && mlp->d_in == fp.C + fp.code_dim,
               "SD_MLP_F16_TC: supports C = 256 with the 39-d positional code (got C=%d, code=%d, d_in=%d)", fp.C,
               fp.code_dim, mlp->d_in);
    SD_REQUIRE(fp.nv_c <= tc::MAX_NVC_TC, "SD_MLP_F16_TC: at most %d colour views (got %d)", tc::MAX_NVC_TC, fp.nv_c);
    SD_REQUIRE(fp.Hf >= 2 && fp.Wf >= 2, "SD_MLP_F16_TC: the feature map must be at least 2 x 2");
    tc::Params P = {};
    P.fp = fp;
    P.src = src;
    if (render) {
        const int K = src.K;
        SD_REQUIRE(K >= 32 && K <= tc::TM, "SD_MLP_F16_TC: fused render needs 32 <= K <= 128 samples per ray (got %d)", K);
        P.mode = tc::MODE_RENDER;
        P.K = K;
        P.upt = tc::TM / K;
        P.n_units = N / K;
        P.cfg = render->cfg;
        P.depth = render->depth; P.dino_ray = render->dino; P.rgb_ray = render->rgb_out;
        P.weights = render->weights; P.alphas = render->alphas;
        P.rgb = render->rgb_samps;   // per-sample colours: the caller's buffer (or, without the tensor-core composite, workspace)
        P.cmma = proj ? render->cmma : 0;
        P.hsum = render->hsum; P.wsum = render->wsum;
        SD_REQUIRE(fp.nv_c == 0 || !P.rgb_ray || P.rgb || P.cmma, "SD_MLP_F16_TC: rgb_out needs a per-sample colour buffer");
    } else {
        P.mode = tc::MODE_POINTS;
        P.K = 1;
        P.upt = tc::TM;
        P.n_units = N;
        P.dino = out.dino;
        P.rgb = out.rgb;
        P.perm = src.xyz ? perm : nullptr;
    }
    P.n_tiles = (P.n_units + P.upt - 1) / P.upt;
    P.sigma = out.sigma; P.invalid = out.invalid; P.invalid_feat = out.invalid_feat;
    return tc_launch(P, mlp, st, proj);
}

int launch_mlp_tc(const sd_mlp *mlp, const float *x, long long N, float *out, cudaStream_t st) {
    SD_REQUIRE(tc_head_ok(mlp), "SD_MLP_F16_TC: head must be d_in <= 312, d_hidden = 128, 2 <= d_out <= 65 and packed");
    if (N == 0) return SD_OK;
    SD_REQUIRE(x && out, "sd_mlp_forward: null pointer");
    tc::Params P = {};
    P.mode = tc::MODE_ROWS;
    P.K = 1; P.upt = tc::TM; P.n_units = N;
    P.n_tiles = (N + tc::TM - 1) / tc::TM;
    P.x_rows = x; P.d_in = mlp->d_in; P.out_rows = out;
    return tc_launch(P, mlp, st);
}

}  // namespace sd
