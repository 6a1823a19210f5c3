// Tile kernel of the field query on a projected scene (sd_field_project): bilinear interpolation ON the tensor cores.
//
//   BTSNet.forward      models/bts.py:476-595     (projection, mask, gather, code, head, softplus, colours)
//
// The points arrive sorted by the 7x7-texel bin of the feature map their 2x2 footprint starts in (binning.cu), so
// every footprint of a bin lies inside one 8x8-texel box = 64 texels.  For a tile of 128 sorted points and each bin
// ("chunk") it touches,
//     hidden_pre[128 x 128] += Wgt[128 x 64] . Pbox[64 x 128]
// where Pbox is the box of the PROJECTED map (P = W_feat . F, field_proj.cu), fetched by ONE pair of TMA tile copies
// straight into the MN-major SWIZZLE_128B layout tcgen05.mma reads, and Wgt holds the four bilinear weights of each
// row (zero elsewhere).  The gather -- 4 taps x 512 B per point through the load/store unit in field_tc.cu, the
// wall of that kernel -- is gone: a point costs four 2-byte shared-memory stores.  The positional code, the
// coordinate hi/lo split, the bias and the learn_empty replacement go through a 48-wide K block as before;
// layer 2 reads the ReLU'd hidden tile from TMEM, written in place over the layer-1 accumulator.
//
// Warp roles (10 warps, one persistent CTA per SM):
//   warps 0-3  epilogue: layer-1 accumulator -> ReLU -> fp16 -> same TMEM columns (A operand of layer 2);
//              layer-2 accumulator -> softplus density + features, staged through shared memory, coalesced stores
//   warp  4    tcgen05.mma issuer + TMEM owner
//   warp  5    TMA producer: one 8x8x128-channel box of P per chunk
//   warps 6-9  one thread per row: point -> projection, mask, tap, colours, weights into the chunk's A operand
//              (an undo log keeps the rest of the operand zero), positional code -> code operand
// Rings: NRING (A chunk, B chunk) pairs released by tcgen05.commit, 2 code operands, layer-1 and layer-2
// accumulators double buffered in TMEM.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {
namespace tb {
using namespace tcx;

constexpr int TM = 128;
constexpr int CHUNK = 16384;                 // A chunk [128 rows][64 slots] fp16 = B chunk [2 halves][64 slots][64 ch] fp16
#ifndef SD_TB_PT_GROUPS
#define SD_TB_PT_GROUPS 1
#endif
constexpr int NRING = 3, NCODE = 2;
constexpr int N_EPI_WARPS = 4, N_PT_WARPS = 4, N_PT_GROUPS = SD_TB_PT_GROUPS;
constexpr int WARP_EPI2 = N_EPI_WARPS, WARP_MMA = 2 * N_EPI_WARPS, WARP_TMA = WARP_MMA + 1, WARP_PT0 = WARP_TMA + 1;
constexpr int NTHREADS = (WARP_PT0 + N_PT_GROUPS * N_PT_WARPS) * 32;
static_assert(NCODE % N_PT_GROUPS == 0, "a code operand is always filled by the same point group");
constexpr int TMEM_COLS = 512;
constexpr int D2_COL = 256, D2_STRIDE = 128;  // layer-1 accumulators at columns 0 / 128, layer-2 at 256 / 384
constexpr int MAX_NVC_TB = 4;
constexpr int W2_BYTES = 3 * 80 * 128;     // W_out: two K blocks + the bias block
constexpr int ONE_COL = 496;                 // 8 TMEM columns holding the constant (1, 1, 0, ...): A operand of the bias K step
constexpr int KCODE = 3;                     // K steps of the code block (48 columns)

constexpr int OFF_WC = 0;
constexpr int OFF_A = OFF_WC + CHUNK;
constexpr int OFF_B = OFF_A + NRING * CHUNK;
constexpr int OFF_CODE = OFF_B + NRING * CHUNK;
constexpr int OFF_W2 = OFF_CODE + NCODE * CHUNK;
constexpr int OFF_STAGE = OFF_W2 + W2_BYTES;
constexpr int STAGE_ROW = 272;               // output row (256 B) + 16 B: thread = row stores are bank-conflict free
constexpr int OFF_BO = OFF_STAGE + TM * STAGE_ROW;   // output bias of the 64 feature columns
constexpr int OFF_DIRTY = OFF_BO + 256;
constexpr int OFF_CAM = OFF_DIRTY + NRING * TM;
constexpr int OFF_BAR = OFF_CAM + 448;
enum { BAR_FULL_A = 0, BAR_FULL_B = NRING, BAR_EMPTY = 2 * NRING, BAR_FULL_C = 3 * NRING, BAR_EMPTY_C = BAR_FULL_C + NCODE,
       BAR_D1 = BAR_EMPTY_C + NCODE, BAR_H = BAR_D1 + 2, BAR_D2 = BAR_H + 2, BAR_D2_EMPTY = BAR_D2 + 2,
       BAR_WLOAD = BAR_D2_EMPTY + 2, NBAR = BAR_WLOAD + 1 };
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_ALLOC = OFF_TMEM + 16 + 1024;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(OFF_W2 % 1024 == 0 && OFF_STAGE % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

__device__ long long g_trace[8 * 64 * 8];   // [role][tile][event] clock64 stamps of CTA 0 (SD_TC_DEBUG & 8192)
#define TB_TRACE(role, j, ev)                                                                    \
    do {                                                                                         \
        if ((P.dbg & 8192) && blockIdx.x == 0 && (j) < 64 && (role) < 8 && (threadIdx.x & 31) == 0)        \
            g_trace[((role) * 64 + (int)(j)) * 8 + (ev)] = clock64();                            \
    } while (0)

struct Params {
    CUtensorMap tmap;          // P as [Hf][Wf][128] fp16, box 8 x 8 x 64 channels, SWIZZLE_128B
    FieldParams fp;
    const float *xyz;
    const unsigned int *perm;
    const unsigned short *pcb;
    const unsigned int *cbin;
    int nbx;
    int dbg;                   // SD_TC_DEBUG & 8192: clock64 trace of CTA 0
    long long N, n_tiles;
    int n2, D;
    const unsigned char *wc_img, *w2_img;
    const float *b_out;
    float *sigma, *dino, *rgb, *invalid;
    unsigned char *invalid_feat;
};

__global__ void __launch_bounds__(NTHREADS, 1) field_bin_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = sm_u + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    float *s_cam = reinterpret_cast<float *>(sm + OFF_CAM);
    unsigned char *s_dirty = sm + OFF_DIRTY;

    // ---- one-time setup ------------------------------------------------------------------------------
    if (tid == 0) {
        for (int e = 0; e < NRING; ++e) {
            mbar_init(BAR(BAR_FULL_A + e), N_PT_WARPS);
            mbar_init(BAR(BAR_FULL_B + e), 1);
            mbar_init(BAR(BAR_EMPTY + e), 1);
        }
        for (int s = 0; s < NCODE; ++s) { mbar_init(BAR(BAR_FULL_C + s), N_PT_WARPS); mbar_init(BAR(BAR_EMPTY_C + s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(BAR_D1 + b), 1); mbar_init(BAR(BAR_H + b), N_EPI_WARPS);
            mbar_init(BAR(BAR_D2 + b), 1); mbar_init(BAR(BAR_D2_EMPTY + b), N_EPI_WARPS);
        }
        mbar_init(BAR(BAR_WLOAD), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 21 * (1 + P.fp.nv_c); i += NTHREADS) {
        const int c = i / 21, e = i - 21 * c;
        const float *K = c == 0 ? P.fp.K_f : P.fp.K_c + 9 * (c - 1);
        const float *W = c == 0 ? P.fp.w2c_f : P.fp.w2c_c + 16 * (c - 1);
        s_cam[i] = e < 9 ? __ldg(K + e) : __ldg(W + (e - 9));
    }
    // the weight operands start out all zero and are kept so by the undo log of the point warps
    for (int i = tid; i < NRING * CHUNK / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm + OFF_A)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < NRING * TM; i += NTHREADS) s_dirty[i] = 0xFF;
    for (int i = tid; i < 64; i += NTHREADS) reinterpret_cast<float *>(sm + OFF_BO)[i] = i < P.D ? __ldg(P.b_out + 1 + i) : 0.0f;
    fence_proxy_async();
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t w2b = 3u * (uint32_t)P.n2 * 128u;
        mbar_expect_tx(BAR(BAR_WLOAD), CHUNK + w2b);
        bulk_g2s(sm_u + OFF_WC, P.wc_img, CHUNK, BAR(BAR_WLOAD));
        bulk_g2s(sm_u + OFF_W2, P.w2_img, w2b, BAR(BAR_WLOAD));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);

    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_tiles = P.n_tiles > first ? (P.n_tiles - first + stride - 1) / stride : 0;
    // compact bins touched by my tile number jj: the first and the last one (the sorted order makes them consecutive).
    // Raw loads only: callers do the arithmetic (count = last - first + 1) an iteration later, so that nobody
    // waits for a load in the iteration that issues it.
    auto span = [&](long long jj, int &c0, int &c1) {
        c0 = 0; c1 = 0;
        if (jj < 0 || jj >= my_tiles) return;
        const long long a = (first + jj * stride) * TM;
        const long long b = (a + TM < P.N ? a + TM : P.N) - 1;
        c0 = (int)__ldg(P.pcb + a);
        c1 = (int)__ldg(P.pcb + b);
    };

    if (warp < N_EPI_WARPS) {
        // =================================== EPILOGUE 1 ==============================================
        // layer-1 accumulator -> ReLU -> fp16 -> the same TMEM columns, A operand of layer 2
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        {   // constant-1 columns (k = 128, 129 of layer 2): the output bias comes out of the MMA
            uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(t_lane + ONE_COL, one);
        }
        for (long long j = 0; j < my_tiles; ++j) {
            const int b = (int)(j & 1);
            mbar_wait(BAR(BAR_D1 + b), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (warp == 0) TB_TRACE(0, j, 2);
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {      // 32 hidden units -> 16 packed columns, written over columns already read
                uint32_t vr[32];
                tmem_ld32_issue(t_lane + b * 128 + kb * 32, vr);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    pk[e] = pack_h2_relu(__uint_as_float(vr[2 * e]), __uint_as_float(vr[2 * e + 1]));
                tmem_st16(t_lane + b * 128 + kb * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(BAR(BAR_H + b));
            if (warp == 0) TB_TRACE(0, j, 3);
        }
    } else if (warp < WARP_MMA) {
        // =================================== EPILOGUE 2 ==============================================
        // layer-2 accumulator -> softplus density + features -> global
        const int wq = warp - WARP_EPI2;                         // TMEM lane quadrant of this warp
        const int row = wq * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
        const int D = P.D;
        unsigned char *stage0 = sm + OFF_STAGE + wq * 8192, *stage1 = stage0 + 4096;
        auto load_grow = [&](long long jj) {
            if (jj >= my_tiles) return -1;
            const long long gpos = (first + jj * stride) * TM + row;
            return gpos < P.N ? (int)__ldg(P.perm + gpos) : -1;
        };
        int grow_next = load_grow(0);
        for (long long j = 0; j < my_tiles; ++j) {
            const int grow_keep = grow_next;                     // point this thread's row of tile j stands for
            grow_next = load_grow(j + 1);
            const int b1 = (int)(j & 1);
            mbar_wait(BAR(BAR_D2 + b1), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (wq == 0) TB_TRACE(0, j, 0);
            const uint32_t t_d2 = t_lane + D2_COL + b1 * D2_STRIDE;
            const bool ok = grow_keep >= 0;
            if (P.dino && D == 64) {
                // Two halves of 32 columns (registers).  Transpose through shared memory: lane = row writes its 8 chunks
                // of a half (XOR-swizzled, conflict free), then 8 lanes read one row back and the warp stores four
                // 128-byte row halves per request.  (One 256-byte bulk copy per row was measured instead: ~22 cycles
                // per copy, slower.)
                uint32_t sr;
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t vr[32];
                    tmem_ld32_issue(t_d2 + hf * 32, vr);
                    if (hf == 1) tmem_ld1_issue(t_d2 + D, sr);                   // density column sits behind the features
                    tmem_ld_wait();
                    unsigned char *stage = hf ? stage1 : stage0;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<uint4 *>(stage + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                            make_uint4(vr[4 * q], vr[4 * q + 1], vr[4 * q + 2], vr[4 * q + 3]);
                }
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_D2_EMPTY + b1));            // (__syncwarp inside: the staged rows are visible)
                if (wq == 0) TB_TRACE(0, j, 4);
                if (ok && P.sigma && !(P.dbg & 2)) P.sigma[grow_keep] = softplus_fast(__uint_as_float(sr));
                if (wq == 0) TB_TRACE(0, j, 5);
                const int c8 = lane & 7;
                // all the shared loads of a batch first, into distinct registers (volatile asm keeps the order): a load
                // that reuses the source registers of a store in flight waits for that store to leave the LSU
#pragma unroll 1
                for (int hb = 0; hb < 2; ++hb) {       // batch = 8 requests of 4 row halves = one 32-column half
                    float4 x[8];
                    float *dp[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int hf = hb, r = 4 * i + (lane >> 3);
                        x[i] = lds128_ordered(smem_u32(hf ? stage1 : stage0) + r * 128 + ((c8 ^ (r & 7)) << 4));
                        const int dst = __shfl_sync(0xffffffffu, grow_keep, r);
                        dp[i] = dst >= 0 ? P.dino + (long long)dst * 64 + hf * 32 + c8 * 4 : nullptr;
                    }
                    __syncwarp();   // (also keeps ptxas from sinking the loads back between the stores)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (dp[i] && !(P.dbg & 1)) stg128_ordered(dp[i], x[i]);
                }
                __syncwarp();
            } else {
                uint32_t vr[64], sr;
                tmem_ld32_issue(t_d2, vr);
                tmem_ld32_issue(t_d2 + 32, vr + 32);
                tmem_ld1_issue(t_d2 + D, sr);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_D2_EMPTY + b1));
                if (ok && P.sigma) P.sigma[grow_keep] = softplus_fast(__uint_as_float(sr));
                if (P.dino && ok) {
                    float *o = P.dino + (long long)grow_keep * D;
#pragma unroll
                    for (int c = 0; c < 64; ++c)
                        if (c < D) o[c] = __uint_as_float(vr[c]);
                }
            }
            if (wq == 0) TB_TRACE(0, j, 1);
        }
    } else if (warp == WARP_MMA) {
        // =================================== MMA ISSUER ===============================================
        if (lane == 0) {
            mbar_wait(BAR(BAR_WLOAD), 0);
            const uint32_t idesc_k = umma_idesc(TM, 128), idesc_mn = idesc_k | UMMA_B_MN_MAJOR, idesc2 = umma_idesc(TM, P.n2);
            auto layer2 = [&](long long jj) {
                const int b = (int)(jj & 1);
                mbar_wait(BAR(BAR_H + b), (uint32_t)((jj >> 1) & 1));
                mbar_wait(BAR(BAR_D2_EMPTY + b), (uint32_t)(((jj >> 1) & 1) ^ 1));
                tc_fence_after();
                TB_TRACE(1, jj + 1, 5);
#pragma unroll
                for (int k = 0; k < 8; ++k)          // K = 16 per instruction = 8 packed columns of the hidden tile
                    umma_ts(tmem_base + D2_COL + b * D2_STRIDE, tmem_base + b * 128 + k * 8,
                            umma_desc(sm_u + OFF_W2 + (k >> 2) * P.n2 * 128 + (k & 3) * 32), idesc2, k != 0);
                umma_ts(tmem_base + D2_COL + b * D2_STRIDE, tmem_base + ONE_COL, umma_desc(sm_u + OFF_W2 + 2 * P.n2 * 128), idesc2, 1);
                umma_commit(BAR(BAR_D2 + b));
            };
            int e = 0;
            uint32_t ph = 0;
            int c0n, c1n;
            span(0, c0n, c1n);
            for (long long j = 0; j < my_tiles; ++j) {
                const int m = c1n - c0n + 1;
                span(j + 1, c0n, c1n);
                if (j > 0) layer2(j - 1);
                const uint32_t d1 = tmem_base + (uint32_t)(j & 1) * 128u;
                uint32_t acc = 0;
                TB_TRACE(1, j, 0);
                for (int i = 0; i < m; ++i) {
                    mbar_wait(BAR(BAR_FULL_B + e), ph);
                    if (i == 0) TB_TRACE(1, j, 1);
                    mbar_wait(BAR(BAR_FULL_A + e), ph);
                    tc_fence_after();
                    if (i == 0) TB_TRACE(1, j, 2);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {    // 16 texel slots per instruction
                        umma(d1, umma_desc(sm_u + OFF_A + e * CHUNK + k * 32),
                             umma_desc_mn(sm_u + OFF_B + e * CHUNK + k * 2048, CHUNK / 2, 1024), idesc_mn, acc);
                        acc = 1;
                    }
                    umma_commit(BAR(BAR_EMPTY + e));
                    if (++e == NRING) { e = 0; ph ^= 1; }
                }
                const int cs = (int)(j % NCODE);
                TB_TRACE(1, j, 3);
                mbar_wait(BAR(BAR_FULL_C + cs), (uint32_t)((j / NCODE) & 1));
                tc_fence_after();
                TB_TRACE(1, j, 4);
#pragma unroll
                for (int k = 0; k < KCODE; ++k)
                    umma(d1, umma_desc(sm_u + OFF_CODE + cs * CHUNK + k * 32), umma_desc(sm_u + OFF_WC + k * 32), idesc_k, 1);
                umma_commit(BAR(BAR_EMPTY_C + cs));
                umma_commit(BAR(BAR_D1 + (int)(j & 1)));
                TB_TRACE(1, j, 6);
            }
            if (my_tiles > 0) layer2(my_tiles - 1);
        }
    } else if (warp == WARP_TMA) {
        // =================================== TMA PRODUCER =============================================
        if (lane == 0) tma_prefetch_desc(&P.tmap);
        int e = 0;
        uint32_t ph = 0;
        int c0, c1, c0n, c1n;
        span(0, c0, c1);
        span(1, c0n, c1n);
        unsigned int mybin = lane <= c1 - c0 ? __ldg(P.cbin + c0 + lane) : 0u;
        for (long long j = 0; j < my_tiles; ++j) {
            // next tile's bins and the span of the tile after it are in flight while this tile's boxes are issued
            const int m = c1 - c0 + 1;
            const unsigned int mybin_n = lane <= c1n - c0n ? __ldg(P.cbin + c0n + lane) : 0u;
            int c0nn, c1nn;
            span(j + 2, c0nn, c1nn);
            for (int base = 0; base < m; base += 32) {
                if (base > 0) mybin = base + lane < m ? __ldg(P.cbin + c0 + base + lane) : 0u;
                const int cnt = m - base < 32 ? m - base : 32;
                for (int i = 0; i < cnt; ++i) {
                    const unsigned int b = __shfl_sync(0xffffffffu, mybin, i);
                    if (lane == 0) {
                        const int by = (int)(b / (unsigned)P.nbx), bx = (int)(b - (unsigned)by * (unsigned)P.nbx);
                        if (i == 0 && base == 0) TB_TRACE(6, j, 0);
                        mbar_wait(BAR(BAR_EMPTY + e), ph ^ 1);
                        if (i == 0 && base == 0) TB_TRACE(6, j, 1);
                        mbar_expect_tx(BAR(BAR_FULL_B + e), CHUNK);
                        const uint32_t dst = sm_u + OFF_B + e * CHUNK;
                        tma_load_3d(dst, &P.tmap, 0, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                        tma_load_3d(dst + CHUNK / 2, &P.tmap, 64, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                    }
                    if (++e == NRING) { e = 0; ph ^= 1; }
                }
                __syncwarp();
            }
            TB_TRACE(6, j, 2);
            if (lane == 0 && (P.dbg & 8192) && blockIdx.x == 0 && j < 64) g_trace[(6 * 64 + (int)j) * 8 + 7] = m;
            c0 = c0n; c1 = c1n; mybin = mybin_n;
            c0n = c0nn; c1n = c1nn;
        }
    } else {
        // =================================== POINT WARPS ================================================
        // Two groups of four warps take alternate tiles (group g: tiles g, g+2, ...), one thread per row.  Both count
        // the chunks of ALL tiles, so that the ring positions agree with the MMA issuer's.
        const int grp = (warp - WARP_PT0) / N_PT_WARPS;
        const int row = tid - (WARP_PT0 + grp * N_PT_WARPS) * 32;
        const int nv_c = P.fp.nv_c;
        const float inv_denom = 1.0f / P.fp.enc.denom;
        const int pt_role = 2 + (warp - WARP_PT0) % N_PT_WARPS + (grp ? 8 : 0);   // trace: group 0 only (slots 2..5)
        // inputs of a tile row, fetched in two stages so that no load is consumed in the iteration that issues it:
        // sorted position -> (point index, compact bin) two of my tiles ahead, point index -> coordinates one ahead
        struct RowIdx { int grow, cr, c0, c1, p0, p1; };
        struct RowIn { int grow; float px, py, pz; int c0, c1, cr, p0, p1; };
        auto fetch_idx = [&](long long jj) {
            RowIdx r;
            r.grow = -1; r.cr = 0;
            span(jj, r.c0, r.c1);
            span(jj - 1, r.p0, r.p1);          // the other group's tile in between (jj = 0: nothing, see below)
            if (jj >= my_tiles) return r;
            const long long gpos = (first + jj * stride) * TM + row;
            if (gpos >= P.N) return r;
            r.grow = (int)__ldg(P.perm + gpos);
            r.cr = (int)__ldg(P.pcb + gpos);
            return r;
        };
        auto fetch_pt = [&](const RowIdx &ix) {
            RowIn r;
            r.grow = ix.grow; r.cr = ix.cr; r.c0 = ix.c0; r.c1 = ix.c1; r.p0 = ix.p0; r.p1 = ix.p1;
            r.px = r.py = r.pz = 0.0f;
            if (ix.grow >= 0) {
                r.px = __ldg(P.xyz + 3ll * ix.grow); r.py = __ldg(P.xyz + 3ll * ix.grow + 1); r.pz = __ldg(P.xyz + 3ll * ix.grow + 2);
            }
            return r;
        };
        int e = 0;
        uint32_t ph = 0;
        RowIdx ix1 = fetch_idx(grp);
        RowIn nxt = fetch_pt(ix1);
        ix1 = fetch_idx(grp + N_PT_GROUPS);
        for (long long j = grp; j < my_tiles; j += N_PT_GROUPS) {
            const RowIn cur = nxt;
            TB_TRACE(pt_role, j, 0);
            nxt = fetch_pt(ix1);
            ix1 = fetch_idx(j + 2 * N_PT_GROUPS);
            if (N_PT_GROUPS == 2 && j > 0) {
                // Walk over the ring positions of the other group's tile j-1, WAITING on each: an mbarrier wait tells
                // apart only adjacent phases, so nobody may get two phases ahead on an entry (a tile can span more
                // chunks than the ring has entries).
                for (int i = cur.p1 - cur.p0 + 1; i > 0; --i) {
                    mbar_wait(BAR(BAR_EMPTY + e), ph ^ 1);
                    if (++e == NRING) { e = 0; ph ^= 1; }
                }
            }
            const bool ok = cur.grow >= 0;
            const long long grow = cur.grow;
            float x = 0.f, y = 0.f, zp = 0.f;
            bool inv = false;
            Tap t = {};
            if (ok) {
                float zc;
                project_point(s_cam, s_cam + 9, cur.px, cur.py, cur.pz, x, y, zc, inv);
                x = clamp_keep_nan(x, -2.0f, 2.0f);
                y = clamp_keep_nan(y, -2.0f, 2.0f);
                zp = znorm_fast(zc, P.fp.enc, inv_denom);
                t = bilinear_tap(x, y, P.fp.Hf, P.fp.Wf);
                clamp_footprint(t, P.fp.Hf, P.fp.Wf);
            }
            // ---- bilinear weights -> this row's 4 slots of its bin's chunk; every other slot of the row stays zero.
            //      Slots (nw, ne) = (s, s+1) share one 16-byte piece of the row (lx <= 6), (sw, se) the next one:
            //      the dirty byte keeps (ly << 3 | lx) and the undo clears the same two pairs.
            const bool plain = ok && !(P.fp.learn_empty && inv);
            const int bx7 = (int)(((unsigned)t.x0 * 9363u) >> 16), by7 = (int)(((unsigned)t.y0 * 9363u) >> 16);   // / 7 for < 2^15
            const int lx = t.x0 - bx7 * SD_BIN, ly = t.y0 - by7 * SD_BIN;
            const int q = cur.cr - cur.c0;
            const uint32_t w_top = pack_h2(t.wnw, t.wne), w_bot = pack_h2(t.wsw, t.wse);
            auto pair_off = [&](int lyy, int lxx) {      // byte offset of slot (lyy, lxx) inside the row
                return (uint32_t)(((lyy ^ (row & 7)) << 4) + lxx * 2);
            };
            TB_TRACE(pt_role, j, 1);
            const int m = cur.c1 - cur.c0 + 1;
            for (int i = 0; i < m; ++i) {
                mbar_wait(BAR(BAR_EMPTY + e), ph ^ 1);
                if (i == 0) TB_TRACE(pt_role, j, 2);
                unsigned char *arow = sm + OFF_A + e * CHUNK + row * 128;
                const int d = s_dirty[e * TM + row];
                if (d != 0xFF) {
                    unsigned short *p0 = reinterpret_cast<unsigned short *>(arow + pair_off(d >> 3, d & 7));
                    unsigned short *p1 = reinterpret_cast<unsigned short *>(arow + pair_off((d >> 3) + 1, d & 7));
                    p0[0] = 0; p0[1] = 0; p1[0] = 0; p1[1] = 0;
                }
                if (plain && i == q) {
                    unsigned short *p0 = reinterpret_cast<unsigned short *>(arow + pair_off(ly, lx));
                    unsigned short *p1 = reinterpret_cast<unsigned short *>(arow + pair_off(ly + 1, lx));
                    p0[0] = (unsigned short)w_top; p0[1] = (unsigned short)(w_top >> 16);
                    p1[0] = (unsigned short)w_bot; p1[1] = (unsigned short)(w_bot >> 16);
                    s_dirty[e * TM + row] = (unsigned char)(ly * 8 + lx);
                } else if (d != 0xFF) {
                    s_dirty[e * TM + row] = 0xFF;
                }
                fence_proxy_async();
                mbar_arrive_warp(BAR(BAR_FULL_A + e));
                if (++e == NRING) { e = 0; ph ^= 1; }
            }
            TB_TRACE(pt_role, j, 3);
            // ---- per-point outputs that do not need the head ---------------------------------------------------
            if (ok) {
                if (P.invalid_feat && !(P.dbg & 2)) P.invalid_feat[grow] = inv ? 1 : 0;
                if (nv_c > 0 && (P.rgb || P.invalid)) {
                    for (int v = 0; v < nv_c; ++v) {
                        float cx, cy, cz;
                        bool cinv;
                        const float *c = s_cam + 21 * (1 + v);
                        project_point(c, c + 9, cur.px, cur.py, cur.pz, cx, cy, cz, cinv);
                        if (P.rgb) {
                            float c3[3];
                            sample_color(P.fp.rgb + (size_t)v * 3 * P.fp.Hc * P.fp.Wc, P.fp.Hc, P.fp.Wc, cx, cy, c3);
                            float *o = P.rgb + (size_t)grow * 3 * nv_c + 3 * v;
                            o[0] = c3[0]; o[1] = c3[1]; o[2] = c3[2];
                        }
                        if (P.invalid) P.invalid[(size_t)grow * nv_c + v] = (cinv || inv) ? 1.0f : 0.0f;
                    }
                }
            }
            // ---- positional code -> code operand (positional_encoding.py:68-80; sin/cos of 1.5*2^k*v by angle
            //      doubling from one accurate sincosf per coordinate) -------------------------------------------
            uint32_t pk[24];
#pragma unroll
            for (int i = 0; i < 24; ++i) pk[i] = 0u;
            if (ok) {
                float code[48];
                code[0] = x; code[1] = y; code[2] = zp;
                code[45] = 1.0f; code[46] = 1.0f;                      // layer-1 bias (hi, lo) comes out of the MMA
                code[47] = plain ? 0.0f : 1.0f;                        // learn_empty: W_feat . empty_feature (bts.py:311-319)
#pragma unroll
                for (int d = 0; d < 3; ++d) {                          // hi/lo split of the raw coordinates (see mlp_pack_kernel)
                    const float hi = __half2float(__float2half_rn(code[d]));
                    code[39 + d] = code[d] - hi;
                    code[42 + d] = hi;
                }
                float s[3], c[3];
                const float a0 = x * P.fp.enc.freq_factor, a1 = y * P.fp.enc.freq_factor, a2 = zp * P.fp.enc.freq_factor;
                if (fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fabsf(a2)) <= 3.2f) {    // the usual case: hardware sin / cos
                    s[0] = __sinf(a0); c[0] = __cosf(a0); s[1] = __sinf(a1); c[1] = __cosf(a1); s[2] = __sinf(a2); c[2] = __cosf(a2);
                } else {                                                         // next to / behind the camera: |z'| is large
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
            const int cs = (int)(j % NCODE);
            TB_TRACE(pt_role, j, 4);
            mbar_wait(BAR(BAR_EMPTY_C + cs), (uint32_t)(((j / NCODE) & 1) ^ 1));
            TB_TRACE(pt_role, j, 5);
            unsigned char *crow = sm + OFF_CODE + cs * CHUNK + row * 128;
#pragma unroll
            for (int qq = 0; qq < 6; ++qq)
                *reinterpret_cast<uint4 *>(crow + ((qq ^ (row & 7)) << 4)) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
            fence_proxy_async();
            mbar_arrive_warp(BAR(BAR_FULL_C + cs));
            TB_TRACE(pt_role, j, 6);
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------
    bulk_wait<0>();            // output rows still in flight (epilogue threads)
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tb

// debug: clock64 trace of the last launch made with SD_TC_DEBUG & 8192 (not part of the public header)
extern "C" int sd_debug_read_trace_bin(long long *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tb::g_trace, sizeof(long long) * 8 * 64 * 8));
    return SD_OK;
}

#ifdef SD_DEBUG_WAIT
extern "C" int sd_debug_read_timeout(unsigned int *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tcx::g_wait_timeout, sizeof(unsigned int) * 260));
    return SD_OK;
}
#endif

bool bin_kernel_supported(const sd_scene *s, const sd_mlp *mlp) {
    return s && mlp && s->feat_proj && mlp->packed && mlp->precision == SD_MLP_F16_TC && s->C == 256 && s->nv_f == 1 &&
           s->include_input && s->num_freqs == 6 && s->Hf >= 2 && s->Wf >= 2 && s->nv_c <= tb::MAX_NVC_TB &&
           mlp->d_hidden == 128 && mlp->d_in == s->C + 39 && mlp->d_out >= 2 && mlp->d_out - 1 <= 64;
}

int launch_field_bin(const sd_scene *scene, const FieldParams &fp, const float *xyz, long long N, const sd_mlp *mlp,
                     const BinOrder &order, const TcOut &out, cudaStream_t st) {
    if (N == 0) return SD_OK;
    SD_REQUIRE(bin_kernel_supported(scene, mlp), "field_bin: unsupported scene / head for the projected-map kernel");
    SD_REQUIRE(order.bw == SD_BIN, "field_bin: the feature map is too large for %d x %d bins", SD_BIN, SD_BIN);
    SD_REQUIRE(N < (1ll << 31), "field_bin: at most 2^31 - 1 points per call (got %lld)", N);
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    const unsigned char *proj = reinterpret_cast<const unsigned char *>(scene->feat_proj);
    SD_REQUIRE(((uintptr_t)blob & 15) == 0 && ((uintptr_t)proj & 15) == 0, "field_bin: packed blobs must be 16-byte aligned");
    SD_REQUIRE(((uintptr_t)out.dino & 15) == 0, "field_bin: dino must be 16-byte aligned");
    tb::Params P = {};
    P.fp = fp;
    P.xyz = xyz;
    P.perm = order.perm; P.pcb = order.pcb; P.cbin = order.cbin; P.nbx = order.nbx;
    P.N = N;
    {
        const char *e = getenv("SD_TC_DEBUG");
        P.dbg = e ? atoi(e) : 0;
    }
    P.n_tiles = (N + tb::TM - 1) / tb::TM;
    P.D = mlp->d_out - 1;
    P.n2 = (mlp->d_out + 15) / 16 * 16;
    P.wc_img = proj;
    P.w2_img = blob + L.off_w_out_h;
    P.b_out = reinterpret_cast<const float *>(blob + L.off_b_out);
    P.sigma = out.sigma; P.dino = out.dino; P.rgb = out.rgb; P.invalid = out.invalid; P.invalid_feat = out.invalid_feat;
    const unsigned long long dims[3] = {128ull, (unsigned long long)fp.Wf, (unsigned long long)fp.Hf};
    const unsigned long long strides[2] = {256ull, 256ull * (unsigned long long)fp.Wf};
    const unsigned int box[3] = {64u, 8u, 8u};
    int rc = make_tmap_f16(&P.tmap, proj + tb::CHUNK, 3, dims, strides, box);
    if (rc) return rc;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        SD_CUDA_OK(cudaGetDevice(&dev));
        SD_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        SD_CUDA_OK(cudaFuncSetAttribute(tb::field_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tb::SMEM_ALLOC));
    }
    const unsigned grid = (unsigned)(P.n_tiles < sm_count ? P.n_tiles : sm_count);
    tb::field_bin_kernel<<<grid, tb::NTHREADS, tb::SMEM_ALLOC, st>>>(P);
    SD_LAUNCH_OK("field_bin_kernel");
    return SD_OK;
}

}  // namespace sd
