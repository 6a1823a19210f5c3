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

#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {
namespace tb {
using namespace tcx;

constexpr int TM = 128;
constexpr int CHUNK = 16384;                 // A chunk [128 rows][64 slots] fp16 = B chunk [2 halves][64 slots][64 ch] fp16
constexpr int NRING = 3, NCODE = 2;
constexpr int N_EPI_WARPS = 4, N_PT_WARPS = 4;
constexpr int WARP_MMA = N_EPI_WARPS, WARP_TMA = WARP_MMA + 1, WARP_PT0 = WARP_TMA + 1;
constexpr int NTHREADS = (WARP_PT0 + N_PT_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int D2_COL = 256, D2_STRIDE = 128;  // layer-1 accumulators at columns 0 / 128, layer-2 at 256 / 384
constexpr int MAX_NVC_TB = 4;
constexpr int W2_BYTES = 2 * 80 * 128;
constexpr int KCODE = 3;                     // K steps of the code block (48 columns)

constexpr int OFF_WC = 0;
constexpr int OFF_A = OFF_WC + CHUNK;
constexpr int OFF_B = OFF_A + NRING * CHUNK;
constexpr int OFF_CODE = OFF_B + NRING * CHUNK;
constexpr int OFF_W2 = OFF_CODE + NCODE * CHUNK;
constexpr int OFF_STAGE = OFF_W2 + W2_BYTES;
constexpr int OFF_DIRTY = OFF_STAGE + N_EPI_WARPS * 8192;
constexpr int OFF_CAM = OFF_DIRTY + NRING * TM;
constexpr int OFF_BAR = OFF_CAM + 448;
enum { BAR_FULL_A = 0, BAR_FULL_B = NRING, BAR_EMPTY = 2 * NRING, BAR_FULL_C = 3 * NRING, BAR_EMPTY_C = BAR_FULL_C + NCODE,
       BAR_D1 = BAR_EMPTY_C + NCODE, BAR_H = BAR_D1 + 2, BAR_D2 = BAR_H + 2, BAR_D2_EMPTY = BAR_D2 + 2,
       BAR_WLOAD = BAR_D2_EMPTY + 2, NBAR = BAR_WLOAD + 1 };
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_ALLOC = OFF_TMEM + 16 + 1024;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(OFF_W2 % 1024 == 0 && OFF_STAGE % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

struct Params {
    CUtensorMap tmap;          // P as [Hf][Wf][128] fp16, box 8 x 8 x 64 channels, SWIZZLE_128B
    FieldParams fp;
    const float *xyz;
    const unsigned int *perm;
    const unsigned short *pcb;
    const unsigned int *cbin;
    int nbx;
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
    fence_proxy_async();
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t w2b = 2u * (uint32_t)P.n2 * 128u;
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
    // compact bins touched by my tile number jj: first one and how many (the sorted order makes them consecutive)
    auto span = [&](long long jj, int &c0, int &m) {
        c0 = 0; m = 1;
        if (jj >= my_tiles) return;
        const long long a = (first + jj * stride) * TM;
        const long long b = (a + TM < P.N ? a + TM : P.N) - 1;
        c0 = (int)__ldg(P.pcb + a);
        m = (int)__ldg(P.pcb + b) - c0 + 1;
    };

    if (warp < N_EPI_WARPS) {
        // =================================== EPILOGUE ================================================
        const int row = tid;
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int D = P.D;
        const float bo_sigma = __ldg(P.b_out);
        float bo4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int c = 4 * (lane & 15) + e; bo4[e] = c < D ? __ldg(P.b_out + 1 + c) : 0.0f; }
        unsigned char *stage0 = sm + OFF_STAGE + warp * 8192, *stage1 = stage0 + 4096;
        int grow_keep = -1;
        for (long long j = 0; j <= my_tiles; ++j) {
            if (j > 0) {
                // ---------------- second epilogue of tile j-1 -----------------------------------------
                const int b1 = (int)((j - 1) & 1);
                mbar_wait(BAR(BAR_D2 + b1), (uint32_t)(((j - 1) >> 1) & 1));
                tc_fence_after();
                uint32_t vr[64], sr;
                const uint32_t t_d2 = t_lane + D2_COL + b1 * D2_STRIDE;
                tmem_ld32_issue(t_d2, vr);
                tmem_ld32_issue(t_d2 + 32, vr + 32);
                tmem_ld1_issue(t_d2 + D, sr);                                    // density column sits behind the features
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_D2_EMPTY + b1));
                float v[64];
#pragma unroll
                for (int c = 0; c < 64; ++c) v[c] = __uint_as_float(vr[c]);
                const float sg = softplus(__uint_as_float(sr) + bo_sigma);
                const bool ok = grow_keep >= 0;
                if (ok && P.sigma) P.sigma[grow_keep] = sg;
                if (P.dino && D == 64) {
                    // transpose through shared memory: lane = row writes its 16 chunks (XOR-swizzled, conflict
                    // free), then 16 lanes read one row back and the warp stores two whole rows per request
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        *reinterpret_cast<float4 *>((q < 8 ? stage0 : stage1) + lane * 128 + (((q & 7) ^ (lane & 7)) << 4)) =
                            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    __syncwarp();
                    const int c16 = lane & 15;
                    const unsigned char *src = (c16 < 8 ? stage0 : stage1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = 2 * i + (lane >> 4);
                        float4 x = *reinterpret_cast<const float4 *>(src + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
                        x.x += bo4[0]; x.y += bo4[1]; x.z += bo4[2]; x.w += bo4[3];
                        const int dst = __shfl_sync(0xffffffffu, grow_keep, r);
                        if (dst >= 0) *reinterpret_cast<float4 *>(P.dino + (long long)dst * 64 + c16 * 4) = x;
                    }
                    __syncwarp();
                } else if (P.dino && ok) {
                    float *o = P.dino + (long long)grow_keep * D;
#pragma unroll
                    for (int c = 0; c < 64; ++c)
                        if (c < D) o[c] = v[c] + __ldg(P.b_out + 1 + c);
                }
            }
            if (j == my_tiles) break;
            // ---------------- first epilogue of tile j ----------------------------------------------------
            const long long gpos = (first + j * stride) * TM + row;
            grow_keep = gpos < P.N ? (int)__ldg(P.perm + gpos) : -1;
            const int b = (int)(j & 1);
            mbar_wait(BAR(BAR_D1 + b), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int kb = 0; kb < 2; ++kb) {      // 64 hidden units -> 32 packed columns, written over columns already read
                uint32_t vr[64], pk[32];
                tmem_ld32_issue(t_lane + b * 128 + kb * 64, vr);
                tmem_ld32_issue(t_lane + b * 128 + kb * 64 + 32, vr + 32);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    pk[e] = pack_h2(fmaxf(__uint_as_float(vr[2 * e]), 0.0f), fmaxf(__uint_as_float(vr[2 * e + 1]), 0.0f));
                tmem_st32(t_lane + b * 128 + kb * 32, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(BAR(BAR_H + b));
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
#pragma unroll
                for (int k = 0; k < 8; ++k)          // K = 16 per instruction = 8 packed columns of the hidden tile
                    umma_ts(tmem_base + D2_COL + b * D2_STRIDE, tmem_base + b * 128 + k * 8,
                            umma_desc(sm_u + OFF_W2 + (k >> 2) * P.n2 * 128 + (k & 3) * 32), idesc2, k != 0);
                umma_commit(BAR(BAR_D2 + b));
            };
            int e = 0;
            uint32_t ph = 0;
            int c0n, mn;
            span(0, c0n, mn);
            for (long long j = 0; j < my_tiles; ++j) {
                const int m = mn;
                span(j + 1, c0n, mn);
                const uint32_t d1 = tmem_base + (uint32_t)(j & 1) * 128u;
                uint32_t acc = 0;
                for (int i = 0; i < m; ++i) {
                    mbar_wait(BAR(BAR_FULL_B + e), ph);
                    mbar_wait(BAR(BAR_FULL_A + e), ph);
                    tc_fence_after();
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
                mbar_wait(BAR(BAR_FULL_C + cs), (uint32_t)((j / NCODE) & 1));
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < KCODE; ++k)
                    umma(d1, umma_desc(sm_u + OFF_CODE + cs * CHUNK + k * 32), umma_desc(sm_u + OFF_WC + k * 32), idesc_k, 1);
                umma_commit(BAR(BAR_EMPTY_C + cs));
                umma_commit(BAR(BAR_D1 + (int)(j & 1)));
                if (j > 0) layer2(j - 1);
            }
            if (my_tiles > 0) layer2(my_tiles - 1);
        }
    } else if (warp == WARP_TMA) {
        // =================================== TMA PRODUCER =============================================
        if (lane == 0) tma_prefetch_desc(&P.tmap);
        int e = 0;
        uint32_t ph = 0;
        int c0, m, c0n, mn;
        span(0, c0, m);
        span(1, c0n, mn);
        unsigned int mybin = lane < m ? __ldg(P.cbin + c0 + lane) : 0u;
        for (long long j = 0; j < my_tiles; ++j) {
            // next tile's bins and the span of the tile after it are in flight while this tile's boxes are issued
            const unsigned int mybin_n = lane < mn ? __ldg(P.cbin + c0n + lane) : 0u;
            int c0nn, mnn;
            span(j + 2, c0nn, mnn);
            for (int base = 0; base < m; base += 32) {
                if (base > 0) mybin = base + lane < m ? __ldg(P.cbin + c0 + base + lane) : 0u;
                const int cnt = m - base < 32 ? m - base : 32;
                for (int i = 0; i < cnt; ++i) {
                    const unsigned int b = __shfl_sync(0xffffffffu, mybin, i);
                    if (lane == 0) {
                        const int by = (int)(b / (unsigned)P.nbx), bx = (int)(b - (unsigned)by * (unsigned)P.nbx);
                        mbar_wait(BAR(BAR_EMPTY + e), ph ^ 1);
                        mbar_expect_tx(BAR(BAR_FULL_B + e), CHUNK);
                        const uint32_t dst = sm_u + OFF_B + e * CHUNK;
                        tma_load_3d(dst, &P.tmap, 0, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                        tma_load_3d(dst + CHUNK / 2, &P.tmap, 64, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                    }
                    if (++e == NRING) { e = 0; ph ^= 1; }
                }
                __syncwarp();
            }
            c0 = c0n; m = mn; mybin = mybin_n;
            c0n = c0nn; mn = mnn;
        }
    } else {
        // =================================== POINT WARPS ================================================
        const int row = tid - WARP_PT0 * 32;
        const int nv_c = P.fp.nv_c;
        struct RowIn { int grow; float px, py, pz; int c0, m, cr; };
        auto fetch = [&](long long jj) {
            RowIn r;
            r.grow = -1; r.px = r.py = r.pz = 0.0f; r.cr = 0;
            span(jj, r.c0, r.m);
            if (jj >= my_tiles) return r;
            const long long gpos = (first + jj * stride) * TM + row;
            if (gpos >= P.N) return r;
            r.grow = (int)__ldg(P.perm + gpos);
            r.cr = (int)__ldg(P.pcb + gpos);
            r.px = __ldg(P.xyz + 3ll * r.grow); r.py = __ldg(P.xyz + 3ll * r.grow + 1); r.pz = __ldg(P.xyz + 3ll * r.grow + 2);
            return r;
        };
        int e = 0;
        uint32_t ph = 0;
        RowIn nxt = fetch(0);
        for (long long j = 0; j < my_tiles; ++j) {
            const RowIn cur = nxt;
            nxt = fetch(j + 1);
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
                zp = znorm(zc, P.fp.enc);
                t = bilinear_tap(x, y, P.fp.Hf, P.fp.Wf);
                clamp_footprint(t, P.fp.Hf, P.fp.Wf);
            }
            // ---- bilinear weights -> this row's 4 slots of its bin's chunk; every other slot of the row stays zero
            const bool plain = ok && !(P.fp.learn_empty && inv);
            const int lx = t.x0 - (t.x0 / SD_BIN) * SD_BIN, ly = t.y0 - (t.y0 / SD_BIN) * SD_BIN;
            const int s00 = ly * 8 + lx;
            const int q = cur.cr - cur.c0;
            const unsigned short h_nw = __half_as_ushort(__float2half_rn(t.wnw)), h_ne = __half_as_ushort(__float2half_rn(t.wne));
            const unsigned short h_sw = __half_as_ushort(__float2half_rn(t.wsw)), h_se = __half_as_ushort(__float2half_rn(t.wse));
            auto slot_ptr = [&](unsigned char *arow, int s) {
                return reinterpret_cast<unsigned short *>(arow + (((s >> 3) ^ (row & 7)) << 4) + (s & 7) * 2);
            };
            for (int i = 0; i < cur.m; ++i) {
                mbar_wait(BAR(BAR_EMPTY + e), ph ^ 1);
                unsigned char *arow = sm + OFF_A + e * CHUNK + row * 128;
                const int d = s_dirty[e * TM + row];
                if (d != 0xFF) {
                    *slot_ptr(arow, d) = 0; *slot_ptr(arow, d + 1) = 0; *slot_ptr(arow, d + 8) = 0; *slot_ptr(arow, d + 9) = 0;
                }
                if (plain && i == q) {
                    *slot_ptr(arow, s00) = h_nw; *slot_ptr(arow, s00 + 1) = h_ne;
                    *slot_ptr(arow, s00 + 8) = h_sw; *slot_ptr(arow, s00 + 9) = h_se;
                    s_dirty[e * TM + row] = (unsigned char)s00;
                } else if (d != 0xFF) {
                    s_dirty[e * TM + row] = 0xFF;
                }
                fence_proxy_async();
                mbar_arrive_warp(BAR(BAR_FULL_A + e));
                if (++e == NRING) { e = 0; ph ^= 1; }
            }
            // ---- per-point outputs that do not need the head ---------------------------------------------------
            if (ok) {
                if (P.invalid_feat) P.invalid_feat[grow] = inv ? 1 : 0;
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
                sincosf(__fmul_rn(x, P.fp.enc.freq_factor), &s[0], &c[0]);
                sincosf(__fmul_rn(y, P.fp.enc.freq_factor), &s[1], &c[1]);
                sincosf(__fmul_rn(zp, P.fp.enc.freq_factor), &s[2], &c[2]);
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
            mbar_wait(BAR(BAR_EMPTY_C + cs), (uint32_t)(((j / NCODE) & 1) ^ 1));
            unsigned char *crow = sm + OFF_CODE + cs * CHUNK + row * 128;
#pragma unroll
            for (int qq = 0; qq < 6; ++qq)
                *reinterpret_cast<uint4 *>(crow + ((qq ^ (row & 7)) << 4)) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
            fence_proxy_async();
            mbar_arrive_warp(BAR(BAR_FULL_C + cs));
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

}  // namespace tb

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
    tb::Params P = {};
    P.fp = fp;
    P.xyz = xyz;
    P.perm = order.perm; P.pcb = order.pcb; P.cbin = order.cbin; P.nbx = order.nbx;
    P.N = N;
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
