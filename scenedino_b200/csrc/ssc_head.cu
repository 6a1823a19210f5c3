// The unsupervised SSC head fused with the feature expansion that feeds it (SURVEY 8f-1 + 8f-2):
//
//   dino_full = F.normalize(W2e relu(W1e f + b1e) + b2e)                   MlpDimReduction.transform_expand
//                                                                          models/backbones/dino/dim_reduction.py:22-25
//   x^   = _norm(dino_full)                                               SemanticHead.forward, semantic_head.py:107-112
//   code = _norm(Wl x^ + bl  +  Wn2 relu(Wn1 x^ + bn1) + bn2)              StegoClusterHead.forward, :285-305 (eval: no dropout)
//   seg  = pseudo_assignment[argmax_k  normalize(code) . normalize(centre_k)]   KMeansParamHead, :308-373
//
// per voxel of the SSC query (models/bts.py:584-592; sscbench/evaluate_model_sscbench.py:850).  The reference writes the
// 768-d row (3 KB per voxel, 6.4 GB per grid) and runs 1.38 MFLOP per voxel of head on it.  Here the 768-d row never
// exists.  With h = relu(W1e f + b1e) (128-d), v = W2e h + b2e and s = 1 / |v|, everything between the two ReLUs is
// linear, so it is folded ONCE per model (sd_ssc_head_pack):
//   |v|^2            = h^T G h + 2 g.h + c            G = W2e^T W2e,  g = W2e^T b2e,  c = |b2e|^2
//   Wn1 x^ + bn1     = s (M1 h + m1)  + bn1           M1 = Wn1 W2e [768 x 128],  m1 = Wn1 b2e
//   relu(.)          = s relu(M1 h + m1 + |v| bn1)    (s > 0)
//   Wl x^            = s (ML h + ml)                  ML = Wl W2e [64 x 128],   ml = Wl b2e
//   code (unnormed)  = s [ ML h + ml + Wn2 relu(M1 h + m1 + |v| bn1) ] + bl + bn2
// 0.36 MFLOP per voxel instead of 1.6, all of it on the tensor cores, results identical up to rounding (measured against
// the oracle in fp32: 2e-7 on the cosine scores; with fp16 operands 3e-4).
//
// Kernel: persistent, one CTA per SM, 10 warps, TWO 128-row tiles ("lanes") in flight that share every streamed weight
// chunk.  Per lane the tensor pipe and the lane's four epilogue warps play strict ping-pong through two mbarriers
// (MMA_DONE / EPI_DONE); while one lane's epilogue converts an accumulator, the other lane's MMAs run.
//   warps 0-3 / 4-7  epilogue of lane 0 / 1 (thread = row = TMEM lane): operand build of the tile (fp32 rows -> fp16
//                    K-major SWIZZLE_128B), then per tile
//                      a  h  = relu(D + b1e)            -> fp16 into TMEM (A operand of every later MMA of the tile)
//                      b  |v| from q = G h (one accumulator) and h
//                      c  x6: relu(D + m1 + |v| bn1)    -> fp16 over the accumulator's own columns (A operand of Wn2's K slice)
//                      d  code -> normalise -> fp16 into TMEM (A operand of the score MMA)
//                      e  27 cosine scores -> argmax -> pseudo-label LUT -> 1 byte per voxel
//   warp 8           streams the weight chunks (48 KB each: [W1e | G], then 6 x [M1 block | Wn2 slice]) through a
//                    3-slot ring with bulk copies: 336 KB from L2 per PAIR of tiles
//   warp 9           issues every MMA (one thread), TMEM owner
// TMEM (512 columns): h of lane 0 / 1 at 0 / 64, code accumulators at 128 / 192, one 128-column accumulator per lane at
// 256 / 384 (layer-1, G, the M1 blocks and the scores take turns in it).
#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {
namespace sh {
using namespace tcx;

constexpr int TM = 128;
constexpr int D_RED = 64, D_LAT = 128, D_CODE = 64, MAX_CLS = 32, MAX_CHUNKS = 8;
constexpr int CHUNK_BYTES = 49152;           // one streamed chunk: 32 KB + 16 KB operand images
constexpr int NSLOT = 3;
// blob (sd_ssc_head_pack): streamed chunks first (chunk 0 = [W1e image 16 KB | G image 32 KB], chunk 1 + c = [M1 block c
// 32 KB | Wn2 K-slice c 16 KB]), then the resident images and the fp32 vectors
constexpr size_t BLOB_OFF_ML(int nch) { return (size_t)(1 + nch) * CHUNK_BYTES; }          // [64 rows][128 k] image, 16 KB
constexpr size_t BLOB_OFF_CEN(int nch) { return BLOB_OFF_ML(nch) + 16384; }              // [32 rows][64 k] image, 4 KB
constexpr size_t BLOB_OFF_VEC(int nch) { return BLOB_OFF_CEN(nch) + 4096; }
// fp32 vectors: b1e[128] | 2g[128] | (m1, bn1) interleaved [d_mid][2] | ml[64] | bl + bn2 [64] | c, n_cls, 0, 0 | lut bytes [32]
constexpr int VEC_B1E = 0, VEC_G2 = 128, VEC_M1 = 256;
constexpr int VEC_ML(int dmid) { return VEC_M1 + 2 * dmid; }
constexpr int VEC_BSUM(int dmid) { return VEC_ML(dmid) + 64; }
constexpr int VEC_CC(int dmid) { return VEC_BSUM(dmid) + 64; }
constexpr int VEC_LUT(int dmid) { return VEC_CC(dmid) + 4; }
constexpr int VEC_FLOATS(int dmid) { return VEC_LUT(dmid) + 8; }
constexpr size_t blob_bytes(int dmid) { return BLOB_OFF_VEC(dmid / 128) + (size_t)VEC_FLOATS(dmid) * 4; }

// shared memory
constexpr int OFF_RING = 0;
constexpr int OFF_ML = OFF_RING + NSLOT * CHUNK_BYTES;
constexpr int OFF_CEN = OFF_ML + 16384;
constexpr int OFF_A = OFF_CEN + 4096;                    // 2 lanes x [128 rows][64 k] fp16
constexpr int OFF_VEC = OFF_A + 2 * 16384;
constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS(MAX_CHUNKS * 128) * 4;
enum { BAR_RING_FULL = 0, BAR_RING_EMPTY = NSLOT, BAR_MMA_DONE = 2 * NSLOT, BAR_EPI_DONE = BAR_MMA_DONE + 2,
       BAR_RES = BAR_EPI_DONE + 2, NBAR = BAR_RES + 1 };
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_ALLOC = OFF_TMEM + 16 + 1024;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(OFF_ML % 1024 == 0 && OFF_CEN % 1024 == 0 && OFF_A % 1024 == 0 && OFF_VEC % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
constexpr int NTHREADS = 320;
constexpr int WARP_TMA = 8, WARP_MMA = 9;
constexpr int TMEM_COLS = 512;
constexpr int HE_COL = 0, CODE_COL = 128, ACC_COL = 256;

struct Params {
    const float *f;                 // [N][64] fp32 features of the field query
    const unsigned int *perm;       // or NULL: row r of the input stands for voxel perm[r] (outputs go to perm[r])
    const unsigned char *blob;
    long long N, n_tiles;
    int nch, n_cls;
    unsigned char *seg, *pseudo;    // [N] labels after / before the pseudo-label LUT (either may be NULL)
    float *scores;                  // [N][n_cls] cosine scores or NULL
};

__global__ void __launch_bounds__(NTHREADS, 1) ssc_head_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;   // (warp index the compiler knows to be uniform)
    const uint32_t bar0 = sm_u + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int dmid = P.nch * 128;
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(BAR(BAR_RING_FULL + s), 1); mbar_init(BAR(BAR_RING_EMPTY + s), 1); }
        for (int l = 0; l < 2; ++l) { mbar_init(BAR(BAR_MMA_DONE + l), 1); mbar_init(BAR(BAR_EPI_DONE + l), 4); }
        mbar_init(BAR(BAR_RES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float *s_vec = reinterpret_cast<float *>(sm + OFF_VEC);
    {
        const float *gv = reinterpret_cast<const float *>(P.blob + BLOB_OFF_VEC(P.nch));
        for (int i = tid; i < VEC_FLOATS(dmid); i += NTHREADS) s_vec[i] = __ldg(gv + i);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(BAR(BAR_RES), 16384 + 4096);
        bulk_g2s(sm_u + OFF_ML, P.blob + BLOB_OFF_ML(P.nch), 16384, BAR(BAR_RES));
        bulk_g2s(sm_u + OFF_CEN, P.blob + BLOB_OFF_CEN(P.nch), 4096, BAR(BAR_RES));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);
    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_tiles = P.n_tiles > first ? (P.n_tiles - first + stride - 1) / stride : 0;
    const long long n_pairs = (my_tiles + 1) / 2;
    const int nsteps = 1 + P.nch;                 // streamed chunks per pair of tiles

    if (warp < 8) {
        // =================================== EPILOGUE WARPS ==========================================
        const int ln = warp >> 2, wq = warp & 3;                       // lane of the pair, TMEM lane quadrant
        const int r_tile = wq * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
        const uint32_t t_he = t_lane + HE_COL + ln * 64, t_code = t_lane + CODE_COL + ln * 64, t_acc = t_lane + ACC_COL + ln * 128;
        unsigned char *a_row = sm + OFF_A + ln * 16384 + r_tile * 128;
        const float *s_b1e = s_vec + VEC_B1E, *s_g2 = s_vec + VEC_G2;
        const float2 *s_m1 = reinterpret_cast<const float2 *>(s_vec + VEC_M1);
        const float *s_ml = s_vec + VEC_ML(dmid), *s_bsum = s_vec + VEC_BSUM(dmid);
        const float cc = s_vec[VEC_CC(dmid)];
        const unsigned char *s_lut = reinterpret_cast<const unsigned char *>(s_vec + VEC_LUT(dmid));
        uint32_t ph = 0;                                             // phase of MMA_DONE[ln] this role waits for next
        auto wait_mma = [&]() { mbar_wait(BAR(BAR_MMA_DONE + ln), ph); ph ^= 1; tc_fence_after(); };
        auto epi_done = [&]() { tc_fence_before(); mbar_arrive_warp(BAR(BAR_EPI_DONE + ln)); };
        // operand build: this thread's row, fp32 -> fp16, 16-byte chunk q at position q ^ (row & 7)
        auto build_a = [&](long long tile) {
            const long long row = tile * TM + r_tile;
            const uint4 *src = reinterpret_cast<const uint4 *>(P.f + row * D_RED);
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
            fence_proxy_async();              // generic-proxy writes -> visible to the MMA's async-proxy reads
        };
        if (ln < my_tiles) build_a(first + ln * stride);
        epi_done();
        for (long long j = ln; j < my_tiles; j += 2) {
            const long long tile = first + j * stride;
            const long long row = tile * TM + r_tile;
            // ---- a: hidden = relu(D1 + b1e) as fp16 pairs --------------------------------------------------------
            wait_mma();
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
                uint32_t vr[32], pk[16];
                tmem_ld32_issue(t_acc + kb * 32, vr);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    pk[e] = pack_h2_relu(__uint_as_float(vr[2 * e]) + s_b1e[kb * 32 + 2 * e],
                                         __uint_as_float(vr[2 * e + 1]) + s_b1e[kb * 32 + 2 * e + 1]);
                tmem_st16(t_he + kb * 16, pk);
            }
            tmem_st_wait();
            epi_done();
            // ---- b: |v|^2 = h.(G h + 2 g) + c  (h as the tensor cores see it: the fp16 values) ----------------------
            wait_mma();
            float vv = cc;
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
                uint32_t qr[32], hr[16];
                tmem_ld32_issue(t_acc + kb * 32, qr);
                tmem_ld16_issue(t_he + kb * 16, hr);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float2 h2 = half2_bits_to_float2(hr[e]);
                    vv = fmaf(h2.x, __uint_as_float(qr[2 * e]) + s_g2[kb * 32 + 2 * e], vv);
                    vv = fmaf(h2.y, __uint_as_float(qr[2 * e + 1]) + s_g2[kb * 32 + 2 * e + 1], vv);
                }
            }
            const float nv = fmaxf(sqrtf(fmaxf(vv, 0.0f)), 1e-12f);     // F.normalize eps (dim_reduction.py:25)
            const float sc = 1.0f / nv;
            epi_done();
            // ---- c: the 768 hidden units of the non-linear path, 128 at a time -------------------------------------
            for (int c = 0; c < P.nch; ++c) {
                wait_mma();
                const float2 *cm = s_m1 + c * 128;
#pragma unroll 1
                for (int kb = 0; kb < 4; ++kb) {
                    uint32_t vr[32], pk[16];
                    tmem_ld32_issue(t_acc + kb * 32, vr);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float2 c0 = cm[kb * 32 + 2 * e], c1 = cm[kb * 32 + 2 * e + 1];
                        pk[e] = pack_h2_relu(fmaf(nv, c0.y, __uint_as_float(vr[2 * e]) + c0.x),
                                             fmaf(nv, c1.y, __uint_as_float(vr[2 * e + 1]) + c1.x));
                    }
                    tmem_st16(t_acc + kb * 16, pk);           // columns already read
                }
                tmem_st_wait();
                epi_done();
            }
            // ---- d: code = s (acc + ml) + bl + bn2, normalised (semantic_head.py:305, 360) -> fp16 A operand of the scores
            wait_mma();
            {
                uint32_t vr[64];
                tmem_ld32_issue(t_code, vr);
                tmem_ld32_issue(t_code + 32, vr + 32);
                tmem_ld_wait();
                float code[64], ss = 0.0f;
#pragma unroll
                for (int k = 0; k < 64; ++k) {
                    code[k] = fmaf(sc, __uint_as_float(vr[k]) + s_ml[k], s_bsum[k]);
                    ss = fmaf(code[k], code[k], ss);
                }
                const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-10f);
                uint32_t pk[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) pk[e] = pack_h2(code[2 * e] * inv, code[2 * e + 1] * inv);
                tmem_st32(t_he, pk);
                tmem_st_wait();
            }
            epi_done();
            // ---- e: cosine scores -> first maximum -> pseudo-label LUT ------------------------------------------------
            wait_mma();
            {
                uint32_t vr[32];
                tmem_ld32_issue(t_acc, vr);
                tmem_ld_wait();
                int best = 0;
                float bv = __uint_as_float(vr[0]);
#pragma unroll
                for (int k = 1; k < MAX_CLS; ++k) {
                    const float v = __uint_as_float(vr[k]);
                    if (k < P.n_cls && v > bv) { bv = v; best = k; }
                }
                if (row < P.N) {
                    const long long dst = P.perm ? (long long)__ldg(P.perm + row) : row;
                    if (P.pseudo) P.pseudo[dst] = (unsigned char)best;
                    if (P.seg) P.seg[dst] = s_lut[best];
                    if (P.scores) {
#pragma unroll
                        for (int k = 0; k < MAX_CLS; ++k)
                            if (k < P.n_cls) P.scores[dst * P.n_cls + k] = __uint_as_float(vr[k]);
                    }
                }
            }
            if (j + 2 < my_tiles) build_a(first + (j + 2) * stride);
            epi_done();
        }
    } else if (warp == WARP_TMA) {
        // (whole warp in lockstep, the copy inside an elect_one() branch: tc_common.cuh, "elected" forms)
        long long g = 0;
        for (long long p = 0; p < n_pairs; ++p)
            for (int s = 0; s < nsteps; ++s, ++g) {
                const int slot = (int)(g % NSLOT);
                mbar_wait_warp(BAR(BAR_RING_EMPTY + slot), (uint32_t)(((g / NSLOT) & 1) ^ 1));
                if (elect_one()) {
                    mbar_expect_tx(BAR(BAR_RING_FULL + slot), CHUNK_BYTES);
                    bulk_g2s(sm_u + OFF_RING + slot * CHUNK_BYTES, P.blob + (size_t)s * CHUNK_BYTES, CHUNK_BYTES, BAR(BAR_RING_FULL + slot));
                }
                __syncwarp();
            }
    } else {
        // =================================== MMA ISSUER ===============================================
        // The whole warp runs the issue loop in lockstep; one elected lane issues (tc_common.cuh: under `if (lane == 0)` every
        // tcgen05 instruction was wrapped in an ELECT loop, ~90 cycles each -- ~240 MMAs per pair of tiles: the issue rate,
        // not the tensor pipe, bounded this kernel).
        {
            mbar_wait_warp(BAR(BAR_RES), 0);
            const uint32_t id128 = umma_idesc(TM, 128), id64 = umma_idesc(TM, 64), id32 = umma_idesc(TM, 32);
            uint32_t pe[2] = {0u, 0u};                       // phase of EPI_DONE[l] to wait for next
            auto wait_epi = [&](int l) { mbar_wait_warp(BAR(BAR_EPI_DONE + l), pe[l]); pe[l] ^= 1; tc_fence_after(); };
            long long g = 0;
            for (long long p = 0; p < n_pairs; ++p) {
                const int nl = (2 * p + 1 < my_tiles) ? 2 : 1;
                // ---- chunk 0: [W1e | G]
                int slot = (int)(g % NSLOT);
                uint32_t ring = sm_u + OFF_RING + slot * CHUNK_BYTES;
                mbar_wait_warp(BAR(BAR_RING_FULL + slot), (uint32_t)((g / NSLOT) & 1));
                tc_fence_after();
                for (int l = 0; l < nl; ++l) {               // layer 1 of the expansion: A from shared memory
                    wait_epi(l);                             // operand built, accumulator drained
                    if (elect_one()) {          // ONE elected thread issues: a single-thread region (tc_common.cuh)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma(tmem_base + ACC_COL + l * 128, umma_desc(sm_u + OFF_A + l * 16384 + k * 32), umma_desc(ring + k * 32), id128, k != 0);
                        umma_commit(BAR(BAR_MMA_DONE + l));
                    }
                    __syncwarp();
                }
                for (int l = 0; l < nl; ++l) {               // q = G h (-> |v|) and the linear path ML h -> code accumulator
                    wait_epi(l);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_ts(tmem_base + ACC_COL + l * 128, tmem_base + HE_COL + l * 64 + k * 8,
                                    umma_desc(ring + 16384 + (k >> 2) * 16384 + (k & 3) * 32), id128, k != 0);
                        umma_commit(BAR(BAR_MMA_DONE + l));
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_ts(tmem_base + CODE_COL + l * 64, tmem_base + HE_COL + l * 64 + k * 8,
                                    umma_desc(sm_u + OFF_ML + (k >> 2) * 8192 + (k & 3) * 32), id64, k != 0);
                        if (l == nl - 1) umma_commit(BAR(BAR_RING_EMPTY + slot));
                    }
                    __syncwarp();
                }
                ++g;
                // ---- chunks 1..nch: [M1 block c | Wn2 K-slice c]
                for (int c = 0; c < P.nch; ++c, ++g) {
                    slot = (int)(g % NSLOT);
                    ring = sm_u + OFF_RING + slot * CHUNK_BYTES;
                    mbar_wait_warp(BAR(BAR_RING_FULL + slot), (uint32_t)((g / NSLOT) & 1));
                    tc_fence_after();
                    for (int l = 0; l < nl; ++l) {
                        if (c == 0) wait_epi(l);             // |v| taken: the accumulator is free (later blocks: the tensor pipe
                                                             // runs this thread's MMAs in order, behind the Wn2 slice that read it)
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                umma_ts(tmem_base + ACC_COL + l * 128, tmem_base + HE_COL + l * 64 + k * 8,
                                        umma_desc(ring + (k >> 2) * 16384 + (k & 3) * 32), id128, k != 0);
                            umma_commit(BAR(BAR_MMA_DONE + l));
                        }
                        __syncwarp();
                    }
                    for (int l = 0; l < nl; ++l) {
                        wait_epi(l);                         // the block's hidden units are in the accumulator's columns as fp16
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                umma_ts(tmem_base + CODE_COL + l * 64, tmem_base + ACC_COL + l * 128 + k * 8,
                                        umma_desc(ring + 32768 + (k >> 2) * 8192 + (k & 3) * 32), id64, 1);
                            if (l == nl - 1) {
                                umma_commit(BAR(BAR_RING_EMPTY + slot));
                                if (c == P.nch - 1)                        // code accumulators complete
                                    for (int l2 = 0; l2 < nl; ++l2) umma_commit(BAR(BAR_MMA_DONE + l2));
                            }
                        }
                        __syncwarp();
                    }
                }
                for (int l = 0; l < nl; ++l) {               // cosine scores against the centres
                    wait_epi(l);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_ts(tmem_base + ACC_COL + l * 128, tmem_base + HE_COL + l * 64 + k * 8, umma_desc(sm_u + OFF_CEN + k * 32), id32, k != 0);
                        umma_commit(BAR(BAR_MMA_DONE + l));
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- one-off folding of the head's weights (sd_ssc_head_pack) -------------------------------------------------------
struct PackIn {
    const float *w1e, *b1e, *w2e, *b2e;      // MlpDimReduction: linear_in [128,64], linear_out [d_full,128]
    const float *wl, *bl, *wn1, *bn1, *wn2, *bn2;   // StegoClusterHead: linear [64,d_full], nonlinear [d_mid,d_full], [64,d_mid]
    const float *centres;                    // [n_cls, 64]
    const long long *lut;                    // [n_cls] pseudo_assignment
    int d_full, d_mid, n_cls;
};

__device__ __forceinline__ void put_h(unsigned char *img, int row, int k, int rows, float v) {
    reinterpret_cast<__half *>(img)[umma_sw128_offset(row, k, rows) / 2] = __float2half_rn(v);
}

__global__ void __launch_bounds__(256) ssc_pack_kernel(PackIn I, unsigned char *blob) {
    const int nch = I.d_mid / 128;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    float *vec = reinterpret_cast<float *>(blob + BLOB_OFF_VEC(nch));
    // chunk 0: W1e image [128 units][64 k], G image [128][128]
    for (long long i = tid; i < 128 * 64; i += nth) put_h(blob, (int)(i >> 6), (int)(i & 63), 128, I.w1e[i]);
    for (long long i = tid; i < 128 * 128; i += nth) {
        const int a = (int)(i >> 7), b = (int)(i & 127);
        double acc = 0.0;
        for (int j = 0; j < I.d_full; ++j) acc += (double)I.w2e[(size_t)j * 128 + a] * I.w2e[(size_t)j * 128 + b];
        put_h(blob + 16384, a, b, 128, (float)acc);
    }
    // chunks 1..nch: M1 = Wn1 W2e in blocks of 128 rows, Wn2 in K slices of 128 columns
    for (long long i = tid; i < (long long)I.d_mid * 128; i += nth) {
        const int m = (int)(i >> 7), k = (int)(i & 127);
        double acc = 0.0;
        const float *w = I.wn1 + (size_t)m * I.d_full;
        for (int j = 0; j < I.d_full; ++j) acc += (double)w[j] * I.w2e[(size_t)j * 128 + k];
        put_h(blob + (size_t)(1 + (m >> 7)) * CHUNK_BYTES, m & 127, k, 128, (float)acc);
    }
    for (long long i = tid; i < (long long)64 * I.d_mid; i += nth) {
        const int o = (int)(i / I.d_mid), m = (int)(i - (long long)o * I.d_mid);
        put_h(blob + (size_t)(1 + (m >> 7)) * CHUNK_BYTES + 32768, o, m & 127, 64, I.wn2[i]);
    }
    // resident: ML = Wl W2e [64][128]; normalised centres [32][64]
    for (long long i = tid; i < 64 * 128; i += nth) {
        const int o = (int)(i >> 7), k = (int)(i & 127);
        double acc = 0.0;
        for (int j = 0; j < I.d_full; ++j) acc += (double)I.wl[(size_t)o * I.d_full + j] * I.w2e[(size_t)j * 128 + k];
        put_h(blob + BLOB_OFF_ML(nch), o, k, 64, (float)acc);
    }
    for (long long i = tid; i < MAX_CLS * 64; i += nth) {
        const int c = (int)(i >> 6), k = (int)(i & 63);
        float v = 0.0f;
        if (c < I.n_cls) {
            double ss = 0.0;
            for (int q = 0; q < 64; ++q) ss += (double)I.centres[c * 64 + q] * I.centres[c * 64 + q];
            const float nrm = (float)sqrt(ss);
            v = I.centres[c * 64 + k] / (nrm > 1e-12f ? nrm : 1e-12f);        // F.normalize (semantic_head.py:359)
        }
        put_h(blob + BLOB_OFF_CEN(nch), c, k, MAX_CLS, v);
    }
    // vectors
    for (long long i = tid; i < 128; i += nth) {
        vec[VEC_B1E + i] = I.b1e[i];
        double acc = 0.0;
        for (int j = 0; j < I.d_full; ++j) acc += (double)I.w2e[(size_t)j * 128 + i] * I.b2e[j];
        vec[VEC_G2 + i] = (float)(2.0 * acc);
    }
    for (long long m = tid; m < I.d_mid; m += nth) {
        double acc = 0.0;
        for (int j = 0; j < I.d_full; ++j) acc += (double)I.wn1[(size_t)m * I.d_full + j] * I.b2e[j];
        vec[VEC_M1 + 2 * m] = (float)acc;
        vec[VEC_M1 + 2 * m + 1] = I.bn1[m];
    }
    for (long long o = tid; o < 64; o += nth) {
        double acc = 0.0;
        for (int j = 0; j < I.d_full; ++j) acc += (double)I.wl[(size_t)o * I.d_full + j] * I.b2e[j];
        vec[VEC_ML(I.d_mid) + o] = (float)acc;
        vec[VEC_BSUM(I.d_mid) + o] = I.bl[o] + I.bn2[o];
    }
    if (tid == 0) {
        double acc = 0.0;
        for (int j = 0; j < I.d_full; ++j) acc += (double)I.b2e[j] * I.b2e[j];
        vec[VEC_CC(I.d_mid)] = (float)acc;
        vec[VEC_CC(I.d_mid) + 1] = (float)I.n_cls;
        vec[VEC_CC(I.d_mid) + 2] = vec[VEC_CC(I.d_mid) + 3] = 0.0f;
        unsigned char *lut = reinterpret_cast<unsigned char *>(vec + VEC_LUT(I.d_mid));
        for (int c = 0; c < MAX_CLS; ++c) lut[c] = c < I.n_cls ? (unsigned char)I.lut[c] : 0;
    }
}

}  // namespace sh
}  // namespace sd

using namespace sd;

static bool ssc_dims_ok(int d_red, int d_lat, int d_full, int d_mid, int d_code, int n_cls) {
    return d_red == sh::D_RED && d_lat == sh::D_LAT && d_code == sh::D_CODE && d_full >= 128 && d_full <= 8192 &&
           d_mid >= 128 && d_mid % 128 == 0 && d_mid <= sh::MAX_CHUNKS * 128 && n_cls >= 1 && n_cls <= sh::MAX_CLS;
}

extern "C" size_t sd_ssc_head_pack_bytes(int d_red, int d_lat, int d_full, int d_mid, int d_code, int n_cls) {
    if (!ssc_dims_ok(d_red, d_lat, d_full, d_mid, d_code, n_cls)) return 0;
    return sh::blob_bytes(d_mid);
}

extern "C" int sd_ssc_head_pack(const float *w1e, const float *b1e, const float *w2e, const float *b2e, const float *wl,
                                const float *bl, const float *wn1, const float *bn1, const float *wn2, const float *bn2,
                                const float *centres, const long long *lut, int d_red, int d_lat, int d_full, int d_mid,
                                int d_code, int n_cls, void *packed, void *stream) {
    SD_REQUIRE(ssc_dims_ok(d_red, d_lat, d_full, d_mid, d_code, n_cls),
               "sd_ssc_head_pack: supports expand 64 -> 128 -> d_full, STEGO d_full -> (d_mid <= 1024, multiple of 128) -> 64, "
               "<= 32 clusters (got %d -> %d -> %d, mid %d, code %d, %d clusters)", d_red, d_lat, d_full, d_mid, d_code, n_cls);
    SD_REQUIRE(w1e && b1e && w2e && b2e && wl && bl && wn1 && bn1 && wn2 && bn2 && centres && lut && packed, "sd_ssc_head_pack: null pointer");
    SD_REQUIRE(((uintptr_t)packed & 1023) == 0, "sd_ssc_head_pack: packed must be 1024-byte aligned");
    sh::PackIn I = {w1e, b1e, w2e, b2e, wl, bl, wn1, bn1, wn2, bn2, centres, lut, d_full, d_mid, n_cls};
    sh::ssc_pack_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(I, reinterpret_cast<unsigned char *>(packed));
    SD_LAUNCH_OK("ssc_pack_kernel");
    return SD_OK;
}

extern "C" int sd_ssc_head(const void *packed, int d_mid, int n_cls, const float *f, const unsigned int *perm, long long N,
                           unsigned char *seg, unsigned char *pseudo, float *scores, void *stream) {
    SD_REQUIRE(packed && d_mid >= 128 && d_mid % 128 == 0 && d_mid <= sh::MAX_CHUNKS * 128 && n_cls >= 1 && n_cls <= sh::MAX_CLS,
               "sd_ssc_head: bad head (d_mid %d, %d clusters)", d_mid, n_cls);
    SD_REQUIRE(N >= 0, "sd_ssc_head: bad N");
    if (N == 0) return SD_OK;
    SD_REQUIRE(f && (seg || pseudo || scores), "sd_ssc_head: null pointer");
    SD_REQUIRE(((uintptr_t)f & 15) == 0 && ((uintptr_t)packed & 1023) == 0, "sd_ssc_head: f must be 16-byte, packed 1024-byte aligned");
    sh::Params P = {};
    P.f = f; P.perm = perm; P.blob = reinterpret_cast<const unsigned char *>(packed);
    P.N = N; P.n_tiles = (N + sh::TM - 1) / sh::TM;
    P.nch = d_mid / 128; P.n_cls = n_cls;
    P.seg = seg; P.pseudo = pseudo; P.scores = scores;
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(sh::ssc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sh::SMEM_ALLOC));
    }
    // two tiles per CTA iteration: a grid of at most ceil(tiles / 2) CTAs keeps both lanes of every CTA busy
    const long long want = (P.n_tiles + 1) / 2;
    const unsigned grid = (unsigned)(want < sm_count ? want : sm_count);
    profile_before((cudaStream_t)stream);
    sh::ssc_head_kernel<<<grid, sh::NTHREADS, sh::SMEM_ALLOC, (cudaStream_t)stream>>>(P);
    profile_after((cudaStream_t)stream);
    SD_LAUNCH_OK("ssc_head_kernel");
    return SD_OK;
}
