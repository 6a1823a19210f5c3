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
// Per query the sort (binning.cu) leaves, at each point's SORTED position, a 32-byte record (coordinates, weights, point
// index, bin) and, per 128-point tile, a table entry (rows, bins it touches).  This kernel never projects a point and in
// steady state issues no global load through the load/store unit.
//
// Warp roles (15 warps, one persistent CTA per SM):
//   warps 0-3    epilogue 1: layer-1 accumulator -> ReLU (in the fp32->fp16 conversion) -> same TMEM columns (A of layer 2)
//   warps 4-7    epilogue 2: layer-2 accumulator -> softplus density + features, transposed through shared memory, coalesced
//                128-bit stores to the rows the perm ring names
//   warp  8      layer-1 MMA issuer + TMEM owner (per chunk 4 MMAs with an MN-major B descriptor, then 3 for the code block)
//   warp  9      layer-2 MMA issuer (8 TS-form MMAs + 1 for the bias block)
//   warp  10     producer: claims tiles (atomic, 4 per claim, last tile first), bulk-copies table entries and each tile's 4 KB
//                record block into rings, issues the TMA boxes
//   warps 11-14  one thread per row: weights into the chunk's A operand (an undo log keeps the rest of it zero), positional
//                code -> code operand, point index -> perm ring
// Rings: 2 weight chunks, 5 box chunks, 2 code operands, 3 record blocks, 8 perm blocks; layer-1 and layer-2 accumulators
// double buffered in TMEM.  Protocol rules (each was a deadlock first): an mbarrier parity wait tells apart only adjacent
// phases, so a role that skips ring positions still waits on each of them, and the closing arrivals at the end of the tile
// stream wait for the same conditions a real tile would.  DESIGN.md section 3.3 has the measurements behind the layout.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

// SD_TB_X3 = 1 (field_bin_x3.cu compiles this file a second time): the rel-1e-4 variant of the same kernel.  Every operand
// of every contraction is carried as an fp16 pair hi + lo (hi = half(v), lo = half(v - hi): ~22 significant bits) and every
// product as the three tensor-core products hi.hi + lo.hi + hi.lo, accumulated in fp32 in TMEM: bilinear weights x
// projected map (two maps P_hi / P_lo, sd_field_project_x3), positional code x code block of W_in, hidden x W_out.  The
// biases are added in fp32 in the epilogues, sin / cos start from an accurate sincosf, sigma uses the accurate softplus.
// Single-slot weight / code rings (shared memory: 2 x 16 KB per operand), rows leave by direct 16-byte stores.
#ifndef SD_TB_X3
#define SD_TB_X3 0
#endif
#if SD_TB_X3
#define TB_NS tbx
#else
#define TB_NS tb
#endif

namespace sd {
namespace TB_NS {
using namespace tcx;
constexpr bool X3 = SD_TB_X3 != 0;
constexpr int XK = X3 ? 2 : 1;                // operand images per ring entry: hi (+ lo)

constexpr int TM = 128;
constexpr int CHUNK = 16384;                 // A chunk [128 rows][64 slots] fp16 = B chunk [2 halves][64 slots][64 ch] fp16
#ifndef SD_TB_PT_GROUPS
#define SD_TB_PT_GROUPS 1
#endif
#ifndef SD_TB_NRA
#define SD_TB_NRA 2
#endif
#ifndef SD_TB_NRB
#define SD_TB_NRB 5
#endif
#ifndef SD_TB_STAGE
#define SD_TB_STAGE 4096      // staging bytes per epilogue-2 warp: one 32-column half at a time
#endif
#ifndef SD_TB_ABLATE
#define SD_TB_ABLATE 0        // timing experiments only (results are garbage): 1 no code computation, 2 no output path in epilogue 2,
#endif                        // 4 no box loads, 8 no layer-2 MMAs, 16 no chunk MMAs, 32 no epilogue-1 conversion
#ifndef SD_TB_BOX_AHEAD
#define SD_TB_BOX_AHEAD 1     // the producer issues the boxes of tile j + BOX_AHEAD while it publishes the records of tile j + 2
#endif
constexpr int NRA = X3 ? 1 : SD_TB_NRA, NRB = X3 ? 2 : SD_TB_NRB, NCODE = X3 ? 1 : 2;    // ring depths: weight (A) chunks, box (B) chunks, code operands
constexpr int N_EPI_WARPS = 4, N_PT_WARPS = 4, N_PT_GROUPS = SD_TB_PT_GROUPS;
constexpr int WARP_EPI2 = N_EPI_WARPS, WARP_MMA = 2 * N_EPI_WARPS, WARP_MMA2 = WARP_MMA + 1, WARP_TMA = WARP_MMA2 + 1,
              WARP_PT0 = WARP_TMA + 1;
constexpr int NTHREADS = (WARP_PT0 + N_PT_GROUPS * N_PT_WARPS) * 32;
static_assert(NCODE % N_PT_GROUPS == 0, "a code operand is always filled by the same point group");
constexpr int TMEM_COLS = 512;
constexpr int D2_COL = 256, D2_STRIDE = 128;  // layer-1 accumulators at columns 0 / 128, layer-2 at 256 / 384
constexpr int MAX_NVC_TB = 4;
constexpr int W2_BYTES = X3 ? 2 * 2 * 80 * 128 : 3 * 80 * 128;     // W_out: two K blocks + the bias block (x3: hi and lo images, no bias block)
constexpr int ONE_COL = 496;                 // 8 TMEM columns holding the constant (1, 1, 0, ...): A operand of the bias K step
constexpr int KCODE = 3;                     // K steps of the code block (48 columns)

constexpr int OFF_WC = 0;
constexpr int OFF_A = OFF_WC + XK * CHUNK;
constexpr int OFF_B = OFF_A + NRA * XK * CHUNK;
constexpr int OFF_CODE = OFF_B + NRB * XK * CHUNK;
constexpr int OFF_W2 = OFF_CODE + NCODE * XK * CHUNK;
constexpr int OFF_STAGE = OFF_W2 + W2_BYTES;
constexpr int STAGE_PER_WARP = X3 ? 0 : SD_TB_STAGE;
constexpr int NREC = 3;                      // ring of per-tile record blocks (128 x 32 B), filled by bulk copies
constexpr int REC_BYTES = TM * 32;
constexpr int NPERM = 8;                     // ring of per-tile point indices handed from the point warps to epilogue 2
constexpr int OFF_REC = OFF_STAGE + N_EPI_WARPS * STAGE_PER_WARP;
constexpr int OFF_PERM = OFF_REC + NREC * REC_BYTES;
constexpr int OFF_HDR = OFF_PERM + NPERM * TM * 4;        // per record-ring entry: the tile's table entry (TileInfo, 16 B)
#ifndef SD_TB_BATCH
#define SD_TB_BATCH 4
#endif
constexpr int BATCH = SD_TB_BATCH;                        // tiles claimed per atomic
constexpr int OFF_TAB = OFF_HDR + NREC * 16;              // two batches of table entries, fetched by bulk copies
constexpr int OFF_MINFO = OFF_TAB + 2 * BATCH * 16;       // per weight-ring entry: chunks of the tile whose first chunk sits there
constexpr int OFF_NTILES = OFF_MINFO + 16;      // tiles this CTA processed, published by the MMA issuer at the end
constexpr int OFF_TIDX = OFF_NTILES + 8;      // per record-ring entry: the tile's index in sorted order (binned output)
constexpr int OFF_PTILE = OFF_TIDX + 16;       // per perm-ring entry: the same, handed on to epilogue 2 by the point warps
constexpr int OFF_DIRTY = OFF_PTILE + NPERM * 4;
constexpr int OFF_CAM = OFF_DIRTY + NRA * TM + (NRA * TM % 8 ? 8 - NRA * TM % 8 : 0);
constexpr int OFF_BIAS = OFF_CAM + 448;            // x3: b_in [128] + b_out [80] fp32
constexpr int OFF_BAR = OFF_BIAS + (X3 ? 832 : 0);
enum { BAR_FULL_A = 0, BAR_EMPTY_A = NRA, BAR_FULL_B = 2 * NRA, BAR_EMPTY_B = 2 * NRA + NRB, BAR_FULL_C = 2 * NRA + 2 * NRB,
       BAR_EMPTY_C = BAR_FULL_C + NCODE,
       BAR_D1 = BAR_EMPTY_C + NCODE, BAR_H = BAR_D1 + 2, BAR_D2 = BAR_H + 2, BAR_D2_EMPTY = BAR_D2 + 2,
       BAR_WLOAD = BAR_D2_EMPTY + 2, BAR_REC_FULL = BAR_WLOAD + 1, BAR_REC_EMPTY = BAR_REC_FULL + NREC,
       BAR_TAB = BAR_REC_EMPTY + NREC, BAR_D1_FREE = BAR_TAB + 2, NBAR = BAR_D1_FREE + 2 };
// consumers of a record-ring entry: the point warps
constexpr int REC_CONSUMERS = N_PT_GROUPS * N_PT_WARPS;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_ALLOC = OFF_TMEM + 16 + 1024;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");
static_assert(SD_TB_BOX_AHEAD >= 0 && SD_TB_BOX_AHEAD <= 2, "headers live in a ring of NREC = 3");
static_assert(OFF_W2 % 1024 == 0 && OFF_STAGE % 1024 == 0 && OFF_BAR % 8 == 0 && NREC * 4 <= 16, "alignment");

#ifndef TB_T0
#define TB_T0 40   // first traced tile of CTA 0 (SD_TC_DEBUG & 8192): tiles TB_T0 .. TB_T0 + 63
#endif
__device__ long long g_trace[8 * 64 * 8];
__device__ long long g_tiles[4 * 2 * 512];          // [cta 0..3][tile][clock64 at the layer-1 issuer's tile start, chunks] (SD_TC_DEBUG & 8192)
__device__ unsigned long long g_cta_ns[256 * 2];   // [cta][start, end] %globaltimer (SD_TC_DEBUG & 8192)   // [role][tile][event] clock64 stamps of CTA 0 (SD_TC_DEBUG & 8192)
#ifdef SD_DEBUG_WAIT
// debug build: last stage every warp reached, frozen when the first wait times out (profiles/stress_bin.py prints it)
__device__ int g_prog[160 * 16 * 4];
#define TB_PROG(jj, stage, a1, a2)                                                                               \
    do {                                                                                                          \
        if ((threadIdx.x & 31) == 0 && *reinterpret_cast<volatile unsigned int *>(&tcx::g_wait_timeout[0]) == 0) { \
            int *pp_ = g_prog + (blockIdx.x * 16 + (threadIdx.x >> 5)) * 4;                                       \
            pp_[0] = (int)(jj); pp_[1] = (stage); pp_[2] = (int)(a1); pp_[3] = (int)(a2);                           \
        }                                                                                                         \
    } while (0)
#else
#define TB_PROG(jj, stage, a1, a2) do {} while (0)
#endif
#ifdef SD_TB_JITTER
// debug build: random delays in every role (per warp, from the clock) to shake out protocol races
#define TB_JIT()                                                                                     \
    do {                                                                                             \
        unsigned int h_ = (unsigned int)clock64() * 2654435761u + (threadIdx.x >> 5) * 40503u;         \
        h_ = __shfl_sync(0xffffffffu, h_, 0);                                                         \
        if ((h_ & 0x30000u) == 0) __nanosleep((h_ >> 20) & 0xFFFu);                                    \
    } while (0)
#else
#define TB_JIT() do {} while (0)
#endif
#define TB_TRACE(role, j, ev)                                                                    \
    do {                                                                                         \
        if ((P.dbg & 8192) && blockIdx.x == 0 && (j) >= TB_T0 && (j) < TB_T0 + 64 && (role) < 8 && (threadIdx.x & 31) == 0) \
            g_trace[((role) * 64 + (int)(j) - TB_T0) * 8 + (ev)] = clock64();                            \
    } while (0)

struct Params {
    CUtensorMap tmap;          // P as [Hf][Wf][128] fp16, box 8 x 8 x 64 channels, SWIZZLE_128B
    CUtensorMap tmap_out;      // binned output: dino as [N][64] fp32, box 32 rows x 32 columns, SWIZZLE_128B
    CUtensorMap tmap_lo;       // x3: the lo half of the projected map
    const float *b_in;         // x3: fp32 biases, added in the epilogues
    FieldParams fp;
    const float *xyz;
    unsigned int *tile_ctr;    // next unclaimed tile (zeroed by the sort): CTAs claim tiles dynamically
    const TileInfo *tiles;     // per-tile table in claim order (binning.cu), padded with empty entries
    const GeoRec *rec;         // per-point records at the sorted positions (binning.cu)
    const unsigned int *cbin;  // compact bin number -> bin id
    int nbx;
    int dbg;                   // SD_TC_DEBUG & 8192: clock64 trace of CTA 0
    long long N, n_tiles;
    int n2, D;
    const unsigned char *wc_img, *w2_img;
    const float *b_out;
    float *sigma, *dino, *rgb, *invalid;
    unsigned char *invalid_feat;
    int binned;                // feature rows leave in SORTED order (row r of tile t -> dino[t * 128 + r]) by TMA tile stores
    unsigned int *perm_out;    // binned: [N] sorted position -> point index (or NULL)
};

__global__ void __launch_bounds__(NTHREADS, 1) field_bin_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_u = smem_u32(sm);
    const int tid = threadIdx.x, warp = warp_uniform(), lane = tid & 31;
    const uint32_t bar0 = sm_u + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    float *s_cam = reinterpret_cast<float *>(sm + OFF_CAM);
    unsigned char *s_dirty = sm + OFF_DIRTY;

    // ---- one-time setup ------------------------------------------------------------------------------
    if ((P.dbg & 8192) && blockIdx.x < 256 && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_cta_ns[2 * blockIdx.x] = t; g_cta_ns[2 * blockIdx.x + 1] = 0;
    }
    if ((P.dbg & 8192) && blockIdx.x < 4 && tid == 0) g_tiles[(blockIdx.x * 512 + 510) * 2] = clock64();
    if (tid == 0) {
        for (int e = 0; e < NRA; ++e) { mbar_init(BAR(BAR_FULL_A + e), N_PT_WARPS); mbar_init(BAR(BAR_EMPTY_A + e), 1); }
        for (int e = 0; e < NRB; ++e) { mbar_init(BAR(BAR_FULL_B + e), 1); mbar_init(BAR(BAR_EMPTY_B + e), 1); }
        for (int s = 0; s < NCODE; ++s) { mbar_init(BAR(BAR_FULL_C + s), N_PT_WARPS); mbar_init(BAR(BAR_EMPTY_C + s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(BAR_D1 + b), 1); mbar_init(BAR(BAR_H + b), N_EPI_WARPS);
            mbar_init(BAR(BAR_D2 + b), 1); mbar_init(BAR(BAR_D2_EMPTY + b), N_EPI_WARPS);
        }
        mbar_init(BAR(BAR_WLOAD), 1);
        for (int r = 0; r < NREC; ++r) { mbar_init(BAR(BAR_REC_FULL + r), 1); mbar_init(BAR(BAR_REC_EMPTY + r), REC_CONSUMERS); }
        mbar_init(BAR(BAR_TAB), 1); mbar_init(BAR(BAR_TAB + 1), 1);
        mbar_init(BAR(BAR_D1_FREE), 1); mbar_init(BAR(BAR_D1_FREE + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 21 * (1 + P.fp.nv_c); i += NTHREADS) {
        const int c = i / 21, e = i - 21 * c;
        const float *K = c == 0 ? P.fp.K_f : P.fp.K_c + 9 * (c - 1);
        const float *W = c == 0 ? P.fp.w2c_f : P.fp.w2c_c + 16 * (c - 1);
        s_cam[i] = e < 9 ? __ldg(K + e) : __ldg(W + (e - 9));
    }
    // the weight operands start out all zero and are kept so by the undo log of the point warps
    for (int i = tid; i < NRA * XK * CHUNK / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm + OFF_A)[i] = make_uint4(0, 0, 0, 0);
    if (X3) {   // fp32 biases for the epilogues; the code operands start out zero (their unused columns stay zero)
        float *s_bias = reinterpret_cast<float *>(sm + OFF_BIAS);
        for (int i = tid; i < 128 + 80; i += NTHREADS) s_bias[i] = i < 128 ? __ldg(P.b_in + i) : (i - 128 <= P.D ? __ldg(P.b_out + (i - 128)) : 0.0f);
        for (int i = tid; i < NCODE * XK * CHUNK / 16; i += NTHREADS) reinterpret_cast<uint4 *>(sm + OFF_CODE)[i] = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < NRA * TM; i += NTHREADS) s_dirty[i] = 0xFF;
    if (tid == 0) *reinterpret_cast<volatile int *>(sm + OFF_NTILES) = 0x7FFFFFFF;
    fence_proxy_async();
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u + OFF_TMEM), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t w2b = (X3 ? 4u : 3u) * (uint32_t)P.n2 * 128u;
        mbar_expect_tx(BAR(BAR_WLOAD), XK * CHUNK + w2b);
        bulk_g2s(sm_u + OFF_WC, P.wc_img, XK * CHUNK, BAR(BAR_WLOAD));
        bulk_g2s(sm_u + OFF_W2, P.w2_img, w2b, BAR(BAR_WLOAD));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);

    // Tiles are claimed dynamically by the TMA producer (atomic counter): the time a tile takes varies with the chunks it
    // touches and, more, from SM to SM (measured: static striding left CTAs finishing between 218 and 335 us).  The j-th
    // tile of this CTA travels through record-ring entry j % NREC: header {rows, chunks}, then 128 records.
    volatile TileInfo *s_hdr = reinterpret_cast<volatile TileInfo *>(sm + OFF_HDR);
    volatile int *s_ntiles = reinterpret_cast<volatile int *>(sm + OFF_NTILES);
    auto rec_wait = [&](long long jj) { mbar_wait(BAR(BAR_REC_FULL + (int)(jj % NREC)), (uint32_t)((jj / NREC) & 1)); };
    auto rec_ptr = [&](long long jj) { return reinterpret_cast<const GeoRec *>(sm + OFF_REC + (int)(jj % NREC) * REC_BYTES); };

    if (warp < N_EPI_WARPS) {
        // =================================== EPILOGUE 1 ==============================================
        // layer-1 accumulator -> ReLU -> fp16 -> the same TMEM columns, A operand of layer 2
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        if (!X3) {   // constant-1 columns (k = 128, 129 of layer 2): the output bias comes out of the MMA
            uint32_t one[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(t_lane + ONE_COL, one);
        }
        for (long long j = 0;; ++j) {
            const int b = (int)(j & 1);
            mbar_wait(BAR(BAR_D1 + b), (uint32_t)((j >> 1) & 1));
            if (j >= *s_ntiles) {                         // the layer-1 issuer's closing arrival, not a tile: pass it on
                mbar_arrive_warp(BAR(BAR_H + b));
                break;
            }
            tc_fence_after();
            if (warp == 0) TB_TRACE(0, j, 2);
#pragma unroll 1
            for (int kb = (SD_TB_ABLATE & 32) ? 4 : 0; kb < 4; ++kb) {      // 32 hidden units -> 16 packed columns, written over columns already read
                uint32_t vr[32];
                tmem_ld32_issue(t_lane + b * 128 + kb * 32, vr);
                tmem_ld_wait();
                uint32_t pk[16];
                if (X3) {
                    // h = relu(acc + b_in) in fp32 -> (hi, lo) halves: hi pairs over columns 32 kb .. +15, lo pairs over
                    // 32 kb + 16 .. +31 (both inside the 32 columns just read)
                    const float *s_b1 = reinterpret_cast<const float *>(sm + OFF_BIAS) + kb * 32;
                    uint32_t pl[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float h0 = fmaxf(__uint_as_float(vr[2 * e]) + s_b1[2 * e], 0.0f), h1 = fmaxf(__uint_as_float(vr[2 * e + 1]) + s_b1[2 * e + 1], 0.0f);
                        const __half2 hi = __floats2half2_rn(h0, h1);
                        const float2 hf = __half22float2(hi);
                        pk[e] = as_u32(hi);
                        pl[e] = pack_h2(h0 - hf.x, h1 - hf.y);
                    }
                    tmem_st16(t_lane + b * 128 + kb * 32, pk);
                    tmem_st16(t_lane + b * 128 + kb * 32 + 16, pl);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        pk[e] = pack_h2_relu(__uint_as_float(vr[2 * e]), __uint_as_float(vr[2 * e + 1]));
                    tmem_st16(t_lane + b * 128 + kb * 16, pk);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            TB_JIT();
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
        unsigned char *stage0 = sm + OFF_STAGE + wq * STAGE_PER_WARP, *stage1 = stage0 + (STAGE_PER_WARP == 8192 ? 4096 : 0);
        // The point index of this thread's row was left in the perm ring by the point warps when they processed the tile.
        // No barrier of its own: the write is ordered before this read by the chain FULL_C -> MMA -> D1 -> H -> D2, and
        // an entry is rewritten 8 tiles later, while the point warps can be at most 5 tiles ahead of this role (two
        // code operands, layer 2 behind layer 1, two D2 buffers).
        const int *s_perm = reinterpret_cast<const int *>(sm + OFF_PERM);
        for (long long j = 0;; ++j) {
            const int b1 = (int)(j & 1);
            mbar_wait(BAR(BAR_D2 + b1), (uint32_t)((j >> 1) & 1));
            if (j >= *s_ntiles) break;                    // the MMA issuer's closing arrival, not a tile
            tc_fence_after();
            if (wq == 0) TB_TRACE(0, j, 0);
#ifdef SD_TB_SLOW_EPI2
            __nanosleep(SD_TB_SLOW_EPI2);        // debug: sustained back-pressure from the last role of the chain
#endif
            const int grow_keep = s_perm[(int)(j % NPERM) * TM + row];       // point this thread's row of tile j stands for
            const uint32_t t_d2 = t_lane + D2_COL + b1 * D2_STRIDE;
            const bool ok = grow_keep >= 0;
            if (X3) {
                // fp32 biases, accurate softplus; the row leaves by direct 16-byte stores (its own row of dino, in the
                // caller's order or -- binned -- at its sorted position)
                const float *s_b2 = reinterpret_cast<const float *>(sm + OFF_BIAS) + 128;    // b_out in nn.Linear order: [sigma, f0 ...]
                const int t = reinterpret_cast<const volatile int *>(sm + OFF_PTILE)[(int)(j % NPERM)];
                uint32_t vr[64], sr;
                tmem_ld32_issue(t_d2, vr);
                tmem_ld32_issue(t_d2 + 32, vr + 32);
                tmem_ld1_issue(t_d2 + D, sr);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_warp(BAR(BAR_D2_EMPTY + b1));
                if (ok) {
                    if (P.sigma) P.sigma[grow_keep] = softplus(__uint_as_float(sr) + s_b2[0]);
                    if (P.binned && P.perm_out) P.perm_out[(long long)t * TM + row] = (unsigned int)grow_keep;
                    if (P.dino) {
                        float *o = P.dino + (P.binned ? (long long)t * TM + row : (long long)grow_keep) * D;
                        if (D == 64) {
#pragma unroll
                            for (int q = 0; q < 16; ++q)
                                reinterpret_cast<float4 *>(o)[q] = make_float4(__uint_as_float(vr[4 * q]) + s_b2[1 + 4 * q], __uint_as_float(vr[4 * q + 1]) + s_b2[2 + 4 * q],
                                                                               __uint_as_float(vr[4 * q + 2]) + s_b2[3 + 4 * q], __uint_as_float(vr[4 * q + 3]) + s_b2[4 + 4 * q]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 64; ++c)
                                if (c < D) o[c] = __uint_as_float(vr[c]) + s_b2[1 + c];
                        }
                    }
                }
            } else if (P.binned) {
                // Binned output: the tile's rows are consecutive rows of dino, so the features leave as four TMA tile stores
                // per warp-quadrant pair (32 rows x 128 B each) straight from the staging buffer -- written once, swizzled
                // the way the tensor map expects (16-byte piece q of row r at q ^ (r & 7)), never read back, no address per
                // row: nothing of the output queues in the load/store unit besides these 16 shared-memory stores per thread.
                // Rows past N (ragged last tile) fall outside the tensor map and are clipped by the copy engine.
                const int t = reinterpret_cast<const volatile int *>(sm + OFF_PTILE)[(int)(j % NPERM)];
                uint32_t sr;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t vr[32];
                    tmem_ld32_issue(t_d2 + hf * 32, vr);
                    if (hf == 1) tmem_ld1_issue(t_d2 + D, sr);
                    tmem_ld_wait();
                    if (hf == 1) {
                        tc_fence_before();
                        mbar_arrive_warp(BAR(BAR_D2_EMPTY + b1));
                    }
                    if (SD_TB_ABLATE & 2) continue;
                    if (elect_one()) bulk_wait_read<(SD_TB_STAGE == 8192 ? 1 : 0)>();     // the store that last read this half (a tile ago) has left
                    __syncwarp();
                    unsigned char *stage = hf ? stage1 : stage0;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<uint4 *>(stage + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                            make_uint4(vr[4 * q], vr[4 * q + 1], vr[4 * q + 2], vr[4 * q + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (elect_one()) {          // (the same lane every time: bulk groups are per thread)
                        tma_store_2d(&P.tmap_out, smem_u32(stage), hf * 32, t * TM + wq * 32);
                        bulk_commit();
                    }
                }
                if (ok && !(SD_TB_ABLATE & 2)) {
                    if (P.sigma && !(P.dbg & 2)) P.sigma[grow_keep] = softplus_fast(__uint_as_float(sr));
                    if (P.perm_out) P.perm_out[(long long)t * TM + row] = (unsigned int)grow_keep;
                }
            } else if (P.dino && D == 64) {
                // Two halves of 32 columns (registers), one after the other through ONE 4 KB staging buffer per warp (the
                // other 4 KB went to a fifth box slot): lane = row writes its 8 chunks of a half (XOR-swizzled, conflict
                // free), then 8 lanes read one row back and the warp stores four 128-byte row halves per request.  (One
                // 256-byte bulk copy per row was measured instead: ~22 cycles per copy, slower.)
                uint32_t sr;
                const int c8 = lane & 7;
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t vr[32];
                    tmem_ld32_issue(t_d2 + hf * 32, vr);
                    if (hf == 1) tmem_ld1_issue(t_d2 + D, sr);                   // density column sits behind the features
                    tmem_ld_wait();
                    if (hf == 1) {                                               // both halves are in registers
                        tc_fence_before();
                        TB_JIT();
                        mbar_arrive_warp(BAR(BAR_D2_EMPTY + b1));
                    }
                    __syncwarp();                   // every lane has finished reading the previous half back
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<uint4 *>(stage0 + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                            make_uint4(vr[4 * q], vr[4 * q + 1], vr[4 * q + 2], vr[4 * q + 3]);
                    __syncwarp();                   // the staged rows are visible
                    // all the shared loads of the half first, into distinct registers (volatile asm keeps the order): a load
                    // that reuses the source registers of a store in flight waits for that store to leave the LSU
                    float4 x[8];
                    float *dp[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + (lane >> 3);
                        x[i] = lds128_ordered(smem_u32(stage0) + r * 128 + ((c8 ^ (r & 7)) << 4));
                        const int dst = __shfl_sync(0xffffffffu, grow_keep, r);
                        dp[i] = dst >= 0 ? P.dino + (long long)dst * 64 + hf * 32 + c8 * 4 : nullptr;
                    }
                    __syncwarp();   // (also keeps ptxas from sinking the loads back between the stores)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (dp[i] && !(P.dbg & 1)) stg128_ordered(dp[i], x[i]);
                }
                if (wq == 0) TB_TRACE(0, j, 4);
                if (ok && P.sigma && !(P.dbg & 2)) P.sigma[grow_keep] = softplus_fast(__uint_as_float(sr));
                if (wq == 0) TB_TRACE(0, j, 5);
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
        // Layer 1, issued by the whole warp in lockstep (tc_common.cuh: "elected" forms -- one lane issues, nothing
        // diverges, the tcgen05 instructions come out back to back).  Layer 2 has a warp of its own (WARP_MMA2): every
        // barrier test is a round trip through the load/store unit and one warp doing all of them was the critical path.
        {
            mbar_wait_warp(BAR(BAR_WLOAD), 0);
            const uint32_t idesc_k = umma_idesc(TM, 128), idesc_mn = idesc_k | UMMA_B_MN_MAJOR;
            int e = 0, eb = 0;                       // ring positions: weight chunks, box chunks
            uint32_t ph = 0, phb = 0;
            long long j = 0;
            volatile int *s_minfo = reinterpret_cast<volatile int *>(sm + OFF_MINFO);
            for (;; ++j) {
                // The tile's first weight chunk carries its chunk count (0: no more tiles) and stands for the code operand
                // too: one barrier test where there were three.
                mbar_wait_warp(BAR(BAR_FULL_A + e), ph);
                const int m = __shfl_sync(0xffffffffu, s_minfo[e], 0);
                TB_PROG(j, 1, m, e);
                // the accumulator (its first 64 columns held the hidden tile of tile j-2) is free once layer 2 of j-2 has run.
                // (Also before the closing arrival below: nobody may complete two phases of a barrier ahead of its waiter.)
                mbar_wait_warp(BAR(BAR_D1_FREE + (int)(j & 1)), (uint32_t)(((j >> 1) & 1) ^ 1));
                if (m == 0) break;
                const uint32_t d1 = tmem_base + (uint32_t)(j & 1) * 128u;
                TB_TRACE(1, j, 0);
                if ((P.dbg & 8192) && blockIdx.x < 4 && j < 500 && lane == 0) { g_tiles[(blockIdx.x * 512 + j) * 2] = clock64(); g_tiles[(blockIdx.x * 512 + j) * 2 + 1] = m; }
                for (int i = 0; i < m; ++i) {
                    TB_PROG(j, 10 + i, m, e);
                    if (i > 0) mbar_wait_warp(BAR(BAR_FULL_A + e), ph);
                    if (i == 0) TB_TRACE(1, j, 1);
                    mbar_wait_warp(BAR(BAR_FULL_B + eb), phb);
                    tc_fence_after();
                    if (i == 0) TB_TRACE(1, j, 2);
                    TB_JIT();
                    // ONE elected thread issues (a single-thread region: even a warp that has split cannot issue twice)
                    if (elect_one()) {
#pragma unroll
                        for (int k = (SD_TB_ABLATE & 16) ? 4 : 0; k < 4; ++k) {    // 16 texel slots per instruction
                            const uint32_t a_h = sm_u + OFF_A + e * XK * CHUNK + k * 32, b_h = sm_u + OFF_B + eb * XK * CHUNK + k * 2048;
                            umma(d1, umma_desc(a_h), umma_desc_mn(b_h, CHUNK / 2, 1024), idesc_mn, (i | k) != 0);
                            if (X3) {       // + w_lo . P_hi + w_hi . P_lo
                                umma(d1, umma_desc(a_h + CHUNK), umma_desc_mn(b_h, CHUNK / 2, 1024), idesc_mn, 1);
                                umma(d1, umma_desc(a_h), umma_desc_mn(b_h + CHUNK, CHUNK / 2, 1024), idesc_mn, 1);
                            }
                        }
                        umma_commit(BAR(BAR_EMPTY_A + e));
                        umma_commit(BAR(BAR_EMPTY_B + eb));
                    }
                    __syncwarp();
                    if (++e == NRA) { e = 0; ph ^= 1; }
                    if (++eb == NRB) { eb = 0; phb ^= 1; }
                }
                const int cs = (int)(j % NCODE);
                TB_TRACE(1, j, 4);
                TB_JIT();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KCODE; ++k) {
                        const uint32_t c_h = sm_u + OFF_CODE + cs * XK * CHUNK + k * 32, w_h = sm_u + OFF_WC + k * 32;
                        umma(d1, umma_desc(c_h), umma_desc(w_h), idesc_k, 1);
                        if (X3) {
                            umma(d1, umma_desc(c_h + CHUNK), umma_desc(w_h), idesc_k, 1);
                            umma(d1, umma_desc(c_h), umma_desc(w_h + CHUNK), idesc_k, 1);
                        }
                    }
                    umma_commit(BAR(BAR_EMPTY_C + cs));
                    umma_commit(BAR(BAR_D1 + (int)(j & 1)));
                }
                __syncwarp();
                TB_TRACE(1, j, 6);
            }
            // closing: tell everybody downstream how many tiles there were and complete the phase the first epilogue waits
            // on; it passes the arrival on to the layer-2 issuer, which passes it on to the second epilogue
            if (lane == 0) *s_ntiles = (int)j;
            __syncwarp();
            mbar_arrive_e(BAR(BAR_D1 + (int)(j & 1)));
        }
    } else if (warp == WARP_MMA2) {
        // =================================== MMA ISSUER, LAYER 2 ======================================
        {   // (whole warp, elected forms: see the layer-1 issuer)
            mbar_wait_warp(BAR(BAR_WLOAD), 0);
            const uint32_t idesc2 = umma_idesc(TM, P.n2);
            for (long long j = 0;; ++j) {
                const int b = (int)(j & 1);
                mbar_wait_warp(BAR(BAR_H + b), (uint32_t)((j >> 1) & 1));
                // (the accumulator drained -- also before the closing arrival: nobody may complete two phases of a barrier
                // ahead of its waiter)
                mbar_wait_warp(BAR(BAR_D2_EMPTY + b), (uint32_t)(((j >> 1) & 1) ^ 1));
                const int nt = __shfl_sync(0xffffffffu, *s_ntiles, 0);
                if (j >= nt) { mbar_arrive_e(BAR(BAR_D2 + b)); break; }
                tc_fence_after();
                TB_TRACE(1, j + 1, 5);
                TB_JIT();
                if (elect_one()) {
                    if (X3) {
                        // hidden (hi, lo) pairs sit in blocks of 16 columns per 32 units (epilogue 1): k-step k reads hi at
                        // 32 (k / 2) + 8 (k % 2), lo 16 columns further; W_out hi image, then the lo image 2 n2 128 bytes behind
                        const uint32_t w2l = 2u * (uint32_t)P.n2 * 128u;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t a_h = tmem_base + b * 128 + (k >> 1) * 32 + (k & 1) * 8;
                            const uint32_t w_h = sm_u + OFF_W2 + (k >> 2) * P.n2 * 128 + (k & 3) * 32;
                            umma_ts(tmem_base + D2_COL + b * D2_STRIDE, a_h, umma_desc(w_h), idesc2, k != 0);
                            umma_ts(tmem_base + D2_COL + b * D2_STRIDE, a_h + 16, umma_desc(w_h), idesc2, 1);
                            umma_ts(tmem_base + D2_COL + b * D2_STRIDE, a_h, umma_desc(w_h + w2l), idesc2, 1);
                        }
                    } else {
#pragma unroll
                        for (int k = (SD_TB_ABLATE & 8) ? 8 : 0; k < 8; ++k)          // K = 16 per instruction = 8 packed columns of the hidden tile
                            umma_ts(tmem_base + D2_COL + b * D2_STRIDE, tmem_base + b * 128 + k * 8,
                                    umma_desc(sm_u + OFF_W2 + (k >> 2) * P.n2 * 128 + (k & 3) * 32), idesc2, k != 0);
                        umma_ts(tmem_base + D2_COL + b * D2_STRIDE, tmem_base + ONE_COL, umma_desc(sm_u + OFF_W2 + 2 * P.n2 * 128), idesc2, 1);
                    }
                    umma_commit(BAR(BAR_D2 + b));
                    umma_commit(BAR(BAR_D1_FREE + b));
                }
                __syncwarp();
            }
        }
    } else if (warp == WARP_TMA) {
        // =================================== TMA PRODUCER =============================================
        // The whole warp runs the role in lockstep and the copies are issued in their "elected" forms (tc_common.cuh): no
        // per-instruction ELECT loop.
        if (lane == 0) { tma_prefetch_desc(&P.tmap); if (X3) tma_prefetch_desc(&P.tmap_lo); }
        __syncwarp();
        int e = 0;
        uint32_t ph = 0;
        // Tiles are claimed BATCH at a time with one atomic; the table entries of a batch (rows, chunk span, first bins: all
        // the producer needs to know about a tile) arrive by a bulk copy.  In steady state this warp issues no load through
        // the load/store unit: a claim and a table fetch have a whole batch of tiles to complete.  Per iteration: publish
        // header + records of tile j+2 (the point warps run ahead of the MMAs), then issue the boxes of tile j + BOX_AHEAD
        // (a box takes ~2 400 cycles to arrive -- a tile period -- so they are requested as far ahead as the ring allows).
        const TileInfo *s_tab = reinterpret_cast<const TileInfo *>(sm + OFF_TAB);
        auto claim = [&]() -> long long { return lane == 0 ? (long long)atomicAdd(P.tile_ctr, (unsigned)BATCH) : 0; };
        auto bcast = [&](long long v) { return __shfl_sync(0xffffffffu, v, 0); };
        auto fetch_tab = [&](long long k, long long base) {        // batch k -> table slot k & 1
            if (base >= P.n_tiles) return;
            if (elect_one()) {
                mbar_expect_tx(BAR(BAR_TAB + (int)(k & 1)), BATCH * 16);
                bulk_g2s(sm_u + OFF_TAB + (int)(k & 1) * BATCH * 16, P.tiles + base, BATCH * 16, BAR(BAR_TAB + (int)(k & 1)));
            }
            __syncwarp();
        };
        // stream of this CTA's tiles: batch k, entry i
        long long k = 0, base = bcast(claim()), base_n;
        int bi = 0;
        fetch_tab(0, base);
        base_n = claim();                                           // (lane 0; broadcast when batch 1 starts)
        bool ended = false, tab_ready = false;
        auto publish_next = [&](long long slot) {                   // next tile of the stream -> ring entry slot % NREC
            if (ended) return;
            const int r = (int)(slot % NREC);
            TileInfo ti;
            ti.c0m = 0; ti.b01 = 0; ti.b23 = 0; ti.rows = 0;
            long long v = -1;                                       // position in claim order
            if (base < P.n_tiles) {
                if (!tab_ready) { mbar_wait_warp(BAR(BAR_TAB + (int)(k & 1)), (uint32_t)((k >> 1) & 1)); tab_ready = true; }
                const TileInfo tl = s_tab[(int)(k & 1) * BATCH + bi];      // (broadcasts: the compiler must SEE that the
                ti.c0m = __shfl_sync(0xffffffffu, tl.c0m, 0);               // control flow below is uniform, or every copy gets
                ti.b01 = __shfl_sync(0xffffffffu, tl.b01, 0);               // its ELECT loop back)
                ti.b23 = __shfl_sync(0xffffffffu, tl.b23, 0);
                ti.rows = __shfl_sync(0xffffffffu, tl.rows, 0);
                v = base + bi;
            }
            TB_PROG(slot, 1, ti.rows, ti.c0m);
            TB_JIT();
            mbar_wait_warp(BAR(BAR_REC_EMPTY + r), (uint32_t)(((slot / NREC) & 1) ^ 1));
            TB_PROG(slot, 2, ti.rows, v);
            if (lane == 0) {
                s_hdr[r].c0m = ti.c0m; s_hdr[r].b01 = ti.b01; s_hdr[r].b23 = ti.b23; s_hdr[r].rows = ti.rows;
                reinterpret_cast<volatile int *>(sm + OFF_TIDX)[r] = (int)(P.n_tiles - 1 - v);
            }
            __syncwarp();                                           // (the header is written before whichever lane arrives)
            if (ti.rows == 0) {
                mbar_arrive_e(BAR(BAR_REC_FULL + r));               // no more tiles: a header alone
                ended = true;
                return;
            }
            const long long t = P.n_tiles - 1 - v;                  // claim order runs from the last tile down
            if (elect_one()) {
                mbar_expect_tx(BAR(BAR_REC_FULL + r), ti.rows * 32u);
                bulk_g2s(sm_u + OFF_REC + r * REC_BYTES, P.rec + t * TM, ti.rows * 32u, BAR(BAR_REC_FULL + r));
            }
            __syncwarp();
            if (++bi == BATCH) {                                    // next batch: its claim was made a batch ago
                bi = 0; ++k; tab_ready = false;
                base = bcast(base_n);
                fetch_tab(k, base);
                base_n = claim();
            }
        };
        publish_next(0);
        publish_next(1);
        // boxes of tile j: from the header published for it (false: past the end)
        auto issue_boxes = [&](long long j) -> bool {
            const int r = (int)(j % NREC);
            const unsigned int rows = __shfl_sync(0xffffffffu, s_hdr[r].rows, 0), c0m = __shfl_sync(0xffffffffu, s_hdr[r].c0m, 0),
                               b01 = __shfl_sync(0xffffffffu, s_hdr[r].b01, 0), b23 = __shfl_sync(0xffffffffu, s_hdr[r].b23, 0);
            if (rows == 0) return false;
            const int c0 = (int)(c0m & 0xFFFFu), m = (int)(c0m >> 16);
            TB_PROG(j, 3, rows, m);
            for (int i = 0; i < m; ++i) {
                unsigned int b = i == 0 ? (b01 & 0xFFFFu) : i == 1 ? (b01 >> 16) : i == 2 ? (b23 & 0xFFFFu) : (b23 >> 16);
                if (i >= 4) b = __shfl_sync(0xffffffffu, __ldg(P.cbin + c0 + i), 0);   // a tile that touches more than four bins (rare)
                const int by = (int)(b / (unsigned)P.nbx), bx = (int)(b - (unsigned)by * (unsigned)P.nbx);
                if (i == 0) TB_TRACE(6, j, 0);
                mbar_wait_warp(BAR(BAR_EMPTY_B + e), ph ^ 1);
                if (i == 0) TB_TRACE(6, j, 1);
                if (SD_TB_ABLATE & 4) {
                    mbar_arrive_e(BAR(BAR_FULL_B + e));
                } else if (elect_one()) {
                    mbar_expect_tx(BAR(BAR_FULL_B + e), XK * CHUNK);
                    const uint32_t dst = sm_u + OFF_B + e * XK * CHUNK;
                    tma_load_3d(dst, &P.tmap, 0, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                    tma_load_3d(dst + CHUNK / 2, &P.tmap, 64, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                    if (X3) {
                        tma_load_3d(dst + CHUNK, &P.tmap_lo, 0, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                        tma_load_3d(dst + CHUNK + CHUNK / 2, &P.tmap_lo, 64, bx * SD_BIN, by * SD_BIN, BAR(BAR_FULL_B + e));
                    }
                }
                __syncwarp();
                if (++e == NRB) { e = 0; ph ^= 1; }
            }
            TB_TRACE(6, j, 2);
            if ((P.dbg & 8192) && blockIdx.x == 0 && j >= TB_T0 && j < TB_T0 + 64 && lane == 0) g_trace[(6 * 64 + (int)j - TB_T0) * 8 + 7] = m;
            return true;
        };
        bool more = true;
        for (int a = 0; a < SD_TB_BOX_AHEAD && more; ++a) more = issue_boxes(a);
        for (long long j = 0; more; ++j) {
            publish_next(j + 2);
            more = issue_boxes(j + SD_TB_BOX_AHEAD);
        }
    } else {
        // =================================== POINT WARPS ================================================
        // Two groups of four warps take alternate tiles (group g: tiles g, g+2, ...), one thread per row.  Both count
        // the chunks of ALL tiles, so that the ring positions agree with the MMA issuer's.
        const int grp = (warp - WARP_PT0) / N_PT_WARPS;
        const int row = tid - (WARP_PT0 + grp * N_PT_WARPS) * 32;
        const int nv_c = P.fp.nv_c;
        const int pt_role = 2 + (warp - WARP_PT0) % N_PT_WARPS + (grp ? 8 : 0);   // trace: group 0 only (slots 2..5)
        // inputs of a tile row: its record in the ring (shared memory: nothing here queues behind the epilogue's store
        // bursts in the global load/store path) and the tile's chunk span
        struct RowIn { bool ok; int cr, c0, c1, slot, grow; float x, y, zp; uint32_t w01, w23; };
        volatile int *s_minfo = reinterpret_cast<volatile int *>(sm + OFF_MINFO);
        int e = 0;
        uint32_t ph = 0;
        for (long long j = 0;; ++j) {
            RowIn cur;
            rec_wait(j);
            const int rows = (int)s_hdr[j % NREC].rows;
            TB_PROG(j, 1, rows, e);
            if (rows == 0) {                               // no more tiles for this CTA: tell the MMA issuer (chunk count 0)
                if ((int)(j % N_PT_GROUPS) == grp) {
                    mbar_wait(BAR(BAR_EMPTY_A + e), ph ^ 1);
                    if (row == 0) s_minfo[e] = 0;
                    mbar_arrive_warp(BAR(BAR_FULL_A + e));
                }
                break;
            }
            const GeoRec *rr = rec_ptr(j);
            const int tile_idx = reinterpret_cast<const volatile int *>(sm + OFF_TIDX)[(int)(j % NREC)];
            cur.c0 = (int)(rr[0].cs & 0xFFFFu);            // compact bins the tile touches: first, last
            cur.c1 = (int)(rr[rows - 1].cs & 0xFFFFu);
            const bool mine = (int)(j % N_PT_GROUPS) == grp;
            cur.ok = mine && row < rows;
            cur.cr = 0; cur.slot = 0xFF; cur.grow = -1; cur.x = cur.y = cur.zp = 0.0f; cur.w01 = cur.w23 = 0u;
            if (cur.ok) {
                const uint4 *rp = reinterpret_cast<const uint4 *>(rr + row);
                const uint4 ra = rp[0], rb = rp[1];
                cur.x = __uint_as_float(ra.x); cur.y = __uint_as_float(ra.y); cur.zp = __uint_as_float(ra.z);
                cur.w01 = ra.w; cur.w23 = rb.x; cur.grow = (int)rb.y;
                cur.cr = (int)(rb.z & 0xFFFFu); cur.slot = (int)((rb.z >> 16) & 0xFFu);
            }
            TB_JIT();
            mbar_arrive_warp(BAR(BAR_REC_EMPTY + (int)(j % NREC)));
            TB_TRACE(pt_role, j, 0);
            if (!mine) {
                // Walk over the ring positions of the other group's tile, WAITING on each: an mbarrier wait tells apart
                // only adjacent phases, so nobody may get two phases ahead on an entry (a tile can span more chunks
                // than the ring has entries).
                for (int i = cur.c1 - cur.c0 + 1; i > 0; --i) {
                    mbar_wait(BAR(BAR_EMPTY_A + e), ph ^ 1);
                    if (++e == NRA) { e = 0; ph ^= 1; }
                }
                continue;
            }
            const bool ok = cur.ok;
            const float x = cur.x, y = cur.y, zp = cur.zp;
            const bool plain = ok && cur.slot != 0xFF;     // bilinear taps (else: the learned empty feature, bts.py:311-319)
            // ---- positional code -> code operand (positional_encoding.py:68-80; sin/cos of 1.5*2^k*v by angle
            //      doubling from one accurate sincosf per coordinate) -------------------------------------------
            uint32_t pk[24], pl[X3 ? 24 : 1];
#pragma unroll
            for (int i = 0; i < 24; ++i) pk[i] = 0u;
#pragma unroll
            for (int i = 0; i < (X3 ? 24 : 1); ++i) pl[i] = 0u;
            if (ok && !(SD_TB_ABLATE & 1)) {
                float code[48];
#pragma unroll
                for (int i = 39; i < 48; ++i) code[i] = 0.0f;
                code[0] = x; code[1] = y; code[2] = zp;
                if (X3) {
                    code[39] = plain ? 0.0f : 1.0f;                    // learn_empty: W_feat . empty_feature (column 39 of the x3 code block)
                } else {
                    code[45] = 1.0f; code[46] = 1.0f;                  // layer-1 bias (hi, lo) comes out of the MMA
                    code[47] = plain ? 0.0f : 1.0f;                    // learn_empty: W_feat . empty_feature (bts.py:311-319)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {                      // hi/lo split of the raw coordinates (see mlp_pack_kernel)
                        const float hi = __half2float(__float2half_rn(code[d]));
                        code[39 + d] = code[d] - hi;
                        code[42 + d] = hi;
                    }
                }
                float s[3], c[3];
                const float a0 = x * P.fp.enc.freq_factor, a1 = y * P.fp.enc.freq_factor, a2 = zp * P.fp.enc.freq_factor;
                if (!X3 && fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fabsf(a2)) <= 3.2f) {    // the usual case: hardware sin / cos
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
                for (int i = 0; i < 24; ++i) {
                    if (X3) {
                        const __half2 hi = __floats2half2_rn(code[2 * i], code[2 * i + 1]);
                        const float2 hf = __half22float2(hi);
                        pk[i] = as_u32(hi);
                        pl[i] = pack_h2(code[2 * i] - hf.x, code[2 * i + 1] - hf.y);
                    } else {
                        pk[i] = pack_h2(code[2 * i], code[2 * i + 1]);
                    }
                }
            }
            const int cs = (int)(j % NCODE);
            TB_TRACE(pt_role, j, 4);
            mbar_wait(BAR(BAR_EMPTY_C + cs), (uint32_t)(((j / NCODE) & 1) ^ 1));
            TB_TRACE(pt_role, j, 5);
            reinterpret_cast<int *>(sm + OFF_PERM)[(int)(j % NPERM) * TM + row] = ok ? cur.grow : -1;
            if (row == 0) reinterpret_cast<volatile int *>(sm + OFF_PTILE)[(int)(j % NPERM)] = tile_idx;
            unsigned char *crow = sm + OFF_CODE + cs * XK * CHUNK + row * 128;
#pragma unroll
            for (int qq = 0; qq < 6; ++qq) {
                *reinterpret_cast<uint4 *>(crow + ((qq ^ (row & 7)) << 4)) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
                if (X3) *reinterpret_cast<uint4 *>(crow + CHUNK + ((qq ^ (row & 7)) << 4)) = make_uint4(pl[4 * qq], pl[4 * qq + 1], pl[4 * qq + 2], pl[4 * qq + 3]);
            }
            // (made visible to the tensor cores by the fence + arrival of the tile's first weight chunk below)
            // ---- bilinear weights -> this row's 4 slots of its bin's chunk; every other slot of the row stays zero.
            //      Slots (nw, ne) = (s, s+1) share one 16-byte piece of the row (lx <= 6), (sw, se) the next one:
            //      the dirty byte keeps (ly << 3 | lx) and the undo clears the same two pairs.
            const int lx = cur.slot & 7, ly = (cur.slot >> 3) & 7;
            const int q = cur.cr - cur.c0;
            uint32_t w_top = cur.w01, w_bot = cur.w23, w_top_l = 0u, w_bot_l = 0u;
            if (X3 && plain) {      // the fp32 weights again (the sort made them from the same x, y: binning.cu) -> (hi, lo) halves
                Tap t = bilinear_tap(x, y, P.fp.Hf, P.fp.Wf);
                clamp_footprint(t, P.fp.Hf, P.fp.Wf);
                const __half2 th = __floats2half2_rn(t.wnw, t.wne), bh = __floats2half2_rn(t.wsw, t.wse);
                const float2 tf = __half22float2(th), bf = __half22float2(bh);
                w_top = as_u32(th); w_bot = as_u32(bh);
                w_top_l = pack_h2(t.wnw - tf.x, t.wne - tf.y); w_bot_l = pack_h2(t.wsw - bf.x, t.wse - bf.y);
            }
            auto pair_off = [&](int lyy, int lxx) {      // byte offset of slot (lyy, lxx) inside the row
                return (uint32_t)(((lyy ^ (row & 7)) << 4) + lxx * 2);
            };
            TB_TRACE(pt_role, j, 1);
            const int m = cur.c1 - cur.c0 + 1;
            TB_PROG(j, 5, m, (cur.c0 << 16) | (cur.c1 & 0xFFFF));
            for (int i = 0; i < m; ++i) {
                mbar_wait(BAR(BAR_EMPTY_A + e), ph ^ 1);
                TB_PROG(j, 20 + i, m, e);
                if (i == 0) TB_TRACE(pt_role, j, 2);
                unsigned char *arow = sm + OFF_A + e * XK * CHUNK + row * 128;
                const int d = s_dirty[e * TM + row];
                if (d != 0xFF) {
                    unsigned short *p0 = reinterpret_cast<unsigned short *>(arow + pair_off(d >> 3, d & 7));
                    unsigned short *p1 = reinterpret_cast<unsigned short *>(arow + pair_off((d >> 3) + 1, d & 7));
                    p0[0] = 0; p0[1] = 0; p1[0] = 0; p1[1] = 0;
                    if (X3) {
                        p0 += CHUNK / 2; p1 += CHUNK / 2;
                        p0[0] = 0; p0[1] = 0; p1[0] = 0; p1[1] = 0;
                    }
                }
                if (plain && i == q) {
                    unsigned short *p0 = reinterpret_cast<unsigned short *>(arow + pair_off(ly, lx));
                    unsigned short *p1 = reinterpret_cast<unsigned short *>(arow + pair_off(ly + 1, lx));
                    p0[0] = (unsigned short)w_top; p0[1] = (unsigned short)(w_top >> 16);
                    p1[0] = (unsigned short)w_bot; p1[1] = (unsigned short)(w_bot >> 16);
                    if (X3) {
                        p0 += CHUNK / 2; p1 += CHUNK / 2;
                        p0[0] = (unsigned short)w_top_l; p0[1] = (unsigned short)(w_top_l >> 16);
                        p1[0] = (unsigned short)w_bot_l; p1[1] = (unsigned short)(w_bot_l >> 16);
                    }
                    s_dirty[e * TM + row] = (unsigned char)cur.slot;
                } else if (d != 0xFF) {
                    s_dirty[e * TM + row] = 0xFF;
                }
                if (i == 0 && row == 0) s_minfo[e] = m;           // chunks of this tile, read by the MMA issuer behind FULL_A
                fence_proxy_async();
                TB_JIT();
                mbar_arrive_warp(BAR(BAR_FULL_A + e));
                if (++e == NRA) { e = 0; ph ^= 1; }
            }
            TB_TRACE(pt_role, j, 3);
            // ---- colours of the render views (bts.py:330-441, 557-569): only when asked for; the point and its
            //      frustum flag are fetched / recomputed here (the SSC query does not take this path)
            if (ok && nv_c > 0 && (P.rgb || P.invalid)) {
                const long long grow = cur.grow;
                const float px = __ldg(P.xyz + 3 * grow), py = __ldg(P.xyz + 3 * grow + 1), pz = __ldg(P.xyz + 3 * grow + 2);
                float ex, ey, ez;
                bool inv;
                project_point(s_cam, s_cam + 9, px, py, pz, ex, ey, ez, inv);
                for (int v = 0; v < nv_c; ++v) {
                    float cx, cy, cz;
                    bool cinv;
                    const float *c = s_cam + 21 * (1 + v);
                    project_point(c, c + 9, px, py, pz, cx, cy, cz, cinv);
                    if (P.rgb) {
                        float c3[3];
                        sample_color(P.fp.rgb + (size_t)v * 3 * P.fp.Hc * P.fp.Wc, P.fp.Hc, P.fp.Wc, cx, cy, c3);
                        float *o = P.rgb + (size_t)grow * 3 * nv_c + 3 * v;
                        o[0] = c3[0]; o[1] = c3[1]; o[2] = c3[2];
                    }
                    if (P.invalid) P.invalid[(size_t)grow * nv_c + v] = (cinv || inv) ? 1.0f : 0.0f;
                }
            }
            TB_TRACE(pt_role, j, 6);
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------
    if ((P.dbg & 8192) && blockIdx.x < 256 && (tid & 31) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(&g_cta_ns[2 * blockIdx.x + 1], t);
    }
    bulk_wait<0>();            // output rows still in flight (epilogue threads)
    tc_fence_before();
    __syncthreads();
    if ((P.dbg & 8192) && blockIdx.x < 4 && tid == 0) g_tiles[(blockIdx.x * 512 + 511) * 2] = clock64();
    if (warp == WARP_MMA) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace TB_NS

#if !SD_TB_X3
// debug: clock64 trace of the last launch made with SD_TC_DEBUG & 8192 (not part of the public header)
extern "C" int sd_debug_read_cta_ns(unsigned long long *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tb::g_cta_ns, sizeof(unsigned long long) * 512));
    return SD_OK;
}
extern "C" int sd_debug_read_tiles_bin(long long *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tb::g_tiles, sizeof(long long) * 4 * 2 * 512));
    return SD_OK;
}
extern "C" int sd_debug_read_trace_bin(long long *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tb::g_trace, sizeof(long long) * 8 * 64 * 8));
    return SD_OK;
}

#ifdef SD_TRAP_REPORT
extern "C" int sd_debug_set_trap_buffer(unsigned int *host_mapped) {
    SD_CUDA_OK(cudaMemcpyToSymbol(tcx::g_trap_report, &host_mapped, sizeof(host_mapped)));
    return SD_OK;
}
#endif
#if defined(SD_DEBUG_WAIT) || defined(SD_DEBUG_LONGWAIT)
#ifdef SD_DEBUG_WAIT
extern "C" int sd_debug_read_progress(int *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tb::g_prog, sizeof(int) * 160 * 16 * 4));
    return SD_OK;
}
#endif
extern "C" int sd_debug_read_timeout(unsigned int *host_out) {
    SD_CUDA_OK(cudaMemcpyFromSymbol(host_out, tcx::g_wait_timeout, sizeof(unsigned int) * 260));
    return SD_OK;
}
#endif

#endif  // !SD_TB_X3

#if SD_TB_X3
bool bin_kernel_supported_x3(const sd_scene *s, const sd_mlp *mlp) {
    return s && mlp && s->feat_proj_x3 && mlp->packed && mlp->precision == SD_MLP_F32_TC && s->C == 256 && s->nv_f == 1 &&
           s->include_input && s->num_freqs == 6 && s->Hf >= 2 && s->Wf >= 2 && s->nv_c <= TB_NS::MAX_NVC_TB &&
           mlp->d_hidden == 128 && mlp->d_in == s->C + 39 && mlp->d_out >= 2 && mlp->d_out - 1 <= 64;
}
#define TB_SUPPORTED bin_kernel_supported_x3
int launch_field_bin_x3(const sd_scene *scene, const FieldParams &fp, const float *xyz, long long N, const sd_mlp *mlp,
                        const BinOrder &order, const TcOut &out, cudaStream_t st) {
#else
bool bin_kernel_supported(const sd_scene *s, const sd_mlp *mlp) {
    return s && mlp && s->feat_proj && mlp->packed && mlp->precision == SD_MLP_F16_TC && s->C == 256 && s->nv_f == 1 &&
           s->include_input && s->num_freqs == 6 && s->Hf >= 2 && s->Wf >= 2 && s->nv_c <= tb::MAX_NVC_TB &&
           mlp->d_hidden == 128 && mlp->d_in == s->C + 39 && mlp->d_out >= 2 && mlp->d_out - 1 <= 64;
}
#define TB_SUPPORTED bin_kernel_supported
int launch_field_bin(const sd_scene *scene, const FieldParams &fp, const float *xyz, long long N, const sd_mlp *mlp,
                     const BinOrder &order, const TcOut &out, cudaStream_t st) {
#endif
    if (N == 0) return SD_OK;
    SD_REQUIRE(TB_SUPPORTED(scene, mlp), "field_bin: unsupported scene / head for the projected-map kernel");
    SD_REQUIRE(order.bw == SD_BIN, "field_bin: the feature map is too large for %d x %d bins", SD_BIN, SD_BIN);
    SD_REQUIRE(N < (1ll << 31), "field_bin: at most 2^31 - 1 points per call (got %lld)", N);
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    const unsigned char *proj = reinterpret_cast<const unsigned char *>(SD_TB_X3 ? scene->feat_proj_x3 : scene->feat_proj);
    SD_REQUIRE(((uintptr_t)blob & 15) == 0 && ((uintptr_t)proj & 15) == 0, "field_bin: packed blobs must be 16-byte aligned");
    SD_REQUIRE(((uintptr_t)out.dino & 15) == 0, "field_bin: dino must be 16-byte aligned");
    TB_NS::Params P = {};
    P.fp = fp;
    P.xyz = xyz;
    P.cbin = order.cbin; P.nbx = order.nbx;
    SD_REQUIRE(order.has_geo, "field_bin: the point order carries no geometry records");
    P.rec = order.rec;
    P.tiles = order.tiles;
    P.tile_ctr = order.tile_ctr;
    P.N = N;
    P.dbg = debug_mask();
    P.n_tiles = (N + TB_NS::TM - 1) / TB_NS::TM;
    P.D = mlp->d_out - 1;
    P.n2 = (mlp->d_out + 15) / 16 * 16;
#if SD_TB_X3
    SD_REQUIRE(P.n2 == 80 || P.n2 * 4 * 128 <= PROJX_W2_BYTES, "field_bin_x3: d_out too large");
    P.wc_img = proj + PROJX_OFF_WC;
    P.w2_img = proj + PROJX_OFF_W2;
    P.b_in = reinterpret_cast<const float *>(blob + L.off_b_in);
#else
    P.wc_img = proj + PROJ_OFF_CODE;
    P.w2_img = blob + L.off_w_out_h;
#endif
    P.b_out = reinterpret_cast<const float *>(blob + L.off_b_out);
    P.sigma = out.sigma; P.dino = out.dino; P.rgb = out.rgb; P.invalid = out.invalid; P.invalid_feat = out.invalid_feat;
    if (out.dino_binned) {
        SD_REQUIRE(P.D == 64 && !out.dino, "field_bin: binned output needs a 64-d head and no caller-order dino");
        SD_REQUIRE(((uintptr_t)out.dino_binned & 15) == 0, "field_bin: dino_binned must be 16-byte aligned");
        P.binned = 1; P.dino = out.dino_binned; P.perm_out = out.perm_out;
#if !SD_TB_X3
        const unsigned long long odims[2] = {64ull, (unsigned long long)N}, ostrides[1] = {256ull};
        const unsigned int obox[2] = {32u, 32u};
        if (int rc_o = make_tmap(&P.tmap_out, out.dino_binned, 4, 2, odims, ostrides, obox)) return rc_o;
#endif
    }
    const unsigned long long dims[3] = {128ull, (unsigned long long)fp.Wf, (unsigned long long)fp.Hf};
    const unsigned long long strides[2] = {256ull, 256ull * (unsigned long long)fp.Wf};
    const unsigned int box[3] = {64u, 8u, 8u};
#if SD_TB_X3
    const size_t map_bytes = (size_t)fp.Hf * fp.Wf * 256;
    int rc = make_tmap_f16(&P.tmap, proj + PROJX_OFF_MAP, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_f16(&P.tmap_lo, proj + PROJX_OFF_MAP + map_bytes, 3, dims, strides, box);
    if (rc) return rc;
#else
    int rc = make_tmap_f16(&P.tmap, proj + PROJ_OFF_MAP, 3, dims, strides, box);
    if (rc) return rc;
#endif
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(TB_NS::field_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_NS::SMEM_ALLOC));
    }
    const unsigned grid = (unsigned)(P.n_tiles < sm_count ? P.n_tiles : sm_count);
    profile_before(st);
    TB_NS::field_bin_kernel<<<grid, TB_NS::NTHREADS, TB_NS::SMEM_ALLOC, st>>>(P);
    profile_after(st);
    SD_LAUNCH_OK(SD_TB_X3 ? "field_bin_kernel (x3)" : "field_bin_kernel");
    return SD_OK;
}

}  // namespace sd
