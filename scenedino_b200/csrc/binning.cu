// Texel binning of query points: a counting sort of the point indices by the 7x7-texel block of the
// feature map their (clamped) bilinear footprint starts in -- so the footprints of a bin cover an 8x8-texel box,
// exactly one 64-row operand chunk of the tile kernel (field_bin.cu).
//
// The field query (BTSNet.forward, models/bts.py:476-595) treats every point independently, so the order in
// which the fused kernel walks the points is free.  In the caller's order (e.g. the SSC voxel grid,
// sscbench/evaluate_model_sscbench.py:270-279, z fastest) neighbouring points project to texels that are 2-11
// texels apart and share nothing: every point pulls its own 4 x 512 B through L2.  Walked bin by bin, the rows
// of a 128-point tile share a footprint of at most 9 x 9 texels, the gather hits L1, and L2 / HBM traffic
// falls towards the unique bytes (measured on the SSC grid: DRAM reads 710 -> 243 MB, L2 reads 3.7 -> 1.1 GB).
// (Sorting by the exact texel instead was measured too: no faster in the fused kernel, and the histogram
// atomics on the few corner texels that collect all behind-camera points made the sort itself 3x slower.)
//
// Two passes over the SAME partition of the points into contiguous ranges, one per block: (1) histogram per range in
// shared memory; its merge into the global histogram (one atomic per non-empty bin) hands back the number of points earlier
// ranges put into the bin -- the range's offset inside the bin, kept in `bbase`; (2) scatter: every block scans the histogram
// for itself (bin starts + a compact numbering of the non-empty bins, in shared memory), then position = start of the bin +
// offset of the range + rank inside the range (shared-memory atomic); (3) the tile table.  The points are read twice and
// projected twice (carrying the projection through memory would cost about what the second projection does).  The order inside a bin depends on atomics and is not reproducible;
// the results per point are (the fused kernel computes each row independently).
#include "common.cuh"
#include "launch.h"
#include "tc_common.cuh"

namespace sd {

constexpr int SCAT_THREADS = 512;
constexpr int MAX_BINS = 12288;     // the scatter pass holds bin starts and compact numbers in 96 KB of shared memory per block
constexpr int BLOCKS_PER_SM = 2;    // (three per SM at 40 registers were measured: no change)
constexpr int MAX_RANGES = 320;     // ranges (= blocks) of the count and scatter passes: BLOCKS_PER_SM per SM, at most this many

struct BinGeom {
    int Hf, Wf, bw, nbx, nbins;   // bins of bw x bw texels (bw = SD_BIN unless the map is huge)
    unsigned int bw_magic;        // ceil(2^32 / bw): v / bw == __umulhi(v, bw_magic) for v * bw < 2^32
};

constexpr int BIN_UNROLL = 4;       // points per thread and trip, their 12 coordinate loads issued together (count pass -2 us;
                                    // beyond that both passes are bound by the instructions of the projection itself:
                                    // profiles/r02_sort_passes.md)

__device__ __forceinline__ int point_bin(const float *cam, const BinGeom &bg, float px, float py, float pz) {
    float x, y, z;
    bool inv;
    project_point(cam, cam + 9, px, py, pz, x, y, z, inv);
    Tap t = bilinear_tap(clamp_keep_nan(x, -2.0f, 2.0f), clamp_keep_nan(y, -2.0f, 2.0f), bg.Hf, bg.Wf);
    clamp_footprint(t, bg.Hf, bg.Wf);   // the same base texel as the kernels that consume the order
    return (int)(__umulhi((unsigned int)t.y0, bg.bw_magic) * (unsigned int)bg.nbx + __umulhi((unsigned int)t.x0, bg.bw_magic));
}

// Histogram with shared-memory pre-aggregation per range (a warp-aggregated version with global atomics only was
// measured: 1.7x slower, the border bins that collect the out-of-frustum points serialise in L2).  bbase[range][bin] =
// what the global counter of the bin held when this range added its share.
template <bool WRITE_BINS>
__global__ void __launch_bounds__(SCAT_THREADS) bin_count_kernel(const float *__restrict__ K, const float *__restrict__ w2c,
                                                                 BinGeom bg, const float *__restrict__ xyz, long long N,
                                                                 unsigned short *__restrict__ bins,
                                                                 unsigned int *__restrict__ hist,
                                                                 unsigned int *__restrict__ bbase,
                                                                 unsigned short *__restrict__ blist,
                                                                 unsigned int *__restrict__ bcount) {
    extern __shared__ unsigned int sh[];
    __shared__ float cam[21];
    __shared__ unsigned int n_list;
    if (threadIdx.x == 0) n_list = 0;
    for (int i = threadIdx.x; i < 21; i += SCAT_THREADS) cam[i] = i < 9 ? __ldg(K + i) : __ldg(w2c + (i - 9));
    for (int b = threadIdx.x; b < bg.nbins; b += SCAT_THREADS) sh[b] = 0;
    __syncthreads();
    const long long per = (N + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x, hi = min(N, lo + per);
    const unsigned int n = hi > lo ? (unsigned int)(hi - lo) : 0u;
    const float *__restrict__ p = xyz + 3 * lo;
    for (unsigned int j0 = threadIdx.x; j0 < n; j0 += BIN_UNROLL * SCAT_THREADS) {   // (32-bit indexing inside a range: N < 2^31)
        float q[BIN_UNROLL][3];
#pragma unroll
        for (int u = 0; u < BIN_UNROLL; ++u) {
            const unsigned int j = j0 + u * SCAT_THREADS;
#pragma unroll
            for (int c = 0; c < 3; ++c) q[u][c] = j < n ? __ldg(p + 3u * j + c) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < BIN_UNROLL; ++u) {
            const unsigned int j = j0 + u * SCAT_THREADS;
            if (j < n) {
                const int b = point_bin(cam, bg, q[u][0], q[u][1], q[u][2]);
                if (WRITE_BINS) bins[lo + j] = (unsigned short)b;
                atomicAdd(&sh[b], 1u);
            }
        }
    }
    __syncthreads();
    // a range of consecutive points touches a small part of the bins: the scatter pass gets the list of them, so that it
    // reads (and this pass writes) only those entries of the range's row
    unsigned int *__restrict__ row = bbase + (size_t)blockIdx.x * bg.nbins;
    unsigned short *__restrict__ list = blist + (size_t)blockIdx.x * bg.nbins;
    for (int b = threadIdx.x; b < bg.nbins; b += SCAT_THREADS) {
        const unsigned int c = sh[b];
        if (c) {
            row[b] = atomicAdd(&hist[b], c);
            list[atomicAdd(&n_list, 1u)] = (unsigned short)b;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) bcount[blockIdx.x] = n_list;
}

// Exclusive scan of the histogram by ONE block (SCAT_THREADS threads), into shared memory: s_start[b] = first sorted position
// of bin b, s_ci[b] = number of non-empty bins before b (the compact number of b when it is non-empty, else 0xFFFFFFFF).
// Every block of the scatter pass does this for itself (40 KB of counts from L2, ~2 us, all blocks at once) instead of waiting
// for a one-block kernel in between (8 us + a launch); block 0 also publishes cbin[c] = bin with compact number c and
// meta[0] = their count.
__device__ __forceinline__ void block_scan_bins(const unsigned int *__restrict__ hist, int nbins, unsigned int *s_start,
                                                unsigned int *s_ci, unsigned int *__restrict__ cbin,
                                                unsigned int *__restrict__ meta) {
    __shared__ unsigned int warp_tot[SCAT_THREADS / 32], warp_ne[SCAT_THREADS / 32];
    for (int i = threadIdx.x; i < nbins; i += SCAT_THREADS) s_start[i] = __ldcg(hist + i);
    __syncthreads();
    const int per = (nbins + SCAT_THREADS - 1) / SCAT_THREADS;
    const int lo = min(nbins, (int)threadIdx.x * per), hi = min(nbins, lo + per);
    unsigned int sum = 0, ne = 0;
    for (int i = lo; i < hi; ++i) { const unsigned int c = s_start[i]; sum += c; ne += c != 0; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int incl = sum, incl_ne = ne;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o), m = __shfl_up_sync(0xffffffffu, incl_ne, o);
        if (lane >= o) { incl += n; incl_ne += m; }
    }
    if (lane == 31) { warp_tot[warp] = incl; warp_ne[warp] = incl_ne; }
    __syncthreads();
    unsigned int run = incl - sum, run_ne = incl_ne - ne;
    for (int w = 0; w < warp; ++w) { run += warp_tot[w]; run_ne += warp_ne[w]; }
    for (int i = lo; i < hi; ++i) {
        const unsigned int c = s_start[i];
        s_start[i] = run;
        s_ci[i] = c ? run_ne : 0xFFFFFFFFu;
        if (c && blockIdx.x == 0) cbin[run_ne] = (unsigned int)i;
        run += c;
        run_ne += c != 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == SCAT_THREADS - 1) meta[0] = run_ne;
    __syncthreads();
}

// Per-point record of the projected-map tile kernel (field_bin.cu), written at the SORTED position of the point as
// ONE 32-byte store (a whole sector: seven separate scattered 4-byte stores were measured at +80 us per 2 M points),
// so that the tile kernel reads it with coalesced loads and never projects anything itself.
struct GeoOut {
    GeoRec *rec;                  // [N]
    unsigned char *invalid_feat;  // caller order [N] or NULL
    EncodeParams enc;
    int learn_empty;
};

template <bool GEO>
__global__ void __launch_bounds__(SCAT_THREADS) bin_scatter_kernel(BinGeom bg, long long N,
                                                                  const unsigned short *__restrict__ bins,
                                                                  const unsigned int *__restrict__ hist,
                                                                  const unsigned int *__restrict__ bbase,
                                                                  const unsigned short *__restrict__ blist,
                                                                  const unsigned int *__restrict__ bcount,
                                                                  unsigned int *__restrict__ cbin,
                                                                  unsigned int *__restrict__ meta,
                                                                  unsigned int *__restrict__ perm,
                                                                  unsigned short *__restrict__ pcb,
                                                                  const float *__restrict__ K, const float *__restrict__ w2c,
                                                                  const float *__restrict__ xyz, GeoOut go) {
    extern __shared__ unsigned int sh[];          // [nbins] bin starts, then the next free position of this range in each bin
    __shared__ float cam[21];                     // it touches; [nbins] compact bin numbers
    unsigned int *s_ci = sh + bg.nbins;
    block_scan_bins(hist, bg.nbins, sh, s_ci, cbin, meta);
    const unsigned int *__restrict__ row = bbase + (size_t)blockIdx.x * bg.nbins;
    const unsigned short *__restrict__ list = blist + (size_t)blockIdx.x * bg.nbins;
    const unsigned int n_list = __ldg(bcount + blockIdx.x);
    for (unsigned int k = threadIdx.x; k < n_list; k += SCAT_THREADS) {
        const unsigned int b = __ldg(list + k);
        sh[b] += __ldg(row + b);
    }
    if (GEO)
        for (int i = threadIdx.x; i < 21; i += SCAT_THREADS) cam[i] = i < 9 ? __ldg(K + i) : __ldg(w2c + (i - 9));
    __syncthreads();
    const long long per = (N + gridDim.x - 1) / gridDim.x;   // the partition of bin_count_kernel (same grid)
    const long long lo = per * blockIdx.x, hi = min(N, lo + per);
    const unsigned int n = hi > lo ? (unsigned int)(hi - lo) : 0u;
    if (!GEO) {
        for (unsigned int j = threadIdx.x; j < n; j += SCAT_THREADS) {
            const int b = bins[lo + j];
            const unsigned int pos = atomicAdd(&sh[b], 1u);
            perm[pos] = (unsigned int)lo + j;
            pcb[pos] = (unsigned short)s_ci[b];
        }
    } else {
        const float *__restrict__ p = xyz + 3 * lo;
        const unsigned int i0 = (unsigned int)lo;
        unsigned char *__restrict__ inv_out = go.invalid_feat ? go.invalid_feat + lo : nullptr;
        // (One atomic per run of equal bins in a warp instead of one per lane was measured: no change -- the pass is bound by
        // the ~250 instructions per point of projection, normalisation and record packing, not by the atomics.)
        for (unsigned int j0 = threadIdx.x; j0 < n; j0 += BIN_UNROLL * SCAT_THREADS) {
            float q[BIN_UNROLL][3];
#pragma unroll
            for (int u = 0; u < BIN_UNROLL; ++u) {
                const unsigned int j = j0 + u * SCAT_THREADS;
#pragma unroll
                for (int c = 0; c < 3; ++c) q[u][c] = j < n ? __ldg(p + 3u * j + c) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < BIN_UNROLL; ++u) {
                const unsigned int j = j0 + u * SCAT_THREADS;
                if (j >= n) continue;
                float x, y, zc;
                bool inv;
                project_point(cam, cam + 9, q[u][0], q[u][1], q[u][2], x, y, zc, inv);
                x = clamp_keep_nan(x, -2.0f, 2.0f);
                y = clamp_keep_nan(y, -2.0f, 2.0f);
                Tap t = bilinear_tap(x, y, bg.Hf, bg.Wf);
                clamp_footprint(t, bg.Hf, bg.Wf);
                const int bx = t.x0 / SD_BIN, by = t.y0 / SD_BIN;            // GEO: bg.bw == SD_BIN
                const int b = by * bg.nbx + bx;                              // the bin bin_count_kernel counted this point in
                const int lx = t.x0 - bx * SD_BIN, ly = t.y0 - by * SD_BIN;
                const unsigned int pos = atomicAdd(&sh[b], 1u);
                const unsigned int ci = s_ci[b];
                const unsigned int slot = (go.learn_empty && inv) ? 0xFFu : (unsigned int)(ly * 8 + lx);
                uint4 *dst = reinterpret_cast<uint4 *>(go.rec + pos);
                dst[0] = make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(znorm(zc, go.enc)), tcx::pack_h2(t.wnw, t.wne));
                dst[1] = make_uint4(tcx::pack_h2(t.wsw, t.wse), i0 + j, ci | (slot << 16), (unsigned int)b);
                if (inv_out) inv_out[j] = inv ? 1 : 0;
            }
        }
    }
}

// Table of the tile kernel's tiles (128 consecutive sorted points each), in the order they are handed out (last tile
// first): what its producer needs, so that it never waits for a load of its own.
__global__ void __launch_bounds__(256) bin_tiles_kernel(const GeoRec *__restrict__ rec, const unsigned int *__restrict__ cbin,
                                                        long long N, long long n_tiles, TileInfo *__restrict__ tiles) {
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    if (v >= n_tiles + 8) return;
    TileInfo ti = {0u, 0u, 0u, 0u};
    if (v < n_tiles) {
        const long long t = n_tiles - 1 - v, a = t * 128;
        const unsigned int rows = (unsigned int)(N - a < 128 ? N - a : 128);
        const unsigned int c0 = rec[a].cs & 0xFFFFu, c1 = rec[a + rows - 1].cs & 0xFFFFu, m = c1 - c0 + 1;
        unsigned int b[4];
        for (int i = 0; i < 4; ++i) b[i] = (unsigned int)i < m ? cbin[c0 + i] : 0u;
        ti.c0m = c0 | (m << 16); ti.b01 = b[0] | (b[1] << 16); ti.b23 = b[2] | (b[3] << 16); ti.rows = rows;
    }
    tiles[v] = ti;
}

static BinGeom bin_geom(int Hf, int Wf) {
    BinGeom g;
    g.Hf = Hf; g.Wf = Wf; g.bw = SD_BIN;
    for (;;) {
        g.nbx = (Wf - 1) / g.bw + 1;
        g.nbins = g.nbx * ((Hf - 1) / g.bw + 1);
        g.bw_magic = (unsigned int)(((1ull << 32) + (unsigned)g.bw - 1) / (unsigned)g.bw);
        if (g.nbins <= MAX_BINS) return g;
        g.bw += SD_BIN;
    }
}

static size_t a256(size_t v) { return (v + 255) / 256 * 256; }

size_t bin_workspace_bytes(int Hf, int Wf, long long N) {
    if (N <= 0 || N >= (1ll << 31)) return 0;
    const BinGeom g = bin_geom(Hf, Wf);
    return a256((size_t)N * 4) + 2 * a256((size_t)N * 2) + 3 * a256((size_t)g.nbins * 4) + 256 +
           a256((size_t)N * sizeof(GeoRec)) +               // per-point records of the tile kernel
           a256((size_t)((N + 127) / 128 + 8) * sizeof(TileInfo)) +
           a256((size_t)MAX_RANGES * g.nbins * 4) +         // offsets of the ranges inside the bins,
           a256((size_t)MAX_RANGES * g.nbins * 2) + a256((size_t)MAX_RANGES * 4);   // the bins each range touches
}

// Sorts the point indices by bin.  Fills `out` with device pointers into the workspace.
int launch_bin_points(const FieldParams &fp, const float *xyz, long long N, void *workspace, size_t workspace_bytes,
                      BinOrder *out, cudaStream_t st, bool want_geo, unsigned char *invalid_feat, bool reuse_sorted) {
    const size_t need = bin_workspace_bytes(fp.Hf, fp.Wf, N);
    if (need == 0 || workspace_bytes < need || !workspace) {
        set_error("binning: workspace of %zu B needed, %zu B given", need, workspace_bytes);
        return SD_ERR_WORKSPACE;
    }
    SD_REQUIRE(fp.Hf >= 2 && fp.Wf >= 2, "binning: the feature map must be at least 2 x 2");
    const BinGeom g = bin_geom(fp.Hf, fp.Wf);
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    unsigned int *perm = reinterpret_cast<unsigned int *>(ws);               ws += a256((size_t)N * 4);
    unsigned short *bins = reinterpret_cast<unsigned short *>(ws);           ws += a256((size_t)N * 2);
    unsigned short *pcb = reinterpret_cast<unsigned short *>(ws);            ws += a256((size_t)N * 2);
    unsigned int *hist = reinterpret_cast<unsigned int *>(ws);               ws += a256((size_t)g.nbins * 4);
    unsigned int *meta = reinterpret_cast<unsigned int *>(ws);               ws += 256;   // right behind hist: one memset
    unsigned int *cidx = reinterpret_cast<unsigned int *>(ws);               ws += a256((size_t)g.nbins * 4);
    unsigned int *cbin = reinterpret_cast<unsigned int *>(ws);               ws += a256((size_t)g.nbins * 4);
    GeoOut go = {};
    go.rec = reinterpret_cast<GeoRec *>(ws);                                                           ws += a256((size_t)N * sizeof(GeoRec));
    TileInfo *tiles = reinterpret_cast<TileInfo *>(ws);                      ws += a256((size_t)((N + 127) / 128 + 8) * sizeof(TileInfo));
    unsigned int *bbase = reinterpret_cast<unsigned int *>(ws);              ws += a256((size_t)MAX_RANGES * g.nbins * 4);
    unsigned short *blist = reinterpret_cast<unsigned short *>(ws);          ws += a256((size_t)MAX_RANGES * g.nbins * 2);
    unsigned int *bcount = reinterpret_cast<unsigned int *>(ws);
    go.invalid_feat = invalid_feat;
    go.enc = fp.enc;
    go.learn_empty = fp.learn_empty;
    static DeviceOnce once;
    int sm_count = 0;
    bool first_use = false;
    if (int rc_dev = device_once(once, &sm_count, &first_use)) return rc_dev;
    if (first_use) {
        SD_CUDA_OK(cudaFuncSetAttribute(bin_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BINS * 4));
        SD_CUDA_OK(cudaFuncSetAttribute(bin_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BINS * 4));
        SD_CUDA_OK(cudaFuncSetAttribute(bin_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BINS * 8));
        SD_CUDA_OK(cudaFuncSetAttribute(bin_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_BINS * 8));
    }
    if (reuse_sorted) {
        // the workspace still holds the sort of these points for these cameras (sd_query_points_sorted): only the tile
        // counter has to start from zero again
        SD_REQUIRE(want_geo && g.bw == SD_BIN, "binning: nothing to reuse for this map size");
        SD_CUDA_OK(cudaMemsetAsync(meta, 0, 256, st));
        out->perm = perm; out->pcb = pcb; out->cbin = cbin; out->bw = g.bw; out->nbx = g.nbx; out->nbins = g.nbins;
        out->has_geo = true; out->rec = go.rec; out->tile_ctr = meta + 1; out->tiles = tiles;
        return SD_OK;
    }
    const long long blocks_wanted = (N + 2 * SCAT_THREADS - 1) / (2 * SCAT_THREADS);
    const long long cap = BLOCKS_PER_SM * sm_count < MAX_RANGES ? BLOCKS_PER_SM * sm_count : MAX_RANGES;
    const unsigned grid = (unsigned)(blocks_wanted < cap ? blocks_wanted : cap);     // ranges: the same for both passes
    const bool geo = want_geo && g.bw == SD_BIN;
    // histogram and meta (meta[1]: tile counter of the tile kernel) in one go
    SD_CUDA_OK(cudaMemsetAsync(hist, 0, a256((size_t)g.nbins * 4) + 256, st));
    if (geo)   // the scatter recomputes the bin from the projection it needs anyway: no bin array
        bin_count_kernel<false><<<grid, SCAT_THREADS, (size_t)g.nbins * 4, st>>>(fp.K_f, fp.w2c_f, g, xyz, N, bins, hist, bbase, blist, bcount);
    else
        bin_count_kernel<true><<<grid, SCAT_THREADS, (size_t)g.nbins * 4, st>>>(fp.K_f, fp.w2c_f, g, xyz, N, bins, hist, bbase, blist, bcount);
    SD_LAUNCH_OK("bin_count_kernel");
    if (geo)
        bin_scatter_kernel<true><<<grid, SCAT_THREADS, (size_t)g.nbins * 8, st>>>(g, N, bins, hist, bbase, blist, bcount, cbin, meta, perm, pcb, fp.K_f, fp.w2c_f, xyz, go);
    else
        bin_scatter_kernel<false><<<grid, SCAT_THREADS, (size_t)g.nbins * 8, st>>>(g, N, bins, hist, bbase, blist, bcount, cbin, meta, perm, pcb, fp.K_f, fp.w2c_f, xyz, go);
    SD_LAUNCH_OK("bin_scatter_kernel");
    out->perm = perm; out->pcb = pcb; out->cbin = cbin; out->bw = g.bw; out->nbx = g.nbx; out->nbins = g.nbins;
    out->has_geo = want_geo && g.bw == SD_BIN;
    out->rec = go.rec;
    out->tile_ctr = meta + 1;
    out->tiles = tiles;
    if (out->has_geo) {
        const long long n_tiles = (N + 127) / 128;
        bin_tiles_kernel<<<(unsigned)((n_tiles + 8 + 255) / 256), 256, 0, st>>>(go.rec, cbin, N, n_tiles, tiles);
        SD_LAUNCH_OK("bin_tiles_kernel");
    }
    return SD_OK;
}

}  // namespace sd
