// Internal launch interface shared by abi.cu, field_simt.cu and field_tc.cu.
#pragma once
#include "common.cuh"

namespace sd {

// Where the query points come from: an explicit [N,3] array, or implicitly o + z*d along rays
// (nerf.py:252-253) so that the [R*K,3] point cloud never exists in HBM.
struct PointSrc {
    const float *xyz;   // [N,3] or NULL
    const float *rays;  // [R,r_dim]  (ray mode: point i = ray i / K at depth z[i])
    const float *z;     // [R,K]
    int r_dim, K;
};

#ifdef __CUDACC__
__device__ __forceinline__ void load_point(const PointSrc &s, long long i, float &px, float &py, float &pz) {
    if (s.xyz) {
        px = __ldg(s.xyz + 3 * i); py = __ldg(s.xyz + 3 * i + 1); pz = __ldg(s.xyz + 3 * i + 2);
    } else {
        const long long r = i / s.K;
        const float *ry = s.rays + r * s.r_dim;
        const float zz = __ldg(s.z + i);
        px = ray_point(__ldg(ry + 0), __ldg(ry + 3), zz);
        py = ray_point(__ldg(ry + 1), __ldg(ry + 4), zz);
        pz = ray_point(__ldg(ry + 2), __ldg(ry + 5), zz);
    }
}
#endif

struct SimtOut {
    float *feat;                 // [N, d_in]           (features mode)
    unsigned char *invalid_feat; // [N]
    float *sigma;                // [N]
    float *dino;                 // [N, d_out-1]
    float *rgb;                  // [N, 3*nv_c]
    float *invalid;              // [N, nv_c]
    float *raw;                  // unused
};

enum { MODE_FEATURES_ = 0, MODE_QUERY_ = 1 };

int launch_field_simt(int mode, const FieldParams &fp, const PointSrc &src, long long N, const sd_mlp *mlp,
                      const SimtOut &out, cudaStream_t st);
int launch_mlp_simt(const sd_mlp *mlp, const float *x, long long N, float *out, bool normalize, cudaStream_t st);

// ---- fused tcgen05 path -----------------------------------------------------------------------
struct TcOut {            // per-point / per-sample outputs (any may be NULL)
    float *sigma;         // [N]
    float *dino;          // [N, D]        (point mode only)
    float *rgb;           // [N, 3*nv_c]   (point mode: BTSNet.forward rgb; render mode: rgb_samps)
    float *invalid;       // [N, nv_c]
    unsigned char *invalid_feat;  // [N]
    // projected-map tile kernel only (field_bin.cu): the 64-d rows in texel-bin (sorted) order, written by TMA tile stores,
    // and the sorted position -> point index map that goes with them
    float *dino_binned;           // [N, 64]
    unsigned int *perm_out;       // [N]
};

struct TcRender {         // per-ray outputs of the fused composite
    sd_render_cfg cfg;
    float *depth, *dino, *rgb_out, *weights, *alphas, *rgb_samps;
    int cmma;             // tc_render_mode(): 0 composite in the epilogue, 1 / 2 on the tensor cores (features / hidden units)
    float *hsum, *wsum;   // cmma == 2: [R,128] per-ray sums of w * relu(hidden) and [R] sums of w, input of launch_head2
};

// true when the fused kernel handles this (scene, head) in render mode with K samples per ray
bool tc_supported(const sd_scene *scene, const sd_mlp *mlp, int K);
int tc_render_mode(const sd_scene *scene, const sd_mlp *mlp);
// `proj`: blob of sd_field_project or NULL.  With it the kernel gathers the 128 projected channels (half the bytes per
// tap) and layer 1 is identity + code block (K 320 -> 192); without it the 256 feature channels and the full W_in.
int launch_field_tc(const FieldParams &fp, const PointSrc &src, long long N, const sd_mlp *mlp,
                    const TcRender *render, const TcOut &out, cudaStream_t st, const unsigned int *perm = nullptr,
                    const void *proj = nullptr);

// ---- texel binning of query points (binning.cu) ------------------------------------------------------
// What the tile kernel needs of a point, at its sorted position: encoder-view coordinates (x, y clamped to +-2, z' of
// positional_encoding.py:13-21), the four bilinear weights as packed halves (nw, ne) / (sw, se), the point's index in
// the caller's order, the compact number of its bin (low 16 bits of cs) and the position of its footprint inside the
// bin's 8x8 box (bits 16-23 of cs: ly * 8 + lx, or 0xFF when the row takes the learned empty feature instead of taps),
// and the id of its bin (row-major over the bin grid).
struct GeoRec {
    float x, y, zp;
    unsigned int w01, w23, perm, cs, bin;
};
static_assert(sizeof(GeoRec) == 32, "one sector per record");

// What the tile kernel's producer needs to know about a 128-point tile, in the order the tiles are handed out (entry v
// describes tile n_tiles - 1 - v): rows (0 = past the end), first compact bin and number of bins it touches (c0 | m << 16)
// and the ids of the first four of them.
struct TileInfo {
    unsigned int c0m, b01, b23, rows;
};
static_assert(sizeof(TileInfo) == 16, "bulk-copied in batches");

struct BinOrder {
    const unsigned int *perm;     // [N] sorted position -> point index
    const unsigned short *pcb;    // [N] compact number of the bin of the point at each sorted position (non-decreasing)
    const unsigned int *cbin;     // [#non-empty bins] compact number -> bin id (row-major over the bin grid)
    int bw, nbx, nbins;           // bin width in texels, bins per row, total bins
    // want_geo: one 32-byte record per sorted position instead of perm / pcb (which then stay unwritten)
    bool has_geo;
    const GeoRec *rec;
    unsigned int *tile_ctr;       // zeroed counter from which the tile kernel's CTAs claim tiles
    const TileInfo *tiles;        // [n_tiles + 8] table in claim order, empty entries behind the end
};
size_t bin_workspace_bytes(int Hf, int Wf, long long N);
int launch_bin_points(const FieldParams &fp, const float *xyz, long long N, void *workspace, size_t workspace_bytes,
                      BinOrder *out, cudaStream_t st, bool want_geo = false, unsigned char *invalid_feat = nullptr,
                      bool reuse_sorted = false);
// ---- projected scene (field_proj.cu): blob layout -----------------------------------------------------------
constexpr int PROJ_OFF_IDENT = 0;          // 2 x 16 KB: UMMA image of the 128 x 128 identity
constexpr int PROJ_OFF_CODE = 32768;       // 16 KB: UMMA image of the code block of W_in (+ projected empty feature in column 47)
constexpr int PROJ_OFF_EMPTY = 49152;      // 128 halves: W_feat . empty_feature
constexpr int PROJ_OFF_MAP = 50176;        // P [Hf*Wf][128] fp16

// ---- projected-map tile kernel (field_proj.cu, field_bin.cu) ---------------------------------------------
// encodes a tiled fp16 tensor map (SWIZZLE_128B) into the 128 bytes at tmap_out (64-byte aligned)
int make_tmap(void *tmap_out, const void *base, int elem_bytes, int rank, const unsigned long long *dims,
              const unsigned long long *strides_bytes, const unsigned int *box);   // elem_bytes: 2 = fp16, 4 = fp32
int make_tmap_f16(void *tmap_out, const void *base, int rank, const unsigned long long *dims,
                  const unsigned long long *strides_bytes, const unsigned int *box);
bool bin_kernel_supported(const sd_scene *scene, const sd_mlp *mlp);
// rel-1e-4 variant (field_bin_x3.cu; scene->feat_proj_x3 from sd_field_project_x3, mlp->precision == SD_MLP_F32_TC)
constexpr int PROJX_OFF_WC = 0;            // 2 x 16 KB: code block of W_in (+ projected empty feature in column 39), hi and lo images
constexpr int PROJX_OFF_W2 = 32768;        // 2 x (2 K blocks x n2 rows x 128 B): W_out hi and lo images (feature rows first, density row last)
constexpr int PROJX_W2_BYTES = 2 * 2 * 80 * 128;
constexpr int PROJX_OFF_MAP = PROJX_OFF_W2 + PROJX_W2_BYTES;   // P_hi [Hf*Wf][128] fp16, then P_lo
bool bin_kernel_supported_x3(const sd_scene *scene, const sd_mlp *mlp);
int launch_field_bin_x3(const sd_scene *scene, const FieldParams &fp, const float *xyz, long long N, const sd_mlp *mlp,
                        const BinOrder &order, const TcOut &out, cudaStream_t st);
int launch_field_bin(const sd_scene *scene, const FieldParams &fp, const float *xyz, long long N, const sd_mlp *mlp,
                     const BinOrder &order, const TcOut &out, cudaStream_t st);
// ResnetFC.forward on explicit rows through the same tcgen05 pipeline (unit test of the MMA path)
int launch_mlp_tc(const sd_mlp *mlp, const float *x, long long N, float *out, cudaStream_t st);
// expand_tc.cu: MlpDimReduction.transform_expand on the tensor cores (64 -> 128 -> ReLU -> d_out, L2-normalised rows)
bool expand_tc_supported(const sd_mlp *mlp);
// the feature rows of the head's second layer on explicit hidden rows: out[r, :D] = W_out[1:] . h[r] + b_out[1:] * scale[r]
// (h [R,128] fp32, scale [R] or NULL = 1): second half of the hidden-composite render (field_tc.cu, cmma == 2)
int launch_head2(const sd_mlp *mlp, const float *h, const float *scale, long long R, float *out, cudaStream_t st);
int launch_expand_tc(const sd_mlp *mlp, const float *f, long long N, float *out, cudaStream_t st);

// sampling.cu: importance + depth samples of the fine pass, merged with the coarse depths and sorted, in one launch
int launch_fine_merge(const float *rays, long long R, int r_dim, const float *weights, const float *z_coarse, const float *depth,
                      int Kc, const float *u0, const float *u1, int Kfi, const float *noise, int Kfd, float depth_std,
                      int lindisp, float *z_all, cudaStream_t st);

}  // namespace sd
