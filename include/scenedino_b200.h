/*
 * scenedino_b200.h -- C ABI of the B200-native feature-field query-and-render path.
 *
 * The reference (tum-vision/scenedino) is pure Python/PyTorch and has no FFI layer; the "plugin
 * surface" it exposes for this path is a set of Python methods.  Each entry
 * point below replaces the body of one of those methods and cites it (paths relative to the
 * reference root).  The Python shim in scenedino_b200/ binds these symbols with ctypes and keeps the
 * reference's method names, argument meaning and error behaviour.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers into buffers allocated and owned by the caller
 *     (PyTorch's caching allocator in the shim).  The library never allocates or frees
 *     caller-visible memory; scratch space is passed in (see *_workspace_bytes).
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; 0 = legacy default).
 *   - Return value: 0 on success, a negative sd_status otherwise; sd_last_error() returns a
 *     thread-local message.  Nothing throws, nothing calls exit().
 *   - Row-major everywhere.  "rays" rows are [origin(3) dir(3) near far ...] with r_dim >= 8 floats
 *     per ray (common/ray_sampler.py:476-484).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point returns
 *     SD_ERR_CUDA.
 */
#ifndef SCENEDINO_B200_H
#define SCENEDINO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 5

typedef enum sd_status {
    SD_OK = 0,
    SD_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    SD_ERR_CUDA = -2,        /* CUDA runtime error (message has the CUDA error string) */
    SD_ERR_WORKSPACE = -3    /* workspace too small */
} sd_status;

typedef enum sd_dtype { SD_F32 = 0, SD_F16 = 1 } sd_dtype;

/* MLP arithmetic.  SD_MLP_FP32: fp32 FFMA on CUDA cores (parity mode, rel 1e-4).
 * SD_MLP_F16_TC: 16-bit operands on tcgen05 tensor cores (kind::f16) with fp32 accumulation in TMEM
 * (reduced-precision mode, rel 2e-2).  The operands are IEEE half, the dtype the reference itself runs
 * the head in under torch autocast (training/base_trainer.py:223): same tensor-core rate as bf16, 8x
 * smaller rounding error -- single-pass bf16 sits AT the 2e-2 bar for a 295-term contraction.
 * SD_MLP_F32_TC: the rel-1e-4 bar ON the tensor cores (point queries on a scene projected by sd_field_project_x3): every
 * operand is an fp16 pair hi + lo (~22 significant bits), every product the three kind::f16 products hi.hi + lo.hi + hi.lo,
 * fp32 accumulation in TMEM, fp32 biases / softplus in the epilogues.  Other entry points treat it like SD_MLP_FP32. */
typedef enum sd_precision { SD_MLP_FP32 = 0, SD_MLP_F16_TC = 1, SD_MLP_F32_TC = 2 } sd_precision;

/* What BTSNet.encode stashes for ONE batch element (models/bts.py:246-257), with the feature map
 * re-laid out channels-last by sd_featmap_pack. */
typedef struct sd_scene {
    const void  *feat;          /* [nv_f, Hf, Wf, C] channels-last, feat_dtype            */
    int          feat_dtype;    /* sd_dtype                                               */
    int          nv_f, C, Hf, Wf;
    const float *K_f;           /* [nv_f,3,3] normalised intrinsics   (bts.py:247)         */
    const float *w2c_f;         /* [nv_f,4,4] world->camera           (bts.py:248)         */
    const float *rgb;           /* [nv_c,3,Hc,Wc] fp32 planar, as grid_c_imgs (bts.py:252) */
    int          nv_c, Hc, Wc;
    const float *K_c;           /* [nv_c,3,3] (bts.py:253) */
    const float *w2c_c;         /* [nv_c,4,4] (bts.py:254) */
    float        d_min, d_max;  /* z_near, z_far of the encoding (bts.py:60)               */
    int          inv_z;         /* bts.py:64                                               */
    int          num_freqs;     /* PositionalEncoding (positional_encoding.py:49-66)       */
    float        freq_factor;
    int          include_input;
    int          learn_empty;   /* bts.py:311-319                                          */
    const float *empty_feature; /* [C] or NULL                                             */
    const void  *feat_proj;     /* optional: blob written by sd_field_project for the head these queries
                                   will use (NULL = absent); enables the projected-map tile kernel */
    const void  *feat_proj_x3;  /* optional: blob written by sd_field_project_x3 (SD_MLP_F32_TC queries) */
} sd_scene;

/* ResnetFC head with n_blocks = 0 (models/prediction_heads/resnetfc.py:90-96,162-199).
 * `packed` is the blob written by sd_mlp_pack (transposed/padded fp32 weights for the CUDA-core
 * path and fp16 UMMA shared-memory images for the tcgen05 path). */
typedef struct sd_mlp {
    const void *packed;
    int         d_in, d_hidden, d_out;
    int         precision;      /* sd_precision */
} sd_mlp;

/* NeRFRenderer options that change arithmetic (renderer/nerf.py:73-119). */
typedef struct sd_render_cfg {
    int   lindisp;
    int   hard_alpha_cap;
    int   white_bkgd;
} sd_render_cfg;

/* ---- library -------------------------------------------------------------------------------- */
int         sd_abi_version(void);
const char *sd_last_error(void);
/* SM count of the current device, or a negative sd_status. */
int         sd_device_sm_count(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
long long   sd_launch_count(void);

/* Measurement hook: the next launch of the path's dominant kernel (the fused field kernel of sd_query_points /
 * sd_render_pass) made by the calling host thread records the two caller-owned cudaEvent_t handles around
 * itself, on the stream it is launched on.  bench.py's roofline uses it; NULL handles switch it off. */
int         sd_profile_next_kernel(void *ev_start, void *ev_stop);

/* ---- one-off packing ------------------------------------------------------------------------ */
/* BTSNet.encode stash (bts.py:214-257): [n_img, C, H, W] fp32 planar -> [n_img, H, W, C]
 * channels-last in dst_dtype.  Replaces the no-op F.interpolate copy at bts.py:217-222. */
int    sd_featmap_pack(const float *nchw, int n_img, int C, int H, int W, void *nhwc, int dst_dtype,
                       void *stream);
/* Size of / writer for the packed weight blob of an MLP head (nn.Linear layout in, fp32). */
size_t sd_mlp_pack_bytes(int d_in, int d_hidden, int d_out);
int    sd_mlp_pack(const float *w_in, const float *b_in, const float *w_out, const float *b_out,
                   int d_in, int d_hidden, int d_out, void *packed, void *stream);

/* Pushes the channels-last fp16 map through the feature columns of the head's first layer, once per
 * encode: P[texel] = W_in[:, :C] . F[texel] (bilinear sampling, bts.py:299-309, and lin_in,
 * resnetfc.py:162-163, are both linear, so sampling P equals lin_in of the sampled features).  With
 * scene->feat_proj set to the result, sd_query_points (SD_MLP_F16_TC, with workspace) interpolates the
 * 128 hidden pre-activations on the tensor cores straight from TMA-fetched tiles of P.  The blob is
 * tied to `mlp` (and to empty_feature when learn_empty); `proj` must be 1024-byte aligned. */
size_t sd_field_project_bytes(const sd_scene *scene);
int    sd_field_project(const sd_scene *scene, const sd_mlp *mlp, void *proj, size_t proj_bytes, void *stream);

/* The same once-per-encode projection for SD_MLP_F32_TC queries: P = W_in[:, :C] . F in fp32 (CUDA cores) from the fp32
 * channels-last map, stored as an fp16 (hi, lo) pair of maps behind (hi, lo) UMMA images of the code block of W_in and of
 * W_out.  Set scene->feat_proj_x3 to the result.  proj 1024-byte aligned, sd_field_project_x3_bytes(scene) bytes. */
size_t sd_field_project_x3_bytes(const sd_scene *scene);
int    sd_field_project_x3(const sd_scene *scene, const sd_mlp *mlp, void *proj, size_t proj_bytes, void *stream);

/* ---- point ops, one per reference function (unfused; for 1:1 parity tests) -------------------- */
/* pts_into_camera + project_to_image + outside_frustum (common/cameras/pinhole.py:40-112) for one
 * camera.  xy [N,2] is unclamped; invalid [N] is 0/1. */
int sd_project_points(const float *K, const float *w2c, const float *xyz, long long N, float *xy,
                      float *z, unsigned char *invalid, void *stream);
/* BTSNet.sample_features (bts.py:271-328), nv_f == 1: feat [N, C+code] fp32, invalid [N]. */
int sd_sample_features(const sd_scene *scene, const float *xyz, long long N, float *feat,
                       unsigned char *invalid, void *stream);
/* BTSNet.sample_colors (bts.py:330-358), default options (bilinear, no combine / frame filter /
 * flow): rgb [N,3*nv_c] (view-major inside a row, as bts.py:559-561 lays it out), invalid [N,nv_c]
 * 0/1 computed on the clamped coordinates like the reference. */
int sd_sample_colors(const sd_scene *scene, const float *xyz, long long N, float *rgb,
                     unsigned char *invalid, void *stream);
/* ResnetFC.forward (resnetfc.py:135-203): x [N,d_in] -> out [N,d_out]. */
int sd_mlp_forward(const sd_mlp *mlp, const float *x, long long N, float *out, void *stream);

/* ---- BTSNet.forward (bts.py:476-595) --------------------------------------------------------- */
/* sigma [N], dino [N,d_out-1], rgb [N,3*nv_c], invalid [N,nv_c] (fp32 0/1, = invalid_colors |
 * all(invalid_features), bts.py:566-569), invalid_feat [N] (0/1).  Any output may be NULL.
 * `workspace` is optional scratch (sd_query_workspace_bytes; may be NULL / 0): with it the tensor-core path
 * first sorts the point indices by the 8x8-texel block of the feature map they project to, so that the
 * rows of a tile share texels (results per point do not depend on the order). */
size_t sd_query_workspace_bytes(const sd_scene *scene, const sd_mlp *mlp, long long N);
int sd_query_points(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N,
                    float *sigma, float *dino, float *rgb, float *invalid,
                    unsigned char *invalid_feat, void *workspace, size_t workspace_bytes, void *stream);

/* The same query when the points and the cameras have not changed since an earlier sd_query_points call that used this
 * workspace (the SSC evaluation queries one fixed voxel grid frame after frame, sscbench/evaluate_model_sscbench.py:
 * 270-279): the texel sort is reused, only the tile kernel runs -- on the scene's CURRENT feature map / projection.
 * Needs a projected scene and SD_MLP_F16_TC.  The frustum mask (invalid_feat) of the earlier call stays valid and is
 * not rewritten. */
int sd_query_points_sorted(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N,
                           float *sigma, float *dino, float *rgb, float *invalid, void *workspace,
                           size_t workspace_bytes, void *stream);

/* The same query with the 64-d features left in TEXEL-BIN ORDER (the order the tile kernel walks the points in) for a
 * consumer that takes a permutation (sd_ssc_head): dino_binned [N,64], row r holds the features of
 * point perm[r]; sigma [N] and invalid_feat [N] stay in the caller's order.  The rows of a tile are consecutive rows of
 * dino_binned, so they leave the SM as TMA tile stores (cp.async.bulk.tensor, 32 rows x 128 B each) instead of one
 * scattered 16-byte store per thread and piece: the reference's chunk loop (sscbench/evaluate_model_sscbench.py:711-717)
 * only ever hands these rows to encoder.expand_dim + the downstream head (models/bts.py:584-592), row by row.
 * reuse_sorted != 0: like sd_query_points_sorted (the workspace holds the sort of the same points and cameras;
 * invalid_feat is not rewritten).  Needs a projected scene, SD_MLP_F16_TC, a 64-d head, dino_binned 16-byte aligned, and
 * enough points for the sorted tile path (16 per texel bin of the map). */
int sd_query_points_binned(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N, float *sigma,
                           float *dino_binned, unsigned int *perm, unsigned char *invalid_feat, void *workspace,
                           size_t workspace_bytes, int reuse_sorted, void *stream);

/* ---- NeRFRenderer sampling (renderer/nerf.py:121-228) --------------------------------------- */
/* sample_coarse (nerf.py:121-141).  u [R,Kc] = torch.rand_like draw, lin [Kc] = torch.linspace. */
int sd_sample_coarse(const float *rays, long long R, int r_dim, const float *u, const float *lin,
                     int Kc, int lindisp, float *z, void *stream);
/* sample_fine (nerf.py:181-212).  u0,u1 [R,Kf]; inds [R,Kf] optional (parity). */
int sd_sample_fine(const float *rays, long long R, int r_dim, const float *weights, int Kc,
                   const float *u0, const float *u1, int Kf, int lindisp, float *z, int *inds,
                   void *stream);
/* sample_fine_depth (nerf.py:214-228).  noise [R,Kfd] = torch.randn_like draw. */
int sd_sample_fine_depth(const float *rays, long long R, int r_dim, const float *depth,
                         const float *noise, int Kfd, float depth_std, float *z, void *stream);
/* sample_coarse_from_dist (nerf.py:143-179), unsorted like the reference. */
int sd_sample_coarse_from_dist(long long R, const float *weights, const float *z_samp, int Kp,
                               const float *u0, const float *u1, int Kc, int lindisp, float *z,
                               int *inds, void *stream);
/* torch.sort(z, dim=-1) values, in place (nerf.py:490,522).  K <= 1024. */
int sd_sort_rows(float *z, long long R, int K, void *stream);

/* ---- NeRFRenderer.composite (nerf.py:230-449) ------------------------------------------------ */
/* Core arithmetic only (nerf.py:246-249,376-405,418-421): z,sigma [R,K]; feat [R,K,D];
 * rgb [R,K,Crgb] -> weights,alphas [R,K]; depth [R]; dino [R,D]; rgb_out [R,Crgb]. */
int sd_composite(const float *z, const float *sigma, const float *feat, const float *rgb,
                 long long R, int K, int D, int Crgb, const sd_render_cfg *cfg, float *weights,
                 float *alphas, float *depth, float *dino, float *rgb_out, void *stream);

/* One full composite() call for one scene: points along the rays at z [R,K], field query,
 * compositing.  Per-ray outputs: depth [R], dino [R,D], rgb_out [R,3nv_c].  Per-sample outputs
 * (any may be NULL): weights, alphas [R,K]; invalid [R,K,nv_c] fp32; invalid_feat [R,K];
 * rgb_samps [R,K,3nv_c]; sigma [R,K].  With mlp->precision == SD_MLP_F16_TC and a supported
 * shape this is ONE fused kernel and needs no workspace. */
size_t sd_render_workspace_bytes(const sd_scene *scene, const sd_mlp *mlp, long long R, int K);
int    sd_render_pass(const sd_scene *scene, const sd_mlp *mlp, const sd_render_cfg *cfg,
                      const float *rays, long long R, int r_dim, const float *z, int K,
                      float *depth, float *dino, float *rgb_out, float *weights, float *alphas,
                      float *invalid, unsigned char *invalid_feat, float *rgb_samps, float *sigma,
                      void *workspace, size_t workspace_bytes, void *stream);

/* ---- NeRFRenderer.forward (renderer/nerf.py:451-539) for one scene, in ONE call ---------------------------------------
 * Coarse sampling -> coarse pass -> (when n_fine > 0) importance + depth samples, merge, sort -> fine pass, launched back
 * to back on `stream` without returning to the caller: sample_coarse kernel, fused field kernel, ONE kernel for
 * sample_fine + sample_fine_depth + cat + sort (fine_merge_kernel), fused field kernel.  The random draws are inputs
 * (u_coarse [R,n_coarse] and lin [n_coarse] as in sd_sample_coarse; u_fine0 / u_fine1 [R, n_fine - n_fine_depth] as in
 * sd_sample_fine; n_depth [R, n_fine_depth] as in sd_sample_fine_depth), so a caller that draws them with the reference's
 * torch calls in the reference's order reproduces its samples.  Results are those of the unfused entry points, bit for
 * bit.  `fine` may be NULL when n_fine == 0.  Any pointer inside sd_render_out may be NULL. */
typedef struct sd_render_out {
    float *depth;                /* [R]           */
    float *dino;                 /* [R, d_out-1]  */
    float *rgb;                  /* [R, 3*nv_c]   */
    float *weights, *alphas;     /* [R, K]        */
    float *z_samps;              /* [R, K]        */
    float *invalid;              /* [R, K, nv_c]  */
    unsigned char *invalid_feat; /* [R, K]        */
    float *rgb_samps;            /* [R, K, 3*nv_c] */
} sd_render_out;
typedef struct sd_sampling {
    int   n_coarse, n_fine, n_fine_depth;
    float depth_std;
} sd_sampling;
size_t sd_render_rays_workspace_bytes(const sd_scene *scene, const sd_mlp *mlp, const sd_sampling *samp, long long R);
int    sd_render_rays(const sd_scene *scene, const sd_mlp *mlp, const sd_render_cfg *cfg, const sd_sampling *samp,
                      const float *rays, long long R, int r_dim, const float *u_coarse, const float *lin,
                      const float *u_fine0, const float *u_fine1, const float *n_depth, const sd_render_out *coarse,
                      const sd_render_out *fine, void *workspace, size_t workspace_bytes, void *stream);

/* ---- section 8f-1: MlpDimReduction.transform_expand (backbones/dino/dim_reduction.py:22-25) --- */
/* `mlp` packs linear_in / linear_out; out [N,d_out] is L2-normalised per row (F.normalize).
 * mlp->precision == SD_MLP_F16_TC and a 64 -> 128 -> k*128 (k <= 8) head: tcgen05 kernel (fp16 operands, fp32
 * accumulation and normalisation, rel 2e-2; f and out 16-byte aligned); anything else: fp32 CUDA-core kernel (rel 1e-4). */
int sd_expand_dim(const sd_mlp *mlp, const float *f, long long N, float *out, void *stream);

/* ---- section 8f-2: the unsupervised SSC head, fused with the expansion that feeds it ---------------------------------
 * SemanticHead.forward(mode = "stego_kmeans") (downstream_head/semantic_head.py:107-112) on
 * MlpDimReduction.transform_expand (dim_reduction.py:22-25) of the 64-d features of a field query -- what
 * BTSNet.forward(predict_segmentation=True) does per voxel (models/bts.py:584-592) -- without ever forming the 768-d
 * rows: everything between the two ReLUs is linear and is folded once per model by sd_ssc_head_pack (ssc_head.cu).
 *   expand:  w1e [128,64], b1e [128], w2e [d_full,128], b2e [d_full]              (linear_in / linear_out)
 *   STEGO:   wl [64,d_full], bl [64]; wn1 [d_mid,d_full], bn1 [d_mid]; wn2 [64,d_mid], bn2 [64]   (1x1-conv weights,
 *            StegoClusterHead.linear_path / nonlinear_path, semantic_head.py:285-305; eval mode: no dropout)
 *   k-means: centres [n_cls,64] (KMeansParamHead.cluster_centers), lut [n_cls] int64 (pseudo_assignment), :308-373
 * all fp32 device pointers in nn.Module layout.  d_red = 64, d_lat = 128, d_code = 64, d_mid a multiple of 128 (<= 1024),
 * n_cls <= 32.  `packed` must be 1024-byte aligned. */
size_t sd_ssc_head_pack_bytes(int d_red, int d_lat, int d_full, int d_mid, int d_code, int n_cls);
int    sd_ssc_head_pack(const float *w1e, const float *b1e, const float *w2e, const float *b2e, const float *wl,
                        const float *bl, const float *wn1, const float *bn1, const float *wn2, const float *bn2,
                        const float *centres, const long long *lut, int d_red, int d_lat, int d_full, int d_mid,
                        int d_code, int n_cls, void *packed, void *stream);
/* f [N,64] fp32 (16-byte aligned) -> seg [N] (labels after the pseudo-label LUT, what SemanticHead.forward returns),
 * pseudo [N] (cluster ids), scores [N,n_cls] (cosine scores): any of the three may be NULL.  perm (or NULL): input row r
 * stands for voxel perm[r], i.e. outputs are written at perm[r].  tcgen05 tensor cores, fp16 operands, fp32 accumulation:
 * scores within 2e-2 of the reference (measured ~3e-4), labels equal wherever the top-2 gap exceeds that. */
int    sd_ssc_head(const void *packed, int d_mid, int n_cls, const float *f, const unsigned int *perm, long long N,
                   unsigned char *seg, unsigned char *pseudo, float *scores, void *stream);

/* ---- PositionalEncoding.forward (common/positional_encoding.py:68-80) on its own --------------------------------------- */
/* x [N,d_in] -> out [N, (include_input ? d_in : 0) + 2*num_freqs*d_in]: [x | per frequency k: sin(f_k x) (d_in), cos(f_k x)
 * (d_in)], f_k = freq_factor * 2^k, cos evaluated as sin(. + pi/2) like the reference.  (Inside the field kernels the code
 * is fused; this entry point is the 1:1 counterpart of the module's forward.) */
int    sd_positional_encoding(const float *x, long long N, int d_in, int num_freqs, float freq_factor, int include_input,
                              float *out, void *stream);

/* ---- section 8f-3: rays of whole views ------------------------------------------------------ */
/* Replaces util.gen_rays / util.unproj_map (common/util.py:253-285, 113-158) and the ray half of
 * ImageRaySampler.sample (common/ray_sampler.py:439-486) for ONE batch element:
 *   c2w [V,4,4] camera-to-world poses, proj [V,3,3] normalised intrinsics (fx, fy, cx, cy are read from
 *   [0][0], [1][1], [0][2], [1][2]), frame_ids [V] or NULL (then 0..V-1), all device pointers;
 *   rays [V*H*W, 11] = origin(3) direction(3) z_near z_far frame-id pixel-x pixel-y, pixels row-major per view
 *   (16-byte aligned).  norm_dir != 0 normalises the directions.  x_shift / y_shift are added to the
 *   pixel-centre coordinates (the reference's xy_offset * pixel size, in NDC units; 0 = none).
 * Bit-identical to torch's CPU result for the same inputs.  H, W >= 2. */
int sd_gen_rays(const float *c2w, const float *proj, const float *frame_ids, int V, int H, int W,
                float z_near, float z_far, int norm_dir, float x_shift, float y_shift, float *rays,
                void *stream);

/* Voxel centres of an SSC grid in the camera frame: what sscbench/evaluate_model_sscbench.py:270-278 builds on the host
 * once (generate_point_grid, sscbench/point_utils.py:46-67) and uploads -- 25 MB for 256 x 256 x 32 -- made on the device
 * instead.  origin [3] (voxel (0,0,0) corner, lidar frame) and T [3][4] (row-major float64 lidar -> camera, the
 * calibration's T_velo_2_cam) are HOST pointers read before the call returns.  xyz [(x1-x0)*ny*nz, 3] fp32, flattened
 * 'ij' order (x slowest); [x0, x1) selects a slab of x indices (voxel-slab sharding).  centre = origin + size*idx + size*0.5
 * with the fp32 origin and index but the size as the double it is in the reference (a Python float inside the numba-compiled
 * TSDFVolume.vox2world, sscbench/fusion.py:205-219): evaluated in double, rounded to fp32 once; then the rigid transform as a
 * float64 dot product rounded once (rigid_transform, fusion.py:407-411; .float() at evaluate_model_sscbench.py:277).
 * Bit-identical to the reference's own grid (tests/golden/voxel_grid.npz, made by running the reference's functions). */
int sd_gen_voxel_grid(const float *origin, double voxel_size, int nx, int ny, int nz, int x0, int x1,
                      const double *T, float *xyz, void *stream);

/* ---- diagnostics (timing experiments; not part of the data path) ----------------------------------------------------
 * With the environment variable SD_TC_DEBUG & 8192 set when the library is loaded, CTA 0 of the tensor-core kernels
 * records clock64 stamps per warp role and tile; these calls copy the trace of the last launch to HOST buffers:
 * sd_debug_read_trace: field_tc_kernel, [4 roles][64 tiles][8 events] long long; sd_debug_read_trace_bin: field_bin_kernel,
 * [8][64][8] long long; sd_debug_read_cta_ns: field_bin_kernel, [256 CTAs][start, end] %globaltimer;
 * sd_debug_read_tiles_bin: field_bin_kernel, [4 CTAs][512 tiles][clock64 at the layer-1 issuer's tile start, operand chunks of
 * the tile] long long (entries 510 / 511 of a CTA: kernel start / end). */
int sd_debug_read_trace(long long *host_out);
int sd_debug_read_trace_bin(long long *host_out);
int sd_debug_read_cta_ns(unsigned long long *host_out);
int sd_debug_read_tiles_bin(long long *host_out);

/* ---- section 8f-4: backward pass of the training step (training/base_trainer.py:223-255) ---------------------------------
 * The fused kernels above are forward-only.  In training mode (autograd on) the Python surface runs the path unfused --
 * sd_sample_features, the head as plain torch modules (its GEMMs and their gradients are library GEMMs), sd_composite --
 * and these two entries supply the gradients of the two custom stages.
 *
 * sd_composite_bwd: gradients of sd_composite (nerf.py:376-421) for the same z / sigma / feat / rgb and render options.
 *   g_depth [R], g_dino [R,D], g_rgb_out [R,Crgb], g_weights [R,K], g_alphas [R,K]: upstream gradients (NULL = zero);
 *   g_sigma [R,K], g_feat [R,K,D], g_rgb [R,K,Crgb]: outputs (any may be NULL).  K <= 256.
 * sd_sample_features_bwd: gradient of BTSNet.sample_features (bts.py:299-319) with respect to the encoder feature map:
 *   g_feat [N, C + code] is scattered (atomic adds) into g_map [Hf,Wf,C] fp32 channels-last, which the caller zeroes;
 *   rows the forward replaced by the learned empty feature add to g_empty [C] instead. */
int sd_composite_bwd(const float *z, const float *sigma, const float *feat, const float *rgb, long long R, int K, int D,
                     int Crgb, const sd_render_cfg *cfg, const float *g_depth, const float *g_dino, const float *g_rgb_out,
                     const float *g_weights, const float *g_alphas, float *g_sigma, float *g_feat, float *g_rgb, void *stream);
int sd_sample_features_bwd(const sd_scene *scene, const float *xyz, long long N, const float *g_feat, float *g_map,
                           float *g_empty, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCENEDINO_B200_H */
