// extern "C" surface of libscenedino_b200.so: error plumbing and the composite entry points that
// dispatch between the fp32 CUDA-core path (field_simt.cu) and the fused tcgen05 path (field_tc.cu).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "launch.h"

namespace sd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return SD_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int debug_mask() {
    static const int mask = [] { const char *e = getenv("SD_TC_DEBUG"); return e ? atoi(e) : 0; }();
    return mask;
}

static std::mutex g_once_mutex;

int device_once(DeviceOnce &seen, int *sm_count, bool *first) {
    std::lock_guard<std::mutex> lock(g_once_mutex);   // launchers may be entered from several host threads
    int dev = 0;
    SD_CUDA_OK(cudaGetDevice(&dev));
    SD_REQUIRE(dev >= 0 && dev < SD_MAX_DEVICES, "device ordinal %d is out of range (max %d devices per process)", dev, SD_MAX_DEVICES);
    *first = seen.sm_count[dev] == 0;
    if (*first) {
        int n = 0;
        SD_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        SD_REQUIRE(n > 0, "device %d reports no multiprocessors", dev);
        seen.sm_count[dev] = n;
    }
    *sm_count = seen.sm_count[dev];
    return SD_OK;
}

// sd_profile_next_kernel: a pair of caller-owned events recorded around the next launch of a field kernel
static thread_local cudaEvent_t g_prof0 = nullptr, g_prof1 = nullptr;
void profile_before(cudaStream_t st) {
    if (g_prof0) cudaEventRecord(g_prof0, st);
    g_prof0 = nullptr;
}
void profile_after(cudaStream_t st) {
    if (g_prof1) cudaEventRecord(g_prof1, st);
    g_prof1 = nullptr;
}

}  // namespace sd

using namespace sd;

extern "C" int sd_abi_version(void) { return SD_ABI_VERSION; }
extern "C" const char *sd_last_error(void) { return g_err; }
extern "C" long long sd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int sd_profile_next_kernel(void *ev_start, void *ev_stop) {
    g_prof0 = reinterpret_cast<cudaEvent_t>(ev_start);
    g_prof1 = reinterpret_cast<cudaEvent_t>(ev_stop);
    return SD_OK;
}

extern "C" int sd_device_sm_count(void) {
    int dev = 0, n = 0;
    SD_CUDA_OK(cudaGetDevice(&dev));
    SD_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

extern "C" int sd_mlp_forward(const sd_mlp *mlp, const float *x, long long N, float *out, void *stream) {
    SD_REQUIRE(mlp, "sd_mlp_forward: mlp is NULL");
    SD_REQUIRE(N >= 0, "sd_mlp_forward: bad N");
    // ResnetFC.forward asserts the input width (resnetfc.py:155); widths are carried by sd_mlp here.
    if (mlp->precision == SD_MLP_F16_TC) return launch_mlp_tc(mlp, x, N, out, (cudaStream_t)stream);
    return launch_mlp_simt(mlp, x, N, out, false, (cudaStream_t)stream);
}

extern "C" size_t sd_query_workspace_bytes(const sd_scene *scene, const sd_mlp *mlp, long long N) {
    if (!scene || !mlp || N <= 0) return 0;
    // only the tensor-core path reorders the points; below a few tiles per SM it is not worth three launches
    if ((mlp->precision != SD_MLP_F16_TC && mlp->precision != SD_MLP_F32_TC) || N < 65536) return 0;
    return bin_workspace_bytes(scene->Hf, scene->Wf, N);
}

static int query_points_impl(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N,
                             float *sigma, float *dino, float *rgb, float *invalid,
                             unsigned char *invalid_feat, void *workspace, size_t workspace_bytes, void *stream, bool reuse_sorted,
                             float *dino_binned = nullptr, unsigned int *perm_out = nullptr);

extern "C" int sd_query_points(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N,
                               float *sigma, float *dino, float *rgb, float *invalid,
                               unsigned char *invalid_feat, void *workspace, size_t workspace_bytes, void *stream) {
    return query_points_impl(scene, mlp, xyz, N, sigma, dino, rgb, invalid, invalid_feat, workspace, workspace_bytes, stream, false);
}

extern "C" int sd_query_points_sorted(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N,
                                      float *sigma, float *dino, float *rgb, float *invalid, void *workspace,
                                      size_t workspace_bytes, void *stream) {
    SD_REQUIRE(scene && mlp && ((scene->feat_proj && mlp->precision == SD_MLP_F16_TC) || (scene->feat_proj_x3 && mlp->precision == SD_MLP_F32_TC)) &&
                   workspace && sd_query_workspace_bytes(scene, mlp, N) > 0 && workspace_bytes >= sd_query_workspace_bytes(scene, mlp, N),
               "sd_query_points_sorted: needs a projected scene, SD_MLP_F16_TC / SD_MLP_F32_TC and the workspace of an earlier sd_query_points call");
    return query_points_impl(scene, mlp, xyz, N, sigma, dino, rgb, invalid, nullptr, workspace, workspace_bytes, stream, true);
}

extern "C" int sd_query_points_binned(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N, float *sigma,
                                      float *dino_binned, unsigned int *perm, unsigned char *invalid_feat, void *workspace,
                                      size_t workspace_bytes, int reuse_sorted, void *stream) {
    SD_REQUIRE(scene && mlp && ((scene->feat_proj && mlp->precision == SD_MLP_F16_TC) || (scene->feat_proj_x3 && mlp->precision == SD_MLP_F32_TC)) &&
                   workspace && sd_query_workspace_bytes(scene, mlp, N) > 0 && workspace_bytes >= sd_query_workspace_bytes(scene, mlp, N),
               "sd_query_points_binned: needs a projected scene, SD_MLP_F16_TC / SD_MLP_F32_TC and a workspace of sd_query_workspace_bytes");
    SD_REQUIRE(dino_binned && mlp->d_out == 65, "sd_query_points_binned: needs dino_binned and a 64-d feature head");
    if (N == 0) return SD_OK;
    return query_points_impl(scene, mlp, xyz, N, sigma, nullptr, nullptr, nullptr, reuse_sorted ? nullptr : invalid_feat, workspace,
                             workspace_bytes, stream, reuse_sorted != 0, dino_binned, perm);
}

static int query_points_impl(const sd_scene *scene, const sd_mlp *mlp, const float *xyz, long long N,
                             float *sigma, float *dino, float *rgb, float *invalid,
                             unsigned char *invalid_feat, void *workspace, size_t workspace_bytes, void *stream, bool reuse_sorted,
                             float *dino_binned, unsigned int *perm_out) {
    FieldParams fp;
    int rc = make_field_params(scene, &fp);
    if (rc) return rc;
    SD_REQUIRE(mlp, "sd_query_points: mlp is NULL");
    SD_REQUIRE(N >= 0 && (xyz || N == 0), "sd_query_points: bad points");
    PointSrc src = {xyz, nullptr, nullptr, 0, 1};
    if (mlp->precision == SD_MLP_F16_TC) {
        TcOut o = {};
        o.sigma = sigma; o.dino = dino; o.rgb = rgb; o.invalid = invalid; o.invalid_feat = invalid_feat;
        o.dino_binned = dino_binned; o.perm_out = perm_out;
        BinOrder order = {};
        const size_t need = sd_query_workspace_bytes(scene, mlp, N);
        if (need && workspace && workspace_bytes >= need) {   // walk the points bin by bin of the feature map
            // projected scene: interpolation on the tensor cores from TMA tiles of P (field_bin.cu).  Worth it when
            // the bins are well filled (a chunk of 64 texels is fetched per bin a tile touches): the sort then also
            // leaves the per-point geometry (coordinates, bilinear weights, frustum mask) at the sorted positions
            const bool tile = scene->feat_proj && bin_kernel_supported(scene, mlp) &&
                              N >= 16ll * ((scene->Wf - 1) / SD_BIN + 1) * ((scene->Hf - 1) / SD_BIN + 1);
            SD_REQUIRE(tile || !reuse_sorted, "sd_query_points_sorted: this query does not take the sorted tile path");
            rc = launch_bin_points(fp, xyz, N, workspace, workspace_bytes, &order, (cudaStream_t)stream, tile,
                                   tile ? invalid_feat : nullptr, reuse_sorted);
            if (rc) return rc;
            if (tile && order.has_geo) return launch_field_bin(scene, fp, xyz, N, mlp, order, o, (cudaStream_t)stream);
        }
        SD_REQUIRE(!dino_binned, "sd_query_points_binned: this query does not take the sorted tile path (too few points for the map)");
        return launch_field_tc(fp, src, N, mlp, nullptr, o, (cudaStream_t)stream, order.perm, scene->feat_proj);
    }
    if (mlp->precision == SD_MLP_F32_TC) {
        // rel-1e-4 on the tensor cores: the same texel sort + tile kernel with every operand as an fp16 (hi, lo) pair
        // (field_bin_x3.cu).  Queries too small for the sorted tile path take the fp32 CUDA-core kernel below -- same bar.
        const size_t need = sd_query_workspace_bytes(scene, mlp, N);
        const bool tile = need && workspace && workspace_bytes >= need && scene->feat_proj_x3 && bin_kernel_supported_x3(scene, mlp) &&
                          N >= 16ll * ((scene->Wf - 1) / SD_BIN + 1) * ((scene->Hf - 1) / SD_BIN + 1);
        SD_REQUIRE(tile || (!reuse_sorted && !dino_binned), "sd_query_points: this SD_MLP_F32_TC query does not take the sorted tile path");
        if (tile) {
            TcOut o = {};
            o.sigma = sigma; o.dino = dino; o.rgb = rgb; o.invalid = invalid; o.invalid_feat = invalid_feat;
            o.dino_binned = dino_binned; o.perm_out = perm_out;
            BinOrder order = {};
            rc = launch_bin_points(fp, xyz, N, workspace, workspace_bytes, &order, (cudaStream_t)stream, true, invalid_feat, reuse_sorted);
            if (rc) return rc;
            SD_REQUIRE(order.has_geo, "sd_query_points: the texel sort left no geometry records");
            return launch_field_bin_x3(scene, fp, xyz, N, mlp, order, o, (cudaStream_t)stream);
        }
        SD_REQUIRE(scene->feat_dtype == SD_F32, "sd_query_points: SD_MLP_F32_TC below the tile-path size needs the fp32 map");
    }
    SD_REQUIRE(mlp->precision == SD_MLP_FP32 || mlp->precision == SD_MLP_F32_TC, "sd_query_points: unknown precision %d", mlp->precision);
    SimtOut out = {};
    out.sigma = sigma; out.dino = dino; out.rgb = rgb; out.invalid = invalid; out.invalid_feat = invalid_feat;
    return launch_field_simt(MODE_QUERY_, fp, src, N, mlp, out, (cudaStream_t)stream);
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

extern "C" size_t sd_render_workspace_bytes(const sd_scene *scene, const sd_mlp *mlp, long long R, int K) {
    if (!scene || !mlp || R <= 0 || K <= 0) return 0;
    const size_t N = (size_t)R * K;
    if (mlp->precision == SD_MLP_F16_TC && tc_supported(scene, mlp, K)) {
        const int mode = tc_render_mode(scene, mlp);
        // projected scene: the composite runs on the tensor cores and nothing per sample touches memory; with more than 64
        // feature outputs the per-ray sums of the 128 hidden units (+ the sum of the weights) wait here for the head2 kernel
        if (mode == 2) return align256((size_t)R * 128 * 4) + align256((size_t)R * 4);
        if (mode == 1) return 0;
        // unprojected scene: the per-sample colours make a round trip through memory (L2-sized tiles of it)
        return scene->nv_c > 0 ? align256(N * 3 * (size_t)scene->nv_c * 4) : 0;
    }
    const int D = mlp->d_out - 1;
    // sigma [N], dino [N,D], rgb [N,3nv_c]
    return align256(N * 4) + align256(N * D * 4) + align256(N * 3 * (size_t)(scene->nv_c > 0 ? scene->nv_c : 1) * 4);
}

extern "C" int sd_render_pass(const sd_scene *scene, const sd_mlp *mlp, const sd_render_cfg *cfg,
                              const float *rays, long long R, int r_dim, const float *z, int K, float *depth,
                              float *dino, float *rgb_out, float *weights, float *alphas, float *invalid,
                              unsigned char *invalid_feat, float *rgb_samps, float *sigma, void *workspace,
                              size_t workspace_bytes, void *stream) {
    FieldParams fp;
    int rc = make_field_params(scene, &fp);
    if (rc) return rc;
    SD_REQUIRE(mlp && cfg, "sd_render_pass: null pointer");
    SD_REQUIRE(R >= 0 && K > 0 && r_dim >= 8, "sd_render_pass: bad shape (rays need >= 8 columns)");
    if (R == 0) return SD_OK;
    SD_REQUIRE(rays && z, "sd_render_pass: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = R * K;
    const int D = mlp->d_out - 1;
    const int Crgb = 3 * fp.nv_c;
    PointSrc src = {nullptr, rays, z, r_dim, K};

    if (mlp->precision == SD_MLP_F16_TC && tc_supported(scene, mlp, K)) {
        TcRender rr = {};
        rr.cfg = *cfg; rr.depth = depth; rr.dino = dino; rr.rgb_out = rgb_out; rr.weights = weights;
        rr.alphas = alphas; rr.rgb_samps = rgb_samps;
        rr.cmma = tc_render_mode(scene, mlp);
        const size_t need = sd_render_workspace_bytes(scene, mlp, R, K);
        const bool use_ws = rr.cmma == 2 || (rr.cmma == 0 && !rgb_samps && Crgb > 0 && rgb_out);
        if (use_ws && need && (workspace_bytes < need || !workspace)) {
            set_error("sd_render_pass: workspace of %zu B needed, %zu B given", need, workspace_bytes);
            return SD_ERR_WORKSPACE;
        }
        if (rr.cmma == 2) {
            SD_REQUIRE(((uintptr_t)workspace & 15) == 0, "sd_render_pass: workspace must be 16-byte aligned");
            rr.hsum = reinterpret_cast<float *>(workspace);
            rr.wsum = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(workspace) + align256((size_t)R * 128 * 4));
        } else if (use_ws) {
            rr.rgb_samps = reinterpret_cast<float *>(workspace);
        }
        TcOut o = {};
        o.sigma = sigma; o.invalid = invalid; o.invalid_feat = invalid_feat;
        rc = launch_field_tc(fp, src, N, mlp, &rr, o, st, nullptr, scene->feat_proj);
        if (rc || rr.cmma != 2 || !dino) return rc;
        // sum_k w_k (W h_k + b) = W (sum_k w_k h_k) + b sum_k w_k: the feature rows of W_out on the per-ray sums
        return launch_head2(mlp, rr.hsum, rr.wsum, R, dino, st);
    }

    const size_t need = sd_render_workspace_bytes(scene, mlp, R, K);
    if (workspace_bytes < need || !workspace) {
        set_error("sd_render_pass: workspace of %zu B needed, %zu B given", need, workspace_bytes);
        return SD_ERR_WORKSPACE;
    }
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    float *w_sigma = sigma ? sigma : reinterpret_cast<float *>(ws);
    ws += align256((size_t)N * 4);
    float *w_dino = reinterpret_cast<float *>(ws);
    ws += align256((size_t)N * D * 4);
    float *w_rgb = rgb_samps ? rgb_samps : reinterpret_cast<float *>(ws);

    if (mlp->precision == SD_MLP_F16_TC) {
        TcOut o = {};
        o.sigma = w_sigma; o.dino = w_dino; o.rgb = Crgb ? w_rgb : nullptr; o.invalid = invalid; o.invalid_feat = invalid_feat;
        rc = launch_field_tc(fp, src, N, mlp, nullptr, o, st, nullptr, scene->feat_proj);
    } else {
        SimtOut out = {};
        out.sigma = w_sigma; out.dino = w_dino; out.rgb = Crgb ? w_rgb : nullptr; out.invalid = invalid;
        out.invalid_feat = invalid_feat;
        rc = launch_field_simt(MODE_QUERY_, fp, src, N, mlp, out, st);
    }
    if (rc) return rc;
    return sd_composite(z, w_sigma, w_dino, Crgb ? w_rgb : nullptr, R, K, D, Crgb, cfg, weights, alphas, depth,
                        dino, rgb_out, stream);
}

// ---- NeRFRenderer.forward for one scene in one call (nerf.py:451-539) ---------------------------------------------------
static size_t rays_ws_layout(const sd_scene *scene, const sd_mlp *mlp, const sd_sampling *sp, long long R, size_t *o_zc,
                             size_t *o_wc, size_t *o_dc, size_t *o_za, size_t *o_pass) {
    const int Kc = sp->n_coarse, K = sp->n_coarse + sp->n_fine;
    size_t o = 0;
    *o_zc = o; o += align256((size_t)R * Kc * 4);
    *o_wc = o; o += align256((size_t)R * Kc * 4);
    *o_dc = o; o += align256((size_t)R * 4);
    *o_za = o; o += align256((size_t)R * K * 4);
    *o_pass = o;
    const size_t a = sd_render_workspace_bytes(scene, mlp, R, Kc), b = sp->n_fine > 0 ? sd_render_workspace_bytes(scene, mlp, R, K) : 0;
    return o + (a > b ? a : b);
}

extern "C" size_t sd_render_rays_workspace_bytes(const sd_scene *scene, const sd_mlp *mlp, const sd_sampling *samp, long long R) {
    if (!scene || !mlp || !samp || R <= 0 || samp->n_coarse <= 0 || samp->n_fine < 0) return 0;
    size_t a, b, c, d, e;
    return rays_ws_layout(scene, mlp, samp, R, &a, &b, &c, &d, &e);
}

extern "C" int sd_render_rays(const sd_scene *scene, const sd_mlp *mlp, const sd_render_cfg *cfg, const sd_sampling *samp,
                              const float *rays, long long R, int r_dim, const float *u_coarse, const float *lin,
                              const float *u_fine0, const float *u_fine1, const float *n_depth, const sd_render_out *coarse,
                              const sd_render_out *fine, void *workspace, size_t workspace_bytes, void *stream) {
    SD_REQUIRE(scene && mlp && cfg && samp && coarse, "sd_render_rays: null pointer");
    const int Kc = samp->n_coarse, Kf = samp->n_fine, Kfd = samp->n_fine_depth, Kfi = Kf - Kfd;
    SD_REQUIRE(Kc > 0 && Kf >= 0 && Kfd >= 0 && Kfi >= 0, "sd_render_rays: bad sample counts (%d coarse, %d fine, %d of them depth-guided)", Kc, Kf, Kfd);
    SD_REQUIRE(Kf == 0 || fine, "sd_render_rays: n_fine > 0 needs the fine outputs");
    SD_REQUIRE(R >= 0 && r_dim >= 8, "sd_render_rays: bad shape (rays need >= 8 columns)");
    if (R == 0) return SD_OK;
    size_t o_zc, o_wc, o_dc, o_za, o_pass;
    const size_t need = rays_ws_layout(scene, mlp, samp, R, &o_zc, &o_wc, &o_dc, &o_za, &o_pass);
    if (!workspace || workspace_bytes < need) {
        set_error("sd_render_rays: workspace of %zu B needed, %zu B given", need, workspace_bytes);
        return SD_ERR_WORKSPACE;
    }
    SD_REQUIRE(((uintptr_t)workspace & 255) == 0, "sd_render_rays: workspace must be 256-byte aligned");
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    float *z_c = coarse->z_samps ? coarse->z_samps : reinterpret_cast<float *>(ws + o_zc);
    float *w_c = coarse->weights ? coarse->weights : reinterpret_cast<float *>(ws + o_wc);
    float *d_c = coarse->depth ? coarse->depth : reinterpret_cast<float *>(ws + o_dc);
    int rc = sd_sample_coarse(rays, R, r_dim, u_coarse, lin, Kc, cfg->lindisp, z_c, stream);
    if (rc) return rc;
    rc = sd_render_pass(scene, mlp, cfg, rays, R, r_dim, z_c, Kc, d_c, coarse->dino, coarse->rgb, Kf > 0 ? w_c : coarse->weights,
                        coarse->alphas, coarse->invalid, coarse->invalid_feat, coarse->rgb_samps, nullptr, ws + o_pass,
                        workspace_bytes - o_pass, stream);
    if (rc || Kf == 0) return rc;
    float *z_a = fine->z_samps ? fine->z_samps : reinterpret_cast<float *>(ws + o_za);
    rc = launch_fine_merge(rays, R, r_dim, w_c, z_c, d_c, Kc, u_fine0, u_fine1, Kfi, n_depth, Kfd, samp->depth_std, cfg->lindisp, z_a,
                           (cudaStream_t)stream);
    if (rc) return rc;
    return sd_render_pass(scene, mlp, cfg, rays, R, r_dim, z_a, Kc + Kf, fine->depth, fine->dino, fine->rgb, fine->weights, fine->alphas,
                          fine->invalid, fine->invalid_feat, fine->rgb_samps, nullptr, ws + o_pass, workspace_bytes - o_pass, stream);
}
