// fp32 CUDA-core implementation of the field query (BTSNet.sample_features / forward,
// models/bts.py:271-328,476-595) and of the ResnetFC head (resnetfc.py:162-199).
//
// This is the PARITY path (sd_precision SD_MLP_FP32): every stage is evaluated in fp32 with the
// oracle's operation order, the MLP with FFMA.  The throughput path is the fused tcgen05 kernel in
// field_tc.cu; both share the device helpers in common.cuh.
//
// Block = 256 threads, tile = 64 points.
//   stage A  one thread per point: project into the encoder view, frustum mask, bilinear tap, z'
//            one thread per (point, colour view): project + RGB bilinear sample -> global
//   stage B  positional code: one thread per (point, code element)
//   stage C  gather: one warp per point, lanes stride channels with 128-bit loads of the
//            channels-last map (4 texels x C), blended in the oracle's order -> X tile in smem
//   stage D  layer 1: thread j owns hidden unit j for 32 points (32 accumulators), X rows are
//            broadcast 128-bit shared loads, W_in^T rows are coalesced global (L1-resident) loads
//   stage E  layer 2 in chunks of 64 outputs, staged through shared memory so that the output rows
//            leave the SM as fully coalesced segments.
#include "common.cuh"
#include "launch.h"

namespace sd {

constexpr int TP = 64;        // points per tile
constexpr int NTHREADS = 256;
constexpr int HS = 129;       // hidden row stride (floats), odd => conflict-free column access
constexpr int MAX_NVC = 8;

// ---- MLP on one tile held in shared memory -------------------------------------------------------
// X: [TP][XS] fp32 (XS multiple of 4, pad columns zero), result handed to `emit(p, o, value)` in
// chunks of 64 outputs staged through `stage` ([TP][65] floats, may alias X).
template <class Emit>
__device__ __forceinline__ void mlp_tile(const float *__restrict__ X, int XS, float *__restrict__ Hs,
                                         float *__restrict__ stage, const unsigned char *__restrict__ blob,
                                         const MlpLayout &L, int n_valid, Emit emit) {
    const int tid = threadIdx.x;
    const float *w_in_t = reinterpret_cast<const float *>(blob + L.off_w_in_t);
    const float *b_in = reinterpret_cast<const float *>(blob + L.off_b_in);
    const float *w_out_t = reinterpret_cast<const float *>(blob + L.off_w_out_t);
    const float *b_out = reinterpret_cast<const float *>(blob + L.off_b_out);
    {   // layer 1 (d_hidden == 128)
        const int j = tid & 127, g = tid >> 7;
        float acc[32];
        const float b = __ldg(b_in + j);
#pragma unroll
        for (int p = 0; p < 32; ++p) acc[p] = b;
        const float4 *X4 = reinterpret_cast<const float4 *>(X) + (size_t)g * 32 * (XS / 4);
        for (int k4 = 0; k4 < XS / 4; ++k4) {
            const float w0 = __ldg(w_in_t + (size_t)(4 * k4 + 0) * 128 + j);
            const float w1 = __ldg(w_in_t + (size_t)(4 * k4 + 1) * 128 + j);
            const float w2 = __ldg(w_in_t + (size_t)(4 * k4 + 2) * 128 + j);
            const float w3 = __ldg(w_in_t + (size_t)(4 * k4 + 3) * 128 + j);
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float4 x = X4[(size_t)p * (XS / 4) + k4];
                acc[p] = fmaf(x.x, w0, acc[p]);
                acc[p] = fmaf(x.y, w1, acc[p]);
                acc[p] = fmaf(x.z, w2, acc[p]);
                acc[p] = fmaf(x.w, w3, acc[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < 32; ++p) Hs[(g * 32 + p) * HS + j] = fmaxf(acc[p], 0.0f);
    }
    __syncthreads();
    // layer 2: thread (p = tid/4, q = tid%4) computes outputs ob + q + 4*i, i < 16
    const int p = tid >> 2, q = tid & 3;
    for (int ob = 0; ob < L.d_out; ob += 64) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int o = ob + q + 4 * i;
            acc[i] = o < L.d_out_pad ? __ldg(b_out + o) : 0.0f;
        }
        for (int k = 0; k < 128; ++k) {
            const float h = Hs[p * HS + k];
            const float *w = w_out_t + (size_t)k * L.d_out_pad + ob + q;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (ob + q + 4 * i < L.d_out_pad) acc[i] = fmaf(h, __ldg(w + 4 * i), acc[i]);
        }
        __syncthreads();  // previous chunk fully emitted / X no longer needed
#pragma unroll
        for (int i = 0; i < 16; ++i) stage[p * 65 + q + 4 * i] = acc[i];
        __syncthreads();
        const int nout = min(64, L.d_out - ob);
        for (int idx = tid; idx < n_valid * nout; idx += NTHREADS) {
            const int pp = idx / nout, oo = idx - pp * nout;
            emit(pp, ob + oo, stage[pp * 65 + oo]);
        }
    }
}

enum { MODE_FEATURES = MODE_FEATURES_, MODE_QUERY = MODE_QUERY_ };

template <int MODE>
__global__ void __launch_bounds__(NTHREADS) field_simt_kernel(FieldParams fp, PointSrc src, long long N,
                                                              const unsigned char *__restrict__ blob, MlpLayout L,
                                                              SimtOut out, int XS) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *X = reinterpret_cast<float *>(smem_raw);                  // [TP][XS]
    float *Hs = X + (size_t)TP * XS;                                 // [TP][HS]
    float *cam = Hs + (size_t)TP * HS;                               // 21 floats per camera, 1 + nv_c cameras
    float *vq = cam + 21 * (1 + MAX_NVC);                            // [TP][3]  (x, y, z')
    float *wq = vq + TP * 3;                                         // [TP][4]  bilinear weights
    int *tq = reinterpret_cast<int *>(wq + TP * 4);                  // [TP][2]  x0, y0
    unsigned char *fq = reinterpret_cast<unsigned char *>(tq + TP * 2);  // [TP] bit0 invalid, bit1 in_x1, bit2 in_y1

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long base = (long long)blockIdx.x * TP;
    const int n_valid = (int)min((long long)TP, N - base);

    for (int i = tid; i < 21 * (1 + fp.nv_c); i += NTHREADS) {
        const int c = i / 21, e = i - 21 * c;
        const float *K = c == 0 ? fp.K_f : fp.K_c + 9 * (c - 1);
        const float *W = c == 0 ? fp.w2c_f : fp.w2c_c + 16 * (c - 1);
        cam[i] = e < 9 ? __ldg(K + e) : __ldg(W + (e - 9));
    }
    __syncthreads();

    // ---- stage A -----------------------------------------------------------------------------
    if (tid < TP) {
        const int p = tid;
        unsigned char flags = 1;
        float x = 0.f, y = 0.f, zp = 0.f;
        Tap t = {};
        if (p < n_valid) {
            float px, py, pz, z;
            bool inv;
            load_point(src, base + p, px, py, pz);
            project_point(cam, cam + 9, px, py, pz, x, y, z, inv);
            x = clamp_keep_nan(x, -2.0f, 2.0f);
            y = clamp_keep_nan(y, -2.0f, 2.0f);
            zp = znorm(z, fp.enc);
            t = bilinear_tap(x, y, fp.Hf, fp.Wf);
            flags = (inv ? 1 : 0) | (t.in_x1 ? 2 : 0) | (t.in_y1 ? 4 : 0);
            if (out.invalid_feat) out.invalid_feat[base + p] = inv ? 1 : 0;
        }
        vq[p * 3 + 0] = x; vq[p * 3 + 1] = y; vq[p * 3 + 2] = zp;
        wq[p * 4 + 0] = t.wnw; wq[p * 4 + 1] = t.wne; wq[p * 4 + 2] = t.wsw; wq[p * 4 + 3] = t.wse;
        tq[p * 2 + 0] = t.x0; tq[p * 2 + 1] = t.y0;
        fq[p] = flags;
    }
    __syncthreads();
    if (MODE == MODE_QUERY && fp.nv_c > 0 && (out.rgb || out.invalid)) {
        for (int idx = tid; idx < n_valid * fp.nv_c; idx += NTHREADS) {
            const int p = idx / fp.nv_c, v = idx - p * fp.nv_c;
            float px, py, pz, x, y, z;
            bool inv;
            load_point(src, base + p, px, py, pz);
            const float *c = cam + 21 * (1 + v);
            project_point(c, c + 9, px, py, pz, x, y, z, inv);
            if (out.rgb) {
                float c3[3];
                sample_color(fp.rgb + (size_t)v * 3 * fp.Hc * fp.Wc, fp.Hc, fp.Wc, x, y, c3);
                float *o = out.rgb + (size_t)(base + p) * 3 * fp.nv_c + 3 * v;
                o[0] = c3[0]; o[1] = c3[1]; o[2] = c3[2];
            }
            if (out.invalid)  // bts.py:566-569 (nv_f == 1: all(invalid_features) == invalid_features)
                out.invalid[(size_t)(base + p) * fp.nv_c + v] = (inv || (fq[p] & 1)) ? 1.0f : 0.0f;
        }
    }
    // ---- stage B: positional code into X[:, C : C+code] (+ zero padding up to XS) --------------
    {
        const int ncode = XS - fp.C;
        for (int idx = tid; idx < TP * ncode; idx += NTHREADS) {
            const int p = idx / ncode, i = idx - p * ncode;
            float v = 0.0f;
            if (p < n_valid && i < fp.code_dim)
                v = code_element(i, vq[p * 3 + 0], vq[p * 3 + 1], vq[p * 3 + 2], fp.enc);
            X[(size_t)p * XS + fp.C + i] = v;
        }
    }
    // ---- stage C: gather -----------------------------------------------------------------------
    for (int p = warp; p < TP; p += NTHREADS / 32) {
        float4 *xrow = reinterpret_cast<float4 *>(X + (size_t)p * XS);
        if (p >= n_valid) {
            for (int c = lane * 4; c < fp.C; c += 128) xrow[c / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const unsigned char fl = fq[p];
        if (fp.learn_empty && (fl & 1)) {  // bts.py:311-319
            for (int c = lane * 4; c < fp.C; c += 128)
                xrow[c / 4] = __ldg(reinterpret_cast<const float4 *>(fp.empty_feature + c));
            continue;
        }
        Tap t;
        t.x0 = tq[p * 2]; t.y0 = tq[p * 2 + 1];
        t.wnw = wq[p * 4]; t.wne = wq[p * 4 + 1]; t.wsw = wq[p * 4 + 2]; t.wse = wq[p * 4 + 3];
        t.in_x1 = fl & 2; t.in_y1 = fl & 4;
        const size_t o_nw = ((size_t)t.y0 * fp.Wf + t.x0) * fp.C;
        const size_t o_ne = o_nw + (t.in_x1 ? fp.C : 0);
        const size_t o_sw = o_nw + (t.in_y1 ? (size_t)fp.Wf * fp.C : 0);
        const size_t o_se = o_sw + (t.in_x1 ? fp.C : 0);
        for (int c = lane * 4; c < fp.C; c += 128) {
            float nw[4], ne[4], sw[4], se[4];
            if (fp.feat_f16) {
                const __half *f = reinterpret_cast<const __half *>(fp.feat);
                const uint2 a = __ldg(reinterpret_cast<const uint2 *>(f + o_nw + c));
                const uint2 b = __ldg(reinterpret_cast<const uint2 *>(f + o_ne + c));
                const uint2 cc = __ldg(reinterpret_cast<const uint2 *>(f + o_sw + c));
                const uint2 d = __ldg(reinterpret_cast<const uint2 *>(f + o_se + c));
                float2 t0 = half2_bits_to_float2(a.x), t1 = half2_bits_to_float2(a.y);
                nw[0] = t0.x; nw[1] = t0.y; nw[2] = t1.x; nw[3] = t1.y;
                t0 = half2_bits_to_float2(b.x); t1 = half2_bits_to_float2(b.y);
                ne[0] = t0.x; ne[1] = t0.y; ne[2] = t1.x; ne[3] = t1.y;
                t0 = half2_bits_to_float2(cc.x); t1 = half2_bits_to_float2(cc.y);
                sw[0] = t0.x; sw[1] = t0.y; sw[2] = t1.x; sw[3] = t1.y;
                t0 = half2_bits_to_float2(d.x); t1 = half2_bits_to_float2(d.y);
                se[0] = t0.x; se[1] = t0.y; se[2] = t1.x; se[3] = t1.y;
            } else {
                const float *f = reinterpret_cast<const float *>(fp.feat);
                const float4 a = __ldg(reinterpret_cast<const float4 *>(f + o_nw + c));
                const float4 b = __ldg(reinterpret_cast<const float4 *>(f + o_ne + c));
                const float4 cc = __ldg(reinterpret_cast<const float4 *>(f + o_sw + c));
                const float4 d = __ldg(reinterpret_cast<const float4 *>(f + o_se + c));
                nw[0] = a.x; nw[1] = a.y; nw[2] = a.z; nw[3] = a.w;
                ne[0] = b.x; ne[1] = b.y; ne[2] = b.z; ne[3] = b.w;
                sw[0] = cc.x; sw[1] = cc.y; sw[2] = cc.z; sw[3] = cc.w;
                se[0] = d.x; se[1] = d.y; se[2] = d.z; se[3] = d.w;
            }
            float4 r;
            r.x = blend4(nw[0], ne[0], sw[0], se[0], t);
            r.y = blend4(nw[1], ne[1], sw[1], se[1], t);
            r.z = blend4(nw[2], ne[2], sw[2], se[2], t);
            r.w = blend4(nw[3], ne[3], sw[3], se[3], t);
            xrow[c / 4] = r;
        }
    }
    __syncthreads();

    if (MODE == MODE_FEATURES) {
        const int d = fp.C + fp.code_dim;
        for (int idx = tid; idx < n_valid * d; idx += NTHREADS) {
            const int p = idx / d, c = idx - p * d;
            out.feat[(size_t)(base + p) * d + c] = X[(size_t)p * XS + c];
        }
        return;
    }
    const int D = L.d_out - 1;
    mlp_tile(X, XS, Hs, X, blob, L, n_valid, [&](int p, int o, float v) {
        if (o == 0) { if (out.sigma) out.sigma[base + p] = softplus(v); }
        else if (out.dino) out.dino[(size_t)(base + p) * D + (o - 1)] = v;
    });
}

// plain ResnetFC.forward on rows given in global memory; NORMALIZE adds F.normalize(dim=-1) for
// MlpDimReduction.transform_expand (second pass over the row just written by this block).
template <bool NORMALIZE>
__global__ void __launch_bounds__(NTHREADS) mlp_simt_kernel(const float *__restrict__ x, long long N,
                                                            const unsigned char *__restrict__ blob, MlpLayout L,
                                                            float *__restrict__ outp, int XS) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *X = reinterpret_cast<float *>(smem_raw);
    const int stage_floats = TP * 65;
    const int xfloats = TP * XS > stage_floats ? TP * XS : stage_floats;
    float *Hs = X + xfloats;
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * TP;
    const int n_valid = (int)min((long long)TP, N - base);
    for (int idx = tid; idx < TP * XS; idx += NTHREADS) {
        const int p = idx / XS, c = idx - p * XS;
        X[idx] = (p < n_valid && c < L.d_in) ? __ldg(x + (size_t)(base + p) * L.d_in + c) : 0.0f;
    }
    __syncthreads();
    mlp_tile(X, XS, Hs, X, blob, L, n_valid,
             [&](int p, int o, float v) { outp[(size_t)(base + p) * L.d_out + o] = v; });
    if (NORMALIZE) {
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        for (int p = warp; p < n_valid; p += NTHREADS / 32) {
            float *row = outp + (size_t)(base + p) * L.d_out;
            float ss = 0.0f;
            for (int c = lane; c < L.d_out; c += 32) { const float v = row[c]; ss = fmaf(v, v, ss); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float den = fmaxf(sqrtf(ss), 1e-12f);
            for (int c = lane; c < L.d_out; c += 32) row[c] = __fdiv_rn(row[c], den);
        }
    }
}

// (camera matrices are read from device memory like everywhere else: no host round trip, graph-capturable)
__global__ void __launch_bounds__(256) project_points_kernel(const float *__restrict__ Kd, const float *__restrict__ Wd,
                                                             const float *__restrict__ xyz, long long N,
                                                             float *__restrict__ xy, float *__restrict__ z,
                                                             unsigned char *__restrict__ invalid) {
    __shared__ float cam[21];
    if (threadIdx.x < 21) cam[threadIdx.x] = threadIdx.x < 9 ? __ldg(Kd + threadIdx.x) : __ldg(Wd + (threadIdx.x - 9));
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float x, y, zz;
    bool inv;
    project_point(cam, cam + 9, __ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), x, y, zz, inv);
    if (xy) { xy[2 * i] = x; xy[2 * i + 1] = y; }
    if (z) z[i] = zz;
    if (invalid) invalid[i] = inv ? 1 : 0;
}

// BTSNet.sample_colors (bts.py:330-358): one thread per (point, colour view)
__global__ void __launch_bounds__(256) sample_colors_kernel(FieldParams fp, const float *__restrict__ xyz, long long N,
                                                            float *__restrict__ rgb,
                                                            unsigned char *__restrict__ invalid) {
    __shared__ float cam[21 * MAX_NVC];
    for (int i = threadIdx.x; i < 21 * fp.nv_c; i += blockDim.x) {
        const int c = i / 21, e = i - 21 * c;
        cam[i] = e < 9 ? __ldg(fp.K_c + 9 * c + e) : __ldg(fp.w2c_c + 16 * c + (e - 9));
    }
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * fp.nv_c) return;
    const long long p = idx / fp.nv_c;
    const int v = (int)(idx - p * fp.nv_c);
    float x, y, z;
    bool inv;
    const float *c = cam + 21 * v;
    project_point(c, c + 9, __ldg(xyz + 3 * p), __ldg(xyz + 3 * p + 1), __ldg(xyz + 3 * p + 2), x, y, z, inv);
    if (rgb) {
        float c3[3];
        sample_color(fp.rgb + (size_t)v * 3 * fp.Hc * fp.Wc, fp.Hc, fp.Wc, x, y, c3);
        float *o = rgb + (size_t)p * 3 * fp.nv_c + 3 * v;
        o[0] = c3[0]; o[1] = c3[1]; o[2] = c3[2];
    }
    if (invalid) invalid[idx] = inv ? 1 : 0;
}

int make_field_params(const sd_scene *s, FieldParams *o) {
    SD_REQUIRE(s, "scene is NULL");
    SD_REQUIRE(s->feat && s->K_f && s->w2c_f, "scene: feature map / cameras missing (call encode first)");
    SD_REQUIRE(s->nv_f == 1, "scene: the default head supports exactly one encoder view (nv_f=%d)", s->nv_f);
    SD_REQUIRE(s->C > 0 && s->C % 8 == 0, "scene: C must be a positive multiple of 8 (got %d)", s->C);
    SD_REQUIRE(s->Hf > 0 && s->Wf > 0, "scene: bad feature map size");
    SD_REQUIRE(s->feat_dtype == SD_F32 || s->feat_dtype == SD_F16, "scene: bad feat_dtype");
    SD_REQUIRE(s->nv_c >= 0 && s->nv_c <= MAX_NVC, "scene: at most %d colour views (got %d)", MAX_NVC, s->nv_c);
    SD_REQUIRE(s->nv_c == 0 || (s->rgb && s->K_c && s->w2c_c && s->Hc > 0 && s->Wc > 0), "scene: colour views incomplete");
    SD_REQUIRE(s->num_freqs >= 0 && s->num_freqs <= 16, "scene: bad num_freqs");
    SD_REQUIRE(!s->learn_empty || s->empty_feature, "scene: learn_empty without empty_feature");
    o->feat = s->feat; o->feat_f16 = s->feat_dtype == SD_F16;
    o->C = s->C; o->Hf = s->Hf; o->Wf = s->Wf;
    o->K_f = s->K_f; o->w2c_f = s->w2c_f;
    o->rgb = s->rgb; o->nv_c = s->nv_c; o->Hc = s->Hc; o->Wc = s->Wc; o->K_c = s->K_c; o->w2c_c = s->w2c_c;
    o->enc.inv_z = s->inv_z;
    o->enc.inv_dmax = (float)(1.0 / (double)s->d_max);
    o->enc.denom = s->inv_z ? (float)(1.0 / (double)s->d_min - 1.0 / (double)s->d_max)
                            : (float)((double)s->d_max - (double)s->d_min);
    o->enc.d_min = s->d_min;
    o->enc.num_freqs = s->num_freqs; o->enc.freq_factor = s->freq_factor; o->enc.include_input = s->include_input;
    o->code_dim = (s->include_input ? 3 : 0) + 6 * s->num_freqs;
    o->learn_empty = s->learn_empty; o->empty_feature = s->empty_feature;
    return SD_OK;
}

static size_t simt_smem_bytes(int XS) {
    return sizeof(float) * ((size_t)TP * XS + (size_t)TP * HS + 21 * (1 + MAX_NVC) + TP * 3 + TP * 4) +
           sizeof(int) * TP * 2 + TP + 64;
}

int launch_field_simt(int mode, const FieldParams &fp, const PointSrc &src, long long N, const sd_mlp *mlp,
                      const SimtOut &out, cudaStream_t st) {
    if (N == 0) return SD_OK;
    MlpLayout L = {};
    const unsigned char *blob = nullptr;
    const int d_feat = fp.C + fp.code_dim;
    if (mode == MODE_QUERY) {
        SD_REQUIRE(mlp && mlp->packed, "mlp: packed weights missing (sd_mlp_pack)");
        SD_REQUIRE(mlp->d_hidden == 128, "mlp: d_hidden must be 128 (got %d)", mlp->d_hidden);
        SD_REQUIRE(mlp->d_in == d_feat, "mlp: d_in=%d but the field produces %d features", mlp->d_in, d_feat);
        L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
        blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    }
    const int XS = (d_feat + 3) / 4 * 4;
    SD_REQUIRE(XS * TP >= TP * 65, "field: feature width too small");
    const size_t smem = simt_smem_bytes(XS);
    SD_REQUIRE(smem <= 227 * 1024, "field: C=%d needs %zu B of shared memory", fp.C, smem);
    const unsigned grid = (unsigned)((N + TP - 1) / TP);
    if (mode == MODE_QUERY) {
        SD_CUDA_OK(cudaFuncSetAttribute(field_simt_kernel<MODE_QUERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        field_simt_kernel<MODE_QUERY><<<grid, NTHREADS, smem, st>>>(fp, src, N, blob, L, out, XS);
    } else {
        SD_CUDA_OK(cudaFuncSetAttribute(field_simt_kernel<MODE_FEATURES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        field_simt_kernel<MODE_FEATURES><<<grid, NTHREADS, smem, st>>>(fp, src, N, blob, L, out, XS);
    }
    SD_LAUNCH_OK("field_simt_kernel");
    return SD_OK;
}

int launch_mlp_simt(const sd_mlp *mlp, const float *x, long long N, float *out, bool normalize, cudaStream_t st) {
    SD_REQUIRE(mlp && mlp->packed, "mlp_forward: null pointer");
    SD_REQUIRE(mlp->d_hidden == 128, "mlp: d_hidden must be 128 (got %d)", mlp->d_hidden);
    SD_REQUIRE(mlp->d_in > 0 && mlp->d_in <= 512 && mlp->d_out > 0, "mlp: unsupported dims");
    if (N == 0) return SD_OK;
    SD_REQUIRE(x && out, "mlp_forward: null pointer");
    const MlpLayout L = mlp_layout(mlp->d_in, mlp->d_hidden, mlp->d_out);
    const int XS = (mlp->d_in + 3) / 4 * 4;
    const size_t xfloats = (size_t)TP * XS > (size_t)TP * 65 ? (size_t)TP * XS : (size_t)TP * 65;
    const size_t smem = sizeof(float) * (xfloats + (size_t)TP * HS);
    const unsigned grid = (unsigned)((N + TP - 1) / TP);
    const unsigned char *blob = reinterpret_cast<const unsigned char *>(mlp->packed);
    if (normalize) {
        SD_CUDA_OK(cudaFuncSetAttribute(mlp_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_simt_kernel<true><<<grid, NTHREADS, smem, st>>>(x, N, blob, L, out, XS);
    } else {
        SD_CUDA_OK(cudaFuncSetAttribute(mlp_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mlp_simt_kernel<false><<<grid, NTHREADS, smem, st>>>(x, N, blob, L, out, XS);
    }
    SD_LAUNCH_OK("mlp_simt_kernel");
    return SD_OK;
}

}  // namespace sd

using namespace sd;

extern "C" int sd_project_points(const float *K, const float *w2c, const float *xyz, long long N, float *xy,
                                 float *z, unsigned char *invalid, void *stream) {
    SD_REQUIRE(N >= 0, "sd_project_points: bad N");
    if (N == 0) return SD_OK;
    SD_REQUIRE(K && w2c && xyz, "sd_project_points: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    project_points_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(K, w2c, xyz, N, xy, z, invalid);
    SD_LAUNCH_OK("project_points_kernel");
    return SD_OK;
}

extern "C" int sd_sample_features(const sd_scene *scene, const float *xyz, long long N, float *feat,
                                  unsigned char *invalid, void *stream) {
    FieldParams fp;
    int rc = make_field_params(scene, &fp);
    if (rc) return rc;
    SD_REQUIRE(N >= 0, "sd_sample_features: bad N");
    if (N == 0) return SD_OK;
    SD_REQUIRE(xyz && feat, "sd_sample_features: null pointer");
    PointSrc src = {xyz, nullptr, nullptr, 0, 1};
    SimtOut out = {};
    out.feat = feat; out.invalid_feat = invalid;
    return launch_field_simt(MODE_FEATURES, fp, src, N, nullptr, out, (cudaStream_t)stream);
}

extern "C" int sd_sample_colors(const sd_scene *scene, const float *xyz, long long N, float *rgb,
                                unsigned char *invalid, void *stream) {
    FieldParams fp;
    int rc = make_field_params(scene, &fp);
    if (rc) return rc;
    SD_REQUIRE(N >= 0, "sd_sample_colors: bad N");
    SD_REQUIRE(fp.nv_c > 0, "sd_sample_colors: the scene has no colour views");
    if (N == 0) return SD_OK;
    SD_REQUIRE(xyz, "sd_sample_colors: null pointer");
    const long long n = N * fp.nv_c;
    sample_colors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fp, xyz, N, rgb, invalid);
    SD_LAUNCH_OK("sample_colors_kernel");
    return SD_OK;
}

extern "C" int sd_expand_dim(const sd_mlp *mlp, const float *f, long long N, float *out, void *stream) {
    // reduced precision: fp16 operands on the tensor cores (expand_tc.cu) for the shipped 64 -> 128 -> 768 shape
    if (mlp && mlp->precision == SD_MLP_F16_TC && expand_tc_supported(mlp))
        return launch_expand_tc(mlp, f, N, out, (cudaStream_t)stream);
    return launch_mlp_simt(mlp, f, N, out, true, (cudaStream_t)stream);
}
