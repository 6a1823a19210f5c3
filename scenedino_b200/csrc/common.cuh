// Shared device helpers for the scenedino_b200 kernels (sm_100a).
//
// Rounding contract: everything the parity tests compare bit-for-bit (projection, frustum mask,
// sample depths, bilinear taps) is written with explicitly rounded intrinsics (__fmul_rn,
// __fmaf_rn, ...) in the SAME order as oracle/sd_oracle.c, so nvcc can neither fuse nor reorder.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scenedino_b200.h"

#define SD_EPS 1e-3f  // common/cameras/pinhole.py:3
#define SD_BIN 7      // texel bins of 7 x 7: the 2 x 2 footprints that start in a bin cover an 8 x 8 box

namespace sd {

// ---- host-side error plumbing ------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);
// One-time setup of a launcher PER DEVICE (kernel attributes such as the dynamic shared-memory limit belong to the device
// the function was loaded on): SM count of the current device in *sm_count; *first says that `seen` had no entry for it.
constexpr int SD_MAX_DEVICES = 64;
struct DeviceOnce { int sm_count[SD_MAX_DEVICES]; };
int device_once(DeviceOnce &seen, int *sm_count, bool *first);
int debug_mask();                       // SD_TC_DEBUG, read once per process (timing experiments; 0 in production)
void profile_before(cudaStream_t st);   // sd_profile_next_kernel: events around the dominant (field) kernel
void profile_after(cudaStream_t st);

#define SD_CUDA_OK(expr)                                                    \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return sd::cuda_fail(_e, #expr);             \
    } while (0)

#define SD_LAUNCH_OK(name)                                                  \
    do {                                                                    \
        sd::count_launch();                                                 \
        cudaError_t _e = cudaGetLastError();                                \
        if (_e != cudaSuccess) return sd::cuda_fail(_e, name);              \
    } while (0)

#define SD_REQUIRE(cond, ...)                                               \
    do {                                                                    \
        if (!(cond)) { sd::set_error(__VA_ARGS__); return SD_ERR_INVALID; } \
    } while (0)

// ---- camera (by value into kernels: lives in constant bank / registers) -------------------------
struct Camera {
    float K[9];
    float w2c[12];  // first three rows of the 4x4
};

struct EncodeParams {
    float inv_dmax;  // (float)(1/d_max)
    float denom;     // (float)(1/d_min - 1/d_max)   [inv_z]   or (float)(d_max - d_min)
    float d_min;
    int inv_z;
    int num_freqs;
    float freq_factor;
    int include_input;
};

// Per-launch description of the field (one batch element).  Camera matrices are read from global
// memory once per block into shared memory by the kernels that need them.
struct FieldParams {
    const void *feat;
    int feat_f16;
    int C, Hf, Wf;
    const float *K_f, *w2c_f;
    const float *rgb;
    int nv_c, Hc, Wc;
    const float *K_c, *w2c_c;
    EncodeParams enc;
    int code_dim;
    int learn_empty;
    const float *empty_feature;
};

int make_field_params(const sd_scene *s, FieldParams *out);

// ---- packed MLP blob layout (written by sd_mlp_pack, see pack.cu) -------------------------------
struct MlpLayout {
    int d_in, d_hidden, d_out;
    int d_in_pad;    // multiple of 16 (tcgen05 K) and of 4 (float4)
    int d_out_pad;   // multiple of 16
    size_t off_w_in_t;   // fp32 [d_in_pad][d_hidden]      (zero rows beyond d_in)
    size_t off_b_in;     // fp32 [d_hidden]
    size_t off_w_out_t;  // fp32 [d_hidden][d_out_pad]     (zero cols beyond d_out)
    size_t off_b_out;    // fp32 [d_out_pad]
    size_t off_w_in_h;   // fp16 UMMA K-major SW128 image of W_in:  [d_in_pad/64 blocks][d_hidden rows][64]
    size_t off_w_out_h;  // fp16 UMMA K-major SW128 image of W_out[1:]: [d_hidden/64 blocks][d_out_pad rows][64]
    size_t off_w_sigma;  // fp32 [d_hidden]  = W_out[0,:]  (density row, evaluated in fp32)
    size_t off_w_sig_h;  // d_hidden == 128: fp16 K-major SW128 image [2 K blocks][16 rows][64] whose row 0 is the density row
                         // W_out[0,:] (layer 2 of the hidden-composite render mode: only sigma is needed per sample)
    size_t off_w_feat_blk;  // d_hidden == 128: W_out[1:] in blocks of 128 output rows, each a complete B operand
                         // [2 K blocks][128 rows][64] (zero rows behind d_out - 1): head2 kernel (expand_tc.cu)
    size_t off_x_w2;     // expand heads (64 -> 128 -> multiple of 128): fp16 K-major SW128 images of W_out in blocks of 128
                         // outputs, [d_out/128][2 K blocks][128 rows][64] (expand_tc.cu); 0 = absent
    size_t total;
};
MlpLayout mlp_layout(int d_in, int d_hidden, int d_out);

// byte offset of element (row, k) inside a K-major SWIZZLE_128B UMMA operand image with `rows` rows:
// [k/64 blocks][rows][128 B]; the 16-byte chunk index is XORed with (row & 7)  (cute Swizzle<3,4,3>).
// Images must start on a 1024-byte boundary of shared memory.
__host__ __device__ __forceinline__ size_t umma_sw128_offset(int row, int k, int rows) {
    const int kb = k >> 6, kk = k & 63;
    const int chunk = (kk >> 3) ^ (row & 7);
    return ((size_t)kb * rows + row) * 128 + (size_t)chunk * 16 + (size_t)(kk & 7) * 2;
}

// ---- exact device arithmetic ---------------------------------------------------------------------
__device__ __forceinline__ float clamp_keep_nan(float v, float lo, float hi) {
    if (v != v) return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

// pinhole.py:40-112 for one camera.  x,y unclamped.
__device__ __forceinline__ void project_point(const float *__restrict__ K, const float *__restrict__ w,
                                              float px, float py, float pz, float &x, float &y,
                                              float &z, bool &invalid) {
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(w[r * 4 + 0], px);
        acc = __fmaf_rn(w[r * 4 + 1], py, acc);
        acc = __fmaf_rn(w[r * 4 + 2], pz, acc);
        c[r] = __fadd_rn(acc, w[r * 4 + 3]);
    }
    float q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(K[r * 3 + 0], c[0]);
        acc = __fmaf_rn(K[r * 3 + 1], c[1], acc);
        q[r] = __fmaf_rn(K[r * 3 + 2], c[2], acc);
    }
    float zc = q[2] > SD_EPS ? q[2] : SD_EPS;
    if (q[2] != q[2]) zc = q[2];
    x = __fdiv_rn(q[0], zc);
    y = __fdiv_rn(q[1], zc);
    z = q[2];
    invalid = (q[2] <= SD_EPS) | (x < -1.0f) | (x > 1.0f) | (y < -1.0f) | (y > 1.0f);
}

// nerf.py:252: o + z*d with separately rounded multiply and add
__device__ __forceinline__ float ray_point(float o, float d, float z) {
    return __fadd_rn(o, __fmul_rn(z, d));
}

// positional_encoding.py:13-21 (encoding_mode "z")
__device__ __forceinline__ float znorm(float z, const EncodeParams &e) {
    float zn;
    if (e.inv_z) {
        float zc = z > SD_EPS ? z : SD_EPS;
        if (z != z) zc = z;
        zn = __fdiv_rn(__fsub_rn(__fdiv_rn(1.0f, zc), e.inv_dmax), e.denom);
    } else {
        zn = __fdiv_rn(__fsub_rn(z, e.d_min), e.denom);
    }
    return __fsub_rn(__fmul_rn(2.0f, zn), 1.0f);
}

// element `i` of the positional code of v = (x, y, z')  (positional_encoding.py:68-80)
__device__ __forceinline__ float code_element(int i, float vx, float vy, float vz, const EncodeParams &e) {
    if (e.include_input) {
        if (i < 3) return i == 0 ? vx : (i == 1 ? vy : vz);
        i -= 3;
    }
    const int k = i / 6, rem = i - 6 * k;
    const int ph = rem / 3, d = rem - 3 * ph;
    const float f = __fmul_rn(e.freq_factor, (float)(1 << k));
    const float v = d == 0 ? vx : (d == 1 ? vy : vz);
    const float arg = __fmaf_rn(v, f, ph ? 1.57079632679489661923f : 0.0f);
    return sinf(arg);
}

// F.grid_sample(bilinear, border, align_corners=False): ATen formulation, see oracle sdo_bilinear_tap
struct Tap {
    int x0, y0;
    float wnw, wne, wsw, wse;
    bool in_x1, in_y1;
};

__device__ __forceinline__ Tap bilinear_tap(float x, float y, int H, int W) {
    Tap t;
    float ix = __fmaf_rn(__fadd_rn(x, 1.0f), (float)W * 0.5f, -0.5f);
    float iy = __fmaf_rn(__fadd_rn(y, 1.0f), (float)H * 0.5f, -0.5f);
    ix = fminf((float)(W - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.0f));
    const float fx = floorf(ix), fy = floorf(iy);
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    const float w = __fsub_rn(ix, fx), e = __fsub_rn(1.0f, w);
    const float n = __fsub_rn(iy, fy), s = __fsub_rn(1.0f, n);
    t.wnw = __fmul_rn(s, e);
    t.wne = __fmul_rn(s, w);
    t.wsw = __fmul_rn(n, e);
    t.wse = __fmul_rn(n, w);
    t.in_x1 = t.x0 + 1 <= W - 1;
    t.in_y1 = t.y0 + 1 <= H - 1;
    return t;
}

// Keeps the 2x2 footprint of a tap inside the map, so that a gather can use fixed +1 texel / +1 row offsets: at
// the last column / row the out-of-range taps have weight zero, so shifting the base by one and moving the
// weights over is exact.  (NaN coordinates give an arbitrary tap: the base is clamped into the map whatever
// happens.)  Used by the tensor-core kernels AND by the texel binning, which must agree on the base texel.
__device__ __forceinline__ void clamp_footprint(Tap &t, int H, int W) {
    if (!t.in_x1) { t.x0 -= 1; t.wne = t.wnw; t.wse = t.wsw; t.wnw = 0.0f; t.wsw = 0.0f; }
    if (!t.in_y1) { t.y0 -= 1; t.wsw = t.wnw; t.wse = t.wne; t.wnw = 0.0f; t.wne = 0.0f; }
    t.x0 = min(max(t.x0, 0), W - 2);
    t.y0 = min(max(t.y0, 0), H - 2);
}

// blend in the oracle's order: nw, ne, sw, se (out-of-range corners are skipped)
__device__ __forceinline__ float blend4(float nw, float ne, float sw, float se, const Tap &t) {
    float acc = __fmul_rn(nw, t.wnw);
    if (t.in_x1) acc = __fmaf_rn(ne, t.wne, acc);
    if (t.in_y1) acc = __fmaf_rn(sw, t.wsw, acc);
    if (t.in_x1 && t.in_y1) acc = __fmaf_rn(se, t.wse, acc);
    return acc;
}

// F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// nerf.py:135-138 / 205-208
__device__ __forceinline__ float depth_from_t(float near, float far, float t, int lindisp) {
    if (!lindisp) {
        const float a = __fmul_rn(near, __fsub_rn(1.0f, t));
        const float b = __fmul_rn(far, t);
        return __fadd_rn(a, b);
    }
    const float a = __fmul_rn(__fdiv_rn(1.0f, near), __fsub_rn(1.0f, t));
    const float b = __fmul_rn(__fdiv_rn(1.0f, far), t);
    return __fdiv_rn(1.0f, __fadd_rn(a, b));
}

// sample RGB of one colour view (bts.py:330-358): planar [3,Hc,Wc] fp32
__device__ __forceinline__ void sample_color(const float *__restrict__ img, int Hc, int Wc, float x,
                                             float y, float *rgb3) {
    const Tap t = bilinear_tap(clamp_keep_nan(x, -2.0f, 2.0f), clamp_keep_nan(y, -2.0f, 2.0f), Hc, Wc);
    const size_t plane = (size_t)Hc * Wc;
    const size_t o = (size_t)t.y0 * Wc + t.x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float *p = img + c * plane + o;
        const float nw = __ldg(p);
        const float ne = t.in_x1 ? __ldg(p + 1) : 0.0f;
        const float sw = t.in_y1 ? __ldg(p + Wc) : 0.0f;
        const float se = (t.in_x1 && t.in_y1) ? __ldg(p + Wc + 1) : 0.0f;
        rgb3[c] = blend4(nw, ne, sw, se, t);
    }
}

__device__ __forceinline__ float2 half2_bits_to_float2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2 *>(&w)); }

}  // namespace sd
