// Small stand-alone kernels: PositionalEncoding.forward on its own (inside the field kernels the code is fused) and the
// backward of the bilinear feature gather (training path).
#include "common.cuh"

namespace sd {

// common/positional_encoding.py:68-80 for x [N, d_in]: out row = [x | k = 0..F-1: sin(f_k x_0..), sin(f_k x_0.. + pi/2)],
// argument formed as fma(x, f, phase) (torch.addcmul on contiguous fp32 CPU tensors), f_k = freq_factor * 2^k.
__global__ void __launch_bounds__(256) positional_encoding_kernel(const float *__restrict__ x, long long N, int d_in, int num_freqs,
                                                                  float freq_factor, int include_input, float *__restrict__ out) {
    const int d_out = (include_input ? d_in : 0) + 2 * num_freqs * d_in;
    const long long total = N * d_out;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d_out;
        int c = (int)(i - r * d_out);
        float v;
        if (include_input && c < d_in) {
            v = __ldg(x + r * d_in + c);
        } else {
            if (include_input) c -= d_in;
            const int blk = c / d_in, d = c - blk * d_in;      // blk = 2 k + phase
            const float f = __fmul_rn(freq_factor, (float)(1 << (blk >> 1)));
            v = sinf(__fmaf_rn(__ldg(x + r * d_in + d), f, (blk & 1) ? 1.57079632679489661923f : 0.0f));
        }
        out[i] = v;
    }
}

// ---- backward of the bilinear feature gather (what autograd derives from bts.py:299-319) -------------------
// g_feat [N, d_in] (d_in = C + code; only the first C columns carry a gradient to the map) is scattered into the
// channels-last map gradient with the forward's own tap (same projection, clamp and weights): one warp per point, a lane
// adds 16-byte pieces (red.global.add.v4.f32).  Rows the forward replaced by the learned empty feature add to g_empty
// instead (per-block partial sums first: every such point hits the same C addresses).
__global__ void __launch_bounds__(256) sample_features_bwd_kernel(FieldParams fp, const float *__restrict__ xyz, long long N,
                                                                  const float *__restrict__ g_feat, int d_in,
                                                                  float *__restrict__ g_map, float *__restrict__ g_empty) {
    extern __shared__ float s_empty[];                      // [C] partial sums of this block
    __shared__ float cam[21];
    const int C = fp.C;
    for (int i = threadIdx.x; i < C; i += 256) s_empty[i] = 0.0f;
    for (int i = threadIdx.x; i < 21; i += 256) cam[i] = i < 9 ? __ldg(fp.K_f + i) : __ldg(fp.w2c_f + (i - 9));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((long long)gridDim.x * 256) >> 5;
    bool any_empty = false;
    for (long long n = wid; n < N; n += nw) {
        float x, y, zc;
        bool inv;
        project_point(cam, cam + 9, __ldg(xyz + 3 * n), __ldg(xyz + 3 * n + 1), __ldg(xyz + 3 * n + 2), x, y, zc, inv);
        x = clamp_keep_nan(x, -2.0f, 2.0f);
        y = clamp_keep_nan(y, -2.0f, 2.0f);
        const float *g = g_feat + (size_t)n * d_in;
        if (fp.learn_empty && inv) {
            any_empty = true;
            for (int c = lane; c < C; c += 32) atomicAdd(&s_empty[c], __ldg(g + c));
            continue;
        }
        const Tap t = bilinear_tap(x, y, fp.Hf, fp.Wf);
        float *m = g_map + ((size_t)t.y0 * fp.Wf + t.x0) * C;
        const size_t dx = C, dy = (size_t)fp.Wf * C;
        for (int c = 4 * lane; c < C; c += 128) {
            float4 v;
            if (((uintptr_t)(g + c) & 15) == 0) v = __ldg(reinterpret_cast<const float4 *>(g + c));
            else v = make_float4(__ldg(g + c), __ldg(g + c + 1), __ldg(g + c + 2), __ldg(g + c + 3));
            auto add = [&](float *dst, float wt) {
                atomicAdd(reinterpret_cast<float4 *>(dst), make_float4(v.x * wt, v.y * wt, v.z * wt, v.w * wt));
            };
            add(m + c, t.wnw);
            if (t.in_x1) add(m + dx + c, t.wne);
            if (t.in_y1) add(m + dy + c, t.wsw);
            if (t.in_x1 && t.in_y1) add(m + dy + dx + c, t.wse);
        }
    }
    __syncthreads();
    if (g_empty && __syncthreads_or(any_empty))
        for (int i = threadIdx.x; i < C; i += 256)
            if (s_empty[i] != 0.0f) atomicAdd(g_empty + i, s_empty[i]);
}

}  // namespace sd

using namespace sd;

extern "C" int sd_sample_features_bwd(const sd_scene *scene, const float *xyz, long long N, const float *g_feat, float *g_map,
                                      float *g_empty, void *stream) {
    FieldParams fp;
    int rc = make_field_params(scene, &fp);
    if (rc) return rc;
    SD_REQUIRE(N >= 0, "sd_sample_features_bwd: bad N");
    if (N == 0) return SD_OK;
    SD_REQUIRE(xyz && g_feat && g_map, "sd_sample_features_bwd: null pointer");
    SD_REQUIRE(fp.C % 4 == 0 && fp.C <= 4096 && ((uintptr_t)g_map & 15) == 0, "sd_sample_features_bwd: C %% 4 == 0 and a 16-byte aligned map gradient");
    SD_REQUIRE(!fp.learn_empty || g_empty, "sd_sample_features_bwd: learn_empty needs g_empty");
    const long long warps = N < 148 * 64 ? N : 148 * 64;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    sample_features_bwd_kernel<<<grid, 256, (size_t)fp.C * 4, (cudaStream_t)stream>>>(fp, xyz, N, g_feat, fp.C + fp.code_dim, g_map, g_empty);
    SD_LAUNCH_OK("sample_features_bwd_kernel");
    return SD_OK;
}

extern "C" int sd_positional_encoding(const float *x, long long N, int d_in, int num_freqs, float freq_factor, int include_input,
                                      float *out, void *stream) {
    SD_REQUIRE(N >= 0 && d_in > 0 && num_freqs >= 0 && num_freqs <= 30, "sd_positional_encoding: bad shape");
    if (N == 0) return SD_OK;
    SD_REQUIRE(x && out, "sd_positional_encoding: null pointer");
    const long long total = N * ((include_input ? d_in : 0) + 2ll * num_freqs * d_in);
    const long long blocks = (total + 255) / 256;
    positional_encoding_kernel<<<(unsigned)(blocks < 65535 * 8 ? blocks : 65535 * 8), 256, 0, (cudaStream_t)stream>>>(
        x, N, d_in, num_freqs, freq_factor, include_input, out);
    SD_LAUNCH_OK("positional_encoding_kernel");
    return SD_OK;
}
