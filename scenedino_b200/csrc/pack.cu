// One-off re-layout kernels: feature map planar -> channels-last, MLP weights -> packed blob.
#include "common.cuh"

namespace sd {

// ---- feature map: [n, C, H, W] fp32 -> [n, H, W, C] fp32 | fp16 ---------------------------------
// Tile = 64 channels x 32 pixels through shared memory: reads are 128 B per warp along W, writes
// are 256 B (fp32) / 128 B (fp16) per warp along C.  HBM-bound: 4*C*H*W read + esize*C*H*W written.
template <bool F16>
__global__ void __launch_bounds__(256) featmap_pack_kernel(const float *__restrict__ src, void *__restrict__ dst,
                                                           int C, long long HW) {
    __shared__ float tile[64][33];
    const int img = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
    const float *s = src + (size_t)img * C * HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + warp * 8 + i;
        const long long p = p0 + lane;
        tile[warp * 8 + i][lane] = (c < C && p < HW) ? __ldg(s + (size_t)c * HW + p) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pl = warp * 4 + i;
        const long long p = p0 + pl;
        if (p >= HW) continue;
        const int c = c0 + lane * 2;
        if (c >= C) continue;
        const float a = tile[lane * 2][pl], b = tile[lane * 2 + 1][pl];
        const size_t o = ((size_t)img * HW + p) * C + c;
        if (F16) {
            __half *d = reinterpret_cast<__half *>(dst);
            if (c + 1 < C) *reinterpret_cast<__half2 *>(d + o) = __floats2half2_rn(a, b);
            else d[o] = __float2half_rn(a);
        } else {
            float *d = reinterpret_cast<float *>(dst);
            if (c + 1 < C) *reinterpret_cast<float2 *>(d + o) = make_float2(a, b);
            else d[o] = a;
        }
    }
}

// ---- MLP blob ---------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

MlpLayout mlp_layout(int d_in, int d_hidden, int d_out) {
    MlpLayout L;
    L.d_in = d_in; L.d_hidden = d_hidden; L.d_out = d_out;
    // whole 64-wide K blocks (SW128 atoms) with at least 8 spare columns behind d_in: the tensor-core
    // image keeps the coordinate hi/lo split there (6) and the layer-1 bias as two constant-1 columns (2)
    L.d_in_pad = (int)align_up((size_t)d_in + 8, 64);
    L.d_out_pad = (int)align_up((size_t)d_out, 16);
    size_t o = 0;
    L.off_w_in_t = o;   o = align_up(o + sizeof(float) * (size_t)L.d_in_pad * d_hidden, 1024);
    L.off_b_in = o;     o = align_up(o + sizeof(float) * (size_t)d_hidden, 1024);
    L.off_w_out_t = o;  o = align_up(o + sizeof(float) * (size_t)d_hidden * L.d_out_pad, 1024);
    L.off_b_out = o;    o = align_up(o + sizeof(float) * (size_t)L.d_out_pad, 1024);
    L.off_w_in_h = o;  o = align_up(o + 2 * (size_t)L.d_in_pad * d_hidden, 1024);
    // tensor-core image of W_out: feature rows 1..d_out-1 first, then the density row 0, padded to 16 rows
    // ... plus one more 64-wide K block whose first two columns carry the output bias as half(b), b - half(b)
    // (field_bin.cu feeds them the constant 1 from TMEM, so the bias comes out of the layer-2 MMA)
    const int n2 = (int)align_up((size_t)d_out, 16);
    L.off_w_out_h = o; o = align_up(o + 2 * (size_t)n2 * align_up((size_t)d_hidden + 2, 64), 1024);
    L.off_w_sigma = o;  o = align_up(o + sizeof(float) * (size_t)d_hidden, 1024);
    L.off_w_sig_h = L.off_w_feat_blk = 0;
    if (d_hidden == 128) {
        L.off_w_sig_h = o;     o = align_up(o + 2 * (size_t)16 * 128, 1024);
        L.off_w_feat_blk = o;  o = align_up(o + (size_t)((d_out - 1 + 127) / 128) * 32768, 1024);
    }
    L.off_x_w2 = 0;
    if (d_in == 64 && d_hidden == 128 && d_out >= 128 && d_out % 128 == 0) {   // MlpDimReduction.transform_expand
        L.off_x_w2 = o; o = align_up(o + 2 * (size_t)d_out * d_hidden, 1024);
    }
    L.total = o;
    return L;
}

__global__ void mlp_pack_kernel(const float *__restrict__ w_in, const float *__restrict__ b_in,
                                const float *__restrict__ w_out, const float *__restrict__ b_out,
                                MlpLayout L, unsigned char *__restrict__ blob) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    float *w_in_t = reinterpret_cast<float *>(blob + L.off_w_in_t);
    float *bi = reinterpret_cast<float *>(blob + L.off_b_in);
    float *w_out_t = reinterpret_cast<float *>(blob + L.off_w_out_t);
    float *bo = reinterpret_cast<float *>(blob + L.off_b_out);
    __half *w_in_h = reinterpret_cast<__half *>(blob + L.off_w_in_h);
    __half *w_out_h = reinterpret_cast<__half *>(blob + L.off_w_out_h);
    float *w_sigma = reinterpret_cast<float *>(blob + L.off_w_sigma);
    const int H = L.d_hidden;
    // fp16 image, K padding put to use.
    //  * The raw coordinates (x, y, z') sit at columns d_in-39 .. d_in-37 of the field's input; z' reaches
    //    +-6e3 for points next to / behind the camera, where one half-precision product is coarse.  Columns
    //    d_in .. d_in+2 repeat half(w) for the low halves of the coordinates and d_in+3 .. d_in+5 hold
    //    w - half(w) for their high halves, so that x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo.
    //  * Columns d_in+6, d_in+7 carry the layer-1 bias as half(b) and b - half(b); the kernel feeds them
    //    the constant 1, so the bias comes out of the MMA and the epilogue needs no per-column constants.
    // The kernel fills the matching A columns; all other padding stays zero on both sides.
    const bool split = L.d_in >= 39;
    for (int i = tid; i < L.d_in_pad * H; i += nth) {
        const int k = i / H, j = i - k * H;
        const float v = k < L.d_in ? w_in[(size_t)j * L.d_in + k] : 0.0f;
        w_in_t[i] = v;
        float vb = v;
        const int e = k - L.d_in;
        if (split && e >= 0 && e < 6) {
            const float w = w_in[(size_t)j * L.d_in + (L.d_in - 39) + (e % 3)];
            vb = e < 3 ? w : w - __half2float(__float2half_rn(w));
        } else if (e == 6) {
            vb = b_in[j];
        } else if (e == 7) {
            vb = b_in[j] - __half2float(__float2half_rn(b_in[j]));
        }
        w_in_h[umma_sw128_offset(j, k, H) / 2] = __float2half_rn(vb);
    }
    for (int i = tid; i < H; i += nth) {
        bi[i] = b_in[i];
        w_sigma[i] = w_out[i];  // row 0 of W_out
    }
    for (int i = tid; i < H * L.d_out_pad; i += nth) {
        const int k = i / L.d_out_pad, o = i - k * L.d_out_pad;
        w_out_t[i] = o < L.d_out ? w_out[(size_t)o * H + k] : 0.0f;
    }
    for (int i = tid; i < L.d_out_pad; i += nth) bo[i] = i < L.d_out ? b_out[i] : 0.0f;
    const int n2 = (L.d_out + 15) / 16 * 16;
    const int Hp = (H + 2 + 63) / 64 * 64;
    for (int i = tid; i < n2 * Hp; i += nth) {
        const int r = i / Hp, k = i - r * Hp;  // image row r <-> W_out row r+1 (features), row d_out-1 <-> W_out row 0 (density)
        const int src = r < L.d_out - 1 ? r + 1 : (r == L.d_out - 1 ? 0 : -1);
        float v = (src >= 0 && k < H) ? w_out[(size_t)src * H + k] : 0.0f;
        if (src >= 0 && k == H) v = b_out[src];
        if (src >= 0 && k == H + 1) v = b_out[src] - __half2float(__float2half_rn(b_out[src]));
        w_out_h[umma_sw128_offset(r, k, n2) / 2] = __float2half_rn(v);
    }
    if (L.off_w_sig_h) {
        __half *w_sig_h = reinterpret_cast<__half *>(blob + L.off_w_sig_h);
        for (int i = tid; i < 16 * 128; i += nth) {
            const int r = i >> 7, k = i & 127;
            w_sig_h[umma_sw128_offset(r, k, 16) / 2] = __float2half_rn(r == 0 ? w_out[k] : 0.0f);
        }
        __half *w_fb = reinterpret_cast<__half *>(blob + L.off_w_feat_blk);
        const int nblk = (L.d_out - 1 + 127) / 128;
        for (int i = tid; i < nblk * 128 * 128; i += nth) {
            const int o = i >> 7, k = i & 127;          // feature output o <-> W_out row o + 1
            const float v = o < L.d_out - 1 ? w_out[(size_t)(o + 1) * H + k] : 0.0f;
            w_fb[((size_t)(o >> 7) * 32768 + umma_sw128_offset(o & 127, k, 128)) / 2] = __float2half_rn(v);
        }
    }
    if (L.off_x_w2) {   // blocks of 128 outputs in their natural order, each a complete B operand (K = 128) of 32 KB
        __half *x_w2 = reinterpret_cast<__half *>(blob + L.off_x_w2);
        for (int i = tid; i < L.d_out * H; i += nth) {
            const int o = i / H, k = i - o * H;
            x_w2[((size_t)(o >> 7) * 32768 + umma_sw128_offset(o & 127, k, 128)) / 2] = __float2half_rn(w_out[(size_t)o * H + k]);
        }
    }
}

}  // namespace sd

extern "C" int sd_featmap_pack(const float *nchw, int n_img, int C, int H, int W, void *nhwc,
                               int dst_dtype, void *stream) {
    SD_REQUIRE(nchw && nhwc, "sd_featmap_pack: null pointer");
    SD_REQUIRE(n_img > 0 && C > 0 && H > 0 && W > 0, "sd_featmap_pack: bad shape");
    SD_REQUIRE(C % 2 == 0, "sd_featmap_pack: C must be even (got %d)", C);
    SD_REQUIRE(dst_dtype == SD_F32 || dst_dtype == SD_F16, "sd_featmap_pack: bad dst_dtype");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 63) / 64), (unsigned)n_img);
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == SD_F16)
        sd::featmap_pack_kernel<true><<<grid, 256, 0, st>>>(nchw, nhwc, C, HW);
    else
        sd::featmap_pack_kernel<false><<<grid, 256, 0, st>>>(nchw, nhwc, C, HW);
    SD_LAUNCH_OK("featmap_pack_kernel");
    return SD_OK;
}

extern "C" size_t sd_mlp_pack_bytes(int d_in, int d_hidden, int d_out) {
    if (d_in <= 0 || d_hidden <= 0 || d_out <= 0) return 0;
    return sd::mlp_layout(d_in, d_hidden, d_out).total;
}

extern "C" int sd_mlp_pack(const float *w_in, const float *b_in, const float *w_out, const float *b_out,
                           int d_in, int d_hidden, int d_out, void *packed, void *stream) {
    SD_REQUIRE(w_in && b_in && w_out && b_out && packed, "sd_mlp_pack: null pointer");
    SD_REQUIRE(d_in > 0 && d_hidden > 0 && d_out > 0, "sd_mlp_pack: bad dims");
    SD_REQUIRE(((uintptr_t)packed & 127) == 0, "sd_mlp_pack: packed must be 128-byte aligned");
    const sd::MlpLayout L = sd::mlp_layout(d_in, d_hidden, d_out);
    sd::mlp_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w_in, b_in, w_out, b_out, L,
                                                              reinterpret_cast<unsigned char *>(packed));
    SD_LAUNCH_OK("mlp_pack_kernel");
    return SD_OK;
}
