// Small stand-alone kernels: PositionalEncoding.forward on its own (inside the field kernels the code is fused).
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

}  // namespace sd

using namespace sd;

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
