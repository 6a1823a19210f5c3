// Fused tcgen05 path (placeholder until the kernel lands): reports "unsupported" loudly.
#include "common.cuh"
#include "launch.h"

namespace sd {

bool tc_supported(const sd_scene *, const sd_mlp *, int) { return false; }

int launch_field_tc(const FieldParams &, const PointSrc &, long long, const sd_mlp *, const TcRender *,
                    const TcOut &, cudaStream_t) {
    set_error("SD_MLP_BF16_TC: fused tcgen05 kernel not built into this library");
    return SD_ERR_INVALID;
}

int launch_mlp_tc(const sd_mlp *, const float *, long long, float *, cudaStream_t) {
    set_error("SD_MLP_BF16_TC: fused tcgen05 kernel not built into this library");
    return SD_ERR_INVALID;
}

}  // namespace sd
