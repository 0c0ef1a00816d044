// vt_rgb.cu -- K1b: NV12 -> packed RGB24.  (filled in once the swscale nv12->rgb24 semantics are pinned)
#include "vt_common.cuh"
namespace vt {
int launch_nv12_to_rgb24(const uint8_t *, int, size_t, int, int, uint8_t *, size_t, int, cudaStream_t) {
    set_error("vt_nv12_to_rgb24: not built yet");
    return VT_ERR_UNSUPPORTED;
}
}  // namespace vt
