// vt_api.cu -- C ABI glue for libvtseg.so (declared in include/vtseg.h).
#include <stdarg.h>

#include "vt_common.cuh"

namespace vt {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;  // B200
    }
    return n;
}

int launch_score(const uint8_t *, int, size_t, int, int, const uint8_t *, int, uint64_t *, uint32_t *, cudaStream_t);
int launch_nv12_to_yuv420p(const uint8_t *, int, size_t, int, int, uint8_t *, size_t, int, cudaStream_t);
int launch_nv12_to_rgb24(const uint8_t *, int, size_t, int, int, uint8_t *, size_t, int, cudaStream_t);
int launch_gather(const uint8_t *, size_t, size_t, const int32_t *, int, uint8_t *, cudaStream_t);

}  // namespace vt

extern "C" {

int vt_version(void) { return 100; }
const char *vt_last_error(void) { return vt::g_err; }
uint64_t vt_launch_count(void) { return vt::g_launches.load(); }

int vt_sad_hist_u8(const uint8_t *luma, int pitch, size_t frame_stride, int w, int h, const uint8_t *prev0,
                   int n_frames, uint64_t *sad, uint32_t *hist, void *stream) {
    return vt::launch_score(luma, pitch, frame_stride, w, h, prev0, n_frames, sad, hist, (cudaStream_t)stream);
}

int vt_nv12_to_yuv420p(const uint8_t *src, int pitch, size_t src_fs, int w, int h, uint8_t *dst, size_t dst_fs,
                       int n_frames, void *stream) {
    return vt::launch_nv12_to_yuv420p(src, pitch, src_fs, w, h, dst, dst_fs, n_frames, (cudaStream_t)stream);
}

int vt_nv12_to_rgb24(const uint8_t *src, int pitch, size_t src_fs, int w, int h, uint8_t *dst, size_t dst_fs,
                     int n_frames, void *stream) {
    return vt::launch_nv12_to_rgb24(src, pitch, src_fs, w, h, dst, dst_fs, n_frames, (cudaStream_t)stream);
}

int vt_gather_frames(const uint8_t *src, size_t src_fs, size_t frame_bytes, const int32_t *index, int count,
                     uint8_t *dst, void *stream) {
    return vt::launch_gather(src, src_fs, frame_bytes, index, count, dst, (cudaStream_t)stream);
}

}  // extern "C"
