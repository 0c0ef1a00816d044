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

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= VT_MAX_DEVICES) dev = 0;
    return dev;
}

int sm_count() {
    static int n[VT_MAX_DEVICES] = {0};
    const int dev = current_device();
    if (!n[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;  // B200
        n[dev] = v;
    }
    return n[dev];
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

/* K5: page-lock an existing host range (a MAP_SHARED mapping of the `.frames` file) so the copy engine can write device
 * frames straight into the file.  On failure the runtime's last-error state is cleared, so the refusal (file systems whose
 * mappings cannot be pinned) does not surface at the next kernel launch. */
int vt_host_register(void *ptr, size_t n_bytes) {
    if (!ptr || !n_bytes) {
        vt::set_error("vt_host_register: bad arguments");
        return VT_ERR_INVALID;
    }
    cudaError_t e = cudaHostRegister(ptr, n_bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return vt::cuda_fail(e, "cudaHostRegister");
    }
    return VT_OK;
}

int vt_host_unregister(void *ptr) {
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return vt::cuda_fail(e, "cudaHostUnregister");
    }
    return VT_OK;
}

/* The same for a read-only use: the bitstream file is mapped privately and the copy engine reads the page cache
 * directly (no staging memcpy).  Tries cudaHostRegisterReadOnly first where the device supports it. */
int vt_host_register_source(void *ptr, size_t n_bytes) {
    if (!ptr || !n_bytes) {
        vt::set_error("vt_host_register_source: bad arguments");
        return VT_ERR_INVALID;
    }
    cudaError_t e = cudaHostRegister(ptr, n_bytes, cudaHostRegisterReadOnly);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = cudaHostRegister(ptr, n_bytes, cudaHostRegisterDefault);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return vt::cuda_fail(e, "cudaHostRegister (source mapping)");
    }
    return VT_OK;
}

int vt_copy_to_device_async(void *dst_dev, const void *src_host, size_t n_bytes, void *stream) {
    if (!dst_dev || !src_host) {
        vt::set_error("vt_copy_to_device_async: bad arguments");
        return VT_ERR_INVALID;
    }
    VT_CUDA(cudaMemcpyAsync(dst_dev, src_host, n_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return VT_OK;
}

/* Asynchronous device -> host copy on `stream` (the landing copy of K5; the host range must be page-locked for the copy
 * to overlap with kernels). */
int vt_copy_to_host_async(void *dst_host, const void *src_dev, size_t n_bytes, void *stream) {
    if (!dst_host || !src_dev) {
        vt::set_error("vt_copy_to_host_async: bad arguments");
        return VT_ERR_INVALID;
    }
    VT_CUDA(cudaMemcpyAsync(dst_host, src_dev, n_bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return VT_OK;
}

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
