// vt_convert.cu -- K1 (NV12 -> planar YUV420P, exact copy semantics) and K5 (segment frame gather).
// Pure streaming kernels: 128-bit non-allocating loads/stores, one pass, grid sized from the SM count.
#include "vt_common.cuh"

namespace vt {

// ---- NV12 -> YUV420P -------------------------------------------------------------------------------------
// Work item = one 16-byte group of a Y row, or one 32-byte group (16 UV pairs) of a UV row.
// Fast path needs: src 16 B aligned with pitch % 16 == 0, w % 32 == 0 (so dst Y rows and dst U/V rows are
// 16 B aligned).  Everything else takes the byte kernel.
__global__ void __launch_bounds__(256)
nv12_to_yuv420p_kernel(const uint8_t *__restrict__ src, int pitch, size_t src_fs, int w, int h,
                       uint8_t *__restrict__ dst, size_t dst_fs, int n_frames) {
    const int cw = w >> 1, ch = h >> 1;
    const int yg = w >> 4;       // 16 B groups per Y row
    const int cg = w >> 5;       // 32 B groups per UV row
    const long long per_frame = (long long)yg * h + (long long)cg * ch;
    const long long total = per_frame * n_frames;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / per_frame);
        long long j = i - (long long)f * per_frame;
        const uint8_t *s = src + (size_t)f * src_fs;
        uint8_t *d = dst + (size_t)f * dst_fs;
        if (j < (long long)yg * h) {
            const int r = (int)(j / yg), g = (int)(j - (long long)r * yg);
            st_stream_u4(d + (size_t)r * w + (size_t)g * 16, ld_stream_u4(s + (size_t)r * pitch + (size_t)g * 16));
        } else {
            j -= (long long)yg * h;
            const int r = (int)(j / cg), g = (int)(j - (long long)r * cg);
            const uint8_t *sp = s + (size_t)pitch * h + (size_t)r * pitch + (size_t)g * 32;
            const uint4 a = ld_stream_u4(sp), b = ld_stream_u4(sp + 16);
            uint4 u, v;
            u.x = __byte_perm(a.x, a.y, 0x6420); v.x = __byte_perm(a.x, a.y, 0x7531);
            u.y = __byte_perm(a.z, a.w, 0x6420); v.y = __byte_perm(a.z, a.w, 0x7531);
            u.z = __byte_perm(b.x, b.y, 0x6420); v.z = __byte_perm(b.x, b.y, 0x7531);
            u.w = __byte_perm(b.z, b.w, 0x6420); v.w = __byte_perm(b.z, b.w, 0x7531);
            uint8_t *du = d + (size_t)w * h + (size_t)r * cw + (size_t)g * 16;
            st_stream_u4(du, u);
            st_stream_u4(du + (size_t)cw * ch, v);
        }
    }
}

__global__ void __launch_bounds__(256)
nv12_to_yuv420p_bytes_kernel(const uint8_t *__restrict__ src, int pitch, size_t src_fs, int w, int h,
                             uint8_t *__restrict__ dst, size_t dst_fs, int n_frames) {
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    const long long per_frame = (long long)w * h + (long long)cw * ch;
    const long long total = per_frame * n_frames;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / per_frame);
        long long j = i - (long long)f * per_frame;
        const uint8_t *s = src + (size_t)f * src_fs;
        uint8_t *d = dst + (size_t)f * dst_fs;
        if (j < (long long)w * h) {
            const int r = (int)(j / w), x = (int)(j - (long long)r * w);
            d[j] = s[(size_t)r * pitch + x];
        } else {
            j -= (long long)w * h;
            const int r = (int)(j / cw), x = (int)(j - (long long)r * cw);
            const uint8_t *sp = s + (size_t)pitch * h + (size_t)r * pitch + 2 * (size_t)x;
            d[(size_t)w * h + j] = sp[0];
            d[(size_t)w * h + (size_t)cw * ch + j] = sp[1];
        }
    }
}

int launch_nv12_to_yuv420p(const uint8_t *src, int pitch, size_t src_fs, int w, int h, uint8_t *dst, size_t dst_fs,
                           int n_frames, cudaStream_t st) {
    if (!src || !dst || w <= 0 || h <= 0 || pitch < w || n_frames <= 0) {
        set_error("vt_nv12_to_yuv420p: bad arguments");
        return VT_ERR_INVALID;
    }
    const bool fast = (w % 32 == 0) && (h % 2 == 0) && (pitch % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                      ((uintptr_t)dst % 16 == 0) && (src_fs % 16 == 0) && (dst_fs % 16 == 0);
    const int grid = sm_count() * 8;
    if (fast) {
        nv12_to_yuv420p_kernel<<<grid, 256, 0, st>>>(src, pitch, src_fs, w, h, dst, dst_fs, n_frames);
        VT_LAUNCHED("nv12_to_yuv420p_kernel");
    } else {
        nv12_to_yuv420p_bytes_kernel<<<grid, 256, 0, st>>>(src, pitch, src_fs, w, h, dst, dst_fs, n_frames);
        VT_LAUNCHED("nv12_to_yuv420p_bytes_kernel");
    }
    return VT_OK;
}

// ---- K5: gather whole frames by index into a contiguous segment buffer --------------------------------------
__global__ void __launch_bounds__(256)
gather_frames_kernel(const uint8_t *__restrict__ src, size_t src_fs, size_t frame_bytes,
                     const int32_t *__restrict__ index, int count, uint8_t *__restrict__ dst) {
    const size_t groups = frame_bytes >> 4;
    const size_t total = groups * (size_t)count;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t k = i / groups, g = i - k * groups;
        const size_t fi = index ? (size_t)index[k] : k;
        st_stream_u4(dst + k * frame_bytes + g * 16, ld_stream_u4(src + fi * src_fs + g * 16));
    }
}
__global__ void __launch_bounds__(256)
gather_frames_bytes_kernel(const uint8_t *__restrict__ src, size_t src_fs, size_t frame_bytes,
                           const int32_t *__restrict__ index, int count, uint8_t *__restrict__ dst) {
    const size_t total = frame_bytes * (size_t)count;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t k = i / frame_bytes, b = i - k * frame_bytes;
        const size_t fi = index ? (size_t)index[k] : k;
        dst[i] = src[fi * src_fs + b];
    }
}

int launch_gather(const uint8_t *src, size_t src_fs, size_t frame_bytes, const int32_t *index, int count,
                  uint8_t *dst, cudaStream_t st) {
    if (!src || !dst || !frame_bytes || count <= 0) {
        set_error("vt_gather_frames: bad arguments");
        return VT_ERR_INVALID;
    }
    const bool fast = (frame_bytes % 16 == 0) && (src_fs % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                      ((uintptr_t)dst % 16 == 0);
    const int grid = sm_count() * 8;
    if (fast) {
        gather_frames_kernel<<<grid, 256, 0, st>>>(src, src_fs, frame_bytes, index, count, dst);
        VT_LAUNCHED("gather_frames_kernel");
    } else {
        gather_frames_bytes_kernel<<<grid, 256, 0, st>>>(src, src_fs, frame_bytes, index, count, dst);
        VT_LAUNCHED("gather_frames_bytes_kernel");
    }
    return VT_OK;
}

}  // namespace vt
