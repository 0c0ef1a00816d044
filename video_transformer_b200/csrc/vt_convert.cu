// vt_convert.cu -- K1 (NV12 -> planar YUV420P, exact copy semantics) and K5 (segment frame gather).
// Pure streaming kernels: 128-bit non-allocating loads/stores, one pass, grid sized from the SM count.
#include <algorithm>

#include "vt_common.cuh"

namespace vt {

// ---- NV12 -> YUV420P -------------------------------------------------------------------------------------
// One block = 16 rows of one picture (Y rows first, then UV rows), 128 x 4 threads; a thread moves one 16-byte group
// (Y) or one 32-byte group of 16 UV pairs (chroma) of four rows, all loads issued before the first store.  No
// divisions: the picture, the row block and the group come from the block and thread indices.
// Fast path needs: src 16 B aligned with pitch % 16 == 0, w % 32 == 0 (so dst Y rows and dst U/V rows are
// 16 B aligned).  Everything else takes the byte kernel.
__global__ void __launch_bounds__(512)
nv12_to_yuv420p_kernel(const uint8_t *__restrict__ src, int pitch, size_t src_fs, int w, int h,
                       uint8_t *__restrict__ dst, size_t dst_fs, int y_blocks) {
    const int cw = w >> 1, ch = h >> 1;
    const uint8_t *s = src + (size_t)blockIdx.z * src_fs;
    uint8_t *d = dst + (size_t)blockIdx.z * dst_fs;
    if ((int)blockIdx.y < y_blocks) {
        const int r0 = blockIdx.y * 16 + threadIdx.y;
        for (int g = threadIdx.x; g < (w >> 4); g += blockDim.x) {
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r0 + 4 * i < h) v[i] = ld_stream_u4(s + (size_t)(r0 + 4 * i) * pitch + (size_t)g * 16);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r0 + 4 * i < h) st_stream_u4(d + (size_t)(r0 + 4 * i) * w + (size_t)g * 16, v[i]);
        }
    } else {
        const int r0 = ((int)blockIdx.y - y_blocks) * 16 + threadIdx.y;
        const uint8_t *suv = s + (size_t)pitch * h;
        uint8_t *du0 = d + (size_t)w * h;
        for (int g = threadIdx.x; g < (w >> 5); g += blockDim.x) {
            uint4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r0 + 4 * i < ch) {
                    const uint8_t *sp = suv + (size_t)(r0 + 4 * i) * pitch + (size_t)g * 32;
                    a[i] = ld_stream_u4(sp);
                    b[i] = ld_stream_u4(sp + 16);
                }
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r0 + 4 * i < ch) {
                    uint4 u, v;
                    u.x = __byte_perm(a[i].x, a[i].y, 0x6420); v.x = __byte_perm(a[i].x, a[i].y, 0x7531);
                    u.y = __byte_perm(a[i].z, a[i].w, 0x6420); v.y = __byte_perm(a[i].z, a[i].w, 0x7531);
                    u.z = __byte_perm(b[i].x, b[i].y, 0x6420); v.z = __byte_perm(b[i].x, b[i].y, 0x7531);
                    u.w = __byte_perm(b[i].z, b[i].w, 0x6420); v.w = __byte_perm(b[i].z, b[i].w, 0x7531);
                    uint8_t *du = du0 + (size_t)(r0 + 4 * i) * cw + (size_t)g * 16;
                    st_stream_u4(du, u);
                    st_stream_u4(du + (size_t)cw * ch, v);
                }
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
        const int yb = (h + 15) / 16, cb = (h / 2 + 15) / 16;
        for (int f0 = 0; f0 < n_frames; f0 += 65535) {                      // gridDim.z limit
            const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
            nv12_to_yuv420p_kernel<<<dim3(1, yb + cb, nf), dim3(128, 4), 0, st>>>(
                src + (size_t)f0 * src_fs, pitch, src_fs, w, h, dst + (size_t)f0 * dst_fs, dst_fs, yb);
            VT_LAUNCHED("nv12_to_yuv420p_kernel");
        }
    } else {
        nv12_to_yuv420p_bytes_kernel<<<grid, 256, 0, st>>>(src, pitch, src_fs, w, h, dst, dst_fs, n_frames);
        VT_LAUNCHED("nv12_to_yuv420p_bytes_kernel");
    }
    return VT_OK;
}

// ---- K5: gather whole frames by index into a contiguous segment buffer --------------------------------------
// blockIdx.y = output frame; a block copies one 16 KB slice of it (256 threads x four 16-byte groups, loads first).
__global__ void __launch_bounds__(256)
gather_frames_kernel(const uint8_t *__restrict__ src, size_t src_fs, size_t frame_bytes,
                     const int32_t *__restrict__ index, int count, uint8_t *__restrict__ dst) {
    const size_t groups = frame_bytes >> 4;
    const size_t k = blockIdx.y;
    const size_t fi = index ? (size_t)index[k] : k;
    const uint8_t *s = src + fi * src_fs;
    uint8_t *d = dst + k * frame_bytes;
    for (size_t g0 = (size_t)blockIdx.x * 1024 + threadIdx.x; g0 < groups; g0 += (size_t)gridDim.x * 1024) {
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (g0 + 256 * i < groups) v[i] = ld_stream_u4(s + (g0 + 256 * i) * 16);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (g0 + 256 * i < groups) st_stream_u4(d + (g0 + 256 * i) * 16, v[i]);
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
        const unsigned gx = (unsigned)std::min<size_t>(64, ((frame_bytes >> 4) + 1023) / 1024);
        for (int k0 = 0; k0 < count; k0 += 65535) {                          // gridDim.y limit
            const int nk = count - k0 < 65535 ? count - k0 : 65535;
            gather_frames_kernel<<<dim3(gx, nk), 256, 0, st>>>(index ? src : src + (size_t)k0 * src_fs, src_fs, frame_bytes,
                                                               index ? index + k0 : nullptr, nk,
                                                               dst + (size_t)k0 * frame_bytes);
            VT_LAUNCHED("gather_frames_kernel");
        }
    } else {
        gather_frames_bytes_kernel<<<grid, 256, 0, st>>>(src, src_fs, frame_bytes, index, count, dst);
        VT_LAUNCHED("gather_frames_bytes_kernel");
    }
    return VT_OK;
}

}  // namespace vt
