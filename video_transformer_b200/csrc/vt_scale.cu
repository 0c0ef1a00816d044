// vt_scale.cu -- K2: libswscale-exact separable polyphase downscale.
//
// Arithmetic (8-bit in/out, the SWS_ACCURATE_RND|SWS_BITEXACT C path of libswscale 9.1.100):
//   horizontal:  mid[r][x] = min( (sum_j src[r][hpos[x]+j] * hcoef[x][j]) >> 7, 32767 )      (14-bit coefs)
//   vertical:    out[y][x] = clip_u8( (2^18 + sum_j mid[vpos[y]+j][x] * vcoef[y][j]) >> 19 ) (12-bit coefs)
//   vertical with one tap (axis not scaled): out = clip_u8( (mid + 64) >> 7 )
// The reference reaches this arithmetic through `ffmpeg -vf scale=-2:360`
// (/root/reference/src/analyzer/content_analyzer.py:193-211).
//
// Two implementations:
//   * generic : hscale kernel -> int16 scratch -> vscale kernel (this file).  Any ratio, any tap count, any
//               alignment.  Parity sweeps and shapes the production kernel does not take.
//   * pair    : vt_scale_pair.cu, the production kernel (one pass, per-warp TMA ring, register-resident
//               intermediates); this file owns the plan, the tensor maps and the dispatch between the two.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "vt_hscale_fast.cuh"
#include "vt_scale_plan.cuh"

namespace vt {

// ---- generic two-kernel path -----------------------------------------------------------------------------
// channel_step = 1 for planar sources, 2 for the interleaved UV plane of NV12 (channel_off picks U or V).
__global__ void __launch_bounds__(256)
hscale_generic_kernel(const uint8_t *__restrict__ src, int src_pitch, size_t src_fs, int sh, int channel_step,
                      int channel_off, int16_t *__restrict__ mid, size_t mid_fs, int dw,
                      const int16_t *__restrict__ coef, const int32_t *__restrict__ pos, int taps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (x >= dw || r >= sh) return;
    src += (size_t)blockIdx.z * src_fs;                              // blockIdx.z = picture of the chunk
    mid += (size_t)blockIdx.z * mid_fs;
    const uint8_t *s = src + (size_t)r * src_pitch + (size_t)pos[x] * channel_step + channel_off;
    const int16_t *c = coef + (size_t)x * taps;
    int v = 0;
    for (int j = 0; j < taps; j++) v += (int)s[(size_t)j * channel_step] * (int)c[j];
    v >>= 7;
    mid[(size_t)r * dw + x] = (int16_t)min(v, 32767);
}

__global__ void __launch_bounds__(256)
vscale_generic_kernel(const int16_t *__restrict__ mid, size_t mid_fs, int dw, uint8_t *__restrict__ dst, int dst_pitch,
                      size_t dst_fs, int dh, const int16_t *__restrict__ coef, const int32_t *__restrict__ pos, int taps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    mid += (size_t)blockIdx.z * mid_fs;
    dst += (size_t)blockIdx.z * dst_fs;
    int v;
    if (taps == 1) {
        v = (mid[(size_t)pos[y] * dw + x] + 64) >> 7;
    } else {
        v = 1 << 18;
        const int16_t *c = coef + (size_t)y * taps;
        const int16_t *m = mid + (size_t)pos[y] * dw + x;
        for (int j = 0; j < taps; j++) v += (int)m[(size_t)j * dw] * (int)c[j];
        v >>= 19;
    }
    dst[(size_t)y * dst_pitch + x] = (uint8_t)max(0, min(255, v));
}

// Vertical taps, VT unrolled (bank padded with zero coefficients; padded taps read a clamped row), 32-bit indices.
// One output row per blockIdx.y, so everything that depends on y is block-uniform: the VT coefficients arrive through
// the widest aligned loads the padded bank allows, and rows whose window lies inside the plane (all but the last few)
// walk a pointer instead of clamping an index per tap -- 3 instructions per tap instead of ~10, which is what the
// 24- and 32-tap banks of the large ratios (3840x2160 -> 640x360) are made of.
template <int VT>
__device__ __forceinline__ void vs_load_coef(const int16_t *__restrict__ c, int (&k)[VT]) {
    if constexpr (VT % 8 == 0) {
#pragma unroll
        for (int i = 0; i < VT / 8; i++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(c) + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                k[8 * i + 2 * j] = (int)(int16_t)(w[j] & 0xFFFFu);
                k[8 * i + 2 * j + 1] = (int)(int16_t)(w[j] >> 16);
            }
        }
    } else if constexpr (VT % 2 == 0) {
#pragma unroll
        for (int i = 0; i < VT / 2; i++) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(c) + i);
            k[2 * i] = (int)(int16_t)(w & 0xFFFFu);
            k[2 * i + 1] = (int)(int16_t)(w >> 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i < VT; i++) k[i] = (int)__ldg(c + i);
    }
}

template <int VT>
__global__ void __launch_bounds__(256)
vscale_fast_kernel(const int16_t *__restrict__ mid, size_t mid_fs, int dw, int sh, uint8_t *__restrict__ dst, int dst_pitch,
                   size_t dst_fs, const int16_t *__restrict__ vc2, const int32_t *__restrict__ vpos) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw) return;
    const int r0 = __ldg(vpos + y);
    int k[VT];
    vs_load_coef<VT>(vc2 + (size_t)y * VT, k);          // rows of the bank are VT * 2 bytes: 16-byte aligned for VT % 8 == 0
    int v = 1 << 18;
    if (r0 + VT <= sh) {                                 // block-uniform
        const int16_t *m = mid + (size_t)blockIdx.z * mid_fs + (size_t)r0 * dw + x;
#pragma unroll
        for (int j = 0; j < VT; j++) {
            v += (int)__ldg(m) * k[j];
            m += dw;
        }
    } else {
        const int16_t *m = mid + (size_t)blockIdx.z * mid_fs + x;
#pragma unroll
        for (int j = 0; j < VT; j++) v += (int)__ldg(m + min(r0 + j, sh - 1) * dw) * k[j];
    }
    dst[(size_t)blockIdx.z * dst_fs + (size_t)y * dst_pitch + x] = (uint8_t)max(0, min(255, v >> 19));
}

static int pad_vt2(int taps) {
    return taps < 2 ? 0 : taps <= 4 ? 4 : taps <= 6 ? 6 : taps <= 8 ? 8 : taps <= 12 ? 12 : taps <= 16 ? 16 : taps <= 24 ? 24
         : taps <= 32 ? 32 : 0;
}

template <int VT>
static void launch_vfast(const vt_scale_plan *p, int c, const int16_t *mid, size_t mid_fs, uint8_t *dst, size_t dst_fs, int nf,
                         cudaStream_t st) {
    const int sh = c ? p->csh : p->sh, dw = c ? p->cdw : p->dw, dh = c ? p->cdh : p->dh;
    vscale_fast_kernel<VT><<<dim3((dw + 255) / 256, dh, nf), 256, 0, st>>>(mid, mid_fs, dw, sh, dst, dw, dst_fs, p->vc2[c], p->vpos[c]);
}
static int vfast(const vt_scale_plan *p, int c, const int16_t *mid, size_t mid_fs, uint8_t *dst, size_t dst_fs, int nf,
                 cudaStream_t st) {
    switch (p->vt2[c]) {
        case 4: launch_vfast<4>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
        case 6: launch_vfast<6>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
        case 8: launch_vfast<8>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
        case 12: launch_vfast<12>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
        case 16: launch_vfast<16>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
        case 24: launch_vfast<24>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
        default: launch_vfast<32>(p, c, mid, mid_fs, dst, dst_fs, nf, st); break;
    }
    VT_LAUNCHED("vscale_fast_kernel");
    return VT_OK;
}

// One plane kind of n_frames NV12 pictures through the dp2a horizontal kernels (vt_hscale_fast.cuh) and the unrolled
// vertical kernel: luma -> dst_a; chroma -> U into dst_a and V into dst_b from ONE horizontal pass over the interleaved
// plane.  Requires can_fast(): word-aligned surfaces, taps within the instantiated range.
int scale_plane_fast(const vt_scale_plan *p, int c, const uint8_t *src, int pitch, size_t src_fs, uint8_t *dst_a,
                     uint8_t *dst_b, size_t dst_fs, int n_frames, cudaStream_t st) {
    const int sh = c ? p->csh : p->sh, dw = c ? p->cdw : p->dw;
    const size_t mid_fs = (size_t)dw * sh;
    int16_t *ma = p->scratch, *mb = p->scratch + (size_t)p->scratch_frames * mid_fs;   // chroma: 2 * cdw*csh <= dw*sh
    for (int f0 = 0; f0 < n_frames; f0 += p->scratch_frames) {
        const int nf = std::min(p->scratch_frames, n_frames - f0);
        const uint8_t *s = src + (size_t)f0 * src_fs;
        const dim3 g((dw + 255) / 256, (sh + HS_RPT - 1) / HS_RPT, nf);
        if (!c) {
#define VT_H(H) case H: hscale_luma_fast<H><<<g, 256, 0, st>>>(s, pitch, src_fs, sh, ma, mid_fs, dw, p->hc2[0], p->hpos[0]); break
            switch (p->hp2[0]) { VT_H(2); VT_H(3); VT_H(4); VT_H(6); VT_H(8); VT_H(12); VT_H(16); }
#undef VT_H
            VT_LAUNCHED("hscale_luma_fast");
            if (int rc = vfast(p, 0, ma, mid_fs, dst_a + (size_t)f0 * dst_fs, dst_fs, nf, st)) return rc;
        } else {
#define VT_H(H) case H: hscale_chroma_fast<H><<<g, 256, 0, st>>>(s, pitch, src_fs, sh, ma, mb, mid_fs, dw, p->hc2[1], p->hpos[1]); break
            switch (p->hp2[1]) { VT_H(2); VT_H(3); VT_H(4); VT_H(6); VT_H(8); VT_H(12); VT_H(16); }
#undef VT_H
            VT_LAUNCHED("hscale_chroma_fast");
            if (int rc = vfast(p, 1, ma, mid_fs, dst_a + (size_t)f0 * dst_fs, dst_fs, nf, st)) return rc;
            if (int rc = vfast(p, 1, mb, mid_fs, dst_b + (size_t)f0 * dst_fs, dst_fs, nf, st)) return rc;
        }
    }
    return VT_OK;
}

// n_frames pictures src_fs / dst_fs bytes apart, in chunks of p->scratch_frames (one launch pair per chunk)
int scale_plane_generic(const vt_scale_plan *p, int chroma, const uint8_t *src, int src_pitch, int channel_step,
                        int channel_off, uint8_t *dst, int dst_pitch, int n_frames, size_t src_fs, size_t dst_fs,
                        cudaStream_t st) {
    const int c = chroma ? 1 : 0;
    const int sh = c ? p->csh : p->sh, dw = c ? p->cdw : p->dw, dh = c ? p->cdh : p->dh;
    const size_t mid_fs = (size_t)p->dw * p->sh;                     // scratch is laid out for the larger (luma) plane
    for (int f0 = 0; f0 < n_frames; f0 += p->scratch_frames) {
        const int nf = std::min(p->scratch_frames, n_frames - f0);
        dim3 b(256), gh((dw + 255) / 256, sh, nf), gv((dw + 255) / 256, dh, nf);
        hscale_generic_kernel<<<gh, b, 0, st>>>(src + (size_t)f0 * src_fs, src_pitch, src_fs, sh, channel_step, channel_off,
                                                p->scratch, mid_fs, dw, p->hcoef[c], p->hpos[c], p->htaps[c]);
        VT_LAUNCHED("hscale_generic_kernel");
        vscale_generic_kernel<<<gv, b, 0, st>>>(p->scratch, mid_fs, dw, dst + (size_t)f0 * dst_fs, dst_pitch, dst_fs, dh,
                                                p->vcoef[c], p->vpos[c], p->vtaps[c]);
        VT_LAUNCHED("vscale_generic_kernel");
    }
    return VT_OK;
}

}  // namespace vt

// ---- plan ----------------------------------------------------------------------------------------------------
namespace {

int upload(const void *h, size_t n, void **d) {
    if (cudaMalloc(d, n ? n : 1) != cudaSuccess) return VT_ERR_NOMEM;
    if (n && cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice) != cudaSuccess) return VT_ERR_CUDA;
    return VT_OK;
}

int make_bank(int src, int dst, int flags, int one, std::vector<int16_t> &coef, std::vector<int32_t> &pos, int *taps) {
    int cap = vt_sws_max_taps(src, dst, flags);
    if (cap < 0) return cap;
    cap = cap < 4 ? 4 : cap;
    coef.assign((size_t)dst * cap, 0);
    pos.assign((size_t)dst, 0);
    int rc = vt_sws_make_filter(src, dst, flags, one, coef.data(), pos.data(), taps);
    if (rc) return rc;
    coef.resize((size_t)dst * *taps);
    return VT_OK;
}

}  // namespace

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

}  // namespace

// 3-D byte tensor (x bytes, rows, frames) with a (tile_w, tile_h, 1) box, no swizzle, zero fill outside.
int vt::make_tmap_u8_3d(void *tmap_out, const uint8_t *base, int row_bytes, int rows, int n_frames, int pitch, size_t frame_stride,
              int tile_w, int tile_h) {
    CUtensorMap *m = (CUtensorMap *)tmap_out;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        vt::set_error("cuTensorMapEncodeTiled not available from the driver");
        return VT_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)row_bytes, (cuuint64_t)rows, (cuuint64_t)n_frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {(cuuint32_t)tile_w, (cuuint32_t)tile_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        vt::set_error("cuTensorMapEncodeTiled failed (%d) row_bytes=%d rows=%d pitch=%d box=%dx%d", (int)r, row_bytes,
                      rows, pitch, tile_w, tile_h);
        return VT_ERR_CUDA;
    }
    return VT_OK;
}

// The same tensor seen as 32-bit elements (rows padded up to a multiple of 4 bytes, which the pitch covers), so that a
// box row can be up to 1024 bytes; x coordinates are then in units of 4 bytes.
int vt::make_tmap_u32_3d(void *tmap_out, const uint8_t *base, int row_bytes, int rows, int n_frames, int pitch,
                         size_t frame_stride, int tile_w, int tile_h) {
    CUtensorMap *m = (CUtensorMap *)tmap_out;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        vt::set_error("cuTensorMapEncodeTiled not available from the driver");
        return VT_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)((row_bytes + 3) / 4), (cuuint64_t)rows, (cuuint64_t)n_frames};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {(cuuint32_t)(tile_w / 4), (cuuint32_t)tile_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        vt::set_error("cuTensorMapEncodeTiled(u32) failed (%d) row_bytes=%d rows=%d pitch=%d box=%dx%d", (int)r, row_bytes,
                      rows, pitch, tile_w, tile_h);
        return VT_ERR_CUDA;
    }
    return VT_OK;
}

extern "C" int vt_scale_plan_stream_info(const vt_scale_plan *p, int chroma, int *info8) {
    if (!p || !info8) return VT_ERR_INVALID;
    const vt_scale_plan::Pair &pr = p->pair[chroma ? 1 : 0];
    info8[0] = p->pair[0].ok && p->pair[1].ok;
    info8[1] = pr.hp; info8[2] = pr.tv; info8[3] = 2 * pr.np; info8[4] = pr.stage_rows;
    info8[5] = pr.tile_w; info8[6] = pr.stage_rows * pr.n_stages; info8[7] = pr.warp_smem;
    return VT_OK;
}

extern "C" int vt_scale_plan_create(int sw, int sh, int dw, int dh, int flags, vt_scale_plan **out) {
    if (!out || sw < 2 || sh < 2 || dw < 2 || dh < 2) {
        vt::set_error("vt_scale_plan_create: bad size %dx%d -> %dx%d", sw, sh, dw, dh);
        return VT_ERR_INVALID;
    }
    vt_scale_plan *p = new (std::nothrow) vt_scale_plan();
    if (!p) return VT_ERR_NOMEM;
    p->sw = sw; p->sh = sh; p->dw = dw; p->dh = dh; p->flags = flags;
    p->csw = (sw + 1) >> 1; p->csh = (sh + 1) >> 1; p->cdw = (dw + 1) >> 1; p->cdh = (dh + 1) >> 1;
    for (int c = 0; c < 2; c++) {
        p->hcoef[c] = p->vcoef[c] = nullptr; p->hpos[c] = p->vpos[c] = nullptr;
        p->hc2[c] = nullptr; p->vc2[c] = nullptr; p->hp2[c] = p->vt2[c] = 0;
    }
    p->scratch = nullptr;
    int rc = VT_OK;
    for (int c = 0; c < 2 && rc == VT_OK; c++) {
        const int s_w = c ? p->csw : sw, s_h = c ? p->csh : sh, d_w = c ? p->cdw : dw, d_h = c ? p->cdh : dh;
        rc = make_bank(s_w, d_w, flags, 1 << 14, p->h_hcoef[c], p->h_hpos[c], &p->htaps[c]);
        if (rc == VT_OK) rc = make_bank(s_h, d_h, flags, 1 << 12, p->h_vcoef[c], p->h_vpos[c], &p->vtaps[c]);
        if (rc == VT_OK) rc = upload(p->h_hcoef[c].data(), p->h_hcoef[c].size() * 2, (void **)&p->hcoef[c]);
        if (rc == VT_OK) rc = upload(p->h_hpos[c].data(), p->h_hpos[c].size() * 4, (void **)&p->hpos[c]);
        if (rc == VT_OK) rc = upload(p->h_vcoef[c].data(), p->h_vcoef[c].size() * 2, (void **)&p->vcoef[c]);
        if (rc == VT_OK) rc = upload(p->h_vpos[c].data(), p->h_vpos[c].size() * 4, (void **)&p->vpos[c]);
    }
    // intermediates of the two-pass kernels: as many pictures per launch as fit 128 MB (at most 32)
    p->scratch_frames = (int)std::max<size_t>(1, std::min<size_t>(32, ((size_t)128 << 20) / ((size_t)dw * sh * sizeof(int16_t))));
    if (rc == VT_OK && cudaMalloc((void **)&p->scratch, (size_t)p->scratch_frames * dw * sh * sizeof(int16_t)) != cudaSuccess)
        rc = VT_ERR_NOMEM;
    for (int c = 0; c < 2 && rc == VT_OK; c++) {               // tables of the fast two-pass kernels
        const int n = c ? p->cdw : dw, nv = c ? p->cdh : dh, ht = p->htaps[c], vtp = p->vtaps[c];
        p->hp2[c] = vt::hscale_fast_pairs(ht);
        p->vt2[c] = vt::pad_vt2(vtp);
        p->hc2[c] = nullptr;
        p->vc2[c] = nullptr;
        if (!p->hp2[c] || !p->vt2[c]) { p->hp2[c] = p->vt2[c] = 0; continue; }
        std::vector<uint32_t> pairs((size_t)n * p->hp2[c], 0);
        for (int x = 0; x < n; x++)
            for (int j = 0; j < ht; j++) {
                const uint32_t v = (uint16_t)p->h_hcoef[c][(size_t)x * ht + j];
                pairs[(size_t)x * p->hp2[c] + j / 2] |= (j & 1) ? (v << 16) : v;
            }
        std::vector<int16_t> vpad((size_t)nv * p->vt2[c], 0);
        for (int y = 0; y < nv; y++)
            for (int j = 0; j < vtp; j++) vpad[(size_t)y * p->vt2[c] + j] = p->h_vcoef[c][(size_t)y * vtp + j];
        rc = upload(pairs.data(), pairs.size() * 4, (void **)&p->hc2[c]);
        if (rc == VT_OK) rc = upload(vpad.data(), vpad.size() * 2, (void **)&p->vc2[c]);
    }
    for (int c = 0; c < 2 && rc == VT_OK; c++) rc = vt::build_pair(p, c);
    if (rc != VT_OK) {
        vt::set_error("vt_scale_plan_create: failed (%d) for %dx%d -> %dx%d flags=0x%x", rc, sw, sh, dw, dh, flags);
        vt_scale_plan_destroy(p);
        return rc;
    }
    *out = p;
    return VT_OK;
}

extern "C" void vt_scale_plan_destroy(vt_scale_plan *p) {
    if (!p) return;
    for (int c = 0; c < 2; c++) {
        cudaFree(p->hcoef[c]); cudaFree(p->hpos[c]); cudaFree(p->vcoef[c]); cudaFree(p->vpos[c]);
        cudaFree(p->hc2[c]); cudaFree(p->vc2[c]);
    }
    vt::free_pair(p);
    cudaFree(p->scratch);
    delete p;
}

extern "C" int vt_scale_plane_u8(const vt_scale_plan *plan, int chroma, const uint8_t *src, int src_pitch,
                                 uint8_t *dst, int dst_pitch, void *stream) {
    if (!plan || !src || !dst) {
        vt::set_error("vt_scale_plane_u8: null argument");
        return VT_ERR_INVALID;
    }
    return vt::scale_plane_generic(plan, chroma, src, src_pitch, 1, 0, dst, dst_pitch, 1, 0, 0, (cudaStream_t)stream);
}

// 1 when vt_scale_score_nv12_to_yuv420p runs K3 inside the luma pass of K2 for this plan (suitably aligned buffers)
extern "C" int vt_scale_plan_fuses_score(const vt_scale_plan *p) { return p && p->pair[0].ok && p->pair[0].score_ok ? 1 : 0; }

// K2 + K3 in one call: scaled frames plus SAD / histogram of the SOURCE luma.  With VT_FUSED_SCORE=1 and a plan that
// allows it (exact 3:2 luma on the adjacent-column layout: 1080p -> 720p, 2160p -> 1440p ...) the luma kernel counts the
// source rows it already holds in shared memory, so the source luma is fetched from HBM once for both results; otherwise
// the score kernel runs first and the scaler after it.  Results are identical either way.
extern "C" int vt_scale_score_nv12_to_yuv420p(const vt_scale_plan *p, const uint8_t *src, int src_pitch, size_t src_fs,
                                              const uint8_t *prev0, uint8_t *dst, size_t dst_fs, int n_frames,
                                              uint64_t *sad_dev, uint32_t *hist_dev, void *stream) {
    if (!p || !src || !dst || !sad_dev || !hist_dev || n_frames <= 0 || src_pitch < p->sw) {
        vt::set_error("vt_scale_score_nv12_to_yuv420p: bad arguments");
        return VT_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = ((uintptr_t)src % 16 == 0) && (src_pitch % 16 == 0) && (src_fs % 16 == 0) &&
                         ((uintptr_t)dst % 8 == 0) && (dst_fs % 8 == 0) && (p->sw % 2 == 0) && (p->sh % 2 == 0) &&
                         (!prev0 || (uintptr_t)prev0 % 4 == 0);
    const char *force = getenv("VT_SCALE_KERNEL");
    // Opt-in (VT_FUSED_SCORE=1).  Measured on B200, 1080p -> 720p, 256 pictures: fused 0.469 ms vs 0.465 ms for the two
    // kernels one after the other, DRAM traffic 4.04 vs 5.04 MB per picture.  The histogram needs 16 KB of lane-private
    // counters per warp, which leaves 8 warps per SM; at that occupancy the fused pass runs at the SUM of its parts
    // (the scaler's multiply pipe and the histogram's shared-memory atomic unit do not overlap within one warp's
    // instruction stream), so the default stays with the separate kernels.
    const char *fuse = getenv("VT_FUSED_SCORE");
    if (aligned && vt_scale_plan_fuses_score(p) && !force && fuse && fuse[0] == '1') {
        VT_CUDA(cudaMemsetAsync(sad_dev, 0, sizeof(uint64_t) * (size_t)n_frames, st));
        VT_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(uint32_t) * 256 * (size_t)n_frames, st));
        const vt::PairScore sc{prev0, sad_dev, hist_dev};
        int rc = vt::launch_pair(p, 0, src, src_pitch, src_fs, dst, dst_fs, n_frames, st, &sc);
        if (rc) return rc;
        if (p->pair[1].ok) return vt::launch_pair(p, 1, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
        // (the chroma planes of such plans always qualify; kept for completeness)
        vt::set_error("vt_scale_score_nv12_to_yuv420p: chroma plane outside the pair kernel");
        return VT_ERR_UNSUPPORTED;
    }
    int rc = vt::launch_score(src, src_pitch, src_fs, p->sw, p->sh, prev0, n_frames, sad_dev, hist_dev, st);
    if (rc) return rc;
    return vt_scale_nv12_to_yuv420p(p, src, src_pitch, src_fs, dst, dst_fs, n_frames, stream);
}

extern "C" int vt_scale_nv12_to_yuv420p(const vt_scale_plan *p, const uint8_t *src, int src_pitch, size_t src_fs,
                                        uint8_t *dst, size_t dst_fs, int n_frames, void *stream) {
    if (!p || !src || !dst || n_frames <= 0 || src_pitch < p->sw) {
        vt::set_error("vt_scale_nv12_to_yuv420p: bad arguments");
        return VT_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ysz = (size_t)p->dw * p->dh, csz = (size_t)p->cdw * p->cdh;
    const bool aligned = ((uintptr_t)src % 16 == 0) && (src_pitch % 16 == 0) && (src_fs % 16 == 0) &&
                         ((uintptr_t)dst % 4 == 0) && (dst_fs % 4 == 0) && (p->sw % 2 == 0) && (p->sh % 2 == 0);
    // VT_SCALE_KERNEL=generic selects the two-pass kernels (A/B measurements only)
    static const char *force = getenv("VT_SCALE_KERNEL");
    // Each plane kind takes the pair kernel when its plan has one (even plane width, taps within the instantiated
    // range ...) and the batched two-pass kernels otherwise -- e.g. 1920x1080 -> 854x480: luma streams, the 427-wide
    // chroma planes (odd width: rows are not 2-byte aligned) go through the general kernels.
    const bool fast = aligned && !(force && !strcmp(force, "generic"));
    // (the fast two-pass kernels need word-aligned surfaces only; `aligned` is stricter)
    const bool fast2 = fast && !(force && !strcmp(force, "twopass-general"));
    const uint8_t *uv = src + (size_t)src_pitch * p->sh;
    int rc;
    if (fast && p->pair[0].ok) rc = vt::launch_pair(p, 0, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
    else if (fast2 && p->hp2[0]) rc = vt::scale_plane_fast(p, 0, src, src_pitch, src_fs, dst, nullptr, dst_fs, n_frames, st);
    else rc = vt::scale_plane_generic(p, 0, src, src_pitch, 1, 0, dst, p->dw, n_frames, src_fs, dst_fs, st);
    if (rc) return rc;
    if (fast && p->pair[1].ok) return vt::launch_pair(p, 1, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
    if (fast2 && p->hp2[1])
        return vt::scale_plane_fast(p, 1, uv, src_pitch, src_fs, dst + ysz, dst + ysz + csz, dst_fs, n_frames, st);
    rc = vt::scale_plane_generic(p, 1, uv, src_pitch, 2, 0, dst + ysz, p->cdw, n_frames, src_fs, dst_fs, st);
    if (rc) return rc;
    rc = vt::scale_plane_generic(p, 1, uv, src_pitch, 2, 1, dst + ysz + csz, p->cdw, n_frames, src_fs, dst_fs, st);
    if (rc) return rc;
    return VT_OK;
}

