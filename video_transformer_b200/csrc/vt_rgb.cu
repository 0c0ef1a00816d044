// vt_rgb.cu -- K1b / config 5: NV12 -> packed RGB24, optionally scaled, with libswscale's nv12->rgb24 semantics
// (BASELINE.json configs[4] "1 fps frame sampling + 768x768 RGB resize for upload").
//
// The reference never produces RGB; the north star defines this output as what
// `ffmpeg -vf scale=W:H -pix_fmt rgb24` would write, i.e. libswscale's general path with SWS_BICUBIC:
//   luma    horizontal bank sw -> dw,            vertical bank sh -> dh
//   chroma  horizontal bank ceil(sw/2) -> ceil(dw/2) (two output pixels share one chroma sample),
//           vertical bank   ceil(sh/2) -> dh          (chroma is interpolated vertically to full height)
//   Y,U,V = (2^18 + sum mid15 * coef12) >> 19, then the table form of the BT.601 limited-range matrix
//   (yuv2rgb.c): T(i) = clip_u8((i*cy - (400<<16) + 0x8000) >> 16), cy = 65536*255/219,
//       R = T(Y + 326 + off(V,crv)), G = T(Y + 326 + off(U,cgu) + off(V,cgv)), B = T(Y + 326 + off(U,cbu)),
//       off(c,k) = ((clip_u8(c)*k) >> 16) - (k >> 9).
// oracle/vt_oracle.c:vto_yuv_to_rgb24 is the CPU restatement, pinned bit-for-bit against libswscale 9.1.100.
//
// This is a low-volume path (one frame per second of video), so it is three simple kernels per chunk of frames:
// horizontal taps into 15-bit intermediates (luma, U, V), then one kernel that runs the three vertical filters
// and the matrix per pixel pair and writes 6 bytes.  HBM traffic is dominated by the intermediates.
#include <algorithm>
#include <mutex>
#include <vector>

#include "vt_common.cuh"

struct vt_rgb_plan {
    int sw, sh, dw, dh, flags;
    int csw, csh, cdw;
    int lht, lvt, cht, cvt;                   // taps: luma h/v, chroma h/v
    int16_t *lhc, *lvc, *chc, *cvc;           // coefficient banks (device)
    int32_t *lhp, *lvp, *chp, *cvp;           // first source sample of every output sample
    int chunk;                                // frames per launch group
    int16_t *my, *mu, *mv;                    // intermediates for `chunk` frames
};

namespace vt {

// horizontal taps of `chunk` frames: blockIdx.z = frame
__global__ void __launch_bounds__(256)
rgb_hscale_kernel(const uint8_t *__restrict__ src, int pitch, size_t src_fs, int rows, int step, int off,
                  int16_t *__restrict__ mid, size_t mid_fs, int dw, const int16_t *__restrict__ coef,
                  const int32_t *__restrict__ pos, int taps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (x >= dw || r >= rows) return;
    const uint8_t *s = src + (size_t)blockIdx.z * src_fs + (size_t)r * pitch + (size_t)pos[x] * step + off;
    const int16_t *c = coef + (size_t)x * taps;
    int v = 0;
    for (int j = 0; j < taps; j++) v += (int)s[(size_t)j * step] * (int)c[j];
    v >>= 7;
    mid[(size_t)blockIdx.z * mid_fs + (size_t)r * dw + x] = (int16_t)min(v, 32767);
}

struct RgbConst {
    int cy, crv, cbu, cgu, cgv;
};

__device__ __forceinline__ int rgb_t(int idx, int cy) {
    const long long v = ((long long)idx * cy - (400LL << 16) + 0x8000) >> 16;
    return (int)max(0LL, min(255LL, v));
}

// one thread per horizontal pixel pair: three vertical filters, the matrix, 6 bytes out
__global__ void __launch_bounds__(256)
rgb_vscale_kernel(const int16_t *__restrict__ my, const int16_t *__restrict__ mu, const int16_t *__restrict__ mv,
                  size_t my_fs, size_t mc_fs, int dw, int cdw, int dh, const int16_t *__restrict__ lvc,
                  const int32_t *__restrict__ lvp, int lvt, const int16_t *__restrict__ cvc,
                  const int32_t *__restrict__ cvp, int cvt, uint8_t *__restrict__ dst, size_t dst_fs, RgbConst k) {
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;    // pair index
    const int y = blockIdx.y;
    if (xp >= cdw || y >= dh) return;
    const int16_t *yy = my + (size_t)blockIdx.z * my_fs + (size_t)lvp[y] * dw + 2 * xp;
    const bool two = 2 * xp + 1 < dw;
    int y0 = 1 << 18, y1 = 1 << 18, u = 1 << 18, v = 1 << 18;
    for (int j = 0; j < lvt; j++) {
        const int c = lvc[(size_t)y * lvt + j];
        y0 += yy[(size_t)j * dw] * c;
        if (two) y1 += yy[(size_t)j * dw + 1] * c;
    }
    const int16_t *uu = mu + (size_t)blockIdx.z * mc_fs + (size_t)cvp[y] * cdw + xp;
    const int16_t *vv = mv + (size_t)blockIdx.z * mc_fs + (size_t)cvp[y] * cdw + xp;
    for (int j = 0; j < cvt; j++) {
        const int c = cvc[(size_t)y * cvt + j];
        u += uu[(size_t)j * cdw] * c;
        v += vv[(size_t)j * cdw] * c;
    }
    y0 >>= 19; y1 >>= 19; u >>= 19; v >>= 19;
    const int uc = max(0, min(255, u)), vc = max(0, min(255, v));
    const int r_off = (int)(((long long)vc * k.crv) >> 16) - (k.crv >> 9);
    const int g_off = (int)(((long long)uc * k.cgu) >> 16) - (k.cgu >> 9) + (int)(((long long)vc * k.cgv) >> 16) - (k.cgv >> 9);
    const int b_off = (int)(((long long)uc * k.cbu) >> 16) - (k.cbu >> 9);
    uint8_t *d = dst + (size_t)blockIdx.z * dst_fs + ((size_t)y * dw + 2 * xp) * 3;
    d[0] = (uint8_t)rgb_t(y0 + 326 + r_off, k.cy);
    d[1] = (uint8_t)rgb_t(y0 + 326 + g_off, k.cy);
    d[2] = (uint8_t)rgb_t(y0 + 326 + b_off, k.cy);
    if (two) {
        d[3] = (uint8_t)rgb_t(y1 + 326 + r_off, k.cy);
        d[4] = (uint8_t)rgb_t(y1 + 326 + g_off, k.cy);
        d[5] = (uint8_t)rgb_t(y1 + 326 + b_off, k.cy);
    }
}

}  // namespace vt

namespace {

int upload(const void *h, size_t n, void **d) {
    if (cudaMalloc(d, n ? n : 1) != cudaSuccess) return VT_ERR_NOMEM;
    if (n && cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice) != cudaSuccess) return VT_ERR_CUDA;
    return VT_OK;
}

int bank(int src, int dst, int flags, int one, int16_t **coef_dev, int32_t **pos_dev, int *taps) {
    int cap = vt_sws_max_taps(src, dst, flags);
    if (cap < 0) return cap;
    cap = std::max(cap, 4);
    std::vector<int16_t> coef((size_t)dst * cap, 0);
    std::vector<int32_t> pos((size_t)dst, 0);
    int rc = vt_sws_make_filter(src, dst, flags, one, coef.data(), pos.data(), taps);
    if (rc) return rc;
    rc = upload(coef.data(), (size_t)dst * *taps * sizeof(int16_t), (void **)coef_dev);
    if (rc == VT_OK) rc = upload(pos.data(), (size_t)dst * sizeof(int32_t), (void **)pos_dev);
    return rc;
}

// truncating division, as C does it in yuv2rgb.c
long long cdiv(long long a, long long b) { return a / b; }

}  // namespace

extern "C" int vt_rgb_plan_create(int sw, int sh, int dw, int dh, int flags, vt_rgb_plan **out) {
    if (!out || sw < 2 || sh < 2 || dw < 2 || dh < 2) {
        vt::set_error("vt_rgb_plan_create: bad size %dx%d -> %dx%d", sw, sh, dw, dh);
        return VT_ERR_INVALID;
    }
    if (dw & 1) {
        vt::set_error("vt_rgb_plan_create: odd output width %d is outside what the oracle pins", dw);
        return VT_ERR_UNSUPPORTED;
    }
    vt_rgb_plan *p = new (std::nothrow) vt_rgb_plan();
    if (!p) return VT_ERR_NOMEM;
    *p = vt_rgb_plan{};
    p->sw = sw; p->sh = sh; p->dw = dw; p->dh = dh; p->flags = flags;
    p->csw = (sw + 1) / 2; p->csh = (sh + 1) / 2; p->cdw = (dw + 1) / 2;
    int rc = bank(sw, dw, flags, 1 << 14, &p->lhc, &p->lhp, &p->lht);
    if (rc == VT_OK) rc = bank(sh, dh, flags, 1 << 12, &p->lvc, &p->lvp, &p->lvt);
    if (rc == VT_OK) rc = bank(p->csw, p->cdw, flags, 1 << 14, &p->chc, &p->chp, &p->cht);
    if (rc == VT_OK) rc = bank(p->csh, dh, flags, 1 << 12, &p->cvc, &p->cvp, &p->cvt);
    // intermediates: as many frames per launch group as fit 256 MB
    const size_t per_frame = ((size_t)dw * sh + 2 * (size_t)p->cdw * p->csh) * sizeof(int16_t);
    p->chunk = (int)std::max<size_t>(1, std::min<size_t>(32, ((size_t)256 << 20) / per_frame));
    if (rc == VT_OK && cudaMalloc((void **)&p->my, (size_t)p->chunk * dw * sh * sizeof(int16_t)) != cudaSuccess) rc = VT_ERR_NOMEM;
    if (rc == VT_OK && cudaMalloc((void **)&p->mu, (size_t)p->chunk * p->cdw * p->csh * sizeof(int16_t)) != cudaSuccess) rc = VT_ERR_NOMEM;
    if (rc == VT_OK && cudaMalloc((void **)&p->mv, (size_t)p->chunk * p->cdw * p->csh * sizeof(int16_t)) != cudaSuccess) rc = VT_ERR_NOMEM;
    if (rc != VT_OK) {
        vt::set_error("vt_rgb_plan_create: failed (%d) for %dx%d -> %dx%d flags=0x%x", rc, sw, sh, dw, dh, flags);
        vt_rgb_plan_destroy(p);
        return rc;
    }
    *out = p;
    return VT_OK;
}

extern "C" void vt_rgb_plan_destroy(vt_rgb_plan *p) {
    if (!p) return;
    cudaFree(p->lhc); cudaFree(p->lvc); cudaFree(p->chc); cudaFree(p->cvc);
    cudaFree(p->lhp); cudaFree(p->lvp); cudaFree(p->chp); cudaFree(p->cvp);
    cudaFree(p->my); cudaFree(p->mu); cudaFree(p->mv);
    delete p;
}

extern "C" int vt_scale_nv12_to_rgb24(const vt_rgb_plan *p, const uint8_t *src, int pitch, size_t src_fs, uint8_t *dst,
                                      size_t dst_fs, int n_frames, void *stream) {
    if (!p || !src || !dst || n_frames <= 0 || pitch < p->sw || pitch < 2 * p->csw) {
        vt::set_error("vt_scale_nv12_to_rgb24: bad arguments");
        return VT_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    vt::RgbConst k;
    k.cy = (int)((65536LL * 255) / 219);
    k.crv = (int)cdiv(104597LL * 65536 + 0x8000, k.cy);
    k.cbu = (int)cdiv(132201LL * 65536 + 0x8000, k.cy);
    k.cgu = (int)cdiv(-25675LL * 65536 + 0x8000, k.cy);
    k.cgv = (int)cdiv(-53279LL * 65536 + 0x8000, k.cy);
    const size_t my_fs = (size_t)p->dw * p->sh, mc_fs = (size_t)p->cdw * p->csh;
    for (int f0 = 0; f0 < n_frames; f0 += p->chunk) {
        const int nf = std::min(p->chunk, n_frames - f0);
        const uint8_t *s = src + (size_t)f0 * src_fs;
        const uint8_t *uv = s + (size_t)pitch * p->sh;
        dim3 b(256);
        vt::rgb_hscale_kernel<<<dim3((p->dw + 255) / 256, p->sh, nf), b, 0, st>>>(s, pitch, src_fs, p->sh, 1, 0, p->my, my_fs,
                                                                             p->dw, p->lhc, p->lhp, p->lht);
        VT_LAUNCHED("rgb_hscale_kernel");
        vt::rgb_hscale_kernel<<<dim3((p->cdw + 255) / 256, p->csh, nf), b, 0, st>>>(uv, pitch, src_fs, p->csh, 2, 0, p->mu,
                                                                               mc_fs, p->cdw, p->chc, p->chp, p->cht);
        VT_LAUNCHED("rgb_hscale_kernel");
        vt::rgb_hscale_kernel<<<dim3((p->cdw + 255) / 256, p->csh, nf), b, 0, st>>>(uv, pitch, src_fs, p->csh, 2, 1, p->mv,
                                                                               mc_fs, p->cdw, p->chc, p->chp, p->cht);
        VT_LAUNCHED("rgb_hscale_kernel");
        vt::rgb_vscale_kernel<<<dim3((p->cdw + 255) / 256, p->dh, nf), b, 0, st>>>(
            p->my, p->mu, p->mv, my_fs, mc_fs, p->dw, p->cdw, p->dh, p->lvc, p->lvp, p->lvt, p->cvc, p->cvp, p->cvt,
            dst + (size_t)f0 * dst_fs, dst_fs, k);
        VT_LAUNCHED("rgb_vscale_kernel");
    }
    return VT_OK;
}

// Same-size NV12 -> RGB24: the scaled path with unit horizontal/vertical luma banks (libswscale takes the same
// route and the oracle pins it).  Plans are cached per (w, h).
namespace vt {
int launch_nv12_to_rgb24(const uint8_t *src, int pitch, size_t src_fs, int w, int h, uint8_t *dst, size_t dst_fs,
                         int n_frames, cudaStream_t st) {
    static std::mutex mu;
    static std::vector<vt_rgb_plan *> cache;
    vt_rgb_plan *p = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (vt_rgb_plan *q : cache)
            if (q->sw == w && q->sh == h && q->dw == w && q->dh == h) p = q;
        if (!p) {
            int rc = vt_rgb_plan_create(w, h, w, h, VT_SWS_BICUBIC, &p);
            if (rc) return rc;
            cache.push_back(p);
        }
    }
    return vt_scale_nv12_to_rgb24(p, src, pitch, src_fs, dst, dst_fs, n_frames, st);
}
}  // namespace vt
