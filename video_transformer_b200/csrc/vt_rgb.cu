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
// This is a low-volume path (one frame per second of video): per chunk of frames, horizontal taps into 15-bit
// intermediates (one launch for luma, one for U and V together), then one kernel that runs the three vertical filters
// and the matrix per pixel pair and writes 6 bytes.  The fast kernels (word-aligned surfaces, taps <= 16) take the
// horizontal taps two at a time through dp2a on funnel-shifted 32-bit words -- interleaved chroma gives U and V from
// the same permuted word -- and unroll the vertical taps; 32-bit arithmetic throughout (every product fits).  The
// first version (byte loads, 64-bit index math) executed 7.9 M warp-instructions per 768x768 picture; the general
// kernels below are kept for surfaces that are not 4-byte aligned and for banks with more than 16 taps.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "vt_common.cuh"
#include "vt_hscale_fast.cuh"

struct vt_rgb_plan {
    int sw, sh, dw, dh, flags;
    int csw, csh, cdw;
    int lht, lvt, cht, cvt;                   // taps: luma h/v, chroma h/v
    int16_t *lhc, *lvc, *chc, *cvc;           // coefficient banks (device)
    int32_t *lhp, *lvp, *chp, *cvp;           // first source sample of every output sample
    int chunk;                                // frames per launch group
    int16_t *my, *mu, *mv;                    // intermediates for `chunk` frames
    // fast kernels: coefficient PAIRS per output sample, padded to hpl / hpc pairs; vertical banks padded to vtl / vtc
    int hpl, hpc, vtl, vtc;                   // 0 = that bank is outside what the fast kernels take
    uint32_t *lhc2, *chc2;                    // dw x hpl, cdw x hpc
    int16_t *lvc2, *cvc2;                     // dh x vtl, dh x vtc
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


// Vertical taps + matrix, fast form: LVT / CVT taps unrolled (banks padded with zero coefficients; padded taps read a
// clamped row), luma intermediates of a pixel pair fetched as one 32-bit word, 32-bit arithmetic (the largest product,
// 255 * cbu, is below 2^25), three 16-bit stores per pair.
__device__ __forceinline__ int rgb_t32(int idx, int cy) { return max(0, min(255, (idx * cy - (400 << 16) + 0x8000) >> 16)); }

template <int LVT, int CVT>
__global__ void __launch_bounds__(256)
rgb_vscale_fast(const int16_t *__restrict__ my, const int16_t *__restrict__ mu, const int16_t *__restrict__ mv,
                size_t my_fs, size_t mc_fs, int dw, int cdw, int sh, int csh, const int16_t *__restrict__ lvc2,
                const int32_t *__restrict__ lvp, const int16_t *__restrict__ cvc2, const int32_t *__restrict__ cvp,
                uint8_t *__restrict__ dst, size_t dst_fs, RgbConst k) {
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;    // pair index; dw is even, so every pair has two pixels
    const int y = blockIdx.y;
    if (xp >= cdw) return;
    const uint32_t *yy = reinterpret_cast<const uint32_t *>(my + (size_t)blockIdx.z * my_fs) + xp;
    const int16_t *uu = mu + (size_t)blockIdx.z * mc_fs + xp;
    const int16_t *vv = mv + (size_t)blockIdx.z * mc_fs + xp;
    const int lr = __ldg(lvp + y), cr = __ldg(cvp + y), hw = dw >> 1;
    uint32_t wy[LVT];
    int su[CVT], sv[CVT];
#pragma unroll
    for (int j = 0; j < LVT; j++) wy[j] = __ldg(yy + min(lr + j, sh - 1) * hw);
#pragma unroll
    for (int j = 0; j < CVT; j++) {
        const int o = min(cr + j, csh - 1) * cdw;
        su[j] = __ldg(uu + o);
        sv[j] = __ldg(vv + o);
    }
    int y0 = 1 << 18, y1 = 1 << 18, u = 1 << 18, v = 1 << 18;
#pragma unroll
    for (int j = 0; j < LVT; j++) {
        const int c = __ldg(lvc2 + y * LVT + j);
        y0 += ((int)(wy[j] << 16) >> 16) * c;
        y1 += ((int)wy[j] >> 16) * c;
    }
#pragma unroll
    for (int j = 0; j < CVT; j++) {
        const int c = __ldg(cvc2 + y * CVT + j);
        u += su[j] * c;
        v += sv[j] * c;
    }
    y0 >>= 19; y1 >>= 19; u >>= 19; v >>= 19;
    const int uc = max(0, min(255, u)), vc = max(0, min(255, v));
    const int r_off = ((vc * k.crv) >> 16) - (k.crv >> 9);
    const int g_off = ((uc * k.cgu) >> 16) - (k.cgu >> 9) + ((vc * k.cgv) >> 16) - (k.cgv >> 9);
    const int b_off = ((uc * k.cbu) >> 16) - (k.cbu >> 9);
    const int r0 = rgb_t32(y0 + 326 + r_off, k.cy), g0 = rgb_t32(y0 + 326 + g_off, k.cy), b0 = rgb_t32(y0 + 326 + b_off, k.cy);
    const int r1 = rgb_t32(y1 + 326 + r_off, k.cy), g1 = rgb_t32(y1 + 326 + g_off, k.cy), b1 = rgb_t32(y1 + 326 + b_off, k.cy);
    unsigned short *d = reinterpret_cast<unsigned short *>(dst + (size_t)blockIdx.z * dst_fs + ((size_t)y * dw + 2 * xp) * 3);
    d[0] = (unsigned short)(r0 | (g0 << 8));
    d[1] = (unsigned short)(b0 | (r1 << 8));
    d[2] = (unsigned short)(g1 | (b1 << 8));
}

// The same for EIGHT pixels per thread (output width a multiple of 8): one 128-bit load per luma tap row, one 64-bit
// load per chroma tap row and channel, and the matrix's clamp comes free with the byte pack (cvt.pack.sat, two
// values per instruction; the affine part of T() is folded per channel and pixel pair:
// (Y + 326 + off) * cy + K = Y * cy + A_channel).  24 output bytes leave as three 64-bit stores.
__device__ __forceinline__ uint32_t rgb_pack4(int b3, int b2, int b1, int b0) {   // sat_u8 of each, b0 in the low byte
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b3), "r"(b2), "r"(0u));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b1), "r"(b0), "r"(t));
    return d;
}

template <int LVT, int CVT>
__global__ void __launch_bounds__(128)
rgb_vscale_fast8(const int16_t *__restrict__ my, const int16_t *__restrict__ mu, const int16_t *__restrict__ mv,
                 size_t my_fs, size_t mc_fs, int dw, int cdw, int sh, int csh, const int16_t *__restrict__ lvc2,
                 const int32_t *__restrict__ lvp, const int16_t *__restrict__ cvc2, const int32_t *__restrict__ cvp,
                 uint8_t *__restrict__ dst, size_t dst_fs, RgbConst k) {
    const int xq = blockIdx.x * blockDim.x + threadIdx.x;    // group of 8 pixels = 4 chroma samples
    const int y = blockIdx.y;
    const int nq = dw >> 3;
    if (xq >= nq) return;
    const uint4 *yy = reinterpret_cast<const uint4 *>(my + (size_t)blockIdx.z * my_fs) + xq;
    const uint2 *uu = reinterpret_cast<const uint2 *>(mu + (size_t)blockIdx.z * mc_fs) + xq;
    const uint2 *vv = reinterpret_cast<const uint2 *>(mv + (size_t)blockIdx.z * mc_fs) + xq;
    const int lr = __ldg(lvp + y), cr = __ldg(cvp + y);
    uint4 wy[LVT];
    uint2 wu[CVT], wv[CVT];
#pragma unroll
    for (int j = 0; j < LVT; j++) wy[j] = __ldg(yy + min(lr + j, sh - 1) * nq);
#pragma unroll
    for (int j = 0; j < CVT; j++) {
        const int o = min(cr + j, csh - 1) * nq;
        wu[j] = __ldg(uu + o);
        wv[j] = __ldg(vv + o);
    }
    int ya[8], ua[4], va[4];
#pragma unroll
    for (int i = 0; i < 8; i++) ya[i] = 1 << 18;
#pragma unroll
    for (int i = 0; i < 4; i++) ua[i] = va[i] = 1 << 18;
#pragma unroll
    for (int j = 0; j < LVT; j++) {
        const int c = __ldg(lvc2 + y * LVT + j);
        const uint32_t w[4] = {wy[j].x, wy[j].y, wy[j].z, wy[j].w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            ya[2 * i] += ((int)(w[i] << 16) >> 16) * c;
            ya[2 * i + 1] += ((int)w[i] >> 16) * c;
        }
    }
#pragma unroll
    for (int j = 0; j < CVT; j++) {
        const int c = __ldg(cvc2 + y * CVT + j);
        ua[0] += ((int)(wu[j].x << 16) >> 16) * c; ua[1] += ((int)wu[j].x >> 16) * c;
        ua[2] += ((int)(wu[j].y << 16) >> 16) * c; ua[3] += ((int)wu[j].y >> 16) * c;
        va[0] += ((int)(wv[j].x << 16) >> 16) * c; va[1] += ((int)wv[j].x >> 16) * c;
        va[2] += ((int)(wv[j].y << 16) >> 16) * c; va[3] += ((int)wv[j].y >> 16) * c;
    }
    const int base = 326 * k.cy - (400 << 16) + 0x8000;
    int px[24];                                              // (T's argument) >> 16 per output byte, before the clamp
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int uc = max(0, min(255, ua[q] >> 19)), vc = max(0, min(255, va[q] >> 19));
        const int r_off = ((vc * k.crv) >> 16) - (k.crv >> 9);
        const int g_off = ((uc * k.cgu) >> 16) - (k.cgu >> 9) + ((vc * k.cgv) >> 16) - (k.cgv >> 9);
        const int b_off = ((uc * k.cbu) >> 16) - (k.cbu >> 9);
        const int ar = r_off * k.cy + base, ag = g_off * k.cy + base, ab = b_off * k.cy + base;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int yv = ya[2 * q + e] >> 19;
            px[6 * q + 3 * e] = (yv * k.cy + ar) >> 16;
            px[6 * q + 3 * e + 1] = (yv * k.cy + ag) >> 16;
            px[6 * q + 3 * e + 2] = (yv * k.cy + ab) >> 16;
        }
    }
    uint2 *d = reinterpret_cast<uint2 *>(dst + (size_t)blockIdx.z * dst_fs + ((size_t)y * dw + 8 * (size_t)xq) * 3);
#pragma unroll
    for (int i = 0; i < 3; i++)
        d[i] = make_uint2(rgb_pack4(px[8 * i + 3], px[8 * i + 2], px[8 * i + 1], px[8 * i]),
                          rgb_pack4(px[8 * i + 7], px[8 * i + 6], px[8 * i + 5], px[8 * i + 4]));
}

}  // namespace vt

namespace {

int upload(const void *h, size_t n, void **d) {
    if (cudaMalloc(d, n ? n : 1) != cudaSuccess) return VT_ERR_NOMEM;
    if (n && cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice) != cudaSuccess) return VT_ERR_CUDA;
    return VT_OK;
}

int bank(int src, int dst, int flags, int one, int16_t **coef_dev, int32_t **pos_dev, int *taps,
         std::vector<int16_t> *coef_host) {
    int cap = vt_sws_max_taps(src, dst, flags);
    if (cap < 0) return cap;
    cap = std::max(cap, 4);
    std::vector<int16_t> coef((size_t)dst * cap, 0);
    std::vector<int32_t> pos((size_t)dst, 0);
    int rc = vt_sws_make_filter(src, dst, flags, one, coef.data(), pos.data(), taps);
    if (rc) return rc;
    coef.resize((size_t)dst * *taps);
    *coef_host = coef;
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
    std::vector<int16_t> hl, vl, hc, vc;
    int rc = bank(sw, dw, flags, 1 << 14, &p->lhc, &p->lhp, &p->lht, &hl);
    if (rc == VT_OK) rc = bank(sh, dh, flags, 1 << 12, &p->lvc, &p->lvp, &p->lvt, &vl);
    if (rc == VT_OK) rc = bank(p->csw, p->cdw, flags, 1 << 14, &p->chc, &p->chp, &p->cht, &hc);
    if (rc == VT_OK) rc = bank(p->csh, dh, flags, 1 << 12, &p->cvc, &p->cvp, &p->cvt, &vc);
    // tables of the fast kernels: horizontal coefficient pairs, vertical banks padded to an instantiated tap count
    auto pad_hp = [](int taps) { const int hp = (taps + 1) / 2; return hp <= 2 ? 2 : hp <= 3 ? 3 : hp <= 4 ? 4 : hp <= 6 ? 6 : hp <= 8 ? 8 : 0; };
    auto pad_vt = [](int taps) { return taps <= 4 ? 4 : taps <= 6 ? 6 : taps <= 8 ? 8 : 0; };
    auto pairs = [](const std::vector<int16_t> &c, int n, int taps, int hp) {
        std::vector<uint32_t> t((size_t)n * hp, 0);
        for (int x = 0; x < n; x++)
            for (int j = 0; j < taps; j++) {
                const uint32_t v = (uint16_t)c[(size_t)x * taps + j];
                t[(size_t)x * hp + j / 2] |= (j & 1) ? (v << 16) : v;
            }
        return t;
    };
    auto padded = [](const std::vector<int16_t> &c, int n, int taps, int vt) {
        std::vector<int16_t> t((size_t)n * vt, 0);
        for (int y = 0; y < n; y++)
            for (int j = 0; j < taps; j++) t[(size_t)y * vt + j] = c[(size_t)y * taps + j];
        return t;
    };
    if (rc == VT_OK) {
        p->hpl = pad_hp(p->lht); p->hpc = pad_hp(p->cht); p->vtl = pad_vt(p->lvt); p->vtc = pad_vt(p->cvt);
        if (p->hpl && p->hpc && p->vtl && p->vtc) {
            const auto a = pairs(hl, dw, p->lht, p->hpl), b = pairs(hc, p->cdw, p->cht, p->hpc);
            const auto c = padded(vl, dh, p->lvt, p->vtl), d = padded(vc, dh, p->cvt, p->vtc);
            rc = upload(a.data(), a.size() * 4, (void **)&p->lhc2);
            if (rc == VT_OK) rc = upload(b.data(), b.size() * 4, (void **)&p->chc2);
            if (rc == VT_OK) rc = upload(c.data(), c.size() * 2, (void **)&p->lvc2);
            if (rc == VT_OK) rc = upload(d.data(), d.size() * 2, (void **)&p->cvc2);
        } else {
            p->hpl = p->hpc = p->vtl = p->vtc = 0;
        }
    }
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
    cudaFree(p->lhc2); cudaFree(p->chc2); cudaFree(p->lvc2); cudaFree(p->cvc2);
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
    // VT_RGB_KERNEL=generic selects the general kernels (A/B measurements and their own parity test)
    static const char *force = getenv("VT_RGB_KERNEL");
    const bool fast = p->hpl && !(force && !strcmp(force, "generic")) && (uintptr_t)src % 4 == 0 && pitch % 4 == 0 &&
                      src_fs % 4 == 0 && (uintptr_t)dst % 2 == 0 && dst_fs % 2 == 0;
    for (int f0 = 0; f0 < n_frames; f0 += p->chunk) {
        const int nf = std::min(p->chunk, n_frames - f0);
        const uint8_t *s = src + (size_t)f0 * src_fs;
        const uint8_t *uv = s + (size_t)pitch * p->sh;
        dim3 b(256);
        uint8_t *d = dst + (size_t)f0 * dst_fs;
        if (fast) {
            const dim3 gl((p->dw + 255) / 256, (p->sh + vt::HS_RPT - 1) / vt::HS_RPT, nf),
                gc((p->cdw + 255) / 256, (p->csh + vt::HS_RPT - 1) / vt::HS_RPT, nf), gv((p->cdw + 255) / 256, p->dh, nf);
#define VT_HL(H) case H: vt::hscale_luma_fast<H><<<gl, b, 0, st>>>(s, pitch, src_fs, p->sh, p->my, my_fs, p->dw, p->lhc2, p->lhp); break
            switch (p->hpl) { VT_HL(2); VT_HL(3); VT_HL(4); VT_HL(6); VT_HL(8); }
#undef VT_HL
            VT_LAUNCHED("rgb_hscale_luma_fast");
#define VT_HC(H) case H: vt::hscale_chroma_fast<H><<<gc, b, 0, st>>>(uv, pitch, src_fs, p->csh, p->mu, p->mv, mc_fs, p->cdw, p->chc2, p->chp); break
            switch (p->hpc) { VT_HC(2); VT_HC(3); VT_HC(4); VT_HC(6); VT_HC(8); }
#undef VT_HC
            VT_LAUNCHED("rgb_hscale_chroma_fast");
            // eight pixels per thread when the rows of the output and of the intermediates are 8-byte aligned
            const bool wide = p->dw % 8 == 0 && (uintptr_t)d % 8 == 0 && dst_fs % 8 == 0;
            const dim3 gv8((p->dw / 8 + 127) / 128, p->dh, nf);
#define VT_V(L, C) if (p->vtl == L && p->vtc == C) { \
        if (wide) vt::rgb_vscale_fast8<L, C><<<gv8, 128, 0, st>>>(p->my, p->mu, p->mv, my_fs, mc_fs, p->dw, p->cdw, p->sh, p->csh, \
                                                                p->lvc2, p->lvp, p->cvc2, p->cvp, d, dst_fs, k); \
        else vt::rgb_vscale_fast<L, C><<<gv, b, 0, st>>>(p->my, p->mu, p->mv, my_fs, mc_fs, p->dw, p->cdw, p->sh, p->csh, \
                                                        p->lvc2, p->lvp, p->cvc2, p->cvp, d, dst_fs, k); }
            VT_V(4, 4) VT_V(4, 6) VT_V(4, 8) VT_V(6, 4) VT_V(6, 6) VT_V(6, 8) VT_V(8, 4) VT_V(8, 6) VT_V(8, 8)
#undef VT_V
            VT_LAUNCHED("rgb_vscale_fast");
            continue;
        }
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
            d, dst_fs, k);
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
