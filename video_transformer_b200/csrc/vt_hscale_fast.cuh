// vt_hscale_fast.cuh -- horizontal polyphase taps through dp2a for word-aligned NV12 surfaces (taps <= 32), shared by
// the RGB path (vt_rgb.cu) and by the batched two-pass scaler (vt_scale.cu) for planes the pair kernel does not take.
// Results are libswscale's 15-bit intermediates: min((sum pixel * coef14) >> 7, 32767).
#pragma once
#include "vt_common.cuh"

namespace vt {

__device__ __forceinline__ int hs_dp2a_lo(uint32_t coef_pair, uint32_t pix, int acc) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef_pair), "r"(pix), "r"(acc));
    return d;
}
__device__ __forceinline__ int hs_dp2a_hi(uint32_t coef_pair, uint32_t pix, int acc) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef_pair), "r"(pix), "r"(acc));
    return d;
}

// HP coefficient pairs of one output sample, fetched with the widest aligned loads (rows of the table are HP words)
template <int HP>
__device__ __forceinline__ void hs_load_pairs(const uint32_t *__restrict__ t, uint32_t (&c)[HP]) {
    if constexpr (HP % 4 == 0) {
#pragma unroll
        for (int i = 0; i < HP / 4; i++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(t) + i);
            c[4 * i] = q.x; c[4 * i + 1] = q.y; c[4 * i + 2] = q.z; c[4 * i + 3] = q.w;
        }
    } else if constexpr (HP % 2 == 0) {
#pragma unroll
        for (int i = 0; i < HP / 2; i++) {
            const uint2 q = __ldg(reinterpret_cast<const uint2 *>(t) + i);
            c[2 * i] = q.x; c[2 * i + 1] = q.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < HP; i++) c[i] = __ldg(t + i);
    }
}

// Horizontal taps, luma: one thread per output sample, HP coefficient pairs.  The taps' bytes are fetched as aligned
// 32-bit words (clamped to the row's last word: bytes past the taps carry zero coefficients) and funnel-shifted into
// place; each dp2a multiplies two pixels by two 14-bit coefficients.
constexpr int HS_RPT = 8;     // source rows per thread of the horizontal kernels: position, shift and coefficients are
                               // fetched once and the per-row work is loads + dp2a (the one-row form spent 40 of its 62
                               // instructions on indices)
template <int HP>
__global__ void __launch_bounds__(256)
hscale_luma_fast(const uint8_t *__restrict__ src, int pitch, size_t src_fs, int rows, int16_t *__restrict__ mid,
                     size_t mid_fs, int dw, const uint32_t *__restrict__ coef2, const int32_t *__restrict__ pos) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= dw) return;
    constexpr int NAW = (HP + 1) / 2, NW = NAW + 1;
    const int r0 = blockIdx.y * HS_RPT, nr = min(HS_RPT, rows - r0);
    const int a = __ldg(pos + x);
    const int w0 = a >> 2, wl = (pitch >> 2) - 1, pw = pitch >> 2;
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    int wi[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) wi[i] = min(w0 + i, wl);
    uint32_t c[HP];
    hs_load_pairs<HP>(coef2 + (size_t)x * HP, c);
    const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (size_t)blockIdx.z * src_fs + (size_t)r0 * pitch);
    int16_t *mo = mid + (size_t)blockIdx.z * mid_fs + (size_t)r0 * dw + x;
    // Long windows (the 12- and 16-pair banks of the large ratios): all but the row's last few samples lie inside the
    // row, and their words then sit at constant offsets from ONE pointer (measured 7 % on 3840x2160 -> 640x360; for the
    // short banks the second code path only costs registers: 122 per thread for the 8-pair chroma kernel).
    constexpr bool DUAL = HP >= 12;
    const bool inside = DUAL && w0 + NW - 1 <= wl;
    const uint32_t *base = row + w0;
#pragma unroll 4
    for (int r = 0; r < nr; r++) {
        uint32_t w[NW];
        if (DUAL && inside) {
#pragma unroll
            for (int i = 0; i < NW; i++) w[i] = __ldg(base + i);
        } else {
#pragma unroll
            for (int i = 0; i < NW; i++) w[i] = __ldg(row + wi[i]);
        }
        base += pw;
        int v = 0;
#pragma unroll
        for (int i = 0; i < HP; i++) {
            const uint32_t al = __funnelshift_r(w[i >> 1], w[(i >> 1) + 1], sh);
            v = (i & 1) ? hs_dp2a_hi(c[i], al, v) : hs_dp2a_lo(c[i], al, v);
        }
        *mo = (int16_t)min(v >> 7, 32767);
        row += pw;
        mo += dw;
    }
}

// Horizontal taps, NV12 chroma: one thread per output sample produces U and V.  Sample pair (t, t+1) is one aligned
// word U_t V_t U_t+1 V_t+1 after the funnel shift; a byte permute makes it (U_t, U_t+1, V_t, V_t+1), so dp2a.lo is
// the U taps and dp2a.hi the V taps with the same coefficient pair.
template <int HP>
__global__ void __launch_bounds__(256)
hscale_chroma_fast(const uint8_t *__restrict__ src, int pitch, size_t src_fs, int rows, int16_t *__restrict__ mu,
                       int16_t *__restrict__ mv, size_t mid_fs, int cdw, const uint32_t *__restrict__ coef2,
                       const int32_t *__restrict__ pos) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= cdw) return;
    const int r0 = blockIdx.y * HS_RPT, nr = min(HS_RPT, rows - r0);
    const int a = 2 * __ldg(pos + x);                  // byte offset of the first U sample
    const int w0 = a >> 2, wl = (pitch >> 2) - 1, pw = pitch >> 2;
    const uint32_t sh = (uint32_t)(a & 3) * 8u;        // 0 or 16
    int wi[HP + 1];
#pragma unroll
    for (int i = 0; i < HP + 1; i++) wi[i] = min(w0 + i, wl);
    uint32_t c[HP];
    hs_load_pairs<HP>(coef2 + (size_t)x * HP, c);
    const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (size_t)blockIdx.z * src_fs + (size_t)r0 * pitch);
    size_t o = (size_t)blockIdx.z * mid_fs + (size_t)r0 * cdw + x;
    constexpr bool DUAL = HP >= 12;                    // see hscale_luma_fast
    const bool inside = DUAL && w0 + HP <= wl;
    const uint32_t *base = row + w0;
#pragma unroll 4
    for (int r = 0; r < nr; r++) {
        uint32_t w[HP + 1];
        if (DUAL && inside) {
#pragma unroll
            for (int i = 0; i < HP + 1; i++) w[i] = __ldg(base + i);
        } else {
#pragma unroll
            for (int i = 0; i < HP + 1; i++) w[i] = __ldg(row + wi[i]);
        }
        base += pw;
        int u = 0, v = 0;
#pragma unroll
        for (int i = 0; i < HP; i++) {
            const uint32_t pw4 = __byte_perm(__funnelshift_r(w[i], w[i + 1], sh), 0u, 0x3120);
            u = hs_dp2a_lo(c[i], pw4, u);
            v = hs_dp2a_hi(c[i], pw4, v);
        }
        mu[o] = (int16_t)min(u >> 7, 32767);
        mv[o] = (int16_t)min(v >> 7, 32767);
        row += pw;
        o += cdw;
    }
}


// coefficient pairs per output sample, padded to an instantiated count (0 = more than 32 taps: general kernels).
// 12 and 16 pairs serve the large ratios (3840x2160 -> 640x360 is 6:1: 24 bicubic taps; 1920x1080 -> 426x240 4.5:1).
inline int hscale_fast_pairs(int taps) {
    const int hp = (taps + 1) / 2;
    return hp <= 2 ? 2 : hp <= 3 ? 3 : hp <= 4 ? 4 : hp <= 6 ? 6 : hp <= 8 ? 8 : hp <= 12 ? 12 : hp <= 16 ? 16 : 0;
}

}  // namespace vt
