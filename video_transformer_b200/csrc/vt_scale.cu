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
//   * generic  : hscale kernel -> int16 scratch -> vscale kernel.  Any ratio, any tap count.  Parity sweeps.
//   * streaming: one pass, one warp per (frame, plane, 128-column strip, row chunk).  The warp's source tile
//                is staged in shared memory by TMA (cp.async.bulk.tensor, completion on an mbarrier); each
//                lane owns 4 output columns, runs the horizontal taps with dp2a (two 14-bit coefs x two
//                pixels per instruction) and keeps the vertical window of 15-bit intermediates in a
//                register ring, so intermediates never touch memory.  NV12 chroma is de-interleaved with
//                PRMT on the way.  HBM traffic = source once + destination once.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "vt_scale_plan.cuh"

namespace vt {

// ---- generic two-kernel path -----------------------------------------------------------------------------
// channel_step = 1 for planar sources, 2 for the interleaved UV plane of NV12 (channel_off picks U or V).
__global__ void __launch_bounds__(256)
hscale_generic_kernel(const uint8_t *__restrict__ src, int src_pitch, int sh, int channel_step, int channel_off,
                      int16_t *__restrict__ mid, int dw, const int16_t *__restrict__ coef,
                      const int32_t *__restrict__ pos, int taps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (x >= dw || r >= sh) return;
    const uint8_t *s = src + (size_t)r * src_pitch + (size_t)pos[x] * channel_step + channel_off;
    const int16_t *c = coef + (size_t)x * taps;
    int v = 0;
    for (int j = 0; j < taps; j++) v += (int)s[(size_t)j * channel_step] * (int)c[j];
    v >>= 7;
    mid[(size_t)r * dw + x] = (int16_t)min(v, 32767);
}

__global__ void __launch_bounds__(256)
vscale_generic_kernel(const int16_t *__restrict__ mid, int dw, uint8_t *__restrict__ dst, int dst_pitch, int dh,
                      const int16_t *__restrict__ coef, const int32_t *__restrict__ pos, int taps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw || y >= dh) return;
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

int scale_plane_generic(const vt_scale_plan *p, int chroma, const uint8_t *src, int src_pitch, int channel_step,
                        int channel_off, uint8_t *dst, int dst_pitch, cudaStream_t st) {
    const int c = chroma ? 1 : 0;
    const int sh = c ? p->csh : p->sh, dw = c ? p->cdw : p->dw, dh = c ? p->cdh : p->dh;
    dim3 b(256), gh((dw + 255) / 256, sh), gv((dw + 255) / 256, dh);
    hscale_generic_kernel<<<gh, b, 0, st>>>(src, src_pitch, sh, channel_step, channel_off, p->scratch, dw,
                                            p->hcoef[c], p->hpos[c], p->htaps[c]);
    VT_LAUNCHED("hscale_generic_kernel");
    vscale_generic_kernel<<<gv, b, 0, st>>>(p->scratch, dw, dst, dst_pitch, dh, p->vcoef[c], p->vpos[c],
                                            p->vtaps[c]);
    VT_LAUNCHED("vscale_generic_kernel");
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

// ==== streaming path ==============================================================================================
namespace vt {

struct StreamArgs {
    const uint32_t *lane_tab;
    const int32_t *vtab;
    const int32_t *strip_x0;
    uint8_t *dst;            // frame 0, first byte of this plane kind's first plane (Y, or U)
    size_t dst_fs;           // bytes between output frames
    size_t dst_plane2;       // chroma: offset from the U plane to the V plane
    int n_frames, n_strips, n_chunks, rows_out, strip_cols;
    int dw, dh;
    int tile_w, tile_h, warp_smem;
};

__device__ __forceinline__ int dp2a_lo_su(uint32_t coef_pair, uint32_t pix, int acc) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef_pair), "r"(pix), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t coef_pair, uint32_t pix, int acc) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef_pair), "r"(pix), "r"(acc));
    return d;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// One warp per work item (frame, strip of 32*CPT output columns, chunk of rows_out output rows).
//   HP   dp2a pairs per output (horizontal taps padded to 2*HP with zero coefficients)
//   TV   vertical taps (front-padded with zeros), RING = 8 or 16 register slots per column
//   CPT  output columns per lane;  UV: source is the interleaved NV12 chroma plane, a lane produces U and V
//   FULL every strip is completely inside the picture (no per-column bounds checks on the stores)
template <int HP, int TV, int CPT, bool UV, bool FULL>
__global__ void __launch_bounds__(128, 5)
scale_stream_kernel(const __grid_constant__ CUtensorMap tmap, const StreamArgs a) {
    constexpr int RING = TV <= 8 ? 8 : 16;
    constexpr int NCH = UV ? 2 : 1;
    constexpr int NA = UV ? HP : (HP + 1) / 2;   // aligned 32-bit words holding one column's taps
    constexpr int NW = NA + 1;                    // words fetched (one extra for the byte misalignment)
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *wbase = smem + (size_t)warp * a.warp_smem;
    const uint32_t tile_s = smem_u32(wbase);
    const int tile_bytes = a.tile_w * a.tile_h;
    int32_t *vt = (int32_t *)(wbase + ((tile_bytes + 32 + 15) & ~15));
    uint64_t *bar = (uint64_t *)(wbase + a.warp_smem - 16);
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncwarp();
    uint32_t phase = 0;
    const int per_frame = a.n_strips * a.n_chunks;
    const long long total = (long long)per_frame * a.n_frames;
    for (long long it = (long long)blockIdx.x * 4 + warp; it < total; it += (long long)gridDim.x * 4) {
        const int f = (int)(it / per_frame);
        const int rem = (int)(it - (long long)f * per_frame);
        const int chunk = rem / a.n_strips, strip = rem - chunk * a.n_strips;
        const int y0 = chunk * a.rows_out, y1 = min(a.dh, y0 + a.rows_out);
        const int rs = a.vtab[(size_t)y0 * (TV + 2)];                    // first source row of the chunk
        const int re = a.vtab[(size_t)(y1 - 1) * (TV + 2) + 1] + 1;      // one past the last source row
        const int x0s = a.strip_x0[strip];
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)tile_bytes);
            tma_load_3d(wbase, &tmap, bar, x0s, rs, f);
        }
        // per-row vertical table of this chunk -> shared (broadcast reads later)
        const int nvt = (y1 - y0) * (TV + 2);
        for (int i = lane; i < nvt; i += 32) vt[i] = a.vtab[(size_t)y0 * (TV + 2) + i];
        // this lane's columns
        // Lane l owns output columns xl + 32*c: neighbouring lanes read neighbouring source bytes, so a warp's
        // shared-memory loads fall in one or two 128-byte wavefronts without bank conflicts.
        const int xl = strip * a.strip_cols + lane;
        uint32_t cpair[CPT][HP];
        uint32_t addr[CPT], shft[CPT];
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const uint32_t *t = a.lane_tab + (size_t)(xl + 32 * c) * (1 + HP);
            const uint32_t off = t[0] - (uint32_t)x0s;                   // byte offset of tap 0 in a tile row
            addr[c] = tile_s + (off & ~3u);
            shft[c] = (off & 3u) * 8u;
#pragma unroll
            for (int p = 0; p < HP; p++) cpair[c][p] = t[1 + p];
        }
        int m[RING][CPT * NCH];
#pragma unroll
        for (int k = 0; k < RING; k++)
#pragma unroll
            for (int c = 0; c < CPT * NCH; c++) m[k][c] = 0;
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1;

        int ynext = y0;
        int vlast = vt[1] - rs;                                          // row (relative) completing output ynext
        const int nrows = re - rs;
        // running pointers: this lane's first output byte of row ynext, and that row's vertical coefficients
        uint8_t *dptr = a.dst + (size_t)f * a.dst_fs + (size_t)y0 * a.dw + xl;
        const int32_t *vc = vt + 2;
        for (int base = 0; base < nrows; base += RING) {
#pragma unroll
            for (int k = 0; k < RING; k++) {
                const int r = base + k;
                if (r < nrows) {
                    const uint32_t rowoff = (uint32_t)r * (uint32_t)a.tile_w;
                    // ---- horizontal pass: CPT columns (x NCH channels) of source row r -> ring slot k
                    uint32_t w[CPT][NW];
#pragma unroll
                    for (int c = 0; c < CPT; c++)
#pragma unroll
                        for (int i = 0; i < NW; i++) w[c][i] = lds32(addr[c] + rowoff + 4u * i);
#pragma unroll
                    for (int c = 0; c < CPT; c++) {
                        uint32_t al[NA];
#pragma unroll
                        for (int i = 0; i < NA; i++) al[i] = __funnelshift_r(w[c][i], w[c][i + 1], shft[c]);
                        if (!UV) {
                            int v = 0;
#pragma unroll
                            for (int p = 0; p < HP; p++)
                                v = (p & 1) ? dp2a_hi_su(cpair[c][p], al[p >> 1], v) : dp2a_lo_su(cpair[c][p], al[p >> 1], v);
                            m[k][c] = min(v >> 7, 32767);
                        } else {
                            int vu = 0, vv = 0;
#pragma unroll
                            for (int q = 0; q < (HP + 1) / 2; q++) {
                                const uint32_t hi = (2 * q + 1 < NA) ? al[2 * q + 1] : al[2 * q];
                                const uint32_t uw = __byte_perm(al[2 * q], hi, 0x6420);
                                const uint32_t vw = __byte_perm(al[2 * q], hi, 0x7531);
                                vu = dp2a_lo_su(cpair[c][2 * q], uw, vu);
                                vv = dp2a_lo_su(cpair[c][2 * q], vw, vv);
                                if (2 * q + 1 < HP) {
                                    vu = dp2a_hi_su(cpair[c][2 * q + 1], uw, vu);
                                    vv = dp2a_hi_su(cpair[c][2 * q + 1], vw, vv);
                                }
                            }
                            m[k][2 * c] = min(vu >> 7, 32767);
                            m[k][2 * c + 1] = min(vv >> 7, 32767);
                        }
                    }
                    // ---- vertical pass for every output row whose window ends at this source row
                    while (vlast == r) {
                        int acc[CPT * NCH];
#pragma unroll
                        for (int c = 0; c < CPT * NCH; c++) acc[c] = 1 << 18;
#pragma unroll
                        for (int j = 0; j < TV; j++) {
                            const int cj = vc[j];
#pragma unroll
                            for (int c = 0; c < CPT * NCH; c++)
                                acc[c] += m[(k + RING - (TV - 1) + j) & (RING - 1)][c] * cj;
                        }
                        if (!UV) {
#pragma unroll
                            for (int c = 0; c < CPT; c++)
                                if (FULL || xl + 32 * c < a.dw) dptr[32 * c] = (uint8_t)__vimin_s32_relu(acc[c] >> 19, 255);
                        } else {
                            uint8_t *dv = dptr + a.dst_plane2;
#pragma unroll
                            for (int c = 0; c < CPT; c++)
                                if (FULL || xl + 32 * c < a.dw) {
                                    dptr[32 * c] = (uint8_t)__vimin_s32_relu(acc[2 * c] >> 19, 255);
                                    dv[32 * c] = (uint8_t)__vimin_s32_relu(acc[2 * c + 1] >> 19, 255);
                                }
                        }
                        ynext++;
                        dptr += a.dw;
                        vc += TV + 2;
                        vlast = ynext < y1 ? vc[-1] - rs : -1;           // -1: no further output in this chunk
                    }
                }
            }
        }
        __syncwarp();   // every lane is done with the tile before lane 0 lets TMA overwrite it
    }
}

}  // namespace vt

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

namespace {

int make_tmap(CUtensorMap *m, const uint8_t *base, int row_bytes, int rows, int n_frames, int pitch, size_t frame_stride,
              int tile_w, int tile_h) {
    return vt::make_tmap_u8_3d(m, base, row_bytes, rows, n_frames, pitch, frame_stride, tile_w, tile_h);
}

int pad_hp(int taps) {
    const int hp = (taps + 1) / 2;
    return hp <= 3 ? 3 : (hp <= 4 ? 4 : (hp <= 6 ? 6 : 0));
}
int pad_tv(int taps) { return taps < 2 ? 0 : (taps <= 6 ? 6 : (taps <= 8 ? 8 : (taps <= 12 ? 12 : 0))); }

// Builds the tables the streaming kernel reads for one plane kind (c = 0 luma, 1 chroma).
int build_stream(vt_scale_plan *p, int c) {
    vt_scale_plan::Stream &s = p->stream[c];
    const bool uv = c == 1;
    const int dw = c ? p->cdw : p->dw, dh = c ? p->cdh : p->dh, sh = c ? p->csh : p->sh;
    const int bpp = uv ? 2 : 1;                      // source bytes per sample step
    const int ht = p->htaps[c], vtaps = p->vtaps[c];
    s.hp = pad_hp(ht);
    s.tv = pad_tv(vtaps);
    if (!s.hp || !s.tv || dw % 8 || (p->dw % 16)) return VT_OK;        // not eligible: generic path stays
    const std::vector<int32_t> &hpos = p->h_hpos[c], &vpos = p->h_vpos[c];
    const std::vector<int16_t> &hco = p->h_hcoef[c], &vco = p->h_vcoef[c];
    // columns per lane: widest strip whose source span fits one 256-byte TMA box
    int cpt = uv ? 2 : 4;
    for (;; cpt >>= 1) {
        const int cols = 32 * cpt;
        int worst = 0;
        for (int x0 = 0; x0 < dw; x0 += cols) {
            const int x1 = std::min(dw, x0 + cols) - 1;
            const int first = (hpos[x0] * bpp) & ~15;
            const int last = (hpos[x1] + 2 * s.hp) * bpp + 4;           // +4: the extra word each column fetches
            worst = std::max(worst, last - first);
        }
        if (worst <= 256) { s.tile_w = (worst + 15) & ~15; break; }
        if (cpt == 1) return VT_OK;
    }
    s.cpt = cpt;
    s.strip_cols = 32 * cpt;
    s.n_strips = (dw + s.strip_cols - 1) / s.strip_cols;
    // output rows per item: keep the staged tile near 12 KB
    int rows_out = 48;
    for (;; rows_out = rows_out > 24 ? 24 : rows_out >> 1) {
        if (rows_out < 1) return VT_OK;
        int worst = 0;
        for (int y0 = 0; y0 < dh; y0 += rows_out) {
            const int y1 = std::min(dh, y0 + rows_out) - 1;
            worst = std::max(worst, vpos[y1] + vtaps - vpos[y0]);
        }
        if (worst <= 256 && (worst * s.tile_w <= 9216 || rows_out <= 3)) { s.tile_h = worst; break; }
    }
    s.rows_out = rows_out;
    s.n_chunks = (dh + rows_out - 1) / rows_out;
    const int vt_bytes = rows_out * (s.tv + 2) * 4;
    s.warp_smem = ((s.tile_w * s.tile_h + 32 + 15) & ~15) + vt_bytes + 16;
    s.warp_smem = (s.warp_smem + 127) & ~127;
    if (s.warp_smem * 4 > 200 * 1024) return VT_OK;

    std::vector<int32_t> sx((size_t)s.n_strips);
    for (int i = 0; i < s.n_strips; i++) sx[i] = (hpos[i * s.strip_cols] * bpp) & ~15;
    const int ncol = s.n_strips * s.strip_cols;
    std::vector<uint32_t> lt((size_t)ncol * (1 + s.hp), 0);
    for (int x = 0; x < ncol; x++) {
        const int xs = std::min(x, dw - 1);
        uint32_t *t = &lt[(size_t)x * (1 + s.hp)];
        t[0] = (uint32_t)(hpos[xs] * bpp);
        if (x < dw)
            for (int j = 0; j < ht; j++) {
                const uint32_t v = (uint16_t)hco[(size_t)x * ht + j];
                t[1 + j / 2] |= (j & 1) ? (v << 16) : v;
            }
    }
    std::vector<int32_t> vtb((size_t)dh * (s.tv + 2), 0);
    for (int y = 0; y < dh; y++) {
        int32_t *t = &vtb[(size_t)y * (s.tv + 2)];
        t[0] = vpos[y];
        t[1] = vpos[y] + vtaps - 1;
        for (int j = 0; j < vtaps; j++) t[2 + (s.tv - vtaps) + j] = vco[(size_t)y * vtaps + j];
    }
    (void)sh;
    int rc = upload(sx.data(), sx.size() * 4, (void **)&s.strip_x0);
    if (rc == VT_OK) rc = upload(lt.data(), lt.size() * 4, (void **)&s.lane_tab);
    if (rc == VT_OK) rc = upload(vtb.data(), vtb.size() * 4, (void **)&s.vtab);
    if (rc != VT_OK) return rc;
    s.ok = true;
    return VT_OK;
}

template <int HP, int TV, int CPT, bool UV, bool FULL>
int launch_full(const CUtensorMap &tm, const vt::StreamArgs &a, int smem, int grid, cudaStream_t st) {
    auto k = vt::scale_stream_kernel<HP, TV, CPT, UV, FULL>;
    static int smem_set = 0;
    if (smem > smem_set) {
        VT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    k<<<grid, 128, smem, st>>>(tm, a);
    VT_LAUNCHED("scale_stream_kernel");
    return VT_OK;
}

template <int HP, int TV, int CPT, bool UV>
int launch_one(const CUtensorMap &tm, const vt::StreamArgs &a, int smem, int grid, cudaStream_t st) {
    if (a.dw % a.strip_cols == 0) return launch_full<HP, TV, CPT, UV, true>(tm, a, smem, grid, st);
    return launch_full<HP, TV, CPT, UV, false>(tm, a, smem, grid, st);
}

template <int CPT, bool UV>
int dispatch_taps(int hp, int tv, const CUtensorMap &tm, const vt::StreamArgs &a, int smem, int grid, cudaStream_t st) {
#define VT_CASE(H, T) if (hp == H && tv == T) return launch_one<H, T, CPT, UV>(tm, a, smem, grid, st)
    VT_CASE(3, 6); VT_CASE(3, 8); VT_CASE(3, 12);
    VT_CASE(4, 6); VT_CASE(4, 8); VT_CASE(4, 12);
    VT_CASE(6, 6); VT_CASE(6, 8); VT_CASE(6, 12);
#undef VT_CASE
    vt::set_error("scale_stream: no instantiation for hp=%d tv=%d", hp, tv);
    return VT_ERR_UNSUPPORTED;
}

int launch_stream(const vt_scale_plan *p, int c, const uint8_t *src, int pitch, size_t src_fs, uint8_t *dst,
                  size_t dst_fs, int n_frames, cudaStream_t st) {
    const vt_scale_plan::Stream &s = p->stream[c];
    const bool uv = c == 1;
    CUtensorMap tm;
    const uint8_t *base = uv ? src + (size_t)pitch * p->sh : src;
    const int row_bytes = uv ? 2 * p->csw : p->sw;
    const int rows = uv ? p->csh : p->sh;
    int rc = make_tmap(&tm, base, row_bytes, rows, n_frames, pitch, src_fs, s.tile_w, s.tile_h);
    if (rc) return rc;
    vt::StreamArgs a;
    a.lane_tab = s.lane_tab; a.vtab = s.vtab; a.strip_x0 = s.strip_x0;
    a.dst = uv ? dst + (size_t)p->dw * p->dh : dst;
    a.dst_fs = dst_fs;
    a.dst_plane2 = (size_t)p->cdw * p->cdh;
    a.n_frames = n_frames; a.n_strips = s.n_strips; a.n_chunks = s.n_chunks; a.rows_out = s.rows_out;
    a.strip_cols = s.strip_cols;
    a.dw = uv ? p->cdw : p->dw; a.dh = uv ? p->cdh : p->dh;
    a.tile_w = s.tile_w; a.tile_h = s.tile_h; a.warp_smem = s.warp_smem;
    const int smem = s.warp_smem * 4;
    const long long items = (long long)s.n_strips * s.n_chunks * n_frames;
    const int per_sm = std::max(1, std::min(5, (220 * 1024) / (smem + 1024)));
    const long long want = (items + 3) / 4;
    const int grid = (int)std::min<long long>(want, (long long)vt::sm_count() * per_sm);
    if (uv) {
        if (s.cpt == 2) return dispatch_taps<2, true>(s.hp, s.tv, tm, a, smem, grid, st);
        if (s.cpt == 1) return dispatch_taps<1, true>(s.hp, s.tv, tm, a, smem, grid, st);
    } else {
        if (s.cpt == 4) return dispatch_taps<4, false>(s.hp, s.tv, tm, a, smem, grid, st);
        if (s.cpt == 2) return dispatch_taps<2, false>(s.hp, s.tv, tm, a, smem, grid, st);
        if (s.cpt == 1) return dispatch_taps<1, false>(s.hp, s.tv, tm, a, smem, grid, st);
    }
    vt::set_error("scale_stream: unsupported cpt=%d", s.cpt);
    return VT_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int vt_scale_plan_stream_info(const vt_scale_plan *p, int chroma, int *info8) {
    if (!p || !info8) return VT_ERR_INVALID;
    const vt_scale_plan::Pair &pr = p->pair[chroma ? 1 : 0];
    if (p->pair[0].ok && p->pair[1].ok) {
        info8[0] = 1; info8[1] = pr.hp; info8[2] = pr.tv; info8[3] = 2 * pr.np; info8[4] = pr.stage_rows;
        info8[5] = pr.tile_w; info8[6] = pr.stage_rows * pr.n_stages; info8[7] = pr.warp_smem;
        return VT_OK;
    }
    const vt_scale_plan::Stream &s = p->stream[chroma ? 1 : 0];
    info8[0] = s.ok; info8[1] = s.hp; info8[2] = s.tv; info8[3] = s.cpt; info8[4] = s.rows_out; info8[5] = s.tile_w;
    info8[6] = s.tile_h; info8[7] = s.warp_smem;
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
    for (int c = 0; c < 2; c++) { p->hcoef[c] = p->vcoef[c] = nullptr; p->hpos[c] = p->vpos[c] = nullptr; }
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
    if (rc == VT_OK && cudaMalloc((void **)&p->scratch, (size_t)dw * sh * sizeof(int16_t)) != cudaSuccess)
        rc = VT_ERR_NOMEM;
    for (int c = 0; c < 2 && rc == VT_OK; c++) rc = build_stream(p, c);
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
        cudaFree(p->stream[c].strip_x0); cudaFree(p->stream[c].lane_tab); cudaFree(p->stream[c].vtab);
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
    return vt::scale_plane_generic(plan, chroma, src, src_pitch, 1, 0, dst, dst_pitch, (cudaStream_t)stream);
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
    // VT_SCALE_KERNEL=stream|generic selects the older kernels (A/B measurements only)
    static const char *force = getenv("VT_SCALE_KERNEL");
    const bool want_pair = !force || !strcmp(force, "pair");
    const bool want_stream = !force || !strcmp(force, "stream");
    if (aligned && want_pair && p->pair[0].ok && p->pair[1].ok) {
        int rc = vt::launch_pair(p, 0, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
        if (rc) return rc;
        return vt::launch_pair(p, 1, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
    }
    if (aligned && want_stream && p->stream[0].ok && p->stream[1].ok) {
        int rc = launch_stream(p, 0, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
        if (rc) return rc;
        return launch_stream(p, 1, src, src_pitch, src_fs, dst, dst_fs, n_frames, st);
    }
    for (int f = 0; f < n_frames; f++) {
        const uint8_t *s = src + (size_t)f * src_fs;
        uint8_t *d = dst + (size_t)f * dst_fs;
        int rc = vt::scale_plane_generic(p, 0, s, src_pitch, 1, 0, d, p->dw, st);
        if (rc) return rc;
        const uint8_t *uv = s + (size_t)src_pitch * p->sh;
        rc = vt::scale_plane_generic(p, 1, uv, src_pitch, 2, 0, d + ysz, p->cdw, st);
        if (rc) return rc;
        rc = vt::scale_plane_generic(p, 1, uv, src_pitch, 2, 1, d + ysz + csz, p->cdw, st);
        if (rc) return rc;
    }
    return VT_OK;
}

