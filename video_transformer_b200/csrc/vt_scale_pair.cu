// vt_scale_pair.cu -- K2 production kernel: libswscale-exact separable polyphase downscale of NV12 -> YUV420P
// (arithmetic stated in vt_scale.cu).  The reference reaches this work
// through `ffmpeg -vf scale=-2:360` (/root/reference/src/analyzer/content_analyzer.py:193-211).
//
// Shape of the kernel (everything below is about instruction count: the pass is issue-bound, not DRAM-bound):
//   * one warp owns a strip of output columns and walks DOWN a segment of output rows; its source rows arrive
//     through a private TMA ring (cp.async.bulk.tensor 3-D boxes, one mbarrier per stage), so every source
//     byte is read from HBM once and there is no block-level synchronisation at all;
//   * a lane owns PAIRS of adjacent output columns (x, x+1).  The even column's taps are aligned with funnel
//     shifts and run through dp2a (two 14-bit coefficients x two pixels per instruction).  The odd column
//     starts 0..3 source samples further right, so it re-uses the SAME aligned words with its coefficients
//     rotated into place (one more dp2a, no loads, no shifts);
//   * the 15-bit horizontal results live in a register ring of TV rows (slot = source row mod TV, static
//     because the row loop is unrolled by TV); whenever a source row completes an output row's window the
//     vertical taps run out of that ring.  Vertical coefficients and the "window ends at row" schedule sit in
//     KERNEL PARAMETER space (constant bank), are fetched with uniform loads and feed IMAD directly as uniform
//     register operands, so the vertical pass costs no shared-memory traffic and its branch is warp-uniform;
//   * results are clamped and packed with cvt.pack.sat (I2IP) and leave as 16-bit stores;
//   * a ragged right edge is handled by sliding the last strip left until it ends at the last column (the
//     overlap is computed twice with identical results), so there is no per-store bounds predicate.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cstring>
#include <mutex>
#include <vector>

#include "vt_scale_plan.cuh"

// VT_ABLATE (measurement builds only, results are wrong): 1 no vertical taps, 2 no horizontal taps, 4 no tile loads,
// 8 no stores
#ifndef VT_ABLATE
#define VT_ABLATE 0
#endif
// VT_PAIR_TIMES (measurement builds, tools/build_variant.sh times -DVT_PAIR_TIMES=1): with VT_PAIR_DEBUG_TIMES=file every
// warp records its start and end time (tools/pair_times.py reads the file).  What it showed: the SM's schedulers favour
// the oldest warp -- with equal static shares the five warps of a scheduler finish at 0.37, 0.51, 0.68, 0.84 and 0.96 of
// the kernel's duration -- and that this costs nothing: handing the work out dynamically (ticketed, shrinking chunks;
// all warps then end within 5 % of each other) made the kernel 10 % SLOWER, because every extra item pays a ring
// warm-up while one or two warps per scheduler already keep the multiply pipe as busy as five do.
#ifndef VT_PAIR_TIMES
#define VT_PAIR_TIMES 0
#endif

namespace vt {

template <int TV>
struct VCfg {
    static constexpr int STRIDE = (TV + 1 + 3) & ~3;  // coefficients[TV] (front padded), last source row, padding
    static constexpr int ROWS = 28672 / (STRIDE * 2); // output rows one launch can cover (parameter space is 32 KB)
};
// 16-bit entries (coefficients are 13-bit, source rows < 32768): 720 rows of a 12-tap plan or 1080 rows of an 8-tap plan
// fit one launch
template <int TV>
struct __align__(16) VTab {
    int16_t t[VCfg<TV>::ROWS * VCfg<TV>::STRIDE];
};

struct PairArgs {
    const uint32_t *lane_tab;
    const int32_t *box_x0;             // n_strips x n_boxes: first source byte of each TMA box (multiple of 16)
    const int32_t *strip_col;          // n_strips: first output column of each strip
    uint8_t *dst;                      // frame 0: first byte of the Y plane (luma) or of the U plane (chroma)
    unsigned long long dst_fs;         // bytes between output frames
    unsigned long long dst_plane2;     // chroma: U plane -> V plane
    int n_frames, n_strips, n_segs, seg_rows;
    unsigned int dw;                   // output width of this plane kind
    int y_begin, y_end;                // output rows covered by this launch (vtab row 0 = y_begin)
    int tile_w;                        // bytes per box row (always PAIR_TILE_W)
    int n_boxes;                       // TMA boxes per stage
    int groups_per_stage;              // a stage holds groups_per_stage * TV source rows
    int box_bytes, stage_bytes, n_stages, warp_smem;
    int round_bias;                    // 1 << 18 (the vertical pass's rounding term)
    // static vertical schedule (kernels instantiated with MASK != 0): output rows [reg_lo, reg_hi) repeat one pattern
    // of window ends and coefficient sets every align_p source rows; segments start on rows = align_r0 (mod align_p)
    int reg_lo, reg_hi, align_p, align_r0;
    unsigned long long *dbg_times;     // VT_PAIR_TIMES builds: per warp {start, end} globaltimer (null otherwise)
    int sc[24];                        // [phase][TV] front-padded coefficients of the pattern's output rows
    // fused scene score (kernels instantiated with SC = true): SAD and 256-bin histogram of the SOURCE luma, taken from
    // the rows the scaler already has in shared memory.  Every source pixel is counted by exactly one warp: a lane
    // owns the 12 source bytes under its 8 output columns, an item owns the source rows from the end of the previous
    // item's last window to the end of its own.
    const uint8_t *sc_src;             // frame 0's luma plane (the surface the tensor map describes)
    const uint8_t *sc_prev0;           // luma of the picture preceding frame 0, or null (then frame 0's SAD is 0)
    unsigned long long sc_fs;          // bytes between frames
    int sc_pitch, sc_rows;             // source pitch and rows
    unsigned long long *sc_sad;        // [n_frames], zeroed before the launch
    uint32_t *sc_hist;                 // [n_frames][256], zeroed before the launch
    int sc_cnt_off;                    // byte offset of the warp's counters inside its shared memory
};

constexpr int PAIR_SC_COUNTER_BYTES = 128 * 32 * 4;   // per warp: 128 rows x 32 lanes of packed u16 pairs (bins r, r + 128)

__device__ __forceinline__ int dp2a_lo(uint32_t coef_pair, uint32_t pix, int acc) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef_pair), "r"(pix), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi(uint32_t coef_pair, uint32_t pix, int acc) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef_pair), "r"(pix), "r"(acc));
    return d;
}
// sum over i < N of coefficient pair i times halfword i of the byte string w[]
template <int N>
__device__ __forceinline__ int dot_halfwords(const uint32_t (&c)[N], const uint32_t *w) {
    int v = 0;
#pragma unroll
    for (int i = 0; i < N; i++) v = (i & 1) ? dp2a_hi(c[i], w[i >> 1], v) : dp2a_lo(c[i], w[i >> 1], v);
    return v;
}
// The odd column of a pair: its taps start 0..2 samples right of the even column's.  Halfwords 1..HP-1 of the aligned
// samples carry HP-1 coefficient pairs in place; the (at most two) samples that fall outside them are fetched into
// one halfword by a byte permute (ALU pipe) and cost one more dp2a -- HP dp2a in all, like the even column.
template <int HP>
__device__ __forceinline__ int dot_odd(const uint32_t (&c)[HP], const uint32_t *w, uint32_t w_last, uint32_t sel) {
    int v = dp2a_lo(c[HP - 1], __byte_perm(w[0], w_last, sel), 0);
#pragma unroll
    for (int i = 1; i < HP; i++) v = (i & 1) ? dp2a_hi(c[i - 1], w[i >> 1], v) : dp2a_lo(c[i - 1], w[i >> 1], v);
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// (sat_u8(hi) << 8) | sat_u8(lo)
__device__ __forceinline__ uint32_t pack_sat_u8x2(int hi, int lo) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(0u));
    return d;
}
// sat_u8 of four results, b0 in the low byte (two I2IP: the second takes the first as its upper half)
__device__ __forceinline__ uint32_t pack_sat_u8x4(int b3, int b2, int b1, int b0) {
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b3), "r"(b2), "r"(0u));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b1), "r"(b0), "r"(t));
    return d;
}
__device__ __forceinline__ void st_u32(uint8_t *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_u32x2(uint8_t *p, uint32_t lo, uint32_t hi) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(lo), "r"(hi) : "memory");
}
// A column pair of a plane whose width is odd (427-wide chroma of 854x480): its rows alternate between even and odd byte
// addresses, so the pair leaves as two byte stores.  The strips still end exactly on the last column (the last strip
// slides left), so no column of a pair is out of range.
__device__ __forceinline__ void st_u8x2(uint8_t *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u8 [%0], %1;" ::"l"(p), "r"(v & 0xFFu) : "memory");
    asm volatile("st.global.L1::no_allocate.u8 [%0], %1;" ::"l"(p + 1), "r"((v >> 8) & 0xFFu) : "memory");
}
__device__ __forceinline__ void st_u16(uint8_t *p, uint32_t v) {
    if (VT_ABLATE & 8) {
        if (v == 0x12345u) asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
        return;
    }
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
}

__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(phase)
        : "memory");
}

#ifndef VT_PAIR_WIDE
#define VT_PAIR_WIDE 1                   // 0: measurement build with two column pairs per lane everywhere
#endif
#ifndef VT_PAIR_BLOCKS
#define VT_PAIR_BLOCKS 5
#endif
constexpr uint32_t PAIR_TILE_W = VT_PAIR_WIDE ? 448 : 256;   // bytes per row of every TMA box (32-bit elements: up to 1024 B)

__host__ __device__ constexpr int pair_np(int hp, int tv) { return (VT_PAIR_WIDE && hp <= 3 && tv <= 8) ? 4 : 2; }   // luma pairs per lane
#ifndef VT_PAIR_SG8
#define VT_PAIR_SG8 2                    // groups per stage / static block for 8 vertical taps: 2 (16-row boxes, three
                                         // blocks per SM) measured 8-12 % faster than 1 (8-row boxes, five blocks)
#endif
#ifndef VT_PAIR_SG12
#define VT_PAIR_SG12 1                   // the same for 12 vertical taps: 2 (24-row boxes, two blocks per SM) measured 5-8 % slower
#endif
__host__ __device__ constexpr int pair_sg(int tv) { return tv <= 6 ? 2 : (tv <= 8 ? VT_PAIR_SG8 : VT_PAIR_SG12); }
__host__ __device__ constexpr int pair_min_blocks(int hp, int tv) {
    return (hp <= 4 && tv <= 6) ? VT_PAIR_BLOCKS : (hp <= 4 && tv <= 8) ? (VT_PAIR_SG8 == 2 ? 3 : VT_PAIR_BLOCKS)
                                                 : (tv > 8 && VT_PAIR_SG12 == 2 ? 2 : 3);
}

// compile-time loop: f(std::integral_constant<int, K>) for K = I .. N-1 (the row index must be a constant expression
// so that ring slots, mask bits and coefficient offsets fold into the instructions)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr int popc_c(unsigned v) { return v ? (int)(v & 1u) + popc_c(v >> 1) : 0; }
__host__ __device__ constexpr int ctz_c(unsigned v) { return (v & 1u) ? 0 : 1 + ctz_c(v >> 1); }

// ---- static horizontal patterns (the kernel's HS parameter): 1 = exact 3:2, 2 = exact 2:1, 3 = exact 3:1 --------------
// First source sample of output column x's nominal tap window (libswscale's interior positions for these ratios).
__host__ __device__ constexpr int hs_nominal(int hs, int x) {
    return hs == 1 ? 3 * (x >> 1) - 2 + (x & 1) : hs == 2 ? 2 * x - 3 : 3 * x - 4;
}
// Samples between the lane's (4-byte aligned) window start and its first column's nominal window.
__host__ __device__ constexpr int hs_pre(int hs, bool uv) { return uv ? (hs == 2 ? 1 : 0) : (hs == 1 ? 2 : hs == 2 ? 1 : 0); }
// Offset (samples) of column c's first tap inside the lane's window.
__host__ __device__ constexpr int hs_off(int hs, bool uv, int c) { return hs_nominal(hs, c) - hs_nominal(hs, 0) + hs_pre(hs, uv); }
// Source bytes a lane advances per lane index: its columns x ratio (x 2 bytes per chroma sample).
__host__ __device__ constexpr int hs_lane_stride(int hs, bool uv, int ncol) {
    return (hs == 1 ? 3 * ncol / 2 : hs == 2 ? 2 * ncol : 3 * ncol) * (uv ? 2 : 1);
}
// 32-bit words a lane fetches per source row.
__host__ __device__ constexpr int hs_words(int hs, bool uv, int ncol, int hp) {
    int hi = 0;
    for (int c = 0; c < ncol; c++) {
        const int off = hs_off(hs, uv, c);
        int last = 0;
        if (uv) {
            const int sp = off + 2 * (hp - 1);                  // last sample pair (sp, sp+1): bytes 2sp .. 2sp+3
            last = (sp & 1) ? (sp >> 1) + 1 : (sp >> 1);
        } else {
            const int hw = (off >> 1) + hp - 1;                 // last halfword of the (byte-shifted, if off is odd) stream
            last = (off & 1) ? (hw >> 1) + 1 : (hw >> 1);
        }
        hi = last > hi ? last : hi;
    }
    return hi + 1;
}

// HP  dp2a pairs of the even column (taps padded to 2*HP); the odd column uses HP+1 pairs (rotated coefficients)
// TV  vertical taps (front padded); UV: the source is NV12's interleaved chroma plane, a lane produces U and V
// MASK, Q  static vertical schedule for ratios whose vertical phases repeat inside a group of TV source rows
//     (3:2, 2:1, 3:1): bit k of MASK = "an output row's window ends at row k of every regular group", Q = number of
//     distinct coefficient sets.  Regular groups then run without any per-row branch or table fetch (coefficients
//     are constant-bank operands of the IMADs); picture edges and MASK == 0 plans take the table-driven path.
// HS  static horizontal pattern (1: exact 3:2 with HP 3, 2: exact 2:1 with HP 4, 3: exact 3:1 with HP 6): a lane owns
//     2*NP ADJACENT output columns, i.e. a fixed step of the source per lane.  Column c's first tap then sits at a
//     fixed offset of the lane's window (3:2 luma: bytes 2,3,5,6,8,9,11,12; chroma: samples 0,1,3,4), so the
//     alignment of every tap pair is a compile-time choice
//     between the loaded words and ONE byte-shifted copy of them (luma), or one byte permute per sample pair that
//     serves U (dp2a.lo) and V (dp2a.hi) at once (chroma).  Five loads per source row per lane instead of twelve,
//     no per-pair shifts or selectors, and the lane's output bytes leave as one 8-byte (2 x 4-byte) store.
template <int HP, int TV, bool UV, int MASK, int Q, int HS, bool SC = false>
__global__ void __launch_bounds__(128, SC ? 2 : pair_min_blocks(HP, TV))
scale_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PairArgs a,
                  const __grid_constant__ VTab<TV> vtab) {
    constexpr int BN = HP + 1;
    constexpr int NP = UV ? pair_np(HP, TV) / 2 : pair_np(HP, TV);   // column pairs per lane
    constexpr int NCH = UV ? 2 : 1;
    constexpr int NM = NP * 2 * NCH;                  // horizontal results per lane per source row
    constexpr int NAW = UV ? BN : (BN + 1) / 2;       // aligned 32-bit words holding a pair's source samples
    constexpr int NW = NAW + 1;                       // words fetched (one extra for the byte misalignment)
    constexpr int NQ = (NAW + 1) / 2;                 // chroma: de-interleaved words per channel
    constexpr int LT = (2 * HP + 2 + 3) & ~3;         // words per lane-table entry: offset, HP + HP pairs, selector
    constexpr int VS = VCfg<TV>::STRIDE;
    constexpr int NOUT = popc_c((unsigned)MASK);      // output rows per regular group
    constexpr int FIRSTK = MASK ? ctz_c((unsigned)MASK) : 0;
    constexpr int SG = pair_sg(TV);                   // groups per static block (= groups per TMA stage)
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform for the compiler
    uint8_t *wbase = smem + (size_t)warp * a.warp_smem;
    const uint32_t wsm = smem_u32(wbase);
    const int nst = a.n_stages;
    const int rg = a.groups_per_stage;
    const int stage_rows = rg * TV;
    uint64_t *bars = (uint64_t *)(wbase + (size_t)nst * a.stage_bytes);
    if (lane == 0) {
        for (int s = 0; s < nst; s++) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t cnt_lane = wsm + (uint32_t)a.sc_cnt_off + 4u * (uint32_t)lane;   // this lane's column of counters
    if constexpr (SC) {
        static_assert(!SC || (HS == 1 && !UV), "the fused score is written for the 3:2 luma layout");
        const uint32_t z = wsm + (uint32_t)a.sc_cnt_off + 16u * (uint32_t)lane;
#pragma unroll 4
        for (int i = 0; i < PAIR_SC_COUNTER_BYTES / (32 * 16); i++)
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(z + 512u * i), "r"(0u) : "memory");
        __syncwarp();
    }
    const uint32_t bar0 = smem_u32(bars);
    constexpr uint32_t group_bytes = (uint32_t)TV * PAIR_TILE_W;   // tile rows are PAIR_TILE_W bytes apart: row offsets
    uint32_t phases = 0;                                           // inside a group are immediates of the loads
    const uint32_t stage_tx = (uint32_t)(stage_rows * PAIR_TILE_W * a.n_boxes);
    // A warp keeps ONE strip for the whole launch (its column tables are loaded once) and walks over that strip's
    // (frame, segment) items; warps beyond a multiple of n_strips have nothing to do.
    const int warps_per_strip = (int)(((long long)gridDim.x * 4) / a.n_strips);
    const int gw = blockIdx.x * 4 + warp;
    if (gw >= warps_per_strip * a.n_strips) return;
    const int strip = gw % a.n_strips;
    const int total = a.n_frames * a.n_segs;
#if VT_PAIR_TIMES
    if (a.dbg_times && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.dbg_times[2 * gw] = t;
    }
#endif
    // this lane's column pairs
    // per pair: tile address, byte shift, even column's HP coefficient pairs; odd column: HP-1 pairs on the aligned
    // halfwords 1..HP-1, one pair on the two leftover samples, and the byte selector that fetches those two samples
    constexpr int NCOL = 2 * NP;                      // HS: adjacent output columns per lane
    uint32_t addr[NP], shb[NP], ca[NP][HP], cb[NP][HP], selx[NP];
    uint32_t hc[HS ? NCOL : 1][HP];                   // HS: coefficient pairs of column c, placed on its nominal window
    int hs_xa = 0;
    if constexpr (HS) {
        static_assert(!HS || (NCOL * HP) % 4 == 0, "lane entries are fetched as 16-byte words");
        const uint4 *t4 = reinterpret_cast<const uint4 *>(a.lane_tab + ((size_t)strip * 32 + lane) * (NCOL * HP));
#pragma unroll
        for (int i = 0; i < NCOL * HP / 4; i++) {
            const uint4 q = __ldg(t4 + i);
            hc[(4 * i) / HP][(4 * i) % HP] = q.x;
            hc[(4 * i + 1) / HP][(4 * i + 1) % HP] = q.y;
            hc[(4 * i + 2) / HP][(4 * i + 2) % HP] = q.z;
            hc[(4 * i + 3) / HP][(4 * i + 3) % HP] = q.w;
        }
        // the strip's source window starts 4 bytes left of lane 0's nominal taps.  A TMA box must start on a 16-byte
        // boundary (measured: 4-byte starts fault) but may start left of column 0 (zero fill), so round down
        const int x0 = a.box_x0[strip];
        hs_xa = x0 & ~15;
        addr[0] = wsm + (uint32_t)hs_lane_stride(HS, UV, NCOL) * (uint32_t)lane + (uint32_t)(x0 - hs_xa);
    }
#pragma unroll
    for (int g = 0; g < (HS ? 0 : NP); g++) {
        const uint4 *t4 = reinterpret_cast<const uint4 *>(a.lane_tab + ((size_t)(strip * NP + g) * 32 + lane) * LT);
        uint32_t wd[LT];
#pragma unroll
        for (int i = 0; i < LT / 4; i++) {
            const uint4 q = __ldg(t4 + i);
            wd[4 * i] = q.x; wd[4 * i + 1] = q.y; wd[4 * i + 2] = q.z; wd[4 * i + 3] = q.w;
        }
        addr[g] = wsm + (wd[0] & ~3u);               // wd[0]: byte offset of the pair's tap 0 inside a stage (row 0)
        shb[g] = (wd[0] & 3u) * 8u;
#pragma unroll
        for (int i = 0; i < HP; i++) ca[g][i] = wd[1 + i];
#pragma unroll
        for (int i = 0; i < HP; i++) cb[g][i] = wd[1 + HP + i];
        selx[g] = wd[1 + 2 * HP];
    }
    const size_t strip_byte = (size_t)a.strip_col[strip] + (HS ? NCOL : 2) * lane;
    // Work units are (frame, segment) pairs in frame-major order; a warp takes a CONTIGUOUS run of them, so vertically
    // adjacent segments of one frame merge into one item (one ring warm-up, one pipeline start) and the unit size only
    // sets the balancing granularity.
    const int per_warp = (total + warps_per_strip - 1) / warps_per_strip;
    int unit = (gw / a.n_strips) * per_warp;
    const int unit_end = min(total, unit + per_warp);
    while (unit < unit_end) {
        const int f = unit / a.n_segs;
        const int seg = unit - f * a.n_segs;
        const int nseg = min(a.n_segs - seg, unit_end - unit);              // segments of this frame in the run
        unit += nseg;
        const int y0 = a.y_begin + seg * a.seg_rows;
        const int y1 = min(a.y_end, y0 + nseg * a.seg_rows);
        int vi = (y0 - a.y_begin) * VS;
        int rs = vtab.t[vi + TV] - (TV - 1);                                // first source row of the segment's windows
        if (MASK) {                                                          // start on the pattern's row phase
            int d = (rs - a.align_r0) % a.align_p;
            rs -= d < 0 ? d + a.align_p : d;
        }
        const int nrows = vtab.t[(y1 - 1 - a.y_begin) * VS + TV] + 1 - rs;  // source rows the segment needs
        const int ngroups = (nrows + TV - 1) / TV;
        const int nloads = (ngroups + rg - 1) / rg;
        __syncwarp();                                                        // previous item's tiles are no longer read
        if (lane == 0) {
            const int pre = min(nst, nloads);
#pragma unroll 1
            for (int s = 0; s < pre; s++) {
                mbar_expect_tx(&bars[s], stage_tx);
#pragma unroll 1
                for (int b = 0; b < a.n_boxes; b++)
                    tma_load_3d(wbase + (size_t)s * a.stage_bytes + (size_t)b * a.box_bytes, &tmap, &bars[s],
                                (HS ? hs_xa : a.box_x0[strip * a.n_boxes + b]) >> 2, rs + s * stage_rows, f);
            }
        }
        int m[TV][NM];
#pragma unroll
        for (int k = 0; k < TV; k++)
#pragma unroll
            for (int c = 0; c < NM; c++) m[k][c] = 0;

        const int rnd = a.round_bias;                                        // 1 << 18, from the parameters so that it lives in a
                                                                             // register and every tap is IMAD acc, m, UR(coef), acc
        int y = y0;
        int vrel = vtab.t[vi + TV] - rs;                                     // row (relative to the current group) completing row y
        const int ylim = min(y1, a.reg_hi);
        bool stale = false;                                                  // vc[] / last_next do not belong to row y
        int vc[TV];
#pragma unroll
        for (int j = 0; j < TV; j++) vc[j] = vtab.t[vi + j];
        int last_next = vtab.t[vi + VS + TV];                                // row y+1's window end, fetched one row ahead so that
                                                                             // the "does a row end here" branch never waits on a load
        uint8_t *dptr = a.dst + (size_t)f * a.dst_fs + (size_t)y0 * a.dw + strip_byte;

        // fused score: rows and bytes this item owns, the previous picture's bytes under them
        int own_lo = 0;
        unsigned own_span = 0;
        uint32_t sad_acc = 0;
        const uint8_t *prev_lane = nullptr;
        if constexpr (SC) {
            own_lo = y0 == a.y_begin ? 0 : vtab.t[(y0 - 1 - a.y_begin) * VS + TV] + 1;
            const int own_hi = y1 >= a.y_end ? a.sc_rows : vtab.t[(y1 - 1 - a.y_begin) * VS + TV] + 1;
            own_span = (unsigned)(own_hi - own_lo);
            // no predecessor: compare the picture with itself (SAD 0 without a branch in the row code)
            const uint8_t *pf = f > 0 ? a.sc_src + (size_t)(f - 1) * a.sc_fs : (a.sc_prev0 ? a.sc_prev0 : a.sc_src);
            prev_lane = pf + (a.box_x0[strip] + 4) + 12 * lane;
        }
        uint32_t ga[NP];
        // ALL:   every row this call may see is owned by the item (the straight-line block must stay branch-free); a
        //        literal at both call sites, so the tests on it fold at compile time once the lambda is inlined
        // pvrow: the previous picture's three words under this lane for that row, already in registers (the static
        //        block requests a whole stage's worth one stage ahead: issued just in time, each row would expose a
        //        memory latency with only two warps per scheduler to hide it)
        auto hpass = [&](int k, int (&out)[NM], int srow, const bool ALL, const uint32_t *pvrow) {   // horizontal pass of row k
            if constexpr (HS != 0) {
                constexpr int NWD = hs_words(HS, UV, NCOL, HP);
                const uint32_t ra = ga[0] + (uint32_t)k * PAIR_TILE_W;
                uint32_t w[NWD];
#pragma unroll
                for (int i = 0; i < NWD; i++) w[i] = lds_u32(ra + 4u * i);
                if constexpr (SC) {
                    // words 1..3 are the 12 source bytes under this lane's 8 output columns
                    if (ALL || (unsigned)(srow - own_lo) < own_span) {      // warp-uniform
                        if (ALL) {
                            sad_acc = sad4(w[1], pvrow[0], sad_acc);
                            sad_acc = sad4(w[2], pvrow[1], sad_acc);
                            sad_acc = sad4(w[3], pvrow[2], sad_acc);
                        } else {
                            const uint32_t *pp = reinterpret_cast<const uint32_t *>(prev_lane + (size_t)srow * a.sc_pitch);
                            sad_acc = sad4(w[1], __ldg(pp), sad_acc);
                            sad_acc = sad4(w[2], __ldg(pp + 1), sad_acc);
                            sad_acc = sad4(w[3], __ldg(pp + 2), sad_acc);
                        }
#pragma unroll
                        for (int i = 1; i <= 3; i++) {
                            const uint32_t lo7 = w[i] & 0x7F7F7F7Fu;        // counter row = bin & 127
                            const uint32_t hi16 = (w[i] >> 3) & 0x10101010u; // shift count of the increment: 16 * (bin >> 7)
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                // address on the multiply pipe (one-hot dp4a: cnt_lane + 128 * byte_b), increment on the ALU.
                                // No "memory" clobber: the counters alias nothing the scaler reads and the flush sits
                                // behind __syncwarp().  (atomicAdd instead of the asm measured 5 % slower.)
                                const uint32_t addr = (uint32_t)__dp4a(lo7, 0x80u << (8 * b), cnt_lane);
                                const uint32_t inc = 1u << __byte_perm(hi16, 0u, 0x4440 + b);
                                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc));
                            }
                        }
                    }
                }
                if constexpr (!UV) {
                    uint32_t s1[NWD - 1];                                    // the same bytes, one byte further on
#pragma unroll
                    for (int i = 0; i < NWD - 1; i++) s1[i] = __funnelshift_r(w[i], w[i + 1], 8);   // (unused ones fold away)
#pragma unroll
                    for (int c = 0; c < NCOL; c++) {
                        const int off = hs_off(HS, false, c), h = off >> 1;
                        int v = 0;
#pragma unroll
                        for (int t = 0; t < HP; t++) {
                            const int hw = h + t;
                            const uint32_t word = (off & 1) ? s1[hw >> 1] : w[hw >> 1];
                            v = (hw & 1) ? dp2a_hi(hc[c][t], word, v) : dp2a_lo(hc[c][t], word, v);
                        }
                        out[c] = min(v >> 7, 32767);
                    }
                } else {
                    constexpr int NPW = hs_off(HS, true, NCOL - 1) + 2 * (HP - 1) + 1;
                    uint32_t pw[NPW];                                        // sample pair s: (U_s, U_s+1, V_s, V_s+1)
#pragma unroll
                    for (int sp = 0; sp < NPW; sp++) {
                        const int wi = sp >> 1;
                        pw[sp] = (sp & 1) ? __byte_perm(w[wi], w[wi + 1 < NWD ? wi + 1 : wi], 0x5342) : __byte_perm(w[wi], 0u, 0x3120);
                    }
#pragma unroll
                    for (int c = 0; c < NCOL; c++) {
                        const int off = hs_off(HS, true, c);
                        int u = 0, v = 0;
#pragma unroll
                        for (int t = 0; t < HP; t++) {
                            u = dp2a_lo(hc[c][t], pw[off + 2 * t], u);
                            v = dp2a_hi(hc[c][t], pw[off + 2 * t], v);
                        }
                        out[4 * (c >> 1) + (c & 1)] = min(u >> 7, 32767);
                        out[4 * (c >> 1) + 2 + (c & 1)] = min(v >> 7, 32767);
                    }
                }
                return;
            }
#pragma unroll
            for (int g = 0; g < NP; g++) {
                uint32_t w[NW], al[NAW];
#pragma unroll
                for (int i = 0; i < NW; i++)
                    w[i] = (VT_ABLATE & 4) ? ga[g] * (uint32_t)(k + i + 1) : lds_u32(ga[g] + (uint32_t)k * PAIR_TILE_W + 4u * i);
#pragma unroll
                for (int i = 0; i < NAW; i++) al[i] = __funnelshift_r(w[i], w[i + 1], shb[g]);
                if (VT_ABLATE & 2) {
#pragma unroll
                    for (int c = 0; c < NM / NP; c++) out[NM / NP * g + c] = (int)(al[c % NAW] ^ ca[g][c % HP]);
                } else if (!UV) {
                    out[2 * g] = min(dot_halfwords<HP>(ca[g], al) >> 7, 32767);
                    out[2 * g + 1] = min(dot_odd<HP>(cb[g], al, al[NAW - 1], selx[g]) >> 7, 32767);
                } else {
                    uint32_t uw[NQ], vw[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        const uint32_t hi = (2 * q + 1 < NAW) ? al[2 * q + 1] : al[2 * q];
                        uw[q] = __byte_perm(al[2 * q], hi, 0x6420);
                        vw[q] = __byte_perm(al[2 * q], hi, 0x7531);
                    }
                    out[4 * g] = min(dot_halfwords<HP>(ca[g], uw) >> 7, 32767);
                    out[4 * g + 1] = min(dot_odd<HP>(cb[g], uw, uw[NQ - 1], selx[g]) >> 7, 32767);
                    out[4 * g + 2] = min(dot_halfwords<HP>(ca[g], vw) >> 7, 32767);
                    out[4 * g + 3] = min(dot_odd<HP>(cb[g], vw, vw[NQ - 1], selx[g]) >> 7, 32767);
                }
            }
        };
        auto vstore = [&](const int (&acc)[NM]) {                            // clamp, pack and store output row y
            if constexpr (HS != 0 && NM == 8) {
                if (!UV) {
                    const uint32_t lo = pack_sat_u8x4(acc[3] >> 19, acc[2] >> 19, acc[1] >> 19, acc[0] >> 19);
                    const uint32_t hi = pack_sat_u8x4(acc[7] >> 19, acc[6] >> 19, acc[5] >> 19, acc[4] >> 19);
                    st_u32x2(dptr, lo, hi);
                } else {
                    st_u32(dptr, pack_sat_u8x4(acc[5] >> 19, acc[4] >> 19, acc[1] >> 19, acc[0] >> 19));
                    st_u32(dptr + a.dst_plane2, pack_sat_u8x4(acc[7] >> 19, acc[6] >> 19, acc[3] >> 19, acc[2] >> 19));
                }
                return;
            }
            if constexpr (HS != 0 && NM == 4 && !UV) {
                st_u32(dptr, pack_sat_u8x4(acc[3] >> 19, acc[2] >> 19, acc[1] >> 19, acc[0] >> 19));
                return;
            }
            // (HS chroma with one column pair per lane stores its two 16-bit halves exactly like the generic layout)
            if (a.dw & 1) {                                                  // odd plane width (warp-uniform): byte stores
#pragma unroll
                for (int g = 0; g < NP; g++) {
                    if (!UV) {
                        st_u8x2(dptr + 64 * g, pack_sat_u8x2(acc[2 * g + 1] >> 19, acc[2 * g] >> 19));
                    } else {
                        st_u8x2(dptr + 64 * g, pack_sat_u8x2(acc[4 * g + 1] >> 19, acc[4 * g] >> 19));
                        st_u8x2(dptr + a.dst_plane2 + 64 * g, pack_sat_u8x2(acc[4 * g + 3] >> 19, acc[4 * g + 2] >> 19));
                    }
                }
                return;
            }
#pragma unroll
            for (int g = 0; g < NP; g++) {
                if (!UV) {
                    st_u16(dptr + 64 * g, pack_sat_u8x2(acc[2 * g + 1] >> 19, acc[2 * g] >> 19));
                } else {
                    st_u16(dptr + 64 * g, pack_sat_u8x2(acc[4 * g + 1] >> 19, acc[4 * g] >> 19));
                    st_u16(dptr + a.dst_plane2 + 64 * g, pack_sat_u8x2(acc[4 * g + 3] >> 19, acc[4 * g + 2] >> 19));
                }
            }
        };
        int rbase = rs;                                                      // absolute source row of the group's row 0
        uint32_t pv[SC ? SG * TV : 1][3], pn[SC ? SG * TV : 1][3];           // fused score: previous picture, this / next stage
        int pn_row = -1;                                                     // first row pn[] holds (-1: nothing requested)
        auto advance = [&]() -> bool {                                       // next output row; false when the item is done
            y++;
            dptr += a.dw;
            if (y >= y1) return false;
            vi += VS;
            vrel = last_next - rbase;
            last_next = vtab.t[vi + VS + TV];                                // (one entry past the launch's rows is padding)
#pragma unroll
            for (int j = 0; j < TV; j++) vc[j] = vtab.t[vi + j];
            return true;
        };
        int s = 0;                                                           // stage of load ld
        uint32_t soff = 0;                                                   // its byte offset in the warp's ring
        int groups_left = ngroups;
        for (int ld = 0; ld < nloads; ld++) {
            mbar_wait_u32(bar0 + 8u * (uint32_t)s, (phases >> s) & 1u);
            phases ^= 1u << s;
            const int ng = min(rg, groups_left);
            groups_left -= ng;
            uint32_t goff = soff;
#pragma unroll 1
            for (int gis = 0; gis < ng;) {
#pragma unroll
                for (int g = 0; g < (HS ? 1 : NP); g++) ga[g] = addr[g] + goff;
                int adv = 1;                                                 // groups this iteration consumes
                if (MASK && gis + SG <= ng && vrel == FIRSTK && y >= a.reg_lo && y + SG * NOUT <= ylim &&
                    (!SC || ((unsigned)(rbase - own_lo) < own_span && (unsigned)(rbase + SG * TV - 1 - own_lo) < own_span))) {
                    // ---- a whole stage of regular groups: the schedule is static (no branch, no table fetch), the
                    // coefficients are parameter-space constants, and horizontal / vertical work of neighbouring rows
                    // interleaves freely because it is one basic block
                    if constexpr (SC) {
                        // the previous picture's words for THIS stage were requested one stage ago; request the next
                        // stage's now, so that their latency hides behind this stage's arithmetic
                        const int ppw = a.sc_pitch >> 2;
                        if (pn_row == rbase) {
#pragma unroll
                            for (int r = 0; r < SG * TV; r++) { pv[r][0] = pn[r][0]; pv[r][1] = pn[r][1]; pv[r][2] = pn[r][2]; }
                        } else {
                            const uint32_t *pp = reinterpret_cast<const uint32_t *>(prev_lane + (size_t)rbase * a.sc_pitch);
#pragma unroll
                            for (int r = 0; r < SG * TV; r++) {
                                pv[r][0] = __ldg(pp + r * ppw); pv[r][1] = __ldg(pp + r * ppw + 1); pv[r][2] = __ldg(pp + r * ppw + 2);
                            }
                        }
                        const int nxt = rbase + SG * TV;
                        if (nxt + SG * TV <= rs + nrows) {
                            const uint32_t *pp = reinterpret_cast<const uint32_t *>(prev_lane + (size_t)nxt * a.sc_pitch);
#pragma unroll
                            for (int r = 0; r < SG * TV; r++) {
                                pn[r][0] = __ldg(pp + r * ppw); pn[r][1] = __ldg(pp + r * ppw + 1); pn[r][2] = __ldg(pp + r * ppw + 2);
                            }
                            pn_row = nxt;
                        } else {
                            pn_row = -1;
                        }
                    }
                    static_for<0, SG * TV>([&](auto kc) {
                        constexpr int kk = decltype(kc)::value, k = kk % TV;
                        hpass(kk, m[k], rbase + kk, true, pv[SC ? kk : 0]);
                        if constexpr ((MASK >> k) & 1) {
                            constexpr int q = (popc_c((unsigned)MASK & ((1u << k) - 1u)) + (kk / TV) * NOUT) % Q;
                            int acc[NM];
#pragma unroll
                            for (int c = 0; c < NM; c++) acc[c] = rnd;
#pragma unroll
                            for (int j = 0; j < ((VT_ABLATE & 1) ? 1 : TV); j++) {
#pragma unroll
                                for (int c = 0; c < NM; c++) acc[c] += m[(k + 1 + j) % TV][c] * a.sc[q * TV + j];
                            }
                            vstore(acc);
                            dptr += a.dw;
                        }
                    });
                    adv = SG;
                    y += SG * NOUT;
                    vi += SG * NOUT * VS;
                    stale = true;
                    if (y >= y1) goto item_done;
                    vrel = vtab.t[vi + TV] - rbase;
                } else {
                    if (MASK && stale) {
                        last_next = vtab.t[vi + VS + TV];
#pragma unroll
                        for (int j = 0; j < TV; j++) vc[j] = vtab.t[vi + j];
                        stale = false;
                    }
#pragma unroll
                    for (int k = 0; k < TV; k++) {
                        hpass(k, m[k], rbase + k, false, nullptr);
                        // ---- vertical pass for every output row whose window ends at this source row
                        while (vrel == k) {
                            int acc[NM];
#pragma unroll
                            for (int c = 0; c < NM; c++) acc[c] = rnd;
#pragma unroll
                            for (int j = 0; j < ((VT_ABLATE & 1) ? 1 : TV); j++) {
#pragma unroll
                                for (int c = 0; c < NM; c++) acc[c] += m[(k + 1 + j) % TV][c] * vc[j];
                            }
                            vstore(acc);
                            if (!advance()) goto item_done;                  // the rest of this group's rows feed nothing
                        }
                    }
                }
                vrel -= adv * TV;
                rbase += adv * TV;
                goff += adv * group_bytes;
                gis += adv;
            }
            __syncwarp();                                                    // every lane is done with this stage
            if (lane == 0 && ld + nst < nloads) {
                mbar_expect_tx(&bars[s], stage_tx);
#pragma unroll 1
                for (int b = 0; b < a.n_boxes; b++)
                    tma_load_3d(wbase + soff + (size_t)b * a.box_bytes, &tmap, &bars[s],
                                (HS ? hs_xa : a.box_x0[strip * a.n_boxes + b]) >> 2, rs + (ld + nst) * stage_rows, f);
            }
            s++;
            soff += (uint32_t)a.stage_bytes;
            if (s == nst) {
                s = 0;
                soff = 0;
            }
        }
    item_done:;
        if constexpr (SC) {
            // ---- flush this item's counters into the picture's histogram and SAD; leave the counters zeroed
            __syncwarp();
            const uint32_t cbase = wsm + (uint32_t)a.sc_cnt_off;
#pragma unroll 1
            for (int r = 0; r < 4; r++) {
                const uint32_t row = 4u * (uint32_t)lane + (uint32_t)r;
                uint32_t lo = 0, hi = 0;
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const uint32_t ad = cbase + row * 128u + (((uint32_t)j + (uint32_t)lane) & 31u) * 4u;   // rotated: no bank conflicts
                    uint32_t wv;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv) : "r"(ad) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ad), "r"(0u) : "memory");
                    lo += wv & 0xFFFFu;
                    hi += wv >> 16;
                }
                if (lo) atomicAdd(a.sc_hist + (size_t)f * 256 + row, lo);
                if (hi) atomicAdd(a.sc_hist + (size_t)f * 256 + 128 + row, hi);
            }
            const uint32_t tot = __reduce_add_sync(0xffffffffu, sad_acc);
            if (lane == 0 && tot) atomicAdd(a.sc_sad + f, (unsigned long long)tot);
            __syncwarp();
        }
    }
#if VT_PAIR_TIMES
    if (a.dbg_times && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.dbg_times[2 * gw + 1] = t;
    }
#endif
}

// ---- host side ---------------------------------------------------------------------------------------------------
namespace {

int pad_hp(int taps) {
    const int hp = (taps + 1) / 2;
    return hp <= 3 ? 3 : (hp <= 4 ? 4 : (hp <= 6 ? 6 : 0));
}
int pad_tv(int taps) { return taps < 2 ? 0 : (taps <= 6 ? 6 : (taps <= 8 ? 8 : (taps <= 12 ? 12 : 0))); }
int vstride_for(int tv) { return (tv + 1 + 3) & ~3; }

int upload(const void *h, size_t n, void **d) {
    if (cudaMalloc(d, n ? n : 1) != cudaSuccess) return VT_ERR_NOMEM;
    if (n && cudaMemcpy(*d, h, n, cudaMemcpyHostToDevice) != cudaSuccess) return VT_ERR_CUDA;
    return VT_OK;
}

template <int HP, int TV, bool UV, int MASK, int Q, int HS = 0, bool SC = false>
int launch_t(const vt_scale_plan::Pair &s, const CUtensorMap &tm, PairArgs a, int rows_total, cudaStream_t st) {
    auto k = scale_pair_kernel<HP, TV, UV, MASK, Q, HS, SC>;
    static int smem_set_dev[VT_MAX_DEVICES] = {0}, blocks_per_sm_dev[VT_MAX_DEVICES] = {0};   // per device ordinal
    if (SC) {                                 // the warp's counters follow its ring
        a.sc_cnt_off = s.warp_smem;
        a.warp_smem = s.warp_smem + PAIR_SC_COUNTER_BYTES;
    }
    const int smem = a.warp_smem * 4;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int &smem_set = smem_set_dev[current_device()], &blocks_per_sm = blocks_per_sm_dev[current_device()];
    if (smem > smem_set || !blocks_per_sm) {
        VT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        VT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k, 128, smem));
        if (blocks_per_sm < 1) {
            set_error("scale_pair_kernel<%d,%d,%d,0x%x>: does not fit an SM (smem %d)", HP, TV, (int)UV, MASK, smem);
            return VT_ERR_UNSUPPORTED;
        }
        smem_set = smem;
    }
    static VTab<TV> vt_host;                 // 28 KB staging for the parameter copy; filled under the launch lock
    const int cap = VCfg<TV>::ROWS - 1;      // the kernel reads one table entry ahead
    if (SC && rows_total > cap) {
        set_error("scale_pair_kernel: fused score needs the whole picture in one launch (%d rows > %d)", rows_total, cap);
        return VT_ERR_UNSUPPORTED;
    }
    for (int yb = 0; yb < rows_total; yb += cap) {
        const int ye = std::min(rows_total, yb + cap);
        for (size_t i = 0, n = (size_t)(ye - yb) * s.vstride; i < n; i++) vt_host.t[i] = (int16_t)s.vtab[(size_t)yb * s.vstride + i];
        a.y_begin = yb;
        a.y_end = ye;
        const int rows = ye - yb;
        // Grid: every SM full.  Work units are (frame, segment); a warp keeps its strip and takes a contiguous run of
        // units (adjacent segments of a frame merge into one item), so the segment height only sets the balancing
        // granularity: pick the height (a multiple of the regular group's output rows) that minimises
        // units per warp x rows per unit + items per warp x (window overlap + set-up).
        static const int grid_cap = getenv("VT_PAIR_GRID_BLOCKS") ? atoi(getenv("VT_PAIR_GRID_BLOCKS")) : 0;   // experiments
        const int grid = sm_count() * ((grid_cap > 0 && grid_cap < blocks_per_sm) ? grid_cap : blocks_per_sm);
        const long long wps = std::max<long long>(1, (long long)grid * 4 / a.n_strips);   // warps per strip
        constexpr int nout = MASK ? popc_c((unsigned)MASK) : 1;
        int best_rows = rows;
        double best_cost = 1e30;
        for (int sr = nout; sr <= rows + nout - 1; sr += nout) {
            if (sr < 2 * nout && sr < rows) continue;
            const int n = (rows + sr - 1) / sr;
            const long long upw = ((long long)a.n_frames * n + wps - 1) / wps;             // units per warp
            const double items = 1.0 + (double)upw / n;                                    // frames a run touches
            const double cost = (double)upw * sr * s.src_rows_per_dst_row + items * ((TV - 1) + 12.0);
            if (cost < best_cost * 0.9995) { best_cost = cost; best_rows = sr; }
        }
        a.seg_rows = best_rows;
        a.n_segs = (rows + a.seg_rows - 1) / a.seg_rows;
        static const char *dbg_path = VT_PAIR_TIMES ? getenv("VT_PAIR_DEBUG_TIMES") : nullptr;   // per-warp start/end times
        if (dbg_path && !a.dbg_times) cudaMalloc((void **)&a.dbg_times, sizeof(unsigned long long) * 2 * grid * 4);
        k<<<grid, 128, smem, st>>>(tm, a, vt_host);
        VT_LAUNCHED("scale_pair_kernel");
        if (dbg_path) {
            std::vector<unsigned long long> h((size_t)2 * grid * 4);
            cudaStreamSynchronize(st);
            cudaMemcpy(h.data(), a.dbg_times, h.size() * 8, cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(dbg_path, UV ? "ab" : "wb")) {
                fwrite(h.data(), 8, h.size(), f);
                fclose(f);
            }
            cudaFree(a.dbg_times);
            a.dbg_times = nullptr;
        }
    }
    return VT_OK;
}

template <bool UV>
int dispatch(int hp, int tv, bool hs, const vt_scale_plan::Pair &s, const CUtensorMap &tm, const PairArgs &a, int rows,
             cudaStream_t st) {
    if constexpr (!UV) {
        if (a.sc_hist) {                      // fused score: only the 3:2 luma layout is instantiated (checked by the caller)
            if (hs && s.mask == 0x36 && s.n_phases == 2 && hp == 3 && tv == 6)
                return launch_t<3, 6, false, 0x36, 2, 1, true>(s, tm, a, rows, st);
            set_error("scale_pair: no fused-score instantiation for this plan");
            return VT_ERR_UNSUPPORTED;
        }
    }
    // static horizontal pattern + static vertical schedule (exact 3:2 both ways: 1080p -> 720p)
    if (hs && s.mask == 0x36 && s.n_phases == 2 && hp == 3 && tv == 6) return launch_t<3, 6, UV, 0x36, 2, 1>(s, tm, a, rows, st);
    if (hs && s.mask == 0xAA && s.n_phases == 1 && hp == 4 && tv == 8) return launch_t<4, 8, UV, 0xAA, 1, 2>(s, tm, a, rows, st);
    if (hs && s.mask == 0x924 && s.n_phases == 1 && hp == 6 && tv == 12) return launch_t<6, 12, UV, 0x924, 1, 3>(s, tm, a, rows, st);
    // static vertical schedules (exact 3:2, 2:1 and 3:1 ratios)
    if (s.mask == 0x36 && s.n_phases == 2 && hp == 3 && tv == 6) return launch_t<3, 6, UV, 0x36, 2>(s, tm, a, rows, st);
    if (s.mask == 0xAA && s.n_phases == 1 && hp == 4 && tv == 8) return launch_t<4, 8, UV, 0xAA, 1>(s, tm, a, rows, st);
    if (s.mask == 0x924 && s.n_phases == 1 && hp == 6 && tv == 12) return launch_t<6, 12, UV, 0x924, 1>(s, tm, a, rows, st);
#define VT_CASE(H, T) if (hp == H && tv == T) return launch_t<H, T, UV, 0, 1>(s, tm, a, rows, st)
    VT_CASE(3, 6); VT_CASE(3, 8); VT_CASE(3, 12);
    VT_CASE(4, 6); VT_CASE(4, 8); VT_CASE(4, 12);
    VT_CASE(6, 6); VT_CASE(6, 8); VT_CASE(6, 12);
#undef VT_CASE
    set_error("scale_pair: no instantiation for hp=%d tv=%d", hp, tv);
    return VT_ERR_UNSUPPORTED;
}

// Finds the static vertical schedule of a plane, if it has one: output rows [reg_lo, reg_hi) whose window positions
// advance by P source rows every Q output rows with Q repeating coefficient sets, P dividing the group size tv, and
// an alignment (row phase r0 mod P) for which the window ends inside a group are exactly the canonical mask.
void find_static_schedule(vt_scale_plan::Pair &s, const std::vector<int32_t> &vpos, const std::vector<int16_t> &vco,
                          int vtaps, int src_rows, int dh) {
    s.mask = 0;
    s.n_phases = 1;
    s.align_p = 1;
    s.align_r0 = 0;
    s.reg_lo = s.reg_hi = 0;
    int g = src_rows, b = dh;
    while (b) { const int t = g % b; g = b; b = t; }
    const int P = src_rows / g, Q = dh / g;
    const unsigned canon = (s.tv == 6 && P == 3 && Q == 2) ? 0x36u : (s.tv == 8 && P == 2 && Q == 1) ? 0xAAu
                         : (s.tv == 12 && P == 3 && Q == 1) ? 0x924u : 0u;
    if (!canon || dh < 8 * Q) return;
    const int yref = (dh / 2) / Q * Q;
    auto regular = [&](int y) {
        int q = (y - yref) % Q;
        if (q < 0) q += Q;
        const int n = (y - yref - q) / Q;
        if (vpos[y] != vpos[yref] + n * P + (vpos[yref + q] - vpos[yref])) return false;
        for (int j = 0; j < vtaps; j++)
            if (vco[(size_t)y * vtaps + j] != vco[(size_t)(yref + q) * vtaps + j]) return false;
        return true;
    };
    int lo = yref, hi = yref;
    while (lo > 0 && regular(lo - 1)) lo--;
    while (hi < dh && regular(hi)) hi++;
    if (hi - lo < 4 * Q) return;
    // alignment: a group starting at row rb (inside the regular region) must see window ends exactly at the mask
    for (int al = 0; al < P; al++) {
        const int rb = vpos[yref] + vtaps + al;                 // first row after yref's window end, plus al
        unsigned m = 0;
        int y_first = -1;
        for (int y = lo; y < hi; y++) {
            const int last = vpos[y] + vtaps - 1;
            if (last >= rb && last < rb + s.tv) {
                m |= 1u << (last - rb);
                if (y_first < 0) y_first = y;
            }
        }
        if (m != canon || y_first < 0 || y_first + Q > hi) continue;
        s.mask = (int)canon;
        s.n_phases = Q;
        s.align_p = P;
        s.align_r0 = ((rb % P) + P) % P;
        s.reg_lo = lo;
        s.reg_hi = hi;
        for (int q = 0; q < Q; q++)
            for (int j = 0; j < s.tv; j++) s.sc[q * s.tv + j] = 0;
        for (int q = 0; q < Q; q++)
            for (int j = 0; j < vtaps; j++)
                s.sc[q * s.tv + (s.tv - vtaps) + j] = vco[(size_t)(y_first + q) * vtaps + j];
        return;
    }
}

}  // namespace

// Builds the tables the pair kernel reads for one plane kind (c = 0 luma, 1 chroma).  Leaves pair[c].ok false
// (generic kernels stay in charge) when the shape is outside what the kernel handles.
int build_pair(vt_scale_plan *p, int c) {
    vt_scale_plan::Pair &s = p->pair[c];
    const bool uv = c == 1;
    const int dw = c ? p->cdw : p->dw, dh = c ? p->cdh : p->dh;
    const int bpp = uv ? 2 : 1;
    const int ht = p->htaps[c], vtaps = p->vtaps[c];
    const std::vector<int32_t> &hpos = p->h_hpos[c], &vpos = p->h_vpos[c];
    const std::vector<int16_t> &hco = p->h_hcoef[c], &vco = p->h_vcoef[c];
    s.hp = pad_hp(ht);
    s.tv = pad_tv(vtaps);
    if (!s.hp || !s.tv) return VT_OK;
    const int bn = s.hp + 1;
    s.np = uv ? pair_np(s.hp, s.tv) / 2 : pair_np(s.hp, s.tv);
    s.strip_cols = s.np * 64;
    // an odd CHROMA width is fine (byte stores; the last strip slides left onto the last column, so its pairs start on
    // odd columns); the luma width is even by construction (`scale=-2:H`)
    if ((!uv && (dw & 1)) || (p->dw & 1) || dw < s.strip_cols) return VT_OK;
    if ((c ? p->csh : p->sh) > 32767) return VT_OK;              // the vertical table holds source rows as int16
    for (int x = 0; x + 1 < dw; x++)
        if (hpos[x + 1] < hpos[x]) return VT_OK;
    for (int x = 0; x + 1 < dw; x += (dw & 1) ? 1 : 2)           // every pair a lane can own
        if (hpos[x + 1] - hpos[x] + ht > 2 * bn) return VT_OK;
    for (int y = 0; y + 1 < dh; y++)
        if (vpos[y + 1] < vpos[y]) return VT_OK;
    s.n_strips = (dw + s.strip_cols - 1) / s.strip_cols;
    const int naw = uv ? bn : (bn + 1) / 2, nw = naw + 1;
    s.lt_words = (2 * s.hp + 2 + 3) & ~3;
    // boxes: a strip is cut into n_boxes runs of box_cols columns whose source span fits a 256-byte TMA box row
    int box_cols = s.strip_cols, need = 0;
    std::vector<int32_t> scol((size_t)s.n_strips);
    for (int strip = 0; strip < s.n_strips; strip++) scol[strip] = std::min(strip * s.strip_cols, dw - s.strip_cols);
    for (;; box_cols >>= 1) {
        if (box_cols < 16) return VT_OK;
        need = 0;
        for (int strip = 0; strip < s.n_strips; strip++)
            for (int b = 0; b * box_cols < s.strip_cols; b++) {
                const int xf = scol[strip] + b * box_cols, xl = xf + box_cols - 2;   // first / last even column
                need = std::max(need, hpos[xl] * bpp + 4 * nw - ((hpos[xf] * bpp) & ~15));
            }
        if (need <= (int)PAIR_TILE_W) break;
    }
    s.n_boxes = s.strip_cols / box_cols;
    s.tile_w = (int)PAIR_TILE_W;
    // stage geometry: TWO stages of about 12 rows.  Measured with tools/ubench_tma.cu on B200: per-warp rings with two
    // stages in flight pull 6.3-6.5 TB/s, three or four stages only 2.8-4.4 TB/s (HBM row conflicts between the many
    // row blocks then in flight); box height and alignment hardly matter.
    const int nb = pair_min_blocks(s.hp, s.tv);
    s.n_stages = 2;
    s.groups_per_stage = pair_sg(s.tv);
    if (getenv("VT_PAIR_RG") && getenv("VT_PAIR_NST")) {                    // geometry experiments
        s.groups_per_stage = atoi(getenv("VT_PAIR_RG"));
        s.n_stages = atoi(getenv("VT_PAIR_NST"));
    }
    for (;; s.groups_per_stage--) {
        if (s.groups_per_stage < 1) return VT_OK;
        s.stage_rows = s.groups_per_stage * s.tv;
        s.box_bytes = (s.stage_rows * s.tile_w + 127) & ~127;
        s.stage_bytes = s.box_bytes * s.n_boxes;
        s.warp_smem = (s.n_stages * s.stage_bytes + 64 + 127) & ~127;
        if (4 * s.warp_smem * nb + nb * 1024 <= 227 * 1024) break;
    }
    if (s.n_stages < 2) return VT_OK;
    std::vector<int32_t> bx((size_t)s.n_strips * s.n_boxes);
    std::vector<uint32_t> lt((size_t)s.n_strips * s.np * 32 * s.lt_words, 0);
    for (int strip = 0; strip < s.n_strips; strip++) {
        for (int b = 0; b < s.n_boxes; b++) bx[(size_t)strip * s.n_boxes + b] = (hpos[scol[strip] + b * box_cols] * bpp) & ~15;
        for (int g = 0; g < s.np; g++)
            for (int lane = 0; lane < 32; lane++) {
                const int xa = scol[strip] + g * 64 + 2 * lane, xb = xa + 1;
                const int b = (g * 64 + 2 * lane) / box_cols;
                uint32_t *t = &lt[((size_t)(strip * s.np + g) * 32 + lane) * s.lt_words];
                t[0] = (uint32_t)(b * s.box_bytes + hpos[xa] * bpp - bx[(size_t)strip * s.n_boxes + b]);
                for (int j = 0; j < ht; j++) {
                    const uint32_t v = (uint16_t)hco[(size_t)xa * ht + j];
                    t[1 + j / 2] |= (j & 1) ? (v << 16) : v;
                }
                const int d = hpos[xb] - hpos[xa];                 // the odd column's taps start d samples further right
                // samples 2..2*hp-1 (halfwords 1..hp-1) keep their place; what falls outside goes to the spare halfword
                const int words = uv ? (bn + 1) / 2 : (bn + 1) / 2; // aligned words (luma) / de-interleaved words (chroma)
                int spare[2] = {-1, -1}, nspare = 0;
                for (int j = 0; j < ht; j++) {
                    const uint32_t v = (uint16_t)hco[(size_t)xb * ht + j];
                    const int sidx = d + j;
                    if (sidx >= 2 && sidx < 2 * s.hp) {
                        t[1 + s.hp + (sidx / 2 - 1)] |= (sidx & 1) ? (v << 16) : v;
                    } else if (v) {
                        if (nspare == 2) return VT_OK;             // cannot happen for d + ht <= 2*hp + 2; stay generic
                        t[1 + s.hp + (s.hp - 1)] |= nspare ? (v << 16) : v;
                        spare[nspare++] = sidx;
                    }
                }
                // byte selector for __byte_perm(first word, last word, sel): low byte = spare[0], next = spare[1]
                uint32_t sel = 0;
                for (int q = 0; q < 2; q++) {
                    int sidx = spare[q] < 0 ? 0 : spare[q];
                    int code = sidx < 4 ? sidx : 4 + (sidx - 4 * (words - 1));
                    if (code < 0 || code > 7) return VT_OK;
                    sel |= (uint32_t)code << (4 * q);
                }
                t[1 + 2 * s.hp] = sel | 0x4400u;                   // upper halfword of the permute: don't care (zero coefficients)
            }
    }
    s.src_rows_per_dst_row = (double)(c ? p->csh : p->sh) / dh;
    find_static_schedule(s, vpos, vco, vtaps, c ? p->csh : p->sh, dh);
    s.vstride = vstride_for(s.tv);
    s.vtab.assign((size_t)dh * s.vstride, 0);
    for (int y = 0; y < dh; y++) {
        int32_t *t = &s.vtab[(size_t)y * s.vstride];
        for (int j = 0; j < vtaps; j++) t[(s.tv - vtaps) + j] = vco[(size_t)y * vtaps + j];
        t[s.tv] = vpos[y] + vtaps - 1;
    }
    // ---- static horizontal pattern (the kernel's HS parameter): exact 3:2, every column's taps inside the six-sample
    // window that starts at 3*(x/2) - 2 + (x&1).  libswscale's edge columns (clamped position, taps folded onto the
    // border sample) fit because the folded coefficients land on in-range samples of that same window; the samples
    // left of column 0 / right of the last column are TMA zero fill and carry zero coefficients.
    s.hs = false;
    const int ncol = 2 * s.np;                                   // adjacent output columns per lane
    const int hsid = (s.hp == 3 && s.tv == 6 && s.mask == 0x36 && s.n_phases == 2) ? 1
                   : (s.hp == 4 && s.tv == 8 && s.mask == 0xAA && s.n_phases == 1) ? 2
                   : (s.hp == 6 && s.tv == 12 && s.mask == 0x924 && s.n_phases == 1) ? 3 : 0;
    std::vector<int32_t> bx_hs((size_t)s.n_strips);
    std::vector<uint32_t> lt_hs((size_t)s.n_strips * 32 * ncol * s.hp, 0);
    bool hs = hsid && s.n_boxes == 1 && dw % (uv ? std::max(ncol, 2) : std::max(ncol, 4)) == 0 && !getenv("VT_PAIR_NO_HS");
    if (hs && hs_lane_stride(hsid, uv, ncol) * 31 + 4 * hs_words(hsid, uv, ncol, s.hp) + 12 > (int)PAIR_TILE_W) hs = false;
    for (int strip = 0; hs && strip < s.n_strips; strip++) {
        if (scol[strip] % ncol) { hs = false; break; }
        bx_hs[strip] = bpp * (hs_nominal(hsid, scol[strip]) - hs_pre(hsid, uv));
        if (bx_hs[strip] & 3) { hs = false; break; }             // lanes fetch aligned 32-bit words
        for (int lane = 0; hs && lane < 32; lane++)
            for (int cidx = 0; hs && cidx < ncol; cidx++) {
                const int x = scol[strip] + ncol * lane + cidx;
                const int nominal = hs_nominal(hsid, x);
                uint32_t *t = &lt_hs[(((size_t)strip * 32 + lane) * ncol + cidx) * s.hp];
                for (int j = 0; j < ht; j++) {
                    const uint32_t v = (uint16_t)hco[(size_t)x * ht + j];
                    if (!v) continue;
                    const int tap = hpos[x] + j - nominal;
                    if (tap < 0 || tap >= 2 * s.hp) { hs = false; break; }
                    t[tap / 2] |= (tap & 1) ? (v << 16) : v;
                }
            }
    }
    int rc = upload(bx.data(), bx.size() * 4, (void **)&s.box_x0);
    if (rc == VT_OK && hs) rc = upload(bx_hs.data(), bx_hs.size() * 4, (void **)&s.box_x0_hs);
    if (rc == VT_OK && hs) rc = upload(lt_hs.data(), lt_hs.size() * 4, (void **)&s.lane_tab_hs);
    s.hs = hs && rc == VT_OK;
    // fused scene score (luma, exact 3:2): strips tile the source exactly (no slid last strip), a lane's 8 columns sit
    // on 12 source bytes, the last window ends on the last source row and the picture fits one launch
    s.score_ok = s.hs && !uv && hsid == 1 && dw % s.strip_cols == 0 && (long long)p->sw * 2 == (long long)dw * 3 &&
                 vpos[dh - 1] + vtaps == p->sh && dh <= VCfg<6>::ROWS - 1;
    for (int strip = 0; s.score_ok && strip < s.n_strips; strip++)
        if (bx_hs[strip] + 4 != scol[strip] / 2 * 3) s.score_ok = false;
    if (rc == VT_OK) rc = upload(scol.data(), scol.size() * 4, (void **)&s.strip_col);
    if (rc == VT_OK) rc = upload(lt.data(), lt.size() * 4, (void **)&s.lane_tab);
    if (rc != VT_OK) return rc;
    s.ok = true;
    return VT_OK;
}

void free_pair(vt_scale_plan *p) {
    for (int c = 0; c < 2; c++) {
        cudaFree(p->pair[c].box_x0);
        cudaFree(p->pair[c].strip_col);
        cudaFree(p->pair[c].lane_tab);
        cudaFree(p->pair[c].box_x0_hs);
        cudaFree(p->pair[c].lane_tab_hs);
        p->pair[c].box_x0 = p->pair[c].strip_col = p->pair[c].box_x0_hs = nullptr;
        p->pair[c].lane_tab = p->pair[c].lane_tab_hs = nullptr;
    }
}

int launch_pair(const vt_scale_plan *p, int c, const uint8_t *src, int pitch, size_t src_fs, uint8_t *dst, size_t dst_fs,
                int n_frames, cudaStream_t st, const PairScore *score) {
    const vt_scale_plan::Pair &s = p->pair[c];
    const bool uv = c == 1;
    CUtensorMap tm;
    const uint8_t *base = uv ? src + (size_t)pitch * p->sh : src;
    const int row_bytes = uv ? 2 * p->csw : p->sw;
    const int rows = uv ? p->csh : p->sh;
    int rc = make_tmap_u32_3d(&tm, base, row_bytes, rows, n_frames, pitch, src_fs, s.tile_w, s.stage_rows);
    if (rc) return rc;
    PairArgs a;
    // the adjacent-column layout stores up to 8 (luma) / 4 (chroma) bytes per lane: needs the planes aligned accordingly
    const size_t ysz = (size_t)p->dw * p->dh, csz = (size_t)p->cdw * p->cdh;
    const bool hs = s.hs && (uv ? ((uintptr_t)dst % 4 == 0 && dst_fs % 4 == 0 && ysz % 4 == 0 && csz % 4 == 0)
                                : ((uintptr_t)dst % 8 == 0 && dst_fs % 8 == 0));
    a.lane_tab = hs ? s.lane_tab_hs : s.lane_tab;
    a.box_x0 = hs ? s.box_x0_hs : s.box_x0;
    a.strip_col = s.strip_col;
    a.dst = uv ? dst + (size_t)p->dw * p->dh : dst;
    a.dst_fs = dst_fs;
    a.dst_plane2 = (size_t)p->cdw * p->cdh;
    a.n_frames = n_frames;
    a.n_strips = s.n_strips;
    a.n_segs = a.seg_rows = 0;
    a.dw = (unsigned)(uv ? p->cdw : p->dw);
    a.y_begin = a.y_end = 0;
    a.tile_w = s.tile_w;
    a.n_boxes = s.n_boxes;
    a.groups_per_stage = s.groups_per_stage;
    a.box_bytes = s.box_bytes;
    a.stage_bytes = s.stage_bytes;
    a.n_stages = s.n_stages;
    a.warp_smem = s.warp_smem;
    a.round_bias = 1 << 18;
    a.dbg_times = nullptr;
    a.sc_src = src;
    a.sc_prev0 = score ? score->prev0 : nullptr;
    a.sc_fs = src_fs;
    a.sc_pitch = pitch;
    a.sc_rows = p->sh;
    a.sc_sad = score ? (unsigned long long *)score->sad : nullptr;
    a.sc_hist = score ? score->hist : nullptr;
    a.sc_cnt_off = 0;
    if (score && (uv || !hs || !s.score_ok)) {
        set_error("scale_pair: this plan / alignment has no fused score path");
        return VT_ERR_UNSUPPORTED;
    }
    a.reg_lo = s.reg_lo;
    a.reg_hi = s.reg_hi;
    a.align_p = s.align_p;
    a.align_r0 = s.align_r0;
    for (int i = 0; i < 24; i++) a.sc[i] = s.sc[i];
    const int dh = uv ? p->cdh : p->dh;
    return uv ? dispatch<true>(s.hp, s.tv, hs, s, tm, a, dh, st) : dispatch<false>(s.hp, s.tv, hs, s, tm, a, dh, st);
}

}  // namespace vt
