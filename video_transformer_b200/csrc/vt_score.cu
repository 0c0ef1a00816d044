// vt_score.cu -- K3: per-frame 256-bin luma histogram + SAD(cur, prev).
//
// HBM-bound byte work: 2 bytes read per pixel (cur + prev luma), ~1 KB written per frame.
// The histogram is the hard part: shared-memory atomics on a per-warp histogram serialise on flat picture
// areas (test patterns, letterboxing) and cost data-dependent bank conflicts everywhere else.  This kernel's
// hot loop is data independent.  Every lane of a warp owns a private column of 256 one-byte counters packed
// four to a 32-bit word; word (row r, lane l) sits at  warp_base + r*128 + l*4, so the bank is the lane and a
// warp's 32 updates never conflict.  Bin b lives in row (b & 63), byte (b >> 6).  An update is one
// shared-memory RED of (1 << 8*(b>>6)) on that word (measured on B200: 9.3 px/clk/SM against 7.6 for
// LDS.U8/IADD/STS.U8 on the same layout and 4.9-7.7 for byte counters that share words between lanes;
// tools/ubench_smem.cu).  A lane sees at most 240 pixels between flushes, so no byte can carry into its
// neighbour.  A flush sums each row over the 32 lanes with rotated (conflict-free) word reads and adds the
// four bins of the row to the block histogram.  SAD is VABSDIFF4 + IDP.4A on the same 128-bit loads.
#include <algorithm>
#include <type_traits>

#include "vt_common.cuh"

namespace vt {

constexpr int SC_WARPS = 8;
constexpr int SC_THREADS = SC_WARPS * 32;
constexpr int SC_CNT_BYTES = 8192;                                   // per warp: 64 rows x 32 lanes x 4 packed u8
constexpr int SC_SMEM = SC_WARPS * SC_CNT_BYTES + SC_CNT_BYTES + 1024;  // + alignment slack + block hist

__device__ __forceinline__ void red_shared(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// 4 pixels of one 32-bit word.  base = (8 KB aligned warp region) | lane*4.
// The ALU pipe (half rate on sm_100a: LOP3/SHF/PRMT issue every other cycle, tools/ubench_pipes.cu) was the busiest
// one in this kernel, so everything that is linear goes to the FMA pipe instead: the counter address of pixel k is
// ONE dp4a, base + 128 * byte_k(rows) (u8 x u8 dot product with the one-hot vector 128 << 8k), and the shift count of
// its increment is another dp4a that picks byte k of j8.  The increment itself is a wrapping funnel shift of 1
// (only the low five bits of the count matter, so pixel 0 needs no extraction at all).
__device__ __forceinline__ uint32_t one_shl_wrap(uint32_t n) {   // 1 << (n & 31)
    uint32_t d;
    asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(0u), "r"(1u), "r"(n));
    return d;
}
__device__ __forceinline__ void hist_word(uint32_t w, uint32_t base) {
    const uint32_t r = w & 0x3F3F3F3Fu;          // row index of each pixel
    const uint32_t j8 = (w >> 3) & 0x18181818u;  // 8 * (pixel >> 6): bit offset of its counter in the word
    red_shared(__dp4a(r, 0x00000080u, base), one_shl_wrap(j8));
    red_shared(__dp4a(r, 0x00008000u, base), one_shl_wrap(__dp4a(j8, 0x00000100u, 0u)));
    red_shared(__dp4a(r, 0x00800000u, base), one_shl_wrap(__dp4a(j8, 0x00010000u, 0u)));
    red_shared(__dp4a(r, 0x80000000u, base), one_shl_wrap(__dp4a(j8, 0x01000000u, 0u)));
}

// Warp-collective: add the lane-private packed counters into the block histogram and clear them.
// Lane l owns rows 2l and 2l+1 and reads each as eight 128-bit chunks, starting at chunk (l & 7) and
// rotating: a 128-bit shared load is served eight lanes at a time, and eight consecutive lanes always ask
// for eight different chunks (different banks), so the loads are conflict free although every lane reads
// a different row.  All 16 loads are issued before the first use.
__device__ __forceinline__ void flush_counters(uint32_t warp_cnt, uint32_t *bhist, int lane) {
    __syncwarp();
    uint32_t even[2], odd[2];  // per row: two 16-bit sums each (bytes 0/2 and bytes 1/3; max 32*255 = 8160)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t a = warp_cnt + (uint32_t)(2 * lane + h) * 128u + (uint32_t)((lane + i) & 7) * 16u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                         : "r"(a)
                         : "memory");
        }
        uint32_t e = 0, o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            e += (v[i].x & 0x00FF00FFu) + (v[i].y & 0x00FF00FFu) + (v[i].z & 0x00FF00FFu) + (v[i].w & 0x00FF00FFu);
            o += ((v[i].x >> 8) & 0x00FF00FFu) + ((v[i].y >> 8) & 0x00FF00FFu) + ((v[i].z >> 8) & 0x00FF00FFu) +
                 ((v[i].w >> 8) & 0x00FF00FFu);
        }
        even[h] = e;
        odd[h] = o;
    }
    __syncwarp();   // every lane has summed its rows: the counters can be cleared
#pragma unroll
    for (int p = 0; p < 16; p++)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(warp_cnt + (uint32_t)(p * 32 + lane) * 16u), "r"(0u)
                     : "memory");
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int r = 2 * lane + h;
        if (even[h] & 0xFFFFu) atomicAdd(&bhist[r], even[h] & 0xFFFFu);
        if (odd[h] & 0xFFFFu) atomicAdd(&bhist[r + 64], odd[h] & 0xFFFFu);
        if (even[h] >> 16) atomicAdd(&bhist[r + 128], even[h] >> 16);
        if (odd[h] >> 16) atomicAdd(&bhist[r + 192], odd[h] >> 16);
    }
    __syncwarp();
}

// Out-of-line copy for the mid-row flush (rows wider than 8160 pixels only): keeps the hot loop's code small enough
// for the instruction cache.
__device__ __noinline__ void flush_counters_cold(uint32_t warp_cnt, uint32_t *bhist, int lane) {
    flush_counters(warp_cnt, bhist, lane);
}

// Aligned fast path: luma/prev base and pitch are multiples of 16 bytes.
// U = passes (of 32 lanes x 16 bytes) per unpredicated block of the row loop: min(4, full passes per row).  Rows are
// walked as  [blocks of U full passes] [single full passes] [one partial pass with the byte mask of a ragged width];
// only the last step carries predicates.  U = 0 keeps the fully general loop (rows wider than 4080 pixels per lane
// budget, i.e. more than 15 passes, where the counters must be flushed inside a row).
// RAG = merged rows whose last group is ragged (width not a multiple of 16): a separate instantiation, because the
// byte masks in the merged passes cost the common case 3 % even when they are never applied.
template <int U, bool RAG>
__global__ void __launch_bounds__(SC_THREADS, 3)
score_kernel(const uint8_t *__restrict__ luma, int pitch, size_t frame_stride, int w, int h,
             const uint8_t *__restrict__ prev0, int rows_per_block, int chunks_per_frame, int chunks_per_block,
             int merge_k, int merge_p, unsigned long long *__restrict__ sad_out, uint32_t *__restrict__ hist_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s0 = smem_u32(smem_raw);
    const uint32_t cnt0 = (s0 + SC_CNT_BYTES - 1) & ~(uint32_t)(SC_CNT_BYTES - 1);  // 8 KB aligned
    uint32_t *bhist = (uint32_t *)(smem_raw + (cnt0 - s0) + SC_WARPS * SC_CNT_BYTES);
    const uint32_t warp_cnt = cnt0 + warp * SC_CNT_BYTES;

    // a block owns chunks_per_block consecutive row chunks of one frame: its 256 global histogram atomics and its
    // SAD atomics are paid once for all of them
    const int blocks_per_frame = (chunks_per_frame + chunks_per_block - 1) / chunks_per_block;
    const int f = blockIdx.x / blocks_per_frame;
    const int chunk0 = (blockIdx.x - f * blocks_per_frame) * chunks_per_block;
    const int chunk1 = min(chunks_per_frame, chunk0 + chunks_per_block);
    const int rows_first = chunk0 * rows_per_block, rows_last = min(h, chunk1 * rows_per_block);
    const uint8_t *cur = luma + (size_t)f * frame_stride;
    const uint8_t *prv = f ? cur - frame_stride : (prev0 ? prev0 : cur);

    bhist[tid] = 0;
#pragma unroll
    for (int p = 0; p < 16; p++)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(warp_cnt + (uint32_t)(p * 32 + lane) * 16u), "r"(0u)
                     : "memory");
    __syncthreads();

    const int ngroups = (w + 15) >> 4;  // 16-byte groups per row
    const int tail = w & 15;            // valid bytes in the last group (0 = full)
    const uint32_t base = warp_cnt | ((uint32_t)lane << 2);
    const int rpw = rows_per_block / SC_WARPS;   // consecutive rows owned by each warp

    // Rows whose width is not a multiple of 512 pixels end in a partial pass with idle lanes.  With merge_k > 1 a warp
    // walks merge_k rows at a time and packs their leftover groups into merge_p full-width passes (virtual index
    // v = pass*32 + lane -> row v / rem, group v % rem): 1280 -> 2 rows share one pass, 1920 -> 4 rows share three.
    constexpr int MAXP = 3;
    const int nfull_all = (w & 15) ? ngroups - 1 : ngroups;   // groups that need no byte mask
    const int gfull = nfull_all & ~31;                        // groups covered by full passes
    const int rem = ngroups - gfull;
    int moff[MAXP];                                           // byte offset from the first row of the group, < 0 = idle
    bool mrag[MAXP];                                          // this lane's group is the row's ragged last one
#pragma unroll
    for (int j = 0; j < MAXP; j++) {
        const int v = j * 32 + lane;
        moff[j] = -1;
        mrag[j] = false;
        if (U > 0 && merge_k > 1 && j < merge_p && v < merge_k * rem) {
            moff[j] = (v / rem) * pitch + (gfull + v % rem) * 16;
            mrag[j] = RAG && (v % rem == rem - 1);
        }
    }

    uint32_t sad = 0;
    for (int chunk = chunk0; chunk < chunk1; chunk++) {
    const int r0 = chunk * rows_per_block;
    const int r1 = min(h, r0 + rows_per_block);
    int budget = 0;                     // upper bound on pixels a lane has counted since the last flush
    const int rstep = (U > 0 && merge_k > 1) ? merge_k : 1;
    for (int rr = 0; rr < rpw; rr += rstep) {
        const int row_a = r0 + warp * rpw + rr;
        if (row_a >= r1) break;
        const int nrows = min(rstep, min(r1 - row_a, rpw - rr));
        const bool merged = rstep > 1 && nrows == rstep;
        if constexpr (U > 0) {
            if (merged) {
                // Full passes of the group's rows as ONE sequence of virtual passes, walked in blocks of U (then 2, then
                // 1) with all loads of a block issued before its first counter update; then the packed leftover passes.
                const uint8_t *ca = cur + (size_t)row_a * pitch;
                const uint8_t *pa = prv + (size_t)row_a * pitch;
                const int fullp = gfull >> 5, total = nrows * fullp;
                size_t off = (size_t)lane * 16;
                int pir = 0;                                 // pass inside the current row
                auto block = [&](auto nbc) {
                    constexpr int NB = decltype(nbc)::value;
                    uint4 c[NB], p[NB];
#pragma unroll
                    for (int k = 0; k < NB; k++) {
                        c[k] = ld_stream_u4(ca + off);
                        p[k] = ld_stream_u4(pa + off);
                        off += 512;
                        if (++pir == fullp) {
                            pir = 0;
                            off += (size_t)pitch - (size_t)fullp * 512;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NB; k++) {
                        sad = sad4(c[k].x, p[k].x, sad);
                        sad = sad4(c[k].y, p[k].y, sad);
                        sad = sad4(c[k].z, p[k].z, sad);
                        sad = sad4(c[k].w, p[k].w, sad);
                        hist_word(c[k].x, base);
                        hist_word(c[k].y, base);
                        hist_word(c[k].z, base);
                        hist_word(c[k].w, base);
                    }
                };
                int q = 0;
                for (; q + U <= total; q += U) block(std::integral_constant<int, (U > 0 ? U : 1)>{});
                if constexpr (U >= 3) {
                    if (q + 2 <= total) {
                        block(std::integral_constant<int, 2>{});
                        q += 2;
                    }
                }
                for (; q < total; q++) block(std::integral_constant<int, 1>{});
                uint4 c[MAXP], p[MAXP];
#pragma unroll
                for (int j = 0; j < MAXP; j++) {
                    c[j] = make_uint4(0, 0, 0, 0);
                    p[j] = c[j];
                    if (j < merge_p && moff[j] >= 0) {
                        c[j] = ld_stream_u4(ca + moff[j]);
                        p[j] = ld_stream_u4(pa + moff[j]);
                    }
                }
#pragma unroll
                if constexpr (RAG) {                         // bytes past the width count as zeros
#pragma unroll
                    for (int j = 0; j < MAXP; j++)
                        if (mrag[j]) {
                            uint32_t m[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int nb = tail - 4 * q;
                                m[q] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
                            }
                            c[j].x &= m[0]; c[j].y &= m[1]; c[j].z &= m[2]; c[j].w &= m[3];
                            p[j].x &= m[0]; p[j].y &= m[1]; p[j].z &= m[2]; p[j].w &= m[3];
                        }
                }
#pragma unroll
                for (int j = 0; j < MAXP; j++) {
                    if (j < merge_p && moff[j] >= 0) {
                        sad = sad4(c[j].x, p[j].x, sad);
                        sad = sad4(c[j].y, p[j].y, sad);
                        sad = sad4(c[j].z, p[j].z, sad);
                        sad = sad4(c[j].w, p[j].w, sad);
                        hist_word(c[j].x, base);
                        hist_word(c[j].y, base);
                        hist_word(c[j].z, base);
                        hist_word(c[j].w, base);
                    }
                }
                continue;
            }
        }
        for (int ri = 0; ri < nrows; ri++) {
        const int row = row_a + ri;
        const uint8_t *crow = cur + (size_t)row * pitch;
        const uint8_t *prow = prv + (size_t)row * pitch;
        if constexpr (U > 0) {
            const int nfull = nfull_all;
            int g0 = 0;
            for (; g0 + 32 * U <= nfull; g0 += 32 * U) {     // blocks of U full passes: all loads first
                uint4 c[U ? U : 1], p[U ? U : 1];
#pragma unroll
                for (int k = 0; k < U; k++) {
                    c[k] = ld_stream_u4(crow + (size_t)(g0 + k * 32 + lane) * 16);
                    p[k] = ld_stream_u4(prow + (size_t)(g0 + k * 32 + lane) * 16);
                }
#pragma unroll
                for (int k = 0; k < U; k++) {
                    sad = sad4(c[k].x, p[k].x, sad);
                    sad = sad4(c[k].y, p[k].y, sad);
                    sad = sad4(c[k].z, p[k].z, sad);
                    sad = sad4(c[k].w, p[k].w, sad);
                    hist_word(c[k].x, base);
                    hist_word(c[k].y, base);
                    hist_word(c[k].z, base);
                    hist_word(c[k].w, base);
                }
            }
            for (; g0 + 32 <= nfull; g0 += 32) {             // single full passes
                const uint4 c = ld_stream_u4(crow + (size_t)(g0 + lane) * 16);
                const uint4 p = ld_stream_u4(prow + (size_t)(g0 + lane) * 16);
                sad = sad4(c.x, p.x, sad);
                sad = sad4(c.y, p.y, sad);
                sad = sad4(c.z, p.z, sad);
                sad = sad4(c.w, p.w, sad);
                hist_word(c.x, base);
                hist_word(c.y, base);
                hist_word(c.z, base);
                hist_word(c.w, base);
            }
            if (g0 + lane < ngroups) {                       // the partial pass; its last group may be ragged
                uint4 c = ld_stream_u4(crow + (size_t)(g0 + lane) * 16);
                uint4 p = ld_stream_u4(prow + (size_t)(g0 + lane) * 16);
                if (tail && g0 + lane == ngroups - 1) {
                    uint32_t m[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int nb = tail - 4 * q;
                        m[q] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
                    }
                    c.x &= m[0]; c.y &= m[1]; c.z &= m[2]; c.w &= m[3];
                    p.x &= m[0]; p.y &= m[1]; p.z &= m[2]; p.w &= m[3];
                }
                sad = sad4(c.x, p.x, sad);
                sad = sad4(c.y, p.y, sad);
                sad = sad4(c.z, p.z, sad);
                sad = sad4(c.w, p.w, sad);
                hist_word(c.x, base);
                hist_word(c.y, base);
                hist_word(c.z, base);
                hist_word(c.w, base);
            }
            // (a lane sees at most 16 * 15 pixels of a row: no mid-row flush)
        } else {
        // four passes (64 bytes per lane) at a time: all eight 128-bit loads are issued before the first counter
        // update, which is what keeps enough bytes in flight per SM
        for (int g0 = 0; g0 < ngroups; g0 += 128) {
            uint4 c[4], p[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int g = g0 + k * 32 + lane;
                c[k] = make_uint4(0, 0, 0, 0);
                p[k] = c[k];
                if (g < ngroups) {
                    c[k] = ld_stream_u4(crow + (size_t)g * 16);
                    p[k] = ld_stream_u4(prow + (size_t)g * 16);
                }
            }
            if (tail && g0 + 128 >= ngroups) {  // the row's last group is in this block: mask bytes past the width
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (g0 + k * 32 + lane == ngroups - 1) {
                        uint32_t m[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int nb = tail - 4 * q;
                            m[q] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
                        }
                        c[k].x &= m[0]; c[k].y &= m[1]; c[k].z &= m[2]; c[k].w &= m[3];
                        p[k].x &= m[0]; p[k].y &= m[1]; p[k].z &= m[2]; p[k].w &= m[3];
                    }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (g0 + k * 32 >= ngroups) break;          // warp-uniform
                if (budget > 255 - 16) {                    // only reachable for rows wider than 8160 pixels
                    flush_counters_cold(warp_cnt, bhist, lane);
                    budget = 0;
                }
                if (g0 + k * 32 + lane < ngroups) {
                    sad = sad4(c[k].x, p[k].x, sad);
                    sad = sad4(c[k].y, p[k].y, sad);
                    sad = sad4(c[k].z, p[k].z, sad);
                    sad = sad4(c[k].w, p[k].w, sad);
                    hist_word(c[k].x, base);
                    hist_word(c[k].y, base);
                    hist_word(c[k].z, base);
                    hist_word(c[k].w, base);
                }
                budget += 16;
            }
        }
        }
        }
    }
    flush_counters(warp_cnt, bhist, lane);
    }

    // SAD: lane partials -> warp sum in 64 bit -> one global atomic per warp
    unsigned long long s64 = sad;
#pragma unroll
    for (int o = 16; o; o >>= 1) s64 += __shfl_xor_sync(0xffffffffu, s64, o);
    if (lane == 0 && s64) atomicAdd(&sad_out[f], s64);

    __syncthreads();
    uint32_t v = bhist[tid];
    if (tid == 0 && tail) v -= (uint32_t)(rows_last - rows_first) * (uint32_t)(16 - tail);  // masked bytes were counted as 0
    if (v) atomicAdd(&hist_out[(size_t)f * 256 + tid], v);
}

// Any alignment / pitch: byte loads, shared-memory atomics.  Correctness path for odd shapes.
__global__ void __launch_bounds__(256)
score_generic_kernel(const uint8_t *__restrict__ luma, int pitch, size_t frame_stride, int w, int h,
                     const uint8_t *__restrict__ prev0, int rows_per_block, int chunks_per_frame,
                     unsigned long long *__restrict__ sad_out, uint32_t *__restrict__ hist_out) {
    __shared__ uint32_t bhist[256];
    __shared__ unsigned long long bsad;
    const int tid = threadIdx.x;
    const int f = blockIdx.x / chunks_per_frame;
    const int chunk = blockIdx.x - f * chunks_per_frame;
    const int r0 = chunk * rows_per_block, r1 = min(h, r0 + rows_per_block);
    const uint8_t *cur = luma + (size_t)f * frame_stride;
    const uint8_t *prv = f ? cur - frame_stride : (prev0 ? prev0 : cur);
    bhist[tid] = 0;
    if (tid == 0) bsad = 0;
    __syncthreads();
    unsigned long long s = 0;
    for (int r = r0; r < r1; r++) {
        const uint8_t *cr = cur + (size_t)r * pitch, *pr = prv + (size_t)r * pitch;
        for (int x = tid; x < w; x += 256) {
            int a = cr[x], b = pr[x];
            atomicAdd(&bhist[a], 1u);
            s += (unsigned)abs(a - b);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0 && s) atomicAdd(&bsad, s);
    __syncthreads();
    if (tid == 0 && bsad) atomicAdd(&sad_out[f], bsad);
    if (bhist[tid]) atomicAdd(&hist_out[(size_t)f * 256 + tid], bhist[tid]);
}

int launch_score(const uint8_t *luma, int pitch, size_t frame_stride, int w, int h, const uint8_t *prev0,
                 int n_frames, uint64_t *sad, uint32_t *hist, cudaStream_t st) {
    if (!luma || !sad || !hist || w <= 0 || h <= 0 || pitch < w || n_frames <= 0) {
        set_error("vt_sad_hist_u8: bad arguments (w=%d h=%d pitch=%d n=%d)", w, h, pitch, n_frames);
        return VT_ERR_INVALID;
    }
    VT_CUDA(cudaMemsetAsync(sad, 0, sizeof(uint64_t) * (size_t)n_frames, st));
    VT_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 256 * (size_t)n_frames, st));
    const bool aligned = ((uintptr_t)luma % 16 == 0) && (pitch % 16 == 0) && (frame_stride % 16 == 0) &&
                         (!prev0 || (uintptr_t)prev0 % 16 == 0) && (((w + 15) & ~15) <= pitch);
    // Rows per block: as many consecutive rows per warp as fit the byte counters (a lane counts
    // 16*ceil(groups/32) pixels per row and may hold 255), so a block flushes its counters exactly once.
    const int px_per_lane_row = 16 * ((((w + 15) >> 4) + 31) / 32);
    int rpw = 255 / px_per_lane_row;
    if (rpw < 1) rpw = 1;
    if (rpw > 8) rpw = 8;
    // Row merging (see the kernel): k rows share ceil(k * rem / 32) passes for their leftover groups.  Pick the k <= 8
    // with the fewest passes per row whose pixel count per lane still fits the byte counters; among equals the largest
    // (measured on B200 at 1280 wide: k = 2 0.74, k = 6 0.82 of the roofline), walked in blocks of three passes
    // (blocks of three beat four at every width measured: 0.85 against 0.80 at 1920 and 3840).
    int merge_k = 1, merge_p = 0;
    {
        const int ng = (w + 15) >> 4, fullp = ng / 32, rem = ng - 32 * fullp;
        static const bool off = getenv("VT_SCORE_NO_MERGE") != nullptr;
        if (aligned && !off && fullp >= 1 && fullp + 1 <= 15) {
            double best = fullp + (rem ? 1.0 : 0.0);
            static const int k_force = getenv("VT_SCORE_K") ? atoi(getenv("VT_SCORE_K")) : 0;   // A/B measurements
            for (int k = 2; k <= 8; k++) {
                if (rem == 0 || (k_force && k != k_force)) continue;
                const int pk = (k * rem + 31) / 32;
                if (pk > 3 || 16 * (k * fullp + pk) > 255) continue;
                const double cost = (double)(k * fullp + pk) / k;
                if (cost <= best + 1e-9 && (cost < best - 1e-9 || merge_k > 1)) {   // ties: the larger group (measured)
                    best = cost; merge_k = k; merge_p = pk;
                }
            }
            if (rem == 0) {
                // no leftover groups, but walking k rows as one run of passes still keeps three loads in flight across
                // row ends: the largest k that fits, preferring runs that are whole blocks of three
                for (int k = 2; k <= 8; k++)
                    if (16 * k * fullp <= 255 && (merge_k == 1 || (k * fullp) % 3 == 0 || (merge_k * fullp) % 3 != 0)) merge_k = k;
                merge_p = 0;
            }
            if (merge_k > 1) {
                // a warp's last group may be cut by the picture's bottom edge; its rows then run unmerged
                const int G = 16 * (merge_k * fullp + merge_p), cut = (merge_k - 1) * 16 * (fullp + 1);
                int groups = 1;
                while ((groups + 1) * merge_k <= 8 && (groups + 1) * G <= 255 && groups * G + cut <= 255) groups++;
                if (cut > 255) merge_k = 1, merge_p = 0;
                else rpw = merge_k * groups;
            }
        }
    }
    const int rows_per_block = aligned ? rpw * SC_WARPS : 32;
    const int chunks = (h + rows_per_block - 1) / rows_per_block;
    // chunks per block: keep about six waves of blocks, fold the rest into fewer global atomics
    int cpb = 1;
    if (aligned) {
        const long long want = 6LL * sm_count() * 3;
        cpb = (int)std::max<long long>(1, std::min<long long>(chunks, (long long)chunks * n_frames / want));
    }
    const long long blocks = (long long)((chunks + cpb - 1) / cpb) * n_frames;
    if (blocks > 0x7fffffffLL) {
        set_error("vt_sad_hist_u8: batch too large");
        return VT_ERR_INVALID;
    }
    if (aligned) {
        const int ngroups = (w + 15) >> 4;
        const int npass = (ngroups + 31) / 32;
        int u = npass > 15 ? 0 : merge_k > 1 ? std::min(3, merge_k * (ngroups / 32)) : std::max(1, std::min(4, (w >> 4) / 32));
        static const int u_force = getenv("VT_SCORE_U") ? atoi(getenv("VT_SCORE_U")) : 0;   // A/B measurements
        if (u_force > 0 && u > 0) u = std::min(4, u_force);
        const bool rag = merge_k > 1 && (w & 15);
        typedef void (*kern_t)(const uint8_t *, int, size_t, int, int, const uint8_t *, int, int, int, int, int,
                               unsigned long long *, uint32_t *);
        static const kern_t table[2][5] = {
            {score_kernel<0, false>, score_kernel<1, false>, score_kernel<2, false>, score_kernel<3, false>, score_kernel<4, false>},
            {score_kernel<0, false>, score_kernel<1, true>, score_kernel<2, true>, score_kernel<3, true>, score_kernel<4, true>}};
        static bool attr_done_dev[VT_MAX_DEVICES] = {false};   // function attributes are per device
        bool &attr_done = attr_done_dev[current_device()];
        if (!attr_done) {
            for (int r = 0; r < 2; r++)
                for (int i = 0; i < 5; i++)
                    VT_CUDA(cudaFuncSetAttribute(table[r][i], cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
            attr_done = true;
        }
        kern_t kern = table[rag ? 1 : 0][u];
        kern<<<(unsigned)blocks, SC_THREADS, SC_SMEM, st>>>(luma, pitch, frame_stride, w, h, prev0, rows_per_block, chunks,
                                                            cpb, merge_k, merge_p, (unsigned long long *)sad, hist);
        VT_LAUNCHED("score_kernel");
    } else {
        score_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(luma, pitch, frame_stride, w, h, prev0, rows_per_block,
                                                              chunks, (unsigned long long *)sad, hist);
        VT_LAUNCHED("score_generic_kernel");
    }
    return VT_OK;
}

}  // namespace vt
