// vt_score.cu -- K3: per-frame 256-bin luma histogram + SAD(cur, prev).
//
// HBM-bound byte work: 2 bytes read per pixel (cur + prev luma), ~1 KB written per frame.
// The histogram is the hard part: shared-memory atomics serialise on flat picture areas (test patterns,
// letterboxing), so this kernel uses no atomics in its hot loop.  Every lane of a warp owns a private
// column of 256 one-byte counters (8 KB per warp, laid out counter[bin][lane] so a warp's 32 accesses
// fall in 8 words x 4 bytes); an update is LDS.U8 / IADD / STS.U8 on an address formed by one shift
// and one LOP3.  A lane sees at most 240 pixels between flushes, so a byte never wraps.  Flushes sum the
// 32 lane bytes of each bin with dp4a and add them to a per-block u32 histogram.
// SAD is __vabsdiffu4 + dp4a on the same 128-bit loads.
#include "vt_common.cuh"

namespace vt {

constexpr int SC_WARPS = 8;
constexpr int SC_THREADS = SC_WARPS * 32;
constexpr int SC_CNT_BYTES = 8192;                                   // per warp: 256 bins x 32 lanes x u8
constexpr int SC_SMEM = SC_WARPS * SC_CNT_BYTES + SC_CNT_BYTES + 1024;  // + alignment slack + block hist

__device__ __forceinline__ void cnt_inc(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    v += 1;
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// 4 pixels of one 32-bit word. base = (8 KB-aligned warp region) | lane, so OR-ing in bin*32 is exact.
__device__ __forceinline__ void hist_word(uint32_t w, uint32_t base) {
    cnt_inc(((w << 5) & 0x1FE0u) | base);
    cnt_inc(((w >> 3) & 0x1FE0u) | base);
    cnt_inc(((w >> 11) & 0x1FE0u) | base);
    cnt_inc(((w >> 19) & 0x1FE0u) | base);
}

// Warp-collective: add the lane-private byte counters into the block histogram and clear them.
__device__ __forceinline__ void flush_counters(uint32_t warp_cnt /* shared addr, 8 KB aligned */, uint32_t *bhist,
                                               int lane) {
    __syncwarp();
#pragma unroll 4
    for (int p = 0; p < 16; p++) {
        uint32_t a = warp_cnt + (uint32_t)(p * 32 + lane) * 16u;
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
        uint32_t s = __dp4a(v.x, 0x01010101u, 0u);
        s = __dp4a(v.y, 0x01010101u, s);
        s = __dp4a(v.z, 0x01010101u, s);
        s = __dp4a(v.w, 0x01010101u, s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (!(lane & 1) && s) atomicAdd(&bhist[p * 16 + (lane >> 1)], s);
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
    }
    __syncwarp();
}

// Aligned fast path: luma/prev base and pitch are multiples of 16 bytes.
__global__ void __launch_bounds__(SC_THREADS, 3)
score_kernel(const uint8_t *__restrict__ luma, int pitch, size_t frame_stride, int w, int h,
             const uint8_t *__restrict__ prev0, int rows_per_block, int chunks_per_frame,
             unsigned long long *__restrict__ sad_out, uint32_t *__restrict__ hist_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s0 = smem_u32(smem_raw);
    const uint32_t cnt0 = (s0 + SC_CNT_BYTES - 1) & ~(uint32_t)(SC_CNT_BYTES - 1);  // 8 KB aligned
    uint32_t *bhist = (uint32_t *)(smem_raw + (cnt0 - s0) + SC_WARPS * SC_CNT_BYTES);
    const uint32_t warp_cnt = cnt0 + warp * SC_CNT_BYTES;

    const int f = blockIdx.x / chunks_per_frame;
    const int chunk = blockIdx.x - f * chunks_per_frame;
    const int r0 = chunk * rows_per_block;
    const int r1 = min(h, r0 + rows_per_block);
    const uint8_t *cur = luma + (size_t)f * frame_stride;
    const uint8_t *prv = f ? cur - frame_stride : (prev0 ? prev0 : cur);

    bhist[tid] = 0;
#pragma unroll
    for (int p = 0; p < 16; p++)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(warp_cnt + (uint32_t)(p * 32 + lane) * 16u), "r"(0u)
                     : "memory");
    __syncthreads();

    const int ngroups = (w + 15) >> 4;           // 16-byte groups per row
    const int nit = (ngroups + 31) >> 5;         // warp passes per row
    const int tail = w & 15;                     // valid bytes in the last group (0 = full)
    const int nrows = (r1 - r0 - warp + SC_WARPS - 1) / SC_WARPS;  // rows this warp owns (may be <= 0)
    const int total = nrows > 0 ? nrows * nit : 0;
    const uint32_t base = warp_cnt | (uint32_t)lane;

    uint32_t sad = 0;
    int budget = 0;
    // software pipeline: loads for pass i+1 are in flight while pass i updates the counters
    uint4 c = make_uint4(0, 0, 0, 0), p = c;
    bool have = false;
    auto fetch = [&](int i, uint4 &cc, uint4 &pp) -> bool {
        int rr = i / nit;
        int g = (i - rr * nit) * 32 + lane;
        if (g >= ngroups) return false;
        size_t off = (size_t)(r0 + warp + rr * SC_WARPS) * pitch + (size_t)g * 16;
        cc = ld_stream_u4(cur + off);
        pp = ld_stream_u4(prv + off);
        if (tail && g == ngroups - 1) {  // zero the bytes past the display width in both operands
            uint32_t m[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int nb = tail - 4 * k;
                m[k] = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
            }
            cc.x &= m[0]; cc.y &= m[1]; cc.z &= m[2]; cc.w &= m[3];
            pp.x &= m[0]; pp.y &= m[1]; pp.z &= m[2]; pp.w &= m[3];
        }
        return true;
    };
    if (total > 0) have = fetch(0, c, p);
    for (int i = 0; i < total; i++) {
        uint4 cn = make_uint4(0, 0, 0, 0), pn = cn;
        bool have_n = false;
        if (i + 1 < total) have_n = fetch(i + 1, cn, pn);
        if (budget > 255 - 16) {
            flush_counters(warp_cnt, bhist, lane);
            budget = 0;
        }
        if (have) {
            sad = sad4(c.x, p.x, sad);
            sad = sad4(c.y, p.y, sad);
            sad = sad4(c.z, p.z, sad);
            sad = sad4(c.w, p.w, sad);
            hist_word(c.x, base);
            hist_word(c.y, base);
            hist_word(c.z, base);
            hist_word(c.w, base);
        }
        budget += 16;
        c = cn; p = pn; have = have_n;
    }
    flush_counters(warp_cnt, bhist, lane);

    // SAD: lane partials (<= 2^32) -> warp sum in 64 bit -> one global atomic per warp
    unsigned long long s64 = sad;
#pragma unroll
    for (int o = 16; o; o >>= 1) s64 += __shfl_xor_sync(0xffffffffu, s64, o);
    if (lane == 0 && s64) atomicAdd(&sad_out[f], s64);

    __syncthreads();
    uint32_t v = bhist[tid];
    if (tid == 0 && tail) v -= (uint32_t)(r1 - r0) * (uint32_t)(16 - tail);  // masked bytes were counted as 0
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
    // 32 rows per block: each warp owns 4 rows; enough blocks per frame to fill 148 SMs x 3 with small batches
    const int rows_per_block = 32;
    const int chunks = (h + rows_per_block - 1) / rows_per_block;
    const long long blocks = (long long)chunks * n_frames;
    if (blocks > 0x7fffffffLL) {
        set_error("vt_sad_hist_u8: batch too large");
        return VT_ERR_INVALID;
    }
    if (aligned) {
        static bool attr_done = false;
        if (!attr_done) {
            VT_CUDA(cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
            attr_done = true;
        }
        score_kernel<<<(unsigned)blocks, SC_THREADS, SC_SMEM, st>>>(luma, pitch, frame_stride, w, h, prev0,
                                                                  rows_per_block, chunks,
                                                                  (unsigned long long *)sad, hist);
        VT_LAUNCHED("score_kernel");
    } else {
        score_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(luma, pitch, frame_stride, w, h, prev0, rows_per_block,
                                                              chunks, (unsigned long long *)sad, hist);
        VT_LAUNCHED("score_generic_kernel");
    }
    return VT_OK;
}

}  // namespace vt
