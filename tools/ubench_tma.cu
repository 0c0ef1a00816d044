// Micro-benchmark: how fast can per-warp TMA rings pull strided strips of a batch of pitch-linear luma planes
// into shared memory?  (Design input for vt_scale_pair.cu: box width/height, stages, warps per SM.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/ubench_tma tools/ubench_tma.cu -lcuda   (build outside the tree; -cudart shared keeps the static runtime out)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *tmap, uint32_t bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Args {
    const uint8_t *src; int pitch; size_t frame_stride;
    int n_frames, rows, n_strips, strip_stride, box_w, box_h, nst, seg_rows, n_segs, warp_smem, stage_bytes, elem;   // elem: bytes per tensor element
    int xoff;   // added to every strip's first byte (alignment experiments)
    int mode;   // 0 TMA tensor boxes, 1 one bulk copy per row, 2 LDG.128 by the warp itself
    unsigned long long *sink;
};

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmap, const Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t wsm = smem_u32(smem + (size_t)warp * a.warp_smem);
    const uint32_t bar0 = wsm + a.nst * a.stage_bytes;
    if (lane == 0) { for (int s = 0; s < a.nst; s++) mbar_init(bar0 + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    uint32_t phases = 0, acc = 0;
    const uint32_t tx = (uint32_t)a.box_w * a.box_h;
    const int per_frame = a.n_strips * a.n_segs;
    const long long total = (long long)per_frame * a.n_frames, nwarps = (long long)gridDim.x * 4;
    for (long long it = (long long)blockIdx.x * 4 + warp; it < total; it += nwarps) {
        const int f = (int)(it / per_frame), rem = (int)(it - (long long)f * per_frame);
        const int seg = rem / a.n_strips, strip = rem - seg * a.n_strips;
        const int r0 = seg * a.seg_rows, r1 = min(a.rows, r0 + a.seg_rows);
        const int nloads = (r1 - r0 + a.box_h - 1) / a.box_h;
        const int x0 = strip * a.strip_stride + (a.xoff < 0 ? 16 * ((strip * 5 + seg) % 8) : a.xoff);
        auto issue = [&](int s, int ld) {
            const uint32_t dst = wsm + s * a.stage_bytes, bar = bar0 + 8 * s;
            if (a.mode == 0) {
                if (lane == 0) { mbar_expect_tx(bar, tx); tma_load_3d(dst, &tmap, bar, x0 / a.elem, r0 + ld * a.box_h, f); }
            } else if (a.mode == 1) {
                if (lane == 0) mbar_expect_tx(bar, tx);
                __syncwarp();
                if (lane < a.box_h)
                    bulk_load_1d(dst + lane * a.box_w, a.src + (size_t)f * a.frame_stride + (size_t)min(r0 + ld * a.box_h + lane, a.rows - 1) * a.pitch + x0, a.box_w, bar);
            }
        };
        __syncwarp();
        for (int s = 0; s < min(a.nst, nloads); s++) issue(s, s);
        int s = 0;
        for (int ld = 0; ld < nloads; ld++) {
            if (a.mode == 2) {
                const uint8_t *p = a.src + (size_t)f * a.frame_stride + (size_t)(r0 + ld * a.box_h) * a.pitch + x0;
                const int per_row = a.box_w / 16, n = per_row * a.box_h;
                for (int i = lane; i < n; i += 32) {
                    const int rr = i / per_row, cc = i - rr * per_row;
                    uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + (size_t)rr * a.pitch + cc * 16));
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                }
            } else {
                mbar_wait(bar0 + 8 * s, (phases >> s) & 1u); phases ^= 1u << s;
                uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(wsm + s * a.stage_bytes + lane * 4));
                acc += v;
                __syncwarp();
                if (ld + a.nst < nloads) issue(s, ld + a.nst);
            }
            s = s + 1 == a.nst ? 0 : s + 1;
        }
    }
    if (acc == 0x12345678u) a.sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int W = 1920, H = 1080, pitch = 2048, F = 256;
    const size_t fs = (size_t)pitch * (H + H / 2);
    uint8_t *src; cudaMalloc(&src, fs * F); cudaMemset(src, 1, fs * F);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    printf("mode box_w box_h nst strip_stride warps/SM l2promo : GB/s of unique bytes (strip_stride x rows)\n");
    struct Cfg { int mode, box_w, box_h, nst, stride, bps, promo, xoff; };
    const Cfg cfgs[] = {
        {0, 256, 6, 4, 192, 5, 1, 0}, {0, 256, 6, 2, 192, 5, 1, 0}, {0, 256, 6, 3, 192, 5, 1, 0}, {0, 256, 12, 2, 192, 5, 1, 0}, {0, 256, 12, 3, 192, 5, 1, 0},
        {0, 256, 12, 2, 192, 5, 1, 16}, {0, 256, 12, 2, 192, 5, 1, -1}, {0, 448, 12, 2, 384, 5, 1, 0}, {0, 448, 12, 2, 384, 5, 1, 16}, {0, 448, 12, 2, 384, 5, 1, -1},
        {0, 448, 12, 2, 384, 4, 1, -1}, {0, 448, 12, 3, 384, 4, 1, -1}, {0, 448, 18, 2, 384, 4, 1, -1}, {0, 448, 24, 2, 384, 2, 1, -1}, {0, 448, 6, 2, 384, 5, 1, -1},
        {0, 512, 12, 2, 384, 4, 1, -1}, {1, 448, 12, 2, 384, 5, 1, -1}, {0, 448, 8, 2, 384, 5, 1, -1}, {0, 448, 8, 3, 384, 5, 1, -1},
    };
    for (const Cfg &c : cfgs) {
        Args a; a.src = src; a.pitch = pitch; a.frame_stride = fs; a.n_frames = F; a.rows = H;
        a.strip_stride = c.stride; a.n_strips = W / c.stride; a.box_w = c.box_w; a.box_h = c.box_h; a.nst = c.nst;
        a.elem = c.box_w > 256 ? 4 : 1; a.mode = c.mode; a.xoff = c.xoff; a.sink = sink;
        a.stage_bytes = (c.box_w * c.box_h + 127) & ~127; a.warp_smem = (a.stage_bytes * c.nst + 64 + 127) & ~127;
        const long long warps = (long long)sms * c.bps * 4;
        long long n_segs = (8 * warps + (long long)F * a.n_strips - 1) / ((long long)F * a.n_strips);
        if (n_segs < 1) n_segs = 1;
        a.seg_rows = (int)((H + n_segs - 1) / n_segs); a.seg_rows = (a.seg_rows + c.box_h - 1) / c.box_h * c.box_h;
        a.n_segs = (H + a.seg_rows - 1) / a.seg_rows;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)(W / a.elem), (cuuint64_t)H, (cuuint64_t)F};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
        cuuint32_t box[3] = {(cuuint32_t)(c.box_w / a.elem), (cuuint32_t)c.box_h, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&tm, a.elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, src, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int smem = a.warp_smem * 4;
        if (smem * c.bps > 225 * 1024) { printf("skip (smem)\n"); continue; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms * c.bps, 128, smem>>>(tm, a);
        cudaEventRecord(e0);
        for (int rep = 0; rep < 5; rep++) k<<<sms * c.bps, 128, smem>>>(tm, a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        cudaError_t err = cudaGetLastError();
        const double bytes = (double)a.n_strips * c.stride * H * F;
        printf("%d %4d %3d %d %4d %2d %d xoff %2d: %7.1f GB/s  (%.3f ms)%s\n", c.mode, c.box_w, c.box_h, c.nst, c.stride, c.bps * 4, c.promo, c.xoff, bytes / ms / 1e6, ms,
               err ? cudaGetErrorString(err) : "");
    }
    return 0;
}
