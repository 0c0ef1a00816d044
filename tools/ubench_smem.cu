// Micro-benchmark: lane-private shared-memory counter updates, three ways (design input for vt_score.cu).
//   A: LDS.U8 / IADD / STS.U8 on byte counters laid out [bin][lane]           (quad lanes share a word)
//   B: LDS.U8 / IADD / STS.U8 on byte counters laid out [row][lane][4]         (bank == lane, conflict free)
//   C: ATOMS.ADD u32 of (1 << 8j) on packed counters laid out [row][lane]      (bank == lane, conflict free)
// Data: pseudo-random bytes (mode 0), a slow ramp (mode 1), constant (mode 2).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VAR>
__global__ void __launch_bounds__(256, 3) k(const uint32_t *__restrict__ data, int iters, uint32_t *out) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t base = ((s0 + 8191) & ~8191u) + warp * 8192;
    for (int i = threadIdx.x; i < 8 * 8192 / 4; i += 256) ((uint32_t *)(smem + (((s0 + 8191) & ~8191u) - s0)))[i] = 0;
    __syncthreads();
    const uint32_t *p = data + (size_t)blockIdx.x * 256 * 4 + threadIdx.x * 4;
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        uint4 v = *(const uint4 *)(p + (size_t)(it & 63) * gridDim.x * 1024);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (VAR == 0) {
                uint32_t b0 = base | lane;
#pragma unroll
                for (int kx = 0; kx < 4; kx++) {
                    uint32_t sh = kx == 0 ? (w[q] << 5) : (w[q] >> (8 * kx - 5));
                    uint32_t a = (sh & 0x1FE0u) | b0, t;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t) : "r"(a) : "memory");
                    t += 1;
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(t) : "memory");
                }
            } else {
                uint32_t R = w[q] & 0x3F3F3F3Fu, J = (w[q] >> 6) & 0x03030303u;
                uint32_t b0 = base | (lane << 2);
#pragma unroll
                for (int kx = 0; kx < 4; kx++) {
                    if (VAR == 1) {
                        // X = (r << 7) | j  built as two ops (layout rows are 128 B apart)
                        uint32_t r7 = kx == 0 ? (R << 7) : (kx == 1 ? (R >> 1) : (kx == 2 ? (R >> 9) : (R >> 17)));
                        uint32_t a = ((r7 & 0x1F80u) | b0) + ((J >> (8 * kx)) & 3u), t;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t) : "r"(a) : "memory");
                        t += 1;
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(t) : "memory");
                    } else {
                        uint32_t r7 = kx == 0 ? (R << 7) : (kx == 1 ? (R >> 1) : (kx == 2 ? (R >> 9) : (R >> 17)));
                        uint32_t a = (r7 & 0x1F80u) | b0;
                        uint32_t inc = 1u << (((J >> (8 * kx)) & 3u) * 8u);
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(inc) : "memory");
                    }
                }
            }
        }
        if ((it & 7) == 7) {   // keep byte counters from wrapping: clear (cost excluded from the comparison: same for all)
            __syncwarp();
            for (int i = lane; i < 8192 / 16; i += 32) *(uint4 *)(smem + (base - s0) + i * 16) = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }
    }
    acc += *(uint32_t *)(smem + (base - s0) + lane * 4);
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main() {
    const int blocks = 148 * 3, iters = 2048;
    size_t n = (size_t)blocks * 1024 * 64;
    uint32_t *h = (uint32_t *)malloc(n * 4), *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, blocks * 256 * 4);
    for (int mode = 0; mode < 3; mode++) {
        uint32_t s = 12345;
        for (size_t i = 0; i < n; i++) {
            uint32_t v;
            if (mode == 0) { s = s * 1664525u + 1013904223u; v = s ^ (s >> 13); }
            else if (mode == 1) { uint32_t b = (uint32_t)((i * 4) / 9) & 0xFF; v = b | (b << 8) | (b << 16) | (b << 24); }
            else v = 0x80808080u;
            h[i] = v;
        }
        cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
        for (int var = 0; var < 3; var++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            auto run = [&](int it) {
                if (var == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728); k<0><<<blocks, 256, 73728>>>(d, it, o); }
                if (var == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728); k<1><<<blocks, 256, 73728>>>(d, it, o); }
                if (var == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728); k<2><<<blocks, 256, 73728>>>(d, it, o); }
            };
            run(64); cudaDeviceSynchronize();
            cudaEventRecord(a); run(iters); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double px = (double)blocks * 256 * 16 * iters;
            printf("mode %d var %c: %.3f ms  %.1f Gpx/s  %.2f px/clk/SM @1.9GHz  err=%s\n", mode, "ABC"[var], ms, px / ms / 1e6,
                   px / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
