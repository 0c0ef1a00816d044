// Micro-benchmark: issue rates of the integer instructions the scaler is made of, per SM (B200, sm_100a).
// Each variant runs ILP independent dependency chains per thread for `iters` iterations; the result is
// warp-instructions per cycle per SM at full occupancy (2048 threads/SM) and at the scaler's occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c) { int d; asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp4a_u(uint32_t a, uint32_t b, int c) { int d; asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int imad(int a, int b, int c) { int d; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t shf(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int vmin(int a, int b) { int d; asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t i2ip(int a, int b) { uint32_t d; asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int ILP = 8;
// MODE: 0 IDP.2A, 1 IMAD, 2 SHF, 3 VIMNMX, 4 I2IP, 5 IDP.2A+SHF alternating, 6 IMAD+SHF alternating, 7 IDP.4A, 8 FFMA,
//       9 IDP+IMAD alternating, 10 IMAD+FFMA alternating, 11 IDP+SHF+SHF (1:2)
struct UTab { int t[64]; };
template <int MODE>
__global__ void k(int iters, uint32_t seed, uint32_t *out, const __grid_constant__ UTab ut) {
    uint32_t x[ILP], y[ILP];
    float fx[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed + threadIdx.x * 7 + i; y[i] = seed * 3 + i; fx[i] = (float)i; }
    const uint32_t c0 = seed | 0x00010001u;
    for (int it = 0; it < iters; it++) {
        const int ui = (it & 7) * 8;               // uniform, dynamic: the loads below become LDCU + UR operands
        const int uc0 = ut.t[ui], uc1 = ut.t[ui + 1], uc2 = ut.t[ui + 2], uc3 = ut.t[ui + 3];
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) x[i] = dp2a_lo(c0, y[i], x[i]);
                if (MODE == 1) x[i] = imad(x[i], c0, y[i]);
                if (MODE == 2) x[i] = shf(x[i], y[i], c0);
                if (MODE == 3) x[i] = vmin(x[i], y[i]);
                if (MODE == 4) x[i] = i2ip(x[i], y[i]);
                if (MODE == 5) { x[i] = dp2a_lo(c0, y[i], x[i]); y[i] = shf(y[i], c0, c0); }
                if (MODE == 6) { x[i] = imad(x[i], c0, c0); y[i] = shf(y[i], c0, c0); }
                if (MODE == 7) x[i] = dp4a_u(c0, y[i], x[i]);
                if (MODE == 8) fx[i] = ffma(fx[i], 1.0001f, 0.5f);
                if (MODE == 9) { x[i] = dp2a_lo(c0, y[i], x[i]); y[i] = imad(y[i], c0, c0); }
                if (MODE == 10) { x[i] = imad(x[i], c0, c0); fx[i] = ffma(fx[i], 1.0001f, 0.5f); }
                if (MODE == 12) x[i] = imad(y[i], (u & 1) ? ((u & 2) ? uc0 : uc1) : ((u & 2) ? uc2 : uc3), x[i]);
                if (MODE == 13) { x[i] = imad(y[i], (u & 1) ? uc0 : uc1, x[i]); y[i] = shf(y[i], c0, c0); }
                if (MODE == 14) { x[i] = dp2a_lo(c0, y[i], x[i]); x[i] = dp2a_lo(c0 + 1, y[i], x[i]); x[i] = (int)x[i] >> 7; x[i] = vmin(x[i], 32767); y[i] = shf(y[i], c0, c0); }
                if (MODE == 11) { x[i] = dp2a_lo(c0, y[i], x[i]); y[i] = shf(y[i], c0, c0); y[i] = shf(y[i], c0, x[i]); }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += x[i] + y[i] + (uint32_t)fx[i];
    if (r == 0x12345678u) out[0] = r;
}

// The scaler's essential instruction mix with none of its bookkeeping: per iteration three "source rows"
// (12 LDS + 8 funnel shifts + 28 dp2a + 8 shifts + 8 clamps each) and two "output rows" (48 IMAD + 8 shifts + 4 packs
// + 4 16-bit stores each), eight independent columns per thread like the real kernel.  Tells how many of these
// instructions per clock an SM can issue at best.
__global__ void __launch_bounds__(128) scaler_mix(int iters, uint32_t seed, unsigned short *out) {
    __shared__ uint32_t tile[128 * 16];
    for (int i = threadIdx.x; i < 128 * 16; i += 128) tile[i] = seed * i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile) + threadIdx.x * 4;
    uint32_t ca[4][3], cb[4][4], sh = ((seed + threadIdx.x) & 3) * 8;
    int m[6][8], coef[6];
    for (int g = 0; g < 4; g++) { for (int i = 0; i < 3; i++) ca[g][i] = seed + g + i + threadIdx.x; for (int i = 0; i < 4; i++) cb[g][i] = seed * 3 + g + i + threadIdx.x; }
    for (int k = 0; k < 6; k++) { coef[k] = (int)seed + k; for (int c = 0; c < 8; c++) m[k][c] = k + c; }
    unsigned short *dst = out + (size_t)(blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 * 32 + (threadIdx.x & 31);   // 64 B per warp store
    for (int it = 0; it < iters; it++) {
        const uint32_t rowbase = base + (uint32_t)(it & 7) * 16;      // loop-variant addresses: nothing can be hoisted
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int slot = r * 2;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                uint32_t w0, w1, w2;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(rowbase + (uint32_t)(r * 512 + g * 12)));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(rowbase + (uint32_t)(r * 512 + g * 12 + 4)));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(rowbase + (uint32_t)(r * 512 + g * 12 + 8)));
                const uint32_t a0 = shf(w0, w1, sh), a1 = shf(w1, w2, sh);
                int va = dp2a_lo(ca[g][0], a0, 0); va = dp2a_lo(ca[g][1], a0, va); va = dp2a_lo(ca[g][2], a1, va);
                int vb = dp2a_lo(cb[g][0], a0, 0); vb = dp2a_lo(cb[g][1], a0, vb); vb = dp2a_lo(cb[g][2], a1, vb); vb = dp2a_lo(cb[g][3], a1, vb);
                m[slot][2 * g] = vmin(va >> 7, 32767);
                m[slot][2 * g + 1] = vmin(vb >> 7, 32767);
            }
            if (r != 1) {   // two of three rows complete an output row
                int acc[8];
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = 1 << 18;
#pragma unroll
                for (int j = 0; j < 6; j++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[c] = imad(m[(slot + 1 + j) % 6][c], coef[j], acc[c]);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const uint32_t pk = i2ip(acc[2 * g + 1] >> 19, acc[2 * g] >> 19);
                    asm volatile("st.global.u16 [%0], %1;" ::"l"(dst + g * 32), "h"((unsigned short)pk) : "memory");
                }
            }
        }
    }
}

void run_scaler_mix(int blocks_per_sm) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned short *out; cudaMalloc(&out, (size_t)sms * blocks_per_sm * 128 * 8 * 2 + 64);
    const int iters = 4000;
    scaler_mix<<<sms * blocks_per_sm, 128>>>(10, 1, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    scaler_mix<<<sms * blocks_per_sm, 128>>>(iters, 1, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("scaler_mix: %s\n", cudaGetErrorString(err));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_iter = 3 * 64 + 2 * 64;                       // instructions per warp per iteration (essentials only)
    const double winst = (double)iters * per_iter * 4 * blocks_per_sm;   // per SM
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double px_per_clk = (double)iters * 2 * 8 * 128 * blocks_per_sm / cycles;   // output pixels per clock per SM
    printf("(%.3f ms) scaler essential mix        warps/SM %2d : %.3f warp-instr/clk/SM (%.3f per SMSP), %.2f output px/clk/SM\n",
           ms, 4 * blocks_per_sm, winst / cycles, winst / cycles / 4, px_per_clk);
    cudaFree(out);
}

template <int MODE>
void run(const char *name, int per_iter, int threads, int blocks_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; cudaMalloc(&out, 4);
    const int iters = 2000;
    UTab ut; for (int i = 0; i < 64; i++) ut.t[i] = 3 + i;
    k<MODE><<<sms * blocks_per_sm, threads>>>(10, 1, out, ut);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, threads>>>(iters, 1, out, ut);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winst = (double)iters * 8 * ILP * per_iter * (threads / 32) * blocks_per_sm;   // warp-instr per SM
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s warps/SM %2d : %.3f warp-instr/clk/SM  (%.3f per SMSP)\n", name, threads / 32 * blocks_per_sm, winst / cycles,
           winst / cycles / 4);
    cudaFree(out);
}

int main() {
    run_scaler_mix(4);
    run_scaler_mix(5);
    run_scaler_mix(8);
    for (int w = 0; w < 2; w++) {
        const int threads = w ? 128 : 1024, bps = w ? 5 : 2;
        run<0>("IDP.2A", 1, threads, bps);
        run<7>("IDP.4A", 1, threads, bps);
        run<1>("IMAD", 1, threads, bps);
        run<8>("FFMA", 1, threads, bps);
        run<2>("SHF", 1, threads, bps);
        run<3>("VIMNMX", 1, threads, bps);
        run<4>("I2IP", 1, threads, bps);
        run<5>("IDP.2A + SHF", 2, threads, bps);
        run<11>("IDP.2A + 2 SHF", 3, threads, bps);
        run<6>("IMAD + SHF", 2, threads, bps);
        run<9>("IDP.2A + IMAD", 2, threads, bps);
        run<10>("IMAD + FFMA", 2, threads, bps);
        run<12>("IMAD (UR coefficient)", 1, threads, bps);
        run<13>("IMAD (UR coef) + SHF", 2, threads, bps);
        run<14>("2 IDP + SHR + MIN + SHF", 5, threads, bps);
    }
    return 0;
}
