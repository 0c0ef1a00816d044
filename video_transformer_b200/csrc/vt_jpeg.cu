// vt_jpeg.cu -- baseline JPEG encoder for the upload-size reducer.
//
// The reference reduces an upload with `ffmpeg -vf scale=-2:360 -c:v libx264 -crf 28 ...`
// (/root/reference/src/analyzer/content_analyzer.py:193-217).  A B200 has no video encoder and the image has no
// software one, so the reducer's pictures leave the GPU as Motion-JPEG: ITU-T T.81 baseline sequential DCT, 4:2:0,
// Annex-K Huffman tables, one restart interval per MCU row.  Arithmetic is the Independent JPEG Group's (integer
// "islow" forward DCT, divisor = 8 * quantval with round-half-away, jpeg_quality_scaling); the CPU statement of the
// same algorithm is oracle/jpeg_oracle.c, which tests pin byte-for-byte against libjpeg-turbo.
//
// Kernels
//   jpeg_row_kernel       one thread block per (picture, MCU row).  Phase 1: every thread transforms and quantises 8x8
//                         blocks into shared memory (zigzag int16).  Phase 2: each thread sizes the entropy code of a
//                         contiguous run of blocks; a block-wide exclusive scan turns sizes into bit offsets.  Phase 3:
//                         threads write their codes into a shared bit buffer (atomicOr on the shared 32-bit words a run
//                         boundary falls in).  Phase 4: 0xFF byte stuffing with a second scan, stuffed bytes go to the
//                         row's scratch area, the row's size to row_bytes.  A restart interval per MCU row is what makes
//                         rows independent: DC prediction restarts and every row is byte aligned.
//   jpeg_offsets_kernel   picture sizes and exclusive offsets (header + rows + RSTn markers + EOI), one block.
//   jpeg_assemble_kernel  copies header, rows and markers of every picture to its place in the packed output.
// The work is byte/bit manipulation with tiny traffic (0.35 MB in, ~0.03 MB out per 640x360 picture): latency- and
// issue-bound, far from any HBM roofline; it runs once per SECOND of video.
#include <algorithm>
#include <cstring>
#include <vector>

#include "vt_common.cuh"

namespace vt {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// ITU-T T.81 Annex K.1 (luminance, chrominance), natural order
const uint8_t kQLum[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                           69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55, 64,
                           81, 104, 113, 92, 49, 64, 78,  87,  103, 121, 120, 101, 72, 92,  95,  98,  112, 100, 103, 99};
const uint8_t kQChr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                           99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                           99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 typical Huffman tables: code counts per length, then symbols in code order
const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
    0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
    0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94,
    0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8,
    0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};
const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
    0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
    0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
    0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92,
    0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};

// (length << 16) | code for every symbol of a table
void build_table(const uint8_t *bits, const uint8_t *vals, uint32_t *out256) {
    memset(out256, 0, 256 * sizeof(uint32_t));
    unsigned code = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1]; i++) out256[vals[k++]] = ((uint32_t)l << 16) | code++;
        code <<= 1;
    }
}

constexpr int JT = 256;                 // threads per block of the row kernel

struct JpegArgs {
    const uint8_t *src;                 // picture f: Y plane (w x h, tight), U, V (cw x ch each) at src + f * frame_stride
    unsigned long long frame_stride;
    int w, h, cw, ch, mw, mh;           // picture size, chroma size, MCUs per row / column
    int expand;                         // 1: limited-range video samples are expanded to full range first
    int cap_words;                      // bit buffer of one MCU row, in 32-bit words
    int row_cap;                        // scratch bytes per (picture, row)
    uint8_t *scratch;                   // [n_frames][mh][row_cap]
    uint32_t *row_bytes;                // [n_frames][mh]
    int *status;                        // 0 ok, 1 an MCU row outgrew its bit buffer, 2 the output buffer is too small
    uint16_t qdiv[2][64];               // 8 * quantval, zigzag order: [0] luminance, [1] chrominance
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c (JDCT_ISLOW): one 1-D pass over eight values spaced `S` apart
template <int S, bool FIRST>
__device__ __forceinline__ void fdct_1d(int *p) {
    constexpr int CB = 13, P1 = 2, SH = FIRST ? CB - P1 : CB + P1;
    int t0 = p[0] + p[7 * S], t7 = p[0] - p[7 * S], t1 = p[S] + p[6 * S], t6 = p[S] - p[6 * S];
    int t2 = p[2 * S] + p[5 * S], t5 = p[2 * S] - p[5 * S], t3 = p[3 * S] + p[4 * S], t4 = p[3 * S] - p[4 * S];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    if (FIRST) {
        p[0] = (t10 + t11) << P1;
        p[4 * S] = (t10 - t11) << P1;
    } else {
        p[0] = descale(t10 + t11, P1);
        p[4 * S] = descale(t10 - t11, P1);
    }
    int z1 = (t12 + t13) * 4433;
    p[2 * S] = descale(z1 + t13 * 6270, SH);
    p[6 * S] = descale(z1 + t12 * (-15137), SH);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    t4 *= 2446;
    t5 *= 16819;
    t6 *= 25172;
    t7 *= 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 = z3 * -16069 + z5;
    z4 = z4 * -3196 + z5;
    p[7 * S] = descale(t4 + z1 + z3, SH);
    p[5 * S] = descale(t5 + z2 + z4, SH);
    p[3 * S] = descale(t6 + z2 + z3, SH);
    p[S] = descale(t7 + z1 + z4, SH);
}

__device__ __forceinline__ int nbits_of(int v) { return 32 - __clz(v < 0 ? -v : v); }

// Walks the entropy code of one block.  EMIT = false: returns its length in bits.  EMIT = true: writes the bits at
// bit position `pos` of `buf` (big-endian 32-bit words) and returns the position after the block.
// zz[k * stride]: coefficient k of the block (the row's coefficients are stored k-major, see jpeg_row_kernel).
template <bool EMIT>
__device__ __forceinline__ uint32_t walk_block(const int16_t *zz, int stride, int pred, const uint32_t *dc_tab,
                                               const uint32_t *ac_tab, uint32_t *buf, uint32_t pos) {
    uint32_t widx = pos >> 5, acc = 0;
    int fill = (int)(pos & 31u);
    uint32_t total = 0;
    auto put = [&](uint32_t bits, int len) {
        if (!EMIT) {
            total += (uint32_t)len;
            return;
        }
        const int space = 32 - fill;
        if (len <= space) {
            acc |= bits << (space - len);       // len >= 1, so the shift is at most 31
            fill += len;
            if (fill == 32) {
                atomicOr(&buf[widx++], acc);
                acc = 0;
                fill = 0;
            }
        } else {
            const int rest = len - space;
            atomicOr(&buf[widx++], acc | (bits >> rest));
            acc = bits << (32 - rest);
            fill = rest;
        }
    };
    const int diff = (int)zz[0] - pred;
    int n = nbits_of(diff);
    uint32_t e = dc_tab[n];
    put(((e & 0xFFFFu) << n) | ((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << n) - 1u)), (int)(e >> 16) + n);
    int run = 0;
#pragma unroll 1
    for (int k = 1; k < 64; k++) {
        const int v = zz[k * stride];
        if (v == 0) {
            run++;
            continue;
        }
        while (run > 15) {
            e = ac_tab[0xF0];
            put(e & 0xFFFFu, (int)(e >> 16));
            run -= 16;
        }
        n = nbits_of(v);
        e = ac_tab[(run << 4) | n];
        put(((e & 0xFFFFu) << n) | ((uint32_t)(v < 0 ? v - 1 : v) & ((1u << n) - 1u)), (int)(e >> 16) + n);
        run = 0;
    }
    if (run > 0) {
        e = ac_tab[0];
        put(e & 0xFFFFu, (int)(e >> 16));
    }
    if (EMIT) {
        if (fill) atomicOr(&buf[widx], acc);
        return (widx << 5) + (uint32_t)fill;
    }
    return total;
}

// exclusive scan of one value per thread over the block (JT threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < JT / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        if (lane < JT / 32) warp_sums[lane] = s;       // inclusive sums of the warps
    }
    __syncthreads();
    const uint32_t base = warp ? warp_sums[warp - 1] : 0;
    *total = warp_sums[JT / 32 - 1];
    __syncthreads();                                   // warp_sums may be reused by the next scan
    return base + inc - v;
}

__constant__ uint32_t c_huff[4][256];   // DC luminance, AC luminance, DC chrominance, AC chrominance

__global__ void __launch_bounds__(JT) jpeg_row_kernel(const __grid_constant__ JpegArgs a) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int row = blockIdx.x, f = blockIdx.y;
    const int nb = a.mw * 6;                             // 8x8 blocks of this MCU row, in scan order
    // coefficients K-MAJOR: coef[k * nbp + b].  Threads walk the same k of neighbouring blocks, so a block-major layout
    // (128 bytes between lanes) put all 32 lanes on one bank: 93 % of the kernel's shared wavefronts were conflicts
    // (profiles/r02_ncu_jpeg.csv, first build); k-major they are adjacent halfwords.
    const int nbp = (nb + 1) & ~1;
    int16_t *coef = reinterpret_cast<int16_t *>(sm);                                   // [64][nbp]
    uint32_t *bitbuf = reinterpret_cast<uint32_t *>(sm + (size_t)nbp * 128);           // [cap_words]
    uint32_t *tabs = bitbuf + a.cap_words;                                             // [4][256]
    __shared__ uint32_t warp_sums[JT / 32];
    for (int i = threadIdx.x; i < 4 * 256; i += JT) tabs[i] = (&c_huff[0][0])[i];
    for (int i = threadIdx.x; i < a.cap_words; i += JT) bitbuf[i] = 0;
    const uint8_t *yp = a.src + (size_t)f * a.frame_stride;
    const uint8_t *up = yp + (size_t)a.w * a.h;
    const uint8_t *vp = up + (size_t)a.cw * a.ch;

    // ---- phase 1: forward DCT + quantisation, one 8x8 block per thread per pass
    for (int b = threadIdx.x; b < nb; b += JT) {
        const int m = b / 6, k = b - 6 * m;
        const uint8_t *plane;
        int pw, ph, bx, by, chroma;
        if (k < 4) {
            plane = yp; pw = a.w; ph = a.h; bx = 2 * m + (k & 1); by = 2 * row + (k >> 1); chroma = 0;
        } else {
            plane = k == 4 ? up : vp; pw = a.cw; ph = a.ch; bx = m; by = row; chroma = 1;
        }
        int d[64];
        const bool inside = bx * 8 + 8 <= pw && by * 8 + 8 <= ph && (pw & 3) == 0 && ((size_t)plane & 3) == 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int y = min(by * 8 + r, ph - 1);
            const uint8_t *p = plane + (size_t)y * pw;
            if (inside) {                                                               // rows are 4-byte aligned here
                const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(p + bx * 8));
                const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t *>(p + bx * 8 + 4));
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    d[r * 8 + c] = (int)((w0 >> (8 * c)) & 0xFFu);
                    d[r * 8 + 4 + c] = (int)((w1 >> (8 * c)) & 0xFFu);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; c++) d[r * 8 + c] = (int)__ldg(p + min(bx * 8 + c, pw - 1));
            }
        }
#pragma unroll
        for (int i = 0; i < 64; i++) {
            int v = d[i];
            if (a.expand) {
                v = chroma ? (((v - 128) * 18652 + 8192) >> 14) + 128 : ((v - 16) * 19077 + 8192) >> 14;
                v = min(max(v, 0), 255);
            }
            d[i] = v - 128;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) fdct_1d<1, true>(d + 8 * i);
#pragma unroll
        for (int i = 0; i < 8; i++) fdct_1d<8, false>(d + i);
        const uint16_t *qd = a.qdiv[chroma];
        int16_t *o = coef + b;
#pragma unroll
        for (int kz = 0; kz < 64; kz++) {
            constexpr uint8_t zig[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
            const int t = d[zig[kz]];
            const unsigned qv = qd[kz];
            const unsigned mag = ((unsigned)(t < 0 ? -t : t) + (qv >> 1)) / qv;
            o[kz * nbp] = (int16_t)(t < 0 ? -(int)mag : (int)mag);
        }
    }
    __syncthreads();

    // ---- phase 2: code sizes.  Thread t owns blocks [t*per, t*per + per) so that offsets follow from one scan
    const int per = (nb + JT - 1) / JT;
    const int b0 = min(nb, (int)threadIdx.x * per), b1 = min(nb, b0 + per);
    auto pred_of = [&](int b) -> int {                   // DC predictor of block b: previous block of its component
        const int m = b / 6, k = b - 6 * m;
        if (k >= 1 && k <= 3) return coef[b - 1];
        if (m == 0) return 0;                            // restart at the beginning of every MCU row
        return coef[b - (k == 0 ? 3 : 6)];
    };
    uint32_t mine = 0;
    for (int b = b0; b < b1; b++) {
        const int chroma = (b % 6) >= 4;
        mine += walk_block<false>(coef + b, nbp, pred_of(b), tabs + (chroma ? 512 : 0), tabs + (chroma ? 768 : 256),
                                  nullptr, 0);
    }
    uint32_t total_bits;
    uint32_t pos = block_exclusive_scan(mine, warp_sums, &total_bits);
    const uint32_t n_bytes = (total_bits + 7) >> 3;
    const size_t row_index = (size_t)f * a.mh + row;
    if (n_bytes > (uint32_t)a.cap_words * 4u) {
        if (threadIdx.x == 0) {
            atomicMax(a.status, 1);
            a.row_bytes[row_index] = 0;
        }
        return;
    }
    // ---- phase 3: emit
    for (int b = b0; b < b1; b++) {
        const int chroma = (b % 6) >= 4;
        pos = walk_block<true>(coef + b, nbp, pred_of(b), tabs + (chroma ? 512 : 0), tabs + (chroma ? 768 : 256),
                               bitbuf, pos);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (total_bits & 7u)) {         // pad the last byte with ones
        const uint32_t padn = 8u - (total_bits & 7u);
        const uint32_t shift = 32u - (total_bits & 31u) - padn;
        bitbuf[total_bits >> 5] |= ((1u << padn) - 1u) << shift;
    }
    __syncthreads();
    // ---- phase 4: byte stuffing.  Thread t owns bytes [t*bper, t*bper + bper)
    const uint32_t bper = (n_bytes + JT - 1) / JT;
    const uint32_t s0 = min(n_bytes, threadIdx.x * bper), s1 = min(n_bytes, s0 + bper);
    auto byte_at = [&](uint32_t i) -> uint32_t { return (bitbuf[i >> 2] >> (24u - 8u * (i & 3u))) & 0xFFu; };
    uint32_t ff = 0;
    for (uint32_t i = s0; i < s1; i++) ff += byte_at(i) == 0xFFu;
    uint32_t ff_total;
    const uint32_t ff_before = block_exclusive_scan(ff, warp_sums, &ff_total);
    const uint32_t out_bytes = n_bytes + ff_total;
    if (out_bytes > (uint32_t)a.row_cap) {
        if (threadIdx.x == 0) {
            atomicMax(a.status, 1);
            a.row_bytes[row_index] = 0;
        }
        return;
    }
    uint8_t *o = a.scratch + row_index * (size_t)a.row_cap + s0 + ff_before;
    for (uint32_t i = s0; i < s1; i++) {
        const uint32_t v = byte_at(i);
        *o++ = (uint8_t)v;
        if (v == 0xFFu) *o++ = 0;
    }
    if (threadIdx.x == 0) a.row_bytes[row_index] = out_bytes;
}

// offsets[f] = first byte of picture f in the packed output, offsets[n] = total (exclusive scan, one block);
// row_off[f][r] = first byte of MCU row r inside its picture
__global__ void jpeg_offsets_kernel(const uint32_t *row_bytes, int n_frames, int mh, uint32_t header_len,
                                    unsigned long long *offsets, uint32_t *row_off, unsigned long long out_cap, int *status) {
    __shared__ unsigned long long chunk_sum[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (n_frames + nt - 1) / nt;
    const int f0 = min(n_frames, t * per), f1 = min(n_frames, f0 + per);
    unsigned long long s = 0;
    for (int f = f0; f < f1; f++) {
        uint32_t at = header_len;
        for (int r = 0; r < mh; r++) {
            row_off[(size_t)f * mh + r] = at;
            at += row_bytes[(size_t)f * mh + r] + 2u;                    // + RSTn (or EOI after the last row)
        }
        s += at;
    }
    chunk_sum[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < nt; i++) {
            const unsigned long long v = chunk_sum[i];
            chunk_sum[i] = run;
            run += v;
        }
        offsets[n_frames] = run;
        if (run > out_cap) atomicMax(status, 2);
    }
    __syncthreads();
    unsigned long long run = chunk_sum[t];
    for (int f = f0; f < f1; f++) {
        offsets[f] = run;
        const size_t last = (size_t)f * mh + mh - 1;
        run += row_off[last] + row_bytes[last] + 2u;
    }
}

// one block per (MCU row or header, picture): header / row bytes and the marker that follows the row
__global__ void __launch_bounds__(128) jpeg_assemble_kernel(const uint8_t *header, uint32_t header_len, const uint8_t *scratch,
                                                            const uint32_t *row_bytes, const uint32_t *row_off, int mh,
                                                            int row_cap, const unsigned long long *offsets,
                                                            unsigned long long out_cap, const int *status, uint8_t *out) {
    const int f = blockIdx.y, r = (int)blockIdx.x - 1;
    if (*status != 0 || offsets[f + 1] > out_cap) return;
    uint8_t *o = out + offsets[f];
    if (r < 0) {
        for (uint32_t i = threadIdx.x; i < header_len; i += blockDim.x) o[i] = header[i];
        return;
    }
    const uint32_t n = row_bytes[(size_t)f * mh + r], at = row_off[(size_t)f * mh + r];
    const uint8_t *s = scratch + ((size_t)f * mh + r) * (size_t)row_cap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) o[at + i] = s[i];
    if (threadIdx.x == 0) {
        o[at + n] = 0xFF;
        o[at + n + 1] = (uint8_t)(r + 1 < mh ? 0xD0 + (r & 7) : 0xD9);     // RSTm after every row but the last, then EOI
    }
}

}  // namespace
}  // namespace vt

struct vt_jpeg_plan {
    int w = 0, h = 0, quality = 0, expand = 0, device = 0;
    int mw = 0, mh = 0, cap_words = 0, row_cap = 0;
    size_t smem = 0;
    uint16_t qdiv[2][64];
    std::vector<uint8_t> header;
    uint8_t *header_dev = nullptr;
    uint8_t *scratch = nullptr;
    uint32_t *row_bytes = nullptr, *row_off = nullptr;
    int scratch_frames = 0;
};

namespace {

void quant_table(int quality, bool chroma, uint8_t *out64) {
    quality = std::min(100, std::max(1, quality));
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    const uint8_t *base = chroma ? vt::kQChr : vt::kQLum;
    for (int i = 0; i < 64; i++) {
        long t = ((long)base[i] * scale + 50L) / 100L;
        out64[i] = (uint8_t)std::min(255L, std::max(1L, t));
    }
}

void put_dht(std::vector<uint8_t> &h, int tc_th, const uint8_t *bits, const uint8_t *vals, int n) {
    h.push_back(0xFF); h.push_back(0xC4);
    const int len = 2 + 1 + 16 + n;
    h.push_back((uint8_t)(len >> 8)); h.push_back((uint8_t)len);
    h.push_back((uint8_t)tc_th);
    h.insert(h.end(), bits, bits + 16);
    h.insert(h.end(), vals, vals + n);
}

}  // namespace

extern "C" int vt_jpeg_plan_create(int w, int h, int quality, int expand_range, vt_jpeg_plan **out) {
    if (!out || w < 2 || h < 2 || w > 16384 || h > 16384 || quality < 1 || quality > 100) {
        vt::set_error("vt_jpeg_plan_create: bad arguments (size 2..16384, quality 1..100)");
        return VT_ERR_INVALID;
    }
    auto *p = new vt_jpeg_plan;
    p->w = w; p->h = h; p->quality = quality; p->expand = expand_range ? 1 : 0;
    p->device = vt::current_device();
    p->mw = (w + 15) / 16;
    p->mh = (h + 15) / 16;
    if (p->mw * 6 > 0xFFFF) {                             // DRI carries the interval in 16 bits
        delete p;
        vt::set_error("vt_jpeg_plan_create: picture too wide for one restart interval per MCU row");
        return VT_ERR_UNSUPPORTED;
    }
    // bit buffer of an MCU row: as many bytes as the row has samples (a 1:1 "compression" is the refusal point)
    p->cap_words = (p->mw * 16 * 16 * 3 / 2 + 3) / 4;
    p->row_cap = p->cap_words * 4 + p->cap_words;         // + 25 % for stuffed zero bytes
    p->smem = (size_t)((p->mw * 6 + 1) & ~1) * 128 + (size_t)p->cap_words * 4 + 4 * 256 * 4;
    if (p->smem > 220 * 1024) {
        delete p;
        vt::set_error("vt_jpeg_plan_create: %d-pixel-wide pictures need %zu bytes of shared memory per MCU row", w, p->smem);
        return VT_ERR_UNSUPPORTED;
    }
    uint8_t q[2][64];
    quant_table(quality, false, q[0]);
    quant_table(quality, true, q[1]);
    for (int t = 0; t < 2; t++)
        for (int k = 0; k < 64; k++) p->qdiv[t][k] = (uint16_t)(q[t][vt::kZigzag[k]] << 3);
    auto &hd = p->header;
    const uint8_t soi_app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    hd.insert(hd.end(), soi_app0, soi_app0 + sizeof(soi_app0));
    for (int t = 0; t < 2; t++) {
        const uint8_t dqt[] = {0xFF, 0xDB, 0, 67, (uint8_t)t};
        hd.insert(hd.end(), dqt, dqt + sizeof(dqt));
        for (int k = 0; k < 64; k++) hd.push_back(q[t][vt::kZigzag[k]]);
    }
    const uint8_t sof[] = {0xFF, 0xC0, 0, 17, 8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3,
                           1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    hd.insert(hd.end(), sof, sof + sizeof(sof));
    put_dht(hd, 0x00, vt::kDcLumBits, vt::kDcVals, 12);
    put_dht(hd, 0x10, vt::kAcLumBits, vt::kAcLumVals, 162);
    put_dht(hd, 0x01, vt::kDcChrBits, vt::kDcVals, 12);
    put_dht(hd, 0x11, vt::kAcChrBits, vt::kAcChrVals, 162);
    const uint8_t dri_sos[] = {0xFF, 0xDD, 0, 4, (uint8_t)(p->mw >> 8), (uint8_t)p->mw,
                               0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    hd.insert(hd.end(), dri_sos, dri_sos + sizeof(dri_sos));
    uint32_t tabs[4][256];
    vt::build_table(vt::kDcLumBits, vt::kDcVals, tabs[0]);
    vt::build_table(vt::kAcLumBits, vt::kAcLumVals, tabs[1]);
    vt::build_table(vt::kDcChrBits, vt::kDcVals, tabs[2]);
    vt::build_table(vt::kAcChrBits, vt::kAcChrVals, tabs[3]);
    cudaError_t e = cudaMemcpyToSymbol(vt::c_huff, tabs, sizeof(tabs));
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->header_dev, hd.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->header_dev, hd.data(), hd.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && p->smem > 48 * 1024)
        e = cudaFuncSetAttribute(vt::jpeg_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    if (e != cudaSuccess) {
        if (p->header_dev) cudaFree(p->header_dev);
        delete p;
        (void)cudaGetLastError();
        return vt::cuda_fail(e, "vt_jpeg_plan_create");
    }
    *out = p;
    return VT_OK;
}

extern "C" void vt_jpeg_plan_destroy(vt_jpeg_plan *p) {
    if (!p) return;
    if (p->header_dev) cudaFree(p->header_dev);
    if (p->scratch) cudaFree(p->scratch);
    if (p->row_bytes) cudaFree(p->row_bytes);
    if (p->row_off) cudaFree(p->row_off);
    delete p;
}

extern "C" size_t vt_jpeg_max_frame_bytes(const vt_jpeg_plan *p) {
    return p ? p->header.size() + (size_t)p->mh * (p->row_cap + 2) : 0;
}

extern "C" int vt_jpeg_header(const vt_jpeg_plan *p, uint8_t *out, size_t cap, size_t *len) {
    if (!p || !len) return VT_ERR_INVALID;
    *len = p->header.size();
    if (out && cap >= p->header.size()) memcpy(out, p->header.data(), p->header.size());
    return VT_OK;
}

extern "C" int vt_jpeg_encode_yuv420p(vt_jpeg_plan *p, const uint8_t *src_dev, size_t src_frame_stride, int n_frames,
                                      uint8_t *out_dev, size_t out_cap, uint64_t *offsets_dev, int32_t *status_dev,
                                      void *stream) {
    if (!p || !src_dev || !out_dev || !offsets_dev || !status_dev || n_frames <= 0 || n_frames > 65535) {
        vt::set_error("vt_jpeg_encode_yuv420p: bad arguments (1..65535 pictures per call)");
        return VT_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_frames > p->scratch_frames) {                  // grows outside the stream's order: rare, synchronise first
        VT_CUDA(cudaStreamSynchronize(st));
        if (p->scratch) cudaFree(p->scratch);
        if (p->row_bytes) cudaFree(p->row_bytes);
        if (p->row_off) cudaFree(p->row_off);
        p->scratch = nullptr;
        p->row_bytes = p->row_off = nullptr;
        p->scratch_frames = 0;
        VT_CUDA(cudaMalloc((void **)&p->scratch, (size_t)n_frames * p->mh * p->row_cap));
        VT_CUDA(cudaMalloc((void **)&p->row_bytes, (size_t)n_frames * p->mh * sizeof(uint32_t)));
        VT_CUDA(cudaMalloc((void **)&p->row_off, (size_t)n_frames * p->mh * sizeof(uint32_t)));
        p->scratch_frames = n_frames;
    }
    VT_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int32_t), st));
    vt::JpegArgs a;
    a.src = src_dev;
    a.frame_stride = src_frame_stride;
    a.w = p->w; a.h = p->h; a.cw = (p->w + 1) / 2; a.ch = (p->h + 1) / 2;
    a.mw = p->mw; a.mh = p->mh;
    a.expand = p->expand;
    a.cap_words = p->cap_words;
    a.row_cap = p->row_cap;
    a.scratch = p->scratch;
    a.row_bytes = p->row_bytes;
    a.status = status_dev;
    memcpy(a.qdiv, p->qdiv, sizeof(a.qdiv));
    vt::jpeg_row_kernel<<<dim3((unsigned)p->mh, (unsigned)n_frames), vt::JT, p->smem, st>>>(a);
    VT_LAUNCHED("jpeg_row_kernel");
    const int nt = n_frames >= 1024 ? 1024 : ((n_frames + 31) / 32) * 32;
    vt::jpeg_offsets_kernel<<<1, nt, 0, st>>>(p->row_bytes, n_frames, p->mh, (uint32_t)p->header.size(),
                                              (unsigned long long *)offsets_dev, p->row_off, (unsigned long long)out_cap,
                                              status_dev);
    VT_LAUNCHED("jpeg_offsets_kernel");
    vt::jpeg_assemble_kernel<<<dim3((unsigned)p->mh + 1, (unsigned)n_frames), 128, 0, st>>>(
                                                                 p->header_dev, (uint32_t)p->header.size(), p->scratch,
                                                                 p->row_bytes, p->row_off, p->mh, p->row_cap,
                                                                 (const unsigned long long *)offsets_dev,
                                                                 (unsigned long long)out_cap, status_dev, out_dev);
    VT_LAUNCHED("jpeg_assemble_kernel");
    return VT_OK;
}
