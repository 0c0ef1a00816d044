/*
 * jpeg_oracle.c -- CPU restatement of the baseline-JPEG encoder behind the upload-size reducer (SURVEY.md section 8 rows
 * a8 / f4).  TEST INFRASTRUCTURE ONLY: nothing in video_transformer_b200/ may link or call this file.
 *
 * What it restates.  The reference reduces an upload with `ffmpeg ... -c:v libx264 -crf 28`
 * (/root/reference/src/analyzer/content_analyzer.py:193-217); a B200 has no video encoder, so the B200-native reducer
 * writes Motion-JPEG: ITU-T T.81 baseline sequential DCT, Huffman coding with the typical tables of Annex K.  The
 * arithmetic follows the Independent JPEG Group's library (third party, not under /root/reference; the copy linked into
 * this image's OpenCV 4.13 is libjpeg-turbo), restated from its published algorithm:
 *   - jfdctint.c (JDCT_ISLOW): Loeffler-Ligtenberg-Moschytz integer forward DCT, CONST_BITS 13, PASS1_BITS 2,
 *     output scaled by 8;
 *   - jcdctmgr.c: sample - 128, divisor = quantval * 8, round half away from zero;
 *   - jcparam.c: jpeg_quality_scaling (q < 50: 5000/q, else 200 - 2q), (base * scale + 50) / 100 clamped to 1..255;
 *   - jchuff.c: DC difference / AC run-length symbols, 0xFF byte stuffing, 1-padding before a restart marker.
 * Pinned by tests/test_jpeg.py: the scan of a grey image encoded here is byte-identical to cv2.imencode's
 * (libjpeg-turbo) at several qualities, and the DQT / DHT segments equal the library's.
 *
 * Input colour: the frames are limited-range BT.601 YCbCr (video); JFIF is full range, so samples are expanded first:
 *   Y' = clamp(((Y - 16) * 19077 + 8192) >> 14),  C' = clamp((((C - 128) * 18652 + 8192) >> 14) + 128)
 * (19077/16384 = 255/219, 18652/16384 = 255/224; >> is an arithmetic shift).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const uint8_t vtj_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                       41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                       30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const uint8_t vtj_qlum[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t vtj_qchr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                     99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t vtj_dc_lum_bits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t vtj_dc_chr_bits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t vtj_dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t vtj_ac_lum_bits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t vtj_ac_lum_vals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
    0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
    0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94,
    0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8,
    0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};
static const uint8_t vtj_ac_chr_bits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t vtj_ac_chr_vals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
    0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
    0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
    0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92,
    0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};

typedef struct {
    uint16_t code[256];
    uint8_t len[256];
} vtj_huff;

static void vtj_build(const uint8_t *bits, const uint8_t *vals, vtj_huff *h) {
    memset(h, 0, sizeof(*h));
    unsigned code = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < bits[l - 1]; i++) {
            h->code[vals[k]] = (uint16_t)code++;
            h->len[vals[k]] = (uint8_t)l;
            k++;
        }
        code <<= 1;
    }
}

void vtj_quant_table(int quality, int chroma, uint8_t *out64 /* natural order */) {
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    const uint8_t *base = chroma ? vtj_qchr : vtj_qlum;
    for (int i = 0; i < 64; i++) {
        long t = ((long)base[i] * scale + 50L) / 100L;
        if (t < 1) t = 1;
        if (t > 255) t = 255;
        out64[i] = (uint8_t)t;
    }
}

#define VTJ_DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

/* jfdctint.c: in-place on 64 ints (already level shifted), output scaled by 8 */
static void vtj_fdct(int *data) {
    enum { CB = 13, P1 = 2 };
    int *d = data;
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 8; i++) {
            int *p = pass == 0 ? data + 8 * i : data + i;
            const int s = pass == 0 ? 1 : 8;
            int t0 = p[0] + p[7 * s], t7 = p[0] - p[7 * s], t1 = p[s] + p[6 * s], t6 = p[s] - p[6 * s];
            int t2 = p[2 * s] + p[5 * s], t5 = p[2 * s] - p[5 * s], t3 = p[3 * s] + p[4 * s], t4 = p[3 * s] - p[4 * s];
            int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            if (pass == 0) {
                p[0] = (t10 + t11) << P1;
                p[4 * s] = (t10 - t11) << P1;
            } else {
                p[0] = VTJ_DESCALE(t10 + t11, P1);
                p[4 * s] = VTJ_DESCALE(t10 - t11, P1);
            }
            const int sh = pass == 0 ? CB - P1 : CB + P1;
            int z1 = (t12 + t13) * 4433;
            p[2 * s] = VTJ_DESCALE(z1 + t13 * 6270, sh);
            p[6 * s] = VTJ_DESCALE(z1 + t12 * (-15137), sh);
            z1 = t4 + t7;
            int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * 9633;
            t4 *= 2446;
            t5 *= 16819;
            t6 *= 25172;
            t7 *= 12299;
            z1 *= -7373;
            z2 *= -20995;
            z3 *= -16069;
            z4 *= -3196;
            z3 += z5;
            z4 += z5;
            p[7 * s] = VTJ_DESCALE(t4 + z1 + z3, sh);
            p[5 * s] = VTJ_DESCALE(t5 + z2 + z4, sh);
            p[3 * s] = VTJ_DESCALE(t6 + z2 + z3, sh);
            p[s] = VTJ_DESCALE(t7 + z1 + z4, sh);
        }
    }
    (void)d;
}

/* One 8x8 block: samples (with edge replication past w,h) -> 64 quantised coefficients in ZIGZAG order */
void vtj_block_coefficients(const uint8_t *plane, int pitch, int w, int h, int bx, int by, const uint8_t *q64,
                            int expand /* 0 none, 1 luma 16..235, 2 chroma 16..240 */, int16_t *out_zz) {
    int d[64];
    for (int r = 0; r < 8; r++) {
        int y = by * 8 + r;
        if (y > h - 1) y = h - 1;
        for (int c = 0; c < 8; c++) {
            int x = bx * 8 + c;
            if (x > w - 1) x = w - 1;
            int v = plane[(size_t)y * pitch + x];
            if (expand == 1) {
                v = ((v - 16) * 19077 + 8192) >> 14;
            } else if (expand == 2) {
                v = (((v - 128) * 18652 + 8192) >> 14) + 128;
            }
            if (v < 0) v = 0;
            if (v > 255) v = 255;
            d[r * 8 + c] = v - 128;
        }
    }
    vtj_fdct(d);
    for (int k = 0; k < 64; k++) {
        const int n = vtj_zigzag[k];
        const int qv = (int)q64[n] << 3;
        int t = d[n];
        if (t < 0) {
            t = -t;
            t += qv >> 1;
            t = t >= qv ? t / qv : 0;
            t = -t;
        } else {
            t += qv >> 1;
            t = t >= qv ? t / qv : 0;
        }
        out_zz[k] = (int16_t)t;
    }
}

typedef struct {
    uint8_t *p;
    size_t n, cap;
    uint32_t acc;
    int nbits;
    int overflow;
} vtj_bits;

static void vtj_byte(vtj_bits *b, int v) {
    if (b->n < b->cap) b->p[b->n] = (uint8_t)v;
    else b->overflow = 1;
    b->n++;
}
static void vtj_put(vtj_bits *b, unsigned code, int len) {
    b->acc = (b->acc << len) | (code & ((1u << len) - 1u));
    b->nbits += len;
    while (b->nbits >= 8) {
        int v = (int)((b->acc >> (b->nbits - 8)) & 0xFF);
        vtj_byte(b, v);
        if (v == 0xFF) vtj_byte(b, 0);
        b->nbits -= 8;
    }
}
static void vtj_flush(vtj_bits *b) {
    if (b->nbits > 0) vtj_put(b, 0x7F, 8 - b->nbits); /* pad with ones */
    b->acc = 0;
    b->nbits = 0;
}
static int vtj_nbits(int v) {
    int n = 0;
    if (v < 0) v = -v;
    while (v) {
        n++;
        v >>= 1;
    }
    return n;
}
static void vtj_encode_block(vtj_bits *b, const int16_t *zz, int *pred, const vtj_huff *dc, const vtj_huff *ac) {
    int diff = zz[0] - *pred;
    *pred = zz[0];
    int n = vtj_nbits(diff);
    vtj_put(b, dc->code[n], dc->len[n]);
    if (n) vtj_put(b, (unsigned)(diff < 0 ? diff - 1 : diff), n);
    int run = 0;
    for (int k = 1; k < 64; k++) {
        int v = zz[k];
        if (v == 0) {
            run++;
            continue;
        }
        while (run > 15) {
            vtj_put(b, ac->code[0xF0], ac->len[0xF0]);
            run -= 16;
        }
        n = vtj_nbits(v);
        int sym = (run << 4) | n;
        vtj_put(b, ac->code[sym], ac->len[sym]);
        vtj_put(b, (unsigned)(v < 0 ? v - 1 : v), n);
        run = 0;
    }
    if (run > 0) vtj_put(b, ac->code[0], ac->len[0]);
}

static void vtj_marker(vtj_bits *b, int m) {
    vtj_byte(b, 0xFF);
    vtj_byte(b, m);
}
static void vtj_u16(vtj_bits *b, int v) {
    vtj_byte(b, v >> 8);
    vtj_byte(b, v & 0xFF);
}
static void vtj_dht(vtj_bits *b, int tc_th, const uint8_t *bits, const uint8_t *vals, int nvals) {
    vtj_marker(b, 0xC4);
    vtj_u16(b, 2 + 1 + 16 + nvals);
    vtj_byte(b, tc_th);
    for (int i = 0; i < 16; i++) vtj_byte(b, bits[i]);
    for (int i = 0; i < nvals; i++) vtj_byte(b, vals[i]);
}

/* Header of a frame: SOI, APP0 (JFIF 1.01, no density), DQT x ncomp tables, SOF0, DHT x 2 or 4, DRI, SOS.
 * Returns its length.  n_comp = 1 (grey) or 3 (YCbCr 4:2:0). */
size_t vtj_header(int w, int h, int n_comp, int quality, int restart_interval, uint8_t *out, size_t cap) {
    vtj_bits b = {out, 0, cap, 0, 0, 0};
    uint8_t q[64];
    vtj_marker(&b, 0xD8);
    vtj_marker(&b, 0xE0);
    vtj_u16(&b, 16);
    const char *jfif = "JFIF";
    for (int i = 0; i < 5; i++) vtj_byte(&b, jfif[i]);
    vtj_byte(&b, 1);
    vtj_byte(&b, 1);
    vtj_byte(&b, 0);
    vtj_u16(&b, 1);
    vtj_u16(&b, 1);
    vtj_byte(&b, 0);
    vtj_byte(&b, 0);
    for (int t = 0; t < (n_comp == 3 ? 2 : 1); t++) {
        vtj_quant_table(quality, t, q);
        vtj_marker(&b, 0xDB);
        vtj_u16(&b, 67);
        vtj_byte(&b, t);
        for (int k = 0; k < 64; k++) vtj_byte(&b, q[vtj_zigzag[k]]);
    }
    vtj_marker(&b, 0xC0);
    vtj_u16(&b, 8 + 3 * n_comp);
    vtj_byte(&b, 8);
    vtj_u16(&b, h);
    vtj_u16(&b, w);
    vtj_byte(&b, n_comp);
    for (int c = 0; c < n_comp; c++) {
        vtj_byte(&b, c + 1);
        vtj_byte(&b, (c == 0 && n_comp == 3) ? 0x22 : 0x11);
        vtj_byte(&b, c == 0 ? 0 : 1);
    }
    vtj_dht(&b, 0x00, vtj_dc_lum_bits, vtj_dc_vals, 12);
    vtj_dht(&b, 0x10, vtj_ac_lum_bits, vtj_ac_lum_vals, 162);
    if (n_comp == 3) {
        vtj_dht(&b, 0x01, vtj_dc_chr_bits, vtj_dc_vals, 12);
        vtj_dht(&b, 0x11, vtj_ac_chr_bits, vtj_ac_chr_vals, 162);
    }
    if (restart_interval > 0) {
        vtj_marker(&b, 0xDD);
        vtj_u16(&b, 4);
        vtj_u16(&b, restart_interval);
    }
    vtj_marker(&b, 0xDA);
    vtj_u16(&b, 6 + 2 * n_comp);
    vtj_byte(&b, n_comp);
    for (int c = 0; c < n_comp; c++) {
        vtj_byte(&b, c + 1);
        vtj_byte(&b, c == 0 ? 0x00 : 0x11);
    }
    vtj_byte(&b, 0);
    vtj_byte(&b, 63);
    vtj_byte(&b, 0);
    return b.overflow ? 0 : b.n;
}

/* Whole picture.  y/u/v planes (u, v NULL for grey); restart_interval in MCUs (0 = none; the GPU encoder uses one MCU row).
 * expand_range: 1 = limited-range video samples are expanded to full range first.  Returns bytes written, 0 on overflow. */
size_t vtj_encode(const uint8_t *y, int y_pitch, const uint8_t *u, const uint8_t *v, int c_pitch, int w, int h, int quality,
                  int restart_interval, int expand_range, uint8_t *out, size_t cap) {
    const int n_comp = (u && v) ? 3 : 1;
    size_t hl = vtj_header(w, h, n_comp, quality, restart_interval, out, cap);
    if (!hl) return 0;
    vtj_bits b = {out, hl, cap, 0, 0, 0};
    vtj_huff dcl, acl, dcc, acc;
    vtj_build(vtj_dc_lum_bits, vtj_dc_vals, &dcl);
    vtj_build(vtj_ac_lum_bits, vtj_ac_lum_vals, &acl);
    vtj_build(vtj_dc_chr_bits, vtj_dc_vals, &dcc);
    vtj_build(vtj_ac_chr_bits, vtj_ac_chr_vals, &acc);
    uint8_t ql[64], qc[64];
    vtj_quant_table(quality, 0, ql);
    vtj_quant_table(quality, 1, qc);
    const int ms = n_comp == 3 ? 16 : 8;
    const int mw = (w + ms - 1) / ms, mh = (h + ms - 1) / ms;
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    int pred[3] = {0, 0, 0};
    int count = 0, rst = 0;
    int16_t zz[64];
    for (int my = 0; my < mh; my++)
        for (int mx = 0; mx < mw; mx++) {
            if (restart_interval > 0 && count == restart_interval) {
                vtj_flush(&b);
                vtj_marker(&b, 0xD0 + (rst & 7));
                rst++;
                count = 0;
                pred[0] = pred[1] = pred[2] = 0;
            }
            if (n_comp == 3) {
                for (int k = 0; k < 4; k++) {
                    vtj_block_coefficients(y, y_pitch, w, h, mx * 2 + (k & 1), my * 2 + (k >> 1), ql, expand_range ? 1 : 0, zz);
                    vtj_encode_block(&b, zz, &pred[0], &dcl, &acl);
                }
                vtj_block_coefficients(u, c_pitch, cw, ch, mx, my, qc, expand_range ? 2 : 0, zz);
                vtj_encode_block(&b, zz, &pred[1], &dcc, &acc);
                vtj_block_coefficients(v, c_pitch, cw, ch, mx, my, qc, expand_range ? 2 : 0, zz);
                vtj_encode_block(&b, zz, &pred[2], &dcc, &acc);
            } else {
                vtj_block_coefficients(y, y_pitch, w, h, mx, my, ql, expand_range ? 1 : 0, zz);
                vtj_encode_block(&b, zz, &pred[0], &dcl, &acl);
            }
            count++;
        }
    vtj_flush(&b);
    vtj_marker(&b, 0xD9);
    return b.overflow ? 0 : b.n;
}
