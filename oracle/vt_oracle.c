/*
 * vt_oracle.c -- CPU restatement of the long-video ingest arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under video_transformer_b200/ may link, import or execute this file; it is the checker for
 * the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
 *
 * What it restates, and where the authority for each function lives:
 *
 *  - vto_sws_*            the reference's only pixel recipe is `ffmpeg -vf scale=-2:360`
 *                         (/root/reference/src/analyzer/content_analyzer.py:193-211), i.e. FFmpeg's
 *                         libswscale with the scale filter's default SWS_BICUBIC.  libswscale is a
 *                         third-party dependency that is NOT vendored in /root/reference and is not
 *                         pinned by requirements.txt (/root/reference/requirements.txt:1-9).  This
 *                         file restates its published algorithm (libswscale/utils.c initFilter,
 *                         swscale.c hScale8To15_c, output.c yuv2planeX_8_c) for the version the image
 *                         carries: libswscale 9.1.100 (FFmpeg 8.0.1, OpenCV wheel).  Parity is PINNED:
 *                         tests/test_oracle.py checks it bit-for-bit against that library run
 *                         with SWS_ACCURATE_RND|SWS_BITEXACT (the CPU-independent C path), and
 *                         tests/golden/sws_*.npz hold outputs generated from the library.
 *  - vto_sad_hist, vto_nv12_* , vto_rgb24_*   no reference code exists; the
 *                         definitions in SURVEY.md section 8a K1/K3 are the authority and this file is
 *                         their first statement ("parity unpinned by the reference", pinned against
 *                         numpy and, for RGB, against libswscale in tests/test_oracle.py).
 *
 * Plain C99, no dependencies.  Build: make -C oracle  (-> oracle/libvtoracle.so)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VTO_SWS_BILINEAR 2
#define VTO_SWS_BICUBIC 4
#define VTO_SWS_AREA 0x20

static int vto_log2(unsigned v) {
    int n = 0;
    while (v >>= 1) n++;
    return n;
}

static int64_t vto_abs64(int64_t v) { return v < 0 ? -v : v; }

/* FFmpeg's ROUNDED_DIV: round half away from zero. */
static int64_t vto_rounded_div(int64_t a, int64_t b) {
    return (a >= 0 ? a + (b >> 1) : a - (b >> 1)) / b;
}

/*
 * Polyphase filter bank for one axis, as libswscale's initFilter builds it with
 * srcPos == dstPos == 128 (progressive yuv420p through the scale filter), no user src/dst filters,
 * SWS_BITEXACT set (coefficients beyond the reduced tap count are dropped, so filterAlign padding
 * carries zeros and the result does not depend on the host CPU).
 *
 *   one   : 1<<14 for the horizontal pass, 1<<12 for the vertical pass
 *   coef  : out, dst_w * (*taps) int16
 *   pos   : out, dst_w int32 (first source index of each output's window)
 * Returns 0, or -1 on bad arguments / allocation failure.  *taps receives the reduced tap count
 * ("minFilterSize"); callers allocate coef with vto_sws_max_taps() columns.
 */
int vto_sws_max_taps(int src_w, int dst_w, int flags) {
    int size_factor = (flags & VTO_SWS_BICUBIC) ? 4 : 2; /* AREA: 1 -> see below */
    int64_t x_inc = (((int64_t)src_w << 16) + (dst_w >> 1)) / dst_w;
    int fs;
    if (flags & VTO_SWS_AREA) size_factor = 1;
    if ((flags & VTO_SWS_BICUBIC)) size_factor = 4;
    else if (flags & VTO_SWS_BILINEAR) size_factor = 2;
    if (x_inc <= (1 << 16)) fs = 1 + size_factor;
    else fs = 1 + (size_factor * src_w + dst_w - 1) / dst_w;
    if (fs > src_w - 2) fs = src_w - 2;
    if (fs < 1) fs = 1;
    return fs;
}

int vto_sws_make_filter(int src_w, int dst_w, int flags, int one, int16_t *coef, int32_t *pos, int *taps) {
    if (src_w < 1 || dst_w < 1 || !coef || !pos || !taps) return -1;
    const int64_t x_inc = (((int64_t)src_w << 16) + (dst_w >> 1)) / dst_w;
    int shift_lim = vto_log2((unsigned)(src_w / dst_w > 0 ? src_w / dst_w : 1));
    if (src_w / dst_w <= 0) shift_lim = 0;
    if (shift_lim > 8) shift_lim = 8;
    const int64_t fone = 1LL << (54 - shift_lim);
    int fsize;
    int64_t *filt = NULL;
    int i, j;

    if (vto_abs64(x_inc - 0x10000) < 10) {
        /* unscaled axis: identity */
        fsize = 1;
        filt = (int64_t *)calloc((size_t)dst_w, sizeof(int64_t));
        if (!filt) return -1;
        for (i = 0; i < dst_w; i++) { filt[i] = fone; pos[i] = i; }
    } else if (x_inc <= (1 << 16) && (flags & VTO_SWS_AREA)) {
        /* area upscale == bilinear upscale in libswscale */
        int64_t x_dst_in_src;
        fsize = 2;
        filt = (int64_t *)calloc((size_t)dst_w * fsize, sizeof(int64_t));
        if (!filt) return -1;
        x_dst_in_src = ((128 * x_inc) >> 8) - ((128 * 0x8000LL) >> 7);
        for (i = 0; i < dst_w; i++) {
            int xx = (int)((x_dst_in_src - ((int64_t)(fsize - 1) << 15) + (1 << 15)) >> 16);
            pos[i] = xx;
            for (j = 0; j < fsize; j++) {
                int64_t c = fone - vto_abs64(((int64_t)xx * (1 << 16)) - x_dst_in_src) * (fone >> 16);
                if (c < 0) c = 0;
                filt[i * fsize + j] = c;
                xx++;
            }
            x_dst_in_src += x_inc;
        }
    } else {
        int size_factor;
        int64_t x_dst_in_src;
        if (flags & VTO_SWS_BICUBIC) size_factor = 4;
        else if (flags & VTO_SWS_AREA) size_factor = 1;
        else if (flags & VTO_SWS_BILINEAR) size_factor = 2;
        else return -1;
        if (x_inc <= (1 << 16)) fsize = 1 + size_factor;
        else fsize = 1 + (size_factor * src_w + dst_w - 1) / dst_w;
        if (fsize > src_w - 2) fsize = src_w - 2;
        if (fsize < 1) fsize = 1;
        filt = (int64_t *)calloc((size_t)dst_w * fsize, sizeof(int64_t));
        if (!filt) return -1;
        x_dst_in_src = ((128 * x_inc) >> 7) - ((128 * 0x10000LL) >> 7);
        for (i = 0; i < dst_w; i++) {
            int xx = (int)((x_dst_in_src - (int64_t)(fsize - 2) * (1LL << 16)) / (1 << 17));
            pos[i] = xx;
            for (j = 0; j < fsize; j++) {
                int64_t d = vto_abs64(((int64_t)xx * (1 << 17)) - x_dst_in_src) << 13;
                int64_t c;
                if (x_inc > (1 << 16)) d = d * dst_w / src_w;
                if (flags & VTO_SWS_BICUBIC) {
                    const int64_t B = 0;
                    const int64_t C = (int64_t)(0.6 * (1 << 24));
                    if (d >= (1LL << 31)) {
                        c = 0;
                    } else {
                        int64_t dd = (d * d) >> 30;
                        int64_t ddd = (dd * d) >> 30;
                        if (d < (1LL << 30))
                            c = (12 * (1 << 24) - 9 * B - 6 * C) * ddd + (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                                (6 * (1 << 24) - 2 * B) * (1LL << 30);
                        else
                            c = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd + (-12 * B - 48 * C) * d +
                                (8 * B + 24 * C) * (1LL << 30);
                    }
                    c /= (1LL << 54) / fone;
                } else if (flags & VTO_SWS_AREA) {
                    int64_t d2 = d - (1 << 29);
                    if (d2 * x_inc < -(1LL << (29 + 16))) c = 1LL << (30 + 16);
                    else if (d2 * x_inc < (1LL << (29 + 16))) c = -d2 * x_inc + (1LL << (29 + 16));
                    else c = 0;
                    c *= fone >> (30 + 16);
                } else { /* bilinear */
                    c = (1 << 30) - d;
                    if (c < 0) c = 0;
                    c *= fone >> 30;
                }
                filt[i * fsize + j] = c;
                xx++;
            }
            x_dst_in_src += 2 * x_inc;
        }
    }

    /* Trim near-zero taps: shift windows right past negligible leading taps, count trailing ones. */
    {
        const double cutoff_lim = 0.002 * (double)fone; /* SWS_MAX_REDUCE_CUTOFF */
        int min_size = 0;
        for (i = dst_w - 1; i >= 0; i--) {
            int min = fsize;
            int64_t cut = 0;
            for (j = 0; j < fsize; j++) {
                int k;
                cut += vto_abs64(filt[i * fsize]);
                if ((double)cut > cutoff_lim) break;
                if (i < dst_w - 1 && pos[i] >= pos[i + 1]) break;
                for (k = 1; k < fsize; k++) filt[i * fsize + k - 1] = filt[i * fsize + k];
                filt[i * fsize + k - 1] = 0;
                pos[i]++;
            }
            cut = 0;
            for (j = fsize - 1; j > 0; j--) {
                cut += vto_abs64(filt[i * fsize + j]);
                if ((double)cut > cutoff_lim) break;
                min--;
            }
            if (min > min_size) min_size = min;
        }
        if (min_size < 1) min_size = 1;
        /* compact to min_size columns (BITEXACT: drop everything beyond) */
        if (min_size < fsize) {
            for (i = 0; i < dst_w; i++)
                for (j = 0; j < min_size; j++) filt[i * min_size + j] = filt[i * fsize + j];
        }
        fsize = min_size;
    }

    /* Fold windows that hang over either edge back inside [0, src_w). */
    for (i = 0; i < dst_w; i++) {
        if (pos[i] < 0) {
            for (j = 1; j < fsize; j++) {
                int left = j + pos[i] > 0 ? j + pos[i] : 0;
                filt[i * fsize + left] += filt[i * fsize + j];
                filt[i * fsize + j] = 0;
            }
            pos[i] = 0;
        }
        if (pos[i] + fsize > src_w) {
            int shift = pos[i] + (fsize - src_w < 0 ? fsize - src_w : 0);
            int64_t acc = 0;
            for (j = fsize - 1; j >= 0; j--) {
                if (pos[i] + j >= src_w) {
                    acc += filt[i * fsize + j];
                    filt[i * fsize + j] = 0;
                }
            }
            for (j = fsize - 1; j >= 0; j--) {
                if (j < shift) filt[i * fsize + j] = 0;
                else filt[i * fsize + j] = filt[i * fsize + j - shift];
            }
            pos[i] -= shift;
            filt[i * fsize + src_w - 1 - pos[i]] += acc;
        }
    }

    /* Normalise each row to `one` with running error feedback. */
    for (i = 0; i < dst_w; i++) {
        int64_t err = 0, sum = 0;
        for (j = 0; j < fsize; j++) sum += filt[i * fsize + j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (j = 0; j < fsize; j++) {
            int64_t v = filt[i * fsize + j] + err;
            int iv = (int)vto_rounded_div(v, sum);
            coef[i * fsize + j] = (int16_t)iv;
            err = v - iv * sum;
        }
    }
    *taps = fsize;
    free(filt);
    return 0;
}

/* Horizontal pass: 8-bit row -> 15-bit intermediates (hScale8To15_c). */
static void vto_hscale_row(const uint8_t *src, int16_t *dst, int dst_w, const int16_t *coef, const int32_t *pos,
                           int taps) {
    for (int i = 0; i < dst_w; i++) {
        int v = 0;
        const uint8_t *s = src + pos[i];
        const int16_t *c = coef + (size_t)i * taps;
        for (int j = 0; j < taps; j++) v += (int)s[j] * c[j];
        v >>= 7;
        dst[i] = (int16_t)(v > 32767 ? 32767 : v);
    }
}

static uint8_t vto_clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/*
 * One plane, separable: every source row through the horizontal bank, then each output row is a
 * vertical FIR over 15-bit rows (yuv2planeX_8_c with the flat dither 64: val = 64<<12; >>19), or the
 * 1-tap form (yuv2plane1_8_c: (v + 64) >> 7) when the vertical axis is unscaled.
 */
int vto_sws_scale_plane(const uint8_t *src, int sw, int sh, int spitch, uint8_t *dst, int dw, int dh, int dpitch,
                        int flags) {
    int ht = vto_sws_max_taps(sw, dw, flags), vt = vto_sws_max_taps(sh, dh, flags);
    if (ht < 4) ht = 4;
    if (vt < 4) vt = 4;
    int16_t *hc = (int16_t *)malloc(sizeof(int16_t) * (size_t)dw * ht);
    int16_t *vc = (int16_t *)malloc(sizeof(int16_t) * (size_t)dh * vt);
    int32_t *hp = (int32_t *)malloc(sizeof(int32_t) * (size_t)dw);
    int32_t *vp = (int32_t *)malloc(sizeof(int32_t) * (size_t)dh);
    int16_t *mid = (int16_t *)malloc(sizeof(int16_t) * (size_t)dw * sh);
    int rc = -1, htaps = 0, vtaps = 0;
    if (!hc || !vc || !hp || !vp || !mid) goto done;
    if (vto_sws_make_filter(sw, dw, flags, 1 << 14, hc, hp, &htaps)) goto done;
    if (vto_sws_make_filter(sh, dh, flags, 1 << 12, vc, vp, &vtaps)) goto done;
    for (int y = 0; y < sh; y++) vto_hscale_row(src + (size_t)y * spitch, mid + (size_t)y * dw, dw, hc, hp, htaps);
    for (int y = 0; y < dh; y++) {
        uint8_t *d = dst + (size_t)y * dpitch;
        if (vtaps == 1) {
            const int16_t *r = mid + (size_t)vp[y] * dw;
            for (int x = 0; x < dw; x++) d[x] = vto_clip_u8((r[x] + 64) >> 7);
        } else {
            for (int x = 0; x < dw; x++) {
                int v = 64 << 12;
                for (int j = 0; j < vtaps; j++) v += mid[(size_t)(vp[y] + j) * dw + x] * vc[(size_t)y * vtaps + j];
                d[x] = vto_clip_u8(v >> 19);
            }
        }
    }
    rc = 0;
done:
    free(hc); free(vc); free(hp); free(vp); free(mid);
    return rc;
}

/* K3: 256-bin luma histogram of `cur` and SAD(cur, prev) over the display w x h. */
void vto_sad_hist(const uint8_t *cur, int cur_pitch, const uint8_t *prev, int prev_pitch, int w, int h,
                  uint64_t *sad, uint32_t *hist) {
    uint64_t s = 0;
    memset(hist, 0, 256 * sizeof(uint32_t));
    for (int y = 0; y < h; y++) {
        const uint8_t *c = cur + (size_t)y * cur_pitch;
        const uint8_t *p = prev ? prev + (size_t)y * prev_pitch : NULL;
        for (int x = 0; x < w; x++) {
            hist[c[x]]++;
            if (p) s += (uint64_t)(c[x] > p[x] ? c[x] - p[x] : p[x] - c[x]);
        }
    }
    *sad = s;
}

/* The same K3, written the way a tuned host implementation would be (bench.py's CPU arm takes whichever of this
 * and OpenCV's norm/calcHist is faster, i.e. the STRONGER baseline): SAD in its own loop so the compiler turns it into
 * psadbw, four interleaved sub-histograms so consecutive equal pixels do not serialise on one counter.  Results are
 * identical to vto_sad_hist (tests/test_oracle.py). */
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
void vto_sad_hist_fast(const uint8_t *cur, int cur_pitch, const uint8_t *prev, int prev_pitch, int w, int h,
                       uint64_t *sad, uint32_t *hist) {
    uint64_t s = 0;
    uint32_t sub[4][256];
    memset(sub, 0, sizeof(sub));
    for (int y = 0; y < h; y++) {
        const uint8_t *c = cur + (size_t)y * cur_pitch;
        int x = 0;
        for (; x + 4 <= w; x += 4) {
            sub[0][c[x]]++;
            sub[1][c[x + 1]]++;
            sub[2][c[x + 2]]++;
            sub[3][c[x + 3]]++;
        }
        for (; x < w; x++) sub[0][c[x]]++;
        if (prev) {
            const uint8_t *p = prev + (size_t)y * prev_pitch;
            uint32_t row = 0;
            for (int i = 0; i < w; i++) row += (uint32_t)(c[i] > p[i] ? c[i] - p[i] : p[i] - c[i]);
            s += row;
        }
    }
    for (int i = 0; i < 256; i++) hist[i] = sub[0][i] + sub[1][i] + sub[2][i] + sub[3][i];
    *sad = s;
}

/* K1a: NV12 (pitch-linear Y plane, interleaved UV plane) -> planar YUV420P. Exact copy semantics. */
void vto_nv12_to_yuv420p(const uint8_t *y, const uint8_t *uv, int pitch, int w, int h, uint8_t *dy, uint8_t *du,
                         uint8_t *dv) {
    int cw = (w + 1) / 2, ch = (h + 1) / 2;
    for (int r = 0; r < h; r++) memcpy(dy + (size_t)r * w, y + (size_t)r * pitch, (size_t)w);
    for (int r = 0; r < ch; r++) {
        const uint8_t *s = uv + (size_t)r * pitch;
        for (int x = 0; x < cw; x++) {
            du[(size_t)r * cw + x] = s[2 * x];
            dv[(size_t)r * cw + x] = s[2 * x + 1];
        }
    }
}

/*
 * K0 restatement for the CPU baseline: one I_PCM picture of the synthetic streams (macroblock-ordered raw
 * samples, 384 per macroblock, two header bytes between macroblocks; see video_transformer_b200/synth.py
 * and ITU-T H.264 7.3.5 "pcm_sample_luma / pcm_sample_chroma") -> planar YUV420P at the display size.
 * `payload` points at macroblock 0's first luma sample.  Decode parity itself is pinned against libavcodec
 * in tests/test_decode.py; this function exists so the host-core baseline pays for the same re-layout.
 */
void vto_pcm_picture_to_yuv420p(const uint8_t *payload, int mb_w, int mb_h, int w, int h, uint8_t *y, uint8_t *u,
                                uint8_t *v) {
    const int cw = (w + 1) / 2, ch = (h + 1) / 2;
    for (int my = 0; my < mb_h; my++) {
        for (int mx = 0; mx < mb_w; mx++) {
            const uint8_t *mb = payload + ((size_t)my * mb_w + mx) * 386;
            for (int r = 0; r < 16; r++) {
                int yy = my * 16 + r;
                if (yy >= h) break;
                int n = w - mx * 16;
                if (n > 16) n = 16;
                if (n > 0) memcpy(y + (size_t)yy * w + mx * 16, mb + r * 16, (size_t)n);
            }
            for (int r = 0; r < 8; r++) {
                int yy = my * 8 + r;
                if (yy >= ch) break;
                int n = cw - mx * 8;
                if (n > 8) n = 8;
                if (n > 0) {
                    memcpy(u + (size_t)yy * cw + mx * 8, mb + 256 + r * 8, (size_t)n);
                    memcpy(v + (size_t)yy * cw + mx * 8, mb + 320 + r * 8, (size_t)n);
                }
            }
        }
    }
}

/*
 * K1b / config 5 restatement: NV12 (or planar 4:2:0) -> packed RGB24 with optional scaling, as libswscale's
 * general path does it for `nv12 -> rgb24` with SWS_BICUBIC (swscale.c + output.c yuv2rgb_X_c_template +
 * yuv2rgb.c ff_yuv2rgb_c_init_tables, ITU-R BT.601 limited range = swscale's default colourspace):
 *   - luma:   horizontal bank sw->dw, vertical bank sh->dh
 *   - chroma: horizontal bank ceil(sw/2)->ceil(dw/2) (pairs of output pixels share chroma),
 *             vertical bank ceil(sh/2)->dh (chroma is interpolated vertically to full height)
 *   - Y,U,V = (2^18 + sum mid15 * coef12) >> 19, then the table form of the matrix:
 *       T(i)  = clip_u8(((i * cy) - (400 << 16) + 0x8000) >> 16),  cy = 65536*255/219
 *       R = T(Y + 326 + off(V, crv)),  G = T(Y + 326 + off(U, cgu) + off(V, cgv)),  B = T(Y + 326 + off(U, cbu))
 *       off(c, k) = ((clip_u8(c) * k) >> 16) - (k >> 9),   k = (K*65536 + 0x8000) / cy  for the BT.601 constants.
 * The constants 326/400 are the table offsets of yuv2rgb.c for limited range; tests/test_oracle.py pins the whole
 * function bit-for-bit against libswscale 9.1.100 (default flags and SWS_ACCURATE_RND|SWS_BITEXACT agree there).
 */
static int64_t vto_cdiv(int64_t a, int64_t b) { return a / b; } /* C division truncates toward zero */

static int vto_rgb_T(int64_t idx) {
    const int64_t cy = (65536LL * 255) / 219;
    int64_t v = (idx * cy - (400LL << 16) + 0x8000) >> 16;
    return (int)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

int vto_yuv_to_rgb24(const uint8_t *y, int y_pitch, const uint8_t *u, const uint8_t *v, int c_pitch, int c_step,
                     int sw, int sh, uint8_t *dst, int dw, int dh) {
    const int flags = VTO_SWS_BICUBIC;
    const int csw = (sw + 1) / 2, csh = (sh + 1) / 2, cdw = (dw + 1) / 2;
    const int64_t cy = (65536LL * 255) / 219;
    const int64_t crv = vto_cdiv(104597LL * 65536 + 0x8000, cy), cbu = vto_cdiv(132201LL * 65536 + 0x8000, cy);
    const int64_t cgu = vto_cdiv(-25675LL * 65536 + 0x8000, cy), cgv = vto_cdiv(-53279LL * 65536 + 0x8000, cy);
    int cap_lh = vto_sws_max_taps(sw, dw, flags), cap_lv = vto_sws_max_taps(sh, dh, flags);
    int cap_ch = vto_sws_max_taps(csw, cdw, flags), cap_cv = vto_sws_max_taps(csh, dh, flags);
    if (cap_lh < 4) cap_lh = 4;
    if (cap_lv < 4) cap_lv = 4;
    if (cap_ch < 4) cap_ch = 4;
    if (cap_cv < 4) cap_cv = 4;
    int16_t *lhc = malloc(sizeof(int16_t) * (size_t)dw * cap_lh), *lvc = malloc(sizeof(int16_t) * (size_t)dh * cap_lv);
    int16_t *chc = malloc(sizeof(int16_t) * (size_t)cdw * cap_ch), *cvc = malloc(sizeof(int16_t) * (size_t)dh * cap_cv);
    int32_t *lhp = malloc(4 * (size_t)dw), *lvp = malloc(4 * (size_t)dh), *chp = malloc(4 * (size_t)cdw),
            *cvp = malloc(4 * (size_t)dh);
    int16_t *my = malloc(sizeof(int16_t) * (size_t)dw * sh), *mu = malloc(sizeof(int16_t) * (size_t)cdw * csh),
            *mv = malloc(sizeof(int16_t) * (size_t)cdw * csh);
    int lht, lvt, cht, cvt, rc = -1;
    if (!lhc || !lvc || !chc || !cvc || !lhp || !lvp || !chp || !cvp || !my || !mu || !mv) goto done;
    if (vto_sws_make_filter(sw, dw, flags, 1 << 14, lhc, lhp, &lht) || vto_sws_make_filter(sh, dh, flags, 1 << 12, lvc, lvp, &lvt) ||
        vto_sws_make_filter(csw, cdw, flags, 1 << 14, chc, chp, &cht) || vto_sws_make_filter(csh, dh, flags, 1 << 12, cvc, cvp, &cvt))
        goto done;
    for (int r = 0; r < sh; r++) vto_hscale_row(y + (size_t)r * y_pitch, my + (size_t)r * dw, dw, lhc, lhp, lht);
    for (int r = 0; r < csh; r++)
        for (int x = 0; x < cdw; x++) {
            int au = 0, av = 0;
            for (int j = 0; j < cht; j++) {
                au += (int)u[(size_t)r * c_pitch + (size_t)(chp[x] + j) * c_step] * chc[(size_t)x * cht + j];
                av += (int)v[(size_t)r * c_pitch + (size_t)(chp[x] + j) * c_step] * chc[(size_t)x * cht + j];
            }
            au >>= 7; av >>= 7;
            mu[(size_t)r * cdw + x] = (int16_t)(au > 32767 ? 32767 : au);
            mv[(size_t)r * cdw + x] = (int16_t)(av > 32767 ? 32767 : av);
        }
    for (int yy = 0; yy < dh; yy++)
        for (int x = 0; x < dw; x++) {
            int Y = 1 << 18, U = 1 << 18, V = 1 << 18;
            for (int j = 0; j < lvt; j++) Y += my[(size_t)(lvp[yy] + j) * dw + x] * lvc[(size_t)yy * lvt + j];
            for (int j = 0; j < cvt; j++) {
                U += mu[(size_t)(cvp[yy] + j) * cdw + x / 2] * cvc[(size_t)yy * cvt + j];
                V += mv[(size_t)(cvp[yy] + j) * cdw + x / 2] * cvc[(size_t)yy * cvt + j];
            }
            Y >>= 19; U >>= 19; V >>= 19;
            int64_t Uc = U < 0 ? 0 : (U > 255 ? 255 : U), Vc = V < 0 ? 0 : (V > 255 ? 255 : V);
            uint8_t *d = dst + ((size_t)yy * dw + x) * 3;
            d[0] = (uint8_t)vto_rgb_T(Y + 326 + ((Vc * crv) >> 16) - (crv >> 9));
            d[1] = (uint8_t)vto_rgb_T(Y + 326 + ((Uc * cgu) >> 16) - (cgu >> 9) + ((Vc * cgv) >> 16) - (cgv >> 9));
            d[2] = (uint8_t)vto_rgb_T(Y + 326 + ((Uc * cbu) >> 16) - (cbu >> 9));
        }
    rc = 0;
done:
    free(lhc); free(lvc); free(chc); free(cvc); free(lhp); free(lvp); free(chp); free(cvp); free(my); free(mu); free(mv);
    return rc;
}
