// vt_swsfilter.cpp -- host-side polyphase filter banks with libswscale's exact integer arithmetic.
//
// The reference resizes with `ffmpeg -vf scale=-2:360` (/root/reference/src/analyzer/content_analyzer.py:198-199),
// i.e. libswscale with SWS_BICUBIC.  To stay within +-1 LSB of that output the GPU kernels must use the very
// same coefficient banks, so this file rebuilds them the way libswscale's filter set-up does (fixed-point
// kernel evaluation in 2^-30 units, near-zero tap trimming at 0.2 %, border folding, per-row normalisation
// with error feedback).  It targets the CPU-independent variant (SWS_BITEXACT): trimmed taps are dropped.
// Sample positions are the centred ones (srcPos = dstPos = 128) that progressive yuv420p gets from the
// scale filter, for luma and chroma alike.
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "../../include/vtseg.h"

namespace {

struct Bank {
    int taps = 0;
    std::vector<int64_t> w;   // dst x taps, 64-bit fixed point
    std::vector<int32_t> pos; // dst
};

inline int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }
inline int ilog2(unsigned v) {
    int n = 0;
    while (v >>= 1) ++n;
    return n;
}

enum Mode { kBicubic, kBilinear, kArea };

bool pick_mode(int flags, Mode *m) {
    if (flags & VT_SWS_BICUBIC) *m = kBicubic;
    else if (flags & VT_SWS_AREA) *m = kArea;
    else if (flags & VT_SWS_BILINEAR) *m = kBilinear;
    else return false;
    return true;
}

int support_factor(Mode m) { return m == kBicubic ? 4 : (m == kBilinear ? 2 : 1); }

int initial_taps(int src, int dst, Mode m, int64_t inc) {
    int t = inc <= 65536 ? 1 + support_factor(m) : 1 + (support_factor(m) * src + dst - 1) / dst;
    t = std::min(t, src - 2);
    return std::max(t, 1);
}

// Kernel weight at distance d (2^-30 source pixels after the downscale stretch), in units of `unit`.
int64_t weight(Mode m, int64_t d, int64_t inc, int64_t unit) {
    switch (m) {
        case kBicubic: {
            // Keys cubic with B = 0, C = 0.6 (libswscale's default parameters), 2^24 fixed point
            const int64_t c = (int64_t)(0.6 * (1 << 24));
            const int64_t k = 1 << 24;
            int64_t v = 0;
            if (d < (1LL << 31)) {
                const int64_t d2 = (d * d) >> 30, d3 = (d2 * d) >> 30;
                if (d < (1LL << 30)) v = (12 * k - 6 * c) * d3 + (6 * c - 18 * k) * d2 + (6 * k) * (1LL << 30);
                else v = (-6 * c) * d3 + (30 * c) * d2 + (-48 * c) * d + (24 * c) * (1LL << 30);
            }
            return v / ((1LL << 54) / unit);
        }
        case kArea: {
            const int64_t e = d - (1 << 29);
            int64_t v;
            if (e * inc < -(1LL << 45)) v = 1LL << 46;
            else if (e * inc < (1LL << 45)) v = (1LL << 45) - e * inc;
            else v = 0;
            return v * (unit >> 46);
        }
        default: {
            const int64_t v = std::max<int64_t>((1 << 30) - d, 0);
            return v * (unit >> 30);
        }
    }
}

bool build_bank(int src, int dst, int flags, int one, Bank *out) {
    Mode mode;
    if (src < 1 || dst < 1 || !pick_mode(flags, &mode)) return false;
    const int64_t inc = (((int64_t)src << 16) + (dst >> 1)) / dst;
    const int ratio = src / dst;
    const int64_t unit = 1LL << (54 - (ratio > 0 ? std::min(ilog2((unsigned)ratio), 8) : 0));

    int taps;
    std::vector<int64_t> w;
    std::vector<int32_t> pos((size_t)dst);

    if (iabs64(inc - 65536) < 10) {  // this axis is not scaled
        taps = 1;
        w.assign((size_t)dst, unit);
        for (int i = 0; i < dst; ++i) pos[i] = i;
    } else if (inc <= 65536 && mode == kArea) {  // area enlargement is linear interpolation
        taps = 2;
        w.resize((size_t)dst * taps);
        int64_t centre = (inc >> 1) - 0x8000;
        for (int i = 0; i < dst; ++i, centre += inc) {
            int first = (int)((centre - ((int64_t)(taps - 1) << 15) + (1 << 15)) >> 16);
            pos[i] = first;
            for (int j = 0; j < taps; ++j) {
                int64_t v = unit - iabs64(((int64_t)(first + j) << 16) - centre) * (unit >> 16);
                w[(size_t)i * taps + j] = std::max<int64_t>(v, 0);
            }
        }
    } else {
        taps = initial_taps(src, dst, mode, inc);
        w.resize((size_t)dst * taps);
        int64_t centre = inc - 65536;  // 17-bit fraction: twice the 16.16 position
        for (int i = 0; i < dst; ++i, centre += 2 * inc) {
            int first = (int)((centre - (int64_t)(taps - 2) * 65536) / (1 << 17));
            pos[i] = first;
            for (int j = 0; j < taps; ++j) {
                int64_t d = iabs64(((int64_t)(first + j) << 17) - centre) << 13;
                if (inc > 65536) d = d * dst / src;
                w[(size_t)i * taps + j] = weight(mode, d, inc, unit);
            }
        }
    }

    // Trim: slide each window right over negligible leading taps (keeping starts monotone), then find the
    // widest span any output still needs.
    const double negligible = 0.002 * (double)unit;
    int need = 1;
    for (int i = dst - 1; i >= 0; --i) {
        int64_t *row = &w[(size_t)i * taps];
        int64_t acc = 0;
        for (int n = 0; n < taps; ++n) {
            acc += iabs64(row[0]);
            if ((double)acc > negligible) break;
            if (i < dst - 1 && pos[i] >= pos[i + 1]) break;
            std::rotate(row, row + 1, row + taps);
            row[taps - 1] = 0;
            ++pos[i];
        }
        int span = taps;
        acc = 0;
        for (int j = taps - 1; j > 0; --j) {
            acc += iabs64(row[j]);
            if ((double)acc > negligible) break;
            --span;
        }
        need = std::max(need, span);
    }
    if (need < taps) {
        for (int i = 0; i < dst; ++i)
            for (int j = 0; j < need; ++j) w[(size_t)i * need + j] = w[(size_t)i * taps + j];
        w.resize((size_t)dst * need);
        taps = need;
    }

    // Windows hanging over an edge are folded back onto the edge sample.
    for (int i = 0; i < dst; ++i) {
        int64_t *row = &w[(size_t)i * taps];
        if (pos[i] < 0) {
            for (int j = 1; j < taps; ++j) {
                int to = std::max(j + pos[i], 0);
                row[to] += row[j];
                row[j] = 0;
            }
            pos[i] = 0;
        }
        if (pos[i] + taps > src) {
            const int shift = pos[i] + std::min(taps - src, 0);
            int64_t spill = 0;
            for (int j = taps - 1; j >= 0; --j)
                if (pos[i] + j >= src) {
                    spill += row[j];
                    row[j] = 0;
                }
            for (int j = taps - 1; j >= 0; --j) row[j] = j < shift ? 0 : row[j - shift];
            pos[i] -= shift;
            row[src - 1 - pos[i]] += spill;
        }
    }

    out->taps = taps;
    out->pos = std::move(pos);
    out->w = std::move(w);
    (void)one;
    return true;
}

}  // namespace

extern "C" int vt_sws_max_taps(int src_size, int dst_size, int flags) {
    Mode m;
    if (src_size < 1 || dst_size < 1 || !pick_mode(flags, &m)) return VT_ERR_INVALID;
    const int64_t inc = (((int64_t)src_size << 16) + (dst_size >> 1)) / dst_size;
    if (inc <= 65536 && m == kArea) return 2;
    return initial_taps(src_size, dst_size, m, inc);
}

extern "C" int vt_sws_make_filter(int src_size, int dst_size, int flags, int one, int16_t *coef, int32_t *pos,
                                  int *taps) {
    Bank b;
    if (!coef || !pos || !taps || one <= 0 || !build_bank(src_size, dst_size, flags, one, &b)) return VT_ERR_INVALID;
    for (int i = 0; i < dst_size; ++i) {
        const int64_t *row = &b.w[(size_t)i * b.taps];
        int64_t total = 0;
        for (int j = 0; j < b.taps; ++j) total += row[j];
        total = (total + one / 2) / one;
        if (!total) total = 1;
        int64_t carry = 0;
        for (int j = 0; j < b.taps; ++j) {
            const int64_t v = row[j] + carry;
            const int64_t q = (v >= 0 ? v + (total >> 1) : v - (total >> 1)) / total;  // round half away from 0
            coef[(size_t)i * b.taps + j] = (int16_t)q;
            carry = v - q * total;
        }
        pos[i] = b.pos[i];
    }
    *taps = b.taps;
    return VT_OK;
}

extern "C" int vt_scale_width_for_height(int src_w, int src_h, int dst_h) {
    if (src_w <= 0 || src_h <= 0 || dst_h <= 0) return VT_ERR_INVALID;
    // scale=-2:H : the scale filter evaluates w = av_rescale(H, src_w, src_h * 2) * 2, and av_rescale rounds
    // to nearest with halves away from zero.
    const int64_t den = (int64_t)src_h * 2;
    int64_t r = ((int64_t)dst_h * src_w + den / 2) / den;
    if (r < 1) r = 1;
    return (int)(r * 2);
}
