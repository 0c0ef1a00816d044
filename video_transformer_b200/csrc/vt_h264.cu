// vt_h264.cu -- K0 on this pool: H.264 "PCM-intra" decode front end.
//
// The north star puts NVDEC here.  On this pool the driver runs behind a paravirtual proxy that exposes
// compute only (cuvidCreateDecoder -> CUDA_ERROR_NO_DEVICE, see DESIGN.md / tools/probe_nvdec.py), and the
// image has no H.264 encoder, so the only H.264 streams that exist here are the synthetic ones our
// generator writes: Constrained-Baseline, CAVLC, one slice per picture, IDR pictures made of I_PCM
// macroblocks and P pictures that are one mb_skip_run.  For exactly that
// subset the decode is HBM-bound byte re-layout (macroblock order -> raster NV12), which is what the
// kernel below does; anything else is VT_ERR_UNSUPPORTED (never a CPU decode).  Parity: bit-exact against
// libavcodec's software decoder on the same bitstream (tests/test_decode.py).
//
// Replaces: libavcodec inside the ffmpeg child (/root/reference/src/utils/video_segmenter.py:141-154,
// /root/reference/src/analyzer/content_analyzer.py:193-211) and ffprobe's stream scan
// (/root/reference/src/utils/video_utils.py:7-38).
#include <string.h>

#include <vector>

#include <string.h>

#include <algorithm>

#include "vt_common.cuh"

namespace {

// ---- bit reader over an RBSP (emulation prevention already removed) ----------------------------------------
struct Bits {
    const uint8_t *p;
    size_t n;      // bytes
    size_t bit = 0;
    bool bad = false;
    Bits(const uint8_t *d, size_t len) : p(d), n(len) {}
    uint32_t u(int k) {
        uint32_t v = 0;
        for (int i = 0; i < k; i++) {
            if ((bit >> 3) >= n) { bad = true; return 0; }
            v = (v << 1) | ((p[bit >> 3] >> (7 - (bit & 7))) & 1u);
            bit++;
        }
        return v;
    }
    uint32_t ue() {
        int z = 0;
        while (!bad && u(1) == 0) if (++z > 31) { bad = true; return 0; }
        if (bad) return 0;
        return z ? ((1u << z) - 1u + u(z)) : 0;
    }
    int32_t se() {
        uint32_t k = ue();
        return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
    }
};

// Copy up to `cap` RBSP bytes of a NAL payload, dropping emulation prevention bytes.  *consumed_map, when
// given, receives for every RBSP byte its offset in the raw payload (to map header length back to the file).
size_t unescape(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, uint32_t *raw_off) {
    size_t o = 0;
    int zeros = 0;
    for (size_t i = 0; i < n && o < cap; i++) {
        if (zeros >= 2 && src[i] == 3) { zeros = 0; continue; }
        if (raw_off) raw_off[o] = (uint32_t)i;
        dst[o++] = src[i];
        zeros = src[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}

struct Sps {
    bool valid = false;
    int profile = 0, level = 0;
    int log2_max_frame_num = 4, poc_type = 0, log2_max_poc_lsb = 4;
    int mb_w = 0, mb_h = 0, frame_mbs_only = 1;
    int crop_l = 0, crop_r = 0, crop_t = 0, crop_b = 0;
    int fps_num = 0, fps_den = 0;
    int chroma_format_idc = 1;
    bool delta_pic_order_always_zero = false;
};
struct Pps {
    bool valid = false;
    int entropy_cabac = 0, bottom_field_pic_order = 0, slice_groups = 1, weighted_pred = 0;
    int deblock_ctrl = 0, redundant_pic_cnt = 0, num_ref_l0 = 1;
};

bool parse_sps(const uint8_t *nal, size_t n, Sps *s) {
    uint8_t buf[256];
    size_t len = unescape(nal + 1, n - 1, buf, sizeof(buf), nullptr);
    Bits b(buf, len);
    s->profile = b.u(8);
    b.u(8);
    s->level = b.u(8);
    b.ue();  // sps id
    if (s->profile == 100 || s->profile == 110 || s->profile == 122 || s->profile == 244 || s->profile == 44 ||
        s->profile == 83 || s->profile == 86 || s->profile == 118 || s->profile == 128) {
        s->chroma_format_idc = b.ue();
        if (s->chroma_format_idc == 3) b.u(1);
        int bd_l = b.ue(), bd_c = b.ue();
        b.u(1);
        if (b.u(1)) return false;  // scaling matrices: outside the subset
        if (bd_l || bd_c) return false;
    }
    s->log2_max_frame_num = b.ue() + 4;
    s->poc_type = b.ue();
    if (s->poc_type == 0) s->log2_max_poc_lsb = b.ue() + 4;
    else if (s->poc_type == 1) {
        s->delta_pic_order_always_zero = b.u(1);
        b.se(); b.se();
        uint32_t k = b.ue();
        for (uint32_t i = 0; i < k && !b.bad; i++) b.se();
    }
    b.ue();  // max_num_ref_frames
    b.u(1);  // gaps
    s->mb_w = b.ue() + 1;
    s->mb_h = b.ue() + 1;
    s->frame_mbs_only = b.u(1);
    if (!s->frame_mbs_only) b.u(1);
    b.u(1);  // direct_8x8
    if (b.u(1)) { s->crop_l = b.ue(); s->crop_r = b.ue(); s->crop_t = b.ue(); s->crop_b = b.ue(); }
    if (b.u(1)) {  // VUI
        if (b.u(1)) { if (b.u(8) == 255) { b.u(16); b.u(16); } }
        if (b.u(1)) b.u(1);
        if (b.u(1)) { b.u(3); b.u(1); if (b.u(1)) { b.u(8); b.u(8); b.u(8); } }
        if (b.u(1)) { b.ue(); b.ue(); }
        if (b.u(1)) {
            uint32_t units = b.u(32), scale = b.u(32);
            b.u(1);
            if (units && scale) { s->fps_num = (int)scale; s->fps_den = (int)(units * 2); }
        }
    }
    s->valid = !b.bad && s->mb_w > 0 && s->mb_h > 0;
    return s->valid;
}

bool parse_pps(const uint8_t *nal, size_t n, Pps *p) {
    uint8_t buf[64];
    size_t len = unescape(nal + 1, n - 1, buf, sizeof(buf), nullptr);
    Bits b(buf, len);
    b.ue(); b.ue();
    p->entropy_cabac = b.u(1);
    p->bottom_field_pic_order = b.u(1);
    p->slice_groups = b.ue() + 1;
    if (p->slice_groups != 1) return false;
    p->num_ref_l0 = b.ue() + 1;
    b.ue();
    p->weighted_pred = b.u(1);
    b.u(2);
    b.se(); b.se(); b.se();
    p->deblock_ctrl = b.u(1);
    b.u(1);
    p->redundant_pic_cnt = b.u(1);
    p->valid = !b.bad;
    return p->valid;
}

struct SliceInfo {
    bool ok = false;      // header parsed and inside the PCM-intra subset
    bool idr = false, all_skip = false, pcm = false;
    size_t payload = 0;   // raw byte offset (from NAL start) of macroblock 0's first sample
};

// Parse one slice NAL far enough to classify it.  `nal` points at the NAL header byte.
SliceInfo parse_slice(const uint8_t *nal, size_t n, const Sps &sps, const Pps &pps) {
    SliceInfo si;
    const int type = nal[0] & 31, ref_idc = (nal[0] >> 5) & 3;
    si.idr = type == 5;
    uint8_t buf[96];
    uint32_t raw[96];
    size_t len = unescape(nal + 1, n - 1, buf, sizeof(buf), raw);
    Bits b(buf, len);
    const int mbs = sps.mb_w * sps.mb_h;
    if (b.ue() != 0) return si;  // first_mb_in_slice: one slice per picture
    uint32_t st = b.ue() % 5;    // 0 P, 2 I
    b.ue();                      // pps id
    b.u(sps.log2_max_frame_num);
    if (!sps.frame_mbs_only) return si;
    if (si.idr) b.ue();
    if (sps.poc_type == 0) { b.u(sps.log2_max_poc_lsb); if (pps.bottom_field_pic_order) b.se(); }
    else if (sps.poc_type == 1 && !sps.delta_pic_order_always_zero) { b.se(); if (pps.bottom_field_pic_order) b.se(); }
    if (pps.redundant_pic_cnt) b.ue();
    if (st == 0) {
        if (b.u(1)) b.ue();     // num_ref_idx_active_override
        if (b.u(1)) return si;  // ref_pic_list_modification: outside the subset
        if (pps.weighted_pred) return si;
    } else if (st != 2) {
        return si;
    }
    if (ref_idc) {
        if (si.idr) { b.u(1); b.u(1); }
        else if (b.u(1)) return si;  // adaptive marking: outside the subset
    }
    if (pps.entropy_cabac) return si;
    b.se();  // slice_qp_delta
    if (pps.deblock_ctrl) {
        uint32_t idc = b.ue();
        if (idc != 1) { b.se(); b.se(); }
    }
    if (b.bad) return si;
    if (st == 0) {
        // P picture: exactly one mb_skip_run covering the picture, then the RBSP stop bit.
        if ((int)b.ue() != mbs || b.bad) return si;
        if (b.u(1) != 1) return si;
        si.all_skip = true;
        si.ok = true;
        return si;
    }
    // I picture: mb_type must be I_PCM (25); samples start at the next byte boundary.
    if (b.ue() != 25 || b.bad) return si;
    size_t byte = (b.bit + 7) >> 3;
    if (byte >= len) return si;
    si.payload = 1 + (size_t)raw[byte];  // +1: NAL header byte
    // Fixed layout check: 384 sample bytes per macroblock, "0D 00" (ue(25) + alignment) between them, then
    // the stop byte 0x80.  Any emulation prevention byte inside would change the length.
    const size_t expect = si.payload + (size_t)mbs * 386 - 2 + 1;
    if (n != expect) return si;
    si.pcm = true;
    si.ok = true;
    return si;
}

struct Nal { size_t off, size; };  // off = NAL header byte, size excludes trailing zeros

void split_annexb(const uint8_t *d, size_t n, std::vector<Nal> &out) {
    size_t i = 0, start = (size_t)-1;
    while (i + 3 <= n) {
        const uint8_t *z = (const uint8_t *)memchr(d + i, 0, n - i);
        if (!z) break;
        i = (size_t)(z - d);
        if (i + 3 <= n && d[i + 1] == 0 && d[i + 2] == 1) {
            if (start != (size_t)-1) {
                size_t end = i;
                while (end > start && d[end - 1] == 0) end--;
                out.push_back({start, end - start});
            }
            start = i + 3;
            i += 3;
        } else {
            i++;
        }
    }
    if (start != (size_t)-1 && start < n) {
        size_t end = n;
        while (end > start && d[end - 1] == 0) end--;
        out.push_back({start, end - start});
    }
}

}  // namespace

extern "C" int vt_h264_scan(const uint8_t *bs, size_t n, vt_stream_info *info, uint64_t *frame_offsets,
                            uint32_t *frame_sizes, uint32_t *frame_flags, int max_frames) {
    if (!bs || !info || n < 8) {
        vt::set_error("vt_h264_scan: bad arguments");
        return VT_ERR_INVALID;
    }
    memset(info, 0, sizeof(*info));
    std::vector<Nal> nals;
    split_annexb(bs, n, nals);
    if (nals.empty()) {
        vt::set_error("vt_h264_scan: no Annex-B start code found");
        return VT_ERR_BITSTREAM;
    }
    Sps sps;
    Pps pps;
    int frames = 0, idr = 0;
    bool subset = true;
    for (const Nal &nl : nals) {
        if (!nl.size) continue;
        const int type = bs[nl.off] & 31;
        if (type == 7) {
            Sps s2;
            if (!parse_sps(bs + nl.off, nl.size, &s2)) { subset = false; continue; }
            if (sps.valid && (s2.mb_w != sps.mb_w || s2.mb_h != sps.mb_h)) {
                vt::set_error("vt_h264_scan: resolution change mid-stream is not supported");
                return VT_ERR_UNSUPPORTED;
            }
            sps = s2;
        } else if (type == 8) {
            if (!parse_pps(bs + nl.off, nl.size, &pps)) subset = false;
        } else if (type == 1 || type == 5) {
            if (!sps.valid || !pps.valid) {
                vt::set_error("vt_h264_scan: slice before SPS/PPS");
                return VT_ERR_BITSTREAM;
            }
            SliceInfo si = parse_slice(bs + nl.off, nl.size, sps, pps);
            if (!si.ok) subset = false;
            if (frames < max_frames) {
                if (frame_offsets) frame_offsets[frames] = nl.off;
                if (frame_sizes) frame_sizes[frames] = (uint32_t)nl.size;
                if (frame_flags) frame_flags[frames] = (si.idr ? 1u : 0u) | (si.all_skip ? 2u : 0u) | (si.ok ? 0u : 4u);
            }
            frames++;
            if (si.idr) idr++;
        }
    }
    if (!sps.valid) {
        vt::set_error("vt_h264_scan: no parsable SPS");
        return VT_ERR_BITSTREAM;
    }
    info->codec = 4;
    info->coded_width = sps.mb_w * 16;
    info->coded_height = sps.mb_h * 16;
    info->width = info->coded_width - 2 * (sps.crop_l + sps.crop_r);
    info->height = info->coded_height - 2 * (sps.crop_t + sps.crop_b);
    info->fps_num = sps.fps_num;
    info->fps_den = sps.fps_den;
    info->n_frames = frames;
    info->n_idr = idr;
    info->pcm_intra_only = (subset && sps.chroma_format_idc == 1 && frames > 0) ? 1 : 0;
    return VT_OK;
}

namespace {
int layout_core(const uint8_t *bs, size_t n, const Sps &sps, const Pps &pps, const uint64_t *frame_offsets,
                const uint32_t *frame_sizes, int n_frames, uint64_t *payload_off) {
    uint64_t last_idr = UINT64_MAX;
    for (int f = 0; f < n_frames; f++) {
        if (frame_sizes[f] < 2 || frame_offsets[f] > n || frame_sizes[f] > n - frame_offsets[f]) {
            vt::set_error("vt_h264_pcm_layout: frame %d is empty or outside the buffer", f);
            return VT_ERR_BITSTREAM;
        }
        SliceInfo si = parse_slice(bs + frame_offsets[f], frame_sizes[f], sps, pps);
        if (!si.ok) {
            vt::set_error("vt_h264_pcm_layout: picture %d is outside the PCM-intra subset "
                          "(needs NVDEC, which this driver refuses)", f);
            return VT_ERR_UNSUPPORTED;
        }
        if (si.pcm) last_idr = frame_offsets[f] + si.payload;
        payload_off[f] = last_idr;  // skip pictures repeat the picture they reference
    }
    return VT_OK;
}
}  // namespace

extern "C" int vt_h264_pcm_layout(const uint8_t *bs, size_t n, const uint64_t *frame_offsets,
                                  const uint32_t *frame_sizes, int n_frames, uint64_t *payload_off) {
    if (!bs || !frame_offsets || !frame_sizes || !payload_off || n_frames <= 0) {
        vt::set_error("vt_h264_pcm_layout: bad arguments");
        return VT_ERR_INVALID;
    }
    // parameter sets: the first SPS/PPS of the stream (vt_h264_scan rejected mid-stream changes)
    std::vector<Nal> head;
    size_t head_len = frame_offsets[0] < n ? (size_t)frame_offsets[0] : n;
    split_annexb(bs, head_len + 4 <= n ? head_len + 4 : n, head);
    Sps sps;
    Pps pps;
    for (const Nal &nl : head) {
        if (!nl.size) continue;
        const int type = bs[nl.off] & 31;
        if (type == 7 && !sps.valid) parse_sps(bs + nl.off, nl.size, &sps);
        if (type == 8 && !pps.valid) parse_pps(bs + nl.off, nl.size, &pps);
    }
    if (!sps.valid || !pps.valid) {
        vt::set_error("vt_h264_pcm_layout: no SPS/PPS ahead of the first picture");
        return VT_ERR_BITSTREAM;
    }
    return layout_core(bs, n, sps, pps, frame_offsets, frame_sizes, n_frames, payload_off);
}

extern "C" int vt_h264_pcm_layout_ps(const uint8_t *bs, size_t n, const uint8_t *sps_nal, size_t sps_len,
                                     const uint8_t *pps_nal, size_t pps_len, const uint64_t *frame_offsets,
                                     const uint32_t *frame_sizes, int n_frames, uint64_t *payload_off) {
    if (!bs || !sps_nal || !pps_nal || sps_len < 4 || pps_len < 2 || !frame_offsets || !frame_sizes ||
        !payload_off || n_frames <= 0) {
        vt::set_error("vt_h264_pcm_layout_ps: bad arguments");
        return VT_ERR_INVALID;
    }
    Sps sps;
    Pps pps;
    if (!parse_sps(sps_nal, sps_len, &sps) || !parse_pps(pps_nal, pps_len, &pps)) {
        vt::set_error("vt_h264_pcm_layout_ps: parameter sets outside the supported subset");
        return VT_ERR_UNSUPPORTED;
    }
    return layout_core(bs, n, sps, pps, frame_offsets, frame_sizes, n_frames, payload_off);
}

// ---- device side: macroblock-ordered PCM samples -> NV12 raster --------------------------------------------
namespace vt {

// One block per (frame, macroblock row).  The row's samples (mb_w x 386 bytes, contiguous in the stream) are
// read once with 128-bit loads into shared memory, then written as 16 luma rows and 8 interleaved chroma
// rows with 128-bit stores.  Unaligned shared reads are assembled with funnel shifts.
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t *s, uint32_t off) {
    const uint32_t *w = (const uint32_t *)(s + (off & ~3u));
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

// Payload offsets of the pictures of one launch travel as a kernel parameter (constant bank): no device allocation,
// no extra copy, nothing to free -- the call stays a single asynchronous launch.
constexpr int PCM_FRAMES_PER_LAUNCH = 1024;
struct PcmOffsets {
    uint64_t off[PCM_FRAMES_PER_LAUNCH];
};

__global__ void __launch_bounds__(256)
h264_pcm_kernel(const uint8_t *__restrict__ bs, const __grid_constant__ PcmOffsets payload, int mb_w, int mb_h,
                int width, int height, const uint8_t *__restrict__ prev, uint8_t *__restrict__ out, int pitch,
                size_t frame_stride) {
    extern __shared__ __align__(16) uint8_t srow[];
    const int f = blockIdx.x / mb_h, my = blockIdx.x - f * mb_h;
    const int tid = threadIdx.x;
    uint8_t *dst = out + (size_t)f * frame_stride;
    const uint64_t p0 = payload.off[f];
    const int wq = (width + 15) >> 4;  // 16 B groups per output row (pitch >= 16*wq is checked on the host)
    if (p0 == UINT64_MAX) {
        // skip picture with no IDR in this batch: repeat the carried-over surface
        for (int i = tid; i < 24 * wq; i += blockDim.x) {
            const int r = i / wq, g = i - r * wq;
            size_t o;
            if (r < 16) {
                const int y = my * 16 + r;
                if (y >= height) continue;
                o = (size_t)y * pitch + (size_t)g * 16;
            } else {
                const int y = my * 8 + (r - 16);
                if (y >= (height + 1) / 2) continue;
                o = (size_t)pitch * height + (size_t)y * pitch + (size_t)g * 16;
            }
            st_stream_u4(dst + o, ld_stream_u4(prev + o));
        }
        return;
    }
    const uint64_t start = p0 + (uint64_t)my * mb_w * 386;
    const uint32_t mis = (uint32_t)(start & 15);
    const uint8_t *src = bs + (start - mis);
    const int nbytes = mb_w * 386 - 2 + (int)mis;
    for (int i = tid; i < (nbytes + 15) / 16 + 1; i += blockDim.x)
        *(uint4 *)(srow + (size_t)i * 16) = ld_stream_u4(src + (size_t)i * 16);
    __syncthreads();
    for (int i = tid; i < 24 * mb_w; i += blockDim.x) {
        const int r = i / mb_w, mx = i - r * mb_w;
        if (mx * 16 >= width) continue;
        const uint32_t mb0 = mis + (uint32_t)mx * 386u;
        uint4 v;
        size_t o;
        if (r < 16) {
            const int y = my * 16 + r;
            if (y >= height) continue;
            const uint32_t a = mb0 + (uint32_t)r * 16u;
            v.x = lds_u32_unaligned(srow, a);
            v.y = lds_u32_unaligned(srow, a + 4);
            v.z = lds_u32_unaligned(srow, a + 8);
            v.w = lds_u32_unaligned(srow, a + 12);
            o = (size_t)y * pitch + (size_t)mx * 16;
        } else {
            const int cr = r - 16, y = my * 8 + cr;
            if (y >= (height + 1) / 2) continue;
            const uint32_t a = mb0 + 256u + (uint32_t)cr * 8u;
            const uint32_t b0 = lds_u32_unaligned(srow, a), b1 = lds_u32_unaligned(srow, a + 4);
            const uint32_t r0 = lds_u32_unaligned(srow, a + 64), r1 = lds_u32_unaligned(srow, a + 68);
            v.x = __byte_perm(b0, r0, 0x5140);
            v.y = __byte_perm(b0, r0, 0x7362);
            v.z = __byte_perm(b1, r1, 0x5140);
            v.w = __byte_perm(b1, r1, 0x7362);
            o = (size_t)pitch * height + (size_t)y * pitch + (size_t)mx * 16;
        }
        st_stream_u4(dst + o, v);
    }
}

}  // namespace vt

extern "C" int vt_h264_pcm_decode(const uint8_t *bs_dev, const uint64_t *payload_off, int n_frames, int width,
                                  int height, const uint8_t *prev_dev, uint8_t *nv12_dev, int pitch,
                                  size_t frame_stride, void *stream) {
    if (!bs_dev || !payload_off || !nv12_dev || n_frames <= 0 || width <= 0 || height <= 0) {
        vt::set_error("vt_h264_pcm_decode: bad arguments");
        return VT_ERR_INVALID;
    }
    const int mb_w = (width + 15) / 16, mb_h = (height + 15) / 16;
    if (pitch < mb_w * 16 || pitch % 16 || frame_stride % 16 || (uintptr_t)nv12_dev % 16 ||
        (prev_dev && (uintptr_t)prev_dev % 16)) {
        vt::set_error("vt_h264_pcm_decode: surfaces need pitch >= %d, 16-byte aligned", mb_w * 16);
        return VT_ERR_INVALID;
    }
    for (int f = 0; f < n_frames; f++)
        if (payload_off[f] == UINT64_MAX && !prev_dev) {
            vt::set_error("vt_h264_pcm_decode: picture %d repeats a picture from before this batch but no "
                          "previous surface was given", f);
            return VT_ERR_INVALID;
        }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)mb_w * 386 + 64;
    static size_t smem_set_dev[VT_MAX_DEVICES] = {0};   // function attributes are per device
    size_t &smem_set = smem_set_dev[vt::current_device()];
    if (smem > 48 * 1024 && smem > smem_set) {
        VT_CUDA(cudaFuncSetAttribute(vt::h264_pcm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    for (int f0 = 0; f0 < n_frames; f0 += vt::PCM_FRAMES_PER_LAUNCH) {
        const int nf = std::min(vt::PCM_FRAMES_PER_LAUNCH, n_frames - f0);
        vt::PcmOffsets po;
        memcpy(po.off, payload_off + f0, sizeof(uint64_t) * (size_t)nf);
        vt::h264_pcm_kernel<<<(unsigned)(nf * mb_h), 256, smem, st>>>(bs_dev, po, mb_w, mb_h, width, height, prev_dev,
                                                                     nv12_dev + (size_t)f0 * frame_stride, pitch,
                                                                     frame_stride);
        VT_LAUNCHED("h264_pcm_kernel");
    }
    return VT_OK;
}

// The per-batch body of the ingest pass as one C call (see include/vtseg.h).
extern "C" int vt_ingest_batch_pcm(const vt_scale_plan *plan, const uint8_t *bs_dev, const uint64_t *payload_off,
                                   int n_frames, int width, int height, const uint8_t *prev_dev, uint8_t *nv12_dev,
                                   int pitch, size_t surface_bytes, uint64_t *sad_dev, uint32_t *hist_dev,
                                   uint8_t *out_dev, size_t out_frame_bytes, void *stream) {
    int rc = vt_h264_pcm_decode(bs_dev, payload_off, n_frames, width, height, prev_dev, nv12_dev, pitch, surface_bytes, stream);
    if (rc) return rc;
    if (plan && out_dev)                      // K2 + K3 (one luma pass where the plan fuses them)
        return vt_scale_score_nv12_to_yuv420p(plan, nv12_dev, pitch, surface_bytes, prev_dev, out_dev, out_frame_bytes,
                                              n_frames, sad_dev, hist_dev, stream);
    rc = vt_sad_hist_u8(nv12_dev, pitch, surface_bytes, width, height, prev_dev, n_frames, sad_dev, hist_dev, stream);
    if (rc || !out_dev) return rc;
    if (plan) return vt_scale_nv12_to_yuv420p(plan, nv12_dev, pitch, surface_bytes, out_dev, out_frame_bytes, n_frames, stream);
    return vt_nv12_to_yuv420p(nv12_dev, pitch, surface_bytes, width, height, out_dev, out_frame_bytes, n_frames, stream);
}
