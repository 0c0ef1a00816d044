/*
 * vtseg.h -- C ABI of libvtseg.so, the B200 (sm_100a) implementation of the long-video ingest hot path.
 *
 * The reference (shizhenneko/Video-Transformer) has no FFI: its seam is three subprocess calls to the
 * ffmpeg/ffprobe binaries made from Python.  Each entry point below names the reference call site whose
 * work it replaces; paths are relative to /root/reference.  The Python host module
 * video_transformer_b200/video_segmenter.py binds these with ctypes.
 *
 * Conventions
 *   - every function returns VT_OK (0) or a negative VT_ERR_* code and never throws;
 *     vt_last_error() returns a thread-local message for the last failure;
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g. torch.Tensor.data_ptr());
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *     on that stream unless stated otherwise;
 *   - no torch types, no C++ types.
 */
#ifndef VTSEG_H
#define VTSEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_OK 0
#define VT_ERR_INVALID (-1)     /* bad argument */
#define VT_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define VT_ERR_UNSUPPORTED (-3) /* stream/shape outside what this build handles */
#define VT_ERR_BITSTREAM (-4)   /* malformed or truncated bitstream */
#define VT_ERR_NOMEM (-5)
#define VT_ERR_NVDEC (-6)       /* libnvcuvid missing or the driver refuses video decode */

/* libswscale flag values (the reference passes none => ffmpeg's scale filter default SWS_BICUBIC,
 * src/analyzer/content_analyzer.py:198-199). */
#define VT_SWS_BILINEAR 2
#define VT_SWS_BICUBIC 4
#define VT_SWS_AREA 0x20

int vt_version(void);
const char *vt_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
uint64_t vt_launch_count(void);

/* ---- K2 host side: polyphase filter bank, libswscale-exact (SWS_ACCURATE_RND|SWS_BITEXACT path) -------
 * Replaces: libswscale's filter set-up inside `ffmpeg -vf scale=-2:360`
 * (src/analyzer/content_analyzer.py:193-211).  `one` = 1<<14 (horizontal) or 1<<12 (vertical). */
int vt_sws_max_taps(int src_size, int dst_size, int flags);
int vt_sws_make_filter(int src_size, int dst_size, int flags, int one, int16_t *coef /* dst*max_taps */,
                       int32_t *pos /* dst */, int *taps);
/* `-2:H` semantics of the scale filter: width from aspect ratio, rounded to a multiple of 2. */
int vt_scale_width_for_height(int src_w, int src_h, int dst_h);

/* ---- K2 device side: one scaling plan = the four filter banks (luma/chroma x h/v) in device memory ---- */
typedef struct vt_scale_plan vt_scale_plan;
int vt_scale_plan_create(int src_w, int src_h, int dst_w, int dst_h, int flags, vt_scale_plan **out);
void vt_scale_plan_destroy(vt_scale_plan *plan);

/* Which kernel vt_scale_nv12_to_yuv420p will use for a plane kind: info8 = {streaming ok, dp2a pairs, vertical taps,
 * columns per lane, output rows per item, tile width, tile height, shared bytes per warp}. */
int vt_scale_plan_stream_info(const vt_scale_plan *plan, int chroma, int *info8);

/* One 8-bit plane, any ratio (generic kernel; used for parity sweeps and odd shapes). */
int vt_scale_plane_u8(const vt_scale_plan *plan, int chroma, const uint8_t *src_dev, int src_pitch,
                      uint8_t *dst_dev, int dst_pitch, void *stream);

/* A batch of NV12 frames -> planar YUV420P at the plan's destination size.
 *   src: frame f at src_dev + f*src_frame_stride; Y plane h rows of src_pitch bytes, UV plane follows at
 *        src_dev + f*src_frame_stride + src_pitch*src_h (the NVDEC surface layout).
 *   dst: frame f at dst_dev + f*dst_frame_stride; tightly packed Y (dw*dh), U, V (cw*ch each).
 * Replaces: decode->swscale->yuv420p inside `ffmpeg -i IN -vf scale=-2:360 ...`
 * (src/analyzer/content_analyzer.py:193-211). */
int vt_scale_nv12_to_yuv420p(const vt_scale_plan *plan, const uint8_t *src_dev, int src_pitch,
                             size_t src_frame_stride, uint8_t *dst_dev, size_t dst_frame_stride, int n_frames,
                             void *stream);

/* K2 + K3 in one call: the scaled frames of vt_scale_nv12_to_yuv420p plus the SAD / histogram of vt_sad_hist_u8 on the
 * SOURCE luma (prev0_dev: luma of the picture before frame 0, or NULL).  With VT_FUSED_SCORE=1 in the environment and a
 * plan that allows it (vt_scale_plan_fuses_score == 1: exact 3:2 luma, e.g. 1080p -> 720p) the luma pass of the scaler
 * counts the source rows it already holds in shared memory, so the source luma is fetched from HBM once (20 % less DRAM
 * traffic, same time -- see DESIGN.md); otherwise the two kernels run one after the other.  Results are identical.  Replaces the same reference work as the two calls it combines. */
int vt_scale_plan_fuses_score(const vt_scale_plan *plan);
int vt_scale_score_nv12_to_yuv420p(const vt_scale_plan *plan, const uint8_t *src_dev, int src_pitch, size_t src_frame_stride,
                                   const uint8_t *prev0_dev, uint8_t *dst_dev, size_t dst_frame_stride, int n_frames,
                                   uint64_t *sad_dev, uint32_t *hist_dev, void *stream);

/* ---- K1: NV12 -> planar YUV420P (same size) and NV12 -> packed RGB24 (swscale nv12->rgb24 semantics) ---- */
int vt_nv12_to_yuv420p(const uint8_t *src_dev, int src_pitch, size_t src_frame_stride, int w, int h,
                       uint8_t *dst_dev, size_t dst_frame_stride, int n_frames, void *stream);
int vt_nv12_to_rgb24(const uint8_t *src_dev, int src_pitch, size_t src_frame_stride, int w, int h,
                     uint8_t *dst_dev, size_t dst_frame_stride, int n_frames, void *stream);

/* ---- K1b + K2 for the upload product (BASELINE.json configs[4]: sampled frames -> 768x768 RGB) -------------
 * NV12 -> packed RGB24 at a different size, with the semantics of `ffmpeg -vf scale=W:H -pix_fmt rgb24`
 * (libswscale general path, SWS_BICUBIC, BT.601 limited range; chroma interpolated vertically to full height,
 * shared by horizontal pixel pairs).  The reference has no RGB anywhere; the nearest call
 * site is the upload-size reducer src/analyzer/content_analyzer.py:167-236.  dst_w must be even.
 *   dst: frame f at dst_dev + f*dst_frame_stride, dst_h rows of 3*dst_w bytes (R,G,B). */
typedef struct vt_rgb_plan vt_rgb_plan;
int vt_rgb_plan_create(int src_w, int src_h, int dst_w, int dst_h, int flags, vt_rgb_plan **out);
void vt_rgb_plan_destroy(vt_rgb_plan *plan);
int vt_scale_nv12_to_rgb24(const vt_rgb_plan *plan, const uint8_t *src_dev, int src_pitch, size_t src_frame_stride,
                           uint8_t *dst_dev, size_t dst_frame_stride, int n_frames, void *stream);

/* ---- K3: per-frame 256-bin luma histogram and SAD against the previous frame ---------------------------
 * Frame f's luma is at luma_dev + f*frame_stride (h rows of `pitch` bytes, display width w).
 * prev for frame 0 is prev0_dev (same pitch) or NULL (then sad[0] = 0); prev for f>0 is frame f-1.
 * sad_dev: n_frames x u64, hist_dev: n_frames x 256 x u32; both are overwritten.
 * No reference counterpart: this is the scene-change measurement the north star
 * adds at the snap_to_keyframe hook (src/utils/video_segmenter.py:157-159). */
int vt_sad_hist_u8(const uint8_t *luma_dev, int pitch, size_t frame_stride, int w, int h,
                   const uint8_t *prev0_dev, int n_frames, uint64_t *sad_dev, uint32_t *hist_dev, void *stream);

/* ---- K5: gather frames [first, first+count) of a strided batch into one contiguous buffer -------------
 * Replaces: the per-segment artefact of extract_segment (src/utils/video_segmenter.py:86-154). */
int vt_gather_frames(const uint8_t *src_dev, size_t src_frame_stride, size_t frame_bytes, const int32_t *index_dev,
                     int count, uint8_t *dst_dev, void *stream);

/* ---- upload-size reducer: baseline JPEG of planar YUV420P pictures (Motion-JPEG samples) ------------------------------
 * Replaces: the libx264 encode inside `ffmpeg -vf scale=-2:360 -c:v libx264 -crf 28` of
 * ContentAnalyzer._compress_video_for_upload (src/analyzer/content_analyzer.py:193-217); a B200 has no video encoder.
 * ITU-T T.81 baseline, 4:2:0, Annex-K Huffman tables, IJG integer DCT and quality scaling, one restart interval per MCU
 * row.  expand_range = 1 expands limited-range video samples (16..235 / 16..240) to JFIF's full range first.
 *   src: picture f at src_dev + f*src_frame_stride, tightly packed Y (w*h), U, V (ceil(w/2)*ceil(h/2) each)
 *   out: pictures packed back to back; offsets_dev[f] = first byte of picture f, offsets_dev[n_frames] = total bytes
 *   status_dev: 0 ok, 1 an MCU row did not compress below its raw size, 2 out_cap too small (then nothing is valid) */
typedef struct vt_jpeg_plan vt_jpeg_plan;
int vt_jpeg_plan_create(int w, int h, int quality, int expand_range, vt_jpeg_plan **out);
void vt_jpeg_plan_destroy(vt_jpeg_plan *plan);
size_t vt_jpeg_max_frame_bytes(const vt_jpeg_plan *plan);
int vt_jpeg_header(const vt_jpeg_plan *plan, uint8_t *out, size_t cap, size_t *len);
int vt_jpeg_encode_yuv420p(vt_jpeg_plan *plan, const uint8_t *src_dev, size_t src_frame_stride, int n_frames,
                           uint8_t *out_dev, size_t out_cap, uint64_t *offsets_dev, int32_t *status_dev, void *stream);

/* Landing of K5: frames go from the device straight into a page-locked mapping of the segment's `.frames` file.
 * vt_host_register pins an existing host range (VT_ERR_CUDA when the range cannot be pinned; the CUDA error state is
 * cleared), vt_copy_to_host_async is the D2H copy on `stream`. */
int vt_host_register(void *ptr, size_t n_bytes);
int vt_host_unregister(void *ptr);
/* Source side: pin a (private) mapping of the bitstream file so that the H2D copy reads the page cache directly;
 * vt_copy_to_device_async is that copy.  VT_ERR_CUDA when the mapping cannot be pinned (then the caller stages). */
int vt_host_register_source(void *ptr, size_t n_bytes);
int vt_copy_to_device_async(void *dst_dev, const void *src_host, size_t n_bytes, void *stream);
int vt_copy_to_host_async(void *dst_host, const void *src_dev, size_t n_bytes, void *stream);

/* ---- K0: decode front end --------------------------------------------------------------------------------
 * Replaces: libavcodec inside the ffmpeg child process (src/utils/video_segmenter.py:141-154,
 * src/analyzer/content_analyzer.py:193-211) and ffprobe (src/utils/video_utils.py:7-38). */
typedef struct vt_stream_info {
    int codec;          /* 4 = H.264 (cudaVideoCodec numbering) */
    int width, height;  /* display size after cropping */
    int coded_width, coded_height;
    int fps_num, fps_den; /* from VUI timing; 0/0 when absent */
    int n_frames;       /* access units found */
    int n_idr;          /* of which IDR */
    int pcm_intra_only; /* 1 when every slice is I_PCM-only or all-P_Skip (decodable by vt_h264_pcm_decode) */
} vt_stream_info;

/* Scan an Annex-B H.264 elementary stream held in host memory. frame_offsets/frame_sizes/frame_flags may be
 * NULL; otherwise they receive up to max_frames entries (flags: bit0 = IDR, bit1 = all-skip P picture). */
int vt_h264_scan(const uint8_t *bitstream, size_t n_bytes, vt_stream_info *info, uint64_t *frame_offsets,
                 uint32_t *frame_sizes, uint32_t *frame_flags, int max_frames);

/* Decode `n_frames` access units of a PCM-intra stream (I_PCM IDR pictures + all-P_Skip pictures) that
 * already sit in device memory, into NV12 surfaces (pitch-linear, UV plane at pitch*height).
 *   bitstream_dev : the elementary stream bytes in device memory
 *   payload_off   : n_frames x u64 (host), byte offset of each picture's first macroblock (from vt_h264_scan
 *                   via vt_h264_pcm_layout), or UINT64_MAX for an all-skip picture
 *   prev_dev      : surface that precedes frame 0 in decode order (needed when frame 0 is a skip picture)
 * Any other H.264 feature is VT_ERR_UNSUPPORTED: on this pool the driver refuses NVDEC. */
int vt_h264_pcm_layout(const uint8_t *bitstream, size_t n_bytes, const uint64_t *frame_offsets,
                       const uint32_t *frame_sizes, int n_frames, uint64_t *payload_off);
/* Same, with the parameter sets given explicitly (MP4: they live in avcC, samples are length-prefixed NALs;
 * frame_offsets then point at each sample's NAL header byte inside the file image `bitstream`). */
int vt_h264_pcm_layout_ps(const uint8_t *bitstream, size_t n_bytes, const uint8_t *sps_nal, size_t sps_len,
                          const uint8_t *pps_nal, size_t pps_len, const uint64_t *frame_offsets,
                          const uint32_t *frame_sizes, int n_frames, uint64_t *payload_off);
/* bitstream_dev must be readable 32 bytes past the last picture (whole 16-byte groups are fetched). */
int vt_h264_pcm_decode(const uint8_t *bitstream_dev, const uint64_t *payload_off, int n_frames, int width,
                       int height, const uint8_t *prev_dev, uint8_t *nv12_dev, int pitch, size_t frame_stride,
                       void *stream);

/* ---- the fused production entry: one call per batch of pictures -------------------------------------------
 * decode (K0) -> SAD + histogram on the decoded luma (K3) -> frames at the output size (K2 when `plan` is given,
 * else K1 same-size conversion), all enqueued on `stream`.  This is the whole per-batch body of the ingest pass
 * (src/analyzer/content_analyzer.py:193-211 is the reference's decode -> scale loop); hosts in any language drive the
 * path with this one call plus their own copies.
 *   prev_dev  : the surface that precedes picture 0 of the batch (NULL for the first batch of a stream)
 *   out_dev   : n_frames x out_frame_bytes planar YUV420P, or NULL to skip the frame output (scores only) */
int vt_ingest_batch_pcm(const vt_scale_plan *plan, const uint8_t *bitstream_dev, const uint64_t *payload_off,
                        int n_frames, int width, int height, const uint8_t *prev_dev, uint8_t *nv12_dev, int pitch,
                        size_t surface_bytes, uint64_t *sad_dev, uint32_t *hist_dev, uint8_t *out_dev,
                        size_t out_frame_bytes, void *stream);

/* NVDEC availability: VT_OK when libnvcuvid.so.1 loads and cuvidGetDecoderCaps reports H.264 8-bit 4:2:0
 * support; VT_ERR_NVDEC otherwise (message in vt_last_error()). */
int vt_nvdec_probe(int *n_engines, int *max_w, int *max_h);

/* ---- K0 proper: NVDEC sessions (libnvcuvid parser + decoder behind dlopen) --------------------------------------
 * Replaces libavcodec inside the ffmpeg child processes (src/utils/video_segmenter.py:141-154,
 * src/analyzer/content_analyzer.py:193-211) for real H.264 / HEVC / VP9 / AV1 content.  codec uses cudaVideoCodec
 * numbering (4 H.264, 8 HEVC, 9 VP9, 11 AV1).  Feed elementary-stream bytes (Annex-B for H.264/HEVC) with their
 * timestamps; next_surface returns VT_OK with a mapped pitch-linear NV12 surface in device memory (chroma plane at
 * surface + pitch * surface_rows) -- the input layout of every kernel above -- or 1 when no picture is displayable
 * yet; release the surface when the kernels reading it have been enqueued on `stream` and completed.
 * vt_decode_open returns VT_ERR_NVDEC (vt_nvdec_probe's message) where the driver exposes no video decode. */
typedef struct vt_decoder vt_decoder;
int vt_decode_open(int codec, int max_surfaces, void *stream, vt_decoder **out);
int vt_decode_feed(vt_decoder *dec, const uint8_t *data, size_t n_bytes, int64_t pts, int end_of_stream);
int vt_decode_next_surface(vt_decoder *dec, uint64_t *surface_dev, int *pitch, int *width, int *height,
                           int *surface_rows, int64_t *pts);
int vt_decode_release_surface(vt_decoder *dec, uint64_t surface_dev);
void vt_decode_close(vt_decoder *dec);

#ifdef __cplusplus
}
#endif
#endif /* VTSEG_H */
