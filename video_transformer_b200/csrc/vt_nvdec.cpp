// vt_nvdec.cpp -- NVDEC reachability probe (K0).
//
// The north star feeds the kernels from NVDEC.  libnvcuvid ships no headers in this image, so the two
// structures needed for the capability query are declared here from the Video Codec SDK ABI.  On this pool
// the driver sits behind a paravirtual proxy with NVIDIA_DRIVER_CAPABILITIES=compute,utility:
// libnvcuvid.so.1 loads, but cuvidGetDecoderCaps and cuvidCreateDecoder both return CUDA_ERROR_NO_DEVICE (100)
// (tools/probe_nvdec.py, profiles/r01_box_probe.md).  vt_nvdec_probe reports that state; the decode entry
// points then refuse streams outside the PCM-intra subset instead of decoding on the CPU.
#include <dlfcn.h>
#include <string.h>

#include "../../include/vtseg.h"

namespace vt { void set_error(const char *fmt, ...); }

namespace {
struct CuvidDecodeCaps {
    int eCodecType;
    int eChromaFormat;
    unsigned int nBitDepthMinus8;
    unsigned int reserved1[3];
    unsigned char bIsSupported;
    unsigned char nNumNVDECs;
    unsigned short nOutputFormatMask;
    unsigned int nMaxWidth;
    unsigned int nMaxHeight;
    unsigned int nMaxMBCount;
    unsigned short nMinWidth;
    unsigned short nMinHeight;
    unsigned char bIsHistogramSupported;
    unsigned char nCounterBitDepth;
    unsigned short nMaxHistogramBins;
    unsigned int reserved3[10];
};
typedef int (*cuvidGetDecoderCaps_t)(CuvidDecodeCaps *);
}  // namespace

extern "C" int vt_nvdec_probe(int *n_engines, int *max_w, int *max_h) {
    if (n_engines) *n_engines = 0;
    if (max_w) *max_w = 0;
    if (max_h) *max_h = 0;
    void *h = dlopen("libnvcuvid.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        vt::set_error("vt_nvdec_probe: libnvcuvid.so.1 not loadable: %s", dlerror());
        return VT_ERR_NVDEC;
    }
    cuvidGetDecoderCaps_t caps_fn = (cuvidGetDecoderCaps_t)dlsym(h, "cuvidGetDecoderCaps");
    if (!caps_fn) {
        vt::set_error("vt_nvdec_probe: cuvidGetDecoderCaps missing");
        dlclose(h);
        return VT_ERR_NVDEC;
    }
    CuvidDecodeCaps c;
    memset(&c, 0, sizeof(c));
    c.eCodecType = 4;     // cudaVideoCodec_H264
    c.eChromaFormat = 1;  // 4:2:0
    const int rc = caps_fn(&c);
    if (rc != 0 || !c.bIsSupported) {
        vt::set_error("vt_nvdec_probe: cuvidGetDecoderCaps rc=%d supported=%d (driver exposes no video decode here)",
                      rc, (int)c.bIsSupported);
        return VT_ERR_NVDEC;
    }
    if (n_engines) *n_engines = c.nNumNVDECs;
    if (max_w) *max_w = (int)c.nMaxWidth;
    if (max_h) *max_h = (int)c.nMaxHeight;
    return VT_OK;
}

// ---- decode sessions (K0 proper) -----------------------------------------------------------------------------------
// vt_decode_open / feed / next_surface / release_surface / close: a thin session over libnvcuvid's parser + decoder
// (SURVEY.md section 8b's proposed surface).  The parser splits the elementary stream, calls back with the sequence
// header (decoder creation), with every picture's parameters (cuvidDecodePicture) and with pictures in display order;
// next_surface maps the next displayable picture as a pitch-linear NV12 surface in device memory -- the layout every
// kernel of this library consumes -- on the caller's stream.  Nothing is copied and nothing runs on the CPU besides the
// bitstream parser inside the driver library.
//
// The Video Codec SDK ships no headers in this image; the structures below are declared from its ABI (nvcuvid.h /
// cuviddec.h, SDK 11-12 layouts).  CUVIDPICPARAMS is passed through opaquely.  On this pool the driver refuses video
// decode (vt_nvdec_probe), so this code is compiled and exported but has never run here: vt_decode_open returns
// VT_ERR_NVDEC with the probe's message, and tests/test_nvdec.py skips with it.
#include <deque>
#include <mutex>
#include <new>

namespace {

struct CuvidEofFormat {                 // CUVIDEOFORMAT
    int codec;
    struct { unsigned int numerator, denominator; } frame_rate;
    unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8, min_num_decode_surfaces;
    unsigned int coded_width, coded_height;
    struct { int left, top, right, bottom; } display_area;
    int chroma_format;
    unsigned int bitrate;
    struct { int x, y; } display_aspect_ratio;
    struct { unsigned char flags, color_primaries, transfer_characteristics, matrix_coefficients; } video_signal_description;
    unsigned int seqhdr_data_length;
};
struct CuvidParserDispInfo {            // CUVIDPARSERDISPINFO
    int picture_index, progressive_frame, top_field_first, repeat_first_field;
    long long timestamp;
};
struct CuvidSourceDataPacket {          // CUVIDSOURCEDATAPACKET
    unsigned long flags, payload_size;
    const unsigned char *payload;
    long long timestamp;
};
typedef int (*SeqCb)(void *, CuvidEofFormat *);
typedef int (*DecCb)(void *, void * /* CUVIDPICPARAMS* */);
typedef int (*DispCb)(void *, CuvidParserDispInfo *);
struct CuvidParserParams {              // CUVIDPARSERPARAMS
    int CodecType;
    unsigned int ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
    unsigned int flags;                 // bAnnexb : 1
    unsigned int uReserved1[4];
    void *pUserData;
    SeqCb pfnSequenceCallback;
    DecCb pfnDecodePicture;
    DispCb pfnDisplayPicture;
    void *pfnGetOperatingPoint, *pfnGetSEIMsg;
    void *pvReserved2[5];
    void *pExtVideoInfo;
};
struct CuvidRect { short left, top, right, bottom; };
struct CuvidDecodeCreateInfo {          // CUVIDDECODECREATEINFO
    unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
    int CodecType, ChromaFormat;
    unsigned long ulCreationFlags, bitDepthMinus8, ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, Reserved1;
    CuvidRect display_area;
    int OutputFormat, DeinterlaceMode;
    unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
    void *vidLock;
    CuvidRect target_rect;
    unsigned long enableHistogram;
    unsigned long Reserved2[4];
};
struct CuvidProcParams {                // CUVIDPROCPARAMS
    int progressive_frame, second_field, top_field_first, unpaired_field;
    unsigned int reserved_flags, reserved_zero;
    unsigned long long raw_input_dptr;
    unsigned int raw_input_pitch, raw_input_format;
    unsigned long long raw_output_dptr;
    unsigned int raw_output_pitch, Reserved1;
    void *output_stream;
    unsigned int Reserved[46];
    unsigned long long *histogram_dptr;
    void *Reserved2[1];
};

struct CuvidApi {
    void *lib = nullptr;
    int (*CreateVideoParser)(void **, CuvidParserParams *) = nullptr;
    int (*ParseVideoData)(void *, CuvidSourceDataPacket *) = nullptr;
    int (*DestroyVideoParser)(void *) = nullptr;
    int (*CreateDecoder)(void **, CuvidDecodeCreateInfo *) = nullptr;
    int (*DestroyDecoder)(void *) = nullptr;
    int (*DecodePicture)(void *, void *) = nullptr;
    int (*MapVideoFrame64)(void *, int, unsigned long long *, unsigned int *, CuvidProcParams *) = nullptr;
    int (*UnmapVideoFrame64)(void *, unsigned long long) = nullptr;
    bool load() {
        if (lib) return true;
        lib = dlopen("libnvcuvid.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib) return false;
#define VT_SYM(field, name) *(void **)(&field) = dlsym(lib, name)
        VT_SYM(CreateVideoParser, "cuvidCreateVideoParser");
        VT_SYM(ParseVideoData, "cuvidParseVideoData");
        VT_SYM(DestroyVideoParser, "cuvidDestroyVideoParser");
        VT_SYM(CreateDecoder, "cuvidCreateDecoder");
        VT_SYM(DestroyDecoder, "cuvidDestroyDecoder");
        VT_SYM(DecodePicture, "cuvidDecodePicture");
        VT_SYM(MapVideoFrame64, "cuvidMapVideoFrame64");
        VT_SYM(UnmapVideoFrame64, "cuvidUnmapVideoFrame64");
#undef VT_SYM
        return CreateVideoParser && ParseVideoData && DestroyVideoParser && CreateDecoder && DestroyDecoder &&
               DecodePicture && MapVideoFrame64 && UnmapVideoFrame64;
    }
};
CuvidApi g_api;

}  // namespace

struct vt_decoder {
    void *parser = nullptr, *decoder = nullptr, *stream = nullptr;
    int codec = 0, max_surfaces = 0;
    int width = 0, height = 0, surface_height = 0;     // display size; rows of the luma plane in a mapped surface
    int fps_num = 0, fps_den = 0;
    int error = 0;
    std::mutex mu;
    std::deque<CuvidParserDispInfo> ready;

    static int on_sequence(void *user, CuvidEofFormat *f) {
        vt_decoder *d = static_cast<vt_decoder *>(user);
        if (f->chroma_format != 1 || f->bit_depth_luma_minus8 != 0) {
            vt::set_error("vt_decode: only 8-bit 4:2:0 streams are handled (chroma_format %d, bit depth %d)",
                          f->chroma_format, 8 + f->bit_depth_luma_minus8);
            d->error = VT_ERR_UNSUPPORTED;
            return 0;
        }
        if (d->decoder) {
            g_api.DestroyDecoder(d->decoder);
            d->decoder = nullptr;
        }
        const int w = f->display_area.right - f->display_area.left, h = f->display_area.bottom - f->display_area.top;
        CuvidDecodeCreateInfo ci;
        memset(&ci, 0, sizeof(ci));
        ci.ulWidth = f->coded_width;
        ci.ulHeight = f->coded_height;
        ci.ulNumDecodeSurfaces = (unsigned long)((f->min_num_decode_surfaces > d->max_surfaces ? f->min_num_decode_surfaces
                                                                                              : d->max_surfaces));
        ci.CodecType = f->codec;
        ci.ChromaFormat = 1;
        ci.ulCreationFlags = 4;               // cudaVideoCreate_PreferCUVID
        ci.ulMaxWidth = f->coded_width;
        ci.ulMaxHeight = f->coded_height;
        ci.display_area.left = (short)f->display_area.left;
        ci.display_area.top = (short)f->display_area.top;
        ci.display_area.right = (short)f->display_area.right;
        ci.display_area.bottom = (short)f->display_area.bottom;
        ci.OutputFormat = 0;                  // cudaVideoSurfaceFormat_NV12
        ci.DeinterlaceMode = f->progressive_sequence ? 0 : 2;   // Weave / Adaptive
        ci.ulTargetWidth = (unsigned long)w;
        ci.ulTargetHeight = (unsigned long)h;
        ci.ulNumOutputSurfaces = 4;
        const int rc = g_api.CreateDecoder(&d->decoder, &ci);
        if (rc != 0) {
            vt::set_error("vt_decode: cuvidCreateDecoder failed (%d) for %ux%u codec %d", rc, f->coded_width,
                          f->coded_height, f->codec);
            d->error = VT_ERR_NVDEC;
            return 0;
        }
        d->width = w;
        d->height = h;
        d->surface_height = (h + 1) & ~1;
        d->fps_num = (int)f->frame_rate.numerator;
        d->fps_den = (int)f->frame_rate.denominator;
        return (int)ci.ulNumDecodeSurfaces;   // tells the parser how many surfaces it may keep in flight
    }
    static int on_decode(void *user, void *pic_params) {
        vt_decoder *d = static_cast<vt_decoder *>(user);
        if (!d->decoder) return 0;
        const int rc = g_api.DecodePicture(d->decoder, pic_params);
        if (rc != 0) {
            vt::set_error("vt_decode: cuvidDecodePicture failed (%d)", rc);
            d->error = VT_ERR_NVDEC;
            return 0;
        }
        return 1;
    }
    static int on_display(void *user, CuvidParserDispInfo *info) {
        vt_decoder *d = static_cast<vt_decoder *>(user);
        if (!info) return 1;                  // end-of-stream notification
        std::lock_guard<std::mutex> lock(d->mu);
        d->ready.push_back(*info);
        return 1;
    }
};

extern "C" int vt_decode_open(int codec, int max_surfaces, void *stream, vt_decoder **out) {
    if (!out || max_surfaces < 1 || max_surfaces > 64) {
        vt::set_error("vt_decode_open: bad arguments");
        return VT_ERR_INVALID;
    }
    int rc = vt_nvdec_probe(nullptr, nullptr, nullptr);
    if (rc != VT_OK) return rc;               // message already set: the driver exposes no video decode here
    if (!g_api.load()) {
        vt::set_error("vt_decode_open: libnvcuvid.so.1 lacks the parser/decoder entry points");
        return VT_ERR_NVDEC;
    }
    vt_decoder *d = new (std::nothrow) vt_decoder();
    if (!d) return VT_ERR_NOMEM;
    d->codec = codec;
    d->max_surfaces = max_surfaces;
    d->stream = stream;
    CuvidParserParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.CodecType = codec;                     // cudaVideoCodec: 4 = H.264, 8 = HEVC, 9 = VP9, 11 = AV1
    pp.ulMaxNumDecodeSurfaces = 1;            // the sequence callback returns the real number
    pp.ulClockRate = 0;                       // timestamps are passed through in the caller's units
    pp.ulMaxDisplayDelay = 0;                 // low latency: pictures are handed over as soon as they are displayable
    pp.pUserData = d;
    pp.pfnSequenceCallback = vt_decoder::on_sequence;
    pp.pfnDecodePicture = vt_decoder::on_decode;
    pp.pfnDisplayPicture = vt_decoder::on_display;
    rc = g_api.CreateVideoParser(&d->parser, &pp);
    if (rc != 0) {
        vt::set_error("vt_decode_open: cuvidCreateVideoParser failed (%d)", rc);
        delete d;
        return VT_ERR_NVDEC;
    }
    *out = d;
    return VT_OK;
}

extern "C" int vt_decode_feed(vt_decoder *d, const uint8_t *data, size_t n_bytes, int64_t pts, int end_of_stream) {
    if (!d || (!data && n_bytes)) {
        vt::set_error("vt_decode_feed: bad arguments");
        return VT_ERR_INVALID;
    }
    CuvidSourceDataPacket pkt;
    memset(&pkt, 0, sizeof(pkt));
    pkt.flags = 2ul /* CUVID_PKT_TIMESTAMP */ | (end_of_stream ? 1ul /* CUVID_PKT_ENDOFSTREAM */ : 0ul);
    pkt.payload_size = (unsigned long)n_bytes;
    pkt.payload = data;
    pkt.timestamp = pts;
    const int rc = g_api.ParseVideoData(d->parser, &pkt);
    if (d->error) return d->error;
    if (rc != 0) {
        vt::set_error("vt_decode_feed: cuvidParseVideoData failed (%d)", rc);
        return VT_ERR_BITSTREAM;
    }
    return VT_OK;
}

// VT_OK: *surface_dev is a pitch-linear NV12 surface (luma rows, then interleaved chroma at surface_dev +
// pitch * surface_rows) valid until vt_decode_release_surface.  1: no picture is displayable yet (feed more data).
extern "C" int vt_decode_next_surface(vt_decoder *d, uint64_t *surface_dev, int *pitch, int *width, int *height,
                                      int *surface_rows, int64_t *pts) {
    if (!d || !surface_dev || !pitch) {
        vt::set_error("vt_decode_next_surface: bad arguments");
        return VT_ERR_INVALID;
    }
    CuvidParserDispInfo info;
    {
        std::lock_guard<std::mutex> lock(d->mu);
        if (d->ready.empty()) return 1;
        info = d->ready.front();
        d->ready.pop_front();
    }
    CuvidProcParams vp;
    memset(&vp, 0, sizeof(vp));
    vp.progressive_frame = info.progressive_frame;
    vp.top_field_first = info.top_field_first;
    vp.unpaired_field = info.repeat_first_field < 0;
    vp.output_stream = d->stream;
    unsigned long long ptr = 0;
    unsigned int pt = 0;
    const int rc = g_api.MapVideoFrame64(d->decoder, info.picture_index, &ptr, &pt, &vp);
    if (rc != 0) {
        vt::set_error("vt_decode_next_surface: cuvidMapVideoFrame64 failed (%d)", rc);
        return VT_ERR_NVDEC;
    }
    *surface_dev = ptr;
    *pitch = (int)pt;
    if (width) *width = d->width;
    if (height) *height = d->height;
    if (surface_rows) *surface_rows = d->surface_height;
    if (pts) *pts = info.timestamp;
    return VT_OK;
}

extern "C" int vt_decode_release_surface(vt_decoder *d, uint64_t surface_dev) {
    if (!d || !d->decoder) return VT_ERR_INVALID;
    const int rc = g_api.UnmapVideoFrame64(d->decoder, surface_dev);
    if (rc != 0) {
        vt::set_error("vt_decode_release_surface: cuvidUnmapVideoFrame64 failed (%d)", rc);
        return VT_ERR_NVDEC;
    }
    return VT_OK;
}

extern "C" void vt_decode_close(vt_decoder *d) {
    if (!d) return;
    if (d->parser) g_api.DestroyVideoParser(d->parser);
    if (d->decoder) g_api.DestroyDecoder(d->decoder);
    delete d;
}
