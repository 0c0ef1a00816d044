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
