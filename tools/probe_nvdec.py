"""Second probe: why does cuvidGetDecoderCaps return 100 on the box?  Not product code."""
import ctypes, json, os, subprocess, sys

out = {}
def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %r" % (e,)

out["LD_LIBRARY_PATH"] = os.environ.get("LD_LIBRARY_PATH")
out["dev"] = sh("ls -la /dev/nvidia* /dev/nvidia-caps 2>&1")
out["libs"] = sh("ls -la /usr/lib/libnvcuvid* /usr/local/nvidia/lib/ 2>&1 | head -80")
out["usr_lib_cuda"] = sh("ls -la /usr/lib/libcuda* /usr/lib/x86_64-linux-gnu/libcuda* /usr/local/cuda/compat/ 2>&1")
out["proc_caps"] = sh("ls /proc/driver/nvidia/ 2>&1; cat /proc/driver/nvidia/version 2>&1; ls /proc/driver/nvidia/capabilities 2>&1; cat /proc/driver/nvidia/params 2>&1 | head -40")

class CAPS(ctypes.Structure):
    _fields_ = [("eCodecType", ctypes.c_int), ("eChromaFormat", ctypes.c_int), ("nBitDepthMinus8", ctypes.c_uint),
                ("reserved1", ctypes.c_uint * 3), ("bIsSupported", ctypes.c_ubyte), ("nNumNVDECs", ctypes.c_ubyte),
                ("nOutputFormatMask", ctypes.c_ushort), ("nMaxWidth", ctypes.c_uint), ("nMaxHeight", ctypes.c_uint),
                ("nMaxMBCount", ctypes.c_uint), ("nMinWidth", ctypes.c_ushort), ("nMinHeight", ctypes.c_ushort),
                ("bIsHistogramSupported", ctypes.c_ubyte), ("nCounterBitDepth", ctypes.c_ubyte),
                ("nMaxHistogramBins", ctypes.c_ushort), ("reserved3", ctypes.c_uint * 10)]

def try_caps(lib, tag):
    c = CAPS(); c.eCodecType = 4; c.eChromaFormat = 1
    rc = lib.cuvidGetDecoderCaps(ctypes.byref(c))
    out["caps_" + tag] = dict(rc=rc, sup=c.bIsSupported, n=c.nNumNVDECs, mw=c.nMaxWidth, mh=c.nMaxHeight)

mode = sys.argv[1] if len(sys.argv) > 1 else "driver"
cu = ctypes.CDLL("libcuda.so.1")
if mode == "driver":
    out["cuInit"] = cu.cuInit(0)
    dev = ctypes.c_int(); out["cuDeviceGet"] = cu.cuDeviceGet(ctypes.byref(dev), 0)
    ctx = ctypes.c_void_p()
    out["cuCtxCreate"] = cu.cuCtxCreate_v2(ctypes.byref(ctx), 0, dev)
else:
    import torch
    torch.zeros(1, device="cuda"); torch.cuda.synchronize()
cur = ctypes.c_void_p(); out["cuCtxGetCurrent"] = (cu.cuCtxGetCurrent(ctypes.byref(cur)), cur.value)
ver = ctypes.c_int(); cu.cuDriverGetVersion(ctypes.byref(ver)); out["drv_ver"] = ver.value

for path in ("libnvcuvid.so.1", "/usr/local/nvidia/lib/libnvcuvid.so.1", "/usr/lib/libnvcuvid.so.1"):
    try:
        lib = ctypes.CDLL(path)
        try_caps(lib, path)
    except OSError as e:
        out["load_" + path] = str(e)
out["maps"] = sh("grep -E 'cuvid|libcuda|nvidia' /proc/%d/maps | awk '{print $6}' | sort -u" % os.getpid())

# try a real decoder create
class RECT(ctypes.Structure):
    _fields_ = [("left", ctypes.c_short), ("top", ctypes.c_short), ("right", ctypes.c_short), ("bottom", ctypes.c_short)]
class CREATEINFO(ctypes.Structure):
    _fields_ = [("ulWidth", ctypes.c_ulong), ("ulHeight", ctypes.c_ulong), ("ulNumDecodeSurfaces", ctypes.c_ulong),
                ("CodecType", ctypes.c_int), ("ChromaFormat", ctypes.c_int), ("ulCreationFlags", ctypes.c_ulong),
                ("bitDepthMinus8", ctypes.c_ulong), ("ulIntraDecodeOnly", ctypes.c_ulong), ("ulMaxWidth", ctypes.c_ulong),
                ("ulMaxHeight", ctypes.c_ulong), ("Reserved1", ctypes.c_ulong), ("display_area", RECT),
                ("OutputFormat", ctypes.c_int), ("DeinterlaceMode", ctypes.c_int), ("ulTargetWidth", ctypes.c_ulong),
                ("ulTargetHeight", ctypes.c_ulong), ("ulNumOutputSurfaces", ctypes.c_ulong), ("vidLock", ctypes.c_void_p),
                ("target_rect", RECT), ("enableHistogram", ctypes.c_ulong), ("Reserved2", ctypes.c_ulong * 4)]
try:
    lib = ctypes.CDLL("libnvcuvid.so.1")
    ci = CREATEINFO()
    ci.ulWidth = 1280; ci.ulHeight = 720; ci.ulNumDecodeSurfaces = 8; ci.CodecType = 4; ci.ChromaFormat = 1
    ci.ulCreationFlags = 4  # PreferCUVID
    ci.ulMaxWidth = 1280; ci.ulMaxHeight = 720
    ci.display_area = RECT(0, 0, 1280, 720); ci.OutputFormat = 0; ci.DeinterlaceMode = 0
    ci.ulTargetWidth = 1280; ci.ulTargetHeight = 720; ci.ulNumOutputSurfaces = 2
    dec = ctypes.c_void_p()
    out["sizeof_createinfo"] = ctypes.sizeof(ci)
    out["cuvidCreateDecoder"] = lib.cuvidCreateDecoder(ctypes.byref(dec), ctypes.byref(ci))
    if dec.value:
        out["cuvidDestroyDecoder"] = lib.cuvidDestroyDecoder(dec)
except Exception as e:  # noqa: BLE001
    out["create_exc"] = repr(e)

# nvjpeg hardware backend as an alternative hardware decode engine
try:
    nj = ctypes.CDLL("libnvjpeg.so.12")
    h = ctypes.c_void_p()
    NVJPEG_BACKEND_HARDWARE = 3
    rc = nj.nvjpegCreateEx(NVJPEG_BACKEND_HARDWARE, None, None, 0, ctypes.byref(h))
    out["nvjpeg_hw_backend_rc"] = rc
except Exception as e:  # noqa: BLE001
    out["nvjpeg_exc"] = repr(e)

print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_nvdec_%s.json" % mode, "w"), indent=1)
