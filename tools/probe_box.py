"""One-shot probe of the GPU box: is NVDEC (libnvcuvid) reachable, which host cores, which libs.

Run under gpurun; writes gpurun_out/probe_box.json. Not part of the product path.
"""
import ctypes, json, os, subprocess, sys, glob

out = {}

def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %r" % (e,)

out["nvidia_smi_L"] = sh("nvidia-smi -L")
out["gpu_query"] = sh("nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv")
out["dec_q"] = sh("nvidia-smi -q | grep -iE 'decoder|encoder|jpeg|ofa' | head -20")
out["nproc"] = os.cpu_count()
out["affinity"] = len(os.sched_getaffinity(0))
out["mem"] = sh("free -g | head -2")
out["caps_env"] = os.environ.get("NVIDIA_DRIVER_CAPABILITIES")
out["ldconfig"] = sh("ldconfig -p | grep -iE 'nvcuvid|nvidia-encode|libcuda|nvjpeg|nvidia-ml|opticalflow'")
out["find_nvcuvid"] = sh("find / -xdev \\( -name 'libnvcuvid*' -o -name 'libnvidia-encode*' -o -name 'libnvidia-opticalflow*' \\) 2>/dev/null | head")
out["find_nvcuvid_all"] = sh("find / \\( -name 'libnvcuvid*' \\) 2>/dev/null | head")
out["shm"] = sh("df -h /dev/shm /tmp | cat")

try:
    import cv2
    out["cv2"] = cv2.__version__
    out["swscale_libs"] = glob.glob(os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs", "libswscale*"))
except Exception as e:  # noqa: BLE001
    out["cv2"] = "ERR %r" % (e,)

import torch
out["torch_cuda"] = torch.cuda.is_available()
if torch.cuda.is_available():
    torch.zeros(1, device="cuda")
    out["props"] = str(torch.cuda.get_device_properties(0))

class CUVIDDECODECAPS(ctypes.Structure):
    _fields_ = [
        ("eCodecType", ctypes.c_int), ("eChromaFormat", ctypes.c_int), ("nBitDepthMinus8", ctypes.c_uint),
        ("reserved1", ctypes.c_uint * 3),
        ("bIsSupported", ctypes.c_ubyte), ("nNumNVDECs", ctypes.c_ubyte), ("nOutputFormatMask", ctypes.c_ushort),
        ("nMaxWidth", ctypes.c_uint), ("nMaxHeight", ctypes.c_uint), ("nMaxMBCount", ctypes.c_uint),
        ("nMinWidth", ctypes.c_ushort), ("nMinHeight", ctypes.c_ushort),
        ("bIsHistogramSupported", ctypes.c_ubyte), ("nCounterBitDepth", ctypes.c_ubyte), ("nMaxHistogramBins", ctypes.c_ushort),
        ("reserved3", ctypes.c_uint * 10),
    ]

caps = {}
try:
    lib = None
    for name in ("libnvcuvid.so.1", "libnvcuvid.so"):
        try:
            lib = ctypes.CDLL(name)
            out["nvcuvid_loaded"] = name
            break
        except OSError as e:
            out["nvcuvid_err_" + name] = str(e)
    if lib is not None:
        for cname, codec in (("h264", 4), ("hevc", 8), ("vp9", 10), ("av1", 11), ("mpeg4", 2), ("jpeg", 5)):
            c = CUVIDDECODECAPS()
            c.eCodecType = codec; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0
            rc = lib.cuvidGetDecoderCaps(ctypes.byref(c))
            caps[cname] = dict(rc=rc, supported=c.bIsSupported, n_nvdec=c.nNumNVDECs, fmt_mask=c.nOutputFormatMask,
                               max_w=c.nMaxWidth, max_h=c.nMaxHeight, max_mb=c.nMaxMBCount, min_w=c.nMinWidth,
                               min_h=c.nMinHeight, hist=c.bIsHistogramSupported)
except Exception as e:  # noqa: BLE001
    out["nvcuvid_exc"] = repr(e)
out["decoder_caps"] = caps

try:
    enc = ctypes.CDLL("libnvidia-encode.so.1")
    out["nvenc_loaded"] = True
except OSError as e:
    out["nvenc_loaded"] = str(e)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe_box.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
