"""Builds libvtseg.so (hand-written sm_100a CUDA + the C ABI in include/vtseg.h) in-tree with nvcc.

Each translation unit is compiled to an object under csrc/_obj/ (only when it or a header changed, several at a
time), then linked into video_transformer_b200/libvtseg.so.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libvtseg.so")
SOURCES = ["vt_api.cu", "vt_score.cu", "vt_convert.cu", "vt_rgb.cu", "vt_scale.cu", "vt_scale_pair.cu",
           "vt_swsfilter.cpp", "vt_h264.cu", "vt_nvdec.cpp", "vt_jpeg.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC,-O2,-Wall"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "vtseg.h")]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return _stale(OUT, srcs + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, os.path.splitext(s)[0] + ".o")
        jobs.append((src, obj, force or _stale(obj, [src] + hdrs)))

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return 0, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        results = list(ex.map(compile_one, jobs))
    log = "".join(out for _, out in results)
    if verbose or any(rc for rc, _ in results):
        sys.stderr.write(log)
    if any(rc for rc, _ in results):
        raise RuntimeError("nvcc failed building libvtseg.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", OUT] + \
          [obj for _, obj, _ in jobs] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed linking libvtseg.so")
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
