#!/usr/bin/env python
"""Kernel micro-bench: each hot kernel launched back to back on synthetic device-resident 1080p NV12 surfaces.
usage: python tools/kbench.py [--frames 256] [--reps 10] [--src 1920x1080] [--dst-h 720] [--kernels scale,score]"""
import argparse, json, os, sys
from ctypes import c_void_p
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_transformer_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--src", default="1920x1080")
ap.add_argument("--dst-h", type=int, default=720)
ap.add_argument("--kernels", default="scale,score,convert")
a = ap.parse_args()
sw, sh = (int(v) for v in a.src.split("x"))
dh = a.dst_h
dw = ops.scale_width_for_height(sw, sh, dh)
L = _lib.lib()
dev = torch.device("cuda:0")
F = a.frames
pitch = (sw + 255) // 256 * 256
rows = sh + sh // 2
surf = torch.randint(0, 256, (F, rows, pitch), dtype=torch.uint8, device=dev)
fb = dw * dh * 3 // 2
out = torch.empty((F, fb), dtype=torch.uint8, device=dev)
sad = torch.empty(F, dtype=torch.int64, device=dev)
hist = torch.empty((F, 256), dtype=torch.int32, device=dev)
plan = ops.ScalePlan(sw, sh, dw, dh, ops.SWS_BICUBIC)
st = torch.cuda.current_stream(dev)
sp = c_void_p(st.cuda_stream)
peak = 6531.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

def run(name, fn, alg_bytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    gbs = alg_bytes * F / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "us_per_picture": round(1000 * ms / F, 4), "alg_gbs": round(gbs, 1),
                      "frac_of_measured_peak": round(gbs / peak, 4), "frames": F, "shape": "%dx%d->%dx%d" % (sw, sh, dw, dh)}))

ks = a.kernels.split(",")
if "scale" in ks:
    run("scale", lambda: _lib.check(L.vt_scale_nv12_to_yuv420p(plan._h, c_void_p(surf.data_ptr()), pitch, rows * pitch,
        c_void_p(out.data_ptr()), fb, F, sp)), sw * sh * 3 // 2 + fb)
    print(plan.stream_info(False), plan.stream_info(True))
if "overlap" in ks:
    # score and scale of the same surfaces on two streams (they do not depend on each other)
    s2 = torch.cuda.Stream(dev)
    sp2 = c_void_p(s2.cuda_stream)
    ev_a, ev_b = torch.cuda.Event(), torch.cuda.Event()
    def both():
        ev_a.record(st)
        s2.wait_event(ev_a)
        _lib.check(L.vt_scale_nv12_to_yuv420p(plan._h, c_void_p(surf.data_ptr()), pitch, rows * pitch,
                                              c_void_p(out.data_ptr()), fb, F, sp))
        _lib.check(L.vt_sad_hist_u8(c_void_p(surf.data_ptr()), pitch, rows * pitch, sw, sh, None, F,
                                    c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), sp2))
        ev_b.record(s2)
        st.wait_event(ev_b)
    run("scale||score", both, sw * sh * 3 // 2 + fb + 2 * sw * sh)
    def serial():
        _lib.check(L.vt_scale_nv12_to_yuv420p(plan._h, c_void_p(surf.data_ptr()), pitch, rows * pitch,
                                              c_void_p(out.data_ptr()), fb, F, sp))
        _lib.check(L.vt_sad_hist_u8(c_void_p(surf.data_ptr()), pitch, rows * pitch, sw, sh, None, F,
                                    c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), sp))
    run("scale;score", serial, sw * sh * 3 // 2 + fb + 2 * sw * sh)
if "fused" in ks:
    print(json.dumps({"fuses_score": plan.fuses_score}))
    run("scale+score (one call)", lambda: _lib.check(L.vt_scale_score_nv12_to_yuv420p(plan._h, c_void_p(surf.data_ptr()), pitch,
        rows * pitch, None, c_void_p(out.data_ptr()), fb, F, c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), sp)),
        sw * sh * 3 // 2 + fb + sw * sh)
if "score" in ks:
    run("score", lambda: _lib.check(L.vt_sad_hist_u8(c_void_p(surf.data_ptr()), pitch, rows * pitch, sw, sh, None, F,
        c_void_p(sad.data_ptr()), c_void_p(hist.data_ptr()), sp)), 2 * sw * sh)
if "convert" in ks:
    out2 = torch.empty((F, sw * sh * 3 // 2), dtype=torch.uint8, device=dev)
    run("nv12_to_yuv420p", lambda: _lib.check(L.vt_nv12_to_yuv420p(c_void_p(surf.data_ptr()), pitch, rows * pitch, sw, sh,
        c_void_p(out2.data_ptr()), sw * sh * 3 // 2, F, sp)), sw * sh * 3)
if "rgb" in ks:
    # config 5's product: NV12 -> 768x768 RGB24 (scaled) and NV12 -> RGB24 at the source size
    Fr = min(F, 64)
    rp = ops.RgbPlan(sw, sh, 768, 768, ops.SWS_BICUBIC)
    orgb = torch.empty((Fr, 768, 768, 3), dtype=torch.uint8, device=dev)
    F_save = F
    F = Fr
    run("nv12_to_rgb24_768", lambda: rp.scale_nv12(surf.view(-1), pitch, Fr, rows * pitch, out=orgb), sw * sh * 3 // 2 + 768 * 768 * 3)
    F = F_save
if "jpeg" in ks:
    # the reducer's Motion-JPEG encoder on 640x360 YUV420P pictures (smooth content + mild noise)
    Fj = min(F, 64)
    w2, h2 = 640, 360
    g = torch.Generator(device="cpu").manual_seed(5)
    yy = torch.arange(h2).view(-1, 1).float(); xx = torch.arange(w2).view(1, -1).float()
    base = (128 + 60 * torch.sin(xx / 37.0) * torch.cos(yy / 23.0)).clamp(16, 235)
    pics = []
    for k in range(Fj):
        y = (base + torch.randint(-6, 7, (h2, w2), generator=g)).clamp(16, 235).to(torch.uint8)
        u = (128 + 30 * torch.sin(xx[:, ::2] / 50.0 + k)).expand(h2 // 2, -1).clamp(16, 240).to(torch.uint8)
        v = (128 + 30 * torch.cos(yy[::2] / 40.0 + k)).expand(-1, w2 // 2).clamp(16, 240).to(torch.uint8)
        pics.append(torch.cat([y.reshape(-1), u.reshape(-1), v.reshape(-1)]))
    frames = torch.stack(pics).to(dev)
    jp = ops.JpegPlan(w2, h2, 75, True)
    outj = torch.empty(Fj * 200000, dtype=torch.uint8, device=dev)
    F_save = F
    F = Fj
    res = {}
    def enc():
        res["r"] = jp.encode(frames, out=outj)
    run("jpeg_640x360_q75", enc, w2 * h2 * 3 // 2)
    off = res["r"][1].cpu()
    print(json.dumps({"jpeg_bytes_per_picture": int(off[-1]) // Fj, "status": int(res["r"][2].cpu()[0])}))
    F = F_save
