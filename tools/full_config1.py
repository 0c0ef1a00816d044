"""BASELINE.json configs[1] at its FULL size, through the plugin's functions, once: a 2-hour 1920x1080@30 source
(216,000 pictures) -> probe -> budget plan (10 segments of 720 s) -> manifest -> extract_segment per segment
(faststart MP4 + 1280x720 yuv420p .frames + .json on /dev/shm), a consumer that removes each segment's artefacts after
reading its sidecar (the reference uploads a segment and moves on, content_analyzer.py:745-775).

The source is the bench clip (3840 pictures, I_PCM IDR / GOP 30 + P_Skip, scene cuts) repeated: its MP4 samples are
tiled into one 24 GB file by isobmff.write_plans (file -> file ranges, nothing held in memory).

    gpurun --timeout 900 -- python tools/full_config1.py            (needs ~60 GB of /dev/shm)

`--config5` runs BASELINE.json configs[4]'s product on the same source instead: one picture per second as 768x768
RGB24 (`output="rgb24", rgb_size=(768, 768), sample_every=30`), hour-long segments (the budget plan of a 10-hour
file), every picture still decoded and scored.
"""
import dataclasses
import json
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402
from video_transformer_b200 import (budget_planner, container, isobmff, landing, video_segmenter)  # noqa: E402
from video_transformer_b200.video_utils import probe_duration  # noqa: E402


def main():
    import torch
    hours = float(os.environ.get("VT_FULL_HOURS", "2"))
    n_total = int(hours * 3600 * bench.FPS)
    work = "/dev/shm/vt_full_%d" % os.getpid()
    os.makedirs(work, exist_ok=True)
    out = {"workload": "configs[1] full size: %.1f h 1920x1080@30 (%d pictures) -> 1280x720 yuv420p + SAD/hist + "
                       "segment MP4s, through probe_duration / plan_segments_with_budget / manifest / extract_segment"
                       % (hours, n_total)}
    try:
        t0 = time.perf_counter()
        raw, small = os.path.join(work, "clip.h264"), os.path.join(work, "clip.mp4")
        bench.make_clip(bench.CLIP_FRAMES, raw)
        container.annexb_to_mp4(raw, small)
        os.unlink(raw)
        movie = isobmff.read_movie(small)
        t = movie.video_track()
        reps = -(-n_total // t.n)
        tile = lambda a: np.tile(a, reps)[:n_total]  # noqa: E731
        deltas = tile(t.deltas)
        big_t = dataclasses.replace(t, sizes=tile(t.sizes), offsets=tile(t.offsets), deltas=deltas,
                                    dts=np.concatenate(([0], np.cumsum(deltas)[:-1])).astype(np.int64),
                                    cts_off=None if t.cts_off is None else tile(t.cts_off), sync=tile(t.sync),
                                    media_duration=int(deltas.sum()))
        big = os.path.join(work, "lecture_2h.mp4")
        isobmff.write_plans(big, [isobmff.plan_whole_track(big_t, movie.timescale)], movie.timescale, movie.ftyp, small)
        os.unlink(small)
        out["source_bytes"] = os.path.getsize(big)
        out["source_build_s"] = round(time.perf_counter() - t0, 2)

        config5 = "--config5" in sys.argv
        if config5:
            video_segmenter.configure(target_height=720, frame_buffers=True, output="rgb24", rgb_size=(768, 768),
                                      sample_every=30)
            out["workload"] = ("configs[4] product on %.1f h of 1920x1080@30 (%d pictures): every picture decoded and "
                               "scored, one per second converted to 768x768 RGB24, hour-long segments" % (hours, n_total))
        else:
            video_segmenter.configure(target_height=720, frame_buffers=True)
        temp = os.path.join(work, "temp")
        t_job = time.perf_counter()
        duration = probe_duration(big)
        t_probe = time.perf_counter() - t_job
        plan = budget_planner.plan_segments_with_budget(duration, {}, 0)
        if config5:                                  # a 10-hour file plans (3600, 0): use that segment length here
            plan = type(plan)(3600, 0, int(-(-duration // 3600)), plan.estimated_calls, plan.available_calls,
                              plan.hard_max_calls, plan.fits_budget)
        manifest = video_segmenter.load_or_create_manifest(video_id="lecture_2h", duration=duration,
                                                           segment_seconds=plan.segment_duration,
                                                           overlap_seconds=plan.overlap, temp_dir=temp)
        mpath = video_segmenter.get_manifest_path("lecture_2h", temp)
        segs = []
        pictures = 0
        cuts = 0
        for entry in manifest["segments"]:
            seg = entry["file_path"]
            ts = time.perf_counter()
            ok = video_segmenter.extract_segment(input_path=big, start=entry["start"], end=entry["end"],
                                                 output_path=seg, stream_copy=True)
            dt = time.perf_counter() - ts
            assert ok, entry
            side = json.loads(open(seg[:-4] + ".json").read())
            frames_bytes = os.path.getsize(seg[:-4] + ".frames")
            assert frames_bytes == side["frames"] * side["frame_bytes"]
            pictures += side["last_picture"] - side["first_picture"]
            cuts += len(side["cuts"])
            segs.append({"id": entry["id"], "seconds": round(dt, 3), "pictures": side["last_picture"] - side["first_picture"],
                         "frames_out": side["frames"],
                         "landing": side["landing"], "recycled": side["landing_recycled"],
                         "mp4_bytes": os.path.getsize(seg),
                         "ms": {k: round(v * 1e3, 1) for k, v in video_segmenter.LAST_TIMINGS.items()}})
            video_segmenter.update_segment_status(manifest, entry["id"], "completed")
            video_segmenter.save_manifest(mpath, manifest)
            for ext in (".frames", ".mp4"):                       # the consumer is done with this segment
                os.unlink(seg[:-4] + ext)
        torch.cuda.synchronize()
        job_s = time.perf_counter() - t_job
        out.update({"duration_s": duration, "plan": [plan.segment_duration, plan.overlap, plan.num_segments],
                    "probe_s": round(t_probe, 3), "job_s": round(job_s, 2), "pictures": pictures, "cuts": cuts,
                    "pictures_per_s": pictures / job_s,
                    "pictures_per_s_after_first_segment": sum(s["pictures"] for s in segs[1:]) /
                    max(1e-9, sum(s["seconds"] for s in segs[1:])),
                    "segments": segs, "arena": landing.stats()})
        print(json.dumps(out))
    finally:
        landing.release_all()
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
