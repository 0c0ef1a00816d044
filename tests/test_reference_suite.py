"""Runs the REFERENCE's own hot-path test files against this repo's drop-in modules (build container only: the
reference tree does not exist on the GPU box).  This is the drop-in proof for the call contract, manifest format and
timestamp plumbing that tests/test_long_video_*.py and tests/test_segment_analysis.py pin."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["test_video_segmenter.py", "test_budget_planner.py", "test_long_video_integration.py",
         "test_long_video_edge_cases.py", "test_segment_analysis.py"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference tree not present")
def test_reference_hot_path_tests_pass_against_the_drop_in(tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "refsuite_support"), os.path.dirname(HERE)])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "vt_redirect",
           "--rootdir", os.path.join(REF, "tests"), "-c", "/dev/null", "--basetemp", str(tmp_path / "bt")] + \
          [os.path.join(REF, "tests", f) for f in FILES]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    # 3 + 3 + 18 tests run; the ffmpeg-binary test skips itself (no ffmpeg in the image)
    import re
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 24, tail
