"""The C-ABI boundary: libvtseg.so loads on a machine without a GPU and exports exactly what include/vtseg.h declares
(no compute calls here).  The Python binding table must cover the same set."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vtseg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(vtlib):
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(os.path.join(ROOT, "video_transformer_b200", "libvtseg.so"))
    for n in names:
        assert hasattr(raw, n), "libvtseg.so does not export %s" % n


def test_binding_table_covers_the_header():
    from video_transformer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_host_only_entry_points_work_without_a_gpu(vtlib):
    assert vtlib.vt_version() >= 100
    assert vtlib.vt_scale_width_for_height(1920, 1080, 720) == 1280
    assert vtlib.vt_sws_max_taps(1920, 1280, 4) >= 6
    assert isinstance(vtlib.vt_last_error(), (bytes, type(None)))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from video_transformer_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.lib()
    except _lib.VtError as e:
        assert "not built" in str(e)
    else:
        raise AssertionError("lib() must raise when libvtseg.so is missing (no CPU fallback)")
