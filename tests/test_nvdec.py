"""K0 proper: the NVDEC session behind the capability probe.

On the pool this repo is developed on the driver refuses video decode (cuvidGetDecoderCaps returns
CUDA_ERROR_NO_DEVICE behind the paravirtual proxy), so the decode part SKIPS with the probe's own message and only
the refusal contract is checked; on a machine where vt_nvdec_probe() == VT_OK the same test decodes the synthetic
PCM-intra clip through NVDEC and compares every surface with the generator's pictures and with the CUDA PCM decoder."""
import ctypes

import numpy as np
import pytest

from video_transformer_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def test_nvdec_session_decodes_or_refuses_loudly(cuda):
    import torch
    from video_transformer_b200 import decode
    L = _lib.lib()
    n, mw, mh = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = L.vt_nvdec_probe(ctypes.byref(n), ctypes.byref(mw), ctypes.byref(mh))
    if rc != _lib.VT_OK:
        msg = (L.vt_last_error() or b"").decode()
        with pytest.raises(_lib.VtError) as exc:
            decode.NvdecSession(decode.NvdecSession.H264)
        assert exc.value.code == _lib.VT_ERR_NVDEC                      # refused, never a CPU fallback
        pytest.skip("NVDEC not reachable here: " + msg)
    w, h, frames = 640, 480, 12
    bs, meta = synth.make_testsrc_h264(w, h, frames, fps=30, gop=4, cuts=[6])
    ses = decode.NvdecSession(decode.NvdecSession.H264, max_surfaces=8, stream=torch.cuda.current_stream())
    ses.feed(bs, 0, end_of_stream=True)
    ref = decode.H264PcmDecoder(bs, device="cuda:0").decode(0, frames, pitch=640).cpu().numpy()
    k = 0
    for ptr, pitch, sw, sh, rows, _pts in ses.surfaces():
        assert (sw, sh) == (w, h)
        host = np.empty((rows + rows // 2) * pitch, np.uint8)
        _lib.check(L.vt_copy_to_host_async(host.ctypes.data, ctypes.c_void_p(ptr), host.size, None))
        torch.cuda.synchronize()
        y = host[: rows * pitch].reshape(rows, pitch)[:h, :w]
        uv = host[rows * pitch:].reshape(rows // 2, pitch)[: h // 2, :w]
        assert np.array_equal(y, ref[k, :h, :w]), k
        assert np.array_equal(uv, ref[k, h:h + h // 2, :w]), k
        k += 1
    assert k == frames
    ses.close()
