"""B200-native long-video ingest: drop-in for the reference's segmenter module functions.

Host side mirrors /root/reference/src/utils/video_segmenter.py, video_utils.py and budget_planner.py
(same names, arguments and error behaviour); pixels go through libvtseg.so (hand-written sm_100a CUDA
behind the C ABI in include/vtseg.h).  There is no CPU pixel path: without the library or a GPU the
pixel entry points raise / return False exactly as the reference does when ffmpeg is missing.
"""
__version__ = "0.1.0"
