"""Offline stand-in for the google-genai SDK (not installed here; Gemini is stubbed offline per the north star).
Only what /root/reference/src/analyzer/content_analyzer.py touches at import and construction time."""
from . import types  # noqa: F401


class Client:  # the reference's tests patch this symbol
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
