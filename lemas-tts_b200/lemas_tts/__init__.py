"""B200-native drop-in for the `lemas_tts` package surface of LEMAS-TTS (reference lemas_tts/__init__.py:1-5).

`TTS` is resolved lazily so that importing helpers (synthetic inputs, host logic) does not load the CUDA library.
"""
__all__ = ["TTS"]
__version__ = "0.1.0"


def __getattr__(name):
    if name == "TTS":
        from .api import TTS

        return TTS
    raise AttributeError(name)
