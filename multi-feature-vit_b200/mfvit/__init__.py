"""mfvit - host-side runtime of the B200-native MF-ViT CA hot path (ctypes over libmfvit.so)."""
from . import _lib  # noqa: F401
from ._lib import MfvError  # noqa: F401
