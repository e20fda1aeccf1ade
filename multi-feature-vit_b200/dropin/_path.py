"""Makes `mfvit` importable when this directory is put on sys.path in place of the reference's moco_pretraining/moco/,
or when its files are laid over that directory (tools/overlay.py leaves the package location in a text file)."""
import os
import sys

try:
    import mfvit  # noqa: F401  already importable
except ImportError:
    _here = os.path.dirname(os.path.abspath(__file__))
    _note = os.path.join(_here, "_mfvit_location.txt")
    _pkg = open(_note).read().strip() if os.path.exists(_note) else os.path.dirname(_here)
    if _pkg not in sys.path:
        sys.path.insert(0, _pkg)
