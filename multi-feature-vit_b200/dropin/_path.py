"""Makes `mfvit` importable when this directory is put on sys.path in place of the reference's moco_pretraining/moco/."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
