"""Drop-in for the reference's absent `vits_returnftrs.py` (MAIN_CA:44 `import vits_returnftrs as vits`): the author's
variant of vits.py whose models also expose `features3D(x) -> [B, 197, 384]` (FUS:80,83,128,133).  Our vits.py already
provides features3D, so this module re-exports it under the name the script looks up (`vits.__dict__[args.arch]`)."""
import _path  # noqa: F401
from vits import (VisionTransformerMoCo, vit_base, vit_base_ori, vit_conv_base, vit_conv_small, vit_small,  # noqa: F401
                  vit_small_ori)

__all__ = ["vit_small", "vit_base", "vit_small_ori", "vit_base_ori", "vit_conv_small", "vit_conv_base"]
