"""Drop-in for moco/builder_vit_mocov3structure_mocov2loss_noprediction_q.py: the same model as
builder_vit_mocov3structure_mocov2loss except that the keys are the momentum encoder's projections, not run through the
predictor (the reference files differ in exactly that line, BLD_NOPRED:174-175)."""
from .builder_vit_mocov3structure_mocov2loss import MoCo as _MoCo
from .builder_vit_mocov3structure_mocov2loss import MoCo_ResNet  # noqa: F401
from .builder_vit_mocov3structure_mocov2loss import MoCo_ViT as _MoCo_ViT
from .builder_vit_mocov3structure_mocov2loss import concat_all_gather  # noqa: F401


class MoCo(_MoCo):
    predictor_on_keys = False


class MoCo_ViT(_MoCo_ViT):
    predictor_on_keys = False
