"""mvd_b200 — B200-native (sm_100a) implementation of MVD's multi-view denoising hot path.

Public names mirror the reference (pananananas/MVD): `ImageCrossAttentionProcessor`,
`get_attention_processor_for_module`, `CameraEncoder`, `ImageEncoder`, `MultiViewUNet`, `UNetOutput`,
`create_mvd_pipeline`, `MVDPipeline`, `ShiftSNRScheduler`.
"""
__version__ = "0.1.0"

from .attention import ImageCrossAttentionProcessor, get_attention_processor_for_module  # noqa: F401
from .camera_encoder import CameraEncoder  # noqa: F401
from .image_encoder import ImageEncoder  # noqa: F401
from .mvd_unet import MultiViewUNet, UNetOutput, create_mvd_pipeline  # noqa: F401
from .pipeline import MVDPipeline  # noqa: F401
from .scheduler import DDPMScheduler, ShiftSNRScheduler  # noqa: F401
from .unet import UNet2DConditionModel  # noqa: F401
