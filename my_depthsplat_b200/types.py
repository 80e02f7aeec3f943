"""I/O types of the rendering hot path.

Mirrors src/model/types.py:7-12 (Gaussians), src/model/decoder/decoder.py:11-22
(DepthRenderingMode, DecoderOutput) of the reference: same field names, shapes and dtypes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

from torch import Tensor

DepthRenderingMode = Literal["depth", "log", "disparity", "relative_disparity"]


@dataclass
class Gaussians:
    means: Tensor        # [batch, gaussian, 3]
    covariances: Tensor  # [batch, gaussian, 3, 3]
    harmonics: Tensor    # [batch, gaussian, 3, d_sh]
    opacities: Tensor    # [batch, gaussian]


@dataclass
class DecoderOutput:
    color: Tensor            # [batch, view, 3, height, width]
    depth: Optional[Tensor]  # [batch, view, height, width] or None


@dataclass
class FusedDecoderOutput(DecoderOutput):
    """DecoderOutput plus what the compositing epilogue computed against a target image (loss_mse.FusedMse)."""
    fused_mse: Optional[object] = None
