"""The colour loss and PSNR the reference computes on the decoder's output -- src/loss/loss_mse.py:21-44 (``LossMse``),
src/evaluation/metrics.py:11-19 (``compute_psnr``) -- with the same names, constructor and call signatures, plus the
fused path of SURVEY.md 8f rank 3: when the decoder was given the target (``DecoderSplattingCUDA.forward_fused_mse(...,
mse_target=...)``) the loss value, its dL/dcolor and the PSNR's squared error were produced by the compositing epilogue
while every pixel was still in registers; ``LossMse.forward`` then just returns that scalar (its backward hands the
stored dL/dcolor to the rasterizer's backward without another pass over the images) and ``compute_psnr`` the per-view
value.  Anything the epilogue does not cover (``clamp_large_error``, a mixed ``valid_depth_mask``, another target or
weight) takes the reference's own tensor expressions, restated below.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Optional

import torch
from torch import Tensor, nn


@dataclass
class LossMseCfg:
    weight: float


@dataclass
class LossMseCfgWrapper:
    mse: LossMseCfg


@dataclass
class FusedMse:
    """What the compositing epilogue computed against ``target``."""
    loss: Tensor          # scalar, differentiable: weight * mean((color - target)^2)  (l1: mean |color - target|)
    sse_clipped: Tensor   # [batch, view] sum over (3, H, W) of (clip(color) - clip(target))^2
    target: Tensor
    weight: float
    l1: bool

    def matches(self, target: Tensor, weight: float, l1: bool) -> bool:
        return (target is self.target or (target.data_ptr() == self.target.data_ptr() and target.shape == self.target.shape)) \
            and float(weight) == self.weight and bool(l1) == self.l1


class LossMse(nn.Module):
    """Same constructor (a wrapper dataclass whose single field holds the cfg, src/loss/loss.py:20-26) and ``forward``
    as the reference's."""

    def __init__(self, cfg: LossMseCfgWrapper) -> None:
        super().__init__()
        (field,) = fields(type(cfg))
        self.cfg = getattr(cfg, field.name)
        self.name = field.name

    def forward(self, prediction, batch, gaussians, global_step: int, l1_loss: bool = False, clamp_large_error: float = 0.0,
                valid_depth_mask: Optional[Tensor] = None) -> Tensor:
        target = batch["target"]["image"]
        fused = getattr(prediction, "fused_mse", None)
        # the mask only filters when it is mixed (loss_mse.py:35); deciding that needs its values, so any mask takes the unfused path
        if fused is not None and clamp_large_error <= 0 and valid_depth_mask is None and fused.matches(target, self.cfg.weight, l1_loss):
            return fused.loss
        delta = prediction.color - target
        if valid_depth_mask is not None and valid_depth_mask.max() > 0.5 and valid_depth_mask.min() < 0.5:
            delta = delta[~valid_depth_mask]
        if clamp_large_error > 0:
            valid_mask = (delta ** 2) < clamp_large_error
            delta = delta[valid_mask]
        if l1_loss:
            return self.cfg.weight * (delta.abs()).mean()
        return self.cfg.weight * (delta ** 2).mean()


@torch.no_grad()
def compute_psnr(ground_truth: Tensor, predicted: Tensor) -> Tensor:
    """[batch, channel, height, width] x 2 -> [batch] (metrics.py:11-19)."""
    ground_truth = ground_truth.clip(min=0, max=1)
    predicted = predicted.clip(min=0, max=1)
    mse = ((ground_truth - predicted) ** 2).mean(dim=(1, 2, 3))
    return -10 * mse.log10()


@torch.no_grad()
def fused_psnr(prediction) -> Tensor:
    """PSNR of every (batch, view) frame from the squared error the epilogue accumulated: [batch, view]."""
    f = prediction.fused_mse
    n = f.target.shape[-3] * f.target.shape[-2] * f.target.shape[-1]
    return -10 * (f.sse_clipped / n).log10()
