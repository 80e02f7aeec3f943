"""Drop-in for the reference's decoder module (src/model/decoder/decoder_splatting_cuda.py:15-91,
decoder.py:19-48, __init__.py:5-13): same class names, constructor and ``forward`` signature, same
``DecoderOutput``; registry key ``"splatting_cuda"``.

The reference flattens (b v), makes V copies of every Gaussian tensor with einops.repeat
(:53-56) and renders depth with a second full rasterization (:69-91).  Here the Gaussians stay
``[B,N,...]`` and colour + depth of all B*V views come out of one rasterizer call.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Generic, Literal, Optional, TypeVar

import torch
from torch import Tensor, nn

from .cuda_splatting import DepthRenderingMode, render_views
from .types import DecoderOutput, FusedDecoderOutput, Gaussians


@dataclass
class DecoderSplattingCUDACfg:
    name: Literal["splatting_cuda"]


T = TypeVar("T")


class Decoder(nn.Module, ABC, Generic[T]):
    cfg: T
    dataset_cfg: Any  # the reference's DatasetCfg; only ``background_color`` is read on this path

    def __init__(self, cfg: T, dataset_cfg: Any) -> None:
        super().__init__()
        self.cfg = cfg
        self.dataset_cfg = dataset_cfg

    @abstractmethod
    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode: Optional[DepthRenderingMode] = None) -> DecoderOutput:
        ...


class DecoderSplattingCUDA(Decoder[DecoderSplattingCUDACfg]):
    background_color: Tensor  # [3]
    grad_reducer = None  # set by dist.ViewShardedDecoder(fused_reduce=True): cross-rank gradient sum inside the backward kernel

    def __init__(self, cfg: DecoderSplattingCUDACfg, dataset_cfg: Any) -> None:
        super().__init__(cfg, dataset_cfg)
        self.register_buffer("background_color", torch.tensor(dataset_cfg.background_color, dtype=torch.float32),
                             persistent=False)

    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode: Optional[DepthRenderingMode] = None) -> DecoderOutput:
        """gaussians [B,N,...]; extrinsics [B,V,4,4]; intrinsics [B,V,3,3]; near/far [B,V] ->
        DecoderOutput(color [B,V,3,H,W], depth [B,V,H,W] | None)."""
        color, depth = render_views(extrinsics, intrinsics, near, far, image_shape, self.background_color, gaussians.means,
                                    gaussians.covariances, gaussians.harmonics, gaussians.opacities, depth_mode=depth_mode,
                                    grad_reducer=self.grad_reducer)
        return DecoderOutput(color, depth)

    def forward_fused_mse(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                          image_shape: tuple[int, int], depth_mode: Optional[DepthRenderingMode] = None, *, mse_target: Tensor,
                          mse_weight: float = 1.0, mse_l1: bool = False, mse_count: Optional[int] = None) -> FusedDecoderOutput:
        """``forward`` plus loss-side fusion: ``mse_target`` [B,V,3,H,W] is the ground truth the
        reference's LossMse compares the colour with; the result also carries that loss (MSE, or L1) and the PSNR
        ingredients, computed in the compositing epilogue.  ``loss_mse.LossMse`` and ``loss_mse.fused_psnr`` pick them up."""
        from .loss_mse import FusedMse
        mse = dict(target=mse_target, weight=mse_weight, l1=mse_l1)
        if mse_count is not None:
            mse["count"] = mse_count
        color, depth = render_views(extrinsics, intrinsics, near, far, image_shape, self.background_color, gaussians.means,
                                    gaussians.covariances, gaussians.harmonics, gaussians.opacities, depth_mode=depth_mode,
                                    grad_reducer=self.grad_reducer, mse=mse)
        return FusedDecoderOutput(color, depth, FusedMse(mse["loss"], mse["sse_clipped"], mse_target, float(mse_weight), bool(mse_l1)))

    def render_depth(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                     image_shape: tuple[int, int], mode: DepthRenderingMode = "depth") -> Tensor:
        """-> [B,V,H,W]."""
        b, n = gaussians.opacities.shape
        dummy = torch.zeros((b, n, 3, 1), dtype=torch.float32, device=gaussians.means.device)
        _, depth = render_views(extrinsics, intrinsics, near, far, image_shape, torch.zeros_like(self.background_color),
                                gaussians.means, gaussians.covariances, dummy, gaussians.opacities, use_sh=False, depth_mode=mode)
        return depth


DECODERS = {"splatting_cuda": DecoderSplattingCUDA}
DecoderCfg = DecoderSplattingCUDACfg


def get_decoder(decoder_cfg: DecoderCfg, dataset_cfg: Any) -> Decoder:
    return DECODERS[decoder_cfg.name](decoder_cfg, dataset_cfg)
