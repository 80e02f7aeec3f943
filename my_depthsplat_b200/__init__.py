"""my_depthsplat_b200 -- B200-native (sm_100a) tile-based Gaussian-splatting rasterizer behind
DepthSplat's rendering API (render_cuda / render_depth_cuda / DecoderSplattingCUDA).

Only the rendering hot path lives here: hand-written CUDA kernels + a C-ABI
library (csrc/, include/b200splat.h) and the host-side mirror of the reference's decoder interface.
Importing the package is cheap and works without a GPU; calling a render function requires the built
library and a CUDA device -- there is no CPU fallback.
"""
from .types import DecoderOutput, DepthRenderingMode, Gaussians  # noqa: F401

__all__ = ["Gaussians", "DecoderOutput", "DepthRenderingMode"]
__version__ = "0.1.0"
