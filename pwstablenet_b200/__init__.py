"""pwstablenet_b200 -- B200-native pixel-wise bilinear warp for PWStableNet.

One hot path only: the grid_sample warp (forward, backward, map composition,
frame sharding) behind the reference's own call signature.  The conv stack,
flags and train/process loops of the reference stay in PyTorch.
"""
from . import consumers, numa, sharding, windows
from .compose import compose_map, warp_fused, warp_stages
from .host import HostInferencePipeline, HostWarpPipeline, warp_host
from .functional import grid_sample, install, uninstall, warp2d_backward, warp2d_forward, warp_taps

__all__ = ["consumers", "windows", "sharding", "numa", "HostWarpPipeline", "HostInferencePipeline", "warp_host", "compose_map", "warp_fused", "warp_stages", "grid_sample", "install", "uninstall", "warp2d_forward", "warp2d_backward", "warp_taps"]
__version__ = "0.1.0"
