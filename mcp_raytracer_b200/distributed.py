"""One-process-per-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards with no data-path collective: every pixel is independent and the scene is
replicated, so rank r renders the 16x16 tiles with (tx + ty) % world == r (the same rule the
kernel applies, rt_megakernel.cu) into its own full-size framebuffer whose un-owned bytes stay
zero.  The ONE exchange step is the framebuffer gather to rank 0 plus the RenderStats merge
(the roles SharedArrayBuffer and RenderStats.merge play in src/raytracer.ts:71-89).  Because
ownership is disjoint and foreign bytes are zero, a SUM-reduce to rank 0 *is* the gather, and
moves one framebuffer per rank instead of `world` of them.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

TILE = 16


def tile_owner_mask(width: int, height: int, part_index: int, part_count: int) -> np.ndarray:
    """bool[H, W]: pixels rendered by `part_index` (mirror of the kernel's ownership rule)."""
    ty, tx = np.meshgrid(np.arange(height) // TILE, np.arange(width) // TILE, indexing="ij")
    if part_count <= 1:
        return np.ones((height, width), bool)
    return ((tx + ty) % part_count) == part_index


def gather_framebuffer(local_rgb8: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """local_rgb8: uint8 [H, W, 3] with zeros outside the rank's tiles.  After the call rank `dst`
    holds the full image in the same tensor; other ranks' tensors are unspecified."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(local_rgb8, dst=dst, op=dist.ReduceOp.SUM)
    return local_rgb8


def merge_stats(sums: torch.Tensor, mins: torch.Tensor, maxs: torch.Tensor, dst: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """RenderStats.merge (src/render-utils/renderStats.ts:42-64) across ranks: int64 tensors of
    totals (SUM), minima (MIN) and maxima (MAX), reduced onto rank `dst`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM)
        dist.reduce(mins, dst=dst, op=dist.ReduceOp.MIN)
        dist.reduce(maxs, dst=dst, op=dist.ReduceOp.MAX)
    return sums, mins, maxs
