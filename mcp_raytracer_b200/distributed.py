"""One-process-per-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards with no data-path collective: every pixel is independent and the scene is
replicated, so rank r renders the 8x4 pixel blocks with block_owner(...) == r (the same rule the
kernels apply, csrc/rt_types.h) into its own full-size framebuffer whose un-owned bytes stay
zero.  The ONE exchange step is getting the owned pixels into rank 0's framebuffer plus the RenderStats
merge (the roles SharedArrayBuffer and RenderStats.merge play in src/raytracer.ts:71-89):
`SharedFramebuffer` lets every rank's render kernel write its pixels straight into rank 0's memory
(CUDA IPC + NVLink peer stores: no collective on the data path); `gather_owned_blocks` is the
fallback that sends each rank's 1/world of the pixels through one NCCL gather.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

BLOCK_W, BLOCK_H = 8, 4


def block_owner(bx, by, blocks_per_row: int, n: int):
    """Mirror of rt::block_owner (csrc/rt_types.h): 8x4 pixel blocks numbered row-major over the whole image; each run of
    `n` consecutive blocks gives every part one block, in an order rotated by a hash of the run's index."""
    i = (np.asarray(by, np.uint64) * np.uint64(blocks_per_row) + np.asarray(bx, np.uint64)) & np.uint64(0xFFFFFFFF)
    g = i // np.uint64(n)
    r = i - g * np.uint64(n)
    m = np.uint64(0xFFFFFFFF)
    h = (g * np.uint64(0x9E3779B1)) & m
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & m
    h ^= h >> np.uint64(13)
    return ((r + h) % np.uint64(n)).astype(np.int64)


def tile_owner_mask(width: int, height: int, part_index: int, part_count: int) -> np.ndarray:
    """bool[H, W]: pixels rendered by `part_index` (mirror of the kernel's ownership rule)."""
    if part_count <= 1:
        return np.ones((height, width), bool)
    by, bx = np.meshgrid(np.arange(height) // BLOCK_H, np.arange(width) // BLOCK_W, indexing="ij")
    return block_owner(bx, by, (width + BLOCK_W - 1) // BLOCK_W, part_count) == part_index


class SharedFramebuffer:
    """RGB8 framebuffer in rank `dst`'s device memory that every rank of the node can write (CUDA IPC through the C ABI:
    rt_shared_buffer_*).  Each rank passes `.ptr` as the rgb8 device pointer of `Camera.renderRegionDevice`; its kernels
    then store the pixels it owns straight into rank dst's memory over NVLink — the exchange step of the path needs no
    collective at all, only the barrier that tells rank dst every rank's kernel has finished.  `ok` is False when the
    ranks cannot map each other's memory (no peer access, IPC refused): use `gather_owned_blocks` then."""

    def __init__(self, width: int, height: int, device: int, dst: int = 0):
        import ctypes as C

        from . import _native

        self.width, self.height, self.device, self.dst = width, height, device, dst
        self.nbytes = width * height * 3
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.ptr = 0
        self._opened = False
        L = _native.lib()
        handle = (C.c_uint8 * 64)()
        status = 0
        if self.rank == dst:
            p = C.c_void_p()
            status = L.rt_shared_buffer_create(device, self.nbytes, C.byref(p), handle)
            self.ptr = p.value or 0
        cuda = torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
        dev = torch.device("cuda", device) if cuda else torch.device("cpu")
        msg = torch.tensor(list(handle) + [status], dtype=torch.int32, device=dev)
        if self.world > 1:
            dist.broadcast(msg, src=dst)
        vals = msg.cpu().tolist()
        ok = vals[64] == 0
        if ok and self.rank != dst:
            h = (C.c_uint8 * 64)(*[v & 0xFF for v in vals[:64]])
            p = C.c_void_p()
            ok = L.rt_shared_buffer_open(device, h, C.byref(p)) == 0
            self.ptr = p.value or 0
            self._opened = ok
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(flag.item())

    def as_tensor(self) -> torch.Tensor:
        """uint8 [H, W, 3] view of the buffer (rank dst only)."""
        import ctypes as C

        class _Arr:  # minimal __cuda_array_interface__ carrier
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (self.height, self.width, 3), "typestr": "|u1", "data": (self.ptr, False), "version": 2}
        return torch.as_tensor(a, device=torch.device("cuda", self.device))

    def close(self) -> None:
        from . import _native

        if self.ptr:
            _native.lib().rt_shared_buffer_release(self.device, self.ptr, 1 if self._opened else 0)
            self.ptr = 0


def owned_pixel_index(width: int, height: int, part_index: int, part_count: int, device=None) -> torch.Tensor:
    """int64 flat pixel indices (j * W + i) of the blocks `part_index` owns, in row-major order."""
    idx = np.flatnonzero(tile_owner_mask(width, height, part_index, part_count).reshape(-1))
    return torch.as_tensor(idx, dtype=torch.int64, device=device)


def gather_owned_blocks(local_rgb8: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Fallback exchange step when ranks cannot write rank dst's memory: every rank sends ONLY the pixels it owns
    (1/world of the image) to rank `dst`, which scatters them into its framebuffer.  local_rgb8: uint8 [H, W, 3]; after
    the call rank dst's tensor holds the whole image."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local_rgb8
    world, rank = dist.get_world_size(), dist.get_rank()
    H, W = local_rgb8.shape[:2]
    flat = local_rgb8.view(-1, 3)
    idx = [owned_pixel_index(W, H, k, world, flat.device) for k in range(world)] if rank == dst else None
    mine = owned_pixel_index(W, H, rank, world, flat.device)
    n_blocks = -(-H // BLOCK_H) * -(-W // BLOCK_W)
    n_max = (-(-n_blocks // world) + 1) * BLOCK_H * BLOCK_W  # blocks are dealt one per part per run: same bound on every rank
    send = torch.zeros((n_max, 3), dtype=torch.uint8, device=flat.device)
    send[: mine.numel()] = flat[mine]
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank == dst:
        for k in range(world):
            if k != dst:
                flat[idx[k]] = recv[k][: idx[k].numel()]
    return local_rgb8


def gather_framebuffer(local_rgb8: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Kept name of the round-1 exchange step; now moves only owned pixels (gather_owned_blocks)."""
    return gather_owned_blocks(local_rgb8, dst)


def merge_stats(sums: torch.Tensor, mins: torch.Tensor, maxs: torch.Tensor, dst: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """RenderStats.merge (src/render-utils/renderStats.ts:42-64) across ranks: int64 tensors of
    totals (SUM), minima (MIN) and maxima (MAX), reduced onto rank `dst`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM)
        dist.reduce(mins, dst=dst, op=dist.ReduceOp.MIN)
        dist.reduce(maxs, dst=dst, op=dist.ReduceOp.MAX)
    return sums, mins, maxs
