"""Ray sharding over the GPUs of one box.

The image is cut into tiles of `tile_px` consecutive row-major pixels; tile t
belongs to rank ``(t + row(t)) % world`` where row(t) is the image row in which
the tile starts – a diagonal deal, so that a subject in the middle of the frame
is spread over all ranks (a plain ``t % world`` deal degenerates into vertical
strips whenever ``W / tile_px`` is a multiple of the world size).  Every rank
builds the same pixel mask, keeps the rays of its own tiles (K1 filters on the
device with the same rule, csrc/k1_rays.cu) and renders them; because the
reference orders rays by ascending pixel index (demo_render.py:200),
re-assembling the tiles reproduces the single-GPU result bit for bit.  The only
communication is one all_gather of fixed-size pixel-tile buffers per frame.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def n_tiles(n_px: int, tile_px: int) -> int:
    return math.ceil(n_px / tile_px)


def owner_of_tile(t, tile_px: int, width: int, world: int):
    """Rank owning tile t (int or tensor)."""
    return (t + (t * tile_px) // width) % world


def max_tiles_per_rank(n_px: int, width: int, tile_px: int, world: int) -> int:
    """Largest number of tiles any rank owns under the diagonal deal (host-side, pure Python ints)."""
    if world <= 1:
        return n_tiles(n_px, tile_px)
    counts = [0] * world
    for t in range(n_tiles(n_px, tile_px)):
        counts[owner_of_tile(t, tile_px, width, world)] += 1
    return max(counts)


def owner_of_pixel(p, tile_px: int, width: int, world: int):
    return owner_of_tile(p // tile_px, tile_px, width, world)


class TilePlan:
    """Index tensors of one (frame size, tile size, world) configuration,
    built once and cached: which pixels each rank owns (padded to a common
    length with -1) and the permutation that re-assembles a gathered frame."""

    _cache = {}

    def __init__(self, n_px, width, tile_px, world, device):
        tiles = torch.arange(n_tiles(n_px, tile_px), device=device)
        own = owner_of_tile(tiles, tile_px, width, world)
        per_rank = [tiles[own == r] for r in range(world)]
        self.tiles_per_rank = max(int(t.numel()) for t in per_rank)
        self.n_local = self.tiles_per_rank * tile_px
        offs = torch.arange(tile_px, device=device)
        idx = torch.full((world, self.n_local), -1, dtype=torch.long, device=device)
        for r, t in enumerate(per_rank):
            px = (t[:, None] * tile_px + offs[None]).reshape(-1)
            px = torch.where(px < n_px, px, torch.full_like(px, -1))
            idx[r, : px.numel()] = px
        self.local_idx = idx                                   # [world, n_local], -1 = padding
        self.local_idx_clamped = idx.clamp(min=0)
        flat = idx.reshape(-1)
        pos = torch.nonzero(flat >= 0).reshape(-1)
        inv = torch.empty(n_px, dtype=torch.long, device=device)
        inv[flat[pos]] = pos                                   # pixel → position in the gathered buffer
        self.inverse = inv
        self.n_px, self.width, self.tile_px, self.world = n_px, width, tile_px, world

    @classmethod
    def get(cls, n_px, width, tile_px, world, device):
        key = (n_px, width, tile_px, world, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(n_px, width, tile_px, world, device)
        return cls._cache[key]

    def pack(self, image_flat: torch.Tensor, rank: int) -> torch.Tensor:
        """[n_px, C] (own tiles populated) → [n_local, C] (one gather kernel)."""
        return image_flat.index_select(0, self.local_idx_clamped[rank])

    def unpack(self, gathered: torch.Tensor) -> torch.Tensor:
        """[world * n_local, C] → [n_px, C] (one gather kernel)."""
        return gathered.index_select(0, self.inverse)


def gather_frame(local_image_flat: torch.Tensor, width: int, tile_px: int, group=None) -> torch.Tensor:
    """One all_gather of this rank's pixel tiles; every rank returns the full
    frame [n_px, C].  NCCL over NVLink on GPUs, gloo in the CPU tests."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local_image_flat
    n_px = local_image_flat.shape[0]
    plan = TilePlan.get(n_px, width, tile_px, world, local_image_flat.device)
    mine = plan.pack(local_image_flat, rank).contiguous()
    gathered = torch.empty((world * plan.n_local,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)      # rank-major concatenation along dim 0
    return plan.unpack(gathered)


def all_gather_sharded_upload(host_tensor: torch.Tensor, out_dev: torch.Tensor, group=None) -> None:
    """Fill the replicated device tensor `out_dev` from a host tensor that every
    rank holds: each rank uploads only its 1/world slice over PCIe and the
    slices are exchanged over NVLink (one all_gather) – the host→device traffic
    per GPU drops by `world`."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    flat_out = out_dev.view(-1)
    n = flat_out.numel()
    if world == 1 or n % world != 0:
        out_dev.copy_(host_tensor, non_blocking=True)
        return
    per = n // world
    flat_out[rank * per:(rank + 1) * per].copy_(host_tensor.reshape(-1)[rank * per:(rank + 1) * per],
                                                 non_blocking=True)
    dist.all_gather_into_tensor(flat_out, flat_out[rank * per:(rank + 1) * per], group=group)
