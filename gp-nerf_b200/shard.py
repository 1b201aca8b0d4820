"""Ray sharding over the GPUs of one box.

The image is cut into tiles of `tile_px` consecutive row-major pixels; tile t
belongs to rank t % world.  Every rank builds the same pixel mask, keeps the
rays of its own tiles (K1 does the filtering on the device) and renders them;
because the reference orders rays by ascending pixel index
(demo_render.py:200), concatenating the ranks' tiles in tile order reproduces
the single-GPU result exactly.  The only communication is one all_gather of
fixed-size pixel-tile buffers per frame.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def n_tiles(n_px: int, tile_px: int) -> int:
    return math.ceil(n_px / tile_px)


def tiles_per_rank(n_px: int, tile_px: int, world: int) -> int:
    return math.ceil(n_tiles(n_px, tile_px) / world)


def owner_of_pixel(p, tile_px: int, world: int):
    return (p // tile_px) % world


def local_pixel_index(n_px: int, tile_px: int, rank: int, world: int, device="cpu") -> torch.Tensor:
    """Flat pixel indices owned by `rank`, in this rank's tile order, padded with
    -1 to the fixed size tiles_per_rank·tile_px."""
    tpr = tiles_per_rank(n_px, tile_px, world)
    tiles = torch.arange(tpr, device=device) * world + rank
    px = tiles[:, None] * tile_px + torch.arange(tile_px, device=device)[None]
    px = torch.where(px < n_px, px, torch.full_like(px, -1))
    return px.reshape(-1)


def pack_local_tiles(image_flat: torch.Tensor, n_px: int, tile_px: int, rank: int, world: int) -> torch.Tensor:
    """image_flat [n_px, C] (full frame, only own tiles populated) → [tpr·tile_px, C]."""
    idx = local_pixel_index(n_px, tile_px, rank, world, image_flat.device)
    out = torch.zeros((idx.numel(),) + tuple(image_flat.shape[1:]), dtype=image_flat.dtype, device=image_flat.device)
    ok = idx >= 0
    out[ok] = image_flat[idx[ok]]
    return out


def unpack_gathered_tiles(gathered: torch.Tensor, n_px: int, tile_px: int, world: int) -> torch.Tensor:
    """gathered [world, tpr·tile_px, C] → full frame [n_px, C]."""
    out = torch.zeros((n_px,) + tuple(gathered.shape[2:]), dtype=gathered.dtype, device=gathered.device)
    for r in range(world):
        idx = local_pixel_index(n_px, tile_px, r, world, gathered.device)
        ok = idx >= 0
        out[idx[ok]] = gathered[r][ok]
    return out


def gather_frame(local_image_flat: torch.Tensor, n_px: int, tile_px: int, group=None) -> torch.Tensor:
    """One all_gather of this rank's pixel tiles; every rank returns the full
    frame [n_px, C].  NCCL on GPUs (NVLink), gloo in the CPU tests."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local_image_flat
    mine = pack_local_tiles(local_image_flat, n_px, tile_px, rank, world).contiguous()
    gathered = torch.empty((world * mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)      # rank-major concatenation along dim 0
    return unpack_gathered_tiles(gathered.view((world,) + tuple(mine.shape)), n_px, tile_px, world)
