"""Peer-memory image exchange between the GPUs of one NVLink box.

K5 (csrc/k4_k5_progressive.cu, ``composite_tiles``) writes every finished pixel
tile straight into the images of the ranks that need it – its own and, through
CUDA-IPC peer pointers, its peers' – and publishes the frame's sequence number
in the peers' arrival flags; ``gpnerf_peer_wait`` is the acquire on the other
side.  No NCCL call sits on the data path: torch.distributed is used once, at
set-up, to pass the 64-byte IPC handles around.

Per rank one IPC allocation::

    [2 halves][n_slots][H*W*3] float   images  (double-buffered by frame parity)
    [2 halves][n_slots][H*W]   uint8   hit masks
    [world]                    int32   arrival flags (flag[k] = last frame rank k finished)
    [1]                        int32   CTA ticket of the local K5 launch
    [1]                        int32   frame counter (advanced by K5 on the device)

The frame counter lives on the device: K5 renders frame ``*seq + 1`` into buffer
half ``(seq + 1) & 1`` and stores the number back when its last CTA retires, so
the host can queue frames ahead (CUDA-graph replays included) without touching
any per-frame state; ``PeerExchange.seq`` only mirrors how many frames were
queued, to know which half holds the newest image.

Two ways of spreading work:

``tiles``   one frame, its pixel tiles dealt over the ranks (K1 keeps the rays
            of the rank's tiles); every rank writes its tiles into EVERY rank's
            image → all ranks hold the full frame (all-gather semantics).
``frames``  one frame per rank (a sweep of `world` novel views); rank r writes
            its whole image into slot r of rank 0's buffer (gather semantics)
            and into its own slot 0.

Why double-buffering is enough: a rank can start writing frame f+1 into a peer
only after it passed the wait of frame f, i.e. after that peer finished its own
K5 of frame f – which the peer issues, in stream order, after whatever consumed
its image of frame f-1 (the half frame f+1 is written to).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import MAX_PEERS, Peer, check


class _DevView:
    """Zero-copy torch view of raw device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr, shape, dtype, device):
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


class PeerExchange:
    def __init__(self, H, W, device, rank, world, mode="tiles", group=None):
        import torch.distributed as dist
        if mode not in ("tiles", "frames"):
            raise _lib.GpnerfError("mode must be 'tiles' or 'frames'")
        if not 1 <= world <= MAX_PEERS:
            raise _lib.GpnerfError(f"peer exchange supports 1..{MAX_PEERS} ranks")
        self.lib = _lib.load()
        self.H, self.W, self.device, self.rank, self.world, self.mode = H, W, torch.device(device), rank, world, mode
        self.n_px = H * W
        self.n_slots = world if mode == "frames" else 1
        img_b, hit_b = self.n_px * 3 * 4, (self.n_px + 255) // 256 * 256
        self.off_img = [[(h * self.n_slots + s) * img_b for s in range(self.n_slots)] for h in range(2)]
        base_hit = 2 * self.n_slots * img_b
        self.off_hit = [[base_hit + (h * self.n_slots + s) * hit_b for s in range(self.n_slots)] for h in range(2)]
        self.off_flags = base_hit + 2 * self.n_slots * hit_b
        self.off_ticket = self.off_flags + 4 * MAX_PEERS
        self.off_seq = self.off_ticket + 64
        self.bytes = self.off_seq + 256
        # --- allocate, exchange the IPC handles, open the peers' buffers
        hb = self.lib.gpnerf_peer_handle_bytes()
        handle = C.create_string_buffer(hb)
        base = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.gpnerf_peer_alloc(self.bytes, C.byref(base), handle), "peer_alloc")
        self.base = [None] * world
        self.base[rank] = base.value
        self._opened = []
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            for k in range(world):
                if k == rank:
                    continue
                p = C.c_void_p()
                with torch.cuda.device(self.device):
                    check(self.lib.gpnerf_peer_open(C.create_string_buffer(handles[k], hb), C.byref(p)),
                          f"peer_open(rank {k})")
                self.base[k] = p.value
                self._opened.append(p.value)
            dist.barrier(group=group)
        self.seq = 0                 # host mirror of the device-side frame counter
        self._views = {}
        self.flags_ptr = self.base[rank] + self.off_flags
        # device-resident gpnerf_peer_t, written once
        p = self._struct()
        staging = torch.empty(C.sizeof(Peer), dtype=torch.uint8).pin_memory()
        C.memmove(staging.data_ptr(), C.addressof(p), C.sizeof(Peer))
        self.peer_dev = torch.empty(C.sizeof(Peer), dtype=torch.uint8, device=self.device)
        self.peer_dev.copy_(staging)
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------
    def _struct(self):
        p = Peer()
        if self.mode == "tiles":
            dst = [(self.rank, 0)] + [(k, 0) for k in range(self.world) if k != self.rank]
        else:
            dst = [(self.rank, 0)] + ([(0, self.rank)] if self.rank != 0 else [])
        p.n_dst = len(dst)
        for half in range(2):
            for i, (k, s) in enumerate(dst):
                p.dst_img[half][i] = self.base[k] + self.off_img[half][s]
                p.dst_hit[half][i] = self.base[k] + self.off_hit[half][s]
        others = [k for k in range(self.world) if k != self.rank]
        p.n_flag = len(others)
        for i, k in enumerate(others):
            p.dst_flag[i] = self.base[k] + self.off_flags + 4 * self.rank
        p.ticket = self.base[self.rank] + self.off_ticket
        p.seq = self.base[self.rank] + self.off_seq
        return p

    def next_frame(self):
        """Book-keeping for one more queued K5 launch (the device advances its
        own counter); returns the frame's sequence number."""
        self.seq += 1
        return self.seq

    def wait(self, stream_ptr):
        """Enqueue the arrival wait of the frame K5 just published (no-op for one rank)."""
        if self.world > 1:
            check(self.lib.gpnerf_peer_wait(C.c_void_p(self.flags_ptr), self.world, self.rank,
                                            C.c_void_p(self.peer_dev.data_ptr()), stream_ptr), "peer_wait")

    # ------------------------------------------------------------------ results of frame `seq`
    def image(self, slot=0, seq=None):
        half = (self.seq if seq is None else seq) & 1
        key = ("img", half, slot)
        if key not in self._views:      # views are built once: torch.as_tensor on a raw pointer is not free
            self._views[key] = _view(self.base[self.rank] + self.off_img[half][slot], (self.n_px, 3), torch.float32,
                                     self.device)
        return self._views[key]

    def hit_mask(self, slot=0, seq=None):
        half = (self.seq if seq is None else seq) & 1
        key = ("hit", half, slot)
        if key not in self._views:
            self._views[key] = _view(self.base[self.rank] + self.off_hit[half][slot], (self.n_px,), torch.uint8,
                                     self.device)
        return self._views[key]

    def close(self):
        self._views = {}
        for p in self._opened:
            self.lib.gpnerf_peer_close(C.c_void_p(p))
        self._opened = []
        if self.base[self.rank]:
            self.lib.gpnerf_peer_free(C.c_void_p(self.base[self.rank]))
            self.base[self.rank] = None
