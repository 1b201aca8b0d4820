"""Host-side driver of the hot path: owns the device buffers of one frame and
issues the K0…K5 launches of libgpnerf_b200.so on torch's current stream.

PyTorch is used for device memory and streams only; every computation on the
path is a kernel of the C-ABI library (gpnerf_b200._lib).  Data-dependent
sizes stay on the device (`counters`), so a frame is a sync-free launch
sequence; the host reads the counters once, together with the image.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import (CNT_P1, CNT_P2, CNT_PIX, CNT_RAYS, N_COUNTERS, PREC_BF16, PREC_FP32, Frame,
                   HeadWeights, check, ptr, ptr_array)

HEAD_KEYS = {
    "geo": ("sigmahead.out_geometry_fc.0",),
    "den": tuple(f"rgbhead.out_geometry_fc.{i}" for i in (0, 2, 4, 6)),
    "base": tuple(f"rgbhead.base_fc.{i}" for i in (0, 2)),
    "vis": tuple(f"rgbhead.vis_fc.{i}" for i in (0, 2)),
    "rgb": tuple(f"rgbhead.rgb_fc.{i}" for i in (0, 2, 4)),
}


# CUDA kernels launched by each C-ABI call (memsets not counted) – used for
# bench.py's `gpu_launches`; keep in sync with csrc/.
KERNELS_PER_CALL = {
    "k0_level_to_channels_last": 1, "k0_featmaps_to_channels_last": 1, "k0_images_to_rgbx": 1,
    "k0_products_to_f16": 1, "k0_sparse_to_f16": 1,
    "k0_build_masks3d": 1, "k1_voxel_pixel_mask": 3, "k1_rays_bbox": 3, "k2_occupancy_compact": 2,
    "k2_gather_volume": 1, "k2_project_gather_meanvar": 1, "k3_density_mlp": 1, "k4_compact_alpha": 2,
    "k4_compact_alpha_fused": 1,      # α and the flags come from the fused kernel: the single-pass compaction only
    "k3_color_mlp": 1, "k5_composite": 1, "k5_raw2outputs": 1, "peer_wait": 1,
    "k23_gather_density_tc": 1, "k3_color_mlp_records": 1, "k3_color_gather_tc": 1, "k3_color_tiles_tc": 1,
}


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def frame_from_batch(batch, H, W, n_views, n_samples, level_dims, src_hw, feat_hw,
                     voxel_size=(0.005,) * 3, mask_threshold=0.1, neg_ray=False, rank=0, world=1,
                     tile_px=64) -> Frame:
    """Pack gpnerf_frame_t from the reference's batch dict (host-side scalars;
    ZjumocapDataset.py:464-517).  K·E is formed with torch's bmm exactly as
    Projector.compute_projections does (BaseRender.py:311-314).  Pure host
    code: needs no GPU."""
    f = Frame()

    def cpu(x):
        return x.detach().to("cpu", torch.float32)
    Rm = cpu(batch["Rh"] if "Rh" in batch else batch["R"]).reshape(3, 3)
    f.R[:] = Rm.flatten().tolist()
    f.Th[:] = cpu(batch["Th"]).flatten().tolist()
    f.bounds_min[:] = cpu(batch["bounds"])[0, 0].tolist()
    f.voxel_size[:] = [float(v) for v in voxel_size]
    f.out_sh[:] = [int(v) for v in batch["out_sh"].reshape(-1, 3).max(0)[0].tolist()]
    for k in range(4):
        f.level_dims[k][:] = [int(v) for v in level_dims[k]]
    f.target_pose[:] = cpu(batch["target_pose"]).reshape(12).tolist()
    f.target_K[:] = cpu(batch["target_K"]).reshape(9).tolist()
    f.target_K_inv[:] = cpu(batch["target_K_inv"]).reshape(9).tolist()
    f.H, f.W = int(H), int(W)
    V = int(n_views)
    if not 1 <= V <= _lib.MAX_VIEWS:
        raise _lib.GpnerfError(f"n_views must be 1..{_lib.MAX_VIEWS}")
    f.n_views = V
    src_poses, src_Ks = cpu(batch["src_poses"]).reshape(V, 3, 4), cpu(batch["src_Ks"]).reshape(V, 3, 3)
    Eh = torch.eye(4).repeat(V, 1, 1)
    Eh[:, :3, :4] = src_poses
    Kh = torch.eye(4).repeat(V, 1, 1)
    Kh[:, :3, :3] = src_Ks
    KE = Kh.bmm(Eh)
    for v in range(V):
        f.src_KE[v][:] = KE[v].flatten().tolist()
    f.src_h, f.src_w = int(src_hw[0]), int(src_hw[1])
    f.feat_h, f.feat_w = int(feat_hw[0]), int(feat_hw[1])
    f.n_samples = int(n_samples)
    f.neg_ray = int(bool(neg_ray))
    f.mask_threshold = float(mask_threshold)
    f.rank, f.world, f.tile_px = int(rank), int(world), int(tile_px)
    return f


class Engine:
    """Buffers + launch sequence for one target resolution / sample count."""

    def __init__(self, H, W, n_samples, n_views, device="cuda:0", precision=PREC_FP32,
                 rank=0, world=1, tile_px=64, max_rays=None, t_min=0.0, voxel_size=(0.005,) * 3,
                 mask_threshold=0.1, fused_gather=True):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GpnerfError("gpnerf_b200 runs on CUDA devices only (no CPU fallback)")
        self.H, self.W, self.S, self.V = int(H), int(W), int(n_samples), int(n_views)
        self.precision = int(precision)
        self.rank, self.world, self.tile_px = int(rank), int(world), int(tile_px)
        self.t_min = float(t_min)
        self.voxel_size = tuple(float(v) for v in voxel_size)
        self.mask_threshold = float(mask_threshold)
        npx = self.H * self.W
        if max_rays is None:
            # capacity = the most pixels any rank can own under the diagonal tile deal (it is not perfectly
            # balanced: 512², tile 64, world 3 gives one rank 1,536 of the 4,096 tiles), not ceil(tiles / world)
            from .shard import max_tiles_per_rank
            max_rays = min(npx, max_tiles_per_rank(npx, self.W, self.tile_px, self.world) * self.tile_px)
        self.max_rays = int(max_rays)
        self.max_pts = self.max_rays * self.S
        if self.max_pts >= 2 ** 31:
            raise _lib.GpnerfError("ray·sample count exceeds int32 indexing")
        dev = self.device
        f32, i32 = torch.float32, torch.int32

        def buf(n, dt=f32):
            return torch.empty(int(n), dtype=dt, device=dev)

        self.counters = torch.zeros(N_COUNTERS, dtype=i32, device=dev)
        self.can_bounds = buf(12)
        self.pix_mask = buf(npx)
        self.ray_pix = buf(npx, i32)
        self.rays_o = buf(3)
        self.rays_d = buf(max(npx, self.max_rays) * 3)
        self.near = buf(max(npx, self.max_rays))
        self.far = buf(max(npx, self.max_rays))
        self.t_vals = torch.linspace(0.0, 1.0, steps=self.S, device="cpu").to(dev)  # BaseRender.py:37
        self.valid = buf(self.max_pts, i32)
        self.z_vals = buf(self.max_pts)
        # fused_gather=False with PREC_BF16: fp32 gathers (reference rounding) feeding the tcgen05 heads – the
        # error-budget configuration (tests/test_gpu_parity.py: which share of the bf16 path's image error comes
        # from the 16-bit storage + HFMA2 interpolation of the fused kernel, which from the bf16 MLPs)
        self.bf16 = self.precision == PREC_BF16 and bool(fused_gather)
        if self.bf16:
            # tensor-core path: gathered features stay on chip; one bf16 record per point for the colour head
            if not 1 <= self.V <= 4:
                raise _lib.GpnerfError("the bf16 tensor-core path supports 1..4 source views")
            # Hand-off from the fused gather → density kernel to the colour head (GPNERF_COLOR_IMPL):
            #   tiles   (default) one block per 128-point tile of the P1 list in the colour head's own operand
            #           layouts, fetched with one bulk copy per tile: the per-view features are gathered once
            #   gather  nothing is handed over; the colour head gathers its inputs again for the survivors
            #   records round 1: one 16(9+5V)-byte record per P1 point, written by round 1's monolithic kernel
            # Unset = auto: starts on gather and is re-decided from the survivor ratio of the frames rendered so far
            # (note_counts): the tile-fed head processes every tile that has a survivor, i.e. ≈P1 points, the
            # gathering head P2; on the benchmark frame (P2 / P1 = 0.90) tiles win by 0.03 ms, on the 1024² S=128
            # frame (0.14) they lose 2 ms – and their record buffer (384 B per point of capacity) is only
            # allocated once a frame asks for it
            import os
            impl = os.environ.get("GPNERF_COLOR_IMPL", "") or "auto"
            if impl not in ("auto", "tiles", "gather", "records"):
                raise _lib.GpnerfError(f"GPNERF_COLOR_IMPL={impl!r}: expected auto, tiles, gather or records")
            self.color_impl_auto = impl == "auto"
            self.color_impl = "gather" if impl == "auto" else impl
            impl = self.color_impl
            self.use_records = impl == "records"
            self.rec_bytes = int(self.lib.gpnerf_k23_record_bytes(self.V))
            self.rec = torch.empty(self.max_pts * self.rec_bytes, dtype=torch.uint8, device=dev) if self.use_records else None
            # tile records: zero-filled once (columns no view owns are never written)
            self.tile_rec_bytes = int(self.lib.gpnerf_k23_tile_record_bytes(self.V))
            self.rec_tiles = (torch.zeros(((self.max_pts + 127) // 128) * self.tile_rec_bytes, dtype=torch.uint8, device=dev)
                              if impl == "tiles" else None)
            self.tiles_min_survival = 0.85
            self.rgb_in = None       # per-view RGB taps [P1][V][3], dense path only (allocated on first use)
            self.vol_feat = self.rgb_feat = self.mask = self.meanvar = None
        else:
            self.rec = self.rec_tiles = None
            self.color_impl, self.color_impl_auto = None, False
            self.vol_feat = buf(self.max_pts * 128)
            self.rgb_feat = buf(self.max_pts * self.V * 35)
            self.mask = buf(self.max_pts * self.V)
            self.meanvar = buf(self.max_pts * 70)
        self.sigma = buf(self.max_pts)
        self.alpha = buf(self.max_pts)
        self.valid1 = buf(self.max_pts, i32)
        self.rgb = buf(self.max_pts * 3)
        self.rgb_map = buf(self.max_rays * 3)
        # K5 rewrites every pixel of this rank's tiles each frame (zeros where no ray); the tiles of other
        # ranks are never touched, so they are cleared once here
        self.pred_img = torch.zeros(npx * 3, dtype=f32, device=dev)
        self.hit_mask = torch.zeros(npx, dtype=torch.uint8, device=dev)
        # CSR offsets left behind by the two compactions: rays of every pixel tile, surviving points of every ray
        self.tile_ray_begin = torch.zeros(math.ceil(npx / self.tile_px) + 1, dtype=i32, device=dev)
        self.ray_pt_begin = torch.zeros(self.max_rays + 1, dtype=i32, device=dev)
        self.exchange = None         # peer.PeerExchange when the image leaves through peer memory
        ws_bytes = self.lib.gpnerf_workspace_bytes(max(self.max_pts, npx))
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        # products of K0 (allocated on first upload)
        self.levels_cl = None
        self.chan_sums = None
        self.masks3d = None
        self.images_rgbx = None
        self.featmaps_cl = None
        self.level_dims = None
        self._weights = None
        self._weight_tensors = None
        self.launches = 0
        # device-resident copy of gpnerf_frame_t (kernels read per-frame values from it, so a
        # captured CUDA graph can be replayed for a new pose) + its pinned staging buffer
        self.frame_pinned = torch.empty(C.sizeof(Frame), dtype=torch.uint8).pin_memory()
        self.frame_dev = torch.empty(C.sizeof(Frame), dtype=torch.uint8, device=dev)
        self._graphs = {}            # (shapes, with_k0) → (CUDAGraph, launches per replay)
        self._static_inputs = None
        self.timing = False          # when set, CUDA events bracket every stage (bench.py)
        self.stage_events = {}

    # ------------------------------------------------------------------ utils
    def _tic(self, name):
        if not self.timing:
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(torch.cuda.current_stream(self.device))
        self.stage_events.setdefault(name, []).append(ev)
        return ev

    def _toc(self, ev):
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))

    def stage_times_ms(self, reset=True):
        """Mean device time per stage (call after a synchronize)."""
        out = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in self.stage_events.items() if v}
        if reset:
            self.stage_events = {}
        return out

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_weights(self, state_dict):
        """Head weights keyed as in the reference state_dict (prefix
        'nerfhead.' optional)."""
        from .ops import pack_head_weights
        hw, keep = pack_head_weights(state_dict, self.device, self.V)
        if keep[2 * (1 + 4 + 2 + 2)].shape[1] != 32 * self.V:
            raise _lib.GpnerfError("rgb_fc.0 expects 32·n_views input features")
        self._weights, self._weight_tensors = hw, keep
        self._graphs = {}           # a captured graph holds the old weight pointers

    def _run(self, name, fn, *args):
        ev = self._tic(name)
        check(fn(*args), name)
        self._toc(ev)
        self.launches += KERNELS_PER_CALL[name]

    # ------------------------------------------------------------- K0 uploads
    def upload_products(self, levels, featmaps, src_imgs):
        """levels: 4 × [1,32,D,H,W] fp32 (SparseConvTensor.dense() layout);
        featmaps [V,32,h,w]; src_imgs [1,V,3,H,W] or [V,3,H,W] in [-1,1].
        Tensors already on the device are used in place; host tensors are
        copied (non-blocking when pinned)."""
        dev, st, L = self.device, self._stream(), self.lib
        lv = [t.to(dev, non_blocking=True).contiguous() for t in levels]
        fm = featmaps.to(dev, non_blocking=True).contiguous()
        im = src_imgs.to(dev, non_blocking=True)
        im = (im[0] if im.dim() == 5 else im).contiguous()
        dims = [tuple(int(v) for v in t.shape[-3:]) for t in lv]
        if self.level_dims != dims:
            self.level_dims = dims
            # tensor-core path: channel-last fp16 inside a zero border (allocated zeroed once; K0 only
            # ever writes the interior), so the fused gather needs no bounds tests
            if self.bf16:
                self.levels_cl = [torch.zeros((d + 2) * (h + 2) * (w + 2) * 32, dtype=torch.float16, device=dev)
                                  for d, h, w in dims]
            else:
                self.levels_cl = [torch.empty(d * h * w * 32, dtype=torch.float32, device=dev) for d, h, w in dims]
            self.chan_sums = [torch.empty(d * h * w, dtype=torch.float32, device=dev) for d, h, w in dims]
            self.masks3d = torch.empty(dims[0][0] * dims[0][1] * dims[0][2], dtype=torch.float32, device=dev)
        assert len(lv) == 4 and all(t.shape[1] == 32 and t.dtype == torch.float32 for t in lv)
        pad = int(self.bf16)
        V, Cc, fh, fw = fm.shape
        assert V == self.V and Cc == 32
        n_fm = V * (fh + 2 * pad) * (fw + 2 * pad) * 32
        if self.featmaps_cl is None or self.featmaps_cl.numel() != n_fm:
            self.featmaps_cl = torch.zeros(n_fm, dtype=torch.float16 if self.bf16 else torch.float32, device=dev)
        if self.bf16:
            # tensor-core path: the 4 levels and the maps → zero-bordered fp16 lines in one launch
            dims_c = ((C.c_int32 * 3) * 4)(*[(C.c_int32 * 3)(*d) for d in dims])
            self._run("k0_products_to_f16", L.gpnerf_k0_products_to_f16, ptr_array(lv), dims_c, ptr(fm), V, fh, fw,
                      ptr_array(self.levels_cl), ptr_array(self.chan_sums), ptr(self.featmaps_cl), st)
        else:
            for t, (d, h, w), cl, cs in zip(lv, dims, self.levels_cl, self.chan_sums):
                self._run("k0_level_to_channels_last", L.gpnerf_k0_level_to_channels_last, ptr(t), d, h, w,
                          0, 0, ptr(cl), ptr(cs), st)
            self._run("k0_featmaps_to_channels_last", L.gpnerf_k0_featmaps_to_channels_last, ptr(fm), V, fh, fw,
                      0, 0, ptr(self.featmaps_cl), st)
        _, _, ih, iw = im.shape
        n_im = V * (ih + 2 * pad) * (iw + 2 * pad) * 4
        if self.images_rgbx is None or self.images_rgbx.numel() != n_im:
            self.images_rgbx = torch.zeros(n_im, dtype=torch.float32, device=dev)
        self._run("k0_images_to_rgbx", L.gpnerf_k0_images_to_rgbx, ptr(im), V, ih, iw, 1, pad,
                  ptr(self.images_rgbx), st)
        self.src_hw, self.feat_hw = (ih, iw), (fh, fw)
        self._keep_inputs = (lv, fm, im)     # keep alive until the stream drains

    def upload_products_sparse(self, levels_sparse, level_dims, featmaps, src_imgs, n_rows_dev=None):
        """The pyramid's levels as the sparse-conv network holds them before
        `.dense()` (SparseConvNet.py:110): per level `(features [N,32] fp32,
        indices [N,3|4] int32 with (d,h,w) last)`, `level_dims` 4 × (D,H,W).
        The active rows are scattered straight into the gather layouts – the
        bordered fp16 volumes of the tensor-core path or the fp32 channel-last
        volumes of the exact path (no dense NCDHW tensor, no K0 transposition)."""
        dev, st, L = self.device, self._stream(), self.lib
        dims = [tuple(int(v) for v in d) for d in level_dims]
        if len(levels_sparse) != 4 or len(dims) != 4:
            raise _lib.GpnerfError("4 levels expected")
        feats = [f.to(dev, non_blocking=True).to(torch.float32).contiguous() for f, _ in levels_sparse]
        idxs = [i.to(dev, non_blocking=True).to(torch.int32).contiguous() for _, i in levels_sparse]
        cols = int(idxs[0].shape[1])
        assert all(f.shape[1] == 32 and i.shape[1] == cols and i.shape[0] == f.shape[0] for f, i in zip(feats, idxs))
        fm = featmaps.to(dev, non_blocking=True).contiguous()
        im = src_imgs.to(dev, non_blocking=True)
        im = (im[0] if im.dim() == 5 else im).contiguous()
        pad = int(self.bf16)
        if self.level_dims != dims:
            self.level_dims = dims
            if self.bf16:
                self.levels_cl = [torch.zeros((d + 2) * (h + 2) * (w + 2) * 32, dtype=torch.float16, device=dev)
                                  for d, h, w in dims]
            else:
                self.levels_cl = [torch.zeros(d * h * w * 32, dtype=torch.float32, device=dev) for d, h, w in dims]
            self.chan_sums = [torch.empty(d * h * w, dtype=torch.float32, device=dev) for d, h, w in dims]
            self.masks3d = torch.empty(dims[0][0] * dims[0][1] * dims[0][2], dtype=torch.float32, device=dev)
        dims_c = ((C.c_int32 * 3) * 4)(*[(C.c_int32 * 3)(*d) for d in dims])
        n_rows = (C.c_int32 * 4)(*[int(f.shape[0]) for f in feats])
        # n_rows_dev: 4 device int32 scalars with the live row counts (the arrays are then capacities:
        # what sparseconv.SparseConvNet hands over without a host sync)
        nrd = None if n_rows_dev is None else ptr_array([t.view(1) for t in n_rows_dev])
        if self.bf16:
            self._run("k0_sparse_to_f16", L.gpnerf_k0_sparse_to_f16, ptr_array(feats), ptr_array(idxs), n_rows, nrd, cols,
                      dims_c, ptr_array(self.levels_cl), ptr_array(self.chan_sums), st)
        else:
            self._run("k0_sparse_to_f16", L.gpnerf_k0_sparse_to_f32, ptr_array(feats), ptr_array(idxs), n_rows, nrd, cols,
                      dims_c, ptr_array(self.levels_cl), ptr_array(self.chan_sums), st)
        V, Cc, fh, fw = fm.shape
        assert V == self.V and Cc == 32
        n_fm = V * (fh + 2 * pad) * (fw + 2 * pad) * 32
        fm_dtype = torch.float16 if self.bf16 else torch.float32
        if self.featmaps_cl is None or self.featmaps_cl.numel() != n_fm or self.featmaps_cl.dtype != fm_dtype:
            self.featmaps_cl = torch.zeros(n_fm, dtype=fm_dtype, device=dev)
        self._run("k0_featmaps_to_channels_last", L.gpnerf_k0_featmaps_to_channels_last, ptr(fm), V, fh, fw,
                  2 if self.bf16 else 0, pad, ptr(self.featmaps_cl), st)
        _, _, ih, iw = im.shape
        n_im = V * (ih + 2 * pad) * (iw + 2 * pad) * 4
        if self.images_rgbx is None or self.images_rgbx.numel() != n_im:
            self.images_rgbx = torch.zeros(n_im, dtype=torch.float32, device=dev)
        self._run("k0_images_to_rgbx", L.gpnerf_k0_images_to_rgbx, ptr(im), V, ih, iw, 1, pad, ptr(self.images_rgbx), st)
        self.src_hw, self.feat_hw = (ih, iw), (fh, fw)
        self._keep_inputs = (feats, idxs, fm, im)

    # ------------------------------------------------------------ frame setup
    _STATIC_KEYS = ("Rh", "R", "Th", "bounds", "out_sh", "src_poses", "src_Ks")

    def make_frame(self, batch, neg_ray=False):
        """gpnerf_frame_t of a batch.  In a sweep only the target camera changes from frame to frame: the
        rest of the struct (SMPL pose, bounds, the K·E products of the source views) is packed once per
        distinct VALUE of those inputs (their bytes are the cache key) and copied."""
        vals = torch.cat([batch[k].detach().to("cpu", torch.float64).reshape(-1) for k in self._STATIC_KEYS if k in batch])
        key = (vals.numpy().tobytes(), tuple(self.level_dims or ()), self.src_hw, self.feat_hw, bool(neg_ray))
        cached = getattr(self, "_frame_cache", None)
        if cached is not None and cached[0] == key:
            f = Frame.from_buffer_copy(cached[1])
            f.target_pose[:] = batch["target_pose"].detach().to("cpu", torch.float32).reshape(12).tolist()
            f.target_K[:] = batch["target_K"].detach().to("cpu", torch.float32).reshape(9).tolist()
            f.target_K_inv[:] = batch["target_K_inv"].detach().to("cpu", torch.float32).reshape(9).tolist()
            return f
        f = frame_from_batch(batch, H=self.H, W=self.W, n_views=self.V, n_samples=self.S,
                             level_dims=self.level_dims, src_hw=self.src_hw, feat_hw=self.feat_hw,
                             voxel_size=self.voxel_size, mask_threshold=self.mask_threshold,
                             neg_ray=neg_ray, rank=self.rank, world=self.world, tile_px=self.tile_px)
        f.self_dev = self.frame_dev.data_ptr()
        self._frame_cache = (key, bytes(f))
        return f

    def upload_frame(self, frame):
        """Refresh the device copy of the frame constants (pinned host → device,
        on the current stream).  Every render entry point calls it first."""
        # four pinned staging slots in rotation, each guarded by an event: a second frame queued while the GPU is
        # still behind must not overwrite constants whose host→device copy has not executed yet
        raw = C.string_at(C.addressof(frame), C.sizeof(Frame))
        if torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture (train.GraphedStep) nothing may wait on an event, and a copy node would
            # re-read a pinned slot that later frames overwrite: the constants must already be on the device
            if raw != getattr(self, "_frame_bytes", None):
                raise _lib.GpnerfError("upload_frame during CUDA-graph capture: run one step with these frame "
                                       "constants before capturing")
            return
        self._frame_bytes = raw
        slots = getattr(self, "_frame_slots", None)
        if slots is None:
            slots = self._frame_slots = [[torch.empty(C.sizeof(Frame), dtype=torch.uint8).pin_memory(), None]
                                         for _ in range(4)]
            self._frame_slot_i = 0
        slot = slots[self._frame_slot_i]
        self._frame_slot_i = (self._frame_slot_i + 1) % len(slots)
        if slot[1] is not None:
            slot[1].synchronize()
        C.memmove(slot[0].data_ptr(), C.addressof(frame), C.sizeof(Frame))
        self.frame_dev.copy_(slot[0], non_blocking=True)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        if not torch.cuda.is_current_stream_capturing():
            slot[1].record(torch.cuda.current_stream(self.device))

    def attach_exchange(self, exchange):
        """Publish the rendered tiles through peer memory (peer.PeerExchange)
        instead of (only) the local pred_img / hit_mask tensors."""
        self.exchange = exchange
        self._graphs = {}

    def result_image(self, slot=0):
        """[H*W, 3] image of the frame rendered last (a view: copy it before
        rendering two more frames when an exchange double-buffers it)."""
        if self.exchange is not None:
            return self.exchange.image(slot)
        return self.pred_img.view(-1, 3)

    def result_hit_mask(self, slot=0):
        if self.exchange is not None:
            return self.exchange.hit_mask(slot)
        return self.hit_mask

    # ------------------------------------------------------------ CUDA graph
    def set_static_inputs(self, levels, featmaps, src_imgs):
        """Device tensors (reference layouts) the captured graph reads its
        upstream products from; refresh their *contents* between replays."""
        dev = self.device
        lv = [t.to(dev).contiguous() for t in levels]
        fm = featmaps.to(dev).contiguous()
        im = src_imgs.to(dev)
        im = (im[0] if im.dim() == 5 else im).contiguous()
        self._static_inputs = (lv, fm, im)
        self._graphs = {}

    def copy_into_static_inputs(self, levels, featmaps, src_imgs, sharded_upload=False):
        """Host (pinned) or device tensors → the static input buffers.  With
        `sharded_upload` (world > 1, host inputs replicated on every rank) each
        rank uploads 1/world of every tensor and the slices are all-gathered
        over NVLink (shard.all_gather_sharded_upload)."""
        im = src_imgs[0] if src_imgs.dim() == 5 else src_imgs
        cur = self._static_inputs
        same = (cur is not None and [tuple(t.shape) for t in cur[0]] == [tuple(t.shape) for t in levels]
                and tuple(cur[1].shape) == tuple(featmaps.shape) and tuple(cur[2].shape) == tuple(im.shape))
        if not same:
            self.set_static_inputs(levels, featmaps, src_imgs)
            return
        lv, fm, imd = self._static_inputs
        pairs = list(zip(lv, levels)) + [(fm, featmaps), (imd, im)]
        if sharded_upload and self.world > 1:
            from .shard import all_gather_sharded_upload
            for d, s in pairs:
                if s.is_cuda:
                    d.copy_(s, non_blocking=True)
                else:
                    all_gather_sharded_upload(s, d)
        else:
            for d, s in pairs:
                d.copy_(s, non_blocking=True)

    def run_progressive_graphed(self, frame, with_k0=True, frame_src=None):
        """One frame as ONE CUDA-graph launch.  The graph is captured on first
        use (after an eager warm-up frame) and replayed afterwards.

        with_k0=True   the graph is [frame constants ← pinned buffer, K0 from the
                       static inputs, K1…K5]: call, then sync before changing
                       the frame (the replay reads the pinned buffer when it runs).
        with_k0=False  the caller has already launched upload_products(); the
                       graph covers K1…K5 only and the frame constants are copied
                       eagerly from `frame_src` (a pinned uint8 tensor the caller
                       keeps untouched until the frame has run) – frames can then
                       be queued ahead of the GPU (Renderer.render_stream)."""
        if with_k0 and self._static_inputs is None:
            raise _lib.GpnerfError("set_static_inputs() first")
        impl = self.color_impl              # read once: note_counts may change it from a worker thread
        key = (tuple(self.level_dims or ()), getattr(self, "src_hw", None), getattr(self, "feat_hw", None), with_k0, impl)
        graphs = self._graphs
        if key not in graphs:
            timing, self.timing = self.timing, False
            if with_k0:
                lv, fm, im = self._static_inputs
                self.upload_products(lv, fm, im)             # eager warm-up: allocations, func attributes
            self.render_progressive(frame, impl=impl)
            torch.cuda.synchronize(self.device)
            key = (tuple(self.level_dims), self.src_hw, self.feat_hw, with_k0, impl)
            C.memmove(self.frame_pinned.data_ptr(), C.addressof(frame), C.sizeof(Frame))
            g = torch.cuda.CUDAGraph()
            l0 = self.launches
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                if with_k0:
                    self.frame_dev.copy_(self.frame_pinned, non_blocking=True)
                    self.upload_products(lv, fm, im)
                self.render_progressive(frame, upload=False, impl=impl)
            graphs[key] = (g, self.launches - l0)
            self.timing = timing
        g, n_launch = graphs[key]
        if with_k0:
            C.memmove(self.frame_pinned.data_ptr(), C.addressof(frame), C.sizeof(Frame))
        else:
            if frame_src is None:
                raise _lib.GpnerfError("with_k0=False needs frame_src (pinned copy of the frame constants)")
            C.memmove(frame_src.data_ptr(), C.addressof(frame), C.sizeof(Frame))
            self.frame_dev.copy_(frame_src, non_blocking=True)
        if self.exchange is not None:
            self.exchange.next_frame()       # one more K5 launch queued (its frame counter lives on the device)
        g.replay()
        self.launches += n_launch
        if self.bf16 and impl == "tiles":
            self._valid1_pending = True

    # --------------------------------------------------------------- launches
    def build_occupancy(self, frame):
        self._run("k0_build_masks3d", self.lib.gpnerf_k0_build_masks3d, ptr_array(self.chan_sums),
                  C.byref(frame), ptr(self.masks3d), self._stream())

    def render_progressive(self, frame, t_rand=None, upload=True, impl=None):
        """demo_render.Renderer.render_rays downstream of the producers.
        Leaves results in self.{rgb_map,pred_img,hit_mask,counters,...}."""
        if self._weights is None:
            raise _lib.GpnerfError("set_weights() has not been called")
        L, st, fr = self.lib, self._stream(), C.byref(frame)
        impl = self._impl_now = impl or self.color_impl      # one hand-off per frame, whatever note_counts does meanwhile
        if upload:
            self.upload_frame(frame)
        self.build_occupancy(frame)
        self._run("k1_voxel_pixel_mask", L.gpnerf_k1_voxel_pixel_mask, ptr(self.masks3d), fr,
                  ptr(self.can_bounds), ptr(self.pix_mask), st)
        self._run("k1_rays_bbox", L.gpnerf_k1_rays_bbox, ptr(self.pix_mask), ptr(self.can_bounds), fr,
                  ptr(self.ray_pix), ptr(self.rays_o), ptr(self.rays_d), ptr(self.near), ptr(self.far),
                  ptr(self.counters), ptr(self.workspace), ptr(self.tile_ray_begin), st)
        self._heads(frame, masks3d=self.masks3d, t_rand=t_rand, n_rays_max=self.max_rays, fuse_alpha=True)
        if self.bf16 and impl == "tiles":
            # tile hand-off: the colour head takes every tile of the P1 list that has a survivor (K5 ignores the
            # colour of a culled point) and counts the survivors; the ordered survivor list (valid1) is not on the
            # frame's path any more – survivor_list() / read_counters() produce it on demand from the flags the
            # fused kernel left in the workspace
            self._run("k3_color_tiles_tc", L.gpnerf_k3_color_tiles_tc, ptr(self.rec_tiles), ptr(self.workspace),
                      C.byref(self._weights), self.V, self.max_pts, ptr(self.counters), CNT_P1, ptr(self.rgb), st)
            self._valid1_pending = True
        else:
            # tensor-core path: α and the survivor flags were written by the fused kernel's epilogue
            self._valid1_pending = False
            self._run("k4_compact_alpha_fused" if self.bf16 else "k4_compact_alpha", L.gpnerf_k4_compact_alpha,
                      None if self.bf16 else ptr(self.sigma), self.max_pts,
                      ptr(self.counters), ptr(self.alpha), ptr(self.valid1), ptr(self.workspace), st)
            self._color(ptr(self.valid1), self.max_pts, CNT_P2, frame)
        ex = self.exchange
        self._run("k5_composite", L.gpnerf_k5_composite, ptr(self.alpha), ptr(self.rgb), ptr(self.ray_pix),
                  ptr(self.tile_ray_begin), ptr(self.ray_pt_begin), fr, C.c_float(self.t_min), ptr(self.rgb_map),
                  ptr(self.pred_img), ptr(self.hit_mask), None if ex is None else ptr(ex.peer_dev), st)
        if ex is not None and not torch.cuda.is_current_stream_capturing():
            ex.next_frame()
        if ex is not None and ex.world > 1:
            ev = self._tic("peer_wait")
            ex.wait(st)
            self._toc(ev)
            self.launches += 1

    def _color(self, valid1_ptr, n_pts_max, slot, frame=None, rgb_in=None):
        L, st = self.lib, self._stream()
        if self.bf16 and not self.use_records:
            self._run("k3_color_gather_tc", L.gpnerf_k3_color_gather_tc, ptr(self.featmaps_cl), ptr(self.images_rgbx),
                      ptr(self.valid), valid1_ptr, ptr(self.rays_o), ptr(self.rays_d), ptr(self.z_vals), C.byref(frame),
                      C.byref(self._weights), n_pts_max, ptr(self.counters), slot, ptr(self.rgb), ptr(rgb_in), st)
        elif self.bf16:
            self._run("k3_color_mlp_records", L.gpnerf_k3_color_mlp_records, ptr(self.rec), valid1_ptr,
                      C.byref(self._weights), self.V, n_pts_max, ptr(self.counters), slot, ptr(self.rgb), st)
        else:
            self._run("k3_color_mlp", L.gpnerf_k3_color_mlp, ptr(self.rgb_feat), ptr(self.meanvar), valid1_ptr,
                      C.byref(self._weights), self.V, n_pts_max, ptr(self.counters), slot, ptr(self.rgb),
                      self.precision, st)

    def _heads(self, frame, masks3d, t_rand, n_rays_max, fuse_alpha=False, rgb_in=None):
        """occupancy (or identity) compaction → gathers → density head."""
        L, st, fr = self.lib, self._stream(), C.byref(frame)
        n_pts_max = n_rays_max * self.S
        self._run("k2_occupancy_compact", L.gpnerf_k2_occupancy_compact, ptr(masks3d), ptr(self.rays_o),
                  ptr(self.rays_d), ptr(self.near), ptr(self.far), ptr(self.t_vals), ptr(t_rand), fr, n_rays_max,
                  ptr(self.valid), ptr(self.z_vals), ptr(self.counters), ptr(self.workspace),
                  ptr(self.ray_pt_begin), st)
        if self.bf16 and self._impl_now == "tiles":
            if self.rec_tiles is None:          # (auto mode that started on another hand-off)
                self.rec_tiles = torch.zeros(((self.max_pts + 127) // 128) * self.tile_rec_bytes, dtype=torch.uint8,
                                             device=self.device)
            self._run("k23_gather_density_tc", L.gpnerf_k23_gather_density_tiles_tc, ptr_array(self.levels_cl),
                      ptr(self.featmaps_cl), ptr(self.images_rgbx), ptr(self.valid), ptr(self.rays_o),
                      ptr(self.rays_d), ptr(self.z_vals), fr, C.byref(self._weights), n_pts_max,
                      ptr(self.counters), ptr(self.sigma), ptr(self.rec_tiles), ptr(rgb_in),
                      ptr(self.alpha) if fuse_alpha else None, ptr(self.workspace) if fuse_alpha else None, st)
            return
        if self.bf16:
            self._run("k23_gather_density_tc", L.gpnerf_k23_gather_density_tc, ptr_array(self.levels_cl),
                      ptr(self.featmaps_cl), ptr(self.images_rgbx), ptr(self.valid), ptr(self.rays_o),
                      ptr(self.rays_d), ptr(self.z_vals), fr, C.byref(self._weights), n_pts_max,
                      ptr(self.counters), ptr(self.sigma), ptr(self.rec) if self.use_records else None,
                      ptr(self.alpha) if fuse_alpha else None, ptr(self.workspace) if fuse_alpha else None, st)
            return
        self._run("k2_gather_volume", L.gpnerf_k2_gather_volume, ptr_array(self.levels_cl), 0, ptr(self.valid),
                  ptr(self.rays_o), ptr(self.rays_d), ptr(self.z_vals), None, fr, n_pts_max, ptr(self.counters),
                  ptr(self.vol_feat), st)
        self._run("k2_project_gather_meanvar", L.gpnerf_k2_project_gather_meanvar, ptr(self.images_rgbx),
                  ptr(self.featmaps_cl), 0, ptr(self.valid), ptr(self.rays_o), ptr(self.rays_d), ptr(self.z_vals),
                  None, fr, n_pts_max, ptr(self.counters), ptr(self.rgb_feat), ptr(self.mask), ptr(self.meanvar), st)
        self._run("k3_density_mlp", L.gpnerf_k3_density_mlp, ptr(self.vol_feat), 0, ptr(self.meanvar),
                  ptr(self.mask), C.byref(self._weights), self.V, n_pts_max, ptr(self.counters), CNT_P1,
                  ptr(self.sigma), None, self.precision, st)

    def render_dense(self, frame, ray_o, ray_d, near, far, t_rand=None, neg_ray=False):
        """BaseRender.Renderer.render_rays semantics: every sample of the given
        rays goes through both heads (no occupancy / density compaction).
        ray_o/ray_d [R,3] (one camera per batch as in the dataset path: row 0 of
        ray_o is the shared origin), near/far [R].  Returns device tensors."""
        if self._weights is None:
            raise _lib.GpnerfError("set_weights() has not been called")
        L, st = self.lib, self._stream()
        self.upload_frame(frame)
        R = int(ray_d.shape[0])
        if R > self.max_rays:
            raise _lib.GpnerfError(f"{R} rays exceed the engine capacity {self.max_rays}")
        dev = self.device
        self.rays_o.copy_(_f32(ray_o, dev).reshape(-1, 3)[0], non_blocking=True)
        self.rays_d[: R * 3].copy_(_f32(ray_d, dev).reshape(-1), non_blocking=True)
        self.near[:R].copy_(_f32(near, dev).reshape(-1), non_blocking=True)
        self.far[:R].copy_(_f32(far, dev).reshape(-1), non_blocking=True)
        self.counters[CNT_RAYS:CNT_RAYS + 1].fill_(R)
        tr = None if t_rand is None else _f32(t_rand, dev).reshape(-1)
        n = R * self.S
        if self.bf16 and not self.use_records and (self.rgb_in is None or self.rgb_in.numel() < n * self.V * 3):
            self.rgb_in = torch.empty(self.max_pts * self.V * 3, dtype=torch.float32, device=dev)
        # dense render: every sample goes through both heads – the tile hand-off unless another one was asked for
        self._impl_now = "tiles" if self.color_impl_auto else self.color_impl
        tiles = self.bf16 and self._impl_now == "tiles"
        self._heads(frame, masks3d=None, t_rand=tr, n_rays_max=R, rgb_in=self.rgb_in if tiles else None)
        # colour head on every point (valid1 = NULL → all rows in order)
        if tiles:
            self._run("k3_color_tiles_tc", L.gpnerf_k3_color_tiles_tc, ptr(self.rec_tiles), None,
                      C.byref(self._weights), self.V, n, ptr(self.counters), CNT_P1, ptr(self.rgb), st)
        elif self.bf16 and not self.use_records:
            self._color(None, n, CNT_P1, frame, self.rgb_in)
        else:
            self._color(None, n, CNT_P1, frame)
        raw = torch.cat([self.rgb[: n * 3].view(n, 3), self.sigma[:n].view(n, 1)], 1).contiguous()
        if self.bf16 and not self.use_records:
            rgb_in = self.rgb_in[: n * self.V * 3].view(n, self.V, 3)
        elif self.bf16:     # per-view RGB sits in chunk 4 of each view block of the record
            rc = self.rec_bytes // 2
            recs = self.rec[: n * self.rec_bytes].view(torch.bfloat16).view(n, rc)
            cols = [(9 + 5 * v + 4) * 8 + c for v in range(self.V) for c in range(3)]
            rgb_in = recs[:, cols].float().view(n, self.V, 3).contiguous()
        else:
            rgb_in = self.rgb_feat[: n * self.V * 35].view(n, self.V, 35)[..., :3].contiguous()
        out = {k: torch.empty(s, dtype=torch.float32, device=dev) for k, s in
               (("rgb_map", (R, 3)), ("disp_map", (R, 1)), ("acc_map", (R, 1)), ("depth_map", (R, 1)),
                ("alpha", (R, self.S)), ("rgb_in_map", (R, self.V * 3)))}
        self._run("k5_raw2outputs", L.gpnerf_k5_raw2outputs, ptr(raw), ptr(self.z_vals), ptr(rgb_in), R, self.S,
                  self.V, int(neg_ray), ptr(out["rgb_map"]), ptr(out["disp_map"]), ptr(out["acc_map"]),
                  ptr(out["depth_map"]), ptr(out["alpha"]), ptr(out["rgb_in_map"]), st)
        out["z_vals"] = self.z_vals[:n].view(R, self.S).clone()
        out["raw"] = raw.view(R, self.S, 4)
        return out

    def note_counts(self, p1, p2):
        """Auto hand-off (GPNERF_COLOR_IMPL unset): pick the colour head for the next frames from the survivor
        ratio of a frame just rendered."""
        if self.bf16 and self.color_impl_auto and p1 > 0:
            self.color_impl = "tiles" if p2 >= self.tiles_min_survival * p1 else "gather"

    def survivor_list(self):
        """The progressive step's ordered survivor list (demo_render.py:312-317) of the last frame: fills
        self.valid1[:P2].  With the tile hand-off it is not needed to render and is produced here, on demand,
        from the flag words of the fused kernel (still in the workspace until the next frame)."""
        if getattr(self, "_valid1_pending", False):
            self._valid1_pending = False
            self._run("k4_compact_alpha_fused", self.lib.gpnerf_k4_compact_alpha, None, self.max_pts,
                      ptr(self.counters), ptr(self.alpha), ptr(self.valid1), ptr(self.workspace), self._stream())
        return self.valid1

    def read_counters(self):
        """One device→host sync: (n_pix, n_rays, P1, P2).  Also brings valid1 up to date (survivor_list)."""
        self.survivor_list()
        c = self.counters.cpu().tolist()
        self.note_counts(c[CNT_P1], c[CNT_P2])
        return {"n_pix": c[CNT_PIX], "n_rays": c[CNT_RAYS], "P1": c[CNT_P1], "P2": c[CNT_P2]}
