"""Mirror of the reference's renderers behind the same plugin API.

``build_render(cfg) -> Renderer`` and ``Renderer.render(batch) -> dict`` as in
libs/renders/BaseRender.py:211-274,367-403 (dense training/validation path,
``progressive=False``) and libs/renders/demo_render.py:429-498,635-671
(progressive inference path, ``progressive=True``).  The hot path runs in
libgpnerf_b200.so through :class:`gpnerf_b200.engine.Engine`.

Upstream products.  The image encoder and the SMPL-code attention + sparse-conv
volume encoder are outside the hot path.  ``render`` takes their
outputs from ``batch['featmaps']`` / ``batch['levels']`` when present (synthetic
benchmarks, tests); otherwise it runs ``self.encoder`` and the reference's own
producer modules held by ``self.nerfhead.sigmahead`` (available when this file
is used inside the reference tree with spconv installed).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import PREC_BF16, PREC_FP32
from .engine import Engine


class Projector:
    """BaseRender.py:278-363 / demo_render.py:501-632 on the B200 kernels."""

    def __init__(self, device, neg_ray=False):
        self.device = device
        self.neg_ray = neg_ray

    @staticmethod
    def _frame(train_cameras, featmaps, neg_ray):
        cams = train_cameras.reshape(-1, 34).detach().cpu().float()
        V = cams.shape[0]
        f = _lib.Frame()
        f.n_views = V
        KE = cams[:, 2:18].reshape(V, 4, 4).bmm(cams[:, 18:].reshape(V, 4, 4))
        for v in range(V):
            f.src_KE[v][:] = KE[v].flatten().tolist()
        f.src_h, f.src_w = int(cams[0, 0]), int(cams[0, 1])
        f.feat_h, f.feat_w = int(featmaps.shape[-2]), int(featmaps.shape[-1])
        f.neg_ray = int(bool(neg_ray))
        f.n_samples = 1
        return f

    def compute(self, xyz, train_imgs, train_cameras, featmaps):
        """xyz [R,S,3]; train_imgs [1,V,3,H,W] already in [0,1]; train_cameras
        [1,V,34]; featmaps [V,C,h,w] → rgb_feat [R,S,V,35], mask [R,S,V,1]."""
        assert train_imgs.shape[0] == 1 and train_cameras.shape[0] == 1   # BaseRender.py:336
        f = self._frame(train_cameras, featmaps, self.neg_ray)
        rgbx = ops.images_to_rgbx(train_imgs[0], unnormalize=False)
        fm = ops.featmaps_to_channels_last(featmaps)
        sh = xyz.shape[:2]
        rgb_feat, mask, _ = ops.project_gather_meanvar(rgbx, fm, f, xyz.reshape(-1, 3))
        return rgb_feat.view(*sh, f.n_views, 35), mask.view(*sh, f.n_views, 1)

    def compute_smpl(self, smpl_xyz, train_cameras, featmaps):
        """demo_render.py:612-632 → [1,6890,V,32]"""
        f = self._frame(train_cameras, featmaps, self.neg_ray)
        dummy = torch.zeros(f.n_views * f.src_h * f.src_w * 4, device=featmaps.device)
        fm = ops.featmaps_to_channels_last(featmaps)
        rgb_feat, _, _ = ops.project_gather_meanvar(dummy, fm, f, smpl_xyz.reshape(-1, 3))
        return rgb_feat.view(*smpl_xyz.shape[:2], f.n_views, 35)[..., 3:]      # a view: K8 takes the strides


# small per-frame constants that end up in gpnerf_frame_t (engine.frame_from_batch) or in host-side set-up
_HOST_KEYS = ("Rh", "R", "Th", "bounds", "out_sh", "src_poses", "src_Ks", "target_pose", "target_K", "target_K_inv")


class Renderer(nn.Module):
    def __init__(self, encoder, nerfhead, is_train=True, neg_ray_train=False, neg_ray_val=False, n_rays=1024,
                 n_samples=64, voxel_size=(0.005, 0.005, 0.005), chunk=64, mesh_th=-1, progressive=False,
                 precision=PREC_FP32, t_min=0.0, rank=0, world=1, tile_px=64, use_cuda_graph=True,
                 shard="tiles", collective="peer"):
        super().__init__()
        self.encoder = encoder
        self.nerfhead = nerfhead
        self.is_train = is_train
        self.neg_ray_train = neg_ray_train
        self.neg_ray_val = neg_ray_val
        self.n_rays = n_rays
        self.n_samples = n_samples
        self.voxel_size = np.array(voxel_size)
        self.chunk = chunk            # kept for API compatibility; 180 GB of HBM needs no ray chunking
        self.mesh_th = mesh_th
        self.progressive = progressive
        self.precision = precision
        self.t_min = t_min
        # Several GPUs of one box (progressive path).  shard = "tiles": one frame, its pixel tiles dealt over
        # the ranks, every rank ends up with the full image; "frames": every rank renders its own batch
        # (a sweep of novel views) and rank 0 additionally receives all images (slot r = rank r).
        # collective = "peer": K5 writes the tiles into the peers' images over NVLink (peer.PeerExchange);
        # "nccl": one all_gather of tile buffers (shard.gather_frame; also the gloo path of the CPU tests).
        if shard not in ("tiles", "frames") or collective not in ("peer", "nccl"):
            raise _lib.GpnerfError("shard must be 'tiles'|'frames', collective 'peer'|'nccl'")
        self.rank, self.world, self.tile_px = rank, world, tile_px
        self.shard, self.collective = shard, collective
        self.use_cuda_graph = use_cuda_graph
        self._engine = None

    # ------------------------------------------------------------------ engine
    def engine_for(self, H, W, V, device, max_rays=None):
        e = self._engine
        key = (H, W, V, str(device), max_rays)
        if e is None or e._key != key:
            tiles = self.shard == "tiles"
            e = Engine(H, W, self.n_samples, V, device=device, precision=self.precision,
                       rank=self.rank if tiles else 0, world=self.world if tiles else 1, tile_px=self.tile_px,
                       t_min=self.t_min, max_rays=max_rays, voxel_size=tuple(float(v) for v in self.voxel_size))
            e._key = key
            self._engine = e
            if self.world > 1 and self.collective == "peer" and self.progressive:
                from .peer import PeerExchange
                e.attach_exchange(PeerExchange(H, W, device, self.rank, self.world, mode=self.shard))
        return e

    @staticmethod
    def _dist_ready():
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()

    def _sync_weights(self, eng):
        """(Re)pack the head weights for the kernels when they changed."""
        plist = getattr(self, "_hot_params", None)
        if plist is None:          # the parameters the kernels read; walking the module tree every frame costs 0.4 ms
            plist = self._hot_params = [p for k, p in self.nerfhead.named_parameters()
                                        if k.startswith("rgbhead.") or k.startswith("sigmahead.out_geometry_fc")] \
                or list(self.nerfhead.parameters())
        ver = tuple((p.data_ptr(), p._version) for p in plist)
        if getattr(eng, "_weights_ver", None) != ver:
            eng.set_weights(self.nerfhead.hot_path_state())
            eng._weights_ver = ver

    def _upstream(self, batch):
        """featmaps and dense levels: from the batch, or from the reference's
        producer modules (demo_render.py:103-157, 442)."""
        src_imgs = batch["src_imgs"]
        produce = "levels" not in batch and "levels_sparse" not in batch
        if produce:
            # host-side constants first: reading them back after the encoder has been queued would wait for it
            cams = self._pack_cameras(batch, src_imgs.shape[-2:], "cpu")
            out_sh = [int(v) for v in torch.as_tensor(batch["out_sh"]).reshape(-1, 3).max(0)[0].tolist()]
        if "featmaps" in batch:
            featmaps = batch["featmaps"]
        else:
            if self.encoder is None:
                raise _lib.GpnerfError("no encoder and no batch['featmaps']")
            featmaps = self.encoder(src_imgs.squeeze(0))
        if "levels" in batch:
            return featmaps, batch["levels"]
        if "levels_sparse" in batch:          # (features, indices) per level + batch['level_dims']: no dense volume at all
            return featmaps, None
        # the reference's producers (demo_render.py:103-157, trainhead.py:44-54) on K2/K8/K7: project the SMPL
        # vertices, attend, run the sparse-conv pyramid; its rows go to the renderer without a dense volume
        sh = self.nerfhead.sigmahead
        device = featmaps.device
        xyz = batch["feature"][..., :3].to(device).float()
        R, Th = batch["Rh"].to(device).float(), batch["Th"].to(device).float()
        smpl_xyz = torch.bmm(xyz, R.transpose(1, 2)) + Th
        feats = Projector(device).compute_smpl(smpl_xyz, cams, featmaps)
        rows, dims, n_dev = sh.encode_geometry(feats, batch["coord"].reshape(-1, batch["coord"].shape[-1]).to(device), out_sh)
        batch["levels_sparse"], batch["level_dims"], batch["levels_sparse_rows"] = rows, dims, n_dev
        return featmaps, None

    @staticmethod
    def _pack_cameras(batch, img_size, device):
        """src_cameras [1,V,34] – BaseRender.py:233-247"""
        src_poses, src_Ks = batch["src_poses"], batch["src_Ks"]
        V = src_poses.shape[1]
        Eh = torch.eye(4, device=device).repeat(1, V, 1, 1)
        Eh[:, :, :3, :4] = src_poses
        Kh = torch.eye(4, device=device).repeat(1, V, 1, 1)
        Kh[:, :, :3, :3] = src_Ks
        cams = torch.ones((1, V, 34), device=device)
        cams[:, :, 0] = img_size[0]
        cams[:, :, 1] = img_size[1]
        cams[:, :, 2:18] = Kh.reshape(1, V, -1)
        cams[:, :, -16:] = Eh.reshape(1, V, -1)
        return cams

    def _neg_ray(self, batch):
        """BaseRender.py:165-168"""
        if "body_msk" in batch and batch["body_msk"].shape[-1] > self.n_rays:
            return self.neg_ray_val
        return self.neg_ray_train

    # ------------------------------------------------------------------ render
    def render(self, batch):
        if getattr(self.nerfhead, "use_rgbhead", True) is False:
            return self.render_mesh(batch)
        return self.render_progressive(batch) if self.progressive else self.render_dense(batch)

    @torch.no_grad()
    def render_progressive(self, batch):
        """demo_render.Renderer.render: returns numpy rgb_map [R,3], pred_img
        [H,W,3] (float64), mask_at_box [H*W] bool, time_slots, etime, rtime."""
        device = batch["src_imgs"].device
        t_host0 = time.perf_counter()
        # etime / rtime (demo_render.py:98-101, 442-447) from stream events instead of two device-wide
        # synchronisations: the host keeps queueing while the producers run
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        # the frame constants are read back now, while the stream is idle: read after the producers have been
        # queued, each of these tiny copies would wait for the encoder and the pyramid
        hostc = {k: batch[k].detach().cpu() for k in _HOST_KEYS if k in batch and torch.is_tensor(batch[k]) and batch[k].is_cuda}
        batch = {**batch, **hostc}               # (a copy: the caller's dict is never modified)
        t_host = time.perf_counter() - t_host0
        ev[0].record(torch.cuda.current_stream(device))
        featmaps, levels = self._upstream(batch)
        ev[1].record(torch.cuda.current_stream(device))
        H, W = batch["src_imgs"].shape[-2:]
        V = batch["src_imgs"].shape[1]
        eng = self.engine_for(int(H), int(W), int(V), device)
        self._sync_weights(eng)
        detail = bool(getattr(self, "time_slots_detail", False))
        eng.timing = detail
        if detail:
            eng.stage_events = {}
        use_graph = self.use_cuda_graph and not detail
        if levels is None:
            # sparse levels: scatter the active rows (no dense volume, no K0 transposition), then K1…K5
            eng.upload_products_sparse(batch["levels_sparse"], batch["level_dims"], featmaps, batch["src_imgs"],
                                       n_rows_dev=batch.get("levels_sparse_rows"))
            frame = eng.make_frame(batch, neg_ray=self._neg_ray(batch))
            if use_graph:
                if getattr(self, "_frame_pinned", None) is None:
                    import ctypes as C
                    from ._lib import Frame
                    self._frame_pinned = torch.empty(C.sizeof(Frame), dtype=torch.uint8).pin_memory()
                eng.run_progressive_graphed(frame, with_k0=False, frame_src=self._frame_pinned)
            else:
                eng.render_progressive(frame)
        elif use_graph:
            # inputs land in static device buffers; the whole frame is one graph launch
            eng.copy_into_static_inputs(levels, featmaps, batch["src_imgs"],
                                        sharded_upload=self.world > 1 and self.shard == "tiles" and self._dist_ready())
            if eng.level_dims is None:
                eng.upload_products(*eng._static_inputs)      # first frame: learn the shapes
            frame = eng.make_frame(batch, neg_ray=self._neg_ray(batch))
            eng.run_progressive_graphed(frame)
        else:
            eng.upload_products(levels, featmaps, batch["src_imgs"])
            frame = eng.make_frame(batch, neg_ray=self._neg_ray(batch))
            eng.render_progressive(frame)
        ev[2].record(torch.cuda.current_stream(device))
        img_d, hit_d = eng.result_image(), eng.result_hit_mask()
        tiled = self.world > 1 and self.shard == "tiles"
        if tiled and eng.exchange is None:
            # the frame's only collective: one all_gather of this rank's pixel tiles (RGB + hit flag)
            import torch.distributed as dist
            from . import shard
            if not dist.is_initialized():
                raise _lib.GpnerfError("world > 1 needs an initialised torch.distributed process group")
            both = torch.cat([img_d, hit_d.view(-1, 1).float()], 1)
            both = shard.gather_frame(both, int(W), self.tile_px)
            img_d, hit_d = both[:, :3].contiguous(), both[:, 3] > 0.5
        # the frame's single host sync: image, hit mask and counters travel together through pinned staging (one
        # event wait; a pageable `.cpu()` per tensor costs three synchronous copies at a third of the bandwidth)
        pin = getattr(self, "_out_pinned", None)
        n_px = int(H) * int(W)
        if pin is None or pin[0].numel() != n_px * 3:
            pin = self._out_pinned = (torch.empty(n_px * 3, dtype=torch.float32).pin_memory(),
                                      torch.empty(n_px, dtype=torch.uint8).pin_memory(),
                                      torch.empty(_lib.N_COUNTERS, dtype=torch.int32).pin_memory(), torch.cuda.Event())
        pin[0].copy_(img_d.reshape(-1), non_blocking=True)
        pin[1].copy_(hit_d.reshape(-1).to(torch.uint8), non_blocking=True)
        pin[2].copy_(eng.counters, non_blocking=True)
        pin[3].record(torch.cuda.current_stream(device))
        pin[3].synchronize()
        c = pin[2].tolist()
        cnt = {"n_pix": c[_lib.CNT_PIX], "n_rays": c[_lib.CNT_RAYS], "P1": c[_lib.CNT_P1], "P2": c[_lib.CNT_P2]}
        eng.note_counts(cnt["P1"], cnt["P2"])          # auto hand-off: the colour head of the next frames
        img32 = pin[0].numpy().reshape(-1, 3)
        mask_at_box = pin[1].numpy().astype(bool)
        rgb_map = img32[np.flatnonzero(mask_at_box)]             # ascending pixel order = the reference's ray order
        pred_img = img32.reshape(H, W, 3).astype(np.float64)     # (fresh arrays: the staging buffers are reused)
        etime, rtime = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
        return {"rgb_map": rgb_map, "pred_img": pred_img, "mask_at_box": mask_at_box,
                "time_slots": self._time_slots(eng, etime, rtime, t_host), "etime": etime, "rtime": rtime, "counts": cnt}

    def _time_slots(self, eng, etime, rtime, t_host):
        """The ten keys of demo_render.render_rays' `time_slots` (demo_render.py:97-357), in seconds.  The reference
        brackets each stage with two device-wide synchronisations; here a frame is one CUDA-graph launch, so by
        default only what stream events can separate without stalling the pipeline is non-zero: `bc_time` (host set-up),
        `sp_encode` (all upstream producers: encoder, SMPL gather, attention, pyramid) and `bc_render` (K1…K5).  With
        `renderer.time_slots_detail = True` the frame is issued kernel by kernel with events around every stage and the
        path's own keys are filled in: `bf_sigma` (layout, pixel mask, rays, sampling, occupancy compaction),
        `sigma_f` (gathers + density head), `bf_rgb` (α + second compaction), `rgb_f` (colour head), `bc_render`
        (compositing)."""
        ts = {k: 0.0 for k in ("bc_time", "sigma_c", "bc_attn", "sigma_attn", "sp_encode", "bf_sigma", "sigma_f",
                               "bf_rgb", "rgb_f", "bc_render")}
        ts["bc_time"], ts["sp_encode"], ts["bc_render"] = t_host, etime, rtime
        if getattr(self, "time_slots_detail", False) and eng.stage_events:
            st = eng.stage_times_ms()
            g = lambda *keys: sum(st.get(k, 0.0) for k in keys) * 1e-3        # noqa: E731
            ts["bf_sigma"] = g("k0_products_to_f16", "k0_sparse_to_f16", "k0_level_to_channels_last",
                               "k0_featmaps_to_channels_last", "k0_images_to_rgbx", "k0_build_masks3d", "k1_voxel_pixel_mask",
                               "k1_rays_bbox", "k2_occupancy_compact")
            ts["sigma_f"] = g("k23_gather_density_tc", "k2_gather_volume", "k2_project_gather_meanvar", "k3_density_mlp")
            ts["bf_rgb"] = g("k4_compact_alpha", "k4_compact_alpha_fused")
            ts["rgb_f"] = g("k3_color_gather_tc", "k3_color_tiles_tc", "k3_color_mlp_records", "k3_color_mlp")
            ts["bc_render"] = g("k5_composite", "peer_wait")
        return ts

    @torch.no_grad()
    def render_stream(self, batches, depth=3):
        """Progressive renders of a sequence of batches (a sweep, a video), as a
        generator yielding the same dicts as ``render`` in order.  The
        host→device upload of batch i+1, i+2 (134 MB per 512² frame: the slowest
        leg of an end-to-end frame) runs on a copy stream into one of `depth`
        staging sets while batch i renders, and the image comes back through
        pinned buffers, so a steady stream is bound by PCIe or by the kernels,
        whichever is slower – not by their sum.  Upstream products come with the
        batch (``levels`` | ``levels_sparse``, ``featmaps``) or, when missing, are
        produced on the device by the encoder and the sigma head's K8/K7 mirrors."""
        import collections
        import ctypes as C
        from ._lib import Frame
        if not self.progressive:
            raise _lib.GpnerfError("render_stream is the progressive (inference) path")
        if self.world > 1 and self.shard == "tiles":
            for b in batches:             # a tile-sharded frame needs all ranks in lock step: no queueing ahead
                yield self.render_progressive(b)
            return
        from concurrent.futures import ThreadPoolExecutor
        st = None
        pending = collections.deque()
        # the host-side tail of a frame (wait for its event, fp32 → float64 image, boolean mask, the
        # compact rgb_map) runs on two worker threads – numpy releases the GIL – while the main thread queues
        # the next frames; results are still yielded in order
        # (the pool and the staging sets below – pinned host buffers, device buffers, a copy stream, events – are
        # kept on the Renderer between calls: allocating them cost 5-20 ms per call, more than ten frames)
        pool = getattr(self, "_stream_pool", None)
        if pool is None:
            pool = self._stream_pool = ThreadPoolExecutor(max_workers=2)
        cache = getattr(self, "_stream_cache", None)
        if cache is None:
            cache = self._stream_cache = {}

        def host_tail(slot, H, W, t_start):
            st["done"][slot].synchronize()
            cnt = dict(zip(("n_pix", "n_rays", "P1", "P2"), st["cnt"][slot][:4].tolist()))
            eng.note_counts(cnt["P1"], cnt["P2"])      # auto hand-off: the colour head of the frames queued from now on
            img32 = st["img"][slot].numpy().reshape(-1, 3)
            mask_at_box = st["hit"][slot].numpy().astype(bool)
            rgb_map = img32[np.flatnonzero(mask_at_box)]                  # ascending pixel order, fp32
            pred_img = img32.reshape(H, W, 3).astype(np.float64)
            rtime = time.time() - t_start
            return {"rgb_map": rgb_map, "pred_img": pred_img, "mask_at_box": mask_at_box,
                    "time_slots": {"bc_render": rtime}, "etime": 0.0, "rtime": rtime, "counts": cnt}

        def finalize(item):
            return item.result()

        for i, batch in enumerate(batches):
            sparse = "levels_sparse" in batch
            produce = "levels" not in batch and not sparse      # pyramid (and encoder) run here, on the device
            need_fm = "featmaps" not in batch
            if need_fm and self.encoder is None:
                raise _lib.GpnerfError("render_stream: no batch['featmaps'] and no encoder")
            src = batch["src_imgs"]
            H, W = int(src.shape[-2]), int(src.shape[-1])
            V = int(src.shape[1])
            device = self._stream_device(batch)
            eng = self.engine_for(H, W, V, device)
            self._sync_weights(eng)
            if st is None:
                key = (H, W, V, str(device), depth, sparse, produce, need_fm,
                       None if (sparse or produce) else tuple(tuple(t.shape) for t in batch["levels"]),
                       None if need_fm else tuple(batch["featmaps"].shape))
                st = cache.get(key)
                if st is not None:          # (a previous stream that was abandoned half way may still have work queued)
                    st["copy"].synchronize()
                    torch.cuda.current_stream(device).synchronize()
            if st is None:
                mk = lambda t: torch.empty(t.shape, dtype=torch.float32, device=device)     # noqa: E731
                im0 = src[0] if src.dim() == 5 else src
                st = cache[key] = {
                    "copy": torch.cuda.Stream(device),
                    "stage": [([] if (sparse or produce) else [mk(t) for t in batch["levels"]],
                               None if need_fm else mk(batch["featmaps"]), mk(im0)) for _ in range(depth)],
                    "sparse": [[] for _ in range(depth)], "small": {},
                    "copied": [torch.cuda.Event() for _ in range(depth)],
                    "free": [torch.cuda.Event() for _ in range(depth)],
                    "done": [torch.cuda.Event() for _ in range(depth)],
                    "frame": [torch.empty(C.sizeof(Frame), dtype=torch.uint8).pin_memory() for _ in range(depth)],
                    "img": [torch.empty(H * W * 3, dtype=torch.float32).pin_memory() for _ in range(depth)],
                    "hit": [torch.empty(H * W, dtype=torch.uint8).pin_memory() for _ in range(depth)],
                    "cnt": [torch.empty(8, dtype=torch.int32).pin_memory() for _ in range(depth)],
                }
            if len(pending) == depth:
                yield finalize(pending.popleft())       # the slot's pinned buffers are free again after this
            slot = i % depth
            t_start = time.time()
            main = torch.cuda.current_stream(device)
            lv_d, fm_d, im_d = st["stage"][slot]
            if i >= depth:
                st["copied"][slot].synchronize()                   # the slot's pinned staging has been read (long ago)
            with torch.cuda.stream(st["copy"]):
                if i >= depth:
                    st["copy"].wait_event(st["free"][slot])        # K0 of the previous tenant has read the set
                if sparse:
                    # row counts change from frame to frame: per-slot staging buffers with spare capacity (fresh
                    # tensors allocated on the copy stream would make the caching allocator fall back to
                    # cudaMalloc – a device-wide synchronisation – whenever the host runs ahead of the GPU)
                    bufs = st["sparse"][slot]
                    lv_s = []
                    for l, (f, ix) in enumerate(batch["levels_sparse"]):
                        n = int(f.shape[0])
                        if l >= len(bufs) or bufs[l][0].shape[0] < n or bufs[l][1].shape[1] != ix.shape[1]:
                            cap = n + n // 4 + 1024
                            pair = (torch.empty(cap, f.shape[1], dtype=torch.float32, device=device),
                                    torch.empty(cap, ix.shape[1], dtype=torch.int32, device=device))
                            if l < len(bufs):
                                bufs[l] = pair
                            else:
                                bufs.append(pair)
                        fb, ib = bufs[l][0][:n], bufs[l][1][:n]
                        fb.copy_(f, non_blocking=True)
                        ib.copy_(ix, non_blocking=True)
                        lv_s.append((fb, ib))
                elif not produce:
                    if not lv_d:      # first dense batch of a stream that started with sparse ones
                        lv_d.extend(torch.empty(t.shape, dtype=torch.float32, device=device) for t in batch["levels"])
                    for d, s_ in zip(lv_d, batch["levels"]):
                        d.copy_(s_, non_blocking=True)
                if not need_fm:
                    if fm_d is None:
                        fm_d = torch.empty(batch["featmaps"].shape, dtype=torch.float32, device=device)
                        st["stage"][slot] = (lv_d, fm_d, im_d)
                    fm_d.copy_(batch["featmaps"], non_blocking=True)
                im_d.copy_(src[0] if src.dim() == 5 else src, non_blocking=True)
                small = {}
                if produce:
                    # the SMPL fit goes through pinned staging too: a pageable host→device copy on the main stream
                    # would first wait for everything queued there, i.e. for the previous frame
                    for k in ("feature", "coord", "Rh", "Th"):
                        t = batch[k]
                        if t.is_cuda:
                            continue
                        key = (slot, k)
                        if key not in st["small"] or st["small"][key][0].shape != t.shape or st["small"][key][0].dtype != t.dtype:
                            st["small"][key] = (torch.empty(t.shape, dtype=t.dtype).pin_memory(),
                                                torch.empty(t.shape, dtype=t.dtype, device=device))
                        pin, dbuf = st["small"][key]
                        pin.copy_(t)
                        dbuf.copy_(pin, non_blocking=True)
                        small[k] = dbuf
                st["copied"][slot].record(st["copy"])
            main.wait_event(st["copied"][slot])
            if need_fm or produce:
                # the producers (f2, f1) on the main stream, fed from the staging set; their results live in
                # module-owned buffers that the sparse upload below consumes before the next frame overwrites them
                b2 = {**batch, **small, "src_imgs": im_d.unsqueeze(0)}
                if not need_fm:
                    b2["featmaps"] = fm_d
                fm_d, _lv = self._upstream(b2)
                if produce:
                    eng.upload_products_sparse(b2["levels_sparse"], b2["level_dims"], fm_d, im_d,
                                               n_rows_dev=b2["levels_sparse_rows"])
                elif sparse:
                    eng.upload_products_sparse(lv_s, batch["level_dims"], fm_d, im_d)
                else:
                    eng.upload_products(lv_d, fm_d, im_d)
            elif sparse:
                eng.upload_products_sparse(lv_s, batch["level_dims"], fm_d, im_d)
            else:
                eng.upload_products(lv_d, fm_d, im_d)              # K0: staging set → gather layouts
            st["free"][slot].record(main)
            frame = eng.make_frame(batch, neg_ray=self._neg_ray(batch))
            if self.use_cuda_graph:
                eng.run_progressive_graphed(frame, with_k0=False, frame_src=st["frame"][slot])
            else:
                C.memmove(st["frame"][slot].data_ptr(), C.addressof(frame), C.sizeof(Frame))
                eng.frame_dev.copy_(st["frame"][slot], non_blocking=True)
                eng.render_progressive(frame, upload=False)
            st["img"][slot].copy_(eng.result_image().reshape(-1), non_blocking=True)
            st["hit"][slot].copy_(eng.result_hit_mask(), non_blocking=True)
            st["cnt"][slot].copy_(eng.counters, non_blocking=True)
            st["done"][slot].record(main)
            pending.append(pool.submit(host_tail, slot, H, W, t_start))
        while pending:
            yield finalize(pending.popleft())

    def _stream_device(self, batch):
        src = batch["src_imgs"]
        if src.is_cuda:
            return src.device
        return torch.device("cuda", torch.cuda.current_device())

    @torch.no_grad()
    def render_mesh(self, batch):
        """The mesh branch (`cfg.head.rgb.use_rgbhead False`; BaseRender.py:255-272,
        demo_render.py:249-268, 366-376): σ of the density head at the grid points
        `batch['pts']` [1,X,Y,Z,3] selected by `batch['inside']` [1,X,Y,Z], α = 1 - exp(-σ)
        scattered into the grid, zero-padded by 10 → ret['cube'] (float64 numpy, as
        the reference builds it); ret['vertices'] / ret['triangles'] / ret['mesh'] from
        PyMCubes + trimesh when importable (as the reference does), else from this
        package's marching tetrahedra (isosurface.py).  Gathers and the head run in
        the library's kernels at explicit points (no rays, no compaction)."""
        device = batch["src_imgs"].device
        batch = dict(batch)                     # _upstream may add the produced rows: never to the caller's dict
        featmaps, levels = self._upstream(batch)
        src = batch["src_imgs"]
        H, W = int(src.shape[-2]), int(src.shape[-1])
        V = int(src.shape[1])
        from .engine import frame_from_batch
        fm = featmaps.to(device)
        if levels is None:                      # sparse rows (from the batch or from the sigma head's own producers)
            rows = [(f.to(device), i.to(device)) for f, i in batch["levels_sparse"]]
            levels_cl, dims = ops.sparse_levels_to_channels_last(rows, batch["level_dims"], batch.get("levels_sparse_rows"))
        else:
            lv = [t.to(device) for t in levels]
            dims = [tuple(int(v) for v in t.shape[-3:]) for t in lv]
            levels_cl = [ops.level_to_channels_last(t)[0] for t in lv]
        frame = frame_from_batch(batch, H=H, W=W, n_views=V, n_samples=1, level_dims=dims, src_hw=(H, W),
                                 feat_hw=tuple(int(v) for v in fm.shape[-2:]),
                                 voxel_size=tuple(float(v) for v in self.voxel_size), neg_ray=self._neg_ray(batch))
        hw, _keep = ops.pack_head_weights(self.nerfhead.hot_path_state(), device, V,
                                          tensor_core_image=self.precision != PREC_FP32)
        fm_cl = ops.featmaps_to_channels_last(fm)
        rgbx = ops.images_to_rgbx(src[0] if src.dim() == 5 else src, unnormalize=True)
        inside = batch["inside"][0].bool()
        pts = batch["pts"][0][inside].reshape(-1, 3).to(device=device, dtype=torch.float32).contiguous()
        sigma = torch.empty(pts.shape[0], dtype=torch.float32, device=device)
        step = 1 << 22                       # explicit points need fp32 rows: 800 B per point and chunk
        for p0 in range(0, pts.shape[0], step):
            p = pts[p0:p0 + step]
            vol = ops.gather_volume(levels_cl, frame, p)
            _rgb_feat, mask, meanvar = ops.project_gather_meanvar(rgbx, fm_cl, frame, p)
            sigma[p0:p0 + step] = ops.density_mlp(vol, meanvar, mask, hw, self.precision)
        alpha = ops.alpha_of_sigma(sigma)
        cube = np.zeros(tuple(inside.shape))
        cube[inside.cpu().numpy()] = alpha.cpu().numpy()
        cube = np.pad(cube, 10, mode="constant")
        ret = {"cube": cube}
        try:                                    # the reference's own path (BaseRender.py:270-272), when its packages exist
            import mcubes
            vertices, triangles = mcubes.marching_cubes(cube, self.mesh_th)
        except ImportError:                     # PyMCubes is not part of the reference tree: marching tetrahedra on the GPU
            from .isosurface import marching_tetrahedra
            vertices, triangles = marching_tetrahedra(torch.from_numpy(cube).to(device), float(self.mesh_th))
        ret["vertices"], ret["triangles"] = vertices, triangles
        try:
            import trimesh
            ret["mesh"] = trimesh.Trimesh(vertices, triangles)
        except ImportError:
            from types import SimpleNamespace
            ret["mesh"] = SimpleNamespace(vertices=vertices, faces=triangles)
        return ret

    def _wants_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def render_dense_train(self, batch):
        """BaseRender.Renderer.render under autograd (what BaseTrainer._forward calls, BaseTrainer.py:99-131): the
        same outputs as `render_dense`, connected to the module's parameters.  The hot path (sampling, gathers,
        both heads, raw2outputs; BaseRender.py:110-157) runs forward and backward in the K6 kernels
        (train.render_dense_autograd) with this module's head parameters as differentiable leaves; the upstream
        producers run in their training form (trainmode.py: batch-statistics BatchNorm, torch autograd ops on the
        same parameters), so gradients reach the encoder, the SMPL codes, the attention and the pyramid too."""
        from . import train, trainmode
        device = batch["src_imgs"].device
        src_imgs = batch["src_imgs"]
        H, W = (int(v) for v in src_imgs.shape[-2:])
        V = int(src_imgs.shape[1])
        if "featmaps" in batch:
            featmaps = batch["featmaps"].to(device)
        elif hasattr(self.encoder, "_run"):                       # this package's mirror: its training form
            featmaps = trainmode.encoder_forward(self.encoder, src_imgs.squeeze(0))
        else:                                                     # the reference's own torch module
            featmaps = self.encoder(src_imgs.squeeze(0))
        if "levels" in batch:
            levels = [t.to(device) for t in batch["levels"]]
        else:
            sh = self.nerfhead.sigmahead
            cams = self._pack_cameras(batch, src_imgs.shape[-2:], device)
            out_sh = [int(v) for v in torch.as_tensor(batch["out_sh"]).reshape(-1, 3).max(0)[0].tolist()]
            xyz = batch["feature"][..., :3].to(device).float()
            Rm, Th = batch["Rh"].to(device).float(), batch["Th"].to(device).float()
            smpl_xyz = torch.bmm(xyz, Rm.transpose(1, 2)) + Th
            feats = trainmode.smpl_features(smpl_xyz, cams, featmaps, self._neg_ray(batch)).flatten(0, 1)
            code = sh.c(torch.arange(sh.c.num_embeddings, device=device))                   # trainhead.py:48
            fused = trainmode.attention_forward(sh.xyzc_attn, code.unsqueeze(1), feats).squeeze(1)
            coord = batch["coord"].reshape(-1, batch["coord"].shape[-1]).to(device)
            rows, dims = trainmode.pyramid_forward(sh.xyzc_net, fused, coord, out_sh)
            levels = [trainmode.rows_to_dense(r, c, d) for (r, c), d in zip(rows, dims)]
        rays_o, rays_d = batch["ray_o"], batch["ray_d"]
        R = int(rays_o.shape[1])
        eng = getattr(self, "_train_engine", None)
        key = (H, W, V, str(device), R)
        if eng is None or eng._key != key:
            eng = Engine(H, W, self.n_samples, V, device=device, precision=PREC_FP32, max_rays=R, t_min=self.t_min,
                         voxel_size=tuple(float(v) for v in self.voxel_size))
            eng._key = key
            self._train_engine = eng
        from .engine import frame_from_batch
        neg = self._neg_ray(batch)
        host = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in batch.items() if k in _HOST_KEYS}
        frame = frame_from_batch(host, H=H, W=W, n_views=V, n_samples=self.n_samples,
                                 level_dims=[tuple(int(v) for v in t.shape[-3:]) for t in levels], src_hw=(H, W),
                                 feat_hw=tuple(int(v) for v in featmaps.shape[-2:]),
                                 voxel_size=tuple(float(v) for v in self.voxel_size), neg_ray=neg)
        frame.self_dev = eng.frame_dev.data_ptr()
        t_rand = torch.rand((1, R, self.n_samples)) if self.is_train else None            # BaseRender.py:40-47
        params = {k: p for k, p in self.nerfhead.named_parameters()
                  if k.startswith("rgbhead.") or k.startswith("sigmahead.out_geometry_fc")}
        rays = tuple(batch[k][0].to(device) for k in ("ray_o", "ray_d", "near", "far"))
        out = train.render_dense_autograd(eng, frame, rays, levels, featmaps, src_imgs.to(device), params, t_rand=t_rand,
                                          neg_ray=neg, precision=getattr(self, "train_precision", train.PREC_TRAIN_TF32))
        keys = ("rgb_map", "disp_map", "acc_map", "depth_map", "alpha", "z_vals", "rgb_in_map")
        return {k: out[k].view(1, R, -1) for k in keys}

    def render_dense(self, batch):
        """BaseRender.Renderer.render: rgb_map [1,R,3], disp/acc/depth [1,R,1],
        alpha (=weights) [1,R,S], z_vals [1,R,S], rgb_in_map [1,R,3V].  With gradients
        enabled and trainable parameters the differentiable route is taken
        (`render_dense_train`); otherwise the forward-only inference kernels."""
        if self._wants_grad():
            return self.render_dense_train(batch)
        with torch.no_grad():
            return self._render_dense_infer(batch)

    def _render_dense_infer(self, batch):
        device = batch["src_imgs"].device
        batch = dict(batch)                     # _upstream may add the produced rows: never to the caller's dict
        featmaps, levels = self._upstream(batch)
        H, W = batch["src_imgs"].shape[-2:]
        V = batch["src_imgs"].shape[1]
        rays_o, rays_d = batch["ray_o"], batch["ray_d"]
        R = int(rays_o.shape[1])
        eng = self.engine_for(int(H), int(W), int(V), device, max_rays=R)
        self._sync_weights(eng)
        if levels is None:                      # sparse rows (from the batch or from the sigma head's own producers)
            eng.upload_products_sparse(batch["levels_sparse"], batch["level_dims"], featmaps, batch["src_imgs"],
                                       n_rows_dev=batch.get("levels_sparse_rows"))
        else:
            eng.upload_products(levels, featmaps, batch["src_imgs"])
        neg = self._neg_ray(batch)
        frame = eng.make_frame(batch, neg_ray=neg)
        t_rand = None
        if self.is_train:          # BaseRender.py:40-47: jitter drawn on the CPU generator
            t_rand = torch.rand((1, R, self.n_samples))
        out = eng.render_dense(frame, rays_o[0], rays_d[0], batch["near"][0], batch["far"][0], t_rand, neg)
        keys = ("rgb_map", "disp_map", "acc_map", "depth_map", "alpha", "z_vals", "rgb_in_map")
        return {k: out[k].view(1, R, -1) for k in keys}


def build_render(cfg, progressive=False):
    """BaseRender.py:367-403 / demo_render.py:635-671.  `cfg.encoder.file` and
    `cfg.head.file` keep their plugin meaning; files that cannot be imported
    fall back to this package's ResUNet / NeRFHead mirrors."""
    from importlib import import_module as impm
    try:
        encoder = getattr(impm(cfg.encoder.file), "build_encoder")(cfg)
    except Exception:
        from .encoder import build_encoder            # the ResUNet mirror (row f2)
        encoder = build_encoder(cfg)
    try:
        nerfhead = getattr(impm(cfg.head.file), "build_head")(cfg)
        if not hasattr(nerfhead, "hot_path_state"):
            raise ImportError
    except Exception:
        from .nerfhead import build_head
        nerfhead = build_head(cfg)
    neg_ray_train = "thuman" in cfg.dataset.train.name
    neg_ray_val = "thuman" in cfg.dataset.test.name
    is_train = nerfhead.training or (encoder is not None and encoder.training)
    chunk = cfg.dataset.train.chunk if is_train else cfg.dataset.test.chunk
    mesh_th = 1.0 / cfg.test.mesh_th if cfg.head.rgb.use_rgbhead is False else -1
    # arithmetic of the inference kernels: `cfg.head.precision` ("fp32" | "bf16", or 0 | 1) when the config names
    # one, else the tensor-core path (bf16 MLPs, north_star).  Training (gradients enabled) always runs the K6 path.
    prec = getattr(cfg.head, "precision", None)
    if prec is None:
        prec = PREC_BF16
    elif isinstance(prec, str):
        prec = {"fp32": PREC_FP32, "bf16": PREC_BF16}[prec.lower()]
    return Renderer(encoder=encoder, nerfhead=nerfhead, is_train=False if progressive else is_train,
                    neg_ray_train=neg_ray_train, neg_ray_val=neg_ray_val, n_rays=cfg.train.n_rays,
                    n_samples=cfg.train.n_samples, voxel_size=cfg.dataset.voxel_size, chunk=chunk,
                    mesh_th=mesh_th, progressive=progressive, precision=int(prec))
