"""Tensor-level wrappers of the stand-alone C-ABI entry points – the operator
granularity of the reference's modules (Projector.compute,
SparseConvNet.forward's gather, fused_mean_variance, the head MLPs,
Renderer.raw2outputs).  Inputs/outputs are torch CUDA tensors; all arithmetic
happens in libgpnerf_b200.so.  No fallback: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import PREC_FP32, Frame, HeadWeights, check, ptr, ptr_array


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.GpnerfError("gpnerf_b200 ops take CUDA tensors only (no CPU fallback)")


def _c(t):
    return t.detach().to(torch.float32).contiguous()


def level_to_channels_last(level):
    """[1,32,D,H,W] → ([D*H*W*32], chan_sum [D*H*W])."""
    _need_cuda(level)
    lib = _lib.load()
    _, ch, D, H, W = level.shape
    assert ch == 32
    out = torch.empty(D * H * W * 32, dtype=torch.float32, device=level.device)
    cs = torch.empty(D * H * W, dtype=torch.float32, device=level.device)
    check(lib.gpnerf_k0_level_to_channels_last(ptr(_c(level)), D, H, W, 0, 0, ptr(out), ptr(cs), _stream(level.device)),
          "k0_level_to_channels_last")
    return out, cs


def sparse_levels_to_channels_last(levels_sparse, level_dims, n_rows_dev=None):
    """The pyramid's levels as active rows (features [N,32], indices [N,3|4] with (d,h,w) last; what a
    SparseConvTensor holds before `.dense()`, SparseConvNet.py:110) → the 4 fp32 channel-last volumes
    [D·H·W·32] that `gather_volume` reads (gpnerf_k0_sparse_to_f32: one scatter, no dense NCDHW tensor)."""
    lib = _lib.load()
    feats = [_c(f.to(torch.float32)) for f, _ in levels_sparse]
    idxs = [i.detach().to(torch.int32).contiguous() for _, i in levels_sparse]
    _need_cuda(*feats, *idxs)
    dev = feats[0].device
    dims = [tuple(int(v) for v in d) for d in level_dims]
    cols = int(idxs[0].shape[1])
    out = [torch.empty(d * h * w * 32, dtype=torch.float32, device=dev) for d, h, w in dims]
    sums = [torch.empty(d * h * w, dtype=torch.float32, device=dev) for d, h, w in dims]
    dims_c = ((C.c_int32 * 3) * 4)(*[(C.c_int32 * 3)(*d) for d in dims])
    n_rows = (C.c_int32 * 4)(*[int(f.shape[0]) for f in feats])
    nrd = None if n_rows_dev is None else _lib.ptr_array([t.view(1) for t in n_rows_dev])
    check(lib.gpnerf_k0_sparse_to_f32(_lib.ptr_array(feats), _lib.ptr_array(idxs), n_rows, nrd, cols, dims_c,
                                      _lib.ptr_array(out), _lib.ptr_array(sums), _stream(dev)), "k0_sparse_to_f32")
    return out, dims


def featmaps_to_channels_last(featmaps):
    _need_cuda(featmaps)
    lib = _lib.load()
    V, ch, h, w = featmaps.shape
    assert ch == 32
    out = torch.empty(V * h * w * 32, dtype=torch.float32, device=featmaps.device)
    check(lib.gpnerf_k0_featmaps_to_channels_last(ptr(_c(featmaps)), V, h, w, 0, 0, ptr(out), _stream(featmaps.device)),
          "k0_featmaps_to_channels_last")
    return out


def images_to_rgbx(src_imgs, unnormalize=True):
    """[V,3,H,W] → [V*H*W*4] (RGB, pad); x*0.5+0.5 on the way if `unnormalize`."""
    _need_cuda(src_imgs)
    lib = _lib.load()
    V, ch, H, W = src_imgs.shape
    assert ch == 3
    out = torch.empty(V * H * W * 4, dtype=torch.float32, device=src_imgs.device)
    check(lib.gpnerf_k0_images_to_rgbx(ptr(_c(src_imgs)), V, H, W, int(bool(unnormalize)), 0, ptr(out), _stream(src_imgs.device)),
          "k0_images_to_rgbx")
    return out


def gather_volume(levels_cl, frame: Frame, points, normalised=False):
    """SparseConvNet.forward's 4-level trilinear gather.  points [n,3] are
    world points, or normalised (x,y,z) grid coordinates when `normalised`."""
    _need_cuda(points, *levels_cl)
    lib = _lib.load()
    pts = _c(points).reshape(-1, 3)
    n = pts.shape[0]
    out = torch.empty((n, 128), dtype=torch.float32, device=pts.device)
    if n:
        check(lib.gpnerf_k2_gather_volume(ptr_array(levels_cl), 2 if normalised else 1, None, None, None, None,
                                          ptr(pts), C.byref(frame), n, None, ptr(out), _stream(pts.device)),
              "k2_gather_volume")
    return out


def project_gather_meanvar(images_rgbx, featmaps_cl, frame: Frame, points):
    """Projector.compute + fused_mean_variance on explicit world points [n,3].
    Returns rgb_feat [n,V,35], mask [n,V], meanvar [n,70]."""
    _need_cuda(points, images_rgbx, featmaps_cl)
    lib = _lib.load()
    pts = _c(points).reshape(-1, 3)
    n, V = pts.shape[0], frame.n_views
    dev = pts.device
    rgb_feat = torch.empty((n, V, 35), dtype=torch.float32, device=dev)
    mask = torch.empty((n, V), dtype=torch.float32, device=dev)
    meanvar = torch.empty((n, 70), dtype=torch.float32, device=dev)
    if n:
        check(lib.gpnerf_k2_project_gather_meanvar(ptr(images_rgbx), ptr(featmaps_cl), 1, None, None, None, None,
                                                   ptr(pts), C.byref(frame), n, None, ptr(rgb_feat), ptr(mask),
                                                   ptr(meanvar), _stream(dev)), "k2_project_gather_meanvar")
    return rgb_feat, mask, meanvar


def mean_variance(rgb_feat):
    """fused_mean_variance: [..., V, 35] → meanvar [n,70]."""
    _need_cuda(rgb_feat)
    lib = _lib.load()
    V = rgb_feat.shape[-2]
    x = _c(rgb_feat).reshape(-1, V, 35)
    n = x.shape[0]
    out = torch.empty((n, 70), dtype=torch.float32, device=x.device)
    if n:
        check(lib.gpnerf_k2_mean_variance(ptr(x), V, n, ptr(out), _stream(x.device)), "k2_mean_variance")
    return out


def pack_head_weights(state_dict, device, n_views=3, tensor_core_image=True):
    """HeadWeights struct (+ the tensors it points into) from reference-keyed
    parameters.  With `tensor_core_image` the bf16 UMMA operand image for the
    tcgen05 path is packed on the device as well."""
    from .engine import HEAD_KEYS

    def get(name):
        for pre in ("", "nerfhead."):
            if pre + name in state_dict:
                return state_dict[pre + name].detach().to(device=device, dtype=torch.float32).contiguous()
        raise KeyError(name)
    hw, keep = HeadWeights(), []

    def pair(name):
        w, b = get(name + ".weight"), get(name + ".bias")
        keep.extend([w, b])
        return w.data_ptr(), b.data_ptr()
    hw.geo_w, hw.geo_b = pair(HEAD_KEYS["geo"][0])
    for field in ("den", "base", "vis", "rgb"):
        for i, nme in enumerate(HEAD_KEYS[field]):
            wp, bp = pair(nme)
            getattr(hw, field + "_w")[i] = wp
            getattr(hw, field + "_b")[i] = bp
    dev = torch.device(device)
    if tensor_core_image and dev.type == "cuda" and 1 <= n_views <= 4 and keep[0].numel() == 64 * 128:
        lib = _lib.load()
        image = torch.empty(lib.gpnerf_k3_packed_weight_bytes(), dtype=torch.uint8, device=dev)
        check(lib.gpnerf_k3_pack_weights(C.byref(hw), n_views, ptr(image), _stream(dev)), "k3_pack_weights")
        hw.tc_image = image.data_ptr()
        keep.append(image)
    return hw, keep


def density_mlp(feat_in, meanvar, mask, weights: HeadWeights, precision=PREC_FP32, want_sigma_feat=False):
    """feat_in [n,128] (volume features) or [n,64] (sigma_feat); returns σ [n]
    (and sigma_feat [n,64])."""
    _need_cuda(feat_in, meanvar, mask)
    lib = _lib.load()
    x = _c(feat_in)
    n, k = x.shape
    assert k in (128, 64)
    V = mask.shape[-1]
    dev = x.device
    sigma = torch.empty(n, dtype=torch.float32, device=dev)
    sfeat = torch.empty((n, 64), dtype=torch.float32, device=dev) if want_sigma_feat else None
    if n:
        check(lib.gpnerf_k3_density_mlp(ptr(x), 0 if k == 128 else 1, ptr(_c(meanvar)), ptr(_c(mask).reshape(n, V)),
                                        C.byref(weights), V, n, None, 0, ptr(sigma), ptr(sfeat), precision,
                                        _stream(dev)), "k3_density_mlp")
    return (sigma, sfeat) if want_sigma_feat else sigma


def color_mlp(rgb_feat, meanvar, weights: HeadWeights, precision=PREC_FP32):
    """rgb_feat [n,V,35], meanvar [n,70] → rgb [n,3]."""
    _need_cuda(rgb_feat, meanvar)
    lib = _lib.load()
    x = _c(rgb_feat)
    n, V, _ = x.shape
    rgb = torch.empty((n, 3), dtype=torch.float32, device=x.device)
    if n:
        check(lib.gpnerf_k3_color_mlp(ptr(x), ptr(_c(meanvar)), None, C.byref(weights), V, n, None, 0, ptr(rgb),
                                      precision, _stream(x.device)), "k3_color_mlp")
    return rgb


def alpha_of_sigma(sigma):
    """α = 1 - exp(-σ) (demo_render.py:313, BaseRender.py:262) through K4's kernel."""
    _need_cuda(sigma)
    lib = _lib.load()
    s = _c(sigma).reshape(-1)
    n = s.numel()
    dev = s.device
    alpha = torch.empty(n, dtype=torch.float32, device=dev)
    if n:
        counters = torch.zeros(_lib.N_COUNTERS, dtype=torch.int32, device=dev)
        counters[_lib.CNT_P1] = n
        valid1 = torch.empty(n, dtype=torch.int32, device=dev)
        ws = torch.empty(int(lib.gpnerf_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        check(lib.gpnerf_k4_compact_alpha(ptr(s), n, ptr(counters), ptr(alpha), ptr(valid1), ptr(ws), _stream(dev)),
              "k4_compact_alpha")
    return alpha


def raw2outputs(raw, z_vals, rgb_in=None, neg=False):
    """Renderer.raw2outputs (+ rgb_in_map): raw [R,S,4], z_vals [R,S],
    rgb_in [R,S,V,3] → rgb_map, disp, acc, weights, depth, rgb_in_map."""
    _need_cuda(raw, z_vals, rgb_in)
    lib = _lib.load()
    raw, z = _c(raw), _c(z_vals)
    R, S, _ = raw.shape
    dev = raw.device
    V = 0 if rgb_in is None else rgb_in.shape[2]
    rgb_map = torch.empty((R, 3), dtype=torch.float32, device=dev)
    disp, acc, depth = (torch.empty(R, dtype=torch.float32, device=dev) for _ in range(3))
    weights = torch.empty((R, S), dtype=torch.float32, device=dev)
    rin = None if rgb_in is None else _c(rgb_in)
    rin_map = None if rgb_in is None else torch.empty((R, V * 3), dtype=torch.float32, device=dev)
    if R:
        check(lib.gpnerf_k5_raw2outputs(ptr(raw), ptr(z), ptr(rin), R, S, V, int(bool(neg)), ptr(rgb_map), ptr(disp),
                                        ptr(acc), ptr(depth), ptr(weights), ptr(rin_map), _stream(dev)),
              "k5_raw2outputs")
    return rgb_map, disp, acc, weights, depth, rin_map


class _Raw2OutputsFn(torch.autograd.Function):
    """Renderer.raw2outputs under autograd: both directions are library kernels
    (gpnerf_k5_raw2outputs / gpnerf_k5_raw2outputs_bwd).  Differentiable in
    `raw`; z_vals and rgb_in are treated as constants (they carry no parameters
    in the reference's training graph, BaseRender.py:110-157)."""

    @staticmethod
    def forward(ctx, raw, z_vals, rgb_in, neg):
        outs = raw2outputs(raw, z_vals, rgb_in, neg)
        ctx.save_for_backward(raw.detach(), z_vals.detach(), rgb_in.detach() if rgb_in is not None else None)
        ctx.neg = bool(neg)
        ctx.mark_non_differentiable(*[o for o in outs[5:] if o is not None and rgb_in is None])
        return outs if rgb_in is not None else outs[:5]

    @staticmethod
    def backward(ctx, g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_rin=None):
        raw, z, rin = ctx.saved_tensors
        lib = _lib.load()
        raw, z = _c(raw), _c(z)
        R, S, _ = raw.shape
        V = 0 if rin is None else rin.shape[2]
        d_raw = torch.empty_like(raw)

        def g(t):
            return None if t is None else _c(t)
        check(lib.gpnerf_k5_raw2outputs_bwd(ptr(raw), ptr(z), ptr(None if rin is None else _c(rin)), R, S, V,
                                            int(ctx.neg), ptr(g(g_rgb_map)), ptr(g(g_disp)), ptr(g(g_acc)),
                                            ptr(g(g_depth)), ptr(g(g_weights)), ptr(g(g_rin)), ptr(d_raw),
                                            _stream(raw.device)), "k5_raw2outputs_bwd")
        return d_raw, None, None, None


def raw2outputs_autograd(raw, z_vals, rgb_in=None, neg=False):
    """Differentiable raw2outputs: (rgb_map, disp, acc, weights, depth[, rgb_in_map])."""
    return _Raw2OutputsFn.apply(raw, z_vals, rgb_in, neg)


def dataset_rays(H, W, K, R, T, bounds, device):
    """The reference's CPU loader rays on the GPU:
    libs/datasets/data_utils.py get_rays (:47-63) + sample_ray's test split
    (:331-337) + get_near_far (:96-130) for every pixel of an H×W view.

    K [3,3], R [3,3], T [3] (or [3,1]) as the dataset holds them, `bounds` [2,3]
    fp32 (world-frame box).  The three tiny host-side steps are the very numpy
    calls the reference makes (two 3×3 inverses, one 3×3 · 3 product); everything
    per pixel runs in the kernel.  Returns device tensors ray_o [R,3], ray_d
    [R,3], near [R], far [R] (fp32, ascending pixel order) and mask_at_box [H*W]
    bool – the batch entries ZjumocapDataset.__getitem__ builds from them."""
    import numpy as np
    lib = _lib.load()
    dev = torch.device(device)
    K, R = np.asarray(K), np.asarray(R)
    T = np.asarray(T).reshape(3)
    R_inv = np.linalg.inv(R)                         # data_utils.py:49
    origin = (-R_inv @ T).ravel().astype(np.float64)  # :50-51
    K_inv = np.linalg.inv(K).astype(np.float64)      # :57
    b = (np.asarray(bounds) + np.array([-0.01, 0.01])[:, None]).astype(np.float64)   # :98
    n = H * W
    i32, f32 = torch.int32, torch.float32
    ray_pix = torch.empty(n, dtype=i32, device=dev)
    ray_o, ray_d = torch.empty(n * 3, dtype=f32, device=dev), torch.empty(n * 3, dtype=f32, device=dev)
    near, far = torch.empty(n, dtype=f32, device=dev), torch.empty(n, dtype=f32, device=dev)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=i32, device=dev)
    ws = torch.empty(int(lib.gpnerf_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    Ki, Ri, Oi, Bi = (np.ascontiguousarray(a, dtype=np.float64) for a in (K_inv, R_inv.astype(np.float64), origin, b))
    check(lib.gpnerf_k1_dataset_rays(Ki.ctypes.data_as(C.c_void_p), Ri.ctypes.data_as(C.c_void_p),
                                     Oi.ctypes.data_as(C.c_void_p), Bi.ctypes.data_as(C.c_void_p), H, W, ptr(ray_pix),
                                     ptr(ray_o), ptr(ray_d), ptr(near), ptr(far), ptr(mask), ptr(cnt), ptr(ws),
                                     _stream(dev)), "k1_dataset_rays")
    r = int(cnt.item())
    return (ray_o[: r * 3].view(r, 3), ray_d[: r * 3].view(r, 3), near[:r], far[:r], mask.bool())
