"""Training path (BASELINE configs[3]): the dense BaseRender render
(libs/renders/BaseRender.py:110-157 → libs/nerfheads/trainhead.py:118-163 →
raw2outputs) under autograd, forward AND backward executed by kernels of
libgpnerf_b200.so.

`render_dense_autograd` is a torch.autograd.Function whose differentiable
inputs are the upstream products (4 dense volume levels, encoder feature maps)
and the head parameters; its outputs are the maps the reference's criterion
consumes (rgb_map, disp, acc, weights, depth, rgb_in_map).  Gradients flow on
into the reference's own producer modules (image encoder, sparse-conv pyramid)
through torch autograd as usual.

First correct training path: gathers use the exact fp32 kernels, the heads run
layer by layer (gpnerf_k6_linear / gpnerf_k6_grad_weights), activations are
kept in HBM between the two passes.  No PyTorch arithmetic is involved.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import CNT_P1, CNT_RAYS, check, ptr, ptr_array
from .engine import Engine

EPI_NONE, EPI_ELU, EPI_RELU, EPI_SIGMOID, EPI_MUL_DELU = 0, 1, 2, 3, 4

# differentiable head parameters, in the order they are passed to the Function
PARAM_KEYS = (
    "sigmahead.out_geometry_fc.0",
    "rgbhead.out_geometry_fc.0", "rgbhead.out_geometry_fc.2", "rgbhead.out_geometry_fc.4", "rgbhead.out_geometry_fc.6",
    "rgbhead.base_fc.0", "rgbhead.base_fc.2", "rgbhead.vis_fc.0", "rgbhead.vis_fc.2",
    "rgbhead.rgb_fc.0", "rgbhead.rgb_fc.2", "rgbhead.rgb_fc.4",
)


def _p(t, off=0):
    return None if t is None else C.c_void_p(t.data_ptr() + 4 * off)


PREC_TRAIN_FP32, PREC_TRAIN_TF32 = 0, 1     # GEMMs of the heads: CUDA-core fp32 (parity) | tcgen05 TF32


class _Kernels:
    def __init__(self, device, precision=PREC_TRAIN_FP32):
        self.lib = _lib.load()
        self.device = device
        self.precision = int(precision)

    def st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def linear(self, X, ldx, K, W, ldw, N, Y, ldy, P, *, xoff=0, woff=0, yoff=0, w_is_kn=False, bias=None, epi=0,
               aux=None, aux_off=0, ld_aux=0, in_aux=None, in_aux_off=0, ld_in_aux=0, in_scale=1.0, add_pre=False,
               add_post=False):
        check(self.lib.gpnerf_k6_linear(_p(X, xoff), ldx, K, C.c_float(in_scale), _p(in_aux, in_aux_off), ld_in_aux,
                                        _p(W, woff), ldw, int(w_is_kn), N, _p(bias), epi, _p(aux, aux_off), ld_aux,
                                        _p(Y, yoff), ldy, int(add_pre), int(add_post), P, self.precision, self.st()),
              "k6_linear")

    def grad_w(self, X, ldx, K, dY, ldy, N, dW, ldw, db, P, *, xoff=0, dyoff=0, dwoff=0, in_scale=1.0, dy_aux=None,
               dy_aux_off=0, ld_dy_aux=0):
        check(self.lib.gpnerf_k6_grad_weights(_p(X, xoff), ldx, K, C.c_float(in_scale), _p(dY, dyoff), ldy, N,
                                              _p(dy_aux, dy_aux_off), ld_dy_aux, _p(dW, dwoff), ldw, _p(db), P,
                                              self.precision, self.st()), "k6_grad_weights")


def _heads_forward(k: _Kernels, w, vol, mv, rf, V, P):
    """Layer-wise forward of both heads on all P points; returns the saved
    activations (each [P, ·] fp32)."""
    dev = vol.device

    def new(n):
        return torch.empty((P, n), dtype=torch.float32, device=dev)
    a = {}
    Wg, bg = w["sigmahead.out_geometry_fc.0"]
    (W0, b0), (W1, b1), (W2, b2), (W3, b3) = (w[f"rgbhead.out_geometry_fc.{i}"] for i in (0, 2, 4, 6))
    (Wb0, bb0), (Wb1, bb1) = w["rgbhead.base_fc.0"], w["rgbhead.base_fc.2"]
    (Wv0, vb0), (Wv1, vb1) = w["rgbhead.vis_fc.0"], w["rgbhead.vis_fc.2"]
    (Wr0, rb0), (Wr1, rb1), (Wr2, rb2) = w["rgbhead.rgb_fc.0"], w["rgbhead.rgb_fc.2"], w["rgbhead.rgb_fc.4"]
    # ---- density head (trainhead.py:39-41, 102-110)
    a["H0"] = new(64)
    k.linear(vol, 128, 128, Wg, 128, 64, a["H0"], 64, P, bias=bg, epi=EPI_ELU)
    a["H1"] = new(64)      # Linear(134→64) on [H0 | mean | var] as two accumulating products
    k.linear(a["H0"], 64, 64, W0, 134, 64, a["H1"], 64, P)
    k.linear(mv, 70, 70, W0, 134, 64, a["H1"], 64, P, woff=64, bias=b0, epi=EPI_ELU, add_pre=True)
    a["H2"] = new(32)
    k.linear(a["H1"], 64, 64, W1, 64, 32, a["H2"], 32, P, bias=b1, epi=EPI_ELU)
    a["H3"] = new(16)
    k.linear(a["H2"], 32, 32, W2, 32, 16, a["H3"], 16, P, bias=b2, epi=EPI_ELU)
    a["S"] = new(1)        # ReLU output; the no-valid-view fill happens in assemble_raw
    k.linear(a["H3"], 16, 16, W3, 16, 1, a["S"], 1, P, bias=b3, epi=EPI_RELU)
    # ---- colour head (trainhead.py:85-100, 128-145)
    a["XB"], a["XV"], a["T"], a["U"] = new(64 * V), new(32 * V), new(32 * V), new(32 * V)
    for v in range(V):
        k.linear(mv, 70, 70, Wb0, 105, 64, a["XB"], 64 * V, P, yoff=64 * v)
        k.linear(rf, 35 * V, 35, Wb0, 105, 64, a["XB"], 64 * V, P, xoff=35 * v, woff=70, yoff=64 * v, bias=bb0,
                 epi=EPI_ELU, add_pre=True)
        k.linear(a["XB"], 64 * V, 64, Wb1, 64, 32, a["XV"], 32 * V, P, xoff=64 * v, yoff=32 * v, bias=bb1, epi=EPI_ELU)
        k.linear(a["XV"], 32 * V, 32, Wv0, 32, 32, a["T"], 32 * V, P, xoff=32 * v, yoff=32 * v, bias=vb0,
                 epi=EPI_ELU, in_scale=1.0 / V)
        k.linear(a["T"], 32 * V, 32, Wv1, 32, 32, a["U"], 32 * V, P, xoff=32 * v, yoff=32 * v, bias=vb1, epi=EPI_ELU)
    a["R0"] = new(32)      # rgb_fc.0 on flat = XV + U, by linearity as two accumulating products
    k.linear(a["XV"], 32 * V, 32 * V, Wr0, 32 * V, 32, a["R0"], 32, P)
    k.linear(a["U"], 32 * V, 32 * V, Wr0, 32 * V, 32, a["R0"], 32, P, bias=rb0, epi=EPI_ELU, add_pre=True)
    a["R1"] = new(16)
    k.linear(a["R0"], 32, 32, Wr1, 32, 16, a["R1"], 16, P, bias=rb1, epi=EPI_ELU)
    a["RGB"] = new(3)
    k.linear(a["R1"], 16, 16, Wr2, 16, 3, a["RGB"], 3, P, bias=rb2, epi=EPI_SIGMOID)
    return a


def _heads_backward(k: _Kernels, w, a, vol, mv, rf, V, P, d_rgb_pre, d_s_pre):
    """Backward of both heads.  Returns (grads of the 12 (W, b) pairs,
    d_vol [P,128], d_mv [P,70], d_rf [P,V*35])."""
    dev = vol.device

    def new(n, zero=False):
        return (torch.zeros if zero else torch.empty)((P, n), dtype=torch.float32, device=dev)
    g = {key: (torch.zeros_like(w[key][0]), torch.zeros_like(w[key][1])) for key in PARAM_KEYS}
    Wg = w["sigmahead.out_geometry_fc.0"][0]
    W0, W1, W2, W3 = (w[f"rgbhead.out_geometry_fc.{i}"][0] for i in (0, 2, 4, 6))
    Wb0, Wb1 = w["rgbhead.base_fc.0"][0], w["rgbhead.base_fc.2"][0]
    Wv0, Wv1 = w["rgbhead.vis_fc.0"][0], w["rgbhead.vis_fc.2"][0]
    Wr0, Wr1, Wr2 = w["rgbhead.rgb_fc.0"][0], w["rgbhead.rgb_fc.2"][0], w["rgbhead.rgb_fc.4"][0]
    d_mv = new(70, zero=True)
    d_rf = new(35 * V)
    # ---- colour head
    gW, gb = g["rgbhead.rgb_fc.4"]
    k.grad_w(a["R1"], 16, 16, d_rgb_pre, 3, 3, gW, 16, gb, P)
    dR1 = new(16)
    k.linear(d_rgb_pre, 3, 3, Wr2, 16, 16, dR1, 16, P, w_is_kn=True, epi=EPI_MUL_DELU, aux=a["R1"], ld_aux=16)
    gW, gb = g["rgbhead.rgb_fc.2"]
    k.grad_w(a["R0"], 32, 32, dR1, 16, 16, gW, 32, gb, P)
    dR0 = new(32)
    k.linear(dR1, 16, 16, Wr1, 32, 32, dR0, 32, P, w_is_kn=True, epi=EPI_MUL_DELU, aux=a["R0"], ld_aux=32)
    gW, gb = g["rgbhead.rgb_fc.0"]
    k.grad_w(a["XV"], 32 * V, 32 * V, dR0, 32, 32, gW, 32 * V, gb, P)
    k.grad_w(a["U"], 32 * V, 32 * V, dR0, 32, 32, gW, 32 * V, None, P)
    dFL = new(32 * V)                                  # ∂L/∂flat = ∂L/∂U = residual part of ∂L/∂XV
    k.linear(dR0, 32, 32, Wr0, 32 * V, 32 * V, dFL, 32 * V, P, w_is_kn=True)
    dT, dXB = new(32 * V), new(64 * V)
    for v in range(V):
        o32, o64 = 32 * v, 64 * v
        # vis_fc.2: dU_pre = dFL_v ⊙ ELU'(U_v) is applied on the fly (dy_aux / in_aux)
        gW, gb = g["rgbhead.vis_fc.2"]
        k.grad_w(a["T"], 32 * V, 32, dFL, 32 * V, 32, gW, 32, gb, P, xoff=o32, dyoff=o32, dy_aux=a["U"],
                 dy_aux_off=o32, ld_dy_aux=32 * V)
        k.linear(dFL, 32 * V, 32, Wv1, 32, 32, dT, 32 * V, P, xoff=o32, yoff=o32, w_is_kn=True, in_aux=a["U"],
                 in_aux_off=o32, ld_in_aux=32 * V, epi=EPI_MUL_DELU, aux=a["T"], aux_off=o32, ld_aux=32 * V)
        # vis_fc.0 (input XV_v / V)
        gW, gb = g["rgbhead.vis_fc.0"]
        k.grad_w(a["XV"], 32 * V, 32, dT, 32 * V, 32, gW, 32, gb, P, xoff=o32, dyoff=o32, in_scale=1.0 / V)
        # dXV_pre = (dFL_v + dT_v·Wv0 / V) ⊙ ELU'(XV_v), in place in dFL_v
        k.linear(dT, 32 * V, 32, Wv0, 32, 32, dFL, 32 * V, P, xoff=o32, yoff=o32, w_is_kn=True, in_scale=1.0 / V,
                 add_pre=True, epi=EPI_MUL_DELU, aux=a["XV"], aux_off=o32, ld_aux=32 * V)
        gW, gb = g["rgbhead.base_fc.2"]
        k.grad_w(a["XB"], 64 * V, 64, dFL, 32 * V, 32, gW, 64, gb, P, xoff=o64, dyoff=o32)
        k.linear(dFL, 32 * V, 32, Wb1, 64, 64, dXB, 64 * V, P, xoff=o32, yoff=o64, w_is_kn=True, epi=EPI_MUL_DELU,
                 aux=a["XB"], aux_off=o64, ld_aux=64 * V)
        gW, gb = g["rgbhead.base_fc.0"]
        k.grad_w(mv, 70, 70, dXB, 64 * V, 64, gW, 105, gb, P, dyoff=o64)
        k.grad_w(rf, 35 * V, 35, dXB, 64 * V, 64, gW, 105, None, P, xoff=35 * v, dyoff=o64, dwoff=70)
        k.linear(dXB, 64 * V, 64, Wb0, 105, 70, d_mv, 70, P, xoff=o64, w_is_kn=True, add_post=True)
        k.linear(dXB, 64 * V, 64, Wb0, 105, 35, d_rf, 35 * V, P, xoff=o64, woff=70, yoff=35 * v, w_is_kn=True)
    # ---- density head
    gW, gb = g["rgbhead.out_geometry_fc.6"]
    k.grad_w(a["H3"], 16, 16, d_s_pre, 1, 1, gW, 16, gb, P)
    dH3 = new(16)
    k.linear(d_s_pre, 1, 1, W3, 16, 16, dH3, 16, P, w_is_kn=True, epi=EPI_MUL_DELU, aux=a["H3"], ld_aux=16)
    gW, gb = g["rgbhead.out_geometry_fc.4"]
    k.grad_w(a["H2"], 32, 32, dH3, 16, 16, gW, 32, gb, P)
    dH2 = new(32)
    k.linear(dH3, 16, 16, W2, 32, 32, dH2, 32, P, w_is_kn=True, epi=EPI_MUL_DELU, aux=a["H2"], ld_aux=32)
    gW, gb = g["rgbhead.out_geometry_fc.2"]
    k.grad_w(a["H1"], 64, 64, dH2, 32, 32, gW, 64, gb, P)
    dH1 = new(64)
    k.linear(dH2, 32, 32, W1, 64, 64, dH1, 64, P, w_is_kn=True, epi=EPI_MUL_DELU, aux=a["H1"], ld_aux=64)
    gW, gb = g["rgbhead.out_geometry_fc.0"]
    k.grad_w(a["H0"], 64, 64, dH1, 64, 64, gW, 134, gb, P)
    k.grad_w(mv, 70, 70, dH1, 64, 64, gW, 134, None, P, dwoff=64)
    dH0 = new(64)
    k.linear(dH1, 64, 64, W0, 134, 64, dH0, 64, P, w_is_kn=True, epi=EPI_MUL_DELU, aux=a["H0"], ld_aux=64)
    k.linear(dH1, 64, 64, W0, 134, 70, d_mv, 70, P, woff=64, w_is_kn=True, add_post=True)
    gW, gb = g["sigmahead.out_geometry_fc.0"]
    k.grad_w(vol, 128, 128, dH0, 64, 64, gW, 128, gb, P)
    d_vol = new(128)
    k.linear(dH0, 64, 64, Wg, 128, 128, d_vol, 128, P, w_is_kn=True)
    return g, d_vol, d_mv, d_rf


class _RenderDenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: Engine, frame, rays, t_rand, neg_ray, src_imgs, n_levels, *tensors):
        levels, featmaps, params = list(tensors[:n_levels]), tensors[n_levels], tensors[n_levels + 1:]
        assert len(params) == 2 * len(PARAM_KEYS)
        dev = eng.device
        k = _Kernels(dev, getattr(eng, "train_precision", PREC_TRAIN_FP32))
        L, st = eng.lib, k.st()
        w = {key: (params[2 * i].detach().float().contiguous(), params[2 * i + 1].detach().float().contiguous())
             for i, key in enumerate(PARAM_KEYS)}
        ray_o, ray_d, near, far = rays
        R, S, V = int(ray_d.shape[0]), eng.S, eng.V
        P = R * S
        if eng.bf16:
            raise _lib.GpnerfError("the training path uses the fp32 engine")
        if R > eng.max_rays:
            raise _lib.GpnerfError(f"{R} rays exceed the engine capacity {eng.max_rays}")
        eng.upload_products([t.detach() for t in levels], featmaps.detach(), src_imgs)
        eng.upload_frame(frame)
        f32 = dict(dtype=torch.float32, device=dev)
        eng.rays_o.copy_(ray_o.detach().to(**f32).reshape(-1, 3)[0])
        eng.rays_d[: R * 3].copy_(ray_d.detach().to(**f32).reshape(-1))
        eng.near[:R].copy_(near.detach().to(**f32).reshape(-1))
        eng.far[:R].copy_(far.detach().to(**f32).reshape(-1))
        eng.counters[CNT_RAYS:CNT_RAYS + 1].fill_(R)
        tr = None if t_rand is None else t_rand.detach().to(**f32).reshape(-1).contiguous()
        fr = C.byref(frame)
        eng._run("k2_occupancy_compact", L.gpnerf_k2_occupancy_compact, None, ptr(eng.rays_o), ptr(eng.rays_d),
                 ptr(eng.near), ptr(eng.far), ptr(eng.t_vals), ptr(tr), fr, R, ptr(eng.valid), ptr(eng.z_vals),
                 ptr(eng.counters), ptr(eng.workspace), None, st)
        eng._run("k2_gather_volume", L.gpnerf_k2_gather_volume, ptr_array(eng.levels_cl), 0, ptr(eng.valid),
                 ptr(eng.rays_o), ptr(eng.rays_d), ptr(eng.z_vals), None, fr, P, ptr(eng.counters),
                 ptr(eng.vol_feat), st)
        eng._run("k2_project_gather_meanvar", L.gpnerf_k2_project_gather_meanvar, ptr(eng.images_rgbx),
                 ptr(eng.featmaps_cl), 0, ptr(eng.valid), ptr(eng.rays_o), ptr(eng.rays_d), ptr(eng.z_vals), None, fr,
                 P, ptr(eng.counters), ptr(eng.rgb_feat), ptr(eng.mask), ptr(eng.meanvar), st)
        vol = eng.vol_feat[: P * 128].view(P, 128).clone()
        rf = eng.rgb_feat[: P * V * 35].view(P, V * 35).clone()
        mv = eng.meanvar[: P * 70].view(P, 70).clone()
        mask = eng.mask[: P * V].view(P, V).clone()
        acts = _heads_forward(k, w, vol, mv, rf, V, P)
        raw = torch.empty((P, 4), **f32)
        check(L.gpnerf_k6_assemble_raw(ptr(acts["RGB"]), ptr(acts["S"]), ptr(mask), V, P, ptr(raw), st),
              "k6_assemble_raw")
        z = eng.z_vals[:P].clone()
        rgb_in = rf.view(P, V, 35)[..., :3].contiguous()
        out = {n: torch.empty(s, **f32) for n, s in (("rgb_map", (R, 3)), ("disp", (R,)), ("acc", (R,)),
                                                     ("depth", (R,)), ("weights", (R, S)), ("rgb_in_map", (R, V * 3)))}
        check(L.gpnerf_k5_raw2outputs(ptr(raw), ptr(z), ptr(rgb_in), R, S, V, int(neg_ray), ptr(out["rgb_map"]),
                                      ptr(out["disp"]), ptr(out["acc"]), ptr(out["depth"]), ptr(out["weights"]),
                                      ptr(out["rgb_in_map"]), st), "k5_raw2outputs")
        ctx.eng, ctx.k, ctx.w, ctx.acts, ctx.frame = eng, k, w, acts, frame
        ctx.saved = (vol, rf, mv, mask, raw, z, rgb_in, eng.valid[:P].clone(), eng.rays_o.clone(),
                     eng.rays_d[: R * 3].clone())
        ctx.dims = (R, S, V, P, bool(neg_ray), n_levels, [tuple(t.shape) for t in levels], tuple(featmaps.shape))
        ctx.z_out = z.view(R, S)
        return (out["rgb_map"], out["disp"], out["acc"], out["weights"], out["depth"], out["rgb_in_map"],
                z.view(R, S).clone())

    @staticmethod
    def backward(ctx, g_rgb_map, g_disp, g_acc, g_weights, g_depth, g_rin, _g_z):
        eng, k, w, a, frame = ctx.eng, ctx.k, ctx.w, ctx.acts, ctx.frame
        vol, rf, mv, mask, raw, z, rgb_in, valid, rays_o, rays_d = ctx.saved
        R, S, V, P, neg, n_levels, level_shapes, fm_shape = ctx.dims
        L, st, dev = eng.lib, k.st(), eng.device
        f32 = dict(dtype=torch.float32, device=dev)

        def c(t):
            return None if t is None else t.detach().to(**f32).contiguous()
        d_raw = torch.empty((P, 4), **f32)
        check(L.gpnerf_k5_raw2outputs_bwd(ptr(raw), ptr(z), ptr(rgb_in), R, S, V, int(neg), ptr(c(g_rgb_map)),
                                          ptr(c(g_disp)), ptr(c(g_acc)), ptr(c(g_depth)), ptr(c(g_weights)),
                                          ptr(c(g_rin)), ptr(d_raw), st), "k5_raw2outputs_bwd")
        d_rgb_pre, d_s_pre = torch.empty((P, 3), **f32), torch.empty((P, 1), **f32)
        check(L.gpnerf_k6_raw_grad_split(ptr(d_raw), ptr(a["RGB"]), ptr(a["S"]), ptr(mask), V, P, ptr(d_rgb_pre),
                                         ptr(d_s_pre), st), "k6_raw_grad_split")
        g, d_vol, d_mv, d_rf = _heads_backward(k, w, a, vol, mv, rf, V, P, d_rgb_pre, d_s_pre)
        check(L.gpnerf_k6_meanvar_bwd(ptr(rf), ptr(mv), ptr(d_mv), V, P, ptr(d_rf), st), "k6_meanvar_bwd")
        fr = C.byref(frame)
        eng.upload_frame(frame)
        need = ctx.needs_input_grad[7:]
        grads_levels = [None] * n_levels
        if any(need[:n_levels]):
            d_cl = [torch.zeros(int(s[-3] * s[-2] * s[-1]) * 32, **f32) for s in level_shapes]
            check(L.gpnerf_k2_gather_volume_bwd(ptr_array(d_cl), 0, ptr(valid), ptr(rays_o), ptr(rays_d), ptr(z), None,
                                                fr, P, None, ptr(d_vol), st), "k2_gather_volume_bwd")
            for i, (s, d) in enumerate(zip(level_shapes, d_cl)):
                out = torch.empty(s, **f32)
                check(L.gpnerf_k6_from_channels_last(ptr(d), 1, int(s[-3] * s[-2] * s[-1]), ptr(out), st),
                      "k6_from_channels_last")
                grads_levels[i] = out
        grad_fm = None
        if need[n_levels]:
            Vv, Cc, fh, fw = fm_shape
            d_fm_cl = torch.zeros(Vv * fh * fw * 32, **f32)
            check(L.gpnerf_k2_project_gather_bwd(ptr(d_fm_cl), 0, ptr(valid), ptr(rays_o), ptr(rays_d), ptr(z), None,
                                                 fr, P, None, ptr(d_rf), st), "k2_project_gather_bwd")
            grad_fm = torch.empty(fm_shape, **f32)
            check(L.gpnerf_k6_from_channels_last(ptr(d_fm_cl), Vv, fh * fw, ptr(grad_fm), st), "k6_from_channels_last")
        grads_params = []
        for key in PARAM_KEYS:
            grads_params.extend(g[key])
        return (None, None, None, None, None, None, None, *grads_levels, grad_fm, *grads_params)


def render_dense_autograd(eng: Engine, frame, rays, levels, featmaps, src_imgs, head_params, t_rand=None,
                          neg_ray=False, precision=PREC_TRAIN_FP32):
    """Differentiable dense render.  `head_params`: dict name → tensor with the
    reference state_dict keys (prefix 'nerfhead.' optional).  Returns a dict
    with the BaseRender output keys (BaseRender.py:148-156).  `precision`:
    PREC_TRAIN_FP32 (CUDA-core fp32 GEMMs, the parity path) or PREC_TRAIN_TF32
    (the heads' forward and backward GEMMs on tcgen05 in TF32)."""
    eng.train_precision = int(precision)
    def get(name):
        for pre in ("", "nerfhead."):
            if pre + name in head_params:
                return head_params[pre + name]
        raise KeyError(name)
    flat = []
    for key in PARAM_KEYS:
        flat.extend([get(key + ".weight"), get(key + ".bias")])
    outs = _RenderDenseFn.apply(eng, frame, rays, t_rand, neg_ray, src_imgs, len(levels), *levels, featmaps, *flat)
    rgb_map, disp, acc, weights, depth, rin, z = outs
    return {"rgb_map": rgb_map, "disp_map": disp[:, None], "acc_map": acc[:, None], "depth_map": depth[:, None],
            "alpha": weights, "z_vals": z, "rgb_in_map": rin}


class GradBucket:
    """Data-parallel training over the GPUs of one box (BASELINE configs[3]: the
    batch's rays are split over the ranks, every rank holds the full heads).

    All head gradients live in ONE flat fp32 buffer (`param.grad` of every
    parameter is a view into it, so autograd accumulates straight into the
    bucket) and are averaged with a single all-reduce per step – 259,716 values
    ≈ 1 MB for the reference's heads – NCCL over NVLink on GPUs,
    gloo in the CPU tests.  The reference wraps the model in DDP but calls
    `.module.render`, which bypasses DDP's reducer; this is the
    reduction it meant to have."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        self.group = group
        self._views = []
        off = 0
        for p in self.params:
            n = p.numel()
            self._views.append(self.flat[off:off + n].view_as(p))
            p.grad = self._views[-1]
            off += n

    def zero(self):
        self.flat.zero_()
        self._reattach()

    def _reattach(self):
        """`optimizer.zero_grad()` (set_to_none=True by default) or an optimizer that replaces `.grad` breaks the
        aliasing between `param.grad` and the flat buffer: gradients that were accumulated elsewhere are copied in
        and the views are attached again, so the all-reduce never runs on stale values."""
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is None:
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v

    def all_reduce_mean(self):
        """Average the bucket over the ranks (no-op without a process group)."""
        import torch.distributed as dist
        self._reattach()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat


class GraphedStep:
    """One training step – forward, backward, the bucket's all-reduce, the optimizer – captured into ONE CUDA graph
    and replayed per step.  A step of the hot path is ~130 launches of 20-100 us kernels (74 of them the K6 GEMMs);
    issued one by one from Python they cost ~2.9 ms of host time, which is what a rank's step takes once its share
    of the rays is small (4096 rays over 8 GPUs).  Requirements on `step_fn`: no host synchronisation, inputs read
    from fixed buffers (e.g. the per-step jitter in a pinned host tensor: its upload is part of the graph), an
    optimizer created with `capturable=True`.  `step_fn` is run `warmup` times on a side stream first (allocations,
    kernel attributes, NCCL channels).  Returns whatever `step_fn` returned at capture time (same tensors each
    replay)."""

    def __init__(self, step_fn, device, warmup=3):
        self.device = torch.device(device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self.out = step_fn()
        torch.cuda.synchronize(self.device)

    def __call__(self):
        self.graph.replay()
        return self.out

