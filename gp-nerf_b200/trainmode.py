"""Training form of the upstream producers, autograd-connected.

The hot path itself (gathers, heads, compositing: rows a3–a17) trains through the K6 kernels of
libgpnerf_b200.so (`train.render_dense_autograd`).  Its producers – image encoder, SMPL-code
attention, sparse-conv pyramid – are "next" rows whose *inference* form runs in this library's K7/K8/K9
kernels; for `tools/train.py` (BaseTrainer.py:99-131: render → criterion → `backward()` → AdamW) they also
need gradients and, for the pyramid, batch-statistics BatchNorm (SparseConvNet.py:21-87 under `.train()`).
This module supplies that form on torch's own differentiable ops, reading the very same parameters
(`nn.Conv2d`, `nn.InstanceNorm2d`, `nn.BatchNorm1d`, `nn.Linear`, spconv-layout weights) the kernels read:

* encoder: the children of `encoder.ResUNet` called in the order of UNet.py:217-234 (cuDNN convolutions);
* attention: MultiHeadAttention.py:40-98 with `sum=False` as three matmuls and a softmax over the views;
* pyramid: site lists and the 27-entry neighbour tables come from the K7 geometry kernels (integer work, no
  gradient); each convolution is `rows[nbr] · W` (index_select + one GEMM), BatchNorm1d with batch statistics
  (running statistics updated as `nn.BatchNorm1d` does), ReLU;
* the levels reach the K6 path as dense NCDHW tensors built by a differentiable scatter (what
  `SparseConvTensor.dense()` returns, SparseConvNet.py:110), so level gradients flow back into the rows.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr


# ----------------------------------------------------------------------------- f2: image encoder
def encoder_forward(enc, x):
    """UNet.py:217-234 on the mirror's own children (same parameters as the K9/cuDNN inference form)."""
    def block(b, t):                                    # UNet.py:38-53
        idt = t
        out = F.relu(b.bn1(b.conv1(t)))
        out = b.bn2(b.conv2(out))
        if b.downsample is not None:
            idt = b.downsample(t)
        return F.relu(out + idt)

    def seq(layer, t):
        for b in layer:
            t = block(b, t)
        return t

    def conv_bn_elu(m, t):                              # UNet.py:106-119
        return F.elu(m.bn(m.conv(t)))

    def up(m, t):                                       # UNet.py:122-131
        return conv_bn_elu(m.conv, F.interpolate(t, scale_factor=m.scale, align_corners=True, mode="bilinear"))

    def skip(x1, x2):                                   # UNet.py:203-215
        dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
        x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        return torch.cat([x2, x1], dim=1)

    x = F.relu(enc.bn1(enc.conv1(x)))
    x1 = seq(enc.layer1, x)
    x2 = seq(enc.layer2, x1)
    x3 = seq(enc.layer3, x2)
    x = conv_bn_elu(enc.iconv3, skip(x2, up(enc.upconv3, x3)))
    x = conv_bn_elu(enc.iconv2, skip(x1, up(enc.upconv2, x)))
    return enc.out_conv(x)


# ----------------------------------------------------------------------------- f1: SMPL-code attention
def attention_forward(attn, q, kv):
    """MultiHeadAttention.py:62-98 with sum=False: q [n,1,d_model], kv [n,V,kv_dim] → [n,1,d_model]."""
    n, lq, _ = q.shape
    V = kv.shape[1]
    h, dk = attn.n_head, attn.d_k
    qh = attn.w_qs(q).view(n, lq, h, dk).transpose(1, 2)
    kh = attn.w_ks(kv).view(n, V, h, dk).transpose(1, 2)
    vh = attn.w_vs(kv).view(n, V, h, dk).transpose(1, 2)
    a = torch.softmax(torch.matmul(qh / dk ** 0.5, kh.transpose(2, 3)), dim=-1)
    out = torch.matmul(a, vh).transpose(1, 2).contiguous().view(n, lq, -1)
    return attn.fc(out)


def smpl_features(smpl_xyz, cams, featmaps, neg_ray=False):
    """Projector.compute_smpl (demo_render.py:612-632 / BaseRender.py:283-363 without the RGB part):
    feature-map samples at the projected SMPL vertices, differentiable in `featmaps` → [1, n, V, C]."""
    V = featmaps.shape[0]
    cams = cams.reshape(-1, 34).to(featmaps.device)
    h, w = cams[0, 0], cams[0, 1]
    KE = cams[:, 2:18].reshape(V, 4, 4).bmm(cams[:, 18:].reshape(V, 4, 4))
    xyz = smpl_xyz.reshape(-1, 3)
    xyz_h = torch.cat([xyz, torch.ones_like(xyz[:, :1])], 1)
    proj = KE.bmm(xyz_h.t()[None].expand(V, -1, -1)).permute(0, 2, 1)[..., :3]            # [V,n,3]
    pix = torch.clamp(proj[..., :2] / proj[..., 2:3], min=-1e6, max=1e6)           # BaseRender.py:315-316
    res = torch.stack([w - 1.0, h - 1.0]).to(pix)
    grid = 2 * pix / res - 1.0
    samp = F.grid_sample(featmaps, grid[:, None], align_corners=True)                     # [V,C,1,n]
    return samp[:, :, 0].permute(2, 0, 1)[None]                                           # [1,n,V,C]


# ----------------------------------------------------------------------------- f1: sparse-conv pyramid
def _geometry(net, coords, spatial_shape, c_in, dev):
    """Site lists + neighbour tables of one pyramid from the K7 geometry kernels (one stream, eager)."""
    lib = _lib.load()
    pl = net._plan(int(coords.shape[0]), int(coords.shape[1]), spatial_shape, c_in, dev)
    pl["coord_in"].copy_(coords.detach())
    dims, caps, counts, coords_l, idx_vol = pl["dims"], pl["caps"], pl["counts"], pl["coords"], pl["idx_vol"]
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    n0, cols = pl["coord_in"].shape
    check(lib.gpnerf_sc_index_input(ptr(pl["coord_in"]), cols, n0, *dims[0], ptr(idx_vol[0]), ptr(pl["owners"]),
                                    ptr(coords_l[0]), ptr(counts[0:1]), ptr(pl["ws"]), st), "sc_index_input")
    tables, level, subm = {}, 0, None
    for li, (_bi, conv, _bn) in enumerate(net._layers()):
        if conv.stride == 2:
            check(lib.gpnerf_sc_strided_sites(ptr(coords_l[level]), ptr(counts[level:level + 1]), caps[level],
                                              *dims[level + 1], ptr(pl["lin"]), ptr(coords_l[level + 1]),
                                              ptr(idx_vol[level + 1]), ptr(counts[level + 1:level + 2]), ptr(pl["ws"]),
                                              st), "sc_strided_sites")
            in_level, out_level, subm = level, level + 1, None
        else:
            in_level = out_level = level
        if conv.stride == 2 or subm is None:
            nbr = pl["nbr"][li]
            check(lib.gpnerf_sc_neighbours(ptr(coords_l[out_level]), ptr(counts[out_level:out_level + 1]),
                                           caps[out_level], conv.stride, ptr(idx_vol[in_level]), *dims[in_level],
                                           ptr(counts[in_level:in_level + 1]), ptr(nbr), st), "sc_neighbours")
            entry = (nbr, in_level, out_level)
            if conv.stride == 1:
                subm = entry
        else:
            entry = (subm[0], in_level, out_level)
        tables[li] = entry
        level = out_level
    n = [int(v) for v in counts.cpu().tolist()]            # the training form sizes its tensors on the host
    return pl, tables, n


def pyramid_forward(net, features, coords, spatial_shape):
    """SparseConvNet.py:104-110 in training form → per level (rows [n_k, C], coords [n_k, 3]) and the 4 level
    dims; differentiable in `features` and in the convolution / BatchNorm parameters."""
    dev = features.device
    pl, tables, n = _geometry(net, coords, spatial_shape, int(features.shape[1]), dev)
    caps = pl["caps"]
    x = features.index_select(0, pl["owners"][: n[0]].long())          # one row per occupied voxel (smallest row id)
    outs = []
    for li, (bi, conv, bn) in enumerate(net._layers()):
        nbr, in_level, out_level = tables[li]
        n_in, n_out = n[in_level], n[out_level]
        idx = nbr.view(27, caps[out_level])[:, :n_out].long()
        idx = torch.where(idx < 0, torch.full_like(idx, n_in), idx)     # "no neighbour" → the appended zero row
        xp = torch.cat([x, x.new_zeros(1, x.shape[1])], 0)
        g = xp[idx.t()]                                                 # [n_out, 27, c_in]
        y = g.reshape(n_out, -1) @ conv.weight.reshape(27 * conv.c_in, conv.c_out)
        y = bn(y) if n_out > 1 or not bn.training else F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight,
                                                                    bn.bias, False, 0.0, bn.eps)
        x = F.relu(y)
        if net._closes_level(bi, conv):
            outs.append((x, pl["coords"][out_level].view(caps[out_level], 3)[:n_out]))
    return outs, pl["dims"][1:]


def rows_to_dense(rows, coords, dims):
    """`SparseConvTensor.dense()` (SparseConvNet.py:110) as a differentiable scatter → [1, C, D, H, W]."""
    D, H, W = dims
    lin = (coords[:, 0].long() * H + coords[:, 1].long()) * W + coords[:, 2].long()
    vol = rows.new_zeros(rows.shape[1], D * H * W)
    vol = vol.index_copy(1, lin, rows.t())
    return vol.view(1, rows.shape[1], D, H, W)
