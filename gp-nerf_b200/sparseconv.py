"""The sparse-conv geometry encoder without spconv.

Mirror of libs/nerfheads/networks/SparseConvNet.py:21-124: the same module tree
and therefore the same ``state_dict`` keys (``net.<i>.<j>.weight`` with spconv's
weight layout ``[kd, kh, kw, in, out]``, ``net.<i>.<j>.{weight,bias,
running_mean,running_var}`` for the BatchNorm1d layers), so a reference
checkpoint loads with ``strict=True``.  The arithmetic runs in
csrc/k7_sparseconv.cu; nothing is computed by PyTorch except the folding of the
BatchNorm running statistics into a per-channel scale/shift (8 tiny vectors,
cached until the parameters change).

``forward(features, coords, spatial_shape)`` takes what the reference wraps in
``spconv.SparseConvTensor(code, coord, out_sh, 1)`` (trainhead.py:54) and returns
the four levels as *sparse rows* – exactly what ``Engine.upload_products_sparse``
consumes – instead of ``x.dense()`` tensors (SparseConvNet.py:110).

Inference form only: BatchNorm uses its running statistics (``eval()``), as the
progressive renderer does.  Sites: several SMPL vertices can fall into the same
5 mm voxel; spconv's hash table then keeps one of them (which one is
implementation-defined), here the smallest row index owns the voxel.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr


class _ConvWeight(nn.Module):
    """Parameter holder with spconv's layout: weight [3, 3, 3, in, out], no bias."""

    def __init__(self, c_in, c_out, stride):
        super().__init__()
        self.c_in, self.c_out, self.stride = c_in, c_out, stride
        self.weight = nn.Parameter(torch.empty(3, 3, 3, c_in, c_out))
        nn.init.kaiming_uniform_(self.weight.view(27 * c_in, c_out).t(), a=5 ** 0.5)


def _block(chs, stride):
    """[(conv, bn, relu)] * len – indices 0,1,2 / 3,4,5 … as in the reference's SparseSequential."""
    mods = []
    for c_in, c_out in chs:
        mods += [_ConvWeight(c_in, c_out, stride), nn.BatchNorm1d(c_out, eps=1e-3, momentum=0.01), nn.ReLU()]
    return nn.Sequential(*mods)


class SparseConvNet(nn.Module):
    def __init__(self, n_layers=4, in_dim=16, out_dim=(32, 32, 32, 32)):
        super().__init__()
        out_dim = list(out_dim)
        assert len(out_dim) == n_layers == 4, "the hot path gathers from 4 levels"
        self.n_layers = n_layers
        self.net = nn.ModuleList()
        prev = in_dim
        for i in range(n_layers):                     # SparseConvNet.py:96-103
            self.net.append(_block([(prev, prev), (prev, prev)], 1))          # double_conv, 'subm<i>'
            self.net.append(_block([(prev, out_dim[i])], 2))                  # stride_conv, 'down<i>'
            prev = out_dim[i]
        self.net.append(_block([(prev, prev), (prev, prev)], 1))              # double_conv, 'subm<n>'
        self._folded = None

    # ------------------------------------------------------------------
    def _layers(self):
        """(conv, bn) pairs in execution order, tagged with the level they end on."""
        out = []
        for bi, blk in enumerate(self.net):
            for j in range(0, len(blk), 3):
                out.append((bi, blk[j], blk[j + 1]))
        return out

    def _fold(self, device):
        ver = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self._folded is None or self._folded[0] != ver:
            packed = []
            for _bi, conv, bn in self._layers():
                inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
                scale = (bn.weight.detach().float() * inv).contiguous().to(device)
                shift = (bn.bias.detach().float() - bn.running_mean.detach().float() * bn.weight.detach().float() * inv)
                packed.append((conv.weight.detach().float().contiguous().to(device), scale,
                               shift.contiguous().to(device)))
            self._folded = (ver, packed)
        return self._folded[1]

    @torch.no_grad()
    def forward(self, features, coords, spatial_shape):
        """features [N, in_dim] fp32, coords [N, 3|4] int (…, d, h, w), spatial_shape (D, H, W) →
        (levels_sparse, level_dims, n_rows_dev): per level (features [cap, 32], coords [cap, 3]) with
        `cap` = capacity and the live row counts in four device int32 scalars (no host sync)."""
        if self.training:
            raise _lib.GpnerfError("SparseConvNet runs in inference form (BatchNorm running statistics): call .eval()")
        lib = _lib.load()
        dev = features.device
        if dev.type != "cuda":
            raise _lib.GpnerfError("gpnerf_b200 runs on CUDA devices only (no CPU fallback)")
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        feats = features.detach().float().contiguous()
        crd = coords.detach().to(device=dev, dtype=torch.int32).contiguous()
        n0, cols = int(crd.shape[0]), int(crd.shape[1])
        dims = [tuple(int(v) for v in spatial_shape)]
        for _ in range(self.n_layers):
            dims.append(tuple((v - 1) // 2 + 1 for v in dims[-1]))       # SparseConv3d(3, 2, padding=1)
        vox = [d * h * w for d, h, w in dims]
        caps = [n0]
        for k in range(1, 5):
            caps.append(min(8 * caps[-1], vox[k]))
        i32 = dict(dtype=torch.int32, device=dev)
        ws = torch.empty(int(lib.gpnerf_workspace_bytes(max(vox[1], n0))), dtype=torch.uint8, device=dev)
        idx_vol = [torch.empty(v, **i32) for v in vox]
        counts = torch.zeros(5, **i32)
        coords_l = [torch.empty(c * 3, **i32) for c in caps]
        owners = torch.empty(n0, **i32)
        check(lib.gpnerf_sc_index_input(ptr(crd), cols, n0, *dims[0], ptr(idx_vol[0]), ptr(owners), ptr(coords_l[0]),
                                        ptr(counts[0:1]), ptr(ws), st), "sc_index_input")
        c_in = feats.shape[1]
        x = torch.empty(n0 * c_in, dtype=torch.float32, device=dev)
        check(lib.gpnerf_sc_gather_rows(ptr(feats), c_in, ptr(owners), ptr(counts[0:1]), n0, ptr(x), st), "sc_gather_rows")
        packed = self._fold(dev)
        level, li = 0, 0
        outs = []
        lin = torch.empty(max(caps[1:]), **i32)
        for bi, conv, _bn in self._layers():
            w, scale, shift = packed[li]
            li += 1
            if conv.stride == 2:
                check(lib.gpnerf_sc_strided_sites(ptr(coords_l[level]), ptr(counts[level:level + 1]), caps[level],
                                                  *dims[level + 1], ptr(lin), ptr(coords_l[level + 1]),
                                                  ptr(idx_vol[level + 1]), ptr(counts[level + 1:level + 2]), ptr(ws), st),
                      "sc_strided_sites")
                out_level = level + 1
            else:
                out_level = level
            y = torch.empty(caps[out_level] * conv.c_out, dtype=torch.float32, device=dev)
            check(lib.gpnerf_sc_conv(ptr(x), conv.c_in, ptr(idx_vol[level]), *dims[level], ptr(counts[level:level + 1]),
                                     ptr(coords_l[out_level]), ptr(counts[out_level:out_level + 1]), caps[out_level],
                                     conv.stride, ptr(w), ptr(scale), ptr(shift), conv.c_out, ptr(y), st), "sc_conv")
            x, level = y, out_level
            # a level's features are final after the double_conv that follows its stride_conv (blocks 2, 4, 6, 8)
            if bi >= 2 and bi % 2 == 0 and conv is self.net[bi][3]:
                outs.append((x.view(caps[level], conv.c_out), coords_l[level].view(caps[level], 3)))
        n_rows_dev = [counts[k:k + 1] for k in range(1, 5)]
        self._keep = (idx_vol, ws, lin, owners, counts)
        return outs, dims[1:], n_rows_dev
