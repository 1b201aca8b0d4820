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


def _pack_tf32x3(w):
    """spconv weight [3,3,3,in,out] → per tap the two UMMA B-operand images of W[k]ᵀ ([out × in], K-major,
    8 × 16-byte core matrices) that csrc/k7_sparseconv_tc.cu multiplies with: the part that is exact in TF32 (13
    low mantissa bits cleared) and the remainder.  float32 [27, 2, out·in]."""
    c_in, c_out = w.shape[3], w.shape[4]
    b = w.reshape(27, c_in, c_out).transpose(1, 2).contiguous()                   # [27, out, in]
    hi = (b.view(torch.int32) & -8192).view(torch.float32)                         # 0xffffe000
    lo = b - hi

    def image(m):                                                                  # (n, c) → (n/8, c/4, n%8, c%4)
        return m.reshape(27, c_out // 8, 8, c_in // 4, 4).permute(0, 1, 3, 2, 4).reshape(27, c_out * c_in)
    return torch.stack([image(hi), image(lo)], 1).contiguous()


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
        self._plist = None
        self._plans = {}
        self.use_cuda_graph = True
        # "fp32": CUDA-core FFMA kernel (cluster tap split) – the default, and the faster one: the pyramid's tiles are
        # latency-bound, not math-bound.  "tf32x3": tcgen05 kind::tf32 with a 3-term split (fp32-grade accuracy),
        # kept as the tensor-core variant (profiles/r01_final_ncu_summary.md §8: 0.58 ms vs 0.47 ms per pyramid)
        self.precision = "fp32"

    # ------------------------------------------------------------------
    def _layers(self):
        """(conv, bn) pairs in execution order, tagged with the level they end on."""
        out = []
        for bi, blk in enumerate(self.net):
            for j in range(0, len(blk), 3):
                out.append((bi, blk[j], blk[j + 1]))
        return out

    def _fold(self, device):
        """BatchNorm (running statistics) as per-channel scale/shift; weights as fp32 [27·in, out].  The device
        tensors are allocated once and refreshed in place, so a captured graph keeps reading the right memory."""
        if self._plist is None or self._plist[0].device != device:
            self._plist = list(self.parameters()) + list(self.buffers())
        ver = tuple((p.data_ptr(), p._version) for p in self._plist)
        if self._folded is not None and self._folded[0] == ver and self._folded[1] == device:
            return self._folded[2]
        fresh = []
        for _bi, conv, bn in self._layers():
            g, inv = bn.weight.detach().float(), torch.rsqrt(bn.running_var.detach().float() + bn.eps)
            w = conv.weight.detach().float().contiguous()
            fresh.append((w, g * inv, bn.bias.detach().float() - bn.running_mean.detach().float() * g * inv,
                          _pack_tf32x3(w)))
        if self._folded is not None and self._folded[1] == device:
            packed = self._folded[2]
            for old, new in zip(packed, fresh):
                for o, n in zip(old, new):
                    o.copy_(n)
        else:
            packed = [tuple(t.to(device).contiguous().clone() for t in trip) for trip in fresh]
            self._plans = {}
        self._folded = (ver, device, packed)
        return packed

    # ------------------------------------------------------------------
    def _plan(self, n0, cols, spatial_shape, c_in, dev):
        """Buffers of one (vertex count, grid) geometry; every row count stays on the device, so the launch
        sequence is the same for every frame and is captured into a CUDA graph on its second use."""
        tc = self.precision == "tf32x3"          # rows travel as [hi | lo] between the layers: twice the width
        key = (n0, cols, tuple(spatial_shape), str(dev), tc)
        pl = self._plans.get(key)
        if pl is not None:
            return pl
        lib = _lib.load()
        dims = [tuple(int(v) for v in spatial_shape)]
        for _ in range(self.n_layers):
            dims.append(tuple((v - 1) // 2 + 1 for v in dims[-1]))       # SparseConv3d(3, 2, padding=1)
        vox = [d * h * w for d, h, w in dims]
        caps = [n0]
        for k in range(1, 5):
            caps.append(min(8 * caps[-1], vox[k]))
        i32 = dict(dtype=torch.int32, device=dev)
        pl = {
            "dims": dims, "caps": caps,
            "feat_in": torch.empty(n0, c_in, dtype=torch.float32, device=dev),
            "coord_in": torch.empty(n0, cols, **i32),
            "ws": torch.empty(int(lib.gpnerf_workspace_bytes(max(vox[1], n0))), dtype=torch.uint8, device=dev),
            "idx_vol": [torch.empty(v, **i32) for v in vox],
            "counts": torch.zeros(5, **i32),
            "coords": [torch.empty(c * 3, **i32) for c in caps],
            "owners": torch.empty(n0, **i32),
            "lin": torch.empty(max(caps[1:]), **i32),
            "nbr": {}, "side": torch.cuda.Stream(dev),
            "x0": torch.empty(n0 * c_in * (2 if tc else 1), dtype=torch.float32, device=dev),
            "y": [], "y_full": {}, "graph": None, "uses": 0,
        }
        level, have_subm = 0, False
        for li, (bi, conv, _bn) in enumerate(self._layers()):
            level += conv.stride == 2
            if conv.stride == 2 or not have_subm:             # one neighbour table per (site list, stride)
                pl["nbr"][li] = torch.empty(27 * caps[level], **i32)
            have_subm = conv.stride == 1
            pl["y"].append(torch.empty(caps[level] * conv.c_out * (2 if tc else 1), dtype=torch.float32, device=dev))
            if tc and self._closes_level(bi, conv):
                pl["y_full"][li] = torch.empty(caps[level] * conv.c_out, dtype=torch.float32, device=dev)
        self._plans[key] = pl
        return pl

    def _closes_level(self, bi, conv):
        """A level's features are final after the double_conv that follows its stride_conv (blocks 2, 4, 6, 8)."""
        return bi >= 2 and bi % 2 == 0 and conv is self.net[bi][3]

    def _launch(self, pl, packed, main):
        """One pyramid.  The site lists, index volumes and neighbour tables depend on the voxel coordinates only, so
        they run on a side stream ahead of the convolutions (which chain through the features on `main`): in the
        captured graph the two chains are parallel branches, joined table by table."""
        lib = _lib.load()
        dims, caps, counts, coords_l, idx_vol = pl["dims"], pl["caps"], pl["counts"], pl["coords"], pl["idx_vol"]
        side = pl["side"]
        st = C.c_void_p(main.cuda_stream)
        st_s = C.c_void_p(side.cuda_stream)
        n0, cols = pl["coord_in"].shape
        check(lib.gpnerf_sc_index_input(ptr(pl["coord_in"]), cols, n0, *dims[0], ptr(idx_vol[0]), ptr(pl["owners"]),
                                        ptr(coords_l[0]), ptr(counts[0:1]), ptr(pl["ws"]), st), "sc_index_input")
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        c_in = pl["feat_in"].shape[1]
        tc = self.precision == "tf32x3"
        gather = lib.gpnerf_sc_gather_rows_split if tc else lib.gpnerf_sc_gather_rows
        check(gather(ptr(pl["feat_in"]), c_in, ptr(pl["owners"]), ptr(counts[0:1]), n0, ptr(pl["x0"]), st), "sc_gather_rows")
        # ---- geometry (side stream): per layer its neighbour table and the event that marks it ready
        tables, level, subm = {}, 0, None
        for li, (bi, conv, _bn) in enumerate(self._layers()):
            if conv.stride == 2:
                check(lib.gpnerf_sc_strided_sites(ptr(coords_l[level]), ptr(counts[level:level + 1]), caps[level],
                                                  *dims[level + 1], ptr(pl["lin"]), ptr(coords_l[level + 1]),
                                                  ptr(idx_vol[level + 1]), ptr(counts[level + 1:level + 2]), ptr(pl["ws"]),
                                                  st_s), "sc_strided_sites")
                in_level, out_level, subm = level, level + 1, None
            else:
                in_level = out_level = level
            if conv.stride == 2 or subm is None:
                nbr = pl["nbr"][li]
                check(lib.gpnerf_sc_neighbours(ptr(coords_l[out_level]), ptr(counts[out_level:out_level + 1]),
                                               caps[out_level], conv.stride, ptr(idx_vol[in_level]), *dims[in_level],
                                               ptr(counts[in_level:in_level + 1]), ptr(nbr), st_s), "sc_neighbours")
                ready = torch.cuda.Event()
                ready.record(side)
                entry = (nbr, ready, out_level)
                if conv.stride == 1:
                    subm = entry
            else:
                entry = (subm[0], None, out_level)              # second convolution of a double_conv: same table
            tables[li] = entry
            level = out_level
        # ---- convolutions (main stream)
        x, outs = pl["x0"], []
        for li, (bi, conv, _bn) in enumerate(self._layers()):
            w, scale, shift, w_tc = packed[li]
            nbr, ready, out_level = tables[li]
            if ready is not None:
                main.wait_event(ready)
            y = pl["y"][li]
            y_full = pl["y_full"].get(li)
            if tc:
                check(lib.gpnerf_sc_conv_tc(ptr(x), conv.c_in, ptr(nbr), ptr(counts[out_level:out_level + 1]),
                                            caps[out_level], ptr(w_tc), ptr(scale), ptr(shift), conv.c_out, ptr(y),
                                            ptr(y_full), st), "sc_conv_tc")
            else:
                check(lib.gpnerf_sc_conv(ptr(x), conv.c_in, ptr(nbr), ptr(counts[out_level:out_level + 1]), caps[out_level],
                                         ptr(w), ptr(scale), ptr(shift), conv.c_out, ptr(y), st), "sc_conv")
            x = y
            if self._closes_level(bi, conv):
                rows = y_full if tc else x
                outs.append((rows.view(caps[out_level], conv.c_out), coords_l[out_level].view(caps[out_level], 3)))
        return outs

    @torch.no_grad()
    def forward(self, features, coords, spatial_shape):
        """features [N, in_dim] fp32, coords [N, 3|4] int (…, d, h, w), spatial_shape (D, H, W) →
        (levels_sparse, level_dims, n_rows_dev): per level (features [cap, 32], coords [cap, 3]) with
        `cap` = capacity and the live row counts in four device int32 scalars (no host sync).  The results
        live in buffers of the module (one set per input geometry) and are overwritten by the next call."""
        if self.training:
            raise _lib.GpnerfError("SparseConvNet runs in inference form (BatchNorm running statistics): call .eval()")
        dev = features.device
        if dev.type != "cuda":
            raise _lib.GpnerfError("gpnerf_b200 runs on CUDA devices only (no CPU fallback)")
        packed = self._fold(dev)
        pl = self._plan(int(coords.shape[0]), int(coords.shape[1]), spatial_shape, int(features.shape[1]), dev)
        pl["feat_in"].copy_(features.detach(), non_blocking=True)
        pl["coord_in"].copy_(coords.detach(), non_blocking=True)
        pl["uses"] += 1
        if self.use_cuda_graph and pl["graph"] is None and pl["uses"] >= 2 and not torch.cuda.is_current_stream_capturing():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                pl["outs"] = self._launch(pl, packed, torch.cuda.current_stream(dev))
            pl["graph"] = g
        if pl["graph"] is not None:
            pl["graph"].replay()
        else:
            pl["outs"] = self._launch(pl, packed, torch.cuda.current_stream(dev))
        return pl["outs"], pl["dims"][1:], [pl["counts"][k:k + 1] for k in range(1, 5)]
