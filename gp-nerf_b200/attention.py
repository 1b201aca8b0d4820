"""Mirror of libs/nerfheads/networks/MultiHeadAttention.py:40-98 on csrc/k8_attention.cu.

Same constructor and parameter names (``w_qs``, ``w_ks``, ``w_vs``, ``fc``,
``layer_norm``) so that ``sigmahead.xyzc_attn.*`` of a reference checkpoint
loads with ``strict=True``.  The reference uses the module in one way only
(trainhead.py:35-36, 50: ``sum=False``, no mask, one query per SMPL vertex, the
V pixel-aligned features of that vertex as keys and values); that is the form
the kernel implements.  ``layer_norm`` is a parameter holder: with ``sum=False``
the reference never applies it either.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr


class MultiHeadAttention(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, kv_dim=None, sum=True):
        super().__init__()
        if d_k != d_v:
            raise _lib.GpnerfError("MultiHeadAttention: d_k == d_v on this path (trainhead.py:35)")
        self.n_head, self.d_k, self.d_v, self.sum_flag = n_head, d_k, d_v, sum
        kv_dim = d_model if kv_dim is None else kv_dim
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(kv_dim, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(kv_dim, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)

    @torch.no_grad()
    def forward(self, q, k, v, mask=None):
        """q [n, 1, d_model]; k = v [n, V, kv_dim] (any strides with a contiguous last axis) →
        (out [n, 1, d_model], None) – the reference's second result (the attention map) is unused by its
        callers (trainhead.py:50 takes [0])."""
        if self.sum_flag or mask is not None or k.data_ptr() != v.data_ptr() or q.shape[1] != 1:
            raise _lib.GpnerfError("MultiHeadAttention: only the reference's call form is implemented "
                                   "(sum=False, mask=None, one query per row, k is v)")
        if q.device.type != "cuda":
            raise _lib.GpnerfError("gpnerf_b200 runs on CUDA devices only (no CPU fallback)")
        lib = _lib.load()
        n, V, kv = (int(s) for s in k.shape)
        code = q.detach().reshape(n, -1).float().contiguous()
        if k.dtype != torch.float32 or k.stride(2) != 1:
            k = k.float().contiguous()
        out = torch.empty(n, code.shape[1], dtype=torch.float32, device=q.device)
        w = [m.weight.detach().float().contiguous() for m in (self.w_qs, self.w_ks, self.w_vs, self.fc)]
        st = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
        check(lib.gpnerf_attn_smpl_code(ptr(code), C.c_void_p(k.data_ptr()), k.stride(1), k.stride(0), n, V, *(ptr(t) for t in w),
                                        int(code.shape[1]), kv, self.n_head, self.d_k, ptr(out), st), "attn_smpl_code")
        return out.unsqueeze(1), None
