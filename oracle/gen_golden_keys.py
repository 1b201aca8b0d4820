"""state_dict keys and shapes of the reference's NeRFHead (libs/nerfheads/trainhead.py) for the two code_dim
settings of the configs, written to tests/golden/head_state_keys.json.  Build container only.

spconv is absent: the stand-in convolution classes below carry what spconv 1.2.1's SubMConv3d / SparseConv3d
register – one `weight` parameter of shape [k, k, k, in, out] and no bias (the reference passes bias=False,
SparseConvNet.py:22-87) – so the key list is the one a real checkpoint has."""
import json
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GPNERF_REFERENCE", "/root/reference")


class _SpConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__()
        k = kernel_size
        self.weight = nn.Parameter(torch.zeros(k, k, k, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))


sp = types.ModuleType("spconv")
sp.SparseSequential = nn.Sequential
sp.SubMConv3d = _SpConv
sp.SparseConv3d = _SpConv
sp.SparseConvTensor = object
sys.modules["spconv"] = sp
for sub in ("", "libs/nerfheads"):
    sys.path.insert(0, os.path.join(REF, sub))
import trainhead  # noqa: E402

out = {}
for code_dim in (16, 32):
    head = trainhead.NeRFHead(code_dim=code_dim)
    out[str(code_dim)] = [(k, list(v.shape)) for k, v in head.state_dict().items()]
    print(code_dim, len(out[str(code_dim)]), "entries")
path = os.path.join(ROOT, "tests", "golden", "head_state_keys.json")
json.dump(out, open(path, "w"))
print(path, os.path.getsize(path), "bytes")
