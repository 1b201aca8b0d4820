"""Golden vectors for the SMPL-code attention (row f1, K8): runs the reference's own
libs/nerfheads/networks/MultiHeadAttention.py (pure torch; loaded by file path because the package's
__init__ imports spconv) in the configuration trainhead.py:35-36 builds, on seeded inputs, and stores
weights + inputs + output in tests/golden/attention.npz.  Build container only (needs /root/reference)."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GPNERF_REFERENCE", "/root/reference")
spec = importlib.util.spec_from_file_location(
    "ref_mha", os.path.join(REF, "libs", "nerfheads", "networks", "MultiHeadAttention.py"))
mha = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mha)

out = {}
for tag, (code_dim, n_head, kv, V, n, seed) in {"a": (16, 4, 32, 3, 257, 5), "b": (16, 4, 32, 5, 64, 6),
                                                 "c": (32, 4, 32, 2, 100, 7)}.items():
    torch.manual_seed(seed)
    m = mha.MultiHeadAttention(n_head, code_dim, code_dim // n_head, code_dim // n_head, kv_dim=kv, sum=False).eval()
    code = torch.randn(n, code_dim)
    feats = torch.randn(n, V, kv) * 2.0
    with torch.no_grad():
        y = m(code.unsqueeze(1), feats, feats)[0].squeeze(1)
    for k, v in m.state_dict().items():
        out[f"{tag}.state.{k}"] = v.numpy()
    out[f"{tag}.code"], out[f"{tag}.feats"], out[f"{tag}.out"] = code.numpy(), feats.numpy(), y.numpy()
    out[f"{tag}.n_head"] = np.int64(n_head)
path = os.path.join(ROOT, "tests", "golden", "attention.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
