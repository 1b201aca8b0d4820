"""Import the *real* reference modules from /root/reference (build container
only – the path does not exist on the GPU box, and nothing under tests -m gpu,
smoke() or bench.py touches this file).

The reference needs `spconv`, `mcubes`, `trimesh` (and `yacs` for configs),
none of which are installed and none of which the hot path's arithmetic uses;
they are replaced by empty stand-ins so that `BaseRender`, `demo_render` and
`trainhead` import and their pure-torch functions can be *called* to produce
golden vectors (oracle/gen_golden.py).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("GPNERF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "libs", "renders"))


def _install_stubs():
    import torch.nn as nn

    if "spconv" not in sys.modules:
        sp = types.ModuleType("spconv")

        class _Conv(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        class SparseConvTensor:  # only constructed by code paths we do not call
            def __init__(self, features, indices, spatial_shape, batch_size):
                self.features, self.indices = features, indices
                self.spatial_shape, self.batch_size = spatial_shape, batch_size

        sp.SparseSequential = nn.Sequential
        sp.SubMConv3d = _Conv
        sp.SparseConv3d = _Conv
        sp.SparseConvTensor = SparseConvTensor
        sys.modules["spconv"] = sp
    for name in ("mcubes", "trimesh"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)


def load():
    """Returns (BaseRender, demo_render, trainhead) reference modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    for sub in ("", "libs/renders", "libs/nerfheads"):
        p = os.path.join(REF_ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib

    base = importlib.import_module("BaseRender")
    demo = importlib.import_module("demo_render")
    head = importlib.import_module("trainhead")
    return base, demo, head
