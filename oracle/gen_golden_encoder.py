"""Golden vectors for the image encoder (row f2): runs the reference's own libs/encoders/UNet.py ResUNet (pure
torch, CPU fp32) with seeded parameters (gpnerf_b200.synth.fill_encoder_params – 8.9 M values, regenerated from
the seed rather than stored) on two seeded image stacks and stores the outputs, plus the reference's
state_dict keys and shapes, in tests/golden/encoder.npz.  Build container only (needs /root/reference)."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
REF = os.environ.get("GPNERF_REFERENCE", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_unet", os.path.join(REF, "libs", "encoders", "UNet.py"))
unet = importlib.util.module_from_spec(spec)
spec.loader.exec_module(unet)

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402

torch.set_num_threads(os.cpu_count())
ref = synth.fill_encoder_params(unet.ResUNet(encoder="resnet34", out_ch=32), seed=42).eval()
out = {"keys": np.array(json.dumps([(k, list(v.shape)) for k, v in ref.state_dict().items()]))}
for tag, (V, H, W, seed) in {"a": (3, 64, 64, 1), "b": (2, 72, 56, 2)}.items():      # b: odd sizes → skip padding
    x = torch.rand(V, 3, H, W, generator=torch.Generator().manual_seed(seed)) * 2 - 1
    with torch.no_grad():
        y = ref(x)
    out[f"{tag}.shape"] = np.array([V, H, W, seed])
    out[f"{tag}.x_sum"] = np.float64(x.double().sum())
    out[f"{tag}.out"] = y.numpy()
    print(tag, tuple(y.shape), float(y.abs().mean()), float(y.abs().max()))
path = os.path.join(ROOT, "tests", "golden", "encoder.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
