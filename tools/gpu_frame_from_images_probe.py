"""ncu target: two eager frames from nothing but source images + SMPL fit (encoder, K8, K7, K0…K5), no CUDA graphs,
so that every kernel is its own line in the launch list."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200._lib import PREC_BF16  # noqa: E402
from gpnerf_b200.encoder import ResUNet  # noqa: E402
from gpnerf_b200.nerfhead import NeRFHead  # noqa: E402
from gpnerf_b200.render import Renderer  # noqa: E402

DEV = "cuda:0"
torch.manual_seed(42)
scene = synth.make_scene("zju", H=512, W=512, V=3, seed=42)
head = NeRFHead(n_views=3, precision=PREC_BF16).eval()
sd = head.state_dict()
for k, v in synth.make_head_weights(V=3, seed=42).items():
    sd[k].copy_(v)
for k, v in sd.items():
    if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
        v.fill_(3.0)
head.load_state_dict(sd)
head.sigmahead.xyzc_net.use_cuda_graph = False
enc = synth.fill_encoder_params(ResUNet(use_cuda_graph=False), seed=42).eval().to(DEV)
r = Renderer(enc, head.to(DEV), is_train=False, n_samples=64, progressive=True, precision=PREC_BF16, use_cuda_graph=False)
batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in scene.items() if k not in ("levels", "featmaps")}
for _ in range(2):
    out = r.render(dict(batch))
print(out["counts"])
