"""ncu target: one pass of the native upstream chain (K2 SMPL gather, K8, K7) at configs[1] geometry."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200.nerfhead import NeRFHead  # noqa: E402

DEV = "cuda:0"
scene = synth.make_scene("zju", H=512, W=512, V=3, seed=42)
head = NeRFHead(n_views=3, precision=1).eval()
sd = head.state_dict()
for k, v in sd.items():
    if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
        v.fill_(3.0)
head.load_state_dict(sd)
head = head.to(DEV)
sh = head.sigmahead
feats = torch.randn(6890, 3, 32, device=DEV)
code = sh.c.weight.detach().unsqueeze(1)
out_sh = [int(v) for v in scene["out_sh"][0]]
coord = scene["coord"][0].to(DEV)
for _ in range(int(os.environ.get("REPS", "2"))):
    fused = sh.xyzc_attn(code, feats, feats)[0].squeeze(1)
    rows, dims, n_dev = sh.xyzc_net(fused, coord, out_sh)
torch.cuda.synchronize()
print([int(n) for n in n_dev])
