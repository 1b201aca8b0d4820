"""compute-sanitizer target: one small progressive frame on each path (fp32
exact kernels, bf16 fused tcgen05 kernels, dense path).  Run as
  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/gpu_sanitize.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200._lib import PREC_BF16, PREC_FP32  # noqa: E402
from gpnerf_b200.engine import Engine  # noqa: E402
import stages  # noqa: E402

scene = synth.make_scene("zju", H=64, W=64, V=3, seed=3, with_rays=True)
w = synth.make_head_weights(V=3, seed=3, random_bias=True)
for prec in (PREC_FP32, PREC_BF16):
    eng, _ = stages.run_engine_progressive(scene, w, 16, precision=prec)
    print("progressive", prec, eng.read_counters(), float(eng.pred_img.sum()))
    R = 300
    rays = tuple(scene[k][0][:R] for k in ("ray_o", "ray_d", "near", "far"))
    e2 = Engine(64, 64, 16, 3, device="cuda:0", max_rays=R, precision=prec)
    e2.set_weights(w)
    d = stages.to_dev(scene, "cuda:0")
    e2.upload_products(d["levels"], d["featmaps"], d["src_imgs"])
    out = e2.render_dense(e2.make_frame(scene), *rays)
    torch.cuda.synchronize()
    print("dense", prec, float(out["rgb_map"].sum()))
# peer exchange (one rank: its own IPC buffer), graph-free; TF32 training step; dataset rays
from gpnerf_b200 import ops  # noqa: E402
from gpnerf_b200.peer import PeerExchange  # noqa: E402
from gpnerf_b200.train import render_dense_autograd  # noqa: E402
import numpy as np  # noqa: E402
e3 = Engine(64, 64, 16, 3, device="cuda:0", precision=PREC_BF16)
e3.set_weights(w)
ex = PeerExchange(64, 64, "cuda:0", 0, 1, mode="tiles")
e3.attach_exchange(ex)
e3.upload_products(d["levels"], d["featmaps"], d["src_imgs"])
for _ in range(2):
    e3.render_progressive(e3.make_frame(scene))
torch.cuda.synchronize()
print("exchange", float(e3.result_image().sum()))
ex.close()
R = 200
rays = tuple(scene[k][0][:R].to("cuda:0") for k in ("ray_o", "ray_d", "near", "far"))
for prec in (0, 1):
    e4 = Engine(64, 64, 16, 3, device="cuda:0", max_rays=R)
    wg = {k: v.clone().to("cuda:0").requires_grad_(True) for k, v in w.items()}
    lv = [t.clone().requires_grad_(True) for t in d["levels"]]
    fm = d["featmaps"].clone().requires_grad_(True)
    e4.set_weights(w)
    e4.upload_products([t.detach() for t in lv], fm.detach(), d["src_imgs"])
    out = render_dense_autograd(e4, e4.make_frame(scene), rays, lv, fm, d["src_imgs"], wg, t_rand=torch.rand(R, 16),
                                precision=prec)
    out["rgb_map"].sum().backward()
    torch.cuda.synchronize()
    print("train", prec, float(fm.grad.abs().sum()))
pose = scene["target_pose"][0].numpy().astype(np.float64)
r = ops.dataset_rays(64, 64, scene["target_K"][0].numpy().astype(np.float64), pose[:, :3], pose[:, 3],
                     scene["can_bounds"][0].numpy(), "cuda:0")
print("dataset rays", r[0].shape)
print("sanitize target done")
# rows f1 + f2: encoder (K9), SMPL attention (K8), sparse-conv pyramid (K7), eager and as captured graphs
from gpnerf_b200.encoder import ResUNet  # noqa: E402
from gpnerf_b200.nerfhead import NeRFHead  # noqa: E402
from gpnerf_b200.render import Renderer  # noqa: E402
head = NeRFHead(n_views=3, precision=PREC_BF16).eval()
sd = head.state_dict()
for k, v in w.items():
    sd[k].copy_(v)
for k, v in sd.items():
    if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
        v.fill_(3.0)
head.load_state_dict(sd)
for prec_e in ("fp16", "fp32"):
    enc = synth.fill_encoder_params(ResUNet(precision=prec_e), seed=1).eval().to("cuda:0")
    rr = Renderer(enc, head.to("cuda:0"), is_train=False, n_samples=16, progressive=True, precision=PREC_BF16)
    base = {k: (v.to("cuda:0") if torch.is_tensor(v) else v) for k, v in scene.items() if k not in ("levels", "featmaps")}
    for _ in range(3):
        out = rr.render(dict(base))
    print("from images", prec_e, out["counts"], float(out["pred_img"].sum()))
odd = torch.rand(2, 3, 72, 56, device="cuda:0") * 2 - 1          # skip tensors that need zero padding
print("encoder odd size", tuple(enc(odd).shape))
print("sanitize target done (f1/f2)")
