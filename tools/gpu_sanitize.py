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
print("sanitize target done")
