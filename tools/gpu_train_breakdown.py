"""Where one training step (4096 rays x 64 samples, TF32 heads) spends its device time: per-kernel totals (1 GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from torch.profiler import profile, ProfilerActivity
import gpnerf_b200  # noqa
from gpnerf_b200 import synth, train
from gpnerf_b200.engine import Engine

dev = torch.device("cuda", 0)
R, S, V, RES = 4096, 64, 3, 512
scene = synth.make_scene("zju", H=RES, W=RES, V=V, seed=42, with_rays=True)
w0 = synth.make_head_weights(V=V, seed=42, random_bias=True)
n_all = scene["ray_o"].shape[1]
sel = (torch.arange(R) * max(1, n_all // R)) % n_all
rays = tuple(scene[k][0][sel].to(dev) for k in ("ray_o", "ray_d", "near", "far"))
eng = Engine(RES, RES, S, V, device=dev, max_rays=R)
w_g = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in w0.items()}
lv = [t.to(dev) for t in scene["levels"]]
fm, im = scene["featmaps"].to(dev), scene["src_imgs"].to(dev)
eng.set_weights(w0)
eng.upload_products(lv, fm, im)
frame = eng.make_frame(scene)
target = torch.rand(R, 3, device=dev)
bucket = train.GradBucket(w_g.values())
opt = torch.optim.AdamW(list(w_g.values()), lr=1e-4)
t_rand = torch.rand(R, S, device=dev)

def step():
    bucket.zero()
    out = train.render_dense_autograd(eng, frame, rays, lv, fm, im, w_g, t_rand=t_rand, precision=train.PREC_TRAIN_TF32)
    loss = ((out["rgb_map"] - target) ** 2).mean()
    loss.backward()
    bucket.all_reduce_mean()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record(); torch.cuda.synchronize()
print("ms/step", a.elapsed_time(b) / 5)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name[:90]
        t, c = rows.get(n, (0.0, 0))
        rows[n] = (t + e.device_time / 3.0, c + 1)
tot = sum(t for t, _ in rows.values())
print("device us per step (sum of kernels): %.1f" % tot)
for n, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%9.1f us  %5.1f %%  x%-4d %s" % (t, 100 * t / tot, c // 3, n))
