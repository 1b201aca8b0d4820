"""Where a blocking Renderer.render(batch) call spends its wall time (1 GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import gpnerf_b200  # noqa
from gpnerf_b200 import synth
from gpnerf_b200._lib import PREC_BF16
from gpnerf_b200.nerfhead import NeRFHead
from gpnerf_b200.render import Renderer

dev = torch.device("cuda", 0)
scene = synth.make_scene("zju", H=512, W=512, V=3, seed=42)
w = synth.make_head_weights(V=3, seed=42)
head = NeRFHead(code_dim=32, n_views=3, precision=PREC_BF16)
sd = head.state_dict(); sd.update(w); head.load_state_dict(sd)
r = Renderer(None, head.to(dev), is_train=False, n_samples=64, progressive=True, precision=PREC_BF16)
batch = {k: v for k, v in scene.items() if torch.is_tensor(v)}
if os.environ.get("SPARSE", "1") == "1":          # levels as sparse rows (4 MB) instead of dense NCDHW tensors (119 MB)
    lv_s, dims_s = synth.sparsify_levels(scene["levels"])
    batch["levels_sparse"] = [(f.pin_memory(), i.pin_memory()) for f, i in lv_s]
    batch["level_dims"] = dims_s
else:
    batch["levels"] = [t.pin_memory() for t in scene["levels"]]
batch["featmaps"] = scene["featmaps"].pin_memory()
batch["src_imgs"] = scene["src_imgs"].pin_memory()

def step():
    b = dict(batch); b["src_imgs"] = batch["src_imgs"].to(dev, non_blocking=True)
    return r.render(b)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("blocking render ms/frame", (time.perf_counter() - t0) / 20 * 1e3)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

# ---- render_stream: per-frame arrival times, then a profile of the main thread
sb = dict(batch)
for _ in r.render_stream(sb for _ in range(4)):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
stamps = []
for out in r.render_stream(sb for _ in range(20)):
    stamps.append((time.perf_counter() - t0) * 1e3)
print("render_stream ms/frame", stamps[-1] / 20, "arrivals", [round(s, 1) for s in stamps])
pr = cProfile.Profile(); pr.enable()
for out in r.render_stream(sb for _ in range(20)):
    pass
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
