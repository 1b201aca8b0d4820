"""compute-sanitizer target: one small tensor-core progressive frame.  compute-sanitizer --tool memcheck python tools/gpu_san_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import gpnerf_b200  # noqa
from gpnerf_b200 import synth
from gpnerf_b200._lib import PREC_BF16
import stages
V = int(os.environ.get("V", "3"))
H = int(os.environ.get("H", "72"))
scene = synth.make_scene("zju", H=H, W=H, V=V, seed=81)
w = synth.make_head_weights(V=V, seed=141, random_bias=True)
eng, _ = stages.run_engine_progressive(scene, w, 40, precision=PREC_BF16, tile_px=48)
print("progressive", eng.read_counters(), float(eng.pred_img.sum()))
