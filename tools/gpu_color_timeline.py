"""Debug build only (GPNERF_DEBUG_COLOR=1): cycle stamps of one colour-chain tile (block 0, chain 0, its second tile)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import gpnerf_b200  # noqa
from gpnerf_b200 import synth
from gpnerf_b200._lib import PREC_BF16
import stages
scene = synth.make_scene("zju", H=512, W=512, V=3, seed=42)
w = synth.make_head_weights(V=3, seed=42)
eng, _ = stages.run_engine_progressive(scene, w, 64, precision=PREC_BF16)
buf = (C.c_longlong * 512)()
eng.lib.gpnerf_debug_color_t(buf)
t = list(buf)
names = ["epi1(64)", "epi2", "epi3", "epi4"]
print("round: leader-sees->bar.sync+fence | ld32+wait | ELU+pack+st issue | (2nd half / wait::st) | fence+arrive | arrive -> next leader-sees")
for r in range(13):
    a = t[r * 8: r * 8 + 6]
    nxt = t[(r + 1) * 8]
    print(f"{r:2d} {names[r % 4]:8s} {a[1]-a[0]:6d} | {a[2]-a[1]:6d} | {a[3]-a[2]:6d} | {a[4]-a[3]:6d} | {a[5]-a[4]:6d} | {nxt-a[5]:6d}   epilogue total {a[5]-a[0]}")
