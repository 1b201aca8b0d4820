"""Debug aid: tcgen05 heads (precision=1) against the fp32 CUDA-core heads on
random inputs.  Run under gpurun with a timeout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import ops, synth  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
for V in (3, 4, 2):
    w = synth.make_head_weights(V=V, seed=5, random_bias=True)
    hw, keep = ops.pack_head_weights(w, dev, V)
    for n in (128, 1000, 40000):
        vol = torch.randn(n, 128, device=dev).clamp(min=0)
        rgb_feat = torch.randn(n, V, 35, device=dev) * 0.5
        mask = (torch.rand(n, V, device=dev) > 0.2).float()
        mv = ops.mean_variance(rgb_feat)
        s0, f0 = ops.density_mlp(vol, mv, mask, hw, 0, want_sigma_feat=True)
        s1, f1 = ops.density_mlp(vol, mv, mask, hw, 1, want_sigma_feat=True)
        torch.cuda.synchronize()
        print(f"V={V} n={n} density: max|dσ|={float((s0 - s1).abs().max()):.4f} mean|σ|={float(s0.abs().mean()):.3f} "
              f"max|dfeat|={float((f0 - f1).abs().max()):.4f}  nz0={int((s0 > 0).sum())} nz1={int((s1 > 0).sum())}")
        c0 = ops.color_mlp(rgb_feat, mv, hw, 0)
        c1 = ops.color_mlp(rgb_feat, mv, hw, 1)
        torch.cuda.synchronize()
        print(f"V={V} n={n} colour : max|drgb|={float((c0 - c1).abs().max()):.4f}")
        s2 = ops.density_mlp(f0, mv, mask, hw, 1)
        torch.cuda.synchronize()
        print(f"V={V} n={n} density(from sigma_feat): max|dσ|={float((s0 - s2).abs().max()):.4f}")
print("done")
