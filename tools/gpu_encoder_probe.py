"""ncu target: one eager pass of the image encoder (3 × 512²)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200.encoder import ResUNet  # noqa: E402

enc = synth.fill_encoder_params(ResUNet(precision=os.environ.get("PREC", "fp16"), use_cuda_graph=False)).eval().to("cuda:0")
x = (torch.rand(3, 3, 512, 512) * 2 - 1).to("cuda:0")
for _ in range(2):
    y = enc(x)
torch.cuda.synchronize()
print(y.shape)
