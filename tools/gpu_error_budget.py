#!/usr/bin/env python
"""Error budget of the tensor-core (bf16) path at BASELINE configs[1] and friends (run on the GPU box):

  python tools/gpu_error_budget.py [--out gpurun_out/error_budget.json]

For each case renders the frame three ways – exact fp32 engine, fp32 gathers + tcgen05 bf16 heads
(`fused_gather=False`), fused 16-bit gather + tcgen05 heads (the benchmarked path) – and compares each with
the CPU oracle over the mask_at_box pixels only (libs/evaluators/if_nerf.py:49-57): max |d|, rms, PSNR,
PSNR delta against a 30 dB pseudo ground truth, survivor-list differences.  The difference between the two
bf16 rows is the share of the 16-bit volume storage + HFMA2 interpolation."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpnerf_b200  # noqa: E402,F401
import gpnerf_oracle as orc  # noqa: E402
import stages  # noqa: E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200._lib import PREC_BF16, PREC_FP32  # noqa: E402

CASES = [  # name, H, V, S, scene seed, weight seed, random bias, neg_ray
    ("configs1_512_v3_s64_seed42", 512, 3, 64, 42, 42, False, False),
    ("192_v3_s64_seed29", 192, 3, 64, 29, 129, True, False),
    ("128_v4_s128_seed7", 128, 4, 128, 7, 107, True, False),
    ("thuman_like_neg_ray_160", 160, 3, 64, 31, 131, True, True),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "error_budget.json"))
    ap.add_argument("--cases", default="")
    args = ap.parse_args()
    res = {}
    for name, H, V, S, seed, wseed, bias, neg in CASES:
        if args.cases and name not in args.cases.split(","):
            continue
        scene = synth.make_scene("zju", H=H, W=H, V=V, seed=seed)
        if neg:
            scene = synth.flip_cameras(scene)
        w = synth.make_head_weights(V=V, seed=wseed, random_bias=bias)
        o = orc.render_progressive(scene, w, S=S, chunk=131072, keep=True, neg_ray=neg)
        ref = o["pred_img"]
        mask = o["mask_at_box"]
        rows = {}
        for tag, prec, fused in (("fp32_exact", PREC_FP32, True), ("fp32_gather+bf16_heads", PREC_BF16, False),
                                 ("fused_16bit_gather+bf16_heads", PREC_BF16, True)):
            eng, _ = stages.run_engine_progressive(scene, w, S, precision=prec, neg_ray=neg, fused_gather=fused)
            c = eng.read_counters()
            img = eng.pred_img.cpu().view(H, H, 3).double()
            st = stages.masked_image_stats(img, ref, mask)
            st["psnr_delta_vs_30dB_pseudo_gt"] = stages.psnr_delta_vs_pseudo_gt(img, ref, mask)
            st["psnr_full_frame"] = orc.psnr(img, ref)
            st["counts"] = c
            st["rays_equal"] = bool(c["n_rays"] == o["n_rays"] and
                                    torch.equal(eng.ray_pix[: c["n_rays"]].cpu().long(), o["ray_pix"].long()))
            st["valid_equal"] = bool(c["P1"] == o["P1"] and torch.equal(eng.valid[: c["P1"]].cpu().long(), o["valid"]))
            diff = np.setxor1d(eng.valid1[: c["P2"]].cpu().numpy(), o["valid1"].numpy())
            st["valid1_xor"] = int(len(diff))
            st["valid1_xor_max_abs_sigma"] = float(o["sigma"][torch.from_numpy(diff).long()].abs().max()) if len(diff) else 0.0
            p1 = c["P1"]
            if st["valid_equal"]:
                sg = eng.sigma[:p1].cpu()
                st["sigma_max_abs"] = float((sg - o["sigma"]).abs().max())
                st["sigma_rms"] = float((sg - o["sigma"]).pow(2).mean().sqrt())
                st["sigma_ref_rms"] = float(o["sigma"].pow(2).mean().sqrt())
            rows[tag] = st
            del eng
            torch.cuda.empty_cache()
        res[name] = {"oracle": {"n_rays": o["n_rays"], "P1": o["P1"], "P2": o["P2"]}, "paths": rows}
        print(name, json.dumps(rows, indent=1), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
