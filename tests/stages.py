"""Stage-by-stage comparison of the CUDA path (through the C ABI) with the CPU
oracle on one scene.  Used by the -m gpu parity tests and by
tools/gpu_stage_report.py."""
import numpy as np
import torch

import gpnerf_oracle as orc
from gpnerf_b200 import synth
from gpnerf_b200._lib import PREC_FP32
from gpnerf_b200.engine import Engine


def to_dev(scene, device):
    out = {}
    for k, v in scene.items():
        if torch.is_tensor(v):
            out[k] = v.to(device)
        elif isinstance(v, list) and v and torch.is_tensor(v[0]):
            out[k] = [t.to(device) for t in v]
        else:
            out[k] = v
    return out


def run_engine_progressive(scene, weights, S, device="cuda:0", precision=PREC_FP32, t_min=0.0,
                           rank=0, world=1, tile_px=64, neg_ray=False, fused_gather=True):
    eng = Engine(scene["H"], scene["W"], S, scene["V"], device=device, precision=precision, t_min=t_min,
                 rank=rank, world=world, tile_px=tile_px, fused_gather=fused_gather)
    eng.set_weights(weights)
    d = to_dev(scene, device)
    eng.upload_products(d["levels"], d["featmaps"], d["src_imgs"])
    frame = eng.make_frame(scene, neg_ray=neg_ray)
    eng.render_progressive(frame)
    torch.cuda.synchronize()
    return eng, frame


def xor_count(a, b):
    return int(len(np.setxor1d(np.asarray(a), np.asarray(b))))


def masked_image_stats(img, ref, mask_at_box):
    """Image error over the `mask_at_box` pixels only – the pixels the reference's evaluator scores
    (libs/evaluators/if_nerf.py:49-57); the background is exactly 0 in both images and would dilute every figure."""
    m = torch.as_tensor(mask_at_box).reshape(-1).bool()
    a, b = img.reshape(-1, 3)[m].double(), ref.reshape(-1, 3)[m].double()
    if a.numel() == 0:
        return {"n_px": 0, "max_abs": 0.0, "rms": 0.0, "psnr_mask": 200.0}
    d = a - b
    return {"n_px": int(m.sum()), "max_abs": float(d.abs().max()), "rms": float(d.pow(2).mean().sqrt()),
            "psnr_mask": orc.psnr_masked(img, ref, m)}


def psnr_delta_vs_pseudo_gt(img_test, img_ref, mask_at_box, gt_db=30.0):
    """north_star: PSNR delta < 0.05 dB with bf16 MLPs.  Random-init weights have no ground truth, so one is
    synthesised: the reference render plus noise that puts the reference render at `gt_db` – the range trained
    GP-NeRF models score in – both PSNRs taken over the mask_at_box pixels as the evaluator does."""
    g = torch.Generator().manual_seed(0)
    gt = img_ref + torch.randn(img_ref.shape, generator=g, dtype=img_ref.dtype) * 10 ** (-gt_db / 20)
    return abs(orc.psnr_masked(img_test, gt, mask_at_box) - orc.psnr_masked(img_ref, gt, mask_at_box))


def compare_progressive(scene, weights, S, precision=PREC_FP32, t_min=0.0, oracle_out=None, neg_ray=False):
    """Returns (report dict, engine, oracle_out)."""
    o = oracle_out if oracle_out is not None else orc.render_progressive(scene, weights, S=S, keep=True,
                                                                         neg_ray=neg_ray)
    eng, _ = run_engine_progressive(scene, weights, S, precision=precision, t_min=t_min, neg_ray=neg_ray)
    c = eng.read_counters()
    rep = {"counts_gpu": c, "counts_oracle": {"n_pix": int(o["pix_idx"].numel()), "n_rays": o["n_rays"],
                                              "P1": o["P1"], "P2": o["P2"]}}
    m3 = eng.masks3d.cpu().view_as(o["masks3d"])
    rep["masks3d_maxabs"] = float((m3 - o["masks3d"]).abs().max())
    rep["masks3d_thr_xor"] = int(((m3 > 0.1) != (o["masks3d"] > 0.1)).sum())
    rep["can_bounds_ne"] = int((eng.can_bounds[:6].cpu() != o["can_bounds"].flatten()).sum())
    pm = eng.pix_mask.cpu()
    rep["pix_mask_ne"] = int((pm != o["pix_mask"]).sum())
    n = c["n_rays"]
    ray_pix = eng.ray_pix[:n].cpu()
    rep["ray_pix_xor"] = xor_count(ray_pix, o["ray_pix"])
    rep["ray_order_ok"] = bool(torch.equal(ray_pix.long(), o["ray_pix"].long()))
    if rep["ray_order_ok"]:
        rep["rays_d_ne"] = int((eng.rays_d[: n * 3].cpu().view(n, 3) != o["rays_d"]).sum())
        rep["near_ne"] = int((eng.near[:n].cpu() != o["near"]).sum())
        rep["far_ne"] = int((eng.far[:n].cpu() != o["far"]).sum())
        rep["z_ne"] = int((eng.z_vals[: n * S].cpu().view(n, S) != o["z_vals"]).sum())
        p1 = c["P1"]
        valid = eng.valid[:p1].cpu()
        rep["valid_xor"] = xor_count(valid, o["valid"])
        rep["valid_order_ok"] = bool(torch.equal(valid.long(), o["valid"].long()))
        if rep["valid_order_ok"]:
            rep["vol_feat_maxabs"] = float((eng.vol_feat[: p1 * 128].cpu().view(p1, 128) - o["vol_feat"]).abs().max())
            V = scene["V"]
            rep["rgb_feat_maxabs"] = float((eng.rgb_feat[: p1 * V * 35].cpu().view(p1, V, 35) - o["rgb_feat"]).abs().max())
            rep["mask_ne"] = int((eng.mask[: p1 * V].cpu().view(p1, V) != o["mask"]).sum())
            mv = eng.meanvar[: p1 * 70].cpu().view(p1, 70)
            rep["mean_maxabs"] = float((mv[:, :35] - o["mean"]).abs().max())
            rep["var_maxabs"] = float((mv[:, 35:] - o["var"]).abs().max())
            sig = eng.sigma[:p1].cpu()
            rep["sigma_maxabs"] = float((sig - o["sigma"]).abs().max())
            rep["sigma_maxrel"] = float(((sig - o["sigma"]).abs() / (o["sigma"].abs() + 1e-3)).max())
            p2 = c["P2"]
            valid1 = eng.valid1[:p2].cpu()
            rep["valid1_xor"] = xor_count(valid1, o["valid1"])
            if rep["valid1_xor"] == 0 and p2:
                rgb = eng.rgb[: p1 * 3].cpu().view(p1, 3)[valid1.long()]
                rep["rgb_maxabs"] = float((rgb - o["rgb"]).abs().max())
        rep["rgb_map_maxabs"] = float((eng.rgb_map[: n * 3].cpu().view(n, 3) - o["rgb_map"]).abs().max())
        img = eng.pred_img.cpu().view(scene["H"], scene["W"], 3)
        rep["pred_img_maxabs"] = float((img.double() - o["pred_img"]).abs().max())
        rep["psnr_vs_oracle"] = orc.psnr(img, o["pred_img"])
        rep["hit_mask_ne"] = int((eng.hit_mask.cpu().bool() != o["mask_at_box"]).sum())
    return rep, eng, o
