"""CPU: the oracle restatement against the golden vectors produced by the real
reference code (oracle/gen_golden.py, run in the build container)."""
import os

import numpy as np
import pytest
import torch

import gpnerf_oracle as orc
from gpnerf_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="module")
def fn():
    return load("functions.npz")


def weights_of(g):
    return {k[2:]: v for k, v in g.items() if k.startswith("w.")}


def test_sampling_bit_exact(fn):
    S = fn["z"].shape[1]
    pts, z = orc.sampling_points(fn["ray_o"], fn["ray_d"], fn["near"], fn["far"], S)
    assert torch.equal(pts, fn["pts"]) and torch.equal(z, fn["z"])
    pts, z = orc.sampling_points(fn["ray_o"], fn["ray_d"], fn["near"], fn["far"], S, fn["t_rand"])
    assert torch.equal(pts, fn["pts_jit"]) and torch.equal(z, fn["z_jit"])


def test_frames_and_grid_coords_bit_exact(fn):
    can = orc.pts_to_can_pts(fn["pts"].reshape(-1, 3), fn["R"], fn["Th"])
    assert torch.equal(can, fn["can"].reshape(-1, 3))
    grid = orc.grid_coords_of(can, fn["bounds"], fn["out_sh"].tolist())
    assert torch.equal(grid, fn["grid"].reshape(-1, 3))


def test_projector(fn):
    rgb_feat, mask = orc.projector_compute(fn["pts"].reshape(-1, 3), fn["imgs01"], fn["cams"], fn["featmaps"])
    assert torch.equal(mask.view(fn["mask"].shape), fn["mask"])
    assert float((rgb_feat.view(fn["rgb_feat"].shape) - fn["rgb_feat"]).abs().max()) < 1e-6
    _, mneg = orc.projector_compute(fn["pts"].reshape(-1, 3), fn["imgs01"], fn["cams"], fn["featmaps"], neg_ray=True)
    assert torch.equal(mneg.view(fn["mask_neg"].shape), fn["mask_neg"])


def test_heads(fn):
    w = weights_of(fn)
    R, S, V, Cf = fn["rgb_feat"].shape
    mean, var = orc.mean_var(fn["rgb_feat"].view(-1, V, Cf))
    assert float((mean - fn["mean"].view(-1, Cf)).abs().max()) < 1e-6
    assert float((var - fn["var"].view(-1, Cf)).abs().max()) < 1e-6
    sfeat = orc.sigma_feat_of(fn["vol_feat"], w)
    assert float((sfeat - fn["sigma_feat"]).abs().max()) < 1e-6
    sigma = orc.density_mlp(sfeat, mean, var, fn["mask"].view(-1, V), w)
    assert float((sigma - fn["sigma_out"].view(-1)).abs().max()) < 1e-6
    rgb = orc.color_mlp(fn["rgb_feat"].view(-1, V, Cf), mean, var, w)
    assert float((rgb - fn["rgb_out"].view(-1, 3)).abs().max()) < 1e-6


def test_raw2outputs(fn):
    raw = torch.cat([fn["rgb_out"], fn["sigma_out"]], -1)
    rgb_map, disp, acc, weights, depth = orc.raw2outputs(raw, fn["z"], False)
    for a, b in ((rgb_map, fn["rgb_map"]), (disp, fn["disp"]), (acc, fn["acc"]), (weights, fn["weights"]),
                 (depth, fn["depth"])):
        assert float((a - b).abs().max()) < 1e-6
    rgb_map, _, _, weights, depth = orc.raw2outputs(raw, fn["z"], True)
    assert float((rgb_map - fn["rgb_map_neg"]).abs().max()) < 1e-6
    assert float((weights - fn["weights_neg"]).abs().max()) < 1e-6
    assert float((depth - fn["depth_neg"]).abs().max()) < 1e-6


@pytest.mark.parametrize("tag", ["mini", "mini_s64"])
def test_whole_path_matches_reference_run(tag):
    g = load(f"whole_{tag}.npz")
    H, S, seed = int(g["H"]), int(g["S"]), int(g["seed"])
    scene = synth.make_scene("zju", H=H, W=H, V=3, seed=seed)
    from hashlib import sha256

    def sha(t):
        return sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()[:16]
    got = (sha(scene["levels"][0]) + sha(scene["featmaps"]) + sha(scene["src_imgs"])).encode()
    assert bytes(g["input_sha"].numpy().tobytes()) == got, "synthetic inputs drifted (torch RNG changed?)"
    o = orc.render_progressive(scene, weights_of(g), S=S, keep=True)
    assert torch.equal(o["can_bounds"], g["can_bounds"])
    assert torch.equal(o["pix_idx"].int(), g["pix_idx"])
    assert torch.equal(o["ray_pix"].int(), g["ray_pix"])
    assert torch.equal(o["near"], g["near"]) and torch.equal(o["far"], g["far"])
    assert torch.equal(o["valid"].int(), g["valid"])
    assert torch.equal(o["valid1"].int(), g["valid1"])
    assert float((o["sigma"] - g["sigma"]).abs().max()) < 1e-6
    assert float((o["rgb_map"] - g["rgb_map"]).abs().max()) < 1e-6


def test_chunked_oracle_is_order_identical():
    scene = synth.make_scene("zju", H=64, W=64, V=3, seed=9)
    w = synth.make_head_weights(V=3, seed=9)
    a = orc.render_progressive(scene, w, S=16, keep=True)
    b = orc.render_progressive(scene, w, S=16, chunk=1000, keep=True)
    assert torch.equal(a["valid"], b["valid"]) and torch.equal(a["valid1"], b["valid1"])
    assert float((a["rgb_map"] - b["rgb_map"]).abs().max()) < 1e-6


def test_empty_volume_gives_no_rays():
    scene = synth.make_scene("zju", H=32, W=32, V=3, seed=1)
    scene["levels"] = [torch.zeros_like(t) for t in scene["levels"]]
    w = synth.make_head_weights(V=3, seed=1)
    masks3d, mask_xyz = orc.build_masks3d(scene["levels"])
    assert mask_xyz.shape[0] == 0 and float(masks3d.max()) == 0.0


def test_dataset_rays_restatement_vs_reference_numpy():
    """Row f3: oracle.dataset_rays against the outputs of the reference's own
    get_rays / get_near_far (oracle/gen_golden_rays.py), bit for bit."""
    z = np.load(os.path.join(GOLD, "dataset_rays.npz"))
    for tag in ("a", "b"):
        H = int(z[f"{tag}.H"])
        o, d, near, far, mask = orc.dataset_rays(H, H, z[f"{tag}.K"], z[f"{tag}.R"], z[f"{tag}.T"], z[f"{tag}.bounds"])
        assert np.array_equal(mask, z[f"{tag}.mask_at_box"]) and mask.sum() > 100
        for got, key in ((o, "ray_o"), (d, "ray_d"), (near, "near"), (far, "far")):
            assert got.dtype == np.float32 and np.array_equal(got, z[f"{tag}.{key}"]), key


def test_attention_restatement_vs_reference_module():
    """Row f1 (K8): oracle.smpl_code_attention against outputs of the reference's own MultiHeadAttention
    module (oracle/gen_golden_attn.py).  fp32 module vs fp64 restatement: 1e-5."""
    z = np.load(os.path.join(GOLD, "attention.npz"))
    for tag in ("a", "b", "c"):
        state = {k.split(".state.")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}.state.")}
        got = orc.smpl_code_attention(state, torch.from_numpy(z[f"{tag}.code"]), torch.from_numpy(z[f"{tag}.feats"]),
                                      n_head=int(z[f"{tag}.n_head"]))
        want = torch.from_numpy(z[f"{tag}.out"])
        assert got.shape == want.shape and float(want.abs().max()) > 0.1
        assert float((got - want).abs().max()) < 1e-5
