"""CPU oracle for the GP-NeRF progressive volume-rendering hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product path (``gp-nerf_b200``) never does.

It is a torch-CPU fp32 restatement of the reference's algorithm, function by
function, each citing the ``/root/reference`` file:line it follows.  The
reference is pure PyTorch, so the restatement uses the same ATen calls where
only *values* matter (``F.grid_sample``, ``F.linear``, ``F.elu``) and spells the
arithmetic out, op by op, on every chain that feeds an integer result (pixel
mask, ray list, ``mask_at_box``, ``valid``), so that the CUDA kernels can follow
the same IEEE-754 sequence and be compared bit for bit:

* small-K matmuls (K = 3, 4) are "first product rounded, then one FMA per
  further term, k ascending" – what ATen's CPU sgemm does on these shapes
  (checked numerically, see oracle/gen_golden.py);
* elementwise torch expressions are separate roundings (no FMA contraction);
* ``x / python_float`` on CPU is a true IEEE division by the fp32 scalar.

Parity pinning: the reference has no tests, golden vectors or fixtures
(SURVEY.md §4, §8c).  The oracle is pinned instead against outputs of the
reference's own functions imported in the build container
(oracle/gen_golden.py → tests/golden/*.npz, committed).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# arithmetic helpers
# --------------------------------------------------------------------------


# SURVEY §8c, second mode: the reference really runs these chains with torch on the GPU – cuBLAS for the K = 3 / 4
# matmuls, ATen's CUDA kernels for norm and grid_sample – whose rounding differs from ATen-CPU's (to which the
# bit-exact pins above are tied).  Inside `with native_ops("cuda:0"):` the three primitives every integer result
# depends on (small matmul, 3-vector norm, grid_sample) are executed by torch on that device exactly as the reference
# issues them; everything else stays as is.  tests/test_gpu_parity.py reports how many integer results move.
_NATIVE = {"device": None}


class native_ops:
    def __init__(self, device):
        self.device = device

    def __enter__(self):
        self.prev, _NATIVE["device"] = _NATIVE["device"], self.device
        return self

    def __exit__(self, *exc):
        _NATIVE["device"] = self.prev
        return False


def _fma(a, b, c):
    """round_fp32(a*b + c) with one rounding (the product of two fp32 is exact
    in fp64; the fp64 sum is then rounded once more to fp32 – double rounding
    differs from a hardware FMA with probability ~2^-29 per op)."""
    return (a.double() * b.double() + c.double()).float()


def mm_seqfma(A, B):
    """A[..., M, K] @ B[..., K, N] as ATen's CPU sgemm rounds it for tiny K:
    acc = a0*b0 (rounded); acc = fma(ak, bk, acc) for k = 1..K-1."""
    if _NATIVE["device"] is not None:
        return torch.matmul(A.to(_NATIVE["device"]), B.to(_NATIVE["device"])).cpu()
    K = A.shape[-1]
    acc = A[..., :, 0:1] * B[..., 0:1, :]
    for k in range(1, K):
        acc = _fma(A[..., :, k:k + 1], B[..., k:k + 1, :], acc)
    return acc


def norm3(x):
    """torch.norm(x, dim=-1) for 3-vectors as ATen's CPU kernel rounds it: x0²
    rounded, two FMAs, then a correctly rounded sqrt (taken in fp64 and rounded
    once more – torch.sqrt's vectorised fp32 kernel is NOT correctly rounded,
    torch.norm's is)."""
    if _NATIVE["device"] is not None:
        return torch.norm(x.to(_NATIVE["device"]), dim=-1).cpu()
    acc = x[..., 0] * x[..., 0]
    acc = _fma(x[..., 1], x[..., 1], acc)
    acc = _fma(x[..., 2], x[..., 2], acc)
    return torch.sqrt(acc.double()).float()


def t_vals(S, device="cpu"):
    """linspace(0,1,S) as torch produces it (BaseRender.py:37)."""
    return torch.linspace(0.0, 1.0, steps=S, device=device)


# --------------------------------------------------------------------------
# a7 (mask build) – libs/nerfheads/networks/SparseConvNet.py:135-141
# --------------------------------------------------------------------------


def build_masks3d(levels, threshold=0.1):
    """Occupancy volume on the level-1 grid and the list of occupied voxels.

    levels: list of [1, C, Dk, Hk, Wk].  Returns masks3d [D1,H1,W1] and
    mask_xyz [N,3] float = (x,y,z) level-1 index × 2 (full-res voxel units)."""
    size = levels[0].shape[-3:]
    ups = []
    for feat in levels:
        chan_sum = feat[0].sum(dim=0)
        ups.append(F.interpolate(chan_sum[None, None], size)[0, 0])   # nearest
    masks3d = torch.stack(ups, 0).sum(0)
    idx = torch.stack(torch.where(masks3d > threshold), 0).permute(1, 0)  # (d,h,w)
    mask_xyz = idx.flip(-1).float() * 2.0
    return masks3d, mask_xyz


# --------------------------------------------------------------------------
# a4 – libs/renders/demo_render.py:166-239
# --------------------------------------------------------------------------


def occupied_voxels_world(mask_xyz, voxel_size, bounds, R, Th):
    """demo_render.py:167-168: voxel → SMPL frame → world."""
    pts = mask_xyz * voxel_size + bounds[0, 0]
    return mm_seqfma(pts, R[0].T.contiguous()) + Th[0, 0]


def world_bounds_of(pts_world):
    """demo_render.py:171-175."""
    lo = pts_world.min(0)[0].clone()
    hi = pts_world.max(0)[0].clone()
    lo[2] -= 0.05
    hi[2] += 0.05
    return torch.stack([lo, hi], 0)


def pixel_mask_from_voxels(pts_world, target_pose, target_K, W, Hh=None):
    """demo_render.py:179-200 with W (hard-coded 512 there) as a parameter.
    Returns the float mask [Hh*W] and the ascending flat pixel indices."""
    Hh = W if Hh is None else Hh
    pose = target_pose[0]
    cam = mm_seqfma(pts_world, pose[:, :3].T.contiguous()) + pose[:, 3:].T
    pix = mm_seqfma(cam, target_K[0].T.contiguous())
    xy = pix[:, :2] / pix[:, 2:]
    minx, miny = xy[:, 0].long(), xy[:, 1].long()          # truncation toward zero
    maxx, maxy = minx + 1, miny + 1
    minx, maxx = minx.clamp(0, W - 1), maxx.clamp(0, W - 1)
    miny, maxy = miny.clamp(0, Hh - 1), maxy.clamp(0, Hh - 1)
    idx = torch.cat([miny * W + minx, maxy * W + minx, miny * W + maxx, maxy * W + maxx], 0)
    mask = torch.zeros(Hh * W)
    mask[idx] = 1.0
    pix_idx = torch.where(mask == 1)[0]
    return mask, pix_idx


def rays_bbox(pix_idx, W, target_pose, target_K_inv, can_bounds, neg_ray=False):
    """demo_render.py:200-239: rays through the masked pixels, 6-plane AABB
    test (exactly two hits), near/far.  Returns a dict; `keep` indexes pix_idx."""
    pose = target_pose[0]
    j = torch.div(pix_idx, W, rounding_mode="floor")
    i = pix_idx - j * W
    xy1 = torch.stack([i, j, torch.ones_like(i)], -1).float()
    origin = mm_seqfma(-pose[:, :3].T.contiguous(), pose[:, 3:].contiguous()).view(-1)   # [3]
    cam = mm_seqfma(xy1, target_K_inv[0].T.contiguous())
    world = mm_seqfma(cam - pose[:, 3:].T, pose[:, :3].contiguous())
    d = world - origin[None]
    o = origin[None].expand(d.shape)
    t = ((can_bounds[None] - o[:, None]) / d[:, None]).reshape(-1, 6)
    p = t[..., None] * d[:, None] + o[:, None]                     # [R0,6,3]
    lo, hi = can_bounds[0], can_bounds[1]
    eps = 1e-6
    inside = ((p[..., 0] >= (lo[0] - eps)) & (p[..., 0] <= (hi[0] + eps))
              & (p[..., 1] >= (lo[1] - eps)) & (p[..., 1] <= (hi[1] + eps))
              & (p[..., 2] >= (lo[2] - eps)) & (p[..., 2] <= (hi[2] + eps)))
    at_box = inside.sum(-1) == 2
    hits = p[at_box][inside[at_box]].reshape(-1, 2, 3)
    o_k, d_k = o[at_box], d[at_box]
    nrm = norm3(d_k)
    d0 = norm3(hits[:, 0] - o_k) / nrm
    d1 = norm3(hits[:, 1] - o_k) / nrm
    if neg_ray:
        d1 = -d1
    return {
        "rays_o": o_k.contiguous(), "rays_d": d_k.contiguous(),
        "near": torch.min(d0, d1), "far": torch.max(d0, d1),
        "mask_at_box": at_box, "ray_pix": pix_idx[at_box],
    }


# --------------------------------------------------------------------------
# a3 – BaseRender.py:35-50 / demo_render.py:59-74
# --------------------------------------------------------------------------


def sampling_points(ray_o, ray_d, near, far, S, t_rand=None):
    """ray_o/d [R,3], near/far [R] → pts [R,S,3], z [R,S].  `t_rand` [R,S] is
    the host-drawn jitter of training mode (BaseRender.py:40-47)."""
    t = t_vals(S).to(near)
    z = near[:, None] * (1.0 - t) + far[:, None] * t
    if t_rand is not None:
        mids = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    pts = ray_o[:, None] + ray_d[:, None] * z[..., None]
    return pts, z


# --------------------------------------------------------------------------
# a6 – BaseRender.py:52-73 / demo_render.py:76-94
# --------------------------------------------------------------------------


def pts_to_can_pts(pts, R, Th):
    """(p − Th)·R, pts [P,3]."""
    return mm_seqfma(pts - Th.reshape(1, 3), R[0].contiguous())


def grid_coords_of(pts_smpl, bounds, out_sh_dhw, voxel=0.005):
    """Normalised (x,y,z) grid coordinates in [-1,1] w.r.t. the padded full-res
    shape.  demo_render.py:86-94 (xyz order) and BaseRender.py:62-73 (dhw order,
    swapped back) perform the same per-component arithmetic."""
    xyz = pts_smpl - bounds[0, 0][None]
    xyz = xyz / voxel
    sh_xyz = torch.as_tensor(out_sh_dhw, dtype=torch.float32).flip(-1)
    return xyz / sh_xyz * 2 - 1


# --------------------------------------------------------------------------
# a7 / a8 – demo_render.py:270-283, SparseConvNet.py:111-122
# --------------------------------------------------------------------------


def trilinear(vol, grid):
    """vol [C,D,H,W], grid [P,3] (x,y,z) → [P,C]; align_corners, zeros padding."""
    if _NATIVE["device"] is not None:
        d = _NATIVE["device"]
        out = F.grid_sample(vol[None].to(d), grid[None, None, None].to(d), padding_mode="zeros", align_corners=True).cpu()
    else:
        out = F.grid_sample(vol[None], grid[None, None, None], padding_mode="zeros", align_corners=True)
    return out[0, :, 0, 0].permute(1, 0)


def occupancy_valid(masks3d, grid):
    """valid = where(trilinear(masks3d) > 0), ascending (demo_render.py:274-281)."""
    occ = trilinear(masks3d[None], grid)[:, 0]
    return torch.where(occ > 0)[0]


def gather_levels(levels, grid):
    """[P, 32·L] in level-major channel order (SparseConvNet.py:111-122)."""
    return torch.cat([trilinear(lv[0], grid) for lv in levels], -1)


# --------------------------------------------------------------------------
# a9 – BaseRender.py:283-363 / demo_render.py:506-609
# --------------------------------------------------------------------------


def pack_cameras(src_poses, src_Ks, H, W):
    """src_cameras [1,V,34] = (H, W, K 4×4, E 4×4) – BaseRender.py:233-247."""
    V = src_poses.shape[1]
    Eh = torch.eye(4).repeat(1, V, 1, 1)
    Eh[:, :, :3, :4] = src_poses
    Kh = torch.eye(4).repeat(1, V, 1, 1)
    Kh[:, :, :3, :3] = src_Ks
    cams = torch.ones(1, V, 34)
    cams[:, :, 0] = H
    cams[:, :, 1] = W
    cams[:, :, 2:18] = Kh.reshape(1, V, 16)
    cams[:, :, 18:] = Eh.reshape(1, V, 16)
    return cams


def project(xyz, cams, neg_ray=False):
    """compute_projections (BaseRender.py:301-323): xyz [P,3] → pix [V,P,2],
    in_front [V,P]."""
    Kh = cams[:, 2:18].reshape(-1, 4, 4)
    Eh = cams[:, 18:].reshape(-1, 4, 4)
    KE = Kh.bmm(Eh)
    xyz_h = torch.cat([xyz, torch.ones_like(xyz[:, :1])], -1)
    proj = mm_seqfma(KE, xyz_h.t()[None].expand(KE.shape[0], -1, -1)).permute(0, 2, 1)  # [V,P,4]
    pix = proj[..., :2] / proj[..., 2:3]
    pix = torch.clamp(pix, min=-1e6, max=1e6)
    in_front = proj[..., 2] < 0 if neg_ray else proj[..., 2] > 0
    return pix, in_front


def projector_compute(xyz, src_imgs01, cams, featmaps, neg_ray=False):
    """Projector.compute (demo_render.py:560-609): xyz [P,3]; src_imgs01
    [V,3,H,W] already ×0.5+0.5; featmaps [V,C,h,w].  Returns rgb_feat
    [P,V,3+C] (RGB first) and mask [P,V] float."""
    cams = cams[0]
    h, w = cams[0][:2]
    pix, in_front = project(xyz, cams, neg_ray)
    resize = torch.stack([w - 1.0, h - 1.0])[None, None]
    norm = 2 * pix / resize - 1.0                                       # [V,P,2]
    if _NATIVE["device"] is not None:
        d = _NATIVE["device"]
        rgb = F.grid_sample(src_imgs01.to(d), norm[:, :, None].to(d), align_corners=True)[..., 0].cpu()
        feat = F.grid_sample(featmaps.to(d), norm[:, :, None].to(d), align_corners=True)[..., 0].cpu()
    else:
        rgb = F.grid_sample(src_imgs01, norm[:, :, None], align_corners=True)[..., 0]      # [V,3,P]
        feat = F.grid_sample(featmaps, norm[:, :, None], align_corners=True)[..., 0]       # [V,C,P]
    rgb_feat = torch.cat([rgb, feat], 1).permute(2, 0, 1).contiguous()                # [P,V,3+C]
    inbound = ((pix[..., 0] <= w - 1.0) & (pix[..., 0] >= 0)
               & (pix[..., 1] <= h - 1.0) & (pix[..., 1] >= 0))
    mask = (inbound & in_front).float().permute(1, 0).contiguous()                     # [P,V]
    return rgb_feat, mask


# --------------------------------------------------------------------------
# a10-a14 – libs/nerfheads/trainhead.py
# --------------------------------------------------------------------------


def mean_var(rgb_feat):
    """fused_mean_variance (trainhead.py:20-24): population stats over views,
    unmasked.  rgb_feat [P,V,Cf] → mean, var [P,Cf]."""
    mean = rgb_feat.mean(1, keepdim=True)
    var = ((rgb_feat - mean) ** 2).mean(1)
    return mean[:, 0], var


def _seq(x, w, prefix, acts):
    """nn.Sequential of Linear/activation pairs keyed `prefix.{0,2,4,..}`."""
    for n, act in enumerate(acts):
        x = F.linear(x, w[f"{prefix}.{2 * n}.weight"], w[f"{prefix}.{2 * n}.bias"])
        if act == "elu":
            x = F.elu(x)
        elif act == "relu":
            x = F.relu(x)
    return x


def sigma_feat_of(vol_feat, w):
    """NeRFSigmaHead.out_geometry_fc (trainhead.py:39-41): 128→64 + ELU."""
    return _seq(vol_feat, w, "sigmahead.out_geometry_fc", ["elu"])


def density_mlp(sigma_feat, mean, var, mask, w):
    """NeRFRGBHead.out_geometry_fc on [sigma_feat | mean | var] and the
    no-valid-view fill (trainhead.py:102-110, 133-137; demo_render.py:298-304).
    Returns sigma [P]."""
    x = torch.cat([sigma_feat, mean, var], -1)
    sigma = _seq(x, w, "rgbhead.out_geometry_fc", ["elu", "elu", "elu", "relu"])[:, 0]
    return sigma.masked_fill(mask.sum(1) < 1, 0.0)


def color_mlp(rgb_feat, mean, var, w):
    """Colour trunk of NeRFRGBHead.forward (trainhead.py:128-145).  rgb_feat
    [P,V,Cf] → rgb [P,3]."""
    V = rgb_feat.shape[1]
    glob = torch.cat([mean, var], -1)[:, None].expand(-1, V, -1)
    x = torch.cat([glob, rgb_feat], -1)
    x = _seq(x, w, "rgbhead.base_fc", ["elu", "elu"])
    x = x + _seq(x * 1.0 / V, w, "rgbhead.vis_fc", ["elu", "elu"])
    return _seq(x.flatten(1, 2), w, "rgbhead.rgb_fc", ["elu", "elu", None]).sigmoid()


# --------------------------------------------------------------------------
# a13, a15, a16
# --------------------------------------------------------------------------


def alpha_valid(sigma):
    """demo_render.py:312-317."""
    alpha = 1.0 - torch.exp(-sigma)
    return alpha, torch.where(alpha > 1e-14)[0]


def composite_progressive(R, S, valid, valid1, alpha, rgb):
    """demo_render.py:335-347: scatter to dense [R,S], exclusive cumprod of
    (1−α+1e-10), weights, Σ w·rgb."""
    hold_rgb = torch.zeros(R * S, 3)
    hold_alpha = torch.zeros(R * S)
    hold_rgb[valid[valid1]] = rgb
    hold_alpha[valid] = alpha
    hold_rgb = hold_rgb.view(R, S, 3)
    hold_alpha = hold_alpha.view(R, S)
    T = torch.cumprod(1.0 - hold_alpha + 1e-10, -1)[..., :-1]
    T = torch.cat([torch.ones_like(T[..., :1]), T], -1)
    weights = hold_alpha * T
    return (weights[..., None] * hold_rgb).sum(1), weights


def raw2outputs(raw, z_vals, neg=False):
    """Renderer.raw2outputs (BaseRender.py:75-107) minus the unused mask."""
    rgb, sigma = raw[:, :, :3], raw[:, :, 3]
    if neg:
        rgb, sigma = torch.flip(rgb, [1]), torch.flip(sigma, [1])
    alpha = 1.0 - torch.exp(-sigma)
    T = torch.cumprod(1.0 - alpha + 1e-10, -1)[:, :-1]
    T = torch.cat([torch.ones_like(T[:, :1]), T], -1)
    weights = alpha * T
    rgb_map = (weights[..., None] * rgb).sum(1)
    depth = (weights * z_vals).sum(-1)
    acc = weights.sum(-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    return rgb_map, disp, acc, weights, depth


# --------------------------------------------------------------------------
# whole-path drivers
# --------------------------------------------------------------------------


def _scene_common(scene):
    H, W = scene["H"], scene["W"]
    cams = pack_cameras(scene["src_poses"], scene["src_Ks"], H, W)
    imgs01 = scene["src_imgs"][0] * 0.5 + 0.5                 # BaseRender.py:231
    out_sh = [int(v) for v in scene["out_sh"][0]]
    return cams, imgs01, out_sh


@torch.no_grad()
def render_progressive(scene, w, S=64, neg_ray=False, chunk=None, keep=False):
    """Device-agnostic restatement of demo_render.Renderer.render_rays
    (demo_render.py:96-365) downstream of the upstream producers, i.e. with
    `levels`/`featmaps` supplied.  `chunk` bounds peak memory of the head stage
    (points per pass, ascending order; results are order-identical).
    Returns rgb_map [R,3], pred_img [H,W,3], mask_at_box [H*W] and – if `keep`
    – every intermediate the parity tests compare."""
    cams, imgs01, out_sh = _scene_common(scene)
    W_img, H_img = scene["W"], scene["H"]
    R_, Th, bounds = scene["R"], scene["Th"], scene["bounds"]
    levels, featmaps = scene["levels"], scene["featmaps"]

    masks3d, mask_xyz = build_masks3d(levels)
    vox_world = occupied_voxels_world(mask_xyz, torch.tensor(VOXEL3), bounds, R_, Th)
    can_bounds = world_bounds_of(vox_world)
    pix_mask, pix_idx = pixel_mask_from_voxels(vox_world, scene["target_pose"], scene["target_K"], W_img, H_img)
    rb = rays_bbox(pix_idx, W_img, scene["target_pose"], scene["target_K_inv"], can_bounds, neg_ray)
    n_rays = rb["near"].shape[0]
    pts, z = sampling_points(rb["rays_o"], rb["rays_d"], rb["near"], rb["far"], S)
    pts = pts.reshape(-1, 3)
    grid = grid_coords_of(pts_to_can_pts(pts, R_, Th), bounds, out_sh)
    valid = occupancy_valid(masks3d, grid)

    P1 = valid.shape[0]
    chunk = P1 if not chunk else chunk
    sig_l, rgbf_l, mean_l, var_l, mask_l, sfeat_l, vfeat_l = [], [], [], [], [], [], []
    for s0 in range(0, max(P1, 1), max(chunk, 1)):
        sel = valid[s0:s0 + chunk]
        rgb_feat, mask = projector_compute(pts[sel], imgs01, cams, featmaps, neg_ray)
        vol_feat = gather_levels(levels, grid[sel])
        sfeat = sigma_feat_of(vol_feat, w)
        mean, var = mean_var(rgb_feat)
        sig_l.append(density_mlp(sfeat, mean, var, mask, w))
        rgbf_l.append(rgb_feat); mean_l.append(mean); var_l.append(var); mask_l.append(mask)
        if keep:
            sfeat_l.append(sfeat); vfeat_l.append(vol_feat)
    sigma = torch.cat(sig_l) if sig_l else torch.zeros(0)
    alpha, valid1 = alpha_valid(sigma)
    rgb_feat = torch.cat(rgbf_l); mean = torch.cat(mean_l); var = torch.cat(var_l)
    rgb_l = []
    P2 = valid1.shape[0]
    for s0 in range(0, max(P2, 1), max(chunk, 1)):
        sel = valid1[s0:s0 + chunk]
        rgb_l.append(color_mlp(rgb_feat[sel], mean[sel], var[sel], w))
    rgb = torch.cat(rgb_l) if rgb_l else torch.zeros(0, 3)
    rgb_map, weights = composite_progressive(n_rays, S, valid, valid1, alpha, rgb)

    pix_mask_final = torch.zeros(H_img * W_img, dtype=torch.bool)
    pix_mask_final[rb["ray_pix"]] = True                     # demo_render.py:348-351
    pred_img = torch.zeros(H_img * W_img, 3, dtype=torch.float64)
    pred_img[rb["ray_pix"]] = rgb_map.double()
    out = {"rgb_map": rgb_map, "pred_img": pred_img.view(H_img, W_img, 3),
           "mask_at_box": pix_mask_final, "n_rays": n_rays, "P": n_rays * S, "P1": P1, "P2": P2}
    if keep:
        out.update({
            "masks3d": masks3d, "mask_xyz": mask_xyz, "can_bounds": can_bounds,
            "pix_mask": pix_mask, "pix_idx": pix_idx, "ray_pix": rb["ray_pix"],
            "box_hit": rb["mask_at_box"], "rays_o": rb["rays_o"], "rays_d": rb["rays_d"],
            "near": rb["near"], "far": rb["far"], "z_vals": z, "valid": valid, "valid1": valid1,
            "sigma": sigma, "alpha": alpha, "rgb": rgb, "weights": weights,
            "rgb_feat": rgb_feat, "mean": mean, "var": var, "mask": torch.cat(mask_l),
            "sigma_feat": torch.cat(sfeat_l), "vol_feat": torch.cat(vfeat_l),
        })
    return out


@torch.no_grad()
def render_dense(scene, w, S=64, neg_ray=False, t_rand=None, chunk=2000, rays=None, keep=False):
    """BaseRender.Renderer.render_rays/batchify_rays (BaseRender.py:110-184)
    with the volume levels supplied: every sample point goes through both
    heads, no compaction.  `rays` = (o,d,near,far) or taken from the scene."""
    cams, imgs01, out_sh = _scene_common(scene)
    R_, Th, bounds = scene["R"], scene["Th"], scene["bounds"]
    levels, featmaps = scene["levels"], scene["featmaps"]
    if rays is None:
        rays = (scene["ray_o"][0], scene["ray_d"][0], scene["near"][0], scene["far"][0])
    o, d, near, far = rays
    V = cams.shape[1]
    outs = {k: [] for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "alpha", "z_vals", "rgb_in_map")}
    raws = []
    for r0 in range(0, o.shape[0], chunk):
        sl = slice(r0, r0 + chunk)
        tr = None if t_rand is None else t_rand[sl]
        pts, z = sampling_points(o[sl], d[sl], near[sl], far[sl], S, tr)
        n = pts.shape[0]
        pts = pts.reshape(-1, 3)
        grid = grid_coords_of(pts_to_can_pts(pts, R_, Th), bounds, out_sh)
        rgb_feat, mask = projector_compute(pts, imgs01, cams, featmaps, neg_ray)
        sfeat = sigma_feat_of(gather_levels(levels, grid), w)
        mean, var = mean_var(rgb_feat)
        sigma = density_mlp(sfeat, mean, var, mask, w)
        rgb = color_mlp(rgb_feat, mean, var, w)
        raw = torch.cat([rgb, sigma[:, None]], -1).view(n, S, 4)
        rgb_map, disp, acc, weights, depth = raw2outputs(raw, z, neg_ray)
        rgb_in = rgb_feat[..., :3].reshape(n, S, V, 3)
        rgb_in_map = (weights[..., None, None] * rgb_in).sum(1)
        for k, v in (("rgb_map", rgb_map), ("disp_map", disp[:, None]), ("acc_map", acc[:, None]),
                     ("depth_map", depth[:, None]), ("alpha", weights), ("z_vals", z),
                     ("rgb_in_map", rgb_in_map.reshape(n, -1))):
            outs[k].append(v)
        if keep:
            raws.append(raw)
    ret = {k: torch.cat(v, 0) for k, v in outs.items()}
    if keep:
        ret["raw"] = torch.cat(raws, 0)
    return ret


VOXEL3 = [0.005, 0.005, 0.005]


def psnr(a, b):
    """libs/evaluators/if_nerf.py:29-32 (10·log10(1/mse))."""
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10.0 * math.log10(1.0 / max(mse, 1e-20))


def psnr_masked(a, b, mask_at_box):
    """PSNR the way the reference's evaluator forms it for a rendered frame
    (libs/evaluators/if_nerf.py:49-57: `rgb_pred = output['pred_img'][mask_at_box]`,
    then mse over those pixels only): the ≈87 % background pixels, which are
    exactly 0 in every render, do not dilute the error."""
    m = torch.as_tensor(mask_at_box).reshape(-1).bool()
    a2, b2 = a.reshape(-1, 3)[m].double(), b.reshape(-1, 3)[m].double()
    mse = float(((a2 - b2) ** 2).mean()) if a2.numel() else 0.0
    return 10.0 * math.log10(1.0 / max(mse, 1e-20))


# --------------------------------------------------------------------------
# Row f3: the dataset path's rays – numpy, as the CPU loader computes them
# --------------------------------------------------------------------------
def dataset_rays(H, W, K, R, T, bounds):
    """libs/datasets/data_utils.py:47-63 (get_rays) followed by the test-split
    branch of sample_ray (:331-337) and get_near_far (:96-130), dtype promotions
    included: the rays are formed in fp64 and cast to fp32, the box test then runs
    in fp64 on those fp32 rays (bounds ± 0.01, |d| < 1e-5 → 1e-5, eps 1e-6, exactly
    two hits, both depths signed by the FIRST hit's side).
    Returns (ray_o [R,3] f32, ray_d [R,3] f32, near [R] f32, far [R] f32,
    mask_at_box [H*W] bool) – numpy arrays."""
    import numpy as np
    K, R, T = np.asarray(K), np.asarray(R), np.asarray(T)
    R_inv = np.linalg.inv(R)
    T = -R_inv @ T.reshape(3)
    rays_o = T.ravel()
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    xy1 = np.stack([i, j, np.ones_like(i)], axis=2)
    pixel_camera = np.dot(xy1, np.linalg.inv(K).T)
    pixel_world = (pixel_camera @ R_inv.T) + T[np.newaxis, ...]
    rays_d = pixel_world - rays_o[None, None]
    ray_o = np.broadcast_to(rays_o, rays_d.shape).reshape(-1, 3).astype(np.float32)
    ray_d = rays_d.reshape(-1, 3).astype(np.float32)
    # get_near_far
    b = np.asarray(bounds) + np.array([-0.01, 0.01])[:, None]
    nominator = b[None] - ray_o[:, None]
    ray_d = ray_d.copy()
    ray_d[np.abs(ray_d) < 1e-5] = 1e-5
    d_intersect = (nominator / ray_d[:, None]).reshape(-1, 6)
    p_intersect = d_intersect[..., None] * ray_d[:, None] + ray_o[:, None]
    min_x, min_y, min_z, max_x, max_y, max_z = b.ravel()
    eps = 1e-6
    inside = ((p_intersect[..., 0] >= (min_x - eps)) * (p_intersect[..., 0] <= (max_x + eps)) *
              (p_intersect[..., 1] >= (min_y - eps)) * (p_intersect[..., 1] <= (max_y + eps)) *
              (p_intersect[..., 2] >= (min_z - eps)) * (p_intersect[..., 2] <= (max_z + eps)))
    mask_at_box = inside.sum(-1) == 2
    p_int = p_intersect[mask_at_box][inside[mask_at_box]].reshape(-1, 2, 3)
    ro, rd = ray_o[mask_at_box], ray_d[mask_at_box]
    norm_ray = np.linalg.norm(rd, axis=1)
    sign = np.array(((p_int[:, 0] - ro) * rd).sum(axis=1) < 0.0, dtype=np.int64) * -2 + 1
    d0 = np.linalg.norm(p_int[:, 0] - ro, axis=1) / norm_ray * sign
    d1 = np.linalg.norm(p_int[:, 1] - ro, axis=1) / norm_ray * sign
    near = np.minimum(d0, d1).astype(np.float32)
    far = np.maximum(d0, d1).astype(np.float32)
    return ro, rd, near, far, mask_at_box


@torch.no_grad()
def mesh_cube(scene, w, pts_grid, inside, neg_ray=False, chunk=65536):
    """Row f4: the mesh branch's occupancy cube, BaseRender.py:255-270
    (demo_render.py:249-268): σ of the density head at the grid points selected by
    `inside`, α = 1 - exp(-σ) scattered into the grid and zero-padded by 10 voxels
    (marching cubes over it is mcubes' job in the reference).  pts_grid [X,Y,Z,3]
    world points, inside [X,Y,Z] bool.  Returns the float64 numpy cube."""
    import numpy as np
    cams, imgs01, out_sh = _scene_common(scene)
    R_, Th, bounds = scene["R"], scene["Th"], scene["bounds"]
    pts = pts_grid[inside].reshape(-1, 3).float()
    sig = []
    for p0 in range(0, pts.shape[0], chunk):
        p = pts[p0:p0 + chunk]
        grid = grid_coords_of(pts_to_can_pts(p, R_, Th), bounds, out_sh)
        rgb_feat, mask = projector_compute(p, imgs01, cams, scene["featmaps"], neg_ray)
        sfeat = sigma_feat_of(gather_levels(scene["levels"], grid), w)
        mean, var = mean_var(rgb_feat)
        sig.append(density_mlp(sfeat, mean, var, mask, w))
    sigma = torch.cat(sig) if sig else torch.zeros(0)
    alpha = (1.0 - torch.exp(-sigma)).numpy()
    cube = np.zeros(tuple(inside.shape))
    cube[inside.numpy().astype(bool)] = alpha
    return np.pad(cube, 10, mode="constant")


# --------------------------------------------------------------------------
# Row f1: the sparse-conv pyramid as a dense emulation
# --------------------------------------------------------------------------
@torch.no_grad()
def sparse_conv_net(state, features, coords, spatial_shape, n_layers=4, eps=1e-3):
    """libs/nerfheads/networks/SparseConvNet.py:21-124 with every spconv layer
    replaced by its dense equivalent (spconv 1.2.1's documented semantics, the
    way its own unit tests check it against torch.nn.Conv3d):
      SubMConv3d(3)       = conv3d(padding=1) evaluated on the input's active sites only
      SparseConv3d(3,2,1) = conv3d(stride=2, padding=1), active where any input site is in reach
    with weight[kd,kh,kw,in,out] ↔ conv3d weight[out,in,kd,kh,kw], followed by
    BatchNorm1d (running statistics) and ReLU on the active sites.  PARITY UNPINNED
    against spconv itself (absent from the reference tree and from this image).
    `state`: the module's state_dict.  Duplicate voxels: the smallest row owns the site.
    Returns the 4 dense levels [1, C, D_k, H_k, W_k] (what x.dense() gives)."""
    D, H, W = [int(v) for v in spatial_shape]
    C0 = features.shape[1]
    dense = torch.zeros(1, C0, D, H, W)
    active = torch.zeros(1, 1, D, H, W)
    crd = coords[:, -3:].long()
    for i in range(crd.shape[0] - 1, -1, -1):          # reverse order: the smallest row index wins
        d, h, w = [int(v) for v in crd[i]]
        if 0 <= d < D and 0 <= h < H and 0 <= w < W:
            dense[0, :, d, h, w] = features[i]
            active[0, 0, d, h, w] = 1.0

    def layer(x, act, prefix, stride):
        wt = state[prefix + ".weight"].permute(4, 3, 0, 1, 2).contiguous()
        bn = prefix.rsplit(".", 1)[0] + "." + str(int(prefix.rsplit(".", 1)[1]) + 1)
        y = F.conv3d(x, wt, stride=stride, padding=1)
        if stride == 2:
            act = (F.max_pool3d(act, 3, 2, 1) > 0).float()
        inv = torch.rsqrt(state[bn + ".running_var"] + eps)
        scale = state[bn + ".weight"] * inv
        shift = state[bn + ".bias"] - state[bn + ".running_mean"] * scale
        y = torch.relu(y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) * act
        return y, act
    x, act = dense, active
    x, act = layer(x, act, "net.0.0", 1)
    x, act = layer(x, act, "net.0.3", 1)
    levels = []
    for i in range(n_layers):
        x, act = layer(x, act, f"net.{2 * i + 1}.0", 2)
        x, act = layer(x, act, f"net.{2 * i + 2}.0", 1)
        x, act = layer(x, act, f"net.{2 * i + 2}.3", 1)
        levels.append(x)
    return levels


def smpl_code_attention(state, code, feats, n_head=4, prefix=""):
    """libs/nerfheads/networks/MultiHeadAttention.py:62-98 as trainhead.py:50 calls it (sum=False, mask=None):
    code [n, d_model] (one query per SMPL vertex), feats [n, V, kv_dim] (keys = values) → [n, d_model].
    Pinned against the reference module itself: tests/golden/attention.npz (oracle/gen_golden_attn.py)."""
    wq, wk, wv, wfc = (state[prefix + k].double() for k in ("w_qs.weight", "w_ks.weight", "w_vs.weight", "fc.weight"))
    n, V, _ = feats.shape
    d_k = wq.shape[0] // n_head
    q = (code.double() @ wq.t()).view(n, 1, n_head, d_k).transpose(1, 2)            # [n, h, 1, dk]
    k = (feats.double() @ wk.t()).view(n, V, n_head, d_k).transpose(1, 2)           # [n, h, V, dk]
    v = (feats.double() @ wv.t()).view(n, V, n_head, d_k).transpose(1, 2)
    attn = torch.softmax((q / d_k ** 0.5) @ k.transpose(2, 3), dim=-1)              # [n, h, 1, V]
    o = (attn @ v).transpose(1, 2).reshape(n, n_head * d_k)
    return (o @ wfc.t()).float()


def smpl_features(smpl_xyz, cams, featmaps, neg_ray=False):
    """Projector.compute_smpl (demo_render.py:612-632): smpl_xyz [n,3], cams [1,V,34], featmaps [V,C,h,w] →
    [n, V, C]: the pixel-aligned features of the SMPL vertices (bilinear, zeros outside, no mask)."""
    dummy = torch.zeros(featmaps.shape[0], 3, int(cams[0, 0, 0]), int(cams[0, 0, 1]))
    rgb_feat, _mask = projector_compute(smpl_xyz, dummy, cams, featmaps, neg_ray)
    return rgb_feat[..., 3:].contiguous()
