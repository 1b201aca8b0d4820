"""Mirror of the reference's head modules (libs/nerfheads/trainhead.py) with
the hot-path arithmetic executed by libgpnerf_b200.so.

Same class names, constructor arguments, method signatures and parameter names
as the reference, so `state_dict` keys line up:
``sigmahead.c.weight``, ``sigmahead.out_geometry_fc.0.{weight,bias}``,
``rgbhead.{base_fc.{0,2},vis_fc.{0,2},rgb_fc.{0,2,4},out_geometry_fc.{0,2,4,6}}``.
The two upstream producers the sigma head owns in the reference –
``xyzc_attn`` (MultiHeadAttention) and ``xyzc_net`` (spconv SparseConvNet) – are
SURVEY.md §8f row 1 ("next"): mirrors on K8/K7 (attention.py, sparseconv.py)
with the reference's parameter names, so the whole ``sigmahead.*`` state_dict
loads; they run in inference form and hand the renderer sparse level rows.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import PREC_FP32
from .attention import MultiHeadAttention
from .sparseconv import SparseConvNet


def weights_init(m):
    """trainhead.py:13-17"""
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class NeRFSigmaHead(nn.Module):
    """trainhead.py:27-76"""

    def __init__(self, in_feat_ch=32, n_smpl=6890, code_dim=16, attn_n_heads=4, spconv_n_layers=4,
                 spconv_out_dim=(32, 32, 32, 32), precision=PREC_FP32):
        super().__init__()
        self.c = nn.Embedding(n_smpl, code_dim)
        self.xyzc_attn = MultiHeadAttention(attn_n_heads, code_dim, code_dim // attn_n_heads, code_dim // attn_n_heads,
                                            kv_dim=in_feat_ch, sum=False)
        self.xyzc_net = SparseConvNet(n_layers=spconv_n_layers, in_dim=code_dim, out_dim=list(spconv_out_dim))
        self.out_geometry_fc = nn.Sequential(nn.Linear(sum(spconv_out_dim), 64), nn.ELU(inplace=True))
        self.out_geometry_fc.apply(weights_init)
        self.precision = precision

    @torch.no_grad()
    def encode_geometry(self, smpl_feat_sampled, coord, out_sh):
        """trainhead.py:44-54 up to the SparseConvTensor, then SparseConvNet.py:104-110 without ``.dense()``:
        smpl_feat_sampled [1|B, n_smpl, V, C] (Projector.compute_smpl), coord [n_smpl, 3|4] voxel indices,
        out_sh (D, H, W) → (levels_sparse, level_dims, n_rows_dev) for Engine.upload_products_sparse."""
        code = self.c.weight.detach()                         # = self.c(arange(n_smpl)), trainhead.py:48
        feats = smpl_feat_sampled.flatten(0, 1)
        fused = self.xyzc_attn(code.unsqueeze(1), feats, feats)[0].squeeze(1)
        return self.xyzc_net(fused, coord, out_sh)

    def volume_features(self, sp_input, grid_coords):
        """The gather half of SparseConvNet.forward (SparseConvNet.py:111-122)
        on the dense levels in sp_input['levels_cl'] → [n,128]."""
        return ops.gather_volume(sp_input["levels_cl"], sp_input["frame"], grid_coords.reshape(-1, 3),
                                 normalised=True)

    def test_forward(self, sp_input, grid_coords, rgb_feat, mask):
        """trainhead.py:61-76 → (sigma_feat [R,S,64], globalfeat [R,S,1,134])"""
        n_rays, n_samples, n_views = rgb_feat.shape[:3]
        vol = self.volume_features(sp_input, grid_coords)
        meanvar = ops.mean_variance(rgb_feat)
        hw, _keep = ops.pack_head_weights(sp_input["head_state"], vol.device, n_views)
        _sigma, sfeat = ops.density_mlp(vol, meanvar, mask.reshape(-1, n_views), hw, self.precision,
                                        want_sigma_feat=True)
        sigma_feat = sfeat.view(n_rays, n_samples, 64)
        globalfeat = torch.cat([sigma_feat.unsqueeze(-2), meanvar.view(n_rays, n_samples, 1, 70)], dim=-1)
        return sigma_feat, globalfeat


class NeRFRGBHead(nn.Module):
    """trainhead.py:79-145"""

    def __init__(self, in_feat_ch=32, n_views=3, precision=PREC_FP32):
        super().__init__()
        cf = in_feat_ch + 3
        self.base_fc = nn.Sequential(nn.Linear(cf * 3, 64), nn.ELU(inplace=True), nn.Linear(64, 32),
                                     nn.ELU(inplace=True))
        self.vis_fc = nn.Sequential(nn.Linear(32, 32), nn.ELU(inplace=True), nn.Linear(32, 32),
                                    nn.ELU(inplace=True))
        self.rgb_fc = nn.Sequential(nn.Linear(32 * n_views, 32), nn.ELU(inplace=True), nn.Linear(32, 16),
                                    nn.ELU(inplace=True), nn.Linear(16, 3))
        self.out_geometry_fc = nn.Sequential(nn.Linear(64 + cf * 2, 64), nn.ELU(inplace=True), nn.Linear(64, 32),
                                             nn.ELU(inplace=True), nn.Linear(32, 16), nn.ELU(inplace=True),
                                             nn.Linear(16, 1), nn.ReLU())
        for seq in (self.out_geometry_fc, self.base_fc, self.vis_fc, self.rgb_fc):
            seq.apply(weights_init)
        self.precision = precision

    def _weights(self, device, n_views):
        sd = {"rgbhead." + k: v for k, v in self.state_dict().items()}
        dummy = torch.zeros(1, device=device)
        if self.precision != PREC_FP32:
            dummy = torch.zeros(64 * 128, device=device)      # the bf16 image packs it; never used here
        sd["sigmahead.out_geometry_fc.0.weight"] = dummy.view(64, -1) if dummy.numel() > 1 else dummy
        sd["sigmahead.out_geometry_fc.0.bias"] = dummy[:64] if dummy.numel() > 1 else dummy
        return ops.pack_head_weights(sd, device, n_views)

    def forward(self, rgb_feat, sigma_feat, mask):
        """(rgb_in, rgb_out, sigma_out) – trainhead.py:118-145"""
        n_rays, n_samples, n_views = rgb_feat.shape[:3]
        hw, _keep = self._weights(rgb_feat.device, n_views)
        meanvar = ops.mean_variance(rgb_feat)
        sigma = ops.density_mlp(sigma_feat.reshape(-1, 64), meanvar, mask.reshape(-1, n_views), hw, self.precision)
        rgb = ops.color_mlp(rgb_feat.reshape(-1, n_views, rgb_feat.shape[-1]), meanvar, hw, self.precision)
        return rgb_feat[..., :3], rgb.view(n_rays, n_samples, 3), sigma.view(n_rays, n_samples, 1)


class NeRFHead(nn.Module):
    """trainhead.py:148-163"""

    def __init__(self, in_feat_ch=32, n_smpl=6890, code_dim=16, attn_n_heads=4, spconv_n_layers=4,
                 spconv_out_dim=(32, 32, 32, 32), use_rgbhead=True, n_views=3, precision=PREC_FP32):
        super().__init__()
        self.sigmahead = NeRFSigmaHead(in_feat_ch, n_smpl, code_dim, attn_n_heads, spconv_n_layers,
                                       spconv_out_dim, precision)
        self.use_rgbhead = use_rgbhead
        self.rgbhead = NeRFRGBHead(in_feat_ch, n_views, precision)

    def hot_path_state(self):
        """The parameters the kernels read, keyed as in the reference state_dict."""
        sd = self.state_dict()
        return {k: v for k, v in sd.items()
                if k.startswith("rgbhead.") or k.startswith("sigmahead.out_geometry_fc")}

    def forward(self, sp_input, grid_coords, smpl_feat_sampled, rgb_feat, mask):
        """(raw [R,S,4], rgb_in [R,S,V,3]) – trainhead.py:159-163 with the dense
        levels supplied in sp_input (the spconv pyramid is upstream)."""
        sp_input = dict(sp_input, head_state=self.hot_path_state())
        sigma_feat, _ = self.sigmahead.test_forward(sp_input, grid_coords, rgb_feat, mask)
        rgb_in, rgb_out, sigma_out = self.rgbhead(rgb_feat, sigma_feat, mask)
        return torch.cat([rgb_out, sigma_out], dim=-1), rgb_in


def _precision_of(cfg):
    """`cfg.head.precision`: "fp32" | "bf16" or 0 | 1 (absent in the reference's config tree: fp32 for the
    module-level calls, which mirror the reference's fp32 arithmetic)."""
    p = getattr(cfg.head, "precision", PREC_FP32)
    if isinstance(p, str):
        return {"fp32": 0, "bf16": 1}[p.lower()]
    return int(p)


def build_head(cfg):
    """trainhead.py:166-177"""
    return NeRFHead(in_feat_ch=cfg.encoder.out_ch, use_rgbhead=cfg.head.rgb.use_rgbhead,
                    n_smpl=cfg.head.sigma.n_smpl, code_dim=cfg.head.sigma.code_dim,
                    attn_n_heads=cfg.head.sigma.n_heads, spconv_n_layers=cfg.head.sigma.n_layers,
                    spconv_out_dim=cfg.head.sigma.outdims, n_views=getattr(cfg, "src_view_num", 3),
                    precision=_precision_of(cfg))
