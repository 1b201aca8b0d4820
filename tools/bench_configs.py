"""Device timings of the other BASELINE.json configurations (not bench lines:
bench.py measures configs[1]).  Prints one JSON line per configuration.

  python tools/bench_configs.py [dense512] [zju1024] [zju512_fp32] [thu512]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200._lib import PREC_BF16, PREC_FP32  # noqa: E402
from gpnerf_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"


def timed(fn, steps=10, warmup=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / steps


def progressive(tag, H, S, V, precision, seed=42, graph=True):
    scene = synth.make_scene("zju", H=H, W=H, V=V, seed=seed)
    w = synth.make_head_weights(V=V, seed=seed)
    eng = Engine(H, H, S, V, device=DEV, precision=precision)
    eng.set_weights(w)
    lv = [t.to(DEV) for t in scene["levels"]]
    fm, im = scene["featmaps"].to(DEV), scene["src_imgs"].to(DEV)
    eng.set_static_inputs(lv, fm, im)
    eng.upload_products(lv, fm, im)
    frame = eng.make_frame(scene)

    def step():
        if graph:
            eng.run_progressive_graphed(frame)
        else:
            eng.upload_products(lv, fm, im)
            eng.render_progressive(frame)
    ms = timed(step)
    c = eng.read_counters()
    eng.timing = True
    eng.stage_events = {}
    for _ in range(5):
        eng.upload_products(lv, fm, im)
        eng.render_progressive(frame)
    torch.cuda.synchronize()
    st = {k: round(v, 4) for k, v in sorted(eng.stage_times_ms().items(), key=lambda kv: -kv[1])}
    print(json.dumps({"config": tag, "H": H, "S": S, "V": V, "precision": "bf16" if precision else "fp32",
                      "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "rays": c["n_rays"],
                      "rays_per_s": c["n_rays"] * 1e3 / ms, "points": c["n_rays"] * S, "P1": c["P1"], "P2": c["P2"],
                      "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30, "stages_ms": st}), flush=True)


def dense(tag, H, S, V, precision, seed=42):
    """BaseRender semantics on a full frame: every pixel's ray, every sample
    through both heads (BASELINE configs[0]'s dense worst case)."""
    scene = synth.make_scene("dense", H=H, W=H, V=V, seed=seed)
    w = synth.make_head_weights(V=V, seed=seed)
    R = scene["ray_o"].shape[1]
    eng = Engine(H, H, S, V, device=DEV, precision=precision, max_rays=R)
    eng.set_weights(w)
    lv = [t.to(DEV) for t in scene["levels"]]
    fm, im = scene["featmaps"].to(DEV), scene["src_imgs"].to(DEV)
    rays = tuple(scene[k][0].to(DEV) for k in ("ray_o", "ray_d", "near", "far"))
    eng.upload_products(lv, fm, im)
    frame = eng.make_frame(scene)

    def step():
        eng.upload_products(lv, fm, im)
        return eng.render_dense(frame, *rays)
    ms = timed(step, steps=5, warmup=2)
    eng.timing = True
    eng.stage_events = {}
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    st = {k: round(v, 4) for k, v in sorted(eng.stage_times_ms().items(), key=lambda kv: -kv[1])}
    print(json.dumps({"config": tag, "H": H, "S": S, "V": V, "precision": "bf16" if precision else "fp32",
                      "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "rays": R, "rays_per_s": R * 1e3 / ms,
                      "points": R * S, "points_per_s": R * S * 1e3 / ms,
                      "rgb_mean": float(out["rgb_map"].mean()), "acc_mean": float(out["acc_map"].mean()),
                      "finite": bool(torch.isfinite(out["rgb_map"]).all()),
                      "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30, "stages_ms": st}), flush=True)


def train(tag, H, S, V, R, seed=42, steps=5, precision=0):
    """BASELINE configs[3]: one training step = dense render of R rays (every
    sample through both heads, jitter on) + MSE on rgb_map + backward to the
    head parameters, the encoder feature maps and the 4 volume levels –
    forward and backward both by the library's kernels (gpnerf_b200.train)."""
    from gpnerf_b200.train import render_dense_autograd
    scene = synth.make_scene("zju", H=H, W=H, V=V, seed=seed, with_rays=True)
    w0 = synth.make_head_weights(V=V, seed=seed, random_bias=True)
    n_all = scene["ray_o"].shape[1]
    sel = (torch.arange(R) * max(1, n_all // R)) % n_all
    rays = tuple(scene[k][0][sel].to(DEV) for k in ("ray_o", "ray_d", "near", "far"))
    eng = Engine(H, H, S, V, device=DEV, max_rays=R)
    w_g = {k: v.clone().to(DEV).requires_grad_(True) for k, v in w0.items()}
    lv_g = [t.clone().to(DEV).requires_grad_(True) for t in scene["levels"]]
    fm_g = scene["featmaps"].clone().to(DEV).requires_grad_(True)
    im = scene["src_imgs"].to(DEV)
    eng.set_weights(w0)
    eng.upload_products([t.detach() for t in lv_g], fm_g.detach(), im)
    frame = eng.make_frame(scene)
    target = torch.rand(R, 3, device=DEV)
    gen = torch.Generator().manual_seed(1)

    def step():
        for t in list(w_g.values()) + lv_g + [fm_g]:
            t.grad = None
        t_rand = torch.rand(R, S, generator=gen)          # BaseRender.py:46: drawn on the CPU generator
        out = render_dense_autograd(eng, frame, rays, lv_g, fm_g, im, w_g, t_rand=t_rand, precision=precision)
        loss = ((out["rgb_map"] - target) ** 2).mean()
        loss.backward()
        return loss
    ms = timed(step, steps=steps, warmup=2)
    loss = float(step())
    gn = float(sum(float(v.grad.pow(2).sum()) for v in w_g.values()) ** 0.5)
    print(json.dumps({"config": tag, "H": H, "S": S, "V": V, "rays": R, "precision": "tf32 tcgen05 heads" if precision else "fp32",
                      "ms_per_step": ms, "rays_per_s": R * 1e3 / ms, "points_per_s": R * S * 1e3 / ms,
                      "loss": loss, "head_grad_norm": gn, "finite": bool(torch.isfinite(torch.tensor(gn))),
                      "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


def upstream(tag, H, S, V, seed=42):
    """Row f1: a frame from the encoder's feature maps – project the SMPL vertices, K8 attention, K7 sparse-conv
    pyramid, sparse rows into the fp16 volumes, then K1…K5 – all on the device (no dense fp32 volume)."""
    from gpnerf_b200.nerfhead import NeRFHead
    from gpnerf_b200.render import Projector, Renderer
    scene = synth.make_scene("zju", H=H, W=H, V=V, seed=seed)
    head = NeRFHead(n_views=V, precision=PREC_BF16).eval()
    sd = head.state_dict()
    for k, v in synth.make_head_weights(V=V, seed=seed).items():
        sd[k].copy_(v)
    for k, v in sd.items():
        if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
            v.fill_(3.0)
    head.load_state_dict(sd)
    head = head.to(DEV)
    r = Renderer(None, head, is_train=False, n_samples=S, progressive=True, precision=PREC_BF16)
    batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in scene.items() if k != "levels"}
    out = r.render(dict(batch))
    sh = head.sigmahead
    xyz = batch["feature"][..., :3].float()
    smpl_xyz = torch.bmm(xyz, batch["Rh"].float().transpose(1, 2)) + batch["Th"].float()
    cams = r._pack_cameras(batch, batch["src_imgs"].shape[-2:], DEV)
    out_sh = [int(v) for v in batch["out_sh"][0]]
    coord = batch["coord"][0]
    proj = Projector(DEV)
    st = {}
    feats = proj.compute_smpl(smpl_xyz, cams, batch["featmaps"])
    st["project+gather SMPL (K2)"] = timed(lambda: proj.compute_smpl(smpl_xyz, cams, batch["featmaps"]))
    code = sh.c.weight.detach().unsqueeze(1)
    f2 = feats.flatten(0, 1)
    st["attention (K8)"] = timed(lambda: sh.xyzc_attn(code, f2, f2))
    fused = sh.xyzc_attn(code, f2, f2)[0].squeeze(1)
    st["sparse-conv pyramid (K7, 14 layers)"] = timed(lambda: sh.xyzc_net(fused, coord, out_sh))
    rows, dims, n_dev = sh.xyzc_net(fused, coord, out_sh)
    eng = r.engine_for(H, H, V, torch.device(DEV))
    st["rows -> fp16 volumes (K0 sparse)"] = timed(lambda: eng.upload_products_sparse(
        rows, dims, batch["featmaps"], batch["src_imgs"], n_rows_dev=n_dev))
    ms_all = timed(lambda: r.render(dict(batch)))
    live = [int(n) for n in n_dev]
    print(json.dumps({"config": tag, "H": H, "S": S, "V": V, "precision": "bf16 heads, fp32 sparse conv",
                      "ms_per_frame_render_call": ms_all, "rays": out["counts"]["n_rays"], "level_rows": live,
                      "stages_ms": {k: round(v, 4) for k, v in st.items()},
                      "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


def encoder(tag, H, V, precision):
    """Row f2: the image encoder on V source views of H×H (120.6 GFLOP at V=3, 512²)."""
    from gpnerf_b200.encoder import ResUNet
    enc = synth.fill_encoder_params(ResUNet(precision=precision), seed=42).eval().to(DEV)
    x = (torch.rand(V, 3, H, H, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(DEV)
    for _ in range(3):
        y = enc(x)
    ms = timed(lambda: enc(x))
    enc.use_cuda_graph = False
    ms_eager = timed(lambda: enc(x))
    # the reference's own module structure in plain torch (channels-last, same dtype, same parameters)
    import torch.nn as nn
    import torch.nn.functional as F
    w = {k: v for k, v in enc._params(x.device).items()}
    dt = next(v for v in w.values() if v.dim() == 4).dtype

    def cv(t, wt, b, stride):
        k = wt.shape[-1]
        if k > 1:
            t = F.pad(t, ((k - 1) // 2,) * 4, mode="reflect")
        return F.conv2d(t, wt, b, stride)

    def inorm(t, pre):
        return F.instance_norm(t, weight=w[pre + ".weight"].to(dt), bias=w[pre + ".bias"].to(dt), eps=1e-5)

    def torch_forward():
        t = x.to(dt).contiguous(memory_format=torch.channels_last)
        t = F.relu(inorm(cv(t, w["conv1.weight"], None, 2), "bn1"))
        feats = []
        for name in ("layer1", "layer2", "layer3"):
            for i, blk in enumerate(getattr(enc, name)):
                pre = f"{name}.{i}"
                y = F.relu(inorm(cv(t, w[pre + ".conv1.weight"], None, blk.stride), pre + ".bn1"))
                y = inorm(cv(y, w[pre + ".conv2.weight"], None, 1), pre + ".bn2")
                if blk.downsample is not None:
                    t = inorm(cv(t, w[pre + ".downsample.0.weight"], None, blk.stride), pre + ".downsample.1")
                t = F.relu(y + t)
            feats.append(t)
        x1, x2, x3 = feats
        up = lambda z: F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=True)      # noqa: E731
        cbe = lambda z, pre: F.elu(inorm(cv(z, w[pre + ".conv.weight"], w[pre + ".conv.bias"], 1), pre + ".bn"))  # noqa: E731
        t = cbe(up(x3), "upconv3.conv")
        t = cbe(torch.cat([t, x2], 1), "iconv3")
        t = cbe(up(t), "upconv2.conv")
        t = cbe(torch.cat([t, x1], 1), "iconv2")
        return F.conv2d(t, w["out_conv.weight"], w["out_conv.bias"]).float()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        y_t = torch_forward()
        ms_torch = timed(torch_forward)
    dev = float((y_t - y).abs().max())
    print(json.dumps({"config": tag, "H": H, "V": V, "precision": precision, "ms_graph": ms, "ms_eager": ms_eager,
                      "ms_eager_plain_torch_ops": ms_torch, "max_abs_diff_vs_plain_torch": dev, "GFLOP": 120.6 * V / 3 * (H / 512) ** 2,
                      "TFLOP_per_s": 120.6 * V / 3 * (H / 512) ** 2 / ms, "out": list(y.shape),
                      "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


def from_images(tag, H, S, V, seed=42):
    """Rows f1 + f2 + the path: images → encoder → SMPL attention → sparse-conv pyramid → K1…K5, device time of
    the whole chain per frame (inputs resident; Renderer.render's own host work and D2H excluded)."""
    from gpnerf_b200.encoder import ResUNet
    from gpnerf_b200.nerfhead import NeRFHead
    from gpnerf_b200.render import Projector, Renderer
    torch.manual_seed(seed)
    scene = synth.make_scene("zju", H=H, W=H, V=V, seed=seed)
    head = NeRFHead(n_views=V, precision=PREC_BF16).eval()
    sd = head.state_dict()
    for k, v in synth.make_head_weights(V=V, seed=seed).items():
        sd[k].copy_(v)
    for k, v in sd.items():
        if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
            v.fill_(3.0)
    head.load_state_dict(sd)
    head = head.to(DEV)
    enc = synth.fill_encoder_params(ResUNet(), seed=seed).eval().to(DEV)
    r = Renderer(enc, head, is_train=False, n_samples=S, progressive=True, precision=PREC_BF16)
    batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in scene.items() if k not in ("levels", "featmaps")}
    out = r.render(dict(batch))
    eng = r.engine_for(H, H, V, torch.device(DEV))
    import ctypes as C
    from gpnerf_b200._lib import Frame
    pinned = torch.empty(C.sizeof(Frame), dtype=torch.uint8).pin_memory()

    def device_chain():
        b = dict(batch)
        fm, _ = r._upstream(b)
        eng.upload_products_sparse(b["levels_sparse"], b["level_dims"], fm, b["src_imgs"], n_rows_dev=b["levels_sparse_rows"])
        eng.run_progressive_graphed(eng.make_frame(b), with_k0=False, frame_src=pinned)
    for _ in range(3):
        device_chain()
    ms = timed(device_chain)
    ms_call = timed(lambda: r.render(dict(batch)))
    print(json.dumps({"config": tag, "H": H, "S": S, "V": V, "precision": "fp16 encoder, fp32 sparse conv, bf16 heads",
                      "ms_per_frame_device_chain": ms, "frames_per_s": 1e3 / ms, "ms_per_render_call": ms_call,
                      "rays": out["counts"]["n_rays"], "rays_per_s": out["counts"]["n_rays"] * 1e3 / ms,
                      "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


CONFIGS = {
    "from_images512": lambda: from_images("frame from source images: encoder + attention + pyramid + render (f2+f1+a)", 512,
                                          64, 3),
    "encoder512": lambda: encoder("image encoder, 3 views 512x512, fp16 cuDNN + K9 norms (row f2)", 512, 3, "fp16"),
    "encoder512_bf16": lambda: encoder("image encoder, 3 views 512x512, bf16", 512, 3, "bf16"),
    "encoder512_fp32": lambda: encoder("image encoder, 3 views 512x512, fp32 parity mode", 512, 3, "fp32"),
    "upstream512": lambda: upstream("frame from feature maps: SMPL attention + sparse-conv pyramid + render (row f1)", 512,
                                    64, 3),
    "zju512_fp32": lambda: progressive("zju512_fp32 (configs[1] geometry, fp32 parity heads)", 512, 64, 3, PREC_FP32,
                                       graph=False),
    "thu512": lambda: progressive("trainthu_valzju shape (configs[2]): same hot-path tensors, other seed", 512, 64, 3,
                                  PREC_BF16, seed=1234),
    "dense512": lambda: dense("dense 512x512 BaseRender path (worst case, no compaction)", 512, 64, 3, PREC_BF16),
    "train4096": lambda: train("training step fwd+bwd, 4096 rays x 64 samples (configs[3] on one GPU)", 512, 64, 3, 4096),
    "train512": lambda: train("training step fwd+bwd, 512 rays x 64 samples (configs[3]: one GPU's share of 8)", 512, 64,
                              3, 512),
    "train4096_tf32": lambda: train("training step fwd+bwd, 4096 rays x 64 samples, TF32 tcgen05 heads", 512, 64, 3, 4096,
                                    precision=1),
    "train512_tf32": lambda: train("training step fwd+bwd, 512 rays x 64 samples, TF32 tcgen05 heads", 512, 64, 3, 512,
                                   precision=1),
    "zju1024": lambda: progressive("1024x1024, S=128, V=4 (configs[4] single-GPU share)", 1024, 128, 4, PREC_BF16),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        t0 = time.time()
        try:
            CONFIGS[n]()
        except Exception as e:   # keep going, report
            print(json.dumps({"config": n, "error": repr(e)}), flush=True)
        torch.cuda.empty_cache()
        print(f"# {n}: {time.time() - t0:.1f} s wall", file=sys.stderr, flush=True)
