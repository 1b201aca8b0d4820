"""BASELINE configs[3] on N GPUs: one training step = 4096 rays x 64 samples split evenly over the ranks,
forward + backward through the library's kernels (TF32 tcgen05 heads by default), head gradients averaged
with ONE flat all-reduce (train.GradBucket).  Run under torchrun; rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29530 \
      tools/gpu_train_ddp.py [--rays 4096] [--steps 10] [--fp32]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200.engine import Engine  # noqa: E402
from gpnerf_b200.train import GradBucket, render_dense_autograd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    H, S, V = 512, 64, 3
    R = args.rays // world
    scene = synth.make_scene("zju", H=H, W=H, V=V, seed=42, with_rays=True)
    w0 = synth.make_head_weights(V=V, seed=42, random_bias=True)
    n_all = scene["ray_o"].shape[1]
    sel = ((torch.arange(args.rays) * max(1, n_all // args.rays)) % n_all)[rank * R:(rank + 1) * R]   # this rank's rays
    rays = tuple(scene[k][0][sel].to(dev) for k in ("ray_o", "ray_d", "near", "far"))
    eng = Engine(H, H, S, V, device=dev, max_rays=R)
    w_g = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in w0.items()}
    lv = [t.to(dev) for t in scene["levels"]]
    fm, im = scene["featmaps"].to(dev), scene["src_imgs"].to(dev)
    eng.set_weights(w0)
    eng.upload_products(lv, fm, im)
    frame = eng.make_frame(scene)
    target = torch.rand(R, 3, device=dev)
    bucket = GradBucket(w_g.values())
    gen = torch.Generator().manual_seed(rank)

    def step():
        bucket.zero()
        t_rand = torch.rand(R, S, generator=gen)
        out = render_dense_autograd(eng, frame, rays, lv, fm, im, w_g, t_rand=t_rand, precision=0 if args.fp32 else 1)
        loss = ((out["rgb_map"] - target) ** 2).mean()
        loss.backward()
        bucket.all_reduce_mean()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = step()
    b.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    chk = bucket.flat.double().sum().reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((hi - lo).abs() <= 1e-9 * hi.abs().clamp_min(1e-30))
    else:
        same = True
    if rank == 0:
        print(json.dumps({"config": "training step fwd+bwd (configs[3])", "n_gpus": world, "rays_per_step": R * world,
                          "rays_per_gpu": R, "samples": S, "precision": "fp32" if args.fp32 else "tf32 tcgen05 heads",
                          "ms_per_step": float(ms), "rays_per_s": R * world * 1e3 / float(ms),
                          "grad_values_all_reduced": int(bucket.flat.numel()), "grads_identical_on_all_ranks": same,
                          "loss": float(loss)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
