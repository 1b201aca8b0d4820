#!/usr/bin/env python
"""Benchmark of the GP-NeRF progressive render hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scene zju|dense]
                  [--precision fp32|bf16] [--impl ours|reference]

One step = one 512×512 novel-view frame of the synthetic ZJU-Mocap-shaped
scene (BASELINE.json configs[1]): layout of the upstream products (K0), pixel
mask + rays + box test (K1), occupancy compaction + gathers (K2), density head
(K3), progressive compaction (K4), colour head (K3), compositing (K5).

At N>1 (`--shard`):
  frames (default)  one frame per GPU and step – N consecutive frames of an orbit
                    sweep (rank r renders the ring camera at 45° + 2°·r);
                    every finished pixel tile is written by K5 straight into slot
                    r of rank 0's image buffer over NVLink (peer memory, no NCCL
                    on the data path).  Per-GPU work is fixed: weak scaling.
  tiles             ONE frame, its pixel tiles dealt over the ranks; K5 writes each
                    tile into every rank's image (all ranks hold the full frame).
                    Strong scaling of a 0.8 ms frame: the replicated K0/K1 passes
                    bound it.
`--collective nccl` replaces the peer-memory writes by one NCCL all_gather.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident
in HBM; `e2e` goes through Renderer.render(batch) with pinned host inputs and
a device→host read of the image.  `--impl reference` times the CPU oracle (the
restatement of the reference's PyTorch path, oracle/gpnerf_oracle.py) on the
host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rays/s"
WORKLOAD = ("zju-like synthetic frame 512x512, V=3, S=64 (BASELINE configs[1], trainzju_valzju inference shape), "
            "progressive path")
SWEEP_STEP_DEG = 2.0          # angular step between the frames of the N-view sweep (frames mode)
S_SAMPLES = 64
VIEWS = 3
RES = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="zju", choices=["zju", "dense"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--tile-px", type=int, default=64)
    ap.add_argument("--shard", default="frames", choices=["frames", "tiles"])
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--mode", default="render", choices=["render", "train"],
                    help="render: the progressive frame (BASELINE configs[1], the default); train: configs[3], one "
                         "4096-ray forward+backward step through Renderer.render + MSE + one flat gradient all-reduce")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling (tiles) sub-record")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Samples that arrived during [t0, t1] widened by one sampling period on each side (the GPU is under
        the same load there: warm-up steps before, per-stage timing steps after).  Waits until nvidia-smi has
        exited: its tear-down holds driver locks and must not overlap the end-to-end timing that follows."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.12 <= ts <= t1 + 0.12):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# per-point algorithmic work of each stage (SURVEY.md §8d / DESIGN.md)
def stage_work(counts, n_level_elems, V, n_map_elems=0, n_img_px=0):
    """Algorithmic work per launch.  Bandwidth-bound stages: COMPULSORY bytes – every distinct input
    byte once plus the outputs – so that achieved / peak is a physical HBM fraction (the gathers re-request the
    L2-resident volumes many times over; that re-request rate is reported separately as `requested_gbs`)."""
    P, P1, P2 = counts["n_rays"] * S_SAMPLES, counts["P1"], counts["P2"]
    fused_compulsory = 2.0 * n_level_elems / 4 + 2.0 * n_map_elems + 16.0 * n_img_px + 20.0 * P1
    return {
        "k3_color_gather_tc": ("tensor", 72160.0 * P2, "72,160 FLOP per surviving point (colour trunk)"),
        "k3_color_tiles_tc": ("tensor", 72160.0 * P1, "72,160 FLOP per P1 point (colour trunk on every tile that has a survivor; tile hand-off)"),
        "k0_level_to_channels_last": ("hbm", 8.0 * n_level_elems / 4, "all 4 calls: 4 B read + 4 B written per element"),
        "k0_products_to_f16": ("hbm", 6.0 * n_level_elems / 4 + 4.0 * n_level_elems / 4 / 32,
                               "4 B read + 2 B written per element, 4 B channel sum per voxel"),
        "k2_occupancy_compact": ("hbm", 36.0 * P, "32 B tap + 4 B z per point"),
        "k2_gather_volume": ("hbm", 4096.0 * P1, "4 levels x 8 corners x 32 ch x 4 B per point"),
        "k2_project_gather_meanvar": ("hbm", V * 4 * 35 * 4.0 * P1, "V x 4 corners x 35 ch x 4 B per point"),
        "k3_density_mlp": ("tensor", 38688.0 * P1, "38,688 FLOP per point"),
        "k3_color_mlp": ("tensor", 72160.0 * P2, "72,160 FLOP per point"),
        # bf16 path: gathers (bf16 storage: 4 levels x 8 corners x 64 B + V x 4 x (64 B + 16 B RGBx)) fused
        # with the density head; 16*(9+5V) B record written per point
        "k23_gather_density_tc": ("hbm", fused_compulsory,
                                  "compulsory bytes: the fp16 volumes, feature maps and RGBx images once (they are L2 "
                                  "resident), 12 B read + 8 B written per P1 point; the kernel REQUESTS 2,048 B volume + "
                                  "320 B/view per point from L1/L2 and runs 38,688 FLOP/point on the tensor pipe"),
        "k3_color_mlp_records": ("tensor", 72160.0 * P2, "72,160 FLOP per point"),
        "k4_compact_alpha": ("hbm", 8.0 * P1, "4 B read + 4 B written per point"),
        "k4_compact_alpha_fused": ("hbm", 0.125 * P1 + 4.0 * P2, "1 flag bit read per point, 4 B written per survivor"),
        "k5_composite": ("hbm", 16.0 * P1, "16 B per surviving sample"),
    }


def cpu_reference_frames(scene, weights, min_seconds, max_frames):
    """Time the CPU oracle (port of the reference's PyTorch path) on full frames."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpnerf_oracle as orc
    torch.set_num_threads(os.cpu_count())
    times, out = [], None
    t_all = time.perf_counter()
    while len(times) < max_frames and (time.perf_counter() - t_all < min_seconds or not times):
        t0 = time.perf_counter()
        out = orc.render_progressive(scene, weights, S=S_SAMPLES, chunk=131072, keep=True)
        times.append(time.perf_counter() - t0)
    return times, out


def run_reference(args, rank):
    """--impl reference: the reference's own algorithm on the host cores."""
    if rank != 0:
        return
    import gpnerf_b200  # noqa: F401
    from gpnerf_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpnerf_oracle as orc
    torch.set_num_threads(os.cpu_count())
    scene = synth.make_scene(args.scene if args.scene == "zju" else "zju", H=RES, W=RES, V=VIEWS, seed=42)
    w = synth.make_head_weights(V=VIEWS, seed=42)
    times, out = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = orc.render_progressive(scene, w, S=S_SAMPLES, chunk=131072)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    rays = out["n_rays"]
    val = rays * len(times) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": len(times) / total,
        "config": {"workload": WORKLOAD, "rays": rays, "P1": out["P1"], "P2": out["P2"]},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "every step = the full frame through oracle/gpnerf_oracle.py "
                                   "(torch CPU restatement of the reference; chunk 131072 points)"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_train(args, rank, world, local_rank):
    """--mode train: BASELINE configs[3] – one training step = 4096 rays x 64 samples split evenly over the ranks,
    forward + backward through the hot path's K6 kernels (TF32 tcgen05 heads), MSE against random targets, head
    gradients averaged with ONE flat all-reduce (train.GradBucket, NCCL over NVLink), AdamW step.  `value` = rays/s
    with the upstream products (levels, feature maps) resident – the hot path of SURVEY §8; `full_pipeline` repeats
    the step through `Renderer.render(batch)` from the source images (encoder + SMPL attention + sparse-conv
    pyramid in their training form, trainmode.py), i.e. the exact sequence of BaseTrainer.train."""
    import torch.distributed as dist
    import gpnerf_b200  # noqa: F401
    from gpnerf_b200 import synth, train
    from gpnerf_b200.encoder import ResUNet
    from gpnerf_b200.engine import Engine
    from gpnerf_b200.nerfhead import NeRFHead
    from gpnerf_b200.render import Renderer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("GPNERF_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    n_rays_total, S, V = 4096, S_SAMPLES, VIEWS
    R = n_rays_total // world
    scene = synth.make_scene("zju", H=RES, W=RES, V=V, seed=42, with_rays=True)
    w0 = synth.make_head_weights(V=V, seed=42, random_bias=True)
    n_all = scene["ray_o"].shape[1]
    sel = ((torch.arange(n_rays_total) * max(1, n_all // n_rays_total)) % n_all)[rank * R:(rank + 1) * R]
    rays = tuple(scene[k][0][sel].to(dev) for k in ("ray_o", "ray_d", "near", "far"))
    eng = Engine(RES, RES, S, V, device=dev, max_rays=R)
    w_g = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in w0.items()}
    lv = [t.to(dev) for t in scene["levels"]]
    fm, im = scene["featmaps"].to(dev), scene["src_imgs"].to(dev)
    eng.set_weights(w0)
    eng.upload_products(lv, fm, im)
    frame = eng.make_frame(scene)
    target = torch.rand(R, 3, device=dev)
    bucket = train.GradBucket(w_g.values())
    opt = torch.optim.AdamW(list(w_g.values()), lr=1e-4, capturable=True)
    gen = torch.Generator().manual_seed(rank)
    t_pin = torch.empty(R, S).pin_memory()

    def device_step():
        bucket.zero()
        out = train.render_dense_autograd(eng, frame, rays, lv, fm, im, w_g, t_rand=t_pin.to(dev, non_blocking=True),
                                          precision=train.PREC_TRAIN_TF32)
        loss = ((out["rgb_map"] - target) ** 2).mean()
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss
    # the whole step as one CUDA graph (train.GraphedStep); --no-graph launches the kernels one by one
    graphed, graph_note = None, "kernels launched one by one (--no-graph)"
    if world > 1 and not os.environ.get("GPNERF_TRAIN_GRAPH_MULTI"):
        # the only attempt at capturing the step with the NCCL all-reduce inside it on 8 ranks hung (round 2, no GPU
        # budget left to find out why): ranks > 1 launch eagerly unless GPNERF_TRAIN_GRAPH_MULTI=1
        graph_note = "kernels launched one by one (CUDA-graph capture of the step is used on one GPU only)"
    elif not args.no_graph:
        try:
            t_pin.copy_(torch.rand(R, S, generator=gen))
            graphed = train.GraphedStep(device_step, dev)
            graph_note = "forward + backward + all-reduce + AdamW replayed as one CUDA graph per step (train.GraphedStep)"
        except Exception as e:       # noqa: BLE001 - report and fall back to eager launches
            graphed, graph_note = None, f"CUDA-graph capture failed ({type(e).__name__}: {e}); eager launches"

    def step():
        t_pin.copy_(torch.rand(R, S, generator=gen))                 # the jitter is drawn on the host (BaseRender.py:40-47)
        return graphed() if graphed is not None else device_step()

    def time_steps(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for a, b in evs:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / k, 1e3 * wall / k], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    ms_dev, ms_wall = time_steps(step, args.steps)
    clocks = sampler.stop(wall0, time.perf_counter()) if rank == 0 else None
    chk = bucket.flat.double().sum().reshape(1)
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((hi - lo).abs() <= 1e-9 * hi.abs().clamp_min(1e-30))
    # ---- the same step through the plugin API from the source images (BaseTrainer._forward's sequence)
    full = None
    try:
        torch.manual_seed(42)
        head = NeRFHead(code_dim=16, n_views=V).to(dev)
        sd = head.state_dict()
        for k, v in w0.items():
            if k in sd and sd[k].shape == v.shape:
                sd[k].copy_(v)
        for k, v in sd.items():
            if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
                v.fill_(3.0)
        head.load_state_dict(sd)
        enc = synth.fill_encoder_params(ResUNet(), seed=42).to(dev)
        r = Renderer(enc, head, is_train=True, n_samples=S, progressive=False).to(dev).train()
        batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in scene.items() if k not in ("levels", "featmaps")}
        for k in ("ray_o", "ray_d", "near", "far"):
            batch[k] = batch[k][:, sel.to(dev)]
        bucket2 = train.GradBucket(r.parameters())
        opt2 = torch.optim.AdamW(list(r.parameters()), lr=1e-4)

        def step_full():
            bucket2.zero()
            ret = r.render(batch)
            loss = ((ret["rgb_map"][0] - target) ** 2).mean()
            loss.backward()
            bucket2.all_reduce_mean()
            opt2.step()
        for _ in range(3):
            step_full()
        f_dev, f_wall = time_steps(step_full, max(3, args.steps // 2))
        full = {"ms_per_step": f_wall, "device_ms_per_step": f_dev, "value": n_rays_total * 1e3 / f_wall, "unit": "rays/s",
                "trainable_values_all_reduced": int(bucket2.flat.numel()),
                "what": "Renderer.render(batch) from the source images + SMPL fit, training mode: encoder, SMPL attention "
                        "and sparse-conv pyramid in training form (torch autograd ops, batch-statistics BatchNorm), "
                        "hot path in the K6 kernels; MSE; one flat all-reduce over ALL parameters; AdamW"}
    except Exception as exc:                # the hot-path number stands on its own
        full = {"error": repr(exc)[:300]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    emit({
        "metric": METRIC, "value": n_rays_total * 1e3 / ms_wall, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_wall, "device_ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic", "mode": "train",
        "config": {"workload": "training step, 4096 rays x 64 samples per step split evenly over the ranks, forward + backward "
                               "through gathers, both heads and raw2outputs (BASELINE configs[3]); zju-like synthetic "
                               "scene 512x512, V=3; products resident, head parameters trainable",
                   "rays_per_gpu": R, "l2": "working set ≈10 GB of activations per step ≫ 126 MB L2",
                   "collective": "one flat all-reduce of the head gradients per step (train.GradBucket)" if world > 1 else "none",
                   "grad_values_all_reduced": int(bucket.flat.numel()), "grads_identical_on_all_ranks": same,
                   "launch": graph_note},
        "clocks": clocks, "full_pipeline": full, "gpu_launches": None,
        "e2e": {"value": n_rays_total * 1e3 / ms_wall, "unit": "rays/s", "h2d_bytes_per_step": int(R * S * 4),
                "d2h_bytes_per_step": 0, "what": "host wall clock around the steps; the per-step jitter is uploaded from pinned memory"},
    })


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to
    stdout on the first collective): fd 1 is pointed at stderr for the whole run and the JSON line goes to a
    private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.mode == "train":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        run_train(args, rank, world, local_rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    import torch.distributed as dist
    import gpnerf_b200  # noqa: F401
    from gpnerf_b200 import shard, synth
    from gpnerf_b200._lib import PREC_BF16, PREC_FP32
    from gpnerf_b200.nerfhead import NeRFHead
    from gpnerf_b200.render import Renderer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("GPNERF_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    prec = PREC_BF16 if args.precision == "bf16" else PREC_FP32

    scene = synth.make_scene("zju", H=RES, W=RES, V=VIEWS, seed=42)
    frames_mode = world > 1 and args.shard == "frames"
    if frames_mode:          # rank r renders frame r of an orbit sweep: consecutive novel views, 2° apart
        scene = synth.retarget(scene, 45.0 + SWEEP_STEP_DEG * rank)
    weights = synth.make_head_weights(V=VIEWS, seed=42)
    head = NeRFHead(code_dim=32, n_views=VIEWS, precision=prec)
    sd = head.state_dict()
    sd.update({k: v for k, v in weights.items()})
    head.load_state_dict(sd)
    head = head.to(dev)
    renderer = Renderer(None, head, is_train=False, n_samples=S_SAMPLES, progressive=True, precision=prec,
                        rank=rank, world=world, tile_px=args.tile_px, use_cuda_graph=not args.no_graph,
                        shard=args.shard, collective=args.collective)
    n_px = RES * RES

    # ---- device-resident inputs for `value`
    host_keys = ("levels", "featmaps", "src_imgs")
    d_levels = [t.to(dev) for t in scene["levels"]]
    d_feat, d_imgs = scene["featmaps"].to(dev), scene["src_imgs"].to(dev)
    eng = renderer.engine_for(RES, RES, VIEWS, dev)
    eng.set_weights(head.hot_path_state())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    eng.set_static_inputs(d_levels, d_feat, d_imgs)
    eng.upload_products(d_levels, d_feat, d_imgs)
    frame = eng.make_frame(scene)

    def step_eager():
        eng.upload_products(d_levels, d_feat, d_imgs)
        eng.render_progressive(frame)

    def step_device():
        if args.no_graph:
            step_eager()
        else:
            eng.run_progressive_graphed(frame)      # K0…K5 as one CUDA-graph launch
        if world > 1 and eng.exchange is None:          # --collective nccl
            if frames_mode:
                out = torch.empty(world * n_px, 3, device=dev)
                dist.all_gather_into_tensor(out, eng.pred_img.view(n_px, 3))
                return out
            return shard.gather_frame(eng.pred_img.view(n_px, 3), RES, args.tile_px)
        return eng.result_image()

    # nvidia-smi starts here so that its start-up (driver locks) falls into the warm-up, not the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step_device()
    torch.cuda.synchronize(dev)
    # keep stepping (untimed, all ranks together) until the first clock sample has arrived, so that samples
    # bracket the timed region with the GPU under the same load
    t_wait = time.perf_counter()
    while True:
        go = torch.tensor([int(rank == 0 and sampler.proc is not None and not sampler.lines and
                               time.perf_counter() - t_wait < 3.0)], device=dev)
        if world > 1:
            dist.broadcast(go, 0)
        if not int(go.item()):
            break
        for _ in range(5):
            flush.zero_()
            step_device()
        torch.cuda.synchronize(dev)
    counts = eng.read_counters()       # (also settles the engine's auto hand-off to the colour head: tiles / gather)
    for _ in range(2):                 # untimed: a changed hand-off is a new CUDA graph, captured here
        flush.zero_()
        step_device()
    torch.cuda.synchronize(dev)
    counts = eng.read_counters()
    ref_image = eng.result_image().cpu().clone()
    ref_hit = eng.result_hit_mask().cpu().clone()
    ref_valid1 = eng.valid1[: counts["P2"]].cpu().clone()
    if world > 1:
        tot = torch.tensor([counts["n_rays"], counts["P1"], counts["P2"]], device=dev)
        dist.all_reduce(tot)
        g_rays, g_p1, g_p2 = [int(v) for v in tot.tolist()]
    else:
        g_rays, g_p1, g_p2 = counts["n_rays"], counts["P1"], counts["P2"]
    # did the tiles written by the peers land?  (hit flags in rank 0's buffers vs. the ranks' ray counts)
    peer_check = None
    if world > 1:
        mine = torch.tensor([counts["n_rays"]], device=dev, dtype=torch.long)
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
        if eng.exchange is not None and rank == 0:
            if frames_mode:
                ok = all(int(eng.exchange.hit_mask(slot=r).sum()) == int(per_rank[r]) for r in range(world))
            else:
                ok = int(eng.exchange.hit_mask().sum()) == g_rays
            peer_check = "ok" if ok else "MISMATCH"

    # ---- timed region: exactly K steps, device-timed, L2 flushed between steps
    launches0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record()
        step_device()
        b.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop(wall0, wall0 + wall) if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = (eng.launches - launches0) // args.steps
    # per-stage device times: CUDA events cannot sit inside a graph replay, so the same steps are
    # re-issued kernel by kernel (same stream, same L2 flush) right after the timed region
    eng.timing = True
    eng.stage_events = {}
    for _ in range(min(args.steps, 10)):
        flush.zero_()
        step_eager()
    torch.cuda.synchronize(dev)
    eng.timing = False
    stage_ms = eng.stage_times_ms()
    # ---- the same frame with the levels handed over as the sparse-conv net holds them before .dense()
    #      (active rows: features + voxel indices): K0's transposition of a >95 % empty
    #      dense volume becomes a scatter.  Reported beside the headline, which keeps the reference's dense layout.
    sparse_dev = None
    if prec == PREC_BF16 and not args.no_graph:
        import ctypes as C
        from gpnerf_b200._lib import Frame
        lv_s, dims_s = synth.sparsify_levels(scene["levels"])
        d_lv_s = [(f.to(dev), i.to(dev)) for f, i in lv_s]
        fpin = torch.empty(C.sizeof(Frame), dtype=torch.uint8).pin_memory()

        def step_sparse():
            eng.upload_products_sparse(d_lv_s, dims_s, d_feat, d_imgs)
            eng.run_progressive_graphed(frame, with_k0=False, frame_src=fpin)
        for _ in range(3):
            flush.zero_()
            step_sparse()
        torch.cuda.synchronize(dev)
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        if world > 1:
            dist.barrier()
        for a, b in ev2:
            flush.zero_()
            a.record()
            step_sparse()
            b.record()
        torch.cuda.synchronize(dev)
        ms_s = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
        same = torch.equal(eng.result_image().cpu(), ref_image) if world == 1 else None
        sparse_dev = {"ms_per_step": float(ms_s), "value": g_rays / (float(ms_s) * 1e-3), "unit": "rays/s",
                      "active_rows": [int(f.shape[0]) for f, _ in lv_s],
                      "input_bytes": int(sum(f.numel() * 4 + i.numel() * 4 for f, i in lv_s)),
                      "image_identical_to_dense_route": same}
    if world > 1:
        t = torch.tensor([dev_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = g_rays / (ms_per_step * 1e-3)
    # per-rank share of the work (frames mode: the views differ, the slowest one paces every step)
    per_rank = None
    if world > 1:
        own = sum(v for k, v in stage_ms.items() if k != "peer_wait")
        mine = torch.tensor([counts["n_rays"], counts["P1"], counts["P2"], own, stage_ms.get("peer_wait", 0.0)],
                            device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rays": int(t[0]), "P1": int(t[1]), "P2": int(t[2]), "stages_ms_sum": round(float(t[3]), 4),
                     "peer_wait_ms": round(float(t[4]), 4)} for t in allr]

    # ---- strong scaling of ONE frame (north_star: "rays are sharded evenly across the GPUs"): the frame's pixel tiles
    #      dealt over the ranks (shard.py), every rank ends with the full image.  Two frames: the headline one and
    #      BASELINE configs[4]'s shape (1024², S=128, V=4).  At N=1 the same code gives the single-GPU reference, so the
    #      per-N records of a scaling run can be divided directly.
    strong = None
    if not args.no_strong and prec == PREC_BF16 and not args.no_graph:
        strong = {}
        for tag, res, views, samples in (("zju512_v3_s64", RES, VIEWS, S_SAMPLES), ("zju1024_v4_s128", 1024, 4, 128)):
            sc = synth.make_scene("zju", H=res, W=res, V=views, seed=42)
            wts = synth.make_head_weights(V=views, seed=42)
            hd = NeRFHead(code_dim=32, n_views=views, precision=prec)
            sdd = hd.state_dict()
            sdd.update({k: v for k, v in wts.items()})
            hd.load_state_dict(sdd)
            rt = Renderer(None, hd.to(dev), is_train=False, n_samples=samples, progressive=True, precision=prec,
                          rank=rank, world=world, tile_px=args.tile_px, shard="tiles", collective=args.collective)
            et = rt.engine_for(res, res, views, dev)
            et.set_weights(hd.hot_path_state())
            lv_t = [t.to(dev) for t in sc["levels"]]
            fm_t, im_t = sc["featmaps"].to(dev), sc["src_imgs"].to(dev)
            et.set_static_inputs(lv_t, fm_t, im_t)
            et.upload_products(lv_t, fm_t, im_t)
            fr_t = et.make_frame(sc)

            def step_t():
                et.run_progressive_graphed(fr_t)
                if world > 1 and et.exchange is None:
                    return shard.gather_frame(et.pred_img.view(res * res, 3), res, args.tile_px)
                return et.result_image()
            for _ in range(3):
                flush.zero_()
                step_t()
            # auto hand-off to the colour head: the engine settles on tiles / gather from this frame's survivor ratio
            # (every rank of a tiles-mode frame decides from its own share); the steps after it are still warm-up
            et.read_counters()
            for _ in range(2):
                flush.zero_()
                step_t()
            torch.cuda.synchronize(dev)
            evt = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            if world > 1:
                dist.barrier()
            for a, b in evt:
                flush.zero_()
                a.record()
                step_t()
                b.record()
            torch.cuda.synchronize(dev)
            ms_t = torch.tensor([sum(a.elapsed_time(b) for a, b in evt) / args.steps], device=dev)
            ct = et.read_counters()
            tot_t = torch.tensor([ct["n_rays"], ct["P1"], ct["P2"]], device=dev)
            if world > 1:
                dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
                dist.all_reduce(tot_t)
            hits = int(et.result_hit_mask().sum())
            strong[tag] = {"ms_per_frame": float(ms_t), "rays": int(tot_t[0]), "P1": int(tot_t[1]), "P2": int(tot_t[2]),
                           "value": int(tot_t[0]) / (float(ms_t) * 1e-3), "unit": "rays/s", "frames_per_s": 1e3 / float(ms_t),
                           "full_image_on_every_rank": bool(hits == int(tot_t[0])),
                           "sharding": f"pixel tiles of {args.tile_px} dealt diagonally over {world} rank(s); " +
                                       ("K5 writes every tile into all ranks' images over NVLink (peer memory)"
                                        if et.exchange is not None else "one NCCL all_gather" if world > 1 else "single GPU")}
            del rt, et, lv_t, fm_t, im_t
            torch.cuda.empty_cache()

    # ---- e2e: Renderer.render(batch) with pinned host inputs, image read back
    e2e = None
    if not args.no_e2e:
        batch = {k: v for k, v in scene.items() if torch.is_tensor(v)}
        batch["levels"] = [t.pin_memory() for t in scene["levels"]]
        batch["featmaps"] = scene["featmaps"].pin_memory()
        batch["src_imgs"] = scene["src_imgs"].pin_memory()
        h2d = sum(t.numel() * 4 for t in batch["levels"]) + batch["featmaps"].numel() * 4 + batch["src_imgs"].numel() * 4

        def e2e_step():
            b = dict(batch)
            # render() takes its device from src_imgs; levels/featmaps stay on the host so that
            # their copies happen inside the call
            b["src_imgs"] = batch["src_imgs"].to(dev, non_blocking=True)
            return renderer.render(b)          # gathers the tiles itself when world > 1

        e2e_windows = {}

        def timed(fn, tag=None):
            """Wall time of K end-to-end steps (max over ranks).  The host-side legs see one-off stalls of
            100-300 ms on some boxes (a third of the runs, any leg): three windows of K steps each, the median
            reported, all three listed under e2e.windows_ms_per_step."""
            ets = []
            for _ in range(3):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize(dev)
                if world > 1:
                    dist.barrier()
                et = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([et], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    et = float(t.item())
                ets.append(et)
            if tag:
                e2e_windows[tag] = [round(1e3 * e / args.steps, 4) for e in ets]
            return sorted(ets)[1]
        # (1) one blocking Renderer.render(batch) call per step (the reference's calling convention)
        for _ in range(6):
            e2e_step()
        et_sync = timed(lambda: [e2e_step() for _ in range(args.steps)], "blocking_call")
        if os.environ.get("GPNERF_PROFILE_BLOCKING"):
            import cProfile
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(args.steps):
                e2e_step()
            torch.cuda.synchronize(dev)
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(18)
        # (2) the same frames through Renderer.render_stream: uploads of the next frames overlap the render of
        #     the current one (every frame's 134 MB still cross PCIe inside the timed region, every image
        #     is read back into host memory)
        stream_ok = not (world > 1 and args.shard == "tiles")
        et = et_sync
        e2e_sparse = None
        if stream_ok:
            n_out = [0]

            def run_stream(k):
                for out in renderer.render_stream(batch for _ in range(k)):
                    n_out[0] += int(out["mask_at_box"].sum() > 0)
            run_stream(8)
            et = timed(lambda: run_stream(args.steps), "render_stream")
            if os.environ.get("GPNERF_PROFILE_BLOCKING"):
                import cProfile
                import pstats
                t0p, arr = time.perf_counter(), []
                for out in renderer.render_stream(batch for _ in range(args.steps)):
                    arr.append(round((time.perf_counter() - t0p) * 1e3, 1))
                print("stream arrivals ms", arr, file=sys.stderr)
                pr = cProfile.Profile()
                pr.enable()
                run_stream(args.steps)
                pr.disable()
                pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(16)
            if prec == PREC_BF16:
                lv_s, dims_s = synth.sparsify_levels(scene["levels"])
                sbatch = {k: v for k, v in batch.items() if k != "levels"}
                sbatch["levels_sparse"] = [(f.pin_memory(), i.pin_memory()) for f, i in lv_s]
                sbatch["level_dims"] = dims_s
                h2d_s = (sum(f.numel() * 4 + i.numel() * 4 for f, i in lv_s) + batch["featmaps"].numel() * 4 +
                         batch["src_imgs"].numel() * 4)

                def run_stream_sparse(k):
                    for out in renderer.render_stream(sbatch for _ in range(k)):
                        n_out[0] += int(out["mask_at_box"].sum() > 0)
                run_stream_sparse(8)
                et_s = timed(lambda: run_stream_sparse(args.steps), "sparse_levels")
                def blocking_sparse():
                    b = dict(sbatch)
                    b["src_imgs"] = sbatch["src_imgs"].to(dev, non_blocking=True)
                    return renderer.render(b)
                for _ in range(4):
                    blocking_sparse()
                et_sb = timed(lambda: [blocking_sparse() for _ in range(args.steps)], "sparse_levels_blocking_call")
                e2e_sparse = {"ms_per_step": 1e3 * et_s / args.steps, "value": g_rays * args.steps / et_s,
                              "blocking_call_ms_per_step": 1e3 * et_sb / args.steps,
                              "unit": "rays/s", "h2d_bytes_per_step": int(h2d_s),
                              "what": "render_stream with the 4 levels as sparse rows (features + indices) instead of "
                                      "dense NCDHW tensors"}
        # (3) nothing precomputed: the batch carries the source images and the SMPL fit only; encoder (f2),
        #     SMPL attention + sparse-conv pyramid (f1) and the path itself all run inside Renderer.render
        e2e_images = None
        if world == 1 and prec == PREC_BF16 and not args.no_graph:
            from gpnerf_b200.encoder import ResUNet
            torch.manual_seed(42)
            head2 = NeRFHead(code_dim=32, n_views=VIEWS, precision=prec).eval()
            sd2 = head2.state_dict()
            sd2.update({k: v for k, v in weights.items()})
            for k, v in sd2.items():        # random-init BatchNorm scales would let the 14-layer pyramid die out
                if "xyzc_net" in k and (k.endswith(".1.weight") or k.endswith(".4.weight")):
                    v.fill_(3.0)
            head2.load_state_dict(sd2)
            enc = synth.fill_encoder_params(ResUNet(), seed=42).eval().to(dev)
            r2 = Renderer(enc, head2.to(dev), is_train=False, n_samples=S_SAMPLES, progressive=True, precision=prec)
            ibatch = {k: v for k, v in batch.items() if k not in ("levels", "featmaps")}
            h2d_i = sum(v.numel() * v.element_size() for v in ibatch.values() if torch.is_tensor(v))

            def images_step():
                b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in ibatch.items()}
                return r2.render(b)
            for _ in range(6):
                out_i = images_step()
            et_i = timed(lambda: [images_step() for _ in range(args.steps)], "from_images")
            if os.environ.get("GPNERF_PROFILE_BLOCKING"):
                import cProfile
                import pstats
                pr = cProfile.Profile()
                pr.enable()
                for _ in range(args.steps):
                    images_step()
                pr.disable()
                pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(25)
            rays_i = int(out_i["counts"]["n_rays"])

            def images_stream(k):
                for out in r2.render_stream(ibatch for _ in range(k)):
                    n_out[0] += int(out["mask_at_box"].sum() > 0)
            images_stream(8)
            et_is = timed(lambda: images_stream(args.steps), "from_images_stream")
            if os.environ.get("GPNERF_PROFILE_BLOCKING"):
                pr = cProfile.Profile()
                pr.enable()
                images_stream(args.steps)
                pr.disable()
                pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(22)
            e2e_images = {"ms_per_step": 1e3 * et_i / args.steps, "value": rays_i * args.steps / et_i, "unit": "rays/s",
                          "rays": rays_i, "h2d_bytes_per_step": int(h2d_i),
                          "stream_ms_per_step": 1e3 * et_is / args.steps, "stream_value": rays_i * args.steps / et_is,
                          "what": "Renderer.render(batch) with only src_imgs + the SMPL fit in the batch (host tensors): "
                                  "image encoder, SMPL-code attention, sparse-conv pyramid and K1..K5 all inside the "
                                  "call; random-init producers, so the ray count differs from the synthetic-volume frame; "
                                  "stream_* = the same batches through Renderer.render_stream"}
        # headline: the levels handed over as the sparse-conv net holds them (active rows) – what the upstream producer
        # owns before `.dense()`; the dense NCDHW hand-off (119 MB of >95 % zeros per frame over PCIe) stays beside it
        dense_leg = {"value": g_rays * args.steps / et, "ms_per_step": 1e3 * et / args.steps, "h2d_bytes_per_step": int(h2d),
                     "what": "the same stream with the 4 levels as dense NCDHW fp32 tensors (SparseConvTensor.dense() layout)"}
        if e2e_sparse is not None:
            et_head, h2d_head = e2e_sparse["ms_per_step"] * 1e-3 * args.steps, e2e_sparse["h2d_bytes_per_step"]
        else:
            et_head, h2d_head = et, h2d
        e2e = {"value": g_rays * args.steps / et_head, "unit": "rays/s", "h2d_bytes_per_step": int(h2d_head),
               "d2h_bytes_per_step": int(n_px * 3 * 4 + n_px + 32), "ms_per_step": 1e3 * et_head / args.steps,
               "frames_per_s": args.steps * (world if frames_mode else 1) / et_head,
               "api": ("gpnerf_b200.render.Renderer.render_stream(batches) – sparse level rows / featmaps / src_imgs of "
                       "every frame in pinned host memory, uploads of the following frames overlapped with the render, "
                       "every image read back to the host") if stream_ok else
                      "gpnerf_b200.render.Renderer.render(batch) – levels/featmaps/src_imgs in pinned host memory",
               "dense_levels": dense_leg,
               "blocking_call_ms_per_step": 1e3 * et_sync / args.steps,
               "blocking_call_api": "gpnerf_b200.render.Renderer.render(batch), one blocking call per frame",
               "sparse_levels": e2e_sparse, "from_images": e2e_images,
               "timing": "median of 3 windows of K steps each (host wall clock around the public API calls)",
               "windows_ms_per_step": e2e_windows}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (this rank's share of the work)
    pk = peaks()
    n_level_elems = sum(t.numel() for t in scene["levels"]) * 4
    n_map_elems = scene["featmaps"].numel()
    n_img_px = VIEWS * (RES + 2) * (RES + 2)
    work = stage_work(counts, n_level_elems, VIEWS, n_map_elems, n_img_px)
    timed = {k: v for k, v in stage_ms.items() if k in work}
    top = max(timed, key=timed.get) if timed else None
    tj = {}
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            tj["_file"] = name
            break
    flops_of = {"k23_gather_density_tc": 38688.0 * counts["P1"], "k3_color_gather_tc": 72160.0 * counts["P2"],
                "k3_color_mlp_records": 72160.0 * counts["P2"], "k3_color_tiles_tc": 72160.0 * counts["P1"]}
    binding = {
        "k23_gather_density_tc": "the producers' gathers: load latency x loads in flight (16 warps x 8 x 16 B per lane; long-scoreboard "
                                 "stall 4.9 warps per issue cycle) with the L1 data pipe at 73 % - what-if runs: producers only 0.28 ms, "
                                 "head only 0.125 ms (profiles/r02_ncu_summary.md section 8); neither HBM nor the tensor pipe bounds this kernel",
        "k3_color_gather_tc": "latency of the 4V+1 dependent MMA -> epilogue rounds of a tile (three chains per SM) and the re-gather",
        "k3_color_tiles_tc": "latency of the 4V+1 dependent MMA -> epilogue rounds of a tile, four chains per SM (TMEM: 4 x 128 columns); "
                             "floors: tensor pipe 0.124 ms (69 MMAs of ~75 cycles per tile), MUFU 0.10 ms, issue slots 0.094 ms",
    }
    requested_of = {"k23_gather_density_tc": (2048.0 + VIEWS * 320.0) * counts["P1"],
                    "k3_color_gather_tc": VIEWS * 320.0 * counts["P2"]}

    def kernel_roofline(name):
        """The contract's roofline object for one kernel: `achieved` = algorithmic (compulsory) bytes or FLOPs per
        launch / its CUDA-event duration; beside it the other physical fractions and what ncu names as the busiest
        unit, so that a latency- or L1-bound kernel is not mistaken for an HBM-bound one."""
        bound, amount, what = work[name]
        n_calls = 4 if name == "k0_level_to_channels_last" else 1
        dur_s = timed[name] * n_calls * 1e-3
        if bound == "hbm":
            ach, peak, unit = amount / dur_s / 1e9, pk["hbm_gbs"], "GB/s"
        else:
            ach, peak, unit = amount / dur_s / 1e12, pk["bf16_tflops"], "TFLOP/s"
        traffic = tj.get(name) if world == 1 else None
        r = {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
             "traffic": traffic,
             "traffic_source": (f"ncu --set full, same workload (profiles/{tj.get('_file')})" if traffic else None),
             "algorithmic": what, "algorithmic_bytes_or_flops": amount, "launch_ms": timed[name] * n_calls,
             "peak_source": pk["source"], "ncu": tj.get("ncu", {}).get(name)}
        if traffic:
            r["frac_dram"] = traffic / dur_s / 1e9 / pk["hbm_gbs"]
        if name in flops_of:
            r["tflops"] = flops_of[name] / dur_s / 1e12
            r["frac_tensor"] = r["tflops"] / pk["bf16_tflops"]
        if name in requested_of:
            r["requested_gbs"] = requested_of[name] / dur_s / 1e9
        if name in binding:
            r["binding_unit"] = binding[name]
        return r
    roofline = kernel_roofline(top) if top else None
    if roofline is not None:
        others = [k for k in ("k23_gather_density_tc", "k3_color_gather_tc", "k3_color_tiles_tc") if k in timed and k != top]
        roofline["other_head_kernels"] = [kernel_roofline(k) for k in others]
        # the whole frame against both roofs: every distinct input byte once + the image, and the heads' FLOPs
        frame_bytes = (6.0 * n_level_elems / 4 + 6.0 * n_map_elems + 28.0 * n_img_px + 36.0 * counts["n_rays"] * S_SAMPLES +
                       40.0 * counts["P1"] + 16.0 * RES * RES)
        frame_flops = 38688.0 * counts["P1"] + 72160.0 * counts["P2"]
        roofline["frame"] = {"compulsory_dram_bytes": frame_bytes, "flops": frame_flops, "ms": ms_per_step,
                             "frac_hbm": frame_bytes / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "frac_tensor": frame_flops / (ms_per_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                             "lower_bound_ms": max(frame_bytes / (pk["hbm_gbs"] * 1e9), frame_flops / (pk["bf16_tflops_sustained"] * 1e12)) * 1e3}
    stages_out = {}
    for k, ms in sorted(stage_ms.items(), key=lambda kv: -kv[1]):
        ent = {"ms": round(ms * (4 if k == "k0_level_to_channels_last" else 1), 4)}
        if k in work:
            bound, amount, _ = work[k]
            dur = ent["ms"] * 1e-3
            if bound == "hbm":
                ent["GB/s"] = round(amount / dur / 1e9, 1); ent["frac_hbm"] = round(amount / dur / 1e9 / pk["hbm_gbs"], 4)
            else:
                ent["TFLOP/s"] = round(amount / dur / 1e12, 2); ent["frac_bf16"] = round(amount / dur / 1e12 / pk["bf16_tflops"], 4)
        stages_out[k] = ent

    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline:
        times, o = cpu_reference_frames(scene, weights, min_seconds=10.0, max_frames=3)
        cpu_baseline = {"value": o["n_rays"] * len(times) / sum(times), "unit": "rays/s",
                        "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{len(times)} full frame(s) of the same scene through oracle/gpnerf_oracle.py "
                                  f"({sum(times):.1f} s of CPU work)",
                        "s_per_frame": sum(times) / len(times)}
        # the benchmarked (tensor-core) frame against the oracle frame just rendered, over the mask_at_box pixels
        # only (libs/evaluators/if_nerf.py:49-57)
        if world == 1:
            import numpy as np
            m = o["mask_at_box"].reshape(-1).bool()
            a = ref_image.reshape(-1, 3)[m].double()
            b = o["pred_img"].reshape(-1, 3)[m].double()
            mse = float(((a - b) ** 2).mean()) if a.numel() else 0.0
            xor = np.setxor1d(ref_valid1.numpy(), o["valid1"].numpy())
            parity = {"against": "oracle/gpnerf_oracle.py (fp32, CPU) on the same scene and weights",
                      "pixels": "mask_at_box only", "n_px": int(m.sum()),
                      "psnr_mask": 10.0 * __import__("math").log10(1.0 / max(mse, 1e-20)),
                      "max_abs": float((a - b).abs().max()) if a.numel() else 0.0, "rms": mse ** 0.5,
                      "rays_equal": bool(counts["n_rays"] == o["n_rays"]), "P1_equal": bool(counts["P1"] == o["P1"]),
                      "hit_mask_equal": bool(torch.equal(ref_hit.bool(), o["mask_at_box"].reshape(-1).bool())),
                      "valid1_xor": int(len(xor)), "P2_oracle": int(o["P2"]),
                      "valid1_xor_max_abs_sigma": float(o["sigma"][torch.from_numpy(xor).long()].abs().max()) if len(xor) else 0.0}

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if (frames_mode or world == 1) else "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "frames_per_s": (world if frames_mode else 1) * 1e3 / ms_per_step,
        "pixel_rays_per_s": (world if frames_mode else 1) * n_px * 1e3 / ms_per_step,
        "config": {"workload": WORKLOAD, "rays": g_rays, "points": g_rays * S_SAMPLES, "P1": g_p1, "P2": g_p2,
                   "l2": "256 MB flush between timed steps; inputs 135 MB > 126 MB L2",
                   "sharding": ("single GPU" if world == 1 else
                                (f"one frame per GPU and step ({world} consecutive frames of an orbit sweep, {SWEEP_STEP_DEG}° apart), "
                                 "images gathered on rank 0"
                                 if frames_mode else
                                 f"one frame, pixel tiles of {args.tile_px} dealt diagonally over {world} ranks") +
                                ("; K5 writes the tiles into the peers' images over NVLink (CUDA-IPC peer memory, "
                                 "arrival flags; no NCCL on the data path)" if eng.exchange is not None else
                                 "; one NCCL all_gather per step")),
                   "peer_check": peer_check, "per_rank": per_rank,
                   "launch": "eager, one launch per kernel" if args.no_graph else
                             "one CUDA-graph replay per frame (frame constants through a pinned buffer)",
                   "stages_ms_from": "eager re-issue of the same steps with CUDA events around every stage",
                   "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "sparse_levels_device": sparse_dev, "strong": strong, "parity": parity,
        "stages_ms": stages_out, "cpu_baseline": cpu_baseline,
    }
    emit(line)


if __name__ == "__main__":
    main()
