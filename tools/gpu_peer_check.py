"""Multi-GPU check of the peer-memory image exchange (run under torchrun, N >= 2):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/gpu_peer_check.py

tiles : one frame sharded by pixel tiles; every rank must end up with the image
        a single GPU renders, bit for bit, through K5's peer stores alone.
frames: rank r renders novel view r; rank 0's slot r must equal what rank r
        holds locally.
Prints one JSON line per mode on rank 0; exits non-zero on mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpnerf_b200  # noqa: F401,E402
from gpnerf_b200 import synth  # noqa: E402
from gpnerf_b200._lib import PREC_BF16  # noqa: E402
from gpnerf_b200.engine import Engine  # noqa: E402
from gpnerf_b200.peer import PeerExchange  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    H, S, V, tile = 256, 64, 3, 64
    scene = synth.make_scene("zju", H=H, W=H, V=V, seed=7)
    w = synth.make_head_weights(V=V, seed=107)
    lv = [t.to(dev) for t in scene["levels"]]
    fm, im = scene["featmaps"].to(dev), scene["src_imgs"].to(dev)
    ok_all = True

    def engine(r, wd, mode, graph):
        e = Engine(H, H, S, V, device=dev, precision=PREC_BF16, rank=r, world=wd, tile_px=tile)
        e.set_weights(w)
        if mode is not None:
            e.attach_exchange(PeerExchange(H, H, dev, rank, world, mode=mode))
        e.set_static_inputs(lv, fm, im)
        e.upload_products(lv, fm, im)
        return e

    # single-GPU reference of the frame (and of this rank's own view for the frames mode)
    ref = engine(0, 1, None, False)
    ref.render_progressive(ref.make_frame(scene))
    torch.cuda.synchronize()
    full = ref.pred_img.view(-1, 3).clone()
    full_hit = ref.hit_mask.clone()
    for graph in (False, True):
        e = engine(rank, world, "tiles", graph)
        fr = e.make_frame(scene)
        for it in range(4):
            if graph:
                e.run_progressive_graphed(fr)
            else:
                e.render_progressive(fr)
            torch.cuda.synchronize()
            ok = torch.equal(e.result_image(), full) and torch.equal(e.result_hit_mask(), full_hit)
            t = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok_all &= bool(t.item())
        if rank == 0:
            print(json.dumps({"mode": "tiles", "graph": graph, "world": world, "bit_exact_on_all_ranks": bool(t.item())}),
                  flush=True)
        dist.barrier()
        e.exchange.close()
    # frames: every rank its own view
    mine = synth.retarget(scene, 45.0 + 360.0 * rank / world)
    ref.render_progressive(ref.make_frame(mine))
    torch.cuda.synchronize()
    own = ref.pred_img.view(-1, 3).clone()
    e = engine(0, 1, "frames", True)
    fr = e.make_frame(mine)
    for it in range(3):
        e.run_progressive_graphed(fr)
        torch.cuda.synchronize()
    gathered = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(gathered, own)
    ok = torch.equal(e.result_image(0), own)
    if rank == 0:
        ok &= all(torch.equal(e.exchange.image(slot=r), gathered[r]) for r in range(world))
    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_all &= bool(t.item())
    if rank == 0:
        print(json.dumps({"mode": "frames", "world": world, "root_slots_match_ranks": bool(t.item())}), flush=True)
    dist.barrier()
    e.exchange.close()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
