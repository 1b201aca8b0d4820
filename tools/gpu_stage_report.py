"""Print the stage-by-stage GPU-vs-oracle report for a few scenes (debug aid;
run under gpurun).  Writes gpurun_out/stage_report.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402

from gpnerf_b200 import synth  # noqa: E402
import stages  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]
    for tag, H, S, seed in (("mini", 128, 16, 5), ("mini_s64", 96, 64, 11), ("half", 256, 64, 42)):
        if only and tag not in only:
            continue
        scene = synth.make_scene("zju", H=H, W=H, V=3, seed=seed)
        w = synth.make_head_weights(V=3, seed=seed + 100, random_bias=(tag != "mini"))
        try:
            rep, _, _ = stages.compare_progressive(scene, w, S)
        except Exception as e:  # keep going: the report is a debugging aid
            rep = {"error": repr(e)}
        out[tag] = rep
        print(tag, json.dumps(rep, indent=1), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stage_report.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
