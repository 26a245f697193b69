#!/usr/bin/env python
"""Run a few un-graphed ScenePipeline steps (for ncu / compute-sanitizer) and
optionally time kernel variants selected through mvsd_set_tuning.

    python tools/run_steps.py --steps 3
    python tools/run_steps.py --sweep          # time tuning variants with CUDA events
"""
import argparse
import itertools
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mvsdet_b200 import _lib  # noqa: E402
from mvsdet_b200.hotpath import MVSDetHotPath  # noqa: E402
from mvsdet_b200.pipeline import ScenePipeline  # noqa: E402
from mvsdet_b200.scene import SceneConfig, make_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--views", type=int, default=20)
    ap.add_argument("--feature-dtype", default="bf16")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--tuning", default="", help="key=value,key=value")
    a = ap.parse_args()
    cfg = SceneConfig(n_views=a.views)
    dev = torch.device("cuda")
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk)
    nbuf = 2
    pipes = []
    for b in range(nbuf):
        scene = make_scene(cfg, seed=b)
        p = ScenePipeline(cfg, dev, feature_dtype=torch.bfloat16 if a.feature_dtype == "bf16" else torch.float32)
        p.set_geometry(mod.geometry(scene["img_meta"], dev))
        p.load_scene(scene)
        pipes.append(p)
    for kv in filter(None, a.tuning.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(int(k), int(v))

    def timed(n):
        timers = {}
        for i in range(n):
            pipes[i % nbuf].step(timers)
        torch.cuda.synchronize()
        return {k: statistics.mean(x.elapsed_time(y) for x, y in v[1:]) for k, v in timers.items()}

    if not a.sweep:
        res = timed(a.steps)
        print(json.dumps({k: round(v, 4) for k, v in res.items()}))
        return
    timed(3)
    for ppw, pw in itertools.product((1, 2, 4), (2, 4, 8, 16)):
        if (8 * ppw) % pw:
            continue
        _lib.set_tuning(0, ppw)
        _lib.set_tuning(1, pw)
        _lib.set_tuning(2, min(ppw, 2))
        try:
            r = timed(8)
        except Exception as e:   # noqa: BLE001
            print("ppw", ppw, "pw", pw, "failed", e)
            continue
        print(f"ppw={ppw} pw={pw} ph={8 * ppw // pw}: fwd {r['plane_sweep_fwd']:.4f} ms, "
              f"bwd(ppw={min(ppw, 2)}) {r['plane_sweep_bwd']:.4f} ms", flush=True)


if __name__ == "__main__":
    main()
