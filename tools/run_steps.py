#!/usr/bin/env python
"""Run a few un-graphed ScenePipeline steps (for ncu / compute-sanitizer): every kernel of the
benchmark step back to back on one stream, per-kernel CUDA-event means printed as JSON.

    python tools/run_steps.py --steps 3 [--feature-dtype f32]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mvsdet_b200.hotpath import MVSDetHotPath  # noqa: E402
from mvsdet_b200.pipeline import ScenePipeline  # noqa: E402
from mvsdet_b200.scene import SceneConfig, make_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--views", type=int, default=20)
    ap.add_argument("--feature-dtype", default="bf16")
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
    def timed(n):
        timers = {}
        for i in range(n):
            pipes[i % nbuf].step(timers)
        torch.cuda.synchronize()
        return {k: statistics.mean(x.elapsed_time(y) for x, y in v[1:]) for k, v in timers.items()}

    res = timed(a.steps)
    print(json.dumps({k: round(v, 4) for k, v in res.items()}))


if __name__ == "__main__":
    main()
