#!/usr/bin/env python
"""Compare plane-sweep backward variants (mvsd_set_tuning key 5) on the full-size
benchmark scene: g_feature of every variant against the default kernel, and
per-kernel CUDA-event times.

    python tools/compare_variants.py --variants 0,5,6 [--feature-dtype bf16|f32] [--views 20]
"""
import argparse
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
    ap.add_argument("--variants", default="0,5,6")
    ap.add_argument("--key", type=int, default=5)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--views", type=int, default=20)
    ap.add_argument("--feature-dtype", default="bf16")
    a = ap.parse_args()
    cfg = SceneConfig(n_views=a.views)
    dev = torch.device("cuda")
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk)
    pipes = []
    for b in range(2):
        scene = make_scene(cfg, seed=b)
        p = ScenePipeline(cfg, dev, feature_dtype=torch.bfloat16 if a.feature_dtype == "bf16" else torch.float32)
        p.set_geometry(mod.geometry(scene["img_meta"], dev))
        p.load_scene(scene)
        pipes.append(p)
    base = None
    for var in [int(x) for x in a.variants.split(",")]:
        _lib.set_tuning(a.key, var)
        timers = {}
        for i in range(a.steps):
            pipes[i % 2].step(timers)
        torch.cuda.synchronize()
        ms = {k: statistics.median(x.elapsed_time(y) for x, y in v[2:]) for k, v in timers.items()}
        pipes[0].step()
        torch.cuda.synchronize()
        g = pipes[0].g_feature.double()
        line = {"variant": var, "plane_sweep_bwd_ms": round(ms["plane_sweep_bwd"], 4),
                "plane_sweep_fwd_ms": round(ms["plane_sweep_fwd"], 4),
                "step_ms": round(sum(ms.values()), 4)}
        if base is None:
            base = g
        else:
            rms = float(base.pow(2).mean().sqrt())
            line["max_abs_err_vs_first"] = float((g - base).abs().max())
            line["rel_to_rms"] = line["max_abs_err_vs_first"] / rms
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
