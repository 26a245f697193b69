#!/usr/bin/env python
"""Per-kernel CUDA-event medians of the benchmark step for the library selected by MVSDET_B200_LIB
(experiment builds, tools/build_exp_lib.py), plus the graph-replayed step time and a checksum of
g_feature against the first run of the same gpurun call (kept under /tmp).

    MVSDET_B200_LIB=mvsdet_b200/lib/exp_g1mb4.so python tools/time_kernels.py [--feature-dtype bf16]
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
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--feature-dtype", default="bf16")
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("MVSDET_B200_LIB", "default")))
    ap.add_argument("--tuning5", type=int, default=0, help="test hook mvsd_set_tuning(5, x) for the backward kernel")
    a = ap.parse_args()
    cfg = SceneConfig(n_views=20)
    dev = torch.device("cuda")
    if a.tuning5:
        from mvsdet_b200 import _lib
        _lib.set_tuning(5, a.tuning5)
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk)
    pipes = []
    for b in range(2):
        scene = make_scene(cfg, seed=b)
        p = ScenePipeline(cfg, dev, feature_dtype=torch.bfloat16 if a.feature_dtype == "bf16" else torch.float32)
        p.set_geometry(mod.geometry(scene["img_meta"], dev))
        p.load_scene(scene)
        pipes.append(p)
    timers = {}
    for i in range(a.steps):
        pipes[i % 2].step(timers)
    torch.cuda.synchronize()
    ms = {k: round(statistics.median(x.elapsed_time(y) for x, y in v[2:]), 4) for k, v in timers.items()}
    graphs = [p.capture() for p in pipes]
    for i in range(6):
        graphs[i % 2].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        graphs[i % 2].replay()
    e1.record()
    torch.cuda.synchronize()
    out = {"lib": a.tag, "tuning5": a.tuning5, "feature_dtype": a.feature_dtype, "step_ms_graph": round(e0.elapsed_time(e1) / 100, 4),
           "kernels_ms": ms}
    # two scenes in flight: both pipelines' steps forked inside ONE graph (tails of one scene's
    # kernels are filled by the other's)
    s_b = torch.cuda.Stream()
    torch.cuda.synchronize()
    pair = torch.cuda.CUDAGraph()
    with torch.cuda.graph(pair):
        cur = torch.cuda.current_stream()
        s_b.wait_stream(cur)
        pipes[0].step()
        with torch.cuda.stream(s_b):
            pipes[1].step()
        cur.wait_stream(s_b)
    for i in range(4):
        pair.replay()
    torch.cuda.synchronize()
    e0.record()
    for i in range(50):
        pair.replay()
    e1.record()
    torch.cuda.synchronize()
    out["step_ms_two_in_flight"] = round(e0.elapsed_time(e1) / 100, 4)
    ref_path = f"/tmp/time_kernels_ref_{a.feature_dtype}.pt"     # lives for one gpurun call
    g = pipes[0].g_feature.double().cpu()
    if os.path.isfile(ref_path):
        ref = torch.load(ref_path)
        out["g_feature_max_err_vs_ref_over_rms"] = float((g - ref).abs().max() / ref.pow(2).mean().sqrt())
    else:
        torch.save(g, ref_path)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
