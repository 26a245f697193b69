"""SURVEY.md 8d "eager-GPU comparison": the reference's PyTorch path (the oracle, ATen
CUDA kernels, fp32) on the same B200 and the same scene as bench.py's headline workload,
timed with CUDA events.  Measurement helper only: nothing here is on the product path.

    python tools/eager_gpu.py [--steps 5] [--warmup 2] > gpurun_out/eager_gpu.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--views", type=int, default=20)
    args = ap.parse_args()
    from mvsdet_b200.scene import SceneConfig, make_scene
    from oracle import mvsdet_oracle as O
    cfg = SceneConfig(n_views=args.views)
    scene = make_scene(cfg, seed=0)
    dev = torch.device("cuda:0")
    feature0 = scene["feature"].to(dev)
    cost0 = scene["cost_out"].to(dev)
    g_var = scene["g_variance"].to(dev)
    g_vol = scene["g_volume_mean"].to(dev)

    def step(backward=True):
        feature = feature0.clone().requires_grad_(backward)
        cost_out = cost0.clone().requires_grad_(backward)
        res = O.hot_path(feature, scene["img_meta"], lambda var: cost_out,
                         near_far_range=cfg.near_far_range, num_depth=cfg.num_depth, topk=cfg.topk,
                         n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, stride=cfg.stride,
                         training=backward)
        if backward:
            torch.autograd.backward([res["variance"], res["volume_mean"]], [g_var, g_vol])
        return res

    out = {"what": "reference PyTorch path (oracle) on CUDA tensors, fp32, eager ATen kernels",
           "views": cfg.n_views, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}
    for name, bwd in (("fwd_bwd", True), ("fwd", False)):
        try:
            ctx = torch.enable_grad() if bwd else torch.no_grad()
            with ctx:
                for _ in range(args.warmup):
                    step(bwd)
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    step(bwd)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            out[name] = {"ms_per_scene": round(ms, 3), "scenes_per_s": round(1e3 / ms, 3),
                         "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2),
                         "steps": args.steps}
        except Exception as exc:          # keep the other leg's number
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
