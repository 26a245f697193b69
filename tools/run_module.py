#!/usr/bin/env python
"""Wall-clock of the autograd drop-in (MVSDetHotPath forward + backward through
torch.autograd) per scene, with and without a cached SceneGeometry -- the cost a
maintainer sees after applying INTEGRATION.md level 1, host work included."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mvsdet_b200.hotpath import MVSDetHotPath
from mvsdet_b200.scene import SceneConfig, make_scene

cfg = SceneConfig(n_views=20)
scene = make_scene(cfg, seed=0)
dev = torch.device("cuda")
hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                    feature_dtype=torch.bfloat16)
feat = scene["feature"].to(dev).requires_grad_(True)
cost = scene["cost_out"].to(dev).requires_grad_(True)
# the gradient arrives in the variance's own memory format (channels_last_3d), as cuDNN's
# Conv3d backward returns it for a channels_last_3d input
gvar = scene["g_variance"].to(dev).contiguous(memory_format=torch.channels_last_3d)
gvol = scene["g_volume_mean"].to(dev)

def step(geo=None):
    res = hot(feat, scene["img_meta"], cost_regularization=lambda var: cost, geometry=geo)
    torch.autograd.backward([res["variance"], res["volume_mean"]], [gvar, gvol])
    feat.grad = None; cost.grad = None

out = {}
for name, geo in (("geometry_every_call", None), ("geometry_cached", hot.geometry(scene["img_meta"], dev))):
    for _ in range(5):
        step(geo)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        step(geo)
    torch.cuda.synchronize()
    out[name + "_ms"] = round((time.perf_counter() - t0) / n * 1e3, 3)
t0 = time.perf_counter()
for _ in range(20):
    hot.geometry(scene["img_meta"], dev)
torch.cuda.synchronize()
out["geometry_host_ms"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
print(json.dumps(out))

# where the host time goes: torch profiler table of one step (top entries)
if os.environ.get("MVSD_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step(None)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))

# host-side hot spots of the drop-in: cProfile over 200 steps (cumulative, top entries)
if os.environ.get("MVSD_CPROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        step(None)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(25)
