"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from mvsdet_b200.scene import SceneConfig, make_scene
from oracle import mvsdet_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ("scannet_tiny", "arkit_tiny", "two_views", "wide_c")


def load_golden(name):
    """-> (scene dict rebuilt from the stored inputs, dict of reference outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfgd = json.loads(str(z["cfg_json"]))
    for key in ("near_far_range", "n_voxels", "voxel_size", "img_shape", "pad_shape",
                "ori_shape", "origin"):
        cfgd[key] = tuple(cfgd[key])
    cfg = SceneConfig(**cfgd)
    intr = z["in_intrinsic"]
    scene = dict(
        cfg=cfg, seed=int(z["seed"]),
        feature=torch.from_numpy(z["in_feature"]),
        cost_out=torch.from_numpy(z["in_cost_out"]),
        g_volume_mean=torch.from_numpy(z["in_g_volume_mean"]),
        g_variance=torch.from_numpy(z["in_g_variance"]),
        img_meta=dict(
            lidar2img=dict(extrinsic=[m for m in z["in_w2c"]],
                           intrinsic=[m for m in intr] if intr.ndim == 3 else intr,
                           origin=np.asarray(cfg.origin, dtype=np.float32)),
            img_shape=tuple(cfg.img_shape), ori_shape=tuple(cfg.ori_shape),
            pad_shape=tuple(cfg.pad_shape)),
    )
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return scene, outs


def oracle_chain(scene, training=True, with_grads=True):
    """Run oracle.hot_path on a scene with the scene's cost_out standing in for
    the cost-regularisation net; optionally the autograd gradients."""
    cfg = scene["cfg"]
    feature = scene["feature"].clone().requires_grad_(with_grads)
    cost_out = scene["cost_out"].clone().requires_grad_(with_grads)
    res = O.hot_path(feature, scene["img_meta"], lambda var: cost_out,
                     near_far_range=cfg.near_far_range, num_depth=cfg.num_depth,
                     topk=cfg.topk, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                     stride=cfg.stride, training=training)
    if with_grads:
        g1, = torch.autograd.grad(res["variance"], feature, scene["g_variance"],
                                  retain_graph=True)
        g2, g3 = torch.autograd.grad(res["volume_mean"], (feature, cost_out),
                                     scene["g_volume_mean"], allow_unused=True)
        res["g_feature_from_variance"] = g1
        res["g_feature_from_voxels"] = g2
        res["g_cost_out"] = g3
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in res.items()}


def assert_close(a, b, rtol, atol, what=""):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), (
        f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance, "
        f"max abs err {float(err.max()):.3e}, max |ref| {float(b.abs().max()):.3e}")
