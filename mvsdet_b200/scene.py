"""Seeded synthetic ScanNet-/ARKit-shaped scenes for parity tests and bench.py.

The reference consumes, per scene, FPN features ``[V,C,Hf,Wf]`` plus an
``img_meta`` dict built by its dataset class
(projects/NeRF-Det/nerfdet/scannet_multiview_dataset.py:100-168 ScanNet,
:263-340 ARKit): ``lidar2img = {extrinsic: [w2c 4x4]*V, intrinsic: 4x4 (or a
list of V 4x4 for ARKit), origin}``, ``img_shape`` (un-padded, 239x320) and
``ori_shape`` (968x1296).  There is no dataset or network on the build and
GPU boxes, so this module fabricates the same structures:
cameras on a jittered helix around the voxel-grid centre looking inward so
that pose-space neighbours overlap, ScanNet colour intrinsics, N(0,1)
features, and a peaky cost-regularisation output.

Everything is generated with numpy's PCG64 on the host from the seed, so the
same seed gives the same scene on every machine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass
class SceneConfig:
    """Hot-path hyper-parameters the drop-in must honour:
    mvsdet_res50_2x_low_res.py:10,17,40,101 and mvsdet_arkit.py:15."""
    n_views: int = 20
    channels: int = 256
    num_depth: int = 12                       # gs_cfg.num_monocular_samples
    topk: int = 3
    near_far_range: Tuple[float, float] = (0.2, 5.0)
    n_voxels: Tuple[int, int, int] = (40, 40, 16)
    voxel_size: Tuple[float, float, float] = (0.16, 0.16, 0.2)
    img_shape: Tuple[int, int] = (239, 320)   # after resize, before padding
    pad_shape: Tuple[int, int] = (240, 320)
    ori_shape: Tuple[int, int] = (968, 1296)
    stride: int = 4
    origin: Tuple[float, float, float] = (0.0, 0.0, 0.5)
    per_view_intrinsics: bool = False         # ARKit layout
    num_neighbors: int = 2                    # k, hard-coded at mvsdet.py:432

    @property
    def feat_hw(self) -> Tuple[int, int]:
        return self.pad_shape[0] // self.stride, self.pad_shape[1] // self.stride

    @property
    def crop_hw(self) -> Tuple[int, int]:
        return self.img_shape[0] // self.stride, self.img_shape[1] // self.stride

    @property
    def depth_interval(self) -> float:
        return (self.near_far_range[1] - self.near_far_range[0]) / self.num_depth

    @property
    def ratio(self) -> float:
        return self.ori_shape[0] / (self.img_shape[0] / self.stride)


SCANNET = SceneConfig()
ARKIT = SceneConfig(n_views=40, near_far_range=(0.5, 5.5), per_view_intrinsics=True,
                    ori_shape=(968, 1296))


def tiny_config(**kw) -> SceneConfig:
    """A small scene the CPU oracle finishes in well under a second."""
    base = dict(n_views=5, channels=32, num_depth=8, topk=3,
                n_voxels=(12, 12, 6), voxel_size=(0.4, 0.4, 0.4),
                img_shape=(47, 64), pad_shape=(48, 64), ori_shape=(188, 256),
                stride=4)
    base.update(kw)
    return SceneConfig(**base)


def _look_at(eye: np.ndarray, target: np.ndarray, roll: float) -> np.ndarray:
    """camera-to-world, OpenCV convention (x right, y down, z forward),
    world z up."""
    fwd = target - eye
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    cr, sr = math.cos(roll), math.sin(roll)
    right, down = cr * right + sr * down, -sr * right + cr * down
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    return c2w


def make_cameras(cfg: SceneConfig, rng: np.random.Generator):
    """-> (w2c list of float32 [4,4], intrinsic float32 [4,4] or list of them)."""
    v = cfg.n_views
    centre = np.asarray(cfg.origin, dtype=np.float64)
    w2c = []
    turns = max(1.0, v / 24.0)
    for i in range(v):
        ang = 2 * math.pi * turns * i / v + rng.normal(0, 0.03)
        rad = 1.6 + 0.25 * math.sin(3.1 * ang) + rng.normal(0, 0.05)
        eye = centre + np.array([rad * math.cos(ang), rad * math.sin(ang),
                                 0.55 + 0.25 * math.sin(1.7 * ang) + rng.normal(0, 0.03)])
        tgt = centre + np.array([rng.normal(0, 0.25), rng.normal(0, 0.25),
                                 -0.25 + rng.normal(0, 0.1)])
        c2w = _look_at(eye, tgt, rng.normal(0, 0.04))
        w2c.append(np.linalg.inv(c2w).astype(np.float32))
    # ScanNet colour camera at 1296x968 (scaled if ori_shape differs)
    sy = cfg.ori_shape[0] / 968.0
    sx = cfg.ori_shape[1] / 1296.0

    def _k(jit):
        k = np.eye(4, dtype=np.float32)
        k[0, 0] = 1170.19 * sx * (1 + jit[0])
        k[1, 1] = 1170.19 * sy * (1 + jit[1])
        k[0, 2] = 647.75 * sx + jit[2]
        k[1, 2] = 483.75 * sy + jit[3]
        return k
    if cfg.per_view_intrinsics:
        intr = [_k(rng.normal(0, [0.01, 0.01, 4.0, 4.0])) for _ in range(v)]
    else:
        intr = _k(np.zeros(4))
    return w2c, intr


def make_scene(cfg: SceneConfig = SCANNET, seed: int = 0, *, with_grads: bool = True,
               cost_scale: float = 3.0) -> Dict:
    """Host-side (CPU, fp32) scene.

    Returns a dict with
      feature   [V,C,Hf,Wf]      N(0,1); rows >= crop h are "padding" but still
                                 random, the reference sweeps the full map
                                 (mvsdet.py:437)
      img_meta  the reference's per-scene meta dict (numpy matrices)
      cost_out  [V,2,D,Hf,Wf]    stand-in for CostRegNet_3DGS's output
      g_volume_mean [C,nx,ny,nz] upstream gradient for the backward
      g_variance    [V,C,D,Hf,Wf] upstream gradient reaching the variance
                                 volume (what CostRegNet's backward would give)
    """
    rng = np.random.default_rng(seed)
    hf, wf = cfg.feat_hw
    v, c, d = cfg.n_views, cfg.channels, cfg.num_depth
    w2c, intr = make_cameras(cfg, rng)
    trng = torch.Generator().manual_seed(int(seed) * 7919 + 13)
    feature = torch.randn(v, c, hf, wf, generator=trng)
    cost_out = torch.randn(v, 2, d, hf, wf, generator=trng) * cost_scale
    scene = dict(
        cfg=cfg, seed=seed, feature=feature, cost_out=cost_out,
        img_meta=dict(
            lidar2img=dict(extrinsic=w2c, intrinsic=intr,
                           origin=np.asarray(cfg.origin, dtype=np.float32)),
            img_shape=tuple(cfg.img_shape), ori_shape=tuple(cfg.ori_shape),
            pad_shape=tuple(cfg.pad_shape)),
    )
    if with_grads:
        scene["g_volume_mean"] = torch.randn(c, *cfg.n_voxels, generator=trng)
        scene["g_variance"] = torch.randn(v, c, d, hf, wf, generator=trng)
    return scene
