"""Device prologue (csrc/scene_setup.cu through geometry.scene_geometry(prologue="device"))
against the host prologue -- the reference's own ATen ops on the CPU: neighbour ids,
homographies and voxel projections must be identical bit for bit on 120 seeded scenes
(ScanNet- and ARKit-shaped, 2..40 views), for whole scenes and for view slices."""
import numpy as np
import pytest
import torch

from mvsdet_b200 import geometry as G
from mvsdet_b200.scene import SceneConfig, make_cameras, tiny_config

pytestmark = pytest.mark.gpu

CONFIGS = [SceneConfig(n_views=20), SceneConfig(n_views=40, near_far_range=(0.5, 5.5), per_view_intrinsics=True),
           SceneConfig(n_views=3), tiny_config(n_views=2)]


def _meta(cfg, seed):
    w2c, intr = make_cameras(cfg, np.random.default_rng(seed))
    return dict(lidar2img=dict(extrinsic=w2c, intrinsic=intr, origin=np.asarray(cfg.origin, dtype=np.float32)),
                img_shape=tuple(cfg.img_shape), ori_shape=tuple(cfg.ori_shape), pad_shape=tuple(cfg.pad_shape))


def _kw(cfg):
    return dict(stride=cfg.stride, near_far_range=cfg.near_far_range, num_depth=cfg.num_depth,
                n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, device="cuda")


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"V{c.n_views}{'pv' if c.per_view_intrinsics else ''}")
def test_device_prologue_equals_host_prologue(cfg):
    for seed in range(30):
        meta = _meta(cfg, seed)
        host = G.scene_geometry(meta, prologue="host", **_kw(cfg))
        dev = G.scene_geometry(meta, prologue="device", **_kw(cfg))
        assert torch.equal(dev.neighbor_ids, host.neighbor_ids), f"neighbour ids (seed {seed})"
        assert torch.equal(dev.hom, host.hom), f"homographies (seed {seed})"
        assert torch.equal(dev.projection, host.projection), f"projection (seed {seed})"
        assert torch.equal(dev.k_feat, host.k_feat)
        assert torch.equal(dev.neighbor_ids_ref(), host.neighbor_ids_host)
        assert dev.points.data_ptr() == host.points.data_ptr()      # cached static parts


def test_device_prologue_view_slice():
    cfg = SceneConfig(n_views=12, per_view_intrinsics=True)
    meta = _meta(cfg, 2)
    full = G.scene_geometry(meta, **_kw(cfg))
    part = G.scene_geometry(meta, view_slice=slice(4, 8), **_kw(cfg))
    assert torch.equal(part.neighbor_ids, full.neighbor_ids[4:8])
    assert torch.equal(part.hom, full.hom[4:8])
    assert torch.equal(part.projection, full.projection[4:8])
    assert torch.equal(part.k_feat, full.k_feat[4:8])
    assert part.depth_values.shape == (4, cfg.num_depth)


def test_single_view_scene():
    cfg = tiny_config(n_views=1)
    meta = _meta(cfg, 0)
    geo = G.scene_geometry(meta, **_kw(cfg))
    host = G.scene_geometry(meta, prologue="host", **_kw(cfg))
    assert geo.k == 0 and tuple(geo.neighbor_ids.shape) == (1, 0) and tuple(geo.hom.shape) == (1, 0, 12)
    assert torch.equal(geo.projection, host.projection)
