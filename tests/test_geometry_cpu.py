"""Host geometry (mvsdet_b200/geometry.py) against the oracle / the loop forms it
replaces: bit-exact, on CPU."""
import numpy as np
import pytest
import torch

from mvsdet_b200 import geometry as G
from mvsdet_b200.scene import SceneConfig, make_cameras, tiny_config
from oracle import mvsdet_oracle as O

CONFIGS = [SceneConfig(n_views=20), SceneConfig(n_views=40, near_far_range=(0.5, 5.5), per_view_intrinsics=True),
           SceneConfig(n_views=3), tiny_config(n_views=2)]


def _meta(cfg, seed):
    """img_meta of scene.make_scene without generating the feature maps."""
    w2c, intr = make_cameras(cfg, np.random.default_rng(seed))
    return dict(lidar2img=dict(extrinsic=w2c, intrinsic=intr, origin=np.asarray(cfg.origin, dtype=np.float32)),
                img_shape=tuple(cfg.img_shape), ori_shape=tuple(cfg.ori_shape), pad_shape=tuple(cfg.pad_shape))


def _meta_tensors(meta):
    intr = torch.as_tensor(np.array(meta["lidar2img"]["intrinsic"]), dtype=torch.float32)
    extr = torch.as_tensor(np.array(meta["lidar2img"]["extrinsic"]), dtype=torch.float32)
    return intr, extr


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"V{c.n_views}{'pv' if c.per_view_intrinsics else ''}")
def test_vectorized_projection_equals_reference_loop(cfg):
    """projection_vectorized spells out the FMA chain of ATen's per-view mm
    (mvsdet.py:1124-1156): identical bits on every seeded scene."""
    for seed in range(30):
        meta = _meta(cfg, seed)
        intr, extr = _meta_tensors(meta)
        ratio = meta["ori_shape"][0] / (meta["img_shape"][0] / cfg.stride)
        fast = G.projection_vectorized(intr, extr, ratio)
        assert torch.equal(fast, G.compute_projection(meta, cfg.stride))
        assert torch.equal(fast, O.compute_projection(intr, extr, ratio))


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"V{c.n_views}{'pv' if c.per_view_intrinsics else ''}")
def test_scene_geometry_matches_oracle_pieces(cfg):
    meta = _meta(cfg, 5)
    geo = G.scene_geometry(meta, stride=cfg.stride, near_far_range=cfg.near_far_range,
                           num_depth=cfg.num_depth, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                           device="cpu")
    intr, w2c = _meta_tensors(meta)
    ratio = meta["ori_shape"][0] / (meta["img_shape"][0] / cfg.stride)
    k = min(2, cfg.n_views - 1)
    nbr = O.get_nearest_pose_ids(torch.inverse(w2c), k)
    assert np.array_equal(geo.neighbor_ids.numpy(), nbr.numpy().astype(np.int32))
    kf = O.feature_intrinsics(intr, ratio)
    ref_proj, nei = O.collect_proj(w2c, kf, nbr)
    for j in range(k):
        rot, trans = O.homography(nei[j], ref_proj)
        assert torch.equal(geo.hom[:, j, :9].reshape(-1, 3, 3), rot)
        assert torch.equal(geo.hom[:, j, 9:], trans.reshape(-1, 3))
    assert torch.equal(geo.points, O.get_points(cfg.n_voxels, cfg.voxel_size, meta["lidar2img"]["origin"]))
    dv = torch.as_tensor(O.depth_values_for(cfg.near_far_range, cfg.num_depth))
    assert torch.equal(geo.depth_values, dv.unsqueeze(0).repeat(cfg.n_views, 1))
    # the static parts are cached per configuration, not rebuilt per scene
    geo2 = G.scene_geometry(meta, stride=cfg.stride, near_far_range=cfg.near_far_range,
                            num_depth=cfg.num_depth, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                            device="cpu")
    assert geo2.points.data_ptr() == geo.points.data_ptr()


def test_view_slice_keeps_full_scene_neighbours():
    cfg = SceneConfig(n_views=12)
    meta = _meta(cfg, 2)
    kw = dict(stride=cfg.stride, near_far_range=cfg.near_far_range, num_depth=cfg.num_depth,
              n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, device="cpu")
    full = G.scene_geometry(meta, **kw)
    part = G.scene_geometry(meta, view_slice=slice(4, 8), **kw)
    assert torch.equal(part.neighbor_ids, full.neighbor_ids[4:8])
    assert torch.equal(part.hom, full.hom[4:8])
    assert torch.equal(part.projection, full.projection[4:8])
    assert part.depth_values.shape == (4, cfg.num_depth)


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"V{c.n_views}{'pv' if c.per_view_intrinsics else ''}")
def test_host_camera_block_reproduces_aten_bits(cfg):
    """The host half of the device prologue (geometry.host_camera_block): numpy fp32 arithmetic
    that must be ATen's, bit for bit -- K_feat (rows 0-1 / ratio), ref_proj = K_feat @ w2c (the
    un-fused chain of ATen's batched small-matrix product) and its inverse -- on 30 seeded
    scenes per configuration (120 scenes)."""
    for seed in range(30):
        meta = _meta(cfg, seed)
        w2c_np, k_feat, ref_proj, inv_ref = G.host_camera_block(meta, cfg.stride)
        intr, w2c = _meta_tensors(meta)
        ratio = meta["ori_shape"][0] / (meta["img_shape"][0] / cfg.stride)
        kf = O.feature_intrinsics(intr, ratio)
        assert np.array_equal(k_feat, kf.numpy())
        nbr = torch.zeros(cfg.n_views, 1, dtype=torch.long)
        want, _ = O.collect_proj(w2c, kf, nbr)
        assert np.array_equal(ref_proj, want.numpy()), f"ref_proj differs from ATen's matmul (seed {seed})"
        assert np.array_equal(inv_ref, torch.inverse(want).numpy())
        assert np.array_equal(w2c_np, w2c.numpy())


def test_setup_kernel_arithmetic_restated_on_host():
    """What csrc/scene_setup.cu computes, restated with numpy fp32/fp64 ops: the un-fused
    4-term chain for nei_proj @ inverse(ref_proj) equals ATen's batched matmul bit for bit, and
    the fp64 closed-form camera centres order the neighbours exactly as the reference's fp32
    knn does on these scenes (the kernel itself is checked on the GPU, tests/test_gpu_geometry.py)."""
    for cfg in CONFIGS[:2]:
        for seed in range(10):
            meta = _meta(cfg, seed)
            w2c_np, k_feat, ref_proj, inv_ref = G.host_camera_block(meta, cfg.stride)
            intr, w2c = _meta_tensors(meta)
            k = min(2, cfg.n_views - 1)
            nbr = O.get_nearest_pose_ids(torch.inverse(w2c), k)
            # neighbour ids from fp64 centres
            loc = np.linalg.inv(w2c_np.astype(np.float64))[:, :3, 3].astype(np.float32).astype(np.float64)
            d2 = ((loc[:, None] - loc[None]) ** 2).sum(-1)
            np.fill_diagonal(d2, np.inf)
            mine = np.argsort(d2, axis=1, kind="stable")[:, :k]
            assert np.array_equal(mine, nbr.numpy())
            # homography chain
            a = ref_proj[nbr.numpy()]                      # [V,k,4,4]
            b = inv_ref[:, None]                           # [V,1,4,4]
            acc = a[..., :, 0:1] * b[..., 0:1, :]
            for kk in (1, 2, 3):
                acc = acc + a[..., :, kk:kk + 1] * b[..., kk:kk + 1, :]
            want = torch.matmul(torch.from_numpy(ref_proj)[nbr], torch.from_numpy(inv_ref).unsqueeze(1))
            assert np.array_equal(acc, want.numpy())
