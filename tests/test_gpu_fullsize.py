"""Full-size (BASELINE.json configs[1]: V=20, 256 ch, 60x80, D=12, 40x40x16)
GPU checks.  The CPU oracle needs minutes for a whole scene at this size, so the
checks are (a) the oracle on a 2-reference-view subset of the same scene and
(b) size-independent properties: the variance is quadratic in the features, so
<g, var(f+u) - var(f-u)>/2 == <bwd(g), u> exactly up to rounding; the warp is
linear, so <warp(x), g> == <x, warp_bwd(g)>; top-k outputs are sorted, distinct
and consistent with the probability volume; view-sharded partial sums + one
reduction reproduce the whole-scene voxel volume bit-exactly in the counts.
"""
import numpy as np
import pytest
import torch

from helpers import oracle_chain
from mvsdet_b200.scene import SceneConfig, make_scene, tiny_config
from oracle import mvsdet_oracle as O

pytestmark = pytest.mark.gpu


def _close(a, b, what, rtol=1e-4, atol_scale=1e-4):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    rms = float(b.pow(2).mean().sqrt())
    err = (a - b).abs()
    bad = err > atol_scale * rms + rtol * b.abs()
    assert not bool(bad.any()), f"{what}: {int(bad.sum())}/{bad.numel()} off, max {float(err.max()):.3e} (rms {rms:.3e})"


@pytest.fixture(scope="module")
def full_scene():
    cfg = SceneConfig(n_views=20)
    return make_scene(cfg, seed=3)


@pytest.fixture(scope="module")
def hot(full_scene):
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = full_scene["cfg"]
    return MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                         stride=cfg.stride)


def test_full_size_subset_against_oracle(full_scene, hot):
    """Reference views 0,1 of the 20-view scene: variance, hypotheses and the
    per-view voxel masks against the oracle run on exactly those views."""
    from mvsdet_b200 import ops
    scene, cfg = full_scene, full_scene["cfg"]
    vs = 2
    cost = scene["cost_out"][:vs]
    ref = O.hot_path(scene["feature"], scene["img_meta"], lambda var: cost,
                     near_far_range=cfg.near_far_range, num_depth=cfg.num_depth, topk=cfg.topk,
                     n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, stride=cfg.stride,
                     training=False, view_subset=vs)
    dev = torch.device("cuda")
    feat_cl = ops.pack_features(scene["feature"].to(dev), torch.float32)
    geo = hot.geometry(scene["img_meta"], dev, view_slice=slice(0, vs))
    var = hot.variance(feat_cl, geo, ref_begin=0)
    _close(var, ref["variance"], "variance[0:2]")
    prob, off, est_depth, est_dens, est_idx, coding = hot.hypotheses(cost.to(dev))
    assert np.array_equal(est_idx.cpu().numpy(), ref["est_idx"].numpy())
    _close(est_depth, ref["est_depth"], "est_depth")
    _close(est_dens, ref["est_densities"], "est_densities")
    vol, count = ops.backproject_aggregate(feat_cl[:vs], geo.points, geo.projection, est_depth,
                                           est_dens, cfg.voxel_size[2], geo.height, geo.width)
    assert np.array_equal(count.cpu().numpy(), ref["count"].reshape(-1).numpy().astype(np.int32))
    _close(vol.reshape(-1), ref["volume_mean"].reshape(-1), "volume_mean (2 views)")


@pytest.mark.parametrize("feat_dtype", [torch.float32, torch.bfloat16])
def test_variance_backward_is_the_exact_adjoint(full_scene, hot, feat_dtype):
    from mvsdet_b200 import ops
    scene = full_scene
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(17)
    geo = hot.geometry(scene["img_meta"], dev)
    # multiples of 1/8 below 16 in magnitude: f, u, f+u and f-u are all exactly
    # representable in bf16 (8 significant bits), so the identity is exact in both modes
    f = (scene["feature"].to(dev) * 8).round().clamp(-120, 120) / 8
    u = (torch.randn(f.shape, generator=g).to(dev) * 8).round().clamp(-120, 120) / 8
    gv = scene["g_variance"].to(dev)

    def var_of(x):
        cl = ops.pack_features(x, feat_dtype)
        return ops.plane_sweep_variance(cl, geo.neighbor_ids, geo.hom, geo.depth_values)

    fp, fm = f + u, f - u
    if feat_dtype == torch.bfloat16:
        for t in (f, fp, fm):
            assert torch.equal(t.to(torch.bfloat16).float(), t)
    lhs = float(((var_of(fp).double() - var_of(fm).double()) * gv.double()).sum() / 2)
    fr = f.clone().requires_grad_(True)
    cl = ops.pack_features(fr, feat_dtype)
    var = ops.plane_sweep_variance(cl, geo.neighbor_ids, geo.hom, geo.depth_values)
    gf, = torch.autograd.grad(var, fr, gv)
    rhs = float((gf.double() * u.double()).sum())
    scale = float((gf.double().abs() * u.double().abs()).sum())
    tol = 2e-5
    assert abs(lhs - rhs) <= tol * scale, f"adjoint identity off: {lhs} vs {rhs} (scale {scale})"


@pytest.mark.parametrize("feat_dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_backward_linear_in_upstream_gradient(full_scene, hot, feat_dtype, tol):
    """bf16 mode hands the gradient back rounded to bf16 (the 1e-2 bar of the
    north star), so linearity holds to that rounding only."""
    from mvsdet_b200 import ops
    scene = full_scene
    dev = torch.device("cuda")
    geo = hot.geometry(scene["img_meta"], dev)
    f = scene["feature"].to(dev).requires_grad_(True)
    cl = ops.pack_features(f, feat_dtype)
    var = ops.plane_sweep_variance(cl, geo.neighbor_ids, geo.hom, geo.depth_values)
    g1 = scene["g_variance"].to(dev)
    g2 = torch.randn(g1.shape, generator=torch.Generator().manual_seed(5)).to(dev)
    a, = torch.autograd.grad(var, f, g1, retain_graph=True)
    b, = torch.autograd.grad(var, f, g2, retain_graph=True)
    c, = torch.autograd.grad(var, f, 0.5 * g1 - 2.0 * g2)
    _close(c, 0.5 * a - 2.0 * b, "bwd(0.5 g1 - 2 g2)", rtol=tol, atol_scale=tol)


def test_warp_adjoint_full_size(full_scene, hot):
    from mvsdet_b200 import ops
    scene = full_scene
    dev = torch.device("cuda")
    geo = hot.geometry(scene["img_meta"], dev)
    x = scene["feature"].to(dev)[:6]
    cl = ops.pack_features(x, torch.float32).requires_grad_(True)
    hom = geo.hom[:6, 0].contiguous()
    out = ops.homo_warp(cl, hom, geo.depth_values[:6])
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(2)).to(dev)
    gx, = torch.autograd.grad(out, cl, g)
    lhs = float((out.double() * g.double()).sum())
    rhs = float((gx.double() * cl.double()).sum())
    scale = float((out.double().abs() * g.double().abs()).sum())
    assert abs(lhs - rhs) <= 1e-5 * scale


def test_topk_properties_full_size(full_scene, hot):
    scene, cfg = full_scene, full_scene["cfg"]
    prob, off, est_depth, est_dens, est_idx, coding = hot.hypotheses(scene["cost_out"].cuda())
    assert float((prob.sum(1) - 1).abs().max()) < 1e-5
    assert bool((est_dens[:, :-1] >= est_dens[:, 1:]).all()), "top-k not descending"
    srt = est_idx.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "top-k indices not distinct"
    assert torch.equal(prob.gather(1, est_idx), est_dens)
    want_idx = prob.topk(cfg.topk, dim=1).indices          # torch's own top-k on the same probabilities
    same = (want_idx == est_idx).float().mean()
    assert float(same) > 0.9999                            # only exact ties may differ
    near, far = cfg.near_far_range
    assert float(est_depth.min()) >= near and float(est_depth.max()) <= far + cfg.depth_interval
    assert float(coding.min()) >= near and float(coding.max()) <= far + cfg.depth_interval


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_view_sharded_partials_match_whole_scene(full_scene, hot, shards):
    """What each rank of mvsdet_b200.sharded computes, run back to back on one
    GPU: per-shard sweep (ref_begin) -> top-k -> BP_SUM partials; their sum,
    normalised once, equals the whole-scene result (counts bit-exact)."""
    from mvsdet_b200 import ops, sharded
    scene, cfg = full_scene, full_scene["cfg"]
    dev = torch.device("cuda")
    feat = scene["feature"].to(dev)
    cost = scene["cost_out"].to(dev)
    whole = hot(feat, scene["img_meta"], cost_regularization=lambda var: cost)
    feat_cl = ops.pack_features(feat, torch.float32)
    c = feat_cl.shape[1]
    n = int(np.prod(cfg.n_voxels))
    total = None
    for r in range(shards):
        b, e = sharded.partition_views(cfg.n_views, shards, r)
        geo = hot.geometry(scene["img_meta"], dev, view_slice=slice(b, e))
        var = hot.variance(feat_cl, geo, ref_begin=b)
        assert torch.equal(var, whole["variance"][b:e]), "sharded sweep must be bit-identical"
        _, _, est_depth, est_dens, _, _ = hot.hypotheses(cost[b:e])
        vol, cnt = ops.backproject_aggregate(feat_cl[b:e], geo.points, geo.projection, est_depth,
                                             est_dens, cfg.voxel_size[2], geo.height, geo.width,
                                             mode="sum")
        buf = sharded.pack_partials(vol, cnt)
        total = buf if total is None else total + buf
    vol_sum, count = sharded.unpack_partials(total, c, n)
    mean = ops.voxel_normalize(vol_sum.contiguous(), count)
    assert torch.equal(count, whole["count"])
    _close(mean.reshape(-1), whole["volume_mean"].reshape(-1), "sharded volume_mean", rtol=1e-5, atol_scale=1e-6)


@pytest.mark.parametrize("world", [1, 4])
def test_sharded_forward_class_with_halo_packing(full_scene, hot, world):
    """ShardedSceneForward.local_partials played for every rank on one GPU (block +
    halo packing, re-based neighbour ids) against the whole-scene forward."""
    from mvsdet_b200 import ops, sharded
    scene, cfg = full_scene, full_scene["cfg"]
    dev = torch.device("cuda")
    feat = scene["feature"].to(dev)
    cost = scene["cost_out"].to(dev)
    whole = hot(feat, scene["img_meta"], cost_regularization=lambda var: cost)
    runner = sharded.ShardedSceneForward(hot)
    c, n = cfg.channels, int(np.prod(cfg.n_voxels))
    total = None
    for rank in range(world):
        b, e = sharded.partition_views(cfg.n_views, world, rank)
        vol, cnt, rng = runner.local_partials(feat, scene["img_meta"], lambda var: cost[b:e],
                                              rank_world=(rank, world))
        assert rng == (b, e)
        buf = sharded.pack_partials(vol, cnt)
        total = buf if total is None else total + buf
    vol_sum, count = sharded.unpack_partials(total, c, n)
    mean = ops.voxel_normalize(vol_sum.contiguous(), count)
    assert torch.equal(count, whole["count"])
    _close(mean.reshape(-1), whole["volume_mean"].reshape(-1), "sharded class volume_mean", rtol=1e-5, atol_scale=1e-6)
    if world == 1:
        out = runner(feat, scene["img_meta"], cost_regularization=lambda var: cost)
        assert torch.equal(out["count"], whole["count"])
        _close(out["volume_mean"].reshape(-1), whole["volume_mean"].reshape(-1), "sharded __call__", rtol=1e-5, atol_scale=1e-6)


def test_arkit_shaped_full_size_subset():
    """configs[3]: per-view intrinsics, near/far [0.5, 5.5], 40 views (train)."""
    from mvsdet_b200 import ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = SceneConfig(n_views=40, near_far_range=(0.5, 5.5), per_view_intrinsics=True)
    scene = make_scene(cfg, seed=8)
    hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk)
    vs = 2
    cost = scene["cost_out"][:vs]
    ref = O.hot_path(scene["feature"], scene["img_meta"], lambda var: cost,
                     near_far_range=cfg.near_far_range, num_depth=cfg.num_depth, topk=cfg.topk,
                     n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, stride=cfg.stride,
                     training=False, view_subset=vs)
    dev = torch.device("cuda")
    feat_cl = ops.pack_features(scene["feature"].to(dev), torch.float32)
    geo = hot.geometry(scene["img_meta"], dev, view_slice=slice(0, vs))
    _close(hot.variance(feat_cl, geo), ref["variance"], "arkit variance[0:2]")
    _, _, est_depth, est_dens, est_idx, _ = hot.hypotheses(cost.to(dev))
    assert np.array_equal(est_idx.cpu().numpy(), ref["est_idx"].numpy())
    vol, count = ops.backproject_aggregate(feat_cl[:vs], geo.points, geo.projection, est_depth,
                                           est_dens, cfg.voxel_size[2], geo.height, geo.width)
    assert np.array_equal(count.cpu().numpy(), ref["count"].reshape(-1).numpy().astype(np.int32))


@pytest.mark.parametrize("d", [16, 32, 64])
def test_depth_plane_sweep_config(d):
    """configs[4]: D in {16..64} at a small spatial size against the oracle."""
    cfg = tiny_config(n_views=4, channels=32, num_depth=d)
    scene = make_scene(cfg, seed=d)
    ref = oracle_chain(scene)
    from test_gpu_parity import cuda_chain
    res = cuda_chain(scene)
    assert np.array_equal(res["est_idx"].cpu().numpy(), ref["est_idx"].numpy())
    assert np.array_equal(res["count"].cpu().numpy().reshape(ref["count"].shape), ref["count"].numpy())
    for key in ("variance", "volume_mean", "g_feature_from_variance", "g_feature_from_voxels"):
        _close(res[key], ref[key], f"D={d}:{key}")
