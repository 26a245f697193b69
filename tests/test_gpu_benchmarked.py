"""Element-wise parity of the BENCHMARKED instantiation (BASELINE.json configs[1]:
V=20, 60x80x256 features, D=12, k=2, T=3, 40x40x16 voxels) through the same object
bench.py times -- ``ScenePipeline.step()`` -- for bf16 and fp32 features.

The CPU oracle needs minutes for the 20-view plane sweep, so the sweep is checked on
the first two reference views: the upstream gradient ``g_variance`` is zero for every
other reference view, which makes the 20-view backward launch (same grid, same kernel
template as the bench) equal to the oracle's ``view_subset=2`` backward.  The
back-projection, prob-norm and top-k kernels are checked on all 20 views (their
oracle takes seconds).

Bars (north_star): integers bit-exact; floats ``|a-b| <= tol*rms(ref) + tol*|ref|`` with
tol = 1e-4 for fp32 features, 1e-2 for bf16 features.  The plain maximum relative error
over the elements with |ref| > rms(ref) is reported next to it (and bounded), and all
numbers land in ``gpurun_out/parity_report.json``.
"""
import json
import os

import numpy as np
import pytest
import torch

from mvsdet_b200.scene import SceneConfig, make_scene
from oracle import mvsdet_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.json")
VS = 2                      # reference views the sweep oracle covers


def _report(key, value):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        data = {}
        if os.path.isfile(REPORT):
            with open(REPORT) as fh:
                data = json.load(fh)
        data[key] = value
        with open(REPORT, "w") as fh:
            json.dump(data, fh, indent=1, sort_keys=True)
    except OSError:
        pass


def _check(a, b, what, tol, abs_floor=0.0):
    """rms-anchored tolerance + the plain relative error on the large elements."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: {tuple(a.shape)} vs {tuple(b.shape)}"
    rms = float(b.pow(2).mean().sqrt())
    err = (a - b).abs()
    bad = err > max(tol * rms, abs_floor) + tol * b.abs()
    big = b.abs() > rms
    max_rel_big = float((err[big] / b.abs()[big]).max()) if bool(big.any()) else 0.0
    _report(what, {"max_abs_err": float(err.max()), "rms_ref": rms, "tol": tol,
                   "max_rel_err_where_ref_gt_rms": max_rel_big, "elements": b.numel(),
                   "out_of_tolerance": int(bad.sum())})
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} out of tolerance, "
                                 f"max abs err {float(err.max()):.3e}, rms {rms:.3e}")
    # plain relative error where the reference is not small: 2x the bar at most
    assert max_rel_big <= 2 * tol, f"{what}: max relative error on |ref|>rms is {max_rel_big:.3e}"


@pytest.fixture(scope="module")
def scene():
    return make_scene(SceneConfig(n_views=20), seed=3)


def _pipeline(scene, feature_dtype, variance_dtype=torch.float32):
    from mvsdet_b200.hotpath import MVSDetHotPath
    from mvsdet_b200.pipeline import ScenePipeline
    cfg = scene["cfg"]
    dev = torch.device("cuda")
    hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        stride=cfg.stride)
    pipe = ScenePipeline(cfg, dev, feature_dtype=feature_dtype, variance_dtype=variance_dtype)
    pipe.set_geometry(hot.geometry(scene["img_meta"], dev))
    return pipe


def _rounded(scene, feature_dtype):
    """the scene with features the kernels and the oracle both see exactly"""
    s = dict(scene)
    if feature_dtype == torch.bfloat16:
        s["feature"] = scene["feature"].to(torch.bfloat16).float()
    return s


@pytest.mark.parametrize("feature_dtype,variance_dtype,tol",
                         [(torch.bfloat16, torch.float32, 1e-2), (torch.float32, torch.float32, 1e-4),
                          (torch.bfloat16, torch.bfloat16, 1e-2)],
                         ids=["bf16", "f32", "bf16_bf16var"])
def test_plane_sweep_fwd_bwd_benchmarked_launch(scene, feature_dtype, variance_dtype, tol):
    """Variance of reference views 0..1 and dL/dfeature of the 20-view sweep backward
    (C=256, G=2, k=2: the kernel template bench.py runs) against the oracle's autograd."""
    s = _rounded(scene, feature_dtype)
    cfg = s["cfg"]
    g_var = s["g_variance"].clone()
    g_var[VS:] = 0
    if variance_dtype == torch.bfloat16:
        g_var = g_var.to(torch.bfloat16).float()
    s["g_variance"] = g_var
    s["g_volume_mean"] = torch.zeros_like(s["g_volume_mean"])
    pipe = _pipeline(s, feature_dtype, variance_dtype)
    pipe.load_scene(s)
    pipe.step()
    torch.cuda.synchronize()

    feature = s["feature"].clone().requires_grad_(True)
    cost = s["cost_out"][:VS]
    ref = O.hot_path(feature, s["img_meta"], lambda var: cost, near_far_range=cfg.near_far_range,
                     num_depth=cfg.num_depth, topk=cfg.topk, n_voxels=cfg.n_voxels,
                     voxel_size=cfg.voxel_size, stride=cfg.stride, training=True, view_subset=VS)
    g_ref, = torch.autograd.grad(ref["variance"], feature, g_var[:VS])
    tag = {torch.bfloat16: "bf16", torch.float32: "f32"}
    name = f"sweep[{tag[feature_dtype]} feat, {tag[variance_dtype]} var]"
    var = pipe.variance[:VS].permute(0, 4, 1, 2, 3).float()
    # the variance itself is computed in fp32 from identical inputs: the fp32 bar holds for
    # bf16 features too (only a bf16 *output* is rounded to 2^-9)
    _check(var, ref["variance"].detach(), name + " variance[0:2]",
           1e-2 if variance_dtype == torch.bfloat16 else 1e-4)
    assert np.array_equal(pipe.geo.neighbor_ids.cpu().numpy(), ref["neighbor_ids"].numpy())
    # fp32 accumulation of the same products in a different order (REDs): the fp32 bar again,
    # unless the upstream gradient / output passes through bf16
    _check(pipe.g_feature, g_ref, name + " g_feature_from_variance",
           1e-4 if variance_dtype == torch.float32 else tol)
    # views that are neither reference views 0..1 nor their neighbours get exactly zero
    touched = set(range(VS)) | set(int(x) for x in ref["neighbor_ids"][:VS].reshape(-1))
    for v in range(cfg.n_views):
        if v not in touched:
            assert float(pipe.g_feature[v].abs().max()) == 0.0


@pytest.mark.parametrize("feature_dtype,tol", [(torch.bfloat16, 1e-2), (torch.float32, 1e-4)],
                         ids=["bf16", "f32"])
def test_backprojection_topk_fwd_bwd_benchmarked_launch(scene, feature_dtype, tol):
    """All 20 views at 59x80x256 / 25 600 voxels: top-k selections, per-view valid masks and
    counts bit-exact; volume_mean, dL/dfeature (back-projection gather) and dL/dcost_out
    (back-projection -> prob-norm -> top-k -> softmax backward) against the oracle."""
    from mvsdet_b200 import ops
    s = _rounded(scene, feature_dtype)
    cfg = s["cfg"]
    s["g_variance"] = torch.zeros_like(s["g_variance"])
    pipe = _pipeline(s, feature_dtype)
    pipe.load_scene(s)
    pipe.step()
    torch.cuda.synchronize()

    h, w = cfg.crop_hw
    v, t = cfg.n_views, cfg.topk
    feature = s["feature"].clone().requires_grad_(True)
    cost_out = s["cost_out"].clone().requires_grad_(True)
    prob, off = O.depth_probability(cost_out)
    est_depth, est_dens, est_idx = O.sample_depth_prob(prob, off, t, cfg.near_far_range[0],
                                                       cfg.depth_interval, return_idx=True)
    depth_r = est_depth[:, :, :h, :w].reshape(v, t, -1).transpose(2, 1).unsqueeze(2)
    dens_r = est_dens[:, :, :h, :w].reshape(v, t, -1).transpose(2, 1).unsqueeze(2)
    ratio = s["img_meta"]["ori_shape"][0] / (s["img_meta"]["img_shape"][0] / cfg.stride)
    projection = O.compute_projection(s["img_meta"]["lidar2img"]["intrinsic"],
                                      s["img_meta"]["lidar2img"]["extrinsic"], ratio)
    points = O.get_points(cfg.n_voxels, cfg.voxel_size, s["img_meta"]["lidar2img"]["origin"])
    volume, valid = O.backproject_weigh(feature[:, :, :h, :w], points, projection, depth_r,
                                        cfg.voxel_size, dens_r)
    mean, count = O.aggregate_views(volume, valid)
    g_feat_ref, g_cost_ref = torch.autograd.grad(mean, (feature, cost_out), s["g_volume_mean"])

    name = f"voxels[{'bf16' if feature_dtype == torch.bfloat16 else 'f32'} feat]"
    assert np.array_equal(pipe.est_idx.cpu().numpy(), est_idx.numpy()), "top-k indices"
    assert np.array_equal(pipe.count.cpu().numpy(), count.reshape(-1).numpy().astype(np.int32)), "counts"
    c = cfg.channels
    _check(pipe.volume_mean.view(c, *cfg.n_voxels), mean.detach(), name + " volume_mean", 1e-4)
    _check(pipe.g_feature, g_feat_ref, name + " g_feature_from_voxels", 1e-4)
    _check(pipe.g_cost_out, g_cost_ref, name + " g_cost_out", 1e-4)
    # per-view masks of the same launch geometry (reference API form of the kernel)
    _, valid_gpu = ops.backproject_per_view(pipe.feat_cl.permute(0, 3, 1, 2), pipe.geo.points,
                                            pipe.geo.projection, depth_r.cuda(), dens_r.detach().cuda(),
                                            cfg.voxel_size[2], h, w)
    assert np.array_equal(valid_gpu.cpu().numpy().reshape(valid.shape), valid.numpy()), "valid masks"
    _report(name + " passing (view, voxel) pairs", int(valid.sum()))


def test_whole_step_is_sum_of_parts(scene):
    """With both upstream gradients live the step's g_feature is the sum of the two parts
    checked above (commutative REDs into one accumulator), to fp32 rounding."""
    s = _rounded(scene, torch.bfloat16)
    parts = []
    for zero in ("g_variance", "g_volume_mean", None):
        x = dict(s)
        if zero:
            x[zero] = torch.zeros_like(s[zero])
        pipe = _pipeline(x, torch.bfloat16)
        pipe.load_scene(x)
        pipe.step()
        torch.cuda.synchronize()
        parts.append(pipe.g_feature.double().cpu())
        del pipe
    _check(parts[2].float(), (parts[0] + parts[1]).float(), "step g_feature = sweep part + voxel part", 1e-4)
