"""NVS-branch consumers emitted by the top-k kernel against the goldens
produced by the reference's own functions (compute_depth_scale[_MultiIntrin] mvsdet.py:1158-1218,
ray-depth conversions :488-494 / :583, opacity :579, process_rgb_raw :319-333), and their
gradients against the oracle's autograd."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, load_golden
from oracle import mvsdet_oracle as O

pytestmark = pytest.mark.gpu


def _close(a, b, what, tol=1e-6):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b)).double()
    assert a.shape == b.shape, f"{what}: {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    bad = err > tol * (1 + b.abs())
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} off, max err {float(err.max()):.3e}"


def _hot(cfg):
    from mvsdet_b200.hotpath import MVSDetHotPath
    return MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                         stride=cfg.stride)


@pytest.mark.parametrize("case", ["scannet_tiny", "arkit_tiny"])
def test_nvs_outputs_match_reference_golden(case):
    scene, gold = load_golden(case)
    nvs = np.load(os.path.join(GOLDEN_DIR, f"nvs_{case}.npz"))
    cfg = scene["cfg"]
    dev = torch.device("cuda")
    res = _hot(cfg)(scene["feature"].to(dev), scene["img_meta"],
                    cost_regularization=lambda var: scene["cost_out"].to(dev), nvs=True)
    # opacity is a selection of the kernel's own probabilities (bit-equal to the top-1 density);
    # against the reference it carries the softmax's rounding like prob_volume does
    assert torch.equal(res["opacity"], res["est_densities"][:, 0])
    assert torch.equal(res["opacity"], res["prob_volume"].max(dim=1)[0])
    _close(res["opacity"], nvs["opacity"], f"{case}: opacity", tol=1e-5)
    _close(res["depth_scale"], nvs["depth_scale"][0], f"{case}: depth_scale")
    _close(res["est_ray_depth"], nvs["est_ray_depth"], f"{case}: est_ray_depth", tol=1e-5)
    _close(res["ray_depth_coding"], nvs["ray_depth_coding"][0], f"{case}: ray_depth_coding", tol=1e-5)
    # the six standard outputs are unchanged by the epilogue
    assert np.array_equal(res["est_idx"].cpu().numpy(), gold["est_idx"])
    _close(res["est_depth"], gold["est_depth"], "est_depth", tol=1e-5)


@pytest.mark.parametrize("case", ["scannet_tiny", "arkit_tiny"])
def test_functional_mirrors(case):
    from mvsdet_b200 import functional as F_
    scene, _ = load_golden(case)
    nvs = np.load(os.path.join(GOLDEN_DIR, f"nvs_{case}.npz"))
    cfg = scene["cfg"]
    h, w = cfg.crop_hw
    fn = F_.compute_depth_scale_MultiIntrin if cfg.per_view_intrinsics else F_.compute_depth_scale
    scale = fn(h, w, torch.device("cuda"), scene["img_meta"], cfg.stride, cfg.n_views)
    assert tuple(scale.shape) == nvs["depth_scale"].shape
    _close(scale, nvs["depth_scale"], "compute_depth_scale mirror")
    rgb = F_.process_rgb_raw(torch.from_numpy(nvs["rgb"]).cuda(), cfg.stride, h, w, nvs["src_id"].tolist())
    assert tuple(rgb.shape) == nvs["rgb_raw"].shape
    _close(rgb, nvs["rgb_raw"], "process_rgb_raw mirror")
    with pytest.raises(ValueError):
        F_.process_rgb_raw(torch.from_numpy(nvs["rgb"]).cuda(), 2, h, w, [0])     # reference asserts ratio == 4


@pytest.mark.parametrize("case", ["scannet_tiny", "arkit_tiny"])
def test_nvs_gradients_match_oracle_autograd(case):
    """dL/dcost_out through est_ray_depth, ray_depth_coding and opacity."""
    from mvsdet_b200 import ops
    scene, _ = load_golden(case)
    cfg = scene["cfg"]
    h, w = cfg.crop_hw
    v, t = cfg.n_views, cfg.topk
    gen = torch.Generator().manual_seed(3)
    cost = scene["cost_out"].clone().requires_grad_(True)
    prob, off = O.depth_probability(cost)
    est_depth, est_dens = O.sample_depth_prob(prob, off, t, cfg.near_far_range[0], cfg.depth_interval)
    coding = O.compute_avg_depth(prob, off, cfg.near_far_range[0], cfg.depth_interval)
    est_depth_r = est_depth[:, :, :h, :w].reshape(v, t, -1).transpose(2, 1).unsqueeze(2)
    ref = O.nvs_consumers(prob, est_depth_r, coding[:, :h, :w].unsqueeze(1), scene["img_meta"], cfg.stride, h, w)
    g_ray = torch.randn(ref["est_ray_depth"].shape, generator=gen)
    g_cod = torch.randn(ref["ray_depth_coding"].shape, generator=gen)
    g_opa = torch.randn(ref["opacity"].shape, generator=gen)
    want, = torch.autograd.grad([ref["est_ray_depth"], ref["ray_depth_coding"], ref["opacity"]], cost,
                                [g_ray, g_cod, g_opa])

    dev = torch.device("cuda")
    hot = _hot(cfg)
    geo = hot.geometry(scene["img_meta"], dev)
    cost_c = scene["cost_out"].to(dev).requires_grad_(True)
    outs = ops.depth_topk_nvs(cost_c, cfg.near_far_range[0], cfg.depth_interval, t, geo.k_feat)
    opacity, scale, ray_depth, ray_coding = outs[6:]
    hf, wf = cfg.feat_hw
    # the oracle's upstream gradients live on the [h,w] crop in the reference's layouts
    g_ray_full = torch.zeros(v, t, hf, wf)
    g_ray_full[:, :, :h, :w] = g_ray.squeeze(2).transpose(2, 1).reshape(v, t, h, w)
    g_cod_full = torch.zeros(v, hf, wf)
    g_cod_full[:, :h, :w] = g_cod.reshape(v, h, w)
    got, = torch.autograd.grad([ray_depth, ray_coding, opacity], cost_c,
                               [g_ray_full.to(dev), g_cod_full.to(dev), g_opa.to(dev)])
    err = (got.cpu().double() - want.double()).abs()
    rms = float(want.double().pow(2).mean().sqrt())
    assert float(err.max()) <= 1e-4 * rms + 1e-4 * float(want.abs().max()), f"max err {float(err.max()):.3e}, rms {rms:.3e}"


def test_nan_logits_give_distinct_hypotheses():
    """a pixel whose logits are NaN / inf has an all-NaN softmax; the kernel must
    still return T distinct plane indices (torch.topk does), not T copies of plane 0."""
    from mvsdet_b200 import ops
    cost = torch.randn(2, 2, 12, 5, 7, device="cuda")
    cost[0, 0, 3, 2, 2] = float("nan")
    cost[1, 0, 5, 1, 1] = float("inf")
    prob, off, est_depth, est_dens, est_idx, coding = ops.depth_topk(cost, 0.2, 0.4, 3)
    for (v, y, x) in ((0, 2, 2), (1, 1, 1)):
        ids = est_idx[v, :, y, x].tolist()
        assert len(set(ids)) == 3, ids
        assert bool(torch.isnan(est_dens[v, :, y, x]).all())
    srt = est_idx.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
