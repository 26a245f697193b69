"""Seed sweep of the bit-exact outputs at the shipped voxel grid (SURVEY.md 8d "guard band").

60 seeded ScanNet-shaped scenes (V=20, 60x80 maps, 40x40x16 voxels; 8 channels -- the
integer outputs do not depend on the channel count): neighbour ids, top-k indices, per-view
valid masks and voxel counts must equal the oracle's bit for bit.  A guard band that
*rejects* near-tie seeds is not usable at this size (512 000 voxel projections per scene put
~200 of them within 1e-4 px of a .5 rounding boundary in every scene), so instead the
near-ties are COUNTED and reported -- |frac(x) - .5| < 1e-4 px, |z - (d +- vs_z)| < 1e-5 m,
top-k probability gaps < 1e-7, second/third-neighbour distance gaps < 1e-6 relative -- and
the bit-exact assertions cover the geometric ones: the kernels reproduce the reference's rounding
(FMA chain of bmm, rintf, strict comparisons), they do not merely avoid the boundaries.
The one place where bit-exactness cannot be had is the ORDER of two depth planes whose softmax
probabilities agree to the last ulp: CUDA's expf and ATen's (Sleef) exp round differently.  Such
swaps are required to be near-ties (relative gap <= 1e-6) and are counted in the report.
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
N_SEEDS = 60


def near_ties(scene, prob, est_depth, projection, points, nbr_margin_only=False):
    """Counts of near-tie events in one scene (oracle-side quantities, fp64 analysis)."""
    cfg = scene["cfg"]
    h, w = cfg.crop_hw
    out = {}
    w2c = torch.as_tensor(np.array(scene["img_meta"]["lidar2img"]["extrinsic"]))
    loc = w2c.inverse()[:, :3, 3].double()
    d2 = (loc[:, None] - loc[None]).pow(2).sum(-1)
    d2.fill_diagonal_(float("inf"))
    srt = d2.sort(dim=1).values
    k = min(cfg.num_neighbors, cfg.n_views - 1)
    gaps = []
    for j in range(k):
        if j + 1 < srt.shape[1]:
            gaps.append(((srt[:, j + 1] - srt[:, j]) / srt[:, j + 1]).min())
    out["knn_min_rel_gap"] = float(min(gaps)) if gaps else None
    out["knn_near_ties"] = int(sum(int(((srt[:, j + 1] - srt[:, j]) / srt[:, j + 1] < 1e-6).sum())
                                   for j in range(k) if j + 1 < srt.shape[1]))
    if nbr_margin_only:
        return out
    v = cfg.n_views
    pts = points.reshape(1, 3, -1).expand(v, 3, -1)
    pts = torch.cat((pts, torch.ones_like(pts[:, :1])), dim=1)
    p = torch.bmm(projection, pts)
    fx, fy, z = (p[:, 0] / p[:, 2]).double(), (p[:, 1] / p[:, 2]).double(), p[:, 2]
    inb = (fx > -0.5) & (fy > -0.5) & (fx < w - 0.5) & (fy < h - 0.5) & (z > 0)
    tie_xy = ((fx - fx.floor() - 0.5).abs() < 1e-4) | ((fy - fy.floor() - 0.5).abs() < 1e-4)
    out["round_near_ties"] = int((tie_xy & inb).sum())
    x = fx.round().long().clamp(0, w - 1)
    y = fy.round().long().clamp(0, h - 1)
    pix = (y * w + x)
    d_at = torch.gather(est_depth[:, :, :h, :w].reshape(v, cfg.topk, -1).transpose(2, 1), 1,
                        pix.unsqueeze(-1).expand(v, pix.shape[1], cfg.topk)).double()
    vs = float(cfg.voxel_size[2])
    zz = z.double().unsqueeze(-1)
    tie_z = (((zz - (d_at - vs)).abs() < 1e-5) | ((zz - (d_at + vs)).abs() < 1e-5)).any(-1)
    out["depth_test_near_ties"] = int((tie_z & inb).sum())
    srt_p = prob.sort(dim=1, descending=True).values.double()
    t = cfg.topk
    gap = (srt_p[:, :t] - srt_p[:, 1:t + 1])
    out["topk_near_ties"] = int((gap < 1e-7).sum())
    return out


def test_sixty_seed_sweep_bit_exact_with_near_tie_report():
    from mvsdet_b200 import ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = SceneConfig(n_views=20, channels=8)
    hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        stride=cfg.stride)
    dev = torch.device("cuda")
    h, w = cfg.crop_hw
    v, t = cfg.n_views, cfg.topk
    totals = {"seeds": 0, "round_near_ties": 0, "depth_test_near_ties": 0, "topk_near_ties": 0,
              "knn_near_ties": 0, "knn_min_rel_gap": 1.0, "topk_swapped_pixels": 0,
              "topk_swap_max_rel_gap": 0.0, "topk_swaps_changing_the_set": 0, "mismatched_seeds": []}
    for seed in range(100, 100 + N_SEEDS):
        scene = make_scene(cfg, seed=seed, with_grads=False)
        # oracle
        prob, off = O.depth_probability(scene["cost_out"])
        est_depth, est_dens, est_idx = O.sample_depth_prob(prob, off, t, cfg.near_far_range[0],
                                                           cfg.depth_interval, return_idx=True)
        ratio = scene["img_meta"]["ori_shape"][0] / (scene["img_meta"]["img_shape"][0] / cfg.stride)
        projection = O.compute_projection(scene["img_meta"]["lidar2img"]["intrinsic"],
                                          scene["img_meta"]["lidar2img"]["extrinsic"], ratio)
        points = O.get_points(cfg.n_voxels, cfg.voxel_size, scene["img_meta"]["lidar2img"]["origin"])
        depth_r = est_depth[:, :, :h, :w].reshape(v, t, -1).transpose(2, 1).unsqueeze(2)
        dens_r = est_dens[:, :, :h, :w].reshape(v, t, -1).transpose(2, 1).unsqueeze(2)
        _, valid = O.backproject_weigh(scene["feature"][:, :, :h, :w], points, projection, depth_r,
                                       cfg.voxel_size, dens_r)
        count = valid.sum(dim=0).reshape(-1)
        w2c = torch.as_tensor(np.array(scene["img_meta"]["lidar2img"]["extrinsic"]))
        nbr = O.get_nearest_pose_ids(w2c.inverse(), 2)
        # kernels
        geo = hot.geometry(scene["img_meta"], dev)
        feat_cl = ops.pack_features(scene["feature"].to(dev), torch.float32)
        _, _, g_depth, g_dens, g_idx, _ = hot.hypotheses(scene["cost_out"].to(dev))
        _, g_count = ops.backproject_aggregate(feat_cl, geo.points, geo.projection, g_depth, g_dens,
                                               cfg.voxel_size[2], geo.height, geo.width)
        _, g_valid = ops.backproject_per_view(feat_cl, geo.points, geo.projection, depth_r.to(dev),
                                              dens_r.to(dev), cfg.voxel_size[2], h, w)
        nt = near_ties(scene, prob, est_depth, projection, points)
        totals["seeds"] += 1
        for key in ("round_near_ties", "depth_test_near_ties", "topk_near_ties", "knn_near_ties"):
            totals[key] += nt[key]
        totals["knn_min_rel_gap"] = min(totals["knn_min_rel_gap"], nt["knn_min_rel_gap"])
        # (1) neighbour ids: bit-exact, always
        if not np.array_equal(geo.neighbor_ids.cpu().numpy(), nbr.numpy()):
            totals["mismatched_seeds"].append((seed, "neighbor_ids"))
        # (2) per-view masks and counts from the ORACLE's hypotheses: bit-exact, always -- this is
        #     where the rounding / depth-test near-ties live
        if not np.array_equal(g_valid.cpu().numpy().reshape(valid.shape), valid.numpy()):
            totals["mismatched_seeds"].append((seed, "valid"))
        # (3) top-k indices: the kernel's softmax (CUDA expf) and ATen's (Sleef) round the
        #     probabilities differently in the last ulp, so two planes whose probabilities agree to
        #     ~1e-7 relative can swap places.  Every differing pixel must be such a near-tie; they
        #     are counted, not hidden.
        diff = (g_idx.cpu() != est_idx)
        n_diff = int(diff.any(dim=1).sum())
        if n_diff:
            pix = diff.any(dim=1)                                       # [V,H,W]
            p_pix = prob.permute(0, 2, 3, 1)[pix]                       # [n, D]
            a_idx = est_idx.permute(0, 2, 3, 1)[pix]
            b_idx = g_idx.cpu().permute(0, 2, 3, 1)[pix]
            pa = torch.gather(p_pix, 1, a_idx).double()
            pb = torch.gather(p_pix, 1, b_idx).double()
            rel = ((pa - pb).abs() / pa.clamp_min(1e-30)).max()
            totals["topk_swapped_pixels"] += n_diff
            totals["topk_swap_max_rel_gap"] = max(totals["topk_swap_max_rel_gap"], float(rel))
            same_set = (a_idx.sort(dim=1).values == b_idx.sort(dim=1).values).all(dim=1)
            totals["topk_swaps_changing_the_set"] += int((~same_set).sum())
            if float(rel) > 1e-6:
                totals["mismatched_seeds"].append((seed, "est_idx beyond a near-tie"))
        else:
            # (4) with identical hypotheses the kernel's own count must equal the oracle's
            if not np.array_equal(g_count.cpu().numpy(), count.numpy().astype(np.int32)):
                totals["mismatched_seeds"].append((seed, "count"))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "seed_sweep_report.json"), "w") as fh:
            json.dump(totals, fh, indent=1)
    except OSError:
        pass
    print("seed sweep:", json.dumps(totals))
    assert totals["round_near_ties"] > 0, "the sweep is supposed to exercise rounding near-ties"
    assert not totals["mismatched_seeds"], f"integer outputs differ on seeds {totals['mismatched_seeds']}"
