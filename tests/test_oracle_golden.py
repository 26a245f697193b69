"""Pin the CPU oracle against the golden vectors produced by the reference.

The goldens are outputs of the reference's own functions
(projects/NeRF-Det/nerfdet/mvsdet.py, mvs_models/module.py) executed by
tests/golden/make_golden.py; see that script for the exact call chain.
Integer / boolean outputs must match bit for bit; floats to 1e-6 (the oracle
calls the same ATen ops, differences can only come from op fusion order).
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, assert_close, load_golden, oracle_chain
from oracle import mvsdet_oracle as O

FLOAT_KEYS = ("variance", "prob_volume", "off_pred", "est_depth", "est_densities",
              "depth_coding", "volume_mean", "projection", "points",
              "g_feature_from_variance", "g_feature_from_voxels", "g_cost_out")
EXACT_KEYS = ("neighbor_ids", "est_idx", "valid", "count")


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference_golden(case):
    scene, gold = load_golden(case)
    res = oracle_chain(scene)
    for key in EXACT_KEYS:
        got = res[key].numpy()
        assert got.shape == gold[key].shape, key
        assert np.array_equal(got, gold[key]), f"{case}:{key} not bit-exact"
    for key in FLOAT_KEYS:
        assert_close(res[key], gold[key], rtol=1e-6, atol=1e-6, what=f"{case}:{key}")


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_eval_branch_equals_train_branch(case):
    """mvsdet.py:463-464 (in-place, eval) and :458-459 (out-of-place, train)
    produce the same variance."""
    scene, gold = load_golden(case)
    res = oracle_chain(scene, training=False, with_grads=False)
    assert_close(res["variance"], gold["variance"], rtol=1e-6, atol=1e-6, what="variance")


@pytest.mark.parametrize("case", GOLDEN_CASES[:2])
def test_closed_form_warp_agrees_with_grid_sample(case):
    """Appendix A.1's closed form (ix = px*W/(W-1) - 0.5, zero padding) is what
    F.grid_sample(align_corners=False) computes on the reference's grid."""
    scene, gold = load_golden(case)
    feat = scene["feature"]
    ref_proj = torch.from_numpy(gold["ref_proj"])
    nei_proj = torch.from_numpy(gold["nei_projs"])[0]
    nbr = torch.from_numpy(gold["neighbor_ids"])
    cfg = scene["cfg"]
    dv = torch.from_numpy(O.depth_values_for(cfg.near_far_range, cfg.num_depth))
    dv = dv.unsqueeze(0).repeat(feat.shape[0], 1)
    src = feat[nbr[:, 0]]
    a = O.homo_warping(src, nei_proj, ref_proj, dv)
    rot, trans = O.homography(nei_proj, ref_proj)
    b = O.warp_closed_form(src, rot, trans, dv)
    assert_close(a[:, :4], gold["warped0"], rtol=1e-6, atol=1e-6, what="warped vs golden")
    assert_close(b, a, rtol=1e-4, atol=2e-4, what="closed form vs grid_sample")
