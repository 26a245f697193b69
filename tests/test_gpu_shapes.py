"""Shapes off the shipped configuration: two 256-channel slices per pixel (C = 512), a ragged second slice
(C = 384), one 128-channel group (C = 128, C = 64), two views (k = 1), 64 depth planes -- whole chain
(variance, top-k, voxels and all three gradients) against the oracle on the same features."""
import numpy as np
import pytest
import torch

from helpers import oracle_chain
from mvsdet_b200.scene import make_scene, tiny_config
from test_gpu_parity import _close, cuda_chain

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,v,d,hw,fdt", [
    (512, 4, 6, (10, 12), torch.float32), (512, 4, 6, (10, 12), torch.bfloat16),
    (128, 5, 16, (12, 16), torch.bfloat16), (384, 3, 8, (9, 11), torch.bfloat16),
    (256, 2, 12, (8, 8), torch.bfloat16), (64, 6, 64, (6, 8), torch.float32)],
    ids=["C512_f32", "C512_bf16", "C128_D16", "C384", "two_views_C256", "D64"])
def test_chain_on_unusual_shapes(c, v, d, hw, fdt):
    h, w = hw
    cfg = tiny_config(n_views=v, channels=c, num_depth=d, img_shape=(4 * h - 1, 4 * w), pad_shape=(4 * h, 4 * w),
                      ori_shape=(16 * h - 4, 16 * w), near_far_range=(0.2, 0.2 + 0.4 * d))
    scene = make_scene(cfg, seed=5)
    if fdt == torch.bfloat16:
        scene["feature"] = scene["feature"].to(torch.bfloat16).float()      # both sides see the rounded features
    ref = oracle_chain(scene)
    res = cuda_chain(scene, feature_dtype=fdt)
    assert np.array_equal(res["count"].cpu().numpy().reshape(ref["count"].shape), ref["count"].numpy())
    assert np.array_equal(res["est_idx"].cpu().numpy(), ref["est_idx"].numpy())
    for key in ("variance", "volume_mean", "g_feature_from_variance", "g_feature_from_voxels", "g_cost_out"):
        _close(res[key], ref[key], f"C={c} V={v} D={d}: {key}", abs_floor=1e-5 if key == "g_cost_out" else 0.0)
