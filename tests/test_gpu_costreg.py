"""the drop-in with the REAL cost-regularisation architecture attached
(CostRegNet_3DGS, mvs_models/mvsnet.py:73-113; restated in oracle/costreg_oracle.py because the GPU
box cannot read /root/reference, pinned against the reference class in
tests/test_oracle_vs_reference.py) instead of the synthetic stand-in tensor.

Checks: the whole chain (sweep -> cuDNN U-Net -> top-k -> back-projection) forward + backward
against the oracle chain with the same net on the CPU; the variance is handed to cuDNN in
channels_last_3d and the gradient comes back in the same layout -- NO hidden 1.2 GB re-layout
(ops.relayout_count stays 0); the bf16 variance hand-off runs."""
import numpy as np
import pytest
import torch

from helpers import oracle_chain
from mvsdet_b200.scene import make_scene, tiny_config
from oracle import mvsdet_oracle as O
from oracle.costreg_oracle import CostRegNet3DGS

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt().clamp_min(1e-30))


@pytest.fixture(scope="module")
def setup():
    cfg = tiny_config(n_views=3, channels=256, num_depth=8)       # D=8, 12x16 maps: divisible by 4
    scene = make_scene(cfg, seed=9)
    torch.manual_seed(0)
    net = CostRegNet3DGS().eval()
    return cfg, scene, net


def test_real_costreg_net_chain_matches_oracle(setup):
    from mvsdet_b200 import ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg, scene, net = setup
    # ---- oracle chain with the net on the CPU
    feature = scene["feature"].clone().requires_grad_(True)
    ref = O.hot_path(feature, scene["img_meta"], net, near_far_range=cfg.near_far_range,
                     num_depth=cfg.num_depth, topk=cfg.topk, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                     stride=cfg.stride, training=True)
    g_ref, = torch.autograd.grad(ref["volume_mean"], feature, scene["g_volume_mean"])
    # ---- drop-in with the net on the GPU (cuDNN, fp32 -- TF32 off for the comparison)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        dev = torch.device("cuda")
        import copy
        net_c = copy.deepcopy(net).to(dev).to(memory_format=torch.channels_last_3d)
        hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                            cost_regularization=net_c, stride=cfg.stride)
        feat_c = scene["feature"].to(dev).requires_grad_(True)
        before = dict(ops.relayout_count)
        res = hot(feat_c, scene["img_meta"])
        g, = torch.autograd.grad(res["volume_mean"], feat_c, scene["g_volume_mean"].to(dev))
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert res["variance"].permute(0, 2, 3, 4, 1).is_contiguous(), "variance must reach cuDNN in channels_last_3d"
    assert ops.relayout_count == before, f"a hidden re-layout of the variance gradient fired: {ops.relayout_count}"
    assert _rel(res["variance"], ref["variance"]) < 1e-4
    assert _rel(res["cost_out"], ref["cost_out"]) < 2e-3
    same = (res["est_idx"].cpu() == ref["est_idx"]).float().mean()
    assert float(same) > 0.995, f"top-k agreement {float(same):.4f}"      # conv rounding can swap near-tied planes
    if float(same) == 1.0:
        assert np.array_equal(res["count"].cpu().numpy().reshape(-1), ref["count"].reshape(-1).numpy())
        assert _rel(res["volume_mean"], ref["volume_mean"]) < 2e-3
        assert _rel(g, g_ref) < 2e-2          # through a 12-layer fp32 U-Net on two different conv libraries


def test_bf16_variance_handoff_runs(setup):
    """variance_dtype=bf16: the sweep writes bf16 channels_last_3d (half the bytes), the net runs in
    bf16, its gradient returns in bf16 channels_last_3d and feeds the bf16 backward instantiation."""
    from mvsdet_b200 import ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    import copy
    cfg, scene, net = setup
    dev = torch.device("cuda")
    net_c = copy.deepcopy(net).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last_3d)
    hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        cost_regularization=lambda var: net_c(var).float(), stride=cfg.stride,
                        feature_dtype=torch.bfloat16, variance_dtype=torch.bfloat16)
    feat_c = scene["feature"].to(dev).requires_grad_(True)
    before = dict(ops.relayout_count)
    res = hot(feat_c, scene["img_meta"])
    assert res["variance"].dtype == torch.bfloat16
    g, = torch.autograd.grad(res["volume_mean"], feat_c, scene["g_volume_mean"].to(dev))
    torch.cuda.synchronize()
    assert ops.relayout_count == before
    assert g.dtype == torch.float32 and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
    # against the fp32 hand-off: bf16 rounding of a cancelling variance is large in relative terms,
    # so only a loose agreement of the variance itself is asserted
    hot32 = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                          cost_regularization=lambda var: net_c(var.to(torch.bfloat16)).float(), stride=cfg.stride,
                          feature_dtype=torch.bfloat16)
    res32 = hot32(feat_c.detach(), scene["img_meta"])
    a, b = res["variance"].float(), res32["variance"]
    # a bf16 output is the fp32 value rounded to 8 significant bits: |a-b| <= 2^-8 |b| (+ denormal slack)
    assert bool(((a - b).abs() <= 2.0 ** -8 * b.abs() + 1e-30).all())
