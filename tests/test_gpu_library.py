"""GPU: the dispatcher-registered ops (torch.ops.mvsdet_b200.*, mvsdet_b200/library.py)
against the reference's golden vectors, against the autograd.Function layer (same
kernels: forward outputs must be bit-identical) and through torch.library.opcheck
(schema, fake-tensor agreement of sizes/strides/dtypes, autograd registration)."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from test_gpu_parity import _close, _module, cuda_chain

pytestmark = pytest.mark.gpu


def _dispatcher_chain(scene, feature_dtype=torch.float32, channels_first=True):
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = scene["cfg"]
    dev = torch.device("cuda")
    feature = scene["feature"].to(dev).requires_grad_(True)
    cost_out = scene["cost_out"].to(dev).requires_grad_(True)
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        stride=cfg.stride, feature_dtype=feature_dtype,
                        channels_first_volume=channels_first, dispatcher_ops=True)
    res = mod(feature, scene["img_meta"], cost_regularization=lambda var: cost_out)
    g1, = torch.autograd.grad(res["variance"], feature, scene["g_variance"].to(dev), retain_graph=True)
    g2, g3 = torch.autograd.grad(res["volume_mean"], (feature, cost_out), scene["g_volume_mean"].to(dev))
    res["g_feature_from_variance"], res["g_feature_from_voxels"], res["g_cost_out"] = g1, g2, g3
    torch.cuda.synchronize()
    return res


@pytest.mark.parametrize("case", ["scannet_tiny", "arkit_tiny"])
@pytest.mark.parametrize("channels_first", [True, False])
def test_dispatcher_chain_vs_reference_golden(case, channels_first):
    from mvsdet_b200 import _lib
    scene, gold = load_golden(case)
    before = _lib.launch_count()
    res = _dispatcher_chain(scene, channels_first=channels_first)
    assert _lib.launch_count() - before >= 8, "the dispatcher ops must launch the library's kernels"
    assert np.array_equal(res["est_idx"].cpu().numpy(), gold["est_idx"])
    assert np.array_equal(res["count"].cpu().numpy().reshape(gold["count"].shape), gold["count"])
    for key in ("variance", "prob_volume", "off_pred", "est_depth", "est_densities", "depth_coding",
                "volume_mean", "g_feature_from_variance", "g_feature_from_voxels"):
        _close(res[key], gold[key], f"{case}:{key}")
    _close(res["g_cost_out"], gold["g_cost_out"], f"{case}:g_cost_out", abs_floor=1e-5)
    # same launchers as the autograd.Function layer: forward results are the same bits
    ref = cuda_chain(scene, channels_first=channels_first, with_grads=False)
    for key in ("variance", "prob_volume", "est_depth", "est_densities", "est_idx", "volume_mean", "count"):
        assert torch.equal(res[key], ref[key]), key


def test_opcheck():
    from mvsdet_b200 import library as L
    from mvsdet_b200 import ops
    scene, gold = load_golden("scannet_tiny")
    cfg = scene["cfg"]
    mod = _module(cfg)
    geo = mod.geometry(scene["img_meta"], torch.device("cuda"))
    feat = ops.pack_features(scene["feature"].cuda(), torch.float32).detach().requires_grad_(True)
    cost_out = scene["cost_out"].cuda().requires_grad_(True)
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(L.plane_sweep_variance, (feat, geo.neighbor_ids, geo.hom, geo.depth_values),
                          test_utils=checks)
    torch.library.opcheck(L.depth_topk, (cost_out, cfg.near_far_range[0], cfg.depth_interval, cfg.topk),
                          test_utils=checks)
    est = L.depth_topk(cost_out.detach(), cfg.near_far_range[0], cfg.depth_interval, cfg.topk)
    est_dens = est[3].detach().requires_grad_(True)
    for channels_first in (True, False):
        torch.library.opcheck(L.backproject_aggregate,
                              (feat, geo.points, geo.projection, est[2], est_dens, cfg.voxel_size[2],
                               geo.height, geo.width, False, channels_first), test_utils=checks)


def test_dispatcher_ops_reject_bad_arguments():
    from mvsdet_b200 import library as L
    x = torch.randn(2, 8, 4, 4, device="cuda")                 # NCHW, not channels_last
    nbr = torch.tensor([[1], [0]], dtype=torch.int32, device="cuda")
    hom = torch.zeros(2, 1, 12, device="cuda")
    dv = torch.ones(2, 4, device="cuda")
    with pytest.raises(ValueError):
        L.plane_sweep_variance(x, nbr, hom, dv)
    cl = x.contiguous(memory_format=torch.channels_last)
    with pytest.raises(ValueError):
        L.plane_sweep_variance(cl, nbr, hom, dv, False, 1)     # reference views exceed feat
    with pytest.raises(ValueError):
        L.depth_topk(torch.zeros(2, 3, 4, 4, 4, device="cuda"), 0.2, 0.4, 3)
