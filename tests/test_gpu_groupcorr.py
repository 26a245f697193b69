"""Group-wise correlation cost volume: mvsd_plane_sweep_groupcorr_fwd/bwd through
the C ABI against the oracle (whose arithmetic tests/test_oracle_vs_reference.py pins to the
reference's own statements, mvs_models/lss_fpn.py:496-503) -- forward values and the autograd
gradient, fp32 and bf16 features, ragged maps / channel counts, all supported group widths."""
import numpy as np
import pytest
import torch

from mvsdet_b200.scene import SceneConfig, make_scene, tiny_config
from oracle import mvsdet_oracle as O

pytestmark = pytest.mark.gpu


def _close(a, b, what, rtol=1e-4, atol_scale=1e-4):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    rms = float(b.pow(2).mean().sqrt())
    err = (a - b).abs()
    bad = err > atol_scale * rms + rtol * b.abs()
    assert not bool(bad.any()), f"{what}: {int(bad.sum())}/{bad.numel()} out of tolerance, max abs err {float(err.max()):.3e} (rms {rms:.3e})"


def _run(cfg, seed, groups, feature_dtype):
    from mvsdet_b200 import ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    dev = torch.device("cuda")
    scene = make_scene(cfg, seed, with_grads=False)
    feature = scene["feature"]
    if feature_dtype == torch.bfloat16:
        feature = feature.to(torch.bfloat16).float()            # both sides see the same rounded features
    hf, wf = cfg.feat_hw
    g = torch.randn(cfg.n_views, 2, groups, cfg.num_depth, hf, wf, generator=torch.Generator().manual_seed(seed))
    # oracle (CPU, autograd)
    f_ref = feature.clone().requires_grad_(True)
    want = O.scene_group_correlation(f_ref, scene["img_meta"], near_far_range=cfg.near_far_range,
                                     num_depth=cfg.num_depth, num_groups=groups, stride=cfg.stride)
    want_g, = torch.autograd.grad(want, f_ref, g)
    # CUDA path
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        stride=cfg.stride, feature_dtype=feature_dtype)
    geo = mod.geometry(scene["img_meta"], dev)
    f_dev = feature.to(dev).requires_grad_(True)
    # the drop-in's form: one fp32 gradient accumulator behind the packed (possibly bf16) features
    feat_cl, sink = ops.pack_features(f_dev, feature_dtype, sink=True)
    got = ops.plane_sweep_group_correlation(feat_cl, geo.neighbor_ids, geo.hom, geo.depth_values, groups,
                                            grad_sink=sink)
    got_g, = torch.autograd.grad(got, f_dev, g.to(dev))
    torch.cuda.synchronize()
    assert tuple(got.shape) == tuple(want.shape)
    return got, want.detach(), got_g, want_g


@pytest.mark.parametrize("feature_dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("channels,groups", [(32, 8), (64, 2), (128, 1), (40, 5), (256, 8), (132, 33)])
def test_group_correlation_matches_oracle(channels, groups, feature_dtype):
    cfg = tiny_config(n_views=5, channels=channels, num_depth=6)
    got, want, got_g, want_g = _run(cfg, 31 + channels, groups, feature_dtype)
    _close(got, want, f"C={channels} G={groups}: cost volume")
    _close(got_g, want_g, f"C={channels} G={groups}: dL/dfeature")


def test_group_correlation_ragged_map():
    h, w = 13, 21
    cfg = tiny_config(n_views=4, channels=64, num_depth=5, img_shape=(4 * h - 1, 4 * w), pad_shape=(4 * h, 4 * w),
                      ori_shape=(16 * h - 4, 16 * w))
    got, want, got_g, want_g = _run(cfg, 77, 8, torch.float32)
    _close(got, want, "ragged map: cost volume")
    _close(got_g, want_g, "ragged map: dL/dfeature")


def test_group_correlation_without_sink_returns_feature_dtype_gradient():
    """without a sink the gradient comes back in the packed tensor's dtype (bf16: 2^-9 rounding)"""
    from mvsdet_b200 import ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = tiny_config(n_views=4, channels=32, num_depth=4)
    dev = torch.device("cuda")
    scene = make_scene(cfg, 5, with_grads=False)
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk, stride=cfg.stride)
    geo = mod.geometry(scene["img_meta"], dev)
    feat = scene["feature"].to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = ops.plane_sweep_group_correlation(feat, geo.neighbor_ids, geo.hom, geo.depth_values, 8)
    g, = torch.autograd.grad(out.sum(), feat)
    assert g.dtype == torch.bfloat16 and tuple(g.shape) == tuple(feat.shape) and bool(torch.isfinite(g.float()).all())


def test_group_correlation_rejects_bad_group_width():
    from mvsdet_b200 import ops
    dev = torch.device("cuda")
    feat = torch.zeros(2, 24, 4, 4, device=dev).contiguous(memory_format=torch.channels_last)
    nbr = torch.tensor([[1], [0]], dtype=torch.int32, device=dev)
    hom = torch.zeros(2, 1, 12, device=dev)
    dv = torch.ones(2, 3, device=dev)
    with pytest.raises(ValueError):                      # 24 / 2 = 12 channels per group: not a divisor of 128
        ops.plane_sweep_group_correlation(feat, nbr, hom, dv, num_groups=2)
    with pytest.raises(ValueError):                      # not divisible at all
        ops.plane_sweep_group_correlation(feat, nbr, hom, dv, num_groups=5)


def test_group_correlation_dispatcher_op_matches_autograd_function():
    """torch.ops.mvsdet_b200.plane_sweep_group_correlation (library.py) == ops.plane_sweep_group_correlation,
    values and gradient, plus torch.library.opcheck of the registration"""
    from mvsdet_b200 import library as L, ops
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = tiny_config(n_views=4, channels=64, num_depth=5)
    dev = torch.device("cuda")
    scene = make_scene(cfg, 9, with_grads=False)
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk, stride=cfg.stride)
    geo = mod.geometry(scene["img_meta"], dev)
    base = scene["feature"].to(dev).contiguous(memory_format=torch.channels_last)
    fa, fb = base.clone().requires_grad_(True), base.clone().requires_grad_(True)
    a = ops.plane_sweep_group_correlation(fa, geo.neighbor_ids, geo.hom, geo.depth_values, 8)
    b = L.plane_sweep_group_correlation(fb, geo.neighbor_ids, geo.hom, geo.depth_values, 8)
    assert torch.equal(a, b)
    g = torch.randn_like(a)
    ga, = torch.autograd.grad(a, fa, g)
    gb, = torch.autograd.grad(b, fb, g)
    _close(gb, ga, "dispatcher op gradient", rtol=1e-5, atol_scale=1e-5)      # fp32 RED order differs
    torch.library.opcheck(torch.ops.mvsdet_b200.plane_sweep_group_correlation.default,
                          (base, geo.neighbor_ids, geo.hom, geo.depth_values, 8, 0),
                          test_utils=("test_schema", "test_faketensor"))


def test_hot_path_with_group_correlation_cost_volume():
    """MVSDetHotPath(cost_volume="group_correlation"): a small cost net on the correlation volumes instead of
    CostRegNet_3DGS on the variance; whole chain forward + backward against the oracle with the same net."""
    from mvsdet_b200.hotpath import MVSDetHotPath
    cfg = tiny_config(n_views=5, channels=32, num_depth=8)
    scene = make_scene(cfg, 13)
    groups = 8
    torch.manual_seed(0)
    net = torch.nn.Conv3d(2 * groups, 2, 3, padding=1)                      # [V, k*G, D, H, W] -> [V, 2, D, H, W]
    hf, wf = cfg.feat_hw

    def cost_net_for(mod):
        # the plane ramp breaks the EXACT ties a zero-padded conv produces where no neighbour has a sample
        # (torch.topk and the kernel order exact ties differently: documented deviation)
        def run(vol):
            ramp = torch.arange(cfg.num_depth, dtype=torch.float32, device=vol.device).view(1, 1, -1, 1, 1)
            return mod(vol.reshape(vol.shape[0], 2 * groups, cfg.num_depth, hf, wf)) * 3.0 + 0.01 * ramp
        return run

    # oracle: the correlation volumes through the same net, then the reference chain from the cost volume on
    f_ref = scene["feature"].clone().requires_grad_(True)
    corr = O.scene_group_correlation(f_ref, scene["img_meta"], near_far_range=cfg.near_far_range,
                                     num_depth=cfg.num_depth, num_groups=groups, stride=cfg.stride)
    want_cost = cost_net_for(net)(corr)
    want = O.hot_path(f_ref, scene["img_meta"], lambda var: want_cost, near_far_range=cfg.near_far_range,
                      num_depth=cfg.num_depth, topk=cfg.topk, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                      stride=cfg.stride)
    g_vol = scene["g_volume_mean"]
    want_g, = torch.autograd.grad(want["volume_mean"], f_ref, g_vol)
    # CUDA
    dev = torch.device("cuda")
    net_d = torch.nn.Conv3d(2 * groups, 2, 3, padding=1).to(dev)
    net_d.load_state_dict(net.state_dict())
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        stride=cfg.stride, cost_volume="group_correlation", num_groups=groups)
    f_dev = scene["feature"].to(dev).requires_grad_(True)
    want_cost_d = want_cost.detach().to(dev)
    seen = {}

    def cost_net_cuda(vol):
        # cuDNN's conv differs from the CPU's by ~1e-6, enough to flip a near-tied third hypothesis: the
        # VALUES handed on are the oracle's (a + (b - a) == b exactly for a ~ b), the GRADIENT flows through
        # the CUDA net -- the test is about the hand-off, not about cuDNN
        cost = cost_net_for(net_d)(vol)
        seen["cost_err"] = float((cost.detach() - want_cost_d).abs().max())
        return cost + (want_cost_d - cost).detach()

    with torch.backends.cudnn.flags(allow_tf32=False):
        res = mod(f_dev, scene["img_meta"], cost_regularization=cost_net_cuda)
        got_g, = torch.autograd.grad(res["volume_mean"], f_dev, g_vol.to(dev))
    torch.cuda.synchronize()
    assert tuple(res["variance"].shape) == tuple(corr.shape)
    _close(res["variance"], corr, "correlation volumes")
    assert seen["cost_err"] <= 1e-4 * float(want_cost_d.abs().max()), seen
    assert torch.equal(res["est_idx"].cpu(), want["est_idx"]), "top-k indices"
    assert torch.equal(res["count"].cpu().reshape(-1), want["count"].reshape(-1).to(torch.int32))
    _close(res["volume_mean"], want["volume_mean"], "volume_mean")
    _close(got_g, want_g, "dL/dfeature through voxels and correlation volumes", rtol=1e-3, atol_scale=1e-3)
