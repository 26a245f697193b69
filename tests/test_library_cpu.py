"""torch.library registration of the path's operators (mvsdet_b200/library.py,
SURVEY.md 8b): schemas, fake (meta) shape / stride / dtype inference, autograd wiring
through the *_bwd ops, and the absence of any CPU kernel.  No compute: runs without a GPU."""
import pytest
import torch

from mvsdet_b200 import library as L

V, C, H, W, D, K, T = 4, 16, 8, 12, 12, 2, 3
NVOX = (5, 4, 3)


def _meta_scene(device="meta", feat_dtype=torch.float32):
    feat = torch.empty((V, H, W, C), dtype=feat_dtype, device=device).permute(0, 3, 1, 2)
    return dict(
        feat=feat.requires_grad_(True),
        nbr=torch.empty((V, K), dtype=torch.int32, device=device),
        hom=torch.empty((V, K, 12), device=device),
        dv=torch.empty((V, D), device=device),
        cost_out=torch.empty((V, 2, D, H, W), device=device, requires_grad=True),
        points=torch.empty((3,) + NVOX, device=device),
        projection=torch.empty((V, 3, 4), device=device))


def test_ops_are_registered_with_the_dispatcher():
    ns = torch.ops.mvsdet_b200
    for name in ("plane_sweep_variance", "plane_sweep_variance_bwd", "depth_topk", "depth_topk_bwd",
                 "backproject_aggregate", "backproject_aggregate_bwd", "voxel_normalize"):
        schema = str(getattr(ns, name).default._schema)
        assert schema.startswith(f"mvsdet_b200::{name}("), schema
    s = str(ns.depth_topk.default._schema)
    assert "-> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)" in s
    # functional ops: nothing is mutated, nothing aliases an input
    for name in ("plane_sweep_variance", "depth_topk", "backproject_aggregate"):
        assert not getattr(ns, name).default._schema.is_mutable


@pytest.mark.parametrize("out_bf16", [False, True])
def test_plane_sweep_fake_layout(out_bf16):
    s = _meta_scene()
    var = L.plane_sweep_variance(s["feat"], s["nbr"], s["hom"], s["dv"], out_bf16)
    assert tuple(var.shape) == (V, C, D, H, W)                       # reference's logical shape
    assert var.permute(0, 2, 3, 4, 1).is_contiguous()                # channels_last_3d memory
    assert var.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    g, = torch.autograd.grad(var, s["feat"], torch.empty_like(var))
    assert tuple(g.shape) == (V, C, H, W) and g.permute(0, 2, 3, 1).is_contiguous()


def test_group_correlation_fake_layout():
    """optional operator: logical [V,k,G,D,H,W], groups innermost in memory"""
    s = _meta_scene()
    cost = L.plane_sweep_group_correlation(s["feat"], s["nbr"], s["hom"], s["dv"], 4)
    assert tuple(cost.shape) == (V, K, 4, D, H, W) and cost.dtype == torch.float32
    assert cost.permute(0, 1, 3, 4, 5, 2).is_contiguous()
    g, = torch.autograd.grad(cost, s["feat"], torch.empty_like(cost))
    assert tuple(g.shape) == (V, C, H, W) and g.permute(0, 2, 3, 1).is_contiguous()
    schema = str(torch.ops.mvsdet_b200.plane_sweep_group_correlation.default._schema)
    assert schema.startswith("mvsdet_b200::plane_sweep_group_correlation(") and "num_groups=8" in schema


def test_view_slice_fake_layout():
    """view-sharded form: V_local reference views out of Vf packed views"""
    s = _meta_scene()
    var = L.plane_sweep_variance(s["feat"], s["nbr"][:2].contiguous(), s["hom"][:2], s["dv"][:2], False, 1)
    assert tuple(var.shape) == (2, C, D, H, W)


def test_chain_fake_shapes_and_gradients():
    s = _meta_scene()
    prob, off, est_depth, est_dens, est_idx, coding = L.depth_topk(s["cost_out"], 0.2, 0.4, T)
    assert tuple(prob.shape) == tuple(off.shape) == (V, D, H, W)
    assert tuple(est_depth.shape) == tuple(est_dens.shape) == tuple(est_idx.shape) == (V, T, H, W)
    assert est_idx.dtype == torch.int64 and not est_idx.requires_grad
    assert tuple(coding.shape) == (V, H, W)
    n = NVOX[0] * NVOX[1] * NVOX[2]
    for channels_first in (True, False):
        for sum_only in (False, True):
            vol, count = L.backproject_aggregate(s["feat"], s["points"], s["projection"], est_depth,
                                                 est_dens, 0.2, H - 1, W, sum_only, channels_first)
            assert tuple(vol.shape) == (C, n) and count.dtype == torch.int32 and tuple(count.shape) == (n,)
            assert (vol if channels_first else vol.t()).is_contiguous()
            assert not count.requires_grad
            gf, gc = torch.autograd.grad(vol, (s["feat"], s["cost_out"]), torch.empty_like(vol),
                                         retain_graph=True)
            assert tuple(gf.shape) == (V, C, H, W) and tuple(gc.shape) == (V, 2, D, H, W)
    out = L.voxel_normalize(vol.detach(), count)
    assert tuple(out.shape) == (C, n)


def test_traces_under_fake_tensor_mode():
    """what a compiler front end does with the surrounding detector: FakeTensorMode with
    'cuda' fake tensors never reaches the library"""
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        s = _meta_scene(device="cuda", feat_dtype=torch.bfloat16)
        var = L.plane_sweep_variance(s["feat"], s["nbr"], s["hom"], s["dv"])
        assert var.device.type == "cuda" and tuple(var.shape) == (V, C, D, H, W)
        outs = L.depth_topk(s["cost_out"], 0.2, 0.4, T)
        vol, count = L.backproject_aggregate(s["feat"], s["points"], s["projection"], outs[2], outs[3],
                                             0.2, H - 1, W)
        assert vol.device.type == "cuda" and vol.grad_fn is not None and count.dtype == torch.int32
        if torch.cuda.is_available():      # the autograd engine wants a device thread for 'cuda' fakes
            gf, = torch.autograd.grad(vol, s["feat"], torch.empty_like(vol))
            assert gf.dtype == torch.bfloat16 and tuple(gf.shape) == (V, C, H, W)


def test_there_is_no_cpu_kernel():
    s = _meta_scene(device="cpu")
    with pytest.raises(NotImplementedError):
        L.plane_sweep_variance(s["feat"].detach(), s["nbr"], s["hom"], s["dv"])
    with pytest.raises(NotImplementedError):
        L.depth_topk(s["cost_out"].detach(), 0.2, 0.4, T)
    with pytest.raises(NotImplementedError):
        L.voxel_normalize(torch.zeros(C, 6), torch.zeros(6, dtype=torch.int32))
