"""View-sharded scene over 2 GPUs: the NVLink peer-memory combine
(mvsd_voxel_reduce_p2p) against the NCCL all-reduce path and the whole-scene
result.  Needs two GPUs; skipped on a single-GPU box."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from mvsdet_b200 import sharded
    from mvsdet_b200.hotpath import MVSDetHotPath
    from mvsdet_b200.scene import make_scene, tiny_config
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = tiny_config(n_views=7, channels=64)
        scene = make_scene(cfg, seed=3, with_grads=False)
        hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                            stride=cfg.stride)
        feat, cost = scene["feature"].to(dev), scene["cost_out"].to(dev)
        begin, end = sharded.partition_views(cfg.n_views, world, rank)
        outs = {}
        for name, p2p in (("nccl", False), ("p2p", True)):
            run = sharded.ShardedSceneForward(hot, p2p=p2p)
            for _ in range(3):                       # the buffers are reused scene after scene
                out = run(feat, scene["img_meta"], cost_regularization=lambda var: cost[begin:end])
            outs[name] = (out["volume_mean"].clone(), out["count"].clone())
        whole = hot(feat, scene["img_meta"], cost_regularization=lambda var: cost)
        torch.cuda.synchronize()
        ok_count = torch.equal(outs["p2p"][1], whole["count"]) and torch.equal(outs["nccl"][1], whole["count"])
        err = float((outs["p2p"][0] - whole["volume_mean"]).abs().max())
        scale = float(whole["volume_mean"].abs().max())
        # identical bits on both ranks
        chk = outs["p2p"][0].double().sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        ret[rank] = (bool(ok_count), err, scale, float(both[0]) == float(both[1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_p2p_combine_matches_nccl_and_whole_scene():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for rank in (0, 1):
        ok_count, err, scale, same = ret[rank]
        assert ok_count, "voxel counts differ from the whole-scene result"
        assert err <= 1e-5 * max(scale, 1.0), f"volume_mean differs by {err}"
        assert same, "replicas are not bit-identical"


def _worker_pipeline(rank, world, port, ret):
    """ShardedScenePipeline forward (p2p + nccl, latency and pipelined) and backward against the
    whole-scene forward + autograd of one GPU."""
    import torch.distributed as dist
    from mvsdet_b200 import sharded
    from mvsdet_b200.hotpath import MVSDetHotPath
    from mvsdet_b200.scene import make_scene, tiny_config
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = tiny_config(n_views=9, channels=64)
        scene = make_scene(cfg, seed=5)
        hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                            stride=cfg.stride, feature_dtype=torch.bfloat16)
        feat = scene["feature"].to(dev).requires_grad_(True)
        cost = scene["cost_out"].to(dev).requires_grad_(True)
        whole = hot(feat, scene["img_meta"], cost_regularization=lambda var: cost)
        g_vol = scene["g_volume_mean"].to(dev)
        g_var = scene["g_variance"].to(dev)
        torch.autograd.backward([whole["variance"], whole["volume_mean"]], [g_var, g_vol])
        pipe = sharded.ShardedScenePipeline(hot, cfg, dev)
        pipe.load(scene["feature"], scene["cost_out"], scene["img_meta"])
        own = pipe.view_ids.to(dev)                          # pose-clustered block: the views this rank owns
        ok = {}
        for mode in ("p2p", "nccl"):
            for _ in range(3):
                res = pipe.forward(mode)
            ok[mode + "_count"] = bool(torch.equal(res["count"], whole["count"]))
            ok[mode + "_err"] = float((res["volume_mean"] - whole["volume_mean"]).abs().max())
            res = pipe.forward_stream(mode, 5)
            torch.cuda.synchronize()
            ok[mode + "_stream_count"] = bool(torch.equal(res["count"], whole["count"]))
            ok[mode + "_stream_err"] = float((res["volume_mean"] - whole["volume_mean"]).abs().max())
        pipe.forward("p2p")
        for _ in range(2):                                   # twice: the accumulators are reused
            g_feat, g_cost = pipe.backward(g_vol, g_var[own].contiguous(memory_format=torch.channels_last_3d))
        torch.cuda.synchronize()
        ok["g_feat_err"] = float((g_feat - feat.grad[own]).abs().max())
        ok["g_feat_scale"] = float(feat.grad.abs().max())
        ok["g_cost_err"] = float((g_cost - cost.grad[own]).abs().max())
        ok["scale"] = float(whole["volume_mean"].abs().max())
        pipe.close()
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_pipeline_forward_backward_matches_whole_scene():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_pipeline, args=(2, _free_port(), ret), nprocs=2, join=True)
    for rank in (0, 1):
        ok = ret[rank]
        for mode in ("p2p", "nccl"):
            assert ok[mode + "_count"] and ok[mode + "_stream_count"], f"{mode}: voxel counts"
            assert ok[mode + "_err"] <= 1e-5 * max(ok["scale"], 1.0), ok
            assert ok[mode + "_stream_err"] <= 1e-5 * max(ok["scale"], 1.0), ok
        assert ok["g_feat_err"] <= 1e-4 * ok["g_feat_scale"], ok
        assert ok["g_cost_err"] <= 1e-4, ok
