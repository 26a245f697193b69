"""Host logic of the view-sharded scene forward (mvsdet_b200/sharded.py) on CPU:
partitioning, packing and a REAL world-size-2 all-reduce over gloo.  The
per-rank partial voxel sums come from the oracle (test infrastructure) -- on a
GPU box the same partials come from mvsd_backproject_fwd(MVSD_BP_SUM), which
tests/test_gpu_fullsize.py (test_view_sharded_partials_match_whole_scene,
test_sharded_forward_class_with_halo_packing) covers.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden, oracle_chain
from mvsdet_b200 import sharded
from mvsdet_b200.scene import SceneConfig, make_scene
from oracle import mvsdet_oracle as O


def test_partition_views_covers_everything_once():
    for v in (0, 1, 2, 5, 20, 80, 81, 100):
        for world in (1, 2, 3, 4, 8, 16):
            blocks = [sharded.partition_views(v, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == v
            for (b0, e0), (b1, e1) in zip(blocks, blocks[1:]):
                assert e0 == b1 and e0 >= b0 and e1 >= b1
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.partition_views(4, 2, 2)


def test_halo_views_rebases_neighbours():
    nbr = torch.tensor([[3, 5], [2, 6], [4, 9], [5, 0]])          # reference views 4..7 of a 10-view scene
    order, local = sharded.halo_views(nbr, 4, 8)
    assert order[:4] == [4, 5, 6, 7] and order[4:] == [0, 2, 3, 9]
    assert local.dtype == torch.int32 and tuple(local.shape) == (4, 2)
    for r in range(4):
        for j in range(2):
            assert order[int(local[r, j])] == int(nbr[r, j])
    order, local = sharded.halo_views(torch.zeros((3, 0), dtype=torch.int64), 0, 3)   # k = 0
    assert order == [0, 1, 2] and tuple(local.shape) == (3, 0)


def test_pack_unpack_roundtrip_both_memory_orders():
    g = torch.Generator().manual_seed(3)
    c, n = 8, 50
    vol = torch.randn(c, n, generator=g)
    cnt = torch.randint(0, 100, (n,), generator=g, dtype=torch.int32)
    buf = sharded.pack_partials(vol, cnt)
    assert buf.shape == (c * n + n,) and buf.dtype == torch.float32
    v2, c2 = sharded.unpack_partials(buf, c, n, channels_first=True)
    assert torch.equal(v2, vol) and torch.equal(c2, cnt)
    vol_t = torch.randn(n, c, generator=g).t()              # logical [C,N], memory [N,C]
    buf = sharded.pack_partials(vol_t, cnt)
    v3, c3 = sharded.unpack_partials(buf, c, n, channels_first=False)
    assert torch.equal(v3, vol_t) and torch.equal(c3, cnt)
    with pytest.raises(ValueError):
        sharded.pack_partials(vol, cnt[:-1])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partials_for(scene, full, begin, end):
    """Partial voxel sums / counts of reference views [begin, end) with the
    oracle's back-projection fed the whole-scene hypotheses."""
    cfg = scene["cfg"]
    h, w = cfg.crop_hw
    v = end - begin
    depth = full["est_depth"][begin:end, :, :h, :w].reshape(v, cfg.topk, -1).transpose(2, 1).unsqueeze(2)
    dens = full["est_densities"][begin:end, :, :h, :w].reshape(v, cfg.topk, -1).transpose(2, 1).unsqueeze(2)
    vol, valid = O.backproject_weigh(scene["feature"][begin:end, :, :h, :w], full["points"],
                                     full["projection"][begin:end], depth, cfg.voxel_size, dens)
    c = vol.shape[1]
    return vol.sum(0).reshape(c, -1), valid.sum(0).reshape(-1).to(torch.int32)


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        scene, _ = load_golden(case)
        full = oracle_chain(scene, with_grads=False)
        begin, end = sharded.partition_views(scene["cfg"].n_views, world, rank)
        vol_sum, count = _partials_for(scene, full, begin, end)
        c, n = vol_sum.shape
        buf = sharded.pack_partials(vol_sum, count)
        sharded.allreduce_partials(buf)                                 # the one collective
        tot, cnt = sharded.unpack_partials(buf, c, n)
        mean = torch.where(cnt.unsqueeze(0) == 0, torch.zeros_like(tot), tot / (cnt + 1e-8))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), mean=mean.numpy(), count=cnt.numpy(),
                 want_mean=full["volume_mean"].reshape(c, n).numpy(),
                 want_count=full["count"].reshape(-1).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["scannet_tiny", "arkit_tiny"])
def test_two_rank_gloo_allreduce_reproduces_whole_scene(case, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for r in res:
        assert np.array_equal(r["count"], r["want_count"]), "voxel counts must be bit-exact"
        np.testing.assert_allclose(r["mean"], r["want_mean"], rtol=1e-5, atol=1e-6)
    # replicas are bit-identical after the all-reduce + identical normalisation
    assert np.array_equal(res[0]["mean"], res[1]["mean"])
    assert np.array_equal(res[0]["count"], res[1]["count"])


def _ring_neighbours(v, k=2, seed=0):
    """pose-neighbour-like ids: mostly adjacent views, a few long links"""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for i in range(v):
        cand = [(i + o) % v for o in (1, -1, 2, -2, 5)]
        perm = torch.randperm(len(cand), generator=g).tolist()
        rows.append([cand[p] for p in perm[:k]])
    return torch.tensor(rows)


@pytest.mark.parametrize("v,world", [(8, 2), (20, 3), (20, 4), (80, 8), (5, 8)])
def test_halo_pull_table_reduces_to_the_whole_scene_gradient(v, world):
    """Backward of the view-sharded scene (host logic of ShardedScenePipeline.backward): every rank
    accumulates gradients for block + halo views; owners pull their peers' halo contributions
    (sharded.halo_pull_table feeds mvsd_halo_reduce_p2p).  Emulated with CPU tensors: the
    owner-side sums must equal the gradient of the unsharded scene."""
    nbr = _ring_neighbours(v, seed=v)
    e = 6
    g = torch.Generator().manual_seed(1)
    # contribution of reference view r to feature view n (itself and its neighbours)
    contrib = {}
    for r in range(v):
        for n_ in [r] + nbr[r].tolist():
            contrib[(r, n_)] = contrib.get((r, n_), 0) + torch.randn(e, generator=g, dtype=torch.float64)
    whole = torch.zeros(v, e, dtype=torch.float64)
    for (r, n_), val in contrib.items():
        whole[n_] += val
    # per-rank local accumulators over (block + halo) views
    local, orders = [], []
    for q in range(world):
        b, en = sharded.partition_views(v, world, q)
        if en <= b:
            local.append(torch.zeros(0, e, dtype=torch.float64)); orders.append([])
            continue
        order, _ = sharded.halo_views(nbr[b:en], b, en)
        buf = torch.zeros(len(order), e, dtype=torch.float64)
        for (r, n_), val in contrib.items():
            if b <= r < en:
                buf[order.index(n_)] += val
        local.append(buf); orders.append(order)
    seen = 0
    for q in range(world):
        b, en = sharded.partition_views(v, world, q)
        if en <= b:
            continue
        offs, src = sharded.halo_pull_table(nbr, world, q)
        assert offs.dtype == torch.int32 and tuple(offs.shape) == (en - b + 1,)
        mine = local[q][:en - b].clone()
        for d_ in range(en - b):
            for j in range(int(offs[d_]), int(offs[d_ + 1])):
                peer, idx = int(src[j, 0]), int(src[j, 1])
                assert peer != q and orders[peer][idx] == b + d_ and idx >= len(range(*sharded.partition_views(v, world, peer)))
                mine[d_] += local[peer][idx]
                seen += 1
        assert torch.allclose(mine, whole[b:en], rtol=0, atol=1e-12)
    total_halo = sum(max(0, len(o) - (sharded.partition_views(v, world, q)[1] - sharded.partition_views(v, world, q)[0]))
                     for q, o in enumerate(orders))
    assert seen == total_halo, "every halo view is pulled exactly once"


def _worker_halo(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        v, e = 9, 5
        nbr = _ring_neighbours(v, seed=4)
        b, en = sharded.partition_views(v, world, rank)
        order, _ = sharded.halo_views(nbr[b:en], b, en)
        g = torch.Generator().manual_seed(100 + rank)
        buf = torch.randn(len(order), e, generator=g, dtype=torch.float64)
        # exchange the local buffers (what the peer pointers give the CUDA kernel)
        bufs = [None] * world
        dist.all_gather_object(bufs, (order, buf))
        offs, src = sharded.halo_pull_table(nbr, world, rank)
        mine = buf[:en - b].clone()
        for d_ in range(en - b):
            for j in range(int(offs[d_]), int(offs[d_ + 1])):
                mine[d_] += bufs[int(src[j, 0])][1][int(src[j, 1])]
        want = torch.zeros(en - b, e, dtype=torch.float64)
        for q, (o, bq) in enumerate(bufs):
            for i, gview in enumerate(o):
                if b <= gview < en:
                    want[gview - b] += bq[i]
        np.savez(os.path.join(out_dir, f"halo{rank}.npz"), mine=mine.numpy(), want=want.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_halo_exchange(tmp_path):
    world = 2
    mp.spawn(_worker_halo, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        z = np.load(tmp_path / f"halo{r}.npz")
        np.testing.assert_allclose(z["mine"], z["want"], rtol=0, atol=1e-12)


def test_pose_order_is_a_permutation_that_never_adds_halos():
    """sharded.pose_order: pose-clustered blocks.  On the multi-turn helix of the synthetic
    scenes the pose neighbours sit one turn away, so contiguous index blocks need many halo maps."""
    from mvsdet_b200 import geometry as G
    for v, world in ((80, 8), (40, 8), (80, 4), (20, 2), (9, 2), (5, 8)):
        cfg = SceneConfig(n_views=v)
        scene = make_scene(cfg, seed=7, with_grads=False)
        geo = G.scene_geometry(scene["img_meta"], stride=cfg.stride, near_far_range=cfg.near_far_range,
                               num_depth=cfg.num_depth, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                               device="cpu", prologue="host")
        order = sharded.pose_order(scene["img_meta"]["lidar2img"]["extrinsic"], geo.neighbor_ids_host, world)
        assert sorted(order) == list(range(v))
        before = sharded.halo_counts(geo.neighbor_ids_host, world)
        after = sharded.halo_counts(geo.neighbor_ids_host, world, order)
        assert (max(after), sum(after)) <= (max(before), sum(before))
        if v == 80 and world == 8:
            assert max(before) >= 10 and max(after) <= 4, (before, after)      # 11-20 halo maps -> 1-3


def test_permuted_scene_has_permuted_neighbours():
    """kNN in pose space is permutation-equivariant: the ids computed from the re-ordered meta dict are the
    re-numbered ids of the original scene (what ShardedScenePipeline.load relies on)."""
    from mvsdet_b200 import geometry as G
    cfg = SceneConfig(n_views=24, per_view_intrinsics=True)
    scene = make_scene(cfg, seed=3, with_grads=False)
    kw = dict(stride=cfg.stride, near_far_range=cfg.near_far_range, num_depth=cfg.num_depth,
              n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, device="cpu", prologue="host")
    geo = G.scene_geometry(scene["img_meta"], **kw)
    order = sharded.pose_order(scene["img_meta"]["lidar2img"]["extrinsic"], geo.neighbor_ids_host, 4)
    meta_p = sharded.permute_img_meta(scene["img_meta"], order)
    assert len(meta_p["lidar2img"]["intrinsic"]) == 24 and meta_p["img_shape"] == scene["img_meta"]["img_shape"]
    geo_p = G.scene_geometry(meta_p, **kw)
    inv = {g: i for i, g in enumerate(order)}
    want = torch.tensor([[inv[int(x)] for x in geo.neighbor_ids_host[g].tolist()] for g in order])
    assert torch.equal(geo_p.neighbor_ids_host, want)
    assert torch.equal(geo_p.projection, geo.projection[order])
