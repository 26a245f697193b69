"""View-sharded forward of ONE scene over the ranks of a process group
(BASELINE.json configs[2]: test-time scenes with 50-100 source views).

Reference views are independent in the plane sweep and in the back-projection
(per-view loops, projects/NeRF-Det/nerfdet/mvsdet.py:453, :1401, :1458); only
the sum over views and the per-voxel valid count couple them (mvsdet.py:511-515).
So every rank sweeps a contiguous block of reference views, back-projects them
into *partial* voxel sums and counts (``MVSD_BP_SUM``), and the partials
(``C*N + N`` words, 26.3 MB for the shipped 40x40x16 grid) are combined once per
scene, giving replicas that are bit-identical across ranks.  Two combines exist:

* ``P2PVoxelReducer`` (default): the back-projection writes its partials into a
  peer-mapped buffer and ``mvsd_voxel_reduce_p2p`` does reduce-scatter +
  ``sum / (count + 1e-8)`` (mvsdet.py:514-515, :681-682) + all-gather in one kernel
  over NVLink peer pointers (V=80 on 8 B200: 0.38 ms per scene against 0.59 ms);
* ``p2p=False``: ONE NCCL all-reduce of the packed partials, then
  ``mvsd_voxel_normalize`` on every rank.  The counts travel as fp32 in the same
  buffer: they are integers <= V <= 2^24, so the sum is exact.

The FPN feature maps of all V views are resident on every rank (the k=2 pose
neighbours of a local view may belong to another rank, SURVEY.md 8e caveat 1);
only the reference-view work is partitioned: a rank packs (fp32 NCHW -> channels-last
bf16/fp32) just its own block plus the few neighbour views outside it ("halo"), and
the neighbour ids are re-based onto that compact buffer.

The partition / packing / all-reduce helpers are device-agnostic host logic and
are exercised with a world-size-2 gloo group on CPU (tests/test_sharded_cpu.py);
``ShardedSceneForward`` itself launches the CUDA kernels and needs a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import CHANNELS_FIRST, CHANNELS_LAST

__all__ = ["partition_views", "halo_views", "pack_partials", "unpack_partials",
           "allreduce_partials", "LocalGeometry", "ShardedSceneForward"]


def halo_views(neighbor_ids: torch.Tensor, begin: int, end: int) -> Tuple[List[int], torch.Tensor]:
    """Views a rank must hold to sweep reference views [begin, end): the block
    itself followed by the neighbour views outside it, ascending.  Returns
    (global view ids in local order, neighbour ids re-based to local indices
    [V_local,k] int32).  Pure host logic (tested on CPU)."""
    nbr = neighbor_ids.to(torch.int64).cpu()
    own = list(range(begin, end))
    extra = sorted(set(int(x) for x in nbr.reshape(-1).tolist()) - set(own))
    order = own + extra
    lut = {g: i for i, g in enumerate(order)}
    local = torch.tensor([[lut[int(x)] for x in row] for row in nbr.tolist()],
                         dtype=torch.int32).reshape(nbr.shape[0], nbr.shape[1] if nbr.dim() > 1 else 0)
    return order, local


@dataclass
class LocalGeometry:
    """A rank's slice of the scene parameter block plus its halo bookkeeping."""
    geo: object                    # SceneGeometry of reference views [begin, end)
    views: List[int]               # global view ids held locally (block first, then halo)
    neighbor_ids_local: torch.Tensor   # [V_local,k] int32 on the device, indices into ``views``
    begin: int
    end: int


def partition_views(n_views: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block [begin, end) of reference views for ``rank``
    (the first ``n_views % world_size`` ranks get one extra view; a rank may get
    an empty block when world_size > n_views)."""
    if n_views < 0 or world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad partition request: V={n_views}, world={world_size}, rank={rank}")
    base, extra = divmod(n_views, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pack_partials(volume_sum: torch.Tensor, count: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[C,N] fp32 sums (any memory order) + [N] integer counts -> one flat fp32
    buffer ``[C*N + N]`` in the memory order of ``volume_sum``."""
    c, n = volume_sum.shape
    if count.numel() != n:
        raise ValueError("count must have one entry per voxel")
    mem = volume_sum if volume_sum.is_contiguous() else volume_sum.t()
    if not mem.is_contiguous():
        raise ValueError("volume_sum must be [C,N] contiguous or the transpose of a contiguous [N,C]")
    if out is None:
        out = torch.empty(c * n + n, dtype=torch.float32, device=volume_sum.device)
    out[:c * n].copy_(mem.reshape(-1))
    out[c * n:].copy_(count.reshape(-1))       # int -> fp32, exact for counts <= 2^24
    return out


def unpack_partials(buf: torch.Tensor, c: int, n: int, channels_first: bool = True):
    """Inverse of pack_partials -> (volume_sum logical [C,N] view, count int32 [N])."""
    vol = buf[:c * n].view(c, n) if channels_first else buf[:c * n].view(n, c).t()
    count = buf[c * n:].round().to(torch.int32)
    return vol, count


def allreduce_partials(buf: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the packed partials over the group, in place.  One collective per
    scene; a no-op without an initialised process group (single-rank run)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class P2PVoxelReducer:
    """Sum-over-ranks + normalise of the voxel partials over NVLink peer memory
    (``mvsd_voxel_reduce_p2p``): one reduce-scatter / normalise / all-gather kernel
    bracketed by two cross-rank barriers, instead of an NCCL all-reduce followed by a
    normalise kernel.  Buffers are torch symmetric memory (P2P-mapped on every rank of
    the group); torch provides the allocation, the pointer exchange and the barrier
    kernel, the data path is ours.

        red = P2PVoxelReducer(C, N, channels_first, device, group)
        volume_mean, count = red(volume_sum, count)        # identical bits on every rank
    """

    def __init__(self, channels: int, n_voxels: int, channels_first: bool, device, group=None):
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("P2PVoxelReducer needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.c, self.n, self.cfirst = int(channels), int(n_voxels), bool(channels_first)
        self.total = self.c * self.n
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.part = symm.empty(self.total + self.n, dtype=torch.float32, device=device)
        self.out = symm.empty(self.total + self.n, dtype=torch.float32, device=device)
        self.h_part = symm.rendezvous(self.part, self.group)
        self.h_out = symm.rendezvous(self.out, self.group)
        self.count_local = torch.empty(self.n, dtype=torch.int32, device=device)

    def partial_buffers(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(sum, count) views of this rank's peer-mapped partial buffer, in the volume's
        memory order: pass them as ``out=`` / ``count_out=`` of the back-projection so the
        kernel writes its partials where the peers will read them."""
        shape = (self.c, self.n) if self.cfirst else (self.n, self.c)
        return self.part[:self.total].view(shape), self.part[self.total:].view(torch.int32)

    def __call__(self, volume_sum: torch.Tensor, count: torch.Tensor):
        from . import _lib
        mem = volume_sum if volume_sum.is_contiguous() else volume_sum.t()
        if not mem.is_contiguous() or mem.numel() != self.total:
            raise ValueError("volume_sum must be [C,N] contiguous or the transpose of a contiguous [N,C]")
        if mem.data_ptr() != self.part.data_ptr():           # not already written in place
            self.part[:self.total].copy_(mem.reshape(-1))
            self.part[self.total:].view(torch.int32).copy_(count.reshape(-1))
        self.h_part.barrier(channel=0)          # every rank's partials are complete and visible
        _lib.call("mvsd_voxel_reduce_p2p", self.h_part.buffer_ptrs_dev, self.h_out.buffer_ptrs_dev,
                  self.count_local.data_ptr(), self.world, self.rank,
                  CHANNELS_FIRST if self.cfirst else CHANNELS_LAST, self.c, self.n,
                  torch.cuda.current_stream().cuda_stream)
        self.h_out.barrier(channel=0)           # every rank's slice has landed in every out buffer
        vol = self.out[:self.total]
        vol = vol.view(self.c, self.n) if self.cfirst else vol.view(self.n, self.c).t()
        return vol, self.out[self.total:].view(torch.int32)


class ShardedSceneForward:
    """Forward of one scene with the reference views split over the group.

        sharded = ShardedSceneForward(hot_path)            # an MVSDetHotPath
        out = sharded(feature, img_meta, cost_regularization)
        out["volume_mean"], out["count"]                   # identical on every rank

    ``feature`` holds ALL views on every rank.  The all-reduce runs on the
    current stream right after the local back-projection (there is nothing left
    to overlap it with inside one scene; a multi-scene caller overlaps it with
    the next scene's sweep by calling from a side stream)."""

    def __init__(self, hot_path, group=None, p2p: bool = True):
        self.hot = hot_path
        self.group = group
        self.p2p = p2p                 # combine over NVLink peer memory (default); False = NCCL all-reduce
        self._reducer = None

    def _world(self) -> Tuple[int, int]:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def local_geometry(self, img_meta: dict, n_views: int, device) -> Optional[LocalGeometry]:
        """This rank's slice of the per-scene parameter block and its halo (host
        work, ~2 ms; a caller that revisits a scene passes it back in as
        ``geometry``)."""
        rank, world = self._world()
        begin, end = partition_views(n_views, world, rank)
        if end <= begin:
            return None
        geo = self.hot.geometry(img_meta, device, view_slice=slice(begin, end), prologue="host")
        views, nbr_local = halo_views(geo.neighbor_ids_host, begin, end)
        return LocalGeometry(geo=geo, views=views, neighbor_ids_local=nbr_local.to(device),
                             begin=begin, end=end)

    def _pack_local(self, feature: torch.Tensor, lg: LocalGeometry) -> torch.Tensor:
        """fp32 NCHW [V,C,H,W] -> channels-last buffer of the views in ``lg.views``
        (contiguous runs of views go through one transpose launch each)."""
        from . import _lib, ops
        v, c, h, w = feature.shape
        dt = self.hot.feature_dtype
        if feature.dtype != torch.float32 or not feature.is_contiguous():
            return ops.pack_features(feature[lg.views], dt)       # already channels-last / bf16 inputs
        out = torch.empty((len(lg.views), h, w, c), dtype=dt, device=feature.device)
        code = _lib.BF16 if dt == torch.bfloat16 else _lib.F32
        st = torch.cuda.current_stream().cuda_stream
        i = 0
        while i < len(lg.views):
            j = i
            while j + 1 < len(lg.views) and lg.views[j + 1] == lg.views[j] + 1:
                j += 1
            _lib.call("mvsd_pack_nchw_to_nhwc", feature[lg.views[i]].data_ptr(), out[i].data_ptr(), code,
                      j - i + 1, c, h, w, st)
            i = j + 1
        return out.permute(0, 3, 1, 2)

    def local_partials(self, feature: torch.Tensor, img_meta: dict, cost_net: Callable,
                       geometry: Optional[LocalGeometry] = None, rank_world: Optional[Tuple[int, int]] = None,
                       partial_out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """This rank's contribution before the collective: (volume_sum logical
        [C,N], count int32 [N], (begin, end)).  ``rank_world`` overrides the
        process group (used by the single-GPU tests to play every rank)."""
        from . import ops
        hot = self.hot
        rank, world = rank_world or self._world()
        v_all = feature.shape[0]
        begin, end = partition_views(v_all, world, rank)
        c = feature.shape[1]
        nx, ny, nz = hot.n_voxels
        n = nx * ny * nz
        dev = feature.device
        if end > begin:
            lg = geometry
            if lg is None:
                geo = hot.geometry(img_meta, dev, view_slice=slice(begin, end), prologue="host")
                views, nbr_local = halo_views(geo.neighbor_ids_host, begin, end)
                lg = LocalGeometry(geo, views, nbr_local.to(dev), begin, end)
            geo = lg.geo
            feat_cl = self._pack_local(feature, lg)            # block + halo only
            variance = ops.plane_sweep_variance(feat_cl, lg.neighbor_ids_local, geo.hom,
                                                geo.depth_values, out_dtype=hot.variance_dtype,
                                                ref_begin=0)
            cost_out = cost_net(variance)
            _, _, est_depth, est_dens, est_idx, _ = hot.hypotheses(cost_out)
            vol_sum, count = ops.backproject_aggregate(
                feat_cl[:end - begin], geo.points, geo.projection, est_depth, est_dens,
                hot.voxel_size[2], geo.height, geo.width, mode="sum",
                channels_first=hot.channels_first_volume,
                out=None if partial_out is None else partial_out[0],
                count_out=None if partial_out is None else partial_out[1])
        else:                                   # more ranks than views: contribute zeros
            shape = (c, n) if hot.channels_first_volume else (n, c)
            vol_sum = torch.zeros(shape, dtype=torch.float32, device=dev)
            vol_sum = vol_sum if hot.channels_first_volume else vol_sum.t()
            count = torch.zeros(n, dtype=torch.int32, device=dev)
        return vol_sum, count, (begin, end)

    def __call__(self, feature: torch.Tensor, img_meta: dict,
                 cost_regularization: Optional[Callable] = None, geometry=None) -> Dict[str, torch.Tensor]:
        from . import ops
        hot = self.hot
        cost_net = cost_regularization or hot.cost_regularization
        if cost_net is None:
            raise ValueError("a cost_regularization callable is required (mvsdet.py:470)")
        c = feature.shape[1]
        nx, ny, nz = hot.n_voxels
        n = nx * ny * nz
        use_p2p = self.p2p and self._world()[1] > 1
        if use_p2p and self._reducer is None:
            self._reducer = P2PVoxelReducer(c, n, hot.channels_first_volume, feature.device, self.group)
        vol_sum, count, (begin, end) = self.local_partials(
            feature, img_meta, cost_net, geometry,
            partial_out=self._reducer.partial_buffers() if use_p2p else None)
        if use_p2p:
            volume_mean, count = self._reducer(vol_sum.detach(), count)
        else:
            buf = pack_partials(vol_sum.detach(), count)
            allreduce_partials(buf, self.group)
            vol_sum, count = unpack_partials(buf, c, n, hot.channels_first_volume)
            volume_mean = ops.voxel_normalize(vol_sum, count)
        vm = (volume_mean.view(c, nx, ny, nz) if volume_mean.is_contiguous()
              else volume_mean.unflatten(1, (nx, ny, nz)))
        return dict(volume_mean=vm, valid=count.view(1, nx, ny, nz).float(), count=count,
                    view_range=(begin, end))
