"""Per-scene camera geometry of the MVSDet hot path.

The reference receives the camera matrices as numpy arrays inside ``img_meta``
and rebuilds small tensors from them every iteration with ~35 tiny ATen calls
and a Python loop over the views
(projects/NeRF-Det/nerfdet/mvsdet.py:407-434, :448-450, :1124-1156, :1316-1327).

``scene_geometry`` has two prologues producing the same parameter block:

* ``prologue="device"`` (default on CUDA): the host computes only what has to be
  ATen's bits -- ``ref_proj = K_feat @ w2c`` and ``inverse(ref_proj)``, because the
  variance volume is sensitive to the rounding of that fp32 inverse
  -- with a handful of numpy ops and ONE ``torch.inverse``; one pinned upload, then
  ONE kernel (``mvsd_scene_setup``, csrc/scene_setup.cu) derives neighbour ids,
  homographies and voxel projections on the device;
* ``prologue="host"``: everything on the host in fp32 with the reference's own ATen
  ops (``torch.inverse``, ``matmul``, ``topk``), packed into one pinned upload.  Used
  where the host needs the neighbour ids (view-sharded halo packing) and as the
  cross-check of the device prologue (tests/test_gpu_geometry.py: bit-identical).

Same names and argument meaning as the reference for the mirrored functions.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

__all__ = ["knn", "get_nearest_pose_ids", "collect_proj", "homography_params",
           "compute_projection", "get_points", "feature_intrinsics", "depth_values_for",
           "SceneGeometry", "scene_geometry", "host_camera_block"]


def knn(x: torch.Tensor, ref: torch.Tensor, k: int, maskself: bool = False) -> torch.Tensor:
    """Reference: mvsdet.py:43-64.  x, ref [B,3,N] -> [B,N,k] indices of the k
    nearest points of ``ref`` (largest negated squared distance first)."""
    neg_d2 = -(ref ** 2).sum(1, keepdim=True) + 2 * torch.matmul(x.transpose(2, 1), ref)
    neg_d2 = neg_d2 - (x ** 2).sum(1, keepdim=True).transpose(2, 1)
    if maskself:
        if x.shape != ref.shape:
            raise ValueError("maskself needs x and ref of the same shape")
        eye = torch.arange(x.shape[2])
        neg_d2[:, eye, eye] = -100000
    return neg_d2.topk(k=k, dim=-1)[1]


def get_nearest_pose_ids(tar_pose: torch.Tensor, ref_poses: torch.Tensor, num_select: int,
                         maskself: bool = False, angular_dist_method: str = "dist") -> torch.Tensor:
    """Reference: mvsdet.py:67-104 (only the live ``'dist'`` method).
    camera-to-world poses [V,4,4] -> neighbour ids [V,k], nearest first."""
    if angular_dist_method != "dist":
        raise ValueError("only angular_dist_method='dist' exists on the MVSDet path")
    num_select = min(num_select, len(ref_poses) - 1)
    tar = tar_pose[:, :3, 3].unsqueeze(0).transpose(2, 1)
    ref = ref_poses[:, :3, 3].unsqueeze(0).transpose(2, 1)
    return knn(tar, ref, k=num_select, maskself=maskself)[0]


def feature_intrinsics(intrinsic: torch.Tensor, ratio: float) -> torch.Tensor:
    """mvsdet.py:422-428: focal lengths / principal point at feature resolution."""
    k = intrinsic.clone()
    if k.dim() == 2:
        k[:2] /= ratio
    else:
        k[:, :2] /= ratio
    return k


def collect_proj(w2c: torch.Tensor, intr: torch.Tensor, neighbor_ids: torch.Tensor):
    """Reference: MVSDet.collect_proj, mvsdet.py:249-264."""
    if intr.dim() == 2:
        intr = intr.unsqueeze(0).expand(w2c.shape[0], 4, 4)
    proj = torch.matmul(intr, w2c)
    v, k = neighbor_ids.shape
    nei = proj[neighbor_ids.reshape(-1)].view(v, k, 4, 4)
    return proj, torch.unbind(nei, dim=1)


def homography_params(src_proj: torch.Tensor, ref_proj: torch.Tensor) -> torch.Tensor:
    """[B,4,4] x2 -> [B,12]: rot rows then trans of src_proj @ inverse(ref_proj)
    (mvs_models/module.py:116-118), the parameter block the warp kernels read."""
    m = torch.matmul(src_proj, torch.inverse(ref_proj))
    return torch.cat((m[:, :3, :3].reshape(-1, 9), m[:, :3, 3]), dim=1).contiguous()


def compute_projection(img_meta: dict, stride: int) -> torch.Tensor:
    """Reference: MVSDet._compute_projection, mvsdet.py:1124-1156 (angles=None).
    -> [V,3,4] fp32, K_feat[:3,:3] @ w2c[:3] per view."""
    intr = img_meta["lidar2img"]["intrinsic"]
    per_view = isinstance(intr, (list, tuple))
    intr_t = torch.as_tensor(np.array(intr), dtype=torch.float32)
    extr = torch.as_tensor(np.array(img_meta["lidar2img"]["extrinsic"]), dtype=torch.float32)
    ratio = img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride)
    out = []
    for i in range(extr.shape[0]):
        k = (intr_t[i] if per_view else intr_t)[:3, :3].clone()
        k[:2] /= ratio
        out.append(k @ extr[i, :3])
    return torch.stack(out)


def projection_vectorized(intr_t: torch.Tensor, extr: torch.Tensor, ratio: float) -> torch.Tensor:
    """Same values as :func:`compute_projection`, bit for bit, without the Python
    loop over views (0.7 ms -> 0.05 ms at V=20).  ATen's CPU ``mm`` for a 3x3 @ 3x4
    product accumulates each output as an FMA chain in k order (checked against the
    loop on 120 seeded scenes, tests/test_geometry_cpu.py); the batched ``bmm`` does
    NOT round the same way, so the chain is spelled out: the products of two fp32
    values are exact in fp64, one rounding per step."""
    v = extr.shape[0]
    k3 = (intr_t[:, :3, :3] if intr_t.dim() == 3 else intr_t[:3, :3].unsqueeze(0).expand(v, 3, 3)).clone()
    k3[:, :2] /= ratio
    a = k3.numpy().astype(np.float64)            # [V,3,3]
    b = extr[:, :3].numpy().astype(np.float64)   # [V,3,4]
    acc = (a[:, :, 0:1] * b[:, 0:1, :]).astype(np.float32)
    for kk in (1, 2):
        acc = (a[:, :, kk:kk + 1] * b[:, kk:kk + 1, :] + acc.astype(np.float64)).astype(np.float32)
    return torch.from_numpy(acc)


_STATIC_CACHE: dict = {}


def static_geometry(n_voxels, voxel_size, origin, near_far_range, num_depth: int, n_views: int, device):
    """Voxel centres and depth planes depend only on the detector configuration
    (mvsdet.py:222-225, :1316-1327): built once per (config, device) and kept on
    the device instead of being rebuilt and re-uploaded (307 KB) for every scene."""
    key = (tuple(int(n) for n in n_voxels), tuple(float(x) for x in voxel_size),
           tuple(float(x) for x in np.asarray(origin).reshape(-1)),
           tuple(float(x) for x in near_far_range), int(num_depth), int(n_views), str(device))
    hit = _STATIC_CACHE.get(key)
    if hit is None:
        points = get_points(n_voxels, voxel_size, origin)
        dvals = torch.as_tensor(depth_values_for(near_far_range, num_depth)).unsqueeze(0).repeat(n_views, 1)
        hit = (points.to(device), dvals.to(device))
        if len(_STATIC_CACHE) > 64:
            _STATIC_CACHE.clear()
        _STATIC_CACHE[key] = hit
    return hit


def get_points(n_voxels, voxel_size, origin) -> torch.Tensor:
    """Reference: get_points, mvsdet.py:1316-1327 -> [3,nx,ny,nz] fp32."""
    n_voxels = torch.as_tensor(n_voxels)
    voxel_size = torch.as_tensor(voxel_size, dtype=torch.float32)
    origin = torch.as_tensor(origin, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(*[torch.arange(int(n)) for n in n_voxels], indexing="ij"))
    new_origin = origin - n_voxels / 2. * voxel_size
    return grid * voxel_size.view(3, 1, 1, 1) + new_origin.view(3, 1, 1, 1)


def depth_values_for(near_far_range: Sequence[float], num_depth: int) -> np.ndarray:
    """mvsdet.py:222-225."""
    interval = (near_far_range[1] - near_far_range[0]) / num_depth
    dv = np.arange(near_far_range[0], near_far_range[1], interval, dtype=np.float32)
    if len(dv) != num_depth:
        raise ValueError(f"near_far_range {near_far_range} / {num_depth} planes does not "
                         f"yield {num_depth} depth values (reference assert, mvsdet.py:225)")
    return dv


def host_camera_block(img_meta: dict, stride: int):
    """The host side of the device prologue: numpy fp32 (w2c [V,4,4], k_feat [4,4] | [V,4,4],
    ref_proj [V,4,4], inv_ref [V,4,4]).

    ``k_feat`` rows 0-1 are the image intrinsics divided by ``ratio`` in fp32 (what
    ``tensor[:2] /= python_float`` does, mvsdet.py:422-428).  ``ref_proj = K_feat @ w2c`` is ATen's
    batched small-matrix product: ``((a0*b0 + a1*b1) + a2*b2) + a3*b3`` with every product and sum
    rounded to fp32 (no FMA) -- reproduced with numpy fp32 ops, bit for bit
    (tests/test_geometry_cpu.py); its inverse is ATen's own ``torch.inverse`` (LAPACK)."""
    w2c = np.asarray(img_meta["lidar2img"]["extrinsic"], dtype=np.float32)
    if w2c.ndim != 3 or w2c.shape[1:] != (4, 4):
        raise ValueError("img_meta['lidar2img']['extrinsic'] must be V matrices of 4x4")
    intr = np.asarray(img_meta["lidar2img"]["intrinsic"], dtype=np.float32)
    if intr.shape not in ((4, 4), (w2c.shape[0], 4, 4)):
        raise ValueError("img_meta['lidar2img']['intrinsic'] must be 4x4 or V matrices of 4x4")
    ratio = np.float32(img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride))
    k_feat = intr.copy()
    k_feat[..., :2, :] /= ratio
    a = k_feat if k_feat.ndim == 3 else k_feat[None]
    ref_proj = a[:, :, 0:1] * w2c[:, 0:1, :]
    for kk in (1, 2, 3):
        ref_proj = ref_proj + a[:, :, kk:kk + 1] * w2c[:, kk:kk + 1, :]
    inv_ref = torch.inverse(torch.from_numpy(ref_proj)).numpy()
    return w2c, k_feat, ref_proj, inv_ref


@dataclass
class SceneGeometry:
    """Device-resident parameter block of one scene."""
    neighbor_ids: torch.Tensor     # [V,k] int32
    hom: torch.Tensor              # [V,k,12] fp32
    depth_values: torch.Tensor     # [V,D] fp32
    projection: torch.Tensor       # [V,3,4] fp32
    points: torch.Tensor           # [3,nx,ny,nz] fp32
    neighbor_ids_host: Optional[torch.Tensor]  # [V,k] int64 (reference dtype); None after the device prologue
    height: int                    # un-padded feature rows (img_shape[0] // stride)
    width: int
    k: int
    k_feat: Optional[torch.Tensor] = None      # [4,4] | [V,4,4] feature-level intrinsics on the device
    n_views: int = 0               # views of the scene (the ids in neighbor_ids index these)

    def neighbor_ids_ref(self) -> torch.Tensor:
        """[V,k] int64 on the host, the reference's dtype (synchronises after the device prologue)."""
        if self.neighbor_ids_host is None:
            self.neighbor_ids_host = self.neighbor_ids.cpu().to(torch.int64)
        return self.neighbor_ids_host


def _scene_geometry_device(img_meta: dict, *, stride: int, near_far_range, num_depth: int, n_voxels,
                           voxel_size, num_neighbors: int, dev: torch.device,
                           view_slice: Optional[slice]) -> SceneGeometry:
    from . import ops
    w2c, k_feat, ref_proj, inv_ref = host_camera_block(img_meta, stride)
    v_all = w2c.shape[0]
    k = min(num_neighbors, v_all - 1)
    begin, end, step = (view_slice or slice(None)).indices(v_all)
    if step != 1 or end <= begin:
        raise ValueError("view_slice must be a non-empty contiguous range of reference views")
    sizes = (w2c.size, k_feat.size, ref_proj.size, inv_ref.size)
    staging = torch.empty(sum(sizes), dtype=torch.float32, pin_memory=True)
    host = staging.numpy()
    o = np.cumsum((0,) + sizes)
    for i, arr in enumerate((w2c, k_feat, ref_proj, inv_ref)):
        host[o[i]:o[i + 1]] = arr.reshape(-1)
    blob = staging.to(dev, non_blocking=True)
    d_w2c = blob[o[0]:o[1]].view(v_all, 4, 4)
    d_k = blob[o[1]:o[2]].view(k_feat.shape)
    nbr, hom, projection = ops.scene_setup(d_w2c, d_k, blob[o[2]:o[3]].view(v_all, 4, 4),
                                           blob[o[3]:o[4]].view(v_all, 4, 4), k, begin, end - begin)
    points, dplanes = static_geometry(n_voxels, voxel_size, img_meta["lidar2img"]["origin"],
                                      near_far_range, num_depth, end - begin, dev)
    return SceneGeometry(neighbor_ids=nbr, hom=hom, depth_values=dplanes, projection=projection,
                         points=points, neighbor_ids_host=None,
                         height=img_meta["img_shape"][0] // stride,
                         width=img_meta["img_shape"][1] // stride, k=k,
                         k_feat=d_k if d_k.dim() == 2 else d_k[begin:end], n_views=v_all)


def scene_geometry(img_meta: dict, *, stride: int, near_far_range, num_depth: int, n_voxels,
                   voxel_size, num_neighbors: int = 2, device="cuda",
                   view_slice: Optional[slice] = None, prologue: Optional[str] = None) -> SceneGeometry:
    """Everything mvsdet.py:407-450 derives from ``img_meta`` as one device-resident parameter
    block.  ``view_slice`` restricts the *reference* views (rows of every per-view array) for
    view-sharded multi-GPU runs; neighbour ids keep indexing the full feature tensor.
    ``prologue``: "device" (default on CUDA) or "host", see the module docstring."""
    if prologue is None:
        prologue = "device" if torch.device(device).type == "cuda" else "host"
    if prologue not in ("device", "host"):
        raise ValueError("prologue must be 'device' or 'host'")
    if prologue == "device":
        if torch.device(device).type != "cuda":
            raise ValueError("the device prologue needs a CUDA device: mvsdet_b200 has no CPU kernels")
        return _scene_geometry_device(img_meta, stride=stride, near_far_range=near_far_range,
                                      num_depth=num_depth, n_voxels=n_voxels, voxel_size=voxel_size,
                                      num_neighbors=num_neighbors, dev=torch.device(device),
                                      view_slice=view_slice)
    extr = img_meta["lidar2img"]["extrinsic"]
    w2c = torch.as_tensor(np.array(extr), dtype=torch.float32)
    v_all = w2c.shape[0]
    intr = torch.as_tensor(np.array(img_meta["lidar2img"]["intrinsic"]), dtype=torch.float32)
    ratio = img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride)
    k_feat = feature_intrinsics(intr, ratio)
    k = min(num_neighbors, v_all - 1)
    c2w = w2c.inverse()
    nbr = get_nearest_pose_ids(c2w, c2w, k, maskself=True)            # [V,k] int64
    ref_proj, nei_projs = collect_proj(w2c, k_feat, nbr)
    if k > 0:
        # src_proj @ inverse(ref_proj) (module.py:116-118) per neighbour; the inverse of the
        # reference projections is the same tensor for every neighbour, so it is taken once
        inv_ref = torch.inverse(ref_proj)
        m = torch.stack([torch.matmul(np_, inv_ref) for np_ in nei_projs], dim=1)     # [V,k,4,4]
        hom = torch.cat((m[:, :, :3, :3].reshape(v_all, k, 9), m[:, :, :3, 3]), dim=2).contiguous()
    else:
        hom = torch.zeros(v_all, 0, 12)
    projection = projection_vectorized(intr, w2c, ratio)
    dev = torch.device(device)
    if view_slice is not None:
        nbr, hom, projection = nbr[view_slice], hom[view_slice], projection[view_slice]
    v = nbr.shape[0]
    points, dplanes = static_geometry(n_voxels, voxel_size, img_meta["lidar2img"]["origin"],
                                      near_far_range, num_depth, v, dev)

    k_rows = k_feat if k_feat.dim() == 2 else (k_feat[view_slice] if view_slice is not None else k_feat)
    parts = [nbr.to(torch.int32).reshape(-1).view(torch.float32) if nbr.numel() else torch.zeros(0),
             hom.reshape(-1), projection.reshape(-1), k_rows.reshape(-1)]
    sizes = [p.numel() for p in parts]
    if dev.type == "cuda":
        staging = torch.empty(sum(sizes), dtype=torch.float32, pin_memory=True)
        torch.cat(parts, out=staging)
        blob = staging.to(dev, non_blocking=True)
    else:
        blob = torch.cat(parts)
    o = np.cumsum([0] + sizes)
    return SceneGeometry(
        neighbor_ids=blob[o[0]:o[1]].view(torch.int32).view(v, k),
        hom=blob[o[1]:o[2]].view(v, k, 12),
        depth_values=dplanes,
        projection=blob[o[2]:o[3]].view(v, 3, 4),
        points=points,
        neighbor_ids_host=nbr,
        height=img_meta["img_shape"][0] // stride,
        width=img_meta["img_shape"][1] // stride,
        k=k, k_feat=blob[o[3]:o[4]].view(k_rows.shape), n_views=v_all)
