"""Drop-in mirrors of the reference's hot-path functions.

Same names, argument meaning, return shapes and error behaviour as
  homo_warping            projects/NeRF-Det/nerfdet/mvs_models/module.py:105-146
  sample_depth_prob       MVSDet.sample_depth_prob, mvsdet.py:266-283
  compute_avg_depth       MVSDet.compute_avg_depth, mvsdet.py:298-317
  backproject_Weigh       mvsdet.py:1372-1492
  compute_depth_scale[_MultiIntrin]   MVSDet.compute_depth_scale*, mvsdet.py:1158-1218
  process_rgb_raw         MVSDet.process_rgb_raw, mvsdet.py:319-333
  get_points, knn, get_nearest_pose_ids, collect_proj, _compute_projection
but every tensor op runs in the sm_100a kernels of this package (ops.py).  The
methods of the reference that read ``self`` take the same values as keyword
arguments (``near``, ``depth_interval``).  The fused entry points the
replacement of mvsdet.py:430-515 actually uses are in ``hotpath.py``.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import geometry as G
from . import ops

knn = G.knn
get_nearest_pose_ids = G.get_nearest_pose_ids
collect_proj = G.collect_proj
get_points = G.get_points
_compute_projection = G.compute_projection

__all__ = ["homo_warping", "sample_depth_prob", "compute_avg_depth", "backproject_Weigh",
           "knn", "get_nearest_pose_ids", "collect_proj", "get_points",
           "_compute_projection", "compute_depth_scale", "compute_depth_scale_MultiIntrin",
           "process_rgb_raw"]


def homo_warping(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor,
                 depth_values: torch.Tensor) -> torch.Tensor:
    """src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D]
    -> warped [B,C,D,H,W] (channels_last_3d memory).  No gradient reaches the
    projection matrices or the depths (the reference builds the grid under
    no_grad, module.py:115)."""
    if depth_values.dim() not in (2, 4):
        raise ValueError("depth_values must be [B,D] or per-pixel [B,D,H,W] (module.py:126-133)")
    with torch.no_grad():
        # 4x4 algebra on the host in fp32 (LAPACK), like the per-scene geometry
        # block: bit-identical to the CPU reference; cuSOLVER's batched inverse
        # differs in the last bits, which moves sample positions by ~1e-5 px.
        hom = G.homography_params(src_proj.detach().float().cpu(),
                                  ref_proj.detach().float().cpu()).to(src_fea.device)
    feat = ops.pack_features(src_fea, src_fea.dtype if src_fea.dtype == torch.bfloat16 else torch.float32)
    return ops.homo_warp(feat, hom, depth_values.to(src_fea.device).float().contiguous(),
                         out_dtype=torch.float32)


def _stack_prob_off(prob_volume, off_pred):
    if prob_volume.shape != off_pred.shape or prob_volume.dim() != 4:
        raise ValueError("prob_volume and off_pred must both be [V,D,H,W]")
    return torch.stack((prob_volume, off_pred), dim=1)


def sample_depth_prob(prob_volume: torch.Tensor, off_pred: torch.Tensor, topk: int = 3, *,
                      near: float, depth_interval: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """prob_volume, off_pred [V,D,H,W] (already softmax-ed / sigmoid-ed) ->
    (est_depth, est_density) each [V,topk,H,W]."""
    res = ops.topk_hypotheses(_stack_prob_off(prob_volume, off_pred), near, depth_interval, topk)
    return res[2], res[3]


def compute_avg_depth(prob_volume: torch.Tensor, off_pred: torch.Tensor, *, near: float,
                      depth_interval: float) -> torch.Tensor:
    """Depth expectation [V,H,W] = sum_d p_d (d*interval + near + off_d*interval)."""
    res = ops.topk_hypotheses(_stack_prob_off(prob_volume, off_pred), near, depth_interval, 1)
    return res[5]


def backproject_Weigh(features, points, projection, depth, voxel_size, prob, gt_depth=None,
                      save_dir=None, img_meta=None, depth_mean=None):
    """features [V,C,h,w] (may be a crop of a larger map), points [3,nx,ny,nz],
    projection [V,3,4], depth/prob [V,h*w,num_surface,T]
    -> (volume [V,C,nx,ny,nz], valid bool [V,1,nx,ny,nz], gap_all, rmse).

    ``gt_depth`` only feeds debug scalars in the reference (mvsdet.py:1431-1486);
    that branch is not part of the path: passing it raises."""
    if gt_depth is not None:
        raise ValueError("the gt_depth debug branch of backproject_Weigh (mvsdet.py:1431-1486) "
                         "is outside the accelerated path")
    v, c, h, w = features.shape
    nx, ny, nz = points.shape[-3:]
    dtype = features.dtype if features.dtype == torch.bfloat16 else torch.float32
    s0, s1, s2, s3 = features.stride()
    if (features.dtype == dtype and s1 == 1 and s3 == c and s2 % c == 0 and s2 >= w * c
            and s0 % s2 == 0 and s0 >= h * s2):
        # a top-left crop of a channels-last map (mvsdet.py:499): address the
        # parent map in place, the kernel only touches rows < h, columns < w
        feat = features.as_strided((v, c, s0 // s2, s2 // c), (s0, 1, s2, c))
    else:
        feat = ops.pack_features(features, dtype)
    volume, valid = ops.backproject_per_view(feat, points.to(feat.device), projection.to(feat.device),
                                             depth, prob, float(voxel_size[-1]), h, w)
    volume = volume.reshape(v, c, nx, ny, nz) if volume.is_contiguous() else \
        volume.unflatten(2, (nx, ny, nz))
    valid = valid.view(v, 1, nx, ny, nz)
    return volume, valid, torch.tensor(1.), torch.tensor(1.)


def _feature_intrinsics_device(img_meta, stride, device):
    import numpy as np
    intr = np.asarray(img_meta["lidar2img"]["intrinsic"], dtype=np.float32).copy()
    intr[..., :2, :] /= np.float32(img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride))
    return torch.from_numpy(intr).to(device)


def compute_depth_scale(height, width, device, img_meta, stride, num_src):
    """MVSDet.compute_depth_scale (mvsdet.py:1158-1187): shared intrinsics -> (1, num_src, h*w, 1)."""
    k = _feature_intrinsics_device(img_meta, stride, device)
    if k.dim() != 2:
        raise ValueError("compute_depth_scale takes one shared 4x4 intrinsic; use compute_depth_scale_MultiIntrin")
    scale = ops.ray_depth_scale(k, 1, height, width)                  # identical for every view
    return scale.reshape(1, 1, height * width, 1).repeat(1, num_src, 1, 1)


def compute_depth_scale_MultiIntrin(height, width, device, img_meta, stride, num_src):
    """MVSDet.compute_depth_scale_MultiIntrin (mvsdet.py:1189-1218): one intrinsic per view
    (``num_src`` is overridden by the list length, as in the reference)."""
    k = _feature_intrinsics_device(img_meta, stride, device)
    if k.dim() != 3:
        raise ValueError("compute_depth_scale_MultiIntrin takes a list of per-view 4x4 intrinsics")
    scale = ops.ray_depth_scale(k, k.shape[0], height, width)
    return scale.reshape(1, k.shape[0], height * width, 1)


def process_rgb_raw(orig_rgb, ratio, height, width, src_id):
    """MVSDet.process_rgb_raw (mvsdet.py:319-333): (n_src,3,H,W) -> (1, num_nei, height*width, 3)."""
    if ratio != 4:
        raise ValueError("process_rgb_raw: the reference asserts ratio == 4 (mvsdet.py:326)")
    return ops.rgb_downsample4(orig_rgb, src_id, height, width)
