"""CPU oracle for the MVSDet plane-sweep / depth-top-k / back-projection path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or
the timed CPU baseline.  The product package ``mvsdet_b200`` never imports it.

What it is: a restatement, in plain PyTorch CPU ops, of the algorithm the
reference runs in ``projects/NeRF-Det/nerfdet/mvsdet.py:430-515`` and the
functions it calls.  Each function cites the reference lines it follows.  The
third-party arithmetic on the path is PyTorch ATen itself (``F.grid_sample``,
``torch.inverse``, ``bmm``, ``topk``, ``softmax``, ``round``; the reference pins
pytorch 2.1.0, this image has 2.11.0 with unchanged semantics for these ops),
so the oracle calls the same ATen ops in the same order instead of
re-deriving them; an independent closed-form bilinear sampler
(``warp_closed_form``) is kept beside it as a cross-check.

How it is pinned: the reference ships no test, golden vector or fixture for
this path.  The pin is therefore the reference's own
code executed on seeded synthetic scenes in the build container
(``oracle/ref_loader.py`` slices the functions out of ``/root/reference`` and
runs them verbatim); ``tests/golden/make_golden.py`` commits those outputs
as ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this oracle
against them (integers bit-exact, floats to 1e-6).

Gradients: the reference has no hand-written backward; its
gradients are whatever autograd computes through these ops.  The oracle is
built from differentiable torch ops with the same data flow, so
``torch.autograd.grad`` through it is the gradient oracle.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "knn", "get_nearest_pose_ids", "collect_proj", "homography",
    "homo_warping", "warp_closed_form", "plane_sweep_variance", "group_correlation", "scene_group_correlation",
    "depth_values_for", "depth_probability", "sample_depth_prob",
    "compute_avg_depth", "compute_projection", "feature_intrinsics",
    "get_points", "backproject_weigh", "aggregate_views", "hot_path",
    "lift", "get_camera_params", "compute_depth_scale", "process_rgb_raw", "nvs_consumers",
]


# --------------------------------------------------------------------------
# a1. neighbour selection -- mvsdet.py:43-64 (knn), :67-104 (method 'dist')
# --------------------------------------------------------------------------
def knn(x: torch.Tensor, ref: torch.Tensor, k: int, maskself: bool = False):
    """x, ref: [B,3,N].  Indices of the k largest *negated squared distances*.

    Follows mvsdet.py:51-63: the matrix is ``-|r|^2 + 2 x.r - |x|^2`` built
    from one matmul and two squared-norm rows; with ``maskself`` the diagonal
    is overwritten with -100000 (not -inf) before ``topk``.
    """
    inner = -2 * torch.matmul(x.transpose(2, 1), ref)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    yy = torch.sum(ref ** 2, dim=1, keepdim=True)
    pairwise = -yy - inner - xx.transpose(2, 1)
    if maskself:
        assert x.shape == ref.shape
        diag = torch.arange(xx.shape[2])
        pairwise[:, diag, diag] = -100000
    return pairwise.topk(k=k, dim=-1)[1]


def get_nearest_pose_ids(c2w: torch.Tensor, num_select: int) -> torch.Tensor:
    """[V,4,4] camera-to-world -> [V,k] int64 neighbour ids, nearest first.

    mvsdet.py:67-104 with ``angular_dist_method='dist'``, ``maskself=True``
    (the only live branch, call site mvsdet.py:432-434); k = min(k, V-1).
    """
    num_select = min(num_select, len(c2w) - 1)
    locs = c2w[:, :3, 3].unsqueeze(0).transpose(2, 1)          # [1,3,V]
    return knn(locs, locs, k=num_select, maskself=True)[0]


# --------------------------------------------------------------------------
# a2. projection matrices -- mvsdet.py:249-264, :423-428
# --------------------------------------------------------------------------
def feature_intrinsics(intrinsic: torch.Tensor, ratio: float) -> torch.Tensor:
    """Rows 0-1 of the 4x4 (or [V,4,4]) image intrinsics divided by ``ratio``
    = ori_h / (img_h / stride)  (mvsdet.py:422-428)."""
    k = intrinsic.clone()
    if k.dim() == 2:
        k[:2] /= ratio
    else:
        k[:, :2] /= ratio
    return k


def collect_proj(w2c: torch.Tensor, intr: torch.Tensor, neighbor_ids: torch.Tensor):
    """proj = K_feat @ w2c ([V,4,4]); neighbour projections gathered per
    reference view -> (proj, [k x [V,4,4]])  (mvsdet.py:249-264)."""
    if intr.dim() == 2:
        intr = intr.unsqueeze(0).repeat(w2c.shape[0], 1, 1)
    proj = torch.matmul(intr, w2c)
    v, k = neighbor_ids.shape
    nei = proj[neighbor_ids.reshape(-1)].view(v, k, 4, 4)
    return proj, torch.unbind(nei, dim=1)


def homography(src_proj: torch.Tensor, ref_proj: torch.Tensor):
    """M = P_src @ inverse(P_ref) -> (rot [B,3,3], trans [B,3,1])
    (mvs_models/module.py:116-118)."""
    proj = torch.matmul(src_proj, torch.inverse(ref_proj))
    return proj[:, :3, :3], proj[:, :3, 3:4]


# --------------------------------------------------------------------------
# a3. homography warp -- mvs_models/module.py:105-146
# --------------------------------------------------------------------------
def _warp_pixel_coords(rot, trans, depth_values, height, width):
    """Source-view pixel coordinates (px, py) of every (plane, pixel).

    module.py:120-136: xyz=(x,y,1) row-major over (y,x); rot@xyz; times depth;
    plus trans; x/z, y/z with no z guard and no epsilon.
    Returns px, py of shape [B, D, H*W].
    """
    b = rot.shape[0]
    d = depth_values.shape[1]
    dev = rot.device
    ys, xs = torch.meshgrid(
        torch.arange(0, height, dtype=torch.float32, device=dev),
        torch.arange(0, width, dtype=torch.float32, device=dev), indexing="ij")
    xyz = torch.stack((xs.reshape(-1), ys.reshape(-1),
                       torch.ones(height * width, device=dev)))       # [3,HW]
    rot_xyz = torch.matmul(rot, xyz.unsqueeze(0).expand(b, 3, -1))     # [B,3,HW]
    if depth_values.dim() == 2:
        dv = depth_values.view(b, 1, d, 1)
    else:                                           # per-pixel depth [B,D,H,W]
        dv = depth_values.reshape(b, d, -1).unsqueeze(1)
    q = rot_xyz.unsqueeze(2) * dv + trans.view(b, 3, 1, 1)            # [B,3,D,HW]
    pxy = q[:, :2] / q[:, 2:3]
    return pxy[:, 0], pxy[:, 1]


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """[B,C,H,W] neighbour features -> [B,C,D,H,W] warped onto the D
    fronto-parallel planes of the reference view (module.py:105-146).

    Normalisation divides by (W-1)/2 and (H-1)/2 (module.py:137-138) while
    ``grid_sample`` runs with its default ``align_corners=False``
    (module.py:142); both are kept -- the mismatch is part of the spec
    (SURVEY.md Appendix A.1).  The grid carries no gradient (module.py:115).
    """
    b, c, h, w = src_fea.shape
    d = depth_values.shape[1]
    with torch.no_grad():
        rot, trans = homography(src_proj, ref_proj)
        px, py = _warp_pixel_coords(rot, trans, depth_values, h, w)
        gx = px / ((w - 1) / 2) - 1
        gy = py / ((h - 1) / 2) - 1
        grid = torch.stack((gx, gy), dim=3)                            # [B,D,HW,2]
    out = F.grid_sample(src_fea, grid.view(b, d * h, w, 2), mode="bilinear",
                        padding_mode="zeros", align_corners=False)
    return out.view(b, c, d, h, w)


def warp_closed_form(src_fea, rot, trans, depth_values):
    """Independent cross-check of ``homo_warping``: explicit 4-tap bilinear
    gather following SURVEY.md Appendix A.1 (ix = px*W/(W-1) - 0.5).  Not used
    by the parity tests as the oracle, only to validate it."""
    b, c, h, w = src_fea.shape
    d = depth_values.shape[1]
    px, py = _warp_pixel_coords(rot, trans, depth_values, h, w)
    ix = ((px / ((w - 1) / 2) - 1) + 1) * (w / 2) - 0.5
    iy = ((py / ((h - 1) / 2) - 1) + 1) * (h / 2) - 0.5
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx, wy = ix - x0, iy - y0
    flat = src_fea.reshape(b, c, h * w)
    out = torch.zeros(b, c, d, h * w, dtype=src_fea.dtype)
    for dy, dx, wt in ((0, 0, (1 - wy) * (1 - wx)), (0, 1, (1 - wy) * wx),
                       (1, 0, wy * (1 - wx)), (1, 1, wy * wx)):
        xt, yt = x0 + dx, y0 + dy
        ok = (xt >= 0) & (xt <= w - 1) & (yt >= 0) & (yt <= h - 1)
        idx = (yt.clamp(0, h - 1) * w + xt.clamp(0, w - 1)).long()
        idx = torch.where(ok, idx, torch.zeros_like(idx))
        g = torch.gather(flat, 2, idx.view(b, 1, d * h * w).expand(b, c, -1))
        wt = torch.where(ok, wt, torch.zeros_like(wt))
        out += g.view(b, c, d, h * w) * wt.unsqueeze(1)
    return out.view(b, c, d, h, w)


# --------------------------------------------------------------------------
# a4. variance accumulation -- mvsdet.py:439-467
# --------------------------------------------------------------------------
def depth_values_for(near_far_range: Sequence[float], num_depth: int) -> np.ndarray:
    """np.arange(near, far, (far-near)/D, float32), must have D entries
    (mvsdet.py:222-225)."""
    interval = (near_far_range[1] - near_far_range[0]) / num_depth
    dv = np.arange(near_far_range[0], near_far_range[1], interval, dtype=np.float32)
    assert len(dv) == num_depth
    return dv


def plane_sweep_variance(feature, w2c, feat_intrinsic, neighbor_ids, depth_values,
                         training: bool = True, view_subset: Optional[int] = None):
    """[V,C,Hf,Wf] -> variance volume [V,C,D,Hf,Wf] (mvsdet.py:439-467).

    S1 = ref + sum_j warped_j, S2 = ref^2 + sum_j warped_j^2 with the
    reference view broadcast over D; var = S2/(k+1) - (S1/(k+1))^2, in exactly
    that order.  ``training`` selects the out-of-place branch (:458-459,
    autograd-safe) or the in-place eval branch (:463-464); the values are the
    same.  ``view_subset`` = n sweeps only the first n reference views (their
    neighbours are still drawn from all views): the bounded CPU-baseline sample
    of bench.py, not a reference feature.
    """
    v = feature.shape[0] if view_subset is None else int(view_subset)
    k = neighbor_ids.shape[1]
    d = depth_values.shape[-1]
    neighbor_ids = neighbor_ids[:v]
    ref_volume = feature[:v].unsqueeze(2).repeat(1, 1, d, 1, 1)
    volume_sum = ref_volume
    volume_sq_sum = ref_volume ** 2
    del ref_volume
    nei_features = feature[neighbor_ids.reshape(-1)].view(v, k, *feature.shape[1:])
    all_proj = torch.matmul(feat_intrinsic if feat_intrinsic.dim() == 3
                            else feat_intrinsic.unsqueeze(0).repeat(w2c.shape[0], 1, 1), w2c)
    ref_proj = all_proj[:v]
    nei_projs = torch.unbind(all_proj[neighbor_ids.reshape(-1)].view(v, k, 4, 4), dim=1)
    if depth_values.dim() == 1:
        depth_values = depth_values.unsqueeze(0).repeat(v, 1)
    for j in range(k):
        warped = homo_warping(nei_features[:, j], nei_projs[j], ref_proj, depth_values)
        if training:
            volume_sum = volume_sum + warped
            volume_sq_sum = volume_sq_sum + warped ** 2
        else:
            volume_sum += warped
            volume_sq_sum += warped.pow_(2)
        del warped
    if training:
        return volume_sq_sum / (k + 1) - (volume_sum / (k + 1)) ** 2
    return volume_sq_sum.div_(k + 1).sub_(volume_sum.div_(k + 1).pow_(2))


def group_correlation(feature, w2c, feat_intrinsic, neighbor_ids, depth_values, num_groups: int):
    """[V,C,Hf,Wf] -> group-wise correlation cost volumes [V,k,G,D,Hf,Wf], one per neighbour
    (SURVEY.md 8f rank 4).  The arithmetic follows mvs_models/lss_fpn.py:485-506 statement by
    statement -- reshape the warped and the reference features into ``num_groups`` contiguous channel
    groups, ``torch.mean(ref.unsqueeze(3) * warped, axis=2)`` -- with MVSDet's own warp
    (``homo_warping``, module.py:105-146) and neighbour selection in place of that file's
    BEVStereo-specific ones.  The reference then feeds every volume to a small 3-D net and averages the
    scores over the neighbours (:506-508); that net is outside the path, the volumes are the hand-off."""
    v, c, hf, wf = feature.shape
    k = neighbor_ids.shape[1]
    d = depth_values.shape[-1]
    all_proj = torch.matmul(feat_intrinsic if feat_intrinsic.dim() == 3
                            else feat_intrinsic.unsqueeze(0).repeat(w2c.shape[0], 1, 1), w2c)
    ref_proj = all_proj[:v]
    nei_projs = torch.unbind(all_proj[neighbor_ids.reshape(-1)].view(v, k, 4, 4), dim=1)
    if depth_values.dim() == 1:
        depth_values = depth_values.unsqueeze(0).repeat(v, 1)
    ref_feat = feature.reshape(v, num_groups, c // num_groups, hf, wf)                      # :499-501
    costs = []
    for j in range(k):
        warped = homo_warping(feature[neighbor_ids[:, j]], nei_projs[j], ref_proj, depth_values)
        warped = warped.reshape(v, num_groups, c // num_groups, d, hf, wf)                  # :496-498
        costs.append(torch.mean(ref_feat.unsqueeze(3) * warped, axis=2))                    # :502-503
    return torch.stack(costs, dim=1)


def scene_group_correlation(feature, img_meta, *, near_far_range, num_depth, num_groups, stride: int = 4):
    """``group_correlation`` with the scene glue of ``hot_path`` (feature-level intrinsics, pose
    neighbours, depth planes: mvsdet.py:422-434, :222-225)."""
    v = feature.shape[0]
    ratio = img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride)
    w2c = torch.as_tensor(np.array(img_meta["lidar2img"]["extrinsic"]))
    intr = torch.as_tensor(np.array(img_meta["lidar2img"]["intrinsic"]))
    k_feat = feature_intrinsics(intr, ratio)
    neighbor_ids = get_nearest_pose_ids(w2c.inverse(), min(2, v - 1))
    dvals = torch.as_tensor(depth_values_for(near_far_range, num_depth))
    return group_correlation(feature, w2c, k_feat, neighbor_ids, dvals, num_groups)


# --------------------------------------------------------------------------
# a5-a7. probabilities and hypotheses -- mvsdet.py:470-482, :266-283, :298-317
# --------------------------------------------------------------------------
def depth_probability(cost_reg_out: torch.Tensor):
    """[V,2,D,H,W] cost-regularisation output -> (softmax over D of channel 0,
    sigmoid of channel 1), each [V,D,H,W]  (mvsdet.py:470-475)."""
    cost_reg, off = torch.unbind(cost_reg_out, dim=1)
    return F.softmax(cost_reg, dim=1), torch.sigmoid(off)


def sample_depth_prob(prob_volume, off_pred, topk, near, depth_interval,
                      return_idx: bool = False):
    """Top-k hypotheses (mvsdet.py:266-283): densities, idx = topk over D;
    depth = idx*interval + near, then += off[idx]*interval."""
    est_densities, est_idx = prob_volume.topk(k=topk, dim=1)
    est_depth = est_idx * depth_interval + torch.tensor(near)
    off_d = torch.gather(off_pred, 1, est_idx) * torch.tensor(depth_interval)
    est_depth = est_depth + off_d
    if return_idx:
        return est_depth, est_densities, est_idx
    return est_depth, est_densities


def compute_avg_depth(prob_volume, off_pred, near, depth_interval):
    """Depth expectation sum_d p_d (d*interval + near + off_d*interval), summed
    in probability-sorted order because the reference obtains the terms from a
    full-length ``topk``  (mvsdet.py:298-317)."""
    d = prob_volume.shape[1]
    dens, idx = prob_volume.topk(k=d, dim=1)
    depth = idx * depth_interval + torch.tensor(near)
    depth = depth + torch.gather(off_pred, 1, idx) * torch.tensor(depth_interval)
    return torch.sum(depth * dens, dim=1)


# --------------------------------------------------------------------------
# a8-a9. voxel geometry -- mvsdet.py:1124-1156, :1316-1327
# --------------------------------------------------------------------------
def compute_projection(intrinsic, extrinsics, ratio: float) -> torch.Tensor:
    """P_i = K_feat[:3,:3] @ w2c_i[:3] -> [V,3,4]  (mvsdet.py:1124-1156, both
    the shared-K branch :1143-1155 and the per-view-K branch :1127-1141)."""
    intrinsic = torch.as_tensor(np.array(intrinsic))
    extr = torch.as_tensor(np.array(extrinsics))
    out = []
    for i in range(extr.shape[0]):
        k = (intrinsic[i] if intrinsic.dim() == 3 else intrinsic)[:3, :3].clone()
        k[:2] /= ratio
        out.append(k @ extr[i][:3])
    return torch.stack(out)


def get_points(n_voxels, voxel_size, origin) -> torch.Tensor:
    """Corner-anchored voxel coordinates [3,nx,ny,nz]:
    idx*voxel_size + (origin - n_voxels/2*voxel_size)  (mvsdet.py:1316-1327)."""
    n_voxels = torch.as_tensor(n_voxels)
    voxel_size = torch.as_tensor(voxel_size, dtype=torch.float32)
    origin = torch.as_tensor(origin, dtype=torch.float32)
    idx = torch.stack(torch.meshgrid(torch.arange(n_voxels[0]),
                                     torch.arange(n_voxels[1]),
                                     torch.arange(n_voxels[2]), indexing="ij"))
    new_origin = origin - n_voxels / 2. * voxel_size
    return idx * voxel_size.view(3, 1, 1, 1) + new_origin.view(3, 1, 1, 1)


# --------------------------------------------------------------------------
# a10. probabilistic back-projection -- mvsdet.py:1372-1492
# --------------------------------------------------------------------------
def backproject_weigh(features, points, projection, depth, voxel_size, prob,
                      return_debug: bool = False):
    """features [V,C,h,w], points [3,nx,ny,nz], projection [V,3,4],
    depth/prob [V,h*w,1,T] -> (volume [V,C,nx,ny,nz], valid [V,1,nx,ny,nz] bool).

    Vectorised restatement of the reference's V*T Python loop
    (mvsdet.py:1401-1427) and of the per-view assignment loop (:1457-1460):
      p = P_i @ (X,1) (bmm, :1384-1386); x,y = round(p.xy/p.z).long();
      in-bounds = x>=0 & y>=0 & x<w & y<h & z>0 (:1388-1391);
      pn = prob / sum_T prob (:1395-1396);
      pass_j = in-bounds & z > depth_j - vs_z & z < depth_j + vs_z (:1407-1408);
      weight = max_j (pass_j ? pn_j : 0) (:1410-1422); valid = any_j pass_j;
      volume = valid ? features[:, y, x] * weight : 0 (:1457-1460).
    """
    v, c, h, w = features.shape
    nx, ny, nz = points.shape[-3:]
    n = nx * ny * nz
    pts = points.reshape(1, 3, -1).expand(v, 3, -1)
    pts = torch.cat((pts, torch.ones_like(pts[:, :1])), dim=1)
    p = torch.bmm(projection, pts)                                     # [V,3,N]
    x = (p[:, 0] / p[:, 2]).round().long()
    y = (p[:, 1] / p[:, 2]).round().long()
    z = p[:, 2]
    inb = (x >= 0) & (y >= 0) & (x < w) & (y < h) & (z > 0)
    pix = (y.clamp(0, h - 1) * w + x.clamp(0, w - 1))                  # [V,N]
    pix = torch.where(inb, pix, torch.zeros_like(pix))
    t = depth.shape[-1] * depth.shape[-2]
    depth_f = depth.reshape(v, h * w, t)
    prob_f = prob.reshape(v, h * w, t)
    prob_norm = prob_f / prob_f.sum(dim=-1, keepdim=True)
    gidx = pix.unsqueeze(-1).expand(v, n, t)
    d_at = torch.gather(depth_f, 1, gidx)                              # [V,N,T]
    pn_at = torch.gather(prob_norm, 1, gidx)
    vs = float(voxel_size[-1])
    passed = inb.unsqueeze(-1) & (z.unsqueeze(-1) > d_at - vs) & (z.unsqueeze(-1) < d_at + vs)
    cand = torch.where(passed, pn_at, torch.zeros_like(pn_at))
    weight = torch.max(cand.permute(2, 0, 1), dim=0)[0]                # [V,N]
    valid = passed.any(dim=-1)
    feat_at = torch.gather(features.reshape(v, c, h * w), 2,
                           pix.unsqueeze(1).expand(v, c, n))           # [V,C,N]
    volume = torch.where(valid.unsqueeze(1), feat_at, torch.zeros_like(feat_at))
    volume = volume * weight.unsqueeze(1)
    volume = volume.view(v, c, nx, ny, nz)
    valid_out = valid.view(v, 1, nx, ny, nz)
    if return_debug:
        return volume, valid_out, dict(x=x, y=y, z=z, inb=inb, weight=weight,
                                       passed=passed)
    return volume, valid_out


# --------------------------------------------------------------------------
# a11. view aggregation -- mvsdet.py:511-515, :681-682
# --------------------------------------------------------------------------
def aggregate_views(volume, valid):
    """sum over views / (count + 1e-8), zero where count == 0.
    Returns (volume_mean [C,nx,ny,nz], count [1,nx,ny,nz] int64)."""
    volume_sum = volume.sum(dim=0)
    count = valid.sum(dim=0)
    volume_mean = volume_sum / (count + 1e-8)
    volume_mean = torch.where((count[0] == 0).unsqueeze(0),
                              torch.zeros_like(volume_mean), volume_mean)
    return volume_mean, count


# --------------------------------------------------------------------------
# f3. NVS-branch consumers of the path -- mvsdet.py:1158-1218, :1272-1313, :319-333, :488-494, :579-583
# --------------------------------------------------------------------------
def lift(x, y, z, intrinsics):
    """Pixel (x, y) at depth z -> homogeneous camera coordinates (mvsdet.py:1300-1313; the
    reference's hard-coded ``.cuda()`` calls are dropped, nothing else)."""
    fx = intrinsics[:, 0, 0]
    fy = intrinsics[:, 1, 1]
    cx = intrinsics[:, 0, 2]
    cy = intrinsics[:, 1, 2]
    sk = intrinsics[:, 0, 1]
    x_lift = (x - cx.unsqueeze(-1) + cy.unsqueeze(-1) * sk.unsqueeze(-1) / fy.unsqueeze(-1)
              - sk.unsqueeze(-1) * y / fy.unsqueeze(-1)) / fx.unsqueeze(-1) * z
    y_lift = (y - cy.unsqueeze(-1)) / fy.unsqueeze(-1) * z
    return torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1)


def get_camera_params(uv, pose, intrinsics):
    """Unit ray directions of pixels uv [B,N,2] for camera-to-world ``pose`` [B,4,4]
    (mvsdet.py:1272-1298) -> (ray_dirs [B,N,3], cam_loc [B,3])."""
    cam_loc = pose[:, :3, 3]
    batch_size, num_samples, _ = uv.shape
    depth = torch.ones((batch_size, num_samples))
    x_cam = uv[:, :, 0].view(batch_size, -1)
    y_cam = uv[:, :, 1].view(batch_size, -1)
    z_cam = depth.view(batch_size, -1)
    pixel_points_cam = lift(x_cam, y_cam, z_cam, intrinsics=intrinsics).permute(0, 2, 1)
    world_coords = torch.bmm(pose, pixel_points_cam).permute(0, 2, 1)[:, :, :3]
    ray_dirs = F.normalize(world_coords - cam_loc[:, None, :], dim=2)
    return ray_dirs, cam_loc


def compute_depth_scale(height, width, img_meta, stride, num_src):
    """z component of the unit ray through every feature-level pixel of a camera with identity
    pose: (1, num_src, h*w, 1).  Both reference variants: shared intrinsics
    (compute_depth_scale, mvsdet.py:1158-1187) and a list of per-view intrinsics
    (compute_depth_scale_MultiIntrin, :1189-1218)."""
    indices = [torch.arange(height), torch.arange(width)]
    uv = torch.stack(torch.meshgrid(*indices, indexing="ij"), dim=0)
    uv = torch.flip(uv, dims=[0]).float()
    uv = uv.reshape(2, -1).transpose(1, 0).unsqueeze(0)                    # (1, h*w, 2) as (x, y)
    ratio = img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride)
    intr = img_meta["lidar2img"]["intrinsic"]
    if isinstance(intr, (list, tuple)):
        num_src = len(intr)
        uv = uv.repeat(num_src, 1, 1)
        intrinsic = torch.tensor(np.array(intr))
        intrinsic[:, :2] /= ratio
        pose = torch.eye(4).unsqueeze(0).repeat(num_src, 1, 1)
        ray_dirs, _ = get_camera_params(uv, pose, intrinsic)
        return ray_dirs[:, :, 2:].unsqueeze(0)
    intrinsic = torch.tensor(np.array(intr))
    intrinsic[:2] /= ratio
    ray_dirs, _ = get_camera_params(uv, torch.eye(4)[None], intrinsic.unsqueeze(0))
    return ray_dirs[0, :, 2:].unsqueeze(0).unsqueeze(0).repeat(1, num_src, 1, 1)


def process_rgb_raw(orig_rgb, ratio, height, width, src_id):
    """(n_src,3,H,W) images -> (1, n_nei, height*width, 3): bilinear 1/ratio down-sampling, crop,
    re-layout (mvsdet.py:319-333)."""
    assert ratio == 4
    new_rgb = F.interpolate(orig_rgb[src_id], scale_factor=1. / ratio, mode="bilinear")
    new_rgb = new_rgb[:, :, :height, :width]
    return new_rgb.reshape(*new_rgb.shape[:2], -1).transpose(2, 1).unsqueeze(0)


def nvs_consumers(prob_volume, est_depth_r, depth_coding, img_meta, stride, height, width):
    """What the NVS branch derives from the path's outputs (mvsdet.py:488-494, :579, :583):
    est_depth_r [V,h*w,1,T], depth_coding [V,1,h,w] ->
    dict(depth_scale [V,h*w,1], est_ray_depth [V,h*w,1,T], ray_depth_coding [V,h*w,1],
    opacity [V,Hf,Wf])."""
    v = prob_volume.shape[0]
    scale = compute_depth_scale(height, width, img_meta, stride, v).squeeze(0)            # (V,h*w,1)
    est_ray_depth = est_depth_r / (scale.unsqueeze(-1).repeat(1, 1, 1, est_depth_r.shape[-1]) + 1e-8)
    coding = depth_coding.reshape(v, 1, -1).transpose(2, 1)                               # (V,h*w,1)
    ray_depth_coding = coding / (scale + 1e-8)
    opacity = torch.max(prob_volume, dim=1)[0]
    return dict(depth_scale=scale, est_ray_depth=est_ray_depth, ray_depth_coding=ray_depth_coding,
                opacity=opacity)


# --------------------------------------------------------------------------
# the inline glue of extract_feat -- mvsdet.py:404-515
# --------------------------------------------------------------------------
def hot_path(feature, img_meta, cost_regularization, *, near_far_range, num_depth,
             topk, n_voxels, voxel_size, stride: int = 4, training: bool = True,
             view_subset: Optional[int] = None):
    """One scene through mvsdet.py:404-515 / :681-682.

    ``feature`` [V,C,Hf,Wf]; ``img_meta`` carries ``lidar2img`` {extrinsic,
    intrinsic, origin}, ``img_shape``, ``ori_shape`` as the reference's dataset
    produces them; ``cost_regularization`` maps the variance volume
    [V,C,D,Hf,Wf] to [V,2,D,Hf,Wf] (CostRegNet_3DGS in the reference; any
    callable here, it is not part of the path).
    Returns a dict of every intermediate the parity tests compare.
    """
    v = feature.shape[0]
    ratio = img_meta["ori_shape"][0] / (img_meta["img_shape"][0] / stride)
    projection = compute_projection(img_meta["lidar2img"]["intrinsic"],
                                    img_meta["lidar2img"]["extrinsic"], ratio)
    points = get_points(n_voxels, voxel_size, img_meta["lidar2img"]["origin"])
    height = img_meta["img_shape"][0] // stride
    width = img_meta["img_shape"][1] // stride
    w2c = torch.as_tensor(np.array(img_meta["lidar2img"]["extrinsic"]))
    intr = torch.as_tensor(np.array(img_meta["lidar2img"]["intrinsic"]))
    k_feat = feature_intrinsics(intr, ratio)

    k = min(2, v - 1)
    neighbor_ids = get_nearest_pose_ids(w2c.inverse(), k)
    depth_interval = (near_far_range[1] - near_far_range[0]) / num_depth
    dvals = torch.as_tensor(depth_values_for(near_far_range, num_depth))
    dev = feature.device                   # CPU for the checker; "cuda" only for tools/eager_gpu.py
    if dev.type != "cpu":
        projection, points, w2c, k_feat, neighbor_ids, dvals = (
            t.to(dev) for t in (projection, points, w2c, k_feat, neighbor_ids, dvals))
    variance = plane_sweep_variance(feature, w2c, k_feat, neighbor_ids, dvals,
                                    training=training, view_subset=view_subset)
    if view_subset is not None:            # bounded CPU-baseline sample (bench.py)
        v = int(view_subset)
        projection = projection[:v]
    cost_out = cost_regularization(variance)
    prob_volume, off_pred = depth_probability(cost_out)
    est_depth, est_dens, est_idx = sample_depth_prob(
        prob_volume, off_pred, topk, near_far_range[0], depth_interval, return_idx=True)
    est_depth_c = est_depth[:, :, :height, :width]
    est_dens_c = est_dens[:, :, :height, :width]
    depth_coding = compute_avg_depth(prob_volume, off_pred, near_far_range[0],
                                     depth_interval)[:, :height, :width].unsqueeze(1)
    depth_r = est_depth_c.reshape(v, topk, -1).transpose(2, 1).unsqueeze(2)
    dens_r = est_dens_c.reshape(v, topk, -1).transpose(2, 1).unsqueeze(2)
    volume, valid = backproject_weigh(feature[:v, :, :height, :width], points,
                                      projection, depth_r, voxel_size, dens_r)
    volume_mean, count = aggregate_views(volume, valid)
    return dict(neighbor_ids=neighbor_ids, variance=variance, cost_out=cost_out,
                prob_volume=prob_volume, off_pred=off_pred, est_depth=est_depth,
                est_densities=est_dens, est_idx=est_idx, depth_coding=depth_coding,
                valid=valid, volume_mean=volume_mean, count=count,
                projection=projection, points=points)
