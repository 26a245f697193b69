#!/usr/bin/env python
"""Host-side replay of the plane-sweep backward's RED-merging schemes on the
benchmark scene (no GPU): how many 1 KB source-pixel reductions leave the SM for

  * no merging (pixel kernel, tuning 5=1),
  * the run kernel's single pending column (default, plane_sweep_bwd_run.cu),
  * R-row blocks with two pending columns per source row (plane_sweep_bwd_rows.cu),
  * the ideal (distinct targets per R x 8 block and per R x 8 x D block).

ncu on B200 measured 3.79 GB of RED payload for the default kernel (3.69 GB
predicted here + 0.10 GB of reference-gradient REDs) and 2.86 GB for the
two-row kernel (2.73 + 0.10 predicted).

    python tools/red_merge_sim.py [--views 20] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvsdet_b200.geometry import scene_geometry  # noqa: E402
from mvsdet_b200.scene import SceneConfig, make_scene  # noqa: E402

NO = -2


def tap_ids(cfg, seed):
    """-> T [V,k,D,H,W,4] source-pixel index of each bilinear tap (order 00,01,10,11), -1 = outside."""
    sc = make_scene(cfg, seed=seed, with_grads=False)
    g = scene_geometry(sc["img_meta"], stride=cfg.stride, near_far_range=cfg.near_far_range,
                       num_depth=cfg.num_depth, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                       device="cpu")
    hom = g.hom.numpy().astype(np.float64)
    dv = g.depth_values.numpy()
    V, k = g.neighbor_ids.shape
    H, W = cfg.feat_hw
    D = cfg.num_depth
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    pix = np.stack([xs.ravel(), ys.ravel(), np.ones(H * W)])
    T = np.empty((V, k, D, H, W, 4), np.int64)
    for v in range(V):
        for j in range(k):
            R, t = hom[v, j, :9].reshape(3, 3), hom[v, j, 9:]
            for d in range(D):
                q = R @ pix * dv[v, d] + t[:, None]
                with np.errstate(all="ignore"):
                    ix = (q[0] / q[2] / ((W - 1) / 2)) * W / 2 - 0.5
                    iy = (q[1] / q[2] / ((H - 1) / 2)) * H / 2 - 0.5
                ix = np.nan_to_num(ix, nan=-1e6, posinf=1e6, neginf=-1e6).clip(-1e6, 1e6)
                iy = np.nan_to_num(iy, nan=-1e6, posinf=1e6, neginf=-1e6).clip(-1e6, 1e6)
                x0 = np.floor(ix).astype(np.int64).reshape(H, W)
                y0 = np.floor(iy).astype(np.int64).reshape(H, W)
                n = 0
                for dy in (0, 1):
                    for dx in (0, 1):
                        xx, yy = x0 + dx, y0 + dy
                        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                        T[v, j, d, :, :, n] = np.where(ok, yy * W + xx, -1)
                        n += 1
    return T


def one_slot(state, l, r, cnt):       # plane_sweep_bwd_run.cu scatter_side_p
    if state[0] == l and l >= 0:
        cnt[0] += 1
    else:
        if state[0] >= 0:
            cnt[0] += 1
        if l >= 0:
            cnt[0] += 1
    state[0] = r if r >= 0 else NO


def one_slot_stay(state, l, r, cnt):   # plane_sweep_bwd_run.cu side_q (default kernel)
    if state[0] == l and l >= 0:
        cnt[0] += 1
        state[0] = r if r >= 0 else NO
    elif state[0] == r and r >= 0:
        cnt[0] += int(l >= 0)
    else:
        cnt[0] += int(state[0] >= 0) + int(l >= 0)
        state[0] = r if r >= 0 else NO


def two_slot(st, l, r, cnt):          # plane_sweep_bwd_rows.cu side_add
    L, R = st
    if l == L and r == R:
        return
    if l >= 0 and l == R:
        cnt[0] += int(L >= 0)
    elif r >= 0 and r == L:
        cnt[0] += int(R >= 0)
    else:
        cnt[0] += int(L >= 0) + int(R >= 0)
    st[0], st[1] = l, r


def replay(T, rows, slots, run=8):
    V, k, D, H, W, _ = T.shape
    cnt = [0]
    add = {1: one_slot, 2: two_slot, "stay": one_slot_stay}[slots]
    flush_on_empty = slots == 1
    slots = 2 if slots == 2 else 1
    for v in range(V):
        for j in range(k):
            for d in range(D):
                t = T[v, j, d]
                for y in range(0, H, rows):
                    rr = list(range(y, min(H, y + rows)))
                    for x0 in range(0, W, run):
                        sides = [[NO] * slots for _ in range(len(rr) + 1)]
                        for x in range(x0, min(W, x0 + run)):
                            for n, r in enumerate(rr):
                                a = t[r, x]
                                if (a < 0).all():
                                    if flush_on_empty:  # the first run kernels flush on an empty sample
                                        for s in (sides[n], sides[n + 1]):
                                            cnt[0] += int(s[0] >= 0)
                                            s[0] = NO
                                    continue
                                add(sides[n], a[0], a[1], cnt)
                                add(sides[n + 1], a[2], a[3], cnt)
                        for s in sides:
                            cnt[0] += sum(1 for q in s if q >= 0)
    return cnt[0]


def distinct(T, bh, bw, bd):
    V, k, D, H, W, _ = T.shape
    gid = ((np.arange(D) // bd)[:, None, None] * 1000000 + (np.arange(H) // bh)[None, :, None] * 1000
           + (np.arange(W) // bw)[None, None, :])
    n = 0
    for v in range(V):
        for j in range(k):
            t = T[v, j]
            m = t >= 0
            n += np.unique(np.broadcast_to(gid[..., None], t.shape)[m] * 10000 + t[m]).size
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    cfg = SceneConfig(n_views=a.views)
    T = tap_ids(cfg, a.seed)
    kb = cfg.channels * 4 / 1e9
    samples = T[..., 0].size
    print(f"samples {samples}, with any tap inside {(T >= 0).any(-1).mean():.3f}, "
          f"unmerged {(T >= 0).sum() * kb:.2f} GB")
    print(f"run kernel (1 row, 1 pending column): {replay(T, 1, 1) * kb:.2f} GB")
    print(f"lean run kernel (+ right tap joins a pending tap when x does not advance): "
          f"{replay(T, 1, 'stay') * kb:.2f} GB")
    for rows in (1, 2, 4):
        print(f"{rows}-row blocks, 2 pending columns:     {replay(T, rows, 2) * kb:.2f} GB"
              f"   (ideal {distinct(T, rows, 8, 1) * kb:.2f} GB)")
    print(f"ideal 4x8 block over all planes:      {distinct(T, 4, 8, cfg.num_depth) * kb:.2f} GB")


if __name__ == "__main__":
    main()
