#!/usr/bin/env python
"""bench.py -- scenes/s of the MVSDet plane-sweep + depth-top-k + voxel
back-projection path (forward + backward) on B200, with the HBM roofline of the
dominant kernel and the CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): one synthetic ScanNet-shaped scene per step
-- V=20 views, 320x240 images -> 60x80 FPN features x 256 channels, D=12 depth
planes, k=2 neighbours, T=3 hypotheses, 40x40x16 voxels -- forward+backward,
bf16 features / fp32 accumulation, fp32 variance volume.  The cost-regularisation
3-D U-Net that sits between the stages is not part of the path:
its output and the gradient it returns to the variance volume are synthetic
inputs of the step.  A "step" = ScenePipeline.step(): pack, sweep fwd, top-k
fwd, voxels fwd, voxels bwd, pn bwd, top-k bwd, sweep bwd, unpack.

N > 1 (torchrun): scenes are independent, so every rank runs its own scenes
(weak scaling, no data-path collective); the timed region is bracketed by a
barrier + synchronize and the slowest rank's device time is used.

`--impl reference` times the CPU oracle (the restatement of the reference's
PyTorch path, oracle/mvsdet_oracle.py) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "scenes/sec for plane-sweep+voxel backproj fwd+bwd; % of HBM roofline"
UNIT = "scenes/s"
NBUF = 3   # rotating scene buffer sets: a set is re-read after >= 2 other steps (> 4 GB of traffic)


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 100):
        self.gpu = gpu_index
        self.samples = []          # (t, sm, max, [reasons])
        self.proc = None
        self.period_ms = period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        threading.Thread(target=self._reader, daemon=True).start()

    def _reader(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            reasons = [n for n, p in zip(names, parts[4:8]) if p.lower().startswith("active")]
            self.samples.append((time.time(), sm, mx, reasons))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({r for s in inside for r in s[3]})
        return {"sm_mhz": statistics.median(s[1] for s in inside),
                "sm_max_mhz": max(s[2] for s in inside), "reasons": reasons,
                "samples": len(inside)}


# ---------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle on the host cores
# ---------------------------------------------------------------------------
def cpu_reference_step(scene, n_sample_views: int):
    """fwd+bwd of the oracle restricted to the first ``n_sample_views``
    reference views (plane sweep, top-k and back-projection all scale linearly
    in the number of reference views; neighbours are still drawn from all V)."""
    from oracle import mvsdet_oracle as O
    cfg = scene["cfg"]
    feature = scene["feature"].clone().requires_grad_(True)
    vs = n_sample_views
    cost_out = scene["cost_out"][:vs].clone().requires_grad_(True)
    res = O.hot_path(feature, scene["img_meta"], lambda var: cost_out,
                     near_far_range=cfg.near_far_range, num_depth=cfg.num_depth, topk=cfg.topk,
                     n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size, stride=cfg.stride,
                     training=True, view_subset=vs)
    torch.autograd.backward([res["variance"], res["volume_mean"]],
                            [scene["g_variance"][:vs], scene["g_volume_mean"]])
    return feature.grad, cost_out.grad


def time_cpu_reference(cfg, steps: int, warmup: int, budget_s: float):
    from mvsdet_b200.scene import make_scene
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    scene = make_scene(cfg, seed=0)
    v = cfg.n_views
    vs = 1
    t0 = time.perf_counter()
    cpu_reference_step(scene, vs)                      # calibration, also a warm-up
    t_one = time.perf_counter() - t0
    total_steps = max(1, steps + warmup)
    for cand in (4, 2):
        if cand <= v and t_one * cand * total_steps <= budget_s:
            vs = cand
            break
    for _ in range(max(0, warmup - 1)):
        cpu_reference_step(scene, vs)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        cpu_reference_step(scene, vs)
        times.append(time.perf_counter() - t0)
    t_step = sum(times) / len(times)
    scenes_per_s = 1.0 / (t_step * v / vs)
    sample = (f"{vs} of {v} reference views per step (plane sweep + top-k + back-projection "
              f"fwd+bwd of the oracle, fp32, torch {torch.__version__} CPU), scaled x{v / vs:g} "
              f"to a scene; {len(times)} timed steps, mean {t_step:.2f} s/step")
    return scenes_per_s, t_step, threads, sample, len(times)


def run_reference_arm(args, cfg, cfg_json):
    rank, local, world = _dist_env()
    if rank != 0:
        return
    val, t_step, threads, sample, nsteps = time_cpu_reference(cfg, args.steps, args.warmup, 150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": nsteps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg_json,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------
def run_own_arm(args, cfg, cfg_json):
    import torch.distributed as dist
    from mvsdet_b200 import _lib
    from mvsdet_b200.hotpath import MVSDetHotPath
    from mvsdet_b200.pipeline import ScenePipeline
    from mvsdet_b200.scene import make_scene

    rank, local, world = _dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: mvsdet_b200 has no CPU path "
                           "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    feat_dtype = torch.bfloat16 if args.feature_dtype == "bf16" else torch.float32
    mod = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                        stride=cfg.stride)
    pipes, graphs = [], []
    for b in range(NBUF):
        scene = make_scene(cfg, seed=1000 * rank + b)
        pipe = ScenePipeline(cfg, dev, feature_dtype=feat_dtype)
        pipe.set_geometry(mod.geometry(scene["img_meta"], dev))
        pipe.load_scene(scene)
        pipes.append(pipe)
        if b == 0:
            host_scene = scene
    torch.cuda.synchronize()
    launches_before = _lib.launch_count()
    pipes[0].step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - launches_before
    use_graph = not args.no_graph
    if use_graph:
        graphs = [p.capture() for p in pipes]

    def one_step(i):
        if use_graph:
            graphs[i % NBUF].replay()
        else:
            pipes[i % NBUF].step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs --------------------------------
    for i in range(args.warmup):
        one_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        one_step(i)
    e1.record()
    barrier()
    wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the public pipeline call -------------
    p0 = pipes[0]
    host = p0.host_buffers()
    host["feature"].copy_(host_scene["feature"])
    host["cost_out"].copy_(host_scene["cost_out"])
    host["g_volume_mean"].copy_(host_scene["g_volume_mean"].reshape(host["g_volume_mean"].shape))
    host["g_variance"].copy_(host_scene["g_variance"].permute(0, 2, 3, 4, 1))
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    from mvsdet_b200.pipeline import HostPipelinedRunner
    runner = HostPipelinedRunner(pipes, graphs if use_graph else None)
    host_in = {n: host[n] for n in p0.INPUTS}
    runner.run(host_in, 2 * NBUF)                       # warm-up: every buffer set, twice
    barrier()
    e2, e3 = runner.run(host_in, e2e_steps)             # H2D / compute / D2H overlapped across steps
    barrier()
    wall2 = time.time()
    ms_e2e = e2.elapsed_time(e3)
    # the same call without overlap (one stream), for reference
    for _ in range(2):
        p0.run_host(graphs[0] if use_graph else None)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in range(min(e2e_steps, 10)):
        p0.run_host(graphs[0] if use_graph else None)
    e5.record()
    barrier()
    ms_e2e_serial = e4.elapsed_time(e5) / min(e2e_steps, 10)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * e2e_steps / (ms_e2e * 1e-3)
    sampler.stop()
    clocks = sampler.summary(wall0, wall1)
    # sanity of the e2e result: compare the host copy of the outputs with the device ones
    e2e_ok = bool(torch.equal(host["count"], p0.count.cpu()))

    # ---- per-kernel device times (CUDA events around each launch, same stream)
    timers = {}
    ksteps = max(3, min(args.steps, 30))
    for i in range(ksteps):
        pipes[i % NBUF].step(timers)
    torch.cuda.synchronize()
    peak, peak_src = _peaks()
    abytes = p0.algorithmic_bytes()
    kernels = {}
    for name, evs in timers.items():
        ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
        kernels[name] = {"ms": round(ms, 5), "algorithmic_mb": round(abytes[name] / 1e6, 2),
                         "gbs": round(abytes[name] / (ms * 1e-3) / 1e9, 1)}
    top = max(kernels, key=lambda n: kernels[n]["ms"])
    total_ms = sum(k["ms"] for k in kernels.values())
    achieved = kernels[top]["gbs"]
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes[top],
                "kernel_ms": kernels[top]["ms"],
                "share_of_step": round(kernels[top]["ms"] / total_ms, 3),
                "path_achieved_gbs": round(sum(abytes.values()) / (ms_per_step * 1e-3) / 1e9, 1),
                "path_frac": round(sum(abytes.values()) / (ms_per_step * 1e-3) / 1e9 / peak, 4)}
    ncu_traffic = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(ncu_traffic):
        try:
            with open(ncu_traffic) as fh:
                tj = json.load(fh)
            roofline["traffic"] = tj.get(top)
            red = tj.get(top + "_red_payload")
            if red:
                # the scatter kernel's second stream: fp32 REDs into L2.  Ceiling measured by
                # tools/microbench_red.cu on this pool's B200 (profiles/r01_b_microbench_red.txt).
                # Up to round 1e this stream (3.6 GB at > 80 % of the ceiling) was the limiter; with
                # the row hand-off it is 2.7 GB and the kernel is bound by instructions per pixel at
                # 12 warps/SM (issue-active 49 %) -- reported for that reason.
                red_gbs = red / (kernels[top]["ms"] * 1e-3) / 1e9
                roofline["limiter"] = {"what": "fp32 red.global.add.v4 payload into L2 (ncu l1tex2xbar write bytes)",
                                       "red_payload_bytes": red, "achieved_gbs": round(red_gbs, 1),
                                       "ceiling_gbs": 5700.0, "frac": round(red_gbs / 5700.0, 3),
                                       "binding": red_gbs / 5700.0 > 0.8}
        except Exception:
            pass

    # ---- the autograd drop-in (MVSDetHotPath + torch.autograd, host geometry every call):
    # what a maintainer gets from INTEGRATION.md level 1, wall clock incl. host work
    module_api = None
    if world == 1:
        hot = MVSDetHotPath(cfg.n_voxels, cfg.voxel_size, cfg.near_far_range, cfg.num_depth, cfg.topk,
                            stride=cfg.stride, feature_dtype=feat_dtype)
        m_feat = p0.feature.detach().clone().requires_grad_(True)
        m_cost = p0.cost_out.detach().clone().requires_grad_(True)
        m_gvar = p0.g_variance.permute(0, 4, 1, 2, 3)       # channels_last_3d view, as cuDNN returns it
        m_gvol = p0.g_volume_mean.view(cfg.channels, *cfg.n_voxels)

        def module_step():
            res = hot(m_feat, host_scene["img_meta"], cost_regularization=lambda var: m_cost)
            torch.autograd.backward([res["variance"], res["volume_mean"]], [m_gvar, m_gvol])
            m_feat.grad = None
            m_cost.grad = None

        for _ in range(5):
            module_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_mod = 30
        for _ in range(n_mod):
            module_step()
        torch.cuda.synchronize()
        ms_mod = (time.perf_counter() - t0) / n_mod * 1e3
        module_api = {"value": 1e3 / ms_mod, "unit": UNIT, "ms_per_scene": round(ms_mod, 4), "steps": n_mod,
                      "what": "MVSDetHotPath forward + torch.autograd backward, scene geometry "
                              "recomputed on the host every call, caching allocator, no CUDA graph"}
        del hot, m_feat, m_cost

    if rank == 0:
        cpu_line = None
        if world == 1 and not args.no_cpu_baseline:
            val, t_step, threads, sample, _ = time_cpu_reference(cfg, 2, 1, 25.0)
            cpu_line = {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 features, f32 accumulate" if feat_dtype == torch.bfloat16 else "f32",
            "data": "synthetic", "config": cfg_json,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": p0.h2d_bytes(),
                    "d2h_bytes_per_step": p0.d2h_bytes(), "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps, "outputs_match_device": e2e_ok,
                    "overlap": "H2D / compute / D2H on three streams over %d buffer sets" % NBUF,
                    "ms_per_step_serial": ms_e2e_serial},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "cuda_graph": use_graph,
            "roofline": roofline, "kernels": kernels,
        }
        if module_api is not None:
            line["module_api"] = module_api
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--views", type=int, default=20)
    ap.add_argument("--feature-dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=30)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl == "own" else 1)

    from mvsdet_b200.scene import SceneConfig
    cfg = SceneConfig(n_views=args.views)
    hf, wf = cfg.feat_hw
    cfg_json = {
        "workload": "BASELINE.json configs[1]: mvsdet_res50_2x_low_res, 1 scene/step, fwd+bwd of "
                    "plane-sweep variance + depth top-k + voxel back-projection",
        "views": cfg.n_views, "channels": cfg.channels, "depth_planes": cfg.num_depth,
        "feature_map": [hf, wf], "image": [cfg.pad_shape[0], cfg.pad_shape[1]],
        "neighbors": cfg.num_neighbors, "topk": cfg.topk, "voxels": list(cfg.n_voxels),
        "feature_in": "fp32 NCHW (FPN layout), packed in-step to channels-last "
                      + ("bf16" if args.feature_dtype == "bf16" else "fp32"),
        "variance": "fp32 channels_last_3d", "scenes_per_step_per_gpu": 1,
        "parallelism": f"scene-parallel x{args.gpus} (no collective)",
        "l2": f"working set 2.6 GB/step >> 126 MB L2; {NBUF} rotating scene buffer sets",
        "cost_reg_net": "excluded (synthetic cost_out / g_variance inputs)",
    }
    if args.impl == "reference":
        run_reference_arm(args, cfg, cfg_json)
    else:
        run_own_arm(args, cfg, cfg_json)


if __name__ == "__main__":
    main()
