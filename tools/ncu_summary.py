#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per captured launch, the
counters the roofline discussion in DESIGN.md uses.

    python tools/ncu_summary.py gpurun_out/prof_sweep.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
]


KERNEL_KEYS = (("sweep_fwd", "plane_sweep_fwd"), ("sweep_bwd", "plane_sweep_bwd"),
               ("backproject_fwd", "backproject_fwd"), ("backproject_bwd", "backproject_bwd"),
               ("depth_topk_fwd", "depth_topk_fwd"), ("depth_topk_bwd", "depth_topk_bwd"),
               ("unpack", "unpack"), ("pack", "pack"), ("prob_norm", "prob_norm_bwd"))    # unpack before pack


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * mult.get(unit, 1)


def traffic(paths, out_path):
    """profiles/traffic.json: per-launch DRAM bytes (read + write) of each kernel,
    mean over the captured launches -- bench.py's roofline.traffic."""
    import json
    acc = {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[2:]:
            for frag, key in KERNEL_KEYS:
                if frag in r[ki]:
                    acc.setdefault(key, []).append(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]))
                    break
    res = {k: int(sum(v) / len(v)) for k, v in acc.items()}
    # RED payload of the plane-sweep backward (bytes leaving the SMs towards L2: the gradient
    # scatter; the kernel has no other global writes) -- bench.py reports it next to the
    # measured fp32-RED ceiling
    red = []
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        if "l1tex__m_l1tex2xbar_write_bytes.sum" not in hdr:
            continue
        ki, wi = hdr.index("Kernel Name"), hdr.index("l1tex__m_l1tex2xbar_write_bytes.sum")
        red += [to_bytes(r[wi], units[wi]) for r in rows[2:] if "sweep_bwd" in r[ki]]
    if red:
        res["plane_sweep_bwd_red_payload"] = int(sum(red) / len(red))
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print(json.dumps(res))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--traffic":
        traffic(sys.argv[3:], sys.argv[2])
        return
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            print(f"## {path}: {r[ki][:90]}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"  {w:78s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
