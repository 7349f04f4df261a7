"""Summarise an `ncu --set full` raw page (`ncu -i X.ncu-rep --page raw --csv`) of the raster kernels:
per launch DRAM traffic, pipe utilisation, occupancy, stall reasons.  Writes profiles/raster_bwd_traffic.json
(read by bench.py for `roofline.traffic`) and prints a short table for profiles/.

  python tools/ncu_traffic.py gpurun_out/r01f_ncu_full_raster_bwd.csv [kernel-substring] [out.json] [workload] [command]
"""
import csv
import json
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__grid_size",
    "launch__block_size",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * m
    return v


def main(path, needle="raster_bwd_kernel<4", out=None, workload="cfg2", command="bench.py --mode eager"):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    rows = [dict(zip(header, r)) for r in rd if r]
    unit_of = dict(zip(header, units))
    picked = [r for r in rows if needle in r.get("Kernel Name", "")] or rows
    summary = []
    for r in picked:
        d = {"kernel": r.get("Kernel Name", "")[:90], "id": r.get("ID")}
        for k in WANT:
            if k in r and r[k] != "":
                d[k] = to_bytes(r[k], unit_of.get(k, "")) if "bytes" in k else float(r[k].replace(",", ""))
        # stall breakdown: every warp-state sample ratio
        stalls = {k.split("smsp__average_warps_issue_stalled_")[-1].replace("_per_issue_active.ratio", ""): float(v.replace(",", ""))
                  for k, v in r.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
        d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        summary.append(d)
    for d in summary:
        print(json.dumps(d))
    if out and summary:
        d = summary[0]
        rec = {
            "kernel": d["kernel"],
            "dram_bytes_per_launch": d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0),
            "dram_bytes_read": d.get("dram__bytes_read.sum"),
            "dram_bytes_write": d.get("dram__bytes_write.sum"),
            "ncu_duration_us": d.get("gpu__time_duration.sum"),
            # what binds a kernel that is not HBM-bound: issue slots and pipes, from the same launch
            "issue": {
                "warp_instructions": d.get("smsp__inst_executed.sum"),
                "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": d.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "alu_pipe_pct": d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "xu_pipe_pct": d.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "lsu_pipe_pct": d.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": d.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "registers_per_thread": d.get("launch__registers_per_thread"),
                "top_stalls_per_issue": d.get("stalls_per_issue"),
            },
            "source": f"ncu --set full --clock-control none, launch id {d['id']} of `{command}` on {workload} ({path.split('/')[-1]})",
        }
        with open(out, "w") as f:
            json.dump(rec, f, indent=1)
        print("wrote", out)


if __name__ == "__main__":
    main(*sys.argv[1:])
