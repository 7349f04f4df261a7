"""Compact per-launch table of an `ncu --set full` capture, for profiles/:

  ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_summary.py X.csv > profiles/X_summary.txt

Columns: duration (us, cold-cache, serialised: compare shares), warp instructions executed, issue-active %, warps
active %, registers, grid, DRAM read / write (MB), DRAM throughput % of peak, average threads per executed instruction
(SIMT lane use), shared-memory LSU wavefronts % of peak, L2 hit rate %, the top three stall reasons per issue."""
import csv
import sys


def f(row, idx, key, scale=1.0, nd=1):
    i = idx.get(key)
    if i is None or row[i] in ("", "no data"):
        return "-"
    try:
        return f"{float(row[i].replace(',', '')) * scale:.{nd}f}"
    except ValueError:
        return row[i]


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    unit = dict(zip(hdr, units))

    def mb(row, key):
        i = idx.get(key)
        if i is None or row[i] == "":
            return "-"
        v = float(row[i].replace(",", ""))
        u = unit[key].lower()
        v *= {"gbyte": 1e3, "mbyte": 1.0, "kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
        return f"{v:.1f}"

    def us(row):
        i = idx["gpu__time_duration.sum"]
        v = float(row[i].replace(",", ""))
        u = unit["gpu__time_duration.sum"].lower()
        return v * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u.replace("second", "s").replace("msecond", "ms"), 1.0) \
            if u in ("ms", "us", "ns", "s") else v * {"msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)

    print(f"# {path}")
    print("id  dur_us  warp_inst  issue%  warps%  regs  grid  dram_rd_MB  dram_wr_MB  dram%  thr/inst  smem_lsu%  l2hit%  kernel | top stalls per issue")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].replace("void <unnamed>::", "").replace("<unnamed>::", "")
        name = name.split("(")[0][:44]
        stalls = sorted(((float(v.replace(",", "")), k.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", ""))
                         for k, v in zip(hdr, r) if k.startswith("smsp__average_warps_issue_stalled_")
                         and k.endswith("_per_issue_active.ratio") and v not in ("", "no data")), reverse=True)[:3]
        st = ", ".join(f"{n} {v:.2f}" for v, n in stalls)
        print(f"{r[idx['ID']]:>2}  {us(r):7.1f}  {f(r, idx, 'smsp__inst_executed.sum', 1e-6, 2):>7}M  "
              f"{f(r, idx, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):>5}  "
              f"{f(r, idx, 'sm__warps_active.avg.pct_of_peak_sustained_active'):>5}  "
              f"{f(r, idx, 'launch__registers_per_thread', 1, 0):>4}  {f(r, idx, 'launch__grid_size', 1, 0):>6}  "
              f"{mb(r, 'dram__bytes_read.sum'):>8}  {mb(r, 'dram__bytes_write.sum'):>8}  "
              f"{f(r, idx, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):>5}  "
              f"{f(r, idx, 'smsp__thread_inst_executed_per_inst_executed.ratio'):>5}  "
              f"{f(r, idx, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):>5}  "
              f"{f(r, idx, 'lts__t_sector_hit_rate.pct'):>5}  {name} | {st}")


if __name__ == "__main__":
    main(sys.argv[1])
