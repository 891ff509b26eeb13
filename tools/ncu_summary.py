"""Turn ncu output brought back in gpurun_out/ into the text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/<tag>/launches.csv  > profiles/<name>_launches.txt
    python tools/ncu_summary.py full     gpurun_out/<tag>/blend.ncu-rep > profiles/<name>_blend_full.txt

`launches`: per-kernel count, total and mean device time and SHARE of the listed launches (ncu's
--metrics gpu__time_duration.sum pass: cold-cache, serialised — shares are meaningful, absolutes are not).
`full`: the counters DESIGN.md / bench.py quote (DRAM bytes, issue-slot utilisation, pipes, shared-memory
wavefronts, atomics, occupancy) for every kernel captured with --set full.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

FULL_KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_xu.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


def launches(path: str) -> None:
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v_us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v_us
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms listed (cold-cache, serialised)")
    print(f"{'share':>7} {'n':>5} {'total_us':>11} {'mean_us':>10}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.2f}% {n:5d} {t:11.1f} {t / n:10.1f}  {k}")


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {path} (ncu --set full --clock-control none), one block per captured launch")
    for d in data:
        print(f"\n== {short(d[hdr.index('Kernel Name')])}  grid {d[hdr.index('Grid Size')]} block {d[hdr.index('Block Size')]}")
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:86s} {d[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
