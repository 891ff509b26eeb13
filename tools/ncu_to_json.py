"""profiles/ncu_current.json from a `ncu --set full` report of `bench.py` (tools/gpu_ncu.sh): the per-launch counters
bench.py quotes in its `roofline` object (DRAM traffic, executed warp instructions, issue-slot utilisation, resident
warps), tied to the CUDA sources they were captured from by eogs2_b200.build.source_hash().  bench.py refuses to quote
the file when the hash does not match the sources it runs.

    python tools/ncu_to_json.py gpurun_out/<tag>/prof.ncu-rep [more.ncu-rep ...]  > profiles/ncu_current.json
"""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from eogs2_b200.build import source_hash          # noqa: E402

KEYS = {
    "duration_ms": ("gpu__time_duration.sum", 1.0),
    "inst_executed": ("smsp__inst_executed.sum", 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "dram_read": ("dram__bytes_read.sum", 1.0),
    "dram_write": ("dram__bytes_write.sum", 1.0),
    "shared_wavefronts": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1.0),
    "shared_bank_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
    "fma_pipe_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "alu_pipe_pct": ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "lsu_pipe_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
    "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    "registers": ("launch__registers_per_thread", 1.0),
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    kernels = {}
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
            name = re.sub(r"[<(].*$", "", name).split("::")[-1]
            k = {}
            for key, (metric, _) in KEYS.items():
                if metric in ix and r[ix[metric]] not in ("", "no data"):
                    v = float(r[ix[metric]].replace(",", ""))
                    k[key] = v * UNIT_SCALE.get(units[ix[metric]], 1.0)
            k["dram_bytes"] = int(k.pop("dram_read", 0) + k.pop("dram_write", 0))
            if "warps_active_pct" in k:
                k["warps_active_per_sm"] = round(k["warps_active_pct"] * 64 / 100, 2)
            kernels[name] = k                    # the last captured launch of a kernel wins
    print(json.dumps({"source_hash": source_hash(), "report": ", ".join(Path(p).name for p in sys.argv[1:]),
                      "how": "ncu --set full --clock-control none, one launch per kernel during `bench.py --steps 2 --warmup 3`",
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
