"""Turn ncu outputs brought back in gpurun_out/ into the small JSON summaries kept under profiles/.

  python tools/summarize_ncu.py launches <launches.csv>              -> per-kernel launch counts / mean us / share of the step
  python tools/summarize_ncu.py full <report.ncu-rep> [kernel-regex] -> selected `--set full` metrics per captured launch

Needs `ncu` on PATH (reads the report with `ncu -i ... --page raw --csv`); no GPU required."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path, skip_first_steps=True):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    per = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[4])
        per.setdefault(name, []).append(float(r[14]) / 1e3)
    out = {k: {"launches": len(v), "mean_us": round(sum(v) / len(v), 2), "total_us": round(sum(v), 2)} for k, v in per.items()}
    tot = sum(v["total_us"] for k, v in out.items() if k.startswith("a0_k"))
    for k, v in out.items():
        v["share_of_a0_kernel_time"] = round(v["total_us"] / tot, 4) if k.startswith("a0_k") else None
    return out


def full(rep, pattern=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[name_col]):
            continue
        d = {"kernel": r[name_col]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        out.append(d)
    return out


if __name__ == "__main__":
    mode = sys.argv[1]
    res = launches(sys.argv[2]) if mode == "launches" else full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    json.dump(res, sys.stdout, indent=1)
    print()
