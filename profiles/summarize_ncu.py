"""Turn ncu captures into the markdown tables kept under profiles/ (run in the dev container; `ncu -i` needs no GPU).

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv          # per-kernel launch counts / time / share
    python profiles/summarize_ncu.py full gpurun_out/prof.ncu-rep [regex]      # --set full: the metrics we track
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def _rows(text):
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = _rows(open(path).read())
    head = rows[0]
    name_i, metric_i, value_i = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    unit_i = head.index("Metric Unit")
    total, count = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[metric_i] != "gpu__time_duration.sum":
            continue
        v = float(r[value_i].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[unit_i], 1.0)
        total[r[name_i]] += v
        count[r[name_i]] += 1
    whole = sum(total.values())
    print("| kernel | launches | total (ns) | avg (us) | share |\n|---|---|---|---|---|")
    for k in sorted(total, key=total.get, reverse=True):
        print(f"| {k[:80]} | {count[k]} | {total[k]:.0f} | {total[k] / count[k] / 1e3:.1f} | {100 * total[k] / whole:.1f} % |")


def full(path, pattern=None):
    if path.endswith(".csv"):  # exported on the GPU box with `ncu -i rep --page raw --csv`
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = _rows(text)
    head, units = rows[0], rows[1]
    name_i = head.index("Kernel Name")
    seen = OrderedDict()
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[name_i]):
            continue
        seen[r[name_i]] = r  # keep the LAST capture of each kernel (warm)
    for name, r in seen.items():
        print(f"### `{name[:110]}`\n\n| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in head:
                i = head.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
