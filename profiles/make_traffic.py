"""profiles/traffic.json from ncu `--set full` captures: per BASELINE config, the DRAM bytes ONE launch of the dominant kernel
moves (dram__bytes_read.sum + dram__bytes_write.sum) -- what bench.py reports as `roofline.traffic` next to the algorithmic
bytes. Runs in the dev container (`ncu -i` needs no GPU):

    python profiles/make_traffic.py cfg2=gpurun_out/r2/prof_final_cfg2.ncu-rep:merge_all_tma cfg5=...:merge_single_view \
                                    cfg3=gpurun_out/r2/prof_final_cfg3.ncu-rep:joint_
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def rows_of(path):
    if path.endswith(".csv"):  # already exported on the GPU box (`ncu -i rep --page raw --csv`): the .ncu-rep stayed there
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def main():
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for spec in sys.argv[1:]:
        cfg, rest = spec.split("=", 1)
        path, pattern = rest.rsplit(":", 1)
        rows = rows_of(path)
        head, units = rows[0], rows[1]
        ni, ri, wi, ti = (head.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        per_kernel = {}
        for r in rows[2:]:
            if not re.search(pattern, r[ni]):
                continue
            name = re.sub(r"<.*|\(.*", "", r[ni]).replace("void ", "").strip()
            rd = float(r[ri].replace(",", "")) * UNITS[units[ri]]
            wr = float(r[wi].replace(",", "")) * UNITS[units[wi]]
            per_kernel.setdefault(name, []).append((rd, wr, r[ti] + " " + units[ti]))
        entry = data.setdefault(cfg, {})
        for name, caps in per_kernel.items():
            rd, wr, dur = caps[-1]  # the last (warm) capture
            entry[name] = {"bytes": int(rd + wr), "read": int(rd), "write": int(wr), "duration_under_ncu": dur,
                           "source": f"ncu --set full --clock-control none, {os.path.basename(path)} ({len(caps)} launch(es) captured)"}
        if cfg == "cfg3" and per_kernel:  # the fused loss is two kernels: report their sum under one key
            tot = sum(v["bytes"] for k, v in entry.items() if k.startswith("joint_") and isinstance(v, dict))
            entry["joint_focal_dice"] = {"bytes": int(tot), "source": "sum of joint_stats_kernel + joint_finalize_kernel + joint_grad_kernel"}
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
