"""Per-kernel opcode histogram of the in-tree libtb200.so (replaces the md5 fingerprint of round 1): which kernels use TMA
(UTMALDG / UTMASTG), mbarriers (SYNCS), vector reductions (REDG), 128-bit shared / global accesses, MUFU, warp match etc.
Runs in the dev container (cuobjdump needs no GPU):

    python profiles/sass_opcodes.py > profiles/sass_opcodes_r2.json     # + a markdown table on stderr
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_toolbelt_b200", "libtb200.so")
TRACK = ["UTMALDG", "UTMASTG", "SYNCS", "REDG", "RED.", "ATOMG", "ATOMS", "LDS.128", "LDS", "STS", "LDG.E.128", "LDG", "STG.E.128",
         "STG.E.EF.128", "STG", "MUFU", "MATCH", "SHFL", "BAR.SYNC", "DADD", "FFMA", "HMMA", "UTCMMA", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for t in TRACK:
            if op.startswith(t) if not t.endswith(".") else op.startswith(t):
                kernels[cur][t] += 1
    names = demangle(list(kernels))
    short = {}
    for k, v in names.items():
        s = re.sub(r"\(.*", "", v).replace("void ", "").replace("tb::", "")
        short[k] = s
    result = {"library": os.path.relpath(LIB, ROOT), "arch": sorted(arch), "kernels": {}}
    for k, c in kernels.items():
        result["kernels"][short[k]] = {t: c[t] for t in ["_total"] + TRACK if c[t]}
    print(json.dumps(result, indent=1))
    # families (template instances summed) as a table
    fam = collections.OrderedDict()
    for k, c in result["kernels"].items():
        f = re.sub(r"<.*", "", k)
        agg = fam.setdefault(f, collections.Counter())
        agg["instances"] += 1
        for t, n in c.items():
            agg[t] += n
    cols = ["instances", "_total", "UTMALDG", "UTMASTG", "SYNCS", "REDG", "ATOMG", "LDS.128", "LDG.E.128", "STG.E.128", "STG.E.EF.128", "MUFU", "MATCH", "DADD", "HMMA", "UTCMMA"]
    sys.stderr.write("| kernel family | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols) + "\n")
    for f, agg in fam.items():
        sys.stderr.write(f"| {f} | " + " | ".join(str(agg[c]) for c in cols) + " |\n")


if __name__ == "__main__":
    main()
