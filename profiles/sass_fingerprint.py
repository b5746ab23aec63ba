"""Per-kernel fingerprint of libtb200.so's SASS (instruction text only, addresses/encodings stripped).

    python profiles/sass_fingerprint.py write profiles/sass_fingerprint_r1.json     # after a GPU-validated build
    python profiles/sass_fingerprint.py diff  profiles/sass_fingerprint_r1.json     # later: which kernels changed?

Used at the end of round 1 to prove that code added after the last GPU run (new kernels, dead-code removal) left every
GPU-validated kernel instruction-identical. Runs in the dev container (cuobjdump needs no GPU)."""
import hashlib
import json
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytorch_toolbelt_b200", "libtb200.so")


def fingerprint(lib=LIB):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", text)[1:]:
        name, body = part.split("\n", 1)
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(.*?);", body)
        out[name.strip()] = {"md5": hashlib.md5("\n".join(ins).encode()).hexdigest(), "instructions": len(ins)}
    return out


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    now = fingerprint()
    if mode == "write":
        json.dump(now, open(path, "w"), indent=0, sort_keys=True)
        print(f"{len(now)} kernels -> {path}")
    else:
        then = json.load(open(path))
        changed = sorted(k for k in now if k in then and now[k]["md5"] != then[k]["md5"])
        print(json.dumps({"kernels": len(now), "changed": changed, "new": sorted(set(now) - set(then)),
                          "removed": sorted(set(then) - set(now))}, indent=1))
