"""bench.py's reference arm (the oracle port timed on host cores) must keep the driver's JSON contract; it is the one
bench leg that can run without a GPU. Also: under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus",
                           env_extra.get("WORLD_SIZE", "1"), "--steps", "1", "--warmup", "0", "--cpu-sample-tiles", "1"],
                          capture_output=True, text=True, timeout=600, cwd=REPO, env=env)


def test_reference_arm_prints_one_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    baseline = json.load(open(os.path.join(REPO, "BASELINE.json")))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "MP/s" and "megapixels/sec" in d["metric"]
    assert str(baseline.get("metric", "")).split()[0].lower() in d["metric"].lower() or "megapixel" in d["metric"].lower()
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "cfg2" in d["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    """bench.py's `roofline.traffic` for the contract config is read from profiles/traffic.json (made from the ncu capture
    by profiles/make_traffic.py) and must sit within 1 % of the algorithmic bytes of the merge kernel: nothing is re-read."""
    sys.path.insert(0, REPO)
    import importlib

    bench = importlib.import_module("bench")
    traffic, source = bench.measured_traffic("cfg2", "merge_all_tma_kernel")
    assert traffic is not None and "ncu" in source
    cfg = bench.CONFIGS["cfg2"]
    H, W, T, S, K = (cfg[k] for k in ("H", "W", "T", "S", "K"))
    n_side = -(-(H - T) // S) + 1
    tiles = n_side * n_side
    algorithmic = tiles * 8 * K * T * T * 4 + K * H * W * 4  # 8 D4 views of every tile's logits in, the cropped canvas out
    assert tiles == 361 and abs(traffic / algorithmic - 1.0) < 0.01
    assert bench.measured_traffic("cfg2", "no_such_kernel") == (None, None)
