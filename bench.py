#!/usr/bin/env python
"""bench.py -- megapixels/sec of merged output for the tiled-inference hot path (BASELINE.json metric).

    python bench.py                                   # this engine, cfg 2 (the contract line), 1 GPU
    python bench.py --config cfg5                     # another BASELINE config (cfg1 | cfg2 | cfg3 | cfg4 | cfg5)
    python bench.py --impl reference                  # the UNMODIFIED reference (baseline/_ref) on the host cores
    torchrun --nproc-per-node 8 bench.py --gpus 8     # ragged strips over 8 GPUs (any tiled config)

Tiled configs (cfg 1, 2, 4, 5): one "step" = one pass of the hot path over one image:
    slice(+augment) kernel: image -> [A*n,Cin,T,T] f32      and      fused de-augment / merge / normalise / crop kernel:
    [A*n,K,T,T] f32 predictions -> [K,H,W] f32.
The model between the two kernels is a black box that is NOT part of the path: predictions are produced once before
timing by the stand-in model of SURVEY 8(d) (`value`, engine-only); `with_model` adds the stand-in kernel; `e2e` goes
through the public API from a pinned HOST image to a pinned HOST result (H2D + kernels + model + D2H inside the timed
region), `e2e_labels` the same with the uint8 class map fused into the merge epilogue (1 byte per pixel downloaded).
Inputs per step are far larger than the 126 MB L2 (cfg 1 excepted: see config.l2), so no explicit flush is needed.
cfg 3: one step = BinaryFocalLoss + DiceLoss forward+backward on 32x4x1024x1024 logits through ONE fused two-pass kernel.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

UNIT = "MP/s"
CONFIGS = {
    # BASELINE.json configs[0..4]; geometry per SURVEY section 8
    "cfg1": dict(H=2048, W=2048, Cin=3, T=256, S=128, K=3, tta=None, weight="pyramid", analytic=False, ref_tiles=None, cpu_tiles=225),
    "cfg2": dict(H=5000, W=5000, Cin=3, T=512, S=256, K=4, tta="d4", weight="pyramid", analytic=False, ref_tiles=None, cpu_tiles=64),
    "cfg4": dict(H=10000, W=10000, Cin=3, T=512, S=256, K=8, tta="flips", weight="pyramid", analytic=False, ref_tiles=78, cpu_tiles=39),
    "cfg5": dict(H=20000, W=20000, Cin=1, T=1024, S=512, K=1, tta=None, weight="gaussian", analytic=True, ref_tiles=78, cpu_tiles=39),
}
METRICS = {
    "cfg1": "megapixels/sec merged (2048x2048, 256 tiles, no TTA)",
    "cfg2": "megapixels/sec merged (5000x5000, 512 tiles, d4 TTA)",
    "cfg4": "megapixels/sec merged (10000x10000, 512 tiles, flips TTA)",
    "cfg5": "megapixels/sec merged (20000x20000, 1024 tiles, no TTA)",
    "cfg3": "Gelem/sec BinaryFocalLoss+DiceLoss fwd+bwd (32x4x1024x1024)",
}
WORKLOADS = {
    "cfg1": "cfg1 (BASELINE configs[0]): 2048x2048x3 u8 -> 225 tiles 256/128 -> no TTA -> K=3 merge, pyramid weight",
    "cfg2": "cfg2 (BASELINE configs[1]): 5000x5000x3 u8 -> 361 tiles 512/256 -> d4 TTA (8 views) -> K=4 logits merge, pyramid weight",
    "cfg4": "cfg4 (BASELINE configs[3]): 10000x10000x3 u8 -> 1521 tiles 512/256 -> fliplr+flipud TTA (3 views) -> K=8 merge, pyramid weight",
    "cfg5": "cfg5 (BASELINE configs[4]): 20000x20000x1 u8 -> 1521 tiles 1024/512 -> no TTA -> K=1 merge, gaussian weight (sigma=T/4)",
}


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(cfg_name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu captures (profiles/traffic.json)."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as f:
            entry = json.load(f).get(cfg_name, {}).get(kernel)
        return (entry["bytes"], entry["source"]) if entry else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


from pytorch_toolbelt_b200.utils.clocks import ClockSampler  # noqa: E402


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own torch-CPU path. Preferred: the UNMODIFIED reference installed under
# baseline/_ref (baseline/install_ref.sh; git-ignored, shipped to the GPU box), driven through ImageSlicer / tta.* /
# TileMerger by baseline/reference_arm.py (kind "reference"). Fallback when that install is absent: the oracle's
# restatement (kind "port"; bit-identical results, tests/test_reference_install.py).
# --------------------------------------------------------------------------------------------------------------
def reference_loop():
    """-> (callable(image, T, S, K, weight, group, batch_size, timings, max_tiles, device), kind)"""
    try:
        from baseline import reference_arm as RA

        if RA.available():
            RA.load()

            def run(image, tile, step, K, weight="pyramid", group=None, batch_size=8, timings=None, max_tiles=None, device="cpu"):
                return RA.tiled_inference(image, tile, step, K, weight=weight, group=group, batch_size=batch_size,
                                          timings=timings, max_tiles=max_tiles, device=device)

            return run, "reference"
    except Exception as e:  # a broken install must not take the bench down: fall back to the port and say so
        sys.stderr.write(f"[bench] reference install unusable ({e!r}); using the oracle port\n")
    from oracle import pipeline as OP

    def run_port(image, tile, step, K, weight="pyramid", group=None, batch_size=8, timings=None, max_tiles=None, device="cpu"):
        return OP.tiled_inference_cpu(image, tile, step, K, weight=weight, group=group or "none", batch_size=batch_size,
                                      timings=timings, max_tiles=max_tiles, device=device)

    return run_port, "port"


_BEST_THREADS = {}


def best_cpu_threads(cores: int) -> int:
    """torch-CPU intra-op threads that maximise the reference loop's throughput on this host: the loop is made of many
    small ATen ops, which get SLOWER with very wide thread pools, so 'all cores' would understate the reference. Probe a
    few pool sizes on a tiny sample and keep the fastest."""
    if cores in _BEST_THREADS:
        return _BEST_THREADS[cores]
    loop, _ = reference_loop()
    rng = np.random.default_rng(1)
    image = rng.integers(0, 256, (1536, 1536, 3)).astype(np.uint8)
    best, best_t = cores, float("inf")
    for th in sorted({t for t in (8, 16, 32, 64, cores) if t <= cores} or {cores}):
        torch.set_num_threads(th)
        loop(image, 512, 256, 4, group="d4", batch_size=8, max_tiles=8)
        dt = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            loop(image, 512, 256, 4, group="d4", batch_size=8, max_tiles=25)
            dt = min(dt, time.perf_counter() - t0)
        if dt < best_t:
            best, best_t = th, dt
    _BEST_THREADS[cores] = best
    return best


def synthetic_image(cfg):
    rng = np.random.default_rng(0)
    return rng.integers(0, 256, (cfg["H"], cfg["W"], cfg["Cin"])).astype(np.uint8)


def cpu_reference_run(cfg, max_tiles=None, threads=0):
    """One pass of the reference loop over the first `max_tiles` tiles (None = all) of the config.
    -> (engine-only seconds for the WHOLE config, with-model seconds, wall seconds of this run, info)."""
    loop, kind = reference_loop()
    cores = host_cores()
    threads = threads or best_cpu_threads(cores)
    torch.set_num_threads(threads)
    image = synthetic_image(cfg)
    from pytorch_toolbelt_b200.inference.tiles import ImageSlicer

    n_total = len(ImageSlicer(image.shape, cfg["T"], cfg["S"]).crops)
    n = n_total if max_tiles is None else min(max_tiles, n_total)
    timings = {}
    t0 = time.perf_counter()
    loop(image, cfg["T"], cfg["S"], cfg["K"], weight=cfg["weight"], group=cfg["tta"], batch_size=8, timings=timings, max_tiles=n)
    wall = time.perf_counter() - t0
    frac = n / n_total
    per_tile = timings["augment"] + timings["deaugment"] + timings["integrate"]
    fixed = timings["split"] + timings["merge"]  # whole-image stages: always complete
    t_engine = fixed + per_tile / frac
    t_with_model = t_engine + timings["model"] / frac
    how = ("the installed reference (baseline/_ref: ImageSlicer / tta.* / TileMerger on torch-CPU)" if kind == "reference"
           else "oracle/pipeline.py (torch-CPU restatement of the reference loop)")
    sample = (f"all {n_total} tiles" if n == n_total else
              f"first {n} of {n_total} tiles, per-tile stages scaled x{1 / frac:.2f}, whole-image split+normalise in full")
    info = {"cores": threads, "host_cores": cores, "kind": kind, "sample": f"{sample} through {how}; wall {wall:.1f}s",
            "stage_seconds_sample": {k: round(v, 3) for k, v in timings.items()}, "cpu_model": _cpu_model(),
            "extrapolated": n != n_total}
    return t_engine, t_with_model, wall, info


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    if args.config == "cfg3":
        return run_reference_losses(args)
    cfg = CONFIGS[args.config]
    mp = cfg["H"] * cfg["W"] / 1e6
    budget_s = 150.0  # keep the whole arm within a few minutes whatever --steps says
    max_tiles = args.ref_tiles if args.ref_tiles else cfg["ref_tiles"]
    for _ in range(max(args.warmup, 0)):
        cpu_reference_run(cfg, max_tiles=8)
    runs, spent = [], 0.0
    want = max(args.steps, 1)
    while len(runs) < want:
        r = cpu_reference_run(cfg, max_tiles=max_tiles)
        runs.append(r)
        spent += r[2]
        if spent + r[2] > budget_s:
            break
    t_engine = float(np.mean([r[0] for r in runs]))
    t_model = float(np.mean([r[1] for r in runs]))
    info = runs[-1][3]
    value = mp / t_engine
    line = {
        "impl": "reference", "metric": METRICS[args.config], "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(runs), "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_engine, "wall_ms_per_step": 1e3 * float(np.mean([r[2] for r in runs])),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config] + "; the reference's torch-CPU loop (README.md:196-227), engine-only: "
                               "split + augment + de-augment + integrate_batch + merge, stand-in model timed separately",
                   "l2": "n/a (CPU)", "steps_note": "steps actually run (capped so that the arm ends within ~150 s of CPU work)"},
        "with_model": mp / t_model,
        "cpu_baseline": {"value": value, "unit": UNIT, **{k: info[k] for k in ("cores", "kind", "sample", "host_cores", "cpu_model",
                                                                              "stage_seconds_sample", "extrapolated")}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# this engine, one GPU, tiled configs
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    if args.config == "cfg3":
        if world > 1:
            raise SystemExit("cfg3 (losses) is a single-GPU bench line")
        return run_ours_losses(args)
    cfg = CONFIGS[args.config]
    if world > 1:
        from pytorch_toolbelt_b200.distributed.bench_multi import run_multi_gpu

        return run_multi_gpu(args, cfg, args.config, METRICS[args.config], UNIT, WORKLOADS[args.config])

    from pytorch_toolbelt_b200 import _native as N
    from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    H, W, Cin, T, S, K = (cfg[k] for k in ("H", "W", "Cin", "T", "S", "K"))
    image_host = torch.from_numpy(synthetic_image(cfg)).pin_memory()
    eng = TiledInferenceEngine((H, W, Cin), T, S, channels=K, weight=cfg["weight"], tta=cfg["tta"], analytic_weight=cfg["analytic"])
    A, n = eng.views, eng.num_tiles
    Wm = torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))
    wm = (C.c_float * (K * Cin))(*[float(v) for v in Wm.flatten().tolist()])
    stream = N.stream_ptr(dev)

    def model(x, out=None):
        y = out if out is not None else torch.empty((x.shape[0], K, T, T), dtype=torch.float32, device=dev)
        N.check(N.lib().tb_standin_model(x.data_ptr(), x.shape[0], Cin, K, T, T, wm, y.data_ptr(), N.stream_ptr(dev)), "standin")
        return y

    image = image_host.to(dev)
    aug = torch.empty((A * n, Cin, T, T), dtype=torch.float32, device=dev)
    pred = torch.empty((A * n, K, T, T), dtype=torch.float32, device=dev)
    out = torch.empty((K, H, W), dtype=torch.float32, device=dev)
    slice_plan = eng.slice_plan(image, out=aug)
    merge_plan = eng.merge_plan(pred, out=out)
    slice_plan.run()
    model(aug, out=pred)
    torch.cuda.synchronize()

    def step(with_model=False):
        slice_plan.run()
        if with_model:
            model(aug, out=pred)
        merge_plan.run()

    def timed(fn, steps):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / steps  # ms per step

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(True)
    torch.cuda.synchronize()
    # small configs: a CUDA graph of the launch pair removes the per-launch host cost from a 0.1 ms step
    graph = None
    if args.config == "cfg1":
        try:
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                gs, gm = eng.slice_plan(image, out=aug), eng.merge_plan(pred, out=out)
                gs.run()
                gm.run()
                with torch.cuda.graph(graph, stream=side):
                    gs.run()
                    gm.run()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
        except Exception as e:  # capture is an optimisation of the submission path only
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e!r}); timing the two launches\n")
            graph = None

    with ClockSampler(local, uuid=str(getattr(torch.cuda.get_device_properties(dev), "uuid", "")) or None) as clocks:
        ms_step = timed(step, args.steps)
        ms_graph = timed(graph.replay, args.steps) if graph is not None else None
        # per-kernel times on the same stream, same warm state (for the roofline of the dominant kernel)
        ms_slice = timed(slice_plan.run, args.steps)
        ms_merge = timed(merge_plan.run, args.steps)
        ms_with_model = timed(lambda: step(True), args.steps)
        deadline = time.time() + 3.0  # keep the same load on (untimed) until the sampler has 3 readings
        while clocks.samples < 3 and clocks.source != "unavailable" and time.time() < deadline:
            step()
            torch.cuda.synchronize()
    clk = clocks.summary()
    ms_launches = None
    if ms_graph is not None and ms_graph < ms_step:
        ms_launches, ms_step = ms_step, ms_graph

    # ---- e2e through the public API with host buffers (pinned): H2D image, kernels + model, D2H result ----
    out_host = torch.empty((K, H, W), dtype=torch.float32).pin_memory()
    rows_per_chunk = 2 if args.config != "cfg1" else 15
    e2e_steps = max(2, args.steps // 2)

    def e2e_step():
        eng.predict_to_host(image_host, model, out_host=out_host, rows_per_chunk=rows_per_chunk)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_labels = None
    if K > 1:
        lab_host = torch.empty((H, W), dtype=torch.uint8).pin_memory()

        def e2e_lab():
            eng.predict_to_host(image_host, model, out_host=lab_host, rows_per_chunk=rows_per_chunk, output="labels")

        for _ in range(2):
            e2e_lab()
        torch.cuda.synchronize()
        ms_lab = timed(e2e_lab, e2e_steps)
        e2e_labels = {"value": H * W / 1e6 / (ms_lab * 1e-3), "unit": UNIT, "ms_per_step": ms_lab, "h2d_bytes_per_step": H * W * Cin,
                      "d2h_bytes_per_step": H * W,
                      "api": "predict_to_host(..., output='labels'): argmax fused into the merge epilogue, uint8 [H,W] downloaded"}

    mp = H * W / 1e6
    peak, peak_src = measured_peaks()
    bytes_slice = H * W * Cin * 1 + n * A * Cin * T * T * 4
    bytes_merge = n * A * K * T * T * 4 + K * H * W * 4
    merge_gbs = bytes_merge / (ms_merge * 1e-3) / 1e9
    slice_gbs = bytes_slice / (ms_slice * 1e-3) / 1e9
    merge_kernel = "merge_single_view_kernel" if A == 1 else "merge_all_tma_kernel"
    slice_kernel = "slice_single_view_kernel" if A == 1 else "slice_augment_fast_kernel"
    traffic, traffic_src = measured_traffic(args.config, merge_kernel)

    t_cpu_engine, t_cpu_model, _, cpu_info = cpu_reference_run(cfg, max_tiles=cfg["cpu_tiles"])
    line = {
        "metric": METRICS[args.config], "value": mp / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOADS[args.config] + "; engine-only (stand-in model excluded, predictions resident)",
            "l2": (f"inputs ({bytes_slice / 1e9:.2f} GB tiles out + {(bytes_merge - K * H * W * 4) / 1e9:.2f} GB predictions in per step) "
                   ">> 126 MB L2; no flush needed") if bytes_merge > 1e9 else
                  "working set 0.42 GB per step > 126 MB L2 (3.3x); not flushed between steps -- say so when quoting",
            "kernels_per_step": [slice_kernel, merge_kernel],
            "weights": "in-register separable gaussian" if (cfg["analytic"] and cfg["weight"] == "gaussian") else "L2-resident table",
            "submission": "CUDA graph replay of the 2-launch step" if ms_launches is not None else "2 prepared launches per step",
        },
        "with_model": {"value": mp / (ms_with_model * 1e-3), "unit": UNIT, "ms_per_step": ms_with_model},
        "e2e": {"value": mp / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": H * W * Cin, "d2h_bytes_per_step": K * H * W * 4,
                "api": f"TiledInferenceEngine.predict_to_host(pinned host image, model) -> pinned host [{K},{H},{W}] f32, pipelined "
                       f"over chunks of {rows_per_chunk} tile rows (H2D / kernels / D2H overlapped on 3 streams)"},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "kernel": merge_kernel, "achieved": merge_gbs, "peak": peak, "unit": "GB/s",
                     "frac": merge_gbs / peak, "traffic": traffic, "peak_source": peak_src, "traffic_source": traffic_src,
                     "algorithmic_bytes": bytes_merge, "ms": ms_merge},
        "roofline_slice": {"bound": "hbm", "kernel": slice_kernel, "achieved": slice_gbs, "peak": peak, "unit": "GB/s",
                           "frac": slice_gbs / peak, "algorithmic_bytes": bytes_slice, "ms": ms_slice},
        "roofline_step": {"algorithmic_bytes": bytes_slice + bytes_merge,
                          "achieved": (bytes_slice + bytes_merge) / (ms_step * 1e-3) / 1e9,
                          "frac": (bytes_slice + bytes_merge) / (ms_step * 1e-3) / 1e9 / peak},
        "cpu_baseline": {"value": mp / t_cpu_engine, "unit": UNIT, "with_model": mp / t_cpu_model, **cpu_info},
        "clocks": clk,
    }
    if e2e_labels:
        line["e2e_labels"] = e2e_labels
    if ms_launches is not None:
        line["ms_per_step_two_launches"] = ms_launches
    if args.config in ("cfg1", "cfg2"):
        # second baseline: the reference's own torch path as CUDA eager on this same GPU (all tiles)
        loop, kind = reference_loop()
        eager = {}
        for _ in range(2):
            eager = {}
            loop(image_host.numpy(), T, S, K, weight=cfg["weight"], group=cfg["tta"], batch_size=8, timings=eager, device="cuda")
        eager_engine_s = sum(eager[k] for k in ("split", "augment", "deaugment", "integrate", "merge"))
        line["torch_cuda_eager_baseline"] = {
            "value": mp / eager_engine_s, "unit": UNIT, "kind": kind,
            "what": "the reference's own loop with device='cuda' (its torch ops as CUDA eager on this GPU), all tiles, batch 8, host "
                    "split + per-batch H2D as in README.md:219, stand-in model excluded",
            "stage_seconds": {k: round(v, 4) for k, v in eager.items()}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# cfg 3: BinaryFocalLoss + DiceLoss forward+backward, 32x4x1024x1024
# --------------------------------------------------------------------------------------------------------------
LOSS_SHAPE = (32, 4, 1024, 1024)


def _loss_inputs(B, device):
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, *LOSS_SHAPE[1:], generator=g) * 3)
    t = (torch.rand(B, *LOSS_SHAPE[1:], generator=g) < 0.3).float()
    return x.to(device), t.to(device)


def cpu_reference_losses(batch=4):
    """The reference's BinaryFocalLoss + DiceLoss (baseline/_ref when present, else the oracle port) forward+backward on a
    `batch`-image slice of cfg 3 on the host cores -> (seconds per FULL 32-image step, extrapolated linearly, info)."""
    kind = "port"
    fn = None
    try:
        from baseline import reference_arm as RA

        if RA.available():
            RA.load()
            from pytorch_toolbelt.losses import BinaryFocalLoss, DiceLoss

            focal, dice = BinaryFocalLoss(), DiceLoss("multilabel")
            fn = lambda z, tt: focal(z, tt) + dice(z, tt)  # noqa: E731
            kind = "reference"
    except Exception as e:
        sys.stderr.write(f"[bench] reference install unusable ({e!r}); using the oracle port\n")
    if fn is None:
        from oracle import losses_oracle as LO

        fn = lambda z, tt: LO.focal_sigmoid(z, tt, alpha=None) + LO.dice_loss(z, tt, "multilabel")  # noqa: E731
    cores = host_cores()
    torch.set_num_threads(cores)
    x, t = _loss_inputs(batch, "cpu")
    best = float("inf")
    for _ in range(2):
        xr = x.clone().requires_grad_(True)
        t0 = time.perf_counter()
        fn(xr, t).backward()
        best = min(best, time.perf_counter() - t0)
    scale = LOSS_SHAPE[0] / batch
    info = {"cores": cores, "host_cores": cores, "kind": kind, "cpu_model": _cpu_model(), "extrapolated": True,
            "sample": f"{batch} of 32 images ({batch}x4x1024x1024) forward+backward on torch-CPU, best of 2, scaled x{scale:.0f} "
                      f"(elementwise + per-class sums: linear in the batch); {best:.2f}s measured"}
    return best * scale, info


def run_reference_losses(args):
    secs, info = cpu_reference_losses(batch=4)
    numel = int(np.prod(LOSS_SHAPE))
    value = numel / secs / 1e9
    line = {"impl": "reference", "metric": METRICS["cfg3"], "value": value, "unit": "Gelem/s", "n_gpus": args.gpus, "steps": 1,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3 (BASELINE configs[2]): BinaryFocalLoss + DiceLoss(multilabel) forward+backward on 32x4x1024x1024 "
                                   "f32 logits, Bernoulli(0.3) float targets; torch-CPU reference", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "Gelem/s", **info},
            "e2e": {"value": value, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def run_ours_losses(args):
    import pytorch_toolbelt_b200.losses as L

    _, _, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    x, t = _loss_inputs(LOSS_SHAPE[0], dev)
    numel = x.numel()
    joint = L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss("multilabel"))
    assert joint.fusable(x, t)

    def step():
        xr = x.detach().requires_grad_(True)
        joint(xr, t).backward()
        return xr.grad

    def timed(fn, steps):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / steps

    warm = max(args.warmup, 3)
    steps = max(args.steps, 10)
    for _ in range(warm):
        step()
    with ClockSampler(local, uuid=str(getattr(torch.cuda.get_device_properties(dev), "uuid", "")) or None) as clocks:
        ms = timed(step, steps)
        deadline = time.time() + 3.0
        while clocks.samples < 3 and clocks.source != "unavailable" and time.time() < deadline:
            step()
            torch.cuda.synchronize()
    clk = clocks.summary()
    # e2e: pinned host logits + targets -> device, loss value back to the host (the gradient stays on the device, as in training)
    xh, th = x.cpu().pin_memory(), t.cpu().pin_memory()
    xd, td = torch.empty_like(x), torch.empty_like(t)
    res = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        xd.copy_(xh, non_blocking=True)
        td.copy_(th, non_blocking=True)
        xr = xd.detach().requires_grad_(True)
        loss = joint(xr, td)
        loss.backward()
        res.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, 3)
    peak, peak_src = measured_peaks()
    alg = 5 * numel * 4  # two passes over (x, t) + one gradient write: SURVEY 8(d), 2.684 GB
    gbs = alg / (ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic("cfg3", "joint_focal_dice")
    cpu_s, cpu_info = cpu_reference_losses(batch=4)
    line = {
        "metric": METRICS["cfg3"], "value": numel / (ms * 1e-3) / 1e9, "unit": "Gelem/s", "n_gpus": 1, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg3 (BASELINE configs[2]): BinaryFocalLoss + DiceLoss(multilabel) forward+backward on 32x4x1024x1024 f32 "
                               "logits, Bernoulli(0.3) float targets, through ONE fused two-pass kernel (BinaryFocalDiceLoss -> "
                               "tb_focal_dice_fwd_bwd); step = loss value + d loss / d logits",
                   "l2": "2.7 GB of traffic per step >> 126 MB L2; no flush needed"},
        "e2e": {"value": numel / (ms_e2e * 1e-3) / 1e9, "unit": "Gelem/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 2 * numel * 4,
                "d2h_bytes_per_step": 4, "api": "pinned host logits+targets -> BinaryFocalDiceLoss(...).backward() -> loss value on the host"},
        "gpu_launches": 4 * steps,
        "roofline": {"bound": "hbm", "kernel": "joint_stats_kernel + joint_grad_kernel", "achieved": gbs, "peak": peak, "unit": "GB/s",
                     "frac": gbs / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes": alg, "ms": ms},
        "cpu_baseline": {"value": numel / cpu_s / 1e9, "unit": "Gelem/s", **cpu_info},
        "clocks": clk,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--ref-tiles", type=int, default=0, help="reference arm: bound the sample to the first N tiles (0 = config default)")
    ap.add_argument("--cpu-sample-tiles", type=int, default=0, help="(compat) same as --ref-tiles")
    args = ap.parse_args()
    if args.cpu_sample_tiles and not args.ref_tiles:
        args.ref_tiles = args.cpu_sample_tiles
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
