"""bench.py body for N > 1 GPUs (one process per GPU): strong scaling of ONE image over ragged strips.

Every rank slices / predicts / merges only its own row-major tile range (distributed/ragged.py); the only cross-rank traffic
is the un-normalised partial sums of the canvas blocks that straddle a range boundary, stored by the merge kernel straight
into the next rank's seed buffer over NVLink (CUDA-IPC mapped peer memory) and ordered by two flags per boundary. NCCL is
used for the benchmark's own barriers / max-over-ranks reduction only, never on the data path. The result stays sharded:
rank g holds the pixels whose last covering tile it owns. value = image megapixels / max-over-ranks device time per step.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from .. import _native as N
from ..inference.engine import TiledInferenceEngine
from ..utils.clocks import ClockSampler
from .ragged import RaggedMerger, RaggedPlan, partition_tiles


def run_multi_gpu(args, CFG, cfg_name, METRIC, UNIT, WORKLOAD):
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    H, W, Cin, T, S, K = (CFG[k] for k in ("H", "W", "Cin", "T", "S", "K"))
    rng = np.random.default_rng(0)
    image_np = rng.integers(0, 256, (H, W, Cin)).astype(np.uint8)
    eng = TiledInferenceEngine((H, W, Cin), T, S, channels=K, weight=CFG["weight"], tta=CFG["tta"], analytic_weight=CFG["analytic"],
                               device=dev)
    ny, nx = eng.slicer.grid_shape
    hp, wp = eng.slicer.target_shape
    top = eng.slicer.margin_top
    plan = RaggedPlan.build(nx, ny, T, T, S, S, hp, rank, world)
    A = eng.views
    t0, t1 = plan.t0, plan.t1
    n_own = t1 - t0
    Wm = torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))
    wm = (C.c_float * (K * Cin))(*[float(v) for v in Wm.flatten().tolist()])

    def model(x, out=None):
        y = out if out is not None else torch.empty((x.shape[0], K, T, T), dtype=torch.float32, device=dev)
        if x.shape[0]:
            N.check(N.lib().tb_standin_model(x.data_ptr(), x.shape[0], Cin, K, T, T, wm, y.data_ptr(), N.stream_ptr(dev)), "standin")
        return y

    # the image rows this rank's tiles read (its slab): only those are uploaded in the e2e leg
    iy0, iy1 = max(plan.rows[0] - top, 0), min(plan.rows[1] - top, H)
    image_host = torch.from_numpy(image_np).pin_memory()
    image = torch.zeros((H, W, Cin), dtype=torch.uint8, device=dev)
    image[iy0:iy1].copy_(image_host[iy0:iy1])
    del image_np
    aug = torch.empty((A * n_own, Cin, T, T), dtype=torch.float32, device=dev)
    pred = torch.empty((A * n_own, K, T, T), dtype=torch.float32, device=dev)
    out = torch.empty((K, H, W), dtype=torch.float32, device=dev)  # only the pixels this rank owns are ever written
    slicer = eng.slice_plan(image, t0, t1, out=aug)
    slicer.run()
    model(aug, out=pred)
    torch.cuda.synchronize()
    merger = RaggedMerger(eng, plan, pred, out)

    def step(with_model=False):
        slicer.run()
        if with_model:
            model(aug, out=pred)
        merger.run()

    def timed(fn, steps):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        ms = torch.tensor([start.elapsed_time(stop) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # device time, max over ranks
        dist.barrier()
        return float(ms.item())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(True)
    with ClockSampler(local, uuid=str(getattr(torch.cuda.get_device_properties(dev), "uuid", "")) or None) as clocks:
        ms_step = timed(step, args.steps)
        ms_slice = timed(slicer.run, args.steps)
        ms_merge = timed(merger.run, args.steps)
        ms_with_model = timed(lambda: step(True), args.steps)
        extra = 0
        while extra < 200:  # same load, untimed, until the sampler has 3 readings (all ranks take the same decision)
            need = torch.tensor([1 if (clocks.samples < 3 and clocks.source != "unavailable") else 0], device=dev)
            dist.broadcast(need, src=0)
            if int(need.item()) == 0:
                break
            for _ in range(10):
                step()
            extra += 10
    clk = clocks.summary()
    merger.check()

    # where a step goes, per rank (CUDA events between the launches of 5 untimed steps; max and min over ranks)
    def phase_probe():
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        marks[0].record()
        slicer.run()
        marks[1].record()
        merger.run(marks=marks[2:3])
        marks[3].record()
        return marks

    dist.barrier()
    torch.cuda.synchronize()
    probes = [phase_probe() for _ in range(5)]
    torch.cuda.synchronize()
    ph = torch.zeros(3, device=dev)
    for m in probes[1:]:
        ph += torch.tensor([m[0].elapsed_time(m[1]), m[1].elapsed_time(m[2]), m[2].elapsed_time(m[3])], device=dev)
    ph /= len(probes) - 1
    ph_max, ph_min = ph.clone(), ph.clone()
    dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(ph_min, op=dist.ReduceOp.MIN)

    # e2e through the public API with host buffers: every rank uploads ITS image slab, runs its tile range, downloads the
    # rows its pixels lie in (the seam rows are downloaded by both neighbours; each holds its own pixels of them)
    oy0, oy1 = iy0, iy1
    rows_out = max(oy1 - oy0, 0)
    out_host = torch.empty((K, max(rows_out, 1), W), dtype=torch.float32).pin_memory()

    def e2e_step():
        image[iy0:iy1].copy_(image_host[iy0:iy1], non_blocking=True)
        slicer.run()
        model(aug, out=pred)
        merger.run()
        if rows_out:
            out_host[:, :rows_out].copy_(out[:, oy0:oy1], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, max(2, args.steps // 2))
    merger.check()

    slab = torch.tensor([float((iy1 - iy0) * W * Cin), float(K * rows_out * W * 4)], device=dev, dtype=torch.float64)
    dist.all_reduce(slab)
    if rank == 0:
        mp = H * W / 1e6
        peak = 6650.0
        pk = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peak = float(json.load(open(pk))["hbm_gbs"])
        n = eng.num_tiles
        b_alg = H * W * Cin + n * A * Cin * T * T * 4 + n * A * K * T * T * 4 + K * H * W * 4
        halo_bytes = (K + 1) * (T - S) * wp * 4
        parts = partition_tiles(n, world)
        line = {
            "metric": METRIC, "value": mp / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD + f"; {n} tiles in row-major ranges of {sorted({b - a for a, b in parts}, reverse=True)} tiles over "
                            f"{world} ranks; engine-only (stand-in model excluded, predictions resident)",
                "l2": "per-rank inputs >> 126 MB L2; no flush needed",
                "parallelism": f"ragged-strips{world}",
                "halo_bytes_per_boundary": halo_bytes,
                "collective": "none on the data path: the merge kernel stores the un-normalised (K+1) x (T-S) x Wp overlap sums of each range "
                              "boundary into the next rank's seed buffer (NVLink peer stores, CUDA IPC) + 2 flags per boundary; "
                              + ("ONE merge launch per rank and step, flags waited / published inside the kernel" if merger.fused else
                                 "three merge launches per rank and step (spill on a side stream) + one-thread flag kernels"),
            },
            "with_model": {"value": mp / (ms_with_model * 1e-3), "unit": UNIT, "ms_per_step": ms_with_model},
            "phases_ms_max_over_ranks": {"slice": ms_slice, "merge_with_exchange": ms_merge},
            "phases_ms_inside_a_step": {"what": "CUDA events between the launches of one step, mean of 4 steps: [max over ranks, min over ranks]",
                                        "slice": [float(ph_max[0]), float(ph_min[0])],
                                        "merge_first_half": [float(ph_max[1]), float(ph_min[1])],
                                        "merge_second_half": [float(ph_max[2]), float(ph_min[2])],
                                        "halves": "fused: [whole merge launch, nothing]; unfused: [final merge (spill on the side stream), "
                                                  "flag wait + seeded merge + join]"},
            "e2e": {"value": mp / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(slab[0].item()), "d2h_bytes_per_step": int(slab[1].item()),
                    "api": "per rank: pinned host image SLAB -> slice / model / RaggedMerger on its tile range -> pinned host rows"},
            "gpu_launches": (2 if merger.fused else 1 + 3 + 4) * args.steps,
            "roofline_step": {"algorithmic_bytes": b_alg, "achieved": b_alg / (ms_step * 1e-3) / 1e9, "peak": peak * world,
                              "frac": b_alg / (ms_step * 1e-3) / 1e9 / (peak * world), "unit": "GB/s",
                              "note": "whole-job algorithmic bytes / max-over-ranks step time / (N x measured single-GPU peak)"},
            "clocks": clk,
        }
        print(json.dumps(line))
    dist.barrier()
    merger.close()
    dist.destroy_process_group()
