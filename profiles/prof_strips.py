"""Where does the multi-GPU step go? Per-phase device times of one strip step on every rank (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        profiles/prof_strips.py [steps]

Phases (CUDA events on the compute stream, averaged over the steps, reported per rank):
  slice      tb_slice_augment of the rank's tiles
  spill      raw band merge of the rows that spill into the next strip
  post       batch_isend_irecv issue (host) -> the compute stream is released again
  interior   band merge of the rows that need no halo (overlaps the transfer)
  wait       time the compute stream sits in the NCCL wait after the interior merge
  top        seeded band merge of the first (T-S) rows
and the same step with the collective removed (every rank merges its strip without the halo: WRONG rows at the strip
tops, timing only) to expose what the exchange costs end to end. Diagnostics for DESIGN.md section 8 item 2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pytorch_toolbelt_b200.distributed.strips import StripMerger, StripPlan, wait_all  # noqa: E402
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H = W = 5000
    Cin, T, S, K = 3, 512, 256, 4
    eng = TiledInferenceEngine((H, W, Cin), T, S, channels=K, weight="pyramid", tta="d4", device=dev)
    ny, nx = eng.slicer.grid_shape
    hp, wp = eng.slicer.target_shape
    plan = StripPlan.build(ny, T, S, hp, rank, world)
    A = eng.views
    n_own = (plan.r1 - plan.r0) * nx
    image = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, Cin)).astype(np.uint8)).to(dev)
    aug = torch.empty((A * n_own, Cin, T, T), dtype=torch.float32, device=dev)
    pred = torch.randn((A * n_own, K, T, T), dtype=torch.float32, device=dev)
    out = torch.empty((K, H, W), dtype=torch.float32, device=dev)

    def band(rows, raw=False, init=None):
        return eng.merge_plan(pred, out=None if raw else out, rows=rows, tile_rows=(plan.r0, plan.r1), raw=raw, init=init)

    merger = StripMerger(band, plan, K, wp, dev)
    slicer = eng.slice_plan(image, plan.r0 * nx, plan.r1 * nx, out=aug)
    no_halo = band((plan.y0, plan.y1))  # timing-only: the whole strip in one launch, no seed

    names = ["slice", "spill", "post", "interior", "wait", "top"]
    acc = {n: 0.0 for n in names}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def instrumented_step():
        marks = [ev()]
        slicer.run(); marks.append(ev())
        ops = []
        if merger.top is not None:
            ops += [dist.P2POp(dist.irecv, merger.recv[0], plan.prev), dist.P2POp(dist.irecv, merger.recv[1], plan.prev)]
        if merger.spill is not None:
            num, den = merger.spill.run()
            ops += [dist.P2POp(dist.isend, num, plan.next), dist.P2POp(dist.isend, den, plan.next)]
        marks.append(ev())
        reqs = dist.batch_isend_irecv(ops) if ops else []
        marks.append(ev())
        if merger.interior is not None:
            merger.interior.run()
        marks.append(ev())
        wait_all(reqs); marks.append(ev())
        if merger.top is not None:
            merger.top.run()
        marks.append(ev())
        return marks

    def timed(fn, n):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for _ in range(5):
        slicer.run(); merger.run()
    all_marks = []
    dist.barrier(); torch.cuda.synchronize()
    for _ in range(steps):
        all_marks.append(instrumented_step())
    torch.cuda.synchronize()
    for marks in all_marks:
        for n, (a, b) in zip(names, zip(marks, marks[1:])):
            acc[n] += a.elapsed_time(b) / steps
    ms_step = timed(lambda: (slicer.run(), merger.run()), steps)
    ms_no_exchange = timed(lambda: (slicer.run(), no_halo.run()), steps)
    ms_slice = timed(slicer.run, steps)
    ms_merge_one_launch = timed(no_halo.run, steps)
    row = {"rank": rank, "tile_rows": plan.r1 - plan.r0, "ms_step": round(ms_step, 4),
           "ms_step_without_exchange_one_launch": round(ms_no_exchange, 4), "ms_slice_alone": round(ms_slice, 4),
           "ms_strip_merge_one_launch": round(ms_merge_one_launch, 4), "phases_ms": {k: round(v, 4) for k, v in acc.items()}}
    rows = [None] * world
    dist.all_gather_object(rows, row)
    if rank == 0:
        print(json.dumps({"world": world, "steps": steps, "ranks": rows}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
