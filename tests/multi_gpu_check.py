"""Multi-GPU parity check (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank builds the same seeded predictions, runs the strip-sharded merge (NCCL halo exchange) on ITS tile rows
and compares the rows it owns, bit for bit, with a full single-GPU merge of all tiles computed locally.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_toolbelt_b200.distributed.strips import StripMerger, StripPlan, sharded_merge  # noqa: E402
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for shape, T, S, K, tta in [((1500, 1100, 3), 256, 128, 4, "d4"), ((900, 700, 3), 128, 64, 3, None),
                                ((1000, 1000, 1), 128, 128, 2, "flips")]:
        eng = TiledInferenceEngine(shape, T, S, channels=K, weight="pyramid", tta=tta, device=dev)
        ny, nx = eng.slicer.grid_shape
        hp, wp = eng.slicer.target_shape
        full_pred = torch.randn(eng.views, eng.num_tiles, K, T, T, generator=torch.Generator().manual_seed(3)).to(dev)
        reference = eng.merge(full_pred.reshape(-1, K, T, T))
        plan = StripPlan.build(ny, T, S, hp, rank, world)
        own = full_pred[:, plan.r0 * nx : plan.r1 * nx].reshape(-1, K, T, T).contiguous()
        out = torch.full_like(reference, float("nan"))

        def band_merge(rows, raw=False, init=None):
            return eng.merge(own, out=None if raw else out, rows=rows, tile_rows=(plan.r0, plan.r1), raw=raw, init=init)

        sharded_merge(band_merge, plan, K, wp, dev)
        torch.cuda.synchronize()
        top = eng.slicer.margin_top
        a, b = max(plan.y0 - top, 0), min(plan.y1 - top, shape[0])
        if not plan.empty and b > a:
            same = torch.equal(out[:, a:b], reference[:, a:b])
            ok = ok and same
            print(f"rank {rank}/{world} {shape} T{T}/S{S} {tta}: rows [{a},{b}) bit-exact={same}", flush=True)

        # steady-state form (prepared launches, persistent halo buffers): same rows, step after step
        def band_plan(rows, raw=False, init=None):
            return eng.merge_plan(own, out=None if raw else out, rows=rows, tile_rows=(plan.r0, plan.r1), raw=raw, init=init)

        merger = StripMerger(band_plan, plan, K, wp, dev)
        for it in range(3):
            out.fill_(float("nan"))
            merger.run()
            torch.cuda.synchronize()
            if not plan.empty and b > a:
                same = torch.equal(out[:, a:b], reference[:, a:b])
                ok = ok and same
                if not same:
                    print(f"rank {rank}/{world} {shape}: StripMerger step {it} MISMATCH", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag) == 1 else "FAIL")
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
