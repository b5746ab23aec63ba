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


def losses_check(rank, world, dev):
    """Sync{Dice,Jaccard}Loss on a batch shard per rank == the single-GPU loss of the whole batch (value and the
    gradient rows of the shard); the only cross-rank traffic is the float64 [C,3] all-reduce."""
    import pytorch_toolbelt_b200.losses as L
    from pytorch_toolbelt_b200.distributed.losses import SyncDiceLoss, SyncJaccardLoss

    per = 2
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(per * world, 4, 96, 80, generator=g) * 3).to(dev)
    t_ml = (torch.rand(per * world, 4, 96, 80, generator=g) < 0.3).float().to(dev)
    t_mc = torch.randint(0, 4, (per * world, 96, 80), generator=g).to(dev)
    mine = slice(per * rank, per * rank + per)
    ok = True
    cases = [
        ("dice multilabel", L.DiceLoss("multilabel"), SyncDiceLoss("multilabel"), t_ml),
        ("dice multiclass log smooth", L.DiceLoss("multiclass", log_loss=True, smooth=1.0),
         SyncDiceLoss("multiclass", log_loss=True, smooth=1.0), t_mc),
        ("jaccard multilabel", L.JaccardLoss("multilabel", smooth=0.5), SyncJaccardLoss("multilabel", smooth=0.5), t_ml),
    ]
    for name, whole_loss, sync_loss, t in cases:
        xw = x.clone().requires_grad_(True)
        whole = whole_loss(xw, t)
        whole.backward()
        xl = x[mine].clone().requires_grad_(True)
        synced = sync_loss(xl, t[mine])
        synced.backward()
        torch.cuda.synchronize()
        v_ok = abs(float(synced) - float(whole)) <= 1e-6 * max(1.0, abs(float(whole)))
        scale = float(xw.grad.abs().max())
        g_ok = float((xl.grad - xw.grad[mine]).abs().max()) <= 1e-5 * scale + 1e-12
        if not (v_ok and g_ok):
            print(f"rank {rank}/{world} {name}: value {float(synced)} vs {float(whole)}, "
                  f"grad max err {float((xl.grad - xw.grad[mine]).abs().max())} (scale {scale}) MISMATCH", flush=True)
        ok = ok and v_ok and g_ok
    if rank == 0:
        print(f"sync losses over {world} ranks: {'ok' if ok else 'MISMATCH'}", flush=True)
    return ok


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
    ok = losses_check(rank, world, dev) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag) == 1 else "FAIL")
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
