"""Worker of tests/test_ragged_gpu.py and of the N-GPU parity check: one process per rank.

    python tests/ragged_worker.py <rank> <world> <port> <out_dir> [same_gpu]

Every rank builds the same seeded predictions, keeps the slice for ITS row-major tile range, runs the ragged-strip merge
(peer stores + flags, distributed/ragged.py) for a few epochs and saves the pixels it wrote (NaN elsewhere). With `same_gpu`
all ranks share cuda:0 (CUDA IPC works between processes on one device), so the multi-rank path is exercised on a 1-GPU box."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_toolbelt_b200.distributed.ragged import RaggedMerger, RaggedPlan  # noqa: E402
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine  # noqa: E402

CASES = [
    # shape, T, S, K, tta, weight
    ((1500, 1100, 3), 256, 128, 4, "d4", "pyramid"),
    ((1000, 1300, 1), 128, 64, 2, "flips", "pyramid"),
    ((1600, 1400, 1), 256, 128, 1, None, "gaussian"),    # single-view kernel
    ((700, 520, 3), 128, 32, 3, None, "pyramid"),         # 4x4 covering tiles per pixel
    ((384, 520, 3), 128, 32, 2, None, "pyramid"),         # 126 tiles, 4x4 covering: with 3 ranks ranges are shorter than a
    ((384, 520, 1), 128, 32, 2, "fliplr", "mean"),        # covering set -> chained (seeded AND spilling) blocks, both kernels
]


def main():
    rank, world, port, out_dir = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    same_gpu = len(sys.argv) > 5 and sys.argv[5] == "same_gpu"
    dev = torch.device("cuda", 0 if same_gpu else rank)
    torch.cuda.set_device(dev)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for ci, (shape, T, S, K, tta, weight) in enumerate(CASES):
        eng = TiledInferenceEngine(shape, T, S, channels=K, weight=weight, tta=tta, device=dev)
        ny, nx = eng.slicer.grid_shape
        hp, wp = eng.slicer.target_shape
        if nx * ny // world < nx:
            continue
        plan = RaggedPlan.build(nx, ny, T, T, S, S, hp, rank, world)
        full = torch.randn(eng.views, eng.num_tiles, K, T, T, generator=torch.Generator().manual_seed(3 + ci)).to(dev)
        own = full[:, plan.t0 : plan.t1].reshape(-1, K, T, T).contiguous()
        # list-of-batches form for one case (per-batch tensor maps / pointers), single tensor otherwise
        preds = own
        if ci == 0:
            n_own = plan.t1 - plan.t0
            view = own.view(eng.views, n_own, K, T, T)
            preds = [view[:, s : s + 7].reshape(-1, K, T, T).contiguous() for s in range(0, n_own, 7)]
        out = torch.full((K, shape[0], shape[1]), float("nan"), dtype=torch.float32, device=dev)
        # multi-view cases run twice: fused single launch (in-kernel flags, the default) and two launches + flag kernels
        results = []
        for fused in ((None, False) if eng.views > 1 else (None,)):
            merger = RaggedMerger(eng, plan, preds, out, timeout_ms=20000, fused=fused)
            for epoch in range(4):  # double-buffered seeds + acks: several epochs back to back without host synchronisation
                out.fill_(float("nan"))
                merger.run()
            merger.check()
            results.append(out.clone())
            dist.barrier()
            merger.close()
        if len(results) == 2:
            same = torch.equal(torch.nan_to_num(results[0], nan=-1e30), torch.nan_to_num(results[1], nan=-1e30))
            ok = ok and same
            if not same:
                print(f"rank {rank}: case {ci}: fused and two-launch results differ", flush=True)
        torch.save(results[0].cpu(), os.path.join(out_dir, f"case{ci}_rank{rank}.pt"))
        if rank == 0:
            torch.save(eng.merge(full.reshape(-1, K, T, T)).cpu(), os.path.join(out_dir, f"case{ci}_ref.pt"))
        dist.barrier()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
