"""CPU, gloo, world_size 2 and 3: the strip partition + halo exchange choreography of the multi-GPU path, with the
oracle's torch-CPU merger standing in for the CUDA band-merge kernel. The sharded result must equal the single-process
merge bit for bit (the receiver seeds its accumulators with the halo, so the tile order is preserved)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tiles_oracle as TO
from pytorch_toolbelt_b200.distributed.strips import StripMerger, StripPlan, partition_tile_rows, sharded_merge


def test_partition_tile_rows():
    assert partition_tile_rows(19, 8) == [(0, 3), (3, 6), (6, 9), (9, 11), (11, 13), (13, 15), (15, 17), (17, 19)]
    assert partition_tile_rows(39, 8)[-1] == (35, 39) and partition_tile_rows(39, 8)[0] == (0, 5)
    assert partition_tile_rows(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for ny in (1, 5, 19, 39):
        for world in (1, 2, 3, 4, 8):
            parts = partition_tile_rows(ny, world)
            assert parts[0][0] == 0 and parts[-1][1] == ny
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    p = StripPlan.build(19, 512, 256, 5120, 3, 8)
    assert (p.r0, p.r1, p.y0, p.y1, p.halo, p.halo_in, p.prev, p.next) == (9, 11, 2304, 2816, 256, 256, 2, 4)
    last = StripPlan.build(19, 512, 256, 5120, 7, 8)
    assert last.y1 == 5120 and last.halo == 0 and last.next is None
    idle = StripPlan.build(2, 512, 256, 768, 3, 4)
    assert idle.empty and idle.prev is None and idle.next is None


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _band_merger(grid, weight, preds, K, tile_rows, nx):
    """Oracle band merge: THIS rank's tiles (rows tile_rows) accumulated in tile order into canvas rows [y0,y1)."""
    w = torch.from_numpy(np.asarray(weight, dtype=np.float32))
    hp, wp = grid["target_shape"]

    def band_merge(rows, raw=False, init=None):
        y0, y1 = rows
        num = torch.zeros((K, y1 - y0, wp)) if init is None else init[0].clone()
        den = torch.zeros((y1 - y0, wp)) if init is None else init[1].clone()
        for t in range(tile_rows[0] * nx, tile_rows[1] * nx):
            x, y, tw, th = (int(v) for v in grid["crops"][t])
            a, b = max(y, y0), min(y + th, y1)
            if a >= b:
                continue
            tile = preds[t]
            num[:, a - y0 : b - y0, x : x + tw] += tile[:, a - y : b - y] * w[a - y : b - y]
            den[a - y0 : b - y0, x : x + tw] += w[a - y : b - y]
        return (num, den) if raw else num / den

    return band_merge


def _worker(rank, world, port, shape, tile, step, K, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = TO.tile_grid(shape[0], shape[1], tile, step)
        n = len(grid["crops"])
        nx = len({int(c[0]) for c in grid["crops"]})
        ny = n // nx
        weight = TO.pyramid_weight(*tile)
        preds = torch.randn(n, K, tile[0], tile[1], generator=torch.Generator().manual_seed(7))
        hp, wp = grid["target_shape"]
        plan = StripPlan.build(ny, tile[0], step[0], hp, rank, world)
        band = _band_merger(grid, weight, preds, K, (plan.r0, plan.r1), nx)
        pieces = sharded_merge(band, plan, K, wp, torch.device("cpu"))
        torch.save([(a, b, t) for a, b, t in pieces], os.path.join(result_dir, f"rank{rank}.pt"))
        dist.barrier()

        # the prepared (steady-state) form must give the same pieces, step after step, from the same buffers
        class _Band:
            def __init__(self, rows, raw=False, init=None):
                self.rows, self.raw, self.init = rows, raw, init

            def run(self):
                return band(self.rows, raw=self.raw, init=self.init)

        merger = StripMerger(_Band, plan, K, wp, torch.device("cpu"))
        for _ in range(2):
            again = merger.run()
            assert len(again) == len(pieces)
            for (a0, b0, t0), (a1, b1, t1) in zip(pieces, again):
                assert (a0, b0) == (a1, b1) and torch.equal(t0, t1)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_matches_single_process(world, tmp_path):
    shape, tile, step, K = (300, 200), (64, 32), (32, 16), 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, tile, step, K, str(tmp_path)), nprocs=world, join=True)
    grid = TO.tile_grid(shape[0], shape[1], tile, step)
    preds = torch.randn(len(grid["crops"]), K, tile[0], tile[1], generator=torch.Generator().manual_seed(7))
    ref = TO.TorchMerger(grid["target_shape"], K, TO.pyramid_weight(*tile))
    ref.integrate_batch(preds, grid["crops"])
    full = ref.merge()
    covered = torch.zeros(grid["target_shape"][0], dtype=torch.bool)
    for r in range(world):
        for y0, y1, t in torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")):
            assert torch.equal(t, full[:, y0:y1]), (r, y0, y1)
            assert not covered[y0:y1].any()
            covered[y0:y1] = True
    assert covered.all()


def _loss_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import losses_oracle as LO
        from pytorch_toolbelt_b200.distributed.losses import global_mean

        g = torch.Generator().manual_seed(11)
        x = torch.randn(4 * world, 3, 16, 16, generator=g) * 3
        t = (torch.rand(4 * world, 3, 16, 16, generator=g) < 0.3).float()
        mine = slice(4 * rank, 4 * rank + 4)
        # (1) what SyncDiceLoss exchanges: the per-class sums over (batch, space) add up over batch shards, and the
        #     loss of the summed statistics is the loss of the whole batch (multilabel, smooth 0, eps 1e-7)
        p = torch.sigmoid(x[mine]).double()
        tt = t[mine].double()
        stats = torch.stack([(p * tt).sum(dim=(0, 2, 3)), (p + tt).sum(dim=(0, 2, 3)), tt.sum(dim=(0, 2, 3))], dim=1)
        dist.all_reduce(stats)
        inter, card, tsum = stats[:, 0], stats[:, 1], stats[:, 2]
        score = (2.0 * inter) / card.clamp_min(1e-7)
        synced = ((1.0 - score) * (tsum > 0)).mean()
        whole = LO.dice_loss(x, t, "multilabel")
        assert abs(float(synced) - float(whole)) <= 1e-6 * max(1.0, abs(float(whole))), (float(synced), float(whole))
        # (2) global_mean: value = mean over ranks, gradient = 1/world to the local term
        xl = x[mine].clone().requires_grad_(True)
        local = LO.focal_sigmoid(xl, t[mine], alpha=None)
        glob = global_mean(local)
        glob.backward()
        xw = x.clone().requires_grad_(True)
        whole_focal = LO.focal_sigmoid(xw, t, alpha=None)
        whole_focal.backward()
        torch.testing.assert_close(glob.detach(), whole_focal.detach(), rtol=1e-6, atol=1e-9)
        torch.testing.assert_close(xl.grad, xw.grad[mine], rtol=1e-5, atol=1e-10)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_data_parallel_loss_reductions(tmp_path):
    world = 2
    mp.spawn(_loss_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def _grad_convention_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import losses_oracle as LO
        from pytorch_toolbelt_b200.distributed.losses import global_mean

        g = torch.Generator().manual_seed(21)
        feats = torch.randn(2 * world, 3, 8, 8, generator=g)
        t = (torch.rand(2 * world, 3, 8, 8, generator=g) < 0.3).float()
        w0 = torch.tensor([0.7, -1.3, 2.1]).view(1, 3, 1, 1)
        mine = slice(2 * rank, 2 * rank + 2)
        # single-process whole-batch parameter gradient
        w = w0.clone().requires_grad_(True)
        LO.focal_sigmoid(feats * w, t, alpha=None).backward()
        whole = w.grad.clone()
        for convention in ("sum", "ddp_mean"):
            wl = w0.clone().requires_grad_(True)
            loss = global_mean(LO.focal_sigmoid(feats[mine] * wl, t[mine], alpha=None), grad_reduction=convention)
            loss.backward()
            gsum = wl.grad.clone()
            dist.all_reduce(gsum)  # SUM over ranks
            reduced = gsum if convention == "sum" else gsum / world  # "ddp_mean": what DDP's averaging produces
            torch.testing.assert_close(reduced, whole, rtol=1e-5, atol=1e-9)
        with pytest.raises(ValueError):
            global_mean(torch.zeros(()), grad_reduction="mean")
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sync_loss_gradient_conventions(tmp_path):
    """d(global loss)/d(local logits) must be SUMMED over ranks; under DDP's averaging use
    grad_reduction="ddp_mean" (backward x world). Checked on parameter gradients against the whole-batch gradient."""
    world = 2
    mp.spawn(_grad_convention_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))
