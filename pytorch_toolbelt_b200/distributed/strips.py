"""Multi-GPU tiled inference: partition the tile grid into horizontal strips, one per rank.

New relative to the reference (its TileMerger is single-device, tiles.py:306-308). Rank g owns a
contiguous block of tile ROWS [r0, r1) and the canvas rows [r0*S, r1*S) (the last rank owns through the bottom).
Its tiles spill (T - S) rows into the next rank's territory; that un-normalised partial sum (K numerator planes +
1 weight plane) is the only data that crosses ranks: one point-to-point transfer per strip boundary (NCCL over
NVLink on GPUs, gloo in the CPU tests). The receiver SEEDS its accumulators with the halo before adding its own
tiles, so the sharded result is bit-identical to the single-GPU merge (same tile order, same rounding).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["partition_tile_rows", "StripPlan", "exchange_halo", "sharded_merge", "StripMerger"]


def partition_tile_rows(ny: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, as-even-as-possible split of ny tile rows over `world` ranks (earlier ranks get the extra row).
    Ranks beyond ny get empty ranges."""
    base, extra = divmod(ny, world)
    out, r = [], 0
    for g in range(world):
        n = base + (1 if g < extra else 0)
        out.append((r, r + n))
        r += n
    return out


@dataclass
class StripPlan:
    rank: int
    world: int
    r0: int          # first tile row owned
    r1: int          # one past the last tile row owned
    y0: int          # first canvas row owned
    y1: int          # one past the last canvas row owned
    halo: int        # rows this strip spills into the next one (T - S), 0 for the last strip
    halo_in: int     # rows the previous strip spills into this one, 0 for the first strip
    prev: Optional[int]  # rank that sends us a halo (None for the first non-empty strip)
    next: Optional[int]  # rank we send our spill to

    @property
    def empty(self) -> bool:
        return self.r1 <= self.r0

    @staticmethod
    def build(ny: int, tile_h: int, step_h: int, canvas_h: int, rank: int, world: int) -> "StripPlan":
        parts = partition_tile_rows(ny, world)
        r0, r1 = parts[rank]
        nonempty = [g for g, (a, b) in enumerate(parts) if b > a]
        y0 = r0 * step_h
        y1 = canvas_h if (r1 >= ny) else r1 * step_h
        halo = max(0, tile_h - step_h)
        prev = nxt = None
        if r1 > r0:
            idx = nonempty.index(rank)
            prev = nonempty[idx - 1] if idx > 0 else None
            nxt = nonempty[idx + 1] if idx + 1 < len(nonempty) else None
        if halo > 0 and nxt is not None:
            # the spill must stay inside the next strip: tiles taller than two strips would need a multi-hop halo
            n0, n1 = parts[nxt]
            next_rows = (canvas_h if n1 >= ny else n1 * step_h) - n0 * step_h
            if halo > next_rows:
                raise ValueError("tile overlap exceeds the neighbouring strip; use fewer ranks for this grid")
        return StripPlan(rank, world, r0, r1, y0, y1, halo if nxt is not None else 0,
                         halo if prev is not None else 0, prev, nxt)


def exchange_halo(send: Optional[Tuple[torch.Tensor, torch.Tensor]], recv_shapes, plan: StripPlan, device,
                  group=None):
    """Post the halo send to `plan.next` and the halo receive from `plan.prev` (non-blocking). Returns
    (recv_num, recv_den, requests); call ``wait_all(requests)`` before using the received tensors."""
    reqs = []
    recv_num = recv_den = None
    ops = []
    if plan.prev is not None and recv_shapes is not None:
        recv_num = torch.empty(recv_shapes[0], dtype=torch.float32, device=device)
        recv_den = torch.empty(recv_shapes[1], dtype=torch.float32, device=device)
        ops += [dist.P2POp(dist.irecv, recv_num, plan.prev, group), dist.P2POp(dist.irecv, recv_den, plan.prev, group)]
    if plan.next is not None and send is not None:
        ops += [dist.P2POp(dist.isend, send[0], plan.next, group), dist.P2POp(dist.isend, send[1], plan.next, group)]
    if ops:
        reqs = dist.batch_isend_irecv(ops)
    return recv_num, recv_den, reqs


def wait_all(reqs) -> None:
    for r in reqs:
        r.wait()


def sharded_merge(band_merge: Callable, plan: StripPlan, k: int, wp: int, device, group=None):
    """Run the three band merges of one strip with the halo exchange in between.

    ``band_merge(rows=(y0,y1), raw=False, init=None)`` must merge THIS RANK'S tiles into canvas rows [y0,y1) and
    return either the normalised band (raw=False) or the un-normalised ``(num, den)`` pair (raw=True); it is
    ``TiledInferenceEngine.merge`` bound to the rank's predictions on GPUs, and an oracle merger in the CPU tests.
    Returns a list of ``(y0, y1, tensor)`` pieces covering the canvas rows this rank owns, top to bottom.
    """
    if plan.empty:
        return []
    pieces = []
    spill = None
    if plan.next is not None and plan.halo > 0:
        spill = band_merge(rows=(plan.y1, plan.y1 + plan.halo), raw=True, init=None)  # small; first, so it can fly
    recv_shapes = None
    top_rows = 0
    if plan.prev is not None and plan.halo_in > 0:
        top_rows = min(plan.halo_in, plan.y1 - plan.y0)
        recv_shapes = ((k, top_rows, wp), (top_rows, wp))
    recv_num, recv_den, reqs = exchange_halo(spill, recv_shapes, plan, device, group)
    # interior rows need nothing from the neighbour: they overlap the transfer
    y_mid = plan.y0 + top_rows
    interior = band_merge(rows=(y_mid, plan.y1), raw=False, init=None) if plan.y1 > y_mid else None
    wait_all(reqs)
    if top_rows:
        top = band_merge(rows=(plan.y0, y_mid), raw=False, init=(recv_num, recv_den))
        pieces.append((plan.y0, y_mid, top))
    if interior is not None:
        pieces.append((y_mid, plan.y1, interior))
    return pieces


class StripMerger:
    """``sharded_merge`` for a steady-state loop: buffers, P2P descriptors and launches prepared once, re-run per image.

    ``make_band(rows=(y0,y1), raw=False, init=None)`` returns a prepared band merge -- an object whose ``run()`` merges
    THIS RANK'S tiles into canvas rows [y0,y1) and returns the band (or the un-normalised ``(num, den)`` pair when
    raw=True), reading ``init`` (the receive buffers, refilled every step) at run time. On GPUs it is
    ``TiledInferenceEngine.merge_plan`` bound to the rank's prediction buffers. One ``run()`` then costs three
    ctypes calls and one ``batch_isend_irecv``; the arithmetic and its order are those of ``sharded_merge``.
    """

    def __init__(self, make_band: Callable, plan: StripPlan, k: int, wp: int, device, group=None):
        self.plan, self.group = plan, group
        self.spill = self.interior = self.top = None
        self.ops = []
        self.pieces: List[Tuple[int, int, object]] = []
        if plan.empty:
            return
        if plan.next is not None and plan.halo > 0:
            self.spill = make_band(rows=(plan.y1, plan.y1 + plan.halo), raw=True, init=None)
        top_rows = min(plan.halo_in, plan.y1 - plan.y0) if (plan.prev is not None and plan.halo_in > 0) else 0
        self.y_mid = plan.y0 + top_rows
        if top_rows:
            self.recv = (torch.empty((k, top_rows, wp), dtype=torch.float32, device=device),
                         torch.empty((top_rows, wp), dtype=torch.float32, device=device))
            self.top = make_band(rows=(plan.y0, self.y_mid), raw=False, init=self.recv)
        if plan.y1 > self.y_mid:
            self.interior = make_band(rows=(self.y_mid, plan.y1), raw=False, init=None)

    def run(self):
        plan = self.plan
        if plan.empty:
            return []
        ops = []
        if self.top is not None:
            ops += [dist.P2POp(dist.irecv, self.recv[0], plan.prev, self.group),
                    dist.P2POp(dist.irecv, self.recv[1], plan.prev, self.group)]
        if self.spill is not None:
            num, den = self.spill.run()  # small; first, so it can fly while the interior is merged
            ops += [dist.P2POp(dist.isend, num, plan.next, self.group), dist.P2POp(dist.isend, den, plan.next, self.group)]
        reqs = dist.batch_isend_irecv(ops) if ops else []
        pieces = []
        interior = self.interior.run() if self.interior is not None else None
        wait_all(reqs)
        if self.top is not None:
            pieces.append((plan.y0, self.y_mid, self.top.run()))
        if interior is not None:
            pieces.append((self.y_mid, plan.y1, interior))
        return pieces
