"""Multi-GPU tiled inference over "ragged strips": the tile grid is cut into `world` contiguous row-major tile RANGES of
(almost) equal length, one per rank, and the overlap between neighbouring ranges is summed through NVLink peer stores
issued by the merge kernel itself -- no NCCL call on the data path.

New relative to the reference (its TileMerger is single-device, tiles.py:306-308). Compared with the
row strips of ``strips.py`` (round 1):

* balance: 361 tiles over 8 ranks are 46/45 tiles each (ceiling 0.98) instead of 3/2 tile ROWS (57/38 tiles, ceiling 0.79);
* exchange: a canvas block whose covering tiles straddle a range boundary is accumulated by the lower rank first, stored
  un-normalised into the NEXT rank's seed buffer (CUDA-IPC mapped peer memory, ``tb_peer_*``), and finished there starting
  from that seed. The accumulation order is the single-GPU tile order, so the sharded result is bit-identical;
* ordering: two uint64 flags per boundary (``ready``: seeds of epoch e are complete; ``ack``: seeds of epoch e were consumed,
  the double buffer may be reused). FUSED form (persistent TMA merge kernel, TTA groups with >= 2 views; the default there):
  ONE merge launch per rank and step -- the kernel walks its blocks in three ordered passes (spill, final, seeded), one thread
  per pipeline waits for ``ack`` / ``ready`` in front of the pass that needs it, and the LAST pipeline of the grid to finish a
  pass publishes the neighbour's flag. UNFUSED form (single-view kernel, or ``fused=False``): three merge launches, the spill
  and its ``ready`` signal on a side stream, one-thread ``tb_peer_signal/wait`` kernels in stream order. The host never blocks.
  Measured (profiles/r2_scaling_summary.md): cfg 2 at 8 GPUs 0.570 ms fused against 0.602 (two launches) / 0.658 (three).

Ownership rule: a canvas pixel belongs to the rank that owns its LAST covering tile (row-major order).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

__all__ = ["partition_tiles", "RaggedPlan", "PeerMailbox", "RaggedMerger", "block_class", "PH_FINAL", "PH_SPILL",
           "PH_SEEDED_FINAL", "PH_SEEDED_SPILL"]

PH_FINAL, PH_SPILL, PH_SEEDED_FINAL, PH_SEEDED_SPILL = 1, 2, 4, 8


def partition_tiles(n_tiles: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, as-even-as-possible row-major tile ranges (earlier ranks get the extra tile)."""
    base, extra = divmod(n_tiles, world)
    out, t = [], 0
    for g in range(world):
        n = base + (1 if g < extra else 0)
        out.append((t, t + n))
        t += n
    return out


def block_class(t0: int, t1: int, nx: int, ra: int, rb: int, ca: int, cb: int) -> int:
    """Host twin of ``peer_block_class`` (csrc/tiles_args.cuh): class of a canvas block covered by tile rows ra..rb and
    columns ca..cb for the rank that owns tiles [t0,t1); 0 = none of its business."""
    own = any(r * nx + ca < t1 and r * nx + cb >= t0 for r in range(ra, rb + 1))
    if not own:
        return 0
    seeded = ra * nx + ca < t0
    spill = rb * nx + cb >= t1
    return (PH_SEEDED_SPILL if spill else PH_SEEDED_FINAL) if seeded else (PH_SPILL if spill else PH_FINAL)


@dataclass
class RaggedPlan:
    rank: int
    world: int
    nx: int
    ny: int
    t0: int
    t1: int
    prev: Optional[int]
    next: Optional[int]
    seed_y0: int      # canvas rows [seed_y0, seed_y0 + window) can hold seeds from the previous rank
    spill_y0: int     # = the next rank's seed_y0
    window: int       # rows of every seed window (= tile height, clipped windows are simply not fully used)
    rows: Tuple[int, int]  # canvas rows this rank's tiles touch
    chained: bool     # blocks that are seeded AND spill exist (ranges shorter than a covering set): signal late

    @staticmethod
    def build(nx: int, ny: int, tile_h: int, tile_w: int, step_h: int, step_w: int, canvas_h: int, rank: int,
              world: int) -> "RaggedPlan":
        n = nx * ny
        parts = partition_tiles(n, world)
        # every rank validates EVERY range (the partition is deterministic), so that all ranks raise together instead of
        # some of them blocking on a flag that will never be written
        if world > 1 and min(b - a for a, b in parts) < nx:
            raise ValueError(f"ragged strips need at least one tile row ({nx} tiles) per rank; {n} tiles over {world} ranks "
                             f"gives {min(b - a for a, b in parts)}: use at most {max(1, ny)} ranks for this grid")
        t0, t1 = parts[rank]
        cover_r = -(-tile_h // step_h)
        cover_c = -(-tile_w // step_w)
        prev = rank - 1 if rank > 0 else None
        nxt = rank + 1 if rank + 1 < world else None
        seed_y0 = (t0 // nx) * step_h
        spill_y0 = (parts[rank + 1][0] // nx) * step_h if nxt is not None else 0
        r_first, r_last = t0 // nx, (t1 - 1) // nx
        rows = (r_first * step_h, min(canvas_h, r_last * step_h + tile_h))
        chained = (t1 - t0) < (cover_r - 1) * nx + cover_c
        return RaggedPlan(rank, world, nx, ny, t0, t1, prev, nxt, seed_y0, spill_y0, tile_h, rows, chained)


class PeerMailbox:
    """This rank's receive side: two seed buffers (double-buffered over epochs) + the flag words, in ONE cudaMalloc'ed
    block that is exported to the neighbours with CUDA IPC.

        word 0    ready     written by the PREVIOUS rank: seeds of epoch `ready` are complete (buffer epoch % 2)
        word 1    ack       written by the NEXT rank: it has consumed the seeds of epoch `ack` that we stored into ITS mailbox
        word 2    status    local: non-zero after a flag wait timed out
        word 3-5  counters  local: arrival counters of the fused single-launch mode (only ever grow)
    """

    HEADER = 256

    def __init__(self, device: torch.device, k: int, rows: int, wp: int):
        from .. import _native as N

        self.N = N
        self.device, self.k, self.rows, self.wp = device, k, rows, wp
        self.num_floats = k * rows * wp
        self.den_floats = rows * wp
        self.buf_bytes = (self.num_floats + self.den_floats) * 4
        self.bytes = self.HEADER + 2 * self.buf_bytes
        ptr = C.c_void_p()
        with torch.cuda.device(device):
            N.check(N.lib().tb_peer_alloc(self.bytes, C.byref(ptr)), "tb_peer_alloc")
        self.ptr = ptr.value
        self._opened: List[int] = []

    # addresses inside a mailbox at `base` (ours or a mapped neighbour's)
    def ready(self, base=None): return (base or self.ptr) + 0
    def ack(self, base=None): return (base or self.ptr) + 8
    def status(self, base=None): return (base or self.ptr) + 16
    def counters(self): return self.ptr + 24
    def num(self, parity: int, base=None): return (base or self.ptr) + self.HEADER + parity * self.buf_bytes
    def den(self, parity: int, base=None): return self.num(parity, base) + self.num_floats * 4

    def handle(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            self.N.check(self.N.lib().tb_peer_export(self.ptr, buf), "tb_peer_export")
        return bytes(buf)

    def open(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64)(*handle)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            self.N.check(self.N.lib().tb_peer_open(buf, C.byref(ptr)), "tb_peer_open")
        self._opened.append(ptr.value)
        return ptr.value

    def read_status(self) -> int:
        """Synchronise the device and read the local status word (0 = no flag wait ever timed out)."""
        torch.cuda.synchronize(self.device)
        val = C.c_uint64(0)
        with torch.cuda.device(self.device):
            self.N.check(self.N.lib().tb_peer_read(self.status(), C.byref(val)), "tb_peer_read")
        return int(val.value)

    def close(self):
        lib = self.N.lib()
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib.tb_peer_close(p)
            self._opened = []
            if self.ptr:
                lib.tb_peer_free(self.ptr)
                self.ptr = 0


class RaggedMerger:
    """Steady-state sharded merge of one rank: everything (mailboxes, IPC mappings, the three prepared merge launches)
    is set up once; ``run()`` enqueues one step on the current stream and returns immediately.

    ``engine``: TiledInferenceEngine of the WHOLE image; ``preds``: this rank's predictions, one tensor
    [A*(t1-t0),K,T,T] or the list of the model's per-batch outputs for tiles [t0,t1); ``out``: [K,H,W] float32 (or the
    uint8 label / mask tensor with ``output=``) -- only the pixels this rank owns are written.
    ``exchange(obj) -> list`` gathers a picklable object from every rank (default: torch.distributed.all_gather_object).
    """

    def __init__(self, engine, plan: RaggedPlan, preds, out: torch.Tensor, output: str = "logits", threshold: float = 0.5,
                 exchange=None, timeout_ms: int = 5000, fused: Optional[bool] = None):
        from .. import _native as N

        self.N, self.engine, self.plan, self.out = N, engine, plan, out
        self.timeout_ms = int(timeout_ms)
        self.epoch = 0
        # fused: ONE merge launch per step, the kernel orders spill -> final -> seeded itself and handles the flags (persistent
        # TMA kernel only, i.e. TTA groups with >= 2 views); otherwise two launches + stream-ordered flag kernels
        if fused is None and os.environ.get("TB200_RAGGED_FUSED") in ("0", "1"):  # A/B switch for the benchmarks
            fused = os.environ["TB200_RAGGED_FUSED"] == "1" and engine.views > 1
        self.fused = (engine.views > 1) if fused is None else bool(fused)
        if self.fused and engine.views == 1:
            raise ValueError("the fused single-launch mode needs the persistent TMA merge kernel (a TTA group with >= 2 views)")
        dev = engine.device
        hp, wp = engine.slicer.target_shape
        k = engine.channels
        self.mailbox = PeerMailbox(dev, k, plan.window, wp)
        if exchange is None:
            import torch.distributed as dist

            def exchange(obj):
                got = [None] * dist.get_world_size()
                dist.all_gather_object(got, obj)
                return got

        handles = exchange(self.mailbox.handle()) if plan.world > 1 else [None]
        self.next_base = self.mailbox.open(handles[plan.next]) if plan.next is not None else None
        self.prev_base = self.mailbox.open(handles[plan.prev]) if plan.prev is not None else None
        mb = self.mailbox
        y0, y1 = plan.rows
        win = plan.window

        def peer(mask, parity):
            return dict(phase_mask=mask, t0=plan.t0, t1=plan.t1,
                        seed_num=mb.num(parity), seed_den=mb.den(parity), seed_y0=plan.seed_y0, seed_rows=win,
                        spill_num=mb.num(parity, self.next_base) if self.next_base else None,
                        spill_den=mb.den(parity, self.next_base) if self.next_base else None,
                        spill_y0=plan.spill_y0, spill_rows=win)

        def launch(mask, rows, parity):
            rows = (max(rows[0], y0), min(rows[1], y1))
            if rows[1] <= rows[0]:
                return None
            return engine.merge_plan(preds, out=out, rows=rows, output=output, threshold=threshold,
                                     first_tile=plan.t0, peer=peer(mask, parity))

        # Unfused mode (single-view kernel, or fused=False): three launches -- SPILL, FINAL, SEEDED -- with the SPILL launch and
        # its `ready` signal on a SIDE stream: the un-normalised overlap sums of a boundary are (K+1)(T-S)Wp*4 bytes of NVLink
        # peer stores (26 MB at cfg 2, 84 MB at cfg 5 = ~93 us at 900 GB/s), which would otherwise sit exposed in front of the
        # merge; on its own stream it overlaps the FINAL launch (HBM-bound, no dependence on anybody). The SEEDED launch follows
        # FINAL behind the wait for the previous rank's flag -- by then that rank's spill has long landed.
        self.spill = self.final = self.seeded = self.one = (None, None)
        self._fused_plans = []
        self.side = None
        if self.fused:
            # ONE launch per step: every block class of this rank in the mask, the kernel walks them in three ordered passes
            # (spill, final, seeded) and waits for / publishes the flags itself (csrc/tiles.cu, merge_all_tma_kernel)
            mask = PH_FINAL | (PH_SPILL if plan.next is not None else 0)
            if plan.prev is not None:
                mask |= PH_SEEDED_FINAL | (PH_SEEDED_SPILL if plan.next is not None else 0)
            plans = []
            for parity in (0, 1):
                pd = peer(mask, parity)
                pd.update(fused=1, chained=int(plan.chained), has_next=int(plan.next is not None),
                          has_prev=int(plan.prev is not None), epoch=1, counters=mb.counters(), ready_flag=mb.ready(),
                          ack_flag=mb.ack(), next_ready=mb.ready(self.next_base) if self.next_base else None,
                          prev_ack=mb.ack(self.prev_base) if self.prev_base else None, status=mb.status(),
                          timeout_ms=self.timeout_ms)
                plans.append(engine.merge_plan(preds, out=out, rows=(y0, y1), output=output, threshold=threshold,
                                               first_tile=plan.t0, peer=pd))
                self._fused_plans.append(pd["struct"])  # `.epoch` is advanced before every run
            self.one = tuple(plans)
            return
        self.side = torch.cuda.Stream(dev) if plan.next is not None else None
        self._ev_start = torch.cuda.Event()
        self._ev_side = torch.cuda.Event()
        if plan.next is not None:
            with torch.cuda.stream(self.side):  # PreparedLaunch binds the stream that is current when the plan is built
                self.spill = tuple(launch(PH_SPILL, (plan.spill_y0, plan.spill_y0 + win), p) for p in (0, 1))
        self.final = tuple(launch(PH_FINAL, (y0, y1), p) for p in (0, 1))
        if plan.prev is not None:  # (the last rank has nobody to spill to -- and no block of its range can need it)
            seeded_mask = PH_SEEDED_FINAL | (PH_SEEDED_SPILL if plan.next is not None else 0)
            self.seeded = tuple(launch(seeded_mask, (plan.seed_y0, plan.seed_y0 + win), p) for p in (0, 1))

    def run(self, marks=None):
        """One step (epoch). Fused: one launch. Otherwise: side stream: [wait ack] spill, signal(ready);  main stream: final,
        wait(ready), seeded, [signal(ready) if chained], signal(ack), join the side stream."""
        N, plan, mb = self.N, self.plan, self.mailbox
        lib = N.lib()
        self.epoch += 1
        e, p = self.epoch, self.epoch % 2
        dev = self.engine.device
        if self.fused:  # one launch: the kernel waits / publishes the flags itself
            self._fused_plans[p].epoch = e
            self.one[p].run()
            if marks:
                marks[0].record(torch.cuda.current_stream(dev))
            return self.out
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            stream = main.cuda_stream
            if plan.next is not None:
                self._ev_start.record(main)
                self.side.wait_event(self._ev_start)  # the predictions of this step are complete
                side = self.side.cuda_stream
                if e > 2:  # the next rank must have consumed the seeds of epoch e-2 before their buffer is overwritten
                    N.check(lib.tb_peer_wait(mb.ack(), e - 2, self.timeout_ms, mb.status(), side), "tb_peer_wait(ack)")
                if self.spill[p] is not None:
                    self.spill[p].run()
                if not plan.chained:
                    N.check(lib.tb_peer_signal(mb.ready(self.next_base), e, side), "tb_peer_signal(ready)")
                    self._ev_side.record(self.side)
            if self.final[p] is not None:
                self.final[p].run()
            if marks:  # (bench instrumentation: an event between the two halves of the step)
                marks[0].record(main)
            if plan.prev is not None:
                N.check(lib.tb_peer_wait(mb.ready(), e, self.timeout_ms, mb.status(), stream), "tb_peer_wait(ready)")
                if self.seeded[p] is not None:
                    self.seeded[p].run()
            if plan.next is not None:
                if plan.chained:  # seeded-AND-spilling blocks exist: the flag can only follow them (side-stream spill first)
                    self._ev_side.record(self.side)
                    main.wait_event(self._ev_side)
                    N.check(lib.tb_peer_signal(mb.ready(self.next_base), e, stream), "tb_peer_signal(ready)")
                else:
                    main.wait_event(self._ev_side)  # the step is complete only when its spill has been handed over
            if plan.prev is not None:
                N.check(lib.tb_peer_signal(mb.ack(self.prev_base), e, stream), "tb_peer_signal(ack)")
        return self.out

    def check(self):
        """Synchronise and raise if any flag wait of this rank timed out (a neighbour died or fell > timeout behind)."""
        st = self.mailbox.read_status()
        if st:
            raise RuntimeError(f"rank {self.plan.rank}: a peer flag wait timed out (last value seen {st - 1}, epoch {self.epoch})")

    def close(self):
        torch.cuda.synchronize(self.engine.device)
        self.mailbox.close()
