"""CPU: the ragged-strip partition of the multi-GPU merge (distributed/ragged.py). The block classification is the same
predicate the kernels evaluate (peer_block_class in csrc/tiles_args.cuh); here it is exercised exhaustively on small grids:
every canvas block is finished by exactly one rank, partial sums travel down an unbroken chain of ranks in tile order, and
every spill / seed lands inside the row windows the mailboxes are sized for."""
import itertools

import pytest

from pytorch_toolbelt_b200.distributed.ragged import (PH_FINAL, PH_SEEDED_FINAL, PH_SEEDED_SPILL, PH_SPILL, RaggedPlan,
                                                      block_class, partition_tiles)


def covering(y0, T, S, n):
    """tile rows (or columns) covering the 32-pixel block starting at y0"""
    ra = 0 if y0 - T + 1 <= 0 else (y0 - T + S) // S
    rb = min((y0 + 31) // S, n - 1)
    return ra, rb


def test_partition_tiles():
    assert partition_tiles(361, 8) == [(0, 46), (46, 91), (91, 136), (136, 181), (181, 226), (226, 271), (271, 316), (316, 361)]
    for n, w in itertools.product((1, 7, 361, 1521), (1, 2, 3, 8)):
        parts = partition_tiles(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


@pytest.mark.parametrize("nx,ny,T,S,world", [
    (19, 19, 512, 256, 8), (19, 19, 512, 256, 2), (39, 39, 512, 256, 8), (39, 39, 1024, 512, 8), (5, 7, 128, 32, 3),
    (4, 9, 128, 64, 4), (6, 6, 64, 64, 5), (3, 8, 96, 32, 8), (15, 15, 256, 128, 4),
])
def test_every_block_is_finished_once_and_chains_are_unbroken(nx, ny, T, S, world):
    hp, wp = (ny - 1) * S + T, (nx - 1) * S + T
    plans = [RaggedPlan.build(nx, ny, T, T, S, S, hp, r, world) for r in range(world)]
    assert [p.t0 for p in plans] == [a for a, _ in partition_tiles(nx * ny, world)]
    for y0 in range(0, hp, 32):
        ra, rb = covering(y0, T, S, ny)
        for x0 in range(0, wp, 32):
            ca, cb = covering(x0, T, S, nx)
            tiles = [r * nx + c for r in range(ra, rb + 1) for c in range(ca, cb + 1)]
            classes = [block_class(p.t0, p.t1, nx, ra, rb, ca, cb) for p in plans]
            finals = [g for g, c in enumerate(classes) if c in (PH_FINAL, PH_SEEDED_FINAL)]
            assert len(finals) == 1, (y0, x0, classes)
            owner = finals[0]
            assert plans[owner].t0 <= max(tiles) < plans[owner].t1  # the owner holds the LAST covering tile
            # emulate the chain with exact integers: partial = sum of the tile ids accumulated so far, in tile order
            carried, order = None, []
            for g, (p, c) in enumerate(zip(plans, classes)):
                if c == 0:
                    assert carried is None or g > owner, "a rank in the middle of a chain must take part"
                    continue
                mine = [t for t in tiles if p.t0 <= t < p.t1]
                assert mine, "a participating rank has tiles of the block"
                if c in (PH_SEEDED_FINAL, PH_SEEDED_SPILL):
                    assert carried is not None, "seeded blocks are preceded by a spill of the previous rank"
                    assert p.seed_y0 <= y0 and y0 + 32 <= p.seed_y0 + p.window  # inside this rank's seed window
                else:
                    assert carried is None
                order += mine
                if c in (PH_SPILL, PH_SEEDED_SPILL):
                    assert p.next == g + 1 and p.spill_y0 <= y0 and y0 + 32 <= p.spill_y0 + p.window
                    assert p.spill_y0 == plans[g + 1].seed_y0
                    if c == PH_SEEDED_SPILL:
                        assert p.chained
                    carried = sum(order)
                else:
                    carried = None
                assert p.rows[0] <= y0 and y0 + 32 <= p.rows[1]
            assert order == sorted(tiles), "tiles are accumulated exactly once, in the reference's row-major order"


def test_too_many_ranks_raise_on_every_rank():
    for rank in range(8):
        with pytest.raises(ValueError):
            RaggedPlan.build(19, 3, 512, 512, 256, 256, 1024, rank, 8)  # 57 tiles / 8 ranks < one tile row each
    p = RaggedPlan.build(19, 19, 512, 512, 256, 256, 5120, 3, 8)
    assert (p.t0, p.t1, p.prev, p.next, p.seed_y0, p.window, p.chained) == (136, 181, 2, 4, 7 * 256, 512, False)
