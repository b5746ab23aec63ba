"""Property tests (hypothesis) of the host-side integer logic: the tile grid against the oracle's independent
restatement of tiles.py:97-142 on ragged geometries, and the strip partition of the multi-GPU path (exact cover,
consistent halos). Bit-exact: these are coordinates and indices."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import tiles_oracle as TO
from pytorch_toolbelt_b200.distributed.strips import StripPlan, partition_tile_rows
from pytorch_toolbelt_b200.inference.tiles import ImageSlicer


@st.composite
def geometries(draw):
    th, tw = draw(st.integers(1, 96)), draw(st.integers(1, 96))
    sh, sw = draw(st.integers(1, th)), draw(st.integers(1, tw))
    h, w = draw(st.integers(1, 400)), draw(st.integers(1, 400))
    margin = draw(st.one_of(st.just(0), st.integers(1, 40),
                            st.tuples(st.integers(0, 30), st.integers(0, 30), st.integers(0, 30), st.integers(0, 30))))
    return h, w, (th, tw), (sh, sw), margin


@settings(max_examples=300, deadline=None)
@given(geometries())
def test_tile_grid_equals_oracle_on_ragged_geometries(geom):
    h, w, tile, step, margin = geom
    want = TO.tile_grid(h, w, tile, step, margin)
    if len(want["crops"]) == 0:  # margins too small for a single tile: nothing to compare
        return
    s = ImageSlicer((h, w, 3), tile, step, image_margin=margin, weight="mean")
    assert s.crops.dtype == np.int64 and np.array_equal(s.crops, want["crops"])
    assert np.array_equal(s.bbox_crops, want["bbox_crops"])
    assert (s.margin_left, s.margin_right, s.margin_top, s.margin_bottom) == tuple(want["margins"])
    assert tuple(s.target_shape) == tuple(want["target_shape"])
    ny, nx = s.grid_shape
    assert ny * nx == len(s.crops)
    # the regular lattice the fused kernels rely on: tile (r, c) sits at (c*Sx, r*Sy)
    lattice = np.array([(c * step[1], r * step[0]) for r in range(ny) for c in range(nx)])
    assert np.array_equal(s.crops[:, :2], lattice)


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 64), st.integers(1, 16), st.integers(2, 64), st.data())
def test_strip_plans_cover_the_canvas_exactly(ny, world, tile, data):
    step = data.draw(st.integers(max(1, (tile + 1) // 2), tile))  # overlap <= 50 %: a spill never crosses two strips
    canvas = (ny - 1) * step + tile
    parts = partition_tile_rows(ny, world)
    assert parts[0][0] == 0 and parts[-1][1] == ny and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    try:
        plans = [StripPlan.build(ny, tile, step, canvas, g, world) for g in range(world)]
    except ValueError:
        return  # a one-row strip narrower than the overlap is rejected loudly, by design
    covered = np.zeros(canvas, dtype=np.int32)
    live = [p for p in plans if not p.empty]
    for p in live:
        covered[p.y0:p.y1] += 1
    assert (covered == 1).all()
    for a, b in zip(live, live[1:]):
        assert a.next == b.rank and b.prev == a.rank and a.y1 == b.y0
        assert a.halo == b.halo_in == tile - step
    assert live[0].prev is None and live[0].halo_in == 0 and live[-1].next is None and live[-1].halo == 0
