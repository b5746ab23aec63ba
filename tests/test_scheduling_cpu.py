"""CPU: tile -> rank scheduling helpers (utils/scheduling.py) against known answers produced by the reference
(utils/bucket_assignment.py, utils/distributed.py:240-309) and, when baseline/_ref is installed, against the reference live."""
import numpy as np
import pytest

from baseline import reference_arm as RA
from pytorch_toolbelt_b200.utils import scheduling as S


def test_known_answers_from_the_reference():
    costs = np.array([5, 5, 5, 1, 1, 9, 9, 2, 2, 2])
    assert S.filler_bucket_assignment(costs, 3).tolist() == [2, 2, 0, 1, 2, 1, 0, 1, 1, 2]
    assert S.split_across_nodes(list("abcdefghij"), 3, 1, cost=costs) == ["d", "f", "h", "i"]
    assert S.split_across_nodes(list(range(10)), 4, 0) == [0, 1] and S.split_across_nodes(list(range(10)), 4, 3) == [7, 8, 9]
    assert S.split_across_nodes(list(range(10)), 1, 0) == list(range(10))
    with pytest.raises(RuntimeError):
        S.split_across_nodes([1, 2, 3], 2, 0, cost=[1, 2])
    loads = [sum(costs[S.filler_bucket_assignment(costs, 3) == b]) for b in range(3)]
    assert max(loads) - min(loads) <= 2 and S.compute_bucket_imbalance_score(costs, S.filler_bucket_assignment(costs, 3)) < 1.0


@pytest.mark.skipif(not RA.available(), reason="baseline/_ref not installed")
def test_against_the_installed_reference():
    RA.load()
    from pytorch_toolbelt.utils import bucket_assignment as RB
    from pytorch_toolbelt.utils.distributed import split_across_nodes as rsplit

    rng = np.random.default_rng(0)
    for n, k in [(10, 4), (361, 8), (1521, 8), (7, 3)]:
        costs = rng.random(n) * 10
        assert np.array_equal(RB.filler_bucket_assignment(costs, k), S.filler_bucket_assignment(costs, k))
        assert np.array_equal(RB.naive_bucket_assignment(costs, k), S.naive_bucket_assignment(costs, k))
        coll = list(range(n))
        for r in range(k):
            assert rsplit(coll, k, r) == S.split_across_nodes(coll, k, r)
            assert rsplit(coll, k, r, cost=costs) == S.split_across_nodes(coll, k, r, cost=costs)
            assert rsplit(coll, k, r, cost=costs, method="naive") == S.split_across_nodes(coll, k, r, cost=costs, method="naive")
