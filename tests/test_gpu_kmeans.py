"""robustkmeans (NMFkCluster.jl:138-289) on the GPU against the oracle's restatement of Clustering.kmeans with
cosine distance: the same seeds per repeat on both sides -> identical best repeat, assignments after sortclustering, counts,
iteration count; total cost / centres / silhouettes to rounding.  Plus the reference's own unit expectation
(test/test_cluster_unit.jl:6-18: two well separated groups give labels {1, 2})."""
import numpy as np
import pytest

import nmfk_b200 as nb
from oracle import nmfk_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context()
    yield c
    c.close()


def _blobs(rng, d, per, k, noise):
    cent = rng.random((d, k))
    X = np.hstack([cent[:, [c]] * (1 + noise * rng.standard_normal((d, per[c]))) for c in range(k)])
    return np.abs(X)[:, rng.permutation(sum(per))]


@pytest.mark.parametrize("d,per,k,repeats,noise", [(6, (40, 25, 10), 3, 12, 0.05), (10, (60, 50, 40, 30), 4, 20, 0.15),
                                                   (3, (30, 30), 2, 8, 0.3), (24, (200, 150, 100, 50, 25), 5, 16, 0.2),
                                                   (5, (50, 40, 30), 4, 10, 0.25), (8, (64,), 1, 3, 0.1)])
def test_robustkmeans_matches_oracle(ctx, d, per, k, repeats, noise):
    rng = np.random.default_rng(d * 100 + k)
    X = _blobs(rng, d, per, len(per), noise)
    seeds = np.stack([nb.kmeanspp_seeds(X, k, rng) for _ in range(repeats)])
    det = {}
    res, sil = nb.robustkmeans(X, k, repeats, seeds=seeds, compute_silhouettes_flag=True, ctx=ctx, details=det)
    ref, rsil = o.robustkmeans(X, k, seeds, compute_silhouettes_flag=True)
    costs = [o.kmeans_lloyd(X, k, s)["totalcost"] for s in seeds]
    assert det["best_repeat"] == int(np.argmin(costs))
    assert np.array_equal(res.assignments, ref["assignments"])
    assert np.array_equal(res.counts, ref["counts"]) and res.iterations == ref["iterations"] and res.converged == ref["converged"]
    assert abs(res.totalcost - ref["totalcost"]) <= 1e-10 * max(ref["totalcost"], 1e-12)
    assert np.allclose(res.centers, ref["centers"], rtol=1e-10) and np.allclose(res.costs, ref["costs"], rtol=1e-8, atol=1e-13)
    assert np.allclose(sil, rsil, atol=1e-9)
    # sortclustering: labels ranked by cluster size
    assert list(res.counts) == sorted(res.counts, reverse=True)


def test_reference_unit_case_two_groups(ctx):
    """test/test_cluster_unit.jl:6-18: X = [group near e1 | group near e2], robustkmeans(X, 2, 5) yields labels {1, 2} that separate
    the groups; the range form picks a k and returns a result."""
    rng = np.random.default_rng(1)
    a = np.array([[1.0], [0.05]]) + 0.02 * rng.random((2, 6))
    b = np.array([[0.05], [1.0]]) + 0.02 * rng.random((2, 6))
    X = np.hstack([a, b])
    res = nb.robustkmeans(X, 2, 5, seed=3, ctx=ctx)
    assert set(res.assignments.tolist()) == {1, 2}
    assert len(set(res.assignments[:6].tolist())) == 1 and len(set(res.assignments[6:].tolist())) == 1
    assert res.assignments[0] != res.assignments[-1]
    best = nb.robustkmeans(X, range(2, 5), 5, seed=3, ctx=ctx)
    assert best is not None and best.assignments.shape == (12,)
    assert nb.robustkmeans(X, range(12, 14), 5, ctx=ctx) is None  # krange[1] >= size(X, 2) (:139-142)
