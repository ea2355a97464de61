"""Parity at BASELINE.json's full size (1M documents x 136 features x 30k queries), through
properties that do not need the oracle on the whole dataset plus an oracle check on a query
sample:
  * 400 sampled queries: per-query NDCG@10 of the exact kernels bit-identical to the oracle, and
    of the batched sweep bit-identical wherever the ranking is (here: everywhere);
  * batched sweep vs exact-order sweep over ALL queries: integer sums within 1e-9 of each other
    in the mean;
  * scale invariance: multiplying a weight vector by a power of two multiplies every score
    exactly, so rankings -- and the integer metric sums -- must not move;
  * a line-search candidate equal to the base weight reproduces evaluate_mean of the base;
  * row-permutation invariance of the mean (sums are order-independent integers).
"""
import numpy as np
import pytest

from tests.helpers import DevDataset, fx_sum, oracle_dataset, synth
from fastrank_b200.kernels import dense_query_index

pytestmark = pytest.mark.gpu

N, D, Q = 1_000_000, 136, 30_000
FX = float(1 << 40)


@pytest.fixture(scope="module")
def full():
    X, y, qid = synth(N, D, Q)
    qidx, nq = dense_query_index(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    yield X, y, qid, dev, nq
    dev.close()


def _line(orig, n):
    c = [0.0] + [orig - 0.05 * (2.0 ** k - 1) for k in range(1, 40)]
    return c[:n]


def test_full_size_sample_against_oracle_and_invariances(oracle, full):
    X, y, qid, dev, nq = full
    rng = np.random.default_rng(11)
    base = rng.uniform(-1, 1, size=(8, D))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int(v) for v in rng.integers(0, D, 8)]
    cands = [_line(base[r, fids[r]], 26) for r in range(8)]
    plan = dev.plan(0, 10)
    fast, pq_fast = None, None
    fast = plan.coord_sweeps(base, fids, cands, fast=True)
    exact = plan.coord_sweeps(base, fids, cands)
    # (1) batched vs exact-order sweep on all 30k queries
    diff = np.abs(fast - exact).astype(np.float64) / FX / nq
    assert diff.max() < 1e-9, diff.max()
    # (2) oracle on a sample of queries, per query, for full weight vectors of two candidates
    uq = np.unique(qid)
    pick = np.sort(rng.choice(len(uq), 400, replace=False))
    rows = np.nonzero(np.isin(qid, uq[pick]))[0]
    Xs, ys, qs = np.ascontiguousarray(X[rows]), y[rows], qid[rows]
    ods = oracle_dataset(oracle, Xs, ys, qs)
    W = []
    for r, k in ((0, 0), (3, 7), (7, 25)):
        w = base[r].copy()
        w[fids[r]] = cands[r][k]
        W.append(w)
    W = np.asarray(W)
    sums_lin, pq_lin = plan.eval_linear(W)
    qpos = {int(v): i for i, v in enumerate(uq)}  # dense query index == order of first appearance (qid sorted)
    for c in range(len(W)):
        exp = oracle.evaluate_scores(ods, oracle.score_linear(Xs, W[c]), "ndcg@10")
        got = pq_lin[c][[qpos[int(v)] for v in uq[pick]]]
        assert np.array_equal(got, exp)
    for c, (r, k) in enumerate(((0, 0), (3, 7), (7, 25))):
        assert int(exact[r, k]) == int(sums_lin[c])          # exact sweep == full rescoring
        assert abs(int(fast[r, k]) - int(sums_lin[c])) / FX / nq < 1e-9
    # (3) scale invariance (exact power-of-two scaling)
    s4, _ = plan.eval_linear(W * 4.0, per_query=False)
    s8, _ = plan.eval_linear(W * 0.125, per_query=False)
    assert s4.tolist() == sums_lin.tolist() and s8.tolist() == sums_lin.tolist()
    # (4) a candidate equal to the base weight reproduces the base's evaluate_mean
    same = plan.coord_sweeps(base[:2], fids[:2], [[base[0, fids[0]]], [base[1, fids[1]]]], fast=True)
    base_sums, _ = plan.eval_linear(base[:2], per_query=False)
    assert abs(int(same[0, 0]) - int(base_sums[0])) / FX / nq < 1e-9
    assert abs(int(same[1, 0]) - int(base_sums[1])) / FX / nq < 1e-9


def test_full_size_row_permutation_invariance(full):
    X, y, qid, dev, nq = full
    rng = np.random.default_rng(12)
    n = 200_000  # a 200k-row prefix (whole queries), permuted
    cut = int(np.searchsorted(qid, qid[n]))
    Xa, ya, qa = X[:cut], y[:cut], qid[:cut]
    perm = rng.permutation(cut)
    W = rng.normal(size=(2, D))
    out = []
    for Xi, yi, qi in ((Xa, ya, qa), (np.ascontiguousarray(Xa[perm]), ya[perm], qa[perm])):
        qidx, nq_i = dense_query_index(qi)
        d2 = DevDataset(Xi, yi.astype(np.float32), qidx, nq_i)
        try:
            sums, _ = d2.plan(0, 10).eval_linear(W, per_query=False)
            out.append(sums.tolist())
        finally:
            d2.close()
    # random float features: no score ties inside a query, so the id tie-break cannot matter
    assert out[0] == out[1]


@pytest.mark.parametrize("name,metric,depth", [("map", 1, -1), ("rr", 2, -1), ("ndcg", 0, -1), ("ndcg@3", 0, 3)])
def test_full_size_other_metrics_fast_vs_exact(oracle, full, name, metric, depth):
    X, y, qid, dev, nq = full
    rng = np.random.default_rng(13)
    base = rng.uniform(-1, 1, size=(3, D))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [5, 77, 135]
    cands = [[0.0, base[r, fids[r]] - 0.05, base[r, fids[r]] + 0.4, 3.0] for r in range(3)]
    plan = dev.plan(metric, depth)
    fast = plan.coord_sweeps(base, fids, cands, fast=True)
    exact = plan.coord_sweeps(base, fids, cands)
    assert (np.abs(fast - exact).astype(np.float64) / FX / nq).max() < 1e-9
    # and the exact kernel against the oracle on a query sample, per query
    uq = np.unique(qid)
    pick = np.sort(rng.choice(len(uq), 200, replace=False))
    rows = np.nonzero(np.isin(qid, uq[pick]))[0]
    Xs, ys, qs = np.ascontiguousarray(X[rows]), y[rows], qid[rows]
    ods = oracle_dataset(oracle, Xs, ys, qs)
    w = base[1].copy()
    w[fids[1]] = cands[1][2]
    _, pq = plan.eval_linear(w[None, :])
    exp = oracle.evaluate_scores(ods, oracle.score_linear(Xs, w), name)
    assert np.array_equal(pq[0][pick], exp)


def test_full_size_forest_scoring(oracle):
    """BASELINE configs[3] at full size: a depth-8 forest over 1M x 136.  Scores bit-identical to
    the oracle on a row sample; linearity of the ensemble sum (score([A, B]) == score(A) +
    score(B) exactly, weights 1.0) and idempotence over all rows."""
    import fastrank_b200 as fr

    X, y, qid = synth(N, D, Q)
    rng = np.random.default_rng(14)

    def tree(depth):
        if depth >= 8 or (depth > 2 and rng.random() < 0.05):
            return {"LeafNode": float(np.round(rng.uniform(0, 4), 3))}
        fid = int(rng.integers(0, D))
        split = float(np.quantile(X[:2000, fid], rng.uniform(0.1, 0.9)))
        return {"FeatureSplit": {"fid": fid, "split": split, "lhs": tree(depth + 1), "rhs": tree(depth + 1)}}

    a = [{"DecisionTree": tree(1)} for _ in range(30)]
    b = [{"DecisionTree": tree(1)} for _ in range(30)]
    ens = lambda ms: {"Ensemble": {"weights": [1.0] * len(ms), "models": ms}}  # noqa: E731
    ds = fr.CDataset.from_numpy(X, y, qid)
    sa = fr.CModel.from_dict(ens(a)).predict_dense(ds)
    sb = fr.CModel.from_dict(ens(b)).predict_dense(ds)
    sab = fr.CModel.from_dict(ens([ens(a), ens(b)])).predict_dense(ds)
    assert np.array_equal(sab, sa + sb)
    assert np.array_equal(fr.CModel.from_dict(ens(a)).predict_dense(ds), sa)
    rows = np.sort(rng.choice(N, 20000, replace=False))
    assert np.array_equal(sa[rows], oracle.score_model(np.ascontiguousarray(X[rows]), ens(a)))


def test_full_size_packed_kernel_equals_general_kernel(full, monkeypatch):
    """At full size: the register-packed NDCG@10 kernel (what train_model and bench.py launch) and
    the general tile kernel produce the same 8 x 51 integer sums, bit for bit, with direct
    publication of the sums on and off."""
    X, y, qid, dev, nq = full
    import bench

    plan = dev.plan(0, 10)
    base, fids, ga, gb = bench.step_inputs(5, D)
    pk = plan.pack_sweeps(base, fids, [a + b for a, b in zip(ga, gb)])
    got = {}
    for label, env in (("packed", {}), ("tile", {"FASTRANK_SWEEP_KERNEL": "tile"}), ("undirect", {"FASTRANK_DIRECT": "0"})):
        for k in ("FASTRANK_SWEEP_KERNEL", "FASTRANK_DIRECT"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got[label] = plan.coord_sweeps_packed(pk).copy()
        assert np.array_equal(plan.coord_sweeps_packed(pk), got[label])  # repeatable (the kernel re-zeroes its state)
    assert np.array_equal(got["packed"], got["tile"])
    assert np.array_equal(got["packed"], got["undirect"])


def test_long_list_shape_select_then_rank_equals_full_count(oracle, monkeypatch):
    """BASELINE configs[4]'s list shape (~120 documents per query) at a size the test can afford:
    the select-then-rank path of the packed kernel against the full count and the general kernel,
    bit for bit, and against the oracle on a few candidates."""
    import bench

    n, q = 400_000, 3_300
    X, y, qid = synth(n, D, q, seed=77)
    qidx, nq = dense_query_index(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    try:
        plan = dev.plan(0, 10)
        assert dev.lib.fr_dev_plan_tile_documents(plan.ptr) == 256
        base, fids, ga, gb = bench.step_inputs(2, D)
        cands = [a + b for a, b in zip(ga, gb)]
        got = {}
        for label, env in (("pruned", {}), ("full", {"FASTRANK_PRUNE_MIN": "0"}), ("tile", {"FASTRANK_SWEEP_KERNEL": "tile"})):
            for k in ("FASTRANK_SWEEP_KERNEL", "FASTRANK_PRUNE_MIN"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got[label] = plan.coord_sweeps(base, fids, cands, fast=True).copy()
        assert np.array_equal(got["pruned"], got["full"])
        assert np.array_equal(got["pruned"], got["tile"])
        ods = oracle_dataset(oracle, X, y, qid)
        for r, k in ((0, 0), (4, 30), (7, 50)):
            w = base[r].copy()
            w[fids[r]] = cands[r][k]
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10")
            assert abs(int(got["pruned"][r, k]) - fx_sum(exp)) / FX / nq < 1e-9
    finally:
        dev.close()
