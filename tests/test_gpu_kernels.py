"""Parity of the CUDA kernels (through the fr_dev_* C ABI) against the CPU oracle:
per-query metric values bit-exact, fixed-point sums exactly equal."""
import numpy as np
import pytest

from tests.helpers import DevDataset, dense_qidx, fx_sum, oracle_dataset, synth

pytestmark = pytest.mark.gpu

METRICS = [("ndcg@10", 0, 10), ("ndcg", 0, -1), ("ndcg@3", 0, 3), ("map", 1, -1), ("rr", 2, -1)]


def _mk(oracle, n, d, q, seed, shuffle=False):
    X, y, qid = synth(n, d, q, seed=seed, shuffle_rows=shuffle)
    ods = oracle_dataset(oracle, X, y, qid)
    qidx, nq = dense_qidx(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    return X, y, qid, ods, dev


@pytest.mark.parametrize("shuffle,tma", [(False, False), (True, False), (False, True), (True, True)])
def test_linear_batch_bit_exact(oracle, shuffle, tma, monkeypatch):
    # tma: the same evaluation with X staged by TMA bulk copies (linear_tma_kernel, FASTRANK_TMA_EVAL=1,
    # which serves up to 8 vectors per pass: the 11 vectors go as 8 + 2 + 1)
    if tma:
        monkeypatch.setenv("FASTRANK_TMA_EVAL", "1")
    else:
        monkeypatch.delenv("FASTRANK_TMA_EVAL", raising=False)
    X, y, qid, ods, dev = _mk(oracle, 6000, 24, 180, seed=3, shuffle=shuffle)
    rng = np.random.default_rng(0)
    W = rng.normal(size=(11, 24))
    W[3, :] = 0.0                      # all scores tie -> pure tie-break order
    W[4, :] = 0.0
    W[4, 2] = 1.0                      # integer column: heavy ties
    try:
        for name, metric, depth in METRICS:
            plan = dev.plan(metric, depth)
            if tma:
                parts = [plan.eval_linear(W[a:b]) for a, b in ((0, 8), (8, 10), (10, 11))]
                sums = np.concatenate([p[0] for p in parts])
                pq = np.concatenate([p[1] for p in parts])
            else:
                sums, pq = plan.eval_linear(W)
            for c in range(W.shape[0]):
                exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), name)
                assert np.array_equal(pq[c], exp), (name, c, np.abs(pq[c] - exp).max())
                assert int(sums[c]) == fx_sum(exp)
    finally:
        dev.close()


def test_ragged_and_long_queries(oracle):
    # query lengths 1 .. ~700 in one dataset: exercises every tile size up to 1024
    rng = np.random.default_rng(9)
    lens = [1, 2, 3, 31, 32, 33, 64, 127, 128, 129, 255, 257, 600, 700, 5, 1]
    qid = np.concatenate([np.full(l, 100 + i) for i, l in enumerate(lens)]).astype(np.int64)
    n = len(qid)
    X = rng.normal(size=(n, 7)).astype(np.float32)
    X[:, 3] = rng.integers(0, 3, n)
    y = rng.integers(0, 5, n).astype(np.float64)
    y[qid == 102] = 0.0                 # a query without relevant documents
    ods = oracle_dataset(oracle, X, y, qid)
    qidx, nq = dense_qidx(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    W = rng.normal(size=(5, 7))
    W[1, :] = 0
    W[1, 3] = -1.0
    try:
        for name, metric, depth in METRICS:
            plan = dev.plan(metric, depth)
            sums, pq = plan.eval_linear(W)
            for c in range(W.shape[0]):
                exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), name)
                assert np.array_equal(pq[c], exp), (name, c)
                assert int(sums[c]) == fx_sum(exp)
    finally:
        dev.close()


def test_negative_and_fractional_gains(oracle):
    rng = np.random.default_rng(4)
    n = 900
    qid = np.sort(rng.integers(0, 40, n)).astype(np.int64)
    X = rng.normal(size=(n, 5)).astype(np.float32)
    y = rng.choice([-1.0, 0.0, 0.5, 1.0, 2.5], size=n)
    ods = oracle_dataset(oracle, X, y, qid)
    qidx, nq = dense_qidx(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    W = rng.normal(size=(3, 5))
    try:
        for name, metric, depth in METRICS:
            plan = dev.plan(metric, depth)
            sums, pq = plan.eval_linear(W)
            for c in range(3):
                exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), name)
                assert np.array_equal(pq[c], exp), (name, c)
                assert int(sums[c]) == fx_sum(exp)
    finally:
        dev.close()


def test_weight_vector_length_truncation(oracle):
    # zip() truncation (dense_dataset.rs:72): shorter and longer weight vectors than D
    X, y, qid, ods, dev = _mk(oracle, 1500, 9, 50, seed=5)
    rng = np.random.default_rng(1)
    try:
        plan = dev.plan(0, 5)
        for wlen in (4, 9, 13):
            W = rng.normal(size=(2, wlen))
            sums, pq = plan.eval_linear(W)
            for c in range(2):
                exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), "ndcg@5")
                assert np.array_equal(pq[c], exp)
    finally:
        dev.close()


@pytest.mark.parametrize("ncand", [1, 2, 5, 26, 51])
def test_coord_sweeps_match_full_rescoring(oracle, ncand):
    X, y, qid, ods, dev = _mk(oracle, 5000, 20, 150, seed=7)
    rng = np.random.default_rng(ncand)
    base = rng.normal(size=(4, 20))
    base[2, 5] = 0.0
    fids = [0, 19, 5, 10]
    cands = []
    for r in range(4):
        orig = base[r, fids[r]]
        c = [0.0] + [orig + s * 0.05 * (2.0 ** k - 1) for k in range(1, 40) for s in (-1, 1)]
        cands.append(c[:ncand])
    try:
        for name, metric, depth in [("ndcg@10", 0, 10), ("map", 1, -1)]:
            plan = dev.plan(metric, depth)
            sums = plan.coord_sweeps(base, fids, cands)
            for r in range(4):
                for k, wv in enumerate(cands[r]):
                    w = base[r].copy()
                    w[fids[r]] = wv
                    exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), name)
                    assert int(sums[r, k]) == fx_sum(exp), (name, r, k)
    finally:
        dev.close()


def test_coord_sweep_feature_beyond_row(oracle):
    # model_dim can exceed D for a Linear model loaded from JSON; the extra weight is inert
    X, y, qid, ods, dev = _mk(oracle, 800, 6, 30, seed=8)
    base = np.random.default_rng(2).normal(size=(1, 9))
    try:
        plan = dev.plan(0, 10)
        sums = plan.coord_sweeps(base, [7], [[0.0, 1.0, -3.0]])
        exp = oracle.evaluate_scores(ods, oracle.score_linear(X, base[0]), "ndcg@10")
        assert [int(v) for v in sums[0]] == [fx_sum(exp)] * 3
    finally:
        dev.close()


def test_query_and_instance_subsets(oracle):
    X, y, qid, ods, dev = _mk(oracle, 3000, 10, 90, seed=11)
    rng = np.random.default_rng(3)
    W = rng.normal(size=(2, 10))
    names = ods.query_names
    pick = [5, 17, 3, 60]
    try:
        # query subset
        plan = dev.plan(0, 5, query_ids=pick)
        sums, pq = plan.eval_linear(W)
        ods.set_view([names[i] for i in pick])
        for c in range(2):
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), "ndcg@5")
            assert np.array_equal(pq[c], exp)
        # instance subset inside those queries
        keep, offs = [], [0]
        for i in pick:
            ids = ods.by_query[names[i]]
            sub = ids[::2]
            keep.extend(sub)
            offs.append(len(keep))
        plan2 = dev.plan(1, -1, query_ids=pick, inst=(offs, keep))
        sums2, pq2 = plan2.eval_linear(W)
        ods.set_view([names[i] for i in pick], instances=keep)
        for c in range(2):
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), "map")
            assert np.array_equal(pq2[c], exp)
    finally:
        ods.set_view(names)
        dev.close()


def test_nan_score_is_reported(oracle):
    X, y, qid, ods, dev = _mk(oracle, 300, 4, 10, seed=12)
    try:
        plan = dev.plan(0, 5)
        W = np.array([[np.nan, 0.0, 0.0, 0.0]])
        with pytest.raises(RuntimeError, match="NaN"):
            plan.eval_linear(W)
    finally:
        dev.close()


def test_fixed_point_sum_is_geometry_independent(oracle):
    # the same queries in a different order (different tiling) give the same integer sum
    X, y, qid, ods, dev = _mk(oracle, 4000, 8, 120, seed=13)
    rng = np.random.default_rng(5)
    W = rng.normal(size=(3, 8))
    try:
        a, _ = dev.plan(0, 10).eval_linear(W, per_query=False)
        perm = rng.permutation(120)
        b, _ = dev.plan(0, 10, query_ids=perm).eval_linear(W, per_query=False)
        assert a.tolist() == b.tolist()
    finally:
        dev.close()


def test_lists_longer_than_the_largest_tile(oracle):
    """Queries of 1025 .. 3000 documents (MSLR-sized lists and beyond) are ranked from HBM by
    long_queries.cu; everything stays bit-identical to the oracle, next to ordinary queries."""
    rng = np.random.default_rng(21)
    lens = [5, 1025, 40, 1500, 1024, 3000, 17]
    qid = np.concatenate([np.full(l, 7 + i) for i, l in enumerate(lens)]).astype(np.int64)
    n = len(qid)
    X = rng.normal(size=(n, 6)).astype(np.float32)
    X[:, 2] = rng.integers(0, 4, n)          # ties
    y = rng.integers(0, 5, n).astype(np.float64)
    ods = oracle_dataset(oracle, X, y, qid)
    qidx, nq = dense_qidx(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    W = rng.normal(size=(19, 6))             # more candidates than one scratch chunk
    W[1, :] = 0.0
    W[1, 2] = 1.0
    try:
        for name, metric, depth in METRICS:
            plan = dev.plan(metric, depth)
            assert dev.lib.fr_dev_plan_has_fast_sweep(plan.ptr) == 0
            sums, pq = plan.eval_linear(W)
            for c in range(W.shape[0]):
                exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), name)
                assert np.array_equal(pq[c], exp), (name, c)
                assert int(sums[c]) == fx_sum(exp)
        plan = dev.plan(0, 10)
        base = rng.normal(size=(2, 6))
        cands = [[0.0, 0.3, -1.0], [base[1, 5], 2.0]]
        sums = plan.coord_sweeps(base, [2, 5], cands)
        for r, f in enumerate([2, 5]):
            for k, wv in enumerate(cands[r]):
                w = base[r].copy()
                w[f] = wv
                assert int(sums[r, k]) == fx_sum(oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10"))
    finally:
        dev.close()


def test_long_lists_through_the_api(oracle):
    import fastrank_b200 as fr

    rng = np.random.default_rng(22)
    lens = [1300, 30, 2100, 64]
    qid = np.concatenate([np.full(l, i) for i, l in enumerate(lens)]).astype(np.int64)
    n = len(qid)
    X = rng.normal(size=(n, 5)).astype(np.float32)
    y = (rng.random(n) < 0.2).astype(np.float64) * rng.integers(1, 4, n)
    ds = fr.CDataset.from_numpy(X, y, qid)
    ods = oracle_dataset(oracle, X, y, qid)
    tree = {"FeatureSplit": {"fid": 1, "split": 0.1, "lhs": {"LeafNode": 1.0},
                             "rhs": {"FeatureSplit": {"fid": 3, "split": -0.2, "lhs": {"LeafNode": 0.5}, "rhs": {"LeafNode": 2.0}}}}}
    for spec in ({"Linear": {"weights": [0.2, -1.0, 0.5, 0.0, 3.0]}}, {"DecisionTree": tree}):
        m = fr.CModel.from_dict(spec)
        for measure in ("ndcg@10", "map", "rr"):
            assert ds.evaluate(m, measure) == oracle.evaluate_model(ods, spec, measure)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@10"
    req.params.num_restarts, req.params.seed, req.params.quiet = 2, 5, True
    model = ds.train_model(req)  # long lists: the exact-order sweep serves the line searches
    w = model.to_dict()["Linear"]["weights"]
    exp = oracle.mean(oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10"))
    assert ds.evaluate_mean(model, "ndcg@10") == pytest.approx(exp, abs=1e-12)


@pytest.mark.parametrize("d,ncand", [(600, 8), (600, 26), (171, 26), (1100, 3), (5, 1), (9, 8)])
def test_wide_matrices_stage_their_weights_in_chunks(oracle, d, ncand):
    """The full-rescore kernel keeps up to 32 KB of candidate weights in shared memory: 600
    features x 8 candidates, 171 x 26 or 1100 x 4 do not fit at once and are staged chunk by
    chunk (barriers inside the tile loop); feature counts that are not a multiple of the 8-value
    register block take the remainder path.  Scores stay the reference's left-to-right sums."""
    rng = np.random.default_rng(d * 31 + ncand)
    n, q = 1500, 40
    qid = np.sort(rng.integers(0, q, n)).astype(np.int64)
    X = rng.normal(size=(n, d)).astype(np.float32)
    X[:, min(2, d - 1)] = rng.integers(0, 3, n)
    y = rng.integers(0, 5, n).astype(np.float64)
    ods = oracle_dataset(oracle, X, y, qid)
    qidx, nq = dense_qidx(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    W = rng.normal(size=(ncand, d))
    try:
        plan = dev.plan(0, 10)
        sums, pq = plan.eval_linear(W)
        for c in range(ncand):
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, W[c]), "ndcg@10")
            assert np.array_equal(pq[c], exp), (d, c)
            assert int(sums[c]) == fx_sum(exp)
        # the exact-order sweep over the same matrix: first, middle and last feature
        base = rng.normal(size=(3, d))
        fids = [0, d // 2, d - 1]
        cands = [[0.0, 0.5, float(base[r, fids[r]])] for r in range(3)]
        got = plan.coord_sweeps(base, fids, cands)
        for r in range(3):
            for k, wv in enumerate(cands[r]):
                w = base[r].copy()
                w[fids[r]] = wv
                assert int(got[r, k]) == fx_sum(oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10"))
    finally:
        dev.close()
