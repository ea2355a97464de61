"""Parity of the batched coordinate-ascent sweep (fr_dev_eval_coord_sweeps_fast) against the
CPU oracle and against the exact-order kernel.

Contract under test (include/fastrank_b200.h): candidate scores are one rounding away from the
reference's left-to-right dot product, everything after the score (ranking, tie-break, metric
terms, summation order) is the reference's.  So:
  * whenever the arithmetic is exact (dyadic weights on small-integer features, all-zero bases)
    results must be BIT-IDENTICAL to the oracle, ties included;
  * on generic float data per-query values are bit-identical wherever the ranking is, i.e.
    everywhere except for documents whose scores agree to the last bits; the mean must be
    within 1e-9 of the oracle (north_star tolerance: 1e-5).
"""
import numpy as np
import pytest

from tests.helpers import DevDataset, dense_qidx, fx_sum, oracle_dataset, synth

pytestmark = pytest.mark.gpu

FX = float(1 << 40)


def _mk(oracle, n, d, q, seed, shuffle=False, X=None, y=None, qid=None):
    if X is None:
        X, y, qid = synth(n, d, q, seed=seed, shuffle_rows=shuffle)
    ods = oracle_dataset(oracle, X, y, qid)
    qidx, nq = dense_qidx(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    return X, y, qid, ods, dev


def _line(orig, n):
    c = [0.0] + [orig + s * 0.05 * (2.0 ** k - 1) for k in range(1, 40) for s in (-1, 1)]
    return c[:n]


def _check(oracle, ods, X, plan, name, base, fids, cands, exact=False, max_flip_frac=2e-4):
    sums, pq = plan.coord_sweeps(base, fids, cands, fast=True, per_query=True)
    nq = pq.shape[2]
    total = flips = 0
    for r in range(len(fids)):
        for k, wv in enumerate(cands[r]):
            w = base[r].copy()
            if fids[r] < len(w):
                w[fids[r]] = wv
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), name)
            got = pq[r, k]
            bad = int((got != exp).sum())
            total += nq
            flips += bad
            assert int(sums[r, k]) == fx_sum(got), "sum is not the sum of the per-query values"
            if exact:
                assert bad == 0, (name, r, k, np.abs(got - exp).max())
                assert int(sums[r, k]) == fx_sum(exp)
            else:
                assert abs(got.mean() - exp.mean()) < 1e-9, (name, r, k)
    assert flips <= max_flip_frac * total, (name, flips, total)
    return flips, total


@pytest.mark.parametrize("n_sweeps,ncand", [(1, 1), (3, 5), (8, 26), (8, 25), (11, 32), (2, 51)])
def test_fast_sweeps_match_oracle(oracle, n_sweeps, ncand):
    X, y, qid, ods, dev = _mk(oracle, 5000, 20, 150, seed=7)
    rng = np.random.default_rng(100 * n_sweeps + ncand)
    base = rng.normal(size=(n_sweeps, 20))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int(v) for v in rng.integers(0, 20, n_sweeps)]
    fids[0] = 0
    fids[-1] = 19
    if n_sweeps > 2:
        fids[1] = fids[2]  # two sweeps on the same coordinate
    cands = [_line(base[r, fids[r]], ncand) for r in range(n_sweeps)]
    try:
        for name, metric, depth in [("ndcg@10", 0, 10), ("ndcg", 0, -1), ("map", 1, -1), ("rr", 2, -1)]:
            plan = dev.plan(metric, depth)
            assert dev.lib.fr_dev_plan_has_fast_sweep(plan.ptr) == 1
            _check(oracle, ods, X, plan, name, base, fids, cands)
    finally:
        dev.close()


def test_fast_sweeps_equal_exact_kernel_on_larger_data(oracle):
    X, y, qid, ods, dev = _mk(oracle, 60000, 48, 1800, seed=21)
    rng = np.random.default_rng(4)
    base = rng.uniform(-1, 1, size=(8, 48))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int(v) for v in rng.integers(0, 48, 8)]
    cands = [_line(base[r, fids[r]], 26) for r in range(8)]
    try:
        plan = dev.plan(0, 10)
        exact = plan.coord_sweeps(base, fids, cands)
        fast = plan.coord_sweeps(base, fids, cands, fast=True)
        diff = np.abs(exact - fast).astype(np.float64) / FX / 1800
        assert diff.max() < 1e-9, diff.max()
        assert (exact == fast).mean() > 0.95
    finally:
        dev.close()


def test_exact_arithmetic_cases_are_bit_identical(oracle):
    """Small-integer features with dyadic weights: every product and sum is exact, so the fast
    path must reproduce the oracle bit for bit -- including massive score ties, where only the
    (gain asc, id asc) tie-break decides the ranking."""
    rng = np.random.default_rng(5)
    n, d, q = 4000, 12, 100
    qid = np.sort(rng.integers(0, q, n)).astype(np.int64)
    X = rng.integers(0, 4, size=(n, d)).astype(np.float32)
    y = rng.integers(0, 5, n).astype(np.float64)
    X, y, qid, ods, dev = _mk(oracle, n, d, q, 0, X=X, y=y, qid=qid)
    base = rng.integers(-4, 5, size=(8, d)) / 8.0
    base[0, :] = 0.0  # all scores tie for candidate 0.0
    base[1, :] = 0.0
    fids = [3, 0, 11, 5, 5, 7, 1, 2]
    cands = [[0.0, 0.5, -0.25, 1.0, 2.0, -8.0, 0.125, 16.0] for _ in range(8)]
    try:
        for name, metric, depth in [("ndcg@10", 0, 10), ("ndcg@3", 0, 3), ("ndcg", 0, -1), ("map", 1, -1), ("rr", 2, -1)]:
            plan = dev.plan(metric, depth)
            _check(oracle, ods, X, plan, name, base, fids, cands, exact=True)
    finally:
        dev.close()


def test_ragged_queries_and_length_limit(oracle):
    rng = np.random.default_rng(9)
    lens = [1, 2, 3, 31, 32, 33, 64, 127, 128, 129, 255, 256, 5, 1, 40]
    qid = np.concatenate([np.full(l, 100 + i) for i, l in enumerate(lens)]).astype(np.int64)
    n = len(qid)
    X = rng.integers(0, 3, size=(n, 7)).astype(np.float32)
    y = rng.integers(0, 5, n).astype(np.float64)
    y[qid == 102] = 0.0  # a query without relevant documents
    X, y, qid, ods, dev = _mk(oracle, n, 7, len(lens), 0, X=X, y=y, qid=qid)
    base = rng.integers(-4, 5, size=(3, 7)) / 4.0
    fids = [0, 3, 6]
    cands = [[0.0, 0.5, -1.0, 2.0, 0.25]] * 3
    try:
        for name, metric, depth in [("ndcg@10", 0, 10), ("ndcg", 0, -1), ("map", 1, -1), ("rr", 2, -1)]:
            plan = dev.plan(metric, depth)
            _check(oracle, ods, X, plan, name, base, fids, cands, exact=True)
    finally:
        dev.close()
    # when more than a quarter of the documents sit in lists beyond the largest sweep tile the
    # plan stays on the exact-order kernels, and the batched entry point must refuse, loudly
    qid2 = np.concatenate([qid, np.full(513, 999)]).astype(np.int64)
    X2 = np.concatenate([X, rng.normal(size=(513, 7)).astype(np.float32)])
    y2 = np.concatenate([y, rng.integers(0, 3, 513).astype(np.float64)])
    qidx, nq = dense_qidx(qid2)
    dev2 = DevDataset(X2, y2.astype(np.float32), qidx, nq)
    try:
        plan = dev2.plan(0, 10)
        assert dev2.lib.fr_dev_plan_has_fast_sweep(plan.ptr) == 0
        with pytest.raises(RuntimeError, match="not available"):
            plan.coord_sweeps(base, fids, cands, fast=True)
        plan.coord_sweeps(base, fids, cands)  # the exact kernel still serves it
    finally:
        dev2.close()


def test_negative_fractional_and_many_distinct_gains(oracle):
    rng = np.random.default_rng(13)
    n, d, q = 3000, 6, 80
    qid = np.sort(rng.integers(0, q, n)).astype(np.int64)
    X = rng.integers(0, 5, size=(n, d)).astype(np.float32)
    base = rng.integers(-4, 5, size=(2, d)) / 4.0
    cands = [[0.0, 1.0, -0.5], [0.25, 2.0, -2.0]]
    # (a) negative and fractional gains: documents below AND above the zero-gain block contribute
    y = rng.choice([-1.0, -0.5, 0.0, 0.0, 0.5, 1.0, 2.5], size=n)
    _, _, _, ods, dev = _mk(oracle, n, d, q, 0, X=X, y=y, qid=qid)
    try:
        for name, metric, depth in [("ndcg@10", 0, 10), ("ndcg", 0, -1), ("map", 1, -1), ("rr", 2, -1)]:
            sums, pq = dev.plan(metric, depth).coord_sweeps(base, [1, 4], cands, fast=True, per_query=True)
            for r, f in enumerate([1, 4]):
                for k, wv in enumerate(cands[r]):
                    w = base[r].copy()
                    w[f] = wv
                    try:
                        exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), name)
                    except Exception:
                        continue
                    assert np.array_equal(pq[r, k], exp), (name, r, k)
    finally:
        dev.close()
    # (b) more than 255 distinct gain values: no discount table, terms are divided on the device
    y = np.round(rng.random(n) * 3.0, 3) * (rng.random(n) < 0.5)
    _, _, _, ods, dev = _mk(oracle, n, d, q, 0, X=X, y=y, qid=qid)
    try:
        assert len(np.unique(y.astype(np.float32))) > 255
        plan = dev.plan(0, 10)
        _check(oracle, ods, X, plan, "ndcg@10", base, [1, 4], cands, exact=True)
    finally:
        dev.close()


def test_subsets_truncation_and_inert_coordinates(oracle):
    X, y, qid, ods, dev = _mk(oracle, 3000, 10, 90, seed=11)
    rng = np.random.default_rng(3)
    names = ods.query_names
    pick = [5, 17, 3, 60, 61, 62]
    try:
        # query subset + instance subset
        keep, offs = [], [0]
        for i in pick:
            ids = ods.by_query[names[i]]
            keep.extend(ids[::2])
            offs.append(len(keep))
        plan = dev.plan(0, 5, query_ids=pick, inst=(offs, keep))
        ods.set_view([names[i] for i in pick], instances=keep)
        base = rng.normal(size=(2, 10))
        _check(oracle, ods, X, plan, "ndcg@5", base, [2, 9], [_line(base[0, 2], 7), _line(base[1, 9], 7)],
               max_flip_frac=0.0)
        ods.set_view(names)
        # weight vector shorter than the row (zip truncation) and a coordinate beyond the row
        plan = dev.plan(0, 10)
        short = rng.normal(size=(2, 6))
        _check(oracle, ods, X, plan, "ndcg@10", short, [1, 5], [_line(short[0, 1], 4), _line(short[1, 5], 4)],
               max_flip_frac=0.0)
        wide = rng.normal(size=(1, 13))
        sums, pq = plan.coord_sweeps(wide, [12], [[0.0, 1.0, -3.0]], fast=True, per_query=True)
        exp = oracle.evaluate_scores(ods, oracle.score_linear(X, wide[0]), "ndcg@10")
        for k in range(3):
            assert np.array_equal(pq[0, k], exp)
        # a sweep without candidates next to one with
        sums = plan.coord_sweeps(short, [1, 5], [[], [0.5]], fast=True)
        w = short[1].copy()
        w[5] = 0.5
        assert int(sums[1, 0]) == fx_sum(oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10"))
    finally:
        ods.set_view(names)
        dev.close()


def test_fast_sweep_reports_nan(oracle):
    X, y, qid, ods, dev = _mk(oracle, 300, 4, 10, seed=12)
    try:
        plan = dev.plan(0, 5)
        with pytest.raises(RuntimeError, match="NaN"):
            plan.coord_sweeps(np.array([[np.nan, 0.0, 0.0, 0.0]]), [1], [[0.0, 1.0]], fast=True)
    finally:
        dev.close()


def test_more_rows_than_one_pass_holds(oracle):
    """8 sweeps x 80 candidates = 640 rows: a sweep group takes 512 rows per pass over X, so the
    call needs a second pass; 3 sweeps x 200 candidates exercises an uneven split."""
    X, y, qid, ods, dev = _mk(oracle, 3000, 10, 90, seed=17)
    rng = np.random.default_rng(6)
    try:
        plan = dev.plan(0, 10)
        for n_sweeps, ncand in ((8, 80), (3, 200)):
            base = rng.normal(size=(n_sweeps, 10))
            fids = [int(v) for v in rng.integers(0, 10, n_sweeps)]
            cands = [[float(v) for v in rng.normal(size=ncand)] for _ in range(n_sweeps)]
            fast = plan.coord_sweeps(base, fids, cands, fast=True)
            exact = plan.coord_sweeps(base, fids, cands)
            assert np.abs(fast - exact).max() / FX / 90 < 1e-9
            for r, k in ((0, 0), (n_sweeps - 1, ncand - 1), (1, ncand // 2)):
                w = base[r].copy()
                w[fids[r]] = cands[r][k]
                exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10")
                assert abs(int(fast[r, k]) - fx_sum(exp)) / FX / 90 < 1e-9
    finally:
        dev.close()


def _ragged(rng, lens, d=6, integer=False):
    qid = np.concatenate([np.full(l, 3 + 2 * i) for i, l in enumerate(lens)]).astype(np.int64)
    n = len(qid)
    if integer:
        X = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    else:
        X = rng.normal(size=(n, d)).astype(np.float32)
        X[:, 2] = rng.integers(0, 4, n)      # ties
    y = (rng.integers(0, 5, n) * (rng.random(n) < 0.5)).astype(np.float64)
    return X, y, qid


@pytest.mark.parametrize("integer", [True, False])
def test_tiles_of_up_to_512_documents(oracle, integer):
    """Lists of 257 .. 512 documents stay on the batched sweep (512-document tiles, one CTA per
    SM); with exact arithmetic the results are bit-identical to the oracle."""
    rng = np.random.default_rng(81)
    lens = [500, 30, 300, 12, 512, 64, 257, 5, 130, 511, 1, 400]
    X, y, qid = _ragged(rng, lens, integer=integer)
    _, _, _, ods, dev = _mk(oracle, 0, 0, 0, 0, X=X, y=y, qid=qid)
    try:
        for name, metric, depth in (("ndcg@10", 0, 10), ("map", 1, -1), ("mrr", 2, -1), ("ndcg", 0, -1)):
            plan = dev.plan(metric, depth)
            assert dev.lib.fr_dev_plan_has_fast_sweep(plan.ptr) == 1
            if integer:
                base = rng.integers(-4, 5, size=(3, 6)).astype(np.float64) / 8.0
                cands = [[float(v) / 4.0 for v in rng.integers(-8, 9, 40)] for _ in range(3)]
            else:
                base = rng.normal(size=(3, 6))
                cands = [_line(base[r, f], 40) for r, f in enumerate([0, 2, 5])]
            _check(oracle, ods, X, plan, name, base, [0, 2, 5], cands, exact=integer, max_flip_frac=0.02)
    finally:
        dev.close()


def test_long_lists_next_to_the_batched_sweep(oracle):
    """A few lists longer than the largest sweep tile (600 .. 2000 documents, under a quarter of
    the documents) must not push the dataset off the batched sweep: they are ranked from HBM,
    everything else in tiles, one call returns both."""
    rng = np.random.default_rng(82)
    lens = [int(v) for v in rng.integers(1, 120, 260)]
    for at, l in ((7, 600), (100, 1100), (259, 2000)):
        lens.insert(at, l)
    X, y, qid = _ragged(rng, lens, integer=True)
    assert sum(l for l in lens if l > 512) * 4 <= len(qid)
    _, _, _, ods, dev = _mk(oracle, 0, 0, 0, 0, X=X, y=y, qid=qid)
    try:
        for name, metric, depth in (("ndcg@10", 0, 10), ("map", 1, -1), ("mrr", 2, -1)):
            plan = dev.plan(metric, depth)
            assert dev.lib.fr_dev_plan_has_fast_sweep(plan.ptr) == 1
            base = rng.integers(-4, 5, size=(4, 6)).astype(np.float64) / 8.0
            cands = [[float(v) / 4.0 for v in rng.integers(-8, 9, 70)] for _ in range(4)]  # > one scratch chunk
            _check(oracle, ods, X, plan, name, base, [1, 2, 3, 9], cands, exact=True)
            # the exact-order entry point and evaluate_mean see the same split
            sums = plan.coord_sweeps(base[:2], [1, 2], [cands[0][:5], cands[1][:5]])
            fast = plan.coord_sweeps(base[:2], [1, 2], [cands[0][:5], cands[1][:5]], fast=True)
            assert np.array_equal(sums, fast)
    finally:
        dev.close()


def test_tile_cap_knob_and_training_with_untiled_lists(oracle, monkeypatch):
    """FASTRANK_TILE_CAP moves the tiled / untiled boundary: a model trained with most lists
    untiled (cap 32) is the model trained with everything tiled."""
    import fastrank_b200 as fr

    rng = np.random.default_rng(83)
    lens = [int(v) for v in rng.integers(1, 90, 60)]
    X, y, qid = _ragged(rng, lens, integer=True)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@5"
    req.params.num_restarts, req.params.seed, req.params.quiet = 2, 5, True
    got = {}
    for cap in (None, "32"):
        if cap is None:
            monkeypatch.delenv("FASTRANK_TILE_CAP", raising=False)
        else:
            monkeypatch.setenv("FASTRANK_TILE_CAP", cap)
        ds = fr.CDataset.from_numpy(X, y, qid)
        m = ds.train_model(req)
        got[cap] = (m.to_dict()["Linear"]["weights"], fr.query_json("last_train_stats")["evals_consumed"],
                    ds.evaluate_mean(m, "ndcg@5"))
    assert got[None] == got["32"]
    ods = oracle_dataset(oracle, X, y, qid)
    res = oracle.coordinate_ascent(ods, "ndcg@5", num_restarts=2, seed=5)
    assert got["32"][1] == res["n_evals"]
    assert np.allclose(got["32"][0], res["weights"], rtol=0, atol=1e-12)


def test_untiled_lists_in_several_scratch_passes(oracle, monkeypatch):
    """The scratch arrays of the untiled lists hold a bounded number of candidates
    (FASTRANK_LONG_CHUNK forces 5): both entry points walk the candidate rows in several passes,
    across sweep-group boundaries, and still return what one pass returns."""
    rng = np.random.default_rng(84)
    lens = [int(v) for v in rng.integers(1, 100, 120)] + [700]
    X, y, qid = _ragged(rng, lens, integer=True)
    base = rng.integers(-4, 5, size=(10, 6)).astype(np.float64) / 8.0   # 10 sweeps = two sweep groups
    fids = [int(v) for v in rng.integers(0, 6, 10)]
    cands = [[float(v) / 4.0 for v in rng.integers(-8, 9, int(rng.integers(1, 9)))] for _ in range(10)]
    got = {}
    for chunk in (None, "5"):
        if chunk is None:
            monkeypatch.delenv("FASTRANK_LONG_CHUNK", raising=False)
        else:
            monkeypatch.setenv("FASTRANK_LONG_CHUNK", chunk)
        _, _, _, ods, dev = _mk(oracle, 0, 0, 0, 0, X=X, y=y, qid=qid)
        try:
            plan = dev.plan(0, 10)
            assert dev.lib.fr_dev_plan_untiled_queries(plan.ptr) == 1
            assert dev.lib.fr_dev_plan_has_fast_sweep(plan.ptr) == 1
            fast, pq = plan.coord_sweeps(base, fids, cands, fast=True, per_query=True)
            exact = plan.coord_sweeps(base, fids, cands)
            assert np.array_equal(fast, exact)
            got[chunk] = (fast.copy(), pq.copy())
            if chunk is not None:
                _check(oracle, ods, X, plan, "ndcg@10", base, fids, cands, exact=True)
        finally:
            dev.close()
    assert np.array_equal(got[None][0], got["5"][0])
    assert np.array_equal(got[None][1], got["5"][1])


@pytest.mark.parametrize("depth", [1, 5, 10, 16])
def test_packed_kernel_equals_general_kernel_bit_for_bit(oracle, monkeypatch, depth):
    """NDCG@k with k <= 16 is served by sweep_packed_kernel (top-k of a candidate in one 64-bit
    register); FASTRANK_SWEEP_KERNEL=tile forces the general tile kernel.  Same scores, same
    comparisons, same fold order: sums and per-query values must agree bit for bit on float data
    (ragged lists up to 512 documents, negative gains, queries without relevant documents, 8 x 51
    candidates so that row groups straddle sweeps, and a plan with few tiles so that quarter items
    are handed out)."""
    rng = np.random.default_rng(90 + depth)
    lens = [int(v) for v in rng.integers(1, 70, 400)] + [130, 255, 1, 2, 33]
    if depth == 10:
        lens += [300, 512]
    X, y, qid = _ragged(rng, lens, d=20)
    y[rng.random(len(y)) < 0.05] = -1.0
    y[qid == 3] = 0.0
    _, _, _, ods, dev = _mk(oracle, 0, 0, 0, 0, X=X, y=y, qid=qid)
    base = rng.normal(size=(8, 20))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int(v) for v in rng.integers(0, 20, 8)]
    cands = [_line(base[r, fids[r]], 51) for r in range(8)]
    got = {}
    try:
        plan = dev.plan(0, depth)
        # FASTRANK_PRUNE_MIN: lists of at least that many documents take the packed kernel's
        # select-then-rank path (default 48; 0 = never; 2 = nearly every list, which also drives
        # lists with more than 32 survivors per candidate into the fall-back to the full count)
        for which, prune in (("tile", None), (None, None), (None, "0"), (None, "2")):
            if which is None:
                monkeypatch.delenv("FASTRANK_SWEEP_KERNEL", raising=False)
            else:
                monkeypatch.setenv("FASTRANK_SWEEP_KERNEL", which)
            if prune is None:
                monkeypatch.delenv("FASTRANK_PRUNE_MIN", raising=False)
            else:
                monkeypatch.setenv("FASTRANK_PRUNE_MIN", prune)
            got[(which, prune)] = plan.coord_sweeps(base, fids, cands, fast=True, per_query=True)
        monkeypatch.delenv("FASTRANK_PRUNE_MIN", raising=False)
        for key in ((None, None), (None, "0"), (None, "2")):
            assert np.array_equal(got[("tile", None)][0], got[key][0]), key
            assert np.array_equal(got[("tile", None)][1], got[key][1]), key
        got = {"tile": got[("tile", None)], None: got[(None, None)]}
        # and against the oracle on a few candidates
        name = "ndcg@%d" % depth
        for r, k in ((0, 0), (3, 17), (7, 50)):
            w = base[r].copy()
            w[fids[r]] = cands[r][k]
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), name)
            assert np.abs(got[None][1][r, k] - exp).max() < 1e-9 or (got[None][1][r, k] != exp).mean() < 0.01
            assert abs(got[None][1][r, k].mean() - exp.mean()) < 1e-9
    finally:
        dev.close()


@pytest.mark.parametrize("name,metric,depth", [("ndcg", 0, -1), ("ndcg@20", 0, 20), ("map", 1, -1), ("rr", 2, -1), ("ndcg@10", 0, 10)])
def test_slot_mode_of_the_packed_kernel_equals_general_kernel(oracle, monkeypatch, name, metric, depth):
    """Measures the register-packed top-k cannot hold (NDCG without cut-off or k > 16, AP, RR) run the
    same warp-item kernel with ranks filed in a per-warp slot buffer (tiles of <= 256 documents).
    FASTRANK_SWEEP_KERNEL=tile forces the general tile kernel, =slots forces slot mode where the
    register mode would apply: same bits everywhere, and the oracle's on exact arithmetic."""
    rng = np.random.default_rng(190 + metric + depth)
    lens = [int(v) for v in rng.integers(1, 70, 300)] + [130, 255, 1, 2, 33, 256]
    X, y, qid = _ragged(rng, lens, d=16)
    y[rng.random(len(y)) < 0.05] = -1.0
    y[qid == 5] = 0.0
    _, _, _, ods, dev = _mk(oracle, 0, 0, 0, 0, X=X, y=y, qid=qid)
    base = rng.normal(size=(8, 16))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int(v) for v in rng.integers(0, 16, 8)]
    cands = [_line(base[r, fids[r]], 51) for r in range(8)]
    got = {}
    try:
        plan = dev.plan(metric, depth)
        assert dev.lib.fr_dev_plan_tile_documents(plan.ptr) == 256
        for which in ("tile", "slots", None):
            if which is None:
                monkeypatch.delenv("FASTRANK_SWEEP_KERNEL", raising=False)
            else:
                monkeypatch.setenv("FASTRANK_SWEEP_KERNEL", which)
            kernel = dev.ffi.string(dev.lib.fr_dev_plan_sweep_kernel(plan.ptr)).decode()
            got[which] = plan.coord_sweeps(base, fids, cands, fast=True, per_query=True) + (kernel,)
        assert got["tile"][2].startswith("sweep_fast_kernel") and got["slots"][2].endswith("slots>")
        assert got[None][2].endswith("slots>") == (name != "ndcg@10")
        for which in ("slots", None):
            assert np.array_equal(got["tile"][0], got[which][0]), which
            assert np.array_equal(got["tile"][1], got[which][1]), which
        for r, k in ((0, 0), (5, 33)):
            w = base[r].copy()
            w[fids[r]] = cands[r][k]
            exp = oracle.evaluate_scores(ods, oracle.score_linear(X, w), name)
            assert abs(got[None][1][r, k].mean() - exp.mean()) < 1e-9
    finally:
        dev.close()


@pytest.mark.parametrize("name,metric", [("mrr", 2), ("map", 1), ("ndcg", 0)])
def test_slot_mode_with_many_tied_scores_is_bit_identical_to_the_oracle(oracle, monkeypatch, name, metric):
    """Small-integer features and dyadic weights: scores are exact and tie all the time, so the tie
    rule decides most ranks -- in particular which relevant document is "the first" for RR, whose
    slot-mode path finds the best relevant document and counts once instead of ranking them all.
    Tiles of <= 256 documents (slot mode), compared with the general kernel and with the oracle."""
    rng = np.random.default_rng(300 + metric)
    lens = [int(v) for v in rng.integers(1, 60, 250)] + [200, 256, 1, 2]
    X, y, qid = _ragged(rng, lens, d=6, integer=True)
    y[qid == 3] = 0.0          # a list without relevant documents
    y[qid == 5] = 2.0          # and one with nothing else
    _, _, _, ods, dev = _mk(oracle, 0, 0, 0, 0, X=X, y=y, qid=qid)
    base = rng.integers(-4, 5, size=(3, 6)).astype(np.float64) / 8.0
    cands = [[float(v) / 4.0 for v in rng.integers(-8, 9, 40)] for _ in range(3)]
    try:
        plan = dev.plan(metric, -1)
        assert dev.lib.fr_dev_plan_tile_documents(plan.ptr) == 256
        monkeypatch.delenv("FASTRANK_SWEEP_KERNEL", raising=False)
        assert dev.ffi.string(dev.lib.fr_dev_plan_sweep_kernel(plan.ptr)).decode().endswith("slots>")
        got = plan.coord_sweeps(base, [0, 2, 5], cands, fast=True, per_query=True)
        monkeypatch.setenv("FASTRANK_SWEEP_KERNEL", "tile")
        ref = plan.coord_sweeps(base, [0, 2, 5], cands, fast=True, per_query=True)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        monkeypatch.delenv("FASTRANK_SWEEP_KERNEL", raising=False)
        _check(oracle, ods, X, plan, name, base, [0, 2, 5], cands, exact=True)
    finally:
        dev.close()
