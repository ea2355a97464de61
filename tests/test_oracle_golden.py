"""Pins the CPU oracle to every known-answer value the reference's own tests hold for the
score -> rank -> metric path.  CPU only."""
import json
import os

import numpy as np
import pytest


def test_single_feature_ndcg5_goldens(oracle, trec_train, goldens, golden_dir):
    # reference tests/test_with_example_data.py:139-167: CA on a one-feature subsample with
    # (1 restart, 1 iteration, step 1.0, no normalise, uniform init) leaves w_f = 1.0.
    names = json.load(open(os.path.join(golden_dir, "trec_news_2018.features.json")))
    name_to_fid = {v: int(k) for k, v in names.items()}
    name_to_fid["0"] = 0
    for name, expected in goldens["single_feature_ndcg5"]["values"].items():
        fid = name_to_fid[name]
        w = np.zeros(trec_train.d)
        w[fid] = 1.0
        per_query = oracle.evaluate_scores(trec_train, oracle.score_linear(trec_train.X, w), "ndcg@5")
        assert per_query.shape == (45,)
        got = float(np.mean(per_query))
        assert abs(got - expected) < 1e-12, (name, got, expected)
        # and through the oracle's CA driver, as the reference test does it
        res = oracle.coordinate_ascent(trec_train, "ndcg@5", num_restarts=1, num_max_iterations=1,
                                       step_base=1.0, normalize=False, init_random=False,
                                       seed=42, features=[fid])
        assert res["weights"][fid] == 1.0
        assert np.count_nonzero(res["weights"]) == 1
        assert abs(res["score"] - expected) < 1e-12


def test_rank_ties_known_answer(oracle, goldens):
    g = goldens["rank_ties"]
    rows = g["input"]
    ids = oracle.rank_query([r[2] for r in rows], [r[0] for r in rows], [r[1] for r in rows])
    assert ids.tolist() == g["expected_ids"]


def test_compute_ndcg_known_answer(oracle, goldens):
    g = goldens["compute_ndcg"]
    ideal = oracle.compute_dcg(g["gains"], None, True)
    actual = oracle.compute_dcg(g["gains"], None, False)
    assert abs(actual / ideal - g["expected"]) <= g["tolerance"]


def test_dcg_depth_pads_and_truncates(oracle):
    # evaluators.rs:262-264: resize(depth) truncates or zero-pads
    g = [3.0, 2.0, 1.0]
    full = oracle.compute_dcg(g, None, False)
    assert oracle.compute_dcg(g, 10, False) == full
    assert oracle.compute_dcg(g, 2, False) == pytest.approx(7.0 / 1.0 + 3.0 / np.log2(3.0), abs=0)
    assert oracle.compute_dcg([], 5, False) == 0.0
    assert oracle.compute_dcg(g[::-1], None, True) == full


def test_dataset_shape(trec_train, goldens):
    g = goldens["dataset_shape"]
    assert (trec_train.n, trec_train.d, trec_train.nq) == (g["n"], g["d"], g["n_queries"])
    assert np.all(trec_train.X[:, 0] == 0.0)


def test_qrel_vs_dataset_norms_agree_on_fixture(oracle, trec_train, golden_dir):
    # reference tests/test_with_example_data.py:243-251
    qrel = oracle.load_qrel(os.path.join(golden_dir, "newsir18-entity.qrel"))
    assert len(qrel) == 50
    rng = np.random.default_rng(5)
    w = rng.normal(size=trec_train.d)
    s = oracle.score_linear(trec_train.X, w)
    a = oracle.evaluate_scores(trec_train, s, "ndcg@5", qrel)
    b = oracle.evaluate_scores(trec_train, s, "ndcg@5")
    assert abs(a.mean() - b.mean()) < 1e-7
    for m in ("map", "mrr", "ndcg"):
        assert oracle.evaluate_scores(trec_train, s, m, qrel).shape == (45,)


def test_metrics_against_independent_numpy(oracle, trec_train):
    # independent numpy restatement of evaluators.rs (ranking key: -score, gain, id)
    rng = np.random.default_rng(11)
    w = rng.normal(size=trec_train.d)
    w[3] = 0.0
    s = oracle.score_linear(trec_train.X, w)
    s_np = np.zeros(trec_train.n)
    for j in range(trec_train.d):
        s_np = s_np + trec_train.X[:, j].astype(np.float64) * w[j]
    assert np.array_equal(s, s_np)
    got = {m: oracle.evaluate_scores(trec_train, s, m) for m in ("ndcg@5", "ndcg", "map", "rr")}
    for k, q in enumerate(trec_train.view_queries):
        ids = np.asarray(trec_train.by_query[q])
        g = trec_train.gains[ids]
        order = np.lexsort((ids, g, -s[ids]))
        rg = g[order].astype(np.float64)
        rel = rg > 0
        disc = np.log2(np.arange(len(rg)) + 2.0)

        def dcg(v, depth):
            v = list(v)
            if depth is not None:
                v = (v + [0.0] * depth)[:depth]
            t = 0.0
            for i, x in enumerate(v):
                t += (2.0 ** x - 1.0) / np.log2(i + 2.0)
            return t

        ideal = sorted(g.astype(np.float64), reverse=True)
        for name, depth in (("ndcg@5", 5), ("ndcg", None)):
            exp = dcg(rg, depth) / dcg(ideal, depth) if rel.any() else 0.0
            assert got[name][k] == pytest.approx(exp, abs=1e-15)
        if rel.any():
            ranks = np.nonzero(rel)[0] + 1
            ap = sum((i + 1) / r for i, r in enumerate(ranks)) / rel.sum()
            assert got["map"][k] == pytest.approx(ap, abs=1e-15)
            assert got["rr"][k] == 1.0 / ranks[0]
        else:
            assert got["map"][k] == 0.0 and got["rr"][k] == 0.0
        del disc


def test_model_bytecode_scoring(oracle):
    X = np.array([[1.0, 5.0, -2.0], [0.5, 6.25, 3.0], [9.0, 7.0, 0.0]], dtype=np.float32)
    tree = {"FeatureSplit": {"fid": 1, "split": 6.25,
                             "lhs": {"LeafNode": 7.0},
                             "rhs": {"FeatureSplit": {"fid": 0, "split": 2.0,
                                                      "lhs": {"LeafNode": -1.0},
                                                      "rhs": {"LeafNode": 12.0}}}}}
    assert oracle.score_model(X, {"DecisionTree": tree}).tolist() == [7.0, 7.0, 12.0]
    ens = {"Ensemble": {"weights": [0.5, 2.0],
                        "models": [{"DecisionTree": tree}, {"Linear": {"weights": [1.0, 0.0, 1.0, 99.0]}}]}}
    assert oracle.score_model(X, ens).tolist() == [0.5 * 7 + 2 * -1.0, 0.5 * 7 + 2 * 3.5, 0.5 * 12 + 2 * 9.0]
    assert oracle.score_model(X, {"SingleFeature": {"fid": 2, "dir": -1.0}}).tolist() == [2.0, -3.0, -0.0]
    assert oracle.score_model(X, {"SingleFeature": {"fid": 7, "dir": 3.0}}).tolist() == [0.0, 0.0, 0.0]


def test_regression_tree_known_answer(oracle, goldens):
    # random_forest.rs:465-506: the tree the reference learns reproduces ys exactly; scoring the
    # tree that SURVEY 8c derived (splits 6.25 then 3.03125) must return the labels.
    g = goldens["regression_tree"]
    X = np.asarray(g["xs"], dtype=np.float32).reshape(-1, 1)
    tree = {"FeatureSplit": {"fid": 0, "split": 6.25,
                             "lhs": {"FeatureSplit": {"fid": 0, "split": 3.03125,
                                                      "lhs": {"LeafNode": 7.0}, "rhs": {"LeafNode": 2.0}}},
                             "rhs": {"LeafNode": 12.0}}}
    assert oracle.score_model(X, {"DecisionTree": tree}).tolist() == [float(y) for y in g["ys"]]


def test_oorandom_known_answers(oracle, goldens):
    # oorandom =11.1.0 (Cargo.toml:18-19): first draws for seed 42 and the learners' default seed
    # Rand64::new(0xdeadbeef).rand_u64() (coordinate_ascent.rs:27-34, random_forest.rs:143-146)
    g = goldens["oorandom_known_answers"]
    rng = oracle.Rng(42)
    assert [rng.u64() for _ in range(3)] == g["seed_42_first_three_u64"]
    assert oracle.Rng(0xDEADBEEF).u64() == g["default_seed"]


def test_lcg_step_matches_numpy_pcg64(oracle):
    # The 128-bit LCG underneath (multiplier, increment convention) is the one numpy's PCG64
    # uses; only the output function differs (oorandom: rotr64(((s >> 29) ^ s) >> 58, s >> 122)).
    bg = np.random.PCG64(1234)
    st = bg.state["state"]
    state, inc = st["state"], st["inc"]
    rng = oracle.Rng(0)
    rng.set_raw(state, inc)
    mult = 47026247687942121848144207491837523525
    mask128, mask64 = (1 << 128) - 1, (1 << 64) - 1
    for _ in range(16):
        x = (((state >> 29) ^ state) >> 58) & mask64
        rot = state >> 122
        expect = ((x >> rot) | (x << ((64 - rot) & 63))) & mask64
        assert rng.u64() == expect
        state = (state * mult + inc) & mask128
    # ... and numpy's next raw draw comes from the same successor state (XSL-RR of it)
    nxt = (st["state"] * mult + inc) & mask128
    xsl = ((nxt >> 64) ^ nxt) & mask64
    r = nxt >> 122
    assert int(bg.random_raw(1)[0]) == ((xsl >> r) | (xsl << ((64 - r) & 63))) & mask64


def test_rng_range_drops_the_range_start(oracle):
    # 11.1.0's Rand64::rand_range returns a draw over [0, end - start): the start is not added
    # (randutil.rs:24 therefore shuffles with rand_range(i..n) in [0, n - i)).
    rng = oracle.Rng(42)
    seen = set()
    for _ in range(400):
        v = rng.range(3, 10)
        assert 0 <= v < 7
        seen.add(v)
        f = rng.float()
        assert 0.0 <= f < 1.0
    assert seen == set(range(7))


def test_random_forest_determinism_golden(oracle, trec_train, goldens):
    # random_forest.rs:427-463 / tests/test_with_example_data.py:175-201: 10 trees, seed 42 ->
    # NDCG@5 0.4367914517387043.  Pins RNG stream + sampling + induction (FeatureStats skip rows
    # that do not carry the feature, normalizers.rs:21-27) + ensemble scoring + NDCG.
    from oracle import random_forest_oracle as rfo

    g = goldens["rf_determinism_ndcg5"]
    assert trec_train.present is not None and not trec_train.present.all()
    model = rfo.learn_forest(trec_train, g["params"])
    got = oracle.mean(oracle.evaluate_scores(trec_train, oracle.score_model(trec_train.X, model), "ndcg@5"))
    assert abs(got - g["expected"]) < g["tolerance"], got


def test_notebook_coordinate_ascent_golden(oracle, trec_train, trec_test, goldens):
    # examples/FastRankDemo.ipynb cells 4-5: the reference's own output for seed 1234567 with the
    # default parameters -- pins reset / shuffle / the whole line-search driver end to end.
    g = goldens["notebook_coordinate_ascent"]
    res = oracle.coordinate_ascent(trec_train, g["request"]["measure"], seed=g["request"]["seed"])
    assert np.allclose(res["weights"], g["weights"], rtol=0, atol=g["weights_tolerance"])
    test_ndcg = oracle.mean(oracle.evaluate_scores(trec_test, oracle.score_linear(trec_test.X, res["weights"]), "ndcg@5"))
    assert "%.3g" % test_ndcg == g["test_ndcg5_printed"]


def test_notebook_random_forest_golden(oracle, trec_train, trec_test, goldens):
    # examples/FastRankDemo.ipynb cells 3, 5: 100 trees, both sampling rates 0.5, seed 1234567
    from oracle import random_forest_oracle as rfo

    g = goldens["notebook_random_forest"]
    params = {"min_leaf_support": 10, "max_depth": 8, "split_candidates": 3, "split_method": "SquaredError"}
    params.update(g["params"])
    model = rfo.learn_forest(trec_train, params)
    test_ndcg = oracle.mean(oracle.evaluate_scores(trec_test, oracle.score_model(trec_test.X, model), "ndcg@5"))
    assert "%.3g" % test_ndcg == g["test_ndcg5_printed"]


def test_ca_improves_and_is_deterministic(oracle, trec_train):
    a = oracle.coordinate_ascent(trec_train, "ndcg@5", num_restarts=2, seed=42)
    b = oracle.coordinate_ascent(trec_train, "ndcg@5", num_restarts=2, seed=42)
    assert np.array_equal(a["all_weights"], b["all_weights"])
    assert a["score"] >= 0.43
    per_query = oracle.evaluate_scores(trec_train, oracle.score_linear(trec_train.X, a["weights"]), "ndcg@5")
    assert oracle.mean(per_query) == a["score"]


def test_oracle_forest_trainer_reproduces_the_regression_tree_known_answer(oracle, goldens):
    # random_forest.rs:465-506 with the sort-based restatement (oracle/random_forest_oracle.py):
    # the learned tree must fit ys exactly, through the splits SURVEY 8c derived (6.25, 3.03125)
    from oracle import random_forest_oracle as rfo

    g = goldens["regression_tree"]
    X = np.asarray(g["xs"], dtype=np.float32).reshape(-1, 1)
    ys = np.asarray(g["ys"], dtype=np.float32)
    ds = oracle.OracleDataset(X, ys, ["query"] * len(ys))
    m = rfo.learn_forest(ds, {"seed": 42, "num_trees": 1, "split_method": "SquaredError", "min_leaf_support": 1,
                              "max_depth": 10, "split_candidates": 32, "feature_sampling_rate": 0.25,
                              "instance_sampling_rate": 0.5})
    tree = m["Ensemble"]["models"][0]["DecisionTree"]
    assert tree["FeatureSplit"]["split"] == 6.25
    assert tree["FeatureSplit"]["lhs"]["FeatureSplit"]["split"] == 3.03125
    assert oracle.score_model(X, m).tolist() == [float(v) for v in ys]


def test_bootstrap_means_against_a_plain_python_restatement(oracle):
    # evaluators.rs:157-171 with the Python Rng wrapper drawing the same stream
    values = np.random.default_rng(4).random(37)
    got = oracle.bootstrap_means(values, 9)
    rng = oracle.Rng(0xDEADBEEF)
    for t in range(9):
        total = 0.0
        for _ in range(len(values)):
            total += values[rng.range(0, len(values))]
        assert got[t] == total / len(values)
    s = np.sort(np.arange(10.0))
    assert oracle.percentile(s, 0.5) == 4.5 and oracle.percentile(np.arange(9.0), 0.5) == 4.0  # stats.rs:190-197
