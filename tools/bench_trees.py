#!/usr/bin/env python
"""BASELINE.json configs[3]: score a 500-tree, depth-8 ensemble over the 1M x 136 synthetic
dataset on one B200 (model.rs:64-84, :104-112 on the GPU) and compare with the CPU oracle on a
bounded sample.  Prints one JSON line; not the headline bench (bench.py is)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def random_tree(rng, X, depth, max_depth):
    if depth >= max_depth or (depth > 2 and rng.random() < 0.05):
        return {"LeafNode": float(np.round(rng.uniform(0, 4), 3))}
    fid = int(rng.integers(0, X.shape[1]))
    split = float(np.quantile(X[:2000, fid], rng.uniform(0.1, 0.9)))
    return {"FeatureSplit": {"fid": fid, "split": split,
                             "lhs": random_tree(rng, X, depth + 1, max_depth),
                             "rhs": random_tree(rng, X, depth + 1, max_depth)}}


def count_nodes(t):
    if "LeafNode" in t:
        return 1
    return 1 + count_nodes(t["FeatureSplit"]["lhs"]) + count_nodes(t["FeatureSplit"]["rhs"])


def main():
    import fastrank_b200 as fr
    from oracle import oracle as orc
    from tests.helpers import synth

    n = int(os.environ.get("N_DOCS", 1_000_000))
    trees, depth = int(os.environ.get("N_TREES", 500)), 8
    X, y, qid = synth(n, 136, max(n // 33, 1))
    rng = np.random.default_rng(4)
    members = [{"DecisionTree": random_tree(rng, X, 1, depth)} for _ in range(trees)]
    spec = {"Ensemble": {"weights": [1.0] * trees, "models": members}}
    nodes = sum(count_nodes(m["DecisionTree"]) for m in members)
    ds = fr.CDataset.from_numpy(X, y, qid)
    model = fr.CModel.from_dict(spec)
    model.predict_dense(ds)  # upload + warm-up
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        scores = model.predict_dense(ds)
    t_predict = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        mean = ds.evaluate_mean(model, "ndcg@10")
    t_eval = (time.perf_counter() - t0) / reps
    m = 20000
    t0 = time.perf_counter()
    exp = orc.score_model(X[:m], spec)
    t_cpu = time.perf_counter() - t0
    assert np.array_equal(scores[:m], exp), "tree scores are not bit-identical to the oracle"
    algo_bytes = n * 136 * 4 + n * 8 + nodes * 16
    print(json.dumps({
        "workload": "%d trees depth<=%d (%d nodes) over %d x 136" % (trees, depth, nodes, n),
        "predict_dense_s": t_predict, "docs_per_s": n / t_predict,
        "evaluate_mean_s": t_eval, "ndcg10": mean,
        "algorithmic_bytes": algo_bytes, "achieved_GBps_predict": algo_bytes / t_predict / 1e9,
        "cpu_oracle_docs_per_s_1thread": m / t_cpu, "bit_exact_sample": m,
    }))


if __name__ == "__main__":
    main()
