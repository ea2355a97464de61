#!/usr/bin/env python
"""train_model(random_forest) on the 1M x 136 synthetic dataset: per-level statistics on the GPU
(rf_induction.cu) against the host trainer (FASTRANK_RF=host).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fastrank_b200 as fr
    from tests.helpers import synth

    n = int(os.environ.get("N_DOCS", 1_000_000))
    trees = int(os.environ.get("N_TREES", 32))
    X, y, qid = synth(n, 136, max(n * 3 // 100, 1))
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = fr.TrainRequest.random_forest()
    req.measure = "ndcg@10"
    req.params.num_trees, req.params.quiet, req.params.seed = trees, True, 5
    out = {"workload": "random_forest %d trees (defaults: depth 8, 3 split candidates, 25%% features, 50%% queries) on %d x 136" % (trees, n)}
    ds.evaluate_mean(fr.CModel.from_dict({"Linear": {"weights": [1.0]}}), "ndcg@10")  # upload before timing
    for where in ("gpu", "host"):
        os.environ["FASTRANK_RF"] = where
        t = time.perf_counter()
        model = ds.train_model(req)
        out["seconds_" + where] = time.perf_counter() - t
        out["ndcg10_" + where] = ds.evaluate_mean(model, "ndcg@10")
    out["speedup"] = out["seconds_host"] / out["seconds_gpu"]
    out["host_threads"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
