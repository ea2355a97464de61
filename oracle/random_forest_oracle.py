"""CPU restatement of the reference's random-forest learner -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/random_forest.rs line by line, including the sort-by-feature the
product's host trainer (fastrank_b200/csrc/random_forest.cpp) replaces with counting passes, so
that the two can be compared tree for tree.  numpy for the sorts, plain Python loops for the
sequential f64 sums (small cases only).  RNG: oracle.Rng (oorandom =11.1.0 restated, see
fastrank_oracle.c).  Pinned end to end by the reference's determinism golden
(random_forest.rs:427-463 -> 0.4367914517387043; tests/test_oracle_golden.py).

Where the reference leaves an order unspecified (sort_unstable among equal keys, HashMap
iteration) this file makes the same deterministic choice as the product: stable sorts, the last
maximal element, queries in order of first appearance.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import oracle as orc


def _shuffle(v: list, rng) -> None:  # randutil.rs:21-27
    n = len(v)
    for i in range(n):
        j = rng.range(i, n)
        v[i], v[j] = v[j], v[i]


def _sample_without_replacement(data: Sequence, rng, count: int) -> list:  # randutil.rs:14-18
    v = list(data)
    _shuffle(v, rng)
    return v[:count]


class _Stats:  # stats.rs:53-124
    def __init__(self):
        self.n, self.mean, self.s = 0, 0.0, 0.0
        self.max, self.min = -1.7976931348623157e308, 1.7976931348623157e308

    def push(self, x: float):
        self.n += 1
        old_mean, old_s = self.mean, self.s
        if self.max < x:
            self.max = x
        if self.min > x:
            self.min = x
        if self.n == 1:
            self.mean = x
            return
        self.mean = old_mean + (x - old_mean) / self.n
        self.s = old_s + (x - old_mean) * (x - self.mean)

    def finished(self) -> bool:
        return self.n > 1

    def variance(self) -> float:
        return self.s / (self.n - 1)


def _compute_output(ids, gains) -> float:  # random_forest.rs:32-41
    if len(ids) == 0:
        return 0.0
    total = 0.0
    for i in ids:
        total += float(gains[i])
    return total / len(ids)


def _squared_error(ids, gains) -> float:  # :42-51
    out = _compute_output(ids, gains)
    sse = 0.0
    for i in ids:
        diff = out - float(gains[i])
        sse += diff * diff
    return sse


def _plogp(x: float) -> float:
    return 0.0 if x == 0.0 else x * math.log2(x)


def _binary(ids, gains, entropy: bool) -> float:  # gini :52-66, entropy :74-87
    if len(ids) == 0:
        return 0.0
    count = float(len(ids))
    positive = sum(1 for i in ids if gains[i] > 0.0)
    p_yes = positive / count
    p_no = (count - positive) / count
    if entropy:
        return -_plogp(p_yes) - _plogp(p_no)
    return p_yes * (1.0 - p_yes) + p_no * (1.0 - p_no)


def _variance(ids, gains) -> float:
    st = _Stats()
    for i in ids:
        st.push(float(gains[i]))
    return st.variance()


def _importance(method: str, lhs, rhs, gains) -> float:  # :90-125
    if method == "SquaredError":
        return -(_squared_error(lhs, gains) + _squared_error(rhs, gains))
    if method == "BinaryGiniImpurity":
        return -(_binary(lhs, gains, False) * len(lhs) + _binary(rhs, gains, False) * len(rhs))
    if method == "InformationGain":
        return -(_binary(lhs, gains, True) * len(lhs) + _binary(rhs, gains, True) * len(rhs))
    return -(_variance(lhs, gains) * len(lhs) + _variance(rhs, gains) * len(rhs))


def _split_candidate(params, fid, X, gains, instances, fstats):  # :211-286
    labels = _Stats()
    for i in instances:
        labels.push(float(gains[i]))
    if not labels.finished() or labels.max == labels.min:
        return None
    k = params["split_candidates"]
    rng_ = fstats.max - fstats.min
    values = np.asarray([float(X[i, fid]) for i in instances], dtype=np.float64)
    order = np.argsort(values, kind="stable")
    scores = values[order]
    ids = [instances[j] for j in order]
    positions = []
    at = 0
    for i in range(1, k):
        position = (i / k) * rng_ + fstats.min
        while at < len(ids) and scores[at] < position:
            at += 1
        if positions and positions[-1][1] == at:
            continue
        positions.append((position, at))
    best = None
    for position, right in positions:
        lhs, rhs = ids[:right], ids[right:]
        if len(lhs) < params["min_leaf_support"] or len(rhs) < params["min_leaf_support"]:
            continue
        imp = _importance(params["split_method"], lhs, rhs, gains)
        if best is None or imp >= best[0]:  # sort by importance, take the last
            best = (imp, position, right)
    if best is None:
        return None
    imp, position, right = best
    return {"fid": fid, "split": position, "importance": imp, "lhs": ids[:right], "rhs": ids[right:]}


def _learn_recursive(params, X, gains, features, instances, depth, trace=None, path="", present=None):  # :362-408
    if not features or not instances:
        return None
    if depth >= params["max_depth"]:
        return None
    if len(instances) < params["min_leaf_support"]:
        return None
    best = None
    for fid in features:
        st = _Stats()  # FeatureStats::compute, normalizers.rs:13-36: missing values are skipped
        for i in instances:
            if present is None or present[i, fid]:
                st.push(float(X[i, fid]))
        if not st.finished():
            continue
        cand = _split_candidate(params, fid, X, gains, instances, st)
        if cand is None:
            continue
        if trace is not None:
            trace.setdefault(path, []).append((int(fid), float(cand["split"]), float(cand["importance"])))
        if best is None or cand["importance"] >= best["importance"]:
            best = cand
    if best is None:
        return None
    lhs = _learn_recursive(params, X, gains, features, best["lhs"], depth + 1, trace, path + "L", present)
    if lhs is None:
        lhs = {"LeafNode": _compute_output(best["lhs"], gains)}
    rhs = _learn_recursive(params, X, gains, features, best["rhs"], depth + 1, trace, path + "R", present)
    if rhs is None:
        rhs = {"LeafNode": _compute_output(best["rhs"], gains)}
    return {"FeatureSplit": {"fid": int(best["fid"]), "split": float(best["split"]), "lhs": lhs, "rhs": rhs}}


def learn_forest(ds: "orc.OracleDataset", params: Dict, traces: Optional[List[Dict]] = None) -> Dict:
    """random_forest.rs:288-342 without the per-tree evaluation (weights 1.0, i.e.
    weight_trees = false).  `params` uses the RandomForestParams field names.  When `traces`
    is a list, one dict per tree is appended: node path ("", "L", "LR", ...) -> the best
    (fid, split, importance) of every feature considered at that node."""
    X, gains = ds.X, ds.gains
    master = orc.Rng(int(params["seed"]))
    seeds = [master.u64() for _ in range(params["num_trees"])]
    features = sorted(range(ds.d))
    queries = sorted(ds.query_names)
    models: List[Dict] = []
    for seed in seeds:
        rng = orc.Rng(seed)
        n_features = max(1, int(len(features) * params["feature_sampling_rate"]))  # sampling.rs:46-47
        n_queries = max(1, int(len(queries) * params["instance_sampling_rate"]))
        fsel = _sample_without_replacement(features, rng, n_features)
        qsel = set(_sample_without_replacement(queries, rng, n_queries))
        instances: List[int] = []
        for q in ds.query_names:  # order of first appearance
            if q in qsel:
                instances.extend(int(i) for i in ds.by_query[q])
        trace = {} if traces is not None else None
        root = _learn_recursive(params, X, gains, fsel, instances, 1, trace, "", ds.present)
        if traces is not None:
            traces.append(trace)
        if root is None:
            root = {"LeafNode": _compute_output(instances, gains)}
        models.append({"DecisionTree": root})
    return {"Ensemble": {"weights": [1.0] * len(models), "models": models}}
