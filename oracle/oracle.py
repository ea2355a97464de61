"""ctypes front-end of the CPU oracle (oracle/fastrank_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(fastrank_b200/) never imports this module.

Besides the thin ctypes layer this file holds the small pieces of the reference's
*data* handling the checker needs (libsvm densification, query grouping, model JSON ->
oracle bytecode), each citing the reference lines it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import math

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfastrank_oracle.so")

METRICS = {"ndcg": 0, "ap": 1, "map": 1, "rr": 2, "mrr": 2}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the .so path."""
    src = os.path.join(_HERE, "fastrank_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


class _CAParams(C.Structure):
    _fields_ = [
        ("num_restarts", C.c_uint32),
        ("num_max_iterations", C.c_uint32),
        ("step_base", C.c_double),
        ("step_scale", C.c_double),
        ("tolerance", C.c_double),
        ("seed", C.c_uint64),
        ("normalize", C.c_int32),
        ("init_random", C.c_int32),
        ("output_ensemble", C.c_int32),
        ("metric", C.c_int32),
        ("depth", C.c_int64),
    ]


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.fro_score_linear.argtypes = [C.c_size_t, C.c_size_t, vp, vp, C.c_size_t, vp]
        L.fro_score_linear.restype = None
        L.fro_score_model.argtypes = [C.c_size_t, C.c_size_t, vp, vp, vp]
        L.fro_score_model.restype = None
        L.fro_rank_query.argtypes = [C.c_size_t, vp, vp, vp, vp]
        L.fro_rank_query.restype = None
        L.fro_compute_dcg.argtypes = [vp, C.c_size_t, C.c_int64, C.c_int]
        L.fro_compute_dcg.restype = C.c_double
        L.fro_evaluate.argtypes = [C.c_int, C.c_int64, C.c_size_t, vp, vp, vp, vp, vp, vp, vp, vp]
        L.fro_evaluate.restype = C.c_int
        L.fro_mean.argtypes = [vp, C.c_size_t]
        L.fro_mean.restype = C.c_double
        L.fro_coordinate_ascent.argtypes = [C.POINTER(_CAParams), C.c_size_t, C.c_size_t] + [vp] * 2 + [
            C.c_size_t
        ] + [vp] * 6 + [C.c_size_t, vp, vp, vp]
        L.fro_coordinate_ascent.restype = C.c_int64
        L.fro_ca_restart_streams.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, vp]
        L.fro_ca_restart_streams.restype = None
        L.fro_rng_seed.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.fro_rng_set_raw.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        L.fro_rng_u64.argtypes = [vp]
        L.fro_rng_u64.restype = C.c_uint64
        L.fro_rng_float.argtypes = [vp]
        L.fro_rng_float.restype = C.c_double
        L.fro_rng_range.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.fro_rng_range.restype = C.c_uint64
        L.fro_sizeof_rng.restype = C.c_size_t
        L.fro_bootstrap_means.argtypes = [vp, C.c_uint64, C.c_uint32, vp]
        L.fro_bootstrap_means.restype = None
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------------------
# measure names: evaluators.rs:132-155
# ----------------------------------------------------------------------------------------
def parse_measure(name: str) -> Tuple[int, int]:
    depth = -1
    base = name
    if "@" in name:
        base, rhs = name.split("@", 1)
        depth = int(rhs)
        if depth < 0:
            raise ValueError(name)
    base = base.lower()
    if base not in METRICS or (base != "ndcg" and False):
        raise ValueError('Invalid training measure: "%s"' % name)
    return METRICS[base], depth


# ----------------------------------------------------------------------------------------
# data handling restated from the reference
# ----------------------------------------------------------------------------------------
class OracleDataset:
    """Dense view: X f32[N,D] row-major, gains f32[N], qid strings, per-query id lists."""

    def __init__(self, X: np.ndarray, gains: np.ndarray, qids: Sequence[str],
                 docids: Optional[List[Optional[str]]] = None,
                 present: Optional[np.ndarray] = None):
        # present[i, f]: instance i carries feature f (instance.rs:61-77: a Dense32 row knows
        # ids below its own length, a Sparse32 row the listed ones).  None = DenseDataset,
        # where nothing is missing (dense_dataset.rs:139-147).
        self.present = present
        self.X = np.ascontiguousarray(X, dtype=np.float32)
        self.gains = np.ascontiguousarray(gains, dtype=np.float32)
        self.qids = [str(q) for q in qids]
        self.docids = docids
        self.n, self.d = self.X.shape
        # dense_dataset.rs:96-109 / dataset.rs:218-226: ids pushed in ascending order
        order: Dict[str, List[int]] = {}
        for i, q in enumerate(self.qids):
            order.setdefault(q, []).append(i)
        self.query_names = list(order.keys())
        self.by_query = order
        self.set_view(self.query_names)

    def set_view(self, query_names: Sequence[str], instances: Optional[Sequence[int]] = None):
        """Restrict evaluation to some queries (sampling.rs:101-115) or instances."""
        self.view_queries = list(query_names)
        offs = [0]
        docs: List[int] = []
        keep = None if instances is None else set(int(i) for i in instances)
        for q in self.view_queries:
            ids = self.by_query[q]
            if keep is not None:
                ids = [i for i in ids if i in keep]
            docs.extend(ids)
            offs.append(len(docs))
        self.qoff = np.asarray(offs, dtype=np.uint64)
        self.qdocs = np.asarray(docs, dtype=np.uint32)
        return self

    @property
    def nq(self) -> int:
        return len(self.view_queries)


def load_libsvm(path: str) -> OracleDataset:
    """libsvm.rs:131-189 (line format) + instance.rs:104-130 (densification) +
    dataset.rs:211-256 (n_dim = max feature id + 1; absent features read as 0.0)."""
    labels, qids, docids, rows = [], [], [], []
    max_fid = 0
    with open(path, "r") as fp:
        for line in fp:
            if not line.strip():
                continue
            comment = None
            if "#" in line:
                line, comment = line.split("#", 1)
                comment = comment.strip()
            toks = line.split()
            labels.append(np.float32(float(toks[0])))
            rest = toks[1:]
            qid = None
            if rest and rest[0].startswith("qid:"):
                qid = rest[0][4:]
                rest = rest[1:]
            if qid is None:
                raise ValueError("Missing qid")
            feats = {}
            for t in rest:
                k, v = t.split(":", 1)
                k = int(k)
                if k in feats:
                    raise ValueError("MultipleDefinitions")
                feats[k] = np.float32(v)
                max_fid = max(max_fid, k)
            qids.append(qid)
            docids.append(comment)
            rows.append(feats)
    X = np.zeros((len(rows), max_fid + 1), dtype=np.float32)
    present = np.zeros((len(rows), max_fid + 1), dtype=bool)
    for i, feats in enumerate(rows):
        for k, v in feats.items():
            X[i, k] = v
        # instance.rs:104-122: dense when at least half of 1..max own id is listed, and then
        # the row is max_own_id + 1 long (ids below that read 0.0, ids beyond are None)
        own_max = max(feats) if feats else 1
        density = len(feats) / own_max if own_max > 0 else float("inf")
        if density >= 0.5:
            present[i, : own_max + 1] = True
        else:
            present[i, list(feats)] = True
    return OracleDataset(X, np.asarray(labels, dtype=np.float32), qids, docids, present)


def load_qrel(path: str) -> Dict[str, Dict[str, float]]:
    """qrel.rs:65-102."""
    out: Dict[str, Dict[str, float]] = {}
    with open(path, "r") as fp:
        for line in fp:
            row = line.split()
            if not row:
                continue
            out.setdefault(row[0], {})[row[2]] = float(np.float32(row[3]))
    return out


def encode_model(model: dict) -> np.ndarray:
    """Reference model JSON (model.rs:10-16 serde layout) -> oracle bytecode."""
    out: List[float] = []

    def tree(node) -> List[float]:
        if "LeafNode" in node:
            return [3.0, float(node["LeafNode"])]
        fs = node["FeatureSplit"]
        lhs = tree(fs["lhs"])
        rhs = tree(fs["rhs"])
        return [2.0, float(fs["fid"]), float(fs["split"]), float(len(lhs))] + lhs + rhs

    def enc(m) -> List[float]:
        (kind, body), = m.items()
        if kind == "SingleFeature":
            return [0.0, float(body["fid"]), float(body["dir"])]
        if kind == "Linear":
            w = [float(v) for v in body["weights"]]
            return [1.0, float(len(w))] + w
        if kind == "DecisionTree":
            return tree(body)
        if kind == "Ensemble":
            ws = [float(v) for v in body["weights"]]
            code = [4.0, float(len(ws))] + ws
            for sub in body["models"]:
                code += enc(sub)
            return code
        raise ValueError(kind)

    out = enc(model)
    return np.asarray(out, dtype=np.float64)


# ----------------------------------------------------------------------------------------
# oracle calls
# ----------------------------------------------------------------------------------------
def score_linear(X: np.ndarray, w: Sequence[float]) -> np.ndarray:
    X = np.ascontiguousarray(X, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty(X.shape[0], dtype=np.float64)
    lib().fro_score_linear(X.shape[0], X.shape[1], _p(X), _p(w), len(w), _p(out))
    return out


def score_model(X: np.ndarray, model: dict) -> np.ndarray:
    X = np.ascontiguousarray(X, dtype=np.float32)
    code = encode_model(model)
    out = np.empty(X.shape[0], dtype=np.float64)
    lib().fro_score_model(X.shape[0], X.shape[1], _p(X), _p(code), _p(out))
    return out


def rank_query(ids, scores, gains) -> np.ndarray:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    gains = np.ascontiguousarray(gains, dtype=np.float32)
    out = np.empty_like(ids)
    lib().fro_rank_query(len(ids), _p(ids), _p(scores), _p(gains), _p(out))
    return out


def compute_dcg(gains, depth: Optional[int], ideal: bool) -> float:
    g = np.ascontiguousarray(gains, dtype=np.float32)
    return lib().fro_compute_dcg(_p(g), len(g), -1 if depth is None else depth, int(ideal))


def _qrel_arrays(ds: OracleDataset, qrel: Optional[Dict[str, Dict[str, float]]]):
    if qrel is None:
        return None, None, None
    present = np.zeros(ds.nq, dtype=np.uint8)
    offs = [0]
    gains: List[float] = []
    for k, q in enumerate(ds.view_queries):
        if q in qrel:
            present[k] = 1
            gains.extend(qrel[q].values())
        offs.append(len(gains))
    return present, np.asarray(offs, dtype=np.uint64), np.asarray(gains, dtype=np.float32)


def evaluate_scores(ds: OracleDataset, scores: np.ndarray, measure: str,
                    qrel: Optional[Dict[str, Dict[str, float]]] = None) -> np.ndarray:
    """Per-query metric values (view query order) for precomputed scores[instance id]."""
    metric, depth = parse_measure(measure)
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    assert scores.shape[0] == ds.n
    present, qoffs, qgains = _qrel_arrays(ds, qrel)
    out = np.empty(ds.nq, dtype=np.float64)
    status = lib().fro_evaluate(metric, depth, ds.nq, _p(ds.qoff), _p(ds.qdocs), _p(scores),
                                _p(ds.gains), _p(present), _p(qoffs), _p(qgains), _p(out))
    if status != 0:
        raise RuntimeError("reference would panic: actual DCG > ideal DCG")
    return out


def evaluate_model(ds: OracleDataset, model: dict, measure: str, qrel=None) -> Dict[str, float]:
    """evaluate_by_query restated (ffi.rs:246-258 -> evaluators.rs:186-204)."""
    vals = evaluate_scores(ds, score_model(ds.X, model), measure, qrel)
    return dict(zip(ds.view_queries, vals.tolist()))


def mean(values: np.ndarray) -> float:
    v = np.ascontiguousarray(values, dtype=np.float64)
    return lib().fro_mean(_p(v), len(v))


def coordinate_ascent(ds: OracleDataset, measure: str, *, num_restarts=5, num_max_iterations=25,
                      step_base=0.05, step_scale=2.0, tolerance=0.001, seed=42, normalize=True,
                      init_random=True, features: Optional[Sequence[int]] = None, qrel=None):
    """coordinate_ascent.rs:197-253.  Returns dict(weights, score, all_weights, all_scores,
    best_restart, n_evals)."""
    metric, depth = parse_measure(measure)
    fids = np.asarray(list(range(ds.d)) if features is None else list(features), dtype=np.uint32)
    dim = int(fids.max()) + 1
    params = _CAParams(num_restarts, num_max_iterations, step_base, step_scale, tolerance, seed,
                       int(normalize), int(init_random), 0, metric, depth)
    out_w = np.zeros((num_restarts, dim), dtype=np.float64)
    out_s = np.zeros(num_restarts, dtype=np.float64)
    n_evals = C.c_uint64(0)
    present, qoffs, qgains = _qrel_arrays(ds, qrel)
    best = lib().fro_coordinate_ascent(C.byref(params), ds.n, ds.d, _p(ds.X), _p(ds.gains), ds.nq,
                                       _p(ds.qoff), _p(ds.qdocs), _p(present), _p(qoffs),
                                       _p(qgains), _p(fids), len(fids), _p(out_w), _p(out_s),
                                       C.cast(C.byref(n_evals), C.c_void_p))
    if best < 0:
        raise RuntimeError("empty dataset")
    return {
        "weights": out_w[best].copy(),
        "score": float(out_s[best]),
        "all_weights": out_w,
        "all_scores": out_s,
        "best_restart": int(best),
        "n_evals": int(n_evals.value),
    }


def bootstrap_means(values: np.ndarray, trials: int = 200) -> np.ndarray:
    """evaluators.rs:157-171: resampled means in trial order."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = np.empty(trials, dtype=np.float64)
    lib().fro_bootstrap_means(_p(v), len(v), trials, _p(out))
    return out


def percentile(sorted_values: np.ndarray, p: float) -> float:
    """stats.rs:142-156 (PercentileStats::percentile), including its interpolation weights."""
    n = p * (len(sorted_values) - 1)
    lhs = int(n)
    rhs = min(len(sorted_values), int(math.ceil(n)))
    interp = n - int(n)
    if lhs == rhs:
        return float(sorted_values[lhs])
    return float(interp * sorted_values[lhs] + (1.0 - interp) * sorted_values[rhs])


class Rng:
    """oorandom::Rand64 (=11.1.0) restatement; pinned by the goldens listed in the C header."""

    def __init__(self, seed: int = 0):
        self._buf = C.create_string_buffer(lib().fro_sizeof_rng() + 16)
        addr = C.addressof(self._buf)
        self._ptr = C.c_void_p((addr + 15) & ~15)
        lib().fro_rng_seed(self._ptr, seed & ((1 << 64) - 1), seed >> 64)

    def set_raw(self, state: int, inc: int):
        m = (1 << 64) - 1
        lib().fro_rng_set_raw(self._ptr, state & m, state >> 64, inc & m, inc >> 64)

    def u64(self) -> int:
        return lib().fro_rng_u64(self._ptr)

    def float(self) -> float:
        return lib().fro_rng_float(self._ptr)

    def range(self, start: int, end: int) -> int:
        return lib().fro_rng_range(self._ptr, start, end)
