"""Python surface of fastrank_b200 -- same classes, methods, argument meaning and error
behaviour as the reference's fastrank/clib.py (CQRel :62-123, CModel :126-206,
CDataset :209-491, query_json :494-505), bound to libfastrank_b200.so instead of the Rust
cdylib.  Everything that scores, ranks or evaluates runs on the GPU behind the C ABI.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Set

import numpy as np

from ._native import ffi, lib

# model.rs:10-16 (ModelEnum variants)
_MODEL_KINDS = ("SingleFeature", "Linear", "DecisionTree", "Ensemble")


def _raise_if_error(payload) -> None:
    """{"error":..,"context":..} objects are how the native side reports failures
    (ffi.rs:22-26); the reference turns them into a plain Exception("error: context")."""
    if isinstance(payload, dict) and "error" in payload and "context" in payload:
        raise Exception("{0}: {1}".format(payload["error"], payload["context"]))


def _take_str(raw) -> str:
    """Copies a native string into Python and releases it with free_str."""
    if raw == ffi.NULL:
        raise ValueError("native call returned NULL")
    try:
        return ffi.string(ffi.cast("char*", raw)).decode("utf-8")
    finally:
        lib.free_str(ffi.cast("void*", raw))


def _take_json(raw):
    value = json.loads(_take_str(raw))
    _raise_if_error(value)
    return value


def _take_result(res):
    """Unpacks a CResult: frees the message and the struct, returns the success pointer or
    raises (clib.py:22-39 semantics)."""
    if res == ffi.NULL:
        raise ValueError("CResult should not be NULL")
    message = None
    if res.error_message != ffi.NULL:
        message = _take_str(res.error_message)
    success = res.success if res.success != ffi.NULL else None
    lib.free_c_result(ffi.cast("CResult*", res))
    if message is not None:
        if "{" in message:
            _raise_if_error(json.loads(message))
        else:
            raise Exception(message)
    return success


def _check_fast_path(raw) -> None:
    """The binary fast paths return NULL on success, an error JSON string otherwise."""
    if raw != ffi.NULL:
        _take_json(raw)


class _Handle:
    """Owns one native pointer and releases it exactly once."""

    _free = None
    _what = "object"

    def __init__(self, pointer=None):
        self.pointer = pointer

    def __del__(self):
        ptr, self.pointer = getattr(self, "pointer", None), None
        if ptr is not None and type(self)._free is not None:
            type(self)._free(ptr)

    def _require_init(self):
        if self.pointer is None:
            raise ValueError("{0} is null!".format(self._what))


class CQRel(_Handle):
    """A loaded set of TREC relevance judgments; some measures need the number of relevant
    documents (MAP) or the ideal gain (NDCG) of the whole judged pool."""

    _what = "CQRel"

    def __init__(self, pointer=None):
        super().__init__(pointer)
        self._queries = None

    @staticmethod
    def load_file(path: str) -> "CQRel":
        return CQRel(ffi.cast("CQRel*", _take_result(lib.load_cqrel(path.encode("utf-8")))))

    @staticmethod
    def from_dict(dictionaries: Dict[str, Dict[str, float]]) -> "CQRel":
        text = json.dumps(dictionaries).encode("utf-8")
        return CQRel(ffi.cast("CQRel*", _take_result(lib.cqrel_from_json(text))))

    def _query_json(self, message="queries"):
        self._require_init()
        return _take_json(lib.cqrel_query_json(self.pointer, message.encode("utf-8")))

    def to_dict(self) -> Dict[str, Dict[str, float]]:
        return self._query_json("to_json")

    def queries(self) -> Set[str]:
        if self._queries is None:
            self._queries = set(self._query_json("queries"))
        return self._queries

    def query_judgments(self, qid: str) -> Dict[str, float]:
        if qid in self.queries():
            return self._query_json(qid)
        raise ValueError("No qid={0} in cqrel: {1}".format(qid, self.queries()))


CQRel._free = staticmethod(lambda p: lib.free_cqrel(p))


class CModel(_Handle):
    """A trained or deserialised model (Linear / SingleFeature / DecisionTree / Ensemble)."""

    _what = "CModel"

    def __init__(self, pointer, params=None):
        super().__init__(pointer)
        self.params = params

    @staticmethod
    def _check_model_json(model_json: Dict):
        [kind] = list(model_json.keys())
        assert kind in _MODEL_KINDS

    @staticmethod
    def from_dict(model_json: Dict) -> "CModel":
        CModel._check_model_json(model_json)
        text = json.dumps(model_json).encode("utf-8")
        return CModel(ffi.cast("CModel*", _take_result(lib.model_from_json(text))))

    def predict_scores(self, dataset: "CDataset") -> Dict[int, float]:
        """instance index -> score for every instance of the dataset (json_api.rs:53-72)."""
        self._require_init()
        dataset._require_init()
        response = _take_json(lib.predict_scores(self.pointer, dataset.pointer))
        return dict((int(k), v) for k, v in response.items())

    def predict_dense_scores(self, dataset: "CDataset", missing: float = float("nan")) -> List[float]:
        """Scores as a list aligned with instance ids; `missing` where the dataset (a sample)
        does not hold the id."""
        output: List[float] = []
        for index, score in sorted(self.predict_scores(dataset).items()):
            while len(output) < index:
                output.append(missing)
            if index == len(output):
                output.append(score)
            else:
                output[index] = score
        return output

    def predict_dense(self, dataset: "CDataset") -> np.ndarray:
        """Binary fast path (no JSON): float64 scores indexed by instance id of the parent
        dataset, NaN for instances a sample does not hold."""
        self._require_init()
        dataset._require_init()
        n = dataset._parent_instances()
        out = np.empty(n, dtype=np.float64)
        _check_fast_path(lib.predict_dense_f64(self.pointer, dataset.pointer,
                                               ffi.cast("double*", out.ctypes.data), n))
        return out

    def _query_json(self, message="to_json"):
        self._require_init()
        return _take_json(lib.model_query_json(ffi.cast("void*", self.pointer), message.encode("utf-8")))

    def to_dict(self):
        return self._query_json("to_json")

    def __str__(self):
        return str(self.to_dict())


CModel._free = staticmethod(lambda p: lib.free_model(p))


class CDataset(_Handle):
    """A dataset owned by the native side: open_ranksvm() for libsvm/ranklib files,
    from_numpy() for in-memory arrays.  The feature matrix moves to GPU memory on first use."""

    _what = "CDataset"

    def __init__(self, pointer=None):
        super().__init__(pointer)
        # from_numpy borrows these buffers (lib.rs:232-234); samples share them
        self.numpy_arrays_to_keep = []
        self._n_parent = None

    def __del__(self):
        super().__del__()
        self.numpy_arrays_to_keep = []

    def _require_init(self):
        if self.pointer is None:
            raise ValueError("Forgot to call open_* or from_numpy on CDataset!")

    @staticmethod
    def open_ranksvm(data_path, feature_names_path=None) -> "CDataset":
        names = ffi.NULL if feature_names_path is None else ffi.new("char[]", feature_names_path.encode("utf-8"))
        path = ffi.new("char[]", data_path.encode("utf-8"))
        return CDataset(ffi.cast("CDataset*", _take_result(lib.load_ranksvm_format(path, names))))

    @staticmethod
    def from_numpy(X, y, qid) -> "CDataset":
        """X float32 (N x D, C-contiguous), y float64 (N), qid int64 (N).  The arrays are
        borrowed, not copied, exactly as in the reference (clib.py:255-297)."""
        (N, D) = X.shape
        assert N > 0
        assert D > 0
        assert len(y) == N
        assert len(qid) == N
        assert X.dtype == "float32"
        assert y.dtype == "float64"
        assert qid.dtype == "int64"
        X = np.ascontiguousarray(X)
        y = np.ascontiguousarray(y).reshape(-1)
        qid = np.ascontiguousarray(qid).reshape(-1)
        dataset = CDataset(
            ffi.cast(
                "CDataset*",
                _take_result(
                    lib.make_dense_dataset_f32_f64_i64(
                        N, D,
                        ffi.cast("float *", X.ctypes.data),
                        ffi.cast("double *", y.ctypes.data),
                        ffi.cast("int64_t *", qid.ctypes.data),
                    )
                ),
            )
        )
        dataset.numpy_arrays_to_keep = [X, y, qid]
        dataset._n_parent = N
        return dataset

    def _child(self, pointer) -> "CDataset":
        child = CDataset(ffi.cast("CDataset*", pointer))
        child.numpy_arrays_to_keep = self.numpy_arrays_to_keep
        child._n_parent = self._parent_instances()
        return child

    def _parent_instances(self) -> int:
        if self._n_parent is None:
            if self.is_sampled():
                ids = [i for ids in self.instances_by_query().values() for i in ids]
                self._n_parent = (max(ids) + 1) if ids else 0
            else:
                self._n_parent = self.num_instances()
        return self._n_parent

    def subsample_queries(self, queries: List[str]) -> "CDataset":
        self._require_init()
        actual = self.queries()
        for q in queries:
            if q not in actual:
                raise ValueError(
                    "Asked for query that does not exist in subsample: {0} not in {1}".format(q, actual)
                )
        request = json.dumps(queries).encode("utf-8")
        return self._child(_take_result(lib.dataset_query_sampling(self.pointer, request)))

    def subsample_feature_names(self, features: List[str]) -> "CDataset":
        name_to_id = self.feature_name_to_index()
        fnums = sorted(set(name_to_id[f] for f in features))
        request = json.dumps(fnums).encode("utf-8")
        return self._child(_take_result(lib.dataset_feature_sampling(self.pointer, request)))

    def train_model(self, train_req: "TrainRequest", sweep: Optional[str] = None) -> CModel:  # noqa: F821
        """The reference's train_model.  `sweep` (extension, coordinate ascent only): "exact" runs
        the line searches on the exact-order kernel -- scores summed in the reference's feature
        order, bit for bit -- instead of the batched sweep, whose scores add the varied coordinate's
        term last and can differ in the last bits; "batched" / None is the default."""
        self._require_init()
        req_dict = train_req.to_dict()
        if sweep is not None:
            req_dict["sweep"] = sweep
        request = ffi.new("char[]", json.dumps(req_dict).encode("utf-8"))
        pointer = _take_result(lib.train_model(request, ffi.cast("void*", self.pointer)))
        return CModel(ffi.cast("CModel*", pointer), train_req)

    def _query_json(self, message="num_features"):
        self._require_init()
        cmd = ffi.new("char[]", message.encode("utf-8"))
        return _take_json(lib.dataset_query_json(ffi.cast("void*", self.pointer), cmd))

    def is_sampled(self) -> bool:
        return self._query_json("is_sampled")

    def num_features(self) -> int:
        return self._query_json("num_features")

    def feature_ids(self) -> Set[int]:
        return set(self._query_json("feature_ids"))

    def feature_names(self) -> Set[str]:
        return set(self._query_json("feature_names"))

    def feature_index_to_name(self) -> Dict[int, str]:
        return dict(zip(self._query_json("feature_ids"), self._query_json("feature_names")))

    def feature_name_to_index(self) -> Dict[str, int]:
        return dict(zip(self._query_json("feature_names"), self._query_json("feature_ids")))

    def num_instances(self) -> int:
        return self._query_json("num_instances")

    def queries(self) -> Set[str]:
        return set(self._query_json("queries"))

    def instances_by_query(self) -> Dict[str, List[int]]:
        return self._query_json("instances_by_query")

    def evaluate(self, model: CModel, evaluator: str, qrel: Optional[CQRel] = None) -> Dict[str, float]:
        """query id -> metric value ("ndcg", "ndcg@5", "map", "mrr", ...)."""
        self._require_init()
        model._require_init()
        qrel_pointer = ffi.NULL
        if qrel is not None:
            qrel._require_init()
            qrel_pointer = qrel.pointer
        return _take_json(
            lib.evaluate_by_query(model.pointer, self.pointer, qrel_pointer, evaluator.encode("utf-8"))
        )

    def evaluate_mean(self, model: CModel, evaluator: str, qrel: Optional[CQRel] = None) -> float:
        """Binary fast path: the mean over queries (evaluators.rs:173-184) without JSON."""
        self._require_init()
        model._require_init()
        qrel_pointer = ffi.NULL if qrel is None else qrel.pointer
        out = ffi.new("double*")
        _check_fast_path(
            lib.evaluate_mean_f64(model.pointer, self.pointer, qrel_pointer, evaluator.encode("utf-8"), out)
        )
        return out[0]

    def predict_scores(self, model: CModel) -> Dict[int, float]:
        return model.predict_scores(self)

    def bootstrap_eval(self, model: CModel, evaluator: str, qrel: Optional[CQRel] = None,
                       num_trials: int = 200) -> Dict[str, float]:
        """SetEvaluator::bootstrap_eval + PercentileStats::summary (evaluators.rs:157-171,
        stats.rs:139-165): the per-query values of `model` are resampled with replacement
        `num_trials` times (Rand64::new(0xdeadbeef), as the reference does); returns the mean and
        the 5/25/50/75/95th percentiles of the resampled means, plus the sorted means."""
        self._require_init()
        model._require_init()
        qrel_pointer = ffi.NULL if qrel is None else qrel.pointer
        means = np.empty(num_trials, dtype=np.float64)
        _check_fast_path(lib.evaluate_bootstrap_f64(model.pointer, self.pointer, qrel_pointer,
                                                   evaluator.encode("utf-8"), num_trials,
                                                   ffi.cast("double*", means.ctypes.data)))

        def percentile(p: float) -> float:  # stats.rs:142-156, weights as the reference has them
            n = p * (len(means) - 1)
            lhs = int(n)
            rhs = min(len(means), int(np.ceil(n)))
            interp = n - int(n)
            if lhs == rhs:
                return float(means[lhs])
            return float(interp * means[lhs] + (1.0 - interp) * means[rhs])

        out = {"mean": self.evaluate_mean(model, evaluator, qrel), "means": means}
        for name, p in (("p5", 0.05), ("p25", 0.25), ("p50", 0.5), ("p75", 0.75), ("p95", 0.95)):
            out[name] = percentile(p)
        return out

    def device_profile(self, enable: Optional[bool] = None, read: bool = False):
        """Per-kernel device timing of this dataset's scoring / ranking launches (CUDA events on
        the library's stream).  enable=True/False switches it; read=True returns
        (launches, total_ms) since the last read and resets the record."""
        self._require_init()
        n = ffi.new("uint64_t*") if read else ffi.NULL
        ms = ffi.new("double*") if read else ffi.NULL
        _check_fast_path(lib.dataset_device_profile(self.pointer, -1 if enable is None else int(bool(enable)), n, ms))
        return (int(n[0]), float(ms[0])) if read else None

    def predict_trecrun(self, model: CModel, output_path: str, system_name: str = "fastrank",
                        quiet=True, depth=0) -> int:
        self._require_init()
        model._require_init()
        response = _take_json(
            lib.predict_to_trecrun(model.pointer, self.pointer, output_path.encode("utf-8"),
                                   system_name.encode("utf-8"), depth)
        )
        if not quiet:
            print("Wrote {} records to {} as {}.".format(response, output_path, system_name))
        return response


CDataset._free = staticmethod(lambda p: lib.free_dataset(p))


def query_json(message: str):
    """Sends a command string to the native query_json entry point and decodes the answer."""
    return _take_json(lib.query_json(message.encode("utf-8")))
