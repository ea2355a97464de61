"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/fastrank_b200.h declares, and the host logic that needs no GPU (JSON, libsvm and
qrel parsing, dataset views, model JSON, error envelopes) behaves like the reference.
No compute calls are made here."""
import ctypes
import json
import os

import pytest


@pytest.fixture(scope="module")
def fr():
    import fastrank_b200

    return fastrank_b200


def test_every_declared_symbol_is_exported():
    from fastrank_b200 import _native

    names = _native.exported_symbols()
    assert len(names) >= 40
    reference_abi = """free_str free_c_result free_dataset free_model free_cqrel load_cqrel cqrel_from_json
        cqrel_query_json load_ranksvm_format dataset_query_sampling dataset_feature_sampling
        dataset_query_json query_json make_dense_dataset_f32_f64_i64 train_model model_from_json
        model_query_json evaluate_by_query predict_scores predict_to_trecrun""".split()
    assert len(reference_abi) == 20
    dll = ctypes.CDLL(_native.LIB_PATH)
    for name in set(names) | set(reference_abi):
        assert hasattr(dll, name), name


def test_defaults_match_reference(fr):
    ca = fr.query_json("coordinate_ascent_defaults")
    assert ca["measure"] == "ndcg" and ca["judgments"] is None
    p = ca["params"]["CoordinateAscent"]
    # coordinate_ascent.rs:25-41
    assert (p["num_restarts"], p["num_max_iterations"], p["step_base"], p["step_scale"], p["tolerance"]) == (5, 25, 0.05, 2.0, 0.001)
    assert (p["normalize"], p["quiet"], p["init_random"], p["output_ensemble"]) == (True, False, True, False)
    rf = fr.query_json("random_forest_defaults")["params"]["RandomForest"]
    # random_forest.rs:141-157
    assert (rf["num_trees"], rf["weight_trees"], rf["split_method"]) == (100, False, {"SquaredError": []})
    assert (rf["instance_sampling_rate"], rf["feature_sampling_rate"]) == (0.5, 0.25)
    assert (rf["min_leaf_support"], rf["split_candidates"], rf["max_depth"]) == (10, 3, 8)
    # Rand64::new(0xdeadbeef).rand_u64() under oorandom =11.1.0 (coordinate_ascent.rs:27-34)
    assert rf["seed"] == p["seed"] == 8208548815909702348
    with pytest.raises(Exception, match="unknown_query_str"):
        fr.query_json("nope")


def test_train_request_roundtrip(fr):
    req = fr.TrainRequest.coordinate_ascent()
    assert isinstance(req.params, fr.CoordinateAscentParams)
    again = fr.TrainRequest.from_dict(req.to_dict())
    assert again.to_dict() == req.to_dict()
    rf = fr.TrainRequest.random_forest()
    assert isinstance(rf.params, fr.RandomForestParams)
    assert rf.clone().to_dict() == rf.to_dict()
    big = req.clone()
    big.params.seed = (1 << 64) - 1
    assert json.loads(json.dumps(big.to_dict()))["params"]["CoordinateAscent"]["seed"] == (1 << 64) - 1


def test_load_dataset_and_introspection(fr, golden_dir):
    # reference tests/test_with_example_data.py:90-104
    rd = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"))
    assert rd.num_instances() == 782 and rd.num_features() == 6
    assert rd.feature_ids() == set(range(6))
    assert rd.feature_names() == set(str(i) for i in range(6))
    assert len(rd.queries()) == 45 and rd.is_sampled() is False
    by_q = rd.instances_by_query()
    assert sum(len(v) for v in by_q.values()) == 782
    named = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"),
                                     os.path.join(golden_dir, "trec_news_2018.features.json"))
    assert named.feature_names() == {"0", "pagerank", "para-fraction", "caption_count", "caption_partial", "caption_position"}
    assert named.feature_name_to_index()["pagerank"] == 4
    with pytest.raises(Exception, match="unknown_dataset_query_str"):
        named._query_json("nope")
    with pytest.raises(Exception):
        fr.CDataset.open_ranksvm("/nonexistent/file")


def test_sampling_views(fr, golden_dir):
    rd = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"),
                                  os.path.join(golden_dir, "trec_news_2018.features.json"))
    sub = rd.subsample_queries(["378", "363"])
    assert sub.is_sampled() and sub.queries() == {"378", "363"}
    assert sub.num_instances() == sum(len(v) for k, v in rd.instances_by_query().items() if k in ("378", "363"))
    assert sub.num_features() == 6
    with pytest.raises(ValueError):
        rd.subsample_queries(["not-a-query"])
    one = rd.subsample_feature_names(["pagerank"])
    assert one.feature_ids() == {4} and one.num_features() == 1 and one.num_instances() == 782
    both = sub.subsample_feature_names(["pagerank", "caption_count"])
    assert both.feature_ids() == {1, 4} and both.queries() == {"378", "363"}
    from fastrank_b200.clib import _take_result, ffi, lib

    with pytest.raises(Exception, match="Missing Features"):
        _take_result(lib.dataset_feature_sampling(rd.pointer, b"[99]"))
    with pytest.raises(Exception, match="No Features"):
        _take_result(lib.dataset_feature_sampling(rd.pointer, b"[]"))
    with pytest.raises(Exception, match="Dataset pointer is null"):
        _take_result(lib.dataset_feature_sampling(ffi.NULL, b"[1]"))


def test_qrel_roundtrip(fr, golden_dir):
    # reference :80-88
    q = fr.CQRel.load_file(os.path.join(golden_dir, "newsir18-entity.qrel"))
    assert len(q.queries()) == 50
    d = q.to_dict()
    q2 = fr.CQRel.from_dict(d)
    assert q2.to_dict() == d
    some = sorted(q.queries())[0]
    assert q.query_judgments(some) == d[some]
    with pytest.raises(ValueError):
        q.query_judgments("nope")
    with pytest.raises(Exception):
        fr.CQRel.load_file("/nonexistent/qrel")


def test_model_json_roundtrip(fr):
    tree = {"FeatureSplit": {"fid": 1, "split": 6.25, "lhs": {"LeafNode": 7.0},
                             "rhs": {"FeatureSplit": {"fid": 0, "split": -2.5, "lhs": {"LeafNode": -1.0}, "rhs": {"LeafNode": 12.0}}}}}
    specs = [
        {"Linear": {"weights": [0.1, -2.0, 1e-300, 3.0]}},
        {"SingleFeature": {"fid": 3, "dir": -1.0}},
        {"DecisionTree": tree},
        {"Ensemble": {"weights": [0.5, 2.0], "models": [{"DecisionTree": tree}, {"Linear": {"weights": [1.0]}}]}},
    ]
    for spec in specs:
        assert fr.CModel.from_dict(spec).to_dict() == spec
    from fastrank_b200.clib import _take_result, lib

    with pytest.raises(Exception, match="unknown variant"):
        _take_result(lib.model_from_json(b'{"Bogus": {}}'))
    with pytest.raises(Exception, match="missing field"):
        _take_result(lib.model_from_json(b'{"Linear": {}}'))
    with pytest.raises(Exception):
        _take_result(lib.model_from_json(b'{"Linear": '))
    with pytest.raises(Exception, match="NULL pointer"):
        _take_result(lib.model_from_json(fr.clib.ffi.NULL))


def test_from_numpy_validation(fr):
    import numpy as np

    X = np.zeros((4, 3), dtype=np.float32)
    y = np.zeros(4)
    with pytest.raises(AssertionError):
        fr.CDataset.from_numpy(X.astype(np.float64), y, np.zeros(4, dtype=np.int64))
    with pytest.raises(Exception, match="TryFromIntError"):
        fr.CDataset.from_numpy(X, y, np.array([0, 1, -5, 2], dtype=np.int64))
    ds = fr.CDataset.from_numpy(X, y, np.array([7, 7, 9, 9], dtype=np.int64))
    assert ds.queries() == {"7", "9"} and ds.num_features() == 3
    assert ds.instances_by_query() == {"7": [0, 1], "9": [2, 3]}


def test_compute_fails_loudly_without_gpu(fr, golden_dir):
    from fastrank_b200._native import lib

    if lib.fr_dev_device_count() > 0:
        pytest.skip("a GPU is present")
    rd = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"))
    m = fr.CModel.from_dict({"Linear": {"weights": [0.0, 1.0]}})
    for call in (lambda: rd.evaluate(m, "ndcg@5"), lambda: m.predict_scores(rd),
                 lambda: rd.train_model(fr.TrainRequest.coordinate_ascent()),
                 lambda: rd.evaluate_mean(m, "map"), lambda: m.predict_dense(rd)):
        with pytest.raises(Exception, match="no CUDA device|no CPU fallback"):
            call()


def test_compat_package_exposes_the_reference_names():
    # compat/ on PYTHONPATH makes `import fastrank` resolve to this implementation
    import importlib
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "compat"))
    try:
        sys.modules.pop("fastrank", None)
        fastrank = importlib.import_module("fastrank")
        from fastrank.clib import CDataset, CModel, CQRel  # noqa: F401
        from fastrank.training import CoordinateAscentParams, RandomForestParams, TrainRequest  # noqa: F401

        for name in ("CQRel", "CDataset", "CModel", "query_json", "TrainRequest", "CoordinateAscentParams",
                     "RandomForestParams"):
            assert hasattr(fastrank, name), name
        req = fastrank.TrainRequest.coordinate_ascent()
        assert req.params.num_restarts == 5 and req.measure == "ndcg"
    finally:
        sys.path.remove(os.path.join(root, "compat"))
        for k in [k for k in sys.modules if k == "fastrank" or k.startswith("fastrank.")]:
            sys.modules.pop(k)


def _load_text(fr, tmp_path, text, name="t.libsvm"):
    path = tmp_path / name
    path.write_text(text)
    return fr.CDataset.open_ranksvm(str(path))


def test_libsvm_parser_cases(fr, tmp_path):
    """The reference's libsvm.rs parser tests (libsvm.rs:318-432) through load_ranksvm_format:
    qid handling, unsorted features, comments as document names, and every parse error."""
    ds = _load_text(fr, tmp_path, "1 qid:A 1:1 2:1 3:1 # docA\n2 qid:B 6:1 3:1 4:0.5\n0 qid:A 2:7\n")
    assert ds.num_instances() == 3
    assert ds.queries() == {"A", "B"}
    assert sorted(len(v) for v in ds.instances_by_query().values()) == [1, 2]
    assert 6 in ds.feature_ids() and 1 in ds.feature_ids()
    for text, what in (
        ("1 qid:A what\n", "FeatureNoColon"),
        ("1 qid:A what:1.7\n", "FeatureNum"),
        ("1 qid:A 1:what\n", "FeatureValNotFloat"),
        ("nope qid:A 1:1\n", "Label"),
        ("nan qid:A 1:1\n", "LabelIsNan"),
        ("1 qid:A 1:1 1:2\n", "MultipleDefinitions"),
        ("1 1:1 2:1\n", "Missing qid"),
    ):
        with pytest.raises(Exception, match=what):
            _load_text(fr, tmp_path, text, name="bad.libsvm")
    with pytest.raises(Exception):
        fr.CDataset.open_ranksvm(str(tmp_path / "does-not-exist"))


def _zstd_compress(data: bytes) -> bytes:
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_compress.restype = ctypes.c_size_t
    z.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    cap = z.ZSTD_compressBound(len(data))
    buf = ctypes.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), 3)
    return buf.raw[:n]


def test_compressed_inputs_by_extension(fr, golden_dir, tmp_path):
    """io_helper.rs:18-29: .gz (multi-member), .bz2 and .zst files are read transparently, for
    datasets and for judgments."""
    import bz2
    import gzip

    train = open(os.path.join(golden_dir, "trec_news_2018.train"), "rb").read()
    qrel = open(os.path.join(golden_dir, "newsir18-entity.qrel"), "rb").read()
    plain_ds = fr.CDataset.open_ranksvm(os.path.join(golden_dir, "trec_news_2018.train"))
    plain_q = fr.CQRel.load_file(os.path.join(golden_dir, "newsir18-entity.qrel")).to_dict()
    half = train.index(b"\n", len(train) // 2) + 1
    codecs = {
        ".gz": lambda b: gzip.compress(b),
        ".bz2": lambda b: bz2.compress(b),
    }
    try:
        _zstd_compress(b"probe")
        codecs[".zst"] = _zstd_compress
    except OSError:
        pass
    for ext, enc in codecs.items():
        p = tmp_path / ("train.libsvm" + ext)
        # two concatenated members / frames for gzip, like MultiGzDecoder accepts
        p.write_bytes(enc(train[:half]) + enc(train[half:]) if ext == ".gz" else enc(train))
        ds = fr.CDataset.open_ranksvm(str(p))
        assert ds.num_instances() == plain_ds.num_instances() == 782
        assert ds.instances_by_query() == plain_ds.instances_by_query()
        assert ds.feature_ids() == plain_ds.feature_ids()
        q = tmp_path / ("judgments.qrel" + ext)
        q.write_bytes(enc(qrel))
        assert fr.CQRel.load_file(str(q)).to_dict() == plain_q
    bad = tmp_path / "broken.gz"
    bad.write_bytes(b"\x1f\x8b\x08\x00 not really gzip")
    with pytest.raises(Exception):
        fr.CQRel.load_file(str(bad))
    with pytest.raises(Exception, match="No such file"):
        fr.CDataset.open_ranksvm(str(tmp_path / "missing.gz"))


def test_large_libsvm_file_is_parsed_in_slices(fr, tmp_path):
    """Files of a few MB are cut at line boundaries and parsed by several threads: instance ids,
    query grouping and the first error (with its line number) must be those of a sequential read."""
    import numpy as np

    rng = np.random.default_rng(3)
    n, d = 30000, 24
    qid = np.sort(rng.integers(0, 900, n))
    X = np.round(rng.normal(size=(n, d)), 4).astype(np.float32)
    y = rng.integers(0, 5, n)
    lines = ["%d qid:%d %s # doc%d" % (y[i], qid[i], " ".join("%d:%r" % (j + 1, float(X[i, j])) for j in range(d)), i)
             for i in range(n)]
    text = "\n".join(lines) + "\n"
    assert len(text) > 4 << 20
    p = tmp_path / "big.libsvm"
    p.write_text(text)
    ds = fr.CDataset.open_ranksvm(str(p))
    assert ds.num_instances() == n and ds.num_features() == d + 1
    by_q = ds.instances_by_query()
    assert len(by_q) == len(np.unique(qid))
    for q in (int(qid[0]), int(qid[n // 2]), int(qid[-1])):
        assert by_q[str(q)] == [int(i) for i in np.nonzero(qid == q)[0]]
    # an error two thirds into the file is reported with its own line number
    bad_line = 2 * n // 3
    lines[bad_line] = "1 qid:7 3:what"
    (tmp_path / "big_bad.libsvm").write_text("\n".join(lines) + "\n")
    with pytest.raises(Exception, match=r"LineParseError\(%d, FeatureValNotFloat" % (bad_line + 1)):
        fr.CDataset.open_ranksvm(str(tmp_path / "big_bad.libsvm"))
    # two errors: the earlier one wins
    lines[100] = "nope qid:7 3:1"
    (tmp_path / "big_bad2.libsvm").write_text("\n".join(lines) + "\n")
    with pytest.raises(Exception, match=r"LineParseError\(101, Label"):
        fr.CDataset.open_ranksvm(str(tmp_path / "big_bad2.libsvm"))
