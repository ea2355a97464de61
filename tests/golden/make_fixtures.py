"""Regenerates tests/golden/ from the read-only reference checkout (run in the authoring
container only; /root/reference does not exist on the GPU box).

* copies the reference's example DATA files (the fixture its own goldens are defined on;
  reference tests/test_with_example_data.py:53-57) -- data, not source code;
* copies the reference's Python integration test module, byte for byte, to
  ref_tests/test_with_example_data.py: it is the acceptance test of the drop-in boundary
   and is RUN, unmodified, against compat/fastrank by
  tests/test_gpu_reference_suite.py -- a fixture, never imported by the product;
* writes reference_goldens.json: every known-answer value the reference's tests hold for
  the score -> rank -> metric path, with the file:line each one comes from.
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = [
    "examples/trec_news_2018.train",
    "examples/trec_news_2018.test",
    "examples/trec_news_2018.features.json",
    "examples/newsir18-entity.qrel",
]

GOLDENS = {
    "single_feature_ndcg5": {
        "source": "tests/test_with_example_data.py:16-23 (asserted :139-167, 7 places)",
        "values": {
            "0": 0.10882970494872854,
            "para-fraction": 0.43942925167146063,
            "caption_position": 0.3838323029697044,
            "caption_count": 0.363671198812673,
            "pagerank": 0.28879573536768505,
            "caption_partial": 0.2119744912371782,
        },
    },
    "rank_ties": {
        "source": "src/evaluators.rs:61-79",
        "input": [[2.0, 0.0, 4], [2.0, 1.0, 3], [2.0, 2.0, 1], [2.0, 2.0, 2], [1.0, 2.0, 5]],
        "expected_ids": [4, 3, 1, 2, 5],
    },
    "compute_ndcg": {
        "source": "src/evaluators.rs:285-295",
        "gains": [0.0, 1.0, 1.0, 1.0, 0.0, 0.0],
        "expected": 0.7328,
        "tolerance": 0.00005,
    },
    "rf_determinism_ndcg5": {
        "source": "src/random_forest.rs:427-463; tests/test_with_example_data.py:175-201",
        "expected": 0.4367914517387043,
        "tolerance": 1e-9,
        "params": {"num_trees": 10, "seed": 42, "min_leaf_support": 1, "max_depth": 10, "split_candidates": 32,
                   "split_method": "SquaredError", "instance_sampling_rate": 0.5, "feature_sampling_rate": 0.25},
        "note": "pins the oorandom =11.1.0 stream (rand_u64 output function, rand_range without range.start)",
    },
    "oorandom_known_answers": {
        "source": "Cargo.toml:18-19 (oorandom =11.1.0); coordinate_ascent.rs:27-34 (default seed)",
        "seed_42_first_three_u64": [12410087264455502793, 359948335059059064, 14020691344464510033],
        "default_seed": 8208548815909702348,
    },
    "notebook_coordinate_ascent": {
        "source": "examples/FastRankDemo.ipynb cells 4-5 (fastrank 0.4.1 output kept in the notebook)",
        "request": {"measure": "ndcg", "seed": 1234567, "init_random": True, "normalize": True},
        "weights": [2.1789363075430108e-09, 3.229859575980813e-06, -6.60633974465588e-07, 0.0,
                    -0.9999773758216621, 1.9152791882e-05],
        "weights_tolerance": 1e-12,
        "test_ndcg5_printed": "0.928",
    },
    "notebook_random_forest": {
        "source": "examples/FastRankDemo.ipynb cells 3, 5",
        "params": {"num_trees": 100, "seed": 1234567, "feature_sampling_rate": 0.5, "instance_sampling_rate": 0.5},
        "test_ndcg5_printed": "0.901",
    },
    "regression_tree": {
        "source": "src/random_forest.rs:465-506",
        "xs": [1, 1, 2, 3, 4, 5, 6, 7, 8, 9],
        "ys": [7, 7, 7, 7, 2, 2, 2, 12, 12, 12],
    },
    "dataset_shape": {
        "source": "tests/test_with_example_data.py:24-52",
        "n": 782, "d": 6, "n_queries": 45, "qrel_queries": 50,
    },
}


def main():
    digests = {}
    for rel in FILES:
        dst = os.path.join(HERE, os.path.basename(rel))
        shutil.copyfile(os.path.join(REF, rel), dst)
        os.chmod(dst, 0o644)
        digests[os.path.basename(rel)] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    os.makedirs(os.path.join(HERE, "ref_tests"), exist_ok=True)
    dst = os.path.join(HERE, "ref_tests", "test_with_example_data.py")
    shutil.copyfile(os.path.join(REF, "tests", "test_with_example_data.py"), dst)
    os.chmod(dst, 0o644)
    digests["ref_tests/test_with_example_data.py"] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    GOLDENS["sha256"] = digests
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as fp:
        json.dump(GOLDENS, fp, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
