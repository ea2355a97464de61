import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def goldens():
    import json

    with open(os.path.join(GOLDEN, "reference_goldens.json")) as fp:
        return json.load(fp)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def trec_train(oracle):
    return oracle.load_libsvm(os.path.join(GOLDEN, "trec_news_2018.train"))


@pytest.fixture(scope="session")
def trec_test(oracle):
    return oracle.load_libsvm(os.path.join(GOLDEN, "trec_news_2018.test"))


# the reference's own test module is a fixture that tests/test_gpu_reference_suite.py runs in a
# subprocess (needs `import fastrank` -> compat/, cwd with examples/): not collected here
collect_ignore_glob = ["golden/ref_tests/*"]
