"""The reference's own integration test module, run UNMODIFIED against this repo.

tests/golden/ref_tests/test_with_example_data.py is a byte-for-byte copy of the reference's
tests/test_with_example_data.py (sha256 pinned in reference_goldens.json).  It does
`import fastrank` and opens `examples/...` relative to the working directory, so it is run in a
subprocess with compat/ on PYTHONPATH (compat/fastrank re-exports fastrank_b200) from a scratch
directory whose examples/ holds the fixture data.  SURVEY.md section 7 step 2's acceptance test.
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TEST = os.path.join(ROOT, "tests", "golden", "ref_tests", "test_with_example_data.py")
DATA = ["trec_news_2018.train", "trec_news_2018.test", "trec_news_2018.features.json", "newsir18-entity.qrel"]


def test_reference_module_is_the_unmodified_copy(goldens):
    digest = hashlib.sha256(open(REF_TEST, "rb").read()).hexdigest()
    assert digest == goldens["sha256"]["ref_tests/test_with_example_data.py"]


def test_reference_test_module_passes_unmodified(tmp_path, golden_dir):
    examples = tmp_path / "examples"
    examples.mkdir()
    for name in DATA:
        shutil.copyfile(os.path.join(golden_dir, name), examples / name)
    shutil.copyfile(REF_TEST, tmp_path / "test_with_example_data.py")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, env.get("PYTHONPATH", "")])
    res = subprocess.run([sys.executable, "-m", "unittest", "-v", "test_with_example_data"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    m = re.search(r"Ran (\d+) tests", out)
    assert m and int(m.group(1)) == 14, out[-2000:]  # every test_ method of TestRustAPI
    assert "test_random_forest " in out and "OK" in out
    assert "skipped" not in out.lower()
