"""The N>1 path on CPU: world_size-2 (and 3) gloo jobs exercising the sharding and the
fixed-point reduction protocol of fastrank_b200.dist.  The NCCL data path itself
is covered by the -m gpu test in tests/test_gpu_dist.py."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import fx_sum, oracle_dataset, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_rows_partitions_whole_queries():
    from fastrank_b200 import dist as frdist

    X, y, qid = synth(5000, 4, 170, seed=2, shuffle_rows=True)
    for world in (1, 2, 3, 8):
        seen = []
        queries = []
        for r in range(world):
            rows = frdist.shard_rows(qid, r, world)
            assert np.all(np.diff(rows) > 0)  # original relative order kept: id tie-break preserved
            seen.append(rows)
            queries.append(set(qid[rows].tolist()))
        allrows = np.concatenate(seen)
        assert sorted(allrows.tolist()) == list(range(len(qid)))
        for a in range(world):
            for b in range(a + 1, world):
                assert not (queries[a] & queries[b])  # a query never straddles ranks
        if world > 1:
            sizes = [len(s) for s in seen]
            assert max(sizes) - min(sizes) < 0.2 * len(qid) / world + 60  # balanced by documents


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_reduction_matches_single_process(oracle, world, tmp_path):
    out = tmp_path / "dist.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", env["MASTER_PORT"],
           os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    res = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.load(open(out))
    # single-process answer
    X, y, qid = synth(6000, 12, 200, seed=5, shuffle_rows=True)
    ods = oracle_dataset(oracle, X, y, qid)
    W = np.random.default_rng(0).normal(size=(2, 12))
    exp = [fx_sum(oracle.evaluate_scores(ods, oracle.score_linear(X, w), "ndcg@10")) for w in W]
    ranks = got["ranks"]
    assert len(ranks) == world
    for sums, nq, uid, nrows, queries in ranks:
        assert sums == exp          # integer sums: independent of the number of ranks
        assert nq == ods.nq
        assert uid == bytes(range(128)).hex()
    assert sum(r[3] for r in ranks) == len(qid)
