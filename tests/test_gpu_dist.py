"""Query-sharded multi-GPU path on real GPUs: 2 ranks over NCCL must produce the
same integer sums -- and therefore the same coordinate-ascent model -- as one GPU holding
everything.  Skipped on a single-GPU box (the gloo tests in test_dist_cpu.py cover the host
protocol there)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_match_one_gpu(tmp_path):
    import fastrank_b200 as fr
    from fastrank_b200._native import lib
    from fastrank_b200.kernels import DevDataset, dense_query_index
    from tests.dist_gpu_worker import OTHER_MEASURES, long_list_data, many_rows, train_long_lists, workload

    if lib.fr_dev_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "dist_gpu.json"
    port = str(_free_port())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", port,
           os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(out)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    got = json.load(open(out))
    X, y, qid, base, fids, cands, W = workload()
    qidx, nq = dense_query_index(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    try:
        plan = dev.plan(0, 10)
        fast = plan.coord_sweeps(base, fids, cands, fast=True)
        exact = plan.coord_sweeps(base, fids, cands)
        lin, _ = plan.eval_linear(W, per_query=False)
        mbase, mfids, mcands = many_rows()
        many = plan.coord_sweeps(mbase, mfids, mcands, fast=True)
        others = {name: dev.plan(metric, -1).coord_sweeps(base, fids, cands, fast=True).tolist()
                  for name, metric in OTHER_MEASURES}
    finally:
        dev.close()
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@10"
    req.params.num_restarts = 2
    req.params.seed = 7
    req.params.quiet = True
    model = ds.train_model(req)
    long_lists = {kind: train_long_lists(fr, *long_list_data(kind)) for kind in ("hybrid", "mixed")}
    for r_fast, r_exact, r_lin, nq_global, weights, mean, r_many, r_long, r_others in got["ranks"]:
        assert r_others == others   # MRR / MAP / NDCG without cut-off: slot mode + the fused reduction
        for kind in ("hybrid", "mixed"):   # lists too long for a sweep tile, evenly and unevenly sharded
            assert r_long[kind][0] == long_lists[kind][0], kind
            assert r_long[kind][1] == long_lists[kind][1], kind
        assert r_many == many.tolist()
        assert nq_global == nq
        assert r_exact == exact.tolist()   # fixed-point sums do not depend on the sharding
        assert r_fast == fast.tolist()
        assert r_lin == lin.tolist()
        assert weights == model.to_dict()["Linear"]["weights"]
        assert mean == ds.evaluate_mean(model, "ndcg@10")
