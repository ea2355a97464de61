"""Worker for tests/test_gpu_dist.py: one rank of a query-sharded NCCL job on real GPUs.
Rank 0 writes the all-reduced sums; the test compares them with a single-GPU run."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


OTHER_MEASURES = (("mrr", 2), ("map", 1), ("ndcg", 0))


def workload():
    from tests.helpers import synth

    X, y, qid = synth(40000, 24, 1200, seed=31, shuffle_rows=True)
    rng = np.random.default_rng(1)
    base = rng.uniform(-1, 1, size=(8, 24))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int(v) for v in rng.integers(0, 24, 8)]
    cands = [[0.0] + [base[r, fids[r]] - 0.05 * (2.0 ** k - 1) for k in range(1, 26)] for r in range(8)]
    W = rng.normal(size=(3, 24))
    return X, y, qid, base, fids, cands, W


def many_rows():
    """8 sweeps x 80 candidates: more rows than one pass holds, so the cross-GPU reduction has to
    ride on the LAST pass only."""
    rng = np.random.default_rng(2)
    base = rng.normal(size=(8, 24))
    fids = [int(v) for v in rng.integers(0, 24, 8)]
    cands = [[float(v) for v in rng.normal(size=80)] for _ in range(8)]
    return base, fids, cands


def long_list_data(kind):
    """Integer features (exact arithmetic: every schedule gives the same bits) and lists too long
    for a sweep tile.  "hybrid": one 700-document list per rank among short ones, so both ranks
    rank it from HBM next to the batched sweep.  "mixed": the long lists dominate the LAST rank's
    shard only -- that rank cannot use the batched sweep, and the ranks have to agree on the
    exact-order entry point instead of waiting on each other."""
    rng = np.random.default_rng(3 if kind == "hybrid" else 4)
    if kind == "hybrid":
        lens = [int(v) for v in rng.integers(5, 80, 300)]
        lens.insert(100, 700)
        lens.append(700)
    else:
        lens = [int(v) for v in rng.integers(20, 60, 60)] + [800, 800]
    qid = np.concatenate([np.full(l, i) for i, l in enumerate(lens)]).astype(np.int64)
    n = len(qid)
    X = rng.integers(-3, 4, size=(n, 5)).astype(np.float32)
    y = (rng.integers(0, 5, n) * (rng.random(n) < 0.5)).astype(np.float64)
    return X, y, qid


def train_long_lists(fr, X, y, qid):
    ds = fr.CDataset.from_numpy(X, y, qid)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@10"
    req.params.num_restarts, req.params.seed, req.params.quiet = 2, 3, True
    req.params.init_random = False
    model = ds.train_model(req)
    return model.to_dict()["Linear"]["weights"], ds.evaluate_mean(model, "ndcg@10")


def main():
    import torch
    import torch.distributed as dist

    import fastrank_b200 as fr
    from fastrank_b200 import dist as frdist
    from fastrank_b200._native import lib
    from fastrank_b200.kernels import DevDataset, dense_query_index

    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = frdist.init_communicator(rank, world, local, install_default=True)
    X, y, qid, base, fids, cands, W = workload()
    rows = frdist.shard_rows(qid, rank, world)
    Xl, yl, ql = np.ascontiguousarray(X[rows]), np.ascontiguousarray(y[rows]), np.ascontiguousarray(qid[rows])
    qidx, nq = dense_query_index(ql)
    dev = DevDataset(Xl, yl.astype(np.float32), qidx, nq, device=local)
    plan = dev.plan(0, 10)
    assert lib.fr_dev_plan_set_comm(plan.ptr, comm.ptr) == 0
    fast = plan.coord_sweeps(base, fids, cands, fast=True)
    exact = plan.coord_sweeps(base, fids, cands)
    lin, _ = plan.eval_linear(W, per_query=False)
    mbase, mfids, mcands = many_rows()
    many = plan.coord_sweeps(mbase, mfids, mcands, fast=True)
    nq_global = int(lib.fr_dev_plan_global_queries(plan.ptr))
    # the measures served by the slot mode of the packed kernel (and RR's one-rank path) reduce
    # through the same kernel tail
    others = {}
    for name, metric in OTHER_MEASURES:
        p2 = dev.plan(metric, -1)
        assert lib.fr_dev_plan_set_comm(p2.ptr, comm.ptr) == 0
        others[name] = p2.coord_sweeps(base, fids, cands, fast=True).tolist()
    dev.close()
    # the reference-compatible surface on a shard: train_model sees all-reduced means
    ds = fr.CDataset.from_numpy(Xl, yl, ql)
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@10"
    req.params.num_restarts = 2
    req.params.seed = 7
    req.params.quiet = True
    model = ds.train_model(req)
    weights = model.to_dict()["Linear"]["weights"]
    mean = ds.evaluate_mean(model, "ndcg@10")
    long_lists = {}
    for kind in ("hybrid", "mixed"):
        X2, y2, q2 = long_list_data(kind)
        rows2 = frdist.shard_rows(q2, rank, world)
        long_lists[kind] = train_long_lists(fr, np.ascontiguousarray(X2[rows2]), np.ascontiguousarray(y2[rows2]),
                                            np.ascontiguousarray(q2[rows2]))
    gathered = [None] * world
    dist.all_gather_object(gathered, (fast.tolist(), exact.tolist(), lin.tolist(), nq_global, weights, mean,
                                      many.tolist(), long_lists, others))
    if rank == 0:
        with open(out_path, "w") as fp:
            json.dump({"world": world, "ranks": gathered}, fp)
    del ds, model
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
