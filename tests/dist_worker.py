"""Worker for tests/test_dist_cpu.py: one rank of a world_size-N gloo job (CPU only).

Each rank takes its query shard (fastrank_b200.dist.shard_rows), computes the per-query metric
of two weight vectors on the shard with the CPU oracle (test infrastructure standing in for
the kernels, which need a GPU), converts to the library's 2^-40 fixed point and combines sums
and query counts through the reduction protocol the NCCL path implements
(fastrank_b200.dist.combine_fixed_point).  Rank 0 writes what it saw to a JSON file.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from fastrank_b200 import dist as frdist
    from oracle import oracle as orc
    from tests.helpers import fx_sum, oracle_dataset, synth

    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, y, qid = synth(6000, 12, 200, seed=5, shuffle_rows=True)
    rows = frdist.shard_rows(qid, rank, world)
    Xl, yl, ql = X[rows], y[rows], qid[rows]
    ods = oracle_dataset(orc, np.ascontiguousarray(Xl), yl, ql)
    rng = np.random.default_rng(0)  # same weights on every rank: replicated host state
    W = rng.normal(size=(2, 12))
    local = [fx_sum(orc.evaluate_scores(ods, orc.score_linear(np.ascontiguousarray(Xl), w), "ndcg@10")) for w in W]
    sums, nq = frdist.combine_fixed_point(np.asarray(local, dtype=np.int64), ods.nq)
    uid = frdist.exchange_unique_id(lambda: bytes(range(128)), rank, world)
    # every rank must hold the same reduced numbers (identical host replay, no broadcast needed)
    gathered = [None] * world
    dist.all_gather_object(gathered, (sums.tolist(), nq, uid.hex(), len(rows), sorted(set(ql.tolist()))))
    if rank == 0:
        with open(out_path, "w") as fp:
            json.dump({"world": world, "ranks": gathered}, fp)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
