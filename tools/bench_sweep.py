#!/usr/bin/env python
"""Tuning harness for the batched sweep kernels: one dataset, one plan, the bench's 408-candidate
step timed (CUDA events around each launch, fr_dev_profile_*) under several settings of the
kernels' environment knobs, which are read per call.

    python tools/bench_sweep.py "FASTRANK_PACKED_MINB=4" "FASTRANK_PACKED_MINB=5" "FASTRANK_SWEEP_KERNEL=tile"
    N=3771125 Q=31531 python tools/bench_sweep.py ""
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from fastrank_b200.kernels import DevDataset, dense_query_index  # noqa: E402


def main():
    n = int(os.environ.get("N", 1_000_000))
    q = int(os.environ.get("Q", 30_000))
    steps = int(os.environ.get("STEPS", 10))
    X, y, qid = bench.make_data(n, 136, q)
    qidx, nq = dense_query_index(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    plan = dev.plan(int(os.environ.get("METRIC", 0)), int(os.environ.get("DEPTH", 10)))  # 0 NDCG, 1 AP, 2 RR; depth -1 = none
    packed = []
    n_sw = int(os.environ.get("SWEEPS", 8))  # restarts per launch (2-D sharding experiments: 4 or 2)
    for s in range(steps + 3):
        base, fids, ga, gb = bench.step_inputs(s, 136)
        packed.append(plan.pack_sweeps(base[:n_sw], fids[:n_sw], [a + b for a, b in zip(ga[:n_sw], gb[:n_sw])]))
    ref = None
    for setting in sys.argv[1:] or [""]:
        pairs = [kv.split("=", 1) for kv in setting.split() if "=" in kv]
        for k, v in pairs:
            os.environ[k] = v
        for s in range(3):
            plan.coord_sweeps_packed(packed[s])
        dev.profile(True)
        dev.profile_read(reset=True)
        sums = [plan.coord_sweeps_packed(packed[3 + s]).copy() for s in range(steps)]
        n_k, ms = dev.profile_read(reset=True)
        dev.profile(False)
        same = None
        if ref is None:
            ref = sums
        else:
            same = all(np.array_equal(a, b) for a, b in zip(ref, sums))
        print(json.dumps({"setting": setting, "launches": n_k, "ms_per_step": ms / steps,
                          "evals_per_s": 51 * n_sw * steps / (ms / 1e3), "same_sums_as_first": same}), flush=True)
        for k, _ in pairs:
            del os.environ[k]
    plan.close()
    dev.close()


if __name__ == "__main__":
    main()
