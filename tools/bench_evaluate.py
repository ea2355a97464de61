#!/usr/bin/env python
"""SURVEY.md 8(d), first row: the full-rescore sweep -- C arbitrary weight vectors scored and
reduced to mean NDCG@10 in ONE pass over the 1M x 136 matrix (evaluators.rs:173-224 for C
models at once; what `dataset.evaluate` and coordinate ascent's start scores call).  Prints one
JSON line per C; not the headline bench (bench.py is)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from fastrank_b200.kernels import DevDataset, dense_query_index
    from tests.helpers import synth

    n = int(os.environ.get("N_DOCS", 1_000_000))
    d = 136
    q = max(n // 33, 1)
    reps = int(os.environ.get("REPS", 20))
    peak = 8000.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    X, y, qid = synth(n, d, q)
    qidx, nq = dense_query_index(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq, device=0)
    plan = dev.plan(0, 10)
    rng = np.random.default_rng(11)
    settings = sys.argv[1:] or [""]  # e.g. "FASTRANK_TMA_EVAL=1 FASTRANK_TMA_EVAL_CFG=2" (read per call)
    for setting in settings:
      pairs = [kv.split("=", 1) for kv in setting.split() if "=" in kv]
      for k, v in pairs:
        os.environ[k] = v
      for c in [int(v) for v in os.environ.get("CANDS", "1,8,26").split(",")]:
          W = np.random.default_rng(11 + c).normal(size=(c, d))
          plan.eval_linear(W, per_query=False)
          dev.profile(True)
          dev.profile_read(reset=True)
          for _ in range(reps):
              sums, _ = plan.eval_linear(W, per_query=False)
          launches, ms = dev.profile_read(reset=True)
          dev.profile(False)
          per_pass_ms = ms / reps
          algo = n * d * 4 + n * 4 + (q + 1) * 4 + c * d * 8 + c * 8
          print(json.dumps({
              "setting": setting,
              "workload": "full rescore, %d weight vectors per pass over %d x %d, ndcg@10" % (c, n, d),
              "launches_per_pass": launches / reps, "ms_per_pass": per_pass_ms,
              "evals_per_s": c / per_pass_ms * 1e3,
              "algorithmic_bytes_per_pass": algo, "achieved_GBps": algo / per_pass_ms / 1e6,
              "peak_GBps": peak, "frac": algo / per_pass_ms / 1e6 / peak,
              "mean_ndcg10_first": float(sums[0]) * 2.0 ** -40 / nq,
          }), flush=True)
      for k, _ in pairs:
        del os.environ[k]
    dev.close()


if __name__ == "__main__":
    main()
