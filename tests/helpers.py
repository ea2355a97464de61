"""Shared test helpers: the BASELINE.md synthetic generator and thin wrappers over the
fr_dev_* kernel ABI (called through cffi, i.e. through the C ABI the product ships)."""
from __future__ import annotations

import numpy as np


def synth(n: int, d: int, q: int, seed: int = 20260417, shuffle_rows: bool = False, max_label: int = 4):
    """SURVEY.md 8d generator: four column families (normal / uniform / small integers for
    ties / tiny exponential), labels {0..4} at MSLR-like marginals from a noisy linear latent,
    qid = sort(integers(0, Q, N))."""
    rng = np.random.default_rng(seed)
    qid = np.sort(rng.integers(0, q, n)).astype(np.int64)
    X = np.empty((n, d), dtype=np.float32)
    for j in range(d):
        fam = j % 4
        if fam == 0:
            X[:, j] = rng.normal(size=n)
        elif fam == 1:
            X[:, j] = rng.random(n)
        elif fam == 2:
            X[:, j] = rng.integers(0, 10, n)
        else:
            X[:, j] = rng.exponential(size=n) * 1e-6
    w_true = rng.normal(size=d)
    Xs = (X - X.mean(0)) / (X.std(0) + 1e-12)
    latent = Xs @ w_true / np.sqrt(d) + rng.normal(size=n)
    cuts = np.quantile(latent, [0.515, 0.84, 0.974, 0.992])
    y = np.searchsorted(cuts, latent).astype(np.float64)
    y = np.minimum(y, max_label)
    if shuffle_rows:
        perm = rng.permutation(n)
        X, y, qid = np.ascontiguousarray(X[perm]), y[perm], qid[perm]
    return X, y, qid


def oracle_dataset(orc, X, y, qid):
    return orc.OracleDataset(X, y.astype(np.float32), [str(int(v)) for v in qid])


FX = float(1 << 40)


def fx_sum(values) -> int:
    """What the device accumulates: sum of round-to-nearest-even(v * 2^40)."""
    return int(np.rint(np.asarray(values, dtype=np.float64) * FX).astype(np.int64).sum())


def __getattr__(name):
    # the kernel wrappers dlopen libfastrank_b200.so: loaded on first use only, so that code which
    # needs the generator alone (bench.py --impl reference) never maps the CUDA library
    if name in ("DevDataset", "DevPlan"):
        from fastrank_b200 import kernels

        return getattr(kernels, name)
    raise AttributeError(name)


def dense_qidx(qid):
    """Dense query numbers in order of first appearance (what the host passes down)."""
    seen = {}
    out = np.empty(len(qid), dtype=np.uint32)
    for i, v in enumerate(qid):
        out[i] = seen.setdefault(int(v), len(seen))
    return out, len(seen)
