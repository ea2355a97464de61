"""Thin Python wrappers over the fr_dev_* kernel ABI (include/fastrank_b200.h, part 2), called
through cffi -- i.e. through the C ABI the product ships.  Used by bench.py, the parity tests
and fastrank_b200.dist; the reference-compatible surface lives in clib.py."""
from __future__ import annotations

import numpy as np

from ._native import ffi as _ffi, lib as _lib

FX_BITS = 40
FX_SCALE = float(1 << FX_BITS)


class DevDataset:
    """fr_dev_dataset + fr_dev_plan through cffi."""

    def __init__(self, X, gains, qidx, nq, device=0):
        self.ffi, self.lib = _ffi, _lib
        self.X = np.ascontiguousarray(X, dtype=np.float32)
        self.gains = np.ascontiguousarray(gains, dtype=np.float32)
        self.qidx = np.ascontiguousarray(qidx, dtype=np.uint32)
        self.nq = int(nq)
        ffi, lib = _ffi, _lib
        out = ffi.new("fr_dev_dataset**")
        rc = lib.fr_dev_dataset_create(device, self.X.shape[0], self.X.shape[1],
                                       ffi.cast("float*", self.X.ctypes.data),
                                       ffi.cast("float*", self.gains.ctypes.data),
                                       ffi.cast("uint32_t*", self.qidx.ctypes.data), self.nq, out)
        self._check(rc)
        self.ptr = out[0]
        self.plans = []

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.ffi.string(self.lib.fr_dev_last_error()).decode())

    def plan(self, metric: int, depth: int = -1, query_ids=None, inst=None, norms=None):
        ffi, lib = self.ffi, self.lib
        desc = ffi.new("fr_dev_plan_desc*")
        keep = []
        desc.metric = metric
        desc.depth = depth
        if query_ids is None:
            desc.n_queries = self.nq
            desc.query_ids = ffi.NULL
        else:
            qi = np.ascontiguousarray(query_ids, dtype=np.uint32)
            keep.append(qi)
            desc.n_queries = len(qi)
            desc.query_ids = ffi.cast("uint32_t*", qi.ctypes.data)
        if inst is not None:
            off, ids = inst
            off = np.ascontiguousarray(off, dtype=np.uint64)
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            keep += [off, ids]
            desc.inst_off = ffi.cast("uint64_t*", off.ctypes.data)
            desc.inst_ids = ffi.cast("uint32_t*", ids.ctypes.data)
        if norms is not None:
            present, value = norms
            present = np.ascontiguousarray(present, dtype=np.uint8)
            value = np.ascontiguousarray(value, dtype=np.float64)
            keep += [present, value]
            desc.norm_present = ffi.cast("uint8_t*", present.ctypes.data)
            desc.norm_value = ffi.cast("double*", value.ctypes.data)
        out = ffi.new("fr_dev_plan**")
        self._check(lib.fr_dev_plan_create(self.ptr, desc, out))
        p = DevPlan(self, out[0], int(desc.n_queries))
        self.plans.append(p)
        return p

    def score_model(self, code: np.ndarray) -> np.ndarray:
        ffi, lib = self.ffi, self.lib
        code = np.ascontiguousarray(code, dtype=np.uint64)
        m = ffi.new("fr_dev_model**")
        self._check(lib.fr_dev_model_create(self.ptr, ffi.cast("uint64_t*", code.ctypes.data), len(code), m))
        out = np.empty(self.X.shape[0], dtype=np.float64)
        rc = lib.fr_dev_score_model(self.ptr, m[0], ffi.cast("double*", out.ctypes.data))
        lib.fr_dev_model_destroy(m[0])
        self._check(rc)
        return out

    # device-side timing on the library's own stream
    def timer_start(self):
        self._check(self.lib.fr_dev_timer_start(self.ptr))

    def timer_stop(self) -> float:
        ms = self.ffi.new("double*")
        self._check(self.lib.fr_dev_timer_stop(self.ptr, ms))
        return float(ms[0])

    def profile(self, on: bool):
        self._check(self.lib.fr_dev_profile_enable(self.ptr, 1 if on else 0))

    def profile_read(self, reset: bool = True):
        n = self.ffi.new("uint64_t*")
        ms = self.ffi.new("double*")
        self._check(self.lib.fr_dev_profile_read(self.ptr, n, ms, 1 if reset else 0))
        return int(n[0]), float(ms[0])

    def bytes(self) -> int:
        return int(self.lib.fr_dev_dataset_bytes(self.ptr))

    def close(self):
        for p in self.plans:
            p.close()
        self.plans = []
        if self.ptr is not None:
            self.lib.fr_dev_dataset_destroy(self.ptr)
            self.ptr = None


class DevPlan:
    def __init__(self, ds: DevDataset, ptr, nq: int):
        self.ds, self.ptr, self.nq = ds, ptr, nq

    def eval_linear(self, W: np.ndarray, per_query: bool = True):
        ffi, lib = self.ds.ffi, self.ds.lib
        W = np.ascontiguousarray(W, dtype=np.float64)
        c, wlen = W.shape
        sums = np.zeros(c, dtype=np.int64)
        pq = np.zeros((c, self.nq), dtype=np.float64) if per_query else None
        rc = lib.fr_dev_eval_linear_batch(self.ptr, ffi.cast("double*", W.ctypes.data), wlen, c,
                                          ffi.cast("int64_t*", sums.ctypes.data),
                                          ffi.cast("double*", pq.ctypes.data) if per_query else ffi.NULL)
        self.ds._check(rc)
        return sums, pq

    def coord_sweeps(self, base_w: np.ndarray, fids, cands, fast: bool = False, per_query: bool = False):
        """cands: list (one per sweep) of candidate-weight lists.  fast=True calls the batched
        sweep (fr_dev_eval_coord_sweeps_fast); per_query (fast only) also returns
        [sweep][candidate][query] values."""
        ffi, lib = self.ds.ffi, self.ds.lib
        base_w = np.ascontiguousarray(base_w, dtype=np.float64)
        r, wlen = base_w.shape
        stride = max(len(c) for c in cands)
        cw = np.zeros((r, stride), dtype=np.float64)
        nc = np.zeros(r, dtype=np.uint32)
        for i, c in enumerate(cands):
            cw[i, : len(c)] = c
            nc[i] = len(c)
        fid = np.ascontiguousarray(fids, dtype=np.uint32)
        sums = np.zeros((r, stride), dtype=np.int64)
        if fast:
            pq = np.zeros((r, stride, self.nq), dtype=np.float64) if per_query else None
            rc = lib.fr_dev_eval_coord_sweeps_fast(self.ptr, r, ffi.cast("double*", base_w.ctypes.data), wlen,
                                                   ffi.cast("uint32_t*", fid.ctypes.data),
                                                   ffi.cast("double*", cw.ctypes.data),
                                                   ffi.cast("uint32_t*", nc.ctypes.data), stride,
                                                   ffi.cast("int64_t*", sums.ctypes.data),
                                                   ffi.cast("double*", pq.ctypes.data) if per_query else ffi.NULL)
            self.ds._check(rc)
            return (sums, pq) if per_query else sums
        rc = lib.fr_dev_eval_coord_sweeps(self.ptr, r, ffi.cast("double*", base_w.ctypes.data), wlen,
                                          ffi.cast("uint32_t*", fid.ctypes.data),
                                          ffi.cast("double*", cw.ctypes.data),
                                          ffi.cast("uint32_t*", nc.ctypes.data), stride,
                                          ffi.cast("int64_t*", sums.ctypes.data))
        self.ds._check(rc)
        return sums

    def pack_sweeps(self, base_w: np.ndarray, fids, cands):
        """Builds the flat arrays fr_dev_eval_coord_sweeps[_fast] takes (done once, outside any
        timed region, by callers that replay the same sweeps)."""
        base_w = np.ascontiguousarray(base_w, dtype=np.float64)
        r = base_w.shape[0]
        stride = max(len(c) for c in cands)
        cw = np.zeros((r, stride), dtype=np.float64)
        nc = np.zeros(r, dtype=np.uint32)
        for i, c in enumerate(cands):
            cw[i, : len(c)] = c
            nc[i] = len(c)
        return base_w, np.ascontiguousarray(fids, dtype=np.uint32), cw, nc, np.zeros((r, stride), dtype=np.int64)

    def coord_sweeps_packed(self, packed, fast: bool = True):
        ffi, lib = self.ds.ffi, self.ds.lib
        base_w, fid, cw, nc, sums = packed
        args = (self.ptr, base_w.shape[0], ffi.cast("double*", base_w.ctypes.data), base_w.shape[1],
                ffi.cast("uint32_t*", fid.ctypes.data), ffi.cast("double*", cw.ctypes.data),
                ffi.cast("uint32_t*", nc.ctypes.data), cw.shape[1], ffi.cast("int64_t*", sums.ctypes.data))
        rc = lib.fr_dev_eval_coord_sweeps_fast(*args, ffi.NULL) if fast else lib.fr_dev_eval_coord_sweeps(*args)
        self.ds._check(rc)
        return sums

    def close(self):
        if self.ptr is not None:
            self.ds.lib.fr_dev_plan_destroy(self.ptr)
            self.ptr = None




def dense_query_index(qid):
    """Dense query numbers in order of first appearance (what the host passes down)."""
    qid = np.asarray(qid)
    _, first, inverse = np.unique(qid, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(order), dtype=np.uint32)
    rank[order] = np.arange(len(order), dtype=np.uint32)
    return rank[inverse].astype(np.uint32), int(len(first))
