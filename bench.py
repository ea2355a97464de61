#!/usr/bin/env python
"""bench.py -- coordinate-ascent NDCG@10 evaluations/s on the 1M x 136 x 30k-query synthetic
DenseDataset (BASELINE.json configs[1]; configs[2] when --gpus > 1: the same dataset sharded
by query, one NCCL all-reduce of the metric sums per group of candidates).

One STEP = one coordinate-ascent line search for each of the 8 restarts
(coordinate_ascent.rs:131-177): direction 0, -1 and +1 = 51 candidate weight vectors per restart
= 408 evaluate_mean equivalents (evaluators.rs:173-184) in ONE call of
fr_dev_eval_coord_sweeps_fast = one kernel launch = one pass over X.  Every candidate is a full
evaluation: all N documents scored, ranked inside their query, NDCG@10 per query, mean over
queries.

  value      evaluations/s with the dataset resident in HBM (device-timed, CUDA events on the
             library's stream, max over ranks)
  e2e        the same metric through the reference-facing C ABI with HOST buffers:
             make_dense_dataset_f32_f64_i64 + train_model (CA, 8 restarts, to convergence);
             the upload of X/y/qid is inside the timed region, evaluations counted are the
             ones the reference's control flow consumes
  roofline   dominant kernel (sweep_packed_kernel, the batched sweep for NDCG@k) vs the measured
             HBM copy bandwidth, plus what actually bounds it (issue slots / FP64 pipe, from the
             committed ncu capture)
  side_trees / side_mslr (N = 1, after the timed region, not part of it)
             BASELINE.json configs[3] (500 trees x depth 8 scored over the same 1M x 136) and
             configs[4]'s shape (3 771 125 x 136 x 31 531 queries, the same CA step)
  cpu_baseline / --impl reference
             the CPU oracle (C restatement of the reference; the Rust reference cannot be
             built in this image) on the host cores, one thread per restart as the reference
             does (coordinate_ascent.rs:216), on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

JSON_OUT = sys.stdout
# BASELINE.json configs[1]; the environment overrides exist for side experiments only (e.g. the
# MSLR-shaped configs[4]: FASTRANK_BENCH_N=3771125 FASTRANK_BENCH_Q=31531) and rename the workload
N_DOCS = int(os.environ.get("FASTRANK_BENCH_N", 1_000_000))
N_FEAT = 136
N_QUERIES = int(os.environ.get("FASTRANK_BENCH_Q", 30_000))
N_RESTARTS = 8
T_ITERS = 25
DEPTH = 10
METRIC = "CA NDCG@10 evaluations/sec on 1M x 136 DenseDataset"
UNIT = "evals/s"
WORKLOAD = "synthetic 1M docs x 136 features x 30k queries, coordinate_ascent 8 restarts, ndcg@10"
LONG_TAIL = os.environ.get("FASTRANK_BENCH_TAIL", "") == "1"  # side experiment: MSLR-like list lengths
if (N_DOCS, N_QUERIES) != (1_000_000, 30_000) or LONG_TAIL:
    WORKLOAD = "synthetic %d docs x 136 features x %d queries%s, coordinate_ascent 8 restarts, ndcg@10" % (
        N_DOCS, N_QUERIES, " (log-normal list lengths, up to 1300 documents)" if LONG_TAIL else "")


def base_config(world: int) -> dict:
    """The keys both arms print (the driver compares them)."""
    return {"workload": WORKLOAD, "evals_per_step": EVALS_PER_STEP,
            "l2": "inputs (544 MB feature matrix per sweep) larger than the 126 MB L2",
            "parallelism": "query-sharded x%d" % world if world > 1 else "single GPU"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_data(n=N_DOCS, d=N_FEAT, q=N_QUERIES):
    from tests.helpers import synth

    t = time.time()
    X, y, qid = synth(n, d, q)
    if LONG_TAIL:
        # list lengths like a web collection's: log-normal around n / q, a tail of lists of more
        # than a thousand documents (MSLR-WEB30K: mean 120, maximum 1251)
        rng = np.random.default_rng(7)
        sigma = 0.8
        lens = rng.lognormal(np.log(n / q) - sigma * sigma / 2, sigma, size=q)
        lens = np.clip(np.rint(lens), 1, 1300).astype(np.int64)
        qid = np.repeat(np.arange(q, dtype=np.int64), lens)
        rest = max(n - len(qid), 0)  # what the draw left over goes into lists of the mean length
        qid = np.concatenate([qid, q + np.arange(rest, dtype=np.int64) // max(n // q, 1)])[:n]
    log("[bench] synthetic data %dx%d, %d queries in %.1fs" % (n, d, len(np.unique(qid)), time.time() - t))
    return X, y, qid


def line_search_candidates(orig: float, direction: int, step_base=0.05, step_scale=2.0, iters=T_ITERS):
    """coordinate_ascent.rs:145-171"""
    step = step_base * direction
    if orig != 0.0 and abs(step) > 0.5 * abs(orig):
        step = step_base * abs(orig) * direction
    total = step
    if direction == 0:
        iters, total = 1, -orig
    out = []
    for _ in range(iters):
        out.append(orig + total)
        step *= step_scale
        total += step
    return out


def step_inputs(step: int, d: int):
    """Deterministic base weights (L1-normalised as CA does) and the feature each restart probes."""
    rng = np.random.default_rng(1000 + step)
    base = rng.uniform(-1.0, 1.0, size=(N_RESTARTS, d))
    base /= np.abs(base).sum(axis=1, keepdims=True)
    fids = [int((step * N_RESTARTS + r * 17) % d) for r in range(N_RESTARTS)]
    group_a = [line_search_candidates(base[r, fids[r]], 0) + line_search_candidates(base[r, fids[r]], -1)
               for r in range(N_RESTARTS)]
    group_b = [line_search_candidates(base[r, fids[r]], +1) for r in range(N_RESTARTS)]
    return base, fids, group_a, group_b


EVALS_PER_STEP = N_RESTARTS * (1 + 2 * T_ITERS)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index
        self.path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % gpu_index)

    def start(self):
        try:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            self.fp = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.fp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fp.close()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "power_w_max": float(max(power)), "samples": len(sm)}
        return out


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ------------------------------------------------------------------------------------------
def cpu_evals_per_sec(X, y, qid, evals_per_thread: int, threads: int = N_RESTARTS):
    """`threads` workers (one per restart, as rayon does in the reference), each running
    `evals_per_thread` full evaluate_mean equivalents with the C oracle."""
    from oracle import oracle as orc
    from tests.helpers import oracle_dataset

    ods = oracle_dataset(orc, X, y, qid)
    d = X.shape[1]

    def work(r):
        rng = np.random.default_rng(77 + r)
        for _ in range(evals_per_thread):
            w = rng.uniform(-1, 1, size=d)
            per_query = orc.evaluate_scores(ods, orc.score_linear(X, w), "ndcg@%d" % DEPTH)
            orc.mean(per_query)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return threads * evals_per_thread / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    X, y, qid = make_data()
    threads = min(N_RESTARTS, os.cpu_count() or 1)
    for _ in range(args.warmup):
        cpu_evals_per_sec(X, y, qid, 1, threads)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        v, dt = cpu_evals_per_sec(X, y, qid, 1, threads)
        t_total += dt
        n_total += threads
    value = n_total / t_total
    sample = "%d steps x %d threads x 1 full evaluate_mean (1M docs) each" % (args.steps, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": base_config(world),
        "details": {"note": "C restatement of the Rust reference (oracle/), not the Rust build; one thread per "
                            "restart as rayon does (coordinate_ascent.rs:216)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "threads": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:  # this arm is pure CPU: the CUDA library must not even be mapped
        line["details"]["native_so_loaded"] = any("libfastrank_b200" in m for m in open("/proc/self/maps"))
    except OSError:
        pass
    print(json.dumps(line), file=JSON_OUT, flush=True)



# ------------------------------------------------------------------------------------------
# side measurements (single GPU, after the timed region): the other BASELINE.json configs
# ------------------------------------------------------------------------------------------
def random_tree(rng, X, depth, max_depth):
    if depth >= max_depth or (depth > 2 and rng.random() < 0.05):
        return {"LeafNode": float(np.round(rng.uniform(0, 4), 3))}
    fid = int(rng.integers(0, X.shape[1]))
    split = float(np.quantile(X[:2000, fid], rng.uniform(0.1, 0.9)))
    return {"FeatureSplit": {"fid": fid, "split": split, "lhs": random_tree(rng, X, depth + 1, max_depth),
                             "rhs": random_tree(rng, X, depth + 1, max_depth)}}


def count_nodes(t):
    if "LeafNode" in t:
        return 1
    return 1 + count_nodes(t["FeatureSplit"]["lhs"]) + count_nodes(t["FeatureSplit"]["rhs"])


def side_trees(fr, X, y, qid, n_trees=500, depth=8, reps=5, cpu_docs=20000):
    """BASELINE.json configs[3]: a 500-tree, depth-8 ensemble (model.rs:64-84, :104-112) scored over
    the bench's 1M x 136 through the reference-facing API (CModel.predict_dense / evaluate_mean),
    kernel time from CUDA events, the C oracle on a bounded sample of the same documents."""
    from oracle import oracle as orc

    rng = np.random.default_rng(4)
    members = [{"DecisionTree": random_tree(rng, X, 1, depth)} for _ in range(n_trees)]
    spec = {"Ensemble": {"weights": [1.0] * n_trees, "models": members}}
    nodes = sum(count_nodes(m["DecisionTree"]) for m in members)
    n = X.shape[0]
    ds = fr.CDataset.from_numpy(X, y, qid)
    model = fr.CModel.from_dict(spec)
    t0 = time.perf_counter()
    scores = model.predict_dense(ds)  # uploads the dataset, lowers and uploads the forest
    t_first = time.perf_counter() - t0
    ds.device_profile(enable=True, read=True)
    t0 = time.perf_counter()
    for _ in range(reps):
        scores = model.predict_dense(ds)
    t_predict = (time.perf_counter() - t0) / reps
    n_k, k_ms = ds.device_profile(read=True)
    t0 = time.perf_counter()
    for _ in range(reps):
        mean = ds.evaluate_mean(model, "ndcg@10")
    t_eval = (time.perf_counter() - t0) / reps
    ds.device_profile(enable=False, read=True)
    t0 = time.perf_counter()
    exp = orc.score_model(X[:cpu_docs], spec)
    t_cpu = time.perf_counter() - t0
    kernel_ms = k_ms / max(n_k, 1)
    algo = n * X.shape[1] * 4 + n * 8 + nodes * 16  # X + scores out + nodes
    visits = float(n) * n_trees * (depth - 1)
    peak = measured_peak_gbs()[0]
    return {"workload": "%d trees, depth <= %d (%d nodes), scored over %d x %d" % (n_trees, depth, nodes, n, X.shape[1]),
            "kernel": "forest_heap_kernel", "kernel_ms": kernel_ms, "predict_dense_api_ms": 1e3 * t_predict,
            "evaluate_mean_api_ms": 1e3 * t_eval, "first_call_ms": 1e3 * t_first,
            "docs_per_s_kernel": n / (kernel_ms / 1e3) if kernel_ms > 0 else None,
            "docs_per_s_api": n / t_predict, "node_visits_per_s": visits / (kernel_ms / 1e3) if kernel_ms > 0 else None,
            "algorithmic_bytes": algo, "achieved_GBps": algo / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else None,
            "frac_of_hbm_peak": algo / (kernel_ms / 1e3) / 1e9 / peak if kernel_ms > 0 else None,
            "ndcg10": mean, "cpu_oracle_docs_per_s_1thread": cpu_docs / t_cpu,
            "bit_exact_vs_oracle": bool(np.array_equal(scores[:cpu_docs], exp)), "bit_exact_sample_docs": cpu_docs}


def side_mslr(steps=5, warmup=3):
    """BASELINE.json configs[4]'s shape: 3 771 125 x 136 documents in 31 531 queries (~120 per
    query, MSLR-WEB30K-like), the same coordinate-ascent step (8 restarts x 51 candidates, NDCG@10)."""
    from fastrank_b200._native import ffi, lib
    from fastrank_b200.kernels import DevDataset, dense_query_index

    n, q = 3_771_125, 31_531
    X, y, qid = make_data(n, N_FEAT, q)
    qidx, nq = dense_query_index(qid)
    dev = DevDataset(X, y.astype(np.float32), qidx, nq)
    try:
        plan = dev.plan(0, DEPTH)
        packed = []
        for s in range(warmup + steps):
            base, fids, ga, gb = step_inputs(s, N_FEAT)
            packed.append(plan.pack_sweeps(base, fids, [a + b for a, b in zip(ga, gb)]))
        for s in range(warmup):
            plan.coord_sweeps_packed(packed[s])
        dev.profile(True)
        dev.profile_read(reset=True)
        dev.timer_start()
        for s in range(steps):
            plan.coord_sweeps_packed(packed[warmup + s])
        ms = dev.timer_stop()
        n_k, k_ms = dev.profile_read(reset=True)
        dev.profile(False)
        algo = n * N_FEAT * 4 + n * 5 + nq * 16 + N_RESTARTS * N_FEAT * 8 + EVALS_PER_STEP * 16
        kernel_ms = k_ms / max(n_k, 1)
        peak = measured_peak_gbs()[0]
        return {"workload": "synthetic %d docs x %d features x %d queries, coordinate_ascent 8 restarts, ndcg@10" % (n, N_FEAT, q),
                "kernel": ffi.string(lib.fr_dev_plan_sweep_kernel(plan.ptr)).decode(),
                "tile_documents": int(lib.fr_dev_plan_tile_documents(plan.ptr)),
                "untiled_queries": int(lib.fr_dev_plan_untiled_queries(plan.ptr)),
                "ms_per_step": ms / steps, "evals_per_s": EVALS_PER_STEP * steps / (ms / 1e3), "steps": steps,
                "kernel_ms_per_launch": kernel_ms, "launches_per_step": n_k / steps,
                "algorithmic_bytes_per_launch": algo, "achieved_GBps": algo / (kernel_ms / 1e3) / 1e9,
                "frac_of_hbm_peak": algo / (kernel_ms / 1e3) / 1e9 / peak,
                "bf16_gemm_variant": "not built: phase 1 (the 8-column X.W product a GEMM would replace) is 12.1 % of the "
                                     "stall samples / 13.2 % of the instructions of this launch (16.7 % / 15.8 % on the "
                                     "1M-document step), ncu source counters in profiles/r02_sweep_packed_*_hot_lines.txt; "
                                     "bf16 inputs would also break the f64 rounding contract (DESIGN.md 8)"}
    finally:
        dev.close()

# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import fastrank_b200 as fr
    from fastrank_b200 import dist as frdist
    from fastrank_b200._native import ffi, lib
    from fastrank_b200.kernels import DevDataset, dense_query_index

    if lib.fr_dev_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; fastrank_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = frdist.init_communicator(rank, world, local_rank, install_default=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    X, y, qid = make_data()
    if world > 1:
        rows = frdist.shard_rows(qid, rank, world)
        Xl, yl, ql = np.ascontiguousarray(X[rows]), np.ascontiguousarray(y[rows]), np.ascontiguousarray(qid[rows])
    else:
        Xl, yl, ql = X, y, qid
    n_local, d = Xl.shape
    qidx, nq_local = dense_query_index(ql)

    # ---- value: dataset resident in HBM ----------------------------------------------------
    dev = DevDataset(Xl, yl.astype(np.float32), qidx, nq_local, device=local_rank)
    plan = dev.plan(0, DEPTH)
    if comm is not None:
        if lib.fr_dev_plan_set_comm(plan.ptr, comm.ptr):
            raise RuntimeError("fr_dev_plan_set_comm failed")

    plan_layout = (int(lib.fr_dev_plan_tile_documents(plan.ptr)), int(lib.fr_dev_plan_untiled_queries(plan.ptr)))
    # the (tiny) inputs of every step are laid out before the clock starts: what is timed is the
    # C-ABI call -- staging of weights and candidates, the kernel, the all-reduce, the read-back
    # train_model submits direction +1 together with directions 0 / -1 (one launch, 8 x 51
    # candidates, one pass over X); --two-launches times the schedule without that speculation
    packed = {}
    for s in range(args.warmup + args.steps):
        base, fids, ga, gb = step_inputs(s, d)
        if args.two_launches or args.exact:
            packed[s] = (plan.pack_sweeps(base, fids, ga), plan.pack_sweeps(base, fids, gb))
        else:
            packed[s] = (plan.pack_sweeps(base, fids, [a + b for a, b in zip(ga, gb)]),)
    launches_per_step = len(packed[0])

    def one_step(s):
        for pk in packed[s]:
            plan.coord_sweeps_packed(pk, fast=not args.exact)

    for s in range(args.warmup):
        one_step(s)
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = int(lib.fr_dev_kernel_launches())
    dev.profile(True)
    dev.profile_read(reset=True)
    sampler.start()
    dev.timer_start()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        one_step(args.warmup + s)
    ms = dev.timer_stop()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    n_kern, kern_ms = dev.profile_read(reset=True)
    dev.profile(False)
    launches = int(lib.fr_dev_kernel_launches()) - launches0
    # side measurement, outside the timed region and reported next to the roofline: SURVEY 8(d)'s
    # first row, the full rescore (evaluate_mean for C arbitrary weight vectors per pass over X)
    full_rescore = []
    if world == 1:
        rng = np.random.default_rng(11)
        peak_gbs = measured_peak_gbs()[0]
        for c in (1, 8):
            W = rng.normal(size=(c, d))
            plan.eval_linear(W, per_query=False)
            dev.profile(True)
            dev.profile_read(reset=True)
            for _ in range(10):
                plan.eval_linear(W, per_query=False)
            n_l, l_ms = dev.profile_read(reset=True)
            dev.profile(False)
            algo = n_local * d * 4 + n_local * 4 + (nq_local + 1) * 4 + c * d * 8 + c * 8
            gbs = algo / (l_ms / max(n_l, 1) / 1e3) / 1e9
            full_rescore.append({"kernel": "linear_batch_kernel<%d,%d>" % (c if plan_layout[0] == 128 else max(c, 2), plan_layout[0]),
                                 "weight_vectors_per_pass": c, "avg_launch_ms": l_ms / max(n_l, 1),
                                 "evals_per_s": c / (l_ms / max(n_l, 1) / 1e3),
                                 "algorithmic_bytes_per_launch": algo, "achieved": gbs, "unit": "GB/s",
                                 "frac": gbs / peak_gbs})
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = EVALS_PER_STEP * args.steps / (ms / 1e3)

    # multi-GPU parity, visible to the driver: the sharded sums of the first timed step (already
    # all-reduced: identical on every rank) are kept here and compared, AFTER the e2e region, with
    # a single-GPU plan over rank 0's full copy
    sharded = None
    if world > 1 and not args.exact:
        sharded = plan.coord_sweeps_packed(packed[args.warmup][0]).copy()

    # roofline of the dominant kernel.  One launch of the batched sweep = ONE pass over this
    # rank's slice of the feature matrix serving all 8 restarts (204 or 200 candidates):
    # algorithmic bytes = X + per-document plan data (position 4 B, gain class 1 B) +
    # per-query plan data (16 B) + the transposed weight table + candidate rows.
    if args.exact:  # exact-order kernel: one pass over X per restart
        per_sweep_bytes = n_local * d * 4 + n_local * 4 + (nq_local + 1) * 4 + 26 * d * 8 + 26 * 8
        bytes_per_launch = N_RESTARTS * per_sweep_bytes
        kname = "coord_sweep_kernel<26,128>"
    else:
        bytes_per_launch = (n_local * d * 4 + n_local * 5 + nq_local * 16 + N_RESTARTS * d * 8
                            + EVALS_PER_STEP // launches_per_step * 16)
        kname = ffi.string(lib.fr_dev_plan_sweep_kernel(plan.ptr)).decode()
    avg_launch_ms = kern_ms / max(n_kern, 1)
    achieved = bytes_per_launch / (avg_launch_ms / 1e3) / 1e9 if avg_launch_ms > 0 else 0.0
    peak, peak_src = measured_peak_gbs()
    traffic = recorded_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                "kernel": kname, "launches_timed": n_kern,
                "avg_launch_ms": avg_launch_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_share_of_step": kern_ms / ms if ms > 0 else None, "peak_source": peak_src,
                "evals_per_launch": EVALS_PER_STEP / launches_per_step,
                # what actually bounds the kernel (from the committed ncu capture of this command)
                "issue_slots_busy_frac": None if traffic is None or "issue_active_pct" not in traffic
                else traffic["issue_active_pct"] / 100.0,
                "fp64_pipe_busy_frac": None if traffic is None or "fp64_pipe_pct" not in traffic
                else traffic["fp64_pipe_pct"] / 100.0,
                "note": "one X pass is shared by all 8 restarts (408 candidate rankings per document per launch), so "
                        "the kernel is FP64-compare bound, not HBM bound: ~240 M warp-level DSETP per launch at one "
                        "per ~2.35 clocks per SM sub-partition (tools/micro/cmp_throughput.cu) plus phase 1 = 0.71 ms "
                        "of FP64-pipe time; see DESIGN.md 3.0"}
    plan.close()
    dev.close()

    # ---- e2e: reference-facing C ABI with host buffers -------------------------------------
    barrier()
    t0 = time.perf_counter()
    ds = fr.CDataset.from_numpy(Xl, yl, ql)
    t_from_numpy = time.perf_counter() - t0
    req = fr.TrainRequest.coordinate_ascent()
    req.measure = "ndcg@%d" % DEPTH
    req.params.num_restarts = N_RESTARTS
    req.params.seed = 42
    req.params.quiet = True
    t1 = time.perf_counter()
    model = ds.train_model(req)
    t_train = time.perf_counter() - t1
    t1 = time.perf_counter()
    final = ds.evaluate_mean(model, req.measure)
    t_eval = time.perf_counter() - t1
    barrier()
    e2e_s = time.perf_counter() - t0
    stats = fr.query_json("last_train_stats")
    import hashlib

    weights = np.asarray(model.to_dict()["Linear"]["weights"], dtype=np.float64)
    weights_sha = hashlib.sha256(weights.tobytes()).hexdigest()
    weights_same = None
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        digests = [None] * world
        dist.all_gather_object(digests, weights_sha)
        weights_same = len(set(digests)) == 1
    cand_per_sweep = 1 + 2 * T_ITERS  # both direction groups ride in one submission
    h2d_total = Xl.nbytes + yl.nbytes + ql.nbytes + stats["sweeps"] * (d * 8 + cand_per_sweep * 16 + 4)
    d2h_total = stats["sweeps"] * cand_per_sweep * 8
    e2e = {"value": stats["evals_consumed"] / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": h2d_total / max(stats["global_steps"], 1),
           "d2h_bytes_per_step": d2h_total / max(stats["global_steps"], 1),
           "seconds": e2e_s, "evals_consumed": stats["evals_consumed"], "evals_computed": stats["evals_computed"],
           "global_steps": stats["global_steps"], "final_train_ndcg10": final,
           "seconds_from_numpy": t_from_numpy, "seconds_train_model": t_train, "seconds_evaluate": t_eval,
           "train_seconds_setup": stats.get("seconds_setup"), "train_seconds_device": stats.get("seconds_device"),
           "sweeps": stats["sweeps"], "trained_weights_sha256": weights_sha,
           "trained_weights_identical_on_all_ranks": weights_same,
           "what": "from_numpy + train_model(CA, 8 restarts, seed 42, to convergence) + evaluate through the C ABI"}
    del ds, model
    clocks = sampler.stop()  # sampled across both timed regions (device-timed steps and e2e)

    dist_equal = None
    if sharded is not None and rank == 0:
        qidx_full, nq_full = dense_query_index(qid)
        dev1 = DevDataset(X, y.astype(np.float32), qidx_full, nq_full, device=local_rank)
        try:
            plan1 = dev1.plan(0, DEPTH)
            base, fids, ga, gb = step_inputs(args.warmup, d)
            single = plan1.coord_sweeps_packed(plan1.pack_sweeps(base, fids, [a + b for a, b in zip(ga, gb)]))
            dist_equal = bool(np.array_equal(single, sharded))
        finally:
            dev1.close()
    if world > 1:
        barrier()  # rank 0 may still be checking; nobody tears the communicator down before that

    # ---- CPU baseline on the host cores (rank 0, single-GPU run only) -----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = min(N_RESTARTS, os.cpu_count() or 1)
        per_thread = args.cpu_evals_per_thread
        v, dt = cpu_evals_per_sec(X, y, qid, per_thread, threads)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "threads": threads, "kind": "port",
               "sample": "%d threads x %d full evaluate_mean (1M docs, ndcg@10) with the C oracle, %.1fs" % (threads, per_thread, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(world),
            "details": {"wall_ms_per_step": wall_ms / max(args.steps, 1), "tile_documents": plan_layout[0],
                        "untiled_queries": plan_layout[1]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        }
        if dist_equal is not None:
            line["dist_sums_equal_single_gpu"] = dist_equal
        if full_rescore:
            line["roofline_full_rescore"] = full_rescore
        if world == 1 and not args.no_side:
            try:
                line["side_trees"] = side_trees(fr, X, y, qid)
            except Exception as e:  # a side measurement must not cost the headline line
                line["side_trees"] = {"error": repr(e)}
            del X, Xl
            try:
                line["side_mslr"] = side_mslr()
            except Exception as e:
                line["side_mslr"] = {"error": repr(e)}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


def claim_stdout():
    """Libraries under us print to stdout (NCCL's version banner, for one).  The contract is ONE
    JSON line there, so fd 1 is pointed at stderr for the duration of the run and the saved
    descriptor is used for the line itself."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    global JSON_OUT
    JSON_OUT = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip side_trees / side_mslr (BASELINE configs[3], [4])")
    ap.add_argument("--two-launches", action="store_true",
                    help="submit direction +1 separately (no speculation): two launches per step")
    ap.add_argument("--exact", action="store_true", help="time the exact-order sweep kernel instead of the batched one")
    ap.add_argument("--cpu-evals-per-thread", type=int, default=60,
                    help="bounded CPU sample: evaluations per host thread (8 threads x 60 ~ 10 s)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
