import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
import fastrank_b200 as fr
X, y, qid = bench.make_data()
for rep in range(2):
    t0 = time.perf_counter()
    ds = fr.CDataset.from_numpy(X, y, qid)
    t1 = time.perf_counter()
    req = fr.TrainRequest.coordinate_ascent(); req.measure = "ndcg@10"
    req.params.num_restarts = 8; req.params.seed = 42; req.params.quiet = True; req.params.num_max_iterations = 1
    m = ds.train_model(req)
    t2 = time.perf_counter()
    print("from_numpy %.1f ms train(1 iter) %.1f ms stats %s" % (1e3*(t1-t0), 1e3*(t2-t1), fr.query_json("last_train_stats")), file=sys.stderr)
    t3 = time.perf_counter(); v = ds.evaluate_mean(m, "ndcg@10"); t4 = time.perf_counter()
    print("evaluate_mean %.1f ms" % (1e3*(t4-t3)), file=sys.stderr)
    del ds, m
