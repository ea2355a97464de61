import os, sys
sys.path.insert(0, os.getcwd())
import fastrank_b200 as fr
from tests.helpers import synth
X, y, qid = synth(1_000_000, 136, 30000)
ds = fr.CDataset.from_numpy(X, y, qid)
req = fr.TrainRequest.random_forest(); req.measure = "ndcg@10"
req.params.num_trees, req.params.quiet, req.params.seed = 4, True, 5
os.environ["FASTRANK_RF"] = "gpu"
ds.train_model(req)
