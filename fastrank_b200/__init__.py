"""fastrank_b200 -- a B200-native drop-in for the training / scoring / evaluation path of
jjfiv/fastrank.  Same Python surface as the reference package (`fastrank/__init__.py:2-8`)."""
from .clib import CDataset, CModel, CQRel, query_json
from .training import CoordinateAscentParams, RandomForestParams, TrainRequest

VERSION_TUPLE = (0, 7, 0)  # API level of the reference surface this mirrors
__version__ = "{}.{}.{}".format(*VERSION_TUPLE)

__all__ = ["clib", "training", "CQRel", "CDataset", "CModel", "query_json", "TrainRequest",
           "CoordinateAscentParams", "RandomForestParams"]
