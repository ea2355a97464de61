"""Drop-in name: with `compat/` on PYTHONPATH, `import fastrank` resolves to fastrank_b200, so
code written against the reference package (reference fastrank/__init__.py:2-8: CQRel,
CDataset, CModel, query_json, TrainRequest, the params dataclasses, the `clib` and `training`
sub-modules) runs on the GPU implementation unchanged."""
import sys as _sys

import fastrank_b200 as _impl
from fastrank_b200 import clib, training  # noqa: F401
from fastrank_b200 import (CDataset, CModel, CoordinateAscentParams, CQRel, RandomForestParams,  # noqa: F401
                           TrainRequest, query_json)

VERSION_TUPLE = _impl.VERSION_TUPLE
__version__ = _impl.__version__

# `from fastrank.clib import ...` / `from fastrank.training import ...`
_sys.modules[__name__ + ".clib"] = clib
_sys.modules[__name__ + ".training"] = training
