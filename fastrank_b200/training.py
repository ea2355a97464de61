"""Training requests -- the dataclasses of the reference's fastrank/training.py (:7-134) with
the same fields, defaults and JSON layout (json_api.rs:13-34)."""
from __future__ import annotations

import random
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, Optional, Union

from .clib import CQRel, query_json


@dataclass
class CoordinateAscentParams:
    """coordinate_ascent.rs:11-23"""

    num_restarts: int = 5
    num_max_iterations: int = 25
    step_base: float = 0.05
    step_scale: float = 2.0
    tolerance: float = 0.001
    normalize: bool = True
    init_random: bool = True
    output_ensemble: bool = False
    seed: int = random.randint(0, (1 << 64) - 1)
    quiet: bool = False

    def name(self) -> str:
        return "CoordinateAscent"

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @staticmethod
    def from_dict(params) -> "CoordinateAscentParams":
        return CoordinateAscentParams(**params)


def _split_method_name(value) -> str:
    # the native side answers {"SquaredError": []} (serde's tuple-variant form)
    if isinstance(value, dict):
        [name] = list(value.keys())
        return name
    return value


@dataclass
class RandomForestParams:
    """random_forest.rs:127-139"""

    num_trees: int = 100
    weight_trees: bool = True
    split_method: Any = "SquaredError"
    instance_sampling_rate: float = 0.5
    feature_sampling_rate: float = 0.25
    min_leaf_support: int = 10
    split_candidates: int = 3
    max_depth: int = 8
    seed: int = random.randint(0, (1 << 64) - 1)
    quiet: bool = False

    def name(self) -> str:
        return "RandomForest"

    def to_dict(self) -> Dict[str, Any]:
        out = asdict(self)
        out["split_method"] = {_split_method_name(self.split_method): []}
        return out

    @staticmethod
    def from_dict(params) -> "RandomForestParams":
        return RandomForestParams(**params)


@dataclass
class TrainRequest:
    """What to optimise (measure), with which learner (params), against which judgments."""

    measure: str = "ndcg"
    params: Union[CoordinateAscentParams, RandomForestParams] = field(default_factory=CoordinateAscentParams)
    judgments: Optional[CQRel] = None

    def to_dict(self) -> Dict[str, Any]:
        return {
            "measure": self.measure,
            "params": {self.params.name(): self.params.to_dict()},
            "judgments": None if self.judgments is None else self.judgments.to_dict(),
        }

    def clone(self) -> "TrainRequest":
        return TrainRequest.from_dict(self.to_dict())

    @staticmethod
    def coordinate_ascent() -> "TrainRequest":
        return TrainRequest.from_dict(query_json("coordinate_ascent_defaults"))

    @staticmethod
    def random_forest() -> "TrainRequest":
        return TrainRequest.from_dict(query_json("random_forest_defaults"))

    @staticmethod
    def from_dict(params) -> "TrainRequest":
        judgments = None
        if params["judgments"] is not None:
            judgments = CQRel.from_dict(params["judgments"])
        learner = params["params"]
        if len(learner) != 1:
            raise ValueError("What do I do with this?: {}".format(learner))
        if "RandomForest" in learner:
            model_params = RandomForestParams.from_dict(learner["RandomForest"])
        elif "CoordinateAscent" in learner:
            model_params = CoordinateAscentParams.from_dict(learner["CoordinateAscent"])
        else:
            raise ValueError("Python doesn't know about model-params: {}".format(learner))
        return TrainRequest(params["measure"], model_params, judgments)
