// random_forest.cpp -- random-forest learner (random_forest.rs:288-408).
#include "host.hpp"

namespace frb {

Model random_forest_learn(const RandomForestParams &, const DatasetView &, const Evaluator &, TrainStats *) {
    throw Error("RandomForest training is not available in this build yet (tree-ensemble SCORING is)");
}

}  // namespace frb
