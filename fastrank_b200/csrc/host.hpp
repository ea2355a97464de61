// host.hpp -- host-side objects behind the reference-compatible C ABI.
//
// The host owns parsing, bookkeeping and the coordinate-ascent control flow; every score,
// rank and metric is produced by the kernels in device.cu through the fr_dev_* ABI.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fastrank_b200.h"
#include "json.hpp"

namespace frb {

// Errors that become {"error":"error","context":...} at the boundary (ffi.rs:45-74).
class Error : public std::runtime_error {
   public:
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};

// ----------------------------------------------------------------------------------------
// oorandom::Rand64 (PCG XSL-RR 128/64).  The reference pins oorandom =11.1.0
// (Cargo.toml:18-19); the crate is not vendored, this restates its published algorithm.
// ----------------------------------------------------------------------------------------
class Rand64 {
   public:
    explicit Rand64(unsigned __int128 seed);
    uint64_t rand_u64();
    double rand_float();
    uint64_t rand_range(uint64_t start, uint64_t end);

   private:
    unsigned __int128 state_, inc_;
};

template <typename T>
void shuffle(std::vector<T> &v, Rand64 &rng) {  // randutil.rs:21-27
    const uint64_t n = v.size();
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t j = rng.rand_range(i, n);
        std::swap(v[i], v[j]);
    }
}

// ----------------------------------------------------------------------------------------
// Judgments (qrel.rs)
// ----------------------------------------------------------------------------------------
struct QueryJudgments {
    std::vector<std::pair<std::string, float>> docs;  // insertion order, unique doc ids
    std::unordered_map<std::string, size_t> index;
    void insert(const std::string &doc, float gain);
    uint32_t num_relevant() const;             // qrel.rs:21-26
    std::vector<float> gain_vector() const;    // qrel.rs:33-39 (positive gains only)
};

struct QRel {
    std::vector<std::string> order;
    std::unordered_map<std::string, QueryJudgments> queries;
    const QueryJudgments *get(const std::string &qid) const;
    QueryJudgments &get_or_create(const std::string &qid);
    static std::shared_ptr<QRel> load_file(const std::string &path);    // qrel.rs:65-102
    static std::shared_ptr<QRel> from_json(const json::Value &v);       // qrel.rs:42-46
    json::Value to_json() const;
};

// ----------------------------------------------------------------------------------------
// Models (model.rs)
// ----------------------------------------------------------------------------------------
struct TreeNode {
    bool leaf = true;
    double value = 0.0;  // leaf
    uint32_t fid = 0;
    double split = 0.0;
    std::unique_ptr<TreeNode> lhs, rhs;
};

struct Model {
    enum Kind { SingleFeature, Linear, DecisionTree, Ensemble };
    Kind kind = Linear;
    uint32_t fid = 0;               // SingleFeature
    double dir = 1.0;               // SingleFeature
    std::vector<double> weights;    // Linear: per feature; Ensemble: per member
    std::unique_ptr<TreeNode> tree; // DecisionTree
    std::vector<Model> members;     // Ensemble

    static Model from_json(const json::Value &v);  // serde layout of model.rs:10-16
    json::Value to_json() const;
    // Lowers to the device program of model_program.hpp.
    std::vector<uint64_t> lower() const;
    static Model linear(std::vector<double> w) {
        Model m;
        m.kind = Linear;
        m.weights = std::move(w);
        return m;
    }
};

// ----------------------------------------------------------------------------------------
// Datasets (dense_dataset.rs, dataset.rs, instance.rs, libsvm.rs, sampling.rs)
// ----------------------------------------------------------------------------------------
// A device plan kept by its dataset so that repeated evaluations of the same (view, measure)
// -- train_model followed by evaluate, evaluate in a loop -- do not rebuild it.
struct PlanHolder {
    fr_dev_plan *plan = nullptr;
    int metric = 0;
    int64_t depth = -1;
    bool sampled = false;
    std::vector<uint32_t> instances;  // the view's instances when sampled
    fr_dev_comm *comm = nullptr;
    uint64_t comm_generation = 0;  // fr_dev_comm_generation(comm) when the plan was built
    ~PlanHolder() {
        if (plan) fr_dev_plan_destroy(plan);
    }
};

struct ParentDataset {
    size_t n = 0;
    size_t d = 0;                 // n_dim: row length of the dense matrix
    const float *x = nullptr;     // row-major n*d; borrowed (from_numpy) or owned_x
    std::vector<float> owned_x;
    std::vector<float> gains;     // f32 label per instance
    std::vector<uint32_t> query_of;            // dense query number per instance
    std::vector<std::string> query_names;      // in order of first appearance
    std::unordered_map<std::string, uint32_t> query_lookup;
    std::vector<std::vector<uint32_t>> by_query;
    std::vector<std::string> docids;           // empty when the dataset has none
    std::vector<uint8_t> has_docid;
    std::vector<uint32_t> row_len;             // loaded data: features present up to this index (0: sparse row)
    // Sparse32 rows (instance.rs:104-122): the ids the row lists, ascending -- Some(0.0) and
    // None stay distinct (normalizers.rs:21-27 skips only what the row does not carry)
    std::unordered_map<uint32_t, std::vector<uint32_t>> sparse_ids;
    std::vector<uint32_t> features;            // ascending ids of the features present
    std::map<uint32_t, std::string> feature_names;
    bool dense_source = false;

    std::mutex dev_mu;
    fr_dev_dataset *dev = nullptr;             // uploaded on first use
    std::recursive_mutex use_mu;               // one evaluator at a time drives the device state
    std::vector<std::shared_ptr<PlanHolder>> plan_cache;  // plans without judgments, newest last
    // models lowered and uploaded for this dataset's device, keyed by CModel::uid, newest last:
    // predict / evaluate in a loop with the same model do not lower, upload and lay out a forest
    // again (500 trees: ~100 ms of host work for a 3 ms kernel).  Guarded by use_mu.
    std::vector<std::pair<uint64_t, fr_dev_model *>> model_cache;

    ~ParentDataset();
    fr_dev_dataset *device();                  // throws Error when no GPU is usable
    // the device form of `m`; uid 0 = a temporary model, built and handed to the caller (*owned)
    fr_dev_model *device_model(const Model &m, uint64_t uid, bool *owned);
    std::string feature_name(uint32_t fid) const;
};

struct DatasetView {
    std::shared_ptr<ParentDataset> parent;
    bool sampled = false;
    std::vector<uint32_t> instances;  // meaningful when sampled
    std::vector<uint32_t> features;   // meaningful when sampled

    std::vector<uint32_t> feature_ids() const;
    uint32_t n_dim() const;           // dataset.rs:124-126: a sample reports features.len()
    size_t num_instances() const;
    // (query number in the parent, instance ids) for every query of this view, parent order
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> instances_by_query() const;
    DatasetView with_queries(const std::vector<std::string> &queries) const;  // sampling.rs:101-115
    DatasetView with_features(const std::vector<uint32_t> &fids) const;       // sampling.rs:76-99
    DatasetView with_instances(std::vector<uint32_t> ids) const;              // sampling.rs:67-73
};

DatasetView load_ranksvm(const std::string &path, const std::string *feature_names_path);
DatasetView make_dense(size_t n, size_t d, const float *x, const double *y, const int64_t *qids);

// io_helper.rs:18-48: whole files, transparently (de)compressed when the name ends in .gz / .bz2 / .zst
std::string read_file_by_extension(const std::string &path);
void write_file_by_extension(const std::string &path, const std::string &content);

// ----------------------------------------------------------------------------------------
// Evaluator (evaluators.rs:98-224): measure parsing + a device plan for one view
// ----------------------------------------------------------------------------------------
struct Measure {
    int metric = FR_METRIC_NDCG;
    int64_t depth = -1;
    std::string display;  // evaluators.rs:343-349 etc.
    static Measure parse(const std::string &name);  // evaluators.rs:132-155
};

class Evaluator {
   public:
    Evaluator(const DatasetView &view, const Measure &measure, const QRel *qrel);
    ~Evaluator();
    Evaluator(const Evaluator &) = delete;
    Evaluator &operator=(const Evaluator &) = delete;

    const std::vector<uint32_t> &view_queries() const { return view_queries_; }
    size_t num_queries() const { return view_queries_.size(); }
    uint64_t global_queries() const { return fr_dev_plan_global_queries(plan_); }
    fr_dev_plan *plan() const { return plan_; }
    const Measure &measure() const { return measure_; }
    const DatasetView &view() const { return view_; }

    double mean_from_fx(int64_t fx) const;
    // evaluate_mean / evaluate_to_map for any model
    // (model_uid != 0: the model belongs to a CModel handle, its device form is cached)
    double evaluate_mean(const Model &m, std::vector<double> *per_query = nullptr, uint64_t model_uid = 0) const;
    // C weight vectors at once
    std::vector<double> evaluate_linear(const std::vector<std::vector<double>> &ws) const;

   private:
    DatasetView view_;
    Measure measure_;
    std::vector<uint32_t> view_queries_;
    std::unique_lock<std::recursive_mutex> use_lock_;
    std::shared_ptr<PlanHolder> holder_;
    fr_dev_plan *plan_ = nullptr;
};

// ----------------------------------------------------------------------------------------
// Training (json_api.rs, coordinate_ascent.rs, random_forest.rs)
// ----------------------------------------------------------------------------------------
struct CoordinateAscentParams {
    uint32_t num_restarts = 5;
    uint32_t num_max_iterations = 25;
    double step_base = 0.05;
    double step_scale = 2.0;
    double tolerance = 0.001;
    uint64_t seed = 0;
    bool normalize = true;
    bool quiet = false;
    bool init_random = true;
    bool output_ensemble = false;
    // not one of the reference's parameters (never serialised): TrainRequest's optional top-level
    // "sweep": "exact" asks for the exact-order kernel (the reference's summation order, bit for bit)
    // instead of the batched sweep
    bool exact_sweep = false;
    static CoordinateAscentParams defaults();  // coordinate_ascent.rs:25-41
    static CoordinateAscentParams from_json(const json::Value &v);
    json::Value to_json() const;
};

struct RandomForestParams {
    uint64_t seed = 0;
    bool quiet = false;
    uint32_t num_trees = 100;
    bool weight_trees = false;
    std::string split_method = "SquaredError";
    double instance_sampling_rate = 0.5;
    double feature_sampling_rate = 0.25;
    uint32_t min_leaf_support = 10;
    uint32_t split_candidates = 3;
    uint32_t max_depth = 8;
    static RandomForestParams defaults();  // random_forest.rs:141-157
    static RandomForestParams from_json(const json::Value &v);
    json::Value to_json() const;
};

struct TrainStats {
    uint64_t evals_consumed = 0;  // evaluate_mean calls the reference's control flow makes
    uint64_t evals_computed = 0;  // candidates actually scored on the GPU (speculation included)
    uint64_t sweeps = 0;
    uint64_t global_steps = 0;
    bool exact_sweep = false;     // the line searches ran on the exact-order kernel
    double seconds_setup = 0.0;   // evaluator construction: dataset upload (first use) + plan
    double seconds_device = 0.0;  // inside the fr_dev_* evaluation calls (copies, kernels, sync)
    double seconds_total = 0.0;   // the whole train_model call
};

Model coordinate_ascent_learn(const CoordinateAscentParams &p, const DatasetView &view,
                              const Evaluator &ev, TrainStats *stats);
Model random_forest_learn(const RandomForestParams &p, const DatasetView &view, const Evaluator &ev,
                          TrainStats *stats);

// last training statistics, readable through query_json("last_train_stats")
TrainStats last_train_stats();
void set_last_train_stats(const TrainStats &s);

}  // namespace frb

// The opaque handles of the C ABI.
struct CDataset {
    frb::DatasetView view;
};
struct CModel {
    frb::Model model;
    uint64_t uid;  // identity of this (immutable) model for per-dataset device caches
    explicit CModel(frb::Model m);
};
struct CQRel {
    std::shared_ptr<frb::QRel> qrel;
};
