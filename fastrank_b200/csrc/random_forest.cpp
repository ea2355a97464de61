// random_forest.cpp -- random-forest learner (random_forest.rs:14-408) on the host.
//
// Tree INDUCTION is host work (sort-by-feature and prefix statistics per node; SURVEY.md 2 row 5
// keeps it off the GPU path); what the forest produces -- a WeightedEnsemble of regression
// trees -- is scored and evaluated on the GPU (device.cu model_score_kernel + scores_eval_kernel):
// once per tree when the reference's per-tree evaluate_mean is observable (progress table,
// weight_trees; random_forest.rs:315-318) and for every later evaluate / predict call.
//
// Trees are induced on a thread per tree (the reference uses rayon, random_forest.rs:305); the
// per-tree seeds are drawn up front from the master generator exactly as the reference does, so
// the result does not depend on the number of threads.  Where the reference's outcome depends on
// unspecified order (sort_unstable among equal keys, HashMap iteration) this file picks the
// deterministic choice: stable sorts, "last maximal element", parent query order.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>

#include "host.hpp"

namespace frb {

namespace {

// stats.rs:53-124 (Welford; Knuth TAOCP vol 2, 3ed, p. 232)
struct StreamingStats {
    uint64_t n = 0;
    double mean = 0.0, s = 0.0, max = -1.7976931348623157e308, min = 1.7976931348623157e308, total = 0.0;
    void push(double x) {
        n += 1;
        const double old_mean = mean, old_s = s;
        if (max < x) max = x;
        if (min > x) min = x;
        total += x;
        if (n == 1) {
            mean = x;
            return;
        }
        mean = old_mean + (x - old_mean) / (double)n;
        s = old_s + (x - old_mean) * (x - mean);
    }
    bool finished() const { return n > 1; }  // finish() is None without a variance
    double variance() const { return s / (double)(n - 1); }
};

struct Ctx {
    const RandomForestParams &p;
    const ParentDataset &ds;
    // dataset.rs:306-308 / dense_dataset.rs:143-147: None when the instance does not carry fid
    bool feature_value(uint32_t inst, uint32_t fid, double *out) const {
        if (fid >= ds.d) return false;
        const float v = ds.x[(size_t)inst * ds.d + fid];
        if (!ds.dense_source) {
            const uint32_t len = ds.row_len[inst];
            // dense row: ids below its length are present; sparse row (len 0): the listed ones,
            // which the densified matrix can only tell apart from absent ones when non-zero
            if (len > 0 ? fid >= len : v == 0.0f) return false;
        }
        *out = (double)v;
        return true;
    }
    double gain(uint32_t inst) const { return (double)ds.gains[inst]; }
};

double compute_output(const Ctx &c, const uint32_t *ids, size_t n) {  // random_forest.rs:32-41
    if (n == 0) return 0.0;
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i) sum += c.gain(ids[i]);
    return sum / (double)n;
}

double squared_error(const Ctx &c, const uint32_t *ids, size_t n) {  // :42-51
    const double output = compute_output(c, ids, n);
    double sse = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double diff = output - c.gain(ids[i]);
        sse += diff * diff;
    }
    return sse;
}

double positive_fraction(const Ctx &c, const uint32_t *ids, size_t n, double *p_no) {
    size_t positive = 0;
    for (size_t i = 0; i < n; ++i) positive += c.ds.gains[ids[i]] > 0.0f;
    const double count = (double)n;
    *p_no = (count - (double)positive) / count;
    return (double)positive / count;
}

double gini_impurity(const Ctx &c, const uint32_t *ids, size_t n) {  // :52-66
    if (n == 0) return 0.0;
    double p_no;
    const double p_yes = positive_fraction(c, ids, n, &p_no);
    return p_yes * (1.0 - p_yes) + p_no * (1.0 - p_no);
}

double plogp(double x) { return x == 0.0 ? 0.0 : x * std::log2(x); }  // :67-73

double entropy(const Ctx &c, const uint32_t *ids, size_t n) {  // :74-87
    if (n == 0) return 0.0;
    double p_no;
    const double p_yes = positive_fraction(c, ids, n, &p_no);
    return -plogp(p_yes) - plogp(p_no);
}

double label_variance(const Ctx &c, const uint32_t *ids, size_t n) {  // label_stats(..).unwrap().variance
    StreamingStats st;
    for (size_t i = 0; i < n; ++i) st.push(c.gain(ids[i]));
    if (!st.finished()) throw Error("TrueVarianceReduction needs at least two instances on each side of a split");
    return st.variance();
}

double importance(const Ctx &c, const uint32_t *lhs, size_t nl, const uint32_t *rhs, size_t nr) {  // :90-125
    const std::string &m = c.p.split_method;
    if (m == "SquaredError") return -(squared_error(c, lhs, nl) + squared_error(c, rhs, nr));
    if (m == "BinaryGiniImpurity") return -(gini_impurity(c, lhs, nl) * (double)nl + gini_impurity(c, rhs, nr) * (double)nr);
    if (m == "InformationGain") return -(entropy(c, lhs, nl) * (double)nl + entropy(c, rhs, nr) * (double)nr);
    return -(label_variance(c, lhs, nl) * (double)nl + label_variance(c, rhs, nr) * (double)nr);
}

struct FeatureSplit {
    bool valid = false;
    uint32_t fid = 0;
    double split = 0.0, importance = 0.0;
    std::vector<uint32_t> lhs, rhs;
};

// random_forest.rs:211-286
FeatureSplit generate_split_candidate(const Ctx &c, uint32_t fid, const std::vector<uint32_t> &instances,
                                      const StreamingStats &fstats) {
    FeatureSplit out;
    StreamingStats labels;
    for (uint32_t i : instances) labels.push(c.gain(i));
    if (!labels.finished() || labels.max == labels.min) return out;
    const uint32_t k = c.p.split_candidates;
    const double range = fstats.max - fstats.min;
    std::vector<std::pair<double, uint32_t>> by_value;
    by_value.reserve(instances.size());
    for (uint32_t i : instances) {
        double v = 0.0;
        c.feature_value(i, fid, &v);  // unwrap_or(0.0)
        by_value.emplace_back(v, i);
    }
    std::stable_sort(by_value.begin(), by_value.end(),
                     [](const auto &a, const auto &b) { return a.first < b.first; });
    std::vector<uint32_t> ids(by_value.size());
    for (size_t i = 0; i < by_value.size(); ++i) ids[i] = by_value[i].second;
    // positions in the sorted order where each of the k-1 evenly spaced thresholds falls
    std::vector<std::pair<double, size_t>> positions;
    size_t at = 0;
    for (uint32_t i = 1; i < k; ++i) {
        const double f = (double)i / (double)k;
        const double position = f * range + fstats.min;
        while (at < ids.size() && by_value[at].first < position) ++at;
        if (!positions.empty() && positions.back().second == at) continue;
        positions.emplace_back(position, at);
    }
    bool have = false;
    size_t best_pos = 0;
    for (const auto &sp : positions) {
        const size_t right = sp.second;
        const size_t nl = right, nr = ids.size() - right;
        if (nl < c.p.min_leaf_support || nr < c.p.min_leaf_support) continue;
        const double imp = importance(c, ids.data(), nl, ids.data() + right, nr);
        if (imp != imp) throw Error("split importance is NaN");
        if (!have || imp >= out.importance) {  // sort by importance, take the last
            have = true;
            out.importance = imp;
            out.split = sp.first;
            best_pos = right;
        }
    }
    if (!have) return out;
    out.valid = true;
    out.fid = fid;
    out.lhs.assign(ids.begin(), ids.begin() + best_pos);
    out.rhs.assign(ids.begin() + best_pos, ids.end());
    return out;
}

std::unique_ptr<TreeNode> leaf(double value) {
    std::unique_ptr<TreeNode> n(new TreeNode());
    n->leaf = true;
    n->value = value;
    return n;
}

// random_forest.rs:362-408; nullptr plays Err(NoTreeReason)
std::unique_ptr<TreeNode> learn_recursive(const Ctx &c, const std::vector<uint32_t> &features,
                                          const std::vector<uint32_t> &instances, uint32_t depth) {
    if (features.empty() || instances.empty()) return nullptr;       // StepDone
    if (depth >= c.p.max_depth) return nullptr;                      // DepthExceeded
    if (instances.size() < (size_t)c.p.min_leaf_support) return nullptr;  // SplitTooSmall
    // FeatureStats::compute (normalizers.rs:13-36): missing values are skipped
    std::vector<StreamingStats> fstats(features.size());
    for (uint32_t inst : instances) {
        for (size_t a = 0; a < features.size(); ++a) {
            double v;
            if (c.feature_value(inst, features[a], &v)) fstats[a].push(v);
        }
    }
    FeatureSplit best;
    for (size_t a = 0; a < features.size(); ++a) {
        if (!fstats[a].finished()) continue;
        FeatureSplit cand = generate_split_candidate(c, features[a], instances, fstats[a]);
        if (!cand.valid) continue;
        if (!best.valid || cand.importance >= best.importance) best = std::move(cand);
    }
    if (!best.valid) return nullptr;  // NoFeatureSplitCandidates
    std::unique_ptr<TreeNode> lhs = learn_recursive(c, features, best.lhs, depth + 1);
    if (!lhs) lhs = leaf(compute_output(c, best.lhs.data(), best.lhs.size()));
    std::unique_ptr<TreeNode> rhs = learn_recursive(c, features, best.rhs, depth + 1);
    if (!rhs) rhs = leaf(compute_output(c, best.rhs.data(), best.rhs.size()));
    std::unique_ptr<TreeNode> node(new TreeNode());
    node->leaf = false;
    node->fid = best.fid;
    node->split = best.split;
    node->lhs = std::move(lhs);
    node->rhs = std::move(rhs);
    return node;
}

uint32_t tree_depth(const TreeNode &n) {  // random_forest.rs:159-166
    return n.leaf ? 1u : 1u + std::max(tree_depth(*n.lhs), tree_depth(*n.rhs));
}

template <typename T>
std::vector<T> sample_without_replacement(std::vector<T> data, Rand64 &rng, size_t count) {  // randutil.rs:14-18
    shuffle(data, rng);
    if (data.size() > count) data.resize(count);
    return data;
}

}  // namespace

Model random_forest_learn(const RandomForestParams &p, const DatasetView &view, const Evaluator &ev,
                          TrainStats *stats_out) {
    TrainStats stats;
    if (p.num_trees == 0) throw Error("Should be at least 1 tree!");
    if (p.split_candidates < 2) throw Error("split_candidates must be at least 2");
    const ParentDataset &parent = *view.parent;
    const Ctx ctx{p, parent};
    Rand64 master((unsigned __int128)p.seed);
    std::vector<uint64_t> seeds(p.num_trees);
    for (uint32_t i = 0; i < p.num_trees; ++i) seeds[i] = master.rand_u64();  // random_forest.rs:293-296

    // the view's features and queries, sorted as sampling.rs:43-44 does before sampling
    std::vector<uint32_t> all_features = view.feature_ids();
    std::sort(all_features.begin(), all_features.end());
    const auto groups = view.instances_by_query();
    std::vector<size_t> query_order(groups.size());
    for (size_t i = 0; i < groups.size(); ++i) query_order[i] = i;
    std::sort(query_order.begin(), query_order.end(), [&](size_t a, size_t b) {
        return parent.query_names[groups[a].first] < parent.query_names[groups[b].first];
    });
    if (all_features.empty()) throw Error("dataset has no features");

    std::vector<std::unique_ptr<TreeNode>> trees(p.num_trees);
    std::vector<std::string> failures(p.num_trees);
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t idx = next.fetch_add(1);
            if (idx >= p.num_trees) return;
            try {
                Rand64 rng((unsigned __int128)seeds[idx]);
                // sampling.rs:38-66
                const size_t n_features =
                    std::max<size_t>(1, (size_t)((double)all_features.size() * p.feature_sampling_rate));
                const size_t n_queries =
                    std::max<size_t>(1, (size_t)((double)query_order.size() * p.instance_sampling_rate));
                const std::vector<uint32_t> features = sample_without_replacement(all_features, rng, n_features);
                const std::vector<size_t> chosen = sample_without_replacement(query_order, rng, n_queries);
                std::vector<uint8_t> keep(groups.size(), 0);
                for (size_t g : chosen) keep[g] = 1;
                std::vector<uint32_t> instances;
                for (size_t g = 0; g < groups.size(); ++g)
                    if (keep[g]) instances.insert(instances.end(), groups[g].second.begin(), groups[g].second.end());
                std::unique_ptr<TreeNode> root = learn_recursive(ctx, features, instances, 1);
                if (!root) root = leaf(compute_output(ctx, instances.data(), instances.size()));  // :344-352
                trees[idx] = std::move(root);
            } catch (const std::exception &e) {
                failures[idx] = e.what();
            }
        }
    };
    {
        unsigned hw = std::thread::hardware_concurrency();
        const unsigned nthreads = std::max(1u, std::min<unsigned>(hw ? hw : 1u, p.num_trees));
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker);
        worker();
        for (std::thread &t : pool) t.join();
    }
    for (const std::string &f : failures)
        if (!f.empty()) throw Error(f);

    if (!p.quiet) {
        printf("-----------------------\n|%7s|%7s|%7s|\n-----------------------\n", "Tree", "Depth",
               ev.measure().display.c_str());
    }
    Model out;
    out.kind = Model::Ensemble;
    const bool need_eval = !p.quiet || p.weight_trees;  // the only uses of the per-tree score (:315-318, :330-335)
    for (uint32_t i = 0; i < p.num_trees; ++i) {
        const uint32_t depth = tree_depth(*trees[i]);
        Model m;
        m.kind = Model::DecisionTree;
        m.tree = std::move(trees[i]);
        double score = 1.0;
        if (need_eval) {
            score = ev.evaluate_mean(m);  // hot loops (c) + (b) on the GPU, once per tree
            stats.evals_consumed += 1;
            stats.evals_computed += 1;
            if (!p.quiet) printf("|%7u|%7u|%7.3f|\n", i + 1, depth, score);
        }
        out.weights.push_back(p.weight_trees ? score : 1.0);
        out.members.push_back(std::move(m));
    }
    if (!p.quiet) printf("-----------------------\n");
    if (stats_out) *stats_out = stats;
    set_last_train_stats(stats);
    return out;
}

}  // namespace frb
