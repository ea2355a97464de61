// random_forest.cpp -- random-forest learner (random_forest.rs:14-408).
//
// The host owns sampling (sampling.rs:38-66) and every decision of the induction: stop rules,
// which thresholds survive, importance, which candidate wins.  The statistics those decisions
// need come from one of two places:
//   * the GPU (rf_induction.cu): a tree grows level by level and the device
//     produces min / max and bucketed label sums for all open nodes at once -- the default for
//     dense datasets of some size with labels that are exact in its integer unit;
//   * counting passes on host threads (learn_recursive below), e.g. for libsvm data with missing
//     features or tiny datasets.
// Either way the k-1 evenly spaced thresholds of a feature are scored from counts and sums, not
// from the reference's sort: a cut at threshold p puts exactly the instances with value < p on
// the left, so the partitions are the reference's.  The forest is scored and evaluated on the
// GPU (trees.cu, device.cu): once per tree when the reference's per-tree evaluate_mean is
// observable (progress table, weight_trees; random_forest.rs:315-318) and for every later
// evaluate / predict call.
//
// The per-tree seeds are drawn up front from the master generator exactly as the reference does
// (random_forest.rs:293-296), so the result does not depend on the number of threads.  Where the
// reference's outcome depends on unspecified order (sort_unstable among equal keys, HashMap
// iteration) this file picks the deterministic choice: node order for sums, "last maximal
// element", parent query order.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <mutex>
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
            // dense row: ids below its length are present; sparse row (len 0): the listed ones
            if (len > 0) {
                if (fid >= len) return false;
            } else {
                auto it = ds.sparse_ids.find(inst);
                if (it == ds.sparse_ids.end() || !std::binary_search(it->second.begin(), it->second.end(), fid))
                    return false;
            }
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

double plogp(double x) { return x == 0.0 ? 0.0 : x * std::log2(x); }  // :67-73

struct FeatureSplit {
    bool valid = false;
    uint32_t fid = 0;
    double split = 0.0, importance = 0.0;
    size_t n_lhs = 0;
};

// random_forest.rs:211-286.  The reference sorts the node's instances by the feature and cuts
// the sorted list at each of the k-1 evenly spaced thresholds; a cut at threshold p puts exactly
// the instances with value < p on the left (the `while scores[i] < position` walk, :238-241).
// The same partitions are produced here without the sort: one pass counts, per threshold, how
// many values fall below it (that count is the reference's split position, so the "same
// position as the previous threshold" de-duplication is identical), and the importance of each
// surviving cut is accumulated over the instances in node order.  Only the order in which
// floating-point sums are taken differs -- an order the reference leaves unspecified anyway
// (sort_unstable among equal feature values).
FeatureSplit generate_split_candidate(const Ctx &c, uint32_t fid, const std::vector<uint32_t> &instances,
                                      const StreamingStats &fstats, const std::vector<double> &values) {
    FeatureSplit out;
    const uint32_t k = c.p.split_candidates;
    const double range = fstats.max - fstats.min;
    const size_t n = instances.size();
    std::vector<double> position;
    for (uint32_t i = 1; i < k; ++i) position.push_back(((double)i / (double)k) * range + fstats.min);
    const size_t m = position.size();
    // below[i] = #{v < position[i]}; thresholds ascend, so count into the first bucket that
    // holds the value and prefix-sum afterwards
    std::vector<size_t> bucket(m + 1, 0);
    for (size_t a = 0; a < n; ++a) {
        const double v = values[a];
        size_t b = 0;
        while (b < m && !(v < position[b])) ++b;
        bucket[b] += 1;
    }
    std::vector<size_t> below(m, 0);
    {
        size_t run = 0;
        for (size_t i = 0; i < m; ++i) {
            run += bucket[i];
            below[i] = run;
        }
    }
    std::vector<size_t> kept;  // thresholds that survive de-duplication and the leaf-size test
    {
        bool have_prev = false;
        size_t prev = 0;
        for (size_t i = 0; i < m; ++i) {
            if (have_prev && prev == below[i]) continue;
            have_prev = true;
            prev = below[i];
            const size_t nl = below[i], nr = n - below[i];
            if (nl < c.p.min_leaf_support || nr < c.p.min_leaf_support) continue;
            kept.push_back(i);
        }
    }
    if (kept.empty()) return out;
    const size_t q = kept.size();
    const std::string &method = c.p.split_method;
    std::vector<double> imp(q, 0.0);
    if (method == "SquaredError") {  // :42-51, two passes like the reference: means, then squared errors
        std::vector<double> suml(q, 0.0), sumr(q, 0.0), ssel(q, 0.0), sser(q, 0.0);
        for (size_t a = 0; a < n; ++a) {
            const double v = values[a], g = c.gain(instances[a]);
            for (size_t e = 0; e < q; ++e) (v < position[kept[e]] ? suml[e] : sumr[e]) += g;
        }
        std::vector<double> meanl(q), meanr(q);
        for (size_t e = 0; e < q; ++e) {
            const size_t nl = below[kept[e]], nr = n - nl;
            meanl[e] = nl ? suml[e] / (double)nl : 0.0;
            meanr[e] = nr ? sumr[e] / (double)nr : 0.0;
        }
        for (size_t a = 0; a < n; ++a) {
            const double v = values[a], g = c.gain(instances[a]);
            for (size_t e = 0; e < q; ++e) {
                if (v < position[kept[e]]) {
                    const double diff = meanl[e] - g;
                    ssel[e] += diff * diff;
                } else {
                    const double diff = meanr[e] - g;
                    sser[e] += diff * diff;
                }
            }
        }
        for (size_t e = 0; e < q; ++e) imp[e] = -(ssel[e] + sser[e]);
    } else if (method == "TrueVarianceReduction") {  // :115-122
        std::vector<StreamingStats> sl(q), sr(q);
        for (size_t a = 0; a < n; ++a) {
            const double v = values[a], g = c.gain(instances[a]);
            for (size_t e = 0; e < q; ++e) (v < position[kept[e]] ? sl[e] : sr[e]).push(g);
        }
        for (size_t e = 0; e < q; ++e) {
            if (!sl[e].finished() || !sr[e].finished())
                throw Error("TrueVarianceReduction needs at least two instances on each side of a split");
            imp[e] = -(sl[e].variance() * (double)sl[e].n + sr[e].variance() * (double)sr[e].n);
        }
    } else {  // BinaryGiniImpurity :52-66,:95-103 / InformationGain :74-87,:104-112
        std::vector<size_t> posl(q, 0), posr(q, 0);
        for (size_t a = 0; a < n; ++a) {
            if (!(c.ds.gains[instances[a]] > 0.0f)) continue;
            const double v = values[a];
            for (size_t e = 0; e < q; ++e) (v < position[kept[e]] ? posl[e] : posr[e]) += 1;
        }
        auto side = [&](size_t positive, size_t count) {
            if (count == 0) return 0.0;
            const double cnt = (double)count;
            const double p_yes = (double)positive / cnt, p_no = (cnt - (double)positive) / cnt;
            if (method == "BinaryGiniImpurity") return (p_yes * (1.0 - p_yes) + p_no * (1.0 - p_no)) * cnt;
            return (-plogp(p_yes) - plogp(p_no)) * cnt;
        };
        for (size_t e = 0; e < q; ++e) {
            const size_t nl = below[kept[e]], nr = n - nl;
            imp[e] = -(side(posl[e], nl) + side(posr[e], nr));
        }
    }
    size_t best = 0;
    for (size_t e = 0; e < q; ++e) {
        if (imp[e] != imp[e]) throw Error("split importance is NaN");
        if (imp[e] >= imp[best]) best = e;  // sort by importance, take the last
    }
    out.valid = true;
    out.fid = fid;
    out.split = position[kept[best]];
    out.importance = imp[best];
    out.n_lhs = below[kept[best]];
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
    // label statistics of the node (random_forest.rs:217-221), the same for every feature
    StreamingStats labels;
    for (uint32_t inst : instances) labels.push(c.gain(inst));
    const bool splittable = labels.finished() && labels.max != labels.min;
    FeatureSplit best;
    std::vector<double> values(instances.size());
    for (size_t a = 0; splittable && a < features.size(); ++a) {
        if (!fstats[a].finished()) continue;
        for (size_t i = 0; i < instances.size(); ++i) {
            double v = 0.0;
            c.feature_value(instances[i], features[a], &v);  // unwrap_or(0.0), :226-229
            values[i] = v;
        }
        FeatureSplit cand = generate_split_candidate(c, features[a], instances, fstats[a], values);
        if (!cand.valid) continue;
        if (!best.valid || cand.importance >= best.importance) best = cand;
    }
    if (!best.valid) return nullptr;  // NoFeatureSplitCandidates
    std::vector<uint32_t> lhs_ids, rhs_ids;
    lhs_ids.reserve(best.n_lhs);
    rhs_ids.reserve(instances.size() - best.n_lhs);
    for (uint32_t inst : instances) {
        double v = 0.0;
        c.feature_value(inst, best.fid, &v);
        (v < best.split ? lhs_ids : rhs_ids).push_back(inst);
    }
    std::unique_ptr<TreeNode> lhs = learn_recursive(c, features, lhs_ids, depth + 1);
    if (!lhs) lhs = leaf(compute_output(c, lhs_ids.data(), lhs_ids.size()));
    std::unique_ptr<TreeNode> rhs = learn_recursive(c, features, rhs_ids, depth + 1);
    if (!rhs) rhs = leaf(compute_output(c, rhs_ids.data(), rhs_ids.size()));
    std::unique_ptr<TreeNode> node(new TreeNode());
    node->leaf = false;
    node->fid = best.fid;
    node->split = best.split;
    node->lhs = std::move(lhs);
    node->rhs = std::move(rhs);
    return node;
}


// ---------------------------------------------------------------------------------------
// The same learner with the per-level statistics produced on the GPU (rf_induction.cu,
// SURVEY.md 8f.2).  The decisions are the host's: stop rules (random_forest.rs:369-378),
// threshold de-duplication and leaf-size test (:236-259), importance (:90-125), "last maximal"
// among candidates (:270, :395).  Sums are integers (labels in units of 2^-FR_RF_GAIN_BITS), so a
// forest does not depend on the order in which the device adds; the squared error of a side is
// sq - sum^2 / n from those exact integers where the two-pass host code sums (mean - y)^2 -- the
// same number up to its last bits, which can only matter between candidates that tie to ~1e-15.
// ---------------------------------------------------------------------------------------
struct TooManyNodes {};

double side_importance(const std::string &method, uint64_t n, uint64_t positive, int64_t sum, int64_t sq,
                       bool *degenerate) {
    if (n == 0) return 0.0;
    if (method == "BinaryGiniImpurity" || method == "InformationGain") {
        const double cnt = (double)n;
        const double p_yes = (double)positive / cnt, p_no = (cnt - (double)positive) / cnt;
        if (method == "BinaryGiniImpurity") return (p_yes * (1.0 - p_yes) + p_no * (1.0 - p_no)) * cnt;
        return (-plogp(p_yes) - plogp(p_no)) * cnt;
    }
    const __int128 num = (__int128)n * (__int128)sq - (__int128)sum * (__int128)sum;  // n * SSE, exact
    const double scale = (double)(1ull << (2 * FR_RF_GAIN_BITS));
    const double sse = (double)num / ((double)n * scale);
    if (method == "SquaredError") return sse;
    if (n < 2) {  // label_stats(..).unwrap() of a single instance
        *degenerate = true;
        return 0.0;
    }
    return sse / (double)(n - 1) * (double)n;  // variance * weight
}

std::unique_ptr<TreeNode> learn_tree_device(const Ctx &c, fr_dev_rf *rf, const std::vector<uint32_t> &features,
                                            const std::vector<uint32_t> &instances) {
    const RandomForestParams &p = c.p;
    auto mean_of = [](int64_t sum, uint64_t n) {
        return n ? (double)sum / (double)(1ull << FR_RF_GAIN_BITS) / (double)n : 0.0;
    };
    // the root's stop rules need no statistics (:369-378)
    if (features.empty() || instances.empty() || 1 >= p.max_depth || instances.size() < (size_t)p.min_leaf_support)
        return leaf(compute_output(c, instances.data(), instances.size()));
    if (fr_dev_rf_begin_tree(rf, instances.data(), instances.size(), features.data(), features.size()))
        throw Error(fr_dev_last_error());
    struct Active {
        std::unique_ptr<TreeNode> *slot;
        uint32_t depth;
    };
    std::unique_ptr<TreeNode> root;
    std::vector<Active> active{{&root, 1u}};
    const uint32_t k = p.split_candidates;
    const size_t F = features.size();
    std::vector<uint64_t> node_n;
    std::vector<int64_t> node_sum, b_sum, b_sq;
    std::vector<float> gmin, gmax, fmin, fmax;
    std::vector<uint32_t> b_n, b_pos, t_fid, f_present;
    std::vector<double> t_split;
    std::vector<int32_t> t_left, t_right;
    while (!active.empty()) {
        const uint32_t na = (uint32_t)active.size();
        if (na > 1024) throw TooManyNodes();
        node_n.assign(na, 0);
        node_sum.assign(na, 0);
        gmin.assign(na, 0.f);
        gmax.assign(na, 0.f);
        fmin.assign((size_t)na * F, 0.f);
        fmax.assign((size_t)na * F, 0.f);
        b_n.assign((size_t)na * F * k, 0);
        b_pos.assign((size_t)na * F * k, 0);
        b_sum.assign((size_t)na * F * k, 0);
        b_sq.assign((size_t)na * F * k, 0);
        f_present.assign((size_t)na * F, 0);
        if (fr_dev_rf_level_stats(rf, na, k, node_n.data(), node_sum.data(), gmin.data(), gmax.data(), fmin.data(),
                                  fmax.data(), b_n.data(), b_pos.data(), b_sum.data(), b_sq.data(), f_present.data()))
            throw Error(fr_dev_last_error());
        t_fid.assign(na, 0xffffffffu);
        t_split.assign(na, 0.0);
        t_left.assign(na, -1);
        t_right.assign(na, -1);
        std::vector<Active> next;
        for (uint32_t a = 0; a < na; ++a) {
            const uint64_t n = node_n[a];
            bool have = false;
            double best_imp = 0.0, best_split = 0.0;
            uint32_t best_f = 0;
            uint64_t best_nl = 0, best_posl = 0;
            int64_t best_suml = 0;
            // label_stats (:217-221) and FeatureStats.finish() both need more than one instance
            if (n > 1 && gmax[a] != gmin[a]) {
                for (size_t fa = 0; fa < F; ++fa) {
                    // FeatureStats.finish(): a feature fewer than two of the node's rows carry has no
                    // statistics and proposes no split (normalizers.rs:29-35, random_forest.rs:387-391)
                    if (f_present[(size_t)a * F + fa] < 2) continue;
                    const double lo = (double)fmin[(size_t)a * F + fa];
                    const double range = (double)fmax[(size_t)a * F + fa] - lo;
                    const size_t base = ((size_t)a * F + fa) * k;
                    uint64_t tot_pos = 0;
                    int64_t tot_sum = 0, tot_sq = 0;
                    for (uint32_t b = 0; b < k; ++b) {
                        tot_pos += b_pos[base + b];
                        tot_sum += b_sum[base + b];
                        tot_sq += b_sq[base + b];
                    }
                    uint64_t nl = 0, posl = 0, prev = 0;
                    int64_t suml = 0, sql = 0;
                    bool have_prev = false, cand_have = false;
                    double cand_imp = 0.0, cand_split = 0.0;
                    uint64_t cand_nl = 0, cand_posl = 0;
                    int64_t cand_suml = 0;
                    for (uint32_t i = 1; i < k; ++i) {
                        nl += b_n[base + i - 1];
                        posl += b_pos[base + i - 1];
                        suml += b_sum[base + i - 1];
                        sql += b_sq[base + i - 1];
                        if (have_prev && prev == nl) continue;
                        have_prev = true;
                        prev = nl;
                        const uint64_t nr = n - nl;
                        if (nl < p.min_leaf_support || nr < p.min_leaf_support) continue;
                        bool degenerate = false;
                        const double imp = -(side_importance(p.split_method, nl, posl, suml, sql, &degenerate) +
                                             side_importance(p.split_method, nr, tot_pos - posl, tot_sum - suml,
                                                             tot_sq - sql, &degenerate));
                        if (degenerate)
                            throw Error("TrueVarianceReduction needs at least two instances on each side of a split");
                        if (imp != imp) throw Error("split importance is NaN");
                        if (!cand_have || imp >= cand_imp) {
                            cand_have = true;
                            cand_imp = imp;
                            cand_split = ((double)i / (double)k) * range + lo;
                            cand_nl = nl;
                            cand_posl = posl;
                            cand_suml = suml;
                        }
                    }
                    if (cand_have && (!have || cand_imp >= best_imp)) {
                        have = true;
                        best_imp = cand_imp;
                        best_split = cand_split;
                        best_f = (uint32_t)fa;
                        best_nl = cand_nl;
                        best_posl = cand_posl;
                        best_suml = cand_suml;
                    }
                }
            }
            (void)best_posl;
            if (!have) {  // NoFeatureSplitCandidates: the parent turns this node into a leaf
                *active[a].slot = leaf(mean_of(node_sum[a], n));
                continue;
            }
            std::unique_ptr<TreeNode> node(new TreeNode());
            node->leaf = false;
            node->fid = features[best_f];
            node->split = best_split;
            TreeNode *raw = node.get();
            *active[a].slot = std::move(node);
            t_fid[a] = features[best_f];
            t_split[a] = best_split;
            const uint32_t child_depth = active[a].depth + 1;
            auto child = [&](std::unique_ptr<TreeNode> *slot, uint64_t cn, int64_t csum, int32_t *id_out) {
                if (cn == 0 || child_depth >= p.max_depth || cn < p.min_leaf_support) {
                    *slot = leaf(mean_of(csum, cn));  // the recursion would return Err here
                    *id_out = -1;
                } else {
                    *id_out = (int32_t)next.size();
                    next.push_back({slot, child_depth});
                }
            };
            child(&raw->lhs, best_nl, best_suml, &t_left[a]);
            child(&raw->rhs, n - best_nl, node_sum[a] - best_suml, &t_right[a]);
        }
        if (fr_dev_rf_partition(rf, na, t_fid.data(), t_split.data(), t_left.data(), t_right.data()))
            throw Error(fr_dev_last_error());
        active.swap(next);
    }
    return root;
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

    // Where do the per-level statistics come from?  The device knows which features a libsvm row
    // carries (fr_dev_dataset_set_row_lengths for Dense32 rows, _set_row_presence when some rows
    // are Sparse32); it needs labels that are exact in its integer unit; FASTRANK_RF=host / gpu
    // overrides the size heuristic.
    const bool device_capable = true;
    bool on_device = device_capable && view.num_instances() >= 50000;
    if (const char *env = getenv("FASTRANK_RF")) {
        if (std::string(env) == "host") on_device = false;
        if (std::string(env) == "gpu") on_device = device_capable;
    }
    if (on_device) {
        const double unit = (double)(1 << FR_RF_GAIN_BITS);
        for (size_t i = 0; i < parent.n && on_device; ++i) {
            const double g = (double)parent.gains[i] * unit;
            if (g != std::floor(g) || std::fabs((double)parent.gains[i]) > 16.0) on_device = false;
        }
    }
    // one induction state (own stream, own buffers) per worker thread: while one tree's level
    // statistics are being turned into decisions on the host, other trees keep the GPU busy
    struct RfHandles {
        std::vector<fr_dev_rf *> p;
        ~RfHandles() {
            for (fr_dev_rf *h : p) fr_dev_rf_destroy(h);
        }
    } rf;
    const unsigned device_workers = on_device ? std::min<unsigned>(4u, p.num_trees) : 0u;
    for (unsigned w = 0; w < device_workers; ++w) {
        fr_dev_rf *h = nullptr;
        if (fr_dev_rf_create(view.parent->device(), &h)) throw Error(fr_dev_last_error());
        rf.p.push_back(h);
    }

    std::vector<std::unique_ptr<TreeNode>> trees(p.num_trees);
    std::vector<std::string> failures(p.num_trees);
    std::atomic<uint32_t> next{0};
    auto worker = [&](unsigned worker_id) {
        fr_dev_rf *my_rf = worker_id < rf.p.size() ? rf.p[worker_id] : nullptr;
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
                std::unique_ptr<TreeNode> root;
                bool done = false;
                if (my_rf) {
                    try {
                        root = learn_tree_device(ctx, my_rf, features, instances);
                        done = true;
                    } catch (const TooManyNodes &) {  // more than 1024 open nodes on a level: host path
                    }
                }
                if (!done) {
                    root = learn_recursive(ctx, features, instances, 1);
                    if (!root) root = leaf(compute_output(ctx, instances.data(), instances.size()));  // :344-352
                }
                trees[idx] = std::move(root);
            } catch (const std::exception &e) {
                failures[idx] = e.what();
            }
        }
    };
    {
        unsigned hw = std::thread::hardware_concurrency();
        const unsigned nthreads =
            std::max(1u, device_workers ? device_workers : std::min<unsigned>(hw ? hw : 1u, p.num_trees));
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(worker, t);
        worker(0);
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
