// coordinate_ascent.cpp -- the coordinate-ascent learner (coordinate_ascent.rs:43-253) as a
// host state machine over GPU line searches.
//
// The reference runs restarts on rayon threads and, inside a restart, calls evaluate_mean up
// to 1 + 2*T times per feature, one weight vector at a time.  The candidate weights of a
// feature depend only on (orig, step_base, step_scale, T), never on a score, so here every
// restart submits its whole next group of candidates to the GPU at once
// (fr_dev_eval_coord_sweeps_fast: ONE pass over the feature matrix for all restarts), and the
// reference's sequential accept / early-break logic is replayed on the returned means.
//   group A = direction 0 (weight -> 0) + direction -1   (1 + T candidates)
//   group B = direction +1                               (T candidates), which the reference
//             only reaches when A did not improve the score (coordinate_ascent.rs:174-176).
// Direction -1 is speculative with respect to the break after direction 0.  With the batched
// sweep, group B is submitted together with group A: on the 1M-document benchmark the reference
// reaches direction +1 in 94 % of the line searches, and one launch with 51 candidates per
// restart shares one pass over the feature matrix instead of two (FASTRANK_SPECULATE=0 restores
// the two-submission schedule).  The statistics keep "consumed" (what the reference's control
// flow evaluates) and "computed" (what the GPU scored) apart.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "host.hpp"

namespace frb {

namespace {

std::mutex g_stats_mu;
TrainStats g_last_stats;

void l1_normalize(std::vector<double> &w) {  // coordinate_ascent.rs:72-82
    double sum = 0.0;
    for (double x : w) sum += std::fabs(x);
    if (sum > 0.0)
        for (double &x : w) x /= sum;
}

struct Restart {
    uint32_t id = 0;
    Rand64 rng;
    std::vector<double> best_w;
    double best_score = 0.0;
    std::vector<uint32_t> order;
    size_t fi = 0;
    size_t successes = 0;
    bool done = false;
    // line search in flight
    std::vector<double> model;
    double start_score = 0.0;
    double orig = 0.0;
    uint32_t feature = 0;
    int group = 0;  // 0 = A pending, 1 = B pending
    explicit Restart(unsigned __int128 seed) : rng(seed) {}
};

// coordinate_ascent.rs:145-171: the weights one direction tries
void direction_candidates(const CoordinateAscentParams &p, double orig, int dir, std::vector<double> &out) {
    double step = p.step_base * (double)dir;
    if (orig != 0.0 && std::fabs(step) > 0.5 * std::fabs(orig)) step = p.step_base * std::fabs(orig) * (double)dir;
    double total = step;
    uint32_t iters = p.num_max_iterations;
    if (dir == 0) {
        iters = 1;
        total = -orig;
    }
    for (uint32_t it = 0; it < iters; ++it) {
        out.push_back(orig + total);
        step *= p.step_scale;
        total += step;
    }
}

bool replace_if_better(Restart &r, double score, double w) {  // core.rs:57-66
    if (score == score && score > r.best_score) {
        r.best_score = score;
        r.best_w = r.model;
        r.best_w[r.feature] = w;
        return true;
    }
    return false;
}

}  // namespace

TrainStats last_train_stats() {
    std::lock_guard<std::mutex> lock(g_stats_mu);
    return g_last_stats;
}

void set_last_train_stats(const TrainStats &s) {
    std::lock_guard<std::mutex> lock(g_stats_mu);
    g_last_stats = s;
}

Model coordinate_ascent_learn(const CoordinateAscentParams &p, const DatasetView &view,
                              const Evaluator &ev, TrainStats *stats_out) {
    TrainStats stats;
    const std::vector<uint32_t> fids = view.feature_ids();
    if (fids.empty()) throw Error("Should be at least one feature!");
    if (view.num_instances() == 0) throw Error("dataset has no instances");
    if (ev.num_queries() == 0 && ev.global_queries() == 0) throw Error("dataset has no queries");
    if (p.num_restarts == 0) throw Error("Should be at least 1 restart!");
    uint32_t dim = 0;
    for (uint32_t f : fids) dim = std::max(dim, f + 1);  // coordinate_ascent.rs:97-103
    const std::string measure_name = ev.measure().display;

    if (!p.quiet) {
        printf("---------------------------\nTraining starts...\n---------------------------\n");
    }
    Rand64 master((unsigned __int128)p.seed);
    std::vector<Restart> rs;
    rs.reserve(p.num_restarts);
    for (uint32_t r = 0; r < p.num_restarts; ++r) {
        rs.emplace_back((unsigned __int128)master.rand_u64());  // coordinate_ascent.rs:211-213
        rs.back().id = r;
        if (!p.quiet) printf("[+] Random restart #%u/%u...\n", r + 1, p.num_restarts);
    }
    // reset(): coordinate_ascent.rs:50-70, then the start score (:110) for every restart at once
    std::vector<std::vector<double>> init(p.num_restarts, std::vector<double>(dim, 0.0));
    for (uint32_t r = 0; r < p.num_restarts; ++r) {
        for (uint32_t f : fids)
            init[r][f] = p.init_random ? (rs[r].rng.rand_float() * 2.0) - 1.0 : 1.0 / (double)fids.size();
    }
    {
        const std::vector<double> start = ev.evaluate_linear(init);
        stats.evals_consumed += p.num_restarts;
        stats.evals_computed += p.num_restarts;
        for (uint32_t r = 0; r < p.num_restarts; ++r) {
            rs[r].best_w = init[r];
            rs[r].best_score = start[r];
        }
    }

    const size_t T = p.num_max_iterations;
    bool use_fast = fr_dev_plan_has_fast_sweep(ev.plan()) != 0;
    if (const char *env = getenv("FASTRANK_SWEEP"))
        if (std::string(env) == "exact") use_fast = false;
    if (p.exact_sweep) use_fast = false;  // asked for in the train request
    stats.exact_sweep = !use_fast;
    bool speculate = use_fast && T > 0;
    if (const char *env = getenv("FASTRANK_SPECULATE")) speculate = speculate && atoi(env) != 0;
    const size_t stride = speculate ? 1 + 2 * T : 1 + T;
    std::vector<double> base_w, cand_w;
    std::vector<uint32_t> fid_arr, ncand;
    std::vector<int64_t> sums;
    // One submitted line search.  A restart submits the line search of its current feature and,
    // when the launch has room (fewer active restarts than the kernel serves per pass over X),
    // LOOKAHEAD line searches for the features that follow in its shuffled order, built from the
    // same best model.  They are valid exactly when the earlier features leave the best model
    // untouched -- which is what the last pass of every restart looks like -- and are discarded
    // otherwise.
    struct Sweep {
        size_t slot;        // index into `active`
        uint32_t feature;
        double orig;
        size_t n_a;         // candidates of directions 0 and -1
        bool has_b;         // direction +1 rides along
        std::vector<double> cands;
    };
    const size_t kLaunchSweeps = 8;  // sweeps sharing one pass over X in fr_dev_eval_coord_sweeps_fast
    // On by default (FASTRANK_LOOKAHEAD=0 turns it off): on the 1M-document benchmark it cuts the
    // launches of a training run by 31 % (816 -> 563) for 9 % more candidates scored -- device time
    // 0.588 -> 0.554 s on one GPU, more where a step is mostly fixed latency (small shards).  The
    // model is the same either way.
    bool lookahead = speculate;
    if (const char *env = getenv("FASTRANK_LOOKAHEAD")) lookahead = speculate && atoi(env) != 0;
    std::vector<Restart *> active;
    std::vector<Sweep> sweeps;
    for (;;) {
        active.clear();
        for (Restart &r : rs)
            if (!r.done) active.push_back(&r);
        if (active.empty()) break;
        // 1. every active restart prepares its next group(s) of candidates
        sweeps.clear();
        // idle sweep slots of the launch are dealt out to the active restarts, one more line search
        // each, the first restarts taking what does not divide evenly
        const size_t idle = lookahead && active.size() < kLaunchSweeps ? kLaunchSweeps - active.size() : 0;
        for (size_t a = 0; a < active.size(); ++a) {
            Restart &r = *active[a];
            const size_t extra = idle / active.size() + (a < idle % active.size() ? 1 : 0);
            if (r.group == 0) {
                if (r.fi == 0) {
                    r.order = fids;
                    shuffle(r.order, r.rng);  // coordinate_ascent.rs:114-116
                    r.successes = 0;
                    if (!p.quiet) {
                        printf("Shuffle features and optimize!\n----------------------------------------\n");
                        printf("%4u|%-16s|%9s|%9s\n", r.id, "Feature", "Weight", measure_name.c_str());
                        printf("----------------------------------------\n");
                    }
                }
                r.model = r.best_w;
                if (p.normalize) l1_normalize(r.model);  // :134-138 (the clone, not the best)
                const size_t n_here = std::min(1 + extra, r.order.size() - r.fi);
                for (size_t m = 0; m < n_here; ++m) {
                    Sweep sw;
                    sw.slot = a;
                    sw.feature = r.order[r.fi + m];
                    sw.orig = r.model[sw.feature];
                    direction_candidates(p, sw.orig, 0, sw.cands);
                    direction_candidates(p, sw.orig, -1, sw.cands);
                    sw.n_a = sw.cands.size();
                    sw.has_b = speculate;
                    if (sw.has_b) direction_candidates(p, sw.orig, +1, sw.cands);
                    sweeps.push_back(std::move(sw));
                }
            } else {  // direction +1 of the feature whose directions 0 / -1 did not improve
                Sweep sw;
                sw.slot = a;
                sw.feature = r.feature;
                sw.orig = r.orig;
                sw.n_a = 0;
                sw.has_b = false;
                direction_candidates(p, r.orig, +1, sw.cands);
                sweeps.push_back(std::move(sw));
            }
        }
        base_w.assign(sweeps.size() * dim, 0.0);
        cand_w.assign(sweeps.size() * stride, 0.0);
        fid_arr.assign(sweeps.size(), 0);
        ncand.assign(sweeps.size(), 0);
        for (size_t i = 0; i < sweeps.size(); ++i) {
            const Sweep &sw = sweeps[i];
            const Restart &r = *active[sw.slot];
            std::copy(r.model.begin(), r.model.end(), base_w.begin() + i * dim);
            std::copy(sw.cands.begin(), sw.cands.end(), cand_w.begin() + i * stride);
            fid_arr[i] = sw.feature;
            ncand[i] = (uint32_t)sw.cands.size();
            stats.evals_computed += sw.cands.size();
        }
        // 2. one GPU pass over the feature matrix for all of them (batched sweep); the
        //    exact-order kernel (one pass per restart) on request or for very long queries
        sums.assign(sweeps.size() * stride, 0);
        const auto t_dev = std::chrono::steady_clock::now();
        const int rc = use_fast
                           ? fr_dev_eval_coord_sweeps_fast(ev.plan(), sweeps.size(), base_w.data(), dim,
                                                           fid_arr.data(), cand_w.data(), ncand.data(),
                                                           stride, sums.data(), nullptr)
                           : fr_dev_eval_coord_sweeps(ev.plan(), sweeps.size(), base_w.data(), dim,
                                                      fid_arr.data(), cand_w.data(), ncand.data(), stride,
                                                      sums.data());
        stats.seconds_device += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_dev).count();
        if (rc) throw Error(fr_dev_last_error());
        stats.sweeps += sweeps.size();
        stats.global_steps += 1;
        // 3. replay the reference's sequential control flow on the means
        for (size_t i = 0; i < sweeps.size();) {
            const size_t a = sweeps[i].slot;
            Restart &r = *active[a];
            size_t end = i;
            while (end < sweeps.size() && sweeps[end].slot == a) ++end;
            for (size_t cur = i; cur < end; ++cur) {
                const Sweep &sw = sweeps[cur];
                bool best_changed = false;
                if (r.group == 0) {
                    r.feature = sw.feature;
                    r.orig = sw.orig;
                    r.start_score = r.best_score;
                }
                const std::string fname = p.quiet ? std::string() : view.parent->feature_name(r.feature);
                auto take = [&](size_t k) {
                    const double score = ev.mean_from_fx(sums[cur * stride + k]);
                    stats.evals_consumed += 1;
                    if (replace_if_better(r, score, sw.cands[k])) {
                        best_changed = true;
                        if (!p.quiet) printf("%4u|%-16s|%9.3f|%9.3f\n", r.id, fname.c_str(), sw.cands[k], score);
                    }
                };
                bool feature_done = false;
                if (r.group == 0) {
                    take(0);  // direction 0
                    if ((r.best_score - r.start_score) > p.tolerance) {
                        feature_done = true;  // :174-176
                    } else {
                        for (size_t k = 1; k < sw.n_a; ++k) take(k);  // direction -1
                        if ((r.best_score - r.start_score) > p.tolerance) feature_done = true;
                        else if (T == 0) feature_done = true;
                        else if (sw.has_b) {
                            for (size_t k = sw.n_a; k < sw.cands.size(); ++k) take(k);  // direction +1
                            feature_done = true;
                        } else r.group = 1;
                    }
                } else {
                    for (size_t k = 0; k < sw.cands.size(); ++k) take(k);  // direction +1
                    feature_done = true;
                }
                if (!feature_done) break;  // direction +1 follows in the next submission
                if ((r.best_score - r.start_score) > p.tolerance) r.successes += 1;  // :181-183
                r.group = 0;
                r.fi += 1;
                if (r.fi == r.order.size()) {
                    r.fi = 0;
                    if (r.successes == 0) r.done = true;  // :185-187
                    else if (!p.quiet) printf("---------------------------\n");
                    break;  // the next pass starts with a new shuffle
                }
                // a lookahead result stands only if this feature left the best model untouched
                if (best_changed) break;
            }
            i = end;
        }
    }
    if (!p.quiet) printf("---------------------------\nFinished successfully.\n");

    Model out;
    if (p.output_ensemble && rs.size() > 1) {  // :232-242
        out.kind = Model::Ensemble;
        for (Restart &r : rs) {
            std::vector<double> w = r.best_w;
            l1_normalize(w);
            out.members.push_back(Model::linear(std::move(w)));
            out.weights.push_back(r.best_score);
        }
    } else {  // :244-251, Iterator::max keeps the LAST maximal element
        size_t best = 0;
        for (size_t r = 1; r < rs.size(); ++r)
            if (rs[r].best_score >= rs[best].best_score) best = r;
        out = Model::linear(rs[best].best_w);
    }
    if (stats_out) *stats_out = stats;
    set_last_train_stats(stats);
    return out;
}

}  // namespace frb
