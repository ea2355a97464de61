// evaluator.cpp -- SetEvaluator on the host (evaluators.rs:98-224): parses the measure name,
// turns a dataset view (+ optional judgments) into a device plan, and forwards every
// evaluation to the kernels.  No metric arithmetic over documents happens here; the only
// numbers computed on the host are the per-query norms that come from a qrel file
// (evaluators.rs:310-318, :397-401), which depend on judged documents, not on the dataset.
#include <algorithm>
#include <cmath>

#include "host.hpp"

namespace frb {

Measure Measure::parse(const std::string &orig) {
    Measure m;
    std::string name = orig;
    const size_t at = orig.find('@');
    if (at != std::string::npos) {
        name = orig.substr(0, at);
        const std::string rhs = orig.substr(at);
        const std::string digits = rhs.substr(1);
        bool ok = !digits.empty() && digits.size() < 19;
        for (char c : digits) ok = ok && c >= '0' && c <= '9';
        if (!ok) throw Error("Couldn't parse after the @ in \"" + orig + "\": " + rhs);
        m.depth = (int64_t)strtoull(digits.c_str(), nullptr, 10);
    }
    for (char &c : name) c = (char)tolower((unsigned char)c);
    if (name == "ap" || name == "map") {
        m.metric = FR_METRIC_AP;
        m.depth = -1;
        m.display = "AP";
    } else if (name == "rr" || name == "mrr") {
        m.metric = FR_METRIC_RR;
        m.depth = -1;
        m.display = "RR";
    } else if (name == "ndcg") {
        m.metric = FR_METRIC_NDCG;
        m.display = m.depth >= 0 ? "NDCG@" + std::to_string(m.depth) : "NDCG";
    } else {
        throw Error("Invalid training measure: \"" + orig + "\"");
    }
    return m;
}

// compute_dcg(gains, depth, ideal=true) for a judged gain vector (evaluators.rs:255-272)
static double ideal_dcg_from_judgments(std::vector<float> gains, int64_t depth) {
    std::sort(gains.begin(), gains.end(), [](float a, float b) { return a > b; });
    size_t lim = gains.size();
    if (depth >= 0 && (size_t)depth < lim) lim = (size_t)depth;
    double dcg = 0.0;
    for (size_t i = 0; i < lim; ++i)
        dcg += (std::pow(2.0, (double)gains[i]) - 1.0) / std::log2((double)i + 2.0);
    return dcg;
}

Evaluator::Evaluator(const DatasetView &view, const Measure &measure, const QRel *qrel)
    : view_(view), measure_(measure) {
    ParentDataset &parent = *view.parent;
    fr_dev_dataset *dev = parent.device();
    use_lock_ = std::unique_lock<std::recursive_mutex>(parent.use_mu);
    std::vector<uint64_t> inst_off{0};
    std::vector<uint32_t> inst_ids;
    bool subset = false;
    if (!view.sampled) {  // every query of the parent, whole: nothing to group
        view_queries_.resize(parent.by_query.size());
        for (uint32_t q = 0; q < view_queries_.size(); ++q) view_queries_[q] = q;
    } else {
        const auto groups = view.instances_by_query();
        for (const auto &g : groups) {
            view_queries_.push_back(g.first);
            if (g.second.size() != parent.by_query[g.first].size()) subset = true;
        }
        if (subset) {
            for (const auto &g : groups) {
                inst_ids.insert(inst_ids.end(), g.second.begin(), g.second.end());
                inst_off.push_back(inst_ids.size());
            }
        }
    }
    // a plan built earlier for the same view and measure (no judgments involved) is reused
    fr_dev_comm *const comm_now = fr_dev_default_comm();
    const bool cacheable = !(qrel && measure.metric != FR_METRIC_RR);
    if (cacheable) {
        for (const std::shared_ptr<PlanHolder> &h : parent.plan_cache) {
            if (h->metric == measure.metric && h->depth == measure.depth && h->comm == comm_now &&
                h->comm_generation == fr_dev_comm_generation(comm_now) &&
                h->sampled == view.sampled && (!view.sampled || h->instances == view.instances)) {
                holder_ = h;
                plan_ = h->plan;
                return;
            }
        }
    }
    std::vector<uint8_t> present;
    std::vector<double> value;
    if (qrel && measure.metric != FR_METRIC_RR) {
        present.assign(view_queries_.size(), 0);
        value.assign(view_queries_.size(), 0.0);
        for (size_t v = 0; v < view_queries_.size(); ++v) {
            const QueryJudgments *qj = qrel->get(parent.query_names[view_queries_[v]]);
            if (!qj) continue;
            present[v] = 1;
            if (measure.metric == FR_METRIC_NDCG) {
                std::vector<float> gv = qj->gain_vector();
                value[v] = gv.empty() ? NAN : ideal_dcg_from_judgments(std::move(gv), measure.depth);
            } else {
                value[v] = (double)qj->num_relevant();
            }
        }
    }
    fr_dev_plan_desc desc;
    desc.metric = measure.metric;
    desc.depth = measure.depth;
    desc.n_queries = (uint32_t)view_queries_.size();
    const bool all_queries = view_queries_.size() == parent.query_names.size();
    desc.query_ids = all_queries ? nullptr : view_queries_.data();
    desc.inst_off = subset ? inst_off.data() : nullptr;
    desc.inst_ids = subset ? inst_ids.data() : nullptr;
    desc.norm_present = present.empty() ? nullptr : present.data();
    desc.norm_value = value.empty() ? nullptr : value.data();
    holder_ = std::make_shared<PlanHolder>();
    if (fr_dev_plan_create(dev, &desc, &holder_->plan)) throw Error(fr_dev_last_error());
    plan_ = holder_->plan;
    if (comm_now) {
        if (fr_dev_plan_set_comm(plan_, comm_now)) throw Error(fr_dev_last_error());
    }
    if (cacheable) {
        holder_->metric = measure.metric;
        holder_->depth = measure.depth;
        holder_->sampled = view.sampled;
        if (view.sampled) holder_->instances = view.instances;
        holder_->comm = comm_now;
        holder_->comm_generation = fr_dev_comm_generation(comm_now);
        if (parent.plan_cache.size() >= 8) parent.plan_cache.erase(parent.plan_cache.begin());
        parent.plan_cache.push_back(holder_);
    }
}

Evaluator::~Evaluator() {}

double Evaluator::mean_from_fx(int64_t fx) const {
    const uint64_t nq = global_queries();
    if (nq == 0) return 0.0;  // evaluators.rs:175-177
    return std::ldexp((double)fx, -FR_FX_BITS) / (double)nq;
}

double Evaluator::evaluate_mean(const Model &m, std::vector<double> *per_query, uint64_t model_uid) const {
    int64_t fx = 0;
    if (per_query) per_query->assign(num_queries(), 0.0);
    double *pq = per_query && !per_query->empty() ? per_query->data() : nullptr;
    if (m.kind == Model::Linear) {
        static const double zero = 0.0;
        const double *w = m.weights.empty() ? &zero : m.weights.data();
        if (fr_dev_eval_linear_batch(plan_, w, m.weights.size(), 1, &fx, pq)) throw Error(fr_dev_last_error());
    } else {
        bool owned = false;
        fr_dev_model *dm = view_.parent->device_model(m, model_uid, &owned);
        const int rc = fr_dev_eval_model(plan_, dm, &fx, pq);
        if (owned) fr_dev_model_destroy(dm);
        if (rc) throw Error(fr_dev_last_error());
    }
    return mean_from_fx(fx);
}

std::vector<double> Evaluator::evaluate_linear(const std::vector<std::vector<double>> &ws) const {
    std::vector<double> out(ws.size(), 0.0);
    if (ws.empty()) return out;
    const size_t wlen = ws[0].size();
    std::vector<double> flat(ws.size() * std::max<size_t>(wlen, 1), 0.0);
    for (size_t c = 0; c < ws.size(); ++c) {
        if (ws[c].size() != wlen) throw Error("evaluate_linear: ragged weight vectors");
        std::copy(ws[c].begin(), ws[c].end(), flat.begin() + c * wlen);
    }
    std::vector<int64_t> fx(ws.size(), 0);
    if (fr_dev_eval_linear_batch(plan_, flat.data(), wlen, ws.size(), fx.data(), nullptr))
        throw Error(fr_dev_last_error());
    for (size_t c = 0; c < ws.size(); ++c) out[c] = mean_from_fx(fx[c]);
    return out;
}

}  // namespace frb
