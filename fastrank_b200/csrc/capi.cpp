// capi.cpp -- the reference-compatible C ABI (reference src/lib.rs:50-326, src/ffi.rs:29-281).
//
// Same symbols, argument meaning, ownership and error envelope as the reference cdylib, so
// fastrank/clib.py (or the mirror in fastrank_b200/clib.py) binds it unchanged:
//   * strings in: NUL-terminated UTF-8, NULL reported as an error (ffi.rs:29-37)
//   * `const CResult*` out: heap struct, exactly one field set (ffi.rs:57-74)
//   * `const void*` out: heap JSON string, either the answer or {"error","context"}
//     (ffi.rs:45-55); released with free_str
#include <algorithm>
#include <charconv>
#include <cstring>
#include <fstream>

#include <chrono>

#include <atomic>

#include <sstream>

#include "host.hpp"

using namespace frb;

namespace {

char *dup_cstr(const std::string &s) {
    char *out = (char *)malloc(s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return out;
}

std::string error_json(const std::string &context) {
    json::Value v = json::Value::object();
    v.set("error", json::Value::string("error"));
    // the reference formats the boxed error with {:?}; for string errors that is the quoted text
    v.set("context", json::Value::string("\"" + context + "\""));
    return json::dump(v);
}

std::string accept_str(const char *name, const void *p) {  // ffi.rs:29-37
    if (!p) throw Error(std::string("NULL pointer: ") + name);
    return std::string((const char *)p);
}

template <typename F>
const void *json_call(F &&body) {  // ffi.rs:45-55
    std::string out;
    try {
        out = body();
    } catch (const std::exception &e) {
        out = error_json(e.what());
    }
    return dup_cstr(out);
}

template <typename F>
const CResult *result_call(F &&body) {  // ffi.rs:57-74
    CResult *res = (CResult *)malloc(sizeof(CResult));
    res->error_message = nullptr;
    res->success = nullptr;
    try {
        res->success = body();
    } catch (const std::exception &e) {
        res->error_message = dup_cstr(error_json(e.what()));
    }
    return res;
}

json::Value parse_json(const std::string &text) {
    try {
        return json::parse(text);
    } catch (const json::ParseError &e) {
        throw Error(std::string("Error(") + e.what() + ")");
    }
}

json::Value train_request_json(const char *kind, json::Value params) {  // ffi.rs:215-236
    json::Value req = json::Value::object();
    req.set("measure", json::Value::string("ndcg"));
    json::Value wrapped = json::Value::object();
    wrapped.set(kind, std::move(params));
    req.set("params", std::move(wrapped));
    req.set("judgments", json::Value::null());
    return req;
}

const CDataset &need(const CDataset *d) {
    if (!d) throw Error("Dataset pointer is null!");
    return *d;
}
const CModel &need(const CModel *m) {
    if (!m) throw Error("Model pointer is null!");
    return *m;
}

// Scores of every instance of the parent dataset (json_api.rs:59-69), straight into `out`
// (parent.n doubles).  The dataset's stream and its score scratch are shared device state: the
// call is serialised with evaluators and trainers on the same dataset (use_mu), so concurrent
// predict / evaluate / train calls from several threads cannot read each other's scores.
void score_all_into(const CModel &m, ParentDataset &parent, double *out) {
    fr_dev_dataset *dev = parent.device();
    std::lock_guard<std::recursive_mutex> lock(parent.use_mu);
    bool owned = false;
    fr_dev_model *dm = parent.device_model(m.model, m.uid, &owned);
    const int rc = fr_dev_score_model(dev, dm, out);
    if (owned) fr_dev_model_destroy(dm);
    if (rc) throw Error(fr_dev_last_error());
}

std::vector<double> score_all(const CModel &m, ParentDataset &parent) {
    std::vector<double> scores(parent.n);
    score_all_into(m, parent, scores.data());
    return scores;
}

std::string rust_display_f64(double v) {  // Display for f64: shortest digits, never scientific
    if (v != v) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

}  // namespace

CModel::CModel(frb::Model m) : model(std::move(m)) {
    static std::atomic<uint64_t> next{1};
    uid = next.fetch_add(1, std::memory_order_relaxed);
}

extern "C" {

void free_str(void *p) { free(p); }
void free_c_result(CResult *p) { free(p); }
void free_dataset(CDataset *p) { delete p; }
void free_model(CModel *p) { delete p; }
void free_cqrel(CQRel *p) { delete p; }

const CResult *load_cqrel(const void *data_path) {
    return result_call([&]() -> const void * {
        auto q = QRel::load_file(accept_str("data_path", data_path));
        return new CQRel{q};
    });
}

const CResult *cqrel_from_json(const void *json_str) {
    return result_call([&]() -> const void * {
        auto q = QRel::from_json(parse_json(accept_str("json_str", json_str)));
        return new CQRel{q};
    });
}

const void *cqrel_query_json(const CQRel *cqrel, const void *query_str) {
    return json_call([&]() -> std::string {
        if (!cqrel) throw Error("cqrel pointer is null!");
        const std::string q = accept_str("query_str", query_str);
        const QRel &qrel = *cqrel->qrel;
        if (q == "to_json") return json::dump(qrel.to_json());
        if (q == "queries") {
            json::Value arr = json::Value::array();
            for (const std::string &name : qrel.order) arr.push(json::Value::string(name));
            return json::dump(arr);
        }
        const QueryJudgments *qj = qrel.get(q);
        if (!qj) throw Error("Unknown request: " + q);
        json::Value docs = json::Value::object();
        for (const auto &kv : qj->docs) docs.set(kv.first, json::Value::number((double)kv.second));
        return json::dump(docs);
    });
}

const CResult *load_ranksvm_format(void *data_path, void *feature_names_path) {
    return result_call([&]() -> const void * {
        const std::string path = accept_str("data_path", data_path);
        std::string names;
        if (feature_names_path) names = accept_str("feature_names_path", feature_names_path);
        return new CDataset{load_ranksvm(path, feature_names_path ? &names : nullptr)};
    });
}

const CResult *dataset_query_sampling(CDataset *dataset, const void *queries_json_list) {
    return result_call([&]() -> const void * {
        const CDataset &d = need(dataset);
        const json::Value v = parse_json(accept_str("queries_json_list", queries_json_list));
        if (v.kind != json::Value::Array) throw Error("invalid type: expected a list of query ids");
        std::vector<std::string> qs;
        for (const json::Value &x : v.arr) {
            if (x.kind != json::Value::String) throw Error("invalid type: expected a string query id");
            qs.push_back(x.s);
        }
        return new CDataset{d.view.with_queries(qs)};
    });
}

const CResult *dataset_feature_sampling(CDataset *dataset, const void *feature_json_list) {
    return result_call([&]() -> const void * {
        const CDataset &d = need(dataset);
        const json::Value v = parse_json(accept_str("feature_json_list", feature_json_list));
        if (v.kind != json::Value::Array) throw Error("invalid type: expected a list of feature ids");
        std::vector<uint32_t> fs;
        for (const json::Value &x : v.arr) {
            if (x.kind != json::Value::UInt || x.u > 0xFFFFFFFFull) throw Error("invalid type: expected u32 feature id");
            fs.push_back((uint32_t)x.u);
        }
        return new CDataset{d.view.with_features(fs)};
    });
}

const void *dataset_query_json(void *dataset, void *json_cmd_str) {
    return json_call([&]() -> std::string {
        const CDataset &d = need((const CDataset *)dataset);
        const std::string cmd = accept_str("dataset_query_json", json_cmd_str);
        const DatasetView &v = d.view;
        const ParentDataset &p = *v.parent;
        if (cmd == "is_sampled") return v.sampled ? "true" : "false";
        if (cmd == "num_features") return std::to_string(v.n_dim());
        if (cmd == "num_instances") return std::to_string(v.num_instances());
        if (cmd == "feature_ids") {
            json::Value arr = json::Value::array();
            for (uint32_t f : v.feature_ids()) arr.push(json::Value::uinteger(f));
            return json::dump(arr);
        }
        if (cmd == "feature_names") {
            json::Value arr = json::Value::array();
            for (uint32_t f : v.feature_ids()) arr.push(json::Value::string(p.feature_name(f)));
            return json::dump(arr);
        }
        if (cmd == "queries") {
            json::Value arr = json::Value::array();
            for (const auto &g : v.instances_by_query()) arr.push(json::Value::string(p.query_names[g.first]));
            return json::dump(arr);
        }
        if (cmd == "instances_by_query") {
            json::Value obj = json::Value::object();
            for (const auto &g : v.instances_by_query()) {
                json::Value ids = json::Value::array();
                for (uint32_t id : g.second) ids.push(json::Value::uinteger(id));
                obj.set(p.query_names[g.first], std::move(ids));
            }
            return json::dump(obj);
        }
        json::Value err = json::Value::object();  // ffi.rs:176-179
        err.set("error", json::Value::string("unknown_dataset_query_str"));
        err.set("context", json::Value::string(cmd));
        return json::dump(err);
    });
}

const void *query_json(const void *json_cmd_str) {
    return json_call([&]() -> std::string {
        const std::string cmd = accept_str("query_json_str", json_cmd_str);
        if (cmd == "coordinate_ascent_defaults")
            return json::dump(train_request_json("CoordinateAscent", CoordinateAscentParams::defaults().to_json()));
        if (cmd == "random_forest_defaults")
            return json::dump(train_request_json("RandomForest", RandomForestParams::defaults().to_json()));
        if (cmd == "last_train_stats") {  // extension: counters of the most recent train_model
            const TrainStats s = last_train_stats();
            json::Value o = json::Value::object();
            o.set("evals_consumed", json::Value::uinteger(s.evals_consumed));
            o.set("evals_computed", json::Value::uinteger(s.evals_computed));
            o.set("sweeps", json::Value::uinteger(s.sweeps));
            o.set("global_steps", json::Value::uinteger(s.global_steps));
            o.set("sweep", json::Value::string(s.exact_sweep ? "exact" : "batched"));
            o.set("seconds_setup", json::Value::number(s.seconds_setup));
            o.set("seconds_device", json::Value::number(s.seconds_device));
            o.set("seconds_total", json::Value::number(s.seconds_total));
            o.set("kernel_launches", json::Value::uinteger(fr_dev_kernel_launches()));
            return json::dump(o);
        }
        if (cmd == "device_count") return std::to_string(fr_dev_device_count());
        json::Value err = json::Value::object();  // ffi.rs:228-231
        err.set("error", json::Value::string("unknown_query_str"));
        err.set("context", json::Value::string(cmd));
        return json::dump(err);
    });
}

const CResult *make_dense_dataset_f32_f64_i64(size_t n, size_t d, const float *x, const double *y,
                                              const int64_t *qids) {
    return result_call([&]() -> const void * { return new CDataset{make_dense(n, d, x, y, qids)}; });
}

const CResult *train_model(void *train_request_json_ptr, void *dataset) {
    return result_call([&]() -> const void * {
        const CDataset &d = need((const CDataset *)dataset);
        const json::Value req = parse_json(accept_str("train_request_json", train_request_json_ptr));
        if (req.kind != json::Value::Object) throw Error("invalid type: expected struct TrainRequest");
        const json::Value *measure = req.find("measure");
        const json::Value *params = req.find("params");
        const json::Value *judgments = req.find("judgments");
        if (!measure) throw Error("missing field `measure`");
        if (!params) throw Error("missing field `params`");
        if (!judgments) throw Error("missing field `judgments`");
        if (measure->kind != json::Value::String) throw Error("invalid type: expected a string for `measure`");
        if (params->kind != json::Value::Object || params->obj.size() != 1)
            throw Error("invalid type: expected a single-key map for enum FastRankModelParams");
        std::shared_ptr<QRel> qrel;
        if (!judgments->is_null()) qrel = QRel::from_json(*judgments);
        const Measure m = Measure::parse(measure->s);  // json_api.rs:40-44
        const std::string &kind = params->obj[0].first;
        TrainStats stats;
        const auto t_begin = std::chrono::steady_clock::now();
        auto seconds_since = [&](std::chrono::steady_clock::time_point t0) {
            return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        };
        if (kind == "CoordinateAscent") {
            CoordinateAscentParams p = CoordinateAscentParams::from_json(params->obj[0].second);
            if (const json::Value *sweep = req.find("sweep")) {  // extension; absent in the reference's requests
                if (sweep->kind != json::Value::String || (sweep->s != "exact" && sweep->s != "batched"))
                    throw Error("invalid value for `sweep`: expected \"exact\" or \"batched\"");
                p.exact_sweep = sweep->s == "exact";
            }
            Evaluator ev(d.view, m, qrel.get());
            const double setup = seconds_since(t_begin);
            CModel *out = new CModel(coordinate_ascent_learn(p, d.view, ev, &stats));
            stats.seconds_setup = setup;
            stats.seconds_total = seconds_since(t_begin);
            set_last_train_stats(stats);
            return out;
        }
        if (kind == "RandomForest") {
            const RandomForestParams p = RandomForestParams::from_json(params->obj[0].second);
            Evaluator ev(d.view, m, qrel.get());
            const double setup = seconds_since(t_begin);
            CModel *out = new CModel(random_forest_learn(p, d.view, ev, &stats));
            stats.seconds_setup = setup;
            stats.seconds_total = seconds_since(t_begin);
            set_last_train_stats(stats);
            return out;
        }
        throw Error("unknown variant `" + kind + "`, expected `CoordinateAscent` or `RandomForest`");
    });
}

const CResult *model_from_json(const void *json_str) {
    return result_call([&]() -> const void * {
        return new CModel(Model::from_json(parse_json(accept_str("json_str", json_str))));
    });
}

const void *model_query_json(const void *model, const void *json_cmd_str) {
    return json_call([&]() -> std::string {
        const CModel &m = need((const CModel *)model);
        const std::string cmd = accept_str("query_json", json_cmd_str);
        if (cmd == "to_json") return json::dump(m.model.to_json());
        json::Value err = json::Value::object();  // ffi.rs:206-209
        err.set("error", json::Value::string("unknown_dataset_query_str"));
        err.set("context", json::Value::string(cmd));
        return json::dump(err);
    });
}

const void *evaluate_by_query(const CModel *model, const CDataset *dataset, const CQRel *qrel,
                              const void *evaluator) {
    return json_call([&]() -> std::string {
        const CModel &m = need(model);
        const CDataset &d = need(dataset);
        const Measure measure = Measure::parse(accept_str("evaluator_name", evaluator));
        Evaluator ev(d.view, measure, qrel ? qrel->qrel.get() : nullptr);  // ffi.rs:253-255
        std::vector<double> per_query;
        ev.evaluate_mean(m.model, &per_query, m.uid);
        json::Value out = json::Value::object();
        const auto &qs = ev.view_queries();
        for (size_t k = 0; k < qs.size(); ++k)
            out.set(d.view.parent->query_names[qs[k]], json::Value::number(per_query[k]));
        return json::dump(out);
    });
}

const void *predict_scores(const CModel *model, const CDataset *dataset) {
    return json_call([&]() -> std::string {  // json_api.rs:53-72
        const CModel &m = need(model);
        const CDataset &d = need(dataset);
        const std::vector<double> scores = score_all(m, *d.view.parent);
        std::string out = "{";
        bool first = true;
        auto emit = [&](uint32_t id) {
            if (scores[id] != scores[id]) throw Error("Model.predict -> NaN");
            if (!first) out.push_back(',');
            first = false;
            out.push_back('"');
            out += std::to_string(id);
            out += "\":";
            json::write_double(out, scores[id]);
        };
        if (d.view.sampled) {
            for (uint32_t id : d.view.instances) emit(id);
        } else {
            for (size_t id = 0; id < d.view.parent->n; ++id) emit((uint32_t)id);
        }
        out.push_back('}');
        return out;
    });
}

const void *predict_dense_f64(const CModel *model, const CDataset *dataset, double *out, size_t n_out) {
    try {
        const CModel &m = need(model);
        const CDataset &d = need(dataset);
        if (!out) throw Error("NULL pointer: out");
        ParentDataset &p = *d.view.parent;
        if (n_out < p.n) throw Error("predict_dense_f64: output buffer smaller than the parent dataset");
        if (d.view.sampled) {
            const std::vector<double> scores = score_all(m, p);
            for (size_t i = 0; i < n_out; ++i) out[i] = NAN;
            for (uint32_t id : d.view.instances) out[id] = scores[id];
        } else {
            score_all_into(m, p, out);
            for (size_t i = p.n; i < n_out; ++i) out[i] = NAN;
        }
        return nullptr;
    } catch (const std::exception &e) {
        return dup_cstr(error_json(e.what()));
    }
}

const void *evaluate_mean_f64(const CModel *model, const CDataset *dataset, const CQRel *qrel,
                              const void *evaluator, double *out_mean) {
    try {
        const CModel &m = need(model);
        const CDataset &d = need(dataset);
        if (!out_mean) throw Error("NULL pointer: out_mean");
        const Measure measure = Measure::parse(accept_str("evaluator_name", evaluator));
        Evaluator ev(d.view, measure, qrel ? qrel->qrel.get() : nullptr);
        *out_mean = ev.evaluate_mean(m.model, nullptr, m.uid);
        return nullptr;
    } catch (const std::exception &e) {
        return dup_cstr(error_json(e.what()));
    }
}

const void *evaluate_bootstrap_f64(const CModel *model, const CDataset *dataset, const CQRel *qrel,
                                   const void *evaluator, uint32_t num_trials, double *out_means) {
    try {
        const CModel &m = need(model);
        const CDataset &d = need(dataset);
        if (!out_means) throw Error("NULL pointer: out_means");
        const Measure measure = Measure::parse(accept_str("evaluator_name", evaluator));
        Evaluator ev(d.view, measure, qrel ? qrel->qrel.get() : nullptr);
        std::vector<double> per_query;
        ev.evaluate_mean(m.model, &per_query, m.uid);  // evaluate_to_vec, evaluators.rs:158
        if (per_query.empty()) throw Error("bootstrap_eval: the view has no queries");
        if (fr_dev_plan_bootstrap(ev.plan(), per_query.data(), per_query.size(), 0xdeadbeefull, num_trials, out_means))
            throw Error(fr_dev_last_error());
        for (uint32_t t = 0; t < num_trials; ++t)
            if (out_means[t] != out_means[t]) throw Error("PercentileStats::NaN");  // stats.rs:134
        std::sort(out_means, out_means + num_trials);  // stats.rs:136
        return nullptr;
    } catch (const std::exception &e) {
        return dup_cstr(error_json(e.what()));
    }
}

const void *dataset_device_profile(const CDataset *dataset, int enable, uint64_t *out_launches,
                                   double *out_total_ms) {
    try {
        const CDataset &d = need(dataset);
        ParentDataset &p = *d.view.parent;
        fr_dev_dataset *dev = p.device();
        std::lock_guard<std::recursive_mutex> lock(p.use_mu);
        if (enable >= 0 && fr_dev_profile_enable(dev, enable)) throw Error(fr_dev_last_error());
        if (out_launches || out_total_ms) {
            uint64_t n = 0;
            double ms = 0.0;
            if (fr_dev_profile_read(dev, &n, &ms, 1)) throw Error(fr_dev_last_error());
            if (out_launches) *out_launches = n;
            if (out_total_ms) *out_total_ms = ms;
        }
        return nullptr;
    } catch (const std::exception &e) {
        return dup_cstr(error_json(e.what()));
    }
}

const void *predict_to_trecrun(const CModel *model, const CDataset *dataset, const void *output_path,
                               const void *system_name, size_t depth) {
    return json_call([&]() -> std::string {  // json_api.rs:75-120
        const CModel &m = need(model);
        const CDataset &d = need(dataset);
        const std::string path = accept_str("output_path", output_path);
        const std::string system = accept_str("system_name", system_name);
        ParentDataset &p = *d.view.parent;
        std::ostringstream out;  // written at the end, compressed by extension (io_helper.rs:31-48)
        const std::vector<double> scores = score_all(m, p);
        size_t written = 0;
        // Export path: scores come from the GPU; ordering the rows of the
        // text file is done here with the reference comparator (evaluators.rs:33-49).
        for (const auto &g : d.view.instances_by_query()) {
            std::vector<uint32_t> ids = g.second;
            std::sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) {
                if (scores[a] != scores[b]) return scores[a] > scores[b];
                if (p.gains[a] != p.gains[b]) return p.gains[a] < p.gains[b];
                return a < b;
            });
            size_t rank = 0;
            for (uint32_t id : ids) {
                ++rank;
                if (depth > 0 && rank > depth) break;
                if (p.has_docid.empty() || !p.has_docid[id])
                    throw Error("Dataset does not contain document ids and therefore cannot save to trecrun!");
                out << p.query_names[g.first] << " Q0 " << p.docids[id] << ' ' << rank << ' '
                    << rust_display_f64(scores[id]) << ' ' << system << '\n';
                ++written;
            }
        }
        write_file_by_extension(path, out.str());
        return std::to_string(written);
    });
}

}  // extern "C"
