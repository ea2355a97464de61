// dataset.cpp -- RNG, judgments and dataset bookkeeping on the host.
//
// Restates the observable behaviour of the reference's dense_dataset.rs, dataset.rs,
// instance.rs, libsvm.rs, qrel.rs and sampling.rs for the C ABI; the feature matrix itself
// is shipped to the GPU once (ParentDataset::device) and never touched on the host again.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>

#include "host.hpp"

namespace frb {

// ---------------------------------------------------------------------------------------
// Rand64
// ---------------------------------------------------------------------------------------
namespace {
typedef unsigned __int128 u128;
inline u128 make_u128(uint64_t hi, uint64_t lo) { return (((u128)hi) << 64) | (u128)lo; }
const u128 kPcgMultiplier = make_u128(0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull);
const u128 kPcgDefaultIncrement = make_u128(0x2FE0E169FFBD06E3ull, 0x5BC307BD4D2F814Full);
}  // namespace

Rand64::Rand64(unsigned __int128 seed) : state_(0), inc_((kPcgDefaultIncrement << 1) | 1) {
    (void)rand_u64();
    state_ += seed;
    (void)rand_u64();
}

uint64_t Rand64::rand_u64() {
    const u128 old = state_;
    state_ = old * kPcgMultiplier + inc_;
    // oorandom 11.1.0 (the version Cargo.toml:18-19 pins): xorshift the 128-bit state by 29,
    // keep bits 58..121, rotate right by the top six bits.
    const unsigned rot = (unsigned)(old >> 122);
    const uint64_t xsh = (uint64_t)(((old >> 29) ^ old) >> 58);
    return (xsh >> rot) | (xsh << ((64 - rot) & 63));
}

double Rand64::rand_float() {
    const uint64_t bits = rand_u64() >> 10;  // 54 bits, as oorandom keeps MANTISSA_DIGITS + 1
    return (double)bits * (1.0 / 18014398509481984.0);
}

uint64_t Rand64::rand_range(uint64_t start, uint64_t end) {
    const uint64_t span = end - start;
    u128 m = (u128)rand_u64() * (u128)span;
    uint64_t low = (uint64_t)m;
    if (low < span) {
        const uint64_t threshold = (0 - span) % span;
        while (low < threshold) {
            m = (u128)rand_u64() * (u128)span;
            low = (uint64_t)m;
        }
    }
    // 11.1.0's Rand64 returns the draw over [0, end-start) without adding `start`; every seeded
    // model of the reference (and its goldens, random_forest.rs:462) depends on that.
    return (uint64_t)(m >> 64);
}

// ---------------------------------------------------------------------------------------
// QRel
// ---------------------------------------------------------------------------------------
void QueryJudgments::insert(const std::string &doc, float gain) {
    auto it = index.find(doc);
    if (it != index.end()) {
        docs[it->second].second = gain;  // HashMap::insert overwrites (qrel.rs:89-92)
        return;
    }
    index.emplace(doc, docs.size());
    docs.emplace_back(doc, gain);
}

uint32_t QueryJudgments::num_relevant() const {
    uint32_t n = 0;
    for (const auto &kv : docs) n += kv.second > 0.0f;
    return n;
}

std::vector<float> QueryJudgments::gain_vector() const {
    std::vector<float> out;
    for (const auto &kv : docs)
        if (kv.second > 0.0f) out.push_back(kv.second);
    return out;
}

const QueryJudgments *QRel::get(const std::string &qid) const {
    auto it = queries.find(qid);
    return it == queries.end() ? nullptr : &it->second;
}

QueryJudgments &QRel::get_or_create(const std::string &qid) {
    auto it = queries.find(qid);
    if (it == queries.end()) {
        order.push_back(qid);
        it = queries.emplace(qid, QueryJudgments()).first;
    }
    return it->second;
}

static float parse_f32_strict(const std::string &tok, bool *ok) {
    errno = 0;
    char *endp = nullptr;
    float v = strtof(tok.c_str(), &endp);
    *ok = !tok.empty() && endp == tok.c_str() + tok.size();
    return v;
}

std::shared_ptr<QRel> QRel::load_file(const std::string &path) {
    std::istringstream in(read_file_by_extension(path));  // .gz / .bz2 / .zst by extension (io_helper.rs:18-29)
    auto out = std::make_shared<QRel>();
    std::string line;
    size_t num = 0;
    while (std::getline(in, line)) {
        ++num;
        std::istringstream ss(line);
        std::vector<std::string> row;
        std::string tok;
        while (ss >> tok) row.push_back(tok);
        if (row.empty()) continue;
        if (row.size() < 4) throw Error(path + ":" + std::to_string(num) + ": expected `qid unused docid gain`");
        bool ok = false;
        const float gain = parse_f32_strict(row[3], &ok);
        if (!ok) throw Error(path + ":" + std::to_string(num) + ": Invalid relevance judgment " + row[3]);
        if (gain != gain) throw Error(path + ":" + std::to_string(num) + ": NaN relevance judgment.");
        out->get_or_create(row[0]).insert(row[2], gain);
    }
    return out;
}

std::shared_ptr<QRel> QRel::from_json(const json::Value &v) {
    if (v.kind != json::Value::Object) throw Error("invalid type: expected a map of query -> {doc: gain}");
    auto out = std::make_shared<QRel>();
    for (const auto &qm : v.obj) {
        if (qm.second.kind != json::Value::Object) throw Error("invalid type: judgments of query " + qm.first + " must be a map");
        QueryJudgments &qj = out->get_or_create(qm.first);
        for (const auto &dm : qm.second.obj) {
            if (!dm.second.is_number()) throw Error("invalid type: gain of " + dm.first + " must be a number");
            const float g = (float)dm.second.as_double();
            if (g != g) throw Error("NaN relevance judgment.");
            qj.insert(dm.first, g);
        }
    }
    return out;
}

json::Value QRel::to_json() const {
    json::Value root = json::Value::object();
    for (const std::string &qid : order) {
        const QueryJudgments &qj = queries.at(qid);
        json::Value docs = json::Value::object();
        for (const auto &kv : qj.docs) docs.set(kv.first, json::Value::number((double)kv.second));
        root.set(qid, std::move(docs));
    }
    return root;
}

// ---------------------------------------------------------------------------------------
// ParentDataset
// ---------------------------------------------------------------------------------------
ParentDataset::~ParentDataset() {
    plan_cache.clear();  // plans reference the device dataset
    for (auto &kv : model_cache) fr_dev_model_destroy(kv.second);
    if (dev) fr_dev_dataset_destroy(dev);
}

fr_dev_model *ParentDataset::device_model(const Model &m, uint64_t uid, bool *owned) {
    std::lock_guard<std::recursive_mutex> lock(use_mu);
    *owned = uid == 0;
    if (uid != 0) {
        for (size_t i = 0; i < model_cache.size(); ++i) {
            if (model_cache[i].first == uid) {
                auto hit = model_cache[i];
                model_cache.erase(model_cache.begin() + (long)i);
                model_cache.push_back(hit);
                return hit.second;
            }
        }
    }
    const std::vector<uint64_t> code = m.lower();
    fr_dev_model *dm = nullptr;
    if (fr_dev_model_create(device(), code.data(), code.size(), &dm)) throw Error(fr_dev_last_error());
    if (uid != 0) {
        if (model_cache.size() >= 4) {
            fr_dev_model_destroy(model_cache.front().second);
            model_cache.erase(model_cache.begin());
        }
        model_cache.emplace_back(uid, dm);
    }
    return dm;
}

fr_dev_dataset *ParentDataset::device() {
    std::lock_guard<std::mutex> lock(dev_mu);
    if (!dev) {
        int which = 0;
        if (const char *env = getenv("FASTRANK_DEVICE")) which = atoi(env);
        else if (const char *lr = getenv("LOCAL_RANK")) which = atoi(lr) % std::max(1, fr_dev_device_count());
        fr_dev_dataset *out = nullptr;
        if (fr_dev_dataset_create(which, n, d, x, gains.data(), query_of.data(),
                                  (uint32_t)query_names.size(), &out))
            throw Error(std::string("GPU dataset upload failed: ") + fr_dev_last_error());
        dev = out;
    }
    return dev;
}

std::string ParentDataset::feature_name(uint32_t fid) const {
    auto it = feature_names.find(fid);
    return it == feature_names.end() ? std::to_string(fid) : it->second;
}

static void index_queries(ParentDataset &ds, const std::vector<std::string> &qid_per_instance) {
    ds.query_of.resize(ds.n);
    for (size_t i = 0; i < ds.n; ++i) {
        const std::string &q = qid_per_instance[i];
        auto it = ds.query_lookup.find(q);
        if (it == ds.query_lookup.end()) {
            it = ds.query_lookup.emplace(q, (uint32_t)ds.query_names.size()).first;
            ds.query_names.push_back(q);
            ds.by_query.emplace_back();
        }
        ds.query_of[i] = it->second;
        ds.by_query[it->second].push_back((uint32_t)i);
    }
}

// libsvm.rs:131-189 (one line), instance.rs:104-130 (densify), dataset.rs:211-256 (collect)
DatasetView load_ranksvm(const std::string &path, const std::string *feature_names_path) {
    auto ds = std::make_shared<ParentDataset>();
    if (feature_names_path) {  // dataset.rs:14-25
        std::ifstream fin(*feature_names_path);
        if (!fin) throw Error("Os { code: 2, kind: NotFound, message: \"No such file or directory\" }: " + *feature_names_path);
        std::stringstream buf;
        buf << fin.rdbuf();
        json::Value names;
        try {
            names = json::parse(buf.str());
        } catch (const json::ParseError &e) {
            throw Error(std::string("feature names: ") + e.what());
        }
        if (names.kind != json::Value::Object) throw Error("feature names: expected a map of feature id -> name");
        for (const auto &m : names.obj) {
            char *endp = nullptr;
            unsigned long long id = strtoull(m.first.c_str(), &endp, 10);
            if (m.first.empty() || *endp != 0) throw Error("ParseIntError { kind: InvalidDigit }");
            if (m.second.kind != json::Value::String) throw Error("feature names: names must be strings");
            ds->feature_names[(uint32_t)id] = m.second.s;
        }
    }
    std::istringstream in(read_file_by_extension(path));  // .gz / .bz2 / .zst by extension (io_helper.rs:18-29)

    struct Row {
        std::vector<std::pair<uint32_t, float>> feats;
    };
    std::vector<Row> rows;
    std::vector<std::string> qids;
    std::set<uint32_t> feature_set;
    std::string line;
    size_t line_num = 0;
    bool any_docid = false;
    while (std::getline(in, line)) {
        ++line_num;
        auto bad = [&](const std::string &what) {
            return Error(path + ": LineParseError(" + std::to_string(line_num) + ", " + what + ")");
        };
        std::string comment;
        bool has_comment = false;
        const size_t hash = line.find('#');
        std::string data = line;
        if (hash != std::string::npos) {
            comment = line.substr(hash + 1);
            const size_t b = comment.find_first_not_of(" \t\r\n");
            const size_t e = comment.find_last_not_of(" \t\r\n");
            comment = b == std::string::npos ? "" : comment.substr(b, e - b + 1);
            has_comment = true;
            data = line.substr(0, hash);
        }
        std::istringstream ss(data);
        std::vector<std::string> toks;
        std::string tok;
        while (ss >> tok) toks.push_back(tok);
        if (toks.empty()) continue;
        errno = 0;
        char *endp = nullptr;
        const double label64 = strtod(toks[0].c_str(), &endp);
        if (endp != toks[0].c_str() + toks[0].size()) throw bad("Label(ParseFloatError { kind: Invalid })");
        const float label = (float)label64;
        if (label != label) throw bad("LabelIsNan(FloatIsNan)");
        size_t k = 1;
        std::string qid;
        bool has_qid = false;
        if (k < toks.size() && toks[k].compare(0, 4, "qid:") == 0) {
            qid = toks[k];
            while (qid.compare(0, 4, "qid:") == 0) qid = qid.substr(4);
            has_qid = true;
            ++k;
        }
        Row row;
        for (; k < toks.size(); ++k) {
            const size_t colon = toks[k].find(':');
            if (colon == std::string::npos) throw bad("FeatureNoColon");
            const std::string fs = toks[k].substr(0, colon), vs = toks[k].substr(colon + 1);
            char *e1 = nullptr;
            errno = 0;
            const unsigned long long fid = strtoull(fs.c_str(), &e1, 10);
            if (fs.empty() || *e1 != 0 || fs[0] == '-' || fs[0] == '+' || fid > 0xFFFFFFFFull || errno == ERANGE)
                throw bad("FeatureNum(ParseIntError)");
            bool ok = false;
            const float val = parse_f32_strict(vs, &ok);
            if (!ok) throw bad("FeatureValNotFloat(Error)");
            row.feats.emplace_back((uint32_t)fid, val);
        }
        if (row.feats.empty()) throw bad("instance without features");
        bool sorted = true;
        for (size_t a = 0; a + 1 < row.feats.size(); ++a)
            if (row.feats[a].first >= row.feats[a + 1].first) sorted = false;
        if (!sorted) {
            std::stable_sort(row.feats.begin(), row.feats.end(),
                             [](const auto &l, const auto &r) { return l.first < r.first; });
            for (size_t a = 0; a + 1 < row.feats.size(); ++a)
                if (row.feats[a].first == row.feats[a + 1].first) throw bad("MultipleDefinitions");
        }
        if (!has_qid) throw Error(path + ": Missing qid");
        // instance.rs:106-122: density >= 0.5 -> dense array of max_feature + 1 entries
        // (feature ids 0..=max become "present"), else only the listed ids are present
        const uint32_t max_feature = row.feats.back().first;
        const double density = (double)row.feats.size() / (double)max_feature;
        uint32_t len;
        if (density >= 0.5) {
            len = max_feature + 1;
            for (uint32_t f = 0; f <= max_feature; ++f) feature_set.insert(f);
        } else {
            len = 0;  // sparse instance
            std::vector<uint32_t> &ids = ds->sparse_ids[(uint32_t)rows.size()];
            for (const auto &fv : row.feats) {
                feature_set.insert(fv.first);
                ids.push_back(fv.first);
            }
        }
        ds->row_len.push_back(len);
        ds->gains.push_back(label);
        qids.push_back(qid);
        ds->docids.push_back(comment);
        ds->has_docid.push_back(has_comment ? 1 : 0);
        any_docid |= has_comment;
        rows.push_back(std::move(row));
    }
    if (rows.empty()) throw Error(path + ": No features defined!");
    if (!any_docid) {
        ds->docids.clear();
        ds->has_docid.clear();
    }
    ds->n = rows.size();
    ds->features.assign(feature_set.begin(), feature_set.end());
    ds->d = (size_t)ds->features.back() + 1;
    // the device layout is dense (instance.rs keeps Sparse32 rows; the kernels stream a matrix)
    if (ds->n * ds->d > ((size_t)1 << 36))
        throw Error(path + ": " + std::to_string(ds->n) + " x " + std::to_string(ds->d) +
                    " values do not fit the dense device layout of this build");
    ds->owned_x.assign(ds->n * ds->d, 0.0f);
    for (size_t i = 0; i < ds->n; ++i)
        for (const auto &fv : rows[i].feats) ds->owned_x[i * ds->d + fv.first] = fv.second;
    ds->x = ds->owned_x.data();
    index_queries(*ds, qids);
    DatasetView view;
    view.parent = ds;
    return view;
}

// dense_dataset.rs:28-55
DatasetView make_dense(size_t n, size_t d, const float *x, const double *y, const int64_t *qids) {
    if (!x || !y || !qids) throw Error("NULL pointer: make_dense_dataset_f32_f64_i64");
    if (n == 0 || d == 0) throw Error("make_dense_dataset_f32_f64_i64: empty matrix");
    auto ds = std::make_shared<ParentDataset>();
    ds->n = n;
    ds->d = d;
    ds->x = x;
    ds->dense_source = true;
    ds->gains.resize(n);
    ds->query_of.resize(n);
    // query ids are interned as integers (dense_dataset.rs:33-47 keeps u32 ids and a name table);
    // consecutive rows of one query -- the usual layout -- skip the hash lookup
    std::unordered_map<int64_t, uint32_t> number_of;
    int64_t last_qid = -1;
    uint32_t last_number = 0;
    for (size_t i = 0; i < n; ++i) {
        const int64_t q = qids[i];
        if (q < 0 || q > 0xFFFFFFFFll) throw Error("TryFromIntError(())");
        if (y[i] != y[i]) throw Error("NaN in ys[" + std::to_string(i) + "]");
        ds->gains[i] = (float)y[i];  // dense_dataset.rs:120
        if (q != last_qid) {
            auto it = number_of.find(q);
            if (it == number_of.end()) {
                it = number_of.emplace(q, (uint32_t)ds->query_names.size()).first;
                ds->query_names.push_back(std::to_string(q));
                ds->query_lookup.emplace(ds->query_names.back(), it->second);
                ds->by_query.emplace_back();
            }
            last_qid = q;
            last_number = it->second;
        }
        ds->query_of[i] = last_number;
        ds->by_query[last_number].push_back((uint32_t)i);
    }
    ds->features.resize(d);
    for (size_t j = 0; j < d; ++j) ds->features[j] = (uint32_t)j;
    DatasetView view;
    view.parent = ds;
    return view;
}

// ---------------------------------------------------------------------------------------
// DatasetView
// ---------------------------------------------------------------------------------------
std::vector<uint32_t> DatasetView::feature_ids() const { return sampled ? features : parent->features; }

uint32_t DatasetView::n_dim() const {
    return sampled ? (uint32_t)features.size() : (uint32_t)parent->d;
}

size_t DatasetView::num_instances() const { return sampled ? instances.size() : parent->n; }

std::vector<std::pair<uint32_t, std::vector<uint32_t>>> DatasetView::instances_by_query() const {
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> out;
    if (!sampled) {
        out.reserve(parent->by_query.size());
        for (uint32_t q = 0; q < parent->by_query.size(); ++q) out.emplace_back(q, parent->by_query[q]);
        return out;
    }
    // dataset.rs:132-140: group the kept instances by their parent query
    std::vector<int64_t> slot(parent->query_names.size(), -1);
    for (uint32_t id : instances) {
        const uint32_t q = parent->query_of[id];
        if (slot[q] < 0) {
            slot[q] = (int64_t)out.size();
            out.emplace_back(q, std::vector<uint32_t>());
        }
        out[(size_t)slot[q]].second.push_back(id);
    }
    std::sort(out.begin(), out.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    return out;
}

DatasetView DatasetView::with_queries(const std::vector<std::string> &queries) const {
    std::set<std::string> wanted(queries.begin(), queries.end());
    DatasetView out;
    out.parent = parent;
    out.sampled = true;
    out.features = feature_ids();
    for (const auto &qi : instances_by_query())
        if (wanted.count(parent->query_names[qi.first]))
            out.instances.insert(out.instances.end(), qi.second.begin(), qi.second.end());
    return out;
}

DatasetView DatasetView::with_features(const std::vector<uint32_t> &fids) const {
    const std::vector<uint32_t> valid_list = feature_ids();
    std::set<uint32_t> valid(valid_list.begin(), valid_list.end()), keep, missing;
    for (uint32_t f : fids) (valid.count(f) ? keep : missing).insert(f);
    if (!missing.empty()) {
        std::string msg = "Missing Features: {";
        bool first = true;
        for (uint32_t f : missing) {
            if (!first) msg += ", ";
            msg += "FeatureId(" + std::to_string(f) + ")";
            first = false;
        }
        throw Error(msg + "}");
    }
    if (keep.empty()) throw Error("No Features!");
    DatasetView out;
    out.parent = parent;
    out.sampled = true;
    out.features.assign(keep.begin(), keep.end());
    if (sampled) {
        out.instances = instances;
    } else {
        out.instances.resize(parent->n);
        for (size_t i = 0; i < parent->n; ++i) out.instances[i] = (uint32_t)i;
    }
    return out;
}

DatasetView DatasetView::with_instances(std::vector<uint32_t> ids) const {
    DatasetView out;
    out.parent = parent;
    out.sampled = true;
    out.features = feature_ids();
    out.instances = std::move(ids);
    return out;
}

}  // namespace frb
