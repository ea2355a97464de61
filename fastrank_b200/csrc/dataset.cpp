// dataset.cpp -- RNG, judgments and dataset bookkeeping on the host.
//
// Restates the observable behaviour of the reference's dense_dataset.rs, dataset.rs,
// instance.rs, libsvm.rs, qrel.rs and sampling.rs for the C ABI; the feature matrix itself
// is shipped to the GPU once (ParentDataset::device) and never touched on the host again.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <thread>

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
        // libsvm data: which feature ids each row carries -- a bitmap when some rows are Sparse32
        // (the listed ids), else the rows' lengths (Dense32: the leading ids)
        if (!dense_source && !sparse_ids.empty() && row_len.size() == n) {
            const uint32_t words = (uint32_t)((d + 31) / 32);
            std::vector<uint32_t> bits(n * (size_t)words, 0u);
            for (size_t i = 0; i < n; ++i) {
                uint32_t *row = bits.data() + i * words;
                if (row_len[i] > 0) {
                    for (uint32_t f = 0; f < row_len[i] && f < d; ++f) row[f >> 5] |= 1u << (f & 31u);
                } else {
                    auto it = sparse_ids.find((uint32_t)i);
                    if (it != sparse_ids.end())
                        for (uint32_t f : it->second)
                            if (f < d) row[f >> 5] |= 1u << (f & 31u);
                }
            }
            if (fr_dev_dataset_set_row_presence(out, bits.data(), words)) {
                const std::string why = fr_dev_last_error();
                fr_dev_dataset_destroy(out);
                throw Error("GPU dataset upload failed: " + why);
            }
        } else if (!dense_source && row_len.size() == n) {
            bool ragged = false;
            for (uint32_t len : row_len) ragged |= len < d;
            if (ragged && fr_dev_dataset_set_row_lengths(out, row_len.data())) {
                const std::string why = fr_dev_last_error();
                fr_dev_dataset_destroy(out);
                throw Error("GPU dataset upload failed: " + why);
            }
        }
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
    // The whole (decompressed, io_helper.rs:18-29) file is parsed from memory: it is cut at line
    // boundaries into one slice per thread, every slice is parsed independently, and the slices are
    // merged in file order, so instance ids, query order and the first error reported are those of
    // a sequential read (libsvm.rs:205-260).
    const bool trace = getenv("FASTRANK_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (trace)
            fprintf(stderr, "[fastrank_b200] load_ranksvm   %-14s +%.1f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    const std::string text = read_file_by_extension(path);
    lap("read");
    struct Row {
        std::vector<std::pair<uint32_t, float>> feats;
        std::string qid, comment;
        float label = 0.f;
        bool has_comment = false;
    };
    struct Slice {
        size_t begin = 0, end = 0, first_line = 0;
        std::vector<Row> rows;
        bool failed = false;
        std::string error;
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t n_slices = std::max<size_t>(1, std::min<size_t>(std::min(hw, 16u), text.size() / ((size_t)1 << 20)));
    std::vector<Slice> slices(n_slices);
    {
        size_t at = 0;
        for (size_t k = 0; k < n_slices; ++k) {
            slices[k].begin = at;
            size_t want = k + 1 == n_slices ? text.size() : text.size() * (k + 1) / n_slices;
            if (want < at) want = at;
            if (k + 1 < n_slices) {
                const size_t nl = text.find('\n', want);
                want = nl == std::string::npos ? text.size() : nl + 1;
            }
            slices[k].end = want;
            at = want;
        }
        size_t line = 0;
        for (size_t k = 0; k < n_slices; ++k) {
            slices[k].first_line = line;
            line += (size_t)std::count(text.begin() + (long)slices[k].begin, text.begin() + (long)slices[k].end, '\n');
        }
    }
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    auto parse_slice = [&](Slice &sl) {
        size_t line_num = sl.first_line;
        const char *base = text.data();
        size_t pos = sl.begin;
        std::vector<std::pair<const char *, const char *>> toks;
        std::string scratch;
        while (pos < sl.end) {
            size_t eol = text.find('\n', pos);
            if (eol == std::string::npos || eol > sl.end) eol = sl.end;
            const char *lb = base + pos, *le = base + eol;
            pos = eol + 1;
            ++line_num;
            auto bad = [&](const std::string &what) {
                sl.failed = true;
                sl.error = path + ": LineParseError(" + std::to_string(line_num) + ", " + what + ")";
            };
            Row row;
            const char *hash = (const char *)memchr(lb, '#', (size_t)(le - lb));
            const char *de = le;
            if (hash) {
                const char *cb = hash + 1, *ce = le;
                while (cb < ce && is_space(*cb)) ++cb;
                while (ce > cb && is_space(ce[-1])) --ce;
                row.comment.assign(cb, ce);
                row.has_comment = true;
                de = hash;
            }
            toks.clear();
            for (const char *c = lb; c < de;) {
                while (c < de && is_space(*c)) ++c;
                if (c >= de) break;
                const char *t0 = c;
                while (c < de && !is_space(*c)) ++c;
                toks.emplace_back(t0, c);
            }
            if (toks.empty()) continue;
            auto cstr = [&](const char *b, const char *e) -> const char * {
                scratch.assign(b, e);
                return scratch.c_str();
            };
            {
                const char *z = cstr(toks[0].first, toks[0].second);
                char *endp = nullptr;
                const double label64 = strtod(z, &endp);
                if (endp != z + scratch.size()) return bad("Label(ParseFloatError { kind: Invalid })");
                row.label = (float)label64;
                if (row.label != row.label) return bad("LabelIsNan(FloatIsNan)");
            }
            size_t k = 1;
            bool has_qid = false;
            if (k < toks.size() && toks[k].second - toks[k].first >= 4 && memcmp(toks[k].first, "qid:", 4) == 0) {
                const char *qb = toks[k].first;
                while (toks[k].second - qb >= 4 && memcmp(qb, "qid:", 4) == 0) qb += 4;
                row.qid.assign(qb, toks[k].second);
                has_qid = true;
                ++k;
            }
            row.feats.reserve(toks.size() - k);
            for (; k < toks.size(); ++k) {
                const char *tb = toks[k].first, *te = toks[k].second;
                const char *colon = (const char *)memchr(tb, ':', (size_t)(te - tb));
                if (!colon) return bad("FeatureNoColon");
                unsigned long long fid = 0;
                bool fid_ok = colon > tb;
                for (const char *c = tb; c < colon && fid_ok; ++c) {
                    if (*c < '0' || *c > '9') fid_ok = false;
                    else fid = fid * 10 + (unsigned)(*c - '0');
                    if (fid > 0xFFFFFFFFull) fid_ok = false;
                }
                if (!fid_ok) return bad("FeatureNum(ParseIntError)");
                // plain decimals go through from_chars (correctly rounded, like the reference's
                // fast-float); whatever it does not take whole ("+1", "inf", hex ...) is strtof's call
                float val = 0.f;
                const auto fc = std::from_chars(colon + 1, te, val);
                if (fc.ec != std::errc() || fc.ptr != te) {
                    const char *z = cstr(colon + 1, te);
                    char *endp = nullptr;
                    val = strtof(z, &endp);
                    if (scratch.empty() || endp != z + scratch.size()) return bad("FeatureValNotFloat(Error)");
                }
                row.feats.emplace_back((uint32_t)fid, val);
            }
            if (row.feats.empty()) return bad("instance without features");
            bool sorted = true;
            for (size_t a = 0; a + 1 < row.feats.size(); ++a)
                if (row.feats[a].first >= row.feats[a + 1].first) sorted = false;
            if (!sorted) {
                std::stable_sort(row.feats.begin(), row.feats.end(),
                                 [](const auto &l, const auto &r) { return l.first < r.first; });
                for (size_t a = 0; a + 1 < row.feats.size(); ++a)
                    if (row.feats[a].first == row.feats[a + 1].first) return bad("MultipleDefinitions");
            }
            if (!has_qid) {
                sl.failed = true;
                sl.error = path + ": Missing qid";
                return;
            }
            sl.rows.push_back(std::move(row));
        }
    };
    if (n_slices == 1) {
        parse_slice(slices[0]);
    } else {
        std::vector<std::thread> pool;
        for (size_t k = 0; k < n_slices; ++k) pool.emplace_back([&, k]() { parse_slice(slices[k]); });
        for (auto &t : pool) t.join();
    }
    for (const Slice &sl : slices)  // the first error in file order, as a sequential reader would hit it
        if (sl.failed) throw Error(sl.error);
    lap("parsed");

    size_t n_rows = 0;
    for (const Slice &sl : slices) n_rows += sl.rows.size();
    if (n_rows == 0) throw Error(path + ": No features defined!");
    std::vector<const Row *> rows;
    rows.reserve(n_rows);
    for (const Slice &sl : slices)
        for (const Row &r : sl.rows) rows.push_back(&r);
    std::vector<std::string> qids(n_rows);
    std::set<uint32_t> feature_set;
    bool any_docid = false, any_dense = false;
    uint32_t dense_max = 0;
    ds->row_len.resize(n_rows);
    ds->gains.resize(n_rows);
    ds->docids.resize(n_rows);
    ds->has_docid.resize(n_rows);
    for (size_t i = 0; i < n_rows; ++i) {
        const Row &row = *rows[i];
        // instance.rs:106-122: density >= 0.5 -> dense array of max_feature + 1 entries
        // (feature ids 0..=max become "present"), else only the listed ids are present
        const uint32_t max_feature = row.feats.back().first;
        const double density = (double)row.feats.size() / (double)max_feature;
        if (density >= 0.5) {
            ds->row_len[i] = max_feature + 1;
            dense_max = any_dense ? std::max(dense_max, max_feature) : max_feature;
            any_dense = true;
        } else {
            ds->row_len[i] = 0;  // sparse instance
            std::vector<uint32_t> &ids = ds->sparse_ids[(uint32_t)i];
            for (const auto &fv : row.feats) {
                feature_set.insert(fv.first);
                ids.push_back(fv.first);
            }
        }
        ds->gains[i] = row.label;
        qids[i] = row.qid;
        ds->docids[i] = row.comment;
        ds->has_docid[i] = row.has_comment ? 1 : 0;
        any_docid |= row.has_comment;
    }
    if (any_dense)
        for (uint32_t f = 0; f <= dense_max; ++f) feature_set.insert(f);
    if (!any_docid) {
        ds->docids.clear();
        ds->has_docid.clear();
    }
    ds->n = n_rows;
    ds->features.assign(feature_set.begin(), feature_set.end());
    ds->d = (size_t)ds->features.back() + 1;
    // the device layout is dense (instance.rs keeps Sparse32 rows; the kernels stream a matrix)
    if (ds->n * ds->d > ((size_t)1 << 36))
        throw Error(path + ": " + std::to_string(ds->n) + " x " + std::to_string(ds->d) +
                    " values do not fit the dense device layout of this build");
    lap("merged");
    ds->owned_x.assign(ds->n * ds->d, 0.0f);
    {
        auto fill = [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i)
                for (const auto &fv : rows[i]->feats) ds->owned_x[i * ds->d + fv.first] = fv.second;
        };
        const size_t workers = n_rows >= 200000 ? std::min<size_t>(hw, 8) : 1;
        if (workers <= 1) {
            fill(0, n_rows);
        } else {
            std::vector<std::thread> pool;
            for (size_t w = 0; w < workers; ++w) pool.emplace_back(fill, n_rows * w / workers, n_rows * (w + 1) / workers);
            for (auto &t : pool) t.join();
        }
    }
    ds->x = ds->owned_x.data();
    lap("densified");
    index_queries(*ds, qids);
    lap("indexed");
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
