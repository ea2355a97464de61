// training.cpp -- learner parameters as JSON (json_api.rs:13-34, coordinate_ascent.rs:11-41,
// random_forest.rs:14-20, :127-157).  Every field is required, as with the reference's
// serde derives (no #[serde(default)]).
#include "host.hpp"

namespace frb {

namespace {

const json::Value &field(const json::Value &obj, const char *key) {
    if (obj.kind != json::Value::Object) throw Error("invalid type: expected a map of parameters");
    const json::Value *v = obj.find(key);
    if (!v) throw Error(std::string("missing field `") + key + "`");
    return *v;
}

uint64_t get_u64(const json::Value &obj, const char *key) {
    const json::Value &v = field(obj, key);
    if (v.kind == json::Value::UInt) return v.u;
    throw Error(std::string("invalid type for `") + key + "`: expected an unsigned integer");
}

uint32_t get_u32(const json::Value &obj, const char *key) {
    const uint64_t v = get_u64(obj, key);
    if (v > 0xFFFFFFFFull) throw Error(std::string("invalid value for `") + key + "`: expected u32");
    return (uint32_t)v;
}

double get_f64(const json::Value &obj, const char *key) {
    const json::Value &v = field(obj, key);
    if (!v.is_number()) throw Error(std::string("invalid type for `") + key + "`: expected f64");
    return v.as_double();
}

bool get_bool(const json::Value &obj, const char *key) {
    const json::Value &v = field(obj, key);
    if (v.kind != json::Value::Bool) throw Error(std::string("invalid type for `") + key + "`: expected a boolean");
    return v.b;
}

uint64_t default_seed() {  // Rand64::new(0xdeadbeef).rand_u64(), coordinate_ascent.rs:27,34
    Rand64 r((unsigned __int128)0xdeadbeefull);
    return r.rand_u64();
}

}  // namespace

CoordinateAscentParams CoordinateAscentParams::defaults() {
    CoordinateAscentParams p;
    p.seed = default_seed();
    return p;
}

CoordinateAscentParams CoordinateAscentParams::from_json(const json::Value &v) {
    CoordinateAscentParams p;
    p.num_restarts = get_u32(v, "num_restarts");
    p.num_max_iterations = get_u32(v, "num_max_iterations");
    p.step_base = get_f64(v, "step_base");
    p.step_scale = get_f64(v, "step_scale");
    p.tolerance = get_f64(v, "tolerance");
    p.seed = get_u64(v, "seed");
    p.normalize = get_bool(v, "normalize");
    p.quiet = get_bool(v, "quiet");
    p.init_random = get_bool(v, "init_random");
    p.output_ensemble = get_bool(v, "output_ensemble");
    if (p.tolerance != p.tolerance) throw Error("tolerance must not be NaN");
    return p;
}

json::Value CoordinateAscentParams::to_json() const {
    json::Value o = json::Value::object();
    o.set("num_restarts", json::Value::uinteger(num_restarts));
    o.set("num_max_iterations", json::Value::uinteger(num_max_iterations));
    o.set("step_base", json::Value::number(step_base));
    o.set("step_scale", json::Value::number(step_scale));
    o.set("tolerance", json::Value::number(tolerance));
    o.set("seed", json::Value::uinteger(seed));
    o.set("normalize", json::Value::boolean(normalize));
    o.set("quiet", json::Value::boolean(quiet));
    o.set("init_random", json::Value::boolean(init_random));
    o.set("output_ensemble", json::Value::boolean(output_ensemble));
    return o;
}

RandomForestParams RandomForestParams::defaults() {
    RandomForestParams p;
    p.seed = default_seed();
    return p;
}

RandomForestParams RandomForestParams::from_json(const json::Value &v) {
    RandomForestParams p;
    p.seed = get_u64(v, "seed");
    p.quiet = get_bool(v, "quiet");
    p.num_trees = get_u32(v, "num_trees");
    p.weight_trees = get_bool(v, "weight_trees");
    const json::Value &sm = field(v, "split_method");
    // serde writes the unit-tuple variants as {"SquaredError":[]}; python callers send that
    // back, or a bare string -- both are accepted
    if (sm.kind == json::Value::String) {
        p.split_method = sm.s;
    } else if (sm.kind == json::Value::Object && sm.obj.size() == 1) {
        p.split_method = sm.obj[0].first;
    } else {
        throw Error("invalid type for `split_method`");
    }
    if (p.split_method != "SquaredError" && p.split_method != "BinaryGiniImpurity" &&
        p.split_method != "InformationGain" && p.split_method != "TrueVarianceReduction")
        throw Error("unknown variant `" + p.split_method +
                    "`, expected one of `SquaredError`, `BinaryGiniImpurity`, `InformationGain`, "
                    "`TrueVarianceReduction`");
    p.instance_sampling_rate = get_f64(v, "instance_sampling_rate");
    p.feature_sampling_rate = get_f64(v, "feature_sampling_rate");
    p.min_leaf_support = get_u32(v, "min_leaf_support");
    p.split_candidates = get_u32(v, "split_candidates");
    p.max_depth = get_u32(v, "max_depth");
    return p;
}

json::Value RandomForestParams::to_json() const {
    json::Value o = json::Value::object();
    o.set("seed", json::Value::uinteger(seed));
    o.set("quiet", json::Value::boolean(quiet));
    o.set("num_trees", json::Value::uinteger(num_trees));
    o.set("weight_trees", json::Value::boolean(weight_trees));
    json::Value sm = json::Value::object();
    sm.set(split_method, json::Value::array());
    o.set("split_method", std::move(sm));
    o.set("instance_sampling_rate", json::Value::number(instance_sampling_rate));
    o.set("feature_sampling_rate", json::Value::number(feature_sampling_rate));
    o.set("min_leaf_support", json::Value::uinteger(min_leaf_support));
    o.set("split_candidates", json::Value::uinteger(split_candidates));
    o.set("max_depth", json::Value::uinteger(max_depth));
    return o;
}

}  // namespace frb
