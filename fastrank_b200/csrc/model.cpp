// model.cpp -- ModelEnum on the host: JSON in/out (serde layout of model.rs:10-16, :29-33,
// :42-45, :53-62, :86-90) and lowering to the device program (model_program.hpp).
#include <cmath>
#include <cstring>

#include "host.hpp"
#include "model_program.hpp"

namespace frb {

namespace {

const json::Value &require(const json::Value &obj, const char *key, const char *owner) {
    if (obj.kind != json::Value::Object)
        throw Error(std::string("invalid type: expected struct ") + owner);
    const json::Value *v = obj.find(key);
    if (!v) throw Error(std::string("missing field `") + key + "`");
    return *v;
}

double require_f64(const json::Value &v, const char *what) {
    if (!v.is_number()) throw Error(std::string("invalid type: expected f64 for ") + what);
    const double d = v.as_double();
    if (d != d) throw Error(std::string("NaN is not allowed for ") + what);
    return d;
}

uint32_t require_u32(const json::Value &v, const char *what) {
    if (v.kind == json::Value::UInt && v.u <= 0xFFFFFFFFull) return (uint32_t)v.u;
    throw Error(std::string("invalid type: expected u32 for ") + what);
}

const json::Member &single_variant(const json::Value &v, const char *owner) {
    if (v.kind != json::Value::Object || v.obj.size() != 1)
        throw Error(std::string("invalid type: expected a single-key map for enum ") + owner);
    return v.obj[0];
}

std::unique_ptr<TreeNode> tree_from_json(const json::Value &v, int depth) {
    if (depth > 2000) throw Error("tree too deep");
    const json::Member &m = single_variant(v, "TreeNode");
    auto node = std::make_unique<TreeNode>();
    if (m.first == "LeafNode") {
        node->leaf = true;
        node->value = require_f64(m.second, "LeafNode");
    } else if (m.first == "FeatureSplit") {
        node->leaf = false;
        node->fid = require_u32(require(m.second, "fid", "FeatureSplit"), "fid");
        node->split = require_f64(require(m.second, "split", "FeatureSplit"), "split");
        node->lhs = tree_from_json(require(m.second, "lhs", "FeatureSplit"), depth + 1);
        node->rhs = tree_from_json(require(m.second, "rhs", "FeatureSplit"), depth + 1);
    } else {
        throw Error("unknown variant `" + m.first + "`, expected `FeatureSplit` or `LeafNode`");
    }
    return node;
}

json::Value tree_to_json(const TreeNode &n) {
    json::Value out = json::Value::object();
    if (n.leaf) {
        out.set("LeafNode", json::Value::number(n.value));
    } else {
        json::Value body = json::Value::object();
        body.set("fid", json::Value::uinteger(n.fid));
        body.set("split", json::Value::number(n.split));
        body.set("lhs", tree_to_json(*n.lhs));
        body.set("rhs", tree_to_json(*n.rhs));
        out.set("FeatureSplit", std::move(body));
    }
    return out;
}

uint64_t f64_bits(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
}

// x_f32 <= split_f64  <=>  x_f32 <= (largest f32 that is <= split)
float round_down_to_f32(double split) {
    float s = (float)split;
    if ((double)s > split) s = std::nextafterf(s, -INFINITY);
    return s;
}

uint32_t lower_tree(const TreeNode &n, std::vector<uint64_t> &nodes) {
    const uint32_t me = (uint32_t)(nodes.size() / 2);
    nodes.push_back(0);
    nodes.push_back(0);
    if (n.leaf) {
        nodes[2 * (size_t)me] = (uint64_t)FR_LEAF;
        nodes[2 * (size_t)me + 1] = f64_bits(n.value);
    } else {
        const float s = round_down_to_f32(n.split);
        uint32_t sbits;
        memcpy(&sbits, &s, 4);
        const uint32_t l = lower_tree(*n.lhs, nodes);
        const uint32_t r = lower_tree(*n.rhs, nodes);
        if (n.fid == FR_LEAF) throw Error("feature id too large");
        nodes[2 * (size_t)me] = (uint64_t)n.fid | ((uint64_t)sbits << 32);
        nodes[2 * (size_t)me + 1] = (uint64_t)l | ((uint64_t)r << 32);
    }
    return me;
}

void lower_into(const Model &m, std::vector<uint64_t> &code, int depth) {
    if (depth >= FR_MODEL_STACK - 1) throw Error("ensemble nesting deeper than the device stack");
    switch (m.kind) {
        case Model::Linear:
            code.push_back(op_word(OP_LINEAR, m.weights.size()));
            for (double w : m.weights) code.push_back(f64_bits(w));
            break;
        case Model::SingleFeature:
            code.push_back(op_word(OP_SINGLE, m.fid));
            code.push_back(f64_bits(m.dir));
            break;
        case Model::DecisionTree: {
            std::vector<uint64_t> nodes;
            lower_tree(*m.tree, nodes);
            code.push_back(op_word(OP_TREE, nodes.size() / 2));
            code.insert(code.end(), nodes.begin(), nodes.end());
            break;
        }
        case Model::Ensemble:
            code.push_back(op_word(OP_ENS_BEGIN, 0));
            for (size_t t = 0; t < m.members.size(); ++t) {
                lower_into(m.members[t], code, depth + 1);
                code.push_back(op_word(OP_ENS_ACC, 0));
                code.push_back(f64_bits(m.weights[t]));
            }
            break;
    }
}

}  // namespace

Model Model::from_json(const json::Value &v) {
    const json::Member &m = single_variant(v, "ModelEnum");
    Model out;
    if (m.first == "Linear") {
        out.kind = Linear;
        const json::Value &w = require(m.second, "weights", "DenseLinearRankingModel");
        if (w.kind != json::Value::Array) throw Error("invalid type: expected a sequence for weights");
        for (const json::Value &x : w.arr) {
            if (!x.is_number()) throw Error("invalid type: expected f64 in weights");
            out.weights.push_back(x.as_double());
        }
    } else if (m.first == "SingleFeature") {
        out.kind = SingleFeature;
        out.fid = require_u32(require(m.second, "fid", "SingleFeatureModel"), "fid");
        const json::Value &d = require(m.second, "dir", "SingleFeatureModel");
        if (!d.is_number()) throw Error("invalid type: expected f64 for dir");
        out.dir = d.as_double();
    } else if (m.first == "DecisionTree") {
        out.kind = DecisionTree;
        out.tree = tree_from_json(m.second, 0);
    } else if (m.first == "Ensemble") {
        out.kind = Ensemble;
        const json::Value &w = require(m.second, "weights", "WeightedEnsemble");
        const json::Value &ms = require(m.second, "models", "WeightedEnsemble");
        if (w.kind != json::Value::Array || ms.kind != json::Value::Array)
            throw Error("invalid type: expected sequences for weights and models");
        for (const json::Value &x : w.arr) out.weights.push_back(require_f64(x, "ensemble weight"));
        for (const json::Value &x : ms.arr) out.members.push_back(Model::from_json(x));
        // model.rs:106: zip() stops at the shorter list
        const size_t k = std::min(out.weights.size(), out.members.size());
        out.weights.resize(k);
        out.members.resize(k);
    } else {
        throw Error("unknown variant `" + m.first +
                    "`, expected one of `SingleFeature`, `Linear`, `DecisionTree`, `Ensemble`");
    }
    return out;
}

json::Value Model::to_json() const {
    json::Value out = json::Value::object();
    switch (kind) {
        case Linear: {
            json::Value body = json::Value::object();
            json::Value w = json::Value::array();
            for (double x : weights) w.push(json::Value::number(x));
            body.set("weights", std::move(w));
            out.set("Linear", std::move(body));
            break;
        }
        case SingleFeature: {
            json::Value body = json::Value::object();
            body.set("fid", json::Value::uinteger(fid));
            body.set("dir", json::Value::number(dir));
            out.set("SingleFeature", std::move(body));
            break;
        }
        case DecisionTree:
            out.set("DecisionTree", tree_to_json(*tree));
            break;
        case Ensemble: {
            json::Value body = json::Value::object();
            json::Value w = json::Value::array();
            for (double x : weights) w.push(json::Value::number(x));
            json::Value ms = json::Value::array();
            for (const Model &m : members) ms.push(m.to_json());
            body.set("weights", std::move(w));
            body.set("models", std::move(ms));
            out.set("Ensemble", std::move(body));
            break;
        }
    }
    return out;
}

std::vector<uint64_t> Model::lower() const {
    std::vector<uint64_t> code;
    lower_into(*this, code, 0);
    code.push_back(op_word(OP_END, 0));
    return code;
}

}  // namespace frb
