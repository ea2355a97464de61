// model_program.hpp -- flattened ModelEnum shared by the host lowering (model.cpp) and the
// device interpreter (device.cu).
//
// The reference scores a model by recursive dynamic dispatch over boxed nodes
// (model.rs:18-27, :64-84, :104-112).  Here a model is lowered once into a postfix program of
// 64-bit words that one GPU thread interprets per document with a tiny register stack:
//
//   OP_END
//   OP_LINEAR   arg = n          followed by n f64 words            push dot(x[0..min(n,D)), w)
//   OP_SINGLE   arg = fid        followed by 1 f64 word (dir)       push dir * x[fid]
//   OP_TREE     arg = n_nodes    followed by 2*n_nodes words        push leaf value
//        node word0: low 32 = fid (FR_LEAF for a leaf), high 32 = float bits of the split
//                    rounded DOWN to f32 (x_f32 <= split_f64  <=>  x_f32 <= rounddown_f32(split))
//        node word1: leaf -> f64 value bits; split -> low 32 = lhs node, high 32 = rhs node
//   OP_ENS_BEGIN                                                    push 0.0
//   OP_ENS_ACC                   followed by 1 f64 word (weight)    v = pop; top = top + w * v
//
// An Ensemble{weights, models} lowers to ENS_BEGIN, (member, ENS_ACC w)*, which reproduces
// the reference's left-to-right `output += weight * member` (model.rs:106-110).
#pragma once
#include <cstdint>

namespace frb {

enum ModelOp : uint32_t {
    OP_END = 0,
    OP_LINEAR = 1,
    OP_SINGLE = 2,
    OP_TREE = 3,
    OP_ENS_BEGIN = 4,
    OP_ENS_ACC = 5,
};

constexpr uint32_t FR_LEAF = 0xFFFFFFFFu;
constexpr int FR_MODEL_STACK = 8;

inline uint64_t op_word(ModelOp op, uint64_t arg) { return (uint64_t)op | (arg << 8); }

}  // namespace frb
