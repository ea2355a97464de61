// trees.cu -- tree-ensemble scoring with the features of a tile staged in shared memory
// (hot loop (c): model.rs:64-84 TreeNode::score, :104-112 WeightedEnsemble::score).
//
// The reference walks boxed nodes recursively, one document and one tree at a time.  Here a
// forest (a single DecisionTree, or an Ensemble whose members are all DecisionTrees -- what
// random_forest.rs produces) is flattened into one array of 16-byte nodes, and one CTA scores a
// tile of 128 consecutive positions:
//   1. the tile's slice of X -- rows 0..max used feature, 128 floats each -- is copied into shared
//      memory with 16-byte cp.async (coalesced 512 B rows; X is read from HBM exactly once);
//   2. thread = document walks the trees four at a time (four independent pointer chases per
//      thread for latency hiding), for a fixed number of levels = the deepest tree; a walker
//      that reached a leaf stays there.  Feature reads are xs[fid][t]: bank = t, so they are
//      conflict-free whatever features the lanes of a warp look at; node reads go through L1
//      (all lanes are inside the same four trees);
//   3. leaves are combined as the reference does: output += weight * leaf, in member order, f64,
//      separate multiply and add (model.rs:106-110) -- scores are bit-identical to the oracle.
// Splits were rounded DOWN to f32 on the host, so `f32 x <= f64 split` is an exact f32 compare.
// Any other model shape goes through the generic interpreter (device.cu model_score_kernel).
#include <cuda_pipeline.h>

#include <limits>

#include "device_common.cuh"
#include "model_program.hpp"
#include "tma.cuh"

namespace {

constexpr int kTile = 128;
constexpr int kWalkers = 8;
constexpr int kHeapLevels = 10;  // implicit-heap layout up to 1024 leaves per tree

__global__ void __launch_bounds__(kTile) forest_tile_kernel(const float *__restrict__ x, size_t ld,
                                                            uint32_t dstage, size_t n,
                                                            const uint4 *__restrict__ nodes,
                                                            const uint32_t *__restrict__ roots,
                                                            const double *__restrict__ weights,
                                                            uint32_t n_trees, uint32_t levels, int weighted,
                                                            const uint32_t *__restrict__ inst_of_pos,
                                                            double *__restrict__ out_pos,
                                                            double *__restrict__ out_inst) {
    extern __shared__ __align__(128) float xs[];  // [dstage][kTile]
    const int t = threadIdx.x;
    const size_t p0 = (size_t)blockIdx.x * kTile;
    // stage: thread t copies 4 consecutive positions of feature row f = it * 4 + t / 32
    {
        const int col = (t & 31) * 4, r0 = t >> 5;
        for (uint32_t f = r0; f < dstage; f += kTile / 32)
            __pipeline_memcpy_async(xs + (size_t)f * kTile + col, x + (size_t)f * ld + p0 + col, 16);
        __pipeline_commit();
        __pipeline_wait_prior(0);
    }
    __syncthreads();
    const size_t p = p0 + t;
    if (p >= n) return;
    const float *__restrict__ mine = xs + t;
    double acc = 0.0;
    for (uint32_t t0 = 0; t0 < n_trees; t0 += kWalkers) {
        uint32_t node[kWalkers];
#pragma unroll
        for (int i = 0; i < kWalkers; ++i) node[i] = __ldg(roots + min(t0 + i, n_trees - 1));
        for (uint32_t lvl = 0; lvl < levels; ++lvl) {
            uint4 w[kWalkers];
#pragma unroll
            for (int i = 0; i < kWalkers; ++i) w[i] = __ldg(nodes + node[i]);
#pragma unroll
            for (int i = 0; i < kWalkers; ++i) {
                if (w[i].x != frb::FR_LEAF) {
                    // model.rs:75: a feature the row does not carry reads as 0.0
                    const float v = w[i].x < dstage ? mine[(size_t)w[i].x * kTile] : 0.0f;
                    node[i] = (v <= __uint_as_float(w[i].y)) ? w[i].z : w[i].w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kWalkers; ++i) {
            if (t0 + i < n_trees) {
                const uint4 w = __ldg(nodes + node[i]);
                const double leaf = __hiloint2double((int)w.w, (int)w.z);
                if (weighted)
                    acc = __dadd_rn(acc, __dmul_rn(__ldg(weights + t0 + i), leaf));
                else
                    acc = leaf;
            }
        }
    }
    if (out_pos) out_pos[p] = acc;
    if (out_inst) out_inst[inst_of_pos[p]] = acc;
}

// Same walk over trees stored as implicit heaps (levels <= kHeapLevels): node i has children
// 2i+1 / 2i+2, 8 bytes per node {fid, split}; a tree is one contiguous block
//   [2^levels - 1 nodes][1 pad][2^levels f64 leaves]  =  2^(levels+4) bytes.
// Shallow leaves are padded down to the last level (always-left dummy splits over copies of the
// leaf value), so every walk takes exactly `levels` steps with no leaf test.  The X tile (one
// 512-byte bulk copy per feature row) and the trees (double-buffered 16 KB batches) are moved by
// the TMA copy engine (cp.async.bulk completing on mbarriers; one elected thread issues them), so
// a walk never waits on L2: every step is two shared-memory reads (node, feature) and a compare.
__global__ void __launch_bounds__(kTile) forest_heap_kernel(const float *__restrict__ x, size_t ld,
                                                            uint32_t dstage, size_t n,
                                                            const uint4 *__restrict__ heap,
                                                            const double *__restrict__ weights,
                                                            uint32_t n_trees, uint32_t levels,
                                                            uint32_t batch, int weighted,
                                                            const uint32_t *__restrict__ inst_of_pos,
                                                            double *__restrict__ out_pos,
                                                            double *__restrict__ out_inst) {
    extern __shared__ __align__(128) float xs[];  // [dstage][kTile], then 2 tree-batch buffers
    __shared__ __align__(8) uint64_t bars[3];     // X tile, tree batches (double-buffered)
    const int t = threadIdx.x;
    const size_t p0 = (size_t)blockIdx.x * kTile;
    const uint32_t tree_u4 = 1u << levels;      // 16-byte units per tree
    const uint32_t batch_u4 = batch * tree_u4;  // ... per batch buffer
    uint4 *tbuf = (uint4 *)(xs + (size_t)dstage * kTile);
    const uint32_t n_batches = (n_trees + batch - 1) / batch;
    // one elected thread drives the copy engine: a 512-byte row of X per feature, 16 KB per batch
    auto stage_batch = [&](uint32_t b) {
        const uint32_t first = b * batch;
        const uint32_t bytes = min(batch, n_trees - first) * tree_u4 * 16u;
        mbar_expect_tx(&bars[1 + (b & 1)], bytes);
        bulk_copy_g2s(tbuf + (size_t)(b & 1) * batch_u4, heap + (size_t)first * tree_u4, bytes, &bars[1 + (b & 1)]);
    };
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[0], dstage * kTile * (unsigned)sizeof(float));
        for (uint32_t f = 0; f < dstage; ++f)
            bulk_copy_g2s(xs + (size_t)f * kTile, x + (size_t)f * ld + p0, kTile * (unsigned)sizeof(float), &bars[0]);
        stage_batch(0);
    }
    __syncthreads();  // barrier objects are initialised for everyone
    mbar_wait(&bars[0], 0);
    const size_t p = p0 + t;
    const float *__restrict__ mine = xs + t;
    double acc = 0.0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        if (t == 0 && b + 1 < n_batches) {
            // the other buffer was last read (generic proxy) before the barrier that closed
            // batch b - 1; order those reads before the async-proxy refill
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            stage_batch(b + 1);
        }
        mbar_wait(&bars[1 + (b & 1)], (b >> 1) & 1u);
        const uint4 *cur = tbuf + (size_t)(b & 1) * batch_u4;
        const uint32_t first = b * batch, in_batch = min(batch, n_trees - first);
        for (uint32_t t0 = 0; t0 < in_batch; t0 += kWalkers) {
            uint32_t node[kWalkers];
            const uint2 *base[kWalkers];
#pragma unroll
            for (int i = 0; i < kWalkers; ++i) {
                node[i] = 0;
                base[i] = (const uint2 *)(cur + (size_t)min(t0 + i, in_batch - 1) * tree_u4);
            }
            for (uint32_t lvl = 0; lvl < levels; ++lvl) {
                uint2 w[kWalkers];
#pragma unroll
                for (int i = 0; i < kWalkers; ++i) w[i] = base[i][node[i]];
#pragma unroll
                for (int i = 0; i < kWalkers; ++i) {
                    const float v = w[i].x < dstage ? mine[(size_t)w[i].x * kTile] : 0.0f;  // model.rs:75
                    node[i] = 2 * node[i] + ((v <= __uint_as_float(w[i].y)) ? 1u : 2u);
                }
            }
#pragma unroll
            for (int i = 0; i < kWalkers; ++i) {
                if (t0 + i < in_batch) {
                    // leaves start one 8-byte slot after the last node
                    const double leaf = ((const double *)base[i])[node[i] + 1];
                    if (weighted)
                        acc = __dadd_rn(acc, __dmul_rn(__ldg(weights + first + t0 + i), leaf));
                    else
                        acc = leaf;
                }
            }
        }
        __syncthreads();  // the buffer is refilled two batches from now
    }
    if (p < n) {
        if (out_pos) out_pos[p] = acc;
        if (out_inst) out_inst[inst_of_pos[p]] = acc;
    }
}

}  // namespace

namespace frbdev {

// Recognises  TREE END  and  ENS_BEGIN (TREE ENS_ACC w)* END  and builds the flat forest.
int build_forest(fr_dev_model *m, const uint64_t *code, size_t n_words) {
    Forest &fo = m->forest;
    fo.ok = false;
    std::vector<uint4> nodes;
    std::vector<uint32_t> roots;
    std::vector<double> weights;
    size_t pc = 0;
    bool weighted = false;
    auto op_of = [&](size_t at) { return (uint32_t)(code[at] & 0xff); };
    if (pc < n_words && op_of(pc) == frb::OP_ENS_BEGIN) {
        weighted = true;
        ++pc;
    }
    uint32_t levels = 0, max_fid = 0;
    while (pc < n_words && op_of(pc) == frb::OP_TREE) {
        const size_t nn = (size_t)(code[pc] >> 8);
        if (pc + 1 + 2 * nn > n_words || nn == 0) return 0;
        const uint32_t base = (uint32_t)nodes.size();
        if ((uint64_t)base + nn > 0xFFFFFFF0ull) return 0;
        std::vector<uint32_t> depth(nn, 0);
        for (size_t k = 0; k < nn; ++k) {  // children follow their parent (pre-order lowering)
            const uint64_t w0 = code[pc + 1 + 2 * k], w1 = code[pc + 2 + 2 * k];
            uint4 nd;
            nd.x = (uint32_t)w0;
            nd.y = (uint32_t)(w0 >> 32);
            if (nd.x == frb::FR_LEAF) {
                nd.z = (uint32_t)w1;          // f64 bits, low word
                nd.w = (uint32_t)(w1 >> 32);  // high word
                levels = std::max(levels, depth[k]);
            } else {
                const uint32_t l = (uint32_t)w1, r = (uint32_t)(w1 >> 32);
                if (l >= nn || r >= nn || l <= k || r <= k) return 0;
                depth[l] = depth[r] = depth[k] + 1;
                nd.z = base + l;
                nd.w = base + r;
                max_fid = std::max(max_fid, nd.x);
            }
            nodes.push_back(nd);
        }
        roots.push_back(base);
        pc += 1 + 2 * nn;
        if (weighted) {
            if (pc + 1 >= n_words || op_of(pc) != frb::OP_ENS_ACC) return 0;
            double w;
            memcpy(&w, &code[pc + 1], 8);
            weights.push_back(w);
            pc += 2;
        } else {
            break;  // a bare tree
        }
    }
    if (pc >= n_words || op_of(pc) != frb::OP_END || roots.empty()) return 0;
    fr_dev_dataset *ds = m->ds;
    const uint32_t dstage = (uint32_t)std::min<size_t>(ds->d, (size_t)max_fid + 1);
    if ((size_t)dstage * kTile * sizeof(float) > 200 * 1024) return 0;  // does not fit: interpreter
    if (weights.empty()) weights.push_back(1.0);
    fo.heap = levels <= (uint32_t)kHeapLevels && !getenv("FASTRANK_NO_HEAP_FOREST") &&
              (size_t)dstage * kTile * sizeof(float) + 2 * 16 * 1024 <= 220 * 1024;
    if (fo.heap) {
        // per tree: [2^levels - 1 nodes][pad][2^levels leaves], 8 bytes each
        const uint32_t n_internal = (1u << levels) - 1u, slots = 2u << levels;
        std::vector<uint64_t> hp((size_t)roots.size() * slots, 0);
        const float inf = std::numeric_limits<float>::infinity();
        uint32_t inf_bits;
        memcpy(&inf_bits, &inf, 4);
        struct Item {
            uint32_t src, heap, depth;
        };
        std::vector<Item> stack;
        for (size_t tr = 0; tr < roots.size(); ++tr) {
            uint64_t *tp = hp.data() + tr * slots;
            stack.push_back({roots[tr], 0u, 0u});
            while (!stack.empty()) {
                const Item it = stack.back();
                stack.pop_back();
                const uint4 nd = nodes[it.src];
                if (it.depth == levels) {  // a leaf slot of the heap (nd is a leaf here)
                    tp[it.heap + 1] = (uint64_t)nd.z | ((uint64_t)nd.w << 32);
                    continue;
                }
                if (nd.x == frb::FR_LEAF) {  // shallow leaf: pad with an always-left split
                    tp[it.heap] = (uint64_t)0u | ((uint64_t)inf_bits << 32);
                    stack.push_back({it.src, 2 * it.heap + 1, it.depth + 1});
                    stack.push_back({it.src, 2 * it.heap + 2, it.depth + 1});
                } else {
                    tp[it.heap] = (uint64_t)nd.x | ((uint64_t)nd.y << 32);
                    stack.push_back({nd.z, 2 * it.heap + 1, it.depth + 1});
                    stack.push_back({nd.w, 2 * it.heap + 2, it.depth + 1});
                }
            }
        }
        (void)n_internal;
        std::vector<uint4> packed(hp.size() / 2);
        memcpy(packed.data(), hp.data(), hp.size() * 8);
        CU(fo.heap_blocks.upload(packed));
        const uint32_t tree_bytes = 16u << levels;
        fo.batch = std::max<uint32_t>(1, (16u * 1024u) / tree_bytes);
    }
    CU(fo.nodes.upload(nodes));
    CU(fo.roots.upload(roots));
    CU(fo.weights.upload(weights));
    CU(cudaStreamSynchronize(0));
    fo.n_trees = (uint32_t)roots.size();
    fo.levels = levels;
    fo.dstage = std::max<uint32_t>(dstage, 1);
    fo.weighted = weighted;
    fo.ok = true;
    return 0;
}

int launch_forest(fr_dev_dataset *ds, const fr_dev_model *m, double *out_pos, double *out_inst,
                  cudaStream_t stream) {
    const Forest &fo = m->forest;
    const size_t smem = (size_t)fo.dstage * kTile * sizeof(float);
    if (smem > 48 * 1024)
        CU(cudaFuncSetAttribute(forest_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((ds->n + kTile - 1) / kTile);
    if (fo.heap) {
        const size_t hsmem = smem + 2 * (size_t)fo.batch * (16u << fo.levels);
        if (hsmem > 48 * 1024)
            CU(cudaFuncSetAttribute(forest_heap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
        auto *ev = ds->prof_slot();
        if (ev) cudaEventRecord(ev->first, stream);
        forest_heap_kernel<<<grid, kTile, hsmem, stream>>>(ds->x.p, ds->ld, fo.dstage, ds->n, fo.heap_blocks.p,
                                                           fo.weights.p, fo.n_trees, fo.levels, fo.batch,
                                                           fo.weighted ? 1 : 0, ds->inst_of_pos_dev.p, out_pos,
                                                           out_inst);
        if (ev) cudaEventRecord(ev->second, stream);
        LAUNCHED();
        CU(cudaGetLastError());
        return 0;
    }
    auto *ev = ds->prof_slot();
    if (ev) cudaEventRecord(ev->first, stream);
    forest_tile_kernel<<<grid, kTile, smem, stream>>>(ds->x.p, ds->ld, fo.dstage, ds->n, fo.nodes.p,
                                                      fo.roots.p, fo.weights.p, fo.n_trees, fo.levels,
                                                      fo.weighted ? 1 : 0, ds->inst_of_pos_dev.p, out_pos,
                                                      out_inst);
    if (ev) cudaEventRecord(ev->second, stream);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

}  // namespace frbdev
