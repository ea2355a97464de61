// device_common.cuh -- state shared by the CUDA translation units (device.cu, sweep_fast.cu):
// error plumbing, device buffers, the run-time NCCL binding and the handle structs behind
// the fr_dev_* ABI.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fastrank_b200.h"
#include "model_program.hpp"

namespace frbdev {

inline thread_local std::string g_last_error;
inline std::atomic<uint64_t> g_kernel_launches{0};

inline int fail(const std::string &msg) {
    g_last_error = msg;
    return 1;
}

#define CU(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(std::string(#expr) + " failed: " + cudaGetErrorString(e__));      \
    } while (0)

#define LAUNCHED() g_kernel_launches.fetch_add(1, std::memory_order_relaxed)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool owns = true;  // false: a slice of an UploadBatch arena
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p && owns) cudaFree(p);
        p = nullptr;
        n = 0;
        owns = true;
    }
    void adopt(T *ptr, size_t count) {
        release();
        p = ptr;
        n = count;
        owns = false;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        return cudaMalloc((void **)&p, sizeof(T) * (count ? count : 1));
    }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t s = 0) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s);
    }
    cudaError_t ensure(size_t count) { return count <= n && p ? cudaSuccess : alloc(count); }
};

// Several host arrays -> ONE device allocation and one copy per array (a plan uploads a dozen
// small arrays; a cudaMalloc apiece costs more than the copies).
struct UploadBatch {
    struct Item {
        void **slot;
        size_t *count_slot;
        bool *owns_slot;
        const void *src;
        size_t bytes, count, offset;
    };
    std::vector<Item> items;
    size_t total = 0;
    template <typename T>
    void add(DevBuf<T> &dst, const std::vector<T> &h) {
        dst.release();
        Item it;
        it.slot = (void **)&dst.p;
        it.count_slot = &dst.n;
        it.owns_slot = &dst.owns;
        it.src = h.data();
        it.bytes = sizeof(T) * h.size();
        it.count = h.size();
        it.offset = total;
        total += (it.bytes + 255) & ~(size_t)255;
        items.push_back(it);
    }
    cudaError_t commit(DevBuf<unsigned char> &arena, cudaStream_t s) {
        const bool trace = getenv("FASTRANK_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = arena.alloc(total ? total : 1);
        if (e != cudaSuccess) return e;
        if (trace)
            fprintf(stderr, "[fastrank_b200]   arena cudaMalloc(%zu) %.2f ms\n", total,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        for (const Item &it : items) {
            *it.slot = arena.p + it.offset;
            *it.count_slot = it.count;
            *it.owns_slot = false;
            if (it.bytes) {
                e = cudaMemcpyAsync(arena.p + it.offset, it.src, it.bytes, cudaMemcpyHostToDevice, s);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    }
};

template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = count;
        return cudaMallocHost((void **)&p, sizeof(T) * (count ? count : 1));
    }
};

constexpr int kMaxTile = 1024;  // largest tile of the exact-order kernels
constexpr int kFastTile = 512;  // largest tile of the batched sweep (sweep_fast.cu)
constexpr int kMaxSweeps = 8;   // sweeps sharing one pass over X in the batched sweep (one blockIdx.y group)
#ifdef __CUDACC__
// cnt += (sj > my) || (sj == my && before): the document at j outranks mine
// (evaluators.rs:33-49: score descending; on equal scores the earlier local position wins,
// local order being the reference's gain-ascending / id-ascending tie-break).  DSETP compares
// -0.0 == +0.0 like NotNan does (evaluators.rs:36); NaN scores are reported before ranking.
__device__ __forceinline__ void count_outranks(unsigned &cnt, double sj, double my, unsigned before) {
    asm("{ .reg .pred p, q; setp.ne.u32 q, %3, 0; setp.eq.and.f64 p, %1, %2, q;"
        " setp.gt.or.f64 p, %1, %2, p; @p add.u32 %0, %0, 1; }"
        : "+r"(cnt)
        : "d"(sj), "d"(my), "r"(before));
}
#endif

constexpr double kFxScale = 1099511627776.0; /* 2^FR_FX_BITS */
static_assert(FR_FX_BITS == 40, "kFxScale must match FR_FX_BITS");

// ---------------------------------------------------------------------------------------
// NCCL, bound at run time so the library loads on machines without it.
// ---------------------------------------------------------------------------------------
struct NcclApi {
    typedef struct {
        char internal[128];
    } UniqueId;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

inline NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        void *h = nullptr;
        for (const char *nm : names) {
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            const char *env = getenv("FASTRANK_NCCL_LIB");
            if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!h) {
            api.why = "libnccl.so.2 not found (import torch first or set FASTRANK_NCCL_LIB)";
            return;
        }
        api.GetUniqueId = (int (*)(NcclApi::UniqueId *))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank =
            (int (*)(void **, int, NcclApi::UniqueId, int))dlsym(h, "ncclCommInitRank");
        api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *,
                                 cudaStream_t))dlsym(h, "ncclAllReduce");
        api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
        api.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
        if (!api.ok) api.why = "libnccl is missing expected symbols";
    });
    return api;
}
constexpr int kNcclUint8 = 1;
constexpr int kNcclInt64 = 4;
constexpr int kNcclUint64 = 5;
constexpr int kNcclSum = 0;

constexpr int ERR_NAN_SCORE = 1;
constexpr int ERR_DCG_ABOVE_IDEAL = 2;
constexpr int ERR_PEER_TIMEOUT = 4;

}  // namespace frbdev
using namespace frbdev;

// Peer-memory mailboxes for the sweep kernel's fused reduction (sweep_fast.cu): every rank maps
// every other rank's mailbox through CUDA IPC (NVLink / NVSwitch peer stores).
//   mailbox = 2 epochs x world slots x kMailWords int64, then 2 x world u32 arrival flags
constexpr uint32_t kMailWords = 4096;
struct Mailbox {
    bool ok = false;
    DevBuf<unsigned char> mem;            // this rank's mailbox
    std::vector<unsigned char *> peer;    // peer[r]: rank r's mailbox as seen from this device
    std::vector<void *> opened;           // IPC mappings to close
    DevBuf<unsigned char *> peer_dev;     // the same pointers, for the kernel
    uint32_t epoch = 0;                   // one per fused reduction, identical on every rank
    __host__ __device__ static size_t slot_bytes() { return sizeof(long long) * kMailWords; }
    __host__ __device__ static size_t flags_offset(int world) { return 2 * (size_t)world * slot_bytes(); }
    static size_t bytes(int world) { return flags_offset(world) + 2 * (size_t)world * sizeof(uint32_t); }
};

struct fr_dev_comm {
    uint64_t generation = 0;  // unique per communicator ever created (a freed address can be reused)
    int device = 0;
    int rank = 0;
    int world = 1;
    void *comm = nullptr;
    cudaStream_t stream = nullptr;
    DevBuf<uint64_t> scratch;
    Mailbox mail;
};

struct fr_dev_dataset {
    int device = 0;
    size_t n = 0, d = 0, ld = 0;
    uint32_t nq = 0;
    DevBuf<float> x;        // [d][ld]
    DevBuf<float> gain;     // [n] by position
    DevBuf<double> gexp;    // [n] by position
    DevBuf<uint32_t> inst_of_pos_dev;
    DevBuf<uint32_t> len_pos;    // optional, by position: features below this id are present in the row (libsvm data)
    DevBuf<uint32_t> present_pos;  // optional, by position: bitmap of the feature ids the row carries (Sparse32 rows)
    uint32_t present_words = 0;    // 32-bit words per row of present_pos
    std::vector<uint32_t> inst_of_pos, pos_of_inst;
    std::vector<uint32_t> q_start, q_len;  // per query, in positions
    std::vector<float> gain_pos;           // host copy, by position
    cudaStream_t stream = nullptr;
    DevBuf<double> scores_pos;   // scratch for model scoring
    DevBuf<double> scores_inst;  // scratch for predict
    // device-side timing (fr_dev_timer_*, fr_dev_profile_*)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    ~fr_dev_dataset() {
        for (auto &pr : prof_events) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        if (stream) cudaStreamDestroy(stream);
    }
    // event pair bracketing one kernel launch while profiling is on
    std::pair<cudaEvent_t, cudaEvent_t> *prof_slot() {
        if (!profile) return nullptr;
        if (prof_used == prof_events.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return nullptr;
            prof_events.emplace_back(a, b);
        }
        return &prof_events[prof_used++];
    }
};

// A model that is a forest of regression trees, flattened for trees.cu.
struct Forest {
    bool ok = false;
    DevBuf<uint4> nodes;       // .x fid (FR_LEAF: leaf)  .y split (f32 bits)  .z/.w children, or leaf f64 lo/hi
    DevBuf<uint32_t> roots;    // first node of every tree
    DevBuf<double> weights;    // ensemble weights
    DevBuf<uint4> heap_blocks; // implicit-heap layout, per tree [2^levels - 1 nodes][pad][2^levels leaves]
    uint32_t n_trees = 0, levels = 0, dstage = 1, batch = 1;
    bool weighted = false, heap = false;
};

struct fr_dev_model {
    fr_dev_dataset *ds = nullptr;
    DevBuf<uint64_t> code;
    size_t n_words = 0;
    Forest forest;
};

// Device-side view of a plan (passed to kernels by value).
struct PlanView {
    const float *x;
    size_t ld;
    uint32_t dfeat;
    const float *gain;
    const double *gexp;
    const double *lg2;           // log2(i + 2)
    const uint32_t *tile_doc_off;  // nt + 1
    const uint32_t *tile_q_off;    // nt + 1
    const uint32_t *pd_pos;        // plan doc -> position
    const uint32_t *pd_q;          // plan doc -> (local query start) | (local query end << 16)
    const uint32_t *pq_local;      // plan query -> (local start) | (len << 16)
    const uint32_t *pq_doc0;       // plan query -> first plan doc
    const double *pq_norm;         // ideal DCG (NaN = none) or num_relevant
    const uint32_t *pq_view;       // plan query -> view (output) index
    uint32_t nt;
    uint32_t nq_plan;
    uint32_t nq_view;
    int metric;
    int depth;  // INT_MAX when absent
};

// Extra plan state of the batched sweep kernel (sweep_fast.cu).
struct FastPlan {
    bool ok = false;          // false: the plan's geometry is outside what the kernel handles
    std::string why;          // ... and why
    int td = 4;               // documents ranked per warp task
    uint32_t n_tasks = 0;
    uint32_t n_cls = 0;       // distinct gain values among the plan's documents
    uint32_t tbl_r = 0;       // discount table rows: min(depth, longest query)
    DevBuf<uint32_t> tile_task_off;  // nt + 1
    DevBuf<uint2> tasks;             // .x = qs | qe << 16, .y = t0 | n << 16 (tile-local documents)
    DevBuf<uint8_t> pd_cls;          // plan doc -> gain class
    DevBuf<double> disc_tbl;         // [n_cls][tbl_r]  (2^gain - 1) / log2(r + 2)
    // NDCG@k with k <= 16 and <= 15 gain classes: sweep_packed_kernel (sweep_packed.cuh)
    bool packed_ok = false;
    bool slots_ok = false;  // the same kernel with ranks filed in shared-memory slots: any measure, tiles <= 256
    DevBuf<uint32_t> pk_q_task_off;     // nq_plan + 1
    DevBuf<uint32_t> pk_tile_task_off;  // nt + 1
    DevBuf<uint4> pk_tasks;             // chunks of <= 16 documents, grouped by query (PackedView::tasks)
    DevBuf<uint16_t> pk_q_order;        // tile-local query indices, costliest first
    DevBuf<double> pk_tbl;              // [n_cls + 1][tbl_r], row 0 zeros
    DevBuf<double> ap_tbl;              // AP: [relevant so far][rank] -> precision, the fold's divisions done once
    uint32_t ap_cols = 0;
    // per-call work buffers (sweep_fast.cu documents the layout): one input blob = transposed
    // weights + the call's candidates flattened into rows, one output blob = sums, error flags,
    // tile counters; each moves with a single copy through pinned memory
    DevBuf<unsigned char> in_dev, out_dev;
    PinnedBuf<unsigned char> in_host, out_host;
    // direct publication (sweep_fast.cu, FastArgs::host_flag): device-side state that the kernel
    // itself leaves zeroed, and a host-mapped block the last CTA writes sums / errors / a flag into
    DevBuf<unsigned char> state_dev;
    unsigned char *pub_host = nullptr;  // cudaHostAlloc(Mapped)
    unsigned char *pub_dev = nullptr;   // the same block as the device sees it
    uint32_t pub_epoch = 0;
    bool direct_open = false;  // a direct call did not complete: the device state may not be zero
    std::vector<uint32_t> len_docs;  // len_docs[l] = tiled documents sitting in lists of exactly l documents
    double long_list_docs_frac(uint32_t min_len) const {
        uint64_t all = 0, lng = 0;
        for (size_t l = 0; l < len_docs.size(); ++l) {
            all += len_docs[l];
            if (l >= min_len) lng += len_docs[l];
        }
        return all ? (double)lng / (double)all : 0.0;
    }
    ~FastPlan() {
        if (pub_host) cudaFreeHost(pub_host);
    }
};
constexpr uint32_t kDirectGroups = 64;                       // tile counters in the direct state block
constexpr size_t kDirectSumsOff = 512;                       // [err i32][done u32][tile ctr u32 x 64] ... [sums]
constexpr size_t kDirectStateBytes = kDirectSumsOff + sizeof(long long) * 4096;
constexpr size_t kPubErrOff = sizeof(long long) * 4096;      // [sums i64 x 4096][err i32][flag u32]

struct FastView {
    const uint32_t *tile_task_off;
    const uint2 *tasks;
    const uint8_t *pd_cls;
    const double *disc_tbl;
    uint32_t tbl_r;
    uint32_t n_cls;
};

// Queries longer than the largest tile (long_queries.cu).
struct LongPlan {
    uint32_t n_long = 0, n_docs = 0;
    uint32_t chunk = 1;  // candidates per pass over the untiled lists
    DevBuf<uint32_t> lq_off, ld_pos, lq_view, out_idx;
    DevBuf<double> lq_norm, scores, slots, w;
};

struct fr_dev_plan {
    fr_dev_dataset *ds = nullptr;
    // one allocation each behind the plan's read-only arrays; declared first so that they are
    // destroyed last (the DevBufs below that point into them do not own their memory)
    DevBuf<unsigned char> arena, fast_arena;
    FastPlan fast;
    LongPlan lng;
    int metric = 0;
    int depth = INT_MAX;
    int tb = 128;
    uint32_t nq_view = 0, nq_plan = 0, nt = 0;
    uint32_t max_len = 0;
    bool contiguous = false;  // every tile covers consecutive positions (bulk-copy friendly)
    DevBuf<double> lg2;
    DevBuf<uint32_t> tile_doc_off, tile_q_off, pd_pos, pd_q, pq_local, pq_doc0, pq_view;
    DevBuf<double> pq_norm;
    // work buffers
    DevBuf<double> w_dev, cand_dev, perq_dev;
    DevBuf<uint32_t> fid_dev, ncand_dev;
    DevBuf<long long> sums_dev;
    DevBuf<int> err_dev;
    PinnedBuf<long long> sums_host;
    PinnedBuf<int> err_host;
    fr_dev_comm *comm = nullptr;
    uint64_t nq_global = 0;
    int sm_count = 148;
    PlanView view() const {
        PlanView v;
        v.x = ds->x.p;
        v.ld = ds->ld;
        v.dfeat = (uint32_t)ds->d;
        v.gain = ds->gain.p;
        v.gexp = ds->gexp.p;
        v.lg2 = lg2.p;
        v.tile_doc_off = tile_doc_off.p;
        v.tile_q_off = tile_q_off.p;
        v.pd_pos = pd_pos.p;
        v.pd_q = pd_q.p;
        v.pq_local = pq_local.p;
        v.pq_doc0 = pq_doc0.p;
        v.pq_norm = pq_norm.p;
        v.pq_view = pq_view.p;
        v.nt = nt;
        v.nq_plan = nq_plan;
        v.nq_view = nq_view;
        v.metric = metric;
        v.depth = depth;
        return v;
    }
};

namespace frbdev {
// defined in device.cu
int check_err_flags(int flags);
int allreduce_sums(fr_dev_plan *pl, long long *dev, size_t count, cudaStream_t stream);
// defined in long_queries.cu
int build_long_plan(fr_dev_plan *pl, const std::vector<std::vector<uint32_t>> &qpos,
                    const std::vector<uint32_t> &long_views, const fr_dev_plan_desc *desc);
int eval_long_linear(fr_dev_plan *pl, const double *w_host, size_t wlen, size_t n_vec,
                     const uint32_t *out_index, long long *sums_dev, double *perq_dev, int *err_dev,
                     cudaStream_t s);
// the batched sweep's staged tables (device pointers, see FastArgs in sweep_fast.cu): n_rows
// candidate rows, row r adds into sums[row_out[r]]
int eval_long_sweep(fr_dev_plan *pl, const double *base_wt, const uint32_t *fid, uint32_t n_sweeps,
                    const double *row_w, const uint32_t *row_meta, const uint32_t *row_out,
                    const uint32_t *grp_row_off, uint32_t n_groups, uint32_t n_rows, uint32_t dm, uint32_t dm8,
                    long long *sums_dev, double *perq_dev, int *err_dev, cudaStream_t s);
int eval_long_scores(fr_dev_plan *pl, const double *scores_pos, long long *sums_dev, double *perq_dev,
                     int *err_dev, cudaStream_t s);
// defined in trees.cu
int build_forest(fr_dev_model *m, const uint64_t *code, size_t n_words);
int launch_forest(fr_dev_dataset *ds, const fr_dev_model *m, double *out_pos, double *out_inst,
                  cudaStream_t stream);
// defined in sweep_fast.cu
int build_fast_plan(fr_dev_plan *pl, const std::vector<uint32_t> &tile_q_off,
                    const std::vector<uint32_t> &pq_local, const std::vector<uint32_t> &pq_doc0,
                    const std::vector<uint32_t> &pd_pos);
}  // namespace frbdev
