// sweep_fast.cu -- the batched coordinate-ascent sweep: every restart's line search in ONE pass
// over the feature matrix (coordinate_ascent.rs:131-177 for all restarts of a global step).
//
// Reference work being replaced, per candidate weight vector (evaluators.rs:173-224):
//   score every document (dense_dataset.rs:67-76), sort each query by (score desc, gain asc,
//   id asc) (evaluators.rs:33-49), NDCG / AP / RR per query, mean.
//
// Shape of the kernel (one CTA = one tile of whole queries, <= TB documents; persistent CTAs):
//   phase 1  stream the tile's slice of X once (feature-major, coalesced, software-prefetched,
//            kept out of L1): per document and per sweep s, with f_s = the coordinate the sweep
//            varies,   T_s = sum_{j != f_s} x_j w_sj   (f64, separate multiply and add, j
//            ascending; the host zeroes w at f_s)   and   xf_s = x_{f_s}.
//   phase 2  the call's candidates are flattened into rows sorted by sweep; 32 rows at a time
//            (lane = row, so lanes are filled across sweep boundaries):
//              score[row][t] = T_s + xf_s * cand_row        (one f64 multiply, one add)
//            then rank by counting: a warp task owns TD documents of one query that can
//            contribute to the metric (gain != 0 for NDCG, gain > 0 for AP / RR) and walks the
//            query once; documents before t count when score >=, documents after t when score >
//            -- the reference's tie-break, because tile order is (gain asc, id asc).  One DSETP
//            (FP64 pipe) + one predicated add (integer pipe) per comparison.
//            rank -> slot[rank] = gain class; one warp per query then folds the slots in rank
//            order (the reference's left-to-right f64 sums over a host-built table of
//            (2^gain - 1) / log2(rank + 2)) and accumulates round(value * 2^40).  The folds of
//            group g ride in the task queue of group g + 1.
//   tail     multi-GPU: the last CTA to finish exchanges the GPU's integer sums with every peer
//            through NVLink-mapped mailboxes and leaves the total in place (no collective launch).
// Scheduling: work items come from an atomic counter -- whole tiles first, the last gridDim.x / 4
// tiles as quarter items (a quarter of the row groups each) so that the SMs drain together.
//
// Arithmetic contract ("fast" mode): a candidate's score is the reference's dot product with
// the varied coordinate's term added last instead of in position -- a few roundings apart
// (~1e-16 relative) from dense_dataset.rs:67-76.  Everything after the score is the
// reference's: whenever the induced ranking of a query is the same -- always, except for
// documents whose scores agree to the last bits -- the per-query metric is bit-identical to the
// oracle.  When the arithmetic is exact (e.g. dyadic weights on small-integer features, or an
// all-zero base) the scores themselves are bit-identical, ties included.  The exact-order
// kernel (device.cu coord_sweep_kernel) stays available behind fr_dev_eval_coord_sweeps.
#include <sched.h>

#include "device_common.cuh"

namespace {

constexpr int kMaxRows = kMaxSweeps * 64;  // 8 sweeps x (1 + 2 x 25) candidates fit one pass
constexpr int kSmemTbl = 64;   // discount-table entries kept in shared memory
constexpr double kFx = 1099511627776.0;
static_assert(FR_FX_BITS == 40, "kFx must match FR_FX_BITS");

// The candidates of a call are flattened into ROWS, sorted by sweep; a warp ranks 32 rows at a
// time (lane = row), so lanes are filled across sweep boundaries (8 x 26 candidates = 7 groups
// of 32 instead of 8 of 26).
struct FastArgs {
    const double *base_wt;        // [n_groups][dm][kMaxSweeps] base weights, transposed per group, dm rounded up to 8
    const uint32_t *fid;          // [n_sweeps]
    const double *row_w;          // [n_rows]
    const uint32_t *row_meta;     // [n_rows] sweep index inside the row's group of kMaxSweeps
    const uint32_t *row_out;      // [n_rows] index into sums / perq
    const uint32_t *grp_row_off;  // [n_groups + 1]
    long long *sums;
    double *perq;  // nullptr or [n_out][nq_view]
    uint32_t n_sweeps, wlen;
    uint32_t dm;  // min(wlen, features): zip() truncation of dense_dataset.rs:67-76
    unsigned *tile_ctr;  // [n_groups] zeroed before the launch: tiles are handed out dynamically
    int *err;
    // fused cross-GPU reduction (nullptr: single GPU, or the NCCL all-reduce follows the kernel)
    unsigned char *const *mail_peers;  // [world] every rank's mailbox, own included
    unsigned *done_ctr;                // CTAs that have flushed their sums (zeroed before the launch)
    uint32_t mail_rank, mail_world, mail_epoch, mail_words;
    uint32_t prune_min;  // sweep_packed_kernel: lists of at least this many documents are pruned before ranking (0: never)
    long long mail_timeout_cycles;     // how long the last CTA waits for its peers (FASTRANK_PEER_TIMEOUT_S, default 30 s)
    uint32_t n_split;  // tiles handed out as quarter items (the last ones of the queue)
    uint32_t split_parts;  // sweep_packed_kernel: parts a split tile is cut into (2 or 4)
    // direct publication (nullptr: the host copies the sums back itself): the last CTA writes the
    // final sums and error flags into host-mapped pinned memory, clears the device-side state for
    // the next launch and raises host_flag = host_epoch, which the host spins on -- no memset, no
    // device-to-host copy and no stream synchronisation per step
    long long *host_sums;
    int *host_err;
    unsigned *host_flag;
    uint32_t host_epoch;
};

// system-scope flag accesses for the peer mailboxes
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// cnt += (a >= b) / (a > b) as DSETP + predicated add (what nvcc emits for the C form is a
// three-instruction add / compare / undo sequence).
__device__ __forceinline__ void count_ge(unsigned &cnt, double a, double b) {
    asm("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt) : "d"(a), "d"(b));
}
__device__ __forceinline__ void count_gt(unsigned &cnt, double a, double b) {
    asm("{ .reg .pred p; setp.gt.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt) : "d"(a), "d"(b));
}

// X is streamed once per launch: keep it out of L1 so the (re-used) weight table stays there.
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// misc words
enum { M_F = 0, M_TILE = 16, M_NEXT = 17, M_CTR = 18, M_SWROW = 20 /* 9 entries */ };

// ---- fused all-reduce over peer memory ----
// Tail of both sweep kernels.  The last CTA of this GPU to flush its sums stores the GPU's
// partial sums into every rank's mailbox (NVLink peer stores), raises its arrival flag there,
// waits for the flags of all ranks in its own mailbox and leaves the rank-ordered integer total
// in A.sums: the kernel that ranks is the kernel that reduces, no separate collective launch
// follows.
template <int TB>
__device__ __forceinline__ void fused_allreduce_tail(const FastArgs &A, int *s_misc) {
    const int t = threadIdx.x;
    __threadfence();
    __syncthreads();
    if (t == 0) s_misc[M_TILE] = atomicAdd(A.done_ctr, 1u) == gridDim.x * gridDim.y - 1 ? 1 : 0;
    __syncthreads();
    if (!s_misc[M_TILE]) return;
    __threadfence();
    const uint32_t nw = A.mail_words;
    if (A.mail_peers != nullptr) {
        const uint32_t world = A.mail_world, me = A.mail_rank, buf = A.mail_epoch & 1u;
        const size_t slot = sizeof(long long) * kMailWords;
        for (uint32_t r = 0; r < world; ++r) {
            long long *dst = (long long *)(A.mail_peers[r] + ((size_t)buf * world + me) * slot);
            for (uint32_t i = t; i < nw; i += TB) dst[i] = __ldcg(A.sums + i);
        }
        __threadfence_system();
        __syncthreads();
        if ((uint32_t)t < world) {
            unsigned *flag = (unsigned *)(A.mail_peers[t] + Mailbox::flags_offset((int)world)) + buf * world + me;
            st_release_sys(flag, A.mail_epoch);
            const unsigned *mine = (const unsigned *)(A.mail_peers[me] + Mailbox::flags_offset((int)world)) + buf * world + t;
            const long long t_begin = clock64();
            while (ld_acquire_sys(mine) != A.mail_epoch) {
                if (clock64() - t_begin > A.mail_timeout_cycles) {  // a peer never made this call
                    atomicOr(A.err, ERR_PEER_TIMEOUT);
                    break;
                }
                __nanosleep(200);
            }
        }
        __syncthreads();
        __threadfence_system();
        const unsigned char *box = A.mail_peers[me] + (size_t)buf * world * slot;
        for (uint32_t i = t; i < nw; i += TB) {
            long long total = 0;
            for (uint32_t r = 0; r < world; ++r)
                total += *(const volatile long long *)(box + (size_t)r * slot + sizeof(long long) * i);
            A.sums[i] = total;
        }
    }
    if (A.host_flag == nullptr) return;
    // publish to the host and leave the device-side state zeroed for the next launch
    __syncthreads();
    for (uint32_t i = t; i < nw; i += TB) {
        A.host_sums[i] = *(volatile long long *)(A.sums + i);
        A.sums[i] = 0;
    }
    for (uint32_t g = t; g < gridDim.y; g += TB) A.tile_ctr[g] = 0u;
    if (t == 0) {
        *A.host_err = atomicExch(A.err, 0);
        *A.done_ctr = 0u;
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) st_release_sys(A.host_flag, A.host_epoch);
}

template <int TB>
struct SlotType {
    typedef uint8_t type;  // document ids 1..128 / gain classes 1..255
};
template <>
struct SlotType<256> {
    typedef uint16_t type;
};
template <>
struct SlotType<512> {
    typedef uint16_t type;
};

struct SmemLayout {
    size_t score, sum, tbl, gexp, slot0, slot1, tasks, cls, rowsw, misc, w, total;
    __host__ __device__ SmemLayout(int tb, int slot_bytes, uint32_t w_doubles) {
        size_t o = 0;
        score = o;  o += align16(sizeof(double) * 32 * (size_t)(tb + 1));
        sum = o;    o += sizeof(unsigned long long) * kMaxRows;
        tbl = o;    // the discount table and the per-document gains are never live together
        gexp = o;   o += sizeof(double) * (tb > kSmemTbl ? tb : kSmemTbl);
        slot0 = o;  o += align16((size_t)slot_bytes * tb * 32);
        slot1 = o;  o += align16((size_t)slot_bytes * tb * 32);
        tasks = o;  o += sizeof(uint2) * tb;
        cls = o;    o += align16(tb);
        rowsw = o;  o += kMaxRows;
        misc = o;   o += 256;
        w = o;      o += sizeof(double) * w_doubles;  // the weight table, when it is staged
        total = o;
    }
};


// One warp task: rank W documents [t0, t0 + n) of the query occupying tile-local [qs, qe)
// under the 32 candidate rows of the lanes (myrow = this lane's row of the score matrix).
template <int W, typename slot_t>
__device__ __forceinline__ void rank_task(const double *__restrict__ myrow, int qs, int qe, int t0, int n,
                                          unsigned lim, bool live, bool tag_cls,
                                          const uint8_t *__restrict__ s_cls, slot_t *__restrict__ slots,
                                          int lane) {
    double st[W];
    unsigned cnt[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const int tt = t0 + i < qe ? t0 + i : qe - 1;
        st[i] = myrow[tt];
        cnt[i] = 0;
    }
    int jq = qs;
#pragma unroll 4
    for (; jq < t0; ++jq) {  // documents that win ties against the task's
        const double sj = myrow[jq];
#pragma unroll
        for (int i = 0; i < W; ++i) count_ge(cnt[i], sj, st[i]);
    }
#pragma unroll
    for (int jj = 0; jj < W; ++jj) {
        if (jj < n) {
            const double sj = myrow[t0 + jj];
#pragma unroll
            for (int i = 0; i < W; ++i) {
                if (i < jj) count_gt(cnt[i], sj, st[i]);
                if (i > jj) count_ge(cnt[i], sj, st[i]);
            }
        }
    }
    jq = t0 + n;
#pragma unroll 4
    for (; jq < qe; ++jq) {  // documents that lose ties
        const double sj = myrow[jq];
#pragma unroll
        for (int i = 0; i < W; ++i) count_gt(cnt[i], sj, st[i]);
    }
    if (live) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (i < n && cnt[i] < lim) {
                // what the fold needs: the gain class (table), else the document
                const unsigned tag = tag_cls ? (unsigned)s_cls[t0 + i] + 1u : (unsigned)(t0 + i + 1);
                slots[(size_t)(qs + cnt[i]) * 32 + lane] = (slot_t)tag;
            }
        }
    }
}

template <int TB, int TD, bool WS>
__global__ void __launch_bounds__(TB, (TB == 128 ? (WS ? 4 : 5) : (TB == 256 ? 2 : 1)))
sweep_fast_kernel(PlanView P, FastView F, FastArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename SlotType<TB>::type slot_t;
    constexpr int NS = kMaxSweeps;
    constexpr int ROW = TB + 1;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t s0 = blockIdx.y * NS;
    const int ns = (int)min((uint32_t)NS, A.n_sweeps - s0);
    const uint32_t dm = A.dm;
    const uint32_t row0 = A.grp_row_off[blockIdx.y];
    const int R = (int)(A.grp_row_off[blockIdx.y + 1] - row0);  // rows of this sweep group
    const int G = (R + 31) >> 5;
    const uint32_t dm8 = (dm + 7) & ~7u;
    const SmemLayout L(TB, (int)sizeof(slot_t), WS ? dm8 * NS : 0u);
    double *s_score = (double *)(smem_raw + L.score);  // [32][ROW]
    unsigned long long *s_sum = (unsigned long long *)(smem_raw + L.sum);
    double *s_tbl = (double *)(smem_raw + L.tbl);
    double *s_gexp = (double *)(smem_raw + L.gexp);
    slot_t *s_slot0 = (slot_t *)(smem_raw + L.slot0);
    const size_t slot_stride = (L.slot1 - L.slot0) / sizeof(slot_t);
    uint2 *s_tasks = (uint2 *)(smem_raw + L.tasks);
    uint8_t *s_cls = (uint8_t *)(smem_raw + L.cls);
    uint8_t *s_rowsw = (uint8_t *)(smem_raw + L.rowsw);
    int *s_misc = (int *)(smem_raw + L.misc);
    const double *__restrict__ wg = A.base_wt + (size_t)blockIdx.y * dm8 * NS;
    double *s_w = (double *)(smem_raw + L.w);
    const bool use_tbl = F.disc_tbl != nullptr;
    const bool tbl_in_smem = use_tbl && F.n_cls * F.tbl_r <= (uint32_t)kSmemTbl;
    const double *tbl = tbl_in_smem ? s_tbl : F.disc_tbl;

    // ---- per-launch setup ----
    for (int idx = t; idx < kMaxRows; idx += TB) {
        s_rowsw[idx] = idx < R ? (uint8_t)A.row_meta[row0 + idx] : (uint8_t)0;
        s_sum[idx] = 0ull;
    }
    if (tbl_in_smem)
        for (uint32_t idx = t; idx < F.n_cls * F.tbl_r; idx += TB) s_tbl[idx] = F.disc_tbl[idx];
    if (WS)
        for (uint32_t idx = t; idx < dm8 * NS; idx += TB) s_w[idx] = wg[idx];
    __syncthreads();
    if (t <= NS) {
        // first row of sweep t (rows are sorted by sweep): lower bound in the staged row table
        int lo = 0, hi = R;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int)s_rowsw[mid] < t) lo = mid + 1;
            else hi = mid;
        }
        s_misc[M_SWROW + t] = lo;
    }
    __syncthreads();
    if (t < NS) {
        const bool has_rows = s_misc[M_SWROW + t + 1] > s_misc[M_SWROW + t];
        s_misc[M_F + t] = (t < ns && has_rows) ? (int)A.fid[s0 + t] : -1;
    }
    if (t == 0) s_misc[M_TILE] = (int)atomicAdd(A.tile_ctr + blockIdx.y, 1u);
    __syncthreads();
    uint32_t fs[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) fs[s] = (uint32_t)s_misc[M_F + s];
    int nan_seen = 0;
    const double *myrow = s_score + (size_t)lane * ROW;


    // Work items: whole tiles first; the last n_split tiles are handed out as quarter items (a
    // quarter of the row groups each, phase 1 repeated) so that the final wave of the persistent
    // grid is made of short items and the SMs drain together.
    const uint32_t n_split = G >= 4 ? min(P.nt, A.n_split) : 0u;
    const uint32_t n_whole = P.nt - n_split;
    const uint32_t n_items = G > 0 ? n_whole + 4u * n_split : 0u;
    uint32_t next_item = n_items;
    for (uint32_t item = G > 0 ? (uint32_t)s_misc[M_TILE] : n_items; item < n_items; item = next_item) {
        // the next item is claimed now; every thread picks the value up BEFORE this item's last
        // barrier, so the claim after next cannot overtake a slow reader
        if (t == 0) s_misc[M_NEXT] = (int)atomicAdd(A.tile_ctr + blockIdx.y, 1u);
        uint32_t tile = item;
        int g_begin = 0, g_end = G;
        if (item >= n_whole) {
            const uint32_t j = item - n_whole, part = j & 3u;
            tile = n_whole + (j >> 2);
            g_begin = (int)((uint32_t)G * part / 4u);
            g_end = (int)((uint32_t)G * (part + 1u) / 4u);
        }
        const uint32_t doc0 = P.tile_doc_off[tile];
        const int nd = (int)(P.tile_doc_off[tile + 1] - doc0);
        const bool active = t < nd;
        const uint32_t pos = active ? P.pd_pos[doc0 + t] : 0u;
        const uint32_t task0 = F.tile_task_off[tile];
        const int ntask = (int)(F.tile_task_off[tile + 1] - task0);
        const uint32_t q0 = P.tile_q_off[tile];
        const int nqt = (int)(P.tile_q_off[tile + 1] - q0);
        if (t < ntask) s_tasks[t] = F.tasks[task0 + t];
        if (!use_tbl) s_gexp[t] = active ? __ldg(P.gexp + pos) : 0.0;
        s_cls[t] = active ? __ldg(F.pd_cls + doc0 + t) : (uint8_t)0;

        // ---- phase 1: one pass over the tile's features for every sweep ----
        // acc[s] = sum_j x_j * w_sj, j ascending, separate multiply and add; the host zeroed
        // w at the coordinate the sweep varies, whose feature value is fetched on its own.
        double acc[NS];
        float xf[NS];
        {
            const float *__restrict__ xp = P.x + pos;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s] = 0.0;
                xf[s] = fs[s] < dm ? ld_stream(xp + (size_t)fs[s] * P.ld) : 0.f;
            }
            float cur[8], nxt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] = (uint32_t)u < dm ? ld_stream(xp + (size_t)u * P.ld) : 0.f;
            // the weight table is zero-padded to a multiple of 8 coordinates: no guards on the math
            for (uint32_t j0 = 0; j0 < dm; j0 += 8) {
                if (j0 + 16 <= dm) {  // prefetch the next block while this one is consumed
#pragma unroll
                    for (int u = 0; u < 8; ++u) nxt[u] = ld_stream(xp + (size_t)(j0 + 8 + u) * P.ld);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        nxt[u] = j0 + 8 + u < dm ? ld_stream(xp + (size_t)(j0 + 8 + u) * P.ld) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double xd = (double)cur[u];
                    const double2 *__restrict__ wj =
                        (const double2 *)((WS ? (const double *)s_w : wg) + (size_t)(j0 + u) * NS);
#pragma unroll
                    for (int s = 0; s < NS; s += 2) {
                        const double2 w2 = WS ? wj[s >> 1] : __ldg(wj + (s >> 1));
                        acc[s] = __dadd_rn(acc[s], __dmul_rn(xd, w2.x));
                        acc[s + 1] = __dadd_rn(acc[s + 1], __dmul_rn(xd, w2.y));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
            }
        }

        // scores of row group g -> s_score, and clear the slot buffer the group will use
        auto score_group = [&](int g) {
            const int rbeg = g << 5, rend = min(R, rbeg + 32);
            double *col = s_score + t;
            int r = rbeg;
            while (r < rend) {
                const int s = s_rowsw[r];
                const int seg_end = min(rend, s_misc[M_SWROW + s + 1]);
                double ss = acc[0];
                float xs32 = xf[0];
#pragma unroll
                for (int u = 1; u < NS; ++u) {
                    if (s == u) {
                        ss = acc[u];
                        xs32 = xf[u];
                    }
                }
                const double xs = (double)xs32;
#pragma unroll 8
                for (; r < seg_end; ++r) {
                    const double sc = __dadd_rn(ss, __dmul_rn(xs, __ldg(A.row_w + row0 + r)));
                    if (sc != sc) nan_seen |= active ? 1 : 0;
                    col[(size_t)(r - rbeg) * ROW] = sc;
                }
            }
            uint4 *z = (uint4 *)(s_slot0 + (size_t)(g & 1) * slot_stride);
            constexpr int nz = (int)((sizeof(slot_t) * TB * 32) / sizeof(uint4));
#pragma unroll
            for (int i = t; i < nz; i += TB) z[i] = make_uint4(0u, 0u, 0u, 0u);
            if (t == 0) s_misc[M_CTR] = 0;
        };

        // one warp folds one query of row group gq: slots in rank order -> value -> fixed point
        auto fold_query = [&](int ql, int gq) {
            const slot_t *slots = s_slot0 + (size_t)(gq & 1) * slot_stride;
            const int nrow = min(32, R - (gq << 5));
            const int g = gq;
            const uint32_t pq = q0 + ql;
            const uint32_t loc = P.pq_local[pq];
            const uint32_t start = loc & 0xffffu, len = loc >> 16;
            const double norm = P.pq_norm[pq];
            const slot_t *sl = slots + (size_t)start * 32 + lane;
            double value = 0.0;
            if (P.metric == FR_METRIC_NDCG) {
                if (norm == norm) {  // Some(ideal), evaluators.rs:351-358
                    const uint32_t lim = len < (uint32_t)P.depth ? len : (uint32_t)P.depth;
                    double dcg = 0.0;
                    for (uint32_t r0 = 0; r0 < lim; r0 += 8) {
                        unsigned id[8];
                        double term[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) id[u] = r0 + u < lim ? sl[(size_t)(r0 + u) * 32] : 0u;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            term[u] = 0.0;
                            if (id[u]) {
                                const uint32_t r = r0 + u;
                                if (use_tbl)
                                    term[u] = tbl[(size_t)(id[u] - 1) * F.tbl_r + r];
                                else
                                    term[u] = s_gexp[id[u] - 1] / __ldg(P.lg2 + r);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) dcg = __dadd_rn(dcg, term[u]);
                    }
                    if (dcg > norm) atomicOr(A.err, ERR_DCG_ABOVE_IDEAL);
                    value = dcg / norm;
                }
            } else if (P.metric == FR_METRIC_AP) {
                if (norm > 0.0) {
                    unsigned recall = 0;
                    double sum = 0.0;
                    for (uint32_t r = 0; r < len; ++r) {
                        if (sl[(size_t)r * 32]) {
                            recall += 1;
                            sum = __dadd_rn(sum, (double)recall / (double)(r + 1));
                        }
                    }
                    value = sum / norm;
                }
            } else {
                for (uint32_t r = 0; r < len; ++r) {
                    if (sl[(size_t)r * 32]) {
                        value = 1.0 / (double)(r + 1);
                        break;
                    }
                }
            }
            if (lane < nrow) {
                const int row = (g << 5) + lane;
                if (A.perq)
                    A.perq[(size_t)A.row_out[row0 + row] * P.nq_view + P.pq_view[pq]] = value;
                const long long fx = __double2ll_rn(value * kFx);
                atomicAdd(&s_sum[row], (unsigned long long)fx);
            }
                };

        score_group(g_begin);
        __syncthreads();
        for (int g = g_begin; g < g_end; ++g) {
            slot_t *slots = s_slot0 + (size_t)(g & 1) * slot_stride;
            const int nrow = min(32, R - (g << 5));  // live lanes of this group
            // -- rank by counting; lane = row, TD documents per task --
            {
                int ti = 0;
                if (lane == 0) ti = atomicAdd(&s_misc[M_CTR], 1);
                ti = __shfl_sync(0xffffffffu, ti, 0);
                // queue = the previous group's folds (latency-bound: slot and table reads, a serial
                // f64 sum, one division) followed by this group's ranking tasks (issue-bound)
                const int nfold = g > g_begin ? nqt : 0;
                while (ti < nfold + ntask) {
                    int tnext = 0;
                    if (lane == 0) tnext = atomicAdd(&s_misc[M_CTR], 1);
                    if (ti < nfold) {
                        fold_query(ti, g - 1);
                        ti = __shfl_sync(0xffffffffu, tnext, 0);
                        continue;
                    }
                    const uint2 tk = s_tasks[ti - nfold];
                    const int qs = (int)(tk.x & 0xffffu), qe = (int)(tk.x >> 16);
                    const int t0 = (int)(tk.y & 0xffffu), n = (int)(tk.y >> 16);
                    const unsigned len = (unsigned)(qe - qs);
                    const unsigned lim =
                        (P.metric == FR_METRIC_NDCG && (unsigned)P.depth < len) ? (unsigned)P.depth : len;
                    const bool tag_cls = P.metric == FR_METRIC_NDCG && use_tbl;
                    // a short last chunk of a query takes the half-width walk
                    if (TD > 4 && n <= TD / 2)
                        rank_task<(TD > 4 ? TD / 2 : TD), slot_t>(myrow, qs, qe, t0, n, lim, lane < nrow, tag_cls, s_cls,
                                                                  slots, lane);
                    else
                        rank_task<TD, slot_t>(myrow, qs, qe, t0, n, lim, lane < nrow, tag_cls, s_cls, slots, lane);
                    ti = __shfl_sync(0xffffffffu, tnext, 0);
                }
            }
            __syncthreads();
            if (g + 1 < g_end) score_group(g + 1);
            __syncthreads();
        }
        // the last group's fold has no ranking phase left to hide in
        for (int ql = warp; ql < nqt; ql += TB / 32) fold_query(ql, g_end - 1);
        next_item = (uint32_t)s_misc[M_NEXT];
        __syncthreads();
    }
    if (nan_seen) atomicOr(A.err, ERR_NAN_SCORE);
    __syncthreads();
    for (int idx = t; idx < R; idx += TB)
        atomicAdd((unsigned long long *)(A.sums + A.row_out[row0 + idx]), s_sum[idx]);
    if (A.mail_peers == nullptr && A.host_flag == nullptr) return;
    fused_allreduce_tail<TB>(A, s_misc);
}

template <int TB, int TD, bool WS>
int launch_fast(fr_dev_plan *pl, const FastArgs &a, uint32_t n_groups, cudaStream_t stream) {
    const uint32_t dm8 = (a.dm + 7) & ~7u;
    const SmemLayout L(TB, (int)sizeof(typename SlotType<TB>::type), WS ? dm8 * kMaxSweeps : 0u);
    auto kernel = sweep_fast_kernel<TB, TD, WS>;
    // attribute + occupancy queries cost several microseconds each: once per (instantiation,
    // device, shared-memory size), not once per launch
    static std::mutex mu;
    static std::map<std::pair<int, size_t>, int> cache;
    int occ = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_pair(pl->ds->device, L.total);
        auto it = cache.find(key);
        if (it == cache.end()) {
            // the attribute is a ceiling, not a reservation: raise it to the device maximum once
            // (setting it per size would LOWER it when a smaller layout comes by later)
            int optin = 0;
            CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pl->ds->device));
            CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TB, L.total));
            it = cache.emplace(key, occ).first;
        }
        occ = it->second;
    }
    if (occ < 1) return fail("sweep_fast_kernel does not fit on an SM");
    uint64_t total = (uint64_t)pl->sm_count * (uint64_t)occ;
    uint32_t gx = (uint32_t)std::max<uint64_t>(1, total / n_groups);
    if (gx > pl->nt) gx = std::max<uint32_t>(pl->nt, 1);  // an empty shard still takes part in the reduction
    PlanView pv = pl->view();
    FastArgs args = a;
    // quarter items: the last gx / 4 tiles, one final wave of short items (splitting every tile
    // of a small shard was measured and is slower: phase 1 is repeated per item)
    args.n_split = gx / 4u;
    FastView fv;
    fv.tile_task_off = pl->fast.tile_task_off.p;
    fv.tasks = pl->fast.tasks.p;
    fv.pd_cls = pl->fast.pd_cls.p;
    fv.disc_tbl = pl->fast.n_cls ? pl->fast.disc_tbl.p : nullptr;
    fv.tbl_r = pl->fast.tbl_r;
    fv.n_cls = pl->fast.n_cls;
    auto *ev = pl->ds->prof_slot();
    if (ev) cudaEventRecord(ev->first, stream);
    kernel<<<dim3(gx, n_groups), TB, L.total, stream>>>(pv, fv, args);
    if (ev) cudaEventRecord(ev->second, stream);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

#include "sweep_packed.cuh"

template <int TB, bool WS, int MINB, int PS, bool SLOTS = false>
int launch_packed(fr_dev_plan *pl, const FastArgs &a, uint32_t n_groups, cudaStream_t stream) {
    const uint32_t dm8 = (a.dm + 7) & ~7u;
    const PackedLayout L(TB, WS ? dm8 * kMaxSweeps : 0u, PS, SLOTS);
    auto kernel = sweep_packed_kernel<TB, WS, MINB, PS, SLOTS>;
    static std::mutex mu;
    static std::map<std::pair<int, size_t>, int> cache;
    int occ = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_pair(pl->ds->device, L.total);
        auto it = cache.find(key);
        if (it == cache.end()) {
            int optin = 0;
            CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pl->ds->device));
            CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TB, L.total));
            it = cache.emplace(key, occ).first;
        }
        occ = it->second;
    }
    if (occ < 1) return fail("sweep_packed_kernel does not fit on an SM");
    uint64_t total = (uint64_t)pl->sm_count * (uint64_t)occ;
    uint32_t gx = (uint32_t)std::max<uint64_t>(1, total / n_groups);
    if (gx > pl->nt) gx = std::max<uint32_t>(pl->nt, 1);  // an empty shard still takes part in the reduction
    PlanView pv = pl->view();
    FastArgs args = a;
    // Work granularity.  A CTA's unit is a tile (phase 1, then ~13 row groups x its queries as warp
    // items).  With many tiles per CTA the last gx / 4 tiles are handed out as quarters (ranges of row
    // groups, phase 1 repeated -- it is ~17 % of a tile) so that the SMs drain together; with fewer
    // than two tiles per CTA repeating phase 1 costs more than the shorter tail saves (measured on
    // 125 k .. 1 M document shards: 0.175 / 0.297 / 0.527 / 0.976 ms).
    args.split_parts = 4;
    args.n_split = (uint64_t)pl->nt >= 2ull * gx ? gx / 4u : 0u;
    if (const char *env = getenv("FASTRANK_NSPLIT_DIV")) {  // tuning knob: 0 = no split items, -1 = every tile
        const int div = atoi(env);
        args.n_split = div > 0 ? gx / (uint32_t)div : (div < 0 ? pl->nt : 0u);
    }
    if (const char *env = getenv("FASTRANK_NSPLIT")) args.n_split = (uint32_t)std::min<long>(std::max<long>(atol(env), 0), pl->nt);
    if (const char *env = getenv("FASTRANK_SPLIT_PARTS")) args.split_parts = atoi(env) == 2 ? 2u : 4u;
    PackedView v;
    v.q_task_off = pl->fast.pk_q_task_off.p;
    v.q_order = pl->fast.pk_q_order.p;
    v.tile_task_off = pl->fast.pk_tile_task_off.p;
    v.tasks = pl->fast.pk_tasks.p;
    v.pd_cls = pl->fast.pd_cls.p;
    v.tbl = pl->fast.pk_tbl.n ? pl->fast.pk_tbl.p : nullptr;
    v.ap_tbl = pl->fast.ap_tbl.n ? pl->fast.ap_tbl.p : nullptr;
    v.ap_cols = pl->fast.ap_cols;
    v.tbl_r = pl->fast.tbl_r;
    v.n_cls = pl->fast.n_cls;
    auto *ev = pl->ds->prof_slot();
    if (ev) cudaEventRecord(ev->first, stream);
    kernel<<<dim3(gx, n_groups), TB, L.total, stream>>>(pv, v, args);
    if (ev) cudaEventRecord(ev->second, stream);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

}  // namespace

namespace frbdev {

// Host half of the fast plan: warp tasks, gain classes, discount table.
int build_fast_plan(fr_dev_plan *pl, const std::vector<uint32_t> &tile_q_off,
                    const std::vector<uint32_t> &pq_local, const std::vector<uint32_t> &pq_doc0,
                    const std::vector<uint32_t> &pd_pos) {
    FastPlan &fp = pl->fast;
    fr_dev_dataset *ds = pl->ds;
    fp.ok = false;
    if (pl->tb > kFastTile) {
        fp.why = "too many documents sit in queries of more than 512 documents";
        return 0;
    }
    int td = 8;
    if (const char *env = getenv("FASTRANK_TD")) td = atoi(env) == 4 ? 4 : 8;
    fp.td = td;
    const bool ndcg = pl->metric == FR_METRIC_NDCG;
    // gain classes
    std::vector<uint32_t> cls_bits;  // bit pattern of class c's gain (flat: a handful of classes)
    std::vector<float> cls_gain;
    std::vector<uint8_t> pd_cls(pd_pos.size(), 0);
    bool table = ndcg;
    if (table) {
        size_t last = 0;
        for (size_t i = 0; i < pd_pos.size(); ++i) {
            const float g = ds->gain_pos[pd_pos[i]];
            uint32_t bits;
            memcpy(&bits, &g, 4);
            if (cls_bits.empty() || cls_bits[last] != bits) {
                size_t at = 0;
                while (at < cls_bits.size() && cls_bits[at] != bits) ++at;
                if (at == cls_bits.size()) {
                    if (cls_gain.size() >= 255) {
                        table = false;
                        break;
                    }
                    cls_bits.push_back(bits);
                    cls_gain.push_back(g);
                }
                last = at;
            }
            pd_cls[i] = (uint8_t)last;
        }
    }
    cudaStream_t s = ds->stream;
    UploadBatch up;  // every array of the sweep plan in one allocation
    std::vector<double> tbl_host, tbl0, ap_host;
    std::vector<uint32_t> q_task_off{0}, pk_tile_off{0};
    std::vector<uint4> pk_tasks;
    std::vector<uint16_t> q_order;
    fp.n_cls = 0;
    fp.tbl_r = 1;
    if (table && !cls_gain.empty()) {
        const uint32_t R = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)pl->depth, pl->max_len));
        std::vector<double> tbl(cls_gain.size() * (size_t)R);
        for (size_t c = 0; c < cls_gain.size(); ++c) {
            const double ge = std::pow(2.0, (double)cls_gain[c]) - 1.0;  // evaluators.rs:268
            for (uint32_t r = 0; r < R; ++r) tbl[c * R + r] = ge / std::log2((double)r + 2.0);
        }
        fp.n_cls = (uint32_t)cls_gain.size();
        fp.tbl_r = R;
        tbl_host = tbl;
    } else {
        std::fill(pd_cls.begin(), pd_cls.end(), 0);
    }
    if (!tbl_host.empty()) up.add(fp.disc_tbl, tbl_host);
    up.add(fp.pd_cls, pd_cls);
    // warp tasks: runs of documents that can contribute, cut into chunks of td
    std::vector<uint32_t> tile_task_off{0};
    std::vector<uint2> tasks;
    for (uint32_t tile = 0; tile + 1 < tile_q_off.size(); ++tile) {
        for (uint32_t pq = tile_q_off[tile]; pq < tile_q_off[tile + 1]; ++pq) {
            const uint32_t start = pq_local[pq] & 0xffffu, len = pq_local[pq] >> 16;
            uint32_t i = 0;
            while (i < len) {
                const float g = ds->gain_pos[pd_pos[pq_doc0[pq] + i]];
                const bool contributes = ndcg ? (g != 0.0f) : (g > 0.0f);
                if (!contributes) {
                    ++i;
                    continue;
                }
                uint32_t n = 1;
                while (n < (uint32_t)td && i + n < len) {
                    const float g2 = ds->gain_pos[pd_pos[pq_doc0[pq] + i + n]];
                    if (!(ndcg ? (g2 != 0.0f) : (g2 > 0.0f))) break;
                    ++n;
                }
                uint2 tk;
                tk.x = start | ((start + len) << 16);
                tk.y = (start + i) | (n << 16);
                tasks.push_back(tk);
                i += n;
            }
        }
        // longest task first: the tile's warps pull tasks from a shared queue, and the phase ends
        // at a barrier, so the long walks must not be the ones left for last
        std::stable_sort(tasks.begin() + tile_task_off.back(), tasks.end(), [](const uint2 &a, const uint2 &b) {
            const uint32_t ca = ((a.x >> 16) - (a.x & 0xffffu)) * (8 + (a.y >> 16));
            const uint32_t cb = ((b.x >> 16) - (b.x & 0xffffu)) * (8 + (b.y >> 16));
            return ca > cb;
        });
        tile_task_off.push_back((uint32_t)tasks.size());
    }
    fp.n_tasks = (uint32_t)tasks.size();
    up.add(fp.tile_task_off, tile_task_off);
    up.add(fp.tasks, tasks);
    fp.ok = true;
    // sweep_packed_kernel: NDCG@k, k <= 16 ranks of 4 bits in one register per candidate
    fp.packed_ok = false;
    fp.slots_ok = false;
    fp.len_docs.assign((size_t)pl->tb + 1, 0);
    for (uint32_t pq = 0; pq < pq_local.size(); ++pq) fp.len_docs[pq_local[pq] >> 16] += pq_local[pq] >> 16;
    const bool can_pack = ndcg && fp.n_cls >= 1 && fp.n_cls <= 15 && pl->depth <= 16;
    // slot mode: NDCG needs the class table (<= 255 classes), AP / RR need nothing; 256-document tiles at most
    const bool can_slot = pl->tb <= 256 && (!ndcg || fp.n_cls >= 1);
    if (can_pack || can_slot) {
        q_order.assign(pq_local.size(), 0);
        std::vector<std::pair<uint64_t, uint16_t>> cost;
        for (uint32_t tile = 0; tile + 1 < tile_q_off.size(); ++tile) {
            cost.clear();
            for (uint32_t pq = tile_q_off[tile]; pq < tile_q_off[tile + 1]; ++pq) {
                const uint32_t start = pq_local[pq] & 0xffffu, len = pq_local[pq] >> 16;
                uint64_t qcost = 64;  // fold + bookkeeping
                uint32_t i = 0;
                auto contributes = [&](uint32_t k) {  // 2^gain - 1 != 0 for NDCG, relevant for AP / RR
                    const float g = ds->gain_pos[pd_pos[pq_doc0[pq] + k]];
                    return ndcg ? g != 0.0f : g > 0.0f;
                };
                while (i < len) {
                    if (!contributes(i)) {
                        ++i;
                        continue;
                    }
                    uint32_t run = 1;
                    while (i + run < len && contributes(i + run)) ++run;
                    // chunks of 16 / 8 / 4: a walk of W documents costs len * (3 + 2 W) instructions
                    while (run > 0) {
                        const uint32_t n = run >= 13 ? std::min<uint32_t>(run, 16) : (run > 8 ? 8 : run);
                        const uint32_t w = n > 8 ? 16 : (n > 4 ? 8 : 4);
                        unsigned long long tags = 0;  // gain class + 1 of each document, 4 bits apiece (register mode)
                        for (uint32_t u = 0; u < n && can_pack; ++u)
                            tags |= (unsigned long long)(pd_cls[pq_doc0[pq] + i + u] + 1u) << (4 * u);
                        pk_tasks.push_back(make_uint4((start + i) | (n << 16), (uint32_t)tags, (uint32_t)(tags >> 32), 0u));
                        qcost += (uint64_t)len * (3 + 2 * w) + (uint64_t)w * w;
                        i += n;
                        run -= n;
                    }
                }
                q_task_off.push_back((uint32_t)pk_tasks.size());
                cost.emplace_back(qcost, (uint16_t)(pq - tile_q_off[tile]));
            }
            std::stable_sort(cost.begin(), cost.end(),
                             [](const std::pair<uint64_t, uint16_t> &a, const std::pair<uint64_t, uint16_t> &b) {
                                 return a.first > b.first;
                             });
            for (size_t k = 0; k < cost.size(); ++k) q_order[tile_q_off[tile] + k] = cost[k].second;
            pk_tile_off.push_back((uint32_t)pk_tasks.size());
        }
        // table with a leading row of zeros: tag 0 = nothing ranked there
        if (ndcg) {
            tbl0.assign((size_t)(fp.n_cls + 1) * fp.tbl_r, 0.0);
            for (size_t c = 0; c < cls_gain.size(); ++c) {
                const double ge = std::pow(2.0, (double)cls_gain[c]) - 1.0;  // evaluators.rs:268
                for (uint32_t r = 0; r < fp.tbl_r; ++r) tbl0[(c + 1) * fp.tbl_r + r] = ge / std::log2((double)r + 2.0);
            }
        }
        up.add(fp.pk_q_task_off, q_task_off);
        up.add(fp.pk_tile_task_off, pk_tile_off);
        up.add(fp.pk_tasks, pk_tasks);
        up.add(fp.pk_q_order, q_order);
        if (!tbl0.empty()) up.add(fp.pk_tbl, tbl0);
        if (pl->metric == FR_METRIC_AP && can_slot) {
            // precision table: recall / (rank + 1) for every (relevant documents so far, rank) the plan's
            // lists can produce -- the same IEEE divisions the reference performs (evaluators.rs:440-444)
            uint32_t max_rel = 0;
            for (uint32_t pq = 0; pq < pq_local.size(); ++pq) {
                uint32_t rel = 0;
                for (uint32_t k = 0; k < (pq_local[pq] >> 16); ++k) rel += ds->gain_pos[pd_pos[pq_doc0[pq] + k]] > 0.0f;
                max_rel = std::max(max_rel, rel);
            }
            const size_t cols = std::max<uint32_t>(pl->max_len, 1), rows = (size_t)max_rel + 1;
            if (rows * cols <= ((size_t)1 << 19)) {  // <= 4 MB
                ap_host.assign(rows * cols, 0.0);
                for (size_t rc = 1; rc < rows; ++rc)
                    for (size_t r = 0; r < cols; ++r) ap_host[rc * cols + r] = (double)rc / (double)(r + 1);
                fp.ap_cols = (uint32_t)cols;
                up.add(fp.ap_tbl, ap_host);
            }
        }
        fp.packed_ok = can_pack;
        fp.slots_ok = can_slot;
    }
    CU(up.commit(pl->fast_arena, s));
    CU(cudaStreamSynchronize(s));  // the host vectors above go out of scope
    return 0;
}

}  // namespace frbdev

extern "C" int fr_dev_plan_has_fast_sweep(const fr_dev_plan *plan) { return plan && plan->fast.ok ? 1 : 0; }

extern "C" const char *fr_dev_plan_sweep_kernel(const fr_dev_plan *plan) {
    if (!plan || !plan->fast.ok) return "";
    bool packed = plan->fast.packed_ok, slots = !packed && plan->fast.slots_ok;
    if (const char *env = getenv("FASTRANK_SWEEP_KERNEL")) {
        const std::string want(env);
        if (want == "tile") packed = slots = false;
        if (want == "slots" && plan->fast.slots_ok) {
            packed = false;
            slots = true;
        }
    }
    if (packed) return plan->tb == 128 ? "sweep_packed_kernel<128>" : plan->tb == 256 ? "sweep_packed_kernel<256>" : "sweep_packed_kernel<512>";
    if (slots) return plan->tb == 128 ? "sweep_packed_kernel<128,slots>" : "sweep_packed_kernel<256,slots>";
    return plan->tb == 128 ? "sweep_fast_kernel<128,8>" : plan->tb == 256 ? "sweep_fast_kernel<256,8>" : "sweep_fast_kernel<512,8>";
}

extern "C" int fr_dev_eval_coord_sweeps_fast(fr_dev_plan *pl, size_t n_sweeps, const double *base_w,
                                             size_t wlen, const uint32_t *fid, const double *cand_w,
                                             const uint32_t *n_cand, size_t cand_stride,
                                             int64_t *out_sum_fx, double *out_per_query) {
    if (!pl || !base_w || !fid || !cand_w || !n_cand || !out_sum_fx)
        return fail("fr_dev_eval_coord_sweeps_fast: NULL argument");
    if (!pl->fast.ok)
        return fail("fr_dev_eval_coord_sweeps_fast: not available for this plan (" + pl->fast.why + ")");
    if (n_sweeps == 0) return 0;
    fr_dev_dataset *ds = pl->ds;
    FastPlan &fp = pl->fast;
    static const bool trace_steps = getenv("FASTRANK_TRACE_STEPS") != nullptr;
    static thread_local double tr_fill = 0, tr_submit = 0, tr_wait = 0, tr_between = 0, tr_copy = 0;
    static thread_local uint64_t tr_calls = 0;
    static thread_local std::chrono::steady_clock::time_point tr_last_exit;
    const auto tr_enter = std::chrono::steady_clock::now();
    if (trace_steps && tr_calls > 0) tr_between += std::chrono::duration<double, std::micro>(tr_enter - tr_last_exit).count();
    CU(cudaSetDevice(ds->device));
    cudaStream_t s = ds->stream;
    const size_t total = n_sweeps * cand_stride;
    // flatten the candidates into rows, sorted by sweep; a sweep group holds at most
    // kMaxSweeps sweeps and kMaxRows rows, so a group with more candidates is served in passes
    std::vector<uint32_t> cursor(n_sweeps, 0);
    for (size_t r = 0; r < n_sweeps; ++r)
        if (n_cand[r] > cand_stride) return fail("fr_dev_eval_coord_sweeps_fast: n_cand > cand_stride");
    const uint32_t n_groups = (uint32_t)((n_sweeps + kMaxSweeps - 1) / kMaxSweeps);
    const size_t dm = std::min<size_t>(wlen, ds->d);
    const size_t dm8 = std::max<size_t>((dm + 7) & ~(size_t)7, 8);
    const size_t wt_count = (size_t)n_groups * dm8 * kMaxSweeps;
    const size_t max_rows = std::min<size_t>(total, (size_t)n_groups * kMaxRows);
    // One pinned staging blob -> one H2D copy per pass:
    //   [wt doubles][row_w doubles][fid u32][row_meta u32][row_out u32][grp_off u32]
    const size_t in_bytes = sizeof(double) * (wt_count + max_rows) +
                            sizeof(uint32_t) * (n_sweeps + 2 * max_rows + n_groups + 1);
    // One device blob cleared with one memset and read back with one D2H copy:
    //   [sums i64 x total][err i32][pad i32][tile counters u32 x n_groups]
    const size_t out_bytes = sizeof(long long) * total + 8 + sizeof(unsigned) * (n_groups + 1);
    CU(fp.in_dev.ensure(in_bytes));
    CU(fp.in_host.ensure(in_bytes));
    if (out_per_query) CU(pl->perq_dev.ensure(total * (size_t)pl->nq_view));
    // passes needed (a sweep group holds kMaxRows candidate rows per pass); the cross-GPU
    // reduction is fused into the last one
    size_t n_passes = 1;
    for (uint32_t g = 0; g < n_groups; ++g) {
        size_t rows = 0;
        for (size_t sw = (size_t)g * kMaxSweeps; sw < std::min(n_sweeps, ((size_t)g + 1) * kMaxSweeps); ++sw)
            rows += n_cand[sw];
        n_passes = std::max(n_passes, (rows + kMaxRows - 1) / kMaxRows);
    }
    fr_dev_comm *comm = pl->comm && pl->comm->world > 1 ? pl->comm : nullptr;
    const bool fuse = comm && comm->mail.ok && total <= kMailWords;
    // Direct publication: the kernel's last CTA hands the sums to the host through mapped pinned
    // memory and re-zeroes the device state, so a step is one small H2D copy and one launch.
    bool direct = !out_per_query && total <= 4096 && n_groups <= kDirectGroups && n_passes == 1 && pl->nt > 0 &&
                  (!comm || fuse);
    if (const char *env = getenv("FASTRANK_DIRECT")) direct = direct && atoi(env) != 0;
    if (direct && !fp.pub_host) {
        if (cudaHostAlloc((void **)&fp.pub_host, kPubErrOff + 8, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&fp.pub_dev, fp.pub_host, 0) != cudaSuccess) {
            cudaGetLastError();
            if (fp.pub_host) cudaFreeHost(fp.pub_host);
            fp.pub_host = nullptr;
            direct = false;
        } else {
            memset(fp.pub_host, 0, kPubErrOff + 8);
            CU(fp.state_dev.alloc(kDirectStateBytes));
            CU(cudaMemsetAsync(fp.state_dev.p, 0, kDirectStateBytes, s));
        }
    }
    long long *sums_dev;
    int *err_dev;
    unsigned *ctr_dev, *done_dev;
    if (direct) {
        if (fp.direct_open) {  // an earlier call failed between launch and publication
            CU(cudaStreamSynchronize(s));
            CU(cudaMemsetAsync(fp.state_dev.p, 0, kDirectStateBytes, s));
        }
        fp.direct_open = true;
        err_dev = (int *)fp.state_dev.p;
        done_dev = (unsigned *)(fp.state_dev.p + 4);
        ctr_dev = (unsigned *)(fp.state_dev.p + 8);
        sums_dev = (long long *)(fp.state_dev.p + kDirectSumsOff);
    } else {
        CU(fp.out_dev.ensure(out_bytes));
        CU(fp.out_host.ensure(sizeof(long long) * total + 8));
        sums_dev = (long long *)fp.out_dev.p;
        err_dev = (int *)(fp.out_dev.p + sizeof(long long) * total);
        ctr_dev = (unsigned *)(fp.out_dev.p + sizeof(long long) * total + 8);
        done_dev = ctr_dev + n_groups;
    }
    // NDCG@k (k <= 16) takes the register-packed kernel; it tests for NaN scores only where T or
    // x_f is not finite, so candidates that are not finite themselves go to the general kernel
    bool use_packed = fp.packed_ok, use_slots = !fp.packed_ok && fp.slots_ok;
    for (size_t r = 0; r < n_sweeps && (use_packed || use_slots); ++r)
        for (uint32_t k = 0; k < n_cand[r]; ++k)
            if (!std::isfinite(cand_w[r * cand_stride + k])) use_packed = use_slots = false;
    if (const char *env = getenv("FASTRANK_SWEEP_KERNEL")) {  // test knob: "tile" forces the general kernel,
        const std::string want(env);                          //            "slots" the slot-buffer instance
        if (want == "tile") use_packed = use_slots = false;
        if (want == "slots" && fp.slots_ok && (use_packed || use_slots)) {
            use_packed = false;
            use_slots = true;
        }
    }
    size_t pass = 0;
    if (!direct) CU(cudaMemsetAsync(fp.out_dev.p, 0, out_bytes, s));
    if (out_per_query)
        CU(cudaMemsetAsync(pl->perq_dev.p, 0, sizeof(double) * total * (size_t)pl->nq_view, s));
    bool first_pass = true;
    for (;;) {
        if (!first_pass) {
            bool more = false;  // rows left for another pass?
            for (size_t sw = 0; sw < n_sweeps && !more; ++sw) more = cursor[sw] < n_cand[sw];
            if (!more) break;
            CU(cudaStreamSynchronize(s));  // the staging blob is about to be rewritten
        }
        unsigned char *hp = fp.in_host.p;
        double *h_wt = (double *)hp;
        double *h_row_w = h_wt + wt_count;
        uint32_t *h_fid = (uint32_t *)(h_row_w + max_rows);
        uint32_t *h_row_meta = h_fid + n_sweeps;
        uint32_t *h_row_out = h_row_meta + max_rows;
        uint32_t *h_grp = h_row_out + max_rows;
        size_t nrows = 0;
        h_grp[0] = 0;
        bool any = false;
        for (uint32_t g = 0; g < n_groups; ++g) {
            const size_t sw0 = (size_t)g * kMaxSweeps, sw1 = std::min(n_sweeps, sw0 + kMaxSweeps);
            uint32_t room = kMaxRows;
            for (size_t sw = sw0; sw < sw1 && room > 0; ++sw) {
                while (cursor[sw] < n_cand[sw] && room > 0) {
                    h_row_w[nrows] = cand_w[sw * cand_stride + cursor[sw]];
                    h_row_meta[nrows] = (uint32_t)(sw - sw0);
                    h_row_out[nrows] = (uint32_t)(sw * cand_stride + cursor[sw]);
                    ++nrows;
                    ++cursor[sw];
                    --room;
                    any = true;
                }
            }
            h_grp[g + 1] = (uint32_t)nrows;
        }
        if (!any) break;
        // base weights transposed per sweep group: wt[g][j][s]; unused sweep columns stay zero,
        // and so does the coordinate each sweep varies
        memset(h_wt, 0, sizeof(double) * wt_count);
        for (size_t sw = 0; sw < n_sweeps; ++sw) {
            double *col = h_wt + (sw / kMaxSweeps) * dm8 * kMaxSweeps + sw % kMaxSweeps;
            for (size_t j = 0; j < dm; ++j) col[j * kMaxSweeps] = j == fid[sw] ? 0.0 : base_w[sw * wlen + j];
            h_fid[sw] = fid[sw];
        }
        const auto tr_filled = std::chrono::steady_clock::now();
        if (trace_steps) tr_fill += std::chrono::duration<double, std::micro>(tr_filled - tr_enter).count();
        CU(cudaMemcpyAsync(fp.in_dev.p, hp, in_bytes, cudaMemcpyHostToDevice, s));
        if (trace_steps) tr_copy += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr_filled).count();
        if (!first_pass) CU(cudaMemsetAsync(ctr_dev, 0, sizeof(unsigned) * n_groups, s));
        first_pass = false;
        unsigned char *dp = fp.in_dev.p;
        FastArgs a;
        a.base_wt = (const double *)dp;
        a.row_w = a.base_wt + wt_count;
        a.fid = (const uint32_t *)(a.row_w + max_rows);
        a.row_meta = a.fid + n_sweeps;
        a.row_out = a.row_meta + max_rows;
        a.grp_row_off = a.row_out + max_rows;
        a.sums = sums_dev;
        a.perq = out_per_query ? pl->perq_dev.p : nullptr;
        a.n_sweeps = (uint32_t)n_sweeps;
        a.wlen = (uint32_t)wlen;
        a.dm = (uint32_t)dm;
        a.err = err_dev;
        a.tile_ctr = ctr_dev;
        ++pass;
        const bool fuse_now = fuse && pass == n_passes;
        a.mail_peers = fuse_now ? comm->mail.peer_dev.p : nullptr;
        a.done_ctr = done_dev;
        a.mail_rank = comm ? (uint32_t)comm->rank : 0u;
        a.mail_world = comm ? (uint32_t)comm->world : 1u;
        a.mail_epoch = fuse_now ? ++comm->mail.epoch : 0u;
        a.mail_words = (uint32_t)total;
        a.prune_min = 48;
        if (const char *env = getenv("FASTRANK_PRUNE_MIN")) a.prune_min = (uint32_t)std::max(0, atoi(env));
        {
            static const long long timeout_cycles = [] {
                double sec = 30.0;  // rank skew (a slow stdout, a debugger, preemption) is not an error
                if (const char *env = getenv("FASTRANK_PEER_TIMEOUT_S")) sec = std::max(0.001, atof(env));
                return (long long)(sec * 2.0e9);
            }();
            a.mail_timeout_cycles = timeout_cycles;
        }
        a.host_sums = direct ? (long long *)fp.pub_dev : nullptr;
        a.host_err = direct ? (int *)(fp.pub_dev + kPubErrOff) : nullptr;
        a.host_flag = direct ? (unsigned *)(fp.pub_dev + kPubErrOff + 4) : nullptr;
        a.host_epoch = direct ? ++fp.pub_epoch : 0u;
        // lists that do not fit a tile are scored from the same staged tables and ranked from HBM
        // (long_queries.cu); their sums land in sums_dev BEFORE the tile kernel runs, so the fused
        // cross-GPU reduction in its tail covers them as well
        if (pl->lng.n_long > 0 &&
            eval_long_sweep(pl, a.base_wt, a.fid, (uint32_t)n_sweeps, a.row_w, a.row_meta, a.row_out, a.grp_row_off,
                            n_groups, (uint32_t)nrows, (uint32_t)dm, (uint32_t)dm8, sums_dev, a.perq, err_dev, s))
            return 1;
        if (pl->nt == 0 && !fuse_now) continue;
        // the weight table is staged in shared memory while that costs no resident CTA
        bool ws = SmemLayout(128, 1, ((a.dm + 7) & ~7u) * kMaxSweeps).total <= 56 * 1024;
        if (const char *env = getenv("FASTRANK_WSMEM")) ws = atoi(env) != 0;
        int rc;
        if (use_packed) {
            // select-then-rank (sweep_packed.cuh) only where lists are long enough to pay for it
            const bool prune = a.prune_min > 0 && fp.long_list_docs_frac(a.prune_min) >= 0.10;
            // resident CTAs per SM at 128 threads: 4 (<= 128 registers), 5 (<= 96) or 6 (<= 80)
            int minb = prune ? 4 : 5;
            if (const char *env = getenv("FASTRANK_PACKED_MINB")) minb = atoi(env);
            minb = std::max(4, std::min(minb, 6));
            const size_t pk_bytes = PackedLayout(pl->tb, ((a.dm + 7) & ~7u) * kMaxSweeps, prune ? kPruneSlots : 0).total;
            const size_t room = pl->tb == 128 ? (size_t)(226 * 1024) / (size_t)minb - 1024
                                              : (pl->tb == 256 ? (size_t)110 * 1024 : (size_t)220 * 1024);
            bool wsp = pk_bytes <= room;
            if (const char *env = getenv("FASTRANK_WSMEM")) wsp = atoi(env) != 0;
#define PK_LAUNCH(TB_, MINB_)                                                                                  \
    (prune ? (wsp ? launch_packed<TB_, true, MINB_, kPruneSlots>(pl, a, n_groups, s)                           \
                  : launch_packed<TB_, false, MINB_, kPruneSlots>(pl, a, n_groups, s))                         \
           : (wsp ? launch_packed<TB_, true, MINB_, 0>(pl, a, n_groups, s)                                     \
                  : launch_packed<TB_, false, MINB_, 0>(pl, a, n_groups, s)))
            if (pl->tb == 128)
                rc = minb <= 4 ? PK_LAUNCH(128, 4) : (minb == 5 ? PK_LAUNCH(128, 5) : PK_LAUNCH(128, 6));
            else if (pl->tb == 256)
                rc = PK_LAUNCH(256, 2);
            else
                rc = PK_LAUNCH(512, 1);
#undef PK_LAUNCH
        } else if (use_slots) {
            // any measure, any cut-off: ranks filed in a per-warp slot buffer (4 CTAs per SM at 128 threads)
            const size_t sl_bytes = PackedLayout(pl->tb, ((a.dm + 7) & ~7u) * kMaxSweeps, 0, true).total;
            bool wsp = sl_bytes <= (pl->tb == 128 ? (size_t)55 * 1024 : (size_t)110 * 1024);
            if (const char *env = getenv("FASTRANK_WSMEM")) wsp = atoi(env) != 0;
            if (pl->tb == 128)
                rc = wsp ? launch_packed<128, true, 4, 0, true>(pl, a, n_groups, s)
                         : launch_packed<128, false, 4, 0, true>(pl, a, n_groups, s);
            else
                rc = wsp ? launch_packed<256, true, 2, 0, true>(pl, a, n_groups, s)
                         : launch_packed<256, false, 2, 0, true>(pl, a, n_groups, s);
        } else if (pl->tb == 128) {
            if (ws)
                rc = fp.td == 8 ? launch_fast<128, 8, true>(pl, a, n_groups, s)
                                : launch_fast<128, 4, true>(pl, a, n_groups, s);
            else
                rc = fp.td == 8 ? launch_fast<128, 8, false>(pl, a, n_groups, s)
                                : launch_fast<128, 4, false>(pl, a, n_groups, s);
        } else if (pl->tb == 256) {
            rc = fp.td == 8 ? launch_fast<256, 8, false>(pl, a, n_groups, s)
                            : launch_fast<256, 4, false>(pl, a, n_groups, s);
        } else {
            rc = fp.td == 8 ? launch_fast<512, 8, false>(pl, a, n_groups, s)
                            : launch_fast<512, 4, false>(pl, a, n_groups, s);
        }
        if (rc) return 1;
    }
    if (direct) {
        const auto tr_submitted = std::chrono::steady_clock::now();
        // spin on the flag the last CTA raises (a stream query now and then catches a failed launch)
        volatile unsigned *flag = (volatile unsigned *)(fp.pub_host + kPubErrOff + 4);
        const unsigned want = fp.pub_epoch;
        // One process per GPU: with several ranks on a host the driver's own wait (which backs off
        // when the CPUs are oversubscribed) beats a spinning thread per rank -- measured at 8 GPUs:
        // 0.247 ms per step against 0.272 ms; alone on the host the spin saves ~15 us per step.
        if (comm) CU(cudaStreamSynchronize(s));
        for (unsigned spins = 0; *flag != want; ++spins) {
            if ((spins & 0xfff) == 0xfff) {
                const cudaError_t q = cudaStreamQuery(s);
                if (q != cudaErrorNotReady) {
                    if (q != cudaSuccess) return fail(std::string("sweep kernel failed: ") + cudaGetErrorString(q));
                    if (*flag != want) return fail("sweep kernel finished without publishing its sums");
                }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xff) == 0xff) sched_yield();  // one process per GPU: do not starve the other ranks' threads
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        fp.direct_open = false;
        if (trace_steps) {
            const auto tr_done = std::chrono::steady_clock::now();
            tr_submit += std::chrono::duration<double, std::micro>(tr_submitted - tr_enter).count();
            tr_wait += std::chrono::duration<double, std::micro>(tr_done - tr_submitted).count();
            tr_last_exit = tr_done;
            if (++tr_calls % 200 == 0) {
                fprintf(stderr, "[fastrank_b200] sweep steps: fill %.1f us, H2D call %.1f us, fill+submit %.1f us, wait %.1f us, between calls %.1f us (avg of %llu)\n",
                        tr_fill / tr_calls, tr_copy / tr_calls, tr_submit / tr_calls, tr_wait / tr_calls, tr_between / tr_calls,
                        (unsigned long long)tr_calls);
            }
        }
        int err_flags;
        memcpy(&err_flags, fp.pub_host + kPubErrOff, sizeof(int));
        if (check_err_flags(err_flags)) return 1;
        memcpy(out_sum_fx, fp.pub_host, sizeof(long long) * total);
        return 0;
    }
    if (!fuse && allreduce_sums(pl, sums_dev, total, s)) return 1;
    CU(cudaMemcpyAsync(fp.out_host.p, fp.out_dev.p, sizeof(long long) * total + 8, cudaMemcpyDeviceToHost, s));
    if (out_per_query)
        CU(cudaMemcpyAsync(out_per_query, pl->perq_dev.p, sizeof(double) * total * (size_t)pl->nq_view,
                           cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    int err_flags;
    memcpy(&err_flags, fp.out_host.p + sizeof(long long) * total, sizeof(int));
    if (check_err_flags(err_flags)) return 1;
    memcpy(out_sum_fx, fp.out_host.p, sizeof(long long) * total);
    return 0;
}
