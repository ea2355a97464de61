// sweep_fast.cu -- the batched coordinate-ascent sweep: every restart's line search in ONE pass
// over the feature matrix (coordinate_ascent.rs:131-177 for all restarts of a global step).
//
// Reference work being replaced, per candidate weight vector (evaluators.rs:173-224):
//   score every document (dense_dataset.rs:67-76), sort each query by (score desc, gain asc,
//   id asc) (evaluators.rs:33-49), NDCG / AP / RR per query, mean.
//
// Shape of the kernel (one CTA = one tile of whole queries, <= TB documents):
//   phase 1  stream the tile's slice of X once (feature-major, coalesced): per document and
//            per sweep s, with f = the coordinate the sweep varies,
//                P = sum_{j<f} x_j w_j     xf = x_f     S = sum_{j>f} x_j w_j
//            all in f64 with separate multiply and add, j ascending -- P is bit-identical to the
//            reference's running sum when it reaches coordinate f.
//   phase 2  per sweep: score[k][t] = (P + xf * cand_k) + S for the <= 32 candidates k (lane =
//            candidate), then rank by counting: a warp task owns TD documents of one query that
//            can contribute to the metric (gain != 0 for NDCG, gain > 0 for AP / RR) and walks
//            the query once; documents before t count when score >= , documents after t when
//            score > -- the reference's tie-break, because tile order is (gain asc, id asc).
//            rank -> slot[rank] = document; one warp per query then folds the slots in rank
//            order (the reference's left-to-right f64 sums) and accumulates round(value * 2^40).
//
// Arithmetic contract ("fast" mode): candidate scores differ from the reference's left-to-right
// dot product only in that the suffix S is summed before it is added (one rounding apart, ~1e-16
// relative).  Whenever the induced ranking of a query is the same -- always, except for
// documents whose scores agree to the last bits -- the per-query metric is bit-identical to the
// oracle, because terms and their summation order are the reference's.  The exact-order kernel
// (device.cu coord_sweep_kernel) stays available behind fr_dev_eval_coord_sweeps.
#include "device_common.cuh"

namespace {

constexpr int kMaxSweeps = 8;  // sweeps sharing one pass over X (one blockIdx.y group)
constexpr double kFx = 1099511627776.0;
static_assert(FR_FX_BITS == 40, "kFx must match FR_FX_BITS");

struct FastArgs {
    const double *base_w;    // [n_sweeps][wlen]
    const uint32_t *fid;     // [n_sweeps]
    const double *cand_w;    // [n_sweeps][cand_stride]
    const uint32_t *n_cand;  // [n_sweeps]
    long long *sums;         // [n_sweeps][cand_stride]
    double *perq;            // nullptr or [n_sweeps][cand_stride][nq_view]
    uint32_t n_sweeps, wlen, cand_stride, cand_off;
    uint32_t kp;  // candidates per pass (score rows, slot row stride), <= 32
    uint32_t dm;  // min(wlen, features): zip() truncation of dense_dataset.rs:67-76
    int *err;
};

// cnt += (a >= b) / (a > b) as DSETP + predicated add (what nvcc emits for the C form is a
// three-instruction add / compare / undo sequence).
__device__ __forceinline__ void count_ge(unsigned &cnt, double a, double b) {
    asm("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt) : "d"(a), "d"(b));
}
__device__ __forceinline__ void count_gt(unsigned &cnt, double a, double b) {
    asm("{ .reg .pred p; setp.gt.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt) : "d"(a), "d"(b));
}

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

struct SmemLayout {
    size_t w, cand, score, gexp, sum, slot, tasks, cls, misc, total;
    __host__ __device__ SmemLayout(int tb, uint32_t dm, uint32_t kp) {
        size_t o = 0;
        w = o;      o += align16(sizeof(double) * (size_t)(dm ? dm : 1) * kMaxSweeps);
        cand = o;   o += sizeof(double) * kMaxSweeps * 32;
        score = o;  o += align16(sizeof(double) * (size_t)kp * (tb + 1));
        gexp = o;   o += sizeof(double) * tb;
        sum = o;    o += sizeof(unsigned long long) * kMaxSweeps * 32;
        slot = o;   o += align16(sizeof(uint16_t) * (size_t)tb * kp);
        tasks = o;  o += sizeof(uint2) * tb;
        cls = o;    o += align16(tb);
        misc = o;   o += 256;
        total = o;
    }
};

// misc words
enum { M_F = 0, M_K = 8, M_SPEC = 16, M_NSPEC = 25, M_CTR = 26 };

template <int TB, int TD>
__global__ void __launch_bounds__(TB, (TB == 128 ? 4 : 1))
sweep_fast_kernel(PlanView P, FastView F, FastArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NS = kMaxSweeps;
    constexpr int ROW = TB + 1;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t s0 = blockIdx.y * NS;
    const int ns = (int)min((uint32_t)NS, A.n_sweeps - s0);
    const uint32_t dm = A.dm, KP = A.kp;
    const SmemLayout L(TB, dm, KP);
    double *s_w = (double *)(smem_raw + L.w);  // [dm][NS]
    double *s_cand = (double *)(smem_raw + L.cand);
    double *s_score = (double *)(smem_raw + L.score);  // [KP][ROW]
    double *s_gexp = (double *)(smem_raw + L.gexp);
    unsigned long long *s_sum = (unsigned long long *)(smem_raw + L.sum);
    uint16_t *s_slot = (uint16_t *)(smem_raw + L.slot);  // [TB][KP]
    uint2 *s_tasks = (uint2 *)(smem_raw + L.tasks);
    uint8_t *s_cls = (uint8_t *)(smem_raw + L.cls);
    int *s_misc = (int *)(smem_raw + L.misc);

    // ---- per-launch setup: weights, candidates, the coordinates that need special handling
    for (uint32_t idx = t; idx < dm * NS; idx += TB) {
        const uint32_t j = idx / NS, s = idx % NS;
        s_w[idx] = (int)s < ns ? A.base_w[(size_t)(s0 + s) * A.wlen + j] : 0.0;
    }
    for (int idx = t; idx < NS * 32; idx += TB) {
        const int s = idx >> 5, k = idx & 31;
        double v = 0.0;
        if (s < ns) {
            const int K = (int)A.n_cand[s0 + s] - (int)A.cand_off;
            if (k < K) v = A.cand_w[(size_t)(s0 + s) * A.cand_stride + A.cand_off + k];
        }
        s_cand[idx] = v;
        s_sum[idx] = 0ull;
    }
    if (t == 0) {
        int nspec = 0;
        for (int s = 0; s < NS; ++s) {
            int K = 0;
            uint32_t f = 0xffffffffu;
            if (s < ns) {
                K = (int)A.n_cand[s0 + s] - (int)A.cand_off;
                K = K < 0 ? 0 : (K > (int)KP ? (int)KP : K);
                f = A.fid[s0 + s];
            }
            s_misc[M_K + s] = K;
            s_misc[M_F + s] = (int)f;
            if (K > 0 && f < dm) {  // insert into the sorted, de-duplicated list
                int p = 0;
                while (p < nspec && (uint32_t)s_misc[M_SPEC + p] < f) ++p;
                if (p == nspec || (uint32_t)s_misc[M_SPEC + p] != f) {
                    for (int u = nspec; u > p; --u) s_misc[M_SPEC + u] = s_misc[M_SPEC + u - 1];
                    s_misc[M_SPEC + p] = (int)f;
                    ++nspec;
                }
            }
        }
        s_misc[M_NSPEC] = nspec;
    }
    __syncthreads();
    uint32_t fs[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) fs[s] = (uint32_t)s_misc[M_F + s];
    const int nspec = s_misc[M_NSPEC];
    int nan_seen = 0;
    const int kk = lane < (int)KP ? lane : (int)KP - 1;  // score row this lane reads
    const double *myrow = s_score + (size_t)kk * ROW;

    for (uint32_t tile = blockIdx.x; tile < P.nt; tile += gridDim.x) {
        const uint32_t doc0 = P.tile_doc_off[tile];
        const int nd = (int)(P.tile_doc_off[tile + 1] - doc0);
        const bool active = t < nd;
        const uint32_t pos = active ? P.pd_pos[doc0 + t] : 0u;
        const uint32_t task0 = F.tile_task_off[tile];
        const int ntask = (int)(F.tile_task_off[tile + 1] - task0);
        if (t < ntask) s_tasks[t] = F.tasks[task0 + t];
        s_gexp[t] = active ? __ldg(P.gexp + pos) : 0.0;
        s_cls[t] = active ? __ldg(F.pd_cls + doc0 + t) : (uint8_t)0;

        // ---- phase 1: one pass over the tile's features for every sweep ----
        double acc[NS], pre[NS], xf[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = pre[s] = xf[s] = 0.0;
        const float *__restrict__ xp = P.x + pos;
        uint32_t j = 0;
        for (int si = 0; si <= nspec; ++si) {
            const uint32_t sp = si < nspec ? (uint32_t)s_misc[M_SPEC + si] : dm;
            for (; j + 8 <= sp; j += 8) {
                float xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) xv[u] = __ldg(xp + (size_t)(j + u) * P.ld);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double xd = (double)xv[u];
                    const double *wj = s_w + (size_t)(j + u) * NS;
#pragma unroll
                    for (int s = 0; s < NS; ++s) acc[s] = __dadd_rn(acc[s], __dmul_rn(xd, wj[s]));
                }
            }
            for (; j < sp; ++j) {
                const double xd = (double)__ldg(xp + (size_t)j * P.ld);
                const double *wj = s_w + (size_t)j * NS;
#pragma unroll
                for (int s = 0; s < NS; ++s) acc[s] = __dadd_rn(acc[s], __dmul_rn(xd, wj[s]));
            }
            if (sp < dm) {  // a coordinate some sweep varies: split that sweep's sum here
                const double xd = (double)__ldg(xp + (size_t)sp * P.ld);
                const double *wj = s_w + (size_t)sp * NS;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (fs[s] == sp) {
                        pre[s] = acc[s];
                        xf[s] = xd;
                        acc[s] = 0.0;
                    } else {
                        acc[s] = __dadd_rn(acc[s], __dmul_rn(xd, wj[s]));
                    }
                }
                j = sp + 1;
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (fs[s] >= dm) {  // the varied coordinate lies beyond the row: candidates are inert
                pre[s] = acc[s];
                acc[s] = 0.0;
                xf[s] = 0.0;
            }
        }

        // ---- phase 2: per sweep, score the candidates, rank, fold the metric ----
        for (int s = 0; s < ns; ++s) {
            const int K = s_misc[M_K + s];
            if (K == 0) continue;
            double ps = pre[0], xs = xf[0], ss = acc[0];
#pragma unroll
            for (int u = 1; u < NS; ++u) {
                if (s == u) {
                    ps = pre[u];
                    xs = xf[u];
                    ss = acc[u];
                }
            }
            {
                const double *cw = s_cand + s * 32;
                double *col = s_score + t;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const double sc = __dadd_rn(__dadd_rn(ps, __dmul_rn(xs, cw[k])), ss);
                    if (sc != sc) nan_seen |= active ? 1 : 0;
                    col[(size_t)k * ROW] = sc;
                }
                uint4 *z = (uint4 *)s_slot;
                const int nz = (int)(((size_t)TB * KP * sizeof(uint16_t)) / sizeof(uint4));
                for (int i = t; i < nz; i += TB) z[i] = make_uint4(0u, 0u, 0u, 0u);
                if (t == 0) s_misc[M_CTR] = 0;
            }
            __syncthreads();
            // -- rank by counting; lane = candidate, TD documents per task --
            {
                int ti = 0;
                if (lane == 0) ti = atomicAdd(&s_misc[M_CTR], 1);
                ti = __shfl_sync(0xffffffffu, ti, 0);
                while (ti < ntask) {
                    int tnext = 0;
                    if (lane == 0) tnext = atomicAdd(&s_misc[M_CTR], 1);
                    const uint2 tk = s_tasks[ti];
                    const int qs = (int)(tk.x & 0xffffu), qe = (int)(tk.x >> 16);
                    const int t0 = (int)(tk.y & 0xffffu), n = (int)(tk.y >> 16);
                    double st[TD];
                    unsigned cnt[TD];
#pragma unroll
                    for (int i = 0; i < TD; ++i) {
                        const int tt = t0 + i < qe ? t0 + i : qe - 1;
                        st[i] = myrow[tt];
                        cnt[i] = 0;
                    }
                    int jq = qs;
#pragma unroll 4
                    for (; jq < t0; ++jq) {  // documents that win ties against the task's
                        const double sj = myrow[jq];
#pragma unroll
                        for (int i = 0; i < TD; ++i) count_ge(cnt[i], sj, st[i]);
                    }
#pragma unroll
                    for (int jj = 0; jj < TD; ++jj) {
                        if (jj < n) {
                            const double sj = myrow[t0 + jj];
#pragma unroll
                            for (int i = 0; i < TD; ++i) {
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
                        for (int i = 0; i < TD; ++i) count_gt(cnt[i], sj, st[i]);
                    }
                    if (lane < K) {
                        const unsigned len = (unsigned)(qe - qs);
                        const unsigned lim =
                            (P.metric == FR_METRIC_NDCG && (unsigned)P.depth < len) ? (unsigned)P.depth : len;
#pragma unroll
                        for (int i = 0; i < TD; ++i)
                            if (i < n && cnt[i] < lim)
                                s_slot[(size_t)(qs + cnt[i]) * KP + lane] = (uint16_t)(t0 + i + 1);
                    }
                    ti = __shfl_sync(0xffffffffu, tnext, 0);
                }
            }
            __syncthreads();
            // -- one warp per query folds the slots in rank order --
            {
                const uint32_t q0 = P.tile_q_off[tile];
                const int nqt = (int)(P.tile_q_off[tile + 1] - q0);
                for (int ql = warp; ql < nqt; ql += TB / 32) {
                    const uint32_t pq = q0 + ql;
                    const uint32_t loc = P.pq_local[pq];
                    const uint32_t start = loc & 0xffffu, len = loc >> 16;
                    const double norm = P.pq_norm[pq];
                    const uint16_t *sl = s_slot + (size_t)start * KP + kk;
                    double value = 0.0;
                    if (P.metric == FR_METRIC_NDCG) {
                        if (norm == norm) {  // Some(ideal), evaluators.rs:351-358
                            const uint32_t lim = len < (uint32_t)P.depth ? len : (uint32_t)P.depth;
                            double dcg = 0.0;
                            for (uint32_t r0 = 0; r0 < lim; r0 += 8) {
                                double term[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    term[u] = 0.0;
                                    const uint32_t r = r0 + u;
                                    if (r < lim) {
                                        const unsigned id = sl[(size_t)r * KP];
                                        if (id) {
                                            if (F.disc_tbl)
                                                term[u] = __ldg(F.disc_tbl + (size_t)s_cls[id - 1] * F.tbl_r + r);
                                            else
                                                term[u] = s_gexp[id - 1] / __ldg(P.lg2 + r);
                                        }
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
                                if (sl[(size_t)r * KP]) {
                                    recall += 1;
                                    sum = __dadd_rn(sum, (double)recall / (double)(r + 1));
                                }
                            }
                            value = sum / norm;
                        }
                    } else {
                        for (uint32_t r = 0; r < len; ++r) {
                            if (sl[(size_t)r * KP]) {
                                value = 1.0 / (double)(r + 1);
                                break;
                            }
                        }
                    }
                    if (lane < K) {
                        if (A.perq)
                            A.perq[((size_t)(s0 + s) * A.cand_stride + A.cand_off + lane) * P.nq_view +
                                   P.pq_view[pq]] = value;
                        const long long fx = __double2ll_rn(value * kFx);
                        atomicAdd(&s_sum[s * 32 + lane], (unsigned long long)fx);
                    }
                }
            }
            __syncthreads();
        }
    }
    if (nan_seen) atomicOr(A.err, ERR_NAN_SCORE);
    __syncthreads();
    for (int idx = t; idx < ns * 32; idx += TB) {
        const int s = idx >> 5, k = idx & 31;
        if (k < s_misc[M_K + s])
            atomicAdd((unsigned long long *)(A.sums + (size_t)(s0 + s) * A.cand_stride + A.cand_off + k),
                      s_sum[idx]);
    }
}

template <int TB, int TD>
int launch_fast(fr_dev_plan *pl, const FastArgs &a, cudaStream_t stream) {
    const SmemLayout L(TB, a.dm, a.kp);
    auto kernel = sweep_fast_kernel<TB, TD>;
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TB, L.total));
    if (occ < 1) return fail("sweep_fast_kernel does not fit on an SM");
    const uint32_t gy = (a.n_sweeps + kMaxSweeps - 1) / kMaxSweeps;
    uint64_t total = (uint64_t)pl->sm_count * (uint64_t)occ;
    uint32_t gx = (uint32_t)std::max<uint64_t>(1, total / gy);
    if (gx > pl->nt) gx = pl->nt;
    PlanView pv = pl->view();
    FastView fv;
    fv.tile_task_off = pl->fast.tile_task_off.p;
    fv.tasks = pl->fast.tasks.p;
    fv.pd_cls = pl->fast.pd_cls.p;
    fv.disc_tbl = pl->fast.n_cls ? pl->fast.disc_tbl.p : nullptr;
    fv.tbl_r = pl->fast.tbl_r;
    auto *ev = pl->ds->prof_slot();
    if (ev) cudaEventRecord(ev->first, stream);
    kernel<<<dim3(gx, gy), TB, L.total, stream>>>(pv, fv, a);
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
    if (pl->tb > 256) {
        fp.why = "a query has more than 256 documents";
        return 0;
    }
    int td = 4;
    if (const char *env = getenv("FASTRANK_TD")) td = atoi(env) == 8 ? 8 : 4;
    fp.td = td;
    const bool ndcg = pl->metric == FR_METRIC_NDCG;
    // gain classes
    std::map<uint32_t, uint32_t> cls_of_bits;
    std::vector<float> cls_gain;
    std::vector<uint8_t> pd_cls(pd_pos.size(), 0);
    bool table = ndcg;
    if (table) {
        for (size_t i = 0; i < pd_pos.size(); ++i) {
            const float g = ds->gain_pos[pd_pos[i]];
            uint32_t bits;
            memcpy(&bits, &g, 4);
            auto it = cls_of_bits.find(bits);
            if (it == cls_of_bits.end()) {
                if (cls_gain.size() >= 255) {
                    table = false;
                    break;
                }
                it = cls_of_bits.emplace(bits, (uint32_t)cls_gain.size()).first;
                cls_gain.push_back(g);
            }
            pd_cls[i] = (uint8_t)it->second;
        }
    }
    cudaStream_t s = ds->stream;
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
        CU(fp.disc_tbl.upload(tbl, s));
    } else {
        std::fill(pd_cls.begin(), pd_cls.end(), 0);
    }
    CU(fp.pd_cls.upload(pd_cls, s));
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
        tile_task_off.push_back((uint32_t)tasks.size());
    }
    fp.n_tasks = (uint32_t)tasks.size();
    CU(fp.tile_task_off.upload(tile_task_off, s));
    CU(fp.tasks.upload(tasks, s));
    fp.ok = true;
    return 0;
}

}  // namespace frbdev

extern "C" int fr_dev_plan_has_fast_sweep(const fr_dev_plan *plan) { return plan && plan->fast.ok ? 1 : 0; }

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
    CU(cudaSetDevice(ds->device));
    cudaStream_t s = ds->stream;
    uint32_t kmax = 0;
    for (size_t r = 0; r < n_sweeps; ++r) {
        if (n_cand[r] > cand_stride) return fail("fr_dev_eval_coord_sweeps_fast: n_cand > cand_stride");
        kmax = std::max(kmax, n_cand[r]);
    }
    const size_t total = n_sweeps * cand_stride;
    CU(pl->w_dev.ensure(n_sweeps * wlen));
    CU(pl->cand_dev.ensure(total));
    CU(pl->fid_dev.ensure(n_sweeps));
    CU(pl->ncand_dev.ensure(n_sweeps));
    CU(pl->sums_dev.ensure(total));
    CU(pl->sums_host.ensure(total));
    if (out_per_query) CU(pl->perq_dev.ensure(total * (size_t)pl->nq_view));
    CU(cudaMemcpyAsync(pl->w_dev.p, base_w, sizeof(double) * n_sweeps * wlen, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(pl->cand_dev.p, cand_w, sizeof(double) * total, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(pl->fid_dev.p, fid, sizeof(uint32_t) * n_sweeps, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(pl->ncand_dev.p, n_cand, sizeof(uint32_t) * n_sweeps, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(pl->sums_dev.p, 0, sizeof(long long) * total, s));
    CU(cudaMemsetAsync(pl->err_dev.p, 0, sizeof(int), s));
    if (out_per_query)
        CU(cudaMemsetAsync(pl->perq_dev.p, 0, sizeof(double) * total * (size_t)pl->nq_view, s));
    for (uint32_t c0 = 0; c0 < kmax; c0 += 32) {
        FastArgs a;
        a.base_w = pl->w_dev.p;
        a.fid = pl->fid_dev.p;
        a.cand_w = pl->cand_dev.p;
        a.n_cand = pl->ncand_dev.p;
        a.sums = pl->sums_dev.p;
        a.perq = out_per_query ? pl->perq_dev.p : nullptr;
        a.n_sweeps = (uint32_t)n_sweeps;
        a.wlen = (uint32_t)wlen;
        a.cand_stride = (uint32_t)cand_stride;
        a.cand_off = c0;
        a.kp = std::min<uint32_t>(32, kmax - c0);
        a.dm = (uint32_t)std::min<size_t>(wlen, ds->d);
        a.err = pl->err_dev.p;
        if (pl->nt == 0) continue;
        int rc;
        if (pl->tb == 128)
            rc = pl->fast.td == 8 ? launch_fast<128, 8>(pl, a, s) : launch_fast<128, 4>(pl, a, s);
        else
            rc = pl->fast.td == 8 ? launch_fast<256, 8>(pl, a, s) : launch_fast<256, 4>(pl, a, s);
        if (rc) return 1;
    }
    if (allreduce_sums(pl, pl->sums_dev.p, total, s)) return 1;
    CU(cudaMemcpyAsync(pl->sums_host.p, pl->sums_dev.p, sizeof(long long) * total,
                       cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(pl->err_host.p, pl->err_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (out_per_query)
        CU(cudaMemcpyAsync(out_per_query, pl->perq_dev.p, sizeof(double) * total * (size_t)pl->nq_view,
                           cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (check_err_flags(pl->err_host.p[0])) return 1;
    for (size_t i = 0; i < total; ++i) out_sum_fx[i] = pl->sums_host.p[i];
    return 0;
}
