// long_queries.cu -- queries that are not tiled: longer than the largest tile (1024 documents
// for the exact-order kernels; 512 when the plan keeps the rest of the dataset on the batched
// sweep, see fr_dev_plan_create).
//
// The tile kernels rank a whole query inside one CTA's shared memory; a query that does not fit
// (MSLR-WEB30K has lists of ~1.2k documents, the reference sorts lists of any length,
// evaluators.rs:206-221) takes this path instead: scores go to a scratch array in HBM (exact
// left-to-right dot products, 8 candidates per read of a feature value), one CTA per
// (query, candidate) stages the list's scores in shared memory and ranks its contributing
// documents by counting, metric terms are scattered to their rank and folded in rank order by
// one thread -- the same arithmetic, in the same order, as rank_and_metric in device.cu, so
// results are bit-identical to the oracle here too.  O(len^2) compares per (list, candidate).
#include "device_common.cuh"

namespace {

constexpr double kFxLong = 1099511627776.0;
static_assert(FR_FX_BITS == 40, "kFxLong must match FR_FX_BITS");
constexpr int kLongChunk = 512;  // most candidates per pass; fewer when the scratch arrays would pass 1 GiB
constexpr int kLongSmem = 4096;  // list length whose scores are staged in shared memory
constexpr int kSelDepth = 64;    // NDCG cut-offs up to this take the pivot selection

struct LongView {
    const uint32_t *lq_off;   // n_long + 1, into ld_pos
    const uint32_t *ld_pos;   // position of every document of the long queries
    const uint32_t *lq_view;  // view (output) index of the query
    const double *lq_norm;    // ideal DCG (NaN = none) or num_relevant
    uint32_t n_docs;
};

// dense_dataset.rs:67-76 for documents of long queries: scores[k][d], left-to-right f64 dot.
// One thread scores one document for kLongKB candidates, so a feature value is read once per
// kLongKB candidates; the loads of 8 features are issued before any of them is consumed.
constexpr int kLongKB = 8;
__global__ void long_linear_scores_kernel(const float *__restrict__ x, size_t ld, uint32_t dm,
                                          LongView L, const double *__restrict__ w, uint32_t wlen,
                                          uint32_t nc, double *__restrict__ scores) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= L.n_docs) return;
    const uint32_t k0 = blockIdx.y * kLongKB;
    const float *__restrict__ xp = x + L.ld_pos[d];
    const double *wk[kLongKB];
    double acc[kLongKB];
#pragma unroll
    for (int k = 0; k < kLongKB; ++k) {
        wk[k] = w + (size_t)min(k0 + k, nc - 1) * wlen;  // rows past nc repeat the last one, unused
        acc[k] = 0.0;
    }
    for (uint32_t j0 = 0; j0 < dm; j0 += 8) {
        float xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = j0 + u < dm ? __ldg(xp + (size_t)(j0 + u) * ld) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u < dm) {
                const double xd = (double)xv[u];
#pragma unroll
                for (int k = 0; k < kLongKB; ++k)
                    acc[k] = __dadd_rn(acc[k], __dmul_rn(xd, __ldg(wk[k] + j0 + u)));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kLongKB; ++k)
        if (k0 + k < nc) scores[(size_t)(k0 + k) * L.n_docs + d] = acc[k];
}

// The batched sweep's scores for documents of untiled lists, from the tables the tile kernel
// uses: T_s = sum_j x_j * w_sj (j ascending, separate multiply and add, the swept coordinate
// zeroed by the host), score(row) = T_s + x_f * row_w -- the arithmetic of sweep_fast_kernel's
// phase 1 and score_group, so tiled and untiled lists of one call follow the same contract.
__global__ void long_sweep_scores_kernel(const float *__restrict__ x, size_t ld, uint32_t dm, uint32_t dm8,
                                         LongView L, const double *__restrict__ base_wt,
                                         const uint32_t *__restrict__ fid, uint32_t n_sweeps,
                                         const double *__restrict__ row_w, const uint32_t *__restrict__ row_meta,
                                         const uint32_t *__restrict__ grp_row_off, uint32_t r_lo, uint32_t r_hi,
                                         double *__restrict__ scores) {
    constexpr int NS = kMaxSweeps;
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = blockIdx.y;
    const uint32_t rb = max(grp_row_off[g], r_lo), re = min(grp_row_off[g + 1], r_hi);
    if (d >= L.n_docs || rb >= re) return;
    const float *__restrict__ xp = x + L.ld_pos[d];
    const double *__restrict__ wt = base_wt + (size_t)g * dm8 * NS;
    double acc[NS];
    float xf[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        acc[s] = 0.0;
        const uint32_t sw = g * NS + s;
        const uint32_t f = sw < n_sweeps ? fid[sw] : 0xffffffffu;
        xf[s] = f < dm ? __ldg(xp + (size_t)f * ld) : 0.f;
    }
    for (uint32_t j0 = 0; j0 < dm; j0 += 8) {
        float xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = j0 + u < dm ? __ldg(xp + (size_t)(j0 + u) * ld) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // the table is zero-padded to dm8 coordinates
            const double xd = (double)xv[u];
            const double2 *wj = reinterpret_cast<const double2 *>(wt + (size_t)(j0 + u) * NS);
#pragma unroll
            for (int s = 0; s < NS; s += 2) {
                const double2 w2 = __ldg(wj + s / 2);
                acc[s] = __dadd_rn(acc[s], __dmul_rn(xd, w2.x));
                acc[s + 1] = __dadd_rn(acc[s + 1], __dmul_rn(xd, w2.y));
            }
        }
    }
    for (uint32_t r = rb; r < re; ++r) {
        const uint32_t s = row_meta[r];
        double ss = acc[0];
        float xs = xf[0];
#pragma unroll
        for (int u = 1; u < NS; ++u) {
            if (s == (uint32_t)u) {
                ss = acc[u];
                xs = xf[u];
            }
        }
        scores[(size_t)(r - r_lo) * L.n_docs + d] = __dadd_rn(ss, __dmul_rn((double)xs, __ldg(row_w + r)));
    }
}

__global__ void long_gather_scores_kernel(LongView L, const double *__restrict__ scores_pos,
                                          double *__restrict__ scores) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < L.n_docs) scores[d] = scores_pos[L.ld_pos[d]];
}

__global__ void __launch_bounds__(256) long_rank_kernel(PlanView P, LongView L, const double *__restrict__ scores,
                                                        double *__restrict__ slots,
                                                        const uint32_t *__restrict__ out_idx, long long *sums,
                                                        double *perq, int *err) {
    const uint32_t q = blockIdx.x, k = blockIdx.y;
    const uint32_t base = L.lq_off[q], len = L.lq_off[q + 1] - base;
    const double *__restrict__ sc = scores + (size_t)k * L.n_docs + base;
    double *sl = slots + (size_t)k * L.n_docs + base;
    // the list's scores are staged in shared memory when they fit, else read through L1
    __shared__ double s_sc[kLongSmem];
    const bool staged = len <= (uint32_t)kLongSmem;
    for (uint32_t t = threadIdx.x; t < len; t += blockDim.x) {
        const double st = sc[t];
        if (st != st) atomicOr(err, ERR_NAN_SCORE);
        if (staged) s_sc[t] = st;
        sl[t] = 0.0;
    }
    __syncthreads();
    const double *__restrict__ src = staged ? s_sc : sc;
    // NDCG@k on a list much longer than k: only the k best documents matter.  A pivot taken from
    // a 32-document sample splits the list; if c >= k documents score above it, every document
    // at or below it has rank >= c >= k and adds nothing, and the documents that outrank a
    // survivor are survivors themselves -- so ranking the c survivors among themselves gives
    // their exact ranks (c is a few dozen, not the list length).
    bool ranked = false;
    if (staged && P.metric == FR_METRIC_NDCG && (unsigned)P.depth <= (unsigned)kSelDepth &&
        len >= 8u * (unsigned)P.depth && len >= 64u) {
        __shared__ double s_smp[32], s_piv[32];
        __shared__ unsigned s_above[32], s_ncand;
        __shared__ uint16_t s_cand[kLongSmem];
        const unsigned depth = (unsigned)P.depth;
        const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        if (warp == 0) {  // the sample, sorted descending by counting
            const double v = s_sc[(size_t)lane * len / 32];
            s_smp[lane] = v;
            s_piv[lane] = v;  // (defined even if NaNs make ranks collide; that run reports an error)
            s_above[lane] = 0;
            __syncwarp();
            unsigned r = 0;
            for (unsigned i = 0; i < 32; ++i) {
                const double o = s_smp[i];
                r += (o > v) || (o == v && i < lane);
            }
            __syncwarp();
            s_piv[r] = v;
        }
        __syncthreads();
        double pivot = 0.0;
        bool have = false;
        for (unsigned i = 0; i < 32 && !have; ++i) {  // pivots from the top until k documents lie above
            const double p = s_piv[i];
            unsigned local = 0;
            for (uint32_t t = threadIdx.x; t < len; t += blockDim.x) local += s_sc[t] > p ? 1u : 0u;
            local = __reduce_add_sync(0xffffffffu, local);
            if (lane == 0 && local) atomicAdd(&s_above[i], local);
            __syncthreads();
            if (s_above[i] >= depth) {
                pivot = p;
                have = true;
            }
        }
        if (have) {
            if (warp == 0) {  // survivors in list order (the tie-break needs it)
                unsigned n = 0;
                for (uint32_t t0 = 0; t0 < len; t0 += 32) {
                    const uint32_t t = t0 + lane;
                    const bool in = t < len && s_sc[t] > pivot;
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in) s_cand[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)t;
                    n += __popc(m);
                }
                if (lane == 0) s_ncand = n;
            }
            __syncthreads();
            const unsigned nc = s_ncand;
            for (unsigned u = threadIdx.x; u < nc; u += blockDim.x) {
                const unsigned d = s_cand[u];
                const double ge = P.gexp[L.ld_pos[base + d]];
                if (ge == 0.0) continue;
                const double st = s_sc[d];
                unsigned cnt = 0;
                for (unsigned v = 0; v < nc; ++v) count_outranks(cnt, s_sc[s_cand[v]], st, v < u ? 1u : 0u);
                if (cnt < depth) sl[cnt] = ge / P.lg2[cnt];  // evaluators.rs:265-270
            }
            ranked = true;
        }
    }
    // otherwise: every document that can contribute is ranked against the whole list (local order
    // is gain-ascending, so they are the tail of the list and whole warps agree)
    for (uint32_t t = threadIdx.x; t < len && !ranked; t += blockDim.x) {
        const uint32_t pos = L.ld_pos[base + t];
        double ge = 0.0;
        bool contrib;
        if (P.metric == FR_METRIC_NDCG) {
            ge = P.gexp[pos];
            contrib = ge != 0.0;
        } else {
            contrib = P.gain[pos] > 0.0f;
        }
        if (!contrib) continue;
        const double st = src[t];
        unsigned cnt = 0;
        // a document that `depth` others already outrank adds nothing to DCG@depth: stop counting
        // (most documents of a long list leave after a few dozen compares)
        const unsigned enough = P.metric == FR_METRIC_NDCG ? (unsigned)min((unsigned)P.depth, len) : len;
        for (uint32_t j0 = 0; j0 < len && cnt < enough; j0 += 32) {
            const uint32_t j1 = min(j0 + 32u, len);
#pragma unroll 4
            for (uint32_t j = j0; j < j1; ++j) count_outranks(cnt, src[j], st, j < t ? 1u : 0u);
        }
        if (st != st) continue;  // NaN: reported above, must not take another document's slot
        if (P.metric == FR_METRIC_NDCG) {
            if ((int)cnt < P.depth) sl[cnt] = ge / P.lg2[cnt];  // evaluators.rs:265-270
        } else {
            sl[cnt] = 1.0;
        }
    }
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double norm = L.lq_norm[q];
    double value = 0.0;
    if (P.metric == FR_METRIC_NDCG) {
        if (norm == norm) {
            const uint32_t lim = len < (uint32_t)P.depth ? len : (uint32_t)P.depth;
            double dcg = 0.0;
            for (uint32_t r = 0; r < lim; ++r) dcg = __dadd_rn(dcg, sl[r]);
            if (dcg > norm) atomicOr(err, ERR_DCG_ABOVE_IDEAL);
            value = dcg / norm;
        }
    } else if (P.metric == FR_METRIC_AP) {
        if (norm > 0.0) {
            unsigned recall = 0;
            double sum = 0.0;
            for (uint32_t r = 0; r < len; ++r) {
                if (sl[r] != 0.0) {
                    recall += 1;
                    sum = __dadd_rn(sum, (double)recall / (double)(r + 1));
                }
            }
            value = sum / norm;
        }
    } else {
        for (uint32_t r = 0; r < len; ++r) {
            if (sl[r] != 0.0) {
                value = 1.0 / (double)(r + 1);
                break;
            }
        }
    }
    const uint32_t o = out_idx[k];
    if (perq) perq[(size_t)o * P.nq_view + L.lq_view[q]] = value;
    atomicAdd((unsigned long long *)(sums + o), (unsigned long long)__double2ll_rn(value * kFxLong));
}

LongView long_view(const fr_dev_plan *pl) {
    LongView v;
    v.lq_off = pl->lng.lq_off.p;
    v.ld_pos = pl->lng.ld_pos.p;
    v.lq_view = pl->lng.lq_view.p;
    v.lq_norm = pl->lng.lq_norm.p;
    v.n_docs = pl->lng.n_docs;
    return v;
}

int rank_chunk(fr_dev_plan *pl, uint32_t n_chunk, const uint32_t *out_idx_dev, long long *sums_dev,
               double *perq_dev, int *err_dev, cudaStream_t s) {
    LongPlan &lp = pl->lng;
    long_rank_kernel<<<dim3(lp.n_long, n_chunk), 256, 0, s>>>(pl->view(), long_view(pl), lp.scores.p, lp.slots.p,
                                                             out_idx_dev, sums_dev, perq_dev, err_dev);
    LAUNCHED();
    CU(cudaGetLastError());
    return 0;
}

}  // namespace

namespace frbdev {

// Host half: the lists that did not fit a tile, their norms (plan_norms_kernel's arithmetic
// restated with the host's libm-built tables) and the scratch arrays.
int build_long_plan(fr_dev_plan *pl, const std::vector<std::vector<uint32_t>> &qpos,
                    const std::vector<uint32_t> &long_views, const fr_dev_plan_desc *desc) {
    LongPlan &lp = pl->lng;
    fr_dev_dataset *ds = pl->ds;
    lp.n_long = (uint32_t)long_views.size();
    lp.n_docs = 0;
    if (lp.n_long == 0) return 0;
    std::vector<uint32_t> lq_off{0}, ld_pos, lq_view;
    std::vector<double> lq_norm;
    for (uint32_t v : long_views) {
        const std::vector<uint32_t> &pos = qpos[v];
        const uint32_t len = (uint32_t)pos.size();
        const bool has_ov = desc->norm_present && desc->norm_value && desc->norm_present[v] != 0;
        double norm = 0.0;
        if (pl->metric == FR_METRIC_NDCG) {
            if (has_ov) {
                norm = desc->norm_value[v];
            } else if (!(ds->gain_pos[pos[len - 1]] > 0.0f)) {  // positions are gain-ascending
                norm = std::nan("");
            } else {
                const uint32_t lim = len < (uint32_t)pl->depth ? len : (uint32_t)pl->depth;
                double dcg = 0.0;
                for (uint32_t i = 0; i < lim; ++i) {
                    const double ge = std::pow(2.0, (double)ds->gain_pos[pos[len - 1 - i]]) - 1.0;
                    dcg += ge / std::log2((double)i + 2.0);
                }
                norm = dcg;
            }
        } else if (pl->metric == FR_METRIC_AP) {
            uint32_t rel = 0;
            for (uint32_t p : pos) rel += ds->gain_pos[p] > 0.0f;
            norm = (has_ov && desc->norm_value[v] > 0.0) ? desc->norm_value[v] : (double)rel;
        }
        ld_pos.insert(ld_pos.end(), pos.begin(), pos.end());
        lq_off.push_back((uint32_t)ld_pos.size());
        lq_view.push_back(v);
        lq_norm.push_back(norm);
    }
    lp.n_docs = (uint32_t)ld_pos.size();
    cudaStream_t s = ds->stream;
    CU(lp.lq_off.upload(lq_off, s));
    CU(lp.ld_pos.upload(ld_pos, s));
    CU(lp.lq_view.upload(lq_view, s));
    CU(lp.lq_norm.upload(lq_norm, s));
    // two f64 scratch arrays of chunk x n_docs: 512 candidates per pass unless that passes 1 GiB
    lp.chunk = (uint32_t)std::max<size_t>(4, std::min<size_t>(kLongChunk, ((size_t)1 << 30) / (16 * (size_t)lp.n_docs)));
    if (const char *env = getenv("FASTRANK_LONG_CHUNK")) {  // test knob: candidates per pass
        const int v = atoi(env);
        if (v >= 1 && v <= kLongChunk) lp.chunk = (uint32_t)v;
    }
    CU(lp.scores.alloc((size_t)lp.chunk * lp.n_docs));
    CU(lp.slots.alloc((size_t)lp.chunk * lp.n_docs));
    CU(lp.out_idx.alloc(lp.chunk));
    CU(lp.w.alloc(1));
    CU(cudaStreamSynchronize(s));
    return 0;
}

// n_vec full weight vectors (host, row-major n_vec x wlen); vector i adds into sums[out_index[i]]
int eval_long_linear(fr_dev_plan *pl, const double *w_host, size_t wlen, size_t n_vec,
                     const uint32_t *out_index, long long *sums_dev, double *perq_dev, int *err_dev,
                     cudaStream_t s) {
    LongPlan &lp = pl->lng;
    if (lp.n_long == 0 || n_vec == 0) return 0;
    fr_dev_dataset *ds = pl->ds;
    const uint32_t dm = (uint32_t)std::min<size_t>(wlen, ds->d);
    CU(lp.w.ensure((size_t)lp.chunk * std::max<size_t>(wlen, 1)));
    for (size_t c0 = 0; c0 < n_vec; c0 += lp.chunk) {
        const uint32_t nc = (uint32_t)std::min<size_t>(lp.chunk, n_vec - c0);
        CU(cudaMemcpyAsync(lp.w.p, w_host + c0 * wlen, sizeof(double) * nc * wlen, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(lp.out_idx.p, out_index + c0, sizeof(uint32_t) * nc, cudaMemcpyHostToDevice, s));
        long_linear_scores_kernel<<<dim3((lp.n_docs + 127) / 128, (nc + kLongKB - 1) / kLongKB), 128, 0, s>>>(
            ds->x.p, ds->ld, dm, long_view(pl), lp.w.p, (uint32_t)wlen, nc, lp.scores.p);
        LAUNCHED();
        CU(cudaGetLastError());
        if (rank_chunk(pl, nc, lp.out_idx.p, sums_dev, perq_dev, err_dev, s)) return 1;
    }
    return 0;
}

int eval_long_sweep(fr_dev_plan *pl, const double *base_wt, const uint32_t *fid, uint32_t n_sweeps,
                    const double *row_w, const uint32_t *row_meta, const uint32_t *row_out,
                    const uint32_t *grp_row_off, uint32_t n_groups, uint32_t n_rows, uint32_t dm, uint32_t dm8,
                    long long *sums_dev, double *perq_dev, int *err_dev, cudaStream_t s) {
    LongPlan &lp = pl->lng;
    if (lp.n_long == 0 || n_rows == 0) return 0;
    fr_dev_dataset *ds = pl->ds;
    for (uint32_t r0 = 0; r0 < n_rows; r0 += lp.chunk) {
        const uint32_t r1 = std::min(n_rows, r0 + lp.chunk);
        long_sweep_scores_kernel<<<dim3((lp.n_docs + 127) / 128, n_groups), 128, 0, s>>>(
            ds->x.p, ds->ld, dm, dm8, long_view(pl), base_wt, fid, n_sweeps, row_w, row_meta, grp_row_off, r0, r1,
            lp.scores.p);
        LAUNCHED();
        CU(cudaGetLastError());
        if (rank_chunk(pl, r1 - r0, row_out + r0, sums_dev, perq_dev, err_dev, s)) return 1;
    }
    return 0;
}

// scores already in HBM by position (trees, ensembles); adds into sums[0]
int eval_long_scores(fr_dev_plan *pl, const double *scores_pos, long long *sums_dev, double *perq_dev,
                     int *err_dev, cudaStream_t s) {
    LongPlan &lp = pl->lng;
    if (lp.n_long == 0) return 0;
    const uint32_t zero = 0;
    CU(cudaMemcpyAsync(lp.out_idx.p, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    long_gather_scores_kernel<<<(lp.n_docs + 127) / 128, 128, 0, s>>>(long_view(pl), scores_pos, lp.scores.p);
    LAUNCHED();
    CU(cudaGetLastError());
    return rank_chunk(pl, 1, lp.out_idx.p, sums_dev, perq_dev, err_dev, s);
}

}  // namespace frbdev
