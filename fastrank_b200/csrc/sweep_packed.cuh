// sweep_packed.cuh -- the batched coordinate-ascent sweep for NDCG@k, k <= 16 (included by
// sweep_fast.cu, which documents the work being replaced and the arithmetic contract; both
// kernels produce the same bits).
//
// What differs from sweep_fast_kernel (the general kernel: any metric, any cut-off):
//   * no score matrix and no CTA-wide phases.  Phase 1 (thread = document, one pass over the
//     tile's slice of X for all sweeps) leaves  T2[s][t] = { T_s(t), x_{f_s}(t) }  in shared memory;
//     after ONE barrier the tile's work is a queue of independent WARP items
//     (query, group of 32 candidate rows): lane = candidate row, scores are formed where they are
//     compared,  score = T + x_f * c_lane  (one LDS.128, one DMUL, one DADD per walked document).
//   * ranking by counting as before (DSETP + predicated add per comparison; the tie rule is the
//     tile order), W = 16 / 8 / 4 documents per walk, pairs inside a chunk compared once.
//   * a candidate's top-k lives in ONE 64-bit register per lane: 4 bits per rank hold the gain
//     class of the document ranked there (<= 15 classes), so there is no slot buffer to clear,
//     scatter into or synchronise on; the same warp folds it at once -- the reference's
//     left-to-right f64 sum of (2^gain - 1) / log2(rank + 2) over ranks 0..k-1
//     (evaluators.rs:255-272) from the host-built table -- and adds round(value * 2^40) to the
//     row's sum.
//   * lists of >= 48 documents are pruned per candidate before they are ranked (rank_pruned below).
// Two barriers per tile instead of two per row group; 0.86 G warp instructions per 408-candidate
// launch on 1M x 136 against 1.21 G, 0.98 ms against 1.66 ms (profiles/r02_sweep_packed_*).
#pragma once

// unroll factors of the two walk loops per chunk width (code size vs. latency hiding: at 4/4/4 the
// kernel is ~75 KB of SASS and stalls on instruction fetch; measured 1.05 ms per bench step at
// 4/4/4, 0.98 at 2/2/2, 0.97 at 2/4/4, 1.08 at 1/1/1)
#ifndef PK_U16
#define PK_U16 2
#endif
#ifndef PK_U8
#define PK_U8 4
#endif
#ifndef PK_U4
#define PK_U4 4
#endif
#ifndef PK_NO_W4
#define PK_NO_W4 0
#endif

struct PackedView {
    const uint32_t *q_task_off;     // [nq_plan + 1] first task of a plan query
    const uint16_t *q_order;        // [nq_plan] tile-local query indices, costliest first
    const uint32_t *tile_task_off;  // [nt + 1]
    const uint4 *tasks;             // .x = t0 | n << 16 (tile-local documents of one query, n <= 16),
                                    // .y/.z = their tags (gain class + 1), 4 bits each, 0 beyond n
    const uint8_t *pd_cls;          // plan doc -> gain class (select-then-rank path)
    const double *tbl;              // [n_cls + 1][tbl_r]: row 0 zeros (empty slot), row c + 1 = class c
    uint32_t tbl_r, n_cls;
    const double *ap_tbl;           // AP: [recall][rank] = (double)recall / (double)(rank + 1), or nullptr
    uint32_t ap_cols;               // ranks per row of ap_tbl
};

constexpr int kPruneSlots = 64;  // survivors a candidate may keep in the select-then-rank path

struct PackedLayout {
    size_t t2, sum, roww, tbl, qd, tasks, order, rowsw, misc, cls, surv, slots, w, total;
    // ps: survivor slots per candidate (0: the kernel instance has no select-then-rank path)
    // slot_mode: ranks are filed in a per-warp shared-memory slot buffer instead of a register
    __host__ __device__ PackedLayout(int tb, uint32_t w_doubles, int ps, bool slot_mode = false) {
        size_t o = 0;
        t2 = o;     o += sizeof(double2) * kMaxSweeps * (size_t)(tb + (ps > 0 ? 1 : 0));  // + one "never ranks" slot per row
        sum = o;    o += sizeof(unsigned long long) * kMaxRows;
        roww = o;   o += sizeof(double) * kMaxRows;
        tbl = o;    o += sizeof(double) * 16 * 16;
        qd = o;     o += sizeof(uint2) * tb;
        tasks = o;  o += sizeof(uint4) * tb;
        order = o;  o += align16(sizeof(uint16_t) * tb);
        rowsw = o;  o += kMaxRows;
        misc = o;   o += 256;
        cls = o;    o += align16((size_t)tb + 1);
        surv = o;   o += (tb < 256 ? 1 : 2) * (size_t)(ps ? ps + 1 : 0) * 32 * (size_t)(tb / 32);  // [warp][slot][lane], + a scratch slot
        slots = o;  o += slot_mode ? (size_t)tb * 32 * (size_t)(tb / 32) : 0;  // [warp][rank][lane] u8
        w = o;      o += sizeof(double) * w_doubles;
        total = o;
    }
};

// documents i < j of one chunk (j later in tile order, so j loses ties): exactly one outranks the other
__device__ __forceinline__ void count_pair(unsigned &ci, unsigned &cj, double si, double sj) {
    asm("{ .reg .pred p; setp.gt.f64 p, %2, %3; @p add.u32 %0, %0, 1; @!p add.u32 %1, %1, 1; }"
        : "+r"(ci), "+r"(cj)
        : "d"(sj), "d"(si));
}

// tag << (4 * rank), 0 when the rank is beyond the register (shl.b64 clamps the shift amount)
__device__ __forceinline__ unsigned long long tag_at_rank(unsigned tag, unsigned rank) {
    unsigned long long v;
    asm("{ .reg .b64 t; cvt.u64.u32 t, %1; shl.b64 %0, t, %2; }" : "=l"(v) : "r"(tag), "r"(rank << 2));
    return v;
}

// Ranks the W documents [t0, t0 + n) of the query occupying tile-local [qs, qe) under this lane's
// candidate c and files their tags by rank into `packed`: 4 bits per rank, ranks >= 16 fall off
// the register and ranks in [lim, 16) are never read.  `tags` holds the chunk's tags, 0 beyond n.
// SLOTS: the tag (gain class + 1 for NDCG, 1 for AP / RR) goes to slot[rank][lane] of the warp's
// shared-memory buffer instead, for ranks below lim -- any cut-off, any number of classes.
template <int W, bool SLOTS>
__device__ __forceinline__ void rank_chunk(const double2 *__restrict__ trow, double c, int qs, int qe, int t0, int n,
                                           unsigned long long tags, unsigned long long &packed, unsigned lim,
                                           const uint8_t *__restrict__ s_cls, uint8_t *__restrict__ slot, bool class_tags) {
    double st[W];
    unsigned cnt[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const int tt = t0 + i < qe ? t0 + i : qe - 1;
        const double2 tx = trow[tt];
        st[i] = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
        cnt[i] = 0;
    }
    int jq = qs;
    constexpr int UNR = W == 16 ? PK_U16 : (W == 8 ? PK_U8 : PK_U4);
#pragma unroll UNR
    for (; jq < t0; ++jq) {  // documents that win ties against the chunk's
        const double2 tx = trow[jq];
        const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
#pragma unroll
        for (int i = 0; i < W; ++i) count_ge(cnt[i], sj, st[i]);
    }
#pragma unroll
    for (int j = 1; j < W; ++j) {
        if (j < n) {
#pragma unroll
            for (int i = 0; i < j; ++i) count_pair(cnt[i], cnt[j], st[i], st[j]);
        }
    }
    jq = t0 + n;
#pragma unroll UNR
    for (; jq < qe; ++jq) {  // documents that lose ties
        const double2 tx = trow[jq];
        const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
#pragma unroll
        for (int i = 0; i < W; ++i) count_gt(cnt[i], sj, st[i]);
    }
    if (SLOTS) {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (i < n && cnt[i] < lim) slot[cnt[i] * 32] = class_tags ? (uint8_t)(s_cls[t0 + i] + 1) : (uint8_t)1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) packed |= tag_at_rank((unsigned)(tags >> (4 * i)) & 15u, cnt[i]);
    }
}

// Select-then-rank for long lists (NDCG@k, k << len).  Per candidate (lane):
//   A. the list is cut into `lim` runs; thr = the smallest of the runs' maxima.  `lim` documents
//      score >= thr, so a document below thr is outranked by at least `lim` others: it cannot be
//      in the top `lim`, and it outranks nobody who is.
//   B. the documents with score >= thr -- the survivors, in list order, which is the tie-break
//      order -- are written to a lane-private list of PS indices; the rest of the list is padded
//      with the tile's dummy slot (score -inf under every candidate, tag 0).
//   B2. when some lane kept many, the same cut is applied to the survivor lists themselves (runs
//      of list slots; a lane whose list is short gets thr = -inf and keeps everything).
//   C. survivors are ranked among themselves by counting, 8 list slots at a time; a survivor's
//      rank among survivors IS its rank in the whole list (everything that outranks a survivor is a
//      survivor), so what lands in ranks < lim is exactly what the full count would put there.
// Returns false (warp-uniform) when some lane has more survivors than slots; the caller then
// takes the full count.
template <int TB>
struct SurvIndex {
    typedef uint16_t type;
};
template <>
struct SurvIndex<128> {
    typedef uint8_t type;  // tile-local indices 0..128 fit a byte
};

template <int TB, int PS>
__device__ __forceinline__ bool rank_pruned(const double2 *__restrict__ trow, double c, int qs, int qe, unsigned lim,
                                            const uint8_t *__restrict__ s_cls,
                                            typename SurvIndex<TB>::type *__restrict__ surv, int lane,
                                            unsigned long long &packed) {
    typedef typename SurvIndex<TB>::type surv_t;
    const double kNegInf = __longlong_as_double(0xfff0000000000000ll);
    const int len = qe - qs;
    const int csize = len / (int)lim;
    // A: two runs at a time (independent maxima: the chains of compares overlap)
    double thr = __longlong_as_double(0x7ff0000000000000ll);
    {
        unsigned ch = 0;
        for (; ch + 2 <= lim; ch += 2) {
            const int ja = qs + (int)ch * csize, jb = ja + csize;
            const int nb = ch + 2 == lim ? qe - jb : csize;  // the last run takes the remainder
            double ma = kNegInf, mb = kNegInf;
#pragma unroll 4
            for (int u = 0; u < csize; ++u) {
                const double2 ta = trow[ja + u], tb = trow[jb + u];
                const double sa = __dadd_rn(ta.x, __dmul_rn(ta.y, c)), sb = __dadd_rn(tb.x, __dmul_rn(tb.y, c));
                ma = sa > ma ? sa : ma;
                mb = sb > mb ? sb : mb;
            }
            for (int u = csize; u < nb; ++u) {
                const double2 tb = trow[jb + u];
                const double sb = __dadd_rn(tb.x, __dmul_rn(tb.y, c));
                mb = sb > mb ? sb : mb;
            }
            const double m2 = ma < mb ? ma : mb;
            thr = m2 < thr ? m2 : thr;
        }
        if (ch < lim) {  // odd lim: the last run alone, with the remainder
            double mx = kNegInf;
            for (int j = qs + (int)ch * csize; j < qe; ++j) {
                const double2 tx = trow[j];
                const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
                mx = sj > mx ? sj : mx;
            }
            thr = mx < thr ? mx : thr;
        }
    }
    // B: branch-free append (a document below thr writes the lane's scratch slot PS)
    surv_t *mine = surv + lane;  // [slot][lane], PS + 1 slots
    unsigned n_s = 0;
#pragma unroll 4
    for (int j = qs; j < qe; ++j) {
        const double2 tx = trow[j];
        const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
        const bool keep = sj >= thr;
        const unsigned slot = keep ? (n_s < (unsigned)PS ? n_s : (unsigned)PS) : (unsigned)PS;
        mine[slot * 32] = (surv_t)j;
        n_s += keep ? 1u : 0u;
    }
    if (__any_sync(0xffffffffu, n_s > (unsigned)PS)) return false;
    int n_max = (int)__reduce_max_sync(0xffffffffu, n_s);
    for (int k = (int)n_s; k < ((n_max + 7) & ~7) && k < PS; ++k) mine[k * 32] = (surv_t)TB;  // padding: never ranks
    // B2
    if (n_max > 24 && n_max >= 2 * (int)lim) {
        const int cs = (n_max + (int)lim - 1) / (int)lim;
        double thr2 = __longlong_as_double(0x7ff0000000000000ll);
        int k = 0;
        for (unsigned ch = 0; ch < lim; ++ch) {
            double mx = kNegInf;
            for (int e = min(n_max, k + cs); k < e; ++k) {
                const double2 tx = trow[mine[k * 32]];
                const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
                mx = sj > mx ? sj : mx;
            }
            thr2 = mx < thr2 ? mx : thr2;
        }
        unsigned w = 0;
#pragma unroll 2
        for (k = 0; k < n_max; ++k) {
            const unsigned id = mine[k * 32];
            const double2 tx = trow[id];
            const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
            const bool keep = id != (unsigned)TB && sj >= thr2;
            mine[(keep ? w : (unsigned)PS) * 32] = (surv_t)id;  // w <= k: in place; the rest goes to the scratch slot
            w += keep ? 1u : 0u;
        }
        const int n_new = (int)__reduce_max_sync(0xffffffffu, w);
        for (k = (int)w; k < ((n_new + 7) & ~7) && k < PS; ++k) mine[k * 32] = (surv_t)TB;
        n_max = n_new;
    }
    // C
    for (int k1 = 0; k1 < n_max; k1 += 8) {
        double st[8];
        unsigned cnt[8], id[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            id[i] = k1 + i < PS ? mine[(k1 + i) * 32] : (unsigned)TB;
            const double2 tx = trow[id[i]];
            st[i] = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
            cnt[i] = 0;
        }
#pragma unroll 4
        for (int k2 = 0; k2 < k1; ++k2) {  // earlier in list order: they win ties
            const double2 tx = trow[mine[k2 * 32]];
            const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
#pragma unroll
            for (int i = 0; i < 8; ++i) count_ge(cnt[i], sj, st[i]);
        }
#pragma unroll
        for (int j = 1; j < 8; ++j) {
#pragma unroll
            for (int i = 0; i < j; ++i) count_pair(cnt[i], cnt[j], st[i], st[j]);
        }
#pragma unroll 4
        for (int k2 = k1 + 8; k2 < n_max; ++k2) {  // later in list order: they lose ties
            const double2 tx = trow[mine[k2 * 32]];
            const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
#pragma unroll
            for (int i = 0; i < 8; ++i) count_gt(cnt[i], sj, st[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) packed |= tag_at_rank(((unsigned)s_cls[id[i]] + 1u) & 15u, cnt[i]);
    }
    return true;
}

template <int TB, bool WS, int MINB, int PS, bool SLOTS>
__global__ void __launch_bounds__(TB, MINB)
sweep_packed_kernel(PlanView P, PackedView V, FastArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NS = kMaxSweeps;
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t s0 = blockIdx.y * NS;
    const int ns = (int)min((uint32_t)NS, A.n_sweeps - s0);
    const uint32_t dm = A.dm;
    const uint32_t row0 = A.grp_row_off[blockIdx.y];
    const int R = (int)(A.grp_row_off[blockIdx.y + 1] - row0);  // rows of this sweep group
    const int G = (R + 31) >> 5;
    const uint32_t dm8 = (dm + 7) & ~7u;
    const PackedLayout L(TB, WS ? dm8 * NS : 0u, PS, SLOTS);
    double2 *s_t2 = (double2 *)(smem_raw + L.t2);  // [NS][TB + 1]
    constexpr int T2S = PS > 0 ? TB + 1 : TB;     // slot TB of every row: a document that never ranks
    uint8_t *s_cls = (uint8_t *)(smem_raw + L.cls);
    typedef typename SurvIndex<TB>::type surv_t;
    surv_t *s_surv = (surv_t *)(smem_raw + L.surv) + (size_t)(t >> 5) * (PS + 1) * 32;
    unsigned long long *s_sum = (unsigned long long *)(smem_raw + L.sum);
    double *s_roww = (double *)(smem_raw + L.roww);
    double *s_tbl = (double *)(smem_raw + L.tbl);
    uint2 *s_qd = (uint2 *)(smem_raw + L.qd);
    uint4 *s_tasks = (uint4 *)(smem_raw + L.tasks);
    uint16_t *s_order = (uint16_t *)(smem_raw + L.order);
    uint8_t *s_rowsw = (uint8_t *)(smem_raw + L.rowsw);
    int *s_misc = (int *)(smem_raw + L.misc);
    const double *__restrict__ wg = A.base_wt + (size_t)blockIdx.y * dm8 * NS;
    double *s_w = (double *)(smem_raw + L.w);
    const uint32_t tbl_r = V.tbl_r;
    uint8_t *s_slot = (uint8_t *)(smem_raw + L.slots) + (size_t)(t >> 5) * TB * 32 + lane;  // [rank][lane] of this warp
    // the register mode's table always fits shared memory (<= 16 classes x 16 ranks); slot mode may
    // have to read a larger one through L1
    const bool tbl_smem = !SLOTS || (V.n_cls + 1) * tbl_r <= 256u;

    // ---- per-launch setup ----
    for (int idx = t; idx < kMaxRows; idx += TB) {
        s_rowsw[idx] = idx < R ? (uint8_t)A.row_meta[row0 + idx] : (uint8_t)0;
        s_roww[idx] = idx < R ? A.row_w[row0 + idx] : 0.0;
        s_sum[idx] = 0ull;
    }
    if (tbl_smem && V.tbl != nullptr)
        for (uint32_t idx = t; idx < (V.n_cls + 1) * tbl_r; idx += TB) s_tbl[idx] = V.tbl[idx];
    if (PS > 0) {
        if (t < NS) s_t2[(size_t)t * T2S + TB] = make_double2(__longlong_as_double(0xfff0000000000000ll), 0.0);
        if (t == 0) s_cls[TB] = (uint8_t)255;  // tag (255 + 1) & 15 = 0: files nothing
    }
    if (WS)
        for (uint32_t idx = t; idx < dm8 * NS; idx += TB) s_w[idx] = wg[idx];
    __syncthreads();
    if (t < NS) {
        // a sweep without rows in this pass is not computed
        bool has_rows = false;
        for (int r = 0; r < R; ++r) has_rows |= (int)s_rowsw[r] == t;
        s_misc[M_F + t] = (t < ns && has_rows) ? (int)A.fid[s0 + t] : -1;
    }
    if (t == 0) s_misc[M_TILE] = (int)atomicAdd(A.tile_ctr + blockIdx.y, 1u);
    __syncthreads();
    uint32_t fs[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) fs[s] = (uint32_t)s_misc[M_F + s];
    int nan_seen = 0;

    // Work items: whole tiles first; the last n_split tiles are handed out as quarter items (a
    // quarter of the row groups each, phase 1 repeated) so that the SMs drain together.
    const uint32_t parts = A.split_parts;  // 2 or 4 row-group ranges per split tile
    const uint32_t n_split = (uint32_t)G >= parts ? min(P.nt, A.n_split) : 0u;
    const uint32_t n_whole = P.nt - n_split;
    const uint32_t n_items = G > 0 ? n_whole + parts * n_split : 0u;
    uint32_t next_item = n_items;
    for (uint32_t item = G > 0 ? (uint32_t)s_misc[M_TILE] : n_items; item < n_items; item = next_item) {
        if (t == 0) s_misc[M_NEXT] = (int)atomicAdd(A.tile_ctr + blockIdx.y, 1u);
        uint32_t tile = item;
        int g_begin = 0, g_end = G;
        if (item >= n_whole) {
            const uint32_t j = item - n_whole, part = j % parts;
            tile = n_whole + j / parts;
            g_begin = (int)((uint32_t)G * part / parts);
            g_end = (int)((uint32_t)G * (part + 1u) / parts);
        }
        const uint32_t doc0 = P.tile_doc_off[tile];
        const int nd = (int)(P.tile_doc_off[tile + 1] - doc0);
        const bool active = t < nd;
        const uint32_t pos = active ? P.pd_pos[doc0 + t] : 0u;
        const uint32_t task0 = V.tile_task_off[tile];
        const int ntask = (int)(V.tile_task_off[tile + 1] - task0);
        const uint32_t q0 = P.tile_q_off[tile];
        const int nqt = (int)(P.tile_q_off[tile + 1] - q0);
        if (t < ntask) s_tasks[t] = V.tasks[task0 + t];
        if (PS > 0 || SLOTS) s_cls[t] = active ? __ldg(V.pd_cls + doc0 + t) : (uint8_t)255;
        if (t < nqt) {
            const uint32_t tb0 = V.q_task_off[q0 + t], tb1 = V.q_task_off[q0 + t + 1];
            s_qd[t] = make_uint2(P.pq_local[q0 + t], (tb0 - task0) | ((tb1 - tb0) << 16));
            s_order[t] = V.q_order[q0 + t];
        }

        // ---- phase 1: one pass over the tile's features for every sweep (as sweep_fast_kernel) ----
        {
            double acc[NS];
            float xf[NS];
            const float *__restrict__ xp = P.x + pos;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s] = 0.0;
                xf[s] = fs[s] < dm ? ld_stream(xp + (size_t)fs[s] * P.ld) : 0.f;
            }
            float cur[8], nxt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] = (uint32_t)u < dm ? ld_stream(xp + (size_t)u * P.ld) : 0.f;
            for (uint32_t j0 = 0; j0 < dm; j0 += 8) {
                if (j0 + 16 <= dm) {
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
            // A score T + x_f * c can only be NaN (the reference's panic, model.rs:49) when T or x_f
            // is not finite -- candidates are checked on the host -- so the per-candidate test runs
            // for such documents only.
            bool odd = false;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                odd |= !(fabs(acc[s]) <= 1.7976931348623157e308) || !(fabsf(xf[s]) <= 3.402823466e38f);
                s_t2[(size_t)s * T2S + t] = make_double2(active ? acc[s] : 0.0, active ? (double)xf[s] : 0.0);
            }
            if (odd && active) {
                for (int r = (g_begin << 5); r < min(R, g_end << 5); ++r) {
                    const int s = s_rowsw[r];
                    double ss = acc[0];
                    float xs32 = xf[0];
#pragma unroll
                    for (int u = 1; u < NS; ++u) {
                        if (s == u) {
                            ss = acc[u];
                            xs32 = xf[u];
                        }
                    }
                    const double sc = __dadd_rn(ss, __dmul_rn((double)xs32, s_roww[r]));
                    if (sc != sc) nan_seen = 1;
                }
            }
        }
        if (t == 0) s_misc[M_CTR] = 0;
        __syncthreads();

        // ---- warp items: (query, row group) ----
        const int ng = g_end - g_begin;
        const int nwork = nqt * ng;
        for (;;) {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(&s_misc[M_CTR], 1);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= nwork) break;
            const int k = idx / ng;
            const int g = g_begin + (idx - k * ng);
            const int ql = s_order[k];
            const uint2 qd = s_qd[ql];
            const int qs = (int)(qd.x & 0xffffu), len = (int)(qd.x >> 16), qe = qs + len;
            const int tk0 = (int)(qd.y & 0xffffu), ntk = (int)(qd.y >> 16);
            const unsigned lim = (P.metric == FR_METRIC_NDCG && (unsigned)P.depth < (unsigned)len) ? (unsigned)P.depth
                                                                                                : (unsigned)len;
            const int row = (g << 5) + lane;
            const bool live = row < R;
            const int rowc = live ? row : R - 1;
            const double c = s_roww[rowc];
            const double2 *trow = s_t2 + (size_t)s_rowsw[rowc] * T2S;
            unsigned long long packed = 0ull;
            bool ranked = false;
            if (PS > 0 && A.prune_min > 0 && len >= (int)A.prune_min && ntk > 0)
                ranked = rank_pruned<TB, (PS > 0 ? PS : 8)>(trow, c, qs, qe, lim, s_cls, s_surv, lane, packed);
            if (!ranked) packed = 0ull;
            // RR needs one rank only, the best relevant document's: the relevant document with the
            // highest score (the earliest of equals -- list order breaks ties), then one count of
            // the documents that outrank it.  O(len) per candidate instead of relevant x len.
            const bool first_rank_only = SLOTS && P.metric == FR_METRIC_RR;
            double value = 0.0;
            if (first_rank_only && ntk > 0) {
                int bi = -1;
                double best = 0.0;
                for (int tk = tk0; tk < tk0 + ntk; ++tk) {
                    const unsigned wx = s_tasks[tk].x;
                    const int t0 = (int)(wx & 0xffffu), n = (int)(wx >> 16);
                    for (int i = t0; i < t0 + n; ++i) {
                        const double2 tx = trow[i];
                        const double si = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
                        if (bi < 0 || si > best) {
                            best = si;
                            bi = i;
                        }
                    }
                }
                unsigned above = 0;
#pragma unroll 4
                for (int jq = qs; jq < qe; ++jq) {
                    const double2 tx = trow[jq];
                    const double sj = __dadd_rn(tx.x, __dmul_rn(tx.y, c));
                    above += (jq < bi ? sj >= best : sj > best) ? 1u : 0u;
                }
                value = 1.0 / (double)(above + 1u);  // evaluators.rs:235-253
            }
            if (SLOTS && !first_rank_only) {  // clear the ranks this list can reach (lane-private column of the warp's buffer)
                for (unsigned r = 0; r < lim; ++r) s_slot[r * 32] = (uint8_t)0;
            }
            const bool class_tags = P.metric == FR_METRIC_NDCG;
            for (int tk = tk0; tk < tk0 + ((ranked || first_rank_only) ? 0 : ntk); ++tk) {
                const uint4 w = s_tasks[tk];
                const int t0 = (int)(w.x & 0xffffu), n = (int)(w.x >> 16);
                const unsigned long long tags = (unsigned long long)w.y | ((unsigned long long)w.z << 32);
                if (n > 8)
                    rank_chunk<16, SLOTS>(trow, c, qs, qe, t0, n, tags, packed, lim, s_cls, s_slot, class_tags);
                else if (n > 4 || PK_NO_W4)
                    rank_chunk<8, SLOTS>(trow, c, qs, qe, t0, n, tags, packed, lim, s_cls, s_slot, class_tags);
                else
                    rank_chunk<4, SLOTS>(trow, c, qs, qe, t0, n, tags, packed, lim, s_cls, s_slot, class_tags);
            }
            // fold: ranks 0 .. lim-1 in order (evaluators.rs:265-270); an empty slot adds +0.0 like a
            // zero-gain document does in the reference
            const uint32_t pq = q0 + ql;
            const double norm = P.pq_norm[pq];
            if (first_rank_only) {
                // value set above
            } else if (!SLOTS || P.metric == FR_METRIC_NDCG) {
                if (norm == norm) {  // Some(ideal), evaluators.rs:351-358
                    double dcg = 0.0;
                    for (unsigned r = 0; r < lim; ++r) {
                        const unsigned tag = SLOTS ? (unsigned)s_slot[r * 32] : (unsigned)(packed >> (r << 2)) & 15u;
                        const double term = (!SLOTS || tbl_smem) ? s_tbl[tag * tbl_r + r] : __ldg(V.tbl + tag * tbl_r + r);
                        dcg = __dadd_rn(dcg, term);
                    }
                    if (dcg > norm) atomicOr(A.err, ERR_DCG_ABOVE_IDEAL);
                    value = dcg / norm;
                }
            } else if (P.metric == FR_METRIC_AP) {  // evaluators.rs:418-448
                if (norm > 0.0) {
                    unsigned recall = 0;
                    double sum = 0.0;
                    // precision at each relevant rank: the quotients come from a host-built table of
                    // the same IEEE divisions when the plan has one (an f64 divide is ~40 instructions)
                    const double *__restrict__ apt = V.ap_tbl;
                    for (unsigned r = 0; r < lim; ++r) {
                        if (s_slot[r * 32]) {
                            recall += 1;
                            const double prec = apt ? __ldg(apt + (size_t)recall * V.ap_cols + r)
                                                    : (double)recall / (double)(r + 1);
                            sum = __dadd_rn(sum, prec);
                        }
                    }
                    value = sum / norm;
                }
            }
            if (live) {
                if (A.perq) A.perq[(size_t)A.row_out[row0 + row] * P.nq_view + P.pq_view[pq]] = value;
                const long long fx = __double2ll_rn(value * kFx);
                atomicAdd(&s_sum[row], (unsigned long long)fx);
            }
        }
        next_item = (uint32_t)s_misc[M_NEXT];
        __syncthreads();  // T2 and the tile tables are rewritten by the next item
    }
    if (nan_seen) atomicOr(A.err, ERR_NAN_SCORE);
    __syncthreads();
    for (int idx = t; idx < R; idx += TB)
        atomicAdd((unsigned long long *)(A.sums + A.row_out[row0 + idx]), s_sum[idx]);
    if (A.mail_peers == nullptr && A.host_flag == nullptr) return;
    fused_allreduce_tail<TB>(A, s_misc);
}
