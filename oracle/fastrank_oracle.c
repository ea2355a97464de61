/*
 * fastrank_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the algorithm on jjfiv/fastrank's
 * score -> rank -> metric hot path and of the coordinate-ascent driver that calls it.
 * It exists so tests/ (and __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * --impl reference legs) can CHECK the CUDA path.  Nothing under fastrank_b200/ may
 * import, link or call this file: the product fails loudly without its CUDA library.
 *
 * Parity status: PINNED for scoring / ranking / NDCG / AP / RR against the reference's
 * own golden vectors (tests/test_oracle_golden.py):
 *   - six single-feature NDCG@5 values   (reference tests/test_with_example_data.py:16-23)
 *   - tie-order known answer [4,3,1,2,5] (reference src/evaluators.rs:61-79)
 *   - compute_dcg known answer 0.7328    (reference src/evaluators.rs:285-295)
 * PINNED for the random-number stream as well: the reference draws from the crate
 * `oorandom` pinned `=11.1.0` (reference Cargo.toml:18-19), which is not vendored under
 * /root/reference.  fro_rng_* restates that version's published algorithm -- including its
 * two quirks: the 128 -> 64 bit output function is `rotr64((u64)(((s >> 29) ^ s) >> 58), s >> 122)`
 * (not PCG's XSL-RR), and Rand64::rand_range() forgets to add range.start -- and is anchored
 * on artefacts the reference itself produced with it (tests/test_oracle_golden.py):
 *   - random-forest determinism golden 0.4367914517387043 (reference src/random_forest.rs:427-463,
 *     tests/test_with_example_data.py:175-201): RNG + sampling + induction + scoring + NDCG@5
 *   - the coordinate-ascent weights and NDCG@5 printed by examples/FastRankDemo.ipynb
 *     (cells 4-6, seed 1234567): RNG + reset + shuffle + the whole line-search driver
 *
 * The reference (Rust) cannot be compiled in this image (no cargo/rustc), so there is
 * no oracle/_ref build; every function below cites the reference lines it follows.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the reference never fuses
 * multiply-add, so neither may this file).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FRO_NDCG 0
#define FRO_AP 1
#define FRO_RR 2

/* ------------------------------------------------------------------------------------
 * RNG: oorandom::Rand64 as of the pinned crate version 11.1.0 (reference Cargo.toml:18-19),
 * restated from the crate's published algorithm.  Call sites in the reference:
 * coordinate_ascent.rs:27,53,199,212; randutil.rs:8,24; random_forest.rs:143,293,306;
 * evaluators.rs:161-165.  Known answers: seed 42 -> 12410087264455502793, 359948335059059064,
 * 14020691344464510033; Rand64::new(0xdeadbeef).rand_u64() = 8208548815909702348 (the default
 * seed of both learners, coordinate_ascent.rs:27-34).
 * ---------------------------------------------------------------------------------- */
typedef unsigned __int128 u128;
typedef struct {
    u128 state;
    u128 inc;
} fro_rng;

static u128 fro_pcg_mult(void) {
    /* 47026247687942121848144207491837523525 = 0x2360ED051FC65DA44385DF649FCCF645 */
    return (((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL;
}
static u128 fro_pcg_default_inc(void) {
    return (((u128)0x2FE0E169FFBD06E3ULL) << 64) | (u128)0x5BC307BD4D2F814FULL;
}

uint64_t fro_rng_u64(fro_rng *r) {
    u128 old = r->state;
    r->state = old * fro_pcg_mult() + r->inc;
    /* 11.1.0's output function: a 128-bit xorshift by 29, bits 58..121 kept, rotated right
     * by the top six state bits (PCG's 64/32 XSH-RR recipe transplanted to 128/64). */
    uint32_t rot = (uint32_t)(old >> 122);
    uint64_t xsh = (uint64_t)(((old >> 29) ^ old) >> 58);
    return (xsh >> rot) | (xsh << ((64 - rot) & 63));
}

void fro_rng_seed(fro_rng *r, uint64_t seed_lo, uint64_t seed_hi) {
    u128 seed = (((u128)seed_hi) << 64) | seed_lo;
    r->state = 0;
    r->inc = (fro_pcg_default_inc() << 1) | 1;
    (void)fro_rng_u64(r);
    r->state += seed;
    (void)fro_rng_u64(r);
}

/* raw state injection, used by the test that cross-checks the 128-bit LCG step against
 * numpy.random.PCG64 (same multiplier and increment convention, different output function). */
void fro_rng_set_raw(fro_rng *r, uint64_t st_lo, uint64_t st_hi, uint64_t inc_lo, uint64_t inc_hi) {
    r->state = (((u128)st_hi) << 64) | st_lo;
    r->inc = (((u128)inc_hi) << 64) | inc_lo;
}

double fro_rng_float(fro_rng *r) {
    /* oorandom keeps MANTISSA_DIGITS+1 = 54 high bits and scales by 2^-54. */
    uint64_t u = fro_rng_u64(r) >> 10;
    return (double)u * (1.0 / 18014398509481984.0);
}

uint64_t fro_rng_range(fro_rng *r, uint64_t start, uint64_t end) {
    /* Lemire's nearly-divisionless bounded draw over [0, end-start).  Version 11.1.0 returns
     * it WITHOUT adding range.start (Rand32 adds it, Rand64 does not) -- the reason the crate
     * is pinned "to prevent seeds from changing" (Cargo.toml:18) -- so shuffle's
     * rand_range(i..n) (randutil.rs:24) yields [0, n-i).  The goldens bake this in. */
    uint64_t s = end - start;
    u128 m = (u128)fro_rng_u64(r) * (u128)s;
    uint64_t leftover = (uint64_t)m;
    if (leftover < s) {
        uint64_t threshold = (0 - s) % s;
        while (leftover < threshold) {
            m = (u128)fro_rng_u64(r) * (u128)s;
            leftover = (uint64_t)m;
        }
    }
    return (uint64_t)(m >> 64);
}

/* evaluators.rs:157-171: SetEvaluator::bootstrap_eval -- `trials` means of n draws with
 * replacement from the per-query values, one Rand64::new(0xdeadbeef) stream across all trials,
 * a sequential f64 sum per trial. */
void fro_bootstrap_means(const double *values, uint64_t n, uint32_t trials, double *out_means) {
    fro_rng r;
    fro_rng_seed(&r, 0xdeadbeefULL, 0);
    for (uint32_t t = 0; t < trials; t++) {
        double sum = 0.0;
        for (uint64_t k = 0; k < n; k++) {
            uint64_t index = fro_rng_range(&r, 0, n);
            sum = sum + values[index];
        }
        out_means[t] = sum / (double)n;
    }
}

/* randutil.rs:21-27 */
static void fro_shuffle_u32(uint32_t *v, size_t n, fro_rng *r) {
    for (size_t i = 0; i < n; i++) {
        size_t j = (size_t)fro_rng_range(r, i, n);
        uint32_t t = v[i];
        v[i] = v[j];
        v[j] = t;
    }
}

/* ------------------------------------------------------------------------------------
 * Scoring
 * ---------------------------------------------------------------------------------- */

/* dense_dataset.rs:67-76 + model.rs:47-51: s_i = sum_j f64(x_ij) * w_j, j ascending,
 * zip() truncates to the shorter of the row and the weight vector. */
void fro_score_linear(size_t n, size_t d, const float *x, const double *w, size_t nw,
                      double *out) {
    size_t m = d < nw ? d : nw;
    for (size_t i = 0; i < n; i++) {
        const float *row = x + i * d;
        double acc = 0.0;
        for (size_t j = 0; j < m; j++) {
            double p = (double)row[j] * w[j];
            acc = acc + p;
        }
        out[i] = acc;
    }
}

/* Model "bytecode" (array of doubles, prefix order) built by oracle/oracle.py from the
 * reference's model JSON (model.rs:10-16):
 *   SingleFeature : 0, fid, dir
 *   Linear        : 1, n, w[0..n)
 *   FeatureSplit  : 2, fid, split, len(lhs code), <lhs>, <rhs>
 *   LeafNode      : 3, value
 *   Ensemble      : 4, m, weight[0..m), <model 0> ... <model m-1>
 */
static double fro_feature_get(const float *row, size_t d, size_t fid) {
    /* instance.rs:63-64 / dense_dataset.rs:139-143: Some(f64(x)) inside the row, callers
     * unwrap_or(0.0) when the feature is absent. */
    return fid < d ? (double)row[fid] : 0.0;
}

static double fro_model_eval(const double *c, size_t *pos, const float *row, size_t d);

static void fro_model_skip(const double *c, size_t *pos) {
    int kind = (int)c[*pos];
    switch (kind) {
    case 0: *pos += 3; break;
    case 1: *pos += 2 + (size_t)c[*pos + 1]; break;
    case 2: {
        size_t lhs_len = (size_t)c[*pos + 3];
        *pos += 4 + lhs_len;
        fro_model_skip(c, pos);
        break;
    }
    case 3: *pos += 2; break;
    default: {
        size_t m = (size_t)c[*pos + 1];
        *pos += 2 + m;
        for (size_t t = 0; t < m; t++) fro_model_skip(c, pos);
    }
    }
}

static double fro_model_eval(const double *c, size_t *pos, const float *row, size_t d) {
    int kind = (int)c[*pos];
    switch (kind) {
    case 0: { /* model.rs:35-40 */
        double v = fro_feature_get(row, d, (size_t)c[*pos + 1]);
        double dir = c[*pos + 2];
        *pos += 3;
        return dir * v;
    }
    case 1: { /* model.rs:47-51 */
        size_t n = (size_t)c[*pos + 1];
        const double *w = c + *pos + 2;
        size_t m = d < n ? d : n;
        double acc = 0.0;
        for (size_t j = 0; j < m; j++) {
            double p = (double)row[j] * w[j];
            acc = acc + p;
        }
        *pos += 2 + n;
        return acc;
    }
    case 2: { /* model.rs:64-84: fval <= split -> lhs else rhs */
        double fval = fro_feature_get(row, d, (size_t)c[*pos + 1]);
        double split = c[*pos + 2];
        size_t lhs_len = (size_t)c[*pos + 3];
        size_t lhs_pos = *pos + 4;
        size_t rhs_pos = lhs_pos + lhs_len;
        double out;
        if (fval <= split) {
            size_t p = lhs_pos;
            out = fro_model_eval(c, &p, row, d);
        } else {
            size_t p = rhs_pos;
            out = fro_model_eval(c, &p, row, d);
        }
        size_t end = rhs_pos;
        fro_model_skip(c, &end);
        *pos = end;
        return out;
    }
    case 3: {
        double v = c[*pos + 1];
        *pos += 2;
        return v;
    }
    default: { /* model.rs:104-112: output += weight * member, member order */
        size_t m = (size_t)c[*pos + 1];
        const double *w = c + *pos + 2;
        *pos += 2 + m;
        double acc = 0.0;
        for (size_t t = 0; t < m; t++) {
            double s = fro_model_eval(c, pos, row, d);
            double p = w[t] * s;
            acc = acc + p;
        }
        return acc;
    }
    }
}

void fro_score_model(size_t n, size_t d, const float *x, const double *code, double *out) {
    for (size_t i = 0; i < n; i++) {
        size_t pos = 0;
        out[i] = fro_model_eval(code, &pos, x + i * d, d);
    }
}

/* ------------------------------------------------------------------------------------
 * Ranking: evaluators.rs:33-49  (score desc, then gain asc, then instance id asc).
 * NotNan ordering: -0.0 == +0.0.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    double score;
    float gain;
    uint32_t id;
} fro_ranked;

static int fro_ranked_cmp(const void *pa, const void *pb) {
    const fro_ranked *a = (const fro_ranked *)pa, *b = (const fro_ranked *)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    if (a->gain < b->gain) return -1;
    if (a->gain > b->gain) return 1;
    if (a->id < b->id) return -1;
    if (a->id > b->id) return 1;
    return 0;
}

/* Sort one query's documents; writes the instance ids in ranked order. */
void fro_rank_query(size_t len, const uint32_t *ids, const double *scores, const float *gains,
                    uint32_t *out_ids) {
    fro_ranked *r = (fro_ranked *)malloc(sizeof(fro_ranked) * (len ? len : 1));
    for (size_t i = 0; i < len; i++) {
        r[i].score = scores[i];
        r[i].gain = gains[i];
        r[i].id = ids[i];
    }
    qsort(r, len, sizeof(fro_ranked), fro_ranked_cmp);
    for (size_t i = 0; i < len; i++) out_ids[i] = r[i].id;
    free(r);
}

static int fro_f32_desc(const void *pa, const void *pb) {
    float a = *(const float *)pa, b = *(const float *)pb;
    return (a < b) - (a > b);
}

/* evaluators.rs:255-272.  depth < 0 means "no depth". */
double fro_compute_dcg(const float *gains, size_t len, int64_t depth, int ideal) {
    size_t m = len;
    if (depth >= 0) m = (size_t)depth;
    float *g = (float *)calloc(m ? m : 1, sizeof(float));
    float *src = (float *)malloc(sizeof(float) * (len ? len : 1));
    memcpy(src, gains, sizeof(float) * len);
    if (ideal) qsort(src, len, sizeof(float), fro_f32_desc);
    for (size_t i = 0; i < m && i < len; i++) g[i] = src[i]; /* resize(): truncate or 0-pad */
    double dcg = 0.0;
    for (size_t i = 0; i < m; i++) {
        double fi = (double)i;
        double gain = (double)g[i];
        dcg += (pow(2.0, gain) - 1.0) / log2(fi + 2.0);
    }
    free(g);
    free(src);
    return dcg;
}

/* ------------------------------------------------------------------------------------
 * Evaluation: evaluators.rs:186-224 (loop), :235-253 (RR), :303-381 (NDCG), :389-448 (AP).
 *
 * Layout: the dataset view is a list of queries; query q owns the instance ids
 * qdocs[qoff[q] .. qoff[q+1]) in the order the reference pushes them (ascending id).
 * scores[] and gains[] are indexed by instance id.
 * Optional judgments (qrel.rs:21-39): qrel_present[q] != 0 when the qrel has the query,
 * with its judged gains in qrel_gains[qrel_off[q] .. qrel_off[q+1]).
 * Returns 0, or 1 when the reference would panic (actual DCG > ideal DCG, :369-374).
 * ---------------------------------------------------------------------------------- */
int fro_evaluate(int metric, int64_t depth, size_t nq, const uint64_t *qoff,
                 const uint32_t *qdocs, const double *scores, const float *gains,
                 const uint8_t *qrel_present, const uint64_t *qrel_off, const float *qrel_gains,
                 double *out_per_query) {
    int status = 0;
    for (size_t q = 0; q < nq; q++) {
        size_t len = (size_t)(qoff[q + 1] - qoff[q]);
        const uint32_t *ids = qdocs + qoff[q];
        fro_ranked *r = (fro_ranked *)malloc(sizeof(fro_ranked) * (len ? len : 1));
        float *dataset_gains = (float *)malloc(sizeof(float) * (len ? len : 1));
        for (size_t i = 0; i < len; i++) {
            r[i].score = scores[ids[i]];
            r[i].gain = gains[ids[i]];
            r[i].id = ids[i];
            dataset_gains[i] = gains[ids[i]];
        }
        qsort(r, len, sizeof(fro_ranked), fro_ranked_cmp);
        int has_qrel = qrel_present != NULL && qrel_present[q];
        double value = 0.0;
        if (metric == FRO_NDCG) {
            /* NDCG::new, :303-339: positive judged gains if the qrel knows the query, else
             * every gain of the query's instances in this dataset. */
            const float *ig = dataset_gains;
            size_t ilen = len;
            float *pos_gains = NULL;
            if (has_qrel) {
                size_t cnt = (size_t)(qrel_off[q + 1] - qrel_off[q]);
                pos_gains = (float *)malloc(sizeof(float) * (cnt ? cnt : 1));
                ilen = 0;
                for (size_t i = 0; i < cnt; i++) {
                    float g = qrel_gains[qrel_off[q] + i];
                    if (g > 0.0f) pos_gains[ilen++] = g;
                }
                ig = pos_gains;
            }
            size_t npos = 0;
            for (size_t i = 0; i < ilen; i++) npos += ig[i] > 0.0f;
            if (npos > 0) {
                double ideal = fro_compute_dcg(ig, ilen, depth, 1);
                float *ranked_gains = (float *)malloc(sizeof(float) * (len ? len : 1));
                for (size_t i = 0; i < len; i++) ranked_gains[i] = r[i].gain;
                double actual = fro_compute_dcg(ranked_gains, len, depth, 0);
                free(ranked_gains);
                if (actual > ideal) status = 1;
                value = actual / ideal;
            }
            free(pos_gains);
        } else if (metric == FRO_AP) {
            /* AveragePrecision::new :389-415 keeps a norm only when it is > 0; score()
             * :422-447 falls back to the ranked list's own relevant count otherwise. */
            uint32_t num_rel = 0;
            if (has_qrel) {
                size_t cnt = (size_t)(qrel_off[q + 1] - qrel_off[q]);
                for (size_t i = 0; i < cnt; i++) num_rel += qrel_gains[qrel_off[q] + i] > 0.0f;
            } else {
                for (size_t i = 0; i < len; i++) num_rel += dataset_gains[i] > 0.0f;
            }
            if (num_rel == 0)
                for (size_t i = 0; i < len; i++) num_rel += r[i].gain > 0.0f;
            if (num_rel > 0) {
                uint32_t recall_points = 0;
                double sum_precision = 0.0;
                for (size_t i = 0; i < len; i++) {
                    if (r[i].gain > 0.0f) {
                        recall_points += 1;
                        sum_precision += (double)recall_points / (double)(i + 1);
                    }
                }
                value = sum_precision / (double)num_rel;
            }
        } else {
            for (size_t i = 0; i < len; i++) {
                if (r[i].gain > 0.0f) {
                    value = 1.0 / (double)(i + 1);
                    break;
                }
            }
        }
        out_per_query[q] = value;
        free(r);
        free(dataset_gains);
    }
    return status;
}

/* evaluators.rs:173-184.  The reference sums in HashMap iteration order (random per
 * call); the oracle fixes the order to the caller's query order. */
double fro_mean(const double *v, size_t n) {
    if (n == 0) return 0.0;
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += v[i];
    return sum / (double)n;
}

/* ------------------------------------------------------------------------------------
 * Coordinate ascent: coordinate_ascent.rs:43-253, core.rs:57-66.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint32_t num_restarts;
    uint32_t num_max_iterations;
    double step_base;
    double step_scale;
    double tolerance;
    uint64_t seed;
    int32_t normalize;
    int32_t init_random;
    int32_t output_ensemble;
    int32_t metric;
    int64_t depth;
} fro_ca_params;

typedef struct {
    size_t n, d;
    const float *x;
    const float *gains;
    size_t nq;
    const uint64_t *qoff;
    const uint32_t *qdocs;
    const uint8_t *qrel_present;
    const uint64_t *qrel_off;
    const float *qrel_gains;
    double *scores;    /* scratch n */
    double *per_query; /* scratch nq */
    uint64_t n_evals;
} fro_ca_ctx;

static double fro_ca_eval(fro_ca_ctx *c, const fro_ca_params *p, const double *w, size_t dim) {
    fro_score_linear(c->n, c->d, c->x, w, dim, c->scores);
    fro_evaluate(p->metric, p->depth, c->nq, c->qoff, c->qdocs, c->scores, c->gains,
                 c->qrel_present, c->qrel_off, c->qrel_gains, c->per_query);
    c->n_evals++;
    return fro_mean(c->per_query, c->nq);
}

/* coordinate_ascent.rs:72-82 */
static void fro_l1_normalize(double *w, size_t dim) {
    double sum = 0.0;
    for (size_t i = 0; i < dim; i++) sum += fabs(w[i]);
    if (sum > 0.0)
        for (size_t i = 0; i < dim; i++) w[i] /= sum;
}

/* coordinate_ascent.rs:87-195.  Returns the restart's best score; best weights in out_w. */
static double fro_ca_optimize(fro_ca_ctx *c, const fro_ca_params *p, fro_rng rng,
                              const uint32_t *fids, size_t nf, size_t dim, double *out_w) {
    static const int SIGN[3] = {0, -1, 1};
    double *model = (double *)calloc(dim, sizeof(double));
    double *best = (double *)calloc(dim, sizeof(double));
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * nf);
    /* reset(): :50-70 */
    for (size_t i = 0; i < nf; i++) {
        if (p->init_random)
            model[fids[i]] = (fro_rng_float(&rng) * 2.0) - 1.0;
        else
            model[fids[i]] = 1.0 / (double)nf;
    }
    double best_score = fro_ca_eval(c, p, model, dim);
    memcpy(best, model, sizeof(double) * dim);
    for (;;) {
        memcpy(order, fids, sizeof(uint32_t) * nf);
        fro_shuffle_u32(order, nf, &rng);
        size_t successes = 0;
        for (size_t fi = 0; fi < nf; fi++) {
            uint32_t f = order[fi];
            double start_score = best_score;
            memcpy(model, best, sizeof(double) * dim);
            if (p->normalize) fro_l1_normalize(model, dim);
            double orig = model[f];
            for (int di = 0; di < 3; di++) {
                int dir = SIGN[di];
                double step = p->step_base * (double)dir;
                if (orig != 0.0 && fabs(step) > 0.5 * fabs(orig))
                    step = p->step_base * fabs(orig) * (double)dir;
                double total = step;
                uint32_t iters = p->num_max_iterations;
                if (dir == 0) {
                    iters = 1;
                    total = -orig;
                }
                for (uint32_t it = 0; it < iters; it++) {
                    double w = orig + total;
                    model[f] = w;
                    double s = fro_ca_eval(c, p, model, dim);
                    if (s == s && s > best_score) { /* core.rs:57-66 strict >, NaN never wins */
                        best_score = s;
                        memcpy(best, model, sizeof(double) * dim);
                    }
                    step *= p->step_scale;
                    total += step;
                }
                if ((best_score - start_score) > p->tolerance) break;
            }
            if ((best_score - start_score) > p->tolerance) successes++;
        }
        if (successes == 0) break;
    }
    memcpy(out_w, best, sizeof(double) * dim);
    free(model);
    free(best);
    free(order);
    return best_score;
}

/* coordinate_ascent.rs:197-253.  fids = dataset.features() (ascending for a dense or
 * loaded dataset).  out_w: num_restarts x dim weights (every restart, for inspection);
 * out_scores: per-restart best score; returns the index of the restart the reference
 * returns as ModelEnum::Linear (Iterator::max: the LAST maximal element). */
int64_t fro_coordinate_ascent(const fro_ca_params *p, size_t n, size_t d, const float *x,
                              const float *gains, size_t nq, const uint64_t *qoff,
                              const uint32_t *qdocs, const uint8_t *qrel_present,
                              const uint64_t *qrel_off, const float *qrel_gains,
                              const uint32_t *fids, size_t nf, double *out_w, double *out_scores,
                              uint64_t *out_n_evals) {
    if (nf == 0 || n == 0 || nq == 0) return -1;
    size_t dim = 0;
    for (size_t i = 0; i < nf; i++)
        if ((size_t)fids[i] + 1 > dim) dim = (size_t)fids[i] + 1;
    fro_ca_ctx c;
    c.n = n; c.d = d; c.x = x; c.gains = gains; c.nq = nq; c.qoff = qoff; c.qdocs = qdocs;
    c.qrel_present = qrel_present; c.qrel_off = qrel_off; c.qrel_gains = qrel_gains;
    c.scores = (double *)malloc(sizeof(double) * n);
    c.per_query = (double *)malloc(sizeof(double) * nq);
    c.n_evals = 0;
    fro_rng master;
    fro_rng_seed(&master, p->seed, 0);
    int64_t best_idx = -1;
    for (uint32_t r = 0; r < p->num_restarts; r++) {
        fro_rng local;
        fro_rng_seed(&local, fro_rng_u64(&master), 0);
        out_scores[r] = fro_ca_optimize(&c, p, local, fids, nf, dim, out_w + (size_t)r * dim);
        if (best_idx < 0 || out_scores[r] >= out_scores[best_idx]) best_idx = (int64_t)r;
    }
    if (out_n_evals) *out_n_evals = c.n_evals;
    free(c.scores);
    free(c.per_query);
    return best_idx;
}

/* Restart RNG derivation alone (coordinate_ascent.rs:199,211-213) plus the first draws a
 * restart makes, so tests can compare the product's host-side RNG with the oracle's. */
void fro_ca_restart_streams(uint64_t seed, uint32_t num_restarts, uint32_t draws_per_restart,
                            double *out_floats) {
    fro_rng master;
    fro_rng_seed(&master, seed, 0);
    for (uint32_t r = 0; r < num_restarts; r++) {
        fro_rng local;
        fro_rng_seed(&local, fro_rng_u64(&master), 0);
        for (uint32_t k = 0; k < draws_per_restart; k++)
            out_floats[(size_t)r * draws_per_restart + k] = fro_rng_float(&local);
    }
}

size_t fro_sizeof_rng(void) { return sizeof(fro_rng); }
