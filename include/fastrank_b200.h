/*
 * fastrank_b200.h -- C ABI of libfastrank_b200.so
 *
 * Two layers, both `extern "C"`, plain pointers and sizes only:
 *
 *  (1) The OUTER boundary: the 20 symbols the reference exports from src/lib.rs and that
 *      fastrank/clib.py binds through cffi (`from .fastrank import lib, ffi`, clib.py:2).
 *      A maintainer swaps the maturin-built cdylib for this library and the Python API
 *      keeps working; see INTEGRATION.md.  Every declaration cites the reference symbol
 *      it replaces.
 *
 *  (2) The INNER boundary (`fr_dev_*`): the thin kernel ABI the host calls at exactly the
 *      seam where the reference's learners call into the hot path
 *      (SetEvaluator::evaluate_mean, coordinate_ascent.rs:110,160; evaluate_to_map,
 *      ffi.rs:256; per-document score, json_api.rs:62).  A Rust host could bind these
 *      unchanged.
 *
 * All compute behind both layers runs on the GPU (sm_100a).  There is no CPU fallback:
 * when no CUDA device is usable every compute entry point reports an error.
 */
#ifndef FASTRANK_B200_H
#define FASTRANK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ======================================================================================
 * (1) Reference-compatible surface (reference src/lib.rs)
 * ==================================================================================== */

typedef struct CDataset CDataset; /* lib.rs:50-53 */
typedef struct CModel CModel;     /* lib.rs:55-57 */
typedef struct CQRel CQRel;       /* lib.rs:59-61 */

/* lib.rs:63-67.  Exactly one field is non-NULL.  error_message is a JSON
 * {"error":..,"context":..} string to be released with free_str; success is an owned
 * CDataset* / CModel* / CQRel*. */
typedef struct CResult {
    const void *error_message;
    const void *success;
} CResult;

void free_str(void *originally_from_native);          /* lib.rs:81  */
void free_c_result(CResult *originally_from_native);  /* lib.rs:89  (not recursive) */
void free_dataset(CDataset *originally_from_native);  /* lib.rs:96  */
void free_model(CModel *originally_from_native);      /* lib.rs:103 */
void free_cqrel(CQRel *originally_from_native);       /* lib.rs:110 */

const CResult *load_cqrel(const void *data_path);     /* lib.rs:117 -> CQRel* */
const CResult *cqrel_from_json(const void *json_str); /* lib.rs:126 -> CQRel* */
/* lib.rs:136; query_str in {"to_json","queries",<qid>} -> JSON string (free_str) */
const void *cqrel_query_json(const CQRel *cqrel, const void *query_str);

/* lib.rs:147 -> CDataset*; feature_names_path may be NULL */
const CResult *load_ranksvm_format(void *data_path, void *feature_names_path);
/* lib.rs:167 -> CDataset* restricted to the JSON list of query ids */
const CResult *dataset_query_sampling(CDataset *dataset, const void *queries_json_list);
/* lib.rs:183 -> CDataset* restricted to the JSON list of feature ids */
const CResult *dataset_feature_sampling(CDataset *dataset, const void *feature_json_list);
/* lib.rs:202; cmd in {is_sampled,num_features,feature_ids,num_instances,queries,
 * instances_by_query,feature_names} (ffi.rs:154-180) */
const void *dataset_query_json(void *dataset, void *json_cmd_str);
/* lib.rs:216; cmd in {coordinate_ascent_defaults, random_forest_defaults} */
const void *query_json(const void *json_cmd_str);

/* lib.rs:224: BORROWS x (n*d, row-major), y (n), qids (n) for the dataset's lifetime. */
const CResult *make_dense_dataset_f32_f64_i64(size_t n, size_t d, const float *x,
                                              const double *y, const int64_t *qids);

const CResult *train_model(void *train_request_json, void *dataset); /* lib.rs:245 -> CModel* */
const CResult *model_from_json(const void *json_str);                /* lib.rs:258 -> CModel* */
const void *model_query_json(const void *model, const void *json_cmd_str); /* lib.rs:268 */

/* lib.rs:283 -> {"qid": score, ...}; qrel may be NULL */
const void *evaluate_by_query(const CModel *model, const CDataset *dataset, const CQRel *qrel,
                              const void *evaluator);
/* lib.rs:299 -> {"instance index": score, ...} */
const void *predict_scores(const CModel *model, const CDataset *dataset);
/* lib.rs:308 -> number of records written (JSON integer) */
const void *predict_to_trecrun(const CModel *model, const CDataset *dataset,
                               const void *output_path, const void *system_name, size_t depth);

/* --------------------------------------------------------------------------------------
 * Binary fast paths beside the JSON forms.  Same semantics as
 * predict_scores / evaluate_by_query+mean, without the JSON round trip.
 * ------------------------------------------------------------------------------------ */
/* Scores for every instance id of the PARENT dataset, NaN where the view does not hold the
 * instance.  out has room for n_out doubles; returns NULL on success, else an error JSON
 * string (free_str). */
const void *predict_dense_f64(const CModel *model, const CDataset *dataset, double *out,
                              size_t n_out);
/* evaluators.rs:173-184 through the boundary: mean of the per-query metric. */
const void *evaluate_mean_f64(const CModel *model, const CDataset *dataset, const CQRel *qrel,
                              const void *evaluator, double *out_mean);

/* evaluators.rs:157-171 (SetEvaluator::bootstrap_eval; the reference calls it from
 * print_standard_eval with 200 trials): the model is evaluated per query, the per-query values are
 * resampled with replacement `num_trials` times with Rand64::new(0xdeadbeef), and out_means
 * receives the resampled means SORTED ascending (what PercentileStats::new holds, stats.rs:131-138).
 * Queries are taken in view order (the reference iterates a HashMap, so its own result varies from
 * run to run).  Returns NULL on success, else an error JSON string (free_str). */
const void *evaluate_bootstrap_f64(const CModel *model, const CDataset *dataset, const CQRel *qrel,
                                   const void *evaluator, uint32_t num_trials, double *out_means);

/* Device-side timing for reports (bench.py): brackets every scoring / ranking kernel this
 * dataset launches with CUDA events on the library's stream (fr_dev_profile_* below, reached
 * through the reference-facing handle).  enable: 1 on, 0 off, -1 leave as is; when out_launches /
 * out_total_ms are non-NULL they receive the launches bracketed since the last call and the sum
 * of their durations, and the record is reset.  Returns NULL on success, else an error JSON
 * string (free_str). */
const void *dataset_device_profile(const CDataset *dataset, int enable, uint64_t *out_launches,
                                   double *out_total_ms);

/* ======================================================================================
 * (2) Kernel ABI
 * ==================================================================================== */

typedef struct fr_dev_dataset fr_dev_dataset; /* column-major feature matrix resident in HBM */
typedef struct fr_dev_plan fr_dev_plan;       /* one SetEvaluator: query tiling, metric, norms */
typedef struct fr_dev_model fr_dev_model;     /* a flattened ModelEnum */
typedef struct fr_dev_comm fr_dev_comm;       /* NCCL communicator for the metric all-reduce */

#define FR_METRIC_NDCG 0 /* evaluators.rs:298-381 */
#define FR_METRIC_AP 1   /* evaluators.rs:383-448 */
#define FR_METRIC_RR 2   /* evaluators.rs:232-253 */

/* Per-query metric values are accumulated as signed fixed point with FR_FX_BITS fractional
 * bits so that sums are independent of reduction order, launch geometry and GPU count. */
#define FR_FX_BITS 40

/* All functions return 0 on success; otherwise a non-zero code, with the message available
 * from fr_dev_last_error() (thread-local, valid until the next failing call). */
const char *fr_dev_last_error(void);
/* Number of usable CUDA devices (0 when there is no driver / no GPU). */
int fr_dev_device_count(void);

/* Upload one dataset (dense_dataset.rs:11-56 / dataset.rs:196-256 after densification).
 *   x           n*d float32, row-major (row stride d)
 *   gains       n float32   (dense_dataset.rs:114-123: f32(y))
 *   query_index n uint32    dense query number of each instance, in [0, n_queries)
 * Rows are regrouped by query and, inside a query, ordered by (gain asc, instance id asc) --
 * the reference's tie-break (evaluators.rs:33-49) -- so that ranking on the device is a
 * stable descending sort by score.  Queries may have any length: lists that fit a tile (up to
 * 1024 documents) are ranked in shared memory, longer ones from HBM. */
int fr_dev_dataset_create(int device, size_t n, size_t d, const float *x, const float *gains,
                          const uint32_t *query_index, uint32_t n_queries, fr_dev_dataset **out);
/* Optional, for data loaded from libsvm files (instance.rs:104-122): row_len[i] = number of leading
 * feature ids instance i carries (a dense row of max own id + 1 entries); ids at or beyond it are
 * MISSING for that instance -- they read as 0.0 when scored or split on (model.rs:75,
 * random_forest.rs:228) but are skipped by FeatureStats (normalizers.rs:21-27), which is what the
 * random-forest statistics below honour.  Without this call nothing is missing. */
int fr_dev_dataset_set_row_lengths(fr_dev_dataset *ds, const uint32_t *row_len);
/* The same for data with Sparse32 rows (instance.rs:118-121: a row that lists fewer than half of the
 * ids up to its largest carries exactly the listed ones): bits[i * words_per_row + (f >> 5)] bit
 * (f & 31) = instance i carries feature f.  Takes precedence over the row lengths. */
int fr_dev_dataset_set_row_presence(fr_dev_dataset *ds, const uint32_t *bits, uint32_t words_per_row);
void fr_dev_dataset_destroy(fr_dev_dataset *ds);
size_t fr_dev_dataset_bytes(const fr_dev_dataset *ds); /* HBM footprint */

typedef struct fr_dev_plan_desc {
    int32_t metric; /* FR_METRIC_* */
    int64_t depth;  /* NDCG@depth; -1 = whole list */
    /* The view (dataset.rs:101-178): queries, in output order.  query_ids == NULL means
     * every query of the dataset in query_index order. */
    uint32_t n_queries;
    const uint32_t *query_ids;
    /* Optional instance subset per view query (sampling.rs:67-72): inst_ids[inst_off[q] ..
     * inst_off[q+1]) are the instance ids kept for view query q.  NULL = all of them. */
    const uint64_t *inst_off;
    const uint32_t *inst_ids;
    /* Optional norms from judgments (evaluators.rs:310-318, :397-401).  When
     * norm_present[q] != 0: NDCG uses norm_value[q] as the ideal DCG (NaN = the qrel has no
     * positive gain => score 0); AP uses it as num_relevant (0 => fall back to the list). */
    const uint8_t *norm_present;
    const double *norm_value;
} fr_dev_plan_desc;

int fr_dev_plan_create(fr_dev_dataset *ds, const fr_dev_plan_desc *desc, fr_dev_plan **out);
void fr_dev_plan_destroy(fr_dev_plan *plan);
/* Attach a communicator: every evaluation of this plan all-reduces its fixed-point sums and
 * query count across ranks. */
int fr_dev_plan_set_comm(fr_dev_plan *plan, fr_dev_comm *comm);
uint64_t fr_dev_plan_global_queries(const fr_dev_plan *plan);
/* How the plan laid the view out: documents per tile (whole queries are packed into tiles of
 * 128 .. 1024 documents, evaluators.rs:206-221 ranks one query at a time) and the number of
 * local queries too long for a tile, which are ranked from HBM instead. */
uint32_t fr_dev_plan_tile_documents(const fr_dev_plan *plan);
uint32_t fr_dev_plan_untiled_queries(const fr_dev_plan *plan);

/* evaluate_mean for C weight vectors in one pass over the matrix
 * (coordinate_ascent.rs:110 / evaluators.rs:173-224 with model.rs:47-51 scoring).
 *   w            C x wlen float64, row-major
 *   out_sum_fx   C fixed-point sums of the per-query metric (all-reduced when a comm is set);
 *                mean = out_sum_fx * 2^-FR_FX_BITS / fr_dev_plan_global_queries()
 *   out_per_query  NULL or C x n_queries float64 (local queries, view order) */
int fr_dev_eval_linear_batch(fr_dev_plan *plan, const double *w, size_t wlen, size_t n_cand,
                             int64_t *out_sum_fx, double *out_per_query);

/* One coordinate-ascent line search per restart (coordinate_ascent.rs:131-177): restart r
 * evaluates weight vectors equal to base_w[r] except that coordinate fid[r] takes each of
 * cand_w[r][0..n_cand[r]).  Arithmetic is the reference's left-to-right f64 dot product; the
 * part of it that does not depend on the candidate is shared.
 *   base_w  n_sweeps x wlen;  cand_w, out_sum_fx  n_sweeps x cand_stride */
int fr_dev_eval_coord_sweeps(fr_dev_plan *plan, size_t n_sweeps, const double *base_w, size_t wlen,
                             const uint32_t *fid, const double *cand_w, const uint32_t *n_cand,
                             size_t cand_stride, int64_t *out_sum_fx);

/* The same line searches, batched: ONE pass over the feature matrix serves every sweep of the
 * call (groups of 8 sweeps); a candidate's score is  T + x_f * cand  with T = the reference's
 * left-to-right f64 sum over the other coordinates, i.e. the reference's dot product with the
 * varied coordinate's term added last instead of in position (a few roundings, ~1e-16 relative,
 * from dense_dataset.rs:67-76).  Ranking, tie-break, metric terms and their summation order are
 * the reference's, so per-query values are bit-identical to the exact path whenever the ranking
 * is.  This is what train_model uses (FASTRANK_SWEEP=exact selects the entry point above
 * instead); it submits all three directions of a line search (1 + 2 * num_max_iterations
 * candidates per restart) in one call.  Tiles of the batched sweep hold up to 512 documents;
 * when at most a quarter of the view's documents sit in longer lists those lists are ranked from
 * HBM next to the tiles (scored with the exact dot product) and the call serves the whole view,
 * otherwise the plan stays on the exact-order kernels (fr_dev_plan_has_fast_sweep == 0).
 * out_per_query: NULL or n_sweeps x cand_stride x n_queries (view order). */
int fr_dev_eval_coord_sweeps_fast(fr_dev_plan *plan, size_t n_sweeps, const double *base_w,
                                  size_t wlen, const uint32_t *fid, const double *cand_w,
                                  const uint32_t *n_cand, size_t cand_stride, int64_t *out_sum_fx,
                                  double *out_per_query);
/* 1 when the batched sweep serves this plan (see above). */
int fr_dev_plan_has_fast_sweep(const fr_dev_plan *plan);
/* Name of the kernel that serves fr_dev_eval_coord_sweeps_fast on this plan (static string):
 * "sweep_packed_kernel<TILE>" for NDCG@k with k <= 16 and at most 15 gain classes (a candidate's
 * top-k kept in one 64-bit register), "sweep_packed_kernel<TILE,slots>" for the other measures on
 * tiles of up to 256 documents (ranks filed in a per-warp slot buffer), "sweep_fast_kernel<TILE,8>"
 * otherwise, "" when the plan has no batched sweep.  For reports (bench.py roofline.kernel). */
const char *fr_dev_plan_sweep_kernel(const fr_dev_plan *plan);

/* Bootstrap resampling of per-query values (evaluators.rs:157-171): `trials` means of n_values
 * draws with replacement from `values` (the per-query output of an evaluation, view order),
 * drawn with oorandom::Rand64::new(seed).rand_range(0..n) exactly as the reference does -- one
 * sequential stream across all trials, one sequential f64 sum per trial.  On the device each
 * trial is one thread started at its position in the stream (128-bit LCG jump-ahead); out_means
 * receives the means in trial order.  Single-GPU plans only. */
int fr_dev_plan_bootstrap(fr_dev_plan *plan, const double *values, size_t n_values, uint64_t seed,
                          uint32_t trials, double *out_means);

/* Flattened ModelEnum (model.rs:10-16).  `code` is a postfix program of 64-bit words; see
 * fastrank_b200/csrc/model_program.hpp for the encoding produced by the host. */
int fr_dev_model_create(fr_dev_dataset *ds, const uint64_t *code, size_t n_words,
                        fr_dev_model **out);
void fr_dev_model_destroy(fr_dev_model *m);
/* Scores of every instance, indexed by instance id (json_api.rs:53-72). */
int fr_dev_score_model(fr_dev_dataset *ds, const fr_dev_model *m, double *out_scores);
/* evaluate_to_map / evaluate_mean for an arbitrary model (ffi.rs:246-258). */
int fr_dev_eval_model(fr_dev_plan *plan, const fr_dev_model *m, int64_t *out_sum_fx,
                      double *out_per_query);

/* --------------------------------------------------------------------------------------
 * Random-forest induction statistics (random_forest.rs:211-286, :362-408; normalizers.rs:13-36),
 * one tree at a time, level by level.  The host keeps the reference's decision logic; the device
 * produces, for all active nodes of a level at once, what those decisions need.
 * Label sums are integers in units of 2^-FR_RF_GAIN_BITS (and their squares), so they do not
 * depend on the order of accumulation; the caller guarantees every label is a multiple of that
 * unit with magnitude <= 16.
 * ------------------------------------------------------------------------------------ */
#define FR_RF_GAIN_BITS 12
typedef struct fr_dev_rf fr_dev_rf;
int fr_dev_rf_create(fr_dev_dataset *ds, fr_dev_rf **out);
void fr_dev_rf_destroy(fr_dev_rf *rf);
/* A new tree over `m` sampled instance ids and `n_features` sampled feature ids; every instance
 * starts in active node 0 (the root). */
int fr_dev_rf_begin_tree(fr_dev_rf *rf, const uint32_t *instances, size_t m, const uint32_t *features,
                         size_t n_features);
/* Statistics of the current level, n_active nodes, k = split_candidates.  Outputs (host):
 *   node_n, node_sum [n_active]            instances and label sum of the node
 *   gmin, gmax [n_active]                  label range (label_stats, random_forest.rs:22-31)
 *   fmin, fmax [n_active][n_features]      FeatureStats min / max
 *   b_n, b_pos, b_sum, b_sq [n_active][n_features][k]
 *        per bucket between consecutive thresholds i/k * (max - min) + min, i = 1..k-1
 *        (bucket b holds the values v with threshold_b <= v < threshold_{b+1}):
 *        instances, instances with label > 0, label sum, sum of squared labels
 *   f_present [n_active][n_features]       NULL, or: instances of the node that carry the feature
 *        (== node_n unless row lengths were set); fmin / fmax range over those only */
int fr_dev_rf_level_stats(fr_dev_rf *rf, uint32_t n_active, uint32_t k, uint64_t *node_n, int64_t *node_sum,
                          float *gmin, float *gmax, float *fmin, float *fmax, uint32_t *b_n, uint32_t *b_pos,
                          int64_t *b_sum, int64_t *b_sq, uint32_t *f_present);
/* Per active node: fid (0xffffffff: the node is a leaf) and threshold; instances with
 * value < split move to node left[.], the others to right[.] (ids of the next level, -1 for a
 * child that is a leaf already). */
int fr_dev_rf_partition(fr_dev_rf *rf, uint32_t n_active, const uint32_t *fid, const double *split,
                        const int32_t *left, const int32_t *right);

/* NCCL bootstrap: rank 0 calls fr_dev_comm_unique_id, ships the 128 bytes to every rank by
 * any means (torch.distributed in fastrank_b200/dist.py), then all ranks call
 * fr_dev_comm_create.  Creation also maps every peer's mailbox through CUDA IPC so that
 * fr_dev_eval_coord_sweeps_fast can finish its cross-GPU sum inside the kernel over NVLink peer
 * memory; when that mapping is unavailable anywhere (or FASTRANK_P2P=0) every rank stays on the
 * ncclAllReduce path.  Either way sums are integers: results do not depend on the GPU count. */
int fr_dev_comm_unique_id(uint8_t out_id[128]);
int fr_dev_comm_create(int device, int rank, int world, const uint8_t id[128], fr_dev_comm **out);
void fr_dev_comm_destroy(fr_dev_comm *comm);
/* A number unique to this communicator for the life of the process (0 for NULL): what caches
 * keyed on a communicator compare, since a destroyed communicator's address can be reused. */
uint64_t fr_dev_comm_generation(const fr_dev_comm *comm);
/* Sum-all-reduce of n uint64 words held in host memory (used for setup-time counts). */
int fr_dev_comm_allreduce_u64(fr_dev_comm *comm, uint64_t *inout, size_t n);
/* Process-wide default: plans created through the reference-compatible surface
 * (train_model, evaluate_by_query, ...) attach this communicator, so that every rank of a
 * query-sharded job sees identical, all-reduced metric sums.  NULL clears it. */
void fr_dev_set_default_comm(fr_dev_comm *comm);
fr_dev_comm *fr_dev_default_comm(void);

/* Counters the bench reads: kernels launched by this library since load. */
uint64_t fr_dev_kernel_launches(void);

/* Device-side timing on the stream the kernels are launched on (bench.py cannot see that
 * stream through torch.cuda.Event).
 *   fr_dev_timer_start/stop  one CUDA-event pair around a region of calls;
 *   fr_dev_profile_enable    when on, every ranking kernel launch (coord sweep / linear batch /
 *                            scores eval) is bracketed by its own event pair;
 *   fr_dev_profile_read      number of bracketed launches and the sum of their durations. */
int fr_dev_timer_start(fr_dev_dataset *ds);
int fr_dev_timer_stop(fr_dev_dataset *ds, double *out_ms);
int fr_dev_profile_enable(fr_dev_dataset *ds, int on);
int fr_dev_profile_read(fr_dev_dataset *ds, uint64_t *out_launches, double *out_total_ms, int reset);

#ifdef __cplusplus
}
#endif
#endif /* FASTRANK_B200_H */
