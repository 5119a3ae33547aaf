/*
 * bias_cuda.h — C ABI of libbias_cuda, the B200 (sm_100a) collapsed-Gibbs engine that
 * replaces the component-assignment sweep of BIAS.jl's LDA / BMM / HDP samplers.
 *
 * The reference has no FFI: its boundary is the exported Julia method set
 * (src/BIAS.jl:8-14).  Each entry point below names the reference interface it
 * replaces (file:line under the reference tree).  A Julia shim `ccall`s these
 * (INTEGRATION.md); the tested caller is the ctypes binding in bias.jl_b200/_lib.py.
 *
 * Conventions
 *  - every function returns 0 on success or a BIAS_E* code; bias_last_error() gives the
 *    thread-local message.  No C++ exception crosses the ABI.
 *  - all buffers are caller-owned HOST memory unless a name ends in _dev.
 *  - indices are 0-based: assignments zz in [0, K), word ids in [0, V).  Ragged Julia
 *    arrays (Vector{Vector{T}}) arrive CSR-flattened: group_ptr[n_groups+1] offsets into
 *    the item arrays.
 *  - there is no CPU fallback: without a CUDA device every compute call fails with BIAS_ENODEV.
 */
#ifndef BIAS_CUDA_H
#define BIAS_CUDA_H
#include <stdint.h>

#if defined(__GNUC__)
#define BIAS_API __attribute__((visibility("default")))
#else
#define BIAS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BIAS_OK        0
#define BIAS_EINVAL    1   /* bad argument (the reference's ArgumentError from @check_args, src/common.jl:233-240) */
#define BIAS_ENODEV    2   /* no usable CUDA device */
#define BIAS_ECUDA     3   /* CUDA runtime error */
#define BIAS_ENOMEM    4
#define BIAS_ERANGE    5   /* n_jk beyond the Stirling table (reference: BoundsError, src/HDP.jl:355) or Kmax exhausted */
#define BIAS_ESTATE    6   /* e.g. delitem! on an empty Gaussian component (reference throws, src/conjugates.jl:66) */

/* conjugate family (src/conjugates.jl) and datum type (src/common.jl:12-18) */
#define BIAS_MD    0   /* MultinomialDirichlet, src/conjugates.jl:131-143 */
#define BIAS_G1G1  1   /* Gaussian1DGaussian1D, src/conjugates.jl:30-47 */
#define BIAS_INT   0   /* one word id per item */
#define BIAS_SENT  1   /* sparse count vector (Sent) per item */
#define BIAS_F64   2   /* one Float64 per item */

typedef struct bias_ctx bias_ctx;
typedef struct bias_model bias_model;    /* a device-resident LDA / BMM / HDP sampler state */

/* The prototype component handed to LDA(c, K, aa) / BMM(...) / HDP(...) */
typedef struct {
    int32_t kind;      /* BIAS_MD | BIAS_G1G1 */
    int32_t datum;     /* BIAS_INT | BIAS_SENT | BIAS_F64 */
    int64_t dd;        /* MD: cardinality V */
    double  aa;        /* MD: per-dimension concentration, i.e. the value the ctor stores (aa/dd, src/conjugates.jl:142) */
    double  m0, v0, vv;/* G1G1: prior mean / prior variance / likelihood variance */
} bias_conjugate;

/* The observations xx.  Exactly one family of pointers is used, chosen by conjugate.datum. */
typedef struct {
    int64_t n_items;
    const int32_t *word;        /* BIAS_INT  [n_items] */
    const double  *x;           /* BIAS_F64  [n_items] */
    const int64_t *sent_ptr;    /* BIAS_SENT [n_items+1] CSR offsets */
    const int32_t *sent_word;   /* BIAS_SENT word ids (unique within an item, as sparsify_sentence makes them) */
    const int32_t *sent_count;  /* BIAS_SENT counts */
} bias_data;

/* What the reference prints once per sweep (src/LDA.jl:108-109) plus engine counters. */
typedef struct {
    double  log_likelihood;   /* the reference's per-sweep diagnostic (posterior-mean sum for MD, src/conjugates.jl:205) */
    int64_t n_moved;          /* items whose assignment changed */
    int32_t KK;               /* components alive after the sweep (HDP); K otherwise */
    int32_t reserved;
    double  gg, aa;           /* HDP concentrations after the sweep */
} bias_sweep_stats;

/* HDP(c, KK, gg, g1, g2, aa, a1, a2), src/HDP.jl:15-33 — KK, gg, aa are updated in place like the Julia object */
typedef struct {
    int32_t KK;
    int32_t Kmax;             /* slot capacity (engine parameter; the reference grows nn by copying) */
    double  gg, g1, g2;
    double  aa, a1, a2;
    int32_t sample_hyperparam;/* src/HDP.jl:171 */
    int32_t n_internals;      /* src/HDP.jl:171 */
} bias_hdp_params;

/* RCRP(c, KK, aa, TT, a1, a2), src/RCRP.jl:18-28.  aa (one concentration per epoch) travels as a
 * separate array; KK is updated in place like the Julia object. */
typedef struct {
    int32_t KK;
    int32_t Kmax;             /* slot capacity (engine parameter) */
    double  a1, a2;
    int32_t sample_hyperparam;/* src/RCRP.jl:81; the sampler switches it off after the burn-in (:166-168) */
    int32_t n_internals;
} bias_rcrp_params;

/* ---- context ------------------------------------------------------------------------- */
BIAS_API int  bias_version(void);
BIAS_API const char *bias_last_error(void);
BIAS_API int  bias_device_count(int *n_out);
/* stream: a cudaStream_t to launch on (e.g. torch's current stream), or NULL to create one. */
BIAS_API int  bias_ctx_create(int device, uint64_t seed, void *stream, bias_ctx **out);
/* A context destroyed while models created on it are alive is released when the last of them is destroyed. */
BIAS_API int  bias_ctx_destroy(bias_ctx *ctx);
BIAS_API int  bias_ctx_sync(bias_ctx *ctx);
BIAS_API int  bias_ctx_num_sms(bias_ctx *ctx, int *n_out);
/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
BIAS_API int  bias_ctx_launch_count(bias_ctx *ctx, int64_t *n_out);

/* ---- one-call samplers: what the Julia shim ccalls ------------------------------------ */
/* collapsed_gibbs_sampler!(lda::LDA, xx, zz, n_burnins, n_lags, n_samples, ...)  src/LDA.jl:67-148.
 * n_sweeps = n_burnins + n_samples*(n_lags+1) (src/LDA.jl:82).  zz is mutated in place.
 * stats: NULL or n_sweeps entries. */
BIAS_API int bias_lda_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data,
                 int64_t n_groups, const int64_t *group_ptr, int32_t *zz,
                 int32_t K, double alpha, int32_t n_sweeps, bias_sweep_stats *stats);
/* collapsed_gibbs_sampler!(bmm::BMM, xx, zz, ...)  src/BMM.jl:46-130 */
BIAS_API int bias_bmm_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data,
                 int32_t *zz, int32_t K, double alpha, int32_t n_sweeps, bias_sweep_stats *stats);
/* collapsed_gibbs_sampler!(hdp::HDP, xx, zz, ...)  src/HDP.jl:166-399.  params (KK, gg, aa) are
 * updated like the Julia object; beta_out[Kmax+1] receives my_beta (src/HDP.jl:394); zz is
 * renumbered to 0..KK-1 in slot order.  If beta_out[0] > 0 on entry, beta_out[0..KK] is taken as the
 * initial my_beta (resuming a chain across calls); otherwise my_beta starts uniform (src/HDP.jl:228). */
BIAS_API int bias_hdp_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data,
                 int64_t n_groups, const int64_t *group_ptr, int32_t *zz,
                 bias_hdp_params *params, int32_t n_sweeps, double *beta_out, bias_sweep_stats *stats);

/* RCRP_gibbs_sampler!(rcrp::RCRP, xx, zz, n_burnins, n_lags, n_samples, sample_hyperparam, n_internals, ...)
 * src/RCRP.jl:77-416.  Epochs arrive as CSR groups (epoch_ptr[n_epochs+1]); aa[n_epochs] and params->KK are
 * updated in place; hyper-parameters are resampled only during the first n_burnins of the n_sweeps sweeps. */
BIAS_API int bias_rcrp_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data,
                  int64_t n_epochs, const int64_t *epoch_ptr, int32_t *zz,
                  bias_rcrp_params *params, double *aa, int32_t n_burnins, int32_t n_sweeps, bias_sweep_stats *stats);

/* ---- resident sampler state: same samplers, split into upload / sweep / download ------- */
#define BIAS_MODEL_LDA 0
#define BIAS_MODEL_BMM 1
#define BIAS_MODEL_HDP 2
#define BIAS_MODEL_RCRP 3
#define BIAS_MODEL_RCRF 5  /* recurrent Chinese restaurant franchise (src/RCRF.jl): created with bias_rcrf_create */
#define BIAS_MODEL_CRF 4   /* HDP sampled by CRF_gibbs_sampler! (src/HDP.jl:402-709): bias_model_create with hdp params */
/* Uploads xx and zz and runs the initialisation pass (src/LDA.jl:91-98, src/BMM.jl:70-75,
 * src/HDP.jl:217-228).  For BMM pass n_groups = 0, group_ptr = NULL; hdp = NULL unless model is HDP. */
BIAS_API int bias_model_create(bias_ctx *ctx, int32_t model, const bias_conjugate *conj, const bias_data *data,
                      int64_t n_groups, const int64_t *group_ptr, const int32_t *zz,
                      int32_t K, double alpha, const bias_hdp_params *hdp, bias_model **out);
/* Same for an RCRP model (epochs are the groups). */
BIAS_API int bias_rcrp_create(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data,
                     int64_t n_epochs, const int64_t *epoch_ptr, const int32_t *zz,
                     const bias_rcrp_params *params, const double *aa, bias_model **out);
BIAS_API int bias_model_get_rcrp(bias_model *m, bias_rcrp_params *params, double *aa /* [n_epochs] */);
/* Switch hyper-parameter resampling on/off between sweeps (HDP, RCRP). */
BIAS_API int bias_model_set_sample_hyperparam(bias_model *m, int32_t on);
/* Re-feeds a resident LDA / BMM model with a new batch (same conjugate, K, alpha): uploads xx and zz into the
 * buffers the model already owns (no allocation; the batch may be smaller than the one the model was created
 * with, BIAS_ERANGE if larger) and reruns the initialisation pass.  Streaming callers alternate two models on
 * two contexts so that one batch uploads while the other is swept. */
BIAS_API int bias_model_reload(bias_model *m, const bias_data *data, int64_t n_groups, const int64_t *group_ptr,
                      const int32_t *zz);
/* 16-bit host formats of the same two calls for LDA x word ids (dd <= 65536, KK <= 65536): word ids and assignments
 * cross PCIe as uint16 and are widened / narrowed on the device — half the bytes per step of a streaming caller
 * (the reference's own arrays are Int64, src/LDA.jl:67-72: the shim converts either way). */
BIAS_API int bias_model_reload_u16(bias_model *m, int64_t n_items, const uint16_t *word, int64_t n_groups,
                          const int64_t *group_ptr, const uint16_t *zz);
BIAS_API int bias_model_get_zz_u16(bias_model *m, uint16_t *zz);
BIAS_API int bias_model_destroy(bias_model *m);
/* n_sweeps asynchronous sweeps on the context stream (one iteration of the loop at src/LDA.jl:103). */
BIAS_API int bias_model_sweep(bias_model *m, int32_t n_sweeps);
/* Re-upload assignments (and rebuild counts) without re-uploading xx. */
BIAS_API int bias_model_set_zz(bias_model *m, const int32_t *zz);
BIAS_API int bias_model_get_zz(bias_model *m, int32_t *zz);
BIAS_API int bias_model_last_stats(bias_model *m, bias_sweep_stats *out);
/* posterior(model, xx, zz) support (src/LDA.jl:151-167): component sufficient statistics.
 * MD:   mi[K*dd] topic-major int64 (the reference's per-topic mi), mm[K], nn[K].
 * G1G1: mu[K] (mean of the assigned data), nn[K].  Unused outputs may be NULL. */
BIAS_API int bias_model_get_components(bias_model *m, int64_t *mi, int64_t *mm, double *mu, int64_t *nn);
/* nn[n_groups*K] row-major group-by-component counts (the reference's nn matrix). */
BIAS_API int bias_model_get_group_counts(bias_model *m, int64_t *nn);
BIAS_API int bias_model_get_hdp(bias_model *m, bias_hdp_params *params, double *beta /* [Kmax+1] */);

/* ---- verification entry points (parity levels (a) and (b) of the north star) ----------- */
/* Conditionals of n_query items on the CURRENT device state with the item itself removed
 * (LDA_sample_zz, src/LDA.jl:26-42; src/BMM.jl:102-107; src/HDP.jl:312-317), evaluated in
 * Float64 with the reference's operation order.  n_out = K (K+1 for HDP).
 * raw[n_query*n_out]: unnormalised log scores; prob[...]: after lognormalize!;
 * draw[n_query]: result of `sample` for the supplied uniforms.  State is not modified. */
BIAS_API int bias_model_conditionals(bias_model *m, int64_t n_query, const int64_t *item_idx,
                            const double *uniforms, double *raw, double *prob, int32_t *draw);
/* One pass of the PRODUCTION sweep kernel with counts frozen and uniforms supplied
 * (uniforms[n_items]); draws[n_items] out.  State is not modified. */
BIAS_API int bias_model_frozen_draws(bias_model *m, const double *uniforms, int32_t *draws);
/* The same with uniforms and draws in DEVICE memory ([n_items] each; synchronises before returning). */
BIAS_API int bias_model_frozen_draws_dev(bias_model *m, const double *uniforms_dev, int32_t *draws_dev);
/* One PRODUCTION sweep (counts updated, zz rewritten) that draws with the caller's uniforms[n_items] instead of the
 * Philox stream: a one-document corpus is then swept serially, token by token, exactly as src/LDA.jl:116-131 does,
 * and can be followed by the oracle (tests/test_gpu_headline.py).  LDA and BMM, single GPU. */
BIAS_API int bias_model_sweep_with_uniforms(bias_model *m, const double *uniforms);
/* HDP table-count conditional (src/HDP.jl:353-358) for n_query (n, a = aa*beta_k) pairs:
 * prob[q*n_max + m-1] for m = 1..n, draws[q] in 1..n. */
BIAS_API int bias_hdp_table_conditionals(bias_ctx *ctx, int64_t n_query, const int32_t *n, const double *a,
                                const double *uniforms, int32_t n_max, double *prob, int32_t *draws);
/* Device-resident normalised log-Stirling table (regenerates StirlingNums_10K.mat, src/HDP.jl:207-209).
 * Copies rows 1..n_rows (packed lower triangle) to out; builds the table on first use. */
BIAS_API int bias_stirling_rows(bias_ctx *ctx, int32_t n_rows, double *out);
/* sample_hyperparam!(hdp, n_group_j, m), src/HDP.jl:124-159, in its two parts.  Device: out[0] = sum_j log w_j with
 * w_j ~ Beta(aa + 1, n_j) and out[1] = sum_j s_j with s_j ~ Bernoulli(n_j / (n_j + aa)) (:133-144) for the context's
 * Philox stream at (sweep, internal); with uniforms[n_groups] supplied, s_j = uniforms[j] < n_j / (n_j + aa). */
BIAS_API int bias_hdp_hyper_aux(bias_ctx *ctx, int64_t n_groups, const int64_t *group_ptr, double aa, uint32_t sweep,
                       uint32_t internal, const double *uniforms /* nullable */, double *out /* [2] */);
/* Host: n independent draws of (alpha0, gamma) (:146-158) given the auxiliaries, each starting from (aa, gg). */
BIAS_API int bias_hdp_hyper_draws(uint64_t seed, double a1, double a2, double g1, double g2, double n_tables, double sum_logw,
                         double sum_s, int32_t KK, double aa, double gg, int32_t n, double *aa_out, double *gg_out);
/* Host: n draws of my_beta ~ Dirichlet([sum_j M_jk ..., gg]), src/HDP.jl:362-364 -> beta_out[n][KK + 1]. */
BIAS_API int bias_hdp_beta_draws(uint64_t seed, int32_t KK, const int64_t *tables, double gg, int32_t n, double *beta_out);

/* Sum over items of log p(x_i | theta_group, phi) under the point estimates of the current state,
 * theta_jk = (n_jk + aa)/(n_j + K aa), phi = posterior mean of each component (LDA x MD x Int).
 * The reference has no perplexity function (src/notes.md:30); perplexity = exp(-sum / n_items).
 * On a sharded run every rank reports its own shard's sum. */
BIAS_API int bias_model_log_likelihood(bias_model *m, double *sum_log_p, int64_t *n_items);
/* Document completion: the same sum over HELD-OUT tokens (CSR over the model's groups: heldout_ptr[n_groups+1],
 * heldout_word[heldout_ptr[n_groups]], host buffers); theta_j comes from the group's training tokens, phi from the
 * training counts.  Held-out perplexity = exp(-sum / n_tokens) (north_star level (c)). */
BIAS_API int bias_model_heldout_log_likelihood(bias_model *m, const int64_t *heldout_ptr, const int32_t *heldout_word,
                      double *sum_log_p, int64_t *n_tokens);

/* ---- CRF (model BIAS_MODEL_CRF) --------------------------------------------------------- */
/* One iteration with caller-supplied uniforms (verification): per item slot i [3i] table draw, [3i+1] dish of a new table; per table slot [3(group_ptr[j]+t)+2] dish draw. */
BIAS_API int bias_crf_sweep_with_uniforms(bias_model *m, const double *uniforms /* [3 * n_items] */);
/* The seating: tji[n_items] table of each customer within its group, njt / kjt [n_items] customers and dish per table
 * (tables of group j at group_ptr[j] ...), n_tbls[n_groups], mm[KK] tables per dish.  Any pointer may be NULL. */
BIAS_API int bias_crf_get_tables(bias_model *m, int32_t *tji, int32_t *njt, int32_t *kjt, int32_t *n_tbls, int32_t *mm);

/* ---- RCRF (src/RCRF.jl:122-938; Float64, word-id and Sent items) -------------------------- */
/* RCRF(c, KK, gg, g1, g2, aa, a1, a2, TT), src/RCRF.jl:16-37; gg and aa (one per epoch) travel as separate arrays. */
typedef struct {
    int32_t KK;
    int32_t Kmax;             /* dish-slot capacity (engine parameter) */
    double  g1, g2, a1, a2;
    int32_t sample_hyperparam;
    int32_t n_internals;
    int32_t join_tables;      /* src/RCRF.jl:130 */
} bias_rcrf_params;
/* xx::Vector{Vector{Vector{T}}} arrives flattened: the restaurants of all epochs are the groups (group_ptr), epoch_ptr
 * [n_epochs + 1] holds the first restaurant of each epoch.  One bias_model_sweep = one iteration of
 * RCRF_gibbs_sampler! (:217-905): per epoch, tables for customers, joining of same-dish tables, dishes for tables, the
 * epoch's hyper-parameters.  bias_crf_get_tables returns mm as [n_epochs][KK]. */
BIAS_API int bias_rcrf_create(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int64_t n_groups,
                     const int64_t *group_ptr, int64_t n_epochs, const int64_t *epoch_ptr, const int32_t *zz,
                     const bias_rcrf_params *params, const double *gg, const double *aa, bias_model **out);
BIAS_API int bias_model_get_rcrf(bias_model *m, int32_t *KK, double *gg /* [n_epochs] */, double *aa /* [n_epochs] */);

/* ---- sharded HDP (groups split over ranks) ---------------------------------- */
/* One sweep of collapsed_gibbs_sampler!(::HDP) (src/HDP.jl:235-373) cut at the points where ranks must agree:
 *   bias_model_delta_begin
 *   bias_hdp_shard_pass      parallel assignment pass; items that draw the new-component slot are queued
 *   for r = 0 .. world-1:    the ranks take turns giving birth, so that a rank sees the components the ranks before
 *                            it opened in this sweep (two ranks meeting the same new atom do not both open one)
 *       rank r:       bias_hdp_shard_births  -> its KK and beta (stick-breaking of src/HDP.jl:325-328)
 *       other ranks:  [broadcast KK, beta from r]  bias_hdp_shard_adopt
 *       all ranks:    bias_model_delta_pack / [all-reduce] / bias_model_delta_apply
 *   bias_hdp_shard_prune     dead components (the same on every rank: counts are global now)
 *   n_internals x { bias_hdp_shard_tables  [all-reduce tables (int64[n_tables]) and hyper (double[2])]  bias_hdp_shard_draw }
 *   bias_hdp_shard_end
 * beta, alpha0 and gamma are drawn on every rank from the same host RNG stream (bias_model_set_host_seed with one
 * seed for all ranks), so replicas stay identical.  bias.jl_b200/sharding.py drives this protocol. */
BIAS_API int bias_model_set_host_seed(bias_model *m, uint64_t seed);
BIAS_API int bias_hdp_shard_pass(bias_model *m);
BIAS_API int bias_hdp_shard_births(bias_model *m, int32_t *KK_out, double *beta_out /* [Kmax+1] */);
BIAS_API int bias_hdp_shard_adopt(bias_model *m, int32_t KK, const double *beta /* [KK+1] */);
BIAS_API int bias_hdp_shard_prune(bias_model *m);
BIAS_API int bias_hdp_shard_tables(bias_model *m, int32_t internal, void **tables_dev, int64_t *n_tables, void **hyper_dev);
BIAS_API int bias_hdp_shard_draw(bias_model *m);
BIAS_API int bias_hdp_shard_end(bias_model *m);

/* ---- multi-GPU exchange (AD-LDA: one allreduce of count deltas per sweep) --------------- */
/* Snapshot the shared component statistics (n_wk, n_k / sumx, n) as the base of the next delta. */
BIAS_API int bias_model_delta_begin(bias_model *m);
/* Base = 0: the next pack/all-reduce/apply merges the ranks' LOCAL initial counts into the global ones
 * (call once after bias_model_create on every rank). */
BIAS_API int bias_model_delta_begin_zero(bias_model *m);
/* Compute delta = live - base into the exchange buffer and return its DEVICE address:
 * n_i32 int32 values followed (8-byte aligned) by n_f64 doubles.  The caller all-reduces
 * (sum) both segments in place across ranks, e.g. with NCCL on the same stream. */
BIAS_API int bias_model_delta_pack(bias_model *m, void **delta_i32_dev, int64_t *n_i32,
                          void **delta_f64_dev, int64_t *n_f64);
/* live = base + (all-reduced) delta. */
BIAS_API int bias_model_delta_apply(bias_model *m);
/* The same exchange with the word-topic deltas in 16 bits (half the all-reduce bytes; LDA x MD x Int).  pack16 writes
 * 16-bit deltas, two per int32 as d1 * 65536 + d0 (NCCL has no int16 sum; all-reduce the buffer as n16 / 2 int32 —
 * exact while the summed deltas fit 16 bits), and a small int32 buffer (n_k / n_items deltas followed by 4 ints whose first is
 * an overflow flag: non-zero after the SUM all-reduce means some |delta| * world exceeded int16 on some rank).
 * All-reduce both buffers, then apply16: it checks the flag ON THE DEVICE and applies nothing when it is set — the
 * deltas stay pending (base is not advanced) and ride along with the next exchange, which the caller makes a 32-bit
 * one (bias_model_delta_pack / apply) once it has seen the flag; no host synchronisation is needed in between. */
BIAS_API int bias_model_delta_pack16(bias_model *m, int32_t world, void **d16_dev, int64_t *n16, void **small_dev, int64_t *n_small);
BIAS_API int bias_model_delta_apply16(bias_model *m);

/* ---- sharded LDA inside the library (csrc/comm.cu): no external collective ------------------------------------
 * The reference is single-process; these replace nothing in it and are what lets its entry point
 * collapsed_gibbs_sampler!(lda, xx, zz, ...) (src/LDA.jl:67-72) use every GPU of a box.  Documents are split over the
 * ranks, every rank keeps a replica of the word-topic counts and once per sweep the replicas are reconciled by the
 * library's own reduce-scatter / all-gather kernel over NVLink peer memory.  LDA x MultinomialDirichlet x word ids,
 * KK <= 1024, documents shorter than 65,536 tokens, at most 8 ranks (one NVSwitch domain). */

/* (1) One process, several GPUs — what the Julia shim calls.  Splits the documents contiguously over `devices`, runs
 * n_sweeps sweeps, returns the assignments in zz (caller-owned, mutated in place like the reference's zz). */
BIAS_API int bias_lda_run_multi(const int32_t *devices, int32_t n_dev, uint64_t seed, const bias_conjugate *conj,
                       const bias_data *data, int64_t n_groups, const int64_t *group_ptr, int32_t *zz /* in/out */,
                       int32_t K, double alpha, int32_t n_sweeps, bias_sweep_stats *stats /* [n_sweeps] or NULL */);
/* Where the most recent bias_lda_run_multi of the calling thread spent its time, in milliseconds of host wall clock:
 * ms[0] set-up (contexts, shard upload, initialisation pass, peer mappings), ms[1] the sweeps with their exchanges,
 * ms[2] download and teardown.  (The reference prints the elapsed time of every iteration, src/LDA.jl:110.) */
BIAS_API int bias_lda_run_multi_timing(double *ms /* [3] */);
/* The same with resident shards: models created by the caller (one per device, or several on one device) become the
 * ranks of a group; bias_group_sweep = sweep of every shard + the exchange, ordered by CUDA events.  After a reload
 * of the shards (all of them) the next bias_group_sweep first merges their local counts.  The group does not own
 * the models. */
typedef struct bias_group bias_group;
BIAS_API int bias_group_create(bias_model **models, int32_t n, bias_group **out);
BIAS_API int bias_group_sweep(bias_group *g, int32_t n_sweeps);
BIAS_API int bias_group_destroy(bias_group *g);

/* (2) One process per GPU (torchrun / MPI / one Julia worker per GPU).  Every rank: create its model on its shard,
 * bias_model_share_prepare(rank, world, sum over ranks of the item capacities), bias_model_share_export, exchange
 * the handles by any host-side means (they are plain bytes), bias_model_share_attach.  From then on
 * bias_model_sweep is COLLECTIVE: every rank must call it with the same n_sweeps; it runs the local sweep and the
 * exchange (ranks are ordered by arrival counters in peer memory; a rank that waits more than 60 s for a peer
 * gives up and the next bias_model_get_components / bias_model_replica_digest reports BIAS_ESTATE). */
typedef struct {
    unsigned char ipc[64];      /* cudaIpcMemHandle_t of the rank's shared block */
    int64_t bytes, dd;
    int32_t Kpad, hot_cap, rank, device;
} bias_share_handle;
BIAS_API int bias_model_share_prepare(bias_model *m, int32_t rank, int32_t world, int64_t n_items_cap_global);
BIAS_API int bias_model_share_export(bias_model *m, bias_share_handle *out);
BIAS_API int bias_model_share_attach(bias_model *m, const bias_share_handle *all /* [world] */, int32_t world);
/* Device time of the most recent exchange kernel of this rank (one process per GPU; measured between the two barriers,
 * i.e. with all ranks present) and the bytes it pulled from its peers (as many go out). */
BIAS_API int bias_model_last_exchange(bias_model *m, float *ms_out, int64_t *bytes_in_out);
/* 64-bit digest of the replica's counts (word-topic table and n_k): equal on all ranks after an exchange. */
BIAS_API int bias_model_replica_digest(bias_model *m, uint64_t *digest);

/* Which representation of the word-topic counts is current on the fast LDA path: *narrow_current = 1 when the sweeps run
 * on the 16-bit rows (+ int32 rows for the *n_hot words whose corpus frequency exceeds 65,535; room for *hot_cap). */
BIAS_API int bias_model_narrow_info(bias_model *m, int32_t *narrow_current, int32_t *n_hot, int32_t *hot_cap);

/* ---- instrumentation ------------------------------------------------------------------ */
/* Device time of the most recent bias_model_sweep call's sweep kernels, measured with CUDA
 * events on the context stream: total ms and launches. */
BIAS_API int bias_model_last_sweep_ms(bias_model *m, float *ms_out, int32_t *launches_out);

#ifdef __cplusplus
}
#endif
#endif
