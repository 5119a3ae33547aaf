/*
 * bias_oracle.h — CPU oracle for the BIAS.jl collapsed-Gibbs assignment sweep.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * serial Julia arithmetic (file:line citations are relative to /root/reference).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product path (libbias_cuda) never does.
 *
 * Pinning status: the conjugate / distribution arithmetic is pinned against the
 * reference's documented REPL transcripts (tests/golden/reference_known_answers.json,
 * SURVEY.md §4).  Nothing in the reference pins sampler draws, sweep trajectories,
 * the Stirling table step or any Distributions.jl draw (no Julia here, no tests
 * upstream): for those this oracle is "parity unpinned" — a restatement only.
 *
 * Conventions: indices are 0-based (the reference is 1-based); all counts are
 * int64_t and all reals double, as in the reference (Int64 / Float64).
 */
#ifndef BIAS_ORACLE_H
#define BIAS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/common.jl ---------------------------------------------------- */
double orc_julia_sum(const double *a, int64_t n);                 /* Base.sum (julia 0.4 reduce.jl, blksize 1024) */
void orc_lognormalize(double *pp, int64_t n);                      /* src/common.jl:42-52 */
int64_t orc_sample(const double *w, int64_t n, double r);          /* src/common.jl:69-79, 0-based result */
void orc_gen_bars(int64_t n_bars, int64_t n_vocab, double noise, double *bars /* [n_bars*n_vocab] row-major */); /* src/common.jl:157-175 */

/* ---- src/distributions.jl -------------------------------------------- */
double orc_gaussian1d_pdf(double mu, double vv, double x);         /* src/distributions.jl:36 */
double orc_gaussian1d_logpdf(double mu, double vv, double x);      /* src/distributions.jl:37 */

/* ---- src/conjugates.jl ------------------------------------------------ */
enum { ORC_MD = 0, ORC_G1G1 = 1 };             /* conjugate kind */
enum { ORC_INT = 0, ORC_SENT = 1, ORC_F64 = 2 }; /* datum kind */

/* K components of one conjugate family, struct-of-arrays. */
typedef struct {
    int32_t kind;      /* ORC_MD | ORC_G1G1 */
    int64_t K;         /* capacity (number of component slots) */
    /* MultinomialDirichlet (src/conjugates.jl:131-143) */
    int64_t dd;        /* cardinality V */
    double  aa;        /* per-dimension concentration (ctor stores aa/dd) */
    int64_t *mi;       /* [K*dd] topic-major, like the reference's per-topic mi */
    int64_t *mm;       /* [K] */
    /* Gaussian1DGaussian1D (src/conjugates.jl:30-47) */
    double m0, v0, vv;
    double *mu;        /* [K] running mean */
    /* both */
    int64_t *nn;       /* [K] number of data items */
} orc_comps;

/* A data set of N items of one datum kind. */
typedef struct {
    int32_t kind;           /* ORC_INT | ORC_SENT | ORC_F64 */
    const int64_t *w;       /* ORC_INT: word id per item (0-based) */
    const double  *x;       /* ORC_F64: value per item */
    const int64_t *sptr;    /* ORC_SENT: CSR offsets [N+1] */
    const int64_t *sw;      /* ORC_SENT: word ids */
    const int64_t *sc;      /* ORC_SENT: counts */
} orc_data;

void   orc_additem(orc_comps *c, int64_t k, const orc_data *d, int64_t i);     /* :60-63, :153-157, :163-170 */
int    orc_delitem(orc_comps *c, int64_t k, const orc_data *d, int64_t i);     /* :64-74 (returns -1 where the reference throws), :158-162, :171-178 */
double orc_logpredictive(const orc_comps *c, int64_t k, const orc_data *d, int64_t i); /* :92-103, :180, :181-198 */
double orc_loglikelihood(const orc_comps *c, int64_t k, const orc_data *d, int64_t i); /* :105, :205-213 */
void   orc_g1g1_posterior(const orc_comps *c, int64_t k, double *mu, double *vv);      /* :77-89 */
void   orc_comps_reset(orc_comps *c, int64_t k);   /* deepcopy(model.component) into slot k */
void   orc_comps_move(orc_comps *c, int64_t dst, int64_t src); /* splice! support */

/* ---- src/LDA.jl ------------------------------------------------------- */
/* Initialisation pass (src/LDA.jl:91-98). nn is [D*K] row-major. Returns the log_likelihood diagnostic. */
double orc_lda_init(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                    const int64_t *z, int64_t *nn, int with_diag);
/* One sweep (src/LDA.jl:113-133). doc_order[D] / tok_order[N] give the randperm
 * orders (global item indices, grouped by document in visiting order); NULL = identity.
 * uniforms[N] is indexed by global item index. Returns the log_likelihood diagnostic. */
double orc_lda_sweep(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                     int64_t *z, int64_t *nn, double alpha,
                     const int64_t *doc_order, const int64_t *tok_order,
                     const double *uniforms, int faithful_diag);
/* Test support: the serial sweep with the conditional in the linear domain (word-id items; see bias_oracle.c). */
void orc_lda_sweep_linear(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr, int64_t *z, int64_t *nn,
                          double alpha, const int64_t *doc_order, const int64_t *tok_order, const double *uniforms);
/* Test support: the same serial sweep following another sampler's assignments other_z[N] (see bias_oracle.c). */
int64_t orc_lda_sweep_follow(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                             int64_t *z, int64_t *nn, double alpha, const double *uniforms,
                             const int64_t *other_z, double *max_edge_dist);
/* Conditional of one item on the current state with the item removed (LDA_sample_zz, src/LDA.jl:26-42).
 * raw[K] <- unnormalised log scores, prob[K] <- after lognormalize!. Returns the 0-based draw for r.
 * State is restored on return. */
int64_t orc_lda_conditional(orc_comps *c, const orc_data *d, int64_t j, int64_t i,
                            int64_t zi, int64_t *nn, double alpha, double r,
                            double *raw, double *prob);

/* ---- src/BMM.jl ------------------------------------------------------- */
double orc_bmm_init(orc_comps *c, const orc_data *d, int64_t N, const int64_t *z, int64_t *nn); /* :70-75 */
double orc_bmm_sweep(orc_comps *c, const orc_data *d, int64_t N, int64_t *z, int64_t *nn,
                     double alpha, const int64_t *order, const double *uniforms);               /* :92-115 */
int64_t orc_bmm_conditional(orc_comps *c, const orc_data *d, int64_t i, int64_t zi, int64_t *nn,
                            double alpha, double r, double *raw, double *prob);

/* ---- src/HDP.jl (direct assignment sampler, :166-399) ------------------ */
typedef struct { uint64_t s[4]; } orc_rng;
void   orc_rng_seed(orc_rng *g, uint64_t seed);
double orc_rand(orc_rng *g);                          /* U[0,1) */
double orc_randn(orc_rng *g);
double orc_rand_gamma(orc_rng *g, double shape);      /* Gamma(shape, 1); Marsaglia-Tsang */
double orc_rand_beta(orc_rng *g, double a, double b);
void   orc_randperm(orc_rng *g, int64_t n, int64_t *out);

typedef struct {
    int64_t KK;        /* active components (hdp.KK) */
    int64_t Kmax;      /* capacity of comps / nn columns / beta */
    double gg, g1, g2; /* top-level concentration and its Gamma prior */
    double aa, a1, a2; /* group-level concentration and its Gamma prior */
    double *beta;      /* [Kmax+1]; beta[KK] is the unrepresented mass */
} orc_hdp;

double orc_hdp_init(orc_hdp *h, orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                    const int64_t *z, int64_t *nn /* [D*Kmax] row-major */);              /* :217-228 */
/* prune (:248-264) + item loop (:279-341). uniforms may be NULL (then rng is used for item draws).
 * Returns the log_likelihood diagnostic, or NaN if Kmax would be exceeded. */
double orc_hdp_sweep(orc_hdp *h, orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                     int64_t *z, int64_t *nn, orc_rng *g, const double *uniforms, int shuffle);
/* K+1 conditional of one item with the item removed (:312-317); no birth/death handling.
 * raw/prob have KK+1 entries. */
int64_t orc_hdp_conditional(const orc_hdp *h, orc_comps *c, const orc_data *d, int64_t j, int64_t i,
                            int64_t zi, int64_t *nn, double r, double *raw, double *prob);
/* Table-count conditional for one (n, a=aa*beta_k) (:353-358). lstir is the normalised
 * log-Stirling table (packed lower triangle, see orc_stirling_table). prob[n] out; returns draw m (1-based count). */
int64_t orc_hdp_table_conditional(const double *lstir, int64_t n, double a, double r, double *prob);
/* n_internals x { table counts (:348-361), beta ~ Dirichlet (:362-364), hyper-parameters (:369-372 -> :124-159) }.
 * M_out[D*Kmax] receives the last table counts. Returns 0, or -1 if some n_jk exceeds nmax. */
int orc_hdp_resample_beta(orc_hdp *h, int64_t D, const int64_t *ptr, const int64_t *nn,
                          const double *lstir, int64_t nmax, int n_internals, int sample_hyper,
                          orc_rng *g, int64_t *M_out);
void orc_hdp_sample_hyperparam(orc_hdp *h, int64_t D, const int64_t *ptr, int64_t m, orc_rng *g); /* :124-159 */

/* ---- src/RCRP.jl (recurrent Chinese restaurant process, :77-416) ---------------------------
 * Epochs are the groups: ptr[TT+1] offsets, nn[TT*Kmax] row-major epoch-by-component counts,
 * aa[TT] per-epoch concentration.  Components are shared by all epochs. */
typedef struct {
    int64_t KK, Kmax, TT;
    double *aa;        /* [TT] */
    double a1, a2;
} orc_rcrp;
void orc_rcrp_init(orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, const int64_t *z, int64_t *nn); /* :98-143 */
/* One iteration of the MCMC loop (:174-398): first / middle / last epoch blocks, deaths, births,
 * chain splits (:309-331), optional hyper-parameter resampling (:60-74).  uniforms may be NULL.
 * Returns the log_likelihood diagnostic of the last epoch block structure (sum over all epochs), NaN on capacity. */
double orc_rcrp_sweep(orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, int64_t *z, int64_t *nn,
                      orc_rng *g, const double *uniforms, int shuffle, int sample_hyper, int n_internals);
/* Conditional of item i of epoch t on the current state with the item removed (:191-215, :263-287, :366-377).
 * raw/prob have KK+1 entries; components that are not alive in the window get -Inf / 0.  No deletion is
 * performed when the removal empties a component (it simply has probability 0).  State restored. */
int64_t orc_rcrp_conditional(const orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, int64_t t,
                             int64_t i, int64_t zi, int64_t *nn, double r, double *raw, double *prob);

/* ---- src/HDP.jl:402-709 CRF_gibbs_sampler! (Chinese restaurant franchise) -----------------------
 * Tables live in CSR arrays with the items' offsets (a group has at most as many tables as items):
 * tji[N] table of each item (0-based within its group), njt[N] customers per table, kjt[N] dish per table,
 * n_tbls[G], mm[Kmax] tables per dish. */
typedef struct {
    int64_t G;
    const int64_t *ptr;
    int64_t *tji, *njt, *kjt, *n_tbls, *mm;
} orc_crf;
void orc_crf_init(const orc_hdp *h, orc_crf *f, orc_comps *c, const orc_data *d, const int64_t *z);     /* :441-470 */
/* One iteration (:506-667): tables for customers, dishes for tables, hyper-parameters.  uniforms (nullable) holds three
 * numbers per item slot: [3i] table draw and [3i+1] dish-of-a-new-table draw of item i, [3(ptr[j]+t)+2] dish draw of
 * table t of group j.  defer_deaths = 0 is the reference (a dish that loses its last table is removed at once and
 * everything above it renumbered); 1 keeps the slot (its weight is log 0) until the end of the iteration — the order
 * the GPU engine uses.  Returns the log_likelihood diagnostic (:669-673), NaN on capacity. */
double orc_crf_sweep(orc_hdp *h, orc_crf *f, orc_comps *c, const orc_data *d, int64_t *z, orc_rng *g,
                     const double *uniforms, int defer_deaths, int sample_hyper, int n_internals);
int64_t orc_crf_table_conditional(const orc_hdp *h, const orc_crf *f, orc_comps *c, const orc_data *d, int64_t j,
                                  int64_t i, int64_t *z, double r, double *raw, double *prob);

/* ---- src/RCRF.jl (recurrent Chinese restaurant franchise, :122-938) ------------------------------
 * Restaurants of all epochs are flattened: eptr[TT+1] = first restaurant of each epoch, ptr[G+1] = first customer of
 * each restaurant; tables in CSR arrays as for the CRF; mm[TT*Kmax] row-major tables per (epoch, dish). */
typedef struct {
    int64_t KK, Kmax, TT;
    const int64_t *eptr, *ptr;
    int64_t *tji, *njt, *kjt, *n_tbls, *mm;
    double *gg, *aa;   /* [TT] */
    double g1, g2, a1, a2;
} orc_rcrf;
void orc_rcrf_init(orc_rcrf *f, orc_comps *c, const orc_data *d, const int64_t *z);                       /* :177-209 */
/* One iteration (:217-905): per epoch, tables for customers, joining of same-dish tables, dishes for tables, and the
 * epoch's hyper-parameters.  uniforms / defer_deaths as for orc_crf_sweep.  NaN on capacity. */
double orc_rcrf_sweep(orc_rcrf *f, orc_comps *c, const orc_data *d, int64_t *z, orc_rng *g, const double *uniforms,
                      int defer_deaths, int join_tables, int sample_hyper, int n_internals);

/* Normalised log unsigned Stirling numbers of the first kind, rows n = 1..nmax, packed:
 * entry (n, m), 1 <= m <= n, at index n*(n-1)/2 + (m-1); each row shifted so its maximum is 0.
 * Regenerates the missing src/StirlingNums_10K.mat (snumbersNormalizedSparse, src/HDP.jl:207-209). */
void orc_stirling_table(int64_t nmax, double *out);

#ifdef __cplusplus
}
#endif
#endif
