/*
 * bias_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see bias_oracle.h).
 *
 * A from-scratch C restatement of the serial Julia arithmetic on BIAS.jl's
 * component-assignment hot path.  Citations are file:line under /root/reference.
 * The product library (libbias_cuda) never links or loads this file.
 */
#include "bias_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================
 * src/common.jl
 * ====================================================================== */

/* Julia 0.4 Base.sum over a dense Float64 array: mapreduce_impl with blksize
 * 1024 — a left-to-right loop below the block size, pairwise halves above it
 * (Julia base/reduce.jl; third-party, unpinned beyond `REQUIRE: julia 0.4`).
 * Used wherever the reference calls sum(::Vector{Float64}) (src/common.jl:48,
 * src/conjugates.jl:193-195, src/distributions.jl:62). */
static double jsum_rec(const double *a, int64_t lo, int64_t hi) /* inclusive */
{
    if (lo == hi) return a[lo];
    if (lo + 1024 > hi) {
        double v = a[lo] + a[lo + 1];
        for (int64_t i = lo + 2; i <= hi; ++i) v += a[i];
        return v;
    }
    int64_t mid = (int64_t)(((uint64_t)lo + (uint64_t)hi) >> 1);
    double v1 = jsum_rec(a, lo, mid);
    double v2 = jsum_rec(a, mid + 1, hi);
    return v1 + v2;
}

double orc_julia_sum(const double *a, int64_t n)
{
    if (n <= 0) return 0.0;
    return jsum_rec(a, 0, n - 1);
}

/* src/common.jl:42-52 — max-shift, exp, sum, divide (four passes). */
void orc_lognormalize(double *pp, int64_t n)
{
    double pp_max = pp[0];
    for (int64_t k = 1; k < n; ++k) if (pp[k] > pp_max) pp_max = pp[k];
    for (int64_t k = 0; k < n; ++k) pp[k] = exp(pp[k] - pp_max);
    double pp_sum = orc_julia_sum(pp, n);
    for (int64_t k = 0; k < n; ++k) pp[k] /= pp_sum;
}

/* src/common.jl:69-79 — linear CDF walk with the strict test `cw < r`,
 * clamped at the last entry.  r is supplied (the reference calls rand()). */
int64_t orc_sample(const double *w, int64_t n, double r)
{
    int64_t i = 0;
    double cw = w[0];
    while (cw < r && i < n - 1) {
        i += 1;
        cw += w[i];
    }
    return i;
}

/* src/common.jl:157-175 — bar topics on a sqrt(V) x sqrt(V) grid, column-major
 * flatten (Julia `b[:]`).  First half of the bars are grid rows, second half
 * grid columns. */
void orc_gen_bars(int64_t n_bars, int64_t n_vocab, double noise, double *bars)
{
    int64_t S = (int64_t)llround(sqrt((double)n_vocab));
    int64_t half = (int64_t)llround((double)n_bars / 2.0);
    /* Julia round(): half away from zero, same as llround */
    double *b = (double *)malloc((size_t)(S * S) * sizeof(double));
    for (int64_t i = 0; i < n_bars * n_vocab; ++i) bars[i] = 0.0;
    for (int64_t kk = 0; kk < n_bars; ++kk) {
        for (int64_t i = 0; i < S * S; ++i) b[i] = 0.0 + noise;
        if (kk < half) {
            for (int64_t c = 0; c < S; ++c) b[kk + c * S] = 1.0;          /* b[kk, :] */
        } else {
            int64_t col = kk - half;
            for (int64_t r = 0; r < S; ++r) b[r + col * S] = 1.0;         /* b[:, kk-half] */
        }
        double s = orc_julia_sum(b, S * S);
        for (int64_t i = 0; i < S * S && i < n_vocab; ++i) bars[kk * n_vocab + i] = b[i] / s;
    }
    free(b);
}

/* ======================================================================
 * src/distributions.jl
 * ====================================================================== */
double orc_gaussian1d_pdf(double mu, double vv, double x)
{
    return exp(-(x - mu) * (x - mu) / (2 * vv)) / (sqrt(2 * M_PI * vv));   /* :36 */
}
double orc_gaussian1d_logpdf(double mu, double vv, double x)
{
    return log(orc_gaussian1d_pdf(mu, vv, x));                             /* :37 */
}

/* ======================================================================
 * src/conjugates.jl
 * ====================================================================== */
void orc_comps_reset(orc_comps *c, int64_t k)
{
    c->nn[k] = 0;
    if (c->kind == ORC_MD) {
        c->mm[k] = 0;
        memset(c->mi + k * c->dd, 0, (size_t)c->dd * sizeof(int64_t));
    } else {
        c->mu[k] = c->m0;                     /* ctor sets mu = m0, src/conjugates.jl:42-47 */
    }
}

void orc_comps_move(orc_comps *c, int64_t dst, int64_t src)
{
    c->nn[dst] = c->nn[src];
    if (c->kind == ORC_MD) {
        c->mm[dst] = c->mm[src];
        memmove(c->mi + dst * c->dd, c->mi + src * c->dd, (size_t)c->dd * sizeof(int64_t));
    } else {
        c->mu[dst] = c->mu[src];
    }
}

void orc_additem(orc_comps *c, int64_t k, const orc_data *d, int64_t i)
{
    if (c->kind == ORC_G1G1) {                      /* src/conjugates.jl:60-63 */
        double x = d->x[i];
        c->nn[k] += 1;
        double n = (double)c->nn[k];
        c->mu[k] = ((n - 1) / n) * c->mu[k] + x / n;
        return;
    }
    int64_t *mi = c->mi + k * c->dd;
    if (d->kind == ORC_INT) {                       /* :153-157 */
        mi[d->w[i]] += 1;
        c->mm[k] += 1;
        c->nn[k] += 1;
    } else {                                        /* :163-170 */
        for (int64_t t = d->sptr[i]; t < d->sptr[i + 1]; ++t) {
            mi[d->sw[t]] += d->sc[t];
            c->mm[k] += d->sc[t];
        }
        c->nn[k] += 1;
    }
}

int orc_delitem(orc_comps *c, int64_t k, const orc_data *d, int64_t i)
{
    if (c->kind == ORC_G1G1) {                      /* src/conjugates.jl:64-74 */
        double x = d->x[i];
        if (c->nn[k] < 1) return -1;                /* the reference throws a String here (:66) */
        if (c->nn[k] == 1) {
            c->mu[k] = c->m0;
            c->nn[k] = 0;
        } else {
            double n = (double)c->nn[k];
            c->mu[k] = c->mu[k] * (n / (n - 1)) - x / (n - 1);
            c->nn[k] -= 1;
        }
        return 0;
    }
    int64_t *mi = c->mi + k * c->dd;
    if (d->kind == ORC_INT) {                       /* :158-162 */
        mi[d->w[i]] -= 1;
        c->mm[k] -= 1;
        c->nn[k] -= 1;
    } else {                                        /* :171-178 */
        for (int64_t t = d->sptr[i]; t < d->sptr[i + 1]; ++t) {
            mi[d->sw[t]] -= d->sc[t];
            c->mm[k] -= d->sc[t];
        }
        c->nn[k] -= 1;
    }
    return 0;
}

void orc_g1g1_posterior(const orc_comps *c, int64_t k, double *mu, double *vv)
{                                                   /* src/conjugates.jl:77-89 */
    if (c->nn[k] > 0) {
        double n = (double)c->nn[k];
        double denom = c->vv + n * c->v0;
        *mu = (n * c->v0 * c->mu[k] + c->vv * c->m0) / denom;
        *vv = c->v0 * c->vv / denom;
    } else {
        *mu = c->m0;
        *vv = c->v0;
    }
}

/* logpredictive of a Sent under MultinomialDirichlet, src/conjugates.jl:181-198.
 * The reference materialises lll = aa + mi over the whole vocabulary and sums
 * lgamma(lll) - lgamma(aa + mi) over all V entries; entries of untouched words are
 * exactly 0.0, so the dense vector is built here too and summed with Julia's sum. */
static double md_logpred_sent(const orc_comps *c, int64_t k, const orc_data *d, int64_t i)
{
    const int64_t *mi = c->mi + k * c->dd;
    int64_t V = c->dd;
    double *lll = (double *)malloc((size_t)V * sizeof(double));
    double *term = (double *)malloc((size_t)V * sizeof(double));
    for (int64_t v = 0; v < V; ++v) lll[v] = c->aa + (double)mi[v];
    int64_t mm = 0;
    int64_t len = d->sptr[i + 1] - d->sptr[i];
    double *lg = (double *)malloc((size_t)(len > 0 ? len : 1) * sizeof(double));
    for (int64_t t = 0; t < len; ++t) {
        int64_t w = d->sw[d->sptr[i] + t], cnt = d->sc[d->sptr[i] + t];
        lll[w] += (double)cnt;
        mm += cnt;
        lg[t] = lgamma((double)cnt + 1.0);
    }
    for (int64_t v = 0; v < V; ++v) term[v] = lgamma(lll[v]) - lgamma(c->aa + (double)mi[v]);
    double ll = lgamma((double)mm + 1.0) - orc_julia_sum(lg, len)
              + lgamma(c->aa * (double)c->dd + (double)c->mm[k])
              - lgamma(c->aa * (double)c->dd + (double)c->mm[k] + (double)mm)
              + orc_julia_sum(term, V);
    free(lll); free(term); free(lg);
    return ll;
}

double orc_logpredictive(const orc_comps *c, int64_t k, const orc_data *d, int64_t i)
{
    if (c->kind == ORC_G1G1) {                      /* src/conjugates.jl:92-103 */
        double mu, vv;
        orc_g1g1_posterior(c, k, &mu, &vv);
        double x = d->x[i];
        double dummy = vv + c->vv;
        return -((x - mu) * (x - mu)) / (2 * dummy) - 0.5 * log(2 * M_PI * dummy);
    }
    if (d->kind == ORC_INT) {                       /* :180 */
        const int64_t *mi = c->mi + k * c->dd;
        return log((c->aa + (double)mi[d->w[i]]) / (c->aa * (double)c->dd + (double)c->mm[k]));
    }
    return md_logpred_sent(c, k, d, i);
}

/* The per-item "likelihood" diagnostic.  For MultinomialDirichlet this is the
 * posterior-MEAN probability, not a log (src/conjugates.jl:205-213 via :200-203 and
 * src/distributions.jl:62): alpha = mi + aa over the vocabulary, divided by sum(alpha). */
double orc_loglikelihood(const orc_comps *c, int64_t k, const orc_data *d, int64_t i)
{
    if (c->kind == ORC_G1G1) {                      /* :105 */
        double mu, vv;
        orc_g1g1_posterior(c, k, &mu, &vv);
        return orc_gaussian1d_logpdf(mu, vv, d->x[i]);
    }
    const int64_t *mi = c->mi + k * c->dd;
    int64_t V = c->dd;
    double *alpha = (double *)malloc((size_t)V * sizeof(double));
    for (int64_t v = 0; v < V; ++v) alpha[v] = (double)mi[v] + c->aa;
    double s = orc_julia_sum(alpha, V);
    double ll;
    if (d->kind == ORC_INT) {
        ll = alpha[d->w[i]] / s;
    } else {
        ll = 0.0;
        for (int64_t t = d->sptr[i]; t < d->sptr[i + 1]; ++t)
            ll += (double)d->sc[t] * (alpha[d->sw[t]] / s);
    }
    free(alpha);
    return ll;
}

/* O(1) closed form of the MD x Int diagnostic (same value up to the rounding of sum(alpha)). */
static double md_int_diag_fast(const orc_comps *c, int64_t k, int64_t w)
{
    return ((double)c->mi[k * c->dd + w] + c->aa) / ((double)c->mm[k] + c->aa * (double)c->dd);
}

/* ======================================================================
 * src/LDA.jl
 * ====================================================================== */
double orc_lda_init(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                    const int64_t *z, int64_t *nn, int with_diag)
{                                                   /* src/LDA.jl:91-98 */
    int64_t K = c->K;
    double ll = 0.0;
    for (int64_t j = 0; j < D; ++j)
        for (int64_t i = ptr[j]; i < ptr[j + 1]; ++i) {
            int64_t k = z[i];
            orc_additem(c, k, d, i);
            nn[j * K + k] += 1;
            if (with_diag) ll += orc_loglikelihood(c, k, d, i);
        }
    return ll;
}

/* LDA_sample_zz scores, src/LDA.jl:35-37. */
static void lda_scores(const orc_comps *c, const orc_data *d, int64_t i, const int64_t *nn_j,
                       double alpha, double *pp)
{
    for (int64_t k = 0; k < c->K; ++k)
        pp[k] = log((double)nn_j[k] + alpha) + orc_logpredictive(c, k, d, i);
}

double orc_lda_sweep(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                     int64_t *z, int64_t *nn, double alpha,
                     const int64_t *doc_order, const int64_t *tok_order,
                     const double *uniforms, int diag_mode)
{                                                   /* src/LDA.jl:113-133 */
    int64_t K = c->K;
    double *pp = (double *)malloc((size_t)K * sizeof(double));
    double ll = 0.0;
    int64_t pos = 0;       /* cursor into tok_order, which is grouped by visiting order */
    for (int64_t jo = 0; jo < D; ++jo) {
        int64_t j = doc_order ? doc_order[jo] : jo;
        int64_t len = ptr[j + 1] - ptr[j];
        for (int64_t io = 0; io < len; ++io) {
            int64_t i = tok_order ? tok_order[pos + io] : ptr[j] + io;
            int64_t k = z[i];
            orc_delitem(c, k, d, i);                /* :118-120 */
            nn[j * K + k] -= 1;
            lda_scores(c, d, i, nn + j * K, alpha, pp);    /* :124 -> :26-42 */
            orc_lognormalize(pp, K);
            k = orc_sample(pp, K, uniforms[i]);
            z[i] = k;                               /* :128-131 */
            orc_additem(c, k, d, i);
            nn[j * K + k] += 1;
            if (diag_mode == 2) ll += orc_loglikelihood(c, k, d, i);
            else if (diag_mode == 1) {
                if (c->kind == ORC_MD && d->kind == ORC_INT) ll += md_int_diag_fast(c, k, d->w[i]);
                else ll += orc_loglikelihood(c, k, d, i);
            }
        }
        pos += len;
    }
    free(pp);
    return ll;
}

/* Test support (convergence fixtures): the same serial sweep for word-id items with the conditional evaluated in the
 * linear domain — exp(log(n_dk+a) + log((b+n_kw)/(bV+n_k)) - max) / sum of src/LDA.jl:35-41 + src/common.jl:42-52 IS the
 * normalised weight (n_dk+a)(b+n_kw)/(bV+n_k) — in Float64, CDF walked in k order with the strict test of
 * src/common.jl:74.  Ten times faster than orc_lda_sweep (no log / exp per topic), which makes chains of hundreds of
 * sweeps at K = 1024 affordable; tests/test_oracle_golden.py checks that it makes the same draws as orc_lda_sweep. */
void orc_lda_sweep_linear(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr, int64_t *z, int64_t *nn,
                          double alpha, const int64_t *doc_order, const int64_t *tok_order, const double *uniforms)
{
    int64_t K = c->K, V = c->dd;
    double *pp = (double *)malloc((size_t)K * sizeof(double));
    double bV = c->aa * (double)V;
    int64_t pos = 0;
    for (int64_t jo = 0; jo < D; ++jo) {
        int64_t j = doc_order ? doc_order[jo] : jo;
        int64_t len = ptr[j + 1] - ptr[j];
        int64_t *nj = nn + j * K;
        for (int64_t io = 0; io < len; ++io) {
            int64_t i = tok_order ? tok_order[pos + io] : ptr[j] + io;
            int64_t w = d->w[i], k = z[i];
            c->mi[k * V + w] -= 1; c->mm[k] -= 1; c->nn[k] -= 1; nj[k] -= 1;
            double tot = 0.0;
            for (int64_t q = 0; q < K; ++q) {
                pp[q] = ((double)nj[q] + alpha) * ((c->aa + (double)c->mi[q * V + w]) / (bV + (double)c->mm[q]));
                tot += pp[q];
            }
            double r = uniforms[i] * tot, cw = pp[0];
            k = 0;
            while (cw < r && k < K - 1) { k += 1; cw += pp[k]; }
            z[i] = k;
            c->mi[k * V + w] += 1; c->mm[k] += 1; c->nn[k] += 1; nj[k] += 1;
        }
        pos += len;
    }
    free(pp);
}

/* Test support: the serial sweep of src/LDA.jl:113-133 (identity order) FOLLOWING another sampler's result.  At every
 * token the oracle computes its own conditional on its own state and its own draw for uniforms[i]; where that differs
 * from other_z[i] it records how far uniforms[i] is from the nearest edge of its CDF (the only legitimate reason for
 * two samplers fed the same uniform to disagree) and then ADOPTS other_z[i], so that the states stay comparable for
 * the rest of the sweep.  Returns the number of disagreements; *max_edge_dist <- the largest recorded distance,
 * *max_rel_logp <- largest |log p_or(k) - log p_or(k')| is not computed here (draws only). */
int64_t orc_lda_sweep_follow(orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                             int64_t *z, int64_t *nn, double alpha, const double *uniforms,
                             const int64_t *other_z, double *max_edge_dist)
{
    int64_t K = c->K, n_diff = 0;
    double *pp = (double *)malloc((size_t)K * sizeof(double));
    double worst = 0.0;
    for (int64_t j = 0; j < D; ++j) {
        for (int64_t i = ptr[j]; i < ptr[j + 1]; ++i) {
            int64_t k = z[i];
            orc_delitem(c, k, d, i);
            nn[j * K + k] -= 1;
            lda_scores(c, d, i, nn + j * K, alpha, pp);
            orc_lognormalize(pp, K);
            k = orc_sample(pp, K, uniforms[i]);
            if (k != other_z[i]) {
                double cw = 0.0, best = 1.0;
                for (int64_t q = 0; q < K; ++q) {
                    cw += pp[q];
                    double dist = fabs(cw - uniforms[i]);
                    if (dist < best) best = dist;
                }
                if (best > worst) worst = best;
                n_diff += 1;
                k = other_z[i];
            }
            z[i] = k;
            orc_additem(c, k, d, i);
            nn[j * K + k] += 1;
        }
    }
    free(pp);
    if (max_edge_dist) *max_edge_dist = worst;
    return n_diff;
}

int64_t orc_lda_conditional(orc_comps *c, const orc_data *d, int64_t j, int64_t i,
                            int64_t zi, int64_t *nn, double alpha, double r,
                            double *raw, double *prob)
{
    int64_t K = c->K;
    double mu_saved = (c->kind == ORC_G1G1) ? c->mu[zi] : 0.0;
    orc_delitem(c, zi, d, i);
    nn[j * K + zi] -= 1;
    lda_scores(c, d, i, nn + j * K, alpha, raw);
    memcpy(prob, raw, (size_t)K * sizeof(double));
    orc_lognormalize(prob, K);
    int64_t k = orc_sample(prob, K, r);
    orc_additem(c, zi, d, i);
    nn[j * K + zi] += 1;
    if (c->kind == ORC_G1G1) c->mu[zi] = mu_saved;  /* undo running-mean drift: state is restored exactly */
    return k;
}

/* ======================================================================
 * src/BMM.jl
 * ====================================================================== */
double orc_bmm_init(orc_comps *c, const orc_data *d, int64_t N, const int64_t *z, int64_t *nn)
{                                                   /* src/BMM.jl:70-75 */
    double ll = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        int64_t k = z[i];
        orc_additem(c, k, d, i);
        nn[k] += 1;
        ll += orc_loglikelihood(c, k, d, i);
    }
    return ll;
}

static void bmm_scores(const orc_comps *c, const orc_data *d, int64_t i, const int64_t *nn,
                       double alpha, double *pp)
{                                                   /* src/BMM.jl:102-105 */
    for (int64_t k = 0; k < c->K; ++k)
        pp[k] = log((double)nn[k] + alpha) + orc_logpredictive(c, k, d, i);
}

double orc_bmm_sweep(orc_comps *c, const orc_data *d, int64_t N, int64_t *z, int64_t *nn,
                     double alpha, const int64_t *order, const double *uniforms)
{                                                   /* src/BMM.jl:92-115 */
    int64_t K = c->K;
    double *pp = (double *)malloc((size_t)K * sizeof(double));
    double ll = 0.0;
    for (int64_t io = 0; io < N; ++io) {
        int64_t i = order ? order[io] : io;
        int64_t k = z[i];
        nn[k] -= 1;
        orc_delitem(c, k, d, i);
        bmm_scores(c, d, i, nn, alpha, pp);
        orc_lognormalize(pp, K);
        k = orc_sample(pp, K, uniforms[i]);
        z[i] = k;
        nn[k] += 1;
        orc_additem(c, k, d, i);
        ll += orc_loglikelihood(c, k, d, i);
    }
    free(pp);
    return ll;
}

int64_t orc_bmm_conditional(orc_comps *c, const orc_data *d, int64_t i, int64_t zi, int64_t *nn,
                            double alpha, double r, double *raw, double *prob)
{
    int64_t K = c->K;
    double mu_saved = (c->kind == ORC_G1G1) ? c->mu[zi] : 0.0;
    nn[zi] -= 1;
    orc_delitem(c, zi, d, i);
    bmm_scores(c, d, i, nn, alpha, raw);
    memcpy(prob, raw, (size_t)K * sizeof(double));
    orc_lognormalize(prob, K);
    int64_t k = orc_sample(prob, K, r);
    nn[zi] += 1;
    orc_additem(c, zi, d, i);
    if (c->kind == ORC_G1G1) c->mu[zi] = mu_saved;
    return k;
}

/* ======================================================================
 * RNG (the reference uses Julia's global MersenneTwister and Distributions.jl,
 * neither reproducible here: draws are "parity unpinned").
 * ====================================================================== */
static uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_rng_seed(orc_rng *g, uint64_t seed)
{
    for (int i = 0; i < 4; ++i) g->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(orc_rng *g)
{   /* xoshiro256** */
    uint64_t *s = g->s;
    uint64_t result = rotl64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
double orc_rand(orc_rng *g) { return (double)(rng_next(g) >> 11) * (1.0 / 9007199254740992.0); }
double orc_randn(orc_rng *g)
{
    double u1, u2;
    do { u1 = orc_rand(g); } while (u1 <= 0.0);
    u2 = orc_rand(g);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}
double orc_rand_gamma(orc_rng *g, double shape)
{
    if (shape < 1.0) {
        double u;
        do { u = orc_rand(g); } while (u <= 0.0);
        return orc_rand_gamma(g, shape + 1.0) * pow(u, 1.0 / shape);
    }
    double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
    for (;;) {
        double x, v;
        do { x = orc_randn(g); v = 1.0 + cc * x; } while (v <= 0.0);
        v = v * v * v;
        double u = orc_rand(g);
        if (u < 1.0 - 0.0331 * x * x * x * x) return dd * v;
        if (u > 0.0 && log(u) < 0.5 * x * x + dd * (1.0 - v + log(v))) return dd * v;
    }
}
double orc_rand_beta(orc_rng *g, double a, double b)
{
    double x = orc_rand_gamma(g, a), y = orc_rand_gamma(g, b);
    return x / (x + y);
}
void orc_randperm(orc_rng *g, int64_t n, int64_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = i;
    for (int64_t i = n - 1; i > 0; --i) {
        int64_t j = (int64_t)(orc_rand(g) * (double)(i + 1));
        if (j > i) j = i;
        int64_t t = out[i]; out[i] = out[j]; out[j] = t;
    }
}

/* ======================================================================
 * src/HDP.jl — direct-assignment sampler
 * ====================================================================== */
double orc_hdp_init(orc_hdp *h, orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                    const int64_t *z, int64_t *nn)
{                                                   /* src/HDP.jl:217-228 */
    int64_t S = h->Kmax;
    double ll = 0.0;
    for (int64_t j = 0; j < D; ++j)
        for (int64_t i = ptr[j]; i < ptr[j + 1]; ++i) {
            int64_t k = z[i];
            orc_additem(c, k, d, i);
            nn[j * S + k] += 1;
            ll += orc_loglikelihood(c, k, d, i);
        }
    for (int64_t k = 0; k <= h->KK; ++k) h->beta[k] = 1.0 / (double)(h->KK + 1);   /* :228 */
    return ll;
}

static int64_t hdp_colsum(const int64_t *nn, int64_t D, int64_t S, int64_t k)
{
    int64_t s = 0;
    for (int64_t j = 0; j < D; ++j) s += nn[j * S + k];
    return s;
}

/* Remove component kk: del_column + splice!(components) + shift every zz > kk down +
 * fold its stick into the unrepresented mass (src/HDP.jl:251-262 and :293-306). */
static void hdp_delete_component(orc_hdp *h, orc_comps *c, int64_t D, const int64_t *ptr,
                                 int64_t *z, int64_t *nn, int64_t kk)
{
    int64_t S = h->Kmax, KK = h->KK;
    for (int64_t j = 0; j < D; ++j)
        for (int64_t k = kk; k < KK - 1; ++k) nn[j * S + k] = nn[j * S + k + 1];
    for (int64_t j = 0; j < D; ++j) nn[j * S + KK - 1] = 0;
    for (int64_t k = kk; k < KK - 1; ++k) orc_comps_move(c, k, k + 1);
    orc_comps_reset(c, KK - 1);
    int64_t N = ptr[D];
    for (int64_t i = 0; i < N; ++i) if (z[i] > kk) z[i] -= 1;
    h->beta[KK] += h->beta[kk];
    for (int64_t k = kk; k < KK; ++k) h->beta[k] = h->beta[k + 1];
    h->KK -= 1;
}

static void hdp_scores(const orc_hdp *h, orc_comps *c, const orc_data *d, int64_t i,
                       const int64_t *nn_j, double *pp)
{                                                   /* src/HDP.jl:312-315 */
    int64_t KK = h->KK;
    for (int64_t l = 0; l < KK; ++l)
        pp[l] = log((double)nn_j[l] + h->aa * h->beta[l]) + orc_logpredictive(c, l, d, i);
    /* the new-component term uses the empty prototype hdp.component: slot Kmax-1 is
     * reserved by the callers as an always-empty scratch component when KK < Kmax */
    pp[KK] = log(h->aa * h->beta[KK]) + orc_logpredictive(c, c->K - 1, d, i);
}

double orc_hdp_sweep(orc_hdp *h, orc_comps *c, const orc_data *d, int64_t D, const int64_t *ptr,
                     int64_t *z, int64_t *nn, orc_rng *g, const double *uniforms, int shuffle)
{
    int64_t S = h->Kmax;
    double *pp = (double *)malloc((size_t)(S + 1) * sizeof(double));
    double ll = 0.0;

    /* :248-264 — `for kk = 1:hdp.KK` evaluates its range once while the body shrinks
     * hdp.KK; after a deletion the column that slid into kk is skipped.  Where the
     * reference would index past the shrunken matrix (BoundsError) we stop. */
    int64_t KK0 = h->KK;
    for (int64_t kk = 0; kk < KK0; ++kk) {
        if (kk >= h->KK) break;
        if (hdp_colsum(nn, D, S, kk) == 0) hdp_delete_component(h, c, D, ptr, z, nn, kk);
    }

    int64_t *dorder = (int64_t *)malloc((size_t)(D > 0 ? D : 1) * sizeof(int64_t));
    if (shuffle) orc_randperm(g, D, dorder); else for (int64_t j = 0; j < D; ++j) dorder[j] = j;
    for (int64_t jo = 0; jo < D; ++jo) {
        int64_t j = dorder[jo];
        int64_t len = ptr[j + 1] - ptr[j];
        int64_t *torder = (int64_t *)malloc((size_t)(len > 0 ? len : 1) * sizeof(int64_t));
        if (shuffle) orc_randperm(g, len, torder); else for (int64_t t = 0; t < len; ++t) torder[t] = t;
        for (int64_t io = 0; io < len; ++io) {
            int64_t i = ptr[j] + torder[io];
            int64_t kk = z[i];
            orc_delitem(c, kk, d, i);               /* :285-287 */
            nn[j * S + kk] -= 1;
            if (hdp_colsum(nn, D, S, kk) == 0)      /* :291-307 */
                hdp_delete_component(h, c, D, ptr, z, nn, kk);

            hdp_scores(h, c, d, i, nn + j * S, pp); /* :312-315 */
            orc_lognormalize(pp, h->KK + 1);        /* :316 */
            double r = uniforms ? uniforms[i] : orc_rand(g);
            kk = orc_sample(pp, h->KK + 1, r);      /* :317 */

            if (kk == h->KK) {                      /* :320-331 */
                if (h->KK + 1 >= S) {               /* capacity (last slot is the empty prototype) */
                    free(torder); free(dorder); free(pp);
                    return NAN;
                }
                orc_comps_reset(c, h->KK);
                double b = orc_rand_beta(g, 1.0, h->gg);
                double b_new = h->beta[h->KK];
                h->beta[h->KK] = b * b_new;
                h->beta[h->KK + 1] = (1 - b) * b_new;
                h->KK += 1;
            }
            z[i] = kk;                              /* :335-338 */
            orc_additem(c, kk, d, i);
            nn[j * S + kk] += 1;
            ll += orc_loglikelihood(c, kk, d, i);
        }
        free(torder);
    }
    free(dorder); free(pp);
    return ll;
}

int64_t orc_hdp_conditional(const orc_hdp *h, orc_comps *c, const orc_data *d, int64_t j, int64_t i,
                            int64_t zi, int64_t *nn, double r, double *raw, double *prob)
{
    int64_t S = h->Kmax, n = h->KK + 1;
    double mu_saved = (c->kind == ORC_G1G1) ? c->mu[zi] : 0.0;
    orc_delitem(c, zi, d, i);
    nn[j * S + zi] -= 1;
    hdp_scores(h, c, d, i, nn + j * S, raw);
    memcpy(prob, raw, (size_t)n * sizeof(double));
    orc_lognormalize(prob, n);
    int64_t k = orc_sample(prob, n, r);
    orc_additem(c, zi, d, i);
    nn[j * S + zi] += 1;
    if (c->kind == ORC_G1G1) c->mu[zi] = mu_saved;
    return k;
}

/* Normalised log unsigned Stirling numbers of the first kind via
 * s(n+1, m) = n s(n, m) + s(n, m-1) in log space; each row shifted so max = 0.
 * Any per-row scaling cancels in lognormalize! (src/HDP.jl:353-358). */
void orc_stirling_table(int64_t nmax, double *out)
{
    if (nmax < 1) return;
    double *prev = (double *)malloc((size_t)(nmax + 2) * sizeof(double));
    double *cur = (double *)malloc((size_t)(nmax + 2) * sizeof(double));
    prev[1] = 0.0;                                  /* s(1,1) = 1 */
    out[0] = 0.0;
    for (int64_t n = 1; n < nmax; ++n) {            /* build row n+1 from row n (unnormalised logs kept in prev) */
        double logn = log((double)n);
        double mx = -INFINITY;
        for (int64_t m = 1; m <= n + 1; ++m) {
            double a = (m <= n) ? logn + prev[m] : -INFINITY;    /* n * s(n, m) */
            double b = (m >= 2) ? prev[m - 1] : -INFINITY;       /* s(n, m-1) */
            double hi = a > b ? a : b, lo = a > b ? b : a;
            double v = (lo == -INFINITY) ? hi : hi + log1p(exp(lo - hi));
            cur[m] = v;
            if (v > mx) mx = v;
        }
        double *row = out + (n + 1) * n / 2;
        for (int64_t m = 1; m <= n + 1; ++m) {
            cur[m] -= mx;                           /* keep the running row normalised too: avoids growth */
            row[m - 1] = cur[m];
        }
        double *t = prev; prev = cur; cur = t;
    }
    free(prev); free(cur);
}

int64_t orc_hdp_table_conditional(const double *lstir, int64_t n, double a, double r, double *prob)
{                                                   /* src/HDP.jl:353-358 */
    const double *row = lstir + n * (n - 1) / 2;
    double la = log(a);
    for (int64_t m = 1; m <= n; ++m) prob[m - 1] = row[m - 1] + (double)m * la;
    orc_lognormalize(prob, n);
    return orc_sample(prob, n, r) + 1;
}

void orc_hdp_sample_hyperparam(orc_hdp *h, int64_t D, const int64_t *ptr, int64_t m, orc_rng *g)
{                                                   /* src/HDP.jl:124-159 */
    double sum_logw = 0.0;
    int64_t sum_s = 0;
    double *logw = (double *)malloc((size_t)(D > 0 ? D : 1) * sizeof(double));
    for (int64_t j = 0; j < D; ++j) {
        double nj = (double)(ptr[j + 1] - ptr[j]);
        double w = orc_rand_beta(g, h->aa + 1.0, nj);          /* :136 */
        logw[j] = log(w);
    }
    for (int64_t j = 0; j < D; ++j) {
        double nj = (double)(ptr[j + 1] - ptr[j]);
        double p = nj / h->aa;                                  /* :138-139 */
        p = p / (p + 1.0);
        sum_s += (orc_rand(g) < p) ? 1 : 0;                     /* Binomial(1, p), :143 */
    }
    sum_logw = orc_julia_sum(logw, D);
    free(logw);
    double aa_shape = h->a1 + (double)m - (double)sum_s;        /* :146 */
    double aa_rate = h->a2 - sum_logw;                          /* :147 */
    h->aa = orc_rand_gamma(g, aa_shape) / aa_rate;              /* :148 */

    double eta = orc_rand_beta(g, h->gg + 1.0, (double)m);      /* :151 */
    double pi_eta = 1.0 / (1.0 + ((double)m * (h->g2 - log(eta))) / (h->g1 + (double)h->KK - 1.0));  /* :152 */
    if (orc_rand(g) < pi_eta)
        h->gg = orc_rand_gamma(g, h->g1 + (double)h->KK) / (h->g2 - log(eta));        /* :155 */
    else
        h->gg = orc_rand_gamma(g, h->g1 + (double)h->KK - 1.0) / (h->g2 - log(eta));  /* :157 */
}

int orc_hdp_resample_beta(orc_hdp *h, int64_t D, const int64_t *ptr, const int64_t *nn,
                          const double *lstir, int64_t nmax, int n_internals, int sample_hyper,
                          orc_rng *g, int64_t *M_out)
{                                                   /* src/HDP.jl:343-373 */
    int64_t S = h->Kmax, KK = h->KK;
    int64_t maxn = 1;
    for (int64_t j = 0; j < D; ++j)
        for (int64_t k = 0; k < KK; ++k) {
            if (nn[j * S + k] > nmax) return -1;    /* the reference hits a BoundsError on the 10k table */
            if (nn[j * S + k] > maxn) maxn = nn[j * S + k];
        }
    double *rr = (double *)malloc((size_t)maxn * sizeof(double));
    double *MM = (double *)malloc((size_t)(KK + 1) * sizeof(double));
    memset(M_out, 0, (size_t)(D * S) * sizeof(int64_t));
    for (int hh = 0; hh < n_internals; ++hh) {
        for (int64_t j = 0; j < D; ++j)
            for (int64_t k = 0; k < KK; ++k) {
                int64_t n = nn[j * S + k];
                if (n == 0) M_out[j * S + k] = 0;
                else M_out[j * S + k] = orc_hdp_table_conditional(lstir, n, h->aa * h->beta[k], orc_rand(g), rr);
            }
        int64_t m = 0;
        for (int64_t k = 0; k < KK; ++k) {
            int64_t s = 0;
            for (int64_t j = 0; j < D; ++j) s += M_out[j * S + k];
            MM[k] = (double)s;
            m += s;
        }
        MM[KK] = h->gg;                              /* :363 */
        double tot = 0.0;                            /* beta ~ Dirichlet(MM), :364 */
        for (int64_t k = 0; k <= KK; ++k) { h->beta[k] = orc_rand_gamma(g, MM[k]); tot += h->beta[k]; }
        for (int64_t k = 0; k <= KK; ++k) h->beta[k] /= tot;
        if (sample_hyper) orc_hdp_sample_hyperparam(h, D, ptr, m, g);     /* :369-372 */
    }
    free(rr); free(MM);
    return 0;
}

/* ======================================================================
 * src/RCRP.jl — recurrent Chinese restaurant process
 * ====================================================================== */
void orc_rcrp_init(orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, const int64_t *z, int64_t *nn)
{                                                   /* src/RCRP.jl:98-104, :131-143 */
    int64_t S = h->Kmax;
    for (int64_t t = 0; t < h->TT; ++t)
        for (int64_t i = ptr[t]; i < ptr[t + 1]; ++i) {
            nn[t * S + z[i]] += 1;
            orc_additem(c, z[i], d, i);
        }
}

static int64_t rcrp_colsum(const int64_t *nn, int64_t S, int64_t t0, int64_t t1, int64_t k)
{
    int64_t s = 0;
    for (int64_t t = t0; t <= t1; ++t) s += nn[t * S + k];
    return s;
}

static void rcrp_delete_component(orc_rcrp *h, orc_comps *c, const int64_t *ptr, int64_t *z, int64_t *nn, int64_t kk)
{                                                   /* :184-193 (and the same block in the other epochs) */
    int64_t S = h->Kmax, KK = h->KK, TT = h->TT;
    for (int64_t i = 0; i < ptr[TT]; ++i) if (z[i] > kk) z[i] -= 1;
    for (int64_t k = kk; k < KK - 1; ++k) orc_comps_move(c, k, k + 1);
    orc_comps_reset(c, KK - 1);
    for (int64_t t = 0; t < TT; ++t) {
        for (int64_t k = kk; k < KK - 1; ++k) nn[t * S + k] = nn[t * S + k + 1];
        nn[t * S + KK - 1] = 0;
    }
    h->KK -= 1;
}

/* sum(log(collect(0:n-1) + base)) as the reference writes it (:176, :201-204, :211) */
static double rcrp_logsum(int64_t n, double base)
{
    if (n <= 0) return 0.0;
    double *v = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; ++i) v[i] = log((double)i + base);
    double s = orc_julia_sum(v, n);
    free(v);
    return s;
}

/* Scores of one item of epoch t (0-based) on the current (item-removed) state.  pp has KK+1 entries. */
static void rcrp_scores(const orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, int64_t t, int64_t i,
                        const int64_t *nn, double *pp)
{
    int64_t S = h->Kmax, KK = h->KK, TT = h->TT;
    int first = (t == 0), last = (t == TT - 1);
    for (int64_t k = 0; k <= KK; ++k) pp[k] = -INFINITY;           /* pp = fill(-Inf, KK+1) */
    if (last) {                                                     /* :366-377 */
        for (int64_t k = 0; k < KK; ++k) {
            int64_t cnt = nn[(t - 1) * S + k] + nn[t * S + k];
            if (cnt > 0) pp[k] = log((double)cnt) + orc_logpredictive(c, k, d, i);
        }
        pp[KK] = log(h->aa[t]) + orc_logpredictive(c, c->K - 1, d, i);
        return;
    }
    int64_t N_t = ptr[t + 1] - ptr[t], N_next = ptr[t + 2] - ptr[t + 1];
    double L2 = rcrp_logsum(N_next, (double)N_t + h->aa[t]);       /* :176, :246 (N_t is the full epoch size) */
    int64_t KK_curnex = 1;
    for (int64_t k = 0; k < KK; ++k)
        if (rcrp_colsum(nn, S, first ? t : t - 1, t + 1, k) > 0) KK_curnex += 1;      /* :196-198, :267-269 */
    double a = h->aa[t] / (double)KK_curnex;
    double *L1 = (double *)calloc((size_t)(KK > 0 ? KK : 1), sizeof(double));
    for (int64_t k = 0; k < KK; ++k)
        if (rcrp_colsum(nn, S, first ? t : t - 1, t + 1, k) > 0)
            L1[k] = rcrp_logsum(nn[(t + 1) * S + k], (double)nn[t * S + k] + a);     /* :200-203 */
    for (int64_t k = 0; k < KK; ++k) {
        int64_t cnt = first ? nn[t * S + k] : nn[(t - 1) * S + k] + nn[t * S + k];
        if (cnt <= 0) continue;                                     /* kk_idx_precur */
        double old = L1[k];
        L1[k] = rcrp_logsum(nn[(t + 1) * S + k], (double)nn[t * S + k] + 1.0 + a);   /* :209-211 */
        pp[k] = log((double)cnt) + orc_logpredictive(c, k, d, i) + orc_julia_sum(L1, KK) - L2;
        L1[k] = old;
    }
    pp[KK] = log(h->aa[t]) + orc_logpredictive(c, c->K - 1, d, i) + orc_julia_sum(L1, KK) - L2;   /* :217 */
    free(L1);
}

static void rcrp_sample_hyperparam(orc_rcrp *h, int64_t t, int64_t n_items, int64_t KK, int iters, orc_rng *g)
{                                                   /* src/RCRP.jl:60-74 */
    /* Reference quirk, reproduced: the loop at :62 is `for n = 1:iters`, which rebinds the argument n
     * (the epoch size), so inside the body n is the iteration number 1..iters, not N_t[tt]. */
    (void)n_items;
    for (int it = 1; it <= iters; ++it) {
        double n = (double)it;
        double eta = orc_rand_beta(g, h->aa[t] + 1.0, n);
        double rr = (h->a1 + (double)KK - 1.0) / (n * (h->a2 - log(eta)));
        double pi_eta = rr / (1.0 + rr);
        if (orc_rand(g) < pi_eta) h->aa[t] = orc_rand_gamma(g, h->a1 + (double)KK) / (h->a2 - log(eta));
        else h->aa[t] = orc_rand_gamma(g, h->a1 + (double)KK - 1.0) / (h->a2 - log(eta));
    }
}

double orc_rcrp_sweep(orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, int64_t *z, int64_t *nn,
                      orc_rng *g, const double *uniforms, int shuffle, int sample_hyper, int n_internals)
{
    int64_t S = h->Kmax, TT = h->TT;
    double *pp = (double *)malloc((size_t)(S + 1) * sizeof(double));
    double ll = 0.0;
    for (int64_t t = 0; t < TT; ++t) {
        if (t == 0 || t == TT - 1) ll = (t == 0) ? 0.0 : ll;       /* the reference resets the diagnostic at the first epoch only (:171) */
        int64_t len = ptr[t + 1] - ptr[t];
        int64_t *order = (int64_t *)malloc((size_t)(len > 0 ? len : 1) * sizeof(int64_t));
        if (shuffle) orc_randperm(g, len, order); else for (int64_t q = 0; q < len; ++q) order[q] = q;
        for (int64_t q = 0; q < len; ++q) {
            int64_t i = ptr[t] + order[q];
            int64_t kk = z[i];
            orc_delitem(c, kk, d, i);
            nn[t * S + kk] -= 1;
            if (rcrp_colsum(nn, S, 0, TT - 1, kk) == 0) rcrp_delete_component(h, c, ptr, z, nn, kk);
            rcrp_scores(h, c, d, ptr, t, i, nn, pp);
            orc_lognormalize(pp, h->KK + 1);
            kk = orc_sample(pp, h->KK + 1, uniforms ? uniforms[i] : orc_rand(g));
            if (kk == h->KK) {                                      /* :222-226 */
                if (h->KK + 1 >= S) { free(order); free(pp); return NAN; }
                orc_comps_reset(c, h->KK);
                h->KK += 1;
            }
            z[i] = kk;
            nn[t * S + kk] += 1;
            orc_additem(c, kk, d, i);
            ll += orc_loglikelihood(c, kk, d, i);
        }
        free(order);
        if (t > 0 && t < TT - 1) {
            /* :309-331 — a chain that is empty at t but alive before and after is split: everything it holds
             * from t+1 on becomes a new component */
            for (int64_t kk = 0; kk < h->KK; ++kk) {
                if (nn[t * S + kk] == 0 && nn[(t + 1) * S + kk] != 0 && rcrp_colsum(nn, S, 0, t, kk) != 0) {
                    if (h->KK + 1 >= S) { free(pp); return NAN; }
                    int64_t kn = h->KK;
                    orc_comps_reset(c, kn);
                    for (int64_t tt = t + 1; tt < TT; ++tt) {
                        for (int64_t i = ptr[tt]; i < ptr[tt + 1]; ++i)
                            if (z[i] == kk) {
                                z[i] = kn;
                                orc_delitem(c, kk, d, i);
                                orc_additem(c, kn, d, i);
                            }
                        nn[tt * S + kn] = nn[tt * S + kk];
                        nn[tt * S + kk] = 0;
                    }
                    h->KK += 1;
                }
            }
        }
        if (sample_hyper) {
            int64_t KK_t = 0;
            for (int64_t k = 0; k < h->KK; ++k) if (nn[t * S + k] != 0) KK_t += 1;
            rcrp_sample_hyperparam(h, t, len, KK_t, n_internals, g);
        }
    }
    free(pp);
    return ll;
}

int64_t orc_rcrp_conditional(const orc_rcrp *h, orc_comps *c, const orc_data *d, const int64_t *ptr, int64_t t,
                             int64_t i, int64_t zi, int64_t *nn, double r, double *raw, double *prob)
{
    int64_t S = h->Kmax, n = h->KK + 1;
    double mu_saved = (c->kind == ORC_G1G1) ? c->mu[zi] : 0.0;
    orc_delitem(c, zi, d, i);
    nn[t * S + zi] -= 1;
    rcrp_scores(h, c, d, ptr, t, i, nn, raw);
    memcpy(prob, raw, (size_t)n * sizeof(double));
    orc_lognormalize(prob, n);
    int64_t k = orc_sample(prob, n, r);
    orc_additem(c, zi, d, i);
    nn[t * S + zi] += 1;
    if (c->kind == ORC_G1G1) c->mu[zi] = mu_saved;
    return k;
}

/* ======================================================================
 * src/HDP.jl:402-709 — CRF_gibbs_sampler! (Chinese restaurant franchise: tables and dishes)
 * ====================================================================== */
void orc_crf_init(const orc_hdp *h, orc_crf *f, orc_comps *c, const orc_data *d, const int64_t *z)
{                                                   /* :441-470 */
    for (int64_t k = 0; k < h->Kmax; ++k) f->mm[k] = 0;
    for (int64_t j = 0; j < f->G; ++j) {
        int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0, nt = 0;
        for (int64_t i = 0; i < n; ++i) {            /* kjt[jj] = unique(zz[jj]): first-appearance order */
            int64_t kk = z[p0 + i], t = -1;
            for (int64_t q = 0; q < nt; ++q) if (f->kjt[p0 + q] == kk) { t = q; break; }
            if (t < 0) { t = nt++; f->kjt[p0 + t] = kk; f->njt[p0 + t] = 0; }
            orc_additem(c, kk, d, p0 + i);
            f->tji[p0 + i] = t;
            f->njt[p0 + t] += 1;
        }
        f->n_tbls[j] = nt;
        for (int64_t q = 0; q < nt; ++q) f->mm[f->kjt[p0 + q]] += 1;
    }
}

/* a dish nobody serves any more leaves the menu (:531-545, :610-621): everything above it moves down by one */
static void crf_delete_dish(orc_hdp *h, orc_crf *f, orc_comps *c, int64_t *z, int64_t kk)
{
    int64_t KK = h->KK, N = f->ptr[f->G];
    for (int64_t i = 0; i < N; ++i) if (z[i] > kk) z[i] -= 1;
    for (int64_t j = 0; j < f->G; ++j)
        for (int64_t q = 0; q < f->n_tbls[j]; ++q) if (f->kjt[f->ptr[j] + q] > kk) f->kjt[f->ptr[j] + q] -= 1;
    for (int64_t k = kk; k < KK - 1; ++k) { orc_comps_move(c, k, k + 1); f->mm[k] = f->mm[k + 1]; }
    orc_comps_reset(c, KK - 1);
    f->mm[KK - 1] = 0;
    h->KK -= 1;
}

/* Conditional of the table of item i of group j with the item removed from its table (:548-556); the table is
 * NOT deleted when the removal empties it (its entry is log(0) = -Inf).  raw/prob have n_tbls[j] + 1 entries. */
int64_t orc_crf_table_conditional(const orc_hdp *h, const orc_crf *f, orc_comps *c, const orc_data *d, int64_t j,
                                  int64_t i, int64_t *z, double r, double *raw, double *prob)
{
    int64_t p0 = f->ptr[j], nt = f->n_tbls[j], kk = z[i], t0 = f->tji[i];
    double mu_saved = (c->kind == ORC_G1G1) ? c->mu[kk] : 0.0;
    orc_delitem(c, kk, d, i);
    for (int64_t t = 0; t < nt; ++t) {
        int64_t n = f->njt[p0 + t] - (t == t0 ? 1 : 0);
        raw[t] = log((double)n) + orc_logpredictive(c, f->kjt[p0 + t], d, i);
    }
    raw[nt] = log(h->aa) + orc_logpredictive(c, c->K - 1, d, i);
    memcpy(prob, raw, (size_t)(nt + 1) * sizeof(double));
    orc_lognormalize(prob, nt + 1);
    int64_t t = orc_sample(prob, nt + 1, r);
    orc_additem(c, kk, d, i);
    if (c->kind == ORC_G1G1) c->mu[kk] = mu_saved;
    return t;
}

double orc_crf_sweep(orc_hdp *h, orc_crf *f, orc_comps *c, const orc_data *d, int64_t *z, orc_rng *g,
                     const double *uniforms, int defer_deaths, int sample_hyper, int n_internals)
{
    int64_t S = h->Kmax, G = f->G;
    int64_t maxn = 1;
    for (int64_t j = 0; j < G; ++j) if (f->ptr[j + 1] - f->ptr[j] > maxn) maxn = f->ptr[j + 1] - f->ptr[j];
    double *pp = (double *)malloc((size_t)((S > maxn ? S : maxn) + 2) * sizeof(double));
    int64_t *tidx = (int64_t *)malloc((size_t)maxn * sizeof(int64_t));
#define CRF_U(idx, which) (uniforms ? uniforms[3 * (idx) + (which)] : orc_rand(g))
    /* 1. resampling tables for customers (:506-587) */
    for (int64_t j = 0; j < G; ++j) {
        int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0;
        for (int64_t ii = 0; ii < n; ++ii) {
            int64_t i = p0 + ii, kk = z[i], tbl = f->tji[i];
            orc_delitem(c, kk, d, i);
            f->njt[p0 + tbl] -= 1;
            if (f->njt[p0 + tbl] == 0) {                        /* the table empties: remove it (:517-546) */
                int64_t nt = f->n_tbls[j];
                f->mm[kk] -= 1;
                for (int64_t q = tbl; q < nt - 1; ++q) { f->kjt[p0 + q] = f->kjt[p0 + q + 1]; f->njt[p0 + q] = f->njt[p0 + q + 1]; }
                f->n_tbls[j] = nt - 1;
                for (int64_t q = 0; q < n; ++q) if (f->tji[p0 + q] > tbl) f->tji[p0 + q] -= 1;
                if (f->mm[kk] == 0 && !defer_deaths) crf_delete_dish(h, f, c, z, kk);
            }
            int64_t nt = f->n_tbls[j];
            for (int64_t t = 0; t < nt; ++t)
                pp[t] = log((double)f->njt[p0 + t]) + orc_logpredictive(c, f->kjt[p0 + t], d, i);      /* :551-554 */
            pp[nt] = log(h->aa) + orc_logpredictive(c, c->K - 1, d, i);                                /* :555 */
            orc_lognormalize(pp, nt + 1);
            tbl = orc_sample(pp, nt + 1, CRF_U(i, 0));
            if (tbl == nt) {                                    /* a new table: draw its dish (:559-581) */
                f->n_tbls[j] = nt + 1;
                f->njt[p0 + tbl] = 0;
                for (int64_t k = 0; k < h->KK; ++k) pp[k] = log((double)f->mm[k]) + orc_logpredictive(c, k, d, i);
                pp[h->KK] = log(h->gg) + orc_logpredictive(c, c->K - 1, d, i);
                orc_lognormalize(pp, h->KK + 1);
                kk = orc_sample(pp, h->KK + 1, CRF_U(i, 1));
                if (kk == h->KK) {
                    if (h->KK + 1 >= S) { free(pp); free(tidx); return NAN; }
                    orc_comps_reset(c, h->KK);
                    f->mm[h->KK] = 0;
                    h->KK += 1;
                }
                f->kjt[p0 + tbl] = kk;
                f->mm[kk] += 1;
            }
            kk = f->kjt[p0 + tbl];
            f->tji[i] = tbl;
            f->njt[p0 + tbl] += 1;
            orc_additem(c, kk, d, i);
            z[i] = kk;
        }
    }
    /* 2. resampling dishes (:593-655) */
    for (int64_t j = 0; j < G; ++j) {
        int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0;
        for (int64_t tbl = 0; tbl < f->n_tbls[j]; ++tbl) {
            int64_t kk = f->kjt[p0 + tbl], nti = 0;
            f->mm[kk] -= 1;
            for (int64_t q = 0; q < n; ++q) if (f->tji[p0 + q] == tbl) tidx[nti++] = p0 + q;
            if (f->mm[kk] == 0 && !defer_deaths) {
                crf_delete_dish(h, f, c, z, kk);                /* the component is dropped with its items (:608-621) */
            } else {
                for (int64_t q = 0; q < nti; ++q) orc_delitem(c, kk, d, tidx[q]);                     /* :626-628 */
            }
            for (int64_t k = 0; k < h->KK; ++k) {               /* :633-639 */
                pp[k] = log((double)f->mm[k]);
                for (int64_t q = 0; q < nti; ++q) pp[k] += orc_logpredictive(c, k, d, tidx[q]);
            }
            pp[h->KK] = 0.0;
            pp[h->KK] += log(h->gg);                            /* :640-643 */
            for (int64_t q = 0; q < nti; ++q) pp[h->KK] += orc_logpredictive(c, c->K - 1, d, tidx[q]);
            orc_lognormalize(pp, h->KK + 1);
            kk = orc_sample(pp, h->KK + 1, CRF_U(p0 + tbl, 2));
            if (kk == h->KK) {
                if (h->KK + 1 >= S) { free(pp); free(tidx); return NAN; }
                orc_comps_reset(c, h->KK);
                f->mm[h->KK] = 0;
                h->KK += 1;
            }
            f->kjt[p0 + tbl] = kk;
            for (int64_t q = 0; q < nti; ++q) { orc_additem(c, kk, d, tidx[q]); z[tidx[q]] = kk; }
            f->mm[kk] += 1;
        }
    }
#undef CRF_U
    if (defer_deaths) {                             /* dishes that lost their last table leave the menu now */
        for (int64_t k = h->KK - 1; k >= 0; --k) if (f->mm[k] == 0) crf_delete_dish(h, f, c, z, k);
    }
    if (sample_hyper) {                             /* :658-667 */
        int64_t m = 0;
        for (int64_t j = 0; j < G; ++j) m += f->n_tbls[j];
        for (int it = 0; it < n_internals; ++it) orc_hdp_sample_hyperparam(h, G, f->ptr, m, g);
    }
    double ll = 0.0;                                /* :669-673 */
    int64_t N = f->ptr[G];
    for (int64_t i = 0; i < N; ++i) ll += orc_logpredictive(c, z[i], d, i);
    free(pp);
    free(tidx);
    return ll;
}

/* ======================================================================
 * src/RCRF.jl — recurrent Chinese restaurant franchise (RCRF_gibbs_sampler!, :122-938)
 * Per epoch: the CRF's two passes with the dish prior coupled to the neighbouring epochs' menus.
 * ====================================================================== */
static int64_t rcrf_mm_colsum(const orc_rcrf *f, int64_t k)
{
    int64_t s = 0;
    for (int64_t t = 0; t < f->TT; ++t) s += f->mm[t * f->Kmax + k];
    return s;
}

void orc_rcrf_init(orc_rcrf *f, orc_comps *c, const orc_data *d, const int64_t *z)
{                                                   /* :177-209 */
    for (int64_t i = 0; i < f->TT * f->Kmax; ++i) f->mm[i] = 0;
    for (int64_t t = 0; t < f->TT; ++t)
        for (int64_t j = f->eptr[t]; j < f->eptr[t + 1]; ++j) {
            int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0, nt = 0;
            for (int64_t i = 0; i < n; ++i) {
                int64_t kk = z[p0 + i], tb = -1;
                for (int64_t q = 0; q < nt; ++q) if (f->kjt[p0 + q] == kk) { tb = q; break; }
                if (tb < 0) { tb = nt++; f->kjt[p0 + tb] = kk; f->njt[p0 + tb] = 0; }
                orc_additem(c, kk, d, p0 + i);
                f->tji[p0 + i] = tb;
                f->njt[p0 + tb] += 1;
            }
            f->n_tbls[j] = nt;
            for (int64_t q = 0; q < nt; ++q) f->mm[t * f->Kmax + f->kjt[p0 + q]] += 1;
        }
}

static void rcrf_delete_dish(orc_rcrf *f, orc_comps *c, int64_t *z, int64_t kk)
{                                                   /* :271-284 and its copies */
    int64_t KK = f->KK, S = f->Kmax, G = f->eptr[f->TT], N = f->ptr[G];
    for (int64_t i = 0; i < N; ++i) if (z[i] > kk) z[i] -= 1;
    for (int64_t j = 0; j < G; ++j)
        for (int64_t q = 0; q < f->n_tbls[j]; ++q) if (f->kjt[f->ptr[j] + q] > kk) f->kjt[f->ptr[j] + q] -= 1;
    for (int64_t k = kk; k < KK - 1; ++k) orc_comps_move(c, k, k + 1);
    orc_comps_reset(c, KK - 1);
    for (int64_t t = 0; t < f->TT; ++t) {
        for (int64_t k = kk; k < KK - 1; ++k) f->mm[t * S + k] = f->mm[t * S + k + 1];
        f->mm[t * S + KK - 1] = 0;
    }
    f->KK -= 1;
}

/* log prior weight of every dish for a table of epoch t, and of a new dish in pr[KK] (:303-327, :534-556, :762-768).
 * Written as the reference writes it, with the O(n) sums; note the reference's L1 uses gg/K' in the base term and aa/K'
 * in the "+1" term (:313 vs :321) — reproduced. */
static void rcrf_dish_prior(const orc_rcrf *f, int64_t t, double *pr)
{
    int64_t S = f->Kmax, KK = f->KK, TT = f->TT;
    const int64_t *mt = f->mm + t * S, *mp = (t > 0) ? f->mm + (t - 1) * S : NULL, *mn = (t < TT - 1) ? f->mm + (t + 1) * S : NULL;
    for (int64_t k = 0; k <= KK; ++k) pr[k] = -INFINITY;
    if (!mn) {                                      /* last epoch */
        for (int64_t k = 0; k < KK; ++k) {
            int64_t cnt = mt[k] + (mp ? mp[k] : 0);
            if (cnt > 0) pr[k] = log((double)cnt);
        }
        pr[KK] = log(f->gg[t]);
        return;
    }
    int64_t KK_curnex = 1, sum_t = 0, sum_n = 0;
    for (int64_t k = 0; k < KK; ++k) { if (mt[k] + mn[k] > 0) KK_curnex += 1; sum_t += mt[k]; sum_n += mn[k]; }
    double *L1 = (double *)calloc((size_t)(KK > 0 ? KK : 1), sizeof(double));
    for (int64_t k = 0; k < KK; ++k)
        if (mt[k] + mn[k] > 0) L1[k] = rcrp_logsum(mn[k], (double)mt[k] + f->gg[t] / (double)KK_curnex);
    double L2 = rcrp_logsum(sum_n, (double)sum_t + f->gg[t]);
    for (int64_t k = 0; k < KK; ++k) {
        int64_t cnt = mt[k] + (mp ? mp[k] : 0);
        if (cnt <= 0) continue;
        double old = L1[k];
        L1[k] = rcrp_logsum(mn[k], (double)mt[k] + 1.0 + f->aa[t] / (double)KK_curnex);
        pr[k] = log((double)cnt) + orc_julia_sum(L1, KK) - L2;
        L1[k] = old;
    }
    pr[KK] = log(f->gg[t]) + orc_julia_sum(L1, KK) - L2;
    free(L1);
}

static void rcrf_sample_hyperparam(orc_rcrf *f, int64_t t, int64_t m, int64_t KK, int iters, orc_rng *g)
{                                                   /* :80-118 */
    int64_t j0 = f->eptr[t], j1 = f->eptr[t + 1], D = j1 - j0;
    double *logw = (double *)malloc((size_t)(D > 0 ? D : 1) * sizeof(double));
    for (int it = 0; it < iters; ++it) {
        int64_t sum_s = 0;
        for (int64_t j = j0; j < j1; ++j) logw[j - j0] = log(orc_rand_beta(g, f->aa[t] + 1.0, (double)(f->ptr[j + 1] - f->ptr[j])));
        for (int64_t j = j0; j < j1; ++j) {
            double p = (double)(f->ptr[j + 1] - f->ptr[j]) / f->aa[t];
            p = p / (p + 1.0);
            sum_s += (orc_rand(g) < p) ? 1 : 0;
        }
        double aa_shape = f->a1 + (double)m - (double)sum_s, aa_rate = f->a2 - orc_julia_sum(logw, D);
        f->aa[t] = orc_rand_gamma(g, aa_shape) / aa_rate;
        double eta = orc_rand_beta(g, f->gg[t] + 1.0, (double)m);
        double pi_eta = 1.0 / (1.0 + ((double)m * (f->g2 - log(eta))) / (f->g1 + (double)KK - 1.0));
        if (orc_rand(g) < pi_eta) f->gg[t] = orc_rand_gamma(g, f->g1 + (double)KK) / (f->g2 - log(eta));
        else f->gg[t] = orc_rand_gamma(g, f->g1 + (double)KK - 1.0) / (f->g2 - log(eta));
    }
    free(logw);
}

double orc_rcrf_sweep(orc_rcrf *f, orc_comps *c, const orc_data *d, int64_t *z, orc_rng *g, const double *uniforms,
                      int defer_deaths, int join_tables, int sample_hyper, int n_internals)
{
    int64_t S = f->Kmax, TT = f->TT, G = f->eptr[TT];
    int64_t maxn = 1;
    for (int64_t j = 0; j < G; ++j) if (f->ptr[j + 1] - f->ptr[j] > maxn) maxn = f->ptr[j + 1] - f->ptr[j];
    double *pp = (double *)malloc((size_t)((S > maxn ? S : maxn) + 2) * sizeof(double));
    double *pr = (double *)malloc((size_t)(S + 2) * sizeof(double));
    int64_t *tidx = (int64_t *)malloc((size_t)maxn * sizeof(int64_t));
    int64_t *tmp = (int64_t *)malloc((size_t)maxn * 3 * sizeof(int64_t));
#define RCRF_U(idx, which) (uniforms ? uniforms[3 * (idx) + (which)] : orc_rand(g))
#define RCRF_FAIL() do { free(pp); free(pr); free(tidx); free(tmp); return NAN; } while (0)
    for (int64_t t = 0; t < TT; ++t) {
        /* tables for customers (:238-349, :477-576, :705-792) */
        for (int64_t j = f->eptr[t]; j < f->eptr[t + 1]; ++j) {
            int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0;
            for (int64_t ii = 0; ii < n; ++ii) {
                int64_t i = p0 + ii, kk = z[i], tbl = f->tji[i];
                orc_delitem(c, kk, d, i);
                f->njt[p0 + tbl] -= 1;
                if (f->njt[p0 + tbl] == 0) {
                    int64_t nt = f->n_tbls[j];
                    kk = f->kjt[p0 + tbl];
                    for (int64_t q = tbl; q < nt - 1; ++q) { f->kjt[p0 + q] = f->kjt[p0 + q + 1]; f->njt[p0 + q] = f->njt[p0 + q + 1]; }
                    f->n_tbls[j] = nt - 1;
                    for (int64_t q = 0; q < n; ++q) if (f->tji[p0 + q] > tbl) f->tji[p0 + q] -= 1;
                    f->mm[t * S + kk] -= 1;
                    if (!defer_deaths && rcrf_mm_colsum(f, kk) == 0) rcrf_delete_dish(f, c, z, kk);
                }
                int64_t nt = f->n_tbls[j];
                for (int64_t q = 0; q < nt; ++q)
                    pp[q] = log((double)f->njt[p0 + q]) + orc_logpredictive(c, f->kjt[p0 + q], d, i);
                pp[nt] = log(f->aa[t]) + orc_logpredictive(c, c->K - 1, d, i);
                orc_lognormalize(pp, nt + 1);
                tbl = orc_sample(pp, nt + 1, RCRF_U(i, 0));
                if (tbl == nt) {
                    f->n_tbls[j] = nt + 1;
                    f->njt[p0 + tbl] = 0;
                    rcrf_dish_prior(f, t, pr);
                    for (int64_t k = 0; k < f->KK; ++k)
                        pp[k] = (pr[k] == -INFINITY) ? -INFINITY : pr[k] + orc_logpredictive(c, k, d, i);
                    pp[f->KK] = pr[f->KK] + orc_logpredictive(c, c->K - 1, d, i);
                    orc_lognormalize(pp, f->KK + 1);
                    kk = orc_sample(pp, f->KK + 1, RCRF_U(i, 1));
                    if (kk == f->KK) {
                        if (f->KK + 1 >= S) RCRF_FAIL();
                        orc_comps_reset(c, f->KK);
                        for (int64_t tt = 0; tt < TT; ++tt) f->mm[tt * S + f->KK] = 0;
                        f->KK += 1;
                    }
                    f->kjt[p0 + tbl] = kk;
                    f->mm[t * S + kk] += 1;
                }
                kk = f->kjt[p0 + tbl];
                f->tji[i] = tbl;
                f->njt[p0 + tbl] += 1;
                orc_additem(c, kk, d, i);
                z[i] = kk;
            }
        }
        /* tables of a restaurant that serve the same dish are joined (:351-377) */
        if (join_tables)
            for (int64_t j = f->eptr[t]; j < f->eptr[t + 1]; ++j) {
                int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0, nt = f->n_tbls[j], nu = 0;
                int64_t *new_k = tmp, *new_n = tmp + maxn, *map = tmp + 2 * maxn;
                for (int64_t q = 0; q < nt; ++q) {
                    int64_t u = -1;
                    for (int64_t r = 0; r < nu; ++r) if (new_k[r] == f->kjt[p0 + q]) { u = r; break; }
                    if (u < 0) { u = nu++; new_k[u] = f->kjt[p0 + q]; new_n[u] = 0; }
                    else f->mm[t * S + f->kjt[p0 + q]] -= 1;
                    new_n[u] += f->njt[p0 + q];
                    map[q] = u;
                }
                if (nu != nt) {
                    for (int64_t q = 0; q < n; ++q) f->tji[p0 + q] = map[f->tji[p0 + q]];
                    for (int64_t r = 0; r < nu; ++r) { f->kjt[p0 + r] = new_k[r]; f->njt[p0 + r] = new_n[r]; }
                    f->n_tbls[j] = nu;
                }
            }
        /* dishes for tables (:380-461, :606-688, :822-894).  Deviation, documented: when the table is the last one of its
         * dish in this epoch but the dish lives on in other epochs the reference skips the delitem! of the table's
         * customers (:395-416, the else belongs to the outer if) and they end up counted twice; here they are removed. */
        for (int64_t j = f->eptr[t]; j < f->eptr[t + 1]; ++j) {
            int64_t p0 = f->ptr[j], n = f->ptr[j + 1] - p0;
            for (int64_t tbl = 0; tbl < f->n_tbls[j]; ++tbl) {
                int64_t kk = f->kjt[p0 + tbl], nti = 0;
                f->mm[t * S + kk] -= 1;
                for (int64_t q = 0; q < n; ++q) if (f->tji[p0 + q] == tbl) tidx[nti++] = p0 + q;
                for (int64_t q = 0; q < nti; ++q) orc_delitem(c, kk, d, tidx[q]);
                if (!defer_deaths && rcrf_mm_colsum(f, kk) == 0) rcrf_delete_dish(f, c, z, kk);
                rcrf_dish_prior(f, t, pr);
                for (int64_t k = 0; k <= f->KK; ++k) {
                    if (pr[k] == -INFINITY) { pp[k] = -INFINITY; continue; }
                    pp[k] = pr[k];
                    for (int64_t q = 0; q < nti; ++q) pp[k] += orc_logpredictive(c, k < f->KK ? k : c->K - 1, d, tidx[q]);
                }
                orc_lognormalize(pp, f->KK + 1);
                kk = orc_sample(pp, f->KK + 1, RCRF_U(p0 + tbl, 2));
                if (kk == f->KK) {
                    if (f->KK + 1 >= S) RCRF_FAIL();
                    orc_comps_reset(c, f->KK);
                    for (int64_t tt = 0; tt < TT; ++tt) f->mm[tt * S + f->KK] = 0;
                    f->KK += 1;
                }
                f->kjt[p0 + tbl] = kk;
                for (int64_t q = 0; q < nti; ++q) { orc_additem(c, kk, d, tidx[q]); z[tidx[q]] = kk; }
                f->mm[t * S + kk] += 1;
            }
        }
        if (sample_hyper) {                         /* :464-468 */
            int64_t KK_t = 0, m = 0;
            for (int64_t k = 0; k < f->KK; ++k) if (f->mm[t * S + k] != 0) KK_t += 1;
            for (int64_t j = f->eptr[t]; j < f->eptr[t + 1]; ++j) m += f->n_tbls[j];
            rcrf_sample_hyperparam(f, t, m, KK_t, n_internals, g);
        }
    }
#undef RCRF_U
#undef RCRF_FAIL
    if (defer_deaths)
        for (int64_t k = f->KK - 1; k >= 0; --k) if (rcrf_mm_colsum(f, k) == 0) rcrf_delete_dish(f, c, z, k);
    free(pp); free(pr); free(tidx); free(tmp);
    double ll = 0.0;
    int64_t N = f->ptr[G];
    for (int64_t i = 0; i < N; ++i) ll += orc_logpredictive(c, z[i], d, i);
    return ll;
}
