// generic.cuh — one warp per item, Float64, log domain: the assignment step for every
// (model, conjugate, datum) combination of the reference, and the verification mode.
//
//   LDA  src/LDA.jl:26-42,116-131      score_k = log(n_jk + aa)          + logpredictive_k(x)
//   BMM  src/BMM.jl:95-113             score_k = log(n_k  + aa)          + logpredictive_k(x)
//   HDP  src/HDP.jl:285-338            score_k = log(n_jk + aa*beta_k)   + logpredictive_k(x), k < KK
//                                      score_KK = log(aa*beta_u) + logpredictive(prototype, x)
//   logpredictive: src/conjugates.jl:92-103 (Gaussian1DGaussian1D), :180 (MD x Int), :181-198 (MD x Sent)
//   lognormalize!: src/common.jl:42-52     sample: src/common.jl:69-79
//
// The item's own contribution is removed arithmetically from the counts it reads
// (delitem!, src/conjugates.jl:64-74,158-178) and the tables are touched with atomics only
// when the assignment changes.
//
// faithful = 1 (verification): after the K scores are in shared memory, lane 0 alone performs
// the max-shifted exp, the left-to-right sum, the division and the strict `cw < r` walk in the
// reference's operation order, so draws for supplied uniforms are reproducible bit for bit up
// to the last-ulp differences between CUDA's and glibc's log/exp/lgamma.
// faithful = 0 (production): the exp and the CDF scan are spread over the warp.
#pragma once
#include "common.cuh"

namespace bias {

struct GenericParams {
    int32_t model;              // BIAS_MODEL_*
    int32_t kind, datum;        // conjugate kind / datum kind
    int32_t K, Kpad, n_out;     // n_out = K, or K+1 for HDP
    int64_t n_work;
    const int64_t *item_idx;    // nullable: work index -> item (verification / pending lists)
    const int32_t *work_idx32;  // nullable: same, 32-bit list (HDP pending list)
    const int32_t *row_of;      // item -> row of prior_rows (group id); nullable => row 0
    int32_t row_by_work;        // 1: row_of is indexed by work index instead of item
    int32_t *prior_rows;        // [rows][Kpad] n_jk (LDA/HDP) or n_items (BMM)
    int32_t prior_is_items;     // BMM: prior_rows aliases n_items
    // observations
    const int32_t *word;
    const double *x;
    const int64_t *sent_ptr;
    const int32_t *sent_word;
    const int32_t *sent_count;
    int32_t *z;
    // component statistics
    int32_t *n_wk;              // MD   [V][Kpad]
    int32_t *n_k;               // MD   [Kpad] token totals (mm)
    int32_t *n_items;           // both [Kpad] items per component (nn)
    double *sumx;               // G1G1 [Kpad] sum of assigned x
    int64_t V;
    double beta;                // MD per-dimension concentration (aa/dd)
    double m0, v0, vv;
    double alpha;               // LDA/BMM aa; HDP aa (alpha0)
    const double *hdp_beta;     // HDP my_beta[KK+1]
    // RCRP (src/RCRP.jl): groups are epochs; the prior couples epoch t with t-1 and t+1
    int32_t rcrp_T;             // number of epochs
    const double *rcrp_aa;      // [T] per-epoch concentration
    int64_t work_base;          // first item of the launch when no index list is given
    // RNG
    uint64_t seed;
    uint32_t sweep;
    uint32_t rng_stream;
    const double *ext_uniforms; // nullable, indexed by work index when item_idx != NULL else by item
    // modes / outputs
    int32_t frozen, faithful;
    double *raw_out, *prob_out; // [n_work][n_out], nullable
    int32_t *draw_out;          // [n_work], frozen mode
    double *diag_sum;
    unsigned long long *moved;
    // HDP deferred births
    int32_t *pending_list;
    unsigned int *pending_count;
};

constexpr int kGenWarps = 4;
constexpr int kGenThreads = kGenWarps * 32;
__host__ __device__ inline size_t generic_smem_bytes(int n_out) {
    return (size_t)kGenWarps * (size_t)((n_out + 3) & ~3) * sizeof(double);
}

#ifdef __CUDACC__

// Julia 0.4 Base.sum association (left-to-right below 1024 elements, pairwise halves above).
__device__ inline double jsum_dev(const double *a, int lo, int hi) {
    if (lo == hi) return a[lo];
    if (lo + 1024 > hi) {
        double v = a[lo] + a[lo + 1];
        for (int i = lo + 2; i <= hi; ++i) v += a[i];
        return v;
    }
    const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
    return jsum_dev(a, lo, mid) + jsum_dev(a, mid + 1, hi);
}

__device__ __forceinline__ double g1g1_logpred(double n, double sx, double x, double m0, double v0, double vv) {
    double mu, vp;
    if (n > 0.0) {                       // posterior, src/conjugates.jl:77-89 with n*v0*mean == v0*sum(x)
        const double denom = vv + n * v0;
        mu = (v0 * sx + vv * m0) / denom;
        vp = v0 * vv / denom;
    } else {
        mu = m0;
        vp = v0;
    }
    const double dummy = vp + vv;        // src/conjugates.jl:98-99
    return -((x - mu) * (x - mu)) / (2.0 * dummy) - 0.5 * log(2.0 * 3.141592653589793 * dummy);
}

__device__ __forceinline__ double g1g1_loglik(double n, double sx, double x, double m0, double v0, double vv) {
    double mu, vp;                       // logpdf(posterior(me), x), src/conjugates.jl:105, src/distributions.jl:36-37
    if (n > 0.0) {
        const double denom = vv + n * v0;
        mu = (v0 * sx + vv * m0) / denom;
        vp = v0 * vv / denom;
    } else {
        mu = m0;
        vp = v0;
    }
    return log(exp(-(x - mu) * (x - mu) / (2.0 * vp)) / sqrt(2.0 * 3.141592653589793 * vp));
}

// ---- building blocks shared by the sweep kernel and the serial HDP birth pass --------------
struct Datum {
    int w;
    double xv;
    int64_t s0, s1;
    double m_tot, coef;
};

template <int KIND, int DATUM>
__device__ __forceinline__ Datum load_datum(const GenericParams &p, int64_t i, int lane) {
    Datum d{0, 0.0, 0, 0, 0.0, 0.0};
    if (DATUM == BIAS_INT) {
        d.w = p.word[i];
    } else if (DATUM == BIAS_F64) {
        d.xv = p.x[i];
    } else {
        d.s0 = p.sent_ptr[i];
        d.s1 = p.sent_ptr[i + 1];
        double mloc = 0.0, lg = 0.0;
        for (int64_t t = d.s0 + lane; t < d.s1; t += 32) {
            const double c = (double)p.sent_count[t];
            mloc += c;
            if (p.faithful) lg += lgamma(c + 1.0);
        }
        d.m_tot = warp_sum(mloc);
        // the multinomial coefficient is k-independent (src/conjugates.jl:193): kept only where raw scores are compared
        if (p.faithful) d.coef = lgamma(d.m_tot + 1.0) - warp_sum(lg);
    }
    return d;
}

// A running product of positive factors kept as mantissa x 2^e: renorm() moves the exponent bits of `m` into `e` and leaves
// m in [1, 2).  Branch-free and exact (no rounding), so products of any length never overflow and ONE log is taken at the
// end.  (The first version folded log(prod) into an accumulator whenever a lane's product passed 1e100: a quarter of the
// kernel's instructions at config 2 were those logs, profiles/r02_other_configs_ncu_summary.md.)
__device__ __forceinline__ void renorm(double &m, int &e) {
    const int hi = __double2hiint(m);
    e += (hi >> 20) - 1023;
    m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(m));
}

// Production scoring of a Sent item (src/conjugates.jl:181-198 without the k-independent multinomial coefficient):
//   sc[k] <- log prior_k + sum_w [lgamma(b + n_kw + c_w) - lgamma(b + n_kw)] + lgamma(bV + n_k) - lgamma(bV + n_k + m)
// for the topics of the lane (2 lane, 2 lane + 1 of every block of 64).  On entry sc[k] holds the PRIOR FACTOR of topic k in the linear
// domain (0 = the topic cannot be chosen).  Counts are integers, so every lgamma difference is the log of a rising
// factorial: score_k = log( prior_k * prod_w prod_{i<c_w} (b + n_kw + i) / prod_{i<m} (bV + n_k + i) ) — two products in
// Float64 with the exponents kept apart (renorm), one division and one log per topic.  The item's (word, count) pairs
// are read once by the lanes and broadcast with shuffles; two topics of a lane advance together (independent chains);
// in the denominator consecutive factors are taken in pairs, f (f + 1) = fma(f, f, f).
template <int DUMMY = 0>
__device__ __forceinline__ void sent_scores_fast(const GenericParams &p, const Datum &d, int zold, int K, int n_out,
                                                 double *sc, int lane) {
    const int Kpad = p.Kpad;
    const double beta = p.beta, bv = p.beta * (double)p.V, m_tot = d.m_tot;
    const int32_t *const n_wk = p.n_wk, *const n_k = p.n_k;
    const int32_t *const sent_word = p.sent_word, *const sent_count = p.sent_count;
    const int64_t s0 = d.s0, s1 = d.s1;
    const int m_int = (int)m_tot;
    for (int kbase = 0; kbase < n_out; kbase += 64) {          // warp-uniform trip count: the loop body shuffles
        // a lane owns two ADJACENT topics: their counts of a word are one 8-byte load
        const int ka = kbase + 2 * lane, kb = ka + 1;
        const bool has_a = ka < n_out, has_b = kb < n_out;
        const bool fresh_a = ka >= K, fresh_b = kb >= K;          // HDP: the empty prototype (or no topic at all)
        const bool own_a = has_a && ka == zold, own_b = has_b && kb == zold;
        const int ia = fresh_a ? 0 : ka;                          // column pair to load (pair 0 when unused; padded columns are 0)
        double pa = has_a ? sc[ka] : 1.0, pb = has_b ? sc[kb] : 1.0;
        const bool dead_a = !(pa > 0.0), dead_b = !(pb > 0.0);
        if (dead_a) pa = 1.0;
        if (dead_b) pb = 1.0;
        int ea = 0, eb = 0;
        renorm(pa, ea);
        renorm(pb, eb);
        // one (word, count) pair: c >= 1 factors per topic
        auto one_word = [&](const int2 n2, const int c) {
            int xa = n2.x, xb = n2.y;
            if (fresh_a) xa = 0; else if (own_a) xa -= c;
            if (fresh_b) xb = 0; else if (own_b) xb -= c;
            double fa = beta + (double)xa, fb = beta + (double)xb;
            // counts are small: the first four factors without a loop, the rest (rare) in one
            pa *= fa; pb *= fb;
            if (c > 1) { fa += 1.0; fb += 1.0; pa *= fa; pb *= fb; }
            if (c > 2) { fa += 1.0; fb += 1.0; pa *= fa; pb *= fb; }
            if (c > 3) { fa += 1.0; fb += 1.0; pa *= fa; pb *= fb; }
            if (c > 4) {
                renorm(pa, ea);
                renorm(pb, eb);
                for (int i = 4; i < c; ++i) {
                    fa += 1.0; fb += 1.0; pa *= fa; pb *= fb;
                    if ((i & 7) == 7) { renorm(pa, ea); renorm(pb, eb); }
                }
                renorm(pa, ea);
                renorm(pb, eb);
            }
        };
        for (int64_t t0 = s0; t0 < s1; t0 += 32) {
            const int n_here = (int)min((int64_t)32, s1 - t0);
            int w_l = 0, c_l = 0;
            if (lane < n_here) {
                w_l = sent_word[t0 + lane];
                c_l = sent_count[t0 + lane];
            }
            // four (word, count) pairs per step: their count loads are issued together
            int t = 0;
            for (; t + 4 <= n_here; t += 4) {
                int ct[4];
                int2 n2[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int wq = __shfl_sync(kFull, w_l, t + q);
                    ct[q] = __shfl_sync(kFull, c_l, t + q);
                    n2[q] = __ldcg(reinterpret_cast<const int2 *>(n_wk + (size_t)wq * Kpad + ia));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ct[q] > 0) one_word(n2[q], ct[q]);       // a pair with count 0 has no factors (uniform branch)
                renorm(pa, ea);                // at most 16 factors below 2^32 since the last one
                renorm(pb, eb);
            }
            for (; t < n_here; ++t) {
                const int wq = __shfl_sync(kFull, w_l, t), cq = __shfl_sync(kFull, c_l, t);
                const int2 n2 = __ldcg(reinterpret_cast<const int2 *>(n_wk + (size_t)wq * Kpad + ia));
                if (cq > 0) one_word(n2, cq);
            }
            renorm(pa, ea);
            renorm(pb, eb);
        }
        // lgamma(bV + n_k) - lgamma(bV + n_k + m) = -log prod_{i<m} (bV + n_k + i)
        double nka = 0.0, nkb = 0.0;
        const int2 nk2 = __ldcg(reinterpret_cast<const int2 *>(n_k + ia));
        if (!fresh_a) nka = (double)nk2.x - (own_a ? m_tot : 0.0);
        if (!fresh_b) nkb = (double)nk2.y - (own_b ? m_tot : 0.0);
        double acc_a, acc_b;
        if (m_int <= (1 << 16)) {
            double qa = 1.0, qb = 1.0, fa = bv + nka, fb = bv + nkb;
            int ga = 0, gb = 0, i = 0;
            for (; i + 8 <= m_int; i += 8) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {                  // four pairs of factors: below 2^(4 * 70)
                    qa *= fma(fa, fa, fa);
                    qb *= fma(fb, fb, fb);
                    fa += 2.0;
                    fb += 2.0;
                }
                renorm(qa, ga);
                renorm(qb, gb);
            }
            for (; i < m_int; ++i) { qa *= fa; qb *= fb; fa += 1.0; fb += 1.0; }    // at most seven more
            renorm(qa, ga);
            renorm(qb, gb);
            acc_a = (double)(ea - ga) * 0.6931471805599453 + log(pa / qa);
            acc_b = (double)(eb - gb) * 0.6931471805599453 + log(pb / qb);
        } else {
            acc_a = (double)ea * 0.6931471805599453 + log(pa) + (lgamma(bv + nka) - lgamma(bv + nka + m_tot));
            acc_b = (double)eb * 0.6931471805599453 + log(pb) + (lgamma(bv + nkb) - lgamma(bv + nkb + m_tot));
        }
        if (has_a) sc[ka] = dead_a ? -INFINITY : acc_a;
        if (has_b) sc[kb] = dead_b ? -INFINITY : acc_b;
    }
}

// K (+1) unnormalised log scores into sc[]; returns the warp-wide maximum.
// zold < 0: the item is currently unassigned (nothing to remove).
template <int KIND, int DATUM>
__device__ __forceinline__ double score_all(const GenericParams &p, const Datum &d, int zold, int K, int n_out,
                                            const int32_t *prow, const double *hdp_beta, double alpha,
                                            double *sc, int lane, int epoch = 0) {
    const bool is_hdp = p.model == BIAS_MODEL_HDP;
    const bool is_rcrp = p.model == BIAS_MODEL_RCRP;
    const int Kpad = p.Kpad;
    double lmax = -INFINITY;
    // RCRP: rows of the neighbouring epochs and a = aa[t] / (1 + #components alive in the window)
    const int32_t *rprev = nullptr, *rnext = nullptr;
    double rc_a = 0.0;
    if (is_rcrp) {
        if (epoch > 0) rprev = prow - Kpad;
        if (epoch < p.rcrp_T - 1) rnext = prow + Kpad;
        if (rnext) {                                   // src/RCRP.jl:196-198, :267-269
            int alive = 0;
            for (int k = lane; k < K; k += 32) {
                const int tot = __ldcg(prow + k) - (k == zold ? 1 : 0) + (rprev ? __ldcg(rprev + k) : 0) + __ldcg(rnext + k);
                alive += (tot > 0) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(kFull, alive, o);
            rc_a = alpha / (double)(alive + 1);
        }
    }
    if (KIND == BIAS_MD && DATUM == BIAS_SENT && !p.faithful) {
        // the prior FACTOR of every outcome (the log of the code below, in the linear domain; 0 = cannot be chosen) goes
        // into the item's one product: no log of its own
        for (int k = lane; k < n_out; k += 32) {
            const bool own = (k == zold), fresh = (k >= K);
            double pf;
            if (is_hdp) {
                pf = fresh ? alpha * hdp_beta[K] : (double)(__ldcg(prow + k) - (own ? 1 : 0)) + alpha * hdp_beta[k];
            } else if (is_rcrp) {
                if (fresh) {
                    pf = alpha;
                } else {
                    const int cur = __ldcg(prow + k) - (own ? 1 : 0);
                    const int cnt = cur + (rprev ? __ldcg(rprev + k) : 0);
                    pf = cnt <= 0 ? 0.0 : (double)cnt;
                    if (cnt > 0 && rnext) pf *= ((double)cur + (double)__ldcg(rnext + k) + rc_a) / ((double)cur + rc_a);
                }
            } else {
                pf = (double)(__ldcg(prow + k) - (own ? 1 : 0)) + alpha;
            }
            sc[k] = pf;
        }
        __syncwarp();
        sent_scores_fast(p, d, zold, K, n_out, sc, lane);
        __syncwarp();
        for (int k = lane; k < n_out; k += 32) lmax = fmax(lmax, sc[k]);
        lmax = warp_max(lmax);
        __syncwarp();
        return lmax;
    }
    for (int k = lane; k < n_out; k += 32) {
        const bool own = (k == zold);
        const bool fresh = (k >= K);                  // HDP new-component slot: the empty prototype
        double prior;
        if (is_hdp) {
            prior = fresh ? log(alpha * hdp_beta[K])
                          : log((double)(__ldcg(prow + k) - (own ? 1 : 0)) + alpha * hdp_beta[k]);
        } else if (is_rcrp) {
            if (fresh) {
                prior = log(alpha);                                        // src/RCRP.jl:217
            } else {
                const int cur = __ldcg(prow + k) - (own ? 1 : 0);
                const int cnt = cur + (rprev ? __ldcg(rprev + k) : 0);   // :213 (first epoch), :283
                if (cnt <= 0) {
                    prior = -INFINITY;                                     // not in kk_idx_precur
                } else {
                    prior = log((double)cnt);
                    if (rnext) {
                        // sum(L1) - L2 is common to every outcome; what differs is L1[k] evaluated with
                        // n_tk + 1 (:209-211), and sum_{i<n'} log(i+n+1+a) - log(i+n+a) telescopes:
                        const double nx = (double)__ldcg(rnext + k);
                        prior += log(((double)cur + nx + rc_a) / ((double)cur + rc_a));
                    }
                }
            }
        } else {
            prior = log((double)(__ldcg(prow + k) - (own ? 1 : 0)) + alpha);
        }
        double lp;
        if (KIND == BIAS_G1G1) {
            double n = 0.0, sx = 0.0;
            if (!fresh) {
                n = (double)(__ldcg(p.n_items + k) - (own ? 1 : 0));
                sx = __ldcg(p.sumx + k) - (own ? d.xv : 0.0);
            }
            lp = g1g1_logpred(n, sx, d.xv, p.m0, p.v0, p.vv);
        } else if (DATUM == BIAS_INT) {
            double nw = 0.0, nk = 0.0;
            if (!fresh) {
                nw = (double)(__ldcg(p.n_wk + (size_t)d.w * Kpad + k) - (own ? 1 : 0));
                nk = (double)(__ldcg(p.n_k + k) - (own ? 1 : 0));
            }
            lp = log((p.beta + nw) / (p.beta * (double)p.V + nk));       // src/conjugates.jl:180
        } else {
            double nk = 0.0;
            if (!fresh) nk = (double)__ldcg(p.n_k + k) - (own ? d.m_tot : 0.0);
            double acc = 0.0;
            if (p.faithful) {
                // the reference's form: lgamma differences, summed in item order
                for (int64_t t = d.s0; t < d.s1; ++t) {
                    const int wt = p.sent_word[t];
                    const double ct = (double)p.sent_count[t];
                    double nw = 0.0;
                    if (!fresh) nw = (double)__ldcg(p.n_wk + (size_t)wt * Kpad + k) - (own ? ct : 0.0);
                    acc += lgamma(p.beta + nw + ct) - lgamma(p.beta + nw);
                }
            }
            const double bv = p.beta * (double)p.V;
            lp = d.coef + lgamma(bv + nk) - lgamma(bv + nk + d.m_tot) + acc;                   // src/conjugates.jl:193-195
        }
        const double s = prior + lp;
        sc[k] = s;
        lmax = fmax(lmax, s);
    }
    lmax = warp_max(lmax);
    __syncwarp();
    return lmax;
}

// lognormalize! + sample in the reference's operation order, by lane 0 (src/common.jl:42-52, 69-79).
// Leaves the normalised probabilities in sc[].
__device__ __forceinline__ int draw_faithful(double *sc, int n_out, double lmax, double u, int lane) {
    int knew = 0;
    if (lane == 0) {
        for (int k = 0; k < n_out; ++k) sc[k] = exp(sc[k] - lmax);
        const double tot = jsum_dev(sc, 0, n_out - 1);
        for (int k = 0; k < n_out; ++k) sc[k] /= tot;
        int idx = 0;
        double cw = sc[0];
        while (cw < u && idx < n_out - 1) {
            idx += 1;
            cw += sc[idx];
        }
        knew = idx;
    }
    knew = __shfl_sync(kFull, knew, 0);
    __syncwarp();
    return knew;
}

// Production draw: exp in parallel, CDF by 32-wide chunk scans in k order, first k with cdf >= u * total.
__device__ __forceinline__ int draw_parallel(double *sc, int n_out, double lmax, double u, int lane) {
    double total = 0.0;
    for (int b = 0; b < n_out; b += 32) {
        const int k = b + lane;
        double e = 0.0;
        if (k < n_out) {
            e = exp(sc[k] - lmax);
            sc[k] = e;
        }
        total += warp_sum(e);
    }
    __syncwarp();
    const double target = u * total;
    double base = 0.0;
    int knew = n_out - 1;
    for (int b = 0; b < n_out; b += 32) {
        const int k = b + lane;
        const double e = (k < n_out) ? sc[k] : 0.0;
        const double incl = base + warp_incl_scan(e, lane);
        const unsigned bal = __ballot_sync(kFull, (k < n_out) && (incl >= target));
        if (bal) {
            knew = b + __ffs(bal) - 1;
            break;
        }
        base = __shfl_sync(kFull, incl, 31);
    }
    __syncwarp();
    return knew;
}

// delitem! from zold (if >= 0) and additem! to kdst (if >= 0) on the shared tables.
template <int KIND, int DATUM>
__device__ __forceinline__ void apply_move(const GenericParams &p, const Datum &d, int zold, int kdst,
                                           int32_t *prow, int lane) {
    const int Kpad = p.Kpad;
    if (KIND == BIAS_MD) {
        if (DATUM == BIAS_INT) {
            if (lane == 0) {
                if (zold >= 0) { atomicAdd(p.n_wk + (size_t)d.w * Kpad + zold, -1); atomicAdd(p.n_k + zold, -1); }
                if (kdst >= 0) { atomicAdd(p.n_wk + (size_t)d.w * Kpad + kdst, 1); atomicAdd(p.n_k + kdst, 1); }
            }
        } else {
            for (int64_t t = d.s0 + lane; t < d.s1; t += 32) {
                const int wt = p.sent_word[t], ct = p.sent_count[t];
                if (zold >= 0) atomicAdd(p.n_wk + (size_t)wt * Kpad + zold, -ct);
                if (kdst >= 0) atomicAdd(p.n_wk + (size_t)wt * Kpad + kdst, ct);
            }
            if (lane == 0) {
                if (zold >= 0) atomicAdd(p.n_k + zold, -(int)d.m_tot);
                if (kdst >= 0) atomicAdd(p.n_k + kdst, (int)d.m_tot);
            }
        }
    } else if (lane == 0) {
        if (zold >= 0) atomicAdd(p.sumx + zold, -d.xv);
        if (kdst >= 0) atomicAdd(p.sumx + kdst, d.xv);
    }
    if (lane == 0) {
        if (zold >= 0) atomicAdd(p.n_items + zold, -1);
        if (kdst >= 0) atomicAdd(p.n_items + kdst, 1);
        if (!p.prior_is_items) {
            if (zold >= 0) atomicAdd(prow + zold, -1);
            if (kdst >= 0) atomicAdd(prow + kdst, 1);
        }
    }
}

// loglikelihood(components[kk], x) after the add (src/LDA.jl:131): valid on lane 0.
template <int KIND, int DATUM>
__device__ __forceinline__ double diag_after(const GenericParams &p, const Datum &d, int knew, int lane) {
    const int Kpad = p.Kpad;
    if (KIND == BIAS_G1G1)
        return g1g1_loglik((double)__ldcg(p.n_items + knew), __ldcg(p.sumx + knew), d.xv, p.m0, p.v0, p.vv);
    if (DATUM == BIAS_INT)
        return ((double)__ldcg(p.n_wk + (size_t)d.w * Kpad + knew) + p.beta) /
               ((double)__ldcg(p.n_k + knew) + p.beta * (double)p.V);
    double part = 0.0;
    const double den = (double)__ldcg(p.n_k + knew) + p.beta * (double)p.V;
    for (int64_t t = d.s0 + lane; t < d.s1; t += 32)
        part += (double)p.sent_count[t] *
                (((double)__ldcg(p.n_wk + (size_t)p.sent_word[t] * Kpad + knew) + p.beta) / den);
    return warp_sum(part);
}

#ifdef BIAS_GENERIC_KERNEL_IMPL      // the kernels themselves are emitted by model.cu only
// ---------------------------------------------------------------------------------------
// Gaussian1DGaussian1D x Float64 with few components (n_out <= kItemMaxOut): ONE THREAD per item.
// With K ~ 20 a warp per item leaves most lanes idle; here a thread evaluates its item's K (+1) scores itself.  The
// component statistics (n, sum x) are the same addresses for every lane, so a warp's loads of them are single
// broadcast transactions; the 32 items of a warp are consecutive, i.e. almost always of one group, so the prior row is
// shared too.  Same arithmetic as the warp-per-item kernel (scores, max-shift, exp, left-to-right sum — the
// reference's order, src/common.jl:42-52 — and the strict `cw < r` walk of :69-79), same uniforms, same updates.
// LDA / BMM / HDP; the RCRP prior and the verification modes stay on the warp-per-item kernel.
// ---------------------------------------------------------------------------------------
constexpr int kItemMaxOut = 64;
constexpr int kItemThreads = 256;
constexpr int kItemPriorN = 64;          // log(n + a_k) is tabulated for counts below this
__host__ __device__ inline size_t gauss_item_smem_bytes(int n_out) { return (size_t)n_out * kItemPriorN * sizeof(double); }

// Everything about a component that does not depend on the item is computed ONCE per batch of kItemThreads items and
// kept in shared memory: the posterior-predictive mean, 1 / (2 (v_post + vv)) and the log-normaliser (src/conjugates.jl:92-103
// needs two divisions and a log per (item, component); the first version spent 70 % of its instructions there and in
// log(n_jk + a_k)), and log(n + a_k) for n < 64 — a group's counts are tens.  What is left per (item, component) is one count
// load, four shared-memory reads and three multiply-adds; the item's OWN component is scored from the snapshot with
// the item taken out; exp() is only taken of scores within 60 of the maximum (the others add less than 1e-26 of the
// total: nothing in Float64).  The snapshot is as stale as one batch of one CTA — the same order as the ~300,000 items
// in flight on the GPU that the per-item loads could not see either.
__global__ void __launch_bounds__(kItemThreads) gauss_item_sweep_kernel(const GenericParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *prior_s = reinterpret_cast<double *>(smem_raw);             // [n_out][kItemPriorN]
    __shared__ double n_s[kItemMaxOut], sx_s[kItemMaxOut], mu_s[kItemMaxOut], i2d_s[kItemMaxOut], lc_s[kItemMaxOut],
        ab_s[kItemMaxOut], vp_s[kItemMaxOut];
    const Philox philox(p.seed);
    const bool is_hdp = p.model == BIAS_MODEL_HDP;
    const int K = p.K, Kpad = p.Kpad, n_out = p.n_out;
    const int tid = threadIdx.x;
    double diag_acc = 0.0;
    unsigned long long moved_acc = 0;
    double sc[kItemMaxOut];
    if (tid < n_out) ab_s[tid] = is_hdp ? p.alpha * p.hdp_beta[tid < K ? tid : K] : p.alpha;
    __syncthreads();
    for (int idx = tid; idx < n_out * kItemPriorN; idx += kItemThreads)
        prior_s[idx] = log((double)(idx % kItemPriorN) + ab_s[idx / kItemPriorN]);
    for (int64_t base = (int64_t)blockIdx.x * kItemThreads; base < p.n_work; base += (int64_t)gridDim.x * kItemThreads) {
        __syncthreads();
        if (tid < n_out) {
            const bool fresh = tid >= K;
            const double n = fresh ? 0.0 : (double)__ldcg(p.n_items + tid), sx = fresh ? 0.0 : __ldcg(p.sumx + tid);
            double mu = p.m0, vp = p.v0;
            if (n > 0.0) {                       // posterior, src/conjugates.jl:77-89
                const double denom = p.vv + n * p.v0;
                mu = (p.v0 * sx + p.vv * p.m0) / denom;
                vp = p.v0 * p.vv / denom;
            }
            const double dummy = vp + p.vv;      // src/conjugates.jl:98-99
            n_s[tid] = n;
            sx_s[tid] = sx;
            mu_s[tid] = mu;
            vp_s[tid] = vp;
            i2d_s[tid] = 1.0 / (2.0 * dummy);
            lc_s[tid] = 0.5 * log(2.0 * 3.141592653589793 * dummy);
        }
        __syncthreads();
        const int64_t q = base + tid;
        if (q >= p.n_work) continue;
        const int64_t i = p.item_idx ? p.item_idx[q] : (p.work_idx32 ? (int64_t)p.work_idx32[q] : p.work_base + q);
        const int zold = p.z[i];                      // -1: the item is currently unassigned (HDP pending)
        int32_t *prow = p.prior_rows + (p.row_of ? (size_t)p.row_of[p.row_by_work ? q : i] * Kpad : 0);
        const double x = p.x[i];
        double lmax = -INFINITY;
        for (int k0 = 0; k0 < n_out; k0 += 4) {
            const int4 c4 = __ldcg(reinterpret_cast<const int4 *>(prow + k0));     // rows are Kpad-aligned, padded with zeros
            const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + e;
                if (k < n_out) {
                    const bool own = (k == zold), fresh = (k >= K);
                    const int cnt = fresh ? 0 : cc[e] - (own ? 1 : 0);
                    const double prior = (cnt >= 0 && cnt < kItemPriorN) ? prior_s[k * kItemPriorN + cnt] : log((double)cnt + ab_s[k]);
                    double lp;
                    if (!own) {
                        const double dx = x - mu_s[k];
                        lp = -(dx * dx) * i2d_s[k] - lc_s[k];
                    } else {
                        lp = g1g1_logpred(n_s[k] - 1.0, sx_s[k] - x, x, p.m0, p.v0, p.vv);
                    }
                    const double s = prior + lp;
                    sc[k] = s;
                    lmax = fmax(lmax, s);
                }
            }
        }
        const int64_t uidx = (p.item_idx || p.work_idx32) ? q : i;
        double u;
        if (p.ext_uniforms) {
            u = p.ext_uniforms[uidx];
        } else {
            uint32_t o[4];
            philox((uint32_t)i, (uint32_t)((uint64_t)i >> 32), p.sweep, p.rng_stream, o);
            u = u01_double(o[0], o[1]);
        }
        double total = 0.0;
        for (int k = 0; k < n_out; ++k) {
            const double dlt = sc[k] - lmax;
            const double e = dlt < -60.0 ? 0.0 : exp(dlt);
            sc[k] = e;
            total += e;
        }
        const double target = u * total;
        int knew = n_out - 1;
        double cw = 0.0;
        for (int k = 0; k < n_out; ++k) {
            cw += sc[k];
            if (cw >= target) { knew = k; break; }
        }
        if (p.frozen) {
            if (p.draw_out) p.draw_out[q] = knew;
            continue;
        }
        const bool born = is_hdp && knew >= K;      // deferred to the serial birth pass
        const int kdst = born ? -1 : knew;
        if (kdst != zold) {
            if (zold >= 0) {
                atomicAdd(p.sumx + zold, -x);
                atomicAdd(p.n_items + zold, -1);
                if (!p.prior_is_items) atomicAdd(prow + zold, -1);
            }
            if (kdst >= 0) {
                atomicAdd(p.sumx + kdst, x);
                atomicAdd(p.n_items + kdst, 1);
                if (!p.prior_is_items) atomicAdd(prow + kdst, 1);
            }
            p.z[i] = kdst;
            if (!born) moved_acc += 1;
        }
        if (born) {
            const unsigned slot = atomicAdd(p.pending_count, 1u);
            p.pending_list[slot] = (int32_t)i;
        } else if (p.diag_sum) {
            // loglikelihood(components[kk], x) with the item seated (src/LDA.jl:131 -> src/conjugates.jl:105): from the
            // snapshot when the item stayed (the snapshot counts it), from the component's statistics when it moved
            if (knew == zold) {
                const double dx = x - mu_s[knew], a = -(dx * dx) / (2.0 * vp_s[knew]);
                diag_acc += a < -745.0 ? -INFINITY : a - 0.5 * log(2.0 * 3.141592653589793 * vp_s[knew]);   // log(exp(a) / sqrt(2 pi vp))
            } else {
                diag_acc += g1g1_loglik((double)__ldcg(p.n_items + knew), __ldcg(p.sumx + knew), x, p.m0, p.v0, p.vv);
            }
        }
    }
    if (!p.frozen) {
        if (p.diag_sum) {
            diag_acc = warp_sum(diag_acc);
            if ((threadIdx.x & 31) == 0) atomicAdd(p.diag_sum, diag_acc);
        }
        if (p.moved) {
            for (int o = 16; o > 0; o >>= 1) moved_acc += __shfl_xor_sync(kFull, moved_acc, o);
            if ((threadIdx.x & 31) == 0 && moved_acc) atomicAdd(p.moved, moved_acc);
        }
    }
}


template <int KIND, int DATUM>
__global__ void __launch_bounds__(kGenThreads, (KIND == BIAS_G1G1 || DATUM == BIAS_INT) ? 8 : 4)
generic_sweep_kernel(const GenericParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_out = p.n_out;
    double *sc = reinterpret_cast<double *>(smem_raw) + (size_t)warp * ((n_out + 3) & ~3);
    const Philox philox(p.seed);
    const int64_t n_warps = (int64_t)gridDim.x * kGenWarps;
    const bool is_hdp = p.model == BIAS_MODEL_HDP;
    const int K = p.K, Kpad = p.Kpad;
    double diag_acc = 0.0;
    unsigned long long moved_acc = 0;

    for (int64_t q = (int64_t)blockIdx.x * kGenWarps + warp; q < p.n_work; q += n_warps) {
        const int64_t i = p.item_idx ? p.item_idx[q] : (p.work_idx32 ? (int64_t)p.work_idx32[q] : p.work_base + q);
        const int zold = p.z[i];                      // -1: the item is currently unassigned (HDP pending)
        int32_t *prow = p.prior_rows + (p.row_of ? (size_t)p.row_of[p.row_by_work ? q : i] * Kpad : 0);
        const Datum d = load_datum<KIND, DATUM>(p, i, lane);
        const int epoch = (p.model == BIAS_MODEL_RCRP) ? p.row_of[i] : 0;
        const double alpha_i = (p.model == BIAS_MODEL_RCRP) ? p.rcrp_aa[epoch] : p.alpha;
        const double lmax = score_all<KIND, DATUM>(p, d, zold, K, n_out, prow, p.hdp_beta, alpha_i, sc, lane, epoch);

        const int64_t uidx = (p.item_idx || p.work_idx32) ? q : i;
        double u;
        if (p.ext_uniforms) {
            u = p.ext_uniforms[uidx];
        } else {
            uint32_t o[4];
            philox((uint32_t)i, (uint32_t)((uint64_t)i >> 32), p.sweep, p.rng_stream, o);
            u = u01_double(o[0], o[1]);
        }

        int knew;
        if (p.faithful) {
            if (p.raw_out)
                for (int k = lane; k < n_out; k += 32) p.raw_out[(size_t)q * n_out + k] = sc[k];
            __syncwarp();
            knew = draw_faithful(sc, n_out, lmax, u, lane);
            if (p.prob_out)
                for (int k = lane; k < n_out; k += 32) p.prob_out[(size_t)q * n_out + k] = sc[k];
        } else {
            knew = draw_parallel(sc, n_out, lmax, u, lane);
        }
        __syncwarp();

        if (p.frozen) {
            if (lane == 0 && p.draw_out) p.draw_out[q] = knew;
            continue;
        }

        // HDP: a draw of the new-component slot is deferred to the serial birth pass
        const bool born = (is_hdp || p.model == BIAS_MODEL_RCRP) && knew >= K;
        if (born) {
            // leave the item unassigned and queue it (also when it was already unassigned)
            if (zold >= 0) apply_move<KIND, DATUM>(p, d, zold, -1, prow, lane);
            if (lane == 0) {
                p.z[i] = -1;
                const unsigned slot = atomicAdd(p.pending_count, 1u);
                p.pending_list[slot] = (int32_t)i;
            }
        } else if (knew != zold) {
            apply_move<KIND, DATUM>(p, d, zold, knew, prow, lane);
            if (lane == 0) {
                p.z[i] = knew;
                moved_acc += 1;
            }
        }
        if (!born && p.diag_sum) {
            const double dv = diag_after<KIND, DATUM>(p, d, knew, lane);
            if (lane == 0) diag_acc += dv;
        }
    }

    if (!p.frozen && lane == 0) {
        if (p.diag_sum) atomicAdd(p.diag_sum, diag_acc);
        if (p.moved && moved_acc) atomicAdd(p.moved, moved_acc);
    }
}

#endif  // BIAS_GENERIC_KERNEL_IMPL

#endif  // __CUDACC__

}  // namespace bias
