// crf.cuh — Chinese-restaurant-franchise samplers (reference src/HDP.jl:402-709 CRF_gibbs_sampler!, src/RCRF.jl
// RCRF_gibbs_sampler!) for the three item kinds of the reference's demos: Float64 x Gaussian1DGaussian1D
// (demo_CRF_*, demo_RCRF_Gaussian1DGaussian1D.jl), word ids and Sent x MultinomialDirichlet (demo_RCRF_Multinomial*.jl).
// Included by hdp.cu inside namespace bias (uses its PhiloxStream, hdp_hyper_kernel, grid1d, compaction kernels).
//
// The reference walks the restaurants one after the other; inside a restaurant every step depends on the one
// before (tables open and close, indices shift).  Here ONE WARP owns a restaurant and performs exactly the
// reference's serial steps for it, lanes sharing the work of each step (scores over tables / dishes, shifting the
// table arrays, re-labelling the customers); restaurants run in parallel against the shared dish statistics
// (sum x, n, tables per dish) with atomics — the same cross-group staleness as the other samplers.
//   pass 1 (crf_tables_kernel)  :506-587  a table for every customer; a new table draws its dish
//   pass 2 (crf_dishes_kernel)  :593-655  a dish for every table (all its customers move together)
// Differences from the serial reference, both documented in DESIGN.md: a dish that loses its last table keeps its slot
// (weight log 0 = never chosen) until the end of the iteration instead of being spliced out at once (:531-545,
// :610-621 renumber every restaurant), and a new dish takes the next free slot with one atomic.  With a single
// restaurant and supplied uniforms the kernels reproduce the serial sampler's iteration exactly (tests/test_gpu_crf.py).
//
// Layout: tables of group j live at [group_ptr[j], group_ptr[j] + n_tbls[j]) of njt / kjt (a group has at most as
// many tables as customers); tji[i] is the table of customer i within its group; mm[k] = tables serving dish k.
#pragma once

struct CrfParams {
    const int64_t *group_ptr;
    int64_t G;
    const double *x;
    int32_t *z, *tji, *njt, *kjt, *n_tbls, *mm;
    int32_t *n_items;           // customers per dish (component statistics, shared with the other samplers)
    double *sumx;
    // MultinomialDirichlet items
    const int32_t *word;
    const int64_t *sent_ptr;
    const int32_t *sent_word, *sent_count;
    int32_t *n_wk, *n_k;
    double beta;
    int64_t V;
    int32_t *d_KK;              // dish slots in use (grows by atomics during a pass)
    int32_t Kmax;
    double m0, v0, vv, aa, gg;
    uint64_t seed;
    uint32_t sweep;
    const double *ext_uniforms; // nullable: [3N] = per item slot i: table draw, dish-of-a-new-table draw; per table slot: dish draw
    int *lock;                  // births are decided one at a time (see crf_lock)
    int *err;                   // 2: dish slots exhausted
    unsigned long long *moved;
    int32_t max_len;            // longest group (capacity of the per-warp staging buffers)
    // RCRF (src/RCRF.jl): the launch covers the restaurants [g0, g1) of epoch t; mm is [T][Kpad] and the dish prior
    // couples epochs t-1, t, t+1.  Plain CRF: rcrf = 0, T = 1, t = 0, g0 = 0, g1 = G.
    int32_t rcrf, T, t, Kpad;
    int64_t g0, g1;
    int32_t *tmp;               // [N] scratch of the table-joining pass
};

constexpr int kCrfWarps = 4;
constexpr int kCrfSnapT = 32;          // tables of a restaurant whose dishes' predictive constants are cached (pass 1, Float64 items)
__host__ __device__ inline size_t crf_warp_doubles(int max_len, int Kmax) {
    const size_t n = (size_t)(max_len > Kmax ? max_len : Kmax) + 2;
    return n + (size_t)max_len + (3 * (size_t)max_len + 1) / 2 + 7 * (size_t)kCrfSnapT;
}
__host__ __device__ inline size_t crf_smem_bytes(int max_len, int Kmax) {
    // per warp: scores double[max(max_len, Kmax) + 2] + staged values double[max_len] + the restaurant's tji / njt / kjt
    // (3 x int32[max_len], pass 1)
    // + log(n) for n = 0 .. max_len, shared by the CTA (pass 1: the weight of a table with n customers)
    return ((size_t)kCrfWarps * crf_warp_doubles(max_len, Kmax) + (size_t)max_len + 1) * sizeof(double);
}

// first-appearance tables from z (src/HDP.jl:441-470): kjt[jj] = unique(zz[jj])
__global__ void crf_init_kernel(CrfParams p) {
    const int64_t j = p.g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.g1) return;
    const int64_t p0 = p.group_ptr[j], n = p.group_ptr[j + 1] - p0;
    int nt = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int kk = p.z[p0 + i];
        int t = -1;
        for (int q = 0; q < nt; ++q)
            if (p.kjt[p0 + q] == kk) { t = q; break; }
        if (t < 0) {
            t = nt++;
            p.kjt[p0 + t] = kk;
            p.njt[p0 + t] = 0;
            atomicAdd(p.mm + (size_t)p.t * p.Kpad + kk, 1);
        }
        p.tji[p0 + i] = t;
        p.njt[p0 + t] += 1;
    }
    p.n_tbls[j] = nt;
}

__device__ __forceinline__ double crf_draw_uniform(const CrfParams &p, int64_t idx, int which) {
    if (p.ext_uniforms) return p.ext_uniforms[3 * idx + which];
    PhiloxStream rs(p.seed, (uint32_t)idx, (uint32_t)((uint64_t)idx >> 32), p.sweep * 8u + 5u + (uint32_t)which);
    return rs.u01();
}

// max over the warp of the scores a lane wrote, then the draw (same arithmetic as the generic kernel)
__device__ __forceinline__ int crf_draw(double *sc, int n_out, double u, int lane) {
    double lmax = -INFINITY;
    for (int k = lane; k < n_out; k += 32) lmax = fmax(lmax, sc[k]);
    lmax = warp_max(lmax);
    __syncwarp();
    return draw_parallel(sc, n_out, lmax, u, lane);
}

// ---- the conjugate, by item kind -------------------------------------------------------------------------
// logpredictive(components[k], item) (src/conjugates.jl:92-103, :180, :181-198); k < 0: the empty prototype.
// own: the item itself is still counted in dish k's statistics and is taken out arithmetically (delitem!).
template <int DATUM>
__device__ __forceinline__ double crf_logpred(const CrfParams &p, int k, int64_t idx, double x, bool own = false) {
    if (DATUM == BIAS_F64) {         // x = p.x[idx], loaded once by the caller (a re-load after every fence goes to L2)
        double cn = k < 0 ? 0.0 : (double)__ldcg(p.n_items + k), sx = k < 0 ? 0.0 : __ldcg(p.sumx + k);
        if (own) { cn -= 1.0; sx -= x; }
        return g1g1_logpred(cn, sx, x, p.m0, p.v0, p.vv);
    } else if (DATUM == BIAS_INT) {
        const double nw = k < 0 ? 0.0 : (double)(__ldcg(p.n_wk + (size_t)p.word[idx] * p.Kpad + k) - (own ? 1 : 0));
        const double nk = k < 0 ? 0.0 : (double)(__ldcg(p.n_k + k) - (own ? 1 : 0));
        return log((p.beta + nw) / (p.beta * (double)p.V + nk));
    } else {
        const double bv = p.beta * (double)p.V;
        double nk = k < 0 ? 0.0 : (double)__ldcg(p.n_k + k);
        double m = 0.0, acc = 0.0, coef = 0.0;
        for (int64_t t = p.sent_ptr[idx]; t < p.sent_ptr[idx + 1]; ++t) {
            const double ct = (double)p.sent_count[t];
            const double nw = k < 0 ? 0.0 : (double)__ldcg(p.n_wk + (size_t)p.sent_word[t] * p.Kpad + k) - (own ? ct : 0.0);
            acc += lgamma(p.beta + nw + ct) - lgamma(p.beta + nw);
            coef -= lgamma(ct + 1.0);
            m += ct;
        }
        if (own) nk -= m;
        return (lgamma(m + 1.0) + coef) + lgamma(bv + nk) - lgamma(bv + nk + m) + acc;       // src/conjugates.jl:193-195
    }
}
// additem! (sign = +1) / delitem! (sign = -1) of one item on dish k; executed by ONE lane
template <int DATUM>
__device__ __forceinline__ void crf_item_move(const CrfParams &p, int k, int64_t idx, int sign, double x) {
    atomicAdd(p.n_items + k, sign);
    if (DATUM == BIAS_F64) {
        atomicAdd(p.sumx + k, (double)sign * x);
    } else if (DATUM == BIAS_INT) {
        atomicAdd(p.n_wk + (size_t)p.word[idx] * p.Kpad + k, sign);
        atomicAdd(p.n_k + k, sign);
    } else {
        int m = 0;
        for (int64_t t = p.sent_ptr[idx]; t < p.sent_ptr[idx + 1]; ++t) {
            atomicAdd(p.n_wk + (size_t)p.sent_word[t] * p.Kpad + k, sign * p.sent_count[t]);
            m += p.sent_count[t];
        }
        atomicAdd(p.n_k + k, sign * m);
    }
}

// Restaurants that meet a new atom at the same moment would each open a dish for it (hundreds of duplicates in the
// first iteration).  A warp whose draw says "new dish" therefore takes a global lock, scores the menu again — it now
// holds whatever the previous lock holders opened — and opens a dish only if the draw (same uniform) still says so.
// With one restaurant this is the serial sampler step for step.
// Only births are serialised: while a lock holder scores the menu, other warps keep moving customers between existing
// dishes with pairs of independent atomics (n, then the sum), so a scorer can catch a dish between the two — its mean
// off by one item, which for a dish of a handful of items is far outside the predictive and makes the holder open
// a duplicate.  A dish is therefore opened only if kCrfBirthLooks consecutive scorings (same uniform) all say
// "new"; with one restaurant every look gives the same answer, so the serial trajectory is unchanged.
constexpr int kCrfBirthLooks = 6;
__device__ __forceinline__ void crf_lock(const CrfParams &p, int lane) {
    if (lane == 0)
        while (atomicCAS(p.lock, 0, 1) != 0) __nanosleep(200);
    __syncwarp();
    __threadfence();
}
__device__ __forceinline__ void crf_unlock(const CrfParams &p, int lane) {
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(p.lock, 0);
}

// ---- the prior weight of a dish for a new table / a table that re-draws its dish ----
// CRF: log(mm[k]) (src/HDP.jl:566, :634).  RCRF, epoch t (src/RCRF.jl:303-327, :534-556, :762-768): dishes served in
// epochs t-1..t are eligible with weight log(mm[t-1,k] + mm[t,k]) plus a look-ahead factor over the tables of epoch
// t+1.  The reference evaluates that factor as sum(L1) - L2 with L1[k] = sum_{i < mm[t+1,k]} log(i + mm[t,k] + gg/K')
// and the candidate's entry replaced by sum_i log(i + mm[t,k] + 1 + aa/K') (gg in one, aa in the other: reproduced);
// everything but the candidate's own two sums is common to all outcomes, and the sums are lgamma differences.
struct CrfLook {
    double a, g;        // aa[t] / K', gg[t] / K'
    bool on;            // epoch t has a successor
};
__device__ __forceinline__ CrfLook crf_look(const CrfParams &p, int KK, int lane) {
    CrfLook L{0.0, 0.0, false};
    if (!p.rcrf || p.t >= p.T - 1) return L;
    const int32_t *mt = p.mm + (size_t)p.t * p.Kpad;
    int alive = 0;
    for (int k = lane; k < KK; k += 32) alive += (__ldcg(mt + k) + __ldcg(mt + p.Kpad + k)) > 0 ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(kFull, alive, o);
    L.a = p.aa / (double)(alive + 1);
    L.g = p.gg / (double)(alive + 1);
    L.on = true;
    return L;
}
__device__ __forceinline__ double crf_dish_prior(const CrfParams &p, const CrfLook &L, int k) {
    const int32_t *mt = p.mm + (size_t)p.t * p.Kpad;
    if (!p.rcrf) return log((double)__ldcg(mt + k));
    const int m = __ldcg(mt + k);
    const int cnt = m + (p.t > 0 ? __ldcg(mt - p.Kpad + k) : 0);
    if (cnt <= 0) return -INFINITY;
    double v = log((double)cnt);
    if (L.on) {
        const int nx = __ldcg(mt + p.Kpad + k);
        if (nx > 0) {
            const double dm = (double)m, dn = (double)nx;
            v += (lgamma(dm + 1.0 + L.a + dn) - lgamma(dm + 1.0 + L.a)) - (lgamma(dm + L.g + dn) - lgamma(dm + L.g));
        }
    }
    return v;
}

// The number of dishes, read by ONE lane and broadcast: other warps open dishes concurrently, and lanes that read
// different values would run the scoring loops and the draw's warp collectives with different trip counts.
__device__ __forceinline__ int crf_read_KK(const CrfParams &p, int lane) {
    int KK = 0;
    if (lane == 0) KK = __ldcg(p.d_KK);
    return __shfl_sync(kFull, KK, 0);
}

// a new dish: the next free slot (one atomic); -1 when the slots are exhausted
__device__ __forceinline__ int crf_new_dish(const CrfParams &p, int lane) {
    int k = 0;
    if (lane == 0) {
        k = atomicAdd(p.d_KK, 1);
        if (k + 2 > p.Kmax) {
            atomicSub(p.d_KK, 1);
            *p.err = 2;
            k = -1;
        }
    }
    return __shfl_sync(kFull, k, 0);
}

// ---- pass 1: a table for every customer (src/HDP.jl:506-587) ----
// The restaurant's table arrays (tji, njt, kjt) and, for Float64 items, its values are staged in shared memory for the
// whole visit: a step's only dependent trips to L2 are the dish statistics it scores against and the atomics that move
// the customer.  (The first version kept the arrays in global memory and fenced twice per customer: 9 + 8 of 20 stalled
// warps per issue were membar and long-scoreboard stalls, profiles/r02_crf_ncu_summary.md.)  z is not read: a customer's
// dish is its table's, z[i] == kjt[tji[i]].
template <int DATUM>
__global__ void __launch_bounds__(kCrfWarps * 32, 5) crf_tables_kernel(CrfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t n_sc = (size_t)(p.max_len > p.Kmax ? p.max_len : p.Kmax) + 2;
    double *sc = reinterpret_cast<double *>(smem_raw) + (size_t)warp * crf_warp_doubles(p.max_len, p.Kmax);
    double *xs = sc + n_sc;
    int32_t *tji = reinterpret_cast<int32_t *>(xs + p.max_len), *njt = tji + p.max_len, *kjt = njt + p.max_len;
    // Float64 items: per table, the constants of its dish's predictive (src/conjugates.jl:92-103) — without the customer being
    // scored (mu, 1 / (2 (v_post + vv)), log-normaliser) and with it taken out of the dish (mean = oA - oB x).  They are rebuilt
    // from the live statistics only when this restaurant changed something (a table opened or closed, a customer changed dish);
    // a customer that stays costs three multiply-adds per table instead of two divisions and a log.
    double *c_mu = sc + crf_warp_doubles(p.max_len, p.Kmax) - 7 * kCrfSnapT, *c_i2d = c_mu + kCrfSnapT, *c_lc = c_i2d + kCrfSnapT,
           *c_oA = c_lc + kCrfSnapT, *c_oB = c_oA + kCrfSnapT, *c_oi2d = c_oB + kCrfSnapT, *c_olc = c_oi2d + kCrfSnapT;
    double *logn = reinterpret_cast<double *>(smem_raw) + (size_t)kCrfWarps * crf_warp_doubles(p.max_len, p.Kmax);
    for (int q = threadIdx.x; q <= p.max_len; q += blockDim.x) logn[q] = log((double)q);
    __syncthreads();
    const double i2d0 = 1.0 / (2.0 * (p.v0 + p.vv)), lc0 = 0.5 * log(2.0 * 3.141592653589793 * (p.v0 + p.vv));   // the empty prototype
    const int64_t n_warps = (int64_t)gridDim.x * kCrfWarps;
    unsigned long long moved = 0;
    int32_t *const mm_t = p.mm + (size_t)p.t * p.Kpad;
    const double log_aa = log(p.aa), log_gg = log(p.gg);
    for (int64_t j = p.g0 + (int64_t)blockIdx.x * kCrfWarps + warp; j < p.g1; j += n_warps) {
        const int64_t p0 = p.group_ptr[j];
        const int n = (int)(p.group_ptr[j + 1] - p0);
        int nt = p.n_tbls[j];
        bool dirty = true;                   // the cached constants do not describe the current tables / statistics
        __syncwarp();
        for (int q = lane; q < n; q += 32) {
            tji[q] = p.tji[p0 + q];
            if (q < nt) {
                njt[q] = p.njt[p0 + q];
                kjt[q] = p.kjt[p0 + q];
            }
            if (DATUM == BIAS_F64) xs[q] = p.x[p0 + q];
        }
        __syncwarp();
        for (int ii = 0; ii < n; ++ii) {
            const int64_t i = p0 + ii;
            int tbl = tji[ii];
            const int kk = kjt[tbl];
            const double x = (DATUM == BIAS_F64) ? xs[ii] : 0.0;
            // remove the customer (:511-515) — from its table here, from its dish ARITHMETICALLY: the statistics of dish kk
            // keep counting it until it is known to leave (every score of kk below takes it out again), so a customer
            // that stays with its dish — nearly all of them once the chain has settled — costs no atomics on the
            // handful of addresses every restaurant shares
            const int left = njt[tbl] - 1;
            __syncwarp();
            if (lane == 0) njt[tbl] = left;
            if (left == 0) {
                // the table empties: it is removed and the tables after it move down (:517-529); the dish keeps its
                // slot even if this was its last table
                if (lane == 0) atomicAdd(mm_t + kk, -1);
                __syncwarp();
                for (int b = tbl; b < nt - 1; b += 32) {
                    const int q = b + lane;
                    int a0 = 0, a1 = 0;
                    if (q < nt - 1) { a0 = kjt[q + 1]; a1 = njt[q + 1]; }
                    __syncwarp();
                    if (q < nt - 1) { kjt[q] = a0; njt[q] = a1; }
                    __syncwarp();
                }
                nt -= 1;
                dirty = true;
                for (int q = lane; q < n; q += 32)
                    if (tji[q] > tbl) tji[q] -= 1;
            }
            // the table count taken off mm must have reached what the lanes are about to read
            __threadfence_block();
            __syncwarp();
            // scores of the tables and of a new one (:549-555)
            if (DATUM == BIAS_F64 && nt <= kCrfSnapT) {
                if (dirty) {
                    if (lane < nt) {
                        const int k = kjt[lane];
                        const double cn = (double)__ldcg(p.n_items + k), sx = __ldcg(p.sumx + k);
                        double mu = p.m0, vp = p.v0;
                        if (cn > 0.0) {
                            const double denom = p.vv + cn * p.v0;
                            mu = (p.v0 * sx + p.vv * p.m0) / denom;
                            vp = p.v0 * p.vv / denom;
                        }
                        c_mu[lane] = mu;
                        c_i2d[lane] = 1.0 / (2.0 * (vp + p.vv));
                        c_lc[lane] = 0.5 * log(2.0 * 3.141592653589793 * (vp + p.vv));
                        double oA = p.m0, oB = 0.0, vp1 = p.v0;          // the dish without one of its customers
                        if (cn - 1.0 > 0.0) {
                            const double d1 = p.vv + (cn - 1.0) * p.v0;
                            oA = (p.v0 * sx + p.vv * p.m0) / d1;
                            oB = p.v0 / d1;
                            vp1 = p.v0 * p.vv / d1;
                        }
                        c_oA[lane] = oA;
                        c_oB[lane] = oB;
                        c_oi2d[lane] = 1.0 / (2.0 * (vp1 + p.vv));
                        c_olc[lane] = 0.5 * log(2.0 * 3.141592653589793 * (vp1 + p.vv));
                    }
                    dirty = false;
                    __syncwarp();
                }
                if (lane < nt) {
                    const bool own = kjt[lane] == kk;
                    const double dx = x - (own ? c_oA[lane] - c_oB[lane] * x : c_mu[lane]);
                    sc[lane] = logn[njt[lane]] + (-(dx * dx) * (own ? c_oi2d[lane] : c_i2d[lane]) - (own ? c_olc[lane] : c_lc[lane]));
                }
                if (lane == (nt & 31)) sc[nt] = log_aa + (-((x - p.m0) * (x - p.m0)) * i2d0 - lc0);
            } else {
                for (int t = lane; t < nt; t += 32) sc[t] = logn[njt[t]] + crf_logpred<DATUM>(p, kjt[t], i, x, kjt[t] == kk);
                if (lane == (nt & 31)) sc[nt] = log_aa + crf_logpred<DATUM>(p, -1, i, x);
            }
            __syncwarp();
            tbl = crf_draw(sc, nt + 1, crf_draw_uniform(p, i, 0), lane);
            if (tbl == nt) {
                dirty = true;
                // a new table draws its dish from the menu (:559-581)
                const int KK = crf_read_KK(p, lane);
                const CrfLook L = crf_look(p, KK, lane);
                for (int k = lane; k < KK; k += 32) {
                    const double pr = crf_dish_prior(p, L, k);
                    sc[k] = (pr == -INFINITY) ? pr : pr + crf_logpred<DATUM>(p, k, i, x, k == kk);
                }
                if (lane == (KK & 31)) sc[KK] = log_gg + crf_logpred<DATUM>(p, -1, i, x);
                __syncwarp();
                const double u_dish = crf_draw_uniform(p, i, 1);
                int knew = crf_draw(sc, KK + 1, u_dish, lane);
                const bool locked = knew == KK;
                if (locked) {
                    crf_lock(p, lane);
                    const int KK2 = crf_read_KK(p, lane);
                    knew = KK2;
                    for (int look = 0; look < kCrfBirthLooks && knew == KK2; ++look) {      // see kCrfBirthLooks
                        const CrfLook L2 = crf_look(p, KK2, lane);
                        for (int k = lane; k < KK2; k += 32) {
                            const double pr = crf_dish_prior(p, L2, k);
                            sc[k] = (pr == -INFINITY) ? pr : pr + crf_logpred<DATUM>(p, k, i, x, k == kk);
                        }
                        if (lane == 0) sc[KK2] = log_gg + crf_logpred<DATUM>(p, -1, i, x);
                        __syncwarp();
                        knew = crf_draw(sc, KK2 + 1, u_dish, lane);
                        __syncwarp();
                    }
                    if (knew == KK2) {
                        knew = crf_new_dish(p, lane);
                        if (knew < 0) knew = KK2 - 1;
                    }
                }
                if (lane == 0) {
                    kjt[tbl] = knew;
                    njt[tbl] = 0;
                    atomicAdd(mm_t + knew, 1);
                }
                nt += 1;
                __syncwarp();
                if (locked) {
                    // seat the customer before the lock is released: the next lock holder must see a dish somebody eats
                    if (lane == 0) {
                        tji[ii] = tbl;
                        njt[tbl] = 1;
                        if (knew != kk) {
                            crf_item_move<DATUM>(p, knew, i, 1, x);
                            crf_item_move<DATUM>(p, kk, i, -1, x);
                            p.z[i] = knew;
                            moved += 1;
                        }
                    }
                    crf_unlock(p, lane);
                    continue;
                }
            }
            const int kn = kjt[tbl];
            if (kn != kk) dirty = true;
            __syncwarp();
            if (lane == 0) {
                tji[ii] = tbl;
                njt[tbl] += 1;
                if (kn != kk) {
                    crf_item_move<DATUM>(p, kn, i, 1, x);
                    crf_item_move<DATUM>(p, kk, i, -1, x);
                    p.z[i] = kn;
                    moved += 1;
                }
            }
            __syncwarp();
        }
        for (int q = lane; q < n; q += 32) {
            p.tji[p0 + q] = tji[q];
            if (q < nt) {
                p.njt[p0 + q] = njt[q];
                p.kjt[p0 + q] = kjt[q];
            }
        }
        if (lane == 0) p.n_tbls[j] = nt;
    }
    if (lane == 0 && moved) atomicAdd(p.moved, moved);
}

// ---- pass 2: a dish for every table (src/HDP.jl:593-655) ----
// score of dish k (k < 0: a new dish) for a whole table: the sum of its customers' individual predictives (:633-643)
template <int DATUM>
__device__ __forceinline__ double crf_table_score(const CrfParams &p, int k, const double *xs, const int32_t *ids, int64_t p0, int nti) {
    double s = 0.0;
    if (DATUM == BIAS_F64) {
        // the dish's posterior predictive (src/conjugates.jl:92-103) is the same Gaussian for every customer of the table: its mean,
        // 1 / (2 (v_post + vv)) and log-normaliser once, then one multiply-add per customer
        const double cn = k < 0 ? 0.0 : (double)__ldcg(p.n_items + k), sx = k < 0 ? 0.0 : __ldcg(p.sumx + k);
        double mu = p.m0, vp = p.v0;
        if (cn > 0.0) {
            const double denom = p.vv + cn * p.v0;
            mu = (p.v0 * sx + p.vv * p.m0) / denom;
            vp = p.v0 * p.vv / denom;
        }
        const double dummy = vp + p.vv, i2d = 1.0 / (2.0 * dummy), lc = 0.5 * log(2.0 * 3.141592653589793 * dummy);
        for (int q = 0; q < nti; ++q) {
            const double dx = xs[q] - mu;
            s += -(dx * dx) * i2d - lc;
        }
    } else {
        for (int q = 0; q < nti; ++q) s += crf_logpred<DATUM>(p, k, p0 + ids[q], 0.0);
    }
    return s;
}

template <int DATUM>
__global__ void __launch_bounds__(kCrfWarps * 32) crf_dishes_kernel(CrfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t n_sc = (size_t)(p.max_len > p.Kmax ? p.max_len : p.Kmax) + 2;
    double *sc = reinterpret_cast<double *>(smem_raw) + (size_t)warp * crf_warp_doubles(p.max_len, p.Kmax);
    double *xs = sc + n_sc;                                   // Float64 items: the values of the table's customers
    int32_t *ids = reinterpret_cast<int32_t *>(xs);           // other items: their positions in the restaurant
    const int64_t n_warps = (int64_t)gridDim.x * kCrfWarps;
    unsigned long long moved = 0;
    int32_t *const mm_t = p.mm + (size_t)p.t * p.Kpad;
    for (int64_t j = p.g0 + (int64_t)blockIdx.x * kCrfWarps + warp; j < p.g1; j += n_warps) {
        const int64_t p0 = p.group_ptr[j];
        const int n = (int)(p.group_ptr[j + 1] - p0);
        int32_t *kjt = p.kjt + p0;
        const int32_t *tji = p.tji + p0;
        const int nt = p.n_tbls[j];
        for (int tbl = 0; tbl < nt; ++tbl) {
            const int kk = kjt[tbl];
            // the table's customers, in index order (tidx = find(x -> x==tbl, tji[jj]), :602)
            int nti = 0;
            double sum = 0.0;
            for (int b = 0; b < n; b += 32) {
                const int q = b + lane;
                const bool mine = q < n && tji[q] == tbl;
                const unsigned bal = __ballot_sync(kFull, mine);
                if (mine) {
                    const int slot = nti + __popc(bal & ((1u << lane) - 1u));
                    if (DATUM == BIAS_F64) {
                        const double xv = p.x[p0 + q];
                        xs[slot] = xv;
                        sum += xv;
                    } else {
                        ids[slot] = q;
                    }
                }
                nti += __popc(bal);
            }
            __syncwarp();
            // the table leaves its dish (:599, :626-628); the table count goes first: a dish must not look alive (prior
            // log(mm)) while its statistics are already gone
            if (lane == 0) atomicAdd(mm_t + kk, -1);
            if (DATUM == BIAS_F64) {
                sum = warp_sum(sum);
                if (lane == 0) {
                    atomicAdd(p.n_items + kk, -nti);
                    atomicAdd(p.sumx + kk, -sum);
                }
            } else {
                for (int q = lane; q < nti; q += 32) crf_item_move<DATUM>(p, kk, p0 + ids[q], -1, 0.0);
            }
            __threadfence();
            __syncwarp();
            const int KK = crf_read_KK(p, lane);
            const CrfLook L = crf_look(p, KK, lane);
            for (int k = lane; k <= KK; k += 32) {
                double s = (k < KK) ? crf_dish_prior(p, L, k) : log(p.gg);
                if (s != -INFINITY) s += crf_table_score<DATUM>(p, k < KK ? k : -1, xs, ids, p0, nti);
                sc[k] = s;
            }
            __syncwarp();
            const double u_dish = crf_draw_uniform(p, p0 + tbl, 2);
            int knew = crf_draw(sc, KK + 1, u_dish, lane);
            const bool locked = knew == KK;
            if (locked) {
                crf_lock(p, lane);
                const int KK2 = crf_read_KK(p, lane);
                knew = KK2;
                for (int look = 0; look < kCrfBirthLooks && knew == KK2; ++look) {          // see kCrfBirthLooks
                    const CrfLook L2 = crf_look(p, KK2, lane);
                    for (int k = lane; k <= KK2; k += 32) {
                        double s_ = (k < KK2) ? crf_dish_prior(p, L2, k) : log(p.gg);
                        if (s_ != -INFINITY) s_ += crf_table_score<DATUM>(p, k < KK2 ? k : -1, xs, ids, p0, nti);
                        sc[k] = s_;
                    }
                    __syncwarp();
                    knew = crf_draw(sc, KK2 + 1, u_dish, lane);
                    __syncwarp();
                }
                if (knew == KK2) {
                    knew = crf_new_dish(p, lane);
                    if (knew < 0) knew = kk;
                }
            }
            if (lane == 0) {
                kjt[tbl] = knew;
                atomicAdd(mm_t + knew, 1);
                if (knew != kk) moved += (unsigned long long)nti;
            }
            if (DATUM == BIAS_F64) {
                if (lane == 0) {
                    atomicAdd(p.n_items + knew, nti);
                    atomicAdd(p.sumx + knew, sum);
                }
            } else {
                for (int q = lane; q < nti; q += 32) crf_item_move<DATUM>(p, knew, p0 + ids[q], 1, 0.0);
            }
            if (locked) crf_unlock(p, lane);
            if (knew != kk)
                for (int q = lane; q < n; q += 32)
                    if (tji[q] == tbl) p.z[p0 + q] = knew;
            __threadfence();
            __syncwarp();
        }
    }
    if (lane == 0 && moved) atomicAdd(p.moved, moved);
}

// RCRF join_tables (src/RCRF.jl:351-377): the tables of a restaurant that serve the same dish become one table
// (first-appearance order of the dishes).  One thread per restaurant; the compaction is done in place.
__global__ void crf_join_kernel(CrfParams p) {
    const int64_t j = p.g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.g1) return;
    const int64_t p0 = p.group_ptr[j];
    const int n = (int)(p.group_ptr[j + 1] - p0), nt = p.n_tbls[j];
    int32_t *kjt = p.kjt + p0, *njt = p.njt + p0, *tji = p.tji + p0, *map = p.tmp + p0;
    int nu = 0;
    for (int q = 0; q < nt; ++q) {
        const int k = kjt[q], cnt = njt[q];
        int u = -1;
        for (int r = 0; r < nu; ++r)
            if (kjt[r] == k) { u = r; break; }
        if (u < 0) {
            u = nu++;
            kjt[u] = k;
            njt[u] = cnt;
        } else {
            njt[u] += cnt;
            atomicAdd(p.mm + (size_t)p.t * p.Kpad + k, -1);
        }
        map[q] = u;
    }
    if (nu != nt) {
        for (int q = 0; q < n; ++q) tji[q] = map[tji[q]];
        p.n_tbls[j] = nu;
    }
}

// diagnostic of src/HDP.jl:669-673: sum of logpredictive(components[zz], xx) with every customer seated
template <int DATUM>
__global__ void crf_diag_kernel(CrfParams p, int64_t N, double *out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        acc += crf_logpred<DATUM>(p, p.z[i], i, (DATUM == BIAS_F64) ? p.x[i] : 0.0);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void crf_sum_tables_kernel(const int32_t *n_tbls, int64_t G, unsigned long long *out) {
    unsigned long long acc = 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < G; j += (int64_t)gridDim.x * blockDim.x) acc += (unsigned)n_tbls[j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
