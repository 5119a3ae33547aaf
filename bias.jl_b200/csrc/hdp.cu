// hdp.cu — direct-assignment HDP sampler (reference src/HDP.jl:166-399) on the GPU.
//
// One sweep =
//   1. parallel assignment pass over all items with K+1 outcomes (generic_sweep_kernel,
//      src/HDP.jl:285-317).  A draw of the (K+1)-th outcome cannot be resolved in parallel —
//      a birth changes K and splits the stick for everyone after it (src/HDP.jl:320-331) —
//      so such items are left unassigned and queued;
//   2. serial birth pass: one warp walks a bounded slice of the queue with the reference's exact
//      sequential semantics (K+1 conditional on the live state, b ~ Beta(1, gamma), stick split);
//      the rest of the queue is then re-sampled in parallel against the enlarged component set,
//      and so on until the queue is empty;
//   3. prune: components that lost their last item are removed, every zz above them shifts down
//      and their stick returns to the unrepresented mass (src/HDP.jl:291-307, 248-264) — done
//      once per sweep as a column compaction instead of at each death;
//   4. n_internals x { table counts from the device-resident Stirling table (src/HDP.jl:348-361),
//      beta ~ Dirichlet (:362-364, host, K+1 gamma draws), alpha0 / gamma resampling
//      (:369-372 -> :124-159; the O(groups) auxiliary draws run on the device) }.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <cooperative_groups.h>
#include "model.hpp"
#include "generic.cuh"

namespace bias {

// ---------------------------------------------------------------------------------------
// Stirling table: normalised log unsigned Stirling numbers of the first kind, packed lower
// triangle, row n (1-based) at offset n(n-1)/2.  s(n+1,m) = n s(n,m) + s(n,m-1) in log space;
// every row is shifted so that its maximum is 0 (any row scaling cancels in lognormalize!).
// Regenerates the reference's missing StirlingNums_10K.mat (src/HDP.jl:207-209).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) stirling_build_kernel(double *tab, int nmax) {
    __shared__ double red[32];
    __shared__ double row_max;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) tab[0] = 0.0;
    __syncthreads();
    for (int n = 1; n < nmax; ++n) {
        const double *prev = tab + (size_t)n * (n - 1) / 2;
        double *cur = tab + (size_t)(n + 1) * n / 2;
        const double logn = log((double)n);
        double mx = -INFINITY;
        for (int m = tid + 1; m <= n + 1; m += 1024) {
            const double a = (m <= n) ? logn + prev[m - 1] : -INFINITY;
            const double b = (m >= 2) ? prev[m - 2] : -INFINITY;
            const double hi = a > b ? a : b, lo = a > b ? b : a;
            const double v = (lo == -INFINITY) ? hi : hi + log1p(exp(lo - hi));
            cur[m - 1] = v;
            mx = fmax(mx, v);
        }
        mx = warp_max(mx);
        if (lane == 0) red[wid] = mx;
        __syncthreads();
        if (wid == 0) {
            double v = red[lane];
            v = warp_max(v);
            if (lane == 0) row_max = v;
        }
        __syncthreads();
        const double rm = row_max;
        for (int m = tid + 1; m <= n + 1; m += 1024) cur[m - 1] -= rm;
        __syncthreads();
    }
}

int ensure_stirling(bias_ctx *ctx, int32_t rows) {
    if (rows < 1) rows = 1;
    if (ctx->stirling && ctx->stirling_rows >= rows) return BIAS_OK;
    if (rows > 10000) rows = 10000;             // the reference's table is 10k x 10k (src/HDP.jl:207)
    if (ctx->stirling && ctx->stirling_rows >= rows) return BIAS_OK;
    BIAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->stirling) cudaFree(ctx->stirling);
    ctx->stirling = nullptr;
    ctx->stirling_rows = 0;
    const size_t n = (size_t)rows * (rows + 1) / 2;
    BIAS_CUDA_TRY(cudaMalloc((void **)&ctx->stirling, n * sizeof(double)));
    stirling_build_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->stirling, rows);
    BIAS_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    ctx->stirling_rows = rows;
    return BIAS_OK;
}

// ---------------------------------------------------------------------------------------
// small counter-based random stream for device-side Gamma / Beta draws
// ---------------------------------------------------------------------------------------
struct PhiloxStream {
    Philox ph;
    uint32_t c0, c1, c2, ctr;
    uint32_t buf[4];
    int have;
    __device__ PhiloxStream(uint64_t seed, uint32_t a, uint32_t b, uint32_t c) : ph(seed), c0(a), c1(b), c2(c), ctr(0), have(0) {}
    __device__ double u01() {
        if (have < 2) {
            ph(c0, c1, c2, ctr++, buf);
            have = 4;
        }
        const double u = u01_double(buf[have - 1], buf[have - 2]);
        have -= 2;
        return u;
    }
    __device__ double normal() {
        const double u1 = u01(), u2 = u01();
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
    __device__ double gamma(double shape) {     // Marsaglia-Tsang
        double boost = 1.0;
        if (shape < 1.0) {
            boost = pow(u01(), 1.0 / shape);
            shape += 1.0;
        }
        const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (int it = 0; it < 64; ++it) {
            double x, v;
            do {
                x = normal();
                v = 1.0 + c * x;
            } while (v <= 0.0);
            v = v * v * v;
            const double u = u01();
            if (u < 1.0 - 0.0331 * x * x * x * x || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
        }
        return boost * d;
    }
};

// ---------------------------------------------------------------------------------------
// table counts: M_jk ~ p(m | n_jk, alpha0 beta_k) ∝ s(n_jk, m) (alpha0 beta_k)^m   (src/HDP.jl:348-361)
// one warp per group; sums of M over groups accumulate in tables[k], the grand total in tables[Kpad]
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int table_draw(const double *row, int n, double la, double u, int lane, bool faithful,
                                          double *prob_out) {
    // scores v_m = row[m-1] + m*la, m = 1..n
    double lmax = -INFINITY;
    for (int m = lane + 1; m <= n; m += 32) lmax = fmax(lmax, row[m - 1] + (double)m * la);
    lmax = warp_max(lmax);
    if (faithful) {
        // lane 0 alone, reference operation order (lognormalize! then sample)
        int res = 1;
        if (lane == 0) {
            double tot = 0.0;
            // Base.sum association for n <= 1024 is left to right; above that pairwise halves
            if (n <= 1024) {
                for (int m = 1; m <= n; ++m) {
                    const double e = exp(row[m - 1] + (double)m * la - lmax);
                    prob_out[m - 1] = e;
                    tot = (m == 1) ? e : tot + e;
                }
            } else {
                for (int m = 1; m <= n; ++m) prob_out[m - 1] = exp(row[m - 1] + (double)m * la - lmax);
                tot = jsum_dev(prob_out, 0, n - 1);
            }
            for (int m = 1; m <= n; ++m) prob_out[m - 1] /= tot;
            int idx = 0;
            double cw = prob_out[0];
            while (cw < u && idx < n - 1) {
                idx += 1;
                cw += prob_out[idx];
            }
            res = idx + 1;
        }
        return __shfl_sync(kFull, res, 0);
    }
    double total = 0.0;
    for (int b = 0; b < n; b += 32) {
        const int m = b + lane + 1;
        const double e = (m <= n) ? exp(row[m - 1] + (double)m * la - lmax) : 0.0;
        total += warp_sum(e);
    }
    const double target = u * total;
    double base = 0.0;
    int res = n;
    for (int b = 0; b < n; b += 32) {
        const int m = b + lane + 1;
        const double e = (m <= n) ? exp(row[m - 1] + (double)m * la - lmax) : 0.0;
        const double incl = base + warp_incl_scan(e, lane);
        const unsigned bal = __ballot_sync(kFull, (m <= n) && (incl >= target));
        if (bal) {
            res = b + __ffs(bal);
            break;
        }
        base = __shfl_sync(kFull, incl, 31);
    }
    return res;
}

// Cells (j, k) with n_jk > kTablesThreadMax: one warp per cell, the cells taken from a list (see hdp_tables_kernel).
__global__ void __launch_bounds__(128)
hdp_tables_big_kernel(const int32_t *rows, int Kpad, int KK, const double *beta, double aa, const double *stirling, int stirling_rows,
                      uint64_t seed, uint32_t sweep, uint32_t internal, unsigned long long *tables, int *err,
                      const long long *cells, const unsigned int *n_cells) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const Philox philox(seed);
    const int64_t n_list = *n_cells;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_list; q += nw) {
        const int64_t c = cells[q], j = c / KK;
        const int k = (int)(c - j * KK);
        const int n = rows[(size_t)j * Kpad + k];
        if (n > stirling_rows) {            // the reference indexes past its 10k table: BoundsError
            if (lane == 0) *err = 1;
            continue;
        }
        uint32_t o[4];
        philox((uint32_t)j, (uint32_t)((uint64_t)j >> 32) ^ ((uint32_t)k << 8), sweep * 64u + internal, kStreamTable, o);
        const double u = u01_double(o[0], o[1]);
        const double *row = stirling + (size_t)n * (n - 1) / 2;
        const int mjk = table_draw(row, n, log(aa * beta[k]), u, lane, false, nullptr);
        if (lane == 0) {
            atomicAdd(tables + k, (unsigned long long)mjk);
            atomicAdd(tables + Kpad, (unsigned long long)mjk);
        }
    }
}

// All cells (j, k) with 0 < n_jk <= kTablesThreadMax: ONE THREAD per cell.  A group uses a handful of the KK components
// and its counts are tens, so a warp per cell left three quarters of the issue slots to idle lanes and to the shuffles
// of 32-wide sums over ~25 terms (0.23 ms per pass over 100,000 groups, ten passes per sweep).  Here a warp scans 32
// cells at a time, queues the non-empty ones in shared memory and hands each lane a cell of its own: the lane walks its
// Stirling row (the rows up to n = 128 are 66 KB: L1-resident) three times — maximum, sum of exp, CDF walk in m order —
// which is the reference's lognormalize! + sample (src/common.jl:42-52, 69-79) with its left-to-right sums.  Same uniforms
// as the warp-per-cell kernel (Philox keyed by group, component, sweep, pass).  The per-component totals are collected
// in shared memory and flushed once per CTA (every cell used to do two atomics on the same ~20 addresses).
constexpr int kTablesThreadMax = 128;
constexpr int kTablesWarps = 4;
__global__ void __launch_bounds__(kTablesWarps * 32)
hdp_tables_kernel(const int32_t *rows, int64_t G, int Kpad, int KK, const double *beta, double aa,
                  const double *stirling, int stirling_rows, uint64_t seed, uint32_t sweep, uint32_t internal,
                  unsigned long long *tables, int *err, long long *big_cells, unsigned int *n_big) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *la_s = reinterpret_cast<double *>(smem_raw);                                   // [KK] log(aa beta_k)
    unsigned long long *tot_s = reinterpret_cast<unsigned long long *>(la_s + KK);         // [KK]
    long long *queue = reinterpret_cast<long long *>(tot_s + KK) + (threadIdx.x >> 5) * 64; // [64] cells of this warp
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < KK; k += blockDim.x) {
        la_s[k] = log(aa * beta[k]);
        tot_s[k] = 0ull;
    }
    __syncthreads();
    const Philox philox(seed);
    const int64_t n_cells = G * (int64_t)KK, n_chunks = (n_cells + 31) / 32;
    const int64_t nw = (int64_t)gridDim.x * kTablesWarps;
    int queued = 0;
    for (int64_t ch = (int64_t)blockIdx.x * kTablesWarps + (threadIdx.x >> 5);; ch += nw) {
        const bool more = ch < n_chunks;                       // warp-uniform
        if (more) {
            const int64_t c = ch * 32 + lane;
            int n = 0;
            if (c < n_cells) {
                const int64_t j = c / KK;
                n = rows[(size_t)j * Kpad + (c - j * KK)];
            }
            if (n > kTablesThreadMax) big_cells[atomicAdd(n_big, 1u)] = c;
            const bool mine = n > 0 && n <= kTablesThreadMax;
            const unsigned bal = __ballot_sync(kFull, mine);
            if (mine) queue[queued + __popc(bal & ((1u << lane) - 1u))] = c;
            queued += __popc(bal);
            __syncwarp();
        }
        // a full warp of cells — or whatever is left at the end
        if (queued >= 32 || (!more && queued > 0)) {
            const int take = min(queued, 32);
            if (lane < take) {
                const int64_t c = queue[queued - take + lane], j = c / KK;
                const int k = (int)(c - j * KK);
                const int n = rows[(size_t)j * Kpad + k];
                uint32_t o[4];
                philox((uint32_t)j, (uint32_t)((uint64_t)j >> 32) ^ ((uint32_t)k << 8), sweep * 64u + internal, kStreamTable, o);
                const double u = u01_double(o[0], o[1]);
                const double *row = stirling + (size_t)n * (n - 1) / 2;
                const double la = la_s[k];
                double lmax = -INFINITY;
                for (int m = 1; m <= n; ++m) lmax = fmax(lmax, row[m - 1] + (double)m * la);
                double total = 0.0;
                for (int m = 1; m <= n; ++m) total += exp(row[m - 1] + (double)m * la - lmax);
                const double target = u * total;
                double cw = 0.0;
                int mjk = n;
                for (int m = 1; m <= n; ++m) {
                    cw += exp(row[m - 1] + (double)m * la - lmax);
                    if (cw >= target) { mjk = m; break; }
                }
                atomicAdd(tot_s + k, (unsigned long long)mjk);
            }
            queued -= take;
            __syncwarp();
        }
        if (!more) break;
    }
    __syncthreads();
    unsigned long long all = 0;
    for (int k = threadIdx.x; k < KK; k += blockDim.x) {
        const unsigned long long v = tot_s[k];
        if (v) atomicAdd(tables + k, v);
        all += v;
    }
    for (int o = 16; o > 0; o >>= 1) all += __shfl_xor_sync(kFull, all, o);
    if (lane == 0 && all) atomicAdd(tables + Kpad, all);
}

__global__ void hdp_table_conditionals_kernel(const int32_t *n_arr, const double *a_arr, const double *u_arr,
                                              int64_t nq, int n_max, const double *stirling, double *prob,
                                              int32_t *draws) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += nw) {
        const int n = n_arr[q];
        const double *row = stirling + (size_t)n * (n - 1) / 2;
        const int m = table_draw(row, n, log(a_arr[q]), u_arr[q], lane, true, prob + (size_t)q * n_max);
        if (lane == 0) draws[q] = m;
    }
}

// alpha0 auxiliaries (src/HDP.jl:133-147): w_j ~ Beta(aa+1, n_j), s_j ~ Bernoulli(n_j / (n_j + aa)).
// out[0] += sum_j log w_j, out[1] += sum_j s_j
__global__ void hdp_hyper_kernel(const int64_t *group_ptr, int64_t G, double aa, uint64_t seed, uint32_t sweep,
                                 uint32_t internal, double *out, const double *ext_uniforms = nullptr) {
    double lw = 0.0, ss = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < G; j += (int64_t)gridDim.x * blockDim.x) {
        const double nj = (double)(group_ptr[j + 1] - group_ptr[j]);
        PhiloxStream rs(seed, (uint32_t)j, (uint32_t)((uint64_t)j >> 32), (sweep * 64u + internal) * 8u + kStreamHyper);
        if (nj > 0.0) {
            const double ga = rs.gamma(aa + 1.0), gb = rs.gamma(nj);
            lw += log(ga / (ga + gb));
            const double u = rs.u01();
            ss += ((ext_uniforms ? ext_uniforms[j] : u) < nj / (nj + aa)) ? 1.0 : 0.0;
        }
    }
    lw = warp_sum(lw);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, lw);
        atomicAdd(out + 1, ss);
    }
}

// ---------------------------------------------------------------------------------------
// serial birth pass (src/HDP.jl:312-338 with the reference's sequential semantics): one warp
// ---------------------------------------------------------------------------------------
// the serial walk over list[0 .. n_list) by ONE warp; sc = (Kmax + 1) doubles of the warp.  Returns the new KK.
template <int KIND, int DATUM>
__device__ __forceinline__ int birth_serial(const GenericParams &p, const int32_t *list, int n_list, int KK, double *beta, double gg,
                                            int Kmax, int *err, double *sc, int lane) {
    double diag_acc = 0.0;
    unsigned long long moved = 0;
    for (int q = 0; q < n_list; ++q) {
        const int64_t i = __ldcg(list + q);
        int32_t *prow = p.prior_rows + (size_t)p.row_of[i] * p.Kpad;
        const Datum d = load_datum<KIND, DATUM>(p, i, lane);
        const int epoch = (p.model == BIAS_MODEL_RCRP) ? p.row_of[i] : 0;
        const double alpha_i = (p.model == BIAS_MODEL_RCRP) ? __ldcg(p.rcrp_aa + epoch) : p.alpha;
        const double lmax = score_all<KIND, DATUM>(p, d, -1, KK, KK + 1, prow, beta, alpha_i, sc, lane, epoch);
        PhiloxStream rs(p.seed, (uint32_t)i, (uint32_t)((uint64_t)i >> 32), p.sweep * 8u + kStreamBirth);
        const double u = rs.u01();
        int knew = draw_parallel(sc, KK + 1, lmax, u, lane);
        if (knew == KK) {
            if (KK + 2 > Kmax) {                 // slot capacity exhausted
                if (lane == 0) *err = 2;
                knew = KK - 1;
            } else {
                const double b = 1.0 - pow(1.0 - rs.u01(), 1.0 / gg);      // Beta(1, gg), src/HDP.jl:325
                if (lane == 0 && p.model == BIAS_MODEL_HDP) {
                    const double b_new = beta[KK];
                    beta[KK] = b * b_new;                                   // :326-328
                    beta[KK + 1] = (1.0 - b) * b_new;
                }
                KK += 1;
                __syncwarp();
            }
        }
        apply_move<KIND, DATUM>(p, d, -1, knew, prow, lane);
        if (lane == 0) p.z[i] = knew;
        __threadfence();
        __syncwarp();
        const double dv = diag_after<KIND, DATUM>(p, d, knew, lane);
        if (lane == 0) diag_acc += dv;
        moved += 1;
    }
    if (lane == 0) {
        atomicAdd(p.diag_sum, diag_acc);
        atomicAdd(p.moved, moved);
    }
    return KK;
}

template <int KIND, int DATUM>
__global__ void __launch_bounds__(32)
hdp_birth_kernel(GenericParams p, const int32_t *list, int n_list, int32_t *d_KK, double *beta, double gg,
                 int Kmax, int *err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sc = reinterpret_cast<double *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int KK = birth_serial<KIND, DATUM>(p, list, n_list, *d_KK, beta, gg, Kmax, err, sc, lane);
    if (lane == 0) *d_KK = KK;
}

// ---------------------------------------------------------------------------------------
// prune: column compaction of an int32 table [rows][Kpad] with remap[k] = new index or -1
// ---------------------------------------------------------------------------------------
__global__ void compact_columns_kernel(int32_t *tab, int64_t rows, int Kpad, int KK_old, const int32_t *remap) {
    extern __shared__ int32_t buf[];
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        int32_t *row = tab + (size_t)r * Kpad;
        for (int k = threadIdx.x; k < KK_old; k += blockDim.x) buf[k] = row[k];
        __syncthreads();
        for (int k = threadIdx.x; k < KK_old; k += blockDim.x) row[k] = 0;
        __syncthreads();
        for (int k = threadIdx.x; k < KK_old; k += blockDim.x) {
            const int nk = remap[k];
            if (nk >= 0) row[nk] = buf[k];
        }
        __syncthreads();
    }
}

__global__ void remap_z_kernel(int32_t *z, int64_t N, const int32_t *remap) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = z[i];
        if (k >= 0) z[i] = remap[k];
    }
}

static int grid1d(bias_ctx *ctx, int64_t n, int threads, int per_sm = 8) {
    int64_t blocks = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)ctx->num_sms * per_sm));
}

static int hdp_prune(bias_model *m) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int KK = m->hdp.KK, Kpad = m->Kpad;
    std::vector<int32_t> ni(Kpad), nk(Kpad);
    std::vector<double> sx;
    BIAS_CUDA_TRY(cudaMemcpyAsync(ni.data(), m->n_items, (size_t)Kpad * 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(nk.data(), m->n_k, (size_t)Kpad * 4, cudaMemcpyDeviceToHost, st));
    if (m->sumx) {
        sx.resize(Kpad);
        BIAS_CUDA_TRY(cudaMemcpyAsync(sx.data(), m->sumx, (size_t)Kpad * 8, cudaMemcpyDeviceToHost, st));
    }
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int32_t> remap(Kpad, -1);
    int alive = 0;
    for (int k = 0; k < KK; ++k)
        if (ni[k] > 0) remap[k] = alive++;
    if (alive == KK) return BIAS_OK;
    if (alive == 0) {            // keep one (empty) component: the reference cannot reach KK = 0 with data present
        remap[0] = 0;
        alive = 1;
    }
    // sticks of the removed components return to the unrepresented mass (src/HDP.jl:303-305)
    std::vector<double> nb(m->h_beta.size(), 0.0);
    double rest = m->h_beta[KK];
    std::vector<int32_t> n_ni(Kpad, 0), n_nk(Kpad, 0);
    std::vector<double> n_sx(Kpad, 0.0);
    for (int k = 0; k < KK; ++k) {
        if (remap[k] >= 0) {
            nb[remap[k]] = m->h_beta[k];
            n_ni[remap[k]] = ni[k];
            n_nk[remap[k]] = nk[k];
            if (m->sumx) n_sx[remap[k]] = sx[k];
        } else {
            rest += m->h_beta[k];
        }
    }
    nb[alive] = rest;
    m->h_beta = nb;
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_remap, remap.data(), (size_t)Kpad * 4, cudaMemcpyHostToDevice, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->n_items, n_ni.data(), (size_t)Kpad * 4, cudaMemcpyHostToDevice, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->n_k, n_nk.data(), (size_t)Kpad * 4, cudaMemcpyHostToDevice, st));
    if (m->sumx) BIAS_CUDA_TRY(cudaMemcpyAsync(m->sumx, n_sx.data(), (size_t)Kpad * 8, cudaMemcpyHostToDevice, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta.data(), m->h_beta.size() * 8, cudaMemcpyHostToDevice, st));
    if (m->N > 0) {
        remap_z_kernel<<<grid1d(ctx, m->N, 256), 256, 0, st>>>(m->z, m->N, m->d_remap);
        ctx->launches += 1;
    }
    if (m->G > 0) {
        compact_columns_kernel<<<grid1d(ctx, m->G * 128, 128), 128, (size_t)KK * 4, st>>>(m->group_rows, m->G, Kpad, KK, m->d_remap);
        ctx->launches += 1;
    }
    if (m->conj.kind == BIAS_MD) {
        compact_columns_kernel<<<grid1d(ctx, m->V * 128, 128), 128, (size_t)KK * 4, st>>>(m->n_wk, m->V, Kpad, KK, m->d_remap);
        ctx->launches += 1;
    }
    BIAS_CUDA_TRY(cudaGetLastError());
    // the host vectors above must outlive the async copies
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    m->hdp.KK = alive;
    return BIAS_OK;
}

int hdp_create(bias_model *m, const bias_hdp_params *hp) {
    m->hdp = *hp;
    const int Kpad = m->Kpad;
    m->h_beta.assign((size_t)Kpad + 1, 0.0);
    for (int k = 0; k <= hp->KK; ++k) m->h_beta[k] = 1.0 / (double)(hp->KK + 1);       // src/HDP.jl:228
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_beta, ((size_t)Kpad + 1) * 8));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta.data(), ((size_t)Kpad + 1) * 8, cudaMemcpyHostToDevice, m->ctx->stream));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->pending, (size_t)std::max<int64_t>(1, m->N) * 2 * 4));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_KK, 4 * sizeof(int32_t)));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->d_KK, 0, 4 * sizeof(int32_t), m->ctx->stream));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_tables, ((size_t)Kpad + 1) * 8));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_hyper, 2 * 8));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_remap, (size_t)Kpad * 4));
    m->host_rng.seed(m->ctx->seed ^ 0x9E3779B97F4A7C15ull);
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    return BIAS_OK;
}

int hdp_destroy(bias_model *m) {
    for (void *p : {(void *)m->d_beta, (void *)m->pending, (void *)m->d_KK, (void *)m->d_tables, (void *)m->d_hyper,
                    (void *)m->d_remap})
        if (p) cudaFree(p);
    m->d_beta = nullptr;
    m->pending = nullptr;
    m->d_KK = nullptr;
    m->d_tables = nullptr;
    m->d_hyper = nullptr;
    m->d_remap = nullptr;
    return BIAS_OK;
}

static GenericParams birth_params(bias_model *m) {
    GenericParams p{};
    p.model = m->model;
    p.rcrp_T = (int32_t)m->G;
    p.rcrp_aa = m->d_aa_t;
    p.kind = m->conj.kind;
    p.datum = m->conj.datum;
    p.K = m->hdp.KK;
    p.Kpad = m->Kpad;
    p.n_out = p.K + 1;
    p.row_of = m->group_of;
    p.prior_rows = m->group_rows;
    p.word = m->word;
    p.x = m->x;
    p.sent_ptr = m->sent_ptr;
    p.sent_word = m->sent_word;
    p.sent_count = m->sent_count;
    p.z = m->z;
    p.n_wk = m->n_wk;
    p.n_k = m->n_k;
    p.n_items = m->n_items;
    p.sumx = m->sumx;
    p.V = m->V;
    p.beta = m->conj.aa;
    p.m0 = m->conj.m0;
    p.v0 = m->conj.v0;
    p.vv = m->conj.vv;
    p.alpha = m->hdp.aa;
    p.seed = m->ctx->seed;
    p.sweep = m->sweep_index;
    p.diag_sum = reinterpret_cast<double *>(m->scratch);
    p.moved = m->scratch + 1;
    return p;
}

// Serial birth pass + parallel re-sampling of the rest of the queue, until the queue is empty.
static int resolve_births(bias_model *m, int32_t *&list_cur, int32_t *&list_next) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int Kpad = m->Kpad, Kmax = m->hdp.Kmax;
    unsigned int *d_pending_count = reinterpret_cast<unsigned int *>(m->scratch + 3);
    int *d_err = reinterpret_cast<int *>(m->scratch + 4);
    uint32_t round = 0;
    for (;;) {
        unsigned int n_pending = 0;
        int err = 0;
        BIAS_CUDA_TRY(cudaMemcpyAsync(&n_pending, d_pending_count, 4, cudaMemcpyDeviceToHost, st));
        BIAS_CUDA_TRY(cudaStreamSynchronize(st));
        if (n_pending == 0) break;
        const int n_serial = (int)std::min<unsigned int>(n_pending, 256u);
        GenericParams bp = birth_params(m);
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_KK, &m->hdp.KK, 4, cudaMemcpyHostToDevice, st));
        const size_t bsm = ((size_t)Kmax + 4) * 8;
        if (bp.kind == BIAS_G1G1)
            hdp_birth_kernel<BIAS_G1G1, BIAS_F64><<<1, 32, bsm, st>>>(bp, list_cur, n_serial, m->d_KK, m->d_beta, m->hdp.gg, Kmax, d_err);
        else if (bp.datum == BIAS_INT)
            hdp_birth_kernel<BIAS_MD, BIAS_INT><<<1, 32, bsm, st>>>(bp, list_cur, n_serial, m->d_KK, m->d_beta, m->hdp.gg, Kmax, d_err);
        else
            hdp_birth_kernel<BIAS_MD, BIAS_SENT><<<1, 32, bsm, st>>>(bp, list_cur, n_serial, m->d_KK, m->d_beta, m->hdp.gg, Kmax, d_err);
        BIAS_CUDA_TRY(cudaGetLastError());
        ctx->launches += 1;
        int32_t kk_new = 0;
        BIAS_CUDA_TRY(cudaMemcpyAsync(&kk_new, m->d_KK, 4, cudaMemcpyDeviceToHost, st));
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->h_beta.data(), m->d_beta, ((size_t)Kpad + 1) * 8, cudaMemcpyDeviceToHost, st));
        BIAS_CUDA_TRY(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
        BIAS_CUDA_TRY(cudaStreamSynchronize(st));
        m->hdp.KK = kk_new;
        if (err == 2)
            return fail(BIAS_ERANGE, "HDP: Kmax = %d component slots exhausted (KK = %d); construct the model with a larger Kmax",
                        Kmax, kk_new);
        const int64_t rest = (int64_t)n_pending - n_serial;
        BIAS_CUDA_TRY(cudaMemsetAsync(d_pending_count, 0, 4, st));
        if (rest > 0) {
            // re-sample the remainder of the queue in parallel against the enlarged component set;
            // items that draw the new slot again land in the other half of the queue buffer
            m->pending = list_next;
            round += 1;
            const uint32_t saved = m->sweep_index;
            BIAS_TRY(launch_generic(m, rest, nullptr, list_cur + n_serial, m->group_of, m->group_rows, nullptr, 0, 0,
                                    nullptr, nullptr, nullptr, kStreamBirth + 8 * round));
            m->sweep_index = saved;
            std::swap(list_cur, list_next);
        }
    }
    return BIAS_OK;
}

// ---- the phases of one HDP sweep (hdp_sweep runs them back to back; the sharded driver puts exchanges between) ----

// parallel assignment pass (src/HDP.jl:279-341); items that draw the new-component slot are queued.  Does not prune.
static int hdp_phase_pass(bias_model *m) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch, 0, 6 * sizeof(unsigned long long), st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta.data(), ((size_t)m->Kpad + 1) * 8, cudaMemcpyHostToDevice, st));
    m->pending_base = m->pending;
    BIAS_TRY(launch_generic(m, m->N, nullptr, nullptr, m->group_of, m->group_rows, nullptr, 0, 0, nullptr, nullptr,
                            nullptr, kStreamItemDraw));
    return BIAS_OK;
}

// births: the serial pass over the queue + parallel re-sampling of the rest (src/HDP.jl:320-331), until it is empty
static int hdp_phase_births(bias_model *m) {
    cudaStream_t st = m->ctx->stream;
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta.data(), ((size_t)m->Kpad + 1) * 8, cudaMemcpyHostToDevice, st));
    int32_t *list_cur = m->pending_base, *list_next = m->pending_base + std::max<int64_t>(1, m->N);
    m->pending = list_cur;
    BIAS_TRY(resolve_births(m, list_cur, list_next));
    m->pending = m->pending_base;
    return BIAS_OK;
}

// table counts M_jk from the Stirling table and the alpha0 auxiliaries of this rank's groups (src/HDP.jl:343-361, :133-147)
static int hdp_phase_tables(bias_model *m, int hh) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int Kpad = m->Kpad, KK = m->hdp.KK;
    int *d_err = reinterpret_cast<int *>(m->scratch + 4);
    BIAS_TRY(ensure_stirling(ctx, (int32_t)std::min<int64_t>(m->max_group_len, 10000)));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->d_tables, 0, ((size_t)Kpad + 1) * 8, st));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->d_hyper, 0, 16, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta.data(), ((size_t)Kpad + 1) * 8, cudaMemcpyHostToDevice, st));
    if (m->G > 0) {
        // cells up to kTablesThreadMax items by one thread each; the others are listed (in the birth queue's buffer, idle in
        // this phase: 2 N int32 = N cells, and at most N / 129 cells can be that large) for the warp-per-cell kernel
        unsigned int *n_big = reinterpret_cast<unsigned int *>(m->scratch + 5);
        long long *big = reinterpret_cast<long long *>(m->pending_base ? m->pending_base : m->pending);
        BIAS_CUDA_TRY(cudaMemsetAsync(n_big, 0, 4, st));
        const size_t tsm = (size_t)KK * 16 + (size_t)kTablesWarps * 64 * 8;
        static size_t tsm_cache[kMaxDevices] = {};
        if (tsm > 48 * 1024 && tsm > tsm_cache[ctx->device % kMaxDevices]) {
            BIAS_CUDA_TRY(cudaFuncSetAttribute(hdp_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
            tsm_cache[ctx->device % kMaxDevices] = tsm;
        }
        const int64_t chunks = (m->G * (int64_t)KK + 31) / 32;
        hdp_tables_kernel<<<grid1d(ctx, chunks * 32, kTablesWarps * 32, 8), kTablesWarps * 32, tsm, st>>>(
            m->group_rows, m->G, Kpad, KK, m->d_beta, m->hdp.aa, ctx->stirling, ctx->stirling_rows, ctx->seed ^ m->seed_mix,
            m->sweep_index, (uint32_t)hh, reinterpret_cast<unsigned long long *>(m->d_tables), d_err, big, n_big);
        ctx->launches += 1;
        if (m->max_group_len > kTablesThreadMax) {
            hdp_tables_big_kernel<<<ctx->num_sms * 4, 128, 0, st>>>(
                m->group_rows, Kpad, KK, m->d_beta, m->hdp.aa, ctx->stirling, ctx->stirling_rows, ctx->seed ^ m->seed_mix,
                m->sweep_index, (uint32_t)hh, reinterpret_cast<unsigned long long *>(m->d_tables), d_err, big, n_big);
            ctx->launches += 1;
        }
        if (m->hdp.sample_hyperparam) {
            hdp_hyper_kernel<<<grid1d(ctx, m->G, 128, 4), 128, 0, st>>>(m->group_ptr, m->G, m->hdp.aa, ctx->seed,
                                                                       m->sweep_index, (uint32_t)hh, m->d_hyper);
            ctx->launches += 1;
        }
        BIAS_CUDA_TRY(cudaGetLastError());
    }
    return BIAS_OK;
}

// alpha0 (Teh et al. eq. 50 with the device-side auxiliaries sum log w_j, sum s_j) and gamma (Escobar-West), src/HDP.jl:124-159,
// as a function of the host RNG and the priors (the model's, or one epoch's of the RCRF: src/RCRF.jl:80-118 with KK = dishes
// of the epoch).  A non-positive Gamma shape or rate is an error, as Distributions.Gamma's argument check is in the reference.
static int host_hyper_draw(std::mt19937_64 &rng, double a1, double a2, double g1, double g2, double mt, double sum_logw,
                           double sum_s, int KK, double &aa_io, double &gg_io) {
    const double aa_shape = a1 + mt - sum_s;            // :146
    const double aa_rate = a2 - sum_logw;               // :147
    if (!(aa_shape > 0.0) || !(aa_rate > 0.0))
        return fail(BIAS_EINVAL, "HDP: Gamma(%g) / %g for alpha0 is not a distribution (a1 = %g, a2 = %g, tables = %g, sum s = %g)",
                    aa_shape, aa_rate, a1, a2, mt, sum_s);
    std::gamma_distribution<double> ga(aa_shape, 1.0);
    aa_io = ga(rng) / aa_rate;                          // :148
    std::gamma_distribution<double> x1d(gg_io + 1.0, 1.0), x2d(std::max(mt, 1e-300), 1.0);
    const double x1 = x1d(rng), x2 = x2d(rng);
    const double eta = x1 / (x1 + x2);                  // Beta(gg + 1, m), :151
    const double le = std::log(eta);
    const double pi_eta = 1.0 / (1.0 + (mt * (g2 - le)) / (g1 + KK - 1.0));   // :152
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    const double shape = (ud(rng) < pi_eta) ? g1 + KK : g1 + KK - 1.0;         // :154-158
    if (!(shape > 0.0) || !(g2 - le > 0.0))
        return fail(BIAS_EINVAL, "HDP: Gamma(%g) / %g for gamma is not a distribution (g1 = %g, g2 = %g, KK = %d)", shape, g2 - le, g1, g2, KK);
    std::gamma_distribution<double> gd(shape, 1.0);
    gg_io = gd(rng) / (g2 - le);
    return BIAS_OK;
}
static int hdp_host_hyper_on(bias_model *m, double mt, double sum_logw, double sum_s, int KK, double &aa_io, double &gg_io) {
    return host_hyper_draw(m->host_rng, m->hdp.a1, m->hdp.a2, m->hdp.g1, m->hdp.g2, mt, sum_logw, sum_s, KK, aa_io, gg_io);
}
static int hdp_host_hyper(bias_model *m, double mt, double sum_logw, double sum_s) {
    return hdp_host_hyper_on(m, mt, sum_logw, sum_s, m->hdp.KK, m->hdp.aa, m->hdp.gg);
}

// beta ~ Dirichlet([sum_j M_j1, ..., sum_j M_jKK, gg]) (src/HDP.jl:362-364) as normalised Gamma draws
static void host_beta_draw(std::mt19937_64 &rng, const unsigned long long *tables, int KK, double gg, double *beta) {
    double tot = 0.0;
    for (int k = 0; k <= KK; ++k) {
        const double a = (k < KK) ? (double)tables[k] : gg;
        double g = 0.0;
        if (a > 0.0) {
            std::gamma_distribution<double> gd(a, 1.0);
            g = gd(rng);
        }
        beta[k] = g;
        tot += g;
    }
    for (int k = 0; k <= KK; ++k) beta[k] = tot > 0.0 ? beta[k] / tot : 1.0 / (KK + 1);
}

// beta ~ Dirichlet([sum_j M_jk ..., gg]) (src/HDP.jl:362-364), alpha0 and gamma (:124-159) from the (possibly
// all-reduced) contents of d_tables / d_hyper; n_groups_total = number of groups over all ranks
static int hdp_phase_draw(bias_model *m) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int Kpad = m->Kpad, KK = m->hdp.KK;
    int *d_err = reinterpret_cast<int *>(m->scratch + 4);
    std::vector<unsigned long long> tables((size_t)Kpad + 1);
    double hyper[2] = {0.0, 0.0};
    int err = 0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(tables.data(), m->d_tables, ((size_t)Kpad + 1) * 8, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(hyper, m->d_hyper, 16, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    if (err == 1)
        return fail(BIAS_ERANGE, "HDP: some n_jk exceeds the %d-row Stirling table (the reference throws BoundsError, src/HDP.jl:355)",
                    ctx->stirling_rows);
    host_beta_draw(m->host_rng, tables.data(), KK, m->hdp.gg, m->h_beta.data());
    for (size_t k = (size_t)KK + 1; k < m->h_beta.size(); ++k) m->h_beta[k] = 0.0;
    if (m->hdp.sample_hyperparam) BIAS_TRY(hdp_host_hyper(m, (double)tables[Kpad], hyper[0], hyper[1]));
    return BIAS_OK;
}

static int hdp_phase_end(bias_model *m) {
    cudaStream_t st = m->ctx->stream;
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta.data(), ((size_t)m->Kpad + 1) * 8, cudaMemcpyHostToDevice, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    m->sweep_index += 1;
    m->stats_pending = true;
    return BIAS_OK;
}

int hdp_sweep(bias_model *m) {
    BIAS_TRY(hdp_prune(m));                       // src/HDP.jl:248-264
    BIAS_TRY(hdp_phase_pass(m));
    BIAS_TRY(hdp_phase_births(m));
    BIAS_TRY(hdp_prune(m));                       // components that died during the pass (src/HDP.jl:291-307)
    for (int hh = 0; hh < m->hdp.n_internals; ++hh) {
        BIAS_TRY(hdp_phase_tables(m, hh));
        BIAS_TRY(hdp_phase_draw(m));
    }
    return hdp_phase_end(m);
}

// ---------------------------------------------------------------------------------------
// CRF (reference src/HDP.jl:402-709): tables and dishes
// ---------------------------------------------------------------------------------------
#include "crf.cuh"

static CrfParams crf_params(bias_model *m, const double *ext_uniforms) {
    CrfParams p{};
    p.group_ptr = m->group_ptr;
    p.G = m->G;
    p.x = m->x;
    p.z = m->z;
    p.tji = m->crf_tji;
    p.njt = m->crf_njt;
    p.kjt = m->crf_kjt;
    p.n_tbls = m->crf_n_tbls;
    p.mm = m->crf_mm;
    p.n_items = m->n_items;
    p.sumx = m->sumx;
    p.word = m->word;
    p.sent_ptr = m->sent_ptr;
    p.sent_word = m->sent_word;
    p.sent_count = m->sent_count;
    p.n_wk = m->n_wk;
    p.n_k = m->n_k;
    p.beta = m->conj.aa;
    p.V = m->V;
    p.d_KK = m->d_KK;
    p.Kmax = m->hdp.Kmax;
    p.m0 = m->conj.m0;
    p.v0 = m->conj.v0;
    p.vv = m->conj.vv;
    p.aa = m->hdp.aa;
    p.gg = m->hdp.gg;
    p.seed = m->ctx->seed;
    p.sweep = m->sweep_index;
    p.ext_uniforms = ext_uniforms;
    p.lock = reinterpret_cast<int *>(m->scratch + 7);
    p.err = reinterpret_cast<int *>(m->scratch + 4);
    p.moved = m->scratch + 1;
    p.max_len = (int32_t)m->max_group_len;
    p.rcrf = 0;
    p.T = 1;
    p.t = 0;
    p.Kpad = m->Kpad;
    p.g0 = 0;
    p.g1 = m->G;
    p.tmp = m->crf_tmp;
    return p;
}

// the launch parameters of epoch t of an RCRF model
static CrfParams rcrf_params(bias_model *m, int t, const double *ext_uniforms) {
    CrfParams p = crf_params(m, ext_uniforms);
    p.rcrf = 1;
    p.T = (int32_t)m->h_epoch_gptr.size() - 1;
    p.t = t;
    p.g0 = m->h_epoch_gptr[t];
    p.g1 = m->h_epoch_gptr[t + 1];
    p.aa = m->h_aa_t[t];
    p.gg = m->h_gg_t[t];
    return p;
}

// the two passes (and the diagnostic) for the model's item kind
static int crf_launch_passes(bias_model *m, const CrfParams &p, int grid, size_t smem, bool join) {
    cudaStream_t st = m->ctx->stream;
    const int64_t Gt = p.g1 - p.g0;
#define CRF_PASSES(D)                                                                          \
    do {                                                                                       \
        crf_tables_kernel<D><<<grid, kCrfWarps * 32, smem, st>>>(p);                            \
        if (join) crf_join_kernel<<<(unsigned)((Gt + 127) / 128), 128, 0, st>>>(p);             \
        crf_dishes_kernel<D><<<grid, kCrfWarps * 32, smem, st>>>(p);                            \
    } while (0)
    if (m->conj.kind == BIAS_G1G1) CRF_PASSES(BIAS_F64);
    else if (m->conj.datum == BIAS_INT) CRF_PASSES(BIAS_INT);
    else CRF_PASSES(BIAS_SENT);
#undef CRF_PASSES
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += join ? 3 : 2;
    return BIAS_OK;
}

static int crf_set_smem(bias_model *m, size_t smem) {
    static size_t smem_cache[kMaxDevices] = {};          // function attributes are per device
    size_t &smem_set = smem_cache[m->ctx->device % kMaxDevices];
    if (smem > 48 * 1024 && smem > smem_set) {
        BIAS_CUDA_TRY(cudaFuncSetAttribute(crf_tables_kernel<BIAS_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BIAS_CUDA_TRY(cudaFuncSetAttribute(crf_dishes_kernel<BIAS_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BIAS_CUDA_TRY(cudaFuncSetAttribute(crf_tables_kernel<BIAS_INT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BIAS_CUDA_TRY(cudaFuncSetAttribute(crf_dishes_kernel<BIAS_INT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BIAS_CUDA_TRY(cudaFuncSetAttribute(crf_tables_kernel<BIAS_SENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BIAS_CUDA_TRY(cudaFuncSetAttribute(crf_dishes_kernel<BIAS_SENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    return BIAS_OK;
}

static int crf_launch_diag(bias_model *m) {
    bias_ctx *ctx = m->ctx;
    if (m->N <= 0) return BIAS_OK;
    const CrfParams p = crf_params(m, nullptr);
    double *out = reinterpret_cast<double *>(m->scratch);
    const int grid = grid1d(ctx, m->N, 256, 4);
    if (m->conj.kind == BIAS_G1G1) crf_diag_kernel<BIAS_F64><<<grid, 256, 0, ctx->stream>>>(p, m->N, out);
    else if (m->conj.datum == BIAS_INT) crf_diag_kernel<BIAS_INT><<<grid, 256, 0, ctx->stream>>>(p, m->N, out);
    else crf_diag_kernel<BIAS_SENT><<<grid, 256, 0, ctx->stream>>>(p, m->N, out);
    BIAS_CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    return BIAS_OK;
}

int crf_create(bias_model *m) {
    cudaStream_t st = m->ctx->stream;
    if (crf_smem_bytes((int)m->max_group_len, m->hdp.Kmax) > 200 * 1024)
        return fail(BIAS_ERANGE, "CRF: groups of %lld items with Kmax = %d do not fit the per-warp staging buffers",
                    (long long)m->max_group_len, m->hdp.Kmax);
    const size_t n = (size_t)std::max<int64_t>(1, m->N), g = (size_t)std::max<int64_t>(1, m->G);
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->crf_tji, n * 4));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->crf_njt, n * 4));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->crf_kjt, n * 4));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->crf_n_tbls, g * 4));
    const size_t T = (m->model == BIAS_MODEL_RCRF) ? m->h_epoch_gptr.size() - 1 : 1;
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->crf_mm, T * (size_t)m->Kpad * 4));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->crf_tmp, n * 4));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->crf_tji, 0, n * 4, st));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->crf_njt, 0, n * 4, st));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->crf_kjt, 0, n * 4, st));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->crf_n_tbls, 0, g * 4, st));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->crf_mm, 0, T * (size_t)m->Kpad * 4, st));
    for (size_t t = 0; t < T; ++t) {
        const CrfParams p = (m->model == BIAS_MODEL_RCRF) ? rcrf_params(m, (int)t, nullptr) : crf_params(m, nullptr);
        if (p.g1 > p.g0) {
            crf_init_kernel<<<(unsigned)((p.g1 - p.g0 + 127) / 128), 128, 0, st>>>(p);
            BIAS_CUDA_TRY(cudaGetLastError());
            m->ctx->launches += 1;
        }
    }
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    return BIAS_OK;
}

int crf_destroy(bias_model *m) {
    for (void *p : {(void *)m->crf_tji, (void *)m->crf_njt, (void *)m->crf_kjt, (void *)m->crf_n_tbls, (void *)m->crf_mm, (void *)m->crf_tmp})
        if (p) cudaFree(p);
    m->crf_tji = m->crf_njt = m->crf_kjt = m->crf_n_tbls = m->crf_mm = m->crf_tmp = nullptr;
    return BIAS_OK;
}

// dishes that lost their last table leave the menu (src/HDP.jl:531-545, :610-621; src/RCRF.jl:271-284), applied once per
// iteration: a dish without tables has no customers, so the test is n_items == 0
static int crf_prune(bias_model *m) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int KK = m->hdp.KK, Kpad = m->Kpad;
    std::vector<int32_t> ni(Kpad);
    BIAS_CUDA_TRY(cudaMemcpyAsync(ni.data(), m->n_items, (size_t)Kpad * 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    int alive = 0;
    for (int k = 0; k < KK; ++k) alive += ni[k] > 0;
    if (alive == KK || alive == 0) return BIAS_OK;
    BIAS_TRY(hdp_prune(m));                      // compacts n_items / sumx / z (alive order) and leaves d_remap on the device
    const int64_t T = (m->model == BIAS_MODEL_RCRF) ? (int64_t)m->h_epoch_gptr.size() - 1 : 1;
    compact_columns_kernel<<<grid1d(ctx, T * 128, 128), 128, (size_t)KK * 4, st>>>(m->crf_mm, T, Kpad, KK, m->d_remap);
    ctx->launches += 1;
    if (m->N > 0) {
        remap_z_kernel<<<grid1d(ctx, m->N, 256), 256, 0, st>>>(m->crf_kjt, m->N, m->d_remap);
        ctx->launches += 1;
    }
    BIAS_CUDA_TRY(cudaGetLastError());
    return BIAS_OK;
}

int crf_sweep(bias_model *m, const double *ext_uniforms_dev) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch, 0, 8 * sizeof(unsigned long long), st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_KK, &m->hdp.KK, 4, cudaMemcpyHostToDevice, st));
    if (m->G > 0) {
        const CrfParams p = crf_params(m, ext_uniforms_dev);
        const size_t smem = crf_smem_bytes(p.max_len, p.Kmax);
        BIAS_TRY(crf_set_smem(m, smem));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((m->G + kCrfWarps - 1) / kCrfWarps, (int64_t)ctx->num_sms * 8));
        BIAS_TRY(crf_launch_passes(m, p, grid, smem, false));
    }
    int32_t kk = 0;
    int err = 0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(&kk, m->d_KK, 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(&err, reinterpret_cast<int *>(m->scratch + 4), 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    m->hdp.KK = kk;
    if (err == 2)
        return fail(BIAS_ERANGE, "CRF: Kmax = %d dish slots exhausted (KK = %d); construct the model with a larger Kmax", m->hdp.Kmax, kk);
    BIAS_TRY(crf_prune(m));
    if (m->hdp.sample_hyperparam && m->G > 0) {     // src/HDP.jl:658-667 -> :124-164
        unsigned long long *d_m = m->scratch + 5;
        crf_sum_tables_kernel<<<grid1d(ctx, m->G, 256, 4), 256, 0, st>>>(m->crf_n_tbls, m->G, d_m);
        ctx->launches += 1;
        unsigned long long mt = 0;
        BIAS_CUDA_TRY(cudaMemcpyAsync(&mt, d_m, 8, cudaMemcpyDeviceToHost, st));
        for (int hh = 0; hh < m->hdp.n_internals; ++hh) {
            double hyper[2] = {0.0, 0.0};
            BIAS_CUDA_TRY(cudaMemsetAsync(m->d_hyper, 0, 16, st));
            hdp_hyper_kernel<<<grid1d(ctx, m->G, 128, 4), 128, 0, st>>>(m->group_ptr, m->G, m->hdp.aa, ctx->seed, m->sweep_index,
                                                                       (uint32_t)hh, m->d_hyper);
            ctx->launches += 1;
            BIAS_CUDA_TRY(cudaMemcpyAsync(hyper, m->d_hyper, 16, cudaMemcpyDeviceToHost, st));
            BIAS_CUDA_TRY(cudaStreamSynchronize(st));
            BIAS_TRY(hdp_host_hyper(m, (double)mt, hyper[0], hyper[1]));
        }
    }
    BIAS_TRY(crf_launch_diag(m));
    m->sweep_index += 1;
    m->stats_pending = true;
    return BIAS_OK;
}

// The hyper-parameters of ONE epoch of the RCRF (src/RCRF.jl:80-118 = sample_hyperparam! of src/HDP.jl:124-159 with the epoch's
// restaurants, tables and dishes): n_internals rounds of { w_j ~ Beta(aa + 1, n_j), s_j ~ Bernoulli(n_j / (n_j + aa)) over the epoch's
// restaurants; aa ~ Gamma(a1 + m - sum s) / (a2 - sum log w); gg by Escobar-West } in one block, no host round trip (the first version
// read two sums back per round: 1,100 synchronisations per sweep of 100 epochs).  err = 4: a Gamma parameter is not positive
// (Distributions.Gamma's argument check in the reference).
__global__ void __launch_bounds__(256)
rcrf_hyper_kernel(const int64_t *group_ptr, const int32_t *n_tbls, int64_t g0, int64_t g1, const int32_t *mm_row, int Kpad,
                  double *aa_t, double *gg_t, int t, double a1, double a2, double g1p, double g2p, int n_internals,
                  uint64_t seed, uint32_t sweep, int *err) {
    __shared__ double s_a[8], s_b[8], s_out[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_sum2 = [&](double a, double b, double &ra, double &rb) {
        a = warp_sum(a);
        b = warp_sum(b);
        __syncthreads();
        if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
        __syncthreads();
        if (tid == 0) {
            double x = 0.0, y = 0.0;
            for (int w = 0; w < 8; ++w) { x += s_a[w]; y += s_b[w]; }
            s_out[0] = x;
            s_out[1] = y;
        }
        __syncthreads();
        ra = s_out[0];
        rb = s_out[1];
    };
    double m_loc = 0.0, kk_loc = 0.0;
    for (int64_t j = g0 + tid; j < g1; j += 256) m_loc += (double)n_tbls[j];
    for (int k = tid; k < Kpad; k += 256) kk_loc += mm_row[k] != 0 ? 1.0 : 0.0;
    double mt, KK;
    block_sum2(m_loc, kk_loc, mt, KK);
    double aa = aa_t[t], gg = gg_t[t];
    for (int hh = 0; hh < n_internals; ++hh) {
        double lw = 0.0, ss = 0.0;
        for (int64_t j = g0 + tid; j < g1; j += 256) {
            const double nj = (double)(group_ptr[j + 1] - group_ptr[j]);
            if (nj > 0.0) {
                PhiloxStream rs(seed, (uint32_t)j, (uint32_t)((uint64_t)j >> 32), ((sweep * 1024u + (uint32_t)t) * 64u + (uint32_t)hh) * 8u + kStreamHyper);
                const double ga = rs.gamma(aa + 1.0), gb = rs.gamma(nj);
                lw += log(ga / (ga + gb));
                ss += (rs.u01() < nj / (nj + aa)) ? 1.0 : 0.0;
            }
        }
        double sum_logw, sum_s;
        block_sum2(lw, ss, sum_logw, sum_s);
        if (tid == 0) {
            PhiloxStream rs(seed ^ 0x5851F42D4C957F2Dull, (uint32_t)t, 0x52435246u, (sweep * 64u + (uint32_t)hh) * 8u + kStreamHyper);
            const double aa_shape = a1 + mt - sum_s, aa_rate = a2 - sum_logw;              // :146-147
            if (!(aa_shape > 0.0) || !(aa_rate > 0.0)) *err = 4;
            else aa = rs.gamma(aa_shape) / aa_rate;                                          // :148
            const double x1 = rs.gamma(gg + 1.0), x2 = rs.gamma(fmax(mt, 1e-300));
            const double le = log(x1 / (x1 + x2));                                          // log of Beta(gg + 1, m), :151
            const double pi_eta = 1.0 / (1.0 + (mt * (g2p - le)) / (g1p + KK - 1.0));       // :152
            const double shape = (rs.u01() < pi_eta) ? g1p + KK : g1p + KK - 1.0;           // :154-158
            if (!(shape > 0.0) || !(g2p - le > 0.0)) *err = 4;
            else gg = rs.gamma(shape) / (g2p - le);
            s_out[0] = aa;
        }
        __syncthreads();
        aa = s_out[0];
    }
    if (tid == 0) {
        aa_t[t] = aa;
        gg_t[t] = gg;
    }
}

// RCRF (reference src/RCRF.jl:122-938): per epoch the CRF's two passes over that epoch's restaurants with the dish
// prior coupled to the neighbouring epochs' menus, same-dish tables joined in between, the epoch's hyper-parameters after
int rcrf_create(bias_model *m, int64_t n_epochs, const int64_t *epoch_ptr, const double *gg, const double *aa, int join_tables) {
    m->h_epoch_gptr.assign(epoch_ptr, epoch_ptr + n_epochs + 1);
    m->h_gg_t.assign(gg, gg + n_epochs);
    m->h_aa_t.assign(aa, aa + n_epochs);
    m->crf_join = join_tables ? 1 : 0;
    if (!m->d_aa_t) BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_aa_t, (size_t)std::max<int64_t>(1, n_epochs) * 8));
    if (!m->d_gg_t) BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_gg_t, (size_t)std::max<int64_t>(1, n_epochs) * 8));
    return crf_create(m);
}

int rcrf_sweep(bias_model *m, const double *ext_uniforms_dev) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int T = (int)m->h_epoch_gptr.size() - 1, Kpad = m->Kpad;
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch, 0, 8 * sizeof(unsigned long long), st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_KK, &m->hdp.KK, 4, cudaMemcpyHostToDevice, st));
    const size_t smem = crf_smem_bytes((int)m->max_group_len, m->hdp.Kmax);
    BIAS_TRY(crf_set_smem(m, smem));
    const uint64_t hyper_seed = m->host_rng();          // the host stream (bias_model_set_host_seed) keys the device draws
    if (m->hdp.sample_hyperparam) {
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_aa_t, m->h_aa_t.data(), (size_t)T * 8, cudaMemcpyHostToDevice, st));
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_gg_t, m->h_gg_t.data(), (size_t)T * 8, cudaMemcpyHostToDevice, st));
    }
    for (int t = 0; t < T; ++t) {
        const CrfParams p = rcrf_params(m, t, ext_uniforms_dev);
        const int64_t Gt = p.g1 - p.g0;
        if (Gt <= 0) continue;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((Gt + kCrfWarps - 1) / kCrfWarps, (int64_t)ctx->num_sms * 8));
        BIAS_TRY(crf_launch_passes(m, p, grid, smem, m->crf_join != 0));
        if (m->hdp.sample_hyperparam) {             // src/RCRF.jl:464-468 -> :80-118, on the device
            rcrf_hyper_kernel<<<1, 256, 0, st>>>(m->group_ptr, m->crf_n_tbls, p.g0, p.g1, m->crf_mm + (size_t)t * Kpad, Kpad, m->d_aa_t,
                                                 m->d_gg_t, t, m->hdp.a1, m->hdp.a2, m->hdp.g1, m->hdp.g2, m->hdp.n_internals,
                                                 hyper_seed, m->sweep_index, reinterpret_cast<int *>(m->scratch + 4));
            BIAS_CUDA_TRY(cudaGetLastError());
            ctx->launches += 1;
        }
    }
    if (m->hdp.sample_hyperparam) {
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->h_aa_t.data(), m->d_aa_t, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->h_gg_t.data(), m->d_gg_t, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
    }
    int32_t kk = 0;
    int err = 0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(&kk, m->d_KK, 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(&err, reinterpret_cast<int *>(m->scratch + 4), 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    m->hdp.KK = kk;
    if (err == 2)
        return fail(BIAS_ERANGE, "RCRF: Kmax = %d dish slots exhausted (KK = %d); construct the model with a larger Kmax", m->hdp.Kmax, kk);
    if (err == 4)
        return fail(BIAS_EINVAL, "RCRF: a Gamma parameter of the hyper-parameter update is not positive (a1 = %g, a2 = %g, g1 = %g, g2 = %g)",
                    m->hdp.a1, m->hdp.a2, m->hdp.g1, m->hdp.g2);
    BIAS_TRY(crf_prune(m));
    BIAS_TRY(crf_launch_diag(m));
    m->hdp.aa = m->h_aa_t[0];
    m->hdp.gg = m->h_gg_t[0];
    m->sweep_index += 1;
    m->stats_pending = true;
    return BIAS_OK;
}

// ---------------------------------------------------------------------------------------
// RCRP (reference src/RCRP.jl:77-416): epochs are visited serially, the items of an epoch in parallel
// ---------------------------------------------------------------------------------------
int rcrp_create(bias_model *m, const bias_rcrp_params *rp, const double *aa) {
    m->h_aa_t.assign(aa, aa + m->G);
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->d_aa_t, (size_t)m->G * 8));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_aa_t, m->h_aa_t.data(), (size_t)m->G * 8, cudaMemcpyHostToDevice, m->ctx->stream));
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    (void)rp;
    return BIAS_OK;
}

// ---- the whole sweep as ONE cooperative kernel: epochs serial, grid-wide barriers between the phases of an epoch ----
//
// Round 1 launched one pass per epoch from the host and read the birth queue (and the whole T x Kpad table, for the
// chain-split rule) back after each: >= 100 host round trips for 50,000 points, 6 ms per sweep.  Here the loop over the
// epochs runs on the device:
//   A  the items of epoch t in parallel, one warp per item (score_all / draw_parallel / apply_move of generic.cuh);
//      a draw of the new-component slot is queued                                     (src/RCRP.jl:191-306, 263-287, 366-377)
//   -- grid barrier --
//   B  while the queue is not empty: warp 0 of block 0 walks up to 256 queued items with the reference's sequential
//      semantics (birth_serial), barrier, the rest of the queue is re-sampled in parallel against the enlarged
//      component set into the other queue, barrier
//   C  chain splits (src/RCRP.jl:309-331), middle epochs only: EVERY block derives the same list of split chains from the
//      table (no barrier needed to agree on it), then the blocks move those chains' later items to the new components
//   -- grid barrier --
// and at the end the per-epoch concentrations are resampled, one thread per epoch (src/RCRP.jl:60-74: the update of
// aa[t] needs only the number of components epoch t uses, which no later epoch changes).
struct RcrpLoop {
    const int64_t *group_ptr;      // [T + 1] first item of every epoch
    int64_t N;
    int32_t *d_KK;                 // in: KK at the start of the sweep; out: at its end
    double *aa_t;                  // [T] per-epoch concentration (GenericParams::rcrp_aa points at the same array)
    int32_t *list_a, *list_b;      // the two birth queues (N entries each)
    unsigned int *cnt;             // [2] their lengths
    int *err;                      // 2: Kmax exhausted by a birth, 3: by a chain split
    int32_t Kmax, sample_hyper, n_internals;
    double a1, a2;
    uint64_t hyper_seed;
};

constexpr int kRcrpSerial = 256;   // queue entries walked serially before the rest is re-sampled in parallel

// one item against K components + the new-component slot
template <int KIND, int DATUM>
__device__ __forceinline__ void rcrp_sample_item(const GenericParams &p, const Philox &philox, int64_t i, int K, uint32_t stream,
                                                 double *sc, int lane, int32_t *queue, unsigned int *queue_len, double &diag_acc,
                                                 unsigned long long &moved_acc) {
    const int zold = __ldcg(p.z + i);                 // -1: queued by an earlier round
    const int epoch = p.row_of[i];
    int32_t *prow = p.prior_rows + (size_t)epoch * p.Kpad;
    const Datum d = load_datum<KIND, DATUM>(p, i, lane);
    const double lmax = score_all<KIND, DATUM>(p, d, zold, K, K + 1, prow, nullptr, __ldcg(p.rcrp_aa + epoch), sc, lane, epoch);
    uint32_t o[4];
    philox((uint32_t)i, (uint32_t)((uint64_t)i >> 32), p.sweep, stream, o);
    const int knew = draw_parallel(sc, K + 1, lmax, u01_double(o[0], o[1]), lane);
    __syncwarp();
    if (knew >= K) {
        if (zold >= 0) apply_move<KIND, DATUM>(p, d, zold, -1, prow, lane);
        if (lane == 0) {
            p.z[i] = -1;
            queue[atomicAdd(queue_len, 1u)] = (int32_t)i;
        }
        return;
    }
    if (knew != zold) {
        apply_move<KIND, DATUM>(p, d, zold, knew, prow, lane);
        if (lane == 0) {
            p.z[i] = knew;
            moved_acc += 1;
        }
    }
    const double dv = diag_after<KIND, DATUM>(p, d, knew, lane);
    if (lane == 0) diag_acc += dv;
}

template <int KIND, int DATUM>
__global__ void __launch_bounds__(256) rcrp_loop_kernel(const GenericParams p, const RcrpLoop q) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int sc_len = (q.Kmax + 4) & ~3;
    double *sc = reinterpret_cast<double *>(smem_raw) + (size_t)warp * sc_len;
    int32_t *split_to = reinterpret_cast<int32_t *>(smem_raw);       // [Kmax], phase C only (overlays warp 0's scores)
    __shared__ int s_nsplit;
    const int64_t gw = (int64_t)blockIdx.x * warps + warp, n_warps = (int64_t)gridDim.x * warps;
    const bool lead = blockIdx.x == 0 && warp == 0;
    const Philox philox(p.seed);
    const int T = p.rcrp_T, Kpad = p.Kpad;
    double diag_acc = 0.0;
    unsigned long long moved_acc = 0;
    int K = __ldcg(q.d_KK);
    int32_t *lcur = q.list_a, *lnext = q.list_b;
    int c = 0;                                          // q.cnt[c] is the length of lcur

    for (int t = 0; t < T; ++t) {
        const int64_t i0 = q.group_ptr[t], i1 = q.group_ptr[t + 1];
        // ---- A ----
        for (int64_t i = i0 + gw; i < i1; i += n_warps)
            rcrp_sample_item<KIND, DATUM>(p, philox, i, K, kStreamItemDraw, sc, lane, lcur, q.cnt + c, diag_acc, moved_acc);
        grid.sync();
        // ---- B ----
        for (uint32_t round = 1;; ++round) {
            const unsigned int n_pending = __ldcg(q.cnt + c);
            if (n_pending == 0) break;
            const int n_serial = (int)min(n_pending, (unsigned int)kRcrpSerial);
            if (lead) {
                K = birth_serial<KIND, DATUM>(p, lcur, n_serial, K, nullptr, 1.0, q.Kmax, q.err, sc, lane);
                if (lane == 0) {
                    *q.d_KK = K;
                    q.cnt[c ^ 1] = 0;
                }
            }
            grid.sync();
            K = __ldcg(q.d_KK);
            const int64_t rest = (int64_t)n_pending - n_serial;
            if (rest > 0) {
                for (int64_t w = gw; w < rest; w += n_warps)
                    rcrp_sample_item<KIND, DATUM>(p, philox, (int64_t)__ldcg(lcur + n_serial + w), K, kStreamBirth + 8u * round, sc, lane,
                                                  lnext, q.cnt + (c ^ 1), diag_acc, moved_acc);
                grid.sync();
            }
            int32_t *tmp = lcur;
            lcur = lnext;
            lnext = tmp;
            c ^= 1;
        }
        // ---- C ----
        if (t > 0 && t < T - 1) {
            __syncthreads();                             // the block's warps are done with their score buffers
            if (warp == 0) {
                const int32_t *row_t = p.prior_rows + (size_t)t * Kpad;
                int n_split = 0;
                for (int kb = 0; kb < K; kb += 32) {
                    const int k = kb + lane;
                    // a chain with a gap at t: empty here, alive at t + 1 and somewhere before t
                    bool gap = k < K && __ldcg(row_t + k) == 0 && __ldcg(row_t + Kpad + k) != 0;
                    if (gap) {
                        int64_t before = 0;
                        for (int tt = 0; tt < t; ++tt) before += __ldcg(p.prior_rows + (size_t)tt * Kpad + k);
                        gap = before != 0;
                    }
                    const unsigned bal = __ballot_sync(kFull, gap);
                    if (k < K) split_to[k] = gap ? K + n_split + __popc(bal & ((1u << lane) - 1u)) : -1;
                    n_split += __popc(bal);
                }
                if (n_split > 0 && K + n_split + 1 > q.Kmax) {
                    if (lead && lane == 0) *q.err = 3;
                    n_split = 0;
                }
                if (lane == 0) s_nsplit = n_split;
            }
            __syncthreads();
            const int n_split = s_nsplit;
            if (n_split > 0) {
                // everything a split chain holds from epoch t + 1 on moves to its new component
                for (int64_t base = i1 + gw * 32; base < q.N; base += n_warps * 32) {
                    const int64_t i = base + lane;
                    const int zf = i < q.N ? __ldcg(p.z + i) : -1;
                    const int kto = (zf >= 0 && zf < K) ? split_to[zf] : -1;
                    unsigned bal = __ballot_sync(kFull, kto >= 0);
                    while (bal) {
                        const int src = __ffs(bal) - 1;
                        bal &= bal - 1;
                        const int64_t ii = base + src;
                        const int kf = __shfl_sync(kFull, zf, src), kt = __shfl_sync(kFull, kto, src);
                        const Datum d = load_datum<KIND, DATUM>(p, ii, lane);
                        apply_move<KIND, DATUM>(p, d, kf, kt, p.prior_rows + (size_t)p.row_of[ii] * Kpad, lane);
                        if (lane == 0) p.z[ii] = kt;
                    }
                }
                K += n_split;
            }
            __syncthreads();                             // split_to is warp 0's score buffer again
            grid.sync();
        }
    }
    if (lead && lane == 0) *q.d_KK = K;
    if (lane == 0) {
        atomicAdd(p.diag_sum, diag_acc);
        if (moved_acc) atomicAdd(p.moved, moved_acc);
    }
    // ---- per-epoch concentration (src/RCRP.jl:60-74), one thread per epoch.  The reference's loop variable
    //      `for n = 1:iters` (:62) shadows the epoch size it was handed, so the "n" of the Escobar-West update is the
    //      iteration number; kept as is. ----
    if (q.sample_hyper && blockIdx.x == 0) {
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            int KK_t = 0;
            for (int k = 0; k < K; ++k) KK_t += __ldcg(p.prior_rows + (size_t)t * Kpad + k) != 0;
            PhiloxStream rs(q.hyper_seed, (uint32_t)t, 0u, p.sweep * 8u + kStreamHyper);
            double aa = q.aa_t[t];
            for (int it = 1; it <= q.n_internals; ++it) {
                const double n = (double)it;
                const double x1 = rs.gamma(aa + 1.0), x2 = rs.gamma(n);
                const double le = log(x1 / (x1 + x2));                       // log of Beta(aa + 1, n)
                const double rr = (q.a1 + KK_t - 1.0) / (n * (q.a2 - le));
                const double shape = (rs.u01() < rr / (1.0 + rr)) ? q.a1 + KK_t : q.a1 + KK_t - 1.0;
                if (shape > 0.0) aa = rs.gamma(shape) / (q.a2 - le);
            }
            q.aa_t[t] = aa;
        }
    }
}

template <int KIND, int DATUM>
static int launch_rcrp_loop(bias_model *m, GenericParams &p, RcrpLoop &q) {
    bias_ctx *ctx = m->ctx;
    // warps per block: as many score buffers of Kmax + 1 doubles as fit 200 KB, at most 8
    const size_t per_warp = (size_t)((m->hdp.Kmax + 4) & ~3) * sizeof(double);
    const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / per_warp));
    const size_t smem = per_warp * warps;
    if (smem > 227 * 1024) return fail(BIAS_ERANGE, "RCRP: Kmax = %d does not fit the score buffers", m->hdp.Kmax);
    if (smem > 48 * 1024)
        BIAS_CUDA_TRY(cudaFuncSetAttribute(rcrp_loop_kernel<KIND, DATUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    BIAS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rcrp_loop_kernel<KIND, DATUM>, warps * 32, smem));
    if (occ < 1) return fail(BIAS_ECUDA, "RCRP: the sweep kernel cannot be resident");
    // no more blocks than the longest epoch has work for: barriers cost by the block
    const int64_t want = (m->max_group_len + warps - 1) / warps;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->num_sms * std::min(occ, 2)));
    void *args[] = {(void *)&p, (void *)&q};
    BIAS_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)rcrp_loop_kernel<KIND, DATUM>, dim3(grid), dim3(warps * 32), args, smem, ctx->stream));
    ctx->launches += 1;
    return BIAS_OK;
}

int rcrp_sweep(bias_model *m) {
    bias_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t TT = m->G;
    BIAS_TRY(hdp_prune(m));                       // src/RCRP.jl:116-128 (and the in-loop deletions, once per sweep)
    BIAS_TRY(ensure_wide(m));
    m->narrow_valid = false;
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch, 0, 6 * sizeof(unsigned long long), st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_KK, &m->hdp.KK, 4, cudaMemcpyHostToDevice, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->d_aa_t, m->h_aa_t.data(), (size_t)TT * 8, cudaMemcpyHostToDevice, st));
    if (m->N > 0 && TT > 0) {
        GenericParams p = birth_params(m);
        p.seed = ctx->seed ^ m->seed_mix;
        RcrpLoop q{};
        q.group_ptr = m->group_ptr;
        q.N = m->N;
        q.d_KK = m->d_KK;
        q.aa_t = m->d_aa_t;
        q.list_a = m->pending;
        q.list_b = m->pending + std::max<int64_t>(1, m->N);
        q.cnt = reinterpret_cast<unsigned int *>(m->scratch + 3);
        q.err = reinterpret_cast<int *>(m->scratch + 4);
        q.Kmax = m->hdp.Kmax;
        q.sample_hyper = m->hdp.sample_hyperparam;
        q.n_internals = m->hdp.n_internals;
        q.a1 = m->hdp.a1;
        q.a2 = m->hdp.a2;
        q.hyper_seed = m->host_rng();             // the host stream (bias_model_set_host_seed) keys the device draws
        if (p.kind == BIAS_G1G1) BIAS_TRY((launch_rcrp_loop<BIAS_G1G1, BIAS_F64>(m, p, q)));
        else if (p.datum == BIAS_INT) BIAS_TRY((launch_rcrp_loop<BIAS_MD, BIAS_INT>(m, p, q)));
        else BIAS_TRY((launch_rcrp_loop<BIAS_MD, BIAS_SENT>(m, p, q)));
    }
    int32_t kk = 0;
    int err = 0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(&kk, m->d_KK, 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(&err, reinterpret_cast<int *>(m->scratch + 4), 4, cudaMemcpyDeviceToHost, st));
    if (m->hdp.sample_hyperparam)
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->h_aa_t.data(), m->d_aa_t, (size_t)TT * 8, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    m->hdp.KK = kk;
    if (err == 2)
        return fail(BIAS_ERANGE, "RCRP: Kmax = %d component slots exhausted (KK = %d); construct the model with a larger Kmax", m->hdp.Kmax, kk);
    if (err == 3) return fail(BIAS_ERANGE, "RCRP: Kmax = %d component slots exhausted while splitting a chain", m->hdp.Kmax);
    BIAS_TRY(hdp_prune(m));
    m->hdp.aa = m->h_aa_t[0];
    m->sweep_index += 1;
    m->stats_pending = true;
    return BIAS_OK;
}

}  // namespace bias

using namespace bias;

extern "C" {

int bias_model_get_hdp(bias_model *m, bias_hdp_params *params, double *beta) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (m->model != BIAS_MODEL_HDP && m->model != BIAS_MODEL_CRF && m->model != BIAS_MODEL_RCRF) return fail(BIAS_EINVAL, "not an HDP model");
    if (params) *params = m->hdp;
    if (beta) {
        const int n = m->hdp.Kmax + 1;
        for (int k = 0; k < n; ++k) beta[k] = (k < (int)m->h_beta.size()) ? m->h_beta[k] : 0.0;
    }
    return BIAS_OK;
}

// ---- sharded HDP: the phases of one sweep, with the caller's exchanges between them (include/bias_cuda.h) ----
static int check_hdp(bias_model *m) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (m->model != BIAS_MODEL_HDP) return fail(BIAS_EINVAL, "not an HDP model");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    return BIAS_OK;
}

int bias_model_set_host_seed(bias_model *m, uint64_t seed) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    m->host_rng.seed(seed);
    return BIAS_OK;
}

int bias_hdp_shard_pass(bias_model *m) {
    BIAS_TRY(check_hdp(m));
    return hdp_phase_pass(m);
}

int bias_hdp_shard_births(bias_model *m, int32_t *KK_out, double *beta_out) {
    BIAS_TRY(check_hdp(m));
    BIAS_TRY(hdp_phase_births(m));
    if (KK_out) *KK_out = m->hdp.KK;
    if (beta_out)
        for (int k = 0; k <= m->hdp.Kmax; ++k) beta_out[k] = (k < (int)m->h_beta.size()) ? m->h_beta[k] : 0.0;
    return BIAS_OK;
}

int bias_hdp_shard_adopt(bias_model *m, int32_t KK, const double *beta) {
    BIAS_TRY(check_hdp(m));
    if (!beta) return fail(BIAS_EINVAL, "null argument");
    if (KK < m->hdp.KK || KK + 2 > m->hdp.Kmax) return fail(BIAS_ERANGE, "HDP shard: KK = %d outside [%d, Kmax - 2 = %d]", KK, m->hdp.KK, m->hdp.Kmax - 2);
    m->hdp.KK = KK;
    for (int k = 0; k < (int)m->h_beta.size(); ++k) m->h_beta[k] = (k <= KK) ? beta[k] : 0.0;
    return BIAS_OK;
}

int bias_hdp_shard_prune(bias_model *m) {
    BIAS_TRY(check_hdp(m));
    return hdp_prune(m);
}

int bias_hdp_shard_tables(bias_model *m, int32_t internal, void **tables_dev, int64_t *n_tables, void **hyper_dev) {
    BIAS_TRY(check_hdp(m));
    BIAS_TRY(hdp_phase_tables(m, internal));
    if (tables_dev) *tables_dev = m->d_tables;
    if (n_tables) *n_tables = (int64_t)m->Kpad + 1;
    if (hyper_dev) *hyper_dev = m->d_hyper;
    return BIAS_OK;
}

int bias_hdp_shard_draw(bias_model *m) {
    BIAS_TRY(check_hdp(m));
    return hdp_phase_draw(m);
}

int bias_hdp_shard_end(bias_model *m) {
    BIAS_TRY(check_hdp(m));
    return hdp_phase_end(m);
}

int bias_crf_sweep_with_uniforms(bias_model *m, const double *uniforms) {
    if (!m || !uniforms) return fail(BIAS_EINVAL, "null argument");
    if (m->model != BIAS_MODEL_CRF && m->model != BIAS_MODEL_RCRF) return fail(BIAS_EINVAL, "not a CRF / RCRF model");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    double *d_u = nullptr;
    const size_t n = (size_t)std::max<int64_t>(1, m->N) * 3;
    BIAS_CUDA_TRY(cudaMalloc((void **)&d_u, n * 8));
    cudaError_t e = cudaMemcpyAsync(d_u, uniforms, (size_t)m->N * 3 * 8, cudaMemcpyHostToDevice, m->ctx->stream);
    int rc = (e != cudaSuccess) ? fail(BIAS_ECUDA, "%s", cudaGetErrorString(e))
                                : (m->model == BIAS_MODEL_RCRF ? rcrf_sweep(m, d_u) : crf_sweep(m, d_u));
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(d_u);
    return rc;
}

int bias_crf_get_tables(bias_model *m, int32_t *tji, int32_t *njt, int32_t *kjt, int32_t *n_tbls, int32_t *mm) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (m->model != BIAS_MODEL_CRF && m->model != BIAS_MODEL_RCRF) return fail(BIAS_EINVAL, "not a CRF / RCRF model");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    if (tji) BIAS_CUDA_TRY(cudaMemcpyAsync(tji, m->crf_tji, (size_t)m->N * 4, cudaMemcpyDeviceToHost, st));
    if (njt) BIAS_CUDA_TRY(cudaMemcpyAsync(njt, m->crf_njt, (size_t)m->N * 4, cudaMemcpyDeviceToHost, st));
    if (kjt) BIAS_CUDA_TRY(cudaMemcpyAsync(kjt, m->crf_kjt, (size_t)m->N * 4, cudaMemcpyDeviceToHost, st));
    if (n_tbls) BIAS_CUDA_TRY(cudaMemcpyAsync(n_tbls, m->crf_n_tbls, (size_t)m->G * 4, cudaMemcpyDeviceToHost, st));
    if (mm) {       // [T][KK] row-major (T = 1 for the CRF)
        const int64_t T = (m->model == BIAS_MODEL_RCRF) ? (int64_t)m->h_epoch_gptr.size() - 1 : 1;
        BIAS_CUDA_TRY(cudaMemcpy2DAsync(mm, (size_t)m->hdp.KK * 4, m->crf_mm, (size_t)m->Kpad * 4, (size_t)m->hdp.KK * 4, (size_t)T,
                                        cudaMemcpyDeviceToHost, st));
    }
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    return BIAS_OK;
}

int bias_model_get_rcrf(bias_model *m, int32_t *KK, double *gg, double *aa) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (m->model != BIAS_MODEL_RCRF) return fail(BIAS_EINVAL, "not an RCRF model");
    if (KK) *KK = m->hdp.KK;
    if (gg) std::copy(m->h_gg_t.begin(), m->h_gg_t.end(), gg);
    if (aa) std::copy(m->h_aa_t.begin(), m->h_aa_t.end(), aa);
    return BIAS_OK;
}

int bias_stirling_rows(bias_ctx *ctx, int32_t n_rows, double *out) {
    if (!ctx || !out) return fail(BIAS_EINVAL, "null argument");
    if (n_rows < 1 || n_rows > 10000) return fail(BIAS_EINVAL, "n_rows must be in [1, 10000]");
    BIAS_CUDA_TRY(cudaSetDevice(ctx->device));
    BIAS_TRY(ensure_stirling(ctx, n_rows));
    const size_t n = (size_t)n_rows * (n_rows + 1) / 2;
    BIAS_CUDA_TRY(cudaMemcpyAsync(out, ctx->stirling, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    BIAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BIAS_OK;
}

int bias_hdp_table_conditionals(bias_ctx *ctx, int64_t n_query, const int32_t *n, const double *a,
                                const double *uniforms, int32_t n_max, double *prob, int32_t *draws) {
    if (!ctx || !n || !a || !uniforms || !prob || !draws) return fail(BIAS_EINVAL, "null argument");
    if (n_query <= 0) return BIAS_OK;
    int32_t need = 1;
    for (int64_t q = 0; q < n_query; ++q) {
        if (n[q] < 1 || n[q] > n_max) return fail(BIAS_EINVAL, "n[%lld] = %d outside [1, n_max]", (long long)q, n[q]);
        if (n[q] > 10000) return fail(BIAS_ERANGE, "n = %d exceeds the 10k Stirling table (reference: BoundsError)", n[q]);
        if (!(a[q] > 0.0)) return fail(BIAS_EINVAL, "a must be > 0");
        need = std::max(need, n[q]);
    }
    BIAS_CUDA_TRY(cudaSetDevice(ctx->device));
    BIAS_TRY(ensure_stirling(ctx, need));
    cudaStream_t st = ctx->stream;
    int32_t *d_n = nullptr, *d_draw = nullptr;
    double *d_a = nullptr, *d_u = nullptr, *d_prob = nullptr;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void **)&d_n, (size_t)n_query * 4);
    A((void **)&d_draw, (size_t)n_query * 4);
    A((void **)&d_a, (size_t)n_query * 8);
    A((void **)&d_u, (size_t)n_query * 8);
    A((void **)&d_prob, (size_t)n_query * n_max * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_prob, 0, (size_t)n_query * n_max * 8, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, n, (size_t)n_query * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_a, a, (size_t)n_query * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, uniforms, (size_t)n_query * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        hdp_table_conditionals_kernel<<<grid1d(ctx, n_query * 32, 128), 128, 0, st>>>(d_n, d_a, d_u, n_query, n_max,
                                                                                      ctx->stirling, d_prob, d_draw);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(prob, d_prob, (size_t)n_query * n_max * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(draws, d_draw, (size_t)n_query * 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (void *p : {(void *)d_n, (void *)d_draw, (void *)d_a, (void *)d_u, (void *)d_prob})
        if (p) cudaFree(p);
    if (e != cudaSuccess) return fail(BIAS_ECUDA, "table conditionals: %s", cudaGetErrorString(e));
    return BIAS_OK;
}

// ---- verification of sample_hyperparam! (src/HDP.jl:124-159) and of the beta draw (:362-364) ----
// The device part: out[0] = sum_j log w_j with w_j ~ Beta(aa + 1, n_j), out[1] = sum_j s_j with s_j ~ Bernoulli(n_j / (n_j + aa)),
// for the Philox stream (seed of ctx, sweep, internal).  With uniforms[n_groups] supplied the Bernoulli draws use them
// (s_j = uniforms[j] < n_j / (n_j + aa)), which makes out[1] comparable exactly.
int bias_hdp_hyper_aux(bias_ctx *ctx, int64_t n_groups, const int64_t *group_ptr, double aa, uint32_t sweep, uint32_t internal,
                       const double *uniforms, double *out) {
    if (!ctx || !group_ptr || !out) return fail(BIAS_EINVAL, "null argument");
    if (n_groups < 0 || !(aa > 0.0)) return fail(BIAS_EINVAL, "n_groups >= 0 and aa > 0 required");
    BIAS_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int64_t *d_ptr = nullptr;
    double *d_u = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_ptr, ((size_t)n_groups + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, 16);
    if (e == cudaSuccess && uniforms) e = cudaMalloc((void **)&d_u, (size_t)std::max<int64_t>(1, n_groups) * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ptr, group_ptr, ((size_t)n_groups + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && uniforms) e = cudaMemcpyAsync(d_u, uniforms, (size_t)n_groups * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, 16, st);
    if (e == cudaSuccess && n_groups > 0) {
        hdp_hyper_kernel<<<grid1d(ctx, n_groups, 128, 4), 128, 0, st>>>(d_ptr, n_groups, aa, ctx->seed, sweep, internal, d_out, d_u);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, 16, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (void *p : {(void *)d_ptr, (void *)d_u, (void *)d_out})
        if (p) cudaFree(p);
    if (e != cudaSuccess) return fail(BIAS_ECUDA, "hyper auxiliaries: %s", cudaGetErrorString(e));
    return BIAS_OK;
}

// The host part: n successive draws of (alpha0, gamma) from the SAME conditioning values (the chain is not advanced:
// every draw starts from aa, gg), with the library's host generator seeded by `seed`.  No device is needed.
int bias_hdp_hyper_draws(uint64_t seed, double a1, double a2, double g1, double g2, double n_tables, double sum_logw, double sum_s,
                         int32_t KK, double aa, double gg, int32_t n, double *aa_out, double *gg_out) {
    if (!aa_out || !gg_out || n < 0) return fail(BIAS_EINVAL, "null argument");
    std::mt19937_64 rng(seed);
    for (int32_t i = 0; i < n; ++i) {
        double a = aa, g = gg;
        BIAS_TRY(host_hyper_draw(rng, a1, a2, g1, g2, n_tables, sum_logw, sum_s, KK, a, g));
        aa_out[i] = a;
        gg_out[i] = g;
    }
    return BIAS_OK;
}

// n draws of beta ~ Dirichlet([tables[0..KK-1], gg]) -> beta_out[n][KK + 1].  No device is needed.
int bias_hdp_beta_draws(uint64_t seed, int32_t KK, const int64_t *tables, double gg, int32_t n, double *beta_out) {
    if (!tables || !beta_out || KK < 1 || n < 0 || !(gg > 0.0)) return fail(BIAS_EINVAL, "bad argument");
    std::mt19937_64 rng(seed);
    std::vector<unsigned long long> t(tables, tables + KK);
    for (int32_t i = 0; i < n; ++i) host_beta_draw(rng, t.data(), KK, gg, beta_out + (size_t)i * (KK + 1));
    return BIAS_OK;
}

}  // extern "C"
