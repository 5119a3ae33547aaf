// lda_fast.cuh — the headline kernel: LDA x MultinomialDirichlet x word-id sweep for sm_100a.
//
// Replaces the inner loop of collapsed_gibbs_sampler!(::LDA) (reference src/LDA.jl:113-133):
//   delitem! (src/conjugates.jl:158-162) -> LDA_sample_zz (src/LDA.jl:26-42: K scores,
//   lognormalize! src/common.jl:42-52, sample src/common.jl:69-79) -> additem! (:153-157).
//
// Mapping: one warp owns one document at a time and visits its tokens serially, so the
// document-topic counts n_dk see every earlier move of the same sweep exactly as in the
// reference; documents run in parallel against the shared word-topic table with
// immediately-visible atomics (AD-LDA style staleness across documents only).
//
// Arithmetic: the reference's exp(log(n_dk+a) + log((b+n_kw)/(bV+n_k)) - max) / sum is the
// normalised weight (n_dk+a)(n_kw+b)/(n_k+bV); it is evaluated in the linear domain in fp32
// (no log/exp: the SFU ceiling would sit below the HBM roofline).
//
// Data layout in HBM:
//   n_wk  int32 [V][Kpad]  word-major: one token's row is one contiguous 4*Kpad-byte span,
//                          read with coalesced 16-byte loads (512 B per warp instruction);
//   n_k   int32 [Kpad];    word int32[N], z int32[N] in document order; doc_ptr int64[D+1].
//   The D x K document-topic table is never stored: a warp rebuilds its document's row in
//   shared memory from z (a few instructions per token), which removes 4*K bytes/token of
//   traffic and 4*D*K bytes of HBM.
// Per-warp shared memory: r[k] = (n_dk+alpha)/(n_k+V*beta) (fp32), cnt[k] = n_dk (int32),
//   R[j] = sum of r over 128-wide chunk j.
// Per token: weight_k = (n_wk + beta) * r[k]; chunk totals by a transposing warp reduction,
//   a scan over chunks picks the chunk, a warp-shuffle inclusive scan inside that chunk forms
//   the CDF in k order (same order as the reference's linear walk) and a ballot finds the
//   first k with cdf >= u * total (the reference's strict `cw < r` test).
#pragma once
#include "common.cuh"

namespace bias {

struct LdaFastParams {
    const int64_t *doc_ptr;
    const int32_t *word;
    int32_t *z;
    int32_t *n_wk;              // [V][Kpad]
    int32_t *n_k;               // [Kpad]
    int64_t D;
    int32_t K, Kpad, V;
    float alpha, beta, vbeta;
    unsigned long long *doc_counter;   // dynamic document scheduler
    uint64_t seed;
    uint32_t sweep;
    const double *ext_uniforms; // nullable: uniforms[N] supplied by the caller (verification)
    int32_t *draws_out;         // frozen mode: draws[N]
    int32_t frozen;             // 1: do not touch z or counts
    int32_t eval;               // 1: no sampling; accumulate sum_t log p(w_t | theta_d, phi) into eval_sum
    double *eval_sum;
    double *diag_sum;           // sum of posterior-mean probabilities (src/conjugates.jl:205)
    unsigned long long *moved;  // tokens whose assignment changed
    const int64_t *eval_ptr;    // eval mode, nullable: CSR of held-out tokens per document (document completion)
    const int32_t *eval_word;   //   their word ids; theta_d still comes from the document's training tokens (z)
    uint16_t *n_wk16;           // lda_half16.cuh: the narrow tables, [V + 2 * hot_cap][Kpad]
    const int32_t *word_hot;    // lda_half16.cuh: [V] hot-row index of each word or -1
    int32_t hot_cap, hot_rep;   // lda_half16.cuh: copy c of hot row h is row V + 2 (h + c * hot_cap); hot_rep copies (narrow.cuh: kHotRep)
    int32_t *nk_rep;            // half-warp kernels: [kNkRep][Kpad] replicas collecting this sweep's changes of n_k
    int32_t pf_dist;            // lda_half.cuh: rows are pulled into L2 this many tokens ahead (0: off; at most 16)
};

// n_k is Kpad addresses that every document in flight on the GPU updates twice per moved token: RED traffic on those
// few L2 lines costs 8 ms of a 47 ms sweep (the hot slices back up and delay the row loads too).  The half-warp kernels
// therefore send their n_k updates to one of kNkRep zero-initialised replicas (chosen by CTA index), read
// n_k + sum of the replicas when a document starts, and a small kernel folds the replicas into n_k after the sweep.
constexpr int kNkRep = 8;


// warps per CTA: 8 up to K = 1024, 4 above (the per-warp rows are 8*Kpad bytes of shared memory)
__host__ __device__ constexpr int lda_fast_warps(int nch) { return nch <= 8 ? 8 : 4; }
__host__ __device__ constexpr int lda_fast_min_ctas(int nch) { return nch <= 8 ? 3 : 1; }

__host__ __device__ inline size_t lda_fast_smem_bytes(int nch) {
    // per warp: r[Kpad] float + cnt[Kpad] int32 + R[Kpad/128] float (padded to 32 floats)
    return (size_t)lda_fast_warps(nch) * ((size_t)nch * 128 * 8 + 128);
}

#ifdef __CUDACC__

// Reduce N per-lane values across the warp so that lane l ends with the total of
// v[l >> (5 - log2 N)].  N/2 + N/4 + ... + 1 shuffles for the transposing steps, then one
// shuffle per remaining butterfly level (9 shuffles for N = 8 instead of 40).
template <int N>
__device__ __forceinline__ float transpose_reduce(float (&v)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            float send = upper ? v[i] : v[i + h];
            float keep = upper ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, off);
        }
        off >>= 1;
    }
    float x = v[0];
#pragma unroll
    for (; off >= 1; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
    return x;
}

template <int NCH>
__device__ __forceinline__ void prefetch_row(const int32_t *row, int lane) {
#pragma unroll
    for (int l = lane; l < NCH * 4; l += 32) prefetch_l2(row + l * 32);     // one 128-byte line per lane
}

template <int NCH>
__global__ void __launch_bounds__(lda_fast_warps(NCH) * 32, lda_fast_min_ctas(NCH))
lda_md_int_sweep_kernel(const LdaFastParams p) {
    constexpr bool KEEP_ROW = NCH <= 8;     // the token's row stays in registers between the two passes
    constexpr int LOG_NCH = (NCH == 1) ? 0 : (NCH == 2) ? 1 : (NCH == 4) ? 2 : (NCH == 8) ? 3 : (NCH == 16) ? 4 : 5;
    constexpr int GSHIFT = 5 - LOG_NCH;     // lane >> GSHIFT = chunk held by the lane after transpose_reduce
    constexpr int KPAD = NCH * 128;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem_raw + (size_t)warp * ((size_t)KPAD * 8 + 128);
    float *r_s = reinterpret_cast<float *>(wbase);
    int32_t *cnt_s = reinterpret_cast<int32_t *>(wbase + (size_t)KPAD * 4);
    float *R_s = reinterpret_cast<float *>(wbase + (size_t)KPAD * 8);

    const Philox philox(p.seed);
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    float diag_acc = 0.f;
    unsigned moved_acc = 0;
    double eval_acc = 0.0, eval_log_norm = 0.0;

    for (;;) {
        // ---- fetch the next document (one atomic per document) ----
        unsigned long long doc = 0;
        if (lane == 0) doc = atomicAdd(p.doc_counter, 1ull);
        doc = __shfl_sync(kFull, doc, 0);
        if (doc >= (unsigned long long)p.D) break;
        const int64_t t_begin = p.doc_ptr[doc], t_end = p.doc_ptr[doc + 1];
        // tokens that are scored / resampled: the document's own, or its held-out tokens in evaluation mode
        const int64_t s_begin = p.eval_ptr ? p.eval_ptr[doc] : t_begin, s_end = p.eval_ptr ? p.eval_ptr[doc + 1] : t_end;
        const int32_t *s_word = p.eval_ptr ? p.eval_word : p.word;
        if (s_end <= s_begin) continue;
        eval_log_norm = (double)logf((float)(t_end - t_begin) + (float)p.K * alpha);

        // ---- rebuild n_dk from z, then r = (n_dk + alpha) / (n_k + V beta) ----
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            reinterpret_cast<int4 *>(cnt_s)[j * 32 + lane] = make_int4(0, 0, 0, 0);
        __syncwarp();
        for (int64_t t = t_begin + lane; t < t_end; t += 32) atomicAdd(&cnt_s[p.z[t]], 1);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int k0 = j * 128 + lane * 4;
            int4 nk = __ldcg(reinterpret_cast<const int4 *>(p.n_k) + j * 32 + lane);
            int4 c = reinterpret_cast<int4 *>(cnt_s)[j * 32 + lane];
            float4 rv;
            rv.x = (k0 + 0 < p.K) ? __fdividef((float)c.x + alpha, (float)nk.x + vbeta) : 0.f;
            rv.y = (k0 + 1 < p.K) ? __fdividef((float)c.y + alpha, (float)nk.y + vbeta) : 0.f;
            rv.z = (k0 + 2 < p.K) ? __fdividef((float)c.z + alpha, (float)nk.z + vbeta) : 0.f;
            rv.w = (k0 + 3 < p.K) ? __fdividef((float)c.w + alpha, (float)nk.w + vbeta) : 0.f;
            reinterpret_cast<float4 *>(r_s)[j * 32 + lane] = rv;
            float s = warp_sum((rv.x + rv.y) + (rv.z + rv.w));
            if (lane == 0) R_s[j] = s;
        }
        __syncwarp();

        // ---- tokens of the document, 32 at a time ----
        for (int64_t t0 = s_begin; t0 < s_end; t0 += 32) {
            const int nb = (int)min((int64_t)32, s_end - t0);
            int my_w = 0, my_z = 0;
            float my_u = 0.5f;
            if (lane < nb) {
                my_w = s_word[t0 + lane];
                my_z = p.eval_ptr ? 0 : p.z[t0 + lane];
                const uint64_t tid = (uint64_t)(t0 + lane);
                if (p.ext_uniforms) {
                    my_u = (float)p.ext_uniforms[tid];
                } else {
                    uint32_t o[4];
                    philox((uint32_t)tid, (uint32_t)(tid >> 32), p.sweep, kStreamItemDraw, o);
                    my_u = u01_float(o[0]);
                }
            }
            int my_new = my_z;
            // warm L2 with the first two rows of the batch
            {
                const int w0 = __shfl_sync(kFull, my_w, 0);
                prefetch_row<NCH>(p.n_wk + (size_t)w0 * KPAD, lane);
                if (nb > 1) {
                    const int w1 = __shfl_sync(kFull, my_w, 1);
                    prefetch_row<NCH>(p.n_wk + (size_t)w1 * KPAD, lane);
                }
            }

            for (int i = 0; i < nb; ++i) {
                const int w = __shfl_sync(kFull, my_w, i);
                const int kold = __shfl_sync(kFull, my_z, i);
                const float u = __shfl_sync(kFull, my_u, i);
                const int4 *row = reinterpret_cast<const int4 *>(p.n_wk + (size_t)w * KPAD);

                // issue the row loads first (NCH x 512 B per warp in flight), then do the removal.
                // ld.global.cg: the table is mutated by other SMs' atomics, L1 must not serve it.
                int4 c[KEEP_ROW ? NCH : 1];
                if constexpr (KEEP_ROW) {
#pragma unroll
                    for (int j = 0; j < NCH; ++j) c[j] = __ldcg(row + j * 32 + lane);
                }
                if (i + 2 < nb) {
                    const int w2 = __shfl_sync(kFull, my_w, i + 2);
                    prefetch_row<NCH>(p.n_wk + (size_t)w2 * KPAD, lane);
                }

                // ---- delitem!: n_dk[kold] -= 1, n_k[kold] -= 1 folded into r[kold] ----
                const int jo = kold >> 7;
                float r_saved = 0.f, R_saved = 0.f, r_rm = 0.f;
                if (lane == 0 && !p.eval) {
                    const int cn = cnt_s[kold];
                    r_saved = r_s[kold];
                    R_saved = R_s[jo];
                    const float inv = __fdividef(r_saved, (float)cn + alpha);      // 1/(n_k + V beta)
                    const float inv_rm = __fdividef(inv, 1.f - inv);               // 1/(n_k - 1 + V beta)
                    r_rm = ((float)(cn - 1) + alpha) * inv_rm;
                    cnt_s[kold] = cn - 1;
                    r_s[kold] = r_rm;
                    R_s[jo] = R_saved + (r_rm - r_saved);
                }
                r_rm = __shfl_sync(kFull, r_rm, 0);
                __syncwarp();

                // ---- pass 1: chunk sums of n_wk * r (beta * sum r is added per chunk) ----
                float acc[NCH];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const float4 rv = reinterpret_cast<const float4 *>(r_s)[j * 32 + lane];
                    int4 cj;
                    if constexpr (KEEP_ROW) cj = c[j]; else cj = __ldcg(row + j * 32 + lane);
                    float a = (float)cj.x * rv.x;
                    a = fmaf((float)cj.y, rv.y, a);
                    a = fmaf((float)cj.z, rv.z, a);
                    a = fmaf((float)cj.w, rv.w, a);
                    acc[j] = a;
                }
                float T = transpose_reduce<NCH>(acc, lane);      // lane l: chunk (l >> GSHIFT)
                const int myj = lane >> GSHIFT;
                T = fmaf(beta, R_s[myj], T);
                if (myj == jo && !p.eval) T -= r_rm;             // the row still counts this token
                T = fmaxf(T, 0.f);
                // inclusive scan over chunks (in chunk order)
                float P = T;
#pragma unroll
                for (int d = 1; d < NCH; d <<= 1) {
                    const float y = __shfl_up_sync(kFull, P, d << GSHIFT);
                    if (myj >= d) P += y;
                }
                const float total = __shfl_sync(kFull, P, 31);
                if (p.eval) {
                    // point-estimate predictive: sum_k theta_dk phi_kw = total / (L_d + K alpha)
                    if (lane == 0) eval_acc += (double)logf(total) - eval_log_norm;
                    continue;
                }
                const float target = u * total;
                unsigned bal = __ballot_sync(kFull, P >= target);
                int src_lane = bal ? (__ffs(bal) - 1) : 31;
                const int jsel = src_lane >> GSHIFT;
                const float base = __shfl_sync(kFull, P - T, src_lane);

                // ---- pass 2: CDF inside the selected chunk ----
                int4 cs;
                if constexpr (KEEP_ROW) {
                    // jsel is warp-uniform: a jump table instead of a 4 x (NCH-1) select chain
                    switch (jsel) {
                        default: cs = c[0]; break;
                        case 1: if constexpr (NCH > 1) cs = c[1 % NCH]; break;
                        case 2: if constexpr (NCH > 2) cs = c[2 % NCH]; break;
                        case 3: if constexpr (NCH > 2) cs = c[3 % NCH]; break;
                        case 4: if constexpr (NCH > 4) cs = c[4 % NCH]; break;
                        case 5: if constexpr (NCH > 4) cs = c[5 % NCH]; break;
                        case 6: if constexpr (NCH > 4) cs = c[6 % NCH]; break;
                        case 7: if constexpr (NCH > 4) cs = c[7 % NCH]; break;
                    }
                } else {
                    cs = __ldcg(row + jsel * 32 + lane);
                }
                const float4 rv = reinterpret_cast<const float4 *>(r_s)[jsel * 32 + lane];
                const int kb = jsel * 128 + lane * 4;
                float n0 = (float)cs.x, n1 = (float)cs.y, n2 = (float)cs.z, n3 = (float)cs.w;
                // own token excluded; the clamp only matters if another SM's update is caught mid-flight
                if (kold == kb + 0) n0 = fmaxf(n0 - 1.f, 0.f);
                if (kold == kb + 1) n1 = fmaxf(n1 - 1.f, 0.f);
                if (kold == kb + 2) n2 = fmaxf(n2 - 1.f, 0.f);
                if (kold == kb + 3) n3 = fmaxf(n3 - 1.f, 0.f);
                const float w0 = (n0 + beta) * rv.x, w1 = (n1 + beta) * rv.y;
                const float w2 = (n2 + beta) * rv.z, w3 = (n3 + beta) * rv.w;
                const float s = (w0 + w1) + (w2 + w3);
                const float incl = base + warp_incl_scan(s, lane);
                bal = __ballot_sync(kFull, incl >= target);
                const int lsel = bal ? (__ffs(bal) - 1) : 31;
                // walk the (at most) four entries of the selected lane
                float cw = (incl - s) + w0;
                int csel = 0;
                float nsel = n0;
                if (cw < target) { csel = 1; nsel = n1; cw += w1;
                    if (cw < target) { csel = 2; nsel = n2; cw += w2;
                        if (cw < target) { csel = 3; nsel = n3; } } }
                int knew = __shfl_sync(kFull, kb + csel, lsel);
                nsel = __shfl_sync(kFull, nsel, lsel);
                if (knew >= p.K) {           // rounding pushed the target past the last real topic
                    knew = p.K - 1;
                    nsel = -1.f;             // diagnostic recomputed below from memory
                }

                if (p.frozen) {
                    // restore the removal, record the draw
                    if (lane == 0) {
                        cnt_s[kold] += 1;
                        r_s[kold] = r_saved;
                        R_s[jo] = R_saved;
                    }
                    if (lane == i) my_new = knew;
                    __syncwarp();
                    continue;
                }

                // ---- additem! ----
                if (lane == 0) {
                    float inv_new;
                    if (knew == kold) {
                        cnt_s[kold] += 1;
                        r_s[kold] = r_saved;
                        R_s[jo] = R_saved;
                        inv_new = __fdividef(r_saved, (float)cnt_s[kold] + alpha);
                    } else {
                        const int cn = cnt_s[knew];
                        const float r_old = r_s[knew];
                        const float inv = __fdividef(r_old, (float)cn + alpha);
                        inv_new = __fdividef(inv, 1.f + inv);                  // 1/(n_k + 1 + V beta)
                        const float r_new = ((float)(cn + 1) + alpha) * inv_new;
                        cnt_s[knew] = cn + 1;
                        r_s[knew] = r_new;
                        R_s[knew >> 7] += r_new - r_old;
                        atomicAdd(p.n_wk + (size_t)w * KPAD + kold, -1);
                        atomicAdd(p.n_wk + (size_t)w * KPAD + knew, 1);
                        atomicAdd(p.n_k + kold, -1);
                        atomicAdd(p.n_k + knew, 1);
                        moved_acc += 1;
                    }
                    // loglikelihood(components[kk], x) after the add: posterior mean of the word
                    if (nsel < 0.f) nsel = (float)__ldcg(p.n_wk + (size_t)w * KPAD + knew) - (knew == kold ? 1.f : 0.f);
                    diag_acc += (nsel + 1.f + beta) * inv_new;
                }
                if (lane == i) my_new = knew;
                __syncwarp();
            }  // tokens of the batch

            if (p.eval) {
                // nothing to write back
            } else if (p.frozen) {
                if (lane < nb) p.draws_out[t0 + lane] = my_new;
            } else if (lane < nb && my_new != my_z) {
                p.z[t0 + lane] = my_new;
            }
        }  // batches
    }  // documents

    if (p.eval) {
        if (lane == 0) atomicAdd(p.eval_sum, eval_acc);
        return;
    }
    if (!p.frozen) {
        // lane 0 of each warp holds the warp's diagnostics
        if (lane == 0) {
            atomicAdd(p.diag_sum, (double)diag_acc);
            atomicAdd(p.moved, (unsigned long long)moved_acc);
        }
    }
}

#endif  // __CUDACC__

}  // namespace bias
