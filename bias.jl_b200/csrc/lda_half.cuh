// lda_half.cuh — LDA x MultinomialDirichlet x word-id sweep, two documents per warp (sm_100a).
//
// Same sampler as lda_fast.cuh (reference src/LDA.jl:113-133: delitem! -> LDA_sample_zz -> additem!),
// same data layout in HBM, same arithmetic (linear-domain fp32 weights, CDF in k order, strict
// `cw < r` walk of src/common.jl:74).  What changes is the mapping:
//
//   * a HALF-warp (16 lanes) owns a document; the two halves of a warp run two documents in lockstep.
//     The per-token bookkeeping (broadcasts, reductions, scans, ballots, the update of the document's
//     row) is one instruction stream for two tokens, so the warp-instruction count per token drops from
//     ~385 to ~240 and the trips through the LSU/shared-memory data pipe (every SHFL is one) from ~126 to
//     ~95 — the pipe ncu shows at 74 % for the one-document-per-warp kernel (profiles/r01_*).
//   * the token's own count is removed arithmetically: nothing is written to shared memory for the
//     removal, and nothing at all when the token keeps its topic.  r[k], n_dk[k] and the chunk sums R[j]
//     are rewritten only when the topic changes.
//   * n_dk is kept as uint16 (documents shorter than 65,536 tokens; longer ones use lda_fast.cuh), which
//     brings the per-document shared-memory state to 6 KB + 64 B at K = 1024: 32 documents per SM.
//
// Chunks are 64 topics wide (16 lanes x int4); a lane keeps the token's whole row in registers
// (NCK int4 = 64 registers at K = 1024) between the chunk pass and the in-chunk pass.
#pragma once
#include "lda_fast.cuh"

namespace bias {

constexpr int kHalfWarps = 8;       // default warps per CTA = 16 documents in flight per CTA

__host__ __device__ inline size_t lda_half_doc_bytes(int nck) { return (size_t)nck * 64 * 6 + 64; }
// 8 warps x 2 CTAs at K = 1024: 2 x (99,328 + 1,024 reserved) = 200,704 B = exactly the 196 KB shared-memory carve-out,
// which leaves 60 KB of L1.  That matters: the rows in flight (32 documents x 4 KB) pass through L1, and with the next
// carve-out (228 KB, 28 KB of L1) the same kernel runs 30 % slower (measured, profiles/r01b_*).
__host__ __device__ inline size_t lda_half_smem_bytes(int nck, int warps = kHalfWarps) {
    return (size_t)warps * 2 * lda_half_doc_bytes(nck);
}

#ifdef __CUDACC__

// Reduce N per-lane values across a 16-lane half so that lane l ends with the total of v[l >> (4 - log2 N)].
template <int N>
__device__ __forceinline__ float transpose_reduce_half(float (&v)[N], int l16) {
    int off = 8;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
        const bool upper = (l16 & off) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = upper ? v[i] : v[i + h];
            const float keep = upper ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, off, 16);
        }
        off >>= 1;
    }
    float x = v[0];
#pragma unroll
    for (; off >= 1; off >>= 1) x += __shfl_xor_sync(kFull, x, off, 16);
    return x;
}

__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o, 16);
    return v;
}

__device__ __forceinline__ float half_incl_scan(float v, int l16) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const float y = __shfl_up_sync(kFull, v, o, 16);
        if (l16 >= o) v += y;
    }
    return v;
}

template <int NCK>
__device__ __forceinline__ void prefetch_row_half(const int32_t *row, int l16) {
#pragma unroll
    for (int l = l16; l < NCK * 2; l += 16) prefetch_l2(row + l * 32);      // 128-byte lines
}

template <int NCK>
__device__ __forceinline__ int4 pick_chunk(const int4 (&c)[NCK], int j) {
    // j is uniform over the half: a jump table instead of a select chain over NCK x 4 registers
    int4 r = c[0];
    switch (j) {
        default: break;
        case 1: if constexpr (NCK > 1) r = c[1 % NCK]; break;
        case 2: if constexpr (NCK > 2) r = c[2 % NCK]; break;
        case 3: if constexpr (NCK > 2) r = c[3 % NCK]; break;
        case 4: if constexpr (NCK > 4) r = c[4 % NCK]; break;
        case 5: if constexpr (NCK > 4) r = c[5 % NCK]; break;
        case 6: if constexpr (NCK > 4) r = c[6 % NCK]; break;
        case 7: if constexpr (NCK > 4) r = c[7 % NCK]; break;
        case 8: if constexpr (NCK > 8) r = c[8 % NCK]; break;
        case 9: if constexpr (NCK > 8) r = c[9 % NCK]; break;
        case 10: if constexpr (NCK > 8) r = c[10 % NCK]; break;
        case 11: if constexpr (NCK > 8) r = c[11 % NCK]; break;
        case 12: if constexpr (NCK > 8) r = c[12 % NCK]; break;
        case 13: if constexpr (NCK > 8) r = c[13 % NCK]; break;
        case 14: if constexpr (NCK > 8) r = c[14 % NCK]; break;
        case 15: if constexpr (NCK > 8) r = c[15 % NCK]; break;
    }
    return r;
}

template <int NCK, int WARPS = kHalfWarps>
__global__ void __launch_bounds__(WARPS * 32, 2)
lda_md_int_sweep_half_kernel(const LdaFastParams p) {
    constexpr int LOG_NCK = (NCK == 2) ? 1 : (NCK == 4) ? 2 : (NCK == 8) ? 3 : 4;
    constexpr int GSHIFT = 4 - LOG_NCK;     // l16 >> GSHIFT = chunk held by the lane after the transposing reduction
    constexpr int KPAD = NCK * 64;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
    unsigned char *hbase = smem_raw + (size_t)(warp * 2 + half) * lda_half_doc_bytes(NCK);
    float *r_s = reinterpret_cast<float *>(hbase);
    uint16_t *cnt_s = reinterpret_cast<uint16_t *>(hbase + (size_t)KPAD * 4);
    float *R_s = reinterpret_cast<float *>(hbase + (size_t)KPAD * 6);
    int32_t *const my_nk = p.nk_rep + (blockIdx.x % kNkRep) * KPAD;

    const Philox philox(p.seed);
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    float diag_acc = 0.f;
    unsigned moved_acc = 0;
    int64_t t_cur = 0, t_end = 0;
    bool live = true;                       // this half may still be handed documents

    // Both halves execute every collective below (full mask, width 16); a half without work runs on dummy
    // operands (word 0, topic 0) and its stores are predicated off.
    for (;;) {
        // ---- a half whose document is finished fetches the next one and rebuilds its row ----
        bool need = live && t_cur >= t_end;
        bool fetched = false;
        while (__any_sync(kFull, need)) {
            unsigned long long doc = 0;
            if (l16 == 0 && need) doc = atomicAdd(p.doc_counter, 1ull);
            doc = __shfl_sync(kFull, doc, 0, 16);
            if (need) {
                if (doc >= (unsigned long long)p.D) {
                    live = false;
                    need = false;
                } else {
                    t_cur = p.doc_ptr[doc];
                    t_end = p.doc_ptr[doc + 1];
                    need = t_end <= t_cur;
                    fetched = !need;
                }
            }
        }
        if (__any_sync(kFull, fetched)) {
            // n_dk from z, then r = (n_dk + alpha) / (n_k + V beta) and the chunk sums R
            if (fetched)
                for (int q = l16; q < KPAD / 8; q += 16) reinterpret_cast<uint4 *>(cnt_s)[q] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (fetched)
                for (int64_t t = t_cur + l16; t < t_end; t += 16) {
                    const int k = p.z[t];
                    atomicAdd(reinterpret_cast<unsigned *>(cnt_s) + (k >> 1), 1u << ((k & 1) * 16));
                }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NCK; ++j) {
                const int k0 = j * 64 + l16 * 4;
                int4 nk = __ldcg(reinterpret_cast<const int4 *>(p.n_k) + j * 16 + l16);
#pragma unroll
                for (int r = 0; r < kNkRep; ++r) {
                    const int4 d = __ldcg(reinterpret_cast<const int4 *>(p.nk_rep + r * KPAD) + j * 16 + l16);
                    nk.x += d.x; nk.y += d.y; nk.z += d.z; nk.w += d.w;
                }
                const uint2 cc = reinterpret_cast<const uint2 *>(cnt_s)[j * 16 + l16];
                float4 rv;
                rv.x = (k0 + 0 < p.K) ? __fdividef((float)(cc.x & 0xffffu) + alpha, (float)nk.x + vbeta) : 0.f;
                rv.y = (k0 + 1 < p.K) ? __fdividef((float)(cc.x >> 16) + alpha, (float)nk.y + vbeta) : 0.f;
                rv.z = (k0 + 2 < p.K) ? __fdividef((float)(cc.y & 0xffffu) + alpha, (float)nk.z + vbeta) : 0.f;
                rv.w = (k0 + 3 < p.K) ? __fdividef((float)(cc.y >> 16) + alpha, (float)nk.w + vbeta) : 0.f;
                const float s = half_sum((rv.x + rv.y) + (rv.z + rv.w));
                if (fetched) {
                    reinterpret_cast<float4 *>(r_s)[j * 16 + l16] = rv;
                    if (l16 == 0) R_s[j] = s;
                }
            }
            __syncwarp();
        }
        const bool first_batch = __any_sync(kFull, fetched);
        const int nb = (int)min((int64_t)16, t_end - t_cur);      // 0 when this half has run out of documents
        const int nb_both = max(nb, __shfl_xor_sync(kFull, nb, 16));
        if (nb_both == 0) break;

        // ---- up to 16 tokens of each half's document ----
        int my_w = 0, my_z = 0;
        float my_u = 0.5f;
        if (l16 < nb) {
            my_w = p.word[t_cur + l16];
            my_z = p.z[t_cur + l16];
            const uint64_t tid = (uint64_t)(t_cur + l16);
            if (p.ext_uniforms) {
                my_u = (float)p.ext_uniforms[tid];
            } else {
                uint32_t o[4];
                philox((uint32_t)tid, (uint32_t)(tid >> 32), p.sweep, kStreamItemDraw, o);
                my_u = u01_float(o[0]);
            }
        }
        int my_new = my_z;
        // words of the following batch, so that rows can be pulled into L2 across the batch boundary
        int nxt_w = 0;
        const int rem = (int)min((int64_t)64, t_end - t_cur);          // tokens of this document still ahead (capped)
        if (16 + l16 < rem) nxt_w = p.word[t_cur + 16 + l16];
        const int pf = p.pf_dist;

        // operands and row of the first token
        int w = __shfl_sync(kFull, my_w, 0, 16);
        int kold = __shfl_sync(kFull, my_z, 0, 16);
        float u = __shfl_sync(kFull, my_u, 0, 16);
        int4 c[NCK];
        {
            const int32_t *row32 = p.n_wk + (size_t)w * KPAD;
#pragma unroll
            for (int j = 0; j < NCK; ++j) c[j] = __ldcg(reinterpret_cast<const int4 *>(row32) + j * 16 + l16);
        }
        int c_own = __ldcg(p.n_wk + (size_t)w * KPAD + kold);
        if (first_batch) {
            // a new document: nothing was prefetched for its first rows
            for (int d = 1; d < pf; ++d) {
                const int wd = __shfl_sync(kFull, my_w, d & 15, 16);
                if (d < nb) prefetch_row_half<NCK>(p.n_wk + (size_t)wd * KPAD, l16);
            }
        }

        for (int i = 0; i < nb_both; ++i) {
            const bool on = i < nb;
            int32_t *row32 = p.n_wk + (size_t)w * KPAD;

            // ---- delitem!, arithmetically: r[kold] with n_dk and n_k one lower; nothing is stored ----
            const int jo = kold >> 6;
            const int cn = cnt_s[kold];
            const float r_old = r_s[kold];
            const float inv = __fdividef(r_old, (float)cn + alpha);            // 1/(n_k + V beta)
            const float inv_rm = __fdividef(inv, 1.f - inv);                   // 1/(n_k - 1 + V beta)
            const float r_rm = ((float)(cn - 1) + alpha) * inv_rm;

            // ---- pass 1: chunk sums of n_wk * r (beta * sum r is added per chunk) ----
            float acc[NCK];
#pragma unroll
            for (int j = 0; j < NCK; ++j) {
                const float4 rv = reinterpret_cast<const float4 *>(r_s)[j * 16 + l16];
                float a = (float)c[j].x * rv.x;
                a = fmaf((float)c[j].y, rv.y, a);
                a = fmaf((float)c[j].z, rv.z, a);
                a = fmaf((float)c[j].w, rv.w, a);
                acc[j] = a;
            }
            float T = transpose_reduce_half<NCK>(acc, l16);                    // lane l: chunk (l >> GSHIFT)
            const int myj = l16 >> GSHIFT;
            T = fmaf(beta, R_s[myj], T);
            if (myj == jo) {
                // the row and r still count this token: swap its term for the removed one
                const float co = (float)c_own;
                T -= (co + beta) * r_old - (fmaxf(co - 1.f, 0.f) + beta) * r_rm;
            }
            T = fmaxf(T, 0.f);
            float P = T;                                                       // inclusive scan in chunk order
#pragma unroll
            for (int d = 1; d < NCK; d <<= 1) {
                const float y = __shfl_up_sync(kFull, P, d << GSHIFT, 16);
                if (myj >= d) P += y;
            }
            const float total = __shfl_sync(kFull, P, 15, 16);
            const float target = u * total;
            unsigned bal = (__ballot_sync(kFull, P >= target) >> (half * 16)) & 0xffffu;
            const int src_lane = bal ? (__ffs(bal) - 1) : 15;
            const int jsel = src_lane >> GSHIFT;
            const float base = __shfl_sync(kFull, P - T, src_lane, 16);
            const int4 cs = pick_chunk<NCK>(c, jsel);

            // ---- the next token's operands and row: in flight during pass 2 and the update ----
            const int inext = (i + 1) & 15;
            const int w_n = __shfl_sync(kFull, my_w, inext, 16);
            const int kold_n = __shfl_sync(kFull, my_z, inext, 16);
            const float u_n = __shfl_sync(kFull, my_u, inext, 16);
            int c_own_n = 0;
            if (i + 1 < nb_both) {
                const int32_t *rown = p.n_wk + (size_t)w_n * KPAD;
#pragma unroll
                for (int j = 0; j < NCK; ++j) c[j] = __ldcg(reinterpret_cast<const int4 *>(rown) + j * 16 + l16);
                c_own_n = __ldcg(rown + kold_n);
                if (pf > 0) {
                    const int ip = i + pf;
                    const int wa = __shfl_sync(kFull, my_w, ip & 15, 16), wb = __shfl_sync(kFull, nxt_w, ip & 15, 16);
                    if (ip < rem) prefetch_row_half<NCK>(p.n_wk + (size_t)(ip < 16 ? wa : wb) * KPAD, l16);
                }
            }

            // ---- pass 2: CDF inside the selected chunk ----
            const float4 rv = reinterpret_cast<const float4 *>(r_s)[jsel * 16 + l16];
            const int kb = jsel * 64 + l16 * 4;
            float n0 = (float)cs.x, n1 = (float)cs.y, n2 = (float)cs.z, n3 = (float)cs.w;
            float q0 = rv.x, q1 = rv.y, q2 = rv.z, q3 = rv.w;
            // own token excluded; the clamp only matters if another SM's update is caught mid-flight
            if (kold == kb + 0) { n0 = fmaxf(n0 - 1.f, 0.f); q0 = r_rm; }
            if (kold == kb + 1) { n1 = fmaxf(n1 - 1.f, 0.f); q1 = r_rm; }
            if (kold == kb + 2) { n2 = fmaxf(n2 - 1.f, 0.f); q2 = r_rm; }
            if (kold == kb + 3) { n3 = fmaxf(n3 - 1.f, 0.f); q3 = r_rm; }
            const float w0 = (n0 + beta) * q0, w1 = (n1 + beta) * q1;
            const float w2 = (n2 + beta) * q2, w3 = (n3 + beta) * q3;
            const float s = (w0 + w1) + (w2 + w3);
            const float incl = base + half_incl_scan(s, l16);
            bal = (__ballot_sync(kFull, incl >= target) >> (half * 16)) & 0xffffu;
            const int lsel = bal ? (__ffs(bal) - 1) : 15;
            // walk the (at most) four entries of the selected lane
            float cw = (incl - s) + w0;
            int csel = 0;
            float nsel = n0;
            if (cw < target) { csel = 1; nsel = n1; cw += w1;
                if (cw < target) { csel = 2; nsel = n2; cw += w2;
                    if (cw < target) { csel = 3; nsel = n3; } } }
            int knew = __shfl_sync(kFull, kb + csel, lsel, 16);
            nsel = __shfl_sync(kFull, nsel, lsel, 16);
            if (knew >= p.K) {           // rounding pushed the target past the last real topic
                knew = p.K - 1;
                nsel = -1.f;             // diagnostic recomputed below from memory
            }
            if (on && l16 == i) my_new = knew;

            const bool moved = on && !p.frozen && knew != kold;
            if (!p.frozen) {
                // ---- additem!: the document's row changes only when the topic does ----
                const int cn2 = cnt_s[knew];
                const float r2 = r_s[knew];
                const float inv2 = __fdividef(r2, (float)cn2 + alpha);
                float inv_new = __fdividef(inv2, 1.f + inv2);                      // 1/(n_k + 1 + V beta)
                const float r_new = ((float)(cn2 + 1) + alpha) * inv_new;
                if (knew == kold) inv_new = inv;
                __syncwarp();
                if (moved) {
                    // the document's row: lane 0; the global counts: ONE fire-and-forget RED instruction with four active
                    // lanes (a single lane issuing four atomics back to back sits on the token's critical path)
                    if (l16 == 0) {
                        cnt_s[kold] = (uint16_t)(cn - 1);
                        r_s[kold] = r_rm;
                        R_s[jo] += r_rm - r_old;
                        cnt_s[knew] = (uint16_t)(cn2 + 1);
                        r_s[knew] = r_new;
                        R_s[knew >> 6] += r_new - r2;
                    }
                    if (l16 >= 8 && l16 < 12) {
                        // lanes 8..11: n_wk[kold]--, n_wk[knew]++, n_k[kold]--, n_k[knew]++
                        const int which = l16 - 8, kk = (which & 1) ? knew : kold;
                        red_add((which < 2 ? row32 : my_nk) + kk, (which & 1) ? 1 : -1);
                    }
                    if (l16 == 0) moved_acc += 1;
                }
                if (l16 == 0 && on) {
                    // loglikelihood(components[kk], x) after the add: posterior mean of the word
                    if (nsel < 0.f) nsel = (float)__ldcg(row32 + knew) - (knew == kold ? 1.f : 0.f);
                    diag_acc += (nsel + 1.f + beta) * inv_new;
                }
                // the next token's row was read before this update: when it is the same word, read it again
                if (__any_sync(kFull, moved && w_n == w && i + 1 < nb)) {
                    if (l16 >= 8 && l16 < 12) __threadfence();
                    __syncwarp();
                    if (moved && w_n == w && i + 1 < nb) {
                        const int32_t *rown = p.n_wk + (size_t)w_n * KPAD;
#pragma unroll
                        for (int j = 0; j < NCK; ++j) c[j] = __ldcg(reinterpret_cast<const int4 *>(rown) + j * 16 + l16);
                        c_own_n = __ldcg(rown + kold_n);
                    }
                }
                __syncwarp();
            }
            w = w_n;
            kold = kold_n;
            u = u_n;
            c_own = c_own_n;
        }  // tokens of the batch

        if (p.frozen) {
            if (l16 < nb) p.draws_out[t_cur + l16] = my_new;
        } else if (l16 < nb && my_new != my_z) {
            p.z[t_cur + l16] = my_new;
        }
        t_cur += nb;
    }  // batches

    if (!p.frozen && l16 == 0) {
        atomicAdd(p.diag_sum, (double)diag_acc);
        atomicAdd(p.moved, (unsigned long long)moved_acc);
    }
}

#endif  // __CUDACC__

}  // namespace bias
