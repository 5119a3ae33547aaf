// lda_tma.cuh — LDA x MultinomialDirichlet x word-id sweep, rows staged in shared memory by
// bulk asynchronous copies (cp.async.bulk + mbarrier: SASS UBLKCP / SYNCS), sm_100a.
//
// Same algorithm and arithmetic as lda_fast.cuh (reference src/LDA.jl:113-133; see that header for
// the derivation).  What changes is how a token's 4*Kpad-byte count row reaches the warp:
//   * each warp owns a two-stage ring of row buffers in shared memory; while it scores token i from
//     stage i&1, one elected lane has already issued the bulk copy of token i+1's row into the other
//     stage, completion tracked by a per-stage mbarrier (expect_tx / try_wait.parity).  The HBM/L2
//     latency of the gather is therefore overlapped with a full token of arithmetic without holding
//     the row in registers (the register version is latency-bound at 16 warps/SM, profiles/r01_*);
//   * both passes read the row with conflict-free LDS.128, so the second pass indexes the selected
//     chunk directly (no register select chain);
//   * n_dk is kept as packed uint16 in shared memory (documents shorter than 65,536 tokens; longer
//     ones take the register kernel), which lets 16 warps x 14 KB fit in one 227 KB CTA per SM.
// A token whose word equals the previous token's word reuses the row already in shared memory,
// patched with the warp's own move, instead of racing a new copy against its own atomics.
#pragma once
#include "lda_fast.cuh"

namespace bias {

constexpr int kTmaWarps = 16;
__host__ __device__ constexpr size_t lda_tma_warp_bytes(int nch) {
    // 2 row stages (int32) + r (fp32) + cnt (uint16) + R[32] floats + 2 mbarriers (+pad)
    return (size_t)nch * 128 * (4 * 2 + 4 + 2) + 128 + 64;
}
__host__ __device__ constexpr size_t lda_tma_smem_bytes(int nch) { return (size_t)kTmaWarps * lda_tma_warp_bytes(nch); }

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int NCH>
__global__ void __launch_bounds__(kTmaWarps * 32, 1)
lda_md_int_sweep_tma_kernel(const LdaFastParams p) {
    constexpr int LOG_NCH = (NCH == 1) ? 0 : (NCH == 2) ? 1 : (NCH == 4) ? 2 : 3;
    constexpr int GSHIFT = 5 - LOG_NCH;
    constexpr int KPAD = NCH * 128;
    constexpr uint32_t ROW_BYTES = KPAD * 4;
    static_assert(NCH <= 8, "the staged kernel covers K <= 1024");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem_raw + (size_t)warp * lda_tma_warp_bytes(NCH);
    int32_t *row_s0 = reinterpret_cast<int32_t *>(wbase);
    int32_t *row_s1 = reinterpret_cast<int32_t *>(wbase + ROW_BYTES);
    float *r_s = reinterpret_cast<float *>(wbase + 2 * ROW_BYTES);
    uint16_t *cnt_s = reinterpret_cast<uint16_t *>(wbase + 3 * ROW_BYTES);
    float *R_s = reinterpret_cast<float *>(wbase + 3 * ROW_BYTES + KPAD * 2);
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + 3 * ROW_BYTES + KPAD * 2 + 128);
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    const uint32_t row0_a = smem_u32(row_s0), row1_a = smem_u32(row_s1);

    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t ph0 = 0, ph1 = 0;          // phase parity of each stage
    int stage = 0;                      // stage holding (or receiving) the current token's row

    const Philox philox(p.seed);
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    float diag_acc = 0.f;
    unsigned moved_acc = 0;

    for (;;) {
        unsigned long long doc = 0;
        if (lane == 0) doc = atomicAdd(p.doc_counter, 1ull);
        doc = __shfl_sync(kFull, doc, 0);
        if (doc >= (unsigned long long)p.D) break;
        const int64_t t_begin = p.doc_ptr[doc], t_end = p.doc_ptr[doc + 1];
        if (t_end <= t_begin) continue;

        // first row of the document: start the copy before rebuilding n_dk
        int w_first = 0;
        if (lane == 0) {
            w_first = p.word[t_begin];
            const uint32_t bar = stage ? bar1 : bar0;
            mbar_expect_tx(bar, ROW_BYTES);
            bulk_g2s(stage ? row1_a : row0_a, p.n_wk + (size_t)w_first * KPAD, ROW_BYTES, bar);
        }

        // ---- rebuild n_dk (packed uint16) from z, then r = (n_dk + alpha) / (n_k + V beta) ----
#pragma unroll
        for (int j = 0; j < NCH / 2 + (NCH == 1); ++j)
            if (j * 256 + lane * 8 < KPAD) reinterpret_cast<int4 *>(cnt_s)[j * 32 + lane] = make_int4(0, 0, 0, 0);
        __syncwarp();
        for (int64_t t = t_begin + lane; t < t_end; t += 32) {
            const int k = p.z[t];
            atomicAdd(reinterpret_cast<unsigned *>(cnt_s) + (k >> 1), 1u << ((k & 1) * 16));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int k0 = j * 128 + lane * 4;
            const int4 nk = __ldcg(reinterpret_cast<const int4 *>(p.n_k) + j * 32 + lane);
            const uint2 cp = reinterpret_cast<const uint2 *>(cnt_s)[j * 32 + lane];
            const float c0 = (float)(cp.x & 0xffffu), c1 = (float)(cp.x >> 16);
            const float c2 = (float)(cp.y & 0xffffu), c3 = (float)(cp.y >> 16);
            float4 rv;
            rv.x = (k0 + 0 < p.K) ? __fdividef(c0 + alpha, (float)nk.x + vbeta) : 0.f;
            rv.y = (k0 + 1 < p.K) ? __fdividef(c1 + alpha, (float)nk.y + vbeta) : 0.f;
            rv.z = (k0 + 2 < p.K) ? __fdividef(c2 + alpha, (float)nk.z + vbeta) : 0.f;
            rv.w = (k0 + 3 < p.K) ? __fdividef(c3 + alpha, (float)nk.w + vbeta) : 0.f;
            reinterpret_cast<float4 *>(r_s)[j * 32 + lane] = rv;
            const float s = warp_sum((rv.x + rv.y) + (rv.z + rv.w));
            if (lane == 0) R_s[j] = s;
        }
        __syncwarp();

        bool row_in_flight = true;       // a copy into `stage` has been issued and not yet waited for
        int w_prev = -1;                 // word whose row currently sits in `stage` (after the wait)
        for (int64_t t0 = t_begin; t0 < t_end; t0 += 32) {
            const int nb = (int)min((int64_t)32, t_end - t0);
            int my_w = 0, my_z = 0;
            float my_u = 0.5f;
            if (lane < nb) {
                my_w = p.word[t0 + lane];
                my_z = p.z[t0 + lane];
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
            // first word of the NEXT batch, so the ring stays full across batch boundaries
            int w_next_batch = -1;
            if (t0 + 32 < t_end) w_next_batch = p.word[t0 + 32];

            for (int i = 0; i < nb; ++i) {
                const int w = __shfl_sync(kFull, my_w, i);
                const int kold = __shfl_sync(kFull, my_z, i);
                const float u = __shfl_sync(kFull, my_u, i);
                const int w_next = (i + 1 < nb) ? __shfl_sync(kFull, my_w, i + 1) : w_next_batch;

                // ---- make sure this token's row is (being) staged in `stage` ----
                if (!row_in_flight && w != w_prev) {
                    if (lane == 0) {
                        const uint32_t bar = stage ? bar1 : bar0;
                        mbar_expect_tx(bar, ROW_BYTES);
                        bulk_g2s(stage ? row1_a : row0_a, p.n_wk + (size_t)w * KPAD, ROW_BYTES, bar);
                    }
                    row_in_flight = true;
                }
                // ---- prefetch the next token's row into the other stage ----
                const bool next_is_new = (w_next >= 0) && (w_next != w);
                if (next_is_new && lane == 0) {
                    const uint32_t bar = stage ? bar0 : bar1;
                    mbar_expect_tx(bar, ROW_BYTES);
                    bulk_g2s(stage ? row0_a : row1_a, p.n_wk + (size_t)w_next * KPAD, ROW_BYTES, bar);
                }

                // ---- delitem!: n_dk[kold] -= 1, n_k[kold] -= 1 folded into r[kold] ----
                const int jo = kold >> 7;
                float r_saved = 0.f, R_saved = 0.f, r_rm = 0.f;
                if (lane == 0) {
                    const int cn = cnt_s[kold];
                    r_saved = r_s[kold];
                    R_saved = R_s[jo];
                    const float inv = __fdividef(r_saved, (float)cn + alpha);
                    const float inv_rm = __fdividef(inv, 1.f - inv);
                    r_rm = ((float)(cn - 1) + alpha) * inv_rm;
                    cnt_s[kold] = (uint16_t)(cn - 1);
                    r_s[kold] = r_rm;
                    R_s[jo] = R_saved + (r_rm - r_saved);
                }
                r_rm = __shfl_sync(kFull, r_rm, 0);

                // ---- wait for the row ----
                if (row_in_flight) {
                    if (stage) { mbar_wait(bar1, ph1); ph1 ^= 1; } else { mbar_wait(bar0, ph0); ph0 ^= 1; }
                    row_in_flight = false;
                }
                __syncwarp();
                const int32_t *row_s = stage ? row_s1 : row_s0;

                // ---- pass 1 ----
                float acc[NCH];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const float4 rv = reinterpret_cast<const float4 *>(r_s)[j * 32 + lane];
                    const int4 cj = reinterpret_cast<const int4 *>(row_s)[j * 32 + lane];
                    float a = (float)cj.x * rv.x;
                    a = fmaf((float)cj.y, rv.y, a);
                    a = fmaf((float)cj.z, rv.z, a);
                    a = fmaf((float)cj.w, rv.w, a);
                    acc[j] = a;
                }
                float T = transpose_reduce<NCH>(acc, lane);
                const int myj = lane >> GSHIFT;
                T = fmaf(beta, R_s[myj], T);
                if (myj == jo) T -= r_rm;
                T = fmaxf(T, 0.f);
                float P = T;
#pragma unroll
                for (int d = 1; d < NCH; d <<= 1) {
                    const float y = __shfl_up_sync(kFull, P, d << GSHIFT);
                    if (myj >= d) P += y;
                }
                const float total = __shfl_sync(kFull, P, 31);
                const float target = u * total;
                unsigned bal = __ballot_sync(kFull, P >= target);
                const int src_lane = bal ? (__ffs(bal) - 1) : 31;
                const int jsel = src_lane >> GSHIFT;
                const float base = __shfl_sync(kFull, P - T, src_lane);

                // ---- pass 2 ----
                const int4 cs = reinterpret_cast<const int4 *>(row_s)[jsel * 32 + lane];
                const float4 rv = reinterpret_cast<const float4 *>(r_s)[jsel * 32 + lane];
                const int kb = jsel * 128 + lane * 4;
                float n0 = (float)cs.x, n1 = (float)cs.y, n2 = (float)cs.z, n3 = (float)cs.w;
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
                float cw = (incl - s) + w0;
                int csel = 0;
                float nsel = n0;
                if (cw < target) { csel = 1; nsel = n1; cw += w1;
                    if (cw < target) { csel = 2; nsel = n2; cw += w2;
                        if (cw < target) { csel = 3; nsel = n3; } } }
                int knew = __shfl_sync(kFull, kb + csel, lsel);
                nsel = __shfl_sync(kFull, nsel, lsel);
                if (knew >= p.K) {
                    knew = p.K - 1;
                    nsel = -1.f;
                }

                if (p.frozen) {
                    if (lane == 0) {
                        cnt_s[kold] += 1;
                        r_s[kold] = r_saved;
                        R_s[jo] = R_saved;
                    }
                    if (lane == i) my_new = knew;
                } else {
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
                            inv_new = __fdividef(inv, 1.f + inv);
                            const float r_new = ((float)(cn + 1) + alpha) * inv_new;
                            cnt_s[knew] = (uint16_t)(cn + 1);
                            r_s[knew] = r_new;
                            R_s[knew >> 7] += r_new - r_old;
                            atomicAdd(p.n_wk + (size_t)w * KPAD + kold, -1);
                            atomicAdd(p.n_wk + (size_t)w * KPAD + knew, 1);
                            atomicAdd(p.n_k + kold, -1);
                            atomicAdd(p.n_k + knew, 1);
                            moved_acc += 1;
                            if (!next_is_new && w_next == w) {
                                // the next token reuses this staged row: apply the warp's own move to it
                                int32_t *rw = stage ? row_s1 : row_s0;
                                rw[kold] -= 1;
                                rw[knew] += 1;
                                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            }
                        }
                        if (nsel < 0.f) nsel = (float)row_s[knew] - (knew == kold ? 1.f : 0.f);
                        diag_acc += (nsel + 1.f + beta) * inv_new;
                    }
                    if (lane == i) my_new = knew;
                }
                __syncwarp();
                w_prev = w;
                if (next_is_new) {
                    stage ^= 1;
                    row_in_flight = true;
                }
            }  // tokens of the batch

            if (p.frozen) {
                if (lane < nb) p.draws_out[t0 + lane] = my_new;
            } else if (lane < nb && my_new != my_z) {
                p.z[t0 + lane] = my_new;
            }
        }  // batches
        // leave the ring idle between documents: nothing in flight here (the last token never prefetches)
        (void)w_first;
    }  // documents

    if (!p.frozen && lane == 0) {
        atomicAdd(p.diag_sum, (double)diag_acc);
        atomicAdd(p.moved, (unsigned long long)moved_acc);
    }
}

#endif  // __CUDACC__

}  // namespace bias
