// lda_half16.cuh — LDA x MultinomialDirichlet x word-id sweep over the narrow count tables (sm_100a): the headline kernel.
//
// Same sampler, mapping and arithmetic as lda_half.cuh (two documents per warp, reference src/LDA.jl:113-133);
// what changes is where a token's row comes from.  On this path the word-topic counts live in the NARROW TABLES:
// uint16 [V][Kpad] rows for every word whose corpus frequency is at most 65,535 (no cell of such a row can leave 16 bits)
// and int32 rows, appended to the same allocation, for the few HOT words above that (see classify_words_kernel).  The
// tables are the canonical state while sweeps run: nothing is rebuilt or folded per sweep.  The int32 table n_wk
// [V][Kpad] that the generic kernels, the verification entry points and the exports read is a VIEW that model.cu
// refreshes on demand (narrow_fold_kernel) and converts back when somebody wrote to it (narrow_build_kernel).
//
// A 16-bit row is 2*Kpad bytes — half the L2->SM bytes and half the LSU wavefronts of an int32 row, 32 instead of
// 64 registers per lane, and a 102 MB table at K = 1024, V = 50k that the 126 MB L2 can hold.  A hot row is 4*Kpad
// bytes, fetched one token ahead like a 16-bit row (same address pattern for its first half, 32 more registers for the
// second: the instantiation with the route runs 2 CTAs x 8 warps per SM at up to 128 registers).  The two halves of a warp may take different routes through the loads and the
// int->float conversions; every shuffle, ballot and shared-memory update is common code.
//
// Chunks are 128 topics wide (16 lanes x 8 uint16 = one LDG.128 per lane); r is stored in shared memory with
// the two float4 halves of a lane's 8 topics de-interleaved, so that both LDS.128 of a chunk are conflict-free.
#pragma once
#include <type_traits>
#include "lda_half.cuh"
#include "narrow.cuh"

namespace bias {

#ifndef BIAS_H16_WARPS
#define BIAS_H16_WARPS 6
#endif
#ifndef BIAS_H16_CTAS
#define BIAS_H16_CTAS 3
#endif
constexpr int kH16Warps = BIAS_H16_WARPS;        // warps per CTA = 12 documents per CTA, 3 CTAs per SM
constexpr int kH16Ctas = BIAS_H16_CTAS;
#ifndef BIAS_H16_WARPS8
#define BIAS_H16_WARPS8 6
#endif
constexpr int kH16Warps8 = BIAS_H16_WARPS8;
// the instantiation with the hot-row route keeps a whole 4 KB row in registers: 2 CTAs of 8 warps at up to 128 registers
// (6 x 3 at 112 registers spills in the hot route: 81.6 ms per Zipf sweep against 55.7)
constexpr int kH16HotWarps = 8, kH16HotCtas = 2;      // warps per CTA when the document counts are bytes (CNT8): the same
                                                 // 36 documents per SM in 186 KB, which leaves 32 KB of L1 (7 x 3 spills)

// per document: r (fp32) + the document's topic counts (uint16, or uint8 when no document has more than 255 tokens)
// + the per-chunk sums of r
__host__ __device__ inline size_t lda_half16_doc_bytes(int nc8, bool cnt8 = false) { return (size_t)nc8 * 128 * (cnt8 ? 5 : 6) + 64; }
__host__ __device__ inline size_t lda_half16_smem_bytes(int nc8, int warps = kH16Warps, bool cnt8 = false) {
    return (size_t)warps * 2 * lda_half16_doc_bytes(nc8, cnt8);
}

#ifdef __CUDACC__

// freq[w] = occurrences of word w in the local shard (row sums of the freshly built local counts)
__global__ void word_freq_kernel(const int32_t *n_wk, int64_t V, int Kpad, int32_t *freq) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < V; w += nw) {
        const int4 *src = reinterpret_cast<const int4 *>(n_wk + (size_t)w * Kpad);
        long long s = 0;
        for (int q = lane; q < Kpad / 4; q += 32) {
            const int4 a = __ldcg(src + q);
            s += (long long)a.x + a.y + a.z + a.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        if (lane == 0) freq[w] = (int32_t)min(s, (long long)INT32_MAX);
    }
}

template <int N>
__device__ __forceinline__ uint4 pick_chunk16(const uint4 (&c)[N], int j) {
    uint4 r = c[0];
    switch (j) {
        default: break;
        case 1: if constexpr (N > 1) r = c[1 % N]; break;
        case 2: if constexpr (N > 2) r = c[2 % N]; break;
        case 3: if constexpr (N > 2) r = c[3 % N]; break;
        case 4: if constexpr (N > 4) r = c[4 % N]; break;
        case 5: if constexpr (N > 4) r = c[5 % N]; break;
        case 6: if constexpr (N > 4) r = c[6 % N]; break;
        case 7: if constexpr (N > 4) r = c[7 % N]; break;
    }
    return r;
}

// Two fp32 values in one 64-bit register: sm_100's packed fp32 instructions (FFMA2 / FADD2 / FMUL2) do two topics per
// issue slot.  cvt2: both 16-bit counts of a word to floats WITHOUT the conversion unit — a byte permute puts each count
// under the exponent of 2^23 (0x4B00xxxx = 2^23 + x exactly) and one packed add takes the 2^23 off again.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 cvt2(unsigned x) {
    const unsigned lo = __byte_perm(x, 0x4B000000u, 0x7410), hi = __byte_perm(x, 0x4B000000u, 0x7432);
    f32x2 f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(f) : "r"(lo), "r"(hi));
    return fadd2(f, pack2(-8388608.f, -8388608.f));
}

__device__ __forceinline__ float dot8(const uint4 x, const float4 a, const float4 b) {
    f32x2 s = fmul2(cvt2(x.x), pack2(a.x, a.y));
    s = ffma2(cvt2(x.y), pack2(a.z, a.w), s);
    s = ffma2(cvt2(x.z), pack2(b.x, b.y), s);
    s = ffma2(cvt2(x.w), pack2(b.z, b.w), s);
    float lo, hi;
    unpack2(s, lo, hi);
    return lo + hi;
}

// the same eight products for a hot row: a = cells of topics +0..3, b = +4..7 (int32, stored in r_pos order)
__device__ __forceinline__ float dot8i(const uint4 a, const uint4 b, const float4 ra, const float4 rb) {
    f32x2 s = fmul2(pack2((float)(int)a.x, (float)(int)a.y), pack2(ra.x, ra.y));
    s = ffma2(pack2((float)(int)a.z, (float)(int)a.w), pack2(ra.z, ra.w), s);
    s = ffma2(pack2((float)(int)b.x, (float)(int)b.y), pack2(rb.x, rb.y), s);
    s = ffma2(pack2((float)(int)b.z, (float)(int)b.w), pack2(rb.z, rb.w), s);
    float lo, hi;
    unpack2(s, lo, hi);
    return lo + hi;
}

// A hot row held in registers: c = pieces 0 .. N-1, d = pieces N .. 2N-1 (piece jj of lane l: topics
// (jj >> 1) * 128 + l * 8 + (jj & 1) * 4 .. + 3).
template <int N, int I>
__device__ __forceinline__ uint4 row_piece(const uint4 (&c)[N], const uint4 (&d)[N]) {
    if constexpr (I < N) return c[I]; else return d[I - N];
}
// chunk sums of the whole row
template <int N, int J = 0>
__device__ __forceinline__ void hot_chunk_sums(const uint4 (&c)[N], const uint4 (&d)[N], const float4 *r4, int l16, float (&acc)[N]) {
    if constexpr (J < N) {
        acc[J] = dot8i(row_piece<N, 2 * J>(c, d), row_piece<N, 2 * J + 1>(c, d), r4[J * 32 + l16], r4[J * 32 + 16 + l16]);
        hot_chunk_sums<N, J + 1>(c, d, r4, l16, acc);
    }
}
// the two pieces of chunk j; j is uniform over the half
template <int N, int J = 0>
__device__ __forceinline__ void pick_pair(const uint4 (&c)[N], const uint4 (&d)[N], int j, uint4 &a, uint4 &b) {
    if constexpr (J < N) {
        if (j == J || J == N - 1) {
            a = row_piece<N, 2 * J>(c, d);
            b = row_piece<N, 2 * J + 1>(c, d);
        } else {
            pick_pair<N, J + 1>(c, d, j, a, b);
        }
    }
}

// HOT = false: the corpus has no word above 65,535 occurrences (decided on the host after every (re)load) and the
// hot-row route is compiled out.  It is a separate instantiation because the route's mere presence in the token loop
// slows the 16-bit route by 20 % (37.6 -> 45.9 ms per 100 M tokens, profiles/r02_hot_route_variants.md).
template <int NC8, int WARPS = kH16Warps, int CTAS = kH16Ctas, bool CNT8 = false, bool HOT = false>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
lda_md_int_sweep_half16_kernel(const LdaFastParams p) {
    using cnt_t = typename std::conditional<CNT8, uint8_t, uint16_t>::type;
    constexpr int CB = (int)sizeof(cnt_t);
    constexpr int LOG_NC = (NC8 == 1) ? 0 : (NC8 == 2) ? 1 : (NC8 == 4) ? 2 : 3;
    constexpr int GSHIFT = 4 - LOG_NC;      // l16 >> GSHIFT = chunk held by the lane after the transposing reduction
    constexpr int KPAD = NC8 * 128;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
    unsigned char *hbase = smem_raw + (size_t)(warp * 2 + half) * lda_half16_doc_bytes(NC8, CNT8);
    float *r_s = reinterpret_cast<float *>(hbase);                                  // permuted, see r_pos
    cnt_t *cnt_s = reinterpret_cast<cnt_t *>(hbase + (size_t)KPAD * 4);
    float *R_s = reinterpret_cast<float *>(hbase + (size_t)KPAD * (4 + CB));
    const float4 *r4 = reinterpret_cast<const float4 *>(r_s);

    int32_t *const my_nk = p.nk_rep + (blockIdx.x % kNkRep) * KPAD;

    const Philox philox(p.seed);
    const float alpha = p.alpha, beta = p.beta, vbeta = p.vbeta;
    float diag_acc = 0.f;
    unsigned moved_acc = 0;
    int64_t t_cur = 0, t_end = 0;
    bool live = true;

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
            if (fetched)
                for (int q = l16; q < KPAD * CB / 16; q += 16) reinterpret_cast<uint4 *>(cnt_s)[q] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (fetched)
                for (int64_t t = t_cur + l16; t < t_end; t += 16) {
                    const int k = p.z[t];
                    if (CNT8) atomicAdd(reinterpret_cast<unsigned *>(cnt_s) + (k >> 2), 1u << ((k & 3) * 8));
                    else atomicAdd(reinterpret_cast<unsigned *>(cnt_s) + (k >> 1), 1u << ((k & 1) * 16));
                }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NC8; ++j) {
                const int k0 = j * 128 + l16 * 8;
                int4 na = __ldcg(reinterpret_cast<const int4 *>(p.n_k) + j * 32 + l16 * 2);
                int4 nb4 = __ldcg(reinterpret_cast<const int4 *>(p.n_k) + j * 32 + l16 * 2 + 1);
#pragma unroll
                for (int r = 0; r < kNkRep; ++r) {
                    const int4 da = __ldcg(reinterpret_cast<const int4 *>(p.nk_rep + r * KPAD) + j * 32 + l16 * 2);
                    const int4 db = __ldcg(reinterpret_cast<const int4 *>(p.nk_rep + r * KPAD) + j * 32 + l16 * 2 + 1);
                    na.x += da.x; na.y += da.y; na.z += da.z; na.w += da.w;
                    nb4.x += db.x; nb4.y += db.y; nb4.z += db.z; nb4.w += db.w;
                }
                float cf[8];             // the document's counts of topics k0 .. k0 + 7
                if (CNT8) {
                    const uint2 cc = reinterpret_cast<const uint2 *>(cnt_s)[j * 16 + l16];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        cf[e] = (float)((cc.x >> (8 * e)) & 0xffu);
                        cf[4 + e] = (float)((cc.y >> (8 * e)) & 0xffu);
                    }
                } else {
                    const uint4 cc = reinterpret_cast<const uint4 *>(cnt_s)[j * 16 + l16];
                    cf[0] = (float)(cc.x & 0xffffu); cf[1] = (float)(cc.x >> 16); cf[2] = (float)(cc.y & 0xffffu); cf[3] = (float)(cc.y >> 16);
                    cf[4] = (float)(cc.z & 0xffffu); cf[5] = (float)(cc.z >> 16); cf[6] = (float)(cc.w & 0xffffu); cf[7] = (float)(cc.w >> 16);
                }
                float4 ra, rb;
                ra.x = (k0 + 0 < p.K) ? __fdividef(cf[0] + alpha, (float)na.x + vbeta) : 0.f;
                ra.y = (k0 + 1 < p.K) ? __fdividef(cf[1] + alpha, (float)na.y + vbeta) : 0.f;
                ra.z = (k0 + 2 < p.K) ? __fdividef(cf[2] + alpha, (float)na.z + vbeta) : 0.f;
                ra.w = (k0 + 3 < p.K) ? __fdividef(cf[3] + alpha, (float)na.w + vbeta) : 0.f;
                rb.x = (k0 + 4 < p.K) ? __fdividef(cf[4] + alpha, (float)nb4.x + vbeta) : 0.f;
                rb.y = (k0 + 5 < p.K) ? __fdividef(cf[5] + alpha, (float)nb4.y + vbeta) : 0.f;
                rb.z = (k0 + 6 < p.K) ? __fdividef(cf[6] + alpha, (float)nb4.z + vbeta) : 0.f;
                rb.w = (k0 + 7 < p.K) ? __fdividef(cf[7] + alpha, (float)nb4.w + vbeta) : 0.f;
                const float s = half_sum(((ra.x + ra.y) + (ra.z + ra.w)) + ((rb.x + rb.y) + (rb.z + rb.w)));
                if (fetched) {
                    reinterpret_cast<float4 *>(r_s)[j * 32 + l16] = ra;
                    reinterpret_cast<float4 *>(r_s)[j * 32 + 16 + l16] = rb;
                    if (l16 == 0) R_s[j] = s;
                }
            }
            __syncwarp();
        }
        const int nb = (int)min((int64_t)16, t_end - t_cur);      // 0 when this half has run out of documents
        const int nb_both = max(nb, __shfl_xor_sync(kFull, nb, 16));
        if (nb_both == 0) break;

        // ---- up to 16 tokens of each half's document ----
        // my_w is the token's ROW in the narrow tables: the word id, or V + 2h for the h-th hot word
        int my_w = 0, my_z = 0;
        float my_u = 0.5f;
        if (l16 < nb) {
            my_w = p.word[t_cur + l16];
            if constexpr (HOT) {
                const int h = __ldg(p.word_hot + my_w);
                if (h >= 0) my_w = p.V + 2 * (h + (int)(blockIdx.x % (unsigned)p.hot_rep) * p.hot_cap);     // this CTA's copy of the row
            }
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

        // Visiting order inside the batch.  Without hot words: index order.  With them: the tokens on 16-bit rows first,
        // then the hot ones (each group in index order), so that the two documents of the warp are mostly on the same
        // route at the same time — a step whose halves differ pays for both routes.  (The reference visits the tokens
        // of a document in randperm order, src/LDA.jl:114; any fixed order is one of its orders.)
        int pos = l16;
        int src = 0;
        if constexpr (HOT) {
            const bool my_hot = l16 < nb && my_w >= p.V;
            const unsigned hm = (__ballot_sync(kFull, my_hot) >> (half * 16)) & 0xffffu;
            const unsigned cm = ((nb >= 16) ? 0xffffu : ((1u << nb) - 1u)) & ~hm;
            const unsigned lt = (1u << l16) - 1u;
            if (l16 < nb) pos = my_hot ? __popc(cm) + __popc(hm & lt) : __popc(cm & lt);
            const unsigned b0 = (__ballot_sync(kFull, pos == 0) >> (half * 16)) & 0xffffu;
            src = __ffs(b0) - 1;
        }

        // operands and row of the first token
        int w = __shfl_sync(kFull, my_w, src, 16);
        int kold = __shfl_sync(kFull, my_z, src, 16);
        float u = __shfl_sync(kFull, my_u, src, 16);
        uint4 c[NC8];
        uint4 d[HOT ? NC8 : 1];            // HOT: the second half of a hot token's row (pieces NC8 .. 2 NC8 - 1)
        {
            const uint4 *row = reinterpret_cast<const uint4 *>(p.n_wk16 + (size_t)w * KPAD);
#pragma unroll
            for (int j = 0; j < NC8; ++j) c[j] = __ldcg(row + j * 16 + l16);
            if constexpr (HOT) {
                if (w >= p.V) {
#pragma unroll
                    for (int j = 0; j < NC8; ++j) d[j] = __ldcg(row + (NC8 + j) * 16 + l16);
                }
            }
        }
        int c_own = (HOT && w >= p.V) ? __ldcg(reinterpret_cast<const int32_t *>(p.n_wk16 + (size_t)w * KPAD) + r_pos(kold))
                               : (int)__ldcg(p.n_wk16 + (size_t)w * KPAD + kold);

        for (int i = 0; i < nb_both; ++i) {
            const bool on = i < nb;
            const bool hot = HOT && w >= p.V;  // uniform over the half; the two halves of the warp may differ

            // ---- delitem!, arithmetically: r[kold] with n_dk and n_k one lower; nothing is stored ----
            const int jo = kold >> 7;
            const int cn = cnt_s[kold];
            const float r_old = r_s[r_pos(kold)];
            const float inv = fast_div(r_old, (float)cn + alpha);            // 1/(n_k + V beta)
            const float inv_rm = fast_div(inv, 1.f - inv);                   // 1/(n_k - 1 + V beta)
            const float r_rm = ((float)(cn - 1) + alpha) * inv_rm;

            // ---- pass 1: chunk sums of n_wk * r (beta * sum r is added per chunk) ----
            float acc[NC8];
            if (!hot) {
#pragma unroll
                for (int j = 0; j < NC8; ++j) acc[j] = dot8(c[j], r4[j * 32 + l16], r4[j * 32 + 16 + l16]);
            } else {
                // the whole row is in registers: piece jj of lane l holds topics (jj >> 1) * 128 + l * 8 + (jj & 1) * 4 .. + 3;
                // c[] = pieces 0 .. NC8 - 1, d[] = pieces NC8 .. 2 NC8 - 1 (only instantiated with HOT)
                if constexpr (HOT) hot_chunk_sums<NC8>(c, d, r4, l16, acc);
            }
            float T = transpose_reduce_half<NC8>(acc, l16);                    // lane l: chunk (l >> GSHIFT)
            const int myj = l16 >> GSHIFT;
            T = fmaf(beta, R_s[myj], T);
            if (myj == jo) {
                const float co = (float)c_own;
                T -= (co + beta) * r_old - (fmaxf(co - 1.f, 0.f) + beta) * r_rm;
            }
            T = fmaxf(T, 0.f);
            float P = T;                                                       // inclusive scan in chunk order
#pragma unroll
            for (int d = 1; d < NC8; d <<= 1) {
                const float y = __shfl_up_sync(kFull, P, d << GSHIFT, 16);
                if (myj >= d) P += y;
            }
            const float total = __shfl_sync(kFull, P, 15, 16);
            const float target = u * total;
            unsigned bal = (__ballot_sync(kFull, P >= target) >> (half * 16)) & 0xffffu;
            const int src_lane = bal ? (__ffs(bal) - 1) : 15;
            const int jsel = src_lane >> GSHIFT;
            const float base = __shfl_sync(kFull, P - T, src_lane, 16);
            uint4 cs = pick_chunk16<NC8>(c, jsel), cs2 = make_uint4(0, 0, 0, 0);
            if constexpr (HOT) {
                if (hot) pick_pair<NC8>(c, d, jsel, cs, cs2);
            }

            // ---- the next token's operands and row: in flight during pass 2 and the update ----
            int inext = (i + 1) & 15;       // lane that holds the token visited next
            if constexpr (HOT) {
                const unsigned bn = (__ballot_sync(kFull, pos == inext) >> (half * 16)) & 0xffffu;
                inext = __ffs(bn) - 1;
            }
            const int w_n = __shfl_sync(kFull, my_w, inext, 16);
            const int kold_n = __shfl_sync(kFull, my_z, inext, 16);
            const float u_n = __shfl_sync(kFull, my_u, inext, 16);
            int c_own_n = 0;
            if (i + 1 < nb_both) {
                const uint16_t *rown = p.n_wk16 + (size_t)w_n * KPAD;
#pragma unroll
                for (int j = 0; j < NC8; ++j) c[j] = __ldcg(reinterpret_cast<const uint4 *>(rown) + j * 16 + l16);
                if constexpr (HOT) {
                    if (w_n >= p.V) {
#pragma unroll
                        for (int j = 0; j < NC8; ++j) d[j] = __ldcg(reinterpret_cast<const uint4 *>(rown) + (NC8 + j) * 16 + l16);
                    }
                }
                c_own_n = (HOT && w_n >= p.V) ? __ldcg(reinterpret_cast<const int32_t *>(rown) + r_pos(kold_n)) : (int)__ldcg(rown + kold_n);
            }

            // ---- pass 2: CDF inside the selected chunk (8 topics per lane) ----
            const float4 qa = r4[jsel * 32 + l16], qb = r4[jsel * 32 + 16 + l16];
            const int kb = jsel * 128 + l16 * 8;
            float n[8];
            if (!hot) {
                // (eight values only: the conversion unit is idle here, and the packed form measured 0.3 ms slower)
                n[0] = (float)(cs.x & 0xffffu); n[1] = (float)(cs.x >> 16); n[2] = (float)(cs.y & 0xffffu); n[3] = (float)(cs.y >> 16);
                n[4] = (float)(cs.z & 0xffffu); n[5] = (float)(cs.z >> 16); n[6] = (float)(cs.w & 0xffffu); n[7] = (float)(cs.w >> 16);
            } else {
                n[0] = (float)(int)cs.x; n[1] = (float)(int)cs.y; n[2] = (float)(int)cs.z; n[3] = (float)(int)cs.w;
                n[4] = (float)(int)cs2.x; n[5] = (float)(int)cs2.y; n[6] = (float)(int)cs2.z; n[7] = (float)(int)cs2.w;
            }
            float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            // own token excluded (the clamp only matters for a count caught mid-update).  The lane-dependent jump this
            // compiles to costs 8 % of the kernel's stall samples, but the branch-free form (r[kold] rewritten in shared
            // memory before this pass, one taken off the packed cell) runs the whole sweep 20 % slower (45.4 against 37.6 ms):
            // the schedule ptxas finds for this loop is fragile (see also profiles/r02_hot_route_variants.md).
            const int eo = kold - kb;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (eo == e) { n[e] = fmaxf(n[e] - 1.f, 0.f); q[e] = r_rm; }
            float wgt[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) wgt[e] = (n[e] + beta) * q[e];
            const float s = ((wgt[0] + wgt[1]) + (wgt[2] + wgt[3])) + ((wgt[4] + wgt[5]) + (wgt[6] + wgt[7]));
            const float incl = base + half_incl_scan(s, l16);
            bal = (__ballot_sync(kFull, incl >= target) >> (half * 16)) & 0xffffu;
            const int lsel = bal ? (__ffs(bal) - 1) : 15;
            // walk the (at most) eight entries of the selected lane: first e whose running sum reaches the target
            float cw = incl - s;
            int csel = 7;
            float nsel = n[7];
            float pre[7];
#pragma unroll
            for (int e = 0; e < 7; ++e) { cw += wgt[e]; pre[e] = cw; }
#pragma unroll
            for (int e = 6; e >= 0; --e)
                if (!(pre[e] < target)) { csel = e; nsel = n[e]; }
            int knew = __shfl_sync(kFull, kb + csel, lsel, 16);
            nsel = __shfl_sync(kFull, nsel, lsel, 16);
            if (knew >= p.K) {           // rounding pushed the target past the last real topic
                knew = p.K - 1;
                nsel = -1.f;             // diagnostic recomputed below from memory
            }
            if (on && l16 == (HOT ? src : i)) my_new = knew;

            const bool moved = on && !p.frozen && knew != kold;
            if (!p.frozen) {
                // ---- additem!: the document's row changes only when the topic does ----
                const int cn2 = cnt_s[knew];
                const float r2 = r_s[r_pos(knew)];
                const float inv2 = fast_div(r2, (float)cn2 + alpha);
                float inv_new = fast_div(inv2, 1.f + inv2);                      // 1/(n_k + 1 + V beta)
                const float r_new = ((float)(cn2 + 1) + alpha) * inv_new;
                if (knew == kold) inv_new = inv;
                __syncwarp();
                if (moved) {
                    // the document's row: lane 0; the global counts: ONE fire-and-forget RED instruction with six active
                    // lanes (a single lane issuing six atomics back to back sits on the token's critical path)
                    if (l16 == 0) {
                        cnt_s[kold] = (cnt_t)(cn - 1);
                        r_s[r_pos(kold)] = r_rm;
                        R_s[jo] += r_rm - r_old;
                        cnt_s[knew] = (cnt_t)(cn2 + 1);
                        r_s[r_pos(knew)] = r_new;
                        R_s[knew >> 7] += r_new - r2;
                    }
                    if constexpr (HOT) {
                        // lanes 6..15: row[kold]--, row[knew]++ on every copy of a hot row (a 16-bit row has one), n_k[kold]--, n_k[knew]++
                        const int which = l16 - 6, kk = (which & 1) ? knew : kold, copy = which >> 1;
                        if (which >= 8 || (which >= 0 && (copy == 0 || (hot && copy < p.hot_rep)))) {
                            unsigned *ptr;
                            unsigned val = 1u;
                            if (which < 8) {
                                const int w_c = hot ? w + 2 * (copy - (int)(blockIdx.x % (unsigned)p.hot_rep)) * p.hot_cap : w;
                                unsigned *rowu = reinterpret_cast<unsigned *>(p.n_wk16 + (size_t)w_c * KPAD);
                                ptr = hot ? rowu + r_pos(kk) : rowu + (kk >> 1);
                                if (!hot) val = 1u << ((kk & 1) * 16);
                            } else {
                                ptr = reinterpret_cast<unsigned *>(my_nk) + kk;
                            }
                            red_add(ptr, (which & 1) ? val : 0u - val);
                        }
                    } else if (l16 >= 8 && l16 < 12) {
                        // lanes 8..11: row[kold]--, row[knew]++, n_k[kold]--, n_k[knew]++ (a 16-bit cell is half of a 32-bit word)
                        const int which = l16 - 8, kk = (which & 1) ? knew : kold;
                        unsigned *ptr;
                        unsigned val;
                        if (which < 2) {
                            unsigned *rowu = reinterpret_cast<unsigned *>(p.n_wk16 + (size_t)w * KPAD);
                            ptr = rowu + (kk >> 1);
                            val = 1u << ((kk & 1) * 16);
                        } else {
                            ptr = reinterpret_cast<unsigned *>(my_nk) + kk;
                            val = 1u;
                        }
                        red_add(ptr, (which & 1) ? val : 0u - val);
                    }
                    if (l16 == 0) moved_acc += 1;
                }
                if (l16 == 0 && on) {
                    if (nsel < 0.f)
                        nsel = (hot ? (float)__ldcg(reinterpret_cast<const int32_t *>(p.n_wk16 + (size_t)w * KPAD) + r_pos(knew))
                                    : (float)__ldcg(p.n_wk16 + (size_t)w * KPAD + knew)) - (knew == kold ? 1.f : 0.f);
                    diag_acc += (nsel + 1.f + beta) * inv_new;
                }
                // the next token's row was read before this update: when it is the same word, read it again
                if (__any_sync(kFull, moved && w_n == w && i + 1 < nb)) {
                    if (l16 >= (HOT ? 6 : 8) && l16 < (HOT ? 16 : 12)) __threadfence();
                    __syncwarp();
                    if (moved && w_n == w && i + 1 < nb) {
                        const uint16_t *rown = p.n_wk16 + (size_t)w_n * KPAD;
#pragma unroll
                        for (int j = 0; j < NC8; ++j) c[j] = __ldcg(reinterpret_cast<const uint4 *>(rown) + j * 16 + l16);
                        if constexpr (HOT) {
                            if (hot) {
#pragma unroll
                                for (int j = 0; j < NC8; ++j) d[j] = __ldcg(reinterpret_cast<const uint4 *>(rown) + (NC8 + j) * 16 + l16);
                            }
                        }
                        c_own_n = hot ? __ldcg(reinterpret_cast<const int32_t *>(rown) + r_pos(kold_n)) : (int)__ldcg(rown + kold_n);
                    }
                }
                __syncwarp();
            }
            src = inext;
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
