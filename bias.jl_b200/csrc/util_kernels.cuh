// util_kernels.cuh — initialisation pass (src/LDA.jl:91-98, src/BMM.jl:70-75, src/HDP.jl:217-224:
// additem! for every item) and the count-delta kernels of the multi-GPU exchange.
#pragma once
#include "common.cuh"

namespace bias {

#ifdef __CUDACC__

// Range checks the reference gets from Julia's bounds checking (BoundsError) and @check_args:
// flag bit 0: zz outside [0, K); bit 1: word id outside [0, V); bit 2: bad Sent entry.
__global__ void validate_kernel(const int32_t *z, int64_t N, int K, const int32_t *word, int64_t V,
                                const int32_t *sent_word, const int32_t *sent_count, int64_t nnz, int *flag) {
    int bad = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const int k = z[i];
        if (k < 0 || k >= K) bad |= 1;
        if (word) {
            const int w = word[i];
            if (w < 0 || w >= V) bad |= 2;
        }
    }
    if (sent_word)
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += stride)
            if (sent_word[t] < 0 || sent_word[t] >= V || sent_count[t] < 0) bad |= 4;
    if (bad) atomicOr(flag, bad);
}

// 16-bit host formats of the streaming path: word ids / assignments travel as uint16 and are widened (narrowed) here
__global__ void __launch_bounds__(256, 6) widen_u16_kernel(const uint16_t *src, int64_t n, int32_t *dst) {
    const int64_t n8 = n / 8, stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t q = t0; q < n8; q += stride) {
        const uint4 x = reinterpret_cast<const uint4 *>(src)[q];
        reinterpret_cast<int4 *>(dst)[2 * q] = make_int4((int)(x.x & 0xffffu), (int)(x.x >> 16), (int)(x.y & 0xffffu), (int)(x.y >> 16));
        reinterpret_cast<int4 *>(dst)[2 * q + 1] = make_int4((int)(x.z & 0xffffu), (int)(x.z >> 16), (int)(x.w & 0xffffu), (int)(x.w >> 16));
    }
    for (int64_t i = n8 * 8 + t0; i < n; i += stride) dst[i] = (int32_t)src[i];
}
__global__ void __launch_bounds__(256, 6) narrow_u16_kernel(const int32_t *src, int64_t n, uint16_t *dst) {
    const int64_t n8 = n / 8, stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t q = t0; q < n8; q += stride) {
        const int4 a = reinterpret_cast<const int4 *>(src)[2 * q], b = reinterpret_cast<const int4 *>(src)[2 * q + 1];
        uint4 o;
        o.x = (unsigned)(a.x & 0xffff) | ((unsigned)a.y << 16);
        o.y = (unsigned)(a.z & 0xffff) | ((unsigned)a.w << 16);
        o.z = (unsigned)(b.x & 0xffff) | ((unsigned)b.y << 16);
        o.w = (unsigned)(b.z & 0xffff) | ((unsigned)b.w << 16);
        reinterpret_cast<uint4 *>(dst)[q] = o;
    }
    for (int64_t i = n8 * 8 + t0; i < n; i += stride) dst[i] = (uint16_t)src[i];
}

// item -> group id by binary search in the CSR offsets
__global__ void group_of_kernel(const int64_t *ptr, int64_t G, int64_t N, int32_t *group_of) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = G;        // invariant: ptr[lo] <= i < ptr[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (ptr[mid] <= i) lo = mid; else hi = mid;
        }
        group_of[i] = (int32_t)lo;
    }
}

// additem! for word-id items: n_wk[w][z] += 1, n_k[z] += 1 (block-private histogram for n_k)
__global__ void build_md_int_kernel(const int32_t *word, const int32_t *z, int64_t N, int Kpad,
                                    int32_t *n_wk, int32_t *n_k, int32_t *n_items) {
    extern __shared__ int32_t hist[];
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = z[i];
        atomicAdd(n_wk + (size_t)word[i] * Kpad + k, 1);
        atomicAdd(&hist[k], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
        const int h = hist[k];
        if (h) {
            atomicAdd(n_k + k, h);
            atomicAdd(n_items + k, h);
        }
    }
}

// additem! for Sent items: one warp per item
__global__ void build_md_sent_kernel(const int64_t *sptr, const int32_t *sw, const int32_t *sc, const int32_t *z,
                                     int64_t N, int Kpad, int32_t *n_wk, int32_t *n_k, int32_t *n_items) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += nw) {
        const int k = z[i];
        int m = 0;
        for (int64_t t = sptr[i] + lane; t < sptr[i + 1]; t += 32) {
            atomicAdd(n_wk + (size_t)sw[t] * Kpad + k, sc[t]);
            m += sc[t];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(kFull, m, o);
        if (lane == 0) {
            atomicAdd(n_k + k, m);
            atomicAdd(n_items + k, 1);
        }
    }
}

// additem! for Float64 items under Gaussian1DGaussian1D: (sum x, n) per component
__global__ void build_g1g1_kernel(const double *x, const int32_t *z, int64_t N, int Kpad,
                                  double *sumx, int32_t *n_items) {
    extern __shared__ unsigned char sraw[];
    double *hs = reinterpret_cast<double *>(sraw);
    int32_t *hn = reinterpret_cast<int32_t *>(hs + Kpad);
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x) { hs[k] = 0.0; hn[k] = 0; }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = z[i];
        atomicAdd(&hs[k], x[i]);
        atomicAdd(&hn[k], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kpad; k += blockDim.x)
        if (hn[k]) { atomicAdd(sumx + k, hs[k]); atomicAdd(n_items + k, hn[k]); }
}

// nn[j, k] (the reference's group-by-component matrix)
__global__ void build_group_rows_kernel(const int32_t *z, const int32_t *group_of, int64_t N, int Kpad,
                                        int32_t *rows) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = z[i];
        if (k >= 0) atomicAdd(rows + (size_t)group_of[i] * Kpad + k, 1);
    }
}

// group rows of a list of query items (verification of models that keep no n_jk table):
// one warp per query, row q of `rows` <- histogram of z over the query item's group.
__global__ void build_query_rows_kernel(const int64_t *item_idx, int64_t nq, const int64_t *ptr, int64_t G,
                                        const int32_t *z, int Kpad, int32_t *rows, int32_t *row_of_query) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += nw) {
        const int64_t i = item_idx[q];
        int64_t lo = 0, hi = G;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (ptr[mid] <= i) lo = mid; else hi = mid;
        }
        for (int64_t t = ptr[lo] + lane; t < ptr[lo + 1]; t += 32) atomicAdd(rows + (size_t)q * Kpad + z[t], 1);
        if (lane == 0) row_of_query[q] = (int32_t)q;
    }
}

// ---- multi-GPU exchange -----------------------------------------------------------------
__global__ void delta_pack_i32_kernel(const int32_t *live, const int32_t *base, int32_t *delta, int64_t n4) {
    // n4 = number of int4 groups (buffers are padded to 16 bytes)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 a = reinterpret_cast<const int4 *>(live)[i], b = reinterpret_cast<const int4 *>(base)[i];
        reinterpret_cast<int4 *>(delta)[i] = make_int4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    }
}
__global__ void delta_apply_i32_kernel(int32_t *live, int32_t *base, const int32_t *delta, int64_t n4,
                                       const int32_t *skip_if_set = nullptr) {
    if (skip_if_set && *skip_if_set != 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 b = reinterpret_cast<const int4 *>(base)[i], d = reinterpret_cast<const int4 *>(delta)[i];
        const int4 v = make_int4(b.x + d.x, b.y + d.y, b.z + d.z, b.w + d.w);
        reinterpret_cast<int4 *>(live)[i] = v;
        reinterpret_cast<int4 *>(base)[i] = v;
    }
}
// 16-bit exchange of the word-topic deltas: 8 cells per thread; *flag is raised when a delta could overflow int16 once
// `world` of them are summed (the caller then repeats the exchange in 32 bits).  NCCL has no int16 sum, so two deltas
// travel in one int32 as d1 * 65536 + d0 (arithmetic, not bitwise): the int32 sum over ranks is D1 * 65536 + D0, and
// with |D0| < 2^15 both come back exactly (D0 = sign-extended low half, D1 = (X - D0) >> 16).
__global__ void delta_pack16_kernel(const int32_t *live, const int32_t *base, int16_t *d16, int64_t n8, int limit, int32_t *flag) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 a0 = reinterpret_cast<const int4 *>(live)[2 * i], a1 = reinterpret_cast<const int4 *>(live)[2 * i + 1];
        const int4 b0 = reinterpret_cast<const int4 *>(base)[2 * i], b1 = reinterpret_cast<const int4 *>(base)[2 * i + 1];
        const int d[8] = {a0.x - b0.x, a0.y - b0.y, a0.z - b0.z, a0.w - b0.w, a1.x - b1.x, a1.y - b1.y, a1.z - b1.z, a1.w - b1.w};
        unsigned o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            bad |= (abs(d[2 * e]) > limit) | (abs(d[2 * e + 1]) > limit);
            o[e] = (unsigned)(d[2 * e + 1] * 65536 + d[2 * e]);
        }
        reinterpret_cast<uint4 *>(d16)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
__global__ void delta_apply16_kernel(int32_t *live, int32_t *base, const int16_t *d16, int64_t n8, const int32_t *flag) {
    if (*flag != 0) return;         // some rank's delta did not fit: nothing is applied, the deltas stay pending
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 x = reinterpret_cast<const uint4 *>(d16)[i];
        const int4 b0 = reinterpret_cast<const int4 *>(base)[2 * i], b1 = reinterpret_cast<const int4 *>(base)[2 * i + 1];
        const int xs[4] = {(int)x.x, (int)x.y, (int)x.z, (int)x.w};
        int lo[4], hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            lo[e] = (int)(int16_t)(xs[e] & 0xffff);
            hi[e] = (xs[e] - lo[e]) >> 16;
        }
        const int4 v0 = make_int4(b0.x + lo[0], b0.y + hi[0], b0.z + lo[1], b0.w + hi[1]);
        const int4 v1 = make_int4(b1.x + lo[2], b1.y + hi[2], b1.z + lo[3], b1.w + hi[3]);
        reinterpret_cast<int4 *>(live)[2 * i] = v0;
        reinterpret_cast<int4 *>(live)[2 * i + 1] = v1;
        reinterpret_cast<int4 *>(base)[2 * i] = v0;
        reinterpret_cast<int4 *>(base)[2 * i + 1] = v1;
    }
}
__global__ void delta_pack_f64_kernel(const double *live, const double *base, double *delta, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        delta[i] = live[i] - base[i];
}
__global__ void delta_apply_f64_kernel(double *live, double *base, const double *delta, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = base[i] + delta[i];
        live[i] = v;
        base[i] = v;
    }
}

// topic-major int64 export for posterior(): mi[k*V + w] = n_wk[w*Kpad + k]
__global__ void export_mi_kernel(const int32_t *n_wk, int64_t V, int K, int Kpad, int64_t *mi) {
    const int64_t total = V * (int64_t)K;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = idx / K;
        const int k = (int)(idx - w * K);
        mi[(int64_t)k * V + w] = (int64_t)n_wk[w * Kpad + k];
    }
}

#endif

// n_k += sum of the replicas that collected this sweep's changes (lda_fast.cuh: kNkRep); the replicas are cleared
__global__ void fold_nk_kernel(int32_t *n_k, int32_t *nk_rep, int Kpad, int n_rep) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Kpad; k += gridDim.x * blockDim.x) {
        int s = 0;
        for (int r = 0; r < n_rep; ++r) {
            s += nk_rep[r * Kpad + k];
            nk_rep[r * Kpad + k] = 0;
        }
        if (s) n_k[k] += s;
    }
}

}  // namespace bias
