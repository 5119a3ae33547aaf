// comm.cu — the sharded (multi-GPU) LDA sweep behind the C ABI: one count exchange per sweep, done by our own kernels
// over peer memory (NVLink / NVSwitch), no library collective.
//
// AD-LDA sharding (the reference is single-process): documents are split over the ranks, every rank keeps a
// full replica of the word-topic counts in its narrow tables (lda_half16.cuh) and samples its shard against it.  After
// a sweep the replicas differ by each rank's moves.  The exchange is a reduce-scatter + all-gather fused with the
// delta arithmetic, directly on the live tables:
//
//   rank r OWNS the rows [V r / n, V (r+1) / n) (and a share of the hot rows).  For its rows it keeps `base`, the counts
//   all ranks agreed on after the previous exchange, reads the live row of EVERY rank through peer memory, forms
//   new = base + sum_p (live_p - base), stores it in base and writes it into every rank's live table.
//
// Per rank and sweep that moves (n-1)/n of the table in and out over NVLink and touches local HBM once for reading and
// once for writing — against pack / all-reduce / apply passes over int32 copies of the whole table (round 1: 1.8 GB of
// local traffic per sweep at any n).  n_k travels as published deltas that every rank sums (Kpad ints).
//
// Ordering between ranks is two barriers per exchange.  One process driving several GPUs (bias_group, what the Julia
// shim calls through bias_lda_run_multi) uses CUDA events across its streams; one process per GPU (torchrun) maps the
// other ranks' tables with CUDA IPC and uses arrival counters in peer memory (peer_barrier_kernel, bounded spin).
#include <algorithm>
#include <cstring>
#include <new>
#include <chrono>
#include <string>
#include <thread>
#include <vector>
#include "model.hpp"
#include "narrow.cuh"

namespace bias {

// ---- the narrow tables -------------------------------------------------------------------------------------------
// One allocation [V + 2 * hot_cap][Kpad] of uint16: rows 0 .. V-1 are the 16-bit rows of the words, and behind them
// every HOT word h owns two rows (= Kpad int32 cells) at row V + 2h.  A word is hot when its frequency in the whole
// corpus (all ranks) exceeds 65,535, i.e. when one of its cells could leave 16 bits; all other words can never overflow
// (a cell is at most the word's frequency), so the classification is static between (re)loads and nothing has to be
// checked per sweep.  The int32 cells of a hot row are stored in r_pos order, which makes the sweep kernel's 16-byte
// loads of a hot row the same address pattern as those of a 16-bit row (lane l, load jj: row + (jj*16 + l) * 16 bytes)
// and lines the cells up with the lane's r values in shared memory.

// word_hot[w] = index of the hot row of w, or -1.  Deterministic (prefix sum in word order): every rank of a sharded
// run derives the same indices from the same global frequencies.  One block.
__global__ void classify_words_kernel(const int32_t *freq_g, int64_t V, int32_t hot_cap, int32_t *word_hot, int32_t *hot_word,
                                      int32_t *n_hot_out) {
    __shared__ int32_t part[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int64_t per = (V + nt - 1) / nt, lo = min(V, (int64_t)t * per), hi = min(V, lo + per);
    int c = 0;
    for (int64_t w = lo; w < hi; ++w) c += freq_g[w] > 65535;
    part[t] = c;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int i = 0; i < nt; ++i) { const int x = part[i]; part[i] = run; run += x; }
        n_hot_out[0] = run;                      // words that need 32-bit cells
        n_hot_out[1] = run > hot_cap ? 1 : 0;    // more than the table has room for: the int32-table kernel takes the sweeps
    }
    __syncthreads();
    int h = part[t];
    for (int64_t w = lo; w < hi; ++w) {
        if (freq_g[w] > 65535 && h < hot_cap) {
            word_hot[w] = h;
            hot_word[h] = (int32_t)w;
            ++h;
        } else {
            word_hot[w] = -1;
        }
    }
}

// int32 table -> narrow tables (one warp per word)
__global__ void narrow_build_kernel(const int32_t *n_wk, const int32_t *word_hot, int64_t V, int Kpad, uint16_t *tab) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < V; w += nw) {
        const int h = word_hot[w];
        const int4 *src = reinterpret_cast<const int4 *>(n_wk + (size_t)w * Kpad);
        uint4 *dst = reinterpret_cast<uint4 *>(tab + (size_t)w * Kpad);
        if (h < 0) {
            for (int q = lane; q < Kpad / 8; q += 32) {
                const int4 a = __ldcg(src + 2 * q), b = __ldcg(src + 2 * q + 1);
                uint4 o;
                o.x = (unsigned)(a.x & 0xffff) | ((unsigned)a.y << 16);
                o.y = (unsigned)(a.z & 0xffff) | ((unsigned)a.w << 16);
                o.z = (unsigned)(b.x & 0xffff) | ((unsigned)b.y << 16);
                o.w = (unsigned)(b.z & 0xffff) | ((unsigned)b.w << 16);
                dst[q] = o;
            }
        } else {
            for (int q = lane; q < Kpad / 8; q += 32) dst[q] = make_uint4(0, 0, 0, 0);      // unused while the word is hot
            int4 *hot = reinterpret_cast<int4 *>(tab + (size_t)(V + 2 * (int64_t)h) * Kpad);
            // topics 4q .. 4q+3 sit at r_pos(4q) .. +3 (r_pos keeps groups of four together)
            for (int q = lane; q < Kpad / 4; q += 32) hot[r_pos(4 * q) >> 2] = __ldcg(src + q);
        }
    }
}

// narrow tables -> int32 table
__global__ void narrow_fold_kernel(const uint16_t *tab, const int32_t *word_hot, int64_t V, int Kpad, int32_t *n_wk) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < V; w += nw) {
        const int h = word_hot[w];
        int4 *dst = reinterpret_cast<int4 *>(n_wk + (size_t)w * Kpad);
        if (h < 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(tab + (size_t)w * Kpad);
            for (int q = lane; q < Kpad / 8; q += 32) {
                const uint4 x = __ldcg(src + q);
                dst[2 * q] = make_int4((int)(x.x & 0xffffu), (int)(x.x >> 16), (int)(x.y & 0xffffu), (int)(x.y >> 16));
                dst[2 * q + 1] = make_int4((int)(x.z & 0xffffu), (int)(x.z >> 16), (int)(x.w & 0xffffu), (int)(x.w >> 16));
            }
        } else {
            const int4 *hot = reinterpret_cast<const int4 *>(tab + (size_t)(V + 2 * (int64_t)h) * Kpad);
            for (int q = lane; q < Kpad / 4; q += 32) dst[q] = __ldcg(hot + (r_pos(4 * q) >> 2));
        }
    }
}

// order-independent digest of an int32 buffer (replica comparison)
__global__ void digest_kernel(const int32_t *x, int64_t n, unsigned long long *out) {
    unsigned long long h = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long v = ((unsigned long long)i + 1) * 0x9E3779B97F4A7C15ull ^ (unsigned long long)(unsigned)x[i] * 0xC2B2AE3D27D4EB4Full;
        v ^= v >> 29;
        v *= 0xBF58476D1CE4E5B9ull;
        v ^= v >> 32;
        if (x[i] != 0) h += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(kFull, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}

struct XchgParams {
    uint16_t *tab[kMaxPeers];          // every rank's narrow tables (own included), addressable from this device
    const int32_t *nk_pub[kMaxPeers];  // every rank's published n_k (delta, or the local counts when merging)
    int32_t world, rank, merge;        // merge: the tables hold local counts only (after a (re)load): new = sum_p live_p
    int64_t V;
    int32_t Kpad;
    const int32_t *hot_info;
    uint16_t *base16;                  // [rows owned][Kpad]
    int32_t *base_hot;                 // [hot_cap][Kpad]; only the owned rows are used
    int32_t *n_k, *nk_base;
};

struct FreqParams {
    const int32_t *freq_pub[kMaxPeers];
    int32_t world;
    int64_t V;
    int32_t *freq_g;
};

struct BarrierParams {
    uint32_t *flags[kMaxPeers];        // every rank's arrival counters [kMaxPeers]
    int32_t world, rank;
    uint32_t epoch;
    int32_t *err;
};

// Loads and stores on the mapped peer pointers.  Every address is read once and written once between the two barriers of
// an exchange (release / acquire at system scope on the arrival counters), so plain L2-only accesses are enough.  Measured:
// 600 - 640 GB/s per direction and GPU on 2 to 8 GPUs (a rank's own reads and its peers' writes arrive together; bench.py
// `exchange`), about 0.8 of the 770 GB/s a peer copy reaches on this pool.
__device__ __forceinline__ uint4 ld_sys(const uint4 *p) { return __ldcg(p); }
__device__ __forceinline__ void st_sys(uint4 *p, const uint4 v) { __stcg(p, v); }

// new = base + sum over ranks of (x_r - base), independently in the two 16-bit halves of a word (mod 2^16: exact, since
// the true count of a cell of a 16-bit row never exceeds the word's corpus frequency <= 65,535)
__device__ __forceinline__ unsigned merge16(const unsigned (&x)[kMaxPeers], int world, unsigned b) {
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
        if (r < world) { lo += x[r] & 0xffffu; hi += x[r] >> 16; }
    const unsigned bl = b & 0xffffu, bh = b >> 16;
    lo = lo + bl - (unsigned)world * bl;
    hi = hi + bh - (unsigned)world * bh;
    return (lo & 0xffffu) | (hi << 16);
}
__device__ __forceinline__ unsigned merge32(const unsigned (&x)[kMaxPeers], int world, unsigned b) {
    unsigned s = 0;
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
        if (r < world) s += x[r];
    return s + b - (unsigned)world * b;
}

// U pieces at once for a run of W ranks (W = capacity of the unrolled loops; p.world <= W): all loads first, W * U of
// them in flight per thread
template <bool HOT, int W, int U>
__device__ __forceinline__ void exchange_pieces(const XchgParams &p, const int64_t (&q)[U], int n, uint4 *base_minus_q0) {
    uint4 x[U][W], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (u < n) {
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < p.world) x[u][r] = ld_sys(reinterpret_cast<const uint4 *>(p.tab[r]) + q[u]);
            b[u] = p.merge ? make_uint4(0, 0, 0, 0) : base_minus_q0[q[u]];
        }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (u < n) {
            unsigned xs[kMaxPeers];
            uint4 y;
#define BIAS_LANE(f)                                                                                  \
    {                                                                                                 \
        _Pragma("unroll") for (int r = 0; r < kMaxPeers; ++r) xs[r] = (r < W && r < p.world) ? x[u][r < W ? r : 0].f : 0u; \
        y.f = HOT ? merge32(xs, p.world, b[u].f) : merge16(xs, p.world, b[u].f);                      \
    }
            BIAS_LANE(x) BIAS_LANE(y) BIAS_LANE(z) BIAS_LANE(w)
#undef BIAS_LANE
            base_minus_q0[q[u]] = y;
#pragma unroll
            for (int r = 0; r < W; ++r)
                if (r < p.world) st_sys(reinterpret_cast<uint4 *>(p.tab[r]) + q[u], y);
        }
}

template <bool HOT, int W, int U>
__device__ __forceinline__ void exchange_range(const XchgParams &p, int64_t q0, int64_t q1, int64_t off, uint4 *base, int64_t t0, int64_t stride) {
    // pieces off + [q0, q1) of the tables, base[0 .. q1 - q0)
    for (int64_t s = q0 + t0; s < q1; s += (int64_t)U * stride) {
        int64_t q[U];
        int n = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            q[u] = off + s + (int64_t)u * stride;
            if (s + (int64_t)u * stride < q1) n = u + 1;
        }
        exchange_pieces<HOT, W, U>(p, q, n, base - (off + q0));
    }
}

template <int W, int U>
__global__ void __launch_bounds__(256) narrow_exchange_kernel(const XchgParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // the 16-bit rows this rank owns, in 16-byte pieces
    const int64_t row0 = p.V * p.rank / p.world, row1 = p.V * (p.rank + 1) / p.world;
    exchange_range<false, W, U>(p, row0 * (p.Kpad / 8), row1 * (p.Kpad / 8), 0, reinterpret_cast<uint4 *>(p.base16), t0, stride);
    // its share of the hot rows (int32 cells; row h starts V*Kpad/8 + h*Kpad/4 pieces into the table)
    const int64_t n_hot = p.hot_info[0];
    const int64_t h0 = n_hot * p.rank / p.world, h1 = n_hot * (p.rank + 1) / p.world;
    exchange_range<true, W, U>(p, h0 * (p.Kpad / 4), h1 * (p.Kpad / 4), p.V * (p.Kpad / 8),
                               reinterpret_cast<uint4 *>(p.base_hot) + h0 * (p.Kpad / 4), t0, stride);
    // n_k: every rank sums the published deltas of all ranks
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < p.Kpad; k += blockDim.x) {
            int s = p.merge ? 0 : p.nk_base[k];
            for (int r = 0; r < p.world; ++r) s += p.nk_pub[r][k];
            p.n_k[k] = s;
            p.nk_base[k] = s;
        }
}

__global__ void nk_publish_kernel(const int32_t *n_k, const int32_t *nk_base, int32_t *nk_pub, int Kpad, int merge) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Kpad; k += gridDim.x * blockDim.x)
        nk_pub[k] = merge ? n_k[k] : n_k[k] - nk_base[k];
}

__global__ void freq_sum_kernel(const FreqParams p) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < p.V; w += (int64_t)gridDim.x * blockDim.x) {
        long long s = 0;
        for (int r = 0; r < p.world; ++r) s += p.freq_pub[r][w];
        p.freq_g[w] = (int32_t)min(s, (long long)INT32_MAX);
    }
}

// One process per GPU: thread t tells rank t "rank `rank` has arrived at barrier `epoch`" and waits for rank t's arrival
// in its own counters.  Everything this stream did before is complete when the kernel starts (stream order), so peers
// that pass the barrier see it; the spin is bounded (a dead peer must not hang the GPU): after 60 s *err is raised.
__global__ void peer_barrier_kernel(const BarrierParams p) {
    const int t = threadIdx.x;
    if (t >= p.world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[t] + p.rank), "r"(p.epoch) : "memory");
    unsigned long long start, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags[p.rank] + t) : "memory");
        if ((int)(v - p.epoch) >= 0) break;
        __nanosleep(100);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - start > 60ull * 1000000000ull) {
            atomicExch(p.err, 1);
            break;
        }
    }
}

// ---------------------------------------------------------------------------------------
// the shared block: narrow tables + what the other ranks read besides them
// ---------------------------------------------------------------------------------------
static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int32_t narrow_hot_cap(int64_t V, int64_t n_items_cap) {
    // a word is hot above 65,535 occurrences: a corpus of n items has fewer than n / 65,536 + 1 of them
    return (int32_t)std::min<int64_t>(V, n_items_cap / 65536 + 1);
}

int narrow_alloc(bias_model *m, int32_t hot_cap) {
    if (m->share_base && hot_cap == m->hot_cap) return BIAS_OK;
    if (m->share_base) {
        BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
        cudaFree(m->share_base);
        m->share_base = nullptr;
        if (m->hot_word) cudaFree(m->hot_word);
        if (m->base_hot) cudaFree(m->base_hot);
        m->hot_word = nullptr;
        m->base_hot = nullptr;
    }
    // the hot rows exist kHotRep times (lda_fast.cuh) unless that would be hundreds of MB; the first copy is the canonical one
    m->hot_rep = hot_cap <= 16384 ? kHotRep : 1;
    m->hot_rep_valid = false;
    const size_t tab = round_up(((size_t)m->V + 2 * (size_t)hot_cap * m->hot_rep) * m->Kpad * sizeof(uint16_t), 256);
    const size_t nk = round_up((size_t)m->Kpad * 4, 256), fr = round_up((size_t)std::max<int64_t>(1, m->V) * 4, 256);
    m->off_nk_pub = tab;
    m->off_freq_pub = tab + nk;
    m->off_flags = tab + nk + fr;
    m->share_bytes = m->off_flags + 256;
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->share_base, m->share_bytes));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->share_base + tab, 0, m->share_bytes - tab, m->ctx->stream));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->hot_word, (size_t)std::max(1, hot_cap) * 4));
    m->n_wk16 = reinterpret_cast<uint16_t *>(m->share_base);
    m->hot_cap = hot_cap;
    m->narrow_valid = false;
    m->class_valid = false;
    m->peer_base[0] = m->share_base;
    return BIAS_OK;
}

void narrow_free(bias_model *m) {
    for (int r = 0; r < kMaxPeers; ++r)
        if (m->peer_ipc[r] && m->peer_base[r]) cudaIpcCloseMemHandle(m->peer_base[r]);
    if (m->ev_x0) cudaEventDestroy(m->ev_x0);
    if (m->ev_x1) cudaEventDestroy(m->ev_x1);
    void *ptrs[] = {m->share_base, m->hot_word, m->base16, m->base_hot, m->nk_base, m->comm_err};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    m->share_base = nullptr;
    m->n_wk16 = nullptr;
}

static int narrow_grid(bias_model *m) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((m->V * 32 + 255) / 256, (int64_t)m->ctx->num_sms * 8));
}

// copies 1 .. rep-1 of the hot rows <- the canonical copy (after a conversion or an exchange wrote it)
__global__ void narrow_replicate_kernel(uint16_t *tab, int64_t V, int Kpad, int hot_cap, int rep, const int32_t *hot_info) {
    const int n_hot = min(hot_info[0], hot_cap);
    const int64_t per_row = Kpad / 4, total = (int64_t)n_hot * per_row;          // 16-byte pieces of the canonical hot rows
    const uint4 *src = reinterpret_cast<const uint4 *>(tab + (size_t)V * Kpad);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcg(src + q);
        for (int c = 1; c < rep; ++c) reinterpret_cast<uint4 *>(tab + ((size_t)V + 2 * (size_t)c * hot_cap) * Kpad)[q] = v;
    }
}

int narrow_replicate(bias_model *m) {
    if (m->hot_rep_valid) return BIAS_OK;
    if (m->hot_rep > 1 && m->n_hot_host > 0) {
        const int64_t total = (int64_t)m->n_hot_host * (m->Kpad / 4);
        narrow_replicate_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)m->ctx->num_sms * 8)), 256, 0,
                                  m->ctx->stream>>>(m->n_wk16, m->V, m->Kpad, m->hot_cap, m->hot_rep, m->hot_info);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
    }
    m->hot_rep_valid = true;
    return BIAS_OK;
}

int narrow_classify(bias_model *m) {
    classify_words_kernel<<<1, 1024, 0, m->ctx->stream>>>(m->word_freq_g, m->V, m->hot_cap, m->word_hot, m->hot_word, m->hot_info);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    m->class_valid = true;
    // the host picks the kernel instantiation by whether hot words exist: one small read-back per (re)load
    int32_t h[2] = {0, 0};
    BIAS_CUDA_TRY(cudaMemcpyAsync(h, m->hot_info, sizeof h, cudaMemcpyDeviceToHost, m->ctx->stream));
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    if (h[1]) return fail(BIAS_ESTATE, "%d words need 32-bit rows but the table has room for %d", h[0], m->hot_cap);
    m->n_hot_host = h[0];
    return BIAS_OK;
}

int narrow_build(bias_model *m) {
    narrow_build_kernel<<<narrow_grid(m), 256, 0, m->ctx->stream>>>(m->n_wk, m->word_hot, m->V, m->Kpad, m->n_wk16);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    m->narrow_valid = true;
    m->hot_rep_valid = false;
    return BIAS_OK;
}

int narrow_fold(bias_model *m) {
    narrow_fold_kernel<<<narrow_grid(m), 256, 0, m->ctx->stream>>>(m->n_wk16, m->word_hot, m->V, m->Kpad, m->n_wk);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    m->wide_valid = true;
    return BIAS_OK;
}

int comm_global_word_freq(bias_model *m) {
    if (m->world > 1) return fail(BIAS_ESTATE, "sharded model: the word frequencies are summed by the (collective) sweep call");
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->word_freq_g, m->word_freq, (size_t)std::max<int64_t>(1, m->V) * 4, cudaMemcpyDeviceToDevice,
                                  m->ctx->stream));
    return BIAS_OK;
}

// ---- the steps of the protocol: local launches only; a barrier over all ranks goes between them ----
static int step_publish_freq(bias_model *m) {
    if (!m->wide_valid) return fail(BIAS_ESTATE, "sharded (re)load: the int32 counts are not current");
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->share_base + m->off_freq_pub, m->word_freq, (size_t)std::max<int64_t>(1, m->V) * 4,
                                  cudaMemcpyDeviceToDevice, m->ctx->stream));
    return BIAS_OK;
}

static int step_classify_build(bias_model *m) {
    cudaStream_t st = m->ctx->stream;
    FreqParams f{};
    for (int r = 0; r < m->world; ++r) f.freq_pub[r] = reinterpret_cast<const int32_t *>(m->peer_base[r] + m->off_freq_pub);
    f.world = m->world;
    f.V = m->V;
    f.freq_g = m->word_freq_g;
    freq_sum_kernel<<<std::max(1, (int)std::min<int64_t>((m->V + 255) / 256, 1024)), 256, 0, st>>>(f);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    BIAS_TRY(narrow_classify(m));
    BIAS_TRY(narrow_build(m));
    nk_publish_kernel<<<(m->Kpad + 255) / 256, 256, 0, st>>>(m->n_k, m->nk_base, reinterpret_cast<int32_t *>(m->share_base + m->off_nk_pub),
                                                          m->Kpad, 1);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

static int step_exchange(bias_model *m, int merge) {
    XchgParams p{};
    for (int r = 0; r < m->world; ++r) {
        p.tab[r] = reinterpret_cast<uint16_t *>(m->peer_base[r]);
        p.nk_pub[r] = reinterpret_cast<const int32_t *>(m->peer_base[r] + m->off_nk_pub);
    }
    p.world = m->world;
    p.rank = m->rank;
    p.merge = merge;
    p.V = m->V;
    p.Kpad = m->Kpad;
    p.hot_info = m->hot_info;
    p.base16 = m->base16;
    p.base_hot = m->base_hot;
    p.n_k = m->n_k;
    p.nk_base = m->nk_base;
    const int grid = m->ctx->num_sms * 8;
    if (m->world <= 2) narrow_exchange_kernel<2, 4><<<grid, 256, 0, m->ctx->stream>>>(p);
    else if (m->world <= 4) narrow_exchange_kernel<4, 2><<<grid, 256, 0, m->ctx->stream>>>(p);
    else narrow_exchange_kernel<8, 1><<<grid, 256, 0, m->ctx->stream>>>(p);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    m->wide_valid = false;
    m->hot_rep_valid = false;       // this rank and its peers write the canonical copy of the hot rows
    return BIAS_OK;
}

static int step_publish_nk(bias_model *m) {
    nk_publish_kernel<<<(m->Kpad + 255) / 256, 256, 0, m->ctx->stream>>>(m->n_k, m->nk_base,
                                                                       reinterpret_cast<int32_t *>(m->share_base + m->off_nk_pub), m->Kpad, 0);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

// ---- barriers ----
static int flag_barrier(bias_model *m) {
    BarrierParams b{};
    for (int r = 0; r < m->world; ++r) b.flags[r] = reinterpret_cast<uint32_t *>(m->peer_base[r] + m->off_flags);
    b.world = m->world;
    b.rank = m->rank;
    b.epoch = ++m->barrier_epoch;
    b.err = m->comm_err;
    peer_barrier_kernel<<<1, 32, 0, m->ctx->stream>>>(b);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

static int comm_buffers(bias_model *m, int world) {
    if (m->base16) return BIAS_OK;
    const size_t rows = (size_t)(m->V / world + 1);        // the rows a rank owns
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->base16, rows * m->Kpad * sizeof(uint16_t)));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->base_hot, (size_t)std::max(1, m->hot_cap) * m->Kpad * sizeof(int32_t)));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->nk_base, (size_t)m->Kpad * sizeof(int32_t)));
    BIAS_CUDA_TRY(cudaMalloc((void **)&m->comm_err, 16));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->nk_base, 0, (size_t)m->Kpad * sizeof(int32_t), m->ctx->stream));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->comm_err, 0, 16, m->ctx->stream));
    return BIAS_OK;
}

static int sharded_ok(const bias_model *m) {
    if (!m->fast || !m->share_base || m->Kpad > 1024 || m->max_group_len > 65535)
        return fail(BIAS_EINVAL, "the in-library sharded sweep covers LDA x MultinomialDirichlet x word ids with K <= 1024 and "
                                 "documents shorter than 65,536 tokens (other models: bias_model_delta_* + an external all-reduce)");
    return BIAS_OK;
}

// one process per GPU: the collective part of bias_model_sweep
int comm_prepare_flags(bias_model *m) {
    BIAS_TRY(sharded_ok(m));
    BIAS_TRY(step_publish_freq(m));
    BIAS_TRY(flag_barrier(m));
    BIAS_TRY(step_classify_build(m));
    BIAS_TRY(flag_barrier(m));
    BIAS_TRY(step_exchange(m, 1));
    BIAS_TRY(flag_barrier(m));
    return BIAS_OK;
}

int comm_exchange_flags(bias_model *m) {
    BIAS_TRY(step_publish_nk(m));
    BIAS_TRY(flag_barrier(m));
    if (!m->ev_x0) {
        BIAS_CUDA_TRY(cudaEventCreate(&m->ev_x0));
        BIAS_CUDA_TRY(cudaEventCreate(&m->ev_x1));
    }
    BIAS_CUDA_TRY(cudaEventRecord(m->ev_x0, m->ctx->stream));      // all ranks have arrived: what follows is transfer time
    BIAS_TRY(step_exchange(m, 0));
    BIAS_CUDA_TRY(cudaEventRecord(m->ev_x1, m->ctx->stream));
    BIAS_TRY(flag_barrier(m));
    return BIAS_OK;
}

int comm_check(bias_model *m) {
    if (m->world <= 1 || !m->comm_err) return BIAS_OK;
    int32_t e = 0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(&e, m->comm_err, 4, cudaMemcpyDeviceToHost, m->ctx->stream));
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    if (e) return fail(BIAS_ESTATE, "a rank of the sharded run did not reach a barrier within 60 s; the counts of this replica are not valid");
    return BIAS_OK;
}

}  // namespace bias

using namespace bias;

struct bias_group {
    std::vector<bias_model *> ms;
    std::vector<cudaEvent_t> ev;
};

// all streams of the group wait for everything enqueued so far on every stream of the group
static int group_barrier(bias_group *g) {
    const size_t n = g->ms.size();
    for (size_t i = 0; i < n; ++i) {
        BIAS_CUDA_TRY(cudaSetDevice(g->ms[i]->ctx->device));
        BIAS_CUDA_TRY(cudaEventRecord(g->ev[i], g->ms[i]->ctx->stream));
    }
    for (size_t i = 0; i < n; ++i) {
        BIAS_CUDA_TRY(cudaSetDevice(g->ms[i]->ctx->device));
        for (size_t j = 0; j < n; ++j)
            if (j != i && g->ms[j]->ctx->stream != g->ms[i]->ctx->stream)
                BIAS_CUDA_TRY(cudaStreamWaitEvent(g->ms[i]->ctx->stream, g->ev[j], 0));
    }
    return BIAS_OK;
}

#define GROUP_EACH(expr)                                          \
    for (bias_model * m : g->ms) {                                \
        BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));             \
        BIAS_TRY(expr);                                           \
    }

static int group_prepare(bias_group *g) {
    GROUP_EACH(step_publish_freq(m));
    BIAS_TRY(group_barrier(g));
    GROUP_EACH(step_classify_build(m));
    BIAS_TRY(group_barrier(g));
    GROUP_EACH(step_exchange(m, 1));
    BIAS_TRY(group_barrier(g));
    return BIAS_OK;
}

extern "C" {

static int group_setup(bias_group *g, int32_t hot_cap) {
    const int n = (int)g->ms.size();
    for (int i = 0; i < n; ++i) {
        bias_model *m = g->ms[i];
        BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
        BIAS_TRY(ensure_wide(m));
        BIAS_TRY(narrow_alloc(m, hot_cap));
        BIAS_CUDA_TRY(cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming));
        for (int j = 0; j < n; ++j)
            if (g->ms[j]->ctx->device != m->ctx->device) {
                int can = 0;
                BIAS_CUDA_TRY(cudaDeviceCanAccessPeer(&can, m->ctx->device, g->ms[j]->ctx->device));
                if (!can) return fail(BIAS_ENODEV, "device %d cannot address the memory of device %d", m->ctx->device, g->ms[j]->ctx->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(g->ms[j]->ctx->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(BIAS_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
    }
    for (int i = 0; i < n; ++i) {
        bias_model *m = g->ms[i];
        BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
        BIAS_TRY(comm_buffers(m, n));
    }
    // nothing can fail from here on: only now do the models become shards
    for (int i = 0; i < n; ++i) {
        bias_model *m = g->ms[i];
        m->rank = i;
        m->world = n;
        m->grouped = true;
        m->seed_mix = 0x9E3779B97F4A7C15ull * (uint64_t)i;
        for (int j = 0; j < n; ++j) m->peer_base[j] = g->ms[j]->share_base;
        m->narrow_valid = false;        // the tables hold this shard's counts only: the first sweep merges them
        m->class_valid = false;
    }
    return BIAS_OK;
}

int bias_group_create(bias_model **models, int32_t n, bias_group **out) {
    if (!models || !out || n < 1) return fail(BIAS_EINVAL, "null argument");
    *out = nullptr;
    if (n > kMaxPeers) return fail(BIAS_EINVAL, "at most %d shards (the GPUs of one NVSwitch domain)", kMaxPeers);
    int64_t cap = 0;
    for (int i = 0; i < n; ++i) {
        if (!models[i]) return fail(BIAS_EINVAL, "null model");
        BIAS_TRY(sharded_ok(models[i]));
        if (models[i]->world != 1) return fail(BIAS_ESTATE, "shard %d already belongs to a sharded run", i);
        if (models[i]->V != models[0]->V || models[i]->Kpad != models[0]->Kpad || models[i]->K != models[0]->K)
            return fail(BIAS_EINVAL, "the shards of a group must be the same model (KK, dd)");
        for (int j = 0; j < i; ++j)
            if (models[j] == models[i]) return fail(BIAS_EINVAL, "shard %d is listed twice", i);
        cap += models[i]->N_cap;
    }
    bias_group *g = new (std::nothrow) bias_group();
    if (!g) return fail(BIAS_ENOMEM, "out of host memory");
    g->ms.assign(models, models + n);
    g->ev.assign(n, nullptr);
    const int rc = group_setup(g, narrow_hot_cap(models[0]->V, cap));
    if (rc != BIAS_OK) {
        for (cudaEvent_t e : g->ev)
            if (e) cudaEventDestroy(e);
        delete g;
        return rc;
    }
    *out = g;
    return BIAS_OK;
}

int bias_group_destroy(bias_group *g) {
    if (!g) return BIAS_OK;
    for (size_t i = 0; i < g->ms.size(); ++i) {
        cudaSetDevice(g->ms[i]->ctx->device);
        cudaStreamSynchronize(g->ms[i]->ctx->stream);
        if (g->ev[i]) cudaEventDestroy(g->ev[i]);
    }
    delete g;
    return BIAS_OK;
}

int bias_group_sweep(bias_group *g, int32_t n_sweeps) {
    if (!g) return fail(BIAS_EINVAL, "null group");
    if (n_sweeps < 0) return fail(BIAS_EINVAL, "n_sweeps < 0");
    for (int32_t s = 0; s < n_sweeps; ++s) {
        bool stale = false;
        for (bias_model *m : g->ms) stale = stale || !m->narrow_valid;
        if (stale) {
            for (bias_model *m : g->ms)
                if (m->narrow_valid) return fail(BIAS_ESTATE, "the shards of a group must be (re)loaded together");
            BIAS_TRY(group_prepare(g));
        }
        for (bias_model *m : g->ms) {
            BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
            BIAS_TRY(model_local_sweep(m));
            BIAS_TRY(step_publish_nk(m));
        }
        BIAS_TRY(group_barrier(g));
        GROUP_EACH(step_exchange(m, 0));
        BIAS_TRY(group_barrier(g));
    }
    for (bias_model *m : g->ms)
        if (m->stage16 && n_sweeps > 0) {
            BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
            BIAS_TRY(model_mirror16(m));
        }
    return BIAS_OK;
}

// collapsed_gibbs_sampler!(lda, xx, zz, ...) on several GPUs of this process: documents split contiguously over
// `devices`, one context and resident model per device, sweeps through bias_group_sweep, assignments gathered back
// into zz.  stats[s] is filled after every sweep only when stats != NULL (n_moved and the diagnostic summed over shards).
// where the most recent bias_lda_run_multi of this thread spent its time: set-up (contexts, shard upload, initialisation pass, peer
// mappings), the sweeps (with their exchanges), the rest (download, teardown); milliseconds of host wall clock with the devices idle at
// both ends of each span
static thread_local double g_run_ms[3] = {0.0, 0.0, 0.0};

int bias_lda_run_multi(const int32_t *devices, int32_t n_dev, uint64_t seed, const bias_conjugate *conj, const bias_data *data,
                       int64_t n_groups, const int64_t *group_ptr, int32_t *zz, int32_t K, double alpha, int32_t n_sweeps,
                       bias_sweep_stats *stats) {
    if (!devices || !conj || !data || !group_ptr || !zz) return fail(BIAS_EINVAL, "null argument");
    if (n_dev < 1 || n_dev > kMaxPeers) return fail(BIAS_EINVAL, "n_dev must be 1..%d", kMaxPeers);
    if (n_sweeps < 0) return fail(BIAS_EINVAL, "n_sweeps < 0");
    if (conj->kind != BIAS_MD || conj->datum != BIAS_INT || !data->word)
        return fail(BIAS_EINVAL, "bias_lda_run_multi covers LDA x MultinomialDirichlet x word ids");
    if (n_groups < 0 || group_ptr[0] != 0 || group_ptr[n_groups] != data->n_items) return fail(BIAS_EINVAL, "group_ptr must start at 0 and end at n_items");
    const auto t_start = std::chrono::steady_clock::now();
    g_run_ms[0] = g_run_ms[1] = g_run_ms[2] = 0.0;
    std::vector<bias_ctx *> ctxs(n_dev, nullptr);
    std::vector<bias_model *> ms(n_dev, nullptr);
    std::vector<std::vector<int64_t>> ptrs(n_dev);
    std::vector<int64_t> item0(n_dev + 1, 0);
    bias_group *g = nullptr;
    int rc = BIAS_OK;
    for (int i = 0; i < n_dev; ++i) {
        const int64_t g0 = n_groups * i / n_dev, g1 = n_groups * (i + 1) / n_dev;
        item0[i] = group_ptr[g0];
        item0[i + 1] = group_ptr[g1];
        ptrs[i].resize((size_t)(g1 - g0) + 1);
        for (int64_t j = g0; j <= g1; ++j) ptrs[i][(size_t)(j - g0)] = group_ptr[j] - group_ptr[g0];
    }
    // the shards are created (allocations, upload, initialisation pass) and later read back and freed by one host thread per
    // device: done one device after the other these are a quarter of a 100-sweep call on 8 GPUs
    std::vector<int> rcs(n_dev, BIAS_OK);
    std::vector<std::string> errs(n_dev);
    auto on_every_device = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (int i = 0; i < n_dev; ++i)
            th.emplace_back([&, i] {
                rcs[i] = fn(i);
                if (rcs[i] != BIAS_OK) errs[i] = bias_last_error();     // the message lives in the worker's thread
            });
        for (auto &t : th) t.join();
        for (int i = 0; i < n_dev; ++i)
            if (rcs[i] != BIAS_OK) return fail(rcs[i], "device %d: %s", devices[i], errs[i].c_str());
        return (int)BIAS_OK;
    };
    rc = on_every_device([&](int i) {
        bias_data d = *data;
        d.n_items = item0[i + 1] - item0[i];
        d.word = data->word + item0[i];
        int r = bias_ctx_create(devices[i], seed, nullptr, &ctxs[i]);
        if (r == BIAS_OK)
            r = bias_model_create(ctxs[i], BIAS_MODEL_LDA, conj, &d, (int64_t)ptrs[i].size() - 1, ptrs[i].data(), zz + item0[i], K, alpha,
                                  nullptr, &ms[i]);
        return r;
    });
    if (rc == BIAS_OK && n_dev > 1) rc = bias_group_create(ms.data(), n_dev, &g);
    for (int i = 0; i < n_dev && rc == BIAS_OK; ++i) rc = bias_ctx_sync(ctxs[i]);
    const auto t_sweeps = std::chrono::steady_clock::now();
    g_run_ms[0] = std::chrono::duration<double, std::milli>(t_sweeps - t_start).count();
    for (int32_t s = 0; s < n_sweeps && rc == BIAS_OK; ++s) {
        rc = n_dev > 1 ? bias_group_sweep(g, 1) : bias_model_sweep(ms[0], 1);
        if (rc == BIAS_OK && stats) {
            bias_sweep_stats acc{};
            for (int i = 0; i < n_dev && rc == BIAS_OK; ++i) {
                bias_sweep_stats one{};
                rc = bias_model_last_stats(ms[i], &one);
                acc.log_likelihood += one.log_likelihood;
                acc.n_moved += one.n_moved;
                acc.KK = one.KK;
                acc.aa = one.aa;
                acc.gg = one.gg;
            }
            stats[s] = acc;
        }
    }
    for (int i = 0; i < n_dev && rc == BIAS_OK; ++i) rc = bias_ctx_sync(ctxs[i]);
    const auto t_done = std::chrono::steady_clock::now();
    g_run_ms[1] = std::chrono::duration<double, std::milli>(t_done - t_sweeps).count();
    if (rc == BIAS_OK) rc = on_every_device([&](int i) { return bias_model_get_zz(ms[i], zz + item0[i]); });
    if (g) bias_group_destroy(g);
    {
        std::string keep = rc != BIAS_OK ? bias_last_error() : "";
        on_every_device([&](int i) {
            if (ms[i]) bias_model_destroy(ms[i]);
            if (ctxs[i]) bias_ctx_destroy(ctxs[i]);
            return (int)BIAS_OK;
        });
        if (rc != BIAS_OK) fail(rc, "%s", keep.c_str());
    }
    g_run_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_done).count();
    return rc;
}

int bias_lda_run_multi_timing(double *ms) {
    if (!ms) return fail(BIAS_EINVAL, "null argument");
    for (int i = 0; i < 3; ++i) ms[i] = g_run_ms[i];
    return BIAS_OK;
}

// ---- one process per GPU ----
int bias_model_share_prepare(bias_model *m, int32_t rank, int32_t world, int64_t n_items_cap_global) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(BIAS_EINVAL, "rank %d of %d: at most %d ranks (the GPUs of one NVSwitch domain)", rank, world, kMaxPeers);
    BIAS_TRY(sharded_ok(m));
    if (n_items_cap_global < m->N_cap) return fail(BIAS_EINVAL, "the global capacity is smaller than this rank's");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_wide(m));
    BIAS_TRY(narrow_alloc(m, narrow_hot_cap(m->V, n_items_cap_global)));
    m->rank = rank;
    m->world_pending = world;
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));      // the counters are zero before any peer can see the block
    return BIAS_OK;
}

int bias_model_share_export(bias_model *m, bias_share_handle *out) {
    if (!m || !out) return fail(BIAS_EINVAL, "null argument");
    if (!m->share_base || m->world_pending < 1) return fail(BIAS_ESTATE, "bias_model_share_prepare has not been called");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    memset(out, 0, sizeof *out);
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof out->ipc, "handle size");
    cudaIpcMemHandle_t h;
    BIAS_CUDA_TRY(cudaIpcGetMemHandle(&h, m->share_base));
    memcpy(out->ipc, &h, sizeof h);
    out->bytes = (int64_t)m->share_bytes;
    out->dd = m->V;
    out->Kpad = m->Kpad;
    out->hot_cap = m->hot_cap;
    out->rank = m->rank;
    out->device = m->ctx->device;
    return BIAS_OK;
}

int bias_model_share_attach(bias_model *m, const bias_share_handle *all, int32_t world) {
    if (!m || !all) return fail(BIAS_EINVAL, "null argument");
    if (world != m->world_pending) return fail(BIAS_ESTATE, "bias_model_share_prepare was called for %d ranks", m->world_pending);
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    for (int r = 0; r < world; ++r) {
        const bias_share_handle &h = all[r];
        if (h.rank != r) return fail(BIAS_EINVAL, "handle %d belongs to rank %d", r, h.rank);
        if (h.bytes != (int64_t)m->share_bytes || h.dd != m->V || h.Kpad != m->Kpad || h.hot_cap != m->hot_cap)
            return fail(BIAS_EINVAL, "rank %d shares a different model (dd, KK or capacity)", r);
        if (r == m->rank) {
            m->peer_base[r] = m->share_base;
            continue;
        }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, h.ipc, sizeof ih);
        void *p = nullptr;
        BIAS_CUDA_TRY(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        m->peer_base[r] = static_cast<unsigned char *>(p);
        m->peer_ipc[r] = true;
    }
    m->world = world;
    m->seed_mix = 0x9E3779B97F4A7C15ull * (uint64_t)m->rank;
    BIAS_TRY(comm_buffers(m, world));
    m->narrow_valid = false;
    m->class_valid = false;
    return BIAS_OK;
}

int bias_model_last_exchange(bias_model *m, float *ms_out, int64_t *bytes_in_out) {
    // device time of the most recent exchange kernel of a one-process-per-GPU run (between its two barriers) and the
    // bytes this rank pulled from its peers in it (the same number goes out): the NVLink figure of bench.py
    if (!m || !ms_out) return fail(BIAS_EINVAL, "null argument");
    if (m->world <= 1 || !m->ev_x1) return fail(BIAS_ESTATE, "no exchange has been timed");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_CUDA_TRY(cudaEventSynchronize(m->ev_x1));
    BIAS_CUDA_TRY(cudaEventElapsedTime(ms_out, m->ev_x0, m->ev_x1));
    if (bytes_in_out) {
        const int64_t rows = m->V * (m->rank + 1) / m->world - m->V * m->rank / m->world;
        const int64_t hot = (int64_t)m->n_hot_host * (m->rank + 1) / m->world - (int64_t)m->n_hot_host * m->rank / m->world;
        *bytes_in_out = (rows * m->Kpad * 2 + hot * m->Kpad * 4) * (m->world - 1);
    }
    return BIAS_OK;
}

int bias_model_replica_digest(bias_model *m, uint64_t *digest) {
    // order-independent 64-bit digest of the replica's word-topic counts and n_k (equal replicas <=> equal digests, up
    // to hash collisions): what a sharded run compares across ranks after an exchange
    if (!m || !digest) return fail(BIAS_EINVAL, "null argument");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(comm_check(m));
    BIAS_TRY(ensure_wide(m));
    unsigned long long *d = nullptr;
    BIAS_CUDA_TRY(cudaMalloc((void **)&d, 8));
    cudaMemsetAsync(d, 0, 8, m->ctx->stream);
    const int64_t n = m->V * (int64_t)m->Kpad + m->Kpad;
    digest_kernel<<<std::max(1, (int)std::min<int64_t>((n + 255) / 256, (int64_t)m->ctx->num_sms * 8)), 256, 0, m->ctx->stream>>>(m->i32_block, n, d);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, m->ctx->stream);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(BIAS_ECUDA, "digest: %s", cudaGetErrorString(e));
    *digest = h;
    return BIAS_OK;
}

}  // extern "C"
