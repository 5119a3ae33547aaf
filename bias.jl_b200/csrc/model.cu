// model.cu — C ABI of libbias_cuda: context, resident sampler state, LDA / BMM sweeps,
// verification entry points and the multi-GPU delta exchange.  HDP lives in hdp.cu.
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <new>
#include "model.hpp"
#include "lda_fast.cuh"
#include "lda_half.cuh"
#include "lda_half16.cuh"
#define BIAS_GENERIC_KERNEL_IMPL
#include "generic.cuh"
#include "util_kernels.cuh"

namespace bias {

static thread_local std::string g_err;

bool is_dynamic_K(const bias_model *m) {
    return m->model == BIAS_MODEL_HDP || m->model == BIAS_MODEL_RCRP || m->model == BIAS_MODEL_CRF || m->model == BIAS_MODEL_RCRF;
}

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

template <typename T>
static int dev_alloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    BIAS_CUDA_TRY(cudaMalloc((void **)p, n * sizeof(T)));
    return BIAS_OK;
}

static int grid_for(bias_ctx *ctx, int64_t n, int threads, int per_sm = 8) {
    int64_t blocks = (n + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->num_sms * per_sm;
    return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

// ---------------------------------------------------------------------------------------
// fast LDA kernel launch
// ---------------------------------------------------------------------------------------
template <int NCH>
static int launch_lda_fast_t(bias_model *m, const LdaFastParams &p) {
    static int ctas_cache[kMaxDevices] = {};          // function attributes are per device
    int &ctas_per_sm = ctas_cache[m->ctx->device % kMaxDevices];
    const size_t smem = lda_fast_smem_bytes(NCH);
    const int threads = lda_fast_warps(NCH) * 32;
    if (ctas_per_sm == 0) {
        BIAS_CUDA_TRY(cudaFuncSetAttribute(lda_md_int_sweep_kernel<NCH>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        BIAS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lda_md_int_sweep_kernel<NCH>, threads, smem));
        if (occ < 1) return fail(BIAS_ECUDA, "LDA sweep kernel does not fit on an SM (K padded to %d)", NCH * 128);
        ctas_per_sm = occ;
    }
    int64_t want = (p.D + lda_fast_warps(NCH) - 1) / lda_fast_warps(NCH);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)m->ctx->num_sms * ctas_per_sm));
    lda_md_int_sweep_kernel<NCH><<<grid, threads, smem, m->ctx->stream>>>(p);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

template <int NCK, int WARPS = kHalfWarps>
static int launch_lda_half_t(bias_model *m, const LdaFastParams &p) {
    static int ctas_cache[kMaxDevices] = {};          // function attributes are per device
    int &ctas_per_sm = ctas_cache[m->ctx->device % kMaxDevices];
    const size_t smem = lda_half_smem_bytes(NCK, WARPS);
    const int threads = WARPS * 32;
    if (ctas_per_sm == 0) {
        BIAS_CUDA_TRY(cudaFuncSetAttribute(lda_md_int_sweep_half_kernel<NCK, WARPS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        BIAS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lda_md_int_sweep_half_kernel<NCK, WARPS>, threads, smem));
        if (occ < 1) return fail(BIAS_ECUDA, "LDA half-warp sweep kernel does not fit on an SM (K padded to %d)", NCK * 64);
        ctas_per_sm = occ;
    }
    const int64_t want = (p.D + WARPS * 2 - 1) / (WARPS * 2);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)m->ctx->num_sms * ctas_per_sm));
    lda_md_int_sweep_half_kernel<NCK, WARPS><<<grid, threads, smem, m->ctx->stream>>>(p);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

template <int NC8, int WARPS, int CTAS, bool CNT8, bool HOT>
static int launch_lda_half16_t(bias_model *m, const LdaFastParams &p) {
    static int ctas_cache[kMaxDevices] = {};          // function attributes are per device
    int &ctas_per_sm = ctas_cache[m->ctx->device % kMaxDevices];
    const size_t smem = lda_half16_smem_bytes(NC8, WARPS, CNT8);
    const int threads = WARPS * 32;
    if (ctas_per_sm == 0) {
        BIAS_CUDA_TRY(cudaFuncSetAttribute(lda_md_int_sweep_half16_kernel<NC8, WARPS, CTAS, CNT8, HOT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        BIAS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lda_md_int_sweep_half16_kernel<NC8, WARPS, CTAS, CNT8, HOT>, threads, smem));
        if (occ < 1) return fail(BIAS_ECUDA, "LDA 16-bit sweep kernel does not fit on an SM (K padded to %d)", NC8 * 128);
        ctas_per_sm = std::min(occ, CTAS);
    }
    const int64_t want = (p.D + WARPS * 2 - 1) / (WARPS * 2);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)m->ctx->num_sms * ctas_per_sm));
    lda_md_int_sweep_half16_kernel<NC8, WARPS, CTAS, CNT8, HOT><<<grid, threads, smem, m->ctx->stream>>>(p);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

// the document's own topic counts fit a byte when no document has more than 255 tokens: 5 KB instead of 6 KB of shared
// memory per document at K = 1024, 42 instead of 36 documents per SM (7 instead of 6 warps per CTA)
template <int NC8>
static int launch_lda_half16(bias_model *m, const LdaFastParams &p) {
    const char *force = getenv("BIAS_LDA_CNT8");
    const bool cnt8 = force ? atoi(force) != 0 && m->max_group_len <= 255 : m->max_group_len <= 255;
    if (m->n_hot_host > 0 || getenv("BIAS_LDA_FORCE_HOT")) {       // (the switch: the same kernel on a corpus without hot words, for profiling)
        // some words need int32 rows: the instantiation with the hot-row route keeps a whole 4 KB row in registers
        // (up to 128 per thread: 2 CTAs of 8 warps per SM)
        if (cnt8) return launch_lda_half16_t<NC8, kH16HotWarps, kH16HotCtas, true, true>(m, p);
        return launch_lda_half16_t<NC8, kH16HotWarps, kH16HotCtas, false, true>(m, p);
    }
    if (cnt8) return launch_lda_half16_t<NC8, kH16Warps8, kH16Ctas, true, false>(m, p);
    return launch_lda_half16_t<NC8, kH16Warps, kH16Ctas, false, false>(m, p);
}

// the narrow tables (lda_half16.cuh): the default whenever they were allocated (K <= 1024) and every document is shorter
// than 65,536 tokens; BIAS_LDA_KERNEL=half / reg keep the sweeps on the int32 table
static bool use_half16_kernel(const bias_model *m) {
    const char *force = getenv("BIAS_LDA_KERNEL");
    if (force && (force[0] == 'r' || force[0] == 'h')) return false;
    return m->n_wk16 && m->Kpad <= 1024 && m->max_group_len <= 65535;
}

// two documents per warp (lda_half.cuh): the default for K <= 1024 and documents shorter than 65,536 tokens;
// BIAS_LDA_KERNEL=reg selects the one-document-per-warp kernel
static bool use_half_kernel(const bias_model *m) {
    const char *force = getenv("BIAS_LDA_KERNEL");
    if (force && force[0] == 'r') return false;
    return m->Kpad <= 1024 && m->max_group_len <= 65535;
}

static int launch_lda_fast(bias_model *m, const double *ext_uniforms, int frozen, int32_t *draws_out, int eval = 0,
                           const int64_t *eval_ptr = nullptr, const int32_t *eval_word = nullptr) {
    LdaFastParams p{};
    p.eval = eval;
    p.eval_ptr = eval_ptr;
    p.eval_word = eval_word;
    p.eval_sum = reinterpret_cast<double *>(m->scratch + 6);
    p.doc_ptr = m->group_ptr;
    p.word = m->word;
    p.z = m->z;
    p.n_wk = m->n_wk;
    p.n_k = m->n_k;
    p.D = m->G;
    p.K = m->K;
    p.Kpad = m->Kpad;
    p.V = (int32_t)m->V;
    p.alpha = (float)m->alpha;
    p.beta = (float)m->conj.aa;
    p.vbeta = (float)(m->conj.aa * (double)m->V);
    p.doc_counter = m->scratch + 2;
    p.seed = m->ctx->seed ^ m->seed_mix;
    p.sweep = m->sweep_index;
    p.ext_uniforms = ext_uniforms;
    p.draws_out = draws_out;
    p.frozen = frozen;
    p.diag_sum = reinterpret_cast<double *>(m->scratch);
    p.moved = m->scratch + 1;
    p.pf_dist = 2;            // rows prefetched into L2 ahead of the token loop (int32-row half-warp kernel)
    p.nk_rep = m->nk_rep;
    const bool half_path = use_half_kernel(m) && !eval && m->nk_rep;
    if (use_half16_kernel(m) && !eval && m->V > 0 && m->nk_rep) {
        BIAS_TRY(ensure_narrow(m));
        BIAS_TRY(narrow_replicate(m));
        p.n_wk16 = m->n_wk16;
        p.word_hot = m->word_hot;
        p.hot_cap = m->hot_cap;
        p.hot_rep = m->hot_rep;
        switch (m->Kpad / 128) {
            case 1: BIAS_TRY(launch_lda_half16<1>(m, p)); break;
            case 2: BIAS_TRY(launch_lda_half16<2>(m, p)); break;
            case 4: BIAS_TRY(launch_lda_half16<4>(m, p)); break;
            case 8: BIAS_TRY(launch_lda_half16<8>(m, p)); break;
            default: return fail(BIAS_EINVAL, "unsupported padded K %d for the 16-bit LDA kernel", m->Kpad);
        }
        if (!frozen) {
            fold_nk_kernel<<<(m->Kpad + 255) / 256, 256, 0, m->ctx->stream>>>(m->n_k, m->nk_rep, m->Kpad, kNkRep);
            BIAS_CUDA_TRY(cudaGetLastError());
            m->ctx->launches += 1;
            m->wide_valid = false;          // the int32 view is refreshed when somebody asks for it (ensure_wide)
        }
        return BIAS_OK;
    }
    BIAS_TRY(ensure_wide(m));
    if (!frozen && !eval) m->narrow_valid = false;
    if (half_path) {
        switch (m->Kpad / 64) {
            case 2: BIAS_TRY(launch_lda_half_t<2>(m, p)); break;
            case 4: BIAS_TRY(launch_lda_half_t<4>(m, p)); break;
            case 8: BIAS_TRY(launch_lda_half_t<8>(m, p)); break;
            case 16: BIAS_TRY(launch_lda_half_t<16>(m, p)); break;
            default: return fail(BIAS_EINVAL, "unsupported padded K %d for the half-warp LDA kernel", m->Kpad);
        }
        if (!frozen) {
            fold_nk_kernel<<<(m->Kpad + 255) / 256, 256, 0, m->ctx->stream>>>(m->n_k, m->nk_rep, m->Kpad, kNkRep);
            BIAS_CUDA_TRY(cudaGetLastError());
            m->ctx->launches += 1;
        }
        return BIAS_OK;
    }
    switch (m->Kpad / 128) {
        case 1: return launch_lda_fast_t<1>(m, p);
        case 2: return launch_lda_fast_t<2>(m, p);
        case 4: return launch_lda_fast_t<4>(m, p);
        case 8: return launch_lda_fast_t<8>(m, p);
        case 16: return launch_lda_fast_t<16>(m, p);
        case 32: return launch_lda_fast_t<32>(m, p);
    }
    return fail(BIAS_EINVAL, "unsupported padded K %d for the fast LDA kernel", m->Kpad);
}

// ---------------------------------------------------------------------------------------
// generic kernel launch
// ---------------------------------------------------------------------------------------
template <int KIND, int DATUM>
static int launch_generic_t(bias_model *m, const GenericParams &p) {
    const size_t smem = generic_smem_bytes(p.n_out);
    static size_t smem_cache[kMaxDevices] = {};
    size_t &smem_set = smem_cache[m->ctx->device % kMaxDevices];
    static int ctas_cache[kMaxDevices] = {};          // function attributes are per device
    int &ctas_per_sm = ctas_cache[m->ctx->device % kMaxDevices];
    if (smem > 48 * 1024 && smem > smem_set) {
        BIAS_CUDA_TRY(cudaFuncSetAttribute(generic_sweep_kernel<KIND, DATUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    int occ = 0;
    BIAS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, generic_sweep_kernel<KIND, DATUM>, kGenThreads, smem));
    ctas_per_sm = std::max(1, occ);
    const int64_t blocks = (p.n_work + kGenWarps - 1) / kGenWarps;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)m->ctx->num_sms * ctas_per_sm));
    generic_sweep_kernel<KIND, DATUM><<<grid, kGenThreads, smem, m->ctx->stream>>>(p);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

int launch_generic(bias_model *m, int64_t n_work, const int64_t *item_idx, const int32_t *work_idx32,
                   const int32_t *row_of, int32_t *prior_rows, const double *ext_uniforms, int frozen,
                   int faithful, double *raw_out, double *prob_out, int32_t *draw_out, uint32_t rng_stream,
                   int row_by_work) {
    if (n_work <= 0) return BIAS_OK;
    BIAS_TRY(ensure_wide(m));
    if (!frozen) m->narrow_valid = false;
    GenericParams p{};
    p.model = m->model;
    p.kind = m->conj.kind;
    p.datum = m->conj.datum;
    p.K = is_dynamic_K(m) ? m->hdp.KK : m->K;
    p.Kpad = m->Kpad;
    p.n_out = is_dynamic_K(m) ? p.K + 1 : p.K;
    p.rcrp_T = (int32_t)m->G;
    p.rcrp_aa = m->d_aa_t;
    p.work_base = m->work_base;
    p.n_work = n_work;
    p.item_idx = item_idx;
    p.work_idx32 = work_idx32;
    p.row_of = row_of;
    p.row_by_work = row_by_work;
    p.prior_rows = prior_rows;
    p.prior_is_items = (prior_rows == m->n_items) ? 1 : 0;
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
    p.alpha = (m->model == BIAS_MODEL_HDP) ? m->hdp.aa : m->alpha;
    p.hdp_beta = m->d_beta;
    p.seed = m->ctx->seed ^ m->seed_mix;
    p.sweep = m->sweep_index;
    p.rng_stream = rng_stream;
    p.ext_uniforms = ext_uniforms;
    p.frozen = frozen;
    p.faithful = faithful;
    p.raw_out = raw_out;
    p.prob_out = prob_out;
    p.draw_out = draw_out;
    p.diag_sum = frozen ? nullptr : reinterpret_cast<double *>(m->scratch);
    p.moved = m->scratch + 1;
    p.pending_list = m->pending;
    p.pending_count = reinterpret_cast<unsigned int *>(m->scratch + 3);

    if (p.kind == BIAS_G1G1 && p.n_out <= kItemMaxOut && !faithful && !raw_out && !prob_out && m->model != BIAS_MODEL_RCRP &&
        !getenv("BIAS_NO_ITEM_KERNEL")) {
        // few Gaussian components: one thread per item (generic.cuh: gauss_item_sweep_kernel)
        const int64_t blocks = (p.n_work + kItemThreads - 1) / kItemThreads;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)m->ctx->num_sms * 8));
        gauss_item_sweep_kernel<<<grid, kItemThreads, gauss_item_smem_bytes(p.n_out), m->ctx->stream>>>(p);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
        return BIAS_OK;
    }
    if (p.kind == BIAS_G1G1) return launch_generic_t<BIAS_G1G1, BIAS_F64>(m, p);
    if (p.datum == BIAS_INT) return launch_generic_t<BIAS_MD, BIAS_INT>(m, p);
    return launch_generic_t<BIAS_MD, BIAS_SENT>(m, p);
}

// ---------------------------------------------------------------------------------------
// the two representations of the word-topic counts on the fast LDA path (lda_half16.cuh)
// ---------------------------------------------------------------------------------------
int ensure_wide(bias_model *m) {
    if (m->wide_valid || !m->n_wk16) return BIAS_OK;
    if (!m->narrow_valid) return fail(BIAS_ESTATE, "neither representation of the counts is current");
    return narrow_fold(m);
}

int ensure_narrow(bias_model *m) {
    if (m->narrow_valid) return BIAS_OK;
    if (!m->n_wk16) return fail(BIAS_ESTATE, "the model has no narrow tables");
    if (!m->wide_valid) return fail(BIAS_ESTATE, "neither representation of the counts is current");
    if (!m->class_valid) {
        BIAS_TRY(comm_global_word_freq(m));         // one GPU: the local frequencies; a sharded run sums them in its sweep call
        BIAS_TRY(narrow_classify(m));
    }
    return narrow_build(m);
}

// ---------------------------------------------------------------------------------------
// initialisation pass: additem! for every item
// ---------------------------------------------------------------------------------------
int rebuild_counts(bias_model *m) {
    cudaStream_t st = m->ctx->stream;
    m->wide_valid = true;           // the pass below writes the int32 table
    m->narrow_valid = false;
    m->class_valid = false;
    BIAS_CUDA_TRY(cudaMemsetAsync(m->i32_block, 0, (size_t)m->n_i32 * sizeof(int32_t), st));
    if (m->n_f64) BIAS_CUDA_TRY(cudaMemsetAsync(m->f64_block, 0, (size_t)m->n_f64 * sizeof(double), st));
    if (m->word_freq) BIAS_CUDA_TRY(cudaMemsetAsync(m->word_freq, 0, (size_t)std::max<int64_t>(1, m->V) * sizeof(int32_t), st));
    const int threads = 256;
    if (m->N > 0) {
        if (m->conj.kind == BIAS_MD && m->conj.datum == BIAS_INT) {
            build_md_int_kernel<<<grid_for(m->ctx, m->N, threads), threads, (size_t)m->Kpad * sizeof(int32_t), st>>>(
                m->word, m->z, m->N, m->Kpad, m->n_wk, m->n_k, m->n_items);
        } else if (m->conj.kind == BIAS_MD) {
            build_md_sent_kernel<<<grid_for(m->ctx, m->N * 32, threads), threads, 0, st>>>(
                m->sent_ptr, m->sent_word, m->sent_count, m->z, m->N, m->Kpad, m->n_wk, m->n_k, m->n_items);
        } else {
            build_g1g1_kernel<<<grid_for(m->ctx, m->N, threads, 2), threads, (size_t)m->Kpad * 12, st>>>(
                m->x, m->z, m->N, m->Kpad, m->sumx, m->n_items);
        }
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
        if (m->word_freq) {
            // occurrences of each word in the local shard: the bound on how far a cell of its row can grow in a sweep
            word_freq_kernel<<<grid_for(m->ctx, m->V * 32, 256), 256, 0, st>>>(m->n_wk, m->V, m->Kpad, m->word_freq);
            BIAS_CUDA_TRY(cudaGetLastError());
            m->ctx->launches += 1;
        }
        if (m->group_rows) {
            BIAS_CUDA_TRY(cudaMemsetAsync(m->group_rows, 0, (size_t)m->G * m->Kpad * sizeof(int32_t), st));
            build_group_rows_kernel<<<grid_for(m->ctx, m->N, threads), threads, 0, st>>>(m->z, m->group_of, m->N,
                                                                                         m->Kpad, m->group_rows);
            BIAS_CUDA_TRY(cudaGetLastError());
            m->ctx->launches += 1;
        }
    }
    return BIAS_OK;
}

// zz / word / Sent range checks on the uploaded arrays; one small D2H + sync
static int validate_on_device(bias_model *m, int32_t K) {
    if (m->N == 0) return BIAS_OK;
    cudaStream_t st = m->ctx->stream;
    int *flag = reinterpret_cast<int *>(m->scratch + 5);
    validate_kernel<<<grid_for(m->ctx, std::max<int64_t>(m->N, m->nnz), 256), 256, 0, st>>>(
        m->z, m->N, K, m->word, m->V, m->sent_word, m->sent_count, m->nnz, flag);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    int h = 0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    BIAS_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
    if (h & 1) return fail(BIAS_EINVAL, "zz holds a value outside [0, %d)", K);
    if (h & 2) return fail(BIAS_EINVAL, "a word id lies outside [0, %lld) (reference: BoundsError)", (long long)m->V);
    if (h & 4) return fail(BIAS_EINVAL, "a Sent entry has a word id outside [0, %lld) or a negative count", (long long)m->V);
    return BIAS_OK;
}

static int validate(const bias_conjugate *c, const bias_data *d, int64_t G, const int64_t *ptr, const int32_t *zz,
                    int32_t K, double alpha, int32_t model) {
    if (!c || !d || !zz) return fail(BIAS_EINVAL, "null argument");
    if (K < 1) return fail(BIAS_EINVAL, "KK must be >= 1");
    if (!(alpha > 0.0)) return fail(BIAS_EINVAL, "aa must be > 0");
    if (d->n_items < 0) return fail(BIAS_EINVAL, "negative item count");
    if (c->kind == BIAS_MD) {
        if (c->datum != BIAS_INT && c->datum != BIAS_SENT)
            return fail(BIAS_EINVAL, "MultinomialDirichlet takes Int or Sent observations (src/conjugates.jl:153-198)");
        if (c->dd < 1 || !(c->aa > 0.0)) return fail(BIAS_EINVAL, "MultinomialDirichlet: dd >= 1 and aa > 0 required");
        if (c->datum == BIAS_INT && !d->word) return fail(BIAS_EINVAL, "data.word is null");
        if (c->datum == BIAS_SENT && (!d->sent_ptr || !d->sent_word || !d->sent_count))
            return fail(BIAS_EINVAL, "Sent data pointers are null");
    } else if (c->kind == BIAS_G1G1) {
        if (c->datum != BIAS_F64) return fail(BIAS_EINVAL, "Gaussian1DGaussian1D takes Float64 observations (src/conjugates.jl:60-103)");
        if (!(c->v0 > 0.0) || !(c->vv > 0.0))
            return fail(BIAS_EINVAL, "Gaussian1D: the condition vv > zero(vv) is not satisfied (src/distributions.jl:19)");
        if (!d->x) return fail(BIAS_EINVAL, "data.x is null");
    } else {
        return fail(BIAS_EINVAL, "unknown conjugate kind %d", c->kind);
    }
    if (model != BIAS_MODEL_BMM) {
        if (!ptr || G < 0) return fail(BIAS_EINVAL, "group_ptr is required for grouped models");
        if (ptr[0] != 0 || ptr[G] != d->n_items) return fail(BIAS_EINVAL, "group_ptr must start at 0 and end at n_items");
        for (int64_t j = 0; j < G; ++j)
            if (ptr[j + 1] < ptr[j]) return fail(BIAS_EINVAL, "group_ptr must be non-decreasing");
    }
    if (c->kind == BIAS_MD && c->datum == BIAS_SENT && d->sent_ptr[0] != 0) return fail(BIAS_EINVAL, "sent_ptr must start at 0");
    // the O(N) range checks on zz / word ids / Sent entries run on the device (validate_kernel)
    return BIAS_OK;
}

int model_local_sweep(bias_model *m) {
    cudaStream_t st = m->ctx->stream;
    if (m->model == BIAS_MODEL_HDP) return hdp_sweep(m);
    if (m->model == BIAS_MODEL_RCRP) return rcrp_sweep(m);
    if (m->model == BIAS_MODEL_CRF) return crf_sweep(m, nullptr);
    if (m->model == BIAS_MODEL_RCRF) return rcrf_sweep(m, nullptr);
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch, 0, 4 * sizeof(unsigned long long), st));
    if (m->fast) {
        BIAS_TRY(launch_lda_fast(m, nullptr, 0, nullptr));
    } else {
        int32_t *prior = (m->model == BIAS_MODEL_BMM) ? m->n_items : m->group_rows;
        const int32_t *row_of = (m->model == BIAS_MODEL_BMM) ? nullptr : m->group_of;
        BIAS_TRY(launch_generic(m, m->N, nullptr, nullptr, row_of, prior, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                kStreamItemDraw));
    }
    m->sweep_index += 1;
    m->stats_pending = true;
    return BIAS_OK;
}


}  // namespace bias

using namespace bias;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int bias_version(void) { return 100; }

const char *bias_last_error(void) { return g_err.c_str(); }

int bias_device_count(int *n_out) {
    if (!n_out) return fail(BIAS_EINVAL, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *n_out = 0;
        return fail(BIAS_ENODEV, "no CUDA device: %s", cudaGetErrorString(e));
    }
    *n_out = n;
    return BIAS_OK;
}

int bias_ctx_create(int device, uint64_t seed, void *stream, bias_ctx **out) {
    if (!out) return fail(BIAS_EINVAL, "null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(BIAS_ENODEV, "libbias_cuda has no CPU fallback and found no CUDA device (%s)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return fail(BIAS_EINVAL, "device %d out of range (have %d)", device, n);
    BIAS_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    BIAS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(BIAS_ENODEV, "libbias_cuda is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    bias_ctx *c = new (std::nothrow) bias_ctx();
    if (!c) return fail(BIAS_ENOMEM, "out of host memory");
    c->device = device;
    c->seed = seed;
    c->num_sms = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) { delete c; return fail(BIAS_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(se)); }
        c->own_stream = true;
    }
    cudaEventCreate(&c->ev_a);
    cudaEventCreate(&c->ev_b);
    *out = c;
    return BIAS_OK;
}

static void ctx_release(bias_ctx *ctx);

int bias_ctx_destroy(bias_ctx *ctx) {
    if (!ctx) return BIAS_OK;
    if (ctx->live_models > 0) {          // models still launch on this stream: the last bias_model_destroy releases it
        ctx->closed = true;
        return BIAS_OK;
    }
    ctx_release(ctx);
    return BIAS_OK;
}

static void ctx_release(bias_ctx *ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->stirling) cudaFree(ctx->stirling);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int bias_ctx_sync(bias_ctx *ctx) {
    if (!ctx) return fail(BIAS_EINVAL, "null context");
    BIAS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BIAS_OK;
}

int bias_ctx_num_sms(bias_ctx *ctx, int *n_out) {
    if (!ctx || !n_out) return fail(BIAS_EINVAL, "null argument");
    *n_out = ctx->num_sms;
    return BIAS_OK;
}

int bias_ctx_launch_count(bias_ctx *ctx, int64_t *n_out) {
    if (!ctx || !n_out) return fail(BIAS_EINVAL, "null argument");
    *n_out = ctx->launches;
    return BIAS_OK;
}

// -----------------------------------------------------------------------------------------
int bias_model_destroy(bias_model *m) {
    if (!m) return BIAS_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    hdp_destroy(m);
    crf_destroy(m);
    void *ptrs[] = {m->word, m->x, m->sent_ptr, m->sent_word, m->sent_count, m->group_ptr, m->group_of, m->z,
                    m->i32_block, m->f64_block, m->group_rows, m->base_i32, m->delta_i32, m->base_f64,
                    m->delta_f64, m->scratch, m->d_aa_t, m->d_gg_t, m->word_freq, m->word_freq_g, m->word_hot, m->hot_info, m->nk_rep, m->stage16};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    narrow_free(m);
    if (m->h_scratch) cudaFreeHost(m->h_scratch);
    bias_ctx *ctx = m->ctx;
    delete m;
    if (--ctx->live_models == 0 && ctx->closed) ctx_release(ctx);
    return BIAS_OK;
}

int bias_model_create(bias_ctx *ctx, int32_t model, const bias_conjugate *conj, const bias_data *data,
                      int64_t n_groups, const int64_t *group_ptr, const int32_t *zz,
                      int32_t K, double alpha, const bias_hdp_params *hdp, bias_model **out) {
    if (!ctx || !out) return fail(BIAS_EINVAL, "null argument");
    *out = nullptr;
    if (model < BIAS_MODEL_LDA || model > BIAS_MODEL_RCRF) return fail(BIAS_EINVAL, "unknown model %d", model);
    const bool dynamic = model == BIAS_MODEL_HDP || model == BIAS_MODEL_RCRP || model == BIAS_MODEL_CRF || model == BIAS_MODEL_RCRF;
    if (dynamic) {
        if (!hdp) return fail(BIAS_EINVAL, "HDP / RCRP parameters are required");
        if (hdp->KK < 1 || hdp->Kmax < hdp->KK + 2) return fail(BIAS_EINVAL, "HDP: need 1 <= KK and Kmax >= KK + 2");
        if (model != BIAS_MODEL_RCRP && (!(hdp->gg > 0.0) || !(hdp->aa > 0.0))) return fail(BIAS_EINVAL, "HDP: gg and aa must be > 0");
        if (hdp->Kmax > 4096) return fail(BIAS_EINVAL, "HDP: Kmax is limited to 4096 component slots");
        if (hdp->n_internals < 0) return fail(BIAS_EINVAL, "HDP: n_internals < 0");
        K = hdp->KK;
        alpha = hdp->aa;
    }
    BIAS_TRY(validate(conj, data, n_groups, group_ptr, zz, K, alpha, model));
    BIAS_CUDA_TRY(cudaSetDevice(ctx->device));

    bias_model *m = new (std::nothrow) bias_model();
    if (!m) return fail(BIAS_ENOMEM, "out of host memory");
    m->ctx = ctx;
    ctx->live_models += 1;
    m->model = model;
    m->conj = *conj;
    m->N = data->n_items;
    m->G = (model == BIAS_MODEL_BMM) ? 0 : n_groups;
    m->K = K;
    m->alpha = alpha;
    m->V = (conj->kind == BIAS_MD) ? conj->dd : 0;
    for (int64_t j = 0; j < m->G; ++j) m->max_group_len = std::max(m->max_group_len, group_ptr[j + 1] - group_ptr[j]);
    const int32_t Kcap = dynamic ? hdp->Kmax : K;
    // pad K to 128 * 2^j (the fast kernel's chunk tree); the generic kernels only need the stride
    int32_t kp = 128;
    while (kp < Kcap) kp *= 2;
    m->Kpad = kp;
    const char *force_generic = getenv("BIAS_FORCE_GENERIC");
    m->fast = (model == BIAS_MODEL_LDA && conj->kind == BIAS_MD && conj->datum == BIAS_INT && kp <= 4096 &&
               !(force_generic && force_generic[0] == '1'));
    if (m->N >= (int64_t)1 << 31) { bias_model_destroy(m); return fail(BIAS_EINVAL, "more than 2^31-1 items per device"); }

    cudaStream_t st = ctx->stream;
    int rc = BIAS_OK;
#define M_TRY(expr) do { rc = (expr); if (rc != BIAS_OK) { bias_model_destroy(m); return rc; } } while (0)
#define M_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { bias_model_destroy(m); \
        return fail(_e == cudaErrorMemoryAllocation ? BIAS_ENOMEM : BIAS_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)

    // observations
    if (conj->datum == BIAS_INT) {
        M_TRY(dev_alloc(&m->word, (size_t)m->N));
        M_CUDA(cudaMemcpyAsync(m->word, data->word, (size_t)m->N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    } else if (conj->datum == BIAS_F64) {
        M_TRY(dev_alloc(&m->x, (size_t)m->N));
        M_CUDA(cudaMemcpyAsync(m->x, data->x, (size_t)m->N * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        m->nnz = data->sent_ptr[m->N];
        M_TRY(dev_alloc(&m->sent_ptr, (size_t)m->N + 1));
        M_TRY(dev_alloc(&m->sent_word, (size_t)m->nnz));
        M_TRY(dev_alloc(&m->sent_count, (size_t)m->nnz));
        M_CUDA(cudaMemcpyAsync(m->sent_ptr, data->sent_ptr, ((size_t)m->N + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        M_CUDA(cudaMemcpyAsync(m->sent_word, data->sent_word, (size_t)m->nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        M_CUDA(cudaMemcpyAsync(m->sent_count, data->sent_count, (size_t)m->nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    M_TRY(dev_alloc(&m->z, (size_t)m->N));
    M_CUDA(cudaMemcpyAsync(m->z, zz, (size_t)m->N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (model != BIAS_MODEL_BMM) {
        M_TRY(dev_alloc(&m->group_ptr, (size_t)m->G + 1));
        M_CUDA(cudaMemcpyAsync(m->group_ptr, group_ptr, ((size_t)m->G + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (!m->fast) {
            M_TRY(dev_alloc(&m->group_of, (size_t)m->N));
            if (m->N > 0) {
                group_of_kernel<<<grid_for(ctx, m->N, 256), 256, 0, st>>>(m->group_ptr, m->G, m->N, m->group_of);
                M_CUDA(cudaGetLastError());
                ctx->launches += 1;
            }
            M_TRY(dev_alloc(&m->group_rows, (size_t)m->G * m->Kpad));
        }
    }
    // component statistics
    const int64_t n_wk_elems = (conj->kind == BIAS_MD) ? m->V * (int64_t)m->Kpad : 0;
    m->n_i32 = n_wk_elems + 2 * (int64_t)m->Kpad;      // Kpad is a multiple of 128: stays 16-byte aligned
    M_TRY(dev_alloc(&m->i32_block, (size_t)m->n_i32));
    m->n_wk = m->i32_block;
    m->n_k = m->i32_block + n_wk_elems;
    m->n_items = m->n_k + m->Kpad;
    if (conj->kind == BIAS_G1G1) {
        m->n_f64 = m->Kpad;
        M_TRY(dev_alloc(&m->f64_block, (size_t)m->n_f64));
        m->sumx = m->f64_block;
    }
    if (m->fast && m->Kpad <= 1024) {
        M_TRY(narrow_alloc(m, narrow_hot_cap(m->V, m->N)));
        M_TRY(dev_alloc(&m->word_freq, (size_t)std::max<int64_t>(1, m->V)));
        M_TRY(dev_alloc(&m->word_freq_g, (size_t)std::max<int64_t>(1, m->V)));
        M_TRY(dev_alloc(&m->word_hot, (size_t)std::max<int64_t>(1, m->V)));
        M_TRY(dev_alloc(&m->hot_info, 4));
        M_TRY(dev_alloc(&m->nk_rep, (size_t)kNkRep * m->Kpad));
        M_CUDA(cudaMemsetAsync(m->nk_rep, 0, (size_t)kNkRep * m->Kpad * sizeof(int32_t), st));
    }
    M_TRY(dev_alloc(&m->scratch, 8));
    M_CUDA(cudaMemsetAsync(m->scratch, 0, 8 * sizeof(unsigned long long), st));
    M_CUDA(cudaMallocHost((void **)&m->h_scratch, 8 * sizeof(unsigned long long)));
    if (model != BIAS_MODEL_BMM) m->h_group_ptr.assign(group_ptr, group_ptr + m->G + 1);
    if (dynamic) M_TRY(hdp_create(m, hdp));
    M_TRY(validate_on_device(m, K));
    M_TRY(rebuild_counts(m));
    if (model == BIAS_MODEL_CRF) M_TRY(crf_create(m));
    // the uploads came from caller-owned pageable memory: finish them before returning
    M_CUDA(cudaStreamSynchronize(st));
#undef M_TRY
#undef M_CUDA
    m->last.KK = K;
    m->N_cap = m->N;
    m->G_cap = m->G;
    m->nnz_cap = m->nnz;
    *out = m;
    return BIAS_OK;
}

int bias_model_reload(bias_model *m, const bias_data *data, int64_t n_groups, const int64_t *group_ptr, const int32_t *zz) {
    if (!m || !data || !zz) return fail(BIAS_EINVAL, "null argument");
    if (is_dynamic_K(m)) return fail(BIAS_EINVAL, "reload is defined for the fixed-K models (LDA, BMM)");
    const int64_t G = (m->model == BIAS_MODEL_BMM) ? 0 : n_groups;
    BIAS_TRY(validate(&m->conj, data, G, group_ptr, zz, m->K, m->alpha, m->model));
    const int64_t nnz = (m->conj.datum == BIAS_SENT) ? data->sent_ptr[data->n_items] : 0;
    if (data->n_items > m->N_cap || G > m->G_cap || nnz > m->nnz_cap)
        return fail(BIAS_ERANGE, "reload: %lld items / %lld groups / %lld Sent entries exceed the capacity the model was "
                    "created with (%lld / %lld / %lld)", (long long)data->n_items, (long long)G, (long long)nnz,
                    (long long)m->N_cap, (long long)m->G_cap, (long long)m->nnz_cap);
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    m->N = data->n_items;
    m->G = G;
    m->nnz = nnz;
    if (m->conj.datum == BIAS_INT) {
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->word, data->word, (size_t)m->N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    } else if (m->conj.datum == BIAS_F64) {
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->x, data->x, (size_t)m->N * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->sent_ptr, data->sent_ptr, ((size_t)m->N + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->sent_word, data->sent_word, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->sent_count, data->sent_count, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->z, zz, (size_t)m->N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    m->mirror16_valid = false;
    if (m->model != BIAS_MODEL_BMM) {
        m->max_group_len = 1;
        for (int64_t j = 0; j < G; ++j) m->max_group_len = std::max(m->max_group_len, group_ptr[j + 1] - group_ptr[j]);
        m->h_group_ptr.assign(group_ptr, group_ptr + G + 1);
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->group_ptr, group_ptr, ((size_t)G + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (m->group_of && m->N > 0) {
            group_of_kernel<<<grid_for(m->ctx, m->N, 256), 256, 0, st>>>(m->group_ptr, m->G, m->N, m->group_of);
            BIAS_CUDA_TRY(cudaGetLastError());
            m->ctx->launches += 1;
        }
    }
    BIAS_TRY(validate_on_device(m, m->K));
    BIAS_TRY(rebuild_counts(m));
    m->sweep_index += 1;        // never reused: a re-fed batch must not replay the uniforms of the batch before it
    m->stats_pending = false;
    m->timing_valid = false;
    // the uploads came from caller-owned memory: finish them before returning
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    return BIAS_OK;
}

// ---- 16-bit host formats (LDA x word ids): half the PCIe bytes of reload / get_zz ----
static int64_t stage16_stride(const bias_model *m) { return (std::max<int64_t>(8, m->N_cap) + 7) / 8 * 8; }

// the uint16 mirror of z (third part of stage16): written at the end of every sweep call once the 16-bit formats
// are in use, so that get_zz_u16 is a plain copy-engine transfer that does not wait for SM time behind another
// stream's sweep
static int mirror16_update(bias_model *m) {
    if (m->N > 0) {
        narrow_u16_kernel<<<grid_for(m->ctx, m->N / 8 + 1, 256), 256, 0, m->ctx->stream>>>(m->z, m->N, m->stage16 + 2 * stage16_stride(m));
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
    }
    m->mirror16_valid = true;
    return BIAS_OK;
}

extern "C++" {
namespace bias {
int model_mirror16(bias_model *m) { return mirror16_update(m); }
}
}

static int stage16_ready(bias_model *m) {
    if (m->model != BIAS_MODEL_LDA || m->conj.kind != BIAS_MD || m->conj.datum != BIAS_INT)
        return fail(BIAS_EINVAL, "the 16-bit host formats are defined for LDA x MultinomialDirichlet x word ids");
    if (m->V > 65536 || m->K > 65536)
        return fail(BIAS_ERANGE, "the 16-bit host formats need dd <= 65536 and KK <= 65536 (dd = %lld, KK = %d)", (long long)m->V, m->K);
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    if (!m->stage16) BIAS_CUDA_TRY(cudaMalloc((void **)&m->stage16, (size_t)stage16_stride(m) * 3 * sizeof(uint16_t)));
    return BIAS_OK;
}

int bias_model_reload_u16(bias_model *m, int64_t n_items, const uint16_t *word, int64_t n_groups, const int64_t *group_ptr,
                          const uint16_t *zz) {
    if (!m || !word || !zz || !group_ptr) return fail(BIAS_EINVAL, "null argument");
    BIAS_TRY(stage16_ready(m));
    const int64_t G = n_groups;
    if (n_items < 0 || G < 0) return fail(BIAS_EINVAL, "negative item or group count");
    if (group_ptr[0] != 0 || group_ptr[G] != n_items) return fail(BIAS_EINVAL, "group_ptr must start at 0 and end at n_items");
    int64_t max_len = 1;
    for (int64_t j = 0; j < G; ++j) {
        if (group_ptr[j + 1] < group_ptr[j]) return fail(BIAS_EINVAL, "group_ptr must be non-decreasing");
        max_len = std::max(max_len, group_ptr[j + 1] - group_ptr[j]);
    }
    if (n_items > m->N_cap || G > m->G_cap)
        return fail(BIAS_ERANGE, "reload: %lld items / %lld groups exceed the capacity the model was created with (%lld / %lld)",
                    (long long)n_items, (long long)G, (long long)m->N_cap, (long long)m->G_cap);
    cudaStream_t st = m->ctx->stream;
    m->N = n_items;
    m->G = G;
    m->nnz = 0;
    uint16_t *w16 = m->stage16, *z16 = m->stage16 + stage16_stride(m);
    m->mirror16_valid = false;
    BIAS_CUDA_TRY(cudaMemcpyAsync(w16, word, (size_t)n_items * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(z16, zz, (size_t)n_items * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    if (n_items > 0) {
        widen_u16_kernel<<<grid_for(m->ctx, n_items / 8 + 1, 256), 256, 0, st>>>(w16, n_items, m->word);
        widen_u16_kernel<<<grid_for(m->ctx, n_items / 8 + 1, 256), 256, 0, st>>>(z16, n_items, m->z);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 2;
    }
    m->max_group_len = max_len;
    m->h_group_ptr.assign(group_ptr, group_ptr + G + 1);
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->group_ptr, group_ptr, ((size_t)G + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (m->group_of && m->N > 0) {
        group_of_kernel<<<grid_for(m->ctx, m->N, 256), 256, 0, st>>>(m->group_ptr, m->G, m->N, m->group_of);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
    }
    BIAS_TRY(validate_on_device(m, m->K));
    BIAS_TRY(rebuild_counts(m));
    m->sweep_index += 1;        // never reused: a re-fed batch must not replay the uniforms of the batch before it
    m->stats_pending = false;
    m->timing_valid = false;
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));     // the uploads came from caller-owned memory
    return BIAS_OK;
}

int bias_model_get_zz_u16(bias_model *m, uint16_t *zz) {
    if (!m || !zz) return fail(BIAS_EINVAL, "null argument");
    BIAS_TRY(stage16_ready(m));
    cudaStream_t st = m->ctx->stream;
    if (!m->mirror16_valid) BIAS_TRY(mirror16_update(m));
    BIAS_CUDA_TRY(cudaMemcpyAsync(zz, m->stage16 + 2 * stage16_stride(m), (size_t)m->N * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    return BIAS_OK;
}

int bias_model_set_zz(bias_model *m, const int32_t *zz) {
    if (!m || !zz) return fail(BIAS_EINVAL, "null argument");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    const int32_t Kcur = is_dynamic_K(m) ? m->hdp.KK : m->K;
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->z, zz, (size_t)m->N * sizeof(int32_t), cudaMemcpyHostToDevice, m->ctx->stream));
    m->mirror16_valid = false;
    BIAS_TRY(validate_on_device(m, Kcur));
    BIAS_TRY(rebuild_counts(m));
    if (m->model == BIAS_MODEL_CRF || m->model == BIAS_MODEL_RCRF) {   // the seating is rebuilt from the new assignments (src/HDP.jl:441-470)
        crf_destroy(m);
        BIAS_TRY(crf_create(m));
    }
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    return BIAS_OK;
}

int bias_model_get_zz(bias_model *m, int32_t *zz) {
    if (!m || !zz) return fail(BIAS_EINVAL, "null argument");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_CUDA_TRY(cudaMemcpyAsync(zz, m->z, (size_t)m->N * sizeof(int32_t), cudaMemcpyDeviceToHost, m->ctx->stream));
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    return BIAS_OK;
}

static int one_sweep(bias_model *m) {
    if (m->world > 1 && !m->grouped) {
        // one process per GPU: the sweep call is collective over the ranks (comm.cu)
        if (!m->narrow_valid) BIAS_TRY(comm_prepare_flags(m));
        BIAS_TRY(model_local_sweep(m));
        return comm_exchange_flags(m);
    }
    if (m->world > 1) return fail(BIAS_ESTATE, "this model is a shard of a bias_group: sweep the group");
    return model_local_sweep(m);
}

int bias_model_sweep(bias_model *m, int32_t n_sweeps) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (n_sweeps < 0) return fail(BIAS_EINVAL, "n_sweeps < 0");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    const int64_t l0 = m->ctx->launches;
    BIAS_CUDA_TRY(cudaEventRecord(m->ctx->ev_a, m->ctx->stream));
    for (int32_t s = 0; s < n_sweeps; ++s) BIAS_TRY(one_sweep(m));
    BIAS_CUDA_TRY(cudaEventRecord(m->ctx->ev_b, m->ctx->stream));
    m->mirror16_valid = false;
    if (m->stage16 && n_sweeps > 0) BIAS_TRY(mirror16_update(m));     // 16-bit formats in use: see mirror16_update
    m->last_launches = (int32_t)(m->ctx->launches - l0);
    m->timing_valid = true;
    return BIAS_OK;
}

int bias_model_last_sweep_ms(bias_model *m, float *ms_out, int32_t *launches_out) {
    if (!m || !ms_out) return fail(BIAS_EINVAL, "null argument");
    if (!m->timing_valid) return fail(BIAS_ESTATE, "no sweep has been timed yet");
    BIAS_CUDA_TRY(cudaEventSynchronize(m->ctx->ev_b));
    BIAS_CUDA_TRY(cudaEventElapsedTime(ms_out, m->ctx->ev_a, m->ctx->ev_b));
    if (launches_out) *launches_out = m->last_launches;
    return BIAS_OK;
}

static int fetch_stats(bias_model *m) {
    if (m->stats_pending) {
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->h_scratch, m->scratch, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                      m->ctx->stream));
        BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
        double d;
        memcpy(&d, &m->h_scratch[0], sizeof d);
        m->last.log_likelihood = d;
        m->last.n_moved = (int64_t)m->h_scratch[1];
        m->stats_pending = false;
    }
    m->last.KK = is_dynamic_K(m) ? m->hdp.KK : m->K;
    m->last.gg = m->hdp.gg;
    m->last.aa = (m->model == BIAS_MODEL_HDP) ? m->hdp.aa : m->alpha;
    return BIAS_OK;
}

int bias_model_last_stats(bias_model *m, bias_sweep_stats *out) {
    if (!m || !out) return fail(BIAS_EINVAL, "null argument");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(fetch_stats(m));
    *out = m->last;
    return BIAS_OK;
}

int bias_model_get_components(bias_model *m, int64_t *mi, int64_t *mm, double *mu, int64_t *nn) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    BIAS_TRY(comm_check(m));
    BIAS_TRY(ensure_wide(m));
    const int K = is_dynamic_K(m) ? m->hdp.KK : m->K;
    std::vector<int32_t> h_nk(m->Kpad), h_ni(m->Kpad);
    BIAS_CUDA_TRY(cudaMemcpyAsync(h_nk.data(), m->n_k, (size_t)m->Kpad * 4, cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaMemcpyAsync(h_ni.data(), m->n_items, (size_t)m->Kpad * 4, cudaMemcpyDeviceToHost, st));
    std::vector<double> h_sx;
    if (m->conj.kind == BIAS_G1G1) {
        h_sx.resize(m->Kpad);
        BIAS_CUDA_TRY(cudaMemcpyAsync(h_sx.data(), m->sumx, (size_t)m->Kpad * 8, cudaMemcpyDeviceToHost, st));
    }
    if (mi && m->conj.kind == BIAS_MD) {
        int64_t *d_mi = nullptr;
        BIAS_TRY(dev_alloc(&d_mi, (size_t)(m->V * K)));
        export_mi_kernel<<<grid_for(m->ctx, m->V * K, 256), 256, 0, st>>>(m->n_wk, m->V, K, m->Kpad, d_mi);
        m->ctx->launches += 1;
        cudaError_t e = cudaMemcpyAsync(mi, d_mi, (size_t)(m->V * K) * 8, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(d_mi);
        if (e != cudaSuccess) return fail(BIAS_ECUDA, "export of mi failed: %s", cudaGetErrorString(e));
    }
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    // the fp32 LDA kernel maintains n_k only; for word-id items nn == mm
    const bool int_items = (m->conj.kind == BIAS_MD && m->conj.datum == BIAS_INT);
    for (int k = 0; k < K; ++k) {
        if (mm && m->conj.kind == BIAS_MD) mm[k] = h_nk[k];
        if (nn) nn[k] = int_items ? h_nk[k] : h_ni[k];
        if (mu && m->conj.kind == BIAS_G1G1) mu[k] = h_ni[k] > 0 ? h_sx[k] / (double)h_ni[k] : m->conj.m0;
    }
    return BIAS_OK;
}

int bias_model_get_group_counts(bias_model *m, int64_t *nn) {
    if (!m || !nn) return fail(BIAS_EINVAL, "null argument");
    if (m->model == BIAS_MODEL_BMM) return fail(BIAS_EINVAL, "BMM has no groups; use bias_model_get_components");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    // computed on the host from z: posterior() is outside the hot path
    std::vector<int32_t> z((size_t)m->N);
    std::vector<int64_t> ptr((size_t)m->G + 1);
    BIAS_CUDA_TRY(cudaMemcpyAsync(z.data(), m->z, (size_t)m->N * 4, cudaMemcpyDeviceToHost, m->ctx->stream));
    BIAS_CUDA_TRY(cudaMemcpyAsync(ptr.data(), m->group_ptr, ((size_t)m->G + 1) * 8, cudaMemcpyDeviceToHost, m->ctx->stream));
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    const int K = is_dynamic_K(m) ? m->hdp.KK : m->K;
    std::fill(nn, nn + (size_t)m->G * K, (int64_t)0);
    for (int64_t j = 0; j < m->G; ++j)
        for (int64_t i = ptr[j]; i < ptr[j + 1]; ++i)
            if (z[i] >= 0 && z[i] < K) nn[j * K + z[i]] += 1;
    return BIAS_OK;
}

// -----------------------------------------------------------------------------------------
// verification
// -----------------------------------------------------------------------------------------
int bias_model_conditionals(bias_model *m, int64_t n_query, const int64_t *item_idx, const double *uniforms,
                            double *raw, double *prob, int32_t *draw) {
    if (!m || !item_idx || !uniforms || !draw) return fail(BIAS_EINVAL, "null argument");
    if (n_query <= 0) return BIAS_OK;
    for (int64_t q = 0; q < n_query; ++q)
        if (item_idx[q] < 0 || item_idx[q] >= m->N) return fail(BIAS_EINVAL, "query item out of range");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    const int K = is_dynamic_K(m) ? m->hdp.KK : m->K;
    const int n_out = is_dynamic_K(m) ? K + 1 : K;
    int64_t *d_idx = nullptr;
    double *d_u = nullptr, *d_raw = nullptr, *d_prob = nullptr;
    int32_t *d_draw = nullptr, *d_rows = nullptr, *d_rowof = nullptr;
    int rc = BIAS_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        for (void *p : {(void *)d_idx, (void *)d_u, (void *)d_raw, (void *)d_prob, (void *)d_draw, (void *)d_rows, (void *)d_rowof})
            if (p) cudaFree(p);
    };
#define V_TRY(expr) do { rc = (expr); if (rc != BIAS_OK) { cleanup(); return rc; } } while (0)
#define V_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(BIAS_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    V_TRY(dev_alloc(&d_idx, (size_t)n_query));
    V_TRY(dev_alloc(&d_u, (size_t)n_query));
    V_TRY(dev_alloc(&d_raw, (size_t)n_query * n_out));
    V_TRY(dev_alloc(&d_prob, (size_t)n_query * n_out));
    V_TRY(dev_alloc(&d_draw, (size_t)n_query));
    V_CUDA(cudaMemcpyAsync(d_idx, item_idx, (size_t)n_query * 8, cudaMemcpyHostToDevice, st));
    V_CUDA(cudaMemcpyAsync(d_u, uniforms, (size_t)n_query * 8, cudaMemcpyHostToDevice, st));
    int32_t *prior = nullptr;
    const int32_t *row_of = nullptr;
    if (m->model == BIAS_MODEL_BMM) {
        prior = m->n_items;
    } else if (m->group_rows) {
        prior = m->group_rows;
        row_of = m->group_of;
    } else {
        // the fast LDA path keeps no n_jk table: build the rows of the queried items' groups
        V_TRY(dev_alloc(&d_rows, (size_t)n_query * m->Kpad));
        V_CUDA(cudaMemsetAsync(d_rows, 0, (size_t)n_query * m->Kpad * 4, st));
        V_TRY(dev_alloc(&d_rowof, (size_t)n_query));
        build_query_rows_kernel<<<grid_for(m->ctx, n_query * 32, 128), 128, 0, st>>>(d_idx, n_query, m->group_ptr, m->G,
                                                                                    m->z, m->Kpad, d_rows, d_rowof);
        V_CUDA(cudaGetLastError());
        m->ctx->launches += 1;
        prior = d_rows;
        row_of = d_rowof;
    }
    V_TRY(launch_generic(m, n_query, d_idx, nullptr, row_of, prior, d_u, 1, 1, d_raw, d_prob, d_draw, kStreamItemDraw,
                         row_of == d_rowof ? 1 : 0));
    if (raw) V_CUDA(cudaMemcpyAsync(raw, d_raw, (size_t)n_query * n_out * 8, cudaMemcpyDeviceToHost, st));
    if (prob) V_CUDA(cudaMemcpyAsync(prob, d_prob, (size_t)n_query * n_out * 8, cudaMemcpyDeviceToHost, st));
    V_CUDA(cudaMemcpyAsync(draw, d_draw, (size_t)n_query * 4, cudaMemcpyDeviceToHost, st));
#undef V_TRY
#undef V_CUDA
    cleanup();
    return BIAS_OK;
}

int bias_model_frozen_draws(bias_model *m, const double *uniforms, int32_t *draws) {
    if (!m || !uniforms || !draws) return fail(BIAS_EINVAL, "null argument");
    if (is_dynamic_K(m)) return fail(BIAS_EINVAL, "frozen draws are defined for LDA and BMM");
    if (m->N == 0) return BIAS_OK;
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    double *d_u = nullptr;
    int32_t *d_draw = nullptr;
    BIAS_TRY(dev_alloc(&d_u, (size_t)m->N));
    if (dev_alloc(&d_draw, (size_t)m->N) != BIAS_OK) { cudaFree(d_u); return BIAS_ENOMEM; }
    int rc = BIAS_OK;
    cudaError_t e = cudaMemcpyAsync(d_u, uniforms, (size_t)m->N * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->scratch + 2, 0, sizeof(unsigned long long), st);
    if (e == cudaSuccess) {
        if (m->fast) {
            rc = launch_lda_fast(m, d_u, 1, d_draw);
        } else {
            int32_t *prior = (m->model == BIAS_MODEL_BMM) ? m->n_items : m->group_rows;
            const int32_t *row_of = (m->model == BIAS_MODEL_BMM) ? nullptr : m->group_of;
            rc = launch_generic(m, m->N, nullptr, nullptr, row_of, prior, d_u, 1, 0, nullptr, nullptr, d_draw, kStreamItemDraw);
        }
    }
    if (e == cudaSuccess && rc == BIAS_OK) e = cudaMemcpyAsync(draws, d_draw, (size_t)m->N * 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d_u);
    cudaFree(d_draw);
    if (e != cudaSuccess) return fail(BIAS_ECUDA, "frozen draws: %s", cudaGetErrorString(e));
    return rc;
}

int bias_model_frozen_draws_dev(bias_model *m, const double *uniforms_dev, int32_t *draws_dev) {
    if (!m || !uniforms_dev || !draws_dev) return fail(BIAS_EINVAL, "null argument");
    if (is_dynamic_K(m)) return fail(BIAS_EINVAL, "frozen draws are defined for LDA and BMM");
    if (m->N == 0) return BIAS_OK;
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch + 2, 0, sizeof(unsigned long long), m->ctx->stream));
    if (m->fast) {
        BIAS_TRY(launch_lda_fast(m, uniforms_dev, 1, draws_dev));
    } else {
        int32_t *prior = (m->model == BIAS_MODEL_BMM) ? m->n_items : m->group_rows;
        const int32_t *row_of = (m->model == BIAS_MODEL_BMM) ? nullptr : m->group_of;
        BIAS_TRY(launch_generic(m, m->N, nullptr, nullptr, row_of, prior, uniforms_dev, 1, 0, nullptr, nullptr, draws_dev, kStreamItemDraw));
    }
    BIAS_CUDA_TRY(cudaStreamSynchronize(m->ctx->stream));
    return BIAS_OK;
}

int bias_model_sweep_with_uniforms(bias_model *m, const double *uniforms) {
    if (!m || !uniforms) return fail(BIAS_EINVAL, "null argument");
    if (is_dynamic_K(m)) return fail(BIAS_EINVAL, "sweeps with supplied uniforms are defined for LDA and BMM");
    if (m->world > 1) return fail(BIAS_ESTATE, "sweeps with supplied uniforms are a single-GPU verification entry point");
    if (m->N == 0) return BIAS_OK;
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    double *d_u = nullptr;
    BIAS_TRY(dev_alloc(&d_u, (size_t)m->N));
    cudaError_t e = cudaMemcpyAsync(d_u, uniforms, (size_t)m->N * 8, cudaMemcpyHostToDevice, st);
    int rc = BIAS_OK;
    if (e == cudaSuccess) e = cudaMemsetAsync(m->scratch, 0, 4 * sizeof(unsigned long long), st);
    if (e == cudaSuccess) {
        if (m->fast) {
            rc = launch_lda_fast(m, d_u, 0, nullptr);
        } else {
            int32_t *prior = (m->model == BIAS_MODEL_BMM) ? m->n_items : m->group_rows;
            const int32_t *row_of = (m->model == BIAS_MODEL_BMM) ? nullptr : m->group_of;
            rc = launch_generic(m, m->N, nullptr, nullptr, row_of, prior, d_u, 0, 0, nullptr, nullptr, nullptr, kStreamItemDraw);
        }
    }
    cudaStreamSynchronize(st);
    cudaFree(d_u);
    if (e != cudaSuccess) return fail(BIAS_ECUDA, "sweep with uniforms: %s", cudaGetErrorString(e));
    if (rc != BIAS_OK) return rc;
    m->sweep_index += 1;
    m->stats_pending = true;
    m->mirror16_valid = false;
    return BIAS_OK;
}

int bias_model_narrow_info(bias_model *m, int32_t *narrow_current, int32_t *n_hot, int32_t *hot_cap) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (narrow_current) *narrow_current = (m->n_wk16 && m->narrow_valid) ? 1 : 0;
    if (hot_cap) *hot_cap = m->hot_cap;
    if (n_hot) *n_hot = (m->n_wk16 && m->class_valid) ? m->n_hot_host : 0;
    return BIAS_OK;
}

int bias_model_log_likelihood(bias_model *m, double *sum_log_p, int64_t *n_items) {
    if (!m || !sum_log_p) return fail(BIAS_EINVAL, "null argument");
    if (!m->fast) return fail(BIAS_EINVAL, "the log-likelihood evaluator covers LDA x MultinomialDirichlet x word ids");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch + 2, 0, sizeof(unsigned long long), st));      // document counter
    BIAS_CUDA_TRY(cudaMemsetAsync(m->scratch + 6, 0, sizeof(double), st));
    BIAS_TRY(launch_lda_fast(m, nullptr, 1, nullptr, 1));
    double h = 0.0;
    BIAS_CUDA_TRY(cudaMemcpyAsync(&h, m->scratch + 6, sizeof(double), cudaMemcpyDeviceToHost, st));
    BIAS_CUDA_TRY(cudaStreamSynchronize(st));
    *sum_log_p = h;
    if (n_items) *n_items = m->N;
    return BIAS_OK;
}

int bias_model_heldout_log_likelihood(bias_model *m, const int64_t *heldout_ptr, const int32_t *heldout_word,
                                      double *sum_log_p, int64_t *n_tokens) {
    if (!m || !heldout_ptr || !sum_log_p) return fail(BIAS_EINVAL, "null argument");
    if (!m->fast) return fail(BIAS_EINVAL, "the held-out evaluator covers LDA x MultinomialDirichlet x word ids");
    const int64_t nh = heldout_ptr[m->G];
    if (heldout_ptr[0] != 0 || nh < 0) return fail(BIAS_EINVAL, "heldout_ptr must start at 0");
    for (int64_t j = 0; j < m->G; ++j)
        if (heldout_ptr[j + 1] < heldout_ptr[j]) return fail(BIAS_EINVAL, "heldout_ptr must be non-decreasing");
    if (nh > 0 && !heldout_word) return fail(BIAS_EINVAL, "heldout_word is null");
    for (int64_t t = 0; t < nh; ++t)
        if (heldout_word[t] < 0 || heldout_word[t] >= m->V)
            return fail(BIAS_EINVAL, "held-out word id %d at %lld is outside 0..%lld", heldout_word[t], (long long)t, (long long)m->V - 1);
    *sum_log_p = 0.0;
    if (n_tokens) *n_tokens = nh;
    if (nh == 0) return BIAS_OK;
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    int64_t *d_ptr = nullptr;
    int32_t *d_word = nullptr;
    BIAS_TRY(dev_alloc(&d_ptr, (size_t)m->G + 1));
    int rc = dev_alloc(&d_word, (size_t)nh);
    if (rc != BIAS_OK) { cudaFree(d_ptr); return rc; }
    double h = 0.0;
    cudaError_t e = cudaMemcpyAsync(d_ptr, heldout_ptr, ((size_t)m->G + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_word, heldout_word, (size_t)nh * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->scratch + 2, 0, sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->scratch + 6, 0, sizeof(double), st);
    if (e == cudaSuccess) {
        rc = launch_lda_fast(m, nullptr, 1, nullptr, 1, d_ptr, d_word);
        if (rc == BIAS_OK) e = cudaMemcpyAsync(&h, m->scratch + 6, sizeof(double), cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d_ptr);
    cudaFree(d_word);
    if (rc != BIAS_OK) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(BIAS_ECUDA, "held-out evaluation: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    *sum_log_p = h;
    return BIAS_OK;
}

// -----------------------------------------------------------------------------------------
// multi-GPU exchange
// -----------------------------------------------------------------------------------------
static int ensure_exchange(bias_model *m) {
    if (!m->base_i32) {
        BIAS_TRY(dev_alloc(&m->base_i32, (size_t)m->n_i32));
        BIAS_TRY(dev_alloc(&m->delta_i32, (size_t)m->n_i32 + 4));      // + the overflow flag of the 16-bit exchange
        if (m->n_f64) {
            BIAS_TRY(dev_alloc(&m->base_f64, (size_t)m->n_f64));
            BIAS_TRY(dev_alloc(&m->delta_f64, (size_t)m->n_f64));
        }
    }
    return BIAS_OK;
}

int bias_model_delta_begin(bias_model *m) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_exchange(m));
    BIAS_TRY(ensure_wide(m));
    BIAS_CUDA_TRY(cudaMemcpyAsync(m->base_i32, m->i32_block, (size_t)m->n_i32 * 4, cudaMemcpyDeviceToDevice, m->ctx->stream));
    if (m->n_f64)
        BIAS_CUDA_TRY(cudaMemcpyAsync(m->base_f64, m->f64_block, (size_t)m->n_f64 * 8, cudaMemcpyDeviceToDevice, m->ctx->stream));
    return BIAS_OK;
}

int bias_model_delta_begin_zero(bias_model *m) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_exchange(m));
    BIAS_CUDA_TRY(cudaMemsetAsync(m->base_i32, 0, (size_t)m->n_i32 * 4, m->ctx->stream));
    if (m->n_f64) BIAS_CUDA_TRY(cudaMemsetAsync(m->base_f64, 0, (size_t)m->n_f64 * 8, m->ctx->stream));
    return BIAS_OK;
}

int bias_model_delta_pack(bias_model *m, void **delta_i32_dev, int64_t *n_i32, void **delta_f64_dev, int64_t *n_f64) {
    if (!m || !delta_i32_dev || !n_i32) return fail(BIAS_EINVAL, "null argument");
    if (!m->base_i32) return fail(BIAS_ESTATE, "bias_model_delta_begin has not been called");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_wide(m));
    cudaStream_t st = m->ctx->stream;
    const int64_t n4 = m->n_i32 / 4;
    delta_pack_i32_kernel<<<grid_for(m->ctx, n4, 256), 256, 0, st>>>(m->i32_block, m->base_i32, m->delta_i32, n4);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    if (m->n_f64) {
        delta_pack_f64_kernel<<<grid_for(m->ctx, m->n_f64, 256), 256, 0, st>>>(m->f64_block, m->base_f64, m->delta_f64, m->n_f64);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
    }
    *delta_i32_dev = m->delta_i32;
    *n_i32 = m->n_i32;
    if (delta_f64_dev) *delta_f64_dev = m->delta_f64;
    if (n_f64) *n_f64 = m->n_f64;
    return BIAS_OK;
}

// 16-bit exchange of the word-topic part (LDA x MD x Int): the int16 deltas live in the first half of the delta buffer,
// the n_k / n_items deltas (int32) and the overflow flag behind the word-topic region
int bias_model_delta_pack16(bias_model *m, int32_t world, void **d16_dev, int64_t *n16, void **small_dev, int64_t *n_small) {
    if (!m || !d16_dev || !n16 || !small_dev || !n_small) return fail(BIAS_EINVAL, "null argument");
    if (!m->base_i32) return fail(BIAS_ESTATE, "bias_model_delta_begin has not been called");
    if (m->conj.kind != BIAS_MD || m->n_f64 != 0 || world < 1) return fail(BIAS_EINVAL, "the 16-bit exchange covers MultinomialDirichlet count tables");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_wide(m));
    cudaStream_t st = m->ctx->stream;
    const int64_t n_wk_elems = m->V * (int64_t)m->Kpad, n_small_i = m->n_i32 - n_wk_elems;
    int32_t *small = m->delta_i32 + n_wk_elems, *flag = m->delta_i32 + m->n_i32;
    BIAS_CUDA_TRY(cudaMemsetAsync(flag, 0, 4 * sizeof(int32_t), st));
    if (n_wk_elems > 0) {
        delta_pack16_kernel<<<grid_for(m->ctx, n_wk_elems / 8, 256), 256, 0, st>>>(m->i32_block, m->base_i32,
            reinterpret_cast<int16_t *>(m->delta_i32), n_wk_elems / 8, 32767 / world, flag);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
    }
    delta_pack_i32_kernel<<<grid_for(m->ctx, n_small_i / 4, 256), 256, 0, st>>>(m->i32_block + n_wk_elems, m->base_i32 + n_wk_elems,
                                                                               small, n_small_i / 4);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    *d16_dev = m->delta_i32;
    *n16 = n_wk_elems;
    *small_dev = small;
    *n_small = n_small_i + 4;
    return BIAS_OK;
}

int bias_model_delta_apply16(bias_model *m) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (!m->base_i32) return fail(BIAS_ESTATE, "bias_model_delta_begin has not been called");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_wide(m));
    m->narrow_valid = false;        // the int32 table is written below
    cudaStream_t st = m->ctx->stream;
    const int64_t n_wk_elems = m->V * (int64_t)m->Kpad, n_small_i = m->n_i32 - n_wk_elems;
    if (n_wk_elems > 0) {
        delta_apply16_kernel<<<grid_for(m->ctx, n_wk_elems / 8, 256), 256, 0, st>>>(m->i32_block, m->base_i32,
            reinterpret_cast<const int16_t *>(m->delta_i32), n_wk_elems / 8, m->delta_i32 + m->n_i32);
        m->ctx->launches += 1;
    }
    delta_apply_i32_kernel<<<grid_for(m->ctx, n_small_i / 4, 256), 256, 0, st>>>(m->i32_block + n_wk_elems, m->base_i32 + n_wk_elems,
                                                                                m->delta_i32 + n_wk_elems, n_small_i / 4,
                                                                                m->delta_i32 + m->n_i32);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    return BIAS_OK;
}

int bias_model_delta_apply(bias_model *m) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (!m->base_i32) return fail(BIAS_ESTATE, "bias_model_delta_begin has not been called");
    BIAS_CUDA_TRY(cudaSetDevice(m->ctx->device));
    BIAS_TRY(ensure_wide(m));
    m->narrow_valid = false;        // the int32 table is written below
    cudaStream_t st = m->ctx->stream;
    const int64_t n4 = m->n_i32 / 4;
    delta_apply_i32_kernel<<<grid_for(m->ctx, n4, 256), 256, 0, st>>>(m->i32_block, m->base_i32, m->delta_i32, n4);
    BIAS_CUDA_TRY(cudaGetLastError());
    m->ctx->launches += 1;
    if (m->n_f64) {
        delta_apply_f64_kernel<<<grid_for(m->ctx, m->n_f64, 256), 256, 0, st>>>(m->f64_block, m->base_f64, m->delta_f64, m->n_f64);
        BIAS_CUDA_TRY(cudaGetLastError());
        m->ctx->launches += 1;
    }
    return BIAS_OK;
}

// -----------------------------------------------------------------------------------------
// one-call samplers
// -----------------------------------------------------------------------------------------
static int run_model(bias_ctx *ctx, int32_t model, const bias_conjugate *conj, const bias_data *data, int64_t G,
                     const int64_t *ptr, int32_t *zz, int32_t K, double alpha, bias_hdp_params *hp, int32_t n_sweeps,
                     double *beta_out, bias_sweep_stats *stats) {
    if (n_sweeps < 0) return fail(BIAS_EINVAL, "n_sweeps < 0");
    bias_model *m = nullptr;
    BIAS_TRY(bias_model_create(ctx, model, conj, data, G, ptr, zz, K, alpha, hp, &m));
    int rc = BIAS_OK;
    if (hp && beta_out && beta_out[0] > 0.0) {
        // resume: the caller hands back the my_beta of an earlier call (src/HDP.jl:173-174 resume path)
        double tot = 0.0;
        for (int k = 0; k <= hp->KK; ++k) tot += beta_out[k];
        if (tot > 0.0) {
            for (int k = 0; k <= hp->KK; ++k) m->h_beta[k] = beta_out[k] / tot;
            cudaMemcpyAsync(m->d_beta, m->h_beta.data(), m->h_beta.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
        }
    }
    for (int32_t s = 0; s < n_sweeps && rc == BIAS_OK; ++s) {
        rc = one_sweep(m);
        if (rc == BIAS_OK && stats) {
            rc = fetch_stats(m);
            stats[s] = m->last;
        }
    }
    if (rc == BIAS_OK) rc = bias_model_get_zz(m, zz);
    if (rc == BIAS_OK && hp) rc = bias_model_get_hdp(m, hp, beta_out);
    bias_model_destroy(m);
    return rc;
}

int bias_rcrf_create(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int64_t n_groups,
                     const int64_t *group_ptr, int64_t n_epochs, const int64_t *epoch_ptr, const int32_t *zz,
                     const bias_rcrf_params *params, const double *gg, const double *aa, bias_model **out) {
    if (!params || !gg || !aa || !epoch_ptr || !out) return fail(BIAS_EINVAL, "null argument");
    if (n_epochs < 2) return fail(BIAS_EINVAL, "RCRF needs at least two epochs (src/RCRF.jl has first, middle and last epoch blocks)");
    if (epoch_ptr[0] != 0 || epoch_ptr[n_epochs] != n_groups) return fail(BIAS_EINVAL, "epoch_ptr must start at 0 and end at n_groups");
    for (int64_t t = 0; t < n_epochs; ++t) {
        if (epoch_ptr[t + 1] < epoch_ptr[t]) return fail(BIAS_EINVAL, "epoch_ptr must be non-decreasing");
        if (!(aa[t] > 0.0) || !(gg[t] > 0.0)) return fail(BIAS_EINVAL, "RCRF: aa[%lld] and gg[%lld] must be > 0", (long long)t, (long long)t);
    }
    bias_hdp_params hp{};
    hp.KK = params->KK;
    hp.Kmax = params->Kmax;
    hp.gg = gg[0];
    hp.g1 = params->g1;
    hp.g2 = params->g2;
    hp.aa = aa[0];
    hp.a1 = params->a1;
    hp.a2 = params->a2;
    hp.sample_hyperparam = params->sample_hyperparam;
    hp.n_internals = params->n_internals;
    BIAS_TRY(bias_model_create(ctx, BIAS_MODEL_RCRF, conj, data, n_groups, group_ptr, zz, params->KK, aa[0], &hp, out));
    int rc = rcrf_create(*out, n_epochs, epoch_ptr, gg, aa, params->join_tables);
    if (rc != BIAS_OK) {
        bias_model_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int bias_rcrp_create(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int64_t n_epochs,
                     const int64_t *epoch_ptr, const int32_t *zz, const bias_rcrp_params *params, const double *aa,
                     bias_model **out) {
    if (!params || !aa || !out) return fail(BIAS_EINVAL, "null argument");
    if (n_epochs < 2) return fail(BIAS_EINVAL, "RCRP needs at least two epochs (src/RCRP.jl:174-398 has first and last epoch blocks)");
    for (int64_t t = 0; t < n_epochs; ++t)
        if (!(aa[t] > 0.0)) return fail(BIAS_EINVAL, "RCRP: aa[%lld] must be > 0", (long long)t);
    bias_hdp_params hp{};
    hp.KK = params->KK;
    hp.Kmax = params->Kmax;
    hp.gg = 1.0;
    hp.aa = aa[0];
    hp.a1 = params->a1;
    hp.a2 = params->a2;
    hp.sample_hyperparam = params->sample_hyperparam;
    hp.n_internals = params->n_internals;
    BIAS_TRY(bias_model_create(ctx, BIAS_MODEL_RCRP, conj, data, n_epochs, epoch_ptr, zz, params->KK, aa[0], &hp, out));
    int rc = rcrp_create(*out, params, aa);
    if (rc != BIAS_OK) {
        bias_model_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int bias_model_get_rcrp(bias_model *m, bias_rcrp_params *params, double *aa) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    if (m->model != BIAS_MODEL_RCRP) return fail(BIAS_EINVAL, "not an RCRP model");
    if (params) {
        params->KK = m->hdp.KK;
        params->Kmax = m->hdp.Kmax;
        params->a1 = m->hdp.a1;
        params->a2 = m->hdp.a2;
        params->sample_hyperparam = m->hdp.sample_hyperparam;
        params->n_internals = m->hdp.n_internals;
    }
    if (aa) std::copy(m->h_aa_t.begin(), m->h_aa_t.end(), aa);
    return BIAS_OK;
}

int bias_model_set_sample_hyperparam(bias_model *m, int32_t on) {
    if (!m) return fail(BIAS_EINVAL, "null model");
    m->hdp.sample_hyperparam = on ? 1 : 0;
    return BIAS_OK;
}

int bias_rcrp_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int64_t n_epochs,
                  const int64_t *epoch_ptr, int32_t *zz, bias_rcrp_params *params, double *aa, int32_t n_burnins,
                  int32_t n_sweeps, bias_sweep_stats *stats) {
    if (!params || !aa) return fail(BIAS_EINVAL, "null argument");
    if (n_sweeps < 0) return fail(BIAS_EINVAL, "n_sweeps < 0");
    bias_model *m = nullptr;
    BIAS_TRY(bias_rcrp_create(ctx, conj, data, n_epochs, epoch_ptr, zz, params, aa, &m));
    int rc = BIAS_OK;
    for (int32_t s = 0; s < n_sweeps && rc == BIAS_OK; ++s) {
        if (s >= n_burnins) m->hdp.sample_hyperparam = 0;        // src/RCRP.jl:166-168
        rc = one_sweep(m);
        if (rc == BIAS_OK && stats) {
            rc = fetch_stats(m);
            stats[s] = m->last;
        }
    }
    if (rc == BIAS_OK) rc = bias_model_get_zz(m, zz);
    if (rc == BIAS_OK) rc = bias_model_get_rcrp(m, params, aa);
    bias_model_destroy(m);
    return rc;
}

int bias_lda_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int64_t n_groups,
                 const int64_t *group_ptr, int32_t *zz, int32_t K, double alpha, int32_t n_sweeps,
                 bias_sweep_stats *stats) {
    return run_model(ctx, BIAS_MODEL_LDA, conj, data, n_groups, group_ptr, zz, K, alpha, nullptr, n_sweeps, nullptr, stats);
}

int bias_bmm_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int32_t *zz, int32_t K,
                 double alpha, int32_t n_sweeps, bias_sweep_stats *stats) {
    return run_model(ctx, BIAS_MODEL_BMM, conj, data, 0, nullptr, zz, K, alpha, nullptr, n_sweeps, nullptr, stats);
}

int bias_hdp_run(bias_ctx *ctx, const bias_conjugate *conj, const bias_data *data, int64_t n_groups,
                 const int64_t *group_ptr, int32_t *zz, bias_hdp_params *params, int32_t n_sweeps,
                 double *beta_out, bias_sweep_stats *stats) {
    if (!params) return fail(BIAS_EINVAL, "null HDP parameters");
    return run_model(ctx, BIAS_MODEL_HDP, conj, data, n_groups, group_ptr, zz, params->KK, params->aa, params, n_sweeps,
                     beta_out, stats);
}

}  // extern "C"
