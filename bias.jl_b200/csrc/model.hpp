// model.hpp — host-side state of libbias_cuda (context and resident sampler state).
#pragma once
#include <vector>
#include <random>
#include "common.cuh"

constexpr int kMaxDevices = 64;     // size of the per-device caches of launch configurations
constexpr int kMaxPeers = 8;        // ranks of a sharded run (the GPUs of one NVSwitch domain)

struct bias_ctx {
    int device = 0;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    int64_t launches = 0;
    // device-resident normalised log-Stirling table (packed lower triangle), built on first use
    double *stirling = nullptr;
    int32_t stirling_rows = 0;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // models created on this context; a context destroyed while models are alive is released by the last of them
    int32_t live_models = 0;
    bool closed = false;
};

struct bias_model {
    bias_ctx *ctx = nullptr;
    int32_t model = 0;
    bias_conjugate conj{};
    int64_t N = 0, G = 0, nnz = 0, V = 0, max_group_len = 1;
    int64_t N_cap = 0, G_cap = 0, nnz_cap = 0;      // allocated capacities (bias_model_reload re-feeds up to these)
    int32_t K = 0, Kpad = 0;
    double alpha = 0.0;
    bool fast = false;              // LDA x MD x Int on the fp32 warp-per-document kernel

    // observations
    int32_t *word = nullptr;
    double *x = nullptr;
    int64_t *sent_ptr = nullptr;
    int32_t *sent_word = nullptr, *sent_count = nullptr;
    int64_t *group_ptr = nullptr;
    int32_t *group_of = nullptr;
    int32_t *z = nullptr;

    // shared component statistics: one int32 block [n_wk | n_k | n_items] and one f64 block [sumx]
    int32_t *i32_block = nullptr;
    int64_t n_i32 = 0;              // padded to a multiple of 4
    int32_t *n_wk = nullptr, *n_k = nullptr, *n_items = nullptr;
    double *f64_block = nullptr;
    int64_t n_f64 = 0;
    double *sumx = nullptr;
    int32_t *group_rows = nullptr;  // [G][Kpad] n_jk for the generic LDA/HDP path
    // fast LDA path (K <= 1024): the narrow tables (lda_half16.cuh) hold the word-topic counts while sweeps run and the
    // int32 table n_wk above is a view refreshed on demand.  Exactly one of the two may be stale at any time.
    uint16_t *n_wk16 = nullptr;     // [V + 2 * hot_cap][Kpad]: 16-bit rows of all words, then the int32 rows of the hot words
    int32_t hot_cap = 0;            // hot rows allocated: no corpus of the model's capacity can have more words above 65,535
    int32_t *word_hot = nullptr;    // [V] index of the word's hot row, or -1
    int32_t *hot_word = nullptr;    // [hot_cap] the reverse map
    int32_t *hot_info = nullptr;    // [0] number of hot words, [1] set when they did not fit (cannot happen by construction)
    int32_t *word_freq = nullptr;   // [V] occurrences of each word in the local shard
    int32_t *word_freq_g = nullptr; // [V] ... in the whole corpus (all ranks of a sharded run); == word_freq on one GPU
    int32_t *nk_rep = nullptr;      // [kNkRep][Kpad] replicas collecting a sweep's changes of n_k (half-warp kernels)
    bool wide_valid = true;         // n_wk holds the current counts
    bool narrow_valid = false;      // the narrow tables hold the current counts
    bool class_valid = false;       // word_hot matches the loaded corpus
    int32_t n_hot_host = 0;         // number of hot words of the loaded corpus (read back after every classification)
    int32_t hot_rep = 1;            // copies of the hot rows in the narrow tables (kHotRep, or 1 when hot_cap is huge)
    bool hot_rep_valid = false;     // the copies equal the first one

    // sharded run (comm.cu): the narrow tables sit at the start of one allocation that the other ranks map
    unsigned char *share_base = nullptr;
    size_t share_bytes = 0, off_nk_pub = 0, off_freq_pub = 0, off_flags = 0;
    int32_t rank = 0, world = 1, world_pending = 0;
    bool grouped = false;                                 // the ranks are models of this process (bias_group): events order them
    unsigned char *peer_base[kMaxPeers] = {};             // every rank's block as addressed from this device
    bool peer_ipc[kMaxPeers] = {};                        // mapped with cudaIpcOpenMemHandle
    uint16_t *base16 = nullptr;                           // [V / world + 1][Kpad] agreed counts of the rows this rank owns
    int32_t *base_hot = nullptr, *nk_base = nullptr, *comm_err = nullptr;
    uint32_t barrier_epoch = 0;
    cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;         // around the most recent exchange kernel
    uint64_t seed_mix = 0;                                // folded into the Philox key: the ranks of a sharded run draw different uniforms
    uint16_t *stage16 = nullptr;    // [3][N_cap rounded up to 8]: uint16 word ids and assignments as uploaded by reload_u16, and the
                                    // uint16 mirror of z that every sweep call leaves behind for get_zz_u16; allocated on first use
    bool mirror16_valid = false;    // the mirror holds the current z

    // multi-GPU exchange buffers (allocated on first use)
    int32_t *base_i32 = nullptr, *delta_i32 = nullptr;
    double *base_f64 = nullptr, *delta_f64 = nullptr;

    // per-sweep scratch: [0] diag (double) [1] moved (u64) [2] doc counter (u64) [3] pending count (u32)
    unsigned long long *scratch = nullptr;
    unsigned long long *h_scratch = nullptr;   // pinned mirror
    uint32_t sweep_index = 0;
    bias_sweep_stats last{};
    bool stats_pending = false;
    int32_t last_launches = 0;
    bool timing_valid = false;

    // HDP
    bias_hdp_params hdp{};
    std::vector<double> h_beta;     // [Kmax+1]
    double *d_beta = nullptr;
    int32_t *pending = nullptr;     // [N] items that drew the new-component slot
    int32_t *d_KK = nullptr;        // device copy of KK used by the serial birth pass
    int64_t *d_tables = nullptr;    // [Kpad] sum_j M_jk, [Kpad] = total
    double *d_hyper = nullptr;      // [0] sum log w, [1] sum s
    int32_t *d_remap = nullptr;     // [Kpad] slot compaction map
    std::mt19937_64 host_rng;
    int32_t *pending_base = nullptr; // start of the two-half pending queue buffer (pending may point at either half)

    // CRF (tables and dishes of the Chinese restaurant franchise, csrc/crf.cuh)
    int32_t *crf_tji = nullptr, *crf_njt = nullptr, *crf_kjt = nullptr, *crf_n_tbls = nullptr, *crf_mm = nullptr;
    int32_t *crf_tmp = nullptr;             // [N] scratch of the table-joining pass
    // RCRF: restaurants of all epochs are the groups; first restaurant of each epoch, per-epoch gg (aa in h_aa_t)
    std::vector<int64_t> h_epoch_gptr;
    std::vector<double> h_gg_t;
    double *d_gg_t = nullptr;       // [T] device copy (the per-epoch hyper-parameter draws run on the device)
    int32_t crf_join = 1;

    // RCRP (epochs are the groups)
    std::vector<double> h_aa_t;     // [T] per-epoch concentration
    double *d_aa_t = nullptr;
    std::vector<int64_t> h_group_ptr;
    int64_t work_base = 0;          // first item of the next generic launch (per-epoch passes)
};

namespace bias {
int ensure_stirling(bias_ctx *ctx, int32_t rows);
int hdp_create(bias_model *m, const bias_hdp_params *hp);
int hdp_sweep(bias_model *m);
int hdp_destroy(bias_model *m);
int crf_create(bias_model *m);
int crf_sweep(bias_model *m, const double *ext_uniforms_dev);
int crf_destroy(bias_model *m);
int rcrf_create(bias_model *m, int64_t n_epochs, const int64_t *epoch_ptr, const double *gg, const double *aa, int join_tables);
int rcrf_sweep(bias_model *m, const double *ext_uniforms_dev);
int rcrp_create(bias_model *m, const bias_rcrp_params *rp, const double *aa);
int rcrp_sweep(bias_model *m);
bool is_dynamic_K(const bias_model *m);
int launch_generic(bias_model *m, int64_t n_work, const int64_t *item_idx, const int32_t *work_idx32,
                   const int32_t *row_of, int32_t *prior_rows, const double *ext_uniforms, int frozen,
                   int faithful, double *raw_out, double *prob_out, int32_t *draw_out, uint32_t rng_stream,
                   int row_by_work = 0);
int rebuild_counts(bias_model *m);
int narrow_replicate(bias_model *m);  // the copies of the hot rows equal the canonical one on return
int ensure_wide(bias_model *m);       // n_wk (int32) is current on return
int ensure_narrow(bias_model *m);     // the narrow tables are current on return
int model_local_sweep(bias_model *m); // one sweep over this model's items, no exchange
int model_mirror16(bias_model *m);
// comm.cu
int32_t narrow_hot_cap(int64_t V, int64_t n_items_cap);
int narrow_alloc(bias_model *m, int32_t hot_cap);
void narrow_free(bias_model *m);
int narrow_classify(bias_model *m);
int narrow_build(bias_model *m);
int narrow_fold(bias_model *m);
int comm_global_word_freq(bias_model *m);
int comm_prepare_flags(bias_model *m);
int comm_exchange_flags(bias_model *m);
int comm_check(bias_model *m);
}  // namespace bias
