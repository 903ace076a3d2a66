// common.cuh — handle layout, error plumbing and small device helpers for libbossgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/bossgpu.h"

namespace boss {

// ------------------------------------------------------------------------------------------------
// error plumbing: no exception crosses the C ABI
// ------------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define BOSS_CUDA(expr)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::boss::fail(BOSSGPU_ECUDA, "%s failed: %s (%s:%d)", #expr,                     \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                       \
    } while (0)

#define BOSS_KERNEL_CHECK() BOSS_CUDA(cudaGetLastError())

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
constexpr int      BIN       = BOSSGPU_BIN;
constexpr int      BUCKET    = BOSSGPU_BUCKET;
constexpr int      FREEZE    = BOSSGPU_FREEZE;
constexpr int      NPAT      = BOSSGPU_N_PATTERNS;
constexpr int      HBINS     = BOSSGPU_HIST_BINS;
constexpr int      NSTEPS    = BOSSGPU_N_STEPS;
constexpr int      TILE      = 2000;          // sites per CTA in the score+bin pass: 20 bins, 1/10 bucket
constexpr int64_t  SITE_ALIGN = 256;          // segment starts on the padded site axis
constexpr double   TINY      = 2.2250738585072014e-308;   // np.finfo(float).tiny (sequences.py:430)

// per-segment geometry, device-visible
struct SegDev {
    int32_t contig;         // global index in contigs_filt order
    int32_t is_tail;        // segment contains the contig's last site
    int64_t contig_len;
    int64_t start;          // first site within the contig (multiple of BUCKET)
    int64_t len;            // sites
    int64_t site_off;       // offset on the padded site axis
    int64_t n_bins;         // bins owned (tail: includes the L//100 + 1 - th bin)
    int64_t ds_off;         // index of local bin 0 in the scores_ds arrays (after the left halo)
    int32_t halo_l, halo_r; // halo capacity available in scores_ds on each side
    int64_t row_off;        // merged-row index of local bin 0, relative to the shard's first merged row
    int64_t n_srows;        // strategy rows owned
    int64_t srow_off;       // strategy-row offset within the shard
    int64_t n_full_buckets; // complete 20 kb buckets in this segment
    int64_t n_sw;           // bucket switch entries owned (tail: n_full_buckets + 1)
    int64_t sw_off;         // offset into bucket arrays
    int64_t tile_off;       // first tile id of this segment
    int64_t n_tiles;
};

// device-side scalars shared between kernels of one update
struct UpdateDev {
    unsigned long long norm_bits;    // max benefit as raw double bits (non-negative => order preserving)
    int32_t  switched_on;
    int32_t  strat_size;
    double   threshold;
    double   normaliser;
    double   ubar0;
    double   fhat_sum;
    double   fhat_scale;             // 1 / fhat_sum (or 1 when the sum is 0)
    unsigned long long n_nonzero;
    unsigned long long n_dropout;
    unsigned long long n_accept[2];
    unsigned long long fsum_hi, fsum_lo;   // exact limbs of sum(fhat_exp)
    unsigned long long mirror_bytes;       // bytes the distribution kernel wrote into the host mirror
    int32_t  e_thr;                  // exponent bin of the threshold: threshold = 2^-e_thr * normaliser
    int32_t  use_codes;              // the one-byte bin codes k_hist left behind decide the masks (else: the benefits)
    int32_t  empty;                  // 1: every benefit is zero (upstream: np.max of an empty array raises); 2: no time_cost (Q14)
    int32_t  fabric_err;             // BOSSGPU_EPEER: a peer shard did not show up at an exchange step (fabric.cuh)
    int32_t  error;                  // BOSSGPU_E* raised on the device; survives the per-update reset
    int32_t  pad_;
};

struct Timers {
    cudaEvent_t ev[16];
};

}  // namespace boss

struct bossgpu_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int nb = 1;
    int len_g = 5;
    int n_seg = 0;
    int n_contigs_total = 0;
    int halo_bins = 0;
    int64_t n_sites_total = 0;
    int64_t n_windows_total = 0;
    // derived global geometry
    int64_t M_rows = 0;        // sum over contigs of L//100 + 1   (rows of the merged benefit array)
    int64_t target_rows = 0;   // n_sites_total // 100             (core.py:179)
    int64_t Tf_rows = 0;       // sum(L) // 100                    (readstartdist.py:29)
    int64_t R0 = 0;            // first merged row of this shard
    int64_t D0 = 0;            // first strategy row (distribution axis) of this shard
    int fhat_shift = 50;       // exact F-hat sums: terms are fhat * 2^fhat_shift (strategy.cuh)
    int ubar_shift = 61;       // exact ubar0 sum: terms are fhat * benefit * 2^(ubar_shift - exponent of the normaliser)
    // shard totals
    int64_t P = 0;             // padded sites
    int64_t n_tiles = 0;
    int64_t ds_len = 0;        // scores_ds slots incl. halos
    int64_t n_rows = 0;        // merged rows owned
    int64_t n_srows = 0;       // strategy rows owned
    int64_t n_sw = 0;          // bucket switch entries owned
    std::vector<boss::SegDev> segs;
    std::vector<int64_t> contig_len_all;
    // device memory
    boss::SegDev* d_segs = nullptr;
    int64_t*  d_tile_start = nullptr;   // [n_seg+1]
    int64_t*  d_row_start = nullptr;    // [n_seg+1]
    int64_t*  d_srow_start = nullptr;   // [n_seg+1]
    int64_t*  d_sm_tile_start = nullptr; // [n_seg+1] first smoothing tile of each segment
    int64_t   n_sm_tiles = 0;
    int64_t*  d_contig_len = nullptr;   // [n_contigs_total]
    void*     d_tiles = nullptr;        // [n_tiles] TileDesc (score_pass.cuh)
    uint8_t*  d_ref = nullptr;          // [P]
    uint16_t* d_cov = nullptr;          // [nb][5][P]
    uint32_t* d_rowflag = nullptr;      // [P] (nb > 1 only)
    double*   d_table = nullptr;        // [NPAT][4] scores
    double*   d_etable = nullptr;       // [NPAT][4] entropies
    double*   d_phi = nullptr;          // [5][len_g]
    double*   d_priors = nullptr;       // [4][len_g]
    double*   d_phi_pow = nullptr;      // [5][len_g][30]
    double    score0 = 0, ent0 = 0;
    double    row0_true[8] = {0};       // table / etable values of the all-zero pattern (row 0 is patched when nb == 1)
    unsigned long long* d_cov_total = nullptr;   // [n_contigs_total]
    int32_t*  d_drop_thr = nullptr;              // [n_contigs_total]  -1 = dropout rule inactive
    double*   d_ds = nullptr;                    // [nb][ds_len]
    double2*  d_benefit = nullptr;               // [nb][n_rows]  (.x forward, .y reverse)
    uint8_t*  d_codes = nullptr;                 // [nb][n_rows][2] exponent-bin code of every benefit entry (k_hist -> k_distribute)
    double2*  d_smu = nullptr;                   // debug
    double2*  d_expected = nullptr;              // debug
    unsigned long long* d_bucket_sum = nullptr;  // [n_sw][nb]
    uint8_t*  d_bucket_sw = nullptr;             // [n_sw][nb]
    uint8_t*  h_bucket_sw = nullptr;             // pinned image, refreshed at the end of every update
    double*   d_fhat_w = nullptr;                // [n_windows_total][2]
    unsigned long long* d_hist = nullptr;        // [3*HBINS + 4]
    uint8_t*  d_strat = nullptr;                 // [n_srows][2][nb]; d_strat_alloc + strat_shift
    uint8_t*  d_strat_alloc = nullptr;
    int       strat_shift = 0;                   // keeps d_strat congruent to the host mirror mod 16
    uint8_t*  h_strat = nullptr;                 // host mirror of d_strat (mapped pinned memory), kept current by k_distribute
    uint8_t*  h_strat_dev = nullptr;             // the same memory as the device sees it
    uint8_t*  h_strat_own = nullptr;             // the library's own allocation (NULL once an external mirror is set)
    void*     reg_base = nullptr;                // page-aligned range registered for an external mirror
    unsigned long long* d_seg_accept = nullptr;  // [n_seg][2] accepted entries per segment and strand
    unsigned long long* h_seg_accept = nullptr;  // pinned
    unsigned long long* d_rs_counts = nullptr;   // [n_windows_total][2] read-start counts (readstartdist.py:26)
    boss::UpdateDev* d_upd = nullptr;
    boss::UpdateDev* h_upd = nullptr;            // pinned
    // staging for ingest
    void*  stage_h = nullptr;  size_t stage_h_bytes = 0;   // pinned
    void*  stage_d = nullptr;  size_t stage_d_bytes = 0;
    void*  scratch_d = nullptr; size_t scratch_d_bytes = 0;
    void*  ingest_d = nullptr;  size_t ingest_d_bytes = 0;     // op records of the batch being ingested (scatter.cuh)
    int32_t* d_ingest_err = nullptr;
    int32_t* h_ingest_err = nullptr;
    // timing
    cudaEvent_t ev[2 * BOSSGPU_N_TIMERS];
    float ms[BOSSGPU_N_TIMERS] = {0};
    bool  ev_valid[BOSSGPU_N_TIMERS] = {false};
    int64_t launches = 0;
    int64_t last_ingest_h2d = 0;                // bytes the last text ingest moved to the device
    int phase_done = -1;
    int n_sm = 148;
    bool score_kernel_ldg = false;
    int score_stages = 2, score_ctas_per_sm = 3;
    // multi-shard exchange state (bossgpu_set_shards)
    int n_shards = 1, shard_index = 0;
    std::vector<int64_t> shard_row_start;
    int64_t mask_stride = 0;
    uint8_t* d_mask_all = nullptr;           // [n_shards][mask_stride] packed masks of every shard's merged rows
    const uint8_t** d_mask_ptrs = nullptr;   // [n_shards] where each shard's packed mask is read from (phase API: slices
                                             // of d_mask_all; fabric: the mask inside each peer's exchange block)
    // peer-memory fabric (fabric.cuh): this shard's exchange block and the peers' blocks as this device sees them
    char*    d_fabric = nullptr;
    size_t   fabric_bytes = 0;
    char**   d_peer_ptrs = nullptr;          // [n_shards]
    const uint8_t** d_fab_mask_ptrs = nullptr; // [n_shards] each peer block's packed mask
    bool     fabric_attached = false;
    unsigned fabric_epoch = 0;
    unsigned long long fabric_timeout_ns = 2000000000ull;
    bool     fused_open = false;             // bossgpu_update_fused_begin without its _end
    int64_t* d_shard_row_start = nullptr;    // [n_shards+1]
    double*  d_halo = nullptr;               // [send L | send R | recv L | recv R] x halo_bins x nb
    bool have_fhat = false;
    bool sticky_on = false;                  // an earlier update saw a bucket on (switches never go off again)
    bool debug_bufs = false;
    // split score/bin pass (bossgpu_prescore): tiles the coming batch does not touch are scored on `stream2` while the
    // host is still packing the batch; the update then only scores the touched tiles
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_pre_thr = nullptr, ev_pre_done = nullptr, ev_main = nullptr;
    int  spec_state = 0;                     // 1: an early pass over every tile is in flight / done on stream2
    int  prescore_state = 0;                 // batch announcement: 0 none, 1 announced, 2 confirmed by the ingest, -1 differs
    int32_t*  d_drop_thr_spec = nullptr;     // [n_contigs_total] thresholds the early pass used
    uint32_t* d_tile_cov = nullptr;          // [n_tiles][nb] per-tile depth totals (score_pass.cuh)
    uint32_t* d_tile_drop = nullptr;         // [n_tiles]
    bool multi_fused = false;                // barcodes: k_score_bin_multi takes a tile through every barcode (no row-summary pass)
    int multi_ctas = 2;                  // resident CTAs per SM of k_score_bin_multi (shared memory per CTA grows with the barcodes)
    bool prescore_ok = false;                // geometry allows it (one segment per contig, no barcodes, staged kernel)
    int64_t pre_n_reads = 0;
    uint64_t pre_hash = 0;
    std::vector<unsigned long long> pre_cov_add;
    int32_t*  d_seg_of_contig = nullptr;     // [n_contigs_total] local segment of a contig, -1 if none
    uint32_t* d_touched = nullptr;           // bitmap over tile ids
    int32_t*  d_touched_list = nullptr;      // [n_tiles]
    unsigned long long* d_pre_misc = nullptr; // [0] list length (low 32 bits used)
    void* pre_stage_h = nullptr; void* pre_stage_d = nullptr; size_t pre_stage_bytes = 0;
    boss::UpdateDev last;
};
