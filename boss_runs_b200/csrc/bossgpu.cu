// bossgpu.cu — C ABI of libbossgpu.so (see include/bossgpu.h for the contract and the reference
// file:line each entry point replaces). Host-side orchestration only; kernels live in the .cuh files.
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <thread>
#include <atomic>
#include <chrono>

#include "common.cuh"
#include "table.cuh"
#include "scatter.cuh"
#include "score_pass.cuh"
#include "strategy.cuh"
#include "fabric.cuh"
#include "synth.cuh"
#include "aeons.cuh"
#include "tokenizer.h"
#include "workerpool.h"

namespace boss {
thread_local std::string g_last_error;
}
using namespace boss;

#define H_CHECK(h)                                                                 \
    do {                                                                           \
        if (!(h)) return fail(BOSSGPU_EINVAL, "null handle");                      \
        BOSS_CUDA(cudaSetDevice((h)->device));                                     \
    } while (0)

template <typename T>
static int dev_alloc(T** p, size_t n, bool zero = true) {
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) return fail(BOSSGPU_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    if (zero) BOSS_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
    return 0;
}
#define TRY(x) do { int _r = (x); if (_r != 0) return _r; } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

extern "C" int bossgpu_abi_version(void) { return BOSSGPU_ABI_VERSION; }
extern "C" const char* bossgpu_last_error(void) { return g_last_error.c_str(); }
extern "C" int bossgpu_device_count(int* n) {
    if (!n) return fail(BOSSGPU_EINVAL, "null out pointer");
    BOSS_CUDA(cudaGetDeviceCount(n));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// scratch / staging buffers (grow on demand)
// ------------------------------------------------------------------------------------------------
static int ensure_scratch(bossgpu_handle* h, size_t bytes) {
    if (bytes <= h->scratch_d_bytes) return 0;
    if (h->scratch_d) cudaFree(h->scratch_d);
    h->scratch_d = nullptr; h->scratch_d_bytes = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&h->scratch_d, want);
    if (e != cudaSuccess) return fail(BOSSGPU_ENOMEM, "scratch cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    h->scratch_d_bytes = want;
    return 0;
}
static int ensure_stage(bossgpu_handle* h, size_t bytes) {
    if (bytes > h->stage_d_bytes) {
        if (h->stage_d) cudaFree(h->stage_d);
        h->stage_d = nullptr; h->stage_d_bytes = 0;
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&h->stage_d, want);
        if (e != cudaSuccess) return fail(BOSSGPU_ENOMEM, "stage cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        h->stage_d_bytes = want;
    }
    if (bytes > h->stage_h_bytes) {
        if (h->stage_h) cudaFreeHost(h->stage_h);
        h->stage_h = nullptr; h->stage_h_bytes = 0;
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMallocHost(&h->stage_h, want);
        if (e != cudaSuccess) return fail(BOSSGPU_ENOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        h->stage_h_bytes = want;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
extern "C" int bossgpu_create(const bossgpu_config* cfg, bossgpu_handle** out) {
    if (!cfg || !out) return fail(BOSSGPU_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->abi_version != BOSSGPU_ABI_VERSION)
        return fail(BOSSGPU_EINVAL, "ABI version mismatch: caller %d, library %d", cfg->abi_version, BOSSGPU_ABI_VERSION);
    if (cfg->n_segments <= 0 || cfg->n_barcodes <= 0 || !cfg->segments || !cfg->ref_codes || !cfg->contig_len_all ||
        cfg->n_contigs_total <= 0)
        return fail(BOSSGPU_EINVAL, "empty geometry");
    if (cfg->len_g != 5 && cfg->len_g != 15) return fail(BOSSGPU_EINVAL, "len_g must be 5 (haploid) or 15 (diploid)");
    if (!cfg->phi || !cfg->priors || !cfg->phi_pow) return fail(BOSSGPU_EINVAL, "missing scoring constants");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev <= 0)
        return fail(BOSSGPU_ECUDA, "no usable CUDA device (%s); libbossgpu has no CPU fallback",
                    e0 == cudaSuccess ? "device count is 0" : cudaGetErrorString(e0));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(BOSSGPU_EINVAL, "device %d out of range [0,%d)", cfg->device, ndev);
    BOSS_CUDA(cudaSetDevice(cfg->device));

    bossgpu_handle* h = new bossgpu_handle();
    h->device = cfg->device;
    h->stream = (cudaStream_t)cfg->stream;
    h->nb = cfg->n_barcodes;
    h->len_g = cfg->len_g;
    h->n_seg = cfg->n_segments;
    h->n_contigs_total = cfg->n_contigs_total;
    h->halo_bins = cfg->halo_bins;
    h->n_sites_total = cfg->n_sites_total;
    h->n_windows_total = cfg->n_windows_total;
    h->score0 = cfg->score0_contig;
    h->ent0 = cfg->entropy0_contig;
    h->contig_len_all.assign(cfg->contig_len_all, cfg->contig_len_all + cfg->n_contigs_total);
    // F-hat terms enter the exact sums as fhat * 2^50 (strategy.cuh, to_limbs_small: one term stays below 2^51 since the
    // normalised F-hat is <= 1; all of them together stay below 2^63: sum(fhat) = 1 per barcode plus tail copies)
    h->fhat_shift = 50;
    {   // ubar0 = sum(fhat * benefit): each term < fhat * 2^ushift once scaled to the normaliser's binade, all of them
        // together < n_barcodes * 2^ushift, kept below 2^62
        int lg = 0;
        while ((1 << lg) < h->nb) ++lg;
        h->ubar_shift = 61 - lg;
    }

    // ---- global axes --------------------------------------------------------------------------
    std::vector<int64_t> o_row(h->n_contigs_total + 1, 0), o_srow(h->n_contigs_total + 1, 0);
    int64_t sumL = 0;
    for (int k = 0; k < h->n_contigs_total; ++k) {
        int64_t L = h->contig_len_all[k];
        if (L < BUCKET) { delete h; return fail(BOSSGPU_EINVAL, "contig %d shorter than one bucket (%lld)", k, (long long)L); }
        o_row[k + 1] = o_row[k] + L / BIN + 1;
        o_srow[k + 1] = o_srow[k] + L / BIN;
        sumL += L;
    }
    h->M_rows = o_row[h->n_contigs_total];
    h->target_rows = h->n_sites_total / BIN;
    h->Tf_rows = sumL / BIN;
    if (h->target_rows < o_srow[h->n_contigs_total] || h->target_rows - h->M_rows > h->M_rows) {
        delete h;
        return fail(BOSSGPU_EINVAL, "n_sites_total inconsistent with contig lengths");
    }

    // ---- segments -----------------------------------------------------------------------------
    h->segs.resize(h->n_seg);
    int64_t P = 0, tiles = 0, ds = 0, rows = 0, srows = 0, sw = 0;
    for (int s = 0; s < h->n_seg; ++s) {
        const bossgpu_segment& in = cfg->segments[s];
        SegDev& S = h->segs[s];
        if (in.contig < 0 || in.contig >= h->n_contigs_total || in.contig_len != h->contig_len_all[in.contig] ||
            in.start < 0 || in.len <= 0 || in.start % BUCKET != 0 || in.start + in.len > in.contig_len ||
            (in.start + in.len != in.contig_len && (in.start + in.len) % BUCKET != 0)) {
            delete h;
            return fail(BOSSGPU_EINVAL, "segment %d is not a bucket-aligned range of its contig", s);
        }
        if (s > 0) {
            const SegDev& Pv = h->segs[s - 1];
            bool contiguous = (in.contig == Pv.contig && in.start == Pv.start + Pv.len) ||
                              (in.contig == Pv.contig + 1 && in.start == 0 && Pv.is_tail);
            if (!contiguous) { delete h; return fail(BOSSGPU_EINVAL, "segments must be a contiguous range of the genome axis"); }
        }
        S.contig = in.contig;
        S.contig_len = in.contig_len;
        S.start = in.start;
        S.len = in.len;
        S.is_tail = (in.start + in.len == in.contig_len) ? 1 : 0;
        S.site_off = P;
        P += round_up(S.len, SITE_ALIGN);
        S.n_bins = S.is_tail ? (S.contig_len / BIN + 1 - S.start / BIN) : S.len / BIN;
        S.halo_l = (S.start > 0) ? h->halo_bins : 0;
        S.halo_r = (!S.is_tail) ? h->halo_bins : 0;
        S.ds_off = ds + S.halo_l;
        ds += S.halo_l + S.n_bins + S.halo_r;
        S.row_off = rows;
        rows += S.n_bins;
        S.n_srows = S.is_tail ? (S.contig_len / BIN - S.start / BIN) : S.len / BIN;
        S.srow_off = srows;
        srows += S.n_srows;
        S.n_full_buckets = S.is_tail ? (S.contig_len / BUCKET - S.start / BUCKET) : S.len / BUCKET;
        S.n_sw = S.n_full_buckets + (S.is_tail ? 1 : 0);
        if (S.is_tail && S.n_full_buckets < 1) {
            delete h;
            return fail(BOSSGPU_EINVAL, "tail segment %d must hold at least one complete bucket", s);
        }
        S.sw_off = sw;
        sw += S.n_sw;
        S.tile_off = tiles;
        S.n_tiles = ceil_div(S.n_bins * BIN, TILE);
        tiles += S.n_tiles;
    }
    h->P = P; h->n_tiles = tiles; h->ds_len = ds; h->n_rows = rows; h->n_srows = srows; h->n_sw = sw;
    h->R0 = o_row[h->segs[0].contig] + h->segs[0].start / BIN;
    h->D0 = o_srow[h->segs[0].contig] + h->segs[0].start / BIN;

    // ---- device memory ------------------------------------------------------------------------
    int rc = 0;
    auto A = [&](int r) { if (rc == 0) rc = r; };
    A(dev_alloc(&h->d_segs, h->n_seg));
    A(dev_alloc(&h->d_tile_start, h->n_seg + 1));
    A(dev_alloc(&h->d_row_start, h->n_seg + 1));
    A(dev_alloc(&h->d_srow_start, h->n_seg + 1));
    A(dev_alloc(&h->d_sm_tile_start, h->n_seg + 1));
    A(dev_alloc(&h->d_contig_len, h->n_contigs_total));
    A(dev_alloc(&h->d_ref, (size_t)(P + SBT_TAIL_PAD)));
    A(dev_alloc(&h->d_cov, (size_t)h->nb * 5 * P + SBT_TAIL_PAD));
    if (h->nb > 1) A(dev_alloc(&h->d_rowflag, (size_t)(P + SBT_TAIL_PAD)));
    A(dev_alloc(&h->d_table, (size_t)(NPAT + 3) * 4));
    A(dev_alloc((TileDesc**)&h->d_tiles, (size_t)tiles));
    A(dev_alloc(&h->d_etable, (size_t)NPAT * 4));
    A(dev_alloc(&h->d_phi, (size_t)5 * h->len_g));
    A(dev_alloc(&h->d_priors, (size_t)4 * h->len_g));
    A(dev_alloc(&h->d_phi_pow, (size_t)5 * h->len_g * FREEZE));
    A(dev_alloc(&h->d_cov_total, (size_t)h->n_contigs_total));
    A(dev_alloc(&h->d_drop_thr, (size_t)h->n_contigs_total));
    A(dev_alloc(&h->d_ds, (size_t)h->nb * ds));
    A(dev_alloc(&h->d_benefit, (size_t)h->nb * rows));
    A(dev_alloc(&h->d_codes, (size_t)h->nb * rows * 2 + 16));
    A(dev_alloc(&h->d_bucket_sum, (size_t)sw * h->nb));
    A(dev_alloc(&h->d_bucket_sw, (size_t)sw * h->nb));
    A(dev_alloc(&h->d_fhat_w, (size_t)h->n_windows_total * 2));
    A(dev_alloc(&h->d_hist, (size_t)3 * HBINS + 4));
    A(dev_alloc(&h->d_strat_alloc, (size_t)srows * 2 * h->nb + 32));
    h->d_strat = h->d_strat_alloc;
    A(dev_alloc(&h->d_seg_accept, (size_t)h->n_seg * 2));
    A(dev_alloc(&h->d_rs_counts, (size_t)h->n_windows_total * 2));
    A(dev_alloc(&h->d_upd, 1));
    A(dev_alloc(&h->d_ingest_err, 1));
    A(dev_alloc((int64_t**)&h->scratch_d, 1));   // placeholder so scratch is never null
    h->scratch_d_bytes = sizeof(int64_t);
    if (rc != 0) { bossgpu_destroy(h); return rc; }
    BOSS_CUDA(cudaMallocHost((void**)&h->h_upd, sizeof(UpdateDev)));
    BOSS_CUDA(cudaMallocHost((void**)&h->h_ingest_err, sizeof(int32_t)));
    BOSS_CUDA(cudaHostAlloc((void**)&h->h_strat_own, std::max<size_t>(16, (size_t)srows * 2 * h->nb), cudaHostAllocMapped | cudaHostAllocPortable));
    h->h_strat = h->h_strat_own;
    BOSS_CUDA(cudaHostGetDevicePointer((void**)&h->h_strat_dev, h->h_strat, 0));
    BOSS_CUDA(cudaMallocHost((void**)&h->h_seg_accept, sizeof(unsigned long long) * 2 * h->n_seg));
    BOSS_CUDA(cudaMallocHost((void**)&h->h_bucket_sw, std::max<size_t>(1, (size_t)sw * h->nb)));
    memset(h->h_bucket_sw, 0, (size_t)sw * h->nb);
    memset(h->h_strat, 1, (size_t)srows * 2 * h->nb);          // Contig.strat starts all-accept (reference.py:118)
    memset(h->h_seg_accept, 0, sizeof(unsigned long long) * 2 * h->n_seg);
    for (auto& ev : h->ev) BOSS_CUDA(cudaEventCreate(&ev));

    // geometry tables
    {
        std::vector<int64_t> ts(h->n_seg + 1), rs(h->n_seg + 1), ss(h->n_seg + 1);
        for (int s = 0; s < h->n_seg; ++s) { ts[s] = h->segs[s].tile_off; rs[s] = h->segs[s].row_off; ss[s] = h->segs[s].srow_off; }
        ts[h->n_seg] = tiles; rs[h->n_seg] = rows; ss[h->n_seg] = srows;
        BOSS_CUDA(cudaMemcpy(h->d_segs, h->segs.data(), sizeof(SegDev) * h->n_seg, cudaMemcpyHostToDevice));
        BOSS_CUDA(cudaMemcpy(h->d_tile_start, ts.data(), sizeof(int64_t) * ts.size(), cudaMemcpyHostToDevice));
        BOSS_CUDA(cudaMemcpy(h->d_row_start, rs.data(), sizeof(int64_t) * rs.size(), cudaMemcpyHostToDevice));
        BOSS_CUDA(cudaMemcpy(h->d_srow_start, ss.data(), sizeof(int64_t) * ss.size(), cudaMemcpyHostToDevice));
        std::vector<int64_t> st(h->n_seg + 1);
        int64_t nt = 0;
        for (int s = 0; s < h->n_seg; ++s) { st[s] = nt; nt += ceil_div(h->segs[s].n_bins, SM_TILE); }
        st[h->n_seg] = nt;
        h->n_sm_tiles = nt;
        BOSS_CUDA(cudaMemcpy(h->d_sm_tile_start, st.data(), sizeof(int64_t) * st.size(), cudaMemcpyHostToDevice));
        BOSS_CUDA(cudaMemcpy(h->d_contig_len, h->contig_len_all.data(), sizeof(int64_t) * h->n_contigs_total, cudaMemcpyHostToDevice));
    }
    // reference bases, segment by segment onto the padded axis
    {
        const uint8_t* src = cfg->ref_codes;
        for (int s = 0; s < h->n_seg; ++s) {
            BOSS_CUDA(cudaMemcpy(h->d_ref + h->segs[s].site_off, src, (size_t)h->segs[s].len, cudaMemcpyHostToDevice));
            src += h->segs[s].len;
        }
    }
    // Contig.strat starts all-accept (reference.py:118)
    BOSS_CUDA(cudaMemset(h->d_strat, 1, (size_t)srows * 2 * h->nb));
    BOSS_CUDA(cudaMemset(h->d_drop_thr, 0xFF, sizeof(int32_t) * h->n_contigs_total));
    // scoring constants + dense table
    BOSS_CUDA(cudaMemcpy(h->d_phi, cfg->phi, sizeof(double) * 5 * h->len_g, cudaMemcpyHostToDevice));
    BOSS_CUDA(cudaMemcpy(h->d_priors, cfg->priors, sizeof(double) * 4 * h->len_g, cudaMemcpyHostToDevice));
    BOSS_CUDA(cudaMemcpy(h->d_phi_pow, cfg->phi_pow, sizeof(double) * 5 * h->len_g * FREEZE, cudaMemcpyHostToDevice));
    k_build_table<<<(unsigned)ceil_div(NPAT, 128), 128, 0, h->stream>>>(h->len_g, h->d_phi, h->d_priors, h->d_phi_pow,
                                                                       h->d_table, h->d_etable);
    BOSS_KERNEL_CHECK();
    {   // special rows behind the pattern table: frozen, never observed, dropped
        double extra[12];
        for (int i = 0; i < 4; ++i) { extra[i] = TINY; extra[4 + i] = h->score0; extra[8 + i] = 0.0; }
        BOSS_CUDA(cudaMemcpyAsync(h->d_table + (size_t)NPAT * 4, extra, sizeof extra, cudaMemcpyHostToDevice, h->stream));
        BOSS_CUDA(cudaStreamSynchronize(h->stream));
        // Without barcodes the all-zero pattern IS "never observed" (Q5: contig score0, not the table value), so
        // row 0 of the working table is patched; the true values are kept for bossgpu_get_score_table.
        BOSS_CUDA(cudaMemcpy(h->row0_true, h->d_table, sizeof(double) * 4, cudaMemcpyDeviceToHost));
        BOSS_CUDA(cudaMemcpy(h->row0_true + 4, h->d_etable, sizeof(double) * 4, cudaMemcpyDeviceToHost));
        if (h->nb == 1) {
            BOSS_CUDA(cudaMemcpy(h->d_table, extra + 4, sizeof(double) * 4, cudaMemcpyHostToDevice));
            double e0[4] = {h->ent0, h->ent0, h->ent0, h->ent0};
            BOSS_CUDA(cudaMemcpy(h->d_etable, e0, sizeof e0, cudaMemcpyHostToDevice));
        }
    }
    {   // tile descriptors of the score/bin pass
        std::vector<TileDesc> td((size_t)tiles);
        for (int s = 0; s < h->n_seg; ++s) {
            const SegDev& S = h->segs[s];
            for (int64_t i = 0; i < S.n_tiles; ++i) {
                TileDesc& d = td[(size_t)(S.tile_off + i)];
                const int64_t local0 = i * TILE;
                d.site_off = S.site_off + local0;
                d.ds_index = S.ds_off + local0 / BIN;
                d.n_sites = (int32_t)std::max<int64_t>(0, std::min<int64_t>(TILE, S.len - local0));
                d.n_bins = (int32_t)std::min<int64_t>(TILE / BIN, S.n_bins - local0 / BIN);
                const int64_t bucket = local0 / BUCKET;
                d.bucket = bucket < S.n_full_buckets ? (int32_t)(S.sw_off + bucket) : -1;
                d.contig = S.contig;
            }
        }
        BOSS_CUDA(cudaMemcpy(h->d_tiles, td.data(), sizeof(TileDesc) * td.size(), cudaMemcpyHostToDevice));
    }
    {   // split score/bin pass
        std::vector<int32_t> soc((size_t)h->n_contigs_total, -1);
        bool one_each = true;
        for (int s = 0; s < h->n_seg; ++s) {
            if (soc[h->segs[s].contig] >= 0) one_each = false;
            soc[h->segs[s].contig] = s;
        }
        A(dev_alloc(&h->d_seg_of_contig, (size_t)h->n_contigs_total));
        BOSS_CUDA(cudaMemcpy(h->d_seg_of_contig, soc.data(), sizeof(int32_t) * soc.size(), cudaMemcpyHostToDevice));
        A(dev_alloc(&h->d_touched, (size_t)(tiles / 32 + 2)));
        A(dev_alloc(&h->d_touched_list, (size_t)tiles + 1));
        A(dev_alloc(&h->d_pre_misc, 2));
        A(dev_alloc(&h->d_drop_thr_spec, (size_t)h->n_contigs_total));
        A(dev_alloc(&h->d_tile_cov, (size_t)(tiles + 1) * h->nb));
        A(dev_alloc(&h->d_tile_drop, (size_t)tiles + 1));
        BOSS_CUDA(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        BOSS_CUDA(cudaEventCreateWithFlags(&h->ev_pre_thr, cudaEventDisableTiming));
        BOSS_CUDA(cudaEventCreateWithFlags(&h->ev_pre_done, cudaEventDisableTiming));
        BOSS_CUDA(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
        h->multi_fused = h->nb > 1 && h->nb <= SBM_MAX_NB && getenv("BOSSGPU_NO_FUSED_BARCODES") == nullptr;
        h->prescore_ok = one_each && (h->nb == 1 || h->multi_fused) && getenv("BOSSGPU_NO_PRESCORE") == nullptr;
        if (h->multi_fused) {
            BOSS_CUDA(cudaFuncSetAttribute(k_score_bin_multi<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbm_smem_bytes(h->nb)));
            BOSS_CUDA(cudaFuncSetAttribute(k_score_bin_multi<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbm_smem_bytes(h->nb)));
            int resident = 0;                    // persistent grid = what the SMs hold at once (registers and shared memory decide)
            BOSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_score_bin_multi<0>, SBM_THREADS, sbm_smem_bytes(h->nb)));
            h->multi_ctas = std::max(1, resident);
        }
        BOSS_CUDA(cudaFuncSetAttribute(k_score_bin_tma<false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbt_smem_bytes(false, 2)));
    }
    BOSS_CUDA(cudaFuncSetAttribute(k_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BOSS_CUDA(cudaFuncSetAttribute(k_score_bin_tma<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbt_smem_bytes(false, 2)));
    BOSS_CUDA(cudaFuncSetAttribute(k_score_bin_tma<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbt_smem_bytes(false, 4)));
    BOSS_CUDA(cudaFuncSetAttribute(k_score_bin_tma<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbt_smem_bytes(true, 2)));
    {
        cudaDeviceProp prop;
        BOSS_CUDA(cudaGetDeviceProperties(&prop, h->device));
        h->n_sm = prop.multiProcessorCount;
        const char* env = getenv("BOSSGPU_SCORE_KERNEL");       // development A/B switch: "ldg" = non-staged variant
        h->score_kernel_ldg = env && strcmp(env, "ldg") == 0;
        const char* e2 = getenv("BOSSGPU_SCORE_STAGES");
        const char* e3 = getenv("BOSSGPU_SCORE_CTAS");
        // measured on B200 (scripts/sweep_score.sh): 2 stages x 3 CTAs/SM beats 4 stages x 2 CTAs/SM — the table
        // gathers want warps (and L1) more than the copies want depth
        h->score_stages = e2 ? atoi(e2) : 2;
        h->score_ctas_per_sm = e3 ? atoi(e3) : 3;
    }
    BOSS_CUDA(cudaFuncSetAttribute(k_smooth_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}

extern "C" int bossgpu_destroy(bossgpu_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {h->d_segs, h->d_tile_start, h->d_row_start, h->d_srow_start, h->d_ref, h->d_cov, h->d_rowflag,
                    h->d_table, h->d_etable, h->d_phi, h->d_priors, h->d_phi_pow, h->d_cov_total, h->d_drop_thr,
                    h->d_ds, h->d_benefit, h->d_codes, h->d_smu, h->d_expected, h->d_bucket_sum, h->d_bucket_sw, h->d_fhat_w,
                    h->d_hist, h->d_strat_alloc, h->d_upd, h->stage_d, h->scratch_d, h->d_ingest_err, h->d_mask_all, h->d_tiles,
                    h->d_shard_row_start, h->d_halo, h->d_sm_tile_start, h->d_contig_len, h->d_seg_accept, h->d_rs_counts,
                    h->d_mask_ptrs, h->d_fabric, h->d_peer_ptrs, h->d_fab_mask_ptrs, h->d_seg_of_contig, h->d_touched,
                    h->d_touched_list, h->d_pre_misc, h->d_drop_thr_spec, h->d_tile_cov, h->d_tile_drop, h->pre_stage_d, h->ingest_d};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->h_upd) cudaFreeHost(h->h_upd);
    if (h->h_ingest_err) cudaFreeHost(h->h_ingest_err);
    if (h->h_strat_own) cudaFreeHost(h->h_strat_own);
    if (h->reg_base) bossgpu_host_unregister(h->reg_base);
    if (h->h_seg_accept) cudaFreeHost(h->h_seg_accept);
    if (h->h_bucket_sw) cudaFreeHost(h->h_bucket_sw);
    if (h->stage_h) cudaFreeHost(h->stage_h);
    if (h->pre_stage_h) cudaFreeHost(h->pre_stage_h);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->ev_pre_thr) cudaEventDestroy(h->ev_pre_thr);
    if (h->ev_pre_done) cudaEventDestroy(h->ev_pre_done);
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    delete h;
    return 0;
}

extern "C" int bossgpu_synchronize(bossgpu_handle* h) {
    H_CHECK(h);
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ingest
// ------------------------------------------------------------------------------------------------
// per batch: an OpRec (32 B), its read offset and its read id per op slot, plus every read's reference span
struct IngestBuf {
    OpRec* rec; int32_t* q0; int32_t* read; int64_t* span;
};
static int ensure_ingest_buf(bossgpu_handle* h, size_t slots, size_t reads, IngestBuf* out) {
    slots = std::max<size_t>(slots, 1); reads = std::max<size_t>(reads, 1);
    const size_t o_q0 = round_up(sizeof(OpRec) * slots, 256), o_rd = o_q0 + round_up(sizeof(int32_t) * slots, 256);
    const size_t o_sp = o_rd + round_up(sizeof(int32_t) * slots, 256), bytes = o_sp + sizeof(int64_t) * reads;
    if (bytes > h->ingest_d_bytes) {
        BOSS_CUDA(cudaStreamSynchronize(h->stream));
        if (h->ingest_d) cudaFree(h->ingest_d);
        h->ingest_d = nullptr; h->ingest_d_bytes = 0;
        const size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&h->ingest_d, want);
        if (e != cudaSuccess) return fail(BOSSGPU_ENOMEM, "op-record cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        h->ingest_d_bytes = want;
    }
    char* base = (char*)h->ingest_d;
    out->rec = (OpRec*)base; out->q0 = (int32_t*)(base + o_q0); out->read = (int32_t*)(base + o_rd); out->span = (int64_t*)(base + o_sp);
    return 0;
}

// Everything of the coverage update that runs once the ops are on the device (scatter.cuh): per-op records, the
// aligned-column check of characters outside ACGT (which settles the batch's error flag), the depth totals, the
// scatter. `d_cig_off` holds n_reads + 1 entries; `n_slots` is its last one. `totals_src` (device, may be NULL):
// reference span of the batch per global contig, added only if the batch is accepted.
static int launch_scatter(bossgpu_handle* h, int64_t n_reads, int64_t n_slots, const int32_t* d_seg, const int64_t* d_tstart,
                          const int32_t* d_bc, const int64_t* d_cig_off, const int64_t* d_cig_end, const uint32_t* d_cig,
                          const int64_t* d_base_off, const uint8_t* d_bases, const uint8_t* d_rev, int ascii, bool count_totals,
                          bool check_q, const unsigned long long* totals_src,
                          PackedBases pk = PackedBases{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, int64_t n_exc = 0,
                          bool record_begin = true) {
    if (record_begin) BOSS_CUDA(cudaEventRecord(h->ev[0], h->stream));
    IngestBuf ib;
    TRY(ensure_ingest_buf(h, (size_t)n_slots, (size_t)n_reads, &ib));
    PrefixArgs pa;
    pa.n_reads = n_reads; pa.cig_off = d_cig_off; pa.cig_end = d_cig_end; pa.cigar = d_cig; pa.seg_of = d_seg; pa.tstart = d_tstart;
    pa.barcode = d_bc; pa.base_off = d_base_off; pa.rev = d_rev; pa.pk_off = pk.data ? pk.off : nullptr; pa.segs = h->d_segs;
    pa.n_seg = h->n_seg; pa.nb = h->nb; pa.P = h->P; pa.check_q = check_q ? 1 : 0; pa.rec = ib.rec; pa.rec_q0 = ib.q0;
    pa.rec_read = (pk.data && n_exc > 0) ? ib.read : nullptr; pa.read_span = ib.span; pa.err = h->d_ingest_err;
    k_op_prefix<<<(unsigned)std::min<int64_t>(n_reads, 1 << 20), PX_THREADS, 0, h->stream>>>(pa);
    BOSS_KERNEL_CHECK();
    h->launches++;
    if (pk.data) {
        if (n_exc > 0) {
            k_check_exc<<<(unsigned)ceil_div(n_exc, 128), 128, 0, h->stream>>>(n_exc, pk, d_rev, d_base_off, d_cig_off, d_cig_end, ib.q0,
                                                                               d_cig, h->d_ingest_err);
            BOSS_KERNEL_CHECK();
            h->launches++;
        }
    } else {
        k_check_bases<<<(unsigned)(h->n_sm * 8), 256, 0, h->stream>>>(n_reads, d_base_off, d_bases, ascii, d_cig_off, d_cig_end, ib.q0,
                                                                      d_cig, h->d_ingest_err);
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    if (totals_src) {
        k_add_u64_if_ok<<<(unsigned)ceil_div(h->n_contigs_total, 256), 256, 0, h->stream>>>(h->d_cov_total, totals_src,
                                                                                          h->n_contigs_total, h->d_ingest_err);
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    if (count_totals) {
        k_add_read_totals<<<(unsigned)ceil_div(n_reads, 256), 256, 0, h->stream>>>(n_reads, d_seg, d_tstart, ib.span, h->d_segs,
                                                                                  h->n_seg, h->d_cov_total, h->d_ingest_err);
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    ScatterArgs a;
    a.n_slots = d_cig_off + n_reads; a.rec = ib.rec; a.rec_read = pa.rec_read; a.bases = d_bases; a.base_is_ascii = ascii; a.pk = pk;
    a.P = h->P; a.cov = h->d_cov; a.err = h->d_ingest_err;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_slots, SC_THREADS), (int64_t)h->n_sm * 32));
    if (pk.data) k_scatter_ops<0><<<grid, SC_THREADS, 0, h->stream>>>(a);
    else if (ascii) k_scatter_ops<1><<<grid, SC_THREADS, 0, h->stream>>>(a);
    else k_scatter_ops<2><<<grid, SC_THREADS, 0, h->stream>>>(a);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaEventRecord(h->ev[1], h->stream));
    h->ev_valid[0] = true;
    h->launches++;
    return 0;
}

static uint64_t batch_hash(int64_t n, const int32_t* contig, const int64_t* tstart, const int64_t* tend) {
    uint64_t hsh = 0x9E3779B97F4A7C15ull;
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t t0 = (uint64_t)std::min(tstart[i], tend[i]), t1 = (uint64_t)std::max(tstart[i], tend[i]);
        hsh = (hsh ^ ((uint64_t)(uint32_t)contig[i] + (t0 << 20) + (t1 << 42) + (t1 >> 22))) * 0x100000001B3ull;
        hsh ^= hsh >> 29;
    }
    return hsh;
}

// The totals the early pass formed its thresholds from must not move before it has read them.
static int spec_order_totals(bossgpu_handle* h) {
    if (h->spec_state == 1) BOSS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pre_thr, 0));
    return 0;
}

// any ingest other than the announced text batch voids the announcement (the update then scores every tile again)
static int prescore_invalidate(bossgpu_handle* h) {
    TRY(spec_order_totals(h));
    if (h->prescore_state == 1 || h->prescore_state == 2) h->prescore_state = -1;
    return 0;
}

static int check_ingest_error(bossgpu_handle* h) {
    BOSS_CUDA(cudaMemcpyAsync(h->h_ingest_err, h->d_ingest_err, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    int32_t e = *h->h_ingest_err;
    if (e != 0) {
        BOSS_CUDA(cudaMemsetAsync(h->d_ingest_err, 0, sizeof(int32_t), h->stream));
        if (e == BOSSGPU_EBASE)
            return fail(BOSSGPU_EBASE, "a read base outside ACGT reached the coverage scatter (index out of bounds upstream)");
        if (e == BOSSGPU_ESHAPE)
            return fail(BOSSGPU_ESHAPE, "CIGAR does not consume exactly the aligned read slice");
        return fail(e, "device-side ingest error %d", e);
    }
    return 0;
}

extern "C" int bossgpu_ingest_packed(bossgpu_handle* h, int64_t n_reads, const int32_t* seg, const int64_t* tstart,
                                     const int32_t* barcode, const int64_t* cig_off, const uint32_t* cigar,
                                     const int64_t* base_off, const uint8_t* bases, int base_is_ascii, int on_device,
                                     const int64_t* contig_cov_add) {
    H_CHECK(h);
    if (n_reads < 0) return fail(BOSSGPU_EINVAL, "negative read count");
    TRY(prescore_invalidate(h));
    if (n_reads == 0) {
        if (!contig_cov_add) return 0;
        const unsigned long long* src = (const unsigned long long*)contig_cov_add;
        if (!on_device) {
            const size_t bytes = sizeof(int64_t) * h->n_contigs_total;
            TRY(ensure_stage(h, bytes));
            memcpy(h->stage_h, contig_cov_add, bytes);
            BOSS_CUDA(cudaMemcpyAsync(h->stage_d, h->stage_h, bytes, cudaMemcpyHostToDevice, h->stream));
            src = (const unsigned long long*)h->stage_d;
        }
        k_add_u64<<<(unsigned)ceil_div(h->n_contigs_total, 256), 256, 0, h->stream>>>(h->d_cov_total, src, h->n_contigs_total);
        BOSS_KERNEL_CHECK();
        h->launches++;
        if (!on_device) BOSS_CUDA(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (!seg || !tstart || !barcode || !cig_off || !cigar || !base_off || !bases) return fail(BOSSGPU_EINVAL, "null batch array");
    if (on_device) {
        // the grid and the op-record buffer are sized by the number of ops, which only the device knows here
        int64_t n_ops = 0;
        BOSS_CUDA(cudaMemcpyAsync(&n_ops, cig_off + n_reads, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        BOSS_CUDA(cudaStreamSynchronize(h->stream));
        if (n_ops < 0) return fail(BOSSGPU_EINVAL, "offset arrays must be non-decreasing");
        TRY(launch_scatter(h, n_reads, n_ops, seg, tstart, barcode, cig_off, cig_off + 1, cigar, base_off, bases, nullptr, base_is_ascii,
                           contig_cov_add == nullptr, /*check_q=*/true, (const unsigned long long*)contig_cov_add));
        return 0;   // errors surface at the next update / synchronize-checked call
    }
    const int64_t n_ops = cig_off[n_reads], n_bases = base_off[n_reads];
    if (cig_off[0] != 0 || base_off[0] != 0 || n_ops < 0 || n_bases < 0) return fail(BOSSGPU_EINVAL, "offset arrays must start at 0");
    // one staging blob: [seg | barcode | tstart | cig_off | base_off | cov_add | cigar | bases]
    size_t o_seg = 0;
    size_t o_bc = o_seg + round_up(sizeof(int32_t) * n_reads, 16);
    size_t o_ts = o_bc + round_up(sizeof(int32_t) * n_reads, 16);
    size_t o_co = o_ts + round_up(sizeof(int64_t) * n_reads, 16);
    size_t o_bo = o_co + round_up(sizeof(int64_t) * (n_reads + 1), 16);
    size_t o_ca = o_bo + round_up(sizeof(int64_t) * (n_reads + 1), 16);
    size_t o_cg = o_ca + round_up(sizeof(int64_t) * h->n_contigs_total, 16);
    size_t o_bs = o_cg + round_up(sizeof(uint32_t) * std::max<int64_t>(n_ops, 1), 16);
    size_t total = o_bs + round_up(std::max<int64_t>(n_bases, 1), 16);
    TRY(ensure_stage(h, total));
    char* hs = (char*)h->stage_h;
    memcpy(hs + o_seg, seg, sizeof(int32_t) * n_reads);
    memcpy(hs + o_bc, barcode, sizeof(int32_t) * n_reads);
    memcpy(hs + o_ts, tstart, sizeof(int64_t) * n_reads);
    memcpy(hs + o_co, cig_off, sizeof(int64_t) * (n_reads + 1));
    memcpy(hs + o_bo, base_off, sizeof(int64_t) * (n_reads + 1));
    if (contig_cov_add) memcpy(hs + o_ca, contig_cov_add, sizeof(int64_t) * h->n_contigs_total);
    memcpy(hs + o_cg, cigar, sizeof(uint32_t) * n_ops);
    memcpy(hs + o_bs, bases, (size_t)n_bases);
    BOSS_CUDA(cudaMemcpyAsync(h->stage_d, hs, total, cudaMemcpyHostToDevice, h->stream));
    char* ds = (char*)h->stage_d;
    TRY(launch_scatter(h, n_reads, n_ops, (const int32_t*)(ds + o_seg), (const int64_t*)(ds + o_ts), (const int32_t*)(ds + o_bc),
                       (const int64_t*)(ds + o_co), (const int64_t*)(ds + o_co) + 1, (const uint32_t*)(ds + o_cg),
                       (const int64_t*)(ds + o_bo), (const uint8_t*)(ds + o_bs), nullptr, base_is_ascii, contig_cov_add == nullptr,
                       /*check_q=*/true, contig_cov_add ? (const unsigned long long*)(ds + o_ca) : nullptr));
    return check_ingest_error(h);
}

extern "C" int64_t bossgpu_tokenize_cigar(const char* text, int64_t len, uint32_t* out, int64_t cap, int64_t* ref_span,
                                          int64_t* query_span) {
    if (!text || len < 0 || (!out && cap > 0)) return fail(BOSSGPU_EINVAL, "bad tokenizer arguments");
    int64_t r = 0, q = 0;
    int64_t n = tokenize_cigar(text, len, out, cap, &r, &q);
    if (n < 0) return fail(BOSSGPU_EINVAL, "CIGAR tokenizer: output capacity %lld too small", (long long)cap);
    if (ref_span) *ref_span = r;
    if (query_span) *query_span = q;
    return n;
}

// Common text path: per read a CIGAR string and the aligned slice of the read, wherever they live.
struct TextRead {
    const char* cigar; int64_t cigar_len;
    const char* seq;   int64_t seq_len;       // the slice itself, original orientation
};

static int ingest_text_impl(bossgpu_handle* h, int64_t n_all, const int32_t* contig, const int64_t* tstart,
                            const int64_t* tend, const int32_t* barcode, const uint8_t* rev,
                            const std::vector<TextRead>& all_reads, int n_threads, const int64_t* batch_cov_add = nullptr) {
    // The caller hands every shard the WHOLE batch. A shard holds one contiguous range of the genome axis, so
    // at most one of its segments belongs to any contig: reads are routed to that segment when they overlap
    // it (the scatter kernel clips at the segment edges, so a read spanning a shard edge lands in both
    // shards), and the depth total of every contig (dropout rule, reference.py:157-158) advances by the
    // reference span of ALL the batch's reads on it, local or not.
    std::vector<int32_t> seg_of_contig(h->n_contigs_total, -1);
    for (int s = 0; s < h->n_seg; ++s) {
        if (seg_of_contig[h->segs[s].contig] >= 0)
            return fail(BOSSGPU_ESTATE, "the text ingest path needs at most one segment per contig in a shard");
        seg_of_contig[h->segs[s].contig] = s;
    }
    const bool trace = getenv("BOSSGPU_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    auto t_begin = now();
    std::vector<int64_t> sel;
    sel.reserve((size_t)n_all);
    std::vector<unsigned long long> cov_add((size_t)h->n_contigs_total, 0ull);
    for (int64_t i = 0; i < n_all; ++i) {
        if (contig[i] < 0 || contig[i] >= h->n_contigs_total)
            return fail(BOSSGPU_EINVAL, "read %lld: contig index out of range", (long long)i);
        const int64_t t0 = std::min(tstart[i], tend[i]), t1 = std::max(tstart[i], tend[i]);
        cov_add[contig[i]] += (unsigned long long)(t1 - t0);
        const int32_t sg = seg_of_contig[contig[i]];
        if (sg < 0) continue;
        const SegDev& S = h->segs[sg];
        if (t1 <= S.start || t0 >= S.start + S.len) continue;
        sel.push_back(i);
    }
    if (batch_cov_add)                // routed batch: the caller hands the reference span of the WHOLE batch per contig
        for (int k = 0; k < h->n_contigs_total; ++k) cov_add[k] = (unsigned long long)batch_cov_add[k];
    TRY(spec_order_totals(h));
    if (h->prescore_state == 1)       // is this the batch that was announced? Same reads, same intervals -> its tile marks hold
        h->prescore_state = (n_all == h->pre_n_reads && batch_hash(n_all, contig, tstart, tend) == h->pre_hash) ? 2 : -1;
    else if (h->prescore_state != 0)  // a second ingest after the announced one: its tiles carry no marks -> score every tile
        h->prescore_state = -1;
    const int64_t n_reads = (int64_t)sel.size();
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char* lws = getenv("LOCAL_WORLD_SIZE")) {       // one process per GPU (torchrun): share the host's cores
        const int n_local = atoi(lws);
        if (n_local > 1) hw = std::max(4u, hw / (unsigned)n_local);
    }
    int T = n_threads > 0 ? n_threads : (int)std::min<unsigned>(hw > 2 ? hw - 1 : hw, 32u);   // one thread keeps issuing copies
    if (n_reads < 64) T = 1;
    T = (int)std::max<int64_t>(1, std::min<int64_t>(T, n_reads));
    // The batch is cut into G groups of consecutive reads; the worker threads tokenise and pack group after
    // group while the copy engine already moves the finished groups to the GPU, so the PCIe transfer hides
    // behind the host work instead of following it.
    const int G = n_reads >= 512 ? 4 : 1;
    // ---- staging blob (pinned host image = device image): per-read scalars | CIGAR text | read bases packed 2 bits
    //      each, every read's run starting on its own byte. The host only copies and packs; the CIGARs are
    //      tokenised on the GPU (k_tokenize) into op slots that never cross PCIe. ------------------------------
    int64_t total_chars = 0, ops_cap = 0, total_text = 0, total_packed = 0;
    for (int64_t j = 0; j < n_reads; ++j) {
        const TextRead& R = all_reads[sel[j]];
        total_chars += R.cigar_len + R.seq_len;
        ops_cap += R.cigar_len / 2 + 1;
        total_text += R.cigar_len;
        total_packed += (R.seq_len + 3) / 4;
    }
    // task (g, t) = reads [cut[g*T+t], cut[g*T+t+1]): equal shares of characters
    std::vector<int64_t> cut((size_t)G * T + 1, n_reads);
    {
        const int64_t parts = (int64_t)G * T;
        const int64_t per = total_chars / parts + 1;
        int64_t acc = 0;
        int64_t t = 1;
        cut[0] = 0;
        for (int64_t j = 0; j < n_reads && t < parts; ++j) {
            acc += all_reads[sel[j]].cigar_len + all_reads[sel[j]].seq_len;
            while (t < parts && acc >= per * t) cut[t++] = j + 1;
        }
    }
    const int64_t nr1 = std::max<int64_t>(n_reads, 1);
    size_t o_seg = 0;
    size_t o_bc = o_seg + round_up(sizeof(int32_t) * nr1, 16);
    size_t o_ts = o_bc + round_up(sizeof(int32_t) * nr1, 16);
    size_t o_co = o_ts + round_up(sizeof(int64_t) * nr1, 16);          // first op slot of each read [n+1]
    size_t o_to = o_co + round_up(sizeof(int64_t) * (nr1 + 1), 16);    // CIGAR text offsets [n+1]
    size_t o_sp = o_to + round_up(sizeof(int64_t) * (nr1 + 1), 16);    // reference span tend - tstart
    size_t o_bo = o_sp + round_up(sizeof(int64_t) * nr1, 16);          // base offsets [n+1] (slice lengths)
    size_t o_po = o_bo + round_up(sizeof(int64_t) * (nr1 + 1), 16);    // byte offset of each read's packed bases
    size_t o_rv = o_po + round_up(sizeof(int64_t) * nr1, 16);
    size_t o_ca = o_rv + round_up((size_t)nr1, 16);
    size_t o_tx = o_ca + round_up(sizeof(unsigned long long) * h->n_contigs_total, 16);
    const size_t small_bytes = o_tx;
    size_t o_pk = o_tx + round_up((size_t)std::max<int64_t>(total_text, 1), 16);
    size_t total = o_pk + round_up(std::max<int64_t>(total_packed, 1), 16);
    TRY(ensure_stage(h, total));
    // device-only: op slots + where each read's ops end
    const size_t x_ops = 0, x_end = round_up(sizeof(uint32_t) * std::max<int64_t>(ops_cap, 1), 16);
    const size_t x_exc = x_end + round_up(sizeof(int64_t) * nr1, 16);
    TRY(ensure_scratch(h, x_exc));
    char* hs = (char*)h->stage_h;
    char* ds = (char*)h->stage_d;
    int32_t* s_seg = (int32_t*)(hs + o_seg);
    int32_t* s_bc = (int32_t*)(hs + o_bc);
    int64_t* s_ts = (int64_t*)(hs + o_ts);
    int64_t* s_co = (int64_t*)(hs + o_co);
    int64_t* s_to = (int64_t*)(hs + o_to);
    int64_t* s_sp = (int64_t*)(hs + o_sp);
    int64_t* s_bo = (int64_t*)(hs + o_bo);
    int64_t* s_po = (int64_t*)(hs + o_po);
    uint8_t* s_rv = (uint8_t*)(hs + o_rv);
    char* s_tx = hs + o_tx;
    uint8_t* s_pk = (uint8_t*)(hs + o_pk);
    memcpy(hs + o_ca, cov_add.data(), sizeof(unsigned long long) * h->n_contigs_total);
    {
        int64_t co = 0, bo = 0, po = 0, to = 0;
        for (int64_t j = 0; j < n_reads; ++j) {
            const int64_t i = sel[j];
            s_seg[j] = seg_of_contig[contig[i]];
            s_bc[j] = barcode[i];
            s_ts[j] = std::min(tstart[i], tend[i]);
            s_sp[j] = std::max(tstart[i], tend[i]) - s_ts[j];
            s_rv[j] = rev[i] ? 1 : 0;
            s_co[j] = co;
            s_to[j] = to;
            s_bo[j] = bo;
            s_po[j] = po;
            co += all_reads[i].cigar_len / 2 + 1;
            to += all_reads[i].cigar_len;
            bo += all_reads[i].seq_len;
            po += (all_reads[i].seq_len + 3) / 4;
        }
        s_bo[n_reads] = bo;
        s_to[n_reads] = to;
        s_co[n_reads] = co;
    }
    BOSS_CUDA(cudaMemcpyAsync(ds, hs, small_bytes, cudaMemcpyHostToDevice, h->stream));
    const double ms_layout = ms_since(t_begin);
    struct Exc { int64_t read; int32_t pos; uint8_t ch; };
    std::vector<std::vector<Exc>> exc((size_t)T);
    std::vector<std::atomic<int>> done((size_t)G);
    for (auto& d : done) d.store(0);
    auto work = [&](int t) {
        for (int g = 0; g < G; ++g) {
            for (int64_t j = cut[(size_t)g * T + t]; j < cut[(size_t)g * T + t + 1]; ++j) {
                const TextRead& R = all_reads[sel[j]];
                memcpy(s_tx + s_to[j], R.cigar, (size_t)R.cigar_len);
                // slices stay in sequencing orientation; the scatter kernel reverse-complements on the fly
                pack_bases((const unsigned char*)R.seq, R.seq_len, s_pk + s_po[j],
                           [&](int64_t pos, unsigned char ch) { exc[t].push_back(Exc{j, (int32_t)pos, ch}); });
            }
            done[g].fetch_add(1, std::memory_order_release);
        }
    };
    auto ship = [&](int g) -> int {
        // reads of group g are tasks (g, 0..T-1): one run of CIGAR text and one run of packed bases
        const int64_t r0 = cut[(size_t)g * T], r1 = cut[(size_t)(g + 1) * T];
        if (r1 <= r0) return 0;
        const size_t c0 = (size_t)s_to[r0], c1 = (size_t)s_to[r1];
        const size_t p0 = (size_t)s_po[r0], p1 = (size_t)(r1 < n_reads ? s_po[r1] : total_packed);
        if (c1 > c0) BOSS_CUDA(cudaMemcpyAsync(ds + o_tx + c0, hs + o_tx + c0, c1 - c0, cudaMemcpyHostToDevice, h->stream));
        if (p1 > p0) BOSS_CUDA(cudaMemcpyAsync(ds + o_pk + p0, hs + o_pk + p0, p1 - p0, cudaMemcpyHostToDevice, h->stream));
        return 0;
    };
    if (T == 1) {
        work(0);
        for (int g = 0; g < G; ++g) TRY(ship(g));
    } else {
        WorkerPool& wp = WorkerPool::instance();
        std::unique_lock<std::mutex> own(wp.owner, std::try_to_lock);
        const std::function<void(int)> job = work;
        std::vector<std::thread> spawned;
        if (own.owns_lock()) wp.start(T, job);
        else for (int t = 0; t < T; ++t) spawned.emplace_back(work, t);
        int rc_ship = 0;
        for (int g = 0; g < G; ++g) {
            while (done[g].load(std::memory_order_acquire) < T) std::this_thread::yield();
            if (rc_ship == 0) rc_ship = ship(g);
        }
        if (own.owns_lock()) wp.wait();
        for (auto& th : spawned) th.join();
        if (rc_ship != 0) return rc_ship;
    }
    const double ms_pass2 = ms_since(t_begin);
    char* xs = (char*)h->scratch_d;
    PackedBases pk{(const uint8_t*)(ds + o_pk), (const int64_t*)(ds + o_po), nullptr, nullptr, nullptr, nullptr};
    size_t n_exc = 0;
    for (auto& v : exc) n_exc += v.size();
    std::vector<char> exc_blob;
    if (n_exc) {
        // characters outside ACGT (rare): [exc_off i64[n+1] | pos i32[n_exc] | read i32[n_exc] | ch u8[n_exc]]
        const size_t e_pos = round_up(sizeof(int64_t) * (n_reads + 1), 16), e_rd = e_pos + round_up(sizeof(int32_t) * n_exc, 16);
        const size_t e_ch = e_rd + round_up(sizeof(int32_t) * n_exc, 16);
        exc_blob.assign(e_ch + round_up(n_exc, 16), 0);
        int64_t* eo = (int64_t*)exc_blob.data();
        int32_t* ep = (int32_t*)(exc_blob.data() + e_pos);
        int32_t* er = (int32_t*)(exc_blob.data() + e_rd);
        uint8_t* ec = (uint8_t*)(exc_blob.data() + e_ch);
        std::vector<Exc> all;
        all.reserve(n_exc);
        for (auto& v : exc) all.insert(all.end(), v.begin(), v.end());
        std::sort(all.begin(), all.end(), [](const Exc& a, const Exc& b) { return a.read != b.read ? a.read < b.read : a.pos < b.pos; });
        size_t x = 0;
        for (int64_t j = 0; j <= n_reads; ++j) {
            while (x < all.size() && all[x].read < j) ++x;
            eo[j] = (int64_t)x;
        }
        for (size_t q = 0; q < all.size(); ++q) { ep[q] = all[q].pos; er[q] = (int32_t)all[q].read; ec[q] = all[q].ch; }
        TRY(ensure_scratch(h, x_exc + exc_blob.size()));
        xs = (char*)h->scratch_d;
        BOSS_CUDA(cudaMemcpyAsync(xs + x_exc, exc_blob.data(), exc_blob.size(), cudaMemcpyHostToDevice, h->stream));
        pk.exc_off = (const int64_t*)(xs + x_exc);
        pk.exc_pos = (const int32_t*)(xs + x_exc + e_pos);
        pk.exc_char = (const uint8_t*)(xs + x_exc + e_ch);
        pk.exc_read = (const int32_t*)(xs + x_exc + e_rd);
    }
    if (n_reads > 0) {
        BOSS_CUDA(cudaEventRecord(h->ev[0], h->stream));
        k_tokenize<<<(unsigned)std::min<int64_t>(n_reads, 1 << 20), TK_THREADS, 0, h->stream>>>(
            n_reads, ds + o_tx, (const int64_t*)(ds + o_to), (const int64_t*)(ds + o_co), (const int64_t*)(ds + o_bo),
            (const int64_t*)(ds + o_sp), (uint32_t*)(xs + x_ops), (int64_t*)(xs + x_end), h->d_ingest_err);
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    if (n_reads > 0) {
        // every contig's depth total advances by the whole batch's reference span on it (whichever shard holds the
        // reads), once the aligned-column check has accepted the batch
        TRY(launch_scatter(h, n_reads, ops_cap, (const int32_t*)(ds + o_seg), (const int64_t*)(ds + o_ts), (const int32_t*)(ds + o_bc),
                           (const int64_t*)(ds + o_co), (const int64_t*)(xs + x_end), (const uint32_t*)(xs + x_ops),
                           (const int64_t*)(ds + o_bo), nullptr, (const uint8_t*)(ds + o_rv), /*ascii=*/1,
                           /*count_totals=*/false, /*check_q=*/false, (const unsigned long long*)(ds + o_ca), pk, (int64_t)n_exc,
                           /*record_begin=*/false));
    } else {
        k_add_u64_if_ok<<<(unsigned)ceil_div(h->n_contigs_total, 256), 256, 0, h->stream>>>(
            h->d_cov_total, (const unsigned long long*)(ds + o_ca), h->n_contigs_total, h->d_ingest_err);
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    int rc = check_ingest_error(h);       // synchronises: exc_blob, the staging buffer and the scratch are free again
    if (rc != 0 && h->prescore_state == 2) h->prescore_state = -1;    // a rejected batch adds nothing to the totals
    h->last_ingest_h2d = (int64_t)(small_bytes + (size_t)total_text + (size_t)total_packed + exc_blob.size());
    if (trace)
        fprintf(stderr, "[bossgpu] ingest %lld of %lld reads, %d threads x %d groups (hw %u): layout %.2f ms, copy+pack (h2d overlapped) %.2f, tokenise + scatter on the GPU %.2f; %zu B staged\n",
                (long long)n_reads, (long long)n_all, T, G, std::thread::hardware_concurrency(), ms_layout, ms_pass2 - ms_layout,
                ms_since(t_begin) - ms_pass2, total);
    return rc;
}

extern "C" int bossgpu_ingest_records(bossgpu_handle* h, int64_t n_reads, const int32_t* contig, const int64_t* tstart,
                                      const int64_t* tend, const int32_t* barcode, const uint8_t* rev,
                                      const int64_t* cig_off, const char* cigar_text, const int64_t* seq_off,
                                      const char* seq_text, int n_threads) {
    H_CHECK(h);
    if (n_reads < 0) return fail(BOSSGPU_EINVAL, "negative read count");
    if (n_reads == 0) return 0;
    if (!contig || !tstart || !tend || !barcode || !rev || !cig_off || !cigar_text || !seq_off || !seq_text)
        return fail(BOSSGPU_EINVAL, "null batch array");
    if (cig_off[0] != 0 || seq_off[0] != 0) return fail(BOSSGPU_EINVAL, "offset arrays must start at 0");
    std::vector<TextRead> reads((size_t)n_reads);
    for (int64_t i = 0; i < n_reads; ++i)
        reads[i] = TextRead{cigar_text + cig_off[i], cig_off[i + 1] - cig_off[i], seq_text + seq_off[i], seq_off[i + 1] - seq_off[i]};
    return ingest_text_impl(h, n_reads, contig, tstart, tend, barcode, rev, reads, n_threads);
}

extern "C" int bossgpu_ingest_records_ptr(bossgpu_handle* h, int64_t n_reads, const int32_t* contig, const int64_t* tstart,
                                          const int64_t* tend, const int32_t* barcode, const uint8_t* rev,
                                          const uint64_t* cigar_ptr, const int64_t* cigar_len, const uint64_t* seq_ptr,
                                          const int64_t* seq_from, const int64_t* seq_to, int n_threads) {
    H_CHECK(h);
    if (n_reads < 0) return fail(BOSSGPU_EINVAL, "negative read count");
    if (n_reads == 0) return 0;
    if (!contig || !tstart || !tend || !barcode || !rev || !cigar_ptr || !cigar_len || !seq_ptr || !seq_from || !seq_to)
        return fail(BOSSGPU_EINVAL, "null batch array");
    std::vector<TextRead> reads((size_t)n_reads);
    for (int64_t i = 0; i < n_reads; ++i) {
        if (seq_to[i] < seq_from[i] || seq_from[i] < 0 || cigar_len[i] < 0)
            return fail(BOSSGPU_EINVAL, "read %lld: bad slice bounds", (long long)i);
        reads[i] = TextRead{(const char*)(uintptr_t)cigar_ptr[i], cigar_len[i],
                            (const char*)(uintptr_t)seq_ptr[i] + seq_from[i], seq_to[i] - seq_from[i]};
    }
    return ingest_text_impl(h, n_reads, contig, tstart, tend, barcode, rev, reads, n_threads);
}

extern "C" int bossgpu_ingest_records_routed(bossgpu_handle* h, int64_t n_reads, const int32_t* contig, const int64_t* tstart,
                                             const int64_t* tend, const int32_t* barcode, const uint8_t* rev,
                                             const uint64_t* cigar_ptr, const int64_t* cigar_len, const uint64_t* seq_ptr,
                                             const int64_t* seq_from, const int64_t* seq_to, const int64_t* batch_cov_add,
                                             int n_threads) {
    H_CHECK(h);
    if (n_reads < 0 || !batch_cov_add) return fail(BOSSGPU_EINVAL, "bad routed batch");
    if (n_reads > 0 && (!contig || !tstart || !tend || !barcode || !rev || !cigar_ptr || !cigar_len || !seq_ptr || !seq_from || !seq_to))
        return fail(BOSSGPU_EINVAL, "null batch array");
    std::vector<TextRead> reads((size_t)n_reads);
    for (int64_t i = 0; i < n_reads; ++i) {
        if (seq_to[i] < seq_from[i] || seq_from[i] < 0 || cigar_len[i] < 0)
            return fail(BOSSGPU_EINVAL, "read %lld: bad slice bounds", (long long)i);
        reads[i] = TextRead{(const char*)(uintptr_t)cigar_ptr[i], cigar_len[i],
                            (const char*)(uintptr_t)seq_ptr[i] + seq_from[i], seq_to[i] - seq_from[i]};
    }
    return ingest_text_impl(h, n_reads, contig, tstart, tend, barcode, rev, reads, n_threads, batch_cov_add);
}

// ------------------------------------------------------------------------------------------------
// update
// ------------------------------------------------------------------------------------------------
static int validate_params(const bossgpu_update_params* p) {
    if (!p) return fail(BOSSGPU_EINVAL, "null update params");
    for (int i = 0; i < NSTEPS; ++i) {
        // bn.move_sum raises for window < 1 (reference.py:259-260 with approx_ccl < 100)
        if (p->w[i] < 1) return fail(BOSSGPU_EINVAL, "staircase window %d is %d bins; Bottleneck's move_sum rejects windows < 1", i, p->w[i]);
        if (i > 0 && p->w[i] < p->w[i - 1]) return fail(BOSSGPU_EINVAL, "staircase windows must be non-decreasing");
        // the smoothing tile (1024 bins + a halo of w - 1 on both sides) has to fit 200 KB of shared memory
        if (p->w[i] > 12000) return fail(BOSSGPU_EINVAL, "staircase window %d exceeds the supported 12000 bins (1.2 Mb reads)", i);
    }
    return 0;
}

#define EV_BEGIN(i) BOSS_CUDA(cudaEventRecord(h->ev[2 * (i)], h->stream))
#define EV_END(i)   do { BOSS_CUDA(cudaEventRecord(h->ev[2 * (i) + 1], h->stream)); h->ev_valid[i] = true; } while (0)

static int ensure_pre_stage(bossgpu_handle* h, size_t want) {
    if (h->pre_stage_bytes >= want) return 0;
    want = want * 2 + 4096;
    if (h->pre_stage_h) cudaFreeHost(h->pre_stage_h);
    if (h->pre_stage_d) cudaFree(h->pre_stage_d);
    h->pre_stage_h = nullptr; h->pre_stage_d = nullptr; h->pre_stage_bytes = 0;
    if (cudaMallocHost(&h->pre_stage_h, want) != cudaSuccess || cudaMalloc(&h->pre_stage_d, want) != cudaSuccess) {
        cudaGetLastError();
        return fail(BOSSGPU_ENOMEM, "staging for the early score pass (%zu bytes)", want);
    }
    h->pre_stage_bytes = want;
    return 0;
}

static ScoreArgs score_args(bossgpu_handle* h) {
    ScoreArgs a;
    a.tiles = (const TileDesc*)h->d_tiles; a.nb = h->nb; a.P = h->P;
    a.ref = h->d_ref; a.cov = h->d_cov; a.rowflag = h->d_rowflag; a.table = h->d_table; a.drop_thr = h->d_drop_thr;
    a.ds = h->d_ds; a.ds_len = h->ds_len; a.tile_cov = h->d_tile_cov; a.tile_drop = h->d_tile_drop;
    return a;
}

static int phase0_scores(bossgpu_handle* h, const bossgpu_update_params* p) {
    // An early pass over every tile may be in flight on stream2 (bossgpu_prescore_begin). If the batch that was ingested
    // since is the one that was announced (bossgpu_prescore), only the tiles it wrote to — and the tiles of contigs
    // whose dropout threshold moved with the new depth total — are scored again; otherwise every tile is.
    const bool begun = h->spec_state == 1;
    const int ann = h->prescore_state;
    const bool split = begun && (ann == 2 || (ann == 1 && h->pre_n_reads == 0));
    h->spec_state = 0;
    h->prescore_state = 0;
    if (begun) BOSS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pre_done, 0));
    // reset per-update device scalars (keeps `error`)
    BOSS_CUDA(cudaMemsetAsync(h->d_upd, 0, offsetof(UpdateDev, error), h->stream));
    BOSS_CUDA(cudaMemsetAsync(h->d_bucket_sum, 0, sizeof(unsigned long long) * h->n_sw * h->nb, h->stream));
    k_drop_thresholds<<<(unsigned)ceil_div(h->n_contigs_total, 128), 128, 0, h->stream>>>(
        h->n_contigs_total, h->d_contig_len, h->nb, h->d_cov_total, h->d_drop_thr);
    BOSS_KERNEL_CHECK();
    h->launches++;
    ScoreArgs a = score_args(h);
    EV_BEGIN(1);
    if (split) {
        k_mark_changed<<<(unsigned)h->n_seg, 256, 0, h->stream>>>(h->d_segs, h->d_drop_thr_spec, h->d_drop_thr, h->d_touched,
                                                                 h->d_touched_list, reinterpret_cast<unsigned*>(h->d_pre_misc));
        BOSS_KERNEL_CHECK();
        a.tile_list = h->d_touched_list;
        a.list_n = reinterpret_cast<const unsigned*>(h->d_pre_misc);
        if (h->multi_fused) {
            dim3 grid((unsigned)std::min<int64_t>(h->n_tiles, (int64_t)h->n_sm * h->multi_ctas), 1u);
            k_score_bin_multi<2><<<grid, SBM_THREADS, sbm_smem_bytes(h->nb), h->stream>>>(a, h->n_tiles);
        } else {
            dim3 grid((unsigned)std::min<int64_t>(h->n_tiles, (int64_t)h->n_sm * h->score_ctas_per_sm), 1u);
            k_score_bin_tma<false, 2, 2><<<grid, SBT_THREADS, sbt_smem_bytes(false, 2), h->stream>>>(a, h->n_tiles);
        }
        BOSS_KERNEL_CHECK();
        h->launches += 2;
    } else if (h->multi_fused) {
        // barcodes: one CTA takes a tile through every barcode (row rules Q6/Q8), each counter read once
        dim3 grid((unsigned)std::min<int64_t>(h->n_tiles, (int64_t)h->n_sm * h->multi_ctas), 1u);
        k_score_bin_multi<0><<<grid, SBM_THREADS, sbm_smem_bytes(h->nb), h->stream>>>(a, h->n_tiles);
        BOSS_KERNEL_CHECK();
        h->launches++;
    } else {
        if (h->nb > 1) {
            k_rowflags<<<(unsigned)ceil_div(h->P / 4, 256), 256, 0, h->stream>>>(h->P / 4, h->nb, h->P, h->d_cov, h->d_rowflag);
            BOSS_KERNEL_CHECK();
            h->launches++;
        }
        if (h->score_kernel_ldg) {
            dim3 grid((unsigned)h->n_tiles, (unsigned)h->nb);
            if (h->nb > 1) k_score_bin<true><<<grid, SB_THREADS, 0, h->stream>>>(a);
            else k_score_bin<false><<<grid, SB_THREADS, 0, h->stream>>>(a);
        } else {
            dim3 grid((unsigned)std::min<int64_t>(h->n_tiles, (int64_t)h->n_sm * h->score_ctas_per_sm), (unsigned)h->nb);
            if (h->nb > 1) k_score_bin_tma<true, 2><<<grid, SBT_THREADS, sbt_smem_bytes(true, 2), h->stream>>>(a, h->n_tiles);
            else if (h->score_stages == 4) k_score_bin_tma<false, 4><<<grid, SBT_THREADS, sbt_smem_bytes(false, 4), h->stream>>>(a, h->n_tiles);
            else k_score_bin_tma<false, 2><<<grid, SBT_THREADS, sbt_smem_bytes(false, 2), h->stream>>>(a, h->n_tiles);
        }
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    EV_END(1);
    // per-tile depth totals and dropout counts -> bucket sums, n_dropout
    k_tile_reduce<<<(unsigned)ceil_div(h->n_tiles, 256), 256, 0, h->stream>>>(h->n_tiles, h->nb, (const TileDesc*)h->d_tiles, h->d_tile_cov,
                                                                             h->d_tile_drop, h->d_bucket_sum, &h->d_upd->n_dropout);
    BOSS_KERNEL_CHECK();
    h->launches++;
    EV_BEGIN(2);
    int64_t max_sw = 0;
    for (auto& S : h->segs) max_sw = std::max(max_sw, S.n_sw);
    dim3 gb((unsigned)ceil_div(max_sw * h->nb, 128), (unsigned)h->n_seg);
    k_buckets<<<gb, 128, 0, h->stream>>>(h->d_segs, h->n_seg, h->nb, p->bucket_threshold, h->d_bucket_sum, h->d_bucket_sw,
                                         &h->d_upd->switched_on);
    BOSS_KERNEL_CHECK();
    h->launches++;
    EV_END(2);
    return 0;
}

// The early half of a split score/bin pass (see include/bossgpu.h). _begin: called when a batch arrives, before anything
// is known about it — every tile is scored at once on a second stream from the counters as they are, with the dropout
// thresholds of the current depth totals, while the host picks records, packs bases and copies. bossgpu_prescore:
// called once the alignment intervals are known — marks the tiles the batch will write to. The update then scores
// only the marked tiles (after the scatter) plus the tiles of any contig whose threshold moved; every other tile's
// bins, depth total and dropout count from the early pass are exactly what the late pass would compute (same counters,
// same threshold). Tiles the scatter modifies while the early pass reads them yield garbage that the late pass
// overwrites (per-tile outputs are plain stores; any counter values index inside the table).
extern "C" int bossgpu_prescore_begin(bossgpu_handle* h) {
    H_CHECK(h);
    if (h->spec_state != 0) BOSS_CUDA(cudaStreamSynchronize(h->stream2));     // an earlier early pass was never consumed
    h->spec_state = 0;
    h->prescore_state = 0;
    if (!h->prescore_ok || h->score_kernel_ldg || h->score_stages != 2 || h->fused_open) return 0;
    cudaStream_t st = h->stream2;
    // whatever the main stream still has queued (a previous ingest, set/synth_coverage) comes first
    BOSS_CUDA(cudaEventRecord(h->ev_main, h->stream));
    BOSS_CUDA(cudaStreamWaitEvent(st, h->ev_main, 0));
    k_drop_thresholds<<<(unsigned)ceil_div(h->n_contigs_total, 128), 128, 0, st>>>(
        h->n_contigs_total, h->d_contig_len, h->nb, h->d_cov_total, h->d_drop_thr_spec);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaEventRecord(h->ev_pre_thr, st));       // the ingest may add to the totals after this point
    ScoreArgs a = score_args(h);
    a.drop_thr = h->d_drop_thr_spec;
    if (h->multi_fused) {
        dim3 grid((unsigned)std::min<int64_t>(h->n_tiles, (int64_t)h->n_sm * h->multi_ctas), 1u);
        k_score_bin_multi<0><<<grid, SBM_THREADS, sbm_smem_bytes(h->nb), st>>>(a, h->n_tiles);
    } else {
        dim3 grid((unsigned)std::min<int64_t>(h->n_tiles, (int64_t)h->n_sm * h->score_ctas_per_sm), 1u);
        k_score_bin_tma<false, 2><<<grid, SBT_THREADS, sbt_smem_bytes(false, 2), st>>>(a, h->n_tiles);
    }
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaEventRecord(h->ev_pre_done, st));
    h->launches += 2;
    h->spec_state = 1;
    return 0;
}

extern "C" int bossgpu_prescore(bossgpu_handle* h, int64_t n_reads, const int32_t* contig, const int64_t* tstart,
                                const int64_t* tend) {
    H_CHECK(h);
    if (n_reads < 0 || (n_reads > 0 && (!contig || !tstart || !tend))) return fail(BOSSGPU_EINVAL, "bad batch arrays");
    h->prescore_state = 0;
    if (h->spec_state != 1) return 0;                    // nothing in flight that the marks could save work for
    const size_t nr = (size_t)std::max<int64_t>(n_reads, 1);
    const size_t o_t0 = round_up(sizeof(int32_t) * nr, 16), o_t1 = o_t0 + sizeof(int64_t) * nr, total = o_t1 + sizeof(int64_t) * nr;
    if (total > h->pre_stage_bytes) {
        BOSS_CUDA(cudaStreamSynchronize(h->stream));     // an earlier announcement may still be copying out of the buffer
        TRY(ensure_pre_stage(h, total));
    }
    char* hs = (char*)h->pre_stage_h;
    int32_t* s_c = (int32_t*)hs; int64_t* s_t0 = (int64_t*)(hs + o_t0); int64_t* s_t1 = (int64_t*)(hs + o_t1);
    for (int64_t i = 0; i < n_reads; ++i) {
        if (contig[i] < 0 || contig[i] >= h->n_contigs_total) return fail(BOSSGPU_EINVAL, "read %lld: contig index out of range", (long long)i);
        s_c[i] = contig[i]; s_t0[i] = std::min(tstart[i], tend[i]); s_t1[i] = std::max(tstart[i], tend[i]);
    }
    cudaStream_t st = h->stream;                         // ahead of the scatter, in stream order
    BOSS_CUDA(cudaMemsetAsync(h->d_touched, 0, sizeof(uint32_t) * (size_t)(h->n_tiles / 32 + 2), st));
    BOSS_CUDA(cudaMemsetAsync(h->d_pre_misc, 0, sizeof(unsigned long long) * 2, st));
    if (n_reads > 0) {
        BOSS_CUDA(cudaMemcpyAsync(h->pre_stage_d, hs, total, cudaMemcpyHostToDevice, st));
        const char* ds = (const char*)h->pre_stage_d;
        k_mark_tiles<<<(unsigned)ceil_div(n_reads, 128), 128, 0, st>>>(
            n_reads, (const int32_t*)ds, (const int64_t*)(ds + o_t0), (const int64_t*)(ds + o_t1), h->d_seg_of_contig, h->d_segs,
            h->d_touched, h->d_touched_list, reinterpret_cast<unsigned*>(h->d_pre_misc));
        BOSS_KERNEL_CHECK();
        h->launches++;
    }
    h->pre_n_reads = n_reads;
    h->pre_hash = batch_hash(n_reads, contig, tstart, tend);
    h->prescore_state = 1;
    return 0;
}

static int ensure_debug(bossgpu_handle* h) {
    if (h->debug_bufs) return 0;
    TRY(dev_alloc(&h->d_smu, (size_t)h->nb * h->n_rows));
    TRY(dev_alloc(&h->d_expected, (size_t)h->nb * h->n_rows));
    h->debug_bufs = true;
    return 0;
}

static int phase1_smooth(bossgpu_handle* h, const bossgpu_update_params* p) {
    if (p->write_debug) TRY(ensure_debug(h));
    EV_BEGIN(3);
    const int64_t t = h->n_sm_tiles;
    SmoothArgs a;
    a.segs = h->d_segs; a.n_seg = h->n_seg; a.nb = h->nb; a.ds = h->d_ds; a.ds_len = h->ds_len;
    a.benefit = h->d_benefit;
    a.smu = p->write_debug ? h->d_smu : nullptr;
    a.expected = p->write_debug ? h->d_expected : nullptr;
    a.n_rows = h->n_rows;
    int wmax = 4, dmax = 1;
    for (int i = 0; i < NSTEPS; ++i) {
        a.mult[i] = p->mult[i];
        wmax = std::max(wmax, p->w[i]);
        dmax = std::max(dmax, p->w[i] - (i ? p->w[i - 1] : 0));
    }
    a.wmax = wmax;
    a.R0 = h->R0; a.target_rows = h->target_rows; a.upd = h->d_upd;
    // levels: as many power-of-two widths as the largest increment needs and shared memory allows (at least the
    // 4-bin level S_mu reads); up to 72 KB keeps three CTAs on an SM
    const size_t span = SM_TILE + 2 * (size_t)(wmax - 1);
    int levels = 3;
    while (levels < SM_MAX_LEVELS && (1 << levels) <= dmax && (size_t)(levels + 1) * span * sizeof(double) <= 72 * 1024) ++levels;
    const bool fits = (size_t)levels * span * sizeof(double) <= 200 * 1024 && span * (size_t)levels < (size_t)1 << 30;
    const bool planned = fits && plan_smoothing(p->w, levels, (int)span, a);
    dim3 grid((unsigned)t, (unsigned)h->nb);
    if (planned) {
        a.n_levels = levels;
        size_t smem = sizeof(double) * span * levels;
        k_smooth<<<grid, SM_THREADS, smem, h->stream>>>(a, h->d_sm_tile_start);
    } else {
        // very long staircase: bin-by-bin sums; windows go through scratch (behind the tile table)
        size_t smem = sizeof(double) * span;
        if (smem > 200 * 1024) return fail(BOSSGPU_EINVAL, "staircase window of %d bins needs %zu B of shared memory", wmax, smem);
        TRY(ensure_scratch(h, 256));
        int32_t* d_w = (int32_t*)h->scratch_d;
        BOSS_CUDA(cudaMemcpyAsync(d_w, p->w, sizeof(int32_t) * NSTEPS, cudaMemcpyHostToDevice, h->stream));
        a.n_levels = 1;
        k_smooth_direct<<<grid, SM_THREADS, smem, h->stream>>>(a, h->d_sm_tile_start, d_w);
    }
    BOSS_KERNEL_CHECK();
    h->launches++;
    EV_END(3);
    return 0;
}

static int upload_fhat(bossgpu_handle* h, const bossgpu_update_params* p) {
    if (p->fhat_from_counts) {
        const int64_t n = h->n_windows_total * 2;
        k_fhat_from_counts<<<(unsigned)ceil_div(n, 256), 256, 0, h->stream>>>(n, h->d_rs_counts, p->rs_alpha, p->rs_denom,
                                                                               p->rs_zero_value, h->d_fhat_w);
        BOSS_KERNEL_CHECK();
        h->launches++;
        h->have_fhat = true;
        return 0;
    }
    if (p->fhat_windows) {
        size_t bytes = sizeof(double) * 2 * h->n_windows_total;
        TRY(ensure_stage(h, bytes));
        memcpy(h->stage_h, p->fhat_windows, bytes);
        BOSS_CUDA(cudaMemcpyAsync(h->d_fhat_w, h->stage_h, bytes, cudaMemcpyHostToDevice, h->stream));
        h->have_fhat = true;
    }
    if (!h->have_fhat) return fail(BOSSGPU_ESTATE, "no F-hat uploaded yet");
    return 0;
}

static int phase2_hist(bossgpu_handle* h, const bossgpu_update_params* p) {
    EV_BEGIN(4);
    FhatGeom fg{h->n_windows_total, h->Tf_rows, h->target_rows};
    const int fsum_shift = 50;
    k_fhat_sum<<<(unsigned)ceil_div(h->n_windows_total, 256), 256, 0, h->stream>>>(fg, h->d_fhat_w, fsum_shift, h->d_upd);
    BOSS_KERNEL_CHECK();
    k_fhat_finish<<<1, 1, 0, h->stream>>>(fsum_shift, h->d_upd);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaMemsetAsync(h->d_hist, 0, sizeof(unsigned long long) * (3 * HBINS + 4), h->stream));
    HistArgs a;
    a.benefit = h->d_benefit; a.n_rows = h->n_rows; a.nb = h->nb; a.R0 = h->R0; a.M = h->M_rows; a.target = h->target_rows;
    a.fg = fg; a.fw = h->d_fhat_w; a.shift = h->fhat_shift; a.ushift = h->ubar_shift; a.hist = h->d_hist; a.upd = h->d_upd; a.codes = h->d_codes;
    const int64_t n_groups = (h->R0 + h->n_rows - 1) / HIST_GROUP - h->R0 / HIST_GROUP + 1;     // groups of the global row axis
    // persistent grid: exactly what the SMs hold at once (a grid of 10 CTAs per SM ran as 1.7 waves with 6 resident)
    static int hist_resident = 0;
    if (hist_resident == 0) {
        BOSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&hist_resident, k_hist, HIST_THREADS, 0));
        hist_resident = std::max(1, hist_resident);
    }
    const int64_t hist_ctas = std::max<int64_t>(1, (int64_t)h->n_sm * hist_resident / h->nb);
    unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_groups, HIST_THREADS), hist_ctas));
    k_hist<<<dim3(gx, (unsigned)h->nb), HIST_THREADS, 0, h->stream>>>(a);
    BOSS_KERNEL_CHECK();
    h->launches += 3;
    EV_END(4);
    return 0;
}

static int phase3_threshold(bossgpu_handle* h, const bossgpu_update_params* p) {
    EV_BEGIN(5);
    k_threshold<<<1, THR_THREADS, 0, h->stream>>>(h->d_hist, h->fhat_shift, h->ubar_shift, p->tc, h->d_upd);
    BOSS_KERNEL_CHECK();
    h->launches++;
    EV_END(5);
    return 0;
}

static int phase4_distribute(bossgpu_handle* h, const uint8_t* const* mask_ptrs) {
    EV_BEGIN(6);
    DistArgs a;
    a.segs = h->d_segs; a.srow_start = h->d_srow_start; a.n_seg = h->n_seg; a.nb = h->nb; a.benefit = h->d_benefit;
    a.codes = h->d_codes;
    a.n_rows = h->n_rows; a.R0 = h->R0; a.D0 = h->D0; a.mask_ptrs = mask_ptrs; a.bucket_sw = h->d_bucket_sw;
    a.shard_row_start = h->d_shard_row_start; a.n_shards = h->n_shards;
    a.strat = h->d_strat; a.strat_host = h->h_strat_dev; a.shift = h->strat_shift;
    a.n_srows = h->n_srows; a.upd = h->d_upd; a.seg_accept = h->d_seg_accept;
    BOSS_CUDA(cudaMemsetAsync(h->d_seg_accept, 0, sizeof(unsigned long long) * 2 * h->n_seg, h->stream));
    const int64_t n_vec = ceil_div(h->n_srows * 2 * h->nb + h->strat_shift, DIST_VEC);
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_vec, DIST_THREADS), (int64_t)h->n_sm * 16));
    if (h->nb == 1) k_distribute<true><<<grid, DIST_THREADS, 0, h->stream>>>(a);
    else k_distribute<false><<<grid, DIST_THREADS, 0, h->stream>>>(a);
    BOSS_KERNEL_CHECK();
    h->launches++;
    // the kernel has refreshed the host mirror where masks changed; only the per-segment accept counts are copied
    BOSS_CUDA(cudaMemcpyAsync(h->h_seg_accept, h->d_seg_accept, sizeof(unsigned long long) * 2 * h->n_seg, cudaMemcpyDeviceToHost, h->stream));
    EV_END(6);
    return 0;
}

static int fetch_result(bossgpu_handle* h, bossgpu_update_result* r) {
    BOSS_CUDA(cudaMemcpyAsync(h->h_upd, h->d_upd, sizeof(UpdateDev), cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaMemcpyAsync(h->h_bucket_sw, h->d_bucket_sw, (size_t)h->n_sw * h->nb, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    h->last = *h->h_upd;
    if (h->last.switched_on) h->sticky_on = true;
    for (int i = 0; i < BOSSGPU_N_TIMERS; ++i)
        if (h->ev_valid[i]) cudaEventElapsedTime(&h->ms[i], h->ev[2 * i], h->ev[2 * i + 1]);
    if (r) {
        memset(r, 0, sizeof *r);
        r->switched_on = h->last.switched_on;
        r->strat_size = h->last.strat_size;
        r->threshold = h->last.threshold;
        r->normaliser = h->last.normaliser;
        r->ubar0 = h->last.ubar0;
        r->fhat_sum = h->last.fhat_sum;
        r->n_nonzero = (int64_t)h->last.n_nonzero;
        r->n_dropout = (int64_t)h->last.n_dropout;
        r->mirror_bytes = (int64_t)h->last.mirror_bytes;
        for (int sg = 0; sg < h->n_seg; ++sg) {
            r->n_accept[0] += (int64_t)h->h_seg_accept[2 * sg];
            r->n_accept[1] += (int64_t)h->h_seg_accept[2 * sg + 1];
        }
    }
    TRY(check_ingest_error(h));
    return 0;
}

extern "C" int bossgpu_update(bossgpu_handle* h, const bossgpu_update_params* p, bossgpu_update_result* r) {
    H_CHECK(h);
    TRY(validate_params(p));
    for (const auto& S : h->segs)
        if (S.start != 0 || !S.is_tail) return fail(BOSSGPU_ESTATE, "bossgpu_update needs whole-contig segments; use the phase API");
    EV_BEGIN(7);
    TRY(phase0_scores(h, p));
    if (h->sticky_on) {
        // Bucket switches are sticky (reference.py:199-211): once an update has seen one on, every later update does. The
        // whole update is enqueued without a host round trip; the kernels of the strategy half look at the device-side flags
        // themselves (all-zero benefit, missing time_cost: the masks stay as they are) and the host checks them at the end.
        TRY(upload_fhat(h, p));
        TRY(phase1_smooth(h, p));
        TRY(phase2_hist(h, p));
        TRY(phase3_threshold(h, p));
        TRY(phase4_distribute(h, nullptr));
        EV_END(7);
        TRY(fetch_result(h, r));
        if (h->last.switched_on && h->last.empty == 2)
            return fail(BOSSGPU_ENOTC, "a bucket is on but there is no time_cost yet (no read length has been observed)");
        if (h->last.switched_on && h->last.empty)
            return fail(BOSSGPU_EEMPTY, "all benefits are zero: upstream np.max of an empty array raises ValueError");
        return 0;
    }
    // The strategy half only runs once some bucket is on (core.py:172). The switch lives on the device;
    // reading it costs one small sync, which also bounds how much work is queued behind an idle update.
    BOSS_CUDA(cudaMemcpyAsync(&h->h_upd->switched_on, &h->d_upd->switched_on, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_upd->switched_on) {
        if (p->tc != p->tc) {                  // Q14: upstream raises before find_strat_thread, every strategy stays as it is
            EV_END(7);
            fetch_result(h, r);
            return fail(BOSSGPU_ENOTC, "a bucket is on but there is no time_cost yet (no read length has been observed)");
        }
        TRY(upload_fhat(h, p));
        TRY(phase1_smooth(h, p));
        TRY(phase2_hist(h, p));
        TRY(phase3_threshold(h, p));
        // upstream raises on an all-zero benefit before touching any strategy (sequences.py:588)
        BOSS_CUDA(cudaMemcpyAsync(&h->h_upd->empty, &h->d_upd->empty, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        BOSS_CUDA(cudaStreamSynchronize(h->stream));
        if (h->h_upd->empty) {
            EV_END(7);
            fetch_result(h, r);
            return fail(BOSSGPU_EEMPTY, "all benefits are zero: upstream np.max of an empty array raises ValueError");
        }
        TRY(phase4_distribute(h, nullptr));
    }
    EV_END(7);
    return fetch_result(h, r);
}

static int pack_own_mask(bossgpu_handle* h, uint8_t* dst);
static int point_masks_at_own_copy(bossgpu_handle* h);

extern "C" int bossgpu_update_phase(bossgpu_handle* h, int phase, const bossgpu_update_params* p, bossgpu_update_result* r) {
    H_CHECK(h);
    TRY(validate_params(p));
    // phase 4 may follow phase 0 directly: with every bucket still off the strategy half is skipped (core.py:172)
    if (phase != 0 && phase != h->phase_done + 1 && !(phase == 4 && h->phase_done == 0))
        return fail(BOSSGPU_ESTATE, "phase %d after phase %d", phase, h->phase_done);
    switch (phase) {
        case 0: EV_BEGIN(7); TRY(phase0_scores(h, p)); break;
        case 1: TRY(upload_fhat(h, p)); TRY(phase1_smooth(h, p)); break;
        case 2: TRY(phase2_hist(h, p)); break;
        case 3: TRY(phase3_threshold(h, p)); TRY(pack_own_mask(h, h->d_mask_all + (size_t)h->shard_index * h->mask_stride)); break;
        case 4:
            // upstream raises on an all-zero benefit before touching any strategy (sequences.py:588); the flag is
            // the same on every shard because the histogram it derives from has been allreduced
            BOSS_CUDA(cudaMemcpyAsync(h->h_upd, h->d_upd, sizeof(UpdateDev), cudaMemcpyDeviceToHost, h->stream));
            BOSS_CUDA(cudaStreamSynchronize(h->stream));
            if (!h->h_upd->switched_on) {          // (allreduced) switch still off: leave every strategy as it is
                EV_END(7);
                TRY(fetch_result(h, r));
                break;
            }
            if (h->phase_done != 3) return fail(BOSSGPU_ESTATE, "a bucket is on but phases 1-3 were skipped");
            if (h->h_upd->empty) {
                const int why = h->h_upd->empty;
                EV_END(7);
                h->phase_done = -1;
                fetch_result(h, r);
                if (why == 2) return fail(BOSSGPU_ENOTC, "a bucket is on but there is no time_cost yet (no read length has been observed)");
                return fail(BOSSGPU_EEMPTY, "all benefits are zero: upstream np.max of an empty array raises ValueError");
            }
            TRY(phase4_distribute(h, h->n_shards > 1 ? h->d_mask_ptrs : nullptr));
            EV_END(7);
            TRY(fetch_result(h, r));
            break;
        default: return fail(BOSSGPU_EINVAL, "unknown phase %d", phase);
    }
    h->phase_done = phase == 4 ? -1 : phase;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// multi-shard support: exchange buffers
// ------------------------------------------------------------------------------------------------
extern "C" int bossgpu_set_shards(bossgpu_handle* h, int32_t n_shards, int32_t shard_index, const int64_t* row_start) {
    H_CHECK(h);
    if (n_shards < 1 || shard_index < 0 || shard_index >= n_shards || !row_start) return fail(BOSSGPU_EINVAL, "bad shard description");
    if (row_start[shard_index] != h->R0 || row_start[shard_index + 1] != h->R0 + h->n_rows)
        return fail(BOSSGPU_EINVAL, "shard row table disagrees with this handle's segments");
    h->n_shards = n_shards;
    h->shard_index = shard_index;
    h->shard_row_start.assign(row_start, row_start + n_shards + 1);
    int64_t max_rows = 0;
    for (int i = 0; i < n_shards; ++i) max_rows = std::max(max_rows, row_start[i + 1] - row_start[i]);
    h->mask_stride = round_up(ceil_div(max_rows * 2 * h->nb, 8), 16);
    if (h->d_mask_all) cudaFree(h->d_mask_all);
    if (h->d_shard_row_start) cudaFree(h->d_shard_row_start);
    TRY(dev_alloc(&h->d_mask_all, (size_t)h->mask_stride * n_shards));
    TRY(dev_alloc(&h->d_shard_row_start, (size_t)n_shards + 1));
    if (h->d_mask_ptrs) cudaFree(h->d_mask_ptrs);
    TRY(dev_alloc(&h->d_mask_ptrs, (size_t)n_shards));
    TRY(point_masks_at_own_copy(h));
    // exchange block of the peer-memory fabric (fabric.cuh); mapped by the peers after bossgpu_fabric_attach
    if (h->d_fabric) cudaFree(h->d_fabric);
    if (h->d_peer_ptrs) cudaFree(h->d_peer_ptrs);
    if (h->d_fab_mask_ptrs) cudaFree(h->d_fab_mask_ptrs);
    h->d_fabric = nullptr; h->d_peer_ptrs = nullptr; h->d_fab_mask_ptrs = nullptr; h->fabric_attached = false; h->fabric_epoch = 0;
    if (n_shards <= FAB_MAX_SHARDS) {
        const FabricLayout FL = fabric_layout(n_shards, h->nb, h->halo_bins, h->mask_stride);
        TRY(dev_alloc(&h->d_fabric, FL.bytes));
        h->fabric_bytes = FL.bytes;
        TRY(dev_alloc(&h->d_peer_ptrs, (size_t)n_shards));
        TRY(dev_alloc(&h->d_fab_mask_ptrs, (size_t)n_shards));
    }
    BOSS_CUDA(cudaMemcpy(h->d_shard_row_start, row_start, sizeof(int64_t) * (n_shards + 1), cudaMemcpyHostToDevice));
    // halo staging: [left send | right send | left recv | right recv], each halo_bins * nb doubles
    if (h->d_halo) cudaFree(h->d_halo);
    TRY(dev_alloc(&h->d_halo, (size_t)4 * std::max(1, h->halo_bins) * h->nb));
    return 0;
}

// phase API: every shard's mask is gathered (by the caller's collective) into this handle's d_mask_all
static int point_masks_at_own_copy(bossgpu_handle* h) {
    std::vector<const uint8_t*> ptrs((size_t)h->n_shards);
    for (int s = 0; s < h->n_shards; ++s) ptrs[s] = h->d_mask_all + (size_t)s * h->mask_stride;
    BOSS_CUDA(cudaMemcpy(h->d_mask_ptrs, ptrs.data(), sizeof(uint8_t*) * h->n_shards, cudaMemcpyHostToDevice));
    return 0;
}

static int pack_own_mask(bossgpu_handle* h, uint8_t* dst) {
    if (h->n_shards <= 1) return 0;
    int64_t n_bits = h->n_rows * 2 * h->nb;
    k_pack_mask<<<(unsigned)ceil_div(ceil_div(n_bits, 8), 256), 256, 0, h->stream>>>(
        h->d_benefit, h->d_codes, h->n_rows, h->nb, h->R0, h->target_rows, h->d_upd, dst, n_bits);
    BOSS_KERNEL_CHECK();
    h->launches++;
    return 0;
}

// copy the edge bins of split contigs into / out of the halo staging buffers
extern "C" int bossgpu_halo_pack(bossgpu_handle* h) {
    H_CHECK(h);
    if (h->halo_bins <= 0 || !h->d_halo) return 0;
    const SegDev& F = h->segs.front();
    const SegDev& L = h->segs.back();
    const int hb = h->halo_bins;
    BOSS_CUDA(cudaMemsetAsync(h->d_halo, 0, sizeof(double) * 2 * hb * h->nb, h->stream));
    for (int b = 0; b < h->nb; ++b) {
        if (F.start > 0) {      // my first bins go to the left neighbour (its right halo)
            int64_t n = std::min<int64_t>(hb, F.n_bins);
            BOSS_CUDA(cudaMemcpyAsync(h->d_halo + (size_t)b * hb, h->d_ds + (size_t)b * h->ds_len + F.ds_off, sizeof(double) * n,
                                      cudaMemcpyDeviceToDevice, h->stream));
        }
        if (!L.is_tail) {       // my last bins go to the right neighbour (its left halo), right-aligned
            int64_t n = std::min<int64_t>(hb, L.n_bins);
            BOSS_CUDA(cudaMemcpyAsync(h->d_halo + (size_t)(h->nb + b) * hb + (hb - n),
                                      h->d_ds + (size_t)b * h->ds_len + L.ds_off + L.n_bins - n, sizeof(double) * n,
                                      cudaMemcpyDeviceToDevice, h->stream));
        }
    }
    return 0;
}

extern "C" int bossgpu_halo_unpack(bossgpu_handle* h) {
    H_CHECK(h);
    if (h->halo_bins <= 0 || !h->d_halo) return 0;
    const SegDev& F = h->segs.front();
    const SegDev& L = h->segs.back();
    const int hb = h->halo_bins;
    double* recv_l = h->d_halo + (size_t)2 * hb * h->nb;     // from the left neighbour: its last bins, right-aligned
    double* recv_r = recv_l + (size_t)hb * h->nb;            // from the right neighbour: its first bins
    for (int b = 0; b < h->nb; ++b) {
        if (F.start > 0)
            BOSS_CUDA(cudaMemcpyAsync(h->d_ds + (size_t)b * h->ds_len + F.ds_off - hb, recv_l + (size_t)b * hb, sizeof(double) * hb,
                                      cudaMemcpyDeviceToDevice, h->stream));
        if (!L.is_tail)
            BOSS_CUDA(cudaMemcpyAsync(h->d_ds + (size_t)b * h->ds_len + L.ds_off + L.n_bins, recv_r + (size_t)b * hb, sizeof(double) * hb,
                                      cudaMemcpyDeviceToDevice, h->stream));
    }
    return 0;
}

extern "C" int bossgpu_exchange_buffer(bossgpu_handle* h, int which, void** dev_ptr, size_t* bytes) {
    H_CHECK(h);
    if (!dev_ptr || !bytes) return fail(BOSSGPU_EINVAL, "null out pointer");
    const size_t hb = (size_t)std::max(1, h->halo_bins) * h->nb * sizeof(double);
    switch (which) {
        case BOSSGPU_BUF_SWITCH: *dev_ptr = &h->d_upd->switched_on; *bytes = sizeof(int32_t); break;
        case BOSSGPU_BUF_NORM: *dev_ptr = &h->d_upd->norm_bits; *bytes = sizeof(unsigned long long); break;
        case BOSSGPU_BUF_HIST: *dev_ptr = h->d_hist; *bytes = sizeof(unsigned long long) * (3 * HBINS + 4); break;
        case BOSSGPU_BUF_MASK:
            if (!h->d_mask_all) return fail(BOSSGPU_ESTATE, "bossgpu_set_shards not called");
            *dev_ptr = h->d_mask_all; *bytes = (size_t)h->mask_stride * h->n_shards; break;
        case BOSSGPU_BUF_HALO_SEND:
            if (!h->d_halo) return fail(BOSSGPU_ESTATE, "bossgpu_set_shards not called");
            *dev_ptr = h->d_halo; *bytes = 2 * hb; break;
        case BOSSGPU_BUF_HALO_RECV:
            if (!h->d_halo) return fail(BOSSGPU_ESTATE, "bossgpu_set_shards not called");
            *dev_ptr = (char*)h->d_halo + 2 * hb; *bytes = 2 * hb; break;
        case BOSSGPU_BUF_COV_TOTAL: *dev_ptr = h->d_cov_total; *bytes = sizeof(unsigned long long) * h->n_contigs_total; break;
        case BOSSGPU_BUF_STRAT: *dev_ptr = h->d_strat; *bytes = (size_t)h->n_srows * 2 * h->nb; break;
        default: return fail(BOSSGPU_EINVAL, "unknown exchange buffer %d", which);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// peer-memory fabric: the exchanges of the sharded update done by the GPUs (fabric.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int bossgpu_fabric_info(bossgpu_handle* h, void** dev_ptr, size_t* bytes, unsigned char ipc_handle[64]) {
    H_CHECK(h);
    if (!h->d_fabric) return fail(BOSSGPU_ESTATE, "no exchange block: call bossgpu_set_shards first (at most %d shards)", FAB_MAX_SHARDS);
    if (dev_ptr) *dev_ptr = h->d_fabric;
    if (bytes) *bytes = h->fabric_bytes;
    if (ipc_handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t mh;
        BOSS_CUDA(cudaIpcGetMemHandle(&mh, h->d_fabric));
        memcpy(ipc_handle, &mh, 64);
    }
    return 0;
}

extern "C" int bossgpu_ipc_open(int device, const unsigned char ipc_handle[64], void** dev_ptr) {
    if (!ipc_handle || !dev_ptr) return fail(BOSSGPU_EINVAL, "null argument");
    BOSS_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t mh;
    memcpy(&mh, ipc_handle, 64);
    *dev_ptr = nullptr;
    BOSS_CUDA(cudaIpcOpenMemHandle(dev_ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int bossgpu_ipc_close(int device, void* dev_ptr) {
    if (!dev_ptr) return 0;
    BOSS_CUDA(cudaSetDevice(device));
    BOSS_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

// CUDA loads a kernel's code on its first launch (lazy module loading, the default since 12.2), and loading
// synchronises the context. Inside a fused update that is fatal for shards sharing one context: the host would
// block loading the next kernel while an exchange kernel of this shard spins for a peer whose work the host has not
// enqueued yet. So every kernel an update can launch is loaded before the first exchange is ever enqueued.
template <typename K>
static int preload_kernel(K kernel, const char* name) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(BOSSGPU_ECUDA, "loading %s failed: %s", name, cudaGetErrorString(e)); }
    return 0;
}
#define PRELOAD(...) TRY(preload_kernel(__VA_ARGS__, #__VA_ARGS__))

static int preload_update_kernels() {
    PRELOAD(k_drop_thresholds); PRELOAD(k_rowflags); PRELOAD(k_buckets);
    PRELOAD(k_score_bin<false>); PRELOAD(k_score_bin<true>);
    PRELOAD(k_score_bin_tma<false, 2>); PRELOAD(k_score_bin_tma<false, 4>); PRELOAD(k_score_bin_tma<true, 2>);
    PRELOAD(k_score_bin_tma<false, 2, 2>); PRELOAD(k_score_bin_multi<0>); PRELOAD(k_score_bin_multi<2>); PRELOAD(k_mark_tiles); PRELOAD(k_mark_changed); PRELOAD(k_tile_reduce);
    PRELOAD(k_fhat_from_counts); PRELOAD(k_fhat_sum); PRELOAD(k_fhat_finish);
    PRELOAD(k_smooth); PRELOAD(k_smooth_direct); PRELOAD(k_hist); PRELOAD(k_threshold); PRELOAD(k_pack_mask);
    PRELOAD(k_distribute<true>); PRELOAD(k_distribute<false>);
    PRELOAD(k_fabric_switch_halo); PRELOAD(k_fabric_norm); PRELOAD(k_fabric_hist); PRELOAD(k_fabric_mask_ready);
    return 0;
}

extern "C" int bossgpu_fabric_attach(bossgpu_handle* h, int32_t n_shards, const uint64_t* peer_ptrs, double timeout_s) {
    H_CHECK(h);
    if (!h->d_fabric) return fail(BOSSGPU_ESTATE, "no exchange block: call bossgpu_set_shards first");
    if (n_shards != h->n_shards || !peer_ptrs) return fail(BOSSGPU_EINVAL, "peer table must hold %d pointers", h->n_shards);
    if (peer_ptrs[h->shard_index] != (uint64_t)(uintptr_t)h->d_fabric)
        return fail(BOSSGPU_EINVAL, "peer table entry %d must be this handle's own exchange block", h->shard_index);
    const FabricLayout FL = fabric_layout(h->n_shards, h->nb, h->halo_bins, h->mask_stride);
    std::vector<char*> pp((size_t)n_shards);
    std::vector<const uint8_t*> mp((size_t)n_shards);
    for (int s = 0; s < n_shards; ++s) {
        if (!peer_ptrs[s]) return fail(BOSSGPU_EINVAL, "null peer pointer for shard %d", s);
        pp[s] = (char*)(uintptr_t)peer_ptrs[s];
        mp[s] = (const uint8_t*)(pp[s] + FL.o_mask);
    }
    BOSS_CUDA(cudaMemcpy(h->d_peer_ptrs, pp.data(), sizeof(char*) * n_shards, cudaMemcpyHostToDevice));
    BOSS_CUDA(cudaMemcpy(h->d_fab_mask_ptrs, mp.data(), sizeof(uint8_t*) * n_shards, cudaMemcpyHostToDevice));
    BOSS_CUDA(cudaMemset(h->d_fabric, 0, FL.o_sw));           // flags: no epoch seen yet
    if (timeout_s > 0) h->fabric_timeout_ns = (unsigned long long)(timeout_s * 1e9);
    TRY(preload_update_kernels());
    h->fabric_epoch = 0;
    h->fabric_attached = true;
    return 0;
}

static FabricArgs fabric_args(bossgpu_handle* h) {
    FabricArgs a;
    a.peer = h->d_peer_ptrs;
    a.L = fabric_layout(h->n_shards, h->nb, h->halo_bins, h->mask_stride);
    a.n = h->n_shards; a.me = h->shard_index;
    a.epoch = h->fabric_epoch;
    a.timeout_ns = h->fabric_timeout_ns;
    a.upd = h->d_upd;
    a.ds = h->d_ds; a.ds_len = h->ds_len; a.nb = h->nb; a.halo_bins = h->halo_bins;
    const SegDev& F = h->segs.front();
    const SegDev& L = h->segs.back();
    a.first_ds_off = F.ds_off; a.first_n_bins = F.n_bins;
    a.last_ds_off = L.ds_off; a.last_n_bins = L.n_bins;
    a.send_left = F.start > 0; a.send_right = !L.is_tail;
    a.hist = h->d_hist;
    return a;
}

// All kernels of one sharded update, exchanges included, enqueued without a host round trip. Kernels of the
// strategy half return at once while every bucket of every shard is still off (core.py:172).
extern "C" int bossgpu_update_fused_begin(bossgpu_handle* h, const bossgpu_update_params* p) {
    H_CHECK(h);
    TRY(validate_params(p));
    if (!h->fabric_attached) return fail(BOSSGPU_ESTATE, "bossgpu_fabric_attach has not been called");
    if (h->fused_open) return fail(BOSSGPU_ESTATE, "previous bossgpu_update_fused_begin has no matching _end");
    if (!p->fhat_from_counts && !p->fhat_windows && !h->have_fhat) return fail(BOSSGPU_ESTATE, "no F-hat uploaded yet");
    // Everything that can fail without the GPU's help happens BEFORE the first exchange kernel is enqueued: a shard that
    // gave up half-way would leave its peers spinning at the next exchange until the fabric timeout. validate_params has
    // bounded the smoothing tile; the buffers the phases may want are allocated here.
    if (p->write_debug) TRY(ensure_debug(h));
    if (p->fhat_windows) TRY(ensure_stage(h, sizeof(double) * 2 * h->n_windows_total));
    TRY(ensure_scratch(h, 256));
    h->fabric_epoch++;
    const FabricArgs fa = fabric_args(h);
    EV_BEGIN(7);
    TRY(phase0_scores(h, p));
    k_fabric_switch_halo<<<1, FAB_THREADS, 0, h->stream>>>(fa);
    BOSS_KERNEL_CHECK();
    TRY(upload_fhat(h, p));
    TRY(phase1_smooth(h, p));
    k_fabric_norm<<<1, FAB_THREADS, 0, h->stream>>>(fa);
    BOSS_KERNEL_CHECK();
    TRY(phase2_hist(h, p));
    k_fabric_hist<<<1, FAB_THREADS, 0, h->stream>>>(fa);
    BOSS_KERNEL_CHECK();
    TRY(phase3_threshold(h, p));
    TRY(pack_own_mask(h, (uint8_t*)(h->d_fabric + fa.L.o_mask)));
    k_fabric_mask_ready<<<1, FAB_THREADS, 0, h->stream>>>(fa);
    BOSS_KERNEL_CHECK();
    h->launches += 4;
    TRY(phase4_distribute(h, h->d_fab_mask_ptrs));
    EV_END(7);
    h->fused_open = true;
    h->phase_done = -1;
    return 0;
}

extern "C" int bossgpu_update_fused_end(bossgpu_handle* h, bossgpu_update_result* r) {
    H_CHECK(h);
    if (!h->fused_open) return fail(BOSSGPU_ESTATE, "bossgpu_update_fused_end without _begin");
    h->fused_open = false;
    TRY(fetch_result(h, r));
    if (h->last.fabric_err) return fail(BOSSGPU_EPEER, "shard %d of %d: a peer did not reach an exchange step of update %u within %.1f s",
                                        h->shard_index, h->n_shards, h->fabric_epoch, h->fabric_timeout_ns * 1e-9);
    if (h->last.switched_on && h->last.empty == 2)
        return fail(BOSSGPU_ENOTC, "a bucket is on but there is no time_cost yet (no read length has been observed)");
    if (h->last.switched_on && h->last.empty)
        return fail(BOSSGPU_EEMPTY, "all benefits are zero: upstream np.max of an empty array raises ValueError");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// getters / setters
// ------------------------------------------------------------------------------------------------
#define SEG_CHECK(h, seg)                                                                         \
    do {                                                                                          \
        if ((seg) < 0 || (seg) >= (h)->n_seg) return fail(BOSSGPU_EINVAL, "segment %d out of range", (int)(seg)); \
    } while (0)

extern "C" int64_t bossgpu_strat_rows(bossgpu_handle* h, int32_t seg) {
    if (!h) return -1;
    if (seg < 0) return h->n_srows;
    if (seg >= h->n_seg) return -1;
    return h->segs[seg].n_srows;
}

extern "C" int bossgpu_get_strat(bossgpu_handle* h, int32_t seg, uint8_t* out, int64_t out_bytes) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    int64_t n = S.n_srows * 2 * h->nb;
    if (!out || out_bytes != n) return fail(BOSSGPU_EINVAL, "strat buffer must hold %lld bytes", (long long)n);
    BOSS_CUDA(cudaMemcpyAsync(out, h->d_strat + (size_t)S.srow_off * 2 * h->nb, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_strat_all(bossgpu_handle* h, uint8_t* out, int64_t out_bytes) {
    H_CHECK(h);
    int64_t n = h->n_srows * 2 * h->nb;
    if (!out || out_bytes != n) return fail(BOSSGPU_EINVAL, "strat buffer must hold %lld bytes", (long long)n);
    BOSS_CUDA(cudaMemcpyAsync(out, h->d_strat, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_strat_packed(bossgpu_handle* h, uint8_t* out, int64_t out_bytes) {
    H_CHECK(h);
    int64_t n = h->n_srows * 2 * h->nb;
    int64_t nbytes = ceil_div(n, 8);
    if (!out || out_bytes != nbytes) return fail(BOSSGPU_EINVAL, "packed strat buffer must hold %lld bytes", (long long)nbytes);
    TRY(ensure_scratch(h, (size_t)nbytes));
    k_pack_strat<<<(unsigned)ceil_div(nbytes, 256), 256, 0, h->stream>>>(h->d_strat, n, (uint8_t*)h->scratch_d);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaMemcpyAsync(out, h->scratch_d, (size_t)nbytes, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_coverage(bossgpu_handle* h, int32_t seg, uint16_t* out, int64_t out_elems) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    int64_t n = S.len * 5 * h->nb;
    if (!out || out_elems != n) return fail(BOSSGPU_EINVAL, "coverage buffer must hold %lld elements", (long long)n);
    TRY(ensure_scratch(h, sizeof(uint16_t) * (size_t)n));
    k_cov_to_ref_layout<<<(unsigned)ceil_div(n, 256), 256, 0, h->stream>>>(S, h->nb, h->P, h->d_cov, (uint16_t*)h->scratch_d);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaMemcpyAsync(out, h->scratch_d, sizeof(uint16_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_set_coverage(bossgpu_handle* h, int32_t seg, const uint16_t* in, int64_t in_elems) {
    H_CHECK(h);
    TRY(prescore_invalidate(h));
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    int64_t n = S.len * 5 * h->nb;
    if (!in || in_elems != n) return fail(BOSSGPU_EINVAL, "coverage buffer must hold %lld elements", (long long)n);
    TRY(ensure_scratch(h, sizeof(uint16_t) * (size_t)n + 16));
    BOSS_CUDA(cudaMemcpyAsync(h->scratch_d, in, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    k_cov_from_ref_layout<<<(unsigned)ceil_div(n, 256), 256, 0, h->stream>>>(S, h->nb, h->P, (const uint16_t*)h->scratch_d, h->d_cov);
    BOSS_KERNEL_CHECK();
    // the contig depth total is state derived from the counters: recompute it for every contig of this shard
    BOSS_CUDA(cudaMemsetAsync(h->d_cov_total, 0, sizeof(unsigned long long) * h->n_contigs_total, h->stream));
    for (const auto& T : h->segs) {
        k_depth_total<<<(unsigned)std::min<int64_t>(ceil_div(T.len, 256), 148 * 8), 256, 0, h->stream>>>(T, h->nb, h->P, h->d_cov, h->d_cov_total);
        BOSS_KERNEL_CHECK();
    }
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_scores(bossgpu_handle* h, int32_t seg, double* scores, double* entropy, int64_t out_elems) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    int64_t n = S.len * h->nb;
    if (out_elems != n) return fail(BOSSGPU_EINVAL, "score buffers must hold %lld elements", (long long)n);
    TRY(ensure_scratch(h, sizeof(double) * (size_t)n * 2));
    double* ds = (double*)h->scratch_d;
    double* de = ds + n;
    k_materialise_scores<<<(unsigned)ceil_div(S.len, 256), 256, 0, h->stream>>>(S, h->nb, h->P, h->d_ref, h->d_cov, h->d_table, h->d_etable,
                                                                                 h->d_drop_thr, h->score0, h->ent0, ds, de);
    BOSS_KERNEL_CHECK();
    if (scores) BOSS_CUDA(cudaMemcpyAsync(scores, ds, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    if (entropy) BOSS_CUDA(cudaMemcpyAsync(entropy, de, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_scores_ds(bossgpu_handle* h, int32_t seg, double* out, int64_t out_elems) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    int64_t n = S.n_bins * h->nb;
    if (!out || out_elems != n) return fail(BOSSGPU_EINVAL, "scores_ds buffer must hold %lld elements", (long long)n);
    TRY(ensure_scratch(h, sizeof(double) * (size_t)n));
    k_gather_bins<<<(unsigned)ceil_div(n, 256), 256, 0, h->stream>>>(h->d_ds, h->ds_len, S.ds_off, S.n_bins, h->nb, 1, (double*)h->scratch_d);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaMemcpyAsync(out, h->scratch_d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_benefit(bossgpu_handle* h, int32_t seg, double* additional, double* smu, double* expected, int64_t out_elems) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    int64_t n = S.n_bins * 2 * h->nb;
    if (out_elems != n) return fail(BOSSGPU_EINVAL, "benefit buffers must hold %lld elements", (long long)n);
    if ((smu || expected) && !h->debug_bufs) return fail(BOSSGPU_ESTATE, "smu/expected need write_debug in the last update");
    TRY(ensure_scratch(h, sizeof(double) * (size_t)n));
    const double2* srcs[3] = {h->d_benefit, h->d_smu, h->d_expected};
    double* dsts[3] = {additional, smu, expected};
    for (int i = 0; i < 3; ++i) {
        if (!dsts[i]) continue;
        k_gather_bins<<<(unsigned)ceil_div(n, 256), 256, 0, h->stream>>>((const double*)srcs[i], h->n_rows * 2, S.row_off * 2, S.n_bins * 2,
                                                                         h->nb, 2, (double*)h->scratch_d);
        BOSS_KERNEL_CHECK();
        BOSS_CUDA(cudaMemcpyAsync(dsts[i], h->scratch_d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        BOSS_CUDA(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

extern "C" int bossgpu_get_buckets(bossgpu_handle* h, int32_t seg, uint8_t* switches, int64_t n, uint8_t* switched_on) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    if (n != S.n_sw * h->nb) return fail(BOSSGPU_EINVAL, "bucket buffer must hold %lld entries", (long long)(S.n_sw * h->nb));
    std::vector<uint8_t> tmp((size_t)n);
    BOSS_CUDA(cudaMemcpyAsync(tmp.data(), h->d_bucket_sw + (size_t)S.sw_off * h->nb, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    if (switches) memcpy(switches, tmp.data(), (size_t)n);
    if (switched_on) {
        // reference.py:203-207: once any bucket of any barcode is on, the whole contig is flagged
        bool any = false;
        for (uint8_t v : tmp) any |= v != 0;
        for (int b = 0; b < h->nb; ++b) switched_on[b] = any ? 1 : 0;
    }
    return 0;
}

extern "C" int bossgpu_set_buckets(bossgpu_handle* h, int32_t seg, const uint8_t* switches, int64_t n) {
    H_CHECK(h);
    SEG_CHECK(h, seg);
    const SegDev& S = h->segs[seg];
    if (!switches || n != S.n_sw * h->nb) return fail(BOSSGPU_EINVAL, "bucket buffer must hold %lld entries", (long long)(S.n_sw * h->nb));
    BOSS_CUDA(cudaMemcpyAsync(h->d_bucket_sw + (size_t)S.sw_off * h->nb, switches, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(h->h_bucket_sw + (size_t)S.sw_off * h->nb, switches, (size_t)n);
    h->sticky_on = false;                 // the caller may have switched buckets off: look at the device flag again
    return 0;
}

extern "C" int bossgpu_get_hist(bossgpu_handle* h, int64_t* counts, double* f_grid) {
    H_CHECK(h);
    std::vector<unsigned long long> tmp(3 * HBINS + 4);
    BOSS_CUDA(cudaMemcpyAsync(tmp.data(), h->d_hist, sizeof(unsigned long long) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    for (int e = 0; e < HBINS; ++e) {
        if (counts) counts[e] = (int64_t)tmp[e];
        if (f_grid) f_grid[e] = from_limbs(tmp[HBINS + e], tmp[2 * HBINS + e], h->fhat_shift);
    }
    return 0;
}

extern "C" int bossgpu_get_fhat(bossgpu_handle* h, int64_t row0, int64_t n_rows, double* out) {
    H_CHECK(h);
    if (!out || n_rows < 0 || row0 < 0) return fail(BOSSGPU_EINVAL, "bad row range");
    if (!h->have_fhat) return fail(BOSSGPU_ESTATE, "no F-hat on the device yet (no update has derived a strategy)");
    if (n_rows == 0) return 0;
    TRY(ensure_scratch(h, sizeof(double) * 2 * (size_t)n_rows));
    FhatGeom fg{h->n_windows_total, h->Tf_rows, h->target_rows};
    k_fhat_rows<<<(unsigned)ceil_div(2 * n_rows, 256), 256, 0, h->stream>>>(fg, h->d_fhat_w, h->d_upd, row0, n_rows, (double*)h->scratch_d);
    BOSS_KERNEL_CHECK();
    BOSS_CUDA(cudaMemcpyAsync(out, h->scratch_d, sizeof(double) * 2 * (size_t)n_rows, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bossgpu_get_score_table(bossgpu_handle* h, double* scores, double* entropies) {
    H_CHECK(h);
    if (scores) {
        BOSS_CUDA(cudaMemcpy(scores, h->d_table, sizeof(double) * (size_t)NPAT * 4, cudaMemcpyDeviceToHost));
        memcpy(scores, h->row0_true, sizeof(double) * 4);
    }
    if (entropies) {
        BOSS_CUDA(cudaMemcpy(entropies, h->d_etable, sizeof(double) * (size_t)NPAT * 4, cudaMemcpyDeviceToHost));
        memcpy(entropies, h->row0_true + 4, sizeof(double) * 4);
    }
    return 0;
}

extern "C" int64_t bossgpu_pattern_rank(const uint16_t c[5]) {
    if (!c) return -1;
    uint32_t s = (uint32_t)c[0] + c[1] + c[2] + c[3] + c[4];
    if (s >= (uint32_t)FREEZE) return -1;
    return pattern_rank(c[0], c[1], c[2], c[3], c[4]);
}

extern "C" int bossgpu_timing(bossgpu_handle* h, float ms[BOSSGPU_N_TIMERS]) {
    if (!h || !ms) return fail(BOSSGPU_EINVAL, "null argument");
    BOSS_CUDA(cudaSetDevice(h->device));
    if (h->ev_valid[0]) { BOSS_CUDA(cudaEventSynchronize(h->ev[1])); cudaEventElapsedTime(&h->ms[0], h->ev[0], h->ev[1]); }
    for (int i = 0; i < BOSSGPU_N_TIMERS; ++i) ms[i] = h->ev_valid[i] ? h->ms[i] : -1.0f;
    return 0;
}

extern "C" int64_t bossgpu_launch_count(bossgpu_handle* h) { return h ? h->launches : -1; }
extern "C" int64_t bossgpu_ingest_bytes(bossgpu_handle* h) { return h ? h->last_ingest_h2d : -1; }

extern "C" int bossgpu_synth_coverage(bossgpu_handle* h, uint64_t seed, double mean_depth, double p_ref, double p_del,
                                      double frac_dropout, double frac_deep) {
    H_CHECK(h);
    TRY(prescore_invalidate(h));            // counters change under an early pass: the update scores every tile again
    BOSS_CUDA(cudaMemsetAsync(h->d_cov_total, 0, sizeof(unsigned long long) * h->n_contigs_total, h->stream));
    for (const auto& S : h->segs) {
        unsigned grid = (unsigned)std::min<int64_t>(ceil_div(S.len, 256), 148 * 32);
        k_synth_coverage<<<grid, 256, 0, h->stream>>>(S, h->nb, h->P, h->d_ref, h->d_cov, h->d_cov_total, seed, mean_depth, p_ref, p_del,
                                                      frac_dropout, frac_deep);
        BOSS_KERNEL_CHECK();
    }
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}


// ------------------------------------------------------------------------------------------------
// host mirror of the masks, per-segment accept counts, device-side read-start counts
// ------------------------------------------------------------------------------------------------
extern "C" int bossgpu_strat_host(bossgpu_handle* h, uint8_t** ptr, int64_t* bytes) {
    if (!h || !ptr || !bytes) return fail(BOSSGPU_EINVAL, "null argument");
    *ptr = h->h_strat;
    *bytes = h->n_srows * 2 * h->nb;
    return 0;
}

extern "C" int bossgpu_buckets_host(bossgpu_handle* h, uint8_t** ptr, int64_t* bytes) {
    if (!h || !ptr || !bytes) return fail(BOSSGPU_EINVAL, "null argument");
    *ptr = h->h_bucket_sw;
    *bytes = h->n_sw * h->nb;
    return 0;
}

static void page_range(const void* p, size_t bytes, uintptr_t* lo, uintptr_t* hi) {
    const uintptr_t page = 4096;
    *lo = (uintptr_t)p & ~(page - 1);
    *hi = ((uintptr_t)p + std::max<size_t>(bytes, 1) + page - 1) & ~(page - 1);
}

extern "C" int bossgpu_host_register(void* host_ptr, int64_t bytes) {
    if (!host_ptr || bytes < 0) return fail(BOSSGPU_EINVAL, "bad host range");
    uintptr_t lo, hi;
    page_range(host_ptr, (size_t)bytes, &lo, &hi);       // shared-memory and malloc mappings cover whole pages
    cudaError_t e = cudaHostRegister((void*)lo, hi - lo, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        // a previous owner of these pages went away without unregistering them (the allocator reused the address)
        cudaGetLastError();
        cudaHostUnregister((void*)lo);
        cudaGetLastError();
        e = cudaHostRegister((void*)lo, hi - lo, cudaHostRegisterMapped | cudaHostRegisterPortable);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(BOSSGPU_ECUDA, "cudaHostRegister(%zu bytes) failed: %s", (size_t)(hi - lo), cudaGetErrorString(e));
    }
    return 0;
}

extern "C" int bossgpu_host_unregister(void* host_ptr) {
    if (!host_ptr) return 0;
    uintptr_t lo, hi;
    page_range(host_ptr, 1, &lo, &hi);
    cudaError_t e = cudaHostUnregister((void*)lo);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(BOSSGPU_ECUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
    return 0;
}

extern "C" int bossgpu_set_strat_mirror(bossgpu_handle* h, void* host_ptr, int64_t bytes, int registered) {
    H_CHECK(h);
    const int64_t total = h->n_srows * 2 * h->nb;
    if (!host_ptr || bytes != total) return fail(BOSSGPU_EINVAL, "the strategy mirror must hold exactly %lld bytes", (long long)total);
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    if (!registered) {
        TRY(bossgpu_host_register(host_ptr, bytes));
        if (h->reg_base) bossgpu_host_unregister(h->reg_base);
        h->reg_base = host_ptr;
    }
    void* dev = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dev, host_ptr, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(BOSSGPU_ECUDA, "cudaHostGetDevicePointer failed: %s (is the range registered?)", cudaGetErrorString(e)); }
    // move the device copy so that both images are congruent mod 16 (vector stores on both sides)
    const int shift = (int)((uintptr_t)host_ptr & 15);
    if (shift != h->strat_shift) {
        TRY(ensure_scratch(h, (size_t)total));
        BOSS_CUDA(cudaMemcpy(h->scratch_d, h->d_strat, (size_t)total, cudaMemcpyDeviceToDevice));
        h->d_strat = h->d_strat_alloc + shift;
        h->strat_shift = shift;
        BOSS_CUDA(cudaMemcpy(h->d_strat, h->scratch_d, (size_t)total, cudaMemcpyDeviceToDevice));
    }
    h->h_strat = (uint8_t*)host_ptr;
    h->h_strat_dev = (uint8_t*)dev;
    // the new mirror starts as an image of the current strategy
    BOSS_CUDA(cudaMemcpy(h->h_strat, h->d_strat, (size_t)total, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int bossgpu_get_seg_accept(bossgpu_handle* h, int64_t* out, int64_t n) {
    if (!h || !out || n != 2 * (int64_t)h->n_seg) return fail(BOSSGPU_EINVAL, "accept buffer must hold 2 * n_segments entries");
    for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)h->h_seg_accept[i];
    return 0;
}

extern "C" int bossgpu_read_starts_add(bossgpu_handle* h, int64_t n, const int64_t* window, const uint8_t* strand) {
    H_CHECK(h);
    if (n < 0) return fail(BOSSGPU_EINVAL, "negative count");
    if (n == 0) return 0;
    if (!window || !strand) return fail(BOSSGPU_EINVAL, "null array");
    size_t o_s = round_up(sizeof(int64_t) * n, 16);
    size_t total = o_s + round_up((size_t)n, 16);
    TRY(ensure_stage(h, total));
    memcpy(h->stage_h, window, sizeof(int64_t) * n);
    memcpy((char*)h->stage_h + o_s, strand, (size_t)n);
    BOSS_CUDA(cudaMemcpyAsync(h->stage_d, h->stage_h, total, cudaMemcpyHostToDevice, h->stream));
    k_count_read_starts<<<(unsigned)ceil_div(n, 256), 256, 0, h->stream>>>(n, (const int64_t*)h->stage_d,
                                                                          (const uint8_t*)((char*)h->stage_d + o_s),
                                                                          h->n_windows_total, h->d_rs_counts);
    BOSS_KERNEL_CHECK();
    h->launches++;
    BOSS_CUDA(cudaStreamSynchronize(h->stream));      // the staging buffer is reused by the next call
    return 0;
}

extern "C" int bossgpu_get_read_starts(bossgpu_handle* h, int64_t* out, int64_t n) {
    H_CHECK(h);
    if (!out || n != 2 * h->n_windows_total) return fail(BOSSGPU_EINVAL, "buffer must hold 2 * n_windows_total entries");
    BOSS_CUDA(cudaMemcpyAsync(out, h->d_rs_counts, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, h->stream));
    BOSS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}


// ------------------------------------------------------------------------------------------------
// BOSS-AEONS benefit / threshold step (aeons.cuh); stateless
// ------------------------------------------------------------------------------------------------
extern "C" int bossgpu_aeons_update(int device, int64_t n_seq, const int64_t* node_off, const double* scores, const uint8_t* e1,
                                    const uint8_t* e2, const bossgpu_aeons_params* p, double* benefit, double* smu_sum,
                                    uint8_t* strat, int64_t* counts, bossgpu_aeons_result* r) {
    if (!p || !node_off || n_seq <= 0 || !scores || !e1 || !e2 || !benefit || !smu_sum)
        return fail(BOSSGPU_EINVAL, "null argument");
    if (p->want_strategy && (!strat || !r)) return fail(BOSSGPU_EINVAL, "strategy outputs missing");
    if (p->mu_ds < 1) return fail(BOSSGPU_EINVAL, "anchor window of %d nodes; Bottleneck's move_sum rejects windows < 1", p->mu_ds);
    for (int i = 0; i < NSTEPS; ++i) {
        if (p->ccl_ds[i] < 1) return fail(BOSSGPU_EINVAL, "staircase window %d is %d nodes; Bottleneck's move_sum rejects windows < 1", i, p->ccl_ds[i]);
        if (i > 0 && p->ccl_ds[i] < p->ccl_ds[i - 1]) return fail(BOSSGPU_EINVAL, "staircase windows must be non-decreasing");
        if (p->ccl_ds[i] > 12000) return fail(BOSSGPU_EINVAL, "staircase window %d exceeds the supported 12000 nodes", i);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(BOSSGPU_ECUDA, "no usable CUDA device; libbossgpu has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(BOSSGPU_EINVAL, "device %d out of range [0,%d)", device, ndev);
    BOSS_CUDA(cudaSetDevice(device));
    const int64_t total = node_off[n_seq];
    if (node_off[0] != 0 || total <= 0) return fail(BOSSGPU_EINVAL, "node offsets must start at 0 and hold at least one node");
    std::vector<int32_t> tseq, tj0;
    for (int64_t i = 0; i < n_seq; ++i) {
        const int64_t n = node_off[i + 1] - node_off[i];
        if (n <= 0) return fail(BOSSGPU_EINVAL, "contig %lld has no nodes", (long long)i);
        for (int64_t j = 0; j < n; j += AE_TILE) { tseq.push_back((int32_t)i); tj0.push_back((int32_t)j); }
    }
    const size_t n_tiles = tseq.size();
    // one device blob: [off | scores | e1 | e2 | tile_seq | tile_j0 | benefit | smu_sum | strat | counts | norm, nnz | out]
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (size_t)round_up((int64_t)bytes, 256); return at; };
    const size_t o_off = take(sizeof(int64_t) * (n_seq + 1)), o_sc = take(sizeof(double) * total), o_e1 = take((size_t)n_seq),
                 o_e2 = take((size_t)n_seq), o_ts = take(sizeof(int32_t) * n_tiles), o_tj = take(sizeof(int32_t) * n_tiles);
    const size_t in_bytes = o;
    const size_t o_ben = take(sizeof(double) * 2 * total), o_ss = take(sizeof(double) * n_seq), o_st = take((size_t)2 * total),
                 o_cnt = take(sizeof(unsigned long long) * HBINS), o_misc = take(sizeof(unsigned long long) * 2), o_out = take(sizeof(AeonsOut));
    char* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, o);
    if (e != cudaSuccess) return fail(BOSSGPU_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", o, cudaGetErrorString(e));
    std::vector<char> hs(in_bytes, 0);
    memcpy(hs.data() + o_off, node_off, sizeof(int64_t) * (n_seq + 1));
    memcpy(hs.data() + o_sc, scores, sizeof(double) * total);
    memcpy(hs.data() + o_e1, e1, (size_t)n_seq);
    memcpy(hs.data() + o_e2, e2, (size_t)n_seq);
    memcpy(hs.data() + o_ts, tseq.data(), sizeof(int32_t) * n_tiles);
    memcpy(hs.data() + o_tj, tj0.data(), sizeof(int32_t) * n_tiles);
    int rc = 0;
    auto finish = [&](int code) { cudaFree(d); return code; };
#define AE_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { rc = fail(BOSSGPU_ECUDA, "%s failed: %s", #x, cudaGetErrorString(_e)); return finish(rc); } } while (0)
    AE_CUDA(cudaMemcpy(d, hs.data(), in_bytes, cudaMemcpyHostToDevice));
    AE_CUDA(cudaMemset(d + o_cnt, 0, o - o_cnt));
    AeonsArgs a;
    a.n_seq = n_seq; a.off = (const int64_t*)(d + o_off); a.scores = (const double*)(d + o_sc); a.e1 = (const uint8_t*)(d + o_e1);
    a.e2 = (const uint8_t*)(d + o_e2); a.tile_seq = (const int32_t*)(d + o_ts); a.tile_j0 = (const int32_t*)(d + o_tj);
    a.mu = p->mu_ds;
    for (int i = 0; i < NSTEPS; ++i) { a.w[i] = p->ccl_ds[i]; a.perc[i] = p->perc[i]; }
    a.benefit = (double*)(d + o_ben); a.norm_bits = (unsigned long long*)(d + o_misc);
    const int wmax = std::max(p->ccl_ds[NSTEPS - 1], p->mu_ds);
    const size_t smem = sizeof(double) * (AE_TILE + 2 * (size_t)wmax);
    if (smem > 200 * 1024) return finish(fail(BOSSGPU_EINVAL, "window of %d nodes needs %zu B of shared memory", wmax, smem));
    AE_CUDA(cudaFuncSetAttribute(k_aeons_benefit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_aeons_benefit<<<(unsigned)n_tiles, AE_TILE, smem>>>(a);
    AE_CUDA(cudaGetLastError());
    k_aeons_smu_sum<<<(unsigned)n_seq, 256>>>(n_seq, a.off, a.scores, a.e1, a.e2, p->mu_ds, (int64_t)p->ccl_ds[NSTEPS - 1], (double*)(d + o_ss));
    AE_CUDA(cudaGetLastError());
    if (p->want_strategy) {
        k_aeons_hist<<<(unsigned)std::min<int64_t>(ceil_div(2 * total, 256), 148 * 8), 256>>>(
            2 * total, a.benefit, a.norm_bits, (unsigned long long*)(d + o_cnt), (unsigned long long*)(d + o_misc) + 1);
        AE_CUDA(cudaGetLastError());
        k_aeons_threshold<<<1, 32>>>(n_seq, (const double*)(d + o_ss), (const unsigned long long*)(d + o_cnt), a.norm_bits,
                                     (const unsigned long long*)(d + o_misc) + 1, p->tc, p->tbar0, (AeonsOut*)(d + o_out));
        AE_CUDA(cudaGetLastError());
        k_aeons_mask<<<(unsigned)n_tiles, AE_TILE>>>(n_seq, a.off, a.benefit, (const AeonsOut*)(d + o_out), a.tile_seq, a.tile_j0,
                                                     (uint8_t*)(d + o_st));
        AE_CUDA(cudaGetLastError());
    }
    AE_CUDA(cudaMemcpy(benefit, d + o_ben, sizeof(double) * 2 * total, cudaMemcpyDeviceToHost));
    AE_CUDA(cudaMemcpy(smu_sum, d + o_ss, sizeof(double) * n_seq, cudaMemcpyDeviceToHost));
    if (p->want_strategy) {
        AeonsOut out;
        AE_CUDA(cudaMemcpy(&out, d + o_out, sizeof out, cudaMemcpyDeviceToHost));
        r->threshold = out.threshold; r->normaliser = out.normaliser; r->ubar0 = out.ubar0; r->n_nonzero = (int64_t)out.n_nonzero;
        r->strat_size = out.strat_size; r->reserved = 0;
        if (counts) AE_CUDA(cudaMemcpy(counts, d + o_cnt, sizeof(unsigned long long) * HBINS, cudaMemcpyDeviceToHost));
        if (out.empty) return finish(fail(BOSSGPU_EEMPTY, "all benefits are zero: upstream np.max of an empty array raises ValueError"));
        AE_CUDA(cudaMemcpy(strat, d + o_st, (size_t)2 * total, cudaMemcpyDeviceToHost));
    }
#undef AE_CUDA
    return finish(0);
}
