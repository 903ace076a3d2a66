// score_pass.cuh — the streaming pass over every site: score selection + dropout masking + 100-site
// binning + 20 kb bucket depth sums, fused.
//
// Replaces Scoring.update_scores (boss/runs/sequences.py:398-455), Contig.modify_scores/_find_dropout
// (boss/runs/reference.py:148-179), the depth sums of Contig.check_buckets (reference.py:196-199) and
// the binning half of Contig.calc_smu (reference.py:227-231).
//
// The reference keeps a float64 score per site and patches it where the batch touched the contig.
// Those stored values are a pure function of the current counters (and of the contig-wide dropout
// threshold), so this pass recomputes the score of every site from its five counters and the
// reference base and never materialises the per-site array:
//
//     row dropped (depth rule active and some barcode of the row has depth <= thr)  -> 0        reference.py:158-161
//     depth >= 30                                                                   -> tiny     sequences.py:419-420,430
//     row ever observed (any barcode, Q6)                                           -> table[rank(counts)][ref]   :428
//     otherwise                                                                     -> score0 of the contig (Q5)
//
// Algorithmic HBM traffic: 10 B of counters + 1 B reference base per site*barcode read, 8 B per bin
// written (0.08 B/site). The 8.9 MB table is L2/L1 resident.
#pragma once
#include "common.cuh"
#include "table.cuh"

namespace boss {

// one entry per CTA of the score/bin pass, built on the host at create time
struct TileDesc {
    int64_t site_off;     // first site of the tile on the padded site axis (multiple of 8)
    int64_t ds_index;     // index of the tile's first bin in scores_ds (barcode 0)
    int32_t n_sites;      // valid sites in the tile (0 for the tile that only holds a contig's empty last bin)
    int32_t n_bins;       // bins of this tile that exist (<= 20)
    int32_t bucket;       // entry of bucket_sum this tile adds to, or -1 (partial bucket at the contig end)
    int32_t contig;       // global contig index (dropout threshold)
};

struct ScoreArgs {
    const TileDesc* tiles;
    int nb;
    int64_t P;
    const uint8_t* ref;
    const uint16_t* cov;
    const uint32_t* rowflag;         // nb > 1: (touched << 31) | min over barcodes of depth
    const double* table;             // [NPAT + 3][4]; the three extra rows hold tiny, score0, 0.0
    const int32_t* drop_thr;         // per global contig; -1 = rule inactive
    double* ds;                      // [nb][ds_len]
    int64_t ds_len;
    uint32_t* tile_cov;              // [n_tiles][nb] depth total of the tile  (plain stores: a tile scored twice in
    uint32_t* tile_drop;             // [n_tiles]     rows zeroed by the depth rule    one update simply overwrites itself)
    // late half of a split pass (k_score_bin_tma MODE 2, see bossgpu_prescore): the tiles to score, in no particular order
    const int32_t* tile_list = nullptr;
    const unsigned* list_n = nullptr;
};

constexpr int ROW_TINY = NPAT;       // depth >= 30: frozen site            (sequences.py:419-420,430)
constexpr int ROW_SCORE0 = NPAT + 1; // row never observed: contig score0   (reference.py:103-104, Q5)
constexpr int ROW_ZERO = NPAT + 2;   // dropped row                         (reference.py:161)

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ starts, int n, int64_t x) {
    int lo = 0, hi = n;                  // starts[lo] <= x < starts[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (starts[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// nb > 1 only: per-row summary over all barcodes (Q6 / Q8 act on whole rows)
__global__ void __launch_bounds__(256)
k_rowflags(int64_t P4, int nb, int64_t P, const uint16_t* __restrict__ cov, uint32_t* __restrict__ rowflag) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // group of 4 sites
    if (g >= P4) return;
    uint32_t mn[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    uint32_t any[4] = {0, 0, 0, 0};
    for (int b = 0; b < nb; ++b) {
        uint32_t cs[4] = {0, 0, 0, 0};
        for (int base = 0; base < 5; ++base) {
            uint2 v = __ldg(reinterpret_cast<const uint2*>(cov + ((size_t)(b * 5 + base)) * P) + g);
            cs[0] += v.x & 0xFFFFu; cs[1] += v.x >> 16; cs[2] += v.y & 0xFFFFu; cs[3] += v.y >> 16;
        }
        for (int i = 0; i < 4; ++i) { mn[i] = min(mn[i], cs[i]); any[i] |= cs[i]; }
    }
    uint4 out;
    out.x = (any[0] ? 0x80000000u : 0u) | mn[0];
    out.y = (any[1] ? 0x80000000u : 0u) | mn[1];
    out.z = (any[2] ? 0x80000000u : 0u) | mn[2];
    out.w = (any[3] ? 0x80000000u : 0u) | mn[3];
    reinterpret_cast<uint4*>(rowflag)[g] = out;
}

constexpr int SB_THREADS = 256;      // 250 active threads x 8 sites = one 2000-site tile
constexpr int SB_SITES = 8;

// Table row of one site. Branch-free: the special cases select a ROW, not a value, so every site costs
// one 8-byte gather. s_T[k][p] = C(p + k + 1, k + 2) are the binomials of the pattern rank (table.cuh).
// Without barcodes "row never observed" is the all-zero pattern itself, so table row 0 holds the contig's
// score0 in that case (patched at create) and no extra select is needed.
template <bool MULTI>
__device__ __forceinline__ uint32_t site_row(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4, uint32_t flag,
                                             int32_t thr, const uint32_t (*s_T)[FREEZE], uint32_t& depth, uint32_t& dropped) {
    const uint32_t p1 = c0, p2 = p1 + c1, p3 = p2 + c2, p4 = p3 + c3, cs = p4 + c4;
    depth = cs;
    // clamped prefixes keep the shared-memory lookups in range when the site is frozen
    const uint32_t rank = min(p1, 29u) + s_T[0][min(p2, 29u)] + s_T[1][min(p3, 29u)] + s_T[2][min(p4, 29u)] +
                          s_T[3][min(cs, 29u)];
    uint32_t row = cs < (uint32_t)FREEZE ? rank : (uint32_t)ROW_TINY;
    if (MULTI) row = (flag >> 31) ? row : (uint32_t)ROW_SCORE0;
    const int32_t row_min = MULTI ? (int32_t)(flag & 0x7FFFFFFFu) : (int32_t)cs;
    const bool drop = row_min <= thr;                    // thr = -1 while the depth rule is inactive
    dropped = drop ? 1u : 0u;
    return drop ? (uint32_t)ROW_ZERO : row;
}

// grid = (tiles, nb); one CTA = 2000 consecutive sites of one segment for one barcode
template <bool MULTI>
__global__ void __launch_bounds__(SB_THREADS)
k_score_bin(ScoreArgs a) {
    __shared__ double s_part[TILE / 4];
    __shared__ uint32_t s_T[4][FREEZE];
    __shared__ unsigned s_cov, s_drop;

    const TileDesc td = a.tiles[blockIdx.x];
    const int b = blockIdx.y;
    const int t = threadIdx.x;
    if (t < 4 * FREEZE) {
        const int k = t / FREEZE, p = t - k * FREEZE;
        const int64_t q = p + k + 1;
        s_T[k][p] = (uint32_t)(k == 0 ? binom2(q) : k == 1 ? binom3(q) : k == 2 ? binom4(q) : binom5(q));
    }
    if (t == 0) { s_cov = 0; s_drop = 0; }
    __syncthreads();
    const int32_t thr_i = a.drop_thr[td.contig];        // -1 while the depth rule is inactive

    unsigned covsum = 0, ndrop = 0;
    if (t < TILE / SB_SITES) {
        const int l = SB_SITES * t;
        double part0 = 0.0, part1 = 0.0;
        if (l < td.n_sites) {
            const size_t g = ((size_t)td.site_off >> 3) + t;             // 8-site group on the padded axis
            const uint16_t* plane = a.cov + (size_t)b * 5 * a.P;
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(plane) + g);
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(plane + a.P) + g);
            const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(plane + 2 * a.P) + g);
            const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(plane + 3 * a.P) + g);
            const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(plane + 4 * a.P) + g);
            const uint2 rb = __ldg(reinterpret_cast<const uint2*>(a.ref) + g);
            uint4 f0 = make_uint4(0, 0, 0, 0), f1 = make_uint4(0, 0, 0, 0);
            if (MULTI) {
                f0 = __ldg(reinterpret_cast<const uint4*>(a.rowflag) + 2 * g);
                f1 = __ldg(reinterpret_cast<const uint4*>(a.rowflag) + 2 * g + 1);
            }
            const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w},
                           w2[4] = {v2.x, v2.y, v2.z, v2.w}, w3[4] = {v3.x, v3.y, v3.z, v3.w},
                           w4[4] = {v4.x, v4.y, v4.z, v4.w};
            const uint32_t fl[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            const int nvalid = td.n_sites - l;                           // >= 8 except in a contig's last tile
            double s[SB_SITES];
            uint32_t c[SB_SITES][5], cs[SB_SITES], mx = 0;
#pragma unroll
            for (int i = 0; i < SB_SITES; ++i) {
                const int wi = i >> 1;
                if (i & 1) { c[i][0] = w0[wi] >> 16; c[i][1] = w1[wi] >> 16; c[i][2] = w2[wi] >> 16; c[i][3] = w3[wi] >> 16; c[i][4] = w4[wi] >> 16; }
                else { c[i][0] = w0[wi] & 0xFFFFu; c[i][1] = w1[wi] & 0xFFFFu; c[i][2] = w2[wi] & 0xFFFFu; c[i][3] = w3[wi] & 0xFFFFu; c[i][4] = w4[wi] & 0xFFFFu; }
                cs[i] = c[i][0] + c[i][1] + c[i][2] + c[i][3] + c[i][4];
                mx = max(mx, cs[i]);
                covsum += cs[i];                                         // padding counters are zero
            }
            if (!MULTI && mx < (uint32_t)FREEZE && nvalid >= SB_SITES) {
                // common case: no frozen site, no padding -> no clamps, no special rows except "dropped"
#pragma unroll
                for (int i = 0; i < SB_SITES; ++i) {
                    const uint32_t p2 = c[i][0] + c[i][1], p3 = p2 + c[i][2], p4 = p3 + c[i][3];
                    const uint32_t rank = c[i][0] + s_T[0][p2] + s_T[1][p3] + s_T[2][p4] + s_T[3][cs[i]];
                    const bool drop = (int32_t)cs[i] <= thr_i;
                    const uint32_t row = drop ? (uint32_t)ROW_ZERO : rank;
                    const uint32_t refb = ((i < 4 ? rb.x : rb.y) >> (8 * (i & 3))) & 0xFFu;
                    s[i] = __ldg(a.table + (size_t)(row * 4 + refb));
                    ndrop += drop ? 1u : 0u;
                }
            } else {
#pragma unroll
                for (int i = 0; i < SB_SITES; ++i) {
                    uint32_t depth, dropped;
                    uint32_t row = site_row<MULTI>(c[i][0], c[i][1], c[i][2], c[i][3], c[i][4], fl[i], thr_i, s_T, depth, dropped);
                    if (i >= nvalid) { row = ROW_ZERO; dropped = 0; }    // padding behind the contig end
                    const uint32_t refb = ((i < 4 ? rb.x : rb.y) >> (8 * (i & 3))) & 0xFFu;
                    s[i] = __ldg(a.table + (size_t)(row * 4 + refb));
                    ndrop += dropped;
                }
            }
            part0 = ((s[0] + s[1]) + s[2]) + s[3];                       // position order within a thread
            part1 = ((s[4] + s[5]) + s[6]) + s[7];
        }
        s_part[2 * t] = part0;
        s_part[2 * t + 1] = part1;
    }
    // block sums of depth (bucket means) and dropped rows (log line)
    covsum = __reduce_add_sync(0xFFFFFFFFu, covsum);
    ndrop = __reduce_add_sync(0xFFFFFFFFu, ndrop);
    if ((t & 31) == 0) {
        if (covsum) atomicAdd(&s_cov, covsum);
        if (ndrop) atomicAdd(&s_drop, ndrop);
    }
    __syncthreads();

    if (t < TILE / BIN) {
        // bin j of the tile = 25 consecutive 4-site partials, added in position order
        if (t < td.n_bins) {
            double acc = 0.0;
#pragma unroll 5
            for (int k = 0; k < 25; ++k) acc += s_part[25 * t + k];
            a.ds[(size_t)b * a.ds_len + td.ds_index + t] = acc;
        }
    } else if (t == 32) {
        a.tile_cov[(size_t)blockIdx.x * a.nb + b] = s_cov;
        if (b == 0) a.tile_drop[blockIdx.x] = s_drop;
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged persistent variant of the pass (the one the library launches).
//
// One CTA = 8 consumer warps + 1 producer warp, grid = (min(tiles, 2 CTAs per SM), barcodes). The
// producer's elected lane walks the CTA's tiles (tile = blockIdx.x + it * gridDim.x) and, four tiles ahead
// of the consumers, arms the stage's "full" mbarrier with the byte count and issues the bulk copies
// (cp.async.bulk global -> shared: 5 counter planes x 4000 B + 2000 B of reference bases, + 8000 B of row
// flags with barcodes). Consumers wait on "full", score their 8 sites out of shared memory, release the
// stage through the "empty" mbarrier and meet at a named barrier for the 100-site bin sums. HBM latency is
// covered by ~88 KB of copies in flight per CTA instead of by occupancy.
// ------------------------------------------------------------------------------------------------
constexpr int SBT_CONSUMERS = 256;
constexpr int SBT_THREADS = SBT_CONSUMERS + 32;
constexpr int SBT_PLANE_BYTES = TILE * 2;                  // 4000
constexpr int SBT_REF_OFF = 5 * SBT_PLANE_BYTES;           // 20000
constexpr int SBT_FLAG_OFF = SBT_REF_OFF + 2048;           // 22048 (reference bases padded to 2048)
constexpr int SBT_STAGE_BYTES = 22144;                     // multiple of 128, without row flags
constexpr int SBT_STAGE_BYTES_MULTI = SBT_FLAG_OFF + TILE * 4 + 96;   // 30144
constexpr int64_t SBT_TAIL_PAD = 2048;                     // elements allocated behind the site arrays: the last
                                                           // tile of a segment is copied whole

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SBT_CONSUMERS) : "memory"); }

// MODE 0: every tile. MODE 2: the tiles of a.tile_list (late half of a split pass, bossgpu_prescore*).
template <bool MULTI, int SBT_STAGES, int MODE = 0>
__global__ void __launch_bounds__(SBT_THREADS, 3)
k_score_bin_tma(ScoreArgs a, int64_t n_tiles) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int STAGE = MULTI ? SBT_STAGE_BYTES_MULTI : SBT_STAGE_BYTES;
    unsigned char* s_stage = s_raw;
    TileDesc* s_td = reinterpret_cast<TileDesc*>(s_raw + SBT_STAGES * STAGE);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_td + SBT_STAGES);     // full[4], empty[4]
    double* s_part = reinterpret_cast<double*>(s_bar + 2 * SBT_STAGES);                       // [2][TILE/4]
    uint32_t (*s_T)[FREEZE] = reinterpret_cast<uint32_t (*)[FREEZE]>(s_part + 2 * (TILE / 4));
    unsigned* s_covw = reinterpret_cast<unsigned*>(s_T + 4);                                   // [2][8]
    unsigned* s_dropw = s_covw + 16;                                                           // [2][8]

    const int t = threadIdx.x;
    const int b = blockIdx.y;
    if (t < 4 * FREEZE) {
        const int k = t / FREEZE, p = t - k * FREEZE;
        const int64_t q = p + k + 1;
        s_T[k][p] = (uint32_t)(k == 0 ? binom2(q) : k == 1 ? binom3(q) : k == 2 ? binom4(q) : binom5(q));
    }
    if (t == 0) {
        for (int s = 0; s < SBT_STAGES; ++s) {
            mbar_init(smem_u32(&s_bar[s]), 1);                              // full: the producer's expect_tx arrival
            mbar_init(smem_u32(&s_bar[SBT_STAGES + s]), SBT_CONSUMERS / 32); // empty: one arrival per consumer warp
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // make the inits visible to the copy engine
    }
    __syncthreads();

    const int64_t first = blockIdx.x, stride = gridDim.x;

    if (t >= SBT_CONSUMERS) {
        // ------------------------------- producer warp -------------------------------
        if (t == SBT_CONSUMERS) {
            const uint16_t* plane0 = a.cov + (size_t)b * 5 * a.P;
            int it = 0;
            const int64_t n_iter = MODE == 2 ? (int64_t)*a.list_n : n_tiles;
            for (int64_t idx = first; idx < n_iter; idx += stride) {
                const int64_t tile = MODE == 2 ? (int64_t)a.tile_list[idx] : idx;
                const int stage = it % SBT_STAGES;
                const uint32_t full = smem_u32(&s_bar[stage]), empty = smem_u32(&s_bar[SBT_STAGES + stage]);
                if (it >= SBT_STAGES) mbar_wait(empty, ((it / SBT_STAGES) - 1) & 1);
                const TileDesc td = a.tiles[tile];
                s_td[stage] = td;
                const uint32_t dst = smem_u32(s_stage + (size_t)stage * STAGE);
                mbar_expect_tx(full, MULTI ? (5 * SBT_PLANE_BYTES + TILE + TILE * 4) : (5 * SBT_PLANE_BYTES + TILE));
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    bulk_g2s(dst + k * SBT_PLANE_BYTES, plane0 + (size_t)k * a.P + td.site_off, SBT_PLANE_BYTES, full);
                bulk_g2s(dst + SBT_REF_OFF, a.ref + td.site_off, TILE, full);
                if (MULTI) bulk_g2s(dst + SBT_FLAG_OFF, a.rowflag + td.site_off, TILE * 4, full);
                ++it;
            }
        }
        return;
    }

    // ----------------------------------- consumers -----------------------------------
    int it = 0, buf = 0;
    const int64_t n_iter = MODE == 2 ? (int64_t)*a.list_n : n_tiles;
    for (int64_t idx = first; idx < n_iter; idx += stride) {
        const int stage = it % SBT_STAGES;
        mbar_wait(smem_u32(&s_bar[stage]), (it / SBT_STAGES) & 1);
        const TileDesc td = s_td[stage];
        const unsigned char* sb = s_stage + (size_t)stage * STAGE;
        const int32_t thr_i = a.drop_thr[td.contig];        // -1 while the depth rule is inactive
        double* part = s_part + buf * (TILE / 4);
        unsigned covsum = 0, ndrop = 0;
        // pull this thread's 8 sites out of the stage, then hand the stage back to the producer at once:
        // the rest of the iteration works from registers
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0, v4 = v0, f0 = v0, f1 = v0;
        uint2 rb = make_uint2(0, 0);
        if (t < TILE / SB_SITES) {
            v0 = reinterpret_cast<const uint4*>(sb)[t];
            v1 = reinterpret_cast<const uint4*>(sb + SBT_PLANE_BYTES)[t];
            v2 = reinterpret_cast<const uint4*>(sb + 2 * SBT_PLANE_BYTES)[t];
            v3 = reinterpret_cast<const uint4*>(sb + 3 * SBT_PLANE_BYTES)[t];
            v4 = reinterpret_cast<const uint4*>(sb + 4 * SBT_PLANE_BYTES)[t];
            rb = reinterpret_cast<const uint2*>(sb + SBT_REF_OFF)[t];
            if (MULTI) {
                f0 = reinterpret_cast<const uint4*>(sb + SBT_FLAG_OFF)[2 * t];
                f1 = reinterpret_cast<const uint4*>(sb + SBT_FLAG_OFF)[2 * t + 1];
            }
        }
        {
            // the arrival must not overtake the loads: fold every loaded word into a value the arriving lane needs
            const unsigned seen = v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w ^ v2.x ^ v2.y ^ v2.z ^ v2.w ^ v3.x ^ v3.y ^
                                  v3.z ^ v3.w ^ v4.x ^ v4.y ^ v4.z ^ v4.w ^ rb.x ^ rb.y ^ f0.x ^ f0.y ^ f0.z ^ f0.w ^ f1.x ^ f1.y ^
                                  f1.z ^ f1.w;
            const unsigned all_seen = __reduce_or_sync(0xFFFFFFFFu, seen);
            if ((t & 31) == 0) {
                asm volatile("" ::"r"(all_seen) : "memory");
                mbar_arrive(smem_u32(&s_bar[SBT_STAGES + stage]));       // this warp is done with the stage
            }
        }
        if (t < TILE / SB_SITES) {
            const int l = SB_SITES * t;
            double part0 = 0.0, part1 = 0.0;
            if (l < td.n_sites) {
                const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w}, w1[4] = {v1.x, v1.y, v1.z, v1.w},
                               w2[4] = {v2.x, v2.y, v2.z, v2.w}, w3[4] = {v3.x, v3.y, v3.z, v3.w},
                               w4[4] = {v4.x, v4.y, v4.z, v4.w};
                const uint32_t fl[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                const int nvalid = td.n_sites - l;                       // >= 8 except in a contig's last tile
                double s[SB_SITES];
                uint32_t c[SB_SITES][5], cs[SB_SITES], mx = 0;
#pragma unroll
                for (int i = 0; i < SB_SITES; ++i) {
                    const int wi = i >> 1;
                    if (i & 1) { c[i][0] = w0[wi] >> 16; c[i][1] = w1[wi] >> 16; c[i][2] = w2[wi] >> 16; c[i][3] = w3[wi] >> 16; c[i][4] = w4[wi] >> 16; }
                    else { c[i][0] = w0[wi] & 0xFFFFu; c[i][1] = w1[wi] & 0xFFFFu; c[i][2] = w2[wi] & 0xFFFFu; c[i][3] = w3[wi] & 0xFFFFu; c[i][4] = w4[wi] & 0xFFFFu; }
                    cs[i] = c[i][0] + c[i][1] + c[i][2] + c[i][3] + c[i][4];
                    mx = max(mx, cs[i]);
                }
                if (!MULTI && mx < (uint32_t)FREEZE && nvalid >= SB_SITES) {
                    // common case: no frozen site, no padding -> no clamps, no special rows except "dropped".
                    // The 30-entry binomial tables sit in 30 different banks: lookups never conflict.
#pragma unroll
                    for (int i = 0; i < SB_SITES; ++i) {
                        const uint32_t p2 = c[i][0] + c[i][1], p3 = p2 + c[i][2], p4 = p3 + c[i][3];
                        // C(p2+1,2) and C(p3+2,3) in registers (3 IMADs + an exact division by 3 via the modular
                        // inverse) instead of two more shared-memory lookups: the L1/shared pipe is this kernel's
                        // busiest unit (ncu: 85 %). Measured on B200 at 3.1 Gb: 6.65 -> 6.49 ms; moving C(p4+3,4) and
                        // C(cs+4,5) over as well costs more issue slots than it saves (6.85, 7.59 ms), and gathering
                        // the table past L1 (ld.global.cg) is 4.6x slower — the hot patterns live in L1
                        const uint32_t b2 = (p2 * (p2 + 1u)) >> 1;
                        const uint32_t b3 = ((p3 * (p3 + 1u) * (p3 + 2u)) >> 1) * 0xAAAAAAABu;
                        const uint32_t rank = c[i][0] + b2 + b3 + s_T[2][p4] + s_T[3][cs[i]];
                        const bool drop = (int32_t)cs[i] <= thr_i;
                        const uint32_t row = drop ? (uint32_t)ROW_ZERO : rank;
                        const uint32_t refb = ((i < 4 ? rb.x : rb.y) >> (8 * (i & 3))) & 0xFFu;
                        s[i] = __ldg(a.table + (size_t)(row * 4 + refb));
                        ndrop += drop ? 1u : 0u;
                        covsum += cs[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < SB_SITES; ++i) {
                        uint32_t depth, dropped;
                        uint32_t row = site_row<MULTI>(c[i][0], c[i][1], c[i][2], c[i][3], c[i][4], fl[i], thr_i, s_T, depth, dropped);
                        if (i >= nvalid) { row = ROW_ZERO; dropped = 0; depth = 0; }   // whatever follows the contig end
                        const uint32_t refb = ((i < 4 ? rb.x : rb.y) >> (8 * (i & 3))) & 0xFFu;
                        s[i] = __ldg(a.table + (size_t)(row * 4 + (refb & 3u)));
                        ndrop += dropped;
                        covsum += depth;
                    }
                }
                part0 = ((s[0] + s[1]) + s[2]) + s[3];                   // position order within a thread
                part1 = ((s[4] + s[5]) + s[6]) + s[7];
            }
            part[2 * t] = part0;
            part[2 * t + 1] = part1;
        }
        covsum = __reduce_add_sync(0xFFFFFFFFu, covsum);
        ndrop = __reduce_add_sync(0xFFFFFFFFu, ndrop);
        if ((t & 31) == 0) {
            s_covw[buf * 8 + (t >> 5)] = covsum;
            s_dropw[buf * 8 + (t >> 5)] = ndrop;
        }
        consumer_bar();
        if (t < TILE / BIN) {
            // bin j of the tile = 25 consecutive 4-site partials, added in position order
            if (t < td.n_bins) {
                double acc = 0.0;
#pragma unroll 5
                for (int k = 0; k < 25; ++k) acc += part[25 * t + k];
                a.ds[(size_t)b * a.ds_len + td.ds_index + t] = acc;
            }
        } else if (t == 32) {
            unsigned cov = 0, dr = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { cov += s_covw[buf * 8 + w]; dr += s_dropw[buf * 8 + w]; }
            const int64_t tile = MODE == 2 ? (int64_t)a.tile_list[idx] : idx;
            a.tile_cov[(size_t)tile * a.nb + b] = cov;
            if (b == 0) a.tile_drop[tile] = dr;
        }
        ++it;
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Barcoded runs (nb > 1): the whole-row rules couple the barcodes of a site — a row observed in ANY barcode is scored
// from the table in all of them (Q6), a row whose depth is at or below the dropout threshold in SOME barcode is zeroed in
// all of them (Q8) — so a site's scores need its counters in every barcode. One CTA takes a tile through ALL barcodes,
// half a tile (1000 sites, 4 per thread) at a time, in two sweeps:
//   sweep 1  every barcode's five counter planes come through the TMA ring once; each thread turns its sites' counters into
//            table rows (19 bits, parked in shared memory: 3 bytes per site and barcode) and keeps the row-wide minimum
//            depth and "ever observed" flag in registers;
//   sweep 2  per barcode: rows back from shared memory (the thread's own), row rules applied, one table gather per site,
//            100-site bin sums in position order.
// Every counter is read from HBM exactly once (10 B per site and barcode + 1 B reference base per site); the separate
// row-summary pass (k_rowflags: all counters a second time) is only kept for more barcodes than fit in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int SBM_CONSUMERS = 256;                 // 250 active: 4 sites each = 1000 sites
constexpr int SBM_THREADS = SBM_CONSUMERS + 32;
constexpr int SBM_HALF = TILE / 2;
constexpr int SBM_SITES = 4;
constexpr int SBM_PLANE_BYTES = SBM_HALF * 2;      // 2000
constexpr int SBM_REF_OFF = 5 * SBM_PLANE_BYTES;   // 10000
constexpr int SBM_STAGE_BYTES = 11136;             // 5 planes + a 1024-byte window of reference bases, padded to a multiple of 128
// bulk copies move multiples of 16 bytes from 16-byte aligned addresses: the first half's 1000 reference bases come as bytes
// [0, 1024) of the tile, the second half's as bytes [992, 2000) (so they start 8 bytes into the window)
constexpr int SBM_REF_BYTES_H0 = 1024, SBM_REF_SRC_H1 = 992, SBM_REF_BYTES_H1 = TILE - SBM_REF_SRC_H1, SBM_REF_SKIP_H1 = SBM_HALF - SBM_REF_SRC_H1;
constexpr int SBM_STAGES = 3;
constexpr int SBM_MAX_NB = 24;

constexpr size_t sbm_smem_bytes(int nb) {
    return (size_t)SBM_STAGES * SBM_STAGE_BYTES + (size_t)nb * SBM_HALF * 3 + 2 * (SBM_HALF / SBM_SITES) * sizeof(double) +
           4 * FREEZE * sizeof(uint32_t) + 2 * SBM_STAGES * sizeof(unsigned long long) + (size_t)nb * sizeof(unsigned) + 64 + 128;
}

template <int MODE>
__global__ void __launch_bounds__(SBM_THREADS, 2)
k_score_bin_multi(ScoreArgs a, int64_t n_tiles) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int nb = a.nb;
    unsigned char* s_stage = s_raw;
    uint16_t* s_rlo = reinterpret_cast<uint16_t*>(s_raw + SBM_STAGES * SBM_STAGE_BYTES);             // [nb][SBM_HALF]
    uint8_t* s_rhi = reinterpret_cast<uint8_t*>(s_rlo + (size_t)nb * SBM_HALF);                       // [nb][SBM_HALF]
    double* s_part = reinterpret_cast<double*>(s_raw + ((SBM_STAGES * SBM_STAGE_BYTES + (size_t)nb * SBM_HALF * 3 + 15) & ~(size_t)15));
    uint32_t (*s_T)[FREEZE] = reinterpret_cast<uint32_t (*)[FREEZE]>(s_part + 2 * (SBM_HALF / SBM_SITES));
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_T + 4);                        // full[], empty[]
    unsigned* s_cov = reinterpret_cast<unsigned*>(s_bar + 2 * SBM_STAGES);                             // [nb]
    unsigned* s_drop = s_cov + nb;

    const int t = threadIdx.x;
    if (t < 4 * FREEZE) {
        const int k = t / FREEZE, p = t - k * FREEZE;
        const int64_t q = p + k + 1;
        s_T[k][p] = (uint32_t)(k == 0 ? binom2(q) : k == 1 ? binom3(q) : k == 2 ? binom4(q) : binom5(q));
    }
    if (t == 0) {
        for (int s = 0; s < SBM_STAGES; ++s) {
            mbar_init(smem_u32(&s_bar[s]), 1);
            mbar_init(smem_u32(&s_bar[SBM_STAGES + s]), SBM_CONSUMERS / 32);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_iter = MODE == 2 ? (int64_t)*a.list_n : n_tiles;

    if (t >= SBM_CONSUMERS) {
        // ------------------------------- producer warp -------------------------------
        if (t == SBM_CONSUMERS) {
            int it = 0;
            for (int64_t idx = blockIdx.x; idx < n_iter; idx += gridDim.x) {
                const int64_t tile = MODE == 2 ? (int64_t)a.tile_list[idx] : idx;
                const int64_t site_off = a.tiles[tile].site_off;
                for (int half = 0; half < 2; ++half) {
                    for (int b = 0; b < nb; ++b, ++it) {
                        const int stage = it % SBM_STAGES;
                        const uint32_t full = smem_u32(&s_bar[stage]), empty = smem_u32(&s_bar[SBM_STAGES + stage]);
                        if (it >= SBM_STAGES) mbar_wait(empty, ((it / SBM_STAGES) - 1) & 1);
                        const uint32_t dst = smem_u32(s_stage + (size_t)stage * SBM_STAGE_BYTES);
                        mbar_expect_tx(full, 5 * SBM_PLANE_BYTES + (b == 0 ? (half ? SBM_REF_BYTES_H1 : SBM_REF_BYTES_H0) : 0));
                        const uint16_t* plane0 = a.cov + (size_t)b * 5 * a.P + site_off + half * SBM_HALF;
#pragma unroll
                        for (int k = 0; k < 5; ++k) bulk_g2s(dst + k * SBM_PLANE_BYTES, plane0 + (size_t)k * a.P, SBM_PLANE_BYTES, full);
                        if (b == 0) bulk_g2s(dst + SBM_REF_OFF, a.ref + site_off + (half ? SBM_REF_SRC_H1 : 0),
                                             half ? SBM_REF_BYTES_H1 : SBM_REF_BYTES_H0, full);
                    }
                }
            }
        }
        return;
    }

    // ----------------------------------- consumers -----------------------------------
    int it = 0, buf = 0;
    const bool active = t < SBM_HALF / SBM_SITES;                      // 250 of the 256 consumer threads hold sites
    if (t < nb) s_cov[t] = 0u;
    if (t == 0) *s_drop = 0u;
    consumer_bar();
    for (int64_t idx = blockIdx.x; idx < n_iter; idx += gridDim.x) {
        const int64_t tile = MODE == 2 ? (int64_t)a.tile_list[idx] : idx;
        const TileDesc td = a.tiles[tile];
        const int32_t thr_i = a.drop_thr[td.contig];                   // -1 while the depth rule is inactive
        for (int half = 0; half < 2; ++half) {
            const int l0 = half * SBM_HALF + SBM_SITES * t;            // first site of this thread within the tile
            uint32_t refw = 0, mn[SBM_SITES], any = 0;
#pragma unroll
            for (int i = 0; i < SBM_SITES; ++i) mn[i] = 0xFFFFFFFFu;
            // ---- sweep 1: counters -> table rows, row-wide minimum depth / observed flag ----
            for (int b = 0; b < nb; ++b, ++it) {
                const int stage = it % SBM_STAGES;
                mbar_wait(smem_u32(&s_bar[stage]), (it / SBM_STAGES) & 1);
                const unsigned char* sb = s_stage + (size_t)stage * SBM_STAGE_BYTES;
                uint2 v[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) v[k] = make_uint2(0, 0);
                if (active) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) v[k] = reinterpret_cast<const uint2*>(sb + k * SBM_PLANE_BYTES)[t];
                    if (b == 0) refw = reinterpret_cast<const uint32_t*>(sb + SBM_REF_OFF + (half ? SBM_REF_SKIP_H1 : 0))[t];
                }
                {
                    // the arrival must not overtake the loads: fold every loaded word into a value the arriving lane needs
                    const unsigned seen = v[0].x ^ v[0].y ^ v[1].x ^ v[1].y ^ v[2].x ^ v[2].y ^ v[3].x ^ v[3].y ^ v[4].x ^ v[4].y ^ refw;
                    const unsigned all_seen = __reduce_or_sync(0xFFFFFFFFu, seen);
                    if ((t & 31) == 0) {
                        asm volatile("" ::"r"(all_seen) : "memory");
                        mbar_arrive(smem_u32(&s_bar[SBM_STAGES + stage]));
                    }
                }
                unsigned covsum = 0;
                if (active) {
#pragma unroll
                    for (int i = 0; i < SBM_SITES; ++i) {
                        uint32_t c[5];
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            const uint32_t w = (i < 2) ? v[k].x : v[k].y;
                            c[k] = (i & 1) ? (w >> 16) : (w & 0xFFFFu);
                        }
                        const uint32_t p2 = c[0] + c[1], p3 = p2 + c[2], p4 = p3 + c[3], cs = p4 + c[4];
                        const uint32_t rank = min(c[0], 29u) + s_T[0][min(p2, 29u)] + s_T[1][min(p3, 29u)] + s_T[2][min(p4, 29u)] +
                                              s_T[3][min(cs, 29u)];
                        const uint32_t row = cs < (uint32_t)FREEZE ? rank : (uint32_t)ROW_TINY;
                        const int site = SBM_SITES * t + i;
                        s_rlo[(size_t)b * SBM_HALF + site] = (uint16_t)(row & 0xFFFFu);
                        s_rhi[(size_t)b * SBM_HALF + site] = (uint8_t)(row >> 16);
                        const bool valid = l0 + i < td.n_sites;
                        mn[i] = min(mn[i], cs);
                        any |= (cs != 0u ? 1u : 0u) << i;
                        covsum += valid ? cs : 0u;
                    }
                }
                covsum = __reduce_add_sync(0xFFFFFFFFu, covsum);
                if ((t & 31) == 0 && covsum) atomicAdd(&s_cov[b], covsum);
            }
            // ---- sweep 2: row rules, table gather, 100-site bins, barcode by barcode ----
            bool drop[SBM_SITES];
            unsigned ndrop = 0;
#pragma unroll
            for (int i = 0; i < SBM_SITES; ++i) {
                const bool valid = active && l0 + i < td.n_sites;
                drop[i] = valid && (int32_t)mn[i] <= thr_i;
                ndrop += drop[i] ? 1u : 0u;
            }
            ndrop = __reduce_add_sync(0xFFFFFFFFu, ndrop);
            if ((t & 31) == 0 && ndrop) atomicAdd(s_drop, ndrop);
            for (int b = 0; b < nb; ++b) {
                double* part = s_part + buf * (SBM_HALF / SBM_SITES);
                if (active) {
                    double sv[SBM_SITES];
#pragma unroll
                    for (int i = 0; i < SBM_SITES; ++i) {
                        const int site = SBM_SITES * t + i;
                        uint32_t row = (uint32_t)s_rlo[(size_t)b * SBM_HALF + site] | ((uint32_t)s_rhi[(size_t)b * SBM_HALF + site] << 16);
                        row = ((any >> i) & 1u) ? row : (uint32_t)ROW_SCORE0;      // Q5/Q6: row never observed in any barcode
                        if (drop[i] || l0 + i >= td.n_sites) row = ROW_ZERO;        // Q8 / whatever follows the contig end
                        const uint32_t refb = (refw >> (8 * i)) & 3u;
                        sv[i] = __ldg(a.table + (size_t)(row * 4 + refb));
                    }
                    part[t] = ((sv[0] + sv[1]) + sv[2]) + sv[3];                    // position order within a thread
                }
                consumer_bar();
                if (t < SBM_HALF / BIN) {
                    // bin j of the half = 25 consecutive 4-site partials, added in position order
                    const int bin = half * (SBM_HALF / BIN) + t;
                    if (bin < td.n_bins) {
                        double acc = 0.0;
#pragma unroll 5
                        for (int k = 0; k < 25; ++k) acc += part[25 * t + k];
                        a.ds[(size_t)b * a.ds_len + td.ds_index + bin] = acc;
                    }
                }
                buf ^= 1;
            }
        }
        consumer_bar();                                              // every warp's depth sums and drop counts are in
        if (t < nb) { a.tile_cov[(size_t)tile * nb + t] = s_cov[t]; s_cov[t] = 0u; }
        if (t == 32) { a.tile_drop[tile] = *s_drop; *s_drop = 0u; }
        consumer_bar();                                              // ... and reset, before the next tile adds to them
    }
}

// per-tile depth totals -> bucket sums (reference.py:196-198 sums whole 20 kb buckets; a tile is a tenth of one), and
// the number of dropped rows for the log line. Integer sums: order-free.
__global__ void k_tile_reduce(int64_t n_tiles, int nb, const TileDesc* __restrict__ tiles, const uint32_t* __restrict__ tile_cov,
                              const uint32_t* __restrict__ tile_drop, unsigned long long* __restrict__ bucket_sum,
                              unsigned long long* __restrict__ n_dropout) {
    const int64_t tile = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long dr = 0;
    if (tile < n_tiles) {
        const int32_t bucket = tiles[tile].bucket;
        if (bucket >= 0)
            for (int b = 0; b < nb; ++b) {
                const uint32_t c = tile_cov[(size_t)tile * nb + b];
                if (c) atomicAdd(&bucket_sum[(size_t)bucket * nb + b], (unsigned long long)c);
            }
        dr = tile_drop[tile];
    }
    for (int o = 16; o > 0; o >>= 1) dr += __shfl_down_sync(0xFFFFFFFFu, dr, o);
    if ((threadIdx.x & 31) == 0 && dr) atomicAdd(n_dropout, dr);
}

// Tiles the coming batch will write to: one thread per read marks the tiles its reference interval overlaps in its
// contig's segment of this shard (the same clipping the scatter applies) and appends newly marked tiles to a list.
__global__ void k_mark_tiles(int64_t n_reads, const int32_t* __restrict__ contig, const int64_t* __restrict__ t0s,
                             const int64_t* __restrict__ t1s, const int32_t* __restrict__ seg_of_contig,
                             const SegDev* __restrict__ segs, uint32_t* __restrict__ touched, int32_t* __restrict__ list,
                             unsigned* __restrict__ list_n) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int32_t sg = seg_of_contig[contig[r]];
    if (sg < 0) return;
    const SegDev S = segs[sg];
    const int64_t lo = max(t0s[r], S.start) - S.start, hi = min(t1s[r], S.start + S.len) - S.start;   // [lo, hi)
    if (hi <= lo) return;
    for (int64_t tile = S.tile_off + lo / TILE; tile <= S.tile_off + (hi - 1) / TILE; ++tile) {
        const uint32_t bit = 1u << (tile & 31);
        const uint32_t old = atomicOr(&touched[tile >> 5], bit);
        if (!(old & bit)) list[atomicAdd(list_n, 1u)] = (int32_t)tile;
    }
}

// Contigs whose dropout threshold differs between the early pass and the update (the batch moved the mean depth across
// a multiple of 8): every tile of their segment goes onto the list. One CTA per local segment.
__global__ void k_mark_changed(const SegDev* __restrict__ segs, const int32_t* __restrict__ thr_early,
                               const int32_t* __restrict__ thr_now, uint32_t* __restrict__ touched, int32_t* __restrict__ list,
                               unsigned* __restrict__ list_n) {
    const SegDev S = segs[blockIdx.x];
    if (thr_early[S.contig] == thr_now[S.contig]) return;
    for (int64_t tile = S.tile_off + threadIdx.x; tile < S.tile_off + S.n_tiles; tile += blockDim.x) {
        const uint32_t bit = 1u << (tile & 31);
        const uint32_t old = atomicOr(&touched[tile >> 5], bit);
        if (!(old & bit)) list[atomicAdd(list_n, 1u)] = (int32_t)tile;
    }
}

constexpr size_t sbt_smem_bytes(bool multi, int SBT_STAGES) {
    return (size_t)SBT_STAGES * (multi ? SBT_STAGE_BYTES_MULTI : SBT_STAGE_BYTES) + SBT_STAGES * sizeof(TileDesc) +
           2 * SBT_STAGES * sizeof(unsigned long long) + 2 * (TILE / 4) * sizeof(double) + 4 * FREEZE * sizeof(uint32_t) +
           32 * sizeof(unsigned) + 64;
}

// per-contig dropout threshold from the running depth total (reference.py:157-158,175-177):
// mean = total / (L * nb); active iff mean > 5; thr = int(mean / 8)
__global__ void k_drop_thresholds(int n_contigs, const int64_t* __restrict__ contig_len, int nb,
                                  const unsigned long long* __restrict__ cov_total, int32_t* __restrict__ thr) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_contigs) return;
    double mean = (double)cov_total[k] / (double)(contig_len[k] * (int64_t)nb);
    thr[k] = mean > 5.0 ? (int32_t)(mean / 8.0) : -1;
}

// bucket switches (reference.py:199-211): mean depth of each complete bucket, last entry repeats its
// predecessor (adjust_length, utils.py:215-217), sticky OR with the previous state
__global__ void k_buckets(const SegDev* __restrict__ segs, int n_seg, int nb, double threshold,
                          const unsigned long long* __restrict__ bucket_sum, uint8_t* __restrict__ sw,
                          int32_t* __restrict__ switched_on) {
    int sg = blockIdx.y;
    const SegDev S = segs[sg];
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= S.n_sw * nb) return;
    int64_t e = i / nb;
    int b = (int)(i % nb);
    int64_t src = e < S.n_full_buckets ? e : S.n_full_buckets - 1;
    bool on = false;
    if (src >= 0) {
        double mean = (double)bucket_sum[(size_t)(S.sw_off + src) * nb + b] / (double)BUCKET;
        on = mean >= threshold;
    }
    size_t idx = (size_t)(S.sw_off + e) * nb + b;
    if (on) sw[idx] = 1;
    if (sw[idx]) atomicMax(switched_on, 1);
}

// ---- on-demand materialisation of Contig.scores / Contig.entropy for the getters -------------------
__global__ void k_materialise_scores(SegDev S, int nb, int64_t P, const uint8_t* __restrict__ ref,
                                     const uint16_t* __restrict__ cov, const double* __restrict__ table,
                                     const double* __restrict__ etable, const int32_t* __restrict__ drop_thr,
                                     double score0, double ent0, double* __restrict__ scores,
                                     double* __restrict__ entropy) {
    int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= S.len) return;
    const int32_t thr = drop_thr[S.contig];
    uint32_t row_min = 0xFFFFFFFFu, any = 0;
    for (int b = 0; b < nb; ++b) {
        uint32_t cs = 0;
        for (int base = 0; base < 5; ++base) cs += cov[((size_t)(b * 5 + base)) * P + S.site_off + l];
        row_min = min(row_min, cs);
        any |= cs;
    }
    const bool dropped = thr >= 0 && row_min <= (uint32_t)thr;
    const unsigned refb = ref[S.site_off + l];
    for (int b = 0; b < nb; ++b) {
        uint32_t c[5], cs = 0;
        for (int base = 0; base < 5; ++base) { c[base] = cov[((size_t)(b * 5 + base)) * P + S.site_off + l]; cs += c[base]; }
        double s, e;
        if (cs >= (uint32_t)FREEZE) { s = TINY; e = nan(""); }
        else if (!any) { s = score0; e = ent0; }
        else {
            size_t r = (size_t)pattern_rank(c[0], c[1], c[2], c[3], c[4]) * 4 + refb;
            s = table[r]; e = etable[r];
        }
        if (dropped) s = 0.0;
        if (scores) scores[(size_t)l * nb + b] = s;
        if (entropy) entropy[(size_t)l * nb + b] = e;
    }
}

}  // namespace boss
