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

struct ScoreArgs {
    const SegDev* segs;
    const int64_t* tile_start;       // [n_seg+1]
    int n_seg;
    int nb;
    int64_t P;
    const uint8_t* ref;
    const uint16_t* cov;
    const uint32_t* rowflag;         // nb > 1: (touched << 31) | min over barcodes of depth
    const double* table;
    const int32_t* drop_thr;         // per global contig; -1 = rule inactive
    double score0;
    double* ds;                      // [nb][ds_len]
    int64_t ds_len;
    unsigned long long* bucket_sum;  // [n_sw][nb]
    unsigned long long* n_dropout;
};

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

__device__ __forceinline__ double site_score(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4,
                                             uint32_t cs, bool dropped, bool touched, unsigned refb,
                                             const double* __restrict__ table, double score0) {
    if (dropped) return 0.0;
    if (cs >= (uint32_t)FREEZE) return TINY;
    if (!touched) return score0;
    return __ldg(table + (size_t)pattern_rank(c0, c1, c2, c3, c4) * 4 + refb);
}

// grid = (tiles, nb); one CTA = 2000 consecutive sites of one segment for one barcode
template <bool MULTI>
__global__ void __launch_bounds__(TILE_THREADS)
k_score_bin(ScoreArgs a) {
    __shared__ double s_part[TILE / 4];
    __shared__ unsigned s_cov[TILE_THREADS / 32];
    __shared__ unsigned s_drop[TILE_THREADS / 32];

    const int64_t tile = blockIdx.x;
    const int b = blockIdx.y;
    const int sg = find_segment(a.tile_start, a.n_seg, tile);
    const SegDev S = a.segs[sg];
    const int64_t local0 = (tile - S.tile_off) * TILE;         // first site of the tile within the segment
    const int t = threadIdx.x;
    const int32_t thr = a.drop_thr[S.contig];
    const bool rule = thr >= 0;

    unsigned covsum = 0, ndrop = 0;
    if (t < TILE / 4) {
        const int64_t l = local0 + 4 * t;
        double part = 0.0;
        if (l < S.len) {
            const size_t g = (size_t)(S.site_off + l) >> 2;
            const uint16_t* plane = a.cov + (size_t)b * 5 * a.P;
            uint2 v0 = __ldg(reinterpret_cast<const uint2*>(plane) + g);
            uint2 v1 = __ldg(reinterpret_cast<const uint2*>(plane + a.P) + g);
            uint2 v2 = __ldg(reinterpret_cast<const uint2*>(plane + 2 * a.P) + g);
            uint2 v3 = __ldg(reinterpret_cast<const uint2*>(plane + 3 * a.P) + g);
            uint2 v4 = __ldg(reinterpret_cast<const uint2*>(plane + 4 * a.P) + g);
            uint32_t rb = __ldg(reinterpret_cast<const uint32_t*>(a.ref) + g);
            uint4 rf = make_uint4(0, 0, 0, 0);
            if (MULTI) rf = __ldg(reinterpret_cast<const uint4*>(a.rowflag) + g);
            const uint32_t rfv[4] = {rf.x, rf.y, rf.z, rf.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t w0 = i < 2 ? v0.x : v0.y, w1 = i < 2 ? v1.x : v1.y, w2 = i < 2 ? v2.x : v2.y,
                               w3 = i < 2 ? v3.x : v3.y, w4 = i < 2 ? v4.x : v4.y;
                const int sh = (i & 1) * 16;
                const uint32_t c0 = (w0 >> sh) & 0xFFFFu, c1 = (w1 >> sh) & 0xFFFFu, c2 = (w2 >> sh) & 0xFFFFu,
                               c3 = (w3 >> sh) & 0xFFFFu, c4 = (w4 >> sh) & 0xFFFFu;
                const uint32_t cs = c0 + c1 + c2 + c3 + c4;
                const bool in = l + i < S.len;
                uint32_t row_min = cs;
                bool touched = cs > 0;
                if (MULTI) { row_min = rfv[i] & 0x7FFFFFFFu; touched = (rfv[i] >> 31) != 0; }
                const bool dropped = rule && row_min <= (uint32_t)thr;
                const double s = site_score(c0, c1, c2, c3, c4, cs, dropped, touched, (rb >> (8 * i)) & 0xFFu,
                                            a.table, a.score0);
                if (in) {
                    part += s;                  // sites of one thread are added in position order
                    covsum += cs;
                    ndrop += dropped ? 1u : 0u;
                }
            }
        }
        s_part[t] = part;
    }
    // block sums of depth (for the bucket means) and of dropped rows (for the log line)
    for (int o = 16; o > 0; o >>= 1) {
        covsum += __shfl_down_sync(0xFFFFFFFFu, covsum, o);
        ndrop += __shfl_down_sync(0xFFFFFFFFu, ndrop, o);
    }
    if ((t & 31) == 0) { s_cov[t >> 5] = covsum; s_drop[t >> 5] = ndrop; }
    __syncthreads();

    if (t < TILE / BIN) {
        // bin j of the tile = 25 consecutive thread partials, added in position order
        const int64_t bin = local0 / BIN + t;
        if (bin < S.n_bins) {
            double acc = 0.0;
#pragma unroll 5
            for (int k = 0; k < 25; ++k) acc += s_part[25 * t + k];
            a.ds[(size_t)b * a.ds_len + S.ds_off + bin] = acc;
        }
    } else if (t == 32) {
        unsigned long long tot = 0, dr = 0;
        for (int w = 0; w < TILE_THREADS / 32; ++w) { tot += s_cov[w]; dr += s_drop[w]; }
        const int64_t bucket = local0 / BUCKET;
        if (bucket < S.n_full_buckets && tot) atomicAdd(&a.bucket_sum[(size_t)(S.sw_off + bucket) * a.nb + b], tot);
        if (dr && b == 0) atomicAdd(a.n_dropout, dr);
    }
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
