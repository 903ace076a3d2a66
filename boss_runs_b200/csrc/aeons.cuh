// aeons.cuh — BOSS-AEONS' benefit / threshold step on the smoothing and exponent-histogram machinery of the RUNS path.
//
// Replaces, for a pool of contigs given as per-node scores (one value per 100-bp node):
//   Benefit.calc_fragment_benefit / _expand_scores / _calc_smu_moving / _calc_benefit_moving   boss/aeons/sequences.py:1555-1641
//   Benefit.benefit_bins                                                                       boss/aeons/sequences.py:1644-1682
//   ContigPool.find_threshold                                                                  boss/aeons/sequences.py:1059-1094
//   Sequence.find_strat_m0                                                                     boss/aeons/sequences.py:398-406
//
// Geometry of one contig of n nodes, C = ccl_ds[9]: padded axis of N = n + 2C entries, sx[u] = e1 for u < C, the scores for
// C <= u < C + n, e2 for C + n <= u < N - 1 and 0 at N - 1 (upstream fills `scoresx[-ccl_max:-1]`). At padded position p:
//   forward benefit  = sum_i perc_i * (sx[p+1] + ... + sx[p+w_i])   where p < N - w_i - 1
//   reverse benefit  = sum_i perc_i * (sx[p-w_i+1] + ... + sx[p])   where w_i <= p < N - 1
//   smu forward      = sx[p-mu+1] + ... + sx[p]                      (partial at the left edge)
//   smu reverse      = sx[N-1-p] + ... + sx[N-1-p+mu-1]              (upstream never mirrors this row back: it is paired
//                                                                    with the window starting at the MIRRORED position)
// and the result is max(benefit - smu, 0) at p = C .. C + n - 1. Every window is summed from the scores themselves, the ten
// nested windows as one running sum per strand.
#pragma once
#include "common.cuh"
#include "strategy.cuh"

namespace boss {

constexpr int AE_TILE = 256;

struct AeonsArgs {
    int64_t n_seq;
    const int64_t* off;           // [n_seq + 1] node offsets
    const double* scores;         // concatenated
    const uint8_t* e1;
    const uint8_t* e2;
    const int32_t* tile_seq;      // [n_tiles]
    const int32_t* tile_j0;       // [n_tiles]
    int32_t mu;
    int32_t w[NSTEPS];
    double perc[NSTEPS];
    double* benefit;              // sequence i: forward row at 2*off[i], reverse row at 2*off[i] + n_i
    unsigned long long* norm_bits;
};

__device__ __forceinline__ double aeons_sx(int64_t u, int64_t n, int64_t C, const double* __restrict__ sc, bool e1, bool e2) {
    if (u < 0) return 0.0;
    if (u < C) return e1 ? 1.0 : 0.0;
    if (u < C + n) return sc[u - C];
    if (u < n + 2 * C - 1) return e2 ? 1.0 : 0.0;
    return 0.0;
}

// one CTA = AE_TILE consecutive nodes of one contig; dynamic smem: (AE_TILE + 2 * wmax) doubles
__global__ void __launch_bounds__(AE_TILE)
k_aeons_benefit(AeonsArgs a) {
    extern __shared__ double s_sx[];
    __shared__ unsigned long long s_max;
    const int sq = a.tile_seq[blockIdx.x];
    const int64_t j0 = a.tile_j0[blockIdx.x];
    const int64_t o = a.off[sq], n = a.off[sq + 1] - o;
    const int64_t C = a.w[NSTEPS - 1], N = n + 2 * C;
    const int wmax = max((int)C, a.mu);
    const double* sc = a.scores + o;
    const bool e1 = a.e1[sq] != 0, e2 = a.e2[sq] != 0;
    const int t = threadIdx.x;
    if (t == 0) s_max = 0ull;
    // stage sx[P0 - wmax, P0 + AE_TILE + wmax) where P0 = C + j0 is the tile's first padded position
    const int64_t P0 = C + j0;
    const int span = AE_TILE + 2 * wmax;
    for (int i = t; i < span; i += AE_TILE) s_sx[i] = aeons_sx(P0 - wmax + i, n, C, sc, e1, e2);
    __syncthreads();
    unsigned long long bits = 0ull;
    const int64_t j = j0 + t;
    if (j < n) {
        const int64_t p = P0 + t;
        const double* c = s_sx + wmax + t;                 // c[k] = sx[p + k]
        double run_f = 0.0, run_r = 0.0, ben_f = 0.0, ben_r = 0.0;
        int k = 0;
#pragma unroll 1
        for (int i = 0; i < NSTEPS; ++i) {
            const int w = a.w[i];
            for (; k < w; ++k) { run_f += c[k + 1]; run_r += c[-k]; }
            if (p < N - w - 1) ben_f += run_f * a.perc[i];
            if (p >= w && p < N - 1) ben_r += run_r * a.perc[i];
        }
        double smu_f = 0.0, smu_r = 0.0;
        for (int q = 0; q < a.mu; ++q) {
            smu_f += c[-q];                                 // positions below 0 were staged as 0
            smu_r += aeons_sx(N - 1 - p + q, n, C, sc, e1, e2);       // 0 at and beyond the last padded entry
        }
        double bf = ben_f - smu_f, br = ben_r - smu_r;
        if (bf < 0.0) bf = 0.0;
        if (br < 0.0) br = 0.0;
        a.benefit[2 * o + j] = bf;
        a.benefit[2 * o + n + j] = br;
        const unsigned long long xf = (unsigned long long)__double_as_longlong(bf), xr = (unsigned long long)__double_as_longlong(br);
        bits = xf > xr ? xf : xr;
    }
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xFFFFFFFFu, bits, d);
        bits = other > bits ? other : bits;
    }
    if ((t & 31) == 0 && bits) atomicMax(&s_max, bits);
    __syncthreads();
    if (t == 0 && s_max) atomicMax(a.norm_bits, s_max);
}

// smu_sum of every contig: the sum of both smu rows over the WHOLE padded axis (sequences.py:1581). Each padded score
// enters min(mu, N - u) forward windows and min(mu, u + 1) mirrored ones. One CTA per contig, fixed summation order.
__global__ void __launch_bounds__(256)
k_aeons_smu_sum(int64_t n_seq, const int64_t* __restrict__ off, const double* __restrict__ scores, const uint8_t* __restrict__ e1v,
                const uint8_t* __restrict__ e2v, int32_t mu, int64_t C, double* __restrict__ smu_sum) {
    __shared__ double s_part[256];
    const int64_t sq = blockIdx.x;
    const int64_t o = off[sq], n = off[sq + 1] - o, N = n + 2 * C;
    const bool e1 = e1v[sq] != 0, e2 = e2v[sq] != 0;
    double acc = 0.0;
    for (int64_t u = threadIdx.x; u < N; u += 256) {
        const double v = aeons_sx(u, n, C, scores + o, e1, e2);
        const int64_t m = min((int64_t)mu, N - u) + min((int64_t)mu, u + 1);
        acc += v * (double)m;
    }
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) s_part[threadIdx.x] += s_part[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) smu_sum[sq] = s_part[0];
}

// exponent histogram of every non-zero benefit (sequences.py:1654-1677): counts only, every node weighs one
__global__ void __launch_bounds__(256)
k_aeons_hist(int64_t n_vals, const double* __restrict__ benefit, const unsigned long long* __restrict__ norm_bits,
             unsigned long long* __restrict__ counts, unsigned long long* __restrict__ n_nonzero) {
    __shared__ unsigned s_cnt[HBINS];
    for (int i = threadIdx.x; i < HBINS; i += 256) s_cnt[i] = 0;
    __syncthreads();
    const unsigned long long nb = *norm_bits;
    const double norm = __longlong_as_double((long long)nb);
    unsigned nnz = 0;
    if (nb != 0ull)
        for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_vals; i += (int64_t)gridDim.x * 256) {
            const double x = benefit[i];
            if (x != 0.0) { atomicAdd(&s_cnt[abs_exponent_of_ratio(x, norm, nb)], 1u); ++nnz; }
        }
    nnz = __reduce_add_sync(0xFFFFFFFFu, nnz);
    if ((threadIdx.x & 31) == 0 && nnz) atomicAdd(n_nonzero, (unsigned long long)nnz);
    __syncthreads();
    for (int i = threadIdx.x; i < HBINS; i += 256)
        if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
}

struct AeonsOut {
    double threshold, normaliser, ubar0;
    int32_t strat_size, empty;
    unsigned long long n_nonzero;
};

// ContigPool.find_threshold (sequences.py:1059-1094): two cumulative sums over the occupied bins and an argmax
__global__ void k_aeons_threshold(int64_t n_seq, const double* __restrict__ smu_sum, const unsigned long long* __restrict__ counts,
                                  const unsigned long long* __restrict__ norm_bits, const unsigned long long* __restrict__ n_nonzero,
                                  double tc, double tbar0, AeonsOut* out) {
    __shared__ int s_e[HBINS];
    __shared__ double s_peak[HBINS];
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double ubar0 = 0.0;
    for (int64_t i = 0; i < n_seq; ++i) ubar0 += smu_sum[i];          // np.sum over the contigs' smu_sum, in order
    const double norm = __longlong_as_double((long long)*norm_bits);
    out->normaliser = norm; out->ubar0 = ubar0; out->n_nonzero = *n_nonzero;
    if (*norm_bits == 0ull || *n_nonzero == 0ull) { out->empty = 1; out->threshold = 0.0; out->strat_size = 0; return; }
    double cs_u = 0.0, cs_t = 0.0;
    int n_occ = 0;
    for (int e = 0; e < HBINS; ++e) {
        if (!counts[e]) continue;                                     // np.nonzero(bincounts)
        const double c = (double)(long long)counts[e];
        cs_u += (ldexp(1.0, -e) * norm) * c;                          // np.cumsum(benefit_bin * counts)
        cs_t += tc * c;                                               // np.cumsum(tc * counts)
        s_e[n_occ] = e;
        s_peak[n_occ] = (cs_u + ubar0) / (cs_t + tbar0);
        ++n_occ;
    }
    double best = s_peak[0];                                          // np.argmax: first maximum, the first NaN beats everything
    int best_i = 0;
    for (int i = 1; i < n_occ && best == best; ++i)
        if (s_peak[i] > best || s_peak[i] != s_peak[i]) { best = s_peak[i]; best_i = i; }
    const int k = best_i + 1;                                         // bin[argmax + 1], the last bin when the peak is the last one
    out->empty = 0;
    out->strat_size = k;
    out->threshold = ldexp(1.0, -s_e[k < n_occ ? k : n_occ - 1]) * norm;
}

// Sequence.find_strat_m0: (n, 2) bool, benefit >= threshold
__global__ void k_aeons_mask(int64_t n_seq, const int64_t* __restrict__ off, const double* __restrict__ benefit, const AeonsOut* __restrict__ out,
                             const int32_t* __restrict__ tile_seq, const int32_t* __restrict__ tile_j0, uint8_t* __restrict__ strat) {
    const int sq = tile_seq[blockIdx.x];
    const int64_t o = off[sq], n = off[sq + 1] - o, j = tile_j0[blockIdx.x] + threadIdx.x;
    if (j >= n) return;
    const double thr = out->threshold;
    strat[2 * (o + j)] = benefit[2 * o + j] >= thr ? 1 : 0;
    strat[2 * (o + j) + 1] = benefit[2 * o + n + j] >= thr ? 1 : 0;
}

}  // namespace boss
