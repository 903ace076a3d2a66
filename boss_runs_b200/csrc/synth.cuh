// synth.cuh — layout conversion helpers for the getters/setters and the synthetic sequencing-state
// generator used by bench.py (BASELINE.json config 3 cannot be prepared on a host: ~150 GB upstream).
#pragma once
#include "common.cuh"

namespace boss {

__global__ void k_add_u64(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// the same, unless an earlier stage of the ingest rejected the batch (nothing may reach the state then)
__global__ void k_add_u64_if_ok(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, int n,
                                const int32_t* __restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && *err == 0) dst[i] += src[i];
}

// device planes [b][base][site]  ->  reference layout [site][base][b] (reference.py:77)
__global__ void k_cov_to_ref_layout(SegDev S, int nb, int64_t P, const uint16_t* __restrict__ cov, uint16_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= S.len * 5 * nb) return;
    int b = (int)(i % nb);
    int base = (int)((i / nb) % 5);
    int64_t l = i / (5 * nb);
    out[i] = cov[((size_t)(b * 5 + base)) * P + S.site_off + l];
}
__global__ void k_cov_from_ref_layout(SegDev S, int nb, int64_t P, const uint16_t* __restrict__ in, uint16_t* __restrict__ cov) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= S.len * 5 * nb) return;
    int b = (int)(i % nb);
    int base = (int)((i / nb) % 5);
    int64_t l = i / (5 * nb);
    cov[((size_t)(b * 5 + base)) * P + S.site_off + l] = in[i];
}

__global__ void k_depth_total(SegDev S, int nb, int64_t P, const uint16_t* __restrict__ cov, unsigned long long* __restrict__ total) {
    unsigned long long acc = 0;
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < S.len; l += (int64_t)gridDim.x * blockDim.x)
        for (int p = 0; p < 5 * nb; ++p) acc += cov[(size_t)p * P + S.site_off + l];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&total[S.contig], acc);
}

// src is [nb][stride] with the wanted run at [off, off+n*inner) per barcode; out is [n][inner][nb] flattened
// (inner = 1 for scores_ds, 2 for the per-strand arrays)
__global__ void k_gather_bins(const double* __restrict__ src, int64_t stride, int64_t off, int64_t n_inner, int nb, int inner,
                              double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_inner * nb) return;
    int b = (int)(i % nb);
    int64_t e = i / nb;          // index over [n][inner]
    out[i] = src[(size_t)b * stride + off + e];
}

// ---- synthetic sequencing state -------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t& st) {
    st = mix64(st);
    return (double)(st >> 11) * (1.0 / 9007199254740992.0);
}

// depth ~ Poisson(mean) per site, split over bases (p_ref on the reference base, p_del deletions, rest
// uniform substitutions); whole 20 kb regions are zero-depth with probability frac_dropout and get
// depth + 30 with probability frac_deep. Deterministic in (seed, contig, position, barcode).
__global__ void k_synth_coverage(SegDev S, int nb, int64_t P, const uint8_t* __restrict__ ref, uint16_t* __restrict__ cov,
                                 unsigned long long* __restrict__ total, uint64_t seed, double mean_depth, double p_ref,
                                 double p_del, double frac_dropout, double frac_deep) {
    unsigned long long acc = 0;
    const double lim = exp(-mean_depth);
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < S.len; l += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = S.start + l;
        uint64_t rs = mix64(seed ^ ((uint64_t)S.contig << 40) ^ (uint64_t)(pos / BUCKET));
        const double region = u01(rs);
        const unsigned refb = ref[S.site_off + l];
        for (int b = 0; b < nb; ++b) {
            uint64_t st = mix64(seed * 0x100000001B3ull + ((uint64_t)S.contig << 44) + (uint64_t)pos * 131 + b);
            int depth = 0;
            if (region >= frac_dropout) {
                double prod = u01(st);                      // Knuth; mean depth is small
                while (prod > lim && depth < 200) { prod *= u01(st); ++depth; }
                if (region < frac_dropout + frac_deep) depth += FREEZE;
            }
            unsigned c[5] = {0, 0, 0, 0, 0};
            for (int k = 0; k < depth; ++k) {
                double u = u01(st);
                unsigned base;
                if (u < p_ref) base = refb;
                else if (u < p_ref + p_del) base = 4;
                else { base = (refb + 1 + (unsigned)((u - p_ref - p_del) / (1.0 - p_ref - p_del) * 3.0)) & 3u; }
                c[base]++;
            }
            for (int base = 0; base < 5; ++base) cov[((size_t)(b * 5 + base)) * P + S.site_off + l] = (uint16_t)c[base];
            acc += depth;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&total[S.contig], acc);
}

}  // namespace boss
