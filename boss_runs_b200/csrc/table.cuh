// table.cuh — dense score table over the C(34,5) count patterns with sum <= 29, built on the device.
//
// Replaces Scoring.init_score_array / calc_posterior / calc_score
// (/root/reference boss/runs/sequences.py:347-393, 485-516, 520-549). The arithmetic is kept in the
// reference's order with explicitly un-fused fp64 ops (__dmul_rn/__dadd_rn) so that the only source
// of difference from NumPy is the last-ulp behaviour of log().
#pragma once
#include "common.cuh"

namespace boss {

// C(q,k) for q <= 34
__host__ __device__ inline int64_t binom2(int64_t q) { return q * (q - 1) / 2; }
__host__ __device__ inline int64_t binom3(int64_t q) { return q * (q - 1) * (q - 2) / 6; }
__host__ __device__ inline int64_t binom4(int64_t q) { return q * (q - 1) * (q - 2) * (q - 3) / 24; }
__host__ __device__ inline int64_t binom5(int64_t q) { return q * (q - 1) * (q - 2) * (q - 3) * (q - 4) / 120; }

// rank of (c0..c4), sum <= 29: q_k = prefix_k + (k-1) is a strictly increasing 5-subset of [0,33]
__host__ __device__ inline int32_t pattern_rank(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4) {
    uint32_t q1 = c0;
    uint32_t q2 = q1 + c1 + 1;
    uint32_t q3 = q2 + c2 + 1;
    uint32_t q4 = q3 + c3 + 1;
    uint32_t q5 = q4 + c4 + 1;
    // all intermediate products fit in 32 bits for q <= 33
    return (int32_t)(q1 + q2 * (q2 - 1) / 2 + q3 * (q3 - 1) * (q3 - 2) / 6 +
                     q4 * (q4 - 1) * (q4 - 2) * (q4 - 3) / 24 +
                     q5 * (q5 - 1) * (q5 - 2) * (q5 - 3) / 24 * (q5 - 4) / 5);
}

__device__ inline void pattern_unrank(int32_t r, int c[5]) {
    int q[6];
    int64_t rem = r;
    int hi = 33;
    for (int k = 5; k >= 1; --k) {
        int qq = hi;
        for (;; --qq) {
            int64_t b = k == 5 ? binom5(qq) : k == 4 ? binom4(qq) : k == 3 ? binom3(qq) : k == 2 ? binom2(qq) : qq;
            if (b <= rem) { rem -= b; break; }
        }
        q[k] = qq;
        hi = qq - 1;
    }
    int p1 = q[1], p2 = q[2] - 1, p3 = q[3] - 2, p4 = q[4] - 3, p5 = q[5] - 4;
    c[0] = p1; c[1] = p2 - p1; c[2] = p3 - p2; c[3] = p4 - p3; c[4] = p5 - p4;
}

// NumPy's float64 add-reduce over a contiguous row of n < 128 elements: 0 + pairwise(all); the
// pairwise kernel is a plain loop for n < 8 and an 8-accumulator block followed by a tail loop otherwise.
__device__ inline double np_row_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 = __dadd_rn(r0, a[i]);     r1 = __dadd_rn(r1, a[i + 1]); r2 = __dadd_rn(r2, a[i + 2]);
        r3 = __dadd_rn(r3, a[i + 3]); r4 = __dadd_rn(r4, a[i + 4]); r5 = __dadd_rn(r5, a[i + 5]);
        r6 = __dadd_rn(r6, a[i + 6]); r7 = __dadd_rn(r7, a[i + 7]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return __dadd_rn(0.0, res);
}

// score and entropy of one posterior vector (sequences.py:520-549)
__device__ inline void score_entropy(const double* p, int len_g, const double* phi, double* score, double* entropy) {
    double logs[15], t[15], np_[15];
    for (int j = 0; j < len_g; ++j) {
        logs[j] = p[j] > 0.0 ? log(p[j]) : 0.0;
        t[j] = __dmul_rn(-p[j], logs[j]);
    }
    double ent = np_row_sum(t, len_g);
    double new_ent = 0.0;
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < len_g; ++j) np_[j] = __dmul_rn(p[j], phi[i * len_g + j]);
        double o = np_row_sum(np_, len_g);
        if (o == 0.0) o = 1e-300;
        for (int j = 0; j < len_g; ++j) {
            np_[j] = __ddiv_rn(np_[j], o);
            if (np_[j] > 0.0) logs[j] = log(np_[j]);      // masked log: stale slot is multiplied by 0
        }
        for (int j = 0; j < len_g; ++j)
            new_ent = __dadd_rn(new_ent, -__dmul_rn(__dmul_rn(o, np_[j]), logs[j]));
    }
    *entropy = ent;
    *score = __dadd_rn(ent, -new_ent);
}

// one thread per pattern; writes table[rank][ref] for the four reference bases
__global__ void k_build_table(int len_g, const double* __restrict__ phi, const double* __restrict__ priors,
                              const double* __restrict__ phi_pow, double* __restrict__ table,
                              double* __restrict__ etable) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= NPAT) return;
    int c[5];
    pattern_unrank(r, c);
    double lik[15];
    for (int j = 0; j < len_g; ++j) {
        double l = 1.0;
        for (int i = 0; i < 5; ++i) l = __dmul_rn(l, phi_pow[(i * len_g + j) * FREEZE + c[i]]);   // sequences.py:506-507
        lik[j] = l;
    }
    for (int h = 0; h < 4; ++h) {
        double post[15];
        for (int j = 0; j < len_g; ++j) post[j] = __dmul_rn(priors[h * len_g + j], lik[j]);
        double z = np_row_sum(post, len_g);
        if (z < 1e-300) z = 1e-300;                                                           // :514
        for (int j = 0; j < len_g; ++j) post[j] = __ddiv_rn(post[j], z);
        double s, e;
        score_entropy(post, len_g, phi, &s, &e);
        table[(size_t)r * 4 + h] = s;
        etable[(size_t)r * 4 + h] = e;
    }
}

}  // namespace boss
