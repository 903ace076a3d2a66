// strategy.cuh — benefit smoothing on the binned scores and strategy derivation.
//
// Replaces Contig.calc_smu / calc_u (boss/runs/reference.py:215-269; bn.move_sum box sums),
// ReadStartDist._expand_fhat (boss/runs/readstartdist.py:121-152), Scoring.merge_benefit +
// adjust_length (boss/runs/sequences.py:553-560, boss/utils.py:206-226), Scoring.find_strat_thread
// (sequences.py:566-649) and BossRuns._distribute_strategy (boss/runs/core.py:125-155).
//
// Axes.  "merged row" r: row of the concatenated per-contig benefit arrays (L_c//100 + 1 rows per
// contig).  "strategy row" d: row of the concatenated Contig.strat arrays (L_c//100 rows per contig).
// Upstream slices the merged mask with strategy-row offsets (core.py:141,155), so strategy row d is
// simply merged row d (parity quirk Q2) — contig k's mask is shifted by k bins.
//
// Determinism: every cross-thread floating-point sum that feeds the threshold is accumulated in exact
// integer limbs (value * 2^shift split into a 64-bit integer part and 32 fractional bits), so the
// histogram, ubar0 and the F-hat normaliser do not depend on thread order, grid size or shard count.
#pragma once
#include "common.cuh"
#include "score_pass.cuh"

namespace boss {

constexpr int SM_TILE = 1024;      // output bins per CTA in the smoothing kernels
constexpr int SM_THREADS = 256;
constexpr int SM_OUT = SM_TILE / SM_THREADS;   // outputs per thread (t, t+256, ...): piece offsets are fetched once for all of them

constexpr int SM_MAX_PIECES = 192;
constexpr int SM_MAX_LEVELS = 14;

struct SmoothArgs {
    const SegDev* segs;
    int n_seg, nb;
    const double* ds;
    int64_t ds_len;
    double2* benefit;             // [nb][n_rows]
    double2* smu;                 // optional
    double2* expected;            // optional
    int64_t n_rows;
    double mult[NSTEPS];
    int32_t wmax;                 // max(w[9], 4)
    int32_t n_levels;             // sliding power-of-two sums kept in shared memory: widths 1, 2, ..., 2^(n_levels-1)
    // the ten nested boxes as a list of power-of-two pieces: piece p adds the 2^lvl bins starting `off` bins ahead
    // (forward strand) / ending `off` bins behind (reverse strand); both are stored as element offsets into the
    // level arrays (lvl * span + off, lvl * span - off - 2^lvl + 1). Step i ends before piece step_end[i].
    int32_t step_end[NSTEPS];
    int32_t piece_f[SM_MAX_PIECES];
    int32_t piece_r[SM_MAX_PIECES];
    int64_t R0, target_rows;      // rows >= target are cut by adjust_length (core.py:179-181)
    UpdateDev* upd;
};

// Host side: split the increments between consecutive staircase windows into power-of-two pieces.
inline bool plan_smoothing(const int32_t w[NSTEPS], int max_levels, int span, SmoothArgs& a) {
    int n = 0, prev = 0;
    for (int i = 0; i < NSTEPS; ++i) {
        int off = prev, d = w[i] - prev;
        while (d > 0) {
            int k = 0;
            while (k + 1 < max_levels && (2 << k) <= d) ++k;     // largest kept width <= d
            if (n >= SM_MAX_PIECES) return false;
            a.piece_f[n] = k * span + off;
            a.piece_r[n] = k * span - off - (1 << k) + 1;
            ++n;
            off += 1 << k;
            d -= 1 << k;
        }
        a.step_end[i] = n;
        prev = w[i];
    }
    return true;
}

// grid = (ceil(bins/SM_TILE) per segment flattened, nb).
// Dynamic smem: n_levels * (SM_TILE + 2*(wmax-1)) doubles. Level k holds, for every position, the sum of the
// 2^k bins starting there (built by doubling), so a box of width w costs popcount-many loads instead of w.
// Every window is still summed from the bins themselves — no running accumulator, no cancellation.
__global__ void __launch_bounds__(SM_THREADS)
k_smooth(SmoothArgs a, const int64_t* __restrict__ sm_tile_start) {
    extern __shared__ double s_lv[];
    __shared__ unsigned long long s_max;
    if (!a.upd->switched_on) return;           // every bucket still off: the strategy half is skipped (core.py:172)
    const int b = blockIdx.y;
    const int sg = find_segment(sm_tile_start, a.n_seg, blockIdx.x);
    const SegDev S = a.segs[sg];
    const int64_t j0 = (blockIdx.x - sm_tile_start[sg]) * SM_TILE;
    const int halo = a.wmax - 1;
    const int span = SM_TILE + 2 * halo;
    const double* src = a.ds + (size_t)b * a.ds_len + S.ds_off;
    if (threadIdx.x == 0) s_max = 0ull;
    // stage ds[j0-halo, j0+SM_TILE+halo); outside the contig (or the exchanged halo) the box sums see 0,
    // which is what min_count=1 partial windows amount to
    for (int i = threadIdx.x; i < span; i += SM_THREADS) {
        int64_t j = j0 - halo + i;
        double v = 0.0;
        if (j >= -(int64_t)S.halo_l && j < S.n_bins + S.halo_r) v = src[j];
        s_lv[i] = v;
    }
    __syncthreads();
    for (int k = 1; k < a.n_levels; ++k) {
        const double* lo = s_lv + (size_t)(k - 1) * span;
        double* hi = s_lv + (size_t)k * span;
        const int half = 1 << (k - 1);
        for (int i = threadIdx.x; i + 2 * half <= span; i += SM_THREADS) hi[i] = lo[i] + lo[i + half];
        __syncthreads();
    }
    const double* base = s_lv + halo + threadIdx.x;            // this thread's outputs sit at base[256 * m]
    double run_f[SM_OUT], run_r[SM_OUT], eb_f[SM_OUT], eb_r[SM_OUT];
#pragma unroll
    for (int m = 0; m < SM_OUT; ++m) { run_f[m] = 0.0; run_r[m] = 0.0; eb_f[m] = 0.0; eb_r[m] = 0.0; }
    // staircase: sum_i mult_i * box_{w_i}; the boxes are nested, so one running sum per strand and output
    int p = 0;
#pragma unroll 1
    for (int i = 0; i < NSTEPS; ++i) {
        const int pe = a.step_end[i];
#pragma unroll 2
        for (; p < pe; ++p) {
            const double* Lf = base + a.piece_f[p];
            const double* Lr = base + a.piece_r[p];
#pragma unroll
            for (int m = 0; m < SM_OUT; ++m) { run_f[m] += Lf[SM_THREADS * m]; run_r[m] += Lr[SM_THREADS * m]; }
        }
        const double mu = a.mult[i];
#pragma unroll
        for (int m = 0; m < SM_OUT; ++m) { eb_f[m] += run_f[m] * mu; eb_r[m] += run_r[m] * mu; }
    }
    unsigned long long mybits = 0ull;
    const double* L2 = base + 2 * span;                        // sums of 4 bins starting at a position (n_levels >= 3)
#pragma unroll
    for (int m = 0; m < SM_OUT; ++m) {
        const int64_t j = j0 + threadIdx.x + SM_THREADS * m;
        if (j >= S.n_bins) continue;
        // S_mu: 4-bin forward / backward box (reference.py:233-237)
        const double smu_f = L2[SM_THREADS * m], smu_r = L2[SM_THREADS * m - 3];
        double ad_f = eb_f[m] - smu_f, ad_r = eb_r[m] - smu_r;
        if (ad_f < 0.0) ad_f = 0.0;                   // reference.py:267-269
        if (ad_r < 0.0) ad_r = 0.0;
        const size_t o = (size_t)b * a.n_rows + S.row_off + j;
        a.benefit[o] = make_double2(ad_f, ad_r);
        if (a.smu) a.smu[o] = make_double2(smu_f, smu_r);
        if (a.expected) a.expected[o] = make_double2(eb_f[m], eb_r[m]);
        if (a.R0 + S.row_off + j < a.target_rows) {
            // non-negative doubles order like their bit patterns; NaN would sort above everything
            unsigned long long bf = (unsigned long long)__double_as_longlong(ad_f);
            unsigned long long br = (unsigned long long)__double_as_longlong(ad_r);
            const unsigned long long mx = bf > br ? bf : br;
            mybits = mx > mybits ? mx : mybits;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xFFFFFFFFu, mybits, o);
        mybits = other > mybits ? other : mybits;
    }
    if ((threadIdx.x & 31) == 0 && mybits) atomicMax(&s_max, mybits);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(&a.upd->norm_bits, s_max);
}

// Fallback for very long staircases (ultra-long reads) whose piece list or level arrays do not fit:
// every box is summed bin by bin from one staged copy of the tile.
__global__ void __launch_bounds__(SM_THREADS)
k_smooth_direct(SmoothArgs a, const int64_t* __restrict__ sm_tile_start, const int32_t* __restrict__ w) {
    extern __shared__ double s_lv[];
    __shared__ unsigned long long s_max;
    if (!a.upd->switched_on) return;
    const int b = blockIdx.y;
    const int sg = find_segment(sm_tile_start, a.n_seg, blockIdx.x);
    const SegDev S = a.segs[sg];
    const int64_t j0 = (blockIdx.x - sm_tile_start[sg]) * SM_TILE;
    const int halo = a.wmax - 1;
    const int span = SM_TILE + 2 * halo;
    const double* src = a.ds + (size_t)b * a.ds_len + S.ds_off;
    if (threadIdx.x == 0) s_max = 0ull;
    for (int i = threadIdx.x; i < span; i += SM_THREADS) {
        int64_t j = j0 - halo + i;
        double v = 0.0;
        if (j >= -(int64_t)S.halo_l && j < S.n_bins + S.halo_r) v = src[j];
        s_lv[i] = v;
    }
    __syncthreads();
    unsigned long long mybits = 0ull;
#pragma unroll 1
    for (int m = 0; m < SM_OUT; ++m) {
        const int64_t j = j0 + threadIdx.x + SM_THREADS * m;
        if (j >= S.n_bins) continue;
        const double* c = s_lv + halo + threadIdx.x + SM_THREADS * m;
        const double smu_f = (c[0] + c[1]) + (c[2] + c[3]);
        const double smu_r = (c[-3] + c[-2]) + (c[-1] + c[0]);
        double run_f = 0.0, run_r = 0.0, eb_f = 0.0, eb_r = 0.0;
        int k = 0;
#pragma unroll 1
        for (int i = 0; i < NSTEPS; ++i) {
            const int wi = w[i];
            for (; k < wi; ++k) { run_f += c[k]; run_r += c[-k]; }
            eb_f += run_f * a.mult[i];
            eb_r += run_r * a.mult[i];
        }
        double ad_f = eb_f - smu_f, ad_r = eb_r - smu_r;
        if (ad_f < 0.0) ad_f = 0.0;
        if (ad_r < 0.0) ad_r = 0.0;
        const size_t o = (size_t)b * a.n_rows + S.row_off + j;
        a.benefit[o] = make_double2(ad_f, ad_r);
        if (a.smu) a.smu[o] = make_double2(smu_f, smu_r);
        if (a.expected) a.expected[o] = make_double2(eb_f, eb_r);
        if (a.R0 + S.row_off + j < a.target_rows) {
            unsigned long long bf = (unsigned long long)__double_as_longlong(ad_f);
            unsigned long long br = (unsigned long long)__double_as_longlong(ad_r);
            const unsigned long long mx = bf > br ? bf : br;
            mybits = mx > mybits ? mx : mybits;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_down_sync(0xFFFFFFFFu, mybits, o);
        mybits = other > mybits ? other : mybits;
    }
    if ((threadIdx.x & 31) == 0 && mybits) atomicMax(&s_max, mybits);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(&a.upd->norm_bits, s_max);
}

// ---- exact limb arithmetic ---------------------------------------------------------------------------
// A value v in [0, 2^63) is split into two signed integers, v ~= hi + lo * 2^-32 (lo carries 32 fractional bits,
// what lies below is dropped), which are then summed as 64-bit integers: order-free and exact.
// Below 2^51 the split uses the 2^52-magic-number rounding (a few fp64 adds instead of three 64-bit
// conversions): hi = rint(v), lo = rint((v - hi) * 2^32), lo in [-2^31, 2^31]. Above, hi = floor(v).
// Which branch is taken depends on v alone, so an element contributes the same limbs wherever it is summed.
__device__ __forceinline__ void to_limbs(double v, unsigned long long& hi, unsigned long long& lo) {
    if (v < 2251799813685248.0) {                                    // 2^51
        const double M = 6755399441055744.0;                         // 2^52 + 2^51
        const double t = __dadd_rn(v, M);
        const long long h = __double_as_longlong(t) - __double_as_longlong(M);
        const double rem = __dadd_rn(v, -__dadd_rn(t, -M));          // exact: |rem| <= 0.5
        const double t2 = __dadd_rn(__dmul_rn(rem, 4294967296.0), M);
        hi = (unsigned long long)h;
        lo = (unsigned long long)(__double_as_longlong(t2) - __double_as_longlong(M));
    } else {
        hi = (unsigned long long)v;
        lo = (unsigned long long)((v - (double)hi) * 4294967296.0);
    }
}
// The same for values known to be below 2^51 (every F-hat term: F-hat <= 1 and the scale is at most 2^50).
__device__ __forceinline__ void to_limbs_small(double v, unsigned long long& hi, unsigned long long& lo) {
    const double M = 6755399441055744.0;                             // 2^52 + 2^51
    const double t = __dadd_rn(v, M);
    const long long h = __double_as_longlong(t) - __double_as_longlong(M);
    const double rem = __dadd_rn(v, -__dadd_rn(t, -M));              // exact: |rem| <= 0.5
    const double t2 = __dadd_rn(__dmul_rn(rem, 4294967296.0), M);
    hi = (unsigned long long)h;
    lo = (unsigned long long)(__double_as_longlong(t2) - __double_as_longlong(M));
}
__host__ __device__ inline double from_limbs(unsigned long long hi, unsigned long long lo, int shift) {
    // hi + lo*2^-32 (both signed sums), scaled back by 2^-shift
    double v = (double)(long long)hi + (double)(long long)lo * (1.0 / 4294967296.0);
    return ldexp(v, -shift);
}

// F-hat geometry (all global): W windows, expanded x20, tail-fixed to Tf rows (readstartdist.py:129-140),
// then tail-fixed again to `target` rows by adjust_length (core.py:184-185)
struct FhatGeom {
    int64_t W, Tf, target;
};
__device__ __forceinline__ int64_t fhat_window_of_row(const FhatGeom& g, int64_t r) {
    if (r >= g.Tf) r -= (g.target - g.Tf);              // appended copy of the last target-Tf rows
    const int64_t e = 20 * g.W;
    if (r >= e) r -= (g.Tf - e);                        // appended copy of the last Tf-20W rows
    return r / 20;
}

// exact sum of the expanded, tail-fixed F-hat (before normalisation): one thread per window
__global__ void k_fhat_sum(FhatGeom g, const double* __restrict__ fw, int shift, UpdateDev* upd) {
    int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long hi = 0, lo = 0;
    if (w < g.W) {
        const int64_t e = 20 * g.W;
        const int64_t lim = g.Tf < e ? g.Tf : e;
        // rows of the plain expansion that survive truncation
        int64_t m = 0;
        int64_t a0 = 20 * w, a1 = 20 * w + 20;
        if (a1 > lim) a1 = lim;
        if (a1 > a0) m += a1 - a0;
        // rows appended from the tail [e-d, e)
        if (g.Tf > e) {
            int64_t d = g.Tf - e;
            int64_t s0 = e - d > 20 * w ? e - d : 20 * w, s1 = 20 * w + 20;
            if (s1 > s0) m += s1 - s0;
        }
        if (m > 0) {
            for (int s = 0; s < 2; ++s) {
                unsigned long long h, l;
                to_limbs(ldexp(fw[2 * w + s], shift), h, l);
                hi += h * (unsigned long long)m;
                lo += l * (unsigned long long)m;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        hi += __shfl_down_sync(0xFFFFFFFFu, hi, o);
        lo += __shfl_down_sync(0xFFFFFFFFu, lo, o);
    }
    if ((threadIdx.x & 31) == 0 && (hi | lo)) {
        atomicAdd(&upd->fsum_hi, hi);
        atomicAdd(&upd->fsum_lo, lo);
    }
}

__global__ void k_fhat_finish(int shift, UpdateDev* upd) {
    double s = from_limbs(upd->fsum_hi, upd->fsum_lo, shift);
    upd->fhat_sum = s;
    upd->fhat_scale = s != 0.0 ? 1.0 / s : 1.0;        // readstartdist.py:145-150, on_target = 1
}

// F-hat of the rows [row0, row0 + n) of the merged, length-adjusted axis, exactly as k_hist consumes it: window value
// x normaliser (readstartdist.py:121-152 + adjust_length, core.py:184-185). For bossgpu_get_fhat (parity tests).
__global__ void k_fhat_rows(FhatGeom g, const double* __restrict__ fw, const UpdateDev* __restrict__ upd, int64_t row0, int64_t n,
                            double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    const int64_t r = row0 + (i >> 1);
    out[i] = (r >= 0 && r < g.target) ? fw[2 * fhat_window_of_row(g, r) + (i & 1)] * upd->fhat_scale : 0.0;
}

// ---- exponent histogram (sequences.py:584-624) -----------------------------------------------------------
struct HistArgs {
    const double2* benefit;       // [nb][n_rows]
    int64_t n_rows;               // merged rows owned by this shard
    int nb;
    int64_t R0, M, target;        // global: first row here, total merged rows, rows after adjust_length
    FhatGeom fg;
    const double* fw;             // [W][2]
    int shift;                    // limbs hold fhat * 2^shift
    int ushift;                   // ubar0 is summed as integers of fhat * benefit * 2^(ushift - exponent of the normaliser)
    unsigned long long* hist;     // [3*HBINS + 4]: counts | hi | lo | ubar_hi, ubar_lo, n_nonzero, -
    uint8_t* codes;               // [nb][n_rows][2] exponent-bin code of every entry (see k_hist), or NULL
    UpdateDev* upd;
};

// |e| with x/norm = m * 2^e, 0.5 <= m < 1 (np.frexp + np.abs, sequences.py:589-593), for 0 < x <= norm.
// Both operands normal and the quotient far from the subnormal range: the exponent follows from the two
// exponent fields and one mantissa comparison (the correctly rounded quotient of mantissas stays on its
// side of 1), no division needed. Anything else takes the literal path.
__device__ __forceinline__ int abs_exponent_of_ratio(double x, double norm, unsigned long long nbits) {
    const unsigned long long xb = (unsigned long long)__double_as_longlong(x);
    const int ex = (int)(xb >> 52), en = (int)(nbits >> 52);
    int e;
    if (ex != 0 && en != 0 && ex - en > -1000) {
        e = ex - en + ((xb & 0xFFFFFFFFFFFFFull) >= (nbits & 0xFFFFFFFFFFFFFull) ? 1 : 0);
    } else {
        frexp(x / norm, &e);
    }
    return e < 0 ? -e : e;
}

constexpr int HIST_THREADS = 128;
constexpr int HIST_GROUP = 20;        // rows per thread: one 2 kb read-start window spans 20 bins (readstartdist.py:13,129)
constexpr int HIST_SLOTS = 4;         // exponent bins a thread counts in registers, starting one below its first row's

// Adds one weighted entry per lane (bin e < 0 = nothing) to the CTA's shared histogram: `cnt` elements whose F-hat
// limbs sum to (h, l), h < 2^63, |l| < 2^37. Lanes hold neighbouring row groups, whose smoothed benefits mostly share
// a binary exponent, so the warp first groups its lanes by bin and sums each group with warp reductions; only the
// group's leader touches shared memory (64-bit shared atomics are CAS loops, and 32 lanes hitting one bin would
// serialise 32-fold). Must be called by all 32 lanes.
__device__ __forceinline__ void warp_hist_add(int e, unsigned cnt, unsigned long long h, unsigned long long l, unsigned* s_cnt,
                                              unsigned long long* s_hi, unsigned long long* s_lo) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const unsigned long long BIAS = 1ull << 37;
    const unsigned long long lb = l + BIAS;                    // (0, 2^38): a 21- and a 17-bit piece
    unsigned remaining = __ballot_sync(FULL, e >= 0);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int le = __shfl_sync(FULL, e, leader);
        const bool mine = e == le;
        const unsigned grp = __ballot_sync(FULL, mine);
        const unsigned n = __reduce_add_sync(FULL, mine ? cnt : 0u);
        const unsigned c0 = __reduce_add_sync(FULL, mine ? (unsigned)(h & 0x1FFFFFull) : 0u);
        const unsigned c1 = __reduce_add_sync(FULL, mine ? (unsigned)((h >> 21) & 0x1FFFFFull) : 0u);
        const unsigned c2 = __reduce_add_sync(FULL, mine ? (unsigned)(h >> 42) : 0u);
        const unsigned c3 = __reduce_add_sync(FULL, mine ? (unsigned)(lb & 0x1FFFFFull) : 0u);
        const unsigned c4 = __reduce_add_sync(FULL, mine ? (unsigned)(lb >> 21) : 0u);
        if (lane == leader) {
            atomicAdd(&s_cnt[le], n);
            atomicAdd(&s_hi[le], (unsigned long long)c0 + ((unsigned long long)c1 << 21) + ((unsigned long long)c2 << 42));
            atomicAdd(&s_lo[le], (unsigned long long)c3 + ((unsigned long long)c4 << 21) - (unsigned long long)__popc(grp) * BIAS);
        }
        remaining &= ~grp;
    }
}

// One thread walks the 20 rows of one group of the GLOBAL merged axis (rows 20G .. 20G+19): they share one read-start
// window, so F-hat is converted to limbs once per group and strand, and the bins are counted in registers (a run of
// neighbouring rows spans one or two binary exponents); the warp merges its 32 groups at the end. Entries outside the
// four register slots, and groups that straddle a window edge (only in the tail-fixed part of F-hat), take the direct
// path. Also leaves the bin of every entry behind as a one-byte code (hist.codes) for the distribution kernel:
// 0 = the maximum itself, c = entry in [norm 2^-c, norm 2^-(c-1)), 254 = anything smaller, 255 = zero / cut row.
__global__ void __launch_bounds__(HIST_THREADS, 6)
k_hist(HistArgs a) {
    __shared__ unsigned s_cnt[HBINS];
    __shared__ unsigned long long s_hi[HBINS];
    __shared__ unsigned long long s_lo[HBINS];
    __shared__ unsigned long long s_u[3];
    if (!a.upd->switched_on) return;
    for (int i = threadIdx.x; i < HBINS; i += blockDim.x) { s_cnt[i] = 0; s_hi[i] = 0; s_lo[i] = 0; }
    if (threadIdx.x < 3) s_u[threadIdx.x] = 0;
    __syncthreads();

    const unsigned long long nbits = a.upd->norm_bits;
    const double norm = __longlong_as_double((long long)nbits);
    const double scale = a.upd->fhat_scale;
    int norm_e;
    frexp(norm, &norm_e);                       // norm < 2^norm_e  =>  fhat*b*2^(shift-norm_e) < fhat*2^shift
    const double two_shift = ldexp(1.0, a.shift);
    const int ue = a.ushift - norm_e;          // fhat * benefit * 2^(ushift - norm_e) < fhat * 2^ushift <= 2^ushift
    const bool ue_ok = ue > -1000 && ue < 1000;
    const double two_ue = ue_ok ? ldexp(1.0, ue) : 0.0;
    const int b = blockIdx.y;
    const int64_t extra = a.target > a.M ? a.target - a.M : 0;    // rows duplicated at the tail (adjust_length pads)
    unsigned long long u_sum = 0, nnz = 0;
    const double2* ben = a.benefit + (size_t)b * a.n_rows;
    uint16_t* codes = a.codes ? reinterpret_cast<uint16_t*>(a.codes) + (size_t)b * a.n_rows : nullptr;

    const int64_t n_groups = (a.R0 + a.n_rows - 1) / HIST_GROUP - a.R0 / HIST_GROUP + 1;         // groups of the global row axis here
    for (int64_t g0 = (int64_t)blockIdx.x * HIST_THREADS; g0 < n_groups; g0 += (int64_t)gridDim.x * HIST_THREADS) {
    const int64_t G = a.R0 / HIST_GROUP + g0 + threadIdx.x;                                      // global row group
    const int64_t r_lo = max(G * HIST_GROUP, a.R0), r_hi = min((G + 1) * HIST_GROUP, a.R0 + a.n_rows);
    // register state: the group's window, its F-hat limbs per strand, counts of HIST_SLOTS bins per strand
    int64_t cur_win = -1;
    double f[2] = {0.0, 0.0};
    unsigned long long fh[2] = {0, 0}, fl[2] = {0, 0};
    int ebase[2] = {-1, -1};
    unsigned cnt[2] = {0u, 0u};                 // HIST_SLOTS counters of 8 bits each (a group holds at most 2 x 20 rows)
    auto direct = [&](int e, unsigned n, unsigned long long h, unsigned long long l) {
        atomicAdd(&s_cnt[e], n);
        atomicAdd(&s_hi[e], h * n);
        atomicAdd(&s_lo[e], l * n);
    };
    auto enter_window = [&](int64_t win) {      // empties the register slots and loads the window's F-hat
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int j = 0; j < HIST_SLOTS; ++j) {
                const unsigned n = (cnt[s] >> (8 * j)) & 0xFFu;
                if (n) direct(ebase[s] + j, n, fh[s], fl[s]);
            }
            cnt[s] = 0u;
            f[s] = a.fw[2 * win + s] * scale;   // np.multiply(fhat_exp, normalizer)
            to_limbs_small(f[s] * two_shift, fh[s], fl[s]);
        }
        cur_win = win;
    };
    // pass 0: every row below `target`; pass 1 (only when adjust_length pads, core.py:179-181 with reject refs):
    // the last `extra` merged rows once more, at their appended positions
    const int n_pass = (extra > 0 && r_hi > a.M - extra) ? 2 : 1;
    // in the plain part of F-hat (before either tail fix) the whole group lies in window G: nothing to look up per row
    const bool plain = n_pass == 1 && (G + 1) * HIST_GROUP <= min(min(20 * a.fg.W, a.fg.Tf), a.target);
    if (plain && r_lo < r_hi) enter_window(G);
    for (int pass = 0; pass < n_pass; ++pass) {
        for (int64_t r4 = r_lo; r4 < r_hi; r4 += 4) {
        double2 v4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // four independent loads in flight per thread
            const int64_t r = r4 + j;
            const bool in = r < r_hi && (pass == 0 ? r < a.target : r >= a.M - extra);
            v4[j] = in ? ben[r - a.R0] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = r4 + j;
            const double2 v = v4[j];
            // np.nonzero (sequences.py:585): zero entries (and rows that are cut, which read 0) count nowhere
            if (!plain && !(v.x == 0.0 && v.y == 0.0)) {
                const int64_t win = fhat_window_of_row(a.fg, pass == 0 ? r : r + extra);
                if (win != cur_win) enter_window(win);
            }
            unsigned code2 = 0u;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const double x = s == 0 ? v.x : v.y;
                const bool nz = x != 0.0;
                const int e = abs_exponent_of_ratio(nz ? x : norm, norm, nbits);
                // one-byte bin code: the maximum (the only entry whose ratio has exponent +1, Q3) gets 0
                const unsigned c = !nz ? 255u : (x == norm ? 0u : (unsigned)min(e + 1, 254));
                code2 |= c << (8 * s);
                ebase[s] = (ebase[s] < 0 && nz) ? max(e - 1, 0) : ebase[s];
                const int rel = e - ebase[s];
                const bool slot = nz && (unsigned)rel < (unsigned)HIST_SLOTS;
                cnt[s] += slot ? (1u << (8 * (rel & 3))) : 0u;
                if (nz && !slot) direct(e, 1u, fh[s], fl[s]);
                // term of ubar0 = sum(fhat * smu), smu := benefit (Q1): rounded once to 2^-ushift of the normaliser's
                // binade and summed as an integer (order-free, the same whatever the grid or the number of shards)
                const double t = f[s] * x;
                u_sum += (unsigned long long)__double2ll_rn(ue_ok ? t * two_ue : ldexp(t, ue));
                nnz += nz ? 1ull : 0ull;
            }
            if (pass == 0 && codes && r < r_hi) codes[r - a.R0] = (uint16_t)code2;
        }
        }
    }
    // merge the warp's 32 groups, slot by slot
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < HIST_SLOTS; ++j) {
            const unsigned n = (cnt[s] >> (8 * j)) & 0xFFu;
            warp_hist_add(n ? ebase[s] + j : -1, n, fh[s] * n, fl[s] * n, s_cnt, s_hi, s_lo);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        u_sum += __shfl_down_sync(0xFFFFFFFFu, u_sum, o);
        nnz += __shfl_down_sync(0xFFFFFFFFu, nnz, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_u[0], u_sum); atomicAdd(&s_u[2], nnz); }
    __syncthreads();
    for (int i = threadIdx.x; i < HBINS; i += blockDim.x) {
        if (s_cnt[i]) {
            atomicAdd(&a.hist[i], (unsigned long long)s_cnt[i]);
            atomicAdd(&a.hist[HBINS + i], s_hi[i]);
            atomicAdd(&a.hist[2 * HBINS + i], s_lo[i]);
        }
    }
    if (threadIdx.x < 3 && s_u[threadIdx.x]) atomicAdd(&a.hist[3 * HBINS + threadIdx.x], s_u[threadIdx.x]);
}

// ---- threshold from the histogram: two cumulative sums and an argmax (sequences.py:607-646) ----------------
// One CTA. The per-bin terms (limb -> double, ldexp, divisions) are formed by all threads; the two cumulative
// sums stay one thread's sequential loop over the occupied bins (np.cumsum's order, <= 1076 dependent adds);
// the ratios are again formed by all threads and the first maximum is picked like np.argmax does.
constexpr int THR_THREADS = 256;
__global__ void __launch_bounds__(THR_THREADS)
k_threshold(const unsigned long long* __restrict__ hist, int shift, int ushift, double tc, UpdateDev* upd) {
    __shared__ double s_u[HBINS], s_t[HBINS];        // per-bin terms, then (compacted) cumulative sums, then s_u = ratio
    __shared__ int s_exp[HBINS];
    __shared__ int s_nocc;
    if (!upd->switched_on) return;
    const double norm = __longlong_as_double((long long)upd->norm_bits);
    const unsigned long long nnz = hist[3 * HBINS + 2];
    if (upd->norm_bits == 0ull || nnz == 0ull) {                          // np.max of an empty array upstream
        if (threadIdx.x == 0) {
            upd->normaliser = norm; upd->n_nonzero = nnz;
            upd->empty = 1; upd->threshold = 0.0; upd->strat_size = 0;
        }
        return;
    }
    if (tc != tc) {                                                        // Q14: upstream has no time_cost yet -> AttributeError
        if (threadIdx.x == 0) {                                            // before find_strat_thread; no strategy is touched
            upd->normaliser = norm; upd->n_nonzero = nnz;
            upd->empty = 2; upd->threshold = 0.0; upd->strat_size = 0;
        }
        return;
    }
    int norm_e;
    frexp(norm, &norm_e);
    const double ubar0 = ldexp((double)(long long)hist[3 * HBINS], norm_e - ushift);
    const double tbar0 = 3.0 + 3.0 + 4.0;                               // alpha + rho + mu in bins (sequences.py:578-580,630)
    // per-bin terms by all threads, compacted to the occupied bins (np.nonzero(bincounts), sequences.py:607) in bin order:
    // a chunk of THR_THREADS bins at a time, positions from warp ballots
    __shared__ int s_warp_n[THR_THREADS / 32];
    int base = 0;
    for (int e0 = 0; e0 < HBINS; e0 += THR_THREADS) {
        const int e = e0 + threadIdx.x;
        const unsigned long long cnt = e < HBINS ? hist[e] : 0ull;
        double tu = 0.0, tt = 0.0;
        if (cnt) {
            const double counts = (double)(long long)cnt;
            const double f_grid = from_limbs(hist[HBINS + e], hist[2 * HBINS + e], shift);
            const double f_mean = f_grid / counts;
            const double bin = ldexp(1.0, -e) * norm;                   // np.power(2.0, -e) * normaliser
            tu = (bin * f_mean) * counts;                               // benefit_bin * f_grid_mean * counts
            tt = (tc * counts) * f_mean;                                // tc * counts * f_grid_mean
        }
        const unsigned occ = __ballot_sync(0xFFFFFFFFu, cnt != 0ull);
        if ((threadIdx.x & 31) == 0) s_warp_n[threadIdx.x >> 5] = __popc(occ);
        __syncthreads();
        int pos = base;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pos += s_warp_n[w];
        int total = 0;
        for (int w = 0; w < THR_THREADS / 32; ++w) total += s_warp_n[w];
        if (cnt) {
            pos += __popc(occ & ((1u << (threadIdx.x & 31)) - 1u));
            s_u[pos] = tu; s_t[pos] = tt; s_exp[pos] = e;
        }
        base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double cs_u = 0.0, cs_t = 0.0;
        for (int i = 0; i < base; ++i) {
            cs_u += s_u[i];                                             // np.cumsum: sequential, in bin order
            cs_t += s_t[i];
            s_u[i] = cs_u; s_t[i] = cs_t;
        }
        s_nocc = base;
    }
    __syncthreads();
    const int n_occ = s_nocc;
    for (int i = threadIdx.x; i < n_occ; i += THR_THREADS) s_u[i] = (s_u[i] + ubar0) / (s_t[i] + tbar0);
    __syncthreads();
    if (threadIdx.x == 0) {
        // np.argmax: first maximum, and the first NaN beats everything
        double best = s_u[0];
        int best_i = 0;
        for (int i = 1; i < n_occ && best == best; ++i) {
            const double p = s_u[i];
            if (p > best || p != p) { best = p; best_i = i; }
        }
        // threshold = bin[argmax + 1], or the last bin when argmax is the last (sequences.py:643-646)
        const int k = best_i + 1;
        const int e_thr = k < n_occ ? s_exp[k] : s_exp[n_occ - 1];
        upd->normaliser = norm;
        upd->n_nonzero = nnz;
        upd->ubar0 = ubar0;
        upd->strat_size = k;
        upd->threshold = ldexp(1.0, -e_thr) * norm;
        upd->e_thr = e_thr;
        // the bin codes of k_hist order like the benefits as long as the threshold bin is below their saturation
        // point and the normaliser is a normal number (its exponent field drives the codes)
        upd->use_codes = (e_thr <= 253 && (upd->norm_bits >> 52) != 0ull) ? 1 : 0;
    }
}

// ---- mask + bucket-gated distribution (sequences.py:648, core.py:125-155) ---------------------------------
// The strategy lives twice: in HBM (d_strat, the state the next update gates against) and in a host mirror
// that Contig.strat views (pinned memory of the library, or caller-provided shared memory handed over with
// bossgpu_set_strat_mirror). The mirror is kept current by the kernel itself: where the new bytes differ from
// the old ones the warp writes its 512 bytes through the mapped host pointer (posted PCIe writes). A
// steady-state update moves a few percent of the 62 MB a 3.1 Gb genome holds instead of copying all of it.
constexpr int DIST_THREADS = 256;
constexpr int DIST_VEC = 16;          // strategy bytes per thread and step; a warp owns 512 contiguous bytes

struct DistArgs {
    const SegDev* segs;
    const int64_t* srow_start;    // [n_seg+1]
    int n_seg, nb;
    const double2* benefit;       // [nb][n_rows], local merged rows
    const uint8_t* codes;         // [nb][n_rows][2] bin code of every benefit entry (k_hist); read instead of the benefits
                                  // when upd->use_codes: entry >= threshold  <=>  code <= upd->e_thr
    int64_t n_rows;
    int64_t R0, D0;
    const uint8_t* const* mask_ptrs; // multi-shard: [n_shards] packed mask of every shard, bit (i*2+s)*nb+b for the
                                  // shard-local merged row i (own HBM or a peer's exchange block);
                                  // NULL: single shard, benefit is compared directly
    const int64_t* shard_row_start;
    int n_shards;
    const uint8_t* bucket_sw;     // [n_sw][nb]
    uint8_t* strat;               // [n_srows][2][nb]; (strat - shift) is 16-byte aligned
    uint8_t* strat_host;          // device-visible alias of the host mirror, congruent to `strat` mod 16
    int shift;
    int64_t n_srows;
    UpdateDev* upd;
    unsigned long long* seg_accept; // [n_seg][2]
};

// A thread forms 16 consecutive bytes of the flattened [row][strand][barcode] strategy (8 rows without
// barcodes: eight 16-byte benefit loads), compares them with the old bytes in one 16-byte load and rewrites
// HBM where they differ. If any lane of the warp saw a change, the warp's 512 bytes are also written through
// the mapped host pointer, so the host mirror follows the device state chunk by chunk.
template <bool NB1>
__global__ void __launch_bounds__(DIST_THREADS)
k_distribute(DistArgs a) {
    __shared__ unsigned long long s_acc[2];
    __shared__ int s_sg;
    const int nb = NB1 ? 1 : a.nb;
    const int64_t total = a.n_srows * 2 * nb;
    if (!a.upd->switched_on || a.upd->empty) return;      // strategy left as it is (core.py:172; sequences.py:588 raises)
    const double thr = a.upd->threshold;
    const bool by_code = a.codes != nullptr && a.upd->use_codes != 0;
    const unsigned e_thr = (unsigned)a.upd->e_thr;
    const int64_t n_vec = (total + a.shift + DIST_VEC - 1) / DIST_VEC;    // virtual byte axis: v = i + shift
    // every CTA owns one contiguous run of vectors, so a thread stays inside one segment (contig) for long
    // stretches: segment geometry lives in registers and the per-segment accept counters see few atomics
    const int64_t per_cta = ((n_vec + gridDim.x - 1) / gridDim.x + DIST_THREADS - 1) / DIST_THREADS * DIST_THREADS;
    const int64_t v_begin = blockIdx.x * per_cta;
    const int64_t v_end = v_begin + per_cta;                             // whole warps stay in the loop (votes)
    // accepted entries per (segment, strand) for the per-contig log line (core.py:152-154)
    int cur_sg = -1;
    unsigned acc0 = 0, acc1 = 0;
    unsigned long long moved = 0;
    int sg = -1;
    int64_t row_lo = 0, row_hi = -1, sw_off = 0;                         // rows [row_lo, row_hi) belong to segment sg
    if (threadIdx.x == 0) { s_acc[0] = 0; s_acc[1] = 0; s_sg = -2; }
    __syncthreads();
    for (int64_t vi = v_begin + threadIdx.x; vi < v_end; vi += DIST_THREADS) {
        const int64_t i0 = vi * DIST_VEC - a.shift;
        const bool whole = i0 >= 0 && i0 + DIST_VEC <= total;
        const bool part = !whole && i0 + DIST_VEC > 0 && i0 < total;
        uint32_t w[4] = {0, 0, 0, 0};
        bool changed = false;
        bool fast = false;
        if (NB1 && whole && by_code) {
            // eight strategy rows of one barcode-less contig: sixteen code bytes against the threshold bin. With several shards
            // the codes of this shard serve every strategy row whose merged row lies here — all but the few rows at a shard
            // edge (Q2 shifts contig k by k rows), which read the gathered masks below
            const int64_t dl0 = i0 >> 1, dl1 = dl0 + 7;
            if (dl0 >= row_hi || dl0 < row_lo) {
                sg = find_segment(a.srow_start, a.n_seg, dl0);
                row_lo = a.srow_start[sg]; row_hi = a.srow_start[sg + 1]; sw_off = a.segs[sg].sw_off;
            }
            const int64_t j0 = dl0 - row_lo;                               // row within the segment
            const int64_t bk = j0 / (BUCKET / BIN);
            const int64_t m0 = a.D0 + dl0 - a.R0;                        // Q2: strategy row d reads merged row d
            const int64_t c0 = 2 * m0;
            const bool local = a.mask_ptrs == nullptr || (m0 >= 0 && m0 + 8 <= a.n_rows);
            if (local && dl1 < row_hi && ((reinterpret_cast<uintptr_t>(a.codes) + c0) & 3) == 0) {
                fast = true;
                const uint4 ov = *reinterpret_cast<const uint4*>(a.strat + i0);
                const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
                // the eight rows lie in one bucket or straddle two: the first n0 rows follow bucket bk, the rest bk + 1
                const int n0 = (int)min((int64_t)8, (bk + 1) * (BUCKET / BIN) - j0);
                const bool g0 = a.bucket_sw[(size_t)(sw_off + bk)] != 0;
                const bool g1 = n0 < 8 ? a.bucket_sw[(size_t)(sw_off + bk + 1)] != 0 : g0;
                if (g0 || g1) {
                    const uint32_t* cw = reinterpret_cast<const uint32_t*>(a.codes + c0);
                    const uint32_t et = e_thr * 0x01010101u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t gm = ((2 * q < n0 ? g0 : g1) ? 0x0000FFFFu : 0u) | ((2 * q + 1 < n0 ? g0 : g1) ? 0xFFFF0000u : 0u);
                        w[q] = (__vcmpleu4(cw[q], et) & 0x01010101u & gm) | (ow[q] & ~gm);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) w[q] = ow[q];
                }
                if (sg != cur_sg) {
                    if (acc0) atomicAdd(&a.seg_accept[2 * cur_sg], (unsigned long long)acc0);
                    if (acc1) atomicAdd(&a.seg_accept[2 * cur_sg + 1], (unsigned long long)acc1);
                    cur_sg = sg; acc0 = 0; acc1 = 0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { acc0 += __popc(w[q] & 0x00010001u); acc1 += __popc(w[q] & 0x01000100u); }
                changed = (w[0] != ow[0]) | (w[1] != ow[1]) | (w[2] != ow[2]) | (w[3] != ow[3]);
                if (changed) *reinterpret_cast<uint4*>(a.strat + i0) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        if (!fast && (whole || part)) {
            uint4 ov = make_uint4(0, 0, 0, 0);
            if (whole) ov = *reinterpret_cast<const uint4*>(a.strat + i0);
            const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
            int64_t row_cached = -1;
            double2 v = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < DIST_VEC; ++q) {
                const int64_t i = i0 + q;
                uint32_t cur = 0;
                if (whole || (i >= 0 && i < total)) {
                    int b, s;
                    int64_t dl;
                    if (NB1) { b = 0; s = (int)(i & 1); dl = i >> 1; }
                    else { b = (int)(i % nb); s = (int)((i / nb) & 1); dl = i / (2 * nb); }
                    if (dl >= row_hi || dl < row_lo) {
                        sg = find_segment(a.srow_start, a.n_seg, dl);
                        row_lo = a.srow_start[sg]; row_hi = a.srow_start[sg + 1]; sw_off = a.segs[sg].sw_off;
                    }
                    const uint32_t j = (uint32_t)(dl - row_lo);         // row within the segment
                    cur = whole ? (ow[q >> 2] >> (8 * (q & 3))) & 0xFFu : a.strat[i];
                    if (a.bucket_sw[(size_t)(sw_off + j / (BUCKET / BIN)) * nb + b]) {
                        const int64_t r = a.D0 + dl;                   // Q2: strategy row d reads merged row d
                        bool m;
                        if (a.mask_ptrs) {
                            const int sh = find_segment(a.shard_row_start, a.n_shards, r);
                            const int64_t bit = ((r - a.shard_row_start[sh]) * 2 + s) * nb + b;
                            m = (__ldcg(a.mask_ptrs[sh] + (bit >> 3)) >> (bit & 7)) & 1;
                        } else if (by_code) {
                            m = (unsigned)a.codes[((size_t)b * a.n_rows + (r - a.R0)) * 2 + s] <= e_thr;
                        } else {
                            if (!NB1 || dl != row_cached) { v = a.benefit[(size_t)b * a.n_rows + (r - a.R0)]; row_cached = dl; }
                            m = (s == 0 ? v.x : v.y) >= thr;
                        }
                        cur = m ? 1u : 0u;
                    }
                    if (sg != cur_sg) {
                        if (acc0) atomicAdd(&a.seg_accept[2 * cur_sg], (unsigned long long)acc0);
                        if (acc1) atomicAdd(&a.seg_accept[2 * cur_sg + 1], (unsigned long long)acc1);
                        cur_sg = sg; acc0 = 0; acc1 = 0;
                    }
                    if (s == 0) acc0 += cur; else acc1 += cur;
                    if (!whole && cur != a.strat[i]) { a.strat[i] = (uint8_t)cur; changed = true; }
                }
                w[q >> 2] |= cur << (8 * (q & 3));
            }
            if (whole) {
                changed = (w[0] != ow[0]) | (w[1] != ow[1]) | (w[2] != ow[2]) | (w[3] != ow[3]);
                if (changed) *reinterpret_cast<uint4*>(a.strat + i0) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        if (__any_sync(0xFFFFFFFFu, changed)) {
            // something in this warp's 512 bytes changed: refresh their image in the host mirror
            if (whole) {
                *reinterpret_cast<uint4*>(a.strat_host + i0) = make_uint4(w[0], w[1], w[2], w[3]);
                moved += DIST_VEC;
            } else if (part) {
                for (int q = 0; q < DIST_VEC; ++q) {
                    const int64_t i = i0 + q;
                    if (i >= 0 && i < total) { a.strat_host[i] = (uint8_t)((w[q >> 2] >> (8 * (q & 3))) & 0xFFu); ++moved; }
                }
            }
        }
    }
    // a CTA usually ends inside one segment: its counts go through shared memory and leave as two atomics
    if (threadIdx.x == 0) s_sg = cur_sg;
    __syncthreads();
    const bool with_cta = cur_sg == s_sg && cur_sg >= 0;
    if (with_cta) {
        const unsigned m = __match_any_sync(__activemask(), 1);
        const unsigned t0 = __reduce_add_sync(m, acc0), t1 = __reduce_add_sync(m, acc1);
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) {
            if (t0) atomicAdd(&s_acc[0], (unsigned long long)t0);
            if (t1) atomicAdd(&s_acc[1], (unsigned long long)t1);
        }
    } else if (cur_sg >= 0) {
        if (acc0) atomicAdd(&a.seg_accept[2 * cur_sg], (unsigned long long)acc0);
        if (acc1) atomicAdd(&a.seg_accept[2 * cur_sg + 1], (unsigned long long)acc1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_sg >= 0) {
        if (s_acc[0]) atomicAdd(&a.seg_accept[2 * s_sg], s_acc[0]);
        if (s_acc[1]) atomicAdd(&a.seg_accept[2 * s_sg + 1], s_acc[1]);
    }
    for (int o = 16; o > 0; o >>= 1) moved += __shfl_down_sync(0xFFFFFFFFu, moved, o);
    if ((threadIdx.x & 31) == 0 && moved) atomicAdd(&a.upd->mirror_bytes, moved);
}

// F-hat per window from the read-start counts kept on the device (readstartdist.py:96-115):
// (alpha + C) / denom where C > 0, the point-mass value elsewhere
__global__ void k_fhat_from_counts(int64_t n, const unsigned long long* __restrict__ counts, double alpha, double denom,
                                   double zero_value, double* __restrict__ fw) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long c = counts[i];
    fw[i] = c ? __ddiv_rn(__dadd_rn(alpha, (double)c), denom) : zero_value;
}

__global__ void k_count_read_starts(int64_t n, const int64_t* __restrict__ win, const uint8_t* __restrict__ strand,
                                    int64_t n_windows, unsigned long long* __restrict__ counts) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t w = win[i];
    if (w >= 0 && w < n_windows) atomicAdd(&counts[2 * w + (strand[i] ? 1 : 0)], 1ull);
}

// packed mask of this shard's merged rows: bit (i*2+s)*nb+b for local row i (rows cut by adjust_length are 0)
__global__ void k_pack_mask(const double2* __restrict__ benefit, const uint8_t* __restrict__ codes, int64_t n_rows, int nb, int64_t R0,
                            int64_t target, const UpdateDev* upd, uint8_t* __restrict__ out_bits, int64_t n_bits) {
    int64_t byte = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (byte * 8 >= n_bits || !upd->switched_on) return;
    const double thr = upd->threshold;
    const bool by_code = codes != nullptr && upd->use_codes != 0;
    const unsigned e_thr = (unsigned)upd->e_thr;
    unsigned v = 0;
    for (int k = 0; k < 8; ++k) {
        int64_t bit = byte * 8 + k;
        if (bit >= n_bits) break;
        int b = (int)(bit % nb);
        int s = (int)((bit / nb) & 1);
        int64_t i = bit / (2 * nb);
        if (R0 + i >= target) continue;
        if (by_code) {
            if ((unsigned)codes[((size_t)b * n_rows + i) * 2 + s] <= e_thr) v |= 1u << k;
            continue;
        }
        double2 x = benefit[(size_t)b * n_rows + i];
        if ((s == 0 ? x.x : x.y) >= thr) v |= 1u << k;
    }
    out_bits[byte] = (uint8_t)v;
}

__global__ void k_pack_strat(const uint8_t* __restrict__ strat, int64_t n, uint8_t* __restrict__ out) {
    int64_t byte = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (byte * 8 >= n) return;
    unsigned v = 0;
    for (int k = 0; k < 8; ++k) {
        int64_t i = byte * 8 + k;
        if (i < n && strat[i]) v |= 1u << k;
    }
    out[byte] = (uint8_t)v;
}

}  // namespace boss
