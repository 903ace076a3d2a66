// scatter.cuh — coverage update: CIGAR-run expansion + scatter-add into the per-site base counters.
//
// Replaces CoverageConverter._parse_cigar (boss/runs/sequences.py:744-794) and
// Contig.increment_coverage (boss/runs/reference.py:122-144).
//
// Counter layout in HBM: uint16 planes cov[barcode][base 0..4][padded site]; two neighbouring sites
// share one 32-bit word, so the increment is a 32-bit atomic on the containing word. Lanes of a warp
// walk consecutive reference positions of one read, so two lanes that hit the two halves of the
// same word (same base at an even/odd site pair) are merged into a single atomic (warp-aggregated
// through one shuffle). Exact mod-2^16 semantics of the reference's uint16 arrays (Q13) are kept by
// undoing the carry whenever a low half wraps.
#pragma once
#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace boss {

constexpr int SC_THREADS = 128;
constexpr int SC_OPS_PER_THREAD = 4;
constexpr int SC_CHUNK = SC_THREADS * SC_OPS_PER_THREAD;   // CIGAR ops staged per pass

struct Span2 {
    int r, q;
    __host__ __device__ Span2 operator+(const Span2& o) const { return Span2{r + o.r, q + o.q}; }
};

// ASCII -> code as upstream: ACGT -> 0..3, everything else -> ord - 48 (mod 256) (sequences.py:666,762-763)
__device__ __forceinline__ unsigned base_code_ascii(unsigned ch) {
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return (ch - 48u) & 0xFFu;
    }
}

__device__ __forceinline__ void add_u16_pair(unsigned* word, unsigned add) {
    unsigned old = atomicAdd(word, add);
    // a carry out of the low counter must not leak into the high counter
    if ((add & 0xFFFFu) && ((old & 0xFFFFu) + (add & 0xFFFFu) > 0xFFFFu)) atomicSub(word, 0x10000u);
}

// one CTA per read
__global__ void __launch_bounds__(SC_THREADS)
k_scatter(int64_t n_reads, const int32_t* __restrict__ seg_of, const int64_t* __restrict__ tstart,
          const int32_t* __restrict__ barcode, const int64_t* __restrict__ cig_off, const int64_t* __restrict__ cig_end,
          const uint32_t* __restrict__ cigar, const int64_t* __restrict__ base_off,
          const uint8_t* __restrict__ bases, const uint8_t* __restrict__ rev, int base_is_ascii,
          const SegDev* __restrict__ segs, int n_seg,
          int nb, int64_t P, uint16_t* __restrict__ cov, unsigned long long* __restrict__ cov_total,
          int count_totals, int32_t* __restrict__ err) {
    using Scan = cub::BlockScan<Span2, SC_THREADS>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ int s_r[SC_CHUNK + 1];
    __shared__ int s_q[SC_CHUNK];
    __shared__ unsigned char s_cls[SC_CHUNK];

    for (int64_t read = blockIdx.x; read < n_reads; read += gridDim.x) {
        int sg = seg_of[read];
        if (sg < 0 || sg >= n_seg) continue;
        const SegDev S = segs[sg];
        int b = barcode[read];
        if (b < 0 || b >= nb) b = 0;                     // Q11: unknown / unclassified -> index 0
        const int64_t c0 = cig_off[read], c1 = cig_end[read];
        const int64_t q0 = base_off[read], q1 = base_off[read + 1];
        // reverse-strand reads arrive in sequencing orientation: walk the slice backwards and complement
        // (boss/utils.py:85-95: ATGC <-> TACG, everything else unchanged)
        const bool is_rev = rev != nullptr && rev[read] != 0;
        const int64_t t0 = tstart[read];
        int64_t ref_done = 0, q_done = 0;
        unsigned long long in_seg = 0;

        for (int64_t cb = c0; cb < c1; cb += SC_CHUNK) {
            // ---- stage one chunk of ops and scan their reference / read spans -------------------
            Span2 mine[SC_OPS_PER_THREAD];
            unsigned char cls[SC_OPS_PER_THREAD];
            Span2 tsum{0, 0};
            for (int k = 0; k < SC_OPS_PER_THREAD; ++k) {
                int64_t i = cb + threadIdx.x * SC_OPS_PER_THREAD + k;
                unsigned op = i < c1 ? cigar[i] : 1u;    // padding: zero-length insertion
                int len = (int)(op >> 4);
                int c = (int)(op & 15u);
                cls[k] = (unsigned char)c;
                mine[k] = Span2{c != 1 ? len : 0, c != 2 ? len : 0};
                tsum = tsum + mine[k];
            }
            Span2 excl, total;
            Scan(scan_tmp).ExclusiveScan(tsum, excl, Span2{0, 0}, [](const Span2& a, const Span2& b) { return a + b; }, total);
            for (int k = 0; k < SC_OPS_PER_THREAD; ++k) {
                int j = threadIdx.x * SC_OPS_PER_THREAD + k;
                s_r[j] = excl.r; s_q[j] = excl.q; s_cls[j] = cls[k];
                excl = excl + mine[k];
            }
            if (threadIdx.x == SC_THREADS - 1) s_r[SC_CHUNK] = total.r;
            __syncthreads();

            // ---- expand: one lane per reference position of the chunk ---------------------------
            const int R = total.r;
            for (int pbase = 0; pbase < R; pbase += SC_THREADS) {
                int p = pbase + threadIdx.x;
                bool live = p < R;
                unsigned long long widx = ~0ull;     // index of the 32-bit word to touch
                unsigned add = 0;
                if (live) {
                    // last op whose reference start <= p (zero-span insertions sort before it)
                    int lo = 0, hi = SC_CHUNK;       // invariant: s_r[lo] <= p < s_r[hi]
                    while (hi - lo > 1) {
                        int mid = (lo + hi) >> 1;
                        if (s_r[mid] <= p) lo = mid; else hi = mid;
                    }
                    unsigned code;
                    if (s_cls[lo] == 2) {
                        code = 4;                    // deletion column (sequences.py:792-793)
                    } else {
                        const int64_t k = q_done + s_q[lo] + (p - s_r[lo]);     // index in alignment orientation
                        const int64_t qi = is_rev ? q1 - 1 - k : q0 + k;
                        unsigned ch = (qi >= q0 && qi < q1) ? bases[qi] : 0xFFu;
                        if (base_is_ascii) {
                            if (is_rev) ch = ch == 'A' ? 'T' : ch == 'T' ? 'A' : ch == 'G' ? 'C' : ch == 'C' ? 'G' : ch;
                            code = base_code_ascii(ch);
                        } else {
                            code = (is_rev && ch < 4u) ? 3u - ch : ch;
                        }
                    }
                    int64_t site = t0 + ref_done + p - S.start;          // segment-local
                    if (code > 4) {
                        atomicExch(err, BOSSGPU_EBASE);
                        live = false;
                    } else if (site >= 0 && site < S.len) {
                        unsigned long long idx = ((unsigned long long)(b * 5 + (int)code)) * (unsigned long long)P +
                                                 (unsigned long long)(S.site_off + site);
                        widx = idx >> 1;
                        add = 1u << ((idx & 1ull) * 16);
                        in_seg++;
                    } else {
                        live = false;
                    }
                }
                // merge the two halves of one word when neighbouring lanes hit them
                unsigned long long up = __shfl_down_sync(0xFFFFFFFFu, widx, 1);
                unsigned long long dn = __shfl_up_sync(0xFFFFFFFFu, widx, 1);
                int lane = threadIdx.x & 31;
                bool low_half = live && add == 1u;
                bool absorbed = live && add == 0x10000u && lane > 0 && dn == widx;   // lower lane owns the low half
                if (low_half && lane < 31 && up == widx) add = 0x10001u;
                if (live && !absorbed) add_u16_pair(reinterpret_cast<unsigned*>(cov) + widx, add);
            }
            ref_done += total.r;
            q_done += total.q;
            __syncthreads();
        }
        if (count_totals) {
            // block-reduce the number of in-segment positions and add it to the contig's depth total
            __shared__ unsigned long long s_cnt;
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            unsigned long long v = in_seg;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt, v);
            __syncthreads();
            if (threadIdx.x == 0 && s_cnt) atomicAdd(&cov_total[S.contig], s_cnt);
            __syncthreads();
        }
    }
}

// span check of a tokenised batch: ref span must equal tend-tstart is checked on the host; this
// kernel verifies that the read slice is exactly consumed (upstream: NumPy shape error at
// sequences.py:785 when len(int_seq[start:end]) != number of non-deletion columns)
__global__ void k_check_spans(int64_t n_reads, const int64_t* __restrict__ cig_off, const int64_t* __restrict__ cig_end,
                              const uint32_t* __restrict__ cigar, const int64_t* __restrict__ base_off, int32_t* __restrict__ err) {
    int64_t read = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (read >= n_reads) return;
    int64_t q = 0;
    for (int64_t i = cig_off[read]; i < cig_end[read]; ++i) {
        unsigned op = cigar[i];
        if ((op & 15u) != 2u) q += op >> 4;
    }
    if (q != base_off[read + 1] - base_off[read]) atomicExch(err, BOSSGPU_ESHAPE);
}

}  // namespace boss
