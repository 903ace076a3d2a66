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

// read bases packed on the host: 2 bits each (A C G T = 0..3), plus the rare other characters as a side list
struct PackedBases {
    const uint8_t* data;          // NULL: bases come as bytes (ASCII or codes)
    const int64_t* off;           // [n_reads] byte offset of each read's slice
    const int64_t* exc_off;       // [n_reads+1] or NULL when the batch holds nothing but ACGT
    const int32_t* exc_pos;       // position within the slice (sequencing orientation)
    const uint8_t* exc_char;      // the character itself; translated like upstream (ord - 48), usually an error
};

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
          const uint8_t* __restrict__ bases, const uint8_t* __restrict__ rev, int base_is_ascii, PackedBases pk,
          const SegDev* __restrict__ segs, int n_seg,
          int nb, int64_t P, uint16_t* __restrict__ cov, unsigned long long* __restrict__ cov_total,
          int count_totals, int32_t* __restrict__ err) {
    using Scan = cub::BlockScan<Span2, SC_THREADS>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ int s_r[SC_CHUNK + 1];
    __shared__ int s_q[SC_CHUNK];
    __shared__ unsigned char s_cls[SC_CHUNK];

    if (*reinterpret_cast<volatile int32_t*>(err) != 0) return;      // the batch was rejected before reaching the counters
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
                        unsigned ch = 0xFFu;
                        if (pk.data) {
                            // 2 bits per base, 4 per byte, every read's slice starting on its own byte
                            const int64_t o = qi - q0;                          // index in sequencing orientation
                            if (o >= 0 && o < q1 - q0) {
                                ch = (pk.data[pk.off[read] + (o >> 2)] >> (2 * (o & 3))) & 3u;      // raw: A C T G = 0 1 2 3
                                ch ^= ch >> 1;                                                       // A C G T = 0 1 2 3
                                if (is_rev) ch = 3u - ch;
                                if (pk.exc_off) {                               // characters outside ACGT, listed apart
                                    for (int64_t x = pk.exc_off[read]; x < pk.exc_off[read + 1]; ++x)
                                        if (pk.exc_pos[x] == o) { ch = base_code_ascii(pk.exc_char[x]); break; }
                                }
                            }
                            code = ch;
                        } else {
                        if (qi >= q0 && qi < q1) ch = bases[qi];
                        if (base_is_ascii) {
                            if (is_rev) ch = ch == 'A' ? 'T' : ch == 'T' ? 'A' : ch == 'G' ? 'C' : ch == 'C' ? 'G' : ch;
                            code = base_code_ascii(ch);
                        } else {
                            code = (is_rev && ch < 4u) ? 3u - ch : ch;
                        }
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

// ------------------------------------------------------------------------------------------------
// CIGAR text -> packed ops, on the device. Replaces the regex of CoverageConverter._parse_cigar
// (boss/runs/sequences.py:672,768-776): `(\d+)([MIDNSHP=XB])` applied with findall, i.e. an op is a maximal digit
// run immediately followed by one of those letters; everything else is skipped. One CTA per read: every thread
// looks at 8 characters, an op letter preceded by a digit is an op, a block scan numbers them, and the thread that
// owns the letter walks back over its digits. Ops go to the read's slot (>= 2 characters per op, so
// text_len/2 + 1 entries always suffice); spans are checked against the alignment's interval and the read slice
// like upstream's asserts (sequences.py:732-733, :785) before any counter is touched.
// ------------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 128;
constexpr int TK_PER_THREAD = 8;

__device__ __forceinline__ int cigar_class(unsigned ch) {
    switch (ch) {
        case 'I': return 1;
        case 'D': return 2;
        case 'M': case 'N': case 'S': case 'H': case 'P': case '=': case 'X': case 'B': return 0;
        default: return -1;
    }
}

__global__ void __launch_bounds__(TK_THREADS)
k_tokenize(int64_t n_reads, const char* __restrict__ text, const int64_t* __restrict__ text_off,
           const int64_t* __restrict__ op_off, const int64_t* __restrict__ base_off, const int64_t* __restrict__ tspan,
           uint32_t* __restrict__ ops, int64_t* __restrict__ op_end, int32_t* __restrict__ err) {
    using Scan = cub::BlockScan<int, TK_THREADS>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ long long s_r, s_q;
    for (int64_t read = blockIdx.x; read < n_reads; read += gridDim.x) {
        const char* s = text + text_off[read];
        const int64_t n = text_off[read + 1] - text_off[read];
        uint32_t* out = ops + op_off[read];
        if (threadIdx.x == 0) { s_r = 0; s_q = 0; }
        __syncthreads();
        int64_t n_out = 0;
        long long r = 0, q = 0;
        for (int64_t base = 0; base < n; base += TK_THREADS * TK_PER_THREAD) {
            const int64_t i0 = base + (int64_t)threadIdx.x * TK_PER_THREAD;
            int cls[TK_PER_THREAD], mine = 0;
#pragma unroll
            for (int k = 0; k < TK_PER_THREAD; ++k) {
                const int64_t i = i0 + k;
                cls[k] = -1;
                if (i > 0 && i < n) {
                    const unsigned prev = (unsigned char)s[i - 1];
                    const int c = cigar_class((unsigned char)s[i]);
                    if (c >= 0 && prev >= '0' && prev <= '9') { cls[k] = c; ++mine; }
                }
            }
            int excl, total;
            Scan(scan_tmp).ExclusiveSum(mine, excl, total);
#pragma unroll
            for (int k = 0; k < TK_PER_THREAD; ++k) {
                if (cls[k] < 0) continue;
                const int64_t i = i0 + k;
                int64_t d = i - 1;
                while (d > 0 && s[d - 1] >= '0' && s[d - 1] <= '9') --d;      // start of the digit run
                unsigned long long num = 0;
                for (; d < i; ++d) num = num * 10ull + (unsigned long long)(s[d] - '0');
                const uint32_t len = (uint32_t)(num & 0x0FFFFFFFull);          // as the host tokenizer (tokenizer.h)
                out[n_out + excl] = (len << 4) | (uint32_t)cls[k];
                ++excl;
                if (cls[k] != 1) r += len;
                if (cls[k] != 2) q += len;
            }
            n_out += total;
            __syncthreads();                                                     // scan storage is reused
        }
        for (int o = 16; o > 0; o >>= 1) { r += __shfl_down_sync(0xFFFFFFFFu, r, o); q += __shfl_down_sync(0xFFFFFFFFu, q, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd((unsigned long long*)&s_r, (unsigned long long)r); atomicAdd((unsigned long long*)&s_q, (unsigned long long)q); }
        __syncthreads();
        if (threadIdx.x == 0) {
            op_end[read] = op_off[read] + n_out;
            if (s_q != base_off[read + 1] - base_off[read] || s_r != tspan[read]) atomicExch(err, BOSSGPU_ESHAPE);
        }
        __syncthreads();
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
