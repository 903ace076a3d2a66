// scatter.cuh — coverage update: CIGAR-run expansion + scatter-add into the per-site base counters.
//
// Replaces CoverageConverter._parse_cigar (boss/runs/sequences.py:744-794) and
// Contig.increment_coverage (boss/runs/reference.py:122-144).
//
// Counter layout in HBM: uint16 planes cov[barcode][base 0..4][padded site]; two neighbouring sites
// share one 32-bit word, so the increment is a 32-bit atomic on the containing word. Exact mod-2^16
// semantics of the reference's uint16 arrays (Q13) are kept by undoing the carry whenever a low half wraps.
//
// Work is balanced over CIGAR ops, not over reads:
//   k_tokenize      (text ingest only) CIGAR text -> packed ops, one CTA per read
//   k_op_prefix     one CTA per read: running reference / read offsets of every op -> one 16-byte record per op
//                   slot (a block scan per 2048 ops; a 60 kb read is three passes)
//   k_check_exc /   every read character outside ACGT that sits in an aligned column rejects the WHOLE batch
//   k_check_bases   (upstream: IndexError inside np.add.at, before `coverage += tmp_cov`, reference.py:138-144)
//                   -> the error flag is final before any counter is touched
//   k_scatter_ops   fixed blocks of 256 op slots, wherever read boundaries fall: the block's reference positions
//                   are numbered by a scan, every position finds its op through a mark + max-scan over a
//                   2048-position window (no per-position search), fetches its base and issues one atomic.
// The time of the coverage update is therefore (positions in the batch) / (atomic throughput) and shrinks with
// the shard, instead of being the walk of the longest read.
#pragma once
#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace boss {

// read bases packed on the host: 2 bits each (A C G T = 0..3), plus the rare other characters as a side list
struct PackedBases {
    const uint8_t* data;          // NULL: bases come as bytes (ASCII or codes)
    const int64_t* off;           // [n_reads] byte offset of each read's slice
    const int64_t* exc_off;       // [n_reads+1] or NULL when the batch holds nothing but ACGT
    const int32_t* exc_pos;       // position within the slice (sequencing orientation)
    const uint8_t* exc_char;      // the character itself; translated like upstream (ord - 48), usually an error
    const int32_t* exc_read;      // [n_exc] read of each entry (k_check_exc)
};

struct Span2 {
    int r, q;
    __host__ __device__ Span2 operator+(const Span2& o) const { return Span2{r + o.r, q + o.q}; }
};

// One op slot after k_op_prefix: everything k_scatter_ops needs to expand the op without looking at its read again.
// Reference position d of the op (0 <= d < len) is counted iff lo <= d < hi (the part inside this shard's segment);
// its counter is word index (base + code * P + d) on the plane axis; its read base sits at slice index q + d (forward)
// or q - d (reverse strand: the slice is walked backwards and complemented) of the slice starting at `boff`.
struct __align__(16) OpRec {
    long long base;       // (barcode * 5) * P + padded-axis site of the op's first reference position
    int32_t lo, hi;
    long long boff;       // where the read's aligned slice starts in the base array (bytes; packed: bytes of 4 bases)
    int32_t q;            // slice index (sequencing orientation) of the op's first read base
    uint32_t op;          // (len << 4) | (reverse ? 8 : 0) | class; 0 = dead slot
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

// ------------------------------------------------------------------------------------------------
// k_op_prefix: one CTA per read. cig_off has n_reads + 1 entries: read r owns the slots
// [cig_off[r], cig_off[r+1]) of which [cig_off[r], cig_end[r]) hold ops (the text ingest reserves
// text_len/2 + 1 slots per read; a pre-tokenised batch has no spare slots and cig_end = cig_off + 1).
// Writes one OpRec per slot plus the op's offset in the aligned slice (alignment orientation; the aligned-column
// checks search it), checks that the ops consume exactly the aligned slice (upstream: NumPy shape error at
// sequences.py:785) when `check_q` is set, and records every read's reference span.
// ------------------------------------------------------------------------------------------------
constexpr int PX_THREADS = 256;
constexpr int PX_OPS = 8;

struct PrefixArgs {
    int64_t n_reads;
    const int64_t* cig_off;
    const int64_t* cig_end;
    const uint32_t* cigar;
    const int32_t* seg_of;
    const int64_t* tstart;
    const int32_t* barcode;
    const int64_t* base_off;
    const uint8_t* rev;             // may be NULL (bases already in alignment orientation)
    const int64_t* pk_off;          // packed bases: byte offset of each read's slice; NULL: bases are bytes at base_off
    const SegDev* segs;
    int n_seg, nb;
    int64_t P;
    int check_q;
    OpRec* rec;
    int32_t* rec_q0;
    int32_t* rec_read;              // may be NULL (only the exception list of the text ingest needs it)
    int64_t* read_span;
    int32_t* err;
};

__global__ void __launch_bounds__(PX_THREADS)
k_op_prefix(PrefixArgs a) {
    using Scan = cub::BlockScan<Span2, PX_THREADS>;
    __shared__ typename Scan::TempStorage scan_tmp;
    for (int64_t read = blockIdx.x; read < a.n_reads; read += gridDim.x) {
        const int64_t c0 = a.cig_off[read], c1 = a.cig_end[read], cs = a.cig_off[read + 1];
        const int sg = a.seg_of[read];
        const bool tracked = sg >= 0 && sg < a.n_seg;
        int64_t s_start = 0, s_len = 0, s_off = 0;
        if (tracked) { s_start = a.segs[sg].start; s_len = a.segs[sg].len; s_off = a.segs[sg].site_off; }
        int b = a.barcode[read];
        if (b < 0 || b >= a.nb) b = 0;                                     // Q11: unknown / unclassified -> index 0
        const bool is_rev = a.rev != nullptr && a.rev[read] != 0;
        const int64_t q_first = a.base_off[read], n = a.base_off[read + 1] - q_first;
        const long long boff = a.pk_off ? a.pk_off[read] : q_first;
        const int64_t local_start = a.tstart[read] - s_start;              // segment-local site of the read's first position
        const long long plane0 = (long long)(b * 5) * a.P + s_off;
        Span2 carry{0, 0};
        for (int64_t cb = c0; cb < cs; cb += PX_THREADS * PX_OPS) {
            const int64_t i0 = cb + (int64_t)threadIdx.x * PX_OPS;
            uint32_t ops[PX_OPS];
            Span2 mine[PX_OPS], tsum{0, 0};
#pragma unroll
            for (int k = 0; k < PX_OPS; ++k) {
                const int64_t i = i0 + k;
                const uint32_t op = i < c1 ? a.cigar[i] : 0u;
                const int len = (int)(op >> 4), c = (int)(op & 15u);
                ops[k] = op;
                mine[k] = Span2{c != 1 ? len : 0, c != 2 ? len : 0};
                tsum = tsum + mine[k];
            }
            Span2 excl, total;
            Scan(scan_tmp).ExclusiveScan(tsum, excl, Span2{0, 0}, [](const Span2& x, const Span2& y) { return x + y; }, total);
            excl = excl + carry;
#pragma unroll
            for (int k = 0; k < PX_OPS; ++k) {
                const int64_t i = i0 + k;
                if (i < cs) {
                    const int len = (int)(ops[k] >> 4), c = (int)(ops[k] & 15u);
                    OpRec r;
                    r.base = 0; r.lo = 0; r.hi = 0; r.boff = boff; r.q = 0; r.op = 0u;
                    if (tracked && c != 1 && len > 0) {
                        const int64_t local0 = local_start + excl.r;        // segment-local site of the op's first position
                        r.lo = (int32_t)max((int64_t)0, -local0);
                        r.hi = (int32_t)max((int64_t)r.lo, min((int64_t)len, s_len - local0));
                        r.base = plane0 + local0;
                        // reverse-strand reads arrive in sequencing orientation: walk the slice backwards and complement
                        // (boss/utils.py:85-95: ATGC <-> TACG, everything else unchanged)
                        r.q = is_rev ? (int32_t)(n - 1 - excl.q) : excl.q;
                        r.op = ((uint32_t)len << 4) | (is_rev ? 8u : 0u) | (uint32_t)c;
                    }
                    a.rec[i] = r;
                    a.rec_q0[i] = i < c1 ? excl.q : 0x7FFFFFFF;
                    if (a.rec_read) a.rec_read[i] = (int32_t)read;
                }
                excl = excl + mine[k];
            }
            carry = carry + total;
            __syncthreads();                                  // scan storage is reused
        }
        if (threadIdx.x == 0) {
            a.read_span[read] = carry.r;
            if (a.check_q && (int64_t)carry.q != n) atomicExch(a.err, BOSSGPU_ESHAPE);
        }
    }
}

// the op of a read that consumes aligned-slice position k (alignment orientation): the last of its op slots [c0, c1)
// whose read offset is <= k. Returns its class if it really covers k, else -1.
__device__ __forceinline__ int op_class_at(const int32_t* __restrict__ q0, const uint32_t* __restrict__ cigar, int64_t c0, int64_t c1,
                                           int64_t k) {
    int64_t lo = c0, hi = c1;             // first slot with q0 > k
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)q0[mid] <= k) lo = mid + 1; else hi = mid;
    }
    if (lo == c0) return -1;
    const uint32_t op = cigar[lo - 1];
    const int cls = (int)(op & 15u);
    const int64_t qlen = cls != 2 ? (int64_t)(op >> 4) : 0;
    return (k >= q0[lo - 1] && k < q0[lo - 1] + qlen) ? cls : -1;
}

// text ingest: the host lists every character outside ACGT (read, position in the slice, character). One thread
// per entry: a character that upstream's translation sends beyond the five counters (ord - 48 > 4,
// sequences.py:762-763) and that an M-class op places in an aligned column raises IndexError upstream.
__global__ void k_check_exc(int64_t n_exc, PackedBases pk, const uint8_t* __restrict__ rev, const int64_t* __restrict__ base_off,
                            const int64_t* __restrict__ cig_off, const int64_t* __restrict__ cig_end,
                            const int32_t* __restrict__ rec_q0, const uint32_t* __restrict__ cigar, int32_t* __restrict__ err) {
    const int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (x >= n_exc) return;
    if (base_code_ascii(pk.exc_char[x]) <= 4u) return;        // '0'..'4': counted as base 0..4, like upstream
    const int64_t read = pk.exc_read[x];
    const int64_t n = base_off[read + 1] - base_off[read];
    const int64_t o = pk.exc_pos[x];
    const int64_t k = (rev != nullptr && rev[read]) ? n - 1 - o : o;
    if (op_class_at(rec_q0, cigar, cig_off[read], cig_end[read], k) == 0) atomicCAS(err, 0, BOSSGPU_EBASE);
}

// pre-tokenised ingest (bases as bytes in alignment orientation): the same check over every base of the batch
__global__ void k_check_bases(int64_t n_reads, const int64_t* __restrict__ base_off, const uint8_t* __restrict__ bases,
                              int base_is_ascii, const int64_t* __restrict__ cig_off, const int64_t* __restrict__ cig_end,
                              const int32_t* __restrict__ rec_q0, const uint32_t* __restrict__ cigar, int32_t* __restrict__ err) {
    const int64_t n_bases = base_off[n_reads];
    auto look = [&](int64_t x, unsigned ch) {
        const unsigned code = base_is_ascii ? base_code_ascii(ch) : ch;
        if (code <= 4u) return;
        int64_t lo = 0, hi = n_reads;                        // base_off[lo] <= x < base_off[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (base_off[mid] <= x) lo = mid; else hi = mid;
        }
        if (op_class_at(rec_q0, cigar, cig_off[lo], cig_end[lo], x - base_off[lo]) == 0) atomicCAS(err, 0, BOSSGPU_EBASE);
    };
    // sixteen bases per thread and step; the common case (nothing but ACGT / codes 0..4) is decided on whole words
    const int64_t head = min(n_bases, (int64_t)((16 - (reinterpret_cast<uintptr_t>(bases) & 15)) & 15));
    const int64_t n_vec = (n_bases - head) / 16;
    const uint4* vec = reinterpret_cast<const uint4*>(bases + head);
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
        const uint4 q = vec[v];
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
        bool clean = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (base_is_ascii) {
                // A C G T = 0x41 0x43 0x47 0x54: every byte must equal one of them
                const unsigned ok = __vcmpeq4(w[j], 0x41414141u) | __vcmpeq4(w[j], 0x43434343u) | __vcmpeq4(w[j], 0x47474747u) |
                                    __vcmpeq4(w[j], 0x54545454u);
                clean &= ok == 0xFFFFFFFFu;
            } else {
                clean &= __vcmpgtu4(w[j], 0x04040404u) == 0u;
            }
        }
        if (clean) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) look(head + 16 * v + j, (w[j >> 2] >> (8 * (j & 3))) & 0xFFu);
    }
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < head) look(g, bases[g]);
    const int64_t tail0 = head + 16 * n_vec;
    if (g < n_bases - tail0) look(tail0 + g, bases[tail0 + g]);
}

// depth totals for callers that let the library count (bossgpu_ingest_packed without contig_cov_add): every
// reference position of a read that lies inside its segment, added only if the batch was accepted
__global__ void k_add_read_totals(int64_t n_reads, const int32_t* __restrict__ seg_of, const int64_t* __restrict__ tstart,
                                  const int64_t* __restrict__ read_span, const SegDev* __restrict__ segs, int n_seg,
                                  unsigned long long* __restrict__ cov_total, const int32_t* __restrict__ err) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads || *err != 0) return;
    const int sg = seg_of[r];
    if (sg < 0 || sg >= n_seg) return;
    const int64_t s0 = segs[sg].start, s1 = s0 + segs[sg].len;
    const int64_t a = max(tstart[r], s0), b = min(tstart[r] + read_span[r], s1);
    if (b > a) atomicAdd(&cov_total[segs[sg].contig], (unsigned long long)(b - a));
}

// ------------------------------------------------------------------------------------------------
// k_scatter_ops: one CTA works through blocks of SC_THREADS consecutive op slots (thread i <-> slot i), wherever read
// boundaries fall. The block's reference positions are numbered 0..total by a scan over the ops' lengths; a window of
// SC_WIN positions at a time, every op marks its first position, a max-scan hands every position its op, and each
// thread expands SC_PER consecutive positions: base fetch, counter word, one atomic (two positions of one word merged).
// ------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_PER = 8;                         // consecutive reference positions per thread and window
constexpr int SC_WIN = SC_THREADS * SC_PER;       // 2048

struct ScatterArgs {
    const int64_t* n_slots;         // device: cig_off[n_reads]
    const OpRec* rec;
    const int32_t* rec_read;        // only read when pk.exc_off is set
    const uint8_t* bases;           // bytes (ASCII or codes) when pk.data is NULL
    int base_is_ascii;
    PackedBases pk;
    int64_t P;
    uint16_t* cov;
    int32_t* err;
};

// MODE 0: bases packed 2 bits each (text ingest), 1: ASCII bytes, 2: code bytes 0..4
template <int MODE>
__global__ void __launch_bounds__(SC_THREADS)
k_scatter_ops(ScatterArgs a) {
    using Scan = cub::BlockScan<int, SC_THREADS>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ long long s_base[SC_THREADS];      // counter index of block position 0 as this op sees it: base - pos0
    __shared__ int2 s_range[SC_THREADS];          // block positions of the op inside the segment: [pos0 + lo, pos0 + hi)
    __shared__ int2 s_q[SC_THREADS];              // .x: slice index of block position 0 (q -/+ pos0); .y: class | reverse << 3
    __shared__ long long s_boff[SC_THREADS];
    __shared__ __align__(16) short s_idx[SC_WIN];

    if (*reinterpret_cast<volatile int32_t*>(a.err) != 0) return;    // the batch was rejected before reaching the counters
    const int64_t n_slots = *a.n_slots;
    const int t = threadIdx.x;
    unsigned* cov32 = reinterpret_cast<unsigned*>(a.cov);
    for (int64_t blk = blockIdx.x; blk * SC_THREADS < n_slots; blk += gridDim.x) {
        const int64_t slot = blk * SC_THREADS + t;
        OpRec r;
        r.base = 0; r.lo = 0; r.hi = 0; r.boff = 0; r.q = 0; r.op = 0u;
        if (slot < n_slots) {
            const uint4* src = reinterpret_cast<const uint4*>(a.rec + slot);
            const uint4 w0 = src[0], w1 = src[1];
            r.base = (long long)(((unsigned long long)w0.y << 32) | w0.x); r.lo = (int)w0.z; r.hi = (int)w0.w;
            r.boff = (long long)(((unsigned long long)w1.y << 32) | w1.x); r.q = (int)w1.z; r.op = w1.w;
        }
        const int reflen = (int)(r.op >> 4);                         // dead slots, insertions and untracked reads carry 0
        int pos0, total;
        Scan(scan_tmp).ExclusiveSum(reflen, pos0, total);
        const bool is_rev = (r.op & 8u) != 0;
        s_base[t] = r.base - pos0;
        s_range[t] = make_int2(pos0 + r.lo, pos0 + r.hi);
        s_q[t] = make_int2(is_rev ? r.q + pos0 : r.q - pos0, (int)(r.op & 15u));
        s_boff[t] = r.boff;
        for (int win0 = 0; win0 < total; win0 += SC_WIN) {
            __syncthreads();                                  // previous window's readers are done; the op tables are visible
            *reinterpret_cast<uint4*>(&s_idx[t * SC_PER]) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            __syncthreads();
            if (reflen > 0 && pos0 + reflen > win0 && pos0 < win0 + SC_WIN) s_idx[max(pos0 - win0, 0)] = (short)t;
            __syncthreads();
            // op of every position = the last mark at or before it
            const uint4 raw = *reinterpret_cast<const uint4*>(&s_idx[t * SC_PER]);
            const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
            int loc[SC_PER], run = -1;
#pragma unroll
            for (int k = 0; k < SC_PER; ++k) {
                const int v = (int)(short)((w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
                run = max(run, v);
                loc[k] = run;
            }
            int before;
            Scan(scan_tmp).ExclusiveScan(run, before, -1, cub::Max());
            // a warp whose eight-position strips all lie behind the block's last position has nothing to expand
            if (win0 + (t & ~31) * SC_PER >= total) continue;
            // all of the thread's positions first (registers only, no data-dependent indexing), then its atomics
            unsigned long long widx[SC_PER];
            unsigned add[SC_PER];
#pragma unroll
            for (int k = 0; k < SC_PER; ++k) {
                const int p = win0 + t * SC_PER + k;
                const int i = max(max(before, loc[k]), 0);
                const int2 rg = s_range[i];
                const bool live = p >= rg.x && p < rg.y;      // also false beyond `total` and outside this shard's segment
                const int2 qi = s_q[i];
                const bool rev = (qi.y & 8) != 0;
                unsigned code = 4;                            // deletion column (sequences.py:792-793)
                if (live && (qi.y & 3) != 2) {
                    const int64_t o = rev ? (int64_t)qi.x - p : (int64_t)qi.x + p;    // slice index, sequencing orientation
                    if (MODE == 0) {
                        // 2 bits per base, 4 per byte, every read's slice starting on its own byte
                        unsigned ch = (a.pk.data[s_boff[i] + (o >> 2)] >> (2 * (o & 3))) & 3u;       // raw: A C T G = 0 1 2 3
                        ch ^= ch >> 1;                                                                // A C G T = 0 1 2 3
                        if (rev) ch = 3u - ch;
                        if (a.pk.exc_off) {                   // characters outside ACGT, listed apart
                            const int rd = a.rec_read[blk * SC_THREADS + i];
                            for (int64_t x = a.pk.exc_off[rd]; x < a.pk.exc_off[rd + 1]; ++x)
                                if (a.pk.exc_pos[x] == o) { ch = base_code_ascii(a.pk.exc_char[x]); break; }
                        }
                        code = ch;
                    } else {
                        const unsigned ch = a.bases[s_boff[i] + o];
                        if (MODE == 1) {
                            // branch-free: (ch >> 1) & 3 sends A C T G to 0 1 2 3, ^= >> 1 makes that A C G T; the reverse
                            // complement of a base is 3 - code; any other character is ord - 48, never complemented
                            const bool acgt = ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T';
                            unsigned c = (ch >> 1) & 3u;
                            c ^= c >> 1;
                            if (rev) c = 3u - c;
                            code = acgt ? c : ((ch - 48u) & 0xFFu);
                        } else {
                            code = (rev && ch < 4u) ? 3u - ch : ch;
                        }
                    }
                }
                const bool ok = live && code <= 4u;           // code > 4 is unreachable after the aligned-column checks
                const unsigned long long idx = (unsigned long long)(s_base[i] + p) + (unsigned long long)code * (unsigned long long)a.P;
                widx[k] = ok ? (idx >> 1) : ~0ull;
                add[k] = ok ? (1u << ((idx & 1ull) * 16)) : 0u;
            }
            // neighbouring positions often share a 32-bit word (same base at an even/odd site pair): one atomic
#pragma unroll
            for (int k = 1; k < SC_PER; ++k)
                if (widx[k] == widx[k - 1] && add[k]) { add[k] += add[k - 1]; add[k - 1] = 0u; }
            unsigned old[SC_PER];
#pragma unroll
            for (int k = 0; k < SC_PER; ++k)
                if (add[k]) old[k] = atomicAdd(cov32 + widx[k], add[k]);
#pragma unroll
            for (int k = 0; k < SC_PER; ++k)        // a carry out of the low counter must not leak into the high counter (Q13)
                if (add[k] && (add[k] & 0xFFFFu) && ((old[k] & 0xFFFFu) + (add[k] & 0xFFFFu) > 0xFFFFu)) atomicSub(cov32 + widx[k], 0x10000u);
        }
        __syncthreads();                                      // the next block overwrites the op tables
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

}  // namespace boss
