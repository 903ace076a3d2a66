// tokenizer.h — host-side CIGAR tokenizer and reverse complement (plain C++; no CUDA).
//
// Replaces the regex/translate pipeline of CoverageConverter._parse_cigar
// (boss/runs/sequences.py:672,768-776) and boss/utils.py:85-95 (reverse_complement).
#pragma once
#include <stdint.h>

namespace boss {

// Upstream's regex `(\d+)([MIDNSHP=XB])` is applied with findall: anything that does not match is
// skipped silently. Op classes as consumed by the scatter kernel: 1 = I (read only), 2 = D (reference
// only), 0 = every other letter (upstream fills those columns from the read and keeps them).
// Returns the number of ops written (counted only when out == nullptr), or -1 if `cap` is too small.
inline int64_t tokenize_cigar(const char* s, int64_t n, uint32_t* out, int64_t cap, int64_t* ref_span, int64_t* query_span) {
    int64_t k = 0, r = 0, q = 0;
    uint64_t num = 0;
    bool have = false;
    for (int64_t i = 0; i < n; ++i) {
        const unsigned char ch = (unsigned char)s[i];
        if (ch >= '0' && ch <= '9') {
            num = num * 10 + (ch - '0');
            have = true;
            continue;
        }
        int cls = -1;
        switch (ch) {
            case 'I': cls = 1; break;
            case 'D': cls = 2; break;
            case 'M': case 'N': case 'S': case 'H': case 'P': case '=': case 'X': case 'B': cls = 0; break;
            default: break;
        }
        if (have && cls >= 0) {
            // lengths are uint32 upstream (np.array(lengths, dtype=np.uint32)); 28 bits is > any real run
            uint32_t len = (uint32_t)(num & 0x0FFFFFFFu);
            if (out) {
                if (k >= cap) return -1;
                out[k] = (len << 4) | (uint32_t)cls;
            }
            ++k;
            if (cls != 1) r += len;
            if (cls != 2) q += len;
        }
        num = 0;
        have = false;
    }
    *ref_span = r;
    *query_span = q;
    return k;
}

// reverse complement of the aligned slice: upstream complements ATGC (upper case only) and leaves every
// other character as is, then reverses (boss/utils.py:92-94)
inline void revcomp_copy(const char* src, int64_t n, char* dst) {
    for (int64_t i = 0; i < n; ++i) {
        char c = src[n - 1 - i];
        switch (c) {
            case 'A': c = 'T'; break;
            case 'T': c = 'A'; break;
            case 'G': c = 'C'; break;
            case 'C': c = 'G'; break;
            default: break;
        }
        dst[i] = c;
    }
}

}  // namespace boss
