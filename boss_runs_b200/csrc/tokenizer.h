// tokenizer.h — host-side CIGAR tokenizer and reverse complement (plain C++; no CUDA).
//
// Replaces the regex/translate pipeline of CoverageConverter._parse_cigar
// (boss/runs/sequences.py:672,768-776) and boss/utils.py:85-95 (reverse_complement).
#pragma once
#include <stdint.h>
#include <string.h>

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

// ---- read bases, 2 bits each -----------------------------------------------------------------------------
// (c >> 1) & 3 sends A C T G to 0 1 2 3 ("raw" codes; the scatter kernel turns them into A C G T = 0 1 2 3 with
// raw ^ (raw >> 1)). Four bases per byte, base k of a read in bits 2(k&3).. of byte k>>2. Characters outside
// ACGT are reported through `bad(pos, ch)`; their two bits are meaningless.
template <typename Bad>
inline void pack_bases_scalar(const unsigned char* src, int64_t from, int64_t n, uint8_t* dst, Bad&& bad) {
    for (int64_t k = from; k < n; k += 4) {
        unsigned v = 0;
        for (int q = 0; q < 4 && k + q < n; ++q) {
            const unsigned c = src[k + q];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') bad(k + q, (unsigned char)c);
            v |= ((c >> 1) & 3u) << (2 * q);
        }
        dst[k >> 2] = (uint8_t)v;
    }
}

#if defined(__x86_64__) && defined(__GNUC__)
#define BOSS_HAVE_AVX2_PACK 1
}  // namespace boss
#include <immintrin.h>
namespace boss {
// 32 characters -> 8 bytes per step
template <typename Bad>
__attribute__((target("avx2"))) inline int64_t pack_bases_avx2(const unsigned char* src, int64_t n, uint8_t* dst, Bad&& bad) {
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);            // bytes (1, 4): t0 + 4*t1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);       // words (1, 16): p0 + 16*p1
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    int64_t k = 0;
    for (; k + 32 <= n; k += 32) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(src + k));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, cA), _mm256_cmpeq_epi8(x, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(x, cG), _mm256_cmpeq_epi8(x, cT)));
        const unsigned okm = (unsigned)_mm256_movemask_epi8(ok);
        if (okm != 0xFFFFFFFFu)
            for (int q = 0; q < 32; ++q)
                if (!((okm >> q) & 1u)) bad(k + q, src[k + q]);
        const __m256i t = _mm256_and_si256(_mm256_srli_epi16(x, 1), three);
        const __m256i p = _mm256_maddubs_epi16(t, w14);
        const __m256i q4 = _mm256_madd_epi16(p, w116);
        const __m256i r = _mm256_shuffle_epi8(q4, pick);
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(r, 0), hi = (uint32_t)_mm256_extract_epi32(r, 4);
        memcpy(dst + (k >> 2), &lo, 4);
        memcpy(dst + (k >> 2) + 4, &hi, 4);
    }
    return k;       // characters done (multiple of 32)
}
#endif

template <typename Bad>
inline void pack_bases(const unsigned char* src, int64_t n, uint8_t* dst, Bad&& bad) {
    int64_t done = 0;
#ifdef BOSS_HAVE_AVX2_PACK
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) done = pack_bases_avx2(src, n, dst, bad);
#endif
    pack_bases_scalar(src, done, n, dst, bad);
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
