// Host side of the 2-bit wire format: ASCII bases -> the canonical packed words of
// include/dbga_b200.h (sequence j of a batch starts at word (offsets[j] >> 4) + j; base i in bits
// 2 (i & 15) of word i >> 4; A C G T = 0 1 2 3), with validation, 32 bases per AVX2 step (run-time
// dispatch; a scalar word-at-a-time form otherwise).  Used by dbga_pack_ascii and by
// dbga_align_batch itself, which packs every chunk of an ASCII batch on host threads while the
// previous chunks are in flight: a quarter of the bytes then cross the host link
// (reference input side: load_sequences, src/dbga/utils.py:95-96).
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define DBGA_X86 1
#endif

namespace dbga {

// 4 ASCII bases -> 8 bits of codes; accumulates the difference to the bases the codes stand for
static inline uint32_t pack4_scalar(uint32_t x, uint32_t &diff) {
    const uint32_t codes = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    const uint32_t lo = codes & 0x01010101u, hi = (codes >> 1) & 0x01010101u;
    const uint32_t expect = 0x41414141u + 2u * lo + 6u * hi + 11u * (lo & hi);      // A C G T = 65 67 71 84
    diff |= expect ^ x;
    return (codes * 0x01041040u) >> 24;
}

static bool pack_seq_scalar(const uint8_t *src, int64_t len, uint32_t *dst) {
    uint32_t diff = 0;
    int64_t w = 0;
    for (; 16 * (w + 1) <= len; ++w) {
        uint32_t word = 0;
        for (int j = 0; j < 4; ++j) {
            uint32_t x;
            memcpy(&x, src + 16 * w + 4 * j, 4);
            word |= pack4_scalar(x, diff) << (8 * j);
        }
        dst[w] = word;
    }
    if (16 * w < len) {
        uint32_t word = 0;
        for (int64_t i = 16 * w; i < len; ++i) {
            const uint8_t c = src[i];
            const uint32_t code = ((c >> 1) ^ (c >> 2)) & 3u;
            diff |= (uint32_t)("ACGT"[code] != (char)c);
            word |= code << (2 * (i - 16 * w));
        }
        dst[w] = word;
    }
    return diff != 0;
}

#ifdef DBGA_X86
__attribute__((target("avx2"))) static bool pack_seq_avx2(const uint8_t *src, int64_t len, uint32_t *dst) {
    const __m256i m3 = _mm256_set1_epi8(3);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                         'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w14 = _mm256_set1_epi16(0x0401);          // bytes (1, 4): c0 + 4 c1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);     // words (1, 16): (c0 + 4 c1) + 16 (c2 + 4 c3)
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    __m256i bad = _mm256_setzero_si256();
    int64_t i = 0, w = 0;
    for (; i + 32 <= len; i += 32, w += 2) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i c = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), m3);
        bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_shuffle_epi8(lut, c), x));
        const __m256i t = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w14), w116);   // one byte of codes per 4 bases
        const __m256i g = _mm256_shuffle_epi8(t, gather);                           // 4 bytes per 128-bit lane
        dst[w] = (uint32_t)_mm256_extract_epi32(g, 0);
        dst[w + 1] = (uint32_t)_mm256_extract_epi32(g, 4);
    }
    bool b = !_mm256_testz_si256(bad, bad);
    if (i < len) b |= pack_seq_scalar(src + i, len - i, dst + w);
    return b;
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2");
#endif

// ---- FASTA record bodies (fasta_io.cu): the bases of [p, end) -- everything but '\n', '\r', ' ', '\t' --
// counted, or copied upper-cased to dst (which holds exactly `room` bytes: never written past, the records of a
// file are filled by different threads side by side).  32 bytes per step: a step without a blank is one load, four
// compares and one store; a blank ends the step behind it.
static inline bool is_blank(uint8_t c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }
static int64_t body_scalar(const uint8_t *p, const uint8_t *end, uint8_t *dst) {
    int64_t n = 0;
    for (; p < end; ++p)
        if (!is_blank(*p)) { if (dst) dst[n] = (uint8_t)(*p - (((uint8_t)(*p - 'a') < 26u) << 5)); ++n; }
    return n;
}
#ifdef DBGA_X86
__attribute__((target("avx2,bmi,popcnt"))) static int64_t body_avx2(const uint8_t *p, const uint8_t *end, uint8_t *dst, int64_t room) {
    const __m256i nl = _mm256_set1_epi8('\n'), cr = _mm256_set1_epi8('\r'), sp = _mm256_set1_epi8(' '), tb = _mm256_set1_epi8('\t');
    const __m256i a1 = _mm256_set1_epi8('a' - 1), z1 = _mm256_set1_epi8('z' + 1), c20 = _mm256_set1_epi8(0x20);
    int64_t n = 0;
    if (!dst) {
        for (; p + 32 <= end; p += 32) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
            const __m256i b = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, nl), _mm256_cmpeq_epi8(v, cr)),
                                              _mm256_or_si256(_mm256_cmpeq_epi8(v, sp), _mm256_cmpeq_epi8(v, tb)));
            n += 32 - __builtin_popcount((unsigned)_mm256_movemask_epi8(b));
        }
        return n + body_scalar(p, end, nullptr);
    }
    while (p + 32 <= end && n + 32 <= room) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
        const __m256i b = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, nl), _mm256_cmpeq_epi8(v, cr)),
                                          _mm256_or_si256(_mm256_cmpeq_epi8(v, sp), _mm256_cmpeq_epi8(v, tb)));
        const unsigned m = (unsigned)_mm256_movemask_epi8(b);
        const __m256i low = _mm256_and_si256(_mm256_cmpgt_epi8(v, a1), _mm256_cmpgt_epi8(z1, v));
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + n), _mm256_sub_epi8(v, _mm256_and_si256(low, c20)));
        if (m == 0) { p += 32; n += 32; }
        else { const int k = __builtin_ctz(m); p += k + 1; n += k; }        // the bytes behind the blank are rewritten by the next step
    }
    return n + body_scalar(p, end, dst + n);
}
#endif
int64_t host_fasta_body(const uint8_t *p, const uint8_t *end, uint8_t *dst, int64_t room) {
#ifdef DBGA_X86
    if (g_have_avx2) return body_avx2(p, end, dst, room);
#endif
    (void)room;
    return body_scalar(p, end, dst);
}

// one sequence; returns true when it holds a character outside ACGT (packed as A)
bool host_pack_seq(const uint8_t *src, int64_t len, uint32_t *dst) {
#ifdef DBGA_X86
    if (g_have_avx2) return pack_seq_avx2(src, len, dst);
#endif
    return pack_seq_scalar(src, len, dst);
}

}  // namespace dbga
