// ck_host_pack.cpp -- the host side of CK_F_PACKED_IN: a multi-threaded packer that turns record bytes into what
// ck_canon_submit_packed / ck_uniq_submit_packed ship over PCIe (2 bits per base instead of 8).
//
// It does, per record, what k_prepare does on the device (ck_kernels.cuh) and what the reference's worker closure does
// before circkit::canonicalize (src/canonicalize.rs:24-27, src/uniq.rs:35-38):
//   flags & CK_F_NORMALIZE  needletail::sequence::normalize(seq, false): white space dropped, acgt -> upper, t/u/U -> T,
//                           . ~ -> -, anything else -> N;   otherwise bytes are taken as they are (library semantics);
//   lane[i] = 2  every symbol is A/C/G/T: the record goes into the dense 2-bit layout (16 bases per 32-bit unit, first base in
//                the top bits, units in address order, record i at 64-bit word (offsets[i] >> 5) + i);
//   lane[i] = 4  symbols of the 16-letter alphabet -ABCDGHKMNRSTVWY (the CLI alphabet {-,A,C,G,N,T} and upper-case IUPAC);
//   lane[i] = 8  anything else;  lanes 4 and 8 travel as (normalised) bytes in `lane_bytes`, concatenated in record order.
// Built by g++ (no CUDA in here) with an AVX2 + BMI2 fast path chosen at run time; the scalar path gives the same bytes.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/circkit_b200.h"

namespace {

struct Tables {
    uint8_t norm[256];    // needletail normalize(_, false): 0 = dropped
    uint8_t cls[256];     // of a (normalised) symbol: 0 = ACGT, 1 = other member of the 16-letter alphabet, 2 = anything else
    Tables()
    {
        for (int c = 0; c < 256; c++) norm[c] = 'N';
        for (const char *p = "ACGTN-"; *p; p++) norm[(uint8_t)*p] = (uint8_t)*p;
        norm['a'] = 'A'; norm['c'] = 'C'; norm['g'] = 'G';
        norm['t'] = 'T'; norm['u'] = 'T'; norm['U'] = 'T';
        norm['.'] = '-'; norm['~'] = '-';
        norm[' '] = 0; norm['\t'] = 0; norm['\r'] = 0; norm['\n'] = 0;
        for (int c = 0; c < 256; c++) cls[c] = 2;
        for (const char *p = "-ABCDGHKMNRSTVWY"; *p; p++) cls[(uint8_t)*p] = 1;
        for (const char *p = "ACGT"; *p; p++) cls[(uint8_t)*p] = 0;
    }
};
const Tables kT;

inline uint32_t code2(uint32_t b) { return ((b >> 1) & 3u) ^ ((b >> 2) & 1u); }   // A,C,G,T (any case), U -> 0..3

// ---- scalar packer of `n` A/C/G/T bytes into 32-bit units (first base in the top bits; the last unit zero-padded)
void pack_scalar(const uint8_t *s, uint32_t n, uint32_t *dst)
{
    uint32_t j = 0;
    for (; 16u * (j + 1) <= n; j++) {
        uint32_t u = 0;
        for (int k = 0; k < 16; k++) u = (u << 2) | code2(s[16 * j + k]);
        dst[j] = u;
    }
    const uint32_t r = n - 16 * j;
    if (r) {
        uint32_t u = 0;
        for (uint32_t k = 0; k < r; k++) u = (u << 2) | code2(s[16 * j + k]);
        dst[j] = u << (2 * (16 - r));
    }
}

// true iff every byte is A/C/G/T (strict) or, with `fold`, one of ACGTacgtUu (what normalisation maps to A/C/G/T without
// dropping anything); packs on the way.  Returns false as soon as another byte shows up (dst then holds garbage).
bool pack_pure_scalar(const uint8_t *s, uint32_t n, uint32_t *dst, bool fold)
{
    uint32_t bad = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t b = s[i];
        const uint8_t m = fold ? kT.norm[b] : b;
        bad |= (m == 0) | kT.cls[m];
    }
    if (bad) return false;
    pack_scalar(s, n, dst);
    return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) inline uint32_t pack16_bmi2(const uint8_t *s)
{
    uint64_t a, b;
    memcpy(&a, s, 8); memcpy(&b, s + 8, 8);
    a = ((a >> 1) & 0x0303030303030303ull) ^ ((a >> 2) & 0x0101010101010101ull);
    b = ((b >> 1) & 0x0303030303030303ull) ^ ((b >> 2) & 0x0101010101010101ull);
    const uint32_t hi = (uint32_t)_pext_u64(__builtin_bswap64(a), 0x0303030303030303ull);
    const uint32_t lo = (uint32_t)_pext_u64(__builtin_bswap64(b), 0x0303030303030303ull);
    return (hi << 16) | lo;
}
// 32 A/C/G/T bytes (any case, U) -> two 32-bit units, without leaving the vector registers: 2-bit codes per byte, then
// c0 * 4 + c1 per byte pair (maddubs), (..) * 16 + (..) per pair of those (madd): one byte of four bases per 32-bit lane,
// first base in the top bits; a byte shuffle gathers the eight bytes in unit order (bases 0-3 in the unit's top byte).
__attribute__((target("avx2,bmi2"))) inline void pack32_avx2(__m256i x, uint32_t *dst)
{
    const __m256i m3 = _mm256_set1_epi8(3), m1 = _mm256_set1_epi8(1);
    const __m256i c = _mm256_xor_si256(_mm256_and_si256(_mm256_srli_epi16(x, 1), m3), _mm256_and_si256(_mm256_srli_epi16(x, 2), m1));
    const __m256i p = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0104));          // byte pairs: c0 * 4 + c1
    const __m256i q = _mm256_madd_epi16(p, _mm256_set1_epi32(0x00010010));         // 16-bit pairs: p0 * 16 + p1 -> low byte of each lane
    // lanes 0..3 (bases 0-15) -> bytes 3,2,1,0 of the 128-bit half's first word: unit 0 / unit 1 per half
    const __m256i sh = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i r = _mm256_shuffle_epi8(q, sh);
    dst[0] = (uint32_t)_mm256_extract_epi32(r, 0);
    dst[1] = (uint32_t)_mm256_extract_epi32(r, 4);
}
__attribute__((target("avx2,bmi2"))) bool pack_pure_avx2(const uint8_t *s, uint32_t n, uint32_t *dst, bool fold)
{
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i cU = _mm256_set1_epi8('U'), up = _mm256_set1_epi8((char)0xDF);
    uint32_t i = 0;
    __m256i bad = _mm256_setzero_si256();
    for (; i + 32 <= n; i += 32) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i));
        __m256i ok;
        if (fold) {
            const __m256i y = _mm256_and_si256(x, up);   // a-z -> A-Z (other bytes change too, but never INTO the accepted set:
                                                         // 0x61/0x63/0x67/0x74/0x75 are the only bytes that fold onto A/C/G/T/U)
            ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(y, cA), _mm256_cmpeq_epi8(y, cC)),
                                 _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(y, cG), _mm256_cmpeq_epi8(y, cT)), _mm256_cmpeq_epi8(y, cU)));
        } else {
            ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, cA), _mm256_cmpeq_epi8(x, cC)),
                                 _mm256_or_si256(_mm256_cmpeq_epi8(x, cG), _mm256_cmpeq_epi8(x, cT)));
        }
        bad = _mm256_or_si256(bad, _mm256_xor_si256(ok, _mm256_set1_epi8(-1)));
        pack32_avx2(x, dst + (i >> 4));
        if ((i & 1023u) == 992u && !_mm256_testz_si256(bad, bad)) return false;      // look at the verdict once per KiB
    }
    if (!_mm256_testz_si256(bad, bad)) return false;
    if (i < n) {
        uint8_t tail[32];
        memset(tail, 'A', 32);
        memcpy(tail, s + i, n - i);
        uint32_t b2 = 0;
        for (uint32_t k = i; k < n; k++) { const uint8_t m = fold ? kT.norm[s[k]] : s[k]; b2 |= (m == 0) | kT.cls[m]; }
        if (b2) return false;
        const uint32_t r = n - i;
        uint32_t u[2];
        pack32_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(tail)), u);
        if (r <= 16) { dst[i >> 4] = u[0] & ~(r == 16 ? 0u : (0xffffffffu >> (2 * r))); }
        else { dst[i >> 4] = u[0]; dst[(i >> 4) + 1] = u[1] & ~(r == 32 ? 0u : (0xffffffffu >> (2 * (r - 16)))); }
    }
    return true;
}
// the same 64 bytes at a time (AVX-512 BW + VBMI: one byte permute gathers the sixteen packed bytes in unit order)
__attribute__((target("avx512f,avx512bw,avx512vbmi,avx2,bmi2"))) bool pack_pure_avx512(const uint8_t *s, uint32_t n, uint32_t *dst, bool fold)
{
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'), cT = _mm512_set1_epi8('T');
    const __m512i cU = _mm512_set1_epi8('U'), up = _mm512_set1_epi8((char)0xDF);
    const __m512i m3 = _mm512_set1_epi8(3), m1 = _mm512_set1_epi8(1);
    const __m512i w1 = _mm512_set1_epi16(0x0104), w2 = _mm512_set1_epi32(0x00010010);
    alignas(64) static const uint8_t idx[64] = {12, 8, 4, 0, 28, 24, 20, 16, 44, 40, 36, 32, 60, 56, 52, 48};
    const __m512i perm = _mm512_load_si512(idx);
    uint32_t i = 0;
    __mmask64 bad = 0;
    for (; i + 64 <= n; i += 64) {
        const __m512i x = _mm512_loadu_si512(s + i);
        __mmask64 ok;
        if (fold) {
            const __m512i y = _mm512_and_si512(x, up);
            ok = _mm512_cmpeq_epi8_mask(y, cA) | _mm512_cmpeq_epi8_mask(y, cC) | _mm512_cmpeq_epi8_mask(y, cG) |
                 _mm512_cmpeq_epi8_mask(y, cT) | _mm512_cmpeq_epi8_mask(y, cU);
        } else {
            ok = _mm512_cmpeq_epi8_mask(x, cA) | _mm512_cmpeq_epi8_mask(x, cC) | _mm512_cmpeq_epi8_mask(x, cG) | _mm512_cmpeq_epi8_mask(x, cT);
        }
        bad |= ~ok;
        const __m512i c = _mm512_xor_si512(_mm512_and_si512(_mm512_srli_epi16(x, 1), m3), _mm512_and_si512(_mm512_srli_epi16(x, 2), m1));
        const __m512i q = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + (i >> 4)), _mm512_castsi512_si128(_mm512_permutexvar_epi8(perm, q)));
        if ((i & 1023u) == 960u && bad) return false;
    }
    if (bad) return false;
    if (i < n) return pack_pure_avx2(s + i, n - i, dst + (i >> 4), fold);
    return true;
}
bool have_avx512_vbmi()
{
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
                           __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    return ok;
}
bool have_avx2_bmi2()
{
    static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    return ok;
}
#endif

inline bool pack_pure(const uint8_t *s, uint32_t n, uint32_t *dst, bool fold)
{
#if defined(__x86_64__)
    if (have_avx512_vbmi()) return pack_pure_avx512(s, n, dst, fold);
    if (have_avx2_bmi2()) return pack_pure_avx2(s, n, dst, fold);
#endif
    return pack_pure_scalar(s, n, dst, fold);
}

struct Stash { uint32_t rec; std::vector<uint8_t> bytes; };

}  // namespace

extern "C" {

uint64_t ck_pack2_words(uint64_t total_bytes, uint32_t n_records) { return (total_bytes >> 5) + (uint64_t)n_records + 1; }

int ck_pack2_host(const uint8_t *bytes, const uint64_t *offsets, uint32_t n_records, uint32_t flags, uint32_t threads,
                  uint64_t *packed2_dense, uint32_t *lens, uint8_t *lane, uint8_t *lane_bytes, uint64_t lane_bytes_capacity,
                  uint64_t *lane_offsets, uint64_t *lane_bytes_total)
{
    if (lane_bytes_total) *lane_bytes_total = 0;
    if (n_records == 0) return CK_OK;
    if (!offsets || !packed2_dense || !lens || !lane || (!bytes && offsets[n_records] != offsets[0])) return CK_ERR_ARG;
    const bool fold = (flags & CK_F_NORMALIZE) != 0;
    uint32_t T = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
    const uint64_t total = offsets[n_records] - offsets[0];
    T = (uint32_t)std::min<uint64_t>(T, std::max<uint64_t>(1, std::min<uint64_t>(n_records, (total >> 16) + 1)));
    std::vector<std::vector<Stash>> stash(T);
    std::atomic<int> bad(0);
    auto work = [&](uint32_t t) {
        // contiguous record ranges of (nearly) equal bytes
        const uint64_t lo_b = offsets[0] + total * t / T, hi_b = offsets[0] + total * (t + 1) / T;
        uint32_t r0 = (uint32_t)(std::lower_bound(offsets, offsets + n_records, lo_b) - offsets);
        uint32_t r1 = t + 1 == T ? n_records : (uint32_t)(std::lower_bound(offsets, offsets + n_records, hi_b) - offsets);
        if (t == 0) r0 = 0;
        std::vector<uint8_t> tmp;
        for (uint32_t i = r0; i < r1; i++) {
            const uint64_t off = offsets[i];
            if (offsets[i + 1] < off || offsets[i + 1] - off > (1ull << 30)) { bad = 1; return; }
            const uint32_t raw = (uint32_t)(offsets[i + 1] - off);
            const uint8_t *s = bytes + off;
            uint32_t *dst = reinterpret_cast<uint32_t *>(packed2_dense + (off >> 5) + i);
            if (pack_pure(s, raw, dst, fold)) { lens[i] = raw; lane[i] = 2; continue; }
            // general path: normalise, classify, then pack or stash
            tmp.resize(raw);
            uint32_t n = 0, fl = 0;
            if (fold) {
                for (uint32_t k = 0; k < raw; k++) { const uint8_t m = kT.norm[s[k]]; if (m) { tmp[n++] = m; fl |= kT.cls[m]; } }
            } else {
                for (uint32_t k = 0; k < raw; k++) { tmp[k] = s[k]; fl |= kT.cls[s[k]]; }
                n = raw;
            }
            lens[i] = n;
            if (fl == 0) { lane[i] = 2; pack_scalar(tmp.data(), n, dst); continue; }
            lane[i] = (fl & 2u) ? 8 : 4;
            stash[t].push_back(Stash{i, std::vector<uint8_t>(tmp.begin(), tmp.begin() + n)});
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    if (bad) return CK_ERR_ARG;
    // byte-lane records: positions in lane_bytes (record order), then the copies
    uint64_t run = 0;
    bool any = false;
    for (uint32_t t = 0; t < T; t++) any = any || !stash[t].empty();
    if (lane_offsets) {
        if (any) {
            for (uint32_t i = 0; i < n_records; i++) { lane_offsets[i] = run; if (lane[i] != 2) run += lens[i]; }
            lane_offsets[n_records] = run;
        } else {
            memset(lane_offsets, 0, sizeof(uint64_t) * ((size_t)n_records + 1));
        }
    } else if (any) {
        return CK_ERR_ARG;                                   // byte-lane records need lane_offsets
    }
    if (lane_bytes_total) *lane_bytes_total = run;
    if (!any) return CK_OK;
    if (run > lane_bytes_capacity || !lane_bytes) return CK_ERR_ARG;
    auto copy = [&](uint32_t t) {
        for (const Stash &e : stash[t]) if (!e.bytes.empty()) memcpy(lane_bytes + lane_offsets[e.rec], e.bytes.data(), e.bytes.size());
    };
    if (T == 1) copy(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < T; t++) th.emplace_back(copy, t);
        for (auto &x : th) x.join();
    }
    return CK_OK;
}

}  // extern "C"
