/*
 * ck_oracle.c -- CPU restatement of circKit's canonicalize / uniq hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product path (circkit_b200/) never links, imports or calls this file.
 *
 * The reference (Benjamin-Lee/circkit) is Rust and cannot be compiled in this
 * image (no cargo/rustc), so every function below restates the reference source
 * it cites; the third-party crates the path calls (bio 1.3.1, needletail 0.5.1,
 * xxhash-rust 0.8.6, seq_io 0.3.2 -- pinned in the reference's Cargo.lock, not
 * vendored) are restated from their published algorithms.
 *
 * Parity pinning: tests/test_oracle.py checks this file against every known
 * answer and fixture the reference's own tests hold for the path
 * (lib/src/canonicalize.rs:65-232, tests/canon_uniq.rs, tests/examples/...) and
 * the XXH3 restatement against python-xxhash (libxxhash 0.8.2) at every length
 * class.  Items that no reference test pins (IUPAC complements, needletail's
 * "everything else -> N", seq_io CR handling) are listed in DESIGN.md as
 * "parity unpinned by the reference".
 */
#define _GNU_SOURCE          /* memmem */
#include <stdint.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define CK_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* lmsr_index: lib/src/canonicalize.rs:5-36                                   */
/* Duval-style least rotation with wrap-around reads.  The reference indexes   */
/* with s.chars().nth(i); for ASCII that is the byte at i, read here in O(1).  */
/* ------------------------------------------------------------------------- */
CK_EXPORT size_t ck_o_lmsr_index(const uint8_t *s, size_t len)
{
    int64_t n = (int64_t)len;
    int64_t res = 0, l = 0;
    while (l < n) {                       /* :11 */
        res = l;                          /* :12 */
        int64_t r = l, p = l + 1;         /* :13-14 */
        while (r < n) {                   /* :16 */
            uint8_t c = (p < n) ? s[p] : s[p - n];   /* :17-21 */
            if (s[r] > c) break;          /* :22-24 */
            if (s[r] < c) r = l - 1;      /* :25-27 */
            r += 1;                       /* :28 */
            p += 1;                       /* :29 */
        }
        int64_t a = r, b = l + p - r;     /* :32  l = max(r, l + p - r) */
        l = a > b ? a : b;
    }
    return (size_t)res;
}

/* Same algorithm, but paying the reference's real access cost: s.chars().nth(i)
 * walks i UTF-8 scalars from the start on every access (lib/src/canonicalize.rs
 * :18,:20,:22,:25), so the reference as written is O(n^2).  Used only for the
 * "faithful-cost" CPU footnote in bench.py; the result is identical. */
static inline uint8_t nth_walk(const uint8_t *s, int64_t i)
{
    const volatile uint8_t *p = s;        /* volatile: keep the O(i) walk */
    int64_t k = 0;
    while (k < i) { (void)p[k]; k++; }
    return s[i];
}
CK_EXPORT size_t ck_o_lmsr_index_faithful_cost(const uint8_t *s, size_t len)
{
    int64_t n = (int64_t)len, res = 0, l = 0;
    while (l < n) {
        res = l;
        int64_t r = l, p = l + 1;
        while (r < n) {
            uint8_t c = (p < n) ? nth_walk(s, p) : nth_walk(s, p - n);
            if (nth_walk(s, r) > c) break;
            if (nth_walk(s, r) < c) r = l - 1;
            r += 1; p += 1;
        }
        int64_t a = r, b = l + p - r;
        l = a > b ? a : b;
    }
    return (size_t)res;
}

/* lmsr_index_simple: lib/src/canonicalize.rs:154-164 (brute force; the
 * reference's own cross-check).  First index whose rotation is strictly
 * smaller than the best so far. */
static int rot_cmp(const uint8_t *s, size_t n, size_t a, size_t b)
{
    for (size_t k = 0; k < n; k++) {
        uint8_t x = s[(a + k) % n], y = s[(b + k) % n];
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}
CK_EXPORT size_t ck_o_lmsr_index_simple(const uint8_t *s, size_t n)
{
    size_t result = 0;
    for (size_t i = 0; i < n; i++)
        if (rot_cmp(s, n, i, result) < 0) result = i;
    return result;
}

/* lmsr_index_2: lib/src/canonicalize.rs:168-214 (textbook Duval on s+s). */
CK_EXPORT size_t ck_o_lmsr_index_2(const uint8_t *s0, size_t n)
{
    if (n == 0) return 0;
    uint8_t *s = (uint8_t *)malloc(2 * n);
    memcpy(s, s0, n); memcpy(s + n, s0, n);
    size_t m = 2 * n, res = 0, l = 0;
    while (l < n) {
        res = l;
        size_t r = l, p = l + 1;
        while (p < m) {
            if (s[r] > s[p]) break;
            if (s[r] == s[p]) { r++; p++; continue; }
            r = l; p++;
        }
        while (l <= r) l += p - r;
    }
    free(s);
    return res;
}

/* lmsr: lib/src/canonicalize.rs:41-47 */
CK_EXPORT void ck_o_lmsr(const uint8_t *s, size_t n, uint8_t *out)
{
    size_t i = ck_o_lmsr_index(s, n);
    memcpy(out, s + i, n - i);
    memcpy(out + (n - i), s, i);
}

/* ------------------------------------------------------------------------- */
/* bio 1.3.1 alphabets::dna::{complement, revcomp} (call site                  */
/* lib/src/canonicalize.rs:56).  256-entry table: identity, then               */
/* AGCTYRWSKMDVHBN -> TCGARYWSMKHBDVN and the same pairs +32 (lowercase).      */
/* ------------------------------------------------------------------------- */
static uint8_t COMP[256];
static int comp_ready = 0;
static void comp_init(void)
{
    if (comp_ready) return;
    static const char *a = "AGCTYRWSKMDVHBN", *b = "TCGARYWSMKHBDVN";
    for (int v = 0; v < 256; v++) COMP[v] = (uint8_t)v;
    for (int i = 0; a[i]; i++) {
        COMP[(uint8_t)a[i]] = (uint8_t)b[i];
        COMP[(uint8_t)a[i] + 32] = (uint8_t)(b[i] + 32);
    }
    comp_ready = 1;
}
CK_EXPORT void ck_o_complement_table(uint8_t out[256]) { comp_init(); memcpy(out, COMP, 256); }
CK_EXPORT void ck_o_revcomp(const uint8_t *s, size_t n, uint8_t *out)
{
    comp_init();
    for (size_t i = 0; i < n; i++) out[i] = COMP[s[n - 1 - i]];
}

/* canonicalize: lib/src/canonicalize.rs:54-63.
 * Returns 0 if the forward LMSR was kept (a < b), 1 if the revcomp LMSR was. */
CK_EXPORT int ck_o_canonicalize(const uint8_t *s, size_t n, uint8_t *out)
{
    if (n == 0) return 1;
    uint8_t *a = (uint8_t *)malloc(n), *rc = (uint8_t *)malloc(n), *b = (uint8_t *)malloc(n);
    ck_o_lmsr(s, n, a);                   /* :55 */
    ck_o_revcomp(a, n, rc);               /* :56 */
    ck_o_lmsr(rc, n, b);                  /* :56 */
    int fwd = memcmp(a, b, n) < 0;        /* :58  lmsr_s < lmsr_revcomp_s (unsigned bytes) */
    memcpy(out, fwd ? a : b, n);
    free(a); free(rc); free(b);
    return fwd ? 0 : 1;
}

/* Faithful-cost twin of canonicalize (two quadratic lmsr_index calls). */
CK_EXPORT int ck_o_canonicalize_faithful_cost(const uint8_t *s, size_t n, uint8_t *out)
{
    if (n == 0) return 1;
    uint8_t *a = (uint8_t *)malloc(n), *rc = (uint8_t *)malloc(n), *b = (uint8_t *)malloc(n);
    size_t i = ck_o_lmsr_index_faithful_cost(s, n);
    memcpy(a, s + i, n - i); memcpy(a + (n - i), s, i);
    ck_o_revcomp(a, n, rc);
    size_t j = ck_o_lmsr_index_faithful_cost(rc, n);
    memcpy(b, rc + j, n - j); memcpy(b + (n - j), rc, j);
    int fwd = memcmp(a, b, n) < 0;
    memcpy(out, fwd ? a : b, n);
    free(a); free(rc); free(b);
    return fwd ? 0 : 1;
}

/* (start, strand) of the canonical form in the coordinates of the ORIGINAL
 * record -- the compact answer the device library returns instead of bytes:
 *   strand 0: canonical[j] = s[(start + j) mod n]
 *   strand 1: canonical[j] = complement(s[(start - j) mod n])
 * strand follows :58 (ties -> revcomp); start follows lmsr_index's
 * smallest-index rule applied to s (strand 0) or to revcomp(s) (strand 1). */
CK_EXPORT void ck_o_canonical_start(const uint8_t *s, size_t n, uint32_t *start, uint8_t *strand)
{
    if (n == 0) { *start = 0; *strand = 1; return; }
    uint8_t *rc = (uint8_t *)malloc(n), *a = (uint8_t *)malloc(n), *b = (uint8_t *)malloc(n);
    size_t i = ck_o_lmsr_index(s, n);
    ck_o_revcomp(s, n, rc);
    size_t j = ck_o_lmsr_index(rc, n);
    memcpy(a, s + i, n - i); memcpy(a + (n - i), s, i);
    memcpy(b, rc + j, n - j); memcpy(b + (n - j), rc, j);
    if (memcmp(a, b, n) < 0) { *start = (uint32_t)i; *strand = 0; }
    else { *start = (uint32_t)(n - 1 - j); *strand = 1; }
    free(rc); free(a); free(b);
}

/* ------------------------------------------------------------------------- */
/* needletail 0.5.1 sequence::normalize(seq, allow_iupac = false)              */
/* (call sites src/canonicalize.rs:24-27, src/uniq.rs:35-38).                  */
/* Returns the normalised length; *changed mirrors Some(..)/None (the caller   */
/* falls back to the raw bytes when nothing changed, which is the same bytes). */
/* ------------------------------------------------------------------------- */
CK_EXPORT size_t ck_o_normalize(const uint8_t *seq, size_t n, uint8_t *out, int *changed)
{
    size_t m = 0; int ch = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = seq[i], o;
        switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case '-': o = c; break;
        case 'a': o = 'A'; ch = 1; break;
        case 'c': o = 'C'; ch = 1; break;
        case 'g': o = 'G'; ch = 1; break;
        case 't': case 'u': case 'U': o = 'T'; ch = 1; break;
        case '.': case '~': o = '-'; ch = 1; break;
        case ' ': case '\t': case '\r': case '\n': o = ' '; ch = 1; break;
        default: o = 'N'; ch = 1; break;   /* incl. every IUPAC code: allow_iupac is false */
        }
        if (o != ' ') out[m++] = o;
    }
    if (changed) *changed = ch;
    return m;
}

/* ------------------------------------------------------------------------- */
/* xxhash-rust 0.8.6 xxh3::xxh3_64 (call site src/uniq.rs:45): one-shot        */
/* XXH3-64, seed 0, default 192-byte secret.  Restated from the XXH3 spec      */
/* (xxHash 0.8.x, XXH3_64bits).                                                */
/* ------------------------------------------------------------------------- */
static const uint8_t kSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};
#define P32_1 0x9E3779B1U
#define P32_2 0x85EBCA77U
#define P32_3 0xC2B2AE3DU
#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P64_5 0x27D4EB2F165667C5ULL
#define PMX1  0x165667919E3779F9ULL
#define PMX2  0x9FB21C651E98DF25ULL

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }   /* little-endian host */
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t mul128_fold64(uint64_t a, uint64_t b)
{
    __uint128_t p = (__uint128_t)a * b;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
}
static inline uint64_t xxh64_avalanche(uint64_t h)
{
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h;
}
static inline uint64_t xxh3_avalanche(uint64_t h)
{
    h ^= h >> 37; h *= PMX1; h ^= h >> 32; return h;
}
static inline uint64_t rrmxmx(uint64_t h, uint64_t len)
{
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= PMX2;
    h ^= (h >> 35) + len;
    h *= PMX2;
    return h ^ (h >> 28);
}
static inline uint64_t mix16(const uint8_t *in, const uint8_t *sec)
{
    return mul128_fold64(rd64(in) ^ rd64(sec), rd64(in + 8) ^ rd64(sec + 8));   /* seed = 0 */
}
static inline void accumulate_512(uint64_t acc[8], const uint8_t *in, const uint8_t *sec)
{
    for (int i = 0; i < 8; i++) {
        uint64_t dv = rd64(in + 8 * i);
        uint64_t dk = dv ^ rd64(sec + 8 * i);
        acc[i ^ 1] += dv;
        acc[i] += (dk & 0xFFFFFFFFULL) * (dk >> 32);
    }
}
CK_EXPORT uint64_t ck_o_xxh3_64(const uint8_t *in, size_t len)
{
    const uint8_t *sec = kSecret;
    if (len == 0) return xxh64_avalanche(rd64(sec + 56) ^ rd64(sec + 64));
    if (len <= 3) {
        uint32_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t comb = (c1 << 16) | (c2 << 24) | c3 | ((uint32_t)len << 8);
        uint64_t flip = (uint64_t)(rd32(sec) ^ rd32(sec + 4));
        return xxh64_avalanche((uint64_t)comb ^ flip);
    }
    if (len <= 8) {
        uint32_t i1 = rd32(in), i2 = rd32(in + len - 4);
        uint64_t flip = rd64(sec + 8) ^ rd64(sec + 16);
        uint64_t in64 = (uint64_t)i2 + ((uint64_t)i1 << 32);
        return rrmxmx(in64 ^ flip, len);
    }
    if (len <= 16) {
        uint64_t f1 = rd64(sec + 24) ^ rd64(sec + 32), f2 = rd64(sec + 40) ^ rd64(sec + 48);
        uint64_t lo = rd64(in) ^ f1, hi = rd64(in + len - 8) ^ f2;
        uint64_t acc = len + __builtin_bswap64(lo) + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (len <= 128) {
        uint64_t acc = len * P64_1;
        int i = (int)((len - 1) / 32);
        do {
            acc += mix16(in + 16 * i, sec + 32 * i);
            acc += mix16(in + len - 16 * (i + 1), sec + 32 * i + 16);
        } while (i-- != 0);
        return xxh3_avalanche(acc);
    }
    if (len <= 240) {
        uint64_t acc = len * P64_1, acc_end;
        unsigned rounds = (unsigned)len / 16;
        for (unsigned i = 0; i < 8; i++) acc += mix16(in + 16 * i, sec + 16 * i);
        acc_end = mix16(in + len - 16, sec + 136 - 17);
        acc = xxh3_avalanche(acc);
        for (unsigned i = 8; i < rounds; i++) acc_end += mix16(in + 16 * i, sec + 16 * (i - 8) + 3);
        return xxh3_avalanche(acc + acc_end);
    }
    {
        uint64_t acc[8] = { P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1 };
        const size_t stripes_per_block = (192 - 64) / 8;       /* 16 */
        const size_t block_len = 64 * stripes_per_block;       /* 1024 */
        size_t nb_blocks = (len - 1) / block_len;
        for (size_t b = 0; b < nb_blocks; b++) {
            for (size_t s = 0; s < stripes_per_block; s++)
                accumulate_512(acc, in + b * block_len + 64 * s, sec + 8 * s);
            for (int i = 0; i < 8; i++) {                      /* scramble */
                uint64_t a = acc[i];
                a ^= a >> 47; a ^= rd64(sec + 192 - 64 + 8 * i); a *= P32_1;
                acc[i] = a;
            }
        }
        size_t nb_stripes = ((len - 1) - block_len * nb_blocks) / 64;
        for (size_t s = 0; s < nb_stripes; s++)
            accumulate_512(acc, in + nb_blocks * block_len + 64 * s, sec + 8 * s);
        accumulate_512(acc, in + len - 64, sec + 192 - 64 - 7);
        uint64_t r = len * P64_1;
        for (int i = 0; i < 4; i++)
            r += mul128_fold64(acc[2 * i] ^ rd64(sec + 11 + 16 * i), acc[2 * i + 1] ^ rd64(sec + 11 + 16 * i + 8));
        return xxh3_avalanche(r);
    }
}

/* ------------------------------------------------------------------------- */
/* Batch drivers (the two closures of src/canonicalize.rs:21-44 and            */
/* src/uniq.rs:33-78 applied to an in-memory batch).                           */
/*   bytes/offsets : concatenated raw record.seq() bytes, n_records+1 offsets  */
/*   flags bit0    : 1 = needletail-normalise first (CLI semantics),           */
/*                   0 = library semantics (bytes used as they are)            */
/*   out           : canonical bytes, record i at out[offsets[i]..+out_len[i]] */
/* Worker stage is sharded over `threads` (parallel_fasta's worker pool).      */
/* ------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *bytes; const uint64_t *offsets; uint64_t n_records; uint32_t flags;
    uint8_t *out; uint32_t *out_len; uint32_t *out_start; uint8_t *out_strand; uint64_t *out_hash;
    int faithful_cost; atomic_ullong next;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const uint64_t chunk = 64;            /* record sets, like parallel_fasta's queue items */
    for (;;) {
        uint64_t b = atomic_fetch_add(&j->next, chunk);
        if (b >= j->n_records) break;
        uint64_t e = b + chunk < j->n_records ? b + chunk : j->n_records;
        for (uint64_t i = b; i < e; i++) {
            const uint8_t *s = j->bytes + j->offsets[i];
            size_t n = (size_t)(j->offsets[i + 1] - j->offsets[i]);
            uint8_t *norm = NULL;
            if (j->flags & 1u) {              /* src/canonicalize.rs:24-27 */
                norm = (uint8_t *)malloc(n ? n : 1);
                n = ck_o_normalize(s, n, norm, NULL);
                s = norm;
            }
            uint8_t *dst = j->out + j->offsets[i];
            if (j->faithful_cost) ck_o_canonicalize_faithful_cost(s, n, dst);
            else ck_o_canonicalize(s, n, dst);    /* src/canonicalize.rs:29 */
            if (j->out_len) j->out_len[i] = (uint32_t)n;
            if (j->out_start && j->out_strand) ck_o_canonical_start(s, n, &j->out_start[i], &j->out_strand[i]);
            if (j->out_hash) j->out_hash[i] = ck_o_xxh3_64(dst, n);
            free(norm);
        }
    }
    return NULL;
}

CK_EXPORT int ck_o_canonicalize_batch(const uint8_t *bytes, const uint64_t *offsets, uint64_t n_records,
                                      uint32_t flags, int threads, uint8_t *out, uint32_t *out_len,
                                      uint32_t *out_start, uint8_t *out_strand, uint64_t *out_hash,
                                      int faithful_cost)
{
    comp_init();
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    batch_job job = { bytes, offsets, n_records, flags, out, out_len, out_start, out_strand, out_hash,
                      faithful_cost, 0 };
    atomic_init(&job.next, 0);
    if (threads == 1) { batch_worker(&job); return 0; }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&tid[t], NULL, batch_worker, &job) == 0) tid[started++] = tid[t];
    if (started == 0) batch_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    free(tid);
    return 0;
}

/* Consumer stage of uniq, src/uniq.rs:42-78, strictly serial as in the
 * reference: hash the canonical bytes (:45), keep the first record seen with
 * each 64-bit hash (:47-48).  first_index[i] = input index of the first record
 * with record i's hash (== i  <=>  record i is written).  Open-addressing map
 * standing in for HashMap<u64, String, NoHash>; the decision depends only on
 * hash equality, exactly as in the reference (64-bit collisions merge). */
CK_EXPORT int ck_o_uniq_consume(const uint8_t *canon, const uint64_t *offsets, const uint32_t *lens,
                                uint64_t n_records, uint64_t *out_hash, uint64_t *first_index)
{
    uint64_t cap = 16;
    while (cap < 2 * n_records + 2) cap <<= 1;
    uint64_t *keys = (uint64_t *)malloc(cap * 8), *vals = (uint64_t *)malloc(cap * 8);
    uint8_t *used = (uint8_t *)calloc(cap, 1);
    if (!keys || !vals || !used) { free(keys); free(vals); free(used); return -1; }
    for (uint64_t i = 0; i < n_records; i++) {
        uint64_t h = ck_o_xxh3_64(canon + offsets[i], lens ? lens[i] : (size_t)(offsets[i + 1] - offsets[i]));
        if (out_hash) out_hash[i] = h;
        uint64_t s = (h * 0x9E3779B97F4A7C15ULL) & (cap - 1);
        for (;;) {
            if (!used[s]) { used[s] = 1; keys[s] = h; vals[s] = i; first_index[i] = i; break; }
            if (keys[s] == h) { first_index[i] = vals[s]; break; }
            s = (s + 1) & (cap - 1);
        }
    }
    free(keys); free(vals); free(used);
    return 0;
}

CK_EXPORT int ck_o_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

/* ---------------------------------------------------------------------------------------------
 * monomerize: lib/src/monomerize.rs:50-141 restated in C -- the compiled CPU baseline of
 * tools/time_monomerize.py and a second restatement beside oracle/monomerize.py (tests/test_oracle_monomerize.py holds
 * both to the reference's 277 unit-test vectors).  memmem stands in for bio's ShiftAnd::find_all: every occurrence of
 * the seed in seq[.. n - seed_len], overlapping ones included, by increasing start.  NONE = SIZE_MAX.
 * identity < 0: not set (then max distance = overlap_dist). */
static size_t mono_first(const uint8_t *seq, size_t n, size_t s, uint64_t overlap_dist, double identity)
{
    if (n <= s || n < 2 * s) return SIZE_MAX;
    const uint8_t *seed = seq + (n - s);
    const size_t text_len = n - s;
    size_t from = 0;
    while (from + s <= text_len) {
        const uint8_t *hit = (const uint8_t *)memmem(seq + from, text_len - from, seed, s);
        if (!hit) break;
        const size_t occ = (size_t)(hit - seq), m = occ + s;
        const uint8_t *starter = seq + (n - m);
        uint64_t dist = 0;
        for (size_t t = 0; t < m; t++) dist += starter[t] != seq[t];          /* bio hamming(starter, successor) */
        const uint64_t maxd = identity >= 0.0 ? (uint64_t)m - (uint64_t)floor((double)m * identity) : overlap_dist;
        if (dist <= maxd) return n - m;
        from = occ + 1;
    }
    return SIZE_MAX;
}
CK_EXPORT size_t ck_o_monomer_end(const uint8_t *seq, size_t n, size_t seed_len, uint64_t overlap_dist, double identity,
                                  int sensitive, int first_only)
{
    size_t idx = mono_first(seq, n, seed_len, overlap_dist, identity);
    if (first_only) return idx;
    while (idx != SIZE_MAX) {                                                  /* :100-125 */
        const size_t nxt = mono_first(seq, idx, seed_len, overlap_dist, identity);
        if (nxt == SIZE_MAX) break;
        idx = nxt;
    }
    if (sensitive) {                                                           /* :127-141 */
        comp_init();
        const size_t M = idx == SIZE_MAX ? n : idx;
        uint8_t *rc = (uint8_t *)malloc(M ? M : 1);
        for (size_t j = 0; j < M; j++) rc[j] = COMP[seq[M - 1 - j]];
        const size_t k = mono_first(rc, M, seed_len, overlap_dist, identity);
        free(rc);
        if (k != SIZE_MAX) idx = M - (M - k);
    }
    return idx;
}
typedef struct {
    const uint8_t *bytes; const uint64_t *offsets; uint64_t n_records; size_t seed_len; uint64_t overlap_dist;
    double identity; int sensitive; uint32_t *out; atomic_ullong next;
} mono_job;
static void *mono_worker(void *arg)
{
    mono_job *j = (mono_job *)arg;
    for (;;) {
        const uint64_t lo = atomic_fetch_add(&j->next, 256);
        if (lo >= j->n_records) break;
        const uint64_t hi = lo + 256 < j->n_records ? lo + 256 : j->n_records;
        for (uint64_t i = lo; i < hi; i++) {
            const size_t e = ck_o_monomer_end(j->bytes + j->offsets[i], (size_t)(j->offsets[i + 1] - j->offsets[i]), j->seed_len,
                                              j->overlap_dist, j->identity, j->sensitive, 0);
            j->out[i] = e == SIZE_MAX ? 0xffffffffu : (uint32_t)e;
        }
    }
    return NULL;
}
CK_EXPORT int ck_o_monomerize_batch(const uint8_t *bytes, const uint64_t *offsets, uint64_t n_records, size_t seed_len,
                                    uint64_t overlap_dist, double identity, int sensitive, int threads, uint32_t *out_end)
{
    comp_init();
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    mono_job job = { bytes, offsets, n_records, seed_len, overlap_dist, identity, sensitive, out_end, 0 };
    atomic_init(&job.next, 0);
    if (threads == 1) { mono_worker(&job); return 0; }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&tid[t], NULL, mono_worker, &job) == 0) tid[started++] = tid[t];
    if (started == 0) mono_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    free(tid);
    return 0;
}
