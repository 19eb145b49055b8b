// ck_device.cuh -- device-side building blocks shared by the canonicalize/uniq kernels.
//
// Symbol lanes.  A record is a circular string over an ordered alphabet; the kernels work on
// a packed big-endian image of it so that unsigned integer order == lexicographic order
// (reference compares unsigned bytes: lib/src/canonicalize.rs:22,25,58):
//   BITS = 2 : A,C,G,T -> 0..3                       (16 symbols per 32-bit unit)
//   BITS = 4 : - A B C D G H K M N R S T V W Y -> 0..15  (8 per unit; covers the CLI alphabet
//              {-,A,C,G,N,T} that needletail::normalize emits and upper-case IUPAC for the lib API)
//   BITS = 8 : raw bytes                             (4 per unit; any other input, lib semantics)
// In shared memory a strand is an array X of 32-bit units, unit j holding symbols
// [j*S, (j+1)*S) MSB-first, extended circularly by >= 64 symbols past n so that a window that
// starts at any p < n can be read without wrap logic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

#define CK_FULL 0xffffffffu

namespace ck {

// ------------------------------------------------------------------ constant tables
// Filled once per process by ck_tables_init() (ck_kernels.cu).
struct Tables {
    u8 norm[256];    // needletail 0.5.1 normalize(_, false): 0 = dropped, else mapped byte (src/canonicalize.rs:24)
    u8 comp[256];    // bio 1.3.1 complement LUT (lib/src/canonicalize.rs:56)
    u8 code4[256];   // byte -> 4-bit code of the 16-symbol alphabet, 0xFF = not representable
    u8 sym4[16];     // 4-bit code -> byte
    u8 comp4[16];    // complement in 4-bit code space
};
__constant__ Tables c_tab;   // single translation unit (ck_lib.cu): defined here

// Per-CTA copy of the tables the byte-level lanes index with per-thread symbols.  A constant-bank load with a divergent
// index is replayed once per distinct address (ncu r01 p17: LDC = 38 % of the stall samples of k_canon_warp<4>); the same
// lookups in shared memory are one wavefront (16-entry tables: 4 words, 4 banks) or two (256-entry tables).  Kernels that
// use the 4- or 8-bit lanes call stab_load() first; the 2-bit kernels never reference it.
struct SharedTables {
    u8 code4[256];
    u8 comp[256];
    u8 sym4[16];
    u8 comp4[16];
};
__shared__ SharedTables s_tab;
__device__ __forceinline__ void stab_load()
{
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) { s_tab.code4[i] = c_tab.code4[i]; s_tab.comp[i] = c_tab.comp[i]; }
    if (threadIdx.x < 16) { s_tab.sym4[threadIdx.x] = c_tab.sym4[threadIdx.x]; s_tab.comp4[threadIdx.x] = c_tab.comp4[threadIdx.x]; }
    __syncthreads();
}

// ------------------------------------------------------------------ packed2 arena
// 2-bit arena: A,C,G,T = 0..3, 16 bases per 32-bit unit, first base in the top bits, units in address order.  Record i
// starts at 32-byte granule (offsets[i] >> 6) + 3 i, so that every record can be streamed with aligned 256-bit loads
// (LDG.E.ENL2.256, sm_100: a lane-private 32-byte load costs one L1TEX wavefront per lane, the same as a 16-byte one), and
// it is stored DOUBLED: units 0 .. (2 n >> 4) + 1 hold the bases S[b mod n] (k_extend_packed2 appends the second copy).
// Every rotation of the circular record -- and the 128 + 15 bases an output round reads from it -- is then a linear
// window of the arena: neither the scan nor the emit pass ever wraps.  The host packer only writes the first copy.
// A batch whose 2-bit records are all short (<= 512 bases: viroid lengths, configs 1 and 5) uses the SINGLE-COPY layout
// instead (dbl == 0): record i at 16-byte granule (offsets[i] >> 6) + 2 i, followed by its circular extension only (units
// 0 .. (n >> 4) + 4 hold S[b mod n]) -- 113 instead of 258 bytes of arena per 325-base record, everything the emit pass reads
// is what the scan just read, and the lane kernel of round 1 (ck_stream2.cuh: 128-bit loads, wrap through the extension)
// takes it.  The layout is a property of the batch: the packer, k_extend_packed2 and every kernel get the same flag.
__host__ __device__ __forceinline__ u64 p2_word(u64 off, u64 rec, u32 dbl)      // u64 word index
{
    return dbl ? 4 * ((off >> 6) + 3 * rec) : 2 * ((off >> 6) + 2 * rec);
}
__host__ __device__ __forceinline__ u64 p2_words(u64 total, u64 n_records, u32 dbl = 1)
{
    return dbl ? 4 * ((total >> 6) + 3 * n_records + 3) : 2 * ((total >> 6) + 2 * n_records + 2);
}
// aligned output arena (CK_F_ALIGNED_OUT): record i's canonical bytes start at byte 32 * ((offsets[i] >> 5) + i).  32 bytes =
// one DRAM / L2 sector: a 64-byte output round of the lane kernel then covers whole sectors only (with 16-byte alignment half
// of the records wrote half sectors at both ends of every round, which the memory system pays for with fill reads).
__host__ __device__ __forceinline__ u64 out_byte(u64 off, u64 rec) { return 32 * ((off >> 5) + rec); }
__host__ __device__ __forceinline__ u64 out_bytes_total(u64 total, u64 n_records) { return 32 * ((total >> 5) + n_records) + 32; }

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ u32 funnel_l(u32 hi, u32 lo, u32 s)   // top 32 bits of (hi:lo) << s, s in [0,31]
{
    return __funnelshift_l(lo, hi, s);
}

// Reverse-complement of one 32-bit unit of 2-bit symbols (16 bases): reverse the order of the
// 2-bit groups and complement each (3 - c == ~c on two bits).
__device__ __forceinline__ u32 revcomp2_u32(u32 x)
{
    u32 r = __brev(x);                                            // reverses bits, also inside each pair
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);      // swap the two bits of each pair back
    return ~r;
}
// Same for 4-bit symbols (8 per unit): reverse nibbles, complement through the 16-entry table.
__device__ __forceinline__ u32 revcomp4_u32(u32 x)
{
    u32 r = __byte_perm(x, 0, 0x0123);                            // reverse bytes
    r = ((r >> 4) & 0x0f0f0f0fu) | ((r & 0x0f0f0f0fu) << 4);      // swap nibbles inside bytes
    u32 o = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) o |= (u32)s_tab.comp4[(r >> (4 * k)) & 15u] << (4 * k);
    return o;
}
__device__ __forceinline__ u32 revcomp8_u32(u32 x)
{
    u32 r = __byte_perm(x, 0, 0x0123);
    u32 o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) o |= (u32)s_tab.comp[(r >> (8 * k)) & 255u] << (8 * k);
    return o;
}
template <int BITS> __device__ __forceinline__ u32 revcomp_unit(u32 x)
{
    if (BITS == 2) return revcomp2_u32(x);
    if (BITS == 4) return revcomp4_u32(x);
    return revcomp8_u32(x);
}

// ASCII of 4 consecutive 2-bit symbols given as the byte b = c0c1c2c3 (c0 first / most significant):
// returns a little-endian u32 whose byte k is the letter of c_k.
__device__ __forceinline__ u32 ascii4_from_2bit(u32 b)
{
    // selector nibble k of PRMT picks the output byte k from the "ACGT" table
    u32 sel = ((b >> 6) & 3u) | (((b >> 4) & 3u) << 4) | (((b >> 2) & 3u) << 8) | ((b & 3u) << 12);
    return __byte_perm(0x54474341u /* 'A','C','G','T' */, 0, sel);
}

// ------------------------------------------------------------------ XXH3-64 (seed 0, default secret)
// xxhash-rust 0.8.6 xxh3_64 (src/uniq.rs:45), restated from the XXH3 spec.  Secret as LE u64s at
// byte offsets 8*i is kept in constant memory; unaligned secret reads are composed from it.
__constant__ __align__(16) u8 c_secret[192 + 8];

__device__ __forceinline__ u64 sec64(int off)        // XXH_readLE64(secret + off), any alignment
{
    const u32 *w = reinterpret_cast<const u32 *>(c_secret);
    int k = off >> 2, s = (off & 3) * 8;
    u32 a = w[k], b = w[k + 1], c = w[k + 2];
    u32 lo = s ? __funnelshift_r(a, b, s) : a;
    u32 hi = s ? __funnelshift_r(b, c, s) : b;
    return ((u64)hi << 32) | lo;
}
#define CK_P32_1 0x9E3779B1ULL
#define CK_P32_2 0x85EBCA77ULL
#define CK_P32_3 0xC2B2AE3DULL
#define CK_P64_1 0x9E3779B185EBCA87ULL
#define CK_P64_2 0xC2B2AE3D27D4EB4FULL
#define CK_P64_3 0x165667B19E3779F9ULL
#define CK_P64_4 0x85EBCA77C2B2AE63ULL
#define CK_P64_5 0x27D4EB2F165667C5ULL
#define CK_PMX1 0x165667919E3779F9ULL
#define CK_PMX2 0x9FB21C651E98DF25ULL

__device__ __forceinline__ u64 mul128_fold64(u64 a, u64 b) { return (a * b) ^ __umul64hi(a, b); }
__device__ __forceinline__ u64 xxh3_avalanche(u64 h) { h ^= h >> 37; h *= CK_PMX1; h ^= h >> 32; return h; }
__device__ __forceinline__ u64 xxh64_avalanche(u64 h)
{
    h ^= h >> 33; h *= CK_P64_2; h ^= h >> 29; h *= CK_P64_3; h ^= h >> 32; return h;
}
__device__ __forceinline__ u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 bswap64(u64 x)
{
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}

}  // namespace ck
