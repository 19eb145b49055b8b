// ck_lane4.cuh -- lane-per-record kernel for the CLI alphabet {-, A, C, G, N, T}: the records of config 3 of BASELINE.json
// (mixed IUPAC / N-containing sequences) once needletail's normalisation (src/canonicalize.rs:24-27) has mapped every IUPAC
// code to N.  The 2-bit lane cannot hold six symbols and the generic warp-per-record byte-lane kernel (k_canon_warp<4>) spends
// 957 warp instructions on a 325-symbol record; this is the structure of ck_stream3.cuh at 4 bits per symbol:
//   * k_pack4   : per record, BOTH strands packed (8 symbols per 32-bit unit, first symbol in the top nibble, codes - A C G N T =
//                 0..5 in byte order, so integer order == lexicographic order), 32-byte aligned, each CONTINUED circularly past its end by
//                 at least 72 symbols (every 64-symbol window that starts below n is then linear) -- the reverse
//                 complement is a string of its own here (a 4-bit complement is a table lookup, not one LOP3 as for 2 bits),
//                 so the kernel below only ever walks forward, on either strand array, with 256-bit loads; the one wrap of the
//                 emit pass is a reload of octs 0 and 1.  (The first form stored both strands doubled, as ck_stream3.cuh does:
//                 no wrap at all, but 850 instead of 520 bytes per 325-symbol record made the kernel DRAM-bound.)
//                 Records with another symbol (library semantics with IUPAC letters) keep lane 4 and the generic kernel;
//   * scan      : one oct (64 symbols) per iteration and strand; per unit 7 funnel shifts give the eight 8-mer keys (32 bits)
//                 of its rotations, 4 VIMNMX their minimum; the octs' minima are tracked with their (oct, strand) tag, two
//                 smallest per lane; the one partial oct of a record is scanned with its positions past n masked;
//   * locate    : the winning oct is replayed, 64 rotations against the minimal key; a unique hit is the canonical rotation,
//                 anything else (and every record below 129 symbols, 241 with a hash) goes to the generic kernel's duel path;
//   * emit      : 64 canonical symbols per round = one XXH3 stripe, from circular position (start + 64 s) mod n: 9 units
//                 selected from the two octs the lane holds, eight
//                 windows, letters by PRMT in the 8-byte table "-ACGNT" (codes < 8: one lookup per four symbols), the conflict-
//                 free shared-memory stage of ck_stream3.cuh, 128-bit stores.
#pragma once
#include "ck_seg2.cuh"

namespace ck {

#define CK_L4_WARPS 8u
#define CK_L4_WARP_BYTES (CK_S3_STAGE + 1024u + 384u)

// byte offset of record i's packed4 region, bytes of one strand, size of the whole arena
// (a strand = the record once plus its circular continuation by at least 72 symbols: what the partial oct of the scan, an
// output round that starts below n, and the last XXH3 stripe read past n)
__host__ __device__ __forceinline__ u64 p4_byte(u64 off, u64 rec) { return 32 * ((off >> 5) + 8 * rec); }
__host__ __device__ __forceinline__ u32 p4_strand_bytes(u32 n) { return 32u * (((n + 135u) >> 6) + 1u); }
__host__ __device__ __forceinline__ u64 p4_bytes_total(u64 total, u64 n_records) { return 32 * ((total >> 5) + 8 * n_records + 10); }

// four normalised bytes (first symbol in the low byte) -> their four codes as nibbles, first symbol in the TOP nibble of the
// 16 bits; `bad` collects the bytes (under `mask`) that are not one of - A C G N T.  (b >> 1) & 7 is a perfect hash of the six
// letters (A 0, C 1, T 2, G 3, - 6, N 7): two PRMT table lookups give the codes and, for the check, the letters back.
__device__ __forceinline__ u32 l4_codes4(u32 x, u32 mask, u32 &bad)
{
    const u32 idx = (x >> 1) & 0x07070707u;
    const u32 sel = __byte_perm(idx | (idx >> 4), 0, 0x4420);
    const u32 codes = __byte_perm(0x03050201u, 0x04000f0fu, sel);         // A 1, C 2, T 5, G 3 | -, -, '-' 0, N 4
    const u32 back = __byte_perm(0x47544341u, 0x4e2d0000u, sel);          // "ACTG" | 0, 0, "-N"
    bad |= (back ^ x) & mask;
    const u32 r = __byte_perm(codes, 0, 0x0123);
    return __byte_perm(r | (r >> 4), 0, 0x4420);
}
// complement of eight codes at once: A 1 <-> T 5 (bit 2 flips when bits 1:0 == 01), C 2 <-> G 3 (bit 0 flips when bits 2:1 == 01)
__device__ __forceinline__ u32 l4_comp8(u32 x)
{
    const u32 m4 = x & ~(x >> 1) & 0x11111111u, m1 = (x >> 1) & ~(x >> 2) & 0x11111111u;
    return x ^ ((m4 << 2) | m1);
}
__device__ __forceinline__ u32 l4_nibrev(u32 x)
{
    const u32 r = __byte_perm(x, 0, 0x0123);
    return ((r & 0x0f0f0f0fu) << 4) | ((r >> 4) & 0x0f0f0f0fu);
}
// the 8 symbols of the circle that start at symbol q (< n) of a first copy X that l4_close() has continued past n
__device__ __forceinline__ u32 l4_win(const u32 *X, u32 q)
{
    const u32 k = q >> 3;
    return __funnelshift_l(X[k + 1], X[k], 4u * (q & 7u));
}
// continue a zero-padded first copy (NU units, n symbols) circularly: the free nibbles of its last unit and two more units
__device__ __forceinline__ void l4_close(u32 *X, u32 n, u32 NU)
{
    const u32 v = n - 8u * (NU - 1u), f0 = X[0], f1 = X[1];          // 1..8 symbols in the last unit
    if (v < 8u) {
        const u32 s = 4u * (8u - v);
        X[NU - 1u] |= f0 >> (4u * v);
        X[NU] = __funnelshift_l(f1, f0, s);
        X[NU + 1u] = __funnelshift_l(X[2], f1, s);
    } else { X[NU] = f0; X[NU + 1u] = f1; }
}

// Half a warp per record of the 4-bit lane with 129 <= n <= 2048 (a 325-symbol record has 41 units: 16 lanes use 85 % of
// their iterations, 32 lanes 64 %): both strands, each continued circularly; lane_bits[i] becomes 3.
//   A: 8 bytes per lane -> one unit of the forward first copy (shared memory, then continued circularly by two units so that
//      every window of the circle is one funnel shift), alphabet check;
//   B: the reverse complement's first copy from it (window of the forward copy, nibbles reversed, complemented);
//   C: units of the continued strands = windows of the first copies at 8 j mod n, 64-byte coalesced stores.
#define CK_P4_UNITS 260u
__global__ void __launch_bounds__(256) k_pack4(const u8 *bytes, const u64 *offsets, const u32 *lens, u8 *lane_bits, u32 n_records, u8 *p4)
{
    __shared__ u32 sh[16][2][CK_P4_UNITS];
    const u32 lane = lane_id(), hl = lane & 15u, half = lane >> 4;
    const u32 hmask = half ? 0xffff0000u : 0x0000ffffu;
    u32 *F = sh[threadIdx.x >> 4][0], *R = sh[threadIdx.x >> 4][1];
    const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    // the next pair's tag / offset / length are requested, and its bytes prefetched into L2, while this pair is packed
    u32 nx_i = 2u * gw + half;
    bool nx_act = nx_i < n_records && lane_bits[nx_i] == 4;
    u64 nx_off = nx_act ? offsets[nx_i] : 0;
    u32 nx_n = nx_act ? (lens ? lens[nx_i] : (u32)(offsets[nx_i + 1] - nx_off)) : 0u;
    for (u32 i0 = 2u * gw; i0 < n_records; i0 += 2u * nw) {
        const u32 i = nx_i;
        const u64 off = nx_off;
        const u32 n = nx_n;
        bool act = nx_act && n >= 129u && n <= 2048u;
        nx_i = i0 + 2u * nw + half;
        nx_act = nx_i < n_records && lane_bits[nx_i] == 4;
        nx_off = nx_act ? offsets[nx_i] : 0;
        nx_n = nx_act ? (lens ? lens[nx_i] : (u32)(offsets[nx_i + 1] - nx_off)) : 0u;
        if (nx_act && 128u * hl < nx_n + 127u && nx_n <= 2048u) prefetch_l2(bytes + (nx_off & ~127ull) + 128u * hl);
        const u32 NU = (n + 7u) >> 3;
        const u32 *w = reinterpret_cast<const u32 *>(bytes + (off & ~3ull));
        const u32 sh8 = 8u * (u32)(off & 3u);
        u32 bad = 0;
        for (u32 j = hl; j < (act ? NU + 2u : 0u); j += 16) {
            u32 u = 0;
            if (j < NU) {
                const u32 w0 = __ldg(w + 2 * j), w1 = __ldg(w + 2 * j + 1), w2 = __ldg(w + 2 * j + 2);
                const u32 x0 = __funnelshift_r(w0, w1, sh8), x1 = __funnelshift_r(w1, w2, sh8);
                const u32 v = min(8u, n - 8u * j);                         // symbols of this unit
                const u32 m0 = v >= 4u ? 0xffffffffu : ~(0xffffffffu << (8u * v));
                const u32 m1 = v >= 8u ? 0xffffffffu : v > 4u ? ~(0xffffffffu << (8u * (v - 4u))) : 0u;
                u = (l4_codes4(x0, m0, bad) << 16) | l4_codes4(x1, m1, bad);
                if (v < 8u) u &= ~(0xffffffffu >> (4u * v));
            }
            F[j] = u;
        }
        __syncwarp();
        if (__ballot_sync(CK_FULL, bad != 0) & hmask) act = false;         // another alphabet: the generic byte-lane kernels take it
        if (act && hl == 0) l4_close(F, n, NU);
        __syncwarp();
        for (u32 j = hl; j < (act ? NU + 2u : 0u); j += 16) {
            u32 u = 0;
            if (j < NU) {
                const u32 v = min(8u, n - 8u * j);
                const int p = (int)n - 8 - 8 * (int)j;                     // the forward window that ends at symbol n - 1 - 8 j (mod n)
                u = l4_comp8(l4_nibrev(l4_win(F, (u32)(p < 0 ? p + (int)n : p))));
                if (v < 8u) u &= ~(0xffffffffu >> (4u * v));
            }
            R[j] = u;
        }
        __syncwarp();
        if (act && hl == 0) l4_close(R, n, NU);
        __syncwarp();
        const u32 U = act ? p4_strand_bytes(n) >> 2 : 0u;
        u32 *Fo = reinterpret_cast<u32 *>(p4 + p4_byte(off, i)), *Ro = Fo + U;
        u32 q = 8u * hl;                                                   // 8 j mod n: 128 symbols per step, n >= 129
        for (u32 j = hl; j < U; j += 16) {
            const u32 k = q >> 3, r = 4u * (q & 7u);
            Fo[j] = __funnelshift_l(F[k + 1], F[k], r);
            Ro[j] = __funnelshift_l(R[k + 1], R[k], r);
            q += 128u;
            if (q >= n) q -= n;
        }
        if (act && hl == 0) lane_bits[i] = 3;
        __syncwarp();
    }
}

// minimum 8-mer key (32 bits) over the 8 rotations that start in unit x0 (x1 = the unit after it)
__device__ __forceinline__ u32 l4_unit_min(u32 x0, u32 x1)
{
    const u32 a1 = __funnelshift_l(x1, x0, 4), a2 = __funnelshift_l(x1, x0, 8), a3 = __funnelshift_l(x1, x0, 12);
    const u32 a4 = __funnelshift_l(x1, x0, 16), a5 = __funnelshift_l(x1, x0, 20), a6 = __funnelshift_l(x1, x0, 24);
    const u32 a7 = __funnelshift_l(x1, x0, 28);
    return min(min(min(x0, a1), min(a2, a3)), min(min(a4, a5), min(a6, a7)));
}
// minimum over the 64 rotations of an oct (O) given the first unit of the next one
__device__ __forceinline__ u32 l4_oct_min(const Oct &O, u32 nx)
{
    const u32 m0 = min(l4_unit_min(O.lo.x, O.lo.y), l4_unit_min(O.lo.y, O.lo.z)), m1 = min(l4_unit_min(O.lo.z, O.lo.w), l4_unit_min(O.lo.w, O.hi.x));
    const u32 m2 = min(l4_unit_min(O.hi.x, O.hi.y), l4_unit_min(O.hi.y, O.hi.z)), m3 = min(l4_unit_min(O.hi.z, O.hi.w), l4_unit_min(O.hi.w, nx));
    return min(min(m0, m1), min(m2, m3));
}
// the same for the first `lim` rotations only (the partial oct of a record), and the bit mask of the rotations equal to `key`
__device__ __forceinline__ u32 l4_oct_min_masked(const Oct &O, u32 nx, u32 lim)
{
    const u32 x[9] = {O.lo.x, O.lo.y, O.lo.z, O.lo.w, O.hi.x, O.hi.y, O.hi.z, O.hi.w, nx};
    u32 m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 w = i ? __funnelshift_l(x[k + 1], x[k], 4 * i) : x[k];
            if ((u32)(8 * k + i) < lim) m = min(m, w);
        }
    }
    return m;
}
__device__ __forceinline__ u64 l4_oct_hits(const Oct &O, u32 nx, u32 key)
{
    const u32 x[9] = {O.lo.x, O.lo.y, O.lo.z, O.lo.w, O.hi.x, O.hi.y, O.hi.z, O.hi.w, nx};
    u32 lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 w = i ? __funnelshift_l(x[k + 1], x[k], 4 * i) : x[k];
            const u32 bit = (w == key) ? 1u : 0u;
            if (k < 4) lo |= bit << (8 * k + i); else hi |= bit << (8 * (k - 4) + i);
        }
    }
    return ((u64)hi << 32) | lo;
}
// the eight 8-symbol windows that start at symbol 8 u + (xs / 4) of the 128 symbols (LO, UP)
__device__ __forceinline__ void l4_windows(const Oct &LO, const Oct &UP, bool b2, bool b1, bool b0, u32 xs, u32 w[8])
{
    const u32 a0 = b2 ? LO.hi.x : LO.lo.x, a1 = b2 ? LO.hi.y : LO.lo.y, a2 = b2 ? LO.hi.z : LO.lo.z;
    const u32 a3 = b2 ? LO.hi.w : LO.lo.w, a4 = b2 ? UP.lo.x : LO.hi.x, a5 = b2 ? UP.lo.y : LO.hi.y;
    const u32 a6 = b2 ? UP.lo.z : LO.hi.z, a7 = b2 ? UP.lo.w : LO.hi.w, a8 = b2 ? UP.hi.x : UP.lo.x;
    const u32 a9 = b2 ? UP.hi.y : UP.lo.y, a10 = b2 ? UP.hi.z : UP.lo.z, a11 = b2 ? UP.hi.w : UP.lo.w;
    const u32 d0 = b1 ? a2 : a0, d1 = b1 ? a3 : a1, d2 = b1 ? a4 : a2, d3 = b1 ? a5 : a3, d4 = b1 ? a6 : a4;
    const u32 d5 = b1 ? a7 : a5, d6 = b1 ? a8 : a6, d7 = b1 ? a9 : a7, d8 = b1 ? a10 : a8, d9 = b1 ? a11 : a9;
    const u32 y0 = b0 ? d1 : d0, y1 = b0 ? d2 : d1, y2 = b0 ? d3 : d2, y3 = b0 ? d4 : d3, y4 = b0 ? d5 : d4;
    const u32 y5 = b0 ? d6 : d5, y6 = b0 ? d7 : d6, y7 = b0 ? d8 : d7, y8 = b0 ? d9 : d8;
    w[0] = __funnelshift_l(y1, y0, xs); w[1] = __funnelshift_l(y2, y1, xs); w[2] = __funnelshift_l(y3, y2, xs);
    w[3] = __funnelshift_l(y4, y3, xs); w[4] = __funnelshift_l(y5, y4, xs); w[5] = __funnelshift_l(y6, y5, xs);
    w[6] = __funnelshift_l(y7, y6, xs); w[7] = __funnelshift_l(y8, y7, xs);
}
// 16 letters of two windows: codes < 8 select straight in the 8-byte table (T0, T1) = "-ACG" "NT"; the lookups come out last
// symbol first, one PRMT turns each around
__device__ __forceinline__ uint4 l4_ascii16(u32 wa, u32 wb)
{
    const u32 T0 = 0x4743412du, T1 = 0x0000544eu;                  // '-','A','C','G' | 'N','T'
    uint4 v;
    v.x = __byte_perm(__byte_perm(T0, T1, wa >> 16), 0, 0x0123);
    v.y = __byte_perm(__byte_perm(T0, T1, wa), 0, 0x0123);
    v.z = __byte_perm(__byte_perm(T0, T1, wb >> 16), 0, 0x0123);
    v.w = __byte_perm(__byte_perm(T0, T1, wb), 0, 0x0123);
    return v;
}

template <int V>
__global__ void __launch_bounds__(32 * CK_L4_WARPS, 2) k_canon_l4(CanonArgs a, const u8 *p4, const u8 *lane_bits)
{
    extern __shared__ __align__(16) u32 smem[];
    constexpr bool want_hash = (V & CK_W2_HASH) != 0, want_out = (V & CK_W2_OUT) != 0;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 aux = (u32)__cvta_generic_to_shared(smem) + wid * CK_L4_WARP_BYTES;
    const u32 count = *a.count;
    a.list += a.count[16];
    const u32 st_w = aux + 128u * (lane >> 1);
    const u32 st_j = 4u * (lane & 1u) + ((lane >> 1) & 3u);
    const u32 st_w0 = st_w + 16u * (st_j & 7u), st_w1 = st_w + 16u * ((st_j + 1u) & 7u);
    const u32 st_w2 = st_w + 16u * ((st_j + 2u) & 7u), st_w3 = st_w + 16u * ((st_j + 3u) & 7u);
    const u32 st_r = aux + 128u * (lane >> 3) + 16u * (((lane & 3u) + 4u * ((lane >> 2) & 1u) + ((lane >> 3) & 3u)) & 7u);
    const u64 *sec = reinterpret_cast<const u64 *>(c_secret);

    for (;;) {
        u32 b0 = 0;
        if (lane == 0) b0 = atomicAdd(a.retry_counts + 30, 1u);
        b0 = __shfl_sync(CK_FULL, b0, 0) * 32u;
        if (b0 >= count) break;
        const u32 idx = b0 + lane;
        const bool have = idx < count;
        const u32 rec = have ? a.list[idx] : a.list[b0];
        const u64 off = a.offsets[rec];
        const u32 n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
        bool fast = have && lane_bits[rec] == 3 && n >= (want_hash ? 241u : 129u);       // 3: k_pack4 packed it
        const u32 nn = fast ? n : 241u;                            // lanes without a fast record walk a dummy geometry on valid memory
        const u8 *base_f = p4 + p4_byte(off, rec), *base_r = base_f + p4_strand_bytes(n);
        // the next batch's records -> L2 while this one is processed
        {
            const u32 nx = b0 + 32u * (gridDim.x * CK_L4_WARPS) + lane;    // a guess at a batch this warp may claim later
            if (nx < count) { const u32 r2 = a.list[nx]; prefetch_l2(p4 + p4_byte(a.offsets[r2], r2)); }
        }
        u8 *dst = want_out ? a.out + out_byte(off, rec) : nullptr;

        // ---- scan: full octs j < J1 (all 64 rotations below n), both strand arrays
        const u32 J1 = nn >> 6, lim = nn & 63u;
        const u32 J1max = __reduce_max_sync(CK_FULL, fast ? J1 : 0u);
        u32 m1k = 0xffffffffu, m1t = 0, m2k = 0xffffffffu;
        {
            // three octs per strand rotate through the roles (current, next, in flight) by unrolling, not by register moves
            Oct FA = ldg256_here(base_f), FB = ldg256_here(base_f + 32), RA = ldg256_here(base_r), RB = ldg256_here(base_r + 32);
            Oct FC, RC;
#define CK_L4_SCAN(A0, A1, A2, B0, B1, B2)                                                                           \
            {                                                                                                       \
                const u32 jn = min(j + 2, J1 + 1);                                                                  \
                A2 = ldg256_here(base_f + 32 * jn); B2 = ldg256_here(base_r + 32 * jn);                             \
                const u32 kf = l4_oct_min(A0, A1.lo.x), kr = l4_oct_min(B0, B1.lo.x);                               \
                if (j < J1) { seg_track(m1k, m1t, m2k, kf, 2 * j); seg_track(m1k, m1t, m2k, kr, 2 * j + 1); }       \
                if (++j >= J1max) break;                                                                            \
            }
            if (J1max) {
                u32 j = 0;
#pragma unroll 1
                for (;;) {
                    CK_L4_SCAN(FA, FB, FC, RA, RB, RC)
                    CK_L4_SCAN(FB, FC, FA, RB, RC, RA)
                    CK_L4_SCAN(FC, FA, FB, RC, RA, RB)
                }
            }
#undef CK_L4_SCAN
        }
        if (lim) {      // the partial oct: rotations 64 J1 .. n - 1
            const Oct F0 = ldg256_here(base_f + 32 * J1), R0 = ldg256_here(base_r + 32 * J1);
            const u32 fx = ldg32(base_f + 32 * J1 + 32), rx = ldg32(base_r + 32 * J1 + 32);
            seg_track(m1k, m1t, m2k, l4_oct_min_masked(F0, fx, lim), 2 * J1);
            seg_track(m1k, m1t, m2k, l4_oct_min_masked(R0, rx, lim), 2 * J1 + 1);
        }
        // ---- locate
        u32 pos = 0, strand = 0;
        {
            const u32 t = fast ? m1t >> 1 : 0u;
            strand = fast ? (m1t & 1u) : 0u;
            const u8 *bs = strand ? base_r : base_f;
            const Oct O = ldg256_here(bs + 32 * t);
            const u32 nx = ldg32(bs + 32 * t + 32);
            u64 hits = l4_oct_hits(O, nx, m1k);
            const u32 left = nn - 64u * t;                         // rotations of this oct below n
            if (left < 64u) hits &= (1ull << left) - 1ull;
            pos = 64u * t + (u32)__ffsll((long long)hits) - 1u;
            if (m1k == m2k || __popcll(hits) != 1) fast = false;
            if (!fast) { pos = 0; strand = 0; }
        }
        const u8 *bs = strand ? base_r : base_f;
        u64 h = 0;
        if (want_out || want_hash) {
            const u32 nchunks = fast ? (nn + 15) >> 4 : 0u;
            const u32 nfull = fast ? (nn - 1) >> 6 : 0u;
            const u32 rounds = __reduce_max_sync(CK_FULL, (nchunks + 3) >> 2);
            // round s emits the 64 symbols that start at circular position cpos = (pos + 64 s) mod n: a linear window of the
            // strand array (the continuation past n covers a round that starts below n); the round after the one that reaches n
            // starts in oct 0 again, with another window phase
            u32 cpos = pos;
            const u32 olim = fast ? 32u * ((nn + 135u) >> 6) : 32u;         // last oct of the strand array
            u32 onext = ((pos >> 6) << 5) + 64u;
            Oct R0 = ldg256_here(bs + ((pos >> 6) << 5)), R1 = ldg256_here(bs + ((pos >> 6) << 5) + 32);
            u64 acc0 = CK_P32_3, acc1 = CK_P64_1, acc2 = CK_P64_2, acc3 = CK_P64_3;
            u64 acc4 = CK_P64_4, acc5 = CK_P32_2, acc6 = CK_P64_5, acc7 = CK_P32_1;
            u32 og[4] = {0, 0, 0, 0}; int orem[4] = {0, 0, 0, 0};
            if (want_out) {
                const u32 gr = (u32)((dst - a.out) >> 4);
#pragma unroll
                for (u32 i = 0; i < 4; i++) {
                    og[i] = __shfl_sync(CK_FULL, gr, 8 * i + (lane >> 2)) + (lane & 3u);
                    orem[i] = (int)__shfl_sync(CK_FULL, nchunks, 8 * i + (lane >> 2)) - (int)(lane & 3u);
                }
            }
#define CK_T2_SCR(A, I) do { A ^= A >> 47; A ^= sec[16 + I]; A *= CK_P32_1; } while (0)
#define CK_L4_ROUND(LO, UP)                                                                                          \
            {                                                                                                       \
                u32 w[8];                                                                                           \
                l4_windows(LO, UP, (cpos & 32u) != 0, (cpos & 16u) != 0, (cpos & 8u) != 0, 4u * (cpos & 7u), w);    \
                if (s + 1 < rounds) {                                                                               \
                    const u32 cn = cpos + 64u;                                                                      \
                    const bool wrap = fast && cn >= nn;                                                             \
                    if (wrap) UP = ldg256_here(bs);                 /* the next round's lower oct is oct 0 */       \
                    LO = ldg256_here(bs + (wrap ? 32u : min(onext, olim)));                                         \
                    onext = wrap ? 64u : onext + 32u;                                                               \
                    cpos = wrap ? cn - nn : cn;                                                                     \
                }                                                                                                   \
                uint4 v[4];                                                                                         \
                v[0] = l4_ascii16(w[0], w[1]); v[1] = l4_ascii16(w[2], w[3]);                                       \
                v[2] = l4_ascii16(w[4], w[5]); v[3] = l4_ascii16(w[6], w[7]);                                       \
                if (want_hash && s < nfull) {                                                                       \
                    const u32 ks = s & 15u;                                                                         \
                    t2_acc16(acc0, acc1, v[0], sec[ks + 0], sec[ks + 1]);                                           \
                    t2_acc16(acc2, acc3, v[1], sec[ks + 2], sec[ks + 3]);                                           \
                    t2_acc16(acc4, acc5, v[2], sec[ks + 4], sec[ks + 5]);                                           \
                    t2_acc16(acc6, acc7, v[3], sec[ks + 6], sec[ks + 7]);                                           \
                    if (ks == 15u) {                                                                                \
                        CK_T2_SCR(acc0, 0); CK_T2_SCR(acc1, 1); CK_T2_SCR(acc2, 2); CK_T2_SCR(acc3, 3);             \
                        CK_T2_SCR(acc4, 4); CK_T2_SCR(acc5, 5); CK_T2_SCR(acc6, 6); CK_T2_SCR(acc7, 7);             \
                    }                                                                                               \
                }                                                                                                   \
                if (want_out) {                                                                                     \
                    sts128(st_w0, v[0]); sts128(st_w1, v[1]); sts128(st_w2, v[2]); sts128(st_w3, v[3]);             \
                    __syncwarp();                                                                                   \
                    uint4 g[4];                                                                                     \
                    _Pragma("unroll") for (u32 i = 0; i < 4; i++) g[i] = lds128(st_r + 512u * i);                   \
                    _Pragma("unroll") for (u32 i = 0; i < 4; i++) {                                                 \
                        if (orem[i] > 0) reinterpret_cast<uint4 *>(a.out)[og[i]] = g[i];                            \
                        og[i] += 4; orem[i] -= 4;                                                                   \
                    }                                                                                               \
                    __syncwarp();                                                                                   \
                }                                                                                                   \
                if (++s >= rounds) break;                                                                           \
            }
            if (rounds) {
                u32 s = 0;
#pragma unroll 1
                for (;;) {
                    CK_L4_ROUND(R0, R1)
                    CK_L4_ROUND(R1, R0)
                }
            }
#undef CK_L4_ROUND
            if (want_hash) {
                // last stripe: canonical symbols [n - 64, n) = the strand array at (pos + n - 64) mod n, linear with the continuation
                u32 q = pos + nn - 64u;
                if (q >= nn) q -= nn;
                const Oct LA = ldg256_here(bs + ((q >> 6) << 5)), LB = ldg256_here(bs + ((q >> 6) << 5) + 32);
                u32 w[8];
                l4_windows(LA, LB, (q & 32u) != 0, (q & 16u) != 0, (q & 8u) != 0, 4u * (q & 7u), w);
                t2_acc16(acc0, acc1, l4_ascii16(w[0], w[1]), c_lastsec[0], c_lastsec[1]);
                t2_acc16(acc2, acc3, l4_ascii16(w[2], w[3]), c_lastsec[2], c_lastsec[3]);
                t2_acc16(acc4, acc5, l4_ascii16(w[4], w[5]), c_lastsec[4], c_lastsec[5]);
                t2_acc16(acc6, acc7, l4_ascii16(w[6], w[7]), c_lastsec[6], c_lastsec[7]);
                u64 r = (u64)nn * CK_P64_1;
                r += mul128_fold64(acc0 ^ c_mergesec[0], acc1 ^ c_mergesec[1]);
                r += mul128_fold64(acc2 ^ c_mergesec[2], acc3 ^ c_mergesec[3]);
                r += mul128_fold64(acc4 ^ c_mergesec[4], acc5 ^ c_mergesec[5]);
                r += mul128_fold64(acc6 ^ c_mergesec[6], acc7 ^ c_mergesec[7]);
                h = xxh3_avalanche(r);
            }
#undef CK_T2_SCR
        }
        if (fast) {
            a.out_start[rec] = strand ? (n - 1 - pos) : pos;
            a.out_strand[rec] = (u8)strand;
            if (want_hash) a.out_hash[rec] = h;
        } else if (have) {
            // the generic warp kernel takes it (k_canon_warp<4> over this class's retry list)
            const u32 k = atomicAdd(a.retry_counts + CLS_W4, 1u);
            a.retry[a.retry_counts[16 + CLS_W4] + k] = rec;
        }
        __syncwarp();
    }
}

}  // namespace ck
