// ck_lane2.cuh -- lane-per-record LMSR + canonical form + XXH3-64 for 2-bit records (the hot kernel of configs
// 1, 2 and 5 of BASELINE.json).
//
// The warp-per-record kernel (ck_warp2.cuh) pays ~300 warp instructions of per-record fixed cost (address
// arithmetic, warp collectives, half-empty rounds) for a 325-base record.  Here a warp takes 32 records at once and
// every lane walks ITS OWN record, so a warp instruction does useful work for 32 records and no collective sits on
// the per-base path:
//   * stage     : the warp copies the packed words of its 32 records (coalesced 64-bit loads) into a shared-memory
//                 tile, one row per lane; the row stride is 4 (mod 32) units so that 128-bit row reads of the 8 lanes
//                 of a quarter-warp hit 32 different banks; each lane appends its own circular extension;
//   * scan      : 8-mer keys as 16-bit halves (14 funnel shifts + 8 VIMNMX3.U16x2 per 32 rotations, as in the warp
//                 kernel).  The reverse-complement strand is never staged: the 48-base block (x0,x1,x2) a step looks
//                 at is reverse-complemented in registers (BREV + LOP3 per unit), which yields the reverse strand's
//                 keys for forward starts 32t+9 .. 32t+40.  Each lane keeps the two smallest (key, step, strand)
//                 triples, so "the minimal 8-mer is unique" is known exactly at the end.  Only the last step of a
//                 record can see positions twice (the extension repeats the head); its units are padded in registers
//                 (T past base n+7 for the forward keys, A past base n+16 for the reverse keys) so that those
//                 duplicates can never win;
//   * locate    : the winning step is replayed once; an XOR / VIMNMX / shift-add chain turns "which of the 32 keys
//                 equal the minimum" into a bit mask; exactly one valid bit => that rotation is the canonical one;
//   * ASCII+hash: each lane streams its canonical form in 16-byte chunks (the reverse strand costs nothing extra: a
//                 per-lane letter table "TGCA" and per-lane PRMT selectors do the complement and the reversal), four
//                 chunks = one XXH3 stripe per round, eight 64-bit accumulators in registers, no cross-lane traffic;
//   * write     : a round's 64 bytes per lane go through a padded shared-memory stage and leave as 128-bit stores
//                 in which four consecutive lanes cover 64 contiguous bytes of one record;
//   * everything a lane cannot finish alone (equal minimal 8-mers, n < 128, the short XXH3 forms of n <= 240) is
//                 done afterwards one record at a time by the whole warp with the generic record code.
#pragma once
#include "ck_warp2.cuh"

namespace ck {

__constant__ u64 c_lastsec[8];      // LE u64 of the secret at byte offsets 121 + 8 i (XXH3 last stripe)
__constant__ u64 c_mergesec[8];     // LE u64 of the secret at byte offsets 11 + 8 i  (XXH3 merge)

__device__ __forceinline__ uint4 lds128(u32 a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(u32 a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// per-warp shared memory: two tiles (32 rows of ROWU units each; one is computed on while cp.async fills the other),
// a 2560-byte area that holds the staging descriptors first and the output stage later, the (offset, end) pairs of the
// two batches in flight, and (work-list launches) the record indices of three batches
#define CK_T2_AUX_BYTES 2560u
template <int ROWU> __host__ __device__ constexpr u32 t2_tile_bytes() { return 32u * ROWU * 4u; }
template <int ROWU> __host__ __device__ constexpr u32 t2_tiles() { return ROWU <= 36 ? 2u : 1u; }   // long rows: one tile, batch latency is amortised
template <int ROWU> __host__ __device__ constexpr u32 t2_warp_bytes(bool list)
{
    return t2_tiles<ROWU>() * t2_tile_bytes<ROWU>() + CK_T2_AUX_BYTES + 1024u + (list ? 384u : 0u);
}
template <int ROWU> __host__ __device__ constexpr u32 t2_max_n() { return 16u * (ROWU - 4); }
// warps per CTA and CTAs per SM: what 227 KB of shared memory hold
template <int ROWU> __host__ __device__ constexpr u32 t2_warps(bool list)
{
    return ROWU <= 36 ? (list ? 5u : 6u) : ROWU <= 132 ? 5u : ROWU <= 260 ? 3u : 1u;
}
template <int ROWU> __host__ __device__ constexpr u32 t2_ctas() { return ROWU <= 36 ? 3u : ROWU <= 260 ? 2u : 3u; }

__device__ __forceinline__ void cp_async8(u32 dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 16 canonical letters from a 16-base window; T = letter table, (sa, sb) = interleave selectors.  Forward strand:
// ("ACGT", 0x2637, 0x0415).  Reverse strand: the window is rotated by 16 bits first and ("TGCA", 0x5140, 0x7362)
// complement and reverse it.
__device__ __forceinline__ uint4 t2_ascii16(u32 w, u32 T, u32 sa, u32 sb)
{
    const u32 E = w & 0x33333333u, O = (w >> 2) & 0x33333333u;
    const u32 pe_lo = __byte_perm(T, 0, E), pe_hi = __byte_perm(T, 0, E >> 16);
    const u32 po_lo = __byte_perm(T, 0, O), po_hi = __byte_perm(T, 0, O >> 16);
    uint4 v;
    v.x = __byte_perm(pe_hi, po_hi, sa);
    v.y = __byte_perm(pe_hi, po_hi, sb);
    v.z = __byte_perm(pe_lo, po_lo, sa);
    v.w = __byte_perm(pe_lo, po_lo, sb);
    return v;
}

// the two smallest (key << 16 | tag) values seen so far
__device__ __forceinline__ void t2_track(u32 &m1, u32 &m2, u32 v)
{
    m2 = min(m2, max(m1, v));
    m1 = min(m1, v);
}
// u16x2 step minimum -> (min of both halves) << 16
__device__ __forceinline__ u32 t2_key_hi(u32 m)
{
    const u32 sw = __byte_perm(m, 0, 0x1032);
    return __vminu2(m, sw) & 0xffff0000u;
}

// one XXH3 stripe piece: 16 input bytes (v) against secret words (k0, k1); accumulators (a0, a1) = acc[2k], acc[2k+1]
__device__ __forceinline__ void t2_acc16(u64 &a0, u64 &a1, uint4 v, u64 k0, u64 k1)
{
    const u32 x0 = v.x ^ (u32)k0, x1 = v.y ^ (u32)(k0 >> 32), y0 = v.z ^ (u32)k1, y1 = v.w ^ (u32)(k1 >> 32);
    a0 += (u64)x0 * (u64)x1 + (((u64)v.w << 32) | v.z);
    a1 += (u64)y0 * (u64)y1 + (((u64)v.y << 32) | v.x);
}

template <int ROWU, int V>
__global__ void __launch_bounds__(32 * t2_warps<ROWU>((V & CK_W2_LIST) != 0), t2_ctas<ROWU>()) k_canon_t2(CanonArgs a)
{
    extern __shared__ __align__(16) u32 smem[];
    constexpr bool want_hash = (V & CK_W2_HASH) != 0, want_out = (V & CK_W2_OUT) != 0, use_list = (V & CK_W2_LIST) != 0;
    constexpr u32 TB = t2_tile_bytes<ROWU>(), WB = t2_warp_bytes<ROWU>(use_list), NMAX = t2_max_n<ROWU>();
    constexpr bool BLOCKS = NMAX > 1024;                           // XXH3 block scrambles can occur
    constexpr bool DB = t2_tiles<ROWU>() == 2;                     // double-buffered tiles
    const u32 lane = lane_id(), wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const u32 wsm = (u32)__cvta_generic_to_shared(smem) + wid * WB;  // this warp's tiles
    const u32 aux = wsm + t2_tiles<ROWU>() * TB;                   // descriptors / output stage
    const u32 offs = aux + CK_T2_AUX_BYTES + 16u * lane;           // + 512 * slot: (offset, end) of this lane's record
    const u32 recs = aux + CK_T2_AUX_BYTES + 1024u + 4u * lane;    // + 128 * (batch % 3): record index (work lists)
    const u32 gw = blockIdx.x * wpb + wid, nw = gridDim.x * wpb;
    u32 *scr = a.scratch + (size_t)gw * a.scratch_stride;
    const u32 count = use_list ? *a.count : a.n_direct;
    if (use_list) a.list += a.count[16];
    const u32 bstride = nw * 32u;

    // request the (offset, end) pair of record `rec` into slot `slot`
    auto fetch_offsets = [&](u32 rec, u32 slot) {
        cp_async8(offs + 512u * slot, a.offsets + rec);
        cp_async8(offs + 512u * slot + 8u, a.offsets + rec + 1);
    };
    // request the packed words of a batch into tile `tile`: descriptors, then the warp copies record after record
    auto fetch_tile = [&](u32 tile, u32 rec, u64 off, u32 n, bool want) {
        const u64 sp = reinterpret_cast<u64>(a.packed2 + p2_word(off, rec));
        sts128(aux + 16u * lane, make_uint4((u32)sp, (u32)(sp >> 32), want ? (n + 31) >> 5 : 0u, 0u));
        __syncwarp();
        const u32 tw = wsm + tile * TB;
        if (ROWU <= 36) {                                          // <= 16 words per record: two records per pass
            const u32 hlf = lane >> 4, k = lane & 15u;
#pragma unroll 4
            for (u32 i = 0; i < 16; i++) {
                const u32 r = 2 * i + hlf;
                const uint4 d = lds128(aux + 16u * r);
                if (k < d.z) cp_async8(tw + 4u * ROWU * r + 8u * k, reinterpret_cast<const uint2 *>(((u64)d.y << 32) | d.x) + k);
            }
        } else {
#pragma unroll 1
            for (u32 r = 0; r < 32; r++) {
                const uint4 d = lds128(aux + 16u * r);
                const uint2 *p = reinterpret_cast<const uint2 *>(((u64)d.y << 32) | d.x);
#pragma unroll 1
                for (u32 k = lane; k < d.z; k += 32) cp_async8(tw + 4u * ROWU * r + 8u * k, p + k);
            }
        }
        __syncwarp();
    };
    auto is_fast = [&](bool have, u32 n) {
        const bool in_class = have && (use_list || (n >= a.min_n && n <= a.max_n));
        return in_class && n >= (want_hash ? 241u : 128u) && n <= NMAX;
    };

    // ---- prologue: offsets of batches 0 and 1, the tile of batch 0, (work lists) record indices of batches 0..2
    u32 rq = 0;                                                    // work lists: record index two batches ahead
    {
        const u32 i0 = gw * 32u + lane, i1 = i0 + bstride, i2 = i1 + bstride;
        u32 r0 = i0, r1 = i1;
        if (use_list) {
            r0 = i0 < count ? a.list[i0] : 0u; r1 = i1 < count ? a.list[i1] : 0u; rq = i2 < count ? a.list[i2] : 0u;
            sts32(recs, r0); sts32(recs + 128, r1);
        }
        if (i0 < count) fetch_offsets(r0, 0);
        if (i1 < count) fetch_offsets(r1, 1);
        cp_async_wait_all();
        __syncwarp();
        if (DB && gw * 32u < count) {
            const uint4 oe = lds128(offs);
            const u64 off = ((u64)oe.y << 32) | oe.x;
            const u32 n = i0 < count ? oe.z - oe.x : 0u;
            fetch_tile(0, r0, off, n, is_fast(i0 < count, n));
        }
    }
    u32 kb = 0;                                                    // batch counter of this warp
    for (u32 b = gw * 32u; b < count; b += bstride, kb++) {
        const u32 sl = kb & 1u, cur = DB ? sl : 0u;                // offsets slot / tile of this batch
        const u32 idx = b + lane;
        const bool have = idx < count;
        cp_async_wait_all();
        __syncwarp();                                              // the offsets of batches kb, kb + 1 (and tile[cur]) have landed
        u32 rec = idx; u64 off = 0; u32 n = 0;
        {
            const uint4 oe = lds128(offs + 512u * sl);
            if (use_list) rec = lds32(recs + 128u * (kb % 3u));
            if (have) { off = ((u64)oe.y << 32) | oe.x; n = oe.z - oe.x; } else rec = 0;
        }
        {   // next batch -> the other tile (or this batch -> the only tile); offsets of the batch after the next -> the slot just read
            const u32 idx1 = idx + bstride, idx2 = idx1 + bstride;
            if (!DB) fetch_tile(0, rec, off, n, is_fast(have, n));
            else if (b + bstride < count) {
                const uint4 oe = lds128(offs + 512u * (sl ^ 1u));
                u32 rec1 = idx1;
                if (use_list) rec1 = lds32(recs + 128u * ((kb + 1u) % 3u));
                const bool have1 = idx1 < count;
                const u32 n1 = have1 ? oe.z - oe.x : 0u;
                fetch_tile(cur ^ 1u, have1 ? rec1 : 0u, have1 ? (((u64)oe.y << 32) | oe.x) : 0ull, n1, is_fast(have1, n1));
            }
            if (idx2 < count) fetch_offsets(use_list ? rq : idx2, sl);
            if (use_list) {
                sts32(recs + 128u * ((kb + 2u) % 3u), rq);
                const u32 idx3 = idx2 + bstride;
                rq = idx3 < count ? a.list[idx3] : 0u;
            }
            if (!DB) { cp_async_wait_all(); __syncwarp(); }
        }
        const u32 tb = wsm + cur * TB + 4u * ROWU * lane;          // this lane's row
        u32 *Xf = smem + (size_t)wid * (WB / 4) + cur * (TB / 4), *Xr = Xf + a.smem_units;   // generic path: linear strands over the tile
        const bool in_class = have && (use_list || (n >= a.min_n && n <= a.max_n));
        // lane-private fast path: 128 <= n <= NMAX (and the long XXH3 form when a hash is wanted)
        bool fast = is_fast(have, n);
        u8 *dst = want_out ? a.out + 16ull * ((off >> 4) + rec) : nullptr;
        u32 os = 0;                                                // (start << 1) | strand, strand's own coordinates
        u64 h = 0;
        const u32 nn = fast ? n : 128u;                            // lanes without a fast record walk a dummy geometry

        // ---- circular extension of this lane's row: units jn .. jn + 2 (real bases)
        const u32 jn = nn >> 4, rem = nn & 15u;
        {
            const u32 u0 = lds32(tb), u1 = lds32(tb + 4), u2 = lds32(tb + 8), uj = lds32(tb + 4 * jn);
            const u32 sh = 32u - 2u * rem;
            const u32 e0 = (uj & ~(0xffffffffu >> (2 * rem))) | (u0 >> (2 * rem));
            const u32 e1 = __funnelshift_lc(u1, u0, sh), e2 = __funnelshift_lc(u2, u1, sh);
            sts32(tb + 4 * jn, e0); sts32(tb + 4 * jn + 4, e1); sts32(tb + 4 * jn + 8, e2);
        }
        // ---- scan: the two smallest (8-mer key, step, strand) over all 2n rotations
        const u32 S1 = ((nn + 31) >> 5) - 1;                       // full steps; step S1 is the padded last one
        const u32 S1max = __reduce_max_sync(CK_FULL, fast ? S1 : 0u);
        u32 m1 = 0xffffffffu, m2 = 0xffffffffu;
        {
            uint4 Q = lds128(tb);
            u32 rqx = w2_revcomp(Q.x);
#pragma unroll 1
            for (u32 i = 0; 2 * i < S1max; i++) {
                const uint4 Qn = lds128(tb + 16 * i + 16);
                const u32 rqy = w2_revcomp(Q.y), rqz = w2_revcomp(Q.z), rqw = w2_revcomp(Q.w), rnx = w2_revcomp(Qn.x);
                {
                    const u32 t = 2 * i;
                    const u32 kf = t2_key_hi(w2_step_min16(Q.x, Q.y, Q.z)), kr = t2_key_hi(w2_step_min16(rqz, rqy, rqx));
                    const u32 tag = (t < S1) ? 2 * t : 0xffff0000u;
                    t2_track(m1, m2, kf | tag);
                    t2_track(m1, m2, kr | (tag + 1));
                }
                {
                    const u32 t = 2 * i + 1;
                    const u32 kf = t2_key_hi(w2_step_min16(Q.z, Q.w, Qn.x)), kr = t2_key_hi(w2_step_min16(rnx, rqw, rqz));
                    const u32 tag = (t < S1) ? 2 * t : 0xffff0000u;
                    t2_track(m1, m2, kf | tag);
                    t2_track(m1, m2, kr | (tag + 1));
                }
                Q = Qn; rqx = rnx;
            }
        }
        {   // last step of this lane's record: positions 32 S1 .. n - 1 are new, the rest repeats the head
            const uint2 x01 = lds64(tb + 8 * S1);
            const u32 x2 = lds32(tb + 8 * S1 + 8);
            const int dT = (int)nn + 7 - 32 * (int)S1;             // first forward-padded base, relative to unit 2 S1
            const int dA = dT + 9;                                 // first reverse-padded base
#define CK_T2_PAD(d) __funnelshift_rc(0xffffffffu, 0u, 2u * (u32)min(max((d), 0), 16))
            const u32 pT0 = CK_T2_PAD(dT), pT1 = CK_T2_PAD(dT - 16), pT2 = CK_T2_PAD(dT - 32);
            const u32 pA0 = CK_T2_PAD(dA), pA1 = CK_T2_PAD(dA - 16), pA2 = CK_T2_PAD(dA - 32);
#undef CK_T2_PAD
            const u32 kf = t2_key_hi(w2_step_min16(x01.x | pT0, x01.y | pT1, x2 | pT2));
            const u32 kr = t2_key_hi(w2_step_min16(w2_revcomp(x2 & ~pA2), w2_revcomp(x01.y & ~pA1), w2_revcomp(x01.x & ~pA0)));
            t2_track(m1, m2, kf | (2 * S1));
            t2_track(m1, m2, kr | (2 * S1 + 1));
        }
        // ---- locate: replay the winning step, find the rotation that carries the minimal 8-mer
        {
            const u32 t = (m1 & 0xffffu) >> 1, strand = m1 & 1u;
            const uint2 x01 = lds64(tb + 8 * t);
            const u32 x2 = lds32(tb + 8 * t + 8);
            const u32 y0 = strand ? w2_revcomp(x2) : x01.x, y1 = strand ? w2_revcomp(x01.y) : x01.y;
            const u32 y2 = strand ? w2_revcomp(x01.x) : x2;
            const u32 bb = (m1 >> 16) * 0x10001u;
            u32 nm = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const u32 wa = i ? __funnelshift_l(y1, y0, 2 * i) : y0, wb = i ? __funnelshift_l(y2, y1, 2 * i) : y1;
                nm += __vminu2(wa ^ bb, 0x00010001u) << i;         // bits 16 + i / i: key i / key 8 + i differs
                nm += __vminu2(wb ^ bb, 0x00010001u) << (i + 8);   // bits 24 + i / 8 + i: key 16 + i / key 24 + i differs
            }
            const u32 match = ~__byte_perm(nm, 0, 0x1302);         // bit s: key s of the step equals the minimum
            // forward: key s is rotation 32 t + s, valid below n.  reverse: key s is forward start 32 t + 40 - s,
            // valid up to n + 8
            const int lim = (int)nn - 32 * (int)t;
            const u32 valid = strand ? (lim >= 32 ? 0xffffffffu : 0xffffffffu << (32 - lim))
                                     : (lim >= 32 ? 0xffffffffu : (1u << lim) - 1u);
            const u32 hits = match & valid;
            const u32 s = __ffs(hits) - 1;
            int st = strand ? (int)nn - 48 - 32 * (int)t + (int)s : 32 * (int)t + (int)s;
            if (st < 0) st += (int)nn;
            os = ((u32)st << 1) | strand;
            if (((m1 ^ m2) < 0x10000u) || __popc(hits) != 1) fast = false;      // equal minima: the duel path decides
            if (!fast) os = 0;                                     // keeps the dummy walk below inside the row
        }
        // ---- canonical ASCII (+ XXH3-64), lane-private; one stripe (4 chunks of 16 bytes) per round
        if (want_out || want_hash) {
            const u32 strand = os & 1u;
            const u32 T = strand ? 0x41434754u : 0x54474341u;      // "TGCA" / "ACGT"
            const u32 sa = strand ? 0x5140u : 0x2637u, sb = strand ? 0x7362u : 0x0415u, rot = strand ? 16u : 0u;
            const int step = strand ? -16 : 16, nstep = strand ? -(int)nn : (int)nn;
            // forward position of canonical chunk 0: the rotation start, or the mirror of reverse position start
            int p = strand ? (int)nn - 16 - (int)(os >> 1) : (int)(os >> 1);
            if (p < 0) p += (int)nn;
            const int p0 = p;
            const u32 nchunks = fast ? (nn + 15) >> 4 : 0u;
            const u32 nfull = fast ? (nn - 1) >> 6 : 0u;           // stripes the stripe loop hashes
            const u32 rounds = __reduce_max_sync(CK_FULL, (nchunks + 3) >> 2);
            u64 acc0 = CK_P32_3, acc1 = CK_P64_1, acc2 = CK_P64_2, acc3 = CK_P64_3;
            u64 acc4 = CK_P64_4, acc5 = CK_P32_2, acc6 = CK_P64_5, acc7 = CK_P32_1;
            // the four records whose bytes this lane carries out of the stage: records 8 i + (lane >> 2), piece lane & 3
            u8 *od0 = nullptr, *od1 = nullptr, *od2 = nullptr, *od3 = nullptr;
            u32 oc0 = 0, oc1 = 0, oc2 = 0, oc3 = 0;
            if (want_out) {
                const u64 dp = reinterpret_cast<u64>(dst), pc = 16u * (lane & 3u);
                const u32 q = lane >> 2;
                od0 = reinterpret_cast<u8 *>(__shfl_sync(CK_FULL, dp, q) + pc);      oc0 = __shfl_sync(CK_FULL, nchunks, q);
                od1 = reinterpret_cast<u8 *>(__shfl_sync(CK_FULL, dp, q + 8) + pc);  oc1 = __shfl_sync(CK_FULL, nchunks, q + 8);
                od2 = reinterpret_cast<u8 *>(__shfl_sync(CK_FULL, dp, q + 16) + pc); oc2 = __shfl_sync(CK_FULL, nchunks, q + 16);
                od3 = reinterpret_cast<u8 *>(__shfl_sync(CK_FULL, dp, q + 24) + pc); oc3 = __shfl_sync(CK_FULL, nchunks, q + 24);
            }
            const u32 ost = aux + 80u * lane;                      // this lane's 64 bytes in the stage (stride 80: no conflicts)
            const u32 ord = aux + 80u * (lane >> 2) + 16u * (lane & 3u);
            const u64 *sec = reinterpret_cast<const u64 *>(c_secret);
#pragma unroll 1
            for (u32 s = 0; s < rounds; s++) {
                uint4 v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const u32 ad = tb + (((u32)p >> 4) << 2);
                    u32 w = __funnelshift_l(lds32(ad + 4), lds32(ad), 2u * (u32)p);
                    w = __funnelshift_l(w, w, rot);
                    v[k] = t2_ascii16(w, T, sa, sb);
                    p += step;
                    if ((u32)p >= nn) p -= nstep;
                }
                if (want_hash && s < nfull) {
                    const u32 ks = BLOCKS ? (s & 15u) : s;
                    t2_acc16(acc0, acc1, v[0], sec[ks + 0], sec[ks + 1]);
                    t2_acc16(acc2, acc3, v[1], sec[ks + 2], sec[ks + 3]);
                    t2_acc16(acc4, acc5, v[2], sec[ks + 4], sec[ks + 5]);
                    t2_acc16(acc6, acc7, v[3], sec[ks + 6], sec[ks + 7]);
                    if (BLOCKS && ks == 15u) {                     // a 1024-byte block is complete (s + 1 <= nfull: more input follows)
#define CK_T2_SCR(A, I) do { A ^= A >> 47; A ^= sec[16 + I]; A *= CK_P32_1; } while (0)
                        CK_T2_SCR(acc0, 0); CK_T2_SCR(acc1, 1); CK_T2_SCR(acc2, 2); CK_T2_SCR(acc3, 3);
                        CK_T2_SCR(acc4, 4); CK_T2_SCR(acc5, 5); CK_T2_SCR(acc6, 6); CK_T2_SCR(acc7, 7);
#undef CK_T2_SCR
                    }
                }
                if (want_out) {
                    sts128(ost, v[0]); sts128(ost + 16, v[1]); sts128(ost + 32, v[2]); sts128(ost + 48, v[3]);
                    __syncwarp();
                    const u32 c = 4 * s + (lane & 3u);
                    const uint4 g0 = lds128(ord), g1 = lds128(ord + 640), g2 = lds128(ord + 1280), g3 = lds128(ord + 1920);
                    if (c < oc0) *reinterpret_cast<uint4 *>(od0 + 64 * (size_t)s) = g0;
                    if (c < oc1) *reinterpret_cast<uint4 *>(od1 + 64 * (size_t)s) = g1;
                    if (c < oc2) *reinterpret_cast<uint4 *>(od2 + 64 * (size_t)s) = g2;
                    if (c < oc3) *reinterpret_cast<uint4 *>(od3 + 64 * (size_t)s) = g3;
                    __syncwarp();
                }
            }
            if (want_hash) {
                // last stripe: canonical bytes [n - 64, n) = four chunks that start 64 bases before chunk 0
                p = p0 - 4 * step;
                if ((u32)p >= nn) p += nstep;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const u32 ad = tb + (((u32)p >> 4) << 2);
                    u32 w = __funnelshift_l(lds32(ad + 4), lds32(ad), 2u * (u32)p);
                    w = __funnelshift_l(w, w, rot);
                    const uint4 vv = t2_ascii16(w, T, sa, sb);
                    if (k == 0) t2_acc16(acc0, acc1, vv, c_lastsec[0], c_lastsec[1]);
                    if (k == 1) t2_acc16(acc2, acc3, vv, c_lastsec[2], c_lastsec[3]);
                    if (k == 2) t2_acc16(acc4, acc5, vv, c_lastsec[4], c_lastsec[5]);
                    if (k == 3) t2_acc16(acc6, acc7, vv, c_lastsec[6], c_lastsec[7]);
                    p += step;
                    if ((u32)p >= nn) p -= nstep;
                }
                u64 r = (u64)nn * CK_P64_1;
                r += mul128_fold64(acc0 ^ c_mergesec[0], acc1 ^ c_mergesec[1]);
                r += mul128_fold64(acc2 ^ c_mergesec[2], acc3 ^ c_mergesec[3]);
                r += mul128_fold64(acc4 ^ c_mergesec[4], acc5 ^ c_mergesec[5]);
                r += mul128_fold64(acc6 ^ c_mergesec[6], acc7 ^ c_mergesec[7]);
                h = xxh3_avalanche(r);
            }
        }
        if (fast) {
            const u32 start = os >> 1, strand = os & 1u;
            a.out_start[rec] = strand ? (n - 1 - start) : start;
            a.out_strand[rec] = (u8)strand;
            if (want_hash) a.out_hash[rec] = h;
        }
        // ---- everything the lanes could not finish alone: one record at a time, whole warp, generic code (the tile
        //      is free now and holds the linear strands)
        {
            u32 slow = __ballot_sync(CK_FULL, in_class && !fast);
            while (slow) {
                const u32 L = __ffs(slow) - 1;
                slow &= slow - 1;
                const u32 rec_l = __shfl_sync(CK_FULL, rec, L), n_l = __shfl_sync(CK_FULL, n, L);
                const u64 off_l = __shfl_sync(CK_FULL, off, L);
                u8 *dst_l = want_out ? a.out + 16ull * ((off_l >> 4) + rec_l) : nullptr;
                __syncwarp();
                const u32 os_l = w2_tiny_record(a.packed2 + p2_word(off_l, rec_l), n_l, dst_l, Xf, Xr, scr, false);
                u64 h_l = 0;
                if (want_hash) h_l = w2_any_hash((os_l & 1u) ? Xr : Xf, n_l, os_l >> 1);
                if (lane == 0) {
                    const u32 start = os_l >> 1, strand = os_l & 1u;
                    a.out_start[rec_l] = strand ? (n_l - 1 - start) : start;
                    a.out_strand[rec_l] = (u8)strand;
                    if (want_hash) a.out_hash[rec_l] = h_l;
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

}  // namespace ck
