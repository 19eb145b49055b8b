// ck_stream2.cuh -- lane-per-record LMSR + canonical form + XXH3-64 for 2-bit records, streaming form: the hot
// kernel of every 2-bit configuration of BASELINE.json (1, 2, 5 and the short end of 4).
//
// A warp takes 32 records and every lane walks ITS OWN record, so a warp instruction does useful work for 32
// records and no collective sits on the per-base path (the warp-per-record kernel of ck_warp2.cuh pays ~300 warp
// instructions of per-record fixed cost for a 325-base record).  Unlike ck_lane2.cuh nothing is staged in shared
// memory: the packed arena already holds every record 16-byte aligned and followed by its own circular extension
// (ck_device.cuh, "packed2 arena"), so each lane simply streams its record with 128-bit loads that are issued two
// iterations ahead (lane-private streaming reaches HBM speed on B200: tools/micro/stream.cu), occupancy is bounded by
// registers only, and the same kernel serves every record length:
//   * scan      : 8-mer keys as 16-bit halves (14 funnel shifts + 8 VIMNMX3.U16x2 per 32 rotations).  The reverse
//                 strand is never materialised: the 48-base block (x0,x1,x2) a step looks at is reverse-complemented
//                 in registers (BREV + LOP3 per unit), which yields the reverse strand's keys for forward starts
//                 32t+9 .. 32t+40.  Each lane keeps the two smallest (key, step, strand) triples, so "the minimal
//                 8-mer is unique" is known exactly.  Only the last step of a record can see positions twice (the
//                 extension repeats the head); its units are padded in registers (T past base n+7 for the forward
//                 keys, A past base n+16 for the reverse keys) so that those duplicates can never win;
//   * locate    : the winning step is replayed once; an XOR / VIMNMX / shift-add chain turns "which of the 32 keys
//                 equal the minimum" into a bit mask; exactly one valid bit => that rotation is the canonical one;
//   * ASCII+hash: each lane streams its canonical form in 16-byte chunks (the reverse strand costs nothing extra: a
//                 per-lane letter table "TGCA" and per-lane PRMT selectors do the complement and the reversal), four
//                 chunks = one XXH3 stripe per round, eight 64-bit accumulators in registers, the next round's
//                 units already in flight;
//   * write     : a round's 64 bytes per lane go through a padded shared-memory stage (the kernel's only shared
//                 memory) and leave as 128-bit stores in which four consecutive lanes cover 64 contiguous bytes;
//   * retry     : records a lane cannot finish alone (equal minimal 8-mers, n < 128, n = 128 when a hash is wanted)
//                 are appended to a retry list that the warp- / CTA-per-record kernels work off afterwards.
// Work lists come sorted by (class, length) (k_classify + radix sort), so the lanes of a warp run the same trip counts.
#pragma once
#include "ck_lane2.cuh"

namespace ck {

#define CK_S2_WARPS 8u
#define CK_S2_WARP_BYTES (CK_T2_AUX_BYTES + 1024u + 384u + 2560u)

__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
// same load as an asm volatile statement: the compiler must leave it where it is written (it would otherwise sink a
// load that is only used in the next unrolled loop instance into that instance, undoing the software pipeline)
__device__ __forceinline__ uint4 ldg128_here(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg64(const void *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ u32 ldg32(const void *p) { return __ldg(reinterpret_cast<const u32 *>(p)); }

__device__ __forceinline__ void cp_async16(u32 dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// the three smallest (key << 16 | step tag) values: two steps that share the minimal 8-mer can be told apart in the lane
// (s3_ext16) as long as a third one does not carry it too
__device__ __forceinline__ void t3_track(u32 &m1, u32 &m2, u32 &m3, u32 v)
{
    m3 = min(m3, max(m2, v));
    m2 = min(m2, max(m1, v));
    m1 = min(m1, v);
}
// the same for the two values of a step (forward, reverse) at once: the three smallest of the merge of (m1 <= m2 <= m3) and
// (lo <= hi) are min over k of max(m[3 - k], b[k]) -- 8 operations instead of 10
__device__ __forceinline__ void t3_track2(u32 &m1, u32 &m2, u32 &m3, u32 a, u32 b)
{
    const u32 lo = min(a, b), hi = max(a, b);
    const u32 n3 = __vimin3_u32(m3, max(m2, lo), max(m1, hi));
    const u32 n2 = __vimin3_u32(m2, max(m1, lo), hi);
    m1 = min(m1, lo); m2 = n2; m3 = n3;
}
// replay step (tag >> 1) of strand (tag & 1): bit s of the result = rotation s of the step carries the 8-mer (m >> 16) and lies below n
__device__ __forceinline__ u32 s3_replay(const u8 *base, u32 nn, u32 m)
{
    const u32 t = (m & 0xffffu) >> 1, strand = m & 1u;
    const uint2 x01 = ldg64(base + 8 * t);
    const u32 x2 = ldg32(base + 8 * t + 8);
    const u32 y0 = strand ? w2_revcomp(x2) : x01.x, y1 = strand ? w2_revcomp(x01.y) : x01.y;
    const u32 y2 = strand ? w2_revcomp(x01.x) : x2;
    const u32 bb = (m >> 16) * 0x10001u;
    u32 nm = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 wa = i ? __funnelshift_l(y1, y0, 2 * i) : y0, wb = i ? __funnelshift_l(y2, y1, 2 * i) : y1;
        nm += __vminu2(wa ^ bb, 0x00010001u) << i;         // bits 16 + i / i: key i / key 8 + i differs
        nm += __vminu2(wb ^ bb, 0x00010001u) << (i + 8);   // bits 24 + i / 8 + i: key 16 + i / key 24 + i differs
    }
    const u32 match = ~__byte_perm(nm, 0, 0x1302);         // bit s: key s of the step equals the minimum
    const int lim = (int)nn - 32 * (int)t;
    const u32 valid = strand ? (lim >= 32 ? 0xffffffffu : 0xffffffffu << (32 - lim))
                             : (lim >= 32 ? 0xffffffffu : (1u << lim) - 1u);
    return match & valid;
}
// rotation start, in the strand's own coordinates, of hit s of step (tag >> 1) of strand (tag & 1)
__device__ __forceinline__ u32 s3_start(u32 nn, u32 m, u32 s)
{
    const u32 t = (m & 0xffffu) >> 1;
    int st = (m & 1u) ? (int)nn - 48 - 32 * (int)t + (int)s : 32 * (int)t + (int)s;
    if (st < 0) st += (int)nn;
    return (u32)st;
}
// bases 8 .. 23 of the rotation that starts at `st` of `strand` (32 bits, first base on top): a linear window of the record in
// either arena layout (the reads end below n + 24, inside the single-copy layout's extension)
__device__ __forceinline__ u32 s3_ext16(const u8 *base, u32 nn, u32 strand, u32 st)
{
    int f0 = strand ? (int)nn - 24 - (int)st : (int)st + 8;      // forward position of the 16 bases (reverse strand: mirrored)
    if (f0 < 0) f0 += (int)nn;
    const u32 w = (u32)f0 >> 4, sh = 2u * ((u32)f0 & 15u);
    const u32 x = __funnelshift_l(ldg32(base + 4 * w + 4), ldg32(base + 4 * w), sh);
    return strand ? w2_revcomp(x) : x;
}

template <int V>
__global__ void __launch_bounds__(32 * CK_S2_WARPS, 2) k_canon_s2(CanonArgs a)
{
    extern __shared__ __align__(16) u32 smem[];
    constexpr bool want_hash = (V & CK_W2_HASH) != 0, want_out = (V & CK_W2_OUT) != 0, use_list = (V & CK_W2_LIST) != 0;
    // two rotations with the minimal 8-mer are told apart in the lane (as in ck_stream3.cuh) in the variants without a hash only:
    // measured on 325-base records, it takes config 1 (canonicalize) from 0.182 to 0.169 ms per step, but slows the hashing
    // variants' kernel by 5 % (1.77 -> 1.87 ms on config 5), more than the 1 % of records it keeps from the retry kernel are worth
    constexpr bool kPair = !want_hash;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const u32 aux = (u32)__cvta_generic_to_shared(smem) + wid * CK_S2_WARP_BYTES;   // output stage
    const u32 offs = aux + CK_T2_AUX_BYTES + 16u * lane;           // + 512 * slot: (offset, end) of this lane's record
    const u32 recs = aux + CK_T2_AUX_BYTES + 1024u + 4u * lane;    // + 128 * (batch % 3): record index (work lists)
    const u32 head = aux + CK_T2_AUX_BYTES + 1024u + 384u + 80u * lane;   // first three quads + last-step units of this lane's record
    const u32 count = use_list ? *a.count : a.n_direct;
    if (use_list) a.list += a.count[16];
    const u8 *arena = reinterpret_cast<const u8 *>(a.packed2);
    // request the head of a record (quads 0..2 and the three units of its last scan step) into this lane's head slot: issued
    // one batch ahead, so no lane waits for the first touch of its record
    auto fetch_head = [&](u64 off1, u32 n1, u32 rec1) {
        const u8 *nb = arena + 8ull * p2_word(off1, rec1, a.p2_dbl);
        const u32 s1 = n1 >= 128u ? ((n1 + 31u) >> 5) - 1u : 0u;
        cp_async16(head, nb); cp_async16(head + 16, nb + 16); cp_async16(head + 32, nb + 32);
        cp_async8(head + 48, nb + 8 * s1); cp_async8(head + 56, nb + 8 * s1 + 8);
        if (n1 <= 512u && n1 > 192u) prefetch_l2(nb + 64);         // short records (short batches): the rest of the record -> L2;
                                                                   // a long batch outlives what L2 would keep
    };

    // (offset, end) of a record, or (offset, normalised length) when the batch carries lengths of its own
    const bool has_lens = a.lens != nullptr;
    auto fetch_offsets = [&](u32 rec, u32 slot) {
        cp_async8(offs + 512u * slot, a.offsets + rec);
        if (has_lens) cp_async4(offs + 512u * slot + 8u, a.lens + rec);
        else cp_async8(offs + 512u * slot + 8u, a.offsets + rec + 1);
    };
    auto length_of = [&](const uint4 &oe) { return has_lens ? oe.z : oe.z - oe.x; };
    // Batches of 32 records are claimed from a launch-wide counter (a.retry_counts[28], zeroed with the class counters), so
    // every warp stays busy until the work runs out.  A warp knows its next three batches: the offsets of batch kb + 2
    // and the heads of batch kb + 1 are in flight while batch kb is processed.
    auto claim = [&]() {
        u32 v = 0;
        if (lane == 0) v = atomicAdd(a.retry_counts + 28, 1u);
        return __shfl_sync(CK_FULL, v, 0) * 32u;
    };
    // ---- prologue: offsets of batches 0 and 1, heads of batch 0, (work lists) record indices of batches 0..2
    u32 b0 = claim(), b1 = claim(), b2 = claim();
    u32 rq = 0;                                                    // work lists: record index two batches ahead
    {
        const u32 i0 = b0 + lane, i1 = b1 + lane, i2 = b2 + lane;
        u32 r0 = i0, r1 = i1;
        if (use_list) {
            r0 = i0 < count ? a.list[i0] : 0u; r1 = i1 < count ? a.list[i1] : 0u; rq = i2 < count ? a.list[i2] : 0u;
            sts32(recs, r0); sts32(recs + 128, r1);
        }
        if (i0 < count) fetch_offsets(r0, 0);
        if (i1 < count) fetch_offsets(r1, 1);
        cp_async_wait_all();
        __syncwarp();
        if (i0 < count) {
            const uint4 oe = lds128(offs);
            fetch_head(((u64)oe.y << 32) | oe.x, length_of(oe), r0);
        }
    }
    u32 kb = 0;                                                    // batch counter of this warp
    for (; b0 < count; kb++) {
        const u32 sl = kb & 1u;
        const u32 idx = b0 + lane;
        const bool have = idx < count;
        const u32 b3 = claim();
        cp_async_wait_all();
        __syncwarp();                                              // the offsets of batches kb and kb + 1 have landed
        u32 rec = idx; u64 off = 0; u32 n = 0;
        {
            const uint4 oe = lds128(offs + 512u * sl);
            if (use_list) rec = lds32(recs + 128u * (kb % 3u));
            if (have) { off = ((u64)oe.y << 32) | oe.x; n = length_of(oe); } else rec = 0;
            const u32 idx2 = b2 + lane;
            if (idx2 < count) fetch_offsets(use_list ? rq : idx2, sl);
            if (use_list) {
                sts32(recs + 128u * ((kb + 2u) % 3u), rq);
                const u32 idx3 = b3 + lane;
                rq = idx3 < count ? a.list[idx3] : 0u;
            }
        }
        // this batch's record heads (requested one batch ago), then the request for the next batch's
        const uint4 H0 = lds128(head), H1 = lds128(head + 16), H2 = lds128(head + 32), HT = lds128(head + 48);
        if (b1 + lane < count) {
            const uint4 oe = lds128(offs + 512u * (sl ^ 1u));
            const u32 rec1 = use_list ? lds32(recs + 128u * ((kb + 1u) % 3u)) : b1 + lane;
            fetch_head(((u64)oe.y << 32) | oe.x, length_of(oe), rec1);
        }
        const bool in_class = have && (use_list || (n >= a.min_n && n <= a.max_n));
        // direct mode (a promise of one class, no work list): this kernel is the only one that looks at every record, so it also
        // reports the records outside the promise (what k_classify does otherwise)
        if (!use_list && have && !in_class) atomicAdd(a.retry_counts + ((int)CLS_HUGE - 32), 1u);
        // lane-private fast path: n >= 128 (with a hash: n >= 129, the 129..240 and the long XXH3 forms)
        bool fast = in_class && n >= (want_hash ? 129u : 128u);
        const u8 *base = arena + 8ull * p2_word(off, rec, a.p2_dbl);         // this lane's record: units 0 .. jn + 4 are valid
        u8 *dst = want_out ? a.out + out_byte(off, rec) : nullptr;
        u32 os = 0;                                                // (start << 1) | strand, strand's own coordinates
        u64 h = 0;
        const u32 nn = fast ? n : 128u;                            // lanes without a fast record walk a dummy geometry

        // ---- scan: the two smallest (8-mer key, step, strand) over all 2n rotations
        const u32 S1 = ((nn + 31) >> 5) - 1;                       // full steps; step S1 is the padded last one
        const u32 S1max = __reduce_max_sync(CK_FULL, fast ? S1 : 0u);
        const u32 qmax = fast ? (S1 + 1) >> 1 : 0u;                // last quad this lane may read (units <= jn + 3)
        u32 m1 = 0xffffffffu, m2 = 0xffffffffu, m3 = 0xffffffffu;
        const uint2 l01 = make_uint2(HT.x, HT.y);                  // units of the last step, needed after the loop
        const u32 l2 = HT.z;
        {
            // four quads in flight (loads run three iterations ahead); the loop is unrolled four times so that they rotate
            // by renaming (a register move of a quad would wait for its load)
            uint4 QA = H0, QB = H1, QC = H2, QD;
            u32 rqx = w2_revcomp(QA.x);
            const u32 iters = (S1max + 1) >> 1;
#define CK_S2_SCAN(Q, Q1, Q2)                                                                                          \
            {                                                                                                          \
                Q2 = ldg128_here(base + 16 * min(i + 3, qmax));                                                        \
                if ((i & 3u) == 0) prefetch_l2(base + 16 * min(i + 16, qmax));                                         \
                const u32 rqy = w2_revcomp(Q.y), rqz = w2_revcomp(Q.z), rqw = w2_revcomp(Q.w), rnx = w2_revcomp(Q1.x); \
                {                                                                                                      \
                    const u32 t = 2 * i;                                                                               \
                    const u32 kf = t2_key_hi(w2_step_min16(Q.x, Q.y, Q.z)), kr = t2_key_hi(w2_step_min16(rqz, rqy, rqx)); \
                    const u32 tag = (t < S1) ? 2 * t : 0xffff0000u;                                                    \
                    if (kPair) t3_track2(m1, m2, m3, kf | tag, kr | (tag + 1));                                        \
                    else { t2_track(m1, m2, kf | tag); t2_track(m1, m2, kr | (tag + 1)); }                             \
                }                                                                                                      \
                {                                                                                                      \
                    const u32 t = 2 * i + 1;                                                                           \
                    const u32 kf = t2_key_hi(w2_step_min16(Q.z, Q.w, Q1.x)), kr = t2_key_hi(w2_step_min16(rnx, rqw, rqz)); \
                    const u32 tag = (t < S1) ? 2 * t : 0xffff0000u;                                                    \
                    if (kPair) t3_track2(m1, m2, m3, kf | tag, kr | (tag + 1));                                        \
                    else { t2_track(m1, m2, kf | tag); t2_track(m1, m2, kr | (tag + 1)); }                             \
                }                                                                                                      \
                rqx = rnx;                                                                                             \
                if (++i >= iters) break;                                                                               \
            }
            if (iters) {
                u32 i = 0;
#pragma unroll 1
                for (;;) {
                    CK_S2_SCAN(QA, QB, QD)
                    CK_S2_SCAN(QB, QC, QA)
                    CK_S2_SCAN(QC, QD, QB)
                    CK_S2_SCAN(QD, QA, QC)
                }
            }
#undef CK_S2_SCAN
        }
        {   // last step of this lane's record: positions 32 S1 .. n - 1 are new, the rest repeats the head
            const int dT = (int)nn + 7 - 32 * (int)S1;             // first forward-padded base, relative to unit 2 S1
            const int dA = dT + 9;                                 // first reverse-padded base
#define CK_T2_PAD(d) __funnelshift_rc(0xffffffffu, 0u, 2u * (u32)min(max((d), 0), 16))
            const u32 pT0 = CK_T2_PAD(dT), pT1 = CK_T2_PAD(dT - 16), pT2 = CK_T2_PAD(dT - 32);
            const u32 pA0 = CK_T2_PAD(dA), pA1 = CK_T2_PAD(dA - 16), pA2 = CK_T2_PAD(dA - 32);
#undef CK_T2_PAD
            const u32 kf = t2_key_hi(w2_step_min16(l01.x | pT0, l01.y | pT1, l2 | pT2));
            const u32 kr = t2_key_hi(w2_step_min16(w2_revcomp(l2 & ~pA2), w2_revcomp(l01.y & ~pA1), w2_revcomp(l01.x & ~pA0)));
            if (kPair) t3_track2(m1, m2, m3, kf | (2 * S1), kr | (2 * S1 + 1));
            else { t2_track(m1, m2, kf | (2 * S1)); t2_track(m1, m2, kr | (2 * S1 + 1)); }
        }
        // ---- locate: replay the winning step, find the rotation that carries the minimal 8-mer.  Exactly two rotations
        //      with it (9 % of 3 kb records) are told apart by their next 16 bases; anything else that ties takes the duel path
        {
            const u32 mm1 = fast ? m1 : 0u;
            const bool eq2 = (m1 ^ m2) < 0x10000u, eq3 = !kPair || (m1 ^ m3) < 0x10000u;
            const u32 hits1 = s3_replay(base, nn, mm1);
            u32 hits2 = 0;
            const bool pair = fast && eq2 && !eq3;
            if (__any_sync(CK_FULL, pair)) hits2 = s3_replay(base, nn, pair ? m2 : mm1);
            const u32 c1 = __popc(hits1), c2 = pair ? __popc(hits2) : 0u;
            u32 st = s3_start(nn, mm1, __ffs(hits1) - 1), strand = mm1 & 1u;
            if (!(c1 == 1 && !eq2)) {
                if (fast && !eq3 && c1 + c2 == 2 && (!eq2 || pair)) {
                    // candidate B: the hit of the second step, or the second hit of the same step
                    const u32 mb = c2 ? m2 : mm1;
                    const u32 stb = s3_start(nn, mb, c2 ? __ffs(hits2) - 1 : 31u - __clz(hits1)), sb = mb & 1u;
                    const u32 ea = s3_ext16(base, nn, strand, st), eb = s3_ext16(base, nn, sb, stb);
                    if (ea == eb) fast = false;
                    else if (eb < ea) { st = stb; strand = sb; }
                } else fast = false;                               // three or more, or none valid: the duel path decides
            }
            os = (st << 1) | strand;
            if (!fast) os = 0;                                     // keeps the dummy walk below inside the record
        }
        // ---- canonical ASCII (+ XXH3-64), lane-private; one stripe (4 chunks of 16 bytes) per round.
        //      A round reads the 64 + 15 bases that start at forward position B (two aligned 128-bit loads; the in-arena
        //      extension makes the read linear): windows B, B+16, B+32, B+48.  Forward strand: chunk k = window k and
        //      B advances by 64; reverse strand: chunk k = reverse complement of window 3 - k and B retreats by 64.
        if (want_out || want_hash) {
            const u32 strand = os & 1u;
            const u32 T = strand ? 0x41434754u : 0x54474341u;      // "TGCA" / "ACGT"
            const u32 sa = strand ? 0x5140u : 0x2637u, sb = strand ? 0x7362u : 0x0415u, rot = strand ? 16u : 0u;
            const int step = strand ? -64 : 64, nstep = strand ? -(int)nn : (int)nn;
            // forward position of canonical chunk 0: the rotation start, or the mirror of reverse position start
            int p0 = strand ? (int)nn - 16 - (int)(os >> 1) : (int)(os >> 1);
            if (p0 < 0) p0 += (int)nn;
            int B = strand ? p0 - 48 : p0;                         // lowest window of round 0
            if (B < 0) B += (int)nn;
            const u32 nchunks = fast ? (nn + 15) >> 4 : 0u;
            const u32 nfull = fast ? (nn - 1) >> 6 : 0u;           // stripes the stripe loop hashes
            const u32 rounds = __reduce_max_sync(CK_FULL, (nchunks + 3) >> 2);
            u64 acc0 = CK_P32_3, acc1 = CK_P64_1, acc2 = CK_P64_2, acc3 = CK_P64_3;
            u64 acc4 = CK_P64_4, acc5 = CK_P32_2, acc6 = CK_P64_5, acc7 = CK_P32_1;
            // XXH3 129..240 form: one 16-byte mix per chunk instead of stripes (work lists group such records in their own warps)
            const u32 nbr = (want_hash && fast && nn <= 240u) ? nn >> 4 : 0u;    // rounds of 16 bytes
            const bool any_mid = want_hash && __any_sync(CK_FULL, nbr != 0);
            u64 mida = 0, midb = 0;
            // the four records whose bytes this lane carries out of the stage (records 8 i + (lane >> 2), piece lane & 3): where
            // the next piece goes (16-byte granules from a.out) and how many of its pieces are left
            u32 og[4] = {0, 0, 0, 0}; int orem[4] = {0, 0, 0, 0};
            if (want_out) {
                const u32 gr = (u32)((dst - a.out) >> 4);
#pragma unroll
                for (u32 i = 0; i < 4; i++) {
                    og[i] = __shfl_sync(CK_FULL, gr, 8 * i + (lane >> 2)) + (lane & 3u);
                    orem[i] = (int)__shfl_sync(CK_FULL, nchunks, 8 * i + (lane >> 2)) - (int)(lane & 3u);
                }
            }
            // output stage, conflict-free for the writes (lane = row) and the reads (4 lanes = one row): row r, chunk c ->
            // 128-byte line r >> 1, 16-byte slot (c + 4 (r & 1) + ((r >> 1) & 3)) & 7 (ck_stream3.cuh)
            const u32 st_w = aux + 128u * (lane >> 1), st_j = 4u * (lane & 1u) + ((lane >> 1) & 3u);
            const u32 st_w0 = st_w + 16u * (st_j & 7u), st_w1 = st_w + 16u * ((st_j + 1u) & 7u);
            const u32 st_w2 = st_w + 16u * ((st_j + 2u) & 7u), st_w3 = st_w + 16u * ((st_j + 3u) & 7u);
            const u32 st_r = aux + 128u * (lane >> 3) + 16u * (((lane & 3u) + 4u * ((lane >> 2) & 1u) + ((lane >> 3) & 3u)) & 7u);
            const u64 *sec = reinterpret_cast<const u64 *>(c_secret);
            // the two quads of a round and its (unit offset, bit shift), in two register sets: the next round's are requested
            // into the other set before this round's are used (the loop is unrolled twice, so nothing is ever moved)
            uint4 XA0, XB0, XA1, XB1; u32 xa0, xs0, xa1, xs1;
#define CK_S2_FETCH(XA, XB, xa, xs)                                                                                 \
            {                                                                                                       \
                const u8 *ad = base + (((u32)B >> 6) << 4);                                                         \
                XA = ldg128_here(ad); XB = ldg128_here(ad + 16);                                                    \
                xa = ((u32)B >> 4) & 3u; xs = 2u * (u32)B;                                                          \
                B += step;                                                                                          \
                if ((u32)B >= nn) B -= nstep;                                                                       \
            }
            // four canonical 16-base windows of the fetched round, in chunk order
#define CK_S2_WINDOWS(W, XA, XB, xa, xs)                                                                            \
            {                                                                                                       \
                const bool a2 = (xa & 2u) != 0, a1 = (xa & 1u) != 0;                                                \
                const u32 y0 = a2 ? XA.z : XA.x, y1 = a2 ? XA.w : XA.y, y2 = a2 ? XB.x : XA.z, y3 = a2 ? XB.y : XA.w; \
                const u32 y4 = a2 ? XB.z : XB.x, y5 = a2 ? XB.w : XB.y;                                             \
                const u32 u0 = a1 ? y1 : y0, u1 = a1 ? y2 : y1, u2 = a1 ? y3 : y2, u3 = a1 ? y4 : y3, u4 = a1 ? y5 : y4; \
                const u32 w0 = __funnelshift_l(u1, u0, xs), w1 = __funnelshift_l(u2, u1, xs);                       \
                const u32 w2 = __funnelshift_l(u3, u2, xs), w3 = __funnelshift_l(u4, u3, xs);                       \
                W[0] = strand ? w3 : w0; W[1] = strand ? w2 : w1; W[2] = strand ? w1 : w2; W[3] = strand ? w0 : w3; \
            }
#define CK_T2_SCR(A, I) do { A ^= A >> 47; A ^= sec[16 + I]; A *= CK_P32_1; } while (0)
#define CK_S2_ROUND(XA, XB, xa, xs, YA, YB, ya, ys)                                                                 \
            {                                                                                                       \
                uint4 v[4];                                                                                         \
                {                                                                                                   \
                    u32 W[4];                                                                                       \
                    CK_S2_WINDOWS(W, XA, XB, xa, xs);                                                               \
                    if (s + 1 < rounds) {                                                                           \
                        CK_S2_FETCH(YA, YB, ya, ys);                                                                \
                        const int pa = B + 3 * step;           /* four rounds ahead, inside the record only */      \
                        if ((u32)pa < nn) prefetch_l2(base + (((u32)pa >> 6) << 4));                                \
                    }                                                                                               \
                    _Pragma("unroll") for (int k = 0; k < 4; k++) v[k] = t2_ascii16(__funnelshift_l(W[k], W[k], rot), T, sa, sb); \
                }                                                                                                   \
                if (want_hash && s < nfull) {                                                                       \
                    const u32 ks = s & 15u;                                                                         \
                    t2_acc16(acc0, acc1, v[0], sec[ks + 0], sec[ks + 1]);                                           \
                    t2_acc16(acc2, acc3, v[1], sec[ks + 2], sec[ks + 3]);                                           \
                    t2_acc16(acc4, acc5, v[2], sec[ks + 4], sec[ks + 5]);                                           \
                    t2_acc16(acc6, acc7, v[3], sec[ks + 6], sec[ks + 7]);                                           \
                    if (ks == 15u) {       /* a 1024-byte block is complete (s + 1 <= nfull: more input follows) */ \
                        CK_T2_SCR(acc0, 0); CK_T2_SCR(acc1, 1); CK_T2_SCR(acc2, 2); CK_T2_SCR(acc3, 3);             \
                        CK_T2_SCR(acc4, 4); CK_T2_SCR(acc5, 5); CK_T2_SCR(acc6, 6); CK_T2_SCR(acc7, 7);             \
                    }                                                                                               \
                }                                                                                                   \
                if (any_mid) {                                                                                      \
                    _Pragma("unroll") for (u32 k = 0; k < 4; k++) {                                                 \
                        const u32 i = 4 * s + k;                                                                    \
                        const u64 lo = ((u64)v[k].y << 32) | v[k].x, hi = ((u64)v[k].w << 32) | v[k].z;             \
                        if (i < 8) { const u64 t = mul128_fold64(lo ^ sec[2 * i], hi ^ sec[2 * i + 1]); if (i < nbr) mida += t; }      \
                        else if (i < 15) { const u64 t = mul128_fold64(lo ^ c_midsec[2 * i - 16], hi ^ c_midsec[2 * i - 15]); if (i < nbr) midb += t; } \
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
                CK_S2_FETCH(XA0, XB0, xa0, xs0);
                u32 s = 0;
#pragma unroll 1
                for (;;) {
                    CK_S2_ROUND(XA0, XB0, xa0, xs0, XA1, XB1, xa1, xs1)
                    CK_S2_ROUND(XA1, XB1, xa1, xs1, XA0, XB0, xa0, xs0)
                }
            }
            if (want_hash) {
                // last stripe: canonical bytes [n - 64, n) = the four chunks that end where chunk 0 starts
                B = strand ? p0 + 16 : p0 - 64;
                if ((u32)B >= nn) B += nstep;
                CK_S2_FETCH(XA0, XB0, xa0, xs0);
                u32 W[4];
                CK_S2_WINDOWS(W, XA0, XB0, xa0, xs0);
                {
                    const uint4 v0 = t2_ascii16(__funnelshift_l(W[0], W[0], rot), T, sa, sb);
                    t2_acc16(acc0, acc1, v0, c_lastsec[0], c_lastsec[1]);
                    const uint4 v1 = t2_ascii16(__funnelshift_l(W[1], W[1], rot), T, sa, sb);
                    t2_acc16(acc2, acc3, v1, c_lastsec[2], c_lastsec[3]);
                    const uint4 v2 = t2_ascii16(__funnelshift_l(W[2], W[2], rot), T, sa, sb);
                    t2_acc16(acc4, acc5, v2, c_lastsec[4], c_lastsec[5]);
                    const uint4 v3 = t2_ascii16(__funnelshift_l(W[3], W[3], rot), T, sa, sb);
                    t2_acc16(acc6, acc7, v3, c_lastsec[6], c_lastsec[7]);
                }
                u64 r = (u64)nn * CK_P64_1;
                r += mul128_fold64(acc0 ^ c_mergesec[0], acc1 ^ c_mergesec[1]);
                r += mul128_fold64(acc2 ^ c_mergesec[2], acc3 ^ c_mergesec[3]);
                r += mul128_fold64(acc4 ^ c_mergesec[4], acc5 ^ c_mergesec[5]);
                r += mul128_fold64(acc6 ^ c_mergesec[6], acc7 ^ c_mergesec[7]);
                h = xxh3_avalanche(r);
            }
            if (any_mid) {
                // the last 16 canonical bytes [n - 16, n): the chunk that ends where chunk 0 starts
                int q = strand ? p0 + 16 : p0 - 16;
                if ((u32)q >= nn) q += nstep;
                const u8 *ad = base + (((u32)q >> 4) << 2);
                const u32 w = __funnelshift_l(ldg32(ad + 4), ldg32(ad), 2u * (u32)q);
                const uint4 vv = t2_ascii16(__funnelshift_l(w, w, rot), T, sa, sb);
                const u64 lo = ((u64)vv.y << 32) | vv.x, hi = ((u64)vv.w << 32) | vv.z;
                const u64 tail = mul128_fold64(lo ^ c_midsec[14], hi ^ c_midsec[15]);
                if (nbr) h = xxh3_avalanche(xxh3_avalanche((u64)nn * CK_P64_1 + mida) + midb + tail);
            }
#undef CK_S2_FETCH
#undef CK_S2_WINDOWS
#undef CK_S2_ROUND
#undef CK_T2_SCR
        }
        if (fast) {
            const u32 start = os >> 1, strand = os & 1u;
            a.out_start[rec] = strand ? (n - 1 - start) : start;
            a.out_strand[rec] = (u8)strand;
            if (want_hash) a.out_hash[rec] = h;
        } else if (in_class) {
            // retry list of the record's class: a.retry + (first entry of the class) + (entries so far)
            const int c = n <= cls_max_n(CLS_W2S) ? CLS_W2S : n <= cls_max_n(CLS_W2M) ? CLS_W2M : n <= cls_max_n(CLS_W2L) ? CLS_W2L
                        : n <= cls_max_n(CLS_W2X) ? CLS_W2X : n <= cls_max_n(CLS_C2A) ? CLS_C2A : CLS_C2B;
            const u32 k = atomicAdd(a.retry_counts + c, 1u);
            a.retry[a.retry_counts[16 + c] + k] = rec;
        }
        __syncwarp();
        b0 = b1; b1 = b2; b2 = b3;
    }
}

}  // namespace ck
