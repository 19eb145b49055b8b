// ck_stream3.cuh -- lane-per-record LMSR + canonical form + XXH3-64 for 2-bit records: the hot kernel of every 2-bit
// configuration of BASELINE.json (1, 2, 5 and the short end of 4).  Successor of ck_stream2.cuh (kept for A/B runs).
//
// A warp takes 32 records and every lane walks ITS OWN record (no collective on the per-base path).  What bounded the
// previous form was not DRAM and not the ALU pipe but the L1TEX data pipe (ncu: 76 % busy on configs 1 and 2): a
// lane-private 128-bit load touches 32 different cache lines, i.e. 32 wavefronts per warp instruction for 512 bytes.
// This form is built around that:
//   * the arena holds every record 32-byte aligned and DOUBLED (ck_device.cuh), so every rotation is a linear window:
//     both passes stream with 256-bit loads (LDG.E.ENL2.256, sm_100) -- half the wavefronts per byte -- and the emit
//     pass has no wrap logic at all;
//   * scan      : one 32-byte oct = 128 bases = four 32-rotation steps per iteration, octs two iterations ahead in three
//                 rotating register sets; 8-mer keys as u16x2 halves, the reverse strand's keys from the block
//                 reverse-complemented in registers, the two smallest (key, step, strand) per lane;
//   * locate    : replay of the winning step, bit mask of the rotations that carry the minimal key; a unique one is the
//                 canonical rotation, anything else goes to the retry list (warp / CTA kernels);
//   * emit      : per iteration 128 canonical bases = two XXH3 stripes.  The window's 9 units are selected from the two
//                 octs a lane holds (31 SEL), the next oct is requested into the registers of the oct that just died
//                 BEFORE the 128 bytes are generated.  Forward-strand lanes walk up the arena and replace the lower oct,
//                 reverse-strand lanes walk down and replace the upper one -- two loads predicated on the strand, so the
//                 (lower, upper) roles of the two register sets alternate uniformly across the warp and nothing is moved;
//   * write     : a round's 64 bytes per lane cross a shared-memory stage whose layout is conflict-free for the 128-bit
//                 writes (lane = row) AND for the 128-bit reads (4 lanes = one row): row r, chunk c -> 128-byte line r >> 1,
//                 16-byte slot (c + 4 (r & 1) + ((r >> 1) & 3)) & 7.
#pragma once
#include "ck_stream2.cuh"

namespace ck {

#define CK_S3_WARPS 8u
#define CK_S3_STAGE 2048u
#define CK_S3_WARP_BYTES (CK_S3_STAGE + 1024u + 384u + 2560u)

struct Oct { uint4 lo, hi; };

__device__ __forceinline__ Oct ldg256_here(const void *p)
{
    Oct o;
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(o.lo.x), "=r"(o.lo.y), "=r"(o.lo.z), "=r"(o.lo.w), "=r"(o.hi.x), "=r"(o.hi.y), "=r"(o.hi.z), "=r"(o.hi.w)
                 : "l"(p));
    return o;
}
// the same load under a predicate: lanes with !on keep what the registers hold
__device__ __forceinline__ void ldg256_if(Oct &o, const void *p, u32 on)
{
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %9, 0;\n@q ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n}"
                 : "+r"(o.lo.x), "+r"(o.lo.y), "+r"(o.lo.z), "+r"(o.lo.w), "+r"(o.hi.x), "+r"(o.hi.y), "+r"(o.hi.z), "+r"(o.hi.w)
                 : "l"(p), "r"(on));
}

template <int V>
__global__ void __launch_bounds__(32 * CK_S3_WARPS, 2) k_canon_s3(CanonArgs a)
{
    extern __shared__ __align__(16) u32 smem[];
    constexpr bool want_hash = (V & CK_W2_HASH) != 0, want_out = (V & CK_W2_OUT) != 0, use_list = (V & CK_W2_LIST) != 0;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 aux = (u32)__cvta_generic_to_shared(smem) + wid * CK_S3_WARP_BYTES;   // output stage
    const u32 offs = aux + CK_S3_STAGE + 16u * lane;               // + 512 * slot: (offset, end) of this lane's record
    const u32 recs = aux + CK_S3_STAGE + 1024u + 4u * lane;        // + 128 * (batch % 3): record index (work lists)
    const u32 head = aux + CK_S3_STAGE + 1024u + 384u + 80u * lane;   // octs 0, 1 + last-step units of this lane's record
    const u32 count = use_list ? *a.count : a.n_direct;
    if (use_list) a.list += a.count[16];
    const u8 *arena = reinterpret_cast<const u8 *>(a.packed2);
    const bool pf_scan = !(a.mode & 0x100u);   // CK_S3_DEBUG bit 8: L2 prefetch of the scan off
    // request the head of a record (octs 0 and 1, the units of its last scan step) into this lane's head slot, one batch
    // ahead, and the rest of a short record into L2
    auto fetch_head = [&](u64 off1, u32 n1, u32 rec1) {
        const u8 *nb = arena + 8ull * p2_word(off1, rec1, 1u);
        const u32 s1 = n1 >= 128u ? ((n1 + 31u) >> 5) - 1u : 0u;
        cp_async16(head, nb); cp_async16(head + 16, nb + 16); cp_async16(head + 32, nb + 32); cp_async16(head + 48, nb + 48);
        cp_async8(head + 64, nb + 8 * s1); cp_async8(head + 72, nb + 8 * s1 + 8);
        // (L2 prefetches of the rest of a short record measured 1-2 % slower on configs 1 and 5: its loads come soon enough)
    };
    const bool has_lens = a.lens != nullptr;
    auto fetch_offsets = [&](u32 rec, u32 slot) {
        cp_async8(offs + 512u * slot, a.offsets + rec);
        if (has_lens) cp_async4(offs + 512u * slot + 8u, a.lens + rec);
        else cp_async8(offs + 512u * slot + 8u, a.offsets + rec + 1);
    };
    auto length_of = [&](const uint4 &oe) { return has_lens ? oe.z : oe.z - oe.x; };
    // batches of 32 records are claimed from a launch-wide counter (a.retry_counts[28], zeroed with the class counters)
    auto claim = [&]() {
        u32 v = 0;
        if (lane == 0) v = atomicAdd(a.retry_counts + 28, 1u);
        return __shfl_sync(CK_FULL, v, 0) * 32u;
    };
    u32 b0 = claim(), b1 = claim(), b2 = claim();
    u32 rq = 0;
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
    // output stage addresses (conflict-free both ways, see the header): this lane's row as a writer, and as a reader the
    // piece (lane & 3) of row 8 i + (lane >> 2)
    const u32 st_w = aux + 128u * (lane >> 1);
    const u32 st_j = 4u * (lane & 1u) + ((lane >> 1) & 3u);
    const u32 st_w0 = st_w + 16u * (st_j & 7u), st_w1 = st_w + 16u * ((st_j + 1u) & 7u);
    const u32 st_w2 = st_w + 16u * ((st_j + 2u) & 7u), st_w3 = st_w + 16u * ((st_j + 3u) & 7u);
    const u32 st_r = aux + 128u * (lane >> 3) + 16u * (((lane & 3u) + 4u * ((lane >> 2) & 1u) + ((lane >> 3) & 3u)) & 7u);   // + 512 i

    u32 kb = 0;
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
        const uint4 H0 = lds128(head), H1 = lds128(head + 16), H2 = lds128(head + 32), H3 = lds128(head + 48), HT = lds128(head + 64);
        if (b1 + lane < count) {
            const uint4 oe = lds128(offs + 512u * (sl ^ 1u));
            const u32 rec1 = use_list ? lds32(recs + 128u * ((kb + 1u) % 3u)) : b1 + lane;
            fetch_head(((u64)oe.y << 32) | oe.x, length_of(oe), rec1);
        }
        const bool in_class = have && (use_list || (n >= a.min_n && n <= a.max_n));
        // direct mode (a promise of one class, no work list): this kernel is the only one that looks at every record, so it also
        // reports the records outside the promise (what k_classify does otherwise)
        if (!use_list && have && !in_class) atomicAdd(a.retry_counts + ((int)CLS_HUGE - 32), 1u);
        bool fast = in_class && n >= (want_hash ? 129u : 128u);
        const u8 *base = arena + 8ull * p2_word(off, rec, 1u);         // this lane's record, doubled
        u8 *dst = want_out ? a.out + out_byte(off, rec) : nullptr;
        u32 os = 0;                                                // (start << 1) | strand, strand's own coordinates
        u64 h = 0;
        const u32 nn = fast ? n : 128u;                            // lanes without a fast record walk a dummy geometry

        // ---- scan: the two smallest (8-mer key, step, strand) over all 2n rotations
        const u32 S1 = ((nn + 31) >> 5) - 1;                       // full steps; step S1 is the padded last one
        const u32 S1max = __reduce_max_sync(CK_FULL, fast ? S1 : 0u);
        const u32 omax = fast ? ((S1 - 1) >> 2) + 1 : 0u;          // last oct this lane reads
        u32 m1 = 0xffffffffu, m2 = 0xffffffffu, m3 = 0xffffffffu;
        const bool pf_long = pf_scan && S1max >= 32u;             // L2 prefetch ahead of the scan: long records only (+2 % on config 2, -1 % on 1 and 5)
        const uint2 l01 = make_uint2(HT.x, HT.y);
        const u32 l2 = HT.z;
        {
            Oct OA, OB, OC;
            OA.lo = H0; OA.hi = H1; OB.lo = H2; OB.hi = H3;
            u32 r0 = w2_revcomp(OA.lo.x);
            const u32 iters = (S1max + 3) >> 2;
#ifndef CK_S3_AHEAD
#ifndef CK_S3_SCAN_UNROLL
#define CK_S3_AHEAD 3u
#else
#define CK_S3_AHEAD 2u
#endif
#endif
#define CK_S3_STEP(t, x0, x1, x2, q0, q1, q2)                                                                          \
            {                                                                                                          \
                const u32 kf = t2_key_hi(w2_step_min16(x0, x1, x2)), kr = t2_key_hi(w2_step_min16(q2, q1, q0));        \
                const u32 tag = ((t) < S1) ? 2 * (t) : 0xffff0000u;                                                    \
                t3_track2(m1, m2, m3, kf | tag, kr | (tag + 1));                                                       \
            }
#define CK_S3_SCAN(O, O1, O2)                                                                                          \
            {                                                                                                          \
                O2 = ldg256_here(base + 32 * min(j + CK_S3_AHEAD, omax));                                              \
                if ((j & 1u) == 0 && pf_long) prefetch_l2(base + 32 * min(j + 8, omax));                                          \
                const u32 r1 = w2_revcomp(O.lo.y), r2 = w2_revcomp(O.lo.z), r3 = w2_revcomp(O.lo.w);                   \
                const u32 r4 = w2_revcomp(O.hi.x), r5 = w2_revcomp(O.hi.y), r6 = w2_revcomp(O.hi.z);                   \
                const u32 r7 = w2_revcomp(O.hi.w), r8 = w2_revcomp(O1.lo.x);                                           \
                const u32 t = 4 * j;                                                                                   \
                CK_S3_STEP(t, O.lo.x, O.lo.y, O.lo.z, r0, r1, r2)                                                      \
                CK_S3_STEP(t + 1, O.lo.z, O.lo.w, O.hi.x, r2, r3, r4)                                                  \
                CK_S3_STEP(t + 2, O.hi.x, O.hi.y, O.hi.z, r4, r5, r6)                                                  \
                CK_S3_STEP(t + 3, O.hi.z, O.hi.w, O1.lo.x, r6, r7, r8)                                                 \
                r0 = r8;                                                                                               \
                if (++j >= iters) break;                                                                               \
            }
            if (iters) {
                u32 j = 0;
#ifndef CK_S3_SCAN_UNROLL
                // one loop body, the three octs rotate by register moves (OB has been used, OC was requested an iteration ago)
                OC = ldg256_here(base + 32 * min(2u, omax));
#if CK_S3_AHEAD == 4
                Oct OD = ldg256_here(base + 32 * min(3u, omax));
#endif
#pragma unroll 1
                for (;;) {
                    Oct ON;
                    CK_S3_SCAN(OA, OB, ON)
#if CK_S3_AHEAD == 4
                    OA = OB; OB = OC; OC = OD; OD = ON;
#else
                    OA = OB; OB = OC; OC = ON;
#endif
                }
#else
#pragma unroll 1
                for (;;) {
                    CK_S3_SCAN(OA, OB, OC)
                    CK_S3_SCAN(OB, OC, OA)
                    CK_S3_SCAN(OC, OA, OB)
                }
#endif
            }
#undef CK_S3_SCAN
#undef CK_S3_STEP
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
            t3_track2(m1, m2, m3, kf | (2 * S1), kr | (2 * S1 + 1));
        }
        // ---- locate: replay the winning step, find the rotation that carries the minimal 8-mer.  Exactly two rotations
        //      with it (9 % of 3 kb records) are told apart by their next 16 bases; anything else that ties takes the duel path
        {
            const u32 mm1 = fast ? m1 : 0u;
            const bool eq2 = (m1 ^ m2) < 0x10000u, eq3 = (m1 ^ m3) < 0x10000u;
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
        // ---- canonical ASCII (+ XXH3-64), lane-private, 128 bases (two stripes) per iteration.  In the doubled record D
        //      (D[b] = S[b mod n]) canonical chunk c (16 bases) is D[p0 + 16 c ..) on the forward strand and the reverse
        //      complement of D[p0 + n - 16 c ..) on the reverse strand: iteration m reads the 128 + 15 bases that start at
        //      Bv + 128 m (forward) / Bv - 128 m (reverse), a linear window either way.
        if (want_out || want_hash) {
            const u32 strand = os & 1u;
            const u32 T = strand ? 0x41434754u : 0x54474341u;      // "TGCA" / "ACGT"
            const u32 sa = strand ? 0x5140u : 0x2637u, sb = strand ? 0x7362u : 0x0415u, rot = strand ? 16u : 0u;
            // forward position of canonical chunk 0: the rotation start, or the mirror of reverse position start
            int p0 = strand ? (int)nn - 16 - (int)(os >> 1) : (int)(os >> 1);
            if (p0 < 0) p0 += (int)nn;
            const u32 Bv = strand ? (u32)p0 + nn - 112u : (u32)p0;  // lowest base of iteration 0's window (nn >= 128)
            const u32 xs = 2u * (Bv & 15u);
            const bool b2 = (Bv & 64u) != 0, b1 = (Bv & 32u) != 0, b0s = (Bv & 16u) != 0;
            const int ostep = fast ? (strand ? -32 : 32) : 0;     // lanes without a fast record stay on their first oct
            const int olim = (int)(((2u * nn + 127u) >> 7) << 5);  // last oct inside the record's own allocation
            const u32 nchunks = fast ? (nn + 15) >> 4 : 0u;
            const u32 nfull = fast ? (nn - 1) >> 6 : 0u;           // stripes the stripe loop hashes
            u64 acc0 = CK_P32_3, acc1 = CK_P64_1, acc2 = CK_P64_2, acc3 = CK_P64_3;
            u64 acc4 = CK_P64_4, acc5 = CK_P32_2, acc6 = CK_P64_5, acc7 = CK_P32_1;
            const u32 nbr = (want_hash && fast && nn <= 240u) ? nn >> 4 : 0u;    // XXH3 129..240 form: rounds of 16 bytes
            const bool any_mid = want_hash && __any_sync(CK_FULL, nbr != 0);
            u64 mida = 0, midb = 0;
            u32 og[4] = {0, 0, 0, 0}; int orem[4] = {0, 0, 0, 0};
            if (want_out) {
                const u32 gr = (u32)((dst - a.out) >> 4);
#pragma unroll
                for (u32 i = 0; i < 4; i++) {
                    og[i] = __shfl_sync(CK_FULL, gr, 8 * i + (lane >> 2)) + (lane & 3u);
                    orem[i] = (int)__shfl_sync(CK_FULL, nchunks, 8 * i + (lane >> 2)) - (int)(lane & 3u);
                }
            }
            const u64 *sec = reinterpret_cast<const u64 *>(c_secret);
            // byte offset (from base) of the oct this lane requests next: forward lanes replace the lower oct by the one
            // above the pair they hold, reverse lanes the upper oct by the one below; never below the record (the
            // iterations past a record's end are masked, their windows only have to be readable)
            const u32 o0 = fast ? (Bv >> 7) << 5 : 0u;
            int onext = fast ? (strand ? (int)o0 - 32 : (int)o0 + 64) : 0;
            Oct R0 = ldg256_here(base + o0), R1 = ldg256_here(base + o0 + 32);
#ifdef CK_S3_OG2
            // destinations as pointers + a warp-uniform round offset: nothing to update per round
            uint4 *op[4];
            _Pragma("unroll") for (u32 i = 0; i < 4; i++) op[i] = reinterpret_cast<uint4 *>(a.out) + og[i];
#define CK_S3_STORE4 _Pragma("unroll") for (u32 i = 0; i < 4; i++) { if (orem[i] > (int)(4u * s)) op[i][4u * s] = g[i]; }
#else
#define CK_S3_STORE4 _Pragma("unroll") for (u32 i = 0; i < 4; i++) { if (orem[i] > 0) reinterpret_cast<uint4 *>(a.out)[og[i]] = g[i]; og[i] += 4; orem[i] -= 4; }
#endif
#define CK_T2_SCR(A, I) do { A ^= A >> 47; A ^= sec[16 + I]; A *= CK_P32_1; } while (0)
            // one 64-byte round: chunks W0..W3 of stripe s.  MID: the warp holds records of the XXH3 129..240 form
#define CK_S3_ROUND(MID)                                                                                            \
            {                                                                                                       \
                uint4 v[4];                                                                                         \
                v[0] = t2_ascii16(__funnelshift_l(W0, W0, rot), T, sa, sb);                                         \
                v[1] = t2_ascii16(__funnelshift_l(W1, W1, rot), T, sa, sb);                                         \
                v[2] = t2_ascii16(__funnelshift_l(W2, W2, rot), T, sa, sb);                                         \
                v[3] = t2_ascii16(__funnelshift_l(W3, W3, rot), T, sa, sb);                                         \
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
                if (MID) {                                                                                          \
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
                    CK_S3_STORE4                                                                                    \
                    __syncwarp();                                                                                   \
                }                                                                                                   \
            }
            // two rounds with (LO, UP) = the lower / upper oct of the window.  The round body exists once per instance (a
            // real loop over the two halves): with both instances and the scan loop the hot code stays well inside the
            // instruction cache (a fully unrolled form ran into `no_instructions` stalls: profiles/r02_c_*).
#ifdef CK_S3_UNROLL_HH
#define CK_S3_HH_LOOP _Pragma("unroll")
#else
#define CK_S3_HH_LOOP _Pragma("unroll 1")
#endif
#define CK_S3_PAIR(LO, UP, MID)                                                                                     \
            {                                                                                                       \
                u32 w[8];                                                                                           \
                {                                                                                                   \
                    /* units u .. u + 8 of the 16 the lane holds, u = (Bv >> 4) & 7 */                              \
                    const u32 a0 = b2 ? LO.hi.x : LO.lo.x, a1 = b2 ? LO.hi.y : LO.lo.y, a2 = b2 ? LO.hi.z : LO.lo.z;   \
                    const u32 a3 = b2 ? LO.hi.w : LO.lo.w, a4 = b2 ? UP.lo.x : LO.hi.x, a5 = b2 ? UP.lo.y : LO.hi.y;   \
                    const u32 a6 = b2 ? UP.lo.z : LO.hi.z, a7 = b2 ? UP.lo.w : LO.hi.w, a8 = b2 ? UP.hi.x : UP.lo.x;   \
                    const u32 a9 = b2 ? UP.hi.y : UP.lo.y, a10 = b2 ? UP.hi.z : UP.lo.z, a11 = b2 ? UP.hi.w : UP.lo.w; \
                    const u32 c0 = b1 ? a2 : a0, c1 = b1 ? a3 : a1, c2 = b1 ? a4 : a2, c3 = b1 ? a5 : a3, c4 = b1 ? a6 : a4; \
                    const u32 c5 = b1 ? a7 : a5, c6 = b1 ? a8 : a6, c7 = b1 ? a9 : a7, c8 = b1 ? a10 : a8, c9 = b1 ? a11 : a9; \
                    const u32 y0 = b0s ? c1 : c0, y1 = b0s ? c2 : c1, y2 = b0s ? c3 : c2, y3 = b0s ? c4 : c3, y4 = b0s ? c5 : c4; \
                    const u32 y5 = b0s ? c6 : c5, y6 = b0s ? c7 : c6, y7 = b0s ? c8 : c7, y8 = b0s ? c9 : c8;      \
                    w[0] = __funnelshift_l(y1, y0, xs); w[1] = __funnelshift_l(y2, y1, xs);                         \
                    w[2] = __funnelshift_l(y3, y2, xs); w[3] = __funnelshift_l(y4, y3, xs);                         \
                    w[4] = __funnelshift_l(y5, y4, xs); w[5] = __funnelshift_l(y6, y5, xs);                         \
                    w[6] = __funnelshift_l(y7, y6, xs); w[7] = __funnelshift_l(y8, y7, xs);                         \
                }                                                                                                   \
                if (s + 2 < rounds) {      /* the oct that just died is replaced by the next one of the walk */     \
                    const u8 *ad = base + (u32)min(max(onext, 0), olim);                                            \
                    ldg256_if(LO, ad, strand ^ 1u);                                                                 \
                    ldg256_if(UP, ad, strand);                                                                      \
                    onext += ostep;                                                                                 \
                }                                                                                                   \
                CK_S3_HH_LOOP for (u32 hh = 0; hh < 2u && s < rounds; hh++, s++) {                                 \
                    u32 W0, W1, W2, W3;                                                                             \
                    if (hh == 0) { W0 = strand ? w[7] : w[0]; W1 = strand ? w[6] : w[1]; W2 = strand ? w[5] : w[2]; W3 = strand ? w[4] : w[3]; } \
                    else { W0 = strand ? w[3] : w[4]; W1 = strand ? w[2] : w[5]; W2 = strand ? w[1] : w[6]; W3 = strand ? w[0] : w[7]; } \
                    CK_S3_ROUND(MID)                                                                                \
                }                                                                                                   \
                if (s >= rounds) break;                                                                             \
            }
            const u32 rounds = __reduce_max_sync(CK_FULL, (nchunks + 3) >> 2);
            if (rounds) {
                u32 s = 0;
                if (any_mid) {
#pragma unroll 1
                    for (;;) {
                        CK_S3_PAIR(R0, R1, true)
                        CK_S3_PAIR(R1, R0, true)
                    }
                } else {
#pragma unroll 1
                    for (;;) {
                        CK_S3_PAIR(R0, R1, false)
                        CK_S3_PAIR(R1, R0, false)
                    }
                }
            }
#undef CK_S3_PAIR
#undef CK_S3_ROUND
#undef CK_S3_STORE4
            if (want_hash) {
                // last stripe: canonical bytes [n - 64, n) = D[p0 + n - 64 ..) forward, the mirror of D[p0 + 16 ..) reverse
                const u32 Bl = strand ? (u32)p0 + 16u : (u32)p0 + nn - 64u;
                const u8 *ad = base + ((Bl >> 6) << 4);
                const uint4 XA = ldg128_here(ad), XB = ldg128_here(ad + 16);
                const u32 xa = (Bl >> 4) & 3u, xl = 2u * Bl;
                const bool a2 = (xa & 2u) != 0, a1 = (xa & 1u) != 0;
                const u32 y0 = a2 ? XA.z : XA.x, y1 = a2 ? XA.w : XA.y, y2 = a2 ? XB.x : XA.z, y3 = a2 ? XB.y : XA.w;
                const u32 y4 = a2 ? XB.z : XB.x, y5 = a2 ? XB.w : XB.y;
                const u32 u0 = a1 ? y1 : y0, u1 = a1 ? y2 : y1, u2 = a1 ? y3 : y2, u3 = a1 ? y4 : y3, u4 = a1 ? y5 : y4;
                const u32 w0 = __funnelshift_l(u1, u0, xl), w1 = __funnelshift_l(u2, u1, xl);
                const u32 w2 = __funnelshift_l(u3, u2, xl), w3 = __funnelshift_l(u4, u3, xl);
                const u32 W0 = strand ? w3 : w0, W1 = strand ? w2 : w1, W2 = strand ? w1 : w2, W3 = strand ? w0 : w3;
                {
                    const uint4 v0 = t2_ascii16(__funnelshift_l(W0, W0, rot), T, sa, sb);
                    t2_acc16(acc0, acc1, v0, c_lastsec[0], c_lastsec[1]);
                    const uint4 v1 = t2_ascii16(__funnelshift_l(W1, W1, rot), T, sa, sb);
                    t2_acc16(acc2, acc3, v1, c_lastsec[2], c_lastsec[3]);
                    const uint4 v2 = t2_ascii16(__funnelshift_l(W2, W2, rot), T, sa, sb);
                    t2_acc16(acc4, acc5, v2, c_lastsec[4], c_lastsec[5]);
                    const uint4 v3 = t2_ascii16(__funnelshift_l(W3, W3, rot), T, sa, sb);
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
                // the last 16 canonical bytes [n - 16, n)
                const u32 q = strand ? (u32)p0 + 16u : (u32)p0 + nn - 16u;
                const u8 *ad = base + ((q >> 4) << 2);
                const u32 w = __funnelshift_l(ldg32(ad + 4), ldg32(ad), 2u * q);
                const uint4 vv = t2_ascii16(__funnelshift_l(w, w, rot), T, sa, sb);
                const u64 lo = ((u64)vv.y << 32) | vv.x, hi = ((u64)vv.w << 32) | vv.z;
                const u64 tail = mul128_fold64(lo ^ c_midsec[14], hi ^ c_midsec[15]);
                if (nbr) h = xxh3_avalanche(xxh3_avalanche((u64)nn * CK_P64_1 + mida) + midb + tail);
            }
#undef CK_T2_SCR
        }
        if (fast) {
            const u32 start = os >> 1, strand = os & 1u;
            a.out_start[rec] = strand ? (n - 1 - start) : start;
            a.out_strand[rec] = (u8)strand;
            if (want_hash) a.out_hash[rec] = h;
        } else if (in_class) {
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
