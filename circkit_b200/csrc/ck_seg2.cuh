// ck_seg2.cuh -- long 2-bit records (plasmid / mtDNA lengths, 8 k - 426 k bases: config 4 of BASELINE.json): one warp per
// record, every LANE walks its own SEGMENT of it with the streaming code of the lane-per-record kernel (ck_stream3.cuh).
//
// A lane per record is the right shape up to a few kb (ck_stream3.cuh); beyond that a single lane would walk 50 KB and a warp's
// 32 records would keep it busy for milliseconds.  A CTA per record (k_canon_cta) stages both strands in shared memory and
// runs every phase behind block-wide barriers: 15 % of HBM.  Here the record is cut into 32 segments instead, so that each
// lane streams ~1/32 of it straight from the (doubled, 32-byte-aligned) arena, exactly like a circRNA-length record of its own:
//   * scan      : lane l takes the octs [l * per, (l + 1) * per) (4 steps of 32 rotations each).  16-mer keys (32 bits: with
//                 8-mers every long record would tie -- 2 n rotations against 65 536 keys), 30 funnel shifts + 16 VIMNMX3 per
//                 step and strand, the reverse strand from the block reverse-complemented in registers; every lane keeps its
//                 two smallest keys and the step / strand of the smallest.  The one partial step of a record is scanned with
//                 its positions past n masked.  A warp reduction finds the record's minimum; two rotations with the same
//                 16-mer (tandem repeats, poly-A: the adversarial 1 %) send the record to the CTA kernel's duel path;
//   * locate    : the winning step is replayed (by every lane, uniformly);
//   * emit      : the canonical string is cut into 1024-byte blocks (16 XXH3 stripes); lane l generates blocks
//                 [l * bl, (l + 1) * bl) with the two-oct window walk of ck_stream3.cuh (a linear window either strand, thanks
//                 to the doubled record) and sends them through the same conflict-free shared-memory stage;
//   * XXH3-64   : a lane cannot scramble (the accumulators of block b feed block b + 1), but what a block ADDS to the
//                 accumulators depends on its bytes only: every lane stores the eight partial sums of each of its blocks, and
//                 lanes 0..7 then chain acc = scramble(acc + sum_b) over the blocks -- one accumulator each.
#pragma once
#include "ck_stream3.cuh"

namespace ck {

#define CK_SEG_WARPS 8u
#define CK_SEG_WARP_BYTES CK_S3_STAGE
#define CK_SEG_SCRATCH_U64 (8u * 448u)       // per warp: partial sums of up to 447 blocks (n <= 425 984) + the leftover

// minimum 16-mer over the 32 rotations that start in units (x0, x1)
__device__ __forceinline__ u32 seg_step_min32(u32 x0, u32 x1, u32 x2)
{
    u32 m = min(x0, x1);
#pragma unroll
    for (int i = 1; i < 16; i += 2) {
        const u32 a = __funnelshift_l(x1, x0, 2 * i), b = i + 1 < 16 ? __funnelshift_l(x1, x0, 2 * i + 2) : a;
        m = min(min(m, a), b);
    }
#pragma unroll
    for (int i = 1; i < 16; i += 2) {
        const u32 a = __funnelshift_l(x2, x1, 2 * i), b = i + 1 < 16 ? __funnelshift_l(x2, x1, 2 * i + 2) : a;
        m = min(min(m, a), b);
    }
    return m;
}
// the same with a validity mask (bit s: rotation s of the step counts)
__device__ __forceinline__ u32 seg_step_min32_masked(u32 x0, u32 x1, u32 x2, u32 valid)
{
    u32 m = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const u32 a = i ? __funnelshift_l(x1, x0, 2 * i) : x0, b = i ? __funnelshift_l(x2, x1, 2 * i) : x1;
        if ((valid >> i) & 1u) m = min(m, a);
        if ((valid >> (16 + i)) & 1u) m = min(m, b);
    }
    return m;
}
__device__ __forceinline__ void seg_track(u32 &m1k, u32 &m1t, u32 &m2k, u32 k, u32 tag)
{
    const bool lt = k < m1k;
    m2k = lt ? m1k : min(m2k, k);
    m1t = lt ? tag : m1t;
    m1k = min(m1k, k);
}

template <int V>
__global__ void __launch_bounds__(32 * CK_SEG_WARPS, 2) k_canon_seg(CanonArgs a, u64 *seg_scratch)
{
    extern __shared__ __align__(16) u32 smem[];
    constexpr bool want_hash = (V & CK_W2_HASH) != 0, want_out = (V & CK_W2_OUT) != 0;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 aux = (u32)__cvta_generic_to_shared(smem) + wid * CK_SEG_WARP_BYTES;
    u64 *S = seg_scratch + (size_t)(blockIdx.x * CK_SEG_WARPS + wid) * CK_SEG_SCRATCH_U64;
    const u32 count = *a.count;
    a.list += a.count[16];
    const u8 *arena = reinterpret_cast<const u8 *>(a.packed2);
    const u32 st_w = aux + 128u * (lane >> 1);
    const u32 st_j = 4u * (lane & 1u) + ((lane >> 1) & 3u);
    const u32 st_w0 = st_w + 16u * (st_j & 7u), st_w1 = st_w + 16u * ((st_j + 1u) & 7u);
    const u32 st_w2 = st_w + 16u * ((st_j + 2u) & 7u), st_w3 = st_w + 16u * ((st_j + 3u) & 7u);
    const u32 st_r = aux + 128u * (lane >> 3) + 16u * (((lane & 3u) + 4u * ((lane >> 2) & 1u) + ((lane >> 3) & 3u)) & 7u);
    const u64 *sec = reinterpret_cast<const u64 *>(c_secret);

    for (;;) {
        u32 e = 0;
        if (lane == 0) e = atomicAdd(a.retry_counts + 29, 1u);
        e = __shfl_sync(CK_FULL, e, 0);
        if (e >= count) break;
        const u32 rec = a.list[e];
        const u64 off = a.offsets[rec];
        const u32 n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
        const u8 *base = arena + 8ull * p2_word(off, rec, 1u);
        u8 *dst = want_out ? a.out + out_byte(off, rec) : nullptr;

        // ---- scan: lane l owns the octs [o0, o1) = the steps [4 o0, 4 o1); step S1 is the record's partial one
        const u32 S1 = ((n + 31) >> 5) - 1;
        const u32 octs = (S1 + 4) >> 2;                            // octs that hold a step <= S1
        const u32 per = (octs + 31) >> 5;
        const u32 o0 = min(lane * per, octs), o1 = min(o0 + per, octs);
        u32 m1k = 0xffffffffu, m1t = 0, m2k = 0xffffffffu;
        if (o0 < o1) {
            Oct OA = ldg256_here(base + 32 * o0), OB = ldg256_here(base + 32 * (o0 + 1)), OC = ldg256_here(base + 32 * (o0 + 2));
            u32 r0 = w2_revcomp(OA.lo.x);
#pragma unroll 1
            for (u32 j = o0; j < o1; j++) {
                const Oct ON = ldg256_here(base + 32 * min(j + 3, octs + 1));
                if ((j & 1u) == 0) prefetch_l2(base + 32 * min(j + 8, octs + 1));
                const u32 r1 = w2_revcomp(OA.lo.y), r2 = w2_revcomp(OA.lo.z), r3 = w2_revcomp(OA.lo.w);
                const u32 r4 = w2_revcomp(OA.hi.x), r5 = w2_revcomp(OA.hi.y), r6 = w2_revcomp(OA.hi.z);
                const u32 r7 = w2_revcomp(OA.hi.w), r8 = w2_revcomp(OB.lo.x);
                const u32 t = 4 * j;
#define CK_SEG_STEP(tt, x0, x1, x2, q0, q1, q2)                                                                      \
                if ((tt) < S1) {                                                                                     \
                    seg_track(m1k, m1t, m2k, seg_step_min32(x0, x1, x2), 2 * (tt));                                  \
                    seg_track(m1k, m1t, m2k, seg_step_min32(q2, q1, q0), 2 * (tt) + 1);                              \
                }
                CK_SEG_STEP(t, OA.lo.x, OA.lo.y, OA.lo.z, r0, r1, r2)
                CK_SEG_STEP(t + 1, OA.lo.z, OA.lo.w, OA.hi.x, r2, r3, r4)
                CK_SEG_STEP(t + 2, OA.hi.x, OA.hi.y, OA.hi.z, r4, r5, r6)
                CK_SEG_STEP(t + 3, OA.hi.z, OA.hi.w, OB.lo.x, r6, r7, r8)
#undef CK_SEG_STEP
                r0 = r8;
                OA = OB; OB = OC; OC = ON;
            }
        }
        if (lane == 0) {
            // the record's partial step S1 (kept out of the loop: it is one step per record): forward rotations s < lim count,
            // reverse ones s >= 32 - lim -- the others repeat positions the full steps have seen
            const uint2 x01 = ldg64(base + 8 * S1);
            const uint2 x23 = ldg64(base + 8 * S1 + 8);
            const u32 lim = n - 32 * S1;
            const u32 vf = lim >= 32 ? 0xffffffffu : (1u << lim) - 1u;
            const u32 vr = lim >= 32 ? 0xffffffffu : 0xffffffffu << (32 - lim);
            seg_track(m1k, m1t, m2k, seg_step_min32_masked(x01.x, x01.y, x23.x, vf), 2 * S1);
            seg_track(m1k, m1t, m2k, seg_step_min32_masked(w2_revcomp(x23.x), w2_revcomp(x01.y), w2_revcomp(x01.x), vr), 2 * S1 + 1);
        }
        __syncwarp();
        // ---- the record's minimum; a second rotation with the same 16-mer anywhere sends the record to the duel path
        const u32 kmin = __reduce_min_sync(CK_FULL, m1k);
        const u32 who = __ballot_sync(CK_FULL, m1k == kmin);
        const u32 wl = __ffs(who) - 1;
        const u32 second = __reduce_min_sync(CK_FULL, lane == wl ? m2k : m1k);
        const u32 wtag = __shfl_sync(CK_FULL, m1t, wl);
        bool ok = second != kmin;
        u32 os = 0;
        {   // locate: replay the winning step (uniform over the warp)
            const u32 t = wtag >> 1, strand = wtag & 1u;
            const uint2 x01 = ldg64(base + 8 * t);
            const u32 x2 = ldg32(base + 8 * t + 8);
            const u32 y0 = strand ? w2_revcomp(x2) : x01.x, y1 = strand ? w2_revcomp(x01.y) : x01.y;
            const u32 y2 = strand ? w2_revcomp(x01.x) : x2;
            u32 match = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const u32 wa = i ? __funnelshift_l(y1, y0, 2 * i) : y0, wb = i ? __funnelshift_l(y2, y1, 2 * i) : y1;
                match |= (wa == kmin ? 1u : 0u) << i;
                match |= (wb == kmin ? 1u : 0u) << (16 + i);
            }
            const int lim = (int)n - 32 * (int)t;
            const u32 valid = strand ? (lim >= 32 ? 0xffffffffu : 0xffffffffu << (32 - lim))
                                     : (lim >= 32 ? 0xffffffffu : (1u << lim) - 1u);
            const u32 hits = match & valid;
            const u32 s = __ffs(hits) - 1;
            int st = strand ? (int)n - 48 - 32 * (int)t + (int)s : 32 * (int)t + (int)s;
            if (st < 0) st += (int)n;
            os = ((u32)st << 1) | strand;
            if (__popc(hits) != 1) ok = false;
        }
        if (!ok) {
            if (lane == 0) {
                const int c = n <= cls_max_n(CLS_C2A) ? CLS_C2A : CLS_C2B;
                const u32 k = atomicAdd(a.retry_counts + c, 1u);
                a.retry[a.retry_counts[16 + c] + k] = rec;
            }
            __syncwarp();
            continue;
        }
        // ---- emit: lane l generates the chunks [c0, c1) of the canonical string (whole 1024-byte blocks)
        u64 h = 0;
        if (want_out || want_hash) {
            const u32 strand = os & 1u;
            const u32 T = strand ? 0x41434754u : 0x54474341u;
            const u32 sa = strand ? 0x5140u : 0x2637u, sb = strand ? 0x7362u : 0x0415u, rot = strand ? 16u : 0u;
            int p0 = strand ? (int)n - 16 - (int)(os >> 1) : (int)(os >> 1);
            if (p0 < 0) p0 += (int)n;
            const u32 C = (n + 15) >> 4, nblk = (C + 63) >> 6, bl = (nblk + 31) >> 5;
            const u32 c0 = min(64u * bl * lane, C), c1 = min(c0 + 64u * bl, C);
            const u32 nfull = (n - 1) >> 6, nb_full = nfull >> 4;   // stripes / whole blocks the stripe loop hashes
            const u32 s0 = c0 >> 2;                                 // first stripe of this lane
            const u32 my_rounds = (c1 - c0 + 3) >> 2;
            const u32 rounds = __reduce_max_sync(CK_FULL, my_rounds);
            // window base of iteration 0, shifted up by one oct so that it never goes negative (the last lanes of a reverse
            // strand reach below the record's start with windows that belong to no chunk of theirs)
            const int Bvi = strand ? p0 + (int)n - 112 - 16 * (int)c0 : p0 + 16 * (int)c0;
            const u32 Bm = (u32)(Bvi + 128);
            const u32 xs = 2u * (Bm & 15u);
            const bool b2 = (Bm & 64u) != 0, b1 = (Bm & 32u) != 0, b0s = (Bm & 16u) != 0;
            const int ostep = strand ? -32 : 32;
            const int olim = (int)(((2u * n + 127u) >> 7) << 5);
            const int ob = (int)((Bm >> 7) << 5) - 32;              // byte offset of the window's lower oct
            int onext = strand ? ob - 32 : ob + 64;
            Oct R0 = ldg256_here(base + (u32)min(max(ob, 0), olim)), R1 = ldg256_here(base + (u32)min(max(ob + 32, 0), olim));
            u64 acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, acc6 = 0, acc7 = 0;
            u32 og[4] = {0, 0, 0, 0}; int orem[4] = {0, 0, 0, 0};
            if (want_out) {
                const u32 gr = (u32)((dst - a.out) >> 4) + c0;
                const u32 nch = c1 - c0;
#pragma unroll
                for (u32 i = 0; i < 4; i++) {
                    og[i] = __shfl_sync(CK_FULL, gr, 8 * i + (lane >> 2)) + (lane & 3u);
                    orem[i] = (int)__shfl_sync(CK_FULL, nch, 8 * i + (lane >> 2)) - (int)(lane & 3u);
                }
            }
            if (want_hash && lane < 8) S[8u * nb_full + lane] = 0;  // the leftover entry, in case no lane owns leftover stripes
            __syncwarp();
            u32 s = 0;                                              // local round = local stripe
#define CK_SEG_ROUND                                                                                                \
            {                                                                                                       \
                uint4 v[4];                                                                                         \
                v[0] = t2_ascii16(__funnelshift_l(W0, W0, rot), T, sa, sb);                                         \
                v[1] = t2_ascii16(__funnelshift_l(W1, W1, rot), T, sa, sb);                                         \
                v[2] = t2_ascii16(__funnelshift_l(W2, W2, rot), T, sa, sb);                                         \
                v[3] = t2_ascii16(__funnelshift_l(W3, W3, rot), T, sa, sb);                                         \
                const u32 sg = s0 + s;                     /* global stripe */                                      \
                if (want_hash && s < my_rounds && sg < nfull) {                                                     \
                    const u32 ks = sg & 15u;                                                                        \
                    t2_acc16(acc0, acc1, v[0], sec[ks + 0], sec[ks + 1]);                                           \
                    t2_acc16(acc2, acc3, v[1], sec[ks + 2], sec[ks + 3]);                                           \
                    t2_acc16(acc4, acc5, v[2], sec[ks + 4], sec[ks + 5]);                                           \
                    t2_acc16(acc6, acc7, v[3], sec[ks + 6], sec[ks + 7]);                                           \
                    if (ks == 15u || sg + 1 == nfull) {    /* a block is complete, or the record's leftover stripes end here */ \
                        u64 *Sb = S + 8u * (sg >> 4);                                                               \
                        Sb[0] = acc0; Sb[1] = acc1; Sb[2] = acc2; Sb[3] = acc3; Sb[4] = acc4; Sb[5] = acc5; Sb[6] = acc6; Sb[7] = acc7; \
                        acc0 = acc1 = acc2 = acc3 = acc4 = acc5 = acc6 = acc7 = 0;                                  \
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
            }
#define CK_SEG_PAIR(LO, UP)                                                                                         \
            {                                                                                                       \
                u32 w[8];                                                                                           \
                {                                                                                                   \
                    const u32 a0 = b2 ? LO.hi.x : LO.lo.x, a1 = b2 ? LO.hi.y : LO.lo.y, a2 = b2 ? LO.hi.z : LO.lo.z;   \
                    const u32 a3 = b2 ? LO.hi.w : LO.lo.w, a4 = b2 ? UP.lo.x : LO.hi.x, a5 = b2 ? UP.lo.y : LO.hi.y;   \
                    const u32 a6 = b2 ? UP.lo.z : LO.hi.z, a7 = b2 ? UP.lo.w : LO.hi.w, a8 = b2 ? UP.hi.x : UP.lo.x;   \
                    const u32 a9 = b2 ? UP.hi.y : UP.lo.y, a10 = b2 ? UP.hi.z : UP.lo.z, a11 = b2 ? UP.hi.w : UP.lo.w; \
                    const u32 d0 = b1 ? a2 : a0, d1 = b1 ? a3 : a1, d2 = b1 ? a4 : a2, d3 = b1 ? a5 : a3, d4 = b1 ? a6 : a4; \
                    const u32 d5 = b1 ? a7 : a5, d6 = b1 ? a8 : a6, d7 = b1 ? a9 : a7, d8 = b1 ? a10 : a8, d9 = b1 ? a11 : a9; \
                    const u32 y0 = b0s ? d1 : d0, y1 = b0s ? d2 : d1, y2 = b0s ? d3 : d2, y3 = b0s ? d4 : d3, y4 = b0s ? d5 : d4; \
                    const u32 y5 = b0s ? d6 : d5, y6 = b0s ? d7 : d6, y7 = b0s ? d8 : d7, y8 = b0s ? d9 : d8;      \
                    w[0] = __funnelshift_l(y1, y0, xs); w[1] = __funnelshift_l(y2, y1, xs);                         \
                    w[2] = __funnelshift_l(y3, y2, xs); w[3] = __funnelshift_l(y4, y3, xs);                         \
                    w[4] = __funnelshift_l(y5, y4, xs); w[5] = __funnelshift_l(y6, y5, xs);                         \
                    w[6] = __funnelshift_l(y7, y6, xs); w[7] = __funnelshift_l(y8, y7, xs);                         \
                }                                                                                                   \
                if (s + 2 < rounds) {                                                                               \
                    const u8 *ad = base + (u32)min(max(onext, 0), olim);                                            \
                    ldg256_if(LO, ad, strand ^ 1u);                                                                 \
                    ldg256_if(UP, ad, strand);                                                                      \
                    onext += ostep;                                                                                 \
                }                                                                                                   \
                _Pragma("unroll 1") for (u32 hh = 0; hh < 2u && s < rounds; hh++, s++) {                            \
                    u32 W0, W1, W2, W3;                                                                             \
                    if (hh == 0) { W0 = strand ? w[7] : w[0]; W1 = strand ? w[6] : w[1]; W2 = strand ? w[5] : w[2]; W3 = strand ? w[4] : w[3]; } \
                    else { W0 = strand ? w[3] : w[4]; W1 = strand ? w[2] : w[5]; W2 = strand ? w[1] : w[6]; W3 = strand ? w[0] : w[7]; } \
                    CK_SEG_ROUND                                                                                    \
                }                                                                                                   \
                if (s >= rounds) break;                                                                             \
            }
            if (rounds) {
#pragma unroll 1
                for (;;) {
                    CK_SEG_PAIR(R0, R1)
                    CK_SEG_PAIR(R1, R0)
                }
            }
#undef CK_SEG_PAIR
#undef CK_SEG_ROUND
            if (want_hash) {
                __syncwarp();
                // chain the blocks: lane i < 8 owns accumulator i
                u64 acc = lane == 0 ? CK_P32_3 : lane == 1 ? CK_P64_1 : lane == 2 ? CK_P64_2 : lane == 3 ? CK_P64_3
                        : lane == 4 ? CK_P64_4 : lane == 5 ? CK_P32_2 : lane == 6 ? CK_P64_5 : CK_P32_1;
                if (lane < 8) {
                    const u64 sk = sec[16 + lane];
                    for (u32 b = 0; b < nb_full; b++) {
                        acc += S[8u * b + lane];
                        acc ^= acc >> 47; acc ^= sk; acc *= CK_P32_1;
                    }
                    acc += S[8u * nb_full + lane];
                }
                // last stripe: canonical bytes [n - 64, n): its four chunks, one per accumulator pair (lanes 0, 2, 4, 6 apply it)
                const u32 Bl = strand ? (u32)p0 + 16u : (u32)p0 + n - 64u;
                const u8 *ad = base + ((Bl >> 6) << 4);
                const uint4 XA = ldg128_here(ad), XB = ldg128_here(ad + 16);
                const u32 xa = (Bl >> 4) & 3u, xl = 2u * Bl;
                const bool q2 = (xa & 2u) != 0, q1 = (xa & 1u) != 0;
                const u32 y0 = q2 ? XA.z : XA.x, y1 = q2 ? XA.w : XA.y, y2 = q2 ? XB.x : XA.z, y3 = q2 ? XB.y : XA.w;
                const u32 y4 = q2 ? XB.z : XB.x, y5 = q2 ? XB.w : XB.y;
                const u32 u0 = q1 ? y1 : y0, u1 = q1 ? y2 : y1, u2 = q1 ? y3 : y2, u3 = q1 ? y4 : y3, u4 = q1 ? y5 : y4;
                const u32 w0 = __funnelshift_l(u1, u0, xl), w1 = __funnelshift_l(u2, u1, xl);
                const u32 w2 = __funnelshift_l(u3, u2, xl), w3 = __funnelshift_l(u4, u3, xl);
                const u32 Wl[4] = {strand ? w3 : w0, strand ? w2 : w1, strand ? w1 : w2, strand ? w0 : w3};
                // every lane needs the accumulators of lanes 0..7 for the merge: gather them
                u64 ac[8];
#pragma unroll
                for (int i = 0; i < 8; i++) ac[i] = __shfl_sync(CK_FULL, acc, i);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 vv = t2_ascii16(__funnelshift_l(Wl[k], Wl[k], rot), T, sa, sb);
                    t2_acc16(ac[2 * k], ac[2 * k + 1], vv, c_lastsec[2 * k], c_lastsec[2 * k + 1]);
                }
                u64 r = (u64)n * CK_P64_1;
#pragma unroll
                for (int k = 0; k < 4; k++) r += mul128_fold64(ac[2 * k] ^ c_mergesec[2 * k], ac[2 * k + 1] ^ c_mergesec[2 * k + 1]);
                h = xxh3_avalanche(r);
            }
        }
        if (lane == 0) {
            const u32 start = os >> 1, strand = os & 1u;
            a.out_start[rec] = strand ? (n - 1 - start) : start;
            a.out_strand[rec] = (u8)strand;
            if (want_hash) a.out_hash[rec] = h;
        }
        __syncwarp();
    }
}

}  // namespace ck
