// ck_record.cuh -- one circular record -> (start, strand) of its canonical form, canonical ASCII,
// XXH3-64 of the canonical ASCII.  Written once for both thread shapes (ck_group.cuh).
//
// What it replaces in the reference (per record):
//   lmsr_index x2 + revcomp + compare      lib/src/canonicalize.rs:5-36, :41-47, :54-63
//   xxh3_64(canonical)                     src/uniq.rs:45
//
// Algorithm (not Duval -- a data-parallel restatement with the same answer):
//   1. stage the record as packed units in shared memory, forward strand and reverse complement,
//      each extended circularly so windows never need wrap logic;
//   2. every thread scans a strided share of the 2n candidate rotations (n per strand), keeping
//      the minimum K-symbol prefix key; a group-wide min gives the minimal key `gmin`;
//   3. if exactly one candidate carries gmin it is the answer (its rotation is strictly smaller
//      than every other rotation of either strand);
//   4. otherwise (repeats, palindromic circles, multimers) the tied candidates of each strand are
//      reduced by duels:  for i < j with LCP(rot_i, rot_j) >= j - i  candidate j can be dropped
//      (it is either not minimal or an equal rotation with a larger index -- lmsr_index returns
//      the smallest index), else the first mismatch decides.  A final whole-string
//      comparison picks the strand; equality keeps the reverse complement exactly like
//      `if lmsr_s < lmsr_revcomp_s {lmsr_s} else {lmsr_revcomp_s}` (lib/src/canonicalize.rs:58-62).
#pragma once
#include "ck_group.cuh"

namespace ck {

template <int BITS> struct Lane {
    static constexpr int S = 32 / BITS;                 // symbols per 32-bit unit
    static constexpr int SU = (BITS == 8) ? 4 : 8;      // candidate positions per scan step
    static constexpr int EXTU = (BITS == 8) ? 8 : 4;    // circular extension, in units
    static constexpr bool KEY64 = (BITS != 2);          // key width: 32 bits (16 bases) or 64 bits
    static constexpr int K = KEY64 ? 2 * S : S;         // symbols in a key (16, 16, 8)
    static constexpr int SMALL_N = (EXTU + 4) * S;      // below this the extension is built symbol-wise
};
template <int BITS> struct KeyOf { typedef u64 type; };
template <> struct KeyOf<2> { typedef u32 type; };

// Units one strand occupies in shared memory for a record of n symbols.
template <int BITS> __host__ __device__ constexpr u32 strand_units(u32 n)
{
    return n / Lane<BITS>::S + 2 + Lane<BITS>::EXTU;
}

// Global scratch (u32 words) the tie path needs for a record of n symbols: one bitmap + two lists.
template <int BITS> __host__ __device__ constexpr u64 tie_scratch_words(u64 n)
{
    return (n + 31) / 32 + 2 * (n / (Lane<BITS>::K + 1) + 2) + 8;
}

struct RecordIn {
    const u64 *packed2;   // BITS == 2: this record's 16-base u32 units (first base in the top bits), 8-byte aligned
    const u8 *bytes;      // BITS != 2: normalised bytes of this record (already offset)
    u32 n;
};

struct RecordOut {
    u32 start;
    u32 strand;
};

// ---------------------------------------------------------------------------------------------
template <int BITS> __device__ __forceinline__ u32 window32(const u32 *X, u32 pos)
{
    constexpr int S = Lane<BITS>::S;
    u32 j = pos / S, s = (pos % S) * BITS;
    return funnel_l(X[j], X[j + 1], s);
}
template <int BITS> __device__ __forceinline__ typename KeyOf<BITS>::type key_at(const u32 *X, u32 pos)
{
    constexpr int S = Lane<BITS>::S;
    u32 j = pos / S, s = (pos % S) * BITS;
    u32 a = X[j], b = X[j + 1];
    if (!Lane<BITS>::KEY64) return (typename KeyOf<BITS>::type)funnel_l(a, b, s);
    u32 c = X[j + 2];
    return (typename KeyOf<BITS>::type)(((u64)funnel_l(a, b, s) << 32) | funnel_l(b, c, s));
}

// symbol t (0 <= t < n) of the forward strand straight from global memory (small-n extension only)
template <int BITS> __device__ __forceinline__ u32 sym_global(const RecordIn &r, u32 t)
{
    if (BITS == 2) return (reinterpret_cast<const u32 *>(r.packed2)[t >> 4] >> (30 - 2 * (t & 15))) & 3u;
    u32 b = r.bytes[t];
    return BITS == 4 ? (u32)s_tab.code4[b] : b;
}

// ---------------------------------------------------------------------------------------------
// Stage forward strand Xf and reverse complement Xr.
template <int BITS, typename G>
__device__ __forceinline__ void stage_record(const RecordIn &r, u32 *Xf, u32 *Xr)
{
    constexpr int S = Lane<BITS>::S;
    constexpr int EXTU = Lane<BITS>::EXTU;
    const u32 n = r.n, rank = G::rank(), gs = G::size();
    const u32 jn = n / S, rem = n % S;
    const u32 xunits = jn + 1 + EXTU;

    // real symbols
    if (BITS == 2) {
        const u32 W = (n + 31) >> 5;
        for (u32 k = rank; k < W; k += gs) {
            u64 w = __ldg(r.packed2 + k);           // two 16-base units, lower address first
            Xf[2 * k] = (u32)w;
            Xf[2 * k + 1] = (u32)(w >> 32);
        }
    } else {
        const u32 NU = (n + S - 1) / S;
        for (u32 j = rank; j < NU; j += gs) {
            u32 v = 0;
#pragma unroll
            for (int i = 0; i < S; i++) {
                u32 t = j * S + i;
                u32 c = 0;
                if (t < n) { u32 b = __ldg(r.bytes + t); c = BITS == 4 ? (u32)s_tab.code4[b] : b; }
                v = (v << BITS) | c;
            }
            Xf[j] = v;
        }
    }
    G::sync();
    // circular extension: units jn .. xunits-1
    if (n >= (u32)Lane<BITS>::SMALL_N) {
        for (u32 j = jn + rank; j < xunits; j += gs) {
            u32 val;
            if (j == jn) {
                u32 g0 = funnel_l(Xf[0], Xf[1], 0);
                u32 top = rem ? (Xf[jn] & ~(0xffffffffu >> (rem * BITS))) : 0u;
                val = rem ? (top | (g0 >> (rem * BITS))) : g0;
            } else {
                u32 t = j * S - n;
                val = window32<BITS>(Xf, t);
            }
            // all reads of this pass touch units < jn only (n >= SMALL_N), except Xf[jn] by its owner
            Xf[j] = val;
        }
    } else {
        for (u32 j = jn + rank; j < xunits; j += gs) {
            u32 v = 0;
            for (int i = 0; i < S; i++) v = (v << BITS) | sym_global<BITS>(r, (j * S + i) % n);
            Xf[j] = v;
        }
    }
    G::sync();
    // reverse complement: rc symbols [jS, jS+S) = revcomp(forward window at n - (j+1)S  (mod n))
    for (u32 j = rank; j < xunits; j += gs) {
        int t = (int)n - (int)((j + 1) * S);          // n <= 2^30: fits
        if (t < 0) { t += (int)n; if (t < 0) { t %= (int)n; if (t < 0) t += (int)n; } }
        Xr[j] = revcomp_unit<BITS>(window32<BITS>(Xf, (u32)t));
    }
    G::sync();
}

// ---------------------------------------------------------------------------------------------
// Duel of two tied candidates a < b of the same strand (see file header).  Returns the survivor.
template <int BITS> __device__ __forceinline__ u32 duel(const u32 *X, u32 n, u32 a, u32 b)
{
    constexpr int S = Lane<BITS>::S;
    const u32 d = b - a;
    u32 pa = a, pb = b;
    for (u32 off = 0; off < d; off += S) {
        u32 wa = window32<BITS>(X, pa), wb = window32<BITS>(X, pb);
        if (wa != wb) {
            u32 ds = __clz(wa ^ wb) / BITS;
            if (off + ds < d) return wa < wb ? a : b;
            return a;                                   // agree on the first d symbols
        }
        pa += S; while (pa >= n) pa -= n;
        pb += S; while (pb >= n) pb -= n;
    }
    return a;
}

// The same duel by a whole warp (all lanes call it with the same a, b): lane l compares symbols [off + l S, off + (l + 1) S),
// the first lane that sees a difference decides.  The late rounds of the tournament have few pairs that lie far apart --
// in a periodic record they agree on all of the b - a symbols between them -- and one thread walking 16 symbols per step
// (while the rest of the group waits at the barrier) was 70 % of the CTA kernel's time on config 4's repeats.
template <int BITS> __device__ __forceinline__ u32 duel_warp(const u32 *X, u32 n, u32 a, u32 b)
{
    constexpr int S = Lane<BITS>::S;
    const u32 d = b - a, lane = threadIdx.x & 31u;
    for (u32 off0 = 0; off0 < d; off0 += 32u * S) {
        const u32 off = off0 + lane * S;
        u32 wa = 0, wb = 0;
        if (off < d) {
            u32 pa = a + off, pb = b + off;                 // a, b < n and off < d < n: one wrap at most
            if (pa >= n) pa -= n;
            if (pb >= n) pb -= n;
            wa = window32<BITS>(X, pa); wb = window32<BITS>(X, pb);
        }
        const u32 m = __ballot_sync(CK_FULL, wa != wb);
        if (m) {
            const int l = __ffs(m) - 1;
            const u32 wal = __shfl_sync(CK_FULL, wa, l), wbl = __shfl_sync(CK_FULL, wb, l);
            const u32 ds = __clz(wal ^ wbl) / BITS;
            if (off0 + (u32)l * S + ds < d) return wal < wbl ? a : b;
            return a;                                       // agree on the first d symbols
        }
    }
    return a;
}

// Reduce the tied candidates of one strand to a single survivor.  Returns 0xffffffff if none.
// scr layout: bitmap[(n+31)/32] | listA[cap] | listB[cap]
template <int BITS, typename G>
__device__ u32 strand_tie_winner(const u32 *X, u32 n, typename KeyOf<BITS>::type gmin, u32 *scr, u32 *red)
{
    constexpr int K = Lane<BITS>::K;
    const u32 rank = G::rank(), gs = G::size();
    const u32 nwords = (n + 31) >> 5;
    const u32 cap = n / (K + 1) + 2;
    u32 *bitmap = scr, *listA = scr + nwords, *listB = listA + cap;

    // 1. bitmap of candidates whose key equals gmin
    for (u32 w = rank; w < nwords; w += gs) {
        u32 m = 0;
        u32 base = w << 5;
        for (u32 i = 0; i < 32 && base + i < n; i++)
            if (key_at<BITS>(X, base + i) == gmin) m |= 1u << i;
        bitmap[w] = m;
    }
    G::sync();
    // 2. drop a candidate if another tied candidate lies within the K positions before it
    //    (their keys agree on K >= distance symbols), then compact the rest in position order.
    u32 total_kept = 0;
    for (u32 w0 = 0; w0 < nwords; w0 += gs) {
        u32 w = w0 + rank;
        u32 kept = 0;
        if (w < nwords) {
            u32 cur = bitmap[w], prev = w ? bitmap[w - 1] : 0u;
            u64 v = ((u64)cur << 32) | prev;
            u64 s = v << 1;                         // positions 1 before
            s |= s << 1;                            // 1..2
            s |= s << 2;                            // 1..4
            s |= s << 4;                            // 1..8
            if (K >= 16) s |= s << 8;               // 1..16
            kept = cur & ~(u32)(s >> 32);
        }
        u32 cnt = __popc(kept), tot;
        u32 pos = G::exscan_u32(cnt, tot, red);
        u32 o = total_kept + pos;
        while (kept) {
            u32 b = __ffs(kept) - 1;
            kept &= kept - 1;
            listA[o++] = (w << 5) + b;
        }
        total_kept += tot;
    }
    G::sync();
    if (total_kept == 0) return 0xffffffffu;
    // 3. tournament of duels between neighbours in position order
    u32 m = total_kept;
    u32 *src = listA, *dst = listB;
    while (m > 1) {
        u32 pairs = m >> 1;
        if (pairs <= (gs >> 4)) {                           // at most two pairs per warp: a warp per pair
            for (u32 t = rank >> 5; t < pairs; t += gs >> 5) {
                const u32 w = duel_warp<BITS>(X, n, src[2 * t], src[2 * t + 1]);
                if ((rank & 31u) == 0) dst[t] = w;
            }
        } else
            for (u32 t = rank; t < pairs; t += gs) dst[t] = duel<BITS>(X, n, src[2 * t], src[2 * t + 1]);
        if ((m & 1u) && rank == 0) dst[pairs] = src[m - 1];
        G::sync();
        m = pairs + (m & 1u);
        u32 *tmp = src; src = dst; dst = tmp;
    }
    u32 win = src[0];
    G::sync();
    return win;
}

// Whole-string comparison of forward rotation f against reverse-complement rotation r.
// Returns true when the forward rotation is strictly smaller.
template <int BITS, typename G>
__device__ bool forward_strictly_smaller(const u32 *Xf, const u32 *Xr, u32 n, u32 f, u32 r, u32 *red)
{
    constexpr int S = Lane<BITS>::S;
    const u32 rank = G::rank(), gs = G::size();
    u32 best = 0xffffffffu;
    for (u32 off = rank * S; off < n; off += gs * S) {
        u32 pf = f + off; if (pf >= n) pf -= n;
        u32 pr = r + off; if (pr >= n) pr -= n;
        u32 wa = window32<BITS>(Xf, pf), wb = window32<BITS>(Xr, pr);
        if (wa != wb) {
            u32 ds = __clz(wa ^ wb) / BITS;
            if (off + ds < n) { best = ((off + ds) << 1) | (wa < wb ? 0u : 1u); break; }
        }
    }
    u32 g = G::min_u32(best, red);
    return g != 0xffffffffu && (g & 1u) == 0u;
}

// ---------------------------------------------------------------------------------------------
// LMSR of both strands + strand choice.
template <int BITS, typename G>
__device__ RecordOut canonical_start(const u32 *Xf, const u32 *Xr, u32 n, u32 *scr, u32 *red, bool fwd_only)
{
    typedef typename KeyOf<BITS>::type key_t;
    constexpr int S = Lane<BITS>::S;
    constexpr int SU = Lane<BITS>::SU;
    const u32 rank = G::rank(), gs = G::size();
    const u32 nsu = n / SU;                 // full scan steps per strand
    const u32 tail = n - nsu * SU;          // leftover candidate positions per strand (< SU)

    key_t best = ~(key_t)0;
    u32 bestsu = 0xffffffffu;
    u32 tie = 0;
    // full steps: step index su in [0, 2*nsu); su >= nsu addresses the reverse complement
    const u32 su_end = fwd_only ? nsu : 2 * nsu;       // fwd_only: lmsr / lmsr_index of the record itself
    const u32 tail_end = fwd_only ? tail : 2 * tail;
    for (u32 su = rank; su < su_end; su += gs) {
        const bool rc = su >= nsu;
        const u32 *X = rc ? Xr : Xf;
        const u32 p0 = (rc ? su - nsu : su) * SU;
        const u32 j = p0 / S, s0 = (p0 % S) * BITS;
        const u32 x0 = X[j], x1 = X[j + 1];
        u32 x2 = 0;
        if (Lane<BITS>::KEY64) x2 = X[j + 2];
        key_t m = ~(key_t)0;
#pragma unroll
        for (int i = 0; i < SU; i++) {
            key_t k;
            if (!Lane<BITS>::KEY64) k = (key_t)funnel_l(x0, x1, s0 + i * BITS);
            else k = (key_t)(((u64)funnel_l(x0, x1, s0 + i * BITS) << 32) | funnel_l(x1, x2, s0 + i * BITS));
            m = k < m ? k : m;
        }
        tie = (m == best) ? 1u : (m < best ? 0u : tie);
        if (m < best) { best = m; bestsu = su; }
    }
    // tail positions: one candidate per thread, encoded as su = 2*nsu + idx
    for (u32 idx = rank; idx < tail_end; idx += gs) {
        const bool rc = idx >= tail;
        const u32 p = nsu * SU + (rc ? idx - tail : idx);
        key_t k = key_at<BITS>(rc ? Xr : Xf, p);
        tie = (k == best) ? 1u : (k < best ? 0u : tie);
        if (k < best) { best = k; bestsu = 2 * nsu + idx; }
    }
    const key_t gmin = grp_min_key<G>(best, red);
    const bool mine = (best == gmin) && bestsu != 0xffffffffu;
    const u32 holders = G::sum_u32(mine ? 1u : 0u, red);
    const u32 anytie = G::or_u32(mine ? tie : 0u, red);

    RecordOut out;
    if (holders == 1 && !anytie) {
        // unique minimal key: its holder decodes the position (first match inside its step; a later
        // match in the same step is < K symbols away and loses the duel by construction)
        u32 enc = 0;
        if (mine) {
            u32 strand, p;
            if (bestsu >= 2 * nsu) {
                u32 idx = bestsu - 2 * nsu;
                strand = idx >= tail;
                p = nsu * SU + (strand ? idx - tail : idx);
            } else {
                strand = bestsu >= nsu;
                u32 p0 = (strand ? bestsu - nsu : bestsu) * SU;
                const u32 *X = strand ? Xr : Xf;
                p = p0;
                for (int i = 0; i < SU; i++)
                    if (key_at<BITS>(X, p0 + i) == gmin) { p = p0 + i; break; }
            }
            enc = (p << 1) | strand;
        }
        enc = G::or_u32(enc, red);          // only one thread contributes
        out.strand = enc & 1u;
        out.start = enc >> 1;
        return out;
    }
    // tie path
    u32 f = strand_tie_winner<BITS, G>(Xf, n, gmin, scr, red);
    u32 r = fwd_only ? 0xffffffffu : strand_tie_winner<BITS, G>(Xr, n, gmin, scr, red);
    if (f == 0xffffffffu) { out.strand = 1; out.start = r; return out; }
    if (r == 0xffffffffu) { out.strand = 0; out.start = f; return out; }
    bool fwd = forward_strictly_smaller<BITS, G>(Xf, Xr, n, f, r, red);
    out.strand = fwd ? 0u : 1u;
    out.start = fwd ? f : r;
    return out;
}

// ---------------------------------------------------------------------------------------------
// Canonical ASCII: 8 consecutive bytes [t, t+8) of the canonical form as a little-endian u64.
// Valid for any t < n (bytes past n wrap around; callers mask what they do not need).
template <int BITS> __device__ __forceinline__ u64 ascii8(const u32 *X, u32 n, u32 start, u32 t)
{
    u32 q = start + t; if (q >= n) q -= n;
    if (BITS == 2) {
        constexpr int S = 16;
        u32 j = q / S, s = (q % S) * 2;
        u32 w = funnel_l(X[j], X[j + 1], s) >> 16;          // 8 symbols, first one most significant
        return ((u64)ascii4_from_2bit(w & 0xffu) << 32) | ascii4_from_2bit(w >> 8);
    } else if (BITS == 4) {
        u32 w = window32<4>(X, q);
        u64 o = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) o |= (u64)s_tab.sym4[(w >> (28 - 4 * k)) & 15u] << (8 * k);
        return o;
    } else {
        u32 a = window32<8>(X, q), b = window32<8>(X, q + 4);     // q + 4 stays inside the extension
        return ((u64)__byte_perm(b, 0, 0x0123) << 32) | __byte_perm(a, 0, 0x0123);
    }
}

}  // namespace ck
