// ck_warp2.cuh -- the hot kernel: warp-per-record LMSR + canonical form + XXH3-64 for 2-bit records
// (viroid / circRNA lengths, configs 1, 2 and 5 of BASELINE.json).
//
// Same answer as ck_record.cuh (the generic path it falls back to for true ties and tiny records).  The
// kernel is issue-bound (integer ALU pipe: SHF / PRMT / VIMNMX3 issue every other cycle per SM sub-partition,
// tools/micro/pipes.cu), so every phase is written to a warp-instruction budget:
//   * staging: one 64-bit load per lane into the shared-memory strand, three circular-extension units,
//     reverse complement by BREV + one LOP3;
//   * scan: 8-mer keys as 16-bit halves; one lane-step covers 32 rotations with 14 funnel shifts and
//     8 VIMNMX3.U16x2 (each 32-bit window holds the keys of rotations i and i+8); one REDUX per round keeps the
//     warp minimum, a per-lane round mask remembers where it occurred;
//   * resolve: the (usually single) step that holds the minimal 8-mer is re-read cooperatively, one rotation
//     per lane, comparing full 16-mers; a unique minimal 16-mer ends the search, anything else takes the duel path;
//   * ASCII: 16 bases -> 16 letters with 13 ALU ops (two masks split the window into PRMT selector nibbles, four
//     PRMT look-ups in the register constant "ACGT", four PRMT interleaves), no shared-memory table;
//   * output: with CK_F_ALIGNED_OUT record i's canonical bytes start at 32 * ((offsets[i] >> 5) + i), so every
//     store is a full 128-bit line piece and the XXH3 stripes read the very same registers (one ASCII
//     generation feeds both);
//   * XXH3-64 (n > 240): lane = (stripe, accumulator pair); the last stripe rides in spare lanes of the
//     final round; accumulators are reduced across lanes once per 1024-byte block with a transposed butterfly.
// Shared memory is addressed with 32-bit shared-window addresses (ld.shared / st.shared) to keep address
// arithmetic to one instruction per access.
#pragma once
#include "ck_kernels.cuh"

namespace ck {

// ---- shared memory by 32-bit address
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 lds64(u32 a)
{
    uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(u32 a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }

// 16 bases starting at position q of the strand whose unit 0 sits at shared address xb
__device__ __forceinline__ u32 w2_win(u32 xb, u32 q)
{
    const u32 a = xb + ((q >> 4) << 2);
    const u32 hi = lds32(a), lo = lds32(a + 4);
    return __funnelshift_l(lo, hi, q + q);                        // shift taken mod 32
}
// reverse complement of 16 packed bases: reverse the bit order, swap the two bits of every pair back, complement
__device__ __forceinline__ u32 w2_revcomp(u32 x)
{
    const u32 r = __brev(x);
    u32 o;
    asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1b;" : "=r"(o) : "r"(r >> 1), "r"(r + r));    // ~((a & c) | (b & ~c))
    return o;
}

__device__ __forceinline__ uint4 w2_ascii16(u32 w)
{
    const u32 T = 0x54474341u;                                // 'A','C','G','T'
    const u32 E = w & 0x33333333u;                            // nibble j: base 15 - 2j
    const u32 O = (w >> 2) & 0x33333333u;                     // nibble j: base 14 - 2j
    const u32 pe_lo = __byte_perm(T, 0, E), pe_hi = __byte_perm(T, 0, E >> 16);
    const u32 po_lo = __byte_perm(T, 0, O), po_hi = __byte_perm(T, 0, O >> 16);
    uint4 v;
    v.x = __byte_perm(pe_hi, po_hi, 0x2637);                  // bases 0..3 (byte 0 = base 0)
    v.y = __byte_perm(pe_hi, po_hi, 0x0415);                  // bases 4..7
    v.z = __byte_perm(pe_lo, po_lo, 0x2637);                  // bases 8..11
    v.w = __byte_perm(pe_lo, po_lo, 0x0415);                  // bases 12..15
    return v;
}
// unaligned output (same offsets as the input): destination-aligned chunks with ragged record edges
__device__ __noinline__ void w2_emit_ragged(u32 xb, u32 n, u32 start, u8 *dst)
{
    const u32 lane = lane_id();
    const u32 a = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const u32 nchunks = (n + a + 15u) >> 4;
    for (u32 c = lane; c < nchunks; c += 32) {
        const int t0 = (int)(16u * c) - (int)a;               // record-relative byte of this chunk's first byte
        u32 q = start + (u32)(t0 < 0 ? t0 + (int)n : t0);     // n >= 128 here, so one wrap suffices
        if (q >= n) q -= n;
        const uint4 v = w2_ascii16(w2_win(xb, q));
        store_chunk(dst, t0, n, ((u64)v.y << 32) | v.x, ((u64)v.w << 32) | v.z);
    }
}

// ---- scan step: minimum 8-mer key over the 32 rotations that start in units (x0, x1); u16x2 result
__device__ __forceinline__ u32 w2_step_min16(u32 x0, u32 x1, u32 x2)
{
    // window i (i = 0..7) of (x0:x1) holds key(i) in its high half and key(i + 8) in its low half
    u32 a1 = __funnelshift_l(x1, x0, 2), a2 = __funnelshift_l(x1, x0, 4), a3 = __funnelshift_l(x1, x0, 6);
    u32 a4 = __funnelshift_l(x1, x0, 8), a5 = __funnelshift_l(x1, x0, 10), a6 = __funnelshift_l(x1, x0, 12);
    u32 a7 = __funnelshift_l(x1, x0, 14);
    u32 b1 = __funnelshift_l(x2, x1, 2), b2 = __funnelshift_l(x2, x1, 4), b3 = __funnelshift_l(x2, x1, 6);
    u32 b4 = __funnelshift_l(x2, x1, 8), b5 = __funnelshift_l(x2, x1, 10), b6 = __funnelshift_l(x2, x1, 12);
    u32 b7 = __funnelshift_l(x2, x1, 14);
    u32 m0 = __vimin3_u16x2(x0, a1, a2), m1 = __vimin3_u16x2(a3, a4, a5), m2 = __vimin3_u16x2(a6, a7, x1);
    u32 m3 = __vimin3_u16x2(b1, b2, b3), m4 = __vimin3_u16x2(b4, b5, b6);
    u32 m5 = __vimin3_u16x2(m0, m1, m2), m6 = __vimin3_u16x2(m3, m4, b7);
    return __vminu2(m5, m6);
}

// ---- per-CTA constants in shared memory (XXH3 secret pieces at the alignments the long hash reads them)
struct W2Const {
    u64 sec[24];        // secret as LE u64 at byte offsets 8 i        (stripe keys, scramble keys = sec[16 + i])
    u64 lastsec[8];     // LE u64 at byte offsets 121 + 8 i            (last stripe: secret + 192 - 64 - 7)
};                      // lastsec must follow sec directly: the hash indexes both through one base address
__device__ __forceinline__ void w2_fill_const(W2Const *K)
{
    const u32 t = threadIdx.x;
    if (t < 24) K->sec[t] = sec64(8 * (int)t);
    else if (t < 32) K->lastsec[t - 24] = sec64(121 + 8 * (int)(t - 24));
}
// lane -> accumulator index it owns after a block reduction: pair k = lane & 3, bit 2 of the lane picks 2k or 2k+1
__device__ __forceinline__ u32 w2_acc_index(u32 lane) { return 2u * (lane & 3u) + ((lane >> 2) & 1u); }

// sum (a0, a1) over the 8 lanes that share lane & 3; afterwards lanes with bit 2 clear hold the total of a0,
// lanes with bit 2 set the total of a1 (transposed first butterfly step: 6 shuffles instead of 12)
__device__ __forceinline__ u64 w2_reduce_pair(u64 a0, u64 a1, u32 lane)
{
    const bool up = (lane & 4u) != 0;
    const u64 send = up ? a0 : a1;
    u64 keep = up ? a1 : a0;
    keep += __shfl_xor_sync(CK_FULL, send, 4);
    keep += __shfl_xor_sync(CK_FULL, keep, 8);
    keep += __shfl_xor_sync(CK_FULL, keep, 16);
    return keep;
}

// slow paths, kept out of line so the common path stays small; results come back packed as (start << 1) | strand
__device__ __noinline__ u32 w2_slow_start(u32 *Xf, u32 *Xr, u32 n, u32 xb, u32 *scr)
{
    // the duel path reads two more extension units than the scan: build units jn+3, jn+4 of both strands
    const u32 lane = lane_id(), jn = n >> 4, rem = n & 15u;
    const u32 stride = (u32)(Xr - Xf);
    if (lane < 2) sts32(xb + 4 * (jn + 3 + lane), w2_win(xb, 16u * (3 + lane) - rem));
    __syncwarp();
    if (lane < 2) {
        int t = (int)n - 16 * (int)(jn + 4 + lane);
        if (t < 0) t += (int)n;
        sts32(xb + 4 * (stride + jn + 3 + lane), w2_revcomp(w2_win(xb, (u32)t)));
    }
    __syncwarp();
    const RecordOut o = canonical_start<2, Grp<false> >(Xf, Xr, n, scr, nullptr, false);
    return (o.start << 1) | o.strand;
}
__device__ __noinline__ u32 w2_tiny_record(const u64 *src, u32 n, u8 *dst, u32 *Xf, u32 *Xr, u32 *scr, bool fwd_only)
{
    typedef Grp<false> G;
    RecordIn in; in.packed2 = src; in.bytes = nullptr; in.n = n;
    stage_record<2, G>(in, Xf, Xr);
    const RecordOut o = canonical_start<2, G>(Xf, Xr, n, scr, nullptr, fwd_only);
    const u32 *X = o.strand ? Xr : Xf;
    if (dst) emit_ascii<2, G>(X, n, o.start, dst);
    return (o.start << 1) | o.strand;
}
__device__ __noinline__ u64 w2_short_hash(const u32 *X, u32 n, u32 start) { return xxh3_short_warp<2>(X, n, start); }
__device__ __noinline__ u64 w2_any_hash(const u32 *X, u32 n, u32 start) { return xxh3_canonical<2, Grp<false> >(X, n, start, nullptr, nullptr); }

// Kernel variants.  V < 0: every option is read from CanonArgs at run time (normalised lengths, forward-only
// library calls, either output layout, optional outputs).  V >= 0: packed lengths from the offsets, both strands,
// out_start / out_strand present; bit 0 = XXH3-64 wanted, bit 1 = canonical bytes wanted in the aligned arena (else
// no bytes), bit 2 = records come from a work list (else the direct range [0, n_direct)).
// SMALL: every record of the launch has n <= 512 (one scan round, one staging pass); otherwise n <= 8192.
#define CK_W2_HASH 1
#define CK_W2_OUT 2
#define CK_W2_LIST 4
template <bool SMALL, int V>
__global__ void __launch_bounds__(256, 4) k_canon_w2(CanonArgs a)
{
    extern __shared__ __align__(16) u32 smem[];
    W2Const *K = reinterpret_cast<W2Const *>(smem);
    w2_fill_const(K);
    __syncthreads();
    constexpr bool RT = V < 0;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const u32 stride = a.smem_units;                              // even, so both strands stay 8-byte aligned
    u32 *Xf = smem + sizeof(W2Const) / 4 + (size_t)wid * 2 * stride, *Xr = Xf + stride;
    const u32 kb = (u32)__cvta_generic_to_shared(smem);           // W2Const
    const u32 xf = (u32)__cvta_generic_to_shared(Xf), xr = xf + 4 * stride;
    const u32 gw = blockIdx.x * wpb + wid, nw = gridDim.x * wpb;
    u32 *scr = a.scratch + (size_t)gw * a.scratch_stride;
    const bool use_list = RT ? a.list != nullptr : (V & CK_W2_LIST) != 0;
    const u32 count = use_list ? *a.count : a.n_direct;
    if (use_list) a.list += a.count[16];
    const bool fwd_only = RT && (a.mode & 1u) != 0;
    const bool aligned_out = RT ? (a.mode & 2u) != 0 : true;
    const bool want_hash = RT ? a.out_hash != nullptr : (V & CK_W2_HASH) != 0;
    const bool want_out = RT ? a.out != nullptr : (V & CK_W2_OUT) != 0;
    const u32 aidx = w2_acc_index(lane);
    u64 h_init = 0, h_scr = 0, h_mrg = 0;
    if (want_hash) { h_init = xxh3_init_acc(aidx); h_scr = K->sec[16 + aidx]; h_mrg = sec64(11 + 8 * (int)aidx); }

    // software pipeline: the offsets (and, in direct mode, the first packed word of every lane) of the next record are
    // requested while the current one is processed, so neither DRAM round trip sits on the critical path
    u32 e = gw;
    u32 rec_n = 0, rec_nn = 0; u64 off_n = 0; u32 end_n = 0; uint2 d_n = make_uint2(0, 0);
    const bool pipelined = !RT;
    // the next record's packed words are requested once its offsets have had time to arrive (after the staging phases)
#define CK_W2_PREFETCH_DATA()                                                                                        \
    do {                                                                                                             \
        if (pipelined && e + nw < count) {                                                                           \
            const u32 nn = end_n - (u32)off_n;                                                                       \
            if (lane < ((nn + 31) >> 5))                                                                             \
                d_n = __ldg(reinterpret_cast<const uint2 *>(a.packed2 + p2_word(off_n, rec_n, a.p2_dbl)) + lane);             \
        }                                                                                                            \
    } while (0)
    if (pipelined && e < count) {
        rec_n = use_list ? a.list[e] : e;
        if (use_list && e + nw < count) rec_nn = a.list[e + nw];
        off_n = a.offsets[rec_n]; end_n = (u32)a.offsets[rec_n + 1];
        const u32 nn = end_n - (u32)off_n;
        if (lane < ((nn + 31) >> 5)) d_n = __ldg(reinterpret_cast<const uint2 *>(a.packed2 + p2_word(off_n, rec_n, a.p2_dbl)) + lane);
    }
    for (; e < count; e += nw) {
        u32 rec; u64 off; u32 n; uint2 d0 = make_uint2(0, 0);
        if (pipelined) {
            rec = rec_n; off = off_n; n = end_n - (u32)off_n; d0 = d_n;
            const u32 e2 = e + nw;
            if (e2 < count) {
                rec_n = use_list ? rec_nn : e2;                       // list entries are fetched two records ahead
                if (use_list && e2 + nw < count) rec_nn = a.list[e2 + nw];
                off_n = a.offsets[rec_n]; end_n = (u32)a.offsets[rec_n + 1];
            }
        } else {
            rec = use_list ? a.list[e] : e;
            off = a.offsets[rec];
            n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
        }
        const bool in_range = use_list || (n >= a.min_n && n <= a.max_n);    // direct mode: k_classify reported the rest
        const u64 *src = a.packed2 + p2_word(off, rec, a.p2_dbl);
        u8 *dst = want_out ? a.out + (aligned_out ? out_byte(off, rec) : off) : nullptr;
        u32 os = 0;                                               // (start << 1) | strand of the canonical rotation
        u64 h = 0;
        if (!in_range) {
            CK_W2_PREFETCH_DATA();
        } else if (n < 128 || fwd_only) {
            // tiny records and the forward-only library calls: generic path
            os = w2_tiny_record(src, n, dst, Xf, Xr, scr, fwd_only);
            if (want_hash) h = w2_any_hash((os & 1u) ? Xr : Xf, n, os >> 1);
            CK_W2_PREFETCH_DATA();
        } else {
            const u32 jn = n >> 4, rem = n & 15u;
            // ---- stage the forward strand
            {
                const u32 W = (n + 31) >> 5;
                if (pipelined) {
                    if (lane < W) sts64(xf + 8 * lane, d0);
                    if (!SMALL) {
#pragma unroll 1
                        for (u32 k = lane + 32; k < W; k += 32) sts64(xf + 8 * k, __ldg(reinterpret_cast<const uint2 *>(src) + k));
                    }
                } else {
#pragma unroll 1
                    for (u32 k = lane; k < W; k += 32) sts64(xf + 8 * k, __ldg(reinterpret_cast<const uint2 *>(src) + k));
                }
            }
            __syncwarp();
            if (lane < 3) {                                       // circular extension: units jn .. jn + 2
                u32 val;
                if (lane == 0) val = (lds32(xf + 4 * jn) & ~(0xffffffffu >> (2 * rem))) | (lds32(xf) >> (2 * rem));
                else val = w2_win(xf, 16u * lane - rem);          // (jn + lane) * 16 - n; reads units < jn (n >= 128)
                sts32(xf + 4 * (jn + lane), val);
            }
            __syncwarp();
            // ---- reverse complement strand, units 0 .. jn + 2
#pragma unroll 1
            for (u32 j = lane; j < jn + 3; j += 32) {
                int t = (int)n - 16 * (int)(j + 1);
                if (t < 0) t += (int)n;
                sts32(xr + 4 * j, w2_revcomp(w2_win(xf, (u32)t)));
            }
            __syncwarp();
            CK_W2_PREFETCH_DATA();
            // ---- scan: minimal 8-mer over the 2n rotations; step t < S: forward units (2t, 2t+1), else reverse
            const u32 S = (n + 31) >> 5, T2 = 2 * S;
            const u32 xrs = xr - 8 * S;                           // step t >= S reads reverse units 2 (t - S)
            u32 cnt = 0, best = 0;
            if (SMALL) {
                const u32 tt = min(lane, T2 - 1);
                const u32 p = (tt >= S ? xrs : xf) + 8 * tt;
                const uint2 x01 = lds64(p);
                const u32 m = w2_step_min16(x01.x, x01.y, lds32(p + 8));
                const u32 mm = lane < T2 ? min(m >> 16, m & 0xffffu) : 0x10000u;
                const u32 gbest = __reduce_min_sync(CK_FULL, mm);
                u32 cand = __ballot_sync(CK_FULL, mm == gbest);
                // resolve: re-read the step(s) holding the minimal 8-mer, one rotation per lane, as 16-mers
                u32 best32 = 0xffffffffu;
                do {
                    const u32 L = 31u - __clz(cand);
                    cand ^= 1u << L;
                    const u32 strand = L >= S ? 1u : 0u;
                    const u32 p0 = 32 * (strand ? L - S : L);
                    const u32 w = w2_win(strand ? xr : xf, p0 + lane);
                    const bool hit = (w >> 16) == gbest && p0 + lane < n;
                    const u32 hits = __ballot_sync(CK_FULL, hit);
                    if (cand == 0 && cnt == 0 && hits != 0 && (hits & (hits - 1)) == 0) {     // the common case: one step, one rotation
                        cnt = 1; best = ((p0 + 31u - __clz(hits)) << 1) | strand;
                        break;
                    }
                    const u32 wv = hit ? w : 0xffffffffu;
                    const u32 g32 = __reduce_min_sync(CK_FULL, wv);
                    const u32 hb = __ballot_sync(CK_FULL, hit && wv == g32);
                    if (g32 < best32) { best32 = g32; cnt = 0; }
                    if (g32 == best32 && hb) { cnt += __popc(hb); best = ((p0 + 31u - __clz(hb)) << 1) | strand; }
                } while (cand);
            } else {
                u32 gbest = 0xffffffffu, rmask = 0, r = 0;
#pragma unroll 1
                for (u32 t = lane; 32 * r < T2; t += 32, r++) {
                    const u32 tt = min(t, T2 - 1);
                    const u32 p = (tt >= S ? xrs : xf) + 8 * tt;
                    const uint2 x01 = lds64(p);
                    const u32 m = w2_step_min16(x01.x, x01.y, lds32(p + 8));
                    const u32 mm = t < T2 ? min(m >> 16, m & 0xffffu) : 0x10000u;
                    const u32 g = __reduce_min_sync(CK_FULL, mm);
                    if (g < gbest) { gbest = g; rmask = 0; }
                    if (mm == gbest) rmask |= 1u << r;
                }
                u32 best32 = 0xffffffffu;
#pragma unroll 1
                for (u32 any = __ballot_sync(CK_FULL, rmask != 0); any; any = __ballot_sync(CK_FULL, rmask != 0)) {
                    const u32 L = 31u - __clz(any);
                    const u32 mL = __shfl_sync(CK_FULL, rmask, L);
                    const u32 rr = 31u - __clz(mL);
                    if (lane == L) rmask ^= 1u << rr;
                    const u32 t = L + 32 * rr;
                    const u32 strand = t >= S ? 1u : 0u;
                    const u32 p0 = 32 * (strand ? t - S : t);
                    const u32 w = w2_win(strand ? xr : xf, p0 + lane);
                    const bool hit = (w >> 16) == gbest && p0 + lane < n;
                    const u32 wv = hit ? w : 0xffffffffu;
                    const u32 g32 = __reduce_min_sync(CK_FULL, wv);
                    const u32 hb = __ballot_sync(CK_FULL, hit && wv == g32);
                    if (g32 < best32) { best32 = g32; cnt = 0; }
                    if (g32 == best32 && hb) { cnt += __popc(hb); best = ((p0 + 31u - __clz(hb)) << 1) | strand; }
                }
            }
            // equal minimal 16-mers (repeats, multimers, palindromic circles): duel-based generic path
            os = cnt == 1 ? best : w2_slow_start(Xf, Xr, n, xf, scr);
            const u32 start = os >> 1;
            const u32 xs = (os & 1u) ? xr : xf;
            // ---- canonical ASCII (+ XXH3-64 of it)
            const bool long_hash = want_hash && n > 240;
            const bool store = dst && aligned_out;
            if (RT && dst && !aligned_out) w2_emit_ragged(xs, n, start, dst);
            if (store || long_hash) {
                const u32 nchunks = (n + 15) >> 4;
                const u32 vbase = (nchunks + 3u) & ~3u;               // last-stripe chunks: 4 spare lanes after the data
                const u32 cend = long_hash ? vbase + 4 : nchunks;
                const u32 nfull = (n - 1) >> 6;                       // stripes consumed by the stripe loop
                const u32 nb_blocks = (n - 1) >> 10;
                u64 base = h_init, acc0 = 0, acc1 = 0;
                u32 r = 0;
#pragma unroll 1
                for (u32 c = lane; 32 * r < cend; c += 32, r++) {
                    const u32 cc = min(c, cend - 1);                  // spare lanes repeat the last chunk (same bytes, same address)
                    const bool last = long_hash && cc >= vbase;
                    u32 q = start + (last ? n - 64 + 16 * (cc - vbase) : 16 * cc);
                    if (q >= n) q -= n;
                    const uint4 v = w2_ascii16(w2_win(xs, q));
                    if (store && cc < nchunks) *reinterpret_cast<uint4 *>(dst + 16 * (size_t)cc) = v;
                    if (long_hash) {
                        const u32 s = cc >> 2;
                        const u32 sidx = last ? 24 + 2 * (cc - vbase) : (s & 15u) + 2 * (cc & 3u);
                        if (c < cend && (last || s < nfull)) {
                            const u64 d0v = ((u64)v.y << 32) | v.x, d1v = ((u64)v.w << 32) | v.z;
                            const uint2 s0 = lds64(kb + 8 * sidx), s1 = lds64(kb + 8 * sidx + 8);
                            const u32 k0l = v.x ^ s0.x, k0h = v.y ^ s0.y, k1l = v.z ^ s1.x, k1h = v.w ^ s1.y;
                            acc0 += (u64)k0l * (u64)k0h + d1v;
                            acc1 += (u64)k1l * (u64)k1h + d0v;
                        }
                    }
                    if (long_hash && !SMALL && (r & 1u) && (r >> 1) < nb_blocks) {      // a 1024-byte block is complete
                        u64 t = base + w2_reduce_pair(acc0, acc1, lane);
                        t ^= t >> 47; t ^= h_scr; t *= CK_P32_1;
                        base = t; acc0 = 0; acc1 = 0;
                    }
                }
                if (long_hash) {
                    const u64 keyed = (base + w2_reduce_pair(acc0, acc1, lane)) ^ h_mrg;    // acc[aidx] ^ secret[11 + 8 aidx]
                    const u64 partner = __shfl_xor_sync(CK_FULL, keyed, 4);
                    u64 term = lane < 4 ? mul128_fold64(keyed, partner) : 0ULL;
                    term += __shfl_xor_sync(CK_FULL, term, 1);
                    term += __shfl_xor_sync(CK_FULL, term, 2);
                    h = xxh3_avalanche((u64)n * CK_P64_1 + term);                            // valid in lanes 0..3
                }
            }
            if (want_hash && n <= 240) h = w2_short_hash((os & 1u) ? Xr : Xf, n, start);
        }
        if (lane == 0 && in_range) {
            const u32 start = os >> 1, strand = os & 1u;
            if (!RT || a.out_start) a.out_start[rec] = strand ? (n - 1 - start) : start;
            if (!RT || a.out_strand) a.out_strand[rec] = (u8)strand;
            if (want_hash) a.out_hash[rec] = h;
        }
        __syncwarp();
    }
}

#undef CK_W2_PREFETCH_DATA

}  // namespace ck
