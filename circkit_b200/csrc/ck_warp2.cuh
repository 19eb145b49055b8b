// ck_warp2.cuh -- the hot kernel: warp-per-record LMSR + canonical form + XXH3-64 for 2-bit records
// (viroid / circRNA lengths, configs 1, 2 and 5 of BASELINE.json).
//
// Same answer as ck_record.cuh (the generic path it falls back to for true ties and tiny records).  The
// kernel is issue-bound (integer ALU pipe: SHF / PRMT / VIMNMX3 issue every other cycle per SM sub-partition,
// tools/micro/pipes.cu), so every phase is written to a warp-instruction budget:
//   * staging: one 64-bit load per lane into the shared-memory strand, five circular-extension units,
//     reverse complement by BREV + pair swap;
//   * scan: 8-mer keys as 16-bit halves; one lane-step covers 32 rotations with 14 funnel shifts and
//     8 VIMNMX3.U16x2 (each 32-bit window holds the keys of rotations i and i+8); one REDUX per round keeps the
//     warp minimum, a per-lane round mask remembers where it occurred;
//   * resolve: the (usually single) step that holds the minimal 8-mer is re-read cooperatively, one rotation
//     per lane, comparing full 16-mers; a unique minimal 16-mer ends the search, anything else takes the duel path;
//   * ASCII: 16 bases -> 16 letters with 13 ALU ops (two masks split the window into PRMT selector nibbles, four
//     PRMT look-ups in the register constant "ACGT", four PRMT interleaves), no shared-memory table;
//   * output: with CK_F_ALIGNED_OUT record i's canonical bytes start at 16 * ((offsets[i] >> 4) + i), so every
//     store is a full 128-bit line piece and the XXH3 stripes read the very same registers (one ASCII
//     generation feeds both);
//   * XXH3-64 (n > 240): lane = (stripe, accumulator pair); the last stripe rides in spare lanes of the
//     final round; accumulators are reduced across lanes once per 1024-byte block with a transposed butterfly.
#pragma once
#include "ck_kernels.cuh"

namespace ck {

// 16 bases starting at position q (< n) of strand X, as 32 bits
__device__ __forceinline__ u32 w2_window(const u32 *X, u32 q)
{
    const u32 j = q >> 4;
    return __funnelshift_l(X[j + 1], X[j], 2u * q);           // shift taken mod 32
}

// ---- unaligned output (same offsets as the input): destination-aligned chunks with ragged record edges
struct W2Lut { const u32 *Xf; };
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
__device__ __forceinline__ void w2_emit_ragged(const u32 *X, u32 n, u32 start, u8 *dst)
{
    const u32 lane = lane_id();
    const u32 a = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const u32 nchunks = (n + a + 15u) >> 4;
    for (u32 c = lane; c < nchunks; c += 32) {
        const int t0 = (int)(16u * c) - (int)a;               // record-relative byte of this chunk's first byte
        u32 q = start + (u32)(t0 < 0 ? t0 + (int)n : t0);     // n >= 128 here, so one wrap suffices
        if (q >= n) q -= n;
        const uint4 v = w2_ascii16(w2_window(X, q));
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
};                      // lastsec must follow sec directly: the hash indexes both through one base pointer
__device__ __forceinline__ void w2_fill_const(W2Const *K)
{
    const u32 t = threadIdx.x;
    if (t < 24) K->sec[t] = sec64(8 * (int)t);
    else if (t < 32) K->lastsec[t - 24] = sec64(121 + 8 * (int)(t - 24));
}

struct W2Hash {
    u64 init;           // initial accumulator of this lane's accumulator index
    u64 scr;            // scramble key  (secret + 128 + 8 idx)
    u64 mrg;            // merge key     (secret + 11 + 8 idx)
};
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

// SMALL: every record of the launch has n <= 512 (one scan round, one staging pass); otherwise n <= 8192.
template <bool SMALL>
__global__ void __launch_bounds__(256) k_canon_w2(CanonArgs a)
{
    extern __shared__ __align__(16) u32 smem[];
    typedef Grp<false> G;
    W2Const *K = reinterpret_cast<W2Const *>(smem);
    w2_fill_const(K);
    __syncthreads();
    const u32 lane = lane_id(), wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const u32 stride = a.smem_units;                              // even, so both strands stay 8-byte aligned
    u32 *Xf = smem + sizeof(W2Const) / 4 + (size_t)wid * 2 * stride, *Xr = Xf + stride;
    const u32 gw = blockIdx.x * wpb + wid, nw = gridDim.x * wpb;
    u32 *scr = a.scratch + (size_t)gw * a.scratch_stride;
    const u32 count = a.list ? *a.count : a.n_direct;
    const bool fwd_only = (a.mode & 1u) != 0;
    const bool aligned_out = (a.mode & 2u) != 0;
    const u32 aidx = w2_acc_index(lane);
    W2Hash hc;
    hc.init = xxh3_init_acc(aidx);
    hc.scr = K->sec[16 + aidx];
    hc.mrg = sec64(11 + 8 * (int)aidx);
    const u64 *ksec = K->sec;

    for (u32 e = gw; e < count; e += nw) {
        const u32 rec = a.list ? a.list[e] : e;
        const u64 off = a.offsets[rec];
        const u32 n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
        if (a.list == nullptr && (n < a.min_n || n > a.max_n)) continue;  // direct mode: k_classify reported it
        const u64 *src = a.packed2 + ((off >> 5) + rec);
        u8 *dst = a.out ? a.out + (aligned_out ? 16ull * ((off >> 4) + rec) : off) : nullptr;
        RecordOut o;
        u64 h = 0;
        if (n < 128 || fwd_only) {
            // tiny records and the forward-only library calls: generic path
            RecordIn in; in.packed2 = src; in.bytes = nullptr; in.n = n;
            stage_record<2, G>(in, Xf, Xr);
            o = canonical_start<2, G>(Xf, Xr, n, scr, nullptr, fwd_only);
            const u32 *X = o.strand ? Xr : Xf;
            if (dst) emit_ascii<2, G>(X, n, o.start, dst);
            if (a.out_hash) h = xxh3_canonical<2, G>(X, n, o.start, nullptr, nullptr);
        } else {
            const u32 jn = n >> 4, rem = n & 15u;
            // ---- stage the forward strand
            {
                const u32 W = (n + 31) >> 5;
                if (SMALL) {
                    if (lane < W) *reinterpret_cast<uint2 *>(Xf + 2 * lane) = __ldg(reinterpret_cast<const uint2 *>(src) + lane);
                } else {
                    for (u32 k = lane; k < W; k += 32)
                        *reinterpret_cast<uint2 *>(Xf + 2 * k) = __ldg(reinterpret_cast<const uint2 *>(src) + k);
                }
            }
            __syncwarp();
            if (lane < 5) {                                       // circular extension: units jn .. jn + 4
                u32 val;
                if (lane == 0) {
                    const u32 g0 = Xf[0];
                    val = rem ? ((Xf[jn] & ~(0xffffffffu >> (2 * rem))) | (g0 >> (2 * rem))) : g0;
                } else {
                    val = w2_window(Xf, 16u * lane - rem);        // (jn + lane) * 16 - n; reads units < jn (n >= 128)
                }
                Xf[jn + lane] = val;
            }
            __syncwarp();
            // ---- reverse complement strand, units 0 .. jn + 4
            for (u32 j = lane; j < jn + 5; j += 32) {
                int t = (int)n - 16 * (int)(j + 1);
                if (t < 0) t += (int)n;
                Xr[j] = revcomp2_u32(w2_window(Xf, (u32)t));
            }
            __syncwarp();
            // ---- scan: minimal 8-mer over the 2n rotations; step t < S: forward units (2t, 2t+1), else reverse
            const u32 S = (n + 31) >> 5, T2 = 2 * S;
            u32 gbest = 0xffffffffu, rmask = 0;
            {
                u32 r = 0;
                for (u32 t = lane; (SMALL ? r < 1 : 32 * r < T2); t += 32, r++) {
                    u32 mm = 0x10000u;
                    if (t < T2) {
                        const u32 u = t >= S ? t - S : t;
                        const u32 *p = Xf + (t >= S ? stride : 0u) + 2 * u;
                        const uint2 x01 = *reinterpret_cast<const uint2 *>(p);
                        const u32 m = w2_step_min16(x01.x, x01.y, p[2]);
                        mm = min(m >> 16, m & 0xffffu);
                    }
                    const u32 g = __reduce_min_sync(CK_FULL, mm);
                    if (g < gbest) { gbest = g; rmask = 0; }
                    if (mm == gbest) rmask |= 1u << r;
                }
            }
            // ---- resolve: every step that holds the minimal 8-mer is re-read, one rotation per lane, as 16-mers
            u32 best32 = 0xffffffffu, cnt = 0, bestpos = 0, beststrand = 0;
            for (u32 any = __ballot_sync(CK_FULL, rmask != 0); any; any = __ballot_sync(CK_FULL, rmask != 0)) {
                const u32 L = __ffs(any) - 1;
                const u32 mL = __shfl_sync(CK_FULL, rmask, L);
                const u32 r = __ffs(mL) - 1;
                if (lane == L) rmask &= rmask - 1;
                const u32 t = L + 32 * r;
                const u32 strand = t >= S ? 1u : 0u;
                const u32 p = 32 * (strand ? t - S : t) + lane;
                const u32 w = w2_window(strand ? Xr : Xf, p);
                const bool hit = (w >> 16) == gbest && p < n;
                const u32 wv = hit ? w : 0xffffffffu;
                const u32 g32 = __reduce_min_sync(CK_FULL, wv);
                const u32 hb = __ballot_sync(CK_FULL, hit && wv == g32);
                if (g32 < best32) { best32 = g32; cnt = 0; }
                if (g32 == best32 && hb) { cnt += __popc(hb); bestpos = p - lane + (__ffs(hb) - 1); beststrand = strand; }
            }
            if (cnt == 1) {
                o.start = bestpos; o.strand = beststrand;
            } else {
                // equal minimal 16-mers (repeats, multimers, palindromic circles): duel-based generic path
                o = canonical_start<2, G>(Xf, Xr, n, scr, nullptr, false);
            }
            const u32 *X = o.strand ? Xr : Xf;
            // ---- canonical ASCII (+ XXH3-64 of it)
            const bool long_hash = a.out_hash != nullptr && n > 240;
            if (dst && !aligned_out) w2_emit_ragged(X, n, o.start, dst);
            if ((dst && aligned_out) || long_hash) {
                const u32 nchunks = (n + 15) >> 4;
                const u32 vbase = (nchunks + 3u) & ~3u;               // last-stripe chunks: 4 spare lanes after the data
                const u32 cend = long_hash ? vbase + 4 : nchunks;
                const u32 nfull = (n - 1) >> 6;                       // stripes consumed by the stripe loop
                const u32 nb_blocks = (n - 1) >> 10;
                const bool store = dst && aligned_out;
                u64 base = hc.init, acc0 = 0, acc1 = 0;
                u32 r = 0;
                for (u32 c = lane; 32 * r < cend; c += 32, r++) {
                    if (c < cend) {
                        const bool last = c >= vbase;
                        u32 q = o.start + (last ? n - 64 + 16 * (c - vbase) : 16 * c);
                        if (q >= n) q -= n;
                        const uint4 v = w2_ascii16(w2_window(X, q));
                        if (store && c < nchunks) *reinterpret_cast<uint4 *>(dst + 16 * (size_t)c) = v;
                        if (long_hash) {
                            const u32 s = c >> 2;
                            const u32 sidx = last ? 24 + 2 * (c - vbase) : (s & 15u) + 2 * (c & 3u);
                            if (last || s < nfull) {
                                const u64 d0 = ((u64)v.y << 32) | v.x, d1 = ((u64)v.w << 32) | v.z;
                                const u64 k0 = d0 ^ ksec[sidx], k1 = d1 ^ ksec[sidx + 1];
                                acc0 += (u64)(u32)k0 * (u64)(u32)(k0 >> 32) + d1;
                                acc1 += (u64)(u32)k1 * (u64)(u32)(k1 >> 32) + d0;
                            }
                        }
                    }
                    if (long_hash && (r & 1u) && (r >> 1) < nb_blocks) {      // a 1024-byte block is complete
                        u64 t = base + w2_reduce_pair(acc0, acc1, lane);
                        t ^= t >> 47; t ^= hc.scr; t *= CK_P32_1;
                        base = t; acc0 = 0; acc1 = 0;
                    }
                }
                if (long_hash) {
                    const u64 keyed = (base + w2_reduce_pair(acc0, acc1, lane)) ^ hc.mrg;    // acc[aidx] ^ secret[11 + 8 aidx]
                    const u64 partner = __shfl_xor_sync(CK_FULL, keyed, 4);
                    u64 term = lane < 4 ? mul128_fold64(keyed, partner) : 0ULL;
                    term += __shfl_xor_sync(CK_FULL, term, 1);
                    term += __shfl_xor_sync(CK_FULL, term, 2);
                    h = xxh3_avalanche((u64)n * CK_P64_1 + term);                            // valid in lanes 0..3
                }
            }
            if (a.out_hash && n <= 240) h = xxh3_short_warp<2>(X, n, o.start);
        }
        if (lane == 0) {
            if (a.out_start) a.out_start[rec] = o.strand ? (n - 1 - o.start) : o.start;
            if (a.out_strand) a.out_strand[rec] = (u8)o.strand;
            if (a.out_hash) a.out_hash[rec] = h;
        }
        __syncwarp();
    }
}

}  // namespace ck
