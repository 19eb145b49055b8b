// ck_warp2.cuh -- the hot kernel: warp-per-record LMSR + canonical form + XXH3-64 for 2-bit records
// (viroid / circRNA lengths, configs 1, 2 and 5 of BASELINE.json).
//
// Same algorithm as ck_record.cuh (the generic path it falls back to for ties and tiny records), with
// the common case -- a unique minimal 16-mer -- written to cost as few warp instructions as possible,
// because this kernel is issue-bound long before it is HBM-bound:
//   * staging: one 64-bit load per lane straight into the shared-memory strand, two extension units,
//     reverse complement by BREV + pair swap, no division anywhere;
//   * scan: each lane takes 8 consecutive rotations per step from two shared-memory words:
//     8 funnel shifts + 4 three-input minima; forward and reverse-complement steps share the loop;
//   * the winner's position is decoded cooperatively (8 lanes, one ballot);
//   * ASCII comes from a 256-entry "4 bases -> 4 letters" table in shared memory; stores are 16-byte,
//     destination-aligned and coalesced, ragged record edges use 8/4/2/1-byte stores;
//   * XXH3 consumes 8 letters per lane and stripe, 4 stripes per warp step.
#pragma once
#include "ck_kernels.cuh"

namespace ck {

struct W2 {
    const u32 *Xf, *Xr;     // strands in shared memory (units 0 .. n/16 + 1 valid)
    const u32 *lut;         // 256 x u32: byte of 4 bases -> 4 ASCII letters (little-endian)
    u32 n;
};

// 16 bases starting at position q (< n) of strand X, as 32 bits
__device__ __forceinline__ u32 w2_window(const u32 *X, u32 q)
{
    u32 j = q >> 4, s = (q & 15u) * 2u;
    return funnel_l(X[j], X[j + 1], s);
}
// canonical bytes [t, t+8) as a little-endian u64 (t < n; bytes past n wrap)
__device__ __forceinline__ u64 w2_ascii8(const u32 *X, const u32 *lut, u32 n, u32 start, u32 t)
{
    u32 q = start + t; if (q >= n) q -= n;
    u32 w = w2_window(X, q);
    return ((u64)lut[(w >> 16) & 0xffu] << 32) | lut[w >> 24];
}
__device__ __forceinline__ uint4 w2_ascii16(const u32 *X, const u32 *lut, u32 q)
{
    u32 w = w2_window(X, q);
    return make_uint4(lut[w >> 24], lut[(w >> 16) & 0xffu], lut[(w >> 8) & 0xffu], lut[w & 0xffu]);
}

// minimum 16-mer key over the 8 rotations of scan step `su` (su < nsu: forward, else reverse complement)
__device__ __forceinline__ u32 w2_step_min(const u32 *Xf, u32 stride, u32 su, u32 nsu)
{
    const bool rc = su >= nsu;
    const u32 u = rc ? su - nsu : su;
    const u32 *X = Xf + (rc ? stride : 0u) + (u >> 1);
    const u32 x0 = X[0], x1 = X[1];
    const u32 s0 = (u & 1u) * 16u;
    const u32 y0 = funnel_l(x0, x1, s0), y1 = x1 << s0;       // align the 8 rotations to shift 0,2,..,14
    u32 m = y0;
#pragma unroll
    for (int i = 1; i < 8; i++) m = min(m, funnel_l(y0, y1, 2 * i));
    return m;
}

// emit canonical ASCII (strand X, rotation `start`) to dst[0 .. n)
__device__ __forceinline__ void w2_emit(const u32 *X, const u32 *lut, u32 n, u32 start, u8 *dst)
{
    const u32 lane = lane_id();
    const u32 a = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const u32 nchunks = (n + a + 15u) >> 4;
    for (u32 c = lane; c < nchunks; c += 32) {
        const int t0 = (int)(16u * c) - (int)a;               // record-relative byte of this chunk's first byte
        u32 q = start + (u32)(t0 < 0 ? t0 + (int)n : t0);     // n >= 64 here, so one wrap suffices
        if (q >= n) q -= n;
        const uint4 v = w2_ascii16(X, lut, q);
        store_chunk(dst, t0, n, ((u64)v.y << 32) | v.x, ((u64)v.w << 32) | v.z);
    }
}

// XXH3-64 of the canonical ASCII, n > 240 (shorter records take the generic path)
__device__ __forceinline__ u64 w2_stripes(const u32 *X, const u32 *lut, u32 n, u32 st, u32 base, u32 nstripes)
{
    const u32 lane = lane_id(), i = lane & 7u, sg = lane >> 3;
    const u64 *sec = reinterpret_cast<const u64 *>(c_secret);
    u64 mul = 0, dv_sum = 0;
    for (u32 s = sg; s < nstripes; s += 4) {
        u64 dv = w2_ascii8(X, lut, n, st, base + 64 * s + 8 * i);
        u64 dk = dv ^ sec[s + i];
        mul += (u64)(u32)dk * (u64)(u32)(dk >> 32);
        dv_sum += dv;
    }
    mul += __shfl_xor_sync(CK_FULL, mul, 8);  dv_sum += __shfl_xor_sync(CK_FULL, dv_sum, 8);
    mul += __shfl_xor_sync(CK_FULL, mul, 16); dv_sum += __shfl_xor_sync(CK_FULL, dv_sum, 16);
    return mul + __shfl_xor_sync(CK_FULL, dv_sum, 1);
}
__device__ __forceinline__ u64 w2_xxh3_long(const u32 *X, const u32 *lut, u32 n, u32 st)
{
    const u32 lane = lane_id(), i = lane & 7u;
    u64 acc = xxh3_init_acc(i);
    const u32 nb_blocks = (n - 1) >> 10;
    for (u32 b = 0; b < nb_blocks; b++) {
        acc += w2_stripes(X, lut, n, st, b << 10, 16);
        acc = xxh3_scramble(acc, i);
    }
    const u32 nstripes = ((n - 1) - (nb_blocks << 10)) >> 6;
    acc += w2_stripes(X, lut, n, st, nb_blocks << 10, nstripes);
    {   // last stripe: input + n - 64, secret + 192 - 64 - 7
        u64 dv = w2_ascii8(X, lut, n, st, n - 64 + 8 * i);
        u64 dk = dv ^ sec64(121 + 8 * (int)i);
        acc += (u64)(u32)dk * (u64)(u32)(dk >> 32) + __shfl_xor_sync(CK_FULL, dv, 1);
    }
    u64 keyed = acc ^ sec64(11 + 8 * (int)i);
    u64 partner = __shfl_xor_sync(CK_FULL, keyed, 1);
    u64 term = (lane < 8 && (lane & 1u) == 0) ? mul128_fold64(keyed, partner) : 0ULL;
    term += __shfl_xor_sync(CK_FULL, term, 2);
    term += __shfl_xor_sync(CK_FULL, term, 4);
    term = __shfl_sync(CK_FULL, term, 0);
    return xxh3_avalanche((u64)n * CK_P64_1 + term);
}

// ROUNDS > 0: scan fully unrolled over ROUNDS warp steps (n <= ROUNDS * 128); ROUNDS == 0: looped scan.
template <int ROUNDS>
__global__ void __launch_bounds__(256) k_canon_w2(CanonArgs a)
{
    extern __shared__ u32 smem[];
    __shared__ u32 lut[256];
    typedef Grp<false> G;
    lut[threadIdx.x & 255u] = ascii4_from_2bit(threadIdx.x & 255u);
    __syncthreads();
    const u32 lane = lane_id(), wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const u32 stride = a.smem_units;
    u32 *Xf = smem + (size_t)wid * 2 * stride, *Xr = Xf + stride;
    const u32 gw = blockIdx.x * wpb + wid, nw = gridDim.x * wpb;
    u32 *scr = a.scratch + (size_t)gw * a.scratch_stride;
    const u32 count = a.list ? *a.count : a.n_direct;
    const bool fwd_only = (a.mode & 1u) != 0;

    for (u32 e = gw; e < count; e += nw) {
        const u32 rec = a.list ? a.list[e] : e;
        const u64 off = a.offsets[rec];
        const u32 n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
        if (a.list == nullptr && (n < a.min_n || n > a.max_n)) continue;  // direct mode: k_classify reported it
        const u64 *src = a.packed2 + ((off >> 5) + rec);
        RecordOut o;
        u64 h = 0;
        if (n < 64 || fwd_only) {
            // tiny records and the forward-only library calls: generic path
            RecordIn in; in.packed2 = src; in.bytes = nullptr; in.n = n;
            stage_record<2, G>(in, Xf, Xr);
            o = canonical_start<2, G>(Xf, Xr, n, scr, nullptr, fwd_only);
            const u32 *X = o.strand ? Xr : Xf;
            if (a.out) emit_ascii<2, G>(X, n, o.start, a.out + off);
            if (a.out_hash) h = xxh3_canonical<2, G>(X, n, o.start, nullptr, nullptr);
        } else {
            const u32 jn = n >> 4, rem = n & 15u;
            // ---- stage forward strand
            for (u32 k = lane; k < ((n + 31) >> 5); k += 32) {
                const uint2 w = __ldg(reinterpret_cast<const uint2 *>(src) + k);
                *reinterpret_cast<uint2 *>(Xf + 2 * k) = w;
            }
            __syncwarp();
            if (lane < 2) {                                   // circular extension: units jn, jn + 1
                u32 val;
                if (lane == 0) {
                    const u32 g0 = Xf[0];
                    val = rem ? ((Xf[jn] & ~(0xffffffffu >> (2 * rem))) | (g0 >> (2 * rem))) : g0;
                } else {
                    val = w2_window(Xf, 16u - rem);           // (jn + 1) * 16 - n
                }
                Xf[jn + lane] = val;
            }
            __syncwarp();
            // ---- reverse complement strand, units 0 .. jn + 1
            for (u32 j = lane; j < jn + 2; j += 32) {
                int t = (int)n - 16 * (int)(j + 1);
                if (t < 0) t += (int)n;
                Xr[j] = revcomp2_u32(w2_window(Xf, (u32)t));
            }
            __syncwarp();
            // ---- scan
            const u32 nsu = n >> 3, total = 2 * nsu, tail = n & 7u;
            u32 gmin, cnt, win_su = 0;
            u32 tail_key = 0xffffffffu;
            const bool tail_lane = lane < 16 && (lane & 7u) < tail;      // leftover rotations, one per lane
            if (tail_lane) tail_key = w2_window((lane >> 3) ? Xr : Xf, 8 * nsu + (lane & 7u));
            if (ROUNDS > 0) {
                u32 m[ROUNDS > 0 ? ROUNDS : 1];
#pragma unroll
                for (int r = 0; r < ROUNDS; r++) {
                    const u32 su = lane + 32 * r;
                    m[r] = su < total ? w2_step_min(Xf, stride, su, nsu) : 0xffffffffu;
                }
                u32 best = tail_key;
#pragma unroll
                for (int r = 0; r < ROUNDS; r++) best = min(best, m[r]);
                gmin = __reduce_min_sync(CK_FULL, best);
                cnt = 0;
#pragma unroll
                for (int r = 0; r < ROUNDS; r++) {
                    const u32 b = __ballot_sync(CK_FULL, m[r] == gmin && lane + 32 * r < total);
                    cnt += __popc(b);
                    if (b) win_su = (__ffs(b) - 1) + 32 * r;
                }
            } else {
                u32 best = 0xffffffffu, bestsu = 0xffffffffu, tie = 0;
                for (u32 su = lane; su < total; su += 32) {
                    const u32 m = w2_step_min(Xf, stride, su, nsu);
                    tie = (m == best) ? 1u : (m < best ? 0u : tie);
                    if (m < best) { best = m; bestsu = su; }
                }
                gmin = __reduce_min_sync(CK_FULL, min(best, tail_key));
                const u32 b = __ballot_sync(CK_FULL, best == gmin && bestsu != 0xffffffffu);
                const u32 t = __ballot_sync(CK_FULL, best == gmin && tie);
                cnt = __popc(b) + (t ? 2u : 0u);
                win_su = __shfl_sync(CK_FULL, bestsu, b ? __ffs(b) - 1 : 0);
            }
            const u32 tb = __ballot_sync(CK_FULL, tail_lane && tail_key == gmin);
            cnt += __popc(tb);
            if (cnt == 1) {
                if (tb) {
                    const u32 l = __ffs(tb) - 1;
                    o.strand = l >> 3; o.start = 8 * nsu + (l & 7u);
                } else {
                    // the winning step holds 8 rotations: find the first one carrying gmin
                    const u32 strand = win_su >= nsu ? 1u : 0u;
                    const u32 p0 = (strand ? win_su - nsu : win_su) * 8u;
                    const u32 k = w2_window(strand ? Xr : Xf, p0 + (lane & 7u));
                    const u32 hit = __ballot_sync(CK_FULL, k == gmin) & 0xffu;
                    o.strand = strand; o.start = p0 + (__ffs(hit) - 1);
                }
            } else {
                // ties (repeats, multimers, palindromic circles): duel-based generic path
                u32 f = strand_tie_winner<2, G>(Xf, n, gmin, scr, nullptr);
                u32 r = strand_tie_winner<2, G>(Xr, n, gmin, scr, nullptr);
                if (f == 0xffffffffu) { o.strand = 1; o.start = r; }
                else if (r == 0xffffffffu) { o.strand = 0; o.start = f; }
                else {
                    const bool fwd = forward_strictly_smaller<2, G>(Xf, Xr, n, f, r, nullptr);
                    o.strand = fwd ? 0u : 1u; o.start = fwd ? f : r;
                }
            }
            const u32 *X = o.strand ? Xr : Xf;
            if (a.out) w2_emit(X, lut, n, o.start, a.out + off);
            if (a.out_hash) h = n > 240 ? w2_xxh3_long(X, lut, n, o.start) : xxh3_short_warp<2>(X, n, o.start);
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
