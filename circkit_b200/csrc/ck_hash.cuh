// ck_hash.cuh -- XXH3-64 (seed 0, default secret) of the canonical ASCII, computed from the packed
// strand in shared memory: the ASCII bytes only ever exist in registers.
// Replaces xxhash_rust::xxh3::xxh3_64(canonicalized) on the consumer thread (src/uniq.rs:45).
// Length classes follow the XXH3 spec: 0, 1-3, 4-8, 9-16, 17-128, 129-240, >240 (1024-byte blocks
// of 16 stripes with a scramble between blocks, then the last stripe and the merge).
#pragma once
#include "ck_record.cuh"

namespace ck {

__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(CK_FULL, v, d);
    return v;
}

// canonical byte t (t < n) as an integer
template <int BITS> __device__ __forceinline__ u32 canon_byte(const u32 *X, u32 n, u32 start, u32 t)
{
    return (u32)(ascii8<BITS>(X, n, start, t) & 0xffu);
}
// XXH_readLE64(canonical + t), t + 8 <= n
template <int BITS> __device__ __forceinline__ u64 canon64(const u32 *X, u32 n, u32 start, u32 t)
{
    return ascii8<BITS>(X, n, start, t);
}
template <int BITS> __device__ __forceinline__ u32 canon32(const u32 *X, u32 n, u32 start, u32 t)
{
    return (u32)ascii8<BITS>(X, n, start, t);
}

template <int BITS> __device__ __forceinline__ u64 mix16(const u32 *X, u32 n, u32 st, u32 t, int soff)
{
    return mul128_fold64(canon64<BITS>(X, n, st, t) ^ sec64(soff), canon64<BITS>(X, n, st, t + 8) ^ sec64(soff + 8));
}

// n <= 240, executed by one full warp (all lanes must call); result valid in every lane.
template <int BITS> __device__ u64 xxh3_short_warp(const u32 *X, u32 n, u32 st)
{
    const u32 lane = lane_id();
    if (n == 0) return xxh64_avalanche(sec64(56) ^ sec64(64));
    if (n <= 3) {
        u32 c1 = canon_byte<BITS>(X, n, st, 0), c2 = canon_byte<BITS>(X, n, st, n >> 1), c3 = canon_byte<BITS>(X, n, st, n - 1);
        u32 comb = (c1 << 16) | (c2 << 24) | c3 | (n << 8);
        u64 flip = (u64)((u32)sec64(0) ^ (u32)(sec64(0) >> 32));
        return xxh64_avalanche((u64)comb ^ flip);
    }
    if (n <= 8) {
        // reads of 4 bytes at t and n-4: ascii8 wraps past n, only the low 4 bytes are used
        u32 i1 = canon32<BITS>(X, n, st, 0), i2 = canon32<BITS>(X, n, st, n - 4);
        u64 flip = sec64(8) ^ sec64(16);
        u64 in64 = (u64)i2 + ((u64)i1 << 32);
        u64 h = in64 ^ flip;
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= CK_PMX2;
        h ^= (h >> 35) + (u64)n;
        h *= CK_PMX2;
        return h ^ (h >> 28);
    }
    if (n <= 16) {
        u64 f1 = sec64(24) ^ sec64(32), f2 = sec64(40) ^ sec64(48);
        u64 lo = canon64<BITS>(X, n, st, 0) ^ f1, hi = canon64<BITS>(X, n, st, n - 8) ^ f2;
        u64 acc = (u64)n + bswap64(lo) + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (n <= 128) {
        // pairs i = 0..(n-1)/32: lane 2i -> front 16 bytes, lane 2i+1 -> back 16 bytes
        u32 npairs = (n - 1) / 32 + 1;
        u64 v = 0;
        if (lane < 2 * npairs) {
            u32 i = lane >> 1;
            v = (lane & 1u) ? mix16<BITS>(X, n, st, n - 16 * (i + 1), 32 * i + 16) : mix16<BITS>(X, n, st, 16 * i, 32 * i);
        }
        u64 acc = (u64)n * CK_P64_1 + warp_sum_u64(v);
        return xxh3_avalanche(acc);
    }
    {   // 129..240
        u32 rounds = n / 16;
        u64 a = 0, b = 0;
        if (lane < 8) a = mix16<BITS>(X, n, st, 16 * lane, 16 * lane);
        else if (lane < rounds) b = mix16<BITS>(X, n, st, 16 * lane, 16 * (lane - 8) + 3);
        else if (lane == 31) b = mix16<BITS>(X, n, st, n - 16, 136 - 17);
        u64 acc = xxh3_avalanche((u64)n * CK_P64_1 + warp_sum_u64(a));
        u64 acc_end = warp_sum_u64(b);
        return xxh3_avalanche(acc + acc_end);
    }
}

// Contribution of `nstripes` consecutive stripes starting at canonical byte `base` to the eight
// accumulators (one warp; lane i < 8 returns what must be ADDED to acc[i]).  Secret offset of
// stripe s is 8*s (XXH_SECRET_CONSUME_RATE), so the read is an aligned u64 of the secret.
template <int BITS>
__device__ __forceinline__ u64 xxh3_stripes_warp(const u32 *X, u32 n, u32 st, u32 base, u32 nstripes)
{
    const u32 lane = lane_id(), i = lane & 7u, sg = lane >> 3;
    const u64 *sec = reinterpret_cast<const u64 *>(c_secret);
    u64 mul = 0, dv_sum = 0;
    for (u32 s = sg; s < nstripes; s += 4) {
        u64 dv = canon64<BITS>(X, n, st, base + 64 * s + 8 * i);
        u64 dk = dv ^ sec[s + i];
        mul += (dk & 0xffffffffULL) * (dk >> 32);
        dv_sum += dv;
    }
    mul += __shfl_xor_sync(CK_FULL, mul, 8);  dv_sum += __shfl_xor_sync(CK_FULL, dv_sum, 8);
    mul += __shfl_xor_sync(CK_FULL, mul, 16); dv_sum += __shfl_xor_sync(CK_FULL, dv_sum, 16);
    u64 other = __shfl_xor_sync(CK_FULL, dv_sum, 1);          // acc[i ^ 1] += data_val
    return mul + other;
}
__device__ __forceinline__ u64 xxh3_scramble(u64 a, u32 i)
{
    const u64 *sec = reinterpret_cast<const u64 *>(c_secret);
    a ^= a >> 47; a ^= sec[(192 - 64) / 8 + i]; a *= CK_P32_1;
    return a;
}
__device__ __forceinline__ u64 xxh3_init_acc(u32 i)
{
    switch (i) {
    case 0: return CK_P32_3; case 1: return CK_P64_1; case 2: return CK_P64_2; case 3: return CK_P64_3;
    case 4: return CK_P64_4; case 5: return CK_P32_2; case 6: return CK_P64_5; default: return CK_P32_1;
    }
}
// last stripe + merge; acc valid in lanes 0..7 of the calling warp; result valid in every lane.
template <int BITS> __device__ __forceinline__ u64 xxh3_finish_warp(const u32 *X, u32 n, u32 st, u64 acc)
{
    const u32 lane = lane_id(), i = lane & 7u;
    {   // last stripe: input + n - 64, secret + 192 - 64 - 7
        u64 dv = canon64<BITS>(X, n, st, n - 64 + 8 * i);
        u64 dk = dv ^ sec64(121 + 8 * (int)i);
        u64 other = __shfl_xor_sync(CK_FULL, dv, 1);
        acc += (dk & 0xffffffffULL) * (dk >> 32) + other;
    }
    // merge: sum_k mul128_fold64(acc[2k] ^ sec(11+16k), acc[2k+1] ^ sec(11+16k+8))
    u64 keyed = acc ^ sec64(11 + 8 * (int)i);
    u64 partner = __shfl_xor_sync(CK_FULL, keyed, 1);
    u64 term = (lane < 8 && (lane & 1u) == 0) ? mul128_fold64(keyed, partner) : 0ULL;
    u64 r = (u64)n * CK_P64_1 + warp_sum_u64(term);
    return xxh3_avalanche(r);
}

// n > 240, one warp does everything (warp-per-record kernels).
template <int BITS> __device__ u64 xxh3_long_warp(const u32 *X, u32 n, u32 st)
{
    const u32 i = lane_id() & 7u;
    u64 acc = xxh3_init_acc(i);
    const u32 nb_blocks = (n - 1) / 1024;
    for (u32 b = 0; b < nb_blocks; b++) {
        acc += xxh3_stripes_warp<BITS>(X, n, st, b * 1024, 16);
        acc = xxh3_scramble(acc, i);
    }
    const u32 nstripes = ((n - 1) - 1024 * nb_blocks) / 64;
    acc += xxh3_stripes_warp<BITS>(X, n, st, nb_blocks * 1024, nstripes);
    return xxh3_finish_warp<BITS>(X, n, st, acc);
}

// Group-level entry.  hbuf: shared u64[32][8] (CTA shape only).  Result valid in every thread.
template <int BITS, typename G>
__device__ u64 xxh3_canonical(const u32 *X, u32 n, u32 st, u64 *hbuf, u32 *red)
{
    if (!G::kCta) return n <= 240 ? xxh3_short_warp<BITS>(X, n, st) : xxh3_long_warp<BITS>(X, n, st);
    const u32 wid = threadIdx.x >> 5, lane = lane_id(), i = lane & 7u, nw = blockDim.x >> 5;
    u64 h = 0;
    if (n <= 240) {
        if (wid == 0) h = xxh3_short_warp<BITS>(X, n, st);
    } else {
        const u32 nb_blocks = (n - 1) / 1024;
        const u32 last_stripes = ((n - 1) - 1024 * nb_blocks) / 64;
        u64 acc = xxh3_init_acc(i);                      // meaningful in warp 0 only
        for (u32 b0 = 0; b0 <= nb_blocks; b0 += nw) {
            u32 b = b0 + wid;
            if (b <= nb_blocks) {
                u64 c = xxh3_stripes_warp<BITS>(X, n, st, b * 1024, b < nb_blocks ? 16u : last_stripes);
                if (lane < 8) hbuf[wid * 8 + lane] = c;
            }
            __syncthreads();
            if (wid == 0) {
                u32 cnt = nb_blocks + 1 - b0; if (cnt > nw) cnt = nw;
                for (u32 w = 0; w < cnt; w++) {
                    acc += hbuf[w * 8 + i];
                    if (b0 + w < nb_blocks) acc = xxh3_scramble(acc, i);
                }
            }
            __syncthreads();
        }
        if (wid == 0) h = xxh3_finish_warp<BITS>(X, n, st, acc);
    }
    // broadcast from warp 0
    __syncthreads();
    if (threadIdx.x == 0) { red[32] = (u32)h; red[33] = (u32)(h >> 32); }
    __syncthreads();
    h = ((u64)red[33] << 32) | red[32];
    return h;
}

}  // namespace ck
