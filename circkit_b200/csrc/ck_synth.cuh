// ck_synth.cuh -- device generators for the synthetic workloads of BASELINE.json (bench + tests).
// Everything is a pure function of (seed, record index), so any shard of any size can be
// regenerated independently on any rank.
#pragma once
#include "ck_device.cuh"

namespace ck {

__host__ __device__ __forceinline__ u64 splitmix64(u64 x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ u64 rnd(u64 seed, u64 i, u64 k)
{
    return splitmix64(splitmix64(seed ^ (i * 0xD1342543DE82EF95ULL)) + k * 0xA0761D6478BD642FULL);
}

// Duplicate structure: record i > 0 is a duplicate with probability dup_permille/1000; its origin is
// found by following "uniform earlier record" links until an original is hit.
__device__ __forceinline__ bool synth_is_dup(u64 seed, u64 i, u32 dup_permille)
{
    return i > 0 && (rnd(seed, i, 1) % 1000ULL) < dup_permille;
}
__device__ __forceinline__ u64 synth_origin(u64 seed, u64 i, u32 dup_permille)
{
    u64 j = i, hop = 0;
    while (synth_is_dup(seed, j, dup_permille)) { j = rnd(seed, j, 2 + (hop & 1)) % j; hop++; }
    return j;
}
__device__ __forceinline__ u32 synth_len(u64 seed, u64 origin, u32 kind, u32 lo, u32 hi)
{
    u64 r = rnd(seed, origin, 7);
    if (kind == 0) return lo + (u32)(r % (u64)(hi - lo + 1));
    // log-uniform: lo * (hi/lo)^u
    double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
    double v = exp(log((double)lo) + u * (log((double)hi) - log((double)lo)));
    u32 n = (u32)(v + 0.5);
    return n < lo ? lo : (n > hi ? hi : n);
}
// word k (32 bases, MSB-first) of ORIGINAL record `o`; content kinds:
//   0 iid ACGT; 1 tandem repeat of a short period / poly-A runs (adversarial, used for 1% of C4)
__device__ __forceinline__ u64 synth_word(u64 seed, u64 o, u32 k) { return rnd(seed, o, 16 + (u64)k); }
__device__ __forceinline__ u32 synth_base(u64 seed, u64 o, u32 t, u32 n, u32 adversarial_permille)
{
    if (adversarial_permille && (rnd(seed, o, 3) % 1000ULL) < adversarial_permille) {
        u64 r = rnd(seed, o, 4);
        u32 mode = (u32)(r % 7u);
        const u32 periods[6] = {1u, 2u, 3u, 7u, 171u, n / 2u ? n / 2u : 1u};
        if (mode < 6) {
            u32 p = periods[mode];
            u32 tp = t % p;
            return (u32)(synth_word(seed, o, tp >> 5) >> (62 - 2 * (tp & 31))) & 3u;
        }
        // mode 6: random sequence with poly-A runs >= 1 kb: positions inside [a, a + 1500) and [b, b + 1000) are A
        u32 a = (u32)((r >> 8) % n), b = (u32)((r >> 36) % n);
        if ((t + n - a) % n < 1500u || (t + n - b) % n < 1000u) return 0u;
    }
    return (u32)(synth_word(seed, o, t >> 5) >> (62 - 2 * (t & 31))) & 3u;
}

struct SynthArgs {
    u64 seed; u32 n_records; u32 kind, lo, hi; u32 dup_permille; u32 adversarial_permille;
    u64 first_index;         // global index of record 0 of this shard (multi-GPU shards)
};

__global__ void k_synth_lens(SynthArgs a, u64 *lens_out)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_records) return;
    u64 gi = a.first_index + i;
    lens_out[i] = synth_len(a.seed, synth_origin(a.seed, gi, a.dup_permille), a.kind, a.lo, a.hi);
}

// one warp per record, one lane per output word
__global__ void __launch_bounds__(256) k_synth_packed2(SynthArgs a, const u64 *offsets, u64 *packed2, u32 dbl)
{
    const u32 lane = threadIdx.x & 31u;
    const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = gw; i < a.n_records; i += nw) {
        const u64 gi = a.first_index + i;
        const u64 off = offsets[i];
        const u32 n = (u32)(offsets[i + 1] - off);
        u64 *dst = packed2 + p2_word(off, i, dbl);
        const u64 o = synth_origin(a.seed, gi, a.dup_permille);
        const bool dup = o != gi;
        const u32 W = (n + 31) >> 5;
        const bool plain = !dup && a.adversarial_permille == 0;
        u32 rot = 0; bool rc = false;
        if (dup) { u64 r = rnd(a.seed, gi, 5); rot = (u32)(r % n); rc = (r >> 40) & 1u; }
        for (u32 k = lane; k < W; k += 32) {
            u64 w;
            if (plain) {
                w = synth_word(a.seed, o, k);
                u32 valid = n - 32 * k;
                if (valid < 32) w &= ~0ULL << (64 - 2 * valid);
            } else {
                w = 0;
                for (u32 b = 0; b < 32; b++) {
                    u32 t = 32 * k + b;
                    u32 c = 0;
                    if (t < n) {
                        // duplicate = rotation by rot of the origin, then optionally reverse-complemented
                        u32 src = rc ? (n - 1 - t) : t;
                        src += rot; if (src >= n) src -= n;
                        c = synth_base(a.seed, o, src, n, a.adversarial_permille);
                        if (rc) c = 3u - c;
                    }
                    w = (w << 2) | c;
                }
            }
            dst[k] = (w << 32) | (w >> 32);      // arena order: unit of bases 0-15 at the lower address
        }
    }
}

// 2-bit arena -> ASCII arena
__global__ void __launch_bounds__(256) k_unpack2(const u64 *packed2, const u64 *offsets, u32 n_records, u8 *out, u32 dbl)
{
    const u32 lane = threadIdx.x & 31u;
    const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = gw; i < n_records; i += nw) {
        const u64 off = offsets[i];
        const u32 n = (u32)(offsets[i + 1] - off);
        const u64 *src = packed2 + p2_word(off, i, dbl);
        for (u32 t = lane; t < n; t += 32) {
            u32 c = (reinterpret_cast<const u32 *>(src)[t >> 4] >> (30 - 2 * (t & 15))) & 3u;
            out[off + t] = "ACGT"[c];
        }
    }
}

}  // namespace ck
