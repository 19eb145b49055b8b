// ck_monomerize.cuh -- `monomerize` on the device: for every record the end index of its last monomer.
//
// What it replaces in the reference (per record):
//   Monomerizer::first_monomer_end_index            lib/src/monomerize.rs:50-99
//   Monomerizer::last_monomer_end_index             lib/src/monomerize.rs:100-125
//   Monomerizer::last_monomer_end_index_sensitive   lib/src/monomerize.rs:127-141
//
// One warp per record, bytes as they are (library semantics).  first(): the seed is the last seed_len bytes; the 32 lanes
// test 128 candidate start positions of the text seq[.. n - seed_len] at a time (four per lane) (bio's ShiftAnd::find_all reports every
// occurrence, overlapping ones included, by increasing start), the occurrences of a round are taken in position order and
// the Hamming distance of prefix seq[.. occ + seed_len] against the suffix of the same length is counted by the whole
// warp (four bytes per lane and step); the first occurrence within the allowed distance ends the search.  The sensitive pass runs the same code over a
// view that reads the monomer backwards through bio's complement table (no reverse complement is materialised).
#pragma once
#include "ck_device.cuh"

namespace ck {

#ifndef CK_MONO_NONE
#define CK_MONO_NONE 0xffffffffu
#endif

struct MonoArgs {
    const u8 *bytes; const u64 *offsets; const u32 *lens; u32 n_records;   // lens: optional normalised lengths (ck_dev_normalize)
    u32 seed_len;          // 1 .. 63
    u32 use_identity;      // 1: max distance = len - floor(len * identity) (f64, as the reference computes it)
    u64 overlap_dist;      // 0: otherwise
    double identity;
    u32 flags;             // bit0: sensitive (reverse-complement pass), bit1: first_monomer_end_index only
    u32 *out_end;          // CK_MONO_NONE = None
};

template <bool RC> struct MonoView {
    const u8 *s; u32 n;    // RC: position j reads comp[s[n - 1 - j]]
    __device__ __forceinline__ u32 at(u32 j) const { return RC ? (u32)s_tab.comp[s[n - 1u - j]] : (u32)s[j]; }
};

template <bool RC> __device__ u32 mono_first(const MonoArgs &a, MonoView<RC> v)
{
    const u32 s = a.seed_len, n = v.n, lane = lane_id();
    if (n <= s || n < 2 * s) return CK_MONO_NONE;          // no room for an occurrence in seq[.. n - s]
    const u32 limit = n - 2 * s + 1;                        // candidate starts 0 .. n - 2 s
    // the first K = min(seed_len, 4) bytes of the seed as one word: a candidate start must match it before the rest of the
    // seed is looked at (1 in 256 random positions passes instead of 1 in 4, so the divergent tail below is rare)
    const u32 K = min(s, 4u);
    const u32 kmask = K == 4 ? 0xffffffffu : (1u << (8 * K)) - 1u;
    u32 seedw = 0;
    for (u32 k = 0; k < K; k++) seedw |= v.at(n - s + k) << (8 * k);
    // 128 candidate starts per round, four consecutive ones per lane (seven independent byte loads)
    for (u32 base = 0; base < limit; base += 128) {
        const u32 p0 = base + 4 * lane;
        u32 hit = 0;
        if (p0 < limit) {
            u32 b[7];
#pragma unroll
            for (u32 k = 0; k < 7; k++) b[k] = p0 + k < n ? v.at(p0 + k) : 0u;
#pragma unroll
            for (u32 q = 0; q < 4; q++) {
                const u32 w = (b[q] | (b[q + 1] << 8) | (b[q + 2] << 16) | (b[q + 3] << 24)) & kmask;
                if (p0 + q < limit && w == seedw) hit |= 1u << q;
            }
        }
        for (u32 h = hit; h; h &= h - 1) {                  // the rest of the seed for the positions that passed
            const u32 q = (u32)__ffs(h) - 1u, p = p0 + q;
            bool match = true;
            for (u32 k = K; match && k < s; k++) match = v.at(p + k) == v.at(n - s + k);
            if (!match) hit &= ~(1u << q);
        }
        u32 lanes = __ballot_sync(CK_FULL, hit != 0);
        while (lanes) {                                     // occurrences in position order: by lane, then inside the lane
            const u32 l = (u32)__ffs(lanes) - 1u;
            lanes &= lanes - 1u;
            u32 hq = __shfl_sync(CK_FULL, hit, l);
            while (hq) {
                const u32 occ = base + 4 * l + (u32)__ffs(hq) - 1u;
                hq &= hq - 1u;
                const u32 m = occ + s;                      // overlap length
                u32 cnt = 0;
                for (u32 t = 4 * lane; t < m; t += 128) {   // four bytes per lane and step
#pragma unroll
                    for (u32 u = 0; u < 4; u++)
                        if (t + u < m) cnt += v.at(n - m + t + u) != v.at(t + u);
                }
                const u64 dist = __reduce_add_sync(CK_FULL, cnt);
                const u64 maxd = a.use_identity ? (u64)m - (u64)floor((double)m * a.identity) : a.overlap_dist;
                if (dist <= maxd) return n - m;
            }
        }
    }
    return CK_MONO_NONE;
}

__global__ void __launch_bounds__(256) k_monomerize(MonoArgs a)
{
    stab_load();
    const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (u32 rec = gw; rec < a.n_records; rec += nw) {
        const u64 off = a.offsets[rec];
        const u32 n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
        const u8 *seq = a.bytes + off;
        u32 idx = mono_first<false>(a, MonoView<false>{seq, n});
        if (!(a.flags & 2u)) {
            while (idx != CK_MONO_NONE) {                   // lib/src/monomerize.rs:103-114
                const u32 nxt = mono_first<false>(a, MonoView<false>{seq, idx});
                if (nxt == CK_MONO_NONE) break;
                idx = nxt;
            }
            if (a.flags & 1u) {                             // lib/src/monomerize.rs:127-141: Some(k) -> M - (M - k) = k
                const u32 M = idx == CK_MONO_NONE ? n : idx;
                const u32 k = mono_first<true>(a, MonoView<true>{seq, M});
                if (k != CK_MONO_NONE) idx = k;
            }
        }
        if (lane_id() == 0) a.out_end[rec] = idx;
    }
}

}  // namespace ck
