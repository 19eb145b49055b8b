// ck_monomerize.cuh -- `monomerize` on the device: for every record the end index of its last monomer.
//
// What it replaces in the reference (per record):
//   Monomerizer::first_monomer_end_index            lib/src/monomerize.rs:50-99
//   Monomerizer::last_monomer_end_index             lib/src/monomerize.rs:100-125
//   Monomerizer::last_monomer_end_index_sensitive   lib/src/monomerize.rs:127-141
//
// One warp per record, bytes as they are (library semantics).  first(): the seed is the last seed_len bytes; the 32 lanes
// test 32 candidate start positions of the text seq[.. n - seed_len] at a time (bio's ShiftAnd::find_all reports every
// occurrence, overlapping ones included, by increasing start), the occurrences of a round are taken in position order and
// the Hamming distance of prefix seq[.. occ + seed_len] against the suffix of the same length is counted by the whole
// warp; the first occurrence within the allowed distance ends the search.  The sensitive pass runs the same code over a
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
    for (u32 base = 0; base < limit; base += 32) {
        const u32 p = base + lane;
        bool match = p < limit;
        for (u32 k = 0; match && k < s; k++) match = v.at(p + k) == v.at(n - s + k);
        u32 mask = __ballot_sync(CK_FULL, match);
        while (mask) {
            const u32 occ = base + (u32)__ffs(mask) - 1u;
            mask &= mask - 1u;
            const u32 m = occ + s;                          // overlap length
            u32 cnt = 0;
            for (u32 t = lane; t < m; t += 32) cnt += v.at(n - m + t) != v.at(t);
            const u64 dist = __reduce_add_sync(CK_FULL, cnt);
            const u64 maxd = a.use_identity ? (u64)m - (u64)floor((double)m * a.identity) : a.overlap_dist;
            if (dist <= maxd) return n - m;
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
