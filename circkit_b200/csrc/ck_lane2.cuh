// ck_lane2.cuh -- building blocks of the lane-per-record kernel (ck_stream2.cuh): the 16-letter ASCII generator with
// per-lane letter table / selectors (the reverse strand costs nothing extra), the (key, step, strand) tracker, the
// per-stripe XXH3 accumulation, shared-memory / cp.async helpers.
#pragma once
#include "ck_warp2.cuh"

namespace ck {

__constant__ u64 c_lastsec[8];      // LE u64 of the secret at byte offsets 121 + 8 i (XXH3 last stripe)
__constant__ u64 c_mergesec[8];     // LE u64 of the secret at byte offsets 11 + 8 i  (XXH3 merge)
__constant__ u64 c_midsec[16];      // LE u64 of the secret at byte offsets 3 + 8 i   (XXH3 129..240: rounds 8.., start offset 3);
                                    // [14], [15]: offsets 119, 127 (the last 16 bytes: 136 - 17)

__device__ __forceinline__ uint4 lds128(u32 a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(u32 a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// bytes of the per-warp output stage: 32 lanes x 64 bytes at a stride of 80 (conflict-free 128-bit accesses)
#define CK_T2_AUX_BYTES 2560u

__device__ __forceinline__ void cp_async8(u32 dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(u32 dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 16 canonical letters from a 16-base window; T = letter table, (sa, sb) = interleave selectors.  Forward strand:
// ("ACGT", 0x2637, 0x0415).  Reverse strand: the window is rotated by 16 bits first and ("TGCA", 0x5140, 0x7362)
// complement and reverse it.
__device__ __forceinline__ uint4 t2_ascii16(u32 w, u32 T, u32 sa, u32 sb)
{
    const u32 E = w & 0x33333333u, O = (w >> 2) & 0x33333333u;
    const u32 pe_lo = __byte_perm(T, 0, E), pe_hi = __byte_perm(T, 0, E >> 16);
    const u32 po_lo = __byte_perm(T, 0, O), po_hi = __byte_perm(T, 0, O >> 16);
    uint4 v;
    v.x = __byte_perm(pe_hi, po_hi, sa);
    v.y = __byte_perm(pe_hi, po_hi, sb);
    v.z = __byte_perm(pe_lo, po_lo, sa);
    v.w = __byte_perm(pe_lo, po_lo, sb);
    return v;
}

// the two smallest (key << 16 | tag) values seen so far
__device__ __forceinline__ void t2_track(u32 &m1, u32 &m2, u32 v)
{
    m2 = min(m2, max(m1, v));
    m1 = min(m1, v);
}
// u16x2 step minimum -> (min of both halves) << 16
__device__ __forceinline__ u32 t2_key_hi(u32 m)
{
    const u32 sw = __byte_perm(m, 0, 0x1032);
    return __vminu2(m, sw) & 0xffff0000u;
}

// one XXH3 stripe piece: 16 input bytes (v) against secret words (k0, k1); accumulators (a0, a1) = acc[2k], acc[2k+1]
__device__ __forceinline__ void t2_acc16(u64 &a0, u64 &a1, uint4 v, u64 k0, u64 k1)
{
    const u32 x0 = v.x ^ (u32)k0, x1 = v.y ^ (u32)(k0 >> 32), y0 = v.z ^ (u32)k1, y1 = v.w ^ (u32)(k1 >> 32);
    a0 += (u64)x0 * (u64)x1 + (((u64)v.w << 32) | v.z);
    a1 += (u64)y0 * (u64)y1 + (((u64)v.y << 32) | v.x);
}

}  // namespace ck
