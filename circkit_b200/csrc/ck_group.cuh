// ck_group.cuh -- the two cooperating-thread shapes the record kernels run in:
//   Grp<false> : one warp per record   (viroid / circRNA lengths; shuffles and REDUX only)
//   Grp<true>  : one CTA per record    (plasmid / mtDNA lengths; warp step + shared-memory step)
// Both expose the same collective vocabulary so the LMSR algorithm is written once.
#pragma once
#include "ck_device.cuh"

namespace ck {

template <bool CTA> struct Grp;

template <> struct Grp<false> {
    static constexpr bool kCta = false;
    __device__ __forceinline__ static u32 size() { return 32u; }
    __device__ __forceinline__ static u32 rank() { return threadIdx.x & 31u; }
    __device__ __forceinline__ static void sync() { __syncwarp(); }
    __device__ __forceinline__ static u32 min_u32(u32 v, u32 *) { return __reduce_min_sync(CK_FULL, v); }
    __device__ __forceinline__ static u32 sum_u32(u32 v, u32 *) { return __reduce_add_sync(CK_FULL, v); }
    __device__ __forceinline__ static u32 or_u32(u32 v, u32 *) { return __reduce_or_sync(CK_FULL, v); }
    // exclusive prefix sum over ranks; total returned through `total`
    __device__ __forceinline__ static u32 exscan_u32(u32 v, u32 &total, u32 *)
    {
        u32 x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(CK_FULL, x, d);
            if ((threadIdx.x & 31u) >= (u32)d) x += y;
        }
        total = __shfl_sync(CK_FULL, x, 31);
        return x - v;
    }
};

// CTA collectives use a caller-provided shared scratch `red` of >= 34 u32.
template <> struct Grp<true> {
    static constexpr bool kCta = true;
    __device__ __forceinline__ static u32 size() { return blockDim.x; }
    __device__ __forceinline__ static u32 rank() { return threadIdx.x; }
    __device__ __forceinline__ static void sync() { __syncthreads(); }
    template <typename Op> __device__ __forceinline__ static u32 reduce(u32 v, u32 *red, Op op, u32 ident)
    {
        u32 w = op.warp(v);
        u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
        __syncthreads();                       // protect `red` from the previous collective's readers
        if (lane == 0) red[wid] = w;
        __syncthreads();
        if (wid == 0) {
            u32 x = lane < nw ? red[lane] : ident;
            x = op.warp(x);
            if (lane == 0) red[32] = x;
        }
        __syncthreads();
        return red[32];
    }
    struct OpMin { __device__ __forceinline__ u32 warp(u32 v) const { return __reduce_min_sync(CK_FULL, v); } };
    struct OpAdd { __device__ __forceinline__ u32 warp(u32 v) const { return __reduce_add_sync(CK_FULL, v); } };
    struct OpOr { __device__ __forceinline__ u32 warp(u32 v) const { return __reduce_or_sync(CK_FULL, v); } };
    __device__ __forceinline__ static u32 min_u32(u32 v, u32 *red) { return reduce(v, red, OpMin(), 0xffffffffu); }
    __device__ __forceinline__ static u32 sum_u32(u32 v, u32 *red) { return reduce(v, red, OpAdd(), 0u); }
    __device__ __forceinline__ static u32 or_u32(u32 v, u32 *red) { return reduce(v, red, OpOr(), 0u); }
    __device__ __forceinline__ static u32 exscan_u32(u32 v, u32 &total, u32 *red)
    {
        u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
        u32 x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(CK_FULL, x, d);
            if (lane >= (u32)d) x += y;
        }
        __syncthreads();
        if (lane == 31) red[wid] = x;
        __syncthreads();
        if (wid == 0) {
            u32 s = lane < nw ? red[lane] : 0u, t = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 y = __shfl_up_sync(CK_FULL, t, d);
                if (lane >= (u32)d) t += y;
            }
            red[lane] = t - s;                 // exclusive warp offsets
            if (lane == 31) red[32] = t;
        }
        __syncthreads();
        total = red[32];
        return red[wid] + x - v;
    }
};

// 64-bit min over a group, as two 32-bit rounds (high word, then low word among the high-word winners).
template <typename G> __device__ __forceinline__ u64 grp_min_u64(u64 v, u32 *red)
{
    u32 hi = (u32)(v >> 32), lo = (u32)v;
    u32 mh = G::min_u32(hi, red);
    u32 ml = G::min_u32(hi == mh ? lo : 0xffffffffu, red);
    return ((u64)mh << 32) | ml;
}
template <typename G> __device__ __forceinline__ u32 grp_min_key(u32 v, u32 *red) { return G::min_u32(v, red); }
template <typename G> __device__ __forceinline__ u64 grp_min_key(u64 v, u32 *red) { return grp_min_u64<G>(v, red); }

}  // namespace ck
