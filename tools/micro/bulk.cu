// Microbenchmark: lane-private record streaming on sm_100a through the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP)
// against lane-private LDG.128, and lane-private output through bulk stores against the staged STS/LDS/STG.128 path.
// Question it answers for k_canon_s2: every lane walks ITS OWN record; a divergent LDG.128 costs 32 L1TEX wavefronts per
// warp instruction (the kernel's L1TEX data pipe is 76 % busy, profiles/r01_l_c2_k_canon_s2.txt).  Can per-lane bulk copies
// of CH bytes into a shared-memory ring feed the lanes instead, and at what cost in issue slots?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bulk bulk.cu && ./bulk
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32; typedef unsigned long long u64;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, u32 src, u32 bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b)
{
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, uint4 a, uint4 b)
{
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// a little ALU work per 16 bytes so that the loads have something to hide behind (WORK dependent ops per quad)
template <int WORK> __device__ __forceinline__ u32 chew(uint4 v, u32 acc)
{
    u32 x = v.x ^ v.y ^ v.z ^ v.w;
#pragma unroll
    for (int i = 0; i < WORK; i++) x = __funnelshift_l(x, acc, 7) ^ (x >> 3);
    return acc + x;
}

// ---- reads.  rows of `quads` 16-byte quads; lane l of warp batch b walks row b + l.
// MODE 0: lane-private LDG.128, loads 3 ahead in registers
// MODE 1: lane-private bulk copies of CH bytes into a per-lane ring of K slots, one mbarrier per (lane, slot)
// MODE 2: the same with ONE mbarrier per (warp, slot): lane 0 expects 32 * CH bytes, every lane issues its copy
template <int MODE, int CH, int K, int WORK, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_read(const uint4 *arena, u32 rows, u32 quads, u32 *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 gw = blockIdx.x * WARPS + wid, nw = gridDim.x * WARPS;
    u32 acc = 0;
    if (MODE == 0) {
        for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
            const uint4 *p = arena + (size_t)(b + lane) * quads;
            uint4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
            for (u32 i = 0; i < quads; i++) {
                uint4 q3 = __ldg(p + min(i + 3, quads - 1));
                acc = chew<WORK>(q0, acc);
                q0 = q1; q1 = q2; q2 = q3;
            }
        }
    } else if (MODE == 3) {
        for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
            const unsigned char *p = reinterpret_cast<const unsigned char *>(arena + (size_t)(b + lane) * quads);
            const u32 octs = quads / 2;
            uint4 a0, a1, b0, b1;
            ldg256(p, a0, a1); ldg256(p + 32 * min(1u, octs - 1), b0, b1);
            for (u32 i = 0; i < octs; i++) {
                uint4 c0, c1;
                ldg256(p + 32 * min(i + 2, octs - 1), c0, c1);
                acc = chew<WORK>(a0, acc); acc = chew<WORK>(a1, acc);
                a0 = b0; a1 = b1; b0 = c0; b1 = c1;
            }
        }
    } else {
        // per warp: ring rows of K * CH bytes per lane at a stride that keeps 8 lanes on distinct 16-byte bank groups
        constexpr u32 STRIDE = K * CH + 16;
        const u32 ring = smem_u32(smem) + wid * (32u * STRIDE + 8u * 32u * K) + lane * STRIDE;
        const u32 bars = smem_u32(smem) + wid * (32u * STRIDE + 8u * 32u * K) + 32u * STRIDE;     // K * 32 mbarriers
        const u32 mybar = MODE == 1 ? bars + 8u * (lane * K) : bars;                               // + 8 * slot
        if (MODE == 1) { for (u32 s = 0; s < K; s++) mbar_init(mybar + 8 * s, 1); }
        else if (lane == 0) { for (u32 s = 0; s < K; s++) mbar_init(mybar + 8 * s, 1); }
        __syncwarp();
        constexpr u32 QPC = CH / 16;
        u32 par = 0;                                           // bit s: parity to wait for on slot s
        for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
            const unsigned char *p = reinterpret_cast<const unsigned char *>(arena + (size_t)(b + lane) * quads);
            const u32 chunks = quads / QPC;
            for (u32 c = 0; c < K - 1 && c < chunks; c++) {
                if (MODE == 1) mbar_expect_tx(mybar + 8 * c, CH);
                else if (lane == 0) mbar_expect_tx(mybar + 8 * c, 32 * CH);
                bulk_g2s(ring + c * CH, p + c * CH, CH, mybar + 8 * c);
            }
            for (u32 c = 0; c < chunks; c++) {
                const u32 s = c % K;
                const u32 cn = c + K - 1;
                if (cn < chunks) {
                    const u32 sn = cn % K;
                    if (MODE == 1) mbar_expect_tx(mybar + 8 * sn, CH);
                    else if (lane == 0) mbar_expect_tx(mybar + 8 * sn, 32 * CH);
                    bulk_g2s(ring + sn * CH, p + cn * CH, CH, mybar + 8 * sn);
                }
                mbar_wait(mybar + 8 * s, (par >> s) & 1u);
                par ^= 1u << s;
#pragma unroll
                for (u32 q = 0; q < QPC; q++) acc = chew<WORK>(lds128(ring + s * CH + 16 * q), acc);
                __syncwarp();
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- writes.  every lane produces `quads` quads for its own row.
// MODE 0: 64 bytes per lane per round through a shared-memory stage, 128-bit stores in which 4 lanes cover 64 contiguous bytes
// MODE 1: CH bytes per lane accumulate in the lane's shared-memory row (two rows, alternating), one bulk store per row
template <int MODE, int CH, int WORK, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_write(uint4 *arena, u32 rows, u32 quads, u32 seed)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const u32 gw = blockIdx.x * WARPS + wid, nw = gridDim.x * WARPS;
    u32 acc = seed + threadIdx.x;
    if (MODE == 0) {
        const u32 st = smem_u32(smem) + wid * 2048u;
        // conflict-free both ways: row r, chunk c -> line r >> 1, slot (c + 4 (r & 1) + ((r >> 1) & 3)) & 7
        const u32 wr = st + 128u * (lane >> 1);
        const u32 wj = 4u * (lane & 1u) + ((lane >> 1) & 3u);
        for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
            for (u32 i = 0; i < quads; i += 4) {
#pragma unroll
                for (u32 c = 0; c < 4; c++) {
                    acc = chew<WORK>(make_uint4(acc, i, c, b), acc);
                    sts128(wr + 16u * ((c + wj) & 7u), make_uint4(acc, acc + 1, acc + 2, acc + 3));
                }
                __syncwarp();
#pragma unroll
                for (u32 k = 0; k < 4; k++) {
                    const u32 r = 8 * k + (lane >> 2), c = lane & 3u;
                    const uint4 v = lds128(st + 128u * (r >> 1) + 16u * ((c + 4u * (r & 1u) + ((r >> 1) & 3u)) & 7u));
                    arena[(size_t)(b + r) * quads + i + c] = v;
                }
                __syncwarp();
            }
        }
    } else if (MODE == 2) {
        const u32 st = smem_u32(smem) + wid * 2048u;
        const u32 wr = st + 128u * (lane >> 1);
        const u32 wj = 4u * (lane & 1u) + ((lane >> 1) & 3u);
        for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
            for (u32 i = 0; i < quads; i += 4) {
#pragma unroll
                for (u32 c = 0; c < 4; c++) {
                    acc = chew<WORK>(make_uint4(acc, i, c, b), acc);
                    sts128(wr + 16u * ((c + wj) & 7u), make_uint4(acc, acc + 1, acc + 2, acc + 3));
                }
                __syncwarp();
#pragma unroll
                for (u32 k = 0; k < 2; k++) {
                    const u32 r = 16 * k + (lane >> 1), c = 2u * (lane & 1u);
                    const u32 ln = st + 128u * (r >> 1), rot = 4u * (r & 1u) + ((r >> 1) & 3u);
                    const uint4 v0 = lds128(ln + 16u * ((c + rot) & 7u)), v1 = lds128(ln + 16u * ((c + 1 + rot) & 7u));
                    stg256(arena + (size_t)(b + r) * quads + i + c, v0, v1);
                }
                __syncwarp();
            }
        }
    } else {
        constexpr u32 STRIDE = 2 * CH + 16;
        const u32 row = smem_u32(smem) + wid * 32u * STRIDE + lane * STRIDE;
        constexpr u32 QPC = CH / 16;
        u32 half = 0;
        for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(arena + (size_t)(b + lane) * quads);
            for (u32 i = 0; i < quads; i += QPC) {
                bulk_wait_read<1>();                           // the store that last read this half has finished with it
#pragma unroll
                for (u32 q = 0; q < QPC; q++) {
                    acc = chew<WORK>(make_uint4(acc, i, q, b), acc);
                    sts128(row + half * CH + 16 * q, make_uint4(acc, acc + 1, acc + 2, acc + 3));
                }
                fence_async_smem();
                bulk_s2g(dst + 16 * i, row + half * CH, CH);
                bulk_commit();
                half ^= 1u;
            }
        }
        bulk_wait_read<0>();
    }
    if (acc == 0x12345u) arena[0].x = acc;
}

template <typename F> static float best_of(F f, int reps = 3)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

template <int MODE, int CH, int K, int WORK, int WARPS> static void run_read(const char *name, const uint4 *arena, size_t bytes, u32 rowbytes, u32 *out, int sms, int ctas)
{
    const u32 quads = rowbytes / 16, rows = (u32)(bytes / rowbytes) & ~31u;
    const u32 smem = (MODE == 0 || MODE == 3) ? 0 : WARPS * (32u * (K * CH + 16) + 8u * 32u * K);
    cudaFuncSetAttribute(k_read<MODE, CH, K, WORK, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float ms = best_of([&] { k_read<MODE, CH, K, WORK, WARPS><<<ctas * sms, 32 * WARPS, smem>>>(arena, rows, quads, out); });
    printf("read  row %5u B work %2d  %-34s %2d warps/SM %8.3f ms %7.1f GB/s  (%s)\n", rowbytes, WORK, name, ctas * WARPS, ms,
           (double)rows * rowbytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
template <int MODE, int CH, int WORK, int WARPS> static void run_write(const char *name, uint4 *arena, size_t bytes, u32 rowbytes, int sms, int ctas)
{
    const u32 quads = rowbytes / 16, rows = (u32)(bytes / rowbytes) & ~31u;
    const u32 smem = (MODE == 0 || MODE == 2) ? WARPS * 2048u : WARPS * 32u * (2 * CH + 16);
    cudaFuncSetAttribute(k_write<MODE, CH, WORK, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float ms = best_of([&] { k_write<MODE, CH, WORK, WARPS><<<ctas * sms, 32 * WARPS, smem>>>(arena, rows, quads, 1u); });
    printf("write row %5u B work %2d  %-34s %2d warps/SM %8.3f ms %7.1f GB/s  (%s)\n", rowbytes, WORK, name, ctas * WARPS, ms,
           (double)rows * rowbytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    const size_t bytes = 1ull << 30;
    uint4 *arena; u32 *out;
    cudaMalloc(&arena, bytes); cudaMemset(arena, 1, bytes); cudaMalloc(&out, 4u * 8 * sms * 512);
    for (u32 rb : {128u, 512u, 2048u}) {
        // WORK 8 ~ the scan loop's ALU work per quad is far larger (60 instr); WORK 8 = 16 dependent ops: a light consumer
        run_read<0, 64, 2, 8, 8>("lane LDG.128 (3 ahead)", arena, bytes, rb, out, sms, 2);
        run_read<3, 64, 2, 8, 8>("lane LDG.256 (2 ahead)", arena, bytes, rb, out, sms, 2);
        run_read<3, 64, 2, 40, 8>("lane LDG.256 (2 ahead)", arena, bytes, rb, out, sms, 2);
        run_read<1, 64, 4, 8, 8>("bulk 64 B x4 ring, lane mbarriers", arena, bytes, rb, out, sms, 2);
        run_read<1, 128, 2, 8, 8>("bulk 128 B x2 ring, lane mbarriers", arena, bytes, rb, out, sms, 2);
        run_read<1, 128, 3, 8, 8>("bulk 128 B x3 ring, lane mbarriers", arena, bytes, rb, out, sms, 2);
        run_read<2, 128, 2, 8, 8>("bulk 128 B x2 ring, warp mbarrier", arena, bytes, rb, out, sms, 2);
        run_read<2, 128, 3, 8, 8>("bulk 128 B x3 ring, warp mbarrier", arena, bytes, rb, out, sms, 2);
        run_read<0, 64, 2, 40, 8>("lane LDG.128 (3 ahead)", arena, bytes, rb, out, sms, 2);
        run_read<1, 128, 2, 40, 8>("bulk 128 B x2 ring, lane mbarriers", arena, bytes, rb, out, sms, 2);
        run_read<2, 128, 2, 40, 8>("bulk 128 B x2 ring, warp mbarrier", arena, bytes, rb, out, sms, 2);
        run_read<2, 64, 4, 40, 8>("bulk 64 B x4 ring, warp mbarrier", arena, bytes, rb, out, sms, 2);
        if (rb >= 512) {
            run_read<2, 256, 2, 8, 8>("bulk 256 B x2 ring, warp mbarrier", arena, bytes, rb, out, sms, 2);
            run_read<2, 256, 2, 40, 8>("bulk 256 B x2 ring, warp mbarrier", arena, bytes, rb, out, sms, 2);
        }
    }
    for (u32 rb : {512u, 2048u}) {
        run_write<0, 64, 8, 8>("stage + STG.128 (4 lanes / 64 B)", arena, bytes, rb, sms, 2);
        run_write<2, 64, 8, 8>("stage + STG.256 (2 lanes / 64 B)", arena, bytes, rb, sms, 2);
        run_write<2, 64, 40, 8>("stage + STG.256 (2 lanes / 64 B)", arena, bytes, rb, sms, 2);
        run_write<1, 64, 8, 8>("bulk store 64 B per lane", arena, bytes, rb, sms, 2);
        run_write<1, 128, 8, 8>("bulk store 128 B per lane", arena, bytes, rb, sms, 2);
        run_write<1, 256, 8, 8>("bulk store 256 B per lane", arena, bytes, rb, sms, 2);
        run_write<0, 64, 40, 8>("stage + STG.128 (4 lanes / 64 B)", arena, bytes, rb, sms, 2);
        run_write<1, 128, 40, 8>("bulk store 128 B per lane", arena, bytes, rb, sms, 2);
        run_write<1, 256, 40, 8>("bulk store 256 B per lane", arena, bytes, rb, sms, 2);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
