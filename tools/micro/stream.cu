// Microbenchmark: cost of lane-private streaming reads (every lane walks its own row of a global arena) against
// warp-coalesced reads, on sm_100a.  Answers "how many LSU cycles does a fully divergent 16-byte load instruction cost".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream stream.cu && ./stream
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32; typedef unsigned long long u64;

__device__ __forceinline__ void cp16(u32 dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MODE 0: lane-private LDG.128   1: lane-private cp.async 16 B through an 8-deep ring   2: lane-private LDG.64   3: coalesced LDG.128
template <int MODE> __global__ void __launch_bounds__(256) k(const uint4 *arena, u32 rows, u32 quads, u32 *out)
{
    __shared__ uint4 ring[8][256];
    const u32 lane = threadIdx.x & 31u, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    u32 acc = 0;
    for (u32 b = gw * 32u; b < rows; b += nw * 32u) {
        if (MODE == 0) {
            const uint4 *p = arena + (size_t)(b + lane) * quads;
#pragma unroll 4
            for (u32 i = 0; i < quads; i++) { uint4 v = __ldg(p + i); acc ^= v.x + v.y + v.z + v.w; }
        } else if (MODE == 2) {
            const uint2 *p = reinterpret_cast<const uint2 *>(arena + (size_t)(b + lane) * quads);
#pragma unroll 4
            for (u32 i = 0; i < 2 * quads; i++) { uint2 v = __ldg(p + i); acc ^= v.x + v.y; }
        } else if (MODE == 3) {
            const uint4 *p = arena + (size_t)b * quads;
#pragma unroll 4
            for (u32 i = lane; i < 32 * quads; i += 32) { uint4 v = __ldg(p + i); acc ^= v.x + v.y + v.z + v.w; }
        } else {
            const uint4 *p = arena + (size_t)(b + lane) * quads;
            const u32 base = (u32)__cvta_generic_to_shared(&ring[0][threadIdx.x]);
            for (u32 i = 0; i < 7 && i < quads; i++) { cp16(base + i * 4096u, p + i); cp_commit(); }
            for (u32 i = 0; i < quads; i++) {
                if (i + 7 < quads) cp16(base + ((i + 7) & 7u) * 4096u, p + i + 7);
                cp_commit();
                cp_wait<7>();
                uint4 v = ring[i & 7u][threadIdx.x];
                acc ^= v.x + v.y + v.z + v.w;
            }
            cp_wait<0>();
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    const size_t bytes = 1ull << 30;
    uint4 *arena; u32 *out;
    cudaMalloc(&arena, bytes); cudaMemset(arena, 1, bytes); cudaMalloc(&out, 4u * 3 * sms * 256);
    const char *names[] = {"lane LDG.128", "lane cp.async16 ring8", "lane LDG.64", "coalesced LDG.128"};
    const u32 rowbytes[] = {96, 1024, 4096};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (u32 rb : rowbytes) {
        const u32 quads = rb / 16, rows = (u32)(bytes / rb) & ~31u;
        for (int m = 0; m < 4; m++) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (m == 0) k<0><<<3 * sms, 256>>>(arena, rows, quads, out);
                if (m == 1) k<1><<<3 * sms, 256>>>(arena, rows, quads, out);
                if (m == 2) k<2><<<3 * sms, 256>>>(arena, rows, quads, out);
                if (m == 3) k<3><<<3 * sms, 256>>>(arena, rows, quads, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double instr = (double)rows / 32.0 * quads * (m == 2 ? 2 : 1);       // warp-level load instructions
            const double cyc = best * 1e-3 * pr.clockRate * 1e3;                         // SM cycles elapsed
            printf("row %5u B  %-24s %8.3f ms  %7.1f GB/s  %6.2f SM-cycles per warp load instr\n", rb, names[m], best,
                   bytes / best / 1e6, cyc * sms / instr);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
