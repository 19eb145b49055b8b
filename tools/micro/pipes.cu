// Throughput microbenchmark of the integer ops the canonicalize kernels are built from (sm_100a).
// Prints warp-instructions per cycle per SM for each op, 16 warps per SMSP, ILP 8.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32; typedef unsigned long long u64;
#define ITERS 2048
#define ILP 8
enum { OP_SHF, OP_MIN3, OP_MIN3_16, OP_MIN2_16, OP_PRMT, OP_LOP3, OP_IMAD, OP_IMADW, OP_IADD3, OP_MIX_SHF_IMAD, OP_LDS, OP_SHFL,
       OP_REDUX, OP_BREV, OP_POPC, OP_MIX_MIN_IMAD, OP_LDS64, OP_LDS128, OP_VOTE, OP_SEL, OP_ISETP_SEL, OP_IMADHI, OP_FUNNEL_FMA, OP_MIX_SHF_FUNNEL_FMA, OP_MIX3_SHF_FUNNEL_FMA, OP_LDS_OWNBANK, OP_ASCII_ALU, OP_ASCII_LUT, OP_ASCII_LUT8, OP_COUNT };
const char *names[] = {"SHF(funnel)", "VIMNMX3.U32", "VIMNMX3.U16x2", "VIMNMX.U16x2", "PRMT", "LOP3", "IMAD", "IMAD.WIDE", "IADD3",
                       "SHF+IMAD 1:1", "LDS.32", "SHFL.IDX", "REDUX.MIN", "BREV", "POPC", "VIMNMX3+IMAD 1:1", "LDS.64", "LDS.128", "VOTE.BALLOT",
                       "SEL", "ISETP+SEL", "IMAD.HI", "funnel = IMAD.HI+IMAD (per funnel)", "SHF : FMA-funnel 1:1 (per funnel)",
                       "SHF : FMA-funnel 3:1 (per funnel)", "LDS.32 own-bank table", "ascii16 ALU (13 ops) per 16 chars", "ascii16 LUT x32 copies per 16 chars",
                       "ascii16 LUT x8 copies per 16 chars"};
template <int OP> __global__ void __launch_bounds__(512) k(u32 *out, u32 seed, long long *cyc)
{
    __shared__ u32 sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    u32 v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = seed * (threadIdx.x + 1 + i);
    u32 a = seed | 1u, b = seed * 3u + 1u;
    u64 w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) w[i] = v[i];
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == OP_SHF) v[i] = __funnelshift_l(v[i], a, (b + i) & 31);
            if (OP == OP_MIN3) v[i] = min(min(v[i], a + i), b);
            if (OP == OP_MIN3_16) v[i] = __vimin3_u16x2(v[i], a + i, b);
            if (OP == OP_MIN2_16) v[i] = __vminu2(v[i], a + i);
            if (OP == OP_PRMT) v[i] = __byte_perm(v[i], a, b + i);
            if (OP == OP_LOP3) v[i] = (v[i] & a) ^ (b + i);
            if (OP == OP_IMAD) v[i] = v[i] * a + b;
            if (OP == OP_IMADW) w[i] = (u64)(u32)w[i] * (u64)a + w[i];
            if (OP == OP_IADD3) v[i] = v[i] + a + b;
            if (OP == OP_MIX_SHF_IMAD) { if (i & 1) v[i] = __funnelshift_l(v[i], a, (b + i) & 31); else v[i] = v[i] * a + b; }
            if (OP == OP_MIX_MIN_IMAD) { if (i & 1) v[i] = min(min(v[i], a + i), b); else v[i] = v[i] * a + b; }
            if (OP == OP_LDS) v[i] = sm[(v[i] + threadIdx.x) & 4095];
            if (OP == OP_LDS64) { uint2 t = *reinterpret_cast<uint2 *>(&sm[((v[i] + threadIdx.x) * 2) & 4094]); v[i] = t.x + t.y; }
            if (OP == OP_LDS128) { uint4 t = *reinterpret_cast<uint4 *>(&sm[((v[i] + threadIdx.x) * 4) & 4092]); v[i] = t.x + t.w; }
            if (OP == OP_SHFL) v[i] = __shfl_sync(0xffffffffu, v[i], (v[i] + i) & 31);
            if (OP == OP_REDUX) v[i] = __reduce_min_sync(0xffffffffu, v[i] + i);
            if (OP == OP_BREV) v[i] = __brev(v[i]) + 1;
            if (OP == OP_POPC) v[i] = __popc(v[i]) + a;
            if (OP == OP_VOTE) v[i] = __ballot_sync(0xffffffffu, v[i] > a) + i;
            if (OP == OP_SEL) v[i] = (b & (1u << i)) ? v[i] : a;
            if (OP == OP_ISETP_SEL) v[i] = (v[i] > a) ? v[i] : b + i;
            if (OP == OP_IMADHI) v[i] = __umulhi(v[i], a) + b;
            if (OP == OP_FUNNEL_FMA) v[i] = v[i] * 64u + __umulhi(v[(i + 1) % ILP], 64u);
            if (OP == OP_MIX_SHF_FUNNEL_FMA) { if (i & 1) v[i] = __funnelshift_l(v[(i + 1) % ILP], v[i], 6); else v[i] = v[i] * 64u + __umulhi(v[(i + 1) % ILP], 64u); }
            if (OP == OP_MIX3_SHF_FUNNEL_FMA) { if (i & 3) v[i] = __funnelshift_l(v[(i + 1) % ILP], v[i], 6); else v[i] = v[i] * 64u + __umulhi(v[(i + 1) % ILP], 64u); }
            if (OP == OP_LDS_OWNBANK) v[i] = sm[((v[i] & 255u) << 5) + (threadIdx.x & 31u)] + i;
            if (OP == OP_ASCII_ALU) {
                const u32 T = 0x54474341u, x = v[i];
                const u32 E = x & 0x33333333u, O = (x >> 2) & 0x33333333u;
                const u32 pe_lo = __byte_perm(T, 0, E), pe_hi = __byte_perm(T, 0, E >> 16);
                const u32 po_lo = __byte_perm(T, 0, O), po_hi = __byte_perm(T, 0, O >> 16);
                v[i] = __byte_perm(pe_hi, po_hi, 0x2637) + __byte_perm(pe_hi, po_hi, 0x0415) + __byte_perm(pe_lo, po_lo, 0x2637) + __byte_perm(pe_lo, po_lo, 0x0415);
            }
            if (OP == OP_ASCII_LUT || OP == OP_ASCII_LUT8) {
                const u32 x = v[i];
                const u32 sh = OP == OP_ASCII_LUT ? 32u : 8u, lb = OP == OP_ASCII_LUT ? (threadIdx.x & 31u) : (threadIdx.x & 7u);
                const u32 i0 = x >> 24, i1 = __byte_perm(x, 0, 0x4442), i2 = __byte_perm(x, 0, 0x4441), i3 = x & 255u;
                v[i] = sm[i0 * sh + lb] + sm[i1 * sh + lb] + sm[i2 * sh + lb] + sm[i3 * sh + lb] + i;
            }
        }
    }
    long long t1 = clock64();
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r += v[i] + (u32)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(u32 *out, long long *cyc, int sms)
{
    k<OP><<<sms * 4, 512>>>(out, 12345u, cyc);   // 4 x 16 warps = 64 warps/SM = 16 per SMSP
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<sms * 4, 512>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double winst = 64.0 * ITERS * ILP;    // per SM
    printf("%-18s %8.3f warp-inst/cyc/SM (%.3f per SMSP)   [%lld cyc, %.3f ms]%s\n", names[OP], winst / h[0], winst / h[0] / 4, h[0], ms,
           OP == OP_MIX_SHF_IMAD || OP == OP_MIX_MIN_IMAD ? "  (counts both ops)" : "");
}
template <int OP> struct R { static void go(u32 *o, long long *c, int s) { run<OP>(o, c, s); R<OP + 1>::go(o, c, s); } };
template <> struct R<OP_COUNT> { static void go(u32 *, long long *, int) {} };
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    u32 *out; long long *cyc;
    cudaMalloc(&out, p.multiProcessorCount * 4 * 512 * 4); cudaMalloc(&cyc, p.multiProcessorCount * 4 * 8);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    R<0>::go(out, cyc, p.multiProcessorCount);
    return 0;
}
