// ck_kernels.cuh -- the kernels of the canonicalize / uniq hot path (sm_100a, no tensor cores:
// nothing here is a contraction; the work is integer scan + hash + hash-table traffic).
//
//   k_prepare      normalise (needletail rules) + classify + pack one batch of raw records
//   k_classify     route records to a symbol lane / thread shape by alphabet and length
//   k_canon_warp   warp-per-record LMSR + canonical form + XXH3-64       (viroid / circRNA sizes)
//   k_canon_cta    CTA-per-record, same algorithm                        (plasmid / mtDNA sizes)
//   k_table_insert first-occurrence table: CAS the key, atomicMin the input index
//   k_table_first  read back first_index per record
#pragma once
#include "ck_hash.cuh"

namespace ck {

// record classes (index into the per-class work lists)
enum : int {
    CLS_W2S = 0,   // 2-bit, warp, n <= 512
    CLS_W2M = 1,   // 2-bit, warp, n <= 2048
    CLS_C2A = 2,   // 2-bit, CTA,  n <= 65536
    CLS_C2B = 3,   // 2-bit, CTA,  n <= 425984
    CLS_W4 = 4,    // 4-bit, warp, n <= 2048
    CLS_C4 = 5,    // 4-bit, CTA,  n <= 212992
    CLS_W8 = 6,    // bytes, warp, n <= 1024
    CLS_C8 = 7,    // bytes, CTA,  n <= 106496
    CLS_HUGE = 8,  // anything longer: strands staged in global scratch
    CLS_EMPTY = 9, // n == 0
    CLS_W2L = 10,  // 2-bit, warp, n <= 4096
    CLS_W2X = 11,  // 2-bit, warp, n <= 8192
    CLS_COUNT = 12
};
__host__ __device__ constexpr u32 cls_max_n(int c)
{
    return c == CLS_W2S ? 512u : c == CLS_W2M ? 2048u : c == CLS_W2L ? 4096u : c == CLS_W2X ? 8192u : c == CLS_C2A ? 65536u
         : c == CLS_C2B ? 425984u : c == CLS_W4 ? 2048u : c == CLS_C4 ? 212992u : c == CLS_W8 ? 1024u : c == CLS_C8 ? 106496u
         : c == CLS_EMPTY ? 0u : 0xffffffffu;
}
// smallest length of a class (the class below it in the same symbol lane ends one short of it)
__host__ __device__ constexpr u32 cls_min_n(int c)
{
    return c == CLS_W2M ? cls_max_n(CLS_W2S) + 1 : c == CLS_W2L ? cls_max_n(CLS_W2M) + 1 : c == CLS_W2X ? cls_max_n(CLS_W2L) + 1
         : c == CLS_C2A ? cls_max_n(CLS_W2X) + 1 : c == CLS_C2B ? cls_max_n(CLS_C2A) + 1 : c == CLS_C4 ? cls_max_n(CLS_W4) + 1
         : c == CLS_C8 ? cls_max_n(CLS_W8) + 1 : c == CLS_EMPTY ? 0u : 1u;
}

// order of the classes in the sorted index list: the lane-kernel classes first, contiguous, by length
__host__ __device__ constexpr u32 cls_rank(int c)
{
    return c == CLS_W2S ? 0u : c == CLS_W2M ? 1u : c == CLS_W2L ? 2u : c == CLS_W2X ? 3u : c == CLS_C2A ? 4u : c == CLS_C2B ? 5u
         : c == CLS_W4 ? 6u : c == CLS_C4 ? 7u : c == CLS_W8 ? 8u : c == CLS_C8 ? 9u : c == CLS_HUGE ? 10u : 11u;
}
__host__ __device__ constexpr int cls_of_rank(u32 r)
{
    return r == 0 ? CLS_W2S : r == 1 ? CLS_W2M : r == 2 ? CLS_W2L : r == 3 ? CLS_W2X : r == 4 ? CLS_C2A : r == 5 ? CLS_C2B
         : r == 6 ? CLS_W4 : r == 7 ? CLS_C4 : r == 8 ? CLS_W8 : r == 9 ? CLS_C8 : r == 10 ? CLS_HUGE : CLS_EMPTY;
}

struct CanonArgs {
    const u64 *packed2;     // 2-bit arena (ck_device.cuh): record i at u64 index p2_word(offsets[i], i)
    const u8 *bytes;        // normalised byte arena: record i at offsets[i]
    const u64 *offsets;     // n_records + 1 symbol offsets
    const u32 *lens;        // optional normalised lengths (else offsets[i+1] - offsets[i])
    const u32 *list;        // optional: the batch's record indices sorted by (class, length); this class's run starts at count[16]
    const u32 *count;       // device count of this class's entries (with list), else null; count[16] = first entry
    u32 n_direct;           // without list: records [0, n_direct)
    u8 *out;                // canonical ASCII arena (same offsets) or null
    u32 *out_start;         // or null
    u8 *out_strand;         // or null
    u64 *out_hash;          // or null
    u32 *scratch;           // tie-path scratch, `scratch_stride` u32 per warp (warp shape) / per CTA
    u64 scratch_stride;
    u32 smem_units;         // u32 units reserved per strand in (shared) staging memory
    u32 *xglobal;           // CLS_HUGE only: strands staged here, 2 * smem_units u32 per CTA
    u32 mode;               // bit0: forward strand only (lmsr / lmsr_index, lib/src/canonicalize.rs:5,41)
                            // bit1: aligned output arena: record i's bytes start at out_byte(offsets[i], i)
    u32 min_n, max_n;       // length range of this launch's class (checked when there is no list)
    u32 *retry;             // lane-per-record kernel: records it leaves to the warp / CTA kernels, appended per class at
    u32 *retry_counts;      // retry[retry_counts[16 + class] + atomicAdd(retry_counts + class, 1)]
};

// ---------------------------------------------------------------------------------------------
// One 16-byte destination-aligned chunk of a record's canonical ASCII.  t0 = record-relative index of the
// chunk's first byte (negative in the record's first chunk when the record does not start on a 16-byte
// line).  Full chunks are one 128-bit store; ragged edges use naturally aligned 1/2/4/8-byte pieces because
// the neighbouring records own the rest of the line.
__device__ __forceinline__ void store_chunk(u8 *dst, int t0, u32 n, u64 lo, u64 hi)
{
    if (t0 >= 0 && (u32)t0 + 16u <= n) {
        *reinterpret_cast<uint4 *>(dst + t0) = make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
        return;
    }
    const u32 kb = t0 < 0 ? (u32)(-t0) : 0u;
    const u32 ke = t0 < 0 ? min(16u, n + kb) : min(16u, n - (u32)t0);
    u8 *p = dst + t0;
#define CK_PIECE(k) (((k) & 8u ? hi : lo) >> (8u * ((k) & 7u)))
    u32 k = kb;
    if ((k & 1u) && k + 1 <= ke) { p[k] = (u8)CK_PIECE(k); k += 1; }
    if ((k & 2u) && k + 2 <= ke) { *reinterpret_cast<u16 *>(p + k) = (u16)CK_PIECE(k); k += 2; }
    if ((k & 4u) && k + 4 <= ke) { *reinterpret_cast<u32 *>(p + k) = (u32)CK_PIECE(k); k += 4; }
    if ((k & 8u) && k + 8 <= ke) { *reinterpret_cast<u64 *>(p + k) = hi; k += 8; }
    if (k + 8 <= ke) { *reinterpret_cast<u64 *>(p + k) = CK_PIECE(k); k += 8; }
    if (k + 4 <= ke) { *reinterpret_cast<u32 *>(p + k) = (u32)CK_PIECE(k); k += 4; }
    if (k + 2 <= ke) { *reinterpret_cast<u16 *>(p + k) = (u16)CK_PIECE(k); k += 2; }
    if (k + 1 <= ke) { p[k] = (u8)CK_PIECE(k); }
#undef CK_PIECE
}

// canonical ASCII -> global memory, generic lanes (coalesced 128-bit stores)
template <int BITS, typename G>
__device__ __forceinline__ void emit_ascii(const u32 *X, u32 n, u32 start, u8 *dst)
{
    const u32 rank = G::rank(), gs = G::size();
    const u32 a = (u32)(reinterpret_cast<uintptr_t>(dst) & 15u);
    const u32 nchunks = (n + a + 15u) >> 4;
    for (u32 c = rank; c < nchunks; c += gs) {
        const int t0 = (int)(16u * c) - (int)a;               // t0 < n always
        int ti = t0;
        if (ti < 0) { ti += (int)n; if (ti < 0) { ti %= (int)n; if (ti < 0) ti += (int)n; } }
        const u32 tt = (u32)ti;
        u32 t8 = tt + 8; if (t8 >= n) { t8 -= n; if (t8 >= n) t8 %= n; }
        const u64 lo = ascii8<BITS>(X, n, start, tt), hi = ascii8<BITS>(X, n, start, t8);
        store_chunk(dst, t0, n, lo, hi);
    }
}

// One record, both shapes.
template <int BITS, typename G>
__device__ __forceinline__ void do_record(const CanonArgs &a, u32 rec, u32 *Xf, u32 *Xr, u32 *scr, u64 *hbuf, u32 *red)
{
    const u64 off = a.offsets[rec];
    const u32 n = a.lens ? a.lens[rec] : (u32)(a.offsets[rec + 1] - off);
    if (a.list == nullptr && (n < a.min_n || n > a.max_n)) return;   // direct mode: k_classify reported it (uniform)
    RecordIn in;
    in.packed2 = a.packed2 ? a.packed2 + p2_word(off, rec) : nullptr;
    in.bytes = a.bytes ? a.bytes + off : nullptr;
    in.n = n;
    stage_record<BITS, G>(in, Xf, Xr);
    RecordOut o = canonical_start<BITS, G>(Xf, Xr, n, scr, red, (a.mode & 1u) != 0);
    const u32 *X = o.strand ? Xr : Xf;
    if (a.out) emit_ascii<BITS, G>(X, n, o.start, a.out + ((a.mode & 2u) ? out_byte(off, rec) : off));
    u64 h = 0;
    if (a.out_hash) h = xxh3_canonical<BITS, G>(X, n, o.start, hbuf, red);
    if (G::rank() == 0) {
        // strand 1 start is reported in forward coordinates: canonical[j] = comp(s[(start - j) mod n])
        if (a.out_start) a.out_start[rec] = o.strand ? (n - 1 - o.start) : o.start;
        if (a.out_strand) a.out_strand[rec] = (u8)o.strand;
        if (a.out_hash) a.out_hash[rec] = h;
    }
    G::sync();      // staging memory is reused by the next record
}

template <int BITS>
__global__ void __launch_bounds__(256) k_canon_warp(CanonArgs a)
{
    extern __shared__ u32 smem[];
    typedef Grp<false> G;
    const u32 wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    u32 *Xf = smem + (size_t)wid * 2 * a.smem_units, *Xr = Xf + a.smem_units;
    const u32 gw = blockIdx.x * wpb + wid, nw = gridDim.x * wpb;
    u32 *scr = a.scratch + (size_t)gw * a.scratch_stride;
    const u32 count = a.list ? *a.count : a.n_direct;
    if (a.list) a.list += a.count[16];
    for (u32 e = gw; e < count; e += nw) {
        const u32 rec = a.list ? a.list[e] : e;
        do_record<BITS, G>(a, rec, Xf, Xr, scr, nullptr, nullptr);
    }
}

template <int BITS, bool XGLOBAL>
__global__ void __launch_bounds__(1024) k_canon_cta(CanonArgs a)
{
    extern __shared__ u32 smem[];
    __shared__ u32 red[40];
    __shared__ u64 hbuf[32 * 8];
    typedef Grp<true> G;
    u32 *Xf = XGLOBAL ? a.xglobal + (size_t)blockIdx.x * 2 * a.smem_units : smem;
    u32 *Xr = Xf + a.smem_units;
    u32 *scr = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
    const u32 count = a.list ? *a.count : a.n_direct;
    if (a.list) a.list += a.count[16];
    for (u32 e = blockIdx.x; e < count; e += gridDim.x) {
        const u32 rec = a.list ? a.list[e] : e;
        do_record<BITS, G>(a, rec, Xf, Xr, scr, hbuf, red);
    }
}

// n == 0 records: canonical form is empty, strand follows the `else` arm, hash is XXH3 of "".
__global__ void k_canon_empty(CanonArgs a)
{
    const u32 count = a.list ? *a.count : a.n_direct;
    if (a.list) a.list += a.count[16];
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const u32 rec = a.list ? a.list[e] : e;
        if (a.list == nullptr && a.offsets[rec + 1] != a.offsets[rec]) continue;
        if (a.out_start) a.out_start[rec] = 0;
        if (a.out_strand) a.out_strand[rec] = 1;
        if (a.out_hash) a.out_hash[rec] = xxh64_avalanche(sec64(56) ^ sec64(64));
    }
}

// ---------------------------------------------------------------------------------------------
// k_prepare: one warp per raw record.
//   flags bit0 (normalise): needletail::sequence::normalize(seq, false) -- src/canonicalize.rs:24-27
//                           (whitespace dropped, acgt -> upper, t/u/U -> T, ./~ -> -, anything else -> N)
//   otherwise             : library semantics, bytes are taken as they are (lib/src/canonicalize.rs:54)
// Writes lens[i], lane[i] (2, 4 or 8 bits per symbol) and the record in that lane's format:
//   lane 2 -> packed2 at word p2_word(offsets[i], i);   lanes 4/8 -> normalised bytes at offsets[i].
struct PrepareArgs {
    const u8 *raw; const u64 *offsets; u32 n_records; u32 flags;
    u64 *packed2; u8 *bytes; u32 *lens; u8 *lane;
};
__device__ __forceinline__ u32 code2_of(u32 b) { return ((b >> 1) & 3u) ^ ((b >> 2) & 1u); }   // A,C,G,T -> 0..3

__global__ void __launch_bounds__(256) k_prepare(PrepareArgs a)
{
    const u32 lane = lane_id();
    const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const bool norm = a.flags & 1u;
    for (u32 rec = gw; rec < a.n_records; rec += nw) {
        const u64 off = a.offsets[rec];
        const u32 rawlen = (u32)(a.offsets[rec + 1] - off);
        const u8 *src = a.raw + off;
        // pass 1: normalised length and alphabet class
        u32 cnt = 0, cls = 0;
        for (u32 base = 0; base < rawlen; base += 32) {
            u32 t = base + lane;
            u32 m = 0;
            if (t < rawlen) { u32 b = __ldg(src + t); m = norm ? (u32)c_tab.norm[b] : (b | 0x100u); }
            if (m) {
                u32 b = m & 0xffu;
                bool acgt = (b == 'A') | (b == 'C') | (b == 'G') | (b == 'T');
                cls |= acgt ? 0u : (c_tab.code4[b] != 0xffu ? 1u : 3u);
            }
            cnt += __popc(__ballot_sync(CK_FULL, m != 0));
        }
        cls = __reduce_or_sync(CK_FULL, cls);
        const u32 lanebits = cls == 0 ? 2u : (cls == 1 ? 4u : 8u);
        if (lane == 0) { a.lens[rec] = cnt; a.lane[rec] = (u8)lanebits; }
        // pass 2: write in the lane's format
        u64 *dstw = a.packed2 + p2_word(off, rec);
        u8 *dstb = a.bytes + off;
        u32 done = 0;                 // normalised symbols written so far
        u32 acc_hi = 0, acc_lo = 0;   // word under construction (uniform across the warp)
        for (u32 base = 0; base < rawlen; base += 32) {
            u32 t = base + lane;
            u32 m = 0;
            if (t < rawlen) { u32 b = __ldg(src + t); m = norm ? (u32)c_tab.norm[b] : (b | 0x100u); }
            const u32 keep = __ballot_sync(CK_FULL, m != 0);
            const u32 idx = done + __popc(keep & ((1u << lane) - 1u));
            if (lanebits != 2) {
                if (m) dstb[idx] = (u8)m;
            } else {
                // symbols of this chunk fall into word A = done >> 5 and possibly A + 1
                const u32 A = done >> 5;
                u32 c_hi0 = 0, c_lo0 = 0, c_hi1 = 0, c_lo1 = 0;
                if (m) {
                    u32 code = code2_of(m & 0xffu), sh = 62 - 2 * (idx & 31u);
                    u32 hi = sh >= 32 ? code << (sh - 32) : 0u, lo = sh < 32 ? code << sh : 0u;
                    if ((idx >> 5) == A) { c_hi0 = hi; c_lo0 = lo; } else { c_hi1 = hi; c_lo1 = lo; }
                }
                acc_hi |= __reduce_or_sync(CK_FULL, c_hi0);
                acc_lo |= __reduce_or_sync(CK_FULL, c_lo0);
                const u32 ndone = done + __popc(keep);
                if ((ndone >> 5) != A) {          // word A is complete
                    if (lane == 0) dstw[A] = ((u64)acc_lo << 32) | acc_hi;   // unit of bases 0-15 at the lower address
                    acc_hi = __reduce_or_sync(CK_FULL, c_hi1);
                    acc_lo = __reduce_or_sync(CK_FULL, c_lo1);
                }
            }
            done += __popc(keep);
        }
        if (lanebits == 2 && (done & 31u) && lane == 0) dstw[done >> 5] = ((u64)acc_lo << 32) | acc_hi;
    }
}

// ---------------------------------------------------------------------------------------------
// k_classify: class of every record.  With a class promise (direct_mask) records outside it are counted as
// "left unprocessed"; otherwise every record gets a sort key (class << 8 | length bin) and the classes are counted:
// a radix sort of (key, index) then yields ONE index list in which each class is a contiguous run ordered by length,
// so the lanes of a lane-per-record warp walk records of (nearly) equal length.
struct ClassifyArgs {
    const u64 *offsets; const u32 *lens; const u8 *lane;   // lane == null: every record is 2-bit
    u32 n_records;
    u32 *keys, *vals;  // n_records sort keys / record indices (list mode)
    u32 *counts;       // 16 class counters (zeroed by the caller), followed by 16 run starts (k_list_starts)
    u32 direct_mask;   // != 0: no keys are written (the classes of the mask run over all records directly);
                       // records of any other class are counted in counts[CLS_HUGE] = "left unprocessed"
};
__global__ void __launch_bounds__(256) k_classify(ClassifyArgs a)
{
    __shared__ u32 cnt[16];                                    // per-CTA class counts: one global atomic per class and CTA
    if (threadIdx.x < 16) cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 nthreads = gridDim.x * blockDim.x;
    const u32 rounds = (a.n_records + nthreads - 1) / nthreads;
    for (u32 r = 0; r < rounds; r++) {
        const u32 i = r * nthreads + blockIdx.x * blockDim.x + threadIdx.x;
        int cls = -1;
        u32 n = 0;
        if (i < a.n_records) {
            n = a.lens ? a.lens[i] : (u32)(a.offsets[i + 1] - a.offsets[i]);
            const u32 bits = a.lane ? a.lane[i] : 2u;
            if (n == 0) cls = CLS_EMPTY;
            else if (bits == 2) cls = n <= cls_max_n(CLS_W2S) ? CLS_W2S : n <= cls_max_n(CLS_W2M) ? CLS_W2M
                                    : n <= cls_max_n(CLS_W2L) ? CLS_W2L : n <= cls_max_n(CLS_W2X) ? CLS_W2X
                                    : n <= cls_max_n(CLS_C2A) ? CLS_C2A : n <= cls_max_n(CLS_C2B) ? CLS_C2B : CLS_HUGE;
            else if (bits == 4) cls = n <= cls_max_n(CLS_W4) ? CLS_W4 : n <= cls_max_n(CLS_C4) ? CLS_C4 : CLS_HUGE;
            else cls = n <= cls_max_n(CLS_W8) ? CLS_W8 : n <= cls_max_n(CLS_C8) ? CLS_C8 : CLS_HUGE;
        }
        if (a.direct_mask) {
            const u32 m = __ballot_sync(CK_FULL, cls >= 0 && !((a.direct_mask >> cls) & 1u));
            if (m && lane_id() == (u32)(__ffs(m) - 1)) atomicAdd(cnt + CLS_HUGE, __popc(m));
            continue;
        }
        if (cls >= 0) {
            // key = (rank of the class in list order) : (length on a 1/4-octave log scale: lanes of a warp differ by < 19 % in
            // length, and a bin is dense enough that its consecutive members sit close together in the arenas; measured on
            // config 2: 1/32 octave 9.7 ms, 1/8 octave 8.6 ms, 1/4 octave 8.4 ms)
            const u32 msb = 31u - __clz(n | 1u);
            const u32 bin = (msb << 5) | ((((n << (31u - msb)) >> 29) & 3u) << 3);
            a.keys[i] = (cls_rank(cls) << 10) | bin; a.vals[i] = i;
        }
        // warp-aggregated counting: one shared-memory atomic per (warp, class present)
        u32 todo = __ballot_sync(CK_FULL, cls >= 0);
        while (todo) {
            const int c = __shfl_sync(CK_FULL, cls, __ffs(todo) - 1);
            const u32 m = __ballot_sync(CK_FULL, cls == c);
            if (lane_id() == (u32)(__ffs(m) - 1)) atomicAdd(cnt + c, __popc(m));
            todo &= ~m;
        }
    }
    __syncthreads();
    if (threadIdx.x < 16 && cnt[threadIdx.x]) atomicAdd(a.counts + threadIdx.x, cnt[threadIdx.x]);
}
// run starts of the sorted index list (classes in cls_rank order).  Layout of the 64 counters:
//   [0, 12)  records per class        [12]  records of the lane-kernel classes together (one launch)
//   [16, 28) first entry per class    [28]  first entry of the lane-kernel classes
//   [32, 44) retry entries per class  [48, 60)  first retry entry per class (= first entry of the class)
__global__ void k_list_starts(u32 *counts)
{
    if (threadIdx.x == 0) {
        u32 s = 0;
        for (u32 r = 0; r < (u32)CLS_COUNT; r++) {
            const int c = cls_of_rank(r);
            counts[16 + c] = s; counts[48 + c] = s;
            s += counts[c];
        }
        counts[12] = counts[CLS_W2S] + counts[CLS_W2M] + counts[CLS_W2L] + counts[CLS_W2X];
        counts[28] = counts[16 + CLS_W2S];
    }
}

// k_extend_packed2: append the circular extension to every packed record (units jn .. jn + 4, jn = n >> 4; the
// partial unit jn keeps its real bases).  One thread per record; runs after the packer (k_prepare / k_synth_packed2).
__global__ void __launch_bounds__(256) k_extend_packed2(u64 *packed2, const u64 *offsets, const u32 *lens, const u8 *lane, u32 n_records)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_records) return;
    if (lane && lane[i] != 2) return;
    const u64 off = offsets[i];
    const u32 n = lens ? lens[i] : (u32)(offsets[i + 1] - off);
    if (n == 0) return;
    u32 *U = reinterpret_cast<u32 *>(packed2 + p2_word(off, i));
    const u32 jn = n >> 4, rem = n & 15u;
    u32 e[5];
    if (n >= 80) {
        const u32 u0 = U[0], u1 = U[1], u2 = U[2], u3 = U[3], u4 = U[4], uj = U[jn];
        const u32 sh = 32u - 2u * rem;
        e[0] = (uj & ~(0xffffffffu >> (2 * rem))) | (u0 >> (2 * rem));
        e[1] = __funnelshift_lc(u1, u0, sh); e[2] = __funnelshift_lc(u2, u1, sh); e[3] = __funnelshift_lc(u3, u2, sh);
        e[4] = __funnelshift_lc(u4, u3, sh);
    } else {
        for (u32 k = 0; k < 5; k++) {
            u32 v = 0;
            for (u32 b = 0; b < 16; b++) {
                const u32 t = (16 * (jn + k) + b) % n;
                v = (v << 2) | ((U[t >> 4] >> (30 - 2 * (t & 15))) & 3u);
            }
            e[k] = v;
        }
    }
    for (u32 k = 0; k < 5; k++) U[jn + k] = e[k];
}

// ---------------------------------------------------------------------------------------------
// First-occurrence table (replaces HashMap<u64, String, NoHash>, src/uniq.rs:27,47-48,66).
// slot = {key, first}; key == EMPTY_KEY marks a free slot, and the one real key equal to EMPTY_KEY is
// kept in a side slot, so every u64 stays a legal XXH3 value.  Keeping the MIN input index per key
// reproduces "first record seen wins" of the serial consumer for any insertion order.
struct TableSlot { u64 key; u64 first; };
#define CK_EMPTY_KEY 0xffffffffffffffffULL
struct TableArgs {
    TableSlot *slots; u64 mask;          // capacity - 1 (power of two)
    u64 *side_first;                     // first index of key == EMPTY_KEY
    const u64 *hash; const u64 *index;   // index == null: base_index + i
    u32 stride;                          // elements between consecutive records in hash / index (2: interleaved pairs)
    u64 base_index; u32 n;
    u64 *slot_of;                        // out: slot found per record (for k_table_first)
    u64 *first_out;                      // out (k_table_first)
    u32 *overflow;                       // set if the table is full
};
__device__ __forceinline__ u64 table_home(u64 key, u64 mask) { return (key ^ (key >> 29)) & mask; }

__global__ void __launch_bounds__(256) k_table_insert(TableArgs a)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const u64 key = a.hash[(size_t)i * a.stride];
    const u64 idx = a.index ? a.index[(size_t)i * a.stride] : a.base_index + i;
    if (key == CK_EMPTY_KEY) { atomicMin(a.side_first, idx); a.slot_of[i] = ~0ULL; return; }
    u64 s = table_home(key, a.mask);
    for (u64 probes = 0; probes <= a.mask; probes++) {
        u64 prev = atomicCAS(&a.slots[s].key, CK_EMPTY_KEY, key);
        if (prev == CK_EMPTY_KEY || prev == key) {
            atomicMin(&a.slots[s].first, idx);
            a.slot_of[i] = s;
            return;
        }
        s = (s + 1) & a.mask;
    }
    *a.overflow = 1; a.slot_of[i] = ~0ULL - 1;
}
__global__ void __launch_bounds__(256) k_table_first(TableArgs a)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const u64 s = a.slot_of[i];
    a.first_out[i] = s == ~0ULL ? *a.side_first : (s == ~0ULL - 1 ? ~0ULL : a.slots[s].first);
}
__global__ void k_table_clear(TableSlot *slots, u64 n, u64 *side_first)
{
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        slots[i].key = CK_EMPTY_KEY; slots[i].first = ~0ULL;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *side_first = ~0ULL;
}

// ---------------------------------------------------------------------------------------------
// Hash-range partition for multi-GPU uniq: owner = floor(hash * world / 2^64).  Two passes over
// the local hashes: counts per owner, then a scatter of (hash, global index) into per-owner runs of the send
// buffers (order inside a run is free: the owner keeps the minimum index per key).  pos[i] = where record i went, so
// the answer that comes back in send-buffer order can be gathered back into input order.
struct OwnerArgs {
    const u64 *hash; u32 n; u32 world; u64 base_index;
    u32 *counts;       // [world] records per owner, then [world] scatter cursors (both zeroed by the caller)
    u64 *send_pairs; u32 *pos;   // send buffer: (hash, global index) pairs, 16 bytes per record
};
// floor(hash * world / 2^64) on the top 32 bits.  Written with __umulhi: nvcc 12.9 turned the 64-bit form
// (((h >> 32) * world) >> 32), used as a shared-memory index, into a low multiply (every record went to owner 0).
__device__ __forceinline__ u32 owner_of_hash(u64 h, u32 world) { return __umulhi((u32)(h >> 32), world); }
__global__ void __launch_bounds__(256) k_owner_count(OwnerArgs a)
{
    __shared__ u32 cnt[32];
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
        atomicAdd(cnt + owner_of_hash(a.hash[i], a.world), 1u);
    __syncthreads();
    if (threadIdx.x < a.world && cnt[threadIdx.x]) atomicAdd(a.counts + threadIdx.x, cnt[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_owner_scatter(OwnerArgs a)
{
    __shared__ u32 cnt[32], basep[32];
    const u32 per = (a.n + gridDim.x - 1) / gridDim.x;         // this CTA's contiguous slice
    const u32 lo = min(a.n, blockIdx.x * per), hi = min(a.n, lo + per);
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (u32 i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(cnt + owner_of_hash(a.hash[i], a.world), 1u);
    __syncthreads();
    if (threadIdx.x < a.world) {
        u32 start = 0;                                         // run start of this owner = counts of the owners below it
        for (u32 o = 0; o < threadIdx.x; o++) start += a.counts[o];
        basep[threadIdx.x] = start + atomicAdd(a.counts + a.world + threadIdx.x, cnt[threadIdx.x]);
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    for (u32 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const u64 h = a.hash[i];
        const u32 o = owner_of_hash(h, a.world);
        const u32 p = basep[o] + atomicAdd(cnt + o, 1u);
        reinterpret_cast<ulonglong2 *>(a.send_pairs)[p] = make_ulonglong2(h, a.base_index + i);
        a.pos[i] = p;
    }
}

}  // namespace ck
