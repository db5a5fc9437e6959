// Hand-written sm_100a kernels of libfbgpu.  Pure integer / bitwise work, HBM-bound: no tensor cores.
//
//   eval_kernel        one CTA per (shard, container slot): runs a compiled bitmap-call program with its
//                      operand stack held as 8 KiB bitmaps in shared memory (replaces executeBitmapCallShard +
//                      Row/roaring set algebra, executor.go:1782, row.go:242-353, roaring.go:736-1623).
//   pair_count_kernel  one warp per container pair: fused Intersect+Count for Count(Intersect(Row,Row))
//                      (replaces roaring.intersectionCount's 9 type-pair kernels, roaring.go:4477-4614).
//   row_count_kernel   one warp per (shard,row): per-row |row ∩ filter| (doTopK executor.go:2705, fragment.top).
//   groupby_kernel     one CTA per (shard, slot): column-keyed join of two fields' rows (groupByIterator :8617).
//   canon_*            canonical (optimize()) container emission for Row results (roaring.go:3412-3461).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fbgpu_types.h"

namespace fbgpu {

#ifndef FBGPU_EVAL_THREADS
#define FBGPU_EVAL_THREADS 256
#endif
#ifndef FBGPU_EVAL_MIN_BLOCKS
#define FBGPU_EVAL_MIN_BLOCKS 7
#endif
#ifndef FBGPU_BATCH_UNROLL
#define FBGPU_BATCH_UNROLL 4
#endif
constexpr int kEvalThreads = FBGPU_EVAL_THREADS;          // 256 or 512
constexpr int kEvalU4PerThread = 512 / kEvalThreads;       // uint4 per thread of an 8 KiB bitmap
constexpr int kEvalW64PerThread = 1024 / kEvalThreads;     // consecutive u64 words per thread in scans
constexpr int kResolveChunk = 256;

__device__ __forceinline__ uint4 ldg_nc(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int popc4(uint4 v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ uint4 xor4(uint4 a, uint4 b) { return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
__device__ __forceinline__ uint4 andn4(uint4 a, uint4 b) { return make_uint4(a.x & ~b.x, a.y & ~b.y, a.z & ~b.z, a.w & ~b.w); }

// Locate the container (fv, shard, row, slot).  5 dependent loads; see fbgpu_types.h.
__device__ __forceinline__ Resolved resolve(const StoreRef& st, uint32_t fv, uint64_t shard, uint64_t row, int slot) {
    Resolved r; r.ptr = nullptr; r.card = 0; r.typ = 0; r.cnt = 0;
    if (fv >= st.n_views) return r;
    ViewTab v = st.views[fv];
    if (shard >= v.n_shards) return r;
    int f = st.shardmap[v.shard_off + shard];
    if (f < 0) return r;
    FragHdr h = st.frags[f];
    uint32_t idx;
    if (h.contiguous) {
        if (row < h.row0 || row - h.row0 >= h.n_rows) return r;
        idx = (uint32_t)(row - h.row0);
    } else {
        uint32_t lo = 0, hi = h.n_rows;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (st.rows[h.row_off + m].row < row) lo = m + 1; else hi = m; }
        if (lo >= h.n_rows) return r;
        idx = lo;
    }
    RowEnt e = st.rows[h.row_off + idx];
    if (e.row != row || !((e.mask >> slot) & 1)) return r;
    ContDesc d = st.descs[e.first_desc + __popc(e.mask & ((1u << slot) - 1u))];
    r.ptr = st.payload + (size_t)d.off16 * 16; r.card = d.card; r.typ = d.typ; r.cnt = d.cnt;
    return r;
}

// ------------------------------------------------------------------------------------------------
// CTA-level helpers on 8 KiB shared-memory bitmaps (uint4[512]); thread t owns uint4 t and t+256.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bm_zero(uint4* d) {
#pragma unroll
    for (int h = 0; h < kEvalU4PerThread; h++) d[threadIdx.x + h * kEvalThreads] = make_uint4(0, 0, 0, 0);
}
// scatter an array container into a bitmap with MODE 0: |=  1: &= ~  2: ^=
template <int MODE>
__device__ __forceinline__ void bm_scatter(uint32_t* bm, const uint16_t* arr, uint32_t n) {
    const uint32_t* a32 = reinterpret_cast<const uint32_t*>(arr);
    uint32_t n2 = (n + 1) >> 1;
    for (uint32_t i = threadIdx.x; i < n2; i += kEvalThreads) {
        uint32_t v = __ldg(a32 + i);
        uint32_t lo = v & 0xffffu, hi = v >> 16;
        if (MODE == 0) atomicOr(&bm[lo >> 5], 1u << (lo & 31)); else if (MODE == 1) atomicAnd(&bm[lo >> 5], ~(1u << (lo & 31))); else atomicXor(&bm[lo >> 5], 1u << (lo & 31));
        if (2 * i + 1 < n) {
            if (MODE == 0) atomicOr(&bm[hi >> 5], 1u << (hi & 31)); else if (MODE == 1) atomicAnd(&bm[hi >> 5], ~(1u << (hi & 31))); else atomicXor(&bm[hi >> 5], 1u << (hi & 31));
        }
    }
}
// for each array element present in `src`, set it in `dst`
__device__ __forceinline__ void bm_filter_scatter(uint32_t* dst, const uint32_t* src, const uint16_t* arr, uint32_t n) {
    const uint32_t* a32 = reinterpret_cast<const uint32_t*>(arr);
    uint32_t n2 = (n + 1) >> 1;
    for (uint32_t i = threadIdx.x; i < n2; i += kEvalThreads) {
        uint32_t v = __ldg(a32 + i);
        uint32_t lo = v & 0xffffu, hi = v >> 16;
        if ((src[lo >> 5] >> (lo & 31)) & 1) atomicOr(&dst[lo >> 5], 1u << (lo & 31));
        if (2 * i + 1 < n && ((src[hi >> 5] >> (hi & 31)) & 1)) atomicOr(&dst[hi >> 5], 1u << (hi & 31));
    }
}
// Expand a run container into `dst` (overwrites).  Delta bitmap (toggle at start and last+1) followed by a
// CTA-wide prefix-XOR scan: O(runs + 1024 words), independent of run lengths (runToBitmap roaring.go:3792).
__device__ __noinline__ void bm_expand_runs(uint4* dst4, const uint16_t* runs, uint32_t n_runs, uint32_t* warp_par /*[8]*/) {
    bm_zero(dst4);
    __syncthreads();
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst4);
    const uint32_t* r32 = reinterpret_cast<const uint32_t*>(runs);
    for (uint32_t i = threadIdx.x; i < n_runs; i += kEvalThreads) {
        uint32_t v = __ldg(r32 + i);
        uint32_t s = v & 0xffffu, e = (v >> 16) + 1;
        atomicXor(&d32[s >> 5], 1u << (s & 31));
        if (e < 65536u) atomicXor(&d32[e >> 5], 1u << (e & 31));
    }
    __syncthreads();
    // thread t owns kEvalW64PerThread consecutive u64 words
    uint64_t* d64 = reinterpret_cast<uint64_t*>(dst4);
    uint64_t w[kEvalW64PerThread]; uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < kEvalW64PerThread; k++) {
        uint64_t x = d64[kEvalW64PerThread * threadIdx.x + k];
        x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16; x ^= x << 32;
        if (carry) x = ~x;
        carry = (uint32_t)(x >> 63);
        w[k] = x;
    }
    unsigned b = __ballot_sync(0xffffffffu, carry);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t excl = __popc(b & ((1u << lane) - 1u)) & 1u;
    if (lane == 0) warp_par[wid] = __popc(b) & 1u;
    __syncthreads();
    for (int k = 0; k < wid; k++) excl ^= warp_par[k];
#pragma unroll
    for (int k = 0; k < kEvalW64PerThread; k++) d64[kEvalW64PerThread * threadIdx.x + k] = excl ? ~w[k] : w[k];
    __syncthreads();
}

enum { K_PUSH = 0, K_OR, K_AND, K_ANDNOT, K_XOR, K_ORAND, K_ORANDNOT };

// top = f(top, g)   or   below |= top & (~)g   with g streamed from global (bitmap container)
__device__ __forceinline__ void bm_apply_global(int kind, uint4* top, uint4* below, const uint4* g) {
#pragma unroll
    for (int h = 0; h < kEvalU4PerThread; h++) {
        int i = threadIdx.x + h * kEvalThreads;
        uint4 x = ldg_nc(g + i);
        switch (kind) {
            case K_PUSH: top[i] = x; break;
            case K_OR: top[i] = or4(top[i], x); break;
            case K_AND: top[i] = and4(top[i], x); break;
            case K_ANDNOT: top[i] = andn4(top[i], x); break;
            case K_XOR: top[i] = xor4(top[i], x); break;
            case K_ORAND: below[i] = or4(below[i], and4(top[i], x)); break;
            default: below[i] = or4(below[i], andn4(top[i], x)); break;
        }
    }
}
__device__ __forceinline__ void bm_apply_smem(int kind, uint4* top, uint4* below, const uint4* s) {
#pragma unroll
    for (int h = 0; h < kEvalU4PerThread; h++) {
        int i = threadIdx.x + h * kEvalThreads;
        uint4 x = s[i];
        switch (kind) {
            case K_PUSH: top[i] = x; break;
            case K_OR: top[i] = or4(top[i], x); break;
            case K_AND: top[i] = and4(top[i], x); break;
            case K_ANDNOT: top[i] = andn4(top[i], x); break;
            case K_XOR: top[i] = xor4(top[i], x); break;
            case K_ORAND: below[i] = or4(below[i], and4(top[i], x)); break;
            default: below[i] = or4(below[i], andn4(top[i], x)); break;
        }
    }
}

template <int MODE>   // 0: |=   1: &= ~   2: ^=
__device__ __forceinline__ void smem_bit_op(uint32_t* bm, uint32_t v) {
    // red.shared (no return value).  `asm volatile` keeps the reductions in program order, which stops ptxas from
    // hoisting dozens of address/mask computations ahead of them (register pressure decides CTA residency here).
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(bm + (v >> 5));
    uint32_t m = 1u << (v & 31);
    if (MODE == 0) asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(m) : "memory");
    else if (MODE == 1) asm volatile("red.shared.and.b32 [%0], %1;" :: "r"(addr), "r"(~m) : "memory");
    else asm volatile("red.shared.xor.b32 [%0], %1;" :: "r"(addr), "r"(m) : "memory");
}
// one warp scatters a whole array container (16-byte loads, 8 elements per lane per step)
template <int MODE>
__device__ __forceinline__ void warp_scatter_smem_mode(uint32_t* bm, const uint16_t* arr, uint32_t n, int lane) {
    const uint4* a4 = reinterpret_cast<const uint4*>(arr);
    uint32_t n8 = (n + 7) >> 3;
    for (uint32_t i = lane; i < n8; i += 32) {
        uint4 v = ldg_nc(a4 + i);
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
        uint32_t base = i * 8;
        if (base + 8 <= n) {
#pragma unroll
            for (int q = 0; q < 4; q++) { smem_bit_op<MODE>(bm, w[q] & 0xffffu); smem_bit_op<MODE>(bm, w[q] >> 16); }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (base + 2 * q < n) smem_bit_op<MODE>(bm, w[q] & 0xffffu);
                if (base + 2 * q + 1 < n) smem_bit_op<MODE>(bm, w[q] >> 16);
            }
        }
    }
}
__device__ __forceinline__ void warp_scatter_smem(uint32_t* bm, const uint16_t* arr, uint32_t n, int lane) { warp_scatter_smem_mode<0>(bm, arr, n, lane); }
// one warp applies a whole bitmap container with word atomics (safe against concurrent warps)
template <int MODE>
__device__ __forceinline__ void warp_bitmap_atomic(uint32_t* bm, const uint4* g, int lane) {
    for (int i = lane; i < 512; i += 32) {
        uint4 v = ldg_nc(g + i);
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (!w[q]) continue;
            if (MODE == 0) atomicOr(&bm[4 * i + q], w[q]); else if (MODE == 1) atomicAnd(&bm[4 * i + q], ~w[q]); else atomicXor(&bm[4 * i + q], w[q]);
        }
    }
}
// scatters the (up to) 8 u16 values of one 16-byte chunk; rolled on purpose (keeps register pressure low so
// that 4+ CTAs stay resident per SM)
template <int MODE>
__device__ __forceinline__ void scatter_chunk(uint32_t* bm, uint4 v, uint32_t base, uint32_t n) {
    uint64_t a = ((uint64_t)v.y << 32) | v.x, b = ((uint64_t)v.w << 32) | v.z;
    uint32_t cnt = n - base;            // valid elements in this chunk (>= 1)
    if (cnt >= 8) {
#pragma unroll 1
        for (int q = 0; q < 4; q++) { smem_bit_op<MODE>(bm, (uint32_t)a & 0xffffu); smem_bit_op<MODE>(bm, (uint32_t)b & 0xffffu); a >>= 16; b >>= 16; }
    } else {
#pragma unroll 1
        for (uint32_t q = 0; q < cnt; q++) { uint64_t x = q < 4 ? a : b; smem_bit_op<MODE>(bm, (uint32_t)(x >> (16 * (q & 3))) & 0xffffu); }
    }
}
// Batch of commuting row operands (OR / ANDNOT / XOR onto the same target), no barrier in between: the CTA is
// split into groups of G threads, one group per operand (G*16 B contiguous per load), and every thread keeps four
// 16-byte loads in flight before it touches shared memory, so the batch is bandwidth- rather than latency-bound.
template <int MODE>
__device__ __noinline__ void batch_rows(uint32_t* T32, const Resolved* res, int n) {
    const int tid = threadIdx.x;
    int G = kEvalThreads / max(n, 1);
    G = G >= 32 ? 32 : G <= 1 ? 1 : (1 << (31 - __clz(G)));
    const int groups = kEvalThreads / G, g = tid & (G - 1);
    for (int j = tid / G; j < n; j += groups) {
        const Resolved r = res[j];
        if (r.ptr == nullptr) continue;
        if (r.typ == kArray) {
            const uint4* a4 = reinterpret_cast<const uint4*>(r.ptr);
            const uint32_t n8 = (r.card + 7) >> 3;
            for (uint32_t i = g; i < n8; i += FBGPU_BATCH_UNROLL * G) {
                uint4 v[FBGPU_BATCH_UNROLL];
#pragma unroll
                for (int q = 0; q < FBGPU_BATCH_UNROLL; q++) if (i + q * G < n8) v[q] = ldg_nc(a4 + i + q * G);
#pragma unroll
                for (int q = 0; q < FBGPU_BATCH_UNROLL; q++) if (i + q * G < n8) scatter_chunk<MODE>(T32, v[q], (i + q * G) * 8, r.card);
            }
        } else if (r.typ == kBitmap) {
            const uint4* g4 = reinterpret_cast<const uint4*>(r.ptr);
            for (int i = g; i < 512; i += 2 * G) {
                uint4 v[2];
#pragma unroll
                for (int q = 0; q < 2; q++) if (i + q * G < 512) v[q] = ldg_nc(g4 + i + q * G);
#pragma unroll
                for (int q = 0; q < 2; q++) if (i + q * G < 512) {
                    uint32_t w[4] = { v[q].x, v[q].y, v[q].z, v[q].w };
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (!w[c]) continue;
                        uint32_t* dst = &T32[4 * (i + q * G) + c];
                        if (MODE == 0) atomicOr(dst, w[c]); else if (MODE == 1) atomicAnd(dst, ~w[c]); else atomicXor(dst, w[c]);
                    }
                }
            }
        }
    }
}

// ---- alternative batch implementations (selected at compile time by FBGPU_BATCH_IMPL)
template <int MODE>
__device__ __forceinline__ void scatter_chunk_unrolled(uint32_t* bm, uint4 v, uint32_t base, uint32_t n) {
    uint32_t w[4] = { v.x, v.y, v.z, v.w };
    if (base + 8 <= n) {
#pragma unroll
        for (int q = 0; q < 4; q++) { smem_bit_op<MODE>(bm, w[q] & 0xffffu); smem_bit_op<MODE>(bm, w[q] >> 16); }
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (base + 2 * q < n) smem_bit_op<MODE>(bm, w[q] & 0xffffu);
            if (base + 2 * q + 1 < n) smem_bit_op<MODE>(bm, w[q] >> 16);
        }
    }
}
// v1: one warp per operand, plain loop
template <int MODE>
__device__ __forceinline__ void batch_rows_v1(uint32_t* T32, const Resolved* res, int n) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int j = wid; j < n; j += kEvalThreads / 32) {
        const Resolved r = res[j];
        if (r.ptr == nullptr) continue;
        if (r.typ == kArray) {
            const uint4* a4 = reinterpret_cast<const uint4*>(r.ptr);
            const uint32_t n8 = (r.card + 7) >> 3;
            for (uint32_t i = lane; i < n8; i += 32) scatter_chunk_unrolled<MODE>(T32, ldg_nc(a4 + i), i * 8, r.card);
        } else if (r.typ == kBitmap) warp_bitmap_atomic<MODE>(T32, reinterpret_cast<const uint4*>(r.ptr), lane);
    }
}
// v3: one warp per operand, three 16-byte loads per lane in flight and the next operand's loads issued before the
// current operand is scattered (register double buffering)
template <int MODE>
__device__ __forceinline__ void batch_rows_v3(uint32_t* T32, const Resolved* res, int n) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = kEvalThreads / 32;
    uint4 cur[3], nxt[3];
    Resolved rc; rc.ptr = nullptr; rc.card = 0; rc.typ = 0; rc.cnt = 0;
    int j = wid;
    auto issue = [&](const Resolved& r, uint4* buf) {
        if (r.ptr != nullptr && r.typ == kArray) {
            const uint4* a4 = reinterpret_cast<const uint4*>(r.ptr);
            const uint32_t n8 = (r.card + 7) >> 3;
#pragma unroll
            for (int q = 0; q < 3; q++) if (lane + 32 * q < n8) buf[q] = ldg_nc(a4 + lane + 32 * q);
        }
    };
    if (j < n) { rc = res[j]; issue(rc, cur); }
    while (j < n) {
        Resolved rn; rn.ptr = nullptr; rn.card = 0; rn.typ = 0; rn.cnt = 0;
        if (j + NW < n) { rn = res[j + NW]; issue(rn, nxt); }
        if (rc.ptr != nullptr) {
            if (rc.typ == kArray) {
                const uint32_t n8 = (rc.card + 7) >> 3;
#pragma unroll
                for (int q = 0; q < 3; q++) if (lane + 32 * q < n8) scatter_chunk_unrolled<MODE>(T32, cur[q], (lane + 32 * q) * 8, rc.card);
                const uint4* a4 = reinterpret_cast<const uint4*>(rc.ptr);
                for (uint32_t i = lane + 96; i < n8; i += 32) scatter_chunk_unrolled<MODE>(T32, ldg_nc(a4 + i), i * 8, rc.card);
            } else if (rc.typ == kBitmap) warp_bitmap_atomic<MODE>(T32, reinterpret_cast<const uint4*>(rc.ptr), lane);
        }
#pragma unroll
        for (int q = 0; q < 3; q++) cur[q] = nxt[q];
        rc = rn; j += NW;
    }
}
#ifndef FBGPU_BATCH_IMPL
#define FBGPU_BATCH_IMPL 1
#endif
template <int MODE>
__device__ __noinline__ void batch_rows_dispatch(uint32_t* T32, const Resolved* res, int n) {
#if FBGPU_BATCH_IMPL == 1
    batch_rows_v1<MODE>(T32, res, n);
#elif FBGPU_BATCH_IMPL == 2
    batch_rows<MODE>(T32, res, n);
#else
    batch_rows_v3<MODE>(T32, res, n);
#endif
}

struct EvalOut {
    unsigned long long* total;      // += count of every unit (may be null)
    unsigned long long* per_shard;  // [n_shards] += (may be null)
    uint4* bitmaps;                 // [n_units][512] result bitmaps (may be null)
    uint2* info;                    // [n_units] {N, runs} (may be null)
};

// One CTA per (shard, slot) unit, persistent over units.  Dynamic smem: (depth+1) x 8 KiB.
__global__ void __launch_bounds__(kEvalThreads, FBGPU_EVAL_MIN_BLOCKS)
eval_kernel(StoreRef st, const DevOp* __restrict__ prog, int n_ops, int depth,
            const uint64_t* __restrict__ shards, long long n_units, EvalOut out) {
    extern __shared__ uint4 smem4[];
    __shared__ Resolved res[kResolveChunk];
    __shared__ uint32_t warp_tmp[kEvalThreads / 32];
    __shared__ uint32_t warp_tmp2[kEvalThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned long long cta_total = 0;

    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint64_t shard = shards[unit >> 4];
        const int slot = (int)(unit & 15);
        // nibble-packed permutation level -> physical bitmap; levels > top are free, index `depth` is the spare
        uint64_t map = 0xFEDCBA9876543210ull;
        int top = -1;
        auto phys = [&](int level) -> uint4* { return smem4 + (size_t)((map >> (4 * level)) & 15u) * 512; };
        auto swap_levels = [&](int a, int b) {
            uint64_t pa = (map >> (4 * a)) & 15u, pb = (map >> (4 * b)) & 15u;
            map &= ~((15ull << (4 * a)) | (15ull << (4 * b)));
            map |= (pb << (4 * a)) | (pa << (4 * b));
        };
        for (int base = 0; base < n_ops; base += kResolveChunk) {
            int chunk = min(kResolveChunk, n_ops - base);
            __syncthreads();
            if (tid < chunk) {
                DevOp op = prog[base + tid];
                Resolved r; r.ptr = nullptr; r.card = 0; r.typ = 0; r.cnt = 0;
                if (op.op >= D_PUSH_ROW && op.op <= D_ORANDNOT_ROW && op.op != D_PUSH_EMPTY) r = resolve(st, op.fv, shard, op.row, slot);
                res[tid] = r;
            }
            __syncthreads();
            for (int k = 0; k < chunk; k++) {
                const uint8_t opc = prog[base + k].op;
                if (opc == D_PUSH_EMPTY) { top++; bm_zero(phys(top)); __syncthreads(); continue; }
                if (opc == D_SWAP) { swap_levels(top, top - 1); continue; }
                if (opc == D_POP) { top--; continue; }
                if (opc >= D_AND && opc <= D_XOR) {
                    int kind = opc == D_AND ? K_AND : opc == D_OR ? K_OR : opc == D_ANDNOT ? K_ANDNOT : K_XOR;
                    bm_apply_smem(kind, phys(top - 1), nullptr, phys(top));
                    top--; __syncthreads(); continue;
                }
                if (opc == D_OR_ROW || opc == D_ANDNOT_ROW || opc == D_XOR_ROW) {
                    // maximal run of the same commuting row op inside this chunk
                    int e = k + 1;
                    while (e < chunk && prog[base + e].op == opc) e++;
                    uint4* T = phys(top);
                    uint32_t* T32 = reinterpret_cast<uint32_t*>(T);
                    if (opc == D_OR_ROW) batch_rows_dispatch<0>(T32, res + k, e - k);
                    else if (opc == D_ANDNOT_ROW) batch_rows_dispatch<1>(T32, res + k, e - k);
                    else batch_rows_dispatch<2>(T32, res + k, e - k);
                    __syncthreads();
                    for (int j = k; j < e; j++) {          // run containers: CTA-wide expansion, one at a time
                        const Resolved r = res[j];
                        if (r.ptr == nullptr || r.typ != kRun) continue;
                        uint4* S = phys(depth);
                        bm_expand_runs(S, reinterpret_cast<const uint16_t*>(r.ptr), r.cnt, warp_tmp);
                        bm_apply_smem(opc == D_OR_ROW ? K_OR : opc == D_ANDNOT_ROW ? K_ANDNOT : K_XOR, T, nullptr, S);
                        __syncthreads();
                    }
                    k = e - 1;
                    continue;
                }
                // row-operand ops
                int kind = opc == D_PUSH_ROW ? K_PUSH : opc == D_OR_ROW ? K_OR : opc == D_AND_ROW ? K_AND : opc == D_ANDNOT_ROW ? K_ANDNOT
                         : opc == D_XOR_ROW ? K_XOR : opc == D_ORAND_ROW ? K_ORAND : K_ORANDNOT;
                const Resolved r = res[k];
                if (kind == K_PUSH) top++;
                uint4* T = phys(top);
                uint4* B = (kind == K_ORAND || kind == K_ORANDNOT) ? phys(top - 1) : nullptr;
                if (r.ptr == nullptr) {                      // absent container == empty (container_stash.go:38)
                    if (kind == K_PUSH || kind == K_AND) bm_zero(T);
                    else if (kind == K_ORANDNOT) bm_apply_smem(K_OR, B, nullptr, T);
                    __syncthreads();
                } else if (r.typ == kBitmap) {
                    bm_apply_global(kind, T, B, reinterpret_cast<const uint4*>(r.ptr));
                    __syncthreads();
                } else if (r.typ == kArray) {
                    const uint16_t* arr = reinterpret_cast<const uint16_t*>(r.ptr);
                    uint32_t* T32 = reinterpret_cast<uint32_t*>(T);
                    if (kind == K_PUSH) { bm_zero(T); __syncthreads(); bm_scatter<0>(T32, arr, r.card); }
                    else if (kind == K_OR) bm_scatter<0>(T32, arr, r.card);
                    else if (kind == K_ANDNOT) bm_scatter<1>(T32, arr, r.card);
                    else if (kind == K_XOR) bm_scatter<2>(T32, arr, r.card);
                    else if (kind == K_ORAND) bm_filter_scatter(reinterpret_cast<uint32_t*>(B), T32, arr, r.card);
                    else if (kind == K_AND) {
                        uint4* S = phys(depth);
                        bm_zero(S); __syncthreads();
                        bm_filter_scatter(reinterpret_cast<uint32_t*>(S), T32, arr, r.card);
                        swap_levels(top, depth);
                    } else {  // K_ORANDNOT
                        uint4* S = phys(depth);
                        bm_zero(S); __syncthreads();
                        bm_scatter<0>(reinterpret_cast<uint32_t*>(S), arr, r.card); __syncthreads();
                        bm_apply_smem(K_ORANDNOT, T, B, S);
                    }
                    __syncthreads();
                } else {                                     // run container
                    const uint16_t* runs = reinterpret_cast<const uint16_t*>(r.ptr);
                    if (kind == K_PUSH) bm_expand_runs(T, runs, r.cnt, warp_tmp);
                    else {
                        uint4* S = phys(depth);
                        bm_expand_runs(S, runs, r.cnt, warp_tmp);
                        bm_apply_smem(kind, T, B, S);
                        __syncthreads();
                    }
                }
            }
        }
        // ---- unit epilogue: popcount (+ optional bitmap / run statistics for canonical emission)
        uint32_t cnt = 0, nruns = 0;
        if (top >= 0) {
            const uint4* R = phys(top);
            const uint64_t* R64 = reinterpret_cast<const uint64_t*>(R);
#pragma unroll
            for (int h = 0; h < kEvalU4PerThread; h++) {
                int i = tid + h * kEvalThreads;
                uint4 a = R[i];
                cnt += popc4(a);
                if (out.bitmaps) out.bitmaps[(size_t)unit * 512 + i] = a;
                if (out.info) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        int wi = 2 * i + q;
                        uint64_t v = R64[wi];
                        uint64_t prev = wi ? (R64[wi - 1] >> 63) : 0ull;
                        nruns += __popcll(v & ~((v << 1) | prev));   // bitmapCountRuns roaring.go:3372
                    }
                }
            }
        } else if (out.bitmaps) {
#pragma unroll
            for (int h = 0; h < kEvalU4PerThread; h++) out.bitmaps[(size_t)unit * 512 + tid + h * kEvalThreads] = make_uint4(0, 0, 0, 0);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        nruns = __reduce_add_sync(0xffffffffu, nruns);
        __syncthreads();
        if (lane == 0) { warp_tmp[wid] = cnt; warp_tmp2[wid] = nruns; }
        __syncthreads();
        if (tid == 0) {
            uint32_t c = 0, rr = 0;
#pragma unroll
            for (int k = 0; k < kEvalThreads / 32; k++) { c += warp_tmp[k]; rr += warp_tmp2[k]; }
            cta_total += c;
            if (out.per_shard && c) atomicAdd(&out.per_shard[unit >> 4], (unsigned long long)c);
            if (out.info) out.info[unit] = make_uint2(c, rr);
        }
    }
    if (tid == 0 && out.total && cta_total) atomicAdd(out.total, cta_total);
}

// ------------------------------------------------------------------------------------------------
// Warp-level intersection count of two located containers; `bm` is the warp's private 8 KiB smem bitmap.
// Follows the dispatch of intersectionCount (roaring.go:4477-4512) incl. the full/empty short-circuits.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_zero(uint32_t* bm, int lane) {
    uint4* b4 = reinterpret_cast<uint4*>(bm);
#pragma unroll 4
    for (int i = lane; i < 512; i += 32) b4[i] = make_uint4(0, 0, 0, 0);
}
// per-lane partial count of array elements found in a shared-memory bitmap
__device__ __forceinline__ uint32_t warp_probe_smem(const uint32_t* bm, const uint16_t* arr, uint32_t n, int lane) {
    const uint4* a4 = reinterpret_cast<const uint4*>(arr);
    uint32_t n8 = (n + 7) >> 3, c = 0;
    for (uint32_t i = lane; i < n8; i += 32) {
        uint4 v = ldg_nc(a4 + i);
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
        uint32_t base = i * 8;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t lo = w[q] & 0xffffu, hi = w[q] >> 16;
            if (base + 2 * q < n) c += (bm[lo >> 5] >> (lo & 31)) & 1u;
            if (base + 2 * q + 1 < n) c += (bm[hi >> 5] >> (hi & 31)) & 1u;
        }
    }
    return c;
}
// per-lane partial count of array elements found in a global-memory bitmap
__device__ __forceinline__ uint32_t warp_probe_global(const uint32_t* g, const uint16_t* arr, uint32_t n, int lane) {
    uint32_t c = 0;
    for (uint32_t i = lane; i < n; i += 32) { uint32_t v = __ldg(arr + i); c += (__ldg(g + (v >> 5)) >> (v & 31)) & 1u; }
    return c;
}
// warp-private run expansion (delta + prefix-xor over 2048 u32 words; lane owns 64 consecutive words)
__device__ __forceinline__ void warp_expand_runs(uint32_t* bm, const uint16_t* runs, uint32_t n_runs, int lane) {
    warp_zero(bm, lane);
    __syncwarp();
    const uint32_t* r32 = reinterpret_cast<const uint32_t*>(runs);
    for (uint32_t i = lane; i < n_runs; i += 32) {
        uint32_t v = __ldg(r32 + i);
        uint32_t s = v & 0xffffu, e = (v >> 16) + 1;
        atomicXor(&bm[s >> 5], 1u << (s & 31));
        if (e < 65536u) atomicXor(&bm[e >> 5], 1u << (e & 31));
    }
    __syncwarp();
    // pass 1: parity of each lane's 64-word block
    uint32_t par = 0;
    for (int k = 0; k < 64; k++) par ^= bm[lane * 64 + k];
    par = __popc(par) & 1u;
    unsigned b = __ballot_sync(0xffffffffu, par);
    uint32_t carry = __popc(b & ((1u << lane) - 1u)) & 1u;
    for (int k = 0; k < 64; k++) {
        uint32_t x = bm[lane * 64 + k];
        x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
        if (carry) x = ~x;
        carry = x >> 31;
        bm[lane * 64 + k] = x;
    }
    __syncwarp();
}
// per-lane partial: |bitmap(global) ∩ smem bitmap|
__device__ __forceinline__ uint32_t warp_and_count_gs(const uint4* g, const uint32_t* bm, int lane) {
    const uint4* s4 = reinterpret_cast<const uint4*>(bm);
    uint32_t c = 0;
#pragma unroll 4
    for (int i = lane; i < 512; i += 32) c += popc4(and4(ldg_nc(g + i), s4[i]));
    return c;
}
// per-lane partial: number of set bits of a u32-word bitmap (smem or global) inside [s, l]
__device__ __forceinline__ uint32_t range_count32(const uint32_t* bm, uint32_t s, uint32_t l) {
    uint32_t ws = s >> 5, wl = l >> 5;
    uint32_t ms = 0xffffffffu << (s & 31), ml = 0xffffffffu >> (31 - (l & 31));
    if (ws == wl) return __popc(bm[ws] & ms & ml);
    uint32_t c = __popc(bm[ws] & ms) + __popc(bm[wl] & ml);
    for (uint32_t k = ws + 1; k < wl; k++) c += __popc(bm[k]);
    return c;
}

// returns the full count (reduced over the warp, valid in all lanes)
__device__ __forceinline__ uint32_t warp_intersection_count(Resolved a, Resolved b, uint32_t* bm, int lane) {
    if (a.ptr == nullptr || b.ptr == nullptr) return 0;
    if (a.card == kFull) return b.card;                       // roaring.go:4478-4483
    if (b.card == kFull) return a.card;
    // order so that typ(a) <= typ(b) in {array, bitmap, run} with arrays first
    if (a.typ != kArray && b.typ == kArray) { Resolved t = a; a = b; b = t; }
    uint32_t c = 0;
    if (a.typ == kArray && b.typ == kArray) {                 // array x array: build the smaller, probe the larger
        if (a.card > b.card) { Resolved t = a; a = b; b = t; }
        warp_zero(bm, lane); __syncwarp();
        warp_scatter_smem(bm, reinterpret_cast<const uint16_t*>(a.ptr), a.card, lane); __syncwarp();
        c = warp_probe_smem(bm, reinterpret_cast<const uint16_t*>(b.ptr), b.card, lane); __syncwarp();
    } else if (a.typ == kArray && b.typ == kBitmap) {         // roaring.go:4596
        c = warp_probe_global(reinterpret_cast<const uint32_t*>(b.ptr), reinterpret_cast<const uint16_t*>(a.ptr), a.card, lane);
    } else if (a.typ == kArray) {                             // array x run: roaring.go:4537
        warp_expand_runs(bm, reinterpret_cast<const uint16_t*>(b.ptr), b.cnt, lane);
        c = warp_probe_smem(bm, reinterpret_cast<const uint16_t*>(a.ptr), a.card, lane); __syncwarp();
    } else if (a.typ == kBitmap && b.typ == kBitmap) {        // roaring.go:4611
        const uint4* x = reinterpret_cast<const uint4*>(a.ptr); const uint4* y = reinterpret_cast<const uint4*>(b.ptr);
#pragma unroll 4
        for (int i = lane; i < 512; i += 32) c += popc4(and4(ldg_nc(x + i), ldg_nc(y + i)));
    } else {
        // one side is a run; make it `b`
        if (a.typ == kRun && b.typ != kRun) { Resolved t = a; a = b; b = t; }
        warp_expand_runs(bm, reinterpret_cast<const uint16_t*>(b.ptr), b.cnt, lane);
        if (a.typ == kBitmap) c = warp_and_count_gs(reinterpret_cast<const uint4*>(a.ptr), bm, lane);   // roaring.go:4588
        else {                                                // run x run: roaring.go:4555
            const uint32_t* r32 = reinterpret_cast<const uint32_t*>(a.ptr);
            for (uint32_t i = lane; i < a.cnt; i += 32) { uint32_t v = __ldg(r32 + i); c += range_count32(bm, v & 0xffffu, v >> 16); }
        }
        __syncwarp();
    }
    return __reduce_add_sync(0xffffffffu, c);
}

constexpr int kPairWarps = 8;

// Count(Intersect(Row(fvA,rowA), Row(fvB,rowB))): one warp per (shard, slot) container pair.
__global__ void __launch_bounds__(kPairWarps * 32)
pair_count_kernel(StoreRef st, uint32_t fvA, uint64_t rowA, uint32_t fvB, uint64_t rowB,
                  const uint64_t* __restrict__ shards, long long n_units,
                  unsigned long long* total, unsigned long long* per_shard) {
    extern __shared__ uint32_t smem32[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* bm = smem32 + wid * 2048;
    unsigned long long acc = 0;
    const long long stride = (long long)gridDim.x * kPairWarps;
    for (long long unit = (long long)blockIdx.x * kPairWarps + wid; unit < n_units; unit += stride) {
        const uint64_t shard = shards[unit >> 4];
        const int slot = (int)(unit & 15);
        // lanes 0 and 1 walk the two descriptor chains concurrently, then broadcast
        Resolved r; r.ptr = nullptr; r.card = 0; r.typ = 0; r.cnt = 0;
        if (lane == 0) r = resolve(st, fvA, shard, rowA, slot);
        else if (lane == 1) r = resolve(st, fvB, shard, rowB, slot);
        Resolved a, b;
        unsigned long long pa = __shfl_sync(0xffffffffu, (unsigned long long)r.ptr, 0), pb = __shfl_sync(0xffffffffu, (unsigned long long)r.ptr, 1);
        uint32_t meta = ((uint32_t)r.typ << 16) | r.cnt;
        a.ptr = (const void*)pa; b.ptr = (const void*)pb;
        a.card = __shfl_sync(0xffffffffu, r.card, 0); b.card = __shfl_sync(0xffffffffu, r.card, 1);
        uint32_t ma = __shfl_sync(0xffffffffu, meta, 0), mb = __shfl_sync(0xffffffffu, meta, 1);
        a.typ = ma >> 16; a.cnt = ma & 0xffff; b.typ = mb >> 16; b.cnt = mb & 0xffff;
        uint32_t c = warp_intersection_count(a, b, bm, lane);
        acc += c;
        if (per_shard && c && lane == 0) atomicAdd(&per_shard[unit >> 4], (unsigned long long)c);
    }
    if (lane == 0 && total && acc) atomicAdd(total, acc);
}

// ------------------------------------------------------------------------------------------------
// Per-row counts (TopK / TopN-with-ids): one warp per (shard, requested row); filter is an optional
// per-unit bitmap produced by eval_kernel.  doTopK executor.go:2719-2738.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_count_vs_global_bitmap(Resolved a, const uint32_t* fb, uint32_t* bm, int lane) {
    uint32_t c = 0;
    if (a.typ == kArray) c = warp_probe_global(fb, reinterpret_cast<const uint16_t*>(a.ptr), a.card, lane);
    else if (a.typ == kBitmap) {
        const uint4* x = reinterpret_cast<const uint4*>(a.ptr); const uint4* y = reinterpret_cast<const uint4*>(fb);
#pragma unroll 4
        for (int i = lane; i < 512; i += 32) c += popc4(and4(ldg_nc(x + i), y[i]));
    } else {
        const uint32_t* r32 = reinterpret_cast<const uint32_t*>(a.ptr);
        for (uint32_t i = lane; i < a.cnt; i += 32) { uint32_t v = __ldg(r32 + i); c += range_count32(fb, v & 0xffffu, v >> 16); }
    }
    return __reduce_add_sync(0xffffffffu, c);
}

__global__ void __launch_bounds__(kPairWarps * 32)
row_count_kernel(StoreRef st, uint32_t fv, const uint64_t* __restrict__ row_ids, int n_rows,
                 const uint64_t* __restrict__ shards, long long n_shards,
                 const uint4* __restrict__ filter_bitmaps /* [n_shards*16][512] or null */,
                 unsigned long long* out_counts /* [n_rows] */) {
    extern __shared__ uint32_t smem32[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* bm = smem32 + wid * 2048;
    const long long n_tasks = n_shards * (long long)n_rows;
    const long long stride = (long long)gridDim.x * kPairWarps;
    for (long long t = (long long)blockIdx.x * kPairWarps + wid; t < n_tasks; t += stride) {
        const long long si = t / n_rows; const int ri = (int)(t - si * n_rows);
        const uint64_t shard = shards[si], row = row_ids[ri];
        // lanes 0..15 resolve the 16 slots of the row concurrently
        Resolved r; r.ptr = nullptr; r.card = 0; r.typ = 0; r.cnt = 0;
        if (lane < 16) r = resolve(st, fv, shard, row, lane);
        unsigned present = __ballot_sync(0xffffffffu, r.ptr != nullptr);
        unsigned long long acc = 0;
        while (present) {
            int s = __ffs(present) - 1; present &= present - 1;
            Resolved a;
            a.ptr = (const void*)__shfl_sync(0xffffffffu, (unsigned long long)r.ptr, s);
            a.card = __shfl_sync(0xffffffffu, r.card, s);
            uint32_t meta = __shfl_sync(0xffffffffu, ((uint32_t)r.typ << 16) | r.cnt, s);
            a.typ = meta >> 16; a.cnt = meta & 0xffff;
            if (!filter_bitmaps) acc += a.card;
            else acc += warp_count_vs_global_bitmap(a, reinterpret_cast<const uint32_t*>(filter_bitmaps + ((size_t)si * 16 + s) * 512), bm, lane);
        }
        if (lane == 0 && acc) atomicAdd(&out_counts[ri], acc);
    }
}

// ------------------------------------------------------------------------------------------------
// Canonical emission of result bitmaps (Row results): optimize() roaring.go:3412-3461 decides the encoding
// on the host from {N, runs}; this kernel writes the payload (array / run / bitmap) at the given offset.
// ------------------------------------------------------------------------------------------------
constexpr int kEmitThreads = 256;   // thread t owns u64 words 4t..4t+3
struct EmitUnit { uint64_t offset; uint32_t unit; uint32_t typ; };

__global__ void __launch_bounds__(kEmitThreads)
canon_emit_kernel(const uint4* __restrict__ bitmaps, const EmitUnit* __restrict__ units, int n_emit, uint8_t* __restrict__ out) {
    __shared__ uint32_t wsum[kEmitThreads / 32], wsum2[kEmitThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int e = blockIdx.x; e < n_emit; e += gridDim.x) {
        EmitUnit u = units[e];
        const uint64_t* src = reinterpret_cast<const uint64_t*>(bitmaps + (size_t)u.unit * 512);
        if (u.typ == kBitmap) {
            uint4* o = reinterpret_cast<uint4*>(out + u.offset);   // offsets of bitmap payloads are only 2-byte aligned in the
            const uint4* s4 = bitmaps + (size_t)u.unit * 512;      // roaring file; the host keeps emit buffers 16 B aligned per unit
            o[tid] = s4[tid]; o[tid + kEmitThreads] = s4[tid + kEmitThreads];
            continue;
        }
        // thread t owns words 4t..4t+3; compute exclusive prefix of element count (array) or start/end counts (run)
        uint64_t w[4]; uint32_t c1 = 0, c2 = 0;
        uint64_t prev = tid ? (src[4 * tid - 1] >> 63) : 0ull;
        uint64_t starts[4], ends[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w[k] = src[4 * tid + k];
            if (u.typ == kArray) c1 += __popcll(w[k]);
            else {
                uint64_t nextbit = (4 * tid + k + 1 < 1024) ? (src[4 * tid + k + 1] & 1ull) : 0ull;
                starts[k] = w[k] & ~((w[k] << 1) | prev);
                ends[k] = w[k] & ~((w[k] >> 1) | (nextbit << 63));
                c1 += __popcll(starts[k]); c2 += __popcll(ends[k]);
                prev = w[k] >> 63;
            }
        }
        // block exclusive scan of c1 (and c2)
        uint32_t i1 = c1, i2 = c2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t x = __shfl_up_sync(0xffffffffu, i1, d), y = __shfl_up_sync(0xffffffffu, i2, d); if (lane >= d) { i1 += x; i2 += y; } }
        __syncthreads();
        if (lane == 31) { wsum[wid] = i1; wsum2[wid] = i2; }
        __syncthreads();
        uint32_t b1 = 0, b2 = 0;
        for (int k = 0; k < wid; k++) { b1 += wsum[k]; b2 += wsum2[k]; }
        uint32_t p1 = b1 + i1 - c1, p2 = b2 + i2 - c2;
        uint16_t* o16 = reinterpret_cast<uint16_t*>(out + u.offset);
        if (u.typ == kArray) {
#pragma unroll
            for (int k = 0; k < 4; k++) { uint64_t v = w[k]; while (v) { int bit = __ffsll((long long)v) - 1; o16[p1++] = (uint16_t)((4 * tid + k) * 64 + bit); v &= v - 1; } }
        } else {   // run payload: u16 count, then {start,last} pairs (roaring.go:19-51)
            if (tid == 0) o16[0] = (uint16_t)0;  // patched below by the thread holding the total
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint64_t v = starts[k]; while (v) { int bit = __ffsll((long long)v) - 1; o16[1 + 2 * (p1++)] = (uint16_t)((4 * tid + k) * 64 + bit); v &= v - 1; }
                v = ends[k]; while (v) { int bit = __ffsll((long long)v) - 1; o16[2 + 2 * (p2++)] = (uint16_t)((4 * tid + k) * 64 + bit); v &= v - 1; }
            }
            __syncthreads();
            if (tid == kEmitThreads - 1) o16[0] = (uint16_t)p1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GroupBy(Rows(a), Rows(b)) [+ filter]: one CTA per (shard, slot).  Column-keyed join instead of the
// reference's |A|x|B| nested intersectionCount loop (executor.go:8880-8934): field-a rows are inserted into a
// 65,536-entry column table in shared memory (chained for multi-valued columns), field-b rows are streamed
// against it and bump counts[i*nB + j].  Dense (bitmap/run) a-rows take a bitmap pass instead.
// Shared memory: head[65536] u16 (128 KiB) + pool[kGbPool] u32 (row<<16|next) + 8 KiB bitmap.
// ------------------------------------------------------------------------------------------------
constexpr int kGbThreads = 256;
constexpr int kGbPool = 16384;          // chained entries per pass (64 KiB)
constexpr uint32_t kGbDenseCard = 4096; // a-rows at/above this cardinality (or non-array) use the bitmap pass

template <class F>
__device__ __forceinline__ void warp_for_each(const Resolved& c, int lane, F f) {
    if (c.typ == kArray) {
        const uint16_t* a = reinterpret_cast<const uint16_t*>(c.ptr);
        for (uint32_t i = lane; i < c.card; i += 32) f((uint32_t)__ldg(a + i));
    } else if (c.typ == kBitmap) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(c.ptr);
        for (uint32_t i = lane; i < 2048; i += 32) { uint32_t v = __ldg(w + i); while (v) { int b = __ffs(v) - 1; f(i * 32 + b); v &= v - 1; } }
    } else {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(c.ptr);
        for (uint32_t i = 0; i < c.cnt; i++) { uint32_t v = __ldg(r + i); uint32_t s = v & 0xffffu, l = v >> 16; for (uint32_t x = s + lane; x <= l; x += 32) f(x); }
    }
}

__global__ void __launch_bounds__(kGbThreads)
groupby_kernel(StoreRef st, uint32_t fvA, const uint64_t* __restrict__ rowsA, int nA,
               uint32_t fvB, const uint64_t* __restrict__ rowsB, int nB,
               const uint64_t* __restrict__ shards, long long n_units,
               const uint4* __restrict__ filter_bitmaps /* per unit or null */,
               unsigned long long* counts /* [nA*nB] */) {
    extern __shared__ uint8_t gsm[];
    uint16_t* head = reinterpret_cast<uint16_t*>(gsm);                       // 128 KiB
    uint32_t* pool = reinterpret_cast<uint32_t*>(gsm + 131072);              // 64 KiB
    uint32_t* fbm = reinterpret_cast<uint32_t*>(gsm + 131072 + kGbPool * 4); // 8 KiB (dense a-row / filter scratch)
    __shared__ uint32_t any_a;
    __shared__ uint32_t s_need[kGbThreads / 32], s_dense[kGbThreads / 32];
    __shared__ Resolved dense_c;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = kGbThreads / 32;
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint64_t shard = shards[unit >> 4];
        const int slot = (int)(unit & 15);
        const uint32_t* flt = filter_bitmaps ? reinterpret_cast<const uint32_t*>(filter_bitmaps + (size_t)unit * 512) : nullptr;
        // executor.go:8769-8772: a shard missing either fragment contributes nothing
        __syncthreads();
        if (tid == 0) {
            bool ok = fvA < st.n_views && fvB < st.n_views;
            if (ok) { ViewTab va = st.views[fvA], vb = st.views[fvB]; ok = shard < va.n_shards && shard < vb.n_shards && st.shardmap[va.shard_off + shard] >= 0 && st.shardmap[vb.shard_off + shard] >= 0; }
            any_a = ok ? 1u : 0u;
        }
        __syncthreads();
        if (!any_a) continue;
        int ia = 0;
        while (ia < nA) {
            // ---- sparse pass: clear the column table, insert a-rows ia.. in row order until the pool is full or a
            //      dense row is met.  Pool slots are assigned by a deterministic prefix over the rows' cardinalities.
            __syncthreads();
            { uint4* h4 = reinterpret_cast<uint4*>(head); for (int i = tid; i < 8192; i += kGbThreads) h4[i] = make_uint4(0, 0, 0, 0); }
            __syncthreads();
            uint32_t pool_base = 0; int pass_end = nA;
            for (int base = ia; base < nA; base += nwarps) {
                int i = base + wid;
                Resolved c; c.ptr = nullptr; c.card = 0; c.typ = 0; c.cnt = 0;
                if (i < nA) {
                    if (lane == 0) c = resolve(st, fvA, shard, rowsA[i], slot);
                    c.ptr = (const void*)__shfl_sync(0xffffffffu, (unsigned long long)c.ptr, 0); c.card = __shfl_sync(0xffffffffu, c.card, 0);
                    uint32_t m = __shfl_sync(0xffffffffu, ((uint32_t)c.typ << 16) | c.cnt, 0); c.typ = m >> 16; c.cnt = m & 0xffff;
                }
                bool is_dense = c.ptr && (c.typ != kArray || c.card >= kGbDenseCard);
                uint32_t need = (c.ptr && !is_dense) ? c.card : 0u;
                if (lane == 0) { s_need[wid] = need; s_dense[wid] = is_dense ? 1u : 0u; }
                __syncthreads();
                uint32_t running = pool_base, my_start = 0; int stop_at = -1;
                for (int w = 0; w < nwarps; w++) {
                    int idx = base + w; if (idx >= nA) break;
                    if (s_dense[w] || running + s_need[w] > (uint32_t)kGbPool) { stop_at = idx; break; }
                    if (w == wid) my_start = running;
                    running += s_need[w];
                }
                bool my_ok = i < nA && (stop_at < 0 || i < stop_at);
                if (my_ok && need) {
                    const uint16_t* a = reinterpret_cast<const uint16_t*>(c.ptr);
                    for (uint32_t k = lane; k < c.card; k += 32) {
                        uint32_t col = __ldg(a + k);
                        if (flt && !((__ldg(flt + (col >> 5)) >> (col & 31)) & 1u)) continue;
                        uint32_t ent = my_start + k + 1;      // 1-based entry index (kGbPool < 65535 fits 16 bits)
                        unsigned short old = head[col], assumed;
                        do { assumed = old; pool[ent - 1] = ((uint32_t)i << 16) | assumed; __threadfence_block(); old = atomicCAS(&head[col], assumed, (unsigned short)ent); } while (old != assumed);
                    }
                }
                pool_base = running;
                __syncthreads();
                if (stop_at >= 0) { pass_end = stop_at; break; }
            }
            __syncthreads();
            const int pass_lo = ia, pass_hi = pass_end;
            // ---- probe: stream b rows
            for (int j = wid; j < nB && pass_hi > pass_lo; j += nwarps) {
                Resolved c; c.ptr = nullptr; c.card = 0; c.typ = 0; c.cnt = 0;
                if (lane == 0) c = resolve(st, fvB, shard, rowsB[j], slot);
                c.ptr = (const void*)__shfl_sync(0xffffffffu, (unsigned long long)c.ptr, 0); c.card = __shfl_sync(0xffffffffu, c.card, 0);
                uint32_t m = __shfl_sync(0xffffffffu, ((uint32_t)c.typ << 16) | c.cnt, 0); c.typ = m >> 16; c.cnt = m & 0xffff;
                if (!c.ptr) continue;
                warp_for_each(c, lane, [&](uint32_t col) {
                    uint32_t ent = head[col];
                    while (ent) { uint32_t pe = pool[ent - 1]; uint32_t i = pe >> 16; if ((int)i >= pass_lo && (int)i < pass_hi) atomicAdd(&counts[(size_t)i * nB + j], 1ull); ent = pe & 0xffffu; }
                });
            }
            __syncthreads();
            ia = pass_hi;
            // ---- dense pass for the row that ended the sparse pass (if it is dense)
            if (ia < nA) {
                Resolved c; c.ptr = nullptr; c.card = 0; c.typ = 0; c.cnt = 0;
                if (tid == 0) { c = resolve(st, fvA, shard, rowsA[ia], slot); }
                // broadcast through smem
                if (tid == 0) dense_c = c;
                __syncthreads();
                c = dense_c;
                bool is_dense = c.ptr && (c.typ != kArray || c.card >= kGbDenseCard);
                if (is_dense) {
                    // expand into fbm (u32[2048]) and AND with filter
                    uint4* f4 = reinterpret_cast<uint4*>(fbm);
                    for (int i = tid; i < 512; i += kGbThreads) f4[i] = make_uint4(0, 0, 0, 0);
                    __syncthreads();
                    if (c.typ == kBitmap) { const uint4* g = reinterpret_cast<const uint4*>(c.ptr); for (int i = tid; i < 512; i += kGbThreads) f4[i] = ldg_nc(g + i); }
                    else if (c.typ == kArray) { const uint16_t* a = reinterpret_cast<const uint16_t*>(c.ptr); for (uint32_t k = tid; k < c.card; k += kGbThreads) { uint32_t v = __ldg(a + k); atomicOr(&fbm[v >> 5], 1u << (v & 31)); } }
                    else { const uint32_t* r = reinterpret_cast<const uint32_t*>(c.ptr);
                        for (uint32_t k = wid; k < c.cnt; k += nwarps) { uint32_t v = __ldg(r + k); uint32_t s = v & 0xffffu, l = v >> 16;
                            for (uint32_t w = (s >> 5) + lane; w <= (l >> 5); w += 32) { uint32_t mask = 0xffffffffu; if (w == (s >> 5)) mask &= 0xffffffffu << (s & 31); if (w == (l >> 5)) mask &= 0xffffffffu >> (31 - (l & 31)); atomicOr(&fbm[w], mask); } } }
                    __syncthreads();
                    if (flt) { const uint4* g = reinterpret_cast<const uint4*>(flt); for (int i = tid; i < 512; i += kGbThreads) f4[i] = and4(f4[i], g[i]); __syncthreads(); }
                    for (int j = wid; j < nB; j += nwarps) {
                        Resolved b; b.ptr = nullptr; b.card = 0; b.typ = 0; b.cnt = 0;
                        if (lane == 0) b = resolve(st, fvB, shard, rowsB[j], slot);
                        b.ptr = (const void*)__shfl_sync(0xffffffffu, (unsigned long long)b.ptr, 0); b.card = __shfl_sync(0xffffffffu, b.card, 0);
                        uint32_t m = __shfl_sync(0xffffffffu, ((uint32_t)b.typ << 16) | b.cnt, 0); b.typ = m >> 16; b.cnt = m & 0xffff;
                        if (!b.ptr) continue;
                        uint32_t cc = 0;
                        if (b.typ == kArray) cc = warp_probe_smem(fbm, reinterpret_cast<const uint16_t*>(b.ptr), b.card, lane);
                        else if (b.typ == kBitmap) cc = warp_and_count_gs(reinterpret_cast<const uint4*>(b.ptr), fbm, lane);
                        else { const uint32_t* r = reinterpret_cast<const uint32_t*>(b.ptr); for (uint32_t k = lane; k < b.cnt; k += 32) { uint32_t v = __ldg(r + k); cc += range_count32(fbm, v & 0xffffu, v >> 16); } }
                        cc = __reduce_add_sync(0xffffffffu, cc);
                        if (lane == 0 && cc) atomicAdd(&counts[(size_t)ia * nB + j], (unsigned long long)cc);
                    }
                    __syncthreads();
                    ia++;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace fbgpu
