// Micro-benchmark (design exploration, not product): array x array intersection-count strategies for ~1 % density
// containers, round 2.  Builds on aa_variants.cu's finding (warp-private bitmap + red.shared is the best shape, limited by
// shared-memory bank conflicts and the per-pair 8 KiB wipe) and measures the candidates for the product kernel:
//   w1   baseline: 1 warp / pair, wipe + scatter + probe (sorted payload)            [= pair_count_kernel of round 1]
//   w2   1 warp / pair, no wipe: scatter, probe, un-scatter                          (sorted | bank-striped payload)
//   w3   w2 + the next pair's chunks are loaded into registers before the current pair is processed
//   c1   CTA-cooperative: T threads / pair, one 8 KiB bitmap per CTA, 3 barriers per pair, next pair prefetched
//   c3   c1 with three rotating bitmaps: one barrier per pair (unscatter k-1, probe k, scatter k+1 share a phase)
// Payload layout: containers back to back, 16-byte aligned (as the product arena), lengths in a side array.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../featurebase_b200/csrc/stripe.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int STRIDE = 832;   // u16 slots per container (1664 B)

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void gen(uint16_t* data, int* len, int n_cont, uint32_t thresh, uint64_t seed) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_cont) return;
  uint16_t* out = data + (size_t)w * STRIDE;
  int n = 0;
  for (int base = 0; base < 65536; base += 32) {
    uint32_t v = base + lane;
    uint32_t r = (uint32_t)(splitmix(seed ^ ((uint64_t)w << 20) ^ v) >> 32);
    bool keep = r < thresh;
    unsigned m = __ballot_sync(0xffffffff, keep);
    if (keep) { int pos = n + __popc(m & ((1u << lane) - 1)); if (pos < STRIDE) out[pos] = (uint16_t)v; }
    n += __popc(m);
  }
  if (lane == 0) len[w] = n < STRIDE ? n : STRIDE;
}

__device__ __forceinline__ uint4 ldg_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t off_lo(uint32_t w) { return __umulhi(w & 0xffe0u, 1u << 29); }
__device__ __forceinline__ uint32_t off_hi(uint32_t w) { return __umulhi(w & 0xffe00000u, 1u << 13); }
template <int MODE> __device__ __forceinline__ void bit_op(uint32_t sb, uint32_t off, uint32_t sh) {
  const uint32_t addr = sb + off, m = 1u << (sh & 31);
  if (MODE == 0) asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(m) : "memory");
  else asm volatile("red.shared.and.b32 [%0], %1;" :: "r"(addr), "r"(~m) : "memory");
}
// the loader pads the last chunk with copies of the last element: OR / AND-NOT of a bit twice is harmless
template <int MODE> __device__ __forceinline__ void scatter8(uint32_t sb, uint4 v) {
  uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
  for (int q = 0; q < 4; q++) { bit_op<MODE>(sb, off_lo(w[q]), w[q]); bit_op<MODE>(sb, off_hi(w[q]), w[q] >> 16); }
}
__device__ __forceinline__ uint32_t lds(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t probe8(uint32_t sb, uint4 v, uint32_t base, uint32_t n) {
  uint32_t w[4] = { v.x, v.y, v.z, v.w }, c = 0;
  if (base + 8 <= n) {
#pragma unroll
    for (int q = 0; q < 4; q++) { c += (lds(sb + off_lo(w[q])) >> (w[q] & 31)) & 1u; c += (lds(sb + off_hi(w[q])) >> ((w[q] >> 16) & 31)) & 1u; }
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (base + 2 * q < n) c += (lds(sb + off_lo(w[q])) >> (w[q] & 31)) & 1u;
      if (base + 2 * q + 1 < n) c += (lds(sb + off_hi(w[q])) >> ((w[q] >> 16) & 31)) & 1u;
    }
  }
  return c;
}

// ---- w1: baseline
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) w1_wipe(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  extern __shared__ uint32_t smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048; const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  unsigned total = 0;
  for (int w = blockIdx.x * WARPS + wid; w < n_pairs; w += gridDim.x * WARPS) {
    const uint4* A = (const uint4*)(data + (size_t)(2 * w) * STRIDE); const uint4* B = (const uint4*)(data + (size_t)(2 * w + 1) * STRIDE);
    const uint32_t na = len[2 * w], nb = len[2 * w + 1], na8 = (na + 7) >> 3, nb8 = (nb + 7) >> 3;
    uint4 va[3], vb[3];
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) va[q] = ldg_nc(A + lane + 32 * q);
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < nb8) vb[q] = ldg_nc(B + lane + 32 * q);
    for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) scatter8<0>(sb, va[q]);
    for (uint32_t i = lane + 96; i < na8; i += 32) scatter8<0>(sb, ldg_nc(A + i));
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < nb8) total += probe8(sb, vb[q], (lane + 32 * q) * 8, nb);
    for (uint32_t i = lane + 96; i < nb8; i += 32) total += probe8(sb, ldg_nc(B + i), i * 8, nb);
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

// ---- w2: no wipe (scatter, probe, un-scatter); w3 = w2 with the next pair's loads issued first
template <int WARPS, bool PREFETCH>
__global__ void __launch_bounds__(WARPS * 32) w2_unscatter(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  extern __shared__ uint32_t smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048; const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  unsigned total = 0;
  const int stride = gridDim.x * WARPS;
  int w = blockIdx.x * WARPS + wid;
  uint4 va[3], vb[3]; uint32_t na = 0, nb = 0;
  auto load = [&](int p, uint4* xa, uint4* xb, uint32_t& la, uint32_t& lb) {
    const uint4* A = (const uint4*)(data + (size_t)(2 * p) * STRIDE); const uint4* B = (const uint4*)(data + (size_t)(2 * p + 1) * STRIDE);
    la = len[2 * p]; lb = len[2 * p + 1];
    const uint32_t na8 = (la + 7) >> 3, nb8 = (lb + 7) >> 3;
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) xa[q] = ldg_nc(A + lane + 32 * q);
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < nb8) xb[q] = ldg_nc(B + lane + 32 * q);
  };
  if (PREFETCH && w < n_pairs) load(w, va, vb, na, nb);
  for (; w < n_pairs; w += stride) {
    uint4 xa[3], xb[3]; uint32_t la, lb;
    if (PREFETCH) {
#pragma unroll
      for (int q = 0; q < 3; q++) { xa[q] = va[q]; xb[q] = vb[q]; }
      la = na; lb = nb;
      if (w + stride < n_pairs) load(w + stride, va, vb, na, nb);
    } else load(w, xa, xb, la, lb);
    const uint32_t na8 = (la + 7) >> 3, nb8 = (lb + 7) >> 3;
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) scatter8<0>(sb, xa[q]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < nb8) total += probe8(sb, xb[q], (lane + 32 * q) * 8, lb);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) scatter8<1>(sb, xa[q]);
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

// ---- c1: CTA-cooperative, T threads per pair (T*8 >= elements), one bitmap, three barriers per pair
template <int T>
__global__ void __launch_bounds__(T) c1_coop(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  __shared__ __align__(16) uint32_t bm[2048];
  const int tid = threadIdx.x;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  for (int i = tid; i < 512; i += T) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  unsigned total = 0;
  int w = blockIdx.x;
  uint4 va = make_uint4(0, 0, 0, 0), vb = va; uint32_t na = 0, nb = 0;
  auto load = [&](int p) {
    na = len[2 * p]; nb = len[2 * p + 1];
    if ((uint32_t)tid * 8 < na) va = ldg_nc((const uint4*)(data + (size_t)(2 * p) * STRIDE) + tid);
    if ((uint32_t)tid * 8 < nb) vb = ldg_nc((const uint4*)(data + (size_t)(2 * p + 1) * STRIDE) + tid);
  };
  if (w < n_pairs) load(w);
  for (; w < n_pairs; w += gridDim.x) {
    const uint4 xa = va, xb = vb; const uint32_t la = na, lb = nb;
    if (w + (int)gridDim.x < n_pairs) load(w + gridDim.x);
    const bool ha = (uint32_t)tid * 8 < la, hb = (uint32_t)tid * 8 < lb;
    if (ha) scatter8<0>(sb, xa);
    __syncthreads();
    if (hb) total += probe8(sb, xb, tid * 8, lb);
    __syncthreads();
    if (ha) scatter8<1>(sb, xa);
    __syncthreads();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if ((tid & 31) == 0 && total) atomicAdd(out, (unsigned long long)total);
}

// ---- c3: three rotating bitmaps, one barrier per pair.  Phase j: scatter A_j into bm[j%3], probe B_{j-1} in bm[(j-1)%3],
// un-scatter A_{j-2} from bm[(j-2)%3]; a barrier separates the phases.
template <int T>
__global__ void __launch_bounds__(T) c3_rot(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  __shared__ __align__(16) uint32_t bm[3][2048];
  const int tid = threadIdx.x;
  for (int i = tid; i < 3 * 512; i += T) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const uint32_t sb0 = (uint32_t)__cvta_generic_to_shared(bm);
  unsigned total = 0;
  const int G = gridDim.x;
  const int m = (int)blockIdx.x < n_pairs ? (n_pairs - (int)blockIdx.x + G - 1) / G : 0;
  const uint4 z = make_uint4(0, 0, 0, 0);
  uint4 aS = z, bS = z, a1 = z, b1 = z, a2 = z; uint32_t laS = 0, lbS = 0, la1 = 0, lb1 = 0, la2 = 0;
  auto load = [&](int k, uint4& xa, uint4& xb, uint32_t& la, uint32_t& lb) {
    const int p = blockIdx.x + k * G;
    la = len[2 * p]; lb = len[2 * p + 1];
    if ((uint32_t)tid * 8 < la) xa = ldg_nc((const uint4*)(data + (size_t)(2 * p) * STRIDE) + tid);
    if ((uint32_t)tid * 8 < lb) xb = ldg_nc((const uint4*)(data + (size_t)(2 * p + 1) * STRIDE) + tid);
  };
  if (m > 0) load(0, aS, bS, laS, lbS);
  int slot = 0;           // j % 3
  for (int j = 0; j < m + 2; j++) {
    uint4 aN = z, bN = z; uint32_t laN = 0, lbN = 0;
    if (j + 1 < m) load(j + 1, aN, bN, laN, lbN);
    const uint32_t sS = sb0 + (uint32_t)slot * 8192u, s1 = sb0 + (uint32_t)((slot + 2) % 3) * 8192u, s2 = sb0 + (uint32_t)((slot + 1) % 3) * 8192u;
    if ((uint32_t)tid * 8 < la2) scatter8<1>(s2, a2);
    if ((uint32_t)tid * 8 < lb1) total += probe8(s1, b1, tid * 8, lb1);
    if ((uint32_t)tid * 8 < laS) scatter8<0>(sS, aS);
    __syncthreads();
    a2 = a1; la2 = la1;
    a1 = aS; b1 = bS; la1 = laS; lb1 = lbS;
    aS = aN; bS = bN; laS = laN; lbS = lbN;
    slot = slot == 2 ? 0 : slot + 1;
  }
  total = __reduce_add_sync(0xffffffff, total);
  if ((tid & 31) == 0 && total) atomicAdd(out, (unsigned long long)total);
}


// ---- w5: two-part layout.  Part 1 = one element per distinct 32-bit bitmap word (plain STS of the bit, no atomic needed on an
// all-zero bitmap), part 2 = the other elements of shared words (ATOMS.OR after part 1).  Clearing = STS 0 to the part-1 words
// (every touched word has exactly one part-1 element): no 8 KiB wipe, no atomic un-scatter.  Both parts are bank-striped and
// padded to whole chunks with copies of their last element.  len1[c] / len2[c] = elements of part 1 / 2.
__device__ __forceinline__ void sts(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
template <bool CLEAR> __device__ __forceinline__ void store8(uint32_t sb, uint4 v) {
  uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
  for (int q = 0; q < 4; q++) {
    sts(sb + off_lo(w[q]), CLEAR ? 0u : (1u << (w[q] & 31)));
    sts(sb + off_hi(w[q]), CLEAR ? 0u : (1u << ((w[q] >> 16) & 31)));
  }
}
template <int WARPS, bool PREFETCH>
__global__ void __launch_bounds__(WARPS * 32) w5_split(const uint16_t* data, const int* len1, const int* len2, int n_pairs, unsigned long long* out) {
  extern __shared__ __align__(128) uint32_t smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048; const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  unsigned total = 0;
  const int stride = gridDim.x * WARPS;
  int w = blockIdx.x * WARPS + wid;
  uint4 va[3], vb[3]; uint32_t a1 = 0, a2 = 0, b1 = 0, b2 = 0;
  auto load = [&](int p, uint4* xa, uint4* xb, uint32_t& la1, uint32_t& la2, uint32_t& lb1, uint32_t& lb2) {
    const uint4* A = (const uint4*)(data + (size_t)(2 * p) * STRIDE); const uint4* B = (const uint4*)(data + (size_t)(2 * p + 1) * STRIDE);
    la1 = len1[2 * p]; la2 = len2[2 * p]; lb1 = len1[2 * p + 1]; lb2 = len2[2 * p + 1];
    const uint32_t na8 = ((la1 + 7) >> 3) + ((la2 + 7) >> 3), nb8 = ((lb1 + 7) >> 3) + ((lb2 + 7) >> 3);
#pragma unroll
    for (int q = 0; q < 3; q++) xa[q] = ldg_nc(A + min((uint32_t)lane + 32 * q, na8 - 1));
#pragma unroll
    for (int q = 0; q < 3; q++) xb[q] = ldg_nc(B + min((uint32_t)lane + 32 * q, nb8 - 1));
  };
  if (PREFETCH && w < n_pairs) load(w, va, vb, a1, a2, b1, b2);
  for (; w < n_pairs; w += stride) {
    uint4 xa[3], xb[3]; uint32_t la1, la2, lb1, lb2;
    if (PREFETCH) {
#pragma unroll
      for (int q = 0; q < 3; q++) { xa[q] = va[q]; xb[q] = vb[q]; }
      la1 = a1; la2 = a2; lb1 = b1; lb2 = b2;
      if (w + stride < n_pairs) load(w + stride, va, vb, a1, a2, b1, b2);
    } else load(w, xa, xb, la1, la2, lb1, lb2);
    const uint32_t ca1 = (la1 + 7) >> 3, ca = ca1 + ((la2 + 7) >> 3), cb1 = (lb1 + 7) >> 3, cb = cb1 + ((lb2 + 7) >> 3);
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < ca1) store8<false>(sb, xa[q]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) { const uint32_t i = lane + 32 * q; if (i >= ca1 && i < ca) scatter8<0>(sb, xa[q]); }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const uint32_t i = lane + 32 * q;
      if (i < cb1) total += probe8(sb, xb[q], i * 8, lb1);
      else if (i < cb) total += probe8(sb, xb[q], (i - cb1) * 8, lb2);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < ca1) store8<true>(sb, xa[q]);
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}


// ---- w6: w1 (wipe) + next pair prefetched + unguarded chunks (tails padded with copies of the last element; the probe's over-count
// of a padded tail is taken out again by one extra lookup of the last element) + probe hits collected by a funnel shift.
// MULHI: the two offset computations go through IMAD.HI with a register multiplier (fma pipe) instead of LEA.HI (alu pipe).
template <bool MULHI> __device__ __forceinline__ uint32_t o_lo(uint32_t w, uint32_t m29) { return MULHI ? __umulhi(w & 0xffe0u, m29) : off_lo(w); }
template <bool MULHI> __device__ __forceinline__ uint32_t o_hi(uint32_t w, uint32_t m13) { return MULHI ? __umulhi(w & 0xffe00000u, m13) : off_hi(w); }
template <int WARPS, bool MULHI>
__global__ void __launch_bounds__(WARPS * 32) w6_tuned(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out, uint32_t m29, uint32_t m13) {
  extern __shared__ __align__(128) uint32_t smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048; const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  unsigned total = 0;
  const int stride = gridDim.x * WARPS;
  int w = blockIdx.x * WARPS + wid;
  uint4 va[3], vb[3]; uint32_t na = 0, nb = 0;
  auto load = [&](int p, uint4* xa, uint4* xb, uint32_t& la, uint32_t& lb) {
    const uint4* A = (const uint4*)(data + (size_t)(2 * p) * STRIDE); const uint4* B = (const uint4*)(data + (size_t)(2 * p + 1) * STRIDE);
    la = len[2 * p]; lb = len[2 * p + 1];
    const uint32_t na8 = (la + 7) >> 3, nb8 = (lb + 7) >> 3;
#pragma unroll
    for (int q = 0; q < 3; q++) xa[q] = ldg_nc(A + min((uint32_t)lane + 32 * q, na8 - 1));
#pragma unroll
    for (int q = 0; q < 3; q++) xb[q] = ldg_nc(B + min((uint32_t)lane + 32 * q, nb8 - 1));
  };
  if (w < n_pairs) load(w, va, vb, na, nb);
  for (; w < n_pairs; w += stride) {
    uint4 xa[3], xb[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { xa[q] = va[q]; xb[q] = vb[q]; }
    const uint32_t la = na, lb = nb;
    const uint32_t last_b = __ldg(data + (size_t)(2 * w + 1) * STRIDE + lb - 1);      // (the product keeps it in the descriptor)
    if (w + stride < n_pairs) load(w + stride, va, vb, na, nb);
    const uint32_t na8 = (la + 7) >> 3, nb8 = (lb + 7) >> 3;
#pragma unroll 4
    for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) {
      uint32_t x[4] = { xa[q].x, xa[q].y, xa[q].z, xa[q].w };
#pragma unroll
      for (int k = 0; k < 4; k++) { bit_op<0>(sb, o_lo<MULHI>(x[k], m29), x[k]); bit_op<0>(sb, o_hi<MULHI>(x[k], m13), x[k] >> 16); }
    }
    __syncwarp();
    uint32_t hits = 0;      // one bit per probed element (<= 24 per lane)
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < nb8) {
      uint32_t x[4] = { xb[q].x, xb[q].y, xb[q].z, xb[q].w };
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t t0 = lds(sb + o_lo<MULHI>(x[k], m29)) >> (x[k] & 31), t1 = lds(sb + o_hi<MULHI>(x[k], m13)) >> ((x[k] >> 16) & 31);
        hits = __funnelshift_l(t0 << 31, hits, 1);      // hits = (hits << 1) | bit
        hits = __funnelshift_l(t1 << 31, hits, 1);
      }
    }
    total += __popc(hits);
    if (lane == 0) { const uint32_t pad = (8 - (lb & 7)) & 7; total -= pad * ((lds(sb + ((last_b >> 5) << 2)) >> (last_b & 31)) & 1u); }
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}


// ---- w8: half-range passes.  Every container is stored as [values < 32768][values >= 32768], each half bank-striped for 8-byte
// pieces (4 elements per lane and round) and padded to whole pieces.  A warp intersects a pair in two passes over a 4 KiB bitmap:
// half the shared memory per warp => twice the pairs in flight per SM (the 8 KiB form is bounded to 27 warps by 227 KiB).
template <int E>
static void stripe_generic(const uint16_t* src, uint16_t* dst, uint32_t n) {      // host: as fbgpu_stripe::stripe_array, E elements per lane-chunk
  if (n == 0) return;
  std::vector<std::vector<uint16_t>> bank(32);
  for (uint32_t i = 0; i < n; i++) bank[(src[i] >> 5) & 31].push_back(src[i]);
  std::vector<size_t> pos(32, 0);
  const uint32_t per_round = 32 * E, rounds = (n + per_round - 1) / per_round;
  for (uint32_t t = 0; t < rounds; t++) for (uint32_t q = 0; q < (uint32_t)E; q++) {
    uint32_t base = per_round * t + q; if (base >= n) continue;
    uint32_t m = (n - base + E - 1) / E; if (m > 32) m = 32;
    int order[32]; for (int b = 0; b < 32; b++) order[b] = b;
    std::sort(order, order + 32, [&](int a, int b) { return bank[a].size() - pos[a] > bank[b].size() - pos[b]; });
    uint32_t lane = 0;
    while (lane < m) for (int k = 0; k < 32 && lane < m; k++) { int b = order[k]; if (pos[b] >= bank[b].size()) continue; dst[base + E * lane] = bank[b][pos[b]++]; lane++; }
  }
}
__device__ __forceinline__ uint2 ldg_nc2(const uint2* p) { uint2 r; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p)); return r; }
__device__ __forceinline__ uint32_t h_lo(uint32_t w) { return __umulhi(w & 0x7fe0u, 1u << 29); }
__device__ __forceinline__ uint32_t h_hi(uint32_t w) { return __umulhi(w & 0x7fe00000u, 1u << 13); }
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) w8_half(const uint16_t* data, const int* len_lo, const int* len_hi, int n_pairs, unsigned long long* out) {
  extern __shared__ __align__(128) uint32_t smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 1024; const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  unsigned total = 0;
  const int stride = gridDim.x * WARPS;
  for (int w = blockIdx.x * WARPS + wid; w < n_pairs; w += stride) {
    const uint16_t* A = data + (size_t)(2 * w) * STRIDE; const uint16_t* B = data + (size_t)(2 * w + 1) * STRIDE;
    const uint32_t al = len_lo[2 * w], ah = len_hi[2 * w], bl = len_lo[2 * w + 1], bh = len_hi[2 * w + 1];
    const uint32_t alp = (al + 3) & ~3u, blp = (bl + 3) & ~3u;
    uint2 xa[2][3], xb[2][3];
    const uint32_t pa[2] = { (al + 3) >> 2, (ah + 3) >> 2 }, pb[2] = { (bl + 3) >> 2, (bh + 3) >> 2 };
    const uint2* Ap[2] = { (const uint2*)A, (const uint2*)(A + alp) }; const uint2* Bp[2] = { (const uint2*)B, (const uint2*)(B + blp) };
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int q = 0; q < 3; q++) { xa[h][q] = ldg_nc2(Ap[h] + min((uint32_t)lane + 32 * q, max(pa[h], 1u) - 1)); xb[h][q] = ldg_nc2(Bp[h] + min((uint32_t)lane + 32 * q, max(pb[h], 1u) - 1)); }
    }
    const uint32_t nb_[2] = { bl, bh };
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll 4
      for (int i = lane; i < 256; i += 32) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 3; q++) if (lane + 32 * q < pa[h]) {
        const uint32_t x[2] = { xa[h][q].x, xa[h][q].y };
#pragma unroll
        for (int k = 0; k < 2; k++) { bit_op<0>(sb, h_lo(x[k]), x[k]); bit_op<0>(sb, h_hi(x[k]), x[k] >> 16); }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 3; q++) if (lane + 32 * q < pb[h]) {
        const uint32_t x[2] = { xb[h][q].x, xb[h][q].y }; const uint32_t base = (lane + 32 * q) * 4;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (base + 2 * k < nb_[h]) total += (lds(sb + h_lo(x[k])) >> (x[k] & 31)) & 1u;
          if (base + 2 * k + 1 < nb_[h]) total += (lds(sb + h_hi(x[k])) >> ((x[k] >> 16) & 31)) & 1u;
        }
      }
      __syncwarp();
    }
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}


// ---- w9: no wipe, no atomic un-scatter: the shared-memory addresses computed for the scatter stay in registers and the same
// words are set back to zero with plain stores after the probe (0 ALU work, ~1/2 of the wipe's shared-memory wavefronts).
template <int WARPS, int MINB, bool PREFETCH>
__global__ void __launch_bounds__(WARPS * 32, MINB) w9_keepaddr(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  extern __shared__ __align__(128) uint32_t smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048; const uint32_t sb = (uint32_t)__cvta_generic_to_shared(bm);
  for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  unsigned total = 0;
  const int stride = gridDim.x * WARPS;
  int w = blockIdx.x * WARPS + wid;
  uint4 va[3], vb[3]; uint32_t na = 0, nb = 0;
  auto load = [&](int p, uint4* xa, uint4* xb, uint32_t& la, uint32_t& lb) {
    const uint4* A = (const uint4*)(data + (size_t)(2 * p) * STRIDE); const uint4* B = (const uint4*)(data + (size_t)(2 * p + 1) * STRIDE);
    la = len[2 * p]; lb = len[2 * p + 1];
    const uint32_t na8 = (la + 7) >> 3, nb8 = (lb + 7) >> 3;
#pragma unroll
    for (int q = 0; q < 3; q++) xa[q] = ldg_nc(A + min((uint32_t)lane + 32 * q, na8 - 1));
#pragma unroll
    for (int q = 0; q < 3; q++) xb[q] = ldg_nc(B + min((uint32_t)lane + 32 * q, nb8 - 1));
  };
  if (PREFETCH && w < n_pairs) load(w, va, vb, na, nb);
  for (; w < n_pairs; w += stride) {
    uint4 xa[3], xb[3]; uint32_t la, lb;
    if (PREFETCH) {
#pragma unroll
      for (int q = 0; q < 3; q++) { xa[q] = va[q]; xb[q] = vb[q]; }
      la = na; lb = nb;
    } else load(w, xa, xb, la, lb);
    const uint32_t na8 = (la + 7) >> 3, nb8 = (lb + 7) >> 3;
    uint32_t addr[3][8];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const uint32_t x[4] = { xa[q].x, xa[q].y, xa[q].z, xa[q].w };
#pragma unroll
      for (int k = 0; k < 4; k++) { addr[q][2 * k] = sb + off_lo(x[k]); addr[q][2 * k + 1] = sb + off_hi(x[k]); }
      if (lane + 32 * q < na8) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr[q][2 * k]), "r"(1u << (x[k] & 31)) : "memory");
          asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr[q][2 * k + 1]), "r"(1u << ((x[k] >> 16) & 31)) : "memory");
        }
      }
    }
    if (PREFETCH && w + stride < n_pairs) load(w + stride, va, vb, na, nb);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < nb8) total += probe8(sb, xb[q], (lane + 32 * q) * 8, lb);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 3; q++) if (lane + 32 * q < na8) {
#pragma unroll
      for (int k = 0; k < 8; k++) sts(addr[q][k], 0u);
    }
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

template <typename F> float timeit(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(int argc, char** argv) {
  int n_pairs = argc > 1 ? atoi(argv[1]) : (1 << 18);
  double p = argc > 2 ? atof(argv[2]) : 0.01;
  const char* filter = argc > 3 ? argv[3] : nullptr;
  int n_cont = 2 * n_pairs;
  uint16_t *data, *sdata; int* len; unsigned long long* out;
  size_t bytes_all = (size_t)n_cont * STRIDE * 2;
  CK(cudaMalloc(&data, bytes_all)); CK(cudaMalloc(&sdata, bytes_all)); CK(cudaMalloc(&len, n_cont * 4)); CK(cudaMalloc(&out, 8));
  CK(cudaMemset(data, 0, bytes_all));
  gen<<<(n_cont + 7) / 8, 256>>>(data, len, n_cont, (uint32_t)(p * 4294967296.0), 0xFEA7B45E5EED0001ull);
  CK(cudaDeviceSynchronize());
  std::vector<int> hl(n_cont); CK(cudaMemcpy(hl.data(), len, n_cont * 4, cudaMemcpyDeviceToHost));
  std::vector<uint16_t> h(bytes_all / 2), hs(bytes_all / 2, 0);
  CK(cudaMemcpy(h.data(), data, bytes_all, cudaMemcpyDeviceToHost));
  // pad tails with the last element (both layouts), stripe a copy
  {
    int nt = std::max(1u, std::thread::hardware_concurrency()); std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] {
      for (int c = t; c < n_cont; c += nt) {
        uint16_t* src = h.data() + (size_t)c * STRIDE; uint16_t* dst = hs.data() + (size_t)c * STRIDE; int n = hl[c];
        if (n == 0) continue;
        fbgpu_stripe::stripe_array(src, dst, n);
        int padded = ((n + 7) / 8) * 8; if (padded > STRIDE) padded = STRIDE;
        fbgpu_stripe::pad_array_tail(src, n, padded); fbgpu_stripe::pad_array_tail(dst, n, padded);
      }
    });
    for (auto& x : th) x.join();
  }
  CK(cudaMemcpy(data, h.data(), bytes_all, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(sdata, hs.data(), bytes_all, cudaMemcpyHostToDevice));
  // two-part layout (w5): from the SORTED data before padding matters (h still holds n valid sorted elements + padded tail)
  std::vector<uint16_t> h2(bytes_all / 2, 0); std::vector<int> l1(n_cont), l2(n_cont);
  {
    int nt = std::max(1u, std::thread::hardware_concurrency()); std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] {
      std::vector<uint16_t> p1(4096), p2(4096);
      for (int c = t; c < n_cont; c += nt) {
        const uint16_t* src = h.data() + (size_t)c * STRIDE; uint16_t* dst = h2.data() + (size_t)c * STRIDE; int n = hl[c];
        int n1 = 0, n2 = 0; int prevw = -1;
        for (int i = 0; i < n; i++) { int wd = src[i] >> 5; if (wd != prevw) { p1[n1++] = src[i]; prevw = wd; } else p2[n2++] = src[i]; }
        int c1 = (n1 + 7) / 8 * 8, c2 = (n2 + 7) / 8 * 8;
        if (c1 + c2 > STRIDE) { n2 = std::max(0, std::min(n2, STRIDE - c1 - 8)); c2 = (n2 + 7) / 8 * 8; }
        if (n1) { fbgpu_stripe::stripe_array(p1.data(), dst, n1); fbgpu_stripe::pad_array_tail(dst, n1, c1); }
        if (n2) { fbgpu_stripe::stripe_array(p2.data(), dst + c1, n2); fbgpu_stripe::pad_array_tail(dst + c1, n2, c2); }
        l1[c] = n1; l2[c] = n2;
      }
    });
    for (auto& x : th) x.join();
  }
  std::vector<uint16_t> h3(bytes_all / 2, 0); std::vector<int> llo(n_cont), lhi(n_cont);
  {
    int nt = std::max(1u, std::thread::hardware_concurrency()); std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] {
      for (int c = t; c < n_cont; c += nt) {
        const uint16_t* src = h.data() + (size_t)c * STRIDE; uint16_t* dst = h3.data() + (size_t)c * STRIDE; int n = hl[c];
        int nlo = 0; while (nlo < n && src[nlo] < 32768) nlo++;
        int nhi = n - nlo, plo = (nlo + 3) / 4 * 4, phi = (nhi + 3) / 4 * 4;
        if (plo + phi > STRIDE) { nhi = std::max(0, STRIDE - plo - 4); phi = (nhi + 3) / 4 * 4; }
        stripe_generic<4>(src, dst, nlo); for (int k = nlo; k < plo; k++) dst[k] = dst[nlo - 1];
        stripe_generic<4>(src + nlo, dst + plo, nhi); for (int k = nhi; k < phi; k++) dst[plo + k] = dst[plo + nhi - 1];
        llo[c] = nlo; lhi[c] = nhi;
      }
    });
    for (auto& x : th) x.join();
  }
  uint16_t* data3; int *dlo, *dhi;
  CK(cudaMalloc(&data3, bytes_all)); CK(cudaMalloc(&dlo, n_cont * 4)); CK(cudaMalloc(&dhi, n_cont * 4));
  CK(cudaMemcpy(data3, h3.data(), bytes_all, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dlo, llo.data(), n_cont * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dhi, lhi.data(), n_cont * 4, cudaMemcpyHostToDevice));
  uint16_t* data2; int *dl1, *dl2;
  CK(cudaMalloc(&data2, bytes_all)); CK(cudaMalloc(&dl1, n_cont * 4)); CK(cudaMalloc(&dl2, n_cont * 4));
  CK(cudaMemcpy(data2, h2.data(), bytes_all, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dl1, l1.data(), n_cont * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dl2, l2.data(), n_cont * 4, cudaMemcpyHostToDevice));
  { double s1 = 0, s2 = 0; for (int c = 0; c < n_cont; c++) { s1 += l1[c]; s2 += l2[c]; } printf("two-part layout: part1 %.1f  part2 %.1f elements per container\n", s1 / n_cont, s2 / n_cont); }
  double elems = 0, wf_sorted = 0, wf_striped = 0; for (int x : hl) elems += x;
  for (int c = 0; c < std::min(n_cont, 4096); c++) { wf_sorted += fbgpu_stripe::total_wavefronts(h.data() + (size_t)c * STRIDE, hl[c]); wf_striped += fbgpu_stripe::total_wavefronts(hs.data() + (size_t)c * STRIDE, hl[c]); }
  double bytes = elems * 2;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int sample = std::min(n_cont, 4096); double se = 0; for (int c = 0; c < sample; c++) se += hl[c];
  printf("pairs=%d p=%g mean_len=%.1f payload=%.3f GB  wavefronts/instr: sorted %.2f striped %.2f\n", n_pairs, p, elems / n_cont, bytes / 1e9,
         wf_sorted / (se / 32.0), wf_striped / (se / 32.0));
  auto report = [&](const char* name, float ms) {
    unsigned long long hc; CK(cudaMemcpy(&hc, out, 8, cudaMemcpyDeviceToHost));
    printf("%-34s %8.3f ms  %8.1f GB/s  frac %.3f  %6.1f clk/pair/SM  count=%llu\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / 6572.0, ms * 1e-3 * 1.92e9 * sms / n_pairs, hc);
  };
#define RUN(NAME, DATA, CALL) if (!filter || strstr(NAME, filter)) { CK(cudaMemset(out, 0, 8)); const uint16_t* D = DATA; float ms = timeit([&] { CALL; }, 3); CK(cudaGetLastError()); report(NAME, ms); }
  for (int W : { 8, 9 }) for (int pf = 0; pf < 2; pf++) {
    char nm[128]; snprintf(nm, sizeof nm, "w5 split %dw x3 %s", W, pf ? "prefetch" : "direct");
    if (W == 8) { auto k = pf ? w5_split<8, true> : w5_split<8, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192); RUN(nm, data2, (k<<<sms * 3, 8 * 32, 8 * 8192>>>(D, dl1, dl2, n_pairs, out))); }
    else { auto k = pf ? w5_split<9, true> : w5_split<9, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 8192); RUN(nm, data2, (k<<<sms * 3, 9 * 32, 9 * 8192>>>(D, dl1, dl2, n_pairs, out))); }
  }
  { char nm[128]; snprintf(nm, sizeof nm, "w5 split 4w x6 prefetch"); auto k = w5_split<4, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192); RUN(nm, data2, (k<<<sms * 6, 4 * 32, 4 * 8192>>>(D, dl1, dl2, n_pairs, out))); }
  { char nm[128]; snprintf(nm, sizeof nm, "w5 split 7w x4 prefetch"); auto k = w5_split<7, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 8192); RUN(nm, data2, (k<<<sms * 4, 7 * 32, 7 * 8192>>>(D, dl1, dl2, n_pairs, out))); }
  {
    char nm[128];
    { auto k = w8_half<16, 3>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 4096);
      snprintf(nm, sizeof nm, "w8 half 16w x3"); RUN(nm, data3, (k<<<sms * 3, 16 * 32, 16 * 4096>>>(D, dlo, dhi, n_pairs, out))); }
    { auto k = w8_half<8, 6>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096);
      snprintf(nm, sizeof nm, "w8 half 8w x6"); RUN(nm, data3, (k<<<sms * 6, 8 * 32, 8 * 4096>>>(D, dlo, dhi, n_pairs, out))); }
    { auto k = w8_half<8, 5>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096);
      snprintf(nm, sizeof nm, "w8 half 8w x5"); RUN(nm, data3, (k<<<sms * 5, 8 * 32, 8 * 4096>>>(D, dlo, dhi, n_pairs, out))); }
    { auto k = w8_half<8, 4>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096);
      snprintf(nm, sizeof nm, "w8 half 8w x4"); RUN(nm, data3, (k<<<sms * 4, 8 * 32, 8 * 4096>>>(D, dlo, dhi, n_pairs, out))); }
  }
  {
    char nm[128];
    { auto k = w9_keepaddr<8, 3, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
      snprintf(nm, sizeof nm, "w9 keepaddr 8w x3 direct striped"); RUN(nm, sdata, (k<<<sms * 3, 8 * 32, 8 * 8192>>>(D, len, n_pairs, out))); }
    { auto k = w9_keepaddr<8, 3, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
      snprintf(nm, sizeof nm, "w9 keepaddr 8w x3 prefetch striped"); RUN(nm, sdata, (k<<<sms * 3, 8 * 32, 8 * 8192>>>(D, len, n_pairs, out))); }
    { auto k = w9_keepaddr<9, 3, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 8192);
      snprintf(nm, sizeof nm, "w9 keepaddr 9w x3 direct striped"); RUN(nm, sdata, (k<<<sms * 3, 9 * 32, 9 * 8192>>>(D, len, n_pairs, out))); }
    { auto k = w9_keepaddr<9, 3, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 8192);
      snprintf(nm, sizeof nm, "w9 keepaddr 9w x3 prefetch striped"); RUN(nm, sdata, (k<<<sms * 3, 9 * 32, 9 * 8192>>>(D, len, n_pairs, out))); }
    { auto k = w9_keepaddr<9, 3, false>;
      snprintf(nm, sizeof nm, "w9 keepaddr 9w x3 direct sorted"); RUN(nm, data, (k<<<sms * 3, 9 * 32, 9 * 8192>>>(D, len, n_pairs, out))); }
  }
  {
    uint32_t m29 = 1u << 29, m13 = 1u << 13; char nm[128];
    { constexpr int W = 8; auto k = w6_tuned<W, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w6 tuned 8w x3 striped"); RUN(nm, sdata, (k<<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out, m29, m13))); }
    { constexpr int W = 8; auto k = w6_tuned<W, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w6 tuned mulhi 8w x3 striped"); RUN(nm, sdata, (k<<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out, m29, m13))); }
    { constexpr int W = 9; auto k = w6_tuned<W, false>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w6 tuned 9w x3 striped"); RUN(nm, sdata, (k<<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out, m29, m13))); }
    { constexpr int W = 9; auto k = w6_tuned<W, true>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w6 tuned mulhi 9w x3 striped"); RUN(nm, sdata, (k<<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out, m29, m13))); }
  }
  for (int layout = 0; layout < 2; layout++) {
    const uint16_t* D0 = layout ? sdata : data; const char* L = layout ? "striped" : "sorted ";
    char nm[128];
    { constexpr int W = 8; cudaFuncSetAttribute(w1_wipe<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w1 wipe 8w x3 %s", L); RUN(nm, D0, (w1_wipe<W><<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out))); }
    { constexpr int W = 8; cudaFuncSetAttribute(w2_unscatter<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w2 unscatter 8w x3 %s", L); RUN(nm, D0, (w2_unscatter<W, false><<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out))); }
    { constexpr int W = 8; cudaFuncSetAttribute(w2_unscatter<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w3 unscatter+prefetch 8w x3 %s", L); RUN(nm, D0, (w2_unscatter<W, true><<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out))); }
    { constexpr int W = 9; cudaFuncSetAttribute(w2_unscatter<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w3 unscatter+prefetch 9w x3 %s", L); RUN(nm, D0, (w2_unscatter<W, true><<<sms * 3, W * 32, W * 8192>>>(D, len, n_pairs, out))); }
    { constexpr int W = 4; cudaFuncSetAttribute(w2_unscatter<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
      snprintf(nm, sizeof nm, "w3 unscatter+prefetch 4w x6 %s", L); RUN(nm, D0, (w2_unscatter<W, true><<<sms * 6, W * 32, W * 8192>>>(D, len, n_pairs, out))); }
    for (int per_sm : { 8, 12, 16, 20 }) {
      snprintf(nm, sizeof nm, "c1 coop T=128 x%d %s", per_sm, L); RUN(nm, D0, (c1_coop<128><<<sms * per_sm, 128>>>(D, len, n_pairs, out)));
    }
    for (int per_sm : { 12, 16, 21 }) {
      snprintf(nm, sizeof nm, "c1 coop T=96 x%d %s", per_sm, L); RUN(nm, D0, (c1_coop<96><<<sms * per_sm, 96>>>(D, len, n_pairs, out)));
    }
    for (int per_sm : { 4, 6, 8 }) {
      snprintf(nm, sizeof nm, "c3 rot T=128 x%d %s", per_sm, L); RUN(nm, D0, (c3_rot<128><<<sms * per_sm, 128>>>(D, len, n_pairs, out)));
    }
    for (int per_sm : { 6, 8, 9 }) {
      snprintf(nm, sizeof nm, "c3 rot T=96 x%d %s", per_sm, L); RUN(nm, D0, (c3_rot<96><<<sms * per_sm, 96>>>(D, len, n_pairs, out)));
    }
  }
  return 0;
}
