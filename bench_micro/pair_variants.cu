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
