// Micro-benchmark (design exploration, not product): array x array intersection-count
// variants for ~1% density containers. One warp per container pair.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int STRIDE = 832;   // u16 slots per container (1664 B, 128B aligned)

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// one warp per container: Bernoulli(p) per value
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

// V0: just stream the data (sum) - memory floor for this layout
__global__ void v0_stream(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_pairs) return;
  unsigned acc = 0;
  for (int s = 0; s < 2; s++) {
    const uint4* p = (const uint4*)(data + (size_t)(2 * w + s) * STRIDE);
    int n16 = (len[2 * w + s] + 7) >> 3;
    for (int i = lane; i < n16; i += 32) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  }
  acc = __reduce_add_sync(0xffffffff, acc);
  if (lane == 0 && acc == 0x12345) atomicAdd(out, 1ull);
}

// V1: warp-private 8KiB bitmap, clear + atomicOr + probe
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) v1_atomic(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  extern __shared__ uint32_t smem[];
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048;
  unsigned total = 0;
  for (int w = blockIdx.x * WARPS + wid; w < n_pairs; w += gridDim.x * WARPS) {
    const uint16_t* A = data + (size_t)(2 * w) * STRIDE; const uint16_t* B = A + STRIDE;
    int na = len[2 * w], nb = len[2 * w + 1];
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = z;
    __syncwarp();
    for (int i = lane; i < na; i += 32) { uint32_t v = A[i]; atomicOr(&bm[v >> 5], 1u << (v & 31)); }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) { uint32_t v = B[i]; total += (bm[v >> 5] >> (v & 31)) & 1; }
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

// V2: tag cells (tag16|bits16), 4096 cells per warp, STS-only build with shuffle de-dup
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) v2_tag(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  extern __shared__ uint32_t smem[];
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* cells = smem + wid * 4096;
  for (int i = lane; i < 4096; i += 32) cells[i] = 0;
  __syncwarp();
  unsigned total = 0; uint32_t tag = 0;
  for (int w = blockIdx.x * WARPS + wid; w < n_pairs; w += gridDim.x * WARPS) {
    tag++;
    if (tag == 65536) { for (int i = lane; i < 4096; i += 32) cells[i] = 0; tag = 1; __syncwarp(); }
    const uint16_t* A = data + (size_t)(2 * w) * STRIDE; const uint16_t* B = A + STRIDE;
    int na = len[2 * w], nb = len[2 * w + 1];
    uint32_t ccell = 0xffffffffu, cbits = 0;
    for (int base = 0; base < na; base += 32) {
      int i = base + lane; bool valid = i < na;
      uint32_t v = valid ? A[i] : 0;
      uint32_t cell = valid ? (v >> 4) : (0x80000000u + lane);
      uint32_t bits = 1u << (v & 15);
      if (lane == 0 && cell == ccell) bits |= cbits;
      #pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        uint32_t ub = __shfl_up_sync(0xffffffff, bits, d);
        uint32_t uc = __shfl_up_sync(0xffffffff, cell, d);
        if (lane >= d && uc == cell) bits |= ub;
      }
      uint32_t nc = __shfl_down_sync(0xffffffff, cell, 1);
      bool tail = valid && (lane == 31 || nc != cell);
      if (tail) cells[cell] = (tag << 16) | bits;
      ccell = __shfl_sync(0xffffffff, cell, 31); cbits = __shfl_sync(0xffffffff, bits, 31);
    }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {
      uint32_t v = B[i]; uint32_t c = cells[v >> 4];
      total += ((c >> 16) == tag) & ((c >> (v & 15)) & 1);
    }
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

// V3: per-lane binary search of B elements into A (global / L1)
__global__ void v3_bsearch(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_pairs) return;
  const uint16_t* A = data + (size_t)(2 * w) * STRIDE; const uint16_t* B = A + STRIDE;
  int na = len[2 * w], nb = len[2 * w + 1];
  unsigned total = 0;
  for (int i = lane; i < nb; i += 32) {
    uint32_t v = B[i]; int lo = 0, hi = na;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (A[mid] < v) lo = mid + 1; else hi = mid; }
    total += (lo < na && A[lo] == v);
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

// V4: warp-private bitmap, non-atomic RMW with de-dup (u32 words), explicit clear
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) v4_rmw(const uint16_t* data, const int* len, int n_pairs, unsigned long long* out) {
  extern __shared__ uint32_t smem[];
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* bm = smem + wid * 2048;
  unsigned total = 0;
  for (int w = blockIdx.x * WARPS + wid; w < n_pairs; w += gridDim.x * WARPS) {
    const uint16_t* A = data + (size_t)(2 * w) * STRIDE; const uint16_t* B = A + STRIDE;
    int na = len[2 * w], nb = len[2 * w + 1];
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 512; i += 32) ((uint4*)bm)[i] = z;
    __syncwarp();
    for (int base = 0; base < na; base += 32) {
      int i = base + lane; bool valid = i < na;
      uint32_t v = valid ? A[i] : 0;
      uint32_t cell = valid ? (v >> 5) : (0x80000000u + lane);
      uint32_t bits = 1u << (v & 31);
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t ub = __shfl_up_sync(0xffffffff, bits, d);
        uint32_t uc = __shfl_up_sync(0xffffffff, cell, d);
        if (lane >= d && uc == cell) bits |= ub;
      }
      uint32_t nc = __shfl_down_sync(0xffffffff, cell, 1);
      bool tail = valid && (lane == 31 || nc != cell);
      if (tail) bm[cell] |= bits;
      __syncwarp();
    }
    for (int i = lane; i < nb; i += 32) { uint32_t v = B[i]; total += (bm[v >> 5] >> (v & 31)) & 1; }
    __syncwarp();
  }
  total = __reduce_add_sync(0xffffffff, total);
  if (lane == 0) atomicAdd(out, (unsigned long long)total);
}

// raw smem atomic throughput: every lane atomically ORs into random words of an 8KiB CTA bitmap
__global__ void raw_atoms(unsigned long long* out, int iters, int mode) {
  __shared__ uint32_t bm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) bm[i] = 0;
  __syncthreads();
  uint64_t s = splitmix(blockIdx.x * 1024 + threadIdx.x);
  unsigned acc = 0;
  for (int i = 0; i < iters; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t v = (uint32_t)(s >> 40) & 0xffff;
    if (mode == 0) atomicOr(&bm[v >> 5], 1u << (v & 31));
    else if (mode == 1) acc += bm[v >> 5];
    else bm[v >> 5] = v;
  }
  __syncthreads();
  if (acc == 0x1234567 || bm[threadIdx.x] == 0xdeadbeef) atomicAdd(out, 1ull);
}

template <typename F> float timeit(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(int argc, char** argv) {
  int n_pairs = argc > 1 ? atoi(argv[1]) : (1 << 20);
  double p = argc > 2 ? atof(argv[2]) : 0.01;
  int n_cont = 2 * n_pairs;
  uint16_t* data; int* len; unsigned long long* out;
  CK(cudaMalloc(&data, (size_t)n_cont * STRIDE * 2)); CK(cudaMalloc(&len, n_cont * 4)); CK(cudaMalloc(&out, 8));
  CK(cudaMemset(data, 0, (size_t)n_cont * STRIDE * 2));
  gen<<<(n_cont + 7) / 8, 256>>>(data, len, n_cont, (uint32_t)(p * 4294967296.0), 0xFEA7B45E5EED0001ull);
  CK(cudaDeviceSynchronize());
  std::vector<int> hl(n_cont); CK(cudaMemcpy(hl.data(), len, n_cont * 4, cudaMemcpyDeviceToHost));
  double elems = 0; for (int x : hl) elems += x;
  double bytes = elems * 2;
  printf("pairs=%d p=%g mean_len=%.1f payload=%.3f GB\n", n_pairs, p, elems / n_cont, bytes / 1e9);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto report = [&](const char* name, float ms) {
    unsigned long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("%-28s %8.3f ms  %8.1f GB/s  %7.2f Gelem/s  count=%llu\n", name, ms, bytes / ms / 1e6, elems / ms / 1e6, h);
  };
  { CK(cudaMemset(out, 0, 8)); float ms = timeit([&] { v0_stream<<<(n_pairs + 7) / 8, 256>>>(data, len, n_pairs, out); }, 5); report("v0_stream", ms); }
  { CK(cudaMemset(out, 0, 8)); constexpr int W = 8; cudaFuncSetAttribute(v1_atomic<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
    float ms = timeit([&] { v1_atomic<W><<<sms * 3, W * 32, W * 8192>>>(data, len, n_pairs, out); }, 3); report("v1_atomic (8w,3cta)", ms); }
  { CK(cudaMemset(out, 0, 8)); constexpr int W = 4; cudaFuncSetAttribute(v2_tag<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 16384);
    float ms = timeit([&] { v2_tag<W><<<sms * 3, W * 32, W * 16384>>>(data, len, n_pairs, out); }, 3); report("v2_tag (4w,3cta)", ms); }
  { CK(cudaMemset(out, 0, 8)); float ms = timeit([&] { v3_bsearch<<<(n_pairs + 7) / 8, 256>>>(data, len, n_pairs, out); }, 3); report("v3_bsearch", ms); }
  { CK(cudaMemset(out, 0, 8)); constexpr int W = 8; cudaFuncSetAttribute(v4_rmw<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 8192);
    float ms = timeit([&] { v4_rmw<W><<<sms * 3, W * 32, W * 8192>>>(data, len, n_pairs, out); }, 3); report("v4_rmw (8w,3cta)", ms); }
  for (int mode = 0; mode < 3; mode++) {
    int iters = 4096; float ms = timeit([&] { raw_atoms<<<sms * 4, 256>>>(out, iters, mode); }, 3);
    double ops = (double)sms * 4 * 256 * iters;
    printf("raw smem %s: %.3f ms, %.2f lane-ops/cycle/SM (at 1.9GHz)\n", mode == 0 ? "atomicOr" : mode == 1 ? "LDS" : "STS", ms, ops / (ms * 1e-3) / 1.9e9 / sms);
  }
  return 0;
}
