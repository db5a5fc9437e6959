// Self-test of the kernel interpreter (tests/emu/cuda_runtime.h): barriers, warp primitives, shared-memory reductions,
// dynamic shared memory, and the deadlock report for a barrier that only part of a block reaches.
#include "cuda_runtime.h"

#include <cassert>
#include <cstring>

static unsigned g_out[1024];

__global__ void scan_kernel(unsigned* out, int n) {
    unsigned* dyn = reinterpret_cast<unsigned*>(emu::g_dyn);
    __shared__ unsigned wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned v = tid < n ? (unsigned)tid + blockIdx.x : 0u, inc = v;
    for (int d = 1; d < 32; d <<= 1) { unsigned x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int k = 0; k < wid; k++) base += wsum[k];
    dyn[tid] = base + inc;                                        // inclusive prefix sum over the block
    __syncthreads();
    const unsigned total = dyn[blockDim.x - 1];
    const unsigned ball = __ballot_sync(0xffffffffu, (tid & 1) != 0);
    const unsigned red = __reduce_add_sync(0xffffffffu, 1u);
    const int all = __syncthreads_and(tid >= 0), any = __syncthreads_or(tid == 77), cnt = __syncthreads_count(tid % 3 == 0);
    // a shared-memory reduction through its 32-bit shared address
    __shared__ unsigned bits[4];
    if (tid < 4) bits[tid] = 0;
    __syncthreads();
    emu::red_or((uint32_t)__cvta_generic_to_shared(bits) + 4u * (tid & 3), 1u << (tid >> 3));
    __syncthreads();
    if (tid == 0) {
        out[blockIdx.x * 8 + 0] = total; out[blockIdx.x * 8 + 1] = ball; out[blockIdx.x * 8 + 2] = red;
        out[blockIdx.x * 8 + 3] = (unsigned)all; out[blockIdx.x * 8 + 4] = (unsigned)any; out[blockIdx.x * 8 + 5] = (unsigned)cnt;
        out[blockIdx.x * 8 + 6] = bits[0] & bits[1] & bits[2] & bits[3];
    }
    if (tid >= 40) return;                                        // early exit: the remaining threads still synchronise
    __syncthreads();
    if (tid == 0) out[blockIdx.x * 8 + 7] = __reduce_add_sync(0xffffffffu, 0u) + 7u;     // (lane 0 alone would hang if lanes 1..31 never arrived)
    else if (tid < 32) (void)__reduce_add_sync(0xffffffffu, 0u);
}

__global__ void diverge_kernel(unsigned* out) {
    // lanes 0..9 of warp 0 wait at a block barrier, lanes 10..31 at a warp barrier that needs lanes 0..9: nobody can proceed
    if (threadIdx.x < 10) __syncthreads();
    else __syncwarp();
    __syncthreads();
    out[0] = 1;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "diverge")) {
        emu::launch(dim3(1), dim3(64), 0, [&] { diverge_kernel(g_out); });
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "overrun")) {                // device buffers end at a guard page: reading past the (16-byte rounded) end faults
        unsigned char* d = nullptr;
        cudaMalloc(&d, 100);
        volatile unsigned char ok = d[111]; (void)ok;             // inside the rounded size
        puts("in bounds ok"); fflush(stdout);
        volatile unsigned char bad = d[112]; (void)bad;
        puts("not reached");
        return 0;
    }
    const int T = 256;
    emu::launch(dim3(3), dim3(T), T * 4, [&] { scan_kernel(g_out, 200); });
    for (unsigned b = 0; b < 3; b++) {
        unsigned exp = 0; for (unsigned t = 0; t < 200; t++) exp += t + b;
        assert(g_out[b * 8 + 0] == exp);
        assert(g_out[b * 8 + 1] == 0xaaaaaaaau && g_out[b * 8 + 2] == 32u);
        assert(g_out[b * 8 + 3] == 1u && g_out[b * 8 + 4] == 1u && g_out[b * 8 + 5] == 86u);
        assert(g_out[b * 8 + 6] == 0xffffffffu && g_out[b * 8 + 7] == 7u);
    }
    unsigned long long c = 5; assert(atomicAdd(&c, 3ull) == 5 && c == 8);
    unsigned m = 9; assert(atomicCAS(&m, 9u, 4u) == 9 && m == 4 && atomicMin(&m, 2u) == 4 && m == 2);
    assert(__umulhi(0xffe0u & 0x12345u, 1u << 29) == ((0x12345u & 0xffe0u) >> 3) && __ffs(8) == 4 && __popcll(~0ull) == 64 && __clz(1) == 31);
    puts("selftest ok");
    return 0;
}
