// Micro-benchmark (design exploration, not product): issue rates of the integer instructions the array scatter / probe loops are made
// of, per SM sub-partition — IMAD.HI on the FMA pipe against LEA.HI / SHF / LOP3 on the ALU pipe, alone and mixed.
// Each warp runs 8 independent dependency chains; 4 warps per SM sub-partition (16 per SM) hide the 4-cycle latency.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
__constant__ uint32_t c_m = 0x9e3779b1u, c_m29 = 1u << 29;
constexpr int ITERS = 2048;
template <int KIND>
__global__ void __launch_bounds__(512) k(uint32_t* out, uint32_t seed, uint32_t base) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed * (threadIdx.x + 1) + i * 0x01000193u;
    uint64_t b64 = (uint64_t)base << 32;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) x[i] = __umulhi(x[i], c_m) + base;                                  // IMAD.HI (32-bit add folded? no: + IADD)
            if (KIND == 1) x[i] = (uint32_t)(((uint64_t)x[i] * c_m + b64) >> 32);              // IMAD.HI with 64-bit addend
            if (KIND == 2) x[i] = (x[i] >> 3) + base;                                           // LEA.HI / SHF+IADD
            if (KIND == 3) x[i] = (x[i] & 0xffe0u) ^ base;                                      // LOP3
            if (KIND == 4) x[i] = x[i] * c_m + base;                                            // IMAD (lo)
            if (KIND == 5) { uint32_t t = x[i] & 0xffe0ffe0u; x[i] = (uint32_t)(((uint64_t)t * c_m29 + b64) >> 32) ^ t; }   // LOP3 + IMAD.HI + LOP3 (2 ALU : 1 FMA)
            if (KIND == 6) { uint32_t t = x[i] & 0xffe0ffe0u; x[i] = ((t >> 3) + base) ^ t; }   // the same on the ALU pipe only (3 ALU)
            if (KIND == 7) x[i] = 1u << (x[i] & 31) | (x[i] >> 7);                               // SHF.L.W + SHF.R + LOP3
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int KIND> int run(const char* name, int ops_per_step, uint32_t* out, int sms) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<KIND><<<sms, 512>>>(out, 12345u, 0x4000u); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k<KIND><<<sms, 512>>>(out, 12345u, 0x4000u); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * mhz * 1e3, steps = (double)ITERS * 8 * 16 / 4;      // chain steps issued per sub-partition (16 warps / 4)
    printf("%-44s %7.3f ms  %.2f cycles per chain step per sub-partition (%d instr per step -> %.2f cycles / instr)\n", name, ms, cyc / steps, ops_per_step, cyc / steps / ops_per_step);
    return 0;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; CK(cudaMalloc(&out, (size_t)sms * 512 * 4));
    run<0>("IMAD.HI + IADD", 2, out, sms);
    run<1>("IMAD.HI with 64-bit addend", 1, out, sms);
    run<2>("shift + add (LEA.HI / SHF+IADD)", 1, out, sms);
    run<3>("LOP3", 1, out, sms);
    run<4>("IMAD (low)", 1, out, sms);
    run<5>("LOP3 + IMAD.HI + LOP3 (2 ALU : 1 FMA)", 3, out, sms);
    run<6>("LOP3 + LEA/SHF + LOP3 (3 ALU)", 3, out, sms);
    run<7>("SHF.L.W + SHF.R + LOP3", 3, out, sms);
    return 0;
}
