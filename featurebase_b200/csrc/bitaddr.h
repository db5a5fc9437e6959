// Byte offset of the bitmap word that holds an array element, for the two u16 elements packed in one 32-bit payload word.
// Written as mask + multiply-high so that ptxas emits LOP3 + LEA.HI (base folded into the LEA) instead of
// SHF.R + LOP3 + IADD: one instruction less per scattered / probed element in the array hot loops (checked with
// cuobjdump -sass; the identity itself is checked on the host by tests/test_stripe.py::test_word_offset_identity).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FBGPU_HD __host__ __device__ __forceinline__
#else
#define FBGPU_HD inline
#endif

namespace fbgpu {
FBGPU_HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// Pipe balance (B200: the integer ALU pipe — LOP3 / SHF / LEA / IADD3 — and the FMA pipe — IMAD — each issue every second cycle per
// SM sub-partition): the scatter / probe loops are ALU-pipe bound (ncu round 2: alu 63 % busy, fma 7 %).  -DFBGPU_ADDR_IMAD moves
// the two word addresses and the `>> 16` of the upper element to IMAD.HI on the FMA pipe (multipliers read from constant memory so
// that ptxas cannot strength-reduce them, the bitmap's shared-memory base riding along as the 64-bit addend).  MEASURED SLOWER
// (bench_micro/pipe_rates.cu: IMAD.HI issues every 4.3 cycles per sub-partition, LEA.HI / SHF / IMAD-low every 2.2; headline query
// 0.381 ms against 0.352 ms), so the shift / LEA forms are the default and the IMAD form is kept only as the record of the experiment.
// w = (hi << 16) | lo.  4 * (lo >> 5)  and  4 * (hi >> 5)
FBGPU_HD uint32_t word_off_lo(uint32_t w) { return mulhi_u32(w & 0xffe0u, 1u << 29); }
FBGPU_HD uint32_t word_off_hi(uint32_t w) { return mulhi_u32(w & 0xffe00000u, 1u << 13); }
#if defined(__CUDACC__) && defined(FBGPU_ADDR_IMAD)
__constant__ uint32_t c_mul29 = 1u << 29, c_mul13 = 1u << 13, c_mul16 = 1u << 16;
#endif
#if defined(__CUDA_ARCH__) && defined(FBGPU_ADDR_IMAD)
typedef uint64_t smem_base_t;                        // shared-space address of a bitmap, kept in the upper half of a register pair
__device__ __forceinline__ smem_base_t smem_base(uint32_t sb) { return (uint64_t)sb << 32; }
__device__ __forceinline__ uint32_t word_addr_lo(smem_base_t b, uint32_t w) { return (uint32_t)(((uint64_t)(w & 0xffe0u) * c_mul29 + b) >> 32); }
__device__ __forceinline__ uint32_t word_addr_hi(smem_base_t b, uint32_t w) { return (uint32_t)(((uint64_t)(w & 0xffe00000u) * c_mul13 + b) >> 32); }
__device__ __forceinline__ uint32_t upper16(uint32_t w) { return __umulhi(w, c_mul16); }
__device__ __forceinline__ void pin_base(smem_base_t& b) { asm volatile("" : "+l"(b)); }
#else
typedef uint32_t smem_base_t;
FBGPU_HD smem_base_t smem_base(uint32_t sb) { return sb; }
FBGPU_HD uint32_t word_addr_lo(smem_base_t b, uint32_t w) { return b + word_off_lo(w); }
FBGPU_HD uint32_t word_addr_hi(smem_base_t b, uint32_t w) { return b + word_off_hi(w); }
FBGPU_HD uint32_t upper16(uint32_t w) { return w >> 16; }
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void pin_base(smem_base_t& b) { asm volatile("" : "+r"(b)); }
#else
inline void pin_base(smem_base_t&) {}
#endif
#endif
}  // namespace fbgpu
