// Host-only harness around featurebase_b200/csrc/stripe.h for tests/test_stripe.py (built with g++ into a temp dir).
#include "stripe.h"
extern "C" void stripe(const uint16_t* src, uint16_t* dst, uint32_t n) { fbgpu_stripe::stripe_array(src, dst, n); }
extern "C" uint64_t wavefronts(const uint16_t* a, uint32_t n) { return fbgpu_stripe::total_wavefronts(a, n); }
extern "C" uint32_t worst(const uint16_t* a, uint32_t n) { return fbgpu_stripe::worst_group_conflict(a, n); }

// featurebase_b200/csrc/bitaddr.h: exhaustive check of the mask + multiply-high word offsets (host build of the same header)
#include <initializer_list>
#include "bitaddr.h"
extern "C" uint64_t word_offset_mismatches() {
    uint64_t bad = 0;
    for (uint32_t lo = 0; lo < 65536; lo++)
        for (uint32_t hi : { 0u, 1u, 31u, 32u, 4095u, 4096u, 32768u, 65535u, lo, 65535u - lo }) {
            const uint32_t w = hi << 16 | lo;
            bad += fbgpu::word_off_lo(w) != 4 * (lo >> 5);
            bad += fbgpu::word_off_hi(w) != 4 * (hi >> 5);
        }
    return bad;
}

// what run_copies() does for one array payload: (striped or plain) copy, then the duplicate tail padding
extern "C" void load_array(const uint16_t* src, uint16_t* dst, uint32_t n, int striped) {
    if (striped) fbgpu_stripe::stripe_array(src, dst, n); else memcpy(dst, src, (size_t)n * 2);
    fbgpu_stripe::pad_array_tail(dst, n, (n + 7) & ~7u);
}
