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

// Line-by-line host model of kernels.cuh:scatter_chunk_sb over a stored array payload (all n8 chunks, as batch_rows /
// warp_intersection_count issue them): MODE 0 |=, 1 &= ~, 2 ^=.  bm: 2048 words, modified in place.
static void model_bit_op(int mode, uint32_t* bm, uint32_t off, uint32_t sh) {
    const uint32_t m = 1u << (sh & 31);
    uint32_t& w = bm[off / 4];
    if (mode == 0) w |= m; else if (mode == 1) w &= ~m; else w ^= m;
}
extern "C" void model_scatter(int mode, const uint16_t* payload, uint32_t n, uint32_t* bm) {
    const uint32_t n8 = (n + 7) >> 3;
    for (uint32_t i = 0; i < n8; i++) {
        uint32_t w[4]; memcpy(w, payload + 8 * i, 16);
        const uint32_t base = i * 8;
        if (mode != 2 || base + 8 <= n) {
            for (int q = 0; q < 4; q++) { model_bit_op(mode, bm, fbgpu::word_off_lo(w[q]), w[q]); model_bit_op(mode, bm, fbgpu::word_off_hi(w[q]), w[q] >> 16); }
        } else {
            for (int q = 0; q < 4; q++) {
                if (base + 2 * q < n) model_bit_op(mode, bm, fbgpu::word_off_lo(w[q]), w[q]);
                if (base + 2 * q + 1 < n) model_bit_op(mode, bm, fbgpu::word_off_hi(w[q]), w[q] >> 16);
            }
        }
    }
}
// kernels.cuh:probe_chunk over all chunks: number of the n elements set in bm
extern "C" uint32_t model_probe(const uint16_t* payload, uint32_t n, const uint32_t* bm) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < (n + 7) >> 3; i++) {
        uint32_t w[4]; memcpy(w, payload + 8 * i, 16);
        const uint32_t base = i * 8;
        if (base + 8 <= n) {
            for (int q = 0; q < 4; q++) { c += (bm[fbgpu::word_off_lo(w[q]) / 4] >> (w[q] & 31)) & 1u; c += (bm[fbgpu::word_off_hi(w[q]) / 4] >> ((w[q] >> 16) & 31)) & 1u; }
        } else {
            for (int q = 0; q < 4; q++) {
                uint32_t lo = w[q] & 0xffffu, hi = w[q] >> 16;
                if (base + 2 * q < n) c += (bm[lo >> 5] >> (lo & 31)) & 1u;
                if (base + 2 * q + 1 < n) c += (bm[hi >> 5] >> (hi & 31)) & 1u;
            }
        }
    }
    return c;
}
