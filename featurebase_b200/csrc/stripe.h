// Bank-striped order for array-container payloads (EXPERIMENTAL, opt-in with FBGPU_ARRAY_STRIPED=1; unmeasured in round 1).
//
// Why: eval_kernel / pair_count_kernel turn an array container into bits of an 8 KiB shared-memory bitmap with one
// red.shared (or ld.shared probe) per element.  With the sorted order of the roaring format the 32 lanes of one such
// instruction hit pseudo-random banks (bank = (v >> 5) & 31): ~3.5 wavefronts per instruction, and the shared-memory
// pipe is the measured limiter of the headline query (profiles/README.md).  The kernels never use the ORDER of an
// array's elements (they scatter or probe them; only eval_wordpar_kernel's wp_slice searches).  So the loader may permute
// a container such that every group of elements one instruction touches has pairwise distinct banks.  Policy
// (fbgpu.cu:add_fragment_locked): only fragments dominated by array containers are striped, which leaves bitmap-heavy
// views (BSI planes) sorted and eligible for the word-parallel kernel; a view holding any striped array never takes it.
//
// Access pattern being matched (kernels.cuh: batch_rows / warp_intersection_count -> scatter_chunk_unrolled / probe_chunk):
// lane L loads the 16-byte chunk i = L + 32 t, i.e. positions 8 i .. 8 i + 7, and issues eight bit operations, the q-th
// for position 8 i + q.  Instruction group (t, q) = { 256 t + 8 L + q : L = 0..31, position < n }.
//
// Algorithm: counting sort by bank, then for every group in order take one element from each of the banks with the most
// elements left (largest-remaining-first keeps the buckets level, so conflicts only appear when a bucket is longer than
// the number of groups).  Pure host code, O(n + groups * 32).
#pragma once
#include <cstdint>
#include <cstring>

namespace fbgpu_stripe {

constexpr uint32_t kMinStripe = 64;        // shorter arrays: at most 2 lanes per instruction are active anyway

// src: n little-endian u16 (any alignment: it points into the caller's roaring file); dst: 2-byte aligned
inline void stripe_array(const void* src_bytes, uint16_t* dst, uint32_t n) {
    if (n < kMinStripe || n > 4096) { memcpy(dst, src_bytes, (size_t)n * 2); return; }
    uint16_t start[33], fill[32], left[32];
    uint16_t src[4096], tmp[4096];
    memcpy(src, src_bytes, (size_t)n * 2);
    memset(left, 0, sizeof(left));
    for (uint32_t i = 0; i < n; i++) left[(src[i] >> 5) & 31]++;
    start[0] = 0;
    for (int b = 0; b < 32; b++) { start[b + 1] = (uint16_t)(start[b] + left[b]); fill[b] = start[b]; }
    for (uint32_t i = 0; i < n; i++) tmp[fill[(src[i] >> 5) & 31]++] = src[i];      // tmp: grouped by bank, sorted inside a bank
    for (int b = 0; b < 32; b++) fill[b] = start[b];                                // fill[b]: next unread element of bank b
    const uint32_t rounds = (n + 255) >> 8;
    uint8_t order[32];
    for (int b = 0; b < 32; b++) order[b] = (uint8_t)b;
    for (uint32_t t = 0; t < rounds; t++) {
        for (uint32_t q = 0; q < 8; q++) {
            // active lanes of this group: positions 256 t + 8 L + q < n
            uint32_t base = 256 * t + q;
            if (base >= n) continue;
            uint32_t m = (n - base + 7) >> 3; if (m > 32) m = 32;
            // order the banks by elements left, descending: insertion sort over the previous group's order, which is already
            // nearly sorted (each bank lost at most one element since), so this is ~32 compares instead of ~32^2 / 4
            for (int b = 1; b < 32; b++) {
                const uint8_t x = order[b];
                int j = b;
                while (j > 0 && left[order[j - 1]] < left[x]) { order[j] = order[j - 1]; j--; }
                order[j] = x;
            }
            // one element per bank in that order; a second walk only happens when fewer than m banks are non-empty
            // (unavoidable conflicts).  Elements left == positions left >= m, so this terminates.
            uint32_t lane = 0;
            while (lane < m)
                for (int k = 0; k < 32 && lane < m; k++) {
                    const uint32_t b = order[k];
                    if (left[b] == 0) continue;
                    dst[base + 8 * lane] = tmp[fill[b]++];
                    left[b]--; lane++;
                }
        }
    }
}

// Tail padding of an array payload (payloads are stored in whole 16-byte chunks): the slots behind the n-th element repeat
// the last element instead of holding zeros.  OR-ing / AND-NOT-ing a bit twice is the same as once, so the scatter loops
// of those two modes run every chunk through the unguarded path (no divergent partial-chunk branch, which cost about a
// third of the instructions of a ~650-element container); XOR and the counting probes still honour n.
inline void pad_array_tail(uint16_t* a, uint32_t n, uint32_t padded_elems) {
    for (uint32_t k = n; k < padded_elems; k++) a[k] = a[n - 1];
}

// largest number of elements of one instruction group that share a bank (1 = conflict free); test helper
inline uint32_t worst_group_conflict(const uint16_t* a, uint32_t n) {
    uint32_t worst = 0;
    for (uint32_t t = 0; t * 256 < n; t++)
        for (uint32_t q = 0; q < 8; q++) {
            uint32_t cnt[32] = { 0 };
            for (uint32_t L = 0; L < 32; L++) { uint32_t p = 256 * t + 8 * L + q; if (p < n) { uint32_t c = ++cnt[(a[p] >> 5) & 31]; if (c > worst) worst = c; } }
        }
    return worst;
}
inline uint64_t total_wavefronts(const uint16_t* a, uint32_t n) {
    uint64_t tot = 0;
    for (uint32_t t = 0; t * 256 < n; t++)
        for (uint32_t q = 0; q < 8; q++) {
            uint32_t cnt[32] = { 0 }, mx = 0;
            for (uint32_t L = 0; L < 32; L++) { uint32_t p = 256 * t + 8 * L + q; if (p < n) { uint32_t c = ++cnt[(a[p] >> 5) & 31]; if (c > mx) mx = c; } }
            tot += mx;
        }
    return tot;
}

}  // namespace fbgpu_stripe
