// Host-only harness around featurebase_b200/csrc/stripe.h for tests/test_stripe.py (built with g++ into a temp dir).
#include "stripe.h"
extern "C" void stripe(const uint16_t* src, uint16_t* dst, uint32_t n) { fbgpu_stripe::stripe_array(src, dst, n); }
extern "C" uint64_t wavefronts(const uint16_t* a, uint32_t n) { return fbgpu_stripe::total_wavefronts(a, n); }
extern "C" uint32_t worst(const uint16_t* a, uint32_t n) { return fbgpu_stripe::worst_group_conflict(a, n); }
