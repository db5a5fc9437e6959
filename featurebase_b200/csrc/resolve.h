// resolve(): locates one container in the shard store tables (fbgpu_types.h).  Host/device: the kernels inline it, and the
// inspection entry point fbgpu_debug_container() runs the very same code over the host copies of the tables, which is how
// the loaders and the table builder are tested on a box without a GPU (tests/test_store_inspect.py).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "bitaddr.h"        // FBGPU_HD
#include "fbgpu_types.h"

#if defined(__CUDA_ARCH__)
#define fbgpu_popc(x) __popc(x)
#else
#define fbgpu_popc(x) __builtin_popcount(x)
#endif

namespace fbgpu {

// Locate the container (fv, shard, row, slot).  5 dependent loads; see fbgpu_types.h.
FBGPU_HD Resolved resolve(const StoreRef& st, uint32_t fv, uint64_t shard, uint64_t row, int slot) {
    Resolved r; r.ptr = nullptr; r.card = 0; r.typ = 0; r.cnt = 0;
    if (fv >= st.n_views) return r;
    ViewTab v = st.views[fv];
    if (shard >= v.n_shards) return r;
    if (v.rt_rows) {                                  // dense (shard,row) directory: views -> rowtab -> descs
        if (row < v.rmin || row - v.rmin >= v.rt_rows) return r;
        RowTabEnt e = st.rowtab[v.rt_off + shard * v.rt_rows + (row - v.rmin)];
        if (!((e.mask >> slot) & 1)) return r;
        ContDesc d = st.descs[e.first_desc + fbgpu_popc(e.mask & ((1u << slot) - 1u))];
        r.ptr = st.payload + (size_t)d.off16 * 16; r.card = d.card; r.typ = d.typ; r.cnt = d.cnt;
        return r;
    }
    int f = st.shardmap[v.shard_off + shard];
    if (f < 0) return r;
    FragHdr h = st.frags[f];
    uint32_t idx;
    if (h.contiguous) {
        if (row < h.row0 || row - h.row0 >= h.n_rows) return r;
        idx = (uint32_t)(row - h.row0);
    } else {
        uint32_t lo = 0, hi = h.n_rows;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (st.rows[h.row_off + m].row < row) lo = m + 1; else hi = m; }
        if (lo >= h.n_rows) return r;
        idx = lo;
    }
    RowEnt e = st.rows[h.row_off + idx];
    if (e.row != row || !((e.mask >> slot) & 1)) return r;
    ContDesc d = st.descs[e.first_desc + fbgpu_popc(e.mask & ((1u << slot) - 1u))];
    r.ptr = st.payload + (size_t)d.off16 * 16; r.card = d.card; r.typ = d.typ; r.cnt = d.cnt;
    return r;
}

}  // namespace fbgpu
