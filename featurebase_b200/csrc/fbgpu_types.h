// Internal device/host data structures of libfbgpu (not part of the C ABI).
//
// HBM layout of the shard store.  Everything a query needs to find a container is four
// small SoA tables plus one payload arena, all resident in HBM:
//
//   views[fv]                      (index,field,view) slot  -> slice of shardmap
//   shardmap[view.shard_off+shard] shard                    -> fragment id (or -1)
//   frags[f]                       fragment                 -> slice of rows[] (sorted by row id)
//   rows[frag.row_off + k]         row of a fragment        -> first descriptor + 16-bit slot-presence mask
//   descs[row.first_desc + rank]   container                -> payload offset, cardinality, encoding
//   rowtab[view.rt_off + shard*R + (row-rmin)]  dense (shard,row) directory when a view's row ids are dense:
//                                  shortens the chain to views -> rowtab -> descs (3 dependent loads)
//   payload[]                      array: u16[n] (16 B aligned) | bitmap: u64[1024] (128 B aligned) |
//                                  run: {u16 start,u16 last}[r] (16 B aligned)
//
// It mirrors what fragment.row() obtains through tx.OffsetRange (fragment.go:318, rbf/tx.go:1586-1637) but
// resolves with 5 dependent loads instead of a B-tree walk, and keeps every fragment of a shard contiguous.
#pragma once
#include <stdint.h>

namespace fbgpu {

constexpr int kSlotsPerRow = 16;          // ShardWidth 2^20 / 2^16 (shardwidth/helper.go:13, roaring/filter.go:24-27)
constexpr int kBitmapWords = 1024;        // roaring.go:44
constexpr uint32_t kFull = 65536;

enum : uint16_t { kArray = 1, kBitmap = 2, kRun = 3 };  // roaring.go:53-58

struct ContDesc {      // 16 B
    uint32_t off16;    // payload offset in 16-byte units
    uint32_t card;     // N, 1..65536
    uint16_t typ;      // kArray / kBitmap / kRun
    uint16_t cnt;      // run: number of intervals (array: unused, n == card)
    uint32_t pad;
};

struct RowEnt {        // 16 B
    uint64_t row;
    uint32_t first_desc;
    uint16_t mask;     // bit s set <=> container for slot s exists; descriptors are stored in slot order
    uint16_t pad;
};

struct FragHdr {       // 32 B
    uint32_t row_off;  // into rows[]
    uint32_t n_rows;
    uint64_t row0;     // first row id
    uint32_t contiguous; // 1 <=> row ids are exactly row0 .. row0+n_rows-1 (direct index, no search)
    uint32_t pad;
    uint64_t pad2;
};

struct ViewTab {       // 32 B
    uint32_t shard_off; // into shardmap[]
    uint32_t n_shards;  // shardmap slice covers shards [0, n_shards)
    uint32_t rt_rows;   // != 0: dense row table present, covering row ids [rmin, rmin + rt_rows)
    uint32_t pad;
    uint64_t rt_off;    // into rowtab[]: entry (shard, row) at rt_off + shard * rt_rows + (row - rmin)
    uint64_t rmin;
};

struct RowTabEnt {     // 8 B: the (shard,row) directory entry of the dense row table (mask == 0 => row absent in that shard)
    uint32_t first_desc;
    uint16_t mask;
    uint16_t pad;
};

struct StoreRef {      // passed to kernels by value
    const ViewTab* views;
    const int32_t* shardmap;
    const FragHdr* frags;
    const RowEnt* rows;
    const ContDesc* descs;
    const uint8_t* payload;
    const RowTabEnt* rowtab;
    uint32_t n_views;
};

// ---- compiled bitmap-call program (device ops) ----
enum : uint8_t {
    D_PUSH_ROW = 1, D_PUSH_EMPTY, D_OR_ROW, D_AND_ROW, D_ANDNOT_ROW, D_XOR_ROW,
    D_ORAND_ROW,      // S[top-1] |= S[top] &  row      (BSI: matched |= remaining ∩ row)
    D_ORANDNOT_ROW,   // S[top-1] |= S[top] & ~row      (BSI: matched |= remaining \ row)
    D_AND, D_OR, D_ANDNOT, D_XOR,  // S[top-1] = S[top-1] op S[top]; pop
    D_SWAP, D_POP
};
constexpr uint32_t kNoView = 0xffffffffu;

struct DevOp {         // 16 B
    uint8_t op;
    uint8_t pad[3];
    uint32_t fv;       // view slot (kNoView => always-empty row)
    uint64_t row;
};

struct Resolved {      // a container located for one (shard, slot); 16 B
    const void* ptr;   // nullptr => absent
    uint32_t card;
    uint16_t typ;
    uint16_t cnt;
};

}  // namespace fbgpu
