// Read-only walker over one shard's RBF database (the reference's roaring b-tree file, rbf/rbf.go), host code only.
// lets the residency layer take leaf cells straight from `<shard>/data` (+ `wal`) instead of asking the Go
// side to re-serialise every fragment to Pilosa roaring first.  Nothing here touches the GPU; fbgpu_load_rbf() feeds the
// containers found here to the same store builder as fbgpu_load_fragment().
//
// Format (all cites rbf/rbf.go unless noted):
//   page size 8192 (:29).  Page 0 = meta: magic "\xFFRBF" @0, pageN u32 BE @8, walID i64 BE @12, root-record pgno u32 BE
//   @20, freelist pgno u32 BE @24 (:120-141).
//   Root-record page: pgno u32 BE @0, flags u32 BE @4 (= 1), overflow pgno u32 BE @8, then records
//   { pgno u32 BE, name-length u16 BE, name } until pgno == 0 or the space runs out (:153-164, 229-256).
//   Bitmap names are "~field;view<" (short_txkey/txkey.go:129-137; index and shard are implied by the file).
//   B-tree page: pgno u32 BE @0, flags u32 BE @4 (2 leaf, 4 branch :47-53), cellN u16 BE @8, cellN x u16 BE cell offsets
//   @10 (:185-205).  Branch cell (native endian): left key u64, flags u32, child pgno u32 (:596-627).  Leaf cell (native
//   endian): key u64 @0, type u32 @8 (1 array, 2 RLE, 4 bitmap-ptr :62-71), elemN u16 @12, bitN u32 @14, payload @18:
//   array elemN x u16, RLE elemN x {start u16, last u16}, bitmap-ptr u32 pgno of a raw 8 KiB bitmap page (:489-512).
//   WAL (rbf/db.go:163-262, 316-345): a sequence of pages; a page whose flags == 8 is a bitmap header naming the page
//   number of the raw bitmap page that follows it; a page starting with the magic is a meta page and commits everything
//   before it; other pages carry their own page number.  Pages after the last meta page are an unfinished transaction.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace fbgpu_rbf {

constexpr uint64_t kPage = 8192;
constexpr uint32_t kTypeRootRecord = 1, kTypeLeaf = 2, kTypeBranch = 4, kTypeBitmapHeader = 8;
constexpr uint32_t kCellArray = 1, kCellRLE = 2, kCellBitmapPtr = 4;

struct Cell { uint64_t key; uint32_t type; uint32_t elem_n; uint32_t bit_n; const uint8_t* data; };   // data: payload (bitmap: the 8 KiB page)
struct RootRecord { std::string name; uint32_t pgno; };

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline uint32_t be16(const uint8_t* p) { return (uint32_t)p[0] << 8 | p[1]; }
inline uint64_t ne64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline uint32_t ne32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t ne16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline bool is_meta(const uint8_t* p) { return p[0] == 0xff && p[1] == 'R' && p[2] == 'B' && p[3] == 'F'; }

class File {
public:
    // data / wal are borrowed for the lifetime of the object; wal may be null / empty
    bool open(const uint8_t* data, uint64_t data_bytes, const uint8_t* wal, uint64_t wal_bytes, std::string& err) {
        data_ = data; data_pages_ = data_bytes / kPage; wal_.clear();
        if (!data || data_pages_ == 0) { err = "rbf: data file smaller than one page"; return false; }
        // number of committed WAL pages = up to and including the last meta page.  A raw bitmap page may look like
        // anything, so the scan is the reference's "methodical" one from the front (db.go:246-262).
        uint64_t wal_pages = wal ? wal_bytes / kPage : 0, committed = 0;
        for (uint64_t i = 0; i < wal_pages; i++) {
            const uint8_t* p = wal + i * kPage;
            if (is_meta(p)) committed = i + 1;
            else if (be32(p + 4) == kTypeBitmapHeader) i++;
        }
        for (uint64_t i = 0; i < committed; i++) {                 // later pages win (db.go:322-345)
            const uint8_t* p = wal + i * kPage;
            uint32_t pgno = 0;
            if (be32(p + 4) == kTypeBitmapHeader && !is_meta(p)) {
                if (i + 1 >= committed) { err = "rbf: last committed WAL page is a bitmap header"; return false; }
                pgno = be32(p); i++; p = wal + i * kPage;
            } else if (!is_meta(p)) pgno = be32(p);
            wal_[pgno] = p;
        }
        meta_ = page(0);
        if (!meta_ || !is_meta(meta_)) { err = "rbf: bad magic"; return false; }
        page_n_ = be32(meta_ + 8);
        if (page_n_ == 0) { err = "rbf: meta page count is zero"; return false; }
        return true;
    }
    uint32_t page_n() const { return page_n_; }

    // readPage rbf/tx.go:1248-1278: WAL copy first, else the data file; null when out of bounds
    const uint8_t* page(uint32_t pgno) const {
        auto it = wal_.find(pgno);
        if (it != wal_.end()) return it->second;
        if (pgno != 0 && page_n_ != 0 && pgno >= page_n_) return nullptr;
        if (pgno >= data_pages_) return nullptr;
        return data_ + (uint64_t)pgno * kPage;
    }

    bool root_records(std::vector<RootRecord>& out, std::string& err) const {
        out.clear();
        uint32_t guard = 0;
        for (uint32_t pgno = be32(meta_ + 20); pgno != 0;) {
            const uint8_t* p = page(pgno);
            if (!p) { err = "rbf: root record page " + std::to_string(pgno) + " out of bounds"; return false; }
            if (++guard > page_n_) { err = "rbf: root record page cycle"; return false; }
            uint64_t off = 12;
            while (off + 6 <= kPage) {
                uint32_t root = be32(p + off);
                if (root == 0) break;
                uint32_t len = be16(p + off + 4);
                if (off + 6 + len > kPage) { err = "rbf: short root record buffer"; return false; }
                out.push_back(RootRecord{ std::string((const char*)p + off + 6, len), root });
                off += 6 + len;
            }
            pgno = be32(p + 8);                                     // overflow page (:153)
        }
        return true;
    }

    // in-order walk of the b-tree rooted at `root`: containers in ascending key order
    bool walk(uint32_t root, std::vector<Cell>& out, std::string& err) const {
        uint64_t visited = 0;
        return walk_page(root, 0, visited, out, err);
    }

private:
    bool walk_page(uint32_t pgno, int depth, uint64_t& visited, std::vector<Cell>& out, std::string& err) const {
        const uint8_t* p = page(pgno);
        if (!p) { err = "rbf: page " + std::to_string(pgno) + " out of bounds"; return false; }
        if (depth > 32 || ++visited > (uint64_t)page_n_ + 1) { err = "rbf: b-tree cycle at page " + std::to_string(pgno); return false; }
        const uint32_t flags = be32(p + 4), n = be16(p + 8);
        if (10 + 2ull * n > kPage) { err = "rbf: cell index overruns page " + std::to_string(pgno); return false; }
        if (flags & kTypeBranch) {
            for (uint32_t i = 0; i < n; i++) {
                uint32_t off = be16(p + 10 + 2 * i);
                if (off + 16ull > kPage) { err = "rbf: branch cell overruns page " + std::to_string(pgno); return false; }
                if (!walk_page(ne32(p + off + 12), depth + 1, visited, out, err)) return false;
            }
            return true;
        }
        if (!(flags & kTypeLeaf)) { err = "rbf: page " + std::to_string(pgno) + " is neither leaf nor branch (flags " + std::to_string(flags) + ")"; return false; }
        for (uint32_t i = 0; i < n; i++) {
            uint32_t off = be16(p + 10 + 2 * i);
            if (off + 18ull > kPage) { err = "rbf: leaf cell overruns page " + std::to_string(pgno); return false; }
            Cell c{ ne64(p + off), ne32(p + off + 8), ne16(p + off + 12), ne32(p + off + 14), p + off + 18 };
            uint64_t bytes = c.type == kCellArray ? 2ull * c.elem_n : c.type == kCellRLE ? 4ull * c.elem_n : 4;
            if (c.type != kCellArray && c.type != kCellRLE && c.type != kCellBitmapPtr) { err = "rbf: invalid cell type " + std::to_string(c.type); return false; }
            if (off + 18ull + bytes > kPage) { err = "rbf: leaf cell payload overruns page " + std::to_string(pgno); return false; }
            if (c.type == kCellBitmapPtr) {
                c.data = page(ne32(p + off + 18));
                if (!c.data) { err = "rbf: bitmap page out of bounds"; return false; }
            }
            if (!out.empty() && out.back().key >= c.key) { err = "rbf: container keys out of order"; return false; }
            out.push_back(c);
        }
        return true;
    }

    const uint8_t* data_ = nullptr; uint64_t data_pages_ = 0;
    const uint8_t* meta_ = nullptr; uint32_t page_n_ = 0;
    std::unordered_map<uint32_t, const uint8_t*> wal_;
};

}  // namespace fbgpu_rbf
