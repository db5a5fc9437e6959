// Host-only harness around featurebase_b200/csrc/rbf_reader.h for tests/test_rbf.py (built with g++ into a temp dir).
// rbf_dump() writes a line-oriented text description of every bitmap and leaf cell:
//   B <name-hex> <root pgno>
//   C <key> <type> <elemN> <bitN> <payload-hex>          (bitmap cells: the 8 KiB page)
// returns the number of bytes needed (call again with a bigger buffer), or -1 with the message in out.
#include "rbf_reader.h"
#include <cstdio>
static void hex(std::string& s, const uint8_t* p, size_t n) { static const char* d = "0123456789abcdef"; for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; } }
extern "C" long long rbf_dump(const uint8_t* data, unsigned long long nd, const uint8_t* wal, unsigned long long nw, char* out, unsigned long long cap) {
    fbgpu_rbf::File f; std::string err, s;
    std::vector<fbgpu_rbf::RootRecord> recs; std::vector<fbgpu_rbf::Cell> cells;
    bool ok = f.open(data, nd, wal, nw, err) && f.root_records(recs, err);
    for (size_t i = 0; ok && i < recs.size(); i++) {
        s += "B "; hex(s, (const uint8_t*)recs[i].name.data(), recs[i].name.size()); s += " " + std::to_string(recs[i].pgno) + "\n";
        cells.clear();
        ok = f.walk(recs[i].pgno, cells, err);
        for (size_t k = 0; ok && k < cells.size(); k++) {
            const auto& c = cells[k];
            s += "C " + std::to_string(c.key) + " " + std::to_string(c.type) + " " + std::to_string(c.elem_n) + " " + std::to_string(c.bit_n) + " ";
            hex(s, c.data, c.type == fbgpu_rbf::kCellArray ? 2ull * c.elem_n : c.type == fbgpu_rbf::kCellRLE ? 4ull * c.elem_n : 8192);
            s += "\n";
        }
    }
    if (!ok) { snprintf(out, cap, "%s", err.c_str()); return -1; }
    if (s.size() + 1 <= cap) memcpy(out, s.c_str(), s.size() + 1);
    return (long long)s.size() + 1;
}
