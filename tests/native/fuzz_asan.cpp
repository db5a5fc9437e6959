// AddressSanitizer / UBSan fuzz driver for the host-only readers (built and run by tests/test_native_asan.py):
//   fuzz_asan <rbf file> <roaring file> <iterations>
// Mutates the inputs (byte flips biased to headers, truncations) and runs rbf_reader.h / roaring_parse.h over them,
// touching every byte of every payload view they return.  Any out-of-bounds read aborts under ASan.
#include "rbf_reader.h"
#include "roaring_parse.h"
#include "program_compiler.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
static std::vector<uint8_t> slurp(const char* p) { std::vector<uint8_t> v; FILE* f = fopen(p, "rb"); if (!f) { perror(p); exit(2); } int c; while ((c = fgetc(f)) != EOF) v.push_back((uint8_t)c); fclose(f); return v; }
static uint64_t touch(const uint8_t* p, size_t n) { uint64_t s = 0; for (size_t i = 0; i < n; i++) s += p[i]; return s; }
int main(int argc, char** argv) {
    if (argc < 4) return 2;
    std::vector<uint8_t> rbf = slurp(argv[1]), roar = slurp(argv[2]);
    const int iters = atoi(argv[3]);
    std::mt19937_64 rng(7);
    uint64_t sink = 0, ok = 0, bad = 0;
    for (int it = 0; it < iters; it++) {
        {   // RBF: exact-size heap copy so that ASan sees the true end of the buffer
            size_t len = rbf.size();
            if (it % 5 == 4) len = rng() % (rbf.size() + 1);
            std::vector<uint8_t> b(rbf.begin(), rbf.begin() + len);
            for (int k = 0, n = (int)(rng() % 6); k < n && !b.empty(); k++) {
                size_t pg = rng() % ((b.size() + 8191) / 8192), off = pg * 8192 + (rng() % 3 == 0 ? rng() % 8192 : rng() % 64);
                if (off < b.size()) b[off] = (uint8_t)rng();
            }
            std::vector<uint8_t> wal;
            if (it % 7 == 0) { wal.assign(b.begin(), b.begin() + std::min<size_t>(b.size(), 8192 * (1 + rng() % 4))); }
            fbgpu_rbf::File f; std::string err; std::vector<fbgpu_rbf::RootRecord> recs; std::vector<fbgpu_rbf::Cell> cells;
            if (f.open(b.data(), b.size(), wal.empty() ? nullptr : wal.data(), wal.size(), err) && f.root_records(recs, err)) {
                for (auto& r : recs) {
                    cells.clear();
                    if (!f.walk(r.pgno, cells, err)) { bad++; continue; }
                    for (auto& c : cells) sink += touch(c.data, c.type == fbgpu_rbf::kCellArray ? 2ull * c.elem_n : c.type == fbgpu_rbf::kCellRLE ? 4ull * c.elem_n : 8192);
                    ok++;
                }
            } else bad++;
        }
        {   // roaring
            size_t len = roar.size();
            if (it % 4 == 3) len = rng() % (roar.size() + 1);
            std::vector<uint8_t> b(roar.begin(), roar.begin() + len);
            for (int k = 0, n = (int)(rng() % 5); k < n && !b.empty(); k++) { size_t off = rng() % std::min<size_t>(b.size(), 4096); b[off] = (uint8_t)rng(); }
            std::vector<fbgpu::ParsedCont> cs; fbgpu::Error e;
            if (fbgpu::parse_roaring(b.data(), b.size(), cs, e) == 0) {
                for (auto& c : cs) sink += touch(c.data, c.typ == fbgpu::kArray ? 2ull * c.n : c.typ == fbgpu::kBitmap ? 8192 : 4ull * c.cnt);
                ok++;
            } else bad++;
        }
    }
    for (int it = 0; it < iters * 20; it++) {          // program compiler: arbitrary (mostly malformed) post-order programs
        int n = (int)(rng() % 12);
        std::vector<fbgpu_op> ops((size_t)n);
        for (auto& o : ops) {
            o.opcode = (uint32_t)(rng() % 12); o.field = (uint32_t)(rng() % 4); o.view = (uint32_t)(rng() % 3); o.argc = (uint32_t)(rng() % 5 == 0 ? rng() : rng() % 4);
            o.a = rng() % 5 == 0 ? rng() : rng() % 70; o.b = rng() % 9; o.lo = (int64_t)rng() >> (rng() % 64); o.hi = (int64_t)rng() >> (rng() % 64);
        }
        std::vector<fbgpu::DevOp> prog; int depth = 0; fbgpu::Error e;
        fbgpu::ViewLookup lookup = [](uint32_t f, uint32_t v) { return f == 3 ? fbgpu::kNoView : f * 4 + v; };
        if (fbgpu::compile(ops.data(), n, lookup, prog, depth, e) == 0) { ok++; sink += prog.size() + (uint64_t)depth; if (depth < 1 || depth > 15) { printf("bad depth %d\n", depth); return 1; } }
        else bad++;
    }
    printf("fuzz_asan done ok=%llu rejected=%llu sink=%llu\n", (unsigned long long)ok, (unsigned long long)bad, (unsigned long long)sink);
    return 0;
}
