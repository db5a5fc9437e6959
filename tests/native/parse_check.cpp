// Host-only harness around featurebase_b200/csrc/roaring_parse.h for tests/test_roaring_parse.py (g++, no CUDA).
// parse_dump(): one text line per container "key type n cnt official payload-hex" (payload as the loader would copy it:
// array n x u16, bitmap 8192 bytes, run cnt x {u16,u16}); returns bytes needed, or -(error code) with the message in out.
#include "roaring_parse.h"
#include <string>
static void hex(std::string& s, const uint8_t* p, size_t n) { static const char* d = "0123456789abcdef"; for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; } }
extern "C" long long parse_dump(const uint8_t* buf, unsigned long long len, char* out, unsigned long long cap) {
    std::vector<fbgpu::ParsedCont> cs; fbgpu::Error e;
    int rc = fbgpu::parse_roaring(buf, len, cs, e);
    if (rc) { snprintf(out, cap, "%s", e.msg); return rc; }
    std::string s;
    for (const auto& c : cs) {
        s += std::to_string(c.key) + " " + std::to_string(c.typ) + " " + std::to_string(c.n) + " " + std::to_string(c.cnt) + " " + (c.official_run ? "1 " : "0 ");
        hex(s, c.data, c.typ == fbgpu::kArray ? 2ull * c.n : c.typ == fbgpu::kBitmap ? 8192 : 4ull * c.cnt);
        s += "\n";
    }
    if (s.size() + 1 <= cap) memcpy(out, s.c_str(), s.size() + 1);
    return (long long)s.size() + 1;
}
