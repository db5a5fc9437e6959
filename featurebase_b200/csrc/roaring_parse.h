// Reader of one fragment's serialised roaring bitmap (host code only; unit-tested on the CPU through
// tests/native/parse_check.cpp): the Pilosa format Bitmap.WriteTo emits and the official RoaringBitmap format the
// reference also accepts, with the reference's reader quirks.  Produces borrowed views into the caller's buffer.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/fbgpu.h"
#include "fbgpu_types.h"
#include "host_error.h"

namespace fbgpu {

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct ParsedCont { uint64_t key; uint16_t typ; uint32_t n; uint32_t cnt; const uint8_t* data; bool official_run; };

// Bytes after the last container are an ops log in the reference (unmarshal_binary.go:66-92: Add / Remove / batch / roaring ops
// applied on top of the containers).  The residency manager is fed snapshotted fragments, so a log is not
// replayed here; loading the containers without it would serve stale data, hence the refusal.
inline int check_no_ops_log(uint64_t end, uint64_t len, Error& err) {
    if (end < len) return err.set(FBGPU_E_FORMAT, "%llu bytes follow the last container (an ops log): snapshot the fragment before loading it", (unsigned long long)(len - end));
    return 0;
}

// Pilosa format (version = byte 2 only, byte 3 is not looked at: newPilosaRoaringIterator :1985): roaring/roaring.go:1984-2029,2124-2178; official: :1942-1980,2194-2260,6943-7006
inline int parse_roaring(const uint8_t* buf, uint64_t len, std::vector<ParsedCont>& out, Error& err) {
    out.clear();
    if (len < 8) return err.set(FBGPU_E_FORMAT, "roaring data too small (%llu bytes)", (unsigned long long)len);
    uint32_t magic = rd16(buf);
    if (magic == 12348) {
        if (buf[2] != 0) return err.set(FBGPU_E_FORMAT, "unsupported pilosa roaring version %u", buf[2]);
        uint64_t keys = rd32(buf + 4);
        if (8 + keys * 16 > len) return err.set(FBGPU_E_FORMAT, "header overruns buffer");
        const uint8_t *hdr = buf + 8, *offs = buf + 8 + keys * 12;
        uint64_t chunk = 0, end = 8 + keys * 16; uint32_t prev = 0;
        out.reserve(keys);
        for (uint64_t i = 0; i < keys; i++) {
            ParsedCont c{}; c.key = rd64(hdr + i * 12); c.typ = rd16(hdr + i * 12 + 8); c.n = (uint32_t)rd16(hdr + i * 12 + 10) + 1;
            uint32_t o32 = rd32(offs + i * 4); if (o32 < prev) chunk += 1ull << 32; prev = o32;
            uint64_t off = chunk + o32;
            if (c.typ == kArray) { if (off + (uint64_t)c.n * 2 > len) return err.set(FBGPU_E_FORMAT, "array container %llu overruns buffer", (unsigned long long)i); c.data = buf + off; }
            else if (c.typ == kBitmap) { if (off + 8192 > len) return err.set(FBGPU_E_FORMAT, "bitmap container %llu overruns buffer", (unsigned long long)i); c.data = buf + off; }
            else if (c.typ == kRun) {
                if (off + 2 > len) return err.set(FBGPU_E_FORMAT, "run container %llu overruns buffer", (unsigned long long)i);
                c.cnt = rd16(buf + off); if (off + 2 + (uint64_t)c.cnt * 4 > len) return err.set(FBGPU_E_FORMAT, "run container %llu overruns buffer", (unsigned long long)i);
                c.data = buf + off + 2;
            } else return err.set(FBGPU_E_FORMAT, "container %llu has unknown type %u", (unsigned long long)i, c.typ);
            if (!out.empty() && out.back().key >= c.key) return err.set(FBGPU_E_FORMAT, "container keys not ascending");
            end = (uint64_t)(c.data - buf) + (c.typ == kArray ? (uint64_t)c.n * 2 : c.typ == kBitmap ? 8192 : (uint64_t)c.cnt * 4);
            out.push_back(c);
        }
        return check_no_ops_log(end, len, err);
    }
    if (magic == 12346 || magic == 12347) {
        // readOfficialHeader :6960-6966: the no-run cookie is compared on all 32 bits, the run cookie on the low 16 (the high 16 hold keys - 1)
        if (magic == 12346 && rd32(buf) != 12346) return err.set(FBGPU_E_FORMAT, "did not find expected serialCookie in header");
        uint64_t keys, pos; const uint8_t* runbits = nullptr; bool have_runs = magic == 12347;
        if (have_runs) { keys = (uint64_t)rd16(buf + 2) + 1; pos = 4; runbits = buf + pos; pos += (keys + 7) / 8; if (pos > len) return err.set(FBGPU_E_FORMAT, "is-run bitmap overruns buffer"); }
        else { keys = rd32(buf + 4); pos = 8; }
        if (keys > (1u << 16)) return err.set(FBGPU_E_FORMAT, "it is logically impossible to have more than (1<<16) containers");
        if (pos + keys * 4 >= len) return err.set(FBGPU_E_FORMAT, "malformed bitmap, key-cardinality slice overruns buffer");
        const uint8_t* hdr = buf + pos; pos += keys * 4;
        const uint8_t* offs = nullptr;
        if (!have_runs) { if (pos + keys * 4 > len) return err.set(FBGPU_E_FORMAT, "insufficient data for offsets"); offs = buf + pos; }
        uint64_t cur = pos;
        for (uint64_t i = 0; i < keys; i++) {
            ParsedCont c{}; c.key = rd16(hdr + i * 4); c.n = (uint32_t)rd16(hdr + i * 4 + 2) + 1;
            bool isrun = have_runs && ((runbits[i / 8] >> (i % 8)) & 1);
            uint64_t off = offs ? rd32(offs + i * 4) : cur;
            if (isrun) {
                if (off + 2 > len) return err.set(FBGPU_E_FORMAT, "run container overruns buffer");
                c.typ = kRun; c.cnt = rd16(buf + off); if (off + 2 + (uint64_t)c.cnt * 4 > len) return err.set(FBGPU_E_FORMAT, "run container overruns buffer");
                c.data = buf + off + 2; c.official_run = true; cur = off + 2 + (uint64_t)c.cnt * 4;
            } else if (c.n < 4096) { c.typ = kArray; if (off + (uint64_t)c.n * 2 > len) return err.set(FBGPU_E_FORMAT, "array container overruns buffer"); c.data = buf + off; cur = off + (uint64_t)c.n * 2; }
            else { c.typ = kBitmap; if (off + 8192 > len) return err.set(FBGPU_E_FORMAT, "bitmap container overruns buffer"); c.data = buf + off; cur = off + 8192; }
            // the reference Put()s each container into its key map, so a repeated key would replace the earlier container and unordered
            // keys would be sorted; no writer of this format produces either, and the descriptor tables index containers by rank of
            // the key, so both are refused here
            if (!out.empty() && out.back().key >= c.key) return err.set(FBGPU_E_FORMAT, "container keys not ascending");
            out.push_back(c);
        }
        return check_no_ops_log(keys ? cur : len, len, err);
    }
    return err.set(FBGPU_E_FORMAT, "unknown roaring cookie %u", magic);
}

}  // namespace fbgpu
