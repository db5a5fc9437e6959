// Host-only harness around featurebase_b200/csrc/program_compiler.h for tests/test_program_compiler.py (g++, no CUDA).
// View slots are synthesised as (field << 2) | view so that the Python side can map them back.
#include "program_compiler.h"
#include <cstring>
extern "C" int compile_ops(const fbgpu_op* ops, int32_t n_ops, fbgpu::DevOp* out, int32_t cap, int32_t* n_out, int32_t* depth, char* err, int32_t err_cap) {
    fbgpu::Error e; std::vector<fbgpu::DevOp> prog; int d = 0;
    fbgpu::ViewLookup lookup = [](uint32_t field, uint32_t view) { return field >= 1000 ? fbgpu::kNoView : (field << 2) | view; };
    int rc = fbgpu::compile(ops, n_ops, lookup, prog, d, e);
    if (rc) { snprintf(err, err_cap, "%s", e.msg); return rc; }
    *n_out = (int32_t)prog.size(); *depth = d;
    if ((int32_t)prog.size() > cap) return 1;
    memcpy(out, prog.data(), prog.size() * sizeof(fbgpu::DevOp));
    return 0;
}
