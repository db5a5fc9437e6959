// Bitmap-call program compiler (host code only; no CUDA, no store access: unit-tested on the CPU through
// tests/native/compile_check.cpp, and used by fbgpu.cu for every query).
//
// Turns the post-order pql.Call program into stack-machine device ops, mirroring executeBitmapCallShard
// (executor.go:1782-1816) and, for BSI, fragment.rangeOp's control flow (fragment.go:937-1303).  The only thing
// dropped is the data-dependent `remaining.Any()` early exit, which cannot change a result.
#pragma once
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "../../include/fbgpu.h"
#include "fbgpu_types.h"
#include "host_error.h"

namespace fbgpu {

using ViewLookup = std::function<uint32_t(uint32_t field, uint32_t view)>;
struct Node { fbgpu_op op; std::vector<int> kids; };

struct Compiler {
    const ViewLookup& lookup; Error& err;
    std::vector<DevOp> out; int depth = 0, max_depth = 0;
    Compiler(const ViewLookup& l, Error& e) : lookup(l), err(e) {}
    uint32_t fv_of(uint32_t field, uint32_t view) { return lookup(field, view); }
    void emit(uint8_t op, uint32_t fv = kNoView, uint64_t row = 0) {
        DevOp d{}; d.op = op; d.fv = fv; d.row = row; out.push_back(d);
        if (op == D_PUSH_ROW || op == D_PUSH_EMPTY) { depth++; max_depth = std::max(max_depth, depth); }
        else if (op == D_AND || op == D_OR || op == D_ANDNOT || op == D_XOR || op == D_POP) depth--;
    }
    static int bitlen(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }
    static uint64_t ones(uint64_t d) { return d >= 64 ? ~0ull : (1ull << d) - 1; }
    static uint64_t shl_ones(uint64_t d) { return d >= 64 ? 0 : ~0ull << d; }
    static uint64_t abs64(int64_t v) { return v > 0 ? (uint64_t)v : v == INT64_MIN ? 9223372036854775808ull : (uint64_t)(-v); }  // fragment.go:952

    // ---- BSI (rows: 0 exists, 1 sign, 2+i bit i; fragment.go:63-65)
    void range_eq(uint32_t fv, uint64_t depth_, int64_t pred) {                      // fragment.go:963-1003
        uint64_t up = abs64(pred);
        if ((uint64_t)bitlen(up) > depth_) { emit(D_PUSH_EMPTY); return; }
        emit(D_PUSH_ROW, fv, 0);
        emit(pred < 0 ? D_AND_ROW : D_ANDNOT_ROW, fv, 1);
        for (int i = (int)depth_ - 1; i >= 0; i--) emit(((up >> i) & 1) ? D_AND_ROW : D_ANDNOT_ROW, fv, 2 + (uint64_t)i);
    }
    void range_neq(uint32_t fv, uint64_t d, int64_t pred) { emit(D_PUSH_ROW, fv, 0); range_eq(fv, d, pred); emit(D_ANDNOT); }  // :1005-1022
    // filter on top of stack -> result on top
    void lt_unsigned(uint32_t fv, uint64_t d, uint64_t pred, bool eq) {               // :1070-1113
        if ((uint64_t)bitlen(pred) > d || (pred == ones(d) && eq)) return;
        if (pred == ones(d) && !eq) {
            emit(D_PUSH_EMPTY); emit(D_SWAP);
            for (uint64_t i = 0; i < d; i++) emit(D_ORANDNOT_ROW, fv, 2 + i);
            emit(D_POP); return;
        }
        if (eq) pred++;
        emit(D_PUSH_EMPTY); emit(D_SWAP);                                              // [matched, remaining]
        for (int i = (int)d - 1; i >= 0 && pred > 0; i--) {
            if ((pred >> i) & 1) { emit(D_ORANDNOT_ROW, fv, 2 + (uint64_t)i); pred &= ~(1ull << i); }
            else emit(D_ANDNOT_ROW, fv, 2 + (uint64_t)i);
        }
        emit(D_POP);
    }
    void gt_unsigned(uint32_t fv, uint64_t d, uint64_t pred, bool eq) {               // :1157-1205
        for (;;) {
            if (pred == 0 && eq) return;
            if (pred == 0 && !eq) {
                emit(D_PUSH_EMPTY); emit(D_SWAP);
                for (uint64_t i = 0; i < d; i++) emit(D_ORAND_ROW, fv, 2 + i);
                emit(D_POP); return;
            }
            if (!eq && (uint64_t)bitlen(pred) > d) { emit(D_POP); emit(D_PUSH_EMPTY); return; }
            if (eq) { pred--; eq = false; continue; }
            break;
        }
        emit(D_PUSH_EMPTY); emit(D_SWAP);
        pred |= shl_ones(d);
        for (int i = (int)d - 1; i >= 0 && pred < ~0ull; i--) {
            if ((pred >> i) & 1) emit(D_AND_ROW, fv, 2 + (uint64_t)i);
            else { emit(D_ORAND_ROW, fv, 2 + (uint64_t)i); pred |= 1ull << i; }
        }
        emit(D_POP);
    }
    void range_lt(uint32_t fv, uint64_t d, int64_t pred, bool eq) {                   // :1024-1067
        if (pred == 1 && !eq) { pred = 0; eq = true; }
        uint64_t up = abs64(pred);
        if (pred == 0 && !eq) { emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); }
        else if (pred == 0 && eq) { emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); range_eq(fv, d, 0); emit(D_OR); }
        else if (pred < 0) { emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); gt_unsigned(fv, d, up, eq); }
        else { emit(D_PUSH_ROW, fv, 0); emit(D_ANDNOT_ROW, fv, 1); lt_unsigned(fv, d, up, eq); emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); emit(D_OR); }
    }
    void range_gt(uint32_t fv, uint64_t d, int64_t pred, bool eq) {                   // :1115-1155
        if (pred == -1 && !eq) { pred = 0; eq = true; }
        uint64_t up = abs64(pred);
        if (pred == 0 && !eq) { range_neq(fv, d, 0); emit(D_ANDNOT_ROW, fv, 1); }
        else if (pred == 0 && eq) { emit(D_PUSH_ROW, fv, 0); emit(D_ANDNOT_ROW, fv, 1); }
        else if (pred >= 0) { emit(D_PUSH_ROW, fv, 0); emit(D_ANDNOT_ROW, fv, 1); gt_unsigned(fv, d, up, eq); }
        else { emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); lt_unsigned(fv, d, up, eq); emit(D_PUSH_ROW, fv, 0); emit(D_ANDNOT_ROW, fv, 1); emit(D_OR); }
    }
    void between_unsigned(uint32_t fv, uint64_t d, uint64_t pmin, uint64_t pmax) {     // :1262-1303
        if (pmax > ones(d)) { gt_unsigned(fv, d, pmin, true); return; }
        if (pmin == 0) { lt_unsigned(fv, d, pmax, true); return; }
        int diff = bitlen(pmax ^ pmin);
        for (int i = (int)d - 1; i >= diff; i--) emit(((pmin >> i) & 1) ? D_AND_ROW : D_ANDNOT_ROW, fv, 2 + (uint64_t)i);
        uint64_t mask = shl_ones((uint64_t)diff);
        pmin &= ~mask; pmax &= ~mask;
        gt_unsigned(fv, (uint64_t)diff, pmin, true);
        lt_unsigned(fv, (uint64_t)diff, pmax, true);
    }
    void range_between(uint32_t fv, uint64_t d, int64_t lo, int64_t hi) {             // :1213-1259
        uint64_t ulo = abs64(lo), uhi = abs64(hi);
        if (lo == hi) { range_eq(fv, d, lo); return; }
        if (lo >= 0) { emit(D_PUSH_ROW, fv, 0); emit(D_ANDNOT_ROW, fv, 1); between_unsigned(fv, d, ulo, uhi); return; }
        if (hi < 0) { emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); between_unsigned(fv, d, uhi, ulo); return; }
        emit(D_PUSH_ROW, fv, 0); emit(D_ANDNOT_ROW, fv, 1); lt_unsigned(fv, d, uhi, true);
        emit(D_PUSH_ROW, fv, 0); emit(D_AND_ROW, fv, 1); lt_unsigned(fv, d, ulo, true);
        emit(D_OR);
    }

    int gen(const std::vector<Node>& nodes, int id) {
        const Node& n = nodes[id]; const fbgpu_op& o = n.op;
        auto is_row_leaf = [&](int k) { return nodes[k].op.opcode == FBGPU_OP_ROW || nodes[k].op.opcode == FBGPU_OP_ALL; };
        auto leaf_fv = [&](int k) { return fv_of(nodes[k].op.field, nodes[k].op.view); };
        // Left fold over the children.  For the commutative/associative ops (and for the subtrahends of Difference)
        // evaluation order cannot change the result, so complex children are evaluated first and all plain Row
        // children are applied afterwards as one run of fused row ops, which the kernel executes as a barrier-free batch.
        auto fold = [&](uint8_t fused, uint8_t binop, bool first_fixed) -> int {
            std::vector<int> complex_kids, leaf_kids;
            size_t k0 = 0;
            int rc;
            if (first_fixed) { rc = gen(nodes, n.kids[0]); if (rc) return rc; k0 = 1; }
            for (size_t k = k0; k < n.kids.size(); k++) (is_row_leaf(n.kids[k]) ? leaf_kids : complex_kids).push_back(n.kids[k]);
            bool have = first_fixed;
            for (int kid : complex_kids) { rc = gen(nodes, kid); if (rc) return rc; if (have) emit(binop); have = true; }
            if (!have) {
                if (fused == D_AND_ROW) { emit(D_PUSH_ROW, leaf_fv(leaf_kids[0]), nodes[leaf_kids[0]].op.a); leaf_kids.erase(leaf_kids.begin()); }
                else emit(D_PUSH_EMPTY);       // OR / XOR onto an empty bitmap
            }
            for (int kid : leaf_kids) emit(fused, leaf_fv(kid), nodes[kid].op.a);
            return 0;
        };
        switch (o.opcode) {
            case FBGPU_OP_ROW: case FBGPU_OP_ALL: emit(D_PUSH_ROW, fv_of(o.field, o.view), o.a); return 0;
            case FBGPU_OP_EMPTY: emit(D_PUSH_EMPTY); return 0;
            case FBGPU_OP_INTERSECT:
                if (n.kids.empty()) return err.set(FBGPU_E_QUERY, "empty Intersect query is currently not supported");   // executor.go:5362
                return fold(D_AND_ROW, D_AND, false);
            case FBGPU_OP_UNION:
                if (n.kids.empty()) { emit(D_PUSH_EMPTY); return 0; }                                                  // executor.go:5386
                return fold(D_OR_ROW, D_OR, false);
            case FBGPU_OP_DIFFERENCE:
                if (n.kids.empty()) return err.set(FBGPU_E_QUERY, "empty Difference query is currently not supported");  // executor.go:2955
                return fold(D_ANDNOT_ROW, D_ANDNOT, true);
            case FBGPU_OP_XOR:
                if (n.kids.empty()) { emit(D_PUSH_EMPTY); return 0; }                                                  // executor.go:5517
                return fold(D_XOR_ROW, D_XOR, false);
            case FBGPU_OP_NOT: {                                                                                        // executor.go:5554-5602
                if (n.kids.size() != 1) return err.set(FBGPU_E_QUERY, "Not() requires a single bitmap input");
                emit(D_PUSH_ROW, fv_of(o.field, o.view), o.a);
                int kid = n.kids[0];
                if (is_row_leaf(kid)) emit(D_ANDNOT_ROW, leaf_fv(kid), nodes[kid].op.a);
                else { int rc = gen(nodes, kid); if (rc) return rc; emit(D_ANDNOT); }
                return 0;
            }
            case FBGPU_OP_BSI_RANGE: {
                uint32_t fv = fv_of(o.field, o.view); uint64_t d = o.a;
                if (d > 64) return err.set(FBGPU_E_INVALID, "bit depth %llu > 64", (unsigned long long)d);
                switch (o.b) {
                    case FBGPU_CMP_EQ: range_eq(fv, d, o.lo); break;
                    case FBGPU_CMP_NEQ: range_neq(fv, d, o.lo); break;
                    case FBGPU_CMP_LT: range_lt(fv, d, o.lo, false); break;
                    case FBGPU_CMP_LTE: range_lt(fv, d, o.lo, true); break;
                    case FBGPU_CMP_GT: range_gt(fv, d, o.lo, false); break;
                    case FBGPU_CMP_GTE: range_gt(fv, d, o.lo, true); break;
                    case FBGPU_CMP_BETWEEN: range_between(fv, d, o.lo, o.hi); break;
                    default: return err.set(FBGPU_E_INVALID, "invalid range operation %llu", (unsigned long long)o.b);   // ErrInvalidRangeOperation
                }
                return 0;
            }
        }
        return err.set(FBGPU_E_INVALID, "unknown opcode %u", o.opcode);
    }
};

// ops[0..n_ops): post-order program (include/fbgpu.h); lookup maps (field, view) to the store's view slot or kNoView.
// On success fills out / depth (operand stack depth the kernels must provide, >= 1) and returns 0; otherwise returns
// the FBGPU_E_* code with the message in err.
inline int compile(const fbgpu_op* ops, int32_t n_ops, const ViewLookup& lookup, std::vector<DevOp>& out, int& depth, Error& err) {
    if (n_ops < 0 || (n_ops > 0 && !ops)) return err.set(FBGPU_E_INVALID, "bad program");
    std::vector<Node> nodes; std::vector<int> stack;
    for (int i = 0; i < n_ops; i++) {
        Node n; n.op = ops[i];
        bool nary = n.op.opcode >= FBGPU_OP_INTERSECT && n.op.opcode <= FBGPU_OP_NOT;
        if (nary) {
            uint32_t argc = n.op.argc;
            if (argc > stack.size()) return err.set(FBGPU_E_INVALID, "op %d pops %u operands but only %zu available", i, argc, stack.size());
            n.kids.assign(stack.end() - argc, stack.end()); stack.resize(stack.size() - argc);
        }
        nodes.push_back(std::move(n)); stack.push_back((int)nodes.size() - 1);
    }
    if (stack.size() != 1) return err.set(FBGPU_E_INVALID, "program must leave exactly one result (leaves %zu)", stack.size());
    Compiler cp(lookup, err);
    int rc = cp.gen(nodes, stack[0]); if (rc) return rc;
    if (cp.max_depth > 15) return err.set(FBGPU_E_INVALID, "program needs operand stack depth %d > 15", cp.max_depth);
    out.swap(cp.out); depth = std::max(cp.max_depth, 1);
    return 0;
}

// PUSH_ROW r  ==  PUSH_EMPTY ; OR_ROW r.  Used by the word-parallel op loop (wp_machine.h), whose
// hot switch then never shifts the register stack.  The operand stack depth is unchanged.
inline void expand_push_row(std::vector<DevOp>& prog) {
    std::vector<DevOp> out; out.reserve(prog.size() + 8);
    for (const DevOp& o : prog) {
        if (o.op == D_PUSH_ROW) { DevOp e{}; e.op = D_PUSH_EMPTY; e.fv = kNoView; out.push_back(e); DevOp r = o; r.op = D_OR_ROW; out.push_back(r); }
        else out.push_back(o);
    }
    prog.swap(out);
}

}  // namespace fbgpu
