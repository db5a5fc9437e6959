"""Naive Python-set model of the roaring/fragment semantics (TEST INFRASTRUCTURE; same role as the
reference's own roaring/naive.go:10-309 — a dumb ground truth that the C restatement is checked against).
Only for small cases."""

SHARD_WIDTH = 1 << 20


def frag_row(frag_bits, row):
    """columns (shard-relative) of a row of a fragment given as a set of positions row*2^20+col"""
    lo, hi = row * SHARD_WIDTH, (row + 1) * SHARD_WIDTH
    return {p - lo for p in frag_bits if lo <= p < hi}


def bsi_values(frag_bits, bit_depth):
    """col -> signed value for a BSI fragment (fragment.go:63-65, 619-657)"""
    exists = frag_row(frag_bits, 0)
    sign = frag_row(frag_bits, 1)
    planes = [frag_row(frag_bits, 2 + i) for i in range(bit_depth)]
    out = {}
    for c in exists:
        mag = sum((1 << i) for i in range(bit_depth) if c in planes[i])
        out[c] = -mag if c in sign else mag
    return out


def bsi_range(frag_bits, bit_depth, op, pred, pred_max=None):
    vals = bsi_values(frag_bits, bit_depth)
    f = {
        "==": lambda v: v == pred, "!=": lambda v: v != pred, "<": lambda v: v < pred, "<=": lambda v: v <= pred,
        ">": lambda v: v > pred, ">=": lambda v: v >= pred, "><": lambda v: pred <= v <= pred_max,
    }[op]
    return {c for c, v in vals.items() if f(v)}


def optimize_type(values):
    """canonical encoding per optimize() (roaring.go:3412-3426): returns 0 (nil), 1 array, 2 bitmap, 3 run"""
    n = len(values)
    if n == 0:
        return 0
    vs = sorted(values)
    runs = 1 + sum(1 for a, b in zip(vs, vs[1:]) if b != a + 1)
    if runs <= 2048 and runs <= n // 2:
        return 3
    if n < 4096:
        return 1
    return 2
