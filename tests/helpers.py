"""Shared helpers for tests: fragment construction the way the reference's tests do it
(SetBit -> pos = row*2^20 + col%2^20, fragment.go:2780; setValue -> exists/sign/bit rows, fragment.go:619-657)."""
import numpy as np

from oracle import oracle as O

SW = 1 << 20


def set_fragments(rows_to_cols):
    """{row: [absolute columns]} -> {shard: oracle.Bitmap fragment}"""
    per = {}
    for row, cols in rows_to_cols.items():
        for c in cols:
            per.setdefault(c // SW, []).append(row * SW + (c % SW))
    return {s: O.Bitmap.from_values(v) for s, v in per.items()}


def bsi_fragment_positions(values, bit_depth):
    """{col(shard-relative): signed value} -> list of fragment positions (setValue, fragment.go:619-657)"""
    pos = []
    for col, v in values.items():
        pos.append(0 * SW + col)
        if v < 0:
            pos.append(1 * SW + col)
        mag = abs(v)
        for i in range(bit_depth):
            if (mag >> i) & 1:
                pos.append((2 + i) * SW + col)
    return pos


def bsi_fragment(values, bit_depth):
    return O.Bitmap.from_values(bsi_fragment_positions(values, bit_depth))
