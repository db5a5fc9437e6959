"""The 10 archetype container contents x 3 encodings used by the reference's TestContainerCombinations
(restated from /root/reference/roaring/roaring_helpers_test.go:12-232,257-305 as value sets)."""
import numpy as np

from oracle import oracle as O

W = 1 << 16


def archetype_values(name):
    if name == "empty":
        return np.zeros(0, dtype=np.int64)
    if name == "full":
        return np.arange(W)
    if name == "firstBitSet":
        return np.array([0])
    if name == "lastBitSet":
        return np.array([W - 1])
    if name == "firstBitUnset":
        return np.arange(1, W)
    if name == "lastBitUnset":
        return np.arange(0, W - 1)
    if name == "innerBitsSet":
        return np.arange(1, W - 1)
    if name == "outerBitsSet":
        return np.array([0, W - 1])
    if name == "oddBitsSet":
        return np.arange(1, W, 2)
    if name == "evenBitsSet":
        return np.arange(0, W, 2)
    raise KeyError(name)


NAMES = ["empty", "full", "firstBitSet", "lastBitSet", "firstBitUnset", "lastBitUnset", "innerBitsSet",
         "outerBitsSet", "oddBitsSet", "evenBitsSet"]


def container(name, typ):
    """explicit-encoding container, like doContainer() (roaring_helpers_test.go:246-257); note the reference's
    run archetypes for odd/even hold 32768 single-value runs"""
    vals = archetype_values(name)
    if typ == O.ARRAY:
        return O.Container.array(vals)
    if typ == O.BITMAP:
        w = np.zeros(1024, dtype=np.uint64)
        if len(vals):
            np.bitwise_or.at(w, vals >> 6, np.uint64(1) << (vals & 63).astype(np.uint64))
        return O.Container.bitmap(w)
    # runs
    if len(vals) == 0:
        return O.Container.run(np.zeros((0, 2), dtype=np.uint16))
    brk = np.nonzero(np.diff(vals) != 1)[0]
    starts = np.concatenate([[vals[0]], vals[brk + 1]])
    lasts = np.concatenate([vals[brk], [vals[-1]]])
    return O.Container.run(np.stack([starts, lasts], axis=1))
