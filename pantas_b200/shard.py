"""Byte-range partitioning of a GAF at line boundaries (DESIGN.md "multi-GPU").

Records are independent (the reference keeps no per-read state: ``last_read`` /
``count`` at REF:136-137,368,371 are dead), so rank r of R takes the records
that START in [b_r, b_{r+1}), where b_r is r*size/R advanced to the byte after
the next '\\n'.  Stamps carry global file offsets, so the reduced result does
not depend on R.
"""
from __future__ import annotations

import os


def _advance_to_line_start(f, pos: int, size: int) -> int:
    if pos <= 0:
        return 0
    if pos >= size:
        return size
    f.seek(pos - 1)
    while True:
        block = f.read(1 << 16)
        if not block:
            return size
        i = block.find(b"\n")
        if i >= 0:
            return f.tell() - len(block) + i + 1
        # keep scanning


def shard_bounds(path: str, world: int) -> list[int]:
    """world+1 offsets; shard r is [bounds[r], bounds[r+1])."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        b = [_advance_to_line_start(f, (size * r) // world, size) for r in range(world)]
    b.append(size)
    for r in range(1, world + 1):          # monotone even for pathological inputs
        if b[r] < b[r - 1]:
            b[r] = b[r - 1]
    return b


def shard_bounds_bytes(buf, world: int) -> list[int]:
    """Same for an in-memory buffer (bytes / numpy uint8)."""
    import numpy as np

    a = np.frombuffer(buf, dtype=np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else buf
    size = int(a.shape[0])
    out = []
    for r in range(world):
        pos = (size * r) // world
        if pos <= 0:
            out.append(0)
            continue
        nl = np.flatnonzero(a[pos - 1:] == 10)
        out.append(size if nl.size == 0 else pos - 1 + int(nl[0]) + 1)
    out.append(size)
    for r in range(1, world + 1):
        if out[r] < out[r - 1]:
            out[r] = out[r - 1]
    return out
