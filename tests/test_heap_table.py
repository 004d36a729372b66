"""Host logic behind the exact coarse tie order (csrc/coarse.cu): the structure table of the
reference's size-k result heap while it fills, and the claim the GPU's level-parallel fill rests
on -- the first J insertions can be done as 'drop d[j-1] into node entry[j], sift down', in any
order that finishes the deeper nodes first.  Checked against a literal replay of
knn_L2sqr_sse / heap_pop / heap_push (/root/reference/Auncel/utils.cpp:417-490, Heap.h:88-142)
on tie-heavy values.  Runs without a GPU: the table is computed on the host."""
import ctypes as C
import random

import numpy as np
import pytest

from auncel_b200 import _lib

NEU = (float("inf"), -1)


def entry_table(k):
    out = (C.c_int32 * (k + 1))()
    assert _lib.lib().auncel_heap_entry_table(k, out) == 0
    return list(out)


def literal_fill(k, vals):
    """heap_heapify + the push loop of knn_L2sqr_sse on a 1-based max-heap of (value, id)."""
    h = [None] + [NEU] * (k + 1)
    for j, x in enumerate(vals):
        if not x < h[1][0]:
            continue
        v, i = h[k], 1  # heap_pop
        while True:
            i1, i2 = 2 * i, 2 * i + 1
            if i1 > k:
                break
            c1, c2 = h[i1], h[i2 if i2 <= k else i1]
            left = i2 == k + 1 or c1[0] > c2[0]
            c = c1 if left else c2
            if v[0] > c[0]:
                break
            h[i] = c
            i = i1 if left else i2
        h[i] = v
        i, nv = k, (x, j)  # heap_push
        while i > 1:
            f = i >> 1
            if not nv[0] > h[f][0]:
                break
            h[i] = h[f]
            i = f
        h[i] = nv
    return h[: k + 1]


def shortcut_fill(k, vals, entry):
    """What heap_order_kernel does: parallel prefix (here: deepest level first), then the literal
    algorithm for the remaining insertions, each pop started at entry[j]."""
    J = min(entry[k], min(k, len(vals)) - 1)
    h = [None] + [NEU] * (k + 1)
    if J >= 1:
        for j in range(1, J + 1):
            h[entry[j]] = (vals[j - 1], j - 1)
        dk = k.bit_length() - 1
        for t in range(dk - 1, -1, -1):
            for i0 in range(1 << t, 1 << (t + 1)):
                v, i = h[i0], i0
                if v == NEU:
                    continue
                while True:
                    i1 = 2 * i
                    if i1 > k:
                        break
                    c1, c2 = h[i1], h[i1 + 1]
                    left = c1[0] > c2[0]
                    c = c1 if left else c2
                    if v[0] > c[0]:
                        break
                    h[i] = c
                    i = i1 + (0 if left else 1)
                h[i] = v
        h[k] = (vals[J], J)
    for j in range(J + 1 if J >= 1 else 0, len(vals)):
        x = vals[j]
        if not x < h[1][0]:
            continue
        v, i = h[k], (entry[j] if j < k else 1)
        while True:
            i1, i2 = 2 * i, 2 * i + 1
            if i1 > k:
                break
            c1, c2 = h[i1], h[i2 if i2 <= k else i1]
            left = i2 == k + 1 or c1[0] > c2[0]
            c = c1 if left else c2
            if v[0] > c[0]:
                break
            h[i] = c
            i = i1 if left else i2
        h[i] = v
        i, nv = k, (x, j)
        while i > 1:
            f = i >> 1
            if not nv[0] > h[f][0]:
                break
            h[i] = h[f]
            i = f
        h[i] = nv
    return h[: k + 1]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 8, 16, 17, 100, 255, 256, 1000, 1024])
def test_entry_table_and_parallel_prefix(k):
    entry = entry_table(k)
    assert len(entry) == k + 1 and all(1 <= e <= k for e in entry[:k])
    J = entry[k]
    assert 0 <= J <= k - 1
    if k >= 4 and (k & (k - 1)) == 0:
        # power of two: everything but the root-to-slot-k path and slot k itself
        assert J == k - k.bit_length()
    dk = k.bit_length() - 1
    for j in range(1, J + 1):  # the prefix stays off the ancestors of slot k
        n = entry[j]
        dn = n.bit_length() - 1
        assert (k >> (dk - dn)) != n
    rng = random.Random(k)
    for nv in (2, 7, 10 ** 6):  # few distinct values = many exact ties
        n = k if nv != 7 else k + rng.randint(0, 40)
        vals = [float(rng.randint(0, nv)) for _ in range(n)]
        assert literal_fill(k, vals) == shortcut_fill(k, vals, entry), (k, nv)


def heap_reorder_literal(h, k):
    """Heap.h:295-322: k pops, each top goes to the slot the pop frees."""
    h = h[:]
    for p in range(k):
        s, top, v, i = k - p, h[1], h[k - p], 1
        while True:
            i1, i2 = 2 * i, 2 * i + 1
            if i1 > s:
                break
            c1, c2 = h[i1], h[i2 if i2 <= s else i1]
            left = i2 == s + 1 or c1[0] > c2[0]
            c = c1 if left else c2
            if v[0] > c[0]:
                break
            h[i] = c
            i = i1 if left else i2
        h[i] = v
        h[s] = top
    return h


def heap_reorder_pipelined(h, k, rng, lanes=32):
    """The schedule of heap_order_kernel's second phase: pop p runs on lane p % lanes, a pop may enter
    at the root every second tick and only while no pop in flight sits on an ancestor of the slot it
    takes its value from; every lane in flight moves one level per tick.  Lanes are stepped in random
    order inside a tick: the result must not depend on it."""
    h = h[:]
    L = [dict(busy=False) for _ in range(lanes)]
    nxt, tick, last = 0, 0, -2
    while nxt < k or any(x["busy"] for x in L):
        if nxt < k and tick - last >= 2:
            s = k - nxt
            ds = s.bit_length() - 1
            hazard = any(x["busy"] and x["depth"] <= ds and (s >> (ds - x["depth"])) == x["node"] for x in L)
            if not hazard:
                x = L[nxt % lanes]
                assert not x["busy"]
                x.update(busy=True, node=1, depth=0, size=s, v=h[s], top=h[1])
                nxt += 1
                last = tick
        order = [x for x in L if x["busy"]]
        rng.shuffle(order)
        for x in order:
            i1 = 2 * x["node"]
            stop = i1 > x["size"]
            if not stop:
                c1, c2 = h[i1], h[i1 + 1 if i1 + 1 <= x["size"] else i1]
                left = i1 + 1 > x["size"] or c1[0] > c2[0]
                c = c1 if left else c2
                stop = x["v"][0] > c[0]
            if stop:
                h[x["node"]] = x["v"]
                h[x["size"]] = x["top"]
                x["busy"] = False
            else:
                h[x["node"]] = c
                x["node"] = i1 + (0 if left else 1)
                x["depth"] += 1
        tick += 1
    return h


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 31, 64, 100, 257, 1024])
def test_pipelined_heap_reorder_equals_literal(k):
    rng = random.Random(1000 + k)
    for nv in (2, 9, 10 ** 6):
        vals = [float(rng.randint(0, nv)) for _ in range(k + (rng.randint(0, 30) if nv == 9 else 0))]
        h = literal_fill(k, vals)
        assert heap_reorder_pipelined(h, k, rng, lanes=rng.choice([32, 8])) == heap_reorder_literal(h, k), (k, nv)
