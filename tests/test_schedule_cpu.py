"""Invariants of the static schedules that drive the loader / MMA issuer / epilogue roles of the fused MLP
kernels (host code in mlp_tc.cu, read through nerfb_debug_schedule; no GPU needed).  A violated invariant is a
deadlock or silent corruption on the device, so it is checked here, on the CPU."""
import ctypes as C

import pytest

from nerf_projects_b200 import _lib

F_FIRST, F_LAST, F_REL_SLOT, F_REL_ENC, F_N128 = 1, 2, 4, 8, 16


def schedule(which):
    lib = _lib.load()
    nl, no, nlay, ring = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    loads = (C.c_int * (2 * 44))()
    ops = (C.c_int * (5 * 84))()
    _lib.check(lib.nerfb_debug_schedule(which, C.byref(nl), C.byref(no), C.byref(nlay), loads, ops, C.byref(ring)),
               "nerfb_debug_schedule")
    L = [(loads[2 * i], loads[2 * i + 1]) for i in range(nl.value)]
    O = [dict(load=ops[5 * i], tile=ops[5 * i + 1], asrc=ops[5 * i + 2], flags=ops[5 * i + 3], blk=ops[5 * i + 4])
         for i in range(no.value)]
    return L, O, nlay.value, ring.value


@pytest.mark.parametrize("which,nlayers,nloads", [(0, 10, 41), (1, 8, 31), (2, 9, 34), (3, 7, 28)])
def test_schedule_invariants(which, nlayers, nloads):
    loads, ops, nlay, ring = schedule(which)
    assert nlay == nlayers and len(loads) == nloads and ring >= 4
    # blocks tile the op list exactly; a block is one tile's pass over one layer
    i, blocks = 0, []
    while i < len(ops):
        n = ops[i]["blk"]
        assert n >= 1, f"op {i} starts no block"
        assert all(o["blk"] == 0 for o in ops[i + 1:i + n]) and all(o["tile"] == ops[i]["tile"] for o in ops[i:i + n])
        assert ops[i]["flags"] & F_FIRST and not any(o["flags"] & F_FIRST for o in ops[i + 1:i + n])
        assert ops[i + n - 1]["flags"] & F_LAST and not any(o["flags"] & F_LAST for o in ops[i:i + n - 1])
        blocks.append((i, n))
        i += n
    assert i == len(ops)
    # per layer: tile 0's block then tile 1's block (ping-pong order the epilogue expects)
    assert len(blocks) == 2 * nlayers and [ops[b]["tile"] for b, _ in blocks] == [0, 1] * nlayers
    # every load is consumed by exactly the two tiles in schedule order, and released exactly once, by its LAST use
    uses = {}
    for k, o in enumerate(ops):
        uses.setdefault(o["load"], []).append(k)
    # (a chunk shared by both tiles is used twice; the encoding part of a skip/view layer is loaded once per tile)
    assert set(uses) == set(range(len(loads)))
    for ld, ks in uses.items():
        rel = [k for k in ks if ops[k]["flags"] & F_REL_SLOT]
        assert rel == [ks[-1]], f"load {ld}: released by ops {rel}, last use is {ks[-1]}"
        assert len(ks) in (1, 2)
    # ring: loads are issued in index order into slot (index % ring); when load q is issued, load q-ring must have been
    # released by an op that does not depend on q or anything later (else the loader and the issuer deadlock)
    first_use = {ld: ks[0] for ld, ks in uses.items()}
    last_use = {ld: ks[-1] for ld, ks in uses.items()}
    for q in range(ring, len(loads)):
        assert last_use[q - ring] < first_use[q], f"load {q} needs slot of load {q - ring}, which is still in use"
    # loads are used in non-decreasing order of their first use (the issuer confirms arrival in load order)
    fu = [first_use[q] for q in range(len(loads))]
    assert fu == sorted(fu)
    # the encoding buffer is released once per use; its uses alternate between the tiles
    enc_ops = [o for o in ops if o["asrc"] == 4]
    assert all(o["flags"] & F_REL_ENC for o in enc_ops) and not any(o["flags"] & F_REL_ENC for o in ops if o["asrc"] != 4)
    assert [o["tile"] for o in enc_ops] == [0, 1] * (len(enc_ops) // 2)
    # chunk sizes: 16 KB per CTA, 8 KB for the N=128 (views) layer
    for o in ops:
        assert loads[o["load"]][1] == (8192 if o["flags"] & F_N128 else 16384)
    # wrap-around: the first loads of the next iteration reuse the slots of the last loads of this one
    for q in range(ring):
        prev = len(loads) - ring + q
        assert last_use[prev] >= 0 and first_use[q] < len(ops)
