"""Invariants of the static schedules that drive the loader / MMA issuer / epilogue roles of the fused MLP
kernels (host code in mlp_tc.cu, read through nerfb_debug_schedule; no GPU needed).  A violated invariant is a
deadlock or silent corruption on the device, so it is checked here, on the CPU."""
import ctypes as C

import pytest

from nerf_projects_b200 import _lib

F_FIRST, F_LAST, F_REL_SLOT, F_REL_ENC, F_N128, F_BIAS0, F_BIAS1 = 1, 2, 4, 8, 16, 32, 64


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
    # bias MMAs (forward schedules): EVERY block starts with one -- and only there -- from the bias buffer the layer's tile
    # was loaded into; buffers alternate from layer to layer (also across the wrap to the next super-tile: an even number
    # of layers); the tile rides on the layer's first load, and the first load of layer j+2 is more than `ring` loads
    # after the first load of layer j+1, so it is issued only after every MMA of layer j (see mlp_tc.cu: build_schedule)
    bias_blocks = [(b, n) for b, n in blocks if ops[b]["flags"] & (F_BIAS0 | F_BIAS1)]
    if which in (0, 1):
        assert bias_blocks == blocks and nlayers % 2 == 0
        bufs = [1 if ops[b]["flags"] & F_BIAS1 else 0 for b, _ in blocks]
        assert bufs == [(i // 2) % 2 for i in range(len(bufs))]
        for b, n in blocks:
            assert not (ops[b]["flags"] & F_BIAS0 and ops[b]["flags"] & F_BIAS1)
            assert not any(o["flags"] & (F_BIAS0 | F_BIAS1) for o in ops[b + 1:b + n])
        firsts = [min(o["load"] for o in ops[b:b + n]) for b, n in blocks[::2]]        # first load of every layer
        assert firsts == sorted(firsts) and firsts[0] == 0
        # the tile of layer j+2 overwrites the buffer of layer j: the loader issues that load (index q, possibly in the
        # next super-tile) only after the op that releases load q - ring has completed; that op must not precede the LAST
        # bias MMA of layer j, which is issued at the start of tile 1's block of layer j (a commit covers every MMA
        # issued before it, so "the same op" is enough)
        nl, nb = len(loads), len(firsts)
        for j in range(nb):
            q = firsts[(j + 2) % nb] + (nl if j + 2 >= nb else 0)
            prev = q - ring
            rel_op = last_use[prev % nl] + (len(ops) if prev >= nl else 0)
            assert rel_op >= blocks[2 * j + 1][0], f"bias buffer of layer {j} may be overwritten while in use"
    else:
        assert not bias_blocks
    # wrap-around: the first loads of the next iteration reuse the slots of the last loads of this one
    for q in range(ring):
        prev = len(loads) - ring + q
        assert last_use[prev] >= 0 and first_use[q] < len(ops)
