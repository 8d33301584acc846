"""Model check of the tile-by-tile dependencies between convolution layers ("Chains of convolutions in one
launch"; sayuri_b200/csrc/conv3x3_tc2.cuh, "Cross-layer dependencies").  No GPU, no library: a restatement of

  * the buffers each block family reads and writes (EnqueueForward in sayuri_b200/csrc/engine.cu: x / t / u rotate, the
    bottleneck families use ia / ib / ic),
  * the rule by which a convolution depends on its producer tile by tile (LaunchConv: the input was written by a
    convolution that published counters, and the residual — if any — is that convolution's own input) or on everything
    before it (griddepcontrol.wait),
  * what one item touches: it READS rows of the input's tiles st-1, st, st+1 (its slab reaches 24 rows into the
    neighbours) and the residual's tile st, and WRITES the output's tile st,

and an exhaustive check, on the happens-before order those dependencies generate, that every read sees exactly the
write it is meant to see: the producing write is ordered before it, and every other write to the same tile of the same
buffer is ordered before that producing write or after the read.  This is the argument of the design in executable
form; the GPU tests (tools/chain_race.py, test_scheduling_knobs...) check the implementation against it bit for bit."""
import itertools

import pytest

TILES = 7   # super tiles per layer: enough for the st-1 .. st+1 stencil to compose over several layers


class Conv:
    def __init__(self, name, src, dst, res=None, pool=False):
        self.name, self.src, self.dst, self.res, self.pool = name, src, dst, res, pool


class Other:   # a kernel that is not a convolution: depends on everything before, everything after depends on it
    def __init__(self, name, reads, writes):
        self.name, self.reads, self.writes = name, reads, writes


def tower(block_types, se_every=0):
    """The launch sequence of EnqueueForward for a tower of the given block families."""
    ops = [Other("unpack", [], ["in"]), Conv("input", "in", "x")]
    x, t, u = "x", "t", "u"
    for b, kind in enumerate(block_types):
        se = se_every and (b + 1) % se_every == 0
        last_res = None if se else x
        skip = x
        if kind == "residual":
            ops += [Conv("b%d.c1" % b, x, t), Conv("b%d.c2" % b, t, u, last_res, pool=se)]
        elif kind == "bottleneck":
            ops += [Conv("b%d.c1" % b, x, "ia"), Conv("b%d.c2" % b, "ia", "ib"), Conv("b%d.c3" % b, "ib", "ic"),
                    Conv("b%d.c4" % b, "ic", u, last_res, pool=se)]
        elif kind == "nested":
            ops += [Conv("b%d.c1" % b, x, "ia"), Conv("b%d.c2" % b, "ia", "ib"), Conv("b%d.c3" % b, "ib", "ic", "ia"),
                    Conv("b%d.c4" % b, "ic", "ib"), Conv("b%d.c5" % b, "ib", "ia", "ic"), Conv("b%d.c6" % b, "ia", u, last_res, pool=se)]
        elif kind == "mixer":
            ops += [Other("b%d.dw" % b, [x], [t]), Conv("b%d.f1" % b, t, "ia"), Conv("b%d.f2" % b, "ia", u, None if se else t, pool=se)]
            skip = t
        if se:
            ops += [Other("b%d.se_fc" % b, ["pool"], ["gb"]), Other("b%d.se_apply" % b, [u, skip, "gb"], [u])]
        x, u = u, x
    ops += [Conv("head", x, "pv"), Other("head_fused", ["pv"], ["out"])]
    return ops


def events_and_order(ops):
    """Events (op index, tile), their reads / writes of (buffer, tile), and the direct happens-before edges."""
    n = len(ops)
    ev = [(i, st) for i in range(n) for st in range(TILES)]
    reads, writes, edges, tiled = {}, {}, set(), []
    producer = {}        # buffer -> index of the convolution that wrote it last (None after another kind of kernel)
    barrier_before = {}  # op index -> ops it waits for as whole grids
    for i, op in enumerate(ops):
        if isinstance(op, Conv):
            p = producer.get(op.src)
            tile_deps = p is not None and (op.res is None or op.res == ops[p].src)
            for st in range(TILES):
                reads[(i, st)] = [(op.src, s) for s in (st - 1, st, st + 1) if 0 <= s < TILES] + ([(op.res, st)] if op.res else [])
                writes[(i, st)] = [(op.dst, st)] + ([("pool", st)] if op.pool else [])
                if tile_deps:
                    for s in (st - 1, st, st + 1):
                        if 0 <= s < TILES:
                            edges.add(((p, s), (i, st)))
            if not tile_deps:
                barrier_before[i] = range(i)
            else:
                tiled.append(i)
            producer[op.dst] = i
        else:
            for st in range(TILES):
                reads[(i, st)] = [(b, st) for b in op.reads]
                writes[(i, st)] = [(b, st) for b in op.writes]
            barrier_before[i] = range(i)
            for b in op.writes:
                producer[b] = None
    for i, before in barrier_before.items():     # whole-grid dependency: every earlier event precedes every event of op i
        for j in before:
            for a, b in itertools.product(range(TILES), range(TILES)):
                edges.add(((j, a), (i, b)))
    return ev, reads, writes, edges, tiled


def closure(ev, edges):
    idx = {e: k for k, e in enumerate(ev)}
    n = len(ev)
    reach = [0] * n        # bitsets: reach[k] = events that happen before ev[k]
    # events are topologically ordered by (op index, tile) because every edge goes to a later op
    preds = [[] for _ in range(n)]
    for a, b in edges:
        preds[idx[b]].append(idx[a])
    for k in range(n):
        r = 0
        for p in preds[k]:
            r |= reach[p] | (1 << p)
        reach[k] = r
    return idx, reach


@pytest.mark.parametrize("blocks,se_every", [
    (["residual"] * 6, 0), (["residual"] * 7, 3), (["bottleneck"] * 3, 0), (["nested"] * 3, 2), (["mixer"] * 4, 2),
    (["residual", "nested", "mixer", "bottleneck", "residual"], 2)])
def test_every_read_sees_exactly_its_producing_write(blocks, se_every):
    ops = tower(blocks, se_every)
    ev, reads, writes, edges, tiled = events_and_order(ops)
    idx, reach = closure(ev, edges)

    def before(a, b):
        return bool(reach[idx[b]] >> idx[a] & 1)

    writers = {}
    for e in ev:
        for cell in writes[e]:
            writers.setdefault(cell, []).append(e)
    for e in ev:
        for cell in reads[e]:
            ws = [w for w in writers.get(cell, []) if w[0] < e[0]]          # program order: earlier launches
            if not ws:
                continue                                                     # external input (weights, planes)
            prod = ws[-1]
            assert before(prod, e), "%s tile %d reads %s written by %s tile %d without waiting for it" % (
                ops[e[0]].name, e[1], cell, ops[prod[0]].name, prod[1])
            for w in writers[cell]:
                if w == prod or w[0] == e[0]:
                    continue
                assert before(w, prod) or before(e, w), "%s tile %d reads %s of %s, but the write by %s tile %d is ordered with neither" % (
                    ops[e[0]].name, e[1], cell, ops[prod[0]].name, ops[w[0]].name, w[1])
    # the rule is not vacuous: a third or more of the convolutions depend on their producer tile by tile (residual towers: all but the first of a group)
    n_conv = sum(isinstance(op, Conv) for op in ops)
    assert len(tiled) >= n_conv // 3, ([ops[i].name for i in tiled], n_conv)


def test_a_too_narrow_wait_would_be_caught():
    """The same check fails when a layer waits for tile st only (the stencil the slab really reads is st-1 .. st+1)."""
    ops = tower(["residual"] * 3)
    ev, reads, writes, edges, tiled = events_and_order(ops)
    edges = {(a, b) for a, b in edges if not (b[0] in tiled and a[1] != b[1])}   # tile-dependent layers wait for tile st only
    idx, reach = closure(ev, edges)
    bad = 0
    for e in ev:
        for cell in reads[e]:
            ws = [w for w in ev if w[0] < e[0] and cell in writes[w]]
            if ws and not (reach[idx[e]] >> idx[ws[-1]] & 1):
                bad += 1
    assert bad > 0
