"""Inverse of ``cgb_measure_plan_build``: rebuild, from the tiles / groups / lanes the fused kernel walks, the bead
tuple every position of the values row measures (test infrastructure)."""

import numpy as np


def decode_measure_plan(plan):
    """{position in the values row: (kind, beads tuple, bond id)} of the staged columns."""
    out = {}
    lanes = plan.lanes.reshape(-1, 4, 32)
    for tile in plan.tiles:
        bead0 = int(tile["bead0"])
        bonds = plan.tile_bonds[int(tile["slot0"]):int(tile["slot0"]) + int(tile["nslots"])]
        for g in range(int(tile["group0"]), int(tile["group0"]) + int(tile["ngroups"])):
            grp = plan.groups[g]
            edges = {}     # record number -> (a, b): the record holds x_b - x_a
            for c in range(4):
                if grp["kind"][c] != 2:
                    continue
                for lane in range(int(grp["nlanes"][c])):
                    ref = int(lanes[g, c, lane]["ref"])
                    a, b = bead0 + (ref & 0xFFFF), bead0 + (ref >> 16)
                    assert a < bead0 + int(tile["span"]) and b < bead0 + int(tile["span"])
                    edges[c * 32 + lane] = (a, b)
                    if lane < int(grp["nemit"][c]):
                        slot = int(lanes[g, c, lane]["slot"])
                        out[int(grp["vcol0"][c]) + lane] = (2, (a, b), int(bonds[slot]))
                    else:
                        assert int(lanes[g, c, lane]["slot"]) == -1
            for c in range(4):
                kind = int(grp["kind"][c])
                if kind < 3:
                    continue
                assert int(grp["nemit"][c]) == int(grp["nlanes"][c])
                for lane in range(int(grp["nlanes"][c])):
                    ref = int(lanes[g, c, lane]["ref"])
                    vecs = []
                    for j in range(kind - 1):
                        byte = (ref >> (8 * j)) & 0xFF
                        a, b = edges[byte & 0x7F]
                        vecs.append((b, a) if byte & 0x80 else (a, b))   # oriented pair (from, to)
                    if kind == 3:
                        # u = x0 - x1 is the pair (1 -> 0); v = x2 - x1 is (1 -> 2)
                        (m1, x0), (m2, x2) = vecs
                        assert m1 == m2
                        beads = (x0, m1, x2)
                    else:
                        (x0, x1), (y1, x2), (y2, x3) = vecs
                        assert x1 == y1 and x2 == y2
                        beads = (x0, x1, x2, x3)
                    slot = int(lanes[g, c, lane]["slot"])
                    out[int(grp["vcol0"][c]) + lane] = (kind, beads, int(bonds[slot]))
    return out


def check_measure_plan(plan):
    """Every column of the plan is measured exactly once, at the position ``col_out`` says, from its own beads."""
    n2, n3, n4 = plan.counts
    decoded = decode_measure_plan(plan)
    assert len(decoded) == plan.n_staged and set(decoded) == set(range(plan.n_staged))
    kinds = np.concatenate([np.full(n2, 2), np.full(n3, 3), np.full(n4, 4)]).astype(int)
    positions = np.asarray(plan.col_out)
    assert sorted(positions.tolist()) == list(range(plan.n_cols))
    staged = 0
    for col in range(plan.n_cols):
        pos = int(positions[col])
        if pos >= plan.n_staged:
            continue
        staged += 1
        kind, beads, bond = decoded[pos]
        assert kind == kinds[col]
        want = tuple(int(x) for x in plan.col_beads[col][:kind])
        if kind == 2:
            assert set(beads) == set(want)       # a length does not depend on the orientation of its record
        else:
            assert beads == want
        assert bond == int(plan.col_bond[col])
    assert staged == plan.n_staged
    # chunk kinds follow the plan's shape; the per-slot lists name every emitting lane-role exactly once
    shape_kinds = {0: (2, 3, 4, 0), 1: (2, 2, 3, 3)}[int(plan.shape)]
    lanes = plan.lanes.reshape(-1, 4, 32)
    for tile in plan.tiles:
        g0, ng = int(tile["group0"]), int(tile["ngroups"])
        for g in range(g0, g0 + ng):
            for c in range(4):
                k = int(plan.groups[g]["kind"][c])
                assert k in (0, shape_kinds[c]) and (k != 0 or int(plan.groups[g]["nlanes"][c]) == 0)
        l0, nslots, nlist = int(tile["list0"]), int(tile["nslots"]), int(tile["nlist"])
        ptr = plan.tile_lists[l0:l0 + nslots + 1].astype(int)
        ent = plan.tile_lists[l0 + nslots + 1:l0 + nslots + 1 + nlist].astype(int)
        assert ptr[0] == 0 and ptr[-1] == nlist and (np.diff(ptr) >= 0).all()
        seen = set()
        for slot in range(nslots):
            for e in ent[ptr[slot]:ptr[slot + 1]]:
                c, rest = divmod(int(e), 224)
                gl, lane = divmod(rest, 32)
                assert gl < ng and int(lanes[g0 + gl, c, lane]["slot"]) == slot
                seen.add(int(e))
        want = {c * 224 + (g - g0) * 32 + lane for g in range(g0, g0 + ng) for c in range(4) for lane in range(32)
                if int(lanes[g, c, lane]["slot"]) >= 0}
        assert seen == want and len(ent) == len(want)
    # wide columns: after the staged ones, in kind order
    w = sum(plan.wide_counts)
    assert plan.n_staged + w == plan.n_cols
    return decoded
