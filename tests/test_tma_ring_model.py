"""CPU model of the residual-box ring of the direct kernel's TMA mode (csrc/conv_gemm_tc_f16d.cu): the producer warp
walks a cursor (tile, pass, 32-channel half, sub-tile) with `advance`, the epilogue warps derive the same box index from
their own loop counters (g = DT * chunk + t).  Both sides must agree on (index -> tensor coordinates, ring slot, phase) for
every tiling the host can choose, or the mbarrier phases drift and the kernel traps."""
import itertools

DBN = 64


def producer_boxes(n_tiles, grid, block, N, M, DT):
    """(g, c0, row0) in the order the producer issues them (Cur / advance / box_of of the kernel)."""
    nt_total = (N + DBN - 1) // DBN
    tile, nt, c, t = block, 0, 0, 0
    out, g = [], 0
    more = tile < n_tiles
    while more:
        row0 = tile * (DT * 128) + t * 128
        if row0 >= M:
            row0 = 0
        out.append((g, nt * DBN + c, row0))
        g += 1
        # advance
        ncol = min(DBN, N - nt * DBN)
        t += 1
        if t < DT:
            continue
        t = 0
        c += 32
        if c < ncol:
            continue
        c = 0
        nt += 1
        if nt < nt_total:
            continue
        nt = 0
        tile += grid
        more = tile < n_tiles
    return out


def consumer_boxes(n_tiles, grid, block, N, M, DT, n_epi=8):
    """(g, c0, row0, group) as the epilogue warps compute them; group = which four warps read the box."""
    nt_total = (N + DBN - 1) // DBN
    out, chunk = [], 0
    for tile in range(block, n_tiles, grid):
        for nt in range(nt_total):
            nb = nt * DBN
            ncol = min(DBN, N - nb)
            for egrp in range(2):
                split_t, split_c = DT == 2 and n_epi == 8, DT == 1 and n_epi == 8
                t_first = egrp if split_t else 0
                t_last = t_first + 1 if split_t else DT
                c_first, c_step = (32 * egrp, 64) if split_c else (0, 32)
                for t in range(t_first, t_last):
                    row0 = tile * (DT * 128) + t * 128
                    for c in range(c_first, ncol, c_step):
                        ck = chunk + (c >> 5)
                        out.append((DT * ck + t, nb + c, row0 if row0 < M else 0, egrp))
            chunk += ncol >> 5
    return out


def test_producer_and_consumers_agree_on_every_box():
    for DT, N, n_tiles, grid in itertools.product((1, 2), (32, 128, 512), (1, 5, 9), (1, 4)):
        M = n_tiles * DT * 128 - 70                                   # the last sub-tile is partly behind the last row
        if DT == 2 and n_tiles > 1:
            M = (n_tiles - 1) * 256 + 100                             # ... or wholly behind it (t = 1 of the last tile)
        for block in range(min(grid, n_tiles)):
            prod = producer_boxes(n_tiles, grid, block, N, M, DT)
            cons = consumer_boxes(n_tiles, grid, block, N, M, DT)
            assert [p[0] for p in prod] == list(range(len(prod)))
            by_g = {}
            for g, c0, row0, grp in cons:
                assert g not in by_g, 'a box is read by exactly one group of four warps'
                by_g[g] = (c0, row0, grp)
            assert sorted(by_g) == [p[0] for p in prod], (DT, N, n_tiles, grid, block)
            for g, c0, row0 in prod:
                assert by_g[g][:2] == (c0, row0), (DT, N, g)
            # each group sees its boxes in increasing order (a ring slot is waited on in issue order)
            for grp in (0, 1):
                seq = [g for g, _, _, e in cons if e == grp]
                assert seq == sorted(seq)


def test_ring_never_deadlocks():
    """Producer blocks on the slot of box g - rs; consumers of different groups advance independently.  Simulate with the
    slowest legal interleaving (each group waits for its next box) and check everything drains."""
    for DT, N, rs in itertools.product((1, 2), (64, 512), (2, 3, 4)):
        n_tiles, grid, block = 6, 2, 1
        M = n_tiles * DT * 128
        prod = producer_boxes(n_tiles, grid, block, N, M, DT)
        cons = consumer_boxes(n_tiles, grid, block, N, M, DT)
        queues = {e: [g for g, _, _, grp in cons if grp == e] for e in (0, 1)}
        loaded, freed, issued = set(), set(), 0
        progress = True
        while progress:
            progress = False
            while issued < len(prod) and (issued < rs or (issued - rs) in freed):
                loaded.add(issued); issued += 1; progress = True
            for e in (0, 1):
                if queues[e] and queues[e][0] in loaded:
                    freed.add(queues[e].pop(0)); progress = True
        assert issued == len(prod) and not queues[0] and not queues[1], (DT, N, rs)
