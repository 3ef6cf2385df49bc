"""CPU model of the residual-box ring of the direct kernel's TMA mode (csrc/conv_gemm_tc_f16d.cu).  The producer warp
walks a cursor (tile, pass, 32-channel half, sub-tile); each group of four epilogue warps owns one half of the ring and
derives (slot, phase) of its k-th box from its own loop counters.  Both sides must agree on (slot, phase, tensor
coordinates) for every tiling the host can choose, every slot must have exactly one consumer group and its phases must
advance one at a time, or the mbarrier parities drift and the kernel reads the wrong box."""
import itertools

DBN = 64


def producer_boxes(n_tiles, grid, block, N, M, DT, rs):
    """(slot, use, c0, row0) in issue order (Cur / advance / box_of + the slot rule of the producer loop)."""
    nt_total = (N + DBN - 1) // DBN
    half = rs >> 1
    tile, nt, c, t = block, 0, 0, 0
    out, ck = [], 0
    more = tile < n_tiles
    while more:
        grp = t if DT == 2 else c >> 5
        k = ck if DT == 2 else ck >> 1
        row0 = tile * (DT * 128) + t * 128
        out.append((grp * half + k % half, k // half, nt * DBN + c, row0 if row0 < M else 0))
        if t == DT - 1:
            ck += 1
        ncol = min(DBN, N - nt * DBN)                                # advance
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


def consumer_boxes(n_tiles, grid, block, N, M, DT, rs):
    """per group: [(slot, use, c0, row0)] in the order its four warps wait for them."""
    nt_total = (N + DBN - 1) // DBN
    half = rs >> 1
    seq = {0: [], 1: []}
    for egrp in (0, 1):
        chunk = 0
        for tile in range(block, n_tiles, grid):
            for nt in range(nt_total):
                nb = nt * DBN
                ncol = min(DBN, N - nb)
                split_t, split_c = DT == 2, DT == 1                   # (8 epilogue warps)
                t_first = egrp if split_t else 0
                t_last = t_first + 1 if split_t else DT
                c_first, c_step = (32 * egrp, 64) if split_c else (0, 32)
                for t in range(t_first, t_last):
                    row0 = tile * (DT * 128) + t * 128
                    for c in range(c_first, ncol, c_step):
                        ck = chunk + (c >> 5)
                        k = ck if DT == 2 else ck >> 1
                        seq[egrp].append((egrp * half + k % half, k // half, nb + c, row0 if row0 < M else 0))
                chunk += ncol >> 5
    return seq


def test_producer_and_consumers_agree_on_every_box():
    for DT, N, n_tiles, grid, rs in itertools.product((1, 2), (64, 128, 512), (1, 5, 9), (1, 4), (2, 4)):
        M = n_tiles * DT * 128 - 70                                   # the last sub-tile is partly behind the last row
        if DT == 2 and n_tiles > 1:
            M = (n_tiles - 1) * 256 + 100                             # ... or wholly behind it (t = 1 of the last tile)
        half = rs >> 1
        for block in range(min(grid, n_tiles)):
            prod = producer_boxes(n_tiles, grid, block, N, M, DT, rs)
            cons = consumer_boxes(n_tiles, grid, block, N, M, DT, rs)
            for grp in (0, 1):
                mine = [b for b in prod if b[0] // half == grp]       # the producer's boxes in this group's half of the ring
                assert mine == cons[grp], (DT, N, n_tiles, grid, rs, block, grp)
            assert len(prod) == len(cons[0]) + len(cons[1])
            # per slot the phases advance one at a time, starting at 0
            for slot in range(rs):
                uses = [b[1] for b in prod if b[0] == slot]
                assert uses == list(range(len(uses))), (DT, N, rs, slot)


def test_ring_never_deadlocks():
    """The producer refills a slot only after its previous box has been released; the two groups advance independently.
    Simulate the slowest legal interleaving and check that everything drains."""
    for DT, N, rs in itertools.product((1, 2), (64, 512), (2, 4)):
        n_tiles, grid, block = 6, 2, 1
        M = n_tiles * DT * 128
        prod = producer_boxes(n_tiles, grid, block, N, M, DT, rs)
        cons = consumer_boxes(n_tiles, grid, block, N, M, DT, rs)
        queues = {e: list(cons[e]) for e in (0, 1)}
        loaded, released, issued = set(), set(), 0
        progress = True
        while progress:
            progress = False
            while issued < len(prod):
                slot, use = prod[issued][:2]
                if use > 0 and (slot, use - 1) not in released:
                    break
                loaded.add((slot, use)); issued += 1; progress = True
            for e in (0, 1):
                if queues[e] and queues[e][0][:2] in loaded:
                    released.add(queues[e].pop(0)[:2]); progress = True
        assert issued == len(prod) and not queues[0] and not queues[1], (DT, N, rs)
