"""Host-only checks of the row-ownership plan of the full-gradient pass
(iid_plan_rows, no device): every (i-tile, j) slot is covered exactly once
across the ranks, every i-tile belongs to one rank, the pieces of a split row
are consecutive side-buffer slots in j order, and the longest-first schedule
on the SM count ends close to the ideal."""
import ctypes
import heapq

import numpy as np
import pytest

from pyiid_b200 import _lib

DIAG, NOF, FLUSH = 1 << 16, 1 << 17, 1 << 18


def plan(types, ntypes, sms, rank, world, div=0):
    lib = _lib.load()
    n = len(types)
    c = [ctypes.c_int64(0) for _ in range(3)]
    assert lib.iid_plan_rows(n, types.ctypes.data, ntypes, sms, rank, world, div,
                             *[ctypes.byref(x) for x in c], None, 0, None, 0, None, 0) == 0
    nj, ns, nf = [x.value for x in c]
    jobs = np.zeros((nj, 4), np.int32)
    segs = np.zeros((ns, 4), np.int32)
    fixes = np.zeros((max(nf, 1), 4), np.int32)
    assert lib.iid_plan_rows(n, types.ctypes.data, ntypes, sms, rank, world, div,
                             *[ctypes.byref(x) for x in c], jobs.ctypes.data, nj,
                             segs.ctypes.data, ns, fixes.ctypes.data, nf) == 0
    return jobs, segs, fixes[:nf]


def job_cost(job, segs):
    return sum((je - jb + 16) * (6.7 if info & NOF else 8.25)
               for jb, je, info, _ in segs[job[1]:job[2]])


@pytest.mark.parametrize('counts,world,sms', [
    ((55,), 1, 148), ((561,), 1, 148), ((700, 300), 3, 148), ((2000,), 2, 37),
    ((5000, 5000), 8, 148), ((33, 1, 64), 2, 4)])
def test_rows_are_covered_exactly_once(counts, world, sms):
    types = np.repeat(np.arange(len(counts)), counts).astype(np.int32)
    npad = int(sum((c + 31) // 32 * 32 for c in counts))
    nt = npad // 32
    # element type of every padded j and the run boundaries
    jtype = np.concatenate([np.full((c + 31) // 32 * 32, e) for e, c in enumerate(counts)])
    cover = np.zeros((nt, npad), np.int32)
    owner = np.full(nt, -1)
    for rank in range(world):
        jobs, segs, fixes = plan(types, len(counts), sms, rank, world)
        split = {}
        for it, s0, s1, dest in jobs:
            assert s1 > s0
            assert owner[it] in (-1, rank)
            owner[it] = rank
            assert it % world == rank
            for jb, je, info, _ in segs[s0:s1]:
                assert 0 <= jb < je <= npad and jb % 32 == 0 and je % 32 == 0
                cover[it, jb:je] += 1
                ty = info & 0xffff
                assert (jtype[jb:je] == ty).all()            # one element run per segment
                lo, hi = it * 32, it * 32 + 32
                if info & DIAG:
                    assert (jb, je) == (lo, hi) and not info & NOF
                elif info & NOF:
                    assert jb >= hi                            # above the diagonal: gradient only
                else:
                    assert je <= lo                            # below it: carries F(Q)
            # a flush between two segments exactly where the element type changes
            for k in range(s0, s1 - 1):
                change = (segs[k][2] & 0xffff) != (segs[k + 1][2] & 0xffff)
                assert bool(segs[k][2] & FLUSH) == change or not change
                if change:
                    assert segs[k][2] & FLUSH
            if dest >= 0:
                split.setdefault(it, []).append((dest, segs[s0][0]))
        # split rows: consecutive slots, in j order, one RowFix each
        assert sorted(split) == sorted(f[0] for f in fixes)
        for it, d0, d1, _ in fixes:
            pcs = sorted(split[it])
            assert [p[0] for p in pcs] == list(range(d0, d1)) and d1 - d0 >= 2
            assert [p[1] for p in pcs] == sorted(p[1] for p in pcs)
        slots = sorted(d for _, _, _, d in jobs if d >= 0)
        assert slots == list(range(len(slots)))
    assert (cover == 1).all()
    assert (owner >= 0).all()


@pytest.mark.parametrize('n,world', [(10000, 1), (50000, 1), (50000, 8), (100000, 8)])
def test_schedule_is_balanced(n, world):
    """List scheduling of the jobs in the order the block scheduler sees them
    (one block per SM) ends within 1.5 % of the ideal."""
    types = np.zeros(n, np.int32)
    jobs, segs, _ = plan(types, 1, 148, 0, world)
    cost = [job_cost(j, segs) + 300 for j in jobs]
    assert cost == sorted(cost, reverse=True)
    h = [0.0] * 148
    heapq.heapify(h)
    for c in cost:
        heapq.heappush(h, heapq.heappop(h) + c)
    assert sum(cost) / 148 / max(h) > 0.985
