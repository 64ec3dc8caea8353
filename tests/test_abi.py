"""The C-ABI library loads and exports every symbol include/iid_b200.h
declares; without a GPU the compute path refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from pyiid_b200 import _lib


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'iid_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(iid_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name


def test_binding_table_covers_the_header():
    names = set(declared_functions())
    bound = set(_lib.SIGNATURES) | {'iid_last_error'}
    assert names == bound, names ^ bound


def test_version_and_bad_arguments():
    lib = _lib.load()
    assert lib.iid_version() >= 100
    assert lib.iid_set_shard(None, 0, 1) != 0
    assert lib.iid_create(0, 7, ctypes.byref(ctypes.c_void_p())) == -1
    assert b'precision' in lib.iid_last_error()


@pytest.mark.skipif(has_gpu(), reason='a GPU is present')
def test_no_cpu_fallback_without_a_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.iid_create(0, 0, ctypes.byref(h))
    assert rc == -5 and not h.value
    from pyiid_b200 import ElasticScatter, structures
    atoms = structures.random_atoms(5, 0)
    with pytest.raises(_lib.IIDError):
        ElasticScatter().get_fq(atoms)
    with pytest.raises(NotImplementedError):
        ElasticScatter().set_processor('CPU', 'flat')


def test_shard_plan_partitions_the_pair_list():
    """Host-only: the per-rank work-item slices cover every (i, j) slot once
    and are balanced (iid_plan_shard, no device)."""
    lib = _lib.load()
    rs = np.random.RandomState(0)
    for n, ntypes in ((1, 1), (33, 1), (561, 1), (2000, 2), (10000, 3)):
        types = np.sort(rs.randint(0, ntypes, n)).astype(np.int32)
        counts = np.bincount(types, minlength=ntypes)
        npad = int(sum((c + 31) // 32 * 32 for c in counts))
        for tri in (0, 1):
            for world in (1, 2, 8):
                slots, mine = [], []
                for rank in range(world):
                    v = [ctypes.c_int64(0) for _ in range(4)]
                    rc = lib.iid_plan_shard(n, types.ctypes.data, ntypes, 148, tri,
                                            rank, world, *[ctypes.byref(x) for x in v])
                    assert rc == 0
                    assert v[3].value == npad
                    slots.append(v[2].value)
                    mine.append(v[1].value)
                    total = v[0].value
                assert sum(mine) == total
                nt = npad // 32
                expect = npad * npad if not tri else (nt * (nt - 1) // 2 + nt) * 1024
                assert sum(slots) == expect
                if total >= 16 * world:
                    assert max(slots) <= 1.15 * (sum(slots) / world) + 32 * 4096


def test_small_structures_get_a_single_wave_triangle_list():
    """Structures of the sampler's size get a triangle list with at most one
    item per SM (the whole evaluation then runs as one cooperative launch,
    iid_fused.cuh) that still covers every tile pair once; larger ones keep the
    list with the shortest makespan over several waves."""
    lib = _lib.load()
    for n, one_wave in ((55, True), (147, True), (309, True), (561, True), (700, True),
                        (923, True), (1100, True), (1415, False), (3000, False)):
        types = np.zeros(n, np.int32)
        v = [ctypes.c_int64(0) for _ in range(4)]
        assert lib.iid_plan_shard(n, types.ctypes.data, 1, 148, 1, 0, 1,
                                  *[ctypes.byref(x) for x in v]) == 0
        total, slots, npad = v[0].value, v[2].value, v[3].value
        nt = npad // 32
        assert slots == (nt * (nt - 1) // 2 + nt) * 1024
        assert (total <= 148) == one_wave, (n, total)


@pytest.mark.parametrize('tier,points,lft0,qh0', [(None, 12, 5, 1. / 3.), (2, 12, 5, 1. / 3.),
                                                  (1, 8, 3, 0.157), (0, 6, 2, 0.0658)])
def test_radial_stencil_interpolates_band_limited_functions(tier, points, lft0, qh0):
    """The Lagrange stencils of iid_stencil.cuh -- 12 points on the grid
    Q_max h = 1/3 (force table of the fused kernel, coarse-grid pair histogram),
    8 points on Q_max h = 0.157 and 6 points on Q_max h = 0.0658 (fine-grid pair
    histograms): the weights are a partition of unity, reproduce polynomials,
    and interpolate sin(Q r)/r to 4e-10 of its amplitude -- the figure DESIGN.md
    quotes."""
    lib = _lib.load()
    if tier is None:
        weights = lib.iid_stencil_weights
    else:
        def weights(*a):
            return lib.iid_hist_stencil_weights(tier, *a)
    w = np.zeros(16)
    npts, left = ctypes.c_int(0), ctypes.c_int(0)
    qh = ctypes.c_double(0.)
    assert weights(0.25, w.ctypes.data, ctypes.byref(npts), ctypes.byref(left),
                   ctypes.byref(qh)) == 0
    n, lft, qmax_h = npts.value, left.value, qh.value
    assert n == points and lft == lft0 and abs(qmax_h - qh0) < 1e-15
    rs = np.random.RandomState(0)
    qmax, worst = 25., 0.
    h = qmax_h / qmax
    for _ in range(400):
        u, k = rs.rand(), rs.randint(0, 9000 * 12 // points)
        assert weights(u, w.ctypes.data, None, None, None) == 0
        ww = w[:n]
        assert abs(ww.sum() - 1.) < 1e-12
        nodes = (k - lft + np.arange(n)) * h
        r = (k + u) * h
        assert abs(ww.dot(nodes ** 3) - r ** 3) < 1e-9 * max(1., r ** 3)
        for q in (qmax, 0.6 * qmax, 3.):
            g = np.where(nodes != 0, np.sin(q * nodes) / np.where(nodes != 0, nodes, 1.), q)
            exact = np.sin(q * r) / r if r > 0 else q
            worst = max(worst, abs(ww.dot(g) - exact) / q)
    assert worst < 4e-10, worst
    # at a node the stencil is the identity
    assert weights(0., w.ctypes.data, None, None, None) == 0
    assert abs(w[lft] - 1.) < 1e-15 and np.abs(np.delete(w[:n], lft)).max() < 1e-15


def test_shard_plan_property_random_structures():
    """Property form of the test above (hypothesis): any element mix, any
    world size -- the ranks' slices partition the work list, the triangle list
    covers every unordered tile pair once and the square list every ordered
    (i, j) slot once (lower items + diagonal tiles + gradient-only mirror
    items above the diagonal)."""
    from hypothesis import given, settings, strategies as st
    lib = _lib.load()

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.integers(1, 900), min_size=1, max_size=4), st.integers(1, 9),
           st.sampled_from([1, 37, 148]))
    def check(counts, world, sms):
        types = np.repeat(np.arange(len(counts)), counts).astype(np.int32)
        n = len(types)
        npad = int(sum((c + 31) // 32 * 32 for c in counts))
        nt = npad // 32
        for tri in (0, 1):
            slots = items = 0
            total = None
            for rank in range(world):
                v = [ctypes.c_int64(0) for _ in range(4)]
                assert lib.iid_plan_shard(n, types.ctypes.data, len(counts), sms, tri, rank,
                                          world, *[ctypes.byref(x) for x in v]) == 0
                total = v[0].value
                items += v[1].value
                slots += v[2].value
                assert v[3].value == npad
            assert items == total
            assert slots == (npad * npad if not tri else (nt * (nt - 1) // 2 + nt) * 1024)

    check()
