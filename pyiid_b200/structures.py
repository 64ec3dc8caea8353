"""Synthetic structures for tests and benchmarks:
random-perturbed fcc spheres and Mackay icosahedra.  Stands in for the
reference's ASE / asap3 builders (``pyiid/utils.py:24-42``,
``examples/Au_NP_PDF.py:14``), which are not installed here."""
import numpy as np

from . import ase_shim

A_AU = 4.0782
A_PT = 3.9242
A_AUPT = 4.0012


def _atoms_class():
    if ase_shim.have_real_ase():  # pragma: no cover
        from ase import Atoms
        return Atoms
    return ase_shim.Atoms


def fcc_sphere_positions(n, a, sigma=0.05, seed=0):
    """The n fcc lattice sites nearest the origin (ties by x, y, z), perturbed
    by N(0, sigma) per coordinate and shifted into the positive octant."""
    m = int(np.ceil((3.0 * n / (16.0 * np.pi)) ** (1.0 / 3.0))) + 2
    rng = np.arange(-m, m + 1)
    i, j, k = np.meshgrid(rng, rng, rng, indexing='ij')
    cells = np.stack([i.ravel(), j.ravel(), k.ravel()], 1).astype(np.float64)
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    pts = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3)
    r2 = np.round((pts ** 2).sum(1) * 4).astype(np.int64)
    order = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0], r2))
    if len(order) < n:
        raise ValueError('lattice block too small')
    pos = pts[order[:n]] * a
    if sigma:
        pos = pos + np.random.RandomState(seed).normal(0, sigma, (n, 3))
    return pos - pos.min(0)


def fcc_sphere(symbol, n, a=None, sigma=0.05, seed=0):
    if a is None:
        a = {'Au': A_AU, 'Pt': A_PT}[symbol]
    Atoms = _atoms_class()
    z = ase_shim.atomic_numbers[symbol]
    return Atoms(numbers=[z] * n, positions=fcc_sphere_positions(n, a, sigma, seed))


def alloy_sphere(n, a=A_AUPT, sigma=0.05, seed=1, symbols=('Au', 'Pt')):
    """50:50 random alloy: species = RandomState(seed).rand(n) < 0.5."""
    Atoms = _atoms_class()
    pick = np.random.RandomState(seed).rand(n) < 0.5
    z0, z1 = (ase_shim.atomic_numbers[s] for s in symbols)
    numbers = np.where(pick, z0, z1)
    return Atoms(numbers=numbers, positions=fcc_sphere_positions(n, a, sigma, seed))


def icosahedron_positions(shells, nn=2.88):
    """Mackay icosahedron with ``shells`` complete shells around a centre atom
    (1, 13, 55, 147, 309, 561 atoms for 0..5 shells); edge spacing ``nn``."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    verts = []
    for s1 in (-1, 1):
        for s2 in (-1, 1):
            verts += [(0, s1, s2 * t), (s1, s2 * t, 0), (s2 * t, 0, s1)]
    verts = np.array(verts, dtype=np.float64)
    edge = 2.0
    d = np.linalg.norm(verts[:, None] - verts[None], axis=2)
    nbr = np.abs(d - edge) < 1e-9
    edges = [(a, b) for a in range(12) for b in range(a + 1, 12) if nbr[a, b]]
    faces = [(a, b, c) for a in range(12) for b in range(a + 1, 12)
             for c in range(b + 1, 12) if nbr[a, b] and nbr[b, c] and nbr[a, c]]
    pts = [np.zeros(3)]
    for n in range(1, shells + 1):
        for v in verts:
            pts.append(v * n)
        for a, b in edges:
            for i in range(1, n):
                pts.append((verts[a] * (n - i) + verts[b] * i))
        for a, b, c in faces:
            for i in range(1, n):
                for j in range(1, n - i):
                    k = n - i - j
                    pts.append(verts[a] * i + verts[b] * j + verts[c] * k)
    pts = np.array(pts) * (nn / edge)
    return pts


def icosahedron(symbol, shells, nn=2.88):
    Atoms = _atoms_class()
    pos = icosahedron_positions(shells, nn)
    z = ase_shim.atomic_numbers[symbol]
    atoms = Atoms(numbers=[z] * len(pos), positions=pos)
    return atoms


def random_atoms(n, seed, box=10.0, symbol='Au'):
    """``tests/__init__.py:83-95`` setup_atoms: uniform random positions in a
    ``box`` A cube, centred."""
    Atoms = _atoms_class()
    q = np.random.RandomState(seed).random_sample((n, 3)) * box
    atoms = Atoms(numbers=[ase_shim.atomic_numbers[symbol]] * n, positions=q)
    atoms.center()
    return atoms


def atomic_square():
    """``tests/__init__.py:131-141``: Au4 3 A square and its 0.75-scaled copy."""
    Atoms = _atoms_class()
    a1 = Atoms('Au4', [[0, 0, 0], [3, 0, 0], [0, 3, 0], [3, 3, 0]])
    a1.center()
    a2 = a1.copy()
    a2.positions *= .75
    return a1, a2
