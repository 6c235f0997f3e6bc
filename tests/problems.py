"""Shared problem builders: the same numpy inputs go to the CPU oracle and to the CUDA path."""
import numpy as np

import oracle
import tinyad_b200 as tad
from tinyad_b200 import meshes


class Problem:
    """d, n_vertices, list of (kind, conn, data); kinds are numerically identical in oracle and product."""

    def __init__(self, d, n_vertices, terms, is_vector=False):
        self.d, self.n_vertices, self.terms, self.is_vector = d, n_vertices, terms, is_vector

    def oracle_terms(self):
        return [oracle.Term(k, c, dt) for k, c, dt in self.terms]

    def gpu(self, **kw):
        f = tad.Function(self.d, self.n_vertices, is_vector=self.is_vector, **kw)
        for k, c, dt in self.terms:
            f.add_term(k, c, dt)
        return f


def planar_newton_problem():
    """tests/NewtonTest.cc:12-58: symmetric Dirichlet on 4 triangles (weight 1/#F) + 2 positional penalties."""
    V_rest, V_init, F, b, bc = meshes.planar_test_mesh()
    data = meshes.tri_rest_data(V_rest, F, weight=1.0 / len(F))
    p = Problem(2, len(V_rest), [(tad.SYMDIRICHLET2D, F, data), (tad.PENALTY2D, b.reshape(-1, 1), bc)])
    return p, V_init.reshape(-1).copy()


def grid_problem(N, seed=0, with_penalty=False):
    V, F = meshes.grid_2d(N)
    data = meshes.tri_rest_data(V, F)
    terms = [(tad.SYMDIRICHLET2D, F, data)]
    if with_penalty:
        b = np.array([[0], [N], [(N + 1) * N]], dtype=np.int32)
        terms.append((tad.PENALTY2D, b, V[b[:, 0]] + 0.01))
    x = meshes.deform(V, 1.0 / N, seed=seed).reshape(-1)
    return Problem(2, len(V), terms), x


def tet_problem(n, seed=0, with_penalty=False, nz=None):
    V, T = meshes.kuhn_cube(n, n, n if nz is None else nz)
    data = meshes.tet_rest_data(V, T)
    terms = [(tad.SYMDIRICHLET3D, T, data)]
    if with_penalty:
        b = np.array([[0], [n], [len(V) - 1]], dtype=np.int32)
        terms.append((tad.PENALTY3D, b, V[b[:, 0]] + 0.01))
    x = meshes.deform(V, 1.0 / n, seed=seed).reshape(-1)
    return Problem(3, len(V), terms), x


def icosphere(subdivisions=1):
    """Closed triangle mesh with vertex valences 5 and 6 (consistently oriented): icosahedron, each face split in four."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    V = [np.array(v, dtype=float) / np.linalg.norm(v) for v in V]
    for _ in range(subdivisions):
        mid, F2 = {}, []

        def m(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                p = V[a] + V[b]
                V.append(p / np.linalg.norm(p))
                mid[key] = len(V) - 1
            return mid[key]
        for a, b, c in F:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            F2 += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        F = F2
    return np.array(V), np.array(F, dtype=np.int32)


def one_ring_table(n_vertices, F, width=9):
    """DynamicElementsTest.cc:96-105: vertex -> neighbours from the DIRECTED face edges (v_i -> v_{i+1}), padded with -1."""
    nbrs = [[] for _ in range(n_vertices)]
    for f in F:
        for i in range(3):
            nbrs[int(f[i])].append(int(f[(i + 1) % 3]))
    tab = -np.ones((n_vertices, width), dtype=np.int32)
    for v, lst in enumerate(nbrs):
        assert len(lst) <= width
        tab[v, :len(lst)] = lst
    return tab


def load_reference_fixtures():
    """tests/golden/reference_fixtures.npz (made by tests/golden/make_reference_fixtures.py from the reference itself):
    returns {name: dict} with name like "s/tet_cube3/" (scalar functions) or "v/sos_planar/" (vector functions); every dict holds
    d, n_vertices, terms [(kind, conn, data)], x and the reference's outputs."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fixtures.npz"))
    cases = {}
    for key in z.files:
        if key.endswith("/meta"):
            p = key[:-4]
            d, nv, nt = (int(v) for v in z[key])
            c = {"d": d, "n_vertices": nv, "x": z[p + "x"],
                 "terms": [(int(z[p + f"kind{i}"][0]), z[p + f"conn{i}"], z[p + f"data{i}"]) for i in range(nt)]}
            for field in ("f", "g", "r", "outer", "inner", "H", "H_proj", "J", "hess_res", "hess_row", "hess_col", "hess_val"):
                if p + field in z.files:
                    c[field] = z[p + field]
            cases[p] = c
    return cases
