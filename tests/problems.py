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
