"""Synthetic meshes and per-element rest data of the BASELINE.json configurations.

C1: (N+1)^2 grid on [0,1]^2, each cell split along the same diagonal into two CCW triangles.
C2 / C5: (n+1)^3 lattice, each cube split into 6 positively oriented Kuhn tetrahedra.
Element order is lexicographic cell order (x fastest), which keeps gathers and scatters local.
"""
import itertools

import numpy as np


def planar_test_mesh():
    """6-vertex / 4-face fixture of the reference (tests/Meshes.hh:12-42): rest, stretched init, F, b, bc."""
    V_rest = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0, 2], [1, 2]], dtype=np.float64)
    V_init = V_rest.copy()
    V_init[:, 0] *= 0.5
    V_init[:, 1] *= 0.25
    F = np.array([[0, 1, 2], [1, 3, 2], [2, 3, 5], [2, 5, 4]], dtype=np.int32)
    b = np.array([0, 4], dtype=np.int32)
    bc = np.array([[0.0, 0.0], [-2.0, 0.0]])
    return V_rest, V_init, F, b, bc


def grid_2d(N):
    """Returns V ((N+1)^2, 2), F (2 N^2, 3) int32."""
    xs = np.arange(N + 1, dtype=np.float64) / N
    X, Y = np.meshgrid(xs, xs, indexing="xy")          # vertex (i, j) -> index i + (N+1) j, position (i/N, j/N)
    V = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="xy")
    v00 = (i + (N + 1) * j).ravel()
    v10, v01, v11 = v00 + 1, v00 + (N + 1), v00 + (N + 2)
    F = np.empty((2 * N * N, 3), dtype=np.int32)
    F[0::2] = np.stack([v00, v10, v11], axis=1)
    F[1::2] = np.stack([v00, v11, v01], axis=1)
    return V, F


def kuhn_cube(nx, ny=None, nz=None, z0=0, nz_total=None):
    """Lattice of nx x ny x nz cubes (cube layers z0 .. z0+nz-1 of a lattice with nz_total layers), 6 Kuhn
    tets each.  Returns V (all (nx+1)(ny+1)(nz_total+1) lattice vertices, spacing 1/nx) and T (6 nx ny nz, 4) int32
    with global vertex indices, so slabs of one big mesh can be generated rank by rank."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    nz_total = nz if nz_total is None else nz_total
    h = 1.0 / nx
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz_total + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    V = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64) * h
    ck, cj, ci = np.meshgrid(np.arange(z0, z0 + nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (ci + sx * cj + sy * ck).ravel().astype(np.int64)
    step = np.array([1, sx, sy], dtype=np.int64)
    tets = []
    for perm in itertools.permutations(range(3)):
        v0 = base
        v1 = v0 + step[perm[0]]
        v2 = v1 + step[perm[1]]
        v3 = v2 + step[perm[2]]
        # orientation = sign of the permutation; swap two vertices of the odd ones
        inv = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b])
        tets.append(np.stack([v0, v1, v2, v3] if inv % 2 == 0 else [v0, v2, v1, v3], axis=1))
    T = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)   # cell-major, 6 consecutive tets per cube
    return V, T


def tri_rest_data(V_rest, F, weight=None):
    """Per-triangle data of the 2-D symmetric Dirichlet functor: Mr (row-major 2x2, columns b-a, c-a), weight."""
    a, b, c = V_rest[F[:, 0]], V_rest[F[:, 1]], V_rest[F[:, 2]]
    Mr = np.stack([b - a, c - a], axis=2)                       # (f, 2, 2), columns
    if weight is None:
        weight = 0.5 * np.abs(Mr[:, 0, 0] * Mr[:, 1, 1] - Mr[:, 0, 1] * Mr[:, 1, 0])   # area weight
    w = np.broadcast_to(np.asarray(weight, dtype=np.float64), (len(F),))
    return np.concatenate([Mr.reshape(-1, 4), w[:, None]], axis=1)


def tet_rest_data(V_rest, T):
    """Per-tet data of the 3-D symmetric Dirichlet functor: Mr^-1 (row-major 3x3), volume."""
    a = V_rest[T[:, 0]]
    Mr = np.stack([V_rest[T[:, 1]] - a, V_rest[T[:, 2]] - a, V_rest[T[:, 3]] - a], axis=2)
    det = np.linalg.det(Mr)
    assert np.all(det > 0), "rest tets must be positively oriented"
    return np.concatenate([np.linalg.inv(Mr).reshape(-1, 9), (det / 6.0)[:, None]], axis=1)


def deform(V, h, seed=0, noise=0.2, smooth=0.15):
    """x = rest + smooth deformation + uniform noise of amplitude noise*h (no element inverts for noise <= 0.2)."""
    rng = np.random.default_rng(seed)
    d = V.shape[1]
    X = V.copy()
    ph = 2.0 * np.pi * V
    if d == 2:
        X[:, 0] += smooth * h * 4 * np.sin(ph[:, 1]) * np.cos(0.5 * ph[:, 0])
        X[:, 1] += smooth * h * 4 * np.sin(ph[:, 0])
    else:
        X[:, 0] += smooth * h * 4 * np.sin(ph[:, 1]) * np.cos(ph[:, 2])
        X[:, 1] += smooth * h * 4 * np.sin(ph[:, 2]) * np.cos(0.5 * ph[:, 0])
        X[:, 2] += smooth * h * 4 * np.sin(ph[:, 0])
    X += noise * h * (rng.random(V.shape) * 2.0 - 1.0) / np.sqrt(d)
    return X
