"""Generates tests/golden/reference_fixtures.npz from the REFERENCE ITSELF.

Run in the build container (needs oracle/_ref/libtinyad_ref.so, i.e. /root/reference):
    python tests/golden/make_reference_fixtures.py
Each case stores the inputs (term kinds, connectivity, per-element data, x) and what TinyAD::ScalarFunction /
VectorFunction of the unmodified reference (oracle/ref_driver.cc over oracle/eigen_shim) returned for them:
f, g, the sparsity pattern and values of H (eval_with_derivatives) and of the projected H (eval_with_hessian_proj, eps = 1e-9),
or r, J (pattern + values), f = r.r, g = 2 J^T r and the Hessian of every residual (eval_with_derivatives) for vector functions.  tests/test_reference_fixtures.py (CPU: oracle) and
tests/test_reference_fixtures_gpu.py (B200: the product through the C ABI) consume the file; neither needs the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from problems import icosphere, one_ring_table  # noqa: E402
from tinyad_b200 import meshes  # noqa: E402


def scalar_cases():
    rng = np.random.default_rng(2024)
    V_rest, V_init, F, b, bc = meshes.planar_test_mesh()
    yield "planar_newton", 2, len(V_rest), [(oracle.SYMDIRICHLET2D, F, meshes.tri_rest_data(V_rest, F, weight=1.0 / len(F))),
                                             (oracle.PENALTY2D, b.reshape(-1, 1), bc)], V_init.reshape(-1).copy()
    V, F = meshes.grid_2d(7)
    bb = np.array([[0], [7], [8 * 7]], dtype=np.int32)
    yield "tri_grid7", 2, len(V), [(oracle.SYMDIRICHLET2D, F, meshes.tri_rest_data(V, F)), (oracle.PENALTY2D, bb, V[bb[:, 0]] + 0.01)], \
        meshes.deform(V, 1.0 / 7, seed=3).reshape(-1)
    V, T = meshes.kuhn_cube(3, 3, 3)
    bb = np.array([[0], [3], [len(V) - 1]], dtype=np.int32)
    yield "tet_cube3", 3, len(V), [(oracle.SYMDIRICHLET3D, T, meshes.tet_rest_data(V, T)), (oracle.PENALTY3D, bb, V[bb[:, 0]] + 0.01)], \
        meshes.deform(V, 1.0 / 3, seed=1).reshape(-1)
    V, T = meshes.kuhn_cube(4, 3, 2)
    yield "tet_slab_strong", 3, len(V), [(oracle.SYMDIRICHLET3D, T, meshes.tet_rest_data(V, T))], meshes.deform(V, 1.0 / 4, seed=8, noise=0.3).reshape(-1)
    V, F = meshes.grid_2d(6)
    yield "arap_grid6", 2, len(V), [(oracle.ARAP2D, F, meshes.tri_rest_data(V, F))], meshes.deform(V, 1.0 / 6, seed=7).reshape(-1)
    nv = 24
    conn = np.stack([rng.permutation(nv)[:2] for _ in range(40)]).astype(np.int32)
    yield "trig_mix", 2, nv, [(oracle.TRIG_MIX2D, conn, rng.random((40, 1)) + 0.5)], rng.random(2 * nv) * 2.0
    yield "repeated_handle", 2, nv, [(oracle.REPEATED_HANDLE, conn, np.zeros((40, 1)))], rng.random(2 * nv) * 2.0
    yield "edge_dirichlet", 1, nv, [(oracle.EDGE_DIRICHLET1D, conn, np.full((40, 1), 0.5))], rng.random(nv)
    V, F = icosphere(1)
    tab = one_ring_table(len(V), F)
    yield "dyn_one_ring", 1, len(V), [(oracle.DYN_ONERING1D, tab, np.zeros(tab.shape))], rng.standard_normal(len(V))
    yield "dyn_sum_sqr", 2, 4, [(oracle.DYN_SUM_SQR2D, np.zeros((4, 1), dtype=np.int32), np.zeros((4, 1)))], np.ones(8)


def vector_cases():
    rng = np.random.default_rng(77)
    V_rest, V_init, F, b, bc = meshes.planar_test_mesh()
    yield "sos_planar", len(V_rest), [(oracle.SOS_SYMDIRICHLET2D, F, meshes.tri_rest_data(V_rest, F, weight=1.0 / np.sqrt(len(F)))),
                                       (oracle.SOS_PENALTY2D, b.reshape(-1, 1), bc)], V_init.reshape(-1).copy()
    V, F = meshes.grid_2d(6)
    edges = np.unique(np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1), axis=0).astype(np.int32)
    e = V[edges[:, 1]] - V[edges[:, 0]]
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    yield "sos_polycurl", len(V), [(oracle.SOS_POLYCURL2D, edges, np.concatenate([e, rng.random((len(edges), 1)) + 0.5], axis=1))], \
        rng.standard_normal(2 * len(V))


def main():
    out = {}
    names = []
    for name, d, nv, terms, x in scalar_cases():
        ot = [oracle.Term(k, c, dt) for k, c, dt in terms]
        r2 = oracle.ref_scalar_eval(d, nv, ot, oracle.DERIVATIVES, x)
        r3 = oracle.ref_scalar_eval(d, nv, ot, oracle.HESSIAN_PROJ, x, eps=1e-9)
        r0 = oracle.ref_scalar_eval(d, nv, ot, oracle.EVAL, x)
        assert np.array_equal(r2.outer, r3.outer) and np.array_equal(r2.inner, r3.inner)
        p = "s/" + name + "/"
        out[p + "meta"] = np.array([d, nv, len(terms)], dtype=np.int64)
        for i, (k, c, dt) in enumerate(terms):
            out[p + f"kind{i}"] = np.array([k], dtype=np.int64)
            out[p + f"conn{i}"] = np.ascontiguousarray(c, dtype=np.int32)
            out[p + f"data{i}"] = np.ascontiguousarray(dt, dtype=np.float64)
        out[p + "x"] = x
        out[p + "f"] = np.array([r0.f, r2.f, r3.f])
        out[p + "g"] = r2.g
        out[p + "outer"], out[p + "inner"] = r2.outer, r2.inner
        out[p + "H"], out[p + "H_proj"] = r2.values, r3.values
        names.append(p)
    for name, nv, terms, x in vector_cases():
        ot = [oracle.Term(k, c, dt) for k, c, dt in terms]
        r = oracle.ref_vector_eval(2, nv, ot, oracle.V_SOS_DERIVATIVES, x)
        p = "v/" + name + "/"
        out[p + "meta"] = np.array([2, nv, len(terms)], dtype=np.int64)
        for i, (k, c, dt) in enumerate(terms):
            out[p + f"kind{i}"] = np.array([k], dtype=np.int64)
            out[p + f"conn{i}"] = np.ascontiguousarray(c, dtype=np.int32)
            out[p + f"data{i}"] = np.ascontiguousarray(dt, dtype=np.float64)
        out[p + "x"] = x
        out[p + "f"] = np.array([r.f])
        out[p + "g"], out[p + "r"] = r.g, r.r
        out[p + "outer"], out[p + "inner"], out[p + "J"] = r.outer, r.inner, r.values
        # VectorFunction::eval_with_derivatives: every stored entry of every residual's n x n Hessian
        rd = oracle.ref_vector_eval(2, nv, ot, oracle.V_DERIVATIVES, x)
        for u, v in ((rd.r, r.r), (rd.values, r.values)):          # last-bit differences only (FMA contraction differs per instantiation)
            assert np.abs(u - v).max() <= 1e-15 * np.abs(v).max()
        hr, hi, hj, hv = rd.phases["residual_hessians"]
        out[p + "hess_res"], out[p + "hess_row"], out[p + "hess_col"], out[p + "hess_val"] = hr, hi, hj, hv
        names.append(p)
    path = os.path.join(HERE, "reference_fixtures.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {len(names)} cases, {os.path.getsize(path) / 1024:.0f} KiB; made with {oracle.ref_lib().ref_description().decode()}")


if __name__ == "__main__":
    main()
