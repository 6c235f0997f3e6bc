"""CPU: the device projection routine (Detail/Projection.hh is __host__ __device__) compiled for the host,
checked against numpy.linalg.eigh and the oracle on random, degenerate, clustered and badly scaled matrices."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "host_projection.cc")
    so = os.path.join(HERE, "host", "libhost_projection.so")
    hdr = os.path.join(ROOT, "tinyad_b200", "include", "TinyAD", "Detail", "Projection.hh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "tinyad_b200", "include"),
                        "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    L.host_project.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
    L.host_project_jacobi.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
    L.host_project_translations.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
    L.host_deflated_dim.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
    L.host_project_reduced.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
    return L


def pack(L, A):
    k = A.shape[0]
    out = np.zeros(k * (k + 1) // 2)
    for i in range(k):
        for j in range(i + 1):
            out[L.host_seq_index(k, i, j)] = A[i, j]
    return out


def unpack(L, p, k):
    A = np.zeros((k, k))
    for i in range(k):
        for j in range(i + 1):
            A[i, j] = A[j, i] = p[L.host_seq_index(k, i, j)]
    return A


def reference_projection(A, eps):
    w, V = np.linalg.eigh(A)
    t = np.abs(w) if eps < 0 else np.maximum(w, eps)
    return (V * t) @ V.T


def matrices(k, rng):
    out = []
    for scale in (1.0, 1e-6, 1e6):
        for _ in range(6):
            A = rng.standard_normal((k, k)); A = (A + A.T) * scale
            out.append(A)
        B = rng.standard_normal((k, max(1, k - 3))); out.append(B @ B.T * scale)           # PSD, null space of dim 3
        B = rng.standard_normal((k, max(1, k // 2))); out.append(-(B @ B.T) * scale)        # NSD, big null space
        Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        w = np.concatenate([np.zeros(min(3, k)), rng.standard_normal(max(0, k - 3))])[:k]
        out.append((Q * w) @ Q.T * scale)                                                  # exact triple zero eigenvalue
        w = np.repeat(rng.standard_normal((k + 2) // 3), 3)[:k]
        out.append((Q * w) @ Q.T * scale)                                                  # triple clusters everywhere
        w = -np.abs(rng.standard_normal(k)) - 0.1
        out.append((Q * w) @ Q.T * scale)                                                  # negative definite (all clamped)
        D = np.diag(rng.standard_normal(k)); out.append(D * scale)                         # diagonal (T splits completely)
        if k >= 4:
            Bd = np.zeros((k, k)); h = k // 2
            X = rng.standard_normal((h, h)); Bd[:h, :h] = X + X.T
            Y = rng.standard_normal((k - h, k - h)); Bd[h:, h:] = Y + Y.T
            out.append(Bd * scale)                                                         # block diagonal
    out.append(np.zeros((k, k)))
    out.append(-np.eye(k))
    out.append(np.ones((k, k)))
    return [0.5 * (A + A.T) for A in out]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 18])
@pytest.mark.parametrize("eps", [1e-9, -1.0, 0.0])
def test_projection_host(hostlib, k, eps):
    rng = np.random.default_rng(1000 + k)
    n_fallback = 0
    mats = matrices(k, rng)
    for idx, A in enumerate(mats):
        p = pack(hostlib, A)
        code = hostlib.host_project(k, p.ctypes.data, eps)
        assert code in (0, 1, 2, 3), (k, idx, code)
        got = unpack(hostlib, p, k)
        if code == 3:
            n_fallback += 1
            assert np.array_equal(got, A)
            continue
        ora, ocode = oracle.project(A, eps)
        # the reference's dominance early-out uses eps as given, even when negative (HessianProjection.hh:37,62)
        ref = A if ocode == 0 else reference_projection(A, eps)
        scale = max(np.abs(A).max(), abs(eps), 1e-300)
        err = np.abs(got - ref).max() / scale
        assert err <= 5e-12, (k, idx, code, err)
        if code < 2:
            assert np.array_equal(got, A)
        assert (code == 0) == (ocode == 0)
        assert np.abs(got - ora).max() / scale <= 5e-12
    assert n_fallback <= len(mats) // 20, n_fallback


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 8])
@pytest.mark.parametrize("eps", [1e-9, -1.0, 0.0])
def test_jacobi_fallback_host(hostlib, k, eps):
    """project_full_jacobi (the in-kernel full solver of the fused small-k element kernel): H + sum (clamp(l) - l) v v^T."""
    rng = np.random.default_rng(2000 + k)
    for idx, A in enumerate(matrices(k, rng)):
        p = pack(hostlib, A)
        code = hostlib.host_project_jacobi(k, p.ctypes.data, eps)
        assert code in (1, 2), (k, idx, code)
        got = unpack(hostlib, p, k)
        ref = reference_projection(A, eps)
        scale = max(np.abs(A).max(), abs(eps), 1e-300)
        assert np.abs(got - ref).max() / scale <= 5e-12, (k, idx, code)
        if code == 1:
            assert np.array_equal(got, A)


def test_projection_host_element_hessians(hostlib):
    """Element Hessians of the tet / triangle energies (translation null space -> clustered zero eigenvalues)."""
    from problems import tet_problem, grid_problem
    import scipy.sparse as sp
    for (p, x), k in ((tet_problem(3, seed=4), 12), (grid_problem(6, seed=4), 6)):
        kind, conn, data = p.terms[0]
        for e in range(0, len(conn), 7):
            # Hessian of one element through the oracle: a 1-element problem on its own vertices
            loc = np.arange(conn.shape[1], dtype=np.int32)[None, :]
            xe = x.reshape(-1, p.d)[conn[e]].reshape(-1)
            r = oracle.scalar_eval(p.d, conn.shape[1], [oracle.Term(kind, loc, data[e:e + 1])], oracle.DERIVATIVES, xe)
            A = sp.csc_matrix((r.values, r.inner, r.outer), shape=(k, k)).toarray()
            A = 0.5 * (A + A.T)
            pk = pack(hostlib, A)
            code = hostlib.host_project(k, pk.ctypes.data, 1e-9)
            assert code == 2
            ref = reference_projection(A, 1e-9)
            assert np.abs(unpack(hostlib, pk, k) - ref).max() <= 1e-11 * np.abs(A).max()


def translation_projector(k, d):
    """Orthogonal projector onto the complement of the d translations of k / d handles (local index = d * handle + component)."""
    n = k // d
    U = np.zeros((k, d))
    for a in range(d):
        U[a::d, a] = 1.0 / np.sqrt(n)
    return np.eye(k) - U @ U.T


@pytest.mark.parametrize("k,d", [(12, 3), (9, 3), (6, 3), (6, 2), (8, 2), (4, 2)])
@pytest.mark.parametrize("eps", [1e-9, 0.0, -1.0, 1e-3])
def test_translation_null_space_deflation_host(hostlib, k, d, eps):
    """proj_tridiagonalize<K, D> + the deflated phases B2 / C: matrices WITH the translation null space (random, with negative
    eigenvalues, with an additional null vector, scaled, zero) and WITHOUT it (generic: the test must say no) against numpy."""
    rng = np.random.default_rng(3000 + 10 * k + d)
    P = translation_projector(k, d)
    mats = []
    for scale in (1.0, 1e-5, 1e4):
        for _ in range(6):
            A = rng.standard_normal((k, k)); A = P @ (A + A.T) @ P * scale            # indefinite, translation invariant
            mats.append(A)
        B = rng.standard_normal((k, k)); mats.append(P @ (B @ B.T) @ P * scale)      # PSD, exactly the translation null space
        B = rng.standard_normal((k, max(1, k - d - 1))); mats.append(P @ (B @ B.T) @ P * scale)   # one more null vector
        A = rng.standard_normal((k, k)); mats.append((A + A.T) * scale)              # not translation invariant
        A = rng.standard_normal((k, k)); A = P @ (A + A.T) @ P
        mats.append((A + 1e-9 * rng.standard_normal((k, k))) * scale)                # invariance violated at 1e-9: must not be deflated
    mats.append(np.zeros((k, k)))
    mats.append(-P)
    n_fallback = 0
    for idx, A in enumerate(mats):
        A = 0.5 * (A + A.T)
        p = pack(hostlib, A)
        code = hostlib.host_project_translations(k, d, p.ctypes.data, eps)
        assert code in (0, 1, 2, 3), (k, d, idx, code)
        got = unpack(hostlib, p, k)
        if code == 3:
            n_fallback += 1
            assert np.array_equal(got, A)
            continue
        ora, ocode = oracle.project(A, eps)
        ref = A if ocode == 0 else reference_projection(A, eps)
        scale = max(np.abs(A).max(), abs(eps), 1e-300)
        assert np.abs(got - ref).max() / scale <= 5e-12, (k, d, idx, code, np.abs(got - ref).max() / scale)
        assert np.abs(got - ora).max() / scale <= 5e-12
        if code < 2:
            assert np.array_equal(got, A)
    assert n_fallback <= 2, n_fallback


def test_translation_deflation_on_element_hessians(hostlib):
    """Real element Hessians of the tet / triangle deformation energies through the deflated path (what the kernels run)."""
    from problems import tet_problem, grid_problem
    import scipy.sparse as sp
    for (p, x), k in ((tet_problem(3, seed=4), 12), (grid_problem(6, seed=4), 6)):
        kind, conn, data = p.terms[0]
        for e in range(0, len(conn), 5):
            loc = np.arange(conn.shape[1], dtype=np.int32)[None, :]
            xe = x.reshape(-1, p.d)[conn[e]].reshape(-1)
            r = oracle.scalar_eval(p.d, conn.shape[1], [oracle.Term(kind, loc, data[e:e + 1])], oracle.DERIVATIVES, xe)
            A = sp.csc_matrix((r.values, r.inner, r.outer), shape=(k, k)).toarray()
            A = 0.5 * (A + A.T)
            for eps in (1e-9, 1e-6):
                pk = pack(hostlib, A)
                if np.abs(A).max() > 0:                                                        # (an inverted element has a zero Hessian)
                    assert hostlib.host_deflated_dim(k, p.d, pk.ctypes.data, eps) == p.d      # the deflated path really runs on them
                code = hostlib.host_project_translations(k, p.d, pk.ctypes.data, eps)
                assert code == 2
                ref = reference_projection(A, eps)
                assert np.abs(unpack(hostlib, pk, k) - ref).max() <= 1e-11 * np.abs(A).max()


@pytest.mark.parametrize("k,d", [(12, 3), (8, 2), (4, 1)])
@pytest.mark.parametrize("eps", [1e-9, 0.0, -1.0, 1e-3])
def test_reduced_pipeline_host(hostlib, k, d, eps):
    """proj_tridiagonalize<K, D, true> (Hadamard reduction to the complement of the translations) + phases B1 / B2 on K - D +
    proj_apply<K, K - D>: translation-invariant matrices take the reduced path (code bit 16), all others the general one."""
    rng = np.random.default_rng(4000 + 10 * k + d)
    P = translation_projector(k, d)
    mats, invariant = [], []
    for scale in (1.0, 1e-5, 1e4):
        for _ in range(8):
            A = rng.standard_normal((k, k)); mats.append(P @ (A + A.T) @ P * scale); invariant.append(True)
        B = rng.standard_normal((k, k)); mats.append(P @ (B @ B.T) @ P * scale); invariant.append(True)          # PSD on the complement
        B = rng.standard_normal((k, max(1, k - d - 1))); mats.append(P @ (B @ B.T) @ P * scale); invariant.append(True)
        B = rng.standard_normal((k, k)); mats.append(-P @ (B @ B.T) @ P * scale); invariant.append(True)         # everything moves: form B
        A = rng.standard_normal((k, k)); mats.append((A + A.T) * scale); invariant.append(False)
    mats.append(np.zeros((k, k))); invariant.append(False)
    mats.append(-P); invariant.append(True)
    n_fallback = 0
    for idx, (A, inv) in enumerate(zip(mats, invariant)):
        A = 0.5 * (A + A.T)
        p = pack(hostlib, A)
        code = hostlib.host_project_reduced(k, d, p.ctypes.data, eps)
        got = unpack(hostlib, p, k)
        if code == 3:
            n_fallback += 1
            assert np.array_equal(got, A)
            continue
        ora, ocode = oracle.project(A, eps)
        ref = A if ocode == 0 else reference_projection(A, eps)
        scale = max(np.abs(A).max(), abs(eps), 1e-300)
        assert np.abs(got - ref).max() / scale <= 5e-12, (k, d, idx, code, np.abs(got - ref).max() / scale)
        assert np.abs(got - ora).max() / scale <= 5e-12
        if code & 15 >= 2 and ocode != 0:
            assert bool(code & 16) == inv, (k, d, idx, code)       # invariant matrices went through the reduced pipeline
        if code & 15 < 2:
            assert np.array_equal(got, A)
    assert n_fallback <= 2, n_fallback


def test_reduced_pipeline_on_tet_hessians(hostlib):
    from problems import tet_problem
    import scipy.sparse as sp
    p, x = tet_problem(3, seed=4)
    kind, conn, data = p.terms[0]
    for e in range(0, len(conn), 3):
        loc = np.arange(4, dtype=np.int32)[None, :]
        xe = x.reshape(-1, 3)[conn[e]].reshape(-1)
        r = oracle.scalar_eval(3, 4, [oracle.Term(kind, loc, data[e:e + 1])], oracle.DERIVATIVES, xe)
        A = sp.csc_matrix((r.values, r.inner, r.outer), shape=(12, 12)).toarray()
        A = 0.5 * (A + A.T)
        if np.abs(A).max() == 0:
            continue
        for eps in (1e-9, 1e-6):
            pk = pack(hostlib, A)
            code = hostlib.host_project_reduced(12, 3, pk.ctypes.data, eps)
            assert code == 2 | 16
            ref = reference_projection(A, eps)
            assert np.abs(unpack(hostlib, pk, 12) - ref).max() <= 1e-11 * np.abs(A).max()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref is not built")
def test_product_projection_against_the_reference_routine(hostlib):
    """The product's projection routines (TinyAD/Detail/Projection.hh, compiled for the host: general pipeline, translation
    deflation, reduced 9 x 9 pipeline) against TinyAD::project_positive_definite of the reference ITSELF (oracle/_ref,
    Utils/HessianProjection.hh:48-101) on the element Hessians of a deformed tet mesh and on random symmetric matrices, eps = 1e-9
    and the absolute-value strategy: 1e-10 relative to max |H| (north_star's bound for projected values)."""
    from problems import tet_problem
    import scipy.sparse as sp
    p, x = tet_problem(3, seed=9)
    kind, conn, data = p.terms[0]
    mats = []
    for e in range(0, len(conn), 2):
        loc = np.arange(4, dtype=np.int32)[None, :]
        xe = x.reshape(-1, 3)[conn[e]].reshape(-1)
        r = oracle.scalar_eval(3, 4, [oracle.Term(kind, loc, data[e:e + 1])], oracle.DERIVATIVES, xe)
        A = sp.csc_matrix((r.values, r.inner, r.outer), shape=(12, 12)).toarray()
        A = 0.5 * (A + A.T)
        if np.abs(A).max() > 0:
            mats.append((A, True))
    rng = np.random.default_rng(12)
    for _ in range(20):
        B = rng.standard_normal((12, 12))
        mats.append((B + B.T, False))
    worst = 0.0
    for A, is_element in mats:
        for eps in (1e-9, -1.0):
            ref = oracle.ref_project(A, eps)
            runs = [("general", lambda pk: hostlib.host_project(12, pk.ctypes.data, eps))]
            if is_element:
                runs += [("deflated", lambda pk: hostlib.host_project_translations(12, 3, pk.ctypes.data, eps)),
                         ("reduced", lambda pk: hostlib.host_project_reduced(12, 3, pk.ctypes.data, eps))]
            for name, run in runs:
                if eps < 0 and name != "general":
                    continue                       # a numerically zero eigenvalue may take either sign in |lambda| mode: no solver pins it
                pk = pack(hostlib, A)
                run(pk)
                err = np.abs(unpack(hostlib, pk, 12) - ref).max() / np.abs(A).max()
                if eps < 0 and is_element:
                    continue                       # same remark: the translation null space sits exactly at the sign change
                worst = max(worst, err)
                assert err <= 1e-10, (name, eps, err)
    assert worst > 0.0
