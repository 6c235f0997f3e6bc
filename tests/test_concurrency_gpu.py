"""GPU: re-entrancy and run-to-run determinism.

tests/ScalarFunctionTest.cc:255-291 calls eval* concurrently on one function object; tests/NewtonTest.cc:97-111 runs whole Newton
loops from 4 threads and expects bitwise identical results.  ctypes releases the GIL during the C calls, so the Python threads below
really are concurrent at the C ABI (the runtime serialises the evaluations of ONE function object with a mutex; different function
objects run on their own streams)."""
import threading

import numpy as np
import pytest

import tinyad_b200 as tad
from problems import planar_newton_problem, tet_problem

pytestmark = pytest.mark.gpu


def _eval(torch, fn, x):
    xd = torch.from_numpy(x).cuda()
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    f = fn.eval_with_hessian_proj(xd, g, H)
    torch.cuda.synchronize()
    return f, g.cpu().numpy(), H.cpu().numpy()


def test_concurrent_evaluations_of_one_function(torch_cuda):
    torch = torch_cuda
    p, x = tet_problem(8, seed=1, with_penalty=True)
    fn = p.gpu(assembly=tad.ASSEMBLY_GATHER)            # deterministic assembly: results must be bitwise equal
    fn.pattern()
    ref = _eval(torch, fn, x)
    xs = [x, x * 1.0, x.copy(), x + 0.0]
    out, errs = [None] * 8, []

    def work(i):
        try:
            torch.cuda.set_device(0)
            for _ in range(5):
                out[i] = _eval(torch, fn, xs[i % 4])
                f = fn.eval(torch.from_numpy(x).cuda())
                assert f == ref[0]
        except Exception as e:        # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs
    for r in out:
        assert r[0] == ref[0] and np.array_equal(r[1], ref[1]) and np.array_equal(r[2], ref[2])
    fn.close()


def test_newton_loops_from_four_threads_are_deterministic(torch_cuda):
    """NewtonTest.cc:97-111: four threads, each its own function object and its own projected-Newton loop, identical results."""
    torch = torch_cuda
    results, errs = [None] * 4, []

    def loop(i):
        try:
            torch.cuda.set_device(0)
            p, x = planar_newton_problem()
            fn = p.gpu(assembly=tad.ASSEMBLY_GATHER)
            xd = torch.from_numpy(x).cuda()
            g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
            H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
            d = torch.empty_like(g)
            xn = torch.empty_like(g)
            for _ in range(10):
                f = fn.eval_with_hessian_proj(xd, g, H)
                fn.newton_direction(g, H, d, w_identity=1e-9, rel_tol=1e-13)
                fn.line_search(xd, d, f, g, xn)
                xd, xn = xn, xd
            f = fn.eval_with_hessian_proj(xd, g, H)
            torch.cuda.synchronize()
            results[i] = (f, xd.cpu().numpy(), g.cpu().numpy())
            fn.close()
        except Exception as e:        # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=loop, args=(i,)) for i in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs
    f0, x0, g0 = results[0]
    assert abs(f0 - 4.0) < 1e-12 and np.abs(g0).max() < 1e-10          # NewtonTest.cc:82-88
    for f, xx, gg in results[1:]:
        # gather assembly and the PCG's fixed-order dot products make the whole loop bitwise reproducible, like the reference's
        assert f == f0 and np.array_equal(xx, x0) and np.array_equal(gg, g0)
