"""GPU: re-entrancy and run-to-run determinism.

tests/ScalarFunctionTest.cc:255-291 calls eval* concurrently on one function object; tests/NewtonTest.cc:97-111 runs whole Newton
loops from 4 threads and expects bitwise identical results.  ctypes releases the GIL during the C calls, so the Python threads below
really are concurrent at the C ABI (the runtime serialises the evaluations of ONE function object with a mutex; different function
objects run on their own streams)."""
import threading

import os

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


def test_concurrent_host_buffer_calls_on_one_function(torch_cuda):
    """The host-vector overloads of the facade (tad_eval*_host) stage x / g / H through device buffers owned by the function: eight
    threads with DIFFERENT x must each get the result of their own x (the scenario of tests/ScalarFunctionTest.cc:255-291 with the
    std::vector overloads; the runtime holds the function mutex from the H2D of x to the last D2H)."""
    torch = torch_cuda
    p, x = tet_problem(7, seed=4, with_penalty=True)
    fn = p.gpu(assembly=tad.ASSEMBLY_GATHER)            # deterministic assembly: bitwise comparable
    fn.set_option(tad.OPT_CHUNK_ELEMENTS, 512)         # several slabs, pipelined D2H of finished rows
    rng = np.random.default_rng(0)
    xs = [x + 1e-3 * rng.standard_normal(x.shape) for _ in range(8)]
    refs = [fn.eval_with_hessian_proj_host(xi) for xi in xs]            # serial
    assert len({r[0] for r in refs}) == 8                                # the inputs really differ
    out, errs = [None] * 8, []

    def work(i):
        try:
            torch.cuda.set_device(0)
            for _ in range(6):
                out[i] = fn.eval_with_hessian_proj_host(xs[i])
                f, g = fn.eval_with_gradient_host(xs[i])
                assert f == refs[i][0] and fn.eval_host(xs[i]) == refs[i][0]
        except Exception as e:        # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs
    for r, ref in zip(out, refs):
        assert r[0] == ref[0] and np.array_equal(r[1], ref[1]) and np.array_equal(r[2], ref[2])
    fn.close()


@pytest.mark.parametrize("assembly", [tad.ASSEMBLY_ATOMIC, tad.ASSEMBLY_GATHER])
def test_slab_size_does_not_change_results(torch_cuda, assembly):
    """TAD_OPT_CHUNK_ELEMENTS / TAD_OPT_LANES: slab-wise evaluation with pipelined D2H gives the same f (bitwise: the partial sums are
    laid out independently of the slab size), g and H (atomic: to rounding; gather: bitwise) as one slab per term."""
    torch = torch_cuda
    p, x = tet_problem(9, seed=5, with_penalty=True)     # 4,374 tets + 3 penalty elements (which touch the first and the last vertex)
    results = []
    for chunk, lanes in ((-1, 1), (256, 1), (512, 2), (1024, 3), (1536, 4)):
        fn = p.gpu(assembly=assembly)
        fn.set_option(tad.OPT_CHUNK_ELEMENTS, chunk)
        fn.set_option(tad.OPT_LANES, lanes)
        results.append(fn.eval_with_hessian_proj_host(x))
        f1, g1 = fn.eval_with_gradient_host(x)
        assert f1 == results[-1][0] and fn.eval_host(x) == f1
        st = fn.projection_stats()
        assert st["decomposed"] == 4374 + 3 - 3 or st["decomposed"] >= 4374   # the 1-vertex penalty blocks are diagonally dominant
        fn.close()
    f0, g0, H0 = results[0]
    for f, g, H in results[1:]:
        assert f == f0
        if assembly == tad.ASSEMBLY_GATHER:
            assert np.array_equal(g, g0) and np.array_equal(H, H0)
        else:
            assert np.abs(g - g0).max() <= 1e-13 * np.abs(g0).max() and np.abs(H - H0).max() <= 1e-13 * np.abs(H0).max()


def test_launch_count_and_caller_stream(torch_cuda):
    """tad_function_launch_count counts this library's kernels and the element kernels; tad_function_set_caller_stream orders an
    evaluation after work queued on the caller's stream (here: x is produced by a torch kernel on a side stream)."""
    torch = torch_cuda
    p, x = tet_problem(6, seed=6)
    fn = p.gpu()
    fn.set_option(tad.OPT_CHUNK_ELEMENTS, -1)
    g = torch.empty(fn.n_vars, dtype=torch.float64, device="cuda")
    H = torch.empty(fn.nnz, dtype=torch.float64, device="cuda")
    xd = torch.from_numpy(x).cuda()
    torch.cuda.synchronize()
    n0 = fn.launch_count()
    f_ref = fn.eval_with_hessian_proj(xd, g, H)
    n1 = fn.launch_count()
    # 4 element kernels (Hessian parts) + 2 reductions + A, B1, B2, list + 2 fused phase-C / assembly kernels = 12; with the reduced
    # pipeline of four-handle elements (default): B1 and B2 once per pipeline and one more fused phase-C / assembly launch = 15
    assert n1 - n0 == (12 if os.environ.get("TAD_REDUCED_PIPELINE", "1") == "0" else 15)
    side = torch.cuda.Stream()
    fn.set_caller_stream(side.cuda_stream)
    big = torch.randn(1 << 24, device="cuda")
    with torch.cuda.stream(side):
        for _ in range(20):
            big = big * 1.0001                      # keeps the side stream busy for a while
        x2 = torch.from_numpy(x).cuda(non_blocking=False) * 1.0
        x2 = x2 + 0.0 * big[: x2.numel()].double()  # depends on the long chain
    f2 = fn.eval_with_hessian_proj(x2, g, H)        # no host sync in between: the runtime waits for `side`
    assert f2 == f_ref
    fn.close()
