// Host build of the device projection routine (tinyad_b200/include/TinyAD/Detail/Projection.hh) for CPU unit tests:
// the function is __host__ __device__, so this exercises exactly the code the CUDA kernel runs.
#include <TinyAD/Detail/Projection.hh>

template <int K>
static int run(double* packed, double eps)
{
    return TinyAD::detail::project_element<K>([&](int s) { return packed[s]; }, [&](int s, double v) { packed[s] = v; }, eps);
}

extern "C" int host_project(int k, double* packed, double eps)
{
    switch (k)
    {
    case 1: return run<1>(packed, eps);
    case 2: return run<2>(packed, eps);
    case 3: return run<3>(packed, eps);
    case 4: return run<4>(packed, eps);
    case 5: return run<5>(packed, eps);
    case 6: return run<6>(packed, eps);
    case 7: return run<7>(packed, eps);
    case 8: return run<8>(packed, eps);
    case 9: return run<9>(packed, eps);
    case 10: return run<10>(packed, eps);
    case 12: return run<12>(packed, eps);
    case 15: return run<15>(packed, eps);
    case 16: return run<16>(packed, eps);
    case 18: return run<18>(packed, eps);
    default: return -1;
    }
}

// element Hessians with the translation null space of D-dimensional handles: deflated form (Projection.hh proj_tridiagonalize<K, D>)
extern "C" int host_project_translations(int k, int d, double* packed, double eps)
{
    auto ld = [&](int s) { return packed[s]; };
    auto st = [&](int s, double v) { packed[s] = v; };
    if (k == 12 && d == 3) return TinyAD::detail::project_element<12, 3>(ld, st, eps);
    if (k == 9 && d == 3) return TinyAD::detail::project_element<9, 3>(ld, st, eps);
    if (k == 6 && d == 3) return TinyAD::detail::project_element<6, 3>(ld, st, eps);
    if (k == 6 && d == 2) return TinyAD::detail::project_element<6, 2>(ld, st, eps);
    if (k == 8 && d == 2) return TinyAD::detail::project_element<8, 2>(ld, st, eps);
    if (k == 4 && d == 2) return TinyAD::detail::project_element<4, 2>(ld, st, eps);
    return -1;
}

// the reduced pipeline (four handles: the complement of the translations through a Hadamard transform, K - D instead of K)
extern "C" int host_project_reduced(int k, int d, double* packed, double eps)
{
    auto ld = [&](int s) { return packed[s]; };
    auto st = [&](int s, double v) { packed[s] = v; };
    if (k == 12 && d == 3) return TinyAD::detail::project_element_reduced<12, 3>(ld, st, eps);
    if (k == 8 && d == 2) return TinyAD::detail::project_element_reduced<8, 2>(ld, st, eps);
    if (k == 4 && d == 1) return TinyAD::detail::project_element_reduced<4, 1>(ld, st, eps);
    return -1;
}

// dimension of the translation null space phase B2 deflated (0 = the plain path was taken), for the tests
template <int K, int D>
static int deflated_dim(const double* packed, double eps)
{
    using namespace TinyAD::detail;
    using L = ProjLayout<K>;
    double R[L::nR], Wb[L::nW] = {0};
    int code = proj_tridiagonalize<K, D>([&](int s) { return packed[s]; }, [&](int i, double v) { R[i] = v; }, eps);
    if (code == PROJ_DOMINANT) return -1;
    if (proj_eigenvalues<K>([&](int i) { return R[i]; }, [&](int i, double v) { R[i] = v; }) == PROJ_FALLBACK) return -2;
    code = proj_select_vectors<K>([&](int i) { return R[i]; }, [&](int i) { return R[L::off_lam + i]; }, [&](int i, double v) { R[L::off_lam + i] = v; },
                                  [&](int i, double v) { Wb[i] = v; }, [&](int jv, const double (&v)[K]) { for (int q = 0; q < K; ++q) Wb[L::off_vec + jv * K + q] = v[q]; },
                                  [&](int jv, double (&v)[K]) { for (int q = 0; q < K; ++q) v[q] = Wb[L::off_vec + jv * K + q]; }, eps);
    if (code != PROJ_REBUILT) return -3;
    return ((int)Wb[1]) >> 3;
}
extern "C" int host_deflated_dim(int k, int d, const double* packed, double eps)
{
    if (k == 12 && d == 3) return deflated_dim<12, 3>(packed, eps);
    if (k == 6 && d == 2) return deflated_dim<6, 2>(packed, eps);
    return -9;
}

// the in-kernel fallback of the fused small-k element kernel (cyclic Jacobi, Projection.hh project_full_jacobi)
template <int K>
static int run_jacobi(double* packed, double eps)
{
    return TinyAD::detail::project_full_jacobi<K>([&](int s) { return packed[s]; }, [&](int s, double v) { packed[s] = v; }, eps);
}

extern "C" int host_project_jacobi(int k, double* packed, double eps)
{
    switch (k)
    {
    case 1: return run_jacobi<1>(packed, eps);
    case 2: return run_jacobi<2>(packed, eps);
    case 3: return run_jacobi<3>(packed, eps);
    case 4: return run_jacobi<4>(packed, eps);
    case 5: return run_jacobi<5>(packed, eps);
    case 6: return run_jacobi<6>(packed, eps);
    case 8: return run_jacobi<8>(packed, eps);
    default: return -1;
    }
}

extern "C" int host_seq_index(int k, int i, int j) { return TinyAD::detail::hess_seq_index(k, i, j); }
