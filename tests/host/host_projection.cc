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
