// tinyad_b200 -- newton_direction (Utils/NewtonDirection.hh:25-48, :56-65): d with (H_proj + w_identity I) d = -g.
// Reference: Eigen::SimplicialLDLT factorisation.  Here: device PCG with a block-Jacobi preconditioner (tad_pcg_solve); the
// sparse direct solve named by BASELINE.json (cuDSS) is not available in this image -- this is a labelled stand-in and is
// timed separately from the assembly.  Throws std::runtime_error("Linear solve failed ...") like the reference (:43-44).
#pragma once

#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/Utils/LinearSolver.hh>

namespace TinyAD
{

// device-resident g / H values / d of a function (no PCIe traffic inside a Newton loop)
template <class FunctionT>
void newton_direction_device(const FunctionT& _func, const double* _g_dev, const double* _H_values_dev, LinearSolver<double>& _solver,
                             double* _d_dev, const double _w_identity = 0.0)
{
    _solver.sparsity_pattern_dirty = false;
    detail::check(tad_newton_direction(_func.handle(), _g_dev, _H_values_dev, _w_identity, _solver.rel_tol, _solver.max_iters, _d_dev,
                                       &_solver.last_iters, &_solver.last_rel_residual));
}

// host vectors + host CSR (the reference's signature with std::vector / TinyAD::SparseMatrix instead of Eigen types)
inline std::vector<double> newton_direction(const std::vector<double>& _g, const SparseMatrix& _H_proj, LinearSolver<double>& _solver,
                                            const double& _w_identity = 0.0)
{
    const int64_t n = (int64_t)_g.size();
    if ((int64_t)_H_proj.outer.size() != n + 1) throw std::runtime_error("[TinyAD-B200] newton_direction: size mismatch");
    struct Dev
    {
        void* p = nullptr;
        ~Dev() { if (p) cudaFree(p); }
    } outer, inner, vals, g, d;
    auto up = [](Dev& b, const void* src, size_t bytes) {
        if (cudaMalloc(&b.p, bytes ? bytes : 8) != cudaSuccess || (bytes && cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess))
            throw std::runtime_error("[TinyAD-B200] newton_direction: device allocation / copy failed");
    };
    up(outer, _H_proj.outer.data(), _H_proj.outer.size() * sizeof(int32_t));
    up(inner, _H_proj.inner.data(), _H_proj.inner.size() * sizeof(int32_t));
    up(vals, _H_proj.values.data(), _H_proj.values.size() * sizeof(double));
    up(g, _g.data(), _g.size() * sizeof(double));
    up(d, nullptr, 0);
    if (n && (cudaFree(d.p) != cudaSuccess || (d.p = nullptr, cudaMalloc(&d.p, (size_t)n * sizeof(double)) != cudaSuccess)))
        throw std::runtime_error("[TinyAD-B200] newton_direction: device allocation failed");
    _solver.sparsity_pattern_dirty = false;
    detail::check(tad_pcg_solve(n, _solver.block_dim, (const int32_t*)outer.p, (const int32_t*)inner.p, (const double*)vals.p, _w_identity,
                                (const double*)g.p, -1.0, (double*)d.p, _solver.rel_tol, _solver.max_iters, &_solver.last_iters,
                                &_solver.last_rel_residual, nullptr));
    std::vector<double> out((size_t)n);
    if (n && cudaMemcpy(out.data(), d.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        throw std::runtime_error("[TinyAD-B200] newton_direction: copy failed");
    for (double v : out)
        if (!std::isfinite(v)) throw std::runtime_error("[TinyAD-B200] Linear solve failed: direction is not finite.");  // :46
    return out;
}

inline std::vector<double> newton_direction(const std::vector<double>& _g, const SparseMatrix& _H_proj, const double& _w_identity = 0.0)
{
    LinearSolver<double> solver;  // :56-65
    return newton_direction(_g, _H_proj, solver, _w_identity);
}

}  // namespace TinyAD
