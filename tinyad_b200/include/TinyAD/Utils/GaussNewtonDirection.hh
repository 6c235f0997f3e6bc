// tinyad_b200 -- gauss_newton_direction (Utils/GaussNewtonDirection.hh:24-47): d = -(J^T J + w_identity I)^-1 J^T r for a
// sum-of-squares function f(x) = sum_i r_i(x)^2.  The reference forms J^T J as a sparse product and factorises it with
// SimplicialLDLT; here the normal-equations operator is applied matrix-free inside a device PCG (tad_gauss_newton_direction):
// r, the Jacobian values and d stay in HBM.  Throws std::runtime_error("Linear solve failed ...") like the reference (:41-42).
#pragma once

#include <TinyAD/ScalarFunction.hh>
#include <TinyAD/Utils/LinearSolver.hh>

namespace TinyAD
{

// _func: the VectorFunction whose eval_sum_of_squares_with_derivatives produced r / J values (device pointers)
template <class VectorFunctionT>
void gauss_newton_direction_device(const VectorFunctionT& _func, const double* _r_dev, const double* _J_values_dev, LinearSolver<double>& _solver,
                                   double* _d_dev, const double _w_identity = 0.0)
{
    _solver.sparsity_pattern_dirty = false;
    detail::check(tad_gauss_newton_direction(_func.handle(), _r_dev, _J_values_dev, _w_identity, _solver.rel_tol, _solver.max_iters, _d_dev,
                                             &_solver.last_iters, &_solver.last_rel_residual));
}

}  // namespace TinyAD
