// tinyad_b200 -- LinearSolver of the reference's solver utilities (Utils/LinearSolver.hh:12-19).
// The reference wraps Eigen::SimplicialLDLT and caches its symbolic analysis (`sparsity_pattern_dirty`).  On the B200 path
// the pattern is fixed by construction and the solve is a device PCG (tad_pcg_solve), so the struct only carries the
// solver's parameters and the statistics of the last solve.
#pragma once

namespace TinyAD
{

template <typename PassiveT = double>
struct LinearSolver
{
    bool sparsity_pattern_dirty = true;  // kept for source compatibility; the device pattern never changes
    double rel_tol = 1e-10;              // relative residual |A d + g| / |g| at which the PCG stops
    int max_iters = 10000;
    int block_dim = 1;                   // size of the block-Jacobi blocks (the variable dimension d); 1 = scalar Jacobi
    int last_iters = 0;
    double last_rel_residual = 0.0;
};

}  // namespace TinyAD
