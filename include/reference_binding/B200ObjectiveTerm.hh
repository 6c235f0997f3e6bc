// The binding a TinyAD maintainer would add: a ScalarObjectiveTermBase -- the reference's own
// plugin interface for the path, include/TinyAD/Detail/ScalarObjectiveTerm.hh:21-45 -- whose three virtuals run on the B200
// through the C ABI of include/tinyad_b200.h.  Host-only C++17: it includes the REFERENCE's headers (and therefore Eigen) and
// nothing of this repository except the C header, so it lives in the reference's translation units; the element functors are
// compiled for the device in a separate .cu translation unit (tinyad_b200/include/TinyAD/Kernels.cuh), which hands over the
// tad_function they were added to.
//
//     auto func = TinyAD::scalar_function<3>(TinyAD::range(n_vertices));                  // the reference's facade, unmodified
//     func.objective_terms.push_back(std::make_unique<TinyAD::B200ScalarObjectiveTerm>(fn, n_elements));
//     func.n_elements += n_elements;
//     auto [f, g, H_proj] = func.eval_with_hessian_proj(x);                               // Detail/ScalarFunctionImpl.hh:378-416
//
// compiled and run as written by oracle/ref_plugin_driver.cc (tests/test_reference_plugin*.py).
#pragma once

#include <TinyAD/Detail/ScalarObjectiveTerm.hh>  // the reference

#include <tinyad_b200.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace TinyAD
{

struct B200ScalarObjectiveTerm : ScalarObjectiveTermBase<double>
{
    /// _fn holds the device terms (one or several tad_function_add_term calls); it is borrowed, not owned.
    B200ScalarObjectiveTerm(tad_function _fn, Eigen::Index _n_elements) : fn(_fn), n(_n_elements)
    {
        if (fn == nullptr) throw std::runtime_error("B200ScalarObjectiveTerm: null tad_function");
        if (tad_function_n_elements(fn) != (int64_t)_n_elements) throw std::runtime_error("B200ScalarObjectiveTerm: element count differs from the device function's");
    }

    Eigen::Index n_elements() const override { return n; }

    // Detail/ScalarObjectiveTerm.hh:162-187
    double eval(const Eigen::VectorX<double>& _x) const override
    {
        check_size(_x);
        double f = 0.0;
        check(tad_eval_host(fn, _x.data(), &f));
        return f;
    }

    // Detail/ScalarObjectiveTerm.hh:189-222
    void eval_with_gradient_add(const Eigen::VectorX<double>& _x, double& _f, Eigen::VectorX<double>& _g) const override
    {
        check_size(_x);
        double f = 0.0;
        std::vector<double> g((std::size_t)_x.size());
        check(tad_eval_with_gradient_host(fn, _x.data(), &f, g.data()));
        _f += f;
        for (Eigen::Index i = 0; i < _g.size(); ++i) _g[i] += g[(std::size_t)i];
    }

    // Detail/ScalarObjectiveTerm.hh:224-278: the assembled rows go back as one triplet per structural entry (explicit zeros
    // included), so that the caller's setFromTriplets reproduces the pattern it would have built from the per-element triplets
    void eval_with_derivatives_add(const Eigen::VectorX<double>& _x, double& _f, Eigen::VectorX<double>& _g,
                                   std::vector<Eigen::Triplet<double>>& _H_triplets, const bool _project_hessian,
                                   const double& _projection_eps) const override
    {
        check_size(_x);
        int64_t n_outer = 0, nnz = 0;
        check(tad_function_pattern(fn, &n_outer, &nnz));
        std::vector<int32_t> outer((std::size_t)n_outer + 1), inner((std::size_t)nnz);
        check(tad_function_pattern_copy(fn, outer.data(), inner.data()));
        std::vector<double> g((std::size_t)_x.size()), values((std::size_t)nnz);
        double f = 0.0;
        check(tad_eval_with_derivatives_host(fn, _x.data(), &f, g.data(), values.data(), _project_hessian ? 1 : 0, _projection_eps));
        _f += f;
        for (Eigen::Index i = 0; i < _g.size(); ++i) _g[i] += g[(std::size_t)i];
        using SparseIndex = typename Eigen::SparseMatrix<double>::StorageIndex;
        _H_triplets.reserve(_H_triplets.size() + (std::size_t)nnz);
        for (int64_t r = 0; r < n_outer; ++r)
            for (int32_t p = outer[(std::size_t)r]; p < outer[(std::size_t)r + 1]; ++p)
                _H_triplets.push_back(Eigen::Triplet<double>((SparseIndex)r, (SparseIndex)inner[(std::size_t)p], values[(std::size_t)p]));
    }

private:
    void check_size(const Eigen::VectorX<double>& _x) const
    {
        if ((int64_t)_x.size() != tad_function_n_vars(fn)) throw std::runtime_error("B200ScalarObjectiveTerm: x has the wrong size");
    }
    static void check(int status)
    {
        if (status != TAD_OK) throw std::runtime_error(std::string("[TinyAD-B200] ") + tad_last_error());
    }

    tad_function fn;
    Eigen::Index n;
};

}  // namespace TinyAD
