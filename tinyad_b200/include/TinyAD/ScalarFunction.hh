// tinyad_b200 -- host facade with the reference's public surface
// (include/TinyAD/ScalarFunction.hh:36-248, Detail/ScalarFunctionImpl.hh in /root/reference):
//   auto func = TinyAD::scalar_function<d>(TinyAD::range(n));
//   func.add_elements<N>(TinyAD::range(m), functor);
//   auto [f, g, H] = func.eval_with_hessian_proj(x);
// Every evaluation goes through the C ABI (include/tinyad_b200.h) into CUDA kernels; there is
// no CPU path.  Errors reported by the runtime become std::runtime_error, like
// TINYAD_ERROR_throw (Utils/Out.hh:73-79), and leave the function usable.
//
// Differences from the reference that the GPU imposes:
//   * the element function is a functor with `template <class E> __host__ __device__ auto
//     operator()(E& element) const` (nvcc rejects generic __host__ __device__ lambdas); its
//     captures are copied to the device by value, so per-element data must be device pointers;
//   * x / g are std::vector<double> (or raw device pointers in the *_device overloads) and H is
//     the plain CSR struct below instead of Eigen::VectorXd / Eigen::SparseMatrix<double>;
//   * handles are integers (Support/Common.hh:30); mesh-library adapters are out of scope.
#pragma once

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <map>
#include <tuple>
#include <utility>
#include <vector>

#include <TinyAD/Kernels.cuh>
#include <TinyAD/Support/Common.hh>
#include <tinyad_b200.h>

#define TINYAD_ScalarFunction_DEFINED

namespace TinyAD
{

// Detail/EvalSettings.hh:14-21.  n_threads is meaningless on the GPU and kept for source compatibility.
struct EvalSettings
{
    int n_threads = -1;
    int device = 0;                       // CUDA device ordinal
    int assembly = TAD_ASSEMBLY_ATOMIC;   // or TAD_ASSEMBLY_GATHER (deterministic, element order)
    int64_t chunk_elements = 0;           // 0 = whole term per launch
};

// Utils/HessianProjection.hh:16
constexpr double default_hessian_projection_eps = 1e-9;

// Utils/Helpers.hh:17-27
inline std::vector<int64_t> range(int64_t n)
{
    std::vector<int64_t> r((size_t)n);
    for (int64_t i = 0; i < n; ++i) r[(size_t)i] = i;
    return r;
}
template <class Range>
int64_t count(const Range& r) { int64_t n = 0; for (auto it = r.begin(); it != r.end(); ++it) ++n; return n; }

// Compressed sparse matrix with int32 indices (Eigen::SparseMatrix<double>'s StorageIndex).  For the
// Hessian the arrays are both CSR and CSC (structural symmetry); for the Jacobian they are CSC.
struct SparseMatrix
{
    int64_t rows = 0, cols = 0;
    std::vector<int32_t> outer;   // outerIndexPtr
    std::vector<int32_t> inner;   // innerIndexPtr, ascending per outer slice
    std::vector<double> values;   // valuePtr (explicit zeros kept)
    int64_t nonZeros() const { return (int64_t)values.size(); }
    double coeff(int64_t r, int64_t c) const  // column-compressed lookup (== row-compressed for symmetric H)
    {
        for (int32_t p = outer[(size_t)c]; p < outer[(size_t)c + 1]; ++p)
            if (inner[(size_t)p] == r) return values[(size_t)p];
        return 0.0;
    }
};

namespace detail
{
inline void check(int status)
{
    if (status != TAD_OK) throw std::runtime_error(std::string("[TinyAD-B200] ") + tad_last_error());
}
template <class Range>
bool is_identity_range(const Range& r, std::vector<int64_t>& out)
{
    bool identity = true;
    int64_t i = 0;
    for (auto h : r) { const int64_t idx = idx_from_handle(h); out.push_back(idx); identity = identity && (idx == i); ++i; }
    return identity;
}
}  // namespace detail

template <int variable_dimension, typename PassiveT = double, typename VariableHandleT = int64_t>
struct ScalarFunction
{
    static_assert(std::is_same<PassiveT, double>::value, "the B200 path computes in FP64 only");
    static_assert(variable_dimension >= 1, "Variable dimension needs to be at least 1.");

    ScalarFunction() = default;  // evaluates to 0 / empty like a default-constructed reference function
    ScalarFunction(std::vector<VariableHandleT> _variable_handles, const EvalSettings& _settings = EvalSettings())
        : settings(_settings), n_vars(variable_dimension * (int64_t)_variable_handles.size()),
          variable_handles(std::move(_variable_handles))
    {
        // ScalarFunctionImpl.hh:57-60: handles must be a permutation of 0..n-1
        std::vector<bool> used(variable_handles.size(), false);
        for (auto v : variable_handles)
        {
            const int64_t iv = idx_from_handle(v);
            if (iv < 0 || iv >= (int64_t)variable_handles.size() || used[(size_t)iv])
                throw std::runtime_error("[TinyAD-B200] variable indices are not compact");
            used[(size_t)iv] = true;
        }
        if (variable_handles.empty()) throw std::runtime_error("[TinyAD-B200] no variables");
        detail::check(tad_function_create(variable_dimension, (int64_t)variable_handles.size(), 0, settings.device, &h));
        detail::check(tad_function_set_option(h, TAD_OPT_ASSEMBLY, settings.assembly));
        detail::check(tad_function_set_option(h, TAD_OPT_CHUNK_ELEMENTS, settings.chunk_elements));
    }
    ScalarFunction(const ScalarFunction&) = delete;             // ScalarFunction.hh:46-51
    ScalarFunction& operator=(const ScalarFunction&) = delete;
    ScalarFunction(ScalarFunction&& o) noexcept { *this = std::move(o); }
    ScalarFunction& operator=(ScalarFunction&& o) noexcept
    {
        if (this != &o)
        {
            if (h) tad_function_destroy(h);
            h = o.h; o.h = nullptr;
            settings = o.settings;
            n_vars = o.n_vars; n_elements = o.n_elements;
            variable_handles = std::move(o.variable_handles);
            o.n_vars = 0; o.n_elements = 0;
        }
        return *this;
    }
    ~ScalarFunction() { if (h) tad_function_destroy(h); }

    // ScalarFunctionImpl.hh:63-100.  The functor is stored by value (LambdaImpl::func) and lives on as kernel argument.
    template <int element_valence, typename ElementHandleRangeT, typename EvalElementFunction>
    void add_elements(const ElementHandleRangeT& _element_range, EvalElementFunction _eval_element)
    {
        static_assert(element_valence >= 0, "Element valence needs to be non-negative.");
        require_handle();
        std::vector<int64_t> handles;
        const bool identity = detail::is_identity_range(_element_range, handles);
        add_term_impl<element_valence, false>(handles, identity, std::move(_eval_element));
    }

    // ScalarFunction.hh:80-108 / ScalarFunctionImpl.hh:134-214: each element may touch a different number of variable handles.
    // The functor is run once on a recorder element (on the device, with the largest static valence) to find each element's
    // valence; elements are grouped by the exact or next larger static valence and one term per non-empty group is added, in
    // the order of the template arguments -- exactly the reference's grouping.  Within a group the slot of a variables() call
    // is a run-time quantity, so these terms use the run-time-indexed element (dense derivative masks).  As in the reference,
    // the projection acts on the padded k x k block while only the accessed variables are assembled.
    template <int... ElementValences, typename ElementHandleRangeT, typename EvalElementFunction>
    void add_elements_dynamic(const ElementHandleRangeT& _element_range, EvalElementFunction _eval_element)
    {
        static_assert(sizeof...(ElementValences) >= 1, "At least one element valence has to be passed.");
        static_assert(((ElementValences >= 0) && ...), "Element valences need to be non-negative.");
        require_handle();
        std::vector<int> static_valences_sorted = {ElementValences...};
        std::sort(static_valences_sorted.begin(), static_valences_sorted.end());
        if (std::unique(static_valences_sorted.begin(), static_valences_sorted.end()) != static_valences_sorted.end())
            throw std::runtime_error("[TinyAD-B200] Element valences passed to add_elements<..>(..) are not unique. Please pass unique element valences.");
        constexpr int max_valence = std::max({ElementValences...});
        std::vector<int64_t> handles;
        const bool identity = detail::is_identity_range(_element_range, handles);
        const int64_t n = (int64_t)handles.size();

        // record: a scratch function with one term of the largest valence; its element -> handle table gives the valences
        std::vector<int32_t> table((size_t)std::max<int64_t>(1, max_valence * n), -1);
        {
            tad_function rec = nullptr;
            detail::check(tad_function_create(variable_dimension, (int64_t)variable_handles.size(), 0, settings.device, &rec));
            using R = TermLauncher<EvalElementFunction, variable_dimension, max_valence, 0, true>;
            R* launcher = new R{_eval_element};
            const int s = tad_function_add_term(rec, max_valence, 0, n, identity ? nullptr : handles.data(), &R::launch, launcher, &R::destroy);
            if (s == TAD_OK && n > 0) detail::check(tad_function_term_table(rec, 0, table.data()));
            tad_function_destroy(rec);
            if (s == TAD_TOO_MANY_VARIABLES)
                throw std::runtime_error("[TinyAD-B200] Element valence exceeds maximum static valence passed to add_elements<..>(..). "
                                         "Please pass a large-enough static valence as template argument.");
            detail::check(s);
        }
        std::map<int, std::vector<int64_t>> groups;
        for (int64_t e = 0; e < n; ++e)
        {
            int valence = 0;
            for (int j = 0; j < max_valence; ++j) valence += table[(size_t)(j * n + e)] >= 0 ? 1 : 0;
            const auto it = std::lower_bound(static_valences_sorted.begin(), static_valences_sorted.end(), valence);
            groups[*it].push_back(handles[(size_t)e]);  // it != end(): valence <= max_valence was checked by the record pass
        }
        (add_dynamic_group<ElementValences>(groups, _eval_element), ...);
    }

    // ScalarFunctionImpl.hh:216-254
    template <class ReadFn>
    std::vector<double> x_from_data(ReadFn&& _read_user_data) const
    {
        std::vector<double> x((size_t)n_vars, NAN);
        for (auto v : variable_handles)
        {
            const auto user_vec = _read_user_data(v);
            for (int i = 0; i < variable_dimension; ++i) x[(size_t)(variable_dimension * idx_from_handle(v) + i)] = user_vec[i];
        }
        for (double xi : x)
            if (!std::isfinite(xi)) throw std::runtime_error("[TinyAD-B200] x_from_data: non-finite entry");
        return x;
    }
    template <class WriteFn>
    void x_to_data(const std::vector<double>& _x, WriteFn&& _write_user_data) const
    {
        check_size(_x);
        for (auto v : variable_handles)
        {
            Vec<double, variable_dimension> vec;
            for (int i = 0; i < variable_dimension; ++i) vec[i] = _x[(size_t)(variable_dimension * idx_from_handle(v) + i)];
            _write_user_data(v, vec);
        }
    }

    // ---- host-vector overloads (H2D / D2H inside) ----
    double eval(const std::vector<double>& _x) const  // ScalarFunctionImpl.hh:256-273
    {
        if (!h) return 0.0;
        check_size(_x);
        double f = 0.0;
        detail::check(tad_eval_host(h, _x.data(), &f));
        return f;
    }
    double operator()(const std::vector<double>& _x) const { return eval(_x); }

    void eval_with_gradient(const std::vector<double>& _x, double& _f, std::vector<double>& _g) const  // :284-299
    {
        _f = 0.0;
        _g.assign((size_t)n_vars, 0.0);
        if (!h) return;
        check_size(_x);
        detail::check(tad_eval_with_gradient_host(h, _x.data(), &_f, _g.data()));
    }
    std::tuple<double, std::vector<double>> eval_with_gradient(const std::vector<double>& _x) const
    {
        double f = 0.0;
        std::vector<double> g;
        eval_with_gradient(_x, f, g);
        return {f, std::move(g)};
    }

    void eval_with_derivatives(const std::vector<double>& _x, double& _f, std::vector<double>& _g, SparseMatrix& _H) const  // :316-336
    {
        eval_second(_x, _f, _g, _H, false, NAN);
    }
    std::tuple<double, std::vector<double>, SparseMatrix> eval_with_derivatives(const std::vector<double>& _x) const
    {
        double f = 0.0;
        std::vector<double> g;
        SparseMatrix H;
        eval_with_derivatives(_x, f, g, H);
        return {f, std::move(g), std::move(H)};
    }
    SparseMatrix eval_hessian(const std::vector<double>& _x) const  // :354-368
    {
        double f = 0.0;
        std::vector<double> g;
        SparseMatrix H;
        eval_with_derivatives(_x, f, g, H);
        return H;
    }
    SparseMatrix eval_hessian_of_quadratic() const { return eval_hessian(std::vector<double>((size_t)n_vars, 0.0)); }  // :370-376

    void eval_with_hessian_proj(const std::vector<double>& _x, double& _f, std::vector<double>& _g, SparseMatrix& _H_proj,
                                const double& _projection_eps = default_hessian_projection_eps) const  // :378-399
    {
        eval_second(_x, _f, _g, _H_proj, true, _projection_eps);
    }
    std::tuple<double, std::vector<double>, SparseMatrix> eval_with_hessian_proj(
        const std::vector<double>& _x, const double& _projection_eps = default_hessian_projection_eps) const
    {
        double f = 0.0;
        std::vector<double> g;
        SparseMatrix H;
        eval_with_hessian_proj(_x, f, g, H, _projection_eps);
        return {f, std::move(g), std::move(H)};
    }

    // ---- device-pointer overloads: x, g, H values stay in HBM (no PCIe traffic in a Newton loop) ----
    double eval_device(const double* x_dev) const
    {
        double f = 0.0;
        if (h) detail::check(tad_eval(h, x_dev, &f));
        return f;
    }
    double eval_with_gradient_device(const double* x_dev, double* g_dev) const
    {
        double f = 0.0;
        if (h) detail::check(tad_eval_with_gradient(h, x_dev, &f, g_dev));
        return f;
    }
    double eval_with_derivatives_device(const double* x_dev, double* g_dev, double* H_values_dev) const
    {
        double f = 0.0;
        if (h) detail::check(tad_eval_with_derivatives(h, x_dev, &f, g_dev, H_values_dev, 0, NAN));
        return f;
    }
    double eval_with_hessian_proj_device(const double* x_dev, double* g_dev, double* H_values_dev,
                                         double eps = default_hessian_projection_eps) const
    {
        double f = 0.0;
        if (h) detail::check(tad_eval_with_derivatives(h, x_dev, &f, g_dev, H_values_dev, 1, eps));
        return f;
    }
    // Fixed pattern of the Hessian (host copy without values).
    SparseMatrix pattern() const
    {
        SparseMatrix P;
        P.rows = P.cols = n_vars;
        if (!h) { P.outer.assign(1, 0); return P; }
        int64_t n_outer = 0, nnz = 0;
        detail::check(tad_function_pattern(h, &n_outer, &nnz));
        P.outer.resize((size_t)n_outer + 1);
        P.inner.resize((size_t)nnz);
        detail::check(tad_function_pattern_copy(h, P.outer.data(), P.inner.data()));
        return P;
    }

    tad_function handle() const { return h; }

    EvalSettings settings;             // ScalarFunction.hh:221
    int64_t n_vars = 0;                // :224
    int64_t n_elements = 0;            // :227
    std::vector<VariableHandleT> variable_handles;  // :230

private:
    template <int element_valence, bool force_dense, typename EvalElementFunction>
    void add_term_impl(const std::vector<int64_t>& handles, bool identity, EvalElementFunction _eval_element)
    {
        using L = TermLauncher<EvalElementFunction, variable_dimension, element_valence, 0, force_dense>;
        L* launcher = new L{std::move(_eval_element)};
        detail::check(tad_function_add_term(h, element_valence, 0, (int64_t)handles.size(), identity ? nullptr : handles.data(), &L::launch,
                                            launcher, &L::destroy));
        n_elements += (int64_t)handles.size();
    }
    template <int element_valence, typename EvalElementFunction>
    void add_dynamic_group(const std::map<int, std::vector<int64_t>>& groups, const EvalElementFunction& _eval_element)
    {
        const auto it = groups.find(element_valence);
        if (it == groups.end()) return;
        add_term_impl<element_valence, true>(it->second, false, _eval_element);
    }
    void require_handle() const
    {
        if (!h) throw std::runtime_error("[TinyAD-B200] function has no variables (default-constructed or moved-from)");
    }
    void check_size(const std::vector<double>& _x) const
    {
        if ((int64_t)_x.size() != n_vars) throw std::runtime_error("[TinyAD-B200] x.size() != n_vars");
    }
    void eval_second(const std::vector<double>& _x, double& _f, std::vector<double>& _g, SparseMatrix& _H, bool project, double eps) const
    {
        _f = 0.0;
        _g.assign((size_t)n_vars, 0.0);
        _H = SparseMatrix();
        _H.rows = _H.cols = n_vars;
        _H.outer.assign((size_t)n_vars + 1, 0);
        if (!h) return;
        check_size(_x);
        _H = pattern();
        _H.values.assign(_H.inner.size(), 0.0);
        detail::check(tad_eval_with_derivatives_host(h, _x.data(), &_f, _g.data(), _H.values.data(), project ? 1 : 0, eps));
    }

    tad_function h = nullptr;
};

// ScalarFunction.hh:242-248 / ScalarFunctionImpl.hh:418-442
template <int variable_dimension, typename PassiveT = double, typename VariableRangeT>
auto scalar_function(const VariableRangeT& _variable_range, const EvalSettings& _settings = EvalSettings())
{
    using VariableHandle = typename std::decay_t<decltype(*_variable_range.begin())>;
    std::vector<VariableHandle> variable_handles;
    for (auto vh : _variable_range) variable_handles.push_back(vh);
    return ScalarFunction<variable_dimension, PassiveT, VariableHandle>(std::move(variable_handles), _settings);
}

}  // namespace TinyAD
