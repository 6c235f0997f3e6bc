// tinyad_b200 -- host facade of the reference's VectorFunction
// (include/TinyAD/VectorFunction.hh:19-204, Detail/VectorFunctionImpl.hh in /root/reference):
//   auto func = TinyAD::vector_function<d>(TinyAD::range(n));
//   func.add_elements<N, M>(TinyAD::range(m), functor);   // functor returns TINYAD_VECTOR_TYPE(element)
//   auto [f, g, r, J] = func.eval_sum_of_squares_with_derivatives(x);
// Residual rows are element-major per term (row = start_t + M*e + j, VectorObjectiveTerm.hh:173,215-238);
// J is returned in compressed COLUMN storage with ascending row indices, like the reference's
// Eigen::SparseMatrix after setFromTriplets.  The per-residual Hessian tensor
// (eval_with_derivatives, VectorFunctionImpl.hh:206-252) is out of scope.
#pragma once

#include <algorithm>

#include <TinyAD/ScalarFunction.hh>

namespace TinyAD
{

template <int variable_dimension, typename PassiveT = double, typename VariableHandleT = int64_t>
struct VectorFunction
{
    static_assert(std::is_same<PassiveT, double>::value, "the B200 path computes in FP64 only");

    VectorFunction() = default;
    VectorFunction(std::vector<VariableHandleT> _variable_handles, const EvalSettings& _settings = EvalSettings())
        : settings(_settings), n_vars(variable_dimension * (int64_t)_variable_handles.size()),
          variable_handles(std::move(_variable_handles))
    {
        std::vector<bool> used(variable_handles.size(), false);
        for (auto v : variable_handles)
        {
            const int64_t iv = idx_from_handle(v);
            if (iv < 0 || iv >= (int64_t)variable_handles.size() || used[(size_t)iv])
                throw std::runtime_error("[TinyAD-B200] variable indices are not compact");
            used[(size_t)iv] = true;
        }
        if (variable_handles.empty()) throw std::runtime_error("[TinyAD-B200] no variables");
        detail::check(tad_function_create(variable_dimension, (int64_t)variable_handles.size(), 1, settings.device, &h));
    }
    VectorFunction(const VectorFunction&) = delete;
    VectorFunction& operator=(const VectorFunction&) = delete;
    VectorFunction(VectorFunction&& o) noexcept { *this = std::move(o); }
    VectorFunction& operator=(VectorFunction&& o) noexcept
    {
        if (this != &o)
        {
            if (h) tad_function_destroy(h);
            h = o.h; o.h = nullptr;
            settings = o.settings;
            n_vars = o.n_vars; n_elements = o.n_elements; n_outputs = o.n_outputs;
            variable_handles = std::move(o.variable_handles);
            term_elements_ = std::move(o.term_elements_);
            o.n_vars = o.n_elements = o.n_outputs = 0;
        }
        return *this;
    }
    ~VectorFunction() { if (h) tad_function_destroy(h); }

    // VectorFunctionImpl.hh:64-101
    template <int element_valence, int outputs_per_element, typename ElementHandleRangeT, typename EvalElementFunction>
    void add_elements(const ElementHandleRangeT& _element_range, EvalElementFunction _eval_element)
    {
        static_assert(outputs_per_element >= 1, "a vector element needs at least one output");
        if (!h) throw std::runtime_error("[TinyAD-B200] function has no variables");
        std::vector<int64_t> handles;
        const bool identity = detail::is_identity_range(_element_range, handles);
        using L = TermLauncher<EvalElementFunction, variable_dimension, element_valence, outputs_per_element>;
        L* launcher = new L{std::move(_eval_element)};
        detail::check(tad_function_add_term(h, element_valence, outputs_per_element, (int64_t)handles.size(),
                                            identity ? nullptr : handles.data(), &L::launch, launcher, &L::destroy));
        n_elements += (int64_t)handles.size();
        n_outputs += outputs_per_element * (int64_t)handles.size();
        term_elements_.push_back((int64_t)handles.size());
    }

    std::vector<double> eval(const std::vector<double>& _x) const  // :143-159
    {
        std::vector<double> r((size_t)n_outputs, 0.0);
        if (!h) return r;
        Scratch s(*this, _x, false);
        detail::check(tad_veval(h, s.x, s.r));
        s.download(nullptr, &r, nullptr);
        return r;
    }
    std::vector<double> operator()(const std::vector<double>& _x) const { return eval(_x); }

    void eval_with_jacobian(const std::vector<double>& _x, std::vector<double>& _r, SparseMatrix& _J) const  // :170-188
    {
        _r.assign((size_t)n_outputs, 0.0);
        _J = pattern();
        if (!h) return;
        Scratch s(*this, _x, true);
        detail::check(tad_veval_with_jacobian(h, s.x, s.r, s.J));
        s.download(nullptr, &_r, &_J.values);
    }
    std::tuple<std::vector<double>, SparseMatrix> eval_with_jacobian(const std::vector<double>& _x) const
    {
        std::vector<double> r;
        SparseMatrix J;
        eval_with_jacobian(_x, r, J);
        return {std::move(r), std::move(J)};
    }

    // VectorFunctionImpl.hh:203-236 (eval_with_derivatives): residuals, Jacobian and one n_vars x n_vars Hessian per residual.
    // The device returns one dense k x k block per residual (tad_veval_with_derivatives); the sparse matrices of the reference's
    // signature are built from them on the host with setFromTriplets semantics (k^2 entries per residual, zeros kept, entries of a
    // handle requested twice summed).  eval_hessians_device keeps the blocks in HBM.
    void eval_with_derivatives(const std::vector<double>& _x, std::vector<double>& _r, SparseMatrix& _J, std::vector<SparseMatrix>& _H) const
    {
        _r.assign((size_t)n_outputs, 0.0);
        _J = pattern();
        _H.assign((size_t)n_outputs, SparseMatrix());
        for (auto& Hm : _H) { Hm.rows = Hm.cols = n_vars; Hm.outer.assign((size_t)n_vars + 1, 0); }
        if (!h) return;
        Scratch s(*this, _x, true);
        int64_t total = 0;
        detail::check(tad_function_residual_hessian_layout(h, -1, nullptr, nullptr, nullptr, &total));
        double* Hb = nullptr;
        Scratch::cuda_check(cudaMalloc(&Hb, sizeof(double) * (size_t)(total + 1)));
        const int status = tad_veval_with_derivatives(h, s.x, s.r, s.J, Hb);
        std::vector<double> blocks((size_t)total);
        if (status == TAD_OK && total) cudaMemcpy(blocks.data(), Hb, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost);
        cudaFree(Hb);
        detail::check(status);
        s.download(nullptr, &_r, &_J.values);
        int64_t row0 = 0;
        for (int t = 0;; ++t)
        {
            int64_t off = 0, n_res = 0;
            int k = 0;
            if (tad_function_residual_hessian_layout(h, t, &off, &k, &n_res, nullptr) != TAD_OK) break;   // past the last term
            const int N = k / variable_dimension;
            const int64_t n_el = term_elements_[(size_t)t];
            const int M = n_el ? (int)(n_res / n_el) : 0;
            std::vector<int32_t> table((size_t)N * (size_t)n_el);
            detail::check(tad_function_term_table(h, t, table.data()));
            for (int64_t e = 0; e < n_el; ++e)
                for (int m = 0; m < M; ++m)
                {
                    const double* B = blocks.data() + off + ((int64_t)M * e + m) * k * k;
                    // dense accumulation over the element's global variables (columns sorted ascending, duplicates summed)
                    std::vector<std::pair<int64_t, int>> vars;   // (global variable, local index)
                    for (int sl = 0; sl < N; ++sl)
                    {
                        const int32_t vh = table[(size_t)sl * (size_t)n_el + (size_t)e];
                        if (vh < 0) continue;
                        for (int c = 0; c < variable_dimension; ++c) vars.push_back({(int64_t)variable_dimension * vh + c, variable_dimension * sl + c});
                    }
                    std::sort(vars.begin(), vars.end());
                    SparseMatrix& Hm = _H[(size_t)(row0 + (int64_t)M * e + m)];
                    for (size_t cj = 0; cj < vars.size(); ++cj)
                    {
                        for (size_t ri = 0; ri < vars.size(); ++ri)
                        {
                            Hm.inner.push_back((int32_t)vars[ri].first);
                            Hm.values.push_back(B[vars[ri].second * k + vars[cj].second]);
                        }
                        Hm.outer[(size_t)vars[cj].first + 1] = (int32_t)vars.size();
                    }
                    for (int64_t c = 0; c < n_vars; ++c) Hm.outer[(size_t)c + 1] += Hm.outer[(size_t)c];
                }
            row0 += n_res;
        }
    }
    std::tuple<std::vector<double>, SparseMatrix, std::vector<SparseMatrix>> eval_with_derivatives(const std::vector<double>& _x) const
    {
        std::vector<double> r;
        SparseMatrix J;
        std::vector<SparseMatrix> H;
        eval_with_derivatives(_x, r, J, H);
        return {std::move(r), std::move(J), std::move(H)};
    }
    // device-resident form: r (n_outputs), J values (nnz of the CSC pattern), H blocks (tad_function_residual_hessian_layout)
    void eval_with_derivatives_device(const double* x_dev, double* r_dev, double* J_values_dev, double* H_blocks_dev) const
    {
        if (!h) return;
        detail::check(tad_veval_with_derivatives(h, x_dev, r_dev, J_values_dev, H_blocks_dev));
    }

    double eval_sum_of_squares(const std::vector<double>& _x) const  // :254-266
    {
        if (!h) return 0.0;
        Scratch s(*this, _x, false);
        double f = 0.0;
        detail::check(tad_veval_sum_of_squares(h, s.x, &f));
        return f;
    }

    void eval_sum_of_squares_with_derivatives(const std::vector<double>& _x, double& _f, std::vector<double>& _g,
                                              std::vector<double>& _r, SparseMatrix& _J) const  // :268-283
    {
        _f = 0.0;
        _g.assign((size_t)n_vars, 0.0);
        _r.assign((size_t)n_outputs, 0.0);
        _J = pattern();
        if (!h) return;
        Scratch s(*this, _x, true);
        detail::check(tad_veval_sum_of_squares_with_derivatives(h, s.x, &_f, s.g, s.r, s.J));
        s.download(&_g, &_r, &_J.values);
    }
    std::tuple<double, std::vector<double>, std::vector<double>, SparseMatrix> eval_sum_of_squares_with_derivatives(
        const std::vector<double>& _x) const
    {
        double f = 0.0;
        std::vector<double> g, r;
        SparseMatrix J;
        eval_sum_of_squares_with_derivatives(_x, f, g, r, J);
        return {f, std::move(g), std::move(r), std::move(J)};
    }

    SparseMatrix pattern() const  // n_outputs x n_vars, CSC
    {
        SparseMatrix P;
        P.rows = n_outputs;
        P.cols = n_vars;
        if (!h) { P.outer.assign((size_t)n_vars + 1, 0); return P; }
        int64_t n_outer = 0, nnz = 0;
        detail::check(tad_function_pattern(h, &n_outer, &nnz));
        P.outer.resize((size_t)n_outer + 1);
        P.inner.resize((size_t)nnz);
        P.values.assign((size_t)nnz, 0.0);
        detail::check(tad_function_pattern_copy(h, P.outer.data(), P.inner.data()));
        return P;
    }

    tad_function handle() const { return h; }

    EvalSettings settings;
    int64_t n_vars = 0, n_elements = 0, n_outputs = 0;
    std::vector<VariableHandleT> variable_handles;

private:
    // device scratch for the host-vector overloads
    struct Scratch
    {
        double *x = nullptr, *g = nullptr, *r = nullptr, *J = nullptr;
        int64_t nv, no, nnz = 0;
        int prev_device = -1;   // the scratch lives on the function's device (settings.device), whatever the caller's current device is
        Scratch(const VectorFunction& fn, const std::vector<double>& _x, bool jac) : nv(fn.n_vars), no(fn.n_outputs)
        {
            if ((int64_t)_x.size() != nv) throw std::runtime_error("[TinyAD-B200] x.size() != n_vars");
            cuda_check(cudaGetDevice(&prev_device));
            if (prev_device != fn.settings.device) cuda_check(cudaSetDevice(fn.settings.device));
            if (jac) detail::check(tad_function_pattern(fn.h, nullptr, &nnz));
            cuda_check(cudaMalloc(&x, sizeof(double) * (size_t)(nv + 1)));
            cuda_check(cudaMalloc(&g, sizeof(double) * (size_t)(nv + 1)));
            cuda_check(cudaMalloc(&r, sizeof(double) * (size_t)(no + 1)));
            cuda_check(cudaMalloc(&J, sizeof(double) * (size_t)(nnz + 1)));
            cuda_check(cudaMemcpy(x, _x.data(), sizeof(double) * (size_t)nv, cudaMemcpyHostToDevice));
        }
        ~Scratch()
        {
            cudaFree(x); cudaFree(g); cudaFree(r); cudaFree(J);
            if (prev_device >= 0) cudaSetDevice(prev_device);
        }
        void download(std::vector<double>* _g, std::vector<double>* _r, std::vector<double>* _J)
        {
            if (_g) { _g->resize((size_t)nv); cuda_check(cudaMemcpy(_g->data(), g, sizeof(double) * (size_t)nv, cudaMemcpyDeviceToHost)); }
            if (_r) { _r->resize((size_t)no); cuda_check(cudaMemcpy(_r->data(), r, sizeof(double) * (size_t)no, cudaMemcpyDeviceToHost)); }
            if (_J) { _J->resize((size_t)nnz); cuda_check(cudaMemcpy(_J->data(), J, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost)); }
        }
        static void cuda_check(cudaError_t e)
        {
            if (e != cudaSuccess) throw std::runtime_error(std::string("[TinyAD-B200] CUDA error: ") + cudaGetErrorString(e));
        }
    };

    tad_function h = nullptr;
    std::vector<int64_t> term_elements_;   // number of elements of every term (layout of the per-residual Hessian blocks)
};

// VectorFunction.hh:198-204
template <int variable_dimension, typename PassiveT = double, typename VariableRangeT>
auto vector_function(const VariableRangeT& _variable_range, const EvalSettings& _settings = EvalSettings())
{
    using VariableHandle = typename std::decay_t<decltype(*_variable_range.begin())>;
    std::vector<VariableHandle> variable_handles;
    for (auto vh : _variable_range) variable_handles.push_back(vh);
    return VectorFunction<variable_dimension, PassiveT, VariableHandle>(std::move(variable_handles), _settings);
}

}  // namespace TinyAD
