// tinyad_b200 -- armijo_condition / line_search (Utils/LineSearch.hh:14-65).
#pragma once

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <tinyad_b200.h>

namespace TinyAD
{

inline bool armijo_condition(const double _f_curr, const double _f_new, const double _s, const std::vector<double>& _d,
                             const std::vector<double>& _g, const double _armijo_const)  // :14-24
{
    double dg = 0.0;
    for (size_t i = 0; i < _d.size(); ++i) dg += _d[i] * _g[i];
    return _f_new <= _f_curr + _armijo_const * _s * dg;
}

namespace detail
{
// Step schedule of the reference (LineSearch.hh:44-60): s_max, s_max * shrink, ... and, when s_max > 1, the value 1.0 is visited
// exactly once at the point where the geometric sequence would jump over it.
inline double next_line_search_step(const double s, const double shrink, const bool visit_one)
{
    return (visit_one && s > 1.0 && s * shrink < 1.0) ? 1.0 : s * shrink;
}
}  // namespace detail

// Host vectors; _eval is any callable double(const std::vector<double>&), e.g. the function object itself (LineSearch.hh:26-65).
// The directional derivative d.g is formed once (the reference re-evaluates the same dot product in every Armijo test).
template <typename EvalFunctionT>
std::vector<double> line_search(const std::vector<double>& _x0, const std::vector<double>& _d, const double _f, const std::vector<double>& _g,
                                const EvalFunctionT& _eval, const double _s_max = 1.0, const double _shrink = 0.8, const int _max_iters = 64,
                                const double _armijo_const = 1e-4)
{
    const size_t n = _x0.size();
    if (n != _g.size() || n != _d.size()) throw std::runtime_error("[TinyAD-B200] line_search: size mismatch");
    if (!(_s_max > 0.0)) throw std::runtime_error("[TinyAD-B200] Max step size not positive.");
    double slope = 0.0;
    for (size_t q = 0; q < n; ++q) slope += _d[q] * _g[q];
    std::vector<double> trial(n);
    double step = _s_max;
    for (int it = 0; it < _max_iters; ++it, step = detail::next_line_search_step(step, _shrink, _s_max > 1.0))
    {
        for (size_t q = 0; q < n; ++q) trial[q] = _x0[q] + step * _d[q];
        const double f_trial = _eval(trial);
        if (f_trial != f_trial) throw std::runtime_error("[TinyAD-B200] line_search: objective is NaN");  // TINYAD_ASSERT_EQ(f_new, f_new)
        if (f_trial <= _f + _armijo_const * step * slope) return trial;                                    // Armijo: sufficient decrease
    }
    std::printf("[TinyAD-B200] WARNING: Line search couldn't find improvement.\n");
    return _x0;
}

// Device-resident vectors: every trial is one value-only evaluation on the GPU (tad_line_search).  Returns the accepted
// step (0 = no improvement found, x_new = x0 like the reference).
template <class FunctionT>
double line_search_device(const FunctionT& _func, const double* _x0_dev, const double* _d_dev, const double _f, const double* _g_dev,
                          double* _x_new_dev, double* _f_new = nullptr, const double _s_max = 1.0, const double _shrink = 0.8,
                          const int _max_iters = 64, const double _armijo_const = 1e-4)
{
    double step = 0.0, f_new = 0.0;
    int n_evals = 0;
    if (tad_line_search(_func.handle(), _x0_dev, _d_dev, _f, _g_dev, _s_max, _shrink, _max_iters, _armijo_const, _x_new_dev, &f_new, &step,
                        &n_evals) != TAD_OK)
        throw std::runtime_error(std::string("[TinyAD-B200] ") + tad_last_error());
    if (_f_new) *_f_new = f_new;
    return step;
}

}  // namespace TinyAD
