// tinyad_b200 -- newton_decrement (Utils/NewtonDecrement.hh:20-26): -0.5 d.g, the difference between f(x) and the minimum
// of the quadratic model; affinely invariant stopping criterion.
#pragma once

#include <stdexcept>
#include <vector>

#include <tinyad_b200.h>

namespace TinyAD
{

inline double newton_decrement(const std::vector<double>& _d, const std::vector<double>& _g)
{
    if (_d.size() != _g.size()) throw std::runtime_error("[TinyAD-B200] newton_decrement: size mismatch");
    double s = 0.0;
    for (size_t i = 0; i < _d.size(); ++i) s += _d[i] * _g[i];  // Eigen's dot: left-to-right
    return -0.5 * s;
}

// device-resident vectors of a function
template <class FunctionT>
double newton_decrement_device(const FunctionT& _func, const double* _d_dev, const double* _g_dev)
{
    double out = 0.0;
    if (tad_newton_decrement(_func.handle(), _d_dev, _g_dev, &out) != TAD_OK)
        throw std::runtime_error(std::string("[TinyAD-B200] ") + tad_last_error());
    return out;
}

}  // namespace TinyAD
