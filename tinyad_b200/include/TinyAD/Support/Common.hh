// tinyad_b200 -- handle adapters (reference: include/TinyAD/Support/Common.hh:13-41 and Support/{OpenMesh,Polymesh,PMP,
// GeometryCentral}.hh).  ScalarFunction / VectorFunction take (vertex, edge, face ...) handles of any mesh library on the HOST:
// scalar_function<d>(mesh.vertices()), add_elements<3>(mesh.faces(), functor).  Internally every handle becomes a contiguous
// integer through idx_from_handle(); on the DEVICE the element functor sees that integer (element.handle) and indexes device
// arrays with it -- mesh-library objects do not exist in device code.
//
// The reference ships one overload per library (OpenMesh::BaseHandle::idx(), pm::primitive_handle::idx.value, pmp::Handle::idx(),
// geometrycentral::Element::getIndex()).  None of these libraries is needed to state what the overloads do, so the adapters here
// are written against the handle INTERFACE: anything integral, anything with .idx() returning an integer, with .idx.value, or
// with .getIndex() is accepted, which covers the four libraries without including their headers.  A further handle type is
// supported by overloading TinyAD::idx_from_handle (or an overload found by ADL) exactly as in the reference.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <type_traits>
#include <utility>

namespace TinyAD
{
namespace detail
{
template <class...> using void_t = void;
template <class H, class = void> struct has_idx_fn : std::false_type {};
template <class H> struct has_idx_fn<H, void_t<decltype(static_cast<std::int64_t>(std::declval<const H&>().idx()))>> : std::true_type {};
template <class H, class = void> struct has_idx_value : std::false_type {};
template <class H> struct has_idx_value<H, void_t<decltype(static_cast<std::int64_t>(std::declval<const H&>().idx.value))>> : std::true_type {};
template <class H, class = void> struct has_get_index : std::false_type {};
template <class H> struct has_get_index<H, void_t<decltype(static_cast<std::int64_t>(std::declval<const H&>().getIndex()))>> : std::true_type {};
}  // namespace detail

// Support/Common.hh:29-32: integer handles (TinyAD::range(n), std::vector<int>)
template <class H, std::enable_if_t<std::is_integral<H>::value, int> = 0>
inline std::int64_t idx_from_handle(H _idx) { return (std::int64_t)_idx; }

// Support/OpenMesh.hh:27-30 (OpenMesh::BaseHandle::idx()), Support/PMP.hh:23-26 (pmp::Handle::idx())
template <class H, std::enable_if_t<!std::is_integral<H>::value && detail::has_idx_fn<H>::value, int> = 0>
inline std::int64_t idx_from_handle(const H& _h) { return (std::int64_t)_h.idx(); }

// Support/Polymesh.hh:27-31 (pm::primitive_handle<tag>::idx.value)
template <class H, std::enable_if_t<!std::is_integral<H>::value && !detail::has_idx_fn<H>::value && detail::has_idx_value<H>::value, int> = 0>
inline std::int64_t idx_from_handle(const H& _h) { return (std::int64_t)_h.idx.value; }

// Support/GeometryCentral.hh:27-31 (geometrycentral::Element<T, M>::getIndex())
template <class H, std::enable_if_t<!std::is_integral<H>::value && !detail::has_idx_fn<H>::value && !detail::has_idx_value<H>::value &&
                                        detail::has_get_index<H>::value, int> = 0>
inline std::int64_t idx_from_handle(const H& _h) { return (std::int64_t)_h.getIndex(); }

// Support/Common.hh:37-41: fallback with the lowest priority in overload resolution
inline std::int64_t idx_from_handle(...)
{
    throw std::runtime_error("[TinyAD-B200] Handle type not supported. Please overload idx_from_handle() for your handle type or include one of "
                             "the provided header files, e.g. TinyAD/Support/OpenMesh.hh.");
}

}  // namespace TinyAD
