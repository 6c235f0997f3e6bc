// tinyad_b200 -- one Hessian part of the Double<12> tet kernel (compile with -DTADX_PART=p, p < TADX_TET_PARTS).
// The parts are independent kernels, so the NP instantiations build in parallel.
#include "energies.cuh"

#ifndef TADX_PART
#error "define TADX_PART"
#endif

namespace TinyAD { namespace detail {
template int launch_second_order_part<tadx::SymDirichlet3D, 3, 4, TADX_TET_PARTS, TADX_PART, false>(const tadx::SymDirichlet3D&, const tad_launch_args&);
} }
