// tinyad_b200 -- PMP handle support (reference: include/TinyAD/Support/PMP.hh).  The adapter is written against the handle
// interface in Support/Common.hh, so this header only has to exist for source compatibility: include it before ScalarFunction.hh
// / VectorFunction.hh as in the reference.
#pragma once
#include <TinyAD/Support/Common.hh>
#if defined(TINYAD_ScalarFunction_DEFINED) || defined(TINYAD_VectorFunction_DEFINED)
#error Please include this file before ScalarFunction.hh / VectorFunction.hh
#endif
