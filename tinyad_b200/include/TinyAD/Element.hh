// tinyad_b200 -- the object handed to a per-element functor on the device.
//
// Mirrors the reference's Element (include/TinyAD/Detail/Element.hh:16-295): variables(vh)
// reads x[d*idx(vh)+i] (Element.hh:165) and seeds active scalars with their local index in
// FIRST-ACCESS order; a handle requested again gets its first slot (Element.hh:222-231).
// Differences forced by device execution:
//   * no std::vector / exceptions: errors set bits in a device error word that the runtime
//     turns into std::runtime_error on the host after the launch;
//   * the local->global index map is not rebuilt at every evaluation: a one-time RECORD
//     pass (RecorderElement, like ScalarFunctionImpl.hh:106-129) writes it to a table that
//     the runtime turns into the fixed CSR pattern and scatter maps;
//   * when no element of a term repeats a handle (the normal case, detected by the record
//     pass) the slot of the j-th variables() call is the compile-time constant j, so seeds
//     are constants and their zero blocks fold away; otherwise the Dedup instantiation
//     does the reference's linear search at run time.
#pragma once

#include <cstdint>

#include <TinyAD/Matrix.hh>
#include <TinyAD/Scalar.hh>

namespace TinyAD
{

// bits of the device error word (index = tad_status)
constexpr int TINYAD_ERR_TOO_MANY_VARIABLES = 1 << 4;
constexpr int TINYAD_ERR_INDEX_OUT_OF_RANGE = 1 << 5;
constexpr int TINYAD_ERR_PATTERN_MISMATCH = 1 << 9;

namespace detail
{
TINYAD_HD TINYAD_INLINE void raise(int32_t* err, int bit)
{
#if defined(__CUDA_ARCH__)
    if (err) atomicOr(err, bit);
#else
    if (err) *err |= bit;
#endif
}
}  // namespace detail

template <int d, int N, int M, typename ScalarT, bool active_mode_, bool Dedup>
struct Element
{
    static constexpr int n_element = d * N;
    static constexpr bool active_mode = active_mode_;
    using ScalarType = ScalarT;
    using VariableVectorType = Vec<ScalarT, d>;
    using PassiveVectorType = Vec<double, d>;
    using OutputVectorType = Vec<ScalarT, (M > 0 ? M : 1)>;

    // _rec (optional): this element's column of the recorded element -> handle table (entry of slot j at _rec[j * _rec_stride]).
    // The reference rebuilds idx_local_to_global at every evaluation (Element.hh:208-260); here the table was recorded once at
    // add_elements time and the CSR pattern / scatter maps were built from it, so an evaluation that requests other handles
    // (a functor that branches on x before or between its variables() calls) must be reported, not silently mis-assembled.
    TINYAD_HD TINYAD_INLINE Element(int64_t _handle, const double* _x, int64_t _n_handles, int32_t* _err, const int32_t* _rec = nullptr,
                                    int64_t _rec_stride = 0)
        : handle(_handle), x(_x), n_handles(_n_handles), err(_err), n_used(0), rec(_rec), rec_stride(_rec_stride), mismatch(false)
    {
        // Static slots: ALL recorded handles of the element are loaded here, back to back (N coalesced int32 loads).  variables()
        // then addresses x through the RECORDED handle of its slot and only remembers the handle the functor asked for; range and
        // equality are checked after the functor (check_recorded_count).  A warp issues in order, so with the functor's own
        // `element.variables(conn(e, j))` every call used to wait for its connectivity load, then for the range check on it,
        // then for x -- N times two dependent memory round trips at 8 warps per SM (ncu: 15 % of the tet kernels' stall samples
        // on the range checks, another 35 % of the first kernel's on the comparison with the recorded handle).  Now the loads
        // of x depend on one round trip that starts at construction.  A functor that requests another handle than recorded
        // computes with the recorded one and the evaluation reports TAD_PATTERN_MISMATCH / TAD_INDEX_OUT_OF_RANGE.
        if constexpr (!Dedup)
        {
            if (rec)
                detail::static_for<N>([&](auto jc) TINYAD_LAMBDA_INLINE {
                    constexpr int j = decltype(jc)::value;
                    rec_val[j] = rec[j * rec_stride];
                });
        }
    }
    Element(const Element&) = delete;  // Element.hh:78

    TINYAD_HD TINYAD_INLINE VariableVectorType variables(int64_t vh)
    {
        if constexpr (!Dedup)
        {
            if (rec)
            {
                int slot = n_used++;
                if (slot >= N)
                {
                    detail::raise(err, TINYAD_ERR_TOO_MANY_VARIABLES);  // Element.hh:237-238
                    slot = N - 1;
                }
                req_val[slot] = (vh < 0 || vh > 0x7fffffffll) ? -2 : (int32_t)vh;   // checked after the functor
                const int32_t rh = rec_val[slot];
                return load_variables(rh >= 0 ? (int64_t)rh : 0, slot);
            }
        }
        if (vh < 0 || vh >= n_handles)
        {
            detail::raise(err, TINYAD_ERR_INDEX_OUT_OF_RANGE);
            vh = 0;
        }
        int slot = n_used;
        if constexpr (Dedup)
        {
            slot = -1;
            for (int j = 0; j < N; ++j)
                if (j < n_used && seen[j] == vh && slot < 0) slot = j;
            if (slot < 0)
            {
                slot = n_used;
                if (n_used < N) seen[n_used] = vh;
                ++n_used;
            }
        }
        else
            ++n_used;
        if (slot >= N)
        {
            detail::raise(err, TINYAD_ERR_TOO_MANY_VARIABLES);  // Element.hh:237-238
            slot = N - 1;
        }
        if constexpr (Dedup)
        {
            if (rec) mismatch = mismatch || ((int64_t)rec[slot * rec_stride] != vh);   // run-time slot: compare at once
        }
        return load_variables(vh, slot);
    }

    // the d variables of handle vh as active scalars with local indices d * slot + i (or passive values)
    TINYAD_HD TINYAD_INLINE VariableVectorType load_variables(const int64_t vh, const int slot) const
    {
        const double* xv = x + d * vh;
        VariableVectorType v;
        detail::static_for<d>([&](auto ic) TINYAD_LAMBDA_INLINE {
            constexpr int i = decltype(ic)::value;
            if constexpr (active_mode)
            {
                if constexpr (Dedup) v.a[i] = ScalarT::active_dense(xv[i], d * slot + i);
                else v.a[i] = ScalarT(xv[i], d * slot + i);
            }
            else v.a[i] = xv[i];
        });
        return v;
    }
    TINYAD_HD TINYAD_INLINE ScalarT variable(int64_t vh)
    {
        static_assert(d == 1, "element.variable(vh) needs variable dimension 1 (Element.hh:268)");
        return variables(vh).a[0];
    }
    TINYAD_HD TINYAD_INLINE PassiveVectorType variables_passive(int64_t vh) const  // Element.hh:272-285
    {
        if (vh < 0 || vh >= n_handles)
        {
            detail::raise(err, TINYAD_ERR_INDEX_OUT_OF_RANGE);
            vh = 0;
        }
        PassiveVectorType v;
        detail::static_for<d>([&](auto ic) TINYAD_LAMBDA_INLINE { constexpr int i = decltype(ic)::value; v.a[i] = x[d * vh + i]; });
        return v;
    }
    TINYAD_HD TINYAD_INLINE double variable_passive(int64_t vh) const
    {
        static_assert(d == 1, "element.variable_passive(vh) needs variable dimension 1");
        return variables_passive(vh).a[0];
    }

    int64_t handle;  // element handle (Element.hh:127)
    const double* x;
    int64_t n_handles;
    int32_t* err;
    int n_used;
    const int32_t* rec;
    int64_t rec_stride;
    bool mismatch;  // some variables() call requested another handle than the recorded one
    int32_t rec_val[Dedup ? 1 : N], req_val[Dedup ? 1 : N];  // recorded / requested handle of every slot, compared after the functor
    int64_t seen[Dedup ? N : 1];

    // after the functor ran: did it request the recorded handles, and exactly the recorded number of (distinct) handles?
    TINYAD_HD TINYAD_INLINE void check_recorded_count(int32_t recorded) const
    {
        const int want = recorded < 0 ? -recorded - 1 : recorded;  // < 0 marks "a handle was requested more than once"
        bool bad = mismatch, out_of_range = false;
        if constexpr (!Dedup)
        {
            if (rec)
                detail::static_for<N>([&](auto jc) TINYAD_LAMBDA_INLINE {
                    constexpr int j = decltype(jc)::value;
                    if (j < n_used)
                    {
                        const int32_t r = req_val[j];
                        if (r == -2 || (int64_t)r >= n_handles) out_of_range = true;   // Element.hh:159-170
                        else if (rec_val[j] != r) bad = true;
                    }
                });
        }
        if (out_of_range) detail::raise(err, TINYAD_ERR_INDEX_OUT_OF_RANGE);
        if (bad || (Dedup ? (n_used != want) : (recorded >= 0 && n_used != want))) detail::raise(err, TINYAD_ERR_PATTERN_MISMATCH);
    }
};

// Passive element used once per term to record which handles an element touches.
template <int d, int N, int M>
struct RecorderElement
{
    static constexpr int n_element = d * N;
    static constexpr bool active_mode = false;
    using ScalarType = double;
    using VariableVectorType = Vec<double, d>;
    using PassiveVectorType = Vec<double, d>;
    using OutputVectorType = Vec<double, (M > 0 ? M : 1)>;

    TINYAD_HD TINYAD_INLINE RecorderElement(int64_t _handle, int64_t _n_handles, int32_t* _err)
        : handle(_handle), n_handles(_n_handles), err(_err), n_used(0), n_calls(0) {}
    RecorderElement(const RecorderElement&) = delete;

    TINYAD_HD TINYAD_INLINE VariableVectorType variables(int64_t vh)
    {
        ++n_calls;
        if (vh < 0 || vh >= n_handles)
        {
            detail::raise(err, TINYAD_ERR_INDEX_OUT_OF_RANGE);
            return VariableVectorType();
        }
        bool found = false;
        for (int j = 0; j < N; ++j)
            if (j < n_used && seen[j] == vh) found = true;
        if (!found)
        {
            if (n_used < N) seen[n_used] = vh;
            else detail::raise(err, TINYAD_ERR_TOO_MANY_VARIABLES);
            ++n_used;
        }
        return VariableVectorType();
    }
    TINYAD_HD TINYAD_INLINE double variable(int64_t vh) { return variables(vh).a[0]; }
    TINYAD_HD TINYAD_INLINE PassiveVectorType variables_passive(int64_t) const { return PassiveVectorType(); }
    TINYAD_HD TINYAD_INLINE double variable_passive(int64_t) const { return 0.0; }

    int64_t handle;
    int64_t n_handles;
    int32_t* err;
    int n_used;   // distinct handles
    int n_calls;  // variables() calls
    int64_t seen[N > 0 ? N : 1];
};

}  // namespace TinyAD

// Element.hh:300-331
#define TINYAD_ACTIVE_MODE(element) std::decay_t<decltype(element)>::active_mode
#define TINYAD_SCALAR_TYPE(element) typename std::decay_t<decltype(element)>::ScalarType
#define TINYAD_VARIABLES_TYPE(element) typename std::decay_t<decltype(element)>::VariableVectorType
#define TINYAD_VECTOR_TYPE(element) typename std::decay_t<decltype(element)>::OutputVectorType
