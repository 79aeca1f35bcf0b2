// ek_ops_fused.cu -- fused multi-output suites: read (t,q,p) or (t,td,p) once, write every requested field.
#include "ek_launch.cuh"

using namespace ek;

template <class Op, typename T>
static int suite(const char* what, ek_operand a, ek_operand b, ek_operand c, void* const* outs, uint32_t out_mask, int64_t n, void* stream) {
    if (!outs) return set_error(EK_ERR_ARG, "%s: outs is NULL", what);
    if (out_mask == 0 || out_mask >= (1u << S_NSLOTS)) return set_error(EK_ERR_ARG, "%s: out_mask=0x%x selects no valid output", what, out_mask);
    void* o[S_NSLOTS];
    for (int k = 0; k < S_NSLOTS; ++k) {
        o[k] = (out_mask >> k) & 1u ? outs[k] : nullptr;
        if (((out_mask >> k) & 1u) && !outs[k]) return set_error(EK_ERR_ARG, "%s: output %d requested but its buffer is NULL", what, k);
    }
    ek_operand ins[3] = {a, b, c};
    return launch<Op, T>(what, ins, o, n, Params{}, stream);
}

template <typename T>
static int impl_suite_tqp(ek_operand t, ek_operand q, ek_operand p, void* const* outs, uint32_t m, int64_t n, void* stream) {
    return suite<OpSuiteTQP, T>("suite_tqp", t, q, p, outs, m, n, stream);
}
EK_API(suite_tqp, (ek_operand t, ek_operand q, ek_operand p, void* const* outs, uint32_t m, int64_t n, void* stream), (t, q, p, outs, m, n, stream))

template <typename T>
static int impl_suite_ttdp(ek_operand t, ek_operand td, ek_operand p, void* const* outs, uint32_t m, int64_t n, void* stream) {
    return suite<OpSuiteTTdP, T>("suite_ttdp", t, td, p, outs, m, n, stream);
}
EK_API(suite_ttdp, (ek_operand t, ek_operand td, ek_operand p, void* const* outs, uint32_t m, int64_t n, void* stream), (t, td, p, outs, m, n, stream))
