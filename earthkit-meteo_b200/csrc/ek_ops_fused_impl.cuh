// ek_ops_fused_impl.cuh -- fused multi-output suites: read (t,q,p) or (t,td,p) once, write every requested field.
// Included by ek_ops_fused_tqp.cu / ek_ops_fused_ttdp.cu (one translation unit per suite, compiled in parallel).
#pragma once
#include "ek_launch.cuh"

using namespace ek;

template <template <uint32_t> class OpM, template <uint32_t> class OpME, typename T>
static int suite(const char* what, ek_operand a, ek_operand b, ek_operand c, void* const* outs, uint32_t out_mask, int64_t n, void* stream) {
    if (!outs) return set_error(EK_ERR_ARG, "%s: outs is NULL", what);
    if (out_mask == 0 || out_mask >= (1u << S_NSLOTS)) return set_error(EK_ERR_ARG, "%s: out_mask=0x%x selects no valid output", what, out_mask);
    void* o[S_NSLOTS];
    for (int k = 0; k < S_NSLOTS; ++k) {
        o[k] = (out_mask >> k) & 1u ? outs[k] : nullptr;
        if (((out_mask >> k) & 1u) && !outs[k]) return set_error(EK_ERR_ARG, "%s: output %d requested but its buffer is NULL", what, k);
    }
    ek_operand ins[3] = {a, b, c};
    // output sets with a dedicated compile-time instantiation (the bench / config workloads); any other set
    // runs the generic kernel that tests the mask at run time
    switch (out_mask) {
        case 0x1F: return launch<OpM<0x1F>, OpME<0x1F>, T>(what, ins, o, n, Params{}, stream);  // theta, es, rh, td|q, tv
        case 0x05: return launch<OpM<0x05>, OpME<0x05>, T>(what, ins, o, n, Params{}, stream);  // theta, rh
        case 0x2C: return launch<OpM<0x2C>, OpME<0x2C>, T>(what, ins, o, n, Params{}, stream);  // rh, td|q, w
        default: return launch<OpM<0>, OpME<0>, T>(what, ins, o, n, Params{}, stream);
    }
}

