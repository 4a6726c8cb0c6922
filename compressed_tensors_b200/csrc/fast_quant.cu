// fast_quant.cu -- streaming instantiations: quantize (to int8 / float8_e4m3fn) and dequantize.
#include "engine.h"
#include "ops.cuh"

namespace ctb {

#define SIG_FAIL(sig)                                                                              \
    do {                                                                                           \
        set_error("no fast kernel for op=%d dtype=%d sel=%d zp=%d", sig.op, sig.p_dt, sig.sel, sig.zp); \
        return CT_E_UNSUPPORTED;                                                                   \
    } while (0)

template <template <class, int, int> class OP, class P>
static int by_kind_zp(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    // for quantize/dequantize the NARROW/WIDE integer kinds share one kernel
    const int kind = (s.sel == QF8) ? QF8 : QI_WIDE;
    if (kind == QF8 && s.zp == 0) return launch_stream<OP<P, QF8, 0>>(lp, device, st);
    if (kind == QF8 && s.zp == 1) return launch_stream<OP<P, QF8, 1>>(lp, device, st);
    if (kind == QI_WIDE && s.zp == 0) return launch_stream<OP<P, QI_WIDE, 0>>(lp, device, st);
    if (kind == QI_WIDE && s.zp == 1) return launch_stream<OP<P, QI_WIDE, 1>>(lp, device, st);
    SIG_FAIL(s);
}

int launch_fast_quant(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    switch (s.p_dt) {
    case CT_BF16: return by_kind_zp<QuantizeOp, BF16>(s, lp, device, st);
    case CT_F16: return by_kind_zp<QuantizeOp, F16>(s, lp, device, st);
    case CT_F32: return by_kind_zp<QuantizeOp, F32>(s, lp, device, st);
    }
    SIG_FAIL(s);
}

int launch_fast_dequant(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    switch (s.p_dt) {
    case CT_BF16: return by_kind_zp<DequantizeOp, BF16>(s, lp, device, st);
    case CT_F16: return by_kind_zp<DequantizeOp, F16>(s, lp, device, st);
    case CT_F32: return by_kind_zp<DequantizeOp, F32>(s, lp, device, st);
    }
    SIG_FAIL(s);
}

}  // namespace ctb
