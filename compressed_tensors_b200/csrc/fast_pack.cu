// fast_pack.cu -- streaming instantiations: fused quantize+pack, fused unpack+dequantize,
// standalone int8 <-> int32 bit packing.
#include "engine.h"
#include "ops.cuh"

namespace ctb {

#define SIG_FAIL(sig)                                                                              \
    do {                                                                                           \
        set_error("no fast kernel for op=%d dtype=%d sel=%d zp=%d", sig.op, sig.p_dt, sig.sel, sig.zp); \
        return CT_E_UNSUPPORTED;                                                                   \
    } while (0)

template <template <class, int, int> class OP, class P>
static int by_bits_zp(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.sel == 4 && s.zp == 0) return launch_stream<OP<P, 4, 0>>(lp, device, st);
    if (s.sel == 4 && s.zp == 1) return launch_stream<OP<P, 4, 1>>(lp, device, st);
    if (s.sel == 8 && s.zp == 0) return launch_stream<OP<P, 8, 0>>(lp, device, st);
    if (s.sel == 8 && s.zp == 1) return launch_stream<OP<P, 8, 1>>(lp, device, st);
    SIG_FAIL(s);
}

int launch_fast_quantpack(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    switch (s.p_dt) {
    case CT_BF16: return by_bits_zp<QuantPackOp, BF16>(s, lp, device, st);
    case CT_F16: return by_bits_zp<QuantPackOp, F16>(s, lp, device, st);
    case CT_F32: return by_bits_zp<QuantPackOp, F32>(s, lp, device, st);
    }
    SIG_FAIL(s);
}

int launch_fast_unpackdeq(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    switch (s.p_dt) {
    case CT_BF16: return by_bits_zp<UnpackDequantOp, BF16>(s, lp, device, st);
    case CT_F16: return by_bits_zp<UnpackDequantOp, F16>(s, lp, device, st);
    case CT_F32: return by_bits_zp<UnpackDequantOp, F32>(s, lp, device, st);
    }
    SIG_FAIL(s);
}

int launch_fast_bits(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.op == F_PACK && s.sel == 4) return launch_stream<PackOp<4>>(lp, device, st);
    if (s.op == F_PACK && s.sel == 8) return launch_stream<PackOp<8>>(lp, device, st);
    if (s.op == F_UNPACK && s.sel == 4) return launch_stream<UnpackOp<4>>(lp, device, st);
    if (s.op == F_UNPACK && s.sel == 8) return launch_stream<UnpackOp<8>>(lp, device, st);
    SIG_FAIL(s);
}

}  // namespace ctb
