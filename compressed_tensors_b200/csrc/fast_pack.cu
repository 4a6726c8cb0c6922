// fast_pack.cu -- streaming instantiations: fused quantize+pack, fused unpack+dequantize,
// standalone int8 <-> int32 bit packing.
#include "engine.h"
#include "ops.cuh"

namespace ctb {

#define SIG_FAIL(sig)                                                                              \
    do {                                                                                           \
        set_error("no fast kernel for op=%d dtype=%d sel=%d zp=%d group=%d", sig.op, sig.p_dt, sig.sel, sig.zp, sig.group); \
        return CT_E_UNSUPPORTED;                                                                   \
    } while (0)

// unit sizes: 4-bit codes -> 4 chunks (32 elements -> 16 B out); 8-bit codes -> 2 chunks
template <class P, int BITS, int ZP>
static int quantpack_g(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    constexpr int GMAX = (BITS == 4) ? 4 : 2;
    if constexpr (P::DT != CT_F32) {
        if (s.group == GMAX) return launch_stream<QuantPackOp<P, BITS, ZP, GMAX>>(lp, device, st);
    }
    if (s.group == 1) return launch_stream<QuantPackOp<P, BITS, ZP, 1>>(lp, device, st);
    SIG_FAIL(s);
}
template <class P>
static int quantpack_p(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.sel == 4 && s.zp == 0) return quantpack_g<P, 4, 0>(s, lp, device, st);
    if (s.sel == 4 && s.zp == 1) return quantpack_g<P, 4, 1>(s, lp, device, st);
    if (s.sel == 8 && s.zp == 0) return quantpack_g<P, 8, 0>(s, lp, device, st);
    if (s.sel == 8 && s.zp == 1) return quantpack_g<P, 8, 1>(s, lp, device, st);
    SIG_FAIL(s);
}
int fast_group_quantpack(int p_dt, int bits) { return p_dt == CT_F32 ? 1 : (bits == 4 ? 4 : 2); }

int launch_fast_quantpack(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    switch (s.p_dt) {
    case CT_BF16: return quantpack_p<BF16>(s, lp, device, st);
    case CT_F16: return quantpack_p<F16>(s, lp, device, st);
    case CT_F32: return quantpack_p<F32>(s, lp, device, st);
    }
    SIG_FAIL(s);
}

template <class P>
static int unpackdeq_p(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.sel == 4 && s.zp == 0) return launch_stream<UnpackDequantOp<P, 4, 0>>(lp, device, st);
    if (s.sel == 4 && s.zp == 1) return launch_stream<UnpackDequantOp<P, 4, 1>>(lp, device, st);
    if (s.sel == 8 && s.zp == 0) return launch_stream<UnpackDequantOp<P, 8, 0>>(lp, device, st);
    if (s.sel == 8 && s.zp == 1) return launch_stream<UnpackDequantOp<P, 8, 1>>(lp, device, st);
    SIG_FAIL(s);
}
int launch_fast_unpackdeq(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    switch (s.p_dt) {
    case CT_BF16: return unpackdeq_p<BF16>(s, lp, device, st);
    case CT_F16: return unpackdeq_p<F16>(s, lp, device, st);
    case CT_F32: return unpackdeq_p<F32>(s, lp, device, st);
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
