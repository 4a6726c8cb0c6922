// fast_fp4.cu -- streaming instantiations of the FP4 functors (fp4_ops.cuh) and the device-side exhaustive check of the
// NVFP4 division shortcut.
#include "engine.h"
#include "fp4_ops.cuh"

namespace ctb {

#define SIG_FAIL(sig)                                                                              \
    do {                                                                                           \
        set_error("no fast fp4 kernel for op=%d dtype=%d sel=%d zp=%d", sig.op, sig.p_dt, sig.sel, sig.zp); \
        return CT_E_UNSUPPORTED;                                                                   \
    } while (0)

// sig.sel: F_FP4_QUANTPACK: 0 = MX (arithmetic in T), 1 = NV with scale in T, 2 = NV with float32 scale, 3 = NV with the group
//          observer fused in (scale is an OUTPUT, float8_e4m3fn)
//          F_FP4_UNPACKDEQ: Fp4ScaleKind of the scale tensor (FS_SAME / FS_F8 / FS_E8M0)
// sig.zp : Fp4ZpKind
template <class P>
static int fp4_quantpack_p(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.sel == 0) {
        if (s.zp == FZ_NONE) return launch_stream<Fp4MxQuantPackOp<P, FZ_NONE>>(lp, device, st);
        if (s.zp == FZ_U8) return launch_stream<Fp4MxQuantPackOp<P, FZ_U8>>(lp, device, st);
        if (s.zp == FZ_I8) return launch_stream<Fp4MxQuantPackOp<P, FZ_I8>>(lp, device, st);
    } else if (s.sel == 1) {
        if (s.zp == FZ_NONE) return launch_stream<Fp4NvQuantPackOp<P, FS_SAME, FZ_NONE>>(lp, device, st);
        if (s.zp == FZ_F8) return launch_stream<Fp4NvQuantPackOp<P, FS_SAME, FZ_F8>>(lp, device, st);
    } else if (s.sel == 2) {
        if (s.zp == FZ_NONE) return launch_stream<Fp4NvQuantPackOp<P, FS_F32, FZ_NONE>>(lp, device, st);
        if (s.zp == FZ_F8) return launch_stream<Fp4NvQuantPackOp<P, FS_F32, FZ_F8>>(lp, device, st);
    }
    SIG_FAIL(s);
}
template <class P>
static int fp4_unpackdeq_p(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.sel == FS_SAME) return launch_stream<Fp4UnpackDequantOp<P, FS_SAME>>(lp, device, st);
    if (s.sel == FS_F8) return launch_stream<Fp4UnpackDequantOp<P, FS_F8>>(lp, device, st);
    if (s.sel == FS_E8M0) return launch_stream<Fp4UnpackDequantOp<P, FS_E8M0>>(lp, device, st);
    SIG_FAIL(s);
}

int launch_fast_fp4(const FastSig& s, const LaunchPlan& lp, int device, cudaStream_t st) {
    if (s.op == F_FP4_QUANTPACK && s.sel == 3) {   // observe + quantize + pack (scale is an output)
        if (s.p_dt == CT_BF16) return launch_stream<Fp4NvObserveQuantPackOp<BF16>>(lp, device, st);
        if (s.p_dt == CT_F16) return launch_stream<Fp4NvObserveQuantPackOp<F16>>(lp, device, st);
        SIG_FAIL(s);
    }
    if (s.op == F_FP4_QUANTPACK) {
        if (s.p_dt == CT_BF16) return fp4_quantpack_p<BF16>(s, lp, device, st);
        if (s.p_dt == CT_F16) return fp4_quantpack_p<F16>(s, lp, device, st);
    } else {
        if (s.p_dt == CT_BF16) return fp4_unpackdeq_p<BF16>(s, lp, device, st);
        if (s.p_dt == CT_F16) return fp4_unpackdeq_p<F16>(s, lp, device, st);
    }
    SIG_FAIL(s);
}

// ---- selftest: the E2M1 code of quotient(x, s) (reciprocal + one residual step) must equal the code of div.rn(x, s) for
// ---- every 16-bit x and every float32 scale significand, at several scale exponents inside the fast range ------------------
// mode 0: E2M1 code of the quantization quotient (weights / effective scale).  mode 1: the float32 VALUE of scale / global_scale
// for every 16-bit scale with |scale| in [2^-40, 2^14] (what Fp4UnpackDequantOp uses recip_div for).
template <class P>
__global__ void __launch_bounds__(256) fp4_division_selftest_kernel(unsigned long long* mismatches, int exp_field, int mode) {
    using Op = Fp4NvQuantPackOp<P, FS_F32, FZ_NONE>;
    unsigned long long local = 0;
    for (uint32_t m = blockIdx.x; m < (1u << 23); m += gridDim.x) {
        const float s = __uint_as_float(((uint32_t)exp_field << 23) | m);
        const float rc = __frcp_rn(s);
        for (uint32_t xp = threadIdx.x; xp < 65536u; xp += blockDim.x) {
            const float x = P::lo(xp);
            if (mode == 1) {
                const float a = fabsf(x);
                if (!(a >= 9.094947017729282e-13f && a <= 16384.0f)) continue;
                if (__float_as_uint(recip_div(x, s, rc)) != __float_as_uint(__fdiv_rn(x, s))) ++local;
                continue;
            }
            const float fast = Op::quotient(P::lo(Op::clamp_x(xp)), s, rc);   // exactly what chunk_fast does; nothing is added without a zero point
            const float ref = __fadd_rn(__fdiv_rn(x, s), 0.0f);
            if (x != x) continue;   // NaN weights are outside the bit-exact contract
            if (f32x2_to_e2m1x2(fast, 0.f) != f32x2_to_e2m1x2(ref, 0.f)) ++local;
        }
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(mismatches, local);
}

}  // namespace ctb

extern "C" int ct_selftest_fp4_division(int dtype, int scale_exponent, int mode, uint64_t* mismatches, int device) {
    using namespace ctb;
    if (!mismatches) { set_error("null output"); return CT_E_ARG; }
    if (mode == 0 && (scale_exponent < -100 || scale_exponent > 9)) { set_error("scale exponent outside the fast range [-100, 9] (|scale| in [2^-100, 2^10))"); return CT_E_ARG; }
    if (mode == 1 && (scale_exponent < -60 || scale_exponent > 59)) { set_error("global-scale exponent outside the fast range [-60, 59]"); return CT_E_ARG; }
    if (mode != 0 && mode != 1) { set_error("mode must be 0 or 1"); return CT_E_ARG; }
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard guard(device);
    unsigned long long* d = nullptr;
    CT_CUDA_TRY(cudaMalloc(&d, sizeof(unsigned long long)));
    CT_CUDA_TRY(cudaMemset(d, 0, sizeof(unsigned long long)));
    const int ef = scale_exponent + 127;
    if (dtype == CT_BF16) fp4_division_selftest_kernel<BF16><<<148 * 16, 256>>>(d, ef, mode);
    else if (dtype == CT_F16) fp4_division_selftest_kernel<F16><<<148 * 16, 256>>>(d, ef, mode);
    else { cudaFree(d); set_error("selftest supports bf16 / f16"); return CT_E_DTYPE; }
    count_launch();
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "fp4 selftest");
    *mismatches = h;
    return CT_OK;
}
