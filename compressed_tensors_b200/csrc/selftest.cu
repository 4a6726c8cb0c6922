// selftest.cu -- device-side exhaustive proof of the reciprocal-based quotient rounding
// (see quant_core.cuh): for every pair of 16-bit patterns (x, s) with s in the fast-path range,
// T(quot(x, s)) must equal T(fp32(x) / fp32(s)) bit for bit (NaN == NaN).
#include "engine.h"
#include "quant_core.cuh"

namespace ctb {

template <class P>
__global__ void __launch_bounds__(256) division_selftest_kernel(unsigned long long* mismatches) {
    unsigned long long local = 0;
    // blockIdx.x enumerates scale patterns (65536), threads sweep x patterns
    for (uint32_t sp = blockIdx.x; sp < 65536u; sp += gridDim.x) {
        const float s = P::lo(sp);
        const ScaleCtx c = make_scale_ctx(s);
        if (c.slow) continue;   // the kernels use IEEE division there
        for (uint32_t xp = threadIdx.x; xp < 65536u; xp += blockDim.x) {
            const float x = P::lo(xp);
            const uint32_t fast = P::pack(quot<P, false>(x, c), 0.f) & 0xffffu;
            const uint32_t ref = P::pack(__fdiv_rn(x, s), 0.f) & 0xffffu;
            const float ff = P::lo(fast), rf = P::lo(ref);
            const bool both_nan = (ff != ff) && (rf != rf);
            if (fast != ref && !both_nan) ++local;
        }
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(mismatches, local);
}

}  // namespace ctb

extern "C" int ct_selftest_division(int dtype, uint64_t* mismatches, int device) {
    using namespace ctb;
    if (!mismatches) { set_error("null output"); return CT_E_ARG; }
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard guard(device);
    unsigned long long* d = nullptr;
    CT_CUDA_TRY(cudaMalloc(&d, sizeof(unsigned long long)));
    CT_CUDA_TRY(cudaMemset(d, 0, sizeof(unsigned long long)));
    if (dtype == CT_BF16) division_selftest_kernel<BF16><<<148 * 8, 256>>>(d);
    else if (dtype == CT_F16) division_selftest_kernel<F16><<<148 * 8, 256>>>(d);
    else { cudaFree(d); set_error("selftest supports bf16 / f16"); return CT_E_DTYPE; }
    count_launch();
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "selftest");
    *mismatches = h;
    return CT_OK;
}
