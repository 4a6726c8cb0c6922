// generic.cuh -- parameter block and per-element arithmetic shared by the generic kernels (generic.cu, fp4.cu)
#pragma once
#include "engine.h"
#include "quant_core.cuh"

namespace ctb {

struct GParams {
    int64_t rows, cols, rdiv, cdiv, srs;
    int x_dt, s_dt, zp_dt, cd, q_dt, out_dt, qtype, bits;
    const void* in;
    const void* scale;
    const void* zp;
    const int32_t* gidx;
    void* out;
    float qmin, qmax;
    const float* gs;   // global scale (one float32 on the device) or nullptr
    int se_dt;         // dtype of scale / global_scale; == s_dt (for stored fp8 / E8M0 scales: the float dtype they decode to) without one
};

__device__ __forceinline__ int64_t scale_index(const GParams& p, int64_t r, int64_t c) {
    int64_t rb = (p.rdiv == 1) ? r : (p.rdiv == CT_DIV_INF ? 0 : r / p.rdiv);
    int64_t cb;
    if (p.gidx) cb = p.gidx[c];
    else cb = (p.cdiv == CT_DIV_INF) ? 0 : c / p.cdiv;
    return rb * p.srs + cb;
}

// the scale every op works with: scale / global_scale in se_dt when there is a global scale (forward_helpers.py:535-536)
__device__ __forceinline__ float eff_scale(const GParams& p, int64_t si) {
    const float s = rnd_dt(load_as_f32(p.scale, si, p.s_dt), p.se_dt);
    return p.gs ? rnd_dt(__fdiv_rn(s, *p.gs), p.se_dt) : s;
}

// quantized value (in compute dtype, as fp32) of x[r, c]
__device__ __forceinline__ float quant_at(const GParams& p, int64_t r, int64_t c) {
    const int64_t si = scale_index(p, r, c);
    const float x = load_as_f32(p.in, r * p.cols + c, p.x_dt);
    const float s = eff_scale(p, si);
    float z = 0.f;
    if (p.zp) z = rnd_dt(load_as_f32(p.zp, si, p.zp_dt), p.x_dt);   // zero_point.to(x.dtype)
    return quant_scalar(x, s, p.zp != nullptr, z, p.cd, p.qtype, p.qmin, p.qmax);
}

// dequantized value of code q (already widened to fp32) at [r, c], rounded per op to scale dtype
__device__ __forceinline__ float dequant_at(const GParams& p, float q, int64_t r, int64_t c) {
    const int64_t si = scale_index(p, r, c);
    float v = rnd_dt(q, p.se_dt);
    const float s = eff_scale(p, si);
    if (p.zp) v = rnd_dt(__fsub_rn(v, rnd_dt(load_as_f32(p.zp, si, p.zp_dt), p.se_dt)), p.se_dt);
    return rnd_dt(__fmul_rn(v, s), p.se_dt);
}

static GParams make_params(const ct_quant_desc& d, const void* in, const void* scale, const void* zp,
                           const int32_t* gidx, void* out) {
    GParams p;
    p.rows = d.rows; p.cols = d.cols; p.rdiv = d.rdiv; p.cdiv = d.cdiv; p.srs = d.s_row_stride;
    p.x_dt = d.x_dtype; p.s_dt = d.scale_dtype; p.zp_dt = d.zp_dtype; p.cd = d.compute_dtype;
    p.q_dt = d.q_dtype; p.out_dt = d.out_dtype; p.qtype = d.qtype; p.bits = d.num_bits;
    p.in = in; p.scale = scale; p.zp = zp; p.gidx = gidx; p.out = out;
    if (d.qtype == CT_Q_INT) {
        const float r = (float)(1 << d.num_bits);
        p.qmax = r / 2 - 1; p.qmin = -r / 2;
    } else if (d.qtype == CT_Q_FP4) {
        p.qmax = 6.f; p.qmin = -6.f;
    } else {
        p.qmax = 448.f; p.qmin = -448.f;
    }
    p.gs = reinterpret_cast<const float*>(d.global_scale);
    p.se_dt = d.global_scale ? d.seff_dtype : (is_float_dt(d.scale_dtype) ? d.scale_dtype : d.seff_dtype);
    return p;
}

static unsigned grid_for(int64_t work_items) {
    int64_t b = (work_items + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)b;
}

}  // namespace ctb
