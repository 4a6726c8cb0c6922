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

CT_HD int64_t scale_index(const GParams& p, int64_t r, int64_t c) {
    int64_t rb = (p.rdiv == 1) ? r : (p.rdiv == CT_DIV_INF ? 0 : r / p.rdiv);
    int64_t cb;
    if (p.gidx) cb = p.gidx[c];
    else cb = (p.cdiv == CT_DIV_INF) ? 0 : c / p.cdiv;
    return rb * p.srs + cb;
}

// the scale every op works with: scale / global_scale in se_dt when there is a global scale (forward_helpers.py:535-536)
CT_HD float eff_scale(const GParams& p, int64_t si) {
    const float s = rnd_dt(load_as_f32(p.scale, si, p.s_dt), p.se_dt);
    return p.gs ? rnd_dt(hd_div(s, *p.gs), p.se_dt) : s;
}

// quantized value (in compute dtype, as fp32) of x[r, c]
CT_HD float quant_at(const GParams& p, int64_t r, int64_t c) {
    const int64_t si = scale_index(p, r, c);
    const float x = load_as_f32(p.in, r * p.cols + c, p.x_dt);
    const float s = eff_scale(p, si);
    float z = 0.f;
    if (p.zp) z = rnd_dt(load_as_f32(p.zp, si, p.zp_dt), p.x_dt);   // zero_point.to(x.dtype)
    return quant_scalar(x, s, p.zp != nullptr, z, p.cd, p.qtype, p.qmin, p.qmax);
}

// dequantized value of code q (already widened to fp32) at [r, c], rounded per op to scale dtype
CT_HD float dequant_at(const GParams& p, float q, int64_t r, int64_t c) {
    const int64_t si = scale_index(p, r, c);
    float v = rnd_dt(q, p.se_dt);
    const float s = eff_scale(p, si);
    if (p.zp) v = rnd_dt(hd_sub(v, rnd_dt(load_as_f32(p.zp, si, p.zp_dt), p.se_dt)), p.se_dt);
    return rnd_dt(hd_mul(v, s), p.se_dt);
}

// ---------------------------------------------------------------------------------------------
// bit packing: one thread owns a group of 32 consecutive elements along the packed dimension,
// which maps to exactly BITS int32 words (helpers.py:62-96).  Words are sums of
// (code + offset) << pos in wrapping int32 arithmetic, like the reference's scatter_add_.
// ---------------------------------------------------------------------------------------------
template <int BITS, class LoadFn>
CT_HD void pack_group(uint32_t (&words)[BITS], int nvalid, LoadFn load) {
#pragma unroll
    for (int k = 0; k < BITS; ++k) words[k] = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
            const int32_t u = load(j) + (1 << (BITS - 1));
            const int bitpos = j * BITS;
            const int w = bitpos >> 5, sh = bitpos & 31;
            words[w] += (uint32_t)u << sh;
            const int ov = sh + BITS - 32;
            if (ov > 0) words[w + 1] += (uint32_t)(u >> (BITS - ov));
        }
    }
}

template <int BITS, class StoreFn>
CT_HD void unpack_group(const uint32_t (&words)[BITS], int nvalid, StoreFn store) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
            const int bitpos = j * BITS;
            const int w = bitpos >> 5, sh = bitpos & 31;
            uint32_t v = words[w] >> sh;
            if (sh + BITS > 32) v |= words[w + 1] << (32 - sh);
            v &= (1u << BITS) - 1u;
            store(j, (int)v - (1 << (BITS - 1)));
        }
    }
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
