/*
 * oracle/ct_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the compressed-tensors compress/decompress and
 * quantize/dequantize hot path.  It is the CHECKER for the CUDA engine in
 * compressed_tensors_b200/: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Nothing under
 * compressed_tensors_b200/ links, imports or calls this file.
 *
 * Parity status: PINNED for pack/unpack, quantize, dequantize, fake_quantize,
 * bitmask pack/unpack and the 2:4 cutlass metadata transform (checked against
 * golden vectors produced by importing the reference, see
 * tests/golden/make_golden.py, and against the reference tests' known-answer
 * vectors).  "parity unpinned" for sparse24_bitmask / sparse_bitmask
 * compress+decompress: those compressors are absent from the reference
 * snapshot; only their bit order (pack_bitmasks) is pinned.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src/compressed_tensors).  The arithmetic of the reference
 * lives in PyTorch ATen CPU kernels (torch>=2.10, installed 2.11.0): true
 * division, add, clamp, round-half-even, and narrowing casts, each rounding to
 * the tensor dtype.  Those IEEE semantics are restated here with explicit
 * fp32 math + explicit round-to-nearest-even narrowing between every op.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* dtype codes shared with oracle/__init__.py */
enum { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2, DT_I8 = 3, DT_F8E4M3 = 4, DT_I32 = 5, DT_U8 = 6, DT_I64 = 7 };
enum { Q_INT = 0, Q_FLOAT = 1, Q_FP4 = 2 };   /* Q_FLOAT = fp8 e4m3, Q_FP4 = fp4 e2m1 (ct_oracle_fp4.c) */

#define ORC_OK 0
#define ORC_E_DTYPE (-1)
#define ORC_E_BITS (-2)
#define ORC_E_SHAPE (-3)

/* ------------------------------------------------------------------------- */
/* scalar format conversions (IEEE round-to-nearest-even)                     */
/* ------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float bf16_to_f32(uint16_t h) { return u2f((uint32_t)h << 16); }

static inline uint16_t f32_to_bf16(float f) {
    uint32_t u = f2u(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0; /* torch: any NaN -> 0x7fc0 */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}

static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    if (exp == 0) {
        if (man == 0) return u2f(sign);
        /* subnormal: value = man * 2^-24 */
        float v = (float)man * 5.9604644775390625e-08f;
        return sign ? -v : v;
    }
    if (exp == 31) return u2f(sign | 0x7f800000u | (man << 13));
    return u2f(sign | ((exp + 112u) << 23) | (man << 13));
}

static inline uint16_t f32_to_f16(float f) {
    uint32_t u = f2u(f);
    uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
    uint32_t a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);      /* NaN */
    if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);     /* >= 65520 -> inf */
    if (a < 0x38800000u) {                                       /* < 2^-14: subnormal half */
        /* round a * 2^24 to integer, ties to even, via fp32 magic add */
        float v = u2f(a) * 16777216.0f;                          /* exact scaling */
        float r = nearbyintf(v);                                 /* RNE in default mode */
        return (uint16_t)(sign | (uint16_t)r);
    }
    uint32_t lsb = (a >> 13) & 1u;
    a += 0xfffu + lsb;
    return (uint16_t)(sign | ((a - 0x38000000u) >> 13));
}

/* float -> float8_e4m3fn, round-to-nearest-even, no saturation: values that
 * round above 448 become NaN (0x7f), exactly like torch's `.to(float8_e4m3fn)`
 * (c10 Float8_e4m3fn, the third-party arithmetic used by
 * quantization/quant_args.py:483).  1 sign, 4 exp (bias 7), 3 mantissa bits,
 * min subnormal 2^-9, max 448, no inf. */
static inline uint8_t f32_to_f8e4m3(float f) {
    uint32_t u = f2u(f);
    uint8_t sign = (uint8_t)((u >> 24) & 0x80u);
    uint32_t a = u & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint8_t)(sign | 0x7fu);   /* NaN */
    float av = u2f(a);
    if (av >= 480.0f) return (uint8_t)(sign | 0x7fu);      /* rounds past 448 (464 ties to even 448; >=465 in bf16/ >464 -> 480) */
    if (av < 0.015625f) {                                   /* below 2^-6: subnormal grid, step 2^-9 */
        float r = nearbyintf(av * 512.0f);                  /* 0..8 */
        return (uint8_t)(sign | (uint8_t)r);               /* r == 8 -> 0x08 = 2^-6, the first normal: correct */
    }
    /* normal: keep 3 mantissa bits */
    uint32_t lsb = (a >> 20) & 1u;
    a += 0x7ffffu + lsb;
    uint32_t e = (a >> 23);                                 /* biased fp32 exponent after rounding */
    uint32_t m = (a >> 20) & 7u;
    uint32_t code = ((e - 120u) << 3) | m;                  /* 127-7 = 120 */
    if (code > 0x7eu) return (uint8_t)(sign | 0x7fu);       /* 464 < |x| < 480 rounds to "480" -> NaN */
    return (uint8_t)(sign | code);
}

static inline float f8e4m3_to_f32(uint8_t b) {
    uint32_t sign = (uint32_t)(b & 0x80u) << 24;
    uint32_t e = (b >> 3) & 0xfu, m = b & 7u;
    float v;
    if (e == 0xf && m == 7) return u2f(sign | 0x7fc00000u);
    if (e == 0) v = (float)m * 0.001953125f;                /* m * 2^-9 */
    else v = u2f(((e + 120u) << 23) | (m << 20));
    return sign ? -v : v;
}

/* narrow an fp32 value to a float dtype and widen back: one ATen op result */
static inline float rnd(float v, int dt) {
    switch (dt) {
    case DT_BF16: return bf16_to_f32(f32_to_bf16(v));
    case DT_F16: return f16_to_f32(f32_to_f16(v));
    default: return v;
    }
}

static inline float load_as_f32(const void* p, int64_t i, int dt) {
    switch (dt) {
    case DT_F32: return ((const float*)p)[i];
    case DT_F16: return f16_to_f32(((const uint16_t*)p)[i]);
    case DT_BF16: return bf16_to_f32(((const uint16_t*)p)[i]);
    case DT_I8: return (float)((const int8_t*)p)[i];
    case DT_U8: return (float)((const uint8_t*)p)[i];
    case DT_F8E4M3: return f8e4m3_to_f32(((const uint8_t*)p)[i]);
    case DT_I32: return (float)((const int32_t*)p)[i];   /* int32 -> float RNE, as torch */
    case DT_I64: return (float)((const int64_t*)p)[i];
    default: return 0.0f;
    }
}

/* float -> int8 as torch's CPU cast does on x86: NaN/inf -> 0, else
 * truncate-to-int32 then wrap to 8 bits.  Only the in-range branch is
 * normative (values were clamped first); NaN is excluded from bit-exact sets. */
static inline int8_t f32_to_i8(float v) {
    if (!(v == v) || isinf(v)) return 0;
    return (int8_t)(int32_t)v;
}

static inline void store_from_f32(void* p, int64_t i, int dt, float v) {
    switch (dt) {
    case DT_F32: ((float*)p)[i] = v; break;
    case DT_F16: ((uint16_t*)p)[i] = f32_to_f16(v); break;
    case DT_BF16: ((uint16_t*)p)[i] = f32_to_bf16(v); break;
    case DT_I8: ((int8_t*)p)[i] = f32_to_i8(v); break;
    case DT_F8E4M3: ((uint8_t*)p)[i] = f32_to_f8e4m3(v); break;
    case DT_I32: ((int32_t*)p)[i] = (int32_t)v; break;
    default: break;
    }
}

static inline int is_float_dt(int dt) { return dt == DT_F32 || dt == DT_F16 || dt == DT_BF16; }

/* exported scalar helpers so the tests can sweep all bit patterns */
void orc_cast_f32_to_f8e4m3(const float* in, uint8_t* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = f32_to_f8e4m3(in[i]);
}
void orc_cast_f8e4m3_to_f32(const uint8_t* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = f8e4m3_to_f32(in[i]);
}
void orc_cast_f32_to_bf16(const float* in, uint16_t* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = f32_to_bf16(in[i]);
}
void orc_cast_f32_to_f16(const float* in, uint16_t* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = f32_to_f16(in[i]);
}
void orc_cast_f16_to_f32(const uint16_t* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = f16_to_f32(in[i]);
}

/* ------------------------------------------------------------------------- */
/* pack_to_int32 / unpack_from_int32                                          */
/* compressors/pack_quantized/helpers.py:20-101 and :104-180                  */
/* ------------------------------------------------------------------------- */

/* One logical row of `n` int8 values with element stride `stride` -> `nw`
 * int32 words.  The reference builds each word as a SUM of (value+offset)<<pos
 * in int32 (scatter_add_, helpers.py:80-93), not an OR, and the straddling
 * high part uses an arithmetic shift; restated exactly so that even
 * out-of-range inputs agree. Padding elements are raw zeros added after the
 * offset (helpers.py:55,66-67) so they contribute nothing. */
static void pack_row(const int8_t* in, int64_t stride, int64_t n, int bits, int32_t* out, int64_t out_stride, int64_t nw) {
    const int32_t offset = 1 << (bits - 1);
    for (int64_t w = 0; w < nw; ++w) out[w * out_stride] = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t u = (int32_t)in[i * stride] + offset;
        int64_t bitpos = i * (int64_t)bits;
        int64_t w = bitpos >> 5;
        int sh = (int)(bitpos & 31);
        uint32_t acc = (uint32_t)out[w * out_stride] + ((uint32_t)u << sh);
        out[w * out_stride] = (int32_t)acc;
        int ov = sh + bits - 32;
        if (ov > 0 && w + 1 < nw) {
            int32_t hi = u >> (bits - ov); /* arithmetic, helpers.py:88 */
            out[(w + 1) * out_stride] = (int32_t)((uint32_t)out[(w + 1) * out_stride] + (uint32_t)hi);
        }
    }
}

/* in: int8 [rows, cols] row-major.
 * packed_dim=1 -> out int32 [rows, ceil(cols*bits/32)] row-major.
 * packed_dim=0 -> out int32 [ceil(rows*bits/32), cols] row-major, i.e. the
 *                 `.contiguous()` of the transposed view the reference returns
 *                 (helpers.py:98-99, pack_quantized/base.py:109-110). */
int orc_pack_int32(const int8_t* in, int32_t* out, int64_t rows, int64_t cols, int bits, int packed_dim) {
    if (bits < 1 || bits > 8) return ORC_E_BITS;
    if (packed_dim == 1) {
        int64_t nw = (cols * bits + 31) / 32;
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < rows; ++r) pack_row(in + r * cols, 1, cols, bits, out + r * nw, 1, nw);
    } else if (packed_dim == 0) {
        int64_t nw = (rows * bits + 31) / 32;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < cols; ++c) pack_row(in + c, cols, rows, bits, out + c, cols, nw);
    } else {
        return ORC_E_SHAPE;
    }
    return ORC_OK;
}

static void unpack_row(const int32_t* in, int64_t in_stride, int64_t nw, int bits, int8_t* out, int64_t stride, int64_t n) {
    const int32_t offset = 1 << (bits - 1);
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t i = 0; i < n; ++i) {
        int64_t bitpos = i * (int64_t)bits;
        int64_t w = bitpos >> 5;
        int sh = (int)(bitpos & 31);
        uint32_t lo = (w < nw) ? (uint32_t)in[w * in_stride] : 0u;   /* zero padded words, helpers.py:148-151 */
        uint32_t v = lo >> sh;
        int lo_bits = 32 - sh;
        if (lo_bits < bits) {
            uint32_t hi = (w + 1 < nw) ? (uint32_t)in[(w + 1) * in_stride] : 0u;
            v = (v & ((1u << lo_bits) - 1u)) | (hi << lo_bits);
        }
        v &= mask;
        out[i * stride] = (int8_t)((int32_t)v - offset);
    }
}

/* in: packed int32 (layout as produced by orc_pack_int32); out: int8 [rows, cols] */
int orc_unpack_int32(const int32_t* in, int8_t* out, int64_t rows, int64_t cols, int bits, int packed_dim) {
    if (bits < 1 || bits > 8) return ORC_E_BITS;
    if (packed_dim == 1) {
        int64_t nw = (cols * bits + 31) / 32;
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < rows; ++r) unpack_row(in + r * nw, 1, nw, bits, out + r * cols, 1, cols);
    } else if (packed_dim == 0) {
        int64_t nw = (rows * bits + 31) / 32;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < cols; ++c) unpack_row(in + c, cols, nw, bits, out + c, cols, rows);
    } else {
        return ORC_E_SHAPE;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* quantize / dequantize / fake_quantize                                      */
/* quantization/lifecycle/forward_helpers.py:523-546 (_quantize),             */
/* :549-572 (_dequantize), :180-215 (_quantize_dequantize);                   */
/* quantization/quant_args.py:460-496 (round_to_quantized_type_args);         */
/* quantization/utils/helpers.py:198-226 (calculate_range).                   */
/*                                                                            */
/* Scale addressing covers every strategy of forward.py:184-241 with one      */
/* formula: sidx = (r / rdiv) * s_row_stride + (g_idx ? g_idx[c] : c / cdiv)  */
/*   TENSOR : rdiv=cdiv=INT64_MAX                                             */
/*   CHANNEL: rdiv=1, s_row_stride=1, cdiv=INT64_MAX                          */
/*   GROUP  : rdiv=1, s_row_stride=ngroups (0 if the scale has one row),      */
/*            cdiv=group_size; with g_idx the column's group is looked up     */
/*            (forward_helpers.py:149-175: argsort-gather, per-group op,      */
/*            inverse gather == direct lookup)                                */
/*   BLOCK  : rdiv=block_h, s_row_stride=ceil(cols/block_w), cdiv=block_w     */
/*            (forward_helpers.py:62-115; zero padding never reaches the      */
/*            sliced output)                                                  */
/* ------------------------------------------------------------------------- */

static inline void q_range(int qtype, int bits, float* qmin, float* qmax) {
    if (qtype == Q_INT) {
        float r = ldexpf(1.0f, bits);
        *qmax = r / 2 - 1;
        *qmin = -r / 2;
    } else if (qtype == Q_FP4) { /* FP4 e2m1, quant_args.py FP4_E2M1_DATA */
        *qmax = 6.0f;
        *qmin = -6.0f;
    } else { /* FP8 e4m3 */
        *qmax = 448.0f;
        *qmin = -448.0f;
    }
}

float orc_fp4_round(float v);   /* cast_to_fp4, ct_oracle_fp4.c */

/* torch.clamp(t, min, max): NaN propagates */
static inline float clampf(float v, float lo, float hi) {
    if (v != v) return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

/* the value of `quantized_ground` before the final .to(dtype), in compute dtype cd */
static inline float quant_core(float x, float s, int has_zp, float zp_in_xdt, int cd, int qtype, float qmin, float qmax) {
    float t = rnd(x / s, cd);                         /* forward_helpers.py:538 */
    if (has_zp) t = rnd(t + zp_in_xdt, cd);           /* :539-540, in-place add in cd */
    t = clampf(t, qmin, qmax);                        /* quant_args.py:481 */
    if (qtype == Q_INT) t = nearbyintf(t);            /* torch.round = half-to-even, :490 */
    else if (qtype == Q_FP4) t = orc_fp4_round(t);    /* FP4_E2M1_DATA.cast_to_fp4, :485 */
    else t = f8e4m3_to_f32(f32_to_f8e4m3(t));         /* .to(float8_e4m3fn) then back, :483,:495 */
    return t;
}

/* global scale (forward_helpers.py:535-536, 559-560, 196-197): `scale = scale / global_scale` first, in  */
/* se_dt = result_type(scale, global_scale); every later "scale.dtype" is se_dt.  gs == NULL: se_dt = s_dt */
static inline float eff_scale(float sv, const float* gs, int se_dt) { return gs ? rnd(sv / gs[0], se_dt) : sv; }

int orc_quantize_gs(const void* x, int x_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                 const int32_t* g_idx, void* out, int out_dt, int64_t rows, int64_t cols,
                 int64_t rdiv, int64_t cdiv, int64_t s_row_stride, int cd, int qtype, int bits,
                 const float* gs, int se_dt) {
    if (!is_float_dt(x_dt) || !is_float_dt(s_dt) || !is_float_dt(cd)) return ORC_E_DTYPE;
    if (qtype == Q_INT && (bits < 1 || bits > 8)) return ORC_E_BITS;
    if (!gs) se_dt = s_dt;
    float qmin, qmax;
    q_range(qtype, bits, &qmin, &qmax);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        int64_t sbase = (r / rdiv) * s_row_stride;
        for (int64_t c = 0; c < cols; ++c) {
            int64_t si = sbase + (g_idx ? (int64_t)g_idx[c] : c / cdiv);
            float xv = load_as_f32(x, r * cols + c, x_dt);
            float sv = eff_scale(load_as_f32(scale, si, s_dt), gs, se_dt);
            float zv = 0.0f;
            if (zp) zv = rnd(load_as_f32(zp, si, zp_dt), x_dt); /* zero_point.to(x.dtype) */
            float q = quant_core(xv, sv, zp != NULL, zv, cd, qtype, qmin, qmax);
            store_from_f32(out, r * cols + c, out_dt, q);
        }
    }
    return ORC_OK;
}

int orc_quantize(const void* x, int x_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                 const int32_t* g_idx, void* out, int out_dt, int64_t rows, int64_t cols,
                 int64_t rdiv, int64_t cdiv, int64_t s_row_stride, int cd, int qtype, int bits) {
    return orc_quantize_gs(x, x_dt, scale, s_dt, zp, zp_dt, g_idx, out, out_dt, rows, cols, rdiv, cdiv, s_row_stride, cd, qtype, bits, NULL, s_dt);
}

int orc_dequantize_gs(const void* q, int q_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                   const int32_t* g_idx, void* out, int out_dt, int64_t rows, int64_t cols,
                   int64_t rdiv, int64_t cdiv, int64_t s_row_stride, const float* gs, int se_dt) {
    if (!is_float_dt(s_dt) || !is_float_dt(out_dt)) return ORC_E_DTYPE;
    if (!gs) se_dt = s_dt;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        int64_t sbase = (r / rdiv) * s_row_stride;
        for (int64_t c = 0; c < cols; ++c) {
            int64_t si = sbase + (g_idx ? (int64_t)g_idx[c] : c / cdiv);
            float v = rnd(load_as_f32(q, r * cols + c, q_dt), se_dt);      /* x_q.to(scale.dtype), :562 */
            float sv = eff_scale(load_as_f32(scale, si, s_dt), gs, se_dt);
            if (zp) v = rnd(v - rnd(load_as_f32(zp, si, zp_dt), se_dt), se_dt); /* :564-565 */
            v = rnd(v * sv, se_dt);                                         /* :567 */
            store_from_f32(out, r * cols + c, out_dt, v);                   /* :569-570 / forward_helpers.py:171 */
        }
    }
    return ORC_OK;
}

int orc_dequantize(const void* q, int q_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                   const int32_t* g_idx, void* out, int out_dt, int64_t rows, int64_t cols,
                   int64_t rdiv, int64_t cdiv, int64_t s_row_stride) {
    return orc_dequantize_gs(q, q_dt, scale, s_dt, zp, zp_dt, g_idx, out, out_dt, rows, cols, rdiv, cdiv, s_row_stride, NULL, s_dt);
}

int orc_fake_quantize_gs(const void* x, int x_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                      const int32_t* g_idx, void* out, int out_dt, int64_t rows, int64_t cols,
                      int64_t rdiv, int64_t cdiv, int64_t s_row_stride, int cd, int qtype, int bits,
                      const float* gs, int se_dt) {
    if (!is_float_dt(x_dt) || !is_float_dt(s_dt) || !is_float_dt(cd) || !is_float_dt(out_dt)) return ORC_E_DTYPE;
    if (!gs) se_dt = s_dt;
    float qmin, qmax;
    q_range(qtype, bits, &qmin, &qmax);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        int64_t sbase = (r / rdiv) * s_row_stride;
        for (int64_t c = 0; c < cols; ++c) {
            int64_t si = sbase + (g_idx ? (int64_t)g_idx[c] : c / cdiv);
            float xv = load_as_f32(x, r * cols + c, x_dt);
            float sv = eff_scale(load_as_f32(scale, si, s_dt), gs, se_dt);
            float zraw = zp ? load_as_f32(zp, si, zp_dt) : 0.0f;
            float qv = quant_core(xv, sv, zp != NULL, rnd(zraw, x_dt), cd, qtype, qmin, qmax);
            float d = rnd(qv, se_dt);                               /* quantized.to(scale.dtype), :209 */
            if (zp) d = rnd(d - rnd(zraw, se_dt), se_dt);          /* :210-211 */
            d = rnd(d * sv, se_dt);                                 /* :213 */
            store_from_f32(out, r * cols + c, out_dt, d);
        }
    }
    return ORC_OK;
}

int orc_fake_quantize(const void* x, int x_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                      const int32_t* g_idx, void* out, int out_dt, int64_t rows, int64_t cols,
                      int64_t rdiv, int64_t cdiv, int64_t s_row_stride, int cd, int qtype, int bits) {
    return orc_fake_quantize_gs(x, x_dt, scale, s_dt, zp, zp_dt, g_idx, out, out_dt, rows, cols, rdiv, cdiv, s_row_stride, cd, qtype, bits, NULL, s_dt);
}

/* ------------------------------------------------------------------------- */
/* pack_bitmasks / unpack_bitmasks  (utils/helpers.py:306-343)                */
/* numpy.packbits(axis=-1, bitorder="little"): bit k of byte b <-> col 8b+k   */
/* ------------------------------------------------------------------------- */
int orc_pack_bitmasks(const uint8_t* bytemask, uint8_t* out, int64_t rows, int64_t cols) {
    int64_t nb = (cols + 7) / 8;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t b = 0; b < nb; ++b) {
            uint8_t v = 0;
            for (int k = 0; k < 8; ++k) {
                int64_t c = b * 8 + k;
                if (c < cols && bytemask[r * cols + c]) v |= (uint8_t)(1u << k);
            }
            out[r * nb + b] = v;
        }
    return ORC_OK;
}

int orc_unpack_bitmasks(const uint8_t* packed, uint8_t* bytemask, int64_t rows, int64_t cols) {
    int64_t nb = (cols + 7) / 8;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) bytemask[r * cols + c] = (packed[r * nb + (c >> 3)] >> (c & 7)) & 1u;
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* Sparse24BitMask compress / decompress -- RESTATED, parity unpinned.        */
/* The compressor is absent from the reference snapshot (config/base.py:18,   */
/* config/sparse_24_bitmask.py:16-29 are the remnants).  Restated from the    */
/* format's public description: for every 4 consecutive elements of a row     */
/* keep the 2 of largest |x| (ties -> lower column first), store kept values  */
/* in column order as [R, C/2] plus pack_bitmasks(mask) as uint8 [R, C/8].    */
/* Elements are moved as raw `esize`-byte patterns.  |x| is compared on the   */
/* value for float dtypes (dt) and on int8 for 1-byte payloads.               */
/* ------------------------------------------------------------------------- */
static inline float abs_key(const void* x, int64_t i, int dt) {
    float v = load_as_f32(x, i, dt);
    return fabsf(v);
}

int orc_sparse24_compress(const void* x, int dt, void* values, uint8_t* bitmask, int64_t rows, int64_t cols) {
    if (cols % 4 != 0) return ORC_E_SHAPE;
    int esize = (dt == DT_F32 || dt == DT_I32) ? 4 : ((dt == DT_F16 || dt == DT_BF16) ? 2 : 1);
    int64_t nb = (cols + 7) / 8;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        for (int64_t b = 0; b < nb; ++b) bitmask[r * nb + b] = 0;
        for (int64_t qd = 0; qd < cols / 4; ++qd) {
            int64_t base = r * cols + qd * 4;
            float k[4];
            for (int j = 0; j < 4; ++j) k[j] = abs_key(x, base + j, dt);
            /* rank: number of elements that beat j (larger key, or equal key at lower index) */
            int keep[4];
            for (int j = 0; j < 4; ++j) {
                int beat = 0;
                for (int m = 0; m < 4; ++m)
                    if (m != j && (k[m] > k[j] || (k[m] == k[j] && m < j))) ++beat;
                keep[j] = beat < 2;
            }
            int o = 0;
            for (int j = 0; j < 4; ++j)
                if (keep[j]) {
                    int64_t c = qd * 4 + j;
                    bitmask[r * nb + (c >> 3)] |= (uint8_t)(1u << (c & 7));
                    memcpy((char*)values + (r * (cols / 2) + qd * 2 + o) * esize, (const char*)x + (base + j) * esize, esize);
                    ++o;
                }
        }
    }
    return ORC_OK;
}

int orc_sparse24_decompress(const void* values, int dt, const uint8_t* bitmask, void* out, int64_t rows, int64_t cols) {
    if (cols % 4 != 0) return ORC_E_SHAPE;
    int esize = (dt == DT_F32 || dt == DT_I32) ? 4 : ((dt == DT_F16 || dt == DT_BF16) ? 2 : 1);
    int64_t nb = (cols + 7) / 8;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        int64_t o = 0;
        for (int64_t c = 0; c < cols; ++c) {
            int bit = (bitmask[r * nb + (c >> 3)] >> (c & 7)) & 1;
            char* dst = (char*)out + (r * cols + c) * esize;
            if (bit) {
                memcpy(dst, (const char*)values + (r * (cols / 2) + o) * esize, esize);
                ++o;
            } else {
                memset(dst, 0, esize);
            }
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* sparse-bitmask (unstructured) compress / decompress -- RESTATED, parity    */
/* unpinned (config/base.py:17, config/sparse_bitmask.py:12-25 remnants).     */
/* values = x[x != 0] in row-major order, bitmask = pack_bitmasks(x != 0),    */
/* row_offsets[r] = number of non-zeros in rows < r.  Returns nnz via *nnz.   */
/* "x != 0" is a value test: -0.0 counts as zero, NaN as non-zero.            */
/* ------------------------------------------------------------------------- */
static inline int is_nonzero(const void* x, int64_t i, int dt) {
    float v = load_as_f32(x, i, dt);
    return !(v == 0.0f);
}

int orc_bitmask_compress(const void* x, int dt, void* values, uint8_t* bitmask, int64_t* row_offsets, int64_t* nnz,
                         int64_t rows, int64_t cols) {
    int esize = (dt == DT_F32 || dt == DT_I32) ? 4 : ((dt == DT_F16 || dt == DT_BF16) ? 2 : 1);
    int64_t nb = (cols + 7) / 8;
    int64_t o = 0;
    for (int64_t r = 0; r < rows; ++r) {
        row_offsets[r] = o;
        for (int64_t b = 0; b < nb; ++b) bitmask[r * nb + b] = 0;
        for (int64_t c = 0; c < cols; ++c)
            if (is_nonzero(x, r * cols + c, dt)) {
                bitmask[r * nb + (c >> 3)] |= (uint8_t)(1u << (c & 7));
                if (values) memcpy((char*)values + o * esize, (const char*)x + (r * cols + c) * esize, esize);
                ++o;
            }
    }
    *nnz = o;
    return ORC_OK;
}

int orc_bitmask_decompress(const void* values, int dt, const uint8_t* bitmask, void* out, int64_t rows, int64_t cols) {
    int esize = (dt == DT_F32 || dt == DT_I32) ? 4 : ((dt == DT_F16 || dt == DT_BF16) ? 2 : 1);
    int64_t nb = (cols + 7) / 8;
    int64_t o = 0;
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            char* dst = (char*)out + (r * cols + c) * esize;
            if ((bitmask[r * nb + (c >> 3)] >> (c & 7)) & 1) {
                memcpy(dst, (const char*)values + o * esize, esize);
                ++o;
            } else {
                memset(dst, 0, esize);
            }
        }
    return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* fused hot-path compositions (what the compressors call back to back)       */
/* pack_quantized/base.py:96-104 (quantize -> pack_to_int32) and :147-161     */
/* (unpack_from_int32 -> dequantize).  `tmp` is caller scratch int8 [R*C].    */
/* ------------------------------------------------------------------------- */
int orc_quantize_pack(const void* x, int x_dt, const void* scale, int s_dt, const void* zp, int zp_dt,
                      const int32_t* g_idx, int32_t* packed, int8_t* tmp, int64_t rows, int64_t cols,
                      int64_t rdiv, int64_t cdiv, int64_t s_row_stride, int cd, int bits) {
    int rc = orc_quantize(x, x_dt, scale, s_dt, zp, zp_dt, g_idx, tmp, DT_I8, rows, cols, rdiv, cdiv, s_row_stride, cd, Q_INT, bits);
    if (rc) return rc;
    return orc_pack_int32(tmp, packed, rows, cols, bits, 1);
}

int orc_unpack_dequantize(const int32_t* packed, const void* scale, int s_dt, const void* zp, int zp_dt,
                          const int32_t* g_idx, void* out, int out_dt, int8_t* tmp, int64_t rows, int64_t cols,
                          int64_t rdiv, int64_t cdiv, int64_t s_row_stride, int bits) {
    int rc = orc_unpack_int32(packed, tmp, rows, cols, bits, 1);
    if (rc) return rc;
    return orc_dequantize(tmp, DT_I8, scale, s_dt, zp, zp_dt, g_idx, out, out_dt, rows, cols, rdiv, cdiv, s_row_stride);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline legs of bench.py ask for all host threads explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* 2:4 "semi-structured" (CUTLASS / marlin-24) value + metadata layout        */
/* utils/semi_structured_conversions.py:33-60 (meta reordering offsets),      */
/* :66-197 (from dense), :204-298 (to dense).  Pinned by golden vectors.      */
/*   dense [m, k]; non-fp32: quads of 4 elements, keep 2 -> sparse [m, k/2];  */
/*   fp32: pairs of 2, keep 1 -> sparse [m, k/2].  One 4-bit code per quad    */
/*   (idx0 | idx1 << 2); meta element = 4 (int16) or 8 (int32, int8 data)     */
/*   codes; meta stored reordered for ColumnMajorInterleaved<2>.              */
/* ------------------------------------------------------------------------- */
static int64_t meta_offset(int64_t r, int64_t c, int64_t m, int meta_bytes) {
    const int64_t gy = (meta_bytes == 2) ? 32 : 16;
    int64_t rr = r / 64 * 64 + (r % 2) * 2 + (r % 8) / 4 + ((r % gy) % 4) / 2 * 32 + ((r % 64) / 8) * 4;
    int tr = (rr % 2 == 0) && (c % 2 == 1);
    int bl = (rr % 2 == 1) && (c % 2 == 0);
    rr += tr - bl;
    int64_t cc = c - (tr - bl);
    return (cc / 2) * m * 2 + rr * 2 + (cc % 2);
}

static inline int elem_nonzero(const void* p, int64_t i, int dt) {
    switch (dt) {
    case DT_I8: return ((const int8_t*)p)[i] != 0;
    case DT_F16: case DT_BF16: return (((const uint16_t*)p)[i] & 0x7fffu) != 0;
    case DT_F32: return (((const uint32_t*)p)[i] & 0x7fffffffu) != 0;
    case DT_I32: return ((const int32_t*)p)[i] != 0;
    default: return 0;
    }
}

/* meta: int32 [m, k/32] for int8 data, else int16 [m, k/16] (fp32: [m, k/8]) */
int orc_semi_structured_from_dense(const void* dense, int dt, void* sparse, void* meta, int64_t m, int64_t k) {
    const int es = (dt == DT_F32 || dt == DT_I32) ? 4 : ((dt == DT_I8) ? 1 : 2);
    const int meta_bytes = (dt == DT_I8) ? 4 : 2;
    const int qpe = meta_bytes * 2;                 /* 4-bit codes per meta element */
    const int ks = (dt == DT_F32) ? 2 : 4;          /* dense elements per code */
    if (k % (ks * qpe) != 0) return ORC_E_SHAPE;
    const int64_t ncols = k / (ks * qpe);
    for (int64_t r = 0; r < m; ++r)
        for (int64_t c = 0; c < ncols; ++c) {
            uint32_t word = 0;
            for (int q = 0; q < qpe; ++q) {
                const int64_t base = r * k + (c * qpe + q) * ks;
                int m0, m1, m2, m3;
                if (ks == 4) {
                    m0 = elem_nonzero(dense, base, dt); m1 = elem_nonzero(dense, base + 1, dt);
                    m2 = elem_nonzero(dense, base + 2, dt); m3 = elem_nonzero(dense, base + 3, dt);
                } else {
                    m0 = m1 = elem_nonzero(dense, base, dt);
                    m2 = m3 = elem_nonzero(dense, base + 1, dt);
                }
                (void)m2;
                const int e0 = m0 & m1, e1 = (!m0) & m1, e2 = (!m0) & (!m1);
                const int bit0 = e1, bit1 = e2, bit2 = e0 | e2 | m3, bit3 = e1 | (!m1);
                const int idx0 = bit0 | (bit1 << 1), idx1 = bit2 | (bit3 << 1);
                word |= (uint32_t)(idx0 | (idx1 << 2)) << (4 * q);
                const int64_t so = r * (k / 2) + (c * qpe + q) * (ks / 2);
                if (ks == 4) {
                    memcpy((char*)sparse + so * es, (const char*)dense + (base + idx0) * es, es);
                    memcpy((char*)sparse + (so + 1) * es, (const char*)dense + (base + idx1) * es, es);
                } else {
                    memcpy((char*)sparse + so * es, (const char*)dense + (base + idx0 / 2) * es, es);
                }
            }
            const int64_t off = meta_offset(r, c, m, meta_bytes);
            if (meta_bytes == 2) ((int16_t*)meta)[off] = (int16_t)word;
            else ((int32_t*)meta)[off] = (int32_t)word;
        }
    return ORC_OK;
}

/* sparse [m, k] (k = sparse columns) -> dense [m, 2k] */
int orc_semi_structured_to_dense(const void* sparse, int dt, const void* meta, void* dense, int64_t m, int64_t k) {
    const int meta_bytes = (dt == DT_I8) ? 4 : 2;
    const int qpe = meta_bytes * 2;
    /* fp32 is processed as pairs of 16-bit halves (semi_structured_conversions.py:289-293) */
    const int es = (dt == DT_I8) ? 1 : 2;
    const int64_t kh = (dt == DT_F32 || dt == DT_I32) ? 2 * k : k;      /* sparse columns in es-sized units */
    if ((2 * kh) % (4 * qpe) != 0) return ORC_E_SHAPE;
    const int64_t ncols = 2 * kh / (4 * qpe);
    memset(dense, 0, (size_t)(m * 2 * kh * es));
    for (int64_t r = 0; r < m; ++r)
        for (int64_t c = 0; c < ncols; ++c) {
            const int64_t off = meta_offset(r, c, m, meta_bytes);
            const uint32_t word = (meta_bytes == 2) ? (uint32_t)(uint16_t)((const int16_t*)meta)[off] : (uint32_t)((const int32_t*)meta)[off];
            for (int q = 0; q < qpe; ++q) {
                const int idx0 = (word >> (4 * q)) & 3, idx1 = (word >> (4 * q + 2)) & 3;
                const int64_t quad = c * qpe + q;
                const int64_t dbase = r * 2 * kh + quad * 4, sbase = r * kh + quad * 2;
                memcpy((char*)dense + (dbase + idx0) * es, (const char*)sparse + sbase * es, es);
                memcpy((char*)dense + (dbase + idx1) * es, (const char*)sparse + (sbase + 1) * es, es);
            }
        }
    return ORC_OK;
}
