/*
 * ct_oracle_fp4.c -- TEST INFRASTRUCTURE ONLY (part of libct_oracle.so, see ct_oracle.c header).
 *
 * CPU restatement of the reference's FP4 (E2M1) and MX (E8M0 scale) pieces; reference paths are under
 * /root/reference/src/compressed_tensors:
 *   cast_to_fp4                      quantization/utils/fp4_utils.py:77-98
 *   pack_fp4_to_uint8                compressors/nvfp4/helpers.py:108-158
 *   unpack_fp4_from_uint8            compressors/nvfp4/helpers.py:162-193
 *   compress_mx_scale / decompress   compressors/mx_utils.py:18-44
 * The quantize / dequantize / fake_quantize arithmetic with FP4 rounding and a global scale lives in
 * ct_oracle.c (orc_*_gs, Q_FP4).  Pinned by tests/golden/fp4.pt.gz (tests/golden/make_golden_fp4.py).
 */

/* cast_to_fp4 on one value already rounded to its tensor dtype (every constant and result below is
 * exact in bf16 / fp16 / fp32, so no further rounding happens in the reference either):
 *   sign = torch.sign(x)  (-1, 0, +1; sign(+-0) = +0; sign(NaN) = NaN)
 *   a = |x| snapped by the closed/open interval ladder of :89-96, then  a * sign  (:97)            */
float orc_fp4_round(float v) {
    if (v != v) return v;
    float a = fabsf(v), r;
    if (a <= 0.25f) r = 0.0f;
    else if (a < 0.75f) r = 0.5f;
    else if (a <= 1.25f) r = 1.0f;
    else if (a < 1.75f) r = 1.5f;
    else if (a <= 2.5f) r = 2.0f;
    else if (a < 3.5f) r = 3.0f;
    else if (a <= 5.0f) r = 4.0f;
    else r = 6.0f;
    if (v > 0.0f) return r;
    if (v < 0.0f) return -r;      /* 0 * -1 = -0.0 for small negatives */
    return 0.0f;                  /* +-0 * sign(+-0) = +0.0 */
}

int orc_cast_to_fp4(const void* x, int dt, void* out, int64_t n) {
    if (!is_float_dt(dt)) return ORC_E_DTYPE;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) store_from_f32(out, i, dt, orc_fp4_round(load_as_f32(x, i, dt)));
    return ORC_OK;
}

/* nibble of one VALID fp4 value: index of |x| in {0,.5,1,1.5,2,3,4,6} via (x*2).to(int8).abs() equality
 * tests (:141-151; anything else maps to 0), bit 3 = torch.signbit(x) (so -0.0 -> 8) */
static inline uint8_t fp4_nibble(float v, int dt) {
    uint8_t sign = (f2u(v) >> 31) & 1u;
    float d = rnd(v * 2.0f, dt);
    int a = (int)(int8_t)(int32_t)d;
    if (a < 0) a = -a;
    uint8_t idx = 0;
    if (a == 1) idx = 1; else if (a == 2) idx = 2; else if (a == 3) idx = 3; else if (a == 4) idx = 4;
    else if (a == 6) idx = 5; else if (a == 8) idx = 6; else if (a >= 12) idx = 7;
    return (uint8_t)(idx | (sign << 3));
}

/* x [rows, cols] (cols even) -> uint8 [rows, cols/2]; element 2j in the low nibble (:154-156) */
int orc_pack_fp4(const void* x, int dt, uint8_t* out, int64_t rows, int64_t cols) {
    if (!is_float_dt(dt)) return ORC_E_DTYPE;
    if (cols % 2) return ORC_E_SHAPE;
    int64_t n = rows * cols / 2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        out[i] = (uint8_t)(fp4_nibble(load_as_f32(x, 2 * i, dt), dt) | (fp4_nibble(load_as_f32(x, 2 * i + 1, dt), dt) << 4));
    return ORC_OK;
}

static const float ORC_E2M1[8] = {0.0f, 0.5f, 1.0f, 1.5f, 2.0f, 3.0f, 4.0f, 6.0f};

/* uint8 [rows, cols/2] -> out_dt [rows, cols]: kE2M1[n & 7] * (n & 8 ? -1 : 1), so nibble 8 is -0.0 (:176-190) */
int orc_unpack_fp4(const uint8_t* in, void* out, int out_dt, int64_t rows, int64_t cols) {
    if (!is_float_dt(out_dt)) return ORC_E_DTYPE;
    if (cols % 2) return ORC_E_SHAPE;
    int64_t n = rows * cols / 2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t b = in[i];
        float lo = ORC_E2M1[b & 7] * ((b & 8) ? -1.0f : 1.0f);
        float hi = ORC_E2M1[(b >> 4) & 7] * ((b & 0x80) ? -1.0f : 1.0f);
        store_from_f32(out, 2 * i, out_dt, lo);
        store_from_f32(out, 2 * i + 1, out_dt, hi);
    }
    return ORC_OK;
}

/* compress_mx_scale (mx_utils.py:30-31): 127 + floor(log2(scale)) with log2 and floor evaluated in the
 * scale's dtype, .to(int32), .to(uint8) (wraps).  For the power-of-two scales the reference's own flow
 * produces (calculate_qparams' MX branch) log2 is exact; for other values the result depends on the
 * host libm's log2f in the last ulp when log2(scale) is within an ulp of an integer. */
int orc_mx_scale_compress(const void* scale, int dt, uint8_t* out, int64_t n) {
    if (!is_float_dt(dt)) return ORC_E_DTYPE;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float l = floorf(rnd(log2f(load_as_f32(scale, i, dt)), dt));
        int32_t e = (l != l || isinf(l)) ? (int32_t)0x80000000 : (int32_t)l;   /* x86 cvttss2si of NaN / inf */
        out[i] = (uint8_t)(uint32_t)(127 + (int64_t)e);
    }
    return ORC_OK;
}

/* decompress_mx_scale (mx_utils.py:43-44): 2.0 ** (e - 127).to(bfloat16) -> bfloat16 (pow in float, rounded to bf16) */
int orc_mx_scale_decompress(const uint8_t* in, uint16_t* out_bf16, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float e = bf16_to_f32(f32_to_bf16((float)((int32_t)in[i] - 127)));
        out_bf16[i] = f32_to_bf16(exp2f(e));
    }
    return ORC_OK;
}
