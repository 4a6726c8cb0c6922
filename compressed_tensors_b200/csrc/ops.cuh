// ops.cuh -- the per-chunk functors (8 elements per call) plugged into the streaming pipelines.
//
// Fast-path contract (checked on the host in dispatch.cu): x, scale and compute dtype are the
// same float dtype P; zero point absent (ZP=0) or int8 (ZP=1); scale index = chunk / dc
// (TENSOR, CHANNEL with cols % 8 == 0, GROUP with cols % group == 0 and group % 8 == 0).
#pragma once

#include "quant_core.cuh"
#include "stream.cuh"

namespace ctb {

struct F32 {
    static constexpr int DT = CT_F32;
};

enum QKind { QI_NARROW = 0, QI_WIDE = 1, QF8 = 2 };

template <class P> struct ElemBytes { static constexpr int v = 2; };
template <> struct ElemBytes<F32> { static constexpr int v = 4; };

// ------------------------------------------------------------------------------------
// scale / zero-point context
// ------------------------------------------------------------------------------------
template <class P> __device__ __forceinline__ float load_scale_f32(const void* scale, uint32_t i) {
    if constexpr (P::DT == CT_BF16) {
        return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(scale) + i) << 16);
    } else if constexpr (P::DT == CT_F16) {
        return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(scale) + i)));
    } else {
        return __ldg(reinterpret_cast<const float*>(scale) + i);
    }
}
template <class P> __device__ __forceinline__ uint32_t load_scale_raw2(const void* scale, uint32_t i) {
    uint32_t h = __ldg(reinterpret_cast<const unsigned short*>(scale) + i);
    return h | (h << 16);
}

template <class P, int ZP> struct QuantCtx {
    ScaleCtx sc;
    uint32_t zp2;   // 16-bit P: zero point as T duplicated; F32: float bits
};
template <class P, int ZP>
__device__ __forceinline__ QuantCtx<P, ZP> quant_prefetch(const Job& J, uint32_t gc) {
    QuantCtx<P, ZP> c;
    const uint32_t si = fd_div(gc, J.dc);
    c.sc = make_scale_ctx(load_scale_f32<P>(J.scale, si));
    c.zp2 = 0;
    if constexpr (ZP == 1) {
        const float z = (float)__ldg(reinterpret_cast<const int8_t*>(J.zp) + si);
        if constexpr (P::DT == CT_F32) c.zp2 = __float_as_uint(z);
        else c.zp2 = dup2<P>(z);
    }
    return c;
}

template <class P, int ZP> struct DequantCtx {
    uint32_t s2;    // 16-bit P: raw scale duplicated; F32: float bits
    uint32_t zp2;
};
template <class P, int ZP>
__device__ __forceinline__ DequantCtx<P, ZP> dequant_prefetch(const Job& J, uint32_t gc) {
    DequantCtx<P, ZP> c;
    const uint32_t si = fd_div(gc, J.dc);
    if constexpr (P::DT == CT_F32) c.s2 = __float_as_uint(__ldg(reinterpret_cast<const float*>(J.scale) + si));
    else c.s2 = load_scale_raw2<P>(J.scale, si);
    c.zp2 = 0;
    if constexpr (ZP == 1) {
        const float z = (float)__ldg(reinterpret_cast<const int8_t*>(J.zp) + si);
        if constexpr (P::DT == CT_F32) c.zp2 = __float_as_uint(z);
        else c.zp2 = dup2<P>(z);
    }
    return c;
}

// fp32 compute dtype: clamped, un-rounded value
template <int ZP>
__device__ __forceinline__ float scaled_clamped_f32(float x, const ScaleCtx& c, uint32_t zpbits, const Common& cm) {
    float t = __fdiv_rn(x, c.s);
    if (ZP) t = __fadd_rn(t, __uint_as_float(zpbits));
    return fminf(fmaxf(t, cm.qmin), cm.qmax);
}

// rounded (integer-valued, sign of zero preserved) pair in T
template <class P, bool WIDE>
__device__ __forceinline__ uint32_t round_int2(uint32_t t) {
    if constexpr (!WIDE) return round_keep_sign2<P>(t);
    else return P::pack(rintf(P::lo(t)), rintf(P::hi(t)));
}
// fp8-grid pair in T: T -> e4m3 (RNE) -> T
template <class P>
__device__ __forceinline__ uint32_t round_fp8_2(uint32_t t) {
    uint32_t h2 = e4m3x2_to_f16x2(f32x2_to_e4m3x2(P::lo(t), P::hi(t)));
    if constexpr (P::DT == CT_F16) return h2;
    else return P::pack(__low2float(*reinterpret_cast<__half2*>(&h2)), __high2float(*reinterpret_cast<__half2*>(&h2)));
}

// unsigned 4-bit codes of 4 rounded pairs -> one packed word (element 0 in bits 0-3)
__device__ __forceinline__ uint32_t nibbles_to_word(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3) {
    uint32_t a = __byte_perm(t0, t1, 0x6420) & 0x0f0f0f0fu;
    uint32_t b = __byte_perm(t2, t3, 0x6420) & 0x0f0f0f0fu;
    a |= a >> 4;
    b |= b >> 4;
    return __byte_perm(a, b, 0x6420) ^ 0x88888888u;
}

// ------------------------------------------------------------------------------------
// QuantPack: x (8 x T) -> BITS*8 bits of the int32 bitstream
//   replaces quantize(dtype=int8) + pack_to_int32 (pack_quantized/base.py:96-104)
// ------------------------------------------------------------------------------------
template <class P, int BITS, int ZP>
struct QuantPackOp {
    static_assert(BITS == 4 || BITS == 8, "fast path packs 4- and 8-bit codes");
    static constexpr int IN_BYTES = 8 * ElemBytes<P>::v;
    using Ctx = QuantCtx<P, ZP>;
    __device__ static __forceinline__ Ctx prefetch(const Job& J, const Common&, uint32_t gc) { return quant_prefetch<P, ZP>(J, gc); }

    __device__ static __forceinline__ void run(const Job& J, const Common& cm, const Ctx& c, uint32_t gc, const uint32_t (&w)[IN_BYTES / 4]) {
        if (c.sc.slow) body<true>(J, cm, c, gc, w);
        else body<false>(J, cm, c, gc, w);
    }
    template <bool SLOW>
    __device__ static __forceinline__ void body(const Job& J, const Common& cm, const Ctx& c, uint32_t gc, const uint32_t (&w)[IN_BYTES / 4]) {
        if constexpr (P::DT == CT_F32) {
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t = scaled_clamped_f32<ZP>(__uint_as_float(w[k]), c.sc, c.zp2, cm);
                const uint32_t u = (uint32_t)(rint_magic_f32(t) + (1 << (BITS - 1))) & ((1u << BITS) - 1u);
                if (BITS == 4) lo |= u << (4 * k);
                else if (k < 4) lo |= u << (8 * k);
                else hi |= u << (8 * (k - 4));
            }
            if (BITS == 4) stg_stream4(J.out + (size_t)gc * 4, lo);
            else stg_stream8(J.out + (size_t)gc * 8, make_uint2(lo, hi));
        } else if constexpr (BITS == 4) {
            uint32_t t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = round_magic2<P>(scaled_clamped2<P, ZP != 0, SLOW>(w[k], c.sc, c.zp2, cm.qmin2, cm.qmax2));
            stg_stream4(J.out + (size_t)gc * 4, nibbles_to_word(t[0], t[1], t[2], t[3]));
        } else {
            uint32_t b[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = scaled_clamped2<P, ZP != 0, SLOW>(w[k], c.sc, c.zp2, cm.qmin2, cm.qmax2);
                b[2 * k] = (uint32_t)(rint_magic_f32(P::lo(t)) + 128) & 0xffu;
                b[2 * k + 1] = (uint32_t)(rint_magic_f32(P::hi(t)) + 128) & 0xffu;
            }
            const uint32_t lo = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
            const uint32_t hi = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
            stg_stream8(J.out + (size_t)gc * 8, make_uint2(lo, hi));
        }
    }
};

// ------------------------------------------------------------------------------------
// Quantize: x (8 x T) -> 8 one-byte codes (int8 or float8_e4m3fn)
//   replaces quantize(dtype=args.pytorch_dtype()) (naive_quantized/base.py:79-86)
// ------------------------------------------------------------------------------------
template <class P, int KIND, int ZP>
struct QuantizeOp {
    static constexpr int IN_BYTES = 8 * ElemBytes<P>::v;
    using Ctx = QuantCtx<P, ZP>;
    __device__ static __forceinline__ Ctx prefetch(const Job& J, const Common&, uint32_t gc) { return quant_prefetch<P, ZP>(J, gc); }

    __device__ static __forceinline__ void run(const Job& J, const Common& cm, const Ctx& c, uint32_t gc, const uint32_t (&w)[IN_BYTES / 4]) {
        if (c.sc.slow) body<true>(J, cm, c, gc, w);
        else body<false>(J, cm, c, gc, w);
    }
    template <bool SLOW>
    __device__ static __forceinline__ void body(const Job& J, const Common& cm, const Ctx& c, uint32_t gc, const uint32_t (&w)[IN_BYTES / 4]) {
        uint32_t b[8];
        if constexpr (P::DT == CT_F32) {
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const float t0 = scaled_clamped_f32<ZP>(__uint_as_float(w[k]), c.sc, c.zp2, cm);
                const float t1 = scaled_clamped_f32<ZP>(__uint_as_float(w[k + 1]), c.sc, c.zp2, cm);
                if constexpr (KIND == QF8) {
                    const uint32_t e = f32x2_to_e4m3x2(t0, t1);
                    b[k] = e & 0xffu;
                    b[k + 1] = e >> 8;
                } else {
                    b[k] = (uint32_t)rint_magic_f32(t0) & 0xffu;
                    b[k + 1] = (uint32_t)rint_magic_f32(t1) & 0xffu;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = scaled_clamped2<P, ZP != 0, SLOW>(w[k], c.sc, c.zp2, cm.qmin2, cm.qmax2);
                if constexpr (KIND == QF8) {
                    const uint32_t e = f32x2_to_e4m3x2(P::lo(t), P::hi(t));
                    b[2 * k] = e & 0xffu;
                    b[2 * k + 1] = e >> 8;
                } else {
                    b[2 * k] = (uint32_t)rint_magic_f32(P::lo(t)) & 0xffu;
                    b[2 * k + 1] = (uint32_t)rint_magic_f32(P::hi(t)) & 0xffu;
                }
            }
        }
        const uint32_t lo = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        const uint32_t hi = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
        stg_stream8(J.out + (size_t)gc * 8, make_uint2(lo, hi));
    }
};

// ------------------------------------------------------------------------------------
// dequantization tail shared by Dequantize / UnpackDequant / FakeQuant (16-bit P)
//   v = q.to(T); v -= zp.to(T); v *= scale     (forward_helpers.py:562-567)
// ------------------------------------------------------------------------------------
template <class P, int ZP>
__device__ __forceinline__ uint32_t dq_tail2(uint32_t v, uint32_t zp2, uint32_t s2) {
    if (ZP) v = sub2<P>(v, zp2);
    return mul2<P>(v, s2);
}
template <int ZP>
__device__ __forceinline__ float dq_tail_f32(float v, uint32_t zpbits, uint32_t sbits) {
    if (ZP) v = __fsub_rn(v, __uint_as_float(zpbits));
    return __fmul_rn(v, __uint_as_float(sbits));
}

template <class P>
__device__ __forceinline__ void store_out8(uint8_t* out, uint32_t gc, const uint32_t (&o)[8 * ElemBytes<P>::v / 4]) {
    if constexpr (P::DT == CT_F32) {
        stg_stream16(out + (size_t)gc * 32, make_uint4(o[0], o[1], o[2], o[3]));
        stg_stream16(out + (size_t)gc * 32 + 16, make_uint4(o[4], o[5], o[6], o[7]));
    } else {
        stg_stream16(out + (size_t)gc * 16, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// unsigned byte u (0..255) -> exact float u - bias, via the 2^23 exponent trick
__device__ __forceinline__ float ubyte_to_f32(uint32_t word, int k, float bias_plus_2p23) {
    // selector picks byte k of `word` as byte 0, bytes 1..3 from 0x4B000000
    const uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)k);
    return __fsub_rn(__uint_as_float(bits), bias_plus_2p23);
}

// ------------------------------------------------------------------------------------
// Dequantize: 8 one-byte codes -> 8 x T       (naive_quantized/base.py:119-124)
// ------------------------------------------------------------------------------------
template <class P, int KIND /* QI_* = int8 codes, QF8 = e4m3 codes */, int ZP>
struct DequantizeOp {
    static constexpr int IN_BYTES = 8;
    using Ctx = DequantCtx<P, ZP>;
    __device__ static __forceinline__ Ctx prefetch(const Job& J, const Common&, uint32_t gc) { return dequant_prefetch<P, ZP>(J, gc); }

    __device__ static __forceinline__ void run(const Job& J, const Common&, const Ctx& c, uint32_t gc, const uint32_t (&w)[2]) {
        float f[8];
        if constexpr (KIND == QF8) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t h2 = e4m3x2_to_f16x2((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
                f[2 * k] = __low2float(*reinterpret_cast<__half2*>(&h2));
                f[2 * k + 1] = __high2float(*reinterpret_cast<__half2*>(&h2));
            }
        } else {
            const uint32_t w0 = w[0] ^ 0x80808080u, w1 = w[1] ^ 0x80808080u;  // int8 -> offset-128 unsigned
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[k] = ubyte_to_f32(w0, k, 8388608.0f + 128.0f);
                f[4 + k] = ubyte_to_f32(w1, k, 8388608.0f + 128.0f);
            }
        }
        uint32_t o[8 * ElemBytes<P>::v / 4];
        if constexpr (P::DT == CT_F32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(dq_tail_f32<ZP>(f[k], c.zp2, c.s2));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = dq_tail2<P, ZP>(P::pack(f[2 * k], f[2 * k + 1]), c.zp2, c.s2);
        }
        store_out8<P>(J.out, gc, o);
    }
};

// ------------------------------------------------------------------------------------
// UnpackDequant: BITS*8 bits of the int32 bitstream -> 8 x T
//   replaces unpack_from_int32 + dequantize (pack_quantized/base.py:159-166)
// ------------------------------------------------------------------------------------
template <class P, int BITS, int ZP>
struct UnpackDequantOp {
    static_assert(BITS == 4 || BITS == 8, "fast path unpacks 4- and 8-bit codes");
    static constexpr int IN_BYTES = BITS;
    using Ctx = DequantCtx<P, ZP>;
    __device__ static __forceinline__ Ctx prefetch(const Job& J, const Common&, uint32_t gc) { return dequant_prefetch<P, ZP>(J, gc); }

    __device__ static __forceinline__ void run(const Job& J, const Common&, const Ctx& c, uint32_t gc, const uint32_t (&w)[BITS / 4]) {
        uint32_t o[8 * ElemBytes<P>::v / 4];
        if constexpr (BITS == 4 && P::DT != CT_F32) {
            // nibble u = n + 8 in [0, 15]; (EXP | u) is the T value EXPVAL + u exactly; subtract EXPVAL + 8
            const uint32_t off2 = P::OFF8_2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo = (w[0] >> (8 * k)) & 0xfu;
                const uint32_t hi = (w[0] >> (8 * k + 4)) & 0xfu;
                const uint32_t p = P::ONE_TWENTY_EIGHT2 | lo | (hi << 16);
                o[k] = dq_tail2<P, ZP>(sub2<P>(p, off2), c.zp2, c.s2);
            }
        } else {
            float f[8];
            if constexpr (BITS == 4) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = (float)((int)((w[0] >> (4 * k)) & 0xfu) - 8);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    f[k] = ubyte_to_f32(w[0], k, 8388608.0f + 128.0f);
                    f[4 + k] = ubyte_to_f32(w[BITS / 4 - 1], k, 8388608.0f + 128.0f);
                }
            }
            if constexpr (P::DT == CT_F32) {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(dq_tail_f32<ZP>(f[k], c.zp2, c.s2));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = dq_tail2<P, ZP>(P::pack(f[2 * k], f[2 * k + 1]), c.zp2, c.s2);
            }
        }
        store_out8<P>(J.out, gc, o);
    }
};

// ------------------------------------------------------------------------------------
// FakeQuant: x (8 x T) -> quantize -> dequantize -> 8 x T   (forward_helpers.py:180-215)
// ------------------------------------------------------------------------------------
template <class P, int ZP> struct FakeCtx {
    ScaleCtx sc;
    uint32_t s2;
    uint32_t zp2;
};
template <class P, int KIND, int ZP>
struct FakeQuantOp {
    static constexpr int IN_BYTES = 8 * ElemBytes<P>::v;
    using Ctx = FakeCtx<P, ZP>;
    __device__ static __forceinline__ Ctx prefetch(const Job& J, const Common&, uint32_t gc) {
        Ctx c;
        const uint32_t si = fd_div(gc, J.dc);
        const float s = load_scale_f32<P>(J.scale, si);
        c.sc = make_scale_ctx(s);
        if constexpr (P::DT == CT_F32) c.s2 = __float_as_uint(s);
        else c.s2 = load_scale_raw2<P>(J.scale, si);
        c.zp2 = 0;
        if constexpr (ZP == 1) {
            const float z = (float)__ldg(reinterpret_cast<const int8_t*>(J.zp) + si);
            if constexpr (P::DT == CT_F32) c.zp2 = __float_as_uint(z);
            else c.zp2 = dup2<P>(z);
        }
        return c;
    }
    __device__ static __forceinline__ void run(const Job& J, const Common& cm, const Ctx& c, uint32_t gc, const uint32_t (&w)[IN_BYTES / 4]) {
        if (c.sc.slow) body<true>(J, cm, c, gc, w);
        else body<false>(J, cm, c, gc, w);
    }
    template <bool SLOW>
    __device__ static __forceinline__ void body(const Job& J, const Common& cm, const Ctx& c, uint32_t gc, const uint32_t (&w)[IN_BYTES / 4]) {
        uint32_t o[IN_BYTES / 4];
        if constexpr (P::DT == CT_F32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float t = scaled_clamped_f32<ZP>(__uint_as_float(w[k]), c.sc, c.zp2, cm);
                if constexpr (KIND == QF8) t = e4m3_to_f32(f32_to_e4m3_byte(t));
                else t = rintf(t);
                o[k] = __float_as_uint(dq_tail_f32<ZP>(t, c.zp2, c.s2));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t t = scaled_clamped2<P, ZP != 0, SLOW>(w[k], c.sc, c.zp2, cm.qmin2, cm.qmax2);
                if constexpr (KIND == QF8) t = round_fp8_2<P>(t);
                else t = round_int2<P, KIND == QI_WIDE>(t);
                o[k] = dq_tail2<P, ZP>(t, c.zp2, c.s2);
            }
        }
        store_out8<P>(J.out, gc, o);
    }
};

// ------------------------------------------------------------------------------------
// standalone bit packing of int8 codes (compressors/pack_quantized/helpers.py:20-180), 4 and 8 bits
// ------------------------------------------------------------------------------------
struct NoCtx {};
template <int BITS>
struct PackOp {
    static constexpr int IN_BYTES = 8;
    using Ctx = NoCtx;
    __device__ static __forceinline__ Ctx prefetch(const Job&, const Common&, uint32_t) { return {}; }
    __device__ static __forceinline__ void run(const Job& J, const Common&, const Ctx&, uint32_t gc, const uint32_t (&w)[2]) {
        if constexpr (BITS == 8) {
            stg_stream8(J.out + (size_t)gc * 8, make_uint2(w[0] ^ 0x80808080u, w[1] ^ 0x80808080u));
        } else {
            // word = sum_j (v_j + 8) << 4j  (mod 2^32), the reference's scatter_add semantics, which
            // also fixes the result for out-of-range int8 inputs.  dp4a forms v0 + 16*v1 etc.
            const int s01 = __dp4a((int)w[0], 0x00001001, 0);   // weights (1, 16, 0, 0)
            const int s23 = __dp4a((int)w[0], 0x10010000, 0);   // weights (0, 0, 1, 16)
            const int s45 = __dp4a((int)w[1], 0x00001001, 0);
            const int s67 = __dp4a((int)w[1], 0x10010000, 0);
            const uint32_t word = (uint32_t)s01 + ((uint32_t)s23 << 8) + ((uint32_t)s45 << 16) + ((uint32_t)s67 << 24) + 0x88888888u;
            stg_stream4(J.out + (size_t)gc * 4, word);
        }
    }
};
template <int BITS>
struct UnpackOp {
    static constexpr int IN_BYTES = BITS;
    using Ctx = NoCtx;
    __device__ static __forceinline__ Ctx prefetch(const Job&, const Common&, uint32_t) { return {}; }
    __device__ static __forceinline__ void run(const Job& J, const Common&, const Ctx&, uint32_t gc, const uint32_t (&w)[BITS / 4]) {
        if constexpr (BITS == 8) {
            stg_stream8(J.out + (size_t)gc * 8, make_uint2(w[0] ^ 0x80808080u, w[BITS / 4 - 1] ^ 0x80808080u));
        } else {
            const uint32_t lo = w[0] & 0x0f0f0f0fu, hi = (w[0] >> 4) & 0x0f0f0f0fu;
            uint32_t a = __byte_perm(lo, hi, 0x5140);   // elements 0..3 (codes 0..15 per byte)
            uint32_t b = __byte_perm(lo, hi, 0x7362);   // elements 4..7
            // code - 8 per byte without inter-byte borrows: keep low 3 bits, set 0xF8 when bit 3 is clear
            a = (a & 0x07070707u) | ((~a & 0x08080808u) * 31u);
            b = (b & 0x07070707u) | ((~b & 0x08080808u) * 31u);
            stg_stream8(J.out + (size_t)gc * 8, make_uint2(a, b));
        }
    }
};

}  // namespace ctb
