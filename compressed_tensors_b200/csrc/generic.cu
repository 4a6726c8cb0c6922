// generic.cu -- shape- and dtype-generic kernels: every strategy (tensor / channel / token /
// group incl. g_idx and one-row scales / block with implicit padding), ragged row lengths, all
// bit widths 1..8, packed_dim 0 and 1, mixed x / scale / zero-point dtypes.  These are the
// catch-all behind the streaming fast path (dispatch.cu decides); they are coalesced and
// bit-exact but not tuned for the roofline.
#include "generic.cuh"

namespace ctb {

__global__ void __launch_bounds__(256) generic_quant_kernel(const __grid_constant__ GParams p, int mode) {
    const int64_t n = p.rows * p.cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / p.cols, c = i - r * p.cols;
        if (mode == G_QUANTIZE) {
            store_from_f32(p.out, i, p.q_dt, quant_at(p, r, c));
        } else if (mode == G_DEQUANTIZE) {
            store_from_f32(p.out, i, p.out_dt, dequant_at(p, load_as_f32(p.in, i, p.q_dt), r, c));
        } else {
            store_from_f32(p.out, i, p.out_dt, dequant_at(p, quant_at(p, r, c), r, c));
        }
    }
}

// packed_dim == 1: in [rows, cols] -> out [rows, nw]; thread = (row, group)
template <int BITS>
__global__ void __launch_bounds__(256) pack_dim1_kernel(const int8_t* __restrict__ in, int32_t* __restrict__ out,
                                                        int64_t rows, int64_t cols, int64_t nw) {
    const int64_t groups = (cols + 31) / 32;
    const int64_t total = rows * groups;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / groups, g = t - r * groups;
        const int64_t c0 = g * 32;
        const int nvalid = (int)min((int64_t)32, cols - c0);
        const int8_t* src = in + r * cols + c0;
        uint32_t words[BITS];
        pack_group<BITS>(words, nvalid, [&](int j) { return (int32_t)src[j]; });
#pragma unroll
        for (int k = 0; k < BITS; ++k)
            if (g * BITS + k < nw) out[r * nw + g * BITS + k] = (int32_t)words[k];
    }
}

// packed_dim == 0: in [rows, cols] packs down the rows -> out [nw, cols]; thread = (group, col)
template <int BITS>
__global__ void __launch_bounds__(256) pack_dim0_kernel(const int8_t* __restrict__ in, int32_t* __restrict__ out,
                                                        int64_t rows, int64_t cols, int64_t nw) {
    const int64_t groups = (rows + 31) / 32;
    const int64_t total = groups * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = t / cols, c = t - g * cols;
        const int64_t r0 = g * 32;
        const int nvalid = (int)min((int64_t)32, rows - r0);
        uint32_t words[BITS];
        pack_group<BITS>(words, nvalid, [&](int j) { return (int32_t)in[(r0 + j) * cols + c]; });
#pragma unroll
        for (int k = 0; k < BITS; ++k)
            if (g * BITS + k < nw) out[(g * BITS + k) * cols + c] = (int32_t)words[k];
    }
}

template <int BITS>
__global__ void __launch_bounds__(256) unpack_dim1_kernel(const int32_t* __restrict__ in, int8_t* __restrict__ out,
                                                          int64_t rows, int64_t cols, int64_t nw) {
    const int64_t groups = (cols + 31) / 32;
    const int64_t total = rows * groups;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / groups, g = t - r * groups;
        const int64_t c0 = g * 32;
        const int nvalid = (int)min((int64_t)32, cols - c0);
        uint32_t words[BITS];
#pragma unroll
        for (int k = 0; k < BITS; ++k) words[k] = (g * BITS + k < nw) ? (uint32_t)in[r * nw + g * BITS + k] : 0u;
        int8_t* dst = out + r * cols + c0;
        unpack_group<BITS>(words, nvalid, [&](int j, int v) { dst[j] = (int8_t)v; });
    }
}

template <int BITS>
__global__ void __launch_bounds__(256) unpack_dim0_kernel(const int32_t* __restrict__ in, int8_t* __restrict__ out,
                                                          int64_t rows, int64_t cols, int64_t nw) {
    const int64_t groups = (rows + 31) / 32;
    const int64_t total = groups * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = t / cols, c = t - g * cols;
        const int64_t r0 = g * 32;
        const int nvalid = (int)min((int64_t)32, rows - r0);
        uint32_t words[BITS];
#pragma unroll
        for (int k = 0; k < BITS; ++k) words[k] = (g * BITS + k < nw) ? (uint32_t)in[(g * BITS + k) * cols + c] : 0u;
        unpack_group<BITS>(words, nvalid, [&](int j, int v) { out[(r0 + j) * cols + c] = (int8_t)v; });
    }
}

// fused generic quantize -> pack (packed_dim 1) and unpack -> dequantize
template <int BITS>
__global__ void __launch_bounds__(256) quantpack_generic_kernel(const __grid_constant__ GParams p, int64_t nw) {
    const int64_t groups = (p.cols + 31) / 32;
    const int64_t total = p.rows * groups;
    int32_t* out = reinterpret_cast<int32_t*>(p.out);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / groups, g = t - r * groups;
        const int64_t c0 = g * 32;
        const int nvalid = (int)min((int64_t)32, p.cols - c0);
        uint32_t words[BITS];
        pack_group<BITS>(words, nvalid, [&](int j) {
            const float q = quant_at(p, r, c0 + j);
            return (q != q) ? 0 : (int32_t)q;   // .to(int8)
        });
#pragma unroll
        for (int k = 0; k < BITS; ++k)
            if (g * BITS + k < nw) out[r * nw + g * BITS + k] = (int32_t)words[k];
    }
}

template <int BITS>
__global__ void __launch_bounds__(256) unpackdeq_generic_kernel(const __grid_constant__ GParams p, int64_t nw) {
    const int64_t groups = (p.cols + 31) / 32;
    const int64_t total = p.rows * groups;
    const int32_t* in = reinterpret_cast<const int32_t*>(p.in);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / groups, g = t - r * groups;
        const int64_t c0 = g * 32;
        const int nvalid = (int)min((int64_t)32, p.cols - c0);
        uint32_t words[BITS];
#pragma unroll
        for (int k = 0; k < BITS; ++k) words[k] = (g * BITS + k < nw) ? (uint32_t)in[r * nw + g * BITS + k] : 0u;
        unpack_group<BITS>(words, nvalid, [&](int j, int v) {
            store_from_f32(p.out, r * p.cols + c0 + j, p.out_dt, dequant_at(p, (float)v, r, c0 + j));
        });
    }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int launch_generic_quant(int mode, const ct_quant_desc& d, const void* in, const void* scale, const void* zp,
                         const int32_t* g_idx, void* out, cudaStream_t stream) {
    const int64_t n = d.rows * d.cols;
    if (n == 0) return CT_OK;
    GParams p = make_params(d, in, scale, zp, g_idx, out);
    generic_quant_kernel<<<grid_for(n), 256, 0, stream>>>(p, mode);
    count_launch();
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}

#define BITS_SWITCH(bits, EXPR)                 \
    switch (bits) {                             \
    case 1: { constexpr int B = 1; EXPR; } break; \
    case 2: { constexpr int B = 2; EXPR; } break; \
    case 3: { constexpr int B = 3; EXPR; } break; \
    case 4: { constexpr int B = 4; EXPR; } break; \
    case 5: { constexpr int B = 5; EXPR; } break; \
    case 6: { constexpr int B = 6; EXPR; } break; \
    case 7: { constexpr int B = 7; EXPR; } break; \
    case 8: { constexpr int B = 8; EXPR; } break; \
    default: set_error("num_bits %d outside [1, 8]", bits); return CT_E_BITS; \
    }

int launch_generic_quantpack(const ct_quant_desc& d, const void* x, const void* scale, const void* zp,
                             const int32_t* g_idx, int32_t* packed, cudaStream_t stream) {
    if (d.rows * d.cols == 0) return CT_OK;
    GParams p = make_params(d, x, scale, zp, g_idx, packed);
    const int64_t nw = (d.cols * d.num_bits + 31) / 32;
    const int64_t items = d.rows * ((d.cols + 31) / 32);
    BITS_SWITCH(d.num_bits, (quantpack_generic_kernel<B><<<grid_for(items), 256, 0, stream>>>(p, nw)));
    count_launch();
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}

int launch_generic_unpackdeq(const ct_quant_desc& d, const int32_t* packed, const void* scale, const void* zp,
                             const int32_t* g_idx, void* out, cudaStream_t stream) {
    if (d.rows * d.cols == 0) return CT_OK;
    GParams p = make_params(d, packed, scale, zp, g_idx, out);
    const int64_t nw = (d.cols * d.num_bits + 31) / 32;
    const int64_t items = d.rows * ((d.cols + 31) / 32);
    BITS_SWITCH(d.num_bits, (unpackdeq_generic_kernel<B><<<grid_for(items), 256, 0, stream>>>(p, nw)));
    count_launch();
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}

int launch_generic_pack(const int8_t* in, int32_t* out, int64_t rows, int64_t cols, int bits, int packed_dim, cudaStream_t stream) {
    if (rows * cols == 0) return CT_OK;
    if (packed_dim == 1) {
        const int64_t nw = (cols * bits + 31) / 32;
        const int64_t items = rows * ((cols + 31) / 32);
        BITS_SWITCH(bits, (pack_dim1_kernel<B><<<grid_for(items), 256, 0, stream>>>(in, out, rows, cols, nw)));
    } else {
        const int64_t nw = (rows * bits + 31) / 32;
        const int64_t items = ((rows + 31) / 32) * cols;
        BITS_SWITCH(bits, (pack_dim0_kernel<B><<<grid_for(items), 256, 0, stream>>>(in, out, rows, cols, nw)));
    }
    count_launch();
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}

int launch_generic_unpack(const int32_t* in, int8_t* out, int64_t rows, int64_t cols, int bits, int packed_dim, cudaStream_t stream) {
    if (rows * cols == 0) return CT_OK;
    if (packed_dim == 1) {
        const int64_t nw = (cols * bits + 31) / 32;
        const int64_t items = rows * ((cols + 31) / 32);
        BITS_SWITCH(bits, (unpack_dim1_kernel<B><<<grid_for(items), 256, 0, stream>>>(in, out, rows, cols, nw)));
    } else {
        const int64_t nw = (rows * bits + 31) / 32;
        const int64_t items = ((rows + 31) / 32) * cols;
        BITS_SWITCH(bits, (unpack_dim0_kernel<B><<<grid_for(items), 256, 0, stream>>>(in, out, rows, cols, nw)));
    }
    count_launch();
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}

}  // namespace ctb
