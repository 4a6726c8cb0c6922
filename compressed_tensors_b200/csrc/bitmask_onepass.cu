// bitmask_onepass.cu -- unstructured bitmask compression / expansion in ONE pass over the dense tensor.
//
//   format (CompressionFormat.sparse_bitmask, config/base.py:17; restated in oracle/ct_oracle.c, "parity unpinned"):
//       values = x[x != 0] row-major, bitmask = pack_bitmasks(x != 0) (utils/helpers.py:306-317), row_offsets = exclusive prefix of the
//       per-row counts, nnz = values.numel()
//
// Stream compaction needs a device-wide exclusive scan of the per-tile non-zero counts.  The two-phase kernels of sparse.cu
// (count -> cub scan -> move) read the dense tensor twice and need the host to learn nnz before `values` can be allocated.  Here the
// scan is a DECOUPLED LOOK-BACK over one 64-bit descriptor per tile ({status, count} in a single word, so publishing it needs no
// fence): a tile publishes its own count as soon as it has it, then sums its predecessors' descriptors backwards until it meets one
// that already holds an inclusive prefix.  Tiles are claimed from a ticket counter, so every predecessor of a running tile has
// itself been claimed by a resident CTA (forward progress).  The dense tensor is read exactly once:
//       compress  : 2 B in + 2 B x density out + 1/8 B mask per element  (3.125 B at 50 % zeros)
//       expand    : 1/8 B mask + 2 B x density in, 2 B out
// `values` is written into a caller buffer of capacity rows * cols; nnz stays on the device.
//
// Work decomposition (2-byte dtypes, cols % 8 == 0): unit = 8 elements = one 16-byte access and one mask byte; tile = 256 threads x
// BM_U units, thread t owns units t + 256 u (coalesced), scan order = unit order.  The kept elements of a tile are staged in shared
// memory so that the compact side is moved coalesced.
#include "engine.h"

namespace ctb {

constexpr int BM_U = 4;
constexpr int BM_TILE = 256 * BM_U;   // units per tile (8192 elements, 16 KB of bf16)

constexpr unsigned long long DESC_A = 1ull << 62;   // value = this tile's count
constexpr unsigned long long DESC_P = 2ull << 62;   // value = inclusive prefix up to and including this tile
constexpr unsigned long long DESC_VAL = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t nz_byte16(const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        byte |= ((w[j] & 0x7fffu) != 0u ? 1u : 0u) << (2 * j);            // -0.0 == 0
        byte |= ((w[j] & 0x7fff0000u) != 0u ? 1u : 0u) << (2 * j + 1);
    }
    return byte;
}

// exclusive prefix of this tile = sum of the counts of tiles [0, tile); executed by warp 0.  `total` = this tile's count.
__device__ __forceinline__ unsigned long long lookback(unsigned long long* desc, uint32_t tile, uint32_t total, int lane) {
    if (lane == 0) st_desc(desc + tile, (tile == 0 ? DESC_P : DESC_A) | (unsigned long long)total);
    unsigned long long excl = 0;
    if (tile == 0) return 0;
    int64_t look = (int64_t)tile - 1;          // lane l inspects tile look - l
    while (true) {
        const int64_t idx = look - lane;
        unsigned long long d;
        do {
            d = (idx >= 0) ? ld_desc(desc + idx) : DESC_P;          // before the first tile: prefix 0
        } while (__any_sync(0xffffffffu, (d >> 62) == 0));          // someone has not published yet: look again
        const uint32_t has_p = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        // lanes up to (and including) the nearest tile that holds a prefix contribute
        const int stop = has_p ? (__ffs(has_p) - 1) : 31;
        unsigned long long part = (lane <= stop) ? (d & DESC_VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        excl += part;
        if (has_p) break;
        look -= 32;
    }
    if (lane == 0) st_desc(desc + tile, DESC_P | (excl + total));
    return excl;
}

// COMPRESS: src = dense [n_units x 16 B], dst = values (compact), bitmask written, row_offsets / nnz written
// !COMPRESS: src = values (compact), bitmask read, dst = dense
template <bool COMPRESS>
__global__ void __launch_bounds__(256) bitmask_lookback_kernel(const void* __restrict__ src, uint8_t* __restrict__ bitmask, void* __restrict__ dst,
                                                               int64_t* __restrict__ row_offsets, int64_t* __restrict__ nnz_out,
                                                               unsigned long long* __restrict__ desc, uint32_t* __restrict__ ticket,
                                                               uint32_t n_units, uint32_t n_tiles, FastDiv upr /* units per row */) {
    __shared__ __align__(16) uint16_t stage[8 * BM_TILE];
    __shared__ int warp_tot[BM_U][8];
    __shared__ uint32_t tile_s;
    __shared__ unsigned long long prefix_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = tile_s;
    if (tile >= n_tiles) return;
    const uint32_t u0 = tile * BM_TILE;

    uint4 v[BM_U];
    uint32_t byte[BM_U];
    int cnt[BM_U], incl[BM_U], off[BM_U];
#pragma unroll
    for (int u = 0; u < BM_U; ++u) {
        const uint32_t i = u0 + u * 256 + threadIdx.x;
        v[u] = make_uint4(0, 0, 0, 0);
        byte[u] = 0;
        if (i < n_units) {
            if (COMPRESS) v[u] = ldg_stream16(reinterpret_cast<const uint4*>(src) + i);
            else byte[u] = __ldg(bitmask + i);
        }
    }
#pragma unroll
    for (int u = 0; u < BM_U; ++u) {
        if (COMPRESS) byte[u] = nz_byte16(v[u]);
        cnt[u] = __popc(byte[u]);
        incl[u] = cnt[u];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int u = 0; u < BM_U; ++u) {
            const int n = __shfl_up_sync(0xffffffffu, incl[u], o);
            if (lane >= o) incl[u] += n;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int u = 0; u < BM_U; ++u) warp_tot[u][warp] = incl[u];
    }
    if (COMPRESS) {
        // mask bytes: four neighbouring lanes combine theirs into one aligned 32-bit store (n_units % 4 == 0)
#pragma unroll
        for (int u = 0; u < BM_U; ++u) {
            const uint32_t i = u0 + u * 256 + threadIdx.x;
            uint32_t w = byte[u];
            w |= __shfl_down_sync(0xffffffffu, w, 1) << 8;
            w |= __shfl_down_sync(0xffffffffu, w, 2) << 16;
            if ((threadIdx.x & 3) == 0 && i < n_units) reinterpret_cast<uint32_t*>(bitmask)[i >> 2] = w;
        }
    }
    __syncthreads();
    int total = 0;
#pragma unroll
    for (int u = 0; u < BM_U; ++u) {
        int before = 0, seg = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int t = warp_tot[u][w];
            if (w < warp) before += t;
            seg += t;
        }
        off[u] = total + before + incl[u] - cnt[u];      // this unit's first slot in the tile's compact run
        total += seg;
    }
    // warp 0 resolves the tile's global prefix while the other warps stage their elements
    if (warp == 0) {
        const unsigned long long excl = lookback(desc, tile, (uint32_t)total, lane);
        if (lane == 0) {
            prefix_s = excl;
            if (COMPRESS && tile == n_tiles - 1) *nnz_out = (int64_t)(excl + (unsigned long long)total);
        }
    }
    if (COMPRESS) {
#pragma unroll
        for (int u = 0; u < BM_U; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            int o = off[u];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((byte[u] >> k) & 1u) stage[o++] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
        }
    }
    __syncthreads();
    const unsigned long long pos = prefix_s;
    uint16_t* cout = reinterpret_cast<uint16_t*>(dst);
    const uint16_t* cin = reinterpret_cast<const uint16_t*>(src);
    // compact run [pos, pos + total): one 2-byte element to reach 4-byte alignment, then pairs, then a possible last element
    const int head = (int)(pos & 1ull) & (total > 0 ? 1 : 0);     // values is at least 4-byte aligned: parity of the element index
    const int pairs = (total - head) >> 1;
    const int tail = total - head - 2 * pairs;
    if (COMPRESS) {
        if (row_offsets) {
#pragma unroll
            for (int u = 0; u < BM_U; ++u) {
                const uint32_t i = u0 + u * 256 + threadIdx.x;
                if (i < n_units) {
                    const uint32_t r = fd_div(i, upr);
                    if (r * upr.d == i) row_offsets[r] = (int64_t)(pos + (unsigned long long)off[u]);
                }
            }
        }
        if (threadIdx.x == 0 && head) cout[pos] = stage[0];
        uint32_t* c32 = reinterpret_cast<uint32_t*>(cout + pos + head);
        for (int j = threadIdx.x; j < pairs; j += 256) c32[j] = (uint32_t)stage[head + 2 * j] | ((uint32_t)stage[head + 2 * j + 1] << 16);
        if (threadIdx.x == 32 && tail) cout[pos + total - 1] = stage[total - 1];
    } else {
        if (threadIdx.x == 0 && head) stage[0] = cin[pos];
        const uint32_t* c32 = reinterpret_cast<const uint32_t*>(cin + pos + head);
        for (int j = threadIdx.x; j < pairs; j += 256) {
            const uint32_t two = ldg_stream4(c32 + j);
            stage[head + 2 * j] = (uint16_t)two;
            stage[head + 2 * j + 1] = (uint16_t)(two >> 16);
        }
        if (threadIdx.x == 32 && tail) stage[total - 1] = cin[pos + total - 1];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < BM_U; ++u) {
            const uint32_t i = u0 + u * 256 + threadIdx.x;
            if (i < n_units) {
                uint32_t e[8];
                int o = off[u];
#pragma unroll
                for (int k = 0; k < 8; ++k) e[k] = ((byte[u] >> k) & 1u) ? (uint32_t)stage[o++] : 0u;
                stg_stream16(reinterpret_cast<uint4*>(dst) + i, make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16)));
            }
        }
    }
}

bool bitmask_lookback_ok(int dtype, int64_t rows, int64_t cols, const void* dense, const void* mask, const void* values) {
    const int64_t n_units = rows * cols / 8;
    return dt_size(dtype) == 2 && cols % 8 == 0 && n_units % 4 == 0 && n_units > 0 && n_units < 0x7fffffffLL && aligned16(dense) &&
           (reinterpret_cast<uintptr_t>(mask) & 3u) == 0 && (reinterpret_cast<uintptr_t>(values) & 3u) == 0;
}

// scratch of one launch: [16 B ticket][n_tiles descriptors], zeroed; stream-ordered like the job tables of dispatch.cu
template <bool COMPRESS>
int launch_bitmask_lookback(const void* src, uint8_t* bitmask, void* dst, int64_t* row_offsets, int64_t* nnz_out, int64_t rows, int64_t cols,
                            int device, cudaStream_t st) {
    const int64_t n_units = rows * cols / 8;
    const uint32_t n_tiles = (uint32_t)((n_units + BM_TILE - 1) / BM_TILE);
    uint8_t* scratch = nullptr;
    const size_t bytes = 16 + (size_t)n_tiles * sizeof(unsigned long long);
    int rc = scratch_alloc(reinterpret_cast<void**>(&scratch), bytes, device, st);
    if (rc) return rc;
    CT_CUDA_TRY(cudaMemsetAsync(scratch, 0, bytes, st));
    bitmask_lookback_kernel<COMPRESS><<<n_tiles, 256, 0, st>>>(src, bitmask, dst, row_offsets, nnz_out,
                                                               reinterpret_cast<unsigned long long*>(scratch + 16), reinterpret_cast<uint32_t*>(scratch),
                                                               (uint32_t)n_units, n_tiles, make_fastdiv((uint64_t)(cols / 8)));
    count_launch();
    cudaFreeAsync(scratch, st);
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}
template int launch_bitmask_lookback<true>(const void*, uint8_t*, void*, int64_t*, int64_t*, int64_t, int64_t, int, cudaStream_t);
template int launch_bitmask_lookback<false>(const void*, uint8_t*, void*, int64_t*, int64_t*, int64_t, int64_t, int, cudaStream_t);

}  // namespace ctb
