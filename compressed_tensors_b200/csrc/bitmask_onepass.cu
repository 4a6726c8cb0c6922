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
#include <cstdlib>

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
    __shared__ unsigned long long lb_sum[8];
    __shared__ int lb_p[8];
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
    // stage first (tile-local offsets only), then the whole CTA looks back: 256 descriptors per step
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
    if (threadIdx.x == 0) st_desc(desc + tile, (tile == 0 ? DESC_P : DESC_A) | (unsigned long long)total);
    unsigned long long excl = 0;
    if (tile > 0) {
        int64_t base = (int64_t)tile - 1;
        while (true) {
            const int64_t idx = base - threadIdx.x;
            unsigned long long d = DESC_P;                 // before the first tile: prefix 0
            if (idx >= 0) {
                do { d = ld_desc(desc + idx); } while ((d >> 62) == 0);
            }
            const uint32_t has_p = __ballot_sync(0xffffffffu, (d >> 62) == 2);
            const int stop = has_p ? (__ffs(has_p) - 1) : 31;
            unsigned long long part = (lane <= stop) ? (d & DESC_VAL) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) { lb_sum[warp] = part; lb_p[warp] = has_p != 0; }
            __syncthreads();
            bool found = false;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (!found) { excl += lb_sum[w]; found = lb_p[w] != 0; }
            }
            __syncthreads();
            if (found) break;
            base -= 256;
        }
    }
    if (threadIdx.x == 0) {
        st_desc(desc + tile, DESC_P | (excl + (unsigned long long)total));
        if (COMPRESS && tile == n_tiles - 1) *nnz_out = (int64_t)(excl + (unsigned long long)total);
    }
    __syncthreads();
    const unsigned long long pos = excl;
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

// ------------------------------------------------------------------------------------------------------------------------------
// v2 of the compressing direction: the same scan, restructured for latency.  v1 above keeps a tile in REGISTERS between its load and its
// write-out, so a CTA has nothing in flight while it waits for its predecessors (4 CTAs / SM at 64 registers: 213 us for 235 MB, slower
// than the two-phase kernels), and a first wave of ~1200 simultaneous CTAs resolves its chain 32 descriptors at a time (~40 us).
//   * a producer lane claims tiles from the ticket counter and streams them with 1-D bulk copies (TMA engine) into a shared-memory
//     ring; the input of the next tile is in flight whatever the consumers are waiting for
//   * 4 consumer warps read their units from shared memory (no 16-byte values held in registers), scan, compact into a staging
//     buffer, RELEASE the ring slot, and only then look back -- 128 descriptors per step (one per consumer thread)
//   * the compact run leaves the staging buffer as 16-byte vectors aligned to GLOBAL 16-byte boundaries: the run starts at an arbitrary
//     element, so every output vector is cut out of two aligned shared-memory vectors with a funnel shift (tile-uniform shift)
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int B2_CW = 4;                       // consumer warps
constexpr int B2_CT = 32 * B2_CW;              // consumer threads
constexpr int B2_TILE = 1024;                  // units per tile (8192 elements, 16 KB)
constexpr int B2_UPT = B2_TILE / B2_CT;        // units per consumer thread (8)
constexpr int B2_STAGES = 2;
constexpr uint32_t B2_HDR = 128;               // full[2] | empty[2] | ids[2] | pad
constexpr uint32_t B2_STAGE_BYTES = 8 * B2_TILE * 2 + 32;   // compact elements of a tile + one vector of slack for the funnel
constexpr uint32_t B2_SMEM = B2_HDR + B2_STAGES * (B2_TILE * 16) + B2_STAGE_BYTES;

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint32_t lds16(uint32_t saddr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
    return v;
}

__global__ void __launch_bounds__(32 * (B2_CW + 1)) bitmask_compress_ring_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ bitmask,
                                                                                  uint16_t* __restrict__ values, int64_t* __restrict__ row_offsets,
                                                                                  int64_t* __restrict__ nnz_out, unsigned long long* __restrict__ desc,
                                                                                  uint32_t* __restrict__ ticket, uint32_t n_units, uint32_t n_tiles, FastDiv upr) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int warp_tot[B2_UPT][B2_CW];
    __shared__ unsigned long long lb_sum[B2_CW];
    __shared__ int lb_p[B2_CW];
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t full0 = sbase, empty0 = sbase + 16, ids0 = sbase + 32;
    const uint32_t ring0 = sbase + B2_HDR;
    const uint32_t stage0 = ring0 + B2_STAGES * (B2_TILE * 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < B2_STAGES; ++s) {
            mbar_init_a(full0 + 8 * s, 1);
            mbar_init_a(empty0 + 8 * s, B2_CW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == B2_CW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t policy = l2_evict_first_policy();
            int s = 0;
            uint32_t ph = 0;
            while (true) {
                const uint32_t tile = atomicAdd(ticket, 1u);
                mbar_wait_a(empty0 + 8 * s, ph ^ 1u);
                if (tile >= n_tiles) {                                  // nothing left: tell the consumers
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ids0 + 4 * s), "r"(0xffffffffu) : "memory");
                    mbar_arrive_a(full0 + 8 * s);
                    break;
                }
                const uint32_t nu = min((uint32_t)B2_TILE, n_units - tile * B2_TILE);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(ids0 + 4 * s), "r"(tile) : "memory");
                mbar_expect_tx_a(full0 + 8 * s, nu * 16);
                bulk_g2s_a(ring0 + (uint32_t)s * (B2_TILE * 16), src + (size_t)tile * (B2_TILE * 16), nu * 16, full0 + 8 * s, policy);
                if (++s == B2_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int ctid = threadIdx.x;      // 0 .. 127
    int s = 0;
    uint32_t ph = 0;
    while (true) {
        mbar_wait_a(full0 + 8 * s, ph);
        uint32_t tile;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tile) : "r"(ids0 + 4 * s) : "memory");
        if (tile == 0xffffffffu) break;
        const uint32_t u0 = tile * B2_TILE;
        const uint32_t nu = min((uint32_t)B2_TILE, n_units - u0);
        const uint32_t ring = ring0 + (uint32_t)s * (B2_TILE * 16);

        // ---- A: mask bytes, counts, scan (unit order: segment k of 128 units, then thread) ----
        uint32_t bytes_lo = 0, bytes_hi = 0;      // mask bytes of units k = 0..3 / 4..7
        int incl[B2_UPT];
#pragma unroll
        for (int k = 0; k < B2_UPT; ++k) {
            const uint32_t i = k * B2_CT + ctid;
            uint32_t b = 0;
            if (i < nu) b = nz_byte16(lds128(ring + i * 16));
            if (k < 4) bytes_lo |= b << (8 * k);
            else bytes_hi |= b << (8 * (k - 4));
            incl[k] = __popc(b);
            // four neighbouring lanes combine their bytes into one aligned 32-bit store (n_units % 4 == 0)
            uint32_t w = b;
            w |= __shfl_down_sync(0xffffffffu, w, 1) << 8;
            w |= __shfl_down_sync(0xffffffffu, w, 2) << 16;
            if ((ctid & 3) == 0 && i < nu) reinterpret_cast<uint32_t*>(bitmask)[(u0 + i) >> 2] = w;
        }
        int cnt_packed_lo = 0, cnt_packed_hi = 0;   // the per-unit counts (<= 8: 4 bits each)
#pragma unroll
        for (int k = 0; k < B2_UPT; ++k) {
            if (k < 4) cnt_packed_lo |= incl[k] << (4 * k);
            else cnt_packed_hi |= incl[k] << (4 * (k - 4));
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < B2_UPT; ++k) {
                const int n = __shfl_up_sync(0xffffffffu, incl[k], o);
                if (lane >= o) incl[k] += n;
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < B2_UPT; ++k) warp_tot[k][warp] = incl[k];
        }
        named_bar_sync(1, B2_CT);
        int total = 0;
        int off[B2_UPT];
#pragma unroll
        for (int k = 0; k < B2_UPT; ++k) {
            int before = 0, seg = 0;
#pragma unroll
            for (int w = 0; w < B2_CW; ++w) {
                const int t = warp_tot[k][w];
                if (w < warp) before += t;
                seg += t;
            }
            const int c = ((k < 4 ? cnt_packed_lo >> (4 * k) : cnt_packed_hi >> (4 * (k - 4))) & 15);
            off[k] = total + before + incl[k] - c;
            total += seg;
        }

        // ---- B: compaction into the staging buffer (element j of the tile's run at stage[j]) ----
#pragma unroll
        for (int k = 0; k < B2_UPT; ++k) {
            const uint32_t i = k * B2_CT + ctid;
            const uint32_t b = (k < 4 ? bytes_lo >> (8 * k) : bytes_hi >> (8 * (k - 4))) & 0xffu;
            if (b) {
                const uint4 v = lds128(ring + i * 16);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t o = stage0 + 2u * (uint32_t)off[k];
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if ((b >> e) & 1u) { sts16(o, w[e >> 1] >> (16 * (e & 1))); o += 2; }
            }
        }
        named_bar_sync(1, B2_CT);
        if (lane == 0) mbar_arrive_a(empty0 + 8 * s);      // the producer may refill this slot while we look back

        // ---- C: exclusive prefix of the tile, 128 descriptors per step ----
        if (ctid == 0) st_desc(desc + tile, (tile == 0 ? DESC_P : DESC_A) | (unsigned long long)total);
        unsigned long long excl = 0;
        if (tile > 0) {
            int64_t base = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = base - ctid;
                unsigned long long d = DESC_P;                 // before the first tile: prefix 0
                if (idx >= 0) {
                    do { d = ld_desc(desc + idx); } while ((d >> 62) == 0);
                }
                const uint32_t has_p = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                const int stop = has_p ? (__ffs(has_p) - 1) : 31;
                unsigned long long part = (lane <= stop) ? (d & DESC_VAL) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lane == 0) { lb_sum[warp] = part; lb_p[warp] = has_p != 0; }
                named_bar_sync(1, B2_CT);
                bool found = false;
#pragma unroll
                for (int w = 0; w < B2_CW; ++w) {
                    if (!found) { excl += lb_sum[w]; found = lb_p[w] != 0; }
                }
                named_bar_sync(1, B2_CT);
                if (found) break;
                base -= B2_CT;
            }
        }
        if (ctid == 0) {
            st_desc(desc + tile, DESC_P | (excl + (unsigned long long)total));
            if (tile == n_tiles - 1) *nnz_out = (int64_t)(excl + (unsigned long long)total);
        }

        // ---- D: row offsets, then the compact run [excl, excl + total) as global-aligned 16-byte vectors ----
        if (row_offsets) {
#pragma unroll
            for (int k = 0; k < B2_UPT; ++k) {
                const uint32_t i = k * B2_CT + ctid;
                if (i < nu) {
                    const uint32_t gu = u0 + i, r = fd_div(gu, upr);
                    if (r * upr.d == gu) row_offsets[r] = (int64_t)(excl + (unsigned long long)off[k]);
                }
            }
        }
        {
            const uint32_t shift = (uint32_t)(excl & 7ull);               // elements of the first global vector that belong to earlier tiles
            const unsigned long long g0 = excl - shift;                    // first element of the first global vector we touch
            const uint32_t span = shift + (uint32_t)total;                 // elements from g0 to the end of the run
            const uint32_t nvec = (span + 7) >> 3;
            uint4* gv = reinterpret_cast<uint4*>(values + g0);
            // vector c holds stage elements [8c - shift, 8c - shift + 8): halves (8 - shift) .. of the pair (stage vector c - 1, stage vector c)
            const uint32_t h0 = 8u - shift;                                // 1 .. 8
            const uint32_t wo = h0 >> 1, odd = h0 & 1u;
            for (uint32_t c = ctid; c < nvec; c += B2_CT) {
                const bool first = (c == 0), last = (8 * c + 8 > span);
                uint4 A = make_uint4(0, 0, 0, 0);
                if (c > 0 || shift == 0) A = lds128(stage0 + 16u * (c - (shift ? 1u : 0u)));   // shift == 0: the vector is stage vector c itself
                uint4 B = A;
                if (shift) B = lds128(stage0 + 16u * c);
                uint32_t o[4];
                if (shift == 0) { o[0] = A.x; o[1] = A.y; o[2] = A.z; o[3] = A.w; }
                else {
                    const uint32_t W[9] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, 0u};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t lo = 0, hi = 0;
                        switch (wo) {                                       // tile-uniform
                        case 0: lo = W[i]; hi = W[i + 1]; break;
                        case 1: lo = W[i + 1]; hi = W[i + 2]; break;
                        case 2: lo = W[i + 2]; hi = W[i + 3]; break;
                        case 3: lo = W[i + 3]; hi = W[i + 4]; break;
                        default: lo = W[i + 4]; hi = W[i + 5 > 8 ? 8 : i + 5]; break;
                        }
                        o[i] = odd ? __funnelshift_r(lo, hi, 16) : lo;
                    }
                }
                if (!first && !last) gv[c] = make_uint4(o[0], o[1], o[2], o[3]);
                else {
                    // partial vector: elements before the run belong to another tile, elements after it to the next one
                    uint16_t* ge = values + g0 + 8ull * c;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t pos = 8 * c + e;
                        if (pos >= shift && pos < span) ge[e] = (uint16_t)(o[e >> 1] >> (16 * (e & 1)));
                    }
                }
            }
        }
        named_bar_sync(1, B2_CT);      // the staging buffer is reused by the next tile
        if (++s == B2_STAGES) { s = 0; ph ^= 1u; }
    }
}

bool bitmask_lookback_ok(int dtype, int64_t rows, int64_t cols, const void* dense, const void* mask, const void* values) {
    const int64_t n_units = rows * cols / 8;
    return dt_size(dtype) == 2 && cols % 8 == 0 && n_units % 4 == 0 && n_units > 0 && n_units < 0x7fffffffLL && aligned16(dense) &&
           (reinterpret_cast<uintptr_t>(mask) & 3u) == 0 && aligned16(values);
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
    if (COMPRESS && !getenv("CT_B200_BITMASK_V1")) {
        auto kfn = bitmask_compress_ring_kernel;
        CT_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B2_SMEM));
        uint32_t grid = (uint32_t)sm_count(device) * 4;
        if (grid > n_tiles) grid = n_tiles;
        kfn<<<grid, 32 * (B2_CW + 1), B2_SMEM, st>>>(reinterpret_cast<const uint8_t*>(src), bitmask, reinterpret_cast<uint16_t*>(dst), row_offsets, nnz_out,
                                                     reinterpret_cast<unsigned long long*>(scratch + 16), reinterpret_cast<uint32_t*>(scratch),
                                                     (uint32_t)n_units, n_tiles, make_fastdiv((uint64_t)(cols / 8)));
    } else {
        bitmask_lookback_kernel<COMPRESS><<<n_tiles, 256, 0, st>>>(src, bitmask, dst, row_offsets, nnz_out,
                                                                   reinterpret_cast<unsigned long long*>(scratch + 16), reinterpret_cast<uint32_t*>(scratch),
                                                                   (uint32_t)n_units, n_tiles, make_fastdiv((uint64_t)(cols / 8)));
    }
    count_launch();
    cudaFreeAsync(scratch, st);
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}
template int launch_bitmask_lookback<true>(const void*, uint8_t*, void*, int64_t*, int64_t*, int64_t, int64_t, int, cudaStream_t);
template int launch_bitmask_lookback<false>(const void*, uint8_t*, void*, int64_t*, int64_t*, int64_t, int64_t, int, cudaStream_t);

}  // namespace ctb
