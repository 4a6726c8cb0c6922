// stream.cuh -- the two persistent streaming pipelines every hot-path op is instantiated on.
//
// Work decomposition (all hot-path ops are elementwise over a flat, contiguous tensor):
//   chunk = 8 consecutive elements
//   unit  = Op::GROUP consecutive chunks = one thread's piece of work.  GROUP is chosen per op so
//           that the thread's OUTPUT is one 16-byte store (4-bit packing: 32 elements -> 4 words;
//           1-byte codes: 16 elements) and the scale / reciprocal is fetched once per unit.
//   tile  = TILE_CHUNKS chunks = 8192 elements (16 KB of bf16) = one CTA's unit of work
//   job   = one tensor; a launch covers a table of jobs (whole-model batches); tiles are numbered
//           globally and CLAIMED from a device counter by a persistent grid of n_SM * ctas_per_sm CTAs
//           (dynamic schedule, see stream_tma_kernel; small launches keep a static round-robin deal).
//
// Pipeline 1 ("tma"): warp-specialised.  One producer lane issues 1-D bulk async copies
// (cp.async.bulk, the TMA engine without a tensor map) of whole input tiles into a ring of
// shared-memory stages, each guarded by a full/empty mbarrier pair (waits use the hardware
// suspend hint, so the idle producer does not burn issue slots).  NCW consumer warps wait on the
// full barrier, read their units from shared memory with bank-conflict-free swizzled LDS.128, do
// the arithmetic in registers and write with coalesced 16-byte streaming stores.  The scale loads
// of a tile are issued BEFORE the barrier wait and consumed after it.
// Pipeline 0 ("direct"): every thread issues its tile's 128-bit global loads up front and relies on
// occupancy for memory-level parallelism (kept as the comparison point and for profiling).
#pragma once

#include "common.cuh"

namespace ctb {

#ifndef CT_TILE_CHUNKS
#define CT_TILE_CHUNKS 1024
#endif
constexpr int TILE_CHUNKS = CT_TILE_CHUNKS;   // default tile: 1024 chunks = 8192 elements (16 KB of bf16); build-time knob for A/B variants
#ifndef CT_BIG_TILE_CHUNKS
#define CT_BIG_TILE_CHUNKS 2048
#endif
// An op may ask for another tile size (`static constexpr int TILE`).  The four headline 16-bit ops run on 2048-chunk tiles with a
// 3-stage x 2-CTA ring: half the per-tile bookkeeping of the consumer warps for the same bytes in flight; +2.2 % / +1.4 % / +2.5 % / +2.5 %
// on quantize+pack / unpack+dequantize / fp8 quantize / fp8 dequantize (profiles/ops_r2.md).  The host side must count tiles with the
// same number: sig_tile_chunks() in engine.h, checked at every launch.
constexpr int BIG_TILE_CHUNKS = CT_BIG_TILE_CHUNKS;
constexpr int CHUNK_ELEMS = 8;
constexpr int NCW = 8;              // consumer warps (tma pipeline)
constexpr int DIRECT_THREADS = 256;
constexpr int MAX_STAGES = 12;

struct Job {
    const uint8_t* in;     // streamed input
    const void* scale;
    const void* zp;
    uint8_t* out;
    uint32_t n_chunks;     // numel / 8
    uint32_t tile_begin;   // global id of this job's first tile
    uint32_t tile_end;
    uint32_t _pad;
    FastDiv dc;            // flat mode: chunks per scale element (scale index = chunk / dc); 2-D mode: chunks per column block
    // 2-D mode (BLOCK strategy, one-row group scales): sidx = (row / rd) * srs + (chunk_in_row / dc)
    FastDiv cpr;           // chunks per row (d == 0: flat mode)
    FastDiv rd;            // rows per scale row (block height; "infinite" for a one-row scale)
    uint32_t srs;          // scale row stride
    uint32_t _pad2;
    const void* aux;       // FP4 ops: the tensor's global scale (one float32) or nullptr
};

struct JobTable {
    const Job* jobs;       // device array when n > 1
    int n;
    Job one;               // inline copy when n == 1 (no table upload needed)
};

// launch-uniform constants
struct Common {
    uint32_t qmin2, qmax2;   // clamp bounds duplicated in both halves, in the compute dtype (16-bit T)
    float qmin, qmax;        // same as fp32
    int bits;
};

__device__ __forceinline__ uint32_t job_tile_end(const JobTable& t, int j) { return t.n == 1 ? t.one.tile_end : t.jobs[j].tile_end; }
// by-value copy: the job lives in registers until a tile of another tensor comes up
__device__ __forceinline__ Job job_at(const JobTable& t, int j) {
    if (t.n == 1) return t.one;
    return t.jobs[j];
}

// ------------------------------------------------------------------------------------
// Op concept:
//   static constexpr int IN_BYTES;     streamed input bytes per chunk (4, 8, 16 or 32)
//   static constexpr int GROUP;        chunks per unit (1, 2 or 4)
//   struct Raw;                        raw scale / zero-point bits of a unit (plain loads only)
//   static Raw prefetch(J, gc0);       issued BEFORE the data arrives; no dependent arithmetic
//   static void run(J, cm, raw, gc0, w[GROUP][IN_BYTES/4], off);   gc0 = first chunk of the unit;
//                                      w[k] holds chunk (k + off) mod GROUP (see swz_off)
//   optional:  struct Tile;  static Tile tile(J);   evaluated once per (thread, tile) and passed to run() as a 7th argument:
//                                      per-tensor constants that are too expensive to rebuild per unit
// ------------------------------------------------------------------------------------

template <class Op, class = void> struct HasTile { static constexpr bool value = false; };
template <class Op> struct HasTile<Op, decltype((void)sizeof(typename Op::Tile))> { static constexpr bool value = true; };
template <class Op, class = void> struct HasShape { static constexpr bool value = false; };
template <class Op> struct HasShape<Op, decltype((void)Op::PREF_CTAS)> { static constexpr bool value = true; };
template <class Op, class = void> struct TileOf { static constexpr int value = TILE_CHUNKS; };
template <class Op> struct TileOf<Op, decltype((void)Op::TILE)> { static constexpr int value = Op::TILE; };
struct NoTile {};
template <class Op> __device__ __forceinline__ auto op_tile(const Job& J) {
    if constexpr (HasTile<Op>::value) return Op::tile(J);
    else return NoTile{};
}
template <class Op, class W, class T>
__device__ __forceinline__ void op_run(const Job& J, const Common& cm, const typename Op::Raw& r, uint32_t gc0, const W& w, int off, const T& tc) {
    if constexpr (HasTile<Op>::value) Op::run(J, cm, r, gc0, w, off, tc);
    else Op::run(J, cm, r, gc0, w, off);
}

template <int BYTES>
__device__ __forceinline__ void lds_chunk(uint32_t saddr, uint32_t (&w)[BYTES / 4]) {
    if constexpr (BYTES == 4) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(saddr));
    } else if constexpr (BYTES == 8) {
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(saddr));
    } else if constexpr (BYTES == 16) {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr));
    } else {
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(saddr + 16));
    }
}
// CACHED = true keeps the line in L1 (several loads of one thread share 128-byte lines)
template <int BYTES, bool CACHED>
__device__ __forceinline__ void ldg_chunk(const uint8_t* p, uint32_t (&w)[BYTES / 4]) {
    if constexpr (BYTES == 4) {
        w[0] = CACHED ? __ldg(reinterpret_cast<const uint32_t*>(p)) : ldg_stream4(p);
    } else if constexpr (BYTES == 8) {
        uint2 v = CACHED ? __ldg(reinterpret_cast<const uint2*>(p)) : ldg_stream8(p);
        w[0] = v.x; w[1] = v.y;
    } else if constexpr (BYTES == 16) {
        uint4 v = CACHED ? __ldg(reinterpret_cast<const uint4*>(p)) : ldg_stream16(p);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        uint4 v = CACHED ? __ldg(reinterpret_cast<const uint4*>(p)) : ldg_stream16(p);
        uint4 u = CACHED ? __ldg(reinterpret_cast<const uint4*>(p + 16)) : ldg_stream16(p + 16);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = u.x; w[5] = u.y; w[6] = u.z; w[7] = u.w;
    }
}

// Bank-conflict-free reads of a unit from shared memory: lane l reads chunk (k + swz_off(l)) mod GROUP
// in its k-th LDS.128, so the 8 lanes of a quarter warp always cover 8 different 16-byte bank groups
// (unit stride = GROUP * 16 bytes).  Register k of the thread therefore holds chunk (k + off) mod GROUP;
// the op rotates its per-chunk outputs back before the store (rotate_out below).
template <int GROUP>
__device__ __forceinline__ int swz_off(int lane) {
    if constexpr (GROUP == 4) return (lane >> 1) & 3;
    else if constexpr (GROUP == 2) return (lane >> 2) & 1;
    else return 0;
}
// o holds GROUP blocks of WPC words, block k belonging to chunk (k + off) mod GROUP: put them in chunk order
template <int GROUP, int WPC>
__device__ __forceinline__ void rotate_out(uint32_t (&o)[GROUP * WPC], int off) {
    if constexpr (GROUP == 2) {
        const bool sw = off & 1;
#pragma unroll
        for (int i = 0; i < WPC; ++i) {
            const uint32_t a = o[i], b = o[WPC + i];
            o[i] = sw ? b : a;
            o[WPC + i] = sw ? a : b;
        }
    } else if constexpr (GROUP == 4) {
        const bool r1 = off & 1, r2 = off & 2;
#pragma unroll
        for (int i = 0; i < WPC; ++i) {
            uint32_t a0 = o[i], a1 = o[WPC + i], a2 = o[2 * WPC + i], a3 = o[3 * WPC + i];
            // block k -> position k+1
            uint32_t b0 = r1 ? a3 : a0, b1 = r1 ? a0 : a1, b2 = r1 ? a1 : a2, b3 = r1 ? a2 : a3;
            // block k -> position k+2
            o[i] = r2 ? b2 : b0;
            o[WPC + i] = r2 ? b3 : b1;
            o[2 * WPC + i] = r2 ? b0 : b2;
            o[3 * WPC + i] = r2 ? b1 : b3;
        }
    }
}

// ------------------------------------------------------------------------------------
// pipeline 1: TMA bulk-copy ring
// ------------------------------------------------------------------------------------
// Tile schedule.  `sched` != nullptr: DYNAMIC -- the producer lane of every CTA claims tile ids from a
// zero-initialised device counter with atomicAdd, one claim in flight ahead of its use.  A static
// round-robin deal leaves bandwidth on the table on B200: SMs do not all see the same HBM bandwidth
// (two dies, eight stacks), so equal shares finish at different times and the launch waits for the
// slowest SM (measured with tools/ubench/mix_ceiling.cu: 6.4-6.8 TB/s static vs 7.2 TB/s dynamic
// for a 2:1 / 4:1 read:write stream).  `sched` == nullptr: static deal (small launches, <= 1 tile a CTA).
// Either way the producer publishes {tile, next tile} of a stage in shared memory before it arms the
// stage's full barrier; consumers pick both up after the wait (release/acquire through the mbarrier).
constexpr uint32_t SMEM_HDR = 384;   // full[12] | empty[12] | ids[12] (8 bytes each), then the data ring

__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void lds_v2(uint32_t saddr, uint32_t& a, uint32_t& b) {
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(saddr) : "memory");
}

template <class Op>
__global__ void __launch_bounds__(32 * (NCW + 1)) stream_tma_kernel(const __grid_constant__ JobTable tbl,
                                                                   const __grid_constant__ Common cm,
                                                                   uint32_t total_tiles, int stages, uint32_t* sched) {
    constexpr int TILE_CHUNKS = TileOf<Op>::value;     // shadows the default on purpose: everything below is per-op
    constexpr int TILE_BYTES = TILE_CHUNKS * Op::IN_BYTES;
    constexpr int CTHREADS = NCW * 32;
    constexpr int G = Op::GROUP;
    constexpr int UNITS = TILE_CHUNKS / G;
    constexpr int ITERS = UNITS / CTHREADS;
    static_assert(UNITS % CTHREADS == 0, "tile must split evenly over the consumer threads");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t full0 = sbase;                        // full[s]  at sbase + 8 s
    const uint32_t empty0 = sbase + 8 * MAX_STAGES;      // empty[s] at sbase + 96 + 8 s
    const uint32_t ids0 = sbase + 16 * MAX_STAGES;       // ids[s]   at sbase + 192 + 8 s
    const uint32_t data0 = sbase + SMEM_HDR;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init_a(full0 + 8 * s, 1);
            mbar_init_a(empty0 + 8 * s, NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t policy = l2_evict_first_policy();
            int s = 0, j = 0;
            uint32_t ph = 0;
            // cur / nxt are known, nn is the claim in flight
            uint32_t cur, nxt, nn;
            if (sched) {
                cur = atomicAdd(sched, 2u);
                nxt = cur + 1;                    // the first claim takes two tiles: both ids are needed at once
            } else {
                cur = blockIdx.x;
                nxt = cur + gridDim.x;
            }
            if (cur >= total_tiles) {             // nothing left for this CTA: tell the consumers
                sts_v2(ids0, cur, cur);
                mbar_arrive_a(full0);
            } else {
                Job J = job_at(tbl, 0);
                while (cur < total_tiles) {
                    if (sched) nn = (nxt < total_tiles) ? atomicAdd(sched, 1u) : nxt;
                    else nn = nxt + gridDim.x;
                    if (cur >= J.tile_end) {
                        while (cur >= job_tile_end(tbl, j)) ++j;
                        J = job_at(tbl, j);
                    }
                    const uint32_t lt = cur - J.tile_begin;
                    const uint32_t base = lt * TILE_CHUNKS;
                    const uint32_t chunks = min((uint32_t)TILE_CHUNKS, J.n_chunks - base);
                    const uint32_t bytes = chunks * Op::IN_BYTES;
                    mbar_wait_a(empty0 + 8 * s, ph ^ 1u);
                    sts_v2(ids0 + 8 * s, cur, nxt);
                    mbar_expect_tx_a(full0 + 8 * s, bytes);
                    bulk_g2s_a(data0 + (uint32_t)s * TILE_BYTES, J.in + (size_t)lt * TILE_BYTES, bytes, full0 + 8 * s, policy);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                    cur = nxt;
                    nxt = nn;
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int ctid = threadIdx.x;  // 0 .. CTHREADS-1
        int s = 0, j = 0;
        uint32_t ph = 0;
        const int off = (Op::IN_BYTES == 16) ? swz_off<G>(lane) : 0;
        // The scale / zero-point loads run ONE TILE AHEAD of the arithmetic: they miss L2 (each is
        // used once) and would otherwise expose a full DRAM round trip per tile.
        Job Jn = job_at(tbl, 0);
        typename Op::Raw raw_next[ITERS];
        uint32_t tile, nxt;
        mbar_wait_a(full0, 0);
        lds_v2(ids0, tile, nxt);
        if (tile < total_tiles) {
            while (tile >= job_tile_end(tbl, j)) ++j;
            Jn = job_at(tbl, j);
            const uint32_t base = (tile - Jn.tile_begin) * TILE_CHUNKS;
            const uint32_t chunks = min((uint32_t)TILE_CHUNKS, Jn.n_chunks - base);
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c0 = (uint32_t)(it * CTHREADS + ctid) * G;
                if (c0 < chunks) raw_next[it] = Op::prefetch(Jn, base + c0);
            }
        }
        // per-tensor constants of the op (Op::tile): rebuilt only when the next tile belongs to another tensor
        decltype(op_tile<Op>(Jn)) tc{};
        bool new_job = true;
        while (tile < total_tiles) {
            // invariant: full[s] of `tile` has been waited for, `nxt` is the tile after it
            const Job J = Jn;
            const uint32_t base = (tile - J.tile_begin) * TILE_CHUNKS;
            const uint32_t chunks = min((uint32_t)TILE_CHUNKS, J.n_chunks - base);
            if (new_job) { tc = op_tile<Op>(J); new_job = false; }
            typename Op::Raw raw[ITERS];
#pragma unroll
            for (int it = 0; it < ITERS; ++it) raw[it] = raw_next[it];
            if (nxt < total_tiles) {
                if (nxt >= Jn.tile_end) {
                    while (nxt >= job_tile_end(tbl, j)) ++j;
                    Jn = job_at(tbl, j);
                    new_job = true;
                }
                const uint32_t nbase = (nxt - Jn.tile_begin) * TILE_CHUNKS;
                const uint32_t nchunks = min((uint32_t)TILE_CHUNKS, Jn.n_chunks - nbase);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const uint32_t c0 = (uint32_t)(it * CTHREADS + ctid) * G;
                    if (c0 < nchunks) raw_next[it] = Op::prefetch(Jn, nbase + c0);
                }
            }
            const uint32_t sp = data0 + (uint32_t)s * TILE_BYTES;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c0 = (uint32_t)(it * CTHREADS + ctid) * G;
                if (c0 < chunks) {
                    uint32_t w[G][Op::IN_BYTES / 4];
#pragma unroll
                    for (int k = 0; k < G; ++k) lds_chunk<Op::IN_BYTES>(sp + (c0 + ((k + off) & (G - 1))) * Op::IN_BYTES, w[k]);
                    op_run<Op>(J, cm, raw[it], base + c0, w, off, tc);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty0 + 8 * s);
            if (++s == stages) { s = 0; ph ^= 1u; }
            tile = nxt;
            if (tile >= total_tiles) break;
            mbar_wait_a(full0 + 8 * s, ph);
            uint32_t t2;
            lds_v2(ids0 + 8 * s, t2, nxt);
        }
    }
}

// ------------------------------------------------------------------------------------
// pipeline 0: direct global loads
// ------------------------------------------------------------------------------------
template <class Op>
__global__ void __launch_bounds__(DIRECT_THREADS) stream_direct_kernel(const __grid_constant__ JobTable tbl,
                                                                       const __grid_constant__ Common cm,
                                                                       uint32_t total_tiles) {
    constexpr int TILE_CHUNKS = TileOf<Op>::value;
    constexpr int G = Op::GROUP;
    constexpr int UNITS = TILE_CHUNKS / G;
    constexpr int ITERS = UNITS / DIRECT_THREADS;
    int j = 0;
    Job J = job_at(tbl, 0);
    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (tile >= J.tile_end) {
            while (tile >= job_tile_end(tbl, j)) ++j;
            J = job_at(tbl, j);
        }
        const uint32_t base = (tile - J.tile_begin) * TILE_CHUNKS;
        const uint32_t chunks = min((uint32_t)TILE_CHUNKS, J.n_chunks - base);
        uint32_t w[ITERS][G][Op::IN_BYTES / 4];
        typename Op::Raw raw[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t c0 = (uint32_t)(it * DIRECT_THREADS + threadIdx.x) * G;
            if (c0 < chunks) {
#pragma unroll
                for (int k = 0; k < G; ++k) ldg_chunk<Op::IN_BYTES, (G > 1)>(J.in + (size_t)(base + c0 + k) * Op::IN_BYTES, w[it][k]);
                raw[it] = Op::prefetch(J, base + c0);
            }
        }
        const auto tc = op_tile<Op>(J);
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t c0 = (uint32_t)(it * DIRECT_THREADS + threadIdx.x) * G;
            if (c0 < chunks) op_run<Op>(J, cm, raw[it], base + c0, w[it], 0, tc);
        }
    }
}

// ------------------------------------------------------------------------------------
// host-side launcher shared by all ops
// ------------------------------------------------------------------------------------
struct LaunchPlan {
    JobTable tbl;
    Common cm;
    uint32_t total_tiles;
    int tile_chunks = TILE_CHUNKS;   // the tile size the host counted tiles with; must equal TileOf<Op>::value of the kernel launched
    uint32_t* sched = nullptr;   // zeroed device counter for the dynamic tile schedule (nullptr = static deal); owned by dispatch.cu
};

template <class Op>
int launch_stream(const LaunchPlan& lp, int device, cudaStream_t stream) {
    if (lp.total_tiles == 0) return CT_OK;
    if (lp.tile_chunks != TileOf<Op>::value) {
        set_error("internal: tiles counted with %d chunks, the kernel uses %d", lp.tile_chunks, TileOf<Op>::value);
        return CT_E_ARG;
    }
    const Tuning tn = tuning();
    const int sms = sm_count(device);
    if (tn.pipe == 1) {
        constexpr int TILE_BYTES = TileOf<Op>::value * Op::IN_BYTES;
        int stages = tn.stages;
        int ctas = tn.ctas_per_sm;
        if constexpr (HasShape<Op>::value) {   // an op's measured preference, unless the caller tuned explicitly
            if (tn.auto_shape) { stages = Op::PREF_STAGES; ctas = Op::PREF_CTAS; }
        }
        if (stages < 2) stages = 2;
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        // keep stages * tile * ctas within ~200 KB of shared memory per SM
        while ((size_t)stages * TILE_BYTES * ctas + 1024 * ctas > 200 * 1024 && stages > 2) --stages;
        while ((size_t)stages * TILE_BYTES * ctas + 1024 * ctas > 200 * 1024 && ctas > 1) --ctas;
        const size_t smem = SMEM_HDR + (size_t)stages * TILE_BYTES;
        auto kfn = stream_tma_kernel<Op>;
        CT_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t grid = (uint32_t)(sms * ctas);
        if (grid > lp.total_tiles) grid = lp.total_tiles;
        kfn<<<grid, 32 * (NCW + 1), smem, stream>>>(lp.tbl, lp.cm, lp.total_tiles, stages, grid < lp.total_tiles ? lp.sched : nullptr);
    } else {
        int ctas = tn.ctas_per_sm > 0 ? tn.ctas_per_sm : 8;
        uint32_t grid = (uint32_t)(sms * ctas);
        if (grid > lp.total_tiles) grid = lp.total_tiles;
        stream_direct_kernel<Op><<<grid, DIRECT_THREADS, 0, stream>>>(lp.tbl, lp.cm, lp.total_tiles);
    }
    count_launch();
    CT_CUDA_TRY(cudaGetLastError());
    return CT_OK;
}

}  // namespace ctb
