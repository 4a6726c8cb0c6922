// stream.cuh -- the two persistent streaming pipelines every hot-path op is instantiated on.
//
// Work decomposition (all hot-path ops are elementwise over a flat, contiguous tensor):
//   chunk = 8 consecutive elements            (one thread's unit of work)
//   tile  = TILE_CHUNKS chunks = 8192 elems   (one CTA's unit of work; 16 KB of bf16)
//   job   = one tensor; a launch covers a table of jobs (whole-model batches), tiles are
//           numbered globally and dealt round-robin to a persistent grid of
//           n_SM * ctas_per_sm CTAs.
//
// Pipeline 1 ("tma"): warp-specialised.  One producer lane issues 1-D bulk async copies
// (cp.async.bulk, the TMA engine without a tensor map) of whole input tiles into a ring of
// shared-memory stages, each guarded by a full/empty mbarrier pair; NCW consumer warps wait on
// the full barrier, read their chunks from shared memory with conflict-free vector LDS, do the
// arithmetic in registers and store results with coalesced streaming stores.  Bytes in flight
// per SM = stages * tile_bytes * ctas_per_sm, independent of register pressure.
// Pipeline 0 ("direct"): every thread issues its tile's 128-bit ld.global.nc loads up front
// (UNROLL = chunks per thread per tile) and relies on occupancy for memory-level parallelism.
#pragma once

#include "common.cuh"

namespace ctb {

constexpr int TILE_CHUNKS = 1024;   // 8192 elements
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
    FastDiv dc;            // chunks per scale element (scale index = chunk / dc)
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
// by-value copy: the job lives in registers for the duration of a tile
__device__ __forceinline__ Job job_at(const JobTable& t, int j) {
    if (t.n == 1) return t.one;
    return t.jobs[j];
}

// ------------------------------------------------------------------------------------
// Op concept:
//   static constexpr int IN_BYTES;            streamed input bytes per chunk (4, 8, 16 or 32)
//   struct Ctx;                               per-chunk scale / zero-point context
//   static Ctx prefetch(J, cm, gc);           issued BEFORE the data arrives
//   static void run(J, cm, ctx, gc, in[]);    in[] = IN_BYTES/4 words of this chunk
// ------------------------------------------------------------------------------------

template <int BYTES> struct InWords { uint32_t w[BYTES / 4]; };

template <int BYTES>
__device__ __forceinline__ void lds_chunk(const uint8_t* p, uint32_t (&w)[BYTES / 4]) {
    if constexpr (BYTES == 4) {
        w[0] = *reinterpret_cast<const uint32_t*>(p);
    } else if constexpr (BYTES == 8) {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        w[0] = v.x; w[1] = v.y;
    } else if constexpr (BYTES == 16) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        uint4 u = *reinterpret_cast<const uint4*>(p + 16);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = u.x; w[5] = u.y; w[6] = u.z; w[7] = u.w;
    }
}
template <int BYTES>
__device__ __forceinline__ void ldg_chunk(const uint8_t* p, uint32_t (&w)[BYTES / 4]) {
    if constexpr (BYTES == 4) {
        w[0] = ldg_stream4(p);
    } else if constexpr (BYTES == 8) {
        uint2 v = ldg_stream8(p);
        w[0] = v.x; w[1] = v.y;
    } else if constexpr (BYTES == 16) {
        uint4 v = ldg_stream16(p);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        uint4 v = ldg_stream16(p);
        uint4 u = ldg_stream16(p + 16);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        w[4] = u.x; w[5] = u.y; w[6] = u.z; w[7] = u.w;
    }
}

// ------------------------------------------------------------------------------------
// pipeline 1: TMA bulk-copy ring
// ------------------------------------------------------------------------------------
template <class Op>
__global__ void __launch_bounds__(32 * (NCW + 1)) stream_tma_kernel(const __grid_constant__ JobTable tbl,
                                                                   const __grid_constant__ Common cm,
                                                                   uint32_t total_tiles, int stages) {
    constexpr int TILE_BYTES = TILE_CHUNKS * Op::IN_BYTES;
    constexpr int CTHREADS = NCW * 32;
    constexpr int ITERS = TILE_CHUNKS / CTHREADS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + MAX_STAGES;
    uint8_t* data = smem_raw + 256;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
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
            for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                while (tile >= job_tile_end(tbl, j)) ++j;
                const Job J = job_at(tbl, j);
                const uint32_t lt = tile - J.tile_begin;
                const uint32_t base = lt * TILE_CHUNKS;
                const uint32_t chunks = min((uint32_t)TILE_CHUNKS, J.n_chunks - base);
                const uint32_t bytes = chunks * Op::IN_BYTES;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(data + (size_t)s * TILE_BYTES, J.in + (size_t)lt * TILE_BYTES, bytes, &full[s], policy);
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int ctid = threadIdx.x;  // 0 .. CTHREADS-1
        int s = 0, j = 0;
        uint32_t ph = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            while (tile >= job_tile_end(tbl, j)) ++j;
            const Job J = job_at(tbl, j);
            const uint32_t base = (tile - J.tile_begin) * TILE_CHUNKS;
            const uint32_t chunks = min((uint32_t)TILE_CHUNKS, J.n_chunks - base);
            typename Op::Ctx ctx[ITERS];
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c = it * CTHREADS + ctid;
                if (c < chunks) ctx[it] = Op::prefetch(J, cm, base + c);
            }
            mbar_wait(&full[s], ph);
            const uint8_t* sp = data + (size_t)s * TILE_BYTES;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c = it * CTHREADS + ctid;
                if (c < chunks) {
                    uint32_t w[Op::IN_BYTES / 4];
                    lds_chunk<Op::IN_BYTES>(sp + (size_t)c * Op::IN_BYTES, w);
                    Op::run(J, cm, ctx[it], base + c, w);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == stages) { s = 0; ph ^= 1u; }
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
    constexpr int ITERS = TILE_CHUNKS / DIRECT_THREADS;
    int j = 0;
    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        while (tile >= job_tile_end(tbl, j)) ++j;
        const Job J = job_at(tbl, j);
        const uint32_t base = (tile - J.tile_begin) * TILE_CHUNKS;
        const uint32_t chunks = min((uint32_t)TILE_CHUNKS, J.n_chunks - base);
        uint32_t w[ITERS][Op::IN_BYTES / 4];
        typename Op::Ctx ctx[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t c = it * DIRECT_THREADS + threadIdx.x;
            if (c < chunks) {
                ldg_chunk<Op::IN_BYTES>(J.in + (size_t)(base + c) * Op::IN_BYTES, w[it]);
                ctx[it] = Op::prefetch(J, cm, base + c);
            }
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t c = it * DIRECT_THREADS + threadIdx.x;
            if (c < chunks) Op::run(J, cm, ctx[it], base + c, w[it]);
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
};

template <class Op>
int launch_stream(const LaunchPlan& lp, int device, cudaStream_t stream) {
    if (lp.total_tiles == 0) return CT_OK;
    const Tuning tn = tuning();
    const int sms = sm_count(device);
    if (tn.pipe == 1) {
        constexpr int TILE_BYTES = TILE_CHUNKS * Op::IN_BYTES;
        int stages = tn.stages;
        int ctas = tn.ctas_per_sm;
        if (stages < 2) stages = 2;
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        // keep stages * tile * ctas within ~200 KB of shared memory per SM
        while ((size_t)stages * TILE_BYTES * ctas + 1024 * ctas > 200 * 1024 && stages > 2) --stages;
        while ((size_t)stages * TILE_BYTES * ctas + 1024 * ctas > 200 * 1024 && ctas > 1) --ctas;
        const size_t smem = 256 + (size_t)stages * TILE_BYTES;
        auto kfn = stream_tma_kernel<Op>;
        CT_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t grid = (uint32_t)(sms * ctas);
        if (grid > lp.total_tiles) grid = lp.total_tiles;
        kfn<<<grid, 32 * (NCW + 1), smem, stream>>>(lp.tbl, lp.cm, lp.total_tiles, stages);
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
