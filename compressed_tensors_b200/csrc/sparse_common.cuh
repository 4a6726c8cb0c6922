// sparse_common.cuh -- 2:4 selection shared by the bitmask kernels (sparse.cu) and the fused 2:4 + int4 compressor (fast_sparse24q.cu)
#pragma once

#include "common.cuh"

namespace ctb {

// keep mask (bit j = element j kept) of one quad held as two packed words {e0,e1}, {e2,e3}, and the kept pair in column order
__device__ __forceinline__ uint32_t quad_select16(uint32_t w0, uint32_t w1, uint32_t& pair) {
    // composite = |x| bits << 2 | (3 - column): all four distinct, larger = wins (larger magnitude, or equal magnitude and lower
    // column).  The two winners come out of a 4-input selection network of integer min / max.
    const uint32_t c0 = ((w0 << 2) & 0x1fffcu) | 3u, c1 = ((w0 >> 14) & 0x1fffcu) | 2u;
    const uint32_t c2 = ((w1 << 2) & 0x1fffcu) | 1u, c3 = ((w1 >> 14) & 0x1fffcu);
    const uint32_t a = max(c0, c1), b = min(c0, c1), c = max(c2, c3), d = min(c2, c3);
    const uint32_t first = max(a, c), second = max(min(a, c), max(b, d));
    const uint32_t ia = 3u - (first & 3u), ib = 3u - (second & 3u);
    const uint32_t keep = (1u << ia) | (1u << ib);
    const uint32_t i0 = min(ia, ib), i1 = max(ia, ib);
    pair = __byte_perm(w0, w1, 0x1010u + i0 * 0x22u + i1 * 0x2200u);   // bytes (2 i0, 2 i0 + 1, 2 i1, 2 i1 + 1)
    return keep;
}

}  // namespace ctb
