// sparse_common.cuh -- 2:4 selection shared by the bitmask kernels (sparse.cu) and the fused 2:4 + int4 compressor (fast_sparse24q.cu)
#pragma once

#include "common.cuh"

namespace ctb {

// keep mask (bit j = element j kept) of one quad held as two packed words {e0,e1}, {e2,e3}, and the kept pair in column order.
// composite key = |x| bits in [17, 32) | (3 - column) in the low bits: the four keys are distinct, larger = wins (larger magnitude,
// or equal magnitude and lower column).  The shift by 17 drops the sign bit of the low-half element for free, so a low-half key is
// ONE multiply-add (x * 2^17 + c) and a high-half key a mask and a multiply-add; the two winners come out of a 4-input selection
// network of integer min / max (7 instructions with the 3-input form); -0.0 has magnitude 0.
__device__ __forceinline__ uint32_t quad_select16(uint32_t w0, uint32_t w1, uint32_t& pair) {
    const uint32_t c0 = w0 * 0x20000u + 3u, c1 = (w0 & 0x7fff0000u) * 2u + 2u;
    const uint32_t c2 = w1 * 0x20000u + 1u, c3 = (w1 & 0x7fff0000u) * 2u;
    const uint32_t a = max(c0, c1), b = min(c0, c1), c = max(c2, c3), d = min(c2, c3);
    const uint32_t first = max(a, c), second = max(max(min(a, c), b), d);
    const uint32_t ja = first & 3u, jb = second & 3u;              // 3 - column of the two winners
    const uint32_t jhi = max(ja, jb), jlo = min(ja, jb);            // jhi belongs to the LOWER column
    // bytes (2 col, 2 col + 1) = 0x76 - 0x22 j: low half = the lower column's element
    pair = __byte_perm(w0, w1, 0x7676u - 0x22u * jhi - 0x2200u * jlo);
    return (8u >> ja) | (8u >> jb);
}

}  // namespace ctb
