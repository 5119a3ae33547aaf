// narrow.cuh — layout of the narrow count tables of the fast LDA path (kernels: lda_half16.cuh; conversions and the
// multi-GPU exchange: comm.cu).
#pragma once
#include "common.cuh"

namespace bias {

// A hot word's row is read by every SM all the time: the most frequent word of a Zipf corpus needs 36 GB per sweep out of the few L2
// slices its 4 KB live in, and the sweep waits for those lines (profiles/r02_hot_route_variants.md).  The hot rows therefore exist in
// kHotRep copies at different addresses: a CTA reads the copy of its index, every update goes to all copies (one RED instruction with
// more active lanes), and whoever else writes the first (canonical) copy — conversion, exchange — has the others refreshed before the
// next sweep.
#ifndef BIAS_HOT_REP
#define BIAS_HOT_REP 2
#endif
constexpr int kHotRep = BIAS_HOT_REP;        // at most 4 (the update is one RED instruction of a 16-lane half); 2, 3 and 4 copies measure the same

#ifdef __CUDACC__
// position of topic k's r inside the shared-memory vector (see the header comment)
__device__ __forceinline__ int r_pos(int k) {
    return (k & ~127) | (((k >> 2) & 1) << 6) | (((k >> 3) & 15) << 2) | (k & 3);
}

#endif

}  // namespace bias
