// narrow.cuh — layout of the narrow count tables of the fast LDA path (kernels: lda_half16.cuh; conversions and the
// multi-GPU exchange: comm.cu).
#pragma once
#include "common.cuh"

namespace bias {

#ifdef __CUDACC__
// position of topic k's r inside the shared-memory vector (see the header comment)
__device__ __forceinline__ int r_pos(int k) {
    return (k & ~127) | (((k >> 2) & 1) << 6) | (((k >> 3) & 15) << 2) | (k & 3);
}

#endif

}  // namespace bias
