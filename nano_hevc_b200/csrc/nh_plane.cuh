// nh_plane.cuh -- neighbour fetch from a plane (block.py:38-55), shared by the frame coders.
#pragma once
#include "nh_common.cuh"

namespace nh {

// Neighbour fetch with the reference's substitution rules (block.py:38-55) and
// replicate-last padding (intra.py:174-178) folded into an index clamp.
// COHERENT: read through L2 (ld.cg) because another SM may just have written the sample.
template <bool COHERENT>
__device__ __forceinline__ int plane_px(const int16_t* plane, int64_t idx) {
    if constexpr (COHERENT) return (int)__ldcg(plane + idx);
    else return (int)__ldg(plane + idx);
}

template <bool COHERENT>
__device__ __forceinline__ int top_ref(const int16_t* plane, int H, int W, int pitch, int x, int y,
                                       int n_top, int k) {  // k = 0 .. 2N
    if (k == 0) return (x == 0 || y == 0) ? 128 : plane_px<COHERENT>(plane, (int64_t)(y - 1) * pitch + x - 1);
    if (y == 0) return 128;
    int last = x + n_top - 1;
    if (last > W - 1) last = W - 1;
    int col = x + k - 1;
    if (col > last) col = last;
    return plane_px<COHERENT>(plane, (int64_t)(y - 1) * pitch + col);
}

template <bool COHERENT>
__device__ __forceinline__ int left_ref(const int16_t* plane, int H, int W, int pitch, int x, int y,
                                        int n_left, int k) {
    if (k == 0) return (x == 0 || y == 0) ? 128 : plane_px<COHERENT>(plane, (int64_t)(y - 1) * pitch + x - 1);
    if (x == 0) return 128;
    int last = y + n_left - 1;
    if (last > H - 1) last = H - 1;
    int row = y + k - 1;
    if (row > last) row = last;
    return plane_px<COHERENT>(plane, (int64_t)row * pitch + x - 1);
}

}  // namespace nh
