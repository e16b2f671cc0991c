// nh_block.cuh -- device building blocks shared by the batched kernels:
//   * RowsTile<N>: an N x N int32 working matrix in shared memory on which a
//     group of N lanes runs the two separable passes (lane = column for the
//     first pass, lane = row for the second), used for N = 16 and 32 and by
//     the frame coder for every N;
//   * packed int16 row helpers;
//   * DC / planar row predictors.
#pragma once
#include "nh_common.cuh"

namespace nh {

template <int N>
struct RowsTile {
    // Row pitch in 32-bit words.  N+4 keeps every row 16-byte aligned and makes
    // both access patterns conflict-free: column sweeps (lane j reads word
    // k*PITCH + j) and per-lane 128-bit row reads (chunk index (PITCH/4)*i + q,
    // with PITCH/4 odd).
    static constexpr int PITCH = N + 4;
    // Blocks sharing a warp are skewed by N words so their columns fall into
    // different banks.
    static constexpr int WORDS = N * PITCH + (N < 32 ? N : 0);
};

// First (column) pass in place: lane j owns column j.
//   forward: temp[i][j] = (sum_k T[i][k] * X[k][j] + rnd) >> shift   transform.py:180-185
//   inverse: temp[i][j] = (sum_k T[k][i] * C[k][j] + rnd) >> shift   transform.py:222-227
template <int N, bool DST, bool INV>
__device__ __forceinline__ void col_pass(int* M, int j) {
    constexpr int P = RowsTile<N>::PITCH;
    int x[N], y[N];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = M[k * P + j];
    pass1d<N, DST, INV>(x, y);
#pragma unroll
    for (int k = 0; k < N; ++k) M[k * P + j] = y[k];
}

// Second (row) pass: lane i owns row i; result stays in registers.
//   forward: coeff[i][j] = (sum_k temp[i][k] * T[j][k] + rnd) >> shift   transform.py:189-194
//   inverse: res[i][j]   = (sum_k temp[i][k] * T[k][j] + rnd) >> shift   transform.py:231-236
template <int N, bool DST, bool INV>
__device__ __forceinline__ void row_pass(const int* M, int i, int (&out)[N]) {
    constexpr int P = RowsTile<N>::PITCH;
    int x[N];
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        int4 v = *reinterpret_cast<const int4*>(M + i * P + 4 * q);
        x[4 * q + 0] = v.x;
        x[4 * q + 1] = v.y;
        x[4 * q + 2] = v.z;
        x[4 * q + 3] = v.w;
    }
    pass1d<N, DST, INV>(x, out);
}

template <int N>
__device__ __forceinline__ void store_row_smem(int* M, int i, const int (&v)[N]) {
    constexpr int P = RowsTile<N>::PITCH;
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
        *reinterpret_cast<int4*>(M + i * P + 4 * q) =
            make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// Both separable passes of a 2-D transform with ONE inlined copy of the butterfly: every pass
// reads a column (lane = column index), transforms it and -- after the first pass -- writes the
// result back as a ROW, i.e. transposed, so that the second pass is again "read column lane".
//   forward: pass 0  temp[i][j] = sum_k T[i][k] X[k][j]      (lane j; stored as Mt[j][i])
//            pass 1  coeff[i][j] = sum_k temp[i][k] T[j][k]  (lane i reads Mt[k][i] = temp[i][k])
//   inverse: same with T^T (transform.py:222-236).
// Sharing the code between the passes halves the instruction footprint of the 16- and 32-point
// kernels (70 KB -> 38 KB at N = 32; profiles/r1_notes.md).  On return `out` is row `lane` of the result.
template <int N, bool DST, bool INV, bool DP = false>
__device__ __forceinline__ void two_pass_transform(int* M, int lane, bool active, int (&out)[N]) {
    constexpr int P = RowsTile<N>::PITCH;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (active) {
            int x[N];
#pragma unroll
            for (int k = 0; k < N; ++k) x[k] = M[k * P + lane];
            pass1d<N, DST, INV, DP>(x, out);
        }
        if (pass == 0) {
            __syncwarp();  // every lane has read its column
            if (active) store_row_smem<N>(M, lane, out);
            __syncwarp();
        }
    }
}

// ---- packed rows: N int16 as N/2 32-bit words --------------------------------
template <int N>
__device__ __forceinline__ void load_row16(const int16_t* p, uint32_t (&w)[N / 2]) {
    if constexpr (N == 4) {
        uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
        w[0] = v.x;
        w[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < N / 8; ++q) {
            uint4 v = ldg_stream(p + 8 * q);
            w[4 * q + 0] = v.x;
            w[4 * q + 1] = v.y;
            w[4 * q + 2] = v.z;
            w[4 * q + 3] = v.w;
        }
    }
}
template <int N>
__device__ __forceinline__ void store_row16(int16_t* p, const uint32_t (&w)[N / 2]) {
    if constexpr (N == 4) {
        __stcs(reinterpret_cast<uint2*>(p), make_uint2(w[0], w[1]));
    } else {
#pragma unroll
        for (int q = 0; q < N / 8; ++q)
            stg_stream(p + 8 * q, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
    }
}
template <int N>
__device__ __forceinline__ void store_row32(int32_t* p, const int (&v)[N]) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
        stg_stream(p + 4 * q, make_uint4((uint32_t)v[4 * q], (uint32_t)v[4 * q + 1],
                                         (uint32_t)v[4 * q + 2], (uint32_t)v[4 * q + 3]));
}
template <int N>
__device__ __forceinline__ void load_row32(const int32_t* p, int (&v)[N]) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        uint4 t = ldg_stream(p + 4 * q);
        v[4 * q + 0] = (int)t.x;
        v[4 * q + 1] = (int)t.y;
        v[4 * q + 2] = (int)t.z;
        v[4 * q + 3] = (int)t.w;
    }
}
template <int N>
__device__ __forceinline__ void unpack_row(const uint32_t (&w)[N / 2], int (&v)[N]) {
#pragma unroll
    for (int q = 0; q < N / 2; ++q) {
        v[2 * q] = lo16(w[q]);
        v[2 * q + 1] = hi16(w[q]);
    }
}
template <int N>
__device__ __forceinline__ void pack_row(const int (&v)[N], uint32_t (&w)[N / 2]) {
#pragma unroll
    for (int q = 0; q < N / 2; ++q) w[q] = pack16(v[2 * q], v[2 * q + 1]);
}
template <int N>
__device__ __forceinline__ int sum_row(const uint32_t (&w)[N / 2]) {
    int s = 0;
#pragma unroll
    for (int q = 0; q < N / 2; ++q) s += lo16(w[q]) + hi16(w[q]);
    return s;
}

// intra.py:109-111, one row y of the planar predictor.
template <int N>
__device__ __forceinline__ void planar_row(int y, int left_y, const int (&top)[N], int tr, int bl,
                                           int (&out)[N]) {
#pragma unroll
    for (int x = 0; x < N; ++x) out[x] = planar_px<N>(x, y, left_y, top[x], tr, bl);
}

// Quant -> (levels) -> dequant on a row held in registers (quant.py:41-123).
template <int N>
__device__ __forceinline__ void quant_dequant_row(const int (&coeff)[N], const QuantParams& qp,
                                                  int (&lv)[N], int (&dq)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        lv[k] = quantize_one(coeff[k], qp);
        dq[k] = dequantize_one(lv[k], qp);
    }
}

// intra.py:70-78: int16 wrap-around add of the (truncated) residual, then clip.
__device__ __forceinline__ int recon_px(int pred, int res, int maxv) {
    return clip_pixel(sext16(pred + sext16(res)), maxv);
}

}  // namespace nh
