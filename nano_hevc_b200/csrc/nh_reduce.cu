// nh_reduce.cu -- K9: the integer reductions behind nano_hevc/metrics.py.
//   mse / psnr  (metrics.py:7-21)   -> exact integer SSE here, float64 finish on the host
//   sad         (metrics.py:24-26)
//   satd_4x4    (metrics.py:29-43)  -> per block, summed over the 4x4 sub-blocks
//   residual_energy (metrics.py:46-48)
//   count_nonzero   (quant.py:171-173)
#include "nh_block.cuh"

namespace nh {

__device__ __forceinline__ void warp_block_atomic_add2(long long s0, long long s1, int64_t* out) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off);
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    }
    __shared__ long long sh[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = 0, t1 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += sh[0][w]; t1 += sh[1][w]; }
        atomicAdd(reinterpret_cast<unsigned long long*>(out), (unsigned long long)t0);
        if (out + 1) atomicAdd(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)t1);
    }
}

__device__ __forceinline__ void acc_pair(uint32_t wa, uint32_t wb, long long& sse, long long& sad) {
    int d0 = lo16(wa) - lo16(wb), d1 = hi16(wa) - hi16(wb);
    sse += (long long)d0 * d0 + (long long)d1 * d1;
    sad += abs(d0) + abs(d1);
}

// a, b: `rows` rows of `width` int16 with pitches pa / pb (a flat array is rows = 1).
__global__ void __launch_bounds__(256)
    sse_sad_kernel(const int16_t* __restrict__ a, int64_t pa, const int16_t* __restrict__ b,
                   int64_t pb, int64_t rows, int64_t width, int vec_ok, int64_t* out) {
    long long sse = 0, sad = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {  // rows start 16-byte aligned: 8 samples per 128-bit load
        const int64_t w8 = width / 8;
        for (int64_t t = tid; t < rows * w8; t += stride) {
            const int64_t r = t / w8, c = (t % w8) * 8;
            uint4 va = ldg_stream(a + r * pa + c), vb = ldg_stream(b + r * pb + c);
            acc_pair(va.x, vb.x, sse, sad);
            acc_pair(va.y, vb.y, sse, sad);
            acc_pair(va.z, vb.z, sse, sad);
            acc_pair(va.w, vb.w, sse, sad);
        }
        const int64_t tail = width - w8 * 8;
        for (int64_t t = tid; t < rows * tail; t += stride) {
            const int64_t r = t / tail, c = w8 * 8 + t % tail;
            int d = (int)a[r * pa + c] - (int)b[r * pb + c];
            sse += (long long)d * d;
            sad += abs(d);
        }
    } else {
        for (int64_t t = tid; t < rows * width; t += stride) {
            const int64_t r = t / width, c = t % width;
            int d = (int)a[r * pa + c] - (int)b[r * pb + c];
            sse += (long long)d * d;
            sad += abs(d);
        }
    }
    warp_block_atomic_add2(sse, sad, out);
}

__global__ void __launch_bounds__(256)
    count_nonzero_kernel(const int32_t* __restrict__ lv, int64_t n, int64_t* out) {
    long long c = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n4 = n / 4;
    for (int64_t i = tid; i < n4; i += stride) {
        uint4 v = ldg_stream(lv + 4 * i);
        c += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += stride) c += lv[i] != 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(out), (unsigned long long)c);
}

// quant.py:153-168 estimate_bits: sum(log2(|l| + 1)) in float64 plus the non-zero count (the host adds
// 2 bits per non-zero level and truncates, like int(np.sum(...))).
__global__ void __launch_bounds__(256)
    level_stats_kernel(const int32_t* __restrict__ lv, int64_t n, int64_t* nnz_out, double* sum_out) {
    long long c = 0;
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = lv[i];
        if (v != 0) {
            ++c;
            const unsigned a = (unsigned)(v < 0 ? -(long long)v : (long long)v);
            s += log2((double)a + 1.0);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, off);
        s += __shfl_xor_sync(0xffffffffu, s, off);
    }
    if ((threadIdx.x & 31) == 0 && c) {
        atomicAdd(reinterpret_cast<unsigned long long*>(nnz_out), (unsigned long long)c);
        atomicAdd(sum_out, s);
    }
}

// Per-block costs: one lane per 4x4 sub-block, N*N/16 lanes per block.
template <int N>
__global__ void __launch_bounds__(256)
    block_costs_kernel(const int16_t* __restrict__ a, const int16_t* __restrict__ b, int64_t n_blocks,
                       int32_t* __restrict__ sad, int32_t* __restrict__ satd,
                       int64_t* __restrict__ energy) {
    constexpr int SB = N * N / 16;
    constexpr int LPB = SB < 32 ? SB : 32;   // lanes per block
    constexpr int SPL = SB / LPB;            // sub-blocks per lane
    constexpr int SBW = N / 4;
    const int64_t total = n_blocks * LPB;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // all lanes of a warp iterate together (total is padded to a multiple of 32 for the shuffles)
    const int64_t total_pad = (total + 31) / 32 * 32;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_pad; t += stride) {
        const bool valid = t < total;
        const int64_t blk = valid ? t / LPB : 0;
        const int l = (int)(t % LPB);
        int s_sad = 0, s_satd = 0;
        long long s_en = 0;
        if (valid) {
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const int sb = l + i * LPB;
                const int sx = (sb % SBW) * 4, sy = (sb / SBW) * 4;
                int d[16];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int64_t off = blk * N * N + (sy + y) * N + sx;
                    uint2 va = *reinterpret_cast<const uint2*>(a + off);
                    uint2 vb = *reinterpret_cast<const uint2*>(b + off);
                    d[4 * y + 0] = lo16(va.x) - lo16(vb.x);
                    d[4 * y + 1] = hi16(va.x) - hi16(vb.x);
                    d[4 * y + 2] = lo16(va.y) - lo16(vb.y);
                    d[4 * y + 3] = hi16(va.y) - hi16(vb.y);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    s_sad += abs(d[e]);
                    s_en += (long long)d[e] * d[e];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int a0 = d[j] + d[4 + j], a1 = d[j] - d[4 + j];
                    int a2 = d[8 + j] + d[12 + j], a3 = d[8 + j] - d[12 + j];
                    d[j] = a0 + a2; d[4 + j] = a1 + a3; d[8 + j] = a0 - a2; d[12 + j] = a1 - a3;
                }
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2) {
                    int a0 = d[4 * i2] + d[4 * i2 + 1], a1 = d[4 * i2] - d[4 * i2 + 1];
                    int a2 = d[4 * i2 + 2] + d[4 * i2 + 3], a3 = d[4 * i2 + 2] - d[4 * i2 + 3];
                    s_satd += abs(a0 + a2) + abs(a1 + a3) + abs(a0 - a2) + abs(a1 - a3);
                }
            }
        }
#pragma unroll
        for (int off = LPB / 2; off > 0; off >>= 1) {
            s_sad += __shfl_xor_sync(0xffffffffu, s_sad, off);
            s_satd += __shfl_xor_sync(0xffffffffu, s_satd, off);
            s_en += __shfl_xor_sync(0xffffffffu, s_en, off);
        }
        if (valid && l == 0) {
            if (sad) sad[blk] = s_sad;
            if (satd) satd[blk] = s_satd;
            if (energy) energy[blk] = s_en;
        }
    }
}


// Wide-input variants behind the per-block metric wrappers (the reference widens before it reduces:
// metrics.py:9 float64, :26 / :33 int32, :48 int64), for callers that hand in uint16 samples or the int32
// output of inverse_transform:
//   out[0] = sum (a - b)^2 with the difference in int64, accumulated modulo 2^64   (residual_energy, :48)
//   out[1] = sum |a - b| with the difference and abs() wrapping in int32, summed in int64   (sad, :26)
//   fsum   = sum of double(a - b)^2, each square rounded like diff ** 2 on float64       (mse, :9-10)
__global__ void __launch_bounds__(256)
    metrics_i32_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n, int64_t* out,
                       double* fsum) {
    unsigned long long sse = 0;
    long long sad = 0;
    double fs = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t x = a[i], y = b ? b[i] : 0;
        const long long d = (long long)x - (long long)y;
        sse += (unsigned long long)d * (unsigned long long)d;
        const uint32_t w = (uint32_t)x - (uint32_t)y;              // int32 wrap-around difference
        const int32_t wa = (int32_t)((w & 0x80000000u) ? 0u - w : w);   // np.abs on int32 (INT_MIN stays INT_MIN)
        sad += (long long)wa;
        const double dd = (double)d;
        fs += dd * dd;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(fsum, fs);
    warp_block_atomic_add2((long long)sse, sad, out);
}

// mse (metrics.py:7-10) for float64 inputs: sum of (a - b)^2 in float64.
__global__ void __launch_bounds__(256)
    sse_f64_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* fsum) {
    double fs = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double d = a[i] - b[i];
        fs += d * d;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(fsum, fs);
}

// satd_4x4 (metrics.py:29-43) on int32 inputs with the reference's int32 wrap-around arithmetic; one thread
// per block pair, out (B,) int64.
__global__ void __launch_bounds__(128)
    satd4_i32_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n_blocks, int64_t* out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_blocks; t += (int64_t)gridDim.x * blockDim.x) {
        uint32_t d[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) d[e] = (uint32_t)a[t * 16 + e] - (uint32_t)b[t * 16 + e];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t a0 = d[j] + d[4 + j], a1 = d[j] - d[4 + j], a2 = d[8 + j] + d[12 + j], a3 = d[8 + j] - d[12 + j];
            d[j] = a0 + a2; d[4 + j] = a1 + a3; d[8 + j] = a0 - a2; d[12 + j] = a1 - a3;
        }
        long long s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t a0 = d[4 * i] + d[4 * i + 1], a1 = d[4 * i] - d[4 * i + 1];
            uint32_t a2 = d[4 * i + 2] + d[4 * i + 3], a3 = d[4 * i + 2] - d[4 * i + 3];
            const uint32_t r[4] = {a0 + a2, a1 + a3, a0 - a2, a1 - a3};
#pragma unroll
            for (int k = 0; k < 4; ++k) s += (long long)(int32_t)((r[k] & 0x80000000u) ? 0u - r[k] : r[k]);
        }
        out[t] = s;
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int launch_sse_sad(const int16_t* a, int64_t pa, const int16_t* b, int64_t pb, int64_t rows,
                          int64_t width, int64_t* out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(int64_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reduction output)");
    if (rows * width == 0) return NH_OK;
    int vec_ok = aligned16(a) && aligned16(b) && (rows == 1 || (pa % 8 == 0 && pb % 8 == 0));
    int grid = grid_for(rows * width / 8 + 1, 256, 4);
    sse_sad_kernel<<<grid, 256, 0, st>>>(a, pa, b, pb, rows, width, vec_ok, out);
    NH_CHECK_LAUNCH("sse_sad_kernel");
    return NH_OK;
}

}  // namespace nh

using namespace nh;

NH_API int nh_reduce_sse_sad(const int16_t* a, const int16_t* b, int64_t n, int64_t* out, void* stream) {
    if (!a || !b || !out || n < 0) { set_error("nh_reduce_sse_sad: null pointer or negative count"); return NH_E_ARG; }
    return launch_sse_sad(a, n, b, n, 1, n, out, reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_reduce_sse_sad_2d(const int16_t* a, int pitch_a, const int16_t* b, int pitch_b, int height,
                                int width, int64_t* out, void* stream) {
    if (!a || !b || !out || height < 0 || width < 0 || pitch_a < width || pitch_b < width) {
        set_error("nh_reduce_sse_sad_2d: bad argument");
        return NH_E_ARG;
    }
    return launch_sse_sad(a, pitch_a, b, pitch_b, height, width, out, reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_reduce_metrics_i32(const int32_t* a, const int32_t* b, int64_t n, int64_t* out, double* fsum,
                                 void* stream) {
    if (!a || !out || !fsum || n < 0) { set_error("nh_reduce_metrics_i32: null pointer or negative count"); return NH_E_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(int64_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(fsum, 0, sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reduction output)");
    if (n == 0) return NH_OK;
    metrics_i32_kernel<<<grid_for(n, 256 * 4, 4), 256, 0, st>>>(a, b, n, out, fsum);
    NH_CHECK_LAUNCH("metrics_i32_kernel");
    return NH_OK;
}

NH_API int nh_reduce_sse_f64(const double* a, const double* b, int64_t n, double* fsum, void* stream) {
    if (!a || !b || !fsum || n < 0) { set_error("nh_reduce_sse_f64: null pointer or negative count"); return NH_E_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(fsum, 0, sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reduction output)");
    if (n == 0) return NH_OK;
    sse_f64_kernel<<<grid_for(n, 256 * 4, 4), 256, 0, st>>>(a, b, n, fsum);
    NH_CHECK_LAUNCH("sse_f64_kernel");
    return NH_OK;
}

NH_API int nh_satd_4x4_i32(const int32_t* a, const int32_t* b, int64_t n_blocks, int64_t* out, void* stream) {
    if (!a || !b || !out || n_blocks < 0) { set_error("nh_satd_4x4_i32: null pointer or negative count"); return NH_E_ARG; }
    if (n_blocks == 0) return NH_OK;
    satd4_i32_kernel<<<grid_for(n_blocks, 128, 8), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, n_blocks, out);
    NH_CHECK_LAUNCH("satd4_i32_kernel");
    return NH_OK;
}

NH_API int nh_block_costs(const int16_t* a, const int16_t* b, int64_t n_blocks, int size, int32_t* sad,
                          int32_t* satd, int64_t* energy, void* stream) {
    if (log2_size(size) < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (!a || !b || n_blocks < 0 || !aligned16(a) || !aligned16(b)) {
        set_error("nh_block_costs: null / misaligned pointer or negative count");
        return NH_E_ARG;
    }
    if (n_blocks == 0) return NH_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int lpb = size * size / 16 < 32 ? size * size / 16 : 32;
    int grid = grid_for(n_blocks * lpb, 256, 8);
    switch (size) {
        case 4: block_costs_kernel<4><<<grid, 256, 0, st>>>(a, b, n_blocks, sad, satd, energy); break;
        case 8: block_costs_kernel<8><<<grid, 256, 0, st>>>(a, b, n_blocks, sad, satd, energy); break;
        case 16: block_costs_kernel<16><<<grid, 256, 0, st>>>(a, b, n_blocks, sad, satd, energy); break;
        default: block_costs_kernel<32><<<grid, 256, 0, st>>>(a, b, n_blocks, sad, satd, energy); break;
    }
    NH_CHECK_LAUNCH("block_costs_kernel");
    return NH_OK;
}

NH_API int nh_count_nonzero(const int32_t* levels, int64_t n, int64_t* out, void* stream) {
    if (!levels || !out || n < 0 || !aligned16(levels)) {
        set_error("nh_count_nonzero: null / misaligned pointer or negative count");
        return NH_E_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int64_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(count)");
    if (n == 0) return NH_OK;
    count_nonzero_kernel<<<grid_for(n / 4 + 1, 256, 4), 256, 0, st>>>(levels, n, out);
    NH_CHECK_LAUNCH("count_nonzero_kernel");
    return NH_OK;
}

NH_API int nh_level_stats(const int32_t* levels, int64_t n, int64_t* nnz_out, double* sum_log2_out,
                          void* stream) {
    if (!levels || !nnz_out || !sum_log2_out || n < 0) {
        set_error("nh_level_stats: null pointer or negative count");
        return NH_E_ARG;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(sum_log2_out, 0, sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(level stats)");
    if (n == 0) return NH_OK;
    level_stats_kernel<<<grid_for(n, 256 * 8, 4), 256, 0, st>>>(levels, n, nnz_out, sum_log2_out);
    NH_CHECK_LAUNCH("level_stats_kernel");
    return NH_OK;
}
