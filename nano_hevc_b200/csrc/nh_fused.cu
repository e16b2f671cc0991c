// nh_fused.cu -- K6: predict -> residual -> forward -> quantize -> dequantize ->
// inverse -> reconstruct -> clip in ONE kernel, DC / planar prediction from
// given N-sample references (BASELINE config 2).
//
// Reference composition: README.md:55-71, docs/frames_and_panes.md:319-338
// (intra.py:46-113, intra.py:65-78, transform.py:154-238, quant.py:41-123).
//
// Two thread mappings, both streaming every tensor through HBM exactly once:
//   * N = 4, 8  ("unit" kernels): one lane owns 64 pixels (one 8x8 block or
//     four consecutive 4x4 blocks) entirely in registers; both transform passes
//     run in-thread with no shuffles.  Global traffic is staged through a
//     per-warp padded shared-memory tile so that every LDG/STG is a fully
//     coalesced 128-bit access (512 contiguous bytes per warp instruction).
//   * N = 16, 32 ("rows" kernels): N lanes own one block; lane = row for global
//     I/O and the second pass, lane = column for the first pass, with the
//     transposition going through an N x (N+4) int32 shared-memory matrix.
#include "nh_block.cuh"

namespace nh {

struct FusedArgs {
    const int16_t* orig;
    const int16_t* top;
    const int16_t* left;
    const int16_t* top_right;
    const int16_t* bottom_left;
    const uint8_t* modes;
    int mode;
    int64_t n_blocks;
    QuantParams qp;
    int maxv;
    int16_t* pred;
    int32_t* coeff;
    int32_t* levels;
    int16_t* recon;
};

// ------------------------------------------------------------ unit kernels
constexpr int kUnitWarps = 8;  // 256 threads per CTA

template <int N, bool DST>
__global__ void __launch_bounds__(kUnitWarps * 32, 2) fused_unit_kernel(const FusedArgs a) {
    constexpr int NN = N * N;
    constexpr int BPU = 64 / NN;  // blocks per unit (lane)
    using T16 = WarpTile<128>;    // 64 int16 per lane
    using T32 = WarpTile<256>;    // 64 int32 per lane
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* s16 = smem_raw + warp * (T16::kBytes + T32::kBytes);
    unsigned char* s32 = s16 + T16::kBytes;

    const int64_t n_units = (a.n_blocks + BPU - 1) / BPU;
    const int64_t n_tiles = (n_units + 31) / 32;
    const int64_t warp_global = (int64_t)blockIdx.x * kUnitWarps + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * kUnitWarps;

    for (int64_t tile = warp_global; tile < n_tiles; tile += warp_stride) {
        const int64_t blk0 = tile * 32 * BPU;  // first block of the tile
        int64_t rem = a.n_blocks - blk0;
        const int blocks_valid = (int)(rem < 32 * BPU ? rem : 32 * BPU);
        const int chunks16 = blocks_valid * (NN * 2 / 16);  // valid 16-byte chunks, int16 tensors
        const int chunks32 = blocks_valid * (NN * 4 / 16);  // ... int32 tensors

        // -- stage the original pixels: global (coalesced) -> shared (padded)
        T16::load(s16, reinterpret_cast<const unsigned char*>(a.orig + blk0 * NN), lane, chunks16);
        __syncwarp();

        int res[BPU][N][N];
        uint4* u16 = T16::unit(s16, lane);
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
            const int64_t b = blk0 + (int64_t)lane * BPU + q;
            const bool valid = b < a.n_blocks;
            // references of this block
            int top[N], left[N], tr = 0, bl = 0, mode = a.mode;
            if (valid) {
                uint32_t tw[N / 2], lw[N / 2];
                load_row16<N>(a.top + b * N, tw);
                load_row16<N>(a.left + b * N, lw);
                unpack_row<N>(tw, top);
                unpack_row<N>(lw, left);
                tr = a.top_right[b];
                bl = a.bottom_left[b];
                if (a.modes) mode = a.modes[b];
            } else {
#pragma unroll
                for (int k = 0; k < N; ++k) top[k] = left[k] = 0;
            }
            int dc = 0;
            if (mode == 1) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) s += top[k] + left[k];
                dc = dc_value<N>(s);
            }
            // 8 pixels per 16-byte chunk: orig -> prediction + residual; the prediction
            // replaces the original pixels in the tile (it is needed again for the recon).
            int* rq = &res[q][0][0];
#pragma unroll
            for (int c = 0; c < NN / 8; ++c) {
                uint4 v = u16[q * (NN / 8) + c];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                int p[8];
                if (mode == 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) p[k] = dc;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int e = 8 * c + k, y = e / N, x = e % N;
                        p[k] = planar_px<N>(x, y, left[y], top[x], tr, bl);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int o = (k & 1) ? hi16(w[k >> 1]) : lo16(w[k >> 1]);
                    rq[8 * c + k] = sext16(o - sext16(p[k]));  // intra.py:65-67
                }
                u16[q * (NN / 8) + c] = make_uint4(pack16(p[0], p[1]), pack16(p[2], p[3]),
                                                   pack16(p[4], p[5]), pack16(p[6], p[7]));
            }
        }
        __syncwarp();
        if (a.pred) T16::store(s16, reinterpret_cast<unsigned char*>(a.pred + blk0 * NN), lane, chunks16);
        __syncwarp();  // the tile keeps the prediction until the reconstruction rewrites it

        // -- forward transform (in-thread, both passes)
        uint4* u32 = T32::unit(s32, lane);
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, false>(res[q]);
        if (a.coeff) {
#pragma unroll
            for (int q = 0; q < BPU; ++q)
#pragma unroll
                for (int e = 0; e < NN / 4; ++e) {
                    const int* r = &res[q][0][0];
                    u32[q * (NN / 4) + e] = make_uint4(r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
                }
            __syncwarp();
            T32::store(s32, reinterpret_cast<unsigned char*>(a.coeff + blk0 * NN), lane, chunks32);
            __syncwarp();
        }
        // -- quantize (levels out) and dequantize in place
#pragma unroll
        for (int q = 0; q < BPU; ++q)
#pragma unroll
            for (int e = 0; e < NN / 4; ++e) {
                int* r = &res[q][0][0] + 4 * e;
                int l0 = quantize_one(r[0], a.qp), l1 = quantize_one(r[1], a.qp);
                int l2 = quantize_one(r[2], a.qp), l3 = quantize_one(r[3], a.qp);
                if (a.levels) u32[q * (NN / 4) + e] = make_uint4(l0, l1, l2, l3);
                r[0] = dequantize_one(l0, a.qp);
                r[1] = dequantize_one(l1, a.qp);
                r[2] = dequantize_one(l2, a.qp);
                r[3] = dequantize_one(l3, a.qp);
            }
        if (a.levels) {
            __syncwarp();
            T32::store(s32, reinterpret_cast<unsigned char*>(a.levels + blk0 * NN), lane, chunks32);
        }
        // -- inverse transform, reconstruct against the prediction still in the int16 tile
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, true>(res[q]);
        if (a.recon) {
#pragma unroll
            for (int q = 0; q < BPU; ++q)
#pragma unroll
                for (int c = 0; c < NN / 8; ++c) {  // 8 pixels per 16-byte chunk
                    uint4 pv = u16[q * (NN / 8) + c];
                    const int* r = &res[q][0][0] + 8 * c;
                    uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, ow[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ow[k] = pack16(recon_px(lo16(pw[k]), r[2 * k], a.maxv),
                                       recon_px(hi16(pw[k]), r[2 * k + 1], a.maxv));
                    u16[q * (NN / 8) + c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            __syncwarp();
            T16::store(s16, reinterpret_cast<unsigned char*>(a.recon + blk0 * NN), lane, chunks16);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------ rows kernels
constexpr int kRowsWarps = 4;  // 128 threads per CTA

template <int N>
__global__ void __launch_bounds__(kRowsWarps * 32) fused_rows_kernel(const FusedArgs a) {
    constexpr int NN = N * N;
    constexpr int BPW = 32 / N;  // blocks per warp
    __shared__ __align__(16) int smem[kRowsWarps][BPW * RowsTile<N>::WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / N, r = lane % N;
    int* M = smem[warp] + g * RowsTile<N>::WORDS;

    const int64_t n_tiles = (a.n_blocks + BPW - 1) / BPW;
    const int64_t warp_global = (int64_t)blockIdx.x * kRowsWarps + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * kRowsWarps;

    for (int64_t tile = warp_global; tile < n_tiles; tile += warp_stride) {
        const int64_t b = tile * BPW + g;
        const bool valid = b < a.n_blocks;
        uint32_t pw[N / 2];  // prediction row, kept packed for the reconstruction
        {
            int o[N], p[N];
            if (valid) {
                uint32_t ow[N / 2];
                load_row16<N>(a.orig + b * NN + r * N, ow);
                unpack_row<N>(ow, o);
                int mode = a.modes ? (int)a.modes[b] : a.mode;
                uint32_t tw[N / 2];
                load_row16<N>(a.top + b * N, tw);
                if (mode == 1) {
                    uint32_t lw[N / 2];
                    load_row16<N>(a.left + b * N, lw);
                    int dc = dc_value<N>(sum_row<N>(tw) + sum_row<N>(lw));
#pragma unroll
                    for (int x = 0; x < N; ++x) p[x] = dc;
                } else {
                    int top[N];
                    unpack_row<N>(tw, top);
                    planar_row<N>(r, (int)a.left[b * N + r], top, (int)a.top_right[b],
                                  (int)a.bottom_left[b], p);
                }
            } else {
#pragma unroll
                for (int x = 0; x < N; ++x) o[x] = p[x] = 0;
            }
            pack_row<N>(p, pw);
            if (valid && a.pred) store_row16<N>(a.pred + b * NN + r * N, pw);
            int res[N];
#pragma unroll
            for (int x = 0; x < N; ++x) res[x] = sext16(o[x] - sext16(p[x]));
            store_row_smem<N>(M, r, res);
        }
        __syncwarp();
        col_pass<N, false, false>(M, r);
        __syncwarp();
        {
            int c[N], lv[N], dq[N];
            row_pass<N, false, false>(M, r, c);
            if (valid && a.coeff) store_row32<N>(a.coeff + b * NN + r * N, c);
            quant_dequant_row<N>(c, a.qp, lv, dq);
            if (valid && a.levels) store_row32<N>(a.levels + b * NN + r * N, lv);
            store_row_smem<N>(M, r, dq);  // lane r read row r and is the only writer of row r
        }
        __syncwarp();
        col_pass<N, false, true>(M, r);
        __syncwarp();
        {
            int res[N];
            row_pass<N, false, true>(M, r, res);
            if (valid && a.recon) {
                uint32_t ow[N / 2];
#pragma unroll
                for (int k = 0; k < N / 2; ++k)
                    ow[k] = pack16(recon_px(lo16(pw[k]), res[2 * k], a.maxv),
                                   recon_px(hi16(pw[k]), res[2 * k + 1], a.maxv));
                store_row16<N>(a.recon + b * NN + r * N, ow);
            }
        }
        __syncwarp();
    }
}

template <int N, bool DST>
static int launch_unit(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPU = 64 / (N * N);
    constexpr int kSmem = kUnitWarps * (WarpTile<128>::kBytes + WarpTile<256>::kBytes);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_unit_kernel<N, DST>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fused_unit_kernel)");
        configured = true;
    }
    int64_t units = (a.n_blocks + BPU - 1) / BPU;
    int grid = grid_for(units, (int64_t)kUnitWarps * 32, 2);
    fused_unit_kernel<N, DST><<<grid, kUnitWarps * 32, kSmem, st>>>(a);
    NH_CHECK_LAUNCH("fused_unit_kernel");
    return NH_OK;
}

template <int N>
static int launch_rows(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPW = 32 / N;
    int grid = grid_for(a.n_blocks, (int64_t)kRowsWarps * BPW, 4);
    fused_rows_kernel<N><<<grid, kRowsWarps * 32, 0, st>>>(a);
    NH_CHECK_LAUNCH("fused_rows_kernel");
    return NH_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace nh

NH_API int nh_fused_pipeline_dcplanar(const int16_t* orig, const int16_t* top, const int16_t* left,
                                      const int16_t* top_right, const int16_t* bottom_left,
                                      const uint8_t* modes, int mode, int64_t n_blocks, int size,
                                      int qp, int is_intra, int use_dst, int bit_depth,
                                      int16_t* pred, int32_t* coeff, int32_t* levels,
                                      int16_t* recon, void* stream) {
    using namespace nh;
    int l2 = log2_size(size);
    if (l2 < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (n_blocks == 0) return NH_OK;
    if (n_blocks < 0 || !orig || !top || !left || !top_right || !bottom_left) {
        set_error("nh_fused_pipeline_dcplanar: null input or negative block count");
        return NH_E_ARG;
    }
    if (!modes && mode != 0 && mode != 1) {
        set_error("nh_fused_pipeline_dcplanar: mode must be 0 (planar) or 1 (DC), got %d", mode);
        return NH_E_ARG;
    }
    if (bit_depth < 1 || bit_depth > 15) {
        set_error("nh_fused_pipeline_dcplanar: bit_depth %d out of range", bit_depth);
        return NH_E_ARG;
    }
    if (!aligned16(orig) || !aligned16(top) || !aligned16(left) || !aligned16(pred) ||
        !aligned16(coeff) || !aligned16(levels) || !aligned16(recon)) {
        set_error("nh_fused_pipeline_dcplanar: tensors must be 16-byte aligned");
        return NH_E_ARG;
    }
    if (n_blocks == 0) return NH_OK;
    FusedArgs a{orig, top, left, top_right, bottom_left, modes, mode, n_blocks,
                make_quant_params(qp, l2, is_intra), (1 << bit_depth) - 1,
                pred, coeff, levels, recon};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (size) {
        case 4: return use_dst ? launch_unit<4, true>(a, st) : launch_unit<4, false>(a, st);
        case 8: return launch_unit<8, false>(a, st);
        case 16: return launch_rows<16>(a, st);
        default: return launch_rows<32>(a, st);
    }
}
