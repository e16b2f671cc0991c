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
#include <cuda.h>

#include <cstdlib>

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
    // "narrow" outputs of the host-buffer pipeline: coefficients / levels as int16 (they fit in the
    // pixel domain: |coeff| <= 32394, |level| <= 13600), halving what has to cross PCIe.  A lane
    // that leaves the pixel domain raises *ood_flag and the host redoes the chunk through int32.
    int16_t* coeff16;
    int16_t* levels16;
    int* ood_flag;
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

// ------------------------------------------------------- unit kernels, v2
// Same lane mapping as above, restructured around asynchronous copies:
//   * the next tile's pixels are prefetched with cp.async (LDGSTS) into the second int16 tile
//     while the current tile is being coded (hides the long-scoreboard stall of v1);
//   * (measured and rejected, profiles/r1_notes.md: per-lane TMA 1-D bulk stores -- UBLKCP takes
//     uniform-register operands, so a per-lane bulk store compiles to a 32-iteration waterfall
//     loop; outputs therefore keep the cooperative, fully coalesced LDS.128 + STG.128 sweep);
//   * quant / dequant run in 32 bits when the lane's inputs lie in the pixel domain [0, 4095]
//     (checked on the packed words as they stream in); a lane that sees anything else recodes
//     its unit with the exact int64 reference arithmetic in a cold, loop-based routine.
constexpr int kV2Warps = 4;  // 128 threads per CTA, 3 CTAs per SM

template <int N>
__device__ __noinline__ void slow_transform(int* blk, bool dst, bool inv) {
    constexpr int shift = Log2<N>::v + 5;
    constexpr int rnd = 1 << (shift - 1);
    int tmp[N * N];
    auto T = [&](int i, int k) -> int {
        return (dst && N == 4) ? dst4(i & 3, k & 3) : cosv((i * (32 / N)) * (2 * k + 1));
    };
#pragma unroll 1
    for (int i = 0; i < N; ++i)
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            unsigned acc = 0;
#pragma unroll 1
            for (int k = 0; k < N; ++k) acc += (unsigned)(inv ? T(k, i) : T(i, k)) * (unsigned)blk[k * N + j];
            tmp[i * N + j] = (int)(acc + (unsigned)rnd) >> shift;
        }
#pragma unroll 1
    for (int i = 0; i < N; ++i)
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            unsigned acc = 0;
#pragma unroll 1
            for (int k = 0; k < N; ++k) acc += (unsigned)tmp[i * N + k] * (unsigned)(inv ? T(k, j) : T(j, k));
            blk[i * N + j] = (int)(acc + (unsigned)rnd) >> shift;
        }
}

// Exact (any int16 input) coding of one block straight from / to global memory.  Cold path.
template <int N>
__device__ __noinline__ void slow_block(const FusedArgs& a, int64_t b, bool dst) {
    constexpr int NN = N * N;
    int blk[NN];
    int16_t pr[NN];
    const int mode = a.modes ? (int)a.modes[b] : a.mode;
    int s = 0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) s += (int)a.top[b * N + k] + (int)a.left[b * N + k];
    const int dc = dc_value<N>(s);
    const int tr = a.top_right[b], bl = a.bottom_left[b];
#pragma unroll 1
    for (int e = 0; e < NN; ++e) {
        const int y = e / N, x = e % N;
        const int p = mode == 1 ? dc : planar_px<N>(x, y, (int)a.left[b * N + y], (int)a.top[b * N + x], tr, bl);
        pr[e] = (int16_t)p;
        blk[e] = sext16((int)a.orig[b * NN + e] - sext16(p));
        if (a.pred) a.pred[b * NN + e] = (int16_t)p;
    }
    slow_transform<N>(blk, dst, false);
#pragma unroll 1
    for (int e = 0; e < NN; ++e) {
        if (a.coeff) a.coeff[b * NN + e] = blk[e];
        const int lv = quantize_one(blk[e], a.qp);
        if (a.levels) a.levels[b * NN + e] = lv;
        blk[e] = dequantize_one(lv, a.qp);
    }
    slow_transform<N>(blk, dst, true);
    if (a.recon) {
#pragma unroll 1
        for (int e = 0; e < NN; ++e) a.recon[b * NN + e] = (int16_t)recon_px((int)pr[e], blk[e], a.maxv);
    }
}

template <int N, bool DST, bool NARROW>
__global__ void __launch_bounds__(kV2Warps * 32, 3) fused_unit_kernel_v2(const FusedArgs a, const FastQuant fq) {
    constexpr int NN = N * N;
    constexpr int BPU = 64 / NN;
    using T16 = WarpTile<128>;
    using T32 = WarpTile<256>;
    constexpr int kWarpBytes = 2 * T16::kBytes + T32::kBytes;  // 2 pixel tiles (double buffer) + 1 int32 tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + warp * kWarpBytes;
    // (a function of the buffer index, not an array of pointers: indexing a pointer array with a run-time value makes
    // the compiler forget the address space -- generic LD / ST instead of LDS / STS, tracked on the long scoreboard)
    auto s16 = [&](int i) -> unsigned char* { return wbase + i * T16::kBytes; };
    unsigned char* s32 = wbase + 2 * T16::kBytes;

    const int64_t n_units = (a.n_blocks + BPU - 1) / BPU;
    const int64_t n_tiles = (n_units + 31) / 32;
    const int64_t warp_stride = (int64_t)gridDim.x * kV2Warps;
    int64_t tile = (int64_t)blockIdx.x * kV2Warps + warp;

    auto prefetch = [&](int64_t t, unsigned char* dst) {
        const int64_t blk0 = t * 32 * BPU;
        const int64_t rem = a.n_blocks - blk0;
        const int chunks = (int)(rem < 32 * BPU ? rem : 32 * BPU) * (NN * 2 / 16);
        const unsigned char* g = reinterpret_cast<const unsigned char*>(a.orig + blk0 * NN);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int c = it * 32 + lane;
            if (c < chunks) cp_async16(smem_u32(dst + (c >> 3) * T16::kPitch + (c & 7) * 16), g + (size_t)c * 16);
        }
    };

    // References of one lane's unit, software-pipelined one tile ahead like the pixels.
    struct Refs {
        uint32_t tw[BPU][N / 2], lw[BPU][N / 2];
        int tr[BPU], bl[BPU], mode[BPU];
    };
    auto load_refs = [&](int64_t t, Refs& r) {
        const int64_t ub = (t * 32 + lane) * BPU;
        const int64_t urem = a.n_blocks - ub;
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
            r.mode[q] = a.mode;
            if (q < urem) {
                load_row16<N>(a.top + (ub + q) * N, r.tw[q]);
                load_row16<N>(a.left + (ub + q) * N, r.lw[q]);
                r.tr[q] = a.top_right[ub + q];
                r.bl[q] = a.bottom_left[ub + q];
                if (a.modes) r.mode[q] = a.modes[ub + q];
            } else {
#pragma unroll
                for (int k = 0; k < N / 2; ++k) r.tw[q][k] = r.lw[q][k] = 0;
                r.tr[q] = r.bl[q] = 0;
            }
        }
    };

    Refs nxt;
    if (tile < n_tiles) {
        prefetch(tile, s16(0));
        load_refs(tile, nxt);
    }
    cp_async_commit();
    int cur = 0;
    for (; tile < n_tiles; tile += warp_stride, cur ^= 1) {
        const int64_t blk0 = tile * 32 * BPU;
        const int64_t ub = blk0 + (int64_t)lane * BPU;  // first block of this lane's unit
        const int64_t trem = a.n_blocks - blk0;
        const int blocks_valid = (int)(trem < 32 * BPU ? trem : 32 * BPU);
        const int chunks16 = blocks_valid * (NN * 2 / 16), chunks32 = blocks_valid * (NN * 4 / 16);
        const int64_t urem = a.n_blocks - ub;
        const int ublocks = urem >= BPU ? BPU : (urem > 0 ? (int)urem : 0);  // valid blocks in the unit

        // this tile's references were fetched one iteration ago
        uint32_t tw[BPU][N / 2], lw[BPU][N / 2];
        int tr[BPU], bl[BPU], mode[BPU];
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
#pragma unroll
            for (int k = 0; k < N / 2; ++k) { tw[q][k] = nxt.tw[q][k]; lw[q][k] = nxt.lw[q][k]; }
            tr[q] = nxt.tr[q]; bl[q] = nxt.bl[q]; mode[q] = nxt.mode[q];
        }
        if (tile + warp_stride < n_tiles) load_refs(tile + warp_stride, nxt);
        // the other pixel tile is free (its reconstruction left at the end of the previous
        // iteration): start fetching the next tile into it, then wait for the current one
        if (tile + warp_stride < n_tiles) prefetch(tile + warp_stride, s16(cur ^ 1));
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();

        int res[BPU][N][N];
        uint4* u16 = T16::unit(s16(cur), lane);
        uint32_t ood = 0;  // out-of-domain bits: any sample outside [0, 4095]
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
            int top[N], left[N];
            unpack_row<N>(tw[q], top);
            unpack_row<N>(lw[q], left);
#pragma unroll
            for (int k = 0; k < N / 2; ++k) ood |= (tw[q][k] | lw[q][k]) & 0xF000F000u;
            ood |= (uint32_t)(tr[q] | bl[q]) & 0xFFFFF000u;
            int* rq = &res[q][0][0];
            if (mode[q] == 1) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) s += top[k] + left[k];
                const int dc = dc_value<N>(s);
                const uint32_t dcw = pack16(dc, dc);
#pragma unroll
                for (int c = 0; c < NN / 8; ++c) {
                    uint4 v = u16[q * (NN / 8) + c];
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ood |= w[k] & 0xF000F000u;
                        rq[8 * c + 2 * k] = lo16(w[k]) - dc;
                        rq[8 * c + 2 * k + 1] = hi16(w[k]) - dc;
                    }
                    u16[q * (NN / 8) + c] = make_uint4(dcw, dcw, dcw, dcw);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NN / 8; ++c) {
                    uint4 v = u16[q * (NN / 8) + c];
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    int p[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int e = 8 * c + k, y = e / N, x = e % N;
                        p[k] = planar_px<N>(x, y, left[y], top[x], tr[q], bl[q]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ood |= w[k] & 0xF000F000u;
                        rq[8 * c + 2 * k] = lo16(w[k]) - p[2 * k];
                        rq[8 * c + 2 * k + 1] = hi16(w[k]) - p[2 * k + 1];
                    }
                    u16[q * (NN / 8) + c] = make_uint4(pack16(p[0], p[1]), pack16(p[2], p[3]),
                                                       pack16(p[4], p[5]), pack16(p[6], p[7]));
                }
            }
        }
        const bool fast = ood == 0;  // in the pixel domain: 32-bit arithmetic is exact
        __syncwarp();
        if (a.pred) T16::store(s16(cur), reinterpret_cast<unsigned char*>(a.pred + blk0 * NN), lane, chunks16);
        __syncwarp();  // the tile keeps the prediction until the reconstruction rewrites it

        // -- forward transform (in-thread, both passes)
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, false>(res[q]);
        int* flat = &res[0][0][0];
        uint4* u32 = T32::unit(s32, lane);
        uint4* n16 = T16::unit(s32, lane);  // the same tile viewed as 64 int16 per lane (narrow outputs)
        if (NARROW) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                n16[e] = make_uint4(pack16(flat[8 * e], flat[8 * e + 1]), pack16(flat[8 * e + 2], flat[8 * e + 3]),
                                    pack16(flat[8 * e + 4], flat[8 * e + 5]), pack16(flat[8 * e + 6], flat[8 * e + 7]));
            __syncwarp();
            T16::store(s32, reinterpret_cast<unsigned char*>(a.coeff16 + blk0 * NN), lane, chunks16);
            __syncwarp();
        } else if (a.coeff) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                u32[e] = make_uint4(flat[4 * e], flat[4 * e + 1], flat[4 * e + 2], flat[4 * e + 3]);
            __syncwarp();
            T32::store(s32, reinterpret_cast<unsigned char*>(a.coeff + blk0 * NN), lane, chunks32);
            __syncwarp();
        }
        // -- quantise (levels out) and dequantise in place, 32-bit pixel-domain arithmetic
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int* r = flat + 8 * e;
            int l[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                l[k] = quantize_fast(r[k], fq);
                r[k] = dequantize_fast(l[k], fq);
            }
            if (NARROW) {
                n16[e] = make_uint4(pack16(l[0], l[1]), pack16(l[2], l[3]), pack16(l[4], l[5]), pack16(l[6], l[7]));
            } else if (a.levels) {
                u32[2 * e] = make_uint4(l[0], l[1], l[2], l[3]);
                u32[2 * e + 1] = make_uint4(l[4], l[5], l[6], l[7]);
            }
        }
        if (NARROW) {
            __syncwarp();
            T16::store(s32, reinterpret_cast<unsigned char*>(a.levels16 + blk0 * NN), lane, chunks16);
        } else if (a.levels) {
            __syncwarp();
            T32::store(s32, reinterpret_cast<unsigned char*>(a.levels + blk0 * NN), lane, chunks32);
        }
        // -- inverse transform, reconstruct against the prediction still in the pixel tile
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, true>(res[q]);
        if (a.recon) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 pv = u16[c];
                const int* r = flat + 8 * c;
                uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, ow[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ow[k] = pack16(recon_px(lo16(pw[k]), r[2 * k], a.maxv),
                                   recon_px(hi16(pw[k]), r[2 * k + 1], a.maxv));
                u16[c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
            __syncwarp();
            T16::store(s16(cur), reinterpret_cast<unsigned char*>(a.recon + blk0 * NN), lane, chunks16);
        }
        __syncwarp();
        // -- a lane whose inputs left the pixel domain recodes its unit exactly (cold path); the
        //    __syncwarp above orders the cooperative stores of its unit before these stores
        if (!fast) {
            if (NARROW && ublocks > 0) *a.ood_flag = 1;
            const FusedArgs a_cold = a;  // address taken here only, not on the hot path
            for (int q = 0; q < ublocks; ++q) slow_block<N>(a_cold, ub + q, DST);
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------- 4x4 kernel, looped (generation 4 at N = 4)
// Generation 2 keeps a lane's four 4x4 blocks (64 samples) in registers with everything unrolled: a
// 64 KB hot loop that ncu shows waiting on instruction fetch (no-instruction 1.16 warps per issue,
// profiles/r1_fused4_ncu_summary.json).  Here a lane codes ONE block at a time in two rolled loops
// over four rounds (round q = block q*32 + lane of the 128-block warp tile), with the coefficients
// parked in the int32 staging tile between the loops: a few more shared-memory reads, a hot loop of
// about 10 KB.  Same tiles, same cooperative 128-bit sweeps, same exact cold path as generation 2.
template <bool DST>
__global__ void __launch_bounds__(kV2Warps * 32, 3) fused_unit4_kernel(const FusedArgs a, const FastQuant fq,
                                                                        unsigned int* tile_counter) {
    constexpr int N = 4, NN = 16;
    using T16 = WarpTile<128>;
    using T32 = WarpTile<256>;
    constexpr int kWarpBytes = 2 * T16::kBytes + T32::kBytes;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + warp * kWarpBytes;
    // (a function of the buffer index, not an array of pointers: indexing a pointer array with a run-time value makes
    // the compiler forget the address space -- generic LD / ST instead of LDS / STS, tracked on the long scoreboard)
    auto s16 = [&](int i) -> unsigned char* { return wbase + i * T16::kBytes; };
    unsigned char* s32 = wbase + 2 * T16::kBytes;
    // block j of the tile lives in unit j >> 2 (padded pitch), slot j & 3
    auto px_of = [&](unsigned char* tile, int j) { return reinterpret_cast<uint4*>(tile + (j >> 2) * T16::kPitch + (j & 3) * 32); };
    auto i32_of = [&](int j) { return reinterpret_cast<uint4*>(s32 + (j >> 2) * T32::kPitch + (j & 3) * 64); };

    const int64_t n_tiles = (a.n_blocks + 127) / 128;
    // tiles are handed out dynamically, one ticket ahead (a static partition waits for the slowest SM)
    auto next_tile = [&]() -> int64_t {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        return (int64_t)__shfl_sync(0xffffffffu, t, 0);
    };
    int64_t tile = next_tile();
    int64_t tile_next = tile < n_tiles ? next_tile() : n_tiles;
    auto prefetch = [&](int64_t t, unsigned char* dst) {
        const int64_t blk0 = t * 128;
        const int64_t rem = a.n_blocks - blk0;
        const int chunks = (int)(rem < 128 ? rem : 128) * 2;
        const unsigned char* g = reinterpret_cast<const unsigned char*>(a.orig + blk0 * NN);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int c = it * 32 + lane;
            if (c < chunks) cp_async16(smem_u32(dst + (c >> 3) * T16::kPitch + (c & 7) * 16), g + (size_t)c * 16);
        }
    };
    struct Refs {
        uint2 tw, lw;
        int tr, bl, mode;
    };
    auto load_refs = [&](int64_t t, int q, Refs& rf) {  // block q*32 + lane of tile t
        const int64_t b = t * 128 + q * 32 + lane;
        if (t < n_tiles && b < a.n_blocks) {
            rf.tw = __ldcs(reinterpret_cast<const uint2*>(a.top + b * N));
            rf.lw = __ldcs(reinterpret_cast<const uint2*>(a.left + b * N));
            rf.tr = a.top_right[b];
            rf.bl = a.bottom_left[b];
            rf.mode = a.modes ? (int)a.modes[b] : a.mode;
        } else {
            rf.tw = rf.lw = make_uint2(0u, 0u);
            rf.tr = rf.bl = 0;
            rf.mode = 1;
        }
    };
    // references of the four rounds of the coming tile, one register set per round, each refilled for the NEXT
    // tile right after its round has used it: a full tile of lead.  (Round 1 kept two sets refilled two rounds
    // ahead -- only one round1() of lead for rounds 2 / 3: ncu put 45 % of the stall samples on the first use of
    // those loads, long-scoreboard 2.7 warps per issue.)
    Refs rf0, rf1, rf2, rf3;
    load_refs(tile, 0, rf0);
    load_refs(tile, 1, rf1);
    load_refs(tile, 2, rf2);
    load_refs(tile, 3, rf3);
    if (tile < n_tiles) prefetch(tile, s16(0));
    cp_async_commit();
    int cur = 0;
    int64_t tile_after = n_tiles;
    for (; tile < n_tiles; tile = tile_next, tile_next = tile_after, cur ^= 1) {
        tile_after = tile_next < n_tiles ? next_tile() : n_tiles;
        const int64_t blk0 = tile * 128;
        const int64_t trem = a.n_blocks - blk0;
        const int blocks_valid = (int)(trem < 128 ? trem : 128);
        const int chunks16 = blocks_valid * 2, chunks32 = blocks_valid * 4;
        if (tile_next < n_tiles) prefetch(tile_next, s16(cur ^ 1));
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        unsigned char* px = s16(cur);
        uint32_t ood_mask = 0;  // bit q: this lane's block of round q left the pixel domain [0, 4095]
        // ---- loop 1: predict, residual, forward transform; prediction and coefficients into the tiles.
        auto round1 = [&](int q, const Refs& rf) {
            const int j = q * 32 + lane;
            uint4* p16 = px_of(px, j);
            const uint4 v0 = p16[0], v1 = p16[1];
            const uint32_t ow[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint32_t ood = (rf.tw.x | rf.tw.y | rf.lw.x | rf.lw.y) & 0xF000F000u;
            ood |= (uint32_t)(rf.tr | rf.bl) & 0xFFFFF000u;
            const int top[4] = {lo16(rf.tw.x), hi16(rf.tw.x), lo16(rf.tw.y), hi16(rf.tw.y)};
            const int left[4] = {lo16(rf.lw.x), hi16(rf.lw.x), lo16(rf.lw.y), hi16(rf.lw.y)};
            int p[16];
            if (rf.mode == 1) {
                const int dc = dc_value<N>(top[0] + top[1] + top[2] + top[3] + left[0] + left[1] + left[2] + left[3]);
#pragma unroll
                for (int e = 0; e < 16; ++e) p[e] = dc;
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) p[e] = planar_px<N>(e & 3, e >> 2, left[e >> 2], top[e & 3], rf.tr, rf.bl);
            }
            int res[4][4];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ood |= ow[k] & 0xF000F000u;
                res[k >> 1][2 * (k & 1)] = lo16(ow[k]) - p[2 * k];
                res[k >> 1][2 * (k & 1) + 1] = hi16(ow[k]) - p[2 * k + 1];
            }
            p16[0] = make_uint4(pack16(p[0], p[1]), pack16(p[2], p[3]), pack16(p[4], p[5]), pack16(p[6], p[7]));
            p16[1] = make_uint4(pack16(p[8], p[9]), pack16(p[10], p[11]), pack16(p[12], p[13]), pack16(p[14], p[15]));
            transform2d<N, DST, false>(res);
            uint4* c32 = i32_of(j);
#pragma unroll
            for (int i = 0; i < 4; ++i) c32[i] = make_uint4(res[i][0], res[i][1], res[i][2], res[i][3]);
            ood_mask |= (ood != 0 ? 1u : 0u) << q;
        };
        round1(0, rf0);
        load_refs(tile_next, 0, rf0);
        round1(1, rf1);
        load_refs(tile_next, 1, rf1);
        round1(2, rf2);
        load_refs(tile_next, 2, rf2);
        round1(3, rf3);
        load_refs(tile_next, 3, rf3);
        __syncwarp();
        if (a.pred) T16::store(px, reinterpret_cast<unsigned char*>(a.pred + blk0 * NN), lane, chunks16);
        if (a.coeff) T32::store(s32, reinterpret_cast<unsigned char*>(a.coeff + blk0 * NN), lane, chunks32);
        __syncwarp();
        // ---- loop 2: quantise (levels replace the coefficients in the tile), dequantise, inverse
        // transform, reconstruct against the prediction still in the pixel tile
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const int j = q * 32 + lane;
            uint4* c32 = i32_of(j);
            int res[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 c = c32[i];
                const int cc[4] = {(int)c.x, (int)c.y, (int)c.z, (int)c.w};
                int l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    l[k] = quantize_fast(cc[k], fq);
                    res[i][k] = dequantize_fast(l[k], fq);
                }
                c32[i] = make_uint4(l[0], l[1], l[2], l[3]);
            }
            transform2d<N, DST, true>(res);
            uint4* p16 = px_of(px, j);
            const uint4 v0 = p16[0], v1 = p16[1];
            const uint32_t pw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint32_t rw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                rw[k] = pack16(recon_px(lo16(pw[k]), res[k >> 1][2 * (k & 1)], a.maxv),
                               recon_px(hi16(pw[k]), res[k >> 1][2 * (k & 1) + 1], a.maxv));
            p16[0] = make_uint4(rw[0], rw[1], rw[2], rw[3]);
            p16[1] = make_uint4(rw[4], rw[5], rw[6], rw[7]);
        }
        __syncwarp();
        if (a.levels) T32::store(s32, reinterpret_cast<unsigned char*>(a.levels + blk0 * NN), lane, chunks32);
        if (a.recon) T16::store(px, reinterpret_cast<unsigned char*>(a.recon + blk0 * NN), lane, chunks16);
        __syncwarp();
        // -- blocks whose inputs left the pixel domain are recoded exactly (cold path; the __syncwarp
        //    above orders the cooperative stores before these stores)
        if (ood_mask) {
            const FusedArgs a_cold = a;
            for (int q = 0; q < 4; ++q) {
                const int j = q * 32 + lane;
                if (((ood_mask >> q) & 1u) && j < blocks_valid) slow_block<N>(a_cold, blk0 + j, DST);
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    release_tile_counter(tile_counter, gridDim.x * kV2Warps);
}

template <bool DST>
static int launch_unit4(const FusedArgs& a, cudaStream_t st) {
    constexpr int kSmem = kV2Warps * (2 * WarpTile<128>::kBytes + WarpTile<256>::kBytes);
    {
        const int rc = ensure_dynamic_smem(fused_unit4_kernel<DST>, kSmem, "cudaFuncSetAttribute(fused_unit4_kernel)");
        if (rc != NH_OK) return rc;
    }
    int grid = grid_for((a.n_blocks + 3) / 4, (int64_t)kV2Warps * 32, 3);
    unsigned int* counter = nullptr;
    {
        const int rc = acquire_tile_counter(st, &counter);
        if (rc != NH_OK) return rc;
    }
    fused_unit4_kernel<DST><<<grid, kV2Warps * 32, kSmem, st>>>(a, make_fast_quant(a.qp), counter);
    NH_CHECK_LAUNCH("fused_unit4_kernel");
    tile_counter_launched(st);
    return NH_OK;
}

// ------------------------------------------------------- unit kernels, v3 (TMA)
// Generation 2 with the staging sweeps replaced by tensor-map TMA copies issued by ONE elected lane:
//   * the next pixel tile arrives through cp.async.bulk.tensor (UTMALDG) into a 128B-swizzled shared
//     tile and is awaited on an mbarrier;
//   * pred / coeff / levels / recon leave through cp.async.bulk.tensor stores (UTMASTG) straight from
//     the swizzled tiles the lanes wrote, tracked by the elected lane's bulk groups.
// Tensors are viewed as [units][64 elements]; a box is 32 units x 128 bytes (int32 tensors: two boxes
// per tile, elements 0-31 and 32-63 of every unit).  With SWIZZLE_128B the 16-byte chunk k of row t
// sits at t*128 + ((k ^ (t & 7)) << 4): a lane's 128-bit accesses to its own row are conflict-free
// and the tile in shared memory is contiguous, so one instruction moves it.
struct TmaMaps {
    CUtensorMap orig, pred, coeff, levels, recon;
};

template <int N, bool DST>
__global__ void __launch_bounds__(kV2Warps * 32, 3)
    fused_unit_kernel_v3(const FusedArgs a, const FastQuant fq, const __grid_constant__ TmaMaps maps,
                         const int64_t n_units) {
    constexpr int NN = N * N;
    constexpr int BPU = 64 / NN;
    constexpr int kTile = 4096;                         // 32 rows x 128 bytes
    constexpr int kWarpBytes = 4 * kTile;               // 2 pixel tiles + 2 int32 half tiles
    extern __shared__ unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1024-byte alignment is required by the 128B swizzle pattern
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t wbase = base + warp * kWarpBytes;
    const uint32_t s16[2] = {wbase, wbase + kTile};
    const uint32_t s32h[2] = {wbase + 2 * kTile, wbase + 3 * kTile};
    const uint32_t mbar[2] = {base + kV2Warps * kWarpBytes + warp * 16, base + kV2Warps * kWarpBytes + warp * 16 + 8};
    // this lane's row in a tile, and its swizzled chunk addresses
    const uint32_t row = lane * 128;
    const uint32_t sw = lane & 7;
    auto chunk = [&](uint32_t tile, int k) -> uint32_t { return tile + row + (((uint32_t)k ^ sw) << 4); };
    auto lds128 = [](uint32_t addr) -> uint4 {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
        return v;
    };
    auto sts128 = [](uint32_t addr, uint4 v) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    };

    if (lane == 0) {
        mbar_init(mbar[0], 1);
        mbar_init(mbar[1], 1);
        fence_mbar_init();
    }
    fence_async_smem();
    __syncwarp();

    const int64_t n_tiles = (n_units + 31) / 32;
    const int64_t warp_stride = (int64_t)gridDim.x * kV2Warps;
    int64_t tile = (int64_t)blockIdx.x * kV2Warps + warp;

    struct Refs {
        uint32_t tw[BPU][N / 2], lw[BPU][N / 2];
        int tr[BPU], bl[BPU], mode[BPU];
    };
    auto load_refs = [&](int64_t t, Refs& r) {
        const int64_t unit = t * 32 + lane;
        const int64_t ub = unit * BPU;
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
            r.mode[q] = a.mode;
            if (unit < n_units) {
                load_row16<N>(a.top + (ub + q) * N, r.tw[q]);
                load_row16<N>(a.left + (ub + q) * N, r.lw[q]);
                r.tr[q] = a.top_right[ub + q];
                r.bl[q] = a.bottom_left[ub + q];
                if (a.modes) r.mode[q] = a.modes[ub + q];
            } else {
#pragma unroll
                for (int k = 0; k < N / 2; ++k) r.tw[q][k] = r.lw[q][k] = 0;
                r.tr[q] = r.bl[q] = 0;
            }
        }
    };
    auto issue_load = [&](int64_t t, int buf) {  // elected lane only
        mbar_arrive_expect_tx(mbar[buf], kTile);
        tma_load_2d(s16[buf], &maps.orig, 0, (int)(t * 32), mbar[buf]);
    };

    Refs nxt;
    if (tile < n_tiles) {
        if (lane == 0) issue_load(tile, 0);
        load_refs(tile, nxt);
    }
    int cur = 0;
    uint32_t it = 0;  // iteration counter: buffer `cur` is on its (it >> 1)-th use
    for (; tile < n_tiles; tile += warp_stride, cur ^= 1, ++it) {
        const int unit0 = (int)(tile * 32);
        const int64_t unit = tile * 32 + lane;
        const int64_t ub = unit * BPU;
        const bool uvalid = unit < n_units;

        uint32_t tw[BPU][N / 2], lw[BPU][N / 2];
        int tr[BPU], bl[BPU], mode[BPU];
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
#pragma unroll
            for (int k = 0; k < N / 2; ++k) { tw[q][k] = nxt.tw[q][k]; lw[q][k] = nxt.lw[q][k]; }
            tr[q] = nxt.tr[q]; bl[q] = nxt.bl[q]; mode[q] = nxt.mode[q];
        }
        if (tile + warp_stride < n_tiles) load_refs(tile + warp_stride, nxt);

        // bulk groups of the previous iteration: everything but its reconstruction has been read
        if (lane == 0) bulk_wait_read<1>();
        mbar_wait(mbar[cur], (it >> 1) & 1);   // this tile's pixels have landed
        __syncwarp();

        int res[BPU][N][N];
        uint32_t ood = 0;
#pragma unroll
        for (int q = 0; q < BPU; ++q) {
            int top[N], left[N];
            unpack_row<N>(tw[q], top);
            unpack_row<N>(lw[q], left);
#pragma unroll
            for (int k = 0; k < N / 2; ++k) ood |= (tw[q][k] | lw[q][k]) & 0xF000F000u;
            ood |= (uint32_t)(tr[q] | bl[q]) & 0xFFFFF000u;
            int* rq = &res[q][0][0];
            int dc = 0;
            if (mode[q] == 1) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) s += top[k] + left[k];
                dc = dc_value<N>(s);
            }
#pragma unroll
            for (int c = 0; c < NN / 8; ++c) {
                const uint32_t addr = chunk(s16[cur], q * (NN / 8) + c);
                const uint4 v = lds128(addr);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                int p[8];
                if (mode[q] == 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) p[k] = dc;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int e = 8 * c + k, y = e / N, x = e % N;
                        p[k] = planar_px<N>(x, y, left[y], top[x], tr[q], bl[q]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ood |= w[k] & 0xF000F000u;
                    rq[8 * c + 2 * k] = lo16(w[k]) - p[2 * k];
                    rq[8 * c + 2 * k + 1] = hi16(w[k]) - p[2 * k + 1];
                }
                sts128(addr, make_uint4(pack16(p[0], p[1]), pack16(p[2], p[3]), pack16(p[4], p[5]), pack16(p[6], p[7])));
            }
        }
        const bool fast = ood == 0 || !uvalid;
        // G1: prediction
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (a.pred) tma_store_2d(&maps.pred, 0, unit0, s16[cur]);
            bulk_commit();
        }

        // -- forward transform (in-thread, both passes)
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, false>(res[q]);
        int* flat = &res[0][0][0];
        // G2, G3: coefficient halves
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (a.coeff) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    sts128(chunk(s32h[h], e), make_uint4(flat[32 * h + 4 * e], flat[32 * h + 4 * e + 1],
                                                         flat[32 * h + 4 * e + 2], flat[32 * h + 4 * e + 3]));
                fence_async_smem();
            }
            __syncwarp();
            if (lane == 0) {
                if (a.coeff) tma_store_2d(&maps.coeff, 32 * h, unit0, s32h[h]);
                bulk_commit();
            }
        }
        // the previous reconstruction (3 groups back) has been read: its tile takes the prefetch
        if (lane == 0) {
            bulk_wait_read<3>();
            if (tile + warp_stride < n_tiles) {
                fence_async_smem();
                issue_load(tile + warp_stride, cur ^ 1);
            }
        }
        // -- quantise (levels out) and dequantise in place; G4, G5: level halves
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (lane == 0) bulk_wait_read<1>();  // the coefficient half that used this buffer has been read
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int* r = flat + 32 * h + 4 * e;
                int l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    l[k] = quantize_fast(r[k], fq);
                    r[k] = dequantize_fast(l[k], fq);
                }
                if (a.levels) sts128(chunk(s32h[h], e), make_uint4(l[0], l[1], l[2], l[3]));
            }
            if (a.levels) fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (a.levels) tma_store_2d(&maps.levels, 32 * h, unit0, s32h[h]);
                bulk_commit();
            }
        }
        // -- inverse transform, reconstruct against the prediction still in the pixel tile
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, true>(res[q]);
        // G1 (prediction) was forced complete by the waits above, the tile may be rewritten
        if (a.recon) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t addr = chunk(s16[cur], c);
                const uint4 pv = lds128(addr);
                const int* r = flat + 8 * c;
                uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w}, ow[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ow[k] = pack16(recon_px(lo16(pw[k]), r[2 * k], a.maxv),
                                   recon_px(hi16(pw[k]), r[2 * k + 1], a.maxv));
                sts128(addr, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            }
            fence_async_smem();
        }
        __syncwarp();
        // G6: reconstruction
        if (lane == 0) {
            if (a.recon) tma_store_2d(&maps.recon, 0, unit0, s16[cur]);
            bulk_commit();
        }
        // -- a lane whose inputs left the pixel domain recodes its unit exactly (cold path), after
        //    the tile's TMA stores have been performed
        if (__any_sync(0xffffffffu, !fast)) {
            if (lane == 0) {
                bulk_wait_all<0>();
                fence_async_smem();
            }
            __syncwarp();
            if (!fast) {
                const FusedArgs a_cold = a;
                for (int q = 0; q < BPU; ++q) slow_block<N>(a_cold, ub + q, DST);
            }
            __syncwarp();
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
}

// ------------------------------------------------------------ rows kernels
constexpr int kRowsWarps = 4;  // 128 threads per CTA

// Cold path of the rows kernels: the tile from the residual rows in M onwards with the plain int32
// butterflies and the int64 quant of the reference -- exact for any int16 input.
template <int N>
__device__ __noinline__ void rows_tile_exact(const FusedArgs& a, int* M, int r, bool valid, int64_t b,
                                             const uint32_t (&pw)[N / 2]) {
    constexpr int NN = N * N;
    {
        int c[N], lv[N], dq[N];
        two_pass_transform<N, false, false, false>(M, r, true, c);
        if (valid && a.coeff) store_row32<N>(a.coeff + b * NN + r * N, c);
        quant_dequant_row<N>(c, a.qp, lv, dq);
        if (valid && a.levels) store_row32<N>(a.levels + b * NN + r * N, lv);
        __syncwarp();
        store_row_smem<N>(M, r, dq);
    }
    __syncwarp();
    int res[N];
    two_pass_transform<N, false, true, false>(M, r, true, res);
    if (valid && a.recon) {
        uint32_t ow[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; ++k)
            ow[k] = pack16(recon_px(lo16(pw[k]), res[2 * k], a.maxv),
                           recon_px(hi16(pw[k]), res[2 * k + 1], a.maxv));
        store_row16<N>(a.recon + b * NN + r * N, ow);
    }
}

template <int N>
__global__ void __launch_bounds__(kRowsWarps * 32, 4) fused_rows_kernel(const FusedArgs a, const FastQuant fq) {
    constexpr int NN = N * N;
    constexpr int BPW = 32 / N;  // blocks per warp
    // IDP.2A odd part: measured +9 % at N = 32 and -3 % at N = 16 (the PRMT packing eats the gain
    // of the smaller odd part), so it is enabled for the 32-point butterfly only.
    constexpr bool kDP = N == 32;
    __shared__ __align__(16) int smem[kRowsWarps][BPW * RowsTile<N>::WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / N, r = lane % N;
    int* M = smem[warp] + g * RowsTile<N>::WORDS;

    const int64_t n_tiles = (a.n_blocks + BPW - 1) / BPW;
    const int64_t warp_stride = (int64_t)gridDim.x * kRowsWarps;
    int64_t tile = (int64_t)blockIdx.x * kRowsWarps + warp;

    // this lane's row of original pixels, software-pipelined one tile ahead
    uint32_t nxt_ow[N / 2];
    auto load_orig = [&](int64_t t) {
        const int64_t b = t * BPW + g;
        if (b < a.n_blocks) {
            load_row16<N>(a.orig + b * NN + r * N, nxt_ow);
        } else {
#pragma unroll
            for (int k = 0; k < N / 2; ++k) nxt_ow[k] = 0;
        }
    };
    if (tile < n_tiles) load_orig(tile);

    for (; tile < n_tiles; tile += warp_stride) {
        const int64_t b = tile * BPW + g;
        const bool valid = b < a.n_blocks;
        uint32_t pw[N / 2];  // prediction row, kept packed for the reconstruction
        uint32_t ood = 0;    // out-of-domain bits: any sample outside [0, 4095]
        {
            int o[N], p[N];
            unpack_row<N>(nxt_ow, o);
#pragma unroll
            for (int k = 0; k < N / 2; ++k) ood |= nxt_ow[k] & 0xF000F000u;
            if (tile + warp_stride < n_tiles) load_orig(tile + warp_stride);
            if (valid) {
                const int mode = a.modes ? (int)a.modes[b] : a.mode;
                uint32_t tw[N / 2], lw[N / 2];
                load_row16<N>(a.top + b * N, tw);
                load_row16<N>(a.left + b * N, lw);
                const int tr = a.top_right[b], bl = a.bottom_left[b];
#pragma unroll
                for (int k = 0; k < N / 2; ++k) ood |= (tw[k] | lw[k]) & 0xF000F000u;
                ood |= (uint32_t)(tr | bl) & 0xFFFFF000u;
                if (mode == 1) {
                    const int dc = dc_value<N>(sum_row<N>(tw) + sum_row<N>(lw));
#pragma unroll
                    for (int x = 0; x < N; ++x) p[x] = dc;
                } else {
                    int top[N];
                    unpack_row<N>(tw, top);
                    planar_row<N>(r, (int)a.left[b * N + r], top, tr, bl, p);
                }
            } else {
#pragma unroll
                for (int x = 0; x < N; ++x) p[x] = 0;
            }
            pack_row<N>(p, pw);
            if (valid && a.pred) store_row16<N>(a.pred + b * NN + r * N, pw);
            int res[N];
#pragma unroll
            for (int x = 0; x < N; ++x) res[x] = sext16(o[x] - sext16(p[x]));
            store_row_smem<N>(M, r, res);
        }
        // Pixel domain [0, 4095] for every sample this warp touched => 32-bit quant and the IDP.2A
        // butterflies are exact (bounds in DESIGN.md section 3); anything else takes the cold path.
        const bool fast = !__any_sync(0xffffffffu, ood != 0);
        __syncwarp();
        if (fast) {
            {
                int c[N], lv[N], dq[N];
                two_pass_transform<N, false, false, kDP>(M, r, true, c);
                if (valid && a.coeff16) {
                    uint32_t cw[N / 2];
                    pack_row<N>(c, cw);
                    store_row16<N>(a.coeff16 + b * NN + r * N, cw);
                } else if (valid && a.coeff) {
                    store_row32<N>(a.coeff + b * NN + r * N, c);
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    lv[k] = quantize_fast(c[k], fq);
                    dq[k] = dequantize_fast(lv[k], fq);
                }
                if (valid && a.levels16) {
                    uint32_t lw2[N / 2];
                    pack_row<N>(lv, lw2);
                    store_row16<N>(a.levels16 + b * NN + r * N, lw2);
                } else if (valid && a.levels) {
                    store_row32<N>(a.levels + b * NN + r * N, lv);
                }
                __syncwarp();  // every lane has read its column of the second forward pass
                store_row_smem<N>(M, r, dq);
            }
            __syncwarp();
            int res[N];
            two_pass_transform<N, false, true, kDP>(M, r, true, res);
            if (valid && a.recon) {
                uint32_t ow[N / 2];
#pragma unroll
                for (int k = 0; k < N / 2; ++k)
                    ow[k] = pack16(recon_px(lo16(pw[k]), res[2 * k], a.maxv),
                                   recon_px(hi16(pw[k]), res[2 * k + 1], a.maxv));
                store_row16<N>(a.recon + b * NN + r * N, ow);
            }
        } else {
            if (a.ood_flag && lane == 0) *a.ood_flag = 1;
            const FusedArgs a_cold = a;  // copies made here so that nothing has its address taken on the hot path
            uint32_t pw_cold[N / 2];
#pragma unroll
            for (int k = 0; k < N / 2; ++k) pw_cold[k] = pw[k];
            rows_tile_exact<N>(a_cold, M, r, valid, b, pw_cold);
        }
        __syncwarp();
    }
}

}  // namespace nh
#include "nh_fused_mma.cuh"
#include "nh_coder8.cuh"
namespace nh {

template <int N, bool DST>
static int launch_unit_v1(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPU = 64 / (N * N);
    constexpr int kSmem = kUnitWarps * (WarpTile<128>::kBytes + WarpTile<256>::kBytes);
    {
        const int rc = ensure_dynamic_smem(fused_unit_kernel<N, DST>, kSmem, "cudaFuncSetAttribute(fused_unit_kernel)");
        if (rc != NH_OK) return rc;
    }
    int64_t units = (a.n_blocks + BPU - 1) / BPU;
    int grid = grid_for(units, (int64_t)kUnitWarps * 32, 2);
    fused_unit_kernel<N, DST><<<grid, kUnitWarps * 32, kSmem, st>>>(a);
    NH_CHECK_LAUNCH("fused_unit_kernel");
    return NH_OK;
}

template <int N, bool DST, bool NARROW = false>
static int launch_unit_v2(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPU = 64 / (N * N);
    constexpr int kSmem = kV2Warps * (2 * WarpTile<128>::kBytes + WarpTile<256>::kBytes);
    {
        const int rc = ensure_dynamic_smem(fused_unit_kernel_v2<N, DST, NARROW>, kSmem, "cudaFuncSetAttribute(fused_unit_kernel_v2)");
        if (rc != NH_OK) return rc;
    }
    int64_t units = (a.n_blocks + BPU - 1) / BPU;
    int grid = grid_for(units, (int64_t)kV2Warps * 32, 3);
    fused_unit_kernel_v2<N, DST, NARROW><<<grid, kV2Warps * 32, kSmem, st>>>(a, make_fast_quant(a.qp));
    NH_CHECK_LAUNCH("fused_unit_kernel_v2");
    return NH_OK;
}

// ---- tensor maps for generation 3 --------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// [n_units][64 elements] view of a block-major tensor; box = 32 units x 128 bytes, 128B swizzle.
static int make_unit_map(CUtensorMap* m, const void* ptr, int64_t n_units, bool is32) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NH_E_CUDA; }
    const cuuint64_t gdim[2] = {64, (cuuint64_t)n_units};
    const cuuint64_t gstride[1] = {(cuuint64_t)(is32 ? 256 : 128)};
    const cuuint32_t box[2] = {(cuuint32_t)(is32 ? 32 : 64), 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, is32 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                    const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return NH_E_CUDA; }
    return NH_OK;
}

template <int N, bool DST>
static int launch_unit_v3(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPU = 64 / (N * N);
    constexpr int NN = N * N;
    constexpr int kSmem = kV2Warps * 4 * 4096 + kV2Warps * 16 + 1024;
    {
        const int rc = ensure_dynamic_smem(fused_unit_kernel_v3<N, DST>, kSmem, "cudaFuncSetAttribute(fused_unit_kernel_v3)");
        if (rc != NH_OK) return rc;
    }
    const int64_t n_units = a.n_blocks / BPU;  // whole units go through TMA
    if (n_units > 0) {
        TmaMaps maps;
        int rc = make_unit_map(&maps.orig, a.orig, n_units, false);
        // outputs that were not requested reuse the input map: the kernel never issues their stores
        if (rc == NH_OK) rc = make_unit_map(&maps.pred, a.pred ? (const void*)a.pred : (const void*)a.orig, n_units, false);
        if (rc == NH_OK) rc = make_unit_map(&maps.recon, a.recon ? (const void*)a.recon : (const void*)a.orig, n_units, false);
        if (rc == NH_OK) rc = a.coeff ? make_unit_map(&maps.coeff, a.coeff, n_units, true) : make_unit_map(&maps.coeff, a.orig, n_units, false);
        if (rc == NH_OK) rc = a.levels ? make_unit_map(&maps.levels, a.levels, n_units, true) : make_unit_map(&maps.levels, a.orig, n_units, false);
        if (rc != NH_OK) return rc;
        FusedArgs b = a;
        b.n_blocks = n_units * BPU;
        int grid = grid_for(n_units, (int64_t)kV2Warps * 32, 3);
        fused_unit_kernel_v3<N, DST><<<grid, kV2Warps * 32, kSmem, st>>>(b, make_fast_quant(a.qp), maps, n_units);
        NH_CHECK_LAUNCH("fused_unit_kernel_v3");
    }
    const int64_t done = n_units * BPU, tail = a.n_blocks - done;
    if (tail > 0) {  // fewer than BPU trailing 4x4 blocks: not a whole 128-byte TMA row
        FusedArgs t = a;
        t.orig += done * NN; t.top += done * N; t.left += done * N;
        t.top_right += done; t.bottom_left += done;
        if (t.modes) t.modes += done;
        if (t.pred) t.pred += done * NN;
        if (t.coeff) t.coeff += done * NN;
        if (t.levels) t.levels += done * NN;
        if (t.recon) t.recon += done * NN;
        t.n_blocks = tail;
        return launch_unit_v2<N, DST>(t, st);
    }
    return NH_OK;
}

// Kernel generation used for N = 4, 8: 4 (default: tensor-core passes at N = 8, the rolled 4x4
// kernel at N = 4), 2 (cp.async + in-thread butterflies), 1 (first generation) or 3 (TMA tensor-map staging);
// set by nh_set_fused_impl() or the NH_FUSED_IMPL=v1|v2|v3|v4 environment variable.
static thread_local int g_fused_impl = 0;   // per calling thread: no shared mutable state between callers
static int fused_impl() {
    if (g_fused_impl == 0) {
        const char* e = getenv("NH_FUSED_IMPL");
        g_fused_impl = (e && e[0] == 'v' && e[1] >= '1' && e[1] <= '4') ? e[1] - '0' : 4;
    }
    return g_fused_impl;
}

template <int N, bool DST>
static int launch_unit(const FusedArgs& a, cudaStream_t st) {
    if (a.coeff16) return launch_unit_v2<N, DST, true>(a, st);  // narrow outputs: generation 2 only
    if (N == 8 && fused_impl() == 4) {
        // tensor-core passes for 8-bit content; deeper content goes to generation 2, whose 32-bit fast
        // path covers samples up to 4095 (the tensor-core kernel would recode every tile exactly)
        if (a.maxv <= 255) return launch_mma8(a, st);
        return launch_unit_v2<N, DST>(a, st);
    }
    if (N == 4 && fused_impl() == 4) return launch_unit4<DST>(a, st);  // rolled 4x4 kernel
    switch (fused_impl()) {
        case 1: return launch_unit_v1<N, DST>(a, st);
        case 3: return launch_unit_v3<N, DST>(a, st);
        default: return launch_unit_v2<N, DST>(a, st);
    }
}

template <int N>
static int launch_rows(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPW = 32 / N;
    int grid = grid_for(a.n_blocks, (int64_t)kRowsWarps * BPW, 4);
    fused_rows_kernel<N><<<grid, kRowsWarps * 32, 0, st>>>(a, make_fast_quant(a.qp));
    NH_CHECK_LAUNCH("fused_rows_kernel");
    return NH_OK;
}

// Kernel behind N = 16, 32: 2 (default) = tensor-core passes (fused_mma_kernel), 1 = CUDA-core
// butterflies (fused_rows_kernel); nh_set_rows_impl() or NH_ROWS_IMPL=1|2.  The int16-output variant
// of the host pipeline always uses the CUDA-core kernel.
static thread_local int g_rows_impl = 0;
int rows_impl() {
    if (g_rows_impl == 0) {
        const char* e = getenv("NH_ROWS_IMPL");
        g_rows_impl = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 2;
    }
    return g_rows_impl;
}

template <int N>
static int launch_16_32(const FusedArgs& a, cudaStream_t st) {
    // tensor-core kernel for 8-bit content; the CUDA-core kernel keeps a fast path up to 12 bits
    if (a.coeff16 || rows_impl() == 1 || a.maxv > 255) return launch_rows<N>(a, st);
    return launch_mma<N>(a, st);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int dispatch_fused(const FusedArgs& a, int size, int use_dst, cudaStream_t st) {
    switch (size) {
        case 4: return use_dst ? launch_unit<4, true>(a, st) : launch_unit<4, false>(a, st);
        case 8: return launch_unit<8, false>(a, st);
        case 16: return launch_16_32<16>(a, st);
        default: return launch_16_32<32>(a, st);
    }
}

}  // namespace nh

NH_API int nh_set_fused_impl(int generation) {
    if (generation < 1 || generation > 4) {
        nh::set_error("nh_set_fused_impl: generation must be 1, 2, 3 or 4, got %d", generation);
        return NH_E_ARG;
    }
    nh::g_fused_impl = generation;
    return NH_OK;
}

NH_API int nh_set_rows_impl(int impl) {
    if (impl < 1 || impl > 2) {
        nh::set_error("nh_set_rows_impl: impl must be 1 (CUDA-core) or 2 (tensor-core), got %d", impl);
        return NH_E_ARG;
    }
    nh::g_rows_impl = impl;
    return NH_OK;
}

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
                pred, coeff, levels, recon, nullptr, nullptr, nullptr};
    return nh::dispatch_fused(a, size, use_dst, reinterpret_cast<cudaStream_t>(stream));
}

// Internal entry of the host-buffer pipeline (nh_host.cu): same kernel, coefficients / levels
// narrowed to int16 on the device side of the PCIe copy; *ood_flag is raised when a block left the
// pixel domain (the narrow tensors are then not valid for that chunk).
int nh::fused_pipeline_dcplanar_narrow(const int16_t* orig, const int16_t* top, const int16_t* left,
                                       const int16_t* top_right, const int16_t* bottom_left,
                                       const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                                       int is_intra, int use_dst, int bit_depth, int16_t* pred,
                                       int16_t* coeff16, int16_t* levels16, int16_t* recon, int* ood_flag,
                                       cudaStream_t st) {
    const int l2 = log2_size(size);
    if (l2 < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (n_blocks <= 0) return NH_OK;
    FusedArgs a{orig, top, left, top_right, bottom_left, modes, mode, n_blocks,
                make_quant_params(qp, l2, is_intra), (1 << bit_depth) - 1,
                pred, nullptr, nullptr, recon, coeff16, levels16, ood_flag};
    return dispatch_fused(a, size, use_dst, st);
}
