// nh_ops.cu -- the batched single-stage operators behind the reference's
// per-block functions: K2 predictors, K3 residual / reconstruct / clip,
// K4 forward / inverse transforms, K5 quantize / dequantize.
#include <cstdlib>

#include "nh_block.cuh"
#include "nh_mma.cuh"

namespace nh {

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// =================================================================== K4
// N = 4, 8: one lane owns 64 coefficients (one 8x8 block / four 4x4 blocks),
// staged through padded shared memory for fully coalesced 128-bit traffic.
constexpr int kXfWarps = 8;

template <int N, bool DST, bool INV, bool IN32>
__global__ void __launch_bounds__(kXfWarps * 32, 2)
    transform_unit_kernel(const void* __restrict__ in, int32_t* __restrict__ out, int64_t n_blocks) {
    constexpr int NN = N * N;
    constexpr int BPU = 64 / NN;
    using TIn = WarpTile<IN32 ? 256 : 128>;
    using TOut = WarpTile<256>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* s = smem_raw + warp * TOut::kBytes;  // one buffer, reused for input and output

    // Programmatic dependent launch (launch_transform_unit): the next launch on the stream may be scheduled while this
    // grid drains, and this grid touches global memory only after the grid in front of it has completed and flushed.
    // Both instructions are no-ops in a launch without the attribute.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int64_t n_units = (n_blocks + BPU - 1) / BPU;
    const int64_t n_tiles = (n_units + 31) / 32;
    const int cta_warps = (int)(blockDim.x >> 5);   // 5 .. kXfWarps: picked per launch (launch_transform_unit)
    for (int64_t tile = (int64_t)blockIdx.x * cta_warps + warp; tile < n_tiles;
         tile += (int64_t)gridDim.x * cta_warps) {
        const int64_t blk0 = tile * 32 * BPU;
        int64_t rem = n_blocks - blk0;
        const int blocks_valid = (int)(rem < 32 * BPU ? rem : 32 * BPU);
        constexpr int kInElem = IN32 ? 4 : 2;
        TIn::load(s, reinterpret_cast<const unsigned char*>(in) + blk0 * NN * kInElem, lane,
                  blocks_valid * (NN * kInElem / 16));
        __syncwarp();
        int v[BPU][N][N];
        int* flat = &v[0][0][0];
        const uint4* ui = TIn::unit(s, lane);
        if constexpr (IN32) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                uint4 t = ui[e];
                flat[4 * e] = (int)t.x; flat[4 * e + 1] = (int)t.y;
                flat[4 * e + 2] = (int)t.z; flat[4 * e + 3] = (int)t.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                uint4 t = ui[e];
                flat[8 * e] = lo16(t.x); flat[8 * e + 1] = hi16(t.x);
                flat[8 * e + 2] = lo16(t.y); flat[8 * e + 3] = hi16(t.y);
                flat[8 * e + 4] = lo16(t.z); flat[8 * e + 5] = hi16(t.z);
                flat[8 * e + 6] = lo16(t.w); flat[8 * e + 7] = hi16(t.w);
            }
        }
        __syncwarp();  // everyone has consumed the input tile before it is overwritten
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, INV>(v[q]);
        uint4* uo = TOut::unit(s, lane);
#pragma unroll
        for (int e = 0; e < 16; ++e)
            uo[e] = make_uint4(flat[4 * e], flat[4 * e + 1], flat[4 * e + 2], flat[4 * e + 3]);
        __syncwarp();
        TOut::store(s, reinterpret_cast<unsigned char*>(out + blk0 * NN), lane,
                    blocks_valid * (NN * 4 / 16));
        __syncwarp();
    }
}

// The same kernel as a software pipeline (default; NH_XF_PIPE=0 keeps the one above): round 1 measured the
// unpipelined form at 0.65-0.87 of the HBM copy bandwidth (16 warps per SM, load -> transform -> store in
// sequence per warp, a static tile partition that waits for its slowest SM).  The recipe of the fused kernels
// applies unchanged: 4 warps per CTA, 3 CTAs per SM, two tile buffers per warp -- the next tile arrives by
// cp.async while this one is transformed in place and streamed out -- and tiles drawn from a ticket counter.
constexpr int kXfPipeWarps = 4;

template <int N, bool DST, bool INV, bool IN32>
__global__ void __launch_bounds__(kXfPipeWarps * 32, 3)
    transform_unit_pipe_kernel(const void* __restrict__ in, int32_t* __restrict__ out, int64_t n_blocks,
                               unsigned int* tile_counter) {
    constexpr int NN = N * N;
    constexpr int BPU = 64 / NN;
    using TIn = WarpTile<IN32 ? 256 : 128>;
    using TOut = WarpTile<256>;
    constexpr int kInElem = IN32 ? 4 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* buf[2] = {smem_raw + (2 * warp) * TOut::kBytes, smem_raw + (2 * warp + 1) * TOut::kBytes};

    const int64_t n_units = (n_blocks + BPU - 1) / BPU;
    const int64_t n_tiles = (n_units + 31) / 32;
    auto next_tile = [&]() -> int64_t {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        return (int64_t)__shfl_sync(0xffffffffu, t, 0);
    };
    auto prefetch = [&](int64_t t, unsigned char* dst) {   // global (linear) -> shared (padded units)
        const int64_t blk0 = t * 32 * BPU;
        const int64_t rem = n_blocks - blk0;
        const int chunks = (int)(rem < 32 * BPU ? rem : 32 * BPU) * (NN * kInElem / 16);
        const unsigned char* g = reinterpret_cast<const unsigned char*>(in) + blk0 * NN * kInElem;
#pragma unroll
        for (int it = 0; it < TIn::kIters; ++it) {
            const int c = it * 32 + lane;
            const int u = c / TIn::kChunksPerUnit, k = c % TIn::kChunksPerUnit;
            if (c < chunks) cp_async16(smem_u32(dst + u * TIn::kPitch + k * 16), g + (size_t)c * 16);
            else *reinterpret_cast<uint4*>(dst + u * TIn::kPitch + k * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    };
    int64_t tile = next_tile();
    int64_t tile_next = tile < n_tiles ? next_tile() : n_tiles;
    if (tile < n_tiles) prefetch(tile, buf[0]);
    cp_async_commit();
    int cur = 0;
    int64_t tile_after = n_tiles;
    for (; tile < n_tiles; tile = tile_next, tile_next = tile_after, cur ^= 1) {
        tile_after = tile_next < n_tiles ? next_tile() : n_tiles;
        const int64_t blk0 = tile * 32 * BPU;
        const int64_t rem = n_blocks - blk0;
        const int blocks_valid = (int)(rem < 32 * BPU ? rem : 32 * BPU);
        if (tile_next < n_tiles) prefetch(tile_next, buf[cur ^ 1]);   // its last reader finished before the syncwarp below
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        unsigned char* s = buf[cur];
        int v[BPU][N][N];
        int* flat = &v[0][0][0];
        const uint4* ui = TIn::unit(s, lane);
        if constexpr (IN32) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                uint4 t = ui[e];
                flat[4 * e] = (int)t.x; flat[4 * e + 1] = (int)t.y;
                flat[4 * e + 2] = (int)t.z; flat[4 * e + 3] = (int)t.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                uint4 t = ui[e];
                flat[8 * e] = lo16(t.x); flat[8 * e + 1] = hi16(t.x);
                flat[8 * e + 2] = lo16(t.y); flat[8 * e + 3] = hi16(t.y);
                flat[8 * e + 4] = lo16(t.z); flat[8 * e + 5] = hi16(t.z);
                flat[8 * e + 6] = lo16(t.w); flat[8 * e + 7] = hi16(t.w);
            }
        }
        __syncwarp();  // everyone has consumed the input tile before it is overwritten
#pragma unroll
        for (int q = 0; q < BPU; ++q) transform2d<N, DST, INV>(v[q]);
        uint4* uo = TOut::unit(s, lane);
#pragma unroll
        for (int e = 0; e < 16; ++e)
            uo[e] = make_uint4(flat[4 * e], flat[4 * e + 1], flat[4 * e + 2], flat[4 * e + 3]);
        __syncwarp();
        TOut::store(s, reinterpret_cast<unsigned char*>(out + blk0 * NN), lane, blocks_valid * (NN * 4 / 16));
        __syncwarp();
    }
    cp_async_wait<0>();
    release_tile_counter(tile_counter, gridDim.x * kXfPipeWarps);
}

// N = 16, 32: N lanes per block through the shared-memory working matrix.
constexpr int kRowsWarps = 4;

template <int N, bool INV, bool IN32>
__global__ void __launch_bounds__(kRowsWarps * 32)
    transform_rows_kernel(const void* __restrict__ in, int32_t* __restrict__ out, int64_t n_blocks) {
    constexpr int NN = N * N;
    constexpr int BPW = 32 / N;
    __shared__ __align__(16) int smem[kRowsWarps][BPW * RowsTile<N>::WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / N, r = lane % N;
    int* M = smem[warp] + g * RowsTile<N>::WORDS;
    const int64_t n_tiles = (n_blocks + BPW - 1) / BPW;
    for (int64_t tile = (int64_t)blockIdx.x * kRowsWarps + warp; tile < n_tiles;
         tile += (int64_t)gridDim.x * kRowsWarps) {
        const int64_t b = tile * BPW + g;
        const bool valid = b < n_blocks;
        int x[N];
        if (valid) {
            if constexpr (IN32) {
                load_row32<N>(reinterpret_cast<const int32_t*>(in) + b * NN + r * N, x);
            } else {
                uint32_t w[N / 2];
                load_row16<N>(reinterpret_cast<const int16_t*>(in) + b * NN + r * N, w);
                unpack_row<N>(w, x);
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) x[k] = 0;
        }
        store_row_smem<N>(M, r, x);
        __syncwarp();
        int y[N];
        two_pass_transform<N, false, INV>(M, r, true, y);  // one inlined butterfly for both passes
        if (valid) store_row32<N>(out + b * NN + r * N, y);
        __syncwarp();
    }
}

// N = 16, 32 on the tensor cores: the two passes of ONE direction as register-chained HMMAs (see
// nh_fused_mma.cuh for the fragment algebra).  pass 1: acc(m, n) = A1(m, k) X(k, n) with the constant
// operand A1 = T (forward) or T^T (inverse) and X through ldmatrix.trans; pass 2: the rounded
// accumulators are the A operand, B = T^T (forward) or T (inverse) comes from the same constant
// table, and the result lands in the natural row layout.  Exact while every input sample lies in
// [-1024, 1023]: |pass-1 result| <= 2048 (an exact f16 integer) and accumulators <= 4.2e6 < 2^24;
// a warp tile with anything larger takes the CUDA-core butterflies.

// int16 pair, each in [-1024, 1023] -> f16 pair, exactly: x + 1024 = 1024 b + l; (0x6400 | l) is the f16
// number 1024 + l, from which 1024 (b = 1) or 2048 (b = 0) is subtracted.
__device__ __forceinline__ uint32_t s16x2_to_h2(uint32_t w) {
    const uint32_t v = __vadd2(w, 0x04000400u);
    const uint32_t h = (v & 0x03ff03ffu) | 0x64006400u;
    const uint32_t c = 0x68006800u - (v & 0x04000400u);
    return h2_bits(__hsub2(bits_h2(h), bits_h2(c)));
}

template <int N, bool INV, bool IN32, int OCC>
__global__ void __launch_bounds__(kMmaWarps * 32, OCC)
    transform_mma_kernel(const void* __restrict__ in, int32_t* __restrict__ out, int64_t n_blocks) {
    using C = MmaConsts<N>;
    constexpr int NN = N * N;
    constexpr int BPW = 32 / N;
    constexpr int MT = N / 16, NT = N / 8, KT = N / 16;
    constexpr int SH = Log2<N>::v + 5;
    constexpr int PITCH = N * 2 + 16;
    constexpr int TILE = N * PITCH;
    constexpr int FAST_BYTES = BPW * TILE;
    constexpr int EXACT_BYTES = BPW * RowsTile<N>::WORDS * 4;
    constexpr int WARP_BYTES = FAST_BYTES > EXACT_BYTES ? FAST_BYTES : EXACT_BYTES;
    constexpr int RW = IN32 ? N : N / 2;  // 32-bit words of one input row
    __shared__ __align__(16) unsigned char smem[kMmaWarps][WARP_BYTES];
    __shared__ __align__(16) uint4 ctab[C::V_END][32];
    stage_mma_consts<N, kMmaWarps * 32>(&ctab[0][0]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / N, r = lane % N;
    const int fg = lane >> 2, ft = lane & 3;
    unsigned char* sm = smem[warp];
    int* M = reinterpret_cast<int*>(sm) + g * RowsTile<N>::WORDS;
    const int lane_off = (((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + (lane >> 4) * 16;
    const uint32_t ctab_lane = smem_u32(&ctab[0][lane]);
    auto cv = [&](int v) -> uint4 { return ld_const_vec(ctab_lane, v); };
    const float rnd = (float)(1 << (SH - 1));

    const int64_t n_tiles = (n_blocks + BPW - 1) / BPW;
    const int64_t warp_stride = (int64_t)gridDim.x * kMmaWarps;
    int64_t tile = (int64_t)blockIdx.x * kMmaWarps + warp;
    // The warp tile (BPW consecutive blocks) is contiguous in memory: lanes sweep it in 16-byte chunks
    // (512 contiguous bytes per warp instruction), one tile ahead, and scatter the chunks into the
    // int16 shared tile.  (Row-per-lane loads touched 32 different lines per instruction and left the
    // kernel waiting on memory: long-scoreboard 9.6 warps per issue.)
    constexpr int ESZ = IN32 ? 4 : 2;
    constexpr int EPC = 16 / ESZ;                    // elements per chunk
    constexpr int CPL = BPW * NN * ESZ / 16 / 32;    // chunks per lane
    static_assert(CPL * 4 == RW, "chunk sweep and row sweep move the same number of words");
    uint4 nxt[CPL];
    auto prefetch = [&](int64_t t) {
        const int64_t e_valid = (n_blocks - t * BPW) * NN;  // elements of the tile that exist
        const unsigned char* p = reinterpret_cast<const unsigned char*>(in) + t * BPW * NN * ESZ;
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
            const int c = it * 32 + lane;
            nxt[it] = (int64_t)c * EPC < e_valid ? ldg_stream(p + 16 * c) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    if (tile < n_tiles) prefetch(tile);

    for (; tile < n_tiles; tile += warp_stride) {
        const int64_t b = tile * BPW + g;
        const bool valid = b < n_blocks;
        // test -1024 <= x <= 1023 on the chunks as they are
        uint32_t ood = 0;
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
            const uint32_t w[4] = {nxt[it].x, nxt[it].y, nxt[it].z, nxt[it].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) ood |= IN32 ? (w[k] + 1024u) & 0xFFFFF800u : __vadd2(w[k], 0x04000400u) & 0xF800F800u;
        }
        const bool fast = !__any_sync(0xffffffffu, ood != 0);
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
            const int e0 = (it * 32 + lane) * EPC;  // first element of the chunk within the tile
            const int u = e0 / NN, row = (e0 % NN) / N, col = e0 % N;
            const uint32_t w[4] = {nxt[it].x, nxt[it].y, nxt[it].z, nxt[it].w};
            if (fast) {
                unsigned char* dst = sm + u * TILE + row * PITCH + col * 2;
                if constexpr (IN32) *reinterpret_cast<uint2*>(dst) = make_uint2(pack16((int)w[0], (int)w[1]), pack16((int)w[2], (int)w[3]));
                else *reinterpret_cast<uint4*>(dst) = nxt[it];
            } else {  // exact path: int32 working matrix of the block
                int* mrow = reinterpret_cast<int*>(sm) + u * RowsTile<N>::WORDS + row * RowsTile<N>::PITCH + col;
                if constexpr (IN32) {
                    *reinterpret_cast<int4*>(mrow) = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
                } else {
                    *reinterpret_cast<int4*>(mrow) = make_int4(lo16(w[0]), hi16(w[0]), lo16(w[1]), hi16(w[1]));
                    *reinterpret_cast<int4*>(mrow + 4) = make_int4(lo16(w[2]), hi16(w[2]), lo16(w[3]), hi16(w[3]));
                }
            }
        }
        if (tile + warp_stride < n_tiles) prefetch(tile + warp_stride);
        __syncwarp();
        if (fast) {
#pragma unroll
            for (int u = 0; u < BPW; ++u) {
                const int64_t bu = tile * BPW + u;
                const bool valid_u = bu < n_blocks;
                const uint32_t so = smem_u32(sm + u * TILE) + lane_off;
                float acc[MT][NT][4];
                uint32_t h[MT][NT][2];
                {
                    uint32_t xb[KT][NT][2];
#pragma unroll
                    for (int ki = 0; ki < KT; ++ki)
#pragma unroll
                        for (int np = 0; np < NT / 2; ++np) {
                            uint32_t ro[4];
                            ldsm_x4_t(ro, so + 16 * ki * PITCH + 32 * np);
#pragma unroll
                            for (int j = 0; j < 4; ++j) xb[ki][2 * np + (j >> 1)][j & 1] = s16x2_to_h2(ro[j]);
                        }
#pragma unroll
                    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                        for (int ki = 0; ki < KT; ++ki) {
                            uint4 af;
                            if constexpr (INV) {  // A = T^T from the B = T fragments
                                const uint4 t = cv(C::V_TB + ki * NT / 2 + mi);
                                af = make_uint4(t.x, t.z, t.y, t.w);
                            } else {
                                af = cv(C::V_TA + mi * KT + ki);
                            }
#pragma unroll
                            for (int ni = 0; ni < NT; ++ni) {
                                if (ki == 0)
                                    hmma16816(acc[mi][ni], af, xb[ki][ni][0], xb[ki][ni][1], rnd, rnd, rnd, rnd);
                                else
                                    hmma16816(acc[mi][ni], af, xb[ki][ni][0], xb[ki][ni][1], acc[mi][ni][0],
                                              acc[mi][ni][1], acc[mi][ni][2], acc[mi][ni][3]);
                            }
                        }
                }
#pragma unroll
                for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NT; ++ni) {
                        h[mi][ni][0] = round_pair_plain<SH>(acc[mi][ni][0], acc[mi][ni][1]);
                        h[mi][ni][1] = round_pair_plain<SH>(acc[mi][ni][2], acc[mi][ni][3]);
                    }
                // second pass: A = first-pass result (C -> A), B = T^T (forward) / T (inverse)
#pragma unroll
                for (int np = 0; np < NT / 2; ++np)
#pragma unroll
                    for (int ki = 0; ki < KT; ++ki) {
                        uint32_t b00, b01, b10, b11;  // (b0, b1) of n-tiles 2np and 2np + 1
                        if constexpr (INV) {
                            const uint4 t = cv(C::V_TB + ki * NT / 2 + np);
                            b00 = t.x; b01 = t.y; b10 = t.z; b11 = t.w;
                        } else {  // B[k][n] = T[n][k]: the A = T fragment of m-tile np
                            const uint4 t = cv(C::V_TA + np * KT + ki);
                            b00 = t.x; b01 = t.z; b10 = t.y; b11 = t.w;
                        }
#pragma unroll
                        for (int mi = 0; mi < MT; ++mi) {
                            const uint4 af = make_uint4(h[mi][2 * ki][0], h[mi][2 * ki][1], h[mi][2 * ki + 1][0],
                                                        h[mi][2 * ki + 1][1]);
                            float(&d0)[4] = acc[mi][2 * np];
                            float(&d1)[4] = acc[mi][2 * np + 1];
                            if (ki == 0) {
                                hmma16816(d0, af, b00, b01, rnd, rnd, rnd, rnd);
                                hmma16816(d1, af, b10, b11, rnd, rnd, rnd, rnd);
                            } else {
                                hmma16816(d0, af, b00, b01, d0[0], d0[1], d0[2], d0[3]);
                                hmma16816(d1, af, b10, b11, d1[0], d1[1], d1[2], d1[3]);
                            }
                        }
                    }
                if (valid_u) {
                    int32_t* op = out + bu * NN + fg * N + 2 * ft;
#pragma unroll
                    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                        for (int ni = 0; ni < NT; ++ni) {
                            int v[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                v[e] = __float_as_int(floor_shift_magic<SH>(acc[mi][ni][e])) - kMagicI;
                            __stcs(reinterpret_cast<int2*>(op + (16 * mi) * N + 8 * ni), make_int2(v[0], v[1]));
                            __stcs(reinterpret_cast<int2*>(op + (16 * mi + 8) * N + 8 * ni), make_int2(v[2], v[3]));
                        }
                }
            }
        } else {
            int y[N];
            two_pass_transform<N, false, INV>(M, r, true, y);
            if (valid) store_row32<N>(out + b * NN + r * N, y);
        }
        __syncwarp();
    }
}

template <int N, bool INV, bool IN32>
static int launch_transform_mma(const void* in, int32_t* out, int64_t n_blocks, cudaStream_t st) {
    static const int occ = [] {  // resident CTAs per SM the kernel is compiled for: NH_XF_OCC=3|4|5 (A/B runs)
        const char* e = getenv("NH_XF_OCC");
        // measured per direction (tools/time_xform_mma.py, 2^20 16x16 / 2^18 32x32 blocks): the forward transform (int16 in,
        // int32 out) runs at 0.82 / 0.89 / 0.97 of the copy bandwidth with 3 / 4 / 5 CTAs per SM at N = 16 (0.93 / 0.95 / 0.96 at
        // N = 32), the inverse (int32 in and out) at 0.90 / 1.00 / 0.91 (1.00 / 0.99 / 0.95)
        return (e && e[0] >= '3' && e[0] <= '5') ? e[0] - '0' : (INV ? 4 : 5);
    }();
    int grid = grid_for(n_blocks, kMmaWarps * (32 / N), occ);
    if (occ == 4) transform_mma_kernel<N, INV, IN32, 4><<<grid, kMmaWarps * 32, 0, st>>>(in, out, n_blocks);
    else if (occ == 5) transform_mma_kernel<N, INV, IN32, 5><<<grid, kMmaWarps * 32, 0, st>>>(in, out, n_blocks);
    else transform_mma_kernel<N, INV, IN32, 3><<<grid, kMmaWarps * 32, 0, st>>>(in, out, n_blocks);
    NH_CHECK_LAUNCH("transform_mma_kernel");
    return NH_OK;
}

template <int N, bool DST, bool INV, bool IN32>
static int launch_transform_unit(const void* in, int32_t* out, int64_t n_blocks, cudaStream_t st) {
    // Measured (tools/sweep_xform.py, profiles/r2_xform_sweep.jsonl): the pipelined kernel wins for the forward
    // transform once a launch holds enough tiles per warp (2^24 4x4 blocks: 0.88 -> 0.97 of the copy bandwidth,
    // 2^24 8x8 blocks: 0.92 -> 0.96); small launches are bounded by ramp-up and the launch gap and prefer the
    // larger grid of the kernel above, and the inverse (int32 in AND out: twice the staging per tile) is no
    // faster pipelined.  NH_XF_PIPE=0|1 forces one of them.
    static const int force = [] { const char* e = getenv("NH_XF_PIPE"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    const bool pipe = force >= 0 ? force == 1 : (!INV && (n_blocks * N * N) / 64 >= (int64_t(1) << 21));
    if (pipe) {
        constexpr int kSmemP = kXfPipeWarps * 2 * WarpTile<256>::kBytes;
        int rc = ensure_dynamic_smem(transform_unit_pipe_kernel<N, DST, INV, IN32>, kSmemP, "cudaFuncSetAttribute(transform_unit_pipe_kernel)");
        if (rc != NH_OK) return rc;
        constexpr int BPUP = 64 / (N * N);
        const int gridp = grid_for((n_blocks + BPUP - 1) / BPUP, kXfPipeWarps * 32, 3);
        unsigned int* counter = nullptr;
        rc = acquire_tile_counter(st, &counter);
        if (rc != NH_OK) return rc;
        transform_unit_pipe_kernel<N, DST, INV, IN32><<<gridp, kXfPipeWarps * 32, kSmemP, st>>>(in, out, n_blocks, counter);
        NH_CHECK_LAUNCH("transform_unit_pipe_kernel");
        tile_counter_launched(st);
        return NH_OK;
    }
    constexpr int kSmem = kXfWarps * WarpTile<256>::kBytes;
    {
        const int rc = ensure_dynamic_smem(transform_unit_kernel<N, DST, INV, IN32>, kSmem, "cudaFuncSetAttribute(transform_unit_kernel)");
        if (rc != NH_OK) return rc;
    }
    constexpr int BPU = 64 / (N * N);
    // Warps per CTA: a small launch hands every warp only a few tiles (2^20 4x4 blocks: 55.4 tiles per SM, 3.46 per warp
    // with 16 warps -- a quarter of the warps runs a fourth round while the rest idles).  Pick the CTA size whose rounds
    // waste the least: 14 warps per SM make that 3.95 -> 4 rounds.
    int cta_warps = kXfWarps;
    {
        const int64_t tiles = ((n_blocks + BPU - 1) / BPU + 31) / 32;
        const double per_sm = (double)tiles / sm_count();
        double best_cost = 1e30;
        for (int w = kXfWarps; w >= 5; --w) {
            const double rounds = per_sm / (2 * w);
            const double cost = (double)(int64_t)(rounds + 0.999999) * (2 * w);   // warp-rounds the SM spends
            if (cost < best_cost * 0.97) { best_cost = cost; cta_warps = w; }      // fewer warps only for a clear gain
        }
        if (per_sm > 400) cta_warps = kXfWarps;   // many rounds: the remainder does not matter
    }
    int grid = grid_for((n_blocks + BPU - 1) / BPU, cta_warps * 32, 2);
    // A 2^20-block launch of 4x4 blocks runs for 20 us: the launch gap and the ramp of the next grid are a tenth of it.
    // With programmatic stream serialization the next grid's CTAs are placed while this one drains (NH_XF_PDL=0: plain launch).
    static const bool pdl = [] { const char* e = getenv("NH_XF_PDL"); return !(e && e[0] == '0'); }();
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(cta_warps * 32);
        cfg.dynamicSmemBytes = kSmem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, transform_unit_kernel<N, DST, INV, IN32>, in, out, n_blocks);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(transform_unit_kernel)");
        return NH_OK;
    }
    transform_unit_kernel<N, DST, INV, IN32><<<grid, cta_warps * 32, kSmem, st>>>(in, out, n_blocks);
    NH_CHECK_LAUNCH("transform_unit_kernel");
    return NH_OK;
}

template <int N, bool INV, bool IN32>
static int launch_transform_rows(const void* in, int32_t* out, int64_t n_blocks, cudaStream_t st) {
    int grid = grid_for(n_blocks, kRowsWarps * (32 / N), 4);
    transform_rows_kernel<N, INV, IN32><<<grid, kRowsWarps * 32, 0, st>>>(in, out, n_blocks);
    NH_CHECK_LAUNCH("transform_rows_kernel");
    return NH_OK;
}

template <bool INV, bool IN32>
static int dispatch_transform(const void* in, int32_t* out, int64_t n_blocks, int size, int use_dst,
                              cudaStream_t st) {
    switch (size) {
        case 4:
            return use_dst ? launch_transform_unit<4, true, INV, IN32>(in, out, n_blocks, st)
                           : launch_transform_unit<4, false, INV, IN32>(in, out, n_blocks, st);
        case 8: return launch_transform_unit<8, false, INV, IN32>(in, out, n_blocks, st);
        case 16:
            return rows_impl() == 1 ? launch_transform_rows<16, INV, IN32>(in, out, n_blocks, st)
                                    : launch_transform_mma<16, INV, IN32>(in, out, n_blocks, st);
        case 32:
            return rows_impl() == 1 ? launch_transform_rows<32, INV, IN32>(in, out, n_blocks, st)
                                    : launch_transform_mma<32, INV, IN32>(in, out, n_blocks, st);
    }
    return NH_E_SIZE;
}

// =================================================================== K5 / K3
// Pure element-wise streams: 128-bit vector body plus a scalar tail.
// Four independent 128-bit loads per thread and trip keep enough bytes in flight; F::fast is the
// 32-bit form (exact while |x| <= 32768), F::exact the reference's int64 arithmetic.
template <class F>
__global__ void __launch_bounds__(256) map_i32_kernel(const int32_t* __restrict__ in,
                                                      int32_t* __restrict__ out, int64_t n, F f) {
    constexpr int U = 4;
    const int64_t n4 = n / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto apply = [&](uint4 v) -> uint4 {
        const int x[4] = {(int)v.x, (int)v.y, (int)v.z, (int)v.w};
        const uint32_t big = ((uint32_t)(x[0] + 32768) | (uint32_t)(x[1] + 32768) | (uint32_t)(x[2] + 32768) |
                              (uint32_t)(x[3] + 32768)) & 0xFFFF0000u;
        // x + 32768 in [0, 65535] for |x| <= 32767; x = 32768 itself is caught as "big" (bit 16)
        if (big == 0)
            return make_uint4((uint32_t)f.fast(x[0]), (uint32_t)f.fast(x[1]), (uint32_t)f.fast(x[2]), (uint32_t)f.fast(x[3]));
        return make_uint4((uint32_t)f.exact(x[0]), (uint32_t)f.exact(x[1]), (uint32_t)f.exact(x[2]), (uint32_t)f.exact(x[3]));
    };
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(in + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u) stg_stream(out + 4 * (i + u * stride), apply(v[u]));
    }
    for (; i < n4; i += stride) stg_stream(out + 4 * i, apply(ldg_stream(in + 4 * i)));
    for (int64_t k = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = f.exact(in[k]);
}

struct QuantF {
    QuantParams p;
    FastQuant fq;
    __device__ int fast(int c) const { return quantize_fast(c, fq); }
    __device__ int exact(int c) const { return quantize_one(c, p); }
};
struct DequantF {
    QuantParams p;
    FastQuant fq;
    __device__ int fast(int l) const { return dequantize_fast(l, fq); }
    __device__ int exact(int l) const { return dequantize_one(l, p); }
};

// kind 0: residual = orig - pred (intra.py:65-67); 1: clip (intra.py:75-78, b unused)
template <int KIND>
__global__ void __launch_bounds__(256) map_i16_kernel(const int16_t* __restrict__ a,
                                                      const int16_t* __restrict__ b,
                                                      int16_t* __restrict__ out, int64_t n, int maxv) {
    auto f = [&](int x, int y) -> int {
        if constexpr (KIND == 0) return sext16(x - y);
        else return clip_pixel(x, maxv);
    };
    const int64_t n8 = n / 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        uint4 va = ldg_stream(a + 8 * i);
        uint4 vb = KIND == 0 ? ldg_stream(b + 8 * i) : make_uint4(0, 0, 0, 0);
        uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w}, wo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            wo[k] = pack16(f(lo16(wa[k]), lo16(wb[k])), f(hi16(wa[k]), hi16(wb[k])));
        stg_stream(out + 8 * i, make_uint4(wo[0], wo[1], wo[2], wo[3]));
    }
    for (int64_t i = 8 * n8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int16_t)f((int)a[i], KIND == 0 ? (int)b[i] : 0);
}

// intra.py:70-72: pred (int16) + residual (int32 truncated to int16), int16 wrap-around.
__global__ void __launch_bounds__(256) reconstruct_kernel(const int16_t* __restrict__ pred,
                                                          const int32_t* __restrict__ res,
                                                          int16_t* __restrict__ out, int64_t n) {
    const int64_t n8 = n / 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        uint4 p = ldg_stream(pred + 8 * i);
        uint4 r0 = ldg_stream(res + 8 * i), r1 = ldg_stream(res + 8 * i + 4);
        uint32_t pw[4] = {p.x, p.y, p.z, p.w};
        int r[8] = {(int)r0.x, (int)r0.y, (int)r0.z, (int)r0.w, (int)r1.x, (int)r1.y, (int)r1.z, (int)r1.w};
        uint32_t wo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            wo[k] = pack16(sext16(lo16(pw[k]) + sext16(r[2 * k])), sext16(hi16(pw[k]) + sext16(r[2 * k + 1])));
        stg_stream(out + 8 * i, make_uint4(wo[0], wo[1], wo[2], wo[3]));
    }
    for (int64_t i = 8 * n8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int16_t)sext16((int)pred[i] + sext16(res[i]));
}

static int grid_elems(int64_t n, int per_thread) {
    return grid_for((n + per_thread - 1) / per_thread, 256, 8);
}

// =================================================================== K2
// One lane = one row of one block; N lanes per block.  Reference arrays are read
// straight from global memory (they are tiny next to the prediction written).
struct GlobalRef {  // accessor for nh::ref_at / angular_sample
    const int16_t* p;   // primary array (2N+1 entries)
    const int16_t* s;   // secondary array
    int c;              // top_left argument
    __device__ int pri(int k) const { return (int)__ldg(p + k); }
    __device__ int sec(int k) const { return (int)__ldg(s + k); }
    __device__ int corner() const { return c; }
};

// kind 0: DC from (B,N) refs; 1: planar from (B,N) refs + tr/bl; 2: modes from padded refs
template <int N, int KIND>
__global__ void __launch_bounds__(256)
    predict_rows_kernel(const int16_t* __restrict__ top, const int16_t* __restrict__ left,
                        const int16_t* __restrict__ aux0, const int16_t* __restrict__ aux1,
                        const uint8_t* __restrict__ modes, int mode, int allow_dc_planar,
                        int16_t* __restrict__ pred, int64_t n_blocks) {
    const int64_t total = n_blocks * N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t b = t / N;
        const int r = (int)(t % N);
        int p[N];
        if constexpr (KIND == 0 || KIND == 1) {
            const int16_t* tp = top + b * N;
            const int16_t* lp = left + b * N;
            if constexpr (KIND == 0) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) s += (int)__ldg(tp + k) + (int)__ldg(lp + k);
                int dc = dc_value<N>(s);
#pragma unroll
                for (int x = 0; x < N; ++x) p[x] = dc;
            } else {
                int tv[N];
#pragma unroll
                for (int k = 0; k < N; ++k) tv[k] = (int)__ldg(tp + k);
                planar_row<N>(r, (int)__ldg(lp + r), tv, (int)__ldg(aux0 + b), (int)__ldg(aux1 + b), p);
            }
        } else {
            const int16_t* tp = top + b * (2 * N + 1);
            const int16_t* lp = left + b * (2 * N + 1);
            const int m = modes ? (int)modes[b] : mode;
            if (m == 1 && allow_dc_planar) {
                int s = 0;
#pragma unroll
                for (int k = 1; k <= N; ++k) s += (int)__ldg(tp + k) + (int)__ldg(lp + k);
                int dc = dc_value<N>(s);
#pragma unroll
                for (int x = 0; x < N; ++x) p[x] = dc;
            } else if (m == 0 && allow_dc_planar) {
                int tv[N];
#pragma unroll
                for (int k = 0; k < N; ++k) tv[k] = (int)__ldg(tp + 1 + k);
                planar_row<N>(r, (int)__ldg(lp + 1 + r), tv, (int)__ldg(tp + N + 1),
                              (int)__ldg(lp + N + 1), p);
            } else {
                const AngleInfo ai = angle_info(m < 2 ? 2 : (m > 34 ? 34 : m));  // launcher validates
                GlobalRef ref;
                ref.c = (int)__ldg(aux0 + b);
                if (ai.vertical) { ref.p = tp; ref.s = lp; } else { ref.p = lp; ref.s = tp; }
#pragma unroll
                for (int x = 0; x < N; ++x)
                    p[x] = ai.vertical ? angular_sample(ref, ai, x, r) : angular_sample(ref, ai, r, x);
            }
        }
        int16_t* out = pred + b * (N * N) + r * N;
        uint32_t w[N / 2];
        pack_row<N>(p, w);
        store_row16<N>(out, w);
    }
}


// intra_dc_predict (intra.py:46-62) sums WHATEVER top / left hold (top.sum() + left.sum()), then divides by
// 2 * size: the (B, n_top) / (B, n_left) form for callers that pass arrays of another length than `size`.
template <int N>
__global__ void __launch_bounds__(256)
    predict_dc_ragged_kernel(const int16_t* __restrict__ top, int n_top, const int16_t* __restrict__ left, int n_left,
                             int16_t* __restrict__ pred, int64_t n_blocks) {
    const int64_t total = n_blocks * N;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / N;
        const int r = (int)(t % N);
        int s = 0;
        for (int k = 0; k < n_top; ++k) s += (int)__ldg(top + b * n_top + k);
        for (int k = 0; k < n_left; ++k) s += (int)__ldg(left + b * n_left + k);
        const int dc = dc_value<N>(s);
        int p[N];
#pragma unroll
        for (int x = 0; x < N; ++x) p[x] = dc;
        uint32_t w[N / 2];
        pack_row<N>(p, w);
        store_row16<N>(pred + b * (N * N) + r * N, w);
    }
}

template <int KIND>
static int launch_predict(const int16_t* top, const int16_t* left, const int16_t* aux0,
                          const int16_t* aux1, const uint8_t* modes, int mode, int allow,
                          int16_t* pred, int64_t n_blocks, int size, cudaStream_t st) {
    if (n_blocks == 0) return NH_OK;
    int grid = grid_for(n_blocks * size, 256, 8);
    switch (size) {
        case 4: predict_rows_kernel<4, KIND><<<grid, 256, 0, st>>>(top, left, aux0, aux1, modes, mode, allow, pred, n_blocks); break;
        case 8: predict_rows_kernel<8, KIND><<<grid, 256, 0, st>>>(top, left, aux0, aux1, modes, mode, allow, pred, n_blocks); break;
        case 16: predict_rows_kernel<16, KIND><<<grid, 256, 0, st>>>(top, left, aux0, aux1, modes, mode, allow, pred, n_blocks); break;
        case 32: predict_rows_kernel<32, KIND><<<grid, 256, 0, st>>>(top, left, aux0, aux1, modes, mode, allow, pred, n_blocks); break;
        default: set_error("Unsupported transform size: %d", size); return NH_E_SIZE;
    }
    NH_CHECK_LAUNCH("predict_rows_kernel");
    return NH_OK;
}


// =================================================================== frame containers
// uint8 <-> int16 sample conversion for the device-side frame containers (frame.py:45-51 reads
// uint8 planes, frame.py:107-111 / :172-178 write them back with numpy's astype(np.uint8), i.e.
// the low 8 bits of every int16 sample).  16 samples per thread, scalar tail.
template <bool TO_I16>
__global__ void __launch_bounds__(256) convert_kernel(const void* __restrict__ in, void* __restrict__ out, int64_t n) {
    const int64_t n16 = n / 16;
    const uint8_t* b = reinterpret_cast<const uint8_t*>(TO_I16 ? in : out);   // the uint8 side
    const int16_t* w = reinterpret_cast<const int16_t*>(TO_I16 ? out : in);   // the int16 side
    const bool vec = ((reinterpret_cast<uintptr_t>(b) & 15) | (reinterpret_cast<uintptr_t>(w) & 15)) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        for (int64_t i = t0; i < n16; i += stride) {
            if constexpr (TO_I16) {
                const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(in) + i);
                const uint32_t s[4] = {v.x, v.y, v.z, v.w};
                uint32_t o[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[2 * k] = __byte_perm(s[k], 0u, 0x4140);      // bytes 0, 1 zero-extended
                    o[2 * k + 1] = __byte_perm(s[k], 0u, 0x4342);  // bytes 2, 3
                }
                uint4* d = reinterpret_cast<uint4*>(out) + 2 * i;
                stg_stream(d, make_uint4(o[0], o[1], o[2], o[3]));
                stg_stream(d + 1, make_uint4(o[4], o[5], o[6], o[7]));
            } else {
                const uint4 a = ldg_stream(reinterpret_cast<const uint4*>(in) + 2 * i);
                const uint4 c = ldg_stream(reinterpret_cast<const uint4*>(in) + 2 * i + 1);
                stg_stream(reinterpret_cast<uint4*>(out) + i,
                           make_uint4(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420),
                                      __byte_perm(c.x, c.y, 0x6420), __byte_perm(c.z, c.w, 0x6420)));
            }
        }
    }
    for (int64_t i = (vec ? n16 * 16 : 0) + t0; i < n; i += stride) {
        if constexpr (TO_I16) reinterpret_cast<int16_t*>(out)[i] = (int16_t)reinterpret_cast<const uint8_t*>(in)[i];
        else reinterpret_cast<uint8_t*>(out)[i] = (uint8_t)reinterpret_cast<const int16_t*>(in)[i];
    }
}

}  // namespace nh

using namespace nh;

#define NH_REQUIRE_SIZE(size)                                      \
    if (log2_size(size) < 0) {                                     \
        set_error("Unsupported transform size: %d", (int)(size)); \
        return NH_E_SIZE;                                          \
    }
#define NH_REQUIRE(cond, msg)  \
    if (!(cond)) {             \
        set_error("%s", msg);  \
        return NH_E_ARG;       \
    }

NH_API int nh_forward_transform(const void* residual, int residual_is_i32, int32_t* coeff,
                                int64_t n_blocks, int size, int use_dst, void* stream) {
    NH_REQUIRE_SIZE(size);
    NH_REQUIRE(residual && coeff && n_blocks >= 0, "nh_forward_transform: null pointer or negative count");
    NH_REQUIRE(aligned16(residual) && aligned16(coeff), "nh_forward_transform: tensors must be 16-byte aligned");
    if (n_blocks == 0) return NH_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return residual_is_i32 ? dispatch_transform<false, true>(residual, coeff, n_blocks, size, use_dst, st)
                           : dispatch_transform<false, false>(residual, coeff, n_blocks, size, use_dst, st);
}

NH_API int nh_inverse_transform(const int32_t* coeff, int32_t* residual, int64_t n_blocks, int size,
                                int use_dst, void* stream) {
    NH_REQUIRE_SIZE(size);
    NH_REQUIRE(residual && coeff && n_blocks >= 0, "nh_inverse_transform: null pointer or negative count");
    NH_REQUIRE(aligned16(residual) && aligned16(coeff), "nh_inverse_transform: tensors must be 16-byte aligned");
    if (n_blocks == 0) return NH_OK;
    return dispatch_transform<true, true>(coeff, residual, n_blocks, size, use_dst,
                                          reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_quantize(const int32_t* coeff, int32_t* level, int64_t n, int qp, int size, int is_intra,
                       void* stream) {
    // quant.py:72 computes int(np.log2(size)) for whatever size it is given: 1..63 is accepted here (the
    // 32-bit fast path is exact up to shift 14 + 8 + 5)
    NH_REQUIRE(size >= 1 && size <= 63, "nh_quantize: size out of range 1..63");
    NH_REQUIRE(coeff && level && n >= 0, "nh_quantize: null pointer or negative count");
    NH_REQUIRE(aligned16(coeff) && aligned16(level), "nh_quantize: tensors must be 16-byte aligned");
    if (n == 0) return NH_OK;
    int l2 = 0;
    while ((2 << l2) <= size) ++l2;   // floor(log2(size))
    QuantF f{make_quant_params(qp, l2, is_intra), {}};
    f.fq = make_fast_quant(f.p);
    map_i32_kernel<<<grid_elems(n, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(coeff, level, n, f);
    NH_CHECK_LAUNCH("nh_quantize");
    return NH_OK;
}

NH_API int nh_dequantize(const int32_t* level, int32_t* coeff, int64_t n, int qp, int size, void* stream) {
    (void)size;  // quant.py:82-123 ignores it
    NH_REQUIRE(coeff && level && n >= 0, "nh_dequantize: null pointer or negative count");
    NH_REQUIRE(aligned16(coeff) && aligned16(level), "nh_dequantize: tensors must be 16-byte aligned");
    if (n == 0) return NH_OK;
    DequantF f{make_quant_params(qp, 2, 1), {}};
    f.fq = make_fast_quant(f.p);
    map_i32_kernel<<<grid_elems(n, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(level, coeff, n, f);
    NH_CHECK_LAUNCH("nh_dequantize");
    return NH_OK;
}

NH_API int nh_residual_block(const int16_t* orig, const int16_t* pred, int16_t* residual, int64_t n,
                             void* stream) {
    NH_REQUIRE(orig && pred && residual && n >= 0, "nh_residual_block: null pointer or negative count");
    NH_REQUIRE(aligned16(orig) && aligned16(pred) && aligned16(residual), "nh_residual_block: tensors must be 16-byte aligned");
    if (n == 0) return NH_OK;
    map_i16_kernel<0><<<grid_elems(n, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(orig, pred, residual, n, 0);
    NH_CHECK_LAUNCH("nh_residual_block");
    return NH_OK;
}

NH_API int nh_clip_to_pixel_range(const int16_t* in, int16_t* out, int64_t n, int bit_depth, void* stream) {
    NH_REQUIRE(in && out && n >= 0, "nh_clip_to_pixel_range: null pointer or negative count");
    NH_REQUIRE(bit_depth >= 1 && bit_depth <= 15, "nh_clip_to_pixel_range: bit_depth out of range 1..15");
    NH_REQUIRE(aligned16(in) && aligned16(out), "nh_clip_to_pixel_range: tensors must be 16-byte aligned");
    if (n == 0) return NH_OK;
    map_i16_kernel<1><<<grid_elems(n, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, in, out, n, (1 << bit_depth) - 1);
    NH_CHECK_LAUNCH("nh_clip_to_pixel_range");
    return NH_OK;
}

NH_API int nh_reconstruct_block(const int16_t* pred, const int32_t* residual, int16_t* out, int64_t n,
                                void* stream) {
    NH_REQUIRE(pred && residual && out && n >= 0, "nh_reconstruct_block: null pointer or negative count");
    NH_REQUIRE(aligned16(pred) && aligned16(residual) && aligned16(out), "nh_reconstruct_block: tensors must be 16-byte aligned");
    if (n == 0) return NH_OK;
    reconstruct_kernel<<<grid_elems(n, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, residual, out, n);
    NH_CHECK_LAUNCH("nh_reconstruct_block");
    return NH_OK;
}

NH_API int nh_intra_dc_predict(const int16_t* top, const int16_t* left, int16_t* pred, int64_t n_blocks,
                               int size, void* stream) {
    NH_REQUIRE_SIZE(size);
    NH_REQUIRE(top && left && pred && n_blocks >= 0, "nh_intra_dc_predict: null pointer or negative count");
    NH_REQUIRE(aligned16(pred), "nh_intra_dc_predict: pred must be 16-byte aligned");
    return launch_predict<0>(top, left, nullptr, nullptr, nullptr, 1, 0, pred, n_blocks, size,
                             reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_intra_dc_predict_ragged(const int16_t* top, int n_top, const int16_t* left, int n_left,
                                      int16_t* pred, int64_t n_blocks, int size, void* stream) {
    NH_REQUIRE_SIZE(size);
    NH_REQUIRE(pred && n_blocks >= 0 && n_top >= 0 && n_left >= 0 && (top || n_top == 0) && (left || n_left == 0),
               "nh_intra_dc_predict_ragged: null pointer or negative count");
    NH_REQUIRE(aligned16(pred), "nh_intra_dc_predict_ragged: pred must be 16-byte aligned");
    if (n_blocks == 0) return NH_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(n_blocks * size, 256, 8);
    switch (size) {
        case 4: predict_dc_ragged_kernel<4><<<grid, 256, 0, st>>>(top, n_top, left, n_left, pred, n_blocks); break;
        case 8: predict_dc_ragged_kernel<8><<<grid, 256, 0, st>>>(top, n_top, left, n_left, pred, n_blocks); break;
        case 16: predict_dc_ragged_kernel<16><<<grid, 256, 0, st>>>(top, n_top, left, n_left, pred, n_blocks); break;
        default: predict_dc_ragged_kernel<32><<<grid, 256, 0, st>>>(top, n_top, left, n_left, pred, n_blocks); break;
    }
    NH_CHECK_LAUNCH("predict_dc_ragged_kernel");
    return NH_OK;
}

NH_API int nh_intra_planar_predict(const int16_t* top, const int16_t* left, const int16_t* top_right,
                                   const int16_t* bottom_left, int16_t* pred, int64_t n_blocks,
                                   int size, void* stream) {
    NH_REQUIRE_SIZE(size);
    NH_REQUIRE(top && left && top_right && bottom_left && pred && n_blocks >= 0,
               "nh_intra_planar_predict: null pointer or negative count");
    NH_REQUIRE(aligned16(pred), "nh_intra_planar_predict: pred must be 16-byte aligned");
    return launch_predict<1>(top, left, top_right, bottom_left, nullptr, 0, 0, pred, n_blocks, size,
                             reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_intra_predict_modes(const int16_t* top, const int16_t* left, const int16_t* top_left,
                                  const uint8_t* modes, int mode, int allow_dc_planar, int16_t* pred,
                                  int64_t n_blocks, int size, void* stream) {
    NH_REQUIRE_SIZE(size);
    NH_REQUIRE(top && left && top_left && pred && n_blocks >= 0,
               "nh_intra_predict_modes: null pointer or negative count");
    NH_REQUIRE(aligned16(pred), "nh_intra_predict_modes: pred must be 16-byte aligned");
    if (!modes) {
        int lo = allow_dc_planar ? 0 : 2;
        if (mode < lo || mode > 34) {
            set_error("nh_intra_predict_modes: mode %d out of range %d..34", mode, lo);
            return NH_E_ARG;
        }
    }
    return launch_predict<2>(top, left, top_left, nullptr, modes, mode, allow_dc_planar, pred, n_blocks,
                             size, reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_convert_u8_to_i16(const uint8_t* src, int16_t* dst, int64_t n, void* stream) {
    NH_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "nh_convert_u8_to_i16: null pointer or negative count");
    if (n == 0) return NH_OK;
    convert_kernel<true><<<grid_for(n, 256 * 16, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, n);
    NH_CHECK_LAUNCH("convert_kernel");
    return NH_OK;
}

NH_API int nh_convert_i16_to_u8(const int16_t* src, uint8_t* dst, int64_t n, void* stream) {
    NH_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "nh_convert_i16_to_u8: null pointer or negative count");
    if (n == 0) return NH_OK;
    convert_kernel<false><<<grid_for(n, 256 * 16, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, n);
    NH_CHECK_LAUNCH("convert_kernel");
    return NH_OK;
}
