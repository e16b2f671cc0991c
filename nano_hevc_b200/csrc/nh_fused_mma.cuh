// nh_fused_mma.cuh -- K10: the fused DC / planar pipeline with the four separable transform passes
// on the tensor cores (included by nh_fused.cu): fused_mma_kernel<N> for N = 16 / 32 and
// fused_mma8_kernel for N = 8 (two blocks per MMA; its own comment below).  The chain of one block
// lives in nh_mma.cuh (mma_block_chain), shared with the frame coders' winner pipeline.
//
// Why tensor cores here: ncu shows the CUDA-core 32x32 kernel limited by the integer pipe and by
// shared-memory instruction issue (78 instr/px, 0.45 of the HBM peak) -- north_star's condition for
// a tensor-core 32x32 transform.  Why warp-level HMMA (mma.sync m16n8k16, f16 x f16 -> f32) and not
// tcgen05: the reference rounds and shifts between the passes (transform.py:180-194, :222-236), so
// the four GEMMs of a block cannot be fused and every pass boundary needs the accumulators in
// registers.  With mma.sync the accumulator fragment of one pass IS the operand fragment of the next
// pass (C -> A directly; C -> B gives the transposed product), so the whole chain
//     temp  = (T  X    + r) >> s      A = T      B = X (ldmatrix.trans)     -> (m = i, n = x)
//     coefT = (T  temp^T + r) >> s    A = T      B = temp  (C -> B)         -> (m = v, n = i)
//     tmp2  = (T^T dq  + r) >> s      A = T^T    B = dq^T  (C -> B)         -> (m = y, n = v)
//     res   = (tmp2 T  + r) >> s      A = tmp2 (C -> A)   B = T             -> (m = y, n = x)
// runs without a single shuffle or shared-memory round trip between passes; tcgen05 would need a
// TMEM load, a rounding pass and a shared-memory store of the next operand at every boundary.
//
// Exactness: every MMA operand is an integer of magnitude <= 2048 (exact in f16) and every
// accumulator stays below 2^24 (exact in f32) when all samples lie in [0, 255]:
//     |residual| <= 255, |temp| <= 511, |coeff| <= 1023, |dequantised| <= 360 / 180 (N = 16 / 32, any
//     QP, intra or inter), |tmp2| <= 661 / 328, accumulators <= 1.05e6
// (tests/test_host_math.py::test_mma_operand_bounds recomputes these from the reference tables).
// The rounding offset r rides in as the accumulator's initial value and the floor shift is one
// FFMA.RM against 1.5 * 2^23 (the integer appears in the low mantissa bits).  Tiles with a sample
// outside [0, 255] (or a clip bound above 1023) take the exact CUDA-core path (rows_tile_exact).
#pragma once
#include "nh_mma.cuh"

namespace nh {

template <int N, int OCC>
__global__ void __launch_bounds__(kMmaWarps * 32, OCC) fused_mma_kernel(const FusedArgs a, const FastQuant fq) {
    using C = MmaConsts<N>;
    constexpr int NN = N * N;
    constexpr int BPW = 32 / N;  // blocks per warp tile
    constexpr int S1 = Log2<N>::v + 1;
    // int16 block tile in shared memory: row pitch N*2 + 16 bytes (an odd number of 16-byte groups),
    // so the 8 rows of an ldmatrix / stmatrix 8x8 tile and the per-lane 128-bit row accesses are
    // conflict-free.
    constexpr int PITCH = N * 2 + 16;
    constexpr int TILE = N * PITCH;
    constexpr int TOP_OFF = 2 * BPW * TILE;  // then one row of top references per block
    constexpr int FAST_BYTES = TOP_OFF + BPW * N * 2;
    constexpr int EXACT_BYTES = BPW * RowsTile<N>::WORDS * 4;
    constexpr int WARP_BYTES = FAST_BYTES > EXACT_BYTES ? FAST_BYTES : EXACT_BYTES;
    __shared__ __align__(16) unsigned char smem[kMmaWarps][WARP_BYTES];
    // Per-lane constants: staged once per CTA into shared memory (vector v of lane l at ctab[v][l]:
    // a warp's 128-bit read is 512 contiguous bytes) and re-read where they are used.  Keeping the 48
    // words in registers across the tile loop made the first version spill its prefetched pixels;
    // the loads are volatile asm so that they stay inside the loop.
    __shared__ __align__(16) uint4 ctab[C::V_END][32];
    stage_mma_consts<N, kMmaWarps * 32>(&ctab[0][0]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / N, r = lane % N;    // I/O mapping: lane = row r of block g
    const int fg = lane >> 2, ft = lane & 3; // fragment mapping
    unsigned char* sm = smem[warp];
    unsigned char* my_o = sm + g * TILE + r * PITCH;          // this lane's row of the original tile
    unsigned char* my_p = sm + (BPW + g) * TILE + r * PITCH;  // ... and of the prediction tile
    int16_t* my_top = reinterpret_cast<int16_t*>(sm + TOP_OFF) + g * N;
    // ldmatrix / stmatrix x4 row address of this lane: 8x8 tile j = lane >> 3 covers rows
    // 8*(j&1).. and columns 8*(j>>1)..  of a 16x16 region
    const int lane_off = (((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + (lane >> 4) * 16;
    const uint32_t ctab_lane = smem_u32(&ctab[0][lane]);
    auto cv = [&](int v) -> uint4 { return ld_const_vec(ctab_lane, v); };

    const bool clip_ok = a.maxv <= 1023;
    const uint32_t clip_lo2 = 0x08000800u;  // reconstruction carries a +2048 bias per 16-bit half
    const uint32_t clip_hi2 = clip_lo2 + (uint32_t)(clip_ok ? a.maxv : 0) * 0x10001u;
    const int dq_rnd_b = fq.dq_rnd + (kOperandBits << fq.dq_shift);  // dequantised value + 0x6600

    const int64_t n_tiles = (a.n_blocks + BPW - 1) / BPW;
    const int64_t warp_stride = (int64_t)gridDim.x * kMmaWarps;
    int64_t tile = (int64_t)blockIdx.x * kMmaWarps + warp;

    // One tile ahead: the warp tile's pixels and this lane's reference samples.  The BPW blocks of a
    // tile are contiguous in memory, so lanes sweep them in 16-byte chunks (512 contiguous bytes per
    // warp instruction) and scatter the chunks into the shared tile; pred / recon leave the same way.
    // (Row-per-lane accesses touch 32 different lines per instruction; the single-stage transform
    // kernels gained 30 % from this change.)
    constexpr int CPL = BPW * NN * 2 / 16 / 32;  // chunks per lane
    auto chunk_ptr = [&](unsigned char* tiles, int it) -> unsigned char* {
        const int e0 = (it * 32 + lane) * 8;     // first pixel of the chunk within the warp tile
        return tiles + (e0 / NN) * TILE + ((e0 % NN) / N) * PITCH + (e0 % N) * 2;
    };
    auto sweep_out = [&](unsigned char* tiles, int16_t* dst, int64_t t) {  // shared tiles -> global, coalesced
        const int64_t e_valid = (a.n_blocks - t * BPW) * NN;
        uint4 v[CPL];
#pragma unroll
        for (int it = 0; it < CPL; ++it) v[it] = *reinterpret_cast<const uint4*>(chunk_ptr(tiles, it));
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
            const int c = it * 32 + lane;
            if ((int64_t)c * 8 < e_valid) stg_stream(dst + t * BPW * NN + 8 * c, v[it]);
        }
    };
    uint4 nxt_px[CPL];
    int nxt_top = 0, nxt_left = 0, nxt_tr = 0, nxt_bl = 0, nxt_mode = 0;
    auto prefetch = [&](int64_t t) {
        const int64_t b = t * BPW + g;
        const int64_t e_valid = (a.n_blocks - t * BPW) * NN;
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
            const int c = it * 32 + lane;
            nxt_px[it] = (int64_t)c * 8 < e_valid ? ldg_stream(a.orig + t * BPW * NN + 8 * c) : make_uint4(0u, 0u, 0u, 0u);
        }
        if (b < a.n_blocks) {
            nxt_top = a.top[b * N + r];
            nxt_left = a.left[b * N + r];
            nxt_tr = a.top_right[b];
            nxt_bl = a.bottom_left[b];
            nxt_mode = a.modes ? (int)a.modes[b] : a.mode;
        } else {
            nxt_top = nxt_left = nxt_tr = nxt_bl = 0;
            nxt_mode = 1;
        }
    };
    if (tile < n_tiles) prefetch(tile);

    for (; tile < n_tiles; tile += warp_stride) {
        const int64_t b = tile * BPW + g;
        const bool valid = b < a.n_blocks;
        uint32_t ood = (uint32_t)(nxt_top | nxt_left | nxt_tr | nxt_bl) & 0xFFFFFF00u;  // outside [0, 255]
#pragma unroll
        for (int it = 0; it < CPL; ++it) {
            *reinterpret_cast<uint4*>(chunk_ptr(sm, it)) = nxt_px[it];
            ood |= (nxt_px[it].x | nxt_px[it].y | nxt_px[it].z | nxt_px[it].w) & 0xFF00FF00u;
        }
        my_top[r] = (int16_t)nxt_top;
        const int left_r = nxt_left, tr = nxt_tr, bl = nxt_bl, mode = nxt_mode;
        // sum of the block's 2N reference samples.  (A partial-mask __reduce_add_sync would split
        // the warp -- WARPSYNC.EXCLUSIVE + REDUX into a uniform register -- and leave the halves
        // diverged in front of the warp-collective instructions below.)
        int ref_sum = nxt_top + nxt_left;
        if constexpr (BPW == 1) {
            ref_sum = __reduce_add_sync(0xffffffffu, ref_sum);
        } else {
#pragma unroll
            for (int o = N / 2; o > 0; o >>= 1) ref_sum += __shfl_xor_sync(0xffffffffu, ref_sum, o);
        }
        const bool fast = clip_ok && !__any_sync(0xffffffffu, ood != 0);
        if (tile + warp_stride < n_tiles) prefetch(tile + warp_stride);
        __syncwarp();
        // `fast` is warp-uniform and the ONLY branch around the warp-collective instructions
        // (ldmatrix / mma / stmatrix .sync.aligned); the per-lane mode test is a select inside it.
        uint32_t pw[N / 2];  // this lane's prediction row, packed
        uint32_t tw[N / 2];
#pragma unroll
        for (int q = 0; q < N / 8; ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(my_top + 8 * q);
            tw[4 * q] = v.x; tw[4 * q + 1] = v.y; tw[4 * q + 2] = v.z; tw[4 * q + 3] = v.w;
        }
        const uint32_t dc2 = (uint32_t)(dc_value<N>(ref_sum) & 0xffff) * 0x10001u;  // intra.py:46-62
        if (fast) {
            // intra.py:109-111 on 16-bit pairs: every term is non-negative and the sum of a pair
            // member stays below 2^16, so one 32-bit multiply-add chain serves two pixels
            const uint32_t wy = (uint32_t)(N - 1 - r);
            const uint32_t c0 = (uint32_t)((r + 1) * bl + N) * 0x10001u;
#pragma unroll
            for (int k = 0; k < N / 2; ++k) {
                const uint32_t ck1 = (uint32_t)(N - 1 - 2 * k) | ((uint32_t)(N - 2 - 2 * k) << 16);
                const uint32_t ck2 = (uint32_t)(2 * k + 1) | ((uint32_t)(2 * k + 2) << 16);
                const uint32_t t = tw[k] * wy + c0 + (uint32_t)left_r * ck1 + (uint32_t)tr * ck2;
                pw[k] = mode == 1 ? dc2 : ((t >> S1) & 0x00FF00FFu);
            }
#pragma unroll
            for (int q = 0; q < N / 8; ++q)
                *reinterpret_cast<uint4*>(my_p + 16 * q) =
                    make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
            __syncwarp();
            if (a.pred) sweep_out(sm + BPW * TILE, a.pred, tile);
#pragma unroll
            for (int u = 0; u < BPW; ++u) {
                const int64_t bu = tile * BPW + u;
                const bool valid_u = bu < a.n_blocks;  // warp-uniform
                const bool want_c = valid_u && a.coeff != nullptr, want_l = valid_u && a.levels != nullptr;
                // element (i = 2ft + ..., v = fg + ...) of the transposed coefficient fragments
                int32_t* cp = a.coeff + bu * NN + (2 * ft) * N + fg;
                int32_t* lp = a.levels + bu * NN + (2 * ft) * N + fg;
                const uint32_t so = smem_u32(sm + u * TILE) + lane_off;
                const uint32_t sp = smem_u32(sm + (BPW + u) * TILE) + lane_off;
                mma_block_chain<N>(so, sp, cv, want_c, cp, want_l, lp, fq, dq_rnd_b, clip_lo2, clip_hi2);
            }
            __syncwarp();
            if (a.recon) sweep_out(sm, a.recon, tile);
        } else {
            // exact CUDA-core path: residual rows into the int32 working matrix (aliases the tiles)
            {
                int top[N], p[N];
                unpack_row<N>(tw, top);
                planar_row<N>(r, left_r, top, tr, bl, p);
                pack_row<N>(p, pw);
#pragma unroll
                for (int k = 0; k < N / 2; ++k) pw[k] = mode == 1 ? dc2 : pw[k];
            }
            if (valid && a.pred) store_row16<N>(a.pred + b * NN + r * N, pw);
            uint32_t ow[N / 2];
#pragma unroll
            for (int q = 0; q < N / 8; ++q) {
                const uint4 vo = *reinterpret_cast<const uint4*>(my_o + 16 * q);
                ow[4 * q] = vo.x; ow[4 * q + 1] = vo.y; ow[4 * q + 2] = vo.z; ow[4 * q + 3] = vo.w;
            }
            __syncwarp();
            int* M = reinterpret_cast<int*>(sm) + g * RowsTile<N>::WORDS;
            {
                int res[N];
#pragma unroll
                for (int k = 0; k < N / 2; ++k) {
                    res[2 * k] = sext16(lo16(ow[k]) - lo16(pw[k]));
                    res[2 * k + 1] = sext16(hi16(ow[k]) - hi16(pw[k]));
                }
                store_row_smem<N>(M, r, res);
            }
            __syncwarp();
            if (a.ood_flag && lane == 0) *a.ood_flag = 1;
            // copies made HERE so that neither the kernel parameters nor pw have their address taken
            // on the hot path (they would live in local memory for the whole loop otherwise)
            const FusedArgs a_cold = a;
            uint32_t pw_cold[N / 2];
#pragma unroll
            for (int k = 0; k < N / 2; ++k) pw_cold[k] = pw[k];
            rows_tile_exact<N>(a_cold, M, r, valid, b, pw_cold);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// N = 8 on the tensor cores.  Two 8x8 blocks share one m16n8k16 MMA through a block-diagonal
// constant operand, so the chain of the 16 / 32 kernel carries over with every fragment full:
//     temp   (m=(blk,i), n=x) = diag(T,T)     [X_a; X_b]            A = {tf, 0, 0, tf}   B = ldmatrix.trans
//     coef^T (m=(blk,v), n=i) = diag(T,T)     [temp_a^T; temp_b^T]  B = previous C fragments
//     tmp2   (m=(blk,y), n=v) = diag(T^T,T^T) [dq_a; dq_b]          A = {ttf, 0, 0, ttf}
//     res    (m=(blk,y), n=x) = [tmp2_a; tmp2_b] T                  m16n8k8, A = previous C, B = {ttf}
// with tf = {T[g][2t], T[g][2t+1]} and ttf = {T[2t][g], T[2t+1][g]}: the whole transform matrix is
// two registers per lane.  Operand magnitudes for 8-bit samples: |temp| <= 511 (biased form),
// |dq| <= 720, |tmp2| <= 1348 (plain f16), |res| <= 2523, accumulators <= 6.5e5.
// Staging is the unit kernel's: a warp tile is 32 blocks, pixels arrive by cp.async one tile ahead
// into padded shared tiles (block u at u * 144 bytes: the 8 rows of a block are one conflict-free
// ldmatrix 8x8 tile), lane u predicts block u, pred / recon leave through the cooperative 128-bit
// sweep.  Coefficients and levels are stored straight from the fragments (each STG.32 of a warp
// fills four whole 32-byte sectors).  A tile with any sample outside [0, 255] is recoded afterwards,
// one block per lane, with the exact reference arithmetic (slow_block).
static __constant__ signed char kc_dct8[64] = {
    64, 64, 64, 64, 64, 64, 64, 64, 89, 75, 50, 18, -18, -50, -75, -89, 83, 36, -36, -83, -83, -36, 36, 83,
    75, -18, -89, -50, 50, 89, 18, -75, 64, -64, -64, 64, 64, -64, -64, 64, 50, -89, 18, 75, -75, -18, 89, -50,
    36, -83, 83, -36, -36, 83, -83, 36, 18, -50, 75, -89, 89, -75, 50, -18};

__device__ __forceinline__ void hmma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, float c0,
                                         float c1, float c2, float c3) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(b0), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}

#ifndef NH_TILES_PER_TICKET
#define NH_TILES_PER_TICKET 1
#endif
#ifndef NH_MMA8_UNROLL
#define NH_MMA8_UNROLL 8
#endif
constexpr int kMma8Unroll = NH_MMA8_UNROLL;  // ldmatrix quads (4 blocks each) unrolled per loop trip
template <int OCC>
__global__ void __launch_bounds__(kV2Warps * 32, OCC) fused_mma8_kernel(const FusedArgs a, const FastQuant fq,
                                                                         unsigned int* tile_counter) {
    constexpr int N = 8, NN = 64, SH = 8, S1 = 4;
    using T16 = WarpTile<128>;
    constexpr int kWarpBytes = 3 * T16::kBytes;  // 2 pixel tiles (double buffer) + 1 prediction tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fg = lane >> 2, ft = lane & 3;
    unsigned char* wbase = smem_raw + warp * kWarpBytes;
    // (a function of the buffer index, not an array of pointers: indexing a pointer array with a run-time value makes
    // the compiler forget the address space -- generic LD / ST instead of LDS / STS, tracked on the long scoreboard)
    auto s16 = [&](int i) -> unsigned char* { return wbase + i * T16::kBytes; };
    unsigned char* sP = wbase + 2 * T16::kBytes;
    // ldmatrix / stmatrix x4: 8x8 tile j = lane >> 3 is block 4q + j, row lane & 7
    const uint32_t lane_off = (uint32_t)((lane >> 3) * T16::kPitch + (lane & 7) * 16);

    const uint4 tf4 = make_uint4(
        pack_h2((float)kc_dct8[fg * 8 + 2 * ft], (float)kc_dct8[fg * 8 + 2 * ft + 1]), 0u, 0u, 0u);
    const uint32_t tf = tf4.x;
    const uint32_t ttf = pack_h2((float)kc_dct8[(2 * ft) * 8 + fg], (float)kc_dct8[(2 * ft + 1) * 8 + fg]);
    const uint4 a_fwd = make_uint4(tf, 0u, 0u, tf), a_inv = make_uint4(ttf, 0u, 0u, ttf);
    const float rnd = (float)(1 << (SH - 1));
    // forward second pass takes temp + 1536: remove 1536 * sum_x T[v][x] (= 512 for v = 0, else 0)
    const float init_f2 = fg == 0 ? rnd - (float)(kOperandBias * 512) : rnd;
    const uint32_t clip_lo2 = 0x10001000u;  // reconstruction carries a +4096 bias per 16-bit half
    const uint32_t clip_hi2 = clip_lo2 + (uint32_t)a.maxv * 0x10001u;  // launcher guarantees maxv <= 1023

    const int64_t n_tiles = (a.n_blocks + 31) / 32;
    // Tiles are handed out dynamically (one atomic per tile, fetched one tile ahead): with a static
    // partition the kernel waits for its slowest SM -- ncu showed SM active cycles spread over
    // 1.02M .. 1.21M for an average of 1.09M.
    constexpr int kTilesPerTicket = NH_TILES_PER_TICKET;
    int64_t ticket_base = 0;
    int ticket_left = 0;
    auto next_tile = [&]() -> int64_t {
        if (ticket_left == 0) {
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(tile_counter, 1u);
            ticket_base = (int64_t)__shfl_sync(0xffffffffu, t, 0) * kTilesPerTicket;
            ticket_left = kTilesPerTicket;
        }
        return ticket_base + (kTilesPerTicket - ticket_left--);
    };
    int64_t tile = next_tile();
    int64_t tile_next = tile < n_tiles ? next_tile() : n_tiles;

    auto prefetch = [&](int64_t t, unsigned char* dst) {
        const int64_t blk0 = t * 32;
        const int64_t rem = a.n_blocks - blk0;
        const int chunks = (int)(rem < 32 ? rem : 32) * 8;
        const unsigned char* gp = reinterpret_cast<const unsigned char*>(a.orig + blk0 * NN);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int c = it * 32 + lane;
            if (c < chunks) cp_async16(smem_u32(dst + (c >> 3) * T16::kPitch + (c & 7) * 16), gp + (size_t)c * 16);
        }
    };
    // references of lane u's block, one tile ahead like the pixels
    uint32_t n_tw[4], n_lw[4];
    int n_tr = 0, n_bl = 0, n_mode = 1;
    auto load_refs = [&](int64_t t) {
        const int64_t b = t * 32 + lane;
        if (b < a.n_blocks) {
            load_row16<N>(a.top + b * N, n_tw);
            load_row16<N>(a.left + b * N, n_lw);
            n_tr = a.top_right[b];
            n_bl = a.bottom_left[b];
            n_mode = a.modes ? (int)a.modes[b] : a.mode;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) n_tw[k] = n_lw[k] = 0;
            n_tr = n_bl = 0;
            n_mode = 1;
        }
    };
    if (tile < n_tiles) {
        prefetch(tile, s16(0));
        load_refs(tile);
    }
    cp_async_commit();
    int cur = 0;
    int64_t tile_after = n_tiles;
    for (; tile < n_tiles; tile = tile_next, tile_next = tile_after, cur ^= 1) {
        const int64_t blk0 = tile * 32;
        const int64_t trem = a.n_blocks - blk0;
        const int blocks_valid = (int)(trem < 32 ? trem : 32);
        const int chunks16 = blocks_valid * 8;
        tile_after = tile_next < n_tiles ? next_tile() : n_tiles;  // ticket for the tile after next
        uint32_t ood = 0;  // any sample outside [0, 255]
        // ---- lane u predicts block u (row layout, 16-bit pairs) into the prediction tile
        {
            uint32_t tw[4], lw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { tw[k] = n_tw[k]; lw[k] = n_lw[k]; ood |= (tw[k] | lw[k]) & 0xFF00FF00u; }
            const int tr = n_tr, bl = n_bl, mode = n_mode;
            ood |= (uint32_t)(tr | bl) & 0xFFFFFF00u;
            if (tile_next < n_tiles) load_refs(tile_next);
            uint4* up = T16::unit(sP, lane);
            if (mode == 1) {  // intra.py:46-62
                int s = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) s += lo16(tw[k]) + hi16(tw[k]) + lo16(lw[k]) + hi16(lw[k]);
                const uint32_t dc2 = (uint32_t)(dc_value<N>(s) & 0xffff) * 0x10001u;
#pragma unroll
                for (int y = 0; y < 8; ++y) up[y] = make_uint4(dc2, dc2, dc2, dc2);
            } else {  // intra.py:109-111, two pixels per multiply-add chain (exact for 8-bit samples)
                uint32_t base[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    base[k] = (uint32_t)tr * ((uint32_t)(2 * k + 1) | ((uint32_t)(2 * k + 2) << 16)) + 0x00080008u;
                const uint32_t bl2 = (uint32_t)bl * 0x10001u;
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    const uint32_t ly = (y & 1) ? (lw[y >> 1] >> 16) : (lw[y >> 1] & 0xffffu);
                    uint32_t p[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ck1 = (uint32_t)(7 - 2 * k) | ((uint32_t)(6 - 2 * k) << 16);
                        const uint32_t t = tw[k] * (uint32_t)(7 - y) + bl2 * (uint32_t)(y + 1) + ly * ck1 + base[k];
                        p[k] = (t >> S1) & 0x00FF00FFu;
                    }
                    up[y] = make_uint4(p[0], p[1], p[2], p[3]);
                }
            }
        }
        // the other pixel tile is free: start fetching the next tile into it, then wait for this one
        if (tile_next < n_tiles) prefetch(tile_next, s16(cur ^ 1));
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        if (a.pred) T16::store(sP, reinterpret_cast<unsigned char*>(a.pred + blk0 * NN), lane, chunks16);
        const uint32_t sO = smem_u32(s16(cur)) + lane_off, sPa = smem_u32(sP) + lane_off;
        uint32_t oodw = 0;
#pragma unroll kMma8Unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t ro[4], rp[4], pc[4], rr[4];
            const uint32_t off = (uint32_t)(4 * q * T16::kPitch);
            ldsm_x4_t(ro, sO + off);
            ldsm_x4_t(rp, sPa + off);
            ldsm_x4(pc, sPa + off);
            oodw |= ro[0] | ro[1] | ro[2] | ro[3];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int ba = 4 * q + 2 * p;  // blocks ba, ba + 1 of the tile
                const bool va = ba < blocks_valid, vb = ba + 1 < blocks_valid;
                float acc[4];
                // forward, first pass
                const uint32_t x0 = h2_bits(__hsub2(bits_h2(ro[2 * p] | 0x64006400u), bits_h2(rp[2 * p] | 0x64006400u)));
                const uint32_t x1 = h2_bits(__hsub2(bits_h2(ro[2 * p + 1] | 0x64006400u), bits_h2(rp[2 * p + 1] | 0x64006400u)));
                hmma16816(acc, a_fwd, x0, x1, rnd, rnd, rnd, rnd);
                uint32_t h0 = round_pair_biased<SH>(acc[0], acc[1]), h1 = round_pair_biased<SH>(acc[2], acc[3]);
                // forward, second pass (transposed): acc[e] = coeff_blk[i = 2t + (e&1)][v = g], blk = e >> 1
                hmma16816(acc, a_fwd, h0, h1, init_f2, init_f2, init_f2, init_f2);
                float dqf[4];
                int32_t* cp = a.coeff + (blk0 + ba) * NN + (2 * ft) * N + fg;
                int32_t* lp = a.levels + (blk0 + ba) * NN + (2 * ft) * N + fg;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = __float_as_int(floor_shift_magic<SH>(acc[e])) - kMagicI;
                    const int lv = quantize_fast(c, fq);
                    const int dq = dequantize_fast(lv, fq);
                    const bool v = (e >> 1) ? vb : va;
                    const int o = (e & 1) * N + (e >> 1) * NN;
                    if (v && a.coeff) __stcs(cp + o, c);
                    if (v && a.levels) __stcs(lp + o, lv);
                    dqf[e] = __int_as_float(dq + kMagicI) - kMagicF;
                }
                h0 = pack_h2(dqf[0], dqf[1]);
                h1 = pack_h2(dqf[2], dqf[3]);
                // inverse, first pass
                hmma16816(acc, a_inv, h0, h1, rnd, rnd, rnd, rnd);
                h0 = round_pair_plain<SH>(acc[0], acc[1]);
                h1 = round_pair_plain<SH>(acc[2], acc[3]);
                // inverse, second pass: acc[e] = res_blk[y = g][x = 2t + (e&1)], blk = e >> 1
                hmma1688(acc, h0, h1, ttf, rnd, rnd, rnd, rnd);
                // reconstruct + clip on 16-bit pairs carrying a +4096 bias (|res| <= 2523)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t m0 = __float_as_uint(__fmaf_rd(acc[2 * j], 1.0f / (float)(1 << SH), kMagicF + 4096.0f));
                    const uint32_t m1 = __float_as_uint(__fmaf_rd(acc[2 * j + 1], 1.0f / (float)(1 << SH), kMagicF + 4096.0f));
                    const uint32_t s = __byte_perm(m0, m1, 0x5410) + pc[2 * p + j];
                    rr[2 * p + j] = __vminu2(__vmaxu2(s, clip_lo2), clip_hi2) - clip_lo2;
                }
            }
            stsm_x4(sO + off, rr);
        }
        ood |= oodw & 0xFF00FF00u;
        __syncwarp();
        if (a.recon) T16::store(s16(cur), reinterpret_cast<unsigned char*>(a.recon + blk0 * NN), lane, chunks16);
        __syncwarp();
        // any sample of the tile outside [0, 255]: recode it exactly, one block per lane (cold path;
        // the __syncwarp above orders the cooperative stores before these)
        if (__any_sync(0xffffffffu, ood != 0)) {
            const FusedArgs a_cold = a;
            if (lane < blocks_valid) slow_block<N>(a_cold, blk0 + lane, false);
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    release_tile_counter(tile_counter, gridDim.x * kV2Warps);
}

template <int OCC>
static int launch_mma8_occ(const FusedArgs& a, cudaStream_t st) {
    constexpr int kSmem = kV2Warps * 3 * WarpTile<128>::kBytes;
    {
        const int rc = ensure_dynamic_smem(fused_mma8_kernel<OCC>, kSmem, "cudaFuncSetAttribute(fused_mma8_kernel)");
        if (rc != NH_OK) return rc;
    }
    int grid = grid_for(a.n_blocks, (int64_t)kV2Warps * 32, OCC);
    unsigned int* counter = nullptr;
    {
        const int rc = acquire_tile_counter(st, &counter);
        if (rc != NH_OK) return rc;
    }
    fused_mma8_kernel<OCC><<<grid, kV2Warps * 32, kSmem, st>>>(a, make_fast_quant(a.qp), counter);
    NH_CHECK_LAUNCH("fused_mma8_kernel");
    tile_counter_launched(st);
    return NH_OK;
}
static int launch_mma8(const FusedArgs& a, cudaStream_t st) {
    static const int occ = [] {   // read once (thread-safe static initialisation)
        const char* e = getenv("NH_MMA_OCC");
        return (e && e[0] == '4') ? 4 : 3;  // measured: 0.87 of the HBM peak at 3 CTAs / SM, 0.80 at 4
    }();
    return occ == 3 ? launch_mma8_occ<3>(a, st) : launch_mma8_occ<4>(a, st);
}

template <int N>
static int launch_mma(const FusedArgs& a, cudaStream_t st) {
    constexpr int BPW = 32 / N;
    static const int occ = [] {  // resident CTAs per SM the kernel is compiled for: NH_MMA_OCC=3|4 (A/B profiling)
        const char* e = getenv("NH_MMA_OCC");
        return (e && e[0] == '3') ? 3 : 4;  // N = 16: 427 vs 414 Gpix/s, N = 32: 445 vs 448
    }();
    int grid = grid_for(a.n_blocks, (int64_t)kMmaWarps * BPW, occ);
    if (occ == 3) fused_mma_kernel<N, 3><<<grid, kMmaWarps * 32, 0, st>>>(a, make_fast_quant(a.qp));
    else fused_mma_kernel<N, 4><<<grid, kMmaWarps * 32, 0, st>>>(a, make_fast_quant(a.qp));
    NH_CHECK_LAUNCH("fused_mma_kernel");
    return NH_OK;
}

}  // namespace nh
