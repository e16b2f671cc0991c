// nh_coder8.cuh -- K7 winner stage for 8x8 blocks of an 8-bit plane on the tensor cores (included by
// nh_fused.cu after nh_fused_mma.cuh).  The modes were decided by search_plane_kernel (nh_search.cuh);
// this kernel is fused_mma8_kernel with a different front and back end:
//   * a warp tile is 32 consecutive blocks of one block row; its pixels arrive by cp.async straight from
//     the plane (row r of the tile = 512 contiguous bytes) and the reconstruction goes back the same way;
//   * lane u gathers the references of block u from the plane (block.py:38-55, as bytes) and predicts
//     its block for ANY of the 35 modes: angular scan lines through predict_line_u8 (the search kernel's
//     interpolation), horizontal modes transposed in registers;
//   * coefficients / levels / prediction leave block-major exactly as in fused_mma8_kernel.
// A tile that holds an undecided block (mode 0xFF: search_plane_kernel saw a sample outside [0, 255])
// is not coded here: all its blocks are marked 0xFF and the exact coder kernel, launched afterwards
// for the marked blocks only, searches and codes them.
#pragma once
#include "nh_plane.cuh"
#include "nh_search.cuh"

namespace nh {

struct Coder8Args {
    const int16_t* src;
    int H, W, pitch;
    uint8_t* modes;        // in: decided modes; out: 0xFF for the blocks of a tile left to the exact coder
    int16_t* pred;
    int32_t* coeff;
    int32_t* levels;
    int16_t* recon_plane;  // pitch as src
    int maxv;
    int n_frames;          // frames stacked `frame_stride` samples apart (src and recon_plane alike);
    int64_t frame_stride;  // block-major tensors hold frame f at block offset f * (bw * bh)
};

// 4x4 byte transpose: c[j] byte i = r[i] byte j
__device__ __forceinline__ void transpose4x4_u8(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t& c0,
                                                uint32_t& c1, uint32_t& c2, uint32_t& c3) {
    const uint32_t u = __byte_perm(r0, r1, 0x5140), v = __byte_perm(r2, r3, 0x5140);
    const uint32_t u2 = __byte_perm(r0, r1, 0x7362), v2 = __byte_perm(r2, r3, 0x7362);
    c0 = __byte_perm(u, v, 0x5410);
    c1 = __byte_perm(u, v, 0x7632);
    c2 = __byte_perm(u2, v2, 0x5410);
    c3 = __byte_perm(u2, v2, 0x7632);
}

constexpr int kC8RefBytes = 76;  // per lane: top[0..27] | left[28..55] | projected extension [56..63] | copy [64..75]

__global__ void __launch_bounds__(kV2Warps * 32, 3) coder8_plane_mma_kernel(const Coder8Args a, const FastQuant fq,
                                                                            unsigned int* tile_counter) {
    constexpr int N = 8, NN = 64, SH = 8, S1 = 4;
    using T16 = WarpTile<128>;
    constexpr int kWarpBytes = 3 * T16::kBytes + 32 * kC8RefBytes;  // 2 pixel tiles + 1 prediction tile + references
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fg = lane >> 2, ft = lane & 3;
    unsigned char* wbase = smem_raw + warp * kWarpBytes;
    // (a function of the buffer index, not an array of pointers: indexing a pointer array with a run-time value makes
    // the compiler forget the address space -- generic LD / ST instead of LDS / STS, tracked on the long scoreboard)
    auto s16 = [&](int i) -> unsigned char* { return wbase + i * T16::kBytes; };
    unsigned char* sP = wbase + 2 * T16::kBytes;
    unsigned char* rb = wbase + 3 * T16::kBytes + lane * kC8RefBytes;   // odd word stride: conflict-free
    const uint32_t lane_off = (uint32_t)((lane >> 3) * T16::kPitch + (lane & 7) * 16);

    const uint32_t tf = pack_h2((float)kc_dct8[fg * 8 + 2 * ft], (float)kc_dct8[fg * 8 + 2 * ft + 1]);
    const uint32_t ttf = pack_h2((float)kc_dct8[(2 * ft) * 8 + fg], (float)kc_dct8[(2 * ft + 1) * 8 + fg]);
    const uint4 a_fwd = make_uint4(tf, 0u, 0u, tf), a_inv = make_uint4(ttf, 0u, 0u, ttf);
    const float rnd = (float)(1 << (SH - 1));
    const float init_f2 = fg == 0 ? rnd - (float)(kOperandBias * 512) : rnd;
    const uint32_t clip_lo2 = 0x10001000u;
    const uint32_t clip_hi2 = clip_lo2 + (uint32_t)a.maxv * 0x10001u;

    const int bw = a.W / N, bh = a.H / N;
    const int tpr = (bw + 31) / 32;                 // tiles per block row
    const int64_t n_tiles = (int64_t)tpr * bh * a.n_frames;   // block rows of all frames, frame after frame
    auto next_tile = [&]() -> int64_t {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        return (int64_t)__shfl_sync(0xffffffffu, t, 0);
    };
    int64_t tile = next_tile();
    int64_t tile_next = tile < n_tiles ? next_tile() : n_tiles;

    auto prefetch = [&](int64_t t, unsigned char* dst) {   // row `it` of block `lane`: 512 contiguous bytes per warp
        const int byg = (int)(t / tpr), bx0 = (int)(t % tpr) * 32;
        const int fr = byg / bh, by = byg - fr * bh;
        if (bx0 + lane < bw) {
            const int16_t* gp = a.src + fr * a.frame_stride + (int64_t)by * N * a.pitch + (bx0 + lane) * N;
#pragma unroll
            for (int it = 0; it < 8; ++it) cp_async16(smem_u32(dst + lane * T16::kPitch + it * 16), gp + (int64_t)it * a.pitch);
        } else {
            // an idle lane's block shares its MMAs with a real block through the block-diagonal operand:
            // stale shared memory read as f16 may be Inf / NaN, and 0 x NaN would reach the real block
#pragma unroll
            for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(dst + lane * T16::kPitch + it * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    };
    int n_mode = 1;
    auto load_mode = [&](int64_t t) {
        const int byg = (int)(t / tpr), bx0 = (int)(t % tpr) * 32;   // frame f, row by = block row f * bh + by
        n_mode = bx0 + lane < bw ? (int)a.modes[(int64_t)byg * bw + bx0 + lane] : 1;
    };
    if (tile < n_tiles) {
        prefetch(tile, s16(0));
        load_mode(tile);
    }
    cp_async_commit();
    int cur = 0;
    int64_t tile_after = n_tiles;
    for (; tile < n_tiles; tile = tile_next, tile_next = tile_after, cur ^= 1) {
        const int byg = (int)(tile / tpr), bx0 = (int)(tile % tpr) * 32;
        const int fr = byg / bh, by = byg - fr * bh;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        const int64_t blk0 = (int64_t)byg * bw + bx0;
        const int blocks_valid = bw - bx0 < 32 ? bw - bx0 : 32;
        const int chunks16 = blocks_valid * 8;
        tile_after = tile_next < n_tiles ? next_tile() : n_tiles;
        const int mode = n_mode;
        const bool undecided = __any_sync(0xffffffffu, mode > 34);
        if (tile_next < n_tiles) load_mode(tile_next);
        if (!undecided) {
            // ---- K1 for block `lane` (an idle lane of a ragged tile gathers the last block again)
            const int bx = bx0 + (lane < blocks_valid ? lane : blocks_valid - 1);
            const int x = bx * N, y = by * N;
            if (x > 0 && y > 0 && x + 2 * N <= a.W && y + 2 * N <= a.H) {   // no substitution, no truncation
                const int16_t* c = srcf + (int64_t)(y - 1) * a.pitch + x - 1;
                int tv[2 * N + 1], lv[2 * N + 1];
#pragma unroll
                for (int k = 0; k <= 2 * N; ++k) {
                    tv[k] = __ldg(c + k);
                    lv[k] = __ldg(c + (int64_t)k * a.pitch);
                }
#pragma unroll
                for (int k = 0; k <= 2 * N; ++k) {
                    rb[k] = (unsigned char)tv[k];
                    rb[28 + k] = (unsigned char)lv[k];
                }
                rb[2 * N + 1] = (unsigned char)tv[2 * N];
                rb[28 + 2 * N + 1] = (unsigned char)lv[2 * N];
            } else {
#pragma unroll
                for (int k = 0; k < 2 * N + 2; ++k) {
                    const int kk = k <= 2 * N ? k : 2 * N;
                    rb[k] = (unsigned char)top_ref<false>(srcf, a.H, a.W, a.pitch, x, y, 2 * N, kk);
                    rb[28 + k] = (unsigned char)left_ref<false>(srcf, a.H, a.W, a.pitch, x, y, 2 * N, kk);
                }
            }
            // ---- lane u predicts block u into the prediction tile (rows of 16-bit samples)
            uint4* up = T16::unit(sP, lane);
            if (mode == 1) {  // intra.py:46-62
                int s = 0;
#pragma unroll
                for (int k = 1; k <= N; ++k) s += (int)rb[k] + (int)rb[28 + k];
                const uint32_t dc2 = (uint32_t)(dc_value<N>(s) & 0xffff) * 0x10001u;
#pragma unroll
                for (int yy = 0; yy < 8; ++yy) up[yy] = make_uint4(dc2, dc2, dc2, dc2);
            } else if (mode == 0) {  // intra.py:109-111, two pixels per multiply-add chain
                const uint32_t tr = rb[N + 1], bl = rb[28 + N + 1];
                uint32_t base[4], tw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    base[k] = tr * ((uint32_t)(2 * k + 1) | ((uint32_t)(2 * k + 2) << 16)) + 0x00080008u;
                    tw[k] = (uint32_t)rb[1 + 2 * k] | ((uint32_t)rb[2 + 2 * k] << 16);
                }
                const uint32_t bl2 = bl * 0x10001u;
#pragma unroll
                for (int yy = 0; yy < 8; ++yy) {
                    const uint32_t ly = rb[28 + 1 + yy];
                    uint32_t p[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ck1 = (uint32_t)(7 - 2 * k) | ((uint32_t)(6 - 2 * k) << 16);
                        const uint32_t t = tw[k] * (uint32_t)(7 - yy) + bl2 * (uint32_t)(yy + 1) + ly * ck1 + base[k];
                        p[k] = (t >> S1) & 0x00FF00FFu;
                    }
                    up[yy] = make_uint4(p[0], p[1], p[2], p[3]);
                }
            } else {  // intra.py:116-207
                const int angle = intra_angle(mode);
                const bool vertical = mode >= 18;
                const int pri = vertical ? 0 : 28, sec = vertical ? 28 : 0;
                if (angle < 0) {  // projected extension ref[-len .. -1] + a copy of ref[0 .. 11] behind it
                    const int inv = inv_angle_of_mode(mode);
                    const int len = -((N * angle) >> 5);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        reinterpret_cast<uint32_t*>(rb + 64)[c] = reinterpret_cast<const uint32_t*>(rb + pri)[c];
                    for (int tt = 0; tt < len; ++tt) {
                        int proj = (-tt * inv + 128) >> 8;   // (k+1) projection, SURVEY Q3
                        proj = proj > 2 * N ? 2 * N : proj;
                        rb[63 - tt] = rb[sec + proj];
                    }
                }
                uint32_t ln[8][2];
                int p = angle;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const uint32_t f8 = ((uint32_t)p & 31u) << 3, g8 = 256u - f8;
                    const int k = 1 + (p >> 5);
                    predict_line_u8<2>(rb, (k < 0 ? 64 : pri) + k, f8, g8, ln[s]);
                    p += angle;
                }
                if (!vertical) {  // scan line = image column: transpose the 8x8 bytes
                    uint32_t t[8][2];
                    transpose4x4_u8(ln[0][0], ln[1][0], ln[2][0], ln[3][0], t[0][0], t[1][0], t[2][0], t[3][0]);
                    transpose4x4_u8(ln[4][0], ln[5][0], ln[6][0], ln[7][0], t[0][1], t[1][1], t[2][1], t[3][1]);
                    transpose4x4_u8(ln[0][1], ln[1][1], ln[2][1], ln[3][1], t[4][0], t[5][0], t[6][0], t[7][0]);
                    transpose4x4_u8(ln[4][1], ln[5][1], ln[6][1], ln[7][1], t[4][1], t[5][1], t[6][1], t[7][1]);
#pragma unroll
                    for (int s = 0; s < 8; ++s) { ln[s][0] = t[s][0]; ln[s][1] = t[s][1]; }
                }
#pragma unroll
                for (int yy = 0; yy < 8; ++yy)
                    up[yy] = make_uint4(__byte_perm(ln[yy][0], 0u, 0x4140), __byte_perm(ln[yy][0], 0u, 0x4342),
                                        __byte_perm(ln[yy][1], 0u, 0x4140), __byte_perm(ln[yy][1], 0u, 0x4342));
            }
        }
        // the other pixel tile is free: start fetching the next tile into it, then wait for this one
        if (tile_next < n_tiles) prefetch(tile_next, s16(cur ^ 1));
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        if (undecided) {  // warp-uniform
            if (lane < blocks_valid) a.modes[blk0 + lane] = 0xFF;
            if (lane == 0) atomicAdd(tile_counter + 2, 1u);   // the exact coder kernel has work to do
            continue;
        }
        if (a.pred) T16::store(sP, reinterpret_cast<unsigned char*>(a.pred + blk0 * NN), lane, chunks16);
        const uint32_t sO = smem_u32(s16(cur)) + lane_off, sPa = smem_u32(sP) + lane_off;
        // unrolled 2x, not 8x as in fused_mma8_kernel: with the 35-mode prediction code in front of it the
        // kernel waited on instruction fetch (ncu: no-instruction 1.08 warps per issue -> 0.08)
#pragma unroll 2
        for (int q = 0; q < 8; ++q) {
            uint32_t ro[4], rp[4], pc[4], rr[4];
            const uint32_t off = (uint32_t)(4 * q * T16::kPitch);
            ldsm_x4_t(ro, sO + off);
            ldsm_x4_t(rp, sPa + off);
            ldsm_x4(pc, sPa + off);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int ba = 4 * q + 2 * p;  // blocks ba, ba + 1 of the tile
                const bool va = ba < blocks_valid, vb = ba + 1 < blocks_valid;
                float acc[4];
                const uint32_t x0 = h2_bits(__hsub2(bits_h2(ro[2 * p] | 0x64006400u), bits_h2(rp[2 * p] | 0x64006400u)));
                const uint32_t x1 = h2_bits(__hsub2(bits_h2(ro[2 * p + 1] | 0x64006400u), bits_h2(rp[2 * p + 1] | 0x64006400u)));
                hmma16816(acc, a_fwd, x0, x1, rnd, rnd, rnd, rnd);
                uint32_t h0 = round_pair_biased<SH>(acc[0], acc[1]), h1 = round_pair_biased<SH>(acc[2], acc[3]);
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
                hmma16816(acc, a_inv, h0, h1, rnd, rnd, rnd, rnd);
                h0 = round_pair_plain<SH>(acc[0], acc[1]);
                h1 = round_pair_plain<SH>(acc[2], acc[3]);
                hmma1688(acc, h0, h1, ttf, rnd, rnd, rnd, rnd);
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
        __syncwarp();
        if (a.recon_plane && lane < blocks_valid) {   // row `it` of block `lane`: 512 contiguous bytes per warp
            int16_t* gp = a.recon_plane + fr * a.frame_stride + (int64_t)by * N * a.pitch + (bx0 + lane) * N;
            uint4 v[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) v[it] = *reinterpret_cast<const uint4*>(s16(cur) + lane * T16::kPitch + it * 16);
#pragma unroll
            for (int it = 0; it < 8; ++it) stg_stream(gp + (int64_t)it * a.pitch, v[it]);
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    release_tile_counter(tile_counter, gridDim.x * kV2Warps);
}

// Declared in nh_common.cuh; called by nh_frame.cu.
int coder8_plane_mma(const int16_t* src, int n_frames, int64_t frame_stride, int H, int W, int pitch, uint8_t* modes,
                     int16_t* pred, int32_t* coeff, int32_t* levels, int16_t* recon_plane, const QuantParams& qp,
                     int maxv, cudaStream_t st, unsigned int** handed_back) {
    constexpr int kSmem = kV2Warps * (3 * WarpTile<128>::kBytes + 32 * kC8RefBytes);
    int rc = ensure_dynamic_smem(coder8_plane_mma_kernel, kSmem, "cudaFuncSetAttribute(coder8_plane_mma_kernel)");
    if (rc != NH_OK) return rc;
    const int bw = W / 8, bh = H / 8;
    const int64_t n_tiles = (int64_t)((bw + 31) / 32) * bh * n_frames;
    const int grid = grid_for(n_tiles, kV2Warps, 3);
    unsigned int* counter = nullptr;
    rc = acquire_tile_counter(st, &counter);
    if (rc != NH_OK) return rc;
    *handed_back = counter + 2;
    Coder8Args a{src, H, W, pitch, modes, pred, coeff, levels, recon_plane, maxv, n_frames, frame_stride};
    coder8_plane_mma_kernel<<<grid, kV2Warps * 32, kSmem, st>>>(a, make_fast_quant(qp), counter);
    NH_CHECK_LAUNCH("coder8_plane_mma_kernel");
    return NH_OK;
}

}  // namespace nh
