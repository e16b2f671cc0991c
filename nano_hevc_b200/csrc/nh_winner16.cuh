// nh_winner16.cuh -- K7 winner stage for 16x16 blocks of an 8-bit plane, two blocks per warp.
//
// coder_kernel<16, 32, SRC_PLANE> gives a warp one 16x16 block: 633 warp instructions per block of which the
// tensor-core chain is under half -- the rest (reference gather, prediction on 16 of the 32 lanes, DC, the decided
// mode, loop and prefetch bookkeeping) is per WARP ITERATION, so a 16x16 block costs twice what a quarter of a
// 32x32 block does (ncu: 79 against 41 thread instructions per pixel; 8 4K frames 373 us against 218 us).
// Here a warp iteration takes a PAIR of blocks: each half-warp gathers the references, builds the projected
// extension and predicts the 16 rows of its own block (one row per lane, all 32 lanes busy), then the warp runs
// the register-chained MMA pipeline of nh_mma.cuh once per block.  Pixels travel as 16-byte chunks, the next
// pair's pixels, references and modes are fetched while this one is coded.
// Only blocks whose mode the search kernel decided are coded here (decided => every sample and reference of the
// block is 8-bit, nh_search*.cuh); the others (mode 0xFF) are counted in *handed_back and left to
// coder_kernel<16, 32, SRC_PLANE> with only_undecided = 1, which exits at once when the count is zero.
#pragma once

namespace nh {

struct Winner16Cfg {
    static constexpr int N = 16;
    static constexpr int PITCH = N * 2 + 16;                 // bytes, the ldmatrix pitch of mma_block_chain<16>
    static constexpr int REF_W = CoderCfg<16, 32>::REF_W;    // int16 entries of a reference array
    static constexpr int NEG_W = CoderCfg<16, 32>::NEG_W;
    static constexpr int REFS_BYTES = ((2 * REF_W * 2 + 15) / 16) * 16;
    static constexpr int NEG_BYTES = ((15 * NEG_W * 2 + 15) / 16) * 16;
    static constexpr int TILE_BYTES = N * PITCH;
    static constexpr int BLOCK_BYTES = REFS_BYTES + NEG_BYTES + 2 * TILE_BYTES;   // refs, projection, O tile, P tile
    static constexpr int WARPS = 4;
    static constexpr int SMEM_BYTES = MmaConsts<16>::V_END * 32 * 16 + WARPS * 2 * BLOCK_BYTES;
};

__global__ void __launch_bounds__(Winner16Cfg::WARPS * 32, 4) winner16_pair_kernel(const CoderArgs a) {
    using C = Winner16Cfg;
    constexpr int N = 16, PITCH = C::PITCH;
    static_assert(PITCH == CoderCfg<16, 32>::O_PITCH * 2, "tiles must have the pitch predict / chain expect");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* ctab = reinterpret_cast<uint4*>(smem_raw);
    stage_mma_consts<16, C::WARPS * 32>(ctab);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = lane >> 4, hl = lane & 15;                // half-warp = block of the pair, lane of the half
    const MmaWinnerCtx mctx = make_mma_winner_ctx<16>(ctab, lane, a.fq, a.maxv);
    unsigned char* wbase = smem_raw + MmaConsts<16>::V_END * 32 * 16 + warp * 2 * C::BLOCK_BYTES;
    auto blk_base = [&](int i) -> unsigned char* { return wbase + i * C::BLOCK_BYTES; };
    int16_t* top = reinterpret_cast<int16_t*>(blk_base(hw));
    int16_t* left = top + C::REF_W;
    int16_t* neg = reinterpret_cast<int16_t*>(blk_base(hw) + C::REFS_BYTES);
    unsigned char* otile = blk_base(hw) + C::REFS_BYTES + C::NEG_BYTES;
    unsigned char* ptile = otile + C::TILE_BYTES;

    const int bw = a.W / N;
    const int64_t n_pairs = (a.n_blocks + 1) / 2;
    auto locate = [&](int64_t bb, int& fx, int& fy) -> int64_t {   // -> sample offset of the block's frame
        const int fr = (int)(bb / a.blocks_per_frame);
        const int64_t lb = bb - fr * a.blocks_per_frame;
        fx = (int)(lb % bw) * N;
        fy = (int)(lb / bw) * N;
        return fr * a.frame_stride;
    };
    // ---- a pair's data in registers: this lane's two 16-byte chunks of its block (chunk c = hl + 16 i: row c / 2,
    // half c % 2), three entries of each reference array (k = hl + 16 i, the last one clamped to 2N) and the mode
    uint4 npx[2];
    int ntv[3], nlv[3], nmode = 0xFF;
    auto fetch = [&](int64_t pair) {
        const int64_t bb = 2 * pair + hw;
        nmode = 0xFF;
        if (bb >= a.n_blocks) return;
        nmode = (int)a.modes_in[bb];
        if (nmode > 34) return;   // not decided: nothing of the block is needed here
        int fx, fy;
        const int16_t* srcf = a.src + locate(bb, fx, fy);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = hl + 16 * i;
            npx[i] = __ldg(reinterpret_cast<const uint4*>(srcf + (int64_t)(fy + (c >> 1)) * a.pitch + fx + 8 * (c & 1)));
        }
        const bool interior = fx > 0 && fy > 0 && fx + 2 * N <= a.W && fy + 2 * N <= a.H;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int k = hl + 16 * i, kk = k <= 2 * N ? k : 2 * N;
            if (interior) {   // no substitution: top[k] = plane[y-1][x-1+k], left[k] = plane[y-1+k][x-1]
                const int16_t* c0 = srcf + (int64_t)(fy - 1) * a.pitch + fx - 1;
                ntv[i] = __ldg(c0 + kk);
                nlv[i] = __ldg(c0 + (int64_t)kk * a.pitch);
            } else {
                ntv[i] = top_ref<false>(srcf, a.H, a.W, a.pitch, fx, fy, 2 * N, kk);
                nlv[i] = left_ref<false>(srcf, a.H, a.W, a.pitch, fx, fy, 2 * N, kk);
            }
        }
    };
    const int64_t first = (int64_t)blockIdx.x * C::WARPS + warp, stride = (int64_t)gridDim.x * C::WARPS;
    if (first < n_pairs) fetch(first);
    for (int64_t pair = first; pair < n_pairs; pair += stride) {
        const int64_t b = 2 * pair + hw;
        const int mode = nmode;
        const bool mine = mode <= 34;   // this half-warp has a block to code (half-uniform)
        if (b < a.n_blocks && !mine && hl == 0) atomicAdd(a.handed_back, 1u);
        int fx = 0, fy = 0;
        int64_t foff = 0;
        if (mine) {
            foff = locate(b, fx, fy);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = hl + 16 * i;
                *reinterpret_cast<uint4*>(otile + (c >> 1) * PITCH + 16 * (c & 1)) = npx[i];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int k = hl + 16 * i;
                if (k < C::REF_W) {
                    top[k] = (int16_t)ntv[i];
                    left[k] = (int16_t)nlv[i];
                }
            }
        }
        if (pair + stride < n_pairs) fetch(pair + stride);
        __syncwarp();
        int dc = 0;
        {   // intra.py:46-62: lane hl adds top[1 + hl] + left[1 + hl], summed over the half-warp
            int s = mine ? (int)top[1 + hl] + (int)left[1 + hl] : 0;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            dc = dc_value<N>(s);
        }
        if (mine) build_neg_array_of_mode<16, 16>(hl, mode, top, left, neg);
        __syncwarp();
        if (mine) {   // row hl of the block's prediction
            int p[N];
            uint32_t pw[N / 2];
            predict_row_u8<16, 32>(mode, hl, top, left, neg, dc, p);
            pack_row<N>(p, pw);
            if (a.out.pred) store_row16<N>(a.out.pred + b * N * N + hl * N, pw);
#pragma unroll
            for (int q = 0; q < N / 8; ++q)
                *reinterpret_cast<uint4*>(ptile + hl * PITCH + 16 * q) = make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
        }
        __syncwarp();
        const int fg = lane >> 2, ft = lane & 3;
        const uint32_t ctab_lane = mctx.ctab_lane;
        auto cv = [&](int v) -> uint4 { return ld_const_vec(ctab_lane, v); };
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {   // the whole warp codes block i of the pair
            const int mode_i = __shfl_sync(0xffffffffu, mode, 16 * i);
            if (mode_i > 34) continue;
            const int64_t bi = 2 * pair + i;
            unsigned char* ot = blk_base(i) + C::REFS_BYTES + C::NEG_BYTES;
            mma_block_chain<16>(smem_u32(ot) + mctx.lane_off, smem_u32(ot + C::TILE_BYTES) + mctx.lane_off, cv,
                                a.out.coeff != nullptr, a.out.coeff + bi * N * N + (2 * ft) * N + fg,
                                a.out.levels != nullptr, a.out.levels + bi * N * N + (2 * ft) * N + fg, a.fq,
                                mctx.dq_rnd_b, mctx.clip_lo2, mctx.clip_hi2);
        }
        __syncwarp();
        if (mine && a.out.recon_plane) {
            int16_t* rp = a.out.recon_plane + foff + (int64_t)fy * a.pitch + fx;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = hl + 16 * i;
                stg_stream(rp + (int64_t)(c >> 1) * a.pitch + 8 * (c & 1),
                           *reinterpret_cast<const uint4*>(otile + (c >> 1) * PITCH + 16 * (c & 1)));
            }
        }
        __syncwarp();   // the tiles are rewritten by the next pair
    }
}

}  // namespace nh
