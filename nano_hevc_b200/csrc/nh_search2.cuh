// nh_search2.cuh -- K7 search stage for 8-bit planes at N = 8 / 16 / 32, second generation ("line-synchronous").
//
// The first generation (nh_search.cuh, still used at N = 4) gives a lane a strip of 4 scan lines x 8 samples; which
// 4 lines depends on the lane, so the integer offset and the fraction of a scan line -- and with them the word
// offset, the funnel-shift amount, the byte selector and the two weights of the interpolation -- are per-lane
// values recomputed on the ALU pipe for every line: 12 of the 36 ALU-pipe instructions of a mirror-pair line
// (ncu, profiles/r1_searchA8_ncu_summary.json: ALU pipe 72 %, FMA pipe 22 %, 162 thread instructions per pixel).
// Here every lane of a warp evaluates THE SAME scan line of the same mode at the same time:
//   * lane = (block, 8-sample segment of the line): N = 8 / 16 / 32 -> 32 / 16 / 8 blocks per warp tile, 1 / 2 / 4
//     lanes per block; the block is walked in passes of 8 scan lines (1 / 2 / 4 passes);
//   * the position of a scan line (intra.py:191-207: k = 1 + ((y+1) * angle >> 5), f = (y+1) * angle & 31) is then
//     warp-uniform: offset, shift, selector and weights live in uniform registers and the per-line work is the 3
//     shared-memory words, 2 funnel shifts, 6 spreading PRMTs, 8 multiply-adds, 2 packing PRMTs and 2 VABSDIFF4
//     of each half of the mirror pair -- nothing else;
//   * the only per-lane address is the block's byte base + 8 * segment.  For that the per-mode arrays of the
//     negative angles hold, behind the projected extension of intra.py:180-186, a copy of the first N + 4 bytes
//     of the primary array (nh_search.cuh copies 12): every window of such a mode, whatever the segment, is read
//     from the mode's own array and no lane has to choose between two arrays;
//   * partial costs of the passes are added up in shared memory (one word per lane and candidate), the lanes of
//     a block are summed by shuffles in the last pass only.
// Candidate order, tie rule and the hand-back of tiles with samples outside [0, 255] are those of nh_search.cuh.
#pragma once
#include "nh_search.cuh"

namespace nh {

template <int N>
struct LineCfg {
    static constexpr int SEG = N / 8;                  // lanes (8-sample segments) per block
    static constexpr int T = 32 / SEG;                 // blocks per warp tile
    static constexpr int G = N / 8;                    // passes of 8 scan lines
    static constexpr int PB = ((2 * N + 9) + 3) / 4 * 4;   // bytes of a positive array: ref[0 .. 2N+1] + word-read slack
    static constexpr int CP = N + 4;                   // bytes of the primary array copied behind a projected extension
    __host__ __device__ static constexpr int neg_len(int mi) { return -((N * neg_angle_at(mi)) >> 5); }
    __host__ __device__ static constexpr int neg_t0(int mi) {   // byte t = 0 of mode 11 + mi, from the block base
        int off = 2 * PB;
        for (int m = 0; m < mi; ++m) off += (neg_len(m) + 3) / 4 * 4 + CP;
        return off + (neg_len(mi) + 3) / 4 * 4;
    }
    static constexpr int neg_bytes() {
        int s = 0;
        for (int mi = 0; mi < 15; ++mi) s += (neg_len(mi) + 3) / 4 * 4 + CP;
        return s;
    }
    static constexpr int BLOCK_WORDS = ((2 * PB + neg_bytes()) / 4) | 1;   // odd: blocks spread over the banks
    static constexpr int WARP_WORDS = T * BLOCK_WORDS;
    static constexpr int GP = T >= 16 ? 1 : 2;         // build: groups of modes per orientation and block
    static constexpr int MPG = 8 / GP;                 // modes per group
    static constexpr int WARPS = N == 8 ? 4 : 8;       // N >= 16: the CTA's scan-line table and the partial costs favour large CTAs
    static constexpr int ACC_WORDS = G > 1 ? 35 * WARPS * 32 : 0;   // partial costs: [candidate][thread]
    static constexpr int TAB_WORDS = 17 * N * 5;       // scan-line table of the CTA: 17 rows x N lines x (int4 + int)
    static constexpr int SMEM_BYTES = (WARPS * WARP_WORDS + ACC_WORDS + TAB_WORDS) * 4;
    static constexpr int PER_SM = N == 8 ? 5 : 2;      // resident CTAs (shared memory: 5 x 44 KB, 2 x 109 KB, 2 x 107 KB)
};

// byte t = 0 of the array of mode 11 + mi (constant bank: a uniform index gives a uniform register)
#define NH_NEGT0_ROW(N) {LineCfg<N>::neg_t0(0), LineCfg<N>::neg_t0(1), LineCfg<N>::neg_t0(2), LineCfg<N>::neg_t0(3),   \
                         LineCfg<N>::neg_t0(4), LineCfg<N>::neg_t0(5), LineCfg<N>::neg_t0(6), LineCfg<N>::neg_t0(7),   \
                         LineCfg<N>::neg_t0(8), LineCfg<N>::neg_t0(9), LineCfg<N>::neg_t0(10), LineCfg<N>::neg_t0(11), \
                         LineCfg<N>::neg_t0(12), LineCfg<N>::neg_t0(13), LineCfg<N>::neg_t0(14), 0}
static __constant__ int kc_line_negt0[3][16] = {NH_NEGT0_ROW(8), NH_NEGT0_ROW(16), NH_NEGT0_ROW(32)};

// (the scan-line table kc_line_tab and prmt() live in nh_search.cuh: the strip kernel uses them at N = 4)
__device__ __forceinline__ void predict_line8(const uint32_t* wp, uint32_t sh, uint32_t sel_last, uint32_t f8, uint32_t g8,
                                              uint32_t (&out)[2]) {
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);   // bytes k .. k+3, k+4 .. k+7
    const uint32_t e0 = __byte_perm(v0, 0u, 0x4240), o0 = __byte_perm(v0, 0u, 0x4341);   // (b0, b2) (b1, b3)
    const uint32_t e2 = __byte_perm(v1, 0u, 0x4240), o2 = __byte_perm(v1, 0u, 0x4341);   // (b4, b6) (b5, b7)
    const uint32_t e1 = __byte_perm(e0, v1, 0x3412);                                     // (b2, b4)
    const uint32_t e3 = prmt(e2, w2, sel_last);                                          // (b6, b8): b8 = byte k & 3 of w2
    const uint32_t t02 = g8 * e0 + 0x00800080u + f8 * o0, t13 = g8 * o0 + 0x00800080u + f8 * e1;
    const uint32_t t46 = g8 * e2 + 0x00800080u + f8 * o2, t57 = g8 * o2 + 0x00800080u + f8 * e3;
    out[0] = __byte_perm(t02, t13, 0x7351);
    out[1] = __byte_perm(t46, t57, 0x7351);
}

// 4x4 byte transpose: c[j] byte i = r[i] byte j
__device__ __forceinline__ void transpose4x4_u8_s(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t& c0,
                                                  uint32_t& c1, uint32_t& c2, uint32_t& c3) {
    const uint32_t u = __byte_perm(r0, r1, 0x5140), v = __byte_perm(r2, r3, 0x5140);
    const uint32_t u2 = __byte_perm(r0, r1, 0x7362), v2 = __byte_perm(r2, r3, 0x7362);
    c0 = __byte_perm(u, v, 0x5410);
    c1 = __byte_perm(u, v, 0x7632);
    c2 = __byte_perm(u2, v2, 0x5410);
    c3 = __byte_perm(u2, v2, 0x7632);
}
// 8x8 byte transpose: in r[i] = (lo, hi) words of row i; out c[j] = (lo, hi) words of column j
__device__ __forceinline__ void transpose8x8_u8(const uint32_t (&r)[8][2], uint32_t (&c)[8][2]) {
    transpose4x4_u8_s(r[0][0], r[1][0], r[2][0], r[3][0], c[0][0], c[1][0], c[2][0], c[3][0]);
    transpose4x4_u8_s(r[4][0], r[5][0], r[6][0], r[7][0], c[0][1], c[1][1], c[2][1], c[3][1]);
    transpose4x4_u8_s(r[0][1], r[1][1], r[2][1], r[3][1], c[4][0], c[5][0], c[6][0], c[7][0]);
    transpose4x4_u8_s(r[4][1], r[5][1], r[6][1], r[7][1], c[4][1], c[5][1], c[6][1], c[7][1]);
}
// lines 4h .. 4h+3 of an 8-line tile as a strip of strip_cost_packed()
__device__ __forceinline__ const uint32_t (&half_tile(const uint32_t (&t)[8][2], int h))[4][2] {
    return *reinterpret_cast<const uint32_t(*)[4][2]>(&t[4 * h]);
}

template <int N, int COST>
__global__ void __launch_bounds__(LineCfg<N>::WARPS * 32, (N > 8 || COST == NH_COST_SAD) ? LineCfg<N>::PER_SM : 4) search_lines_kernel(const SearchArgs a) {
    using C = LineCfg<N>;
    constexpr int SEG = C::SEG, T = C::T, G = C::G, S = Log2<N>::v, PB = C::PB;
    extern __shared__ __align__(16) uint32_t smem_w0[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the scan-line table, copied from the constant bank (lines 0 .. N-1 of every row): indexed constant loads in
    // the hot loop compete with instruction fetch (ncu: no-instruction and short-scoreboard stalls), broadcast
    // shared-memory loads do not
    int4* s_tab = reinterpret_cast<int4*>(smem_w0);
    int* s_k4 = reinterpret_cast<int*>(smem_w0 + 17 * N * 4);
    for (int i = threadIdx.x; i < 17 * N; i += blockDim.x) {
        s_tab[i] = kc_line_tab.e[i / N][i % N];
        s_k4[i] = kc_line_tab.k4[i / N][i % N];
    }
    __syncthreads();
    uint32_t* const smem_w = smem_w0 + C::TAB_WORDS;
    int* acc = reinterpret_cast<int*>(smem_w + C::WARPS * C::WARP_WORDS) + threadIdx.x;   // + CTA size * candidate position
    uint32_t* wbase = smem_w + warp * C::WARP_WORDS;
    const int bi = lane / SEG, sg = lane % SEG;      // block of the tile, segment of the scan line
    const int px_ = 8 * sg;
    unsigned char* blk = reinterpret_cast<unsigned char*>(wbase + bi * C::BLOCK_WORDS);
    const unsigned char* tb = blk;                   // top[0 .. 2N+1]   (index 0 = corner slot)
    const unsigned char* lb = blk + PB;              // left[0 .. 2N+1]
    const unsigned char* lane_v = blk + px_;         // + uniform byte offset = the lane's window of a vertical mode
    const int* negt0 = kc_line_negt0[S - 3];
    const int bw = a.W / N;
    const int64_t n_tiles = (a.n_blocks + T - 1) / T;

    for (int64_t tile = (int64_t)blockIdx.x * C::WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * C::WARPS) {
        // ---- block coordinates (invalid blocks of a ragged tile recompute the last block; nothing is written)
        int64_t b = tile * T + bi;
        const bool valid = b < a.n_blocks;
        if (!valid) b = a.n_blocks - 1;
        const int fr = (int)(b / a.blocks_per_frame);
        const int64_t bf = b - fr * a.blocks_per_frame;   // block index inside its frame
        const int x = (int)(bf % bw) * N, y = (int)(bf / bw) * N;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int ood = 0;
        __syncwarp();   // the previous tile's arrays are no longer read

        // ---- K1: references with the substitution rules of block.py:38-55, as bytes
        const bool interior = __all_sync(0xffffffffu, x > 0 && y > 0 && x + 2 * N <= a.W && y + 2 * N <= a.H);
        constexpr int RE = T * (2 * N + 2), RI = (RE + 31) / 32;
        // all loads first, then all stores: written as one loop the compiler keeps load -> store order and the
        // tile pays RI global-memory round trips in a row (ncu: a third of the kernel's stall samples)
        int tv[RI], lv[RI];
#pragma unroll
        for (int it = 0; it < RI; ++it) {   // uniform trip count (the shuffles need every lane)
            const int e = it * 32 + lane < RE ? it * 32 + lane : RE - 1;
            const int i = e / (2 * N + 2), k = e % (2 * N + 2);
            const int xi = __shfl_sync(0xffffffffu, x, (i * SEG) & 31), yi = __shfl_sync(0xffffffffu, y, (i * SEG) & 31);
            const int16_t* srci = a.src + __shfl_sync(0xffffffffu, fr, (i * SEG) & 31) * a.frame_stride;
            const int kk = k <= 2 * N ? k : 2 * N;   // entry 2N+1: replicate-last padding (only read with weight 0)
            if (interior) {   // no substitution, no truncation: top[k] = plane[y-1][x-1+k], left[k] = plane[y-1+k][x-1]
                const int16_t* c = srci + (int64_t)(yi - 1) * a.pitch + xi - 1;
                tv[it] = __ldg(c + kk);
                lv[it] = __ldg(c + (int64_t)kk * a.pitch);
            } else {
                tv[it] = top_ref<false>(srci, a.H, a.W, a.pitch, xi, yi, 2 * N, kk);
                lv[it] = left_ref<false>(srci, a.H, a.W, a.pitch, xi, yi, 2 * N, kk);
            }
        }
#pragma unroll
        for (int it = 0; it < RI; ++it) {
            const int e = it * 32 + lane < RE ? it * 32 + lane : RE - 1;
            const int i = e / (2 * N + 2), k = e % (2 * N + 2);
            unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
            zb[k] = (unsigned char)tv[it];
            zb[PB + k] = (unsigned char)lv[it];
            ood |= tv[it] | lv[it];
        }
        if (__any_sync(0xffffffffu, (ood & ~0xff) != 0)) {   // leave the tile to the coder kernel's exact search
            if (valid && sg == 0) a.modes[b] = 0xFF;
            continue;
        }
        __syncwarp();

        // ---- projected extensions of the negative-angle modes (intra.py:180-186) + the copy of the primary array
        // behind them.  Unit = (block, orientation, group of MPG modes); horizontal mode 11 + q and vertical mode
        // 25 - q share the angle, hence the length and every projected index (constant offsets after unrolling).
        for (int u0 = 0; u0 < 2 * T * C::GP; u0 += 32) {
            const int u = u0 + lane;
            if (u < 2 * T * C::GP) {
                const int i = u / (2 * C::GP), r = u % (2 * C::GP), o = r / C::GP, g = r % C::GP;
                unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
                const unsigned char* sec = zb + (o ? PB : 0);     // vertical: secondary = left, primary = top
                uint32_t pw[C::CP / 4];
#pragma unroll
                for (int c = 0; c < C::CP / 4; ++c) pw[c] = reinterpret_cast<const uint32_t*>(zb + (o ? 0 : PB))[c];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int len = C::neg_len(14 - q);
                    const int inv = inv_angle(neg_angle_at(14 - q));
                    if (q / C::MPG == g && (q < 7 || o)) {   // mode 18 (q = 7) is vertical only
                        unsigned char* dst = zb + (o ? C::neg_t0(14 - q) : C::neg_t0(q < 7 ? q : 6));
#pragma unroll
                        for (int c = 0; c < C::CP / 4; ++c) reinterpret_cast<uint32_t*>(dst)[c] = pw[c];   // ref[t], t >= 0
#pragma unroll
                        for (int tt = 0; tt < len; ++tt) {                                               // t = -1 - tt
                            const int proj = (-tt * inv + 128) >> 8;                                     // (k+1) projection, Q3
                            dst[-1 - tt] = sec[proj > 2 * N ? 2 * N : proj];
                        }
                    }
                }
            }
        }

        // ---- DC (intra.py:46-62): top[1..N] + left[1..N], summed by the block's lanes
        int rs = 0;
#pragma unroll
        for (int k = 0; k < 2 * N / SEG; ++k) {
            const int kk = sg + k * SEG;
            rs += kk < N ? (int)tb[1 + kk] : (int)lb[1 + kk - N];
        }
#pragma unroll
        for (int off = SEG / 2; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
        const uint32_t dc4 = (uint32_t)dc_value<N>(rs) * 0x01010101u;
        __syncwarp();

        int best = 0x7fffffff;
        bool bad = false;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            // ---- the lane's pixels of this pass: ov = rows 8g .. 8g+7 x columns 8 sg .. (vertical modes, DC, planar),
            // oh = columns 8g .. 8g+7 x rows 8 sg .. transposed (horizontal modes: scan line = image column)
            uint32_t ov[8][2], oh[8][2];
            {
                uint32_t raw = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(srcf + (int64_t)(y + 8 * g + j) * a.pitch + x + px_));
                    raw |= v.x | v.y | v.z | v.w;
                    ov[j][0] = __byte_perm(v.x, v.y, 0x6420);
                    ov[j][1] = __byte_perm(v.z, v.w, 0x6420);
                }
                if (N == 8) {
                    transpose8x8_u8(ov, oh);
                } else {
                    uint32_t rw[8][2];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(srcf + (int64_t)(y + px_ + i) * a.pitch + x + 8 * g));
                        raw |= v.x | v.y | v.z | v.w;
                        rw[i][0] = __byte_perm(v.x, v.y, 0x6420);
                        rw[i][1] = __byte_perm(v.z, v.w, 0x6420);
                    }
                    transpose8x8_u8(rw, oh);
                }
                if (__any_sync(0xffffffffu, (raw & 0xFF00FF00u) != 0)) { bad = true; break; }
            }
            const bool first = g == 0, last = g == G - 1;
            // a candidate's cost so far -> its key in the last pass
            auto settle = [&](int c, int pos) {
                if (G > 1) {
                    if (!first) c += acc[C::WARPS * 32 * pos];
                    if (!last) acc[C::WARPS * 32 * pos] = c;
                }
                if (last) {
#pragma unroll
                    for (int off = SEG / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
                    const int key = (c << 6) | pos;
                    best = key < best ? key : best;
                }
            };
            {   // position 0: DC
                const uint32_t d[4][2] = {{dc4, dc4}, {dc4, dc4}, {dc4, dc4}, {dc4, dc4}};
                settle(strip_cost_packed<2>(d, half_tile(ov, 0), COST) + strip_cost_packed<2>(d, half_tile(ov, 1), COST), 0);
            }
            {   // position 1: planar (intra.py:109-111), two samples per multiply-add chain; the weights carry a
                // factor 2^(7-S) so that the sample is the high byte of its 16-bit lane (max 65408)
                constexpr uint32_t SCL = 1u << (7 - S);
                const uint32_t tr = (uint32_t)tb[N + 1], bl = (uint32_t)lb[N + 1];
                uint32_t kc[4], c1[4], zt[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t X = (uint32_t)(px_ + 2 * i);
                    c1[i] = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 2) - X) << 16)) * SCL;
                    kc[i] = tr * (((X + 1) | ((X + 2) << 16)) * SCL);
                    zt[i] = (uint32_t)tb[1 + px_ + 2 * i] | ((uint32_t)tb[2 + px_ + 2 * i] << 16);   // (top[1+X], top[2+X])
                }
                int c = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t pr[4][2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int yy = 8 * g + 4 * h + j;
                        const uint32_t ly = (uint32_t)lb[1 + yy];
                        const uint32_t vy = (uint32_t)(N - 1 - yy) * SCL;
                        const uint32_t by = ((uint32_t)(yy + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
                        uint32_t t[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) t[i] = ly * c1[i] + kc[i] + vy * zt[i] + by;
                        pr[j][0] = __byte_perm(t[0], t[1], 0x7531);
                        pr[j][1] = __byte_perm(t[2], t[3], 0x7531);
                    }
                    c += strip_cost_packed<2>(pr, half_tile(ov, h), COST);
                }
                settle(c, 1);
            }
            // ---- positions 2..34: angular modes (intra.py:116-207): horizontal mode m together with its mirror,
            // vertical mode 36 - m (same angle: same table row).  Mode 18 is its own mirror: its horizontal half
            // is computed and dropped.
#pragma unroll 1
            for (int r = 0; r <= 16; ++r) {
                // negative angles (r > 8) read the mode's own array (projection + copy), the others the primary array
                const unsigned char* bv = lane_v + (r > 8 ? negt0[23 - r] : 0);
                const unsigned char* bh = lane_v + (r > 8 ? negt0[r - 9] : PB);
                const int4* tab = s_tab + r * N + 8 * g;
                const int* tk4 = s_k4 + r * N + 8 * g;
                int cv = 0, ch = 0;
                if (r == 0 || r == 8 || r == 16) {   // modes 2 / 34, 10 / 26, 18: every fraction is 0
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t pv[4][2], ph[4][2];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k4 = tk4[4 * h + j];
                            const uint32_t sh = (uint32_t)tab[4 * h + j].x;
                            copy_line_w<2>(reinterpret_cast<const uint32_t*>(bv + k4), sh, pv[j]);
                            copy_line_w<2>(reinterpret_cast<const uint32_t*>(bh + k4), sh, ph[j]);
                        }
                        cv += strip_cost_packed<2>(pv, half_tile(ov, h), COST);
                        ch += strip_cost_packed<2>(ph, half_tile(oh, h), COST);
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t pv[4][2], ph[4][2];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k4 = tk4[4 * h + j];
                            const int4 e = tab[4 * h + j];
                            predict_line8(reinterpret_cast<const uint32_t*>(bv + k4), (uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z,
                                          (uint32_t)e.w, pv[j]);
                            predict_line8(reinterpret_cast<const uint32_t*>(bh + k4), (uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z,
                                          (uint32_t)e.w, ph[j]);
                        }
                        cv += strip_cost_packed<2>(pv, half_tile(ov, h), COST);
                        ch += strip_cost_packed<2>(ph, half_tile(oh, h), COST);
                    }
                }
                settle(cv, 34 - r);
                if (r < 16) settle(ch, r + 2);
            }
        }
        if (bad) {
            if (valid && sg == 0) a.modes[b] = 0xFF;
            continue;
        }
        if (valid && sg == 0) {
            a.modes[b] = (uint8_t)mode_of_key(best);
            if (a.costs) a.costs[b] = best >> 6;
        }
    }
}

}  // namespace nh
