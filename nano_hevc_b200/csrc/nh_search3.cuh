// nh_search3.cuh -- K7 search stage for 8-bit planes at N = 16 / 32, third generation ("fraction-major").
//
// The first two generations interpolate every sample of every angular candidate: 33 N^2 two-tap filters per block
// (intra.py:191-207), 2 thread instructions each after packing -- the ALU pipe is the limiter of both kernels
// (ncu, profiles/r2_search_lines8_ncu_summary.json: ALU 62 %, 126 thread instructions per pixel).  But the predicted
// sample of scan line y at position x,
//     ((32 - f) ref[k + x] + f ref[k + x + 1] + 16) >> 5,   k = 1 + ((y+1) angle >> 5),  f = (y+1) angle & 31,
// depends on the mode only through (k, f): the line IS a window of the reference array filtered with fraction f.
// So this kernel filters each reference array ONCE per fraction,
//     F[o][f][t] = ((32 - f) ref_o[t] + f ref_o[t+1] + 16) >> 5,   o = top / left, f = 0 .. 31 (f = 0: the array itself),
// 2 x 31 x (2N+1) filters per block instead of 28 N^2 (8x fewer at N = 32, 4x at N = 16), and every scan line of
// every candidate becomes the fraction-0 "copy path" of the older kernels: three shared-memory words, two funnel
// shifts, two VABSDIFF4 -- with the array (f) and the window start (k) of the line read from a table.
// Negative angles extend the reference array below zero with projected samples of the other side
// (intra.py:180-186, the (k+1) projection of SURVEY Q3), which is mode-specific.  Each filtered array therefore has a
// "porch" of Z bytes in front of t = 0; before a negative-angle mirror pair is evaluated, the porches of the
// fractions it uses are filled with the mode's projected extension filtered by that fraction (the only per-mode
// interpolation left: sum over lines of the negative part, ~N^2 / 5 samples per mode).  A window then reads
// F[o][f] + k whatever the sign of k: the hot loop is the same for every angular mode.
// Lane = a strip of 4 scan lines x 8 samples as in nh_search.cuh (N = 32: one block per warp, N = 16: four), so
// SAD and the 4x4-Hadamard SATD share the code.  Candidate order, tie rule and the hand-back of tiles with samples
// outside [0, 255] are those of nh_search.cuh.
#pragma once
#include "nh_search.cuh"

namespace nh {

template <int N>
struct FracCfg {
    static_assert(N == 16 || N == 32, "fraction-major search: N = 16 / 32");
    static constexpr int SB = N * N / 32;              // strips (lanes) per block
    static constexpr int T = 32 / SB;                  // blocks per warp tile
    static constexpr int SPR = N / 8;                  // strips per row of strips
    static constexpr int PB = (2 * N + 4) / 4 * 4;    // bytes of an array from t = 0: ref[0 .. 2N+1] + the slack of three-word reads (a window starts at 2N - 7 at most)
    static constexpr int NMAX = -(1 + ((N * -26) >> 5));   // deepest negative window start of a fractional mode (angle -26)
    static constexpr int Z = (NMAX + 3) / 4 * 4;       // porch bytes in front of t = 0
    static constexpr int AS = (((Z + PB) / 4) | 1) * 4;    // array stride (odd number of words: fractions spread over the banks)
    static constexpr int HPAD = N == 32 ? 64 : 32;     // the left-based arrays start HPAD bytes late (bank spread of the build)
    static constexpr int ARR_H = 32 * AS + HPAD;       // t = 0 of F[left][0] relative to t = 0 of F[top][0]
    static constexpr int FW = (1 + ((N * 26) >> 5) + N - 1) / 4 + 1;   // words of F[o][f >= 1] a scan line can touch
    static constexpr int CP = 12;                      // bytes of the primary array copied behind a projected extension
    static constexpr int NEG0 = 64 * AS + HPAD;        // projected extensions of modes 11 .. 25 (block-relative, from byte 0)
    __host__ __device__ static constexpr int neg_len(int mi) { return -((N * neg_angle_at(mi)) >> 5); }
    __host__ __device__ static constexpr int neg_t0(int mi) {   // byte t = 0 of mode 11 + mi, from byte 0 of the block
        int off = NEG0;
        for (int m = 0; m < mi; ++m) off += (neg_len(m) + 3) / 4 * 4 + CP;
        return off + (neg_len(mi) + 3) / 4 * 4;
    }
    static constexpr int BLOCK_BYTES = neg_t0(14) + CP;
    static constexpr int BLOCK_WORDS = (BLOCK_BYTES / 4) | 1;   // odd: blocks spread over the banks
    static constexpr int WARPS = 4;
    static constexpr int WARP_WORDS = T * BLOCK_WORDS;
    static constexpr int LIST_STRIDE = N == 32 ? 120 : 32;      // porch quads of one negative angle (max 115 / 30)
    static constexpr int nproj() {                               // projected entries of modes 11 .. 25
        int s = 0;
        for (int mi = 0; mi < 15; ++mi) s += neg_len(mi);
        return s;
    }
    static constexpr int NPROJ = nproj();
    // CTA tables: scan-line table [17][N], porch lists [7][LIST_STRIDE], counts [8] + neg_t0 [16], projection list
    static constexpr int TAB_WORDS = (17 * N + 7 * LIST_STRIDE + 24 + NPROJ + 3) / 4 * 4;
    static constexpr int SMEM_BYTES = (TAB_WORDS + WARPS * WARP_WORDS) * 4;
    static constexpr int PER_SM = N == 32 ? 5 : 3;      // (6 at N = 32 fits and is no faster: the kernel is bound by the shared-memory data pipe)
};

// One predicted quad: bytes t .. t+3 of an array filtered with fraction f8 / 8, from the words holding ref[t .. t+3]
// and ref[t+4 ..] (t a multiple of 4).  Weights scaled by 8: the sample is the high byte of its 16-bit lane.
__device__ __forceinline__ uint32_t filter_quad(uint32_t w0, uint32_t w1, uint32_t f8, uint32_t g8) {
    const uint32_t e0 = __byte_perm(w0, 0u, 0x4240), o0 = __byte_perm(w0, 0u, 0x4341);   // (b0, b2) (b1, b3)
    const uint32_t e1 = __byte_perm(e0, w1, 0x3412);                                     // (b2, b4)
    const uint32_t t02 = g8 * e0 + 0x00800080u + f8 * o0, t13 = g8 * o0 + 0x00800080u + f8 * e1;
    return __byte_perm(t02, t13, 0x7351);
}

template <int N, int COST>
__global__ void __launch_bounds__(FracCfg<N>::WARPS * 32, FracCfg<N>::PER_SM) search_frac_kernel(const SearchArgs a) {
    using C = FracCfg<N>;
    constexpr int SB = C::SB, T = C::T, S = Log2<N>::v, Z = C::Z, AS = C::AS;
    extern __shared__ __align__(16) uint32_t smem_w0[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- CTA tables
    // s_tab[r][y], r = 0 .. 16 (mirror pair: horizontal mode r + 2, vertical mode 34 - r), y = scan line:
    //   (byte offset of the word holding sample k of F[.][f], from t = 0 of F[.][0]) * 32 + 8 (k & 3)
    //   -- the funnel shift takes its amount modulo 32, so the entry itself is the shift operand
    // s_list[r - 9][..], r = 9 .. 15: the porch quads of the pair, (q << 8) | 8 f: bytes -4 (q+1) .. -4 q - 1 of F[.][f]
    int* s_tab = reinterpret_cast<int*>(smem_w0);
    int* s_list = s_tab + 17 * N;
    int* s_qn = s_list + 7 * C::LIST_STRIDE;
    int* s_negt0 = s_qn + 8;
    for (int i = threadIdx.x; i < 17 * N; i += blockDim.x) {
        const int r = i / N, yy = i % N;
        const int p = (yy + 1) * intra_angle(r + 2);
        const int k = 1 + (p >> 5), f = p & 31;
        s_tab[i] = (f * AS + (k & ~3)) * 32 + 8 * (k & 3);
    }
    if (threadIdx.x < 7) {
        const int angle = intra_angle(11 + (int)threadIdx.x);
        int* lst = s_list + threadIdx.x * C::LIST_STRIDE;
        int cnt = 0;
        for (int f = 0; f < 32; ++f) {
            int nmax = 0;   // deepest window start among the lines of this fraction
            for (int yy = 0; yy < N; ++yy) {
                const int p = (yy + 1) * angle;
                if ((p & 31) == f && -(1 + (p >> 5)) > nmax) nmax = -(1 + (p >> 5));
            }
            for (int q = 0; 4 * q < nmax; ++q) lst[cnt++] = (q << 8) | (8 * f);
        }
        s_qn[threadIdx.x] = cnt;
    }
    // s_proj[..]: the projected extensions of modes 11 .. 25 (intra.py:180-186, the (k+1) projection of SURVEY Q3) as
    // byte moves inside a block: (destination << 16) | source
    int* s_proj = s_negt0 + 16;
    if (threadIdx.x >= 32 && threadIdx.x < 47) {
        const int mi = (int)threadIdx.x - 32;
        const int t0 = C::neg_t0(mi), len = C::neg_len(mi), inv = inv_angle_of_mode(11 + mi);
        s_negt0[mi] = t0;
        int start = 0;
        for (int m = 0; m < mi; ++m) start += C::neg_len(m);
        const int sec = Z + (mi >= 7 ? C::ARR_H : 0);   // vertical modes (18 .. 25) project from the left array
        for (int tt = 0; tt < len; ++tt) {              // t = -1 - tt
            const int proj = (-tt * inv + 128) >> 8;
            s_proj[start + tt] = ((t0 - 1 - tt) << 16) | (sec + (proj > 2 * N ? 2 * N : proj));
        }
    }
    __syncthreads();

    uint32_t* wbase = smem_w0 + C::TAB_WORDS + warp * C::WARP_WORDS;
    const int bi = lane / SB, st = lane % SB;        // block of the tile, strip of the block
    const int px_ = (st % C::SPR) * 8;               // base offset of the strip  (x vertical / y horizontal)
    const int py_ = (st / C::SPR) * 4;               // scan offset of the strip  (y vertical / x horizontal)
    unsigned char* blk = reinterpret_cast<unsigned char*>(wbase + bi * C::BLOCK_WORDS);
    const unsigned char* tb = blk + Z;               // top[0 .. 2N+1]   (index 0 = corner slot) = F[top][0]
    const unsigned char* lb = tb + C::ARR_H;         // left[0 .. 2N+1]  = F[left][0]
    const unsigned char* lane_v = tb + px_;          // + table offset = the lane's window in a top-based array
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

        // ---- K1: references with the substitution rules of block.py:38-55, as bytes (load phase, then store phase)
        const bool interior = __all_sync(0xffffffffu, x > 0 && y > 0 && x + 2 * N <= a.W && y + 2 * N <= a.H);
        constexpr int RE = T * (2 * N + 2), RI = (RE + 31) / 32;
        int tv[RI], lv[RI];
#pragma unroll
        for (int it = 0; it < RI; ++it) {   // uniform trip count (the shuffles need every lane)
            const int e = it * 32 + lane < RE ? it * 32 + lane : RE - 1;
            const int i = e / (2 * N + 2), k = e % (2 * N + 2);
            const int xi = __shfl_sync(0xffffffffu, x, (i * SB) & 31), yi = __shfl_sync(0xffffffffu, y, (i * SB) & 31);
            const int16_t* srci = a.src + __shfl_sync(0xffffffffu, fr, (i * SB) & 31) * a.frame_stride;
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
            unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS) + Z;
            zb[k] = (unsigned char)tv[it];
            zb[C::ARR_H + k] = (unsigned char)lv[it];
            ood |= tv[it] | lv[it];
        }

        // ---- the lane's strip, packed bytes: ov = image orientation, oh = transposed (horizontal modes)
        uint32_t ov[4][2], oh[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(srcf + (int64_t)(y + py_ + j) * a.pitch + x + px_ + 4 * q));
                ood |= (int)((v.x | v.y) & 0xFF00FF00u);
                ov[j][q] = __byte_perm(v.x, v.y, 0x6420);
            }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint2 r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[i] = __ldg(reinterpret_cast<const uint2*>(srcf + (int64_t)(y + px_ + 4 * q + i) * a.pitch + x + py_));
                ood |= (int)((r[i].x | r[i].y) & 0xFF00FF00u);
            }
            // 4x4 byte transpose: oh[j][q] byte i = row i, column j
            const uint32_t u0 = __byte_perm(r[0].x, r[1].x, 0x6240), v0 = __byte_perm(r[2].x, r[3].x, 0x6240);
            const uint32_t u1 = __byte_perm(r[0].y, r[1].y, 0x6240), v1 = __byte_perm(r[2].y, r[3].y, 0x6240);
            oh[0][q] = __byte_perm(u0, v0, 0x5410);
            oh[1][q] = __byte_perm(u0, v0, 0x7632);
            oh[2][q] = __byte_perm(u1, v1, 0x5410);
            oh[3][q] = __byte_perm(u1, v1, 0x7632);
        }
        if (__any_sync(0xffffffffu, (ood & ~0xff) != 0)) {   // leave the tile to the coder kernel's exact search
            if (valid && st == 0) a.modes[b] = 0xFF;
            continue;
        }
        __syncwarp();

        // ---- projected extensions of the negative-angle modes + the first CP bytes of the primary array behind them:
        // table-driven byte moves spread over all lanes (blocks innermost: the table entry is a broadcast), loads of a
        // batch before its stores
        {
            constexpr int NPT = T * C::NPROJ;
#pragma unroll 1
            for (int e0 = 0; e0 < NPT; e0 += 128) {
                int ent[4];
                unsigned char v[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int e = e0 + 32 * h + lane;
                    ent[h] = s_proj[e < NPT ? e / T : 0];   // (an idle lane repeats entry 0; its store is skipped)
                    const unsigned char* zb = reinterpret_cast<const unsigned char*>(wbase + (e % T) * C::BLOCK_WORDS);
                    v[h] = zb[ent[h] & 0xffff];
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int e = e0 + 32 * h + lane;
                    unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + (e % T) * C::BLOCK_WORDS);
                    if (e < NPT) zb[ent[h] >> 16] = v[h];
                }
            }
            for (int u = lane; u < 15 * T; u += 32) {
                const int i = u % T, mi = u / T;
                unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
                const uint32_t* pri = reinterpret_cast<const uint32_t*>(zb + Z + (mi >= 7 ? 0 : C::ARR_H));   // vertical: primary = top
                uint32_t* dst = reinterpret_cast<uint32_t*>(zb + s_negt0[mi]);
#pragma unroll
                for (int c = 0; c < C::CP / 4; ++c) dst[c] = pri[c];
            }
        }

        // ---- the filtered arrays F[o][1 .. 31]: unit = (block, orientation, word of the array); the spread operands of a
        // word serve all 31 fractions (4 multiply-adds, one PRMT and one store per quad)
        {
            constexpr int FP = T * 2 * C::FW;
            for (int u0 = 0; u0 < FP; u0 += 32) {
                const int u = u0 + lane;
                if (u < FP) {
                    const int i = u / (2 * C::FW), rem = u % (2 * C::FW), o = rem / C::FW, c = rem % C::FW;
                    uint32_t* ab = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS) + Z +
                                                               (o ? C::ARR_H : 0)) + c;
                    const uint32_t w0 = ab[0], w1 = ab[1];
                    const uint32_t e0 = __byte_perm(w0, 0u, 0x4240), o0 = __byte_perm(w0, 0u, 0x4341);
                    const uint32_t e1 = __byte_perm(e0, w1, 0x3412);
#pragma unroll
                    for (int f = 1; f < 32; ++f) {
                        const uint32_t f8 = 8u * f, g8 = 256u - 8u * f;
                        const uint32_t t02 = g8 * e0 + 0x00800080u + f8 * o0, t13 = g8 * o0 + 0x00800080u + f8 * e1;
                        ab[f * (AS / 4)] = __byte_perm(t02, t13, 0x7351);
                    }
                }
            }
        }

        // ---- DC (intra.py:46-62): top[1..N] + left[1..N], summed by the block's SB lanes
        int rs = 0;
#pragma unroll
        for (int k = st; k < 2 * N; k += SB) rs += k < N ? (int)tb[1 + k] : (int)lb[1 + k - N];
#pragma unroll
        for (int off = SB / 2; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
        const int dc = dc_value<N>(rs);

        auto block_sum = [&](int c) -> int {
            if constexpr (SB == 32) {
                return __reduce_add_sync(0xffffffffu, c);
            } else {
#pragma unroll
                for (int off = SB / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
                return c;
            }
        };
        uint32_t pr[4][2], prh[4][2];
        int best;
        {   // position 0: DC
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) pr[j][q] = (uint32_t)dc * 0x01010101u;
            best = block_sum(strip_cost_packed<2>(pr, ov, COST)) << 6;
        }
        {   // position 1: planar (intra.py:109-111), two samples per multiply-add chain; the weights carry a
            // factor 2^(7-S) so that the sample is the high byte of its 16-bit lane (max 65408)
            constexpr uint32_t SC = 1u << (7 - S);
            const uint32_t tr = (uint32_t)tb[N + 1], bl = (uint32_t)lb[N + 1];
            uint32_t kc[4], c1[4], zt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t X = (uint32_t)(px_ + 2 * i);
                c1[i] = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 2) - X) << 16)) * SC;
                kc[i] = tr * (((X + 1) | ((X + 2) << 16)) * SC);
                zt[i] = (uint32_t)tb[1 + px_ + 2 * i] | ((uint32_t)tb[2 + px_ + 2 * i] << 16);   // (top[1+X], top[2+X])
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = py_ + j;
                const uint32_t ly = (uint32_t)lb[1 + yy];
                const uint32_t vy = (uint32_t)(N - 1 - yy) * SC;
                const uint32_t by = ((uint32_t)(yy + 1) * bl + (uint32_t)N) * SC * 0x10001u;
                uint32_t t[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = ly * c1[i] + kc[i] + vy * zt[i] + by;
#pragma unroll
                for (int q = 0; q < 2; ++q) pr[j][q] = __byte_perm(t[2 * q], t[2 * q + 1], 0x7531);
            }
            const int key = (block_sum(strip_cost_packed<2>(pr, ov, COST)) << 6) | 1;
            best = key < best ? key : best;
        }
        __syncwarp();   // F arrays and projected extensions complete

        // ---- positions 2..34: angular modes, the mirror pair (horizontal mode r + 2, vertical mode 34 - r) together
        int4 e4n = *reinterpret_cast<const int4*>(s_tab + py_);
#pragma unroll 1
        for (int r = 0; r < 16; ++r) {
            const int4 e4 = e4n;   // the lane's four scan lines of this pair; the next pair's entry is fetched a pair ahead
            e4n = *reinterpret_cast<const int4*>(s_tab + (r + 1) * N + py_);
            if (r >= 9) {
                // porches of the pair: the projected extension of the vertical mode filtered into F[top][f], that of
                // the horizontal mode into F[left][f], for the fractions whose lines start below zero
                __syncwarp();   // the previous pair no longer reads the porches
                const int qn = s_qn[r - 9];
                const int* lst = s_list + (r - 9) * C::LIST_STRIDE;
                const int t0v = s_negt0[23 - r], t0h = s_negt0[r - 9];
                const int total = 2 * T * qn;
#pragma unroll 1
                for (int u0 = 0; u0 < total; u0 += 64) {   // two units per lane: both loads in flight together
                    uint32_t w0[2], w1[2], f8[2];
                    uint32_t* dp[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int u = u0 + 32 * h + lane < total ? u0 + 32 * h + lane : total - 1;
                        const int i = u % T, rem = u / T;       // blocks innermost: the list entry is a broadcast
                        const int o = rem >= qn ? 1 : 0, ent = lst[rem - o * qn];
                        const int q4 = 4 * (ent >> 8) + 4;
                        f8[h] = (uint32_t)ent & 0xffu;
                        unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
                        const uint32_t* ew = reinterpret_cast<const uint32_t*>(zb + (o ? t0h : t0v) - q4);
                        w0[h] = ew[0];
                        w1[h] = ew[1];
                        dp[h] = reinterpret_cast<uint32_t*>(zb + Z + (o ? C::ARR_H : 0) + (int)(f8[h] >> 3) * AS - q4);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (u0 + 32 * h + lane < total) *dp[h] = filter_quad(w0[h], w1[h], f8[h], 256u - f8[h]);
                }
                __syncwarp();
            }
            const int ent[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned char* wv = lane_v + (ent[j] >> 5);
                copy_line_w<2>(reinterpret_cast<const uint32_t*>(wv), (uint32_t)ent[j], pr[j]);
                copy_line_w<2>(reinterpret_cast<const uint32_t*>(wv + C::ARR_H), (uint32_t)ent[j], prh[j]);
            }
            const int cv = block_sum(strip_cost_packed<2>(pr, ov, COST));
            const int ch = block_sum(strip_cost_packed<2>(prh, oh, COST));
            const int keyv = (cv << 6) | (34 - r);
            const int keyh = (ch << 6) | (r + 2);
            best = keyv < best ? keyv : best;
            best = keyh < best ? keyh : best;
        }
        {   // mode 18 (angle -32, every fraction 0, its own mirror): window start px - y, below zero from the
            // mode's own array (projection + copy of top[0 .. CP))
            const int t018 = C::neg_t0(7);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = px_ - (py_ + j);
                const unsigned char* bp = (k < 0 ? blk + t018 : tb) + (k & ~3);
                copy_line_w<2>(reinterpret_cast<const uint32_t*>(bp), (uint32_t)(k & 3) * 8u, pr[j]);
            }
            const int key = (block_sum(strip_cost_packed<2>(pr, ov, COST)) << 6) | 18;
            best = key < best ? key : best;
        }
        if (valid && st == 0) {
            a.modes[b] = (uint8_t)mode_of_key(best);
            if (a.costs) a.costs[b] = best >> 6;
        }
    }
}

}  // namespace nh
