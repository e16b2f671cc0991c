// nh_search4.cuh -- K7 search stage for 8-bit planes with the SATD cost (sum of satd_4x4, metrics.py:29-43), fourth
// generation: the 4x4 Hadamard transforms run on the tensor cores.
//
// ncu on the packed SATD of the earlier search kernels (profiles/r1_searchA8satd_ncu_summary.json): 345 thread
// instructions per pixel against 162 for SAD, ALU pipe 81 % -- the Hadamard butterflies of 33 candidates on biased
// 16-bit pairs cost more than predicting the candidates.  But sum |H d H^T| of a 4x4 sub-block is sum |(H (x) H) d|
// over its 16 samples: a 16 x 16 matrix of +-1 applied to a 16-vector, i.e. one row of an m16n8k16 MMA.
//   * A row of the A operand = the 16 predicted samples of one sub-block of one candidate, as the f16 numbers
//     1024 + p (bit pattern 0x6400 | p: ONE PRMT turns two interpolated samples into an operand register -- it
//     replaces the PRMT that packed them as bytes for VABSDIFF4); B = H (x) H as +-1.0 (two n8 halves: 2 HMMA per 16
//     sub-blocks).  Products and sums are small integers: exact in f16 x f16 -> f32.
//   * The transform is linear: (H (x) H)(p - o) = (H (x) H) p - (H (x) H) o.  The second term does not depend on the
//     candidate, so it is computed once per block (the same MMAs on the original samples, negated B) and kept as
//     the START VALUE of the accumulator; the bias 16 * 1024 of the first coefficient cancels the same way.  The
//     cost of a candidate's sub-blocks is then sum |d| over the accumulators: 8 FADD |.| per lane and MMA pair, on
//     the FMA pipe -- no subtraction, no butterflies, nothing on the ALU pipe that bounds the search.
//   * An MMA row spreads its 16 k-slots over the four lanes of a quad, so the four lanes of a quad work on four
//     CONSECUTIVE SCAN LINES of the same 8-sample segment (lane = (segment unit g, line t)); a lane's 8 samples are
//     the row of the left sub-block (MMA row g) and of the right one (MMA row g + 8).  The k-slot <-> sample map is
//     free (B's rows are permuted to match): slots 2t, 2t+1 hold samples (0, 2) of line t, slots 2t+8, 2t+9 samples
//     (1, 3) -- the pairs the interpolation of nh_search.cuh produces.
//   * Scan-line positions differ between the lanes of a quad; they come from the shared-memory table of
//     nh_search2.cuh as four neighbouring entries per load (conflict-free).
// Horizontal modes are evaluated on the transposed block (the sum over a sub-block is transposition invariant),
// mirror pairs share a table row, candidate order / tie rule / hand-back as in nh_search.cuh.
#pragma once
#include <type_traits>

#include "nh_mma.cuh"
#include "nh_search2.cuh"

namespace nh {

template <int N>
struct QuadCfg {
    using L = LineCfg<N>;
    static constexpr int T = N == 8 ? 8 : (N == 16 ? 2 : 1);   // blocks per warp tile
    static constexpr int LPB = 32 / T;                 // lanes per block
    static constexpr int GL = 8 / T;                   // quads per block
    static constexpr int SEG = N / 8;                  // 8-sample segments per scan line
    static constexpr int QS = GL / SEG;                // groups of 4 scan lines a block covers per step
    static constexpr int STEPS = (N / 4) / QS;         // 2, 2, 4
    static constexpr int PB = L::PB, CP = L::CP, BLOCK_WORDS = L::BLOCK_WORDS;
    static constexpr int WARP_WORDS = T * BLOCK_WORDS;
    static constexpr int GP = 16 / T > 8 ? 8 : 16 / T; // build: groups of modes per orientation and block
    static constexpr int MPG = 8 / GP;
    static constexpr int WARPS = 4;
    static constexpr int TAB_WORDS = 17 * N * 5;
    static constexpr int SMEM_BYTES = (WARPS * WARP_WORDS + TAB_WORDS) * 4;
    static constexpr int PER_SM = N == 32 ? 3 : 5;
};

// two interpolated samples (high bytes of the 16-bit lanes of t) as the f16 pair (1024 + s0, 1024 + s1)
__device__ __forceinline__ uint32_t hi_bytes_f16(uint32_t t) { return __byte_perm(t, 0x64646464u, 0x4341); }
// bytes (0, 2) / (1, 3) of a packed word as f16 pairs
__device__ __forceinline__ uint32_t even_bytes_f16(uint32_t w) { return __byte_perm(w, 0x64646464u, 0x4240); }
__device__ __forceinline__ uint32_t odd_bytes_f16(uint32_t w) { return __byte_perm(w, 0x64646464u, 0x4341); }

// One scan line of 8 predicted samples as the A-operand registers of the lane:
// {(s0, s2), (s4, s6), (s1, s3), (s5, s7)} = rows g / g+8 slots 2t.., rows g / g+8 slots 2t+8..
__device__ __forceinline__ uint4 predict_line8_f16(const uint32_t* wp, uint32_t sh, uint32_t sel_last, uint32_t f8, uint32_t g8) {
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
    const uint32_t e0 = __byte_perm(v0, 0u, 0x4240), o0 = __byte_perm(v0, 0u, 0x4341);
    const uint32_t e2 = __byte_perm(v1, 0u, 0x4240), o2 = __byte_perm(v1, 0u, 0x4341);
    const uint32_t e1 = __byte_perm(e0, v1, 0x3412);
    const uint32_t e3 = prmt(e2, w2, sel_last);
    const uint32_t t02 = g8 * e0 + 0x00800080u + f8 * o0, t13 = g8 * o0 + 0x00800080u + f8 * e1;
    const uint32_t t46 = g8 * e2 + 0x00800080u + f8 * o2, t57 = g8 * o2 + 0x00800080u + f8 * e3;
    return make_uint4(hi_bytes_f16(t02), hi_bytes_f16(t46), hi_bytes_f16(t13), hi_bytes_f16(t57));
}
__device__ __forceinline__ uint4 copy_line8_f16(const uint32_t* wp, uint32_t sh) {
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
    return make_uint4(even_bytes_f16(v0), even_bytes_f16(v1), odd_bytes_f16(v0), odd_bytes_f16(v1));
}
__device__ __forceinline__ uint4 packed_line8_f16(uint32_t w0, uint32_t w1) {
    return make_uint4(even_bytes_f16(w0), even_bytes_f16(w1), odd_bytes_f16(w0), odd_bytes_f16(w1));
}

// sum |(H (x) H) a + c| over the lane's share of the 16 x 16 outputs
__device__ __forceinline__ float satd_mma(const uint4& av, const uint32_t (&hb)[2][2], const float (&c)[8]) {
    float d0[4], d1[4];
    hmma16816(d0, av, hb[0][0], hb[0][1], c[0], c[1], c[2], c[3]);
    hmma16816(d1, av, hb[1][0], hb[1][1], c[4], c[5], c[6], c[7]);
    return ((fabsf(d0[0]) + fabsf(d0[1])) + (fabsf(d0[2]) + fabsf(d0[3]))) +
           ((fabsf(d1[0]) + fabsf(d1[1])) + (fabsf(d1[2]) + fabsf(d1[3])));
}

template <int N, int OCC = QuadCfg<N>::PER_SM>
__global__ void __launch_bounds__(QuadCfg<N>::WARPS * 32, OCC) search_quad_kernel(const SearchArgs a) {
    using C = QuadCfg<N>;
    constexpr int T = C::T, S = Log2<N>::v, PB = C::PB, STEPS = C::STEPS, LPB = C::LPB;
    extern __shared__ __align__(16) uint32_t smem_w0[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int4* s_tab = reinterpret_cast<int4*>(smem_w0);
    int* s_k4 = reinterpret_cast<int*>(smem_w0 + 17 * N * 4);
    for (int i = threadIdx.x; i < 17 * N; i += blockDim.x) {
        s_tab[i] = kc_line_tab.e[i / N][i % N];
        s_k4[i] = kc_line_tab.k4[i / N][i % N];
    }
    __syncthreads();
    uint32_t* const wbase = smem_w0 + C::TAB_WORDS + warp * C::WARP_WORDS;
    const int t = lane & 3, g = lane >> 2;
    const bool odd = lane & 1;
    const int bi = g / C::GL, gl = g % C::GL;          // block of the tile, quad of the block
    const int sg = gl % C::SEG, ql = gl / C::SEG;      // segment of the scan line, group of 4 lines within a step
    const int px_ = 8 * sg;
    const int y0 = 4 * ql + t;                         // the lane's scan line in step s: y0 + 4 QS s
    unsigned char* blk = reinterpret_cast<unsigned char*>(wbase + bi * C::BLOCK_WORDS);
    const unsigned char* tb = blk;
    const unsigned char* lb = blk + PB;
    const unsigned char* lane_v = blk + px_;
    const int* negt0 = kc_line_negt0[S - 3];
    const int bw = a.W / N;
    const int64_t n_tiles = (a.n_blocks + T - 1) / T;

    // B operand: column n = coefficient (u, v) = (n >> 2, n & 3) of H d H^T, row k = sample (row, x) of the sub-block
    // with row = (k & 7) >> 1, x = 2 (k & 1) + (k >> 3); H = rows ++++, ++--, +--+, +-+- (metrics.py:35-40).
    // hb = +(H (x) H) for the candidates, hn = -(H (x) H) for the original samples.
    uint32_t hb[2][2], hn[2][2];
    {
        auto hneg = [](int i, int j) { return i == 1 ? j >= 2 : (i == 2 ? (j == 1 || j == 2) : (i == 3 ? (j & 1) : 0)); };
        auto entry = [&](int n, int k) -> uint32_t {
            const int row = (k & 7) >> 1, x = 2 * (k & 1) + (k >> 3);
            return (hneg(n >> 2, row) ^ hneg(n & 3, x)) ? 0xBC00u : 0x3C00u;
        };
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                hb[m][h] = entry(g + 8 * m, 2 * t + 8 * h) | (entry(g + 8 * m, 2 * t + 1 + 8 * h) << 16);
                hn[m][h] = hb[m][h] ^ 0x80008000u;
            }
    }

    for (int64_t tile = (int64_t)blockIdx.x * C::WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * C::WARPS) {
        int64_t b = tile * T + bi;
        const bool valid = b < a.n_blocks;
        if (!valid) b = a.n_blocks - 1;
        const int fr = (int)(b / a.blocks_per_frame);
        const int64_t bf = b - fr * a.blocks_per_frame;
        const int x = (int)(bf % bw) * N, y = (int)(bf / bw) * N;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int ood = 0;
        __syncwarp();   // the previous tile's arrays are no longer read

        // ---- the lane's original samples: scan line y0 + 4 QS s of the block (vertical modes, DC, planar) and of the
        // transposed block (horizontal modes: scan line = image column), 8 samples from px_ each.  Issued before the
        // reference gather so that the loads overlap it.
        uint4 ovr[STEPS];
        int ohr[STEPS][8];
#pragma unroll
        for (int s = 0; s < STEPS; ++s) {
            const int yy = y0 + 4 * C::QS * s;
            ovr[s] = __ldg(reinterpret_cast<const uint4*>(srcf + (int64_t)(y + yy) * a.pitch + x + px_));
#pragma unroll
            for (int j = 0; j < 8; ++j) ohr[s][j] = __ldg(srcf + (int64_t)(y + px_ + j) * a.pitch + x + yy);
        }

        // ---- K1: references with the substitution rules of block.py:38-55, as bytes (loads, then stores)
        const bool interior = __all_sync(0xffffffffu, x > 0 && y > 0 && x + 2 * N <= a.W && y + 2 * N <= a.H);
        constexpr int RE = T * (2 * N + 2), RI = (RE + 31) / 32;
        int tv[RI], lv[RI];
#pragma unroll
        for (int it = 0; it < RI; ++it) {
            const int e = it * 32 + lane < RE ? it * 32 + lane : RE - 1;
            const int i = e / (2 * N + 2), k = e % (2 * N + 2);
            const int xi = __shfl_sync(0xffffffffu, x, (i * LPB) & 31), yi = __shfl_sync(0xffffffffu, y, (i * LPB) & 31);
            const int16_t* srci = a.src + __shfl_sync(0xffffffffu, fr, (i * LPB) & 31) * a.frame_stride;
            const int kk = k <= 2 * N ? k : 2 * N;
            if (interior) {
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

        // ---- start values of the accumulators: -(H (x) H) applied to the original samples (f16: 1024 + o)
        float cin[STEPS][2][8];
        {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                ood |= (int)((ovr[s].x | ovr[s].y | ovr[s].z | ovr[s].w) & 0xFF00FF00u) ? 0x100 : 0;
                const uint4 av = packed_line8_f16(__byte_perm(ovr[s].x, ovr[s].y, 0x6420), __byte_perm(ovr[s].z, ovr[s].w, 0x6420));
                uint32_t h0 = 0, h1 = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ood |= ohr[s][j] | ohr[s][4 + j];
                    h0 |= (uint32_t)(ohr[s][j] & 0xff) << (8 * j);
                    h1 |= (uint32_t)(ohr[s][4 + j] & 0xff) << (8 * j);
                }
                const uint4 ah = packed_line8_f16(h0, h1);
                float d0[4], d1[4];
                hmma16816(d0, av, hn[0][0], hn[0][1], z[0], z[1], z[2], z[3]);
                hmma16816(d1, av, hn[1][0], hn[1][1], z[0], z[1], z[2], z[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) { cin[s][0][i] = d0[i]; cin[s][0][4 + i] = d1[i]; }
                hmma16816(d0, ah, hn[0][0], hn[0][1], z[0], z[1], z[2], z[3]);
                hmma16816(d1, ah, hn[1][0], hn[1][1], z[0], z[1], z[2], z[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) { cin[s][1][i] = d0[i]; cin[s][1][4 + i] = d1[i]; }
            }
        }
        if (__any_sync(0xffffffffu, (ood & ~0xff) != 0)) {   // leave the tile to the coder kernel's exact search
            if (valid && lane % LPB == 0) a.modes[b] = 0xFF;
            continue;
        }
        __syncwarp();

        // ---- projected extensions of the negative-angle modes (intra.py:180-186) + the copy of the primary array
        // behind them (as nh_search2.cuh)
        for (int u0 = 0; u0 < 2 * T * C::GP; u0 += 32) {
            const int u = u0 + lane;
            if (u < 2 * T * C::GP) {
                const int i = u / (2 * C::GP), r = u % (2 * C::GP), o = r / C::GP, gq = r % C::GP;
                unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
                const unsigned char* sec = zb + (o ? PB : 0);
                uint32_t pw[C::CP / 4];
#pragma unroll
                for (int c = 0; c < C::CP / 4; ++c) pw[c] = reinterpret_cast<const uint32_t*>(zb + (o ? 0 : PB))[c];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int len = C::L::neg_len(14 - q);
                    const int inv = inv_angle(neg_angle_at(14 - q));
                    if (q / C::MPG == gq && (q < 7 || o)) {
                        unsigned char* dst = zb + (o ? C::L::neg_t0(14 - q) : C::L::neg_t0(q < 7 ? q : 6));
#pragma unroll
                        for (int c = 0; c < C::CP / 4; ++c) reinterpret_cast<uint32_t*>(dst)[c] = pw[c];
#pragma unroll
                        for (int tt = 0; tt < len; ++tt) {
                            const int proj = (-tt * inv + 128) >> 8;
                            dst[-1 - tt] = sec[proj > 2 * N ? 2 * N : proj];
                        }
                    }
                }
            }
        }

        // ---- DC (intra.py:46-62): top[1..N] + left[1..N], summed by the block's lanes
        int rs = 0;
        {
            const int li = lane % LPB;
#pragma unroll
            for (int k = 0; k < 2 * N / LPB; ++k) {
                const int kk = li + k * LPB;
                rs += kk < N ? (int)tb[1 + kk] : (int)lb[1 + kk - N];
            }
#pragma unroll
            for (int off = LPB / 2; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
        }
        const int dc = dc_value<N>(rs);
        __syncwarp();

        int best = 0x7fffffff;
        // Two candidates at a time: the lanes of a quad first swap one of the two values with their neighbour (even lanes
        // collect the first candidate, odd lanes the second), then every lane sums ITS candidate over the block's
        // lanes and keeps its key -- half the shuffles of two full reductions; the keys meet once per tile.
        // The call for one mirror pair is issued with the work of the next one, so that its shuffle chain
        // (ncu: 30 % of the stall samples, short scoreboard) overlaps independent instructions.
        auto settle2 = [&](float v0, float v1, int pos) {
            const float give = odd ? v0 : v1, keep = odd ? v1 : v0;
            float c = keep + __shfl_xor_sync(0xffffffffu, give, 1);
#pragma unroll
            for (int off = 2; off < LPB; off <<= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
            const int key = (__float2int_rn(c) << 6) | pos;
            best = key < best ? key : best;
        };
        constexpr float kNoCand = 16777216.f / LPB;   // summed over the block's lanes: 2^24, above every cost (<= 4.2e6 at N = 32), key still positive
        float p0, p1;   // the pair waiting to be settled
        int ppos;
        {   // position 0: DC
            const uint32_t d2 = (0x6400u | (uint32_t)dc) * 0x10001u;
            const uint4 av = make_uint4(d2, d2, d2, d2);
            float c = 0.f;
#pragma unroll
            for (int s = 0; s < STEPS; ++s) c += satd_mma(av, hb, cin[s][0]);
            p0 = c;
        }
        {   // position 1: planar (intra.py:109-111); weights scaled so that the sample is the high byte of its 16-bit lane
            constexpr uint32_t SCL = 1u << (7 - S);
            const uint32_t tr = (uint32_t)tb[N + 1], bl = (uint32_t)lb[N + 1];
            uint32_t kc[4], c1[4], zt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // sample pairs (0, 2) (4, 6) (1, 3) (5, 7): the A-operand order
                const uint32_t X = (uint32_t)(px_ + (i & 1) * 4 + (i >> 1));
                c1[i] = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 3) - X) << 16)) * SCL;
                kc[i] = tr * (((X + 1) | ((X + 3) << 16)) * SCL);
                zt[i] = (uint32_t)tb[1 + X] | ((uint32_t)tb[3 + X] << 16);
            }
            float c = 0.f;
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                const int yy = y0 + 4 * C::QS * s;
                const uint32_t ly = (uint32_t)lb[1 + yy];
                const uint32_t vy = (uint32_t)(N - 1 - yy) * SCL;
                const uint32_t by = ((uint32_t)(yy + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
                uint32_t tt[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) tt[i] = ly * c1[i] + kc[i] + vy * zt[i] + by;
                c += satd_mma(make_uint4(hi_bytes_f16(tt[0]), hi_bytes_f16(tt[1]), hi_bytes_f16(tt[2]), hi_bytes_f16(tt[3])), hb, cin[s][0]);
            }
            p1 = c;
            ppos = odd;   // even lanes: position 0, odd lanes: position 1
        }
        // ---- positions 2..34: horizontal mode r + 2 together with its mirror, vertical mode 34 - r (same angle: same
        // table row).  Even lanes end up with the vertical candidate, odd lanes with the horizontal one.
        // FRAC = false: every fraction of the pair is 0 (modes 2 / 34, 10 / 26, 18).
        auto eval_pair = [&](int r, const unsigned char* bv, const unsigned char* bh, auto frac_tag, bool has_h) {
            constexpr bool FRAC = decltype(frac_tag)::value;
            settle2(p0, p1, ppos);
            const int4* tab = s_tab + r * N + y0;
            const int* tk4 = s_k4 + r * N + y0;
            float cv = 0.f, ch = 0.f;
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                const int k4 = tk4[4 * C::QS * s];
                const int4 e = tab[4 * C::QS * s];
                if (FRAC) {
                    cv += satd_mma(predict_line8_f16(reinterpret_cast<const uint32_t*>(bv + k4), (uint32_t)e.x, (uint32_t)e.y,
                                                     (uint32_t)e.z, (uint32_t)e.w), hb, cin[s][0]);
                    ch += satd_mma(predict_line8_f16(reinterpret_cast<const uint32_t*>(bh + k4), (uint32_t)e.x, (uint32_t)e.y,
                                                     (uint32_t)e.z, (uint32_t)e.w), hb, cin[s][1]);
                } else {
                    cv += satd_mma(copy_line8_f16(reinterpret_cast<const uint32_t*>(bv + k4), (uint32_t)e.x), hb, cin[s][0]);
                    ch += satd_mma(copy_line8_f16(reinterpret_cast<const uint32_t*>(bh + k4), (uint32_t)e.x), hb, cin[s][1]);
                }
            }
            p0 = cv;
            p1 = has_h ? ch : kNoCand;
            ppos = odd ? r + 2 : 34 - r;
        };
        const std::true_type frac_t{};
        const std::false_type copy_t{};
        eval_pair(0, lane_v, lane_v + PB, copy_t, true);
#pragma unroll 1
        for (int r = 1; r < 8; ++r) eval_pair(r, lane_v, lane_v + PB, frac_t, true);
        eval_pair(8, lane_v, lane_v + PB, copy_t, true);
#pragma unroll 1
        for (int r = 9; r < 16; ++r) eval_pair(r, lane_v + negt0[23 - r], lane_v + negt0[r - 9], frac_t, true);
        eval_pair(16, lane_v + negt0[7], lane_v + negt0[7], copy_t, false);   // mode 18: vertical only
        settle2(p0, p1, ppos);
        {   // even and odd lanes hold different candidates
            const int other = __shfl_xor_sync(0xffffffffu, best, 1);
            best = other < best ? other : best;
        }
        if (valid && lane % LPB == 0) {
            a.modes[b] = (uint8_t)mode_of_key(best);
            if (a.costs) a.costs[b] = best >> 6;
        }
    }
}

}  // namespace nh
