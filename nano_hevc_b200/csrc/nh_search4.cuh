// nh_search4.cuh -- K7 search stage for 8-bit planes with the SATD cost (sum of satd_4x4, metrics.py:29-43), fourth
// generation: the 4x4 Hadamard transforms run on the tensor cores.
//
// ncu on the packed SATD of the earlier search kernels (profiles/r1_searchA8satd_ncu_summary.json): 345 thread
// instructions per pixel against 162 for SAD, ALU pipe 81 % -- the Hadamard butterflies of 33 candidates on biased
// 16-bit pairs cost more than predicting the candidates.  But sum |H d H^T| of a 4x4 sub-block is sum |(H (x) H) d|
// over its 16 samples: a 16 x 16 matrix of +-1 applied to a 16-vector, i.e. one row of an m16n8k16 MMA.
//   * A row of the A operand = the 16 predicted samples of one sub-block of one candidate.  A sample s becomes the
//     f16 SUBNORMAL s * 2^-24, whose bit pattern is 0x00ss: ONE PRMT turns two interpolated samples into an operand
//     register (it replaces the PRMT that packed them as bytes for VABSDIFF4), there is no bias to remove, and
//     mma.sync keeps subnormal operands and results.
//   * The transform is linear: T(p - o) = T p - T o.  T o does not depend on the candidate: it is computed once per
//     block (the same MMAs on the original samples, negated B) and kept as the START VALUE of the accumulators.
//   * First stage, ONE BUTTERFLY SHORT, f16 accumulators: with a[0..7] the transform of rows 0, 1 of the sub-block
//     (vertical sum / difference x four horizontal coefficients) and c[0..7] that of rows 2, 3, the 16 coefficients
//     are a[n] +- c[n] and |a + c| + |a - c| = 2 max(|a|, |c|).  a and c are sums of 8 sample differences, |.| <=
//     2040: exact in f16 (the full coefficients, <= 4080, are not), so the accumulators are 2 registers per MMA and
//     max(|a|, |c|) is one HMNMX2 with |.| operand modifiers per two values.
//   * Second stage: a SELECTOR MMA (f32 accumulators) adds up the 8 maxima of a sub-block -- which sit in the four
//     lanes of a quad -- over both sub-blocks of a unit, over the steps and into the accumulator column of the
//     candidate: B = 2.0 in the candidate's column.  After four mirror pairs lane t of the quad reads the finished
//     costs of pair t from its own accumulator registers.  No shuffle, no FADD chain, no per-candidate reduction
//     (the first version -- f32 accumulators, 8 FADD |.| per MMA pair, shuffle reductions per candidate -- ran at
//     232 thread instructions per pixel, this one at 170; profiles/r4_search_quad8_*).
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
    static constexpr int PER_SM = 5;                   // 96 registers; 6 CTAs (80 registers) measured +1 % / 0 / 0, the r loops unrolled 2x -0.3 / -1 / -6 % (N = 8 / 16 / 32)
};

// two interpolated samples (high bytes of the 16-bit lanes of t) as an f16 pair.  The bit pattern 0x00ss is the
// f16 SUBNORMAL s * 2^-24: no bias, every sum below stays a multiple of 2^-24 and the MMA units keep subnormals.
__device__ __forceinline__ uint32_t hi_bytes_f16(uint32_t t) { return __byte_perm(t, 0u, 0x4341); }
// bytes (0, 2) / (1, 3) of a packed word as f16 pairs
__device__ __forceinline__ uint32_t even_bytes_f16(uint32_t w) { return __byte_perm(w, 0u, 0x4240); }
__device__ __forceinline__ uint32_t odd_bytes_f16(uint32_t w) { return __byte_perm(w, 0u, 0x4341); }

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
// fraction 0: the samples are the bytes themselves, and (b0, b2) (b1, b3) are what the interpolation spreads anyway
__device__ __forceinline__ uint4 copy_line8_f16(const uint32_t* wp, uint32_t sh) {
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
    return make_uint4(even_bytes_f16(v0), even_bytes_f16(v1), odd_bytes_f16(v0), odd_bytes_f16(v1));
}
__device__ __forceinline__ uint4 packed_line8_f16(uint32_t w0, uint32_t w1) {
    return make_uint4(even_bytes_f16(w0), even_bytes_f16(w1), odd_bytes_f16(w0), odd_bytes_f16(w1));
}

// D = A(16x16, row) * B(16x8, col) + C with f16 accumulators (exact here: every partial result is an integer of at
// most 11 bits times 2^-24)
__device__ __forceinline__ void hmma16816_h(uint32_t (&d)[2], const uint4& a, uint32_t b0, uint32_t b1, uint32_t c0, uint32_t c1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%9};"
        : "=r"(d[0]), "=r"(d[1])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "r"(c0), "r"(c1));
}
// The transform one butterfly stage short.  With rows 0, 1 of a sub-block transformed into a[0..7] and rows 2, 3 into
// c[0..7] (vertical sum / difference of the two rows x the four horizontal coefficients), the 16 coefficients of
// H d H^T are a[n] + c[n] and a[n] - c[n], and |a + c| + |a - c| = 2 max(|a|, |c|).  Both halves are sums of 8 sample
// differences, |.| <= 2040: exact in f16, which the full coefficients (<= 4080) are not.  hb[0] produces a, hb[1] c;
// c4 = the start values -(transform of the original samples).  Returns max(|a|, |c|) for the lane's 2 x 2 outputs.
__device__ __forceinline__ void satd_halves(uint32_t (&m)[2], const uint4& av, const uint32_t (&hb)[2][2], const uint32_t (&c4)[4]) {
    uint32_t da[2], dc[2];
    hmma16816_h(da, av, hb[0][0], hb[0][1], c4[0], c4[1]);
    hmma16816_h(dc, av, hb[1][0], hb[1][1], c4[2], c4[3]);
    m[0] = h2_bits(__hmax2(__habs2(bits_h2(da[0])), __habs2(bits_h2(dc[0]))));
    m[1] = h2_bits(__hmax2(__habs2(bits_h2(da[1])), __habs2(bits_h2(dc[1]))));
}
// acc[row][n] += 2 * (sum over the slots k < 8 of mx[row][k] if n == column of x) + (the same for my, k >= 8): the sum
// over a sub-block's outputs AND over the four lanes of the quad in one MMA, one accumulator column per candidate
__device__ __forceinline__ void satd_sum(float (&acc)[4], const uint32_t (&mx)[2], const uint32_t (&my)[2], uint32_t sx, uint32_t sy) {
    hmma16816(acc, make_uint4(mx[0], mx[1], my[0], my[1]), sx, sy, acc[0], acc[1], acc[2], acc[3]);
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

    // B operands of the first stage.  Slot k holds sample (row, x) of the sub-block, row = (k & 7) >> 1 = the lane's t,
    // x = 2 (k & 1) + (k >> 3).  Column n = (vs, v) = (n >> 2, n & 3): vertical sum (vs = 0) or difference (vs = 1) of
    // the two rows of a half, horizontal coefficient v of H = rows ++++, ++--, +--+, +-+- (metrics.py:35-40).
    // hb[0] takes rows 0, 1 (zero in lanes t >= 2), hb[1] rows 2, 3.
    uint32_t hb[2][2];
    {
        auto hneg = [](int i, int j) { return i == 1 ? j >= 2 : (i == 2 ? (j == 1 || j == 2) : (i == 3 ? (j & 1) : 0)); };
        const int vs = g >> 2, v = g & 3;
        const bool vneg = vs && (t & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t lo = (vneg ^ (bool)hneg(v, h)) ? 0xBC00u : 0x3C00u;          // x = h
            const uint32_t hi = (vneg ^ (bool)hneg(v, 2 + h)) ? 0xBC00u : 0x3C00u;      // x = 2 + h
            const uint32_t w = lo | (hi << 16);
            hb[0][h] = t < 2 ? w : 0u;
            hb[1][h] = t < 2 ? 0u : w;
        }
    }
    const uint32_t kSel = 0x40004000u;   // (2.0, 2.0): the factor of 2 max(|a|, |c|)

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

        // ---- start values of the first-stage accumulators: minus the half transforms of the original samples
        uint32_t cin[STEPS][2][4];
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
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                uint32_t da[2], dc[2];
                hmma16816_h(da, o ? ah : av, hb[0][0] ^ 0x80008000u, hb[0][1] ^ 0x80008000u, 0u, 0u);
                hmma16816_h(dc, o ? ah : av, hb[1][0] ^ 0x80008000u, hb[1][1] ^ 0x80008000u, 0u, 0u);
                cin[s][o][0] = da[0]; cin[s][o][1] = da[1]; cin[s][o][2] = dc[0]; cin[s][o][3] = dc[1];
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
        // Accumulator columns 2j, 2j+1 collect the vertical / horizontal candidate of the j-th mirror pair of a group of
        // four; lane t of a quad ends up with columns 2t, 2t+1 of the unit's left (row g) and right (row g+8) sub-block
        // column: the cost of ONE pair, already summed over the quad and the steps.  The costs are scaled by 2^-24.
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        auto take = [&](int pos_v, int pos_h, bool has_v, bool has_h) {
            float cv = acc[0] + acc[2], ch = acc[1] + acc[3];
#pragma unroll
            for (int off = 4; off < LPB; off <<= 1) {
                cv += __shfl_xor_sync(0xffffffffu, cv, off);
                ch += __shfl_xor_sync(0xffffffffu, ch, off);
            }
            const int kv = (__float2int_rn(cv * 16777216.f) << 6) | pos_v;
            const int kh = (__float2int_rn(ch * 16777216.f) << 6) | pos_h;
            if (has_v) best = kv < best ? kv : best;
            if (has_h) best = kh < best ? kh : best;
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        };
        // ---- positions 2..34: horizontal mode r + 2 together with its mirror, vertical mode 34 - r (same angle: same
        // table row).  FRAC = false: every fraction of the pair is 0 (modes 2 / 34, 10 / 26, 18).
        auto eval_pair = [&](int r, const unsigned char* bv, const unsigned char* bh, auto frac_tag) {
            constexpr bool FRAC = decltype(frac_tag)::value;
            const int4* tab = s_tab + r * N + y0;
            const int* tk4 = s_k4 + r * N + y0;
            const uint32_t sx = g == 2 * (r & 3) ? kSel : 0u, sy = g == 2 * (r & 3) + 1 ? kSel : 0u;
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                const int k4 = tk4[4 * C::QS * s];
                const int4 e = tab[4 * C::QS * s];
                uint32_t mv[2], mh[2];
                if (FRAC) {
                    satd_halves(mv, predict_line8_f16(reinterpret_cast<const uint32_t*>(bv + k4), (uint32_t)e.x, (uint32_t)e.y,
                                                      (uint32_t)e.z, (uint32_t)e.w), hb, cin[s][0]);
                    satd_halves(mh, predict_line8_f16(reinterpret_cast<const uint32_t*>(bh + k4), (uint32_t)e.x, (uint32_t)e.y,
                                                      (uint32_t)e.z, (uint32_t)e.w), hb, cin[s][1]);
                } else {
                    satd_halves(mv, copy_line8_f16(reinterpret_cast<const uint32_t*>(bv + k4), (uint32_t)e.x), hb, cin[s][0]);
                    satd_halves(mh, copy_line8_f16(reinterpret_cast<const uint32_t*>(bh + k4), (uint32_t)e.x), hb, cin[s][1]);
                }
                satd_sum(acc, mv, mh, sx, sy);
            }
        };
        const std::true_type frac_t{};
        const std::false_type copy_t{};
        eval_pair(0, lane_v, lane_v + PB, copy_t);
#pragma unroll 1
        for (int r = 1; r < 4; ++r) eval_pair(r, lane_v, lane_v + PB, frac_t);
        take(34 - t, 2 + t, true, true);
#pragma unroll 1
        for (int r = 4; r < 8; ++r) eval_pair(r, lane_v, lane_v + PB, frac_t);
        take(30 - t, 6 + t, true, true);
        eval_pair(8, lane_v, lane_v + PB, copy_t);
#pragma unroll 1
        for (int r = 9; r < 12; ++r) eval_pair(r, lane_v + negt0[23 - r], lane_v + negt0[r - 9], frac_t);
        take(26 - t, 10 + t, true, true);
#pragma unroll 1
        for (int r = 12; r < 16; ++r) eval_pair(r, lane_v + negt0[23 - r], lane_v + negt0[r - 9], frac_t);
        take(22 - t, 14 + t, true, true);
        {   // last group: column 0 = mode 18 (vertical only), columns 2 / 3 = DC / planar (positions 0 / 1)
            const uint32_t s0 = g == 0 ? kSel : 0u, s2 = g == 2 ? kSel : 0u, s3 = g == 3 ? kSel : 0u;
            constexpr uint32_t SCL = 1u << (7 - S);
            const uint32_t tr = (uint32_t)tb[N + 1], bl = (uint32_t)lb[N + 1];
            uint32_t kc[4], c1[4], zt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // planar (intra.py:109-111), sample pairs (0, 2) (4, 6) (1, 3) (5, 7): the A-operand order
                const uint32_t X = (uint32_t)(px_ + (i & 1) * 4 + (i >> 1));
                c1[i] = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 3) - X) << 16)) * SCL;
                kc[i] = tr * (((X + 1) | ((X + 3) << 16)) * SCL);
                zt[i] = (uint32_t)tb[1 + X] | ((uint32_t)tb[3 + X] << 16);
            }
            const uint32_t d2 = (uint32_t)dc * 0x10001u;
            const int4* tab = s_tab + 16 * N + y0;
            const int* tk4 = s_k4 + 16 * N + y0;
            const unsigned char* b18 = lane_v + negt0[7];
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                const int yy = y0 + 4 * C::QS * s;
                uint32_t m18[2], mdc[2], mpl[2];
                satd_halves(m18, copy_line8_f16(reinterpret_cast<const uint32_t*>(b18 + tk4[4 * C::QS * s]), (uint32_t)tab[4 * C::QS * s].x),
                            hb, cin[s][0]);
                satd_halves(mdc, make_uint4(d2, d2, d2, d2), hb, cin[s][0]);
                const uint32_t ly = (uint32_t)lb[1 + yy];
                const uint32_t vy = (uint32_t)(N - 1 - yy) * SCL;
                const uint32_t by = ((uint32_t)(yy + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
                uint32_t tt[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) tt[i] = ly * c1[i] + kc[i] + vy * zt[i] + by;
                satd_halves(mpl, make_uint4(hi_bytes_f16(tt[0]), hi_bytes_f16(tt[1]), hi_bytes_f16(tt[2]), hi_bytes_f16(tt[3])), hb, cin[s][0]);
                const uint32_t zero[2] = {0u, 0u};
                satd_sum(acc, m18, zero, s0, 0u);
                satd_sum(acc, mdc, mpl, s2, s3);
            }
            take(t == 0 ? 18 : 0, 1, t < 2, t == 1);
        }
#pragma unroll
        for (int off = 1; off < 4; off <<= 1) {   // the lanes of a quad hold different candidates
            const int other = __shfl_xor_sync(0xffffffffu, best, off);
            best = other < best ? other : best;
        }
        if (valid && lane % LPB == 0) {
            a.modes[b] = (uint8_t)mode_of_key(best);
            if (a.costs) a.costs[b] = best >> 6;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// N = 4 (one 4x4 sub-block per block): the same two MMA stages, with the winner stage of search_plane_kernel<4, ., true>
// behind them.  That kernel gives a lane a whole block; an MMA row wants the block's four lines in the four lanes of a
// quad.  So for the SATD search -- and only there -- lane t of a quad predicts LINE t of each of the quad's four blocks
// (the same four predict calls per lane and mode, each against another block's arrays; the scan-line position is now
// the same for all four, one table entry per lane and mode instead of four), which is the operand layout as it
// stands: blocks 0 / 2 of the quad are rows g / g + 8 of one MMA, blocks 1 / 3 of a second one.  The finished costs of
// mirror pair t of a group come out in lane t for all four blocks; every lane keeps four running minima and the quad
// merges them once per tile.  Gather, projected extensions, DC and the winner stage are those of search_plane_kernel.
__device__ __forceinline__ void predict_line4_f16(const uint32_t* wp, uint32_t sh, uint32_t sel_last, uint32_t f8, uint32_t g8,
                                                  uint32_t& p02, uint32_t& p13) {
    const uint32_t w0 = wp[0], w1 = wp[1];
    const uint32_t v0 = __funnelshift_r(w0, w1, sh);
    const uint32_t e0 = __byte_perm(v0, 0u, 0x4240), o0 = __byte_perm(v0, 0u, 0x4341);
    const uint32_t e1 = prmt(e0, w1, sel_last);
    p02 = hi_bytes_f16(g8 * e0 + 0x00800080u + f8 * o0);
    p13 = hi_bytes_f16(g8 * o0 + 0x00800080u + f8 * e1);
}
__device__ __forceinline__ void copy_line4_f16(const uint32_t* wp, uint32_t sh, uint32_t& p02, uint32_t& p13) {
    const uint32_t v0 = __funnelshift_r(wp[0], wp[1], sh);
    p02 = even_bytes_f16(v0);
    p13 = odd_bytes_f16(v0);
}

struct Quad4Cfg {
    using C = SearchCfg<4>;
    static constexpr int SCR_WORDS = 32 * 8;   // per warp: the tile's pixels as rows (4 words) and columns (4 words) per block
    static constexpr int SMEM_BYTES = C::SMEM_BYTES + C::WARPS * SCR_WORDS * 4;
    static constexpr int PER_SM = 4;
};

template <bool CODE>
__global__ void __launch_bounds__(SearchCfg<4>::WARPS * 32, Quad4Cfg::PER_SM) search_quad4_kernel(const SearchArgs a) {
    constexpr int N = 4, S = 2, T = 32;
    using C = SearchCfg<4>;
    constexpr int PB = C::PB;
    extern __shared__ __align__(16) uint32_t smem_w[];
    int* negT0 = reinterpret_cast<int*>(smem_w + C::WARPS * C::WARP_WORDS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 15) negT0[threadIdx.x] = C::neg_t0((int)threadIdx.x);
    int4* s_tab = reinterpret_cast<int4*>(smem_w + C::WARPS * C::WARP_WORDS + 16);
    int* s_k4 = reinterpret_cast<int*>(smem_w + C::WARPS * C::WARP_WORDS + 16 + 17 * 4 * 4);
    for (int i = threadIdx.x; i < 17 * 4; i += blockDim.x) {
        s_tab[i] = kc_line_tab.e[i / 4][i % 4];
        s_k4[i] = kc_line_tab.k4[i / 4][i % 4];
    }
    __syncthreads();
    uint32_t* wbase = smem_w + warp * C::WARP_WORDS;
    uint32_t* scr = smem_w + C::WARPS * C::WARP_WORDS + 16 + C::TAB_WORDS + warp * Quad4Cfg::SCR_WORDS;
    const int t = lane & 3, g = lane >> 2, q0 = lane & ~3;
    unsigned char* blk = reinterpret_cast<unsigned char*>(wbase + lane * C::BLOCK_WORDS);   // lane = block for K1 / DC / winner
    const unsigned char* tb = blk;
    const unsigned char* lb = blk + PB;
    const unsigned char* qblk = reinterpret_cast<const unsigned char*>(wbase + q0 * C::BLOCK_WORDS);   // block j of the quad: + j * BLOCK_WORDS * 4
    const int bw = a.W / N;
    const int64_t n_tiles = (a.n_blocks + T - 1) / T;

    uint32_t hb[2][2];   // first-stage B operands, as in search_quad_kernel
    {
        auto hneg = [](int i, int j) { return i == 1 ? j >= 2 : (i == 2 ? (j == 1 || j == 2) : (i == 3 ? (j & 1) : 0)); };
        const int vs = g >> 2, v = g & 3;
        const bool vneg = vs && (t & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t lo = (vneg ^ (bool)hneg(v, h)) ? 0xBC00u : 0x3C00u;
            const uint32_t hi = (vneg ^ (bool)hneg(v, 2 + h)) ? 0xBC00u : 0x3C00u;
            const uint32_t w = lo | (hi << 16);
            hb[0][h] = t < 2 ? w : 0u;
            hb[1][h] = t < 2 ? 0u : w;
        }
    }
    const uint32_t kSel = 0x40004000u;

    for (int64_t tile = (int64_t)blockIdx.x * C::WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * C::WARPS) {
        int64_t b = tile * T + lane;
        const bool valid = b < a.n_blocks;
        if (!valid) b = a.n_blocks - 1;
        const int fr = (int)(b / a.blocks_per_frame);
        const int64_t bf = b - fr * a.blocks_per_frame;
        const int x = (int)(bf % bw) * N, y = (int)(bf / bw) * N;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int ood = 0;
        __syncwarp();   // the previous tile's arrays and scratch are no longer read

        // ---- K1: references with the substitution rules of block.py:38-55, as bytes (loads, then stores)
        const bool interior = __all_sync(0xffffffffu, x > 0 && y > 0 && x + 2 * N <= a.W && y + 2 * N <= a.H);
        constexpr int RE = T * (2 * N + 2), RI = (RE + 31) / 32;
        int tv[RI], lv[RI];
#pragma unroll
        for (int it = 0; it < RI; ++it) {
            const int e = it * 32 + lane;
            const int i = e / (2 * N + 2), k = e % (2 * N + 2);
            const int xi = __shfl_sync(0xffffffffu, x, i), yi = __shfl_sync(0xffffffffu, y, i);
            const int16_t* srci = a.src + __shfl_sync(0xffffffffu, fr, i) * a.frame_stride;
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
            const int e = it * 32 + lane;
            const int i = e / (2 * N + 2), k = e % (2 * N + 2);
            unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
            zb[k] = (unsigned char)tv[it];
            zb[PB + k] = (unsigned char)lv[it];
            ood |= tv[it] | lv[it];
        }

        // ---- the lane's block as packed rows (ov: the winner stage needs them) and packed columns (oh)
        uint32_t ov[4][1], oh[4][1];
        {
            uint2 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                r[j] = __ldg(reinterpret_cast<const uint2*>(srcf + (int64_t)(y + j) * a.pitch + x));
                ood |= (int)((r[j].x | r[j].y) & 0xFF00FF00u);
                ov[j][0] = __byte_perm(r[j].x, r[j].y, 0x6420);
            }
            transpose4x4_u8_s(ov[0][0], ov[1][0], ov[2][0], ov[3][0], oh[0][0], oh[1][0], oh[2][0], oh[3][0]);
        }
        const bool fast8 = !__any_sync(0xffffffffu, (ood & ~0xff) != 0);
        if (!fast8) {   // leave the tile to the coder kernel's exact search
            if (valid) a.modes[b] = 0xFF;
            if (CODE && lane == 0) atomicAdd(a.handed_back, 1u);
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            scr[lane * 8 + j] = ov[j][0];
            scr[lane * 8 + 4 + j] = oh[j][0];
        }
        __syncwarp();

        // ---- projected extensions of the negative-angle modes (as search_plane_kernel: unit = (block, orientation))
        for (int u0 = 0; u0 < 2 * T * C::GP; u0 += 32) {
            const int u = u0 + lane;
            const int i = u / (2 * C::GP), r = u % (2 * C::GP), o = r / C::GP, gq = r % C::GP;
            unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
            const unsigned char* sec = zb + (o ? PB : 0);
            uint32_t pw[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) pw[c] = reinterpret_cast<const uint32_t*>(zb + (o ? 0 : PB))[c];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int len = C::neg_len(14 - q);
                const int inv = inv_angle(neg_angle_at(14 - q));
                if (q / C::MPG == gq && (q < 7 || o)) {
                    unsigned char* dst = zb + (o ? C::neg_t0(14 - q) : C::neg_t0(q < 7 ? q : 6));
#pragma unroll
                    for (int c = 0; c < 2; ++c) reinterpret_cast<uint32_t*>(dst)[c] = pw[c];
#pragma unroll
                    for (int tt = 0; tt < len; ++tt) {
                        const int proj = (-tt * inv + 128) >> 8;
                        dst[-1 - tt] = sec[proj > 2 * N ? 2 * N : proj];
                    }
                }
            }
        }

        // ---- DC of the lane's block (intra.py:46-62)
        int rs = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) rs += (int)tb[1 + k] + (int)lb[1 + k];
        const int dc = dc_value<N>(rs);
        __syncwarp();

        // ---- start values: minus the half transforms of line t of the quad's four blocks (rows, and columns for the
        // horizontal modes); cin[o][p] belongs to the MMA of blocks p / p + 2
        uint32_t cin[2][2][4];
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const uint32_t wa = scr[(q0 + p) * 8 + 4 * o + t], wb = scr[(q0 + p + 2) * 8 + 4 * o + t];
                const uint4 av = make_uint4(even_bytes_f16(wa), even_bytes_f16(wb), odd_bytes_f16(wa), odd_bytes_f16(wb));
                uint32_t da[2], dcc[2];
                hmma16816_h(da, av, hb[0][0] ^ 0x80008000u, hb[0][1] ^ 0x80008000u, 0u, 0u);
                hmma16816_h(dcc, av, hb[1][0] ^ 0x80008000u, hb[1][1] ^ 0x80008000u, 0u, 0u);
                cin[o][p][0] = da[0]; cin[o][p][1] = da[1]; cin[o][p][2] = dcc[0]; cin[o][p][3] = dcc[1];
            }

        int best[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};   // blocks 0 .. 3 of the quad, this lane's candidates
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        // columns 2t / 2t+1 of blocks p (row g) and p + 2 (row g + 8): the vertical / horizontal candidate of mirror pair t
        auto take = [&](int pos_v, int pos_h, bool has_v, bool has_h) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int kva = (__float2int_rn(acc[p][0] * 16777216.f) << 6) | pos_v, kha = (__float2int_rn(acc[p][1] * 16777216.f) << 6) | pos_h;
                const int kvb = (__float2int_rn(acc[p][2] * 16777216.f) << 6) | pos_v, khb = (__float2int_rn(acc[p][3] * 16777216.f) << 6) | pos_h;
                if (has_v) { best[p] = kva < best[p] ? kva : best[p]; best[p + 2] = kvb < best[p + 2] ? kvb : best[p + 2]; }
                if (has_h) { best[p] = kha < best[p] ? kha : best[p]; best[p + 2] = khb < best[p + 2] ? khb : best[p + 2]; }
                acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
            }
        };
        // one mirror pair: line t of the four blocks, vertical (arrays at voff) and horizontal (arrays at hoff)
        auto eval_pair = [&](int r, int voff, int hoff, auto frac_tag) {
            constexpr bool FRAC = decltype(frac_tag)::value;
            const int4 e = s_tab[r * 4 + t];
            const int k4 = s_k4[r * 4 + t];
            const uint32_t sx = g == 2 * (r & 3) ? kSel : 0u, sy = g == 2 * (r & 3) + 1 ? kSel : 0u;
            uint32_t ve[4], vo[4], he[4], ho[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned char* bj = qblk + j * (C::BLOCK_WORDS * 4) + k4;
                if (FRAC) {
                    predict_line4_f16(reinterpret_cast<const uint32_t*>(bj + voff), (uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z, (uint32_t)e.w, ve[j], vo[j]);
                    predict_line4_f16(reinterpret_cast<const uint32_t*>(bj + hoff), (uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z, (uint32_t)e.w, he[j], ho[j]);
                } else {
                    copy_line4_f16(reinterpret_cast<const uint32_t*>(bj + voff), (uint32_t)e.x, ve[j], vo[j]);
                    copy_line4_f16(reinterpret_cast<const uint32_t*>(bj + hoff), (uint32_t)e.x, he[j], ho[j]);
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                uint32_t mv[2], mh[2];
                satd_halves(mv, make_uint4(ve[p], ve[p + 2], vo[p], vo[p + 2]), hb, cin[0][p]);
                satd_halves(mh, make_uint4(he[p], he[p + 2], ho[p], ho[p + 2]), hb, cin[1][p]);
                satd_sum(acc[p], mv, mh, sx, sy);
            }
        };
        // ONE instance of the pair code for all 16 mirror pairs (the fraction-0 pairs r = 0 and 8 go through the
        // interpolation with weight 0): with an instance per loop as in search_quad_kernel this kernel -- which also
        // carries the winner stage -- waited on instruction fetch (ncu: no-instruction 1.3 warps per issue, issue 56 %)
        const std::true_type frac_t{};
#pragma unroll 1
        for (int r = 0; r < 16; ++r) {
            eval_pair(r, r > 8 ? negT0[23 - r] : 0, r > 8 ? negT0[r - 9] : PB, frac_t);
            if ((r & 3) == 3) take(37 - r - t, r - 1 + t, true, true);
        }
        {   // last group: column 0 = mode 18 (vertical only), columns 2 / 3 = DC / planar (positions 0 / 1)
            const uint32_t s0 = g == 0 ? kSel : 0u, s2 = g == 2 ? kSel : 0u, s3 = g == 3 ? kSel : 0u;
            constexpr uint32_t SCL = 1u << (7 - S);
            const int4 e = s_tab[16 * 4 + t];
            const int k4 = s_k4[16 * 4 + t];
            const int off18 = negT0[7];
            uint32_t e18[4], o18[4], pl0[4], pl1[4], dq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned char* bj = qblk + j * (C::BLOCK_WORDS * 4);
                copy_line4_f16(reinterpret_cast<const uint32_t*>(bj + off18 + k4), (uint32_t)e.x, e18[j], o18[j]);
                dq[j] = (uint32_t)__shfl_sync(0xffffffffu, dc, q0 + j) * 0x10001u;
                // planar (intra.py:109-111), line t of block j, sample pairs (0, 2) and (1, 3)
                const unsigned char* tj = bj;
                const unsigned char* lj = bj + PB;
                const uint32_t tr = (uint32_t)tj[N + 1], bl = (uint32_t)lj[N + 1];
                const uint32_t ly = (uint32_t)lj[1 + t];
                const uint32_t vy = (uint32_t)(N - 1 - t) * SCL;
                const uint32_t by = ((uint32_t)(t + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
                const uint32_t c10 = (3u | (1u << 16)) * SCL, c11 = (2u | (0u << 16)) * SCL;      // (N-1-X, N-3-X), X = 0 / 1
                const uint32_t kc0 = tr * ((1u | (3u << 16)) * SCL), kc1 = tr * ((2u | (4u << 16)) * SCL);
                const uint32_t zt0 = (uint32_t)tj[1] | ((uint32_t)tj[3] << 16), zt1 = (uint32_t)tj[2] | ((uint32_t)tj[4] << 16);
                pl0[j] = hi_bytes_f16(ly * c10 + kc0 + vy * zt0 + by);
                pl1[j] = hi_bytes_f16(ly * c11 + kc1 + vy * zt1 + by);
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                uint32_t m18[2], mdc[2], mpl[2];
                const uint32_t zero[2] = {0u, 0u};
                satd_halves(m18, make_uint4(e18[p], e18[p + 2], o18[p], o18[p + 2]), hb, cin[0][p]);
                satd_halves(mdc, make_uint4(dq[p], dq[p + 2], dq[p], dq[p + 2]), hb, cin[0][p]);
                satd_halves(mpl, make_uint4(pl0[p], pl0[p + 2], pl1[p], pl1[p + 2]), hb, cin[0][p]);
                satd_sum(acc[p], m18, zero, s0, 0u);
                satd_sum(acc[p], mdc, mpl, s2, s3);
            }
            take(t == 0 ? 18 : 0, 1, t < 2, t == 1);
        }
        // ---- the quad's lanes hold different candidates: merge, then lane t keeps the key of ITS block
        int mine = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = best[j];
#pragma unroll
            for (int off = 1; off < 4; off <<= 1) {
                const int other = __shfl_xor_sync(0xffffffffu, k, off);
                k = other < k ? other : k;
            }
            mine = t == j ? k : mine;
        }
        if (valid) {
            a.modes[b] = (uint8_t)mode_of_key(mine);
            if (a.costs) a.costs[b] = mine >> 6;
        }
        if constexpr (CODE) code_winner4(a, mode_of_key(mine), dc, blk, negT0, ov, valid, b, fr, x, y);
    }
}

}  // namespace nh
