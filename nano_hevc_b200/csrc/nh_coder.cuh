// nh_coder.cuh -- the group-cooperative block coder shared by
//   * nh_fused_pipeline_modes  (any of the 35 modes from given padded references),
//   * nh_encode_frame, source neighbours        (K7: 35-mode search + winner pipeline),
//   * nh_encode_frame, reconstructed neighbours (K8: anti-diagonal wavefront).
//
// G lanes cooperate on one N x N block (G = N: several blocks per warp, throughput mode;
// G = 32: one block per warp, latency mode for the wavefront).  Shared memory per group:
//   refs   top[0..2N], left[0..2N]      (index 0 = corner slot)
//   O      N x N int16                  (original pixels, later the reconstruction)
//   M      RowsTile<N> int32            (working matrix of the separable transforms)
#pragma once
#include "nh_block.cuh"

namespace nh {

template <int N, int G>
struct CoderCfg {
    static constexpr int SB = N * N / 16;               // 4x4 sub-blocks per block
    static constexpr int SPL = SB >= G ? SB / G : 1;    // sub-blocks per lane
    static constexpr int MS = SB >= G ? 1 : G / SB;     // mode splits (lanes sharing a sub-block)
    static constexpr int SBL = SB >= G ? G : SB;        // lanes that sum one mode's cost
    static constexpr int REF_W = 2 * N + 4;             // 2N+1 entries + padding read by the packed loads
    static constexpr int NEG_W = N + 8;                 // projected extension (N) + ref[0..7] (9-sample windows)
    static constexpr int NEG_MODES = 15;                // modes 11..25 have a negative angle
    // int16 elements.  Odd word pitch spreads banks; the latency-mode coders at N >= 16 (G = 32)
    // use rows that are 16-byte aligned with an odd number of 16-byte groups instead, the layout
    // ldmatrix / stmatrix need for the tensor-core winner pipeline
    static constexpr int O_PITCH = (N >= 16 && G == 32) ? N + 8 : N + 2;
    static constexpr int REFS_BYTES = 2 * REF_W * 2;
    static constexpr int NEG_BYTES = ((NEG_MODES * NEG_W * 2 + 15) / 16) * 16;
    static constexpr int O_BYTES = ((N * O_PITCH * 2 + 15) / 16) * 16;
    static constexpr int M_BYTES = RowsTile<N>::WORDS * 4;
    static constexpr int REFS_PAD = ((REFS_BYTES + 15) / 16) * 16;
    static constexpr int GROUP_BYTES = REFS_PAD + NEG_BYTES + O_BYTES + M_BYTES;
};

struct SmemRef {  // accessor for nh::ref_at / angular_sample over shared-memory references
    const int16_t* p;
    const int16_t* s;
    int c;
    __device__ __forceinline__ int pri(int k) const { return (int)p[k]; }
    __device__ __forceinline__ int sec(int k) const { return (int)s[k]; }
    __device__ __forceinline__ int corner() const { return c; }
};

// One predicted sample of `mode` at (x, y) from padded shared-memory references.
template <int N>
__device__ __forceinline__ int predict_px(int mode, int x, int y, const int16_t* top,
                                          const int16_t* left, int corner, int dc) {
    if (mode == 1) return dc;
    if (mode == 0) return planar_px<N>(x, y, left[1 + y], top[1 + x], top[N + 1], left[N + 1]);
    const AngleInfo ai = angle_info(mode);
    SmemRef r;
    r.c = corner;
    if (ai.vertical) { r.p = top; r.s = left; return angular_sample(r, ai, x, y); }
    r.p = left; r.s = top;
    return angular_sample(r, ai, y, x);
}

// Cost of one 4x4 sub-block at (sx, sy): SAD (metrics.py:24-26) or satd_4x4 (metrics.py:29-43).
template <int N>
__device__ __forceinline__ int subblock_cost(int mode, int sx, int sy, const int (&o)[16],
                                             const int16_t* top, const int16_t* left, int corner,
                                             int dc, int cost_kind) {
    int d[16];
    if (mode >= 2) {
        // Angular: (int_part, frac) depend on the scan line only, hoist them.
        const AngleInfo ai = angle_info(mode);
        SmemRef r;
        r.c = corner;
        if (ai.vertical) { r.p = top; r.s = left; } else { r.p = left; r.s = top; }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int scan = (ai.vertical ? sy : sx) + s;
            const int p = (scan + 1) * ai.angle;
            const int ip = p >> 5, f = p & 31;
            const int b0 = (ai.vertical ? sx : sy) + 1 + ip;
            int v[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] = (k < 4 || f != 0) ? ref_at(r, b0 + k, ai.inv) : 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pv = angular_px(v[k], v[k + 1], f);
                const int e = ai.vertical ? (s * 4 + k) : (k * 4 + s);  // (y, x) of this sample
                d[e] = o[e] - pv;
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
            d[e] = o[e] - predict_px<N>(mode, sx + (e & 3), sy + (e >> 2), top, left, corner, dc);
    }
    int c = 0;
    if (cost_kind == NH_COST_SAD) {
#pragma unroll
        for (int e = 0; e < 16; ++e) c += abs(d[e]);
    } else {
        // H . d . H^T with the 4x4 +-1 Hadamard of metrics.py:36-41 (row order irrelevant
        // for the sum of absolute values).
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // columns
            int a0 = d[j] + d[4 + j], a1 = d[j] - d[4 + j];
            int a2 = d[8 + j] + d[12 + j], a3 = d[8 + j] - d[12 + j];
            d[j] = a0 + a2; d[4 + j] = a1 + a3; d[8 + j] = a0 - a2; d[12 + j] = a1 - a3;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // rows
            int a0 = d[4 * i] + d[4 * i + 1], a1 = d[4 * i] - d[4 * i + 1];
            int a2 = d[4 * i + 2] + d[4 * i + 3], a3 = d[4 * i + 2] - d[4 * i + 3];
            c += abs(a0 + a2) + abs(a1 + a3) + abs(a0 - a2) + abs(a1 - a3);
        }
    }
    return c;
}

// Exhaustive 35-mode search.  Candidate order 1, 0, 2, 3, ..., 34; the first strict
// minimum wins (DC beats planar on ties -- __main__.py:96 -- then the lowest angular mode).
// `gl` = lane within the group.  Returns key = (cost << 6) | order position, identical on
// every lane of the group.
template <int N, int G>
__device__ __forceinline__ int search_modes(int gl, const int16_t* O, const int16_t* top,
                                            const int16_t* left, int corner, int dc, int cost_kind,
                                            int it0 = 0, int it_step = 1) {
    using Cfg = CoderCfg<N, G>;
    constexpr int SBW = N / 4;
    int o[Cfg::SPL][16];
    int sx[Cfg::SPL], sy[Cfg::SPL];
#pragma unroll
    for (int i = 0; i < Cfg::SPL; ++i) {
        const int sb = (Cfg::MS == 1) ? gl + i * G : gl % Cfg::SB;
        sx[i] = (sb % SBW) * 4;
        sy[i] = (sb / SBW) * 4;
#pragma unroll
        for (int e = 0; e < 16; ++e) o[i][e] = (int)O[(sy[i] + (e >> 2)) * Cfg::O_PITCH + sx[i] + (e & 3)];
    }
    const int ms = (Cfg::MS == 1) ? 0 : gl / Cfg::SB;
    int best = 0x7fffffff;
    constexpr int ITERS = (35 + Cfg::MS - 1) / Cfg::MS;
    // (it0, it_step): several warps may share one block, each taking every it_step-th iteration
    for (int it = it0; it < ITERS; it += it_step) {
        const int pos = it * Cfg::MS + ms;          // position in the candidate order
        const bool active = pos < 35;
        const int mode = !active ? 1 : (pos == 0 ? 1 : (pos == 1 ? 0 : pos));
        int c = 0;
#pragma unroll
        for (int i = 0; i < Cfg::SPL; ++i)
            c += subblock_cost<N>(mode, sx[i], sy[i], o[i], top, left, corner, dc, cost_kind);
#pragma unroll
        for (int off = Cfg::SBL / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        const int key = active ? ((c << 6) | pos) : 0x7fffffff;
        best = key < best ? key : best;
    }
    // warp-level argmin across the mode splits
#pragma unroll
    for (int off = G / 2; off >= Cfg::SBL; off >>= 1) {
        int other = __shfl_xor_sync(0xffffffffu, best, off);
        best = other < best ? other : best;
    }
    return best;
}

// ------------------------------------------------------------------ 8-bit fast search
// When every reference and every original sample of the tile lies in [0, 255] the search runs on
// packed data: four predicted samples per register (bytes), interpolation on 16-bit pairs
//   ((32-f)*(R[k],R[k+1]) + f*(R[k+1],R[k+2]) + (16,16)) >> 5        -- two pixels per IMAD pair,
// SAD with VABSDIFF4.U8.ACC (four pixels per instruction).  Values <= 8176 never carry between the
// 16-bit lanes and never wrap int16, so the result equals intra.py:191-207 exactly.
//
// Horizontal modes (mode < 18) are evaluated on the TRANSPOSED sub-block: pred[y][x] for scan
// line x and base y is the "vertical" formula applied to the left references, and both SAD and
// sum|H d H^T| are invariant under transposition of d.
//
// Negative angles read ref[k < 0] = secondary[((k+1)*inv + 128) >> 8].  build_neg_arrays() lays
// that projection out contiguously per mode: neg[m][j] = ref_m[j - N] for j = 0 .. N+3, so a
// 5-sample window that starts below zero is a plain contiguous read.
template <int N, int G>
__device__ __forceinline__ void build_neg_arrays(int gl, const int16_t* top, const int16_t* left,
                                                 int16_t* neg) {
    using Cfg = CoderCfg<N, G>;
    constexpr int POS = Cfg::NEG_W - N;          // ref[0 .. POS-1] copied behind the projected part
    for (int t = gl; t < Cfg::NEG_MODES * N; t += G) {
        const int mi = t / N, j = t % N;         // mode 11 + mi, element k = j - N
        const int mode = 11 + mi;
        const int k = j - N;
        const int16_t* sec = mode >= 18 ? left : top;
        const int proj = ((k + 1) * inv_angle_of_mode(mode) + 128) >> 8;
        // entries below (N*angle)>>5 are never read; clamp the index so the load stays in bounds
        neg[mi * Cfg::NEG_W + j] = sec[proj > 2 * N ? 2 * N : proj];
    }
    for (int t = gl; t < Cfg::NEG_MODES * POS; t += G) {
        const int mi = t / POS, j = t % POS;
        const int16_t* pri = (11 + mi) >= 18 ? top : left;
        neg[mi * Cfg::NEG_W + N + j] = pri[j];   // pri[0] is the corner in the plane coders
    }
}

// The same for one mode only (the winner pipeline of a block whose mode is already decided).
template <int N, int G>
__device__ __forceinline__ void build_neg_array_of_mode(int gl, int mode, const int16_t* top, const int16_t* left,
                                                        int16_t* neg) {
    using Cfg = CoderCfg<N, G>;
    constexpr int POS = Cfg::NEG_W - N;
    if (mode < 11 || mode > 25) return;
    const int mi = mode - 11;
    const int16_t* sec = mode >= 18 ? left : top;
    const int16_t* pri = mode >= 18 ? top : left;
    const int inv = inv_angle_of_mode(mode);
    for (int j = gl; j < N + POS; j += G) {
        int16_t v;
        if (j < N) {
            const int proj = ((j - N + 1) * inv + 128) >> 8;
            v = sec[proj > 2 * N ? 2 * N : proj];
        } else {
            v = pri[j - N];
        }
        neg[mi * Cfg::NEG_W + j] = v;
    }
}

// SW predicted samples of one scan line as SW/2 words (bytes 0 and 2 of word i = samples 2i, 2i+1):
// R = halfword array (4-byte aligned), h = index of the first sample's ref[k], f = fraction.
template <int SW>
__device__ __forceinline__ void pred_row_pairs(const int16_t* R, int h, int f, uint32_t (&out)[SW / 2]) {
    const uint32_t* W = reinterpret_cast<const uint32_t*>(R) + (h >> 1);
    uint32_t w[SW / 2 + 1];
#pragma unroll
    for (int i = 0; i <= SW / 2; ++i) w[i] = W[i];
    const uint32_t selA = (h & 1) ? 0x5432u : 0x3210u, selB = selA + 0x2222u;
    const uint32_t g = 32u - (uint32_t)f;
#pragma unroll
    for (int i = 0; i < SW / 2; ++i) {
        const uint32_t a = __byte_perm(w[i], w[i + 1], selA);   // (R[k+2i],   R[k+2i+1])
        const uint32_t b = __byte_perm(w[i], w[i + 1], selB);   // (R[k+2i+1], R[k+2i+2])
        out[i] = (g * a + (uint32_t)f * b + 0x00100010u) >> 5;
    }
}

// Sum of satd_4x4 (metrics.py:29-43) over the 4x4 sub-blocks of a strip held as packed bytes (pr = predicted,
// o = original; 4 scan lines x WPS words).
template <int WPS>
__device__ __forceinline__ int satd_strip_packed(const uint32_t (&pr)[4][WPS], const uint32_t (&o)[4][WPS]) {
    // Sum of satd_4x4 on 16-bit pairs.  Every lane carries a bias of 0x4000, restored by the constant of
    // each three-input add, so lanes stay in [0x4000 - 2040, 0x4000 + 2040] and nothing crosses between
    // them.  Lanes = columns (0, 2) and (1, 3): the column pass is lane-wise; the row pass stops one stage
    // early because |a + c| + |a - c| = 2 max(|a|, |c|).
    constexpr uint32_t B2 = 0x40004000u;
    uint32_t acc = 0;
#pragma unroll
    for (int q = 0; q < WPS; ++q) {   // one 4x4 sub-block per packed word column
        uint32_t e[4], f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            e[j] = __byte_perm(o[j][q], 0u, 0x4240) + B2 - __byte_perm(pr[j][q], 0u, 0x4240);   // (d0, d2)
            f[j] = __byte_perm(o[j][q], 0u, 0x4341) + B2 - __byte_perm(pr[j][q], 0u, 0x4341);   // (d1, d3)
        }
        auto had4 = [&](uint32_t (&x)[4]) {   // 4-point Hadamard down the rows, two columns per word
            const uint32_t u0 = x[0] + x[1] - B2, u1 = x[0] - x[1] + B2, u2 = x[2] + x[3] - B2, u3 = x[2] - x[3] + B2;
            x[0] = u0 + u2 - B2; x[1] = u1 + u3 - B2; x[2] = u0 - u2 + B2; x[3] = u1 - u3 + B2;
        };
        had4(e);
        had4(f);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t sd[2] = {e[r] + f[r] - B2, e[r] - f[r] + B2};   // (a, c) and (b, e')
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t m = __vmaxu2(sd[h], 0x80008000u - sd[h]);   // |.| + bias in both lanes
                const uint32_t mx = __vmaxu2(m, __byte_perm(m, m, 0x1032));   // max of the two lanes, in both
                acc = acc + mx - B2;
            }
        }
    }
    return (int)(2u * (acc & 0xffffu));
}

// Strip geometry of the packed search: a lane owns 4 scan lines x SW samples (SW = 8 for N >= 8,
// which amortises the per-scan-line address arithmetic over twice the pixels; 4 for N = 4).
template <int N, int G>
struct StripCfg {
    static constexpr int SW = N >= 8 ? 8 : 4;
    static constexpr int SB = N * N / (4 * SW);          // strips per block
    static constexpr int SPL = SB >= G ? SB / G : 1;     // strips per lane
    static constexpr int MS = SB >= G ? 1 : G / SB;      // mode splits
    static constexpr int SBL = SB >= G ? G : SB;         // lanes that sum one mode's cost
    static constexpr int SPR = N / SW;                   // strips per row of strips
    static constexpr int WPS = SW / 4;                   // packed words per scan line
};

// Cost of one strip.  (b0, s0) = first base / first scan line of the strip in the orientation of the
// mode: vertical modes use (x, y) of the image, horizontal modes (y, x) -- the strip is then the
// transposed region and `oh` holds the transposed original samples.
template <int N, int G>
__device__ __forceinline__ int strip_cost_u8(int mode, int b0, int s0,
                                             const uint32_t (&ov)[4][StripCfg<N, G>::WPS],
                                             const uint32_t (&oh)[4][StripCfg<N, G>::WPS],
                                             const int16_t* top, const int16_t* left,
                                             const int16_t* neg, int dc, int cost_kind) {
    using Cfg = CoderCfg<N, G>;
    using SC = StripCfg<N, G>;
    constexpr int SW = SC::SW, WPS = SC::WPS;
    uint32_t pr[4][WPS];   // predicted scan lines, 4 bytes per word
    bool transposed = false;
    if (mode >= 2) {
        const int angle = intra_angle(mode);
        const bool vertical = mode >= 18;
        transposed = !vertical;
        const int16_t* pos = vertical ? top : left;
        const int16_t* ng = neg + (mode - 11) * Cfg::NEG_W + N;   // only dereferenced for modes 11..25
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = (s0 + j + 1) * angle;
            const int ip = p >> 5, f = p & 31;
            const int k = b0 + 1 + ip;
            const int16_t* R = k < 0 ? ng : pos;   // every array starts 4-byte aligned
            uint32_t w[SW / 2];
            pred_row_pairs<SW>(R, k, f, w);
#pragma unroll
            for (int q = 0; q < WPS; ++q) pr[j][q] = __byte_perm(w[2 * q], w[2 * q + 1], 0x6420);
        }
    } else if (mode == 1) {
        const uint32_t d4 = (uint32_t)dc * 0x01010101u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < WPS; ++q) pr[j][q] = d4;
    } else {
        // planar (intra.py:109-111), image orientation: b0 = x, s0 = y.  Two samples per multiply-add chain: every term
        // is non-negative for 8-bit references and the weights carry a factor 2^(7 - log2 N), so that the sample
        // ((N-1-x) l[y] + (x+1) tr + (N-1-y) t[x] + (y+1) bl + N) >> (log2 N + 1) is the high byte of its 16-bit lane
        // (largest lane value 65408).  (One planar_px per sample made the DC / planar / angular iteration of a warp --
        // three diverged branches -- the longest phase of a block in the multi-warp wavefront kernels.)
        constexpr uint32_t SCL = 1u << (7 - Log2<N>::v);
        const uint32_t tr = (uint32_t)(uint16_t)top[N + 1], bl = (uint32_t)(uint16_t)left[N + 1];
        uint32_t c1[2 * WPS], kc[2 * WPS], zt[2 * WPS];
#pragma unroll
        for (int i = 0; i < 2 * WPS; ++i) {
            const uint32_t X = (uint32_t)(b0 + 2 * i);
            c1[i] = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 2) - X) << 16)) * SCL;
            kc[i] = tr * (((X + 1) | ((X + 2) << 16)) * SCL);
            zt[i] = (uint32_t)(uint16_t)top[1 + b0 + 2 * i] | ((uint32_t)(uint16_t)top[2 + b0 + 2 * i] << 16);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = s0 + j;
            const uint32_t ly = (uint32_t)(uint16_t)left[1 + y];
            const uint32_t vy = (uint32_t)(N - 1 - y) * SCL;
            const uint32_t by = ((uint32_t)(y + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
#pragma unroll
            for (int q = 0; q < WPS; ++q) {
                const uint32_t t0 = ly * c1[2 * q] + kc[2 * q] + vy * zt[2 * q] + by;
                const uint32_t t1 = ly * c1[2 * q + 1] + kc[2 * q + 1] + vy * zt[2 * q + 1] + by;
                pr[j][q] = __byte_perm(t0, t1, 0x7531);
            }
        }
    }
    int c = 0;
    if (cost_kind == NH_COST_SAD) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < WPS; ++q)
                c = (int)(__vsadu4(pr[j][q], transposed ? oh[j][q] : ov[j][q]) + (uint32_t)c);
    } else {
        c = transposed ? satd_strip_packed<WPS>(pr, oh) : satd_strip_packed<WPS>(pr, ov);
    }
    return c;
}

// Same contract as search_modes(); requires all samples of the block in [0, 255] and
// build_neg_arrays() done (and a __syncwarp() since).
template <int N, int G>
__device__ __forceinline__ int search_modes_u8(int gl, const int16_t* O, const int16_t* top,
                                               const int16_t* left, const int16_t* neg, int dc,
                                               int cost_kind, int it0 = 0, int it_step = 1) {
    using Cfg = CoderCfg<N, G>;
    using SC = StripCfg<N, G>;
    constexpr int SW = SC::SW, WPS = SC::WPS;
    // ov: the lane's strip for vertical / DC / planar modes (SW wide, 4 tall at (px, py));
    // oh: its strip for horizontal modes, stored transposed (4 wide, SW tall at (py, px) mirrored:
    //     scan line j = image column qx + j, base i = image row qy + i)
    uint32_t ov[SC::SPL][4][WPS], oh[SC::SPL][4][WPS];
    int px[SC::SPL], py[SC::SPL];
#pragma unroll
    for (int s = 0; s < SC::SPL; ++s) {
        const int st = (SC::MS == 1) ? gl + s * G : gl % SC::SB;
        px[s] = (st % SC::SPR) * SW;   // base offset of the strip (x for vertical, y for horizontal)
        py[s] = (st / SC::SPR) * 4;    // scan offset of the strip (y for vertical, x for horizontal)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < WPS; ++q) {
                uint32_t wv = 0, wh = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = px[s] + 4 * q + i, sc = py[s] + j;
                    wv |= (uint32_t)(uint16_t)O[sc * Cfg::O_PITCH + b] << (8 * i);   // (y = sc, x = b)
                    wh |= (uint32_t)(uint16_t)O[b * Cfg::O_PITCH + sc] << (8 * i);   // (y = b, x = sc)
                }
                ov[s][j][q] = wv;
                oh[s][j][q] = wh;
            }
    }
    const int ms = (SC::MS == 1) ? 0 : gl / SC::SB;
    int best = 0x7fffffff;
    constexpr int ITERS = (35 + SC::MS - 1) / SC::MS;
    for (int it = it0; it < ITERS; it += it_step) {
        const int pos = it * SC::MS + ms;
        const bool active = pos < 35;
        const int mode = !active ? 1 : (pos == 0 ? 1 : (pos == 1 ? 0 : pos));
        int c = 0;
#pragma unroll
        for (int s = 0; s < SC::SPL; ++s)
            c += strip_cost_u8<N, G>(mode, px[s], py[s], ov[s], oh[s], top, left, neg, dc, cost_kind);
#pragma unroll
        for (int off = SC::SBL / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        const int key = active ? ((c << 6) | pos) : 0x7fffffff;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int off = G / 2; off >= SC::SBL; off >>= 1) {
        int other = __shfl_xor_sync(0xffffffffu, best, off);
        best = other < best ? other : best;
    }
    return best;
}

// search_modes_u8() with the block's pixels given as packed bytes: Ob = N rows of N bytes, ObT = the transposed block.
// The int16 form costs every lane 2 x 16 WPS dependent 16-bit loads plus the packing per block (the multi-warp wavefront
// kernels ran that preamble in every warp: ~2500 of a 16x16 block's 5400 cycles); here a strip is 8 WPS word loads.
template <int N, int G>
__device__ __forceinline__ int search_modes_u8_pk(int gl, const unsigned char* Ob, const unsigned char* ObT, const int16_t* top,
                                               const int16_t* left, const int16_t* neg, int dc,
                                               int cost_kind, int it0 = 0, int it_step = 1) {
    using Cfg = CoderCfg<N, G>;
    using SC = StripCfg<N, G>;
    constexpr int SW = SC::SW, WPS = SC::WPS;
    // ov: the lane's strip for vertical / DC / planar modes (SW wide, 4 tall at (px, py));
    // oh: its strip for horizontal modes, stored transposed (4 wide, SW tall at (py, px) mirrored:
    //     scan line j = image column qx + j, base i = image row qy + i)
    uint32_t ov[SC::SPL][4][WPS], oh[SC::SPL][4][WPS];
    int px[SC::SPL], py[SC::SPL];
#pragma unroll
    for (int s = 0; s < SC::SPL; ++s) {
        const int st = (SC::MS == 1) ? gl + s * G : gl % SC::SB;
        px[s] = (st % SC::SPR) * SW;   // base offset of the strip (x for vertical, y for horizontal)
        py[s] = (st / SC::SPR) * 4;    // scan offset of the strip (y for vertical, x for horizontal)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < WPS; ++q) {
                const int w = (py[s] + j) * N + px[s] + 4 * q;   // bytes (sc, b .. b+3); px is a multiple of 4
                ov[s][j][q] = *reinterpret_cast<const uint32_t*>(Ob + w);    // (y = sc, x = b + i)
                oh[s][j][q] = *reinterpret_cast<const uint32_t*>(ObT + w);   // (y = b + i, x = sc)
            }
    }
    const int ms = (SC::MS == 1) ? 0 : gl / SC::SB;
    int best = 0x7fffffff;
    constexpr int ITERS = (35 + SC::MS - 1) / SC::MS;
    for (int it = it0; it < ITERS; it += it_step) {
        // Candidate positions rotated by one iteration: iteration 1 holds positions 0 .. MS-1 -- DC, planar and the first
        // angular modes, three diverged branches -- and iteration 0 the last ones.  With W warps sharing the iterations
        // (it0 = warp, it_step = W) the warp that runs one iteration more than the others is warp 0, which then gets
        // uniform angular iterations only; the key carries the true position, so the winner does not depend on the order.
        // (MS == 1, N = 32: every iteration is uniform already and the plain order happens to be the better balanced.)
        int pos = it * SC::MS + ms - (SC::MS > 1 ? SC::MS : 0);
        if (pos < 0) pos += ITERS * SC::MS;
        const bool active = pos < 35;
        const int mode = !active ? 1 : (pos == 0 ? 1 : (pos == 1 ? 0 : pos));
        int c = 0;
#pragma unroll
        for (int s = 0; s < SC::SPL; ++s)
            c += strip_cost_u8<N, G>(mode, px[s], py[s], ov[s], oh[s], top, left, neg, dc, cost_kind);
#pragma unroll
        for (int off = SC::SBL / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        const int key = active ? ((c << 6) | pos) : 0x7fffffff;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int off = G / 2; off >= SC::SBL; off >>= 1) {
        int other = __shfl_xor_sync(0xffffffffu, best, off);
        best = other < best ? other : best;
    }
    return best;
}

__device__ __forceinline__ int mode_of_key(int key) {
    const int pos = key & 63;
    return pos == 0 ? 1 : (pos == 1 ? 0 : pos);
}

struct CoderOut {
    uint8_t* modes;     // (B,)
    int32_t* costs;     // (B,)
    int16_t* pred;      // (B,N,N)
    int32_t* coeff;     // (B,N,N)
    int32_t* levels;    // (B,N,N)
    int16_t* recon;     // (B,N,N) block-major, or NULL
    int16_t* recon_plane;  // (H, pitch) or NULL
    int pitch;
};

// Row r of the winning mode for the 8-bit fast path: references come from the contiguous per-mode
// arrays of build_neg_arrays() (k < 0) or the plain reference arrays (k >= 0); no int16 wrap can
// occur for 8-bit samples, so the interpolation is the plain formula of intra.py:206-207.
template <int N, int G>
__device__ __forceinline__ void predict_row_u8(int mode, int r, const int16_t* top, const int16_t* left,
                                               const int16_t* neg, int dc, int (&p)[N]) {
    using Cfg = CoderCfg<N, G>;
    if (mode == 1) {
#pragma unroll
        for (int x = 0; x < N; ++x) p[x] = dc;
        return;
    }
    if (mode == 0) {
        int tv[N];
#pragma unroll
        for (int x = 0; x < N; ++x) tv[x] = top[1 + x];
        planar_row<N>(r, (int)left[1 + r], tv, (int)top[N + 1], (int)left[N + 1], p);
        return;
    }
    const int angle = intra_angle(mode);
    const bool vertical = mode >= 18;
    const int16_t* pos = vertical ? top : left;
    const int16_t* ng = neg + (mode - 11) * Cfg::NEG_W + N;  // only dereferenced when k < 0
    auto ref = [&](int k) -> int { return (int)(k < 0 ? ng : pos)[k]; };
    if (vertical) {
        const int pr = (r + 1) * angle;
        const int ip = pr >> 5, f = pr & 31;
        int v[N + 1];
#pragma unroll
        for (int x = 0; x <= N; ++x) v[x] = (x < N || f != 0) ? ref(x + 1 + ip) : 0;
#pragma unroll
        for (int x = 0; x < N; ++x) p[x] = ((32 - f) * v[x] + f * v[x + 1] + 16) >> 5;
    } else {
#pragma unroll
        for (int x = 0; x < N; ++x) {
            const int pr = (x + 1) * angle;
            const int ip = pr >> 5, f = pr & 31;
            const int k = r + 1 + ip;
            const int a = ref(k);
            const int b2 = f != 0 ? ref(k + 1) : 0;
            p[x] = ((32 - f) * a + f * b2 + 16) >> 5;
        }
    }
}

// Samples x0 .. x0 + 7 of row r: the formulas of predict_row_u8() on a segment, so that two lanes can share a row of a
// 16x16 block (its winner stage would otherwise run the prediction on half a warp).
template <int N, int G>
__device__ __forceinline__ void predict_seg8_u8(int mode, int r, int x0, const int16_t* top, const int16_t* left,
                                                const int16_t* neg, int dc, int (&p)[8]) {
    using Cfg = CoderCfg<N, G>;
    if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = dc;
        return;
    }
    if (mode == 0) {
        const int ly = (int)left[1 + r], tr = (int)top[N + 1], bl = (int)left[N + 1];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = planar_px<N>(x0 + i, r, ly, (int)top[1 + x0 + i], tr, bl);
        return;
    }
    const int angle = intra_angle(mode);
    const bool vertical = mode >= 18;
    const int16_t* pos = vertical ? top : left;
    const int16_t* ng = neg + (mode - 11) * Cfg::NEG_W + N;  // only dereferenced when k < 0
    auto ref = [&](int k) -> int { return (int)(k < 0 ? ng : pos)[k]; };
    if (vertical) {
        const int pr = (r + 1) * angle;
        const int ip = pr >> 5, f = pr & 31;
        int v[9];
#pragma unroll
        for (int i = 0; i <= 8; ++i) v[i] = (i < 8 || f != 0) ? ref(x0 + i + 1 + ip) : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = ((32 - f) * v[i] + f * v[i + 1] + 16) >> 5;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pr = (x0 + i + 1) * angle;
            const int ip = pr >> 5, f = pr & 31;
            const int k = r + 1 + ip;
            const int a = ref(k);
            const int b2 = f != 0 ? ref(k + 1) : 0;
            p[i] = ((32 - f) * a + f * b2 + 16) >> 5;
        }
    }
}

// Winner pipeline for one block: row owner r (< N) predicts row r, then the K6 chain through
// the shared working matrix.  All lanes of the warp must call this (it contains __syncwarp).
// The reconstruction is left in O (pitch O_PITCH) for the caller to copy out.
// fast8 (warp-uniform): every sample of the block is 8-bit and `neg` is built -> packed-domain
// prediction and exact 32-bit quant (FastQuant); otherwise the generic int16 / int64 arithmetic.
template <int N, int G>
__device__ __forceinline__ void code_block(int gl, bool valid, int64_t b, int mode, int16_t* O,
                                           int* M, const int16_t* top, const int16_t* left,
                                           int corner, int dc, const QuantParams& qp,
                                           const FastQuant& fq, bool fast8, const int16_t* neg,
                                           int maxv, bool use_dst, const CoderOut& out) {
    using Cfg = CoderCfg<N, G>;
    constexpr int NN = N * N;
    const int r = gl;
    const bool rowlane = r < N;
    uint32_t pw[N / 2];
    if (rowlane) {
        int p[N], res[N];
        if (fast8) {
            predict_row_u8<N, G>(mode, r, top, left, neg, dc, p);
        } else {
#pragma unroll
            for (int x = 0; x < N; ++x) p[x] = predict_px<N>(mode, x, r, top, left, corner, dc);
        }
        pack_row<N>(p, pw);
        if (valid && out.pred) store_row16<N>(out.pred + b * NN + r * N, pw);
#pragma unroll
        for (int x = 0; x < N; ++x) res[x] = sext16((int)O[r * Cfg::O_PITCH + x] - sext16(p[x]));
        store_row_smem<N>(M, r, res);
    }
    __syncwarp();
    {
        int c[N], lv[N], dq[N];
        if (N == 4 && use_dst) two_pass_transform<N, N == 4, false>(M, r, rowlane, c);
        else two_pass_transform<N, false, false>(M, r, rowlane, c);
        if (rowlane) {
            if (valid && out.coeff) store_row32<N>(out.coeff + b * NN + r * N, c);
            if (fast8) {
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    lv[k] = quantize_fast(c[k], fq);
                    dq[k] = dequantize_fast(lv[k], fq);
                }
            } else {
                quant_dequant_row<N>(c, qp, lv, dq);
            }
            if (valid && out.levels) store_row32<N>(out.levels + b * NN + r * N, lv);
        }
        __syncwarp();  // every lane has read its column of the second forward pass
        if (rowlane) store_row_smem<N>(M, r, dq);
    }
    __syncwarp();
    {
        int res[N];
        if (N == 4 && use_dst) two_pass_transform<N, N == 4, true>(M, r, rowlane, res);
        else two_pass_transform<N, false, true>(M, r, rowlane, res);
        if (rowlane) {
            uint32_t ow[N / 2];
#pragma unroll
            for (int k = 0; k < N / 2; ++k) {
                const int a = recon_px(lo16(pw[k]), res[2 * k], maxv);
                const int c2 = recon_px(hi16(pw[k]), res[2 * k + 1], maxv);
                ow[k] = pack16(a, c2);
                *reinterpret_cast<uint32_t*>(O + r * Cfg::O_PITCH + 2 * k) = ow[k];
            }
            if (valid && out.recon) store_row16<N>(out.recon + b * NN + r * N, ow);
        }
    }
    __syncwarp();
}

// DC value from padded shared-memory references (top[1..N], left[1..N]).
template <int N>
__device__ __forceinline__ int dc_from_refs(const int16_t* top, const int16_t* left) {
    int s = 0;
#pragma unroll
    for (int k = 1; k <= N; ++k) s += (int)top[k] + (int)left[k];
    return dc_value<N>(s);
}

// The same sum by a whole (converged) warp: two loads per lane and one REDUX instead of 2N dependent shared-memory loads
// per thread (the multi-warp wavefront kernels compute the DC value on a block's critical path).
template <int N>
__device__ __forceinline__ int dc_from_refs_warp(int lane, const int16_t* top, const int16_t* left) {
    int s = 0;
#pragma unroll
    for (int k = lane; k < N; k += 32) s += (int)top[1 + k] + (int)left[1 + k];
    return dc_value<N>(__reduce_add_sync(0xffffffffu, s));
}

}  // namespace nh
