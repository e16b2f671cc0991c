// nh_search.cuh -- K7 search stage for 8-bit planes: the exhaustive 35-mode decision of every full
// block of a plane (source neighbours), written as modes (uint8) + costs (int32).  The winner
// pipeline runs afterwards in coder_kernel, which is handed the modes (nh_frame.cu).
//
// Why a kernel of its own: in the one-kernel coder the candidate modes of a block are split across
// the lanes that share it, so every mode-dependent quantity (angle, orientation, fraction, window
// parity) is per-lane data and the hot loop spends more instructions selecting than interpolating
// (ncu, 8x8 / SAD: 467 thread instructions per pixel, ALU pipe 73 %).  Here
//   * a warp tile holds enough blocks that lane = one 4 x SW strip and ALL lanes evaluate the SAME
//     modes (N = 4 / 8 / 16 / 32: 32 / 16 / 4 / 1 blocks per warp): angle and orientation are uniform;
//   * horizontal mode m and vertical mode 36 - m share the angle, hence the integer part, fraction and
//     window start of every scan line: they are evaluated together (left references against the
//     transposed strip, top references against the strip);
//   * references are kept as BYTES.  A first version held pair words (ref[t], ref[t+1]) for every t --
//     no alignment work, but 8 shared-memory wavefronts per scan line, and ncu showed the kernel bound
//     by the 128 B/clk shared-memory crossbar (LSU pipe 57 %, short-scoreboard 2.3 warps per issue,
//     issue 53 %).  A scan line of 8 samples needs 9 consecutive bytes: three words, one funnel shift
//     each, and PRMTs against RZ spread them into the 16-bit lanes (b0,b2) (b1,b3) (b2,b4) ...;
//   * weights are scaled by 8: 8 * (32 * 255 + 16) = 65408 < 2^16, so the predicted sample is the HIGH
//     BYTE of its 16-bit lane and one PRMT packs four of them -- no shift, no mask;
//   * negative angles read the projected extension of _build_ref_array (intra.py:180-186, the (k+1)
//     projection of SURVEY Q3) from per-mode byte arrays holding only the entries the reference itself
//     fills (k >= (N * angle) >> 5) followed by a copy of the first bytes of the primary array, so that
//     a window that starts below zero is one contiguous read.
// Same candidate order and tie rule as search_modes(): positions 1, 0, 2 .. 34, first strict minimum.
// A tile with any sample outside [0, 255] is not decided here: its blocks get mode 0xFF and the coder
// kernel runs its own exact search for them.
#pragma once
#include "nh_coder.cuh"
#include "nh_plane.cuh"

namespace nh {

// INTRA_PRED_ANGLE of mode 11 + mi (the negative angles), as a function so that device code can fold it
// after unrolling: -2, -5, -9, -13, -17, -21, -26, -32, -26, ..., -2
__host__ __device__ constexpr int neg_angle_at(int mi) {
    const int d = mi <= 7 ? mi : 14 - mi;
    return d == 0 ? -2 : d == 1 ? -5 : d == 2 ? -9 : d == 3 ? -13 : d == 4 ? -17 : d == 5 ? -21 : d == 6 ? -26 : -32;
}

// Position of scan line yy of the mirror pair (mode, 36 - mode), mode = 2 .. 18 (intra.py:191-207), ready to use:
//   k4 = byte offset of the word that holds ref[k]  (k = 1 + ((yy+1) * angle >> 5), k4 = k & ~3, may be negative)
//   x  = 8 k     (funnel-shift amount: the hardware takes it modulo 32 = 8 (k & 3))
//   y  = selector of the last sample's byte         (0x3412 + ((k & 3) << 8))
//   z  = 8 f     (f = (yy+1) * angle & 31; weights scaled by 8: the sample is the high byte of its 16-bit lane)
//   w  = 8 (32 - f)
// The compiler keeps per-mode values in vector registers and recomputed all of this on the ALU pipe for every
// line (13 instructions per mirror-pair line); constant loads replace them.  k4 has a table of its own: it only
// enters addresses, and a value that is not a vector operand can stay in a uniform register.
struct LineTab { int4 e[17][32]; int k4[17][32]; };
constexpr LineTab make_line_tab() {
    constexpr int ang[17] = {32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32};
    LineTab t{};
    for (int m = 0; m < 17; ++m)
        for (int yy = 0; yy < 32; ++yy) {
            const int p = (yy + 1) * ang[m];
            const int k = 1 + (p >> 5);
            t.k4[m][yy] = k & ~3;
            t.e[m][yy].x = 8 * k;
            t.e[m][yy].y = 0x3412 + ((k & 3) << 8);
            t.e[m][yy].z = (p & 31) << 3;
            t.e[m][yy].w = 256 - ((p & 31) << 3);
        }
    return t;
}
static __constant__ LineTab kc_line_tab = make_line_tab();

// PRMT with a run-time selector handed over as it is (__byte_perm masks it with 0x7777 first: one more ALU-pipe
// instruction per use)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

template <int N>
struct SearchCfg {
    static constexpr int SW = N >= 8 ? 8 : 4;          // strip width (a lane owns 4 scan lines x SW samples)
    static constexpr int SB = N * N / (4 * SW);        // strips per block
    static constexpr int T = 32 / SB;                  // blocks per warp tile
    static constexpr int SPR = N / SW;                 // strips per row of strips
    static constexpr int WPS = SW / 4;                 // packed words per scan line
    static constexpr int PB = ((2 * N + 9) + 3) / 4 * 4;   // bytes of a positive array: ref[0 .. 2N+1] + word-read slack
    static constexpr int CP = 4 * (WPS + 1);           // bytes of the primary array copied behind a projected extension
    __host__ __device__ static constexpr int neg_len(int mi) { return -((N * neg_angle_at(mi)) >> 5); }   // entries the reference fills
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
    static constexpr int GP = T >= 16 ? 1 : (T == 4 ? 4 : 8);   // build: groups of modes per orientation and block
    static constexpr int MPG = 8 / GP;                 // modes per group
    static constexpr int WARPS = 4;
    static constexpr int TAB_WORDS = N == 4 ? 17 * 4 * 5 : 0;   // N = 4: the CTA's copy of the scan-line table (int4 + int per entry)
    static constexpr int SMEM_BYTES = (WARPS * WARP_WORDS + 16 + TAB_WORDS) * 4;
};

struct SearchArgs {
    const int16_t* src;
    int H, W, pitch;
    int cost_kind;
    int64_t n_blocks;          // blocks of all frames
    uint8_t* modes;
    int32_t* costs;
    int64_t blocks_per_frame;  // frames are stacked `frame_stride` samples apart; block b belongs to frame
    int64_t frame_stride;      // b / blocks_per_frame (a warp tile may straddle two frames)
    // CODE = true (N = 4): the winner is coded by the lane that searched the block
    FastQuant fq;
    int maxv;
    int16_t* pred;             // (B, 4, 4) or NULL
    int32_t* coeff;            // (B, 4, 4) or NULL
    int32_t* levels;           // (B, 4, 4) or NULL
    int16_t* recon_plane;      // frames as src (pitch, frame_stride)
    unsigned int* handed_back; // counts the tiles left to the exact coder kernel (mode 0xFF)
};

// SAD (metrics.py:24-26) or the sum of satd_4x4 (metrics.py:29-43) of a strip held as packed bytes.
template <int WPS>
__device__ __forceinline__ int strip_cost_packed(const uint32_t (&pr)[4][WPS], const uint32_t (&o)[4][WPS],
                                                 int cost_kind) {
    int c = 0;
    if (cost_kind == NH_COST_SAD) {
        // two accumulate chains (VABSDIFF4.U8.ACC with a live accumulator): written as asm because the
        // compiler otherwise starts every instruction from RZ and adds the partial sums on the ALU pipe,
        // which is the pipe that bounds this kernel
        uint32_t c0 = 0, c1 = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < WPS; ++q) {
                asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(c0) : "r"(pr[j][q]), "r"(o[j][q]));
                asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(c1) : "r"(pr[j + 2][q]), "r"(o[j + 2][q]));
            }
        c = (int)(c0 + c1);
    } else {
        c = satd_strip_packed<WPS>(pr, o);
    }
    return c;
}

// One scan line of SW predicted samples (packed bytes):
//   ((32-f) ref[k+i] + f ref[k+i+1] + 16) >> 5  for i = 0 .. SW-1   (intra.py:191-207; f8 = 8 f, g8 = 8 (32 - f)).
// wp = word that holds ref[k], sh = 8 (k & 3), sel_last = 0x3412 + ((k & 3) << 8): the three depend on k
// only, so the two halves of a mirror pair share them.
template <int WPS, bool RAW_SEL = false>
__device__ __forceinline__ void predict_line_w(const uint32_t* wp, uint32_t sh, uint32_t sel_last, uint32_t f8,
                                               uint32_t g8, uint32_t (&out)[WPS]) {
    uint32_t w[WPS + 1], v[WPS];
#pragma unroll
    for (int i = 0; i <= WPS; ++i) w[i] = wp[i];
#pragma unroll
    for (int i = 0; i < WPS; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);   // bytes k+4i .. k+4i+3
#pragma unroll
    for (int q = 0; q < WPS; ++q) {
        const uint32_t e0 = __byte_perm(v[q], 0u, 0x4240);        // (b0, b2)
        const uint32_t o0 = __byte_perm(v[q], 0u, 0x4341);        // (b1, b3)
        const uint32_t e1 = q + 1 < WPS ? __byte_perm(e0, v[q + 1], 0x3412)      // (b2, b4)
                                        : (RAW_SEL ? prmt(e0, w[WPS], sel_last) : __byte_perm(e0, w[WPS], sel_last));   // b4 = byte k & 3 of the last word
        const uint32_t t02 = g8 * e0 + 0x00800080u + f8 * o0;     // samples 0, 2 in the high bytes
        const uint32_t t13 = g8 * o0 + 0x00800080u + f8 * e1;     // samples 1, 3
        out[q] = __byte_perm(t02, t13, 0x7351);
    }
}
template <int WPS>
__device__ __forceinline__ void predict_line_u8(const unsigned char* bytes, int ob, uint32_t f8, uint32_t g8,
                                                uint32_t (&out)[WPS]) {
    const uint32_t k3 = (uint32_t)ob & 3u;
    predict_line_w<WPS>(reinterpret_cast<const uint32_t*>(bytes + (ob & ~3)), k3 * 8u, 0x3412u + (k3 << 8), f8, g8, out);
}

// The same for a scan line whose fraction is 0 (angles 0 and +-32): the samples are the bytes themselves.
template <int WPS>
__device__ __forceinline__ void copy_line_w(const uint32_t* wp, uint32_t sh, uint32_t (&out)[WPS]) {
    uint32_t w[WPS + 1];
#pragma unroll
    for (int i = 0; i <= WPS; ++i) w[i] = wp[i];
#pragma unroll
    for (int i = 0; i < WPS; ++i) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

// Winner stage of the 4x4 search kernels, lane = block: prediction of the decided mode as four packed rows, then
// residual, forward DST-VII, quantise, dequantise, inverse, reconstruct in this lane's registers (8-bit samples: the
// 32-bit quantiser forms are exact).  blk = the block's reference arrays (top | left | per-mode
// negative-angle arrays), negT0 = byte t = 0 of mode 11 + mi from blk, ov = the block's pixels as packed rows.
__device__ __forceinline__ void code_winner4(const SearchArgs& a, int wmode, int dc, const unsigned char* blk, const int* negT0,
                                             const uint32_t (&ov)[4][1], bool valid, int64_t b, int fr, int x, int y) {
    constexpr int N = 4;
    constexpr int PB4 = ((2 * N + 9) + 3) / 4 * 4;
    const unsigned char* tb = blk;
    const unsigned char* lb = blk + PB4;
    uint32_t pw[4];
    if (wmode == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pw[j] = (uint32_t)dc * 0x01010101u;
    } else if (wmode == 0) {   // intra.py:109-111
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                w |= (uint32_t)planar_px<N>(i, j, lb[1 + j], tb[1 + i], tb[N + 1], lb[N + 1]) << (8 * i);
            pw[j] = w;
        }
    } else {   // intra.py:116-207 with this lane's own angle
        const int wangle = intra_angle(wmode);
        const bool vert = wmode >= 18;
        const int pri = vert ? 0 : PB4;
        const int ngo = wangle < 0 ? negT0[wmode - 11] : pri;
        uint32_t ln[4][1];
        int p = wangle;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t f8 = ((uint32_t)p & 31u) << 3;
            const int k = 1 + (p >> 5);
            predict_line_u8<1>(blk, (k < 0 ? ngo : pri) + k, f8, 256u - f8, ln[j]);
            p += wangle;
        }
        if (vert) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pw[j] = ln[j][0];
        } else {   // scan line = image column: 4x4 byte transpose
            const uint32_t u0 = __byte_perm(ln[0][0], ln[1][0], 0x5140), v0 = __byte_perm(ln[2][0], ln[3][0], 0x5140);
            const uint32_t u1 = __byte_perm(ln[0][0], ln[1][0], 0x7362), v1 = __byte_perm(ln[2][0], ln[3][0], 0x7362);
            pw[0] = __byte_perm(u0, v0, 0x5410);
            pw[1] = __byte_perm(u0, v0, 0x7632);
            pw[2] = __byte_perm(u1, v1, 0x5410);
            pw[3] = __byte_perm(u1, v1, 0x7632);
        }
    }
    // ---- residual, forward DST-VII, quantise, dequantise, inverse, reconstruct: all in this lane's registers
    // (8-bit samples: the 32-bit quantiser forms are exact)
    int r[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            r[j][i] = (int)((ov[j][0] >> (8 * i)) & 0xffu) - (int)((pw[j] >> (8 * i)) & 0xffu);
    transform2d<4, true, false>(r);
    int lv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) lv[j][i] = quantize_fast(r[j][i], a.fq);
    if (valid) {
        if (a.pred) {
            uint4* pp = reinterpret_cast<uint4*>(a.pred + b * 16);
            __stcs(pp, make_uint4(__byte_perm(pw[0], 0u, 0x4140), __byte_perm(pw[0], 0u, 0x4342),
                                  __byte_perm(pw[1], 0u, 0x4140), __byte_perm(pw[1], 0u, 0x4342)));
            __stcs(pp + 1, make_uint4(__byte_perm(pw[2], 0u, 0x4140), __byte_perm(pw[2], 0u, 0x4342),
                                      __byte_perm(pw[3], 0u, 0x4140), __byte_perm(pw[3], 0u, 0x4342)));
        }
        if (a.coeff) {
            uint4* cp = reinterpret_cast<uint4*>(a.coeff + b * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) __stcs(cp + j, make_uint4(r[j][0], r[j][1], r[j][2], r[j][3]));
        }
        if (a.levels) {
            uint4* lp = reinterpret_cast<uint4*>(a.levels + b * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) __stcs(lp + j, make_uint4(lv[j][0], lv[j][1], lv[j][2], lv[j][3]));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) r[j][i] = dequantize_fast(lv[j][i], a.fq);
    transform2d<4, true, true>(r);
    if (valid && a.recon_plane) {
        int16_t* rp = a.recon_plane + fr * a.frame_stride + (int64_t)y * a.pitch + x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = recon_px((int)((pw[j] >> (8 * i)) & 0xffu), r[j][i], a.maxv);
            *reinterpret_cast<uint2*>(rp + (int64_t)j * a.pitch) = make_uint2(pack16(q[0], q[1]), pack16(q[2], q[3]));
        }
    }
}

// COST is a template parameter so that the SAD instance is not compiled around the registers of the SATD code
// (6 resident CTAs per SM without spills; the SATD instance runs at 5).
// CODE (N = 4 only; lane = block there): after the search the lane codes its block's winner -- it already holds the
// pixels (packed bytes) and every reference array the prediction can need, so the winner stage costs one more
// candidate evaluation, the DST-VII passes in registers (transform.py:180-236) and the stores, instead of a second
// kernel that gathers the references again (round 2: search 613 us + coder 610 us for 8 4K frames).
template <int N, int COST, bool CODE = false>
__global__ void __launch_bounds__(SearchCfg<N>::WARPS * 32, CODE ? 4 : (COST == NH_COST_SAD ? 6 : 5)) search_plane_kernel(const SearchArgs a) {
    static_assert(!CODE || N == 4, "the fused winner stage exists for 4x4 blocks");
    using C = SearchCfg<N>;
    constexpr int SW = C::SW, SB = C::SB, T = C::T, WPS = C::WPS, S = Log2<N>::v;
    extern __shared__ __align__(16) uint32_t smem_w[];
    int* negT0 = reinterpret_cast<int*>(smem_w + C::WARPS * C::WARP_WORDS);   // byte t = 0 of mode 11+mi, from the block base
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 15) negT0[threadIdx.x] = C::neg_t0((int)threadIdx.x);
    // N = 4 (lane = block: every lane is on the same scan line): the scan-line positions come from the table
    // instead of 13 ALU-pipe instructions per mirror-pair line; see nh_search2.cuh for the measurements
    const int4* s_tab = reinterpret_cast<const int4*>(smem_w + C::WARPS * C::WARP_WORDS + 16);
    const int* s_k4 = reinterpret_cast<const int*>(smem_w + C::WARPS * C::WARP_WORDS + 16 + 17 * 4 * 4);
    if constexpr (N == 4) {
        for (int i = threadIdx.x; i < 17 * 4; i += blockDim.x) {
            const_cast<int4*>(s_tab)[i] = kc_line_tab.e[i / 4][i % 4];
            const_cast<int*>(s_k4)[i] = kc_line_tab.k4[i / 4][i % 4];
        }
    }
    __syncthreads();

    uint32_t* wbase = smem_w + warp * C::WARP_WORDS;
    const int bi = lane / SB, st = lane % SB;        // block of the tile, strip of the block
    const int px_ = (st % C::SPR) * SW;              // base offset of the strip  (x vertical / y horizontal)
    const int py_ = (st / C::SPR) * 4;               // scan offset of the strip  (y vertical / x horizontal)
    unsigned char* blk = reinterpret_cast<unsigned char*>(wbase + bi * C::BLOCK_WORDS);
    const unsigned char* tb = blk;                   // top[0 .. 2N+1]   (index 0 = corner slot)
    const unsigned char* lb = blk + C::PB;           // left[0 .. 2N+1]
    const int bw = a.W / N;
    const int64_t n_tiles = (a.n_blocks + T - 1) / T;

    for (int64_t tile = (int64_t)blockIdx.x * C::WARPS + warp; tile < n_tiles;
         tile += (int64_t)gridDim.x * C::WARPS) {
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
        // every block of the tile away from the frame edges (warp-uniform): the common case
        const bool interior = __all_sync(0xffffffffu, x > 0 && y > 0 && x + 2 * N <= a.W && y + 2 * N <= a.H);
        constexpr int RE = T * (2 * N + 2), RI = (RE + 31) / 32;
        // all loads first, then all stores: written as one loop the compiler keeps load -> store order and the
        // tile pays RI global-memory round trips in a row
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
            unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
            zb[k] = (unsigned char)tv[it];
            zb[C::PB + k] = (unsigned char)lv[it];
            ood |= tv[it] | lv[it];
        }

        // ---- the lane's strip, packed bytes: ov = image orientation, oh = transposed (horizontal modes)
        uint32_t ov[4][WPS], oh[4][WPS];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < WPS; ++q) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(srcf + (int64_t)(y + py_ + j) * a.pitch + x + px_ + 4 * q));
                ood |= (int)((v.x | v.y) & 0xFF00FF00u);
                ov[j][q] = __byte_perm(v.x, v.y, 0x6420);
            }
#pragma unroll
        for (int q = 0; q < WPS; ++q) {
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
        const bool fast8 = !__any_sync(0xffffffffu, (ood & ~0xff) != 0);
        if (!fast8) {   // leave the tile to the coder kernel's exact search
            if (valid && st == 0) a.modes[b] = 0xFF;
            if (CODE && lane == 0) atomicAdd(a.handed_back, 1u);
            continue;
        }
        __syncwarp();

        // ---- projected extensions of the negative-angle modes (intra.py:180-186).  Unit = (block,
        // orientation, group of MPG modes).  Horizontal mode 11 + q and vertical mode 25 - q share the angle,
        // hence the length and every projected index: unrolled over q, each entry is one byte load and one
        // byte store at constant offsets (a loop over table-driven items cost 15 % of the kernel at N = 8).
        for (int u0 = 0; u0 < 2 * T * C::GP; u0 += 32) {
            const int u = u0 + lane;
            if (u < 2 * T * C::GP) {
                const int i = u / (2 * C::GP), r = u % (2 * C::GP), o = r / C::GP, g = r % C::GP;
                unsigned char* zb = reinterpret_cast<unsigned char*>(wbase + i * C::BLOCK_WORDS);
                const unsigned char* sec = zb + (o ? C::PB : 0);     // vertical: secondary = left, primary = top
                uint32_t pw[WPS + 1];
#pragma unroll
                for (int c = 0; c <= WPS; ++c) pw[c] = reinterpret_cast<const uint32_t*>(zb + (o ? 0 : C::PB))[c];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int len = C::neg_len(14 - q);
                    const int inv = inv_angle(neg_angle_at(14 - q));
                    if (q / C::MPG == g && (q < 7 || o)) {   // mode 18 (q = 7) is vertical only
                        unsigned char* dst = zb + (o ? C::neg_t0(14 - q) : C::neg_t0(q < 7 ? q : 6));
#pragma unroll
                        for (int c = 0; c <= WPS; ++c) reinterpret_cast<uint32_t*>(dst)[c] = pw[c];   // ref[t], t >= 0
#pragma unroll
                        for (int tt = 0; tt < len; ++tt) {                                           // t = -1 - tt
                            const int proj = (-tt * inv + 128) >> 8;                                 // (k+1) projection, Q3
                            dst[-1 - tt] = sec[proj > 2 * N ? 2 * N : proj];
                        }
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
        __syncwarp();

        uint32_t pr[4][WPS], prh[4][WPS];
        int best;
        {   // position 0: DC
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int q = 0; q < WPS; ++q) pr[j][q] = (uint32_t)dc * 0x01010101u;
            int c = strip_cost_packed<WPS>(pr, ov, COST);
#pragma unroll
            for (int off = SB / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
            best = c << 6;
        }
        {   // position 1: planar (intra.py:109-111), two samples per multiply-add chain; the weights carry a
            // factor 2^(7-S) so that the sample is the high byte of its 16-bit lane (max 65408)
            constexpr uint32_t SC = 1u << (7 - S);
            const uint32_t tr = (uint32_t)tb[N + 1], bl = (uint32_t)lb[N + 1];
            uint32_t kc[SW / 2], c1[SW / 2], zt[SW / 2];
#pragma unroll
            for (int i = 0; i < SW / 2; ++i) {
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
                uint32_t t[SW / 2];
#pragma unroll
                for (int i = 0; i < SW / 2; ++i) t[i] = ly * c1[i] + kc[i] + vy * zt[i] + by;
#pragma unroll
                for (int q = 0; q < WPS; ++q) pr[j][q] = __byte_perm(t[2 * q], t[2 * q + 1], 0x7531);
            }
            int c = strip_cost_packed<WPS>(pr, ov, COST);
#pragma unroll
            for (int off = SB / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
            const int key = (c << 6) | 1;
            best = key < best ? key : best;
        }
        // ---- positions 2..34: angular modes (intra.py:116-207), the same modes on every lane: horizontal
        // mode m together with its mirror, vertical mode 36 - m.  Mode 18 is its own mirror: its
        // horizontal half is computed and dropped.
#pragma unroll 1
        for (int mode = 2; mode <= 18; ++mode) {
            const int angle = intra_angle(mode);
            const int vmode = 36 - mode;
            const int mi_c = mode < 11 ? 0 : mode - 11;
            const int negh = negT0[mi_c];          // only used when k < 0 (modes 11..25)
            const int negv = negT0[14 - mi_c];
            int p = (py_ + 1) * angle;
            if constexpr (N == 4) {
                // a negative-angle mode reads its own array on every line (k <= 0 there, and the array holds a copy
                // of ref[0 .. 7] behind the projection): one base per mode and orientation
                const unsigned char* bv = blk + (angle < 0 ? negv : 0);
                const unsigned char* bh = blk + (angle < 0 ? negh : C::PB);
                const int4* tab = s_tab + (mode - 2) * 4;
                const int* tk4 = s_k4 + (mode - 2) * 4;
                if ((angle & 31) == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k4 = tk4[j];
                        const uint32_t sh = (uint32_t)tab[j].x;
                        copy_line_w<WPS>(reinterpret_cast<const uint32_t*>(bv + k4), sh, pr[j]);
                        copy_line_w<WPS>(reinterpret_cast<const uint32_t*>(bh + k4), sh, prh[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k4 = tk4[j];
                        const int4 e = tab[j];
                        predict_line_w<WPS, true>(reinterpret_cast<const uint32_t*>(bv + k4), (uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z,
                                                  (uint32_t)e.w, pr[j]);
                        predict_line_w<WPS, true>(reinterpret_cast<const uint32_t*>(bh + k4), (uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z,
                                                  (uint32_t)e.w, prh[j]);
                    }
                }
            } else if ((angle & 31) == 0) {   // modes 2 / 34, 10 / 26, 18: every fraction is 0
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = px_ + 1 + (p >> 5);
                    const int k4 = k & ~3;
                    const uint32_t sh = (uint32_t)(k & 3) * 8u;
                    copy_line_w<WPS>(reinterpret_cast<const uint32_t*>(blk + (k < 0 ? negv : 0) + k4), sh, pr[j]);
                    copy_line_w<WPS>(reinterpret_cast<const uint32_t*>(blk + (k < 0 ? negh : C::PB) + k4), sh, prh[j]);
                    p += angle;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t f8 = ((uint32_t)p & 31u) << 3, g8 = 256u - f8;
                    const int k = px_ + 1 + (p >> 5);
                    // every array starts on a word: the word offset, the byte shift and the selector of the
                    // last byte depend on k only and serve both halves
                    const int k4 = k & ~3;
                    const uint32_t k3 = (uint32_t)k & 3u, sh = k3 * 8u, sel_last = 0x3412u + (k3 << 8);
                    predict_line_w<WPS>(reinterpret_cast<const uint32_t*>(blk + (k < 0 ? negv : 0) + k4), sh, sel_last, f8,
                                        g8, pr[j]);
                    predict_line_w<WPS>(reinterpret_cast<const uint32_t*>(blk + (k < 0 ? negh : C::PB) + k4), sh, sel_last,
                                        f8, g8, prh[j]);
                    p += angle;
                }
            }
            int cv = strip_cost_packed<WPS>(pr, ov, COST);
            int ch = strip_cost_packed<WPS>(prh, oh, COST);
#pragma unroll
            for (int off = SB / 2; off > 0; off >>= 1) {
                cv += __shfl_xor_sync(0xffffffffu, cv, off);
                ch += __shfl_xor_sync(0xffffffffu, ch, off);
            }
            const int keyv = (cv << 6) | vmode;
            const int keyh = mode < 18 ? ((ch << 6) | mode) : 0x7fffffff;
            best = keyv < best ? keyv : best;
            best = keyh < best ? keyh : best;
        }
        if (valid && st == 0) {
            a.modes[b] = (uint8_t)mode_of_key(best);
            if (a.costs) a.costs[b] = best >> 6;
        }
        if constexpr (CODE) code_winner4(a, mode_of_key(best), dc, blk, negT0, ov, valid, b, fr, x, y);
    }
}

}  // namespace nh
