// nh_wave.cuh -- K8, the reconstructed-neighbour (raster-order) coder for 8-bit planes at N = 8 and N = 4
// (included by nh_frame.cu after CoderArgs and the generic coder).  docs/frames_and_panes.md:344,
// block.py:68-74: every block predicts from the RECONSTRUCTION of its left, above-left, above and
// above-right neighbours, so a frame is an anti-diagonal wavefront whose critical path is bw + 2 bh
// dependent block times.  What bounds a frame is therefore the latency of ONE block, and these kernels
// are built around that:
//   * a lane's search unit is FIXED for the whole kernel -- (candidate mode, 4-line strip) -- so every
//     mode-dependent quantity (angle, integer offset and fraction of each scan line, the word / shift of
//     its reference window, the projected-extension indices of intra.py:180-186) is computed once per
//     kernel and lives in registers; per block a unit is a handful of shared-memory words, the packed
//     interpolation of nh_search.cuh and a VABSDIFF4 chain;
//   * N = 8: a CTA of three warps owns a block row -- 70 units (35 modes x 2 strips) evaluated at once,
//     argmin by REDUX + one barrier; warp 0 then codes the winner on the tensor cores (the register-
//     chained 8x8 MMA pipeline of nh_fused_mma.cuh with the second block of the block-diagonal operand
//     left empty) while warps 1 / 2 stage the next block's pixels;
//   * N = 4: one warp owns a block row (19 units: 16 mirror pairs, mode 18, DC, planar; no CTA barrier at
//     all) and codes the 4x4 winner across 16 lanes, the DST passes as shuffles;
//   * the exchange-row protocol is the generic coder's (nh_frame.cu): a block publishes its reconstructed
//     bottom row in global memory, -1 = not written yet, the row below polls with ld.cg; the next block's
//     reference samples are requested before the current winner is coded so that the L2 round trip is off
//     the critical path whenever the row above is far enough ahead;
//   * a block with a source sample outside [0, 255] is coded by warp 0 with the generic int16 / int64
//     routines (search_modes / code_block), so the result is exact for every int16 input; planes declared
//     deeper than 8 bits use the generic kernel.
#pragma once
#include "nh_search.cuh"

namespace nh {

// Per-phase cycle counters of the first block row (development only: make NVCCFLAGS+=-DNH_WAVE_PROF; the
// counters land in the scratch header behind the ticket, 8 x int64 at byte 64).
#ifdef NH_WAVE_PROF
#define NH_PROF_DECL long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define NH_PROF_MARK(i) { const long long now__ = clock64(); prof_acc[i] += now__ - prof_t; prof_t = now__; }
#define NH_PROF_DUMP(cond) if (cond) { for (int i__ = 0; i__ < 8; ++i__) atomicAdd(reinterpret_cast<unsigned long long*>(a.ticket) + 8 + i__, (unsigned long long)prof_acc[i__]); }
#else
#define NH_PROF_DECL
#define NH_PROF_MARK(i)
#define NH_PROF_DUMP(cond)
#endif

static __constant__ signed char kc_wave_dct8[64] = {
    64, 64, 64, 64, 64, 64, 64, 64, 89, 75, 50, 18, -18, -50, -75, -89, 83, 36, -36, -83, -83, -36, 36, 83,
    75, -18, -89, -50, 50, 89, 18, -75, 64, -64, -64, 64, 64, -64, -64, 64, 50, -89, 18, 75, -75, -18, 89, -50,
    36, -83, 83, -36, -36, 83, -83, 36, 18, -50, 75, -89, 89, -75, 50, -18};

__device__ __forceinline__ uint32_t ldsm_x1(uint32_t addr) {
    uint32_t r;
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ldsm_x1_t(uint32_t addr) {
    uint32_t r;
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void stsm_x1(uint32_t addr, uint32_t r) {
    asm volatile("stmatrix.sync.aligned.m8n8.x1.shared.b16 [%0], {%1};" ::"r"(addr), "r"(r) : "memory");
}
__device__ __forceinline__ void hmma1688_w(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, float c0, float c1,
                                           float c2, float c3) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(b0), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}

// Cold path of the wavefront kernels: one block coded by ONE warp with the generic, exact routines of
// nh_coder.cuh (int16 interpolation, int64 quantisation) from byte references `refb` (tb at 0, lb at PB;
// reconstructed samples, hence 8-bit) and the source pixels in global memory.  The reconstruction is left
// in r16 (N rows of N int16).  Kept out of line so that its registers do not count against the hot path.
template <int N>
__device__ __noinline__ void wave_block_generic(const CoderArgs& a, const unsigned char* refb, int pb,
                                                const int16_t* srcf, int x, int y, int64_t b, unsigned char* gen,
                                                int16_t* r16) {
    using GC = CoderCfg<N, 32>;
    const int lane = threadIdx.x & 31;
    int16_t* g_top = reinterpret_cast<int16_t*>(gen);
    int16_t* g_left = g_top + GC::REF_W;
    int16_t* g_neg = reinterpret_cast<int16_t*>(gen + GC::REFS_PAD);
    int16_t* g_O = reinterpret_cast<int16_t*>(gen + GC::REFS_PAD + GC::NEG_BYTES);
    int* g_M = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(g_O) + GC::O_BYTES);
    for (int k = lane; k < GC::REF_W; k += 32) {
        const int kk = k <= 2 * N ? k : 2 * N;
        g_top[k] = (int16_t)refb[kk];
        g_left[k] = (int16_t)refb[pb + kk];
    }
    for (int e = lane; e < N * N; e += 32)
        g_O[(e / N) * GC::O_PITCH + (e % N)] = srcf[(int64_t)(y + e / N) * a.pitch + x + e % N];
    __syncwarp();
    const int corner = (int)g_top[0];
    const int dc = dc_from_refs<N>(g_top, g_left);
    const int key = search_modes<N, 32>(lane, g_O, g_top, g_left, corner, dc, a.cost_kind);
    const int wmode = mode_of_key(key);
    if (lane == 0) {
        if (a.out.modes) a.out.modes[b] = (uint8_t)wmode;
        if (a.out.costs) a.out.costs[b] = key >> 6;
    }
    code_block<N, 32>(lane, true, b, wmode, g_O, g_M, g_top, g_left, corner, dc, a.qp, a.fq, false, g_neg, a.maxv,
                      a.use_dst != 0, a.out);
    for (int e = lane; e < N * N; e += 32) r16[e] = g_O[(e / N) * GC::O_PITCH + (e % N)];
    __syncwarp();
}

// ------------------------------------------------------------------------------------------ N = 8
// Roles inside the CTA of a block row (four warps).  Units: warps 0 / 1 hold modes 2 .. 33, warp 2 mode 34, warp 3 DC
// and planar (a warp that mixed unit kinds would run them one after the other and hold up the barrier).  Beyond that
//   warp 0  codes the winner: picks the winner's prediction tile (every unit has written its predicted strip
//           into a tile of its mode, so nothing is predicted twice and no mode-dependent code sits behind the
//           argmin), runs the MMA chain, publishes the bottom row, stores prediction / coefficients / levels /
//           reconstruction straight from its fragments and derives the next block's left references from them;
//   warp 1  stages the next block's pixels (bytes, transposed bytes, int16 ldmatrix tile);
//   warp 2  is the service warp: it polls the exchange row for the next block's top references while warp 0
//           codes the winner, and writes the block's mode / cost.
// Two barriers per block: references + pixels ready, partial minima ready.
struct Wave8Smem {
    static constexpr int N = 8;
    using SC = SearchCfg<8>;
    using GC = CoderCfg<8, 32>;
    static constexpr int REF_BYTES = SC::BLOCK_WORDS * 4;   // tb | lb | projected extensions (nh_search.cuh layout)
    static constexpr int OB = 0;                            // 2 x { 64 B pixels, 64 B transposed } as bytes
    static constexpr int O16 = OB + 2 * 128;                // 2 x 8 rows of 8 int16 (ldmatrix tiles)
    static constexpr int TILES = O16 + 2 * 128;             // 35 prediction tiles (by candidate position), 8 x 8 int16;
                                                            // horizontal modes hold the TRANSPOSED prediction
    static constexpr int R16 = TILES + 35 * 128;            // reconstruction of a generic-path block
    static constexpr int REFS = R16 + 128;
    static constexpr int GEN = (REFS + REF_BYTES + 15) / 16 * 16;   // generic coder's group (out-of-domain blocks)
    static constexpr int TOTAL = GEN + GC::GROUP_BYTES;
};

template <int COST, int OCC>
__global__ void __launch_bounds__(128, OCC) wave8_kernel(const CoderArgs a) {
    constexpr int N = 8, NN = 64, SH = 8;
    using SC = SearchCfg<8>;
    using L = Wave8Smem;
    constexpr int PB = SC::PB;
    __shared__ __align__(16) unsigned char smem[L::TOTAL];
    __shared__ int s_keys[4];
    __shared__ int s_row;
    __shared__ int s_topsum, s_leftsum;
    __shared__ int s_ood[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* refb = smem + L::REFS;          // tb = refb, lb = refb + PB

    // ---- this lane's search unit, fixed for the whole kernel
    const bool unit = tid < 66 || (tid >= 96 && tid < 100);
    const int strip = tid & 1;
    const bool angular = tid < 66;
    const int mode = angular ? 2 + (tid >> 1) : (tid < 98 ? 1 : 0);
    const int pos = angular ? mode : (tid < 98 ? 0 : 1);            // candidate order 1, 0, 2 .. 34
    const bool vertical = mode >= 18;
    const int angle = angular ? intra_angle(mode) : 0;
    const bool negmode = angular && angle < 0;
    const int negoff = negmode ? SC::neg_t0(mode - 11) : 0;
    int woff[4];
    uint32_t sh[4], f8[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = (4 * strip + j + 1) * angle;
        const int k = 1 + (p >> 5);
        woff[j] = (k < 0 ? negoff : (vertical ? 0 : PB)) + (k & ~3);
        sh[j] = (uint32_t)(k & 3) * 8u;
        f8[j] = ((uint32_t)p & 31u) << 3;
    }
    const int obase = L::OB + ((angular && !vertical) ? 64 : 0) + 32 * strip;   // the unit's 4 scan lines of pixels
    unsigned char* my_tile = smem + L::TILES + (unit ? pos : 0) * 128 + 64 * strip;   // its 4 rows of the mode's tile
    // projected extension (intra.py:180-186, the (k+1) projection of SURVEY Q3): the two strip lanes of a
    // negative-angle mode build the mode's array together, entries tt = strip, strip + 2, ...
    int nsrc[4] = {0, 0, 0, 0};
    int nent = 0;
    const int ndst = negoff - 1 - strip;
    if (negmode) {
        const int len = -((N * angle) >> 5), inv = inv_angle_of_mode(mode);
        const int sec = vertical ? PB : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int tt = strip + 2 * i;
            int proj = (-tt * inv + 128) >> 8;
            proj = proj > 2 * N ? 2 * N : proj;
            nsrc[i] = sec + proj;
            nent += tt < len;
        }
    }
    const int npri = vertical ? 0 : PB;

    // ---- winner pipeline constants (warp 0): the 8-point matrix as two registers per lane
    const int fg = lane >> 2, ft = lane & 3;
    const uint32_t tf = pack_h2((float)kc_wave_dct8[fg * 8 + 2 * ft], (float)kc_wave_dct8[fg * 8 + 2 * ft + 1]);
    const uint32_t ttf = pack_h2((float)kc_wave_dct8[(2 * ft) * 8 + fg], (float)kc_wave_dct8[(2 * ft + 1) * 8 + fg]);
    const uint4 a_fwd = make_uint4(tf, 0u, 0u, tf), a_inv = make_uint4(ttf, 0u, 0u, ttf);
    const float rnd = (float)(1 << (SH - 1));
    const float init_f2 = fg == 0 ? rnd - (float)(kOperandBias * 512) : rnd;
    const uint32_t clip_lo2 = 0x10001000u;
    const uint32_t clip_hi2 = clip_lo2 + (uint32_t)a.maxv * 0x10001u;
    const FastQuant fq = a.fq;
    const uint32_t row_addr = (uint32_t)(lane & 7) * 16u;   // ldmatrix x1: lanes 0..7 address the 8 rows
    const uint32_t tiles_addr = smem_u32(smem + L::TILES) + row_addr;
    const uint32_t o16_addr = smem_u32(smem + L::O16) + row_addr;

    // kernel parameters, read once
    const int bw = a.W / N, bh = a.H / N, W = a.W, pitch = a.pitch, n_frames = a.n_frames;
    const bool vec_exch = (W % 2) == 0;                 // 4-byte aligned exchange-row pairs
    uint8_t* const o_modes = a.out.modes;
    int32_t* const o_costs = a.out.costs;
    int16_t* const o_pred = a.out.pred;
    int32_t* const o_coeff = a.out.coeff;
    int32_t* const o_levels = a.out.levels;
    const unsigned poll_sleep = a.poll_sleep_ns;

    for (;;) {
        __syncthreads();   // everyone is done with the previous row (and has read s_row)
        if (tid == 0) s_row = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int tk = s_row;   // frames interleaved: ticket t = row t / F of frame t % F
        const int by = tk / n_frames, fr = tk - by * n_frames;
        if (by >= bh) break;
        const int y = by * N;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int16_t* bottomf = a.bottom + (int64_t)fr * bh * W;
        const int16_t* up = bottomf + (int64_t)(by - 1) * W;   // exchange row above (by > 0)
        const int64_t blk0 = fr * a.blocks_per_frame + (int64_t)by * bw;
        // per-row pointers of this lane
        const int16_t* px_ptr = srcf + (int64_t)(y + (lane >> 1)) * pitch + 4 * (lane & 1);           // warp 1 staging
        int16_t* recon_ptr = a.out.recon_plane + fr * a.frame_stride + (int64_t)(y + fg) * pitch + 2 * ft;   // warp 0
        int16_t* bottom_ptr = bottomf + (int64_t)by * W + 2 * ft;
        int16_t* pred_ptr = o_pred + blk0 * NN + fg * 8 + 2 * ft;
        int32_t* coeff_ptr = o_coeff + blk0 * NN + (2 * ft) * N + fg;
        int32_t* levels_ptr = o_levels + blk0 * NN + (2 * ft) * N + fg;

        // ---- row prologue
        uint2 npx = make_uint2(0u, 0u);   // warp 1, lanes 0..15: 4 pixels of the block being staged
        auto fetch_px = [&](int bx) {     // row l / 2, columns 4 (l & 1) .. (launcher: pitch % 4 == 0)
            if (lane < 16) npx = __ldg(reinterpret_cast<const uint2*>(px_ptr + bx * N));
        };
        auto stage_px = [&](int par) {   // registers -> int16 tile, byte matrix, transposed byte matrix
            int bad = 0;
            if (lane < 16) {
                const int r = lane >> 1, c4 = 4 * (lane & 1);
                *reinterpret_cast<uint2*>(smem + L::O16 + par * 128 + r * 16 + c4 * 2) = npx;
                const uint32_t b4 = __byte_perm(npx.x, npx.y, 0x6420);
                *reinterpret_cast<uint32_t*>(smem + L::OB + par * 128 + r * 8 + c4) = b4;
                unsigned char* t = smem + L::OB + par * 128 + 64 + r;   // transposed: [column][row]
                t[(c4 + 0) * 8] = (unsigned char)b4;
                t[(c4 + 1) * 8] = (unsigned char)(b4 >> 8);
                t[(c4 + 2) * 8] = (unsigned char)(b4 >> 16);
                t[(c4 + 3) * 8] = (unsigned char)(b4 >> 24);
                bad = (int)((npx.x | npx.y) & 0xFF00FF00u);
            }
            bad = __any_sync(0xffffffffu, bad != 0);
            if (lane == 0) s_ood[par] = bad;
        };
        // top references of block bx (service warp; lane k < 18 = entry k of tb): poll the exchange row above
        // until the data is there, then write the bytes, the corner slot of lb and the sum of top[1..N]
        auto stage_top = [&](int bx) {
            int v = 128;
            if (by > 0) {
                const int x = bx * N;
                int last = x + 2 * N - 1;
                if (last > W - 1) last = W - 1;
                int col = x + (lane > 2 * N ? 2 * N : lane) - 1;
                if (col > last) col = last;
                const bool fixed = lane >= 18 || (lane == 0 && x == 0);   // corner of the first column: 128
                unsigned spins = 0;
                for (;;) {
                    v = fixed ? 128 : (int)__ldcg(up + col);
                    if (__all_sync(0xffffffffu, v >= 0)) break;
                    if (poll_sleep) __nanosleep(poll_sleep);
                    if (++spins > (1u << 25)) __trap();   // > 10 s of polling: a protocol error, fail loudly instead of hanging
                }
            }
            if (lane < 18) refb[lane] = (unsigned char)v;
            if (lane == 0) refb[PB] = (unsigned char)v;
            const int ts = __reduce_add_sync(0xffffffffu, (lane >= 1 && lane <= N) ? v : 0);
            if (lane == 0) s_topsum = ts;
        };
        if (warp == 1) {
            fetch_px(0);
            stage_px(0);
            if (bw > 1) fetch_px(1);
        } else if (warp == 2) {
            stage_top(0);
        } else if (warp == 0) {
            if (lane >= 1 && lane < 18) refb[PB + lane] = 128;   // left references of the first block (block.py:45-50)
            if (lane == 0) s_leftsum = N * 128;
        }
        NH_PROF_DECL
        for (int bx = 0; bx < bw; ++bx) {
            const int par = bx & 1, x = bx * N;
            NH_PROF_MARK(7)
            __syncthreads();   // #1: references, sums and this block's pixels are in shared memory
            NH_PROF_MARK(2)
            const bool ood = s_ood[par] != 0;   // CTA-uniform
            if (!ood) {
                // ---- search: one unit per lane
                if (negmode) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const unsigned char v = refb[nsrc[i]];
                        if (i < nent) refb[ndst - 2 * i] = v;
                    }
                    if (strip == 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            reinterpret_cast<uint32_t*>(refb + negoff)[c] = reinterpret_cast<const uint32_t*>(refb + npri)[c];
                    }
                }
                __syncwarp();   // the two strip lanes of a mode sit next to each other in one warp
                int c = 0;
                if (unit) {
                    uint32_t pr[4][2], o[4][2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint2 v = *reinterpret_cast<const uint2*>(smem + obase + par * 128 + 8 * j);
                        o[j][0] = v.x;
                        o[j][1] = v.y;
                    }
                    if (angular) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            predict_line_w<2>(reinterpret_cast<const uint32_t*>(refb + woff[j]), sh[j], 0x3412u + (sh[j] << 5),
                                              f8[j], 256u - f8[j], pr[j]);
                    } else if (mode == 1) {
                        const uint32_t d4 = (uint32_t)dc_value<N>(s_topsum + s_leftsum) * 0x01010101u;   // intra.py:46-62
#pragma unroll
                        for (int j = 0; j < 4; ++j) pr[j][0] = pr[j][1] = d4;
                    } else {   // planar (intra.py:109-111), two samples per multiply-add chain, sample = high byte
                        constexpr uint32_t SCL = 1u << (7 - 3);
                        const unsigned char* tb = refb;
                        const unsigned char* lb = refb + PB;
                        const uint32_t tr = tb[N + 1], bl = lb[N + 1];
                        uint32_t kc[4], c1[4], zt[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t X = (uint32_t)(2 * i);
                            c1[i] = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 2) - X) << 16)) * SCL;
                            kc[i] = tr * (((X + 1) | ((X + 2) << 16)) * SCL);
                            zt[i] = (uint32_t)tb[1 + 2 * i] | ((uint32_t)tb[2 + 2 * i] << 16);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int yy = 4 * strip + j;
                            const uint32_t ly = lb[1 + yy];
                            const uint32_t vy = (uint32_t)(N - 1 - yy) * SCL;
                            const uint32_t byv = ((uint32_t)(yy + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
                            uint32_t t[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) t[i] = ly * c1[i] + kc[i] + vy * zt[i] + byv;
                            pr[j][0] = __byte_perm(t[0], t[1], 0x7531);
                            pr[j][1] = __byte_perm(t[2], t[3], 0x7531);
                        }
                    }
                    // the unit's predicted strip into the tile of its mode, as int16 (ldmatrix tile rows)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(my_tile + 16 * j) =
                            make_uint4(__byte_perm(pr[j][0], 0u, 0x4140), __byte_perm(pr[j][0], 0u, 0x4342),
                                       __byte_perm(pr[j][1], 0u, 0x4140), __byte_perm(pr[j][1], 0u, 0x4342));
                    c = strip_cost_packed<2>(pr, o, COST);
                }
                c += __shfl_xor_sync(0xffffffffu, c, 1);
                const int key = unit ? ((c << 6) | pos) : 0x7fffffff;
                const int wmin = (int)__reduce_min_sync(0xffffffffu, (unsigned)key);
                if (lane == 0) s_keys[warp] = wmin;
            }
            NH_PROF_MARK(3)
            __syncthreads();   // #2: the three partial minima, the prediction tiles
            NH_PROF_MARK(4)
            if (warp == 1) {          // next block's pixels, while warp 0 codes the winner
                if (bx + 1 < bw) stage_px(par ^ 1);
                if (bx + 2 < bw) fetch_px(bx + 2);
            } else if (warp == 2) {   // mode / cost of this block, top references of the next one
                if (!ood && lane == 0) {
                    int best = s_keys[0];
                    best = s_keys[1] < best ? s_keys[1] : best;
                    best = s_keys[2] < best ? s_keys[2] : best;
                    best = s_keys[3] < best ? s_keys[3] : best;
                    if (o_modes) o_modes[blk0 + bx] = (uint8_t)mode_of_key(best);
                    if (o_costs) o_costs[blk0 + bx] = best >> 6;
                }
                // (an out-of-domain block is coded by warp 0 from the CURRENT references: wait for it below)
                if (!ood && bx + 1 < bw) stage_top(bx + 1);
            } else if (warp == 0) {
                uint32_t rr;   // reconstructed pair (row fg, columns 2 ft, 2 ft + 1)
                if (!ood) {
                    int best = s_keys[0];
                    best = s_keys[1] < best ? s_keys[1] : best;
                    best = s_keys[2] < best ? s_keys[2] : best;
                    best = s_keys[3] < best ? s_keys[3] : best;
                    const int wpos = best & 63;
                    const bool transposed = wpos >= 2 && wpos < 18;   // horizontal modes: the tile holds P^T
                    const uint32_t tile = tiles_addr + (uint32_t)wpos * 128u;
                    const uint32_t t_t = ldsm_x1_t(tile), t_n = ldsm_x1(tile);
                    const uint32_t rp = transposed ? t_n : t_t;   // P as B fragment (k = row, n = column)
                    const uint32_t pc = transposed ? t_t : t_n;   // P in accumulator layout (row fg, columns 2 ft ..)
                    const uint32_t ro = ldsm_x1_t(o16_addr + (uint32_t)par * 128u);
                    if (o_pred) *reinterpret_cast<uint32_t*>(pred_ptr + bx * NN) = pc;
                    NH_PROF_MARK(5)
                    // ---- the four passes on the tensor cores (see nh_fused_mma.cuh; block b of the pair is empty)
                    float acc[4];
                    const uint32_t x0 = h2_bits(__hsub2(bits_h2(ro | 0x64006400u), bits_h2(rp | 0x64006400u)));
                    hmma16816(acc, a_fwd, x0, 0u, rnd, rnd, rnd, rnd);
                    uint32_t h0 = round_pair_biased<SH>(acc[0], acc[1]), h1 = round_pair_biased<SH>(acc[2], acc[3]);
                    hmma16816(acc, a_fwd, h0, h1, init_f2, init_f2, init_f2, init_f2);
                    float dqf[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cf = __float_as_int(floor_shift_magic<SH>(acc[e])) - kMagicI;
                        const int lvq = quantize_fast(cf, fq);
                        const int dq = dequantize_fast(lvq, fq);
                        if (o_coeff) __stcs(coeff_ptr + bx * NN + e * N, cf);
                        if (o_levels) __stcs(levels_ptr + bx * NN + e * N, lvq);
                        dqf[e] = __int_as_float(dq + kMagicI) - kMagicF;
                    }
                    h0 = pack_h2(dqf[0], dqf[1]);
                    hmma16816(acc, a_inv, h0, 0u, rnd, rnd, rnd, rnd);
                    h0 = round_pair_plain<SH>(acc[0], acc[1]);
                    h1 = round_pair_plain<SH>(acc[2], acc[3]);
                    hmma1688_w(acc, h0, h1, ttf, rnd, rnd, rnd, rnd);
                    const uint32_t m0 = __float_as_uint(__fmaf_rd(acc[0], 1.0f / (float)(1 << SH), kMagicF + 4096.0f));
                    const uint32_t m1 = __float_as_uint(__fmaf_rd(acc[1], 1.0f / (float)(1 << SH), kMagicF + 4096.0f));
                    const uint32_t sum = __byte_perm(m0, m1, 0x5410) + pc;
                    rr = __vminu2(__vmaxu2(sum, clip_lo2), clip_hi2) - clip_lo2;
                    NH_PROF_MARK(6)
                } else {
                    // ---- exact generic path: int16 references, int64 quantisation; reconstruction through R16
                    int16_t* r16 = reinterpret_cast<int16_t*>(smem + L::R16);
                    wave_block_generic<N>(a, refb, PB, srcf, x, y, blk0 + bx, smem + L::GEN, r16);
                    rr = *reinterpret_cast<const uint32_t*>(r16 + fg * 8 + 2 * ft);
                }
                // ---- publish the bottom row first (the row below is polling for it), then the plane
                if (fg == 7) {
                    if (vec_exch) {
                        __stcg(reinterpret_cast<uint32_t*>(bottom_ptr + x), rr);
                    } else {
                        __stcg(bottom_ptr + x, (int16_t)(rr & 0xffffu));
                        __stcg(bottom_ptr + x + 1, (int16_t)(rr >> 16));
                    }
                }
                *reinterpret_cast<uint32_t*>(recon_ptr + x) = rr;
                // ---- left references of the next block: this block's right-most column, replicated below
                // (bottom-left is not reconstructed yet: n_left = N), and their sum for the DC predictor
                const int rc = (int)(rr >> 16);                       // column 2 ft + 1
                const int r77 = __shfl_sync(0xffffffffu, rc, 31);     // sample (7, 7)
                if (ft == 3) {
                    refb[PB + 1 + fg] = (unsigned char)rc;
                    refb[PB + 9 + fg] = (unsigned char)r77;
                    if (fg == 7) refb[PB + 17] = (unsigned char)r77;
                }
                const int ls = __reduce_add_sync(0xffffffffu, ft == 3 ? rc : 0);
                if (lane == 0) s_leftsum = ls;
            }
            if (ood) {   // CTA-uniform, rare: the generic coder has finished reading the references
                __syncthreads();
                if (warp == 2 && bx + 1 < bw) stage_top(bx + 1);
            }
        }
        NH_PROF_DUMP(tid == 0 && by == 0 && fr == 0)
    }
}

// ------------------------------------------------------------------------------------------ N = 4
static __constant__ signed char kc_wave_dst4[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};

// One warp per block row.  Search units: lane l < 16 = horizontal mode 2 + l together with its mirror, vertical
// mode 34 - l (same angle: same window and fraction per scan line); lane 16 = mode 18; lane 17 = DC; lane 18 =
// planar.  Winner: lanes 0..15 own pixel (y, x) = (lane / 4, lane % 4); the separable DST-VII passes of
// transform.py:180-194 / :222-236 run as four shuffles + four multiply-adds each.
template <int COST>
__global__ void __launch_bounds__(32, 24) wave4_kernel(const CoderArgs a) {
    constexpr int N = 4, NN = 16, SHT = 7;   // transform shift log2(N) + 5
    using SC = SearchCfg<4>;
    using GC = CoderCfg<4, 32>;
    constexpr int PB = SC::PB;
    constexpr int REF_BYTES = (SC::BLOCK_WORDS * 4 + 15) / 16 * 16;
    constexpr int GEN = REF_BYTES + 16 + 32;
    __shared__ __align__(16) unsigned char smem[GEN + GC::GROUP_BYTES];
    __shared__ int s_negT0[15];
    const int lane = threadIdx.x;
    unsigned char* refb = smem;                                       // tb = refb, lb = refb + PB, projected arrays behind
    uint32_t* obw = reinterpret_cast<uint32_t*>(smem + REF_BYTES);   // the block's 4 rows as packed bytes
    int16_t* r16 = reinterpret_cast<int16_t*>(smem + REF_BYTES + 16);   // reconstruction of a generic-path block
    if (lane < 15) s_negT0[lane] = SC::neg_t0(lane);
    __syncwarp();

    // ---- this lane's search unit
    const bool pairlane = lane < 17;
    const int hm = lane < 16 ? 2 + lane : 18;        // horizontal mode (18: its own mirror, coded as vertical only)
    const int vm = 36 - hm;
    const int angle = pairlane ? intra_angle(hm) : 0;
    const bool negmode = pairlane && angle < 0;
    const int negoff_h = (negmode && hm < 18) ? SC::neg_t0(hm - 11) : 0;
    const int negoff_v = negmode ? SC::neg_t0(vm - 11) : 0;
    int woffv[4], woffh[4];
    uint32_t sh[4], f8[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = (j + 1) * angle;
        const int k = 1 + (p >> 5);
        woffv[j] = (k < 0 ? negoff_v : 0) + (k & ~3);
        woffh[j] = (k < 0 ? negoff_h : PB) + (k & ~3);
        sh[j] = (uint32_t)(k & 3) * 8u;
        f8[j] = ((uint32_t)p & 31u) << 3;
    }
    uint32_t nsrc = 0;   // projected indices of entries tt = 0 .. 3 (intra.py:180-186, (k+1) projection)
    int nlen = 0;
    if (negmode) {
        nlen = -((N * angle) >> 5);
        const int inv = inv_angle_of_mode(hm);
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
            int proj = (-tt * inv + 128) >> 8;
            proj = proj > 2 * N ? 2 * N : proj;
            nsrc |= (uint32_t)proj << (8 * tt);
        }
    }
    // ---- winner constants: pixel (py, px) of lanes 0..15, rows / columns of the DST-VII matrix
    const int py = (lane >> 2) & 3, px = lane & 3;
    int t_row_y[4], t_row_x[4], t_col_y[4], t_col_x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t_row_y[k] = kc_wave_dst4[py * 4 + k];   // T[py][k]
        t_row_x[k] = kc_wave_dst4[px * 4 + k];   // T[px][k]
        t_col_y[k] = kc_wave_dst4[k * 4 + py];   // T[k][py]
        t_col_x[k] = kc_wave_dst4[k * 4 + px];   // T[k][px]
    }
    const FastQuant fq = a.fq;
    const int bw = a.W / N, bh = a.H / N;
    const bool vec_exch = (a.W % 4) == 0;   // 8-byte aligned exchange rows

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(a.ticket, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        const int by = tk / a.n_frames, fr = tk - by * a.n_frames;   // frames interleaved
        if (by >= bh) break;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int16_t* reconf = a.out.recon_plane + fr * a.frame_stride;
        int16_t* bottomf = a.bottom + (int64_t)fr * bh * a.W;
        const int16_t* up = bottomf + (int64_t)(by - 1) * a.W;
        const int y = by * N;
        auto top_load = [&](int bx) -> int {   // lane k < 10 holds entry k of tb
            if (by == 0) return 128;
            const int x = bx * N;
            if (lane == 0 && x == 0) return 128;
            int last = x + 2 * N - 1;
            if (last > a.W - 1) last = a.W - 1;
            int col = x + (lane > 2 * N ? 2 * N : lane) - 1;
            if (col > last) col = last;
            return lane < 10 ? (int)__ldcg(up + col) : 0;
        };
        auto px_load = [&](int bx) -> uint2 {   // lane r < 4: row r of the block (launcher: pitch % 4 == 0)
            return lane < 4 ? __ldg(reinterpret_cast<const uint2*>(srcf + (int64_t)(y + lane) * a.pitch + bx * N))
                            : make_uint2(0u, 0u);
        };
        int ntop = top_load(0);
        uint2 npx = px_load(0);
        if (lane >= 1 && lane < 10) refb[PB + lane] = 128;   // left references of the first block
        NH_PROF_DECL
        for (int bx = 0; bx < bw; ++bx) {
            const int x = bx * N;
            NH_PROF_MARK(7)
            const int64_t b = fr * a.blocks_per_frame + (int64_t)by * bw + bx;
            // ---- references
            unsigned spins = 0;
            while (!__all_sync(0xffffffffu, ntop >= 0)) {
                if (a.poll_sleep_ns) __nanosleep(a.poll_sleep_ns);
                if (++spins > (1u << 25)) __trap();   // > 10 s of polling: a protocol error, fail loudly instead of hanging
                ntop = top_load(bx);
            }
            NH_PROF_MARK(0)
            if (lane < 10) refb[lane] = (unsigned char)ntop;
            if (lane == 0) refb[PB] = (unsigned char)ntop;   // corner slot of the left array
            if (lane < 4) obw[lane] = __byte_perm(npx.x, npx.y, 0x6420);
            const bool ood = __any_sync(0xffffffffu, ((npx.x | npx.y) & 0xFF00FF00u) != 0);
            __syncwarp();
            int rec = 0;   // this lane's reconstructed pixel (lanes 0..15)
            if (!ood) {
                const int sref = (lane >= 1 && lane <= N) ? (int)refb[lane] + (int)refb[PB + lane] : 0;
                const int dc = dc_value<N>(__reduce_add_sync(0xffffffffu, sref));   // intra.py:46-62
                const uint4 o4 = *reinterpret_cast<const uint4*>(obw);
                if (bx + 1 < bw) npx = px_load(bx + 1);
                NH_PROF_MARK(1)
                // ---- projected extensions of this lane's two modes
                if (negmode) {
#pragma unroll
                    for (int tt = 0; tt < 4; ++tt)
                        if (tt < nlen) {
                            const int pj = (int)((nsrc >> (8 * tt)) & 0xffu);
                            refb[negoff_v - 1 - tt] = refb[PB + pj];                 // vertical: secondary = left
                            if (hm < 18) refb[negoff_h - 1 - tt] = refb[pj];         // horizontal: secondary = top
                        }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        reinterpret_cast<uint32_t*>(refb + negoff_v)[c] = reinterpret_cast<const uint32_t*>(refb)[c];
                        if (hm < 18) reinterpret_cast<uint32_t*>(refb + negoff_h)[c] = reinterpret_cast<const uint32_t*>(refb + PB)[c];
                    }
                }
                __syncwarp();
                // ---- search
                const uint32_t ov[4][1] = {{o4.x}, {o4.y}, {o4.z}, {o4.w}};
                int key = 0x7fffffff;
                if (pairlane) {
                    uint32_t oh[4][1];
                    {   // 4x4 byte transpose: oh[j] byte i = row i, column j
                        const uint32_t u0 = __byte_perm(o4.x, o4.y, 0x5140), v0 = __byte_perm(o4.z, o4.w, 0x5140);
                        const uint32_t u1 = __byte_perm(o4.x, o4.y, 0x7362), v1 = __byte_perm(o4.z, o4.w, 0x7362);
                        oh[0][0] = __byte_perm(u0, v0, 0x5410);
                        oh[1][0] = __byte_perm(u0, v0, 0x7632);
                        oh[2][0] = __byte_perm(u1, v1, 0x5410);
                        oh[3][0] = __byte_perm(u1, v1, 0x7632);
                    }
                    uint32_t pv[4][1], ph[4][1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t sl = 0x3412u + (sh[j] << 5), g8 = 256u - f8[j];
                        predict_line_w<1>(reinterpret_cast<const uint32_t*>(refb + woffv[j]), sh[j], sl, f8[j], g8, pv[j]);
                        predict_line_w<1>(reinterpret_cast<const uint32_t*>(refb + woffh[j]), sh[j], sl, f8[j], g8, ph[j]);
                    }
                    const int cv = strip_cost_packed<1>(pv, ov, COST), ch = strip_cost_packed<1>(ph, oh, COST);
                    const int keyv = (cv << 6) | vm;
                    const int keyh = hm < 18 ? ((ch << 6) | hm) : 0x7fffffff;
                    key = keyv < keyh ? keyv : keyh;
                } else if (lane == 17) {
                    uint32_t pr[4][1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) pr[j][0] = (uint32_t)dc * 0x01010101u;
                    key = strip_cost_packed<1>(pr, ov, COST) << 6;   // position 0
                } else if (lane == 18) {   // planar (intra.py:109-111)
                    uint32_t pr[4][1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            w |= (uint32_t)planar_px<N>(i, j, refb[PB + 1 + j], refb[1 + i], refb[N + 1], refb[PB + N + 1]) << (8 * i);
                        pr[j][0] = w;
                    }
                    key = (strip_cost_packed<1>(pr, ov, COST) << 6) | 1;
                }
                const int best = (int)__reduce_min_sync(0xffffffffu, (unsigned)key);
                const int wmode = mode_of_key(best);
                NH_PROF_MARK(3)
                if (bx + 1 < bw) ntop = top_load(bx + 1);   // requested now, needed after the winner is coded
                if (lane == 0) {
                    if (a.out.modes) a.out.modes[b] = (uint8_t)wmode;
                    if (a.out.costs) a.out.costs[b] = best >> 6;
                }
                // ---- winner: prediction of pixel (py, px)
                int pred;
                if (wmode == 1) {
                    pred = dc;
                } else if (wmode == 0) {
                    pred = planar_px<N>(px, py, refb[PB + 1 + py], refb[1 + px], refb[N + 1], refb[PB + N + 1]);
                } else {   // intra.py:116-207; 8-bit samples: no int16 wrap in the weighted sum
                    const int wang = intra_angle(wmode);
                    const bool wvert = wmode >= 18;
                    const int scan = wvert ? py : px, base = wvert ? px : py;
                    const int pp = (scan + 1) * wang;
                    const int idx = base + 1 + (pp >> 5), f = pp & 31;
                    const int pri = wvert ? 0 : PB;
                    const int ng = wang < 0 ? s_negT0[wmode - 11] : 0;
                    const int r0 = refb[(idx < 0 ? ng : pri) + idx];
                    const int r1 = refb[(idx + 1 < 0 ? ng : pri) + idx + 1];
                    pred = f ? ((32 - f) * r0 + f * r1 + 16) >> 5 : r0;
                }
                const int orig = (int)reinterpret_cast<const unsigned char*>(obw)[lane & 15];   // pixel (py, px)
                const int res = orig - pred;
                NH_PROF_MARK(5)
                // ---- forward DST-VII (transform.py:180-194): temp = (T X + 64) >> 7, coeff = (temp T^T + 64) >> 7
                int acc = 1 << (SHT - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += t_row_y[k] * __shfl_sync(0xffffffffu, res, 4 * k + px);
                const int temp = acc >> SHT;
                acc = 1 << (SHT - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += t_row_x[k] * __shfl_sync(0xffffffffu, temp, 4 * py + k);
                const int coef = acc >> SHT;   // coeff[py][px]
                const int lvq = quantize_fast(coef, fq);
                const int dq = dequantize_fast(lvq, fq);
                // ---- inverse (transform.py:222-236): temp2 = (T^T C + 64) >> 7, res = (temp2 T + 64) >> 7
                acc = 1 << (SHT - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += t_col_y[k] * __shfl_sync(0xffffffffu, dq, 4 * k + px);
                const int temp2 = acc >> SHT;
                acc = 1 << (SHT - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += t_col_x[k] * __shfl_sync(0xffffffffu, temp2, 4 * py + k);
                const int rres = acc >> SHT;
                rec = clip_pixel(pred + rres, a.maxv);   // intra.py:70-78
                NH_PROF_MARK(6)
                if (lane < 16) {
                    if (a.out.coeff) __stcs(a.out.coeff + b * NN + lane, coef);
                    if (a.out.levels) __stcs(a.out.levels + b * NN + lane, lvq);
                }
                if (a.out.pred) {
                    uint32_t w = (uint32_t)pred | ((uint32_t)__shfl_down_sync(0xffffffffu, pred, 1) << 16);
                    const uint32_t w2 = __shfl_down_sync(0xffffffffu, w, 2);
                    if (lane < 16 && px == 0) *reinterpret_cast<uint2*>(a.out.pred + b * NN + 4 * py) = make_uint2(w, w2);
                }
            } else {
                if (bx + 1 < bw) npx = px_load(bx + 1);
                wave_block_generic<N>(a, refb, PB, srcf, x, y, b, smem + GEN, r16);
                if (bx + 1 < bw) ntop = top_load(bx + 1);
                rec = lane < 16 ? (int)r16[lane] : 0;
            }
            // ---- publish the bottom row first (the row below is polling for it), then the plane
            {
                const uint32_t w = (uint32_t)rec | ((uint32_t)__shfl_down_sync(0xffffffffu, rec, 1) << 16);
                const uint32_t w2 = __shfl_down_sync(0xffffffffu, w, 2);
                if (lane == 12) {
                    if (vec_exch) {
                        __stcg(reinterpret_cast<uint2*>(bottomf + (int64_t)by * a.W + x), make_uint2(w, w2));
                    } else {
                        int16_t* e = bottomf + (int64_t)by * a.W + x;
                        __stcg(e, (int16_t)(w & 0xffff)); __stcg(e + 1, (int16_t)(w >> 16));
                        __stcg(e + 2, (int16_t)(w2 & 0xffff)); __stcg(e + 3, (int16_t)(w2 >> 16));
                    }
                }
                if (lane < 16 && px == 0) *reinterpret_cast<uint2*>(reconf + (int64_t)(y + py) * a.pitch + x) = make_uint2(w, w2);
            }
            __syncwarp();   // every lane is done with this block's references
            // left references of the next block: this block's right-most column; below it is not reconstructed
            // yet: replicate (n_left = N)
            if (lane < 16 && px == 3) refb[PB + 1 + py] = (unsigned char)rec;
            if (lane == 15) {
#pragma unroll
                for (int k = N + 1; k < 10; ++k) refb[PB + k] = (unsigned char)rec;
            }
        }
        NH_PROF_DUMP(lane == 0 && by == 0 && fr == 0)
    }
}

// ------------------------------------------------------------------------------------------ N = 4, four warps
// wave4_kernel above does everything in one warp: about 700 dependent-ish instructions per block, 4230 cycles
// (profiles/r2d_wave_phase_cycles.jsonl: search 1683, DST chain 760, publish 824).  This kernel splits the block
// the way wave8_kernel does:
//   * search: one candidate per lane -- threads 0..32 = angular modes 2..34 (4 scan lines of 4 samples, positions
//     fixed for the whole kernel), thread 96 = DC, threads 100..103 = one planar row each; every unit leaves its
//     predicted 4x4 in a 16-byte tile of its candidate position, argmin by REDUX + one barrier;
//   * warp 0 codes the winner across 16 lanes from the winner's tile (nothing is predicted twice), the DST-VII
//     passes as shuffles, publishes the bottom row, derives the next block's left references;
//   * warp 1 stages the next block's pixels, warp 2 polls the exchange row for its top references and writes
//     mode / cost -- both while warp 0 codes.
struct Wave4Smem {
    using SC = SearchCfg<4>;
    using GC = CoderCfg<4, 32>;
    static constexpr int REF_BYTES = (SC::BLOCK_WORDS * 4 + 15) / 16 * 16;   // tb | lb | projected extensions
    static constexpr int OB = 0;                       // 2 x { 16 B pixel rows, 16 B transposed }
    static constexpr int TILES = OB + 2 * 32;          // 35 predictions of 16 bytes (horizontal modes: transposed)
    static constexpr int R16 = TILES + 35 * 16;        // reconstruction of a generic-path block (16 int16)
    static constexpr int REFS = R16 + 32;
    static constexpr int GEN = REFS + REF_BYTES;       // generic coder's group (out-of-domain blocks)
    static constexpr int TOTAL = GEN + GC::GROUP_BYTES;
};

template <int COST, int OCC>
__global__ void __launch_bounds__(128, OCC) wave4mw_kernel(const CoderArgs a) {
    constexpr int N = 4, NN = 16, SHT = 7;   // transform shift log2(N) + 5
    using SC = SearchCfg<4>;
    using L = Wave4Smem;
    constexpr int PB = SC::PB;
    __shared__ __align__(16) unsigned char smem[L::TOTAL];
    __shared__ int s_keys[4];
    __shared__ int s_row;
    __shared__ int s_topsum, s_leftsum;
    __shared__ int s_ood[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* refb = smem + L::REFS;          // tb = refb, lb = refb + PB

    // ---- this lane's search unit, fixed for the whole kernel
    const bool angular = tid < 33;
    const bool is_dc = tid == 96, is_planar = tid >= 100 && tid < 104;
    const bool unit = angular || is_dc || is_planar;
    const int mode = angular ? 2 + tid : (is_dc ? 1 : 0);
    const int pos = angular ? mode : (is_dc ? 0 : 1);               // candidate order 1, 0, 2 .. 34
    const bool vertical = mode >= 18;
    const int angle = angular ? intra_angle(mode) : 0;
    const bool negmode = angular && angle < 0;
    const int negoff = negmode ? SC::neg_t0(mode - 11) : 0;
    int woff[4];
    uint32_t sh[4], f8[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = (j + 1) * angle;
        const int k = 1 + (p >> 5);
        woff[j] = (k < 0 ? negoff : (vertical ? 0 : PB)) + (k & ~3);
        sh[j] = (uint32_t)(k & 3) * 8u;
        f8[j] = ((uint32_t)p & 31u) << 3;
    }
    const int obase = L::OB + ((angular && !vertical) ? 16 : 0);   // the unit's 4 scan lines of pixels
    uint32_t nsrc = 0;   // projected indices of entries tt = 0 .. 3 (intra.py:180-186, (k+1) projection)
    int nlen = 0;
    if (negmode) {
        nlen = -((N * angle) >> 5);
        const int inv = inv_angle_of_mode(mode);
        const int sec = vertical ? PB : 0;   // vertical: secondary = left
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
            int proj = (-tt * inv + 128) >> 8;
            proj = proj > 2 * N ? 2 * N : proj;
            nsrc |= (uint32_t)(sec + proj) << (8 * tt);
        }
    }
    const int npri = vertical ? 0 : PB;
    unsigned char* my_tile = smem + L::TILES + (unit ? pos : 0) * 16;

    // ---- winner constants (warp 0): pixel (py, px) of lanes 0..15, rows / columns of the DST-VII matrix
    const int py = (lane >> 2) & 3, px = lane & 3;
    int t_row_y[4], t_row_x[4], t_col_y[4], t_col_x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t_row_y[k] = kc_wave_dst4[py * 4 + k];   // T[py][k]
        t_row_x[k] = kc_wave_dst4[px * 4 + k];   // T[px][k]
        t_col_y[k] = kc_wave_dst4[k * 4 + py];   // T[k][py]
        t_col_x[k] = kc_wave_dst4[k * 4 + px];   // T[k][px]
    }
    const FastQuant fq = a.fq;
    const int bw = a.W / N, bh = a.H / N, W = a.W, pitch = a.pitch, n_frames = a.n_frames, maxv = a.maxv;
    const bool vec_exch = (W % 4) == 0;   // 8-byte aligned exchange rows
    uint8_t* const o_modes = a.out.modes;
    int32_t* const o_costs = a.out.costs;
    int16_t* const o_pred = a.out.pred;
    int32_t* const o_coeff = a.out.coeff;
    int32_t* const o_levels = a.out.levels;
    const unsigned poll_sleep = a.poll_sleep_ns;

    for (;;) {
        __syncthreads();   // everyone is done with the previous row (and has read s_row)
        if (tid == 0) s_row = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int tk = s_row;   // frames interleaved: ticket t = row t / F of frame t % F
        const int by = tk / n_frames, fr = tk - by * n_frames;
        if (by >= bh) break;
        const int y = by * N;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int16_t* bottomf = a.bottom + (int64_t)fr * bh * W;
        const int16_t* up = bottomf + (int64_t)(by - 1) * W;   // exchange row above (by > 0)
        const int64_t blk0 = fr * a.blocks_per_frame + (int64_t)by * bw;
        const int16_t* px_ptr = srcf + (int64_t)(y + (lane & 3)) * pitch;                                   // warp 1 staging
        int16_t* recon_ptr = a.out.recon_plane + fr * a.frame_stride + (int64_t)(y + py) * pitch;          // warp 0
        int16_t* bottom_ptr = bottomf + (int64_t)by * W;

        uint2 npx = make_uint2(0u, 0u);   // warp 1, lanes 0..3: row `lane` of the block being staged
        auto fetch_px = [&](int bx) {     // (launcher: pitch % 4 == 0, 8-byte aligned plane)
            if (lane < 4) npx = __ldg(reinterpret_cast<const uint2*>(px_ptr + bx * N));
        };
        auto stage_px = [&](int par) {   // registers -> byte rows, transposed bytes
            int bad = 0;
            if (lane < 4) {
                const uint32_t b4 = __byte_perm(npx.x, npx.y, 0x6420);
                *reinterpret_cast<uint32_t*>(smem + L::OB + par * 32 + 4 * lane) = b4;
                unsigned char* t = smem + L::OB + par * 32 + 16 + lane;   // transposed: [column][row]
                t[0] = (unsigned char)b4;
                t[4] = (unsigned char)(b4 >> 8);
                t[8] = (unsigned char)(b4 >> 16);
                t[12] = (unsigned char)(b4 >> 24);
                bad = (int)((npx.x | npx.y) & 0xFF00FF00u);
            }
            bad = __any_sync(0xffffffffu, bad != 0);
            if (lane == 0) s_ood[par] = bad;
        };
        // top references of block bx (service warp; lane k < 10 = entry k of tb): poll the exchange row above
        auto stage_top = [&](int bx) {
            int v = 128;
            if (by > 0) {
                const int x = bx * N;
                int last = x + 2 * N - 1;
                if (last > W - 1) last = W - 1;
                int col = x + (lane > 2 * N ? 2 * N : lane) - 1;
                if (col > last) col = last;
                const bool fixed = lane >= 10 || (lane == 0 && x == 0);   // corner of the first column: 128
                unsigned spins = 0;
                for (;;) {
                    v = fixed ? 128 : (int)__ldcg(up + col);
                    if (__all_sync(0xffffffffu, v >= 0)) break;
                    if (poll_sleep) __nanosleep(poll_sleep);
                    if (++spins > (1u << 25)) __trap();   // > 10 s of polling: a protocol error, fail loudly instead of hanging
                }
            }
            if (lane < 10) refb[lane] = (unsigned char)v;
            if (lane == 0) refb[PB] = (unsigned char)v;
            const int ts = __reduce_add_sync(0xffffffffu, (lane >= 1 && lane <= N) ? v : 0);
            if (lane == 0) s_topsum = ts;
        };
        if (warp == 1) {
            fetch_px(0);
            stage_px(0);
            if (bw > 1) fetch_px(1);
        } else if (warp == 2) {
            stage_top(0);
        } else if (warp == 0) {
            if (lane >= 1 && lane < 10) refb[PB + lane] = 128;   // left references of the first block (block.py:45-50)
            if (lane == 0) s_leftsum = N * 128;
        }
        NH_PROF_DECL
        for (int bx = 0; bx < bw; ++bx) {
            const int par = bx & 1, x = bx * N;
            NH_PROF_MARK(7)
            __syncthreads();   // #1: references, sums and this block's pixels are in shared memory
            NH_PROF_MARK(2)
            const bool ood = s_ood[par] != 0;   // CTA-uniform
            if (!ood) {
                // ---- search: one candidate per lane
                int c = 0;
                if (unit) {
                    const uint4 o4 = *reinterpret_cast<const uint4*>(smem + obase + par * 32);
                    const uint32_t o[4][1] = {{o4.x}, {o4.y}, {o4.z}, {o4.w}};
                    uint32_t pr[4][1];
                    if (angular) {
                        if (negmode) {   // this mode's projected extension + a copy of ref[0 .. 7] behind it
#pragma unroll
                            for (int tt = 0; tt < 4; ++tt)
                                if (tt < nlen) refb[negoff - 1 - tt] = refb[(nsrc >> (8 * tt)) & 0xffu];
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc)
                                reinterpret_cast<uint32_t*>(refb + negoff)[cc] = reinterpret_cast<const uint32_t*>(refb + npri)[cc];
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            predict_line_w<1>(reinterpret_cast<const uint32_t*>(refb + woff[j]), sh[j], 0x3412u + (sh[j] << 5), f8[j],
                                              256u - f8[j], pr[j]);
                        c = strip_cost_packed<1>(pr, o, COST);
                        *reinterpret_cast<uint4*>(my_tile) = make_uint4(pr[0][0], pr[1][0], pr[2][0], pr[3][0]);
                    } else if (is_dc) {
                        const uint32_t d4 = (uint32_t)dc_value<N>(s_topsum + s_leftsum) * 0x01010101u;   // intra.py:46-62
#pragma unroll
                        for (int j = 0; j < 4; ++j) pr[j][0] = d4;
                        c = strip_cost_packed<1>(pr, o, COST);
                        *reinterpret_cast<uint4*>(my_tile) = make_uint4(d4, d4, d4, d4);
                    } else {   // planar (intra.py:109-111): row yy = lane & 3, two samples per multiply-add chain
                        constexpr uint32_t SCL = 1u << (7 - 2);
                        const unsigned char* tb = refb;
                        const unsigned char* lb = refb + PB;
                        const uint32_t tr = tb[N + 1], bl = lb[N + 1];
                        const int yy = lane & 3;
                        const uint32_t ly = lb[1 + yy];
                        const uint32_t vy = (uint32_t)(N - 1 - yy) * SCL;
                        const uint32_t byv = ((uint32_t)(yy + 1) * bl + (uint32_t)N) * SCL * 0x10001u;
                        uint32_t t[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint32_t X = (uint32_t)(2 * i);
                            const uint32_t c1 = (((uint32_t)(N - 1) - X) | (((uint32_t)(N - 2) - X) << 16)) * SCL;
                            const uint32_t kc = tr * (((X + 1) | ((X + 2) << 16)) * SCL);
                            const uint32_t zt = (uint32_t)tb[1 + 2 * i] | ((uint32_t)tb[2 + 2 * i] << 16);
                            t[i] = ly * c1 + kc + vy * zt + byv;
                        }
                        const uint32_t prow = __byte_perm(t[0], t[1], 0x7531);
                        reinterpret_cast<uint32_t*>(my_tile)[yy] = prow;
                        // cost: SAD adds up over the rows; SATD needs the whole 4x4 in one lane
                        // (only the four planar lanes, 4..7 of warp 3, are in this branch)
                        const uint32_t r0 = __shfl_sync(0xF0u, prow, 4), r1 = __shfl_sync(0xF0u, prow, 5);
                        const uint32_t r2 = __shfl_sync(0xF0u, prow, 6), r3 = __shfl_sync(0xF0u, prow, 7);
                        pr[0][0] = r0; pr[1][0] = r1; pr[2][0] = r2; pr[3][0] = r3;
                        c = strip_cost_packed<1>(pr, o, COST);
                    }
                }
                const int key = (unit && !(is_planar && (lane & 3))) ? ((c << 6) | pos) : 0x7fffffff;
                const int wmin = (int)__reduce_min_sync(0xffffffffu, (unsigned)key);
                if (lane == 0) s_keys[warp] = wmin;
            }
            NH_PROF_MARK(3)
            __syncthreads();   // #2: the partial minima, the prediction tiles
            NH_PROF_MARK(4)
            if (warp == 1) {          // next block's pixels, while warp 0 codes the winner
                if (bx + 1 < bw) stage_px(par ^ 1);
                if (bx + 2 < bw) fetch_px(bx + 2);
            } else if (warp == 2) {   // mode / cost of this block, top references of the next one
                if (!ood && lane == 0) {
                    int best = s_keys[0];
                    best = s_keys[1] < best ? s_keys[1] : best;
                    best = s_keys[3] < best ? s_keys[3] : best;
                    if (o_modes) o_modes[blk0 + bx] = (uint8_t)mode_of_key(best);
                    if (o_costs) o_costs[blk0 + bx] = best >> 6;
                }
                if (!ood && bx + 1 < bw) stage_top(bx + 1);
            } else if (warp == 0) {
                const int64_t b = blk0 + bx;
                int rec;   // this lane's reconstructed pixel (lanes 0..15)
                if (!ood) {
                    int best = s_keys[0];
                    best = s_keys[1] < best ? s_keys[1] : best;
                    best = s_keys[3] < best ? s_keys[3] : best;
                    const int wpos = best & 63;
                    const bool transposed = wpos >= 2 && wpos < 18;   // horizontal modes: the tile holds P^T
                    const int pred = (int)smem[L::TILES + wpos * 16 + (transposed ? px * 4 + py : py * 4 + px)];
                    const int orig = (int)smem[L::OB + par * 32 + (lane & 15)];   // pixel (py, px)
                    const int res = orig - pred;
                    NH_PROF_MARK(5)
                    // ---- forward DST-VII (transform.py:180-194): temp = (T X + 64) >> 7, coeff = (temp T^T + 64) >> 7
                    int acc = 1 << (SHT - 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc += t_row_y[k] * __shfl_sync(0xffffffffu, res, 4 * k + px);
                    const int temp = acc >> SHT;
                    acc = 1 << (SHT - 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc += t_row_x[k] * __shfl_sync(0xffffffffu, temp, 4 * py + k);
                    const int coef = acc >> SHT;   // coeff[py][px]
                    const int lvq = quantize_fast(coef, fq);
                    const int dq = dequantize_fast(lvq, fq);
                    // ---- inverse (transform.py:222-236): temp2 = (T^T C + 64) >> 7, res = (temp2 T + 64) >> 7
                    acc = 1 << (SHT - 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc += t_col_y[k] * __shfl_sync(0xffffffffu, dq, 4 * k + px);
                    const int temp2 = acc >> SHT;
                    acc = 1 << (SHT - 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc += t_col_x[k] * __shfl_sync(0xffffffffu, temp2, 4 * py + k);
                    const int rres = acc >> SHT;
                    rec = clip_pixel(pred + rres, maxv);   // intra.py:70-78
                    NH_PROF_MARK(6)
                    if (lane < 16) {
                        if (o_coeff) __stcs(o_coeff + b * NN + lane, coef);
                        if (o_levels) __stcs(o_levels + b * NN + lane, lvq);
                    }
                    if (o_pred) {
                        const uint32_t w = (uint32_t)pred | ((uint32_t)__shfl_down_sync(0xffffffffu, pred, 1) << 16);
                        const uint32_t w2 = __shfl_down_sync(0xffffffffu, w, 2);
                        if (lane < 16 && px == 0) *reinterpret_cast<uint2*>(o_pred + b * NN + 4 * py) = make_uint2(w, w2);
                    }
                } else {
                    // ---- exact generic path: int16 references, int64 quantisation; reconstruction through R16
                    int16_t* r16 = reinterpret_cast<int16_t*>(smem + L::R16);
                    wave_block_generic<N>(a, refb, PB, srcf, x, y, b, smem + L::GEN, r16);
                    rec = (int)r16[lane & 15];
                }
                // ---- publish the bottom row first (the row below is polling for it), then the plane
                {
                    const uint32_t w = (uint32_t)rec | ((uint32_t)__shfl_down_sync(0xffffffffu, rec, 1) << 16);
                    const uint32_t w2 = __shfl_down_sync(0xffffffffu, w, 2);
                    if (lane == 12) {
                        if (vec_exch) {
                            __stcg(reinterpret_cast<uint2*>(bottom_ptr + x), make_uint2(w, w2));
                        } else {
                            int16_t* e = bottom_ptr + x;
                            __stcg(e, (int16_t)(w & 0xffff)); __stcg(e + 1, (int16_t)(w >> 16));
                            __stcg(e + 2, (int16_t)(w2 & 0xffff)); __stcg(e + 3, (int16_t)(w2 >> 16));
                        }
                    }
                    if (lane < 16 && px == 0) *reinterpret_cast<uint2*>(recon_ptr + x) = make_uint2(w, w2);
                }
                // ---- left references of the next block: this block's right-most column; below it is not
                // reconstructed yet: replicate (n_left = N); and their sum for the DC predictor
                if (lane < 16 && px == 3) refb[PB + 1 + py] = (unsigned char)rec;
                if (lane == 15) {
#pragma unroll
                    for (int k = N + 1; k < 10; ++k) refb[PB + k] = (unsigned char)rec;
                }
                const int ls = __reduce_add_sync(0xffffffffu, (lane < 16 && px == 3) ? rec : 0);
                if (lane == 0) s_leftsum = ls;
            }
            if (ood) {   // CTA-uniform, rare: the generic coder has finished reading the references
                __syncthreads();
                if (warp == 2 && bx + 1 < bw) stage_top(bx + 1);
            }
        }
        NH_PROF_DUMP(tid == 0 && by == 0 && fr == 0)
    }
}

}  // namespace nh
