// nh_frame.cu -- frame-level entry points:
//   K1  nh_gather_refs / nh_plane_to_blocks / nh_blocks_to_plane   (block.py:38-74)
//   K6' nh_fused_pipeline_modes   (any of the 35 modes from given padded references)
//   K7  nh_encode_frame, recon_neighbours = 0   (35-mode search + winner pipeline)
//   K8  nh_encode_frame, recon_neighbours = 1   (anti-diagonal wavefront on the recon plane)
#include <cstdlib>

#include "nh_coder.cuh"
#include "nh_mma.cuh"
#include "nh_plane.cuh"
#include "nh_search.cuh"

namespace nh {

// ------------------------------------------------------------------------ K1
__global__ void __launch_bounds__(256)
    gather_refs_kernel(const int16_t* __restrict__ plane, int H, int W, int pitch, int N, int n_top,
                       int n_left, int16_t* __restrict__ top, int16_t* __restrict__ left,
                       int16_t* __restrict__ corner) {
    const int bw = W / N, bh = H / N, RW = 2 * N + 1;
    const int64_t total = (int64_t)bw * bh * RW;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / RW;
        const int k = (int)(t % RW);
        const int x = (int)(b % bw) * N, y = (int)(b / bw) * N;
        const int tv = top_ref<false>(plane, H, W, pitch, x, y, n_top, k);
        const int lv = left_ref<false>(plane, H, W, pitch, x, y, n_left, k);
        top[t] = (int16_t)tv;
        left[t] = (int16_t)lv;
        if (k == 0) corner[b] = (int16_t)tv;
    }
}

template <bool TO_BLOCKS>
__global__ void __launch_bounds__(256)
    reblock_kernel(const int16_t* __restrict__ in, int H, int W, int pitch, int N,
                   int16_t* __restrict__ out) {
    const int bw = W / N, bh = H / N;
    const int64_t total = (int64_t)bw * bh * N * N;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        // t enumerates plane samples of the covered region row by row (coalesced on the plane side)
        const int cw = bw * N;
        const int py = (int)(t / cw), px = (int)(t % cw);
        const int64_t b = (int64_t)(py / N) * bw + px / N;
        const int64_t bi = b * N * N + (py % N) * N + (px % N);
        const int64_t pi = (int64_t)py * pitch + px;
        if (TO_BLOCKS) out[bi] = in[pi];
        else out[pi] = in[bi];
    }
}

// --------------------------------------------------------------- coder kernel
enum { SRC_ARRAYS = 0, SRC_PLANE = 1, SRC_WAVEFRONT = 2 };

// Winner pipeline of one N x N block (N = 16 / 32) on the tensor cores, for a whole warp whose
// block sits in O with the ldmatrix pitch (CoderCfg<N, 32>::O_PITCH): the row lanes predict the
// winning mode into `ptile` (same layout), then the register-chained passes of nh_mma.cuh code the
// block and leave the reconstruction in O.  8-bit samples only (the callers' fast8 flag).
struct MmaWinnerCtx {
    uint32_t ctab_lane;  // shared-memory address of this lane's constant vectors (MmaConsts<N>)
    int lane_off;        // this lane's ldmatrix / stmatrix row offset
    int dq_rnd_b;
    uint32_t clip_lo2, clip_hi2;
};
template <int N>
__device__ __forceinline__ MmaWinnerCtx make_mma_winner_ctx(const uint4* ctab, int lane, const FastQuant& fq, int maxv) {
    constexpr int PITCH = N * 2 + 16;
    MmaWinnerCtx c;
    c.ctab_lane = smem_u32(ctab + lane);
    c.lane_off = (((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + (lane >> 4) * 16;
    c.dq_rnd_b = fq.dq_rnd + (kOperandBits << fq.dq_shift);
    c.clip_lo2 = 0x08000800u;
    c.clip_hi2 = c.clip_lo2 + (uint32_t)(maxv <= 1023 ? maxv : 0) * 0x10001u;
    return c;
}
template <int N>
__device__ __forceinline__ void winner_mma(int lane, int64_t b, int mode, int16_t* O, unsigned char* ptile,
                                           const int16_t* top, const int16_t* left, const int16_t* neg,
                                           int dc, const FastQuant& fq, const MmaWinnerCtx& ctx,
                                           const CoderOut& out) {
    constexpr int PITCH = N * 2 + 16;
    static_assert(PITCH == CoderCfg<N, 32>::O_PITCH * 2, "O tile must have the ldmatrix pitch");
    if constexpr (N == 16) {   // two lanes per row: lane = (row, half)
        const int r = lane >> 1, x0 = (lane & 1) * 8;
        int p[8];
        predict_seg8_u8<N, 32>(mode, r, x0, top, left, neg, dc, p);
        const uint4 w = make_uint4(pack16(p[0], p[1]), pack16(p[2], p[3]), pack16(p[4], p[5]), pack16(p[6], p[7]));
        if (out.pred) stg_stream(out.pred + b * N * N + r * N + x0, w);
        *reinterpret_cast<uint4*>(ptile + r * PITCH + 2 * x0) = w;
    } else if (lane < N) {
        int p[N];
        uint32_t pw[N / 2];
        predict_row_u8<N, 32>(mode, lane, top, left, neg, dc, p);
        pack_row<N>(p, pw);
        if (out.pred) store_row16<N>(out.pred + b * N * N + lane * N, pw);
#pragma unroll
        for (int q = 0; q < N / 8; ++q)
            *reinterpret_cast<uint4*>(ptile + lane * PITCH + 16 * q) =
                make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
    }
    __syncwarp();
    const int fg = lane >> 2, ft = lane & 3;
    const uint32_t ctab_lane = ctx.ctab_lane;
    auto cv = [&](int v) -> uint4 { return ld_const_vec(ctab_lane, v); };
    mma_block_chain<N>(smem_u32(O) + ctx.lane_off, smem_u32(ptile) + ctx.lane_off, cv, out.coeff != nullptr,
                       out.coeff + b * N * N + (2 * ft) * N + fg, out.levels != nullptr,
                       out.levels + b * N * N + (2 * ft) * N + fg, fq, ctx.dq_rnd_b, ctx.clip_lo2, ctx.clip_hi2);
    __syncwarp();
    if (out.recon && lane < N) {  // block-major reconstruction (not requested by the plane coders)
#pragma unroll
        for (int q = 0; q < N / 8; ++q)
            stg_stream(out.recon + b * N * N + lane * N + 8 * q,
                       *reinterpret_cast<const uint4*>(reinterpret_cast<unsigned char*>(O) + lane * PITCH + 16 * q));
    }
}


struct CoderArgs {
    // SRC_ARRAYS
    const int16_t* orig;   // (B,N,N)
    const int16_t* top;    // (B,2N+1)
    const int16_t* left;   // (B,2N+1)
    const int16_t* top_left;  // (B,)
    const uint8_t* modes_in;  // (B,) or NULL
    int mode;
    int only_undecided;    // SRC_PLANE: skip the warp tiles whose modes_in are all decided (another kernel coded them)
    int vec_ok;            // SRC_PLANE: pitch % 8 == 0 and 16-byte aligned planes -> 128-bit pixel moves (N >= 16, G = 32)
    unsigned int* handed_back;  // with only_undecided: {tiles the other kernel left to this one, CTAs that have looked}
    // SRC_PLANE / SRC_WAVEFRONT
    const int16_t* src;    // (H, pitch)
    int H, W, pitch;
    int cost_kind;
    int* ticket;           // K8: row ticket counter
    int16_t* bottom;       // K8: [bh][W] bottom rows of the reconstructed blocks, -1 = not yet written
    unsigned poll_sleep_ns;  // K8: back-off between two polls of a row that is not ready yet
    // SRC_PLANE / SRC_WAVEFRONT: frames stacked `frame_stride` samples apart (src and recon_plane alike);
    // block-major outputs hold frame f at block offset f * blocks_per_frame, exchange rows at f * bh * W
    int n_frames;
    int64_t frame_stride;
    int64_t blocks_per_frame;
    // common
    int64_t n_blocks;   // blocks of all frames
    QuantParams qp;
    FastQuant fq;
    int maxv;
    int use_dst;
    CoderOut out;
};

// Resident CTAs per SM the kernel is compiled for.  The 16x16 tensor-core winner instance (one block per warp, from
// a plane) takes 246 registers when left alone: 2 CTAs per SM, issue slots 38 % busy, the stalls are fixed-latency
// waits.  Capped at 168 registers (3 CTAs, 32 bytes of spills) with a grid of 3 CTAs per SM: 407 -> 373 us for 8 4K
// frames.  (4 CTAs: 340 bytes of spills, slower; the 32x32 instance gains nothing from the same cap.)
#ifndef NH_WAVE1_PREFETCH
#define NH_WAVE1_PREFETCH 1
#endif
#ifndef NH_WINNER_OCC
#define NH_WINNER_OCC 3
#endif
template <int N, int G, int SRC>
constexpr int coder_occ() { return (N == 16 && G == 32 && SRC == 1) ? NH_WINNER_OCC : 1; }
template <int N, int G, int SRC>
__global__ void __launch_bounds__(SRC == SRC_WAVEFRONT ? 32 : 128, coder_occ<N, G, SRC>()) coder_kernel(const CoderArgs a) {
    using Cfg = CoderCfg<N, G>;
    constexpr int GPW = 32 / G;  // groups (blocks) per warp
    constexpr int WARPS = SRC == SRC_WAVEFRONT ? 1 : 4;
    __shared__ __align__(16) unsigned char smem[WARPS * GPW * Cfg::GROUP_BYTES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / G, gl = lane % G;
    unsigned char* base = smem + (warp * GPW + g) * Cfg::GROUP_BYTES;
    int16_t* top = reinterpret_cast<int16_t*>(base);
    int16_t* left = top + Cfg::REF_W;
    int16_t* neg = reinterpret_cast<int16_t*>(base + Cfg::REFS_PAD);
    int16_t* O = reinterpret_cast<int16_t*>(base + Cfg::REFS_PAD + Cfg::NEG_BYTES);
    int* M = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(O) + Cfg::O_BYTES);

    const int bw = SRC == SRC_ARRAYS ? 1 : a.W / N;
    const int bh = SRC == SRC_ARRAYS ? 1 : a.H / N;
    // 32x32 blocks (one per warp): the winner pipeline runs on the tensor cores when the block is 8-bit
    constexpr bool kMmaWinner = N >= 16 && G == 32 && SRC == SRC_PLANE || N == 32 && G == 32 && SRC == SRC_WAVEFRONT;
    constexpr int NM = kMmaWinner ? N : 32;
    __shared__ __align__(16) uint4 ctab[kMmaWinner ? MmaConsts<NM>::V_END : 1][32];
    MmaWinnerCtx mctx{};
    if constexpr (kMmaWinner) {
        stage_mma_consts<NM, WARPS * 32>(&ctab[0][0]);
        __syncthreads();
        mctx = make_mma_winner_ctx<NM>(&ctab[0][0], lane, a.fq, a.maxv);
    }

    if constexpr (SRC == SRC_PLANE) {
        // Called only to pick up what the tensor-core winner kernel handed back: usually nothing.  Every
        // CTA looks at the count, and the last one to look re-arms the pair for the next call.
        if (a.only_undecided && a.handed_back) {
            __shared__ unsigned int s_count;
            if (threadIdx.x == 0) {
                s_count = *reinterpret_cast<volatile unsigned int*>(a.handed_back);
                __threadfence();
                if (atomicAdd(a.handed_back + 1, 1u) == gridDim.x - 1) {
                    a.handed_back[0] = 0;
                    a.handed_back[1] = 0;
                }
            }
            __syncthreads();
            if (s_count == 0) return;
        }
    }
    if constexpr (SRC == SRC_WAVEFRONT) {
        // One warp per block row, rows handed out in order by a ticket counter so that a
        // waiting warp only ever waits on a row that a resident warp already owns.
        // Tickets interleave the frames (ticket t = row t / F of frame t % F), so all frames advance
        // together and the row above always holds a smaller ticket.
        int* ticket = a.ticket;
        for (;;) {
            int tk = 0;
            if (lane == 0) tk = atomicAdd(ticket, 1);
            tk = __shfl_sync(0xffffffffu, tk, 0);
            const int by = tk / a.n_frames, fr = tk - by * a.n_frames;
            if (by >= bh) break;
            const int16_t* srcf = a.src + fr * a.frame_stride;
            int16_t* reconf = a.out.recon_plane + fr * a.frame_stride;
            int16_t* bottomf = a.bottom + (int64_t)fr * bh * a.W;
            // original pixels do not depend on any neighbour: they are fetched one block ahead (the loads of block
            // bx + 1 are issued as soon as those of block bx sit in O and complete under its search)
            constexpr int OPL = (N * N + G - 1) / G;  // orig samples per lane
            int ov[OPL];
            auto fetch_px = [&](int bxn) {
#pragma unroll
                for (int i = 0; i < OPL; ++i) {
                    const int e = gl + i * G;
                    ov[i] = e < N * N ? (int)__ldg(srcf + (int64_t)(by * N + e / N) * a.pitch + bxn * N + e % N) : 0;
                }
            };
            fetch_px(0);
            for (int bx = 0; bx < bw; ++bx) {
                const int x = bx * N, y = by * N;
                const int64_t b = fr * a.blocks_per_frame + (int64_t)by * bw + bx;
                if (!NH_WAVE1_PREFETCH && bx > 0) fetch_px(bx);
                int ood = 0;  // any sample outside [0, 255] disables the packed 8-bit search
                // left references = right-most column of the block this warp has just reconstructed
                // (still in O); bottom-left is not reconstructed yet -> replicate (n_left = N)
                for (int k = gl + 1; k < Cfg::REF_W; k += G) {
                    const int kk = k <= N ? k : N;
                    const int lv = bx == 0 ? 128 : (int)O[(kk - 1) * Cfg::O_PITCH + (N - 1)];
                    left[k] = (int16_t)lv;
                    ood |= lv;
                }
                __syncwarp();  // O is about to be overwritten with the next block's pixels
                // Top references (corner, above, above-right) come from the exchange rows: every block
                // publishes its reconstructed bottom row there and -1 marks "not written yet", so the
                // data is its own flag -- one L2 round trip, no fence, no separate progress counter.
                // Reconstructed samples are clipped to [0, 2^bit_depth - 1], never negative.
                if (by == 0) {
                    for (int k = gl; k < Cfg::REF_W; k += G) top[k] = 128;
                    ood |= 0;
                } else {
                    const int16_t* up = bottomf + (int64_t)(by - 1) * a.W;
                    int last = x + 2 * N - 1;
                    if (last > a.W - 1) last = a.W - 1;
                    bool ready;
                    do {
                        ready = true;
                        for (int k = gl; k < Cfg::REF_W; k += G) {
                            int v;
                            if (k == 0 && x == 0) {
                                v = 128;
                            } else {
                                int col = x + k - 1;
                                if (col > last) col = last;
                                v = (int)__ldcg(up + col);
                            }
                            if (v < 0) ready = false;
                            else top[k] = (int16_t)v;
                        }
                        ready = __all_sync(0xffffffffu, ready);
                        // A spinning warp takes issue slots from the warps it waits for as soon as
                        // several frames share the SMs (frames on concurrent streams): back off.
                        if (!ready && a.poll_sleep_ns) __nanosleep(a.poll_sleep_ns);
                    } while (!ready);
                    for (int k = gl; k < Cfg::REF_W; k += G) ood |= (int)top[k];
                }
                if (gl == 0) left[0] = top[0];
#pragma unroll
                for (int i = 0; i < OPL; ++i) {
                    const int e = gl + i * G;
                    if (e < N * N) O[(e / N) * Cfg::O_PITCH + (e % N)] = (int16_t)ov[i];
                    ood |= ov[i];
                }
                if (NH_WAVE1_PREFETCH && bx + 1 < bw) fetch_px(bx + 1);
                const bool fast8 = !__any_sync(0xffffffffu, (ood & ~0xff) != 0);
                __syncwarp();
                const int corner = (int)top[0];
                const int dc = dc_from_refs_warp<N>(lane, top, left);   // G == 32: the whole warp codes one block
                int key;
                if (fast8) {
                    build_neg_arrays<N, G>(gl, top, left, neg);
                    __syncwarp();
                    key = search_modes_u8<N, G>(gl, O, top, left, neg, dc, a.cost_kind);
                } else {
                    key = search_modes<N, G>(gl, O, top, left, corner, dc, a.cost_kind);
                }
                const int mode = mode_of_key(key);
                if (gl == 0) {
                    if (a.out.modes) a.out.modes[b] = (uint8_t)mode;
                    if (a.out.costs) a.out.costs[b] = key >> 6;
                }
                bool done = false;
                if constexpr (kMmaWinner) {
                    if (fast8 && a.maxv <= 1023) {
                        winner_mma<32>(lane, b, mode, O, reinterpret_cast<unsigned char*>(M), top, left, neg, dc,
                                       a.fq, mctx, a.out);
                        done = true;
                    }
                }
                if (!done)
                    code_block<N, G>(gl, true, b, mode, O, M, top, left, corner, dc, a.qp, a.fq, fast8, neg,
                                     a.maxv, a.use_dst != 0, a.out);
                // publish the bottom row first (the row below is polling for it), then the plane
                for (int e = gl; e < N; e += G)
                    __stcg(bottomf + (int64_t)by * a.W + x + e, O[(N - 1) * Cfg::O_PITCH + e]);
                for (int e = gl; e < N * N; e += G)
                    reconf[(int64_t)(y + e / N) * a.pitch + x + e % N] = O[(e / N) * Cfg::O_PITCH + (e % N)];
                __syncwarp();
            }
        }
        return;
    } else {
        const int64_t n_tiles = (a.n_blocks + GPW - 1) / GPW;
        // One block per warp at N >= 16: the pixels move as 16-byte chunks (chunk c = row c / (N/8)), the
        // next block's chunks are fetched while this one is coded.
        constexpr bool kVec = SRC == SRC_PLANE && G == 32 && N >= 16;
        constexpr int CPL = kVec ? N * N / 8 / 32 : 1;   // chunks per lane
        constexpr int CW = N >= 8 ? N / 8 : 1;           // chunks per row
        const bool vec = kVec && a.vec_ok && !a.only_undecided;
        constexpr int RPL = (Cfg::REF_W + 31) / 32;      // reference entries per lane
        uint4 nxt[CPL];
        int ntv[RPL], nlv[RPL];
        int nmode = 0xFF;
        // plane sources: block b -> frame b / blocks_per_frame, raster position inside the frame
        auto locate = [&](int64_t bb, int& fx, int& fy) -> const int16_t* {
            const int fr = (int)(bb / a.blocks_per_frame);
            const int64_t lb = bb - fr * a.blocks_per_frame;
            fx = (int)(lb % bw) * N;
            fy = (int)(lb / bw) * N;
            return a.src + fr * a.frame_stride;
        };
        auto fetch = [&](int64_t t) {
            int fx, fy;
            const int16_t* srcf = locate(t, fx, fy);
            if (a.modes_in) nmode = (int)a.modes_in[t];   // one block per warp: block index = tile index
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i, row = c / CW, c8 = c % CW;
                nxt[i] = __ldg(reinterpret_cast<const uint4*>(srcf + (int64_t)(fy + row) * a.pitch + fx + 8 * c8));
            }
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int k = lane + 32 * i, kk = k <= 2 * N ? k : 2 * N;
                ntv[i] = top_ref<false>(srcf, a.H, a.W, a.pitch, fx, fy, 2 * N, kk);
                nlv[i] = left_ref<false>(srcf, a.H, a.W, a.pitch, fx, fy, 2 * N, kk);
            }
        };
        if (vec && (int64_t)blockIdx.x * WARPS + warp < n_tiles) fetch((int64_t)blockIdx.x * WARPS + warp);
        // several blocks per warp: at least the decided mode of the next tile is fetched a tile ahead
        auto load_mode = [&](int64_t t) -> int {
            const int64_t bb = t * GPW + g;
            return (a.modes_in && bb < a.n_blocks) ? (int)a.modes_in[bb] : (a.modes_in ? 1 : 0xFF);
        };
        // ... and at N <= 8 its references and pixels too (scalar loads, a handful of registers per lane)
        constexpr bool kSp = SRC == SRC_PLANE && N <= 8;
        constexpr int RS = kSp ? (Cfg::REF_W + G - 1) / G : 1, PS = kSp ? N * N / G : 1;
        const bool sp = kSp && !a.only_undecided;
        int stv[RS], slv[RS], spx[PS];
        auto sfetch = [&](int64_t t) {
            const int64_t bb = t * GPW + g;
            if (bb < a.n_blocks) {
                int fx, fy;
                const int16_t* srcf = locate(bb, fx, fy);
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    const int k = gl + i * G, kk = k <= 2 * N ? k : 2 * N;
                    stv[i] = top_ref<false>(srcf, a.H, a.W, a.pitch, fx, fy, 2 * N, kk);
                    slv[i] = left_ref<false>(srcf, a.H, a.W, a.pitch, fx, fy, 2 * N, kk);
                }
#pragma unroll
                for (int i = 0; i < PS; ++i) {
                    const int e = gl + i * G;
                    spx[i] = __ldg(srcf + (int64_t)(fy + e / N) * a.pitch + fx + e % N);
                }
            }
        };
        if (sp && (int64_t)blockIdx.x * WARPS + warp < n_tiles) sfetch((int64_t)blockIdx.x * WARPS + warp);
        int pmode = 0xFF;
        if (SRC == SRC_PLANE && !vec && (int64_t)blockIdx.x * WARPS + warp < n_tiles)
            pmode = load_mode((int64_t)blockIdx.x * WARPS + warp);
        for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < n_tiles;
             tile += (int64_t)gridDim.x * WARPS) {
            const int64_t b = tile * GPW + g;
            const bool valid = b < a.n_blocks;
            int corner = 0, x = 0, y = 0, ood = 0;
            const int16_t* srcf = a.src;
            int16_t* reconf = a.out.recon_plane;
            int mode_in = 0xFF;
            bool given = false;
            if constexpr (SRC == SRC_PLANE) {
                // modes decided by search_plane_kernel (nh_search.cuh); 0xFF = not decided (the tile held a
                // sample outside [0, 255]) -> the exact search below, for every block of this warp
                if (vec) {
                    mode_in = nmode;   // fetched with the pixels during the previous block
                } else {
                    mode_in = pmode;
                    const int64_t tn = tile + (int64_t)gridDim.x * WARPS;
                    if (tn < n_tiles) pmode = load_mode(tn);
                }
                given = !__any_sync(0xffffffffu, mode_in > 34);
                if (given && a.only_undecided) continue;
            }
            if constexpr (SRC == SRC_ARRAYS) {
                if (valid) {
                    for (int k = gl; k <= 2 * N; k += G) {
                        top[k] = a.top[b * (2 * N + 1) + k];
                        left[k] = a.left[b * (2 * N + 1) + k];
                    }
                    for (int e = gl; e < N * N; e += G)
                        O[(e / N) * Cfg::O_PITCH + (e % N)] = __ldg(a.orig + b * N * N + e);
                    corner = (int)a.top_left[b];
                }
            } else {
                if (valid) {
                    srcf = locate(b, x, y);
                    if (reconf) reconf += srcf - a.src;
                    if (vec) {
#pragma unroll
                        for (int i = 0; i < RPL; ++i) {
                            const int k = lane + 32 * i;
                            if (k < Cfg::REF_W) {
                                top[k] = (int16_t)ntv[i];
                                left[k] = (int16_t)nlv[i];
                                ood |= ntv[i] | nlv[i];
                            }
                        }
                    } else if (sp) {
#pragma unroll
                        for (int i = 0; i < RS; ++i) {
                            const int k = gl + i * G;
                            if (k < Cfg::REF_W) {
                                top[k] = (int16_t)stv[i];
                                left[k] = (int16_t)slv[i];
                                ood |= stv[i] | slv[i];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < PS; ++i) {
                            const int e = gl + i * G;
                            O[(e / N) * Cfg::O_PITCH + (e % N)] = (int16_t)spx[i];
                            ood |= spx[i];
                        }
                    } else {
#pragma unroll
                        for (int k = gl; k < Cfg::REF_W; k += G) {   // unrolled: all loads in flight together
                            const int kk = k <= 2 * N ? k : 2 * N;
                            const int tv = top_ref<false>(srcf, a.H, a.W, a.pitch, x, y, 2 * N, kk);
                            const int lv = left_ref<false>(srcf, a.H, a.W, a.pitch, x, y, 2 * N, kk);
                            top[k] = (int16_t)tv;
                            left[k] = (int16_t)lv;
                            ood |= tv | lv;
                        }
                    }
                    if (vec) {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const int c = lane + 32 * i, row = c / CW, c8 = c % CW;
                            *reinterpret_cast<uint4*>(O + row * Cfg::O_PITCH + 8 * c8) = nxt[i];
                            ood |= (int)((nxt[i].x | nxt[i].y | nxt[i].z | nxt[i].w) & 0xFF00FF00u);
                        }
                        const int64_t tn = tile + (int64_t)gridDim.x * WARPS;
                        if (tn < n_tiles) fetch(tn);
                    } else if (!sp) {
#pragma unroll (N <= 8 ? N : 4)
                        for (int e = gl; e < N * N; e += G) {
                            const int v = __ldg(srcf + (int64_t)(y + e / N) * a.pitch + x + e % N);
                            O[(e / N) * Cfg::O_PITCH + (e % N)] = (int16_t)v;
                            ood |= v;
                        }
                    }
                }
            }
            if constexpr (kSp) {
                if (sp) {
                    const int64_t tn = tile + (int64_t)gridDim.x * WARPS;
                    if (tn < n_tiles) sfetch(tn);
                }
            }
            if (!valid) {  // keep shared memory defined for the idle groups of a ragged tile
                for (int k = gl; k < Cfg::REF_W; k += G) top[k] = left[k] = 0;
                for (int e = gl; e < N * N; e += G) O[(e / N) * Cfg::O_PITCH + (e % N)] = 0;
            }
            const bool fast8 = !__any_sync(0xffffffffu, (ood & ~0xff) != 0);
            __syncwarp();
            if constexpr (SRC == SRC_PLANE) corner = (int)top[0];
            int dc;
            if constexpr (G == 32 && N >= 16) {   // one block per warp: lane k adds top[1+k] + left[1+k]
                const int s = lane < N ? (int)top[1 + lane] + (int)left[1 + lane] : 0;
                dc = dc_value<N>(__reduce_add_sync(0xffffffffu, s));
            } else {
                dc = dc_from_refs<N>(top, left);
            }
            int mode;
            if constexpr (SRC == SRC_ARRAYS) {
                mode = (valid && a.modes_in) ? (int)a.modes_in[b] : a.mode;
                if (mode > 34) mode = 34;  // launcher validates the scalar; clamp per-block garbage
            } else {
                mode = mode_in;
                if (given) {
                    if (fast8) {
                        build_neg_array_of_mode<N, G>(gl, mode, top, left, neg);
                        __syncwarp();
                    }
                } else {
                    int key;
                    if (fast8) {  // warp-uniform: the shuffles inside both searches span the warp
                        build_neg_arrays<N, G>(gl, top, left, neg);
                        __syncwarp();
                        key = search_modes_u8<N, G>(gl, O, top, left, neg, dc, a.cost_kind);
                    } else {
                        key = search_modes<N, G>(gl, O, top, left, corner, dc, a.cost_kind);
                    }
                    mode = mode_of_key(key);
                    if (valid && gl == 0) {
                        if (a.out.modes) a.out.modes[b] = (uint8_t)mode;
                        if (a.out.costs) a.out.costs[b] = key >> 6;
                    }
                }
            }
            bool done = false;
            if constexpr (kMmaWinner) {
                if (fast8 && valid && a.maxv <= 1023) {  // one block per warp: valid is warp-uniform
                    winner_mma<NM>(lane, b, mode, O, reinterpret_cast<unsigned char*>(M), top, left, neg, dc, a.fq,
                                   mctx, a.out);
                    done = true;
                }
            }
            if (!done)
                code_block<N, G>(gl, valid, b, mode, O, M, top, left, corner, dc, a.qp, a.fq,
                                 SRC != SRC_ARRAYS && fast8, neg, a.maxv, a.use_dst != 0, a.out);
            if constexpr (SRC == SRC_PLANE) {
                if (valid && a.out.recon_plane) {
                    if (vec) {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const int c = lane + 32 * i, row = c / CW, c8 = c % CW;
                            stg_stream(reconf + (int64_t)(y + row) * a.pitch + x + 8 * c8,
                                       *reinterpret_cast<const uint4*>(O + row * Cfg::O_PITCH + 8 * c8));
                        }
                    } else {
                        for (int e = gl; e < N * N; e += G)
                            reconf[(int64_t)(y + e / N) * a.pitch + x + e % N] = O[(e / N) * Cfg::O_PITCH + (e % N)];
                    }
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace nh
#include "nh_winner16.cuh"
namespace nh {

// K8, N = 16 / 32: WPB warps per block row.  With one warp per row the critical path of the wavefront
// is bw + 2 bh block times, and a 32x32 block is ~17 us of dependent work for a single warp, most of
// it the 35-mode search.  Here a CTA owns the row: every warp searches a share of the candidate
// modes (the argmin is merged through shared memory; the key order keeps the tie rule), warp 0 polls
// the exchange row above and runs the winner pipeline, all threads move pixels.  Same exchange-row
// protocol and ticket order as the one-warp kernel, so a waiting CTA only ever waits on a row that a
// resident CTA owns.
// Per-phase cycle counters of the first block row (development only, `make prof`; see nh_wave.cuh)
#ifdef NH_WAVE_PROF
#define NH_MWPROF_DECL long long mwp_t = clock64(), mwp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define NH_MWPROF_MARK(i) { const long long now__ = clock64(); mwp_acc[i] += now__ - mwp_t; mwp_t = now__; }
#define NH_MWPROF_DUMP(cond) if (cond) { for (int i__ = 0; i__ < 8; ++i__) atomicAdd(reinterpret_cast<unsigned long long*>(a.ticket) + 8 + i__, (unsigned long long)mwp_acc[i__]); }
#else
#define NH_MWPROF_DECL
#define NH_MWPROF_MARK(i)
#define NH_MWPROF_DUMP(cond)
#endif
// resident CTAs per SM the 32x32 instances are held to (168 registers: what they took before the look-ahead poll)
template <int N, int WPB>
constexpr int mw_occ() { return N == 32 ? (WPB == 4 ? 3 : WPB == 2 ? 6 : 1) : 1; }
template <int N, int WPB>
__global__ void __launch_bounds__(32 * WPB, mw_occ<N, WPB>()) coder_wave_mw_kernel(const CoderArgs a) {
    using Cfg = CoderCfg<N, 32>;
    constexpr int T = 32 * WPB;
    using MC = MmaConsts<N>;
    static_assert(Cfg::M_BYTES >= N * (N * 2 + 16), "the prediction tile aliases the working matrix");
    __shared__ __align__(16) unsigned char smem[Cfg::GROUP_BYTES];
    __shared__ __align__(16) uint4 ctab[MC::V_END][32];        // per-lane MMA constants
    __shared__ int s_keys[WPB];
    __shared__ int s_row;
    __shared__ __align__(16) int16_t s_top2[(Cfg::REF_W + 7) / 8 * 8];   // top references of the NEXT block (polled ahead)
    __shared__ __align__(16) unsigned char s_ob[2 * N * N];              // the block's pixels as bytes, and transposed
    static_assert(WPB >= 2, "warp 1 polls ahead while warp 0 codes the winner");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    stage_mma_consts<N, T>(&ctab[0][0]);  // made visible by the first barrier of the row loop
    const MmaWinnerCtx mctx = make_mma_winner_ctx<N>(&ctab[0][0], lane, a.fq, a.maxv);
    const bool mma_ok = a.maxv <= 1023;
    int16_t* const top = reinterpret_cast<int16_t*>(smem);
    int16_t* const left = top + Cfg::REF_W;
    int16_t* neg = reinterpret_cast<int16_t*>(smem + Cfg::REFS_PAD);
    int16_t* O = reinterpret_cast<int16_t*>(smem + Cfg::REFS_PAD + Cfg::NEG_BYTES);
    int* M = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(O) + Cfg::O_BYTES);
    const int bw = a.W / N, bh = a.H / N;
    constexpr int OPL = (N * N + T - 1) / T;  // original samples per thread
    for (;;) {
        if (tid == 0) s_row = atomicAdd(a.ticket, 1);
        __syncthreads();
        const int tk = s_row;   // frames interleaved: ticket t = row t / F of frame t % F
        __syncthreads();  // everyone has read s_row before thread 0 draws the next ticket
        const int by = tk / a.n_frames, fr = tk - by * a.n_frames;
        if (by >= bh) break;
        const int16_t* srcf = a.src + fr * a.frame_stride;
        int16_t* reconf = a.out.recon_plane + fr * a.frame_stride;
        int16_t* bottomf = a.bottom + (int64_t)fr * bh * a.W;
        // original pixels do not depend on any neighbour: they are fetched one block ahead (the loads of block bx + 1
        // are issued as soon as those of block bx have been stored into O, and complete under its search)
        int ov[OPL];
        auto fetch_px = [&](int bxn) {
#pragma unroll
            for (int i = 0; i < OPL; ++i) {
                const int e = tid + i * T;
                ov[i] = e < N * N ? (int)__ldg(srcf + (int64_t)(by * N + e / N) * a.pitch + bxn * N + e % N) : 0;
            }
        };
        fetch_px(0);
        // one warp fetches the top references of block bxn into dst; returns the OR of the samples (domain check)
        auto poll_top = [&](int bxn, int16_t* dst) -> int {
            int o = 0;
            if (by == 0) {
                for (int k = lane; k < Cfg::REF_W; k += 32) dst[k] = 128;
                return 128;
            }
            const int xn = bxn * N;
            const int16_t* up = bottomf + (int64_t)(by - 1) * a.W;
            int last = xn + 2 * N - 1;
            if (last > a.W - 1) last = a.W - 1;
            bool ready;
            do {
                ready = true;
                for (int k = lane; k < Cfg::REF_W; k += 32) {
                    int v;
                    if (k == 0 && xn == 0) {
                        v = 128;
                    } else {
                        int col = xn + k - 1;
                        if (col > last) col = last;
                        v = (int)__ldcg(up + col);
                    }
                    if (v < 0) ready = false;
                    else dst[k] = (int16_t)v;
                }
                ready = __all_sync(0xffffffffu, ready);
                if (!ready && a.poll_sleep_ns) __nanosleep(a.poll_sleep_ns);
            } while (!ready);
            for (int k = lane; k < Cfg::REF_W; k += 32) o |= (int)dst[k];
            return o;
        };
        int ood_top_next = 0;   // warp 1: domain check of the references polled ahead
        NH_MWPROF_DECL
        for (int bx = 0; bx < bw; ++bx) {
            NH_MWPROF_MARK(7)
            const int x = bx * N, y = by * N;
            const int64_t b = fr * a.blocks_per_frame + (int64_t)by * bw + bx;
            int ood = 0;
            // left references = right-most column of the block this CTA has just reconstructed (still
            // in O); bottom-left is not reconstructed yet -> replicate (n_left = N)
            for (int k = tid + 1; k < Cfg::REF_W; k += T) {
                const int kk = k <= N ? k : N;
                const int lv = bx == 0 ? 128 : (int)O[(kk - 1) * Cfg::O_PITCH + (N - 1)];
                left[k] = (int16_t)lv;
                ood |= lv;
            }
            // top references from the exchange row above (data-as-flag, -1 = not written yet).  Block 0 of a row:
            // warp 0 polls here; every later block: warp 1 has polled them into s_top2 while warp 0 coded the previous
            // winner (the L2 round trip of the poll is then off the row's critical path whenever the row above is
            // ahead), and they are copied over here.  The other warps wait at the barrier below.
            if (bx == 0) {
                if (warp == 0) {
                    ood |= poll_top(0, top);
                    if (lane == 0) left[0] = top[0];   // lane 0 wrote top[0] itself
                }
            } else if (warp == 1) {
                for (int k = lane; k < Cfg::REF_W; k += 32) top[k] = s_top2[k];
                if (lane == 0) left[0] = s_top2[0];
                ood |= ood_top_next;
            }
            __syncthreads();  // O (previous reconstruction) has been consumed, top / left are in place
            NH_MWPROF_MARK(0)
            // the per-mode arrays of the negative angles only need the references: built beside the pixel stores, the
            // barrier of the domain vote below publishes both (unused when the block leaves the 8-bit domain)
            build_neg_arrays<N, T>(tid, top, left, neg);
#pragma unroll
            for (int i = 0; i < OPL; ++i) {
                const int e = tid + i * T;
                if (e < N * N) {
                    O[(e / N) * Cfg::O_PITCH + (e % N)] = (int16_t)ov[i];
                    s_ob[e] = (unsigned char)ov[i];                               // only read on the 8-bit path
                    s_ob[N * N + (e % N) * N + e / N] = (unsigned char)ov[i];
                }
                ood |= ov[i];
            }
            if (bx + 1 < bw) fetch_px(bx + 1);
            const bool fast8 = __syncthreads_or((ood & ~0xff) != 0) == 0;
            NH_MWPROF_MARK(1)
            const int corner = (int)top[0];
            const int dc = dc_from_refs_warp<N>(lane, top, left);
            int key = fast8 ? search_modes_u8_pk<N, 32>(lane, s_ob, s_ob + N * N, top, left, neg, dc, a.cost_kind, warp, WPB)
                            : search_modes<N, 32>(lane, O, top, left, corner, dc, a.cost_kind, warp, WPB);
            NH_MWPROF_MARK(2)
            if (lane == 0) s_keys[warp] = key;
            __syncthreads();
            NH_MWPROF_MARK(3)
#pragma unroll
            for (int w = 0; w < WPB; ++w) key = s_keys[w] < key ? s_keys[w] : key;
            const int mode = mode_of_key(key);
            if (warp == 0) {
                if (lane == 0) {
                    if (a.out.modes) a.out.modes[b] = (uint8_t)mode;
                    if (a.out.costs) a.out.costs[b] = key >> 6;
                }
                if (fast8 && mma_ok) {  // winner pipeline on the tensor cores (8-bit samples)
                    winner_mma<N>(lane, b, mode, O, reinterpret_cast<unsigned char*>(M), top, left, neg, dc, a.fq,
                                  mctx, a.out);
                } else {
                    code_block<N, 32>(lane, true, b, mode, O, M, top, left, corner, dc, a.qp, a.fq, fast8, neg,
                                      a.maxv, a.use_dst != 0, a.out);
                }
                // publish the bottom row first: the row below is polling for it
                for (int e = lane; e < N; e += 32)
                    __stcg(bottomf + (int64_t)by * a.W + x + e, O[(N - 1) * Cfg::O_PITCH + e]);
            } else if (warp == 1 && bx + 1 < bw) {
                ood_top_next = poll_top(bx + 1, s_top2);
            }
            NH_MWPROF_MARK(4)
            __syncthreads();  // the reconstruction is in O (and the next block's top references in s_top2)
            NH_MWPROF_MARK(5)
            for (int e = tid; e < N * N; e += T)
                reconf[(int64_t)(y + e / N) * a.pitch + x + e % N] = O[(e / N) * Cfg::O_PITCH + (e % N)];
            // no barrier needed here: the next block only reads O (left references) before the
            // barrier that precedes its overwrite
            NH_MWPROF_MARK(6)
        }
        NH_MWPROF_DUMP(by == 0 && fr == 0 && tid == 0)
    }
}

}  // namespace nh
#include "nh_wave.cuh"   // latency-oriented wavefront kernels for 8-bit planes (N = 8, N = 4); needs CoderArgs
#include "nh_search2.cuh"
#include "nh_search3.cuh"   // line-synchronous search kernel (N = 8 / 16 / 32)
#include "nh_search4.cuh"   // SATD search with the Hadamard transforms on the tensor cores
namespace nh {

// Exchange rows of the wavefront coder: -1 where a block will publish its bottom row, 0 in the
// columns no full block covers (the reference reads the zero-initialised plane there).
__global__ void __launch_bounds__(256) init_bottom_kernel(int16_t* bottom, int64_t rows, int W, int covered, int* ticket) {
    const int64_t total = rows * W;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        bottom[t] = (int)(t % W) < covered ? (int16_t)-1 : (int16_t)0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0;
}

// Per-frame statistics of coded frames, one launch for all of them (grid.y = frame):
//   stats[f] = { sum (src - recon)^2 over the WHOLE plane (metrics.py:7-21 numerator; uncovered rows count,
//               __main__.py:135-137), H * W, sum of the winners' costs, number of non-zero levels }.
__global__ void __launch_bounds__(256)
    frame_stats_kernel(const int16_t* __restrict__ src, const int16_t* __restrict__ recon, int64_t frame_stride, int H,
                       int W, int pitch, int vec_ok, const int32_t* __restrict__ costs,
                       const int32_t* __restrict__ levels, int64_t blocks_per_frame, int nn, int64_t* __restrict__ stats) {
    const int fr = blockIdx.y;
    const int16_t* a = src + fr * frame_stride;
    const int16_t* b = recon + fr * frame_stride;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    long long sse = 0, nnz = 0, cost = 0;
    auto acc2 = [&](uint32_t wa, uint32_t wb) {
        const int d0 = lo16(wa) - lo16(wb), d1 = hi16(wa) - hi16(wb);
        sse += (long long)d0 * d0 + (long long)d1 * d1;
    };
    if (vec_ok) {
        const int w8 = W / 8;
        for (int64_t t = tid; t < (int64_t)H * w8; t += stride) {
            const int64_t o = (t / w8) * pitch + (t % w8) * 8;
            const uint4 va = ldg_stream(a + o), vb = ldg_stream(b + o);
            acc2(va.x, vb.x); acc2(va.y, vb.y); acc2(va.z, vb.z); acc2(va.w, vb.w);
        }
        const int tail = W - w8 * 8;
        for (int64_t t = tid; t < (int64_t)H * tail; t += stride) {
            const int64_t o = (t / tail) * pitch + w8 * 8 + t % tail;
            const int d = (int)a[o] - (int)b[o];
            sse += (long long)d * d;
        }
    } else {
        for (int64_t t = tid; t < (int64_t)H * W; t += stride) {
            const int64_t o = (t / W) * pitch + t % W;
            const int d = (int)a[o] - (int)b[o];
            sse += (long long)d * d;
        }
    }
    if (levels) {
        const int32_t* lv = levels + fr * blocks_per_frame * nn;   // 16-byte aligned: nn >= 16
        const int64_t n4 = blocks_per_frame * nn / 4;
        for (int64_t i = tid; i < n4; i += stride) {
            const uint4 v = ldg_stream(lv + 4 * i);
            nnz += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
        }
    }
    if (costs)
        for (int64_t i = tid; i < blocks_per_frame; i += stride) cost += costs[fr * blocks_per_frame + i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sse += __shfl_xor_sync(0xffffffffu, sse, off);
        nnz += __shfl_xor_sync(0xffffffffu, nnz, off);
        cost += __shfl_xor_sync(0xffffffffu, cost, off);
    }
    __shared__ long long sh[3][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = sse; sh[1][warp] = cost; sh[2][warp] = nnz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = 0, t1 = 0, t2 = 0;
        for (int w = 0; w < 8; ++w) { t0 += sh[0][w]; t1 += sh[1][w]; t2 += sh[2][w]; }
        unsigned long long* o = reinterpret_cast<unsigned long long*>(stats + 4 * fr);
        atomicAdd(o, (unsigned long long)t0);
        atomicAdd(o + 2, (unsigned long long)t1);
        atomicAdd(o + 3, (unsigned long long)t2);
        if (blockIdx.x == 0) stats[4 * fr + 1] = (int64_t)H * W;
    }
}

template <int N, int G, int SRC>
static int launch_coder(const CoderArgs& a, int grid, cudaStream_t st) {
    coder_kernel<N, G, SRC><<<grid, SRC == SRC_WAVEFRONT ? 32 : 128, 0, st>>>(a);
    NH_CHECK_LAUNCH("coder_kernel");
    return NH_OK;
}

// How the wavefront coder (recon_neighbours = 1) is laid out on the GPU.  Every kernel below gives the same results;
// which one is fastest depends on the block rows in flight (rows of a frame x frames of the call), so the default
// picks per call.  nh_set_wave_impl(warps, build) forces a choice for the calling thread (tests, profiling), the
// environment gives a thread's initial setting: NH_WAVE_WARPS=1|2|4|8 (N = 16 / 32), NH_WAVE4=1|4 (N = 4),
// NH_WAVE_OCC=lat|thr|hi (N = 4 / 8 builds).
struct WaveImpl { int warps, build; bool init; };
static thread_local WaveImpl g_wave_impl = {0, 0, false};
static const WaveImpl& wave_impl() {
    if (!g_wave_impl.init) {
        const char* w = getenv("NH_WAVE_WARPS");
        const char* w4 = getenv("NH_WAVE4");
        const char* o = getenv("NH_WAVE_OCC");
        if (w && atoi(w) == 12) g_wave_impl.warps = 12;
        else if (w && (w[0] == '1' || w[0] == '2' || w[0] == '4' || w[0] == '8')) g_wave_impl.warps = w[0] - '0';
        else if (w4 && (w4[0] == '1' || w4[0] == '4')) g_wave_impl.warps = w4[0] - '0';
        if (o) g_wave_impl.build = o[0] == 'l' ? 1 : o[0] == 'h' ? 3 : 2;
        g_wave_impl.init = true;
    }
    return g_wave_impl;
}
// Warps per block row at N = 16 / 32, from the rows in flight R (measured on one B200, 1 ... 32 4K frames per call,
// profiles/r5_wave_warps.txt): a row's warps share the candidate modes, so more warps shorten the dependent block
// time, but the resident rows per SM drop (8 warps: 1 CTA per SM, 4: 2-3, 2: 4-6, 1: 8+), and once a call has more
// rows in flight than fit, the rows that wait cost more than the slower block.  Thresholds re-measured with the
// look-ahead / packed-search kernels (profiles/r5_wave_more.txt, last sweep): N = 16: 8 warps up to two 4K frames, 4 up
// to ~5, then 2 (the one-warp kernel no longer wins anywhere up to 48 frames; 32 frames: 10.6 -> 30.9 Gpix/s);
// N = 32: 12 warps up to ~8 frames, 4 up to ~44, then 1 (32 frames: 23.9 -> 43.4 Gpix/s).
static int wave_warps_for(int size, int64_t rows) {
    const int64_t sm = sm_count();
    if (size == 16) return rows * 10 <= sm * 25 ? 8 : rows * 10 <= sm * 47 ? 4 : 2;
    // N = 32: one candidate per warp iteration, 35 iterations: 12 warps run 3 each where 8 run 5 (one 4K frame 1.244 ->
    // 1.202 ms, 8 frames 27.5 (4 warps) -> 28.0 Gpix/s; 9 / 10 / 16 / 18 warps are slower than 8)
    return rows <= sm * 4 ? 12 : rows <= sm * 20 ? 4 : 1;
}

template <int SRC>
static int dispatch_coder(const CoderArgs& a, int size, cudaStream_t st) {
    if (SRC == SRC_WAVEFRONT) {
        const int64_t rows = (int64_t)(a.H / size) * a.n_frames;
        const WaveImpl& wi = wave_impl();
        int grid = rows < sm_count() * 16 ? (int)rows : sm_count() * 16;
        if (grid < 1) grid = 1;
        // 8-bit planes, N = 8: the latency-oriented kernel of nh_wave.cuh (NH_WAVE_IMPL=1 keeps the generic one)
        static const bool wave_new = [] { const char* e = getenv("NH_WAVE_IMPL"); return !(e && e[0] == '1'); }();
        const bool wave_ok = wave_new && a.maxv <= 255 && (a.pitch % 4) == 0 && (a.frame_stride % 4) == 0 &&
                             (reinterpret_cast<uintptr_t>(a.src) & 7) == 0 &&
                             (reinterpret_cast<uintptr_t>(a.out.recon_plane) & 7) == 0;
        if (wave_ok && size == 8) {
            // few rows (one or two frames): the latency build, registers to spare (3 CTAs per SM); many rows: the build
            // that keeps 4 CTAs resident; SAD with more than ~5 4K frames in flight: 5 CTAs per SM (96 registers, 28
            // bytes of spills; 8 / 16 / 32 frames 20.6 / 23.3 / 24.6 -> 22.1 / 24.7 / 26.1 Gpix/s, fewer frames and SATD
            // lose or stay; 6 CTAs at 80 registers: slower everywhere; profiles/r5_wave8_occ.txt).
            const int build = wi.build ? wi.build
                              : rows <= (int64_t)sm_count() * 3 ? 1
                              : (a.cost_kind == NH_COST_SAD && rows > (int64_t)sm_count() * 8) ? 3 : 2;
            if (build == 1) {
                if (a.cost_kind == NH_COST_SAD) wave8_kernel<NH_COST_SAD, 3><<<grid, 128, 0, st>>>(a);
                else wave8_kernel<NH_COST_SATD, 3><<<grid, 128, 0, st>>>(a);
            } else if (build == 3) {
                if (a.cost_kind == NH_COST_SAD) wave8_kernel<NH_COST_SAD, 5><<<grid, 128, 0, st>>>(a);
                else wave8_kernel<NH_COST_SATD, 5><<<grid, 128, 0, st>>>(a);
            } else {
                if (a.cost_kind == NH_COST_SAD) wave8_kernel<NH_COST_SAD, 4><<<grid, 128, 0, st>>>(a);
                else wave8_kernel<NH_COST_SATD, 4><<<grid, 128, 0, st>>>(a);
            }
            NH_CHECK_LAUNCH("wave8_kernel");
            return NH_OK;
        }
        if (wave_ok && size == 4) {
            // up to ~4 4K frames in flight: four warps per block row (one 4K frame 4.16 -> 2.57 ms); beyond that the
            // one-warp kernel, whose 24 resident rows per SM give the higher batch rate (32 frames: 14.7 vs 11.5
            // Gpix/s, profiles/r2_wave4_probe.jsonl).  nh_set_wave_impl(1 | 4, ...) / NH_WAVE4=1|4 forces the one-warp / four-warp kernel.
            const bool one_warp = wi.warps ? wi.warps == 1 : rows > (int64_t)sm_count() * 16;
            if (one_warp) {
                if (a.cost_kind == NH_COST_SAD) wave4_kernel<NH_COST_SAD><<<grid, 32, 0, st>>>(a);
                else wave4_kernel<NH_COST_SATD><<<grid, 32, 0, st>>>(a);
                NH_CHECK_LAUNCH("wave4_kernel");
                return NH_OK;
            }
            // few rows: the latency build; many rows: the build that keeps twice the CTAs resident (rows in flight
            // x 16 pixels per dependent block time is what bounds a batch).  nh_set_wave_impl(..., 1 | 2) / NH_WAVE_OCC=lat|thr forces one of them.
            const bool lat4 = wi.build ? wi.build == 1 : rows <= (int64_t)sm_count() * 8;   // two 4K frames: 3.35 -> 2.93 ms, four: 4.70 / 3.99 ms
            if (lat4) {
                if (a.cost_kind == NH_COST_SAD) wave4mw_kernel<NH_COST_SAD, 4><<<grid, 128, 0, st>>>(a);
                else wave4mw_kernel<NH_COST_SATD, 4><<<grid, 128, 0, st>>>(a);
            } else {
                if (a.cost_kind == NH_COST_SAD) wave4mw_kernel<NH_COST_SAD, 8><<<grid, 128, 0, st>>>(a);
                else wave4mw_kernel<NH_COST_SATD, 8><<<grid, 128, 0, st>>>(a);
            }
            NH_CHECK_LAUNCH("wave4mw_kernel");
            return NH_OK;
        }
        int wave_warps = size >= 16 ? (wi.warps ? wi.warps : wave_warps_for(size, rows)) : 1;
        if (size == 16 && wave_warps == 12) wave_warps = 8;   // 12 warps per row exist at N = 32 only
        if (size >= 16 && wave_warps > 1) {
            // (measured and not kept, profiles/r5_wave_more.txt: register caps for more resident rows -- 168 / 128 registers,
            // up to 8 CTAs per SM -- gain nothing over the choice of warps per row, and the grid size beyond the resident
            // CTAs makes no difference)
            if (grid > sm_count() * 4) grid = sm_count() * 4;
            if (size == 16) {
                if (wave_warps == 2) coder_wave_mw_kernel<16, 2><<<grid, 64, 0, st>>>(a);
                else if (wave_warps == 8) coder_wave_mw_kernel<16, 8><<<grid, 256, 0, st>>>(a);
                else coder_wave_mw_kernel<16, 4><<<grid, 128, 0, st>>>(a);
            } else {
                if (wave_warps == 2) coder_wave_mw_kernel<32, 2><<<grid, 64, 0, st>>>(a);
                else if (wave_warps == 12) coder_wave_mw_kernel<32, 12><<<grid, 384, 0, st>>>(a);
                else if (wave_warps == 8) coder_wave_mw_kernel<32, 8><<<grid, 256, 0, st>>>(a);
                else coder_wave_mw_kernel<32, 4><<<grid, 128, 0, st>>>(a);
            }
            NH_CHECK_LAUNCH("coder_wave_mw_kernel");
            return NH_OK;
        }
        switch (size) {
            case 4: return launch_coder<4, 32, SRC_WAVEFRONT>(a, grid, st);
            case 8: return launch_coder<8, 32, SRC_WAVEFRONT>(a, grid, st);
            case 16: return launch_coder<16, 32, SRC_WAVEFRONT>(a, grid, st);
            default: return launch_coder<32, 32, SRC_WAVEFRONT>(a, grid, st);
        }
    }
    int grid = grid_for(a.n_blocks, 4 * (32 / size), 8);
    switch (size) {
        case 4: return launch_coder<4, 4, SRC == SRC_WAVEFRONT ? SRC_PLANE : SRC>(a, grid, st);
        case 8: return launch_coder<8, 8, SRC == SRC_WAVEFRONT ? SRC_PLANE : SRC>(a, grid, st);
        case 16: return launch_coder<16, 16, SRC == SRC_WAVEFRONT ? SRC_PLANE : SRC>(a, grid, st);
        default: return launch_coder<32, 32, SRC == SRC_WAVEFRONT ? SRC_PLANE : SRC>(a, grid, st);
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// K7 as two launches (8-bit content): search_plane_kernel decides the modes, coder_kernel codes the
// winners (one block per warp and the tensor-core winner pipeline at N = 16 / 32).
// 2 (default) = search kernel + winner kernel, 1 = everything in the single coder kernel (A/B profiling);
// 3 / 4 = as 2 with the line-synchronous / the strip search kernel forced (2 picks per call, see
// launch_search_cost), 5 = as 2 with the fraction-major search kernel forced at N = 16 / 32 (nh_search3.cuh),
// 6 = as 2 with the tensor-core SATD search kernel forced (nh_search4.cuh; SATD, N >= 8);
// nh_set_search_impl() or NH_SEARCH_IMPL=1|2|3|4|5|6.
static thread_local int g_search_impl = 0;   // per calling thread: no shared mutable state between callers
static int search_impl() {
    if (g_search_impl == 0) {
        const char* e = getenv("NH_SEARCH_IMPL");
        g_search_impl = (e && e[0] >= '1' && e[0] <= '6') ? e[0] - '0' : 2;
    }
    return g_search_impl;
}
static int split_impl() { return search_impl() >= 2; }
static bool frac_default() {
    static const bool on = [] { const char* e = getenv("NH_SEARCH_FRAC"); return !(e && e[0] == '0'); }();
    return on;
}

static bool quad_default() {
    static const bool on = [] { const char* e = getenv("NH_SEARCH_QUAD"); return !(e && e[0] == '0'); }();
    return on;
}

template <int N, int COST>
static int launch_search_cost(const CoderArgs& a, cudaStream_t st) {
    SearchArgs s{a.src, a.H, a.W, a.pitch, a.cost_kind, a.n_blocks, a.out.modes, a.out.costs, a.blocks_per_frame,
                 a.frame_stride};
    if constexpr (N >= 8 && COST == NH_COST_SATD) {
        // SATD with the 4x4 Hadamard transforms on the tensor cores (nh_search4.cuh); needs 16-byte pixel loads
        using Q = QuadCfg<N>;
        const bool can = (a.pitch % 8) == 0 && (a.frame_stride % 8) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 15) == 0;
        const int impl = search_impl();
        if (can && (impl == 6 || (impl == 2 && quad_default()))) {
            int rc = ensure_dynamic_smem(search_quad_kernel<N>, Q::SMEM_BYTES, "search_quad_kernel");
            if (rc != NH_OK) return rc;
            const int grid = grid_for(a.n_blocks, (int64_t)Q::WARPS * Q::T, Q::PER_SM);
            search_quad_kernel<N><<<grid, Q::WARPS * 32, Q::SMEM_BYTES, st>>>(s);
            NH_CHECK_LAUNCH("search_quad_kernel");
            return NH_OK;
        }
    }
    if constexpr (N >= 16) {
        // The fraction-major kernel (nh_search3.cuh): every scan line is a window of a reference array filtered once
        // per fraction.  Measured against the line-synchronous kernel (32 4K frames, search + winners): N = 32 SAD
        // +6 %, SATD +17 %, one frame +11 %; at N = 16 it loses on SAD (four blocks per warp: 17 KB of arrays per warp,
        // 12 warps per SM) and wins 7 % on SATD, so it is the default at N = 32 and for SATD at N = 16 (NH_SEARCH_FRAC=0 turns that off).
        using F = FracCfg<N>;
        const int impl = search_impl();
        if (impl == 5 || (impl == 2 && (N == 32 || COST == NH_COST_SATD) && frac_default())) {
            int rc = ensure_dynamic_smem(search_frac_kernel<N, COST>, F::SMEM_BYTES, "search_frac_kernel");
            if (rc != NH_OK) return rc;
            const int grid = grid_for(a.n_blocks, (int64_t)F::WARPS * F::T, F::PER_SM);
            search_frac_kernel<N, COST><<<grid, F::WARPS * 32, F::SMEM_BYTES, st>>>(s);
            NH_CHECK_LAUNCH("search_frac_kernel");
            return NH_OK;
        }
    }
    if constexpr (N >= 8) {
        // The line-synchronous kernel (nh_search2.cuh) needs 16-byte pixel loads.  Measured (profiles/r2_search_*):
        // it wins on SAD once every warp gets a few of its (larger) tiles -- 32 4K frames: N = 8 / 16 / 32
        // 103 -> 117, 77 -> 84, 98 -> 116 Gpix/s for search + winners -- and on SATD at N = 32; with SATD at
        // N = 8 / 16 its 128 registers cost more occupancy than the uniform scan lines save.
        using L = LineCfg<N>;
        const int per_sm = (N > 8 || COST == NH_COST_SAD) ? L::PER_SM : 4;
        const bool can = (a.pitch % 8) == 0 && (a.frame_stride % 8) == 0 && (reinterpret_cast<uintptr_t>(a.src) & 15) == 0;
        const int64_t n_tiles = (a.n_blocks + L::T - 1) / L::T;
        const bool wins = (COST == NH_COST_SAD || N == 32) && n_tiles >= (int64_t)3 * sm_count() * per_sm * L::WARPS;
        const int impl = search_impl();
        if (can && (impl == 3 || (impl != 4 && wins))) {
            int rc = ensure_dynamic_smem(search_lines_kernel<N, COST>, L::SMEM_BYTES, "search_lines_kernel");
            if (rc != NH_OK) return rc;
            const int grid = grid_for(a.n_blocks, (int64_t)L::WARPS * L::T, per_sm);
            search_lines_kernel<N, COST><<<grid, L::WARPS * 32, L::SMEM_BYTES, st>>>(s);
            NH_CHECK_LAUNCH("search_lines_kernel");
            return NH_OK;
        }
    }
    using C = SearchCfg<N>;
    int rc = ensure_dynamic_smem(search_plane_kernel<N, COST>, C::SMEM_BYTES, "search_plane_kernel");
    if (rc != NH_OK) return rc;
    const int per_sm = COST == NH_COST_SAD ? 6 : 5;  // the kernel's __launch_bounds__; 6 x SMEM_BYTES <= 170 KB for every N
    const int grid = grid_for(a.n_blocks, (int64_t)C::WARPS * C::T, per_sm);
    search_plane_kernel<N, COST><<<grid, C::WARPS * 32, C::SMEM_BYTES, st>>>(s);
    NH_CHECK_LAUNCH("search_plane_kernel");
    return NH_OK;
}
// N = 4 with the winner stage fused into the search kernel (nh_search.cuh, CODE = true).  *handed_back = the
// stream's counter of tiles left to the exact coder kernel.
template <int COST>
static int launch_search_code4(const CoderArgs& a, cudaStream_t st, unsigned int** handed_back) {
    using C = SearchCfg<4>;
    int rc = ensure_dynamic_smem(search_plane_kernel<4, COST, true>, C::SMEM_BYTES, "search_plane_kernel<code>");
    if (rc != NH_OK) return rc;
    unsigned int* counter = nullptr;
    rc = acquire_tile_counter(st, &counter);
    if (rc != NH_OK) return rc;
    *handed_back = counter + 2;
    SearchArgs s{a.src, a.H, a.W, a.pitch, a.cost_kind, a.n_blocks, a.out.modes, a.out.costs, a.blocks_per_frame,
                 a.frame_stride, a.fq, a.maxv, a.out.pred, a.out.coeff, a.out.levels, a.out.recon_plane, counter + 2};
    if constexpr (COST == NH_COST_SATD) {
        // SATD with the Hadamard transforms on the tensor cores (nh_search4.cuh, search_quad4_kernel)
        const int impl = search_impl();
        if (impl == 6 || (impl == 2 && quad_default())) {
            rc = ensure_dynamic_smem(search_quad4_kernel<true>, Quad4Cfg::SMEM_BYTES, "search_quad4_kernel");
            if (rc != NH_OK) return rc;
            const int gridq = grid_for(a.n_blocks, (int64_t)C::WARPS * C::T, Quad4Cfg::PER_SM);
            search_quad4_kernel<true><<<gridq, C::WARPS * 32, Quad4Cfg::SMEM_BYTES, st>>>(s);
            NH_CHECK_LAUNCH("search_quad4_kernel");
            return NH_OK;
        }
    }
    const int grid = grid_for(a.n_blocks, (int64_t)C::WARPS * C::T, 4);
    search_plane_kernel<4, COST, true><<<grid, C::WARPS * 32, C::SMEM_BYTES, st>>>(s);
    NH_CHECK_LAUNCH("search_plane_kernel<code>");
    return NH_OK;
}

template <int N>
static int launch_search(const CoderArgs& a, cudaStream_t st) {
    return a.cost_kind == NH_COST_SAD ? launch_search_cost<N, NH_COST_SAD>(a, st) : launch_search_cost<N, NH_COST_SATD>(a, st);
}

static int dispatch_search_then_code(CoderArgs a, int size, cudaStream_t st) {
    int rc;
    // N = 4: the lane that searched a block codes it (NH_SEARCH4_FUSED=0 keeps the two-kernel form)
    static const bool fused4 = [] { const char* e = getenv("NH_SEARCH4_FUSED"); return !(e && e[0] == '0'); }();
    if (size == 4 && fused4 && a.use_dst && a.out.recon_plane && (reinterpret_cast<uintptr_t>(a.out.recon_plane) & 7) == 0) {
        rc = a.cost_kind == NH_COST_SAD ? launch_search_code4<NH_COST_SAD>(a, st, &a.handed_back)
                                        : launch_search_code4<NH_COST_SATD>(a, st, &a.handed_back);
        if (rc != NH_OK) return rc;
        a.modes_in = a.out.modes;
        a.only_undecided = 1;   // the exact coder only touches the tiles marked 0xFF (samples outside [0, 255])
        rc = dispatch_coder<SRC_PLANE>(a, size, st);
        tile_counter_launched(st);
        return rc;
    }
    switch (size) {
        case 4: rc = launch_search<4>(a, st); break;
        case 8: rc = launch_search<8>(a, st); break;
        case 16: rc = launch_search<16>(a, st); break;
        default: rc = launch_search<32>(a, st); break;
    }
    if (rc != NH_OK) return rc;
    a.modes_in = a.out.modes;
    if (size == 8 && (a.pitch % 8) == 0 && (a.frame_stride % 8) == 0 && aligned16(a.src) && aligned16(a.out.recon_plane)) {
        // winners on the tensor cores (nh_coder8.cuh); it hands the tiles it cannot take (undecided
        // blocks) back by marking them 0xFF, and the exact coder below only touches marked blocks
        rc = coder8_plane_mma(a.src, a.n_frames, a.frame_stride, a.H, a.W, a.pitch, a.out.modes, a.out.pred, a.out.coeff,
                              a.out.levels, a.out.recon_plane, a.qp, a.maxv, st, &a.handed_back);
        if (rc != NH_OK) return rc;
        a.only_undecided = 1;
    }
    if (size == 16) {
        // two blocks per warp on the tensor cores (nh_winner16.cuh); undecided blocks (0xFF) go to the exact coder
        static const bool pair = [] { const char* e = getenv("NH_WINNER16_PAIR"); return !(e && e[0] == '0'); }();
        if (pair && (a.pitch % 8) == 0 && (a.frame_stride % 8) == 0 && aligned16(a.src) && aligned16(a.out.recon_plane) &&
            a.maxv <= 1023) {
            rc = ensure_dynamic_smem(winner16_pair_kernel, Winner16Cfg::SMEM_BYTES, "winner16_pair_kernel");
            if (rc != NH_OK) return rc;
            unsigned int* counter = nullptr;
            rc = acquire_tile_counter(st, &counter);
            if (rc != NH_OK) return rc;
            a.handed_back = counter + 2;
            winner16_pair_kernel<<<grid_for((a.n_blocks + 1) / 2, Winner16Cfg::WARPS, 4), Winner16Cfg::WARPS * 32,
                                   Winner16Cfg::SMEM_BYTES, st>>>(a);
            NH_CHECK_LAUNCH("winner16_pair_kernel");
            a.only_undecided = 1;
            rc = launch_coder<16, 32, SRC_PLANE>(a, grid_for(a.n_blocks, 4, NH_WINNER_OCC), st);
            tile_counter_launched(st);
            return rc;
        }
        return launch_coder<16, 32, SRC_PLANE>(a, grid_for(a.n_blocks, 4, NH_WINNER_OCC), st);
    }
    rc = dispatch_coder<SRC_PLANE>(a, size, st);
    if (a.only_undecided) tile_counter_launched(st);   // last launch that reads the stream's counter slot
    return rc;
}


}  // namespace nh

using namespace nh;

NH_API int nh_gather_refs(const int16_t* plane, int height, int width, int pitch, int size, int n_top,
                          int n_left, int16_t* top, int16_t* left, int16_t* corner, void* stream) {
    if (log2_size(size) < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (!plane || !top || !left || !corner || height < 0 || width < 0 || pitch < width ||
        n_top < 1 || n_top > 2 * size || n_left < 1 || n_left > 2 * size) {
        set_error("nh_gather_refs: bad argument (null pointer, pitch < width, or n_top/n_left outside 1..2N)");
        return NH_E_ARG;
    }
    int64_t B = (int64_t)(height / size) * (width / size);
    if (B == 0) return NH_OK;
    int grid = grid_for(B * (2 * size + 1), 256, 8);
    gather_refs_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        plane, height, width, pitch, size, n_top, n_left, top, left, corner);
    NH_CHECK_LAUNCH("gather_refs_kernel");
    return NH_OK;
}

static int reblock(const int16_t* in, int height, int width, int pitch, int size, int16_t* out,
                   bool to_blocks, void* stream) {
    if (log2_size(size) < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (!in || !out || height < 0 || width < 0 || pitch < width) {
        set_error("plane/block conversion: bad argument");
        return NH_E_ARG;
    }
    int64_t n = (int64_t)(height / size) * (width / size) * size * size;
    if (n == 0) return NH_OK;
    int grid = grid_for(n, 256, 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (to_blocks) reblock_kernel<true><<<grid, 256, 0, st>>>(in, height, width, pitch, size, out);
    else reblock_kernel<false><<<grid, 256, 0, st>>>(in, height, width, pitch, size, out);
    NH_CHECK_LAUNCH("reblock_kernel");
    return NH_OK;
}

NH_API int nh_plane_to_blocks(const int16_t* plane, int height, int width, int pitch, int size,
                              int16_t* blocks, void* stream) {
    return reblock(plane, height, width, pitch, size, blocks, true, stream);
}

NH_API int nh_blocks_to_plane(const int16_t* blocks, int height, int width, int pitch, int size,
                              int16_t* plane, void* stream) {
    return reblock(blocks, height, width, pitch, size, plane, false, stream);
}

NH_API int nh_fused_pipeline_modes(const int16_t* orig, const int16_t* top, const int16_t* left,
                                   const int16_t* top_left, const uint8_t* modes, int mode,
                                   int64_t n_blocks, int size, int qp, int is_intra, int use_dst,
                                   int bit_depth, int16_t* pred, int32_t* coeff, int32_t* levels,
                                   int16_t* recon, void* stream) {
    int l2 = log2_size(size);
    if (l2 < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (n_blocks == 0) return NH_OK;
    if (!orig || !top || !left || !top_left || n_blocks < 0) {
        set_error("nh_fused_pipeline_modes: null input or negative block count");
        return NH_E_ARG;
    }
    if (!modes && (mode < 0 || mode > 34)) {
        set_error("nh_fused_pipeline_modes: mode %d out of range 0..34", mode);
        return NH_E_ARG;
    }
    if (bit_depth < 1 || bit_depth > 15) {
        set_error("nh_fused_pipeline_modes: bit_depth %d out of range", bit_depth);
        return NH_E_ARG;
    }
    if (!aligned16(pred) || !aligned16(coeff) || !aligned16(levels) || !aligned16(recon)) {
        set_error("nh_fused_pipeline_modes: output tensors must be 16-byte aligned");
        return NH_E_ARG;
    }
    if (n_blocks == 0) return NH_OK;
    CoderArgs a{};
    a.orig = orig; a.top = top; a.left = left; a.top_left = top_left; a.modes_in = modes; a.mode = mode;
    a.n_blocks = n_blocks;
    a.n_frames = 1;
    a.blocks_per_frame = n_blocks;
    a.qp = make_quant_params(qp, l2, is_intra);
    a.fq = make_fast_quant(a.qp);
    a.maxv = (1 << bit_depth) - 1;
    a.use_dst = use_dst;
    a.out = CoderOut{nullptr, nullptr, pred, coeff, levels, recon, nullptr, 0};
    return dispatch_coder<SRC_ARRAYS>(a, size, reinterpret_cast<cudaStream_t>(stream));
}

NH_API int nh_set_search_impl(int impl) {
    if (impl < 1 || impl > 6) {
        set_error("nh_set_search_impl: impl must be 1 (single coder kernel), 2 (search + winner kernels), 3 / 4 / 5 / 6 (as 2, "
                  "line-synchronous / strip / fraction-major / tensor-core SATD search kernel forced), got %d", impl);
        return NH_E_ARG;
    }
    g_search_impl = impl;
    return NH_OK;
}

NH_API int nh_set_wave_impl(int warps, int build) {
    if (!(warps == 0 || warps == 1 || warps == 2 || warps == 4 || warps == 8 || warps == 12) || build < 0 || build > 3) {
        set_error("nh_set_wave_impl: warps must be 0 (pick per call), 1, 2, 4, 8 or 12 and build 0 (pick per call), 1 (latency), "
                  "2 (throughput) or 3 (throughput, highest occupancy), got %d, %d", warps, build);
        return NH_E_ARG;
    }
    g_wave_impl.warps = warps;
    g_wave_impl.build = build;
    g_wave_impl.init = true;
    return NH_OK;
}

NH_API int64_t nh_encode_frames_scratch_bytes(int n_frames, int height, int width, int size) {
    if (log2_size(size) < 0 || height < 0 || width < 0 || n_frames < 0) return 0;
    return 256 + (int64_t)n_frames * (height / size) * width * 2;  // ticket counter + exchange rows of every frame
}

NH_API int64_t nh_encode_frame_scratch_bytes(int height, int width, int size) {
    return nh_encode_frames_scratch_bytes(1, height, width, size);
}

NH_API int nh_encode_frames(const int16_t* src, int n_frames, int64_t frame_stride, int height, int width, int pitch,
                            int size, int cost_kind, int qp, int recon_neighbours, int bit_depth, uint8_t* modes,
                            int32_t* costs, int16_t* pred, int32_t* coeff, int32_t* levels, int16_t* recon_planes,
                            int64_t* stats, void* scratch, int64_t scratch_bytes, void* stream) {
    int l2 = log2_size(size);
    if (l2 < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (!src || n_frames < 0 || height < 0 || width < 0 || pitch < width ||
        (n_frames > 1 && frame_stride < (int64_t)height * pitch) ||
        (cost_kind != NH_COST_SAD && cost_kind != NH_COST_SATD) || bit_depth < 1 || bit_depth > 15) {
        set_error("nh_encode_frames: bad argument");
        return NH_E_ARG;
    }
    if ((recon_neighbours || stats) && !recon_planes) {
        set_error("nh_encode_frames: recon_planes is required when recon_neighbours != 0 or stats are requested");
        return NH_E_ARG;
    }
    if (!aligned16(pred) || !aligned16(coeff) || !aligned16(levels)) {
        set_error("nh_encode_frames: output tensors must be 16-byte aligned");
        return NH_E_ARG;
    }
    if (n_frames == 0) return NH_OK;
    if (n_frames == 1) frame_stride = (int64_t)height * pitch;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (recon_planes) {  // Plane.zeros semantics (frame.py:41-43): uncovered samples stay 0
        cudaError_t e = frame_stride == (int64_t)height * pitch
                            ? cudaMemsetAsync(recon_planes, 0, (size_t)n_frames * height * pitch * sizeof(int16_t), st)
                            : cudaMemset2DAsync(recon_planes, (size_t)frame_stride * sizeof(int16_t), 0,
                                                (size_t)height * pitch * sizeof(int16_t), (size_t)n_frames, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(recon_planes)");
    }
    if (stats) {
        cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)n_frames * 4 * sizeof(int64_t), st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(stats)");
    }
    const int bw = width / size, bh = height / size;
    const bool planes_vec = (pitch % 8) == 0 && (frame_stride % 8) == 0 && aligned16(src) && aligned16(recon_planes);
    auto launch_stats = [&]() -> int {
        if (!stats) return NH_OK;
        const int64_t bpf = (int64_t)bw * bh;
        dim3 grid((unsigned)(sm_count() * 4 / (n_frames < sm_count() * 4 ? n_frames : sm_count() * 4) + 1), (unsigned)n_frames);
        frame_stats_kernel<<<grid, 256, 0, st>>>(src, recon_planes, frame_stride, height, width, pitch, planes_vec ? 1 : 0,
                                                 bpf ? costs : nullptr, bpf ? levels : nullptr, bpf, size * size, stats);
        NH_CHECK_LAUNCH("frame_stats_kernel");
        return NH_OK;
    };
    if (bw == 0 || bh == 0) return launch_stats();
    CoderArgs a{};
    a.src = src; a.H = height; a.W = width; a.pitch = pitch; a.cost_kind = cost_kind;
    a.n_frames = n_frames;
    a.frame_stride = frame_stride;
    a.blocks_per_frame = (int64_t)bw * bh;
    a.n_blocks = a.blocks_per_frame * n_frames;
    a.qp = make_quant_params(qp, l2, 1);
    a.fq = make_fast_quant(a.qp);
    a.maxv = (1 << bit_depth) - 1;
    a.use_dst = size == 4;  // docs/frames_and_panes.md:328-329
    a.out = CoderOut{modes, costs, pred, coeff, levels, nullptr, recon_planes, pitch};
    a.vec_ok = planes_vec;
    int rc;
    if (!recon_neighbours) {
        // the search kernel reads the plane in 8-byte pieces and needs the modes tensor as its output
        const bool split_ok = split_impl() && bit_depth <= 8 && modes && (pitch % 4) == 0 && (frame_stride % 4) == 0 &&
                              (reinterpret_cast<uintptr_t>(src) & 7) == 0;
        rc = split_ok ? dispatch_search_then_code(a, size, st) : dispatch_coder<SRC_PLANE>(a, size, st);
        return rc != NH_OK ? rc : launch_stats();
    }
    const int64_t need = nh_encode_frames_scratch_bytes(n_frames, height, width, size);
    if (!scratch || scratch_bytes < need) {
        set_error("nh_encode_frames: scratch of %lld bytes required, got %lld", (long long)need,
                  (long long)scratch_bytes);
        return NH_E_NOMEM;
    }
    if ((reinterpret_cast<uintptr_t>(scratch) & 3) != 0) {
        set_error("nh_encode_frames: scratch must be 4-byte aligned");
        return NH_E_ARG;
    }
    {
        static const int sleep_ns = [] {  // NH_WAVE_SLEEP_NS overrides the default back-off
            const char* e = getenv("NH_WAVE_SLEEP_NS");
            const int v = e ? atoi(e) : 0;  // measured: polling without back-off is as fast or faster (profiles/r1_notes.md)
            return v < 0 ? 0 : v;
        }();
        a.poll_sleep_ns = (unsigned)sleep_ns;
    }
    a.ticket = reinterpret_cast<int*>(scratch);
    a.bottom = reinterpret_cast<int16_t*>(reinterpret_cast<unsigned char*>(scratch) + 256);
    const int64_t rows = (int64_t)bh * n_frames;
    init_bottom_kernel<<<grid_for(rows * width, 256, 4), 256, 0, st>>>(a.bottom, rows, width, bw * size, a.ticket);
    NH_CHECK_LAUNCH("init_bottom_kernel");
    rc = dispatch_coder<SRC_WAVEFRONT>(a, size, st);
    return rc != NH_OK ? rc : launch_stats();
}

NH_API int nh_encode_frame(const int16_t* src, int height, int width, int pitch, int size,
                           int cost_kind, int qp, int recon_neighbours, int bit_depth, uint8_t* modes,
                           int32_t* costs, int16_t* pred, int32_t* coeff, int32_t* levels,
                           int16_t* recon_plane, void* scratch, int64_t scratch_bytes, void* stream) {
    return nh_encode_frames(src, 1, (int64_t)height * pitch, height, width, pitch, size, cost_kind, qp, recon_neighbours,
                            bit_depth, modes, costs, pred, coeff, levels, recon_plane, nullptr, scratch, scratch_bytes,
                            stream);
}
