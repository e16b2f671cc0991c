// nh_mma.cuh -- warp-level tensor-core building blocks shared by the fused pipeline kernels
// (nh_fused_mma.cuh), the single-stage transform kernels (nh_ops.cu) and the winner pipeline of the
// frame coders (nh_frame.cu): ldmatrix / stmatrix / HMMA wrappers, the magic-number rounding helpers,
// the compile-time per-lane constant tables of the 16- and 32-point transform matrices and
// mma_block_chain, the register-chained pipeline of one block.  See nh_fused_mma.cuh for how the
// fragments chain and why the arithmetic is exact.
#pragma once
#include <cuda_fp16.h>

#include "nh_common.cuh"

namespace nh {


constexpr int kMmaWarps = 4;
constexpr float kMagicF = 12582912.0f;  // 1.5 * 2^23
constexpr int kMagicI = 0x4B400000;     // its bit pattern: float(kMagicF + k) has bits kMagicI + k
// Operand bias: integers k in [-512, 511] travel as the f16 number k + 1536, whose bit pattern is
// 0x6600 + k.  FFMA.RM against (magic + 0x6600) leaves exactly those 16 bits in the low half of the
// f32 result, so one PRMT packs two operands; the constant 1536 * (row or column sum of T) is taken
// out again through the next pass's accumulator start value.
constexpr int kOperandBias = 1536;
constexpr int kOperandBits = 0x6600;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, const uint32_t (&r)[4]) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// D = A(16x16, row) * B(16x8, col) + C, f16 operands, f32 accumulate.  C is a separate operand so
// that a pass can start from constant registers without copying them into the accumulators first.
__device__ __forceinline__ void hmma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1,
                                          float c0, float c1, float c2, float c3) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 bits_h2(uint32_t w) { return *reinterpret_cast<__half2*>(&w); }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { return h2_bits(__floats2half2_rn(lo, hi)); }

// floor(acc / 2^SH) of an accumulator that already holds the rounding offset, as kMagicF + k
template <int SH>
__device__ __forceinline__ float floor_shift_magic(float acc) {
    return __fmaf_rd(acc, 1.0f / (float)(1 << SH), kMagicF);
}
// Pass boundary, plain form: f32 integer -> f16 pair (FFMA.RM, FADD, half an F2FP per value).
template <int SH>
__device__ __forceinline__ uint32_t round_pair_plain(float a0, float a1) {
    return pack_h2(floor_shift_magic<SH>(a0) - kMagicF, floor_shift_magic<SH>(a1) - kMagicF);
}
// Pass boundary, biased form for k in [-512, 511] (FFMA.RM + half a PRMT per value).
template <int SH>
__device__ __forceinline__ uint32_t round_pair_biased(float a0, float a1) {
    const float m = kMagicF + (float)kOperandBits;
    return __byte_perm(__float_as_uint(__fmaf_rd(a0, 1.0f / (float)(1 << SH), m)),
                       __float_as_uint(__fmaf_rd(a1, 1.0f / (float)(1 << SH), m)), 0x5410);
}

// ---- per-lane constants, generated at compile time ------------------------------------------------
// f16 bit pattern of a small integer (|v| <= 2048, exact)
constexpr uint32_t f16_bits_of_int(int v) {
    if (v == 0) return 0;
    const uint32_t s = v < 0 ? 0x8000u : 0u;
    const uint32_t m = (uint32_t)(v < 0 ? -v : v);
    int e = 0;
    while ((m >> e) > 1) ++e;
    const uint32_t frac = e <= 10 ? (m << (10 - e)) : (m >> (e - 10));
    return s | ((uint32_t)(e + 15) << 10) | (frac & 0x3ffu);
}
template <int N>
struct MmaConsts {
    static constexpr int MT = N / 16, NT = N / 8, KT = N / 16;
    // 128-bit vectors per lane:
    //   V_TA + mi*KT + ki : A = T fragment    {T[16mi+g][16ki+2t..+1], rows +8, cols +8, both}
    //   V_TB + ki*NT/2+np : B = T fragments   {b0, b1 of n-tile 2np, b0, b1 of n-tile 2np+1},
    //                       b0 = {T[16ki+2t][8ni+g], T[16ki+2t+1][8ni+g]}, b1 = rows +8
    //   accumulator start values (integers; converted to f32 when the CTA stages the table):
    //   V_F2: r - 1536 * sum_x T[v][x], v = 16mi + g + 8h, at word 2mi + h   (forward, second pass)
    //   V_I1: r - 1536 * sum_i T[i][y], y likewise                           (inverse, first pass)
    //   V_I2: r - 1536 * sum_v T[v][x], x = 8ni + 2t + p, at word 2ni + p    (inverse, second pass;
    //         plain r when the operand of that pass is not biased)
    static constexpr int V_TA = 0, V_TB = MT * KT, V_F2 = V_TB + KT * NT / 2, V_I1 = V_F2 + (2 * MT + 3) / 4,
                         V_I2 = V_I1 + (2 * MT + 3) / 4, V_END = V_I2 + (2 * NT + 3) / 4;
    int32_t w[V_END][32][4];
};
template <int N>
constexpr MmaConsts<N> make_mma_consts(bool bias_tmp2) {
    using C = MmaConsts<N>;
    C c{};
    const int rnd = 1 << (Log2<N>::v + 4);
    auto pair = [](int r0, int c0, int r1, int c1) -> int32_t {
        return (int32_t)(f16_bits_of_int(dct<N>(r0, c0)) | (f16_bits_of_int(dct<N>(r1, c1)) << 16));
    };
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        for (int mi = 0; mi < C::MT; ++mi)
            for (int ki = 0; ki < C::KT; ++ki) {
                const int i0 = 16 * mi + g, k0 = 16 * ki + 2 * t;
                int32_t(&v)[4] = c.w[C::V_TA + mi * C::KT + ki][lane];
                v[0] = pair(i0, k0, i0, k0 + 1);
                v[1] = pair(i0 + 8, k0, i0 + 8, k0 + 1);
                v[2] = pair(i0, k0 + 8, i0, k0 + 9);
                v[3] = pair(i0 + 8, k0 + 8, i0 + 8, k0 + 9);
            }
        for (int ki = 0; ki < C::KT; ++ki)
            for (int ni = 0; ni < C::NT; ++ni) {
                const int k0 = 16 * ki + 2 * t, x0 = 8 * ni + g;
                int32_t(&v)[4] = c.w[C::V_TB + ki * C::NT / 2 + ni / 2][lane];
                v[2 * (ni & 1)] = pair(k0, x0, k0 + 1, x0);
                v[2 * (ni & 1) + 1] = pair(k0 + 8, x0, k0 + 9, x0);
            }
        for (int mi = 0; mi < C::MT; ++mi)
            for (int hh = 0; hh < 2; ++hh) {
                const int v = 16 * mi + g + 8 * hh;
                int rs = 0, cs = 0;
                for (int k = 0; k < N; ++k) {
                    rs += dct<N>(v, k);
                    cs += dct<N>(k, v);
                }
                c.w[C::V_F2 + (2 * mi + hh) / 4][lane][(2 * mi + hh) % 4] = rnd - kOperandBias * rs;
                c.w[C::V_I1 + (2 * mi + hh) / 4][lane][(2 * mi + hh) % 4] = rnd - kOperandBias * cs;
            }
        for (int ni = 0; ni < C::NT; ++ni)
            for (int p = 0; p < 2; ++p) {
                const int x = 8 * ni + 2 * t + p;
                int cs = 0;
                for (int k = 0; k < N; ++k) cs += dct<N>(k, x);
                c.w[C::V_I2 + (2 * ni + p) / 4][lane][(2 * ni + p) % 4] = bias_tmp2 ? rnd - kOperandBias * cs : rnd;
            }
    }
    return c;
}
// |tmp2| <= 328 at N = 32 fits the biased operand form, 661 at N = 16 does not
template <int N> struct MmaBiasTmp2 { static constexpr bool v = N == 32; };
__device__ constexpr MmaConsts<16> kMmaConsts16 = make_mma_consts<16>(MmaBiasTmp2<16>::v);
__device__ constexpr MmaConsts<32> kMmaConsts32 = make_mma_consts<32>(MmaBiasTmp2<32>::v);
template <int N>
__device__ __forceinline__ const int32_t* mma_consts_words() {
    if constexpr (N == 16) return &kMmaConsts16.w[0][0][0];
    else return &kMmaConsts32.w[0][0][0];
}


// Stage the per-lane constant table of the N-point transform into shared memory (vector v of lane l
// at ctab[v][l]); the start values are stored as integers and converted here.  Callers
// __syncthreads() afterwards.
template <int N, int THREADS>
__device__ __forceinline__ void stage_mma_consts(uint4* ctab) {
    using C = MmaConsts<N>;
    for (int i = threadIdx.x; i < C::V_END * 32; i += THREADS) {
        const int4 v = reinterpret_cast<const int4*>(mma_consts_words<N>())[i];
        uint4 o = make_uint4((uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w);
        if (i >= C::V_F2 * 32)
            o = make_uint4(__float_as_uint((float)v.x), __float_as_uint((float)v.y), __float_as_uint((float)v.z),
                           __float_as_uint((float)v.w));
        ctab[i] = o;
    }
}
// this lane's constant vector v (volatile so that the load stays where it is used)
__device__ __forceinline__ uint4 ld_const_vec(uint32_t ctab_lane, int v) {
    uint4 o;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(ctab_lane + 512u * (uint32_t)v));
    return o;
}

// acc(m, n) = init(m) + A_const(m, k) * B(k, n) where B[k][n] = P[n][k] and P is the previous pass's
// result held as packed C fragments `h` (C -> B: the product comes out transposed w.r.t. using P as
// A).  `i4` holds this lane's start values (word 2mi + half), `afrag(mi, ki)` fetches an A fragment.
template <int MT, int NT, int KT, class AFrag>
__device__ __forceinline__ void mma_const_a(float (&acc)[MT][NT][4], const uint32_t (&h)[MT][NT][2],
                                            const uint4& i4, AFrag afrag) {
    const float init[4] = {__uint_as_float(i4.x), __uint_as_float(i4.y), __uint_as_float(i4.z), __uint_as_float(i4.w)};
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
        for (int ki = 0; ki < KT; ++ki) {
            const uint4 af = afrag(mi, ki);
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                const uint32_t b0 = h[ni >> 1][2 * ki][ni & 1], b1 = h[ni >> 1][2 * ki + 1][ni & 1];
                if (ki == 0)
                    hmma16816(acc[mi][ni], af, b0, b1, init[2 * mi], init[2 * mi], init[2 * mi + 1], init[2 * mi + 1]);
                else
                    hmma16816(acc[mi][ni], af, b0, b1, acc[mi][ni][0], acc[mi][ni][1], acc[mi][ni][2], acc[mi][ni][3]);
            }
        }
    }
}

// The register-chained pipeline of ONE N x N block (N = 16 / 32) for a whole warp: residual from
// the original / prediction tiles in shared memory (`so` / `sp` = this lane's ldmatrix row address in
// each, row pitch PITCH bytes), the four transform passes, quant / dequant in between, coefficients
// and levels stored straight from the fragments (`cp` / `lp` already point at this lane's element
// (i = 2 ft, v = fg)), reconstruction + clip written back over the original tile.  `cv(v)` fetches
// vector v of the lane's constants (MmaConsts<N>).  Valid for 8-bit samples only (see above).
template <int N, class CV>
__device__ __forceinline__ void mma_block_chain(uint32_t so, uint32_t sp, CV cv, bool want_c, int32_t* cp,
                                                bool want_l, int32_t* lp, const FastQuant& fq, int dq_rnd_b,
                                                uint32_t clip_lo2, uint32_t clip_hi2) {
    using C = MmaConsts<N>;
    constexpr int MT = N / 16, NT = N / 8, KT = N / 16;
    constexpr int SH = Log2<N>::v + 5;
    constexpr int PITCH = N * 2 + 16;
    constexpr bool kBiasTmp2 = MmaBiasTmp2<N>::v;
    const float rnd = (float)(1 << (SH - 1));
    float acc[MT][NT][4];
    uint32_t h[MT][NT][2];
    // ---- forward, first pass: temp = (T X + r) >> s,  X = orig - pred as B fragments
    {
        uint32_t xb[KT][NT][2];
#pragma unroll
        for (int ki = 0; ki < KT; ++ki)
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t ro[4], rp[4];
                const uint32_t off = 16 * ki * PITCH + 32 * np;
                ldsm_x4_t(ro, so + off);
                ldsm_x4_t(rp, sp + off);
#pragma unroll
                for (int j = 0; j < 4; ++j)  // (1024 + o) - (1024 + p), exact in f16
                    xb[ki][2 * np + (j >> 1)][j & 1] = h2_bits(
                        __hsub2(bits_h2(ro[j] | 0x64006400u), bits_h2(rp[j] | 0x64006400u)));
            }
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int ki = 0; ki < KT; ++ki) {
                const uint4 af = cv(C::V_TA + mi * KT + ki);
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) {
                    if (ki == 0)
                        hmma16816(acc[mi][ni], af, xb[ki][ni][0], xb[ki][ni][1], rnd, rnd, rnd, rnd);
                    else
                        hmma16816(acc[mi][ni], af, xb[ki][ni][0], xb[ki][ni][1], acc[mi][ni][0],
                                  acc[mi][ni][1], acc[mi][ni][2], acc[mi][ni][3]);
                }
            }
    }
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {  // |temp| <= 511: biased operand
            h[mi][ni][0] = round_pair_biased<SH>(acc[mi][ni][0], acc[mi][ni][1]);
            h[mi][ni][1] = round_pair_biased<SH>(acc[mi][ni][2], acc[mi][ni][3]);
        }
    // ---- forward, second pass (transposed): coeff^T(m = v, n = i) = (T temp^T + r) >> s
    mma_const_a<MT, NT, KT>(acc, h, cv(C::V_F2), [&](int mi, int ki) { return cv(C::V_TA + mi * KT + ki); });
    // ---- coefficients out, quant, levels out, dequant -> biased operand of the inverse
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            int dqb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = __float_as_int(floor_shift_magic<SH>(acc[mi][ni][e])) - kMagicI;
                const int off = (8 * ni + (e & 1)) * N + 16 * mi + 8 * (e >> 1);
                const int lv = quantize_fast(c, fq);
                if (want_c) __stcs(cp + off, c);
                if (want_l) __stcs(lp + off, lv);
                dqb[e] = (lv * fq.dq_mult + dq_rnd_b) >> fq.dq_shift;  // 0x6600 + dq, |dq| <= 360
            }
            h[mi][ni][0] = __byte_perm((uint32_t)dqb[0], (uint32_t)dqb[1], 0x5410);
            h[mi][ni][1] = __byte_perm((uint32_t)dqb[2], (uint32_t)dqb[3], 0x5410);
        }
    // ---- inverse, first pass: tmp2(m = y, n = v) = (T^T dq + r) >> s.  A = T^T comes from
    // the B = T fragments: a0 = b0 of n-tile 2mi, a1 = b0 of 2mi+1, a2 / a3 = their b1
    mma_const_a<MT, NT, KT>(acc, h, cv(C::V_I1), [&](int mi, int ki) {
        const uint4 t = cv(C::V_TB + ki * NT / 2 + mi);
        return make_uint4(t.x, t.z, t.y, t.w);
    });
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            if constexpr (kBiasTmp2) {
                h[mi][ni][0] = round_pair_biased<SH>(acc[mi][ni][0], acc[mi][ni][1]);
                h[mi][ni][1] = round_pair_biased<SH>(acc[mi][ni][2], acc[mi][ni][3]);
            } else {
                h[mi][ni][0] = round_pair_plain<SH>(acc[mi][ni][0], acc[mi][ni][1]);
                h[mi][ni][1] = round_pair_plain<SH>(acc[mi][ni][2], acc[mi][ni][3]);
            }
        }
    // ---- inverse, second pass: res(m = y, n = x) = (tmp2 T + r) >> s,  A = tmp2 (C -> A)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
        const uint4 i4 = cv(C::V_I2 + np);  // start values of n-tiles 2np, 2np + 1
        const float i0 = __uint_as_float(i4.x), i1 = __uint_as_float(i4.y),
                    i2 = __uint_as_float(i4.z), i3 = __uint_as_float(i4.w);
#pragma unroll
        for (int ki = 0; ki < KT; ++ki) {
            const uint4 t = cv(C::V_TB + ki * NT / 2 + np);
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) {
                const uint4 af = make_uint4(h[mi][2 * ki][0], h[mi][2 * ki][1], h[mi][2 * ki + 1][0],
                                            h[mi][2 * ki + 1][1]);
                float(&d0)[4] = acc[mi][2 * np];
                float(&d1)[4] = acc[mi][2 * np + 1];
                if (ki == 0) {
                    hmma16816(d0, af, t.x, t.y, i0, i1, i0, i1);
                    hmma16816(d1, af, t.z, t.w, i2, i3, i2, i3);
                } else {
                    hmma16816(d0, af, t.x, t.y, d0[0], d0[1], d0[2], d0[3]);
                    hmma16816(d1, af, t.z, t.w, d1[0], d1[1], d1[2], d1[3]);
                }
            }
        }
    }
    // ---- reconstruct + clip (intra.py:70-78) on 16-bit pairs: FFMA.RM against magic + 2048
    // leaves res + 2048 (> 0, |res| <= 1214) in the low 16 bits; add the prediction pair,
    // clamp both halves to [2048, 2048 + max] and drop the bias.  The tile of original
    // pixels is reused for the result.
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            uint32_t rp[4], ro[4];
            const uint32_t off = 16 * mi * PITCH + 32 * np;
            ldsm_x4(rp, sp + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float(&c)[4] = acc[mi][2 * np + (j >> 1)];
                const uint32_t m0 = __float_as_uint(
                    __fmaf_rd(c[2 * (j & 1)], 1.0f / (float)(1 << SH), kMagicF + 2048.0f));
                const uint32_t m1 = __float_as_uint(
                    __fmaf_rd(c[2 * (j & 1) + 1], 1.0f / (float)(1 << SH), kMagicF + 2048.0f));
                const uint32_t s = __byte_perm(m0, m1, 0x5410) + rp[j];
                ro[j] = __vminu2(__vmaxu2(s, clip_lo2), clip_hi2) - clip_lo2;
            }
            stsm_x4(so + off, ro);
        }
}

}  // namespace nh
