// nh_mma.cuh -- warp-level tensor-core building blocks shared by the fused pipeline kernels
// (nh_fused_mma.cuh) and the single-stage transform kernels (nh_ops.cu): ldmatrix / stmatrix / HMMA
// wrappers, the magic-number rounding helpers and the compile-time per-lane constant tables of the
// 16- and 32-point transform matrices.  See nh_fused_mma.cuh for how the fragments chain.
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

}  // namespace nh
