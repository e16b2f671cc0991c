// nh_math.cuh -- integer arithmetic of the block-coding path, written once and
// used by every kernel.  Everything here is __host__ __device__ so the CPU test
// harness (tests/host_shim.cpp) can run exactly the same code against the
// oracle before a GPU is involved.
//
// Reference semantics (paths relative to the reference checkout):
//   transforms   nano_hevc/transform.py:154-238  (same shift log2N+5 in both passes, Q1)
//   quant        nano_hevc/quant.py:41-79        (shift 14+per+log2N, int64, Q2)
//   dequant      nano_hevc/quant.py:82-123       (arithmetic shift on signed base)
//   predictors   nano_hevc/intra.py:46-62, 81-113, 116-207 (k+1 projection, Q3)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NH_HD __host__ __device__ __forceinline__
#else
#define NH_HD inline __attribute__((always_inline))
#endif

namespace nh {

// ---------------------------------------------------------------- tables
// Quarter-wave table of the HEVC core transform: T32[i][j] = cosv(i*(2j+1)),
// T_N[i][j] = T32[i*32/N][j] (transform.py:28-135 hold the expanded matrices).
NH_HD constexpr int cos_q(int m) {
    constexpr int c[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                           61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
    return c[m];
}
NH_HD constexpr int cosv(int m) {
    m &= 127;
    return m <= 32 ? cos_q(m) : m <= 64 ? -cos_q(64 - m) : m <= 96 ? -cos_q(m - 64) : cos_q(128 - m);
}
template <int N>
NH_HD constexpr int dct(int i, int j) {
    return cosv((i * (32 / N)) * (2 * j + 1));
}
// transform.py:20-25
NH_HD constexpr int dst4(int i, int j) {
    constexpr int t[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};
    return t[i * 4 + j];
}

template <int N> struct Log2;
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };
template <> struct Log2<32> { static constexpr int v = 5; };

// intra.py:24-29 / 31-34.  Device code reads the tables from constant memory (a
// runtime-indexed constexpr local array would live in local memory).
#define NH_ANGLE_TABLE {32,  26,  21,  17,  13,  9,   5,  2,  0,  -2, -5, \
                        -9,  -13, -17, -21, -26, -32, -26, -21, -17, -13, -9, \
                        -5,  -2,  0,   2,   5,   9,   13,  17,  21,  26, 32}
#define NH_INV_ANGLE_TABLE {0, 0, 0, 0, 0, 0, 0, 0, 0, -4096, -1638, -910, -630, -482, -390, -315, -256, \
                            -315, -390, -482, -630, -910, -1638, -4096, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#if defined(__CUDACC__)
static __constant__ signed char kc_intra_angle[33] = NH_ANGLE_TABLE;
static __constant__ short kc_inv_angle[33] = NH_INV_ANGLE_TABLE;
#endif
NH_HD int intra_angle(int mode) {
#if defined(__CUDA_ARCH__)
    return kc_intra_angle[mode - 2];
#else
    constexpr int a[33] = NH_ANGLE_TABLE;
    return a[mode - 2];
#endif
}
// INV_ANGLE[INTRA_PRED_ANGLE[mode - 2]], 0 for the non-negative angles.
NH_HD int inv_angle_of_mode(int mode) {
#if defined(__CUDA_ARCH__)
    return kc_inv_angle[mode - 2];
#else
    constexpr int a[33] = NH_INV_ANGLE_TABLE;
    return a[mode - 2];
#endif
}
NH_HD constexpr int inv_angle(int angle) {
    return angle == -2 ? -4096 : angle == -5 ? -1638 : angle == -9 ? -910 : angle == -13 ? -630
         : angle == -17 ? -482 : angle == -21 ? -390 : angle == -26 ? -315 : angle == -32 ? -256 : 0;
}

// ------------------------------------------------------ 1-D transform cores
// Exact int32 sums (no rounding) via the even/odd decomposition; identical to
// the direct N-term dot products modulo 2^32, so results match the reference's
// wrap-around int32 accumulators for every input.
//
// fwd_core: y[i*S] = sum_k T_N[i][k] * x[k]            (x contiguous, y strided)
template <int N, int S>
NH_HD void fwd_core(const int* x, int* y) {
    if constexpr (N == 2) {
        y[0] = 64 * (x[0] + x[1]);
        y[S] = 64 * (x[0] - x[1]);
    } else {
        int e[N / 2], o[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            e[k] = x[k] + x[N - 1 - k];
            o[k] = x[k] - x[N - 1 - k];
        }
        fwd_core<N / 2, 2 * S>(e, y);
#pragma unroll
        for (int m = 0; m < N / 2; ++m) {
            int acc = 0;
#pragma unroll
            for (int k = 0; k < N / 2; ++k) acc += dct<N>(2 * m + 1, k) * o[k];
            y[(2 * m + 1) * S] = acc;
        }
    }
}

// inv_core: out[j] = sum_k T_N[k][j] * c[k*S]           (c strided, out contiguous)
template <int N, int S>
NH_HD void inv_core(const int* c, int* out) {
    if constexpr (N == 2) {
        out[0] = 64 * (c[0] + c[S]);
        out[1] = 64 * (c[0] - c[S]);
    } else {
        int e[N / 2], o[N / 2];
        inv_core<N / 2, 2 * S>(c, e);
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            int acc = 0;
#pragma unroll
            for (int m = 0; m < N / 2; ++m) acc += dct<N>(2 * m + 1, k) * c[(2 * m + 1) * S];
            o[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            out[k] = e[k] + o[k];
            out[N - 1 - k] = e[k] - o[k];
        }
    }
}

// ---- 2-way int16 x int8 dot products (IDP.2A) for the widest odd part ----------------------
// B200 issues IDP.2A at the IMAD rate (tools/ubench_int.cu), so packing two 16-bit operands per
// register halves the multiply instructions of the (N/2)^2-term odd part, which is 3/4 of all
// multiplies of a butterfly.  Exact as long as every packed operand fits int16; the kernels that
// use it prove that from the pixel domain of their inputs and fall back to
// the plain int32 cores otherwise.  The host build emulates the instruction INCLUDING the int16
// truncation of its lanes, so an out-of-range operand shows up as a mismatch in the CPU tests.
NH_HD int pack_s16(int lo, int hi) {
#if defined(__CUDA_ARCH__)
    return (int)__byte_perm((unsigned)lo, (unsigned)hi, 0x5410);
#else
    return (int)(((unsigned)lo & 0xffffu) | ((unsigned)hi << 16));
#endif
}
NH_HD int dp2a_s16s8(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, b, c);
#else
    return (int)(short)(a & 0xffff) * (int)(signed char)(b & 0xff) +
           (int)(short)((unsigned)a >> 16) * (int)(signed char)((b >> 8) & 0xff) + c;
#endif
}
NH_HD constexpr int coef_pair(int c0, int c1) { return (c0 & 0xff) | ((c1 & 0xff) << 8); }

template <int N, int S>
NH_HD void fwd_core_dp(const int* x, int* y) {
    int e[N / 2], o[N / 2], pk[N / 4];
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        e[k] = x[k] + x[N - 1 - k];
        o[k] = x[k] - x[N - 1 - k];
    }
    fwd_core<N / 2, 2 * S>(e, y);
#pragma unroll
    for (int t = 0; t < N / 4; ++t) pk[t] = pack_s16(o[2 * t], o[2 * t + 1]);
#pragma unroll
    for (int m = 0; m < N / 2; ++m) {
        int acc = 0;
#pragma unroll
        for (int t = 0; t < N / 4; ++t)
            acc = dp2a_s16s8(pk[t], coef_pair(dct<N>(2 * m + 1, 2 * t), dct<N>(2 * m + 1, 2 * t + 1)), acc);
        y[(2 * m + 1) * S] = acc;
    }
}

template <int N, int S>
NH_HD void inv_core_dp(const int* c, int* out) {
    int e[N / 2], o[N / 2], pk[N / 4];
    inv_core<N / 2, 2 * S>(c, e);
#pragma unroll
    for (int t = 0; t < N / 4; ++t) pk[t] = pack_s16(c[(4 * t + 1) * S], c[(4 * t + 3) * S]);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        int acc = 0;
#pragma unroll
        for (int t = 0; t < N / 4; ++t)
            acc = dp2a_s16s8(pk[t], coef_pair(dct<N>(4 * t + 1, k), dct<N>(4 * t + 3, k)), acc);
        o[k] = acc;
    }
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        out[k] = e[k] + o[k];
        out[N - 1 - k] = e[k] - o[k];
    }
}

// One rounded 1-D pass.  FWD: out[i] = (sum_k T[i][k] in[k] + rnd) >> shift
//                        INV: out[i] = (sum_k T[k][i] in[k] + rnd) >> shift
// DP = true selects the IDP.2A form of the widest odd part (operands must fit int16).
template <int N, bool DST, bool INV, bool DP = false>
NH_HD void pass1d(const int (&in)[N], int (&out)[N]) {
    constexpr int shift = Log2<N>::v + 5;
    constexpr int rnd = 1 << (shift - 1);
    int acc[N];
    if constexpr (DST && N == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int a = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += (INV ? dst4(k, i) : dst4(i, k)) * in[k];
            acc[i] = a;
        }
    } else if constexpr (DP && N >= 8) {
        if constexpr (INV) inv_core_dp<N, 1>(in, acc);
        else fwd_core_dp<N, 1>(in, acc);
    } else if constexpr (INV) {
        inv_core<N, 1>(in, acc);
    } else {
        fwd_core<N, 1>(in, acc);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = (acc[i] + rnd) >> shift;
}

// Whole 2-D transform of a block held by ONE thread (N = 4, 8).
// forward: temp = T.X (columns), coeff = temp.T^T (rows)  -- transform.py:180-194
// inverse: temp = T^T.C (columns), res = temp.T (rows)    -- transform.py:222-236
template <int N, bool DST, bool INV>
NH_HD void transform2d(int (&b)[N][N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        int col[N], o[N];
#pragma unroll
        for (int k = 0; k < N; ++k) col[k] = b[k][j];
        pass1d<N, DST, INV>(col, o);
#pragma unroll
        for (int k = 0; k < N; ++k) b[k][j] = o[k];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        int o[N];
        pass1d<N, DST, INV>(b[i], o);
#pragma unroll
        for (int k = 0; k < N; ++k) b[i][k] = o[k];
    }
}

// ------------------------------------------------------------------ quant
struct QuantParams {
    int mf;            // QUANT_SCALE[qp % 6]                        quant.py:21
    int q_shift;       // 14 + qp//6 + log2(size)                    quant.py:73
    long long q_offset;  // (1 << shift) // 3 (intra) or // 6 (inter)  quant.py:74
    int scale;         // DEQUANT_SCALE[qp % 6]                      quant.py:22
    int per;           // qp // 6
};

NH_HD QuantParams make_quant_params(int qp, int log2_size, int is_intra) {
    const int qs[6] = {26214, 23302, 20560, 18396, 16384, 14564};
    const int ds[6] = {40, 45, 51, 57, 64, 72};
    qp = qp < 0 ? 0 : (qp > 51 ? 51 : qp);  // quant.py:35
    QuantParams p;
    p.per = qp / 6;
    int rem = qp % 6;
    p.mf = qs[rem];
    p.scale = ds[rem];
    p.q_shift = 14 + p.per + log2_size;
    p.q_offset = is_intra ? ((1LL << p.q_shift) / 3) : ((1LL << p.q_shift) / 6);
    return p;
}

// quant.py:76-79 in exact int64 (np.abs on int32 wraps INT32_MIN first).
NH_HD int quantize_one(int c, const QuantParams& p) {
    int a32 = c < 0 ? (int)(0u - (unsigned)c) : c;
    long long lv = ((long long)a32 * (long long)p.mf + p.q_offset) >> p.q_shift;
    long long r = c > 0 ? lv : (c < 0 ? -lv : 0);
    return (int)(unsigned)(unsigned long long)r;
}

// quant.py:112-123
NH_HD int dequantize_one(int level, const QuantParams& p) {
    long long base = (long long)level * (long long)p.scale;
    long long c;
    if (p.per < 4) {
        int sh = 4 - p.per;
        c = (base + (1LL << (sh - 1))) >> sh;
    } else {
        c = (long long)((unsigned long long)base << (p.per - 4));
    }
    return (int)(unsigned)(unsigned long long)c;
}


// 32-bit quant / dequant for the PIXEL DOMAIN: exact whenever |coeff| * 26214 + offset < 2^31,
// which holds for every coefficient the forward transform can produce from residuals in
// [-4095, 4095] (|coeff| <= 32394).  Kernels check the domain of their inputs and
// fall back to the int64 form above otherwise.
struct FastQuant {
    int mf, q_shift, off_pos, off_neg;  // level = (c*mf + (c < 0 ? off_neg : off_pos)) >> q_shift
    int dq_mult, dq_rnd, dq_shift;      // coeff = (level*dq_mult + dq_rnd) >> dq_shift
};
NH_HD FastQuant make_fast_quant(const QuantParams& p) {
    FastQuant f;
    f.mf = p.mf;
    f.q_shift = p.q_shift;
    f.off_pos = (int)p.q_offset;
    // -floor((|c|*mf + off) / 2^s) == floor((c*mf + 2^s - 1 - off) / 2^s) for c < 0
    f.off_neg = (int)((1LL << p.q_shift) - 1 - p.q_offset);
    if (p.per < 4) {
        f.dq_mult = p.scale;
        f.dq_shift = 4 - p.per;
        f.dq_rnd = 1 << (f.dq_shift - 1);
    } else {
        f.dq_mult = p.scale << (p.per - 4);
        f.dq_shift = 0;
        f.dq_rnd = 0;
    }
    return f;
}
NH_HD int quantize_fast(int c, const FastQuant& f) {
    return (c * f.mf + (c < 0 ? f.off_neg : f.off_pos)) >> f.q_shift;
}
NH_HD int dequantize_fast(int level, const FastQuant& f) {
    return (level * f.dq_mult + f.dq_rnd) >> f.dq_shift;
}

// ------------------------------------------------------------ small helpers
NH_HD int clip_pixel(int v, int maxv) { return v < 0 ? 0 : (v > maxv ? maxv : v); }
NH_HD int sext16(int v) { return (int)(short)v; }

// intra.py:46-62: (sum + N) // (2N); python floor division.
template <int N>
NH_HD int dc_value(int sum) {
    return (sum + N) >> (Log2<N>::v + 1);  // floor shift == floor division by 2N
}

// intra.py:109-111
template <int N>
NH_HD int planar_px(int x, int y, int left_y, int top_x, int tr, int bl) {
    int h = (N - 1 - x) * left_y + (x + 1) * tr;
    int v = (N - 1 - y) * top_x + (y + 1) * bl;
    return (h + v + N) >> (Log2<N>::v + 1);
}

// intra.py:191-207.  Weighted sum evaluated in int16 like the reference (Q4);
// a and b are ref[idx], ref[idx+1].  For frac == 0 the formula returns a.
NH_HD int angular_px(int a, int b, int frac) {
    if (frac == 0) return a;
    int s = sext16(sext16(sext16((32 - frac) * a) + sext16(frac * b)) + 16);
    return s >> 5;
}

// ------------------------------------------------------------ angular modes
// intra.py:116-207.  `Ref` is any accessor with
//   int pri(int k)  -- primary array entry k (1..2N; the caller guarantees it is padded)
//   int sec(int k)  -- secondary array entry k (0 = that array's corner SLOT, Q3)
//   int corner()    -- the top_left ARGUMENT (ref[0])
struct AngleInfo {
    int angle;      // INTRA_PRED_ANGLE[mode - 2]
    int inv;        // INV_ANGLE[angle] (0 when angle >= 0)
    bool vertical;  // mode >= 18: primary = top, scan = row
};
NH_HD AngleInfo angle_info(int mode) {
    AngleInfo a;
    a.angle = intra_angle(mode);
    a.inv = inv_angle_of_mode(mode);
    a.vertical = mode >= 18;
    return a;
}

// ref[k] of _build_ref_array (intra.py:159-188) without materialising the array:
//   k == 0 -> top_left;  k > 0 -> primary[k];
//   k <  0 -> secondary[((k + 1) * inv_angle + 128) >> 8]   (note k+1: Q3)
template <class Ref>
NH_HD int ref_at(const Ref& r, int k, int inv) {
    if (k > 0) return r.pri(k);
    if (k == 0) return r.corner();
    return r.sec(((k + 1) * inv + 128) >> 8);
}

// One predicted sample: base b (position along the primary array), scan s.
template <class Ref>
NH_HD int angular_sample(const Ref& r, const AngleInfo& a, int b, int s) {
    int p = (s + 1) * a.angle;
    int ip = p >> 5, f = p & 31;
    int idx = b + 1 + ip;
    int v0 = ref_at(r, idx, a.inv);
    if (f == 0) return v0;
    return angular_px(v0, ref_at(r, idx + 1, a.inv), f);
}

}  // namespace nh
