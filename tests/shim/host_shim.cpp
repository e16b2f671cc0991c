// CPU harness around nano_hevc_b200/csrc/nh_math.cuh: runs the SAME integer
// code the kernels use (transform cores, quant, dequant, predictor formulas) on
// the host so tests/test_host_math.py can compare it with the oracle without a GPU.
#include "../../nano_hevc_b200/csrc/nh_math.cuh"
#include <cstdint>

using namespace nh;

template <int N, bool DST, bool INV>
static void run2d(const int32_t* in, int32_t* out) {
    int b[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) b[i][j] = in[i * N + j];
    transform2d<N, DST, INV>(b);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) out[i * N + j] = b[i][j];
}

// Same data flow as two_pass_transform (nh_block.cuh) with the IDP.2A butterflies.
template <int N, bool INV>
static void run2d_dp(const int32_t* in, int32_t* out) {
    static int m[N][N], t[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) m[i][j] = in[i * N + j];
    for (int j = 0; j < N; ++j) {  // pass 0: column j -> row j of the transposed intermediate
        int x[N], y[N];
        for (int k = 0; k < N; ++k) x[k] = m[k][j];
        pass1d<N, false, INV, true>(x, y);
        for (int i = 0; i < N; ++i) t[j][i] = y[i];
    }
    for (int i = 0; i < N; ++i) {  // pass 1: column i of the transposed intermediate -> row i
        int x[N], y[N];
        for (int k = 0; k < N; ++k) x[k] = t[k][i];
        pass1d<N, false, INV, true>(x, y);
        for (int j = 0; j < N; ++j) out[i * N + j] = y[j];
    }
}

extern "C" {
int shim_transform2d_dp(int n, int inv, const int32_t* in, int32_t* out) {
    switch (n) {
        case 8: inv ? run2d_dp<8, true>(in, out) : run2d_dp<8, false>(in, out); return 0;
        case 16: inv ? run2d_dp<16, true>(in, out) : run2d_dp<16, false>(in, out); return 0;
        case 32: inv ? run2d_dp<32, true>(in, out) : run2d_dp<32, false>(in, out); return 0;
    }
    return -1;
}
int shim_transform2d(int n, int dst, int inv, const int32_t* in, int32_t* out) {
    switch (n) {
        case 4:
            if (dst) { inv ? run2d<4, true, true>(in, out) : run2d<4, true, false>(in, out); }
            else { inv ? run2d<4, false, true>(in, out) : run2d<4, false, false>(in, out); }
            return 0;
        case 8: inv ? run2d<8, false, true>(in, out) : run2d<8, false, false>(in, out); return 0;
        case 16: inv ? run2d<16, false, true>(in, out) : run2d<16, false, false>(in, out); return 0;
        case 32: inv ? run2d<32, false, true>(in, out) : run2d<32, false, false>(in, out); return 0;
    }
    return -1;
}
void shim_quant(const int32_t* c, int32_t* out, int64_t n, int qp, int log2n, int intra) {
    QuantParams p = make_quant_params(qp, log2n, intra);
    for (int64_t i = 0; i < n; ++i) out[i] = quantize_one(c[i], p);
}
void shim_dequant(const int32_t* l, int32_t* out, int64_t n, int qp) {
    QuantParams p = make_quant_params(qp, 2, 1);
    for (int64_t i = 0; i < n; ++i) out[i] = dequantize_one(l[i], p);
}
void shim_quant_fast(const int32_t* c, int32_t* out, int64_t n, int qp, int log2n, int intra) {
    FastQuant f = make_fast_quant(make_quant_params(qp, log2n, intra));
    for (int64_t i = 0; i < n; ++i) out[i] = quantize_fast(c[i], f);
}
void shim_dequant_fast(const int32_t* l, int32_t* out, int64_t n, int qp) {
    FastQuant f = make_fast_quant(make_quant_params(qp, 2, 1));
    for (int64_t i = 0; i < n; ++i) out[i] = dequantize_fast(l[i], f);
}
int shim_matrix(int n, int i, int j) {
    switch (n) { case 4: return dct<4>(i, j); case 8: return dct<8>(i, j);
                 case 16: return dct<16>(i, j); case 32: return dct<32>(i, j); }
    return 0;
}
int shim_angle(int mode) { return intra_angle(mode); }
int shim_inv_angle(int a) { return inv_angle(a); }
}

// ---- predictors through the same nh_math.cuh helpers the kernels use ----
namespace {
struct HostRef {
    const int16_t* p;
    const int16_t* s;
    int c;
    int pri(int k) const { return p[k]; }
    int sec(int k) const { return s[k]; }
    int corner() const { return c; }
};
template <int N>
void predict_mode(const int16_t* top, const int16_t* left, int corner, int mode, int16_t* out) {
    if (mode == 1) {
        int s = 0;
        for (int k = 1; k <= N; ++k) s += top[k] + left[k];
        int dc = dc_value<N>(s);
        for (int e = 0; e < N * N; ++e) out[e] = (int16_t)dc;
        return;
    }
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            int v;
            if (mode == 0) {
                v = planar_px<N>(x, y, left[1 + y], top[1 + x], top[N + 1], left[N + 1]);
            } else {
                AngleInfo ai = angle_info(mode);
                HostRef r;
                r.c = corner;
                if (ai.vertical) { r.p = top; r.s = left; v = angular_sample(r, ai, x, y); }
                else { r.p = left; r.s = top; v = angular_sample(r, ai, y, x); }
            }
            out[y * N + x] = (int16_t)v;
        }
}
}  // namespace

extern "C" int shim_predict_mode(int n, const int16_t* top, const int16_t* left, int corner, int mode,
                                 int16_t* out) {
    switch (n) {
        case 4: predict_mode<4>(top, left, corner, mode, out); return 0;
        case 8: predict_mode<8>(top, left, corner, mode, out); return 0;
        case 16: predict_mode<16>(top, left, corner, mode, out); return 0;
        case 32: predict_mode<32>(top, left, corner, mode, out); return 0;
    }
    return -1;
}
