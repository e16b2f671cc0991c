// Integer-pipe micro-benchmark: issue rate of IMAD, IDP.2A (dp2a), IDP.4A (dp4a), PRMT, VABSDIFF4,
// VIADD.16x2, SHF, LOP3 on sm_100a.  8 independent chains per thread, 1024 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int tools/ubench_int.cu && ./ubench_int
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed, int iters) {
    unsigned a[8], b = seed | 1u, c = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 8 + i + seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = a[i] * b + c;                                   // IMAD
            if (OP == 1) a[i] = __dp2a_lo((int)a[i], (int)b, (int)c) ^ a[i];    // IDP.2A + LOP (see OP 7)
            if (OP == 2) a[i] = __dp4a((int)a[i], (int)b, (int)a[i]);           // IDP.4A
            if (OP == 3) a[i] = __byte_perm(a[i], b, c & 0x7777);               // PRMT
            if (OP == 4) a[i] = __vsadu4(a[i], b) + a[i];                       // VABSDIFF4.ACC
            if (OP == 5) a[i] = __vadd2(a[i], b);                               // VIADD.16x2
            if (OP == 6) a[i] = __funnelshift_r(a[i], b, 5);                    // SHF
            if (OP == 7) a[i] = (a[i] ^ b) & c;                                 // LOP3
            if (OP == 8) a[i] = __dp2a_lo((int)a[i], (int)b, (int)a[i]);        // IDP.2A accumulate chain
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, unsigned* out, int sms, double opsPerIter) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<sms * 4, 256>>>(out, 12345u, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<sms * 4, 256>>>(out, 12345u, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double threadOps = (double)sms * 4 * 256 * iters * 8 * opsPerIter;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double perSmPerClk = threadOps / (ms * 1e-3) / sms / (clk * 1e3);
    printf("%-28s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, perSmPerClk, clk / 1000);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* out;
    cudaMalloc(&out, sms * 4 * 256 * sizeof(unsigned));
    run<0>("IMAD", out, sms, 1);
    run<7>("LOP3 (x1)", out, sms, 1);
    run<1>("IDP.2A + LOP3", out, sms, 2);
    run<8>("IDP.2A chain", out, sms, 1);
    run<2>("IDP.4A", out, sms, 1);
    run<3>("PRMT", out, sms, 1);
    run<4>("VABSDIFF4.ACC", out, sms, 1);
    run<5>("VIADD.16x2", out, sms, 1);
    run<6>("SHF", out, sms, 1);
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
