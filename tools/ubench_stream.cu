// ubench_stream.cu -- what HBM delivers for the fused pipeline's traffic MIX (about 1 byte read per
// 4.7 bytes written) with no arithmetic at all, next to a plain copy (1:1, the mix behind
// MEASURED_PEAKS.json).  Gives the practical ceiling for a write-dominated stream on this part.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_stream tools/ubench_stream.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

template <int NW>  // one 16-byte read -> NW 16-byte writes (NW distinct output streams)
__global__ void __launch_bounds__(256) stream_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 v = __ldcs(in + i);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint4 o = make_uint4(v.x + w, v.y, v.z, v.w);
            __stcs(out + (size_t)w * n + i, o);
        }
    }
}
// the pipeline's real mix: read 2 + 0.5625 B/px, write 2 + 4 + 4 + 2 B/px  (per 8 px: 16 B orig, refs 4.5 B)
__global__ void __launch_bounds__(256) mix_kernel(const uint4* __restrict__ orig, const uint4* __restrict__ refs,
                                                  uint4* __restrict__ p16, uint4* __restrict__ c32,
                                                  uint4* __restrict__ l32, uint4* __restrict__ r16, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 v = __ldcs(orig + i);
        if ((i & 3) == 0) { uint4 r = __ldcs(refs + (i >> 2)); v.x ^= r.x; }  // ~4 B of refs per 16 B of pixels
        __stcs(p16 + i, v);
        __stcs(r16 + i, v);
        __stcs(c32 + 2 * i, v);
        __stcs(c32 + 2 * i + 1, v);
        __stcs(l32 + 2 * i, v);
        __stcs(l32 + 2 * i + 1, v);
    }
}

int main(int argc, char** argv) {
    const size_t n = (size_t)(argc > 1 ? atol(argv[1]) : 32) << 20;  // 16-byte elements (default 512 MB in)
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint4 *in, *out;
    cudaMalloc(&in, n * 16 + (n / 4 + 1) * 16);
    cudaMalloc(&out, n * 16 * 6);
    cudaMemset(in, 1, n * 16 + (n / 4 + 1) * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time = [&](auto launch, const char* name, double bytes) {
        for (int i = 0; i < 3; ++i) launch();
        float best = 1e30f, tot = 0;
        for (int i = 0; i < 10; ++i) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
            tot += ms;
        }
        printf("{\"kernel\": \"%s\", \"GBs_best\": %.1f, \"GBs_mean\": %.1f, \"bytes\": %.0f}\n", name,
               bytes / best / 1e6, bytes / (tot / 10) / 1e6, bytes);
    };
    for (int cps = 4; cps <= 16; cps *= 2) {
        const int grid = sms * cps;
        printf("# %d CTAs of 256 threads per SM\n", cps);
        time([&] { stream_kernel<1><<<grid, 256>>>(in, out, n); }, "copy 1R:1W", n * 16.0 * 2);
        time([&] { stream_kernel<4><<<grid, 256>>>(in, out, n); }, "stream 1R:4W", n * 16.0 * 5);
        time([&] { stream_kernel<5><<<grid, 256>>>(in, out, n); }, "stream 1R:5W", n * 16.0 * 6);
        time([&] { mix_kernel<<<grid, 256>>>(in, in + n, out, out + n, out + 3 * n, out + 5 * n, n); },
             "pipeline mix 2.56R:12W per px", n * 16.0 * 7 + (n / 4) * 16.0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
