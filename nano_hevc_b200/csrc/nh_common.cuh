// nh_common.cuh -- error plumbing, launch geometry and the warp-level staging
// helpers shared by every translation unit of libnh_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nh_b200.h"
#include "nh_math.cuh"

#define NH_API extern "C" __attribute__((visibility("default")))

namespace nh {

// Thread-local description of the last failure (nh_last_error()).
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

inline int log2_size(int size) {
    switch (size) {
        case 4: return 2;
        case 8: return 3;
        case 16: return 4;
        case 32: return 5;
        default: return -1;
    }
}

// Persistent-style grid: `ctas_per_sm` CTAs on every SM of the current device,
// never more CTAs than there are work items of `items_per_cta`.
int sm_count();
inline int grid_for(int64_t items, int64_t items_per_cta, int ctas_per_sm) {
    int64_t need = (items + items_per_cta - 1) / items_per_cta;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// nh_fused.cu: K6 with int16 coefficient / level outputs, used by the host-buffer pipeline.
int fused_pipeline_dcplanar_narrow(const int16_t* orig, const int16_t* top, const int16_t* left,
                                   const int16_t* top_right, const int16_t* bottom_left,
                                   const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                                   int is_intra, int use_dst, int bit_depth, int16_t* pred,
                                   int16_t* coeff16, int16_t* levels16, int16_t* recon, int* ood_flag,
                                   cudaStream_t st);

// nh_fused.cu (nh_coder8.cuh): K7 winner stage for 8x8 blocks of an 8-bit plane whose modes are decided.
struct QuantParams;
// *handed_back = device counter pair {tiles handed back to the exact coder, readers} of this stream.
int coder8_plane_mma(const int16_t* src, int n_frames, int64_t frame_stride, int H, int W, int pitch, uint8_t* modes,
                     int16_t* pred, int32_t* coeff, int32_t* levels, int16_t* recon_plane, const QuantParams& qp,
                     int maxv, cudaStream_t st, unsigned int** handed_back);

// nh_fused.cu: 2 = tensor-core kernels for N = 16 / 32 (default), 1 = CUDA-core butterflies
// (nh_set_rows_impl / NH_ROWS_IMPL); also selects the single-stage transform kernels of nh_ops.cu.
int rows_impl();

// Opt a kernel into more than 48 KB of dynamic shared memory, once per (kernel, device): function
// attributes are per context and a process may drive several GPUs.  nh_api.cu.
int ensure_dynamic_smem_impl(const void* kernel, int bytes, const char* what);
template <class Kernel>
inline int ensure_dynamic_smem(Kernel kernel, int bytes, const char* what) {
    return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kernel), bytes, what);
}

// Work counters for a kernel that hands out its warp tiles dynamically: counter[0] = next ticket,
// counter[1] = warps that have finished.  The pair is zero when a launch starts and the last warp to
// finish zeroes it again (release_tile_counter), so no reset has to be enqueued between launches.
// Counters live in a static device array (nothing is allocated); a stream keeps the slot it was
// first given, so launches on one stream (serialised by the stream) and on different streams
// (different slots) never interfere.  counter[2] / counter[3] belong to the frame coder (nh_coder8.cuh:
// tiles handed back to the exact coder kernel, and the CTAs of that kernel that have looked).
// tile_counter_launched() is called behind the LAST launch that uses the pair: it records the event
// that lets the slot move to another stream once that launch has finished (only needed after more
// than 4096 distinct streams).  nh_api.cu.
int acquire_tile_counter(cudaStream_t stream, unsigned int** counter);
void tile_counter_launched(cudaStream_t stream);

#define NH_CHECK_LAUNCH(what)                                  \
    do {                                                       \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return nh::cuda_fail(e__, what); \
    } while (0)

#if defined(__CUDACC__)
// Called by every warp (all lanes) after it has drawn its last ticket: the last of the launch's
// `total_warps` warps re-arms the counter pair for the next launch.
__device__ __forceinline__ void release_tile_counter(unsigned int* counter, unsigned int total_warps) {
    if ((threadIdx.x & 31) == 0) {
        const unsigned int done = atomicAdd(counter + 1, 1u);
        if (done == total_warps - 1) {
            counter[0] = 0;
            counter[1] = 0;
        }
    }
}

// ----------------------------------------------------------- device helpers
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)w >> 16; }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) {
    return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
}

// Streaming (evict-first) 128-bit global accesses: every tensor of the
// pipeline is touched exactly once, so nothing should linger in L2.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    return __ldcs(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
    __stcs(reinterpret_cast<uint4*>(p), v);
}


// ---- asynchronous copies -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// LDGSTS: 16 bytes global -> shared without a register round trip (L2 only, no L1 allocation).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
#ifdef NH_CPASYNC_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}
// TMA 1-D bulk store shared -> global (SASS: UBLKCP), tracked by per-thread bulk groups.
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until at most PENDING of this thread's bulk groups still have to READ their source.
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
// Make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store.
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA tensor-map copies + mbarrier (warp-uniform: issued by ONE elected lane) ------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    } while (!done);
}
// 2-D tile global -> shared (SASS: UTMALDG); completion is signalled on `mbar` as transaction bytes.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(mbar)
        : "memory");
}
// 2-D tile shared -> global (SASS: UTMASTG), tracked by the issuing thread's bulk groups.
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, uint32_t smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap),
                 "r"(c0), "r"(c1), "r"(smem_src)
                 : "memory");
}
template <int PENDING>
__device__ __forceinline__ void bulk_wait_all() {  // full completion (global writes performed)
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(PENDING) : "memory");
}

// Warp tile staging.  A warp tile is 32 "units" of UNIT_BYTES contiguous
// bytes in global memory (one unit per lane).  In shared memory unit u lives
// at u * (UNIT_BYTES + 16): the 16-byte pad makes the per-lane 128-bit
// accesses (lane l touching its own unit) conflict-free, while the cooperative
// copy below (8 or 16 consecutive lanes sweep one unit) stays conflict-free too.
template <int UNIT_BYTES>
struct WarpTile {
    static constexpr int kPitch = UNIT_BYTES + 16;
    static constexpr int kBytes = 32 * kPitch;
    static constexpr int kChunksPerUnit = UNIT_BYTES / 16;
    static constexpr int kIters = kChunksPerUnit;  // 32 lanes x 16 B per iteration

    // global (linear) -> shared (padded); chunks_valid = number of valid 16-byte chunks
    // (< 32 * kChunksPerUnit only on the ragged last tile).
    static __device__ __forceinline__ void load(unsigned char* smem, const unsigned char* gmem,
                                                int lane, int chunks_valid) {
        uint4 v[kIters];  // all loads in flight before the first shared-memory store
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            int c = it * 32 + lane;
            v[it] = (c < chunks_valid) ? ldg_stream(gmem + (size_t)c * 16) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            int c = it * 32 + lane;
            int u = c / kChunksPerUnit, k = c % kChunksPerUnit;
            *reinterpret_cast<uint4*>(smem + u * kPitch + k * 16) = v[it];
        }
    }
    // shared (padded) -> global (linear)
    static __device__ __forceinline__ void store(const unsigned char* smem, unsigned char* gmem,
                                                 int lane, int chunks_valid) {
        uint4 v[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            int c = it * 32 + lane;
            int u = c / kChunksPerUnit, k = c % kChunksPerUnit;
            v[it] = *reinterpret_cast<const uint4*>(smem + u * kPitch + k * 16);
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            int c = it * 32 + lane;
            if (c < chunks_valid) stg_stream(gmem + (size_t)c * 16, v[it]);
        }
    }
    static __device__ __forceinline__ uint4* unit(unsigned char* smem, int lane) {
        return reinterpret_cast<uint4*>(smem + lane * kPitch);
    }
};
#endif  // __CUDACC__

}  // namespace nh
