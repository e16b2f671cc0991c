// nh_host.cu -- host-buffer entry point of the K6 pipeline (the end-to-end path a caller with
// numpy / C arrays uses).  PCIe is the bound here (14.56 B/px would have to cross it), so:
//   * the batch is cut into chunks that rotate over three streams: H2D of chunk i+1, the kernel of
//     chunk i and D2H of chunk i-1 overlap (B200 has separate copy engines per direction);
//   * coefficients and levels cross the bus as int16 (in the pixel domain |coeff| <= 32394 and
//     |level| <= 13600) into pinned staging buffers and are widened to the
//     caller's int32 arrays by a small pool of host threads (AVX2 sign-extension + streaming
//     stores) while the next chunks are in flight: 8 instead of 12 output bytes per pixel;
//   * on top of that a compact wire format produced by a small device kernel: 64-element segments
//     whose levels are all zero -- the overwhelming majority in coded video, HEVC signals exactly
//     this as the coded-block flag -- are not transferred at all (the host zero-fills and scatters
//     the non-zero segments from a fixed-capacity list); coefficients travel as int8 with the rare
//     segments holding |coeff| > 127 in an int16 exception list.  5.3 instead of 8 output bytes per
//     pixel; a chunk that overflows a list fetches that tensor as int16 after all;
//   * a chunk in which any block left the pixel domain raises a device flag and is redone through
//     the plain int32 path, so the result is bit-exact for every input.
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "nh_common.cuh"

namespace nh {

constexpr int kSlots = 3;

// ------------------------------------------------------------------ host thread pool
class Pool {
  public:
    explicit Pool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }
    // fn(part, parts) on every worker and on the caller; returns when all parts are done
    void run(const std::function<void(int, int)>& fn) {
        if (n_ == 1) { fn(0, 1); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            pending_ = n_ - 1;
            ++epoch_;
        }
        cv_.notify_all();
        fn(0, n_);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
    }

  private:
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(idx, n_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)>* fn_ = nullptr;
    int pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

__attribute__((target("avx2"))) static void widen_avx2(const int16_t* src, int32_t* dst, size_t n) {
    size_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i lo = _mm256_cvtepi16_epi32(_mm256_castsi256_si128(v));
        const __m256i hi = _mm256_cvtepi16_epi32(_mm256_extracti128_si256(v, 1));
        if (aligned) {
            _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), lo);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 8), hi);
        } else {
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), lo);
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i + 8), hi);
        }
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

static void widen_scalar(const int16_t* src, int32_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}

static void widen(const int16_t* src, int32_t* dst, size_t n) {
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) widen_avx2(src, dst, n);
    else widen_scalar(src, dst, n);
}


__attribute__((target("avx2"))) static void widen8_avx2(const int8_t* src, int32_t* dst, size_t n) {
    size_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m256i lo = _mm256_cvtepi8_epi32(v);
        const __m256i hi = _mm256_cvtepi8_epi32(_mm_srli_si128(v, 8));
        if (aligned) {
            _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), lo);
            _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 8), hi);
        } else {
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), lo);
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i + 8), hi);
        }
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}
static void widen8(const int8_t* src, int32_t* dst, size_t n) {
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) widen8_avx2(src, dst, n);
    else for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}
__attribute__((target("avx2"))) static void zero_fill_avx2(int32_t* dst, size_t n) {
    size_t i = 0;
    const __m256i z = _mm256_setzero_si256();
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31) != 0; ++i) dst[i] = 0;
    for (; i + 8 <= n; i += 8) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), z);
    for (; i < n; ++i) dst[i] = 0;
    _mm_sfence();
}
static void zero_fill(int32_t* dst, size_t n) {
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) zero_fill_avx2(dst, n);
    else for (size_t i = 0; i < n; ++i) dst[i] = 0;
}

// ------------------------------------------------------------------ compact wire format
constexpr int kSeg = 64;  // elements per segment (128 bytes of int16)

struct WireCounters {  // written by pack_wire_kernel, read by the host before the D2H copies
    int ood;           // a block left the pixel domain (set by the pipeline kernel)
    int n_wide;        // coefficient segments with an |value| > 127
    int n_nz;          // level segments with a non-zero value
    int pad;
};

// Half a warp per segment: coefficients narrowed to int8, segments that do not fit appended to the
// `wide` list, level segments with any non-zero value appended to the `nz` list (index + the 64
// int16 values).  Lists are capped at `cap` entries; the counters keep counting so the host can see
// an overflow and fall back to the int16 format for that tensor.
__global__ void __launch_bounds__(256) pack_wire_kernel(const int16_t* __restrict__ coeff16,
                                                        const int16_t* __restrict__ levels16, int64_t elems,
                                                        int8_t* __restrict__ coeff8, int* __restrict__ wide_idx,
                                                        int16_t* __restrict__ wide_val, int* __restrict__ nz_idx,
                                                        int16_t* __restrict__ nz_val, int cap,
                                                        WireCounters* __restrict__ cnt) {
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int64_t n_seg = (elems + kSeg - 1) / kSeg;
    const int64_t n_pairs = (n_seg + 1) / 2;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_pairs; p += warps) {
        const int64_t seg = 2 * p + half;
        const int64_t e0 = seg * kSeg + 4 * hl;  // this lane's 4 elements
        uint2 c = make_uint2(0u, 0u), l = make_uint2(0u, 0u);
        const bool in = seg < n_seg;
        if (in && e0 + 4 <= elems) {
            if (coeff16) c = *reinterpret_cast<const uint2*>(coeff16 + e0);
            if (levels16) l = *reinterpret_cast<const uint2*>(levels16 + e0);
        } else if (in) {
            short cv[4] = {0, 0, 0, 0}, lv[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k)
                if (e0 + k < elems) {
                    if (coeff16) cv[k] = coeff16[e0 + k];
                    if (levels16) lv[k] = levels16[e0 + k];
                }
            c = make_uint2((uint32_t)(uint16_t)cv[0] | ((uint32_t)(uint16_t)cv[1] << 16),
                           (uint32_t)(uint16_t)cv[2] | ((uint32_t)(uint16_t)cv[3] << 16));
            l = make_uint2((uint32_t)(uint16_t)lv[0] | ((uint32_t)(uint16_t)lv[1] << 16),
                           (uint32_t)(uint16_t)lv[2] | ((uint32_t)(uint16_t)lv[3] << 16));
        }
        const int c0 = (short)(c.x & 0xffff), c1 = (int)c.x >> 16, c2 = (short)(c.y & 0xffff), c3 = (int)c.y >> 16;
        const bool wide_l = (unsigned)(c0 + 128) > 255u || (unsigned)(c1 + 128) > 255u ||
                            (unsigned)(c2 + 128) > 255u || (unsigned)(c3 + 128) > 255u;
        const bool nz_l = (l.x | l.y) != 0;
        const unsigned wide_b = __ballot_sync(0xffffffffu, wide_l) & hmask;
        const unsigned nz_b = __ballot_sync(0xffffffffu, nz_l) & hmask;
        if (in && coeff16) {
            const uint32_t packed = (uint32_t)(c0 & 0xff) | ((uint32_t)(c1 & 0xff) << 8) |
                                    ((uint32_t)(c2 & 0xff) << 16) | ((uint32_t)(c3 & 0xff) << 24);
            if (e0 + 4 <= elems) *reinterpret_cast<uint32_t*>(coeff8 + e0) = packed;
            else for (int k = 0; k < 4; ++k) if (e0 + k < elems) coeff8[e0 + k] = (int8_t)(packed >> (8 * k));
        }
        int slot_w = -1, slot_z = -1;
        if (hl == 0 && in) {
            if (wide_b) slot_w = atomicAdd(&cnt->n_wide, 1);
            if (nz_b) slot_z = atomicAdd(&cnt->n_nz, 1);
        }
        slot_w = __shfl_sync(0xffffffffu, slot_w, half * 16);
        slot_z = __shfl_sync(0xffffffffu, slot_z, half * 16);
        if (slot_w >= 0 && slot_w < cap) {
            if (hl == 0) wide_idx[slot_w] = (int)seg;
            *reinterpret_cast<uint2*>(wide_val + (int64_t)slot_w * kSeg + 4 * hl) = c;
        }
        if (slot_z >= 0 && slot_z < cap) {
            if (hl == 0) nz_idx[slot_z] = (int)seg;
            *reinterpret_cast<uint2*>(nz_val + (int64_t)slot_z * kSeg + 4 * hl) = l;
        }
    }
}

// ------------------------------------------------------------------ per-device context
struct SlotLayout {
    int64_t orig, top, left, tr, bl, modes, pred, coeff, levels, recon, coeff16, levels16;
    int64_t coeff8, wide_idx, wide_val, nz_idx, nz_val, cnt, total;
    int64_t cap;  // list capacity in segments
};

static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

// Exception / non-zero lists hold at most 1/16 of the chunk's segments and always travel at that
// size (+0.26 bytes per pixel), so that a chunk's copies can be enqueued without waiting for its
// counters; a chunk that overflows a list fetches that tensor as int16 afterwards.
static int64_t list_cap(int size, int64_t chunk) {
    const int64_t segs = (chunk * size * size + kSeg - 1) / kSeg;
    return segs / 16 + 1;
}

static SlotLayout slot_layout(int size, int64_t chunk) {
    const int64_t nn = (int64_t)size * size;
    SlotLayout l;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += align256(bytes); return o; };
    l.cap = list_cap(size, chunk);
    l.orig = take(chunk * nn * 2);
    l.top = take(chunk * size * 2);
    l.left = take(chunk * size * 2);
    l.tr = take(chunk * 2);
    l.bl = take(chunk * 2);
    l.modes = take(chunk);
    l.pred = take(chunk * nn * 2);
    l.coeff = take(chunk * nn * 4);     // int32 areas: only used when a chunk has to be redone
    l.levels = take(chunk * nn * 4);
    l.recon = take(chunk * nn * 2);
    l.coeff16 = take(chunk * nn * 2);
    l.levels16 = take(chunk * nn * 2);
    l.coeff8 = take(chunk * nn);
    l.wide_idx = take(l.cap * 4);
    l.wide_val = take(l.cap * kSeg * 2);
    l.nz_idx = take(l.cap * 4);
    l.nz_val = take(l.cap * kSeg * 2);
    l.cnt = take(sizeof(WireCounters));
    l.total = off;
    return l;
}

struct Staging {  // pinned host memory of one slot
    int16_t* coeff16 = nullptr;
    int16_t* levels16 = nullptr;
    int8_t* coeff8 = nullptr;
    int* wide_idx = nullptr;
    int16_t* wide_val = nullptr;
    int* nz_idx = nullptr;
    int16_t* nz_val = nullptr;
};

struct DeviceCtx {
    cudaStream_t s[kSlots];
    cudaEvent_t done[kSlots];     // the chunk's D2H copies are complete
    Staging st[kSlots];
    WireCounters* cnt = nullptr;  // kSlots entries, pinned
    int64_t stage_elems = 0, stage_cap = 0;
    bool ready = false;
};

static std::mutex g_mu;
static int64_t g_last_h2d = 0, g_last_d2h = 0;  // bytes the last host pipeline call moved over PCIe
static DeviceCtx g_ctx[64];
static Pool* g_pool = nullptr;

static Pool& pool() {
    if (!g_pool) {
        int n = 0;
        if (const char* e = getenv("NH_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            n = (int)std::thread::hardware_concurrency();
            if (n > 6) n = 6;   // more threads than saturate host memory only delay the DMA traffic (profiles/r4_host_threads.txt)
        }
        g_pool = new Pool(n);
    }
    return *g_pool;
}

static int get_ctx(int64_t stage_elems, int64_t cap, DeviceCtx** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return NH_E_ARG; }
    DeviceCtx& c = g_ctx[dev];
    if (!c.ready) {
        for (int i = 0; i < kSlots; ++i) {
            e = cudaStreamCreateWithFlags(&c.s[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithFlags");
            e = cudaEventCreateWithFlags(&c.done[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreateWithFlags");
        }
        e = cudaHostAlloc(reinterpret_cast<void**>(&c.cnt), kSlots * sizeof(WireCounters), cudaHostAllocDefault);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc(counters)");
        c.ready = true;
    }
    if (c.stage_elems < stage_elems || c.stage_cap < cap) {
        if (stage_elems < c.stage_elems) stage_elems = c.stage_elems;
        if (cap < c.stage_cap) cap = c.stage_cap;
        for (int i = 0; i < kSlots; ++i) {
            Staging& t = c.st[i];
            void* olds[] = {t.coeff16, t.levels16, t.coeff8, t.wide_idx, t.wide_val, t.nz_idx, t.nz_val};
            for (void* p : olds) if (p) cudaFreeHost(p);
            t = Staging{};
            auto alloc = [&](auto** p, size_t bytes) {
                if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(p), bytes, cudaHostAllocDefault);
            };
            e = cudaSuccess;
            alloc(&t.coeff16, stage_elems * 2);
            alloc(&t.levels16, stage_elems * 2);
            alloc(&t.coeff8, stage_elems);
            alloc(&t.wide_idx, cap * 4);
            alloc(&t.wide_val, cap * kSeg * 2);
            alloc(&t.nz_idx, cap * 4);
            alloc(&t.nz_val, cap * kSeg * 2);
            if (e != cudaSuccess) { c.stage_elems = c.stage_cap = 0; return cuda_fail(e, "cudaHostAlloc(staging)"); }
        }
        c.stage_elems = stage_elems;
        c.stage_cap = cap;
    }
    *out = &c;
    return NH_OK;
}

// NH_WIRE=int16 disables the compact format (A/B measurements).
static bool compact_wire_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NH_WIRE");
        v = (e && e[0] == 'i') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace nh

using namespace nh;

NH_API int nh_host_pipeline_last_transfer(int64_t* h2d_bytes, int64_t* d2h_bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h2d_bytes) *h2d_bytes = g_last_h2d;
    if (d2h_bytes) *d2h_bytes = g_last_d2h;
    return NH_OK;
}

NH_API int64_t nh_host_pipeline_scratch_bytes(int size, int64_t chunk_blocks) {
    if (log2_size(size) < 0 || chunk_blocks <= 0) return 0;
    return kSlots * slot_layout(size, chunk_blocks).total;
}

NH_API int nh_host_pipeline_dcplanar(const int16_t* orig, const int16_t* top, const int16_t* left,
                                     const int16_t* top_right, const int16_t* bottom_left,
                                     const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                                     int is_intra, int use_dst, int bit_depth, int16_t* pred,
                                     int32_t* coeff, int32_t* levels, int16_t* recon,
                                     void* device_scratch, int64_t scratch_bytes, int64_t chunk_blocks) {
    if (log2_size(size) < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (n_blocks == 0) return NH_OK;
    if (!orig || !top || !left || !top_right || !bottom_left || n_blocks < 0 || chunk_blocks <= 0) {
        set_error("nh_host_pipeline_dcplanar: null input, negative count or chunk_blocks <= 0");
        return NH_E_ARG;
    }
    if (!modes && mode != 0 && mode != 1) {
        set_error("nh_host_pipeline_dcplanar: mode must be 0 (planar) or 1 (DC), got %d", mode);
        return NH_E_ARG;
    }
    if (bit_depth < 1 || bit_depth > 15) {
        set_error("nh_host_pipeline_dcplanar: bit_depth %d out of range", bit_depth);
        return NH_E_ARG;
    }
    const SlotLayout L = slot_layout(size, chunk_blocks);
    if (!device_scratch || scratch_bytes < kSlots * L.total) {
        set_error("nh_host_pipeline_dcplanar: device scratch of %lld bytes required, got %lld",
                  (long long)(kSlots * L.total), (long long)scratch_bytes);
        return NH_E_NOMEM;
    }
    const int64_t nn = (int64_t)size * size;
    if (chunk_blocks * nn / kSeg >= (int64_t)1 << 31) {
        set_error("nh_host_pipeline_dcplanar: chunk_blocks too large");
        return NH_E_ARG;
    }
    std::lock_guard<std::mutex> lk(g_mu);  // one host pipeline at a time per process (shared staging)
    DeviceCtx* ctx = nullptr;
    int rc = get_ctx(chunk_blocks * nn, L.cap, &ctx);
    if (rc != NH_OK) return rc;
    Pool& workers = pool();
    const bool compact = compact_wire_enabled();
    // Which format a chunk's coefficients / levels travel in is decided when the chunk is ENQUEUED, from the list
    // counters of the chunks that have already come back: content whose segments overflow the exception lists
    // (noise: every level segment non-zero) would otherwise pay the compact transfer, a host round trip and a
    // second, unoverlapped int16 transfer per chunk.  A wrong guess costs time, never correctness (compact + overflow
    // still falls back below; int16 is always complete).  Per call, starting from the compact format: no state is
    // kept between calls.
    bool s_coeff8 = true, s_levelz = true;
    bool fmt_coeff8[kSlots] = {false, false, false}, fmt_levelz[kSlots] = {false, false, false};
    unsigned char* base = reinterpret_cast<unsigned char*>(device_scratch);
    const int64_t n_chunks = (n_blocks + chunk_blocks - 1) / chunk_blocks;
    // On any error path copies may still be in flight on the internal streams, targeting the caller's buffers, the
    // pinned staging and device_scratch: nothing returns before they have drained.
    struct Drain {
        DeviceCtx* c;
        bool armed;
        ~Drain() { if (armed) for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->s[i]); }
    } drain{ctx, true};

    g_last_h2d = g_last_d2h = 0;
#define NH_CP(dst, src, bytes, kind, s)                                        \
    do {                                                                       \
        cudaError_t e__ = cudaMemcpyAsync(dst, src, (size_t)(bytes), kind, s); \
        if (e__ != cudaSuccess) return cuda_fail(e__, "cudaMemcpyAsync");      \
        (kind == cudaMemcpyHostToDevice ? g_last_h2d : g_last_d2h) += (int64_t)(bytes); \
    } while (0)

    static const bool trace = getenv("NH_HOST_TRACE") != nullptr;
    double t_acc[4] = {0, 0, 0, 0};
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto chunk_range = [&](int64_t i, int64_t& first, int64_t& n) {
        first = i * chunk_blocks;
        n = n_blocks - first < chunk_blocks ? n_blocks - first : chunk_blocks;
    };
    // Everything a chunk needs is enqueued in one go (no host round trip in the middle): inputs up,
    // pipeline kernel, wire packing, then the outputs in the compact format.  The two exception lists
    // travel at their fixed capacity; the counters tell the host how many entries are valid and
    // whether a list overflowed (then that tensor is fetched as int16 after all).
    auto enqueue = [&](int64_t i) -> int {
        int64_t first, n;
        chunk_range(i, first, n);
        const int slot = (int)(i % kSlots);
        cudaStream_t s = ctx->s[slot];
        unsigned char* d = base + (int64_t)slot * L.total;
        Staging& st = ctx->st[slot];
        WireCounters* dcnt = reinterpret_cast<WireCounters*>(d + L.cnt);
        cudaError_t e = cudaMemsetAsync(dcnt, 0, sizeof(WireCounters), s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(counters)");
        NH_CP(d + L.orig, orig + first * nn, n * nn * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.top, top + first * size, n * size * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.left, left + first * size, n * size * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.tr, top_right + first, n * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.bl, bottom_left + first, n * 2, cudaMemcpyHostToDevice, s);
        if (modes) NH_CP(d + L.modes, modes + first, n, cudaMemcpyHostToDevice, s);
        int r = fused_pipeline_dcplanar_narrow(
            reinterpret_cast<int16_t*>(d + L.orig), reinterpret_cast<int16_t*>(d + L.top),
            reinterpret_cast<int16_t*>(d + L.left), reinterpret_cast<int16_t*>(d + L.tr),
            reinterpret_cast<int16_t*>(d + L.bl), modes ? d + L.modes : nullptr, mode, n, size, qp, is_intra,
            use_dst, bit_depth, pred ? reinterpret_cast<int16_t*>(d + L.pred) : nullptr,
            reinterpret_cast<int16_t*>(d + L.coeff16), reinterpret_cast<int16_t*>(d + L.levels16),
            recon ? reinterpret_cast<int16_t*>(d + L.recon) : nullptr, &dcnt->ood, s);
        if (r != NH_OK) return r;
        const int64_t elems = n * nn, segs = (elems + kSeg - 1) / kSeg;
        const int64_t cap = L.cap < segs ? L.cap : segs;  // list entries that travel
        const bool c8 = compact && coeff && s_coeff8, lz = compact && levels && s_levelz;
        fmt_coeff8[slot] = c8;
        fmt_levelz[slot] = lz;
        if (compact && (coeff || levels)) {   // always run: its counters steer the format of the chunks behind this one
            const int64_t pairs = (segs + 1) / 2;
            pack_wire_kernel<<<grid_for(pairs, 8, 8), 256, 0, s>>>(
                coeff ? reinterpret_cast<int16_t*>(d + L.coeff16) : nullptr,
                levels ? reinterpret_cast<int16_t*>(d + L.levels16) : nullptr, elems,
                reinterpret_cast<int8_t*>(d + L.coeff8), reinterpret_cast<int*>(d + L.wide_idx),
                reinterpret_cast<int16_t*>(d + L.wide_val), reinterpret_cast<int*>(d + L.nz_idx),
                reinterpret_cast<int16_t*>(d + L.nz_val), (int)L.cap, dcnt);
            NH_CHECK_LAUNCH("pack_wire_kernel");
        }
        NH_CP(ctx->cnt + slot, dcnt, sizeof(WireCounters), cudaMemcpyDeviceToHost, s);
        if (pred) NH_CP(pred + first * nn, d + L.pred, n * nn * 2, cudaMemcpyDeviceToHost, s);
        if (coeff) {
            if (c8) {
                NH_CP(st.coeff8, d + L.coeff8, elems, cudaMemcpyDeviceToHost, s);
                NH_CP(st.wide_idx, d + L.wide_idx, cap * 4, cudaMemcpyDeviceToHost, s);
                NH_CP(st.wide_val, d + L.wide_val, cap * kSeg * 2, cudaMemcpyDeviceToHost, s);
            } else {
                NH_CP(st.coeff16, d + L.coeff16, elems * 2, cudaMemcpyDeviceToHost, s);
            }
        }
        if (levels) {
            if (lz) {
                NH_CP(st.nz_idx, d + L.nz_idx, cap * 4, cudaMemcpyDeviceToHost, s);
                NH_CP(st.nz_val, d + L.nz_val, cap * kSeg * 2, cudaMemcpyDeviceToHost, s);
            } else {
                NH_CP(st.levels16, d + L.levels16, elems * 2, cudaMemcpyDeviceToHost, s);
            }
        }
        if (recon) NH_CP(recon + first * nn, d + L.recon, n * nn * 2, cudaMemcpyDeviceToHost, s);
        e = cudaEventRecord(ctx->done[slot], s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
        return NH_OK;
    };
    // widen / zero-fill / scatter into the caller's int32 arrays on the host threads
    auto finish = [&](int64_t i) -> int {
        int64_t first, n;
        chunk_range(i, first, n);
        const int slot = (int)(i % kSlots);
        cudaStream_t s = ctx->s[slot];
        const double w0 = trace ? now() : 0;
        cudaError_t e = cudaEventSynchronize(ctx->done[slot]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
        const double w1 = trace ? now() : 0;
        unsigned char* d = base + (int64_t)slot * L.total;
        const WireCounters c = ctx->cnt[slot];
        const Staging& st = ctx->st[slot];
        if (c.ood != 0) {
            // a block left the pixel domain: redo the chunk through the int32 outputs (inputs are still
            // resident in this slot)
            int r = nh_fused_pipeline_dcplanar(
                reinterpret_cast<int16_t*>(d + L.orig), reinterpret_cast<int16_t*>(d + L.top),
                reinterpret_cast<int16_t*>(d + L.left), reinterpret_cast<int16_t*>(d + L.tr),
                reinterpret_cast<int16_t*>(d + L.bl), modes ? d + L.modes : nullptr, mode, n, size, qp,
                is_intra, use_dst, bit_depth, pred ? reinterpret_cast<int16_t*>(d + L.pred) : nullptr,
                coeff ? reinterpret_cast<int32_t*>(d + L.coeff) : nullptr,
                levels ? reinterpret_cast<int32_t*>(d + L.levels) : nullptr,
                recon ? reinterpret_cast<int16_t*>(d + L.recon) : nullptr, s);
            if (r != NH_OK) return r;
            if (pred) NH_CP(pred + first * nn, d + L.pred, n * nn * 2, cudaMemcpyDeviceToHost, s);
            if (coeff) NH_CP(coeff + first * nn, d + L.coeff, n * nn * 4, cudaMemcpyDeviceToHost, s);
            if (levels) NH_CP(levels + first * nn, d + L.levels, n * nn * 4, cudaMemcpyDeviceToHost, s);
            if (recon) NH_CP(recon + first * nn, d + L.recon, n * nn * 2, cudaMemcpyDeviceToHost, s);
            e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
            return NH_OK;
        }
        const size_t elems = (size_t)(n * nn);
        int32_t* dc = coeff ? coeff + first * nn : nullptr;
        int32_t* dl = levels ? levels + first * nn : nullptr;
        if (!dc && !dl) return NH_OK;
        // a list overflowed: that tensor comes over as int16 after all (still resident on the device)
        bool use8 = fmt_coeff8[slot] && dc, usez = fmt_levelz[slot] && dl;
        if (compact) {   // steer the chunks that are enqueued from now on (hysteresis: back to compact below 3/4 of the capacity)
            const int64_t segs_i = ((int64_t)elems + kSeg - 1) / kSeg, cap_i = L.cap < segs_i ? L.cap : segs_i;
            if (dc) s_coeff8 = s_coeff8 ? c.n_wide <= cap_i : 4 * (int64_t)c.n_wide <= 3 * cap_i;
            if (dl) s_levelz = s_levelz ? c.n_nz <= cap_i : 4 * (int64_t)c.n_nz <= 3 * cap_i;
        }
        if (use8 && c.n_wide > L.cap) {
            NH_CP(st.coeff16, d + L.coeff16, elems * 2, cudaMemcpyDeviceToHost, s);
            use8 = false;
        }
        if (usez && c.n_nz > L.cap) {
            NH_CP(st.levels16, d + L.levels16, elems * 2, cudaMemcpyDeviceToHost, s);
            usez = false;
        }
        if ((dc && fmt_coeff8[slot] && !use8) || (dl && fmt_levelz[slot] && !usez)) {   // a guess was wrong: wait for the second transfer
            e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
        }
        workers.run([&](int part, int parts) {
            const size_t per = ((elems + parts - 1) / parts + kSeg - 1) / kSeg * kSeg;  // whole segments
            const size_t lo = (size_t)part * per, hi = lo + per < elems ? lo + per : elems;
            if (lo >= hi) return;
            if (dc) {
                if (use8) widen8(st.coeff8 + lo, dc + lo, hi - lo);
                else widen(st.coeff16 + lo, dc + lo, hi - lo);
            }
            if (dl) {
                if (usez) zero_fill(dl + lo, hi - lo);
                else widen(st.levels16 + lo, dl + lo, hi - lo);
            }
        });
        const double w2 = trace ? now() : 0;
        // the few exception segments, after the bulk pass (every worker has finished)
        auto scatter = [&](const int* idx, const int16_t* val, int count, int32_t* dst) {
            for (int k = 0; k < count; ++k) {
                const size_t e0 = (size_t)idx[k] * kSeg;
                const size_t m = e0 + kSeg <= elems ? kSeg : (e0 < elems ? elems - e0 : 0);
                for (size_t j = 0; j < m; ++j) dst[e0 + j] = val[(size_t)k * kSeg + j];
            }
        };
        if (use8) scatter(st.wide_idx, st.wide_val, c.n_wide, dc);
        if (usez) scatter(st.nz_idx, st.nz_val, c.n_nz, dl);
        if (trace) { t_acc[1] += w1 - w0; t_acc[2] += w2 - w1; t_acc[3] += now() - w2; }
        return NH_OK;
    };

    for (int64_t i = 0; i < n_chunks; ++i) {
        if (i >= kSlots) {
            rc = finish(i - kSlots);
            if (rc != NH_OK) return rc;
        }
        const double t0 = trace ? now() : 0;
        rc = enqueue(i);
        if (rc != NH_OK) return rc;
        if (trace) t_acc[0] += now() - t0;
    }
    for (int64_t i = n_chunks > kSlots ? n_chunks - kSlots : 0; i < n_chunks; ++i) {
        rc = finish(i);
        if (rc != NH_OK) return rc;
    }
    if (trace)
        fprintf(stderr, "[nh_host] chunks %lld: enqueue %.2f ms, wait copies %.2f ms, host widen/fill %.2f ms, "
                "scatter %.2f ms\n", (long long)n_chunks, t_acc[0], t_acc[1], t_acc[2], t_acc[3]);
#undef NH_CP
    drain.armed = false;   // every chunk was finished: the streams are idle
    return NH_OK;
}

// The same pipeline with coefficients and levels DELIVERED as int16 (an option next to the reference's int32
// dtypes): the results then go by DMA straight into the caller's arrays and no host thread touches them -- the
// int32 entry point above writes 12 bytes per pixel of widened results through the host's caches and DRAM, which
// is what bounds it and keeps it from scaling over the GPUs of one box (bench.py e2e: 8 GPUs at 0.21 efficiency
// in round 1).  Valid in the pixel domain (|coeff| <= 32394, |level| <= 13600); a batch
// with a block outside it fails with NH_E_ARG after the transfers have drained and must use the int32 entry.
NH_API int nh_host_pipeline_dcplanar_i16(const int16_t* orig, const int16_t* top, const int16_t* left,
                                         const int16_t* top_right, const int16_t* bottom_left,
                                         const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                                         int is_intra, int use_dst, int bit_depth, int16_t* pred,
                                         int16_t* coeff16, int16_t* levels16, int16_t* recon,
                                         void* device_scratch, int64_t scratch_bytes, int64_t chunk_blocks) {
    if (log2_size(size) < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (n_blocks == 0) return NH_OK;
    if (!orig || !top || !left || !top_right || !bottom_left || n_blocks < 0 || chunk_blocks <= 0) {
        set_error("nh_host_pipeline_dcplanar_i16: null input, negative count or chunk_blocks <= 0");
        return NH_E_ARG;
    }
    if (!modes && mode != 0 && mode != 1) {
        set_error("nh_host_pipeline_dcplanar_i16: mode must be 0 (planar) or 1 (DC), got %d", mode);
        return NH_E_ARG;
    }
    if (bit_depth < 1 || bit_depth > 15) {
        set_error("nh_host_pipeline_dcplanar_i16: bit_depth %d out of range", bit_depth);
        return NH_E_ARG;
    }
    const SlotLayout L = slot_layout(size, chunk_blocks);
    if (!device_scratch || scratch_bytes < kSlots * L.total) {
        set_error("nh_host_pipeline_dcplanar_i16: device scratch of %lld bytes required, got %lld",
                  (long long)(kSlots * L.total), (long long)scratch_bytes);
        return NH_E_NOMEM;
    }
    const int64_t nn = (int64_t)size * size;
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceCtx* ctx = nullptr;
    int rc = get_ctx(0, 0, &ctx);
    if (rc != NH_OK) return rc;
    struct Drain {
        DeviceCtx* c;
        ~Drain() { for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->s[i]); }
    } drain{ctx};
    unsigned char* base = reinterpret_cast<unsigned char*>(device_scratch);
    g_last_h2d = g_last_d2h = 0;
    auto cp = [&](void* dst, const void* src, int64_t bytes, cudaMemcpyKind kind, cudaStream_t s) -> int {
        cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, kind, s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync");
        (kind == cudaMemcpyHostToDevice ? g_last_h2d : g_last_d2h) += bytes;
        return NH_OK;
    };
    // one out-of-domain flag per slot, only ever raised: read once at the end
    for (int slot = 0; slot < kSlots; ++slot) {
        cudaError_t e = cudaMemsetAsync(base + (int64_t)slot * L.total + L.cnt, 0, sizeof(WireCounters), ctx->s[slot]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(counters)");
    }
    const int64_t n_chunks = (n_blocks + chunk_blocks - 1) / chunk_blocks;
    for (int64_t i = 0; i < n_chunks; ++i) {   // a chunk's copies and kernel go back to back on its slot's stream
        const int64_t first = i * chunk_blocks, n = n_blocks - first < chunk_blocks ? n_blocks - first : chunk_blocks;
        const int slot = (int)(i % kSlots);
        cudaStream_t s = ctx->s[slot];
        unsigned char* d = base + (int64_t)slot * L.total;
        WireCounters* dcnt = reinterpret_cast<WireCounters*>(d + L.cnt);
        if ((rc = cp(d + L.orig, orig + first * nn, n * nn * 2, cudaMemcpyHostToDevice, s)) != NH_OK ||
            (rc = cp(d + L.top, top + first * size, n * size * 2, cudaMemcpyHostToDevice, s)) != NH_OK ||
            (rc = cp(d + L.left, left + first * size, n * size * 2, cudaMemcpyHostToDevice, s)) != NH_OK ||
            (rc = cp(d + L.tr, top_right + first, n * 2, cudaMemcpyHostToDevice, s)) != NH_OK ||
            (rc = cp(d + L.bl, bottom_left + first, n * 2, cudaMemcpyHostToDevice, s)) != NH_OK)
            return rc;
        if (modes && (rc = cp(d + L.modes, modes + first, n, cudaMemcpyHostToDevice, s)) != NH_OK) return rc;
        rc = fused_pipeline_dcplanar_narrow(
            reinterpret_cast<int16_t*>(d + L.orig), reinterpret_cast<int16_t*>(d + L.top),
            reinterpret_cast<int16_t*>(d + L.left), reinterpret_cast<int16_t*>(d + L.tr),
            reinterpret_cast<int16_t*>(d + L.bl), modes ? d + L.modes : nullptr, mode, n, size, qp, is_intra,
            use_dst, bit_depth, pred ? reinterpret_cast<int16_t*>(d + L.pred) : nullptr,
            reinterpret_cast<int16_t*>(d + L.coeff16), reinterpret_cast<int16_t*>(d + L.levels16),
            recon ? reinterpret_cast<int16_t*>(d + L.recon) : nullptr, &dcnt->ood, s);
        if (rc != NH_OK) return rc;
        if (pred && (rc = cp(pred + first * nn, d + L.pred, n * nn * 2, cudaMemcpyDeviceToHost, s)) != NH_OK) return rc;
        if (coeff16 && (rc = cp(coeff16 + first * nn, d + L.coeff16, n * nn * 2, cudaMemcpyDeviceToHost, s)) != NH_OK) return rc;
        if (levels16 && (rc = cp(levels16 + first * nn, d + L.levels16, n * nn * 2, cudaMemcpyDeviceToHost, s)) != NH_OK) return rc;
        if (recon && (rc = cp(recon + first * nn, d + L.recon, n * nn * 2, cudaMemcpyDeviceToHost, s)) != NH_OK) return rc;
    }
    for (int slot = 0; slot < kSlots; ++slot)
        if ((rc = cp(ctx->cnt + slot, base + (int64_t)slot * L.total + L.cnt, sizeof(WireCounters), cudaMemcpyDeviceToHost,
                     ctx->s[slot])) != NH_OK)
            return rc;
    for (int slot = 0; slot < kSlots; ++slot) {
        cudaError_t e = cudaStreamSynchronize(ctx->s[slot]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    }
    for (int slot = 0; slot < kSlots; ++slot)
        if (ctx->cnt[slot].ood != 0) {
            set_error("nh_host_pipeline_dcplanar_i16: a block left the pixel domain (samples outside [0, 4095]); its coefficients "
                      "do not fit int16 -- use nh_host_pipeline_dcplanar");
            return NH_E_ARG;
        }
    return NH_OK;
}
