// nh_host.cu -- host-buffer entry point of the K6 pipeline (the end-to-end path a caller with
// numpy / C arrays uses).  PCIe is the bound here (14.56 B/px would have to cross it), so:
//   * the batch is cut into chunks that rotate over three streams: H2D of chunk i+1, the kernel of
//     chunk i and D2H of chunk i-1 overlap (B200 has separate copy engines per direction);
//   * coefficients and levels cross the bus as int16 (in the pixel domain |coeff| <= 32394 and
//     |level| <= 13600) into pinned staging buffers and are widened to the
//     caller's int32 arrays by a small pool of host threads (AVX2 sign-extension + streaming
//     stores) while the next chunks are in flight: 8 instead of 12 output bytes per pixel;
//   * a chunk in which any block left the pixel domain raises a device flag and is redone through
//     the plain int32 path, so the result is bit-exact for every input.
#include <immintrin.h>

#include <atomic>
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

// ------------------------------------------------------------------ per-device context
struct SlotLayout {
    int64_t orig, top, left, tr, bl, modes, pred, coeff, levels, recon, coeff16, levels16, flag, total;
};

static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

static SlotLayout slot_layout(int size, int64_t chunk) {
    const int64_t nn = (int64_t)size * size;
    SlotLayout l;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += align256(bytes); return o; };
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
    l.flag = take(4);
    l.total = off;
    return l;
}

struct DeviceCtx {
    cudaStream_t s[kSlots];
    cudaEvent_t done[kSlots];
    int16_t* stage_coeff[kSlots] = {nullptr, nullptr, nullptr};   // pinned host staging
    int16_t* stage_levels[kSlots] = {nullptr, nullptr, nullptr};
    int* stage_flag = nullptr;                                     // kSlots ints, pinned
    int64_t stage_elems = 0;
    bool ready = false;
};

static std::mutex g_mu;
static DeviceCtx g_ctx[64];
static Pool* g_pool = nullptr;

static Pool& pool() {
    if (!g_pool) {
        int n = 0;
        if (const char* e = getenv("NH_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            n = (int)std::thread::hardware_concurrency();
            if (n > 8) n = 8;
        }
        g_pool = new Pool(n);
    }
    return *g_pool;
}

static int get_ctx(int64_t stage_elems, DeviceCtx** out) {
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
        e = cudaHostAlloc(reinterpret_cast<void**>(&c.stage_flag), kSlots * sizeof(int), cudaHostAllocDefault);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc(flags)");
        c.ready = true;
    }
    if (c.stage_elems < stage_elems) {
        for (int i = 0; i < kSlots; ++i) {
            if (c.stage_coeff[i]) cudaFreeHost(c.stage_coeff[i]);
            if (c.stage_levels[i]) cudaFreeHost(c.stage_levels[i]);
            e = cudaHostAlloc(reinterpret_cast<void**>(&c.stage_coeff[i]), stage_elems * 2, cudaHostAllocDefault);
            if (e == cudaSuccess)
                e = cudaHostAlloc(reinterpret_cast<void**>(&c.stage_levels[i]), stage_elems * 2, cudaHostAllocDefault);
            if (e != cudaSuccess) { c.stage_elems = 0; return cuda_fail(e, "cudaHostAlloc(staging)"); }
        }
        c.stage_elems = stage_elems;
    }
    *out = &c;
    return NH_OK;
}

}  // namespace nh

using namespace nh;

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
    std::lock_guard<std::mutex> lk(g_mu);  // one host pipeline at a time per process (shared staging)
    DeviceCtx* ctx = nullptr;
    int rc = get_ctx(chunk_blocks * nn, &ctx);
    if (rc != NH_OK) return rc;
    Pool& workers = pool();
    unsigned char* base = reinterpret_cast<unsigned char*>(device_scratch);
    const int64_t n_chunks = (n_blocks + chunk_blocks - 1) / chunk_blocks;

#define NH_CP(dst, src, bytes, kind, s)                                        \
    do {                                                                       \
        cudaError_t e__ = cudaMemcpyAsync(dst, src, (size_t)(bytes), kind, s); \
        if (e__ != cudaSuccess) return cuda_fail(e__, "cudaMemcpyAsync");      \
    } while (0)

    auto chunk_range = [&](int64_t i, int64_t& first, int64_t& n) {
        first = i * chunk_blocks;
        n = n_blocks - first < chunk_blocks ? n_blocks - first : chunk_blocks;
    };
    auto enqueue = [&](int64_t i) -> int {
        int64_t first, n;
        chunk_range(i, first, n);
        const int slot = (int)(i % kSlots);
        cudaStream_t s = ctx->s[slot];
        unsigned char* d = base + (int64_t)slot * L.total;
        cudaError_t e = cudaMemsetAsync(d + L.flag, 0, 4, s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(flag)");
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
            recon ? reinterpret_cast<int16_t*>(d + L.recon) : nullptr, reinterpret_cast<int*>(d + L.flag), s);
        if (r != NH_OK) return r;
        NH_CP(ctx->stage_flag + slot, d + L.flag, 4, cudaMemcpyDeviceToHost, s);
        if (pred) NH_CP(pred + first * nn, d + L.pred, n * nn * 2, cudaMemcpyDeviceToHost, s);
        if (coeff) NH_CP(ctx->stage_coeff[slot], d + L.coeff16, n * nn * 2, cudaMemcpyDeviceToHost, s);
        if (levels) NH_CP(ctx->stage_levels[slot], d + L.levels16, n * nn * 2, cudaMemcpyDeviceToHost, s);
        if (recon) NH_CP(recon + first * nn, d + L.recon, n * nn * 2, cudaMemcpyDeviceToHost, s);
        e = cudaEventRecord(ctx->done[slot], s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
        return NH_OK;
    };
    auto finish = [&](int64_t i) -> int {
        int64_t first, n;
        chunk_range(i, first, n);
        const int slot = (int)(i % kSlots);
        cudaStream_t s = ctx->s[slot];
        cudaError_t e = cudaEventSynchronize(ctx->done[slot]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
        unsigned char* d = base + (int64_t)slot * L.total;
        if (ctx->stage_flag[slot] != 0) {
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
        const int16_t* sc = ctx->stage_coeff[slot];
        const int16_t* sl = ctx->stage_levels[slot];
        int32_t* dc = coeff ? coeff + first * nn : nullptr;
        int32_t* dl = levels ? levels + first * nn : nullptr;
        if (dc || dl) {
            workers.run([&](int part, int parts) {
                const size_t per = ((elems + parts - 1) / parts + 15) / 16 * 16;
                const size_t lo = (size_t)part * per, hi = lo + per < elems ? lo + per : elems;
                if (lo >= hi) return;
                if (dc) widen(sc + lo, dc + lo, hi - lo);
                if (dl) widen(sl + lo, dl + lo, hi - lo);
            });
        }
        return NH_OK;
    };

    for (int64_t i = 0; i < n_chunks; ++i) {
        if (i >= kSlots) {
            rc = finish(i - kSlots);
            if (rc != NH_OK) return rc;
        }
        rc = enqueue(i);
        if (rc != NH_OK) return rc;
    }
    for (int64_t i = n_chunks > kSlots ? n_chunks - kSlots : 0; i < n_chunks; ++i) {
        rc = finish(i);
        if (rc != NH_OK) return rc;
    }
#undef NH_CP
    return NH_OK;
}
