// nh_host.cu -- host-buffer entry point of the K6 pipeline (the end-to-end path a
// caller with numpy / C arrays uses).  The batch is cut into chunks; each chunk is
// copied in, coded and copied out on one of three internal streams so that the
// H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1
// overlap (B200 has separate copy engines per direction).
#include <mutex>

#include "nh_common.cuh"

namespace nh {

constexpr int kSlots = 3;

struct SlotLayout {
    int64_t orig, top, left, tr, bl, modes, pred, coeff, levels, recon, total;
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
    l.coeff = take(chunk * nn * 4);
    l.levels = take(chunk * nn * 4);
    l.recon = take(chunk * nn * 2);
    l.total = off;
    return l;
}

struct DeviceStreams {
    cudaStream_t s[kSlots];
    bool ready = false;
};

static std::mutex g_mu;
static DeviceStreams g_streams[64];

static int get_streams(cudaStream_t** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return NH_E_ARG; }
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceStreams& ds = g_streams[dev];
    if (!ds.ready) {
        for (int i = 0; i < kSlots; ++i) {
            e = cudaStreamCreateWithFlags(&ds.s[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithFlags");
        }
        ds.ready = true;
    }
    *out = ds.s;
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
    if (!orig || !top || !left || !top_right || !bottom_left || n_blocks < 0 || chunk_blocks <= 0) {
        set_error("nh_host_pipeline_dcplanar: null input, negative count or chunk_blocks <= 0");
        return NH_E_ARG;
    }
    const SlotLayout L = slot_layout(size, chunk_blocks);
    if (!device_scratch || scratch_bytes < kSlots * L.total) {
        set_error("nh_host_pipeline_dcplanar: device scratch of %lld bytes required, got %lld",
                  (long long)(kSlots * L.total), (long long)scratch_bytes);
        return NH_E_NOMEM;
    }
    if (n_blocks == 0) return NH_OK;
    cudaStream_t* st = nullptr;
    int rc = get_streams(&st);
    if (rc != NH_OK) return rc;
    const int64_t nn = (int64_t)size * size;
    unsigned char* base = reinterpret_cast<unsigned char*>(device_scratch);
    int64_t done = 0;
    int chunk_idx = 0;
#define NH_CP(dst, src, bytes, kind, s)                                        \
    do {                                                                       \
        cudaError_t e__ = cudaMemcpyAsync(dst, src, (size_t)(bytes), kind, s); \
        if (e__ != cudaSuccess) return cuda_fail(e__, "cudaMemcpyAsync");      \
    } while (0)
    while (done < n_blocks) {
        const int64_t n = n_blocks - done < chunk_blocks ? n_blocks - done : chunk_blocks;
        const int slot = chunk_idx % kSlots;
        cudaStream_t s = st[slot];
        unsigned char* d = base + (int64_t)slot * L.total;
        NH_CP(d + L.orig, orig + done * nn, n * nn * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.top, top + done * size, n * size * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.left, left + done * size, n * size * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.tr, top_right + done, n * 2, cudaMemcpyHostToDevice, s);
        NH_CP(d + L.bl, bottom_left + done, n * 2, cudaMemcpyHostToDevice, s);
        if (modes) NH_CP(d + L.modes, modes + done, n, cudaMemcpyHostToDevice, s);
        rc = nh_fused_pipeline_dcplanar(
            reinterpret_cast<int16_t*>(d + L.orig), reinterpret_cast<int16_t*>(d + L.top),
            reinterpret_cast<int16_t*>(d + L.left), reinterpret_cast<int16_t*>(d + L.tr),
            reinterpret_cast<int16_t*>(d + L.bl), modes ? d + L.modes : nullptr, mode, n, size, qp,
            is_intra, use_dst, bit_depth, pred ? reinterpret_cast<int16_t*>(d + L.pred) : nullptr,
            coeff ? reinterpret_cast<int32_t*>(d + L.coeff) : nullptr,
            levels ? reinterpret_cast<int32_t*>(d + L.levels) : nullptr,
            recon ? reinterpret_cast<int16_t*>(d + L.recon) : nullptr, s);
        if (rc != NH_OK) return rc;
        if (pred) NH_CP(pred + done * nn, d + L.pred, n * nn * 2, cudaMemcpyDeviceToHost, s);
        if (coeff) NH_CP(coeff + done * nn, d + L.coeff, n * nn * 4, cudaMemcpyDeviceToHost, s);
        if (levels) NH_CP(levels + done * nn, d + L.levels, n * nn * 4, cudaMemcpyDeviceToHost, s);
        if (recon) NH_CP(recon + done * nn, d + L.recon, n * nn * 2, cudaMemcpyDeviceToHost, s);
        done += n;
        ++chunk_idx;
    }
#undef NH_CP
    for (int i = 0; i < kSlots; ++i) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    }
    return NH_OK;
}
