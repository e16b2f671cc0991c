// nh_host_frames.cu -- host-buffer entry point of the frame coders (K7 / K8): the end-to-end path of BASELINE
// configs 3 and 5 for a caller whose frames and results live in host memory (numpy arrays, C buffers).
//
// The batch is cut into chunks of `frames_per_chunk` frames that rotate over three internal streams: the upload
// of chunk i+1, nh_encode_frames of chunk i and the download of chunk i-1 overlap (separate copy engines per
// direction).  A chunk's copies and launches are enqueued back to back on ITS stream, so the stream order alone
// keeps slot reuse safe -- chunk i+3 cannot start its upload before chunk i has finished its download -- and no
// event is needed.  Results are delivered in the reference's dtypes straight into the caller's arrays: with every
// output requested that is 12 + 5/N^2 bytes per pixel down against 2 up, so PCIe bounds this path (4-5 Gpix/s
// per GPU on a Gen5 x16 link); outputs passed as NULL are neither computed for delivery nor transferred.  Pinned
// (page-locked) caller buffers give asynchronous copies at link speed; pageable ones work, the driver stages them.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "nh_common.cuh"

namespace nh {

constexpr int kFrameSlots = 3;

struct FrameSlot {   // byte offsets of one slot inside the caller's device scratch
    int64_t src, recon, modes, costs, pred, coeff, levels, stats, wave, total;
};

static int64_t up256(int64_t v) { return (v + 255) / 256 * 256; }

static FrameSlot frame_slot(int fc, int H, int W, int size, int recon_neighbours) {
    const int64_t px = (int64_t)fc * H * W, blocks = (int64_t)fc * (H / size) * (W / size), nn = (int64_t)size * size;
    FrameSlot l;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { int64_t o = off; off += up256(bytes); return o; };
    l.src = take(px * 2);
    l.recon = take(px * 2);
    l.modes = take(blocks);
    l.costs = take(blocks * 4);
    l.pred = take(blocks * nn * 2);
    l.coeff = take(blocks * nn * 4);
    l.levels = take(blocks * nn * 4);
    l.stats = take((int64_t)fc * 4 * 8);
    l.wave = take(recon_neighbours ? nh_encode_frames_scratch_bytes(fc, H, W, size) : 16);
    l.total = off;
    return l;
}

struct FrameCtx {
    cudaStream_t s[kFrameSlots];
    bool ready = false;
};
static std::mutex g_frames_mu;
static FrameCtx g_frames_ctx[64];
static int64_t g_frames_h2d = 0, g_frames_d2h = 0;

static int frames_ctx(FrameCtx** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return NH_E_ARG; }
    FrameCtx& c = g_frames_ctx[dev];
    if (!c.ready) {
        for (int i = 0; i < kFrameSlots; ++i) {
            e = cudaStreamCreateWithFlags(&c.s[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithFlags");
        }
        c.ready = true;
    }
    *out = &c;
    return NH_OK;
}

}  // namespace nh

using namespace nh;

NH_API int64_t nh_host_encode_frames_scratch_bytes(int frames_per_chunk, int height, int width, int size,
                                                   int recon_neighbours) {
    if (log2_size(size) < 0 || frames_per_chunk < 1 || height < 0 || width < 0) return 0;
    return kFrameSlots * frame_slot(frames_per_chunk, height, width, size, recon_neighbours).total;
}

NH_API int nh_host_encode_frames_last_transfer(int64_t* h2d_bytes, int64_t* d2h_bytes) {
    std::lock_guard<std::mutex> lk(g_frames_mu);
    if (h2d_bytes) *h2d_bytes = g_frames_h2d;
    if (d2h_bytes) *d2h_bytes = g_frames_d2h;
    return NH_OK;
}

NH_API int nh_host_encode_frames(const int16_t* src, int n_frames, int height, int width, int size, int cost_kind,
                                 int qp, int recon_neighbours, int bit_depth, uint8_t* modes, int32_t* costs,
                                 int16_t* pred, int32_t* coeff, int32_t* levels, int16_t* recon_planes,
                                 int64_t* stats, int frames_per_chunk, void* device_scratch, int64_t scratch_bytes) {
    if (log2_size(size) < 0) { set_error("Unsupported transform size: %d", size); return NH_E_SIZE; }
    if (!src || n_frames < 0 || height < 0 || width < 0 || frames_per_chunk < 1) {
        set_error("nh_host_encode_frames: bad argument (null src, negative shape or frames_per_chunk < 1)");
        return NH_E_ARG;
    }
    if (n_frames == 0) return NH_OK;
    const int fc = frames_per_chunk < n_frames ? frames_per_chunk : n_frames;
    const FrameSlot l = frame_slot(fc, height, width, size, recon_neighbours);
    if (!device_scratch || scratch_bytes < kFrameSlots * l.total) {
        set_error("nh_host_encode_frames: device scratch of %lld bytes required, got %lld",
                  (long long)(kFrameSlots * l.total), (long long)scratch_bytes);
        return NH_E_NOMEM;
    }
    if ((reinterpret_cast<uintptr_t>(device_scratch) & 255) != 0) {
        set_error("nh_host_encode_frames: device scratch must be 256-byte aligned");
        return NH_E_ARG;
    }
    std::lock_guard<std::mutex> lk(g_frames_mu);   // one call at a time per process: the slots' streams are shared
    FrameCtx* ctx = nullptr;
    int rc = frames_ctx(&ctx);
    if (rc != NH_OK) return rc;
    const int64_t plane = (int64_t)height * width, bpf = (int64_t)(height / size) * (width / size), nn = (int64_t)size * size;
    int64_t up = 0, down = 0;
    auto drain = [&]() { for (int i = 0; i < kFrameSlots; ++i) cudaStreamSynchronize(ctx->s[i]); };
    auto fail = [&](int code) { drain(); return code; };   // nothing of ours may still touch the caller's buffers
    int chunk = 0;
    for (int f0 = 0; f0 < n_frames; f0 += fc, ++chunk) {
        const int nf = n_frames - f0 < fc ? n_frames - f0 : fc;
        cudaStream_t st = ctx->s[chunk % kFrameSlots];
        unsigned char* base = reinterpret_cast<unsigned char*>(device_scratch) + (int64_t)(chunk % kFrameSlots) * l.total;
        auto dp = [&](int64_t off) { return base + off; };
        cudaError_t e = cudaMemcpyAsync(dp(l.src), src + (int64_t)f0 * plane, (size_t)nf * plane * 2, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMemcpyAsync(src)"));
        up += (int64_t)nf * plane * 2;
        rc = nh_encode_frames(reinterpret_cast<const int16_t*>(dp(l.src)), nf, plane, height, width, width, size, cost_kind, qp,
                              recon_neighbours, bit_depth, dp(l.modes),   // always: the search + winner kernels need the modes tensor
                              (costs || stats) ? reinterpret_cast<int32_t*>(dp(l.costs)) : nullptr,
                              pred ? reinterpret_cast<int16_t*>(dp(l.pred)) : nullptr,
                              coeff ? reinterpret_cast<int32_t*>(dp(l.coeff)) : nullptr,
                              (levels || stats) ? reinterpret_cast<int32_t*>(dp(l.levels)) : nullptr,
                              reinterpret_cast<int16_t*>(dp(l.recon)),   // always coded (neighbours, statistics); delivered on request
                              stats ? reinterpret_cast<int64_t*>(dp(l.stats)) : nullptr, dp(l.wave),
                              recon_neighbours ? nh_encode_frames_scratch_bytes(nf, height, width, size) : 16, st);
        if (rc != NH_OK) return fail(rc);
        auto fetch = [&](void* host, int64_t off, int64_t bytes, const char* what) -> bool {
            if (!host || bytes == 0) return true;
            e = cudaMemcpyAsync(host, dp(off), (size_t)bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) { rc = cuda_fail(e, what); return false; }
            down += bytes;
            return true;
        };
        const int64_t b0 = (int64_t)f0 * bpf, nb = (int64_t)nf * bpf;
        if (!fetch(modes ? modes + b0 : nullptr, l.modes, nb, "cudaMemcpyAsync(modes)") ||
            !fetch(costs ? costs + b0 : nullptr, l.costs, nb * 4, "cudaMemcpyAsync(costs)") ||
            !fetch(pred ? pred + b0 * nn : nullptr, l.pred, nb * nn * 2, "cudaMemcpyAsync(pred)") ||
            !fetch(coeff ? coeff + b0 * nn : nullptr, l.coeff, nb * nn * 4, "cudaMemcpyAsync(coeff)") ||
            !fetch(levels ? levels + b0 * nn : nullptr, l.levels, nb * nn * 4, "cudaMemcpyAsync(levels)") ||
            !fetch(recon_planes ? recon_planes + (int64_t)f0 * plane : nullptr, l.recon, (int64_t)nf * plane * 2, "cudaMemcpyAsync(recon)") ||
            !fetch(stats ? stats + (int64_t)f0 * 4 : nullptr, l.stats, (int64_t)nf * 4 * 8, "cudaMemcpyAsync(stats)"))
            return fail(rc);
    }
    for (int i = 0; i < kFrameSlots; ++i) {
        cudaError_t e = cudaStreamSynchronize(ctx->s[i]);
        if (e != cudaSuccess) return fail(cuda_fail(e, "cudaStreamSynchronize"));
    }
    g_frames_h2d = up;
    g_frames_d2h = down;
    return NH_OK;
}
