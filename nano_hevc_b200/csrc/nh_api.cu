// nh_api.cu -- meta entry points, error plumbing and host-side tables of
// libnh_b200.so (see include/nh_b200.h).
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <map>
#include <set>
#include <utility>

#include "nh_common.cuh"

namespace nh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return NH_E_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

int ensure_dynamic_smem_impl(const void* kernel, int bytes, const char* what) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({kernel, dev})) return NH_OK;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, what);
    done.insert({kernel, dev});
    return NH_OK;
}

constexpr int kTileCounters = 4096;
__device__ unsigned int g_tile_counters[4 * kTileCounters];  // {ticket, finished warps, undecided tiles, readers} per slot, zero at load

int acquire_tile_counter(cudaStream_t stream, unsigned int** counter) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, int> slots;  // (device, stream) -> slot
    static int next_slot[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    int slot;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = slots.find({dev, stream});
        if (it == slots.end()) {
            int& n = next_slot[dev & 63];
            if (n >= kTileCounters) {  // more distinct streams than counters: recycle (streams come and go)
                n = 0;
                for (auto i = slots.begin(); i != slots.end();)
                    i = i->first.first == dev ? slots.erase(i) : std::next(i);
            }
            it = slots.emplace(std::make_pair(dev, stream), n++).first;
        }
        slot = it->second;
    }
    static void* bases[64] = {};  // device address of g_tile_counters, per device
    void* base = bases[dev & 63];
    if (!base) {
        e = cudaGetSymbolAddress(&base, g_tile_counters);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetSymbolAddress(g_tile_counters)");
        bases[dev & 63] = base;
    }
    *counter = reinterpret_cast<unsigned int*>(base) + 4 * slot;
    return NH_OK;
}

}  // namespace nh

NH_API int nh_version(void) { return 100; }

NH_API const char* nh_last_error(void) { return nh::g_err; }

NH_API int nh_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        nh::set_error("no CUDA device visible");
        return 0;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        nh::set_error("device %d has compute capability %d.x; this library is built for sm_100a only",
                      dev, major);
        return 0;
    }
    return 1;
}

// nano_hevc/transform.py:20-135
NH_API int nh_get_transform_matrix(int size, int use_dst, int32_t* out) {
    if (!out) { nh::set_error("nh_get_transform_matrix: null output"); return NH_E_ARG; }
    if (nh::log2_size(size) < 0) {
        nh::set_error("Unsupported transform size: %d", size);
        return NH_E_SIZE;
    }
    for (int i = 0; i < size; ++i)
        for (int j = 0; j < size; ++j) {
            int v;
            if (use_dst && size == 4) v = nh::dst4(i, j);       // transform.py:140-141
            else v = nh::cosv((i * (32 / size)) * (2 * j + 1));  // one quarter-wave family
            out[i * size + j] = v;
        }
    return NH_OK;
}

// nano_hevc/intra.py:24-29
NH_API int nh_get_intra_pred_angle(int mode, int* angle_out) {
    if (mode < 2 || mode > 34 || !angle_out) {
        nh::set_error("nh_get_intra_pred_angle: mode %d out of range 2..34", mode);
        return NH_E_ARG;
    }
    *angle_out = nh::intra_angle(mode);
    return NH_OK;
}

// nano_hevc/quant.py:25-38
NH_API int nh_get_qp_params(int qp, int* per_out, int* rem_out) {
    if (!per_out || !rem_out) { nh::set_error("nh_get_qp_params: null output"); return NH_E_ARG; }
    qp = qp < 0 ? 0 : (qp > 51 ? 51 : qp);
    *per_out = qp / 6;
    *rem_out = qp % 6;
    return NH_OK;
}

// nano_hevc/quant.py:21-22
NH_API int nh_get_quant_scales(int rem, int* q, int* dq) {
    if (rem < 0 || rem > 5 || !q || !dq) {
        nh::set_error("nh_get_quant_scales: rem %d out of range 0..5", rem);
        return NH_E_ARG;
    }
    nh::QuantParams p = nh::make_quant_params(rem, 2, 1);
    *q = p.mf;
    *dq = p.scale;
    return NH_OK;
}
