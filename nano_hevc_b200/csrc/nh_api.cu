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

// A stream keeps the slot it was first given.  When every slot is taken, a new stream takes over the slot
// of a stream whose last launch has provably finished (the event recorded behind that launch by
// tile_counter_launched() has completed), one slot at a time: a counter pair that a launch still in
// flight is using is never handed to anyone else.
namespace {
struct CounterSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;   // recorded behind the last launch that used the slot
    bool pending = false;         // acquired, launch not yet recorded
};
struct DeviceCounters {
    void* base = nullptr;         // device address of g_tile_counters
    int used = 0;
    std::map<cudaStream_t, int> by_stream;
    CounterSlot slots[kTileCounters];
};
std::mutex g_counter_mu;
std::map<int, DeviceCounters*> g_counter_devs;
}  // namespace

int acquire_tile_counter(cudaStream_t stream, unsigned int** counter) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> lk(g_counter_mu);
    DeviceCounters*& dc = g_counter_devs[dev];
    if (!dc) dc = new DeviceCounters();
    if (!dc->base) {
        e = cudaGetSymbolAddress(&dc->base, g_tile_counters);
        if (e != cudaSuccess) return cuda_fail(e, "cudaGetSymbolAddress(g_tile_counters)");
    }
    int slot = -1;
    auto it = dc->by_stream.find(stream);
    if (it != dc->by_stream.end()) {
        slot = it->second;
    } else if (dc->used < kTileCounters) {
        slot = dc->used++;
    } else {
        for (int i = 0; i < kTileCounters && slot < 0; ++i) {
            CounterSlot& c = dc->slots[i];
            if (c.pending) continue;
            if (c.done && cudaEventQuery(c.done) != cudaSuccess) {   // still running (or captured): leave it alone
                cudaGetLastError();
                continue;
            }
            dc->by_stream.erase(c.stream);
            slot = i;
        }
        if (slot < 0) {
            set_error("all %d work-counter slots of device %d belong to streams with launches in flight", kTileCounters, dev);
            return NH_E_CUDA;
        }
    }
    dc->by_stream[stream] = slot;
    dc->slots[slot].stream = stream;
    dc->slots[slot].pending = true;
    *counter = reinterpret_cast<unsigned int*>(dc->base) + 4 * slot;
    return NH_OK;
}

void tile_counter_launched(cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lk(g_counter_mu);
    auto d = g_counter_devs.find(dev);
    if (d == g_counter_devs.end()) return;
    auto it = d->second->by_stream.find(stream);
    if (it == d->second->by_stream.end()) return;
    CounterSlot& c = d->second->slots[it->second];
    if (!c.done && cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        c.done = nullptr;
        return;   // without an event the slot simply stays with its stream (pending is never cleared)
    }
    if (cudaEventRecord(c.done, stream) == cudaSuccess) c.pending = false;
    else cudaGetLastError();
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
