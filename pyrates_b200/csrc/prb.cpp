// libprb — C ABI for JIT-compiling and launching generated PyRates vector-field kernels on B200.
// See include/prb.h for the contract and the reference interfaces each entry point replaces.
//
// Design notes
//  * NVRTC (linked) turns generated CUDA C into an sm_100a cubin; this needs no GPU, so the build
//    container can compile and inspect every kernel.
//  * The CUDA *driver* API is resolved at run time with dlopen("libcuda.so.1"): the library loads on
//    machines without a driver (CPU CI) and fails loudly — PRB_ERR_NO_DEVICE — on any compute call.
//    There is deliberately no CPU execution path.
//  * Kernels take one by-value prb_kernel_args block; cooperative launches are used for the
//    persistent network kernels (grid-wide barrier between rhs evaluations).
#include "../../include/prb.h"

#include <cuda.h>
#include <nvrtc.h>

#include <dlfcn.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// ---------------------------------------------------------------- driver API (lazy dlopen)
#define PRB_DRV_FUNCS(X)                                                                          \
    X(cuInit) X(cuDeviceGetCount) X(cuDeviceGet) X(cuDeviceGetName) X(cuDeviceGetAttribute)       \
    X(cuDeviceTotalMem_v2) X(cuDevicePrimaryCtxRetain) X(cuCtxSetCurrent) X(cuCtxGetCurrent)      \
    X(cuCtxSynchronize) X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction)              \
    X(cuFuncGetAttribute) X(cuFuncSetAttribute) X(cuOccupancyMaxActiveBlocksPerMultiprocessor)    \
    X(cuLaunchKernel) X(cuLaunchCooperativeKernel) X(cuMemAlloc_v2) X(cuMemFree_v2)               \
    X(cuMemsetD8Async) X(cuMemcpyHtoDAsync_v2) X(cuMemcpyDtoHAsync_v2) X(cuMemcpyDtoDAsync_v2)    \
    X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemHostRegister_v2) X(cuMemHostUnregister) X(cuStreamCreate) X(cuStreamDestroy_v2)                    \
    X(cuStreamSynchronize) X(cuStreamWaitEvent) X(cuEventCreate) X(cuEventDestroy_v2) X(cuEventRecord)                 \
    X(cuEventSynchronize) X(cuEventElapsedTime) X(cuGetErrorString) X(cuGetErrorName)             \
    X(cuIpcGetMemHandle) X(cuIpcOpenMemHandle_v2) X(cuIpcCloseMemHandle)

struct Driver {
    void* lib = nullptr;
    bool ok = false;
    std::string why;
#define X(name) decltype(&::name) name = nullptr;
    PRB_DRV_FUNCS(X)
#undef X
};

Driver g_drv;
std::once_flag g_drv_once;
std::mutex g_ctx_mu;
CUcontext g_ctx = nullptr;
int g_dev = -1;
std::map<int, CUcontext> g_ctx_of;     // retained primary contexts, one per device ever selected

void load_driver() {
    g_drv.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!g_drv.lib) g_drv.lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!g_drv.lib) {
        const char* why = dlerror();
        g_drv.why = std::string("cannot load libcuda.so.1: ") + (why ? why : "unknown dlopen error");
        return;
    }
    // Some driver symbols carry versioned names; CUDA's header maps e.g. cuMemAlloc -> cuMemAlloc_v2.
#define X(name)                                                                      \
    g_drv.name = reinterpret_cast<decltype(&::name)>(dlsym(g_drv.lib, #name));       \
    if (!g_drv.name) { g_drv.why = "libcuda lacks symbol " #name; return; }
    PRB_DRV_FUNCS(X)
#undef X
    CUresult r = g_drv.cuInit(0);
    if (r != CUDA_SUCCESS) {
        g_drv.why = "cuInit failed with code " + std::to_string((int)r);
        return;
    }
    g_drv.ok = true;
}

int need_driver() {
    std::call_once(g_drv_once, load_driver);
    if (!g_drv.ok)
        return fail(PRB_ERR_NO_DEVICE, "CUDA driver unavailable (%s); libprb has no CPU execution path",
                    g_drv.why.c_str());
    return PRB_OK;
}

int cu_fail(CUresult r, const char* what) {
    const char *name = nullptr, *msg = nullptr;
    if (g_drv.ok) {
        g_drv.cuGetErrorName(r, &name);
        g_drv.cuGetErrorString(r, &msg);
    }
    return fail(PRB_ERR_CUDA, "%s: %s (%s)", what, name ? name : "CUDA_ERROR", msg ? msg : "?");
}

#define CU(call)                                                  \
    do {                                                          \
        CUresult _r = g_drv.call;                                 \
        if (_r != CUDA_SUCCESS) return cu_fail(_r, #call);        \
    } while (0)

int need_ctx() {
    if (int e = need_driver()) return e;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!g_ctx) {
        CUdevice dev;
        CU(cuDeviceGet(&dev, 0));
        CU(cuDevicePrimaryCtxRetain(&g_ctx, dev));
        g_dev = 0;
        g_ctx_of[0] = g_ctx;
    }
    CUcontext cur = nullptr;
    g_drv.cuCtxGetCurrent(&cur);
    if (cur != g_ctx) CU(cuCtxSetCurrent(g_ctx));
    return PRB_OK;
}

}  // namespace

// A module is a cubin; it is loaded lazily into the context of whichever device is current at launch time, once per
// device (CudaBackend(device=0) and CudaBackend(device=1) may share one compiled model inside a process).
struct prb_loaded {
    CUmodule mod = nullptr;
    std::map<std::string, CUfunction> funcs;
};
struct prb_module {
    std::string name;
    std::string log;
    std::vector<char> cubin;
    std::map<int, prb_loaded> on;      // device ordinal -> loaded image
    std::mutex mu;
};

namespace {

int module_function(prb_module* m, const char* kernel, CUfunction* out) {
    std::lock_guard<std::mutex> lk(m->mu);
    prb_loaded& L = m->on[g_dev];
    if (!L.mod) {
        if (m->cubin.empty()) return fail(PRB_ERR_INVALID, "module `%s` holds no cubin", m->name.c_str());
        CU(cuModuleLoadData(&L.mod, m->cubin.data()));
    }
    auto it = L.funcs.find(kernel);
    if (it == L.funcs.end()) {
        CUfunction f;
        CUresult r = g_drv.cuModuleGetFunction(&f, L.mod, kernel);
        if (r != CUDA_SUCCESS)
            return fail(PRB_ERR_NOT_FOUND, "kernel `%s` not found in module `%s`", kernel, m->name.c_str());
        it = L.funcs.emplace(kernel, f).first;
    }
    *out = it->second;
    return PRB_OK;
}

}  // namespace

extern "C" {

int prb_abi_version(void) { return PRB_ABI_VERSION; }
const char* prb_last_error(void) { return g_err.c_str(); }
int64_t prb_launch_count(void) { return g_launches.load(); }
int prb_nvrtc_version(int* major, int* minor) {
    if (!major || !minor) return fail(PRB_ERR_INVALID, "NULL argument");
    return nvrtcVersion(major, minor) == NVRTC_SUCCESS ? PRB_OK : fail(PRB_ERR_COMPILE, "nvrtcVersion failed");
}

// ---------------------------------------------------------------- device
int prb_device_count(int* count) {
    if (!count) return fail(PRB_ERR_INVALID, "count is NULL");
    if (int e = need_driver()) return e;
    CU(cuDeviceGetCount(count));
    return PRB_OK;
}

int prb_set_device(int ordinal) {
    if (int e = need_driver()) return e;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx && g_dev == ordinal) return g_drv.cuCtxSetCurrent(g_ctx) == CUDA_SUCCESS ? PRB_OK : fail(PRB_ERR_CUDA, "cuCtxSetCurrent");
    CUcontext ctx = nullptr;
    auto it = g_ctx_of.find(ordinal);
    if (it != g_ctx_of.end()) {
        ctx = it->second;                       // retained once per device: switching back and forth leaks nothing
    } else {
        CUdevice dev;
        CU(cuDeviceGet(&dev, ordinal));
        CU(cuDevicePrimaryCtxRetain(&ctx, dev));
        g_ctx_of[ordinal] = ctx;
    }
    CU(cuCtxSetCurrent(ctx));
    g_ctx = ctx;
    g_dev = ordinal;
    return PRB_OK;
}

int prb_get_device_info(prb_device_info* info) {
    if (!info) return fail(PRB_ERR_INVALID, "info is NULL");
    if (int e = need_ctx()) return e;
    CUdevice dev;
    CU(cuDeviceGet(&dev, g_dev));
    memset(info, 0, sizeof *info);
    CU(cuDeviceGetName(info->name, sizeof info->name, dev));
    CU(cuDeviceGetAttribute(&info->cc_major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev));
    CU(cuDeviceGetAttribute(&info->cc_minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev));
    CU(cuDeviceGetAttribute(&info->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    CU(cuDeviceGetAttribute(&info->max_smem_per_block_optin, CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN, dev));
    CU(cuDeviceGetAttribute(&info->l2_bytes, CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE, dev));
    CU(cuDeviceGetAttribute(&info->clock_khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE, dev));
    size_t total = 0;
    CU(cuDeviceTotalMem_v2(&total, dev));
    info->total_mem = (int64_t)total;
    return PRB_OK;
}

int prb_device_synchronize(void) {
    if (int e = need_ctx()) return e;
    CU(cuCtxSynchronize());
    return PRB_OK;
}

// ---------------------------------------------------------------- JIT
int prb_compile(const char* src, const char* name, const char* arch, const char* const* opts, int n_opts,
                prb_module** out) {
    if (!src || !out) return fail(PRB_ERR_INVALID, "src/out is NULL");
    *out = nullptr;
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, src, name ? name : "prb_kernel.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return fail(PRB_ERR_COMPILE, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::vector<std::string> o;
    o.push_back(std::string("--gpu-architecture=") + (arch && *arch ? arch : "sm_100a"));
    for (int i = 0; i < n_opts; ++i)
        if (opts && opts[i]) o.push_back(opts[i]);
    std::vector<const char*> oc;
    for (auto& s : o) oc.push_back(s.c_str());
    r = nvrtcCompileProgram(prog, (int)oc.size(), oc.data());
    auto* m = new prb_module();
    m->name = name ? name : "prb_kernel.cu";
    size_t n = 0;
    if (nvrtcGetProgramLogSize(prog, &n) == NVRTC_SUCCESS && n > 1) {
        m->log.resize(n);
        nvrtcGetProgramLog(prog, &m->log[0]);
    }
    *out = m;  // the log is useful on failure as well
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(PRB_ERR_COMPILE, "NVRTC compilation of `%s` failed: %s\n%s", m->name.c_str(),
                    nvrtcGetErrorString(r), m->log.c_str());
    }
    r = nvrtcGetCUBINSize(prog, &n);
    if (r != NVRTC_SUCCESS || n == 0) {
        nvrtcDestroyProgram(&prog);
        return fail(PRB_ERR_COMPILE, "no cubin produced for `%s` (arch must be a real sm_XX target)", m->name.c_str());
    }
    m->cubin.resize(n);
    nvrtcGetCUBIN(prog, m->cubin.data());
    nvrtcDestroyProgram(&prog);
    return PRB_OK;
}

int prb_module_from_cubin(const void* cubin, size_t size, prb_module** out) {
    if (!cubin || !size || !out) return fail(PRB_ERR_INVALID, "cubin/out is NULL");
    auto* m = new prb_module();
    m->name = "cached.cubin";
    m->cubin.assign((const char*)cubin, (const char*)cubin + size);
    *out = m;
    return PRB_OK;
}

int prb_module_cubin(const prb_module* m, const void** data, size_t* size) {
    if (!m || !data || !size) return fail(PRB_ERR_INVALID, "NULL argument");
    *data = m->cubin.data();
    *size = m->cubin.size();
    return PRB_OK;
}

int prb_module_log(const prb_module* m, const char** log) {
    if (!m || !log) return fail(PRB_ERR_INVALID, "NULL argument");
    *log = m->log.c_str();
    return PRB_OK;
}

int prb_module_free(prb_module* m) {
    if (!m) return PRB_OK;
    if (g_drv.ok) {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        CUcontext cur = nullptr;
        g_drv.cuCtxGetCurrent(&cur);
        for (auto& kv : m->on) {
            auto c = g_ctx_of.find(kv.first);
            if (!kv.second.mod || c == g_ctx_of.end()) continue;
            g_drv.cuCtxSetCurrent(c->second);   // a module is unloaded in the context that loaded it
            g_drv.cuModuleUnload(kv.second.mod);
        }
        if (cur) g_drv.cuCtxSetCurrent(cur);
    }
    delete m;
    return PRB_OK;
}

int prb_kernel_get_info(prb_module* m, const char* kernel, uint32_t block_dim, uint32_t smem_bytes,
                        prb_kernel_info* info) {
    if (!m || !kernel || !info) return fail(PRB_ERR_INVALID, "NULL argument");
    if (int e = need_ctx()) return e;
    CUfunction f;
    if (int e = module_function(m, kernel, &f)) return e;
    CU(cuFuncGetAttribute(&info->regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f));
    CU(cuFuncGetAttribute(&info->static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, f));
    CU(cuFuncGetAttribute(&info->local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f));
    CU(cuFuncGetAttribute(&info->max_threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, f));
    info->max_blocks_per_sm = 0;
    if (block_dim) {
        if (smem_bytes > 48 * 1024)
            CU(cuFuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_bytes));
        CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&info->max_blocks_per_sm, f, (int)block_dim, smem_bytes));
    }
    return PRB_OK;
}

// ---------------------------------------------------------------- run
int prb_run(prb_module* m, const char* kernel, const prb_run_desc* d) {
    if (!m || !kernel || !d) return fail(PRB_ERR_INVALID, "NULL argument");
    if (d->struct_size != sizeof(prb_run_desc))
        return fail(PRB_ERR_INVALID, "prb_run_desc size mismatch (got %u, library expects %zu)", d->struct_size,
                    sizeof(prb_run_desc));
    if (d->block_dim == 0 || d->block_dim > 1024) return fail(PRB_ERR_INVALID, "block_dim %u out of range", d->block_dim);
    const prb_kernel_args& a = d->args;
    if (a.n_inst <= 0 || a.n_steps < 0 || a.store_step <= 0)
        return fail(PRB_ERR_INVALID, "n_inst/n_steps/store_step out of range");
    if (!a.y) return fail(PRB_ERR_INVALID, "state pointer y is NULL");
    if (int e = need_ctx()) return e;
    CUfunction f;
    if (int e = module_function(m, kernel, &f)) return e;
    if (d->smem_bytes > 48 * 1024)
        CU(cuFuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)d->smem_bytes));
    unsigned grid = d->grid_dim;
    if (d->cooperative) {
        int per_sm = 0, sms = 0;
        CUdevice dev;
        CU(cuDeviceGet(&dev, g_dev));
        CU(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
        CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, (int)d->block_dim, d->smem_bytes));
        if (per_sm < 1) return fail(PRB_ERR_LAUNCH, "kernel `%s` cannot be resident with block %u / smem %u", kernel, d->block_dim, d->smem_bytes);
        unsigned max_grid = (unsigned)(per_sm * sms);
        if (grid == 0) grid = (unsigned)sms;          // persistent: one CTA per SM
        if (grid > max_grid)
            return fail(PRB_ERR_LAUNCH, "cooperative grid %u exceeds co-resident capacity %u", grid, max_grid);
        if (!a.sync) return fail(PRB_ERR_INVALID, "cooperative kernels need args.sync (grid barrier words)");
        if (a.world > 1 && !a.peers) return fail(PRB_ERR_INVALID, "multi-GPU network kernels need args.peers");
    } else if (grid == 0) {
        grid = (unsigned)((a.n_inst + d->block_dim - 1) / d->block_dim);
    }
    prb_kernel_args args = a;
    void* params[] = {&args};
    CUstream s = (CUstream)d->stream;
    if (d->cooperative) {
        CU(cuLaunchCooperativeKernel(f, grid, 1, 1, d->block_dim, 1, 1, d->smem_bytes, s, params));
    } else {
        CU(cuLaunchKernel(f, grid, 1, 1, d->block_dim, 1, 1, d->smem_bytes, s, params, nullptr));
    }
    g_launches.fetch_add(1);
    return PRB_OK;
}

// ---------------------------------------------------------------- memory / streams / events
int prb_malloc(void** dptr, size_t bytes) {
    if (!dptr) return fail(PRB_ERR_INVALID, "dptr is NULL");
    if (int e = need_ctx()) return e;
    CUdeviceptr p = 0;
    CU(cuMemAlloc_v2(&p, bytes ? bytes : 1));
    *dptr = (void*)p;
    return PRB_OK;
}
int prb_free(void* dptr) {
    if (!dptr) return PRB_OK;
    if (int e = need_ctx()) return e;
    CU(cuMemFree_v2((CUdeviceptr)dptr));
    return PRB_OK;
}
int prb_memset(void* dptr, int byte, size_t bytes, void* stream) {
    if (int e = need_ctx()) return e;
    if (bytes) CU(cuMemsetD8Async((CUdeviceptr)dptr, (unsigned char)byte, bytes, (CUstream)stream));
    return PRB_OK;
}
int prb_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (int e = need_ctx()) return e;
    if (bytes) CU(cuMemcpyHtoDAsync_v2((CUdeviceptr)dst, src, bytes, (CUstream)stream));
    return PRB_OK;
}
int prb_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    if (int e = need_ctx()) return e;
    if (bytes) CU(cuMemcpyDtoHAsync_v2(dst, (CUdeviceptr)src, bytes, (CUstream)stream));
    return PRB_OK;
}
int prb_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (int e = need_ctx()) return e;
    if (bytes) CU(cuMemcpyDtoDAsync_v2((CUdeviceptr)dst, (CUdeviceptr)src, bytes, (CUstream)stream));
    return PRB_OK;
}
int prb_ipc_get_handle(void* dptr, unsigned char* handle) {
    if (!dptr || !handle) return fail(PRB_ERR_INVALID, "NULL argument");
    if (int e = need_ctx()) return e;
    static_assert(sizeof(CUipcMemHandle) == PRB_IPC_HANDLE_BYTES, "IPC handle size");
    CUipcMemHandle h;
    CU(cuIpcGetMemHandle(&h, (CUdeviceptr)dptr));
    memcpy(handle, &h, sizeof h);
    return PRB_OK;
}
int prb_ipc_open_handle(const unsigned char* handle, void** dptr) {
    if (!dptr || !handle) return fail(PRB_ERR_INVALID, "NULL argument");
    if (int e = need_ctx()) return e;
    CUipcMemHandle h;
    memcpy(&h, handle, sizeof h);
    CUdeviceptr p = 0;
    CU(cuIpcOpenMemHandle_v2(&p, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS));
    *dptr = (void*)p;
    return PRB_OK;
}
int prb_ipc_close_handle(void* dptr) {
    if (!dptr) return PRB_OK;
    if (int e = need_ctx()) return e;
    CU(cuIpcCloseMemHandle((CUdeviceptr)dptr));
    return PRB_OK;
}
int prb_host_alloc(void** hptr, size_t bytes) {
    if (!hptr) return fail(PRB_ERR_INVALID, "hptr is NULL");
    if (int e = need_ctx()) return e;
    CU(cuMemHostAlloc(hptr, bytes ? bytes : 1, CU_MEMHOSTALLOC_PORTABLE));
    return PRB_OK;
}
int prb_host_free(void* hptr) {
    if (!hptr) return PRB_OK;
    if (int e = need_ctx()) return e;
    CU(cuMemFreeHost(hptr));
    return PRB_OK;
}
int prb_host_register(void* hptr, size_t bytes) {
    if (!hptr || !bytes) return fail(PRB_ERR_INVALID, "hptr/bytes is NULL");
    if (int e = need_ctx()) return e;
    CU(cuMemHostRegister_v2(hptr, bytes, CU_MEMHOSTREGISTER_PORTABLE));
    return PRB_OK;
}
int prb_host_unregister(void* hptr) {
    if (!hptr) return PRB_OK;
    if (int e = need_ctx()) return e;
    CU(cuMemHostUnregister(hptr));
    return PRB_OK;
}
int prb_stream_create(void** stream) {
    if (!stream) return fail(PRB_ERR_INVALID, "stream is NULL");
    if (int e = need_ctx()) return e;
    CUstream s;
    CU(cuStreamCreate(&s, CU_STREAM_NON_BLOCKING));
    *stream = s;
    return PRB_OK;
}
int prb_stream_destroy(void* stream) {
    if (int e = need_ctx()) return e;
    if (stream) CU(cuStreamDestroy_v2((CUstream)stream));
    return PRB_OK;
}
int prb_stream_synchronize(void* stream) {
    if (int e = need_ctx()) return e;
    CU(cuStreamSynchronize((CUstream)stream));
    return PRB_OK;
}
int prb_stream_wait_event(void* stream, void* event) {
    if (int e = need_ctx()) return e;
    CU(cuStreamWaitEvent((CUstream)stream, (CUevent)event, 0));
    return PRB_OK;
}
int prb_event_create(void** event) {
    if (!event) return fail(PRB_ERR_INVALID, "event is NULL");
    if (int e = need_ctx()) return e;
    CUevent ev;
    CU(cuEventCreate(&ev, CU_EVENT_DEFAULT));
    *event = ev;
    return PRB_OK;
}
int prb_event_destroy(void* event) {
    if (int e = need_ctx()) return e;
    if (event) CU(cuEventDestroy_v2((CUevent)event));
    return PRB_OK;
}
int prb_event_record(void* event, void* stream) {
    if (int e = need_ctx()) return e;
    CU(cuEventRecord((CUevent)event, (CUstream)stream));
    return PRB_OK;
}
int prb_event_synchronize(void* event) {
    if (int e = need_ctx()) return e;
    CU(cuEventSynchronize((CUevent)event));
    return PRB_OK;
}
int prb_event_elapsed_ms(void* start, void* stop, float* ms) {
    if (!ms) return fail(PRB_ERR_INVALID, "ms is NULL");
    if (int e = need_ctx()) return e;
    CU(cuEventElapsedTime(ms, (CUevent)start, (CUevent)stop));
    return PRB_OK;
}

}  // extern "C"
