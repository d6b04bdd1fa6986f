//------------------------------------------------------------------------------
//  runtime.cpp -- implementation of include/gfb200.h (the device layer).
//
//  B200-native replacement for the device half of
//  /root/reference/graph_framework/cuda_context.hpp: NVRTC straight to an
//  sm_100a cubin (the reference goes through generic compute_XY PTX with a
//  128-register cap, cuda_context.hpp:21,218-299), explicit device buffers with
//  pinned host mirrors (the reference uses managed memory for everything,
//  :334-356), a grid-wide max reduction (the reference uses one block, :566-575)
//  and deferred launches that fuse consecutive steps of one kernel into a
//  single register-resident multi-step launch.
//
//  The CUDA driver is reached through cudaGetDriverEntryPoint so the library
//  loads (and NVRTC-only entry points work) on machines without libcuda.
//------------------------------------------------------------------------------
#include "../../include/gfb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#include <unistd.h>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

extern "C" {
int gfb_k_max(const double *in, unsigned long long n, unsigned long long *result, int sms, cudaStream_t s);
double gfb_k_unorder(unsigned long long key);
int gfb_k_deposit(const double *x, const double *y, const double *z, const double *w, unsigned long long n,
                  double *hist, const double *lo, const double *hi, const int *bins, int sms, cudaStream_t s);
int gfb_k_fp64_peak(double *scratch, int iters, int sms, cudaStream_t s);
int gfb_k_fp64_peak_regs(double *scratch, int iters, int sms, cudaStream_t s);
int gfb_k_bin_permutation(const double *const values[3], unsigned n, const double lo[2], const double hi[2],
                          const unsigned cells01[2], unsigned *work, unsigned *perm, int sms, cudaStream_t s);
int gfb_k_permute(double *dst, const double *src, const unsigned *perm, unsigned n, int scatter, int sms, cudaStream_t s);
int gfb_k_bin_breaks(const double *const values[3], unsigned n, const double lo[2], const double hi[2],
                     const unsigned cells01[2], unsigned *breaks, int sms, cudaStream_t s);
int gfb_k_compose(unsigned *total, const unsigned *first, const unsigned *second, unsigned n, int sms, cudaStream_t s);
int gfb_k_fill(double *p, size_t n, double v, int sms, cudaStream_t s);
int gfb_k_sum_peers(double *dst, const double *const *src, int num, size_t count, int sms, cudaStream_t s);
}

namespace {
const char *skeleton_text =
#include "skeleton_text.inc"
;
//  Special functions (erfi) are only compiled into modules that call them.
const char *special_text =
#include "special_text.inc"
;

thread_local std::string last_error;

std::string full_source(const char *source) {
    std::string text(skeleton_text);
    if (std::strstr(source, "gfb::erfi(")) text += special_text;
    return text + source;
}

//  GFB_DEBUG=1: trace every device-layer call and print a raw backtrace on SIGSEGV
//  (resolve the offsets with addr2line on libgfb200.so).
bool debug_enabled() {
    static const bool on = std::getenv("GFB_DEBUG") != nullptr;
    return on;
}
void segv_handler(int sig) {
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "gfb200: fatal signal, backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
struct debug_init {
    debug_init() {
        if (debug_enabled()) {
            signal(SIGSEGV, segv_handler);
            signal(SIGBUS, segv_handler);
        }
    }
} debug_init_instance;
#define GFB_TRACE(...) do { if (debug_enabled()) { std::fprintf(stderr, "[gfb] " __VA_ARGS__); std::fprintf(stderr, "\n"); std::fflush(stderr); } } while (0)

int fail(const std::string &what) {
    last_error = what;
    return 1;
}
int check(const cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    return fail(std::string(what) + ": " + cudaGetErrorString(e));
}

//  Mirror of the device-side argument block in skeleton.cuh.
constexpr int max_ptrs = 56;
struct device_args {
    void *ptr[max_ptrs];
    unsigned long long n;
    unsigned steps;
    unsigned flags;
    double scalar[4];
};

struct driver_api {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void **, void **) = nullptr;
    CUresult (*FuncGetAttribute)(int *, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    bool loaded = false;
};
driver_api driver;

template<typename F>
bool entry(const char *name, F &f) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
    f = reinterpret_cast<F> (p);
    return true;
}
std::once_flag driver_once;
int load_driver_once() {
    if (!entry("cuModuleLoadData", driver.ModuleLoadData) ||
        !entry("cuModuleUnload", driver.ModuleUnload) ||
        !entry("cuModuleGetFunction", driver.ModuleGetFunction) ||
        !entry("cuLaunchKernel", driver.LaunchKernel) ||
        !entry("cuFuncGetAttribute", driver.FuncGetAttribute) ||
        !entry("cuFuncSetAttribute", driver.FuncSetAttribute) ||
        !entry("cuGetErrorString", driver.GetErrorString)) {
        return 1;
    }
    driver.loaded = true;
    return 0;
}
//  One host thread per device is the reference's model (xrays.cpp:419-527): contexts may be created
//  concurrently, so the entry points are resolved exactly once.
int load_driver() {
    std::call_once(driver_once, [] { load_driver_once(); });
    return driver.loaded ? 0 : fail("CUDA driver entry points unavailable (no NVIDIA driver?)");
}
int check_cu(const CUresult r, const char *what) {
    if (r == CUDA_SUCCESS) return 0;
    const char *s = nullptr;
    if (driver.GetErrorString) driver.GetErrorString(r, &s);
    return fail(std::string(what) + ": " + (s ? s : "unknown driver error"));
}

struct buffer {
    void *dev = nullptr;
    size_t bytes = 0;
    bool owned = true;
    void *host = nullptr;       // pinned mirror, created on demand
};

int nvrtc_compile(const std::string &full_source, const char *options, std::vector<char> &cubin, std::string &log) {
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, full_source.c_str(), "gfb_kernels.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        return fail("nvrtcCreateProgram failed");
    }
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    bool has_unroll = false;
    if (options) {
        std::istringstream is(options);
        std::string o;
        while (is >> o) {
            opts.push_back(o);
            if (o.rfind("-DGFB_UNROLL_STAGES", 0) == 0) has_unroll = true;
        }
    }
    if (!has_unroll) opts.push_back("-DGFB_UNROLL_STAGES=0");
    std::vector<const char *> copts;
    for (auto &o : opts) copts.push_back(o.c_str());
    const nvrtcResult r = nvrtcCompileProgram(prog, static_cast<int> (copts.size()), copts.data());
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    log.assign(log_size, '\0');
    if (log_size) nvrtcGetProgramLog(prog, log.data());
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(std::string("NVRTC: ") + nvrtcGetErrorString(r) + "\n" + log);
    }
    size_t size = 0;
    nvrtcGetCUBINSize(prog, &size);
    cubin.resize(size);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return 0;
}
}  // namespace

struct gfb_kernel {
    gfb_ctx *ctx;
    std::string name;
    std::vector<uint64_t> slot_keys;     // buffer key behind every pointer slot
    CUfunction function;
    device_args args;
    unsigned block;
    unsigned grid;
    size_t smem;
    int kind;
    bool can_repeat;
    bool serial = false;            // every step is its own launch (the kernel gathers from an array it rewrites)
};

struct gfb_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    CUmodule module = nullptr;
    std::map<uint64_t, buffer> buffers;
    std::vector<std::unique_ptr<gfb_kernel>> kernels;
    std::string source, log, options;      // options: what NVRTC was finally given (incl. the chosen GFB_MIN_BLOCKS)
    gfb_kernel *pending = nullptr;
    unsigned pending_steps = 0;
    unsigned max_fused = 1024;
    int min_blocks = 0;
    uint64_t launches = 0;
    unsigned long long *scratch = nullptr;      // device scalar for reductions
    unsigned long long *scratch_host = nullptr; // pinned
    double *flush_buffer = nullptr;
    size_t flush_count = 0;
//  Trajectory snapshots: two staging slots, a copy stream, events for the hand-over.
    cudaStream_t copy_stream = nullptr;
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    cudaEvent_t staged[2] = {nullptr, nullptr};     // slot filled (compute stream)
    cudaEvent_t drained[2] = {nullptr, nullptr};    // slot copied out (copy stream)
    unsigned next_slot = 0;
    cudaStream_t upload_stream = nullptr;
    cudaStream_t stream2 = nullptr;         // second compute stream of the chunked host pipeline
//  Ray binning: the permutation in force (perm[slot] = original ray), scratch for the moves.
    unsigned *bin_perm = nullptr;
    unsigned *bin_step = nullptr;       // the permutation of one re-sort, before it is composed into bin_perm
    unsigned *bin_work = nullptr;
    double *bin_scratch = nullptr;
    size_t bin_capacity = 0, bin_cells = 0, bin_n = 0;
    bool binned = false;
//  All-reduce over peer memory (gfb_allreduce_sum_f64): this device's summed slice, hand-over events.
    double *reduce_scratch = nullptr;
    size_t reduce_capacity = 0;
    cudaEvent_t reduce_ready = nullptr, reduce_summed = nullptr, reduce_done = nullptr;
};

namespace {
int launch_now(gfb_kernel *k, const unsigned steps) {
    GFB_TRACE("launch %s steps=%u grid=%u block=%u smem=%zu", k->name.c_str(), steps, k->grid, k->block, k->smem);
    gfb_ctx *c = k->ctx;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    k->args.steps = k->can_repeat || k->kind == 2 ? steps : (steps ? 1u : 0u);
    void *params[] = {&k->args};
    c->launches++;
    return check_cu(driver.LaunchKernel(k->function, k->grid, 1, 1, k->block, 1, 1,
                                        static_cast<unsigned> (k->smem), reinterpret_cast<CUstream> (c->stream),
                                        params, nullptr), k->name.c_str());
}
int flush(gfb_ctx *c) {
    if (!c->pending) return 0;
    gfb_kernel *k = c->pending;
    const unsigned steps = c->pending_steps;
    c->pending = nullptr;
    c->pending_steps = 0;
    return launch_now(k, steps);
}
}  // namespace

extern "C" {

const char *gfb_last_error(void) { return last_error.c_str(); }
void gfb_set_last_error(const char *text) { last_error = text ? text : ""; }
const char *gfb_version(void) { return "gfb200 0.1 (sm_100a)"; }

int gfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

gfb_ctx *gfb_ctx_create(int device) {
    GFB_TRACE("ctx_create %d", device);
    int n = gfb_device_count();
    if (n <= 0) {
        fail("no CUDA device: the B200 back end has no CPU fallback");
        return nullptr;
    }
    device = device%n;
    if (check(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    if (check(cudaFree(nullptr), "context init")) return nullptr;
    if (load_driver()) return nullptr;
    auto c = std::make_unique<gfb_ctx> ();
    c->device = device;
    cudaDeviceProp prop;
    if (check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return nullptr;
    if (prop.major != 10) {
        fail(std::string("device ") + prop.name + " is not sm_100: this back end targets B200 only");
        return nullptr;
    }
    c->sms = prop.multiProcessorCount;
    if (check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return nullptr;
    cudaEventCreate(&c->ev_start);
    cudaEventCreate(&c->ev_stop);
    if (check(cudaMalloc(&c->scratch, 64), "cudaMalloc")) return nullptr;
    if (check(cudaMallocHost(&c->scratch_host, 64), "cudaMallocHost")) return nullptr;
    return c.release();
}

void gfb_ctx_destroy(gfb_ctx *c) {
    GFB_TRACE("ctx_destroy %p", (void *)c);
    if (!c) return;
    cudaSetDevice(c->device);
    flush(c);
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->buffers) {
        if (kv.second.owned && kv.second.dev) cudaFree(kv.second.dev);
        if (kv.second.host) cudaFreeHost(kv.second.host);
    }
    if (c->module) driver.ModuleUnload(c->module);
    if (c->scratch) cudaFree(c->scratch);
    if (c->scratch_host) cudaFreeHost(c->scratch_host);
    if (c->bin_perm) cudaFree(c->bin_perm);
    if (c->bin_step) cudaFree(c->bin_step);
    if (c->bin_work) cudaFree(c->bin_work);
    if (c->bin_scratch) cudaFree(c->bin_scratch);
    if (c->flush_buffer) cudaFree(c->flush_buffer);
    if (c->reduce_scratch) cudaFree(c->reduce_scratch);
    for (cudaEvent_t e : {c->reduce_ready, c->reduce_summed, c->reduce_done}) if (e) cudaEventDestroy(e);
    if (c->upload_stream) {
        cudaStreamSynchronize(c->upload_stream);
        cudaStreamDestroy(c->upload_stream);
    }
    if (c->stream2) {
        cudaStreamSynchronize(c->stream2);
        cudaStreamDestroy(c->stream2);
    }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
        for (int i = 0; i < 2; i++) {
            if (c->stage[i]) cudaFree(c->stage[i]);
            if (c->staged[i]) cudaEventDestroy(c->staged[i]);
            if (c->drained[i]) cudaEventDestroy(c->drained[i]);
        }
    }
    cudaEventDestroy(c->ev_start);
    cudaEventDestroy(c->ev_stop);
    cudaStreamDestroy(c->stream);
    delete c;
}

int gfb_ctx_device_info(gfb_ctx *c, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor) {
    cudaDeviceProp prop;
    if (check(cudaGetDeviceProperties(&prop, c->device), "cudaGetDeviceProperties")) return 1;
    if (name && name_len) {
        std::strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = '\0';
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int gfb_compile_to_cubin(const char *source, const char *options, void **cubin, size_t *cubin_size, char **log) {
    std::vector<char> image;
    std::string text;
    const int r = nvrtc_compile(full_source(source), options, image, text);
    if (log) {
        const std::string &msg = r ? last_error : text;
        *log = static_cast<char *> (std::malloc(msg.size() + 1));
        std::memcpy(*log, msg.c_str(), msg.size() + 1);
    }
    if (r) return r;
    *cubin = std::malloc(image.size());
    std::memcpy(*cubin, image.data(), image.size());
    *cubin_size = image.size();
    return 0;
}
void gfb_free(void *p) { std::free(p); }

int gfb_compile(gfb_ctx *c, const char *source, const char *const *names, int num_names, const char *options) {
    GFB_TRACE("compile %zu bytes", std::strlen(source));
    (void)names; (void)num_names;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (flush(c)) return 1;
    c->source = full_source(source);
    if (c->module) {
        cudaStreamSynchronize(c->stream);
        driver.ModuleUnload(c->module);
        c->module = nullptr;
        c->kernels.clear();
    }
//  Occupancy choice.  The emitted kernels carry __launch_bounds__(128, GFB_MIN_BLOCKS).  Compile once
//  without a promise to learn the natural register count R and the blocks/SM m0 that fit with it;
//  then try to squeeze in one or two more blocks (m0 + 2, m0 + 1): FP64- and gather-latency-bound
//  bodies want warps, but a body that spills pays for them in L1 traffic.  A few spilled doubles are
//  cheaper than a lost block (measured, profiles/r1_sweep4_*.txt: X-mode wins with 48-104 B spilled
//  at 3 blocks/SM, cold plasma loses with 344 B), so a candidate is accepted when it adds at most
//  128 bytes of local memory per thread.
    const std::string user = options ? options : "";
//  Nothing to choose when the caller pinned the macro or no kernel uses it.
    const bool pinned = user.find("-DGFB_MIN_BLOCKS") != std::string::npos ||
                        std::string(source).find("GFB_MIN_BLOCKS)") == std::string::npos;
    struct variant { CUmodule module = nullptr; int regs = 0; int local = 0; };
    auto options_for = [&] (const int mb) {
        return (pinned || mb == 0) ? user : user + " -DGFB_MIN_BLOCKS=" + std::to_string(mb);
    };
    auto load = [&] (const int mb, const std::vector<char> &image, variant &v) -> int {
        if (check_cu(driver.ModuleLoadData(&v.module, image.data()), "cuModuleLoadData")) return 1;
        for (int i = 0; i < num_names; i++) {
            CUfunction f;
            int value = 0;
            if (driver.ModuleGetFunction(&f, v.module, names[i]) != CUDA_SUCCESS) continue;
            if (driver.FuncGetAttribute(&value, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f) == CUDA_SUCCESS && value > v.local) v.local = value;
            if (driver.FuncGetAttribute(&value, CU_FUNC_ATTRIBUTE_NUM_REGS, f) == CUDA_SUCCESS && value > v.regs) v.regs = value;
        }
        GFB_TRACE("compile min_blocks=%d regs=%d local=%d", mb, v.regs, v.local);
        return 0;
    };
    variant natural;
    {
        std::vector<char> image;
        if (nvrtc_compile(c->source, options_for(1).c_str(), image, c->log)) {
            std::fprintf(stderr, "%s\n", last_error.c_str());
            return 1;
        }
        if (load(1, image, natural)) return 1;
    }
    c->module = natural.module;
    c->min_blocks = 0;
    c->options = options_for(1);
    if (pinned) return 0;
    const int regs8 = (natural.regs + 7)/8*8;
    int m0 = regs8 > 0 ? 65536/(128*regs8) : 1;
    m0 = m0 < 1 ? 1 : (m0 > 8 ? 8 : m0);
    c->min_blocks = m0;
//  The candidates are independent NVRTC runs: compile them side by side, then take the first
//  (highest) one that qualifies.
    std::vector<int> candidates;
    for (int mb = (m0 + 2 > 8 ? 8 : m0 + 2); mb > m0; mb--) candidates.push_back(mb);
    std::vector<std::vector<char>> images(candidates.size());
    std::vector<std::string> logs(candidates.size()), errors(candidates.size());
    std::vector<int> status(candidates.size(), 0);
    std::vector<std::thread> workers;
    for (size_t i = 0; i < candidates.size(); i++) {
        workers.emplace_back([&, i] {
            status[i] = nvrtc_compile(c->source, options_for(candidates[i]).c_str(), images[i], logs[i]);
            if (status[i]) errors[i] = last_error;      // last_error is per thread
        });
    }
    for (auto &w : workers) w.join();
    for (size_t i = 0; i < candidates.size(); i++) {
        if (status[i]) {
            std::fprintf(stderr, "%s\n", errors[i].c_str());
            return fail(errors[i]);
        }
        variant v;
        if (load(candidates[i], images[i], v)) return 1;
        if (v.local - natural.local <= 128) {
            driver.ModuleUnload(natural.module);
            c->module = v.module;
            c->min_blocks = candidates[i];
            c->log = logs[i];
            c->options = options_for(candidates[i]);
            return 0;
        }
        driver.ModuleUnload(v.module);
    }
    return 0;
}
int gfb_compiled_min_blocks(gfb_ctx *c) { return c->min_blocks; }
const char *gfb_source(gfb_ctx *c) { return c->source.c_str(); }
const char *gfb_compile_log(gfb_ctx *c) { return c->log.c_str(); }
const char *gfb_compile_options(gfb_ctx *c) { return c->options.c_str(); }

int gfb_buffer(gfb_ctx *c, uint64_t key, size_t bytes, const void *init, void **device_ptr) {
    GFB_TRACE("buffer key=%llx bytes=%zu init=%p", (unsigned long long)key, bytes, init);
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) {
        buffer b;
        b.bytes = bytes;
        if (check(cudaMalloc(&b.dev, bytes ? bytes : 8), "cudaMalloc")) return 1;
        if (init) {
            if (check(cudaMemcpyAsync(b.dev, init, bytes, cudaMemcpyHostToDevice, c->stream), "upload")) return 1;
            if (check(cudaStreamSynchronize(c->stream), "upload sync")) return 1;   // `init` may be a temporary
        } else {
            if (check(cudaMemsetAsync(b.dev, 0, bytes, c->stream), "cudaMemset")) return 1;
        }
        it = c->buffers.emplace(key, b).first;
    } else if (it->second.bytes < bytes) {
        return fail("gfb_buffer: key exists with a smaller size");
    }
    if (device_ptr) *device_ptr = it->second.dev;
    return 0;
}

int gfb_buffer_import(gfb_ctx *c, uint64_t key, void *device_ptr, size_t bytes) {
    auto it = c->buffers.find(key);
    if (it != c->buffers.end()) {
        if (flush(c)) return 1;
        if (it->second.owned && it->second.dev) {
            cudaStreamSynchronize(c->stream);
            cudaFree(it->second.dev);
        }
        it->second.dev = device_ptr;
        it->second.bytes = bytes;
        it->second.owned = false;
        if (it->second.host) {
            cudaFreeHost(it->second.host);
            it->second.host = nullptr;
        }
//  Kernels created earlier must see the adopted memory.
        for (auto &k : c->kernels) {
            for (size_t i = 0; i < k->slot_keys.size(); i++) {
                if (k->slot_keys[i] == key) k->args.ptr[i] = device_ptr;
            }
        }
        return 0;
    }
    buffer b;
    b.dev = device_ptr;
    b.bytes = bytes;
    b.owned = false;
    c->buffers.emplace(key, b);
    return 0;
}

int gfb_buffer_lookup(gfb_ctx *c, uint64_t key, void **device_ptr, size_t *bytes) {
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) {
        if (device_ptr) *device_ptr = nullptr;
        if (bytes) *bytes = 0;
        return fail("gfb_buffer_lookup: unknown key");
    }
    if (device_ptr) *device_ptr = it->second.dev;
    if (bytes) *bytes = it->second.bytes;
    return 0;
}

int gfb_kernel_create(gfb_ctx *c, const char *name, const uint64_t *ptr_keys, int num_ptrs,
                      size_t num_rays, unsigned block_size, size_t dynamic_smem, int kind, int can_repeat,
                      gfb_kernel **kernel) {
    GFB_TRACE("kernel_create %s ptrs=%d n=%zu block=%u smem=%zu kind=%d", name, num_ptrs, num_rays, block_size, dynamic_smem, kind);
    if (!c->module) return fail("gfb_kernel_create: nothing compiled");
    if (num_ptrs > max_ptrs) return fail("gfb_kernel_create: too many pointer arguments");
    auto k = std::make_unique<gfb_kernel> ();
    k->ctx = c;
    k->name = name;
    if (check_cu(driver.ModuleGetFunction(&k->function, c->module, name), name)) return 1;
    std::memset(&k->args, 0, sizeof(k->args));
    for (int i = 0; i < num_ptrs; i++) {
        auto it = c->buffers.find(ptr_keys[i]);
        if (it == c->buffers.end()) return fail(std::string("gfb_kernel_create: missing buffer for ") + name);
        k->args.ptr[i] = it->second.dev;
        k->slot_keys.push_back(ptr_keys[i]);
    }
    k->args.n = num_rays;
    k->block = block_size;
    k->grid = static_cast<unsigned> ((num_rays + block_size - 1)/block_size);
    if (k->grid == 0) k->grid = 1;
    k->smem = dynamic_smem;
    k->kind = kind;
    k->can_repeat = can_repeat == 1;
    k->serial = can_repeat == 2;
    if (dynamic_smem > 48*1024) {
        if (check_cu(driver.FuncSetAttribute(k->function, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                             static_cast<int> (dynamic_smem)), "smem attribute")) return 1;
    }
    *kernel = k.get();
    c->kernels.push_back(std::move(k));
    return 0;
}

int gfb_kernel_run(gfb_kernel *k) {
    GFB_TRACE("kernel_run %s", k->name.c_str());
    gfb_ctx *c = k->ctx;
    if (k->serial) {
        if (flush(c)) return 1;
        return launch_now(k, 1);
    }
    if (c->pending == k && c->pending_steps < c->max_fused && (k->kind != 2)) {
        c->pending_steps++;
        return 0;
    }
    if (flush(c)) return 1;
    if (k->kind == 2) {
        return launch_now(k, k->args.steps);    // Newton: steps holds max iterations
    }
    c->pending = k;
    c->pending_steps = 1;
    return 0;
}

int gfb_kernel_launch(gfb_kernel *k, unsigned steps) {
    if (flush(k->ctx)) return 1;
    if (k->serial) {
        for (unsigned s = 0; s < steps; s++) if (launch_now(k, 1)) return 1;
        return 0;
    }
    return launch_now(k, steps);
}

int gfb_kernel_run_from_host(gfb_kernel *k, unsigned steps, int num_ray_slots,
                             const void *const *host_src, void *const *host_dst, int chunks) {
    gfb_ctx *c = k->ctx;
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (num_ray_slots > max_ptrs) return fail("gfb_kernel_run_from_host: too many slots");
    if (!c->copy_stream) {
        if (check(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "copy stream")) return 1;
        for (int i = 0; i < 2; i++) {
            cudaEventCreateWithFlags(&c->staged[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->drained[i], cudaEventDisableTiming);
        }
    }
    if (!c->upload_stream) {
        if (check(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking), "upload stream")) return 1;
    }
    if (!c->stream2) {
        if (check(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking), "second compute stream")) return 1;
    }
    const unsigned long long total = k->args.n;
    if (chunks < 1) chunks = 1;
//  Chunk boundaries on whole waves (blocks/SM x SMs x block) so only the last piece has a tail.  The
//  first upload and the last read-back cannot overlap any kernel, so with three or more chunks the first
//  piece is a single wave and the last one is the (partial) tail wave; the pieces between share the rest.
    const unsigned long long wave = static_cast<unsigned long long> (k->block)*c->sms*(c->min_blocks > 0 ? c->min_blocks : 1);
    const unsigned long long waves = (total + wave - 1)/wave;
    std::vector<std::pair<unsigned long long, unsigned long long>> pieces;       // offset, count
    if (chunks >= 3 && waves >= 2ull*static_cast<unsigned long long> (chunks)) {
        const unsigned long long middle_waves = waves - 2;
        const unsigned long long per_middle = (middle_waves + chunks - 3)/(chunks - 2);
        unsigned long long off = 0;
        pieces.push_back({off, wave});
        off += wave;
        for (unsigned long long done = 0; done < middle_waves; done += per_middle) {
            const unsigned long long cnt = std::min(per_middle, middle_waves - done)*wave;
            pieces.push_back({off, cnt});
            off += cnt;
        }
        pieces.push_back({off, total - off});
    } else {
        unsigned long long per = (total + chunks - 1)/chunks;
        per = (per + wave - 1)/wave*wave;
        for (unsigned long long off = 0; off < total; off += per) pieces.push_back({off, std::min(per, total - off)});
    }
//  The uploads overwrite arrays that work already queued on the compute stream may still read or
//  write (an earlier step, the scatter that restores the caller's ray order): order them after it.
    {
        cudaEvent_t quiet;
        cudaEventCreateWithFlags(&quiet, cudaEventDisableTiming);
        cudaEventRecord(quiet, c->stream);
        cudaStreamWaitEvent(c->upload_stream, quiet, 0);
        cudaStreamWaitEvent(c->stream2, quiet, 0);
        cudaEventDestroy(quiet);
    }
    std::vector<cudaEvent_t> uploaded, computed;
    int rc = 0;
    for (size_t piece_index = 0; piece_index < pieces.size() && !rc; piece_index++) {
        const unsigned long long off = pieces[piece_index].first, cnt = pieces[piece_index].second;
        cudaEvent_t up, done;
        cudaEventCreateWithFlags(&up, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        uploaded.push_back(up);
        computed.push_back(done);
        for (int s = 0; s < num_ray_slots && !rc; s++) {
            if (host_src && host_src[s]) {
                rc = check(cudaMemcpyAsync(static_cast<char *> (k->args.ptr[s]) + off*sizeof(double),
                                           static_cast<const char *> (host_src[s]) + off*sizeof(double),
                                           cnt*sizeof(double), cudaMemcpyHostToDevice, c->upload_stream), "chunk h2d");
            }
        }
        if (rc) break;
        cudaEventRecord(up, c->upload_stream);
//  Pieces alternate between two compute streams: a launch of 100 dependent steps ends with a tail in which
//  the machine empties (one thread's chain takes ~0.2 ms however few rays are left); the next piece, on the
//  other stream, fills it.  Pieces are disjoint ranges of rays, so they may overlap freely.
        cudaStream_t compute = (piece_index & 1) ? c->stream2 : c->stream;
        cudaStreamWaitEvent(compute, up, 0);
        device_args piece = k->args;
        for (int s = 0; s < num_ray_slots; s++) piece.ptr[s] = static_cast<char *> (k->args.ptr[s]) + off*sizeof(double);
        piece.n = cnt;
        piece.steps = k->can_repeat ? steps : (steps ? 1u : 0u);
        void *params[] = {&piece};
        c->launches++;
        rc = check_cu(driver.LaunchKernel(k->function, static_cast<unsigned> ((cnt + k->block - 1)/k->block), 1, 1, k->block, 1, 1,
                                          static_cast<unsigned> (k->smem), reinterpret_cast<CUstream> (compute), params, nullptr),
                      k->name.c_str());
        if (rc) break;
        cudaEventRecord(done, compute);
        cudaStreamWaitEvent(c->copy_stream, done, 0);
        for (int s = 0; s < num_ray_slots && !rc; s++) {
            if (host_dst && host_dst[s]) {
                rc = check(cudaMemcpyAsync(static_cast<char *> (host_dst[s]) + off*sizeof(double),
                                           static_cast<const char *> (k->args.ptr[s]) + off*sizeof(double),
                                           cnt*sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream), "chunk d2h");
            }
        }
    }
    const int sync_rc = check(cudaStreamSynchronize(c->copy_stream), "pipeline sync") |
                        check(cudaStreamSynchronize(c->stream), "pipeline sync") |
                        check(cudaStreamSynchronize(c->stream2), "pipeline sync") |
                        check(cudaStreamSynchronize(c->upload_stream), "pipeline sync");
    for (auto e : uploaded) cudaEventDestroy(e);
    for (auto e : computed) cudaEventDestroy(e);
    return rc | sync_rc;
}

int gfb_kernel_set_scalar(gfb_kernel *k, int index, double value) {
    if (index < 0 || index >= 4) return fail("scalar index out of range");
    if (flush(k->ctx)) return 1;
    k->args.scalar[index] = value;
    return 0;
}

int gfb_kernel_attributes(gfb_kernel *k, int *regs, int *static_smem, int *local_bytes, int *max_threads) {
    int v = 0;
    if (regs) { if (check_cu(driver.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, k->function), "attr")) return 1; *regs = v; }
    if (static_smem) { if (check_cu(driver.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, k->function), "attr")) return 1; *static_smem = v; }
    if (local_bytes) { if (check_cu(driver.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, k->function), "attr")) return 1; *local_bytes = v; }
    if (max_threads) { if (check_cu(driver.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, k->function), "attr")) return 1; *max_threads = v; }
    return 0;
}

uint64_t gfb_launch_count(gfb_ctx *c) { return c->launches; }
int gfb_set_max_fused_steps(gfb_ctx *c, unsigned steps) {
    if (flush(c)) return 1;
    c->max_fused = steps ? steps : 1;
    return 0;
}
int gfb_flush(gfb_ctx *c) { return flush(c); }

int gfb_max(gfb_ctx *c, uint64_t key, size_t n, double *result) {
    if (flush(c)) return 1;
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) return fail("gfb_max: unknown key");
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    c->launches++;
    if (gfb_k_max(static_cast<const double *> (it->second.dev), n, c->scratch, c->sms, c->stream)) return fail("max kernel launch failed");
    if (check(cudaMemcpyAsync(c->scratch_host, c->scratch, 8, cudaMemcpyDeviceToHost, c->stream), "max readback")) return 1;
    if (check(cudaStreamSynchronize(c->stream), "max sync")) return 1;
    *result = gfb_k_unorder(*c->scratch_host);
    return 0;
}

int gfb_wait(gfb_ctx *c) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    for (auto &kv : c->buffers) {
        if (kv.second.host) {
            if (check(cudaMemcpyAsync(kv.second.host, kv.second.dev, kv.second.bytes, cudaMemcpyDeviceToHost, c->stream), "mirror")) return 1;
        }
    }
    if (c->copy_stream && check(cudaStreamSynchronize(c->copy_stream), "copy stream sync")) return 1;
    return check(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
}

int gfb_snapshot_async(gfb_ctx *c, const uint64_t *keys, int num_keys, size_t bytes_each, void *host_destination) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    const size_t total = bytes_each*static_cast<size_t> (num_keys);
    if (!c->copy_stream) {
        if (check(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "copy stream")) return 1;
        for (int i = 0; i < 2; i++) {
            cudaEventCreateWithFlags(&c->staged[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->drained[i], cudaEventDisableTiming);
        }
    }
    if (c->stage_bytes < total) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->stream);
        for (int i = 0; i < 2; i++) {
            if (c->stage[i]) cudaFree(c->stage[i]);
            if (check(cudaMalloc(&c->stage[i], total), "staging buffer")) return 1;
        }
        c->stage_bytes = total;
    }
    const unsigned slot = c->next_slot;
    c->next_slot ^= 1u;
//  The slot may still be draining from two snapshots ago.
    if (check(cudaStreamWaitEvent(c->stream, c->drained[slot], 0), "wait drained")) return 1;
    for (int i = 0; i < num_keys; i++) {
        auto it = c->buffers.find(keys[i]);
        if (it == c->buffers.end()) return fail("gfb_snapshot_async: unknown key");
        if (it->second.bytes < bytes_each) return fail("gfb_snapshot_async: buffer smaller than bytes_each");
        char *staged = static_cast<char *> (c->stage[slot]) + bytes_each*i;
        if (c->binned) {
//  Rays are sorted by cell: the staging copy puts them back in the caller's order on the way out.
            if (bytes_each != c->bin_n*sizeof(double)) return fail("gfb_snapshot_async: binned rays need whole ray arrays");
            if (gfb_k_permute(reinterpret_cast<double *> (staged), static_cast<const double *> (it->second.dev), c->bin_perm,
                              static_cast<unsigned> (c->bin_n), 1, c->sms, c->stream)) return fail("snapshot scatter failed");
            c->launches++;
        } else if (check(cudaMemcpyAsync(staged, it->second.dev, bytes_each, cudaMemcpyDeviceToDevice, c->stream), "snapshot d2d")) {
            return 1;
        }
    }
    if (check(cudaEventRecord(c->staged[slot], c->stream), "record staged")) return 1;
    if (check(cudaStreamWaitEvent(c->copy_stream, c->staged[slot], 0), "wait staged")) return 1;
    if (check(cudaMemcpyAsync(host_destination, c->stage[slot], total, cudaMemcpyDeviceToHost, c->copy_stream), "snapshot d2h")) return 1;
    return check(cudaEventRecord(c->drained[slot], c->copy_stream), "record drained");
}

int gfb_copy_h2d(gfb_ctx *c, uint64_t key, const void *source, size_t bytes) {
    GFB_TRACE("copy_h2d key=%llx", (unsigned long long)key);
    if (flush(c)) return 1;
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) return fail("gfb_copy_h2d: unknown key");
    if (bytes == 0 || bytes > it->second.bytes) bytes = it->second.bytes;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (check(cudaMemcpyAsync(it->second.dev, source, bytes, cudaMemcpyHostToDevice, c->stream), "h2d")) return 1;
    return check(cudaStreamSynchronize(c->stream), "h2d sync");
}

int gfb_copy_d2h(gfb_ctx *c, uint64_t key, void *destination, size_t bytes) {
    GFB_TRACE("copy_d2h key=%llx dst=%p bytes=%zu", (unsigned long long)key, destination, bytes);
    if (flush(c)) return 1;
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) return fail("gfb_copy_d2h: unknown key");
    if (bytes == 0 || bytes > it->second.bytes) bytes = it->second.bytes;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (check(cudaMemcpyAsync(destination, it->second.dev, bytes, cudaMemcpyDeviceToHost, c->stream), "d2h")) return 1;
    return check(cudaStreamSynchronize(c->stream), "d2h sync");
}

int gfb_host_ptr(gfb_ctx *c, uint64_t key, void **host_ptr) {
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) return fail("gfb_host_ptr: unknown key");
    if (!it->second.host) {
        if (check(cudaMallocHost(&it->second.host, it->second.bytes ? it->second.bytes : 8), "cudaMallocHost")) return 1;
        if (gfb_copy_d2h(c, key, it->second.host, 0)) return 1;
    }
    *host_ptr = it->second.host;
    return 0;
}

int gfb_check_value(gfb_ctx *c, uint64_t key, size_t index, double *value) {
    if (flush(c)) return 1;
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) return fail("gfb_check_value: unknown key");
    if ((index + 1)*sizeof(double) > it->second.bytes) return fail("gfb_check_value: index out of range");
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (check(cudaMemcpyAsync(value, static_cast<const double *> (it->second.dev) + index, sizeof(double),
                              cudaMemcpyDeviceToHost, c->stream), "check_value")) return 1;
    return check(cudaStreamSynchronize(c->stream), "check_value sync");
}

int gfb_timer_start(gfb_ctx *c) {
    if (flush(c)) return 1;
    return check(cudaEventRecord(c->ev_start, c->stream), "cudaEventRecord");
}
int gfb_timer_stop(gfb_ctx *c, float *ms) {
    if (flush(c)) return 1;
    if (check(cudaEventRecord(c->ev_stop, c->stream), "cudaEventRecord")) return 1;
    if (check(cudaEventSynchronize(c->ev_stop), "cudaEventSynchronize")) return 1;
    return check(cudaEventElapsedTime(ms, c->ev_start, c->ev_stop), "cudaEventElapsedTime");
}
void *gfb_stream(gfb_ctx *c) { return c->stream; }

namespace {
int move_rays(gfb_ctx *c, const unsigned *perm, const uint64_t *keys, int num_keys, size_t n, int scatter) {
    for (int i = 0; i < num_keys; i++) {
        auto it = c->buffers.find(keys[i]);
        if (it == c->buffers.end()) return fail("ray binning: unknown buffer key");
        if (it->second.bytes < n*sizeof(double)) return fail("ray binning: buffer shorter than the ray count");
        double *data = static_cast<double *> (it->second.dev);
        if (gfb_k_permute(c->bin_scratch, data, perm, static_cast<unsigned> (n), scatter, c->sms, c->stream)) {
            return fail("ray binning: permute launch failed");
        }
        if (check(cudaMemcpyAsync(data, c->bin_scratch, n*sizeof(double), cudaMemcpyDeviceToDevice, c->stream), "bin copy back")) return 1;
        c->launches++;
    }
    return 0;
}
}

namespace {
int bin_rays(gfb_ctx *c, const uint64_t *sort_keys, int num_sort_keys, const double lo[2], const double hi[2],
             const unsigned cells01[2], const uint64_t *keys, int num_keys, size_t n) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    const size_t cells = static_cast<size_t> (cells01[0])*(cells01[1] ? cells01[1] : 1u);
    if (cells == 0 || cells > (1u << 24) || !(hi[0] > lo[0]) || (cells01[1] && !(hi[1] > lo[1])) || n == 0 || n > 0xffffffffull) {
        return fail("gfb_bin_rays: bad arguments");
    }
    if (c->bin_capacity < n || c->bin_cells < cells) {
        if (c->binned) return fail("gfb_bin_rays: cannot grow while binned, call gfb_unbin_rays first");
        cudaStreamSynchronize(c->stream);
        if (c->bin_perm) cudaFree(c->bin_perm);
        if (c->bin_step) cudaFree(c->bin_step);
        if (c->bin_work) cudaFree(c->bin_work);
        if (c->bin_scratch) cudaFree(c->bin_scratch);
        if (check(cudaMalloc(&c->bin_perm, n*sizeof(unsigned)), "bin perm")) return 1;
        if (check(cudaMalloc(&c->bin_step, n*sizeof(unsigned)), "bin step")) return 1;
        if (check(cudaMalloc(&c->bin_work, (n + 2*cells)*sizeof(unsigned)), "bin work")) return 1;
        if (check(cudaMalloc(&c->bin_scratch, n*sizeof(double)), "bin scratch")) return 1;
        c->bin_capacity = n;
        c->bin_cells = cells;
    }
    const double *values[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < num_sort_keys; i++) {
        auto it = c->buffers.find(sort_keys[i]);
        if (it == c->buffers.end()) return fail("gfb_bin_rays: unknown sort key");
        if (it->second.bytes < n*sizeof(double)) return fail("gfb_bin_rays: sort array shorter than the ray count");
        values[i] = static_cast<const double *> (it->second.dev);
    }
//  Already binned: sort the current order again and fold the new permutation into the one in force,
//  so that gfb_unbin_rays still restores the caller's order with one scatter per array.
    unsigned *perm = c->binned ? c->bin_step : c->bin_perm;
    if (gfb_k_bin_permutation(values, static_cast<unsigned> (n), lo, hi, cells01, c->bin_work, perm, c->sms, c->stream)) {
        return fail("gfb_bin_rays: launch failed");
    }
    c->launches += 3;
    if (move_rays(c, perm, keys, num_keys, n, 0)) return 1;
    if (c->binned) {
        unsigned *total = c->bin_work + 2*cells;        // the cell_of area is free again
        if (gfb_k_compose(total, c->bin_perm, c->bin_step, static_cast<unsigned> (n), c->sms, c->stream)) return fail("gfb_bin_rays: compose failed");
        if (check(cudaMemcpyAsync(c->bin_perm, total, n*sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream), "bin compose copy")) return 1;
        c->launches++;
    }
    c->binned = true;
    c->bin_n = n;
    return 0;
}
}

int gfb_bin_rays(gfb_ctx *c, uint64_t sort_key, double lo, double hi, unsigned cells,
                 const uint64_t *keys, int num_keys, size_t n) {
    const double lo2[2] = {lo, 0.0}, hi2[2] = {hi, 1.0};
    const unsigned cells2[2] = {cells, 0u};
    return bin_rays(c, &sort_key, 1, lo2, hi2, cells2, keys, num_keys, n);
}
int gfb_bin_rays_rz(gfb_ctx *c, const uint64_t *xyz_keys, const double *lo, const double *hi, const unsigned *cells,
                    const uint64_t *keys, int num_keys, size_t n) {
    if (!cells[1]) return fail("gfb_bin_rays_rz: needs cells in both directions");
    return bin_rays(c, xyz_keys, 3, lo, hi, cells, keys, num_keys, n);
}
int gfb_bin_disorder(gfb_ctx *c, const uint64_t *sort_keys, int num_sort_keys, const double *lo, const double *hi,
                     const unsigned *cells, size_t n, double *fraction) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (num_sort_keys != 1 && num_sort_keys != 3) return fail("gfb_bin_disorder: 1 or 3 sort arrays");
    const double *values[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < num_sort_keys; i++) {
        auto it = c->buffers.find(sort_keys[i]);
        if (it == c->buffers.end() || it->second.bytes < n*sizeof(double)) return fail("gfb_bin_disorder: bad sort key");
        values[i] = static_cast<const double *> (it->second.dev);
    }
    const double lo2[2] = {lo[0], num_sort_keys == 3 ? lo[1] : 0.0}, hi2[2] = {hi[0], num_sort_keys == 3 ? hi[1] : 1.0};
    const unsigned cells2[2] = {cells[0], num_sort_keys == 3 ? cells[1] : 0u};
    unsigned *counter = reinterpret_cast<unsigned *> (c->scratch);
    if (gfb_k_bin_breaks(values, static_cast<unsigned> (n), lo2, hi2, cells2, counter, c->sms, c->stream)) return fail("gfb_bin_disorder: launch failed");
    c->launches++;
    unsigned *host = reinterpret_cast<unsigned *> (c->scratch_host);
    if (check(cudaMemcpyAsync(host, counter, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream), "disorder d2h")) return 1;
    if (check(cudaStreamSynchronize(c->stream), "disorder sync")) return 1;
    *fraction = n > 1 ? static_cast<double> (*host)/static_cast<double> (n - 1) : 0.0;
    return 0;
}
int gfb_is_binned(gfb_ctx *c) { return c->binned ? 1 : 0; }
int gfb_copy_rays_d2h(gfb_ctx *c, uint64_t key, void *destination, size_t n) {
    if (!c->binned) return gfb_copy_d2h(c, key, destination, n*sizeof(double));
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    auto it = c->buffers.find(key);
    if (it == c->buffers.end()) return fail("gfb_copy_rays_d2h: unknown key");
    if (n != c->bin_n || it->second.bytes < n*sizeof(double)) return fail("gfb_copy_rays_d2h: not a whole per-ray array");
    if (gfb_k_permute(c->bin_scratch, static_cast<const double *> (it->second.dev), c->bin_perm, static_cast<unsigned> (n), 1,
                      c->sms, c->stream)) return fail("gfb_copy_rays_d2h: scatter failed");
    c->launches++;
    if (check(cudaMemcpyAsync(destination, c->bin_scratch, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream), "d2h")) return 1;
    return check(cudaStreamSynchronize(c->stream), "d2h sync");
}
int gfb_unbin_rays(gfb_ctx *c, const uint64_t *keys, int num_keys, size_t n) {
    if (!c->binned) return 0;
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (move_rays(c, c->bin_perm, keys, num_keys, n, 1)) return 1;
    c->binned = false;
    return 0;
}

int gfb_host_alloc(size_t bytes, void **host_ptr) {
    return check(cudaMallocHost(host_ptr, bytes ? bytes : 8), "cudaMallocHost") ? 1 : 0;
}
int gfb_host_free(void *host_ptr) {
    return check(cudaFreeHost(host_ptr), "cudaFreeHost") ? 1 : 0;
}

int gfb_deposit(gfb_ctx *c, const double *x, const double *y, const double *z, const double *weight,
                size_t n, double *hist, const double *lo, const double *hi, const int *bins) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    c->launches++;
    if (gfb_k_deposit(x, y, z, weight, n, hist, lo, hi, bins, c->sms, c->stream)) return fail("deposit launch failed");
    return 0;
}

//------------------------------------------------------------------------------
//  Sum of one buffer per device over every device, result in every buffer (the binned
//  power-deposition profile is the one quantity of a sharded ray trace that is reduced).  For the
//  reference's model -- one process, one host thread per device (xrays.cpp:419-527) -- where NCCL
//  communicators are not at hand: a reduce-scatter + all-gather written directly on peer memory.
//    phase 1  device g sums slice g of all G buffers with P2P loads over NVLink (sum_peers_kernel,
//             fixed device order => bit-identical result everywhere) into its scratch;
//    phase 2  device g pulls the G - 1 other summed slices from its peers' scratch (copy engines).
//  Cross-device ordering is by events only; the host never blocks.  Without peer access between a pair
//  (no NVLink/NVSwitch) everything is staged through device 0 with cudaMemcpyPeerAsync.
//------------------------------------------------------------------------------
int gfb_allreduce_sum_f64(gfb_ctx *const *ctxs, int num_ctx, const uint64_t *keys, size_t n) {
    if (num_ctx < 1 || num_ctx > 16) return fail("gfb_allreduce_sum_f64: 1 to 16 contexts");
    if (n == 0) return 0;
    std::vector<double *> data(num_ctx);
    for (int g = 0; g < num_ctx; g++) {
        gfb_ctx *c = ctxs[g];
        if (flush(c)) return 1;
        auto it = c->buffers.find(keys[g]);
        if (it == c->buffers.end()) return fail("gfb_allreduce_sum_f64: unknown key");
        if (it->second.bytes < n*sizeof(double)) return fail("gfb_allreduce_sum_f64: buffer shorter than n");
        data[g] = static_cast<double *> (it->second.dev);
        for (int h = 0; h < g; h++) if (ctxs[h] == c) return fail("gfb_allreduce_sum_f64: a context appears twice");
    }
    if (num_ctx == 1) return 0;
//  Slices of whole 16-byte pairs; the last device takes the remainder.
    const size_t per = ((n + num_ctx - 1)/num_ctx + 1)/2*2;
    auto slice_begin = [&] (const int g) { return std::min(n, per*static_cast<size_t> (g)); };
    auto slice_count = [&] (const int g) { return std::min(n, per*static_cast<size_t> (g + 1)) - slice_begin(g); };
//  First pass: can every device read every other one?  (Decided before any scratch is sized.)
    bool peers = true;
    for (int g = 0; g < num_ctx; g++) {
        gfb_ctx *c = ctxs[g];
        if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
        for (int h = 0; h < num_ctx; h++) {
            if (h == g || ctxs[h]->device == c->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->device, ctxs[h]->device);
            if (!can) { peers = false; continue; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[h]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peers = false;
            cudaGetLastError();
        }
    }
    for (int g = 0; g < num_ctx; g++) {
        gfb_ctx *c = ctxs[g];
        if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
        if (!c->reduce_ready) {
            cudaEventCreateWithFlags(&c->reduce_ready, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->reduce_summed, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->reduce_done, cudaEventDisableTiming);
        }
        const size_t need = peers ? per : n;
        if (c->reduce_capacity < need) {
            cudaStreamSynchronize(c->stream);
            if (c->reduce_scratch) cudaFree(c->reduce_scratch);
            c->reduce_scratch = nullptr;
            c->reduce_capacity = 0;
            if (check(cudaMalloc(&c->reduce_scratch, need*sizeof(double)), "reduce scratch")) return 1;
            c->reduce_capacity = need;
        }
        if (check(cudaEventRecord(c->reduce_ready, c->stream), "reduce ready")) return 1;
    }
    if (!peers) {
//  Staged: every buffer to device 0's scratch in turn, added there, then sent back.
        gfb_ctx *root = ctxs[0];
        if (check(cudaSetDevice(root->device), "cudaSetDevice")) return 1;
        for (int g = 1; g < num_ctx; g++) {
            cudaStreamWaitEvent(root->stream, ctxs[g]->reduce_ready, 0);
            if (check(cudaMemcpyPeerAsync(root->reduce_scratch, root->device, data[g], ctxs[g]->device, n*sizeof(double), root->stream), "reduce stage")) return 1;
            const double *pair[2] = {data[0], root->reduce_scratch};
            if (gfb_k_sum_peers(data[0], pair, 2, n, root->sms, root->stream)) return fail("reduce sum launch failed");
            root->launches++;
        }
        cudaEventRecord(root->reduce_summed, root->stream);
        for (int g = 1; g < num_ctx; g++) {
            gfb_ctx *c = ctxs[g];
            if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
            cudaStreamWaitEvent(c->stream, root->reduce_summed, 0);
            if (check(cudaMemcpyPeerAsync(data[g], c->device, data[0], root->device, n*sizeof(double), c->stream), "reduce broadcast")) return 1;
            cudaEventRecord(c->reduce_done, c->stream);
        }
        if (check(cudaSetDevice(root->device), "cudaSetDevice")) return 1;
        for (int g = 1; g < num_ctx; g++) cudaStreamWaitEvent(root->stream, ctxs[g]->reduce_done, 0);
        return 0;
    }
//  Phase 1: reduce-scatter.
    for (int g = 0; g < num_ctx; g++) {
        gfb_ctx *c = ctxs[g];
        if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
        for (int h = 0; h < num_ctx; h++) if (h != g) cudaStreamWaitEvent(c->stream, ctxs[h]->reduce_ready, 0);
        const size_t count = slice_count(g);
        if (count) {
            std::vector<const double *> src(num_ctx);
            for (int h = 0; h < num_ctx; h++) src[h] = data[h] + slice_begin(g);
            if (gfb_k_sum_peers(c->reduce_scratch, src.data(), num_ctx, count, c->sms, c->stream)) return fail("reduce sum launch failed");
            c->launches++;
        }
        if (check(cudaEventRecord(c->reduce_summed, c->stream), "reduce summed")) return 1;
    }
//  Phase 2: all-gather (own slice device-to-device, the others pulled from the peers).
    for (int g = 0; g < num_ctx; g++) {
        gfb_ctx *c = ctxs[g];
        if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
        for (int h = 0; h < num_ctx; h++) if (h != g) cudaStreamWaitEvent(c->stream, ctxs[h]->reduce_summed, 0);
        for (int h = 0; h < num_ctx; h++) {
            const size_t count = slice_count(h);
            if (!count) continue;
            if (check(cudaMemcpyPeerAsync(data[g] + slice_begin(h), c->device, ctxs[h]->reduce_scratch, ctxs[h]->device,
                                          count*sizeof(double), c->stream), "reduce gather")) return 1;
        }
        if (check(cudaEventRecord(c->reduce_done, c->stream), "reduce done")) return 1;
    }
//  A device's scratch may be rewritten by the next reduction only after every peer has pulled it.
    for (int g = 0; g < num_ctx; g++) {
        if (check(cudaSetDevice(ctxs[g]->device), "cudaSetDevice")) return 1;
        for (int h = 0; h < num_ctx; h++) if (h != g) cudaStreamWaitEvent(ctxs[g]->stream, ctxs[h]->reduce_done, 0);
    }
    return 0;
}

namespace {
int measure_peak(gfb_ctx *c, const int variant, double *tflops, float *milliseconds) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    const int iters = 4096;
    double *scratch = reinterpret_cast<double *> (c->scratch);
    auto launch = [&] (const int n) {
        return variant ? gfb_k_fp64_peak_regs(scratch, n, c->sms, c->stream) : gfb_k_fp64_peak(scratch, n, c->sms, c->stream);
    };
    launch(64);      // warm up
    float best = 1.0e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(c->ev_start, c->stream);
        if (launch(iters)) return fail("peak kernel launch failed");
        cudaEventRecord(c->ev_stop, c->stream);
        if (check(cudaEventSynchronize(c->ev_stop), "peak sync")) return 1;
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop);
        if (ms < best) best = ms;
    }
    const double blocks_per_sm = variant ? 4.0 : 8.0;
    const double flops = 2.0*8.0*16.0*static_cast<double> (iters)*256.0*blocks_per_sm*c->sms;
    *tflops = flops/(best*1.0e-3)/1.0e12;
    if (milliseconds) *milliseconds = best;
    return 0;
}
}
int gfb_measure_fp64_peak(gfb_ctx *c, double *tflops, float *milliseconds) { return measure_peak(c, 0, tflops, milliseconds); }
int gfb_measure_fp64_peak_registers(gfb_ctx *c, double *tflops, float *milliseconds) { return measure_peak(c, 1, tflops, milliseconds); }

int gfb_flush_l2(gfb_ctx *c) {
    if (flush(c)) return 1;
    if (check(cudaSetDevice(c->device), "cudaSetDevice")) return 1;
    if (!c->flush_buffer) {
        c->flush_count = (256ull << 20)/sizeof(double);    // 256 MiB > 126 MB L2
        if (check(cudaMalloc(&c->flush_buffer, c->flush_count*sizeof(double)), "cudaMalloc")) return 1;
    }
    if (gfb_k_fill(c->flush_buffer, c->flush_count, 0.0, c->sms, c->stream)) return fail("fill launch failed");
    return 0;
}

}  // extern "C"
