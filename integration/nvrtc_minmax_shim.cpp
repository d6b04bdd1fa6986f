// TEST/BASELINE INFRASTRUCTURE for integration/_build/ref_driver_cuda (the reference's OWN
// gpu::cuda_context, unmodified, timed on the same B200 as a baseline).
//
// The reference's generated kernels call `min<double>(max<double>(x, 0), n)` (piecewise.hpp:26-65).
// NVRTC of CUDA 12.9 has min/max as plain overloads, not templates, so the reference's kernels do not
// compile there ("device kernel image is invalid").  This file interposes nvrtcCreateProgram and
// prepends the two function templates the kernel text expects; nothing else of the reference's path is
// touched -- same source text otherwise, same NVRTC options, same cuModuleLoadDataEx, same launches.
#include <dlfcn.h>
#include <nvrtc.h>
#include <string>

extern "C" nvrtcResult nvrtcCreateProgram(nvrtcProgram *prog, const char *src, const char *name, int numHeaders,
                                          const char *const *headers, const char *const *includeNames) {
    typedef nvrtcResult (*create_t)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    static create_t real = reinterpret_cast<create_t> (dlsym(RTLD_NEXT, "nvrtcCreateProgram"));
    static const char *prefix =
        "template<typename T> __device__ __forceinline__ T min(const T a, const T b) { return a < b ? a : b; }\n"
        "template<typename T> __device__ __forceinline__ T max(const T a, const T b) { return a > b ? a : b; }\n";
    const std::string patched = std::string(prefix) + (src ? src : "");
    return real(prog, patched.c_str(), name, numHeaders, headers, includeNames);
}
