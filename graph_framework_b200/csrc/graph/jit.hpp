//------------------------------------------------------------------------------
//  jit.hpp -- jit::context for the B200 back end.
//
//  Same public calls as /root/reference/graph_framework/jit.hpp:48-339
//  (add_kernel, add_max_reduction, compile, create_kernel_call, create_max_call,
//  print, check_value, wait, copy_to_device, copy_to_host, get_buffer,
//  max_concurrency, print_source, save_source).  The reference picks a device
//  context class at compile time (jit.hpp:63-71); here there is exactly one:
//  the C-ABI device layer of include/gfb200.h.  No CPU path exists: if the CUDA
//  library or device is missing every call fails loudly.
//
//  Extensions used by solver::rk2/rk4 and solver::newton: add_runge_kutta and
//  add_newton register kernels that use the staged / device-resident skeletons
//  of skeleton.cuh instead of a one-step generic item.
//------------------------------------------------------------------------------
#ifndef gfb_graph_jit_hpp
#define gfb_graph_jit_hpp

#include <fstream>
#include <thread>

#include "../../../include/gfb200.h"
#include "emit.hpp"

namespace jit {
    inline void check(const int rc, const char *what) {
        if (rc) {
            std::cerr << "gfb200 error in " << what << ": " << gfb_last_error() << std::endl;
            std::abort();
        }
    }

    template<typename T=double, bool SAFE_MATH=false>
    class context {
    private:
        gfb_ctx *gpu;
        std::ostringstream source_buffer;
        std::vector<kernel_info> kernels;
        std::vector<std::string> kernel_names;
        std::map<std::string, gfb_kernel *> handles;
        uint64_t next_table_key;
        std::string nvrtc_options;

        static uint64_t key(const graph::leaf_ptr &n) { return reinterpret_cast<uint64_t> (n.get()); }

        kernel_info &find(const std::string &name) {
            for (auto &k : kernels) if (k.name == name) return k;
            std::cerr << "Unknown kernel " << name << std::endl;
            std::abort();
        }

    public:
        static_assert(std::is_same<T, double>::value && !SAFE_MATH, "B200 back end: double, SAFE_MATH=false only.");
///  Emission switches (block size, table staging, reciprocal sharing).
        emit_options options;
        constexpr static size_t random_state_size = 0;

        static size_t max_concurrency() {
            const size_t num = static_cast<size_t> (gfb_device_count());
            std::cout << "Located " << num << " B200 (sm_100a) device" << (num == 1 ? "." : "s.") << std::endl;
            return num;
        }

        context(const size_t index) : gpu(gfb_ctx_create(static_cast<int> (index))), next_table_key(1) {
            if (!gpu) {
                std::cerr << "gfb200: cannot create device context: " << gfb_last_error() << std::endl;
                std::abort();
            }
        }
        ~context() { gfb_ctx_destroy(gpu); }
        context(const context &) = delete;
        context &operator=(const context &) = delete;

        gfb_ctx *device() { return gpu; }
        void set_nvrtc_options(const std::string &o) { nvrtc_options = o; }

///  jit.hpp:118-194.
        void add_kernel(const std::string name,
                        graph::input_nodes<T, SAFE_MATH> inputs,
                        graph::output_nodes<T, SAFE_MATH> outputs,
                        graph::map_nodes<T, SAFE_MATH> setters,
                        graph::shared_random_state<T, SAFE_MATH> state,
                        const size_t size) {
            assert(!state.get() && "Random states are not supported by the B200 ray path.");
            kernel_names.push_back(name);
            kernels.push_back(emit_item(source_buffer, options, kernel_kind::generic, name,
                                        inputs, outputs, setters, size));
        }

///  Device-resident per-ray Newton iteration of the same item shape as
///  solver::newton registers (newton.hpp:42-50).
        void add_newton(const std::string name,
                        graph::input_nodes<T, SAFE_MATH> inputs,
                        graph::output_nodes<T, SAFE_MATH> outputs,
                        graph::map_nodes<T, SAFE_MATH> setters,
                        const size_t size) {
            kernel_names.push_back(name);
            kernels.push_back(emit_item(source_buffer, options, kernel_kind::newton, name,
                                        inputs, outputs, setters, size));
        }

///  Staged Runge-Kutta kernel (solver.hpp:638-665, 811-869).
        void add_runge_kutta(const std::string name, const int order,
                             graph::input_nodes<T, SAFE_MATH> inputs,
                             std::vector<graph::leaf_ptr> evolved,
                             std::vector<graph::leaf_ptr> rates,
                             graph::leaf_ptr time, graph::leaf_ptr dt,
                             graph::leaf_ptr residual, const size_t size) {
            kernel_names.push_back(name);
            kernels.push_back(emit_runge_kutta(source_buffer, options,
                                               order == 2 ? kernel_kind::rk2 : kernel_kind::rk4, name,
                                               inputs, evolved, rates, time, dt, residual, size));
        }

///  jit.hpp:201-203.  The reduction is a static kernel of the library.
        void add_max_reduction(const size_t) {}

        void print_source() { std::cout << std::endl << source_buffer.str() << std::endl; }
        std::string get_source() { return source_buffer.str(); }
        void save_source() {
            const std::string s = source_buffer.str();
            std::ostringstream name;
            name << std::hash<std::string> {} (s) << std::hash<std::thread::id> {} (std::this_thread::get_id()) << ".cu";
            std::ofstream f(name.str());
            f << s;
        }
        const std::vector<kernel_info> &get_kernels() const { return kernels; }

///  jit.hpp:238-245.
        void compile(const bool add_reduction=false) {
            (void)add_reduction;
            std::vector<const char *> names;
            for (auto &n : kernel_names) names.push_back(n.c_str());
            const std::string s = source_buffer.str();
//  Runge-Kutta stage loop: unrolling the four stages lets ptxas drop the unused residual of
//  stages 2-4 and schedule across stages, but quadruples the body.  Measured on B200 (profiles/):
//  it wins for the ~570 statement X-mode body and loses for the ~900 statement cold-plasma body.
            std::string opts = nvrtc_options;
            if (opts.find("GFB_UNROLL_STAGES") == std::string::npos) {
                bool unroll = false;
                for (auto &k : kernels) {
                    if (k.kind == kernel_kind::rk2 || k.kind == kernel_kind::rk4) {
//  A body with a Fourier mode loop (VMEC) carries 26 accumulators through that loop: rolled, the stage loop
//  fits 3 blocks/SM at 168 registers (+12 %, measured); unrolled it needs 230 and gets 2.
                        unroll = k.num_statements <= options.unroll_stages_below && !k.has_mode_loop;
                    }
                }
                opts += unroll ? " -DGFB_UNROLL_STAGES=1" : " -DGFB_UNROLL_STAGES=0";
            }
            check(gfb_compile(gpu, s.c_str(), names.data(), static_cast<int> (names.size()), opts.c_str()), "compile");
            handles.clear();
        }

///  jit.hpp:257-265.  Allocates and uploads every buffer the kernel touches
///  (cuda_context.hpp:330-364) and returns the deferred-launch callable.
        std::function<void(void)> create_kernel_call(const std::string kernel_name,
                                                     graph::input_nodes<T, SAFE_MATH> inputs,
                                                     graph::output_nodes<T, SAFE_MATH> outputs,
                                                     graph::shared_random_state<T, SAFE_MATH> state,
                                                     const size_t num_rays) {
            (void)inputs; (void)outputs; (void)state;
            gfb_kernel *handle = get_kernel(kernel_name, num_rays);
            return [handle] () { check(gfb_kernel_run(handle), "kernel run"); };
        }

        gfb_kernel *get_kernel(const std::string &kernel_name, const size_t num_rays) {
            auto found = handles.find(kernel_name);
            if (found != handles.end()) return found->second;
            kernel_info &k = find(kernel_name);
            std::vector<uint64_t> keys;
            for (auto &in : k.inputs) {
                check(gfb_buffer(gpu, key(in), in->size()*sizeof(T), in->data(), nullptr), "input buffer");
                keys.push_back(key(in));
            }
            for (auto &out : k.outputs) {
                check(gfb_buffer(gpu, key(out), num_rays*sizeof(T), nullptr, nullptr), "output buffer");
                keys.push_back(key(out));
            }
            for (auto &g : k.groups) {
                if (g.alias_input >= 0) continue;           // index_1D/2D read a kernel input in place
                const uint64_t tk = 0x8000000000000000ull | next_table_key++;
                check(gfb_buffer(gpu, tk, g.bytes(), g.packed.data(), nullptr), "table buffer");
                keys.push_back(tk);
            }
            bool can_repeat = false;
            for (const bool w : k.input_written) can_repeat = can_repeat || w;
//  A kernel that gathers from an array it also rewrites needs a grid-wide boundary between steps:
//  mode 2 = one launch per step.
            const int repeat_mode = k.indexed_written ? 2 : (can_repeat ? 1 : 0);
            gfb_kernel *handle = nullptr;
            check(gfb_kernel_create(gpu, kernel_name.c_str(), keys.data(), static_cast<int> (keys.size()),
                                    num_rays, options.block_size, k.smem_bytes,
                                    k.kind == kernel_kind::generic ? 0 : (k.kind == kernel_kind::newton ? 2 : 1),
                                    repeat_mode, &handle), "kernel create");
            handles[kernel_name] = handle;
            return handle;
        }

///  jit.hpp:274-277.
        std::function<T(void)> create_max_call(graph::shared_leaf<T, SAFE_MATH> &argument,
                                               std::function<void(void)> run) {
            gfb_ctx *g = gpu;
            const uint64_t k = key(argument);
            return [g, k, run] () {
                run();
                void *p = nullptr;
                size_t bytes = 0;
                check(gfb_buffer_lookup(g, k, &p, &bytes), "max lookup");
                double result = 0.0;
                check(gfb_max(g, k, bytes/sizeof(T), &result), "max");
                return result;
            };
        }

        void print(const size_t index, const graph::output_nodes<T, SAFE_MATH> &nodes) {
            for (auto &n : nodes) std::cout << check_value(index, n) << " ";
            std::cout << std::endl;
        }
        T check_value(const size_t index, const graph::shared_leaf<T, SAFE_MATH> &node) {
            double v = 0.0;
            check(gfb_check_value(gpu, key(node), index, &v), "check_value");
            return v;
        }
        void wait() { check(gfb_wait(gpu), "wait"); }
        void copy_to_device(graph::shared_leaf<T, SAFE_MATH> &node, T *source) {
            check(gfb_copy_h2d(gpu, key(node), source, 0), "copy_to_device");
        }
        void copy_to_host(graph::shared_leaf<T, SAFE_MATH> &node, T *destination) {
            check(gfb_copy_d2h(gpu, key(node), destination, 0), "copy_to_host");
        }
        T *get_buffer(graph::shared_leaf<T, SAFE_MATH> &node) {
            void *p = nullptr;
            check(gfb_host_ptr(gpu, key(node), &p), "get_buffer");
            return static_cast<T *> (p);
        }
///  Raw device pointer of a node's buffer (for callers that share memory with torch).
        void *device_pointer(const graph::shared_leaf<T, SAFE_MATH> &node) {
            void *p = nullptr;
            size_t bytes = 0;
            check(gfb_buffer_lookup(gpu, key(node), &p, &bytes), "device_pointer");
            return p;
        }
    };
}

#endif /* gfb_graph_jit_hpp */
