//------------------------------------------------------------------------------
//  b200_context.hpp -- drop-in device context for the UNMODIFIED reference headers.
//
//  The reference selects its device back end with one typedef (jit.hpp:63-71) and talks to
//  it through 18 members (cuda_context.hpp:111-1004).  This header provides those members on
//  top of the C ABI of libgfb200.so (include/gfb200.h), so that every reference program --
//  xrays, xrays_bench, xkorc, the C and Fortran bindings -- runs on the B200 device layer
//  without touching any other reference file:
//
//      g++ -std=c++20 -DUSE_CUDA -Dcuda_context_h -include integration/b200_context.hpp ...
//          (plus a cpu_context stand-in, because jit.hpp:20 includes cpu_context.hpp unconditionally)
//
//  or, as a one-line patch to jit.hpp, `#include "b200_context.hpp"` next to cuda_context.hpp.
//
//  What changes relative to gpu::cuda_context:
//    * NVRTC straight to an sm_100a cubin (no generic PTX, no 128-register cap);
//    * explicit device buffers + pinned host mirrors instead of managed memory;
//    * the emitted kernel keeps the work item's variables in registers and loops over
//      `steps`: consecutive run()s of one kernel are deferred by the device layer and issued as
//      ONE multi-step launch (state is only observable at wait/copy/check_value/max);
//    * spline tables go to global memory (L1/L2), never to __constant__ (divergent indices
//      serialise in the constant cache): remaining_const_memory = 0;
//    * a grid-wide max reduction instead of one block of 1024 threads.
//  The expression body is still the reference's own (its nodes' compile() methods write it), so
//  this path shows what the boundary alone buys; the full back end (graph_framework_b200/csrc/graph)
//  additionally replaces the front end.
//
//  Only real double is supported (the FP64 ray path).
//------------------------------------------------------------------------------
#ifndef gfb_b200_context_hpp
#define gfb_b200_context_hpp
#ifndef cuda_context_h
#define cuda_context_h
#endif

#include <cstring>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

#include "random.hpp"
#include "../include/gfb200.h"
#include "text_passes.hpp"

namespace gpu {
    template<jit::float_scalar T, bool SAFE_MATH=false>
    class b200_context {
    private:
        gfb_ctx *ctx;
        std::map<std::string, gfb_kernel *> kernels;
///  Pointer-slot order (distinct inputs, then distinct outputs) of the kernel being emitted;
///  jit::context::add_kernel calls prefix, the nodes and postfix back to back (jit.hpp:164-181).
        std::vector<std::string> slots;

        static void check(const int rc, const char *what) {
            if (rc) {
                std::cerr << "b200_context: " << what << ": " << gfb_last_error() << std::endl;
#ifndef NDEBUG
                std::abort();
#endif
            }
        }
        static uint64_t key(const void *p) { return reinterpret_cast<uint64_t> (p); }

    public:
///  cuda_context.hpp:111.
        constexpr static size_t random_state_size = 1;
///  cuda_context.hpp:114.  Zero: tables are never placed in __constant__ memory.
        int remaining_const_memory;

///  cuda_context.hpp:121-126.
        static size_t max_concurrency() { return static_cast<size_t> (gfb_device_count()); }
///  cuda_context.hpp:130-132.
        static std::string device_type() { return "B200 (sm_100a)"; }

///  cuda_context.hpp:139-148.
        b200_context(const size_t index) : ctx(gfb_ctx_create(static_cast<int> (index))), remaining_const_memory(0) {
            static_assert(std::is_same<T, double>::value, "b200_context implements the FP64 path only.");
            if (!ctx) {
                std::cerr << "b200_context: " << gfb_last_error() << std::endl;
                std::exit(1);
            }
        }
///  cuda_context.hpp:153-185.
        ~b200_context() { gfb_ctx_destroy(ctx); }

///  cuda_context.hpp:194-302.
        void compile(const std::string kernel_source, std::vector<std::string> names, const bool add_reduction=false) {
            (void)add_reduction;
            std::vector<const char *> cnames;
            for (auto &n : names) cnames.push_back(n.c_str());
            const std::string text = gfb_text::share_reciprocals(kernel_source);
            check(gfb_compile(ctx, text.c_str(), cnames.data(), static_cast<int> (cnames.size()),
                              "--device-as-default-execution-space"), "compile");
        }

///  cuda_context.hpp:316-531.
        std::function<void(void)> create_kernel_call(const std::string kernel_name,
                                                     graph::input_nodes<T, SAFE_MATH> inputs,
                                                     graph::output_nodes<T, SAFE_MATH> outputs,
                                                     graph::shared_random_state<T, SAFE_MATH> state,
                                                     const size_t num_rays,
                                                     const jit::texture1d_list &, const jit::texture2d_list &) {
            if (state.get()) {
                std::cerr << "b200_context: random states are not on the FP64 ray path." << std::endl;
                std::exit(1);
            }
            std::vector<uint64_t> keys;
            std::unordered_set<void *> seen;
            for (auto &in : inputs) {
                if (!seen.insert(in.get()).second) continue;
                backend::buffer<T> b = in->evaluate();
                check(gfb_buffer(ctx, key(in.get()), b.size()*sizeof(T), b.data(), nullptr), "input buffer");
                keys.push_back(key(in.get()));
            }
            for (auto &out : outputs) {
                if (!seen.insert(out.get()).second) continue;
                check(gfb_buffer(ctx, key(out.get()), num_rays*sizeof(T), nullptr, nullptr), "output buffer");
                keys.push_back(key(out.get()));
            }
            gfb_kernel *k = nullptr;
            check(gfb_kernel_create(ctx, kernel_name.c_str(), keys.data(), static_cast<int> (keys.size()),
                                    num_rays, 128, 0, 0, 1, &k), "kernel create");
            kernels[kernel_name] = k;
            return [k] () { check(gfb_kernel_run(k), "run"); };
        }

///  cuda_context.hpp:540-576.
        std::function<T(void)> create_max_call(graph::shared_leaf<T, SAFE_MATH> &argument,
                                               std::function<void(void)> run) {
            gfb_ctx *c = ctx;
            const uint64_t k = key(argument.get());
            return [c, k, run] () mutable {
                run();
                void *p = nullptr;
                size_t bytes = 0;
                check(gfb_buffer_lookup(c, k, &p, &bytes), "max lookup");
                double result = 0.0;
                check(gfb_max(c, k, bytes/sizeof(T), &result), "max");
                return static_cast<T> (result);
            };
        }

///  cuda_context.hpp:581-584.
        void wait() { check(gfb_wait(ctx), "wait"); }
///  cuda_context.hpp:592-606.
        void print_results(const size_t index, const graph::output_nodes<T, SAFE_MATH> &nodes) {
            for (auto &out : nodes) std::cout << check_value(index, out) << " ";
            std::cout << std::endl;
        }
///  cuda_context.hpp:613-617.
        T check_value(const size_t index, const graph::shared_leaf<T, SAFE_MATH> &node) {
            double v = 0.0;
            check(gfb_check_value(ctx, key(node.get()), index, &v), "check_value");
            return static_cast<T> (v);
        }
///  cuda_context.hpp:625-631.
        void copy_to_device(graph::shared_leaf<T, SAFE_MATH> node, T *source) {
            check(gfb_copy_h2d(ctx, key(node.get()), source, 0), "copy_to_device");
        }
///  cuda_context.hpp:638-643.
        void copy_to_host(graph::shared_leaf<T, SAFE_MATH> node, T *destination) {
            check(gfb_copy_d2h(ctx, key(node.get()), destination, 0), "copy_to_host");
        }
///  cuda_context.hpp:1002-1004.
        T *get_buffer(graph::shared_leaf<T, SAFE_MATH> &node) {
            void *p = nullptr;
            check(gfb_host_ptr(ctx, key(node.get()), &p), "get_buffer");
            return static_cast<T *> (p);
        }

//  ---- emission (writes into jit::context's stream) ------------------------------
///  cuda_context.hpp:650-706.  The skeleton text (gfb_args, launch bounds macro) is prepended by
///  the device layer; nothing else is needed for real double.
        void create_header(std::ostringstream &source_buffer) {
            source_buffer << "typedef unsigned int uint32_t;" << std::endl
                          << "typedef unsigned short uint16_t;" << std::endl
//  The reference's index expressions call min<double>(max<double>(x, 0), n) (piecewise.hpp:26-65).
//  NVRTC 12.9 has no such templates (the reference's own cuda_context fails to compile its EFIT
//  kernels on this toolchain for that reason), so they are supplied here.
                          << "template<typename T> __device__ __forceinline__ T min(const T a, const T b) { return a < b ? a : b; }" << std::endl
                          << "template<typename T> __device__ __forceinline__ T max(const T a, const T b) { return a > b ? a : b; }" << std::endl;
        }

///  cuda_context.hpp:713-849.  Opens the kernel, loads every argument into a register and
///  opens the fused step loop; the reference's nodes then write their statements.
        void create_kernel_prefix(std::ostringstream &source_buffer,
                                  const std::string name,
                                  graph::input_nodes<T, SAFE_MATH> &inputs,
                                  graph::output_nodes<T, SAFE_MATH> &outputs,
                                  graph::shared_random_state<T, SAFE_MATH> state,
                                  const size_t size,
                                  const std::vector<bool> &is_constant,
                                  jit::register_map &registers,
                                  const jit::register_usage &usage,
                                  jit::texture1d_list &, jit::texture2d_list &) {
            (void)state; (void)size; (void)usage;
            source_buffer << std::endl
                          << "extern \"C\" __global__ void __launch_bounds__(128, GFB_MIN_BLOCKS) " << name
                          << "(const __grid_constant__ gfb_args a) {" << std::endl
                          << "    const unsigned long long index = static_cast<unsigned long long> (blockIdx.x)*blockDim.x + threadIdx.x;" << std::endl
                          << "    if (index >= a.n) return;" << std::endl;
            std::unordered_set<void *> seen;
            slots.clear();
            for (size_t i = 0; i < inputs.size(); i++) {
                if (!seen.insert(inputs[i].get()).second) continue;
                const std::string slot = std::to_string(slots.size());
                source_buffer << "    double " << jit::to_string('v', inputs[i].get()) << " = "
                              << (is_constant[i] ? "__ldg(a.ptr[" + slot + "] + index)" : "a.ptr[" + slot + "][index]")
                              << "; // " << inputs[i]->get_symbol() << std::endl;
                slots.push_back(jit::to_string('v', inputs[i].get()));
            }
            for (auto &out : outputs) {
                if (!seen.insert(out.get()).second) continue;
                source_buffer << "    double " << jit::to_string('o', out.get()) << " = 0.0;" << std::endl;
                slots.push_back(jit::to_string('o', out.get()));
            }
            source_buffer << "#pragma unroll 1" << std::endl
                          << "    for (unsigned step = 0; step < a.steps; step++) {" << std::endl;
            for (auto &input : inputs) {
                registers[input.get()] = jit::to_string('r', input.get());
                source_buffer << "        const double " << registers[input.get()] << " = "
                              << jit::to_string('v', input.get()) << ";" << std::endl;
            }
        }

///  cuda_context.hpp:862-946.  Applies the setters to the register copies, closes the step loop
///  and stores written inputs and outputs once.
        void create_kernel_postfix(std::ostringstream &source_buffer,
                                   graph::output_nodes<T, SAFE_MATH> &outputs,
                                   graph::map_nodes<T, SAFE_MATH> &setters,
                                   graph::shared_random_state<T, SAFE_MATH> state,
                                   jit::register_map &registers,
                                   jit::register_map &indices,
                                   const jit::register_usage &usage) {
            (void)state;
            std::vector<std::string> stores;
            std::unordered_set<void *> done;
            for (auto &[out, in] : setters) {
                if (out->is_match(in)) continue;
                auto a = out->compile(source_buffer, registers, indices, usage);
                source_buffer << "        " << jit::to_string('v', in.get()) << " = " << registers[a.get()] << ";" << std::endl;
                stores.push_back(jit::to_string('v', in.get()));
                done.insert(out.get());
            }
            for (auto &out : outputs) {
                if (graph::variable_cast(out).get() || done.contains(out.get())) continue;
                auto a = out->compile(source_buffer, registers, indices, usage);
                source_buffer << "        " << jit::to_string('o', out.get()) << " = " << registers[a.get()] << ";" << std::endl;
                done.insert(out.get());
            }
            source_buffer << "    }" << std::endl;
            auto slot_of = [this] (const std::string &n) {
                for (size_t i = 0; i < slots.size(); i++) if (slots[i] == n) return i;
                return slots.size();
            };
            for (auto &v : stores) {
                source_buffer << "    a.ptr[" << slot_of(v) << "][index] = " << v << ";" << std::endl;
            }
            std::unordered_set<std::string> out_done;
            for (auto &out : outputs) {
                if (graph::variable_cast(out).get()) continue;
                const std::string n = jit::to_string('o', out.get());
                if (!out_done.insert(n).second) continue;
                source_buffer << "    a.ptr[" << slot_of(n) << "][index] = " << n << ";" << std::endl;
            }
            source_buffer << "}" << std::endl;
        }

///  cuda_context.hpp:954-995: nothing to emit, the reduction is a static kernel of libgfb200.
        void create_reduction(std::ostringstream &, const size_t) {}

    };

///  With -DUSE_CUDA the reference instantiates gpu::cuda_context (jit.hpp:65).
    template<jit::float_scalar T, bool SAFE_MATH=false>
    using cuda_context = b200_context<T, SAFE_MATH>;
}

#endif /* gfb_b200_context_hpp */
