// gpu::cpu_context stand-in: g++ + dlopen instead of in-process clang/LLVM ORC.
// TEST INFRASTRUCTURE (oracle/_ref build only) -- not product code.
//
// The reference's CPU back end (`/root/reference/graph_framework/cpu_context.hpp`)
// needs clang+LLVM libraries that this image does not have.  Everything above the
// back end (graph, reductions, autodiff, code emission by the nodes) is compiled
// UNMODIFIED from /root/reference; this header only supplies the 18 members that
// jit::context calls (jit.hpp:80-338).  The kernel text it writes has the same
// shape as the reference's (cpu_context.hpp:400-584: one `extern "C"` function per
// kernel taking `map<size_t, T*>&`, a serial loop over rays, registers named by
// jit::to_string) so the arithmetic the oracle executes is the reference's own;
// only the compiler differs (g++ -O3 -ffast-math instead of clang -O3 -ffast-math).
//
// Build with:  -Dcpu_context_h -include gxx_cpu_context.hpp
#ifndef GFB_ORACLE_GXX_CPU_CONTEXT_HPP
#define GFB_ORACLE_GXX_CPU_CONTEXT_HPP

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <thread>
#include <unistd.h>
#include <unordered_set>

#include "random.hpp"

namespace gpu {
inline std::atomic<int> &gxx_unit_counter() { static std::atomic<int> c{0}; return c; }

template<jit::float_scalar T, bool SAFE_MATH=false>
class cpu_context {
private:
    void *handle = nullptr;
    std::map<graph::leaf_node<T, SAFE_MATH> *, std::vector<T>> device_side;
    std::map<graph::leaf_node<T, SAFE_MATH> *, std::vector<T>> host_side;

    static std::string ptr_key(const void *p) {
        return std::to_string(reinterpret_cast<size_t> (p));
    }
    void store(std::ostringstream &s, const std::string &lhs, const std::string &reg) {
        s << "        " << lhs << "[i] = ";
        if constexpr (SAFE_MATH && !jit::complex_scalar<T>) {
            s << "isnan(" << reg << ") ? 0.0 : " << reg;
        } else {
            s << reg;
        }
        s << ";" << std::endl;
    }

public:
    constexpr static size_t random_state_size = 1;
    int remaining_const_memory = 0;

    static size_t max_concurrency() {
        if (const char *e = std::getenv("GFB_ORACLE_THREADS")) return std::max(1, std::atoi(e));
        return std::thread::hardware_concurrency();
    }
    static std::string device_type() { return "CPU(g++)"; }

    cpu_context(const size_t) {}
    ~cpu_context() { if (handle) dlclose(handle); }

    void compile(const std::string kernel_source,
                 std::vector<std::string> names,
                 const bool add_reduction=false) {
        (void)names; (void)add_reduction;
        const char *tmp = std::getenv("TMPDIR");
        std::ostringstream stem;
        stem << (tmp ? tmp : "/tmp") << "/gfb_oracle_" << getpid() << "_" << gxx_unit_counter()++;
        const std::string src = stem.str() + ".cpp", lib = stem.str() + ".so";
        { std::ofstream f(src); f << kernel_source; }
        const char *inc = std::getenv("GFB_REFERENCE_INCLUDE");
        const char *keep = std::getenv("GFB_ORACLE_KEEP_SOURCE");
        std::ostringstream cmd;
        cmd << "g++ -std=gnu++2a -shared -fPIC -O3 -ffast-math -w ";
        if (inc) cmd << "-I" << inc << " ";
        cmd << src << " -o " << lib;
        if (std::system(cmd.str().c_str()) != 0) {
            std::cerr << "oracle kernel compile failed: " << cmd.str() << std::endl;
            std::exit(-1);
        }
        handle = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) { std::cerr << dlerror() << std::endl; std::exit(-1); }
        if (keep) {
            std::cerr << "oracle kernel source kept: " << src << std::endl;
        } else {
            unlink(src.c_str());
        }
        unlink(lib.c_str());
    }

    std::function<void(void)> create_kernel_call(const std::string kernel_name,
                                                 graph::input_nodes<T, SAFE_MATH> inputs,
                                                 graph::output_nodes<T, SAFE_MATH> outputs,
                                                 graph::shared_random_state<T, SAFE_MATH> state,
                                                 const size_t num_rays,
                                                 const jit::texture1d_list &,
                                                 const jit::texture2d_list &) {
        using plain_fn = void (*)(std::map<size_t, T *> &);
        using rand_fn = void (*)(std::map<size_t, T *> &, typename graph::random_state_node<T, SAFE_MATH>::mt_state *);
        void *sym = dlsym(handle, kernel_name.c_str());
        if (!sym) { std::cerr << "oracle: missing kernel " << kernel_name << std::endl; std::exit(-1); }
        std::map<size_t, T *> args;
        for (auto &in : inputs) {
            if (!device_side.contains(in.get())) {
                backend::buffer<T> b = in->evaluate();
                device_side[in.get()] = std::vector<T> (b.data(), b.data() + b.size());
            }
            args[reinterpret_cast<size_t> (in.get())] = device_side[in.get()].data();
        }
        for (auto &out : outputs) {
            if (!device_side.contains(out.get())) device_side[out.get()] = std::vector<T> (num_rays);
            args[reinterpret_cast<size_t> (out.get())] = device_side[out.get()].data();
        }
        if (state.get()) {
            auto fn = reinterpret_cast<rand_fn> (sym);
            return [fn, args, state] () mutable { fn(args, state->data()); };
        }
        auto fn = reinterpret_cast<plain_fn> (sym);
        return [fn, args] () mutable { fn(args); };
    }

    std::function<T(void)> create_max_call(graph::shared_leaf<T, SAFE_MATH> &argument,
                                           std::function<void(void)> run) {
        std::vector<T> *buf = &device_side[argument.get()];
        return [run, buf] () mutable {
            run();
            if constexpr (jit::complex_scalar<T>) {
                return *std::max_element(buf->cbegin(), buf->cend(),
                                         [] (const T a, const T b) { return std::abs(a) < std::abs(b); });
            } else {
                return *std::max_element(buf->cbegin(), buf->cend());
            }
        };
    }

    void wait() {
        for (auto &kv : host_side) kv.second = device_side[kv.first];
    }
    void print_results(const size_t index, const graph::output_nodes<T, SAFE_MATH> &nodes) {
        for (auto &n : nodes) {
            const T v = device_side[n.get()][index];
            if constexpr (jit::complex_scalar<T>) std::cout << std::real(v) << " " << std::imag(v) << " ";
            else std::cout << v << " ";
        }
        std::cout << std::endl;
    }
    T check_value(const size_t index, const graph::shared_leaf<T, SAFE_MATH> &node) {
        return device_side[node.get()][index];
    }
    void copy_to_device(graph::shared_leaf<T, SAFE_MATH> node, T *source) {
        auto &d = device_side[node.get()];
        std::memcpy(d.data(), source, sizeof(T)*d.size());
    }
    void copy_to_host(const graph::shared_leaf<T, SAFE_MATH> node, T *destination) {
        auto &d = device_side[node.get()];
        std::memcpy(destination, d.data(), sizeof(T)*d.size());
    }
    T *get_buffer(graph::shared_leaf<T, SAFE_MATH> &node) {
        if (!host_side.contains(node.get())) host_side[node.get()] = device_side[node.get()];
        return host_side[node.get()].data();
    }

    // ---- emission -----------------------------------------------------------
    void create_header(std::ostringstream &s) {
        s << "#include <map>\n#include <array>\n#include <cstdint>\n";
        if (jit::complex_scalar<T>) s << "#include <complex>\n#include <special_functions.hpp>\n";
        else s << "#include <cmath>\n";
        s << "using namespace std;" << std::endl;
    }

    void create_kernel_prefix(std::ostringstream &s,
                              const std::string name,
                              graph::input_nodes<T, SAFE_MATH> &inputs,
                              graph::output_nodes<T, SAFE_MATH> &outputs,
                              graph::shared_random_state<T, SAFE_MATH> state,
                              const size_t size,
                              const std::vector<bool> &is_constant,
                              jit::register_map &registers,
                              const jit::register_usage &,
                              jit::texture1d_list &,
                              jit::texture2d_list &) {
        const std::string type = jit::get_type_string<T> ();
        s << std::endl << "extern \"C\" void " << name << "(" << std::endl
          << "    map<size_t, " << type << " *> &args";
        if (state.get()) s << "," << std::endl << "    mt_state *" << jit::to_string('s', state.get());
        s << ") {" << std::endl;
        std::unordered_set<void *> seen;
        for (size_t i = 0; i < inputs.size(); i++) {
            if (seen.insert(inputs[i].get()).second) {
                s << "    " << (is_constant[i] ? "const " : "") << type << " *"
                  << jit::to_string('v', inputs[i].get()) << " = args[" << ptr_key(inputs[i].get()) << "];" << std::endl;
            }
        }
        for (auto &out : outputs) {
            if (seen.insert(out.get()).second) {
                s << "    " << type << " *" << jit::to_string('o', out.get())
                  << " = args[" << ptr_key(out.get()) << "];" << std::endl;
            }
        }
        if (state.get()) {
            registers[state.get()] = jit::to_string('r', state.get());
            s << "    mt_state &" << registers[state.get()] << " = "
              << jit::to_string('s', state.get()) << "[0];" << std::endl;
        }
        s << "    for (size_t i = 0; i < " << size << "; i++) {" << std::endl;
        for (auto &in : inputs) {
            registers[in.get()] = jit::to_string('r', in.get());
            s << "        const " << type << " " << registers[in.get()] << " = "
              << jit::to_string('v', in.get()) << "[i]; // " << in->get_symbol() << std::endl;
        }
    }

    void create_kernel_postfix(std::ostringstream &s,
                               graph::output_nodes<T, SAFE_MATH> &outputs,
                               graph::map_nodes<T, SAFE_MATH> &setters,
                               graph::shared_random_state<T, SAFE_MATH>,
                               jit::register_map &registers,
                               jit::register_map &indices,
                               const jit::register_usage &usage) {
        std::unordered_set<void *> written;
        for (auto &[out, in] : setters) {
            if (!out->is_match(in)) {
                auto a = out->compile(s, registers, indices, usage);
                store(s, jit::to_string('v', in.get()), registers[a.get()]);
                written.insert(out.get());
            }
        }
        for (auto &out : outputs) {
            if (!graph::variable_cast(out).get() && !written.contains(out.get())) {
                auto a = out->compile(s, registers, indices, usage);
                store(s, jit::to_string('o', out.get()), registers[a.get()]);
                written.insert(out.get());
            }
        }
        s << "    }" << std::endl << "}" << std::endl;
    }

    void create_reduction(std::ostringstream &, const size_t) {}
};
}  // namespace gpu

#endif
