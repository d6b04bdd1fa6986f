//------------------------------------------------------------------------------
//  workflow.hpp -- work items and the workflow manager.
//
//  Same classes and calls as /root/reference/graph_framework/workflow.hpp
//  (work_item :22-76, loop_item :84-120, converge_item :128-206, manager
//  :215-425).  Launches are deferred by the device layer, so a loop_item or a
//  host loop around run() costs one multi-step launch, not one launch per step.
//
//  Extensions: runge_kutta_item and newton_item use the staged and
//  device-resident skeletons.
//------------------------------------------------------------------------------
#ifndef gfb_graph_workflow_hpp
#define gfb_graph_workflow_hpp

#include <limits>

#include "jit.hpp"

namespace workflow {
    template<typename T=double, bool SAFE_MATH=false>
    class work_item {
    protected:
        const std::string kernel_name;
        const size_t kernel_size;
        graph::input_nodes<T, SAFE_MATH> inputs;
        graph::output_nodes<T, SAFE_MATH> outputs;
        graph::shared_random_state<T, SAFE_MATH> state;
        std::function<void(void)> kernel;
        struct no_emit {};
        work_item(no_emit, graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                  const std::string name, const size_t size) :
        kernel_name(name), kernel_size(size), inputs(in), outputs(out) {}
    public:
        work_item(graph::input_nodes<T, SAFE_MATH> in,
                  graph::output_nodes<T, SAFE_MATH> out,
                  graph::map_nodes<T, SAFE_MATH> maps,
                  graph::shared_random_state<T, SAFE_MATH> state,
                  const std::string name, const size_t size,
                  jit::context<T, SAFE_MATH> &context) :
        kernel_name(name), kernel_size(size), inputs(in), outputs(out), state(state) {
            context.add_kernel(name, in, out, maps, state, size);
        }
        virtual ~work_item() {}
        virtual void create_kernel_call(jit::context<T, SAFE_MATH> &context) {
            kernel = context.create_kernel_call(kernel_name, inputs, outputs, state, kernel_size);
        }
        virtual void run() { kernel(); }
    };

    template<typename T=double, bool SAFE_MATH=false>
    class loop_item final : public work_item<T, SAFE_MATH> {
        const size_t num_iterations;
    public:
        loop_item(graph::input_nodes<T, SAFE_MATH> inputs,
                  graph::output_nodes<T, SAFE_MATH> outputs,
                  graph::map_nodes<T, SAFE_MATH> maps,
                  graph::shared_random_state<T, SAFE_MATH> state,
                  const std::string name, const size_t size,
                  jit::context<T, SAFE_MATH> &context,
                  const size_t iterations) :
        work_item<T, SAFE_MATH> (inputs, outputs, maps, state, name, size, context),
        num_iterations(iterations) {}
        virtual void run() {
            for (size_t i = 0; i < num_iterations; i++) work_item<T, SAFE_MATH>::run();
        }
    };

///  Host-driven convergence loop with the reference's exact stopping rule on the
///  ENSEMBLE maximum (workflow.hpp:179-205).  Kept for parity; the ray path uses
///  newton_item below.
    template<typename T=double, bool SAFE_MATH=false>
    class converge_item final : public work_item<T, SAFE_MATH> {
    private:
        std::function<T(void)> max_kernel;
        const T tolerance;
        const size_t max_iterations;
    public:
        converge_item(graph::input_nodes<T, SAFE_MATH> inputs,
                      graph::output_nodes<T, SAFE_MATH> outputs,
                      graph::map_nodes<T, SAFE_MATH> maps,
                      graph::shared_random_state<T, SAFE_MATH> state,
                      const std::string name, const size_t size,
                      jit::context<T, SAFE_MATH> &context,
                      const T tol=1.0E-30, const size_t max_iter=1000) :
        work_item<T, SAFE_MATH> (inputs, outputs, maps, state, name, size, context),
        tolerance(tol), max_iterations(max_iter) {
            context.add_max_reduction(size);
        }
        virtual void create_kernel_call(jit::context<T, SAFE_MATH> &context) {
            work_item<T, SAFE_MATH>::create_kernel_call(context);
            max_kernel = context.create_max_call(this->outputs.back(), this->kernel);
        }
        virtual void run() {
            size_t iterations = 0;
            T max_residual = max_kernel();
            T last_max = std::numeric_limits<T>::max();
            T off_last_max = std::numeric_limits<T>::max();
            while (std::abs(max_residual) > std::abs(tolerance)                &&
                   std::abs(last_max - max_residual) > std::abs(tolerance)     &&
                   std::abs(off_last_max - max_residual) > std::abs(tolerance) &&
                   iterations++ < max_iterations) {
                last_max = max_residual;
                if (!(iterations%2)) off_last_max = max_residual;
                max_residual = max_kernel();
            }
            if (iterations > max_iterations) {
                std::cerr << "Workitem failed to converge with in given iterations." << std::endl;
                std::cerr << "Minimum residual reached: " << max_residual << std::endl;
            }
        }
    };

///  Newton iteration that never leaves the device: one launch, every ray
///  iterates to its own convergence (skeleton.cuh newton_item).
    template<typename T=double, bool SAFE_MATH=false>
    class newton_item final : public work_item<T, SAFE_MATH> {
    private:
        const T tolerance;
        const size_t max_iterations;
        gfb_kernel *handle;
    public:
        newton_item(graph::input_nodes<T, SAFE_MATH> inputs,
                    graph::output_nodes<T, SAFE_MATH> outputs,
                    graph::map_nodes<T, SAFE_MATH> maps,
                    const std::string name, const size_t size,
                    jit::context<T, SAFE_MATH> &context,
                    const T tol=1.0E-30, const size_t max_iter=1000) :
        work_item<T, SAFE_MATH> (typename work_item<T, SAFE_MATH>::no_emit(), inputs, outputs, name, size),
        tolerance(tol), max_iterations(max_iter), handle(nullptr) {
            context.add_newton(name, inputs, outputs, maps, size);
        }
        virtual void create_kernel_call(jit::context<T, SAFE_MATH> &context) {
            handle = context.get_kernel(this->kernel_name, this->kernel_size);
            jit::check(gfb_kernel_set_scalar(handle, 0, tolerance), "newton tolerance");
        }
        virtual void run() {
            jit::check(gfb_kernel_launch(handle, static_cast<unsigned> (max_iterations)), "newton launch");
        }
    };

///  Staged Runge-Kutta item (skeleton.cuh runge_kutta).
    template<typename T=double, bool SAFE_MATH=false>
    class runge_kutta_item final : public work_item<T, SAFE_MATH> {
    public:
        runge_kutta_item(const int order,
                         graph::input_nodes<T, SAFE_MATH> inputs,
                         std::vector<graph::leaf_ptr> evolved,
                         std::vector<graph::leaf_ptr> rates,
                         graph::leaf_ptr time, graph::leaf_ptr dt, graph::leaf_ptr residual,
                         const std::string name, const size_t size,
                         jit::context<T, SAFE_MATH> &context) :
        work_item<T, SAFE_MATH> (typename work_item<T, SAFE_MATH>::no_emit(), inputs, {residual}, name, size) {
            context.add_runge_kutta(name, order, inputs, evolved, rates, time, dt, residual, size);
        }
    };

    template<typename T=double, bool SAFE_MATH=false>
    class manager {
    private:
        jit::context<T, SAFE_MATH> context;
        std::vector<std::unique_ptr<work_item<T, SAFE_MATH>>> preitems;
        std::vector<std::unique_ptr<work_item<T, SAFE_MATH>>> items;
        std::vector<std::unique_ptr<work_item<T, SAFE_MATH>>> side_items;
        bool add_reduction;
    public:
        manager(const size_t index) : context(index), add_reduction(false) {}

        void add_preitem(graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                         graph::map_nodes<T, SAFE_MATH> maps, graph::shared_random_state<T, SAFE_MATH> state,
                         const std::string name, const size_t size) {
            preitems.push_back(std::make_unique<work_item<T, SAFE_MATH>> (in, out, maps, state, name, size, context));
        }
        void add_item(graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                      graph::map_nodes<T, SAFE_MATH> maps, graph::shared_random_state<T, SAFE_MATH> state,
                      const std::string name, const size_t size) {
            items.push_back(std::make_unique<work_item<T, SAFE_MATH>> (in, out, maps, state, name, size, context));
        }
        void add_loop_item(graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                           graph::map_nodes<T, SAFE_MATH> maps, graph::shared_random_state<T, SAFE_MATH> state,
                           const std::string name, const size_t size, const size_t iterations) {
            items.push_back(std::make_unique<loop_item<T, SAFE_MATH>> (in, out, maps, state, name, size, context, iterations));
        }
        void add_converge_item(graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                               graph::map_nodes<T, SAFE_MATH> maps, graph::shared_random_state<T, SAFE_MATH> state,
                               const std::string name, const size_t size,
                               const T tol=1.0E-30, const size_t max_iter=1000) {
            add_reduction = true;
            items.push_back(std::make_unique<converge_item<T, SAFE_MATH>> (in, out, maps, state, name, size,
                                                                           context, tol, max_iter));
        }
        void add_newton_item(graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                             graph::map_nodes<T, SAFE_MATH> maps, const std::string name, const size_t size,
                             const T tol=1.0E-30, const size_t max_iter=1000) {
            items.push_back(std::make_unique<newton_item<T, SAFE_MATH>> (in, out, maps, name, size, context, tol, max_iter));
        }
        void add_runge_kutta_item(const int order, graph::input_nodes<T, SAFE_MATH> in,
                                  std::vector<graph::leaf_ptr> evolved, std::vector<graph::leaf_ptr> rates,
                                  graph::leaf_ptr time, graph::leaf_ptr dt, graph::leaf_ptr residual,
                                  const std::string name, const size_t size) {
            items.push_back(std::make_unique<runge_kutta_item<T, SAFE_MATH>> (order, in, evolved, rates, time, dt,
                                                                              residual, name, size, context));
        }

///  Extension: an item that shares this manager's device buffers but is NOT part of run(); it is
///  launched on request with run_side(id).  Used for kernels that act between blocks of steps
///  (absorption, diagnostics) on state that stays in HBM.
        size_t add_side_item(graph::input_nodes<T, SAFE_MATH> in, graph::output_nodes<T, SAFE_MATH> out,
                             graph::map_nodes<T, SAFE_MATH> maps, graph::shared_random_state<T, SAFE_MATH> state,
                             const std::string name, const size_t size) {
            side_items.push_back(std::make_unique<work_item<T, SAFE_MATH>> (in, out, maps, state, name, size, context));
            return side_items.size() - 1;
        }
        void run_side(const size_t id) { side_items.at(id)->run(); }

        void compile() {
            context.compile(add_reduction);
            for (auto &item : preitems) item->create_kernel_call(context);
            for (auto &item : items) item->create_kernel_call(context);
            for (auto &item : side_items) item->create_kernel_call(context);
        }
        void pre_run() { for (auto &item : preitems) item->run(); }
        void run() { for (auto &item : items) item->run(); }
        void wait() { context.wait(); }
        void copy_to_device(graph::shared_leaf<T, SAFE_MATH> &node, T *source) { context.copy_to_device(node, source); }
        void copy_to_host(graph::shared_leaf<T, SAFE_MATH> &node, T *destination) { context.copy_to_host(node, destination); }
        void print(const size_t index, const graph::output_nodes<T, SAFE_MATH> &nodes) { context.print(index, nodes); }
        T check_value(const size_t index, const graph::shared_leaf<T, SAFE_MATH> &node) {
            return context.check_value(index, node);
        }
        jit::context<T, SAFE_MATH> &get_context() { return context; }
    };
}

#endif /* gfb_graph_workflow_hpp */
