//------------------------------------------------------------------------------
//  newton.hpp -- Newton's method as a work item.
//
//  /root/reference/graph_framework/newton.hpp:34-51: for each unknown x the
//  setter is  x <- x - step*f/(df/dx)  and the item's output is f^2.
//  `mode` selects the reference's host-driven ensemble-maximum loop
//  (converge_item) or the device-resident per-ray iteration (newton_item).
//------------------------------------------------------------------------------
#ifndef gfb_graph_newton_hpp
#define gfb_graph_newton_hpp

#include "workflow.hpp"

namespace solver {
    enum class newton_mode { per_ray, ensemble };

    template<typename T=double, bool SAFE_MATH=false>
    void newton(workflow::manager<T, SAFE_MATH> &work,
                graph::output_nodes<T, SAFE_MATH> vars,
                graph::input_nodes<T, SAFE_MATH> inputs,
                graph::shared_leaf<T, SAFE_MATH> func,
                graph::shared_random_state<T, SAFE_MATH> state,
                const T tolerance = 1.0E-30,
                const size_t max_iterations = 1000,
                const T step = 1.0,
                const newton_mode mode = newton_mode::ensemble) {
        graph::map_nodes<T, SAFE_MATH> setters;
        for (auto x : vars) {
            setters.push_back({x - step*func/func->df(x), graph::variable_cast(x)});
        }
        if (mode == newton_mode::ensemble) {
            work.add_converge_item(inputs, {func*func}, setters, state, "loss_kernel",
                                   inputs.back()->size(), tolerance, max_iterations);
        } else {
            work.add_newton_item(inputs, {func*func}, setters, "loss_kernel",
                                 inputs.back()->size(), tolerance, max_iterations);
        }
    }
}

#endif /* gfb_graph_newton_hpp */
