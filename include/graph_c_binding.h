/*------------------------------------------------------------------------------
 *  graph_c_binding.h -- C binding of the graph front end on the B200 back end.
 *
 *  Same entry points, argument order and meaning as the reference's C binding
 *  (/root/reference/graph_c_binding/graph_c_binding.h:70-657), which is what the
 *  Fortran module and any FFI user bind against.  Differences:
 *    - only type DOUBLE with use_safe_math == false is accepted (the FP64 ray
 *      path of the north star); other types make graph_construct_context
 *      return NULL and print why;
 *    - graph_random_state / graph_random / graph_constant_c are not on the ray path and are absent;
 *    - graph_erfi takes a real argument (the reference's needs a complex context;
 *      ray state is real, special_functions.hpp:1504-1512 is the branch it takes).
 *  Nodes are opaque `void *`; identical expressions give identical pointers
 *  (the reference's c_binding_test.c:43-68 relies on that).
 *----------------------------------------------------------------------------*/
#ifndef GFB_GRAPH_C_BINDING_H
#define GFB_GRAPH_C_BINDING_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#define STRUCT_TAG
#else
#define STRUCT_TAG struct
#endif

typedef void *graph_node;

enum graph_type { FLOAT, DOUBLE, COMPLEX_FLOAT, COMPLEX_DOUBLE };

struct graph_c_context {
    enum graph_type type;
    bool safe_math;
};

/* graph_c_binding.h:93-101 */
STRUCT_TAG graph_c_context *graph_construct_context(const enum graph_type type, const bool use_safe_math);
void graph_destroy_context(STRUCT_TAG graph_c_context *c);

/* Leaves: graph_c_binding.h:110-171 */
graph_node graph_variable(STRUCT_TAG graph_c_context *c, const size_t size, const char *symbol);
graph_node graph_constant(STRUCT_TAG graph_c_context *c, const double value);
void graph_set_variable(STRUCT_TAG graph_c_context *c, graph_node var, const void *source);
graph_node graph_pseudo_variable(STRUCT_TAG graph_c_context *c, graph_node var);
graph_node graph_remove_pseudo(STRUCT_TAG graph_c_context *c, graph_node var);

/* Arithmetic: graph_c_binding.h:183-225 */
graph_node graph_add(STRUCT_TAG graph_c_context *c, graph_node left, graph_node right);
graph_node graph_sub(STRUCT_TAG graph_c_context *c, graph_node left, graph_node right);
graph_node graph_mul(STRUCT_TAG graph_c_context *c, graph_node left, graph_node right);
graph_node graph_div(STRUCT_TAG graph_c_context *c, graph_node left, graph_node right);
graph_node graph_fma(STRUCT_TAG graph_c_context *c, graph_node a, graph_node b, graph_node d);

/* Math and trigonometry: graph_c_binding.h:237-327 */
graph_node graph_sqrt(STRUCT_TAG graph_c_context *c, graph_node arg);
graph_node graph_exp(STRUCT_TAG graph_c_context *c, graph_node arg);
graph_node graph_log(STRUCT_TAG graph_c_context *c, graph_node arg);
graph_node graph_erfi(STRUCT_TAG graph_c_context *c, graph_node arg);      /* graph_c_binding.h:349 */
graph_node graph_pow(STRUCT_TAG graph_c_context *c, graph_node left, graph_node right);
graph_node graph_sin(STRUCT_TAG graph_c_context *c, graph_node arg);
graph_node graph_cos(STRUCT_TAG graph_c_context *c, graph_node arg);
graph_node graph_atan(STRUCT_TAG graph_c_context *c, graph_node left, graph_node right);

/* Piecewise constants: graph_c_binding.h:363-401 */
graph_node graph_piecewise_1D(STRUCT_TAG graph_c_context *c, graph_node arg, const double scale,
                              const double offset, const void *source, const size_t source_size);
graph_node graph_piecewise_2D(STRUCT_TAG graph_c_context *c, const size_t num_cols,
                              graph_node x_arg, const double x_scale, const double x_offset,
                              graph_node y_arg, const double y_scale, const double y_offset,
                              const void *source, const size_t source_size);
/* Gathers from a variable: graph_c_binding.h:458-486 */
graph_node graph_index_1D(STRUCT_TAG graph_c_context *c, graph_node variable, graph_node arg,
                          const double scale, const double offset);
graph_node graph_index_2D(STRUCT_TAG graph_c_context *c, graph_node variable, const size_t num_cols,
                          graph_node x_arg, const double x_scale, const double x_offset,
                          graph_node y_arg, const double y_scale, const double y_offset);

/* Derivative: graph_c_binding.h:649-653 */
graph_node graph_df(STRUCT_TAG graph_c_context *c, graph_node fnode, graph_node xnode);

/* Workflow: graph_c_binding.h:456-640 */
size_t graph_get_max_concurrency(STRUCT_TAG graph_c_context *c);
void graph_set_device_number(STRUCT_TAG graph_c_context *c, const size_t num);
void graph_add_pre_item(STRUCT_TAG graph_c_context *c, graph_node *inputs, size_t num_inputs,
                        graph_node *outputs, size_t num_outputs,
                        graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                        graph_node random_state, const char *name, const size_t size);
void graph_add_item(STRUCT_TAG graph_c_context *c, graph_node *inputs, size_t num_inputs,
                    graph_node *outputs, size_t num_outputs,
                    graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                    graph_node random_state, const char *name, const size_t size);
void graph_add_converge_item(STRUCT_TAG graph_c_context *c, graph_node *inputs, size_t num_inputs,
                             graph_node *outputs, size_t num_outputs,
                             graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                             graph_node random_state, const char *name, const size_t size,
                             const double tol, const size_t max_iter);
void graph_compile(STRUCT_TAG graph_c_context *c);
void graph_pre_run(STRUCT_TAG graph_c_context *c);
void graph_run(STRUCT_TAG graph_c_context *c);
void graph_wait(STRUCT_TAG graph_c_context *c);
void graph_copy_to_device(STRUCT_TAG graph_c_context *c, graph_node node, void *source);
void graph_copy_to_host(STRUCT_TAG graph_c_context *c, graph_node node, void *destination);
void graph_print(STRUCT_TAG graph_c_context *c, const size_t index, graph_node *nodes, const size_t num_nodes);

/* Extensions (not in the reference binding): host evaluation of a node
 * (leaf_node::evaluate, node.hpp:378) and the emitted source text. */
size_t graph_evaluate(STRUCT_TAG graph_c_context *c, graph_node node, double *destination, const size_t capacity);
const char *graph_get_source(STRUCT_TAG graph_c_context *c);
/* Extension: arithmetic mode of the kernels compiled for this context.  on (default): a/b is a times
 * a refined hardware reciprocal (<= 1 ulp, but x/0 and x/inf give NaN where IEEE gives inf and 0), sqrt
 * comes from rsqrt, and table indices use (x - offset)*(1/scale) -- what -ffast-math makes of the
 * reference's own kernels (cpu_context.hpp:155-157; pinned cell by cell in tests/golden/ref_cells_efit.npz).
 * off: IEEE division and sqrt, indices by (x - offset)/scale exactly as piecewise.hpp:26-65 is written.
 * Call before graph_compile. */
void graph_set_fast_division(STRUCT_TAG graph_c_context *c, const bool on);

#ifdef __cplusplus
}
#endif
#endif /* GFB_GRAPH_C_BINDING_H */
