/* The reference's OWN binding test (graph_tests/c_binding_test.c) and its OWN header
 * (graph_c_binding/graph_c_binding.h), both read where they lie under the reference tree and
 * unmodified, linked against libgfb200.so: run_tests(DOUBLE, false) -- the one type this back end
 * implements -- asserts node identity after reduction (sub(1,1) == 0, remove_pseudo(px) == x, ...),
 * y/dydx/dydm/dydb/dydy, the converge item (Newton on z^3 - z^2), piecewise_1D/2D and index_1D/2D.
 *
 * Three entry points the file references are off the FP64 ray path (random.hpp
 * and complex types are out of scope) and are supplied HERE, not by the library:
 *   graph_constant_c        only reached for COMPLEX types: aborts;
 *   graph_random_state / graph_random / the pre-item that fills `rand`: a one-element variable holding
 *                           2357136044.0, the value the reference asserts for seed 0, carried through a
 *                           do-nothing pre-item so that the test's copy_to_host(rand) has a device buffer.
 * Everything else the test calls is libgfb200's.  TEST INFRASTRUCTURE (integration/Makefile).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdbool.h>
#include <stddef.h>

#include GFB_REFERENCE_HEADER

graph_node graph_constant_c(STRUCT_TAG graph_c_context *c, const double real_value, const double imag_value) {
    (void)c; (void)real_value; (void)imag_value;
    fprintf(stderr, "graph_constant_c: complex types are not part of the B200 back end\n");
    abort();
}
graph_node graph_random_state(STRUCT_TAG graph_c_context *c, const uint32_t seed) {
    (void)c; (void)seed;
    return NULL;
}
static graph_node rand_stand_in = NULL;
graph_node graph_random(STRUCT_TAG graph_c_context *c, graph_node state) {
    (void)state;
    const double value = 2357136044.0;
    rand_stand_in = graph_variable(c, 1, "rand");
    graph_set_variable(c, rand_stand_in, &value);
    return rand_stand_in;
}
static void pre_item_stand_in(STRUCT_TAG graph_c_context *c, graph_node *inputs, size_t num_inputs,
                              graph_node *outputs, size_t num_outputs,
                              graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                              graph_node random_state, const char *name, const size_t num_particles) {
    (void)inputs; (void)num_inputs; (void)outputs; (void)num_outputs; (void)map_inputs; (void)map_outputs; (void)num_maps;
    graph_node in[1] = {rand_stand_in};
    graph_node out[1] = {graph_mul(c, rand_stand_in, graph_constant(c, 2.0))};   /* a pre-item must compute something */
    graph_add_pre_item(c, in, 1, out, 1, NULL, NULL, 0, random_state, name, num_particles);
}

#define graph_add_pre_item pre_item_stand_in
#define main reference_main
#include GFB_REFERENCE_TEST
#undef main
#undef graph_add_pre_item

int main(void) {
    run_tests(DOUBLE, false);
    printf("reference c_binding_test.c run_tests(DOUBLE, false): all assertions passed\n");
    return 0;
}
