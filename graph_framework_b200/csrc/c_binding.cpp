//------------------------------------------------------------------------------
//  c_binding.cpp -- extern "C" facades over the host front end.
//
//    include/graph_c_binding.h : mirror of the reference's C binding
//        (/root/reference/graph_c_binding/graph_c_binding.cpp: every call forwards
//        1:1 to the graph / workflow layer; nodes are kept alive by the context).
//    include/gfb_rays.h        : the ray tracing call sequence of
//        /root/reference/graph_benchmark/xrays_bench.cpp:53-101 and the Boris push
//        of /root/reference/graph_korc/xkorc.cpp:40-121 as flat C.
//------------------------------------------------------------------------------
#include "../../include/graph_c_binding.h"
#include "../../include/gfb_rays.h"

#include "graph/graph_framework.hpp"
#include "graph/boris.hpp"

using graph::leaf_ptr;

//  The node caches of the front end are per thread and keep every interned node -- and through
//  the variable nodes the n-double host copies of the ray arrays -- alive.  Handles are counted per
//  thread; when the last one created on a thread is destroyed there, the caches are dropped.
namespace {
thread_local long live_handles = 0;
void handle_created() { live_handles++; }
void handle_destroyed() {
    if (live_handles > 0 && --live_handles == 0) graph::clear_caches();
}
}

//******************************************************************************
//  graph_c_binding
//******************************************************************************
namespace {
struct c_context : public graph_c_context {
    std::map<void *, leaf_ptr> nodes;                          // keeps nodes alive (graph_c_binding.cpp:19-29)
    std::unique_ptr<workflow::manager<>> work;
    size_t device = 0;
    std::string source;

    graph_node keep(leaf_ptr n) {
        nodes[n.get()] = n;
        return n.get();
    }
    leaf_ptr get(graph_node n) {
        auto it = nodes.find(n);
        if (it == nodes.end()) {
            std::cerr << "graph_c_binding: unknown node handle." << std::endl;
            std::abort();
        }
        return it->second;
    }
    workflow::manager<> &manager() {
        if (!work) work = std::make_unique<workflow::manager<>> (device);
        return *work;
    }
    std::vector<leaf_ptr> list(graph_node *n, const size_t count) {
        std::vector<leaf_ptr> v;
        for (size_t i = 0; i < count; i++) v.push_back(get(n[i]));
        return v;
    }
    graph::map_nodes<> maps(graph_node *in, graph_node *out, const size_t count) {
        graph::map_nodes<> m;
        for (size_t i = 0; i < count; i++) m.push_back({get(out[i]), get(in[i])});
        return m;
    }
};
c_context *cast(graph_c_context *c) { return static_cast<c_context *> (c); }
}  // namespace

extern "C" {
graph_c_context *graph_construct_context(const enum graph_type type, const bool use_safe_math) {
    if (type != DOUBLE || use_safe_math) {
        std::cerr << "graph_construct_context: the B200 back end implements DOUBLE with safe_math = false only."
                  << std::endl;
        return nullptr;
    }
    auto c = new c_context;
    c->type = type;
    c->safe_math = use_safe_math;
    handle_created();
    return c;
}
void graph_destroy_context(graph_c_context *c) {
    if (!c) return;
    delete cast(c);
    handle_destroyed();
}

graph_node graph_variable(graph_c_context *c, const size_t size, const char *symbol) {
    return cast(c)->keep(graph::variable(size, symbol));
}
graph_node graph_constant(graph_c_context *c, const double value) { return cast(c)->keep(graph::constant(value)); }
void graph_set_variable(graph_c_context *c, graph_node var, const void *source) {
    auto v = cast(c)->get(var);
    const double *s = static_cast<const double *> (source);
    v->set(std::vector<double> (s, s + v->size()));
}
graph_node graph_pseudo_variable(graph_c_context *c, graph_node var) {
    return cast(c)->keep(graph::pseudo_variable(cast(c)->get(var)));
}
graph_node graph_remove_pseudo(graph_c_context *c, graph_node var) {
    return cast(c)->keep(cast(c)->get(var)->remove_pseudo());
}
#define GFB_BINARY(NAME, EXPR)                                                              \
    graph_node NAME(graph_c_context *c, graph_node left, graph_node right) {                \
        auto l = cast(c)->get(left);                                                        \
        auto r = cast(c)->get(right);                                                       \
        return cast(c)->keep(EXPR);                                                         \
    }
#define GFB_UNARY(NAME, EXPR)                                                               \
    graph_node NAME(graph_c_context *c, graph_node arg) {                                   \
        auto a = cast(c)->get(arg);                                                         \
        return cast(c)->keep(EXPR);                                                         \
    }
GFB_BINARY(graph_add, l + r)
GFB_BINARY(graph_sub, l - r)
GFB_BINARY(graph_mul, l*r)
GFB_BINARY(graph_div, l/r)
GFB_BINARY(graph_pow, graph::pow(l, r))
GFB_BINARY(graph_atan, graph::atan(l, r))
GFB_UNARY(graph_sqrt, graph::sqrt(a))
GFB_UNARY(graph_exp, graph::exp(a))
GFB_UNARY(graph_log, graph::log(a))
GFB_UNARY(graph_erfi, graph::erfi(a))
GFB_UNARY(graph_sin, graph::sin(a))
GFB_UNARY(graph_cos, graph::cos(a))
graph_node graph_fma(graph_c_context *c, graph_node a, graph_node b, graph_node d) {
    return cast(c)->keep(graph::fma(cast(c)->get(a), cast(c)->get(b), cast(c)->get(d)));
}
graph_node graph_piecewise_1D(graph_c_context *c, graph_node arg, const double scale, const double offset,
                              const void *source, const size_t source_size) {
    const double *s = static_cast<const double *> (source);
    return cast(c)->keep(graph::piecewise_1D(std::vector<double> (s, s + source_size), cast(c)->get(arg), scale, offset));
}
graph_node graph_piecewise_2D(graph_c_context *c, const size_t num_cols,
                              graph_node x_arg, const double x_scale, const double x_offset,
                              graph_node y_arg, const double y_scale, const double y_offset,
                              const void *source, const size_t source_size) {
    const double *s = static_cast<const double *> (source);
    return cast(c)->keep(graph::piecewise_2D(std::vector<double> (s, s + source_size), num_cols,
                                             cast(c)->get(x_arg), x_scale, x_offset,
                                             cast(c)->get(y_arg), y_scale, y_offset));
}
graph_node graph_index_1D(graph_c_context *c, graph_node variable, graph_node arg, const double scale, const double offset) {
    return cast(c)->keep(graph::index_1D(cast(c)->get(variable), cast(c)->get(arg), scale, offset));
}
graph_node graph_index_2D(graph_c_context *c, graph_node variable, const size_t num_cols,
                          graph_node x_arg, const double x_scale, const double x_offset,
                          graph_node y_arg, const double y_scale, const double y_offset) {
    return cast(c)->keep(graph::index_2D(cast(c)->get(variable), num_cols, cast(c)->get(x_arg), x_scale, x_offset,
                                         cast(c)->get(y_arg), y_scale, y_offset));
}
graph_node graph_df(graph_c_context *c, graph_node fnode, graph_node xnode) {
    return cast(c)->keep(cast(c)->get(fnode)->df(cast(c)->get(xnode)));
}

size_t graph_get_max_concurrency(graph_c_context *) { return static_cast<size_t> (gfb_device_count()); }
void graph_set_device_number(graph_c_context *c, const size_t num) {
//  Replaces the manager, like graph_c_binding.cpp:1939-1979.
    cast(c)->device = num;
    cast(c)->work.reset();
}
void graph_add_pre_item(graph_c_context *c, graph_node *inputs, size_t num_inputs,
                        graph_node *outputs, size_t num_outputs,
                        graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                        graph_node, const char *name, const size_t size) {
    auto x = cast(c);
    x->manager().add_preitem(x->list(inputs, num_inputs), x->list(outputs, num_outputs),
                             x->maps(map_inputs, map_outputs, num_maps), graph::shared_random_state<> (), name, size);
}
void graph_add_item(graph_c_context *c, graph_node *inputs, size_t num_inputs,
                    graph_node *outputs, size_t num_outputs,
                    graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                    graph_node, const char *name, const size_t size) {
    auto x = cast(c);
    x->manager().add_item(x->list(inputs, num_inputs), x->list(outputs, num_outputs),
                          x->maps(map_inputs, map_outputs, num_maps), graph::shared_random_state<> (), name, size);
}
void graph_add_converge_item(graph_c_context *c, graph_node *inputs, size_t num_inputs,
                             graph_node *outputs, size_t num_outputs,
                             graph_node *map_inputs, graph_node *map_outputs, size_t num_maps,
                             graph_node, const char *name, const size_t size,
                             const double tol, const size_t max_iter) {
    auto x = cast(c);
    x->manager().add_converge_item(x->list(inputs, num_inputs), x->list(outputs, num_outputs),
                                   x->maps(map_inputs, map_outputs, num_maps), graph::shared_random_state<> (),
                                   name, size, tol, max_iter);
}
void graph_compile(graph_c_context *c) { cast(c)->manager().compile(); }
void graph_pre_run(graph_c_context *c) { cast(c)->manager().pre_run(); }
void graph_run(graph_c_context *c) { cast(c)->manager().run(); }
void graph_wait(graph_c_context *c) { cast(c)->manager().wait(); }
void graph_copy_to_device(graph_c_context *c, graph_node node, void *source) {
    auto n = cast(c)->get(node);
    cast(c)->manager().copy_to_device(n, static_cast<double *> (source));
}
void graph_copy_to_host(graph_c_context *c, graph_node node, void *destination) {
    auto n = cast(c)->get(node);
    cast(c)->manager().copy_to_host(n, static_cast<double *> (destination));
}
void graph_print(graph_c_context *c, const size_t index, graph_node *nodes, const size_t num_nodes) {
    cast(c)->manager().print(index, cast(c)->list(nodes, num_nodes));
}
size_t graph_evaluate(graph_c_context *c, graph_node node, double *destination, const size_t capacity) {
    auto b = cast(c)->get(node)->evaluate();
    for (size_t i = 0; i < b.size() && i < capacity; i++) destination[i] = b[i];
    return b.size();
}
void graph_set_fast_division(graph_c_context *c, const bool on) {
    cast(c)->manager().get_context().options.fast_division = on;
}
const char *graph_get_source(graph_c_context *c) {
    cast(c)->source = cast(c)->manager().get_context().get_source();
    return cast(c)->source.c_str();
}
}  // extern "C"

//******************************************************************************
//  gfb_rays
//******************************************************************************
namespace {
equilibrium::shared<> make_equilibrium(const std::string &name, const std::string &table_file) {
    if (name == "efit") return equilibrium::make_efit<> (table_file);
    if (name == "vmec") return equilibrium::make_vmec<> (table_file);
    if (name == "slab") return equilibrium::make_slab<> ();
    if (name == "slab_density") return equilibrium::make_slab_density<> ();
    if (name == "slab_field") return equilibrium::make_slab_field<> ();
    if (name == "no_magnetic_field") return equilibrium::make_no_magnetic_field<> ();
    if (name == "gaussian_density") return equilibrium::make_gaussian_density<> ();
    return equilibrium::shared<> ();
}

struct run_options {
    jit::emit_options emit;
    int unroll_stages = -1;         // -1: decided by body size (jit::context::compile)
    unsigned fused_steps = 0;
    bool absorption = false;
    int bin_rays = -1;              // -1: on when the equilibrium has tables; 0 off; > 0 re-sort period in steps
    int reference_defects = -1;     // -1: dispersion::reference_defects() as it stands; 0 true derivatives; 1 the reference's
};
run_options parse_options(const char *options) {
    run_options o;
    if (!options) return o;
    std::istringstream is(options);
    std::string kv;
    while (is >> kv) {
        const size_t eq = kv.find('=');
        if (eq == std::string::npos) continue;
        const std::string k = kv.substr(0, eq);
        const long v = std::atol(kv.substr(eq + 1).c_str());
        if (k == "block") o.emit.block_size = static_cast<unsigned> (v);
        else if (k == "minblocks") o.emit.min_blocks = static_cast<unsigned> (v);
        else if (k == "stage_tables") o.emit.stage_tables = v != 0;
        else if (k == "share_rcp") o.emit.share_reciprocals = v != 0;
        else if (k == "fast_div") o.emit.fast_division = v != 0;
        else if (k == "mode_unroll") o.emit.mode_loop_unroll = static_cast<unsigned> (v);
        else if (k == "mode_recurrence") o.emit.mode_recurrence = v != 0;
        else if (k == "unroll_stages") o.unroll_stages = v != 0 ? 1 : 0;
        else if (k == "fused_steps") o.fused_steps = static_cast<unsigned> (v);
        else if (k == "absorption") o.absorption = v != 0;
        else if (k == "bin_rays") o.bin_rays = static_cast<int> (v);
        else if (k == "reference_defects") o.reference_defects = v != 0 ? 1 : 0;
    }
    return o;
}

struct tracer_base {
    virtual ~tracer_base() {}
    virtual std::vector<leaf_ptr> state() = 0;
    virtual void init(leaf_ptr var, double tol, size_t max_iter, int mode) = 0;
    virtual void compile() = 0;
    virtual void step(size_t n) = 0;
    virtual void wait() = 0;
    virtual void sync_host() = 0;
    virtual void sync_device() = 0;
    virtual leaf_ptr residual() = 0;
    virtual jit::context<> &context() = 0;
    virtual std::vector<leaf_ptr> rhs() = 0;
    virtual workflow::manager<> &work() = 0;
    virtual workflow::ray_order<> &order() = 0;
    virtual void set_ray_order(bool on, size_t period) = 0;

//  Absorption attached to the solver's device context (absorption.hpp): state that the
//  Runge-Kutta kernel leaves in HBM is read in place.
    leaf_ptr step_variable;         // adaptive_rk4: the per-ray dt
    leaf_ptr kamp_re, kamp_im, x_last, y_last, z_last, power, k_sum;
    std::unique_ptr<absorption::weak_damping<>> damping;
    std::unique_ptr<absorption::power_item<>> deposition;
    size_t reset_item = 0;
    void attach_absorption(equilibrium::shared<> &eq, const size_t n) {
        auto s = state();
        kamp_re = graph::variable(n, 0.0, "kamp_re");
        kamp_im = graph::variable(n, 0.0, "kamp_im");
        x_last = graph::variable(n, 0.0, "x_last");
        y_last = graph::variable(n, 0.0, "y_last");
        z_last = graph::variable(n, 0.0, "z_last");
        power = graph::variable(n, 1.0, "power");
        k_sum = graph::variable(n, 0.0, "k_sum");
        damping = std::make_unique<absorption::weak_damping<>> (work(), kamp_re, kamp_im, s[GFB_W], s[GFB_KX], s[GFB_KY],
                                                                 s[GFB_KZ], s[GFB_X], s[GFB_Y], s[GFB_Z], s[GFB_T], eq);
        deposition = std::make_unique<absorption::power_item<>> (work(), s[GFB_X], s[GFB_Y], s[GFB_Z], x_last, y_last, z_last,
                                                                  kamp_im, power, k_sum, eq);
//  Start of a power calculation (xrays.cpp:745-755): X_last = X, power = 1, k_sum = 0.
        reset_item = work().add_side_item({s[GFB_X], s[GFB_Y], s[GFB_Z], x_last, y_last, z_last, power, k_sum}, {},
                                          {{s[GFB_X], x_last}, {s[GFB_Y], y_last}, {s[GFB_Z], z_last},
                                           {graph::one(), power}, {graph::zero(), k_sum}},
                                          graph::shared_random_state<> (), "power_reset", n);
    }
};

template<class SOLVER>
struct tracer final : public tracer_base {
    SOLVER solve;
    tracer(std::vector<leaf_ptr> &s, leaf_ptr dt, equilibrium::shared<> &eq, const size_t n, const int device,
           const run_options &o) :
    solve(s[GFB_W], s[GFB_KX], s[GFB_KY], s[GFB_KZ], s[GFB_X], s[GFB_Y], s[GFB_Z], s[GFB_T], dt, eq, "", n, device) {
        solve.get_work().get_context().options = o.emit;
        if (o.unroll_stages >= 0) {
            solve.get_work().get_context().set_nvrtc_options(o.unroll_stages ? "-DGFB_UNROLL_STAGES=1" : "-DGFB_UNROLL_STAGES=0");
        }
        if (o.fused_steps) gfb_set_max_fused_steps(solve.get_work().get_context().device(), o.fused_steps);
    }
    std::vector<leaf_ptr> state() override { return solve.state(); }
    void init(leaf_ptr var, double tol, size_t max_iter, int mode) override {
        solve.set_newton_mode(mode ? solver::newton_mode::ensemble : solver::newton_mode::per_ray);
        if (var.get()) solve.init(var, tol, max_iter);
        else solve.init();
    }
    void compile() override { solve.compile(); }
    void step(size_t n) override { solve.step(n); }
    void wait() override { solve.wait(); }
    void sync_host() override { solve.sync_host(); }
    void sync_device() override { solve.sync_device(); }
    leaf_ptr residual() override { return solve.get_residual(); }
    jit::context<> &context() override { return solve.get_work().get_context(); }
    workflow::manager<> &work() override { return solve.get_work(); }
    workflow::ray_order<> &order() override { return solve.get_order(); }
    void set_ray_order(bool on, size_t period) override { solve.set_ray_order(on, period); }
    std::vector<leaf_ptr> rhs() override {
        auto &D = solve.get_dispersion();
        return {D.get_dxdt(), D.get_dydt(), D.get_dzdt(), D.get_dkxdt(), D.get_dkydt(), D.get_dkzdt(), D.get_d()};
    }
};

template<class DF>
tracer_base *make_tracer(const std::string &solver_name, std::vector<leaf_ptr> &s, leaf_ptr dt,
                         equilibrium::shared<> &eq, const size_t n, const int device, const run_options &o) {
    if (solver_name == "rk4") return new tracer<solver::rk4<DF, true>> (s, dt, eq, n, device, o);
    if (solver_name == "rk2") return new tracer<solver::rk2<DF, true>> (s, dt, eq, n, device, o);
    if (solver_name == "rk4_graph") return new tracer<solver::rk4<DF, false>> (s, dt, eq, n, device, o);
    if (solver_name == "rk2_graph") return new tracer<solver::rk2<DF, false>> (s, dt, eq, n, device, o);
    if (solver_name == "adaptive_rk4") {
//  dt becomes a per-ray variable that the solver's own Newton item rewrites before every step (solver.hpp:881-1006).
        auto dt_var = graph::variable(n, dt->value, "dt");
        auto t = new tracer<solver::adaptive_rk4<DF>> (s, dt_var, eq, n, device, o);
        t->step_variable = dt_var;
        return t;
    }
    if (solver_name == "split_simplextic") {
//  Only separable Hamiltonians (the constructor aborts otherwise, like the reference's assert).
        if constexpr (std::is_same<DF, dispersion::bohm_gross<>>::value || std::is_same<DF, dispersion::light_wave<>>::value ||
                      std::is_same<DF, dispersion::simple<>>::value || std::is_same<DF, dispersion::acoustic_wave<>>::value) {
            return new tracer<solver::split_simplextic<DF>> (s, dt, eq, n, device, o);
        }
    }
    return nullptr;
}
}  // namespace

struct gfb_rays {
    size_t n;
    int device;
    std::vector<leaf_ptr> vars;       // t, w, x, y, z, kx, ky, kz
    std::unique_ptr<tracer_base> impl;
    bool compiled = false;
    bool absorption_started = false;
    char profile_tag = 0;             // its address keys the deposition profile buffer
    std::string source;
};

namespace {
int rays_fail(const std::string &s) {
    gfb_set_last_error(s.c_str());
    std::cerr << "gfb_rays: " << s << std::endl;
    return 1;
}
}

extern "C" {
gfb_rays *gfb_rays_create(const char *dispersion_name, const char *equilibrium_name, const char *table_file,
                          const char *solver_name, size_t num_rays, double dt, int device, const char *options) {
    if (gfb_device_count() <= 0) {
        rays_fail("no CUDA device: the B200 back end has no CPU fallback");
        return nullptr;
    }
    auto eq = make_equilibrium(equilibrium_name, table_file ? table_file : "");
    if (!eq) {
        rays_fail(std::string("unknown equilibrium ") + equilibrium_name);
        return nullptr;
    }
    auto r = std::make_unique<gfb_rays> ();
    r->n = num_rays;
    r->device = device;
    static const char *symbols[GFB_NUM_STATE] = {"t", "\\omega", "x", "y", "z", "k_{x}", "k_{y}", "k_{z}"};
    for (int i = 0; i < GFB_NUM_STATE; i++) r->vars.push_back(graph::variable(num_rays, symbols[i]));
    auto dtc = graph::constant(dt);
    const run_options o = parse_options(options);
    const std::string d = dispersion_name, s = solver_name;
    tracer_base *t = nullptr;
    const bool defects_before = dispersion::reference_defects();
    if (o.reference_defects >= 0) dispersion::reference_defects() = o.reference_defects != 0;
    if (d == "cold_plasma") t = make_tracer<dispersion::cold_plasma<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "ordinary_wave") t = make_tracer<dispersion::ordinary_wave<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "extra_ordinary_wave") t = make_tracer<dispersion::extra_ordinary_wave<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "bohm_gross") t = make_tracer<dispersion::bohm_gross<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "simple") t = make_tracer<dispersion::simple<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "light_wave") t = make_tracer<dispersion::light_wave<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "acoustic_wave") t = make_tracer<dispersion::acoustic_wave<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "gaussian_well") t = make_tracer<dispersion::gaussian_well<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "ion_cyclotron") t = make_tracer<dispersion::ion_cyclotron<>> (s, r->vars, dtc, eq, num_rays, device, o);
    else if (d == "stiff") t = make_tracer<dispersion::stiff<>> (s, r->vars, dtc, eq, num_rays, device, o);
    dispersion::reference_defects() = defects_before;
    if (!t) {
        rays_fail(std::string("unknown dispersion/solver ") + d + "/" + s);
        return nullptr;
    }
    r->impl.reset(t);
    if (o.absorption) t->attach_absorption(eq, num_rays);
//  Tabulated equilibria: the solver keeps rays sorted by table cell while stepping (binning.hpp);
//  options: bin_rays=0 off, bin_rays=<steps> how often the order is checked.
    if (o.bin_rays >= 0) t->set_ray_order(o.bin_rays != 0, o.bin_rays > 0 ? static_cast<size_t> (o.bin_rays) : 0);
    handle_created();
    return r.release();
}
void gfb_rays_destroy(gfb_rays *r) {
    if (!r) return;
    delete r;
    handle_destroyed();
}

int gfb_rays_set_state(gfb_rays *r, const double *const state[GFB_NUM_STATE]) {
    for (int i = 0; i < GFB_NUM_STATE; i++) {
        if (state[i]) r->vars[i]->set(std::vector<double> (state[i], state[i] + r->n));
    }
    if (r->compiled) r->impl->sync_device();        // restores the caller's ray order first
    return 0;
}
int gfb_rays_init(gfb_rays *r, const char *var, double tolerance, size_t max_iterations, int mode) {
    static const char *names[GFB_NUM_STATE] = {"t", "w", "x", "y", "z", "kx", "ky", "kz"};
    leaf_ptr v;
    if (var && var[0]) {
        for (int i = 0; i < GFB_NUM_STATE; i++) if (std::string(var) == names[i]) v = r->vars[i];
        if (!v) return rays_fail(std::string("unknown state variable ") + var);
    }
    r->impl->init(v, tolerance, max_iterations, mode);
    return 0;
}
int gfb_rays_compile(gfb_rays *r) {
    r->impl->compile();
    r->compiled = true;
    tracer_base &t = *r->impl;
//  The running absorption state travels with its ray when the order changes.
    if (t.damping) t.order().add_arrays({t.kamp_re, t.kamp_im, t.x_last, t.y_last, t.z_last, t.power, t.k_sum});
    return 0;
}
int gfb_rays_step(gfb_rays *r, size_t num_steps) {
    if (!r->compiled) return rays_fail("step before compile");
    r->impl->step(num_steps);
    return gfb_flush(r->impl->context().device());
}
int gfb_rays_wait(gfb_rays *r) {
    r->impl->wait();
    return 0;
}
int gfb_rays_set_binning(gfb_rays *r, int which_state, double lo, double hi, unsigned cells, size_t rebin_every) {
    if (!r->compiled) return rays_fail("set_binning before compile");
    if (which_state >= GFB_NUM_STATE) return rays_fail("bad state index");
    tracer_base &t = *r->impl;
    if (which_state < 0) {
        t.order().disable();
        return 0;
    }
    if (cells == 0 || !(hi > lo)) return rays_fail("bad binning grid");
    equilibrium::cell_grid grid;
    grid.dims = 1;
    grid.lo[0] = lo;
    grid.hi[0] = hi;
    grid.cells[0] = cells;
    std::vector<leaf_ptr> arrays = r->vars;
    if (t.damping) for (auto v : {t.kamp_re, t.kamp_im, t.x_last, t.y_last, t.z_last, t.power, t.k_sum}) arrays.push_back(v);
    t.order().configure(t.work(), grid, {r->vars[which_state]}, arrays, {t.residual()}, r->n, rebin_every);
    return 0;
}
int gfb_rays_get_state(gfb_rays *r, double *const state[GFB_NUM_STATE], double *residual) {
//  Straight from the device into the caller's (ideally pinned) memory, in the caller's ray order; rays
//  that are binned at the moment stay binned on the device.
    gfb_ctx *ctx = r->compiled ? r->impl->context().device() : nullptr;
    if (state) {
        for (int i = 0; i < GFB_NUM_STATE; i++) {
            if (!state[i]) continue;
            if (r->compiled) {
                if (gfb_copy_rays_d2h(ctx, reinterpret_cast<uint64_t> (r->vars[i].get()), state[i], r->n)) return 1;
            } else {
                std::memcpy(state[i], r->vars[i]->data(), sizeof(double)*r->n);
            }
        }
    }
    if (residual && r->compiled) {
        if (gfb_copy_rays_d2h(ctx, reinterpret_cast<uint64_t> (r->impl->residual().get()), residual, r->n)) return 1;
    }
    return 0;
}
int gfb_rays_put_state(gfb_rays *r, const double *const state[GFB_NUM_STATE]) {
    if (!r->compiled) return gfb_rays_set_state(r, state);
    r->impl->order().restore();
    for (int i = 0; i < GFB_NUM_STATE; i++) {
        if (state[i]) r->impl->context().copy_to_device(r->vars[i], const_cast<double *> (state[i]));
    }
    return 0;
}
int gfb_rays_step_host(gfb_rays *r, size_t num_steps, const double *const state_in[GFB_NUM_STATE],
                       double *const state_out[GFB_NUM_STATE], double *residual_out, int chunks) {
    if (!r->compiled) return rays_fail("step_host before compile");
    r->impl->order().restore();
//  Pointer slots of solver_kernel: the 8 inputs in solver order (= GFB_T..GFB_KZ), then the residual.
    const void *src[GFB_NUM_STATE + 1];
    void *dst[GFB_NUM_STATE + 1];
    for (int i = 0; i < GFB_NUM_STATE; i++) {
        src[i] = state_in ? state_in[i] : nullptr;
        dst[i] = state_out ? state_out[i] : nullptr;
    }
    src[GFB_NUM_STATE] = nullptr;
    dst[GFB_NUM_STATE] = residual_out;
    gfb_kernel *k = r->impl->context().get_kernel("solver_kernel", r->n);
    return gfb_kernel_run_from_host(k, static_cast<unsigned> (num_steps), GFB_NUM_STATE + 1, src, dst, chunks);
}
int gfb_rays_trace(gfb_rays *r, size_t num_blocks, size_t sub_steps, double *out) {
    if (!r->compiled) return rays_fail("trace before compile");
    std::vector<uint64_t> keys;
    for (auto &v : r->vars) keys.push_back(reinterpret_cast<uint64_t> (v.get()));
    keys.push_back(reinterpret_cast<uint64_t> (r->impl->residual().get()));
    gfb_ctx *ctx = r->impl->context().device();
    for (size_t b = 0; b < num_blocks; b++) {
        r->impl->step(sub_steps);
//  Records are in the caller's ray order: while the rays are binned the snapshot un-permutes them
//  on its way to the staging buffer (state and residual are both in the order of the last launch).
        if (gfb_snapshot_async(ctx, keys.data(), static_cast<int> (keys.size()), sizeof(double)*r->n,
                               out + b*keys.size()*r->n)) return 1;
    }
    return gfb_wait(ctx);
}
int gfb_rays_absorption_reset(gfb_rays *r) {
    if (!r->compiled) return rays_fail("absorption_reset before compile");
    if (!r->impl->damping) return rays_fail("created without absorption=1");
    r->impl->work().run_side(r->impl->reset_item);
    r->absorption_started = true;
    return 0;
}
int gfb_rays_trace_absorb(gfb_rays *r, size_t num_blocks, size_t sub_steps, double *records, double *absorbed,
                          double *profile, const double *lo, const double *hi, const int *bins) {
    if (!r->compiled) return rays_fail("trace before compile");
    if (!r->impl->damping) return rays_fail("created without absorption=1");
    tracer_base &t = *r->impl;
    gfb_ctx *ctx = t.context().device();
    if (!r->absorption_started && gfb_rays_absorption_reset(r)) return 1;
    std::vector<uint64_t> state_keys, absorb_keys;
    for (auto &v : r->vars) state_keys.push_back(reinterpret_cast<uint64_t> (v.get()));
    state_keys.push_back(reinterpret_cast<uint64_t> (t.residual().get()));
    auto d_power = t.deposition->get_d_power();
    for (auto v : {t.kamp_im, t.power, d_power}) absorb_keys.push_back(reinterpret_cast<uint64_t> (v.get()));

    double *profile_device = nullptr;
    const uint64_t profile_key = reinterpret_cast<uint64_t> (&r->profile_tag);
    size_t cells = 0;
    if (profile) {
        if (!lo || !hi || !bins) return rays_fail("profile requested without lo/hi/bins");
        cells = static_cast<size_t> (bins[0])*bins[1]*bins[2];
        void *p = nullptr;
        if (gfb_buffer(ctx, profile_key, sizeof(double)*cells, nullptr, &p)) return 1;
        profile_device = static_cast<double *> (p);
        if (gfb_copy_h2d(ctx, profile_key, profile, sizeof(double)*cells)) return 1;      // accumulate onto the caller's array
    }
    const double *xd = static_cast<const double *> (t.context().device_pointer(r->vars[GFB_X]));
    const double *yd = static_cast<const double *> (t.context().device_pointer(r->vars[GFB_Y]));
    const double *zd = static_cast<const double *> (t.context().device_pointer(r->vars[GFB_Z]));
    const double *wd = static_cast<const double *> (t.context().device_pointer(d_power));
    for (size_t b = 0; b < num_blocks; b++) {
        t.step(sub_steps);
        t.damping->run();
        t.deposition->run();
        if (profile && gfb_deposit(ctx, xd, yd, zd, wd, r->n, profile_device, lo, hi, bins)) return 1;
        if (records && gfb_snapshot_async(ctx, state_keys.data(), static_cast<int> (state_keys.size()),
                                          sizeof(double)*r->n, records + b*state_keys.size()*r->n)) return 1;
        if (absorbed && gfb_snapshot_async(ctx, absorb_keys.data(), static_cast<int> (absorb_keys.size()),
                                           sizeof(double)*r->n, absorbed + b*absorb_keys.size()*r->n)) return 1;
    }
    if (gfb_wait(ctx)) return 1;
    if (profile && gfb_copy_d2h(ctx, profile_key, profile, sizeof(double)*cells)) return 1;
    return gfb_wait(ctx);
}
int gfb_rays_deposit_block(gfb_rays *r, size_t sub_steps, double *profile_device,
                           const double *lo, const double *hi, const int *bins) {
    if (!r->compiled) return rays_fail("deposit_block before compile");
    if (!r->impl->damping) return rays_fail("created without absorption=1");
    if (!profile_device || !lo || !hi || !bins) return rays_fail("deposit_block needs a device profile with lo/hi/bins");
    tracer_base &t = *r->impl;
    gfb_ctx *ctx = t.context().device();
    if (!r->absorption_started && gfb_rays_absorption_reset(r)) return 1;
    auto d_power = t.deposition->get_d_power();
    if (sub_steps) t.step(sub_steps);
    t.damping->run();
    t.deposition->run();
//  State and d_power are in the same (possibly binned) ray order; the histogram does not care.
    const double *xd = static_cast<const double *> (t.context().device_pointer(r->vars[GFB_X]));
    const double *yd = static_cast<const double *> (t.context().device_pointer(r->vars[GFB_Y]));
    const double *zd = static_cast<const double *> (t.context().device_pointer(r->vars[GFB_Z]));
    const double *wd = static_cast<const double *> (t.context().device_pointer(d_power));
    return gfb_deposit(ctx, xd, yd, zd, wd, r->n, profile_device, lo, hi, bins);
}
int gfb_rays_get_dt(gfb_rays *r, double *out) {
    if (!r->compiled) return rays_fail("get_dt before compile");
    if (!r->impl->step_variable) return rays_fail("get_dt: only solver adaptive_rk4 has a per-ray step");
    r->impl->work().copy_to_host(r->impl->step_variable, out);
    return 0;
}
int gfb_rays_get_absorbed(gfb_rays *r, double *const out[3]) {
    if (!r->compiled) return rays_fail("get_absorbed before compile");
    if (!r->impl->damping) return rays_fail("created without absorption=1");
    tracer_base &t = *r->impl;
    gfb_ctx *ctx = t.context().device();
    const leaf_ptr nodes[3] = {t.kamp_im, t.power, t.deposition->get_d_power()};
    for (int i = 0; i < 3; i++) {
        if (out[i] && gfb_copy_rays_d2h(ctx, reinterpret_cast<uint64_t> (nodes[i].get()), out[i], r->n)) return 1;
    }
    return 0;
}
int gfb_rays_profile(gfb_rays *r, uint64_t *key, size_t *cells) {
    const uint64_t k = reinterpret_cast<uint64_t> (&r->profile_tag);
    size_t bytes = 0;
    if (!r->compiled || gfb_buffer_lookup(r->impl->context().device(), k, nullptr, &bytes)) {
        return rays_fail("gfb_rays_profile: no profile yet (gfb_rays_trace_absorb with a profile creates it)");
    }
    if (key) *key = k;
    if (cells) *cells = bytes/sizeof(double);
    return 0;
}
int gfb_rays_device_ptr(gfb_rays *r, int which, void **device_ptr) {
    if (!r->compiled) return rays_fail("device_ptr before compile");
    r->impl->order().restore();
    if (which < 0 || which > GFB_NUM_STATE) return rays_fail("bad state index");
    *device_ptr = r->impl->context().device_pointer(which == GFB_NUM_STATE ? r->impl->residual() : r->vars[which]);
    return 0;
}
gfb_ctx *gfb_rays_ctx(gfb_rays *r) { return r->impl->context().device(); }
const char *gfb_rays_source(gfb_rays *r) {
    r->source = r->impl->context().get_source();
    return r->source.c_str();
}
int gfb_rays_kernel_stats(gfb_rays *r, int *statements, int *divides, int *reciprocals,
                          int *registers, int *local_bytes, int *smem_bytes) {
    if (!r->compiled) return rays_fail("kernel_stats before compile");
    for (auto &k : r->impl->context().get_kernels()) {
        if (k.name == "solver_kernel") {
            if (statements) *statements = static_cast<int> (k.num_statements);
            if (divides) *divides = static_cast<int> (k.num_divides);
            if (reciprocals) *reciprocals = static_cast<int> (k.num_reciprocals);
            if (smem_bytes) *smem_bytes = static_cast<int> (k.smem_bytes);
            gfb_kernel *h = r->impl->context().get_kernel(k.name, r->n);
            return gfb_kernel_attributes(h, registers, nullptr, local_bytes, nullptr);
        }
    }
    return rays_fail("no solver_kernel");
}
int gfb_rays_rhs(gfb_rays *r, double *const out[7]) {
    workflow::manager<> work(r->device);
    auto outputs = r->impl->rhs();
    work.add_item(r->vars, outputs, {}, graph::shared_random_state<> (), "rhs_kernel", r->n);
    work.compile();
    work.run();
    for (int i = 0; i < 7; i++) {
        if (!out[i]) continue;
        if (outputs[i]->is_constant()) {
            for (size_t j = 0; j < r->n; j++) out[i][j] = outputs[i]->value;
        } else {
            work.copy_to_host(outputs[i], out[i]);
        }
    }
    return 0;
}
}  // extern "C"

//******************************************************************************
//  gfb_boris
//******************************************************************************
struct gfb_boris {
    size_t n;
    std::vector<leaf_ptr> vars;       // x, y, z, ux, uy, uz, gamma
    std::unique_ptr<workflow::manager<>> work;
    double b0 = 0.0, larmor = 0.0;
    bool compiled = false;
//  Particles are kept sorted by the (R, Z) cell of the field tables (binning.hpp).
    equilibrium::cell_grid grid;
    size_t order_period = 1000;
    workflow::ray_order<> order;
    void setup_order() {
        if (grid.dims == 2) order.configure(*work, grid, {vars[0], vars[1], vars[2]}, vars, {}, n, order_period);
    }
};

extern "C" {
gfb_boris *gfb_boris_create(const char *equilibrium_name, const char *table_file, size_t num_particles,
                            double dt_value, int device, const char *options) {
    if (gfb_device_count() <= 0) {
        rays_fail("no CUDA device: the B200 back end has no CPU fallback");
        return nullptr;
    }
    auto eq = make_equilibrium(equilibrium_name, table_file ? table_file : "");
    if (!eq) {
        rays_fail(std::string("unknown equilibrium ") + equilibrium_name);
        return nullptr;
    }
    const run_options o = parse_options(options);
    auto b = std::make_unique<gfb_boris> ();
    b->n = num_particles;
//  xkorc.cpp:40-64
    auto b0 = eq->get_characteristic_field(device);
    b->b0 = b0->evaluate().at(0);
    static const char *symbols[7] = {"x", "y", "z", "u_{x}", "u_{y}", "u_{z}", "\\gamma"};
    for (int i = 0; i < 7; i++) b->vars.push_back(graph::variable(num_particles, symbols[i]));
    auto x = b->vars[0], y = b->vars[1], z = b->vars[2], ux = b->vars[3], uy = b->vars[4], uz = b->vars[5];
    auto gamma = b->vars[6];
//  xkorc.cpp:66-121 (csrc/graph/boris.hpp)
    const boris::push_graph push = boris::build(eq, b->vars, b0, dt_value);
    b->larmor = push.larmor_radius;

    b->work = std::make_unique<workflow::manager<>> (device);
    b->work->get_context().options = o.emit;
    if (o.fused_steps) gfb_set_max_fused_steps(b->work->get_context().device(), o.fused_steps);
    b->work->add_preitem({ux, uy, uz, gamma}, {}, push.initialize, graph::shared_random_state<> (), "initialize_gamma", num_particles);
    b->work->add_item({x, y, z, ux, uy, uz, gamma}, {}, push.step, graph::shared_random_state<> (), "step", num_particles);
//  Particles of one (R, Z) cell share the coefficient rows of the field tables: keep them sorted by
//  cell while stepping (options: bin_rays=0 off, bin_rays=<steps> how often the order is checked).
    if (o.bin_rays != 0) b->grid = eq->get_cell_grid();
    if (o.bin_rays > 0) b->order_period = static_cast<size_t> (o.bin_rays);
    handle_created();
    return b.release();
}
void gfb_boris_destroy(gfb_boris *b) {
    if (!b) return;
    delete b;
    handle_destroyed();
}
int gfb_boris_set_binning(gfb_boris *b, const double *lo, const double *hi, const unsigned *cells, size_t rebin_every) {
    b->order.disable();
    b->grid = equilibrium::cell_grid();
    if (cells && cells[0] && cells[1]) {
        b->grid.dims = 2;
        for (int i = 0; i < 2; i++) {
            b->grid.lo[i] = lo[i];
            b->grid.hi[i] = hi[i];
            b->grid.cells[i] = cells[i];
        }
    }
    b->order_period = rebin_every;
    if (b->compiled) b->setup_order();
    return 0;
}
int gfb_boris_set_state(gfb_boris *b, const double *const state[6]) {
//  After compile every call re-runs initialize_gamma (u <- gamma u), so the momenta must all be the
//  caller's physical u/c again: a partial update would scale the untouched components twice.
    if (b->compiled) {
        for (int i = 0; i < 6; i++) {
            if (!state[i]) return rays_fail("gfb_boris_set_state: after compile all six arrays x, y, z, ux, uy, uz are required");
        }
    }
    b->order.restore();
    for (int i = 0; i < 6; i++) {
        if (!state[i]) continue;
        b->vars[i]->set(std::vector<double> (state[i], state[i] + b->n));
        if (b->compiled) b->work->copy_to_device(b->vars[i], b->vars[i]->data());
    }
    if (b->compiled) b->work->pre_run();
    return 0;
}
int gfb_boris_compile(gfb_boris *b) {
    b->work->compile();
    b->work->pre_run();
    b->compiled = true;
    b->setup_order();
    return 0;
}
int gfb_boris_step(gfb_boris *b, size_t num_steps) {
    if (!b->compiled) return rays_fail("step before compile");
    size_t left = num_steps;
    while (left) {
        const size_t piece = b->order.prepare(left);
        for (size_t i = 0; i < piece; i++) b->work->run();
        if (gfb_flush(b->work->get_context().device())) return 1;
        b->order.advanced(piece);
        left -= piece;
    }
    return 0;
}
int gfb_boris_get_state(gfb_boris *b, double *const state[7]) {
    for (int i = 0; i < 7; i++) {
        if (!state[i]) continue;
        if (b->order.active()) b->order.copy_to_host(b->vars[i], state[i]);     // caller's order, device order kept
        else b->work->copy_to_host(b->vars[i], state[i]);
    }
    return 0;
}
int gfb_boris_info(gfb_boris *b, double *b0, double *larmor_radius) {
    if (b0) *b0 = b->b0;
    if (larmor_radius) *larmor_radius = b->larmor;
    return 0;
}
gfb_ctx *gfb_boris_ctx(gfb_boris *b) { return b->work->get_context().device(); }
}  // extern "C"
