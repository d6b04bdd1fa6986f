// Test helper: build a ray-tracing case with the host front end and dump the emitted
// CUDA source (and packed table groups) without needing a device.
//   emit_case <dispersion> <equilibrium> <kind: rk4|rk2|rk4_graph|newton|rhs> <out.cu> [tables.bin] [efit.gfbt]
#include <fstream>
#include <iostream>
#include "../../graph_framework_b200/csrc/graph/graph_framework.hpp"
#include "../../graph_framework_b200/csrc/graph/boris.hpp"

using graph::leaf_ptr;

static equilibrium::shared<> make_eq(const std::string &name, const std::string &efit) {
    if (name == "efit") return equilibrium::make_efit<> (efit);
    if (name == "vmec") return equilibrium::make_vmec<> (std::getenv("GFB_VMEC_FILE") ? std::getenv("GFB_VMEC_FILE") : "tests/golden/vmec.gfbt");
    if (name == "slab") return equilibrium::make_slab<> ();
    if (name == "slab_density") return equilibrium::make_slab_density<> ();
    if (name == "slab_field") return equilibrium::make_slab_field<> ();
    if (name == "no_magnetic_field") return equilibrium::make_no_magnetic_field<> ();
    if (name == "gaussian_density") return equilibrium::make_gaussian_density<> ();
    std::cerr << "unknown equilibrium " << name << std::endl;
    std::exit(2);
}

template<class DF>
static jit::kernel_info build(const std::string &kind, equilibrium::shared<> eq, std::ostringstream &src,
                              const jit::emit_options &opt) {
    const size_t n = 1;
    auto w = graph::variable(n, "w"), kx = graph::variable(n, "kx"), ky = graph::variable(n, "ky"), kz = graph::variable(n, "kz");
    auto x = graph::variable(n, "x"), y = graph::variable(n, "y"), z = graph::variable(n, "z"), t = graph::variable(n, "t");
    graph::input_nodes<> inputs = {t, w, x, y, z, kx, ky, kz};
    auto dt = graph::constant(std::getenv("GFB_DT") ? std::atof(std::getenv("GFB_DT")) : 1.0e-3);
    dispersion::dispersion_interface<DF> D(w, kx, ky, kz, x, y, z, t, eq);
    if (kind == "rk4" || kind == "rk2") {
        return jit::emit_runge_kutta(src, opt, kind == "rk4" ? jit::kernel_kind::rk4 : jit::kernel_kind::rk2, "solver_kernel",
                                     inputs, {kx, ky, kz, x, y, z},
                                     {D.get_dkxdt(), D.get_dkydt(), D.get_dkzdt(), D.get_dxdt(), D.get_dydt(), D.get_dzdt()},
                                     t, dt, D.get_residual(), n);
    } else if (kind == "newton" || kind == "loss") {
        auto f = D.get_d();
        leaf_ptr var = kx;
        if (const char *v = std::getenv("GFB_NEWTON_VAR")) {
            const std::string name = v;
            var = name == "x" ? x : name == "y" ? y : name == "z" ? z : name == "ky" ? ky : name == "kz" ? kz : name == "w" ? w : kx;
        }
        graph::map_nodes<> setters = {{var - 1.0*f/f->df(var), var}};
        return jit::emit_item(src, opt, kind == "newton" ? jit::kernel_kind::newton : jit::kernel_kind::generic,
                              "loss_kernel", inputs, {f*f}, setters, n);
    } else if (kind == "rhs") {
        return jit::emit_item(src, opt, jit::kernel_kind::generic, "rhs_kernel", inputs,
                              {D.get_dxdt(), D.get_dydt(), D.get_dzdt(), D.get_dkxdt(), D.get_dkydt(), D.get_dkzdt(), D.get_d()},
                              {}, n);
    }
    std::cerr << "unknown kind " << kind << std::endl;
    std::exit(2);
}

//  The two absorption kernels (absorption.hpp): emitted through a manager-less path so that no
//  device is needed.  Argument orders match absorption::weak_damping / absorption::power_item.
static jit::kernel_info build_absorption(const std::string &kind, equilibrium::shared<> eq, std::ostringstream &src,
                                         const jit::emit_options &opt) {
    const size_t n = 1;
    auto w = graph::variable(n, "w"), kx = graph::variable(n, "kx"), ky = graph::variable(n, "ky"), kz = graph::variable(n, "kz");
    auto x = graph::variable(n, "x"), y = graph::variable(n, "y"), z = graph::variable(n, "z"), t = graph::variable(n, "t");
    if (kind == "kamp") {
        auto kamp_re = graph::variable(n, "kamp_re"), kamp_im = graph::variable(n, "kamp_im");
        auto k_vec = kx*eq->get_esup1(x, y, z) + ky*eq->get_esup2(x, y, z) + kz*eq->get_esup3(x, y, z);
        auto Dc = dispersion::cold_plasma_expansion<> ().D(w, k_vec, x, y, z, t, eq);
        auto Dw = dispersion::hot_plasma_expansion<> ().D_complex(w, k_vec, x, y, z, t, eq);
        auto grad = graph::gradient(Dc, {kx, ky, kz});
        auto slope = k_vec->unit()->dot(grad[0]*eq->get_esup1(x, y, z) + grad[1]*eq->get_esup2(x, y, z) +
                                        grad[2]*eq->get_esup3(x, y, z));
        graph::map_nodes<> setters = {{k_vec->length() - Dw.re/slope, kamp_re}, {-1.0*(Dw.im/slope), kamp_im}};
        return jit::emit_item(src, opt, jit::kernel_kind::generic, "weak_damping_kimg_kernel",
                              {kamp_re, kamp_im, kx, ky, kz, x, y, z, t, w}, {}, setters, n);
    }
    auto x_last = graph::variable(n, "x_last"), y_last = graph::variable(n, "y_last"), z_last = graph::variable(n, "z_last");
    auto kamp = graph::variable(n, "kamp"), power = graph::variable(n, "power"), k_sum = graph::variable(n, "k_sum");
    auto dl = graph::vector(eq->get_x(x, y, z) - eq->get_x(x_last, y_last, z_last),
                            eq->get_y(x, y, z) - eq->get_y(x_last, y_last, z_last),
                            eq->get_z(x, y, z) - eq->get_z(x_last, y_last, z_last))->length();
    auto p_next = graph::exp(-2.0*k_sum);
    auto difference = p_next - power;
    graph::map_nodes<> setters = {{x, x_last}, {y, y_last}, {z, z_last}, {p_next, power}, {kamp*dl + k_sum, k_sum}};
    return jit::emit_item(src, opt, jit::kernel_kind::generic, "power",
                          {x, y, z, x_last, y_last, z_last, kamp, power, k_sum},
                          {graph::sqrt(difference*difference)}, setters, n);
}

//  The two kernels of the reference's particle-in-cell example (graph_pic/xpic.cpp:63-125) at test
//  size: an RK4 particle push that gathers the field with index_1D, and the field accumulation
//  that walks the particle array with index_1D on a running index (batch of 4 per step).
static jit::kernel_info build_pic(const std::string &kind, std::ostringstream &src, const jit::emit_options &opt) {
    const size_t n = 64;            // the clamp of an index node is the length of the indexed variable
    const double dt = 1.0e-2, scale = 2.0/63.0, offset = -1.0;
    auto x = graph::variable(n, "x"), vpara = graph::variable(n, "v");
    auto epara = graph::variable(n, "e"), dens = graph::variable(n, "n");
    auto grid_position = graph::variable(n, "xi"), particle_index = graph::variable(n, "i");
    if (kind == "pic_push") {
        auto x1 = dt*vpara;
        auto v1 = -1.0*graph::index_1D(epara, x, scale, offset);
        auto x2 = dt*(vpara + v1/2.0);
        auto v2 = -1.0*graph::index_1D(epara, x + x1/2.0, scale, offset);
        auto x3 = dt*(vpara + v2/2.0);
        auto v3 = -1.0*graph::index_1D(epara, x + x2/2.0, scale, offset);
        auto x4 = dt*(vpara + v3);
        auto v4 = -1.0*graph::index_1D(epara, x + x3, scale, offset);
        graph::map_nodes<> setters = {{x + (x1 + 2.0*(x2 + x3) + x4)/6.0, x}, {vpara + (v1 + 2.0*(v2 + v3) + v4)/6.0, vpara}};
        return jit::emit_item(src, opt, jit::kernel_kind::generic, "Particle_Push", {x, vpara, epara}, {}, setters, n);
    }
    auto density = [] (leaf_ptr d) { return graph::exp(d*d/-0.01); };
    auto next_index = particle_index, next_e = epara, next_n = dens;
    for (int b = 0; b < 4; b++) {
        auto particle = graph::index_1D(x, next_index, 1.0, 0.0);
        next_index = next_index + 1.0;
        auto d = particle - grid_position;
        auto nd = density(d);
        next_e = next_e + -1.0/nd*nd->df(d);
        next_n = next_n + nd;
    }
    graph::map_nodes<> setters = {{next_e, epara}, {next_index, particle_index}, {next_n, dens}};
    return jit::emit_item(src, opt, jit::kernel_kind::generic, "Compute_efield",
                          {epara, dens, grid_position, particle_index, x}, {}, setters, n);
}

int main(int argc, char **argv) {
    if (argc < 5) { std::cerr << "usage: emit_case <dispersion> <equilibrium> <kind> <out.cu> [tables.bin] [efit.gfbt]" << std::endl; return 2; }
    const std::string d = argv[1], e = argv[2], kind = argv[3];
    const std::string efit = argc > 6 ? argv[6] : "tests/golden/efit.gfbt";
    auto eq = make_eq(e, efit);
    jit::emit_options opt;
    if (const char *s = std::getenv("GFB_NO_STAGE")) opt.stage_tables = false;
    if (const char *s = std::getenv("GFB_NO_RCP")) opt.share_reciprocals = false;
    if (const char *s = std::getenv("GFB_NO_FASTDIV")) opt.fast_division = false;
    if (const char *s = std::getenv("GFB_BLOCK")) opt.block_size = std::atoi(s);
    if (const char *s = std::getenv("GFB_MINB")) opt.min_blocks = std::atoi(s);
    std::ostringstream src;
    jit::kernel_info info;
    if (kind == "boris") {
//  The xkorc push with the on-axis field given (GFB_B0; the device search for it is tested on the GPU).
        const size_t n = 1;
        std::vector<leaf_ptr> vars;
        for (const char *name : {"x", "y", "z", "ux", "uy", "uz", "gamma"}) vars.push_back(graph::variable(n, name));
        const auto push = boris::build(eq, vars, graph::constant(std::atof(std::getenv("GFB_B0") ? std::getenv("GFB_B0") : "1.0")),
                                       std::getenv("GFB_DT") ? std::atof(std::getenv("GFB_DT")) : 0.5);
        info = jit::emit_item(src, opt, jit::kernel_kind::generic, "step", vars, {}, push.step, n);
    } else if (kind == "kamp" || kind == "power") info = build_absorption(kind, eq, src, opt);
    else if (kind == "pic_push" || kind == "pic_field") info = build_pic(kind, src, opt);
    else if (d == "cold_plasma") info = build<dispersion::cold_plasma<>> (kind, eq, src, opt);
    else if (d == "ordinary_wave") info = build<dispersion::ordinary_wave<>> (kind, eq, src, opt);
    else if (d == "extra_ordinary_wave") info = build<dispersion::extra_ordinary_wave<>> (kind, eq, src, opt);
    else if (d == "bohm_gross") info = build<dispersion::bohm_gross<>> (kind, eq, src, opt);
    else if (d == "simple") info = build<dispersion::simple<>> (kind, eq, src, opt);
    else if (d == "light_wave") info = build<dispersion::light_wave<>> (kind, eq, src, opt);
    else { std::cerr << "unknown dispersion " << d << std::endl; return 2; }
    std::ofstream(argv[4]) << src.str();
    if (argc > 5) {
        std::ofstream tb(argv[5], std::ios::binary);
        for (auto &g : info.groups) {
            if (g.alias_input >= 0) continue;
            const uint64_t n = g.packed.size();
            tb.write(reinterpret_cast<const char *> (&n), 8);
            tb.write(reinterpret_cast<const char *> (g.packed.data()), 8*n);
        }
    }
    std::cout << "{\"kernel\": \"" << info.name << "\", \"statements\": " << info.num_statements
              << ", \"divides\": " << info.num_divides << ", \"reciprocals\": " << info.num_reciprocals
              << ", \"groups\": [";
    for (size_t g = 0; g < info.groups.size(); g++) {
        auto &grp = info.groups[g];
        std::cout << (g ? ", " : "") << "{\"members\": " << grp.members.size() << ", \"cells\": " << grp.cells
                  << ", \"stride\": " << grp.stride << ", \"staged\": " << (grp.staged ? "true" : "false")
                  << ", \"bytes\": " << grp.bytes() << "}";
    }
    std::cout << "], \"smem_bytes\": " << info.smem_bytes << "}" << std::endl;
    return 0;
}
