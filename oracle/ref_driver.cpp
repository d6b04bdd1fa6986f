// Reference-oracle driver.  TEST INFRASTRUCTURE -- not product code.
//
// Compiled in the build container against the UNMODIFIED reference headers under
// /root/reference/graph_framework (nothing is copied into this repo) together
// with oracle/ref_shim/{gxx_cpu_context.hpp,netcdf.h}.  The binary lands in
// oracle/_ref/ (git-ignored, travels to the GPU box) and needs only g++ at run
// time.  It runs the reference's own graph -> reduce -> df -> emit pipeline and
// its own solver::rk4 / dispersion_interface::solve, so its outputs ARE the
// reference cpu path's results up to the C++ compiler used for the kernel text.
//
// usage:
//   ref_driver trace <dispersion> <equilibrium> <solver> <N> <dt> <nsteps> <save_every> <init:kx|ky|kz|none> <in.bin> <out.bin>
//       in.bin : 8 arrays of N doubles  t,w,x,y,z,kx,ky,kz
//       out.bin: record 0 = state after init (9 arrays: t,w,x,y,z,kx,ky,kz,residual[=0]),
//                then one record after every <save_every> steps (residual = D^2 at
//                the pre-step state of the last step, solver.hpp:316-319).
//       solver adaptive_rk4: dt is the initial value of the per-ray step variable; every record holds a tenth
//       array, the current dt.
//   ref_driver rhs <dispersion> <equilibrium> <N> <in.bin> <out.bin>
//       out.bin: 7 arrays dxdt,dydt,dzdt,dkxdt,dkydt,dkzdt,D  (dispersion.hpp:1387-1433)
//   ref_driver bench <dispersion> <equilibrium> <N> <dt> <nsteps> <threads> <in.bin|-> [blocks]
//       prints one JSON line; rays are split batch/extra like xrays_bench.cpp:38-51.  With
//       [blocks] > 1 the stepping phase is repeated that many times after one setup/init/compile
//       and every block is timed separately ("block_s").
//   ref_driver korc <equilibrium> <N> <nsteps> <in.bin> <out.bin>
//       in.bin: 6 arrays x,y,z,ux,uy,uz (physical u/c, as xkorc.cpp:47-64); out: 7 arrays x,y,z,ux,uy,uz,gamma
//   ref_driver source <dispersion> <equilibrium> <solver>     (dump kernel text; set GFB_ORACLE_KEEP_SOURCE=1)
//   ref_driver absorb <equilibrium> <N> <nrec> <in.bin> <out.bin>
//       in.bin : nrec records of 8 arrays t,w,x,y,z,kx,ky,kz (a trajectory as written by trace)
//       out.bin: nrec records of 4 arrays  Re kamp, Im kamp, power, d_power.  kamp is the setter of
//                absorption::weak_damping (absorption.hpp:395-412) evaluated through the reference's
//                own JIT path with T = std::complex<double>, SAFE_MATH = true as xrays.cpp:1099-1104
//                does; power/d_power follow the bin_power stage (xrays.cpp:674-793, T = double) fed
//                with Im kamp (reference_imag_variable, xrays.cpp:743).  Record 0 holds power = 1,
//                d_power = 0 (the state before the first bin_power kernel).
//   ref_driver cells <N> <scale> <offset> <ncells> <in.bin> <out.bin>
//       which table cell the reference's compiled kernel picks: piecewise_1D over the table
//       [0, 1, ..., ncells - 1] evaluated through its own JIT path (g++ -O3 -ffast-math, as
//       cpu_context.hpp:155-157 compiles) at N arguments; out = N doubles (the cell numbers).
//   ref_driver reducer
//       the reference's reducer defect behind its wrong cold-plasma dD/dz, with four plain variables:
//       ((A W)^2 B)/(C^2 W^4) built from graph nodes, evaluated by the reference's host evaluate(), next
//       to the same arithmetic in plain doubles; plus the cut-down expression of the ray equations
//       d/dz [ (z Q - x S)^2 / ((1 + z^2) W^2) ]  against a central difference.  Prints one JSON line.
//   ref_driver erfi <N> <in.bin> <out.bin>
//       in: N doubles x; out: 2 arrays special::w_im(x), Re special::erfi(x + 0i) (special_functions.hpp)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "solver.hpp"
#include "special_functions.hpp"
#include "timing.hpp"

#ifndef GFB_EFIT_FILE
#error "define GFB_EFIT_FILE"
#endif

typedef double T;
using leaf = graph::shared_leaf<T>;

static std::string efit_path() {
    if (const char *e = std::getenv("GFB_EFIT_FILE")) return e;
    return GFB_EFIT_FILE;
}
static std::string vmec_path() {
    if (const char *e = std::getenv("GFB_VMEC_FILE")) return e;
    return "tests/golden/vmec.gfbt";
}

static equilibrium::shared<T> make_eq(const std::string &name) {
    if (name == "efit") return equilibrium::make_efit<T> (efit_path());
    if (name == "vmec") return equilibrium::make_vmec<T> (vmec_path());
    if (name == "slab") return equilibrium::make_slab<T> ();
    if (name == "slab_density") return equilibrium::make_slab_density<T> ();
    if (name == "slab_field") return equilibrium::make_slab_field<T> ();
    if (name == "no_magnetic_field") return equilibrium::make_no_magnetic_field<T> ();
    if (name == "gaussian_density") return equilibrium::make_gaussian_density<T> ();
    std::cerr << "unknown equilibrium " << name << std::endl;
    std::exit(2);
}

static std::vector<std::vector<double>> read_arrays(const std::string &path, size_t narr, size_t n) {
    std::vector<std::vector<double>> a(narr, std::vector<double> (n));
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::cerr << "cannot open " << path << std::endl; std::exit(2); }
    for (auto &v : a) f.read(reinterpret_cast<char *> (v.data()), sizeof(double)*n);
    if (!f) { std::cerr << "short read " << path << std::endl; std::exit(2); }
    return a;
}

struct state_vars {
    leaf t, w, x, y, z, kx, ky, kz;
    state_vars(const size_t n) :
    t(graph::variable<T> (n, "t")), w(graph::variable<T> (n, "\\omega")),
    x(graph::variable<T> (n, "x")), y(graph::variable<T> (n, "y")), z(graph::variable<T> (n, "z")),
    kx(graph::variable<T> (n, "k_{x}")), ky(graph::variable<T> (n, "k_{y}")), kz(graph::variable<T> (n, "k_{z}")) {}
    std::vector<leaf> all() { return {t, w, x, y, z, kx, ky, kz}; }
    void set(const std::vector<std::vector<double>> &a, const size_t offset, const size_t n) {
        auto v = all();
        for (size_t k = 0; k < 8; k++) {
            graph::variable_cast(v[k])->set(std::vector<T> (a[k].begin() + offset, a[k].begin() + offset + n));
        }
    }
    void set_bench_defaults() {      // xrays_bench.cpp:62-71
        graph::variable_cast(t)->set(0.0);  graph::variable_cast(w)->set(500.0);
        graph::variable_cast(x)->set(2.5);  graph::variable_cast(y)->set(0.0);  graph::variable_cast(z)->set(0.0);
        graph::variable_cast(kx)->set(-600.0); graph::variable_cast(ky)->set(0.0); graph::variable_cast(kz)->set(0.0);
    }
    leaf by_name(const std::string &s) {
        if (s == "kx") return kx; if (s == "ky") return ky; if (s == "kz") return kz;
        if (s == "w") return w; if (s == "x") return x; if (s == "y") return y; if (s == "z") return z;
        return leaf();
    }
};

static void write_record(std::ofstream &f, state_vars &s, const std::vector<double> &residual) {
    for (auto &v : s.all()) {
        auto var = graph::variable_cast(v);
        f.write(reinterpret_cast<const char *> (var->data()), sizeof(double)*var->size());
    }
    f.write(reinterpret_cast<const char *> (residual.data()), sizeof(double)*residual.size());
}

template<class SOLVER>
static int trace_impl(int argc, char **argv) {
    const std::string eqn = argv[3];
    const size_t n = std::stoul(argv[5]);
    const double dt = std::stod(argv[6]);
    const size_t nsteps = std::stoul(argv[7]);
    const size_t every = std::stoul(argv[8]);
    const std::string init = argv[9];
    auto in = read_arrays(argv[10], 8, n);
    std::ofstream out(argv[11], std::ios::binary);

    state_vars s(n);
    s.set(in, 0, n);
    auto eq = make_eq(eqn);
    auto dtc = graph::constant<T> (dt);
    SOLVER solve(s.w, s.kx, s.ky, s.kz, s.x, s.y, s.z, s.t, dtc, eq, "", n, 0);
    if (init == "none") solve.init();
    else solve.init(s.by_name(init));
    std::vector<double> residual(n, 0.0);
    write_record(out, s, residual);
    solve.compile();
    for (size_t i = 1; i <= nsteps; i++) {
        solve.step();
        if (i%every == 0 || i == nsteps) {
            solve.sync_host();
            for (size_t r = 0; r < n; r++) residual[r] = solve.check_residual(r);
            write_record(out, s, residual);
        }
    }
    return 0;
}

//  solver::adaptive_rk4 (solver.hpp:881-1006): dt is a per-ray VARIABLE that the solver's own Newton item
//  rewrites before every step; records carry it as a tenth array.
template<class D>
static int trace_adaptive_impl(int argc, char **argv) {
    const std::string eqn = argv[3];
    const size_t n = std::stoul(argv[5]);
    const double dt = std::stod(argv[6]);
    const size_t nsteps = std::stoul(argv[7]);
    const size_t every = std::stoul(argv[8]);
    const std::string init = argv[9];
    auto in = read_arrays(argv[10], 8, n);
    std::ofstream out(argv[11], std::ios::binary);
    state_vars s(n);
    s.set(in, 0, n);
    auto eq = make_eq(eqn);
    auto dtv = graph::variable<T> (n, "dt");
    graph::variable_cast(dtv)->set(static_cast<T> (dt));
    struct open_solver : public solver::adaptive_rk4<D> {           // `work` is protected; nothing else is touched
        using solver::adaptive_rk4<D>::adaptive_rk4;
        void read(leaf node, T *destination) { this->work.copy_to_host(node, destination); }
    };
    open_solver solve(s.w, s.kx, s.ky, s.kz, s.x, s.y, s.z, s.t, dtv, eq, "", n, 0);
    if (init == "none") solve.init();
    else solve.init(s.by_name(init));
    std::vector<double> residual(n, 0.0);
    auto record = [&] {
        write_record(out, s, residual);
        auto var = graph::variable_cast(dtv);
        out.write(reinterpret_cast<const char *> (var->data()), sizeof(double)*n);
    };
    record();
    solve.compile();
    for (size_t i = 1; i <= nsteps; i++) {
        solve.step();
        if (i%every == 0 || i == nsteps) {
            solve.sync_host();
            solve.read(dtv, graph::variable_cast(dtv)->data());
            for (size_t r = 0; r < n; r++) residual[r] = solve.check_residual(r);
            record();
        }
    }
    return 0;
}

template<class D>
static int rhs_impl(int argc, char **argv) {
    const std::string eqn = argv[3];
    const size_t n = std::stoul(argv[4]);
    auto in = read_arrays(argv[5], 8, n);
    std::ofstream out(argv[6], std::ios::binary);
    state_vars s(n);
    s.set(in, 0, n);
    auto eq = make_eq(eqn);
    dispersion::dispersion_interface<D> di(s.w, s.kx, s.ky, s.kz, s.x, s.y, s.z, s.t, eq);
    graph::input_nodes<T> inputs;
    for (auto &v : s.all()) inputs.push_back(graph::variable_cast(v));
    graph::output_nodes<T> outputs = {di.get_dxdt(), di.get_dydt(), di.get_dzdt(),
                                      di.get_dkxdt(), di.get_dkydt(), di.get_dkzdt(), di.get_d()};
    workflow::manager<T> work(0);
    work.add_item(inputs, outputs, {}, graph::shared_random_state<T> (), "rhs_kernel", n);
    work.compile();
    work.run();
    std::vector<double> buf(n);
    for (auto &o : outputs) {
        work.copy_to_host(o, buf.data());
        out.write(reinterpret_cast<const char *> (buf.data()), sizeof(double)*n);
    }
    return 0;
}

template<class SOLVER>
static int bench_impl(int argc, char **argv) {
    const std::string eqn = argv[3];
    const size_t n = std::stoul(argv[4]);
    const double dt = std::stod(argv[5]);
    const size_t nsteps = std::stoul(argv[6]);
    const size_t nthreads = std::max<size_t> (1, std::min<size_t> (std::stoul(argv[7]), n));
    const std::string inpath = argv[8];
    const size_t blocks = argc > 9 ? std::max<size_t> (1, std::stoul(argv[9])) : 1;
    std::vector<std::vector<double>> in;
    if (inpath != "-") in = read_arrays(inpath, 8, n);
    const size_t batch = n/nthreads, extra = n%nthreads;   // xrays_bench.cpp:38-51
    std::vector<double> t_setup(nthreads), t_init(nthreads), t_compile(nthreads), t_steps(nthreads);
    std::vector<std::vector<double>> t_block(nthreads, std::vector<double> (blocks, 0.0));
    std::vector<std::thread> threads(nthreads);
    std::vector<size_t> offsets(nthreads + 1, 0);
    for (size_t i = 0; i < nthreads; i++) offsets[i + 1] = offsets[i] + batch + (extra > i ? 1 : 0);
    for (size_t i = 0; i < nthreads; i++) {
        threads[i] = std::thread([&, i] () {
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto secs = [] (auto a, auto b) { return std::chrono::duration<double> (b - a).count(); };
            const size_t local = offsets[i + 1] - offsets[i];
            auto t0 = now();
            state_vars s(local);
            if (in.empty()) s.set_bench_defaults(); else s.set(in, offsets[i], local);
            auto eq = make_eq(eqn);
            auto dtc = graph::constant<T> (dt);
            SOLVER solve(s.w, s.kx, s.ky, s.kz, s.x, s.y, s.z, s.t, dtc, eq, "", local, i);
            auto t1 = now();
            solve.init(s.kx);
            auto t2 = now();
            solve.compile();
            auto t3 = now();
            for (size_t b = 0; b < blocks; b++) {
                auto b0 = now();
                for (size_t j = 0; j < nsteps; j++) solve.step();
                solve.sync_host();
                t_block[i][b] = secs(b0, now());
            }
            auto t4 = now();
            t_setup[i] = secs(t0, t1); t_init[i] = secs(t1, t2); t_compile[i] = secs(t2, t3); t_steps[i] = secs(t3, t4);
        });
    }
    for (auto &t : threads) t.join();
    auto mx = [] (const std::vector<double> &v) { double m = 0; for (double x : v) m = std::max(m, x); return m; };
    const double steps_s = mx(t_steps)/static_cast<double> (blocks);
    std::printf("{\"impl\": \"reference\", \"rays\": %zu, \"steps\": %zu, \"threads\": %zu, \"setup_s\": %.4f, \"init_s\": %.4f, "
                "\"compile_s\": %.4f, \"steps_s\": %.6f, \"ray_steps_per_s\": %.6e, \"block_s\": [",
                n, nsteps, nthreads, mx(t_setup), mx(t_init), mx(t_compile), steps_s,
                static_cast<double> (n)*static_cast<double> (nsteps)/steps_s);
    for (size_t b = 0; b < blocks; b++) {
        double worst = 0.0;
        for (size_t i = 0; i < nthreads; i++) worst = std::max(worst, t_block[i][b]);
        std::printf("%s%.6f", b ? ", " : "", worst);
    }
    std::printf("]}\n");
    return 0;
}

//  Boris push exactly as graph_korc/xkorc.cpp:40-121 builds it.
static int korc(int argc, char **argv) {
    const std::string eqn = argv[2];
    const size_t n = std::stoul(argv[3]);
    const size_t nsteps = std::stoul(argv[4]);
    auto in = read_arrays(argv[5], 6, n);
    std::ofstream out(argv[6], std::ios::binary);

    auto eq = make_eq(eqn);
    auto b0 = eq->get_characteristic_field(0);
    const T q = 1.602176634E-19;
    const T me = 9.1093837139E-31;
    const T c = 299792458.0;
    auto gryo_period = me/(q*b0);
    auto larmor_radius = c*gryo_period;

    auto ux = graph::variable<T> (n, "u_{x}");
    auto uy = graph::variable<T> (n, "u_{y}");
    auto uz = graph::variable<T> (n, "u_{z}");
    auto x = graph::variable<T> (n, "x");
    auto y = graph::variable<T> (n, "y");
    auto z = graph::variable<T> (n, "z");
    graph::variable_cast(x)->set(in[0]);  graph::variable_cast(y)->set(in[1]);  graph::variable_cast(z)->set(in[2]);
    graph::variable_cast(ux)->set(in[3]); graph::variable_cast(uy)->set(in[4]); graph::variable_cast(uz)->set(in[5]);
    auto pos = graph::vector(x, y, z);
    auto u_vec = graph::vector(ux, uy, uz);
    auto gamma = graph::variable<T> (n, "\\gamma");
    auto dt = graph::constant<T> (0.5);
    auto gamma_init = 1.0/graph::sqrt(1.0 - u_vec->dot(u_vec));
    auto u_init = gamma_init*u_vec;
    auto b_vec = eq->get_magnetic_field(pos->get_x(), pos->get_y(), pos->get_z())/b0;

    workflow::manager<T> work(0);
    work.add_preitem({graph::variable_cast(ux), graph::variable_cast(uy), graph::variable_cast(uz), graph::variable_cast(gamma)}, {}, {
        {u_init->get_x(), graph::variable_cast(ux)},
        {u_init->get_y(), graph::variable_cast(uy)},
        {u_init->get_z(), graph::variable_cast(uz)},
        {gamma_init, graph::variable_cast(gamma)}
    }, graph::shared_random_state<T> (), "initialize_gamma", n);

    auto u_prime = u_vec - dt*u_vec->cross(b_vec)/(2.0*gamma);
    auto tau = -0.5*dt*b_vec;
    auto tau_sq = tau->dot(tau);
    auto speed_sq = u_prime->dot(u_prime);
    auto sigma = 1.0 + speed_sq - tau_sq;
    auto ustar = u_prime->dot(tau);
    auto gamma_next = graph::sqrt(0.5*(sigma + graph::sqrt(sigma*sigma + 4.0*(tau_sq + ustar*ustar))));
    auto t = tau/gamma_next;
    auto s = 1.0 + t->dot(t);
    auto u_prime_dot_t = u_prime->dot(t);
    auto u_next = (u_prime + u_prime_dot_t*t + u_prime->cross(t))/s;
    auto pos_next = pos + larmor_radius*dt*u_next/gamma_next;

    work.add_item({graph::variable_cast(x), graph::variable_cast(y), graph::variable_cast(z),
                   graph::variable_cast(ux), graph::variable_cast(uy), graph::variable_cast(uz),
                   graph::variable_cast(gamma)}, {}, {
        {pos_next->get_x(), graph::variable_cast(x)},
        {pos_next->get_y(), graph::variable_cast(y)},
        {pos_next->get_z(), graph::variable_cast(z)},
        {u_next->get_x(), graph::variable_cast(ux)},
        {u_next->get_y(), graph::variable_cast(uy)},
        {u_next->get_z(), graph::variable_cast(uz)},
        {gamma_next, graph::variable_cast(gamma)}
    }, graph::shared_random_state<T> (), "step", n);
    work.compile();
    work.pre_run();
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < nsteps; i++) work.run();
    work.wait();
    const double secs = std::chrono::duration<double> (std::chrono::steady_clock::now() - t0).count();
    std::vector<double> buf(n);
    for (auto v : std::vector<leaf> {x, y, z, ux, uy, uz, gamma}) {
        work.copy_to_host(v, buf.data());
        out.write(reinterpret_cast<const char *> (buf.data()), sizeof(double)*n);
    }
    std::printf("{\"impl\": \"reference\", \"b0\": %.17g, \"larmor_radius\": %.17g, \"particles\": %zu, \"steps\": %zu, \"steps_s\": %.6f, \"particle_steps_per_s\": %.6e}\n",
                b0->evaluate().at(0), larmor_radius->evaluate().at(0), n, nsteps, secs,
                static_cast<double> (n)*static_cast<double> (nsteps)/secs);
    return 0;
}

//  The weak damping stage followed by the power stage, as xrays.cpp main() chains them.
//  (GFB_REF_DRIVER_REAL_ONLY: builds of this driver on a device context that implements the FP64
//  path only -- integration/Makefile -- leave the complex<double> stage out.)
#ifndef GFB_REF_DRIVER_REAL_ONLY
static int absorb(int argc, char **argv) {
    typedef std::complex<double> C;
    constexpr bool SAFE = true;
    const std::string eqn = argv[2];
    const size_t n = std::stoul(argv[3]);
    const size_t nrec = std::stoul(argv[4]);
    auto in = read_arrays(argv[5], 8*nrec, n);
    std::ofstream out(argv[6], std::ios::binary);

    std::vector<std::vector<double>> kre(nrec, std::vector<double> (n)), kim(nrec, std::vector<double> (n));
    {
        auto t = graph::variable<C, SAFE> (n, "t");
        auto w = graph::variable<C, SAFE> (n, "\\omega");
        auto x = graph::variable<C, SAFE> (n, "x");
        auto y = graph::variable<C, SAFE> (n, "y");
        auto z = graph::variable<C, SAFE> (n, "z");
        auto kx = graph::variable<C, SAFE> (n, "k_{x}");
        auto ky = graph::variable<C, SAFE> (n, "k_{y}");
        auto kz = graph::variable<C, SAFE> (n, "k_{z}");
        auto kamp = graph::variable<C, SAFE> (n, "kamp");
        equilibrium::shared<C, SAFE> eq;
        if (eqn == "efit") eq = equilibrium::make_efit<C, SAFE> (efit_path());
        else if (eqn == "slab_density") eq = equilibrium::make_slab_density<C, SAFE> ();
        else if (eqn == "slab") eq = equilibrium::make_slab<C, SAFE> ();
        else { std::cerr << "absorb: unsupported equilibrium " << eqn << std::endl; return 2; }

//  absorption.hpp:395-412, the weak_damping constructor body.
        auto k_vec = kx*eq->get_esup1(x, y, z) + ky*eq->get_esup2(x, y, z) + kz*eq->get_esup3(x, y, z);
        auto k_unit = k_vec->unit();
        auto Dc = dispersion::cold_plasma_expansion<C, SAFE> ().D(w, k_vec, x, y, z, t, eq);
        auto Dw = dispersion::hot_plasma_expansion<C, dispersion::z_erfi<C, SAFE>, SAFE> ().D(w, k_vec, x, y, z, t, eq);
        auto kamp1 = k_vec->length() - Dw/k_unit->dot(Dc->df(kx)*eq->get_esup1(x, y, z) +
                                                      Dc->df(ky)*eq->get_esup2(x, y, z) +
                                                      Dc->df(kz)*eq->get_esup3(x, y, z));
        graph::input_nodes<C, SAFE> inputs = {
            graph::variable_cast(kamp), graph::variable_cast(kx), graph::variable_cast(ky), graph::variable_cast(kz),
            graph::variable_cast(x), graph::variable_cast(y), graph::variable_cast(z), graph::variable_cast(t),
            graph::variable_cast(w)
        };
        graph::map_nodes<C, SAFE> setters = {{kamp1, graph::variable_cast(kamp)}};
        workflow::manager<C, SAFE> work(0);
        work.add_item(inputs, {}, setters, graph::shared_random_state<C, SAFE> (), "weak_damping_kimg_kernel", n);
        work.compile();

        std::vector<graph::shared_leaf<C, SAFE>> order = {t, w, x, y, z, kx, ky, kz};
        std::vector<C> buffer(n);
        for (size_t j = 0; j < nrec; j++) {
            for (size_t k = 0; k < 8; k++) {
                for (size_t i = 0; i < n; i++) buffer[i] = C(in[8*j + k][i], 0.0);
                graph::variable_cast(order[k])->set(buffer);
                work.copy_to_device(order[k], graph::variable_cast(order[k])->data());
            }
            work.run();
            work.wait();
            work.copy_to_host(kamp, buffer.data());
            for (size_t i = 0; i < n; i++) { kre[j][i] = std::real(buffer[i]); kim[j][i] = std::imag(buffer[i]); }
        }
    }

//  xrays.cpp:693-776, the bin_power stage.
    std::vector<std::vector<double>> power(nrec, std::vector<double> (n, 1.0)), d_power_rec(nrec, std::vector<double> (n, 0.0));
    {
        auto x = graph::variable<T> (n, "x");
        auto y = graph::variable<T> (n, "y");
        auto z = graph::variable<T> (n, "z");
        auto x_last = graph::variable<T> (n, "x_last");
        auto y_last = graph::variable<T> (n, "y_last");
        auto z_last = graph::variable<T> (n, "z_last");
        auto kamp = graph::variable<T> (n, "kamp");
        auto pw = graph::variable<T> (n, static_cast<T> (1.0), "power");
        auto k_sum = graph::variable<T> (n, static_cast<T> (0.0), "k_sum");
        auto eq = make_eq(eqn);
        auto dlvec = graph::vector(eq->get_x(x, y, z) - eq->get_x(x_last, y_last, z_last),
                                   eq->get_y(x, y, z) - eq->get_y(x_last, y_last, z_last),
                                   eq->get_z(x, y, z) - eq->get_z(x_last, y_last, z_last));
        auto dl = dlvec->length();
        auto kdl = kamp*dl;
        auto k_next = kdl + k_sum;
        auto p_next = graph::exp(-2.0*k_sum);
        auto d_power = p_next - pw;
        d_power = graph::sqrt(d_power*d_power);
        workflow::manager<T> work(0);
        work.add_item({graph::variable_cast(x), graph::variable_cast(y), graph::variable_cast(z),
                       graph::variable_cast(x_last), graph::variable_cast(y_last), graph::variable_cast(z_last),
                       graph::variable_cast(kamp), graph::variable_cast(pw), graph::variable_cast(k_sum)},
                      {d_power},
                      {{x, graph::variable_cast(x_last)}, {y, graph::variable_cast(y_last)}, {z, graph::variable_cast(z_last)},
                       {p_next, graph::variable_cast(pw)}, {k_next, graph::variable_cast(k_sum)}},
                      graph::shared_random_state<T> (), "power", n);
        work.compile();
        graph::variable_cast(x_last)->set(in[2]);
        graph::variable_cast(y_last)->set(in[3]);
        graph::variable_cast(z_last)->set(in[4]);
        work.copy_to_device(x_last, graph::variable_cast(x_last)->data());
        work.copy_to_device(y_last, graph::variable_cast(y_last)->data());
        work.copy_to_device(z_last, graph::variable_cast(z_last)->data());
        for (size_t j = 1; j < nrec; j++) {
            graph::variable_cast(x)->set(in[8*j + 2]);
            graph::variable_cast(y)->set(in[8*j + 3]);
            graph::variable_cast(z)->set(in[8*j + 4]);
            graph::variable_cast(kamp)->set(kim[j]);
            work.copy_to_device(x, graph::variable_cast(x)->data());
            work.copy_to_device(y, graph::variable_cast(y)->data());
            work.copy_to_device(z, graph::variable_cast(z)->data());
            work.copy_to_device(kamp, graph::variable_cast(kamp)->data());
            work.run();
            work.wait();
            work.copy_to_host(pw, power[j].data());
            work.copy_to_host(d_power, d_power_rec[j].data());
        }
    }
    for (size_t j = 0; j < nrec; j++) {
        for (auto *v : {&kre[j], &kim[j], &power[j], &d_power_rec[j]}) {
            out.write(reinterpret_cast<const char *> (v->data()), sizeof(double)*n);
        }
    }
    return 0;
}

#endif

static int erfi_values(int argc, char **argv) {
    const size_t n = std::stoul(argv[2]);
    auto in = read_arrays(argv[3], 1, n);
    std::ofstream out(argv[4], std::ios::binary);
    std::vector<double> a(n), b(n);
    for (size_t i = 0; i < n; i++) {
        a[i] = special::w_im<double> (in[0][i]);
        b[i] = std::real(special::erfi<double> (std::complex<double> (in[0][i], 0.0)));
    }
    out.write(reinterpret_cast<const char *> (a.data()), sizeof(double)*n);
    out.write(reinterpret_cast<const char *> (b.data()), sizeof(double)*n);
    return 0;
}

#define DISPATCH_SOLVER(FN, DNAME, SNAME)                                                         \
    if (DNAME == "cold_plasma" && SNAME == "rk4") return FN<solver::rk4<dispersion::cold_plasma<T>>> (argc, argv);          \
    if (DNAME == "ordinary_wave" && SNAME == "rk4") return FN<solver::rk4<dispersion::ordinary_wave<T>>> (argc, argv);      \
    if (DNAME == "extra_ordinary_wave" && SNAME == "rk4") return FN<solver::rk4<dispersion::extra_ordinary_wave<T>>> (argc, argv); \
    if (DNAME == "bohm_gross" && SNAME == "rk4") return FN<solver::rk4<dispersion::bohm_gross<T>>> (argc, argv);            \
    if (DNAME == "simple" && SNAME == "rk4") return FN<solver::rk4<dispersion::simple<T>>> (argc, argv);                    \
    if (DNAME == "cold_plasma" && SNAME == "rk2") return FN<solver::rk2<dispersion::cold_plasma<T>>> (argc, argv);          \
    if (DNAME == "simple" && SNAME == "rk2") return FN<solver::rk2<dispersion::simple<T>>> (argc, argv);                    \
    if (DNAME == "bohm_gross" && SNAME == "split_simplextic") return FN<solver::split_simplextic<dispersion::bohm_gross<T>>> (argc, argv); \
    if (DNAME == "light_wave" && SNAME == "split_simplextic") return FN<solver::split_simplextic<dispersion::light_wave<T>>> (argc, argv); \
    if (DNAME == "light_wave" && SNAME == "rk4") return FN<solver::rk4<dispersion::light_wave<T>>> (argc, argv);

//  The cell a compiled reference kernel selects (piecewise.hpp:26-65 under -ffast-math).
static int cells(int argc, char **argv) {
    const size_t n = std::stoul(argv[2]);
    const double scale = std::stod(argv[3]), offset = std::stod(argv[4]);
    const size_t ncells = std::stoul(argv[5]);
    auto in = read_arrays(argv[6], 1, n);
    std::ofstream out(argv[7], std::ios::binary);
    auto x = graph::variable<T> (n, "x");
    graph::variable_cast(x)->set(in[0]);
    std::vector<T> ids(ncells);
    for (size_t i = 0; i < ncells; i++) ids[i] = static_cast<T> (i);
    auto cell = graph::piecewise_1D<T> (ids, x, scale, offset);
    workflow::manager<T> work(0);
    work.add_item({graph::variable_cast(x)}, {cell}, {}, graph::shared_random_state<T> (), "cells_kernel", n);
    work.compile();
    work.run();
    std::vector<double> buf(n);
    work.copy_to_host(cell, buf.data());
    out.write(reinterpret_cast<const char *> (buf.data()), sizeof(double)*n);
    return 0;
}

//  Minimal reproduction of the reducer rule at fault.
static int reducer_defect() {
    auto A = graph::variable<T> (1, "A"), B = graph::variable<T> (1, "B"), C = graph::variable<T> (1, "C"), W = graph::variable<T> (1, "W");
    const double a = 1.3, b = 0.7, c = 2.2, w = 500.0;
    graph::variable_cast(A)->set(a); graph::variable_cast(B)->set(b); graph::variable_cast(C)->set(c); graph::variable_cast(W)->set(w);
    auto two = graph::constant<T> (2.0), four = graph::constant<T> (4.0);
    auto f = (graph::pow(A*W, two)*B)/(graph::pow(C, two)*graph::pow(W, four));
    const double from_graph = f->evaluate().at(0);
    const double direct = (a*w)*(a*w)*b/(c*c*w*w*w*w);
    const double what_it_became = a*a*b/(c*c*c*c);

    auto x = graph::variable<T> (1, "x"), z = graph::variable<T> (1, "z"), Q = graph::variable<T> (1, "Q"), S = graph::variable<T> (1, "S");
    graph::variable_cast(x)->set(1.5); graph::variable_cast(z)->set(0.2); graph::variable_cast(Q)->set(-20.0); graph::variable_cast(S)->set(30.0);
    auto n = z*Q - x*S;
    auto h = (n*n)/((1.0 + z*z)*(W*W));
    const double symbolic = h->df(z)->evaluate().at(0);
    const double step = 1.0e-6;
    graph::variable_cast(z)->set(0.2 + step);
    const double hp = h->evaluate().at(0);
    graph::variable_cast(z)->set(0.2 - step);
    const double hm = h->evaluate().at(0);
    std::printf("{\"expression\": \"((A W)^2 B)/(C^2 W^4)\", \"A\": %.17g, \"B\": %.17g, \"C\": %.17g, \"W\": %.17g, "
                "\"reference_graph\": %.17g, \"direct\": %.17g, \"a2b_over_c4\": %.17g, "
                "\"derivative_expression\": \"d/dz (z Q - x S)^2/((1 + z^2) W^2) at x=1.5 z=0.2 Q=-20 S=30 W=500\", "
                "\"reference_df\": %.17g, \"central_difference\": %.17g}\n",
                a, b, c, w, from_graph, direct, what_it_became, symbolic, (hp - hm)/(2.0*step));
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { std::cerr << "usage: see header of oracle/ref_driver.cpp" << std::endl; return 2; }
    const std::string mode = argv[1];
    if (mode == "trace" && argc == 12) {
        const std::string d = argv[2], s = argv[4];
        if (s == "adaptive_rk4" && d == "cold_plasma") return trace_adaptive_impl<dispersion::cold_plasma<T>> (argc, argv);
        if (s == "adaptive_rk4" && d == "ordinary_wave") return trace_adaptive_impl<dispersion::ordinary_wave<T>> (argc, argv);
        DISPATCH_SOLVER(trace_impl, d, s)
    } else if (mode == "bench" && (argc == 9 || argc == 10)) {
        const std::string d = argv[2], s = "rk4";
        DISPATCH_SOLVER(bench_impl, d, s)
    } else if (mode == "rhs" && argc == 7) {
        const std::string d = argv[2];
        if (d == "cold_plasma") return rhs_impl<dispersion::cold_plasma<T>> (argc, argv);
        if (d == "ordinary_wave") return rhs_impl<dispersion::ordinary_wave<T>> (argc, argv);
        if (d == "extra_ordinary_wave") return rhs_impl<dispersion::extra_ordinary_wave<T>> (argc, argv);
        if (d == "bohm_gross") return rhs_impl<dispersion::bohm_gross<T>> (argc, argv);
        if (d == "simple") return rhs_impl<dispersion::simple<T>> (argc, argv);
    } else if (mode == "korc" && argc == 7) {
        return korc(argc, argv);
#ifndef GFB_REF_DRIVER_REAL_ONLY
    } else if (mode == "absorb" && argc == 7) {
        return absorb(argc, argv);
#endif
    } else if (mode == "erfi" && argc == 5) {
        return erfi_values(argc, argv);
    } else if (mode == "cells" && argc == 8) {
        return cells(argc, argv);
    } else if (mode == "reducer") {
        return reducer_defect();
    }
    std::cerr << "bad arguments; see header of oracle/ref_driver.cpp" << std::endl;
    return 2;
}
