// Known-answer tests of the ray path written against THIS repository's C++ front end
// (graph_framework_b200/csrc/graph), GPU required.  Scenarios and expected values are the ones the
// reference pins in graph_tests/physics_test.cpp (:28-68 k.x - wt invariant, :112-169 Bohm-Gross
// parabola, :209-260 light-wave parabola, :286-333 acoustic speed, :341-378 O-mode cut-off,
// :380-470 cold-plasma cut-offs, :472-530 reflection, :583-618 EFIT reflection),
// graph_tests/solver_test.cpp:28-60 (D^2 stays below tolerance for five steps) and
// graph_tests/dispersion_test.cpp:25-64 (Newton solve for every wavenumber component and w).
// Tolerances are the reference's CUDA-branch values.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "../../graph_framework_b200/csrc/graph/graph_framework.hpp"

#ifndef EFIT_FILE
#define EFIT_FILE "tests/golden/efit.gfbt"
#endif

using graph::leaf_ptr;
static int failures = 0;
#define EXPECT(cond, what) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, what); failures++; } else { std::printf("ok   %s\n", what); } } while (0)

struct ray {
    leaf_ptr w, kx, ky, kz, x, y, z, t;
    explicit ray(const size_t n=1) :
    w(graph::variable(n, "\\omega")), kx(graph::variable(n, "k_{x}")), ky(graph::variable(n, "k_{y}")),
    kz(graph::variable(n, "k_{z}")), x(graph::variable(n, "x")), y(graph::variable(n, "y")),
    z(graph::variable(n, "z")), t(graph::variable(n, "t")) {}
    void set(const double w0, const double kx0, const double ky0, const double kz0,
             const double x0, const double y0, const double z0) {
        w->set(w0); kx->set(kx0); ky->set(ky0); kz->set(kz0); x->set(x0); y->set(y0); z->set(z0); t->set(0.0);
    }
};
template<class SOLVER>
SOLVER make(ray &r, const double dt, equilibrium::shared<> &eq) {
    return SOLVER(r.w, r.kx, r.ky, r.kz, r.x, r.y, r.z, r.t, graph::constant(dt), eq);
}

namespace constants {
    const double q = 1.602176634E-19, me = 9.1093837015E-31, mi = 3.34449469E-27;
    const double mu0 = M_PI*4.0E-7, epsilon0 = 8.8541878138E-12;
    const double c = 1.0/std::sqrt(mu0*epsilon0);
}

// k.x - w t is conserved for D = n^2 - 1 (any equilibrium).
static void invariant() {
    ray r;
    r.set(0.37, 0.81, 0.0, 0.0, 0.42, 0.93, 0.18);
    auto eq = equilibrium::make_slab<> ();
    auto solve = make<solver::rk2<dispersion::simple<>>> (r, 1.0, eq);
    solve.init(r.kx);
    solve.compile();
    auto invariant = r.kx*r.x + r.ky*r.y + r.kz*r.z - r.w*r.t;
    const double before = invariant->evaluate().at(0);
    for (int i = 0; i < 10; i++) { solve.step(); solve.sync_host(); }
    EXPECT(std::abs(before - invariant->evaluate().at(0)) < 5.0E-15, "k.x - wt preserved over 10 rk2 steps");
}

// Linear density ramp, no field: x(t) is a parabola for Bohm-Gross and light waves.
template<class SOLVER>
static void parabola(const char *what, const double tolerance, const double k_guess, const bool thermal) {
    using namespace constants;
    ray r;
    r.set(600.0, k_guess, 0.0, 0.0, -1.0, 0.0, 0.0);
    auto eq = equilibrium::make_no_magnetic_field<> ();
    auto solve = make<SOLVER> (r, 0.1, eq);
    solve.init(r.kx, tolerance);
    solve.compile();
    for (int i = 0; i < 20; i++) { solve.step(); solve.sync_host(); }
    const double w0 = 600.0, ne0 = 1.0E19;
    const double wp2 = ne0*0.9*q*q/(epsilon0*me*c*c), wp2_slope = ne0*0.1*q*q/(epsilon0*me*c*c);
    const double time = r.t->evaluate().at(0);
    double expected;
    if (thermal) {
        const double vth2 = 2.0*q*1000.0/(me*c*c);
        const double k0 = std::sqrt(2.0/3.0*(w0*w0 - wp2)/vth2);
        expected = -3.0/8.0*vth2*wp2_slope/(w0*w0)*time*time + 3.0/2.0*vth2/w0*k0*time - 1.0;
    } else {
        const double k0 = std::sqrt(w0*w0 - wp2);
        expected = -wp2_slope/(4.0*w0*w0)*time*time + k0/w0*time - 1.0;
    }
    const double d = r.x->evaluate().at(0) - expected;
    EXPECT(d*d < tolerance, what);
}

static void acoustic(const double tolerance) {
    using namespace constants;
    ray r;
    r.set(1.0, 600.0, 0.0, 0.0, 0.0, 0.0, 0.0);
    auto eq = equilibrium::make_no_magnetic_field<> ();
    auto solve = make<solver::rk4<dispersion::acoustic_wave<>>> (r, 0.0001, eq);
    solve.init(r.kx, tolerance);
    solve.compile();
    for (int i = 0; i < 20; i++) { solve.step(); solve.sync_host(); }
    const double vs = std::sqrt((q*1000.0 + 3.0*q*1000.0)/mi)/c;
    const double d = r.x->evaluate().at(0)/r.t->evaluate().at(0) - vs;
    EXPECT(d*d < tolerance, "acoustic wave travels at the sound speed");
}

static void o_mode_cutoff() {
    using namespace constants;
    ray r;
    r.set(1000.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
    auto eq = equilibrium::make_slab_density<> ();
    auto solve = make<solver::rk4<dispersion::ordinary_wave<>>> (r, 0.0001, eq);
    solve.init(r.x);
    const double wp2 = 1.0E19*q*q/(epsilon0*me*c*c);
    const double x_cut = (1000.0*1000.0 - 1.0 - wp2)/(wp2*0.1);
    const double d = r.x->evaluate().at(0) - x_cut;
    EXPECT(d*d < 8.0E-10, "Newton in x finds the O-mode cut-off");
}

static void cold_plasma_cutoffs() {
    ray r(2);
    r.set(1100.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
    auto eq = equilibrium::make_slab_density<> ();
    auto solve = make<solver::rk4<dispersion::cold_plasma<>>> (r, 0.1, eq);
    r.x->set(0, 25.0);
    r.x->set(1, 5.0);
    solve.init(r.x);
    double plasma_cut = r.x->evaluate().at(0);
    const double right_cut = r.x->evaluate().at(1);
    r.x->set(0.0);
    r.kx->set(0, 1000.0);       // O-mode
    r.kx->set(1, 500.0);        // X-mode
    solve.init(r.kx);
    solve.compile();
    while (std::abs(r.t->evaluate().at(0)) < 30.0) { solve.step(); solve.sync_host(); }
    auto pos = r.x->evaluate();
    EXPECT(pos.at(0) > right_cut && pos.at(0) < plasma_cut, "O-mode passes the right cut-off, stops before the plasma cut-off");
    EXPECT(pos.at(1) < right_cut, "X-mode is reflected at the right cut-off");

    r.w->set(800.0);
    r.x->set(0, 25.0);
    r.x->set(1, 5.0);
    r.kx->set(0.0);
    r.t->set(0.0);
    solve.init(r.x, 5.0E-30);
    solve.sync_device();
    plasma_cut = r.x->evaluate().at(1);
    r.x->set(0.0);
    r.kx->set(0, 500.0);        // O-mode
    r.kx->set(1, 1500.0);       // X-mode
    solve.init(r.kx);
    solve.sync_device();
    while (std::abs(r.t->evaluate().at(0)) < 60.0) { solve.step(); solve.sync_host(); }
    pos = r.x->evaluate();
    EXPECT(pos.at(0) < plasma_cut, "trapped case: O-mode stays below the plasma cut-off");
    EXPECT(pos.at(1) > plasma_cut, "trapped case: X-mode crosses the plasma cut-off");
}

static void reflection(const double tolerance) {
    using namespace constants;
    const double wce = -q/(me*c);
    ray r;
    r.set(wce, 0.0, 0.0, 0.7*wce, 0.1, 0.0, 0.0);
    auto eq = equilibrium::make_slab<> ();
    auto solve = make<solver::rk4<dispersion::cold_plasma<>>> (r, 0.0001, eq);
    solve.init(r.x, tolerance);
    const double cutoff = r.x->evaluate().at(0);
    r.x->set(cutoff - 0.00001*cutoff);
    r.kx->set(22.0);
    solve.init(r.kx, tolerance);
    solve.compile();
    solve.sync_host();
    double furthest = r.x->evaluate().at(0), now = furthest;
    bool inside = true;
    size_t steps = 0;
    do {
        solve.step();
        solve.sync_host();
        now = r.x->evaluate().at(0);
        furthest = std::max(now, furthest);
        inside = inside && std::abs(furthest - cutoff) < 1.9E-6;
    } while (furthest == now && ++steps < 200000);
    EXPECT(inside && steps < 200000, "ray turns around within 1.9e-6 of the cut-off");
}

static void efit_reflects() {
    ray r;
    r.set(590.0, -600.0, 0.0, 0.0, 2.5, 0.0, 0.0);
    auto eq = equilibrium::make_efit<> (EFIT_FILE);
    auto solve = make<solver::rk4<dispersion::ordinary_wave<>>> (r, 0.0001, eq);
    solve.init(r.kx);
    solve.compile();
    for (int i = 0; i < 10000; i++) solve.step();          // one fused launch sequence, no host round trips
    solve.sync_host();
    EXPECT(r.kx->evaluate().at(0) > 0.0, "O-mode ray launched inward at 590 reflects in the EFIT equilibrium (kx changes sign)");
}

// adaptive_rk4 (solver.hpp:881-1006) has no reference test and is only selectable from the xrays
// command line.  Its step-length rule (Newton on 1/dt + lambda D_next^2 for dt AND lambda) is
// reproduced as written; checked here: the two work items alternate on the device, every ray gets
// its own dt, time advances and the state stays finite.
static void adaptive() {
    ray r(4);
    r.set(900.0, 1000.0, 0.25, 0.15, 0.0, 0.0, 0.0);
    auto eq = equilibrium::make_gaussian_density<> ();
    auto dt = graph::variable(4, 0.5/10000.0, "dt");
    solver::adaptive_rk4<dispersion::cold_plasma<>> solve(r.w, r.kx, r.ky, r.kz, r.x, r.y, r.z, r.t, dt, eq);
    solve.init(r.kx, 1.0E-30);
    solve.compile();
    for (int i = 0; i < 3; i++) solve.step();
    solve.sync_host();
    const double t_end = r.t->evaluate().at(0);
    std::printf("     adaptive_rk4: t after 3 steps = %g, x = %g\n", t_end, r.x->evaluate().at(0));
    EXPECT(t_end != 0.0 && std::isfinite(t_end) && std::isfinite(r.x->evaluate().at(3)), "adaptive_rk4 runs: time changes by the solved dt (the rule does not constrain its sign), state finite");
}

template<class SOLVER>
static void keeps_dispersion(const char *what, const double tolerance, const double w0, const double kx0, const double dt) {
    ray r;
    r.set(w0, kx0, 0.25, 0.15, 0.0, 0.0, 0.0);
    auto eq = equilibrium::make_gaussian_density<> ();
    auto solve = make<SOLVER> (r, dt, eq);
    solve.init(r.kx, tolerance);
    solve.compile();
    bool ok = true;
    for (int i = 0; i < 5; i++) {
        solve.step();
        ok = ok && std::abs(solve.check_residual(0)) < tolerance;
    }
    EXPECT(ok, what);
}

template<class DISPERSION>
static void solves_every_unknown(const char *what, const double tolerance, const double w0, const double guess,
                                 equilibrium::shared<> eq) {
    ray r;
    r.set(w0, 0.25, 0.25, 0.15, 0.0, 0.0, 0.0);
    dispersion::dispersion_interface<DISPERSION> D(r.w, r.kx, r.ky, r.kz, r.x, r.y, r.z, r.t, eq);
    graph::input_nodes<> inputs = {r.w, r.x, r.y, r.z, r.kx, r.ky, r.kz, r.t};
    auto residual = [&] () {
        auto d = D.get_d()->evaluate().at(0);
        return d*d;
    };
    bool ok = true;
    r.kx->set(guess);
    D.solve(r.kx, inputs, 0, tolerance);
    ok = ok && residual() < 1.0E-24;
    r.kx->set(0.2);
    D.solve(r.ky, inputs, 0, tolerance);
    ok = ok && residual() < 1.0E-24;
    r.ky->set(0.25);
    r.kz->set(guess);
    D.solve(r.kz, inputs, 0, tolerance);
    ok = ok && residual() < 1.0E-24;
    r.kz->set(0.15);
    r.kx->set(guess);
    D.solve(r.w, inputs, 0, tolerance);
    ok = ok && residual() < 1.0E-24;
    EXPECT(ok, what);
}

// absorption::weak_damping with the reference's calling sequence (own device context, sync_device,
// run, sync_host; absorption.hpp:466-483): the kernel against the host evaluation of the same
// expression (jit_test.cpp style), on rays that straddle the electron cyclotron resonance of the
// EFIT case, plus the analytic limits far from it (no damping, Re k_amp = |k|).
static void absorption_standalone() {
    const size_t n = 96;
    ray r(n);
    auto kamp_re = graph::variable(n, "kamp_re"), kamp_im = graph::variable(n, "kamp_im");
    std::vector<double> xs(n), ws(n), kys(n);
    for (size_t i = 0; i < n; i++) {
        xs[i] = 1.7 + 0.7*static_cast<double> (i)/static_cast<double> (n - 1);
        ws[i] = 690.0 + static_cast<double> (i%7)*3.0;
        kys[i] = -100.0 + static_cast<double> (i%5)*4.0;
    }
    r.w->set(ws); r.kx->set(-600.0); r.ky->set(kys); r.kz->set(7.0);
    r.x->set(xs); r.y->set(0.02); r.z->set(0.03); r.t->set(0.0);
    auto eq = equilibrium::make_efit<> (EFIT_FILE);
    absorption::weak_damping<> damping(kamp_re, kamp_im, r.w, r.kx, r.ky, r.kz, r.x, r.y, r.z, r.t, eq);
    damping.compile();
    damping.sync_device();
    damping.run();
    damping.sync_host();
    const auto im_host = damping.get_imaginary_expression()->evaluate();
    const auto re_host = damping.get_real_expression()->evaluate();
    double worst_im = 0.0, worst_re = 0.0, largest = 0.0;
    size_t undamped = 0;
    for (size_t i = 0; i < n; i++) {
        const double im = kamp_im->data()[i], re = kamp_re->data()[i];
        largest = std::max(largest, std::abs(im));
        worst_im = std::max(worst_im, std::abs(im - im_host.at(i))/std::max(std::abs(im_host.at(i)), 1.0E-6));
        if (std::abs(im_host.at(i)) > 1.0E-3) worst_re = std::max(worst_re, std::abs(re - re_host.at(i))/std::abs(re_host.at(i)));
        const double klen = std::sqrt(600.0*600.0 + kys[i]*kys[i] + 49.0);
        if (im == 0.0 && std::abs(re - klen) < 1.0E-9*klen) undamped++;
    }
    EXPECT(worst_im < 1.0E-8, "weak_damping kernel == host evaluation (Im k_amp)");
    EXPECT(worst_re < 1.0E-10, "weak_damping kernel == host evaluation (Re k_amp)");
    EXPECT(largest > 1.0, "rays at the resonance are damped");
    EXPECT(undamped > 5, "far from the resonance: Im k_amp = 0 and Re k_amp = |k|");
}

int main() {
    const double tolerance = 1.6E-21;       // physics_test.cpp:652, the reference's CUDA branch
    invariant();
    parabola<solver::rk4<dispersion::bohm_gross<>>> ("rk4: Bohm-Gross ray follows the analytic parabola", tolerance, 1000.0, true);
    parabola<solver::split_simplextic<dispersion::bohm_gross<>>> ("split_simplextic: Bohm-Gross parabola", tolerance, 1000.0, true);
    parabola<solver::rk4<dispersion::light_wave<>>> ("rk4: light wave follows the analytic parabola", tolerance, 100.0, false);
    parabola<solver::split_simplextic<dispersion::light_wave<>>> ("split_simplextic: light-wave parabola", tolerance, 100.0, false);
    acoustic(tolerance);
    o_mode_cutoff();
    reflection(tolerance);
    cold_plasma_cutoffs();
    efit_reflects();
    adaptive();
    absorption_standalone();
    keeps_dispersion<solver::rk2<dispersion::simple<>>> ("rk2 simple keeps D^2 < 1e-30", 1.0E-30, 0.5, 0.25, 1.0);
    keeps_dispersion<solver::rk4<dispersion::simple<>>> ("rk4 simple keeps D^2 < 1e-30", 1.0E-30, 0.5, 0.25, 1.0);
    keeps_dispersion<solver::rk2<dispersion::gaussian_well<>>> ("rk2 gaussian_well keeps D^2 < 1e-30", 1.0E-30, 0.5, 0.25, 0.00001);
    keeps_dispersion<solver::rk4<dispersion::gaussian_well<>>> ("rk4 gaussian_well keeps D^2 < 1e-30", 1.0E-30, 0.5, 0.25, 0.00001);
    keeps_dispersion<solver::rk2<dispersion::cold_plasma<>>> ("rk2 cold_plasma keeps D^2 < 1e-30", 1.0E-30, 900.0, 1000.0, 0.5/10000.0);
    keeps_dispersion<solver::rk4<dispersion::cold_plasma<>>> ("rk4 cold_plasma keeps D^2 < 1e-30", 1.0E-30, 900.0, 1000.0, 0.5/10000.0);
    solves_every_unknown<dispersion::simple<>> ("Newton: simple, every unknown", 1.0E-30, 0.5, 1.0, equilibrium::make_gaussian_density<> ());
    solves_every_unknown<dispersion::acoustic_wave<>> ("Newton: acoustic_wave, every unknown", 1.0E-30, 1.0, 600.0, equilibrium::make_no_magnetic_field<> ());
    solves_every_unknown<dispersion::gaussian_well<>> ("Newton: gaussian_well, every unknown", 1.0E-30, 0.5, 1.0, equilibrium::make_gaussian_density<> ());
    solves_every_unknown<dispersion::cold_plasma<>> ("Newton: cold_plasma, every unknown", 1.0E-30, 900.0, 1000.0, equilibrium::make_gaussian_density<> ());
    std::printf("%d failure(s)\n", failures);
    return failures ? 1 : 0;
}
