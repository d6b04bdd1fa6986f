// Host-side checks of the cubic-spline node families (graph::spline_1d / spline_2d, node.hpp) that the EFIT
// equilibrium is built on.  No device needed: everything is leaf_node::evaluate().
//   1. the family construction and the reference's Horner chains of piecewise nodes (build_1D_spline /
//      build_psi, GFB_SPLINE_NODES=0) give the same psi, B, n_e and the same first AND second derivatives;
//   2. reverse-mode gradient() and forward df() agree through spline nodes;
//   3. members beyond the third derivative are zero, and derivatives of members are members.
// usage: spline_host <efit.gfbt>      prints one line per check, exit code = failures
#include <algorithm>
#include <cstdio>
#include <random>
#include "../../graph_framework_b200/csrc/graph/graph_framework.hpp"

using graph::leaf_ptr;
static int failures = 0;

//  Relative deviations with a floor of 1e-9 x the ensemble scale (tests/conftest.py rel_devs): {median, worst}.
struct deviation { double median, worst; };
static deviation compare(const leaf_ptr &a, const leaf_ptr &b) {
    auto va = a->evaluate(), vb = b->evaluate();
    double scale = 0.0;
    for (size_t i = 0; i < vb.size(); i++) scale = std::max(scale, std::abs(vb[i]));
    std::vector<double> d(vb.size());
    for (size_t i = 0; i < vb.size(); i++) {
        const double x = va.size() == 1 ? va[0] : va[i];
        d[i] = std::abs(x - vb[i])/std::max(std::abs(vb[i]), 1.0e-9*scale + 1.0e-300);
    }
    std::sort(d.begin(), d.end());
    return {d[d.size()/2], d.back()};
}
static double worst(const leaf_ptr &a, const leaf_ptr &b) { return compare(a, b).worst; }
static void expect(const char *what, const double dev, const double tolerance) {
    std::printf("%-58s %.3e (tolerance %.0e) %s\n", what, dev, tolerance, dev <= tolerance ? "ok" : "FAILED");
    if (!(dev <= tolerance)) failures++;
}
//  Two evaluations of the same EFIT quantity: psi ~ 0.3 is a sum of folded coefficients up to 4e7
//  (equilibrium.hpp:1121-1133), so ANY change of operation order (the chains' r-direction Horner is mul + add, the
//  family's is fma) moves a few points by up to ~1e-7 while the typical point agrees to rounding: the same two bounds
//  as tests/conftest.py assert_rhs_close.
static void expect_same_field(const char *what, const leaf_ptr &a, const leaf_ptr &b, const double median_tolerance) {
    const deviation d = compare(a, b);
    const bool ok = d.median <= median_tolerance && d.worst <= 1.0e-6;
    std::printf("%-58s median %.3e (tolerance %.0e), worst %.3e (1e-06) %s\n", what, d.median, median_tolerance, d.worst, ok ? "ok" : "FAILED");
    if (!ok) failures++;
}

int main(int argc, char **argv) {
    const std::string file = argc > 1 ? argv[1] : "tests/golden/efit.gfbt";
    const size_t n = 4000;
    std::mt19937_64 engine(7);
    std::uniform_real_distribution<double> ux(0.7, 2.6), uy(-0.4, 0.4), uz(-1.7, 1.7);       // includes points outside the grid (clamped cells)
    std::vector<double> xs(n), ys(n), zs(n);
    for (size_t i = 0; i < n; i++) { xs[i] = ux(engine); ys[i] = uy(engine); zs[i] = uz(engine); }
    auto x = graph::variable(xs, "x"), y = graph::variable(ys, "y"), z = graph::variable(zs, "z");

    struct fields { leaf_ptr ne, te, bx, by, bz; };
    auto build = [&] (const bool nodes) {
        equilibrium::spline_nodes() = nodes;
        auto eq = equilibrium::make_efit<> (file);
        auto b = eq->get_magnetic_field(x, y, z);
        return fields {eq->get_electron_density(x, y, z), eq->get_electron_temperature(x, y, z), b->get_x(), b->get_y(), b->get_z()};
    };
    const fields f = build(true), c = build(false);

//  1. values, first and second derivatives: families against chains.  The chains evaluate the same cubics in the
//     same order for values and first derivatives; second derivatives differ in operation order, and psi ~ 0.3 is a
//     sum of folded coefficients up to 4e7, hence the looser bound there.
    expect_same_field("n_e: family vs Horner chain", f.ne, c.ne, 1.0e-12);
    expect_same_field("T_e: family vs Horner chain", f.te, c.te, 1.0e-12);
    expect_same_field("B_x: family vs Horner chain", f.bx, c.bx, 1.0e-12);
    expect_same_field("B_y: family vs Horner chain", f.by, c.by, 1.0e-12);
    expect_same_field("B_z: family vs Horner chain", f.bz, c.bz, 1.0e-12);
    expect_same_field("dn_e/dx", f.ne->df(x), c.ne->df(x), 1.0e-11);
    expect_same_field("dn_e/dz", f.ne->df(z), c.ne->df(z), 1.0e-11);
    expect_same_field("dB_x/dz (second derivatives of psi)", f.bx->df(z), c.bx->df(z), 1.0e-11);
    expect_same_field("dB_z/dx", f.bz->df(x), c.bz->df(x), 1.0e-11);
    expect_same_field("dB_y/dy", f.by->df(y), c.by->df(y), 1.0e-11);

//  2. reverse mode through spline nodes = forward mode.
    auto h = f.ne*f.bx*f.bx + f.bz*graph::sqrt(f.by*f.by + 1.0) + f.te;
    auto g = graph::gradient(h, {x, y, z});
    expect("gradient(h)[x] vs h->df(x)", worst(g[0], h->df(x)), 1.0e-11);
    expect("gradient(h)[y] vs h->df(y)", worst(g[1], h->df(y)), 1.0e-11);
    expect("gradient(h)[z] vs h->df(z)", worst(g[2], h->df(z)), 1.0e-11);

//  3. closure of the families.
    {
        equilibrium::spline_nodes() = true;
        auto r = graph::variable(std::vector<double> {1.3, 1.9}, "r"), zz = graph::variable(std::vector<double> {0.1, -0.4}, "zz");
        std::array<std::array<leaf_ptr, 4>, 4> k;
        std::vector<double> table(6*5);
        for (size_t i = 0; i < 4; i++) for (size_t j = 0; j < 4; j++) {
            for (size_t t = 0; t < table.size(); t++) table[t] = std::sin(1.0 + i + 3.0*j + 0.37*t);
            k[i][j] = graph::piecewise_2D(table, 5, r, 0.3, 0.9, zz, 0.5, -1.0);
        }
        auto s = graph::spline_2d(k, r, zz);
        bool closed = s->df(r)->op == graph::op_t::spline_2d && s->df(zz)->df(zz)->op == graph::op_t::spline_2d &&
                      s->df(r)->df(zz) == s->df(zz)->df(r) &&
                      s->df(r)->df(r)->df(r)->df(r)->is_constant(0.0) && graph::spline_member(s.get(), 0, 4)->is_constant(0.0);
        expect("spline_2d: closed under df(), mixed derivatives commute, order 4 = 0", closed ? 0.0 : 1.0, 0.0);
//     value against the plain double sum, derivative against a central difference
        const double rr = 1.3, zv = 0.1, u = (rr - 0.9)/0.3;
        const size_t cell = static_cast<size_t> (u)*5 + static_cast<size_t> ((zv + 1.0)/0.5);
        double sum = 0.0;
        for (size_t i = 0; i < 4; i++) for (size_t j = 0; j < 4; j++) sum += k[i][j]->table->values[cell]*std::pow(u, i)*std::pow(zv, j);
        expect("spline_2d value vs the double sum", std::abs(s->evaluate()[0] - sum)/std::abs(sum), 1.0e-14);
        const double step = 1.0e-6;
        auto at = [&] (const double a, const double b) { r->set(std::vector<double> {a, 1.9}); zz->set(std::vector<double> {b, -0.4}); return s->evaluate()[0]; };
        const double fd_r = (at(rr + step, zv) - at(rr - step, zv))/(2.0*step), fd_z = (at(rr, zv + step) - at(rr, zv - step))/(2.0*step);
        at(rr, zv);
        expect("spline_2d d/dr vs central difference", std::abs(s->df(r)->evaluate()[0] - fd_r)/std::abs(fd_r), 1.0e-8);
        expect("spline_2d d/dz vs central difference", std::abs(s->df(zz)->evaluate()[0] - fd_z)/std::abs(fd_z), 1.0e-8);
    }
    std::printf("%d failure(s)\n", failures);
    return failures;
}
