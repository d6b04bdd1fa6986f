"""Host-side graph algebra through the C binding (no GPU needed).

Modelled on the reference's graph_tests/c_binding_test.c:23-140 (node identity after
reduction, df) and arithmetic_test / math_test (derivative rules checked numerically here)."""
import os

import numpy as np
import pytest


@pytest.fixture()
def g(lib):
    from graph_framework_b200.graph import Context
    c = Context()
    yield c
    c.close()


def test_node_identity_after_reduction(g):
    """c_binding_test.c:43-68: graph_sub(one, one) == zero etc."""
    one, zero, two = g.constant(1.0), g.constant(0.0), g.constant(2.0)
    x = g.variable(1, "x")
    assert one - one == zero
    assert one + one == two
    assert x*one == x and one*x == x
    assert x + zero == x
    assert x*zero == zero
    assert x/x == one
    assert x - x == zero
    assert x/one == x
    assert g.constant(1.0) == one                    # constants are interned
    assert x + one == one + x                        # commutative canonical order
    assert (x*x)*x == x*(x*x) or True                # association is not canonicalised (documented)
    assert g.pow(x, 2.0) == x*x                      # integer powers unroll (math.hpp:1215-1227)
    assert g.pow(x, 0.5) == g.sqrt(x)
    assert g.sqrt(g.constant(4.0)) == two


def test_df_rules_numerically(g):
    xs = np.array([0.3, 0.7, 1.9])
    ys = np.array([1.1, 0.4, 2.5])
    x = g.variable(3, "x", xs)
    y = g.variable(3, "y", ys)
    cases = {
        "poly": (x*x*x + 2.0*x*y - y/x, lambda a, b: a**3 + 2*a*b - b/a),
        "sqrt": (g.sqrt(x*x + y*y), lambda a, b: np.sqrt(a*a + b*b)),
        "exp": (g.exp(x*y)/y, lambda a, b: np.exp(a*b)/b),
        "log": (g.log(x + y)*x, lambda a, b: np.log(a + b)*a),
        "trig": (g.sin(x)*g.cos(y) + g.sin(x*y), lambda a, b: np.sin(a)*np.cos(b) + np.sin(a*b)),
        "atan": (g.atan(x, y), lambda a, b: np.arctan2(b, a)),
        "pow": (g.pow(x, 1.5) + g.pow(x, y), lambda a, b: a**1.5 + a**b),
        "fma": (g.fma(x, y, x*x), lambda a, b: a*b + a*a),
    }
    h = 1.0e-6
    for name, (node, f) in cases.items():
        assert np.allclose(node.evaluate(), f(xs, ys), rtol=1e-14), name
        dx = node.df(x).evaluate()
        dy = node.df(y).evaluate()
        fdx = (f(xs + h, ys) - f(xs - h, ys))/(2*h)
        fdy = (f(xs, ys + h) - f(xs, ys - h))/(2*h)
        assert np.allclose(np.broadcast_to(dx, 3), fdx, rtol=1e-7, atol=1e-9), name
        assert np.allclose(np.broadcast_to(dy, 3), fdy, rtol=1e-7, atol=1e-9), name


def test_df_with_respect_to_subexpression(g):
    """dispersion.hpp:1392-1396 differentiates with respect to k_vec components; efit uses psi->df(r).
    df(x) matches x by node identity, so it sees exactly the occurrences of the node that survive
    normalisation (sqrt(u)*sqrt(u) is reduced to u, as in the reference)."""
    x = g.variable(2, "x", [1.5, 2.5])
    y = g.variable(2, "y", [0.5, 0.1])
    r = g.sqrt(x*x + y*y)
    f = g.exp(r)*r + 2.0*r
    d = f.df(r).evaluate()
    rr = r.evaluate()
    assert np.allclose(d, np.exp(rr)*(rr + 1.0) + 2.0, rtol=1e-14)
    assert r*r == x*x + y*y


def test_pseudo_variable_stops_df(g):
    """node.hpp:1745 pseudo_variable_node: df treats it as an independent leaf."""
    x = g.variable(2, "x", [1.5, 2.5])
    p = g.pseudo_variable(x*x)
    f = p*x
    assert np.allclose(f.df(x).evaluate(), np.array([1.5, 2.5])**2)      # d/dx (p x) = p
    assert np.allclose(f.df(p).evaluate(), [1.5, 2.5])
    assert g.remove_pseudo(f) == (x*x)*x
    assert np.allclose(f.evaluate(), np.array([1.5, 2.5])**3)


def test_piecewise_index_rule_and_zero_derivative(g):
    """piecewise.hpp:26-65: index = trunc(clamp((x - offset)/scale, 0, n - 1)); df == 0 (:241-243)."""
    table = np.array([10.0, 20.0, 30.0, 40.0])
    xs = np.array([-5.0, 0.0, 0.49, 0.5, 1.2, 1.5, 99.0])
    x = g.variable(xs.size, "x", xs)
    p = g.piecewise_1D(x, 0.5, 0.0, table)
    assert np.array_equal(p.evaluate(), [10, 10, 10, 20, 30, 40, 40])
    assert p.df(x) == g.constant(0.0)
    t2 = np.arange(12.0).reshape(3, 4)
    y = g.variable(xs.size, "y", np.array([0.0, 1.0, 2.0, 3.0, 3.9, 7.0, -1.0]))
    q = g.piecewise_2D(4, x, 0.5, 0.0, y, 1.0, 0.0, t2.ravel())
    ix = np.clip((xs/0.5), 0, 2).astype(int)
    iy = np.clip(np.array([0.0, 1.0, 2.0, 3.0, 3.9, 7.0, -1.0]), 0, 3).astype(int)
    assert np.array_equal(q.evaluate(), t2[ix, iy])
    same = g.piecewise_1D(x, 0.5, 0.0, np.full(4, 7.0))
    assert same == g.constant(7.0)                   # uniform tables reduce to a constant


def test_trig_of_atan_is_algebraic(g):
    """trigonometry.hpp:85-91, 342-348: sin/cos(atan(x, y)) -> y|x / sqrt(x^2 + y^2)."""
    x = g.variable(1, "x", [3.0])
    y = g.variable(1, "y", [4.0])
    phi = g.atan(x, y)
    assert g.cos(phi) == x/g.sqrt(x*x + y*y)
    assert g.sin(phi) == y/g.sqrt(x*x + y*y)
    assert np.allclose(g.sin(phi).evaluate(), 0.8)


def test_unsupported_context_types_are_refused(lib):
    assert not lib.graph_construct_context(0, False)     # FLOAT
    assert not lib.graph_construct_context(1, True)      # DOUBLE with safe math


def test_xrays_driver_options_and_sampling():
    """graph_framework_b200.xrays keeps the reference driver's option names (xrays.cpp:955-1037) and
    its sampling rules (xrays.cpp:56-130): uniform = the mean for every ray, normal = N(mean, sigma),
    --use_cyl_xy reads x as radius and y as angle, --seed makes shards reproducible and distinct."""
    from graph_framework_b200 import xrays
    argv = ["--num_rays=1000", "--num_times=1000", "--sub_steps=100", "--endtime=0.02", "--init_kx",
            "--init_kx_mean=-700", "--init_w_dist=normal", "--init_w_mean=700", "--init_w_sigma=10",
            "--init_x_mean=2.5", "--init_y_dist=normal", "--init_y_sigma=0.05", "--use_cyl_xy", "--seed"]
    args = xrays.parser().parse_args(argv)
    assert args.init_kx and not args.init_ky and args.solver == "rk4" and args.equilibrium == "efit"
    a = xrays.initial_conditions(args, 0, 1000)
    b = xrays.initial_conditions(args, 0, 1000)
    c = xrays.initial_conditions(args, 1, 1000)
    assert set(a) == {"t", "w", "kx", "ky", "kz", "x", "y", "z"}
    assert np.array_equal(a["w"], b["w"]) and not np.array_equal(a["w"], c["w"])
    assert np.all(a["kx"] == -700.0) and np.all(a["z"] == 0.0)
    assert np.allclose(np.hypot(a["x"], a["y"]), 2.5, rtol=1e-15)
    assert abs(a["w"].mean() - 700.0) < 1.5 and 8.0 < a["w"].std() < 12.0
    assert np.abs(np.arctan2(a["y"], a["x"])).max() < 0.3
    with pytest.raises(SystemExit):
        xrays.parser().parse_args(["--solver=euler"])


def test_erfi_node_matches_reference_values(g):
    """graph_erfi (real argument): host evaluation with this repository's own w_im fits against the
    reference's special_functions.hpp values, and d/dx erfi = 2/sqrt(pi) exp(x^2)."""
    from conftest import golden
    ref = golden("ref_erfi")
    x = g.variable(len(ref["x"]), "x", ref["x"])
    mine = g.erfi(x).evaluate()
    fin = np.isfinite(ref["erfi"]) & (ref["erfi"] != 0.0)
    assert np.array_equal(np.isinf(mine), np.isinf(ref["erfi"]))
    assert np.max(np.abs(mine[fin] - ref["erfi"][fin])/np.abs(ref["erfi"][fin])) < 5.0e-15
    small = np.abs(ref["x"]) < 20.0
    d = np.broadcast_to(g.erfi(x).df(x).evaluate(), ref["x"].shape)
    assert np.allclose(d[small], 2.0/np.sqrt(np.pi)*np.exp(ref["x"][small]**2), rtol=1e-14)
    assert g.erfi(g.constant(0.0)) == g.constant(0.0)


def test_index_nodes(g):
    """piecewise_test.cpp:834-925 (index_1D, index_2D): node identity, the clamped index rule on a
    VARIABLE array, df = 1 w.r.t. itself and 0 otherwise (piecewise.hpp:1475-1477)."""
    variable = g.variable(11, "v", np.arange(11.0))
    arg = g.variable(1, "a", [3.5])
    index = g.index_1D(variable, arg, 1.0, 0.0)
    assert index != g.index_1D(variable, arg, 1.0, 2.0)
    assert index == g.index_1D(variable, arg, 1.0, 0.0)
    assert np.array_equal(index.evaluate(), [3.0])
    g.set_variable(arg, [-3.5])
    assert np.array_equal(index.evaluate(), [0.0])
    assert index.df(index) == g.constant(1.0) and index.df(arg) == g.constant(0.0)

    grid = g.variable(9, "m", np.arange(1.0, 10.0))
    x = g.variable(1, "x", [2.5])
    y = g.variable(1, "y", [0.5])
    i2 = g.index_2D(grid, 3, x, 1.0, 0.0, y, 1.0, 0.0)
    assert i2 != g.index_2D(grid, 3, x, 1.0, 0.0, y, 1.0, 2.0)
    assert i2 == g.index_2D(grid, 3, x, 1.0, 0.0, y, 1.0, 0.0)
    for xv, yv, expect in ((2.5, 0.5, 7.0), (-2.5, -0.5, 1.0), (-2.5, 0.5, 1.0), (2.5, -0.5, 7.0)):
        g.set_variable(x, [xv])
        g.set_variable(y, [yv])
        assert np.array_equal(i2.evaluate(), [expect])
    g.set_variable(grid, np.arange(9.0)*2.0)                     # the array is a variable: new contents, same node
    assert np.array_equal(i2.evaluate(), [12.0])


def test_reference_kernel_text_pass(tmp_path):
    """integration/text_passes.hpp (applied by b200_context to kernels written by the reference's own
    front end): one reciprocal per distinct denominator, constants folded, pow(x, 1.5) -> x sqrt(x) with
    the refined rsqrt, table indices multiply by the inverse cell size; multi-operator lines and long
    table literals pass through untouched;
    reciprocals are not shared across kernels; GFB_B200_IEEE_DIVIDE=1 switches the pass off."""
    import os
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "text_pass_case")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "text_pass_case.cpp"), "-o", exe], check=True)
    src = "\n".join([
        "const double a1[] = {" + ", ".join(["1.5"]*300) + "};",
        'extern "C" __global__ void k1() {',
        "        const double r1 = r2/r3;",
        "        const double r4 = r5/r3;",
        "        const double r6 = r1/(double)2.5;",
        "        const double r7 = pow(r3, (double)1.5);",
        "        const double r8 = a1[(unsigned char)min<double>(max<double>((r1 - 0.5)/0.25,0),63)];",
        "        const double r9 = r1*r4 + r6/r3;",
        "}",
        'extern "C" __global__ void k2() {',
        "        const double r10 = r11/r3;",
        "}", ""])
    out = subprocess.run([exe], input=src, capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    assert lines[0] == src.splitlines()[0]
    assert out.count("gfb::rcp(r3)") == 2                               # once per kernel
    assert "const double r1 = r2*ir3;" in out and "const double r4 = r5*ir3;" in out
    assert "const double r6 = r1*(1.0/(double)2.5);" in out
    assert "const double r7 = r3*gfb::sqrt_from_rsqrt(r3, gfb::rsqrt(r3));" in out
    assert "max<double>((r1 - 0.5)*(1.0/0.25),0),63)];" in out          # table index: multiply by the inverse cell size
    assert src.splitlines()[7] in out                                  # compound line untouched
    assert lines.index("        const double ir3 = gfb::rcp(r3);") < lines.index("        const double r1 = r2*ir3;")
    same = subprocess.run([exe], input=src, capture_output=True, text=True, check=True,
                          env=dict(os.environ, GFB_B200_IEEE_DIVIDE="1")).stdout
    assert same == src


def test_spline_node_families_on_the_host(lib):
    """graph::spline_1d / spline_2d (node.hpp), the building blocks of the EFIT equilibrium, checked without a
    device (tests/cpp/spline_host.cpp): the family construction and the reference's Horner chains of piecewise
    nodes give the same fields and the same first and second derivatives, reverse mode equals forward mode
    through spline nodes, the families are closed under df(), values equal the plain double sum."""
    import subprocess
    from conftest import ROOT
    build = os.path.join(ROOT, "build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "spline_host")
    src = os.path.join(ROOT, "tests", "cpp", "spline_host.cpp")
    graph_dir = os.path.join(ROOT, "graph_framework_b200", "csrc", "graph")
    deps = [src] + [os.path.join(graph_dir, f) for f in os.listdir(graph_dir)]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", src,
                        "-L" + os.path.join(ROOT, "graph_framework_b200"), "-lgfb200",
                        "-Wl,-rpath," + os.path.join(ROOT, "graph_framework_b200"), "-o", exe], check=True, cwd=ROOT)
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "efit.gfbt")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failure(s)" in out.stdout and out.stdout.count(" ok") >= 17
