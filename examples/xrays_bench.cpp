// xrays_bench on the B200 back end: the reference benchmark's workload
// (/root/reference/graph_benchmark/xrays_bench.cpp:34-104: cold-plasma RK4 in the EFIT
// equilibrium, every ray started at x = 2.5 with kx = -600, w = 500, Newton solve for kx, 1000
// steps, one host thread per device, rays split batch/extra) written against this repository's
// C++ front end.  Usage: xrays_bench_b200 [rays=100000] [steps=1000] [efit.gfbt]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../graph_framework_b200/csrc/graph/graph_framework.hpp"

int main(int argc, char **argv) {
    const size_t num_rays = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 100000;
    const size_t num_steps = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 1000;
    const std::string efit = argc > 3 ? argv[3] : "tests/golden/efit.gfbt";
    const size_t devices = std::max<size_t> (1, std::min(jit::context<>::max_concurrency(), num_rays));
    const size_t batch = num_rays/devices, extra = num_rays%devices;
    std::vector<double> setup(devices), init(devices), compile(devices), stepping(devices);
    std::vector<std::thread> threads;
    for (size_t d = 0; d < devices; d++) {
        threads.emplace_back([&, d] () {
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto seconds = [] (auto a, auto b) { return std::chrono::duration<double> (b - a).count(); };
            const size_t n = batch + (extra > d ? 1 : 0);
            const auto t0 = now();
            auto w = graph::variable(n, 500.0, "\\omega");
            auto kx = graph::variable(n, -600.0, "k_{x}"), ky = graph::variable(n, 0.0, "k_{y}"), kz = graph::variable(n, 0.0, "k_{z}");
            auto x = graph::variable(n, 2.5, "x"), y = graph::variable(n, 0.0, "y"), z = graph::variable(n, 0.0, "z");
            auto t = graph::variable(n, 0.0, "t");
            auto eq = equilibrium::make_efit<> (efit);
            auto dt = graph::constant(1.0/static_cast<double> (num_steps));
            solver::rk4<dispersion::cold_plasma<>> solve(w, kx, ky, kz, x, y, z, t, dt, eq, "", n, d);
            const auto t1 = now();
            solve.init(kx);
            const auto t2 = now();
            solve.compile();
            const auto t3 = now();
            for (size_t s = 0; s < num_steps; s++) solve.step();
            solve.sync_host();
            const auto t4 = now();
            setup[d] = seconds(t0, t1); init[d] = seconds(t1, t2); compile[d] = seconds(t2, t3); stepping[d] = seconds(t3, t4);
            if (d == 0) std::printf("ray 0: kx(0) solved, x = %.12f  kx = %.9f  t = %.6f\n", x->evaluate().at(0), kx->evaluate().at(0), t->evaluate().at(0));
        });
    }
    for (auto &th : threads) th.join();
    auto worst = [] (const std::vector<double> &v) { double m = 0; for (double e : v) m = std::max(m, e); return m; };
    std::printf("{\"program\": \"xrays_bench (B200 back end)\", \"devices\": %zu, \"rays\": %zu, \"steps\": %zu, \"setup_s\": %.4f, "
                "\"init_s\": %.4f, \"compile_s\": %.4f, \"steps_s\": %.6f, \"ray_steps_per_s\": %.6e}\n",
                devices, num_rays, num_steps, worst(setup), worst(init), worst(compile), worst(stepping),
                static_cast<double> (num_rays)*static_cast<double> (num_steps)/worst(stepping));
    return 0;
}
