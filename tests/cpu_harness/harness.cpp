// Runs one emitted kernel on the CPU, one "thread" per ray.  TEST INFRASTRUCTURE.
//   harness <tables.bin> <in.bin> <out.bin> <n> <steps> <num_ptr_inputs> <num_outputs> [scalar0]
// GFB_KERNEL_FILE is the emitted source, GFB_KERNEL_NAME the kernel to call.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>
#include "cuda_stub.hpp"
#include "../../graph_framework_b200/csrc/skeleton.cuh"
#include "../../graph_framework_b200/csrc/special.cuh"
namespace gfb { __attribute__((aligned(128))) unsigned char smem[256*1024]; }
#ifndef GFB_UNROLL_STAGES
#define GFB_UNROLL_STAGES 0
#endif
#include GFB_KERNEL_FILE

int main(int argc, char **argv) {
    if (argc < 8) return 2;
    const size_t n = std::atol(argv[4]);
    const unsigned steps = std::atoi(argv[5]);
    const int ni = std::atoi(argv[6]), no = std::atoi(argv[7]);
    std::vector<std::vector<double>> bufs(ni + no, std::vector<double> (n, 0.0));
    {
        std::ifstream f(argv[2], std::ios::binary);
        for (int i = 0; i < ni; i++) f.read(reinterpret_cast<char *> (bufs[i].data()), 8*n);
    }
    std::vector<std::vector<double>> tables;
    {
        std::ifstream f(argv[1], std::ios::binary);
        unsigned long long cnt = 0;
        while (f.read(reinterpret_cast<char *> (&cnt), 8)) {
            tables.emplace_back(cnt);
            f.read(reinterpret_cast<char *> (tables.back().data()), 8*cnt);
        }
    }
    gfb_args a = {};
    int p = 0;
    for (auto &b : bufs) a.ptr[p++] = b.data();
    for (auto &t : tables) a.ptr[p++] = t.data();
    a.n = n;
    a.steps = steps;
    a.scalar[0] = argc > 8 ? std::atof(argv[8]) : 1.0e-30;
    blockDim.x = 1;
    for (size_t i = 0; i < n; i++) {
        blockIdx.x = static_cast<unsigned> (i);
        threadIdx.x = 0;
        GFB_KERNEL_NAME(a);
    }
    std::ofstream o(argv[3], std::ios::binary);
    for (auto &b : bufs) o.write(reinterpret_cast<const char *> (b.data()), 8*n);
    return 0;
}
