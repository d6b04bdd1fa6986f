//------------------------------------------------------------------------------
//  kernels.cu -- statically compiled sm_100a helper kernels of libgfb200.
//
//    * max reduction (replaces the generated single-block `max_reduction` of
//      /root/reference/graph_framework/cuda_context.hpp:954-995),
//    * power-deposition histogram (algorithm of /root/reference/utilities/bin.py:53-106),
//    * FP64 FMA peak probe (roofline denominator),
//    * L2 flush.
//------------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <cstdint>

namespace {
//  Order preserving map double -> uint64 so atomicMax works on any sign.
__device__ __forceinline__ unsigned long long ordered(const double d) {
    const unsigned long long u = static_cast<unsigned long long> (__double_as_longlong(d));
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) max_kernel(const double *__restrict__ in, const unsigned long long n,
                                                  unsigned long long *__restrict__ result) {
//  NaN elements are ignored, like the reference's max() (CUDA max = fmax, cuda_context.hpp:973-985):
//  one bad ray must not end a convergence loop for all the others (workflow.hpp:179-205).  Key 0 =
//  "no number seen" (all NaN, or n == 0) and decodes to NaN.
    double m = __longlong_as_double(0xfff0000000000000ll);        // -inf
    bool seen = false;
    for (unsigned long long i = static_cast<unsigned long long> (blockIdx.x)*blockDim.x + threadIdx.x; i < n;
         i += static_cast<unsigned long long> (gridDim.x)*blockDim.x) {
        const double v = __ldg(in + i);
        if (v == v) {
            m = fmax(m, v);
            seen = true;
        }
    }
    unsigned long long key = seen ? ordered(m) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    __shared__ unsigned long long warp_max[8];
    if ((threadIdx.x & 31) == 0) {
        warp_max[threadIdx.x >> 5] = key;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        key = warp_max[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_down_sync(0xffu, key, o);
            key = other > key ? other : key;
        }
        if (threadIdx.x == 0) {
            atomicMax(result, key);
        }
    }
}

//  A traced beam is narrow: most lanes of a warp land in a handful of bins and, away from the
//  resonance, carry a weight of exactly zero.  Zero weights are skipped (adding +0 changes nothing)
//  and lanes with the same bin are summed in the warp first (match.any), so one atomic is issued per
//  distinct bin per warp instead of one per ray.
__global__ void __launch_bounds__(256) deposit_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                      const double *__restrict__ z, const double *__restrict__ w,
                                                      const unsigned long long n, double *__restrict__ hist,
                                                      const double x0, const double y0, const double z0,
                                                      const double ix, const double iy, const double iz,
                                                      const int nx, const int ny, const int nz) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long stride = static_cast<unsigned long long> (gridDim.x)*blockDim.x;
    const unsigned long long rounded = (n + 31ull) & ~31ull;          // whole warps stay in the loop together
    for (unsigned long long i = static_cast<unsigned long long> (blockIdx.x)*blockDim.x + threadIdx.x; i < rounded;
         i += stride) {
        bool active = false;
        unsigned long long bin = 0;
        double weight = 0.0;
        if (i < n) {
            weight = __ldg(w + i);
            const double fx = floor((__ldg(x + i) - x0)*ix);
            const double fy = floor((__ldg(y + i) - y0)*iy);
            const double fz = floor((__ldg(z + i) - z0)*iz);
            active = weight != 0.0 && fx >= 0.0 && fx < nx && fy >= 0.0 && fy < ny && fz >= 0.0 && fz < nz;
            if (active) {
                bin = (static_cast<unsigned long long> (fx)*ny + static_cast<unsigned long long> (fy))*nz +
                      static_cast<unsigned long long> (fz);
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const unsigned peers = __match_any_sync(mask, bin);
            const unsigned leader = __ffs(peers) - 1u;
            double sum = 0.0;
            for (unsigned rest = peers; rest; rest &= rest - 1u) {
                sum += __shfl_sync(peers, weight, __ffs(rest) - 1);
            }
            if (lane == leader) {
                atomicAdd(hist + bin, sum);
            }
        }
    }
}

//------------------------------------------------------------------------------
//  Binning of rays by table cell (counting sort).  Rays are independent, so their order in the SoA
//  arrays is free; putting rays of the same cell next to each other lets a warp share the
//  coefficient rows it gathers (VMEC: 86 modes x 3 quantities per cell).  The order inside a cell is
//  whatever the atomics give -- per-ray results do not depend on the slot a ray occupies.
//------------------------------------------------------------------------------
//  Cell of ray i: 1-D grid on array a, or the (R, Z) grid of an axisymmetric table on arrays
//  (a, b, c) = (x, y, z) with R = sqrt(x^2 + y^2).  Same clamp as the table look-ups; NaN lands in cell 0.
struct cell_grid {
    const double *a, *b, *c;
    double lo0, inv0, lo1, inv1;
    unsigned n0, n1;                // n1 = 0: one dimensional
    __device__ __forceinline__ unsigned operator()(const unsigned i) const {
        const double u = n1 ? sqrt(__ldg(a + i)*__ldg(a + i) + __ldg(b + i)*__ldg(b + i)) : __ldg(a + i);
        const unsigned i0 = static_cast<unsigned> (fmin(fmax((u - lo0)*inv0, 0.0), static_cast<double> (n0 - 1u)));
        if (!n1) return i0;
        const unsigned i1 = static_cast<unsigned> (fmin(fmax((__ldg(c + i) - lo1)*inv1, 0.0), static_cast<double> (n1 - 1u)));
        return i0*n1 + i1;
    }
};
__global__ void __launch_bounds__(256) bin_count_kernel(const cell_grid grid, const unsigned n,
                                                        unsigned *__restrict__ cell_of, unsigned *__restrict__ count) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned rounded = (n + 31u) & ~31u;
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < rounded; i += gridDim.x*blockDim.x) {
        const bool active = i < n;
        unsigned cell = 0;
        if (active) {
            cell = grid(i);
            cell_of[i] = cell;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const unsigned peers = __match_any_sync(mask, cell);
            if (lane == __ffs(peers) - 1u) atomicAdd(count + cell, static_cast<unsigned> (__popc(peers)));
        }
    }
}
//  How far the order has decayed: the number of neighbouring slots whose rays sit in different cells
//  (occupied cells - 1 right after a sort, ~n for a random order).
__global__ void __launch_bounds__(256) bin_breaks_kernel(const cell_grid grid, const unsigned n, unsigned *__restrict__ breaks) {
    unsigned mine = 0;
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x + 1u; i < n; i += gridDim.x*blockDim.x) {
        mine += grid(i) != grid(i - 1u) ? 1u : 0u;
    }
    for (int offset = 16; offset > 0; offset >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, offset);
    if ((threadIdx.x & 31u) == 0u && mine) atomicAdd(breaks, mine);
}

//  Exclusive scan of the cell counts by one block (cells <= 2^24); also primes the placement cursors.
__global__ void __launch_bounds__(1024) bin_scan_kernel(const unsigned *__restrict__ count, unsigned *__restrict__ cursor,
                                                        const unsigned cells) {
    __shared__ unsigned partial[1024];
    const unsigned per = (cells + blockDim.x - 1u)/blockDim.x;
    const unsigned begin = min(threadIdx.x*per, cells), end = min(begin + per, cells);
    unsigned sum = 0;
    for (unsigned c = begin; c < end; c++) sum += count[c];
    partial[threadIdx.x] = sum;
    __syncthreads();
    for (unsigned stride = 1; stride < blockDim.x; stride <<= 1) {      // Hillis-Steele inclusive scan
        const unsigned v = threadIdx.x >= stride ? partial[threadIdx.x - stride] : 0u;
        __syncthreads();
        partial[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned offset = threadIdx.x ? partial[threadIdx.x - 1u] : 0u;
    for (unsigned c = begin; c < end; c++) {
        cursor[c] = offset;
        offset += count[c];
    }
}
//  Lanes of a warp that go to the same cell reserve their slots with ONE atomic (a narrow beam puts
//  most of a warp in one cell) and keep their relative order.
__global__ void __launch_bounds__(256) bin_place_kernel(const unsigned *__restrict__ cell_of, const unsigned n,
                                                        unsigned *__restrict__ cursor, unsigned *__restrict__ perm) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned rounded = (n + 31u) & ~31u;
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < rounded; i += gridDim.x*blockDim.x) {
        const bool active = i < n;
        const unsigned cell = active ? cell_of[i] : 0u;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const unsigned peers = __match_any_sync(mask, cell);
            const unsigned leader = __ffs(peers) - 1u;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(cursor + cell, static_cast<unsigned> (__popc(peers)));
            base = __shfl_sync(peers, base, leader);
            perm[base + __popc(peers & ((1u << lane) - 1u))] = i;
        }
    }
}
__global__ void __launch_bounds__(256) gather_kernel(double *__restrict__ dst, const double *__restrict__ src,
                                                     const unsigned *__restrict__ perm, const unsigned n) {
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) dst[i] = __ldg(src + perm[i]);
}
__global__ void __launch_bounds__(256) scatter_kernel(double *__restrict__ dst, const double *__restrict__ src,
                                                      const unsigned *__restrict__ perm, const unsigned n) {
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) dst[perm[i]] = __ldg(src + i);
}

//  Composition of two permutations: total[i] = first[second[i]].
__global__ void __launch_bounds__(256) compose_kernel(unsigned *__restrict__ total, const unsigned *__restrict__ first,
                                                      const unsigned *__restrict__ second, const unsigned n) {
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) total[i] = first[second[i]];
}

//  8 independent FMA chains per thread keep the FP64 pipe saturated.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, const int iters, const double a, const double b) {
    double r0 = threadIdx.x, r1 = r0 + 1.0, r2 = r0 + 2.0, r3 = r0 + 3.0;
    double r4 = r0 + 4.0, r5 = r0 + 5.0, r6 = r0 + 6.0, r7 = r0 + 7.0;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    const double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 12345.678) {
        out[0] = s;
    }
}

//  One slice of an all-reduce over peer memory: dst[i] = sum over devices of src[d][i], always in
//  device order, so every device ends up with bit-identical sums.  src[d] may live on another GPU of
//  the NVSwitch domain (peer access enabled): the loads travel over NVLink, 16 bytes per lane.
struct peer_sources {
    const double *ptr[16];
};
__global__ void __launch_bounds__(256) sum_peers_kernel(double *__restrict__ dst, const peer_sources src, const int num,
                                                        const size_t count) {
    const size_t stride = static_cast<size_t> (gridDim.x)*blockDim.x;
    const size_t pairs = count/2;
    for (size_t i = static_cast<size_t> (blockIdx.x)*blockDim.x + threadIdx.x; i < pairs; i += stride) {
        double2 acc = reinterpret_cast<const double2 *> (src.ptr[0])[i];
        for (int d = 1; d < num; d++) {
            const double2 v = reinterpret_cast<const double2 *> (src.ptr[d])[i];
            acc.x += v.x;
            acc.y += v.y;
        }
        reinterpret_cast<double2 *> (dst)[i] = acc;
    }
    if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = src.ptr[0][count - 1];
        for (int d = 1; d < num; d++) acc += src.ptr[d][count - 1];
        dst[count - 1] = acc;
    }
}

//  The same stream with all three FMA operands in (non-uniform) registers, the shape real ray code has:
//  8 chains r_i = fma(r_i, s_j, t_k) whose multipliers and addends rotate through 8 + 8 per-thread values.
//  Measures what the register file can feed the FP64 pipe, which the uniform-operand probe above does not.
__global__ void __launch_bounds__(256) fp64_peak_regs_kernel(double *out, const int iters, const double a, const double b) {
    double r[8], s[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        r[j] = threadIdx.x + j;
        s[j] = a + 1.0e-9*(threadIdx.x + 3*j);
        t[j] = b*(1.0 + 1.0e-3*(threadIdx.x + 5*j));
    }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = fma(r[j], s[(j + k) & 7], t[(j + 3*k) & 7]);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) sum += r[j];
    if (sum == 12345.678) {
        out[0] = sum;
    }
}

__global__ void __launch_bounds__(256) fill_kernel(double *p, const size_t n, const double v) {
    for (size_t i = static_cast<size_t> (blockIdx.x)*blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t> (gridDim.x)*blockDim.x) {
        p[i] = v;
    }
}
}  // namespace

extern "C" {
int gfb_k_max(const double *in, unsigned long long n, unsigned long long *result, int sms, cudaStream_t s) {
    const unsigned long long blocks_needed = (n + 255)/256;
    const unsigned grid = static_cast<unsigned> (blocks_needed < static_cast<unsigned long long> (sms)*8 ? blocks_needed : sms*8);
    cudaMemsetAsync(result, 0, sizeof(unsigned long long), s);
    max_kernel<<<grid ? grid : 1, 256, 0, s>>> (in, n, result);
    return static_cast<int> (cudaGetLastError());
}
double gfb_k_unorder(unsigned long long key) {
    if (key == 0ull) {                      // nothing but NaN (or nothing at all)
        const unsigned long long nan_bits = 0x7ff8000000000000ull;
        double d;
        __builtin_memcpy(&d, &nan_bits, 8);
        return d;
    }
    const unsigned long long u = (key & 0x8000000000000000ull) ? (key & 0x7fffffffffffffffull) : ~key;
    double d;
    __builtin_memcpy(&d, &u, 8);
    return d;
}
int gfb_k_deposit(const double *x, const double *y, const double *z, const double *w, unsigned long long n,
                  double *hist, const double *lo, const double *hi, const int *bins, int sms, cudaStream_t s) {
    const unsigned long long blocks_needed = (n + 255)/256;
    const unsigned grid = static_cast<unsigned> (blocks_needed < static_cast<unsigned long long> (sms)*8 ? blocks_needed : sms*8);
    deposit_kernel<<<grid ? grid : 1, 256, 0, s>>> (x, y, z, w, n, hist, lo[0], lo[1], lo[2],
                                                     bins[0]/(hi[0] - lo[0]), bins[1]/(hi[1] - lo[1]), bins[2]/(hi[2] - lo[2]),
                                                     bins[0], bins[1], bins[2]);
    return static_cast<int> (cudaGetLastError());
}
//  perm[slot] = ray that moves into `slot`.  work: cells*2 + n unsigned (count, cursor, cell_of).
//  values[1], values[2] non-null select the (R, Z) grid; cells = cells0*max(cells1, 1).
int gfb_k_bin_permutation(const double *const values[3], unsigned n, const double lo[2], const double hi[2],
                          const unsigned cells01[2], unsigned *work, unsigned *perm, int sms, cudaStream_t s) {
    const unsigned cells = cells01[0]*(cells01[1] ? cells01[1] : 1u);
    unsigned *count = work, *cursor = work + cells, *cell_of = work + 2*cells;
    const unsigned grid = static_cast<unsigned> ((n + 255u)/256u < static_cast<unsigned> (sms)*8u ? (n + 255u)/256u : sms*8);
    cell_grid g;
    g.a = values[0]; g.b = values[1]; g.c = values[2];
    g.lo0 = lo[0]; g.inv0 = cells01[0]/(hi[0] - lo[0]); g.n0 = cells01[0];
    g.lo1 = lo[1]; g.inv1 = cells01[1] ? cells01[1]/(hi[1] - lo[1]) : 0.0; g.n1 = cells01[1];
    cudaMemsetAsync(count, 0, sizeof(unsigned)*cells, s);
    bin_count_kernel<<<grid ? grid : 1, 256, 0, s>>> (g, n, cell_of, count);
    bin_scan_kernel<<<1, 1024, 0, s>>> (count, cursor, cells);
    bin_place_kernel<<<grid ? grid : 1, 256, 0, s>>> (cell_of, n, cursor, perm);
    return static_cast<int> (cudaGetLastError());
}
int gfb_k_bin_breaks(const double *const values[3], unsigned n, const double lo[2], const double hi[2],
                     const unsigned cells01[2], unsigned *breaks, int sms, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned> ((n + 255u)/256u < static_cast<unsigned> (sms)*8u ? (n + 255u)/256u : sms*8);
    cell_grid g;
    g.a = values[0]; g.b = values[1]; g.c = values[2];
    g.lo0 = lo[0]; g.inv0 = cells01[0]/(hi[0] - lo[0]); g.n0 = cells01[0];
    g.lo1 = lo[1]; g.inv1 = cells01[1] ? cells01[1]/(hi[1] - lo[1]) : 0.0; g.n1 = cells01[1];
    cudaMemsetAsync(breaks, 0, sizeof(unsigned), s);
    bin_breaks_kernel<<<grid ? grid : 1, 256, 0, s>>> (g, n, breaks);
    return static_cast<int> (cudaGetLastError());
}
int gfb_k_permute(double *dst, const double *src, const unsigned *perm, unsigned n, int scatter, int sms, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned> ((n + 255u)/256u < static_cast<unsigned> (sms)*16u ? (n + 255u)/256u : sms*16);
    if (scatter) scatter_kernel<<<grid ? grid : 1, 256, 0, s>>> (dst, src, perm, n);
    else gather_kernel<<<grid ? grid : 1, 256, 0, s>>> (dst, src, perm, n);
    return static_cast<int> (cudaGetLastError());
}
int gfb_k_compose(unsigned *total, const unsigned *first, const unsigned *second, unsigned n, int sms, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned> ((n + 255u)/256u < static_cast<unsigned> (sms)*16u ? (n + 255u)/256u : sms*16);
    compose_kernel<<<grid ? grid : 1, 256, 0, s>>> (total, first, second, n);
    return static_cast<int> (cudaGetLastError());
}
int gfb_k_fp64_peak(double *scratch, int iters, int sms, cudaStream_t s) {
    fp64_peak_kernel<<<sms*8, 256, 0, s>>> (scratch, iters, 0.999999, 1.0e-6);
    return static_cast<int> (cudaGetLastError());
}
int gfb_k_fp64_peak_regs(double *scratch, int iters, int sms, cudaStream_t s) {
    fp64_peak_regs_kernel<<<sms*4, 256, 0, s>>> (scratch, iters, 0.999999, 1.0e-6);
    return static_cast<int> (cudaGetLastError());
}
//  src: `num` (<= 16) pointers to `count` doubles each, 16-byte aligned; dst likewise.
int gfb_k_sum_peers(double *dst, const double *const *src, int num, size_t count, int sms, cudaStream_t s) {
    if (num < 1 || num > 16) return 1;
    peer_sources p;
    for (int d = 0; d < 16; d++) p.ptr[d] = d < num ? src[d] : nullptr;
    const size_t blocks_needed = (count/2 + 255)/256;
    const unsigned grid = static_cast<unsigned> (blocks_needed < static_cast<size_t> (sms)*4 ? blocks_needed : static_cast<size_t> (sms)*4);
    sum_peers_kernel<<<grid ? grid : 1, 256, 0, s>>> (dst, p, num, count);
    return static_cast<int> (cudaGetLastError());
}
int gfb_k_fill(double *p, size_t n, double v, int sms, cudaStream_t s) {
    fill_kernel<<<sms*8, 256, 0, s>>> (p, n, v);
    return static_cast<int> (cudaGetLastError());
}
}
