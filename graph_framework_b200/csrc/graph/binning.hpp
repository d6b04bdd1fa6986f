//------------------------------------------------------------------------------
//  binning.hpp -- keep the rays of a workflow sorted by table cell while stepping.
//
//  No reference counterpart: the reference never reorders rays.  Rays (and particles) are
//  independent, so their order in the SoA arrays is free; when the rays of a warp sit in the same
//  cell of the coefficient tables the gathers of the kernel become broadcasts (EFIT +6 %, Boris
//  push 2x, VMEC +50 %).  workflow::ray_order owns the policy:
//    * before a block of steps the arrays are sorted by cell if they are in the caller's order;
//    * every `check_every` steps the disorder (share of neighbouring rays in different cells) is
//      measured and the arrays are re-sorted when it exceeds one cell boundary per warp;
//    * anything that reads or writes rays by index first restores the caller's order, and copies
//      to the host un-permute on the way (gfb_copy_rays_d2h, gfb_snapshot_async).
//  The device work is done by gfb_bin_rays / gfb_bin_rays_rz / gfb_unbin_rays (include/gfb200.h).
//------------------------------------------------------------------------------
#ifndef gfb_graph_binning_hpp
#define gfb_graph_binning_hpp

#include "equilibrium.hpp"

namespace workflow {
    template<typename T=double, bool SAFE_MATH=false>
    class ray_order {
    private:
        manager<T, SAFE_MATH> *work = nullptr;
        equilibrium::cell_grid grid;
        std::vector<graph::leaf_ptr> sort_vars;         ///< 1 array (1-D grid) or x, y, z ((R, Z) grid)
        std::vector<graph::leaf_ptr> arrays;            ///< per-ray arrays that travel with their ray
        std::vector<graph::leaf_ptr> launch_outputs;    ///< written by every launch: restored, never sorted
        size_t num_rays = 0;
        size_t check_every = 0, since_check = 0;
        double threshold = 1.0/32.0;

        static uint64_t key(const graph::leaf_ptr &n) { return reinterpret_cast<uint64_t> (n.get()); }
        gfb_ctx *device() { return work->get_context().device(); }
        std::vector<uint64_t> keys(const bool with_outputs) const {
            std::vector<uint64_t> k;
            for (auto &a : arrays) k.push_back(key(a));
            if (with_outputs) for (auto &o : launch_outputs) k.push_back(key(o));
            return k;
        }
        int sort() {
            auto k = keys(false);
            int rc;
            if (grid.dims == 1) {
                rc = gfb_bin_rays(device(), key(sort_vars[0]), grid.lo[0], grid.hi[0], grid.cells[0],
                                  k.data(), static_cast<int> (k.size()), num_rays);
            } else {
                const uint64_t xyz[3] = {key(sort_vars[0]), key(sort_vars[1]), key(sort_vars[2])};
                rc = gfb_bin_rays_rz(device(), xyz, grid.lo, grid.hi, grid.cells, k.data(), static_cast<int> (k.size()), num_rays);
            }
            since_check = 0;
            return rc;
        }

    public:
        bool active() const { return work && grid.dims && num_rays > 1; }

///  `arrays` must contain the sort arrays.  check_every = 0: sort once, never look again.
        void configure(manager<T, SAFE_MATH> &manager, const equilibrium::cell_grid &g,
                       std::vector<graph::leaf_ptr> sort_by, std::vector<graph::leaf_ptr> move,
                       std::vector<graph::leaf_ptr> outputs, const size_t n, const size_t period) {
            restore();
            work = &manager;
            grid = g;
            sort_vars = sort_by;
            arrays = move;
            launch_outputs = outputs;
            num_rays = n;
            check_every = period;
            since_check = 0;
            assert((grid.dims == 0 || sort_vars.size() == (grid.dims == 1 ? 1u : 3u)) && "Sort arrays do not match the grid.");
        }
        void disable() {
            restore();
            grid.dims = 0;
        }
        void add_arrays(const std::vector<graph::leaf_ptr> &more) {
            restore();
            arrays.insert(arrays.end(), more.begin(), more.end());
        }
        void set_check_every(const size_t period) { check_every = period; }
        size_t get_check_every() const { return check_every; }

///  Call before launching steps.  Returns how many steps may run before the next look at the order.
        size_t prepare(const size_t wanted) {
            if (!active()) return wanted;
            if (!gfb_is_binned(device())) {
                jit::check(sort(), "ray binning");
            } else if (check_every && since_check >= check_every) {
//  A coherent beam keeps its order for thousands of steps, rays with random directions lose it
//  within tens: measure, and sort only when more than one cell boundary per warp has appeared.
                const uint64_t xyz[3] = {key(sort_vars[0]), grid.dims == 2 ? key(sort_vars[1]) : 0, grid.dims == 2 ? key(sort_vars[2]) : 0};
                double disorder = 1.0;
                jit::check(gfb_bin_disorder(device(), xyz, grid.dims == 1 ? 1 : 3, grid.lo, grid.hi, grid.cells, num_rays, &disorder),
                           "ray order check");
                if (disorder >= threshold) jit::check(sort(), "ray binning");
                else since_check = 0;
            }
            if (!check_every) return wanted;
            return std::min(wanted, check_every - std::min(since_check, check_every - 1));
        }
        void advanced(const size_t steps) { since_check += steps; }

///  Back to the caller's order (no-op when the rays are not sorted at the moment).
        void restore() {
            if (!work || !grid.dims) return;
            if (!gfb_is_binned(device())) return;
            auto k = keys(true);
            jit::check(gfb_unbin_rays(device(), k.data(), static_cast<int> (k.size()), num_rays), "ray unbinning");
        }
///  Per-ray array to the host in the caller's order; the device order is left alone.
        void copy_to_host(graph::leaf_ptr node, T *destination) {
            if (!work) return;
            const bool per_ray = std::find(arrays.begin(), arrays.end(), node) != arrays.end() ||
                                 std::find(launch_outputs.begin(), launch_outputs.end(), node) != launch_outputs.end();
            if (active() && per_ray) {
                jit::check(gfb_copy_rays_d2h(device(), key(node), destination, num_rays), "copy rays to host");
            } else {
                work->copy_to_host(node, destination);
            }
        }
    };
}

#endif /* gfb_graph_binning_hpp */
