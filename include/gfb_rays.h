/*------------------------------------------------------------------------------
 *  gfb_rays.h -- C ABI of the ray tracing hot path.
 *
 *  The reference exposes this path only as C++ templates
 *  (solver::rk4<dispersion::cold_plasma<T>> etc.) driven by
 *  /root/reference/graph_benchmark/xrays_bench.cpp:34-104 and
 *  /root/reference/graph_driver/xrays.cpp:413-529.  This header is the flat C
 *  view of exactly that call sequence so that Python (ctypes), C or Fortran can
 *  drive it:   create -> set_state -> init (Newton) -> compile -> step ... -> get_state.
 *  Every function returns 0 on success; gfb_last_error() (gfb200.h) has the text.
 *----------------------------------------------------------------------------*/
#ifndef GFB_RAYS_H
#define GFB_RAYS_H

#include <stddef.h>
#include <stdint.h>
#include "gfb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfb_rays gfb_rays;

/* State array order everywhere: t, w, x, y, z, kx, ky, kz  (solver.hpp:304-314). */
enum { GFB_T = 0, GFB_W, GFB_X, GFB_Y, GFB_Z, GFB_KX, GFB_KY, GFB_KZ, GFB_NUM_STATE };

/* dispersion : "cold_plasma" | "ordinary_wave" | "extra_ordinary_wave" | "bohm_gross" | "simple" |
 *              "light_wave" | "acoustic_wave" | "gaussian_well" | "ion_cyclotron" | "stiff"
 * equilibrium: "efit" (table_file = GFBT path) | "slab" | "slab_density" | "slab_field" |
 *              "no_magnetic_field" | "gaussian_density"
 * solver     : "rk4" | "rk2" (staged skeleton) | "rk4_graph" | "rk2_graph" (stages unrolled in the
 *              graph, the reference's construction) | "split_simplextic" (separable dispersions only)
 * options    : NULL or space separated key=value list:
 *              block=<threads> minblocks=<n> stage_tables=<0|1> share_rcp=<0|1> fast_div=<0|1> unroll_stages=<0|1>
 *              fused_steps=<max steps per launch> mode_unroll=<n> mode_recurrence=<0|1> (Fourier mode loops)
 *              absorption=<0|1>  also build the weak-damping and power kernels (gfb_rays_trace_absorb)
 *              bin_rays=<0|steps> rays of tabulated equilibria (efit, vmec) are kept sorted by table cell
 *                                 while stepping (default on; the order is checked after about half a cell of
 *                                 travel, 0.5*cell/dt steps within [200 (efit) | 20 (vmec), 5000], and the rays are re-sorted when
 *                                 more than 1/32 of neighbouring rays sit in different cells); 0 = off
 * Mirrors the constructor sequence of xrays_bench.cpp:53-85. */
gfb_rays *gfb_rays_create(const char *dispersion, const char *equilibrium, const char *table_file,
                          const char *solver, size_t num_rays, double dt, int device, const char *options);
void gfb_rays_destroy(gfb_rays *r);

/* variable->set(...) for the eight state arrays (host pointers, num_rays doubles each). */
int gfb_rays_set_state(gfb_rays *r, const double *const state[GFB_NUM_STATE]);
/* solver_interface::init(var, tol, max_iter)  (solver.hpp:254-274).  var = "kx"|"ky"|"kz"|"w"|"x"|"y"|"z"
 * or "" for init() without a solve.  mode 0 = device-resident per-ray Newton, 1 = the reference's
 * host-driven ensemble-maximum loop (workflow.hpp:179-205). */
int gfb_rays_init(gfb_rays *r, const char *var, double tolerance, size_t max_iterations, int mode);
/* solver_interface::compile  (solver.hpp:303-349). */
int gfb_rays_compile(gfb_rays *r);
/* num_steps calls of solver_interface::step  (solver.hpp:382-384). */
int gfb_rays_step(gfb_rays *r, size_t num_steps);
/* work.wait(). */
int gfb_rays_wait(gfb_rays *r);
/* solver_interface::sync_host + read back  (solver.hpp:368-377). Any pointer may be NULL. */
int gfb_rays_get_state(gfb_rays *r, double *const state[GFB_NUM_STATE], double *residual);
/* solver_interface::sync_device from caller memory  (solver.hpp:354-363). */
int gfb_rays_put_state(gfb_rays *r, const double *const state[GFB_NUM_STATE]);
/* sync_device + num_steps x step + sync_host as ONE pipelined call from and to host memory
 * (solver.hpp:354-384): the ensemble is cut into `chunks` pieces whose upload, stepping and
 * read-back overlap.  state_in / state_out: GFB_NUM_STATE host arrays (NULL entries are skipped);
 * residual_out may be NULL.  Pinned memory recommended. */
int gfb_rays_step_host(gfb_rays *r, size_t num_steps, const double *const state_in[GFB_NUM_STATE],
                       double *const state_out[GFB_NUM_STATE], double *residual_out, int chunks);
/* Trajectory output (solver_interface::write_step called every sub_steps, xrays.cpp:246-259):
 * num_blocks times { sub_steps RK steps; snapshot of t, w, x, y, z, kx, ky, kz, residual }.
 * `out` receives num_blocks records of 9 arrays of num_rays doubles ([block][9][ray]); the
 * device->host transfer of block b overlaps the stepping of block b + 1.  Pinned memory recommended. */
int gfb_rays_trace(gfb_rays *r, size_t num_blocks, size_t sub_steps, double *out);
/* Override the automatic choice (see bin_rays above) with a 1-D grid on one state array.
 * Keeps the rays sorted by the table cell of that array while stepping (gfb_bin_rays in gfb200.h):
 * cell = trunc(clamp((state[which] - lo)/(hi - lo)*cells, 0, cells - 1)).  The state is re-sorted before a
 * block of steps when it is in the caller's order, or after `rebin_every` steps (0 = never re-sort while
 * stepping), and restored before every call that reads or writes rays by index, so the caller never sees
 * the permutation.  For VMEC (which = GFB_X, the radial coordinate s in [0, 1], cells = the radial grid) a
 * warp then shares the Fourier coefficient rows it reads.  which_state < 0 switches binning off. */
int gfb_rays_set_binning(gfb_rays *r, int which_state, double lo, double hi, unsigned cells, size_t rebin_every);
/* Trace with power absorption (options "absorption=1").  Replaces the second and third stage of the
 * reference driver, which re-read the trajectory files: absorption::weak_damping
 * (absorption.hpp:327-484, run per record at xrays.cpp:556-558) and bin_power (xrays.cpp:674-793),
 * plus the binning of utilities/bin.py:53-106.  Per block: sub_steps RK steps, then on the state left
 * in device memory  k_amp = |k| - D_warm/(k_hat . dD_cold/dk),  dl = |X - X_last|,
 * p_next = exp(-2 k_sum), d_power = |p_next - power|, k_sum += Im k_amp dl, power = p_next.
 *   records : NULL or [num_blocks][9][num_rays] as gfb_rays_trace
 *   absorbed: NULL or [num_blocks][3][num_rays]  Im k_amp, power, d_power
 *   profile : NULL or [bins[0]][bins[1]][bins[2]] doubles; d_power of every record is ADDED at the ray
 *             position (half-open uniform bins between lo[3] and hi[3]); pass zeros to start
 * The first call (or gfb_rays_absorption_reset) starts a power calculation at the current state:
 * X_last = X, power = 1, k_sum = 0; further calls continue it. */
int gfb_rays_trace_absorb(gfb_rays *r, size_t num_blocks, size_t sub_steps, double *records, double *absorbed,
                          double *profile, const double *lo, const double *hi, const int *bins);
int gfb_rays_absorption_reset(gfb_rays *r);
/* One output block of the same pipeline (replaces one iteration of the record loops of
 * absorption::weak_damping::run, absorption.hpp:466-483, and bin_power, xrays.cpp:757-776, plus the masks of
 * utilities/bin.py:53-106 for that record) with the profile RESIDENT ON THE DEVICE and no host
 * synchronisation: sub_steps RK steps, weak damping, power, then d_power is ADDED to `profile_device`
 * (bins[0]*bins[1]*bins[2] doubles of device memory owned by the caller, e.g. a torch tensor that is
 * then all-reduced with NCCL) on the tracer's stream (gfb_stream(gfb_rays_ctx(r))).  Config 3 of
 * BASELINE.json: the per-GPU histogram that is summed over GPUs once per output block. */
int gfb_rays_deposit_block(gfb_rays *r, size_t sub_steps, double *profile_device,
                           const double *lo, const double *hi, const int *bins);
/* solver "adaptive_rk4" (solver.hpp:881-1006): the per-ray step length the solver's Newton item chose last. */
int gfb_rays_get_dt(gfb_rays *r, double *out);
/* Running absorption state in the caller's ray order (the variables kamp / power / d_power the reference
 * writes per record, absorption.hpp:452-456, xrays.cpp:716-724): out[0..2] = Im k_amp, power, d_power
 * (num_rays doubles each, NULL entries skipped) as left by the last absorption block. */
int gfb_rays_get_absorbed(gfb_rays *r, double *const out[3]);
/* The device-resident profile the last gfb_rays_trace_absorb accumulated into: buffer key in
 * gfb_rays_ctx(r) (for gfb_allreduce_sum_f64 / gfb_copy_d2h) and its number of cells. */
int gfb_rays_profile(gfb_rays *r, uint64_t *key, size_t *cells);
/* Device pointer of state array `which` (GFB_T..GFB_KZ) or of the residual (which = GFB_NUM_STATE). */
int gfb_rays_device_ptr(gfb_rays *r, int which, void **device_ptr);
/* The underlying device context (timers, launch counters, deposit, ...). */
gfb_ctx *gfb_rays_ctx(gfb_rays *r);
/* Emitted CUDA text of the step kernel; statements / remaining divides / shared reciprocals in it. */
const char *gfb_rays_source(gfb_rays *r);
int gfb_rays_kernel_stats(gfb_rays *r, int *statements, int *divides, int *reciprocals,
                          int *registers, int *local_bytes, int *smem_bytes);

/* Right-hand side only: evaluates dx/dt,dy/dt,dz/dt,dkx/dt,dky/dt,dkz/dt,D (dispersion.hpp:1387-1433)
 * for host states; out = 7 arrays of num_rays doubles.  Used by parity tests. */
int gfb_rays_rhs(gfb_rays *r, double *const out[7]);

/* Boris push in an equilibrium field exactly as graph_korc/xkorc.cpp:40-121 builds it.
 * state order: x, y, z, ux, uy, uz (u in units of c), then gamma is produced by the pre-item. */
typedef struct gfb_boris gfb_boris;
gfb_boris *gfb_boris_create(const char *equilibrium, const char *table_file, size_t num_particles,
                            double dt, int device, const char *options);
void gfb_boris_destroy(gfb_boris *b);
int gfb_boris_set_state(gfb_boris *b, const double *const state[6]);
int gfb_boris_compile(gfb_boris *b);          /* also runs the initialize_gamma pre-item */
int gfb_boris_step(gfb_boris *b, size_t num_steps);
int gfb_boris_get_state(gfb_boris *b, double *const state[7]);
/* Keep the particles sorted by the (R, Z) cell of the field tables while stepping (gfb_bin_rays_rz in
 * gfb200.h): lo/hi/cells = {R, Z} grid, re-sorted every `rebin_every` steps (0: only when the order was
 * restored by a call that reads or writes particles by index).  cells = NULL switches it off. */
int gfb_boris_set_binning(gfb_boris *b, const double *lo, const double *hi, const unsigned *cells, size_t rebin_every);
int gfb_boris_info(gfb_boris *b, double *b0, double *larmor_radius);
gfb_ctx *gfb_boris_ctx(gfb_boris *b);

#ifdef __cplusplus
}
#endif
#endif /* GFB_RAYS_H */
