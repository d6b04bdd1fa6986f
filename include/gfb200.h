/*------------------------------------------------------------------------------
 *  gfb200.h -- C ABI of the B200 device layer (libgfb200.so).
 *
 *  This is the drop-in boundary: every entry point replaces one member of the
 *  reference's device context, `gpu::cuda_context<T, SAFE_MATH>`
 *  (/root/reference/graph_framework/cuda_context.hpp), which is the only class
 *  of the reference that touches a device API.  `jit::context` (jit.hpp:80-338)
 *  is its only caller.  Signatures use plain pointers and sizes only.
 *
 *  Error behaviour: the reference asserts in debug builds and ignores CUresults
 *  in release builds (cuda_context.hpp:30-53).  Here every call returns 0 on
 *  success, non-zero on failure, and gfb_last_error() returns the text; the C++
 *  adapters print it and abort, so a failure is never silent.  There is no CPU
 *  fallback: without a CUDA device gfb_ctx_create fails.
 *----------------------------------------------------------------------------*/
#ifndef GFB200_H
#define GFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfb_ctx gfb_ctx;          /* one device context (cuda_context object) */
typedef struct gfb_kernel gfb_kernel;    /* one launchable (the lambda of create_kernel_call) */

/* cuda_context::max_concurrency  (cuda_context.hpp:121-126): number of CUDA devices. */
int gfb_device_count(void);
/* Text of the last failure on this thread ("" if none). */
const char *gfb_last_error(void);
/* Used by the host-side bindings of this library to report their own failures. */
void gfb_set_last_error(const char *text);
/* Library version string. */
const char *gfb_version(void);

/* cuda_context::cuda_context(index)  (cuda_context.hpp:139-148). */
gfb_ctx *gfb_ctx_create(int device);
/* cuda_context::~cuda_context  (cuda_context.hpp:153-185): frees module, buffers, stream. */
void gfb_ctx_destroy(gfb_ctx *ctx);
/* Device name / SM count / clock of the context's device. */
int gfb_ctx_device_info(gfb_ctx *ctx, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor);

/* cuda_context::compile  (cuda_context.hpp:194-302): NVRTC for sm_100a -> cubin -> module.
 * `source` is the emitted bodies; the hand-written skeleton text is prepended by
 * the library.  `options` may be NULL or a space separated list of extra NVRTC flags. */
int gfb_compile(gfb_ctx *ctx, const char *source, const char *const *names, int num_names, const char *options);
/* Blocks per SM the last gfb_compile settled on (4..1: the largest without register spills;
 * 0 when the caller pinned -DGFB_MIN_BLOCKS itself). */
int gfb_compiled_min_blocks(gfb_ctx *ctx);
/* NVRTC only, no device needed: used by the CPU test-suite and by tools.  On success
 * *cubin (malloc'ed, caller frees with gfb_free) holds the sm_100a cubin. */
int gfb_compile_to_cubin(const char *source, const char *options, void **cubin, size_t *cubin_size, char **log);
void gfb_free(void *p);
/* Full text (skeleton + bodies) of the last gfb_compile on this context; jit::context::print_source. */
const char *gfb_source(gfb_ctx *ctx);
const char *gfb_compile_log(gfb_ctx *ctx);
/* The options NVRTC was given for the module in use (the reference's fixed option array is
 * cuda_context.hpp:245-254; here: caller's options + the blocks/SM promise that was chosen): together with
 * gfb_source() this identifies the binary. */
const char *gfb_compile_options(gfb_ctx *ctx);

/* Buffers are keyed by a 64-bit key (the host node address, as the reference keys its
 * std::map<leaf_node *, CUdeviceptr>, cuda_context.hpp:82).  First call allocates `bytes`
 * and uploads `init` (may be NULL = zero fill); later calls return the same buffer.
 * (create_kernel_call's allocation loop, cuda_context.hpp:330-364.) */
int gfb_buffer(gfb_ctx *ctx, uint64_t key, size_t bytes, const void *init, void **device_ptr);
/* Adopt externally owned device memory (e.g. a torch tensor) under a key. */
int gfb_buffer_import(gfb_ctx *ctx, uint64_t key, void *device_ptr, size_t bytes);
/* Device pointer and size of a buffer, 0 size if unknown. */
int gfb_buffer_lookup(gfb_ctx *ctx, uint64_t key, void **device_ptr, size_t *bytes);

/* cuda_context::create_kernel_call  (cuda_context.hpp:316-531).
 * ptr_keys: buffer keys in kernel pointer order (inputs, outputs, table groups).
 * kind: 0 generic item, 1 runge-kutta item, 2 device-resident Newton item.
 * can_repeat: 1 when repeated runs may be fused into one multi-step launch (the item has
 * setters and each ray only reads its own state); 0 makes repeated runs collapse to a single
 * step (no setters: the item is idempotent); 2 launches every run separately (the item gathers
 * from an array it also rewrites, so steps need a grid-wide boundary). */
int gfb_kernel_create(gfb_ctx *ctx, const char *name, const uint64_t *ptr_keys, int num_ptrs,
                      size_t num_rays, unsigned block_size, size_t dynamic_smem, int kind, int can_repeat,
                      gfb_kernel **kernel);
/* The launch lambda.  Launches are DEFERRED: consecutive runs of the same kernel are
 * counted and issued as one launch whose step loop stays in registers; any
 * observation (wait, copies, max, host pointer, a different kernel) flushes. */
int gfb_kernel_run(gfb_kernel *kernel);
/* Immediate launch of `steps` fused steps (flushes anything pending first). */
int gfb_kernel_launch(gfb_kernel *kernel, unsigned steps);
/* Host-to-host pipelined run: the rays are cut into `chunks` contiguous pieces; for each piece the
 * per-ray pointer slots that have a host source are uploaded, `steps` fused steps run on that
 * piece, and the slots that have a host destination are read back -- upload, compute and
 * read-back of different pieces overlap on three streams (two copy engines + SMs).
 * host_src / host_dst: one pointer per per-ray slot (the first num_ray_slots pointer slots of the
 * kernel: inputs then outputs), NULL = no transfer for that slot.  Pinned memory recommended.
 * Pieces end on whole waves of thread blocks; with 3 or more chunks the first piece is one wave and
 * the last one the tail wave, because the first upload and the last read-back overlap nothing.
 * Returns after everything has completed. */
int gfb_kernel_run_from_host(gfb_kernel *kernel, unsigned steps, int num_ray_slots,
                             const void *const *host_src, void *const *host_dst, int chunks);
/* Set scalar[index] passed to the kernel (Newton tolerance, ...). */
int gfb_kernel_set_scalar(gfb_kernel *kernel, int index, double value);
/* registers/thread, static smem, local (spill) bytes/thread, max threads/block. */
int gfb_kernel_attributes(gfb_kernel *kernel, int *regs, int *static_smem, int *local_bytes, int *max_threads);
/* Number of device launches issued so far by this context (bench.py's gpu_launches). */
uint64_t gfb_launch_count(gfb_ctx *ctx);
/* Upper bound of fused steps per launch (default 1024). */
int gfb_set_max_fused_steps(gfb_ctx *ctx, unsigned steps);
int gfb_flush(gfb_ctx *ctx);

/* cuda_context::create_max_call  (cuda_context.hpp:540-576) + create_reduction (:954-995):
 * maximum of n doubles of a buffer, returned to the host.  NaN elements are IGNORED, as by the
 * reference's max() (= fmax): a convergence loop on this value keeps iterating for the good rays.
 * The result is NaN only when every element is NaN or n == 0. */
int gfb_max(gfb_ctx *ctx, uint64_t key, size_t n, double *result);
/* cuda_context::wait  (cuda_context.hpp:581-584). */
int gfb_wait(gfb_ctx *ctx);
/* cuda_context::copy_to_device / copy_to_host  (cuda_context.hpp:625-643). bytes == 0 copies the whole buffer. */
int gfb_copy_h2d(gfb_ctx *ctx, uint64_t key, const void *source, size_t bytes);
int gfb_copy_d2h(gfb_ctx *ctx, uint64_t key, void *destination, size_t bytes);
/* cuda_context::get_buffer  (cuda_context.hpp:1002-1004): host-dereferenceable view, a
 * pinned mirror refreshed by gfb_wait (the reference returns managed memory). */
int gfb_host_ptr(gfb_ctx *ctx, uint64_t key, void **host_ptr);
/* cuda_context::check_value  (cuda_context.hpp:613-617). */
int gfb_check_value(gfb_ctx *ctx, uint64_t key, size_t index, double *value);

/* Asynchronous snapshot of `num_keys` buffers to host memory (trajectory output; the reference's
 * write_step, solver.hpp:418-424, hands the copy to a side thread).  Flushes pending steps, copies
 * every buffer device->device into one of two staging slots on the compute stream, then moves
 * the slot to `host_destination` (num_keys consecutive blocks of `bytes_each`; pinned memory
 * recommended) on a second stream, so the next block of steps overlaps the transfer.
 * gfb_wait() completes all outstanding snapshots.  While rays are binned (gfb_bin_rays) the keys must be
 * whole per-ray arrays; the snapshot delivers them in the caller's order. */
int gfb_snapshot_async(gfb_ctx *ctx, const uint64_t *keys, int num_keys, size_t bytes_each, void *host_destination);

/* Ray binning.  Rays are independent, so their order in the SoA arrays is free: gfb_bin_rays sorts the
 * rays by the table cell of one array (cell = trunc(clamp((v - lo)/(hi - lo)*cells, 0, cells - 1)), the
 * piecewise index rule) and applies the permutation to every listed per-ray buffer, so that a warp
 * works on rays of the same cell and shares the coefficient rows it gathers.  gfb_unbin_rays restores
 * the caller's order (no-op when not binned).  No reference counterpart: the reference never reorders. */
int gfb_bin_rays(gfb_ctx *ctx, uint64_t sort_key, double lo, double hi, unsigned cells,
                 const uint64_t *keys, int num_keys, size_t n);
/* The same with the (R, Z) cell of an axisymmetric 2-D table: xyz_keys = the x, y, z arrays,
 * R = sqrt(x^2 + y^2); lo/hi/cells = {R, Z}; cell = i_R*cells[1] + i_Z (the piecewise_2D layout). */
int gfb_bin_rays_rz(gfb_ctx *ctx, const uint64_t *xyz_keys, const double *lo, const double *hi, const unsigned *cells,
                    const uint64_t *keys, int num_keys, size_t n);
int gfb_unbin_rays(gfb_ctx *ctx, const uint64_t *keys, int num_keys, size_t n);
int gfb_is_binned(gfb_ctx *ctx);
/* How far the order has decayed: the fraction of neighbouring slots whose rays sit in different cells
 * (about occupied_cells/n right after a sort, towards 1 for a random order).  1 sort key: the 1-D grid;
 * 3 sort keys: the (R, Z) grid of gfb_bin_rays_rz.  Synchronises the stream. */
int gfb_bin_disorder(gfb_ctx *ctx, const uint64_t *sort_keys, int num_sort_keys, const double *lo, const double *hi,
                     const unsigned *cells, size_t n, double *fraction);
/* Device -> host copy of one per-ray array of n doubles in the CALLER's order, whether or not the rays
 * are binned at the moment (binned: un-permuted through a scratch buffer; the device order is kept). */
int gfb_copy_rays_d2h(gfb_ctx *ctx, uint64_t key, void *destination, size_t n);

/* Page-locked host memory for snapshot / step-from-host buffers (any thread, any context). */
int gfb_host_alloc(size_t bytes, void **host_ptr);
int gfb_host_free(void *host_ptr);

/* Device timing on the context's stream (CUDA events). */
int gfb_timer_start(gfb_ctx *ctx);
int gfb_timer_stop(gfb_ctx *ctx, float *milliseconds);
/* The context's cudaStream_t, for callers that share it with torch. */
void *gfb_stream(gfb_ctx *ctx);

/* Power-deposition profile (north-star config 3; algorithm of utilities/bin.py:53-106):
 * hist[ix, iy, iz] += weight[i] for rays inside the half-open box. All device pointers. */
int gfb_deposit(gfb_ctx *ctx, const double *x, const double *y, const double *z, const double *weight,
                size_t n, double *hist, const double *lo, const double *hi, const int *bins);

/* Sum of one buffer per device over all devices of ONE process, result left in every buffer
 * (the reduction of the binned power-deposition profile for the reference's
 * thread-per-device model, xrays.cpp:419-527, where torch.distributed/NCCL communicators do not exist).
 * keys[g] names n doubles in ctxs[g].  Reduce-scatter + all-gather written on NVLink peer memory: device g
 * sums slice g of every buffer with P2P loads, always in device order (bit-identical sums everywhere), then
 * pulls the other slices; ordering is by CUDA events, the host does not block (call gfb_wait to observe).
 * Devices without peer access are staged through device 0.  Call from one thread while no other thread
 * uses these contexts.  (One process per GPU: use NCCL -- graph_framework_b200/parallel.py.) */
int gfb_allreduce_sum_f64(gfb_ctx *const *ctxs, int num_ctx, const uint64_t *keys, size_t n);

/* Measured FP64 FMA rate of the device in TFLOP/s: independent DFMA chains whose multiplier and addend
 * are uniform values (the friendliest stream there is) ... */
int gfb_measure_fp64_peak(gfb_ctx *ctx, double *tflops, float *milliseconds);
/* ... and the same chains with all three operands in per-thread registers, the shape compiled ray code has. */
int gfb_measure_fp64_peak_registers(gfb_ctx *ctx, double *tflops, float *milliseconds);
/* L2 flush helper for benchmarks: writes a buffer larger than L2. */
int gfb_flush_l2(gfb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GFB200_H */
