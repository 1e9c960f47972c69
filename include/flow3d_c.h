/* flow3d_c.h -- C ABI of the B200-native dense 3D variational optical-flow solve.
 *
 * This is the drop-in boundary for the hot path of axruff/cuda-flow3d:
 * plain pointers and sizes, no C++/torch types, every entry point returns an int status
 * (0 = FLOW3D_OK, negative = error), never throws, never aborts.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * tree).  Stage functions work on DEVICE pointers; the solver object owns its device arena and
 * offers a host-buffer call (the reference's OpticalFlowE::ComputeFlow contract) and a
 * device-buffer call.
 *
 * Volume layout (all stage functions): x-fastest fp32, element (x,y,z) at  (z*h + y)*ld + x,
 * `ld` = row pitch in floats, ld >= w, ld % 4 == 0, base pointer 16-byte aligned.  Columns
 * [w, ld) are padding: never interpreted, may be overwritten.  The reference instead keeps every
 * level in the top-left-front corner of a full-resolution pitched container
 * (src/kernels/solve_3d.cu:26); that is a storage choice with identical results.
 *
 * `stream` arguments are cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef FLOW3D_C_H_
#define FLOW3D_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOW3D_OK 0
#define FLOW3D_ERR_INVALID_ARG (-1)   /* null pointer, bad dims, misaligned ld/pointer */
#define FLOW3D_ERR_UNSUPPORTED (-2)   /* e.g. median radius outside {1,3,5,7}, blur radius > 32 */
#define FLOW3D_ERR_CUDA (-3)          /* a CUDA runtime call failed; see flow3d_last_cuda_error() */
#define FLOW3D_ERR_NO_DEVICE (-4)     /* no CUDA device: there is NO CPU fallback */
#define FLOW3D_ERR_OUT_OF_MEMORY (-5)
#define FLOW3D_ERR_NOT_INITIALIZED (-6)

/* The nine named solver parameters of the reference (src/main.cpp:77-85,157-165; read at
 * src/optical_flow/optical_flow_e.cpp:150-158).  Same names, same types. */
typedef struct flow3d_params {
  size_t warp_levels_count;      /* default 40   */
  float warp_scale_factor;       /* default 0.95 */
  size_t outer_iterations_count; /* default 40   */
  size_t inner_iterations_count; /* default 5    */
  float equation_alpha;          /* default 7.5  */
  float equation_smoothness;     /* default 0.001 */
  float equation_data;           /* default 0.001 */
  size_t median_radius;          /* default 5 (window edge length: 1 = off, 3, 5, 7) */
  float gaussian_sigma;          /* default 2.0 (<= 0 = no pre-blur) */
} flow3d_params;

typedef struct flow3d_solver flow3d_solver; /* opaque */

/* z-slab view of a level that is sharded along z across GPUs (one process per GPU).  The buffers
 * passed with a slab hold `dims[2]` consecutive planes of the level, local plane 0 being global plane
 * `z0_global`; mirror boundary conditions and the face weights apply at the GLOBAL faces
 * (0 and depth_global-1) only, elsewhere the z neighbours are ordinary (ghost) planes of the buffer.
 * [z_begin, z_end) is the range of local planes the call computes; every plane it reads (z_begin-1 ..
 * z_end for the 7-point stencils, +-2 for the 5^3 median) must exist in the buffer unless it lies
 * beyond a global face.  The reference's own precedent is the z-slab + 1-plane mirrored halo scheme of
 * its out-of-core path (src/cuda_operations/partial_data/cuda_operation_solve_p.cpp:358-417). */
typedef struct flow3d_zslab {
  size_t z0_global;
  size_t depth_global;
  size_t z_begin;
  size_t z_end;
} flow3d_zslab;

/* ---- misc ------------------------------------------------------------------------------- */
int flow3d_version(void);
const char* flow3d_status_string(int status);
/* text of the last CUDA error seen by this thread's calls ("" if none) */
const char* flow3d_last_cuda_error(void);
/* number of visible CUDA devices, or a negative status */
int flow3d_device_count(void);
void flow3d_default_params(flow3d_params* p);
/* kernels launched by this library since process start / last reset (bench.py's gpu_launches) */
uint64_t flow3d_launch_count(void);
void flow3d_reset_launch_count(void);

/* ---- level schedule (host arithmetic only) ------------------------------------------------
 * replaces OpticalFlowBase::GetMaxWarpLevel (src/optical_flow/optical_flow_base.cpp:31-56) */
size_t flow3d_max_warp_level(size_t width, size_t height, size_t depth, float scale_factor);
/* replaces the per-level size/spacing arithmetic of src/optical_flow/optical_flow_e.cpp:262-268;
 * dims = {w,h,d} of `level`, h = {hx,hy,hz} */
int flow3d_level_geometry(size_t width, size_t height, size_t depth, float scale_factor, int level,
                          size_t dims[3], float h[3]);
/* row pitch (floats) the solver uses for a level of width w */
size_t flow3d_aligned_ld(size_t w);

/* ---- device memory helpers (so callers need no CUDA headers) ------------------------------ */
int flow3d_set_device(int device);
int flow3d_malloc(void** dev_ptr, size_t bytes);
int flow3d_free(void* dev_ptr);
int flow3d_memset(void* dev_ptr, int value, size_t bytes, void* stream);
/* copy a tight host volume (w*h*d floats) into a pitched device volume and back */
int flow3d_upload(const float* host, float* dev, const size_t dims[3], size_t ld, void* stream);
int flow3d_download(const float* dev, float* host, const size_t dims[3], size_t ld, void* stream);
int flow3d_stream_synchronize(void* stream);
/* page-locked host memory for fast asynchronous H2D/D2H (FLOW3D_ERR_NO_DEVICE without a device;
 * the reference has the same option behind ALLOCATE_PINNED_MEMORY, src/data_types/data3d.cpp:28-60) */
int flow3d_host_alloc(void** host_ptr, size_t bytes);
int flow3d_host_free(void* host_ptr);
/* name of a device into buf (NUL-terminated, truncated to n) */
int flow3d_device_name(int device, char* buf, size_t n);

/* ---- stage functions (device pointers) ----------------------------------------------------- */

/* Separable Gaussian pre-blur, zero padding, taps from sigma exactly as the reference computes
 * them.  Replaces CudaOperationConvolution3D::Execute
 * (src/cuda_operations/entire_data/cuda_operation_convolution.cpp:134-184) and
 * convolutionRows/Columns/SlicesKernel (src/kernels/convolution_3d.cu:75-372), with guarded
 * edges (the reference kernels write out of bounds when a dimension is not a multiple of 4).
 * in != out; tmp is a scratch volume of the same shape. */
int flow3d_gauss_blur(const float* in, float* out, float* tmp, const size_t dims[3], size_t ld,
                      float sigma, void* stream);

/* Separable box (area-average) resample X -> Y -> Z, shrink or grow.  Replaces
 * CudaOperationResample::Execute (.../cuda_operation_resample.cpp:72-106) and
 * resample_{x,y,z}_3d (src/kernels/resample_3d.cu:28-161).
 * tmp_a: >= flow3d_aligned_ld(out_w)*in_h*in_d floats, tmp_b: >= ld(out_w)*out_h*in_d floats. */
int flow3d_resample(const float* in, const size_t in_dims[3], size_t in_ld, float* out,
                    const size_t out_dims[3], size_t out_ld, float* tmp_a, float* tmp_b,
                    void* stream);

/* Backward trilinear warp of frame 1 by the flow; out-of-volume / NaN targets fall back to
 * frame 0.  Replaces CudaOperationRegistration::Execute (.../cuda_operation_registration.cpp:
 * 70-131) and registration_3d (src/kernels/registration_3d.cu:28-82).  out must not alias f1. */
int flow3d_warp(const float* f0, const float* f1, const float* u, const float* v, const float* w,
                const size_t dims[3], size_t ld, const float h[3], float* out, void* stream);

/* Image derivatives of the (frame0, warped frame1) pair: fx, fy, fz (central differences of
 * (f0+f1w)/(4h), mirror borders) and ft = f1w - f0.  These are the quantities the reference
 * recomputes inside every launch (src/kernels/solve_3d.cu:220-233 and :425-438); computing them
 * once per level gives bit-identical values. */
int flow3d_derivatives(const float* f0, const float* f1w, const size_t dims[3], size_t ld,
                       const float h[3], float* fx, float* fy, float* fz, float* ft, void* stream);

/* Fused warp + derivatives: same results as flow3d_warp followed by flow3d_derivatives, without
 * writing the warped volume. */
int flow3d_warp_derivatives(const float* f0, const float* f1, const float* u, const float* v,
                            const float* w, const size_t dims[3], size_t ld, const float h[3],
                            float* fx, float* fy, float* fz, float* ft, void* stream);

/* Robust weights phi (smoothness) and ksi (data).  Replaces compute_phi_ksi_3d
 * (src/kernels/solve_3d.cu:33-262; launched at .../cuda_operation_solve.cpp:194-221). */
int flow3d_phi_ksi(const float* fx, const float* fy, const float* fz, const float* ft,
                   const float* u, const float* v, const float* w, const float* du,
                   const float* dv, const float* dw, const size_t dims[3], size_t ld,
                   const float h[3], float eps_smooth, float eps_data, float* phi, float* ksi,
                   void* stream);

/* One Jacobi sweep (du,dv,dw) -> (du_out,dv_out,dw_out).  Replaces solve_3d
 * (src/kernels/solve_3d.cu:264-508; launched at .../cuda_operation_solve.cpp:223-257). */
int flow3d_sweep(const float* fx, const float* fy, const float* fz, const float* ft,
                 const float* u, const float* v, const float* w, const float* du, const float* dv,
                 const float* dw, const float* phi, const float* ksi, const size_t dims[3],
                 size_t ld, const float h[3], float alpha, float* du_out, float* dv_out,
                 float* dw_out, void* stream);

/* The whole inner solver of one level: du=dv=dw=0, then `outer` x (phi/ksi + `inner` sweeps).
 * Replaces CudaOperationSolve::Execute (.../cuda_operation_solve.cpp:75-281).  On return
 * du/dv/dw hold the final iterate.  scratch: 5 volumes of ld*h*d floats (phi, ksi, 3 ping-pong). */
int flow3d_solve_level(const float* fx, const float* fy, const float* fz, const float* ft,
                       const float* u, const float* v, const float* w, float* du, float* dv,
                       float* dw, float* scratch, const size_t dims[3], size_t ld,
                       const float h[3], size_t outer, size_t inner, float alpha, float eps_smooth,
                       float eps_data, void* stream);

/* u += du, v += dv, w += dw in one launch.  Replaces three add_3d launches
 * (src/kernels/add_3d.cu:26-41; src/optical_flow/optical_flow_e.cpp:420-438). */
int flow3d_add3(float* u, float* v, float* w, const float* du, const float* dv, const float* dw,
                const size_t dims[3], size_t ld, void* stream);

/* radius^3 median with mirror borders; radius = window edge length (1 = copy, even -> radius-1,
 * supported 3/5/7).  Replaces CudaOperationMedian::Execute (.../cuda_operation_median.cpp:72-149)
 * and median_3d (src/kernels/median_3d.cu:49-299).  in != out. */
int flow3d_median(const float* in, float* out, const size_t dims[3], size_t ld, size_t radius,
                  void* stream);

/* Host only: the taps flow3d_gauss_blur uses for `sigma` -- CudaOperationConvolution3D::ComputeGaussianKernel
 * with precision 3 and pixel size 1.0 as Execute() calls it
 * (src/cuda_operations/entire_data/cuda_operation_convolution.cpp:85-108, :159): radius (size_t)(3*sigma),
 * 2*radius+1 taps.  Returns FLOW3D_ERR_UNSUPPORTED if the radius exceeds 32 (the kernels' limit) or
 * FLOW3D_ERR_INVALID_ARG if `capacity` floats cannot hold the taps.  Checked against the reference's own function
 * by tests/test_cabi_cpu.py. */
int flow3d_gauss_taps(float sigma, float* taps, size_t capacity, size_t* radius);

/* ---- z-slab (multi-GPU) variants: identical arithmetic, boundary handling per flow3d_zslab ------ */
/* pre-blur of a z-slab of the full-resolution frame: zero padding at the global faces only; the
 * slab must hold (size_t)(3*sigma) ghost planes around [z_begin, z_end) (or reach a global face);
 * planes outside the range receive the x/y-blurred intermediate and must not be used */
int flow3d_gauss_blur_slab(const float* in, float* out, float* tmp, const size_t dims[3], size_t ld,
                           const flow3d_zslab* slab, float sigma, void* stream);
int flow3d_sweep_slab(const float* fx, const float* fy, const float* fz, const float* ft,
                      const float* u, const float* v, const float* w, const float* du,
                      const float* dv, const float* dw, const float* phi, const float* ksi,
                      const size_t dims[3], size_t ld, const flow3d_zslab* slab, const float h[3],
                      float alpha, float* du_out, float* dv_out, float* dw_out, void* stream);
int flow3d_phi_ksi_slab(const float* fx, const float* fy, const float* fz, const float* ft,
                        const float* u, const float* v, const float* w, const float* du,
                        const float* dv, const float* dw, const size_t dims[3], size_t ld,
                        const flow3d_zslab* slab, const float h[3], float eps_smooth, float eps_data,
                        float* phi, float* ksi, void* stream);
/* f1 lives in its own slab (f1_dims[2] planes starting at global plane f1_z0_global) because the warp
 * reaches max|w|/hz planes beyond the planes being computed */
int flow3d_warp_derivatives_slab(const float* f0, const float* f1, size_t f1_z0_global,
                                 size_t f1_depth_local, const float* u, const float* v,
                                 const float* w, const size_t dims[3], size_t ld,
                                 const flow3d_zslab* slab, const float h[3], float* fx, float* fy,
                                 float* fz, float* ft, void* stream);
/* One Jacobi sweep with an explicit launch shape (tests and tuning scripts; production launches take the
 * shape from the tuning table): variant 0 = register-marching warps (vec = voxels per lane 1/2/4), 1 / 2 =
 * TMA-staged 64x8 / 32x16 tiles; nchunks = z chunks (0 = default).  ksi_out != NULL: the sweep computes
 * the data-term weight from the iterate it reads (first sweep of an outer iteration), stores it there and
 * ignores `ksi`.  slab may be NULL (whole volume).  Every shape gives identical bits.
 * FLOW3D_ERR_UNSUPPORTED when the variant cannot run this level. */
int flow3d_sweep_shape(const float* fx, const float* fy, const float* fz, const float* ft, const float* u,
                       const float* v, const float* w, const float* du, const float* dv, const float* dw,
                       const float* phi, const float* ksi, const size_t dims[3], size_t ld,
                       const flow3d_zslab* slab, const float h[3], float alpha, float eps_data, float* du_out,
                       float* dv_out, float* dw_out, float* ksi_out, int variant, int vec, int nchunks,
                       void* stream);
/* One outer iteration on a z-slab in ONE call: phi/ksi on the slab's [z_begin, z_end), then `inner`
 * Jacobi sweeps on ranges that shrink by one plane per sweep on every side that is NOT a global face
 * (sweep j computes [z_begin + j, z_end - j) there), ping-ponging between (du,dv,dw) and (tdu,tdv,tdw).
 * This is the communication-free part of the z-sharded solve (ghost depth inner+1).  On return
 * *result_in_tmp is 1 when the final iterate is in (tdu,tdv,tdw), 0 when it is in (du,dv,dw). */
int flow3d_outer_iteration_slab(const float* fx, const float* fy, const float* fz, const float* ft,
                                const float* u, const float* v, const float* w, float* du, float* dv,
                                float* dw, float* tdu, float* tdv, float* tdw, float* phi, float* ksi,
                                const size_t dims[3], size_t ld, const flow3d_zslab* slab,
                                const float h[3], size_t inner, float alpha, float eps_smooth,
                                float eps_data, int* result_in_tmp, void* stream);
/* The same outer iteration split for communication overlap.  [own_begin, own_end) = the local planes this
 * rank owns (the rest of [z_begin, z_end) are ghost planes filled by a neighbour exchange).  part 1
 * ("early") runs, for phi and every sweep, only the planes whose values cannot depend on the ghosts
 * (step j on [own_begin+j+1, own_end-j-1), extended to a global face where there is one): it may start
 * before the ghosts of the starting iterate have arrived.  part 2 ("late") runs the remaining planes of
 * every step and must follow part 1 and the exchange.  part 1 + part 2 = part 0 = flow3d_outer_iteration_slab,
 * bit for bit; early and late launches touch disjoint planes of every buffer at every step. */
int flow3d_outer_iteration_slab_part(const float* fx, const float* fy, const float* fz, const float* ft,
                                     const float* u, const float* v, const float* w, float* du, float* dv,
                                     float* dw, float* tdu, float* tdv, float* tdw, float* phi, float* ksi,
                                     const size_t dims[3], size_t ld, const flow3d_zslab* slab,
                                     const float h[3], size_t inner, float alpha, float eps_smooth,
                                     float eps_data, int part, size_t own_begin, size_t own_end,
                                     int* result_in_tmp, void* stream);
int flow3d_median_slab(const float* in, float* out, const size_t dims[3], size_t ld,
                       const flow3d_zslab* slab, size_t radius, void* stream);
/* resample with the input and the output each given as a slab of its level; x and y passes run on
 * all local input planes, the z pass produces output planes [out_slab->z_begin, z_end).
 * tmp_a: >= ld(out_w)*in_h*in_depth_local, tmp_b: >= ld(out_w)*out_h*in_depth_local floats */
int flow3d_resample_slab(const float* in, const size_t in_dims[3], size_t in_ld,
                         const flow3d_zslab* in_slab, float* out, const size_t out_dims[3],
                         size_t out_ld, const flow3d_zslab* out_slab, float* tmp_a, float* tmp_b,
                         void* stream);
/* max |x| over the w*h*d samples of a pitched device volume (padding columns ignored), written to
 * *out_dev (one device float) */
int flow3d_absmax(const float* dev, const size_t dims[3], size_t ld, float* out_dev, void* stream);

/* ---- solver object -------------------------------------------------------------------------
 * Replaces OpticalFlowE::{Initialize, ComputeFlow, Destroy}
 * (src/optical_flow/optical_flow_e.cpp:42-49, 132-601, 603-622). */

/* bytes of device memory a solver for a W x H x D volume allocates */
size_t flow3d_solver_workspace_bytes(size_t width, size_t height, size_t depth);
int flow3d_solver_create(size_t width, size_t height, size_t depth, int device,
                         flow3d_solver** out);
int flow3d_solver_destroy(flow3d_solver* s);

/* Host-buffer solve = the reference's ComputeFlow contract: tight W*H*D fp32 host volumes in,
 * three tight host volumes out; blocks until the results are in host memory. */
int flow3d_solver_compute_host(flow3d_solver* s, const float* frame_0, const float* frame_1,
                               const flow3d_params* params, float* flow_u, float* flow_v,
                               float* flow_w);

/* Device-buffer solve: inputs/outputs are device volumes with pitch `ld` (as laid out by
 * flow3d_upload); asynchronous on `stream`. */
int flow3d_solver_compute_device(flow3d_solver* s, const float* frame_0, const float* frame_1,
                                 size_t ld, const flow3d_params* params, float* flow_u,
                                 float* flow_v, float* flow_w, void* stream);

/* ---- launch shapes ---------------------------------------------------------------------------------
 * The solver kernels (Jacobi sweep, phi) exist in several launch shapes / variants that give identical
 * bits; which one is fastest depends on the level size.  Launches only LOOK UP a per-device table (a
 * missing entry = static heuristic), so flow3d_solver_compute_device and the stage calls stay
 * asynchronous and capturable.  These calls fill the table; they time candidates with CUDA events and
 * are SYNCHRONOUS.  The table is persisted in $FLOW3D_TUNE_CACHE (default ~/.cache/flow3d_b200/, "off"
 * disables); FLOW3D_AUTOTUNE=0 ignores it.  flow3d_solver_compute_host runs a quick pass on first use (static
 * vector width, four chunk lengths, one timing round; kept in memory only); flow3d_solver_tune runs the full pass.
 * flow3d_tune_kernels: scratch = 16 volumes of the slab's size (contents destroyed). */
int flow3d_solver_tune(flow3d_solver* s, const flow3d_params* params);
int flow3d_tune_kernels(const size_t dims[3], size_t ld, const flow3d_zslab* slab, const float h[3],
                        float* scratch, size_t scratch_floats, void* stream);
/* table entry of kernel 0 = sweep, 1 = phi+ksi, 2 = sweep computing ksi, 3 = phi: returns 1 and
 * out = {voxels per lane, z chunks, variant (0 register-marching, 1/2 TMA tiles 64x8 / 32x16)}, 0 if none */
int flow3d_tune_query(int kernel, const size_t dims[3], size_t ld, const flow3d_zslab* slab, int out[3]);

/* Programmatic dependent launch of a level's phi / sweep chain (flow3d_solve_level and the solver object's
 * solves): each of the 6 x outer launches may become resident while its predecessor drains and waits on the
 * device for it to complete -- same kernels, same bits, shorter launch gaps on the small levels.
 * mode 1 = on, 0 = off, -1 = the environment's choice ($FLOW3D_PDL, default on).  Returns the mode in effect
 * (0 also when the driver refused a launch with the attribute and the library fell back). */
int flow3d_set_pdl(int mode);

/* milliseconds of the last compute call, measured with CUDA events: [0] whole call (host call:
 * including H2D/D2H, the reference's own bracket optical_flow_e.cpp:169->579), [1] device-only */
int flow3d_solver_last_timing(const flow3d_solver* s, float ms[2]);

/* Per-stage device timing (CUDA events on the solve's stream), off by default.  When enabled, each
 * compute call accumulates, per stage, the elapsed milliseconds and the units of work processed:
 * voxels for every stage (FLOW3D_STAGE_SWEEP counts voxel-sweeps = voxels x sweeps, the unit of
 * SURVEY.md 8d; FLOW3D_STAGE_PHI_KSI counts voxel-updates).  Query after the call has finished. */
#define FLOW3D_STAGE_BLUR 0
#define FLOW3D_STAGE_RESAMPLE 1
#define FLOW3D_STAGE_WARP 2
#define FLOW3D_STAGE_PHI_KSI 3
#define FLOW3D_STAGE_SWEEP 4
#define FLOW3D_STAGE_UPDATE 5   /* memset of du + flow update */
#define FLOW3D_STAGE_MEDIAN 6
#define FLOW3D_STAGE_COPY 7     /* H2D/D2H or D2D staging copies */
#define FLOW3D_STAGE_COUNT 8
int flow3d_solver_set_profiling(flow3d_solver* s, int enable);
/* enable != 0: every compute call prints the reference's per-level line "Solve level %2d (%4d x%4d x%4d)"
 * (what OpticalFlowE prints when `silent` is false, src/optical_flow/optical_flow_e.cpp:270-271) as the
 * level is enqueued */
int flow3d_solver_set_verbose(flow3d_solver* s, int enable);
int flow3d_solver_stage_times(flow3d_solver* s, float ms[FLOW3D_STAGE_COUNT],
                              double units[FLOW3D_STAGE_COUNT], uint64_t launches[FLOW3D_STAGE_COUNT]);

/* optional per-level observer for tests: called (after a stream sync) with the level's flow */
typedef void (*flow3d_level_callback)(int level, const size_t dims[3], size_t ld,
                                      const float* dev_u, const float* dev_v, const float* dev_w,
                                      void* user);
int flow3d_solver_set_level_callback(flow3d_solver* s, flow3d_level_callback cb, void* user);

/* ---- convergence diagnostics (NOT part of the reference's path) ------------
 * The reference runs a fixed number of sweeps and never evaluates a residual
 * (src/cuda_operations/entire_data/cuda_operation_solve.cpp:194-257).  flow3d_update_norm measures
 * the last Jacobi update: out_dev[0] = sum over voxels and the three components of (a - b)^2,
 * out_dev[1] = max |a - b| (NaN if any), over the slab's [z_begin, z_end) (whole volume if NULL).
 * Deterministic (fixed-order fold of per-block partials; warp shuffles inside a block).
 * workspace: flow3d_update_norm_workspace_bytes() bytes of device memory. */
size_t flow3d_update_norm_workspace_bytes(void);
int flow3d_update_norm(const float* a0, const float* a1, const float* a2, const float* b0,
                       const float* b1, const float* b2, const size_t dims[3], size_t ld,
                       const flow3d_zslab* slab, double* out_dev, void* workspace, void* stream);
/* enable != 0: every compute call records, per level and outer iteration, the update norm between the
 * last two sweeps.  update_tolerance > 0 additionally stops a level's outer loop once the RMS update
 * (voxel units) drops below it -- an opt-in fast mode whose results differ from the reference's. */
int flow3d_solver_set_diagnostics(flow3d_solver* s, int enable, float update_tolerance);
/* records of the last compute call, levels coarsest first: *n_levels, outer_per_level[l] = outer
 * iterations run on level l, rms[k] / max_abs[k] for the k-th record overall (k < *n_records).
 * Arrays may be NULL; at most `capacity` entries are written to each. */
int flow3d_solver_diagnostics(flow3d_solver* s, size_t* n_levels, size_t* outer_per_level, double* rms,
                              double* max_abs, size_t capacity, size_t* n_records);

/* ---- self-test ------------------------------------------------------------------------------------
 * The sweep kernels divide with the branch-free fast path of div.rn.f32 (csrc/common.cuh: div_fast) and
 * fall back to the IEEE division for operands outside [2^-60, 2^60].  This runs n_pairs divisions on the
 * device both ways: mode 0 = random operands, mode 1 = every divisor mantissa (n_pairs = k * 2^23).  Modes 2 / 3
 * check the branch-free sqrt / reciprocal of the robust weights (sqrt_fast, rcp_fast) against __fsqrt_rn / __frcp_rn on
 * the float whose bit pattern is i, for i < n_pairs (2^32 = every float).
 * out[0] = quotients that differ from IEEE division (must be 0), out[1] = pairs sent to the fallback,
 * out[2] = pairs tested.  Synchronous. */
int flow3d_selftest_fast_div(uint64_t n_pairs, uint64_t seed, int mode, uint64_t out[3]);

/* ---- synthetic test volumes ---------------------------------
 * Analytic texture pair with a known rigid motion, generated on the device in double precision.
 * z0/nz select a z-slab (for sharded generation); out_* may be NULL.  truth_* receive the
 * ground-truth flow in the reference's convention (f1(x + flow) = f0(x)). */
int flow3d_synth_pair(size_t width, size_t height, size_t depth, size_t z0, size_t nz, size_t ld,
                      uint64_t seed, float* frame_0, float* frame_1, float* truth_u,
                      float* truth_v, float* truth_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOW3D_C_H_ */
