/* flow3d_mgpu_c.h -- C ABI of the z-sharded multi-GPU flow solve (libflow3d_b200_mgpu.so).
 *
 * One rank per GPU (one process per GPU under torchrun / MPI, or one thread per GPU inside one process);
 * the ranks of one solve share an NCCL communicator.  Large pyramid levels are split along z with
 * H = inner_iterations + 1 ghost planes per side: an outer iteration (phi + all inner Jacobi sweeps) then
 * runs without communication on shrinking ranges, followed by ONE neighbour exchange of H planes of
 * (du,dv,dw) over NVLink (ncclSend/ncclRecv in one group on the solve's stream: no host round trip).
 * The solver is point-Jacobi, so the sharded flow is bit-identical to the single-GPU flow.
 *
 * What this replaces in the reference: the z-slab driver of its out-of-core path,
 * src/cuda_operations/partial_data/cuda_operation_solve_p.cpp:358-417 (slabs + one-plane halo, one GPU),
 * and the role of src/optical_flow/optical_flow_p.cpp:58-323 (the solver class for volumes larger than
 * one device).  The arithmetic is the single-GPU library's (include/flow3d_c.h, *_slab entry points).
 *
 * This library links NCCL; libflow3d_b200.so (single GPU) does not.
 */
#ifndef FLOW3D_MGPU_C_H_
#define FLOW3D_MGPU_C_H_

#include "flow3d_c.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flow3d_sharded flow3d_sharded;

#define FLOW3D_MGPU_ID_BYTES 128

/* ncclGetUniqueId: rank 0 calls it and hands the 128 bytes to every rank (torch.distributed broadcast,
 * MPI_Bcast, a file, or shared memory between threads) */
int flow3d_mgpu_unique_id(void* id128);

/* Owned planes [*a, *b) of rank `rank` for a level of depth d.  Balanced by COST, not by plane count: an
 * interior rank recomputes ghost planes on two sides, an edge rank on one, so edge ranks own a few
 * planes more (same function on every rank; any partition gives the same flow). */
void flow3d_sharded_own_range(size_t d, int rank, int world, size_t* a, size_t* b);

/* planes [*lo, *hi) of the RAW full-resolution frames rank `rank` must be given: its own planes plus
 * frame_ghost planes (data-dependent reach of the warp, coarse-level source intervals) plus the blur
 * radius (size_t)(3*sigma) */
void flow3d_sharded_input_planes(size_t depth, int rank, int world, float sigma, size_t frame_ghost,
                                 size_t* lo, size_t* hi);

/* Joins the communicator (collective over the `world` ranks: every rank must call it).  world == 1 is
 * allowed (no NCCL traffic). */
int flow3d_sharded_create(size_t width, size_t height, size_t depth, int device, int rank, int world,
                          const void* id128, flow3d_sharded** out);
int flow3d_sharded_destroy(flow3d_sharded* s);

/* One coarse-to-fine solve (collective).  raw_0 / raw_1: this rank's z-slab of the two raw frames on the
 * device, raw_planes planes starting at global plane raw_z0, pitch ld (must cover
 * flow3d_sharded_input_planes).  On return (work enqueued on `stream`; the call synchronises the stream
 * once per level to read the warp reach) flow_u/v/w hold this rank's owned planes [*out_a, *out_b) of the
 * finest level, pitch ld, at most out_capacity_planes planes -- all planes [0, D) on every rank if the
 * finest level was too small to shard. */
int flow3d_sharded_compute(flow3d_sharded* s, const float* raw_0, const float* raw_1, size_t raw_z0,
                           size_t raw_planes, size_t ld, const flow3d_params* params, size_t frame_ghost,
                           float* flow_u, float* flow_v, float* flow_w, size_t out_capacity_planes,
                           size_t* out_a, size_t* out_b, void* stream);

/* A level is sharded when every rank gets at least max(min_planes_per_rank, 2H) + 3 planes and
 * min_voxels_per_rank voxels (defaults 12 and 2^18; thinner levels are computed by every rank).  Must be
 * set identically on every rank.  Environment overrides at create time: FLOW3D_MGPU_MIN_PLANES,
 * FLOW3D_MGPU_MIN_VOXELS. */
int flow3d_sharded_set_thresholds(flow3d_sharded* s, size_t min_planes_per_rank, size_t min_voxels_per_rank);

/* Host-only dry run (no device, no communicator) of the partition arithmetic flow3d_sharded_compute relies on,
 * for every level and every rank of a `world`-rank solve: prolongation sources inside the previous level's
 * valid planes, matching ghost-exchange sizes between neighbours, every rank's piece of an all-gathered level
 * frame inside its own frame slab (own planes +- frame_ghost).  FLOW3D_OK, or FLOW3D_ERR_INVALID_ARG with the
 * first offending level / rank (either may be NULL).  flow3d_sharded_compute runs the same check before it
 * touches the device, so an unsupported geometry fails identically on every rank instead of leaving the others
 * inside a collective.  What it can refuse: a frame ghost shorter than the coarsest level's source interval
 * (use flow3d_sharded_frame_ghost), and parameter sets whose ghost depth inner_iterations_count + 1 does not
 * cover the median's reach plus the prolongation's source margin -- inner_iterations_count >= median_radius / 2 + 1
 * always passes (the defaults: 5 >= 3). */
int flow3d_sharded_plan_check(size_t width, size_t height, size_t depth, int world, const flow3d_params* params,
                              size_t min_planes_per_rank, size_t min_voxels_per_rank, size_t frame_ghost,
                              int* bad_level, int* bad_rank);

/* smallest frame_ghost (32, 64, ... up to the depth) for which flow3d_sharded_plan_check accepts the solve; 0 if
 * none does.  32 planes cover the default pyramid (scale 0.95: 40 levels need 8, 60 levels 22); steeper or
 * deeper pyramids need more (scale 0.9 x 40 levels: 61).  flow3d_mgpu_compute_host uses it. */
size_t flow3d_sharded_frame_ghost(size_t width, size_t height, size_t depth, int world, const flow3d_params* params,
                                  size_t min_planes_per_rank, size_t min_voxels_per_rank);

/* planes [*a, *b) flow3d_sharded_compute will deliver on this rank for these parameters (the owned range
 * of the finest level, or [0, D) when that level is too small to shard) */
int flow3d_sharded_output_planes(const flow3d_sharded* s, const flow3d_params* params, size_t* a, size_t* b);

/* per-phase device time of the last compute call (CUDA events on the solve's stream), in ms */
#define FLOW3D_MGPU_PHASE_PROLONGATION 0
#define FLOW3D_MGPU_PHASE_FLOW_EXCHANGE 1
#define FLOW3D_MGPU_PHASE_LEVEL_FRAMES 2
#define FLOW3D_MGPU_PHASE_WARP 3
#define FLOW3D_MGPU_PHASE_SOLVER 4
#define FLOW3D_MGPU_PHASE_HALO_EXCHANGE 5
#define FLOW3D_MGPU_PHASE_UPDATE 6
#define FLOW3D_MGPU_PHASE_MEDIAN 7
#define FLOW3D_MGPU_PHASE_BLUR 8
#define FLOW3D_MGPU_PHASE_COUNT 9
int flow3d_sharded_set_profiling(flow3d_sharded* s, int enable);
int flow3d_sharded_phase_ms(flow3d_sharded* s, float ms[FLOW3D_MGPU_PHASE_COUNT]);
/* counters of the last compute call: [0] sharded levels, [1] replicated levels, [2] neighbour exchanges,
 * [3] bytes this rank sent in them, [4] level-frame all-gathers, [5] voxel-sweeps computed by this rank
 * (incl. ghost planes), [6] phi voxel-updates, [7] peak device bytes held by the solve */
int flow3d_sharded_stats(flow3d_sharded* s, double out[8]);
/* SYNCHRONOUS launch-shape tuning of every level's slab (see flow3d_c.h "launch shapes") */
int flow3d_sharded_tune(flow3d_sharded* s, const flow3d_params* params);

/* Whole-volume host call on n_devices GPUs of this process: one host thread and one rank per device,
 * tight W*H*D host volumes in and out (the reference's ComputeFlow contract,
 * src/optical_flow/optical_flow_e.cpp:132-601).  ms_out (may be NULL): max over ranks of the CUDA-event
 * time from before the uploads to after the downloads.  persistent != 0 keeps the ranks (communicator,
 * memory pool, tuned shapes) alive for the next call with the same dims/devices; call with
 * n_devices == 0 to release them. */
int flow3d_mgpu_compute_host(size_t width, size_t height, size_t depth, int n_devices, const int* devices,
                             const float* frame_0, const float* frame_1, const flow3d_params* params,
                             float* flow_u, float* flow_v, float* flow_w, float* ms_out, int persistent);

#ifdef __cplusplus
}
#endif
#endif /* FLOW3D_MGPU_C_H_ */
