// flow3d_cabi.cu -- the C ABI (include/flow3d_c.h): argument checking, the stage wrappers and the
// solver object that runs the coarse-to-fine loop of OpticalFlowE::ComputeFlow
// (reference: src/optical_flow/optical_flow_e.cpp:132-601) on compact per-level device buffers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace f3d {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

void note_cuda_error(cudaError_t e, const char* what) {
  g_last_error = std::string(what ? what : "cuda") + ": " + cudaGetErrorName(e) + " (" +
                 cudaGetErrorString(e) + ")";
}
static thread_local bool g_count_suppressed = false;
void suppress_launch_count(bool on) { g_count_suppressed = on; }
void count_launch(unsigned n) {
  if (!g_count_suppressed) g_launches.fetch_add(n, std::memory_order_relaxed);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    note_cuda_error(e, what);
    return FLOW3D_ERR_CUDA;
  }
  return FLOW3D_OK;
}

int sm_count() {  // per device (a process may drive several GPUs)
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;
    }
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static bool fuse_ksi() {
  static const int on = [] {
    const char* e = getenv("FLOW3D_FUSE_KSI");
    return (e && *e) ? atoi(e) : 1;
  }();
  return on != 0;
}

// Stage timer: mark(tag) records an event; the time until the next mark belongs to `tag`.
struct StageTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  std::vector<int> tags;
  size_t used = 0;
  float ms[FLOW3D_STAGE_COUNT] = {0};
  double units[FLOW3D_STAGE_COUNT] = {0};
  uint64_t launches[FLOW3D_STAGE_COUNT] = {0};
  uint64_t launch_mark = 0;
  int cur_tag = -1;
  void reset() {
    used = 0;
    tags.clear();
    cur_tag = -1;
    for (int i = 0; i < FLOW3D_STAGE_COUNT; ++i) { ms[i] = 0; units[i] = 0; launches[i] = 0; }
  }
  void mark(int tag, cudaStream_t st, double work = 0.0) {
    if (!enabled) return;
    const uint64_t now = g_launches.load(std::memory_order_relaxed);
    if (cur_tag >= 0) launches[cur_tag] += now - launch_mark;
    launch_mark = now;
    cur_tag = tag;
    if (used == pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { enabled = false; return; }
      pool.push_back(e);
    }
    cudaEventRecord(pool[used++], st);
    tags.push_back(tag);
    if (tag >= 0) units[tag] += work;
  }
  // call after the stream has been synchronized; the last mark must be an end marker (tag -1)
  void finish() {
    if (!enabled) return;
    for (size_t i = 0; i + 1 < used; ++i) {
      if (tags[i] < 0) continue;
      float t = 0.f;
      if (cudaEventElapsedTime(&t, pool[i], pool[i + 1]) == cudaSuccess) ms[tags[i]] += t;
    }
  }
  void destroy() {
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pool.clear();
  }
};

#define F3D_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) {                             \
      f3d::note_cuda_error(e_, #call);                   \
      return e_ == cudaErrorMemoryAllocation ? FLOW3D_ERR_OUT_OF_MEMORY : FLOW3D_ERR_CUDA; \
    }                                                    \
  } while (0)

#define F3D_TRY(expr)                  \
  do {                                 \
    int rc_ = (expr);                  \
    if (rc_ != FLOW3D_OK) return rc_;  \
  } while (0)

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Gaussian taps exactly as cuda_operation_convolution.cpp:85-108 (precision 3, pixel size 1.0):
// double-precision evaluation, float storage, float running sum, float division.
static int gauss_taps(float sigma, float* taps, int max_radius) {
  const size_t precision = 3;
  const float pixel_size = 1.0f;
  const size_t radius = (size_t)(precision * sigma / pixel_size);
  if ((int)radius > max_radius) return -1;
  const int r = (int)radius;
  for (int i = -r; i <= r; i++) {
    float val = 1.0 / (sigma * std::sqrt(2.0 * 3.1415926)) *
                std::exp(-(i * i * pixel_size * pixel_size) / (2.0 * sigma * sigma));
    taps[i + r] = val;
  }
  float sum = 0.0;
  for (int i = 0; i < 2 * r + 1; i++) sum = sum + taps[i];
  for (int i = 0; i < 2 * r + 1; i++) taps[i] = taps[i] / sum;
  return r;
}

// zr: planes whose z pass is computed (x and y passes run on every local plane: the z pass reads
// `radius` planes around its range)
static int gauss_blur(const float* in, float* out, float* tmp, Dims g, float sigma, cudaStream_t st,
                      const ZRange* zr = nullptr) {
  float taps[65];
  const int r = gauss_taps(sigma, taps, 32);
  if (r < 0) return FLOW3D_ERR_UNSUPPORTED;
  const ZRange all{0, g.d};
  F3D_TRY(launch_conv_axis(in, out, g, taps, r, 0, all, st));   // rows    (cuda_operation_convolution.cpp:170)
  F3D_TRY(launch_conv_axis(out, tmp, g, taps, r, 1, all, st));  // columns (:174)
  F3D_TRY(launch_conv_axis(tmp, out, g, taps, r, 2, zr ? *zr : all, st));  // slices  (:178)
  return FLOW3D_OK;
}

// rows start on 128 B boundaries: full-sector warp accesses at every level width, 16 B-aligned TMA strides
static inline size_t aligned_ld(size_t w) { return (w + 31) & ~(size_t)31; }

// X -> Y -> Z (cuda_operation_resample.cpp:95-105) through two scratch volumes.  With slabs, id[2] /
// od[2] are local depths; x and y passes run on every local input plane, the z pass on the output range.
static int resample(const float* in, const size_t id[3], size_t in_ld, float* out, const size_t od[3],
                    size_t out_ld, float* ta, float* tb, cudaStream_t st,
                    const flow3d_zslab* in_slab = nullptr, const flow3d_zslab* out_slab = nullptr) {
  size_t d0[3] = {id[0], id[1], id[2]};
  size_t d1[3] = {od[0], id[1], id[2]};
  size_t d2[3] = {od[0], od[1], id[2]};
  const size_t l1 = aligned_ld(od[0]);
  const Dims g0 = make_slab_dims(d0, in_ld, in_slab), g1 = make_slab_dims(d1, l1, in_slab),
             g2 = make_slab_dims(d2, l1, in_slab), g3 = make_slab_dims(od, out_ld, out_slab);
  const ZRange all_in{0, (int)id[2]};
  F3D_TRY(launch_resample_axis(in, g0, ta, g1, 0, all_in, st));
  F3D_TRY(launch_resample_axis(ta, g1, tb, g2, 1, all_in, st));
  F3D_TRY(launch_resample_axis(tb, g2, out, g3, 2, make_range(g3, out_slab), st));
  return FLOW3D_OK;
}

// Opt-in convergence diagnostics (kernels_diag.cu): one {sum |update|^2, max |update|} record per
// outer iteration, measured between the last two Jacobi iterates.  With tol > 0 the level stops as
// soon as the RMS update falls below tol (NOT the reference's behaviour: it always runs `outer`
// iterations; results then differ from the parity path).
struct Diag {
  double* dev = nullptr;    // 2 doubles per record
  double* host = nullptr;   // pinned mirror, same layout
  void* workspace = nullptr;
  size_t capacity = 0;      // records
  size_t used = 0;
  float tol = 0.f;
};

static int solve_level(const float* fx, const float* fy, const float* fz, const float* ft,
                       const float* u, const float* v, const float* w, float* du, float* dv,
                       float* dw, float* phi, float* ksi, float* tdu, float* tdv, float* tdw, Dims g,
                       const float h[3], size_t outer, size_t inner, float alpha, float eps_s,
                       float eps_d, cudaStream_t st, StageTimer* tm = nullptr, Diag* dg = nullptr,
                       size_t* outer_done = nullptr) {
  const size_t bytes = (size_t)g.ps * g.d * sizeof(float);
  const double nvox = (double)g.w * g.h * g.d;
  // the phi / sweep launches below form one dependent chain on `st`: launched as programmatic dependents
  // (common.cuh; FLOW3D_PDL=0 turns it off), which hides the launch gap that is a visible share of the 5-150 us
  // kernels of the small levels; the large levels (gap < 1 % of a launch) keep the ordinary launches
  static const double pdl_max_voxels = [] {
    const char* e = getenv("FLOW3D_PDL_MAX_VOXELS");
    return (e && *e) ? atof(e) : 33554432.0;  // 2^25 ~ 322^3
  }();
  PdlScope pdl_chain(nvox <= pdl_max_voxels);
  if (tm) tm->mark(FLOW3D_STAGE_UPDATE, st);
  // cuda_operation_solve.cpp:183-188
  F3D_CUDA(cudaMemsetAsync(du, 0, bytes, st));
  F3D_CUDA(cudaMemsetAsync(dv, 0, bytes, st));
  F3D_CUDA(cudaMemsetAsync(dw, 0, bytes, st));
  float *a0 = du, *a1 = dv, *a2 = dw, *b0 = tdu, *b1 = tdv, *b2 = tdw;
  for (size_t i = 0; i < outer; ++i) {  // :194-257
    // The robust weights of compute_phi_ksi_3d are split: phi (needs neighbours) in its own launch, ksi
    // (pointwise in the iterate) inside the first sweep, which reads every operand of it anyway --
    // same operations, five fewer words of traffic per voxel and outer iteration.
    const bool fuse = fuse_ksi() && inner > 0;
    if (tm) tm->mark(FLOW3D_STAGE_PHI_KSI, st, nvox);
    F3D_TRY(launch_phi_ksi(fx, fy, fz, ft, u, v, w, a0, a1, a2, g, ZRange{0, g.d}, h[0], h[1], h[2], eps_s,
                           eps_d, phi, fuse ? nullptr : ksi, st));
    if (tm) tm->mark(FLOW3D_STAGE_SWEEP, st, nvox * (double)inner);
    for (size_t j = 0; j < inner; ++j) {
      F3D_TRY(launch_sweep(fx, fy, fz, ft, u, v, w, a0, a1, a2, phi, ksi, g, ZRange{0, g.d}, h[0], h[1], h[2],
                           alpha, b0, b1, b2, st, (fuse && j == 0) ? ksi : nullptr, eps_d));
      std::swap(a0, b0);
      std::swap(a1, b1);
      std::swap(a2, b2);
    }
    if (outer_done) *outer_done = i + 1;
    if (dg && dg->dev && dg->used < dg->capacity && inner > 0) {
      if (tm) tm->mark(FLOW3D_STAGE_UPDATE, st);
      double* rec = dg->dev + 2 * dg->used;
      F3D_TRY(launch_update_norm(a0, a1, a2, b0, b1, b2, g, ZRange{0, g.d}, rec, dg->workspace, st));
      if (dg->tol > 0.f) {  // adaptive stopping: the host has to see the value (one sync per outer iteration)
        F3D_CUDA(cudaMemcpyAsync(dg->host + 2 * dg->used, rec, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        F3D_CUDA(cudaStreamSynchronize(st));
        const double rms = std::sqrt(dg->host[2 * dg->used] / (3.0 * nvox));
        ++dg->used;
        if (rms < (double)dg->tol) break;
      } else {
        ++dg->used;
      }
    }
  }
  if (tm) tm->mark(FLOW3D_STAGE_UPDATE, st);
  if (a0 != du) {  // odd number of sweeps: bring the iterate home
    F3D_CUDA(cudaMemcpyAsync(du, a0, bytes, cudaMemcpyDeviceToDevice, st));
    F3D_CUDA(cudaMemcpyAsync(dv, a1, bytes, cudaMemcpyDeviceToDevice, st));
    F3D_CUDA(cudaMemcpyAsync(dw, a2, bytes, cudaMemcpyDeviceToDevice, st));
  }
  return FLOW3D_OK;
}

}  // namespace f3d

using namespace f3d;

// ================================================================================================
// solver object
// ================================================================================================
struct flow3d_solver {
  size_t W = 0, H = 0, D = 0;
  size_t ld = 0;        // full-resolution pitch
  size_t vol = 0;       // floats per arena volume = ld*H*D
  int device = 0;
  float* arena = nullptr;
  // 17 full-resolution volumes (the reference keeps 15 containers, optical_flow_e.h:40, but stores neither
  // the precomputed image derivatives nor a second ping-pong set): blurred frames 2, fx..ft 4, u,v,w 3,
  // du,dv,dw 3, ping-pong 3 (also resample / median scratch), phi + ksi 2 -- the level frames share the
  // phi / ksi slots: they are consumed by the warp before the solver writes its first weight.
  static constexpr int kVolumes = 17;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_ms[2] = {0.f, 0.f};
  flow3d_level_callback cb = nullptr;
  void* cb_user = nullptr;
  StageTimer timer;
  // convergence diagnostics (off by default)
  bool diag_enabled = false;
  Diag diag;
  std::vector<size_t> diag_level_outer;  // outer iterations run per level, coarsest first
  std::vector<double> diag_records;      // {sum_sq, max_abs} per outer iteration, levels concatenated
  std::vector<double> diag_level_voxels;
  bool verbose = false;  // print the reference's per-level line (optical_flow_e.cpp:270-271)
  // launch shapes tuned for this (scale factor, level count)?  (flow3d_solver_tune)
  float tuned_scale = -1.f;
  size_t tuned_levels = 0;
  bool tuned_full = false;
  float* buf(int i) const { return arena + (size_t)i * vol; }
};

// arena slots
enum { B_F0 = 0, B_F1, B_FX, B_FY, B_FZ, B_FT, B_U, B_V, B_W, B_DU, B_DV, B_DW, B_TDU, B_TDV, B_TDW, B_PHI, B_KSI,
       B_F0L = B_PHI, B_F1L = B_KSI };

// The coarse-to-fine loop (optical_flow_e.cpp:179-473).  in0/in1: device frames with pitch in_ld.
// On return u/v/w slots (tracked through the swaps) hold the full-resolution flow.
static int run_pyramid(flow3d_solver* s, const float* in0, const float* in1, size_t in_ld,
                       const flow3d_params* p, float** out_u, float** out_v, float** out_w,
                       cudaStream_t st) {
  const size_t full[3] = {s->W, s->H, s->D};
  const Dims gfull = make_dims(full, s->ld);
  float *F0 = s->buf(B_F0), *F1 = s->buf(B_F1), *f0l = s->buf(B_F0L), *f1l = s->buf(B_F1L);
  float *fx = s->buf(B_FX), *fy = s->buf(B_FY), *fz = s->buf(B_FZ), *ft = s->buf(B_FT);
  float *u = s->buf(B_U), *v = s->buf(B_V), *w = s->buf(B_W);
  float *du = s->buf(B_DU), *dv = s->buf(B_DV), *dw = s->buf(B_DW);
  float *tdu = s->buf(B_TDU), *tdv = s->buf(B_TDV), *tdw = s->buf(B_TDW);
  float *phi = s->buf(B_PHI), *ksi = s->buf(B_KSI);

  const size_t max_level = flow3d_max_warp_level(s->W, s->H, s->D, p->warp_scale_factor);
  int level = (int)std::min(p->warp_levels_count, max_level) - 1;  // :180

  StageTimer* tm = &s->timer;
  const double nfull = (double)s->W * s->H * s->D;
  // :213-257 pre-blur (or plain copy into the solver's own frames)
  tm->mark(FLOW3D_STAGE_BLUR, st, 2.0 * nfull);
  if (p->gaussian_sigma > 0.0) {
    if (in_ld != s->ld) {  // bring to the solver pitch first
      F3D_CUDA(cudaMemcpy2DAsync(f0l, s->ld * 4, in0, in_ld * 4, s->W * 4, s->H * s->D, cudaMemcpyDeviceToDevice, st));
      F3D_CUDA(cudaMemcpy2DAsync(f1l, s->ld * 4, in1, in_ld * 4, s->W * 4, s->H * s->D, cudaMemcpyDeviceToDevice, st));
      in0 = f0l;
      in1 = f1l;
    }
    F3D_TRY(gauss_blur(in0, F0, tdu, gfull, p->gaussian_sigma, st));
    F3D_TRY(gauss_blur(in1, F1, tdu, gfull, p->gaussian_sigma, st));
  } else {
    F3D_CUDA(cudaMemcpy2DAsync(F0, s->ld * 4, in0, in_ld * 4, s->W * 4, s->H * s->D, cudaMemcpyDeviceToDevice, st));
    F3D_CUDA(cudaMemcpy2DAsync(F1, s->ld * 4, in1, in_ld * 4, s->W * 4, s->H * s->D, cudaMemcpyDeviceToDevice, st));
  }

  Diag* dg = nullptr;
  s->diag_level_outer.clear();
  s->diag_level_voxels.clear();
  s->diag_records.clear();
  if (s->diag_enabled) {
    const size_t need = (size_t)(level + 1) * p->outer_iterations_count;
    if (need > s->diag.capacity) {
      if (s->diag.dev) cudaFree(s->diag.dev);
      if (s->diag.host) cudaFreeHost(s->diag.host);
      s->diag.dev = nullptr; s->diag.host = nullptr; s->diag.capacity = 0;
      F3D_CUDA(cudaMalloc(&s->diag.dev, need * 2 * sizeof(double)));
      F3D_CUDA(cudaMallocHost(&s->diag.host, need * 2 * sizeof(double)));
      s->diag.capacity = need;
    }
    if (!s->diag.workspace) F3D_CUDA(cudaMalloc(&s->diag.workspace, update_norm_workspace_bytes()));
    s->diag.used = 0;
    dg = &s->diag;
  }

  size_t prev[3] = {0, 0, 0};
  size_t prev_ld = 0;
  while (level >= 0) {  // :261
    size_t cur[3];
    float h[3];
    flow3d_level_geometry(s->W, s->H, s->D, p->warp_scale_factor, level, cur, h);
    const size_t ld = aligned_ld(cur[0]);
    const Dims g = make_dims(cur, ld);
    const size_t bytes = (size_t)g.ps * g.d * sizeof(float);
    if (s->verbose)  // the reference's own line, printed as the level is enqueued (the device runs behind)
      std::printf("Solve level %2d (%4d x%4d x%4d) \n", level, (int)cur[0], (int)cur[1], (int)cur[2]);

    const double nvox = (double)cur[0] * cur[1] * cur[2];
    tm->mark(FLOW3D_STAGE_RESAMPLE, st, 5.0 * nvox);
    const float *pf0, *pf1;
    if (level == 0) {  // :275-277
      pf0 = F0;
      pf1 = F1;
    } else {  // :279-299, always from full resolution
      F3D_TRY(resample(F0, full, s->ld, f0l, cur, ld, tdu, tdv, st));
      F3D_TRY(resample(F1, full, s->ld, f1l, cur, ld, tdu, tdv, st));
      pf0 = f0l;
      pf1 = f1l;
    }
    if (prev[0] == 0) {  // :304-310
      F3D_CUDA(cudaMemsetAsync(u, 0, bytes, st));
      F3D_CUDA(cudaMemsetAsync(v, 0, bytes, st));
      F3D_CUDA(cudaMemsetAsync(w, 0, bytes, st));
    } else {  // :311-344 prolongation (values not rescaled)
      F3D_TRY(resample(u, prev, prev_ld, du, cur, ld, tdu, tdv, st));
      F3D_TRY(resample(v, prev, prev_ld, dv, cur, ld, tdu, tdv, st));
      F3D_TRY(resample(w, prev, prev_ld, dw, cur, ld, tdu, tdv, st));
      std::swap(u, du);
      std::swap(v, dv);
      std::swap(w, dw);
    }
    // :348-369 warp, fused with the derivative stencils the solver kernels would recompute
    tm->mark(FLOW3D_STAGE_WARP, st, nvox);
    F3D_TRY(launch_warp_derivatives(pf0, pf1, 0, g.d, u, v, w, g, ZRange{0, g.d}, h[0], h[1], h[2], fx, fy, fz,
                                    ft, st));
    // :372-417
    size_t outer_done = 0;
    F3D_TRY(solve_level(fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, tdu, tdv, tdw, g, h,
                        p->outer_iterations_count, p->inner_iterations_count, p->equation_alpha,
                        p->equation_smoothness, p->equation_data, st, tm, dg, &outer_done));
    if (dg) {
      s->diag_level_outer.push_back(p->inner_iterations_count > 0 ? outer_done : 0);
      s->diag_level_voxels.push_back(nvox);
    }
    // :420-438
    tm->mark(FLOW3D_STAGE_UPDATE, st, nvox);
    F3D_TRY(launch_add3(u, v, w, du, dv, dw, g, st));
    prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
    prev_ld = ld;
    --level;
    // :443-473 median of each component (through a temp, then swap)
    tm->mark(FLOW3D_STAGE_MEDIAN, st, 3.0 * nvox);
    {
      size_t r = p->median_radius;
      const ZRange all{0, g.d};
      int rc = launch_median(u, tdu, g, all, (int)r, st);
      if (rc == FLOW3D_OK) {
        std::swap(u, tdu);
        F3D_TRY(launch_median(v, tdu, g, all, (int)r, st));
        std::swap(v, tdu);
        F3D_TRY(launch_median(w, tdu, g, all, (int)r, st));
        std::swap(w, tdu);
      } else if (rc != FLOW3D_ERR_UNSUPPORTED) {
        return rc;
      } else {
        return rc;  // unsupported radius: report instead of handing back an unfiltered/garbage flow
      }
    }
    tm->mark(-1, st);
    if (s->cb) {
      F3D_CUDA(cudaStreamSynchronize(st));
      s->cb(level + 1, cur, ld, u, v, w, s->cb_user);
    }
  }
  if (dg && dg->used) {  // bring the records home (after the last level; the caller's sync covers it)
    F3D_CUDA(cudaMemcpyAsync(dg->host, dg->dev, dg->used * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  *out_u = u;
  *out_v = v;
  *out_w = w;
  return FLOW3D_OK;
}

static int check_params(const flow3d_params* p) {
  if (!p) return FLOW3D_ERR_INVALID_ARG;
  if (p->warp_levels_count == 0) return FLOW3D_ERR_INVALID_ARG;
  if (!(p->warp_scale_factor > 0.f)) return FLOW3D_ERR_INVALID_ARG;
  size_t r = p->median_radius;
  if (r == 0) return FLOW3D_ERR_INVALID_ARG;
  if (r != 1 && (r % 2 == 0)) r -= 1;
  if (!(r == 1 || r == 3 || r == 5 || r == 7)) return FLOW3D_ERR_UNSUPPORTED;
  if (p->gaussian_sigma > 0.f && (size_t)(3 * p->gaussian_sigma) > 32) return FLOW3D_ERR_UNSUPPORTED;
  return FLOW3D_OK;
}

extern "C" {

int flow3d_version(void) { return 100; }

const char* flow3d_status_string(int status) {
  switch (status) {
    case FLOW3D_OK: return "ok";
    case FLOW3D_ERR_INVALID_ARG: return "invalid argument";
    case FLOW3D_ERR_UNSUPPORTED: return "unsupported parameter value";
    case FLOW3D_ERR_CUDA: return "CUDA error";
    case FLOW3D_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case FLOW3D_ERR_OUT_OF_MEMORY: return "out of device memory";
    case FLOW3D_ERR_NOT_INITIALIZED: return "solver not initialized";
    default: return "unknown status";
  }
}

const char* flow3d_last_cuda_error(void) { return g_last_error.c_str(); }

int flow3d_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    note_cuda_error(e, "cudaGetDeviceCount");
    cudaGetLastError();
    return FLOW3D_ERR_NO_DEVICE;
  }
  return n;
}

void flow3d_default_params(flow3d_params* p) {  // src/main.cpp:77-85
  if (!p) return;
  p->warp_levels_count = 40;
  p->warp_scale_factor = 0.95f;
  p->outer_iterations_count = 40;
  p->inner_iterations_count = 5;
  p->equation_alpha = 7.5f;
  p->equation_smoothness = 0.001f;
  p->equation_data = 0.001f;
  p->median_radius = 5;
  p->gaussian_sigma = 2.0f;
}

uint64_t flow3d_launch_count(void) { return g_launches.load(); }
void flow3d_reset_launch_count(void) { g_launches.store(0); }

size_t flow3d_max_warp_level(size_t width, size_t height, size_t depth, float scale_factor) {
  size_t r_width = 1, r_height = 1, r_depth = 1;
  size_t level_counter = 1;
  while (scale_factor < 1.f) {
    const float scale = std::pow(scale_factor, static_cast<float>(level_counter));
    r_width = static_cast<size_t>(std::ceil(width * scale));
    r_height = static_cast<size_t>(std::ceil(height * scale));
    r_depth = static_cast<size_t>(std::ceil(depth * scale));
    if (r_width < 4 || r_height < 4 || r_depth < 4) break;
    ++level_counter;
  }
  if (r_width == 1 || r_height == 1 || r_depth == 1) --level_counter;
  return level_counter;
}

int flow3d_level_geometry(size_t width, size_t height, size_t depth, float scale_factor, int level,
                          size_t dims[3], float h[3]) {
  if (!dims || !h || level < 0) return FLOW3D_ERR_INVALID_ARG;
  const float scale = std::pow(scale_factor, static_cast<float>(level));
  dims[0] = static_cast<size_t>(std::ceil(width * scale));
  dims[1] = static_cast<size_t>(std::ceil(height * scale));
  dims[2] = static_cast<size_t>(std::ceil(depth * scale));
  if (dims[0] == 0 || dims[1] == 0 || dims[2] == 0) return FLOW3D_ERR_INVALID_ARG;
  h[0] = width / static_cast<float>(dims[0]);
  h[1] = height / static_cast<float>(dims[1]);
  h[2] = depth / static_cast<float>(dims[2]);
  return FLOW3D_OK;
}

size_t flow3d_aligned_ld(size_t w) { return aligned_ld(w); }

int flow3d_set_device(int device) {
  F3D_CUDA(cudaSetDevice(device));
  return FLOW3D_OK;
}
int flow3d_malloc(void** dev_ptr, size_t bytes) {
  if (!dev_ptr) return FLOW3D_ERR_INVALID_ARG;
  F3D_CUDA(cudaMalloc(dev_ptr, bytes));
  return FLOW3D_OK;
}
int flow3d_free(void* dev_ptr) {
  F3D_CUDA(cudaFree(dev_ptr));
  return FLOW3D_OK;
}
int flow3d_memset(void* dev_ptr, int value, size_t bytes, void* stream) {
  F3D_CUDA(cudaMemsetAsync(dev_ptr, value, bytes, S(stream)));
  return FLOW3D_OK;
}
int flow3d_upload(const float* host, float* dev, const size_t dims[3], size_t ld, void* stream) {
  if (!host) return FLOW3D_ERR_INVALID_ARG;
  F3D_TRY(check_volume(dev, dims, ld));
  F3D_CUDA(cudaMemcpy2DAsync(dev, ld * 4, host, dims[0] * 4, dims[0] * 4, dims[1] * dims[2],
                             cudaMemcpyHostToDevice, S(stream)));
  return FLOW3D_OK;
}
int flow3d_download(const float* dev, float* host, const size_t dims[3], size_t ld, void* stream) {
  if (!host) return FLOW3D_ERR_INVALID_ARG;
  F3D_TRY(check_volume(dev, dims, ld));
  F3D_CUDA(cudaMemcpy2DAsync(host, dims[0] * 4, dev, ld * 4, dims[0] * 4, dims[1] * dims[2],
                             cudaMemcpyDeviceToHost, S(stream)));
  return FLOW3D_OK;
}
int flow3d_stream_synchronize(void* stream) {
  F3D_CUDA(cudaStreamSynchronize(S(stream)));
  return FLOW3D_OK;
}

int flow3d_host_alloc(void** host_ptr, size_t bytes) {
  if (!host_ptr) return FLOW3D_ERR_INVALID_ARG;
  *host_ptr = nullptr;
  if (flow3d_device_count() <= 0) return FLOW3D_ERR_NO_DEVICE;
  F3D_CUDA(cudaHostAlloc(host_ptr, bytes, cudaHostAllocDefault));
  return FLOW3D_OK;
}
int flow3d_host_free(void* host_ptr) {
  if (!host_ptr) return FLOW3D_OK;
  F3D_CUDA(cudaFreeHost(host_ptr));
  return FLOW3D_OK;
}
int flow3d_device_name(int device, char* buf, size_t n) {
  if (!buf || n == 0) return FLOW3D_ERR_INVALID_ARG;
  if (flow3d_device_count() <= 0) return FLOW3D_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  F3D_CUDA(cudaGetDeviceProperties(&prop, device));
  std::snprintf(buf, n, "%s", prop.name);
  return FLOW3D_OK;
}

// ---- stage wrappers ----------------------------------------------------------------------------
int flow3d_gauss_blur(const float* in, float* out, float* tmp, const size_t dims[3], size_t ld,
                      float sigma, void* stream) {
  F3D_TRY(check_volume(in, dims, ld));
  F3D_TRY(check_volume(out, dims, ld));
  F3D_TRY(check_volume(tmp, dims, ld));
  if (in == out || !(sigma > 0.f)) return FLOW3D_ERR_INVALID_ARG;
  return gauss_blur(in, out, tmp, make_dims(dims, ld), sigma, S(stream));
}

int flow3d_resample(const float* in, const size_t in_dims[3], size_t in_ld, float* out,
                    const size_t out_dims[3], size_t out_ld, float* tmp_a, float* tmp_b,
                    void* stream) {
  F3D_TRY(check_volume(in, in_dims, in_ld));
  F3D_TRY(check_volume(out, out_dims, out_ld));
  if (!tmp_a || !tmp_b || !aligned16(tmp_a) || !aligned16(tmp_b) || in == out) return FLOW3D_ERR_INVALID_ARG;
  return resample(in, in_dims, in_ld, out, out_dims, out_ld, tmp_a, tmp_b, S(stream));
}

int flow3d_warp(const float* f0, const float* f1, const float* u, const float* v, const float* w,
                const size_t dims[3], size_t ld, const float h[3], float* out, void* stream) {
  const void* ps[] = {f0, f1, u, v, w, out};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!h || out == f1) return FLOW3D_ERR_INVALID_ARG;
  return launch_warp(f0, f1, u, v, w, make_dims(dims, ld), h[0], h[1], h[2], out, S(stream));
}

int flow3d_derivatives(const float* f0, const float* f1w, const size_t dims[3], size_t ld,
                       const float h[3], float* fx, float* fy, float* fz, float* ft, void* stream) {
  const void* ps[] = {f0, f1w, fx, fy, fz, ft};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!h) return FLOW3D_ERR_INVALID_ARG;
  return launch_derivatives(f0, f1w, make_dims(dims, ld), h[0], h[1], h[2], fx, fy, fz, ft, S(stream));
}

int flow3d_warp_derivatives(const float* f0, const float* f1, const float* u, const float* v,
                            const float* w, const size_t dims[3], size_t ld, const float h[3],
                            float* fx, float* fy, float* fz, float* ft, void* stream) {
  const void* ps[] = {f0, f1, u, v, w, fx, fy, fz, ft};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!h) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_dims(dims, ld);
  return launch_warp_derivatives(f0, f1, 0, g.d, u, v, w, g, ZRange{0, g.d}, h[0], h[1], h[2], fx, fy, fz, ft,
                                 S(stream));
}

int flow3d_phi_ksi(const float* fx, const float* fy, const float* fz, const float* ft,
                   const float* u, const float* v, const float* w, const float* du,
                   const float* dv, const float* dw, const size_t dims[3], size_t ld,
                   const float h[3], float eps_smooth, float eps_data, float* phi, float* ksi,
                   void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!h) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_dims(dims, ld);
  return launch_phi_ksi(fx, fy, fz, ft, u, v, w, du, dv, dw, g, ZRange{0, g.d}, h[0], h[1], h[2], eps_smooth,
                        eps_data, phi, ksi, S(stream));
}

int flow3d_sweep(const float* fx, const float* fy, const float* fz, const float* ft,
                 const float* u, const float* v, const float* w, const float* du, const float* dv,
                 const float* dw, const float* phi, const float* ksi, const size_t dims[3],
                 size_t ld, const float h[3], float alpha, float* du_out, float* dv_out,
                 float* dw_out, void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, du_out, dv_out, dw_out};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!h || du_out == du || dv_out == dv || dw_out == dw) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_dims(dims, ld);
  return launch_sweep(fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, g, ZRange{0, g.d}, h[0], h[1], h[2], alpha,
                      du_out, dv_out, dw_out, S(stream));
}

int flow3d_solve_level(const float* fx, const float* fy, const float* fz, const float* ft,
                       const float* u, const float* v, const float* w, float* du, float* dv,
                       float* dw, float* scratch, const size_t dims[3], size_t ld,
                       const float h[3], size_t outer, size_t inner, float alpha, float eps_smooth,
                       float eps_data, void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, scratch};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!h) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_dims(dims, ld);
  const size_t n = (size_t)g.ps * g.d;
  return solve_level(fx, fy, fz, ft, u, v, w, du, dv, dw, scratch, scratch + n, scratch + 2 * n,
                     scratch + 3 * n, scratch + 4 * n, g, h, outer, inner, alpha, eps_smooth, eps_data,
                     S(stream));
}

int flow3d_add3(float* u, float* v, float* w, const float* du, const float* dv, const float* dw,
                const size_t dims[3], size_t ld, void* stream) {
  const void* ps[] = {u, v, w, du, dv, dw};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  return launch_add3(u, v, w, du, dv, dw, make_dims(dims, ld), S(stream));
}

int flow3d_median(const float* in, float* out, const size_t dims[3], size_t ld, size_t radius,
                  void* stream) {
  F3D_TRY(check_volume(in, dims, ld));
  F3D_TRY(check_volume(out, dims, ld));
  if (in == out || radius == 0 || radius > 64) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_dims(dims, ld);
  return launch_median(in, out, g, ZRange{0, g.d}, (int)radius, S(stream));
}

// ---- z-slab variants -------------------------------------------------------------------------------
int flow3d_gauss_blur_slab(const float* in, float* out, float* tmp, const size_t dims[3], size_t ld,
                           const flow3d_zslab* slab, float sigma, void* stream) {
  F3D_TRY(check_volume(in, dims, ld));
  F3D_TRY(check_volume(out, dims, ld));
  F3D_TRY(check_volume(tmp, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (in == out || !(sigma > 0.f) || !slab) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  const ZRange zr = make_range(g, slab);
  // every tap of the z pass that lies inside the level must be a plane of the buffer
  const int r = (int)(size_t)(3 * sigma);
  if ((g.z0g + zr.begin - r < g.z0g && g.z0g + zr.begin - r >= 0) ||
      (g.z0g + zr.end - 1 + r >= g.z0g + g.d && g.z0g + zr.end - 1 + r < g.dg))
    return FLOW3D_ERR_INVALID_ARG;
  return gauss_blur(in, out, tmp, g, sigma, S(stream), &zr);
}

int flow3d_sweep_slab(const float* fx, const float* fy, const float* fz, const float* ft,
                      const float* u, const float* v, const float* w, const float* du,
                      const float* dv, const float* dw, const float* phi, const float* ksi,
                      const size_t dims[3], size_t ld, const flow3d_zslab* slab, const float h[3],
                      float alpha, float* du_out, float* dv_out, float* dw_out, void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, du_out, dv_out, dw_out};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h || du_out == du || dv_out == dv || dw_out == dw) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  return launch_sweep(fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, g, make_range(g, slab), h[0], h[1], h[2],
                      alpha, du_out, dv_out, dw_out, S(stream));
}

int flow3d_sweep_shape(const float* fx, const float* fy, const float* fz, const float* ft, const float* u,
                       const float* v, const float* w, const float* du, const float* dv, const float* dw,
                       const float* phi, const float* ksi, const size_t dims[3], size_t ld,
                       const flow3d_zslab* slab, const float h[3], float alpha, float eps_data, float* du_out,
                       float* dv_out, float* dw_out, float* ksi_out, int variant, int vec, int nchunks,
                       void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, phi, du_out, dv_out, dw_out};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  F3D_TRY(check_volume(ksi_out ? ksi_out : ksi, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h || du_out == du || dv_out == dv || dw_out == dw) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  return launch_sweep_shape(fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, g, make_range(g, slab), h[0], h[1], h[2],
                            alpha, du_out, dv_out, dw_out, S(stream), ksi_out, eps_data, variant, vec, nchunks);
}

int flow3d_phi_ksi_slab(const float* fx, const float* fy, const float* fz, const float* ft,
                        const float* u, const float* v, const float* w, const float* du,
                        const float* dv, const float* dw, const size_t dims[3], size_t ld,
                        const flow3d_zslab* slab, const float h[3], float eps_smooth, float eps_data,
                        float* phi, float* ksi, void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  return launch_phi_ksi(fx, fy, fz, ft, u, v, w, du, dv, dw, g, make_range(g, slab), h[0], h[1], h[2],
                        eps_smooth, eps_data, phi, ksi, S(stream));
}

int flow3d_warp_derivatives_slab(const float* f0, const float* f1, size_t f1_z0_global,
                                 size_t f1_depth_local, const float* u, const float* v,
                                 const float* w, const size_t dims[3], size_t ld,
                                 const flow3d_zslab* slab, const float h[3], float* fx, float* fy,
                                 float* fz, float* ft, void* stream) {
  const void* ps[] = {f0, u, v, w, fx, fy, fz, ft};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h || !f1 || !aligned16(f1) || f1_depth_local == 0) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  if (f1_z0_global + f1_depth_local > (size_t)g.dg) return FLOW3D_ERR_INVALID_ARG;
  return launch_warp_derivatives(f0, f1, (int)f1_z0_global, (int)f1_depth_local, u, v, w, g, make_range(g, slab),
                                 h[0], h[1], h[2], fx, fy, fz, ft, S(stream));
}

// The outer iteration of a z-slab, optionally split for communication overlap.  part 0: everything (phi on
// [z_begin,z_end), sweep j on the range shrunk by j planes on every side that is not a global face).
// part 1 ("early"): only the planes whose values do not depend on the ghost planes outside the owned range
// [own_begin, own_end): step j (phi = 0, sweep j) on [own_begin+j+1, own_end-j-1), extended to the buffer
// end on a side without ghost planes -- a value at plane p after step j depends on the starting iterate on
// [p-j-1, p+j+1] only.  part 2 ("late"): the remaining planes of every step, to be run once the ghosts
// have arrived.  Early and late launches touch disjoint planes of every buffer at every step (the late
// part of step j reads planes <= own_begin+j+1 of the buffer the early part of step j+1 writes from
// own_begin+j+2 on), so early(i+1) may overlap the exchange that follows late(i).
static int outer_iteration_parts(const float* fx, const float* fy, const float* fz, const float* ft, const float* u,
                                 const float* v, const float* w, float* du, float* dv, float* dw, float* tdu,
                                 float* tdv, float* tdw, float* phi, float* ksi, const Dims& g, ZRange r0,
                                 const float* h, size_t inner, float alpha, float eps_smooth, float eps_data, int part,
                                 int own_begin, int own_end, int* result_in_tmp, cudaStream_t st) {
  const bool lo_face = (g.z0g + r0.begin) == 0, hi_face = (g.z0g + r0.end) == g.dg;
  // ksi is pointwise in the iterate, so the first sweep computes it (every later sweep of the iteration
  // works inside the first one's range)
  const bool fuse = fuse_ksi() && inner > 0;
  float *a0 = du, *a1 = dv, *a2 = dw, *b0 = tdu, *b1 = tdv, *b2 = tdw;
  for (size_t j = 0; j <= inner; ++j) {
    const ZRange full{(lo_face || j == 0) ? r0.begin : r0.begin + (int)j, (hi_face || j == 0) ? r0.end : r0.end - (int)j};
    // a side has ghosts when the buffer holds planes beyond the owned range there
    ZRange early = full;
    if (own_begin > 0) early.begin = std::max(full.begin, own_begin + (int)j + 1);
    if (own_end < g.d) early.end = std::min(full.end, own_end - (int)j - 1);
    if (early.end < early.begin) early.end = early.begin;
    ZRange todo[2];
    int n = 0;
    if (part == 0) todo[n++] = full;
    else if (part == 1) todo[n++] = early;
    else {
      todo[n++] = ZRange{full.begin, std::min(early.begin, full.end)};
      todo[n++] = ZRange{std::max(early.end, full.begin), full.end};
    }
    for (int k = 0; k < n; ++k) {
      if (todo[k].end <= todo[k].begin) continue;
      if (j == 0) {
        F3D_TRY(launch_phi_ksi(fx, fy, fz, ft, u, v, w, du, dv, dw, g, todo[k], h[0], h[1], h[2], eps_smooth, eps_data,
                               phi, fuse ? nullptr : ksi, st));
      } else {
        F3D_TRY(launch_sweep(fx, fy, fz, ft, u, v, w, a0, a1, a2, phi, ksi, g, todo[k], h[0], h[1], h[2], alpha, b0, b1,
                             b2, st, (fuse && j == 1) ? ksi : nullptr, eps_data));
      }
    }
    if (j >= 1) {
      std::swap(a0, b0);
      std::swap(a1, b1);
      std::swap(a2, b2);
    }
  }
  *result_in_tmp = (a0 == tdu) ? 1 : 0;
  return FLOW3D_OK;
}

int flow3d_outer_iteration_slab(const float* fx, const float* fy, const float* fz, const float* ft,
                                const float* u, const float* v, const float* w, float* du, float* dv,
                                float* dw, float* tdu, float* tdv, float* tdw, float* phi, float* ksi,
                                const size_t dims[3], size_t ld, const flow3d_zslab* slab,
                                const float h[3], size_t inner, float alpha, float eps_smooth,
                                float eps_data, int* result_in_tmp, void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, tdu, tdv, tdw, phi, ksi};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h || !slab || !result_in_tmp) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  return outer_iteration_parts(fx, fy, fz, ft, u, v, w, du, dv, dw, tdu, tdv, tdw, phi, ksi, g, make_range(g, slab), h,
                               inner, alpha, eps_smooth, eps_data, 0, 0, 0, result_in_tmp, S(stream));
}

int flow3d_outer_iteration_slab_part(const float* fx, const float* fy, const float* fz, const float* ft,
                                     const float* u, const float* v, const float* w, float* du, float* dv,
                                     float* dw, float* tdu, float* tdv, float* tdw, float* phi, float* ksi,
                                     const size_t dims[3], size_t ld, const flow3d_zslab* slab,
                                     const float h[3], size_t inner, float alpha, float eps_smooth,
                                     float eps_data, int part, size_t own_begin, size_t own_end,
                                     int* result_in_tmp, void* stream) {
  const void* ps[] = {fx, fy, fz, ft, u, v, w, du, dv, dw, tdu, tdv, tdw, phi, ksi};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h || !slab || !result_in_tmp || part < 0 || part > 2) return FLOW3D_ERR_INVALID_ARG;
  if (own_begin > own_end || own_end > dims[2]) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  return outer_iteration_parts(fx, fy, fz, ft, u, v, w, du, dv, dw, tdu, tdv, tdw, phi, ksi, g, make_range(g, slab), h,
                               inner, alpha, eps_smooth, eps_data, part, (int)own_begin, (int)own_end, result_in_tmp,
                               S(stream));
}

int flow3d_median_slab(const float* in, float* out, const size_t dims[3], size_t ld,
                       const flow3d_zslab* slab, size_t radius, void* stream) {
  F3D_TRY(check_volume(in, dims, ld));
  F3D_TRY(check_volume(out, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (in == out || radius == 0 || radius > 64) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  return launch_median(in, out, g, make_range(g, slab), (int)radius, S(stream));
}

int flow3d_resample_slab(const float* in, const size_t in_dims[3], size_t in_ld,
                         const flow3d_zslab* in_slab, float* out, const size_t out_dims[3],
                         size_t out_ld, const flow3d_zslab* out_slab, float* tmp_a, float* tmp_b,
                         void* stream) {
  F3D_TRY(check_volume(in, in_dims, in_ld));
  F3D_TRY(check_volume(out, out_dims, out_ld));
  F3D_TRY(check_slab(in_dims, in_slab));
  F3D_TRY(check_slab(out_dims, out_slab));
  if (!tmp_a || !tmp_b || !aligned16(tmp_a) || !aligned16(tmp_b) || in == out) return FLOW3D_ERR_INVALID_ARG;
  return resample(in, in_dims, in_ld, out, out_dims, out_ld, tmp_a, tmp_b, S(stream), in_slab, out_slab);
}

int flow3d_absmax(const float* dev, const size_t dims[3], size_t ld, float* out_dev, void* stream) {
  F3D_TRY(check_volume(dev, dims, ld));
  if (!out_dev) return FLOW3D_ERR_INVALID_ARG;
  return launch_absmax(dev, make_dims(dims, ld), out_dev, S(stream));
}

// ---- solver ----------------------------------------------------------------------------------------
size_t flow3d_solver_workspace_bytes(size_t width, size_t height, size_t depth) {
  return aligned_ld(width) * height * depth * sizeof(float) * (size_t)flow3d_solver::kVolumes;
}

int flow3d_solver_create(size_t width, size_t height, size_t depth, int device, flow3d_solver** out) {
  if (!out || width < 2 || height < 2 || depth < 2) return FLOW3D_ERR_INVALID_ARG;
  if (width > (1u << 30) || height > (1u << 30) || depth > (1u << 30)) return FLOW3D_ERR_INVALID_ARG;
  *out = nullptr;
  int n = flow3d_device_count();
  if (n <= 0) return FLOW3D_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return FLOW3D_ERR_INVALID_ARG;
  F3D_CUDA(cudaSetDevice(device));
  flow3d_solver* s = new flow3d_solver();
  s->W = width; s->H = height; s->D = depth;
  s->ld = aligned_ld(width);
  s->vol = s->ld * height * depth;
  s->device = device;
  cudaError_t e = cudaMalloc(&s->arena, s->vol * sizeof(float) * (size_t)flow3d_solver::kVolumes);
  if (e != cudaSuccess) {
    note_cuda_error(e, "cudaMalloc(arena)");
    cudaGetLastError();
    delete s;
    return FLOW3D_ERR_OUT_OF_MEMORY;
  }
  e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
  if (e != cudaSuccess) {
    note_cuda_error(e, "stream/event create");
    flow3d_solver_destroy(s);
    return FLOW3D_ERR_CUDA;
  }
  *out = s;
  return FLOW3D_OK;
}

int flow3d_solver_destroy(flow3d_solver* s) {
  if (!s) return FLOW3D_OK;
  cudaSetDevice(s->device);
  for (int i = 0; i < 4; ++i)
    if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  if (s->stream) cudaStreamDestroy(s->stream);
  s->timer.destroy();
  if (s->diag.dev) cudaFree(s->diag.dev);
  if (s->diag.host) cudaFreeHost(s->diag.host);
  if (s->diag.workspace) cudaFree(s->diag.workspace);
  if (s->arena) cudaFree(s->arena);
  delete s;
  return FLOW3D_OK;
}

int flow3d_solver_set_level_callback(flow3d_solver* s, flow3d_level_callback cb, void* user) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  s->cb = cb;
  s->cb_user = user;
  return FLOW3D_OK;
}

int flow3d_solver_last_timing(const flow3d_solver* s, float ms[2]) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!ms) return FLOW3D_ERR_INVALID_ARG;
  ms[0] = s->last_ms[0];
  ms[1] = s->last_ms[1];
  return FLOW3D_OK;
}

int flow3d_solver_compute_device(flow3d_solver* s, const float* frame_0, const float* frame_1,
                                 size_t ld, const flow3d_params* params, float* flow_u,
                                 float* flow_v, float* flow_w, void* stream) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  F3D_TRY(check_params(params));
  const size_t full[3] = {s->W, s->H, s->D};
  const void* ps[] = {frame_0, frame_1, flow_u, flow_v, flow_w};
  for (const void* p : ps) F3D_TRY(check_volume(p, full, ld));
  F3D_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = S(stream);
  s->timer.reset();
  F3D_CUDA(cudaEventRecord(s->ev[0], st));
  float *u, *v, *w;
  F3D_TRY(run_pyramid(s, frame_0, frame_1, ld, params, &u, &v, &w, st));
  const size_t wb = s->W * 4, rows = s->H * s->D;
  s->timer.mark(FLOW3D_STAGE_COPY, st);
  F3D_CUDA(cudaMemcpy2DAsync(flow_u, ld * 4, u, s->ld * 4, wb, rows, cudaMemcpyDeviceToDevice, st));
  F3D_CUDA(cudaMemcpy2DAsync(flow_v, ld * 4, v, s->ld * 4, wb, rows, cudaMemcpyDeviceToDevice, st));
  F3D_CUDA(cudaMemcpy2DAsync(flow_w, ld * 4, w, s->ld * 4, wb, rows, cudaMemcpyDeviceToDevice, st));
  s->timer.mark(-1, st);
  F3D_CUDA(cudaEventRecord(s->ev[1], st));
  return FLOW3D_OK;
}

static int solver_tune(flow3d_solver* s, const flow3d_params* p, bool quick);

int flow3d_solver_compute_host(flow3d_solver* s, const float* frame_0, const float* frame_1,
                               const flow3d_params* params, float* flow_u, float* flow_v,
                               float* flow_w) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!frame_0 || !frame_1 || !flow_u || !flow_v || !flow_w) return FLOW3D_ERR_INVALID_ARG;
  F3D_TRY(check_params(params));
  F3D_CUDA(cudaSetDevice(s->device));
  // launch shapes: a QUICK tuning pass once per solver and parameter set, OUTSIDE the timed bracket (this call
  // is synchronous by contract; the asynchronous device-buffer call never tunes); flow3d_solver_tune runs the
  // full pass, whose results are persisted
  F3D_TRY(solver_tune(s, params, /*quick=*/true));
  cudaStream_t st = s->stream;
  const size_t wb = s->W * 4, rows = s->H * s->D;
  // same bracket as the reference's timer: H2D -> all levels -> D2H (optical_flow_e.cpp:169 -> :579)
  s->timer.reset();
  F3D_CUDA(cudaEventRecord(s->ev[0], st));
  s->timer.mark(FLOW3D_STAGE_COPY, st);
  float* in0 = s->buf(B_F0L);
  float* in1 = s->buf(B_F1L);
  F3D_CUDA(cudaMemcpy2DAsync(in0, s->ld * 4, frame_0, wb, wb, rows, cudaMemcpyHostToDevice, st));
  F3D_CUDA(cudaMemcpy2DAsync(in1, s->ld * 4, frame_1, wb, wb, rows, cudaMemcpyHostToDevice, st));
  F3D_CUDA(cudaEventRecord(s->ev[2], st));
  float *u, *v, *w;
  F3D_TRY(run_pyramid(s, in0, in1, s->ld, params, &u, &v, &w, st));
  F3D_CUDA(cudaEventRecord(s->ev[3], st));
  s->timer.mark(FLOW3D_STAGE_COPY, st);
  F3D_CUDA(cudaMemcpy2DAsync(flow_u, wb, u, s->ld * 4, wb, rows, cudaMemcpyDeviceToHost, st));
  F3D_CUDA(cudaMemcpy2DAsync(flow_v, wb, v, s->ld * 4, wb, rows, cudaMemcpyDeviceToHost, st));
  F3D_CUDA(cudaMemcpy2DAsync(flow_w, wb, w, s->ld * 4, wb, rows, cudaMemcpyDeviceToHost, st));
  s->timer.mark(-1, st);
  F3D_CUDA(cudaEventRecord(s->ev[1], st));
  F3D_CUDA(cudaStreamSynchronize(st));
  s->timer.finish();
  F3D_CUDA(cudaEventElapsedTime(&s->last_ms[0], s->ev[0], s->ev[1]));
  F3D_CUDA(cudaEventElapsedTime(&s->last_ms[1], s->ev[2], s->ev[3]));
  return FLOW3D_OK;
}

int flow3d_solver_set_verbose(flow3d_solver* s, int enable) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  s->verbose = enable != 0;
  return FLOW3D_OK;
}

int flow3d_solver_set_profiling(flow3d_solver* s, int enable) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  s->timer.enabled = enable != 0;
  s->timer.reset();
  return FLOW3D_OK;
}

int flow3d_solver_stage_times(flow3d_solver* s, float ms[FLOW3D_STAGE_COUNT],
                              double units[FLOW3D_STAGE_COUNT], uint64_t launches[FLOW3D_STAGE_COUNT]) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!ms || !units || !launches) return FLOW3D_ERR_INVALID_ARG;
  if (s->timer.used && s->timer.ms[FLOW3D_STAGE_SWEEP] == 0.f) {  // device-buffer call: not finished yet
    F3D_CUDA(cudaEventSynchronize(s->timer.pool[s->timer.used - 1]));
    s->timer.finish();
  }
  for (int i = 0; i < FLOW3D_STAGE_COUNT; ++i) {
    ms[i] = s->timer.ms[i];
    units[i] = s->timer.units[i];
    launches[i] = s->timer.launches[i];
  }
  return FLOW3D_OK;
}

// ---- launch-shape tuning (synchronous by contract) ------------------------------------------------
int flow3d_tune_kernels(const size_t dims[3], size_t ld, const flow3d_zslab* slab, const float h[3],
                        float* scratch, size_t scratch_floats, void* stream) {
  F3D_TRY(check_volume(scratch, dims, ld));
  F3D_TRY(check_slab(dims, slab));
  if (!h) return FLOW3D_ERR_INVALID_ARG;
  const Dims g = make_slab_dims(dims, ld, slab);
  const size_t n = (size_t)g.ps * g.d;
  if (scratch_floats < 16 * n) return FLOW3D_ERR_INVALID_ARG;
  float* bufs[16];
  for (int i = 0; i < 16; ++i) bufs[i] = scratch + (size_t)i * n;
  cudaStream_t st = S(stream);
  F3D_TRY(tune_level_kernels(g, make_range(g, slab), bufs, h[0], h[1], h[2], st));
  F3D_CUDA(cudaStreamSynchronize(st));
  return FLOW3D_OK;
}

int flow3d_tune_query(int kernel, const size_t dims[3], size_t ld, const flow3d_zslab* slab, int out[3]) {
  if (!dims || !out || kernel < 0 || kernel > 3) return FLOW3D_ERR_INVALID_ARG;
  F3D_TRY(check_slab(dims, slab));
  const Dims g = make_slab_dims(dims, ld, slab);
  return tune_query(kernel, g, make_range(g, slab), out);
}

static int solver_tune(flow3d_solver* s, const flow3d_params* p, bool quick) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  F3D_TRY(check_params(p));
  F3D_CUDA(cudaSetDevice(s->device));
  if (s->tuned_scale == p->warp_scale_factor && s->tuned_levels == p->warp_levels_count && (quick || s->tuned_full))
    return FLOW3D_OK;
  const size_t max_level = flow3d_max_warp_level(s->W, s->H, s->D, p->warp_scale_factor);
  const int top = (int)std::min(p->warp_levels_count, max_level) - 1;
  for (int level = top; level >= 0; --level) {
    size_t cur[3];
    float h[3];
    flow3d_level_geometry(s->W, s->H, s->D, p->warp_scale_factor, level, cur, h);
    const Dims g = make_dims(cur, aligned_ld(cur[0]));
    float* bufs[16];
    for (int i = 0; i < 16; ++i) bufs[i] = s->buf(flow3d_solver::kVolumes - 16 + i);
    F3D_TRY(tune_level_kernels(g, ZRange{0, g.d}, bufs, h[0], h[1], h[2], s->stream, quick));
  }
  F3D_CUDA(cudaStreamSynchronize(s->stream));
  s->tuned_scale = p->warp_scale_factor;
  s->tuned_levels = p->warp_levels_count;
  s->tuned_full = !quick;
  return FLOW3D_OK;
}

int flow3d_solver_tune(flow3d_solver* s, const flow3d_params* p) { return solver_tune(s, p, false); }

int flow3d_set_pdl(int mode) { return pdl_set_mode(mode); }

int flow3d_gauss_taps(float sigma, float* taps, size_t capacity, size_t* radius) {
  if (!taps || !radius || !(sigma > 0.f)) return FLOW3D_ERR_INVALID_ARG;
  float t[2 * 32 + 1];
  const int r = gauss_taps(sigma, t, 32);
  if (r < 0) return FLOW3D_ERR_UNSUPPORTED;
  if (capacity < (size_t)(2 * r + 1)) return FLOW3D_ERR_INVALID_ARG;
  for (int i = 0; i < 2 * r + 1; ++i) taps[i] = t[i];
  *radius = (size_t)r;
  return FLOW3D_OK;
}

size_t flow3d_update_norm_workspace_bytes(void) { return update_norm_workspace_bytes(); }

int flow3d_update_norm(const float* a0, const float* a1, const float* a2, const float* b0,
                       const float* b1, const float* b2, const size_t dims[3], size_t ld,
                       const flow3d_zslab* slab, double* out_dev, void* workspace, void* stream) {
  const void* ps[] = {a0, a1, a2, b0, b1, b2};
  for (const void* p : ps) F3D_TRY(check_volume(p, dims, ld));
  if (!out_dev || !workspace) return FLOW3D_ERR_INVALID_ARG;
  F3D_TRY(check_slab(dims, slab));
  const Dims g = make_slab_dims(dims, ld, slab);
  return launch_update_norm(a0, a1, a2, b0, b1, b2, g, make_range(g, slab), out_dev, workspace, S(stream));
}

int flow3d_solver_set_diagnostics(flow3d_solver* s, int enable, float update_tolerance) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!(update_tolerance >= 0.f)) return FLOW3D_ERR_INVALID_ARG;
  s->diag_enabled = enable != 0;
  s->diag.tol = enable ? update_tolerance : 0.f;
  return FLOW3D_OK;
}

int flow3d_solver_diagnostics(flow3d_solver* s, size_t* n_levels, size_t* outer_per_level, double* rms,
                              double* max_abs, size_t capacity, size_t* n_records) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!n_levels || !n_records) return FLOW3D_ERR_INVALID_ARG;
  F3D_CUDA(cudaSetDevice(s->device));
  F3D_CUDA(cudaDeviceSynchronize());  // the record copy of a device-buffer solve may still be in flight
  const size_t nl = s->diag_level_outer.size();
  size_t total = 0;
  for (size_t l = 0; l < nl; ++l) total += s->diag_level_outer[l];
  if (total > s->diag.used) total = s->diag.used;
  *n_levels = nl;
  *n_records = total;
  size_t k = 0;
  for (size_t l = 0; l < nl; ++l) {
    if (outer_per_level && l < capacity) outer_per_level[l] = s->diag_level_outer[l];
    for (size_t i = 0; i < s->diag_level_outer[l] && k < total; ++i, ++k) {
      if (k < capacity) {
        if (rms) rms[k] = std::sqrt(s->diag.host[2 * k] / (3.0 * s->diag_level_voxels[l]));
        if (max_abs) max_abs[k] = s->diag.host[2 * k + 1];
      }
    }
  }
  return FLOW3D_OK;
}

int flow3d_selftest_fast_div(uint64_t n_pairs, uint64_t seed, int mode, uint64_t out[3]) {
  if (!out || n_pairs == 0 || mode < 0 || mode > 3) return FLOW3D_ERR_INVALID_ARG;
  if (flow3d_device_count() <= 0) return FLOW3D_ERR_NO_DEVICE;
  unsigned long long r[3] = {0, 0, 0};
  F3D_TRY(launch_fast_div_selftest(n_pairs, seed, mode, r));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
  return FLOW3D_OK;
}

int flow3d_synth_pair(size_t width, size_t height, size_t depth, size_t z0, size_t nz, size_t ld,
                      uint64_t seed, float* frame_0, float* frame_1, float* truth_u,
                      float* truth_v, float* truth_w, void* stream) {
  if (width == 0 || height == 0 || depth == 0 || ld < width) return FLOW3D_ERR_INVALID_ARG;
  return launch_synth(width, height, depth, z0, nz, ld, seed, frame_0, frame_1, truth_u, truth_v,
                      truth_w, S(stream));
}

}  // extern "C"
