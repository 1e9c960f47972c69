// sharded_solver.cu -- libflow3d_b200_mgpu.so: the z-sharded multi-GPU flow solve (include/flow3d_mgpu_c.h).
//
// Host-side C++ over two things only: the single-GPU library's C ABI (the *_slab stage functions of
// include/flow3d_c.h -- every kernel launch happens there) and NCCL (neighbour send/recv over NVLink, one
// scalar max, one all-gather for coarse level frames).  The orchestration follows
// OpticalFlowE::ComputeFlow (reference: src/optical_flow/optical_flow_e.cpp:179-473) level by level; the
// slab scheme's precedent in the reference is src/cuda_operations/partial_data/cuda_operation_solve_p.cpp:358-417.
//
// Per rank and level:  own planes [a,b), buffer [A,B) = [a-H, b+H) clipped to the level, H = inner+1.
//   prolongation of the previous level's flow onto [a,b)  -> neighbour exchange of H ghost planes
//   level frames resampled from this rank's z-slab of the blurred full-resolution frames
//   warp + derivatives on [A+1, B-1)
//   outer x { phi on [A+1,B-1), sweep j on [A+1+j, B-1-j): ONE C call, no communication;
//             one grouped ncclSend/ncclRecv of H planes of du,dv,dw per neighbour on the same stream }
//   u += du; 5^3 median on [A+2, B-2).
// Levels too thin to shard are computed by every rank (replicas): no communication, and every rank holds
// the whole flow when the first sharded level starts.  (BASELINE config 5 words this as "coarse levels
// gathered onto one GPU"; replicating costs the same wall time -- the other GPUs would idle -- and saves
// the scatter.)
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "flow3d_mgpu_c.h"

namespace {

#define M_TRY(expr)                    \
  do {                                 \
    int rc_ = (expr);                  \
    if (rc_ != FLOW3D_OK) return rc_;  \
  } while (0)
#define M_CUDA(call)                                                                                  \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      std::fprintf(stderr, "flow3d_mgpu: %s: %s\n", #call, cudaGetErrorString(e_));                   \
      return e_ == cudaErrorMemoryAllocation ? FLOW3D_ERR_OUT_OF_MEMORY : FLOW3D_ERR_CUDA;            \
    }                                                                                                 \
  } while (0)
#define M_NCCL(call)                                                                                  \
  do {                                                                                                \
    ncclResult_t r_ = (call);                                                                         \
    if (r_ != ncclSuccess) {                                                                          \
      std::fprintf(stderr, "flow3d_mgpu: %s: %s\n", #call, ncclGetErrorString(r_));                   \
      return FLOW3D_ERR_CUDA;                                                                         \
    }                                                                                                 \
  } while (0)

// a z-slab of one field of one level: planes [A, A+dl) of a level whose depth is dg
struct Slab {
  float* p = nullptr;
  size_t A = 0, dl = 0;
  size_t w = 0, h = 0, ld = 0, dg = 0;
  size_t plane() const { return ld * h; }
  size_t B() const { return A + dl; }
  float* at(size_t g) const { return p + (g - A) * plane(); }
  size_t floats() const { return plane() * dl; }
};
Slab view(const Slab& s, size_t g_lo, size_t g_hi) {
  Slab v = s;
  v.p = s.at(g_lo);
  v.A = g_lo;
  v.dl = g_hi - g_lo;
  return v;
}

// input planes [lo, hi) read by output planes [o_lo, o_hi) of an axis resampled from a to b samples: the
// kernels' own float arithmetic (reference: src/kernels/resample_3d.cu:41-48)
void source_range(size_t o_lo, size_t o_hi, size_t a, size_t b, size_t* lo, size_t* hi) {
  const float delta = (float)a / (float)b;
  const float flo = std::floor((float)o_lo * delta);
  const float fhi = std::fmin((float)a, std::ceil((float)o_hi * delta));
  long long l = (long long)flo, h = (long long)fhi;
  if (l < 0) l = 0;
  if (h > (long long)a) h = (long long)a;
  *lo = (size_t)l;
  *hi = (size_t)h;
}

struct Level {
  int level;
  size_t dims[3];
  float h[3];
};

}  // namespace

struct flow3d_sharded {
  size_t W = 0, H = 0, D = 0;
  int device = 0, rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t st = nullptr;  // stream of the running compute call
  float* scalar_dev = nullptr;
  float* scalar_host = nullptr;  // pinned
  // overlapped ghost exchange: NCCL runs on its own stream, ordered against the solve's stream by events
  cudaStream_t comm_st = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_comm = nullptr;
  bool overlap = true;
  size_t overlap_min_plane = 360000;  // voxels per plane from which the overlapped exchange is used
  bool use_arena = true;
  bool log = false;
  int log_level = 0;
  double log_t = 0;
  // memory: stream-ordered pool, nothing is returned to the OS between solves
  size_t live_bytes = 0, peak_bytes = 0;
  std::map<float*, size_t> sizes;
  // profiling
  bool profiling = false;
  std::vector<std::pair<int, cudaEvent_t>> marks;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  float phase_ms[FLOW3D_MGPU_PHASE_COUNT] = {0};
  double stats[8] = {0};
  float tuned_scale = -1.f;
  size_t tuned_levels = 0, tuned_inner = 0;
  size_t min_planes = 12, min_voxels = (size_t)1 << 18;

  // Device memory of a solve.  The FIRST solve allocates through the stream-ordered pool and records its
  // high-water mark; after it one arena of that size (+10 %) is allocated and every later solve carves its
  // buffers out of it with a first-fit free list on the host: no driver call in the hot path (the pool was
  // measured to block the host for 100-300 ms per large level while it re-mapped memory), and the same
  // deterministic layout every solve.  Reuse of a freed block is safe without events: all work is ordered
  // on the solve's stream, and the communication stream is joined before anything it touched is released.
  // A request the arena cannot place (fragmentation, larger reach) falls back to the pool.
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::map<size_t, size_t> arena_free;  // offset -> size
  size_t pool_peak = 0;                 // high-water mark of a solve that ran without the arena
  float* arena_take(size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    for (auto it = arena_free.begin(); it != arena_free.end(); ++it) {
      if (it->second < bytes) continue;
      const size_t off = it->first, rest = it->second - bytes;
      arena_free.erase(it);
      if (rest) arena_free[off + bytes] = rest;
      return reinterpret_cast<float*>(arena + off);
    }
    return nullptr;
  }
  void arena_give(float* p, size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    size_t off = (size_t)(reinterpret_cast<char*>(p) - arena);
    auto next = arena_free.lower_bound(off);
    if (next != arena_free.begin()) {
      auto prev = std::prev(next);
      if (prev->first + prev->second == off) { off = prev->first; bytes += prev->second; arena_free.erase(prev); }
    }
    if (next != arena_free.end() && off + bytes == next->first) { bytes += next->second; arena_free.erase(next); }
    arena_free[off] = bytes;
  }
  bool in_arena(const float* p) const {
    const char* c = reinterpret_cast<const char*>(p);
    return arena && c >= arena && c < arena + arena_bytes;
  }
  int alloc(size_t floats, float** out) {
    if (floats == 0) floats = 4;
    const size_t bytes = floats * sizeof(float);
    *out = arena ? arena_take(bytes) : nullptr;
    if (!*out) M_CUDA(cudaMallocAsync(reinterpret_cast<void**>(out), bytes, st));
    sizes[*out] = bytes;
    live_bytes += bytes;
    peak_bytes = std::max(peak_bytes, live_bytes);
    return FLOW3D_OK;
  }
  int zeros(size_t floats, float** out) {
    M_TRY(alloc(floats, out));
    M_CUDA(cudaMemsetAsync(*out, 0, std::max<size_t>(floats, 4) * sizeof(float), st));
    return FLOW3D_OK;
  }
  void release(float* p) {
    if (!p) return;
    auto it = sizes.find(p);
    if (it == sizes.end()) return;
    live_bytes -= it->second;
    if (in_arena(p)) arena_give(p, it->second);
    else cudaFreeAsync(p, st);
    sizes.erase(it);
  }
  double host_now() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  std::vector<std::pair<int, double>> host_marks;  // log mode: host clock at every phase mark
  void mark(int phase) {
    if (log) host_marks.emplace_back(phase, host_now());
    if (!profiling) return;
    if (events_used == event_pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { profiling = false; return; }
      event_pool.push_back(e);
    }
    cudaEvent_t e = event_pool[events_used++];
    cudaEventRecord(e, st);
    marks.emplace_back(phase, e);
  }
};

namespace {

size_t aligned_ld(size_t w) { return flow3d_aligned_ld(w); }

bool is_sharded(const flow3d_sharded* s, const size_t dims[3], size_t Hg) {
  if (s->world == 1) return false;
  const size_t per = dims[2] / (size_t)s->world;
  return per >= std::max(s->min_planes, 2 * Hg) + 3 && dims[0] * dims[1] * per >= s->min_voxels;
}

void level_ranges(const flow3d_sharded* s, const size_t dims[3], size_t Hg, int rank, bool sharded, size_t* a, size_t* b,
                  size_t* A, size_t* B) {
  const size_t d = dims[2];
  if (sharded) {
    flow3d_sharded_own_range(d, rank, s->world, a, b);
    *A = *a >= Hg ? *a - Hg : 0;
    *B = std::min(d, *b + Hg);
  } else {
    *a = 0; *b = d; *A = 0; *B = d;
  }
}

// fill ghosts [A,a) and [b,B) of every field from the z neighbours' owned planes: one NCCL group on the
// solve's stream (no host synchronisation; the kernels queued behind it wait on the device)
int exchange(flow3d_sharded* s, float* const* fields, int nf, size_t plane, size_t A, size_t B, size_t a, size_t b,
             size_t Hg) {
  if (s->world == 1) return FLOW3D_OK;
  const bool lo = s->rank > 0 && a > A, hi = s->rank < s->world - 1 && B > b;
  if (!lo && !hi) return FLOW3D_OK;
  const size_t nsend = std::min(Hg, b - a);
  M_NCCL(ncclGroupStart());
  for (int f = 0; f < nf; ++f) {
    float* t = fields[f];
    if (lo) {
      M_NCCL(ncclRecv(t, (a - A) * plane, ncclFloat, s->rank - 1, s->comm, s->st));
      M_NCCL(ncclSend(t + (a - A) * plane, nsend * plane, ncclFloat, s->rank - 1, s->comm, s->st));
    }
    if (hi) {
      M_NCCL(ncclSend(t + (b - A - nsend) * plane, nsend * plane, ncclFloat, s->rank + 1, s->comm, s->st));
      M_NCCL(ncclRecv(t + (b - A) * plane, (B - b) * plane, ncclFloat, s->rank + 1, s->comm, s->st));
    }
  }
  M_NCCL(ncclGroupEnd());
  s->stats[2] += 1;
  s->stats[3] += (double)nf * ((lo ? nsend : 0) + (hi ? nsend : 0)) * plane * sizeof(float);
  return FLOW3D_OK;
}

// The same exchange, overlappable: the boundary planes are first copied into `sendbuf` on the solve's
// stream (the next iteration's early sweeps overwrite them), then the NCCL group runs on the communication
// stream: sends from the copy, receives straight into the ghost planes.  The caller makes the solve's stream
// wait for ev_comm before anything touches the ghosts.
int exchange_async(flow3d_sharded* s, float* const* fields, int nf, size_t plane, size_t A, size_t B, size_t a, size_t b,
                   size_t Hg, float* sendbuf) {
  const bool lo = s->rank > 0 && a > A, hi = s->rank < s->world - 1 && B > b;
  if (!lo && !hi) return FLOW3D_OK;
  const size_t nsend = std::min(Hg, b - a), chunk = nsend * plane;
  for (int f = 0; f < nf; ++f) {
    if (lo)
      M_CUDA(cudaMemcpyAsync(sendbuf + (size_t)(2 * f) * chunk, fields[f] + (a - A) * plane, chunk * sizeof(float),
                             cudaMemcpyDeviceToDevice, s->st));
    if (hi)
      M_CUDA(cudaMemcpyAsync(sendbuf + (size_t)(2 * f + 1) * chunk, fields[f] + (b - A - nsend) * plane,
                             chunk * sizeof(float), cudaMemcpyDeviceToDevice, s->st));
  }
  M_CUDA(cudaEventRecord(s->ev_pack, s->st));
  M_CUDA(cudaStreamWaitEvent(s->comm_st, s->ev_pack, 0));
  M_NCCL(ncclGroupStart());
  for (int f = 0; f < nf; ++f) {
    if (lo) {
      M_NCCL(ncclRecv(fields[f], (a - A) * plane, ncclFloat, s->rank - 1, s->comm, s->comm_st));
      M_NCCL(ncclSend(sendbuf + (size_t)(2 * f) * chunk, chunk, ncclFloat, s->rank - 1, s->comm, s->comm_st));
    }
    if (hi) {
      M_NCCL(ncclSend(sendbuf + (size_t)(2 * f + 1) * chunk, chunk, ncclFloat, s->rank + 1, s->comm, s->comm_st));
      M_NCCL(ncclRecv(fields[f] + (b - A) * plane, (B - b) * plane, ncclFloat, s->rank + 1, s->comm, s->comm_st));
    }
  }
  M_NCCL(ncclGroupEnd());
  M_CUDA(cudaEventRecord(s->ev_comm, s->comm_st));
  s->stats[2] += 1;
  s->stats[3] += (double)nf * ((lo ? nsend : 0) + (hi ? nsend : 0)) * plane * sizeof(float);
  return FLOW3D_OK;
}

// box resample of a slab of one level onto planes [out_lo, out_hi) of a slab of another level
int resample_slab(flow3d_sharded* s, const Slab& src, size_t idg, Slab& out, size_t out_lo, size_t out_hi) {
  const size_t in_dims[3] = {src.w, src.h, src.dl};
  const size_t out_dims[3] = {out.w, out.h, out.dl};
  const flow3d_zslab in_slab = {src.A, idg, 0, src.dl};
  const flow3d_zslab out_slab = {out.A, out.dg, out_lo - out.A, out_hi - out.A};
  float *ta = nullptr, *tb = nullptr;
  M_TRY(s->alloc(src.dl * src.h * aligned_ld(out.w), &ta));
  M_TRY(s->alloc(src.dl * out.h * aligned_ld(out.w), &tb));
  const int rc = flow3d_resample_slab(src.p, in_dims, src.ld, &in_slab, out.p, out_dims, out.ld, &out_slab, ta, tb, s->st);
  s->release(ta);
  s->release(tb);
  return rc;
}

// this rank's z-slab [Va, Vb) of the two blurred full-resolution frames, and the level frames made from it
struct Frames {
  flow3d_sharded* s = nullptr;
  size_t ghost = 0, Va = 0, Vb = 0;
  float* owned[2] = {nullptr, nullptr};  // buffers the slabs point into (nullptr: caller's memory)
  Slab slab[2];
  float* gathered[2] = {nullptr, nullptr};
  int gathered_level[2] = {-1, -1};

  int init(flow3d_sharded* solver, const float* raw0, const float* raw1, size_t raw_z0, size_t raw_planes, size_t ld,
           float sigma, size_t frame_ghost) {
    s = solver;
    ghost = frame_ghost;
    size_t fa, fb;
    flow3d_sharded_own_range(s->D, s->rank, s->world, &fa, &fb);
    Va = fa >= ghost ? fa - ghost : 0;
    Vb = std::min(s->D, fb + ghost);
    const float* raws[2] = {raw0, raw1};
    for (int i = 0; i < 2; ++i) {
      Slab rs;
      rs.p = const_cast<float*>(raws[i]);
      rs.A = raw_z0; rs.dl = raw_planes; rs.w = s->W; rs.h = s->H; rs.ld = ld; rs.dg = s->D;
      if (sigma > 0.f) {
        const size_t r = (size_t)(3 * sigma);
        const size_t need_lo = Va >= r ? Va - r : 0, need_hi = std::min(s->D, Vb + r);
        if (rs.A > need_lo || rs.B() < need_hi) return FLOW3D_ERR_INVALID_ARG;  // raw slab lacks blur halo planes
        float *out = nullptr, *tmp = nullptr;
        M_TRY(s->alloc(rs.floats(), &out));
        M_TRY(s->alloc(rs.floats(), &tmp));
        const size_t dims[3] = {s->W, s->H, rs.dl};
        const flow3d_zslab sl = {rs.A, s->D, Va - rs.A, Vb - rs.A};
        const int rc = flow3d_gauss_blur_slab(rs.p, out, tmp, dims, ld, &sl, sigma, s->st);
        s->release(tmp);
        if (rc != FLOW3D_OK) { s->release(out); return rc; }
        owned[i] = out;
        rs.p = out;
      } else if (rs.A > Va || rs.B() < Vb) {
        return FLOW3D_ERR_INVALID_ARG;
      }
      slab[i] = view(rs, Va, Vb);
    }
    return FLOW3D_OK;
  }
  void destroy() {
    for (int i = 0; i < 2; ++i) {
      s->release(owned[i]);
      s->release(gathered[i]);
      owned[i] = gathered[i] = nullptr;
    }
  }
  // ranges[r] = planes rank r requests; true if every request's source interval lies in that rank's slab
  bool fits_everywhere(const size_t dims[3], const std::vector<std::pair<size_t, size_t>>& ranges) const {
    for (int r = 0; r < s->world; ++r) {
      size_t fa, fb;
      flow3d_sharded_own_range(s->D, r, s->world, &fa, &fb);
      const size_t va = fa >= ghost ? fa - ghost : 0, vb = std::min(s->D, fb + ghost);
      size_t s_lo = ranges[r].first, s_hi = ranges[r].second;
      if (dims[2] != s->D) source_range(ranges[r].first, ranges[r].second, s->D, dims[2], &s_lo, &s_hi);
      if (s_lo < va || s_hi > vb) return false;
    }
    return true;
  }
  // planes [lo,hi) of a level frame box-resampled from the full-resolution slab (optical_flow_e.cpp:279-299:
  // always from full resolution); *owned_out receives the buffer to release (nullptr: a view)
  int level_frame(int which, int level, const size_t dims[3], size_t lo, size_t hi,
                  const std::vector<std::pair<size_t, size_t>>& ranges, Slab* out, float** owned_out) {
    *owned_out = nullptr;
    Slab o;
    o.w = dims[0]; o.h = dims[1]; o.ld = aligned_ld(dims[0]); o.dg = dims[2]; o.A = lo; o.dl = hi - lo;
    if (fits_everywhere(dims, ranges)) {
      if (level == 0) {
        *out = view(slab[which], lo, hi);
        return FLOW3D_OK;
      }
      size_t s_lo, s_hi;
      source_range(lo, hi, s->D, dims[2], &s_lo, &s_hi);
      M_TRY(s->alloc(o.floats(), &o.p));
      const Slab src = view(slab[which], s_lo, s_hi);
      const int rc = resample_slab(s, src, s->D, o, lo, hi);
      if (rc != FLOW3D_OK) { s->release(o.p); return rc; }
      *owned_out = o.p;
      *out = o;
      return FLOW3D_OK;
    }
    // (level 0 included: a finest level too thin to shard is computed by every rank, each of which then needs
    // every plane -- the all-gather's "resample" from D to D planes is the identity, bit for bit)
    if (gathered_level[which] != level) {
      s->release(gathered[which]);
      gathered[which] = nullptr;
      M_TRY(gather(which, dims, &gathered[which]));
      gathered_level[which] = level;
    }
    Slab full = o;
    full.p = gathered[which]; full.A = 0; full.dl = dims[2];
    *out = view(full, lo, hi);
    return FLOW3D_OK;
  }
  // coarse level whose source intervals are longer than the frame ghost: every rank resamples the planes
  // whose source interval STARTS in its own full-resolution range, one all-gather assembles the level
  int gather(int which, const size_t dims[3], float** full_out) {
    const size_t w = dims[0], hh = dims[1], d = dims[2];
    const size_t ldl = aligned_ld(w), plane = ldl * hh;
    const int world = s->world;
    const float delta = (float)s->D / (float)d;
    std::vector<size_t> bounds(world + 1, d);
    for (int r = 0; r < world; ++r) {
      size_t fa, fb;
      flow3d_sharded_own_range(s->D, r, world, &fa, &fb);
      size_t o = 0;
      while (o < d && (long long)std::floor((float)o * delta) < (long long)fa) ++o;
      bounds[r] = o;
    }
    bounds[0] = 0;
    size_t n_max = 1;
    for (int r = 0; r < world; ++r) n_max = std::max(n_max, bounds[r + 1] - bounds[r]);
    const size_t p_lo = bounds[s->rank], p_hi = bounds[s->rank + 1];
    float *piece = nullptr, *all = nullptr, *full = nullptr;
    M_TRY(s->zeros(n_max * plane, &piece));
    if (p_hi > p_lo) {
      size_t s_lo, s_hi;
      source_range(p_lo, p_hi, s->D, d, &s_lo, &s_hi);
      if (s_lo < Va || s_hi > Vb) return FLOW3D_ERR_INVALID_ARG;  // frame ghost smaller than one coarse source interval
      Slab o;
      o.p = piece; o.A = p_lo; o.dl = p_hi - p_lo; o.w = w; o.h = hh; o.ld = ldl; o.dg = d;
      M_TRY(resample_slab(s, view(slab[which], s_lo, s_hi), s->D, o, p_lo, p_hi));
    }
    M_TRY(s->alloc((size_t)world * n_max * plane, &all));
    M_NCCL(ncclAllGather(piece, all, n_max * plane, ncclFloat, s->comm, s->st));
    M_TRY(s->alloc(d * plane, &full));
    for (int r = 0; r < world; ++r) {
      const size_t n = bounds[r + 1] - bounds[r];
      if (n > 0)
        M_CUDA(cudaMemcpyAsync(full + bounds[r] * plane, all + (size_t)r * n_max * plane, n * plane * sizeof(float),
                               cudaMemcpyDeviceToDevice, s->st));
    }
    s->release(piece);
    s->release(all);
    s->stats[4] += 1;
    *full_out = full;
    return FLOW3D_OK;
  }
};

std::vector<Level> schedule(const flow3d_sharded* s, const flow3d_params* p) {
  const size_t max_level = flow3d_max_warp_level(s->W, s->H, s->D, p->warp_scale_factor);
  const int top = (int)std::min(p->warp_levels_count, max_level) - 1;
  std::vector<Level> out;
  for (int lv = top; lv >= 0; --lv) {
    Level L;
    L.level = lv;
    flow3d_level_geometry(s->W, s->H, s->D, p->warp_scale_factor, lv, L.dims, L.h);
    out.push_back(L);
  }
  return out;
}

int solve(flow3d_sharded* s, Frames& frames, const flow3d_params* P, float* out_u, float* out_v, float* out_w,
          size_t out_ld, size_t out_capacity, size_t* out_a, size_t* out_b) {
  const size_t inner = P->inner_iterations_count, outer = P->outer_iterations_count;
  const size_t Hg = inner + 1;
  const std::vector<Level> sched = schedule(s, P);
  bool have_prev = false;
  size_t pdims[3] = {0, 0, 0}, pv_lo = 0, pv_hi = 0;
  Slab pflow[3];
  size_t a = 0, b = 0, A = 0, B = 0;
  for (const Level& L : sched) {
    const size_t w = L.dims[0], hh = L.dims[1], d = L.dims[2];
    const size_t ldl = aligned_ld(w), plane = ldl * hh;
    const bool sharded = is_sharded(s, L.dims, Hg);
    level_ranges(s, L.dims, Hg, s->rank, sharded, &a, &b, &A, &B);
    const size_t dl = B - A;
    s->stats[sharded ? 0 : 1] += 1;
    const size_t dims_l[3] = {w, hh, dl};

    // ---- flow at this level: zeros, or box prolongation of the previous level (optical_flow_e.cpp:304-344) ----
    s->mark(FLOW3D_MGPU_PHASE_PROLONGATION);
    Slab flow[3];
    for (int c = 0; c < 3; ++c) {
      flow[c].w = w; flow[c].h = hh; flow[c].ld = ldl; flow[c].dg = d; flow[c].A = A; flow[c].dl = dl;
      M_TRY(s->zeros(dl * plane, &flow[c].p));
    }
    if (have_prev) {
      size_t s_lo, s_hi;
      source_range(a, b, pdims[2], d, &s_lo, &s_hi);
      if (s_lo < pv_lo || s_hi > pv_hi) return FLOW3D_ERR_INVALID_ARG;  // prolongation would read invalid planes
      for (int c = 0; c < 3; ++c) {
        M_TRY(resample_slab(s, view(pflow[c], s_lo, s_hi), pdims[2], flow[c], a, b));
        s->release(pflow[c].p);
      }
      s->mark(FLOW3D_MGPU_PHASE_FLOW_EXCHANGE);
      float* fl[3] = {flow[0].p, flow[1].p, flow[2].p};
      M_TRY(exchange(s, fl, 3, plane, A, B, a, b, Hg));
    }
    // ---- level frames: frame 1 with the data-dependent z reach of the warp ---------------------------------
    s->mark(FLOW3D_MGPU_PHASE_LEVEL_FRAMES);
    float wmax = 0.f;
    if (have_prev) {
      M_TRY(flow3d_absmax(flow[2].p, dims_l, ldl, s->scalar_dev, s->st));
      if (s->world > 1)  // same reach on every rank => same local / gather decision
        M_NCCL(ncclAllReduce(s->scalar_dev, s->scalar_dev, 1, ncclFloat, ncclMax, s->comm, s->st));
      M_CUDA(cudaMemcpyAsync(s->scalar_host, s->scalar_dev, sizeof(float), cudaMemcpyDeviceToHost, s->st));
      M_CUDA(cudaStreamSynchronize(s->st));
      wmax = *s->scalar_host;
      if (!(wmax == wmax) || wmax > 1e9f) wmax = 1e9f;  // NaN / absurd flow: clamp the reach to the level
    }
    size_t reach = have_prev ? (size_t)std::ceil((double)wmax / (double)L.h[2]) + 2 : 1;
    if (reach > d) reach = d;
    const size_t A1 = A >= reach ? A - reach : 0, B1 = std::min(d, B + reach);
    std::vector<std::pair<size_t, size_t>> rng0(s->world), rng1(s->world);
    for (int r = 0; r < s->world; ++r) {
      size_t ra, rb, rA, rB;
      level_ranges(s, L.dims, Hg, r, sharded, &ra, &rb, &rA, &rB);
      rng0[r] = {rA, rB};
      rng1[r] = {rA >= reach ? rA - reach : 0, std::min(d, rB + reach)};
    }
    Slab f0l, f1l;
    float *f0_owned = nullptr, *f1_owned = nullptr;
    M_TRY(frames.level_frame(0, L.level, L.dims, A, B, rng0, &f0l, &f0_owned));
    M_TRY(frames.level_frame(1, L.level, L.dims, A1, B1, rng1, &f1l, &f1_owned));
    // ---- warp + derivatives on every plane the solver touches (optical_flow_e.cpp:348-369) ------------------
    const size_t lo1 = A == 0 ? A : A + 1, hi1 = B == d ? B : B - 1;
    const flow3d_zslab sl1 = {A, d, lo1 - A, hi1 - A};
    s->mark(FLOW3D_MGPU_PHASE_WARP);
    float* terms[4];
    for (int i = 0; i < 4; ++i) M_TRY(s->alloc(dl * plane, &terms[i]));
    M_TRY(flow3d_warp_derivatives_slab(f0l.p, f1l.p, f1l.A, f1l.dl, flow[0].p, flow[1].p, flow[2].p, dims_l, ldl, &sl1,
                                       L.h, terms[0], terms[1], terms[2], terms[3], s->st));
    s->release(f0_owned);
    s->release(f1_owned);
    // ---- solver (cuda_operation_solve.cpp:183-257) ------------------------------------------------------------
    float *dc[3], *da[3], *phi = nullptr, *ksi = nullptr;
    for (int c = 0; c < 3; ++c) {
      M_TRY(s->zeros(dl * plane, &dc[c]));
      M_TRY(s->zeros(dl * plane, &da[c]));
    }
    M_TRY(s->zeros(dl * plane, &phi));
    M_TRY(s->zeros(dl * plane, &ksi));
    // Sharded levels overlap the ghost exchange with compute: the exchange that follows iteration i runs on
    // the communication stream while the EARLY part of iteration i+1 (every plane that cannot depend on the
    // ghosts: ~95 % of the work) runs on the solve's stream; the LATE part waits for it.
    // Splitting costs ~12 extra small launches per iteration (~0.1 ms); the exchange costs latency + bytes /
    // ~150 GB/s (NCCL send/recv between two peers).  Measured on B200s: a wash at 512^2 planes, a win from
    // ~600^2 on -- smaller levels keep the serial exchange on the solve's stream.
    const bool ovl = sharded && s->overlap && (b - a) >= 4 * Hg && w * hh >= s->overlap_min_plane;
    float* sendbuf = nullptr;
    if (ovl) M_TRY(s->alloc(6 * Hg * plane, &sendbuf));
    bool pending = false;  // an exchange into dc's ghosts is in flight on the communication stream
    auto count_units = [&]() {
      for (size_t j = 1; j <= inner; ++j)
        s->stats[5] += (double)w * hh * (double)((hi1 == d ? hi1 : hi1 - j) - (lo1 == 0 ? lo1 : lo1 + j));
      s->stats[6] += (double)w * hh * (double)(hi1 - lo1);
    };
    for (size_t it = 0; it < outer; ++it) {
      s->mark(FLOW3D_MGPU_PHASE_SOLVER);
      int in_tmp = 0;
      if (!pending) {
        M_TRY(flow3d_outer_iteration_slab(terms[0], terms[1], terms[2], terms[3], flow[0].p, flow[1].p, flow[2].p, dc[0],
                                          dc[1], dc[2], da[0], da[1], da[2], phi, ksi, dims_l, ldl, &sl1, L.h, inner,
                                          P->equation_alpha, P->equation_smoothness, P->equation_data, &in_tmp, s->st));
      } else {
        M_TRY(flow3d_outer_iteration_slab_part(terms[0], terms[1], terms[2], terms[3], flow[0].p, flow[1].p, flow[2].p,
                                               dc[0], dc[1], dc[2], da[0], da[1], da[2], phi, ksi, dims_l, ldl, &sl1, L.h,
                                               inner, P->equation_alpha, P->equation_smoothness, P->equation_data, 1,
                                               a - A, b - A, &in_tmp, s->st));
        s->mark(FLOW3D_MGPU_PHASE_HALO_EXCHANGE);  // what is left of the exchange after the early part
        M_CUDA(cudaStreamWaitEvent(s->st, s->ev_comm, 0));
        s->mark(FLOW3D_MGPU_PHASE_SOLVER);
        pending = false;
        M_TRY(flow3d_outer_iteration_slab_part(terms[0], terms[1], terms[2], terms[3], flow[0].p, flow[1].p, flow[2].p,
                                               dc[0], dc[1], dc[2], da[0], da[1], da[2], phi, ksi, dims_l, ldl, &sl1, L.h,
                                               inner, P->equation_alpha, P->equation_smoothness, P->equation_data, 2,
                                               a - A, b - A, &in_tmp, s->st));
      }
      if (in_tmp)
        for (int c = 0; c < 3; ++c) std::swap(dc[c], da[c]);
      count_units();
      if (sharded) {
        if (ovl) {
          M_TRY(exchange_async(s, dc, 3, plane, A, B, a, b, Hg, sendbuf));
          pending = true;
        } else {
          s->mark(FLOW3D_MGPU_PHASE_HALO_EXCHANGE);
          M_TRY(exchange(s, dc, 3, plane, A, B, a, b, Hg));
        }
      }
    }
    if (pending) {  // the update, the median and the next prolongation read the iterate's ghost planes
      s->mark(FLOW3D_MGPU_PHASE_HALO_EXCHANGE);
      M_CUDA(cudaStreamWaitEvent(s->st, s->ev_comm, 0));
    }
    s->mark(FLOW3D_MGPU_PHASE_UPDATE);
    s->release(sendbuf);
    for (int i = 0; i < 4; ++i) s->release(terms[i]);
    s->release(phi);
    s->release(ksi);
    // ---- u += du (optical_flow_e.cpp:420-438): valid on the whole buffer, the last exchange refreshed du -------
    M_TRY(flow3d_add3(flow[0].p, flow[1].p, flow[2].p, dc[0], dc[1], dc[2], dims_l, ldl, s->st));
    for (int c = 0; c < 3; ++c) s->release(dc[c]);
    // ---- median (:443-473) on everything whose +-r/2 neighbourhood is valid -----------------------------------
    size_t r = P->median_radius;
    if (r % 2 == 0 && r > 1) r -= 1;
    const size_t r2 = r / 2;
    const size_t m_lo = A == 0 ? A : A + r2, m_hi = B == d ? B : B - r2;
    const flow3d_zslab slm = {A, d, m_lo - A, m_hi - A};
    s->mark(FLOW3D_MGPU_PHASE_MEDIAN);
    for (int c = 0; c < 3; ++c) {
      M_TRY(flow3d_median_slab(flow[c].p, da[c], dims_l, ldl, &slm, P->median_radius, s->st));
      s->release(flow[c].p);
      flow[c].p = da[c];
      pflow[c] = flow[c];
    }
    pdims[0] = w; pdims[1] = hh; pdims[2] = d;
    pv_lo = m_lo; pv_hi = m_hi;
    have_prev = true;
    s->mark(-1);
    if (s->log) {  // FLOW3D_MGPU_LOG=1: host clock per level (synchronising; 2 = not synchronising; debugging aid)
      if (s->log_level == 1) cudaStreamSynchronize(s->st);
      const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
      double host_ms[FLOW3D_MGPU_PHASE_COUNT] = {0};  // host time spent ENQUEUING each phase
      for (size_t i = 0; i + 1 < s->host_marks.size(); ++i)
        if (s->host_marks[i].first >= 0)
          host_ms[s->host_marks[i].first] += (s->host_marks[i + 1].second - s->host_marks[i].second) * 1e3;
      s->host_marks.clear();
      std::fprintf(stderr, "[flow3d mgpu r%d] level %d %zux%zux%zu %s [%zu,%zu): %.1f ms since previous; host enqueue ms: "
                   "prol %.1f xchg %.1f frames %.1f warp %.1f solver %.1f halo %.1f upd %.1f med %.1f; pool live %.2f GB\n",
                   s->rank, L.level, w, hh, d, sharded ? (ovl ? "sharded+overlap" : "sharded") : "replicated", a, b,
                   s->log_t > 0 ? (now - s->log_t) * 1e3 : 0.0, host_ms[0], host_ms[1], host_ms[2], host_ms[3], host_ms[4],
                   host_ms[5], host_ms[6], host_ms[7], s->live_bytes / 1e9);
      s->log_t = now;
    }
  }
  // ---- this rank's owned planes of the finest level ---------------------------------------------------------
  if (b - a > out_capacity) return FLOW3D_ERR_INVALID_ARG;
  float* outs[3] = {out_u, out_v, out_w};
  for (int c = 0; c < 3; ++c) {
    M_CUDA(cudaMemcpy2DAsync(outs[c], out_ld * 4, pflow[c].at(a), pflow[c].ld * 4, s->W * 4, s->H * (b - a),
                             cudaMemcpyDeviceToDevice, s->st));
    s->release(pflow[c].p);
  }
  *out_a = a;
  *out_b = b;
  return FLOW3D_OK;
}

}  // namespace

extern "C" {

int flow3d_mgpu_unique_id(void* id128) {
  if (!id128) return FLOW3D_ERR_INVALID_ARG;
  static_assert(sizeof(ncclUniqueId) <= FLOW3D_MGPU_ID_BYTES, "id size");
  ncclUniqueId id;
  M_NCCL(ncclGetUniqueId(&id));
  std::memset(id128, 0, FLOW3D_MGPU_ID_BYTES);
  std::memcpy(id128, &id, sizeof(id));
  return FLOW3D_OK;
}

void flow3d_sharded_own_range(size_t d, int rank, int world, size_t* a, size_t* b) {
  // Cost model: an outer iteration recomputes ~2.5 plane-equivalents of ghost work per neighbour side (phi on
  // H-1 = 5 extra planes, sweep j on 5-j), so with g interior cuts below it a boundary sits 2.5 - 5g/world
  // planes beyond the even split: edge ranks own ~2.5 planes more than interior ones.
  auto cut = [&](int g) -> size_t {
    if (g <= 0) return 0;
    if (g >= world) return d;
    const double x = (double)g * (double)d / (double)world + 2.5 - 5.0 * (double)g / (double)world;
    long long c = (long long)std::floor(x + 0.5);
    if (c < g) c = g;
    if (c > (long long)d - (world - g)) c = (long long)d - (world - g);
    return (size_t)c;
  };
  if (world <= 1 || d < (size_t)(8 * world)) {  // tiny depth: plain even split
    *a = (size_t)rank * d / (size_t)std::max(world, 1);
    *b = (size_t)(rank + 1) * d / (size_t)std::max(world, 1);
    return;
  }
  *a = cut(rank);
  *b = cut(rank + 1);
}

void flow3d_sharded_input_planes(size_t depth, int rank, int world, float sigma, size_t frame_ghost, size_t* lo,
                                 size_t* hi) {
  size_t fa, fb;
  flow3d_sharded_own_range(depth, rank, world, &fa, &fb);
  const size_t r = sigma > 0.f ? (size_t)(3 * sigma) : 0;
  const size_t ext = frame_ghost + r;
  *lo = fa >= ext ? fa - ext : 0;
  *hi = std::min(depth, fb + ext);
}

int flow3d_sharded_create(size_t width, size_t height, size_t depth, int device, int rank, int world, const void* id128,
                          flow3d_sharded** out) {
  if (!out || width < 2 || height < 2 || depth < 2 || world < 1 || rank < 0 || rank >= world) return FLOW3D_ERR_INVALID_ARG;
  if (world > 1 && !id128) return FLOW3D_ERR_INVALID_ARG;
  *out = nullptr;
  const int n = flow3d_device_count();
  if (n <= 0) return FLOW3D_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return FLOW3D_ERR_INVALID_ARG;
  M_CUDA(cudaSetDevice(device));
  flow3d_sharded* s = new flow3d_sharded();
  s->W = width; s->H = height; s->D = depth;
  s->device = device; s->rank = rank; s->world = world;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;  // freed blocks stay in the pool: the second solve allocates nothing
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (cudaMalloc(&s->scalar_dev, 16) != cudaSuccess || cudaMallocHost(&s->scalar_host, 16) != cudaSuccess) {
    flow3d_sharded_destroy(s);
    return FLOW3D_ERR_OUT_OF_MEMORY;
  }
  // highest priority: the exchange's copy kernels must get SM slots while a sweep's grid is still draining
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&s->comm_st, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_pack, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming) != cudaSuccess) {
    flow3d_sharded_destroy(s);
    return FLOW3D_ERR_CUDA;
  }
  if (const char* e = getenv("FLOW3D_MGPU_OVERLAP")) s->overlap = std::atoi(e) != 0;
  if (const char* e = getenv("FLOW3D_MGPU_OVERLAP_MIN_PLANE")) s->overlap_min_plane = std::strtoull(e, nullptr, 10);
  if (const char* e = getenv("FLOW3D_MGPU_ARENA")) s->use_arena = std::atoi(e) != 0;
  if (const char* e = getenv("FLOW3D_MGPU_LOG")) { s->log_level = std::atoi(e); s->log = s->log_level != 0; }
  if (world > 1) {
    // NCCL's send/recv between two peers defaults to 2 channels (~55 GB/s per direction measured on B200 /
    // NVLink 5); more channels per peer raise it ~1.6x.  Only defaults: an explicit setting wins.
    setenv("NCCL_NCHANNELS_PER_PEER", "32", 0);
    setenv("NCCL_MIN_P2P_NCHANNELS", "32", 0);
    setenv("NCCL_MAX_P2P_NCHANNELS", "64", 0);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    const ncclResult_t r = ncclCommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) {
      std::fprintf(stderr, "flow3d_mgpu: ncclCommInitRank: %s\n", ncclGetErrorString(r));
      s->comm = nullptr;
      flow3d_sharded_destroy(s);
      return FLOW3D_ERR_CUDA;
    }
  }
  // testing knobs: shard levels smaller than the production thresholds
  if (const char* e = getenv("FLOW3D_MGPU_MIN_PLANES")) s->min_planes = std::max<size_t>(std::strtoull(e, nullptr, 10), 1);
  if (const char* e = getenv("FLOW3D_MGPU_MIN_VOXELS")) s->min_voxels = std::max<size_t>(std::strtoull(e, nullptr, 10), 1);
  *out = s;
  return FLOW3D_OK;
}

int flow3d_sharded_destroy(flow3d_sharded* s) {
  if (!s) return FLOW3D_OK;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  if (s->comm) ncclCommDestroy(s->comm);
  if (s->ev_pack) cudaEventDestroy(s->ev_pack);
  if (s->ev_comm) cudaEventDestroy(s->ev_comm);
  if (s->comm_st) cudaStreamDestroy(s->comm_st);
  for (cudaEvent_t e : s->event_pool) cudaEventDestroy(e);
  if (s->arena) cudaFree(s->arena);
  if (s->scalar_dev) cudaFree(s->scalar_dev);
  if (s->scalar_host) cudaFreeHost(s->scalar_host);
  delete s;
  return FLOW3D_OK;
}

int flow3d_sharded_output_planes(const flow3d_sharded* s, const flow3d_params* params, size_t* a, size_t* b) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!params || !a || !b) return FLOW3D_ERR_INVALID_ARG;
  const size_t dims[3] = {s->W, s->H, s->D};
  size_t A, B;
  level_ranges(s, dims, params->inner_iterations_count + 1, s->rank,
               is_sharded(s, dims, params->inner_iterations_count + 1), a, b, &A, &B);
  return FLOW3D_OK;
}

// Host-only dry run of the partition arithmetic of a solve (no device, no communicator): for every level and
// rank it checks what solve() relies on -- the prolongation's source planes lie inside the previous level's
// valid planes of the same rank, neighbours agree on the size of every ghost exchange, and every all-gather
// piece of a level frame comes from the rank's own frame slab.  flow3d_sharded_compute runs it first, so that a
// geometry it cannot handle fails on EVERY rank before any rank has entered a collective.
int flow3d_sharded_plan_check(size_t width, size_t height, size_t depth, int world, const flow3d_params* p,
                              size_t min_planes_per_rank, size_t min_voxels_per_rank, size_t frame_ghost,
                              int* bad_level, int* bad_rank) {
  if (!p || world < 1 || width < 2 || height < 2 || depth < 2) return FLOW3D_ERR_INVALID_ARG;
  flow3d_sharded t;
  t.W = width; t.H = height; t.D = depth; t.world = world;
  t.min_planes = std::max<size_t>(min_planes_per_rank, 1);
  t.min_voxels = std::max<size_t>(min_voxels_per_rank, 1);
  const size_t Hg = p->inner_iterations_count + 1;
  size_t r = p->median_radius;
  if (r % 2 == 0 && r > 1) r -= 1;
  const size_t r2 = r / 2;
  auto fail = [&](int level, int rank) {
    if (bad_level) *bad_level = level;
    if (bad_rank) *bad_rank = rank;
    return FLOW3D_ERR_INVALID_ARG;
  };
  const std::vector<Level> sched = schedule(&t, p);
  std::vector<size_t> pv_lo(world, 0), pv_hi(world, 0);
  size_t pd = 0;
  bool have_prev = false;
  for (const Level& L : sched) {
    const size_t d = L.dims[2];
    const bool sharded = is_sharded(&t, L.dims, Hg);
    std::vector<size_t> a(world), b(world), A(world), B(world);
    for (int k = 0; k < world; ++k) level_ranges(&t, L.dims, Hg, k, sharded, &a[k], &b[k], &A[k], &B[k]);
    for (int k = 0; k < world; ++k) {
      if (b[k] <= a[k] || A[k] > a[k] || B[k] < b[k] || B[k] > d) return fail(L.level, k);
      if (sharded) {
        if (k + 1 < world && (b[k] != a[k + 1] || std::min(Hg, b[k] - a[k]) != a[k + 1] - A[k + 1] ||
                              std::min(Hg, b[k + 1] - a[k + 1]) != B[k] - b[k]))
          return fail(L.level, k);  // what one side sends != the ghost planes the other side receives into
        if (k == 0 && a[k] != 0) return fail(L.level, k);
        if (k == world - 1 && b[k] != d) return fail(L.level, k);
      }
      if (have_prev) {
        size_t s_lo, s_hi;
        source_range(a[k], b[k], pd, d, &s_lo, &s_hi);
        if (s_lo < pv_lo[k] || s_hi > pv_hi[k]) return fail(L.level, k);
      }
    }
    // frames: a level whose source intervals (with the warp reach, data-dependent) do not fit every rank's slab
    // is all-gathered, so on every level each rank's all-gather piece must come from its own frame slab;
    // the finest level of a sharded solve must be reachable locally for the assumed reach
    std::vector<size_t> Va(world), Vb(world);
    for (int k = 0; k < world; ++k) {
      size_t fa, fb;
      flow3d_sharded_own_range(depth, k, world, &fa, &fb);
      Va[k] = fa >= frame_ghost ? fa - frame_ghost : 0;
      Vb[k] = std::min(depth, fb + frame_ghost);
    }
    if (world > 1) {
      const float delta = (float)depth / (float)d;
      std::vector<size_t> bounds(world + 1, d);
      for (int k = 0; k < world; ++k) {
        size_t fa, fb;
        flow3d_sharded_own_range(depth, k, world, &fa, &fb);
        size_t o = 0;
        while (o < d && (long long)std::floor((float)o * delta) < (long long)fa) ++o;
        bounds[k] = o;
      }
      bounds[0] = 0;
      for (int k = 0; k < world; ++k) {
        if (bounds[k + 1] <= bounds[k]) continue;
        size_t s_lo, s_hi;
        source_range(bounds[k], bounds[k + 1], depth, d, &s_lo, &s_hi);
        if (s_lo < Va[k] || s_hi > Vb[k]) return fail(L.level, k);
      }
    }
    for (int k = 0; k < world; ++k) {
      const size_t m_lo = A[k] == 0 ? A[k] : A[k] + r2, m_hi = B[k] == d ? B[k] : B[k] - r2;
      pv_lo[k] = m_lo;
      pv_hi[k] = m_hi;
    }
    pd = d;
    have_prev = true;
  }
  return FLOW3D_OK;
}

// Smallest frame ghost (32, 64, 128, ... capped at the depth = every rank holds the whole frames) the plan
// accepts: the coarsest level's source interval is depth / level_depth planes long, which a deep or steep
// pyramid (scale 0.9 x 40 levels: 61 planes; 0.5 x 7: 128) pushes past the default 32.
size_t flow3d_sharded_frame_ghost(size_t width, size_t height, size_t depth, int world, const flow3d_params* p,
                                  size_t min_planes_per_rank, size_t min_voxels_per_rank) {
  if (!p || world < 1) return 0;
  for (size_t g = std::min<size_t>(32, depth);; g = std::min(depth, 2 * g)) {
    if (flow3d_sharded_plan_check(width, height, depth, world, p, min_planes_per_rank, min_voxels_per_rank, g, nullptr,
                                  nullptr) == FLOW3D_OK)
      return g;
    if (g >= depth) return 0;
  }
}

int flow3d_sharded_set_thresholds(flow3d_sharded* s, size_t min_planes_per_rank, size_t min_voxels_per_rank) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  s->min_planes = std::max<size_t>(min_planes_per_rank, 1);
  s->min_voxels = std::max<size_t>(min_voxels_per_rank, 1);
  return FLOW3D_OK;
}

int flow3d_sharded_set_profiling(flow3d_sharded* s, int enable) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  s->profiling = enable != 0;
  return FLOW3D_OK;
}

int flow3d_sharded_phase_ms(flow3d_sharded* s, float ms[FLOW3D_MGPU_PHASE_COUNT]) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!ms) return FLOW3D_ERR_INVALID_ARG;
  M_CUDA(cudaSetDevice(s->device));
  if (!s->marks.empty()) {
    M_CUDA(cudaEventSynchronize(s->marks.back().second));
    for (size_t i = 0; i + 1 < s->marks.size(); ++i) {
      if (s->marks[i].first < 0) continue;
      float t = 0.f;
      if (cudaEventElapsedTime(&t, s->marks[i].second, s->marks[i + 1].second) == cudaSuccess)
        s->phase_ms[s->marks[i].first] += t;
    }
    s->marks.clear();
    s->events_used = 0;
  }
  for (int i = 0; i < FLOW3D_MGPU_PHASE_COUNT; ++i) ms[i] = s->phase_ms[i];
  return FLOW3D_OK;
}

int flow3d_sharded_stats(flow3d_sharded* s, double out[8]) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!out) return FLOW3D_ERR_INVALID_ARG;
  for (int i = 0; i < 7; ++i) out[i] = s->stats[i];
  out[7] = (double)(s->arena ? s->arena_bytes : s->peak_bytes);
  return FLOW3D_OK;
}

int flow3d_sharded_tune(flow3d_sharded* s, const flow3d_params* p) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!p) return FLOW3D_ERR_INVALID_ARG;
  M_CUDA(cudaSetDevice(s->device));
  if (s->tuned_scale == p->warp_scale_factor && s->tuned_levels == p->warp_levels_count &&
      s->tuned_inner == p->inner_iterations_count)
    return FLOW3D_OK;
  cudaStream_t st = nullptr;
  M_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  s->st = st;
  const size_t Hg = p->inner_iterations_count + 1;
  int rc = FLOW3D_OK;
  for (const Level& L : schedule(s, p)) {
    size_t a, b, A, B;
    const bool sharded = is_sharded(s, L.dims, Hg);
    level_ranges(s, L.dims, Hg, s->rank, sharded, &a, &b, &A, &B);
    const size_t d = L.dims[2], dl = B - A;
    const size_t dims_l[3] = {L.dims[0], L.dims[1], dl};
    const size_t ldl = aligned_ld(L.dims[0]), n = ldl * L.dims[1] * dl;
    const size_t lo1 = A == 0 ? A : A + 1, hi1 = B == d ? B : B - 1;
    const flow3d_zslab sl = {A, d, lo1 - A, hi1 - A};
    float* scratch = nullptr;
    rc = s->alloc(16 * n, &scratch);
    if (rc != FLOW3D_OK) break;
    rc = flow3d_tune_kernels(dims_l, ldl, sharded ? &sl : nullptr, L.h, scratch, 16 * n, st);
    s->release(scratch);
    if (rc != FLOW3D_OK) break;
  }
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  s->st = nullptr;
  if (rc == FLOW3D_OK) {
    s->tuned_scale = p->warp_scale_factor;
    s->tuned_levels = p->warp_levels_count;
    s->tuned_inner = p->inner_iterations_count;
  }
  return rc;
}

int flow3d_sharded_compute(flow3d_sharded* s, const float* raw_0, const float* raw_1, size_t raw_z0, size_t raw_planes,
                           size_t ld, const flow3d_params* params, size_t frame_ghost, float* flow_u, float* flow_v,
                           float* flow_w, size_t out_capacity_planes, size_t* out_a, size_t* out_b, void* stream) {
  if (!s) return FLOW3D_ERR_NOT_INITIALIZED;
  if (!raw_0 || !raw_1 || !params || !flow_u || !flow_v || !flow_w || !out_a || !out_b) return FLOW3D_ERR_INVALID_ARG;
  if (ld < s->W || (ld & 3) || raw_z0 + raw_planes > s->D || raw_planes == 0) return FLOW3D_ERR_INVALID_ARG;
  if (params->warp_levels_count == 0 || params->median_radius == 0) return FLOW3D_ERR_INVALID_ARG;
  {  // same verdict on every rank, before any of them enters a collective
    int lv = -1, rk = -1;
    if (flow3d_sharded_plan_check(s->W, s->H, s->D, s->world, params, s->min_planes, s->min_voxels, frame_ghost, &lv,
                                  &rk) != FLOW3D_OK) {
      std::fprintf(stderr, "flow3d_mgpu: %zux%zux%zu over %d ranks with %zu frame ghost planes cannot be partitioned "
                   "(level %d, rank %d)\n", s->W, s->H, s->D, s->world, frame_ghost, lv, rk);
      return FLOW3D_ERR_INVALID_ARG;
    }
  }
  M_CUDA(cudaSetDevice(s->device));
  s->st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < 7; ++i) s->stats[i] = 0;
  for (int i = 0; i < FLOW3D_MGPU_PHASE_COUNT; ++i) s->phase_ms[i] = 0.f;
  s->marks.clear();
  s->events_used = 0;
  s->peak_bytes = s->live_bytes;
  Frames frames;
  s->mark(FLOW3D_MGPU_PHASE_BLUR);
  int rc = frames.init(s, raw_0, raw_1, raw_z0, raw_planes, ld, params->gaussian_sigma, frame_ghost);
  if (rc == FLOW3D_OK) rc = solve(s, frames, params, flow_u, flow_v, flow_w, ld, out_capacity_planes, out_a, out_b);
  frames.destroy();
  s->mark(-1);
  if (rc == FLOW3D_OK && !s->arena && s->use_arena) {  // first solve done: fix the memory layout for the next ones
    s->pool_peak = s->peak_bytes;
    const size_t want = s->pool_peak + s->pool_peak / 10 + ((size_t)64 << 20);
    cudaStreamSynchronize(s->st);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);  // hand the pool's memory over
    void* p = nullptr;
    if (cudaMalloc(&p, want) == cudaSuccess) {
      s->arena = static_cast<char*>(p);
      s->arena_bytes = want;
      s->arena_free.clear();
      s->arena_free[0] = want;
    } else {
      cudaGetLastError();
      s->use_arena = false;  // not enough room for pool + arena: stay on the pool
    }
  }
  if (rc != FLOW3D_OK) {  // drop whatever a failed level left behind
    std::vector<float*> left;
    for (auto& kv : s->sizes) left.push_back(kv.first);
    for (float* p : left) s->release(p);
  }
  return rc;
}

// ---- whole-volume host call: one thread + one rank per device --------------------------------------------------
namespace {
struct HostGroup {
  size_t W = 0, H = 0, D = 0;
  std::vector<int> devices;
  std::vector<flow3d_sharded*> ranks;
};
std::mutex g_group_mu;
HostGroup* g_group = nullptr;

void release_group() {
  if (!g_group) return;
  std::vector<std::thread> th;
  for (flow3d_sharded* r : g_group->ranks) th.emplace_back([r] { flow3d_sharded_destroy(r); });
  for (auto& t : th) t.join();
  delete g_group;
  g_group = nullptr;
}
}  // namespace

int flow3d_mgpu_compute_host(size_t W, size_t H, size_t D, int n_devices, const int* devices, const float* frame_0,
                             const float* frame_1, const flow3d_params* params, float* flow_u, float* flow_v,
                             float* flow_w, float* ms_out, int persistent) {
  std::lock_guard<std::mutex> lk(g_group_mu);
  if (n_devices == 0) { release_group(); return FLOW3D_OK; }
  if (n_devices < 0 || !devices || !frame_0 || !frame_1 || !params || !flow_u || !flow_v || !flow_w)
    return FLOW3D_ERR_INVALID_ARG;
  const std::vector<int> devs(devices, devices + n_devices);
  if (g_group && (g_group->W != W || g_group->H != H || g_group->D != D || g_group->devices != devs)) release_group();
  std::vector<int> rcs(n_devices, FLOW3D_OK);
  if (!g_group) {
    char id[FLOW3D_MGPU_ID_BYTES];
    if (n_devices > 1) M_TRY(flow3d_mgpu_unique_id(id));
    HostGroup* g = new HostGroup();
    g->W = W; g->H = H; g->D = D; g->devices = devs;
    g->ranks.assign(n_devices, nullptr);
    std::vector<std::thread> th;
    for (int r = 0; r < n_devices; ++r)
      th.emplace_back([&, r] { rcs[r] = flow3d_sharded_create(W, H, D, devs[r], r, n_devices, id, &g->ranks[r]); });
    for (auto& t : th) t.join();
    g_group = g;
    for (int rc : rcs)
      if (rc != FLOW3D_OK) { release_group(); return rc; }
  }
  const size_t ghost = flow3d_sharded_frame_ghost(W, H, D, n_devices, params, g_group->ranks[0]->min_planes,
                                                  g_group->ranks[0]->min_voxels);
  if (ghost == 0) {
    std::fprintf(stderr, "flow3d_mgpu: %zux%zux%zu cannot be z-sharded over %d devices with these parameters (the ghost "
                 "depth inner_iterations_count + 1 must cover the median's reach: inner >= median_radius / 2 + 1)\n",
                 W, H, D, n_devices);
    return FLOW3D_ERR_INVALID_ARG;
  }
  const size_t ld = flow3d_aligned_ld(W);
  std::vector<float> ms(n_devices, 0.f);
  std::vector<std::thread> th;
  for (int r = 0; r < n_devices; ++r) {
    th.emplace_back([&, r] {
      flow3d_sharded* s = g_group->ranks[r];
      auto run = [&]() -> int {
        M_CUDA(cudaSetDevice(s->device));
        M_TRY(flow3d_sharded_tune(s, params));
        size_t lo, hi, a, b, pa, pb;
        flow3d_sharded_input_planes(D, r, n_devices, params->gaussian_sigma, ghost, &lo, &hi);
        flow3d_sharded_own_range(D, r, n_devices, &a, &b);
        M_TRY(flow3d_sharded_output_planes(s, params, &pa, &pb));
        const size_t cap = pb - pa;  // a replicated finest level returns every plane
        cudaStream_t st;
        cudaEvent_t e0, e1;
        M_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        M_CUDA(cudaEventCreate(&e0));
        M_CUDA(cudaEventCreate(&e1));
        float *r0 = nullptr, *r1 = nullptr, *o[3] = {nullptr, nullptr, nullptr};
        const size_t slab_bytes = (hi - lo) * H * ld * sizeof(float);
        M_CUDA(cudaMallocAsync((void**)&r0, slab_bytes, st));
        M_CUDA(cudaMallocAsync((void**)&r1, slab_bytes, st));
        for (int c = 0; c < 3; ++c) M_CUDA(cudaMallocAsync((void**)&o[c], cap * H * ld * sizeof(float), st));
        M_CUDA(cudaEventRecord(e0, st));
        M_CUDA(cudaMemcpy2DAsync(r0, ld * 4, frame_0 + lo * H * W, W * 4, W * 4, H * (hi - lo), cudaMemcpyHostToDevice, st));
        M_CUDA(cudaMemcpy2DAsync(r1, ld * 4, frame_1 + lo * H * W, W * 4, W * 4, H * (hi - lo), cudaMemcpyHostToDevice, st));
        size_t oa = 0, ob = 0;
        int rc = flow3d_sharded_compute(s, r0, r1, lo, hi - lo, ld, params, ghost, o[0], o[1], o[2], cap, &oa, &ob, st);
        if (rc == FLOW3D_OK) {
          // a replicated finest level leaves every plane on every rank: each rank still delivers its own share
          const size_t da = (ob - oa == D && n_devices > 1) ? a : oa, db = (ob - oa == D && n_devices > 1) ? b : ob;
          float* outs[3] = {flow_u, flow_v, flow_w};
          for (int c = 0; c < 3; ++c)
            M_CUDA(cudaMemcpy2DAsync(outs[c] + da * H * W, W * 4, o[c] + (da - oa) * H * ld, ld * 4, W * 4, H * (db - da),
                                     cudaMemcpyDeviceToHost, st));
        }
        M_CUDA(cudaEventRecord(e1, st));
        cudaFreeAsync(r0, st);
        cudaFreeAsync(r1, st);
        for (int c = 0; c < 3; ++c) cudaFreeAsync(o[c], st);
        M_CUDA(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ms[r], e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaStreamDestroy(st);
        return rc;
      };
      rcs[r] = run();
    });
  }
  for (auto& t : th) t.join();
  if (ms_out) *ms_out = *std::max_element(ms.begin(), ms.end());
  int rc = FLOW3D_OK;
  for (int x : rcs)
    if (x != FLOW3D_OK) rc = x;
  if (!persistent || rc != FLOW3D_OK) release_group();
  return rc;
}

}  // extern "C"
