// kernels_warp.cu -- backward trilinear warp of frame 1 and the image-derivative stencils, separate
// and fused, for sm_100a.  Arithmetic transcribed operation by operation from the reference PTX
// (registration_3d.cu:46-80; solve_3d.cu:425-438).
#include "common.cuh"

namespace f3d {

struct WarpGeom {
  Dims g;
  float rhx, rhy, rhz;  // rcp.rn(h): the reference evaluates (1.f / h) with a correctly rounded rcp
  float wmax, hmax, dmax;  // (float)(dim - 1), dmax from the GLOBAL depth
  int f1_z0g, f1_d;        // frame 1 may live in its own (taller) z-slab
};

// Warped frame-1 value at voxel (x,y,z): trilinear sample of f1 at x + u/h, or f0 when the target
// leaves the volume or is NaN.  Pure function of the voxel => the fused kernel may evaluate it at
// neighbours and get the very bits a separate warp pass would have stored.
__device__ __forceinline__ float warp_at(const float* __restrict__ f0, const float* __restrict__ f1,
                                         const float* __restrict__ u, const float* __restrict__ v,
                                         const float* __restrict__ w, const WarpGeom& q, int x, int y,
                                         int z) {
  const Dims& g = q.g;
  const unsigned c = (unsigned)z * (unsigned)g.ps + (unsigned)y * g.ld + x;  // < 2^32 elements per volume
  const float x_f = __fmaf_rn(q.rhx, __ldg(u + c), (float)(unsigned)x);
  const float y_f = __fmaf_rn(q.rhy, __ldg(v + c), (float)(unsigned)y);
  const float z_f = __fmaf_rn(q.rhz, __ldg(w + c), (float)(unsigned)(g.z0g + z));  // global plane
  if ((x_f < 0.f) || (x_f > q.wmax) || (y_f < 0.f) || (y_f > q.hmax) || (z_f < 0.f) ||
      (z_f > q.dmax) || isnan(x_f) || isnan(y_f) || isnan(z_f)) {
    return __ldg(f0 + c);
  }
  const float xfl = floorf(x_f), yfl = floorf(y_f), zfl = floorf(z_f);
  const int xi = (int)xfl, yi = (int)yfl, zi = (int)zfl;
  const float dx = __fsub_rn(x_f, (float)xi);
  const float dy = __fsub_rn(y_f, (float)yi);
  const float dz = __fsub_rn(z_f, (float)zi);
  const int x1 = min(g.w - 1, xi + 1);
  const int y1 = min(g.h - 1, yi + 1);
  const int z1g = min(g.dg - 1, zi + 1);
  // global -> local plane of the frame-1 slab; clamped so that an under-sized slab can never fault
  // (the caller sizes the slab from max|w|, see cuda_flow3d_b200/dist.py)
  const int z0l = min(max(zi - q.f1_z0g, 0), q.f1_d - 1);
  const int z1l = min(max(z1g - q.f1_z0g, 0), q.f1_d - 1);
  const float ox = __fsub_rn(1.f, dx), oy = __fsub_rn(1.f, dy);
  const float w00 = __fmul_rn(ox, oy);
  const float w10 = __fmul_rn(dx, oy);
  const float w01 = __fmul_rn(ox, dy);
  const float w11 = __fmul_rn(dx, dy);
  const unsigned p0 = (unsigned)z0l * (unsigned)g.ps, p1 = (unsigned)z1l * (unsigned)g.ps;
  const unsigned ra = (unsigned)yi * g.ld, rb = (unsigned)y1 * g.ld;
  const unsigned r00 = p0 + ra, r01 = p0 + rb, r10 = p1 + ra, r11 = p1 + rb;
  float v0 = __fmul_rn(w10, __ldg(f1 + r00 + x1));
  v0 = __fmaf_rn(w00, __ldg(f1 + r00 + xi), v0);
  v0 = __fmaf_rn(w01, __ldg(f1 + r01 + xi), v0);
  v0 = __fmaf_rn(w11, __ldg(f1 + r01 + x1), v0);
  float v1 = __fmul_rn(w10, __ldg(f1 + r10 + x1));
  v1 = __fmaf_rn(w00, __ldg(f1 + r10 + xi), v1);
  v1 = __fmaf_rn(w01, __ldg(f1 + r11 + xi), v1);
  v1 = __fmaf_rn(w11, __ldg(f1 + r11 + x1), v1);
  return __fmaf_rn(__fsub_rn(1.f, dz), v0, __fmul_rn(dz, v1));
}

static WarpGeom make_geom(Dims g, float hx, float hy, float hz, int f1_z0g = 0, int f1_d = -1) {
  WarpGeom q;
  q.g = g;
  q.f1_z0g = f1_z0g;
  q.f1_d = f1_d < 0 ? g.d : f1_d;
  q.rhx = 1.f / hx;  // IEEE division of 1 == rcp.rn
  q.rhy = 1.f / hy;
  q.rhz = 1.f / hz;
  q.wmax = (float)(g.w - 1);
  q.hmax = (float)(g.h - 1);
  q.dmax = (float)(g.dg - 1);
  return q;
}

__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ f0,
                                                   const float* __restrict__ f1,
                                                   const float* __restrict__ u,
                                                   const float* __restrict__ v,
                                                   const float* __restrict__ w, WarpGeom q,
                                                   float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z;
  if (x >= q.g.w || y >= q.g.h) return;
  out[(long long)z * q.g.ps + (long long)y * q.g.ld + x] = warp_at(f0, f1, u, v, w, q, x, y, z);
}

int launch_warp(const float* f0, const float* f1, const float* u, const float* v, const float* w,
                Dims g, float hx, float hy, float hz, float* out, cudaStream_t st) {
  dim3 block(32, 8, 1);
  dim3 grid((g.w + 31) / 32, (g.h + 7) / 8, g.d);
  warp_kernel<<<grid, block, 0, st>>>(f0, f1, u, v, w, make_geom(g, hx, hy, hz), out);
  count_launch();
  return check_launch("warp_kernel");
}

// fx = (((f0[p]-f0[m]) + f1w[p]) - f1w[m]) / (4h)   (solve_3d.cu:425-436), ft = f1w - f0 (:437-438)
__device__ __forceinline__ float deriv(float a_p, float a_m, float b_p, float b_m, ConstDiv four_h) {
  return div_const(__fsub_rn(__fadd_rn(__fsub_rn(a_p, a_m), b_p), b_m), four_h);
}

__global__ void __launch_bounds__(256) derivatives_kernel(const float* __restrict__ f0,
                                                          const float* __restrict__ f1w, Dims g,
                                                          float hx, float hy, float hz,
                                                          float* __restrict__ fx,
                                                          float* __restrict__ fy,
                                                          float* __restrict__ fz,
                                                          float* __restrict__ ft) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = blockIdx.z;
  if (x >= g.w || y >= g.h) return;
  const long long c = (long long)z * g.ps + (long long)y * g.ld + x;
  const long long ixp = c - x + mirror_idx(x + 1, g.w), ixm = c - x + mirror_idx(x - 1, g.w);
  const long long rb = (long long)z * g.ps + x;
  const long long iyp = rb + (long long)mirror_idx(y + 1, g.h) * g.ld;
  const long long iym = rb + (long long)mirror_idx(y - 1, g.h) * g.ld;
  const long long cb = (long long)y * g.ld + x;
  const long long izp = cb + (long long)z_neighbour(g, z, 1) * g.ps;
  const long long izm = cb + (long long)z_neighbour(g, z, -1) * g.ps;
  fx[c] = deriv(__ldg(f0 + ixp), __ldg(f0 + ixm), __ldg(f1w + ixp), __ldg(f1w + ixm), make_const_div(__fmul_rn(hx, 4.f)));
  fy[c] = deriv(__ldg(f0 + iyp), __ldg(f0 + iym), __ldg(f1w + iyp), __ldg(f1w + iym), make_const_div(__fmul_rn(hy, 4.f)));
  fz[c] = deriv(__ldg(f0 + izp), __ldg(f0 + izm), __ldg(f1w + izp), __ldg(f1w + izm), make_const_div(__fmul_rn(hz, 4.f)));
  ft[c] = __fsub_rn(__ldg(f1w + c), __ldg(f0 + c));
}

int launch_derivatives(const float* f0, const float* f1w, Dims g, float hx, float hy, float hz,
                       float* fx, float* fy, float* fz, float* ft, cudaStream_t st) {
  dim3 block(32, 8, 1);
  dim3 grid((g.w + 31) / 32, (g.h + 7) / 8, g.d);
  derivatives_kernel<<<grid, block, 0, st>>>(f0, f1w, g, hx, hy, hz, fx, fy, fz, ft);
  count_launch();
  return check_launch("derivatives_kernel");
}

// Fused warp + derivatives: a CTA owns a 32x8 (x,y) tile and marches along z keeping three planes
// of warped values (tile + 1-voxel xy halo) in shared memory, so every warped value is gathered
// once per tile (+ halo) and the warped volume never goes to HBM.
#define WD_TX 32
#define WD_TY 8
__global__ void __launch_bounds__(WD_TX* WD_TY) warp_derivatives_kernel(
    const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ u,
    const float* __restrict__ v, const float* __restrict__ w, WarpGeom q, float hx, float hy,
    float hz, int zchunk, int zs, int ze, float* __restrict__ fx, float* __restrict__ fy, float* __restrict__ fz,
    float* __restrict__ ft) {
  const Dims& g = q.g;
  __shared__ float sw[3][WD_TY + 2][WD_TX + 2];  // ring of warped planes
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * WD_TX + tx;
  const int bx = blockIdx.x * WD_TX, by = blockIdx.y * WD_TY;
  const int x = bx + tx, y = by + ty;
  const int z_begin = zs + blockIdx.z * zchunk;
  const int z_end = min(ze, z_begin + zchunk);
  if (z_begin >= z_end) return;
  const bool valid = (x < g.w) && (y < g.h);
  const ConstDiv fhx = make_const_div(__fmul_rn(hx, 4.f)), fhy = make_const_div(__fmul_rn(hy, 4.f)),
                 fhz = make_const_div(__fmul_rn(hz, 4.f));
  constexpr int CELLS = (WD_TY + 2) * (WD_TX + 2);
  constexpr int PER = (CELLS + WD_TX * WD_TY - 1) / (WD_TX * WD_TY);  // ring cells per thread (2)

  // Everything that does not change along the march is computed once per thread: the (mirrored) voxel
  // coordinates of the ring cells this thread fills -- cell (r,cx) <-> voxel (mirror(bx-1+cx), mirror(by-1+r)):
  // mirrored coordinates reproduce the reference's halo -- and the in-plane offsets of the derivative
  // stencil.  Offsets are 32-bit (a volume holds < 2^32 elements, checked by the launcher's callers).
  int cell_x[PER], cell_y[PER], cell_s[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = tid + k * WD_TX * WD_TY;
    const int r = i / (WD_TX + 2), cx = i - r * (WD_TX + 2);
    cell_x[k] = mirror_idx(bx - 1 + cx, g.w);
    cell_y[k] = mirror_idx(by - 1 + r, g.h);
    cell_s[k] = (i < CELLS) ? r * (WD_TX + 2) + cx : -1;
  }
  float* const ring = &sw[0][0][0];
  auto fill = [&](int slot, int zz) {  // zz = local plane index, possibly one beyond a global face
    const int zsrc = z_neighbour(g, zz, 0);
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (cell_s[k] >= 0) ring[slot * CELLS + cell_s[k]] = warp_at(f0, f1, u, v, w, q, cell_x[k], cell_y[k], zsrc);
  };
  const unsigned ps = (unsigned)g.ps;
  const unsigned o_c = (unsigned)y * g.ld + x;
  const unsigned o_xp = (unsigned)y * g.ld + mirror_idx(x + 1, g.w), o_xm = (unsigned)y * g.ld + mirror_idx(x - 1, g.w);
  const unsigned o_yp = (unsigned)mirror_idx(y + 1, g.h) * g.ld + x, o_ym = (unsigned)mirror_idx(y - 1, g.h) * g.ld + x;
  fill(0, z_begin - 1);
  fill(1, z_begin);
  int sp = 0, sc = 1, sn = 2;
  for (int z = z_begin; z < z_end; ++z) {
    fill(sn, z + 1);
    __syncthreads();
    if (valid) {
      const unsigned pl = (unsigned)z * ps;
      const unsigned c = pl + o_c;
      const unsigned izp = (unsigned)z_neighbour(g, z, 1) * ps + o_c;
      const unsigned izm = (unsigned)z_neighbour(g, z, -1) * ps + o_c;
      const float wc = sw[sc][ty + 1][tx + 1];
      fx[c] = deriv(__ldg(f0 + pl + o_xp), __ldg(f0 + pl + o_xm), sw[sc][ty + 1][tx + 2], sw[sc][ty + 1][tx], fhx);
      fy[c] = deriv(__ldg(f0 + pl + o_yp), __ldg(f0 + pl + o_ym), sw[sc][ty + 2][tx + 1], sw[sc][ty][tx + 1], fhy);
      fz[c] = deriv(__ldg(f0 + izp), __ldg(f0 + izm), sw[sn][ty + 1][tx + 1], sw[sp][ty + 1][tx + 1], fhz);
      ft[c] = __fsub_rn(wc, __ldg(f0 + c));
    }
    __syncthreads();
    const int t = sp; sp = sc; sc = sn; sn = t;
  }
}

int launch_warp_derivatives(const float* f0, const float* f1, int f1_z0g, int f1_d, const float* u,
                            const float* v, const float* w, Dims g, ZRange zr, float hx, float hy,
                            float hz, float* fx, float* fy, float* fz, float* ft, cudaStream_t st) {
  if (zr.end <= zr.begin) return FLOW3D_OK;
  const int nz = zr.end - zr.begin;
  dim3 block(WD_TX, WD_TY, 1);
  const int gx = (g.w + WD_TX - 1) / WD_TX, gy = (g.h + WD_TY - 1) / WD_TY;
  const long long per_plane = (long long)gx * gy;
  long long nchunks = ((long long)sm_count() * 8 + per_plane - 1) / per_plane;
  if (nchunks < 1) nchunks = 1;
  long long len = (nz + nchunks - 1) / nchunks;
  if (len < 8) len = 8;
  if (len > nz) len = nz;
  dim3 grid(gx, gy, (unsigned)((nz + len - 1) / len));
  warp_derivatives_kernel<<<grid, block, 0, st>>>(f0, f1, u, v, w, make_geom(g, hx, hy, hz, f1_z0g, f1_d),
                                                   hx, hy, hz, (int)len, zr.begin, zr.end, fx, fy, fz, ft);
  count_launch();
  return check_launch("warp_derivatives_kernel");
}

}  // namespace f3d
