// kernels_solve.cu -- robust weights (phi, ksi), Jacobi sweep and flow update for sm_100a.
//
// Arithmetic contract ("strict"): every floating-point operation below is an explicit
// round-to-nearest intrinsic (__fmaf_rn / __fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn /
// __fsqrt_rn / __frcp_rn) placed exactly where the reference's kernels (nvcc -ptx of
// src/kernels/solve_3d.cu:177-260 and :423-507) perform an fma.rn / mul / add / div.rn / sqrt.rn /
// rcp.rn, so results are bit-identical to the reference's CUDA build while the kernel structure
// (z-marching warps, vector loads, register-rotated z neighbours, shuffle x neighbours, per-level
// precomputed image derivatives) is new.  A sensitivity study shows that merely changing
// FMA contraction moves the final 128^3 flow by up to 9e-3 voxel, so this is what the 1e-3 gate needs.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "solve_args.cuh"

namespace f3d {

#define F3D_TRY_RC(expr)                \
  do {                                 \
    int rc_ = (expr);                  \
    if (rc_ != FLOW3D_OK) return rc_;  \
  } while (0)

// ------------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------------
template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
};
template <>
struct Vec<2> {
  float v[2];
};
template <>
struct Vec<4> {
  float v[4];
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv(const float* __restrict__ p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void stv(float* p, const Vec<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

// L2 prefetch of the cache line(s) this thread will load `pf` planes ahead: converts the DRAM
// latency of the z-march into an L2 hit without holding registers for data in flight.
__device__ __forceinline__ void prefetch_l2(const float* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const float* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

template <int VEC>
__device__ __forceinline__ Vec<VEC> addv(const Vec<VEC>& a, const Vec<VEC>& b) {
  Vec<VEC> r;
#pragma unroll
  for (int i = 0; i < VEC; ++i) r.v[i] = __fadd_rn(a.v[i], b.v[i]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// Jacobi sweep
// ------------------------------------------------------------------------------------------------
// Lane layout: a warp covers 32/lpr consecutive rows x (lpr*VEC) columns (lpr = lanes per row, a power
// of two chosen per level so that ceil(w / (lpr*VEC)) tiles waste as few lanes as possible -- level
// widths like 397 would otherwise leave a quarter of the lanes idle).
struct LaneMap {
  int sub;         // lane index within its row segment
  int x0, y;       // first column of this lane's VEC group (clamped for idle lanes), row
  bool active;     // false: idle lane (x beyond the volume), shadows the last group, stores nothing
  bool row_ok;     // false: row beyond the volume (idle, clamped to the last row)
  bool left_edge, right_edge;
  int tile_x, first_row;  // this warp's x tile index and first row
};
// wx = warps of a block placed side by side along x (the rest stack along y): wx > 1 makes a block
// read longer contiguous runs of every row
template <int VEC>
__device__ __forceinline__ LaneMap lane_map(int lpr, int w, int h, int wx = 1) {
  LaneMap m;
  const int lane = threadIdx.x;
  m.sub = lane & (lpr - 1);
  const int rows_per_warp = 32 / lpr;
  const int warp_x = threadIdx.y % wx, warp_y = threadIdx.y / wx;
  const int wy = blockDim.y / wx;
  m.tile_x = blockIdx.x * wx + warp_x;
  m.first_row = (blockIdx.y * wy + warp_y) * rows_per_warp;
  const int y_raw = m.first_row + lane / lpr;
  m.row_ok = y_raw < h;
  m.y = m.row_ok ? y_raw : h - 1;
  const int x0_raw = (m.tile_x * lpr + m.sub) * VEC;
  m.active = (x0_raw < w) && m.row_ok;
  m.x0 = (x0_raw < w) ? x0_raw : ((w - 1) / VEC) * VEC;
  m.left_edge = m.sub == 0;
  m.right_edge = m.sub == lpr - 1;
  return m;
}

// x-neighbour values of a VEC-wide register group: left[i] / right[i] are the values at x-1 / x+1
// of element i, taken from the group itself, the adjacent lanes (shuffle) or, at the ends of the
// lane's row segment, from `halo` (first lane: value at x0-1, last lane: value at x0+VEC).  At the volume
// faces the reflect-101 neighbour is substituted (x=0 -> value at 1, x=w-1 -> value at w-2), which
// is what the reference's shared-memory halo holds (solve_3d.cu:326-355).
template <int VEC>
__device__ __forceinline__ void x_neighbours(const Vec<VEC>& c, float halo, bool left_edge, bool right_edge,
                                             int x0, int w, Vec<VEC>& left, Vec<VEC>& right) {
  float from_left = __shfl_up_sync(0xffffffffu, c.v[VEC - 1], 1);
  float from_right = __shfl_down_sync(0xffffffffu, c.v[0], 1);
  if (left_edge) from_left = halo;
  if (right_edge) from_right = halo;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float l = (i > 0) ? c.v[i - 1] : from_left;
    float r = (i < VEC - 1) ? c.v[i + 1] : from_right;
    const int x = x0 + i;
    left.v[i] = (x == 0) ? r : l;
    right.v[i] = (x == w - 1) ? l : r;
  }
}

// Each warp owns 32/lpr row segments of lpr*VEC voxels and marches through a chunk of z planes.
// S = u + du (one rounded add per voxel, shared by all six consumers of that voxel) and phi are kept
// in registers for three consecutive planes; y neighbours are read straight from global memory
// (L1-resident: they are the centre rows of the adjacent warps of the same CTA).
//
// The three planes live in three register sets whose ROLES (previous / current / next) rotate through
// an explicitly 3x-unrolled loop, so the rotation costs no register moves; and the plane body is
// compiled twice, for warps whose tile touches an x face of the volume (EDGE: mirror selects, zeroed
// face weights) and for interior warps (no selects at all).
template <int VEC>
struct PlaneRegs {
  Vec<VEC> Su, Sv, Sw, ph;  // live while the plane is next / current / previous
  Vec<VEC> u, v, w, dv, dw; // live while the plane is next / current
  Vec<VEC> du;              // KSI variant only (dead code otherwise)
};

template <int VEC>
struct SweepCtx {
  unsigned ps, row_c, row_m, row_p, row_h;
  float hx2, hz2, wyp, wym;
  int x0, w;
  bool left_edge, right_edge, active;
};

template <int VEC>
__device__ __forceinline__ void load_plane(const SweepArgs& a, unsigned o, PlaneRegs<VEC>& r) {
  r.u = ldv<VEC>(a.u + o);
  r.v = ldv<VEC>(a.v + o);
  r.w = ldv<VEC>(a.w + o);
  r.du = ldv<VEC>(a.du + o);
  r.dv = ldv<VEC>(a.dv + o);
  r.dw = ldv<VEC>(a.dw + o);
  r.ph = ldv<VEC>(a.phi + o);
  r.Su = addv<VEC>(r.u, r.du);
  r.Sv = addv<VEC>(r.v, r.dv);
  r.Sw = addv<VEC>(r.w, r.dw);
}

// x neighbours of an interior tile: no volume face inside the warp's tile, plain shuffles + halo
template <int VEC>
__device__ __forceinline__ void x_neighbours_interior(const Vec<VEC>& c, float halo, bool left_edge,
                                                      bool right_edge, Vec<VEC>& left, Vec<VEC>& right) {
  float from_left = __shfl_up_sync(0xffffffffu, c.v[VEC - 1], 1);
  float from_right = __shfl_down_sync(0xffffffffu, c.v[0], 1);
  if (left_edge) from_left = halo;
  if (right_edge) from_right = halo;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    left.v[i] = (i > 0) ? c.v[i - 1] : from_left;
    right.v[i] = (i < VEC - 1) ? c.v[i + 1] : from_right;
  }
}

// one plane of the sweep: P = previous plane, C = current, N = receives plane z+1
template <int VEC, bool EDGE, bool KSI>
__device__ __forceinline__ void sweep_plane(const SweepArgs& a, const Dims& g, const SweepCtx<VEC>& c, int z,
                                            const PlaneRegs<VEC>& P, const PlaneRegs<VEC>& C, PlaneRegs<VEC>& N) {
  const unsigned ps = c.ps;
  const unsigned pl = (unsigned)z * ps;
  if (a.pf > 0) {
    const int zp1 = z + 1 + a.pf;  // stencil fields are consumed one plane ahead
    if (zp1 < g.d) {
      const unsigned o = (unsigned)zp1 * ps + c.row_c;
      prefetch_l2(a.u + o); prefetch_l2(a.v + o); prefetch_l2(a.w + o);
      prefetch_l2(a.du + o); prefetch_l2(a.dv + o); prefetch_l2(a.dw + o);
      prefetch_l2(a.phi + o);
    }
    const int zp0 = z + a.pf;
    if (zp0 < g.d) {
      const unsigned o = (unsigned)zp0 * ps + c.row_c;
      prefetch_l2(a.fx + o); prefetch_l2(a.fy + o); prefetch_l2(a.fz + o);
      prefetch_l2(a.ft + o);
      if constexpr (!KSI) prefetch_l2(a.ksi + o);
    }
  }
  // ---- next plane (reflect at the rear face) --------------------------------------------------
  load_plane<VEC>(a, (unsigned)z_neighbour(g, z, 1) * ps + c.row_c, N);
  // ---- centre-only fields of the current plane ------------------------------------------------
  const unsigned oc = pl + c.row_c;
  const Vec<VEC> fx = ldv<VEC>(a.fx + oc);
  const Vec<VEC> fy = ldv<VEC>(a.fy + oc);
  const Vec<VEC> fz = ldv<VEC>(a.fz + oc);
  const Vec<VEC> ft = ldv<VEC>(a.ft + oc);
  Vec<VEC> ks;
  if constexpr (!KSI) ks = ldv<VEC>(a.ksi + oc);
  // ---- y neighbours of the current plane -------------------------------------------------------
  const unsigned om = pl + c.row_m;
  const unsigned op = pl + c.row_p;
  const Vec<VEC> Su_ym = addv<VEC>(ldv<VEC>(a.u + om), ldv<VEC>(a.du + om));
  const Vec<VEC> Sv_ym = addv<VEC>(ldv<VEC>(a.v + om), ldv<VEC>(a.dv + om));
  const Vec<VEC> Sw_ym = addv<VEC>(ldv<VEC>(a.w + om), ldv<VEC>(a.dw + om));
  const Vec<VEC> ph_ym = ldv<VEC>(a.phi + om);
  const Vec<VEC> Su_yp = addv<VEC>(ldv<VEC>(a.u + op), ldv<VEC>(a.du + op));
  const Vec<VEC> Sv_yp = addv<VEC>(ldv<VEC>(a.v + op), ldv<VEC>(a.dv + op));
  const Vec<VEC> Sw_yp = addv<VEC>(ldv<VEC>(a.w + op), ldv<VEC>(a.dw + op));
  const Vec<VEC> ph_yp = ldv<VEC>(a.phi + op);
  // ---- x halo: every lane loads (no branch, so these loads are issued with the batch above: one
  // memory round trip per plane); only the first / last lane of a row segment uses the value
  const unsigned oh = pl + c.row_h;
  const float hSu = __fadd_rn(__ldg(a.u + oh), __ldg(a.du + oh));
  const float hSv = __fadd_rn(__ldg(a.v + oh), __ldg(a.dv + oh));
  const float hSw = __fadd_rn(__ldg(a.w + oh), __ldg(a.dw + oh));
  const float hph = __ldg(a.phi + oh);
  Vec<VEC> Su_xm, Su_xp, Sv_xm, Sv_xp, Sw_xm, Sw_xp, ph_xm, ph_xp;
  if constexpr (EDGE) {
    x_neighbours<VEC>(C.Su, hSu, c.left_edge, c.right_edge, c.x0, c.w, Su_xm, Su_xp);
    x_neighbours<VEC>(C.Sv, hSv, c.left_edge, c.right_edge, c.x0, c.w, Sv_xm, Sv_xp);
    x_neighbours<VEC>(C.Sw, hSw, c.left_edge, c.right_edge, c.x0, c.w, Sw_xm, Sw_xp);
    x_neighbours<VEC>(C.ph, hph, c.left_edge, c.right_edge, c.x0, c.w, ph_xm, ph_xp);
  } else {
    x_neighbours_interior<VEC>(C.Su, hSu, c.left_edge, c.right_edge, Su_xm, Su_xp);
    x_neighbours_interior<VEC>(C.Sv, hSv, c.left_edge, c.right_edge, Sv_xm, Sv_xp);
    x_neighbours_interior<VEC>(C.Sw, hSw, c.left_edge, c.right_edge, Sw_xm, Sw_xp);
    x_neighbours_interior<VEC>(C.ph, hph, c.left_edge, c.right_edge, ph_xm, ph_xp);
  }

  const int zg = g.z0g + z;  // faces are the GLOBAL ones when the level is sharded
  const float wzp = (zg < g.dg - 1) ? c.hz2 : 0.f;
  const float wzm = (zg > 0) ? c.hz2 : 0.f;

  Vec<VEC> rdu, rdv, rdw;
  // phase 1: everything up to the three dependent divisions, for the lane's VEC voxels
  float numU[VEC], denU[VEC], denV[VEC], denW[VEC], sV[VEC], sW[VEC], j12[VEC], j13[VEC], j23[VEC], j24[VEC], j34[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float wxp = c.hx2, wxm = c.hx2;
    if constexpr (EDGE) {
      const int x = c.x0 + i;
      wxp = (x < c.w - 1) ? c.hx2 : 0.f;
      wxm = (x > 0) ? c.hx2 : 0.f;
    }
    const float J11 = __fmul_rn(fx.v[i], fx.v[i]);
    const float J22 = __fmul_rn(fy.v[i], fy.v[i]);
    const float J33 = __fmul_rn(fz.v[i], fz.v[i]);
    const float J12 = __fmul_rn(fx.v[i], fy.v[i]);
    const float J13 = __fmul_rn(fx.v[i], fz.v[i]);
    const float J23 = __fmul_rn(fy.v[i], fz.v[i]);
    const float J14 = __fmul_rn(fx.v[i], ft.v[i]);
    const float J24 = __fmul_rn(fy.v[i], ft.v[i]);
    const float J34 = __fmul_rn(fz.v[i], ft.v[i]);
    const float pc = C.ph.v[i];
    // face weights: w * (phi_n + phi_c)/2  (solve_3d.cu:462-469); plain products
    const float axp = __fmul_rn(wxp, __fmul_rn(__fadd_rn(ph_xp.v[i], pc), 0.5f));
    const float axm = __fmul_rn(wxm, __fmul_rn(__fadd_rn(ph_xm.v[i], pc), 0.5f));
    const float ayp = __fmul_rn(c.wyp, __fmul_rn(__fadd_rn(ph_yp.v[i], pc), 0.5f));
    const float aym = __fmul_rn(c.wym, __fmul_rn(__fadd_rn(ph_ym.v[i], pc), 0.5f));
    const float azp = __fmul_rn(wzp, __fmul_rn(__fadd_rn(N.ph.v[i], pc), 0.5f));
    const float azm = __fmul_rn(wzm, __fmul_rn(__fadd_rn(P.ph.v[i], pc), 0.5f));
    const float sumH = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(axp, axm), ayp), aym), azp), azm);
    const float uc = C.u.v[i], vc = C.v.v[i], wc = C.w.v[i];
    // solve_3d.cu:470-490: seed with the rounded x- product, then one fma per face
    float sumU = __fmul_rn(axm, __fsub_rn(Su_xm.v[i], uc));
    sumU = __fmaf_rn(axp, __fsub_rn(Su_xp.v[i], uc), sumU);
    sumU = __fmaf_rn(ayp, __fsub_rn(Su_yp.v[i], uc), sumU);
    sumU = __fmaf_rn(aym, __fsub_rn(Su_ym.v[i], uc), sumU);
    sumU = __fmaf_rn(azp, __fsub_rn(N.Su.v[i], uc), sumU);
    sumU = __fmaf_rn(azm, __fsub_rn(P.Su.v[i], uc), sumU);
    float sumV = __fmul_rn(axm, __fsub_rn(Sv_xm.v[i], vc));
    sumV = __fmaf_rn(axp, __fsub_rn(Sv_xp.v[i], vc), sumV);
    sumV = __fmaf_rn(ayp, __fsub_rn(Sv_yp.v[i], vc), sumV);
    sumV = __fmaf_rn(aym, __fsub_rn(Sv_ym.v[i], vc), sumV);
    sumV = __fmaf_rn(azp, __fsub_rn(N.Sv.v[i], vc), sumV);
    sumV = __fmaf_rn(azm, __fsub_rn(P.Sv.v[i], vc), sumV);
    float sumW = __fmul_rn(axm, __fsub_rn(Sw_xm.v[i], wc));
    sumW = __fmaf_rn(axp, __fsub_rn(Sw_xp.v[i], wc), sumW);
    sumW = __fmaf_rn(ayp, __fsub_rn(Sw_yp.v[i], wc), sumW);
    sumW = __fmaf_rn(aym, __fsub_rn(Sw_ym.v[i], wc), sumW);
    sumW = __fmaf_rn(azp, __fsub_rn(N.Sw.v[i], wc), sumW);
    sumW = __fmaf_rn(azm, __fsub_rn(P.Sw.v[i], wc), sumW);
    // solve_3d.cu:492-502; numerators as the reference's SASS evaluates them (ptxas contracts the
    // PTX's mul+sub pairs): n = fma(-J13, dw, fma(-J12, dv, -J14))
    if constexpr (KSI) {
      // data-term weight of this outer iteration, from the iterate the iteration starts with: the
      // xi half of compute_phi_ksi_3d (solve_3d.cu:250-260), operation for operation as in
      // phi_ksi_kernel below; stored for the remaining sweeps of the iteration
      const float gt = ft.v[i];
      const float du = C.du.v[i], dv = C.dv.v[i], dw = C.dw.v[i];
      const float r1 = __fadd_rn(J14, __fmaf_rn(J13, dw, __fmaf_rn(J11, du, __fmul_rn(J12, dv))));
      const float r2 = __fadd_rn(J24, __fmaf_rn(J23, dw, __fmaf_rn(J12, du, __fmul_rn(J22, dv))));
      const float r3 = __fadd_rn(J34, __fmaf_rn(J33, dw, __fmaf_rn(J23, dv, __fmul_rn(J13, du))));
      const float r4 = __fmaf_rn(gt, gt, __fmaf_rn(J34, dw, __fmaf_rn(J14, du, __fmul_rn(J24, dv))));
      float sv = __fadd_rn(__fmaf_rn(dw, r3, __fmaf_rn(du, r1, __fmul_rn(dv, r2))), r4);
      sv = __fmul_rn(sv, (sv > 0.f) ? 1.f : 0.f);
      const float arg = __fmaf_rn(a.eps_d, a.eps_d, sv);
      bool ok2 = true;
      float sq2 = sqrt_fast(arg, ok2);
      float kk = rcp_fast(__fadd_rn(sq2, sq2), ok2);
      if (!ok2) {
        sq2 = __fsqrt_rn(arg);
        kk = __frcp_rn(__fadd_rn(sq2, sq2));
      }
      ks.v[i] = kk;
    }
    const float k = ks.v[i];
    const float ndu = __fmaf_rn(-J13, C.dw.v[i], __fmaf_rn(-J12, C.dv.v[i], -J14));
    numU[i] = __fmaf_rn(k, ndu, sumU);
    denU[i] = __fmaf_rn(J11, k, sumH);
    denV[i] = __fmaf_rn(J22, k, sumH);
    denW[i] = __fmaf_rn(J33, k, sumH);
    sV[i] = sumV; sW[i] = sumW;
    j12[i] = J12; j13[i] = J13; j23[i] = J23; j24[i] = J24; j34[i] = J34;
  }
  // phase 2: du' -> dv' -> dw' (each needs the previous quotient) as three rounds over the VEC voxels,
  // branch-free (common.cuh: div_fast), so the dependent chains of the voxels overlap
  bool ok[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { ok[i] = true; rdu.v[i] = div_fast(numU[i], denU[i], ok[i]); }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float ndv = __fmaf_rn(-j23[i], C.dw.v[i], __fmaf_rn(-j12[i], rdu.v[i], -j24[i]));
    rdv.v[i] = div_fast(__fmaf_rn(ks.v[i], ndv, sV[i]), denV[i], ok[i]);
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float ndw = __fmaf_rn(-j23[i], rdv.v[i], __fmaf_rn(-j13[i], rdu.v[i], -j34[i]));
    rdw.v[i] = div_fast(__fmaf_rn(ks.v[i], ndw, sW[i]), denW[i], ok[i]);
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (!ok[i]) {  // an operand outside the fast path's range (zero, tiny, huge, NaN): IEEE division
      rdu.v[i] = __fdiv_rn(numU[i], denU[i]);
      const float ndv = __fmaf_rn(-j23[i], C.dw.v[i], __fmaf_rn(-j12[i], rdu.v[i], -j24[i]));
      rdv.v[i] = __fdiv_rn(__fmaf_rn(ks.v[i], ndv, sV[i]), denV[i]);
      const float ndw = __fmaf_rn(-j23[i], rdv.v[i], __fmaf_rn(-j13[i], rdu.v[i], -j34[i]));
      rdw.v[i] = __fdiv_rn(__fmaf_rn(ks.v[i], ndw, sW[i]), denW[i]);
    }
  }
  if (c.active) {
    stv<VEC>(a.odu + oc, rdu);
    stv<VEC>(a.odv + oc, rdv);
    stv<VEC>(a.odw + oc, rdw);
    if constexpr (KSI) stv<VEC>(a.oksi + oc, ks);
  }
}

template <int VEC, bool EDGE, bool KSI>
__device__ __forceinline__ void sweep_march(const SweepArgs& a, const Dims& g, const SweepCtx<VEC>& c, int z_begin,
                                            int z_end) {
  PlaneRegs<VEC> A, B, Cc;
  // prologue: plane z_begin-1 (reflected at the front face) and plane z_begin
  load_plane<VEC>(a, (unsigned)z_neighbour(g, z_begin, -1) * c.ps + c.row_c, A);
  load_plane<VEC>(a, (unsigned)z_begin * c.ps + c.row_c, B);
  for (int z = z_begin; z < z_end; z += 3) {
    sweep_plane<VEC, EDGE, KSI>(a, g, c, z, A, B, Cc);
    if (z + 1 >= z_end) break;
    sweep_plane<VEC, EDGE, KSI>(a, g, c, z + 1, B, Cc, A);
    if (z + 2 >= z_end) break;
    sweep_plane<VEC, EDGE, KSI>(a, g, c, z + 2, Cc, A, B);
  }
}

// single-body loop: the roles rotate by copying (the compiler turns most copies into renames)
template <int VEC, bool EDGE, bool KSI>
__device__ __forceinline__ void sweep_march1(const SweepArgs& a, const Dims& g, const SweepCtx<VEC>& c, int z_begin,
                                             int z_end) {
  PlaneRegs<VEC> P, C, N;
  load_plane<VEC>(a, (unsigned)z_neighbour(g, z_begin, -1) * c.ps + c.row_c, P);
  load_plane<VEC>(a, (unsigned)z_begin * c.ps + c.row_c, C);
#pragma unroll 1
  for (int z = z_begin; z < z_end; ++z) {
    sweep_plane<VEC, EDGE, KSI>(a, g, c, z, P, C, N);
    P.Su = C.Su; P.Sv = C.Sv; P.Sw = C.Sw; P.ph = C.ph;
    C = N;
  }
}

template <int VEC, int MINB, int ROT, int SPEC, bool KSI = false>
__global__ void __launch_bounds__(128, MINB) sweep_kernel(const SweepArgs a) {
  pdl_trigger();  // the next launch of the chain may become resident while this grid drains ...
  pdl_wait();     // ... and this one touches memory only after its predecessor has completed (common.cuh)
  const Dims g = a.g;
  const LaneMap lm = lane_map<VEC>(a.lpr, g.w, g.h, a.wx);
  if (lm.first_row >= g.h || lm.tile_x * a.lpr * VEC >= g.w) return;  // whole warp leaves together
  const int y = lm.y;
  const int x0 = lm.x0;  // idle lanes shadow the last group / last row
  const int z_begin = a.zs + blockIdx.z * a.zchunk;
  const int z_end = min(a.ze, z_begin + a.zchunk);
  if (z_begin >= z_end) return;

  // halo column of this lane: first lane of a row segment -> x0-1, last -> x0+VEC (mirrored at the
  // faces); the other lanes re-read their own first element (an L1 hit) so the load needs no branch
  const int xh = lm.left_edge ? mirror_idx(x0 - 1, g.w) : (lm.right_edge ? mirror_idx(x0 + VEC, g.w) : x0);

  SweepCtx<VEC> c;
  // 32-bit element offsets (volumes hold < 2^32 elements; checked by the launcher): one
  // IMAD.WIDE per load instead of a 64-bit add pair
  c.ps = (unsigned)g.ps;
  c.row_c = (unsigned)y * g.ld + x0;
  c.row_m = (unsigned)mirror_idx(y - 1, g.h) * g.ld + x0;
  c.row_p = (unsigned)mirror_idx(y + 1, g.h) * g.ld + x0;
  c.row_h = (unsigned)y * g.ld + xh;
  // weights (solve_3d.cu:451-460): alpha / (h*h), zeroed at the volume faces
  c.hx2 = __fdiv_rn(a.alpha, __fmul_rn(a.hx, a.hx));
  const float hy2 = __fdiv_rn(a.alpha, __fmul_rn(a.hy, a.hy));
  c.hz2 = __fdiv_rn(a.alpha, __fmul_rn(a.hz, a.hz));
  c.wyp = (y < g.h - 1) ? hy2 : 0.f;
  c.wym = (y > 0) ? hy2 : 0.f;
  c.x0 = x0;
  c.w = g.w;
  c.left_edge = lm.left_edge;
  c.right_edge = lm.right_edge;
  c.active = lm.active;

  // does this warp's tile [tx0, tx1) contain x = 0 or x = w-1 (or run past the volume)?
  const int tx0 = lm.tile_x * a.lpr * VEC, tx1 = tx0 + a.lpr * VEC;
  const bool edge = (tx0 == 0) || (tx1 > g.w - 1);  // warp-uniform
  if constexpr (ROT) {
    if (edge || !SPEC) sweep_march<VEC, true, KSI>(a, g, c, z_begin, z_end);
    else sweep_march<VEC, false, KSI>(a, g, c, z_begin, z_end);
  } else {
    if (edge || !SPEC) sweep_march1<VEC, true, KSI>(a, g, c, z_begin, z_end);
    else sweep_march1<VEC, false, KSI>(a, g, c, z_begin, z_end);
  }
}

static thread_local bool g_pdl_active = false;
static std::atomic<bool> g_pdl_refused{false};
static std::atomic<int> g_pdl_mode{-1};  // -1: $FLOW3D_PDL (default on), 0 off, 1 on  (flow3d_set_pdl)
static bool pdl_env_default() {
  static const bool enabled = [] {
    const char* e = getenv("FLOW3D_PDL");
    return !(e && *e) || atoi(e) != 0;
  }();
  return enabled;
}
bool pdl_wanted() {
  const int m = g_pdl_mode.load(std::memory_order_relaxed);
  return (m < 0 ? pdl_env_default() : m != 0) && !g_pdl_refused.load(std::memory_order_relaxed);
}
int pdl_set_mode(int mode) {
  g_pdl_mode.store(mode < 0 ? -1 : (mode != 0), std::memory_order_relaxed);
  return pdl_wanted() ? 1 : 0;
}
bool pdl_active() { return g_pdl_active && pdl_wanted(); }
void pdl_disable() { g_pdl_refused.store(true, std::memory_order_relaxed); }
PdlScope::PdlScope(bool on) : prev(g_pdl_active) { g_pdl_active = on; }
PdlScope::~PdlScope() { g_pdl_active = prev; }

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

// lanes per row segment: the power of two in {32,16,8} that wastes the fewest lanes for this width
static int pick_lpr(int w, int vec) {
  int best = 32;
  long long best_cols = -1;
  for (int lpr = 32; lpr >= 8; lpr >>= 1) {
    const int tile = lpr * vec;
    const long long cols = (long long)((w + tile - 1) / tile) * tile;
    if (best_cols < 0 || cols < best_cols) { best = lpr; best_cols = cols; }
  }
  return best;
}

// z chunking of a z-marching launch.  `resident` = CTAs of this kernel that fit on one SM.  Each CTA
// marches `len` planes after a prologue of two; CTAs run in waves of sm_count*resident, and a partly
// filled last wave costs as much as a full one, so the chunk count is chosen to minimise
// waves * (len + prologue) instead of aiming at a fixed CTA count (512^3, VEC=4: 4 chunks = 6.9 waves
// instead of 5 chunks = 8.65 waves).
static void pick_grid(const Dims& g, ZRange zr, int vec, int lpr, int warps_per_block, dim3& grid, dim3& block,
                      int& zchunk, int wx = 1, int resident = 4, int forced_len = 0) {
  const int nz = zr.end - zr.begin;
  block = dim3(32, warps_per_block, 1);
  const int rows_per_block = (warps_per_block / wx) * (32 / lpr);
  const int gx = (g.w + lpr * vec * wx - 1) / (lpr * vec * wx);
  const int gy = (g.h + rows_per_block - 1) / rows_per_block;
  const long long per_plane = (long long)gx * gy;
  static const int min_chunk = env_int("FLOW3D_MIN_ZCHUNK", 8);
  static const int model = env_int("FLOW3D_GRID_MODEL", 0);
  long long len;
  if (forced_len > 0) {
    len = forced_len;
    if (len > nz) len = nz;
    zchunk = (int)len;
    grid = dim3(gx, gy, (nz + zchunk - 1) / zchunk);
    return;
  }
  if (model) {
    const long long slots = (long long)sm_count() * resident;
    const int max_chunks = nz / min_chunk > 1 ? nz / min_chunk : 1;
    double best = -1.0;
    len = nz;
    for (int n = 1; n <= max_chunks && n <= 256; ++n) {
      const long long l = (nz + n - 1) / n;
      const long long chunks = (nz + l - 1) / l;
      const long long total = per_plane * chunks;
      const long long waves = (total + slots - 1) / slots;
      // a wave that is not the only one overlaps its neighbours a little: blend ceil and exact
      const double w_eff = waves > 1 ? 0.75 * (double)waves + 0.25 * (double)total / (double)slots : 1.0;
      const double cost = w_eff * (double)(l + 3);
      if (best < 0.0 || cost < best * 0.995) { best = cost; len = l; }
    }
  } else {
    // enough z chunks for >= ~16 CTAs per SM, but chunks of at least min_chunk planes
    const long long want = (long long)sm_count() * 16;
    long long nchunks = (want + per_plane - 1) / per_plane;
    if (nchunks < 1) nchunks = 1;
    len = (nz + nchunks - 1) / nchunks;
  }
  if (len < min_chunk) len = min_chunk;
  if (len > nz) len = nz;
  if (len < 1) len = 1;
  zchunk = (int)len;
  grid = dim3(gx, gy, (nz + zchunk - 1) / zchunk);
}


static int pick_vec(const Dims& g) {
  static const int forced = env_int("FLOW3D_SWEEP_VEC", 0);  // tuning knob (results are identical)
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  if (g.w >= 224) return 4;  // measured on B200: VEC=2 (more warps) wins below ~200 columns
  if (g.w >= 32) return 2;
  return 1;
}

// ------------------------------------------------------------------------------------------------
// Launch shapes.  Every vector width, z chunking and kernel variant (register-marching warps here, the
// TMA-staged tiles of kernels_sweep_tma.cu) gives identical results, and which is fastest depends on the
// level size in ways a closed-form model misses (wave quantisation at small levels, load balance across
// the two dies at large ones: measured, profiles/).  Shapes therefore come from a per-device table:
//   * launch_sweep / launch_phi_ksi only LOOK UP the table (never time, never synchronise: they are
//     legal under stream capture and truly asynchronous); a missing entry means the static heuristic;
//   * tune_level_kernels() -- reached through flow3d_solver_tune / flow3d_tune_kernels, both documented
//     as synchronous -- times the candidates on scratch buffers and fills the table; its launches are
//     not counted in flow3d_launch_count;
//   * the table is persisted (FLOW3D_TUNE_CACHE or ~/.cache/flow3d_b200/) and reloaded per device.
// FLOW3D_AUTOTUNE=0 ignores the table.
// ------------------------------------------------------------------------------------------------
struct TuneKey {
  int dev, kernel, w, h, ld, nzq;
  bool operator<(const TuneKey& o) const {
    return std::tie(dev, kernel, w, h, ld, nzq) < std::tie(o.dev, o.kernel, o.w, o.h, o.ld, o.nzq);
  }
};
enum { TK_SWEEP = 0, TK_PHI_KSI = 1, TK_SWEEP_KSI = 2, TK_PHI = 3 };
struct TuneEntry {
  TuneCfg cfg;
  bool full;  // false: from a quick pass (implicit first-use tuning); a full pass replaces it; never persisted
};
static std::mutex g_tune_mu;
static std::map<TuneKey, TuneEntry> g_tune;
static unsigned long long g_tune_loaded = 0;  // bit per device: persisted table read

static bool autotune_enabled() {
  static const int on = env_int("FLOW3D_AUTOTUNE", 1);
  return on != 0;
}
static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return dev;
}
static std::string tune_cache_path(int dev) {
  const char* e = getenv("FLOW3D_TUNE_CACHE");
  if (e && *e) return std::string(e) == "off" ? std::string() : std::string(e);
  const char* home = getenv("HOME");
  if (!home || !*home) return std::string();
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return std::string(); }
  std::string name = prop.name;
  for (char& c : name)
    if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9'))) c = '_';
  const std::string dir = std::string(home) + "/.cache/flow3d_b200";
  const std::string cmd = "mkdir -p '" + dir + "' 2>/dev/null";
  if (std::system(cmd.c_str()) != 0) return std::string();
  return dir + "/tune_v3_" + name + "_" + std::to_string(prop.multiProcessorCount) + ".txt";
}
// g_tune_mu held
static void tune_load_locked(int dev) {
  if (dev < 0 || dev >= 64 || (g_tune_loaded & (1ull << dev))) return;
  g_tune_loaded |= 1ull << dev;
  const std::string path = tune_cache_path(dev);
  if (path.empty()) return;
  FILE* f = std::fopen(path.c_str(), "r");
  if (!f) return;
  TuneKey k;
  TuneCfg c;
  k.dev = dev;
  while (std::fscanf(f, "%d %d %d %d %d %d %d %d", &k.kernel, &k.w, &k.h, &k.ld, &k.nzq, &c.vec, &c.nchunks,
                     &c.variant) == 8) {
    if ((c.vec == 1 || c.vec == 2 || c.vec == 4) && c.nchunks >= 0 && c.variant >= 0 && c.variant < SWEEP_VARIANT_COUNT)
      g_tune.emplace(k, TuneEntry{c, true});
  }
  std::fclose(f);
}
static void tune_store(const TuneKey& k, const TuneCfg& c, bool full) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  g_tune[k] = TuneEntry{c, full};
  if (!full) return;
  const std::string path = tune_cache_path(k.dev);
  if (path.empty()) return;
  if (FILE* f = std::fopen(path.c_str(), "a")) {
    std::fprintf(f, "%d %d %d %d %d %d %d %d\n", k.kernel, k.w, k.h, k.ld, k.nzq, c.vec, c.nchunks, c.variant);
    std::fclose(f);
  }
}
// need_full: an entry from a quick pass does not count
static bool tune_lookup(const TuneKey& k, TuneCfg* c, bool need_full = false) {
  if (!autotune_enabled()) return false;
  std::lock_guard<std::mutex> lk(g_tune_mu);
  tune_load_locked(k.dev);
  auto it = g_tune.find(k);
  if (it == g_tune.end() || (need_full && !it->second.full)) return false;
  *c = it->second.cfg;
  return true;
}
// whole-level launches are keyed by their depth; slab launches (ranges that shrink sweep by sweep inside
// one local buffer) by the depth of that buffer
static TuneKey tune_key(int kernel, const Dims& g, ZRange zr) {
  const int nz = zr.end - zr.begin;
  const bool whole = zr.begin == 0 && zr.end == g.d && g.d == g.dg;
  return TuneKey{current_device(), kernel, g.w, g.h, g.ld, whole ? -nz : g.d};
}
static inline int chunk_len(int nz, int nchunks) {
  if (nchunks < 1) nchunks = 1;
  int len = (nz + nchunks - 1) / nchunks;
  return len < 1 ? 1 : len;
}
// candidate chunk counts: chunk lengths around the static heuristic's choice, plus short chunks (2..8
// planes) that shorten the dependent z march of small, latency-bound levels
static void chunk_candidates(int nz, long long per_plane, std::vector<int>& out, bool quick = false) {
  const long long want = (long long)sm_count() * 16;
  long long n0 = (want + per_plane - 1) / per_plane;
  if (n0 < 1) n0 = 1;
  if (n0 > nz) n0 = nz;
  const int len0 = chunk_len(nz, (int)n0);
  int lens[16];
  int nl = 0;
  const double f[] = {4.0, 2.0, 1.333, 1.0, 0.667, 0.5};
  const double fq[] = {2.0, 1.0, 0.5};
  if (quick) for (double k : fq) lens[nl++] = (int)(len0 * k + 0.5);
  else for (double k : f) lens[nl++] = (int)(len0 * k + 0.5);
  const int small_lens[] = {2, 4, 8};
  for (int l : small_lens)
    if (l < len0 && (!quick || l == 4)) lens[nl++] = l;
  out.clear();
  for (int i = 0; i < nl; ++i) {
    int len = lens[i] < 2 ? 2 : lens[i];
    if (len > nz) len = nz;
    const int chunks = (nz + len - 1) / len;
    bool dup = false;
    for (int c : out) dup = dup || c == chunks;
    if (!dup) out.push_back(chunks);
  }
}
// Times every candidate on the caller's (scratch) arguments: out != in, so repeating a launch is
// idempotent.  Two rounds of >= 4 back-to-back launches (>= ~2 ms) per candidate, minimum of the rounds.
// SYNCHRONOUS; the launches are excluded from flow3d_launch_count.
template <class Launch>
static int tune_pick(const std::vector<TuneCfg>& cands, Launch&& launch, cudaStream_t st, TuneCfg* best,
                     const char* what = "", int rounds = 2) {
  if (cands.empty()) return FLOW3D_OK;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess) return FLOW3D_ERR_CUDA;
  if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return FLOW3D_ERR_CUDA; }
  suppress_launch_count(true);
  int rc = FLOW3D_OK;
  auto timed = [&](const TuneCfg& c, int reps, float* ms) {
    if (cudaEventRecord(e0, st) != cudaSuccess) return FLOW3D_ERR_CUDA;
    for (int r = 0; r < reps; ++r) {
      const int k = launch(c);
      if (k != FLOW3D_OK) return k;
    }
    if (cudaEventRecord(e1, st) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess ||
        cudaEventElapsedTime(ms, e0, e1) != cudaSuccess)
      return FLOW3D_ERR_CUDA;
    return FLOW3D_OK;
  };
  float warm_ms = 1.f;
  rc = timed(cands[0], 1, &warm_ms);
  int reps = warm_ms > 0.f ? (int)(2.0f / warm_ms + 0.999f) : 4;
  reps = reps < 4 ? 4 : (reps > 40 ? 40 : reps);
  std::vector<float> t(cands.size(), -1.f);
  for (int round = 0; round < rounds && rc == FLOW3D_OK; ++round) {
    for (size_t i = 0; i < cands.size() && rc == FLOW3D_OK; ++i) {
      float ms = 0.f;
      const int k = timed(cands[i], reps, &ms);
      if (k == FLOW3D_ERR_UNSUPPORTED) continue;  // variant not usable for this level
      rc = k;
      if (rc == FLOW3D_OK && (t[i] < 0.f || ms < t[i])) t[i] = ms;
    }
  }
  float best_ms = -1.f;
  for (size_t i = 0; i < cands.size(); ++i)
    if (t[i] >= 0.f && (best_ms < 0.f || t[i] < best_ms)) { best_ms = t[i]; *best = cands[i]; }
  static const int log = env_int("FLOW3D_TUNE_LOG", 0);
  if (log) {  // per-candidate milliseconds per launch (stderr)
    std::fprintf(stderr, "[flow3d tune] %s reps=%d:", what, reps);
    for (size_t i = 0; i < cands.size(); ++i)
      std::fprintf(stderr, " v%d/x%d/c%d=%.4f%s", cands[i].variant, cands[i].vec, cands[i].nchunks, t[i] / reps,
                   (cands[i].variant == best->variant && cands[i].vec == best->vec && cands[i].nchunks == best->nchunks) ? "*" : "");
    std::fprintf(stderr, "\n");
  }
  suppress_launch_count(false);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

static int launch_sweep_cfg(SweepArgs a, const Dims& g, ZRange zr, TuneCfg cfg, cudaStream_t st) {
  if (sweep_variant_is_tma(cfg.variant))
    return launch_sweep_tma(a, cfg.variant, cfg.nchunks > 0 ? chunk_len(zr.end - zr.begin, cfg.nchunks) : 0, st);
  const int vec = cfg.vec;
  static const int pf = env_int("FLOW3D_SWEEP_PF", 2);
  static const int rows = env_int("FLOW3D_SWEEP_ROWS", 4);
  static const int forced_lpr = env_int("FLOW3D_LPR", 0);
  a.pf = pf;
  a.pf1 = 0;
  a.lpr = (forced_lpr == 8 || forced_lpr == 16 || forced_lpr == 32) ? forced_lpr : pick_lpr(g.w, vec);
  static const int wx_env = env_int("FLOW3D_SWEEP_WX", 1);
  a.wx = (wx_env == 2 || wx_env == 4) && (rows % wx_env == 0) ? wx_env : 1;
  dim3 grid, block;
  pick_grid(g, zr, vec, a.lpr, rows, grid, block, a.zchunk, a.wx, vec == 4 ? 2 : 4,
            cfg.nchunks > 0 ? chunk_len(zr.end - zr.begin, cfg.nchunks) : 0);
  static const int rot = env_int("FLOW3D_SWEEP_ROT", 0), spec = env_int("FLOW3D_SWEEP_SPEC", 0);
  if (a.oksi) {  // first sweep of an outer iteration: computes and stores ksi
    if (vec == 4) launch_chain_kernel(sweep_kernel<4, 2, 0, 0, true>, grid, block, st, a);
    else if (vec == 2) launch_chain_kernel(sweep_kernel<2, 4, 0, 0, true>, grid, block, st, a);
    else launch_chain_kernel(sweep_kernel<1, 4, 0, 0, true>, grid, block, st, a);
    count_launch();
    return check_launch("sweep_kernel<KSI>");
  }
  if (vec == 4) {
    if (rot && spec) launch_chain_kernel(sweep_kernel<4, 2, 1, 1>, grid, block, st, a);
    else if (rot) launch_chain_kernel(sweep_kernel<4, 2, 1, 0>, grid, block, st, a);
    else if (spec) launch_chain_kernel(sweep_kernel<4, 2, 0, 1>, grid, block, st, a);
    else launch_chain_kernel(sweep_kernel<4, 2, 0, 0>, grid, block, st, a);
  } else if (vec == 2) {
    launch_chain_kernel(sweep_kernel<2, 4, 0, 0>, grid, block, st, a);
  } else {
    launch_chain_kernel(sweep_kernel<1, 4, 0, 0>, grid, block, st, a);
  }
  count_launch();
  return check_launch("sweep_kernel");
}

// static choice for a level nobody tuned (FLOW3D_SWEEP_VARIANT forces a kernel variant: tests use it)
static TuneCfg static_sweep_cfg(const Dims& g, ZRange zr) {
  static const int forced_variant = env_int("FLOW3D_SWEEP_VARIANT", -1);
  TuneCfg cfg{pick_vec(g), 0, SWEEP_VARIANT_REG};
  int variant = forced_variant;
  if (variant < 0) variant = SWEEP_VARIANT_REG;
  if (sweep_variant_is_tma(variant) && sweep_tma_usable(g, variant)) {
    cfg.variant = variant;
    const int nz = zr.end - zr.begin;
    cfg.nchunks = (nz + 31) / 32;  // ~32-plane chunks: prologue of 3 planes, enough CTAs to balance the SMs
  }
  return cfg;
}

static void sweep_candidates(const Dims& g, ZRange zr, std::vector<TuneCfg>& cands, bool quick = false) {
  const int nz = zr.end - zr.begin;
  const int v0 = pick_vec(g);
  int vecs[2] = {v0, 0};
  if (v0 == 4) vecs[1] = 2;
  else if (v0 == 2 && g.w >= 128) vecs[1] = 4;
  else if (v0 == 2) vecs[1] = 1;  // narrow levels: more, thinner warps
  if (quick) vecs[1] = 0;
  std::vector<int> chunks;
  for (int vi = 0; vi < 2; ++vi) {
    const int vec = vecs[vi];
    if (!vec) continue;
    const int lpr = pick_lpr(g.w, vec);
    const long long per_plane =
        (long long)((g.w + lpr * vec - 1) / (lpr * vec)) * ((g.h + 4 * (32 / lpr) - 1) / (4 * (32 / lpr)));
    chunk_candidates(nz, per_plane, chunks, quick);
    for (int c : chunks) cands.push_back(TuneCfg{vec, c, SWEEP_VARIANT_REG});
  }
  // The TMA-staged tile kernel (kernels_sweep_tma.cu) is a candidate only on request: measured on B200 it
  // trails the register kernel on every level (512^3: 1.42 ms vs 1.34 ms), so timing it
  // only lengthens the tuning pass.  FLOW3D_TUNE_TMA=1 adds it.
  static const int tune_tma = env_int("FLOW3D_TUNE_TMA", 0);
  for (int variant = 1; tune_tma && !quick && variant < SWEEP_VARIANT_COUNT; ++variant) {
    if (!sweep_variant_is_tma(variant) || !sweep_tma_usable(g, variant) || g.w < 128 || g.h < 32 || nz < 32) continue;
    const int lens[] = {64, 128};
    int seen[8], ns = 0;
    for (int len : lens) {
      if (len > nz) len = nz;
      const int c = (nz + len - 1) / len;
      bool dup = false;
      for (int i = 0; i < ns; ++i) dup = dup || seen[i] == c;
      if (dup) continue;
      seen[ns++] = c;
      cands.push_back(TuneCfg{4, c, variant});
    }
  }
}

int launch_sweep(const float* fx, const float* fy, const float* fz, const float* ft,
                 const float* u, const float* v, const float* w, const float* du, const float* dv,
                 const float* dw, const float* phi, const float* ksi, Dims g, ZRange zr, float hx,
                 float hy, float hz, float alpha, float* odu, float* odv, float* odw, cudaStream_t st,
                 float* ksi_out, float eps_d) {
  if (zr.end <= zr.begin) return FLOW3D_OK;
  SweepArgs a{fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, odu, odv, odw, g, hx, hy, hz, alpha, 0, 0,
              zr.begin, zr.end, 32, 0, 1, ksi_out, eps_d};
  TuneCfg cfg = static_sweep_cfg(g, zr);
  static const int forced_vec = env_int("FLOW3D_SWEEP_VEC", 0), forced_variant = env_int("FLOW3D_SWEEP_VARIANT", -1);
  // a short range inside a deep slab (the "late" launches of the overlapped multi-GPU iteration: a few
  // ghost-dependent planes) must not inherit the slab's chunk count: static shape, one chunk
  const bool short_part = (zr.end - zr.begin) * 2 < g.d;
  if (forced_vec == 0 && forced_variant < 0 && !short_part)
    tune_lookup(tune_key(ksi_out ? TK_SWEEP_KSI : TK_SWEEP, g, zr), &cfg);
  const int rc = launch_sweep_cfg(a, g, zr, cfg, st);
  if (rc == FLOW3D_ERR_UNSUPPORTED && sweep_variant_is_tma(cfg.variant))  // no TMA entry point in this driver
    return launch_sweep_cfg(a, g, zr, TuneCfg{pick_vec(g), 0, SWEEP_VARIANT_REG}, st);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// phi / ksi
// ------------------------------------------------------------------------------------------------
struct PhiKsiArgs {
  const float *fx, *fy, *fz, *ft;
  const float *u, *v, *w;
  const float *du, *dv, *dw;
  float *phi, *ksi;
  Dims g;
  float hx, hy, hz, eps_s, eps_d;
};

// central difference of (f + df) exactly as solve_3d.cu:177-214: ((f[p]-f[m]) + df[p]) - df[m], / (2h)
__device__ __forceinline__ float cdiff(const float* __restrict__ f, const float* __restrict__ df,
                                       long long ip, long long im, float two_h) {
  const float t = __fsub_rn(__fadd_rn(__fsub_rn(__ldg(f + ip), __ldg(f + im)), __ldg(df + ip)), __ldg(df + im));
  return __fdiv_rn(t, two_h);
}

// ((f[p]-f[m]) + df[p]) - df[m] from already-loaded neighbours
__device__ __forceinline__ float cnum(float fp, float fm, float dfp, float dfm) {
  return __fsub_rn(__fadd_rn(__fsub_rn(fp, fm), dfp), dfm);
}

// Same structure as the sweep: one warp per row segment of 32*VEC voxels marching along z, the six
// stencil fields (u,du,v,dv,w,dw) register-rotated in z, x neighbours by shuffle, y neighbours from
// the adjacent rows through L1.
// WITH_KSI = false: phi only (the solver computes ksi in the first sweep of the outer iteration, which
// holds every operand of it anyway; see sweep_plane<.., KSI>)
// EDGE: the warp's tile touches an x face of the volume (mirror selects needed); interior warps run a
// copy of the body without them
template <int VEC, bool WITH_KSI, bool EDGE>
__device__ __forceinline__ void phi_ksi_body(const PhiKsiArgs& a, const LaneMap& lm, int zchunk, int pf, int zs,
                                             int ze) {
  const Dims g = a.g;
  const int y = lm.y;
  const bool active = lm.active;
  const int x0 = lm.x0;
  const int z_begin = zs + blockIdx.z * zchunk;
  const int z_end = min(ze, z_begin + zchunk);
  if (z_begin >= z_end) return;
  const int xh = lm.left_edge ? mirror_idx(x0 - 1, g.w) : (lm.right_edge ? mirror_idx(x0 + VEC, g.w) : x0);
  const unsigned ps = (unsigned)g.ps;
  const unsigned row_c = (unsigned)y * g.ld + x0;
  const unsigned row_m = (unsigned)mirror_idx(y - 1, g.h) * g.ld + x0;
  const unsigned row_p = (unsigned)mirror_idx(y + 1, g.h) * g.ld + x0;
  const unsigned row_h = (unsigned)y * g.ld + xh;
  // divisors 2h are loop invariants: exact division through the FMA sequence of common.cuh
  const ConstDiv thx = make_const_div(__fadd_rn(a.hx, a.hx)), thy = make_const_div(__fadd_rn(a.hy, a.hy)),
                 thz = make_const_div(__fadd_rn(a.hz, a.hz));
  const bool fast_div = thx.fast && thy.fast && thz.fast;

  const float* F[6] = {a.u, a.du, a.v, a.dv, a.w, a.dw};
  Vec<VEC> prev[6], cur[6];
  {
    const unsigned op = (unsigned)z_neighbour(g, z_begin, -1) * ps + row_c;
    const unsigned oc = (unsigned)z_begin * ps + row_c;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      prev[f] = ldv<VEC>(F[f] + op);
      cur[f] = ldv<VEC>(F[f] + oc);
    }
  }
  for (int z = z_begin; z < z_end; ++z) {
    const unsigned pl = (unsigned)z * ps;
    if (pf > 0) {
      const int zp1 = z + 1 + pf;
      if (zp1 < g.d) {
        const unsigned o = (unsigned)zp1 * ps + row_c;
#pragma unroll
        for (int f = 0; f < 6; ++f) prefetch_l2(F[f] + o);
      }
      if constexpr (WITH_KSI) {
        const int zp0 = z + pf;
        if (zp0 < g.d) {
          const unsigned o = (unsigned)zp0 * ps + row_c;
          prefetch_l2(a.fx + o); prefetch_l2(a.fy + o); prefetch_l2(a.fz + o); prefetch_l2(a.ft + o);
        }
      }
    }
    const unsigned on = (unsigned)z_neighbour(g, z, 1) * ps + row_c;
    const unsigned oc = pl + row_c, om = pl + row_m, op = pl + row_p;
    Vec<VEC> next[6], ym[6], yp[6], xm[6], xp[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      next[f] = ldv<VEC>(F[f] + on);
      ym[f] = ldv<VEC>(F[f] + om);
      yp[f] = ldv<VEC>(F[f] + op);
    }
    Vec<VEC> fx, fy, fz, ft;
    if constexpr (WITH_KSI) {
      fx = ldv<VEC>(a.fx + oc); fy = ldv<VEC>(a.fy + oc); fz = ldv<VEC>(a.fz + oc); ft = ldv<VEC>(a.ft + oc);
    }
    float halo[6];
    {
      const unsigned oh = pl + row_h;
#pragma unroll
      for (int f = 0; f < 6; ++f) halo[f] = __ldg(F[f] + oh);
    }
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      if constexpr (EDGE) x_neighbours<VEC>(cur[f], halo[f], lm.left_edge, lm.right_edge, x0, g.w, xm[f], xp[f]);
      else x_neighbours_interior<VEC>(cur[f], halo[f], lm.left_edge, lm.right_edge, xm[f], xp[f]);
    }

    Vec<VEC> ophi, oksi;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      // numerators ((f[p]-f[m]) + df[p]) - df[m] of the nine central differences (solve_3d.cu:177-214)
      float nm[9];
      nm[0] = cnum(xp[0].v[i], xm[0].v[i], xp[1].v[i], xm[1].v[i]);
      nm[1] = cnum(yp[0].v[i], ym[0].v[i], yp[1].v[i], ym[1].v[i]);
      nm[2] = cnum(next[0].v[i], prev[0].v[i], next[1].v[i], prev[1].v[i]);
      nm[3] = cnum(xp[2].v[i], xm[2].v[i], xp[3].v[i], xm[3].v[i]);
      nm[4] = cnum(yp[2].v[i], ym[2].v[i], yp[3].v[i], ym[3].v[i]);
      nm[5] = cnum(next[2].v[i], prev[2].v[i], next[3].v[i], prev[3].v[i]);
      nm[6] = cnum(xp[4].v[i], xm[4].v[i], xp[5].v[i], xm[5].v[i]);
      nm[7] = cnum(yp[4].v[i], ym[4].v[i], yp[5].v[i], ym[5].v[i]);
      nm[8] = cnum(next[4].v[i], prev[4].v[i], next[5].v[i], prev[5].v[i]);
      // Quotients by the loop-invariant 2h through the FMA sequence (common.cuh), with no range test up
      // front: the sequence is exact for |numerator| <= 2^100, and a numerator beyond that (or Inf / NaN)
      // makes the sum of squares below Inf or NaN, which is what triggers the IEEE-division path.  Tiny
      // numerators need no test: a quotient below 2^-80 only enters the sum as an exact zero square.
      // Sum: solve_3d.cu:217-218 as contracted by nvcc, mul(duy,duy) first, then one fma per term.
      auto sum_sq = [&](const float (&q)[9]) {
        float acc = __fmul_rn(q[1], q[1]);
        acc = __fmaf_rn(q[0], q[0], acc);
        acc = __fmaf_rn(q[2], q[2], acc);
        acc = __fmaf_rn(q[3], q[3], acc);
        acc = __fmaf_rn(q[4], q[4], acc);
        acc = __fmaf_rn(q[5], q[5], acc);
        acc = __fmaf_rn(q[6], q[6], acc);
        acc = __fmaf_rn(q[7], q[7], acc);
        acc = __fmaf_rn(q[8], q[8], acc);
        return __fmaf_rn(a.eps_s, a.eps_s, acc);
      };
      float q[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) q[k] = div_const_unchecked(nm[k], (k % 3 == 0) ? thx : (k % 3 == 1) ? thy : thz);
      float acc = sum_sq(q);
      if (!(fast_div && acc <= 3.0e38f)) {
#pragma unroll
        for (int k = 0; k < 9; ++k) q[k] = __fdiv_rn(nm[k], (k % 3 == 0) ? thx.c : (k % 3 == 1) ? thy.c : thz.c);
        acc = sum_sq(q);
      }
      {  // 1 / (2 sqrt(acc)): branch-free fast paths (common.cuh), IEEE intrinsics outside their range
        bool ok = true;
        float sq = sqrt_fast(acc, ok);
        float ph = rcp_fast(__fadd_rn(sq, sq), ok);
        if (!ok) {
          sq = __fsqrt_rn(acc);
          ph = __frcp_rn(__fadd_rn(sq, sq));
        }
        ophi.v[i] = ph;
      }
      if constexpr (WITH_KSI) {
        const float gx = fx.v[i], gy = fy.v[i], gz = fz.v[i], gt = ft.v[i];
        const float J11 = __fmul_rn(gx, gx), J22 = __fmul_rn(gy, gy), J33 = __fmul_rn(gz, gz);
        const float J12 = __fmul_rn(gx, gy), J13 = __fmul_rn(gx, gz), J23 = __fmul_rn(gy, gz);
        const float J14 = __fmul_rn(gx, gt), J24 = __fmul_rn(gy, gt), J34 = __fmul_rn(gz, gt);
        const float du = cur[1].v[i], dv = cur[3].v[i], dw = cur[5].v[i];
        // solve_3d.cu:250-254, operation by operation (row 3 keeps J13*du as the rounded product)
        const float r1 = __fadd_rn(J14, __fmaf_rn(J13, dw, __fmaf_rn(J11, du, __fmul_rn(J12, dv))));
        const float r2 = __fadd_rn(J24, __fmaf_rn(J23, dw, __fmaf_rn(J12, du, __fmul_rn(J22, dv))));
        const float r3 = __fadd_rn(J34, __fmaf_rn(J33, dw, __fmaf_rn(J23, dv, __fmul_rn(J13, du))));
        const float r4 = __fmaf_rn(gt, gt, __fmaf_rn(J34, dw, __fmaf_rn(J14, du, __fmul_rn(J24, dv))));
        float sv = __fadd_rn(__fmaf_rn(dw, r3, __fmaf_rn(du, r1, __fmul_rn(dv, r2))), r4);
        sv = __fmul_rn(sv, (sv > 0.f) ? 1.f : 0.f);
        const float arg = __fmaf_rn(a.eps_d, a.eps_d, sv);
        bool ok2 = true;
        float sq2 = sqrt_fast(arg, ok2);
        float ks = rcp_fast(__fadd_rn(sq2, sq2), ok2);
        if (!ok2) {
          sq2 = __fsqrt_rn(arg);
          ks = __frcp_rn(__fadd_rn(sq2, sq2));
        }
        oksi.v[i] = ks;
      }
    }
    if (active) {
      stv<VEC>(a.phi + oc, ophi);
      if constexpr (WITH_KSI) stv<VEC>(a.ksi + oc, oksi);
    }
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      prev[f] = cur[f];
      cur[f] = next[f];
    }
  }
}

template <int VEC, bool WITH_KSI = true>
__global__ void __launch_bounds__(128) phi_ksi_kernel(const PhiKsiArgs a, int zchunk, int pf, int zs, int ze,
                                                      int lpr) {
  pdl_trigger();
  pdl_wait();
  const LaneMap lm = lane_map<VEC>(lpr, a.g.w, a.g.h);
  if (lm.first_row >= a.g.h) return;
  const int tx0 = lm.tile_x * lpr * VEC, tx1 = tx0 + lpr * VEC;
  const bool edge = (tx0 == 0) || (tx1 > a.g.w - 1);  // warp-uniform
  if (edge) phi_ksi_body<VEC, WITH_KSI, true>(a, lm, zchunk, pf, zs, ze);
  else phi_ksi_body<VEC, WITH_KSI, false>(a, lm, zchunk, pf, zs, ze);
}

static int launch_phi_ksi_cfg(const PhiKsiArgs& a, const Dims& g, ZRange zr, TuneCfg cfg, cudaStream_t st) {
  static const int pf = env_int("FLOW3D_PHIKSI_PF", 2);
  const int vec = cfg.vec;
  dim3 grid, block;
  int zchunk = 0;
  static const int forced_lpr = env_int("FLOW3D_LPR", 0);
  const int lpr = (forced_lpr == 8 || forced_lpr == 16 || forced_lpr == 32) ? forced_lpr : pick_lpr(g.w, vec);
  pick_grid(g, zr, vec, lpr, 4, grid, block, zchunk, 1, 5, cfg.nchunks > 0 ? chunk_len(zr.end - zr.begin, cfg.nchunks) : 0);
  if (!a.ksi) {
    if (vec == 4) launch_chain_kernel(phi_ksi_kernel<4, false>, grid, block, st, a, zchunk, pf, zr.begin, zr.end, lpr);
    else if (vec == 2) launch_chain_kernel(phi_ksi_kernel<2, false>, grid, block, st, a, zchunk, pf, zr.begin, zr.end, lpr);
    else launch_chain_kernel(phi_ksi_kernel<1, false>, grid, block, st, a, zchunk, pf, zr.begin, zr.end, lpr);
  } else if (vec == 4) launch_chain_kernel(phi_ksi_kernel<4>, grid, block, st, a, zchunk, pf, zr.begin, zr.end, lpr);
  else if (vec == 2) launch_chain_kernel(phi_ksi_kernel<2>, grid, block, st, a, zchunk, pf, zr.begin, zr.end, lpr);
  else launch_chain_kernel(phi_ksi_kernel<1>, grid, block, st, a, zchunk, pf, zr.begin, zr.end, lpr);
  count_launch();
  return check_launch("phi_ksi_kernel");
}

// one sweep with an explicit launch shape (tests, scripts): FLOW3D_ERR_UNSUPPORTED if the variant cannot
// run this level
int launch_sweep_shape(const float* fx, const float* fy, const float* fz, const float* ft, const float* u,
                       const float* v, const float* w, const float* du, const float* dv, const float* dw,
                       const float* phi, const float* ksi, Dims g, ZRange zr, float hx, float hy, float hz,
                       float alpha, float* odu, float* odv, float* odw, cudaStream_t st, float* ksi_out, float eps_d,
                       int variant, int vec, int nchunks) {
  if (zr.end <= zr.begin) return FLOW3D_OK;
  if (variant < 0 || variant >= SWEEP_VARIANT_COUNT || !(vec == 1 || vec == 2 || vec == 4) || nchunks < 0)
    return FLOW3D_ERR_INVALID_ARG;
  SweepArgs a{fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, odu, odv, odw, g, hx, hy, hz, alpha, 0, 0,
              zr.begin, zr.end, 32, 0, 1, ksi_out, eps_d};
  return launch_sweep_cfg(a, g, zr, TuneCfg{vec, nchunks, variant}, st);
}

static int static_phi_vec(const Dims& g) {
  static const int forced = env_int("FLOW3D_PHIKSI_VEC", 0);
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  return (g.w >= 48) ? 2 : 1;  // measured: VEC=2 (16+ warps/SM) beats VEC=4 (191 regs)
}

int launch_phi_ksi(const float* fx, const float* fy, const float* fz, const float* ft,
                   const float* u, const float* v, const float* w, const float* du,
                   const float* dv, const float* dw, Dims g, ZRange zr, float hx, float hy, float hz,
                   float eps_s, float eps_d, float* phi, float* ksi, cudaStream_t st) {
  if (zr.end <= zr.begin) return FLOW3D_OK;
  PhiKsiArgs a{fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, g, hx, hy, hz, eps_s, eps_d};
  static const int forced = env_int("FLOW3D_PHIKSI_VEC", 0);
  TuneCfg cfg{static_phi_vec(g), 0, 0};
  if (forced == 0 && (zr.end - zr.begin) * 2 >= g.d) tune_lookup(tune_key(ksi ? TK_PHI_KSI : TK_PHI, g, zr), &cfg);
  return launch_phi_ksi_cfg(a, g, zr, cfg, st);
}

// ------------------------------------------------------------------------------------------------
// explicit tuning of one level's solver kernels (SYNCHRONOUS).  `bufs`: 16 scratch volumes of the level's
// size; they are filled with a finite positive pattern (0x3f3f3f3f = 0.747) so that no candidate runs
// through denormal / NaN slow paths.
// ------------------------------------------------------------------------------------------------
// quick: the implicit first-use pass of flow3d_solver_compute_host -- the static vector width, four chunk
// lengths, one timing round (~5x cheaper); its entries stay in memory and are replaced by a full pass.
int tune_level_kernels(Dims g, ZRange zr, float* const bufs[16], float hx, float hy, float hz, cudaStream_t st,
                       bool quick) {
  if (!autotune_enabled()) return FLOW3D_OK;
  const int nz = zr.end - zr.begin;
  if (nz < 4) return FLOW3D_OK;
  const TuneKey k_sweep = tune_key(TK_SWEEP, g, zr), k_ksi = tune_key(TK_SWEEP_KSI, g, zr),
                k_phi = tune_key(TK_PHI, g, zr);
  TuneCfg tmp;
  const bool have_sweep = tune_lookup(k_sweep, &tmp, !quick), have_ksi = tune_lookup(k_ksi, &tmp, !quick),
             have_phi = tune_lookup(k_phi, &tmp, !quick);
  const int rounds = quick ? 1 : 2;
  if (have_sweep && have_ksi && have_phi) return FLOW3D_OK;
  const size_t bytes = (size_t)g.ps * g.d * sizeof(float);
  for (int i = 0; i < 16; ++i)
    if (cudaMemsetAsync(bufs[i], 0x3f, bytes, st) != cudaSuccess) { cudaGetLastError(); return FLOW3D_ERR_CUDA; }
  const float alpha = 7.5f, eps = 0.001f;
  SweepArgs a{bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5], bufs[6], bufs[7], bufs[8], bufs[9], bufs[10],
              bufs[11], bufs[12], bufs[13], bufs[14], g, hx, hy, hz, alpha, 0, 0, zr.begin, zr.end, 32, 0, 1,
              nullptr, eps};
  std::vector<TuneCfg> cands;
  TuneCfg best;
  char lbl[96];
  auto label = [&](const char* k) {
    std::snprintf(lbl, sizeof(lbl), "%s %dx%dx%d[%d,%d)", k, g.w, g.h, g.d, zr.begin, zr.end);
    return (const char*)lbl;
  };
  if (!have_sweep) {
    sweep_candidates(g, zr, cands, quick);
    best = static_sweep_cfg(g, zr);
    F3D_TRY_RC(tune_pick(cands, [&](TuneCfg c) { return launch_sweep_cfg(a, g, zr, c, st); }, st, &best, label("sweep"), rounds));
    tune_store(k_sweep, best, !quick);
  }
  if (!have_ksi) {
    SweepArgs ak = a;
    ak.oksi = bufs[15];
    cands.clear();
    sweep_candidates(g, zr, cands, quick);
    best = static_sweep_cfg(g, zr);
    F3D_TRY_RC(tune_pick(cands, [&](TuneCfg c) { return launch_sweep_cfg(ak, g, zr, c, st); }, st, &best, label("sweep+ksi"), rounds));
    tune_store(k_ksi, best, !quick);
  }
  if (!have_phi) {
    PhiKsiArgs pa{bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], bufs[5], bufs[6], bufs[7], bufs[8], bufs[9], bufs[10],
                  nullptr, g, hx, hy, hz, eps, eps};
    cands.clear();
    const int v0 = static_phi_vec(g);
    int vecs[2] = {v0, (v0 == 2 && g.w >= 128 && !quick) ? 4 : 0};
    std::vector<int> chunks;
    for (int vi = 0; vi < 2; ++vi) {
      const int vc = vecs[vi];
      if (!vc) continue;
      const int lpr = pick_lpr(g.w, vc);
      const long long per_plane =
          (long long)((g.w + lpr * vc - 1) / (lpr * vc)) * ((g.h + 4 * (32 / lpr) - 1) / (4 * (32 / lpr)));
      chunk_candidates(nz, per_plane, chunks, quick);
      for (int c : chunks) cands.push_back(TuneCfg{vc, c, 0});
    }
    best = TuneCfg{v0, 0, 0};
    F3D_TRY_RC(tune_pick(cands, [&](TuneCfg c) { return launch_phi_ksi_cfg(pa, g, zr, c, st); }, st, &best, label("phi"), rounds));
    tune_store(k_phi, best, !quick);
  }
  return FLOW3D_OK;
}

// read-only view of the table (tests, scripts/level_table.py)
int tune_query(int kernel, Dims g, ZRange zr, int out[3]) {
  TuneCfg c;
  if (!tune_lookup(tune_key(kernel, g, zr), &c)) return 0;
  out[0] = c.vec; out[1] = c.nchunks; out[2] = c.variant;
  return 1;
}

// ------------------------------------------------------------------------------------------------
// u += du (three components in one launch; add_3d.cu:37-40 is a plain rounded add)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add3_kernel(float* __restrict__ u, float* __restrict__ v,
                                                   float* __restrict__ w, const float* __restrict__ du,
                                                   const float* __restrict__ dv,
                                                   const float* __restrict__ dw, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4* U = reinterpret_cast<float4*>(u) + i;
  float4* V = reinterpret_cast<float4*>(v) + i;
  float4* W = reinterpret_cast<float4*>(w) + i;
  const float4 a = *U, b = __ldg(reinterpret_cast<const float4*>(du) + i);
  *U = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
  const float4 c = *V, d = __ldg(reinterpret_cast<const float4*>(dv) + i);
  *V = make_float4(__fadd_rn(c.x, d.x), __fadd_rn(c.y, d.y), __fadd_rn(c.z, d.z), __fadd_rn(c.w, d.w));
  const float4 e = *W, f = __ldg(reinterpret_cast<const float4*>(dw) + i);
  *W = make_float4(__fadd_rn(e.x, f.x), __fadd_rn(e.y, f.y), __fadd_rn(e.z, f.z), __fadd_rn(e.w, f.w));
}

int launch_add3(float* u, float* v, float* w, const float* du, const float* dv, const float* dw,
                Dims g, cudaStream_t st) {
  // padding columns are added too (harmless: never interpreted)
  const long long n4 = g.ps * g.d / 4;
  const int block = 256;
  const long long grid = (n4 + block - 1) / block;
  add3_kernel<<<(unsigned)grid, block, 0, st>>>(u, v, w, du, dv, dw, n4);
  count_launch();
  return check_launch("add3_kernel");
}

// ------------------------------------------------------------------------------------------------
// max |x| (used to size the frame-1 halo of a z-sharded warp): warp shuffles + one atomicMax on the
// non-negative float's integer image
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ in, Dims g,
                                                     unsigned* __restrict__ out) {
  float m = 0.f;
  const long long rows = (long long)g.h * g.d;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = in + r * g.ld;
    for (int x = threadIdx.x; x < g.w; x += blockDim.x) {
      const float v = fabsf(__ldg(row + x));
      m = (v > m || v != v) ? v : m;  // NaN sticks (largest bit pattern in the atomicMax below)
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, m, o);
    m = (t > m || t != t) ? t : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

int launch_absmax(const float* in, Dims g, float* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
  if (e != cudaSuccess) { note_cuda_error(e, "cudaMemsetAsync"); return FLOW3D_ERR_CUDA; }
  long long rows = (long long)g.h * g.d;
  long long blocks = rows < (long long)sm_count() * 16 ? rows : (long long)sm_count() * 16;
  absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, g, reinterpret_cast<unsigned*>(out));
  count_launch();
  return check_launch("absmax_kernel");
}

}  // namespace f3d
