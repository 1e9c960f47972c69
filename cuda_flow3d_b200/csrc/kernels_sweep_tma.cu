// kernels_sweep_tma.cu -- the Jacobi sweep (reference: src/kernels/solve_3d.cu:264-508) as a TMA-staged,
// shared-memory-tiled 2.5D z-marching stencil for sm_100a.
//
// A CTA owns a BX x BY column of voxels and marches it through a chunk of z planes.  A dedicated producer
// warp (one elected lane) streams every plane into shared memory three planes ahead with
// cp.async.bulk.tensor (TMA, 3D tensor maps, one box per field and plane, out-of-volume cells zero-filled
// by the hardware) and signals arrival through mbarriers, so neither DRAM latency nor the descriptor
// set-up of the twelve copies per plane is ever exposed to the four arithmetic warps:
//
//   ring A (4 slots)  u,v,w and du,dv,dw, box (BX+8) x (BY+2): the tile plus its one-voxel halo (the box
//                     starts 4 columns left of the tile so every thread's own float4 stays 16 B-aligned)
//   ring B (3 slots)  phi, same box
//   ring C (2 slots)  fx,fy,fz,ft,ksi, box BX x BY (used at the centre only)
//
// Per plane step q:  main(q)  -- every thread updates its own 4 voxels of plane q from shared memory;
//                    pre(q+2) -- the CTA turns the landed du,dv,dw of plane q+2 into S = u+du, v+dv, w+dw
//                                IN PLACE (one rounded add per voxel, shared by its six consumers, exactly
//                                the value the reference forms per neighbour: solve_3d.cu:470-490) and
//                                copies dv,dw (the centre-only raw increments) to a compact ring D;
//                    one CTA-wide named barrier (producer warp included).
// The z-1 and centre values of S and phi of a thread's own column are carried in registers, which is what
// lets four A slots cover a prefetch distance of three planes.  Rings B and C are refilled one step before
// their first use (shared memory is spent on ring A: 112 KB per CTA, two CTAs per SM), so the arithmetic
// threads prefetch their own 16 B of those fields into L2 three planes ahead.
//
// Status: bit-identical to the register-marching kernel of kernels_solve.cu on every test,
// 1.42 ms against its 1.34 ms per 512^3 sweep on a B200 -- both end up bound by the ~185 instructions per
// voxel-sweep at two warps per scheduler, not by memory.  It is therefore not in the default tuning
// candidates (FLOW3D_TUNE_TMA=1 adds it); flow3d_sweep_shape / FLOW3D_SWEEP_VARIANT select it explicitly.
//
// Arithmetic: the same explicit round-to-nearest operation sequence as kernels_solve.cu (bit-identical to
// the reference's compiled kernels); mirror (reflect-101) neighbours are substituted exactly where the
// reference's shared-memory halo holds them, so even the signs of zeros agree.
#include <cuda.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

#include "solve_args.cuh"

namespace f3d {

// ------------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier + TMA)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// CTA-wide barrier reached from the producer and the consumer branch alike (a named barrier: legal in
// divergent code, unlike __syncthreads)
template <int THREADS>
__device__ __forceinline__ void cta_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tile geometry and shared-memory layout
// ------------------------------------------------------------------------------------------------
enum { M_U = 0, M_V, M_W, M_DU, M_DV, M_DW, M_PHI, M_FX, M_FY, M_FZ, M_FT, M_KSI, M_COUNT };
struct TmaMaps {
  CUtensorMap m[M_COUNT];
};

template <int BX, int BY, bool KSI>
struct TL {
  static constexpr int LX = BX / 4;         // threads per tile row (4 voxels each)
  static constexpr int CONSUMERS = LX * BY;     // arithmetic threads
  static constexpr int THREADS = CONSUMERS + 32;  // + the TMA producer warp
  static constexpr int HX = BX + 8;         // haloed box: columns x0-4 .. x0+BX+3
  static constexpr int HY = BY + 2;         // rows y0-1 .. y0+BY
  static constexpr int HX4 = HX / 4;
  static constexpr int HBOX = HX * HY * 4;  // bytes one haloed box delivers
  static constexpr int HP = ((HBOX + 127) / 128) * 128;  // slot pitch (TMA destinations are 128 B-aligned)
  static constexpr int CP = BX * BY * 4;
  static constexpr int NC = KSI ? 4 : 5;    // centre-only TMA fields: fx,fy,fz,ft(,ksi)
  static constexpr int ND = KSI ? 3 : 2;    // raw increments kept for the centre: (du,) dv, dw
  static constexpr int NSA = 4, NSB = 3, NSC = 2, NSD = 3;
  static constexpr int OFF_S = 0;                          // [NSA][3] du,dv,dw -> S in place
  static constexpr int OFF_U = OFF_S + NSA * 3 * HP;       // [NSA][3] u,v,w
  static constexpr int OFF_P = OFF_U + NSA * 3 * HP;       // [NSB]    phi
  static constexpr int OFF_C = OFF_P + NSB * HP;           // [NSC][NC]
  static constexpr int OFF_D = OFF_C + NSC * NC * CP;      // [NSD][ND]
  static constexpr int OFF_BAR = OFF_D + NSD * ND * CP;    // NSA + NSB + NSC mbarriers
  static constexpr int BYTES = OFF_BAR + 16 * 8 + 128;     // + slack to align the base to 128 B
  static_assert(CP % 128 == 0, "compact planes must keep 128 B alignment");
  static_assert(CONSUMERS % 32 == 0 && 32 % LX == 0, "a warp must cover whole tile rows");
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// x-1 / x+1 neighbours of a thread's four voxels: inside the float4, from the adjacent lanes of the same
// tile row (shuffle) and, for the first / last lane of a row, from the halo columns in shared memory.
// At the volume faces the reflect-101 neighbour is substituted (x = 0 -> value at 1, x = w-1 -> value at
// w-2), what the reference's shared-memory halo holds (solve_3d.cu:326-355).  One code path serves every
// tile (a two-variant build stalled on instruction fetch), and the faces cost almost nothing in it:
//   * x = 0 can only be element 0 of a row's first lane: it takes c.y instead of the halo (face_lo);
//   * x = w-1 as element 3 takes c.z instead of its right neighbour (face_hi3);
//   * x = w-1 as element i < 3 (w not a multiple of 4) is handled where the vector is LOADED, by
//     mirror_tail(): element i+1 (column w, out of the volume, TMA zero fill) is overwritten with element
//     i-1, so the plain in-vector neighbour is already the mirrored one.
template <int LX>
__device__ __forceinline__ void x_nb(const float4& c, const float* row_c, int lx, bool face_lo, bool face_hi3,
                                     float (&l)[4], float (&r)[4]) {
  float from_left = __shfl_up_sync(0xffffffffu, c.w, 1);
  float from_right = __shfl_down_sync(0xffffffffu, c.x, 1);
  if (lx == 0) from_left = row_c[-1];
  if (lx == LX - 1) from_right = row_c[4];
  if (face_lo) from_left = c.y;
  if (face_hi3) from_right = c.z;
  l[0] = from_left; l[1] = c.x; l[2] = c.y; l[3] = c.z;
  r[0] = c.y; r[1] = c.z; r[2] = c.w; r[3] = from_right;
}
// tail = index (1..3) of the first out-of-volume element of a vector that holds x = w-1 as element
// tail-1 < 3, else 0: that element receives the mirror neighbour of x = w-1 (element tail-2; for tail = 1
// that is the previous lane's last element, which the caller passes as `left`)
__device__ __forceinline__ void mirror_tail(float4& c, int tail, float left) {
  if (tail == 1) c.y = left;
  else if (tail == 2) c.z = c.x;
  else if (tail == 3) c.w = c.y;
}

// IEEE-division recomputation of one voxel whose operands left div_fast's range (kept out of line: rare,
// and the hot loop must stay small)
__device__ __noinline__ float3 slow_voxel(float numU, float denU, float denV, float denW, float k, float sV, float sW,
                                          float j12, float j13, float j23, float j24, float j34, float dw) {
  const float du = __fdiv_rn(numU, denU);
  const float ndv = __fmaf_rn(-j23, dw, __fmaf_rn(-j12, du, -j24));
  const float dv = __fdiv_rn(__fmaf_rn(k, ndv, sV), denV);
  const float ndw = __fmaf_rn(-j23, dv, __fmaf_rn(-j13, du, -j34));
  return make_float3(du, dv, __fdiv_rn(__fmaf_rn(k, ndw, sW), denW));
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
constexpr int kPf = 3;  // planes between a thread's L2 prefetch and the TMA copy of that plane

template <int BX, int BY, bool KSI>
__device__ __forceinline__ void sweep_tma_run(const TmaMaps& maps, const SweepArgs& a, unsigned char* smem) {
  using L = TL<BX, BY, KSI>;
  const Dims g = a.g;
  const int tid = threadIdx.x;
  const int lx = tid % L::LX, ly = tid / L::LX;
  const int tx0 = blockIdx.x * BX, ty0 = blockIdx.y * BY;
  const int gx0 = tx0 + 4 * lx, gy = ty0 + ly;
  const bool active = gx0 < g.w && gy < g.h;
  const int zb = a.zs + blockIdx.z * a.zchunk;
  const int n = min(a.ze, zb + a.zchunk) - zb;  // planes this CTA updates: sequence index q = 0 .. n-1
  if (n <= 0) return;

  const uint32_t sbase = smem_u32(smem);
  const uint32_t barA = sbase + L::OFF_BAR, barB = barA + 8 * L::NSA, barC = barB + 8 * L::NSB;
  // local plane of sequence index q (q = -1 and q = n are the z neighbours of the chunk: reflected at the
  // GLOBAL faces, ghost planes of the slab elsewhere)
  auto zplane = [&](int q) { return z_neighbour(g, zb + q, 0); };

  auto issueA = [&](int q) {  // u,v,w,du,dv,dw of plane q, haloed
    const int k = q + 1, s = k & 3;
    const uint32_t bar = barA + 8 * s;
    mbar_expect_tx(bar, 6 * L::HBOX);
    const int z = zplane(q);
#pragma unroll 1
    for (int f = 0; f < 3; ++f) {
      tma_load_3d(sbase + L::OFF_U + (s * 3 + f) * L::HP, &maps.m[M_U + f], tx0 - 4, ty0 - 1, z, bar);
      tma_load_3d(sbase + L::OFF_S + (s * 3 + f) * L::HP, &maps.m[M_DU + f], tx0 - 4, ty0 - 1, z, bar);
    }
  };
  auto issueB = [&](int q) {  // phi of plane q, haloed
    const int k = q + 1, s = k % 3;
    const uint32_t bar = barB + 8 * s;
    mbar_expect_tx(bar, L::HBOX);
    tma_load_3d(sbase + L::OFF_P + s * L::HP, &maps.m[M_PHI], tx0 - 4, ty0 - 1, zplane(q), bar);
  };
  auto issueC = [&](int q) {  // centre-only fields of plane q
    const int s = q & 1;
    const uint32_t bar = barC + 8 * s;
    mbar_expect_tx(bar, L::NC * L::CP);
    const int z = zb + q;
#pragma unroll 1
    for (int f = 0; f < L::NC; ++f)
      tma_load_3d(sbase + L::OFF_C + (s * L::NC + f) * L::CP, &maps.m[M_FX + f], tx0, ty0, z, bar);
  };
  auto waitA = [&](int q) { const int k = q + 1; mbar_wait(barA + 8 * (k & 3), (k >> 2) & 1); };
  auto waitB = [&](int q) { const int k = q + 1; mbar_wait(barB + 8 * (k % 3), (k / 3) & 1); };
  auto waitC = [&](int q) { mbar_wait(barC + 8 * (q & 1), (q >> 1) & 1); };

  float* const fS = reinterpret_cast<float*>(smem + L::OFF_S);
  float* const fU = reinterpret_cast<float*>(smem + L::OFF_U);
  float* const fP = reinterpret_cast<float*>(smem + L::OFF_P);
  float* const fC = reinterpret_cast<float*>(smem + L::OFF_C);
  float* const fD = reinterpret_cast<float*>(smem + L::OFF_D);
  constexpr int HPf = L::HP / 4, CPf = L::CP / 4;

  // S = u + du etc. of plane q in place, over the whole haloed box; raw increments of the tile to ring D.
  // Thread t handles the float4 cells t, t + CONSUMERS, ...: their offsets are loop invariants.
  constexpr int NCELL = L::HX4 * L::HY;
  constexpr int PRE_ITERS = (NCELL + L::CONSUMERS - 1) / L::CONSUMERS;
  int pre_off[PRE_ITERS], pre_coff[PRE_ITERS];  // pre_coff < 0: not a tile cell (halo) or no cell
#pragma unroll
  for (int it = 0; it < PRE_ITERS; ++it) {
    const int i = tid + it * L::CONSUMERS;
    const int r = i / L::HX4, c = i - r * L::HX4;
    pre_off[it] = (i < NCELL) ? r * L::HX + 4 * c : -1;
    const bool inner = (i < NCELL) && (r >= 1) && (r <= BY) && (c >= 1) && (c <= L::LX);
    pre_coff[it] = inner ? (r - 1) * BX + 4 * (c - 1) : -1;
  }
  auto prepass = [&](int q) {
    const int k = q + 1, sa = k & 3, sd = k % 3;
    float* S0 = fS + (sa * 3) * HPf;
    const float* U0 = fU + (sa * 3) * HPf;
    float* D0 = fD + (sd * L::ND) * CPf;
#pragma unroll
    for (int it = 0; it < PRE_ITERS; ++it) {
      const int off = pre_off[it], coff = pre_coff[it];
      if (off < 0) continue;
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const float4 d = ld4(S0 + f * HPf + off);
        const float4 uu = ld4(U0 + f * HPf + off);
        st4(S0 + f * HPf + off, add4(uu, d));
        if (coff >= 0) {
          if constexpr (KSI) st4(D0 + f * CPf + coff, d);
          else if (f > 0) st4(D0 + (f - 1) * CPf + coff, d);
        }
      }
    }
  };

  // ---- prologue ------------------------------------------------------------------------------------
  const bool producer = tid >= L::CONSUMERS;
  if (tid == L::CONSUMERS) {
    for (int i = 0; i < L::NSA + L::NSB + L::NSC; ++i) mbar_init(barA + 8 * i, 1);
    fence_barrier_init();
  }
  cta_sync<L::THREADS>();
  if (producer) {
    // The producer warp: lane 0 issues the copies of step q right after the barrier that ended step q-1
    // (every slot it refills was last read before that barrier); the warp then waits at the next barrier.
    const bool lead = tid == L::CONSUMERS;
    if (lead) {
#pragma unroll 1
      for (int i = 0; i < M_COUNT; ++i)
        if (!(KSI && i == M_KSI)) tma_prefetch_desc(&maps.m[i]);
#pragma unroll 1
      for (int q = -1; q <= 2; ++q) {  // planes -1 .. 2 of ring A, -1 .. 1 of ring B, plane 0 of ring C
        if (q <= n) issueA(q);
        if (q <= 1) issueB(q);
        if (q == 0) issueC(0);
      }
    }
    cta_sync<L::THREADS>();  // consumers: S of planes -1, 0, 1 formed
    cta_sync<L::THREADS>();  // consumers: register-carried planes loaded (slot 0 may be refilled)
#pragma unroll 1
    for (int q = 0; q < n; ++q) {
      if (lead) {
        if (q + 3 <= n) issueA(q + 3);
        if (q + 2 <= n) issueB(q + 2);
        if (q + 1 <= n - 1) issueC(q + 1);
      }
      cta_sync<L::THREADS>();
    }
    return;
  }
  // per-thread constants
  const int hx0 = 4 + 4 * lx, hy = 1 + ly;
  const int rc = hy * L::HX + hx0;
  const int rm = ((gy == 0) ? hy + 1 : hy - 1) * L::HX + hx0;       // reflect-101 in y
  const int rp = ((gy == g.h - 1) ? hy - 1 : hy + 1) * L::HX + hx0;
  const int cc = ly * BX + 4 * lx;
  const float hx2 = __fdiv_rn(a.alpha, __fmul_rn(a.hx, a.hx));  // solve_3d.cu:451-460
  const float hy2 = __fdiv_rn(a.alpha, __fmul_rn(a.hy, a.hy));
  const float hz2 = __fdiv_rn(a.alpha, __fmul_rn(a.hz, a.hz));
  const float wyp = (gy < g.h - 1) ? hy2 : 0.f;
  const float wym = (gy > 0) ? hy2 : 0.f;
  const unsigned row_g = (unsigned)gy * g.ld + gx0;
  // x faces: loop-invariant per thread (x does not change along the march)
  const bool face_lo = gx0 == 0;
  const bool face_hi3 = gx0 + 3 == g.w - 1;
  const int tail = (g.w - 1 >= gx0 && g.w - 1 < gx0 + 3) ? (g.w - gx0) : 0;  // see mirror_tail
  float wxp4[4], wxm4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = gx0 + i;
    wxp4[i] = (x < g.w - 1) ? hx2 : 0.f;
    wxm4[i] = (x > 0) ? hx2 : 0.f;
  }
  // own column of a haloed plane buffer, with the in-vector face mirrored
  auto load_own = [&](const float* plane) {
    float4 v = ld4(plane + rc);
    if (tail) mirror_tail(v, tail, plane[rc - 1]);
    return v;
  };

#pragma unroll 1
  for (int q = -1; q <= 1; ++q) { waitA(q); prepass(q); }
  fence_proxy_async();
  cta_sync<L::THREADS>();
  waitB(-1);
  waitB(0);
  // own column, planes q-1 and q, carried in registers
  float4 Pu = load_own(fS + (0 * 3 + 0) * HPf), Pv = load_own(fS + (0 * 3 + 1) * HPf),
         Pw = load_own(fS + (0 * 3 + 2) * HPf), Pp = load_own(fP + 0 * HPf);
  float4 Cu = load_own(fS + (1 * 3 + 0) * HPf), Cv = load_own(fS + (1 * 3 + 1) * HPf),
         Cw = load_own(fS + (1 * 3 + 2) * HPf), Cp = load_own(fP + 1 * HPf);
  cta_sync<L::THREADS>();  // slot 0 (plane -1) is refilled by the first step's TMA: everyone must have read it

  int sb_c = 1, sb_n = 2;  // ring B slots of planes q and q+1
  int sd_c = 1;            // ring D slot of plane q
#pragma unroll 1
  for (int q = 0; q < n; ++q) {
    // Rings B and C are refilled only one step before their first use (shared memory goes to the deep
    // ring A), so their copies must be L2 hits: every thread prefetches its own 16 B of phi and of the
    // centre-only fields a few planes ahead (plain prefetch.global.L2; a TMA tensor prefetch of the same
    // boxes measured slower and raised DRAM traffic by 28 %).
    if (active) {
      const int zp = zb + q + kPf;
      if (zp < a.ze) {
        const unsigned o = (unsigned)zp * (unsigned)g.ps + row_g;
        prefetch_l2(a.fx + o); prefetch_l2(a.fy + o); prefetch_l2(a.fz + o); prefetch_l2(a.ft + o);
        if constexpr (!KSI) prefetch_l2(a.ksi + o);
        if (zp + 1 < g.d) prefetch_l2(a.phi + o + (unsigned)g.ps);
      }
    }
    waitB(q + 1);
    waitC(q);
    const int sa_c = (q + 1) & 3, sa_n = (q + 2) & 3, sc = q & 1;
    const float* Sc = fS + (sa_c * 3) * HPf;
    const float* Sn = fS + (sa_n * 3) * HPf;
    const float* Uc = fU + (sa_c * 3) * HPf;
    const float* Pc = fP + sb_c * HPf;
    const float* Pn = fP + sb_n * HPf;
    const float* Cc = fC + (sc * L::NC) * CPf;
    const float* Dc = fD + (sd_c * L::ND) * CPf;

    // plane q+1 of the own column (becomes the centre of the next step)
    const float4 Nu = load_own(Sn + 0 * HPf), Nv = load_own(Sn + 1 * HPf), Nw = load_own(Sn + 2 * HPf),
                 Np = load_own(Pn);
    // x neighbours (shuffles: every lane takes part)
    float Su_xm[4], Su_xp[4], Sv_xm[4], Sv_xp[4], Sw_xm[4], Sw_xp[4], ph_xm[4], ph_xp[4];
    x_nb<L::LX>(Cu, Sc + 0 * HPf + rc, lx, face_lo, face_hi3, Su_xm, Su_xp);
    x_nb<L::LX>(Cv, Sc + 1 * HPf + rc, lx, face_lo, face_hi3, Sv_xm, Sv_xp);
    x_nb<L::LX>(Cw, Sc + 2 * HPf + rc, lx, face_lo, face_hi3, Sw_xm, Sw_xp);
    x_nb<L::LX>(Cp, Pc + rc, lx, face_lo, face_hi3, ph_xm, ph_xp);

    if (active) {
      const float4 Su_ym = ld4(Sc + 0 * HPf + rm), Su_yp = ld4(Sc + 0 * HPf + rp);
      const float4 Sv_ym = ld4(Sc + 1 * HPf + rm), Sv_yp = ld4(Sc + 1 * HPf + rp);
      const float4 Sw_ym = ld4(Sc + 2 * HPf + rm), Sw_yp = ld4(Sc + 2 * HPf + rp);
      const float4 ph_ym = ld4(Pc + rm), ph_yp = ld4(Pc + rp);
      const float4 u4 = ld4(Uc + 0 * HPf + rc), v4 = ld4(Uc + 1 * HPf + rc), w4 = ld4(Uc + 2 * HPf + rc);
      const float4 fx4 = ld4(Cc + 0 * CPf + cc), fy4 = ld4(Cc + 1 * CPf + cc), fz4 = ld4(Cc + 2 * CPf + cc),
                   ft4 = ld4(Cc + 3 * CPf + cc);
      float4 ks4 = make_float4(0.f, 0.f, 0.f, 0.f), du4 = ks4;
      if constexpr (!KSI) ks4 = ld4(Cc + 4 * CPf + cc);
      if constexpr (KSI) du4 = ld4(Dc + 0 * CPf + cc);
      const float4 dv4 = ld4(Dc + (L::ND - 2) * CPf + cc), dw4 = ld4(Dc + (L::ND - 1) * CPf + cc);

      const int zg = g.z0g + zb + q;  // faces are the GLOBAL ones when the level is sharded
      const float wzp = (zg < g.dg - 1) ? hz2 : 0.f;
      const float wzm = (zg > 0) ? hz2 : 0.f;

      // phase 1: everything up to the three dependent divisions, for the four voxels
      float numU[4], denU[4], denV[4], denW[4], rks[4], sV[4], sW[4], j12[4], j13[4], j23[4], j24[4], j34[4], dwv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float wxp = wxp4[i], wxm = wxm4[i];
        const float gx = comp(fx4, i), gyv = comp(fy4, i), gz = comp(fz4, i), gt = comp(ft4, i);
        const float J11 = __fmul_rn(gx, gx);
        const float J22 = __fmul_rn(gyv, gyv);
        const float J33 = __fmul_rn(gz, gz);
        const float J12 = __fmul_rn(gx, gyv);
        const float J13 = __fmul_rn(gx, gz);
        const float J23 = __fmul_rn(gyv, gz);
        const float J14 = __fmul_rn(gx, gt);
        const float J24 = __fmul_rn(gyv, gt);
        const float J34 = __fmul_rn(gz, gt);
        const float pc = comp(Cp, i);
        // face weights: w * (phi_n + phi_c)/2  (solve_3d.cu:462-469); plain products
        const float axp = __fmul_rn(wxp, __fmul_rn(__fadd_rn(ph_xp[i], pc), 0.5f));
        const float axm = __fmul_rn(wxm, __fmul_rn(__fadd_rn(ph_xm[i], pc), 0.5f));
        const float ayp = __fmul_rn(wyp, __fmul_rn(__fadd_rn(comp(ph_yp, i), pc), 0.5f));
        const float aym = __fmul_rn(wym, __fmul_rn(__fadd_rn(comp(ph_ym, i), pc), 0.5f));
        const float azp = __fmul_rn(wzp, __fmul_rn(__fadd_rn(comp(Np, i), pc), 0.5f));
        const float azm = __fmul_rn(wzm, __fmul_rn(__fadd_rn(comp(Pp, i), pc), 0.5f));
        const float sumH = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(axp, axm), ayp), aym), azp), azm);
        const float uc = comp(u4, i), vc = comp(v4, i), wc = comp(w4, i);
        // solve_3d.cu:470-490: seed with the rounded x- product, then one fma per face
        float sumU = __fmul_rn(axm, __fsub_rn(Su_xm[i], uc));
        sumU = __fmaf_rn(axp, __fsub_rn(Su_xp[i], uc), sumU);
        sumU = __fmaf_rn(ayp, __fsub_rn(comp(Su_yp, i), uc), sumU);
        sumU = __fmaf_rn(aym, __fsub_rn(comp(Su_ym, i), uc), sumU);
        sumU = __fmaf_rn(azp, __fsub_rn(comp(Nu, i), uc), sumU);
        sumU = __fmaf_rn(azm, __fsub_rn(comp(Pu, i), uc), sumU);
        float sumV = __fmul_rn(axm, __fsub_rn(Sv_xm[i], vc));
        sumV = __fmaf_rn(axp, __fsub_rn(Sv_xp[i], vc), sumV);
        sumV = __fmaf_rn(ayp, __fsub_rn(comp(Sv_yp, i), vc), sumV);
        sumV = __fmaf_rn(aym, __fsub_rn(comp(Sv_ym, i), vc), sumV);
        sumV = __fmaf_rn(azp, __fsub_rn(comp(Nv, i), vc), sumV);
        sumV = __fmaf_rn(azm, __fsub_rn(comp(Pv, i), vc), sumV);
        float sumW = __fmul_rn(axm, __fsub_rn(Sw_xm[i], wc));
        sumW = __fmaf_rn(axp, __fsub_rn(Sw_xp[i], wc), sumW);
        sumW = __fmaf_rn(ayp, __fsub_rn(comp(Sw_yp, i), wc), sumW);
        sumW = __fmaf_rn(aym, __fsub_rn(comp(Sw_ym, i), wc), sumW);
        sumW = __fmaf_rn(azp, __fsub_rn(comp(Nw, i), wc), sumW);
        sumW = __fmaf_rn(azm, __fsub_rn(comp(Pw, i), wc), sumW);
        const float dvc = comp(dv4, i), dwc = comp(dw4, i);
        float k = comp(ks4, i);
        if constexpr (KSI) {
          // data-term weight of this outer iteration from the iterate it starts with: the xi half of
          // compute_phi_ksi_3d (solve_3d.cu:250-260), operation for operation
          const float duc = comp(du4, i);
          const float r1 = __fadd_rn(J14, __fmaf_rn(J13, dwc, __fmaf_rn(J11, duc, __fmul_rn(J12, dvc))));
          const float r2 = __fadd_rn(J24, __fmaf_rn(J23, dwc, __fmaf_rn(J12, duc, __fmul_rn(J22, dvc))));
          const float r3 = __fadd_rn(J34, __fmaf_rn(J33, dwc, __fmaf_rn(J23, dvc, __fmul_rn(J13, duc))));
          const float r4 = __fmaf_rn(gt, gt, __fmaf_rn(J34, dwc, __fmaf_rn(J14, duc, __fmul_rn(J24, dvc))));
          float sv = __fadd_rn(__fmaf_rn(dwc, r3, __fmaf_rn(duc, r1, __fmul_rn(dvc, r2))), r4);
          sv = __fmul_rn(sv, (sv > 0.f) ? 1.f : 0.f);
          const float arg = __fmaf_rn(a.eps_d, a.eps_d, sv);
          bool ok2 = true;
          float sq2 = sqrt_fast(arg, ok2);
          k = rcp_fast(__fadd_rn(sq2, sq2), ok2);
          if (!ok2) {
            sq2 = __fsqrt_rn(arg);
            k = __frcp_rn(__fadd_rn(sq2, sq2));
          }
        }
        // solve_3d.cu:492-502; numerators as the reference's SASS evaluates them
        const float ndu = __fmaf_rn(-J13, dwc, __fmaf_rn(-J12, dvc, -J14));
        numU[i] = __fmaf_rn(k, ndu, sumU);
        denU[i] = __fmaf_rn(J11, k, sumH);
        denV[i] = __fmaf_rn(J22, k, sumH);
        denW[i] = __fmaf_rn(J33, k, sumH);
        rks[i] = k; sV[i] = sumV; sW[i] = sumW;
        j12[i] = J12; j13[i] = J13; j23[i] = J23; j24[i] = J24; j34[i] = J34; dwv[i] = dwc;
      }
      // phase 2: du' -> dv' -> dw' (each needs the previous quotient) as three rounds over the four voxels,
      // branch-free (common.cuh: div_fast), so the four dependent chains overlap
      float rdu[4], rdv[4], rdw[4];
      bool ok[4] = {true, true, true, true};
#pragma unroll
      for (int i = 0; i < 4; ++i) rdu[i] = div_fast(numU[i], denU[i], ok[i]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ndv = __fmaf_rn(-j23[i], dwv[i], __fmaf_rn(-j12[i], rdu[i], -j24[i]));
        rdv[i] = div_fast(__fmaf_rn(rks[i], ndv, sV[i]), denV[i], ok[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ndw = __fmaf_rn(-j23[i], rdv[i], __fmaf_rn(-j13[i], rdu[i], -j34[i]));
        rdw[i] = div_fast(__fmaf_rn(rks[i], ndw, sW[i]), denW[i], ok[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!ok[i]) {  // an operand outside the fast path's range (zero, tiny, huge, NaN): IEEE division
          const float3 r = slow_voxel(numU[i], denU[i], denV[i], denW[i], rks[i], sV[i], sW[i], j12[i], j13[i], j23[i],
                                      j24[i], j34[i], dwv[i]);
          rdu[i] = r.x; rdv[i] = r.y; rdw[i] = r.z;
        }
      }
      const unsigned o = (unsigned)(zb + q) * (unsigned)g.ps + row_g;
      st4(a.odu + o, make_float4(rdu[0], rdu[1], rdu[2], rdu[3]));
      st4(a.odv + o, make_float4(rdv[0], rdv[1], rdv[2], rdv[3]));
      st4(a.odw + o, make_float4(rdw[0], rdw[1], rdw[2], rdw[3]));
      if constexpr (KSI) st4(a.oksi + o, make_float4(rks[0], rks[1], rks[2], rks[3]));
    }
    // rotate the register-carried column
    Pu = Cu; Pv = Cv; Pw = Cw; Pp = Cp;
    Cu = Nu; Cv = Nv; Cw = Nw; Cp = Np;
    sb_c = sb_n; sb_n = (sb_n == 2) ? 0 : sb_n + 1;
    sd_c = (sd_c == 2) ? 0 : sd_c + 1;

    if (q + 2 <= n) {
      waitA(q + 2);
      prepass(q + 2);
    }
    fence_proxy_async();
    cta_sync<L::THREADS>();
  }
}

template <int BX, int BY, bool KSI>
__global__ void __launch_bounds__(TL<BX, BY, KSI>::THREADS, 2)
    sweep_tma_kernel(const __grid_constant__ TmaMaps maps, const SweepArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  sweep_tma_run<BX, BY, KSI>(maps, a, smem);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps (cached) and launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct MapKey {
  const void* p;
  int w, h, d, ld, bx, by;
  bool operator==(const MapKey& o) const {
    return p == o.p && w == o.w && h == o.h && d == o.d && ld == o.ld && bx == o.bx && by == o.by;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t x = reinterpret_cast<size_t>(k.p) * 0x9E3779B97F4A7C15ull;
    x ^= ((size_t)k.w << 40) ^ ((size_t)k.h << 20) ^ (size_t)k.d ^ ((size_t)k.ld << 50) ^ ((size_t)k.bx << 8) ^
         ((size_t)k.by << 16);
    return x * 0xC2B2AE3D27D4EB4Full;
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

static int get_map(const float* p, const Dims& g, int bx, int by, CUtensorMap* out) {
  const MapKey key{p, g.w, g.h, g.d, g.ld, bx, by};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return FLOW3D_OK; }
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return FLOW3D_ERR_UNSUPPORTED;
  const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)g.d};
  const cuuint64_t strides[2] = {(cuuint64_t)g.ld * 4, (cuuint64_t)g.ps * 4};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return FLOW3D_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return FLOW3D_OK;
}

// cudaFuncSetAttribute is per device: one flag per (kernel instance, device)
template <class K>
static int ensure_smem(K kernel, int bytes, unsigned long long* flags) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (*flags & (1ull << dev)) return FLOW3D_OK;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { note_cuda_error(e, "cudaFuncSetAttribute(sweep_tma_kernel)"); return FLOW3D_ERR_CUDA; }
  *flags |= 1ull << dev;
  return FLOW3D_OK;
}

template <int BX, int BY>
static int launch_shape(const SweepArgs& a0, int zchunk_len, cudaStream_t st) {
  SweepArgs a = a0;
  const Dims& g = a.g;
  const int nz = a.ze - a.zs;
  int len = zchunk_len > 0 ? zchunk_len : nz;
  if (len > nz) len = nz;
  a.zchunk = len;
  const bool ksi = a.oksi != nullptr;
  TmaMaps maps;
  std::memset(&maps, 0, sizeof(maps));
  const float* hal[7] = {a.u, a.v, a.w, a.du, a.dv, a.dw, a.phi};
  for (int i = 0; i < 7; ++i) {
    const int rc = get_map(hal[i], g, BX + 8, BY + 2, &maps.m[M_U + i]);
    if (rc != FLOW3D_OK) return rc;
  }
  const float* cen[5] = {a.fx, a.fy, a.fz, a.ft, a.ksi};
  for (int i = 0; i < (ksi ? 4 : 5); ++i) {
    const int rc = get_map(cen[i], g, BX, BY, &maps.m[M_FX + i]);
    if (rc != FLOW3D_OK) return rc;
  }
  const dim3 grid((g.w + BX - 1) / BX, (g.h + BY - 1) / BY, (nz + len - 1) / len);
  static unsigned long long flag_plain = 0, flag_ksi = 0;
  if (ksi) {
    using L = TL<BX, BY, true>;
    const int rc = ensure_smem(sweep_tma_kernel<BX, BY, true>, L::BYTES, &flag_ksi);
    if (rc != FLOW3D_OK) return rc;
    sweep_tma_kernel<BX, BY, true><<<grid, L::THREADS, L::BYTES, st>>>(maps, a);
  } else {
    using L = TL<BX, BY, false>;
    const int rc = ensure_smem(sweep_tma_kernel<BX, BY, false>, L::BYTES, &flag_plain);
    if (rc != FLOW3D_OK) return rc;
    sweep_tma_kernel<BX, BY, false><<<grid, L::THREADS, L::BYTES, st>>>(maps, a);
  }
  count_launch();
  return check_launch("sweep_tma_kernel");
}

bool sweep_tma_usable(const Dims& g, int variant) {
  if (!sweep_variant_is_tma(variant)) return false;
  if (g.w < 8 || g.h < 2 || (g.ld & 3)) return false;
  const int by = variant == SWEEP_VARIANT_TMA_64x8 ? 8 : 16;
  return (g.h + by - 1) / by <= 65535 && encode_fn() != nullptr;
}

int launch_sweep_tma(const SweepArgs& a, int variant, int zchunk_len, cudaStream_t st) {
  if (!sweep_tma_usable(a.g, variant)) return FLOW3D_ERR_UNSUPPORTED;
  if (a.ze <= a.zs) return FLOW3D_OK;
  if (variant == SWEEP_VARIANT_TMA_64x8) return launch_shape<64, 8>(a, zchunk_len, st);
  return launch_shape<32, 16>(a, zchunk_len, st);
}

}  // namespace f3d
