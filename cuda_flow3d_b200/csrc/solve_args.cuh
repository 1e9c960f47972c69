// solve_args.cuh -- argument block and launch-shape descriptor shared by the two Jacobi-sweep kernels
// (the register-marching warp kernel in kernels_solve.cu and the TMA-staged tile kernel in
// kernels_sweep_tma.cu).  Both give identical bits; the tuner picks per level.
#pragma once
#include "common.cuh"

namespace f3d {

struct SweepArgs {
  const float *fx, *fy, *fz, *ft;
  const float *u, *v, *w;
  const float *du, *dv, *dw;
  const float *phi, *ksi;
  float *odu, *odv, *odw;
  Dims g;
  float hx, hy, hz, alpha;
  int zchunk;
  int pf;  // prefetch distance in planes (0 = off)
  int zs, ze;  // compute range (local planes)
  int lpr;     // lanes per row segment (32, 16 or 8)
  int pf1;     // L1 prefetch of the next plane (0 = off)
  int wx;      // warps side by side along x within a block
  float* oksi; // KSI variant: the data-term weight is computed here (not read from `ksi`) and stored
  float eps_d;
};

// launch shape of a z-marching solver kernel (results do not depend on it)
struct TuneCfg {
  int vec;      // register kernel: voxels per lane (1, 2, 4)
  int nchunks;  // z chunks (0 = static heuristic)
  int variant;  // SWEEP_VARIANT_*
};

enum { SWEEP_VARIANT_REG = 0, SWEEP_VARIANT_TMA_64x8 = 1, SWEEP_VARIANT_TMA_32x16 = 2, SWEEP_VARIANT_COUNT };
inline bool sweep_variant_is_tma(int v) { return v == SWEEP_VARIANT_TMA_64x8 || v == SWEEP_VARIANT_TMA_32x16; }

// TMA-staged sweep (kernels_sweep_tma.cu); returns FLOW3D_ERR_UNSUPPORTED when the level cannot use it
int launch_sweep_tma(const SweepArgs& a, int variant, int zchunk_len, cudaStream_t st);
bool sweep_tma_usable(const Dims& g, int variant);

}  // namespace f3d
