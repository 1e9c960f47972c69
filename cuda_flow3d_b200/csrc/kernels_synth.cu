// kernels_synth.cu -- analytic synthetic volume pair with a known rigid motion (SURVEY.md 8d,
// configs 3-5): a band-limited texture T(x) = 127.5 + S*sum_k A_k sin(2*pi f_k.x + p_k) evaluated in
// double precision on the device; frame_0 = T, frame_1(y) = T(R^T (y - c - t) + c); ground-truth
// flow (reference convention f1(x + flow) = f0(x), registration_3d.cu:47-49) = (R - I)(x - c) + t.
#include <cmath>
#include <random>

#include "common.cuh"

namespace f3d {

#define SYNTH_WAVES 32
struct SynthParams {
  double fx[SYNTH_WAVES], fy[SYNTH_WAVES], fz[SYNTH_WAVES];  // cycles / voxel
  double amp[SYNTH_WAVES], phase[SYNTH_WAVES];
  double scale;        // S
  double R[9];         // rotation (row major)
  double c[3], t[3];
};

static SynthParams make_params(size_t W, size_t H, size_t D, uint64_t seed) {
  SynthParams p;
  std::mt19937_64 rng(seed);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_real_distribution<double> ud(0.0, 1.0);
  const double two_pi = 6.283185307179586476925286766559;
  double asum = 0.0;
  for (int k = 0; k < SYNTH_WAVES; ++k) {
    double dx = nd(rng), dy = nd(rng), dz = nd(rng);
    double n = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (n == 0.0) { dx = 1.0; n = 1.0; }
    const double mag = 1.0 / 64.0 + (1.0 / 8.0 - 1.0 / 64.0) * ud(rng);
    p.fx[k] = dx / n * mag;
    p.fy[k] = dy / n * mag;
    p.fz[k] = dz / n * mag;
    p.amp[k] = 0.5 + 0.5 * ud(rng);
    p.phase[k] = two_pi * ud(rng);
    asum += p.amp[k];
  }
  p.scale = 127.5 / asum;
  // rotation by theta about (1,1,1)/sqrt(3); theta shrinks with size so the peak displacement
  // stays ~5.6 voxels
  const double theta = 0.5 * (512.0 / (double)W) * (two_pi / 360.0);
  const double a = 1.0 / std::sqrt(3.0), cs = std::cos(theta), sn = std::sin(theta), oc = 1.0 - cs;
  const double ax[3] = {a, a, a};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) p.R[i * 3 + j] = oc * ax[i] * ax[j] + (i == j ? cs : 0.0);
  p.R[0 * 3 + 1] -= sn * ax[2]; p.R[0 * 3 + 2] += sn * ax[1];
  p.R[1 * 3 + 0] += sn * ax[2]; p.R[1 * 3 + 2] -= sn * ax[0];
  p.R[2 * 3 + 0] -= sn * ax[1]; p.R[2 * 3 + 1] += sn * ax[0];
  p.c[0] = ((double)W - 1.0) / 2.0;
  p.c[1] = ((double)H - 1.0) / 2.0;
  p.c[2] = ((double)D - 1.0) / 2.0;
  p.t[0] = 1.5; p.t[1] = -1.0; p.t[2] = 0.75;
  return p;
}

__device__ __forceinline__ double texture_at(const SynthParams& p, double x, double y, double z) {
  const double two_pi = 6.283185307179586476925286766559;
  double s = 0.0;
#pragma unroll 4
  for (int k = 0; k < SYNTH_WAVES; ++k)
    s += p.amp[k] * sin(two_pi * (p.fx[k] * x + p.fy[k] * y + p.fz[k] * z) + p.phase[k]);
  return 127.5 + p.scale * s;
}

__global__ void __launch_bounds__(256) synth_kernel(const SynthParams p, int W, int H, int z0, int nz,
                                                    int ld, float* f0, float* f1, float* tu, float* tv,
                                                    float* tw) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int zl = blockIdx.z;
  if (x >= W || y >= H || zl >= nz) return;
  const int z = z0 + zl;
  const long long c = ((long long)zl * H + y) * ld + x;
  const double px = x, py = y, pz = z;
  if (f0) f0[c] = (float)texture_at(p, px, py, pz);
  const double qx = px - p.c[0], qy = py - p.c[1], qz = pz - p.c[2];
  if (f1) {
    // R^T (y - c - t) + c
    const double rx = qx - p.t[0], ry = qy - p.t[1], rz = qz - p.t[2];
    const double sx = p.R[0] * rx + p.R[3] * ry + p.R[6] * rz + p.c[0];
    const double sy = p.R[1] * rx + p.R[4] * ry + p.R[7] * rz + p.c[1];
    const double sz = p.R[2] * rx + p.R[5] * ry + p.R[8] * rz + p.c[2];
    f1[c] = (float)texture_at(p, sx, sy, sz);
  }
  if (tu) tu[c] = (float)((p.R[0] - 1.0) * qx + p.R[1] * qy + p.R[2] * qz + p.t[0]);
  if (tv) tv[c] = (float)(p.R[3] * qx + (p.R[4] - 1.0) * qy + p.R[5] * qz + p.t[1]);
  if (tw) tw[c] = (float)(p.R[6] * qx + p.R[7] * qy + (p.R[8] - 1.0) * qz + p.t[2]);
}

int launch_synth(size_t W, size_t H, size_t D, size_t z0, size_t nz, size_t ld, uint64_t seed,
                 float* f0, float* f1, float* tu, float* tv, float* tw, cudaStream_t st) {
  if (z0 + nz > D || nz == 0) return FLOW3D_ERR_INVALID_ARG;
  const SynthParams p = make_params(W, H, D, seed);
  dim3 block(32, 8, 1);
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)nz);
  synth_kernel<<<grid, block, 0, st>>>(p, (int)W, (int)H, (int)z0, (int)nz, (int)ld, f0, f1, tu, tv, tw);
  count_launch();
  return check_launch("synth_kernel");
}

}  // namespace f3d
