// kernels_pyramid.cu -- Gaussian pre-blur and box resampling (pyramid construction, flow
// prolongation) for sm_100a.  Arithmetic is transcribed operation by operation from the reference
// kernels' PTX (see kernels_solve.cu for the contract).
#include "common.cuh"

namespace f3d {

// ------------------------------------------------------------------------------------------------
// separable Gaussian, zero padding (convolution_3d.cu:161-168, 254-262, 358-366)
// ------------------------------------------------------------------------------------------------
#define F3D_MAX_BLUR_RADIUS 32
struct ConvTaps {
  float t[2 * F3D_MAX_BLUR_RADIUS + 1];
};

// One output voxel per thread; the 2r+1 taps are read through L1 (neighbouring threads share
// them).  Accumulation order and the fma-per-tap form follow the reference exactly:
// sum = fma(k[r-j], in[c+j], sum) for j = -r..r, starting from +0.
template <int AXIS>
__global__ void __launch_bounds__(256) conv_axis_kernel(const float* __restrict__ in,
                                                        float* __restrict__ out, Dims g,
                                                        ConvTaps taps, int radius, int zs) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zs + blockIdx.z;
  if (x >= g.w || y >= g.h) return;
  const long long c = (long long)z * g.ps + (long long)y * g.ld + x;
  // along z the zero padding starts at the GLOBAL faces (z-slabs: every tap inside the level must
  // be a plane of the buffer -- the caller provides `radius` ghost planes)
  const int pos = AXIS == 0 ? x : (AXIS == 1 ? y : (g.z0g + z));
  const int n = AXIS == 0 ? g.w : (AXIS == 1 ? g.h : g.dg);
  const long long stride = AXIS == 0 ? 1 : (AXIS == 1 ? (long long)g.ld : g.ps);
  float sum = 0.f;
  for (int j = -radius; j <= radius; ++j) {
    const int p = pos + j;
    const float v = (p >= 0 && p < n) ? __ldg(in + c + (long long)j * stride) : 0.f;
    sum = __fmaf_rn(taps.t[radius - j], v, sum);
  }
  out[c] = sum;
}

int launch_conv_axis(const float* in, float* out, Dims g, const float* taps_host, int radius,
                     int axis, ZRange zr, cudaStream_t st) {
  if (radius < 0 || radius > F3D_MAX_BLUR_RADIUS) return FLOW3D_ERR_UNSUPPORTED;
  if (zr.end <= zr.begin) return FLOW3D_OK;
  ConvTaps t;
  for (int i = 0; i < 2 * radius + 1; ++i) t.t[i] = taps_host[i];
  for (int i = 2 * radius + 1; i < 2 * F3D_MAX_BLUR_RADIUS + 1; ++i) t.t[i] = 0.f;
  dim3 block(32, 8, 1);
  dim3 grid((g.w + 31) / 32, (g.h + 7) / 8, zr.end - zr.begin);
  if (axis == 0) conv_axis_kernel<0><<<grid, block, 0, st>>>(in, out, g, t, radius, zr.begin);
  else if (axis == 1) conv_axis_kernel<1><<<grid, block, 0, st>>>(in, out, g, t, radius, zr.begin);
  else conv_axis_kernel<2><<<grid, block, 0, st>>>(in, out, g, t, radius, zr.begin);
  count_launch();
  return check_launch("conv_axis_kernel");
}

// ------------------------------------------------------------------------------------------------
// box (area-average) resample along one axis (resample_3d.cu:41-69)
// ------------------------------------------------------------------------------------------------
template <int AXIS>
__global__ void __launch_bounds__(256) resample_axis_kernel(const float* __restrict__ in, Dims gi,
                                                            float* __restrict__ out, Dims go, int zs) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zs + blockIdx.z;  // local output plane
  if (x >= go.w || y >= go.h) return;
  // along z the tap arithmetic runs in GLOBAL plane indices (slabs: z0g/dg), then maps to local
  const unsigned o = AXIS == 0 ? x : (AXIS == 1 ? y : (go.z0g + z));
  const unsigned long long a = AXIS == 0 ? gi.w : (AXIS == 1 ? gi.h : gi.dg);
  const unsigned long long b = AXIS == 0 ? go.w : (AXIS == 1 ? go.h : go.dg);
  const float fa = (float)a, fb = (float)b;
  const float delta = __fdiv_rn(fa, fb);
  const float norm = __fdiv_rn(fb, fa);
  const float left_f = __fmul_rn((float)o, delta);
  const float right_f = __fmul_rn((float)(o + 1u), delta);
  const int left_i = (int)floorf(left_f);
  const int right_i = (int)fminf(fa, (float)(unsigned long long)ceilf(right_f));
  const int cnt = right_i - left_i;
  // input element (left_i + j) along AXIS, same other coordinates
  const long long base = AXIS == 0 ? ((long long)z * gi.ps + (long long)y * gi.ld + left_i)
                       : AXIS == 1 ? ((long long)z * gi.ps + (long long)left_i * gi.ld + x)
                                   : ((long long)(left_i - gi.z0g) * gi.ps + (long long)y * gi.ld + x);
  const long long stride = AXIS == 0 ? 1 : (AXIS == 1 ? (long long)gi.ld : gi.ps);
  const float first = __fsub_rn((float)(left_i + 1), left_f);
  float value = 0.f;
  for (int j = 0; j < cnt; ++j) {
    float frac = 1.f;
    if (j == 0) frac = first;
    if (j == cnt - 1) frac = __fsub_rn(right_f, (float)(left_i + j));
    if (cnt == 1) frac = delta;
    value = __fmaf_rn(frac, __ldg(in + base + (long long)j * stride), value);
  }
  out[(long long)z * go.ps + (long long)y * go.ld + x] = __fmul_rn(norm, value);
}

int launch_resample_axis(const float* in, Dims gin, float* out, Dims gout, int axis, ZRange zr,
                         cudaStream_t st) {
  if (zr.end <= zr.begin) return FLOW3D_OK;
  dim3 block(32, 8, 1);
  dim3 grid((gout.w + 31) / 32, (gout.h + 7) / 8, zr.end - zr.begin);
  if (axis == 0) resample_axis_kernel<0><<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  else if (axis == 1) resample_axis_kernel<1><<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  else resample_axis_kernel<2><<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  count_launch();
  return check_launch("resample_axis_kernel");
}

}  // namespace f3d
