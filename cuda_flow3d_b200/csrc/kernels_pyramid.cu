// kernels_pyramid.cu -- Gaussian pre-blur and box resampling (pyramid construction, flow
// prolongation) for sm_100a.  Arithmetic is transcribed operation by operation from the reference
// kernels' PTX (see kernels_solve.cu for the contract).
#include <cstdlib>

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

// float4 forms (north-star (a): vectorised, coalesced rows).  y / z passes: a thread owns four consecutive
// x and reads one float4 per tap.  x pass (compile-time radius R): a thread owns four consecutive outputs
// and reads its 2R+4-wide window once.  Per output the operations are the scalar kernel's: one fma per tap,
// j ascending, zero for taps outside the volume.  Columns >= w of a row are padding: whatever is computed
// there is never interpreted.
template <int AXIS>
__global__ void __launch_bounds__(256) conv_axis_vec4_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             Dims g, ConvTaps taps, int radius, int zs) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zs + blockIdx.z;
  if (x >= g.w || y >= g.h) return;
  const long long c = (long long)z * g.ps + (long long)y * g.ld + x;
  const int pos = AXIS == 1 ? y : (g.z0g + z);
  const int n = AXIS == 1 ? g.h : g.dg;
  const long long stride = AXIS == 1 ? (long long)g.ld : g.ps;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int j = -radius; j <= radius; ++j) {
    const int p = pos + j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p >= 0 && p < n) v = __ldg(reinterpret_cast<const float4*>(in + c + (long long)j * stride));
    const float t = taps.t[radius - j];
    s0 = __fmaf_rn(t, v.x, s0);
    s1 = __fmaf_rn(t, v.y, s1);
    s2 = __fmaf_rn(t, v.z, s2);
    s3 = __fmaf_rn(t, v.w, s3);
  }
  *reinterpret_cast<float4*>(out + c) = make_float4(s0, s1, s2, s3);
}

template <int R>
__global__ void __launch_bounds__(256) conv_x_vec4_kernel(const float* __restrict__ in, float* __restrict__ out, Dims g,
                                                          ConvTaps taps, int zs) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zs + blockIdx.z;
  if (x >= g.w || y >= g.h) return;
  const long long row = (long long)z * g.ps + (long long)y * g.ld;
  float win[2 * R + 4];  // in[x-R .. x+R+3], zero outside [0, w)
#pragma unroll
  for (int k = 0; k < 2 * R + 4; ++k) {
    const int p = x - R + k;
    win[k] = (p >= 0 && p < g.w) ? __ldg(in + row + p) : 0.f;
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = -R; j <= R; ++j) {
    const float t = taps.t[R - j];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = __fmaf_rn(t, win[i + j + R], s[i]);
  }
  *reinterpret_cast<float4*>(out + row + x) = make_float4(s[0], s[1], s[2], s[3]);
}

int launch_conv_axis(const float* in, float* out, Dims g, const float* taps_host, int radius,
                     int axis, ZRange zr, cudaStream_t st) {
  if (radius < 0 || radius > F3D_MAX_BLUR_RADIUS) return FLOW3D_ERR_UNSUPPORTED;
  if (zr.end <= zr.begin) return FLOW3D_OK;
  ConvTaps t;
  for (int i = 0; i < 2 * radius + 1; ++i) t.t[i] = taps_host[i];
  for (int i = 2 * radius + 1; i < 2 * F3D_MAX_BLUR_RADIUS + 1; ++i) t.t[i] = 0.f;
  dim3 block(32, 8, 1);
  static const int scalar = [] { const char* e = getenv("FLOW3D_BLUR_SCALAR"); return (e && *e) ? atoi(e) : 0; }();
  const bool vec_ok = !scalar && (g.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (vec_ok && (axis != 0 || radius == 6 || radius == 3)) {  // sigma = 2 (default) -> 6, sigma = 1 -> 3
    dim3 grid4((g.w + 127) / 128, (g.h + 7) / 8, zr.end - zr.begin);
    if (axis == 1) conv_axis_vec4_kernel<1><<<grid4, block, 0, st>>>(in, out, g, t, radius, zr.begin);
    else if (axis == 2) conv_axis_vec4_kernel<2><<<grid4, block, 0, st>>>(in, out, g, t, radius, zr.begin);
    else if (radius == 6) conv_x_vec4_kernel<6><<<grid4, block, 0, st>>>(in, out, g, t, zr.begin);
    else conv_x_vec4_kernel<3><<<grid4, block, 0, st>>>(in, out, g, t, zr.begin);
    count_launch();
    return check_launch("conv_axis_vec4_kernel");
  }
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

// Tap geometry of one output index (resample_3d.cu:41-48), shared by the kernels below
struct ResampleTaps {
  float delta, norm, left_f, right_f, first;
  int left_i, cnt;
};
__device__ __forceinline__ ResampleTaps resample_taps(unsigned o, unsigned long long a, unsigned long long b) {
  ResampleTaps t;
  const float fa = (float)a, fb = (float)b;
  t.delta = __fdiv_rn(fa, fb);
  t.norm = __fdiv_rn(fb, fa);
  t.left_f = __fmul_rn((float)o, t.delta);
  t.right_f = __fmul_rn((float)(o + 1u), t.delta);
  t.left_i = (int)floorf(t.left_f);
  const int right_i = (int)fminf(fa, (float)(unsigned long long)ceilf(t.right_f));
  t.cnt = right_i - t.left_i;
  t.first = __fsub_rn((float)(t.left_i + 1), t.left_f);
  return t;
}
__device__ __forceinline__ float resample_frac(const ResampleTaps& t, int j) {
  float frac = 1.f;
  if (j == 0) frac = t.first;
  if (j == t.cnt - 1) frac = __fsub_rn(t.right_f, (float)(t.left_i + j));
  if (t.cnt == 1) frac = t.delta;
  return frac;
}

// y / z pass, four consecutive x per thread: every tap is one float4 load (rows are 16-byte aligned and
// padded to a multiple of four floats), and the first three taps are loaded before the first fma so
// the loads overlap.  Same per-element arithmetic and tap order as resample_axis_kernel.
template <int AXIS>
__global__ void __launch_bounds__(256) resample_axis_vec4_kernel(const float* __restrict__ in, Dims gi,
                                                                 float* __restrict__ out, Dims go, int zs) {
  static_assert(AXIS == 1 || AXIS == 2, "x pass is scalar");
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zs + blockIdx.z;
  if (x >= go.w || y >= go.h) return;
  const unsigned o = AXIS == 1 ? y : (go.z0g + z);
  const ResampleTaps t = resample_taps(o, AXIS == 1 ? gi.h : gi.dg, AXIS == 1 ? go.h : go.dg);
  const long long base = AXIS == 1 ? ((long long)z * gi.ps + (long long)t.left_i * gi.ld + x)
                                   : ((long long)(t.left_i - gi.z0g) * gi.ps + (long long)y * gi.ld + x);
  const long long stride = AXIS == 1 ? (long long)gi.ld : gi.ps;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto tap = [&](int j) { return __ldg(reinterpret_cast<const float4*>(in + base + (long long)j * stride)); };
  auto fma4 = [&](float f, const float4& v) {
    acc.x = __fmaf_rn(f, v.x, acc.x); acc.y = __fmaf_rn(f, v.y, acc.y);
    acc.z = __fmaf_rn(f, v.z, acc.z); acc.w = __fmaf_rn(f, v.w, acc.w);
  };
  const int cnt = t.cnt;
  const float4 v0 = tap(0);
  const float4 v1 = tap(cnt > 1 ? 1 : 0);
  const float4 v2 = tap(cnt > 2 ? 2 : 0);
  if (cnt > 0) fma4(resample_frac(t, 0), v0);
  if (cnt > 1) fma4(resample_frac(t, 1), v1);
  if (cnt > 2) fma4(resample_frac(t, 2), v2);
  for (int j = 3; j < cnt; ++j) fma4(resample_frac(t, j), tap(j));
  float4 r;
  r.x = __fmul_rn(t.norm, acc.x); r.y = __fmul_rn(t.norm, acc.y);
  r.z = __fmul_rn(t.norm, acc.z); r.w = __fmul_rn(t.norm, acc.w);
  *reinterpret_cast<float4*>(out + (long long)z * go.ps + (long long)y * go.ld + x) = r;
}

// x pass: one output per thread, the first three taps loaded up front (fine levels have <= 3 taps)
__global__ void __launch_bounds__(256) resample_x_kernel(const float* __restrict__ in, Dims gi,
                                                         float* __restrict__ out, Dims go, int zs) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int z = zs + blockIdx.z;
  if (x >= go.w || y >= go.h) return;
  const ResampleTaps t = resample_taps((unsigned)x, gi.w, go.w);
  const float* p = in + (long long)z * gi.ps + (long long)y * gi.ld + t.left_i;
  const int cnt = t.cnt;
  const float v0 = __ldg(p);
  const float v1 = __ldg(p + (cnt > 1 ? 1 : 0));
  const float v2 = __ldg(p + (cnt > 2 ? 2 : 0));
  float value = 0.f;
  if (cnt > 0) value = __fmaf_rn(resample_frac(t, 0), v0, value);
  if (cnt > 1) value = __fmaf_rn(resample_frac(t, 1), v1, value);
  if (cnt > 2) value = __fmaf_rn(resample_frac(t, 2), v2, value);
  for (int j = 3; j < cnt; ++j) value = __fmaf_rn(resample_frac(t, j), __ldg(p + j), value);
  out[(long long)z * go.ps + (long long)y * go.ld + x] = __fmul_rn(t.norm, value);
}

int launch_resample_axis(const float* in, Dims gin, float* out, Dims gout, int axis, ZRange zr,
                         cudaStream_t st) {
  if (zr.end <= zr.begin) return FLOW3D_OK;
  dim3 block(32, 8, 1);
  dim3 grid((gout.w + 31) / 32, (gout.h + 7) / 8, zr.end - zr.begin);
  static const bool scalar_only = [] { const char* e = getenv("FLOW3D_RESAMPLE_SCALAR"); return e && *e == '1'; }();
  const bool vec_ok = !scalar_only && (gin.ld % 4 == 0) && (gout.ld % 4 == 0) && (gin.ps % 4 == 0) && (gout.ps % 4 == 0) &&
                      aligned16(in) && aligned16(out);
  if (axis != 0 && vec_ok) {
    dim3 g4((gout.w + 127) / 128, (gout.h + 7) / 8, zr.end - zr.begin);
    if (axis == 1) resample_axis_vec4_kernel<1><<<g4, block, 0, st>>>(in, gin, out, gout, zr.begin);
    else resample_axis_vec4_kernel<2><<<g4, block, 0, st>>>(in, gin, out, gout, zr.begin);
  } else if (axis == 0 && !scalar_only) resample_x_kernel<<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  else if (axis == 0) resample_axis_kernel<0><<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  else if (axis == 1) resample_axis_kernel<1><<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  else resample_axis_kernel<2><<<grid, block, 0, st>>>(in, gin, out, gout, zr.begin);
  count_launch();
  return check_launch("resample_axis_kernel");
}

}  // namespace f3d
