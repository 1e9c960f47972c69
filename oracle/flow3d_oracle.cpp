// flow3d_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Scalar C++ restatement, on the CPU, of the dense 3D variational optical-flow solve of
// axruff/cuda-flow3d (the reference has no CPU path; its algorithm lives in CUDA kernels).
// Every function cites the reference file:line it follows (paths relative to /root/reference).
//
// Parity status: the reference ships no golden vectors or tests, so this
// oracle is pinned against outputs of the reference's own CUDA build run on a B200
// (oracle/build_ref.sh -> oracle/_ref, fixtures under tests/golden/, see tests/golden/README.md).
//
// Numerics: IEEE fp32.  The reference kernels are compiled by `nvcc -ptx` with FMA contraction ON
// (Makefile:6), and 8,000 nonlinear sweeps + medians amplify a 1-ulp difference to ~1e-2 voxel
// (measured: DESIGN.md "sensitivity"), i.e. beyond the 1e-3 parity gate.  So this file states the
// arithmetic at the level of individual rounded operations: every std::fmaf below is a place
// where the reference's PTX (nvcc 12.9, `-ptx -std=c++11`, file oracle/_ref/kernels/*.ptx) holds an
// fma.rn.f32, every other +,-,* is a separately rounded operation (compile with -ffp-contract=off
// so the host compiler adds none of its own), and / and sqrt are IEEE (div.rn / sqrt.rn / rcp.rn).
// With that the oracle is intended to be BIT-EXACT with the reference's CUDA build.
//
// Layout: compact x-fastest volumes, index = (z*h + y)*w + x (the reference's pitched
// "sub-box of a full-size container" layout is a pure storage choice).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load this.

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

inline size_t IDX(size_t x, size_t y, size_t z, size_t w, size_t h) { return (z * h + y) * w + x; }

// reflect-101 mirror used by solve_3d.cu:73-75,89-90,104-105 and median_3d.cu:70-72
inline long mirror(long i, long n) {
  if (i < 0) return -i;
  if (i >= n) return 2 * n - i - 2;
  return i;
}

// z-slab view (multi-GPU sharding): local plane zl of a buffer whose plane 0 is global plane z0g of
// a level of global depth dg; reflect-101 applies at the global faces only.
struct Slab {
  long z0g, dg, zs, ze;
};
inline long zn(const Slab& s, long zl, long dz) { return mirror(s.z0g + zl + dz, s.dg) - s.z0g; }

}  // namespace

extern "C" {

int o_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void o_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// optical_flow_base.cpp:31-56
size_t o_max_warp_level(size_t width, size_t height, size_t depth, float scale_factor) {
  size_t r_width = 1, r_height = 1, r_depth = 1;
  size_t level_counter = 1;
  while (scale_factor < 1.f) {
    float scale = std::pow(scale_factor, static_cast<float>(level_counter));
    r_width = static_cast<size_t>(std::ceil(width * scale));
    r_height = static_cast<size_t>(std::ceil(height * scale));
    r_depth = static_cast<size_t>(std::ceil(depth * scale));
    if (r_width < 4 || r_height < 4 || r_depth < 4) break;
    ++level_counter;
  }
  if (r_width == 1 || r_height == 1 || r_depth == 1) --level_counter;
  return level_counter;
}

// optical_flow_e.cpp:262-268
void o_level_geometry(size_t W, size_t H, size_t D, float scale_factor, int level, size_t* dims,
                      float* h) {
  float scale = std::pow(scale_factor, static_cast<float>(level));
  dims[0] = static_cast<size_t>(std::ceil(W * scale));
  dims[1] = static_cast<size_t>(std::ceil(H * scale));
  dims[2] = static_cast<size_t>(std::ceil(D * scale));
  h[0] = W / static_cast<float>(dims[0]);
  h[1] = H / static_cast<float>(dims[1]);
  h[2] = D / static_cast<float>(dims[2]);
}

// cuda_operation_convolution.cpp:85-108 (precision = 3, pixel_size = 1.0 at :159)
// taps must hold >= 2*radius+1 floats; returns the radius.
int o_gauss_taps(float sigma, float* taps) {
  const size_t precision = 3;
  const float pixel_size = 1.0f;
  size_t radius = (size_t)(precision * sigma / pixel_size);
  int r = static_cast<int>(radius);
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

// One separable pass of convolution_3d.cu:161-168 along `axis` (0=x rows :75-172, 1=y columns
// :186-271, 2=z slices :284-372); zero padding outside the volume (:113,122,130); accumulation in
// ascending j with tap index radius-j.  This is the GUARDED semantics.
void o_conv_axis(const float* in, float* out, size_t w, size_t h, size_t d, const float* taps,
                 int radius, int axis) {
  const long n[3] = {(long)w, (long)h, (long)d};
#pragma omp parallel for schedule(static)
  for (long z = 0; z < (long)d; ++z)
    for (long y = 0; y < (long)h; ++y)
      for (long x = 0; x < (long)w; ++x) {
        long c[3] = {x, y, z};
        float sum = 0;
        for (int j = -radius; j <= radius; j++) {
          long p[3] = {x, y, z};
          p[axis] = c[axis] + j;
          float v = (p[axis] >= 0 && p[axis] < n[axis]) ? in[IDX(p[0], p[1], p[2], w, h)] : 0.f;
          sum = std::fmaf(taps[radius - j], v, sum);  // fma.rn in convolution_3d.ptx
        }
        out[IDX(x, y, z, w, h)] = sum;
      }
}

// z pass on a slab: local plane 0 = global z0g of a volume of global depth dg; zero padding at the
// global faces only; computes local planes [zs, ze)
void o_conv_z_slab(const float* in, float* out, size_t w, size_t h, long z0g, long dg, long zs, long ze,
                   const float* taps, int radius) {
#pragma omp parallel for schedule(static)
  for (long z = zs; z < ze; ++z)
    for (long y = 0; y < (long)h; ++y)
      for (long x = 0; x < (long)w; ++x) {
        float sum = 0;
        for (int j = -radius; j <= radius; j++) {
          const long zg = z0g + z + j;
          float v = (zg >= 0 && zg < dg) ? in[IDX(x, y, z + j, w, h)] : 0.f;
          sum = std::fmaf(taps[radius - j], v, sum);
        }
        out[IDX(x, y, z, w, h)] = sum;
      }
}

// cuda_operation_convolution.cpp:159-180: rows(in->out), columns(out->tmp), slices(tmp->out)
void o_gauss_blur(const float* in, float* out, float* tmp, size_t w, size_t h, size_t d,
                  float sigma) {
  float taps[128];
  int r = o_gauss_taps(sigma, taps);
  o_conv_axis(in, out, w, h, d, taps, r, 0);
  o_conv_axis(out, tmp, w, h, d, taps, r, 1);
  o_conv_axis(tmp, out, w, h, d, taps, r, 2);
}

// resample_3d.cu:41-69 for one output index `o` along an axis of input size `a`, output size `b`.
// Emits the (first input index, count, weights[]) of the box filter; returns normalisation.
static inline float resample_taps(size_t a, size_t b, unsigned o, int* left_i_out, int* cnt_out,
                                  float* frac /*>= cnt*/, int max_cnt) {
  float delta = a / static_cast<float>(b);
  float normalization = b / static_cast<float>(a);
  float left_f = o * delta;
  float right_f = (o + 1) * delta;
  int left_i = static_cast<int>(std::floor(left_f));
  int right_i = std::fmin((float)a, (float)static_cast<size_t>(std::ceil(right_f)));
  int cnt = right_i - left_i;
  for (int j = 0; j < cnt && j < max_cnt; j++) {
    float f = 1.f;
    if (j == 0) f = static_cast<float>(left_i + 1) - left_f;
    if (j == cnt - 1) f = right_f - static_cast<float>(left_i + j);
    if (cnt == 1) f = delta;
    frac[j] = f;
  }
  *left_i_out = left_i;
  *cnt_out = cnt;
  return normalization;
}

// resample_{x,y,z}_3d (resample_3d.cu:28-161): `axis` resampled from in_dims[axis] to out_n; the
// other two extents are unchanged.
void o_resample_axis(const float* in, const size_t* in_dims, float* out, size_t out_n, int axis) {
  size_t od[3] = {in_dims[0], in_dims[1], in_dims[2]};
  od[axis] = out_n;
  const size_t a = in_dims[axis], b = out_n;
  const size_t iw = in_dims[0], ih = in_dims[1];
  const size_t ow = od[0], oh = od[1], odp = od[2];
#pragma omp parallel for schedule(static)
  for (long z = 0; z < (long)odp; ++z) {
    std::vector<float> frac(a + 4);
    for (size_t y = 0; y < oh; ++y)
      for (size_t x = 0; x < ow; ++x) {
        size_t c[3] = {x, y, (size_t)z};
        int li, cnt;
        float norm = resample_taps(a, b, (unsigned)c[axis], &li, &cnt, frac.data(), (int)a + 4);
        float value = 0.f;
        for (int j = 0; j < cnt; j++) {
          size_t p[3] = {x, y, (size_t)z};
          p[axis] = (size_t)(li + j);
          value = std::fmaf(frac[j], in[IDX(p[0], p[1], p[2], iw, ih)], value);  // fma.rn
        }
        out[IDX(x, y, z, ow, oh)] = value * norm;
      }
  }
}

// cuda_operation_resample.cpp:95-105: X (in->out), Y (out->tmp), Z (tmp->out).
// tmp_a, tmp_b: scratch of >= max(in,out) voxels each.
void o_resample(const float* in, const size_t* in_dims, float* out, const size_t* out_dims,
                float* tmp_a, float* tmp_b) {
  size_t d0[3] = {in_dims[0], in_dims[1], in_dims[2]};
  o_resample_axis(in, d0, tmp_a, out_dims[0], 0);
  d0[0] = out_dims[0];
  o_resample_axis(tmp_a, d0, tmp_b, out_dims[1], 1);
  d0[1] = out_dims[1];
  o_resample_axis(tmp_b, d0, out, out_dims[2], 2);
}

// registration_3d.cu:46-80.  Slab form: f0,u,v,w,out are slabs (local plane 0 = global z0g) of a level of
// global depth `depth`; f1 is its own slab starting at global plane f1_z0g.  Computes local planes
// [zs, ze).  The plain form is the slab form with z0g = 0 and the whole range.
void o_warp_slab(const float* f0, const float* f1, long f1_z0g, const float* u, const float* v,
                 const float* wv, size_t width, size_t height, size_t depth, long z0g, long zs, long ze,
                 float hx, float hy, float hz, float* out) {
#pragma omp parallel for schedule(static)
  for (long lz = zs; lz < ze; ++lz)
    for (size_t gy = 0; gy < height; ++gy)
      for (size_t gx = 0; gx < width; ++gx) {
        const long gz = z0g + lz;
        size_t c = IDX(gx, gy, lz, width, height);
        // registration_3d.ptx: rcp.rn(h) then ONE fma.rn per coordinate
        float x_f = std::fmaf(1.f / hx, u[c], (float)(unsigned)gx);
        float y_f = std::fmaf(1.f / hy, v[c], (float)(unsigned)gy);
        float z_f = std::fmaf(1.f / hz, wv[c], (float)(unsigned)gz);
        if ((x_f < 0.) || (x_f > width - 1) || (y_f < 0.) || (y_f > height - 1) || (z_f < 0.) ||
            (z_f > depth - 1) || std::isnan(x_f) || std::isnan(y_f) || std::isnan(z_f)) {
          out[c] = f0[c];
        } else {
          int x = (int)std::floor(x_f);
          int y = (int)std::floor(y_f);
          int z = (int)std::floor(z_f);
          float delta_x = x_f - (float)x;
          float delta_y = y_f - (float)y;
          float delta_z = z_f - (float)z;
          int x_1 = std::fmin((float)(width - 1), (float)size_t(x + 1));
          int y_1 = std::fmin((float)(height - 1), (float)size_t(y + 1));
          int z_1 = std::fmin((float)(depth - 1), (float)size_t(z + 1));
          // weights are plain products; the 4-term sums contract as
          // fma(w11,f11, fma(w01,f01, fma(w00,f00, w10*f10)))
          float w00 = (1.f - delta_x) * (1.f - delta_y);
          float w10 = (delta_x) * (1.f - delta_y);
          float w01 = (1.f - delta_x) * (delta_y);
          float w11 = (delta_x) * (delta_y);
          const long zl0 = z - f1_z0g, zl1 = z_1 - f1_z0g;  // planes of the frame-1 slab
          float value_0 = w10 * f1[IDX(x_1, y, zl0, width, height)];
          value_0 = std::fmaf(w00, f1[IDX(x, y, zl0, width, height)], value_0);
          value_0 = std::fmaf(w01, f1[IDX(x, y_1, zl0, width, height)], value_0);
          value_0 = std::fmaf(w11, f1[IDX(x_1, y_1, zl0, width, height)], value_0);
          float value_1 = w10 * f1[IDX(x_1, y, zl1, width, height)];
          value_1 = std::fmaf(w00, f1[IDX(x, y, zl1, width, height)], value_1);
          value_1 = std::fmaf(w01, f1[IDX(x, y_1, zl1, width, height)], value_1);
          value_1 = std::fmaf(w11, f1[IDX(x_1, y_1, zl1, width, height)], value_1);
          out[c] = std::fmaf(1.f - delta_z, value_0, delta_z * value_1);
        }
      }
}

void o_warp(const float* f0, const float* f1, const float* u, const float* v, const float* wv,
            size_t width, size_t height, size_t depth, float hx, float hy, float hz, float* out) {
  o_warp_slab(f0, f1, 0, u, v, wv, width, height, depth, 0, 0, (long)depth, hx, hy, hz, out);
}

// compute_phi_ksi_3d, solve_3d.cu:177-260 (neighbour mirroring :73-75,87-170).  Slab form: buffers are
// z-slabs (plane 0 = global z0g) of a level of global depth dg; local planes [zs, ze) are computed.
void o_phi_ksi_slab(const float* f0, const float* f1, const float* u, const float* v, const float* wv,
                    const float* du, const float* dv, const float* dw, size_t width, size_t height,
                    long z0g, long dg, long zs, long ze, float hx, float hy, float hz, float eq_smooth,
                    float eq_data, float* phi, float* ksi) {
  const long W = width, H = height;
  const Slab sl{z0g, dg, zs, ze};
#pragma omp parallel for schedule(static)
  for (long z = zs; z < ze; ++z)
    for (long y = 0; y < H; ++y)
      for (long x = 0; x < W; ++x) {
        size_t c = IDX(x, y, z, W, H);
        size_t xp = IDX(mirror(x + 1, W), y, z, W, H), xm = IDX(mirror(x - 1, W), y, z, W, H);
        size_t yp = IDX(x, mirror(y + 1, H), z, W, H), ym = IDX(x, mirror(y - 1, H), z, W, H);
        size_t zp = IDX(x, y, zn(sl, z, 1), W, H), zm = IDX(x, y, zn(sl, z, -1), W, H);

        float dux = (u[xp] - u[xm] + du[xp] - du[xm]) / (2.f * hx);
        float duy = (u[yp] - u[ym] + du[yp] - du[ym]) / (2.f * hy);
        float duz = (u[zp] - u[zm] + du[zp] - du[zm]) / (2.f * hz);
        float dvx = (v[xp] - v[xm] + dv[xp] - dv[xm]) / (2.f * hx);
        float dvy = (v[yp] - v[ym] + dv[yp] - dv[ym]) / (2.f * hy);
        float dvz = (v[zp] - v[zm] + dv[zp] - dv[zm]) / (2.f * hz);
        float dwx = (wv[xp] - wv[xm] + dw[xp] - dw[xm]) / (2.f * hx);
        float dwy = (wv[yp] - wv[ym] + dw[yp] - dw[ym]) / (2.f * hy);
        float dwz = (wv[zp] - wv[zm] + dw[zp] - dw[zm]) / (2.f * hz);

        // solve_3d.ptx (compute_phi_ksi_3d): mul(duy,duy) first, then one fma.rn per further term
        float acc = duy * duy;
        acc = std::fmaf(dux, dux, acc);
        acc = std::fmaf(duz, duz, acc);
        acc = std::fmaf(dvx, dvx, acc);
        acc = std::fmaf(dvy, dvy, acc);
        acc = std::fmaf(dvz, dvz, acc);
        acc = std::fmaf(dwx, dwx, acc);
        acc = std::fmaf(dwy, dwy, acc);
        acc = std::fmaf(dwz, dwz, acc);
        acc = std::fmaf(eq_smooth, eq_smooth, acc);
        float sq = std::sqrt(acc);
        phi[c] = 1.f / (sq + sq);

        float fx = (f0[xp] - f0[xm] + f1[xp] - f1[xm]) / (4.f * hx);
        float fy = (f0[yp] - f0[ym] + f1[yp] - f1[ym]) / (4.f * hy);
        float fz = (f0[zp] - f0[zm] + f1[zp] - f1[zm]) / (4.f * hz);
        float ft = f1[c] - f0[c];

        float J11 = fx * fx, J22 = fy * fy, J33 = fz * fz;
        float J12 = fx * fy, J13 = fx * fz, J23 = fy * fz;
        float J14 = fx * ft, J24 = fy * ft, J34 = fz * ft;

        float a = du[c], b = dv[c], cc = dw[c];
        // rows of the quadratic form, transcribed operation by operation from the PTX (which product
        // of a row stays a separately rounded mul is the compiler's choice and differs in row 3)
        float r1 = J14 + std::fmaf(J13, cc, std::fmaf(J11, a, J12 * b));
        float r2 = J24 + std::fmaf(J23, cc, std::fmaf(J12, a, J22 * b));
        float r3 = J34 + std::fmaf(J33, cc, std::fmaf(J23, b, J13 * a));  // NB: here J13*du is the rounded product
        float r4 = std::fmaf(ft, ft, std::fmaf(J34, cc, std::fmaf(J14, a, J24 * b)));
        float s = std::fmaf(cc, r3, std::fmaf(a, r1, b * r2)) + r4;
        s = s * ((s > 0) ? 1.f : 0.f);
        float sq2 = std::sqrt(std::fmaf(eq_data, eq_data, s));
        ksi[c] = 1.f / (sq2 + sq2);
      }
}

void o_phi_ksi(const float* f0, const float* f1, const float* u, const float* v, const float* wv,
               const float* du, const float* dv, const float* dw, size_t width, size_t height,
               size_t depth, float hx, float hy, float hz, float eq_smooth, float eq_data,
               float* phi, float* ksi) {
  o_phi_ksi_slab(f0, f1, u, v, wv, du, dv, dw, width, height, 0, (long)depth, 0, (long)depth, hx, hy, hz,
                 eq_smooth, eq_data, phi, ksi);
}

// solve_3d, solve_3d.cu:423-507: one Jacobi sweep (du,dv,dw) -> (tdu,tdv,tdw); slab form as above
void o_sweep_slab(const float* f0, const float* f1, const float* u, const float* v, const float* wv,
                  const float* du, const float* dv, const float* dw, const float* phi,
                  const float* ksi, size_t width, size_t height, long z0g, long dg, long zs, long ze,
                  float hx, float hy, float hz, float alpha, float* tdu, float* tdv, float* tdw) {
  const long W = width, H = height;
  const Slab sl{z0g, dg, zs, ze};
#pragma omp parallel for schedule(static)
  for (long z = zs; z < ze; ++z)
    for (long y = 0; y < H; ++y)
      for (long x = 0; x < W; ++x) {
        const long zglob = z0g + z;
        size_t c = IDX(x, y, z, W, H);
        size_t ixp = IDX(mirror(x + 1, W), y, z, W, H), ixm = IDX(mirror(x - 1, W), y, z, W, H);
        size_t iyp = IDX(x, mirror(y + 1, H), z, W, H), iym = IDX(x, mirror(y - 1, H), z, W, H);
        size_t izp = IDX(x, y, zn(sl, z, 1), W, H), izm = IDX(x, y, zn(sl, z, -1), W, H);

        float fx = (f0[ixp] - f0[ixm] + f1[ixp] - f1[ixm]) / (4.f * hx);
        float fy = (f0[iyp] - f0[iym] + f1[iyp] - f1[iym]) / (4.f * hy);
        float fz = (f0[izp] - f0[izm] + f1[izp] - f1[izm]) / (4.f * hz);
        float ft = f1[c] - f0[c];

        float J11 = fx * fx, J22 = fy * fy, J33 = fz * fz;
        float J12 = fx * fy, J13 = fx * fz, J23 = fy * fz;
        float J14 = fx * ft, J24 = fy * ft, J34 = fz * ft;

        float hx_2 = alpha / (hx * hx);
        float hy_2 = alpha / (hy * hy);
        float hz_2 = alpha / (hz * hz);

        float xp = (x < W - 1) * hx_2;
        float xm = (x > 0) * hx_2;
        float yp = (y < H - 1) * hy_2;
        float ym = (y > 0) * hy_2;
        float zp = (zglob < dg - 1) * hz_2;
        float zm = (zglob > 0) * hz_2;

        float phi_xp = (phi[ixp] + phi[c]) / 2.f;
        float phi_xm = (phi[ixm] + phi[c]) / 2.f;
        float phi_yp = (phi[iyp] + phi[c]) / 2.f;
        float phi_ym = (phi[iym] + phi[c]) / 2.f;
        float phi_zp = (phi[izp] + phi[c]) / 2.f;
        float phi_zm = (phi[izm] + phi[c]) / 2.f;

        // solve_3d.ptx (solve_3d): face weights are plain products shared by sumH and sumU/V/W;
        // sumH is a chain of plain adds; sumX = fma-chain seeded with the rounded x- product.
        float axp = xp * phi_xp, axm = xm * phi_xm;
        float ayp = yp * phi_yp, aym = ym * phi_ym;
        float azp = zp * phi_zp, azm = zm * phi_zm;
        float sumH = ((((axp + axm) + ayp) + aym) + azp) + azm;

        float uc = u[c], vc = v[c], wc = wv[c];
        float sumU = axm * ((u[ixm] + du[ixm]) - uc);
        sumU = std::fmaf(axp, (u[ixp] + du[ixp]) - uc, sumU);
        sumU = std::fmaf(ayp, (u[iyp] + du[iyp]) - uc, sumU);
        sumU = std::fmaf(aym, (u[iym] + du[iym]) - uc, sumU);
        sumU = std::fmaf(azp, (u[izp] + du[izp]) - uc, sumU);
        sumU = std::fmaf(azm, (u[izm] + du[izm]) - uc, sumU);
        float sumV = axm * ((v[ixm] + dv[ixm]) - vc);
        sumV = std::fmaf(axp, (v[ixp] + dv[ixp]) - vc, sumV);
        sumV = std::fmaf(ayp, (v[iyp] + dv[iyp]) - vc, sumV);
        sumV = std::fmaf(aym, (v[iym] + dv[iym]) - vc, sumV);
        sumV = std::fmaf(azp, (v[izp] + dv[izp]) - vc, sumV);
        sumV = std::fmaf(azm, (v[izm] + dv[izm]) - vc, sumV);
        float sumW = axm * ((wv[ixm] + dw[ixm]) - wc);
        sumW = std::fmaf(axp, (wv[ixp] + dw[ixp]) - wc, sumW);
        sumW = std::fmaf(ayp, (wv[iyp] + dw[iyp]) - wc, sumW);
        sumW = std::fmaf(aym, (wv[iym] + dw[iym]) - wc, sumW);
        sumW = std::fmaf(azp, (wv[izp] + dw[izp]) - wc, sumW);
        sumW = std::fmaf(azm, (wv[izm] + dw[izm]) - wc, sumW);

        // numerators: nvcc's PTX keeps J12*dv and J13*dw as separate mul.f32 + sub.f32, but those
        // carry no .rn and ptxas (offline 12.9 and the driver JIT alike; checked in the SASS and
        // against the reference kernels on a B200) contracts each pair:
        //   n = fma(-J13, dw, fma(-J12, dv, -J14));   num = fma(ksi, n, sum);   den = fma(Jii, ksi, sumH)
        float k = ksi[c];
        float n_du = std::fmaf(-J13, dw[c], std::fmaf(-J12, dv[c], -J14));
        float r_du = std::fmaf(k, n_du, sumU) / std::fmaf(J11, k, sumH);
        float n_dv = std::fmaf(-J23, dw[c], std::fmaf(-J12, r_du, -J24));
        float r_dv = std::fmaf(k, n_dv, sumV) / std::fmaf(J22, k, sumH);
        float n_dw = std::fmaf(-J23, r_dv, std::fmaf(-J13, r_du, -J34));
        float r_dw = std::fmaf(k, n_dw, sumW) / std::fmaf(J33, k, sumH);

        tdu[c] = r_du;
        tdv[c] = r_dv;
        tdw[c] = r_dw;
      }
}

void o_sweep(const float* f0, const float* f1, const float* u, const float* v, const float* wv,
             const float* du, const float* dv, const float* dw, const float* phi, const float* ksi,
             size_t width, size_t height, size_t depth, float hx, float hy, float hz, float alpha,
             float* tdu, float* tdv, float* tdw) {
  o_sweep_slab(f0, f1, u, v, wv, du, dv, dw, phi, ksi, width, height, 0, (long)depth, 0, (long)depth, hx, hy,
               hz, alpha, tdu, tdv, tdw);
}

// CudaOperationSolve::Execute, cuda_operation_solve.cpp:183-257.  On return du/dv/dw hold the
// final iterate (the reference swaps pointers; here we copy when the sweep count is odd).
// scratch: 5 volumes (phi, ksi, tdu, tdv, tdw).
void o_solve_level(const float* f0, const float* f1w, const float* u, const float* v,
                   const float* wv, float* du, float* dv, float* dw, float* scratch, size_t width,
                   size_t height, size_t depth, float hx, float hy, float hz, size_t outer,
                   size_t inner, float alpha, float eq_smooth, float eq_data) {
  const size_t n = width * height * depth;
  float* phi = scratch;
  float* ksi = scratch + n;
  float* a[3] = {du, dv, dw};
  float* b[3] = {scratch + 2 * n, scratch + 3 * n, scratch + 4 * n};
  std::memset(du, 0, n * sizeof(float));
  std::memset(dv, 0, n * sizeof(float));
  std::memset(dw, 0, n * sizeof(float));
  for (size_t i = 0; i < outer; ++i) {
    o_phi_ksi(f0, f1w, u, v, wv, a[0], a[1], a[2], width, height, depth, hx, hy, hz, eq_smooth,
              eq_data, phi, ksi);
    for (size_t j = 0; j < inner; ++j) {
      o_sweep(f0, f1w, u, v, wv, a[0], a[1], a[2], phi, ksi, width, height, depth, hx, hy, hz,
              alpha, b[0], b[1], b[2]);
      std::swap(a[0], b[0]);
      std::swap(a[1], b[1]);
      std::swap(a[2], b[2]);
    }
  }
  if (a[0] != du) {
    std::memcpy(du, a[0], n * sizeof(float));
    std::memcpy(dv, a[1], n * sizeof(float));
    std::memcpy(dw, a[2], n * sizeof(float));
  }
}

// add_3d.cu:37-40
void o_add(float* a, const float* b, size_t n) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) a[i] += b[i];
}

// median_3d.cu:282-297 with radius rules of cuda_operation_median.cpp:95-106 (radius = window edge
// length; 1 -> copy; even -> radius-1; supported 3,5,7).  Selection == element [len/2] of the sorted
// window.  Returns 0 on success, -1 for an unsupported radius (the reference prints and leaves the
// output untouched).
int o_median_slab(const float* in, float* out, size_t width, size_t height, long z0g, long dg, long zs,
                  long ze, size_t radius) {
  const Slab sl{z0g, dg, zs, ze};
  const size_t plane = width * height;
  if (radius == 1) {
    std::memcpy(out + zs * plane, in + zs * plane, (size_t)(ze - zs) * plane * sizeof(float));
    return 0;
  }
  if (radius % 2 == 0) radius -= 1;
  if (radius < 3 || radius > 7) return -1;
  const long r2 = (long)radius / 2;
  const long W = width, H = height;
  const size_t len = radius * radius * radius;
#pragma omp parallel for schedule(static)
  for (long z = zs; z < ze; ++z) {
    std::vector<float> buf(len);
    for (long y = 0; y < H; ++y)
      for (long x = 0; x < W; ++x) {
        size_t k = 0;
        for (long iz = -r2; iz <= r2; ++iz)
          for (long iy = -r2; iy <= r2; ++iy)
            for (long ix = -r2; ix <= r2; ++ix)
              buf[k++] = in[IDX(mirror(x + ix, W), mirror(y + iy, H), zn(sl, z, iz), W, H)];
        std::nth_element(buf.begin(), buf.begin() + len / 2, buf.end());
        out[IDX(x, y, z, W, H)] = buf[len / 2];
      }
  }
  return 0;
}

int o_median(const float* in, float* out, size_t width, size_t height, size_t depth,
             size_t radius) {
  return o_median_slab(in, out, width, height, 0, (long)depth, 0, (long)depth, radius);
}

// z pass of the resample between slabs: input slab (plane 0 = global in_z0g, global depth in_dg) ->
// output slab (out_z0g, out_dg), output local planes [zs, ze); same tap arithmetic in global indices.
void o_resample_z_slab(const float* in, size_t w, size_t h, long in_z0g, long in_dg, float* out,
                       long out_z0g, long out_dg, long zs, long ze) {
#pragma omp parallel for schedule(static)
  for (long z = zs; z < ze; ++z) {
    std::vector<float> frac((size_t)in_dg + 4);
    int li, cnt;
    float norm = resample_taps((size_t)in_dg, (size_t)out_dg, (unsigned)(out_z0g + z), &li, &cnt, frac.data(),
                               (int)in_dg + 4);
    for (size_t y = 0; y < h; ++y)
      for (size_t x = 0; x < w; ++x) {
        float value = 0.f;
        for (int j = 0; j < cnt; j++)
          value = std::fmaf(frac[j], in[IDX(x, y, (size_t)(li + j - in_z0g), w, h)], value);
        out[IDX(x, y, z, w, h)] = value * norm;
      }
  }
}

struct OracleParams {
  size_t warp_levels_count;
  float warp_scale_factor;
  size_t outer_iterations_count;
  size_t inner_iterations_count;
  float equation_alpha;
  float equation_smoothness;
  float equation_data;
  size_t median_radius;
  float gaussian_sigma;
};

typedef void (*o_level_cb)(int level, const size_t* dims, const float* u, const float* v,
                           const float* w, void* user);

// OpticalFlowE::ComputeFlow, optical_flow_e.cpp:132-601.
// frame_0/frame_1: W*H*D floats; flow_u/v/w: W*H*D floats (output).
// level_cb (optional) is called after each level's median with that level's flow.
int o_compute_flow(const float* frame_0, const float* frame_1, size_t W, size_t H, size_t D,
                   const OracleParams* p, float* flow_u, float* flow_v, float* flow_w,
                   o_level_cb level_cb, void* user) {
  const size_t N = W * H * D;
  std::vector<float> f0(frame_0, frame_0 + N), f1(frame_1, frame_1 + N);
  std::vector<float> t0(N), t1(N);

  size_t max_level = o_max_warp_level(W, H, D, p->warp_scale_factor);
  int level = (int)std::min(p->warp_levels_count, max_level) - 1;  // :180

  if (p->gaussian_sigma > 0.0) {  // :213-239
    o_gauss_blur(frame_0, f0.data(), t0.data(), W, H, D, p->gaussian_sigma);
    o_gauss_blur(frame_1, f1.data(), t0.data(), W, H, D, p->gaussian_sigma);
  }

  std::vector<float> f0l(N), f1l(N), f1w(N), u(N), v(N), w(N), du(N), dv(N), dw(N), scratch(5 * N);
  size_t prev[3] = {0, 0, 0};
  const size_t full[3] = {W, H, D};

  while (level >= 0) {  // :261
    size_t cur[3];
    float h[3];
    o_level_geometry(W, H, D, p->warp_scale_factor, level, cur, h);
    const size_t n = cur[0] * cur[1] * cur[2];

    const float* pf0;
    const float* pf1;
    if (level == 0) {  // :275-277
      pf0 = f0.data();
      pf1 = f1.data();
    } else {  // :279-299 (always from full resolution)
      o_resample(f0.data(), full, f0l.data(), cur, t0.data(), t1.data());
      o_resample(f1.data(), full, f1l.data(), cur, t0.data(), t1.data());
      pf0 = f0l.data();
      pf1 = f1l.data();
    }

    if (prev[0] == 0) {  // :304-310
      std::fill(u.begin(), u.begin() + n, 0.f);
      std::fill(v.begin(), v.begin() + n, 0.f);
      std::fill(w.begin(), w.begin() + n, 0.f);
    } else {  // :311-344 (values are NOT rescaled)
      o_resample(u.data(), prev, du.data(), cur, t0.data(), t1.data());
      o_resample(v.data(), prev, dv.data(), cur, t0.data(), t1.data());
      o_resample(w.data(), prev, dw.data(), cur, t0.data(), t1.data());
      std::swap(u, du);
      std::swap(v, dv);
      std::swap(w, dw);
    }

    // :348-369
    o_warp(pf0, pf1, u.data(), v.data(), w.data(), cur[0], cur[1], cur[2], h[0], h[1], h[2],
           f1w.data());

    // :372-417
    o_solve_level(pf0, f1w.data(), u.data(), v.data(), w.data(), du.data(), dv.data(), dw.data(),
                  scratch.data(), cur[0], cur[1], cur[2], h[0], h[1], h[2],
                  p->outer_iterations_count, p->inner_iterations_count, p->equation_alpha,
                  p->equation_smoothness, p->equation_data);

    // :420-438
    o_add(u.data(), du.data(), n);
    o_add(v.data(), dv.data(), n);
    o_add(w.data(), dw.data(), n);

    prev[0] = cur[0];
    prev[1] = cur[1];
    prev[2] = cur[2];
    --level;

    // :443-473
    int rc = o_median(u.data(), t0.data(), cur[0], cur[1], cur[2], p->median_radius);
    if (rc == 0) std::swap_ranges(t0.begin(), t0.begin() + n, u.begin());
    rc = o_median(v.data(), t0.data(), cur[0], cur[1], cur[2], p->median_radius);
    if (rc == 0) std::swap_ranges(t0.begin(), t0.begin() + n, v.begin());
    rc = o_median(w.data(), t0.data(), cur[0], cur[1], cur[2], p->median_radius);
    if (rc == 0) std::swap_ranges(t0.begin(), t0.begin() + n, w.begin());

    if (level_cb) level_cb(level + 1, cur, u.data(), v.data(), w.data(), user);
  }

  // :574-576
  std::memcpy(flow_u, u.data(), N * sizeof(float));
  std::memcpy(flow_v, v.data(), N * sizeof(float));
  std::memcpy(flow_w, w.data(), N * sizeof(float));
  return 0;
}

}  // extern "C"
