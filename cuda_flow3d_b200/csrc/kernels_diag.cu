// kernels_diag.cu -- convergence diagnostic of the Jacobi relaxation.
//
// The reference runs a fixed number of sweeps and never looks at a residual.  This
// kernel measures the size of the last update, sum |a - b|^2 and max |a - b| over the three flow
// increments of two successive iterates, so that a caller can watch (or, opt-in, act on) convergence.
// It is NOT on the parity path: nothing the solve computes depends on it unless a tolerance is set.
//
// Reduction: per-thread accumulation in double, warp shuffles, one shared-memory step per block, a
// per-block partial in global memory, and the last block to finish (ticket counter) folds the
// partials in index order -- the result is deterministic for a given launch geometry.
#include "common.cuh"

namespace f3d {

struct NormPartial {
  double sum;
  float mx;
  float pad;
};

static constexpr int kNormMaxBlocks = 4096;

__global__ void __launch_bounds__(256) update_norm_kernel(const float* __restrict__ a0, const float* __restrict__ a1,
                                                          const float* __restrict__ a2, const float* __restrict__ b0,
                                                          const float* __restrict__ b1, const float* __restrict__ b2,
                                                          Dims g, int zs, int ze, NormPartial* __restrict__ part,
                                                          unsigned* __restrict__ ticket, double* __restrict__ out) {
  double sum = 0.0;
  float mx = 0.f;
  const long long rows = (long long)g.h * (ze - zs);
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int z = zs + (int)(r / g.h), y = (int)(r % g.h);
    const long long o = (long long)z * g.ps + (long long)y * g.ld;
    for (int x = threadIdx.x; x < g.w; x += blockDim.x) {
      const float d0 = __fsub_rn(__ldg(a0 + o + x), __ldg(b0 + o + x));
      const float d1 = __fsub_rn(__ldg(a1 + o + x), __ldg(b1 + o + x));
      const float d2 = __fsub_rn(__ldg(a2 + o + x), __ldg(b2 + o + x));
      sum += (double)d0 * (double)d0 + (double)d1 * (double)d1 + (double)d2 * (double)d2;
      const float m0 = fabsf(d0), m1 = fabsf(d1), m2 = fabsf(d2);
      mx = (m0 > mx || m0 != m0) ? m0 : mx;  // NaN sticks (fmaxf would drop it)
      mx = (m1 > mx || m1 != m1) ? m1 : mx;
      mx = (m2 > mx || m2 != m2) ? m2 : mx;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    const float t = __shfl_down_sync(0xffffffffu, mx, o);
    mx = (t > mx || t != t) ? t : mx;
  }
  __shared__ double s_sum[8];
  __shared__ float s_mx[8];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = sum; s_mx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float m = 0.f;
    for (int i = 0; i < 8; ++i) {
      t += s_sum[i];
      m = (s_mx[i] > m || s_mx[i] != s_mx[i]) ? s_mx[i] : m;
    }
    part[blockIdx.x].sum = t;
    part[blockIdx.x].mx = m;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    float m = 0.f;
    for (unsigned i = 0; i < gridDim.x; ++i) {
      const volatile NormPartial* p = part + i;
      t += p->sum;
      const float v = p->mx;
      m = (v > m || v != v) ? v : m;
    }
    out[0] = t;
    out[1] = (double)m;
  }
}

size_t update_norm_workspace_bytes() { return sizeof(NormPartial) * kNormMaxBlocks + 256; }

int launch_update_norm(const float* a0, const float* a1, const float* a2, const float* b0, const float* b1,
                       const float* b2, Dims g, ZRange zr, double* out_dev, void* workspace, cudaStream_t st) {
  if (zr.end <= zr.begin) return FLOW3D_ERR_INVALID_ARG;
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  NormPartial* part = reinterpret_cast<NormPartial*>(reinterpret_cast<char*>(workspace) + 256);
  const long long rows = (long long)g.h * (zr.end - zr.begin);
  long long blocks = (long long)sm_count() * 8;
  if (blocks > rows) blocks = rows;
  if (blocks > kNormMaxBlocks) blocks = kNormMaxBlocks;
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) { note_cuda_error(e, "cudaMemsetAsync"); return FLOW3D_ERR_CUDA; }
  update_norm_kernel<<<(unsigned)blocks, 256, 0, st>>>(a0, a1, a2, b0, b1, b2, g, zr.begin, zr.end, part, ticket,
                                                       out_dev);
  count_launch();
  return check_launch("update_norm_kernel");
}

// ------------------------------------------------------------------------------------------------
// self-test of the branch-free division (common.cuh: div_fast) against __fdiv_rn, on the device (the
// sequence starts from MUFU.RCP, which has no host equivalent).  mode 0: random sign / exponent in
// [2^-60, 2^60) / mantissa for both operands; mode 1: EVERY divisor mantissa (2^23) against a handful
// of dividends per launch slice.  out[0] = mismatching quotients among accepted pairs, out[1] = pairs
// the range flag rejected, out[2] = pairs tested.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float make_operand(unsigned long long r) {
  const unsigned mant = (unsigned)(r & 0x7fffffu);
  const unsigned ex = 67u + (unsigned)((r >> 23) % 120u);  // 2^-60 .. 2^59
  const unsigned sign = (unsigned)((r >> 40) & 1u) << 31;
  return __uint_as_float(sign | (ex << 23) | mant);
}
__global__ void __launch_bounds__(256) fast_div_selftest_kernel(unsigned long long n, unsigned long long seed, int mode,
                                                                unsigned long long* out) {
  unsigned long long bad = 0, rejected = 0, tested = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    if (mode >= 2) {  // unary: every float bit pattern i (mode 2: sqrt_fast, mode 3: rcp_fast)
      const float x = __uint_as_float((unsigned)i);
      bool ok = true;
      const float q = mode == 2 ? sqrt_fast(x, ok) : rcp_fast(x, ok);
      const float ref = mode == 2 ? __fsqrt_rn(x) : __frcp_rn(x);
      ++tested;
      if (!ok) ++rejected;
      else if (__float_as_uint(q) != __float_as_uint(ref)) ++bad;
      continue;
    }
    float a, b;
    if (mode == 0) {
      a = make_operand(mix64(seed + 2 * i));
      b = make_operand(mix64(seed + 2 * i + 1));
    } else {
      a = make_operand(mix64(seed + (i >> 23)));
      b = __uint_as_float((127u << 23) | (unsigned)(i & 0x7fffffu));
      if ((i >> 23) & 1) b = -b;
    }
    bool ok = true;
    const float q = div_fast(a, b, ok);
    const float ref = __fdiv_rn(a, b);
    ++tested;
    if (!ok) ++rejected;
    else if (__float_as_uint(q) != __float_as_uint(ref)) ++bad;
  }
  for (int o = 16; o > 0; o >>= 1) {
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
    rejected += __shfl_xor_sync(0xffffffffu, rejected, o);
    tested += __shfl_xor_sync(0xffffffffu, tested, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 0, bad);
    atomicAdd(out + 1, rejected);
    atomicAdd(out + 2, tested);
  }
}

int launch_fast_div_selftest(unsigned long long n, unsigned long long seed, int mode, unsigned long long out_host[3]) {
  unsigned long long* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, 3 * sizeof(unsigned long long));
  if (e != cudaSuccess) { note_cuda_error(e, "cudaMalloc"); return FLOW3D_ERR_CUDA; }
  cudaMemset(dev, 0, 3 * sizeof(unsigned long long));
  fast_div_selftest_kernel<<<sm_count() * 8, 256>>>(n, seed, mode, dev);
  e = cudaMemcpy(out_host, dev, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) { note_cuda_error(e, "fast_div_selftest"); return FLOW3D_ERR_CUDA; }
  return FLOW3D_OK;
}

}  // namespace f3d
