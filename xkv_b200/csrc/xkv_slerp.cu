// xkv_b200 — SLERP / MiniCache branch (layer_merge_impl == "slerp", group size 2).
//
// Reference: slerp_merge_rows_batch (fake_layer_merge_dynamic_cache.py:32-90) and fake_minicache_merge
// (:93-100), called at :183-197 on the two layers' K (and V) reshaped to (bs*H*S, D) rows.
// Row-wise: norms, angle Omega between the rows, SLERP direction E; rows whose angle exceeds
// d_min + (d_max - d_min) * gamma ("diverge_mask") are replaced by E scaled back to each layer's own norm,
// all other rows keep their original values (that is what the reference code does, :96-99), rows with
// Omega < 1e-7 use the linear interpolation (:74, :82-87).  Two HBM-bound passes: one to get every row's
// angle and the global min / max (atomics on the non-negative float bit patterns), one to apply.
// Arithmetic is fp32 here (the reference runs these torch ops in the cache dtype, bf16).
#include <cfloat>

#include "xkv_common.cuh"
#include "xkv_host.h"

namespace xkv {

__device__ __forceinline__ float block_row_dot(const __nv_bfloat16* a, const __nv_bfloat16* b, int d, int lane) {
  float acc = 0.f;
  for (int i = lane * 2; i < d; i += 64) {
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(a + i);
    const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(b + i);
    acc = fmaf(__bfloat162float(x.x), __bfloat162float(y.x), fmaf(__bfloat162float(x.y), __bfloat162float(y.y), acc));
  }
  return warp_sum(acc);
}

// one warp per row: omega[row], n1[row], n2[row]; minmax[0] = min omega, minmax[1] = max omega (as uint bits)
__global__ void __launch_bounds__(256) slerp_stats_kernel(const __nv_bfloat16* __restrict__ x1,
                                                          const __nv_bfloat16* __restrict__ x2, long long rows, int d,
                                                          long long ld, float* __restrict__ stats,
                                                          unsigned int* __restrict__ minmax) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  unsigned int lo = 0x7F800000u, hi = 0u;  // +inf, 0
  for (long long r = warp0; r < rows; r += nwarps) {
    const __nv_bfloat16* a = x1 + r * ld;
    const __nv_bfloat16* b = x2 + r * ld;
    const float n1 = sqrtf(block_row_dot(a, a, d, lane));
    const float n2 = sqrtf(block_row_dot(b, b, d, lane));
    float c = block_row_dot(a, b, d, lane) / (n1 * n2);
    c = fminf(1.f, fmaxf(-1.f, c));
    const float om = acosf(c);
    if (lane == 0) {
      stats[3 * r] = om;
      stats[3 * r + 1] = n1;
      stats[3 * r + 2] = n2;
      const unsigned int bits = __float_as_uint(om);
      lo = min(lo, bits);
      hi = max(hi, bits);
    }
  }
  if (lane == 0 && hi >= lo) {
    atomicMin(&minmax[0], lo);
    atomicMax(&minmax[1], hi);
  }
}

__global__ void __launch_bounds__(256) slerp_apply_kernel(const __nv_bfloat16* __restrict__ x1,
                                                          const __nv_bfloat16* __restrict__ x2, long long rows, int d,
                                                          long long ld, const float* __restrict__ stats,
                                                          const unsigned int* __restrict__ minmax, float t, float gamma,
                                                          __nv_bfloat16* __restrict__ e1, __nv_bfloat16* __restrict__ e2,
                                                          long long ldo) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float dmin = __uint_as_float(minmax[0]), dmax = __uint_as_float(minmax[1]);
  const float threshold = dmin + (dmax - dmin) * gamma;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float om = stats[3 * r], n1 = stats[3 * r + 1], n2 = stats[3 * r + 2];
    const bool diverge = om > threshold;
    const bool parallel = om < 1e-7f;
    const float so = sinf(om);
    const float alpha = sinf((1.f - t) * om) / so, beta = sinf(t * om) / so;
    const __nv_bfloat16* a = x1 + r * ld;
    const __nv_bfloat16* b = x2 + r * ld;
    for (int i = lane * 2; i < d; i += 64) {
      const __nv_bfloat162 xa = *reinterpret_cast<const __nv_bfloat162*>(a + i);
      const __nv_bfloat162 xb = *reinterpret_cast<const __nv_bfloat162*>(b + i);
      __nv_bfloat162 o1 = xa, o2 = xb;
      if (diverge) {
        float ev[2];
        const float av[2] = {__bfloat162float(xa.x), __bfloat162float(xa.y)};
        const float bv[2] = {__bfloat162float(xb.x), __bfloat162float(xb.y)};
#pragma unroll
        for (int u = 0; u < 2; ++u)
          ev[u] = parallel ? (1.f - t) * av[u] + t * bv[u] : alpha * (av[u] / n1) + beta * (bv[u] / n2);
        o1 = __floats2bfloat162_rn(ev[0] * n1, ev[1] * n1);
        o2 = __floats2bfloat162_rn(ev[0] * n2, ev[1] * n2);
      }
      *reinterpret_cast<__nv_bfloat162*>(e1 + r * ldo + i) = o1;
      *reinterpret_cast<__nv_bfloat162*>(e2 + r * ldo + i) = o2;
    }
  }
}

__global__ void slerp_init_kernel(unsigned int* minmax) {
  minmax[0] = 0x7F800000u;
  minmax[1] = 0u;
}

}  // namespace xkv

using namespace xkv;

extern "C" size_t xkv_slerp_workspace_bytes(int64_t rows) { return static_cast<size_t>(rows) * 3 * sizeof(float) + 256; }

extern "C" int xkv_slerp_merge(const void* x1, const void* x2, int64_t rows, int d, int64_t ld, float t, float gamma,
                               void* e1, void* e2, int64_t ld_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  XKV_REQUIRE(x1 && x2 && e1 && e2 && workspace, "slerp: null argument");
  XKV_REQUIRE(rows >= 1 && d >= 2 && d % 2 == 0 && ld % 2 == 0 && ld_out % 2 == 0, "slerp: bad sizes (d, ld must be even)");
  XKV_REQUIRE(workspace_bytes >= xkv_slerp_workspace_bytes(rows), "slerp: workspace too small");
  cudaStream_t st = as_stream(stream);
  unsigned int* minmax = static_cast<unsigned int*>(workspace);
  float* stats = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  long long grid = (rows + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  slerp_init_kernel<<<1, 1, 0, st>>>(minmax);
  XKV_LAUNCHED();
  slerp_stats_kernel<<<static_cast<int>(grid), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x1),
                                                             static_cast<const __nv_bfloat16*>(x2), rows, d, ld, stats,
                                                             minmax);
  XKV_LAUNCHED();
  slerp_apply_kernel<<<static_cast<int>(grid), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x1), static_cast<const __nv_bfloat16*>(x2), rows, d, ld, stats, minmax, t, gamma,
      static_cast<__nv_bfloat16*>(e1), static_cast<__nv_bfloat16*>(e2), ld_out);
  XKV_LAUNCHED();
  return 0;
}
