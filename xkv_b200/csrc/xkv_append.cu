// xkv_b200 — append (north-star step 4): project new token rows onto a group's right factor.
//
//   a_new (T x r) = x_new (T x n) * V (n x r)
//
// x_new is the group's token-major row(s) (gathered with xkv_pack_group from the layers' new pre-RoPE keys
// or values), V the stored right factor.  For the few tokens of a decode step this is a GEMV: HBM-bound on
// reading V once (n*r*2 bytes: 4.2 MB for K, 6.3 MB for V at config 2).  Two deterministic stages: per
// 128-row chunk partial sums (coalesced bf16x2 reads of V rows), then a chunk reduction that writes bf16.
// The reference never compresses decode tokens (mode != 'prefill' skips merging, cache:131), so callers
// enable this explicitly (FakeLayerMergingCache(..., compress_decode_tokens=True)).
#include "xkv_common.cuh"
#include "xkv_host.h"

namespace xkv {

constexpr int AP_ROWS = 128;   // rows of V per chunk
constexpr int AP_TMAX = 8;     // new tokens handled per launch

__global__ void __launch_bounds__(128) append_partial_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int T,
                                                             const __nv_bfloat16* __restrict__ V, long long ldv, int n,
                                                             int r, float* __restrict__ partial) {
  __shared__ float xs[AP_TMAX][AP_ROWS];
  const int chunk = blockIdx.y;
  const int i0 = chunk * AP_ROWS;
  const int rows = min(AP_ROWS, n - i0);
  for (int e = threadIdx.x; e < T * AP_ROWS; e += blockDim.x) {
    const int t = e / AP_ROWS, i = e - t * AP_ROWS;
    xs[t][i] = i < rows ? __bfloat162float(x[t * ldx + i0 + i]) : 0.f;
  }
  __syncthreads();
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= r) return;
  float acc0[AP_TMAX], acc1[AP_TMAX];
#pragma unroll
  for (int t = 0; t < AP_TMAX; ++t) acc0[t] = acc1[t] = 0.f;
  const __nv_bfloat16* vp = V + static_cast<long long>(i0) * ldv + j;
#pragma unroll 4
  for (int i = 0; i < rows; ++i) {
    const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vp + static_cast<long long>(i) * ldv);
    const float v0 = __bfloat162float(v2.x), v1 = __bfloat162float(v2.y);
#pragma unroll
    for (int t = 0; t < AP_TMAX; ++t) {
      if (t < T) {
        acc0[t] = fmaf(xs[t][i], v0, acc0[t]);
        acc1[t] = fmaf(xs[t][i], v1, acc1[t]);
      }
    }
  }
  for (int t = 0; t < T; ++t) {
    float* p = partial + (static_cast<long long>(chunk) * T + t) * r + j;
    p[0] = acc0[t];
    p[1] = acc1[t];
  }
}

__global__ void __launch_bounds__(256) append_reduce_kernel(const float* __restrict__ partial, int chunks, int T, int r,
                                                            __nv_bfloat16* __restrict__ out, long long ldo) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T * r) return;
  const int t = e / r, j = e - t * r;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += partial[(static_cast<long long>(c) * T + t) * r + j];
  out[t * ldo + j] = __float2bfloat16_rn(acc);
}

}  // namespace xkv

using namespace xkv;

extern "C" size_t xkv_append_workspace_bytes(int T, int n, int r) {
  const size_t chunks = (static_cast<size_t>(n) + AP_ROWS - 1) / AP_ROWS;
  return chunks * static_cast<size_t>(T < AP_TMAX ? T : AP_TMAX) * r * sizeof(float) + 256;
}

extern "C" int xkv_append_project(const void* x_new, int64_t ldx, int T, const void* V, int64_t ldv, int n, int r,
                                  void* a_out, int64_t lda, void* workspace, size_t workspace_bytes, void* stream) {
  XKV_REQUIRE(x_new && V && a_out && workspace, "append: null argument");
  XKV_REQUIRE(T >= 1 && n >= 1 && r >= 2 && r % 2 == 0 && ldv % 2 == 0, "append: bad sizes (r and ldv must be even)");
  XKV_REQUIRE(workspace_bytes >= xkv_append_workspace_bytes(T, n, r), "append: workspace too small");
  const int chunks = (n + AP_ROWS - 1) / AP_ROWS;
  cudaStream_t st = as_stream(stream);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_new);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a_out);
  for (int t0 = 0; t0 < T; t0 += AP_TMAX) {
    const int tt = T - t0 < AP_TMAX ? T - t0 : AP_TMAX;
    dim3 grid((r / 2 + 127) / 128, chunks);
    append_partial_kernel<<<grid, 128, 0, st>>>(x + t0 * ldx, ldx, tt, static_cast<const __nv_bfloat16*>(V), ldv, n, r,
                                                static_cast<float*>(workspace));
    XKV_LAUNCHED();
    append_reduce_kernel<<<(tt * r + 255) / 256, 256, 0, st>>>(static_cast<const float*>(workspace), chunks, tt, r,
                                                               out + t0 * lda, lda);
    XKV_LAUNCHED();
  }
  return 0;
}
