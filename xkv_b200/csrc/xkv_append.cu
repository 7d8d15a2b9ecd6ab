// xkv_b200 — append (north-star step 4): project new token rows onto a group's right factor.
//
//   a_new (T x r) = x_new (T x n) * V (n x r)
//
// x_new is the group's token-major row(s) (gathered with xkv_pack_group from the layers' new pre-RoPE keys
// or values), V the stored right factor.  For the few tokens of a decode step this is a GEMV: HBM-bound on
// reading V once (n*r*2 bytes: 4.2 MB for K, 6.3 MB for V at config 2), and at a few microseconds of HBM time it is
// launch-bound unless everything happens in ONE launch.  So: up to XKV_APPEND_MAX_PROBLEMS projections (a group's K and
// V factor) share a launch; a CTA owns 64 rows x 256 columns of V, its 8 warps take 8 rows each with all their 16-byte
// loads issued up front (128 B in flight per thread, ~5 MB across the grid); the per-CTA partial sums go to the
// workspace and the LAST CTA to finish a column block (a counter per block) adds them in a fixed order and writes
// bf16 — deterministic, no second launch.
// The reference never compresses decode tokens (mode != 'prefill' skips merging, cache:131), so callers
// enable this explicitly (FakeLayerMergingCache(..., compress_decode_tokens=True)).
#include "xkv_common.cuh"
#include "xkv_host.h"

namespace xkv {

constexpr int AP_ROWS = 128;    // rows of V per CTA
constexpr int AP_COLS = 256;    // columns of V per CTA (32 lanes x 8)
constexpr int AP_TMAX = 8;      // new tokens handled per launch
constexpr int AP_THREADS = 256;

struct AppendProb {
  const __nv_bfloat16* x;
  const __nv_bfloat16* V;
  __nv_bfloat16* out;
  float* partial;          // [chunks][T][r]
  unsigned int* counters;  // one per column block, zero on entry and on exit
  long long ldx, ldv, ldo;
  int n, r, chunks, col_blocks;
};
struct AppendParams {
  AppendProb p[XKV_APPEND_MAX_PROBLEMS];
  int T;
};

// TT: compile-time bound on the tokens of the launch (1, 2, 4 or 8: accumulator registers); VEC: every problem's columns
// can be read as aligned 16-byte vectors (r % 8 == 0, ldv % 8 == 0, 16-byte aligned V) -- then the 16 row loads of a
// thread carry no branch and are all in flight together, which is what the kernel's speed rests on.
template <int TT, bool VEC>
__global__ void __launch_bounds__(AP_THREADS) append_kernel(const __grid_constant__ AppendParams P) {
  extern __shared__ float ap_smem[];   // xs[T][AP_ROWS] | part[8 warps][T][AP_COLS]
  const AppendProb& pr = P.p[blockIdx.z];
  if (static_cast<int>(blockIdx.x) >= pr.col_blocks || static_cast<int>(blockIdx.y) >= pr.chunks) return;
  const int T = P.T;
  float* xs = ap_smem;
  float* part = ap_smem + T * AP_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.y * AP_ROWS;
  const int rows = min(AP_ROWS, pr.n - i0);
  const int j0 = blockIdx.x * AP_COLS + lane * 8;     // this thread's 8 columns
  // ---- V rows first (the long-latency loads), then the x values ----
  // Rows past the end of V (ragged last chunk) and columns past r (ragged last column block) read a clamped, valid
  // address: their x value is zero / their result is never stored, so the loads need no predicate.
  uint4 v[AP_ROWS / 8];
  if (VEC) {
    const int jc = min(j0, pr.r - 8);
#pragma unroll
    for (int k = 0; k < AP_ROWS / 8; ++k) {
      const int i = min(warp + 8 * k, rows - 1);
      v[k] = __ldg(reinterpret_cast<const uint4*>(pr.V + static_cast<long long>(i0 + i) * pr.ldv + jc));
    }
  } else {
#pragma unroll
    for (int k = 0; k < AP_ROWS / 8; ++k) {
      const int i = min(warp + 8 * k, rows - 1);
      const __nv_bfloat16* src = pr.V + static_cast<long long>(i0 + i) * pr.ldv;
      __nv_bfloat16 tmp[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) tmp[c] = src[min(j0 + c, pr.r - 1)];
      v[k] = *reinterpret_cast<uint4*>(tmp);
    }
  }
  for (int e = threadIdx.x; e < T * AP_ROWS; e += AP_THREADS) {
    const int t = e / AP_ROWS, i = e - t * AP_ROWS;
    xs[e] = i < rows ? __bfloat162float(pr.x[t * pr.ldx + i0 + i]) : 0.f;
  }
  __syncthreads();
  // tokens in groups of at most 4 over the SAME loaded rows: 32 accumulator registers instead of 64 at T = 8
  constexpr int TG = TT < 4 ? TT : 4;
#pragma unroll
  for (int tg = 0; tg < TT; tg += TG) {
    float acc[TG][8];
#pragma unroll
    for (int t = 0; t < TG; ++t)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
#pragma unroll
    for (int k = 0; k < AP_ROWS / 8; ++k) {
      const int i = warp + 8 * k;
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[k]);
      float f[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        f[2 * c] = __bfloat162float(h[c].x);
        f[2 * c + 1] = __bfloat162float(h[c].y);
      }
#pragma unroll
      for (int t = 0; t < TG; ++t) {
        if (tg + t < T) {
          const float xv = xs[(tg + t) * AP_ROWS + i];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[t][c] = fmaf(xv, f[c], acc[t][c]);
        }
      }
    }
    // ---- the 8 warps' partial sums meet in shared memory; thread j adds them for column j (fixed order) ----
#pragma unroll
    for (int t = 0; t < TG; ++t) {
      if (tg + t < T) {
        float* dst = part + (warp * T + tg + t) * AP_COLS + lane * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
      }
    }
  }
  __syncthreads();
  const int col = blockIdx.x * AP_COLS + threadIdx.x;
  for (int t = 0; t < T; ++t) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[(w * T + t) * AP_COLS + threadIdx.x];
    if (col < pr.r) pr.partial[(static_cast<long long>(blockIdx.y) * T + t) * pr.r + col] = s;
  }
  // ---- last CTA of this column block: reduce the chunks' partial sums, write bf16, re-arm the counter ----
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int seen = atomicAdd(&pr.counters[blockIdx.x], 1u);
    is_last = seen == static_cast<unsigned int>(pr.chunks - 1);
    if (is_last) pr.counters[blockIdx.x] = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // warp w adds the chunks w, w + 8, ... for the lane's 8 columns (all loads of a token in flight together), shared
  // memory adds the eight warps: a fixed order whatever CTA ends up last
  const int cbase = blockIdx.x * AP_COLS + lane * 8;
  const bool cvec = cbase + 8 <= pr.r && (pr.r & 3) == 0;
  // all tokens' partial sums are fetched before any is used: one L2 round trip for the whole tail, not one per token
  float s8[TT][8];
  const long long step = static_cast<long long>(T) * pr.r;
#pragma unroll
  for (int t = 0; t < TT; ++t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s8[t][k] = 0.f;
    if (t < T) {
      const float* src = pr.partial + static_cast<long long>(t) * pr.r + cbase;
#pragma unroll 4
      for (int c = warp; c < pr.chunks; c += 8) {
        if (cvec) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(src + c * step));
          const float4 b = __ldcg(reinterpret_cast<const float4*>(src + c * step + 4));
          s8[t][0] += a.x, s8[t][1] += a.y, s8[t][2] += a.z, s8[t][3] += a.w;
          s8[t][4] += b.x, s8[t][5] += b.y, s8[t][6] += b.z, s8[t][7] += b.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (cbase + k < pr.r) s8[t][k] += __ldcg(src + c * step + k);
        }
      }
    }
  }
  __syncthreads();   // `part` is free again: everybody has read the tile sums
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    if (t < T) {
      float* dst = part + (warp * T + t) * AP_COLS + lane * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(s8[t][0], s8[t][1], s8[t][2], s8[t][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(s8[t][4], s8[t][5], s8[t][6], s8[t][7]);
    }
  }
  __syncthreads();
  if (col < pr.r) {
    for (int t = 0; t < T; ++t) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += part[(w * T + t) * AP_COLS + threadIdx.x];
      pr.out[t * pr.ldo + col] = __float2bfloat16_rn(sum);
    }
  }
}

static inline size_t ap_align(size_t x) { return (x + 255) / 256 * 256; }
constexpr size_t AP_COUNTER_BYTES = 4096;   // counters of every problem of a call, at the start of the workspace
static inline size_t ap_problem_bytes(int T, int n, int r) {
  const size_t chunks = (static_cast<size_t>(n) + AP_ROWS - 1) / AP_ROWS;
  const size_t tt = static_cast<size_t>(T < AP_TMAX ? T : AP_TMAX);
  return ap_align(chunks * tt * r * sizeof(float));
}

}  // namespace xkv

using namespace xkv;

extern "C" size_t xkv_append_workspace_bytes(int T, int n, int r) { return ap_problem_bytes(T, n, r) + AP_COUNTER_BYTES; }

extern "C" int xkv_append_project_batch(const xkv_append_problem* problems, int count, int T, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  XKV_REQUIRE(problems && workspace && count >= 1 && count <= XKV_APPEND_MAX_PROBLEMS, "append: 1..%d problems per call",
              XKV_APPEND_MAX_PROBLEMS);
  XKV_REQUIRE(T >= 1, "append: no tokens");
  XKV_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "append: workspace must be 256-byte aligned");
  size_t need = AP_COUNTER_BYTES;
  size_t total_blocks = 0;
  for (int i = 0; i < count; ++i) {
    const xkv_append_problem& q = problems[i];
    XKV_REQUIRE(q.x_new && q.V && q.a_out, "append: null argument");
    XKV_REQUIRE(q.n >= 1 && q.r >= 1, "append: bad sizes");
    need += ap_problem_bytes(T, q.n, q.r);
    total_blocks += (static_cast<size_t>(q.r) + AP_COLS - 1) / AP_COLS;
  }
  XKV_REQUIRE(total_blocks * sizeof(unsigned int) <= AP_COUNTER_BYTES, "append: ranks too wide for one call");
  XKV_REQUIRE(workspace_bytes >= need, "append: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  bool vec = true;
  for (int i = 0; i < count; ++i)
    vec = vec && problems[i].r % 8 == 0 && problems[i].ldv % 8 == 0 && (reinterpret_cast<uintptr_t>(problems[i].V) & 15) == 0;
  for (int t0 = 0; t0 < T; t0 += AP_TMAX) {
    const int tt = T - t0 < AP_TMAX ? T - t0 : AP_TMAX;
    AppendParams P;
    std::memset(&P, 0, sizeof(P));
    P.T = tt;
    // the kernel re-arms its counters, but a caller may hand over a fresh workspace every time: one small memset per launch
    unsigned int* counters = static_cast<unsigned int*>(workspace);
    XKV_CHECK_CUDA(cudaMemsetAsync(counters, 0, total_blocks * sizeof(unsigned int), st));
    char* w = static_cast<char*>(workspace) + AP_COUNTER_BYTES;
    int gx = 1, gy = 1;
    for (int i = 0; i < count; ++i) {
      const xkv_append_problem& q = problems[i];
      AppendProb& a = P.p[i];
      a.x = static_cast<const __nv_bfloat16*>(q.x_new) + static_cast<long long>(t0) * q.ldx;
      a.V = static_cast<const __nv_bfloat16*>(q.V);
      a.out = static_cast<__nv_bfloat16*>(q.a_out) + static_cast<long long>(t0) * q.lda;
      a.ldx = q.ldx, a.ldv = q.ldv, a.ldo = q.lda;
      a.n = q.n, a.r = q.r;
      a.chunks = (q.n + AP_ROWS - 1) / AP_ROWS;
      a.col_blocks = (q.r + AP_COLS - 1) / AP_COLS;
      a.partial = reinterpret_cast<float*>(w);
      w += ap_align(static_cast<size_t>(a.chunks) * tt * q.r * sizeof(float));
      a.counters = counters;
      counters += a.col_blocks;
      gx = a.col_blocks > gx ? a.col_blocks : gx;
      gy = a.chunks > gy ? a.chunks : gy;
    }
    const size_t smem = static_cast<size_t>(tt * AP_ROWS + 8 * tt * AP_COLS) * sizeof(float);
    const dim3 grid(gx, gy, count);
    auto launch = [&](auto kern) -> int {
      XKV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>((AP_TMAX * AP_ROWS + 8 * AP_TMAX * AP_COLS) * sizeof(float))));
      kern<<<grid, AP_THREADS, smem, st>>>(P);
      XKV_LAUNCHED();
      return 0;
    };
    int rc;
    if (vec)
      rc = tt <= 1 ? launch(append_kernel<1, true>) : tt <= 2 ? launch(append_kernel<2, true>)
           : tt <= 4 ? launch(append_kernel<4, true>) : launch(append_kernel<8, true>);
    else
      rc = tt <= 1 ? launch(append_kernel<1, false>) : tt <= 2 ? launch(append_kernel<2, false>)
           : tt <= 4 ? launch(append_kernel<4, false>) : launch(append_kernel<8, false>);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int xkv_append_project(const void* x_new, int64_t ldx, int T, const void* V, int64_t ldv, int n, int r,
                                  void* a_out, int64_t lda, void* workspace, size_t workspace_bytes, void* stream) {
  xkv_append_problem q;
  q.x_new = x_new, q.ldx = ldx, q.V = V, q.ldv = ldv, q.n = n, q.r = r, q.a_out = a_out, q.lda = lda;
  return xkv_append_project_batch(&q, 1, T, workspace, workspace_bytes, stream);
}
