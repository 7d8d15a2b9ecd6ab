// xkv_b200 — blocked Cholesky (S + shift*I) = L L^T with explicit inverse Linv = L^{-1}, ONE launch:
// a thread-block cluster per matrix, cluster barriers instead of kernel boundaries.
//
// This is the serial heart of CholeskyQR (the orthogonalisation after every power step of the
// factorisation that replaces torch.linalg.svd, fake_layer_merge_dynamic_cache.py:20).  The previous
// version spent 2 launches per 64-column panel plus 2 per inverse level (~35 launches, ~0.6 ms for
// l = 576); here the whole thing is one kernel whose CTAs exchange 64x64 tiles through L2.
//
// Layout trick: every block product C = A B reads both operands "k-major" ([k][row] / [k][col]) from shared
// memory.  To get that layout straight from row-major global tiles, the factor is kept TRANSPOSED in the upper
// triangle of S (block (k,i), k < i, holds L[i,k]^T = R[k,i]), and the inverse is kept both plain (lower
// blocks of Linv) and transposed (upper blocks of Linv) until the final phase clears the upper blocks.
//
//   Phase F, step k:  every CTA factors the diagonal block S(k,k) redundantly (register-resident
//                     Gauss-Jordan sweep that yields X = L_kk^{-1} directly)            -> Xt in smem
//                     panel   L[i,k]^T -> S(k,i)      for i > k, dealt round-robin over the cluster
//                     --- cluster barrier ---
//                     update  S(j,i) -= L[j,k] L[i,k]^T   for k < j <= i, dealt round-robin
//                     --- cluster barrier ---
//   Phase I, level s = 1, 2, 4, ...: for each aligned pair of diagonal super-blocks [A 0; B C]
//                     pass 0: T = B A^{-1}             (staged in the free lower blocks of S)
//                     pass 1: Linv[C,A] = -C^{-1} T    (written plain and transposed)
//   Final:            zero the upper blocks of Linv, optionally emit its bf16 limbs (the GEMM operand).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "xkv_common.cuh"
#include "xkv_host.h"

namespace cg = cooperative_groups;

namespace xkv {

constexpr int CB = 64;             // block size
constexpr int CBK = CB + 4;        // shared-memory row stride of a k-major tile (rows stay 16-byte aligned)
constexpr int CTILE = CB * CBK;
constexpr int CH_THREADS = 256;
constexpr int CH_SMEM_FLOATS = 3 * CTILE + 2 * 128 + 2 * 64;

struct CholFusedParams {
  float* S[XKV_MAX_BATCH];
  float* Linv[XKV_MAX_BATCH];
  __nv_bfloat16* hi[XKV_MAX_BATCH];
  __nv_bfloat16* mid[XKV_MAX_BATCH];
  __nv_bfloat16* lo[XKV_MAX_BATCH];
  int l, nblk;
  long long ld, ld_limb;
  float pivot_floor, shift;
};

__device__ __forceinline__ void split3f(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  l = __float2bfloat16_rn(r2);
}

// row-major 64x64 global tile -> smem [k][idx] (stride CBK).  L2 loads: the tile was written by another CTA.
__device__ __forceinline__ void load_tile(float* dst, const float* src, long long ld) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = threadIdx.x + it * CH_THREADS;
    const int row = e >> 4, c4 = (e & 15) * 4;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + static_cast<long long>(row) * ld + c4));
    *reinterpret_cast<float4*>(dst + row * CBK + c4) = v;
  }
}

// acc[i][j] += sum_k Ak[k][tr+i] * Bk[k][tc+j]
__device__ __forceinline__ void tile_mma(float (&acc)[4][4], const float* Ak, const float* Bk) {
  const int tr = (threadIdx.x >> 4) * 4;
  const int tc = (threadIdx.x & 15) * 4;
#pragma unroll 8
  for (int k = 0; k < CB; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(Ak + k * CBK + tr);
    const float4 b = *reinterpret_cast<const float4*>(Bk + k * CBK + tc);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// Factor the 64x64 SPD block D (+ shift on the diagonal) and produce X = L^{-1} (lower triangular), with the
// augmented matrix [D | I] (64 x 128) held in registers: thread (cc, par) owns column cc, rows 2u + par.
// Right-looking sweep, one barrier per column j: the owners of row j and of column j publish them, then every
// thread applies  Aug[i][cc] -= D[i][j] * Aug[j][cc] / D[j][j]  to its rows i > j.  Row j of the identity half
// scaled by 1/sqrt(d) is row j of X (forward substitution in right-looking form).  L itself is never needed.
//   Xt[kk][c] = X[c][kk]   (k-major operand for  P = S[i,k] X^T  and for the inverse products)
//   x_plain (optional, global): X row-major with zeros above the diagonal.
// (No __restrict__ on the shared-memory pointers: rowj / lraw are written by OTHER threads between barriers, and
// with restrict the compiler forwards this thread's stale loads across __syncthreads.)
__device__ __forceinline__ void diag_factor(const float* D, long long ld, float shift, float pivot_floor, float* Xt,
                                            float* rowj, float* lraw, float* x_plain) {
  const int cc = threadIdx.x & 127, par = threadIdx.x >> 7;
  float v[32];
  if (cc < CB) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const int i = 2 * u + par;
      v[u] = __ldcg(D + static_cast<long long>(i) * ld + cc) + (i == cc ? shift : 0.f);
    }
  } else {
#pragma unroll
    for (int u = 0; u < 32; ++u) v[u] = (2 * u + par == cc - CB) ? 1.f : 0.f;
  }
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    float* rj = rowj + (j & 1) * 128;
    float* lr = lraw + (j & 1) * 64;
    const int u_first = (j / 2) & ~3;           // first 4-group that still contains a row > j
    if (par == (j & 1)) rj[cc] = v[j >> 1];
    if (cc == j) {
#pragma unroll
      for (int u4 = 0; u4 < 32; u4 += 4)
        if (u4 >= u_first)
          *reinterpret_cast<float4*>(lr + par * 32 + u4) = make_float4(v[u4], v[u4 + 1], v[u4 + 2], v[u4 + 3]);
    }
    __syncthreads();
    if (cc >= j) {                              // finished columns of the D half hold final values
      const float d = fmaxf(rj[j], pivot_floor);
      const float rjp = __fdividef(rj[cc], d);
#pragma unroll
      for (int u4 = 0; u4 < 32; u4 += 4) {
        if (u4 >= u_first) {
          const float4 l4 = *reinterpret_cast<const float4*>(lr + par * 32 + u4);
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int u = u4 + q;
            if (2 * u > j || (2 * u + 1 > j && par == 1)) v[u] = fmaf(-lv[q], rjp, v[u]);
          }
        }
      }
      if (par == (j & 1)) v[j >> 1] *= rsqrtf(d);
    }
  }
  // identity half now holds X: thread (64 + c, par) has X[2u + par][c]
  if (cc >= CB) {
    const int c = cc - CB;
#pragma unroll
    for (int u = 0; u < 32; ++u) Xt[c * CBK + 2 * u + par] = v[u];
    if (x_plain != nullptr) {
#pragma unroll
      for (int u = 0; u < 32; ++u) x_plain[static_cast<long long>(2 * u + par) * ld + c] = v[u];
    }
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1) chol_cluster_kernel(const __grid_constant__ CholFusedParams p) {
  extern __shared__ __align__(16) float chol_sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = static_cast<int>(cluster.num_blocks());
  const int c = static_cast<int>(cluster.block_rank());
  float* Ak = chol_sm;
  float* Bk = Ak + CTILE;
  float* Xt = Bk + CTILE;
  float* rowj = Xt + CTILE;
  float* lraw = rowj + 2 * 128;
  float* S = p.S[blockIdx.y];
  float* Li = p.Linv[blockIdx.y];
  const long long ld = p.ld;
  const int nblk = p.nblk;
  const int tid = threadIdx.x, tr = (tid >> 4) * 4, tc = (tid & 15) * 4;
  auto blk = [&](float* base, int bi, int bj) -> float* {
    return base + static_cast<long long>(bi) * CB * ld + static_cast<long long>(bj) * CB;
  };

  // ---------------- phase F ----------------
  for (int k = 0; k < nblk; ++k) {
    const bool owner = (c == k % CL);
    diag_factor(blk(S, k, k), ld, p.shift, p.pivot_floor, Xt, rowj, lraw, owner ? blk(Li, k, k) : nullptr);
    __syncthreads();
    for (int i = k + 1 + c; i < nblk; i += CL) {
      float* g = blk(S, k, i);
      load_tile(Ak, g, ld);
      __syncthreads();
      float acc[4][4] = {};
      tile_mma(acc, Ak, Xt);   // P[r][c'] = sum_kk S[i,k][r][kk] X[c'][kk]
#pragma unroll
      for (int b = 0; b < 4; ++b)   // stored transposed: S(k,i)[c'][r] = L[i,k][r][c']
        *reinterpret_cast<float4*>(g + static_cast<long long>(tc + b) * ld + tr) =
            make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
      __syncthreads();
    }
    cluster.sync();
    if (owner) {
      // X_k^T, the k-major operand of the inverse products, replaces the (now dead) diagonal block of S.  Written
      // only after the barrier: before it other CTAs may still be loading S(k,k) for their own factorisation.
      float* g = blk(S, k, k);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int e = tid + it * CH_THREADS;
        const int row = e >> 4, c4 = (e & 15) * 4;
        *reinterpret_cast<float4*>(g + static_cast<long long>(row) * ld + c4) =
            *reinterpret_cast<const float4*>(Xt + row * CBK + c4);
      }
    }
    const int nt = nblk - k - 1;
    const int ntiles = nt * (nt + 1) / 2;
    int ii = 0, base = 0;   // tile e = ii (ii + 1) / 2 + jj,  0 <= jj <= ii < nt
    for (int e = (c + CL - (k % CL)) % CL; e < ntiles; e += CL) {
      while (e >= base + ii + 1) {
        base += ii + 1;
        ++ii;
      }
      const int jj = e - base;
      const int i = k + 1 + ii, j = k + 1 + jj;
      load_tile(Ak, blk(S, k, j), ld);
      load_tile(Bk, blk(S, k, i), ld);
      __syncthreads();
      float acc[4][4] = {};
      tile_mma(acc, Ak, Bk);   // (L[j,k] L[i,k]^T)[r][c']
      float* g = blk(S, j, i);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float4* q = reinterpret_cast<float4*>(g + static_cast<long long>(tr + a) * ld + tc);
        float4 v = __ldcg(q);
        v.x -= acc[a][0];
        v.y -= acc[a][1];
        v.z -= acc[a][2];
        v.w -= acc[a][3];
        *q = v;
      }
      __syncthreads();
    }
    cluster.sync();
  }

  // ---------------- phase I ----------------
  for (int s = 1; s < nblk; s <<= 1) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      int cnt = 0;
      for (int i = 0; i < nblk; ++i) {
        const int a = (i / (2 * s)) * 2 * s;       // pair containing row i
        if (i < a + s) continue;                   // i must lie in the C half
        for (int j = a; j < a + s; ++j, ++cnt) {
          if ((cnt + i) % CL != c) continue;
          float acc[4][4] = {};
          const int t0 = pass == 0 ? j : a + s;
          const int t1 = pass == 0 ? a + s : i + 1;
          for (int t = t0; t < t1; ++t) {
            if (pass == 0) {
              load_tile(Ak, blk(S, t, i), ld);                          // L[i,t]^T
              load_tile(Bk, blk(Li, t, j), ld);                         // Linv[t,j]  (t == j: X_j, zero upper)
            } else {
              load_tile(Ak, t < i ? blk(Li, t, i) : blk(S, i, i), ld);  // Linv[i,t]^T  (t == i: X_i^T)
              load_tile(Bk, blk(S, t, j), ld);                          // T[t,j]
            }
            __syncthreads();
            tile_mma(acc, Ak, Bk);
            __syncthreads();
          }
          if (pass == 0) {
            float* g = blk(S, i, j);
#pragma unroll
            for (int r = 0; r < 4; ++r)
              *reinterpret_cast<float4*>(g + static_cast<long long>(tr + r) * ld + tc) =
                  make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
          } else {
            float* g = blk(Li, i, j);
            float* gt = blk(Li, j, i);
#pragma unroll
            for (int r = 0; r < 4; ++r)
              *reinterpret_cast<float4*>(g + static_cast<long long>(tr + r) * ld + tc) =
                  make_float4(-acc[r][0], -acc[r][1], -acc[r][2], -acc[r][3]);
#pragma unroll
            for (int b = 0; b < 4; ++b)
              *reinterpret_cast<float4*>(gt + static_cast<long long>(tc + b) * ld + tr) =
                  make_float4(-acc[0][b], -acc[1][b], -acc[2][b], -acc[3][b]);
          }
        }
      }
      cluster.sync();
    }
  }

  // ---------------- final: clear the transposed copies, emit bf16 limbs ----------------
  __nv_bfloat16* hi = p.hi[blockIdx.y];
  __nv_bfloat16* mid = p.mid[blockIdx.y];
  __nv_bfloat16* lo = p.lo[blockIdx.y];
  const int l4 = p.l >> 2;
  for (int r = c; r < p.l; r += CL) {
    const int bi = r / CB;
    float* row = Li + static_cast<long long>(r) * ld;
    for (int q = tid; q < l4; q += CH_THREADS) {
      const int col = q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col / CB > bi)
        *reinterpret_cast<float4*>(row + col) = v;
      else
        v = __ldcg(reinterpret_cast<const float4*>(row + col));
      if (hi != nullptr) {
        __nv_bfloat16 h[4], m[4], w[4];
        split3f(v.x, h[0], m[0], w[0]);
        split3f(v.y, h[1], m[1], w[1]);
        split3f(v.z, h[2], m[2], w[2]);
        split3f(v.w, h[3], m[3], w[3]);
        const long long o = static_cast<long long>(r) * p.ld_limb + col;
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(h);
        if (mid != nullptr) *reinterpret_cast<uint2*>(mid + o) = *reinterpret_cast<uint2*>(m);
        if (lo != nullptr) *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(w);
      }
    }
  }
}

}  // namespace xkv

using namespace xkv;

extern "C" int xkv_cholesky_inverse_limbs(float* const* S_host, float* const* Linv_host, void* const* hi_host,
                                          void* const* mid_host, void* const* lo_host, int batch, int l, int64_t ld,
                                          int64_t ld_limb, float shift, float pivot_floor, void* stream) {
  XKV_REQUIRE(S_host && Linv_host && batch >= 1 && batch <= XKV_MAX_BATCH, "cholesky: bad batch");
  XKV_REQUIRE(l > 0 && l % CB == 0, "cholesky: l=%d must be a positive multiple of %d", l, CB);
  XKV_REQUIRE(ld % 4 == 0 && ld >= l, "cholesky: ld must be a multiple of 4 and >= l");
  XKV_REQUIRE(!hi_host || (ld_limb % 4 == 0 && ld_limb >= l), "cholesky: ld_limb must be a multiple of 4 and >= l");
  CholFusedParams p;
  std::memset(&p, 0, sizeof(p));
  for (int b = 0; b < batch; ++b) {
    XKV_REQUIRE(S_host[b] && Linv_host[b], "cholesky: null matrix %d", b);
    XKV_REQUIRE((reinterpret_cast<uintptr_t>(S_host[b]) & 15) == 0 && (reinterpret_cast<uintptr_t>(Linv_host[b]) & 15) == 0,
                "cholesky: matrices must be 16-byte aligned");
    p.S[b] = S_host[b];
    p.Linv[b] = Linv_host[b];
    p.hi[b] = hi_host ? static_cast<__nv_bfloat16*>(hi_host[b]) : nullptr;
    p.mid[b] = mid_host ? static_cast<__nv_bfloat16*>(mid_host[b]) : nullptr;
    p.lo[b] = lo_host ? static_cast<__nv_bfloat16*>(lo_host[b]) : nullptr;
  }
  p.l = l;
  p.nblk = l / CB;
  p.ld = ld;
  p.ld_limb = ld_limb;
  p.pivot_floor = pivot_floor;
  p.shift = shift;
  const int smem = CH_SMEM_FLOATS * static_cast<int>(sizeof(float));
  static bool attr_set = false;
  if (!attr_set) {
    XKV_CHECK_CUDA(cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int CL = p.nblk >= 6 ? 8 : (p.nblk >= 3 ? 4 : (p.nblk == 2 ? 2 : 1));
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CL, batch, 1);
  cfg.blockDim = dim3(CH_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XKV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, chol_cluster_kernel, p));
  XKV_LAUNCHED();
  return 0;
}

extern "C" int xkv_cholesky_inverse(float* const* S_host, float* const* Linv_host, int batch, int l, int64_t ld,
                                    float shift, float pivot_floor, void* stream) {
  return xkv_cholesky_inverse_limbs(S_host, Linv_host, nullptr, nullptr, nullptr, batch, l, ld, 0, shift, pivot_floor,
                                    stream);
}
