// xkv_b200 — prefill gather ("pack") of a layer group's K or V into the token-major matrix X.
//
// Replaces torch.cat(keys, dim=1) (fake_layer_merge_dynamic_cache.py:170-171) followed by
// tensor.transpose(1, 2).reshape(bs, sl, nh*hd) (:13-14): column order (layer, head, dim),
// row = token, batch outermost.  The reference applies no centering or scaling,
// so the gather is a pure strided copy: HBM-bound, 4 B of traffic per bf16 element
// (one read + one write).
//
// Each thread moves 16-byte vectors; a (token, layer, head) segment is head_dim*2 bytes
// contiguous on both sides, so reads and writes are full 128-byte lines. The grid is a
// multiple of the SM count and grid-strides over the vectors with 4 loads in flight.
#include "xkv_common.cuh"
#include "xkv_host.h"

namespace xkv {

struct PackParams {
  const uint4* src[XKV_MAX_GROUP_LAYERS];  // per-layer base pointers
  uint4* X;
  long long stride_b, stride_h, stride_s;  // in 16-byte vectors
  int num_layers, bs, heads, seq, vec_per_head;  // vec_per_head = head_dim / 8
  int rows_per_cta;
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// DIR = 0: layers -> X (pack), DIR = 1: X -> layers (unpack)
// grid = (row chunks, bs). A thread owns a fixed 16-byte column slot of X, so the
// (layer, head, dim) decomposition is done once; walking down the tokens is pointer increments.
template <int DIR>
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackParams p) {
  const int b = blockIdx.y;
  const int vph = p.vec_per_head;
  const int row_vecs = p.num_layers * p.heads * vph;  // vectors per X row
  const int s0 = blockIdx.x * p.rows_per_cta;
  const int s1 = min(s0 + p.rows_per_cta, p.seq);
  constexpr int ILP = 4;
  for (int c = threadIdx.x; c < row_vecs; c += blockDim.x) {
    const int lh = c / vph;
    const int dv = c - lh * vph;
    const int l = lh / p.heads;
    const int h = lh - l * p.heads;
    const uint4* layer = p.src[l] + b * p.stride_b + h * p.stride_h + dv;
    uint4* x = p.X + (static_cast<long long>(b) * p.seq) * row_vecs + c;
    for (int s = s0; s < s1; s += ILP) {
      uint4 v[ILP];
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        if (s + u < s1) {
          if (DIR == 0)
            v[u] = ld_stream(layer + (s + u) * p.stride_s);
          else
            v[u] = ld_stream(x + static_cast<long long>(s + u) * row_vecs);
        }
      }
#pragma unroll
      for (int u = 0; u < ILP; ++u) {
        if (s + u < s1) {
          if (DIR == 0)
            st_stream(x + static_cast<long long>(s + u) * row_vecs, v[u]);
          else
            st_stream(const_cast<uint4*>(layer) + (s + u) * p.stride_s, v[u]);
        }
      }
    }
  }
}

static int launch_pack(int dir, const void* const* layer_ptrs, int num_layers, int bs, int heads, int seq,
                       int head_dim, long long stride_b, long long stride_h, long long stride_s, void* X,
                       void* stream) {
  if (seq == 0 && num_layers >= 1) return 0;  // empty prefill: nothing to gather (pointers may be null)
  XKV_REQUIRE(layer_ptrs != nullptr && X != nullptr, "pack: null pointer");
  XKV_REQUIRE(num_layers >= 1 && num_layers <= XKV_MAX_GROUP_LAYERS, "pack: num_layers=%d (max %d)", num_layers,
              XKV_MAX_GROUP_LAYERS);
  XKV_REQUIRE(bs >= 1 && heads >= 1 && seq >= 0 && head_dim >= 8, "pack: bad shape");
  XKV_REQUIRE(head_dim % 8 == 0, "pack: head_dim must be a multiple of 8 (16-byte vectors)");
  XKV_REQUIRE(stride_b % 8 == 0 && stride_h % 8 == 0 && stride_s % 8 == 0,
              "pack: strides must be multiples of 8 elements");
  XKV_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, "pack: X must be 16-byte aligned");
  PackParams p;
  std::memset(&p, 0, sizeof(p));
  for (int l = 0; l < num_layers; ++l) {
    XKV_REQUIRE(layer_ptrs[l] != nullptr && (reinterpret_cast<uintptr_t>(layer_ptrs[l]) & 15) == 0,
                "pack: layer %d pointer null or not 16-byte aligned", l);
    p.src[l] = reinterpret_cast<const uint4*>(layer_ptrs[l]);
  }
  p.X = reinterpret_cast<uint4*>(X);
  p.stride_b = stride_b / 8;
  p.stride_h = stride_h / 8;
  p.stride_s = stride_s / 8;
  p.num_layers = num_layers;
  p.bs = bs;
  p.heads = heads;
  p.seq = seq;
  p.vec_per_head = head_dim / 8;
  if (seq == 0) return 0;  // empty prefill: nothing to gather
  int dev = 0, sms = 148;
  XKV_CHECK_CUDA(cudaGetDevice(&dev));
  XKV_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // aim for ~8 CTAs per SM over the whole grid, at least 4 token rows per CTA
  int chunks = (sms * 8 + bs - 1) / bs;
  int rows = (seq + chunks - 1) / chunks;
  if (rows < 4) rows = 4;
  p.rows_per_cta = rows;
  chunks = (seq + rows - 1) / rows;
  dim3 grid(chunks, bs);
  if (dir == 0)
    pack_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(p);
  else
    pack_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(p);
  XKV_LAUNCHED();
  return 0;
}

}  // namespace xkv

extern "C" int xkv_pack_group(const void* const* layer_ptrs_host, int num_layers, int bs, int heads, int seq,
                              int head_dim, int64_t stride_b, int64_t stride_h, int64_t stride_s, void* X,
                              void* stream) {
  return xkv::launch_pack(0, layer_ptrs_host, num_layers, bs, heads, seq, head_dim, stride_b, stride_h, stride_s, X,
                          stream);
}

extern "C" int xkv_unpack_group(const void* X, int num_layers, int bs, int heads, int seq, int head_dim,
                                int64_t stride_b, int64_t stride_h, int64_t stride_s, void* const* layer_ptrs_host,
                                void* stream) {
  return xkv::launch_pack(1, const_cast<const void* const*>(layer_ptrs_host), num_layers, bs, heads, seq, head_dim,
                          stride_b, stride_h, stride_s, const_cast<void*>(X), stream);
}
