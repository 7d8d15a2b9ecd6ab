// Micro-benchmark + layout check of tcgen05.mma (kind::f16, bf16 operands) issue rates on sm_100a:
// how many clocks does one instruction of a given shape take when it is issued back to back, with the operands
// resident (no TMA, no waits), in the forms the decode scores kernel can use?
//
//   form 0: cta_group::1, A and B in shared memory (SS)         M = 128
//   form 1: cta_group::1, A in tensor memory (TS)               M = 128
//   form 2: cta_group::2, SS: a CTA pair, M = 256 (128 token rows per CTA), B split N/2 rows per CTA
//   form 3: cta_group::2, TS
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_mma_rate tools/probe_mma_rate.cu
//   tools/probe_mma_rate rate   <form> <N> <mmas> [b_blocks a_stages commit_every]  -> one JSON line: clocks per instruction (busiest CTA), time
//   tools/probe_mma_rate verify <form> <N>              -> checks D against a scalar product on integer data
//
// Operand layout: K-major tiles with the 128-byte swizzle, exactly as TMA writes a {64, rows} box (xkv_common.cuh).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../xkv_b200/csrc/xkv_common.cuh"

using namespace xkv;

template <int CG>
__device__ __forceinline__ void p_tmem_alloc(uint32_t* slot, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void p_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int CG, int TS>
__device__ __forceinline__ void p_mma(uint32_t d, uint64_t adesc, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1 && TS == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  if constexpr (CG == 2 && TS == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  if constexpr (CG == 1 && TS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  if constexpr (CG == 2 && TS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void p_commit(uint64_t* bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

struct Params {
  int N, mmas, verify;
  int random_data;   // rate mode: operands = hashed bit patterns with sane exponents instead of a constant ramp
  int b_blocks, a_stages, commit_every;   // rate mode: distinct B blocks cycled (right-factor footprint), ring stages, MMA groups per commit (0: one at the end)
  const __nv_bfloat16* A;   // verify: (CG * 128) x 64, row-major
  const __nv_bfloat16* B;   // verify: N x 64, row-major
  float* D;                 // verify: (CG * 128) x N
  unsigned long long* clk;  // per CTA
};

constexpr int A_STAGES = 5;
constexpr int B_REGION = 128 * 1024;
constexpr int A_BYTES = 128 * 64 * 2;

// row r, column k of a K-major SW128 tile (rows of 128 bytes, 16-byte chunks XORed with the row's low 3 bits)
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return static_cast<uint32_t>(r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}

template <int CG, int TS>
__global__ void __launch_bounds__(192, 1) mma_kernel(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                               // A_STAGES x 16 KiB
  uint8_t* sB = smem + A_STAGES * A_BYTES;          // this CTA's B rows (N / CG) x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + B_REGION);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  const int crank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int nb_rows = P.N / CG;
  // operands
  if (P.verify) {
    for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
      const int r = e >> 6, k = e & 63;
      *reinterpret_cast<__nv_bfloat16*>(sA + sw128_off(r, k)) = P.A[(crank * 128 + r) * 64 + k];
    }
    for (int e = threadIdx.x; e < nb_rows * 64; e += blockDim.x) {
      const int r = e >> 6, k = e & 63;
      *reinterpret_cast<__nv_bfloat16*>(sB + sw128_off(r, k)) = P.B[(crank * nb_rows + r) * 64 + k];
    }
  } else {
    for (int e = threadIdx.x; e < (A_STAGES * A_BYTES + B_REGION) / 4; e += blockDim.x)
    {
      uint32_t w = 0x3C003C00u + (e & 0xff);
      if (P.random_data) {
        uint32_t h = static_cast<uint32_t>(e) * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        w = (h & 0x80FF80FFu) | 0x3F003F00u;   // random sign and mantissa, exponents 126 / 127
      }
      reinterpret_cast<uint32_t*>(smem)[e] = w;
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) p_tmem_alloc<CG>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(CG * 128, P.N, 0, 0);
  if (TS && P.verify) {
    // A operand into tensor memory: lane = row, column c holds k = 2c, 2c + 1 (packed bf16), at columns 256..287
    if (warp >= 2) {   // warps 2..5: lane quarter = warp & 3
      const int row = (warp & 3) * 32 + (threadIdx.x & 31);
      uint32_t w[16];
      for (int half = 0; half < 2; ++half) {
        for (int c = 0; c < 16; ++c) {
          const int k = half * 32 + 2 * c;
          const uint16_t lo = *reinterpret_cast<const uint16_t*>(sA + sw128_off(row, k));
          const uint16_t hi = *reinterpret_cast<const uint16_t*>(sA + sw128_off(row, k + 1));
          w[c] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
        }
        tmem_st_32x16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256u + half * 16, w);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
  }
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 0) {
    t0 = clock64();
    if (crank == 0) {
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      if (elect_one()) {
        const int groups = P.verify ? 1 : P.mmas / 4;
        for (int g = 0; g < groups; ++g) {
          const uint32_t a_st = a_base + static_cast<uint32_t>(g % P.a_stages) * A_BYTES;
          const uint32_t b_blk = b_base + static_cast<uint32_t>(g % P.b_blocks) * static_cast<uint32_t>(nb_rows * 128);
          const uint32_t d = tmem_base + static_cast<uint32_t>((g & 1) * (P.N <= 128 ? 128 : 0));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            p_mma<CG, TS>(d, umma_desc_sw128(a_st + k * 32, 16, 1024), tmem_base + 256u + static_cast<uint32_t>(k * 8),
                          umma_desc_sw128(b_blk + k * 32, 16, 1024), idesc, (P.verify ? k > 0 : 1) ? 1u : 0u);
          if (P.commit_every > 0 && (g + 1) % P.commit_every == 0) p_commit<CG>(&bar[1]);
        }
        p_commit<CG>(&bar[0]);
      }
      __syncwarp();
    }
    mbar_wait(&bar[0], 0);
    t1 = clock64();
    tc_fence_after();
    if (threadIdx.x == 0) P.clk[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (P.verify && warp >= 2) {
    tc_fence_after();
    const int row = (warp & 3) * 32 + (threadIdx.x & 31);
    for (int c0 = 0; c0 < P.N; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      for (int c = 0; c < 16; ++c) P.D[static_cast<long long>(crank * 128 + row) * P.N + c0 + c] = __uint_as_float(v[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    p_tmem_dealloc<CG>(tmem_base, 512);
  }
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

template <int CG, int TS>
static void run(const Params& p, int grid, float* ms) {
  const int smem = A_STAGES * A_BYTES + B_REGION + 1024 + 64;
  CK(cudaFuncSetAttribute(mma_kernel<CG, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(192, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, mma_kernel<CG, TS>, p));   // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, mma_kernel<CG, TS>, p));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(ms, e0, e1));
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s rate|verify <form> <N> [mmas]\n", argv[0]);
    return 2;
  }
  const bool verify = strcmp(argv[1], "verify") == 0;
  const int form = atoi(argv[2]), N = atoi(argv[3]);
  const int mmas = argc > 4 ? atoi(argv[4]) : 1024;
  const int b_blocks = argc > 5 ? atoi(argv[5]) : 1;
  const int a_stages = argc > 6 ? atoi(argv[6]) : 4;
  const int commit_every = argc > 7 ? atoi(argv[7]) : 0;
  const int random_data = argc > 8 ? atoi(argv[8]) : 0;
  const int CG = form >= 2 ? 2 : 1, TS = form & 1;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = verify ? CG : (sms / CG) * CG;
  Params p;
  memset(&p, 0, sizeof(p));
  p.N = N;
  p.mmas = mmas;
  p.verify = verify ? 1 : 0;
  p.random_data = random_data;
  p.b_blocks = b_blocks;
  p.a_stages = a_stages;
  p.commit_every = commit_every;
  if (a_stages < 1 || a_stages > A_STAGES || b_blocks < 1 || b_blocks * (N / CG) * 128 > B_REGION) {
    fprintf(stderr, "bad b_blocks / a_stages\n");
    return 2;
  }
  const int M = CG * 128;
  std::vector<__nv_bfloat16> hA(M * 64), hB(N * 64);
  for (int i = 0; i < M * 64; ++i) hA[i] = __float2bfloat16(static_cast<float>((i * 7 + (i >> 6) * 3) % 13 - 6));
  for (int i = 0; i < N * 64; ++i) hB[i] = __float2bfloat16(static_cast<float>((i * 5 + (i >> 6) * 11) % 9 - 4));
  __nv_bfloat16 *dA, *dB;
  float* dD;
  unsigned long long* dclk;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, static_cast<size_t>(M) * N * 4));
  CK(cudaMalloc(&dclk, grid * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, static_cast<size_t>(M) * N * 4));
  CK(cudaMemset(dclk, 0, grid * 8));
  p.A = dA;
  p.B = dB;
  p.D = dD;
  p.clk = dclk;
  float ms = 0.f;
  if (form == 0) run<1, 0>(p, grid, &ms);
  if (form == 1) run<1, 1>(p, grid, &ms);
  if (form == 2) run<2, 0>(p, grid, &ms);
  if (form == 3) run<2, 1>(p, grid, &ms);
  if (verify) {
    std::vector<float> hD(static_cast<size_t>(M) * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += __bfloat162float(hA[i * 64 + k]) * __bfloat162float(hB[j * 64 + k]);
        const double err = fabs(ref - hD[static_cast<size_t>(i) * N + j]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3 && bad++ < 5) fprintf(stderr, "mismatch at (%d, %d): %f vs %f\n", i, j, hD[static_cast<size_t>(i) * N + j], ref);
      }
    printf("{\"mode\": \"verify\", \"form\": %d, \"cta_group\": %d, \"a_from_tmem\": %d, \"M\": %d, \"N\": %d, \"mismatches\": %d, \"max_abs_err\": %g}\n",
           form, CG, TS, M, N, bad, maxerr);
    return bad ? 1 : 0;
  }
  std::vector<unsigned long long> hclk(grid);
  CK(cudaMemcpy(hclk.data(), dclk, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0;
  for (int i = 0; i < grid; ++i)
    if (hclk[i] > mx) mx = hclk[i];
  const int issued = (mmas / 4) * 4;
  const double flop = 2.0 * 128 * N * 16 * issued * grid;   // per CTA: 128 rows x N x 16 per instruction
  printf("{\"mode\": \"rate\", \"form\": %d, \"cta_group\": %d, \"a_from_tmem\": %d, \"M_per_cta\": 128, \"N\": %d, \"ctas\": %d, \"mmas_per_issuer\": %d, \"b_blocks\": %d, \"a_stages\": %d, \"commit_every\": %d, \"random_data\": %d, "
         "\"clk_per_mma\": %.1f, \"ideal_clk\": %.1f, \"ms\": %.4f, \"tflops\": %.1f}\n",
         form, CG, TS, N, grid, issued, b_blocks, a_stages, commit_every, random_data, static_cast<double>(mx) / issued, 128.0 * N * 16 * 2 / 8192.0, ms, flop / (ms * 1e-3) / 1e12);
  return 0;
}
