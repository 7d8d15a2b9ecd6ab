// Micro-benchmark: how fast can 148 persistent CTAs stream a (S x rk) bf16 token-factor matrix through a TMA ring,
// in the access pattern of the decode scores kernel (every one of `heads` CTA groups reads every 128-token tile,
// k-block by k-block)?  No MMAs: one consumer thread waits for each stage and releases it at once.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_tma_ring tools/probe_tma_ring.cu -lcuda
//   tools/probe_tma_ring <stages> <box_rows> <heads> <cluster> <hold_clk>
//
// stages x (box_rows x 128 B) bytes are in flight per CTA; cluster > 1 multicasts each box to `cluster` CTAs;
// hold_clk > 0 makes the consumer keep a stage for that many clocks (stands in for the MMAs that read it).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../xkv_b200/csrc/xkv_common.cuh"

using namespace xkv;

struct Params {
  CUtensorMap map;   // (S x rk), box {64, box_rows / cluster}
  int S, nkb, heads, stages, box_rows, cluster, hold;
};

__global__ void __launch_bounds__(128, 1) ring_kernel(const __grid_constant__ Params P, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = P.box_rows * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + P.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CL = P.cluster;
  const int crank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int nhb = P.heads / CL;
  const int hb = cid % nhb;
  const int slot = cid / nhb;
  const int nslots = (ncl - hb + nhb - 1) / nhb;
  const int ntiles = (P.S + P.box_rows - 1) / P.box_rows;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CL);
    }
    mbar_fence_init();
    tma_prefetch_desc(&P.map);
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  const uint16_t mask = static_cast<uint16_t>((1u << CL) - 1u);
  const int slice_rows = P.box_rows / CL;
  if (warp == 0 && lane == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int tile = slot; tile < ntiles; tile += nslots)
      for (int kb = 0; kb < P.nkb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        if (CL > 1)
          tma_load_2d_multicast(smem + s * stage_bytes + crank * slice_rows * 128, &P.map, &full_bar[s], kb * 64,
                                tile * P.box_rows + crank * slice_rows, mask);
        else
          tma_load_2d(smem + s * stage_bytes, &P.map, &full_bar[s], kb * 64, tile * P.box_rows);
        if (++s == P.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
  } else if (warp == 1 && lane == 0) {
    int s = 0;
    uint32_t ph = 0;
    unsigned long long acc = 0;
    for (int tile = slot; tile < ntiles; tile += nslots)
      for (int kb = 0; kb < P.nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        acc += *reinterpret_cast<volatile unsigned int*>(smem + s * stage_bytes);
        if (P.hold > 0) {
          const long long t0 = clock64();
          while (clock64() - t0 < P.hold) {
          }
        }
        if (CL > 1) {
          for (int r = 0; r < CL; ++r) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty_bar[s])), "r"(r));
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
          }
        } else {
          mbar_arrive(&empty_bar[s]);
        }
        if (++s == P.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    if (acc == 0x12345678ull) sink[0] = acc;
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
}

int main(int argc, char** argv) {
  const int stages = argc > 1 ? atoi(argv[1]) : 5;
  const int box_rows = argc > 2 ? atoi(argv[2]) : 128;
  const int heads = argc > 3 ? atoi(argv[3]) : 8;
  const int cluster = argc > 4 ? atoi(argv[4]) : 1;
  const int hold = argc > 5 ? atoi(argv[5]) : 0;
  const int S = argc > 6 ? atoi(argv[6]) : 65536;
  const int rk = 512;
  __nv_bfloat16* A;
  cudaMalloc(&A, static_cast<size_t>(S) * rk * 2);
  cudaMemset(A, 0, static_cast<size_t>(S) * rk * 2);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &q);
  Params P;
  memset(&P, 0, sizeof(P));
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(rk), static_cast<cuuint64_t>(S)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(rk) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows / cluster)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(&P.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    return 1;
  }
  P.S = S, P.nkb = rk / 64, P.heads = heads, P.stages = stages, P.box_rows = box_rows, P.cluster = cluster, P.hold = hold;
  const size_t smem = static_cast<size_t>(stages) * box_rows * 128 + 1024 + 16 * stages + 64;
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster > 1 ? 1 : 0;
  int ncl = sms / cluster;
  if (cluster > 1) {
    cfg.gridDim = dim3(cluster, 1, 1);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, ring_kernel, &cfg) == cudaSuccess && n > 0 && n < ncl) ncl = n;
  }
  ncl = ncl / (heads / cluster) * (heads / cluster);   // whole rounds of head blocks
  cfg.gridDim = dim3(ncl * cluster, 1, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaLaunchKernelEx(&cfg, ring_kernel, P, sink);
  cudaDeviceSynchronize();
  const int reps = 10;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, ring_kernel, P, sink);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  cudaError_t err = cudaGetLastError();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = 1e3 * ms / reps;
  const double bytes = static_cast<double>(S) * rk * 2 * heads;
  printf("{\"stages\": %d, \"box_rows\": %d, \"heads\": %d, \"cluster\": %d, \"hold_clk\": %d, \"ctas\": %d, \"in_flight_KiB\": %d, "
         "\"us\": %.1f, \"sm_ingest_TBps\": %.2f, \"err\": \"%s\"}\n",
         stages, box_rows, heads, cluster, hold, ncl * cluster, stages * box_rows / 8, us, bytes / us / 1e6,
         cudaGetErrorString(err));
  return 0;
}
