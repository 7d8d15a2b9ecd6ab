// Bisect of the decode scores kernel on the device: the library's own translation units compiled with XKV_PROBE
// (switches that are compiled out of libxkv_b200.so), the scores launch timed alone with CUDA events.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/probe_decode_scores \
//        tools/probe_decode_scores.cu -lcuda
//   tools/probe_decode_scores <dbg bits> [kernel: 0 = CTA pairs, 1 = single CTAs] [rk] [S] [stages]
// bits: 1 no reconstruction MMAs, 2 ring stages released by a plain mbarrier arrive instead of tcgen05.commit,
//       4 no epilogue pipeline at all, 8 no RoPE (no cos / sin loads), 16 epilogue stops after reading the accumulator
#define XKV_PROBE 1
#include "../xkv_b200/csrc/xkv_capi.cu"
#include "../xkv_b200/csrc/xkv_gemm.cu"
#include "../xkv_b200/csrc/xkv_decode.cu"

__global__ void clock_rate_kernel(double* mhz) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < 200000ull);
  const long long c1 = clock64();
  *mhz = static_cast<double>(c1 - c0) / static_cast<double>(t1 - t0) * 1e3;
}

int main(int argc, char** argv) {
  const int dbg = argc > 1 ? atoi(argv[1]) : 0;
  const int cluster = argc > 2 ? atoi(argv[2]) : 1;
  const int rk = argc > 3 ? atoi(argv[3]) : 512;
  const int S = argc > 4 ? atoi(argv[4]) : 65536;
  const int H = 8, D = 128, Hq = 32;
  __nv_bfloat16 *A, *B, *q, *cs, *sn;
  float* scores;
  cudaMalloc(&A, static_cast<size_t>(S) * rk * 2);
  cudaMalloc(&B, static_cast<size_t>(H) * D * rk * 2);
  cudaMalloc(&q, Hq * D * 2);
  cudaMalloc(&cs, static_cast<size_t>(S) * D * 2);
  cudaMalloc(&sn, static_cast<size_t>(S) * D * 2);
  cudaMalloc(&scores, static_cast<size_t>(Hq) * (S + 64) * 4);
  cudaMemset(A, 0x3c, static_cast<size_t>(S) * rk * 2);
  cudaMemset(B, 0x3c, static_cast<size_t>(H) * D * rk * 2);
  cudaMemset(q, 0x3c, Hq * D * 2);
  cudaMemset(cs, 0x3c, static_cast<size_t>(S) * D * 2);
  cudaMemset(sn, 0x3c, static_cast<size_t>(S) * D * 2);
  xkv::g_probe_dbg = dbg;
  xkv::g_probe_stages = argc > 5 ? atoi(argv[5]) : 0;
  if (cluster == 1) xkv::g_scores_variant = 2;
  if (cluster == 0) xkv::g_scores_variant = 3;   // CTA pairs
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 300; ++i)
    if (xkv::launch_scores(q, Hq, H, D, A, rk, rk, B, rk, S, cs, sn, D, 0.088f, scores, S + 64, 0)) {
      printf("launch failed: %s\n", xkv_last_error());
      return 1;
    }
  cudaDeviceSynchronize();
  const int reps = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) xkv::launch_scores(q, Hq, H, D, A, rk, rk, B, rk, S, cs, sn, D, 0.088f, scores, S + 64, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double* mhz_d;
  double mhz = 0;
  cudaMalloc(&mhz_d, 8);
  clock_rate_kernel<<<1, 1>>>(mhz_d);
  cudaMemcpy(&mhz, mhz_d, 8, cudaMemcpyDeviceToHost);
  printf("{\"dbg\": %d, \"cluster\": %d, \"rk\": %d, \"S\": %d, \"stages\": %d, \"us\": %.1f, \"sm_mhz\": %.0f, \"err\": \"%s\"}\n", dbg, cluster, rk, S, xkv::g_probe_stages,
         1e3 * ms / reps, mhz, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
