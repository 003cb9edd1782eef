// Per-k-block timeline of ONE CTA of linear_tma_kernel under full load (M = 614400 rows):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o linear_trace linear_trace.cu && ./linear_trace K N
#include <cstdio>
#include <cstdlib>
#include <vector>
#define LIN_TRACE
#define LIN_TRACE_BX 9000
thread_local long long r3dfs_launches = 0;
#include "../../r3dfsseg_b200/csrc/tc_gemm_tma.cu"

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 192, N = argc > 2 ? atoi(argv[2]) : 512;
  const int64_t M = 614400;
  float *x, *w, *y;
  cudaMalloc(&x, sizeof(float) * M * K);
  cudaMalloc(&w, sizeof(float) * N * K);
  cudaMalloc(&y, sizeof(float) * M * N);
  cudaMemset(x, 0, sizeof(float) * M * K);
  cudaMemset(w, 0, sizeof(float) * N * K);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = launch_linear_tma(x, K, w, nullptr, nullptr, ACT_LRELU, M, K, N, y, N, identity_map(), 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("rc %d err %s  %.3f ms  (M %lld K %d N %d)\n", rc, cudaGetErrorString(e), ms, (long long)M, K, N);
  }
  std::vector<long long> t(8 * 1024);
  cudaMemcpyFromSymbol(t.data(), g_lin_trace, sizeof(long long) * 8 * 1024);
  const long long t0 = t[7 * 1024 + 2];
  printf("kb | mma:operands mma:issued | cvt:loop raw_landed regs_loaded stage_free stored\n");
  for (int kb = 0; kb < K / 16; ++kb)
    printf("%2d | %6lld %6lld | %6lld %6lld %6lld %6lld %6lld\n", kb, t[kb] - t0, t[1024 + kb] - t0,
           t[2 * 1024 + kb] - t0, t[3 * 1024 + kb] - t0, t[4 * 1024 + kb] - t0, t[5 * 1024 + kb] - t0,
           t[6 * 1024 + kb] - t0);
  printf("all MMAs done %lld | epilogue done %lld\n", t[7 * 1024] - t0, t[7 * 1024 + 1] - t0);
  return 0;
}
