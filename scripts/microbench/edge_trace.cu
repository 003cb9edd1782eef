// Per-tile timeline of ONE CTA of edge_tc_kernel under full load (300 clouds x 2048 points, k = 20):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o edge_trace edge_trace.cu && ./edge_trace
#include <cstdio>
#include <cstdlib>
#include <vector>
#define EDGE_TRACE
#define EDGE_TRACE_BX 3
#define EDGE_TRACE_BY 150
thread_local long long r3dfs_launches = 0;
#include "../../r3dfsseg_b200/csrc/tc_edge.cu"

int main() {
  const int B = 300, N = 2048, k = 20;
  const int64_t M = (int64_t)B * N;
  std::vector<float> h(M * 128);
  std::vector<int> hi(M * k);
  srand(1);
  for (auto& v : h) v = rand() / (float)RAND_MAX - 0.5f;
  for (auto& v : hi) v = rand() % N;
  float *pq, *w2, *s2, *t2, *y;
  int* idx;
  cudaMalloc(&pq, sizeof(float) * M * 128);
  cudaMalloc(&idx, sizeof(int) * M * k);
  cudaMalloc(&w2, sizeof(float) * 64 * 64);
  cudaMalloc(&s2, sizeof(float) * 64);
  cudaMalloc(&t2, sizeof(float) * 64);
  cudaMalloc(&y, sizeof(float) * M * 64);
  cudaMemcpy(pq, h.data(), sizeof(float) * M * 128, cudaMemcpyHostToDevice);
  cudaMemcpy(idx, hi.data(), sizeof(int) * M * k, cudaMemcpyHostToDevice);
  cudaMemcpy(w2, h.data(), sizeof(float) * 64 * 64, cudaMemcpyHostToDevice);
  cudaMemcpy(s2, h.data(), sizeof(float) * 64, cudaMemcpyHostToDevice);
  cudaMemcpy(t2, h.data(), sizeof(float) * 64, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = launch_edge_mlp_tc(pq, idx, w2, s2, t2, B, N, k, y, 64, identity_map(), 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("rc %d err %s  %.3f ms\n", rc, cudaGetErrorString(e), ms);
  }
  std::vector<long long> t(8 * 256);
  cudaMemcpyFromSymbol(t.data(), g_edge_trace, sizeof(long long) * 8 * 256);
  const long long t0 = t[3 * 256];
  printf("tile | prod:top gathered stage_free stored | mma:operands acc_free issued | epi:released\n");
  for (int u = 0; u < 40; ++u)
    printf("%3d | %6lld %6lld %6lld %6lld | %6lld %6lld %6lld | %6lld\n", u, t[3 * 256 + u] - t0,
           t[4 * 256 + u] - t0, t[5 * 256 + u] - t0, t[6 * 256 + u] - t0, t[u] - t0, t[256 + u] - t0,
           t[2 * 256 + u] - t0, t[7 * 256 + u] - t0);
  return 0;
}
