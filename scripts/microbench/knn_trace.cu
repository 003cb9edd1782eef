// Per-tile timeline of ONE CTA of the TMA-fed kNN kernel under full load (300 clouds x 2048 points):
// clock64() at the pipeline's hand-over points, printed as deltas.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o knn_trace knn_trace.cu && ./knn_trace [C]
#include <cstdio>
#include <cstdlib>
#include <vector>
#define KNN_TRACE
#define KNN_TRACE_BX 5
#define KNN_TRACE_BY 150
thread_local long long r3dfs_launches = 0;
#include "../../r3dfsseg_b200/csrc/tc_knn.cu"

__global__ void norms_kernel(const float* x, int64_t M, int C, float* xx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float s = 0;
  for (int c = 0; c < C; ++c) s += x[i * C + c] * x[i * C + c];
  xx[i] = s;
}

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 64;
  const int B = 300, N = 2048, k = 20;
  const int64_t M = (int64_t)B * N;
  std::vector<float> h(M * C);
  srand(1);
  for (auto& v : h) v = rand() / (float)RAND_MAX;
  float *x, *xx;
  int32_t* idx;
  void* split;
  cudaMalloc(&x, sizeof(float) * M * C);
  cudaMalloc(&xx, sizeof(float) * M);
  cudaMalloc(&idx, sizeof(int32_t) * M * k);
  const size_t sb = knn_split_bytes(C, B, N, k);
  cudaMalloc(&split, sb);
  cudaMemcpy(x, h.data(), sizeof(float) * M * C, cudaMemcpyHostToDevice);
  norms_kernel<<<(unsigned)((M + 255) / 256), 256>>>(x, M, C, xx);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = launch_knn_tc(x, C, C, xx, B, N, k, idx, nullptr, 0, split, sb);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("rc %d err %s  %.3f ms (split + knn)\n", rc, cudaGetErrorString(e), ms);
  }
  std::vector<long long> t(8 * 4096);
  cudaMemcpyFromSymbol(t.data(), g_knn_trace, sizeof(long long) * 8 * 4096);
  const int T = N / 128;
  const long long t0 = t[0];
  printf("tile | prod:stage_free | mma:operands mma:acc_free mma:issued | sel:arrive sel:full sel:ld sel:released | (cycles from the first copy)\n");
  for (int j = 0; j < 2 * T; ++j)
    printf("%3d | %7lld | %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", j, t[j] - t0,
           t[4096 + j] - t0, t[2 * 4096 + j] - t0, t[3 * 4096 + j] - t0, t[7 * 4096 + j] - t0,
           t[4 * 4096 + j] - t0, t[5 * 4096 + j] - t0, t[6 * 4096 + j] - t0);
  const long long* e = t.data() + 6 * 4096 + 4000;
  printf("CTA: entry %lld | tmem allocated %lld | first copy 0 | pass 2 done %lld | drained %lld | lists stored %lld | exit %lld\n",
         e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0);
  // the split kernel alone
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(e0);
    if (C <= 16)
      knn_split_kernel<4><<<dim3(T, B), 256, KnnSplit<4>::BLK>>>(x, C, C, xx, N, (unsigned char*)split);
    else
      knn_split_kernel<16><<<dim3(T, B), 256, KnnSplit<16>::BLK>>>(x, C, C, xx, N, (unsigned char*)split);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("split kernel alone: %.3f ms (%s)\n", ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
