// Per-tile timeline of ONE CTA of attention_tc2_kernel under full load (300 clouds x 2048 points):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o att_trace att_trace.cu && ./att_trace [scale]
#include <cstdio>
#include <cstdlib>
#include <vector>
#define ATT_TRACE
#define ATT_TRACE_BX 5
#define ATT_TRACE_BY 150
thread_local long long r3dfs_launches = 0;
#include "../../r3dfsseg_b200/csrc/tc_attention.cu"

int main(int argc, char** argv) {
  const float scale = argc > 1 ? atof(argv[1]) : 3.0f;  // large: the first sweep runs
  const int B = 300, N = 2048;
  const int64_t M = (int64_t)B * N;
  std::vector<float> h(M * 192);
  srand(1);
  for (auto& v : h) v = scale * (rand() / (float)RAND_MAX - 0.5f);
  float *qkv, *y, *kmax;
  void* split;
  cudaMalloc(&qkv, sizeof(float) * M * 192);
  cudaMalloc(&y, sizeof(float) * M * 64);
  cudaMalloc(&kmax, sizeof(float) * B);
  const size_t sb = attention_split_bytes(B, N);
  cudaMalloc(&split, sb);
  cudaMemcpy(qkv, h.data(), sizeof(float) * M * 192, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    int rc = launch_attention_tc(qkv, 192, B, N, y, 64, identity_map(), 0, kmax, split, sb);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("rc %d err %s  %.3f ms (kmax + split + attention)\n", rc, cudaGetErrorString(e), ms);
  }
  std::vector<long long> t(8 * 256);
  cudaMemcpyFromSymbol(t.data(), g_att_trace, sizeof(long long) * 8 * 256);
  const long long t0 = t[0];
  printf("S tiles (g): K landed, S issued\n");
  for (int g = 0; g < 32; ++g) printf("%2d: %7lld %7lld\n", g, t[g] - t0, t[256 + g] - t0);
  printf("sweep-2 tiles (j): P(A) ready | PV(A) issued | P(B) ready | PV(B) issued || workers: S seen, P(B) written\n");
  for (int j = 0; j < 16; ++j)
    printf("%2d: %7lld %7lld %7lld %7lld || %7lld %7lld\n", j, t[2 * 256 + j] - t0, t[4 * 256 + j] - t0,
           t[3 * 256 + j] - t0, t[5 * 256 + j] - t0, t[6 * 256 + j] - t0, t[7 * 256 + j] - t0);
  return 0;
}
