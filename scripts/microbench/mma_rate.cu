// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (K = 8 per instruction) on sm_100a with the
// K-major no-swizzle operand layout the library uses, as a function of N, of the number of
// independent TMEM accumulators the instruction stream alternates between, and of CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../r3dfsseg_b200/csrc/tc.cuh"

template <int N, int NACC, int KIND>  // KIND 0 = tf32 (K=8), 1 = bf16 (K=16)
__global__ void mma_rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  constexpr int LBOA = tc::tile_lbo(128), LBOB = tc::tile_lbo(N);
  constexpr int A_BYTES = 16 * LBOA;  // K = 64 tf32: 16 chunks
  for (int i = tid; i < (A_BYTES + 16 * LBOB) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  if (tid < 32) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  uint32_t idesc = tc::make_idesc_tf32(128, N);
  if (KIND == 1) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t a = tc::smem_u32(smem), b = a + A_BYTES;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = tc::make_desc(a + ks * 2 * LBOA, LBOA, 128);
        const uint64_t db = tc::make_desc(b + ks * 2 * LBOB, LBOB, 128);
        const uint32_t d = tmem_d + ((it * 8 + ks) % NACC) * N;
        if (KIND == 0) {
          tc::mma_tf32(d, da, db, idesc, 1);
        } else if (KIND == 2) {  // A operand from TMEM (columns 448..511), B from shared memory
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
              "r"(tmem_d + 448 + ks * 8), "l"(db), "r"(idesc), "r"(1)
              : "memory");
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
              "l"(da), "l"(db), "r"(idesc), "r"(1)
              : "memory");
        }
      }
    }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem_d, 512);
}

template <int N, int NACC, int KIND>
void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  const int iters = 2000;
  const int smem = 16 * tc::tile_lbo(128) + 16 * tc::tile_lbo(N) + 256;
  cudaFuncSetAttribute(mma_rate_kernel<N, NACC, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<N, NACC, KIND><<<grid, 128, smem>>>(10, d);
  cudaDeviceSynchronize();
  mma_rate_kernel<N, NACC, KIND><<<grid, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = mx / (iters * 8.0);
  const double kk = KIND == 1 ? 16 : 8;
  printf("%-28s grid %4d  clk/MMA %7.1f  MAC/clk/SM %7.1f  (%s)\n", name, grid, per,
         128.0 * N * kk / per * (grid > 148 ? 2 : 1), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 1, 0>("tf32 N=64  1 acc", 148);
  run<64, 2, 0>("tf32 N=64  2 acc", 148);
  run<64, 4, 0>("tf32 N=64  4 acc", 148);
  run<128, 1, 0>("tf32 N=128 1 acc", 148);
  run<128, 2, 0>("tf32 N=128 2 acc", 148);
  run<128, 4, 0>("tf32 N=128 4 acc", 148);
  run<256, 1, 0>("tf32 N=256 1 acc", 148);
  run<256, 2, 0>("tf32 N=256 2 acc", 148);
  run<128, 1, 0>("tf32 N=128 1 acc 2 CTA/SM", 296);
  run<64, 1, 0>("tf32 N=64  1 acc 2 CTA/SM", 296);
  run<64, 1, 2>("tf32 N=64  A in TMEM", 148);
  run<64, 4, 2>("tf32 N=64  A in TMEM 4 acc", 148);
  run<128, 1, 2>("tf32 N=128 A in TMEM", 148);
  run<128, 2, 2>("tf32 N=128 A in TMEM 2 acc", 148);
  run<128, 1, 1>("bf16 N=128 1 acc", 148);
  run<256, 1, 1>("bf16 N=256 1 acc", 148);
  run<256, 2, 1>("bf16 N=256 2 acc", 148);
  return 0;
}
