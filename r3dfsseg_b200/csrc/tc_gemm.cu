// Tensor-core (tcgen05, TMEM accumulators) 3xTF32 GEMM with the layer epilogue:
//     Y[map(m)][n] = act(s[n] * sum_k X[m][k] W[n][k] + t[n])
// Same contract as the SIMT linear_kernel in encoder.cu.  One CTA = 128 rows x BN columns:
//   - all 8 warps stream X / W k-blocks from global memory, split every value into its TF32 high
//     and low parts and store them as K-major UMMA operand tiles in shared memory (2 stages);
//   - one thread issues, per 8-wide k-step, three tcgen05.mma (hi.hi + hi.lo + lo.hi) that
//     accumulate in TMEM; tcgen05.commit -> mbarrier releases the stage for the next loads;
//   - the epilogue reads the accumulator with tcgen05.ld (one row per thread), applies the folded
//     BatchNorm affine + activation and writes the rows.
#include "common.cuh"
#include "tc.cuh"

#define TCG_BM 128
#define TCG_BK 16             // floats per k-block (small stages -> 3 CTAs per SM hide the load latency)
#define TCG_KC4 (TCG_BK / 4)  // 16-byte chunks per k-block
#define TCG_THREADS 256

template <int BN>
struct TcgSmem {
  static constexpr int A_BYTES = tc::tile_bytes(TCG_BM, TCG_KC4);
  static constexpr int B_BYTES = tc::tile_bytes(BN, TCG_KC4);
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;  // A hi, A lo, B hi, B lo
  static constexpr int TOTAL = 2 * STAGE + 64;
};

// load one 128-bit chunk of row-major [rows][ld] (zero outside rows_end / k_end)
__device__ __forceinline__ float4 ld_chunk(const float* __restrict__ src, int64_t ld, int64_t row,
                                           int64_t rows_end, int k, int k_end, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows_end) {
    const float* p = src + row * ld + k;
    if (vec_ok && k + 3 < k_end) {
      v = *reinterpret_cast<const float4*>(p);
    } else {
      if (k + 0 < k_end) v.x = p[0];
      if (k + 1 < k_end) v.y = p[1];
      if (k + 2 < k_end) v.z = p[2];
      if (k + 3 < k_end) v.w = p[3];
    }
  }
  return v;
}

// DIST = true turns the epilogue into the squared-distance form used by the affinity graph:
//     Y[m][n] = (s[m] + t[n]) - 2 * acc      with s = t = squared row norms, X = W (square, BN = 128),
// and blockIdx.z walks over independent problems `zs` elements apart (X, W, s, t, Y alike).
template <int BN, bool DIST>
__global__ __launch_bounds__(TCG_THREADS, 3) void linear_tc_kernel(
    const float* __restrict__ X, int ldx, const float* __restrict__ W, const float* __restrict__ s,
    const float* __restrict__ t, int act, int64_t M, int K, int Nout, float* __restrict__ Y,
    int ldy, RowMap map, int64_t zs_x, int64_t zs_w, int64_t zs_v, int64_t zs_y) {
  extern __shared__ __align__(128) unsigned char smem[];
  // the distance matrix is symmetric: only the tiles on and above the diagonal are computed, an
  // off-diagonal tile is stored twice (as is, and transposed) — half the operand conversion and
  // tensor work for the same bytes written
  if (DIST && blockIdx.y < blockIdx.x) return;
  X += blockIdx.z * zs_x;
  W += blockIdx.z * zs_w;
  Y += blockIdx.z * zs_y;
  if (s) s += blockIdx.z * zs_v;
  if (t) t += blockIdx.z * zs_v;
  using S = TcgSmem<BN>;
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_sc[BN], s_sh[BN];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * TCG_BM;
  const int n0 = blockIdx.y * BN;
  if (!DIST && tid < BN) {
    const int n = n0 + tid;
    s_sc[tid] = (s && n < Nout) ? s[n] : 1.f;
    s_sh[tid] = (t && n < Nout) ? t[n] : 0.f;
  }
  constexpr int LBO_A = tc::tile_lbo(TCG_BM), LBO_B = tc::tile_lbo(BN);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(TCG_BM, BN);
  constexpr int A_CH = TCG_BM * TCG_KC4 / TCG_THREADS;  // chunks per thread per k-block (4)
  constexpr int B_CH = BN * TCG_KC4 / TCG_THREADS;      // 4 (BN=128) or 2 (BN=64)

  if (tid == 0) {
    tc::mbar_init(&bar_mma[0], 1);
    tc::mbar_init(&bar_mma[1], 1);
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  const bool vec_x = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const bool vec_w = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int KB = (K + TCG_BK - 1) / TCG_BK;
  for (int kb = 0; kb < KB; ++kb) {
    const int st = kb & 1;
    const int k0 = kb * TCG_BK;
    // global -> registers; a warp covers 4 rows x 8 chunks = 4 full 128 B lines
    float4 av[A_CH], bv[B_CH];
#pragma unroll
    for (int i = 0; i < A_CH; ++i) {
      const int c = tid + i * TCG_THREADS;
      const int r = c / TCG_KC4, kc = c % TCG_KC4;
      av[i] = ld_chunk(X, ldx, m0 + r, M, k0 + 4 * kc, K, vec_x);
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      const int c = tid + i * TCG_THREADS;
      const int r = c / TCG_KC4, kc = c % TCG_KC4;
      bv[i] = ld_chunk(W, K, n0 + r, Nout, k0 + 4 * kc, K, vec_w);
    }
    // the MMAs that read this stage two k-blocks ago must have finished
    if (kb >= 2) tc::mbar_wait(&bar_mma[st], ((kb >> 1) - 1) & 1);
    unsigned char* sA_hi = smem + st * S::STAGE;
    unsigned char* sA_lo = sA_hi + S::A_BYTES;
    unsigned char* sB_hi = sA_lo + S::A_BYTES;
    unsigned char* sB_lo = sB_hi + S::B_BYTES;
#pragma unroll
    for (int i = 0; i < A_CH; ++i) {
      const int c = tid + i * TCG_THREADS;
      const int r = c / TCG_KC4, kc = c % TCG_KC4;
      float4 hi, lo;
      tc::split4(av[i], hi, lo);
      *reinterpret_cast<float4*>(sA_hi + kc * LBO_A + r * 16) = hi;
      *reinterpret_cast<float4*>(sA_lo + kc * LBO_A + r * 16) = lo;
    }
#pragma unroll
    for (int i = 0; i < B_CH; ++i) {
      const int c = tid + i * TCG_THREADS;
      const int r = c / TCG_KC4, kc = c % TCG_KC4;
      float4 hi, lo;
      tc::split4(bv[i], hi, lo);
      *reinterpret_cast<float4*>(sB_hi + kc * LBO_B + r * 16) = hi;
      *reinterpret_cast<float4*>(sB_lo + kc * LBO_B + r * 16) = lo;
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const int ksteps = min(TCG_BK, K - k0 + 7) / 8;  // 8-wide k-steps holding real data
      const uint32_t a_hi = tc::smem_u32(sA_hi), a_lo = tc::smem_u32(sA_lo);
      const uint32_t b_hi = tc::smem_u32(sB_hi), b_lo = tc::smem_u32(sB_lo);
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t dah = tc::make_desc(a_hi + ks * 2 * LBO_A, LBO_A, 128);
        const uint64_t dal = tc::make_desc(a_lo + ks * 2 * LBO_A, LBO_A, 128);
        const uint64_t dbh = tc::make_desc(b_hi + ks * 2 * LBO_B, LBO_B, 128);
        const uint64_t dbl = tc::make_desc(b_lo + ks * 2 * LBO_B, LBO_B, 128);
        tc::mma_tf32(tmem_d, dal, dbh, IDESC, (kb | ks) != 0);
        tc::mma_tf32(tmem_d, dah, dbl, IDESC, 1);
        tc::mma_tf32(tmem_d, dah, dbh, IDESC, 1);
      }
      tc::mma_commit(&bar_mma[st]);
    }
  }
  // accumulator complete when the last commit has arrived
  tc::mbar_wait(&bar_mma[(KB - 1) & 1], ((KB - 1) >> 1) & 1);
  tc::tc_fence_after();

  // epilogue: warp w reads TMEM lanes 32*(w%4).., columns half (w/4)
  {
    // all MMAs are complete, so the operand stages are free: reuse them as per-warp transpose
    // buffers for coalesced stores
    float* wbuf = reinterpret_cast<float*>(smem) + w * (32 * 33);
    const int rbase = 32 * (w & 3);
    constexpr int HALF = BN / 2;
    const int cbase = (w >> 2) * HALF;
    const bool vec_y = ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
#pragma unroll
    for (int cc = 0; cc < HALF; cc += 32) {
      float v[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)rbase << 16) + (uint32_t)(cbase + cc), v);
      const int nb = n0 + cbase + cc;
      if (DIST) {
        const int64_t m = m0 + rbase + lane;
        const float sm = (m < M) ? s[m] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (sm + ((nb + j < Nout) ? t[nb + j] : 0.f)) - 2.f * v[j];
        if (blockIdx.y > blockIdx.x && m < M) {
          // mirrored tile: a lane holds one row, so for a fixed column the warp's 32 values are
          // consecutive in the transposed row — coalesced as they stand
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < Nout) Y[(int64_t)(nb + j) * ldy + m] = v[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = apply_act(fmaf(s_sc[cbase + cc + j], v[j], s_sh[cbase + cc + j]), act);
      }
      tc::store_chunk_coalesced(wbuf, v, lane, Nout - nb, vec_y, [&](int r) -> float* {
        const int64_t m = m0 + rbase + r;
        return (m < M) ? Y + map(m) * (int64_t)ldy + nb : nullptr;
      });
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem_d, BN);
}

template <int BN>
static int launch_linear_tc_bn(const float* X, int ldx, const float* W, const float* s,
                               const float* t, int act, int64_t M, int K, int Nout, float* Y,
                               int ldy, RowMap map, cudaStream_t st) {
  using S = TcgSmem<BN>;
  cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<BN, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((unsigned)((M + TCG_BM - 1) / TCG_BM), (Nout + BN - 1) / BN);
  linear_tc_kernel<BN, false><<<grid, TCG_THREADS, S::TOTAL, st>>>(X, ldx, W, s, t, act, M, K, Nout,
                                                                   Y, ldy, map, 0, 0, 0, 0);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// D2[g][i][j] = |f_i|^2 + |f_j|^2 - 2 f_i.f_j for G graphs of nn nodes (rows of D floats)
int launch_gram_dist_tc(const float* F, int64_t graph_rows, int64_t row_off, int G, int nn, int D,
                        const float* norms, float* D2, cudaStream_t st) {
  using S = TcgSmem<128>;
  cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<128, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  if (e != cudaSuccess) return (int)e;
  const float* F0 = F + row_off * D;
  dim3 grid((nn + TCG_BM - 1) / TCG_BM, (nn + 127) / 128, G);
  linear_tc_kernel<128, true><<<grid, TCG_THREADS, S::TOTAL, st>>>(
      F0, D, F0, norms, norms, 0, nn, D, nn, D2, nn, identity_map(), graph_rows * D,
      graph_rows * D, nn, (int64_t)nn * nn);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

int launch_linear_tc(const float* X, int ldx, const float* W, const float* s, const float* t,
                     int act, int64_t M, int K, int Nout, float* Y, int ldy, RowMap map,
                     cudaStream_t st) {
  if (Nout % 128 == 0 || Nout > 128)
    return launch_linear_tc_bn<128>(X, ldx, W, s, t, act, M, K, Nout, Y, ldy, map, st);
  return launch_linear_tc_bn<64>(X, ldx, W, s, t, act, M, K, Nout, Y, ldy, map, st);
}


// ---------------------------------------------------------------------------------------------
// "TN" product for the weight gradients of the training path:  C[m][n] = sum_r A[r][m] * B[r][n]
// with A (R x M) and B (R x N) row-major — the reduction runs over the ROWS (R = points or edges,
// 24 576 ... 409 600), so R is cut into `splits` ranges (blockIdx.z), each CTA writes its partial
// 128 x 128 tile to partial[split][M][N] and a fixed-order reduce kernel adds them (deterministic).
// Same 3xTF32 tcgen05 pipeline as linear_tc_kernel; the only difference is the loader, which builds
// each 16-byte K-major chunk from four consecutive rows of one column (a warp reads 128 contiguous
// bytes of a row at a time).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_chunk_t(const float* __restrict__ src, int64_t ld, int col,
                                             int ncols, int64_t r, int64_t r_end) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < ncols) {
    const float* p = src + r * ld + col;
    if (r + 0 < r_end) v.x = __ldg(p);
    if (r + 1 < r_end) v.y = __ldg(p + ld);
    if (r + 2 < r_end) v.z = __ldg(p + 2 * ld);
    if (r + 3 < r_end) v.w = __ldg(p + 3 * ld);
  }
  return v;
}

__global__ __launch_bounds__(TCG_THREADS, 3) void gemm_tn_tc_kernel(
    const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int M,
    int N, int64_t R, int64_t rchunk, float* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem[];
  using S = TcgSmem<128>;
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int m0 = blockIdx.x * TCG_BM, n0 = blockIdx.y * 128;
  const int64_t r_begin = (int64_t)blockIdx.z * rchunk, r_end = min(R, r_begin + rchunk);
  constexpr int LBO = tc::tile_lbo(128);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, 128);
  constexpr int CH = 128 * TCG_KC4 / TCG_THREADS;  // 2 chunks of each operand per thread
  if (tid == 0) {
    tc::mbar_init(&bar_mma[0], 1);
    tc::mbar_init(&bar_mma[1], 1);
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 128);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const int KB = (int)((r_end - r_begin + TCG_BK - 1) / TCG_BK);
  for (int kb = 0; kb < KB; ++kb) {
    const int st = kb & 1;
    const int64_t r0 = r_begin + (int64_t)kb * TCG_BK;
    float4 av[CH], bv[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = tid + i * TCG_THREADS;  // chunk (kc, col): col fastest -> coalesced row reads
      const int col = c & 127, kc = c >> 7;
      av[i] = ld_chunk_t(A, lda, m0 + col, M, r0 + 4 * kc, r_end);
      bv[i] = ld_chunk_t(B, ldb, n0 + col, N, r0 + 4 * kc, r_end);
    }
    if (kb >= 2) tc::mbar_wait(&bar_mma[st], ((kb >> 1) - 1) & 1);
    unsigned char* sA_hi = smem + st * S::STAGE;
    unsigned char* sA_lo = sA_hi + S::A_BYTES;
    unsigned char* sB_hi = sA_lo + S::A_BYTES;
    unsigned char* sB_lo = sB_hi + S::B_BYTES;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = tid + i * TCG_THREADS;
      const int col = c & 127, kc = c >> 7;
      float4 hi, lo;
      tc::split4(av[i], hi, lo);
      *reinterpret_cast<float4*>(sA_hi + kc * LBO + col * 16) = hi;
      *reinterpret_cast<float4*>(sA_lo + kc * LBO + col * 16) = lo;
      tc::split4(bv[i], hi, lo);
      *reinterpret_cast<float4*>(sB_hi + kc * LBO + col * 16) = hi;
      *reinterpret_cast<float4*>(sB_lo + kc * LBO + col * 16) = lo;
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint64_t dah = tc::make_desc(tc::smem_u32(sA_hi), LBO, 128);
      const uint64_t dal = tc::make_desc(tc::smem_u32(sA_lo), LBO, 128);
      const uint64_t dbh = tc::make_desc(tc::smem_u32(sB_hi), LBO, 128);
      const uint64_t dbl = tc::make_desc(tc::smem_u32(sB_lo), LBO, 128);
      constexpr uint64_t KS = tc::desc_kstep(LBO);
      if (kb == 0) tc::mma_tf32_c<false>(tmem_d, dal, dbh, IDESC);
      else tc::mma_tf32_c<true>(tmem_d, dal, dbh, IDESC);
      tc::mma_tf32_c<true>(tmem_d, dah, dbl, IDESC);
      tc::mma_tf32_c<true>(tmem_d, dah, dbh, IDESC);
      tc::mma_tf32_c<true>(tmem_d, dal + KS, dbh + KS, IDESC);  // rows beyond r_end are zeros
      tc::mma_tf32_c<true>(tmem_d, dah + KS, dbl + KS, IDESC);
      tc::mma_tf32_c<true>(tmem_d, dah + KS, dbh + KS, IDESC);
      tc::mma_commit(&bar_mma[st]);
    }
  }
  if (KB > 0) tc::mbar_wait(&bar_mma[(KB - 1) & 1], ((KB - 1) >> 1) & 1);
  tc::tc_fence_after();
  {
    float* out = partial + (int64_t)blockIdx.z * M * N;
    const int rbase = 32 * (w & 3), cbase = (w >> 2) * 64;
    const int m = m0 + rbase + lane;
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      float v[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)rbase << 16) + (uint32_t)(cbase + cc), v);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + cbase + cc + j;
          if (n < N) out[(int64_t)m * N + n] = KB > 0 ? v[j] : 0.f;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem_d, 128);
}

// partial: splits * M * N floats; returns the number of splits used through *splits_out
int launch_gemm_tn_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N,
                      int64_t R, int max_splits, float* partial, int* splits_out, cudaStream_t st) {
  using S = TcgSmem<128>;
  if (M <= 0 || N <= 0 || R <= 0 || max_splits < 1) return R3DFS_E_BADARG;
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  int64_t splits = (3 * 148 * 2 + tiles - 1) / tiles;  // ~2 waves of 3 CTAs per SM
  const int64_t by_r = (R + 511) / 512;
  if (splits > by_r) splits = by_r;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rchunk = (R + splits - 1) / splits;
  rchunk = (rchunk + TCG_BK - 1) / TCG_BK * TCG_BK;
  splits = (R + rchunk - 1) / rchunk;
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((M + 127) / 128, (N + 127) / 128, (unsigned)splits);
  gemm_tn_tc_kernel<<<grid, TCG_THREADS, S::TOTAL, st>>>(A, lda, B, ldb, M, N, R, rchunk, partial);
  R3DFS_CHECK_LAUNCH();
  *splits_out = (int)splits;
  return 0;
}
