// Hard assignment of set points to their FPS seeds (reference models/mpti.py:618-622):
//     argmin_j || f - seed_j + 1e-6 ||_2      (torch<=1.8 pairwise_distance, first minimum)
// as "tensor-core filter + exact verify":
//   1. score_j = |s_j|^2 - 2 f.s_j - 2e-6 * sum(s_j)  — the j-dependent part of the squared norm —
//      from a 3xTF32 tcgen05 GEMM (128 points x 128 seed slots per CTA, K = D);
//   2. a point whose best score beats every other seed by more than a safety margin (far above the
//      Gram-vs-direct rounding difference) is assigned right away; the rare ambiguous point is
//      re-evaluated over ALL seeds with the reference's direct FP32 arithmetic
//      (sum (f - s + 1e-6)^2, sqrt, strict <), so the result equals the direct evaluation.
#include "common.cuh"
#include "proto.cuh"
#include "tc.cuh"

#define AS_THREADS 256
#define AS_BK 16
#define AS_KC4 (AS_BK / 4)

struct AssignSmem {
  static constexpr int TILE = tc::tile_bytes(128, AS_KC4);  // hi or lo
  static constexpr int STAGE = 4 * TILE;                    // A hi, A lo, B hi, B lo
  static constexpr int TOTAL = 2 * STAGE + 64;
};

// per seed: squared norm and element sum (one warp per seed)
__global__ void seed_stats_kernel(const float* __restrict__ feat, int D,
                                  const int32_t* __restrict__ set_off,
                                  const int32_t* __restrict__ seeds,
                                  const int32_t* __restrict__ proto_cnt, int m_max,
                                  float* __restrict__ sn, float* __restrict__ ss) {
  const int set = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int m = proto_cnt[set];
  for (int j = w; j < 128; j += nw) {
    float a = 0.f, b = 0.f;
    if (j < m) {
      const int sidx = seeds[(int64_t)set * m_max + j];
      if (sidx >= 0) {
        const float* p = feat + ((int64_t)set_off[set] + sidx) * D;
        for (int d = lane; d < D; d += 32) {
          const float v = p[d];
          a = fmaf(v, v, a);
          b += v;
        }
      }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      sn[set * 128 + j] = a;
      ss[set * 128 + j] = b;
    }
  }
}

__global__ __launch_bounds__(AS_THREADS, 3) void assign_tc_kernel(
    const float* __restrict__ feat, int D, const int32_t* __restrict__ set_off,
    const int32_t* __restrict__ set_n, const int32_t* __restrict__ seeds,
    const int32_t* __restrict__ proto_cnt, const float* __restrict__ sn,
    const float* __restrict__ ss, int m_max, int k, int32_t* __restrict__ assign) {
  extern __shared__ __align__(128) unsigned char smem[];
  using S = AssignSmem;
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_seed[128];
  const int set = blockIdx.y;
  const int n = set_n[set];
  const int p0 = blockIdx.x * 128;
  if (p0 >= n) return;
  const int64_t row0 = set_off[set];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (n <= k) {  // every point is its own prototype (models/mpti.py:631-634)
    for (int p = p0 + tid; p < min(n, p0 + 128); p += AS_THREADS) assign[row0 + p] = p;
    return;
  }
  const int m = proto_cnt[set];
  constexpr int LBO = tc::tile_lbo(128);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, 128);
  if (tid < 128) s_seed[tid] = tid < m ? seeds[(int64_t)set * m_max + tid] : -1;
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

  const int KB = (D + AS_BK - 1) / AS_BK;
  for (int kb = 0; kb < KB; ++kb) {
    const int st = kb & 1;
    const int k0 = kb * AS_BK;
    float4 av[2], bv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * AS_THREADS;
      const int r = c / AS_KC4, kc = c % AS_KC4;
      const int kk = k0 + 4 * kc;
      av[i] = (p0 + r < n && kk < D)
                  ? *reinterpret_cast<const float4*>(feat + (row0 + p0 + r) * (int64_t)D + kk)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      const int sidx = s_seed[r];
      bv[i] = (sidx >= 0 && kk < D)
                  ? *reinterpret_cast<const float4*>(feat + (row0 + sidx) * (int64_t)D + kk)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (kb >= 2) tc::mbar_wait(&bar_mma[st], ((kb >> 1) - 1) & 1);
    unsigned char* a_hi = smem + st * S::STAGE;
    unsigned char* a_lo = a_hi + S::TILE;
    unsigned char* b_hi = a_lo + S::TILE;
    unsigned char* b_lo = b_hi + S::TILE;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * AS_THREADS;
      const int r = c / AS_KC4, kc = c % AS_KC4;
      float4 hi, lo;
      tc::split4(av[i], hi, lo);
      *reinterpret_cast<float4*>(a_hi + kc * LBO + r * 16) = hi;
      *reinterpret_cast<float4*>(a_lo + kc * LBO + r * 16) = lo;
      tc::split4(bv[i], hi, lo);
      *reinterpret_cast<float4*>(b_hi + kc * LBO + r * 16) = hi;
      *reinterpret_cast<float4*>(b_lo + kc * LBO + r * 16) = lo;
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo);
      const uint32_t bh = tc::smem_u32(b_hi), bl = tc::smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < AS_BK / 8; ++ks) {
        const uint64_t dah = tc::make_desc(ah + ks * 2 * LBO, LBO, 128);
        const uint64_t dal = tc::make_desc(al + ks * 2 * LBO, LBO, 128);
        const uint64_t dbh = tc::make_desc(bh + ks * 2 * LBO, LBO, 128);
        const uint64_t dbl = tc::make_desc(bl + ks * 2 * LBO, LBO, 128);
        tc::mma_tf32(tmem_d, dal, dbh, IDESC, (kb | ks) != 0);
        tc::mma_tf32(tmem_d, dah, dbl, IDESC, 1);
        tc::mma_tf32(tmem_d, dah, dbh, IDESC, 1);
      }
      tc::mma_commit(&bar_mma[st]);
    }
  }
  tc::mbar_wait(&bar_mma[(KB - 1) & 1], ((KB - 1) >> 1) & 1);
  tc::tc_fence_after();

  // All MMAs are complete (the last commit covers every earlier one), so the operand stages are
  // free: they become the 128 x 128 matrix of approximate scores, element j of row r at
  // r * 128 + ((j + r) & 127) (the rotation keeps a warp's column-wise accesses conflict-free).
  __syncthreads();
  float* s_score = reinterpret_cast<float*>(smem);
  {
    const int row = 32 * (w & 3) + lane;
    const int half = w >> 2;
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      float v[32];
      tc::tmem_ld32(tmem_d + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(64 * half + cc), v);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = 64 * half + cc + e;
        const float sc = j < m ? fmaf(-2.f, v[e], sn[set * 128 + j]) - 2e-6f * ss[set * 128 + j]
                               : INFINITY;
        s_score[row * 128 + ((j + row) & 127)] = sc;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem_d, 128);
  // thread = point: the best approximate score; every seed within a safety margin of it (far above
  // the Gram-vs-direct rounding difference: 3xTF32 products + FP32 accumulation of 192 terms stay
  // below 1.2e-5 of |f||s|) is re-evaluated with the reference's direct FP32 arithmetic
  // (sum (f - s + 1e-6)^2, sqrt, strict <, seeds in index order), so the result equals the direct
  // evaluation over all seeds.  Usually there is exactly one candidate and nothing is re-evaluated.
  if (tid < 128) {
    const int p = p0 + tid;
    if (p < n) {
      const float* sr = s_score + tid * 128;
      float bb = INFINITY;
      int ba = 0;
      for (int j = 0; j < m; ++j) {
        const float sc = sr[(j + tid) & 127];
        if (sc < bb) {
          bb = sc;
          ba = j;
        }
      }
      const float margin = 5e-5f * fmaxf(1.f, sn[set * 128 + ba] + fabsf(bb));
      int n_cand = 0;
      for (int j = 0; j < m; ++j) n_cand += sr[(j + tid) & 127] <= bb + margin;
      int result = ba;
      if (n_cand > 1) {
        const float4* f4 = reinterpret_cast<const float4*>(feat + (row0 + p) * (int64_t)D);
        float bestd = INFINITY;
        for (int j = 0; j < m; ++j) {
          if (!(sr[(j + tid) & 127] <= bb + margin)) continue;
          const float4* s4 = reinterpret_cast<const float4*>(feat + (row0 + s_seed[j]) * (int64_t)D);
          float acc = 0.f;
          for (int d4 = 0; d4 < (D >> 2); ++d4) {  // same order as the scalar loop: d ascending
            const float4 a = f4[d4], b = s4[d4];
            float dv = __fadd_rn(a.x - b.x, 1e-6f);
            acc = fmaf(dv, dv, acc);
            dv = __fadd_rn(a.y - b.y, 1e-6f);
            acc = fmaf(dv, dv, acc);
            dv = __fadd_rn(a.z - b.z, 1e-6f);
            acc = fmaf(dv, dv, acc);
            dv = __fadd_rn(a.w - b.w, 1e-6f);
            acc = fmaf(dv, dv, acc);
          }
          const float dn = sqrtf(acc);
          if (dn < bestd) {
            bestd = dn;
            result = j;
          }
        }
      }
      assign[row0 + p] = result;
    }
  }
}

int launch_assign_tc(const float* feat, int D, const int32_t* set_off, const int32_t* set_n,
                     const int32_t* seeds, const int32_t* proto_cnt, int n_sets, int n_cap,
                     int m_max, int k, float* sn, float* ss, int32_t* assign, cudaStream_t st) {
  if ((D & 3) != 0 || m_max > 128) return R3DFS_E_UNSUPPORTED;
  seed_stats_kernel<<<n_sets, 256, 0, st>>>(feat, D, set_off, seeds, proto_cnt, m_max, sn, ss);
  R3DFS_CHECK_LAUNCH();
  cudaError_t e = cudaFuncSetAttribute(assign_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AssignSmem::TOTAL);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((n_cap + 127) / 128, n_sets);
  assign_tc_kernel<<<grid, AS_THREADS, AssignSmem::TOTAL, st>>>(feat, D, set_off, set_n, seeds,
                                                              proto_cnt, sn, ss, m_max, k, assign);
  R3DFS_CHECK_LAUNCH();
  return 0;
}
