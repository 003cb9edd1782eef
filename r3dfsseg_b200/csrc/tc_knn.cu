// knn on the tensor cores: the pairwise-distance GEMM (3xTF32, tcgen05.mma, accumulators in TMEM)
// feeds a per-row top-k selection that reads the accumulator straight out of TMEM — the (N, N)
// matrix never exists in shared or global memory  (reference models/dgcnn.py:17-23).
//
// One CTA = 128 query points of one cloud, warp-specialised:
//   warps 4-7  load 128-candidate tiles (global -> TF32 hi/lo split -> K-major UMMA tiles in shared
//              memory, 2 stages); one of their threads issues the MMAs into one of two TMEM
//              accumulator buffers and commits to an mbarrier;
//   warps 0-3  own one TMEM lane (= query row) per thread: tcgen05.ld 32 columns at a time, form
//              the reference's ranking key, queue the candidates that beat the row's current k-th
//              best, then merge the queue into a sorted k-list kept in registers.
// While the selectors work on tile j the tensor core computes tile j+1 and the loaders fetch j+2.
#include "common.cuh"
#include "tc.cuh"

#define KT_TQ 128
#define KT_TC 128
#define KT_THREADS 384  // warps 0-7 selectors (2 per TMEM lane quarter), warps 8-11 loaders

template <int KC4>
struct KnnTcSmem {
  static constexpr int TILE = tc::tile_bytes(128, KC4);  // one of hi / lo
  static constexpr int Q_OFF = 0;                        // Q hi, Q lo
  static constexpr int B_OFF = 2 * TILE;                 // 2 stages x (hi, lo)
  static constexpr int QV_OFF = B_OFF + 4 * TILE;        // queue values  [16][256] float
  static constexpr int QI_OFF = QV_OFF + 16 * 256 * 4;   // queue columns [16][256] uint8
  static constexpr int TOTAL = QI_OFF + 16 * 256 + 64;   // (the queue is reused for the final merge)
};

template <int KC4>
__device__ __forceinline__ void knn_store_tile(unsigned char* hi_base, unsigned char* lo_base,
                                               const float* __restrict__ x, int ld, int C,
                                               int64_t row0, int64_t rows_end, int t, int nthr,
                                               bool vec_ok) {
  constexpr int LBO = tc::tile_lbo(128);
  constexpr int CH = 128 * KC4;
  // a warp covers 32/KC4' rows x chunks; chunk index fastest so global reads are contiguous
  for (int c = t; c < CH; c += nthr) {
    const int r = c / KC4, kc = c % KC4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t row = row0 + r;
    const int k = 4 * kc;
    if (row < rows_end && k < C) {
      const float* p = x + row * (int64_t)ld + k;
      if (vec_ok && k + 3 < C) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        v.x = p[0];
        if (k + 1 < C) v.y = p[1];
        if (k + 2 < C) v.z = p[2];
        if (k + 3 < C) v.w = p[3];
      }
    }
    float4 hi, lo;
    tc::split4(v, hi, lo);
    *reinterpret_cast<float4*>(hi_base + kc * LBO + r * 16) = hi;
    *reinterpret_cast<float4*>(lo_base + kc * LBO + r * 16) = lo;
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// sorted (descending) insertion into a register-resident k-list; c[i] = key > lv[i]
template <int KL>
__device__ __forceinline__ void list_insert(float (&lv)[KL], int (&li)[KL], float key, int idx) {
#pragma unroll
  for (int i = KL - 1; i > 0; --i) {
    const bool up = key > lv[i - 1];  // everything from i-1 on moves down one slot
    const bool here = key > lv[i];
    const float nv = up ? lv[i - 1] : (here ? key : lv[i]);
    const int ni = up ? li[i - 1] : (here ? idx : li[i]);
    lv[i] = nv;
    li[i] = ni;
  }
  if (key > lv[0]) {
    lv[0] = key;
    li[0] = idx;
  }
}

template <int KC4, int KL>
__global__ __launch_bounds__(KT_THREADS, 1) void knn_tc_kernel(const float* __restrict__ x, int ld,
                                                               int C, const float* __restrict__ xx,
                                                               int N, int k,
                                                               int32_t* __restrict__ idx32,
                                                               int64_t* __restrict__ idx64) {
  extern __shared__ __align__(128) unsigned char smem[];
  using S = KnnTcSmem<KC4>;
  __shared__ uint64_t bar_full[2];   // accumulator buffer b ready (= MMAs of its tile complete)
  __shared__ uint64_t bar_tfree[2];  // accumulator buffer b drained by the 128 selector threads
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * KT_TQ;
  const int64_t base = (int64_t)b * N;
  constexpr int LBO = tc::tile_lbo(128);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, KT_TC);
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int T = (N + KT_TC - 1) / KT_TC;
  const int ksteps = (min(C, 4 * KC4) + 7) / 8;

  if (tid == 0) {
    tc::mbar_init(&bar_full[0], 1);
    tc::mbar_init(&bar_full[1], 1);
    tc::mbar_init(&bar_tfree[0], 256);
    tc::mbar_init(&bar_tfree[1], 256);
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 2 * KT_TC);
  // query tile (all threads)
  knn_store_tile<KC4>(smem + S::Q_OFF, smem + S::Q_OFF + S::TILE, x, ld, C, base + q0, base + N,
                      tid, KT_THREADS, vec_ok);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (w >= 8) {
    // ------------------------------ loaders + MMA issue -------------------------------------
    const int lt = tid - 256;
    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      if (j >= 2) tc::mbar_wait(&bar_full[st], ((j >> 1) - 1) & 1);  // stage's previous MMAs done
      unsigned char* hi = smem + S::B_OFF + st * 2 * S::TILE;
      unsigned char* lo = hi + S::TILE;
      knn_store_tile<KC4>(hi, lo, x, ld, C, base + (int64_t)j * KT_TC, base + N, lt, 128, vec_ok);
      tc::fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (lt == 0) {
        if (j >= 2) tc::mbar_wait(&bar_tfree[st], ((j >> 1) - 1) & 1);  // buffer drained
        tc::tc_fence_after();
        const uint32_t q_hi = tc::smem_u32(smem + S::Q_OFF), q_lo = q_hi + S::TILE;
        const uint32_t b_hi = tc::smem_u32(hi), b_lo = b_hi + S::TILE;
        const uint32_t d = tmem_d + st * KT_TC;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t dqh = tc::make_desc(q_hi + ks * 2 * LBO, LBO, 128);
          const uint64_t dql = tc::make_desc(q_lo + ks * 2 * LBO, LBO, 128);
          const uint64_t dbh = tc::make_desc(b_hi + ks * 2 * LBO, LBO, 128);
          const uint64_t dbl = tc::make_desc(b_lo + ks * 2 * LBO, LBO, 128);
          tc::mma_tf32(d, dql, dbh, IDESC, ks != 0);
          tc::mma_tf32(d, dqh, dbl, IDESC, 1);
          tc::mma_tf32(d, dqh, dbh, IDESC, 1);
        }
        tc::mma_commit(&bar_full[st]);
      }
    }
  } else {
    // ------------------------------ selectors ------------------------------------------------
    // thread = (query row, column half): warps w and w+4 share TMEM lane quarter w%4; warp w < 4
    // scans columns 0..63 of every tile, warp w >= 4 columns 64..127; the two partial k-lists of a
    // row are merged through shared memory at the end.
    float* qv = reinterpret_cast<float*>(smem + S::QV_OFF);
    unsigned char* qi = smem + S::QI_OFF;
    const int row = 32 * (w & 3) + lane;
    const int half = w >> 2;
    const int q = q0 + row;
    const float nq = (q < N) ? -xx[base + q] : 0.f;
    const bool vecn = (N & 3) == 0;
    float lv[KL];
    int li[KL];
#pragma unroll
    for (int i = 0; i < KL; ++i) {
      lv[i] = -INFINITY;
      li[i] = 0;
    }
    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      tc::mbar_wait(&bar_full[st], (j >> 1) & 1);
      tc::tc_fence_after();
      const int c0 = j * KT_TC;
#pragma unroll 1
      for (int cc = 64 * half; cc < 64 * half + 64; cc += 32) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(st * KT_TC + cc), v);
        // 16 columns at a time: queue what beats the current k-th best, then merge the queue
#pragma unroll
        for (int hh = 0; hh < 32; hh += 16) {
          const float thr = lv[KL - 1];
          int cnt = 0;
#pragma unroll
          for (int u = hh; u < hh + 16; u += 4) {
            const int cg = c0 + cc + u;
            float4 cn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cg + 3 < N && vecn) {
              cn = __ldg(reinterpret_cast<const float4*>(xx + base + cg));
            } else {
              if (cg + 0 < N) cn.x = xx[base + cg + 0];
              if (cg + 1 < N) cn.y = xx[base + cg + 1];
              if (cg + 2 < N) cn.z = xx[base + cg + 2];
              if (cg + 3 < N) cn.w = xx[base + cg + 3];
            }
            const float cnv[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // reference key: -xx_i - (-2 x_i.x_j) - xx_j
              const float inner = -2.f * v[u + e];
              const float key = (nq - inner) - cnv[e];
              if (key > thr && cg + e < N) {
                qv[cnt * 256 + tid] = key;
                qi[cnt * 256 + tid] = (unsigned char)(cc + u + e);
                ++cnt;
              }
            }
          }
          const int mx = __reduce_max_sync(0xffffffffu, cnt);
          for (int e = 0; e < mx; ++e) {
            if (e < cnt) {
              const float key = qv[e * 256 + tid];
              if (key > lv[KL - 1]) list_insert<KL>(lv, li, key, c0 + (int)qi[e * 256 + tid]);
            }
          }
        }
      }
      tc::tc_fence_before();
      mbar_arrive(&bar_tfree[st]);
    }
    // merge: the upper-half warps publish their lists, the lower-half warps absorb them
    int* mi = reinterpret_cast<int*>(smem + S::B_OFF);  // loaders are done with the B stages by now
    asm volatile("bar.sync 2, 256;" ::: "memory");      // every selector has consumed its last tile
    if (half == 1) {
#pragma unroll
      for (int i = 0; i < KL; ++i) {
        qv[i * 128 + row] = lv[i];
        mi[i * 128 + row] = li[i];
      }
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (half == 0) {
#pragma unroll 1
      for (int i = 0; i < KL; ++i) {
        const float key = qv[i * 128 + row];
        if (key > lv[KL - 1]) list_insert<KL>(lv, li, key, mi[i * 128 + row]);
      }
      if (q < N) {
#pragma unroll
        for (int i = 0; i < KL; ++i) {
          if (i < k) {
            const int64_t o = (base + q) * k + i;
            if (idx32) idx32[o] = li[i];
            if (idx64) idx64[o] = li[i];
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem_d, 2 * KT_TC);
}

template <int KC4, int KL>
static int launch_knn_tc_t(const float* x, int ld, int C, const float* xx, int64_t B, int N, int k,
                           int32_t* idx32, int64_t* idx64, cudaStream_t st) {
  using S = KnnTcSmem<KC4>;
  cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<KC4, KL>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((N + KT_TQ - 1) / KT_TQ, (unsigned)B);
  knn_tc_kernel<KC4, KL><<<grid, KT_THREADS, S::TOTAL, st>>>(x, ld, C, xx, N, k, idx32, idx64);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// returns R3DFS_E_UNSUPPORTED for shapes the tensor-core kernel is not built for (C > 64)
int launch_knn_tc(const float* x, int ld, int C, const float* xx, int64_t B, int N, int k,
                  int32_t* idx32, int64_t* idx64, cudaStream_t st) {
  if (k < 1 || k > 32 || C > 64 || N < k) return R3DFS_E_UNSUPPORTED;
  if (C <= 16) {
    if (k <= 20) return launch_knn_tc_t<4, 20>(x, ld, C, xx, B, N, k, idx32, idx64, st);
    return launch_knn_tc_t<4, 32>(x, ld, C, xx, B, N, k, idx32, idx64, st);
  }
  if (k <= 20) return launch_knn_tc_t<16, 20>(x, ld, C, xx, B, N, k, idx32, idx64, st);
  return launch_knn_tc_t<16, 32>(x, ld, C, xx, B, N, k, idx32, idx64, st);
}
