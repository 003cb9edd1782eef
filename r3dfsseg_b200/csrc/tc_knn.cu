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
#include <stdlib.h>

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
  static constexpr int XS_OFF = QI_OFF + 16 * 256;       // |x_j|^2 of 4 tiles in flight [4][128]
  static constexpr int TOTAL = XS_OFF + 4 * 128 * 4 + 64;  // (the queue is reused for the final merge)
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

// Candidate-tile loader split in two so that a thread's global loads of the NEXT tile are all in
// flight while the current tile is converted and the MMAs are issued (the straight loop above
// serialises one L2 round trip per 16-byte chunk, which is what used to bound the whole kernel).
template <int KC4, int ROWS, int NTHR>
struct TileRegs {
  static constexpr int NCH = ROWS * KC4 / NTHR;
  float4 v[NCH];
};

template <int KC4, int ROWS, int NTHR>
__device__ __forceinline__ void tile_load(TileRegs<KC4, ROWS, NTHR>& tr, const float* __restrict__ x,
                                          int ld, int C, int64_t row0, int64_t rows_end, int t,
                                          bool vec_ok) {
#pragma unroll
  for (int i = 0; i < TileRegs<KC4, ROWS, NTHR>::NCH; ++i) {
    const int c = t + i * NTHR;
    const int r = c / KC4, kc = c % KC4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t row = row0 + r;
    const int k = 4 * kc;
    if (row < rows_end && k < C) {
      const float* p = x + row * (int64_t)ld + k;
      if (vec_ok && k + 3 < C) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        v.x = __ldg(p);
        if (k + 1 < C) v.y = __ldg(p + 1);
        if (k + 2 < C) v.z = __ldg(p + 2);
        if (k + 3 < C) v.w = __ldg(p + 3);
      }
    }
    tr.v[i] = v;
  }
}

template <int KC4, int ROWS, int NTHR>
__device__ __forceinline__ void tile_store(const TileRegs<KC4, ROWS, NTHR>& tr,
                                           unsigned char* hi_base, unsigned char* lo_base, int t) {
  constexpr int LBO = tc::tile_lbo(ROWS);
#pragma unroll
  for (int i = 0; i < TileRegs<KC4, ROWS, NTHR>::NCH; ++i) {
    const int c = t + i * NTHR;
    const int r = c / KC4, kc = c % KC4;
    float4 hi, lo;
    tc::split4(tr.v[i], hi, lo);
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
    TileRegs<KC4, 128, 128> tr;
    float* xs = reinterpret_cast<float*>(smem + S::XS_OFF);
    tile_load(tr, x, ld, C, base, base + N, lt, vec_ok);
    float xn = lt < N ? __ldg(xx + base + lt) : 0.f;
    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      if (j >= 2) tc::mbar_wait(&bar_full[st], ((j >> 1) - 1) & 1);  // stage's previous MMAs done
      unsigned char* hi = smem + S::B_OFF + st * 2 * S::TILE;
      unsigned char* lo = hi + S::TILE;
      tile_store(tr, hi, lo, lt);
      // the tile's |x_j|^2 travel with it (ring of 4: slot j is rewritten by tile j + 4, whose
      // loaders have waited for MMA j + 2, which was issued after the selectors released tile j)
      xs[(j & 3) * KT_TC + lt] = xn;
      if (j + 1 < T) {
        tile_load(tr, x, ld, C, base + (int64_t)(j + 1) * KT_TC, base + N, lt, vec_ok);
        const int cn_ = (j + 1) * KT_TC + lt;
        xn = cn_ < N ? __ldg(xx + base + cn_) : 0.f;
      }
      tc::fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (lt == 0) {
        if (j >= 2) tc::mbar_wait(&bar_tfree[st], ((j >> 1) - 1) & 1);  // buffer drained
        tc::tc_fence_after();
        const uint32_t q_hi = tc::smem_u32(smem + S::Q_OFF), q_lo = q_hi + S::TILE;
        const uint32_t b_hi = tc::smem_u32(hi), b_lo = b_hi + S::TILE;
        const uint32_t d = tmem_d + st * KT_TC;
        // one descriptor per operand tile; k-step ks is a constant added to its address field
        const uint64_t dqh = tc::make_desc(q_hi, LBO, 128), dql = tc::make_desc(q_lo, LBO, 128);
        const uint64_t dbh = tc::make_desc(b_hi, LBO, 128), dbl = tc::make_desc(b_lo, LBO, 128);
        constexpr uint64_t KS = tc::desc_kstep(LBO);
        tc::mma_tf32_c<false>(d, dql, dbh, IDESC);
        tc::mma_tf32_c<true>(d, dqh, dbl, IDESC);
        tc::mma_tf32_c<true>(d, dqh, dbh, IDESC);
#pragma unroll
        for (int ks = 1; ks < 2 * KC4 / 4; ++ks) {
          if (ks < ksteps) {
            tc::mma_tf32_c<true>(d, dql + ks * KS, dbh + ks * KS, IDESC);
            tc::mma_tf32_c<true>(d, dqh + ks * KS, dbl + ks * KS, IDESC);
            tc::mma_tf32_c<true>(d, dqh + ks * KS, dbh + ks * KS, IDESC);
          }
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
    const float* xs = reinterpret_cast<const float*>(smem + S::XS_OFF);
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
            const float4 cn = *reinterpret_cast<const float4*>(xs + (j & 3) * KT_TC + cc + u);
            const float cnv[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // reference key: -xx_i - (-2 x_i.x_j) - xx_j
              const float key = fmaf(2.f, v[u + e], nq) - cnv[e];  // = (nq - (-2 v)) - cn exactly
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


// ---------------------------------------------------------------------------------------------
// Two-pass variant (k <= 20, 1024 <= N <= 65535): the per-row selection above pays one
// warp-divergent 120-instruction list insertion for each of the ~k (1 + ln(N / k)) candidates that
// beat a running threshold.  Here the distance GEMM runs TWICE (it is cheap: the tensor pipe idles
// otherwise) and the first pass only produces a tight, exact lower bound of every row's k-th best
// key:  the row's candidates are cut into 64 groups, the k-th largest of the 64 group maxima is
// <= the k-th largest key, and only ~k (1 + 1/3) keys lie above it.  The second pass recomputes the
// keys (bit-identical), appends the few that reach the bound to a per-thread queue in shared
// memory, and the 20-lists are built once at the end, all lanes inserting together.
//   candidate tiles are 64 wide here (two 33 KB operand stages instead of 66 KB) to make room for
//   the queues; warps 0-7 = (TMEM lane quarter, 32-column half), warps 8-11 loaders + MMA issue.
// ---------------------------------------------------------------------------------------------
#define K2_TC 64
#define K2_CAP 56  // queue entries per selector thread; drained when fewer than 32 slots remain
#define K2_G 32    // group maxima per selector thread (64 per row)
#define K2_THREADS 416  // warps 0-7 selectors, 8-11 loaders, 12 MMA issue

template <int KC4>
struct Knn2Smem {
  static constexpr int TQ = tc::tile_bytes(128, KC4);
  static constexpr int TB = tc::tile_bytes(K2_TC, KC4);
  static constexpr int Q_OFF = 0;                            // Q hi, Q lo
  static constexpr int B_OFF = 2 * TQ;                       // 2 stages x (hi, lo)
  static constexpr int QK_OFF = B_OFF + 4 * TB;              // queue keys [CAP][256] float
  static constexpr int QC_OFF = QK_OFF + K2_CAP * 256 * 4;   // queue columns [CAP][256] u16
  static constexpr int XS_OFF = QC_OFF + K2_CAP * 256 * 2;   // |x_j|^2 of 4 tiles in flight [4][64]
  static constexpr int TOTAL = XS_OFF + 4 * K2_TC * 4 + 64;
};

template <int KC4, int ROWS>
__device__ __forceinline__ void knn_store_tile_r(unsigned char* hi_base, unsigned char* lo_base,
                                                 const float* __restrict__ x, int ld, int C,
                                                 int64_t row0, int64_t rows_end, int t, int nthr,
                                                 bool vec_ok) {
  constexpr int LBO = tc::tile_lbo(ROWS);
  constexpr int CH = ROWS * KC4;
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

// descending bitonic sort of 64 registers (fully unrolled: every index is a compile-time constant)
__device__ __forceinline__ void sort64_desc(float (&v)[64]) {
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const float a = v[i], b = v[j];
          v[i] = desc ? fmaxf(a, b) : fminf(a, b);
          v[j] = desc ? fminf(a, b) : fmaxf(a, b);
        }
      }
    }
  }
}

// the reference's ranking keys of 32 consecutive candidates: -xx_i - (-2 x_i.x_j) - xx_j
__device__ __forceinline__ void knn_keys32(float (&v)[32], float nq, const float* xs32) {
#pragma unroll
  for (int u = 0; u < 32; u += 4) {
    const float4 cn = *reinterpret_cast<const float4*>(xs32 + u);
    // (nq - (-2 v)) - cn: the doubling is exact, so one FMA gives the reference's rounding
    v[u + 0] = fmaf(2.f, v[u + 0], nq) - cn.x;
    v[u + 1] = fmaf(2.f, v[u + 1], nq) - cn.y;
    v[u + 2] = fmaf(2.f, v[u + 2], nq) - cn.z;
    v[u + 3] = fmaf(2.f, v[u + 3], nq) - cn.w;
  }
}

template <int KC4>
__global__ __launch_bounds__(K2_THREADS, 1) void knn_tc2_kernel(const float* __restrict__ x, int ld,
                                                                int C, const float* __restrict__ xx,
                                                                int N, int k,
                                                                int32_t* __restrict__ idx32,
                                                                int64_t* __restrict__ idx64) {
  constexpr int KL = 20;
  extern __shared__ __align__(128) unsigned char smem[];
  using S = Knn2Smem<KC4>;
  __shared__ uint64_t bar_full[2];   // accumulator b ready = its MMAs done (also: operand stage b free)
  __shared__ uint64_t bar_tfree[2];  // accumulator b drained by the 256 selector threads
  __shared__ uint64_t bar_sfull[2];  // operand stage b written by the 128 loader threads
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * KT_TQ;
  const int64_t base = (int64_t)b * N;
  constexpr int LBOQ = tc::tile_lbo(128), LBOB = tc::tile_lbo(K2_TC);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, K2_TC);
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int T = (N + K2_TC - 1) / K2_TC;
  const int ksteps = (min(C, 4 * KC4) + 7) / 8;

  if (tid == 0) {
    tc::mbar_init(&bar_full[0], 1);
    tc::mbar_init(&bar_full[1], 1);
    tc::mbar_init(&bar_tfree[0], 256);
    tc::mbar_init(&bar_tfree[1], 256);
    tc::mbar_init(&bar_sfull[0], 128);
    tc::mbar_init(&bar_sfull[1], 128);
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 2 * K2_TC);
  knn_store_tile_r<KC4, 128>(smem + S::Q_OFF, smem + S::Q_OFF + S::TQ, x, ld, C, base + q0,
                             base + N, tid, K2_THREADS, vec_ok);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (w == 12) {
    // ------------------------------ MMA issue: its own warp ------------------------------------
    // (when a loader thread issued the MMAs, the loaders stood still for as long as the tensor
    // pipe took to accept the 24 instructions of a tile, and the next tile's conversion could not
    // overlap them: MMA and operand staging ran back to back instead of side by side)
    if (lane == 0) {
      const uint32_t q_hi = tc::smem_u32(smem + S::Q_OFF), q_lo = q_hi + S::TQ;
      const uint64_t dqh = tc::make_desc(q_hi, LBOQ, 128), dql = tc::make_desc(q_lo, LBOQ, 128);
      constexpr uint64_t KQ = tc::desc_kstep(LBOQ), KB = tc::desc_kstep(LBOB);
      for (int j = 0; j < 2 * T; ++j) {
        const int st = j & 1;
        tc::mbar_wait(&bar_sfull[st], (j >> 1) & 1);                      // operands staged
        if (j >= 2) tc::mbar_wait(&bar_tfree[st], ((j >> 1) - 1) & 1);    // accumulator drained
        tc::tc_fence_after();
        const uint32_t b_hi = tc::smem_u32(smem + S::B_OFF + st * 2 * S::TB), b_lo = b_hi + S::TB;
        const uint32_t d = tmem_d + st * K2_TC;
        const uint64_t dbh = tc::make_desc(b_hi, LBOB, 128), dbl = tc::make_desc(b_lo, LBOB, 128);
        tc::mma_tf32_c<false>(d, dql, dbh, IDESC);
        tc::mma_tf32_c<true>(d, dqh, dbl, IDESC);
        tc::mma_tf32_c<true>(d, dqh, dbh, IDESC);
#pragma unroll
        for (int ks = 1; ks < 2 * KC4 / 4; ++ks) {
          if (ks < ksteps) {
            tc::mma_tf32_c<true>(d, dql + ks * KQ, dbh + ks * KB, IDESC);
            tc::mma_tf32_c<true>(d, dqh + ks * KQ, dbl + ks * KB, IDESC);
            tc::mma_tf32_c<true>(d, dqh + ks * KQ, dbh + ks * KB, IDESC);
          }
        }
        tc::mma_commit(&bar_full[st]);
      }
    }
  } else if (w >= 8) {
    // ------------------------------ loaders: every tile twice ----------------------------------
    // Two tiles of global loads are in flight per thread (two register sets, prefetch distance 2):
    // with one, every tile paid a full L2 round trip between "registers free again" and "data
    // there" (ncu: the loaders' first use of the loaded data was the hottest loader instruction and
    // the tensor pipe idled 70 % of the time waiting for operands).
    const int lt = tid - 256;
    TileRegs<KC4, K2_TC, 128> trA, trB;
    float* xs = reinterpret_cast<float*>(smem + S::XS_OFF);
    auto tile_of = [&](int j) { return j < T ? j : j - T; };
    auto norm_of = [&](int j) {
      const int cn_ = tile_of(j) * K2_TC + lt;
      return (lt < K2_TC && cn_ < N) ? __ldg(xx + base + cn_) : 0.f;
    };
    tile_load(trA, x, ld, C, base, base + N, lt, vec_ok);
    float xnA = norm_of(0), xnB = 0.f;
    if (2 * T > 1) {
      tile_load(trB, x, ld, C, base + (int64_t)tile_of(1) * K2_TC, base + N, lt, vec_ok);
      xnB = norm_of(1);
    }
    for (int j = 0; j < 2 * T; j += 2) {  // 2 T is even
      {
        const int st = 0;
        if (j >= 2) tc::mbar_wait(&bar_full[st], ((j >> 1) - 1) & 1);  // stage's previous MMAs done
        unsigned char* hi = smem + S::B_OFF + st * 2 * S::TB;
        unsigned char* lo = hi + S::TB;
        tile_store(trA, hi, lo, lt);
        if (lt < K2_TC) xs[(j & 3) * K2_TC + lt] = xnA;  // ring of 4, see knn_tc_kernel
        if (j + 2 < 2 * T) {
          tile_load(trA, x, ld, C, base + (int64_t)tile_of(j + 2) * K2_TC, base + N, lt, vec_ok);
          xnA = norm_of(j + 2);
        }
        tc::fence_async_smem();
        mbar_arrive(&bar_sfull[st]);
      }
      {
        const int st = 1, j1 = j + 1;
        if (j1 >= 2) tc::mbar_wait(&bar_full[st], ((j1 >> 1) - 1) & 1);
        unsigned char* hi = smem + S::B_OFF + st * 2 * S::TB;
        unsigned char* lo = hi + S::TB;
        tile_store(trB, hi, lo, lt);
        if (lt < K2_TC) xs[(j1 & 3) * K2_TC + lt] = xnB;
        if (j1 + 2 < 2 * T) {
          tile_load(trB, x, ld, C, base + (int64_t)tile_of(j1 + 2) * K2_TC, base + N, lt, vec_ok);
          xnB = norm_of(j1 + 2);
        }
        tc::fence_async_smem();
        mbar_arrive(&bar_sfull[st]);
      }
    }
  } else {
    // ------------------------------ selectors ------------------------------------------------
    float* qk = reinterpret_cast<float*>(smem + S::QK_OFF);
    unsigned short* qc = reinterpret_cast<unsigned short*>(smem + S::QC_OFF);
    const int row = 32 * (w & 3) + lane;
    const int half = w >> 2;
    const int q = q0 + row;
    const float nq = (q < N) ? -xx[base + q] : 0.f;
    const float* xs = reinterpret_cast<const float*>(smem + S::XS_OFF) + 32 * half;
    const uint32_t taddr = tmem_d + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(32 * half);
    // ---- pass 1: group maxima --------------------------------------------------------------
#pragma unroll 1
    for (int s = 0; s < K2_G; ++s) qk[s * 256 + tid] = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      tc::mbar_wait(&bar_full[st], (j >> 1) & 1);
      tc::tc_fence_after();
      float v[32];
      tc::tmem_ld32(taddr + (uint32_t)(st * K2_TC), v);
      const int c0 = j * K2_TC + 32 * half;
      knn_keys32(v, nq, xs + (j & 3) * K2_TC);
      tc::tc_fence_before();
      mbar_arrive(&bar_tfree[st]);  // accumulator buffer and the tile's norms are consumed
      float m = -INFINITY;
      if (c0 + 32 <= N) {
#pragma unroll
        for (int e = 0; e < 32; ++e) m = fmaxf(m, v[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (c0 + e < N) m = fmaxf(m, v[e]);
      }
      float* g = qk + (j & (K2_G - 1)) * 256 + tid;
      *g = fmaxf(*g, m);
    }
    // ---- bound: k-th largest of the row's 64 group maxima ---------------------------------------
    asm volatile("bar.sync 2, 256;" ::: "memory");
    float tau;
    {
      float gm[64];
#pragma unroll
      for (int s = 0; s < K2_G; ++s) {
        gm[s] = qk[s * 256 + tid];
        gm[K2_G + s] = qk[s * 256 + (tid ^ 128)];
      }
      sort64_desc(gm);
      tau = gm[0];
#pragma unroll
      for (int i = 1; i < KL; ++i)
        if (i == k - 1) tau = gm[i];
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");  // the maxima are read: the region becomes the queue
    // ---- pass 2: queue everything that reaches the bound, build the lists lazily -------------------
    float lv[KL];
    int li[KL];
#pragma unroll
    for (int i = 0; i < KL; ++i) {
      lv[i] = -INFINITY;
      li[i] = 0;
    }
    int qcnt = 0;
    auto drain = [&]() {
      const int mx = __reduce_max_sync(0xffffffffu, qcnt);
      for (int e = 0; e < mx; ++e) {
        if (e < qcnt) {
          const float key = qk[e * 256 + tid];
          if (key > lv[KL - 1]) list_insert<KL>(lv, li, key, (int)qc[e * 256 + tid]);
        }
      }
      qcnt = 0;
    };
#pragma unroll 1
    for (int j2 = 0; j2 < T; ++j2) {
      const int j = T + j2;
      const int st = j & 1;
      tc::mbar_wait(&bar_full[st], (j >> 1) & 1);
      tc::tc_fence_after();
      float v[32];
      tc::tmem_ld32(taddr + (uint32_t)(st * K2_TC), v);
      const int c0 = j2 * K2_TC + 32 * half;
      knn_keys32(v, nq, xs + (j & 3) * K2_TC);
      tc::tc_fence_before();
      mbar_arrive(&bar_tfree[st]);
      // one comparison per key: v >= tau and v > thr  <=>  v >= max(tau, next float above thr)
      const float thr = lv[KL - 1];
      const int tb = __float_as_int(thr);  // next float above thr (-inf -> -FLT_MAX, -0 -> denorm min)
      const float cut = fmaxf(tau, __int_as_float(tb >= 0 ? tb + 1 : (tb == (int)0x80000000 ? 1 : tb - 1)));
      const int nvalid = N - c0;
      if (nvalid >= 32) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (v[e] >= cut) {
            qk[qcnt * 256 + tid] = v[e];
            qc[qcnt * 256 + tid] = (unsigned short)(c0 + e);
            ++qcnt;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (v[e] >= cut && e < nvalid) {
            qk[qcnt * 256 + tid] = v[e];
            qc[qcnt * 256 + tid] = (unsigned short)(c0 + e);
            ++qcnt;
          }
        }
      }
      if (__any_sync(0xffffffffu, qcnt > K2_CAP - 32)) drain();
    }
    drain();
    // merge the two column halves of every row (as in knn_tc_kernel)
    float* qv = qk;
    int* mi = reinterpret_cast<int*>(smem + S::B_OFF);
    asm volatile("bar.sync 2, 256;" ::: "memory");
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
  if (w == 0) tc::tmem_dealloc(tmem_d, 2 * K2_TC);
}

template <int KC4>
static int launch_knn_tc2_t(const float* x, int ld, int C, const float* xx, int64_t B, int N, int k,
                            int32_t* idx32, int64_t* idx64, cudaStream_t st) {
  using S = Knn2Smem<KC4>;
  cudaError_t e = cudaFuncSetAttribute(knn_tc2_kernel<KC4>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((N + KT_TQ - 1) / KT_TQ, (unsigned)B);
  knn_tc2_kernel<KC4><<<grid, K2_THREADS, S::TOTAL, st>>>(x, ld, C, xx, N, k, idx32, idx64);
  R3DFS_CHECK_LAUNCH();
  return 0;
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

static bool knn_single_pass_forced() {  // A/B switch: R3DFS_KNN_SINGLE_PASS=1
  static const bool v = [] {
    const char* e = R3DFS_GETENV("R3DFS_KNN_SINGLE_PASS");
    return e && e[0] == '1';
  }();
  return v;
}

// returns R3DFS_E_UNSUPPORTED for shapes the tensor-core kernel is not built for (C > 64)
int launch_knn_tc(const float* x, int ld, int C, const float* xx, int64_t B, int N, int k,
                  int32_t* idx32, int64_t* idx64, cudaStream_t st) {
  if (k < 1 || k > 32 || C > 64 || N < k) return R3DFS_E_UNSUPPORTED;
  // two-pass selection (measured per 300 clouds of 2048 points: C = 9: 2.03 vs 3.05 ms single
  // pass, C = 64: 3.20 vs 3.43 ms); R3DFS_KNN_SINGLE_PASS=1 switches it off (A/B measurements).
  if (k <= 20 && N >= 1024 && N <= 65535 && !knn_single_pass_forced()) {
    if (C <= 16) return launch_knn_tc2_t<4>(x, ld, C, xx, B, N, k, idx32, idx64, st);
    return launch_knn_tc2_t<16>(x, ld, C, xx, B, N, k, idx32, idx64, st);
  }
  if (C <= 16) {
    if (k <= 20) return launch_knn_tc_t<4, 20>(x, ld, C, xx, B, N, k, idx32, idx64, st);
    return launch_knn_tc_t<4, 32>(x, ld, C, xx, B, N, k, idx32, idx64, st);
  }
  if (k <= 20) return launch_knn_tc_t<16, 20>(x, ld, C, xx, B, N, k, idx32, idx64, st);
  return launch_knn_tc_t<16, 32>(x, ld, C, xx, B, N, k, idx32, idx64, st);
}
