// SelfAttention (eval) on the tensor cores  (reference models/attention.py:39-48):
//     y = softmax((q / 8)^T k) v ,  single head, d = 64, per cloud of N points.
// One CTA = 128 query points; thread = (query row, 16-column quarter) for everything read from TMEM.
// Two sweeps over the key tiles (64 keys each), both with 3xTF32 tcgen05.mma:
//   sweep 1:  S = Q K^T  -> running row maximum m            (no exponentials)
//   sweep 2:  S again (bit-identical), P = exp(S - m) written as a K-major UMMA A tile (hi/lo),
//             l += rowsum(P),  O += P V  with V^T staged as the K-major B tile; O stays in TMEM.
// Knowing m before the second sweep means O never has to be rescaled inside TMEM.
// Softmax is shift invariant, so sweep 1 only has to deliver SOME m' >= max_j S_ij that keeps
// exp(S - m') in the normal FP32 range:
//   - if b = |q_i / 8| * max_j |k_j| (Cauchy-Schwarz; the per-cloud key-norm maximum comes from a
//     small pre-kernel) is <= 43 for every row of the CTA, m' = b and sweep 1 is skipped: every
//     S_ij lies in [-b, b], so exp(S - b) >= exp(-2 b) >= exp(-86) = 4.5e-38 stays a normal FP32
//     number even for a row whose keys are all anti-aligned with its query (the row sum cannot
//     underflow);
//   - otherwise sweep 1 runs with ONE TF32 product per k-step (Qhi.Khi, a third of the MMAs, no lo
//     tiles) and m' = its row maximum + 2^-10 b, which bounds what the dropped products can add.
// The (N, N) attention map exists only as 128 x 64 tiles in TMEM / shared memory.
#include "common.cuh"
#include "tc.cuh"

#define AT_THREADS 512  // 16 worker warps: TMEM lane quarter = w % 4, 16-column quarter = w / 4
#define AT_ALL_THREADS 544  // + warp 16: MMA issue
#define AT_BQ 128
#define AT_BK 64

struct AttTcSmem {
  static constexpr int Q_TILE = tc::tile_bytes(128, 16);  // hi or lo, 128 rows x 64 (d)
  static constexpr int K_TILE = tc::tile_bytes(64, 16);   // 64 keys x 64 (d)
  static constexpr int V_TILE = tc::tile_bytes(64, 16);   // 64 (d) rows x 64 keys  (V^T)
  static constexpr int P_TILE = tc::tile_bytes(128, 16);  // 128 rows x 64 keys
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_OFF + 2 * Q_TILE;
  static constexpr int V_OFF = K_OFF + 2 * K_TILE;
  static constexpr int P_OFF = V_OFF + 2 * V_TILE;
  static constexpr int X_OFF = P_OFF + 2 * P_TILE;  // exchange: 4 x 128 floats
  static constexpr int TOTAL = X_OFF + 4 * 128 * 4 + 64;
};

// 64 rows x 64 columns of `src` (row stride ld, starting column col_off) -> K-major hi/lo tiles
__device__ __forceinline__ void at_load_rows(float4 (&v)[2], const float* __restrict__ src, int ld,
                                             int64_t row0, int64_t rows_end, int col_off, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * AT_THREADS;
    const int r = c >> 4, kc = c & 15;
    v[i] = (row0 + r < rows_end)
               ? *reinterpret_cast<const float4*>(src + (row0 + r) * (int64_t)ld + col_off + 4 * kc)
               : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void at_store_rows(const float4 (&v)[2], unsigned char* hi,
                                              unsigned char* lo, int tid) {
  constexpr int LBO = tc::tile_lbo(64);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * AT_THREADS;
    const int r = c >> 4, kc = c & 15;
    float4 h, l;
    tc::split4(v[i], h, l);
    *reinterpret_cast<float4*>(hi + kc * LBO + r * 16) = h;
    *reinterpret_cast<float4*>(lo + kc * LBO + r * 16) = l;
  }
}
// TF32 hi parts only (the approximate first sweep does not need the lo tile)
__device__ __forceinline__ void at_store_rows_hi(const float4 (&v)[2], unsigned char* hi, int tid) {
  constexpr int LBO = tc::tile_lbo(64);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * AT_THREADS;
    const int r = c >> 4, kc = c & 15;
    float4 h;
    h.x = tc::tf32_rn(v[i].x); h.y = tc::tf32_rn(v[i].y);
    h.z = tc::tf32_rn(v[i].z); h.w = tc::tf32_rn(v[i].w);
    *reinterpret_cast<float4*>(hi + kc * LBO + r * 16) = h;
  }
}
// V rows (keys) -> V^T tile: element (d, key) at (key/4)*LBO + d*16 + (key%4)*4
__device__ __forceinline__ void at_load_v(float4 (&v)[2], const float* __restrict__ src, int ld,
                                          int64_t row0, int64_t rows_end, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * AT_THREADS;
    const int key = c & 63, d4 = c >> 6;
    v[i] = (row0 + key < rows_end)
               ? *reinterpret_cast<const float4*>(src + (row0 + key) * (int64_t)ld + 128 + 4 * d4)
               : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void at_store_vT(const float4 (&v)[2], unsigned char* hi,
                                            unsigned char* lo, int tid) {
  constexpr int LBO = tc::tile_lbo(64);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * AT_THREADS;
    const int key = c & 63, d4 = c >> 6;
    float4 h, l;
    tc::split4(v[i], h, l);
    const int base = (key >> 2) * LBO + (key & 3) * 4 + (4 * d4) * 16;
    *reinterpret_cast<float*>(hi + base) = h.x;
    *reinterpret_cast<float*>(hi + base + 16) = h.y;
    *reinterpret_cast<float*>(hi + base + 32) = h.z;
    *reinterpret_cast<float*>(hi + base + 48) = h.w;
    *reinterpret_cast<float*>(lo + base) = l.x;
    *reinterpret_cast<float*>(lo + base + 16) = l.y;
    *reinterpret_cast<float*>(lo + base + 32) = l.z;
    *reinterpret_cast<float*>(lo + base + 48) = l.w;
  }
}

// max_j |k_j|^2 per cloud (k = columns 64..127 of the qkv rows)
__global__ __launch_bounds__(256) void att_kmax_kernel(const float* __restrict__ qkv, int ld, int N,
                                                       float* __restrict__ kmax2) {
  __shared__ float s_red[8];
  const int b = blockIdx.x;
  float mx = 0.f;
  for (int r = threadIdx.x; r < N; r += 256) {
    const float4* p = reinterpret_cast<const float4*>(qkv + ((int64_t)b * N + r) * ld + 64);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 v = __ldg(p + c);
      s = fmaf(v.x, v.x, s);
      s = fmaf(v.y, v.y, s);
      s = fmaf(v.z, v.z, s);
      s = fmaf(v.w, v.w, s);
    }
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) mx = fmaxf(mx, s_red[q]);
    kmax2[b] = mx;
  }
}

__global__ __launch_bounds__(AT_ALL_THREADS, 1) void attention_tc_kernel(const float* __restrict__ qkv,
                                                                     int ld, int N,
                                                                     float* __restrict__ Y, int ldy,
                                                                     RowMap map,
                                                                     const float* __restrict__ kmax2) {
  extern __shared__ __align__(128) unsigned char smem[];
  using S = AttTcSmem;
  __shared__ uint64_t bar_s, bar_pv;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * AT_BQ;
  const int64_t base = (int64_t)b * N;
  constexpr int LBO_Q = tc::tile_lbo(128), LBO_K = tc::tile_lbo(64);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, 64);
  const int T = (N + AT_BK - 1) / AT_BK;
  const int row = 32 * (w & 3) + lane;  // TMEM lane = query row
  const int quarter = w >> 2;           // columns [16*quarter, 16*quarter+16) of every 64-wide tile
  float* xch = reinterpret_cast<float*>(smem + S::X_OFF);

  if (tid == 0) {
    tc::mbar_init(&bar_s, 1);
    tc::mbar_init(&bar_pv, 1);
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 256);  // S0 | S1 | O | (unused)
  // Q tile (pre-divided by the temperature 8 = sqrt(64), exactly as the reference does)
  for (int c = tid; c < 128 * 16; c += AT_ALL_THREADS) {
    const int r = c >> 4, kc = c & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < N) v = *reinterpret_cast<const float4*>(qkv + (base + q0 + r) * (int64_t)ld + 4 * kc);
    v.x /= 8.f; v.y /= 8.f; v.z /= 8.f; v.w /= 8.f;
    float4 h, l;
    tc::split4(v, h, l);
    *reinterpret_cast<float4*>(smem + S::Q_OFF + kc * LBO_Q + r * 16) = h;
    *reinterpret_cast<float4*>(smem + S::Q_OFF + S::Q_TILE + kc * LBO_Q + r * 16) = l;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = tmem_base_s, tmem_o = tmem_base_s + 128;
  const bool worker = w < 16;             // warps 0-15 move data; warp 16 only issues MMAs
  const bool issuer = w == 16;  // the whole warp walks the issue code, one elected lane issues
  const uint32_t q_hi = tc::smem_u32(smem + S::Q_OFF), q_lo = q_hi + S::Q_TILE;
  unsigned char* k_hi_p = smem + S::K_OFF;
  unsigned char* k_lo_p = k_hi_p + S::K_TILE;
  unsigned char* v_hi_p = smem + S::V_OFF;
  unsigned char* v_lo_p = v_hi_p + S::V_TILE;
  unsigned char* p_hi_p = smem + S::P_OFF;
  unsigned char* p_lo_p = p_hi_p + S::P_TILE;
  const uint32_t k_hi = tc::smem_u32(k_hi_p), k_lo = tc::smem_u32(k_lo_p);
  const uint32_t v_hi = tc::smem_u32(v_hi_p), v_lo = tc::smem_u32(v_lo_p);
  const uint32_t p_hi = tc::smem_u32(p_hi_p), p_lo = tc::smem_u32(p_lo_p);

  auto issue_s = [&](uint32_t tmem_s) {  // S[128 x 64] = Q . K^T  (K = 64 -> 8 k-steps)
#pragma unroll 1
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t dqh = tc::make_desc(q_hi + ks * 2 * LBO_Q, LBO_Q, 128);
      const uint64_t dql = tc::make_desc(q_lo + ks * 2 * LBO_Q, LBO_Q, 128);
      const uint64_t dkh = tc::make_desc(k_hi + ks * 2 * LBO_K, LBO_K, 128);
      const uint64_t dkl = tc::make_desc(k_lo + ks * 2 * LBO_K, LBO_K, 128);
      tc::mma_tf32_elect(tmem_s, dql, dkh, IDESC, ks != 0);
      tc::mma_tf32_elect(tmem_s, dqh, dkl, IDESC, 1);
      tc::mma_tf32_elect(tmem_s, dqh, dkh, IDESC, 1);
    }
    tc::mma_commit_elect(&bar_s);
  };

  auto issue_s_hi = [&]() {  // S ~= Qhi . Khi^T only (1 of the 3 TF32 products): row-max estimate
    const uint64_t dqh = tc::make_desc(q_hi, LBO_Q, 128), dkh = tc::make_desc(k_hi, LBO_K, 128);
    constexpr uint64_t KQ = tc::desc_kstep(LBO_Q), KK = tc::desc_kstep(LBO_K);
    tc::mma_tf32_elect(tmem_s, dqh, dkh, IDESC, 0);
#pragma unroll
    for (int ks = 1; ks < 8; ++ks) tc::mma_tf32_elect(tmem_s, dqh + ks * KQ, dkh + ks * KK, IDESC, 1);
    tc::mma_commit_elect(&bar_s);
  };

  uint32_t ph_s = 0, ph_pv = 0;
  // ---------------- row bound m' = |q/8| max|k| (or exact row maxima when the bound is too loose) ---
  float m_row = 0.f, margin = 0.f;
  bool sweep1 = true;              // run a first sweep for the row maxima?
  const bool approx = kmax2 != nullptr;  // ... with single-TF32 MMAs (needs the norm bound)
  if (approx) {
    float qs = 0.f;
    if (worker && q0 + row < N) {
      const float4* qp =
          reinterpret_cast<const float4*>(qkv + (base + q0 + row) * (int64_t)ld + 16 * quarter);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 v = __ldg(qp + c);
        v.x /= 8.f; v.y /= 8.f; v.z /= 8.f; v.w /= 8.f;
        qs = fmaf(v.x, v.x, qs);
        qs = fmaf(v.y, v.y, qs);
        qs = fmaf(v.z, v.z, qs);
        qs = fmaf(v.w, v.w, qs);
      }
    }
    if (worker) xch[quarter * 128 + row] = qs;
    __syncthreads();
    const int xr = worker ? row : 0;
    const float q2 = (xch[xr] + xch[128 + xr]) + (xch[256 + xr] + xch[384 + xr]);
    // |q_i / 8| max_j |k_j| >= max_j S_ij (Cauchy-Schwarz); 1e-4 relative + 1e-6 absolute slack
    // covers the rounding of the norms and of the 3xTF32 S
    m_row = sqrtf(q2) * sqrtf(__ldg(kmax2 + b)) * 1.0001f + 1e-6f;
    // Qhi.Khi drops q_lo.k + q_hi.k_lo: |error| <= 2 * 2^-11 |q||k|  (TF32 rounding is 2^-11 relative)
    margin = m_row * (1.0f / 1024.0f) * 1.01f + 1e-6f;
    sweep1 = __syncthreads_or(!(m_row <= 43.f));
  }
  if (sweep1) {
    // ---------------- sweep 1: row maxima (approximate + safety margin, or exact) ----------------
    float m_run = -INFINITY;
    {
      float4 kv[2];
      if (worker) at_load_rows(kv, qkv, ld, base, base + N, 64, tid);
      for (int j = 0; j < T; ++j) {
        if (worker) {
          if (approx) at_store_rows_hi(kv, k_hi_p, tid);
          else at_store_rows(kv, k_hi_p, k_lo_p, tid);
        }
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncthreads();  // K tile complete; everybody has finished reading S of the previous tile
        if (issuer) {
          tc::tc_fence_after();
          if (approx) issue_s_hi();
          else issue_s(tmem_s);
        }
        if (worker && j + 1 < T)
          at_load_rows(kv, qkv, ld, base + (int64_t)(j + 1) * AT_BK, base + N, 64, tid);
        tc::mbar_wait(&bar_s, ph_s);
        ph_s ^= 1;
        tc::tc_fence_after();
        if (worker) {
          float v[16];
          tc::tmem_ld16(tmem_s + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(16 * quarter), v);
          const int c0 = j * AT_BK + 16 * quarter;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c0 + e < N) m_run = fmaxf(m_run, v[e]);
        }
      }
    }
    // combine the four column quarters of every row
    __syncthreads();
    if (worker) xch[quarter * 128 + row] = m_run;
    tc::tc_fence_before();
    __syncthreads();
    const int xr = worker ? row : 0;
    m_row = fmaxf(fmaxf(xch[xr], xch[128 + xr]), fmaxf(xch[256 + xr], xch[384 + xr])) + margin;
  }
  __syncthreads();

  // ---------------- sweep 2: P = exp(S - m), l, O += P V ---------------------------------------
  // Software pipeline over the key tiles: S lives in two TMEM buffers, so S(j+1) = Q K(j+1)^T is
  // issued as soon as S(j) has landed (which also frees the K tile) and runs on the tensor pipe
  // while the workers turn S(j) into P(j); P(j).V(j) is queued behind it.  The tensor pipe then
  // always has the next MMA batch waiting: per tile it is busy for both GEMMs back to back.
  float l_run = 0.f;
  {
    float4 kv[2], vv[2];
    if (worker) {
      at_load_rows(kv, qkv, ld, base, base + N, 64, tid);
      at_load_v(vv, qkv, ld, base, base + N, tid);
      at_store_rows(kv, k_hi_p, k_lo_p, tid);
      if (T > 1) at_load_rows(kv, qkv, ld, base + AT_BK, base + N, 64, tid);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (issuer) {
      tc::tc_fence_after();
      issue_s(tmem_s);
    }
    for (int j = 0; j < T; ++j) {
      const uint32_t s_cur = tmem_s + (uint32_t)((j & 1) * 64);
      tc::mbar_wait(&bar_s, ph_s);  // S(j) complete: its K tile is free
      ph_s ^= 1;
      tc::tc_fence_after();
      if (worker && j + 1 < T) at_store_rows(kv, k_hi_p, k_lo_p, tid);  // K(j+1)
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (issuer && j + 1 < T) {
        tc::tc_fence_after();
        issue_s(tmem_s + (uint32_t)(((j + 1) & 1) * 64));
      }
      float4 pv4[4];
      if (worker) {
        if (j + 2 < T)
          at_load_rows(kv, qkv, ld, base + (int64_t)(j + 2) * AT_BK, base + N, 64, tid);
        float v[16];
        tc::tmem_ld16(s_cur + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(16 * quarter), v);
        const int c0 = j * AT_BK + 16 * quarter;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float4 p;
          p.x = (c0 + e + 0 < N) ? expf(v[e + 0] - m_row) : 0.f;
          p.y = (c0 + e + 1 < N) ? expf(v[e + 1] - m_row) : 0.f;
          p.z = (c0 + e + 2 < N) ? expf(v[e + 2] - m_row) : 0.f;
          p.w = (c0 + e + 3 < N) ? expf(v[e + 3] - m_row) : 0.f;
          l_run += (p.x + p.y) + (p.z + p.w);
          pv4[e >> 2] = p;
        }
      }
      // V^T and P buffers are free once the previous tile's P.V MMAs have completed
      if (j > 0) {
        tc::mbar_wait(&bar_pv, ph_pv);
        ph_pv ^= 1;
      }
      if (worker) {
        at_store_vT(vv, v_hi_p, v_lo_p, tid);
        if (j + 1 < T) at_load_v(vv, qkv, ld, base + (int64_t)(j + 1) * AT_BK, base + N, tid);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float4 h, l;
          tc::split4(pv4[e >> 2], h, l);
          const int kc = (16 * quarter + e) >> 2;  // 16-byte chunk along the key dimension
          *reinterpret_cast<float4*>(p_hi_p + kc * LBO_Q + row * 16) = h;
          *reinterpret_cast<float4*>(p_lo_p + kc * LBO_Q + row * 16) = l;
        }
      }
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();  // P and V^T complete, S(j) consumed
      if (issuer) {
        tc::tc_fence_after();
        const uint64_t dph = tc::make_desc(p_hi, LBO_Q, 128), dpl = tc::make_desc(p_lo, LBO_Q, 128);
        const uint64_t dvh = tc::make_desc(v_hi, LBO_K, 128), dvl = tc::make_desc(v_lo, LBO_K, 128);
        constexpr uint64_t KP = tc::desc_kstep(LBO_Q), KV = tc::desc_kstep(LBO_K);
        // O[128 x 64] += P[128 x 64 keys] . V[64 keys x 64]
        tc::mma_tf32_elect(tmem_o, dpl, dvh, IDESC, j != 0);
        tc::mma_tf32_elect(tmem_o, dph, dvl, IDESC, 1);
        tc::mma_tf32_elect(tmem_o, dph, dvh, IDESC, 1);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks) {
          tc::mma_tf32_elect(tmem_o, dpl + ks * KP, dvh + ks * KV, IDESC, 1);
          tc::mma_tf32_elect(tmem_o, dph + ks * KP, dvl + ks * KV, IDESC, 1);
          tc::mma_tf32_elect(tmem_o, dph + ks * KP, dvh + ks * KV, IDESC, 1);
        }
        tc::mma_commit_elect(&bar_pv);
      }
    }
  }
  // ---------------- epilogue: O / l ------------------------------------------------------------
  if (worker) xch[quarter * 128 + row] = l_run;
  tc::mbar_wait(&bar_pv, ph_pv);
  tc::tc_fence_after();
  __syncthreads();
  if (worker) {
    const float inv = 1.f / ((xch[row] + xch[128 + row]) + (xch[256 + row] + xch[384 + row]));
    float v[16];
    tc::tmem_ld16(tmem_o + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(16 * quarter), v);
    const int q = q0 + row;
    if (q < N) {
      float* y = Y + map(base + q) * (int64_t)ldy + 16 * quarter;
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        *reinterpret_cast<float4*>(y + e) =
            make_float4(v[e] * inv, v[e + 1] * inv, v[e + 2] * inv, v[e + 3] * inv);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem_base_s, 256);
}

// ---------------------------------------------------------------------------------------------
// TMA-fed variant, operands in TMEM  (what the per-tile trace of the kNN kernel showed applies
// here unchanged: an MMA whose A and B both come from shared memory saturates the SM's 128 B/clk
// at N <= 128, an N = 64 instruction never runs below ~45 cycles, the issuing thread is
// synchronous, and the 16 worker warps of attention_tc_kernel spend most of a tile fetching,
// splitting and storing K, V^T and P behind three CTA-wide barriers).
//   att_split_kernel  splits K and V^T of every cloud ONCE into 128-key UMMA tiles in global
//                     memory (every query block of a cloud used to redo it, twice);
//   attention_tc2_kernel, one CTA = 128 queries:
//     Q/8 (hi, lo) lives in TMEM and is the A operand of S = Q K^T (128 x 128 x 8 instructions,
//     full rate, B tiles arrive by cp.async.bulk);  the workers (thread = row x 16-column
//     quarter of a 64-key half) read S from TMEM, form P = exp(S - m), and write P's hi part over
//     the S columns it came from and its lo part next to O — P never touches shared memory and is
//     the A operand of O += P V;  S is double-buffered in TMEM so S(j+1) runs while P(j) is
//     formed;  one producer lane, one elected MMA lane, mbarriers instead of CTA-wide barriers.
//   TMEM (512 columns): Q hi 0-63 | Q lo 64-127 | S0 / P0 hi 128-255 | S1 / P1 hi 256-383 |
//                       O 384-447 | P lo (one 64-key half) 448-511.
//   Same split arithmetic, same k-step and key order as attention_tc_kernel: bit-identical output.
// ---------------------------------------------------------------------------------------------
// per-tile timestamps of one CTA (scripts/microbench/att_trace.cu compiles this file with ATT_TRACE)
#ifdef ATT_TRACE
__device__ long long g_att_trace[8 * 256];
#define ATT_TR(slot, j) \
  do { if (blockIdx.x == ATT_TRACE_BX && blockIdx.y == ATT_TRACE_BY) g_att_trace[(slot) * 256 + (j)] = clock64(); } while (0)
#else
#define ATT_TR(slot, j) do { } while (0)
#endif
#define A2_THREADS 576  // warps 0-15 workers, 16 copy producer, 17 MMA issue
struct Att2 {
  static constexpr int K_TILE = tc::tile_bytes(128, 16);  // 128 keys x 64 (d), hi or lo
  static constexpr int V_HALF = tc::tile_bytes(64, 16);   // 64 (d) rows x 64 keys (V^T), hi or lo
  static constexpr int BLK = 2 * K_TILE + 4 * V_HALF;     // K hi, K lo, V0 hi, V0 lo, V1 hi, V1 lo
  static constexpr int K_OFF = 0;                         // 2 stages x (hi, lo)
  static constexpr int V_OFF = K_OFF + 4 * K_TILE;        // 2 half buffers x (hi, lo)
  static constexpr int X_OFF = V_OFF + 4 * V_HALF;        // exchange: 4 x 128 floats
  static constexpr int TOTAL = X_OFF + 4 * 128 * 4 + 64;
};

__global__ __launch_bounds__(256) void att_split_kernel(const float* __restrict__ qkv, int ld, int N,
                                                        unsigned char* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int LBO_K = tc::tile_lbo(128), LBO_V = tc::tile_lbo(64);
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int T = gridDim.x;
  const int64_t base = (int64_t)b * N, row0 = (int64_t)t * 128;
  float4* dst = reinterpret_cast<float4*>(out + ((int64_t)b * T + t) * Att2::BLK);
  // ---- K tile: row r, chunk kc -> kc * LBO + r * 16
  for (int c = tid; c < 128 * 16; c += 256) {
    const int r = c >> 4, kc = c & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N)
      v = *reinterpret_cast<const float4*>(qkv + (base + row0 + r) * (int64_t)ld + 64 + 4 * kc);
    float4 h, l;
    tc::split4(v, h, l);
    *reinterpret_cast<float4*>(sm + kc * LBO_K + r * 16) = h;
    *reinterpret_cast<float4*>(sm + Att2::K_TILE + kc * LBO_K + r * 16) = l;
  }
  if (tid < 32)
    *reinterpret_cast<float4*>(sm + (tid >> 4) * Att2::K_TILE + (tid & 15) * LBO_K + 128 * 16) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < 2 * Att2::K_TILE / 16; i += 256) dst[i] = reinterpret_cast<float4*>(sm)[i];
  __syncthreads();
  // ---- V^T halves: element (d, key) of half h -> (key / 4) * LBO + d * 16 + (key % 4) * 4
  for (int c = tid; c < 128 * 16; c += 256) {
    const int key = c & 127, d4 = c >> 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + key < N)
      v = *reinterpret_cast<const float4*>(qkv + (base + row0 + key) * (int64_t)ld + 128 + 4 * d4);
    float4 h, l;
    tc::split4(v, h, l);
    const int half = key >> 6, kk = key & 63;
    unsigned char* hi = sm + half * 2 * Att2::V_HALF;
    unsigned char* lo = hi + Att2::V_HALF;
    const int o = (kk >> 2) * LBO_V + (kk & 3) * 4 + (4 * d4) * 16;
    *reinterpret_cast<float*>(hi + o) = h.x;
    *reinterpret_cast<float*>(hi + o + 16) = h.y;
    *reinterpret_cast<float*>(hi + o + 32) = h.z;
    *reinterpret_cast<float*>(hi + o + 48) = h.w;
    *reinterpret_cast<float*>(lo + o) = l.x;
    *reinterpret_cast<float*>(lo + o + 16) = l.y;
    *reinterpret_cast<float*>(lo + o + 32) = l.z;
    *reinterpret_cast<float*>(lo + o + 48) = l.w;
  }
  if (tid < 64)
    *reinterpret_cast<float4*>(sm + (tid >> 4) * Att2::V_HALF + (tid & 15) * LBO_V + 64 * 16) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  dst += 2 * Att2::K_TILE / 16;
  for (int i = tid; i < 4 * Att2::V_HALF / 16; i += 256) dst[i] = reinterpret_cast<float4*>(sm)[i];
}

__device__ __forceinline__ void a2_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ __launch_bounds__(A2_THREADS, 1) void attention_tc2_kernel(
    const float* __restrict__ qkv, int ld, int N, const unsigned char* __restrict__ split,
    float* __restrict__ Y, int ldy, RowMap map, const float* __restrict__ kmax2) {
  extern __shared__ __align__(128) unsigned char smem[];
  using S = Att2;
  __shared__ uint64_t bar_kfull[2], bar_kfree[2];    // K stage landed / read by its S MMAs
  __shared__ uint64_t bar_vfull[2], bar_vfree[2];    // V^T half buffer landed / read by its PV MMAs
  __shared__ uint64_t bar_sfull[2], bar_sfree[2];    // S buffer complete / read by the workers (sweep 1)
  __shared__ uint64_t bar_pfull[2], bar_pvdone[2];   // P half written / its PV MMAs complete
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  const int64_t base = (int64_t)b * N;
  constexpr int LBO_K = tc::tile_lbo(128), LBO_V = tc::tile_lbo(64);
  constexpr uint32_t IDESC_S = tc::make_idesc_tf32(128, 128), IDESC_O = tc::make_idesc_tf32(128, 64);
  constexpr uint32_t COL_QH = 0, COL_QL = 64, COL_S = 128, COL_O = 384, COL_PL = 448;
  const int T = (N + 127) / 128;
  const bool worker = w < 16;
  const int row = 32 * (w & 3) + lane;  // TMEM lane = query row
  const int cq = (w >> 2) & 3;          // 16-column quarter of every 64-key half
  float* xch = reinterpret_cast<float*>(smem + S::X_OFF);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bar_kfull[i], 1);
      tc::mbar_init(&bar_kfree[i], 1);
      tc::mbar_init(&bar_vfull[i], 1);
      tc::mbar_init(&bar_vfree[i], 1);
      tc::mbar_init(&bar_sfull[i], 1);
      tc::mbar_init(&bar_sfree[i], 512);
      tc::mbar_init(&bar_pfull[i], 512);
      tc::mbar_init(&bar_pvdone[i], 1);
    }
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(32 * (w & 3)) << 16);

  // ---- Q/8 -> TMEM (exact division, as the reference), |q/8|^2 for the row bound ----------------
  float qs = 0.f;
  if (worker) {
    float hi[16], lo[16];
    const bool ok = q0 + row < N;
    const float4* qp = reinterpret_cast<const float4*>(qkv + (base + q0 + row) * (int64_t)ld + 16 * cq);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 v = ok ? __ldg(qp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x /= 8.f; v.y /= 8.f; v.z /= 8.f; v.w /= 8.f;
      qs = fmaf(v.x, v.x, qs);
      qs = fmaf(v.y, v.y, qs);
      qs = fmaf(v.z, v.z, qs);
      qs = fmaf(v.w, v.w, qs);
      float4 h, l;
      tc::split4(v, h, l);
      hi[4 * c] = h.x; hi[4 * c + 1] = h.y; hi[4 * c + 2] = h.z; hi[4 * c + 3] = h.w;
      lo[4 * c] = l.x; lo[4 * c + 1] = l.y; lo[4 * c + 2] = l.z; lo[4 * c + 3] = l.w;
    }
    tc::tmem_st16(trow + COL_QH + 16 * cq, hi);
    tc::tmem_st16(trow + COL_QL + 16 * cq, lo);
    tc::tmem_st_wait();
    xch[cq * 128 + row] = qs;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  // ---- row bound m' = |q/8| max|k| (see attention_tc_kernel) ---------------------------------------
  float m_row = 0.f, margin = 0.f;
  bool sweep1 = true;
  const bool approx = kmax2 != nullptr;
  if (approx) {
    const int xr = worker ? row : 0;
    const float q2 = (xch[xr] + xch[128 + xr]) + (xch[256 + xr] + xch[384 + xr]);
    m_row = sqrtf(q2) * sqrtf(__ldg(kmax2 + b)) * 1.0001f + 1e-6f;
    margin = m_row * (1.0f / 1024.0f) * 1.01f + 1e-6f;
    sweep1 = __syncthreads_or(!(m_row <= 43.f));
  }
  const int T1 = sweep1 ? T : 0;  // tiles of the first sweep; global tile index g = T1 + j in sweep 2
  const unsigned char* src0 = split + (int64_t)b * T * S::BLK;

  if (w == 16) {
    // ------------------------------ copy producer ---------------------------------------------
    // K runs one tile ahead of the V^T halves: a V half buffer is free only when the previous
    // tile's PV MMAs are done, and waiting for that must not delay the next K tile
    if (lane == 0) {
      const int G = T1 + T;
      auto load_k = [&](int g) {
        const int kb = g & 1;
        if (g >= 2) tc::mbar_wait(&bar_kfree[kb], ((g >> 1) - 1) & 1);
        const bool s1 = g < T1;
        const int t = s1 ? g : g - T1;
        const uint32_t kbytes = (s1 && approx) ? S::K_TILE : 2 * S::K_TILE;  // hi only: approximate sweep
        tc::mbar_arrive_expect_tx(&bar_kfull[kb], kbytes);
        tc::bulk_g2s(smem + S::K_OFF + kb * 2 * S::K_TILE, src0 + (int64_t)t * S::BLK, kbytes,
                     &bar_kfull[kb]);
      };
      if (G > 0) load_k(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) load_k(g + 1);
        if (g >= T1) {
          const int t = g - T1;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            if (t >= 1) tc::mbar_wait(&bar_vfree[h], (t - 1) & 1);
            tc::mbar_arrive_expect_tx(&bar_vfull[h], 2 * S::V_HALF);
            tc::bulk_g2s(smem + S::V_OFF + h * 2 * S::V_HALF,
                         src0 + (int64_t)t * S::BLK + 2 * S::K_TILE + h * 2 * S::V_HALF,
                         2 * S::V_HALF, &bar_vfull[h]);
          }
        }
      }
    }
  } else if (w == 17) {
    // ------------------------------ MMA issue (whole warp, elected lane) ----------------------
    constexpr uint64_t KK = tc::desc_kstep(LBO_K), KV = tc::desc_kstep(LBO_V);
    // S(g) = Q K^T into S buffer g & 1, in two halves of four k-steps (part 0 waits for the operands,
    // part 1 commits).  In sweep 2 the halves are interleaved with the two P.V halves of the previous
    // tile — S(j+1) a | PV(j, A) | S(j+1) b | PV(j, B): the single P lo buffer can only be rewritten
    // for half B once PV(j, A) has completed, and with S(j+1) issued in one piece in front of both
    // PV halves the pipe idled ~650 cycles per tile waiting for that hand-over (trace:
    // scripts/microbench/att_trace.cu); now S(j+1) b runs meanwhile.  Same MMA order per accumulator.
    auto issue_s = [&](int g, int part) {
      const int kb = g & 1;
      if (part != 1) {
        tc::mbar_wait(&bar_kfull[kb], (g >> 1) & 1);
        if (lane == 0) ATT_TR(0, g);
        // the buffer's previous content was read by the workers (sweep 1) — in sweep 2 it holds P hi
        // of tile g - 2, whose PV MMAs were issued before this point and execute in issue order
        if (g >= 2 && g - 2 < T1) tc::mbar_wait(&bar_sfree[kb], ((g >> 1) - 1) & 1);
        tc::tc_fence_after();
      }
      const uint32_t k_hi = tc::smem_u32(smem + S::K_OFF + kb * 2 * S::K_TILE), k_lo = k_hi + S::K_TILE;
      const uint64_t dkh = tc::make_desc(k_hi, LBO_K, 128), dkl = tc::make_desc(k_lo, LBO_K, 128);
      const uint32_t d = tmem + COL_S + kb * 128;
      const int ks0 = part == 1 ? 4 : 0, ks1 = part == 0 ? 4 : 8;  // part 2: all eight k-steps
      if (g < T1 && approx) {
        for (int ks = ks0; ks < ks1; ++ks)
          tc::mma_tf32_ts_e(d, tmem + COL_QH + 8 * ks, dkh + ks * KK, IDESC_S, ks != 0);
      } else {
        for (int ks = ks0; ks < ks1; ++ks) {
          tc::mma_tf32_ts_e(d, tmem + COL_QL + 8 * ks, dkh + ks * KK, IDESC_S, ks != 0);
          tc::mma_tf32_ts_e(d, tmem + COL_QH + 8 * ks, dkl + ks * KK, IDESC_S, 1);
          tc::mma_tf32_ts_e(d, tmem + COL_QH + 8 * ks, dkh + ks * KK, IDESC_S, 1);
        }
      }
      if (part != 0) {
        tc::mma_commit_elect(&bar_kfree[kb]);
        tc::mma_commit_elect(&bar_sfull[kb]);
        if (lane == 0) ATT_TR(1, g);
      }
    };
    auto issue_pv = [&](int j, int h) {
      const uint32_t p_hi = tmem + COL_S + ((T1 + j) & 1) * 128;
      tc::mbar_wait(&bar_pfull[h], j & 1);
      if (lane == 0) ATT_TR(2 + h, j);
      tc::mbar_wait(&bar_vfull[h], j & 1);
      tc::tc_fence_after();
      const uint32_t v_hi = tc::smem_u32(smem + S::V_OFF + h * 2 * S::V_HALF), v_lo = v_hi + S::V_HALF;
      const uint64_t dvh = tc::make_desc(v_hi, LBO_V, 128), dvl = tc::make_desc(v_lo, LBO_V, 128);
      const uint32_t ah = p_hi + 64 * h, al = tmem + COL_PL;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        tc::mma_tf32_ts_e(tmem + COL_O, al + 8 * ks, dvh + ks * KV, IDESC_O, (j | h | ks) != 0);
        tc::mma_tf32_ts_e(tmem + COL_O, ah + 8 * ks, dvl + ks * KV, IDESC_O, 1);
        tc::mma_tf32_ts_e(tmem + COL_O, ah + 8 * ks, dvh + ks * KV, IDESC_O, 1);
      }
      tc::mma_commit_elect(&bar_vfree[h]);
      tc::mma_commit_elect(&bar_pvdone[h]);
      if (lane == 0) ATT_TR(4 + h, j);
    };
    for (int g = 0; g < T1; ++g) issue_s(g, 2);
    if (T > 0) issue_s(T1, 2);
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_s(T1 + j + 1, 0);
      issue_pv(j, 0);
      if (j + 1 < T) issue_s(T1 + j + 1, 1);
      issue_pv(j, 1);
    }
  } else {
    // ------------------------------ workers ---------------------------------------------------
    if (sweep1) {
      float m_run = -INFINITY;
      for (int g = 0; g < T1; ++g) {
        const int sb = g & 1;
        tc::mbar_wait(&bar_sfull[sb], (g >> 1) & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[16];
          tc::tmem_ld16(trow + COL_S + sb * 128 + 64 * h + 16 * cq, v);
          const int c0 = g * 128 + 64 * h + 16 * cq;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c0 + e < N) m_run = fmaxf(m_run, v[e]);
        }
        tc::tc_fence_before();
        a2_arrive(&bar_sfree[sb]);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");  // xch (the |q|^2 partials) has been read
      xch[cq * 128 + row] = m_run;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      m_row = fmaxf(fmaxf(xch[row], xch[128 + row]), fmaxf(xch[256 + row], xch[384 + row])) + margin;
      asm volatile("bar.sync 1, 512;" ::: "memory");  // read before the row sums reuse xch
    }
    float l_run = 0.f;
    for (int j = 0; j < T; ++j) {
      const int g = T1 + j, sb = g & 1;
      tc::mbar_wait(&bar_sfull[sb], (g >> 1) & 1);
      if (tid == 0) ATT_TR(6, j);
      tc::tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[16], lo[16];
        const uint32_t col = COL_S + sb * 128 + 64 * h + 16 * cq;
        tc::tmem_ld16(trow + col, v);
        const int c0 = j * 128 + 64 * h + 16 * cq;
        const bool full = (j + 1) * 128 <= N;  // CTA-uniform: no key of this tile lies beyond N
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float4 p;
          if (full) {
            p.x = expf(v[e + 0] - m_row);
            p.y = expf(v[e + 1] - m_row);
            p.z = expf(v[e + 2] - m_row);
            p.w = expf(v[e + 3] - m_row);
          } else {
            p.x = (c0 + e + 0 < N) ? expf(v[e + 0] - m_row) : 0.f;
            p.y = (c0 + e + 1 < N) ? expf(v[e + 1] - m_row) : 0.f;
            p.z = (c0 + e + 2 < N) ? expf(v[e + 2] - m_row) : 0.f;
            p.w = (c0 + e + 3 < N) ? expf(v[e + 3] - m_row) : 0.f;
          }
          l_run += (p.x + p.y) + (p.z + p.w);
          float4 ph, pl;
          tc::split4(p, ph, pl);
          v[e] = ph.x; v[e + 1] = ph.y; v[e + 2] = ph.z; v[e + 3] = ph.w;
          lo[e] = pl.x; lo[e + 1] = pl.y; lo[e + 2] = pl.z; lo[e + 3] = pl.w;
        }
        tc::tmem_st16(trow + col, v);  // P hi over the S columns it came from
        // the single P lo buffer is free once the PV MMAs of the previous half have completed
        if (h == 0) {
          if (j >= 1) tc::mbar_wait(&bar_pvdone[1], (j - 1) & 1);
        } else {
          tc::mbar_wait(&bar_pvdone[0], j & 1);
        }
        tc::tc_fence_after();
        tc::tmem_st16(trow + COL_PL + 16 * cq, lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        a2_arrive(&bar_pfull[h]);
        if (tid == 0 && h == 1) ATT_TR(7, j);
      }
    }
    // ---------------- epilogue: O / l ----------------------------------------------------------
    asm volatile("bar.sync 1, 512;" ::: "memory");
    xch[cq * 128 + row] = l_run;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const float inv = 1.f / ((xch[row] + xch[128 + row]) + (xch[256 + row] + xch[384 + row]));
    tc::mbar_wait(&bar_pvdone[1], (T - 1) & 1);
    tc::tc_fence_after();
    float v[16];
    tc::tmem_ld16(trow + COL_O + 16 * cq, v);
    const int q = q0 + row;
    if (q < N) {
      float* y = Y + map(base + q) * (int64_t)ldy + 16 * cq;
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        *reinterpret_cast<float4*>(y + e) =
            make_float4(v[e] * inv, v[e + 1] * inv, v[e + 2] * inv, v[e + 3] * inv);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem, 512);
}

size_t attention_split_bytes(int64_t B, int N) {
  return (size_t)B * (size_t)((N + 127) / 128) * Att2::BLK;
}

static bool att_v1_forced() {  // A/B switch: R3DFS_ATT_V1=1 (register-fed kernel)
  static const bool v = [] {
    const char* e = R3DFS_GETENV("R3DFS_ATT_V1");
    return e && e[0] == '1';
  }();
  return v;
}

// kmax_ws: B floats of scratch for the per-cloud key-norm maxima; nullptr -> always two sweeps
int launch_attention_tc(const float* qkv, int ld, int64_t B, int N, float* Y, int ldy, RowMap map,
                        cudaStream_t st, float* kmax_ws, void* split_ws, size_t split_bytes) {
  if ((ld & 3) != 0 || (ldy & 3) != 0) return R3DFS_E_UNSUPPORTED;
  if (kmax_ws) {
    att_kmax_kernel<<<(unsigned)B, 256, 0, st>>>(qkv, ld, N, kmax_ws);
    R3DFS_CHECK_LAUNCH();
  }
  if (split_ws && split_bytes >= attention_split_bytes(B, N) && N >= 1 && !att_v1_forced() &&
      (reinterpret_cast<uintptr_t>(split_ws) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(qkv) & 15) == 0) {
    const int T = (N + 127) / 128;
    cudaError_t e = cudaFuncSetAttribute(att_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         4 * Att2::V_HALF);
    if (e != cudaSuccess) return (int)e;
    att_split_kernel<<<dim3(T, (unsigned)B), 256, 4 * Att2::V_HALF, st>>>(qkv, ld, N,
                                                                        (unsigned char*)split_ws);
    R3DFS_CHECK_LAUNCH();
    e = cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Att2::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attention_tc2_kernel<<<dim3(T, (unsigned)B), A2_THREADS, Att2::TOTAL, st>>>(
        qkv, ld, N, (const unsigned char*)split_ws, Y, ldy, map, kmax_ws);
    R3DFS_CHECK_LAUNCH();
    return 0;
  }
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttTcSmem::TOTAL);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((N + AT_BQ - 1) / AT_BQ, (unsigned)B);
  attention_tc_kernel<<<grid, AT_ALL_THREADS, AttTcSmem::TOTAL, st>>>(qkv, ld, N, Y, ldy, map,
                                                                  kmax_ws);
  R3DFS_CHECK_LAUNCH();
  return 0;
}
