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
//     small pre-kernel) is <= 60 for every row of the CTA, m' = b and sweep 1 is skipped
//     (exp(-60) = 9e-27 is far above FP32's 1e-38);
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
    sweep1 = __syncthreads_or(!(m_row <= 60.f));
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

// kmax_ws: B floats of scratch for the per-cloud key-norm maxima; nullptr -> always two sweeps
int launch_attention_tc(const float* qkv, int ld, int64_t B, int N, float* Y, int ldy, RowMap map,
                        cudaStream_t st, float* kmax_ws) {
  if ((ld & 3) != 0 || (ldy & 3) != 0) return R3DFS_E_UNSUPPORTED;
  if (kmax_ws) {
    att_kmax_kernel<<<(unsigned)B, 256, 0, st>>>(qkv, ld, N, kmax_ws);
    R3DFS_CHECK_LAUNCH();
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
