// EdgeConv tail on the tensor cores (reference models/dgcnn.py:56-57 second conv + BN + LeakyReLU,
// :118 max over the k neighbours), fused with the neighbour gather:
//     h1(i, j) = LReLU(P[nbr(i, j)] + Q[i])           (first conv already applied per point)
//     y(i)     = max_j LReLU(s2 * (W2 h1(i, j)) + t2)
// The (B, 64, N, k) edge activations only ever exist as UMMA operand tiles in shared memory and as
// accumulators in TMEM.
//
// One CTA walks over groups of 128 points.  MMA tile t of a group = neighbour slot t of its 128
// points (row = point), so a TMEM lane always belongs to the same point: the max over neighbours
// is a running max in the registers of the thread that owns the point — no exchange.
//   warps 4-11: producers.  Each thread owns 8 fixed (row, 16-byte chunk) cells of the A tile; its
//               Q values stay in registers for the whole group, per tile it gathers P[nbr] (indices
//               and rows gathered one tile ahead), adds, LeakyReLU, TF32 hi/lo split, stores
//               (2 stages), arrives on the stage's mbarrier.
//   warp 12:    one thread issues 3 x 8 tcgen05.mma (3xTF32, K = 64) per tile against the resident
//               W2 tile as soon as the stage is full and the accumulator buffer is drained
//               (decoupled from the producers by mbarriers) and commits to an mbarrier.
//   warps 0-3:  epilogue.  tcgen05.ld the 64 output channels of their row, BN affine +
//               LeakyReLU, running max; after the last slot the row is written out.
#include "common.cuh"
#include "tc.cuh"

// per-tile timestamps of one CTA (scripts/microbench/edge_trace.cu compiles this file with EDGE_TRACE)
#ifdef EDGE_TRACE
__device__ long long g_edge_trace[8 * 256];
#define EDGE_TR(slot, j) \
  do { if (blockIdx.x == EDGE_TRACE_BX && blockIdx.y == EDGE_TRACE_BY) g_edge_trace[(slot) * 256 + (j)] = clock64(); } while (0)
#else
#define EDGE_TR(slot, j) do { } while (0)
#endif
#define ET_THREADS 416  // warps 0-3 epilogue, 4-11 producers, 12 MMA issue
#define ET_PRODUCERS 256
#define ET_GROUP 128  // points per group (= MMA rows)

struct EdgeTcSmem {
  static constexpr int A_TILE = tc::tile_bytes(128, 16);  // one of hi / lo, K = 64
  static constexpr int W_TILE = tc::tile_bytes(64, 16);
  static constexpr int A_OFF = 0;           // 2 stages x (hi, lo)
  static constexpr int W_OFF = 4 * A_TILE;  // W2 hi, lo
  static constexpr int TOTAL = W_OFF + 2 * W_TILE + 64;
};

__device__ __forceinline__ void et_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ __launch_bounds__(ET_THREADS, 1) void edge_tc_kernel(
    const float* __restrict__ PQ, const int32_t* __restrict__ idx, const float* __restrict__ w2,
    const float* __restrict__ s2, const float* __restrict__ t2, int N, int k, int groups_per_cta,
    float* __restrict__ Y, int ldy, RowMap map) {
  extern __shared__ __align__(128) unsigned char smem[];
  using S = EdgeTcSmem;
  __shared__ uint64_t bar_full[2];   // accumulator b ready (also: operand stage b free again)
  __shared__ uint64_t bar_tfree[2];  // accumulator b drained by the 128 epilogue threads
  __shared__ float s_aff[128];       // BN scale (0..63) and shift (64..127) of the second conv
  __shared__ uint64_t bar_sfull[2];  // operand stage b written by the 256 producer threads
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int64_t base = (int64_t)b * N;
  const int n_groups = (N + ET_GROUP - 1) / ET_GROUP;
  const int g_begin = blockIdx.x * groups_per_cta;
  const int g_end = min(n_groups, g_begin + groups_per_cta);
  const int U = (g_end - g_begin) * k;  // one tile per (group, neighbour slot)
  constexpr int LBO_A = tc::tile_lbo(128), LBO_W = tc::tile_lbo(64);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, 64);

  if (tid == 0) {
    tc::mbar_init(&bar_full[0], 1);
    tc::mbar_init(&bar_full[1], 1);
    tc::mbar_init(&bar_tfree[0], 128);
    tc::mbar_init(&bar_tfree[1], 128);
    tc::mbar_init(&bar_sfull[0], ET_PRODUCERS);
    tc::mbar_init(&bar_sfull[1], ET_PRODUCERS);
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 128);
  if (tid < 64) s_aff[tid] = __ldg(s2 + tid);
  else if (tid < 128) s_aff[tid] = __ldg(t2 + tid - 64);
  // resident W2 (64 x 64, row-major [c][kk] = K-major) as hi / lo tiles
  for (int c = tid; c < 64 * 16; c += ET_THREADS) {
    const int r = c >> 4, kc = c & 15;
    const float4 v = *reinterpret_cast<const float4*>(w2 + r * 64 + 4 * kc);
    float4 hi, lo;
    tc::split4(v, hi, lo);
    *reinterpret_cast<float4*>(smem + S::W_OFF + kc * LBO_W + r * 16) = hi;
    *reinterpret_cast<float4*>(smem + S::W_OFF + S::W_TILE + kc * LBO_W + r * 16) = lo;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  if (U <= 0) {
    if (w == 0) tc::tmem_dealloc(tmem_d, 128);
    return;
  }

  if (w == 12) {
    // --------------------------- MMA issue: its own warp ----------------------------------------
    // (a producer thread issuing the MMAs holds its whole producer group at the tile barrier for
    // as long as the tensor pipe takes to accept them, so staging and MMAs could not overlap)
    // the whole warp walks the loop and one elected lane issues (warp-uniform control flow: the
    // compiler wraps every MMA issued under `if (lane == 0)` in a convergence loop)
    {
      const uint32_t wh = tc::smem_u32(smem + S::W_OFF), wl = wh + S::W_TILE;
      const uint64_t dwh = tc::make_desc(wh, LBO_W, 128), dwl = tc::make_desc(wl, LBO_W, 128);
      constexpr uint64_t KA = tc::desc_kstep(LBO_A), KW = tc::desc_kstep(LBO_W);
      for (int u = 0; u < U; ++u) {
        const int st = u & 1;
        tc::mbar_wait(&bar_sfull[st], (u >> 1) & 1);                      // A tile staged
        if (lane == 0) EDGE_TR(0, u);
        if (u >= 2) tc::mbar_wait(&bar_tfree[st], ((u >> 1) - 1) & 1);    // accumulator drained
        if (lane == 0) EDGE_TR(1, u);
        tc::tc_fence_after();
        const uint32_t ah = tc::smem_u32(smem + S::A_OFF + st * 2 * S::A_TILE), al = ah + S::A_TILE;
        const uint32_t d = tmem_d + st * 64;
        const uint64_t dah = tc::make_desc(ah, LBO_A, 128), dal = tc::make_desc(al, LBO_A, 128);
        tc::mma_tf32_elect(d, dal, dwh, IDESC, 0);
        tc::mma_tf32_elect(d, dah, dwl, IDESC, 1);
        tc::mma_tf32_elect(d, dah, dwh, IDESC, 1);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks) {
          tc::mma_tf32_elect(d, dal + ks * KA, dwh + ks * KW, IDESC, 1);
          tc::mma_tf32_elect(d, dah + ks * KA, dwl + ks * KW, IDESC, 1);
          tc::mma_tf32_elect(d, dah + ks * KA, dwh + ks * KW, IDESC, 1);
        }
        tc::mma_commit_elect(&bar_full[st]);
        if (lane == 0) EDGE_TR(2, u);
      }
    }
  } else if (w >= 4) {
    // --------------------------- producers -----------------------------------------------------
    const int lt = tid - 128;
    constexpr int NCH = 128 * 16 / ET_PRODUCERS;  // 8 cells per thread: rows (lt>>4) + 16 i
    const int kc = lt & 15;
    const int r0 = lt >> 4;
    int nb[NCH];
    float4 qv4[NCH];
    auto load_nb = [&](int u2) {  // neighbour indices of tile u2 (prefetched one tile ahead)
      const int g2 = g_begin + u2 / k, j = u2 % k;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int p = g2 * ET_GROUP + r0 + 16 * i;
        nb[i] = (u2 < U && p < N) ? idx[(base + p) * k + j] : -1;
      }
    };
    // Software pipeline: the gathers of tile u + 1 are issued as soon as tile u has been written
    // to its stage (their registers are free again), so their L2 round trip overlaps the fence,
    // the producers' barrier, the MMA issue and the wait for the next stage; the neighbour
    // indices run one tile further ahead.
    float4 hv[NCH];
    int live = 0;
    auto issue_gather = [&]() {  // rows nb[] of P -> hv (all loads in flight together)
      live = 0;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int nbi = nb[i] >= 0 ? nb[i] : 0;
        live |= (nb[i] >= 0) << i;
        hv[i] = __ldg(reinterpret_cast<const float4*>(PQ + (base + nbi) * 128 + 4 * kc));
      }
    };
    load_nb(0);
    issue_gather();
    load_nb(1);
    for (int u = 0; u < U; ++u) {
      const int st = u & 1;
      const int g = g_begin + u / k;
      if (lt == 0) EDGE_TR(3, u);
      if (u % k == 0) {  // new group: this thread's Q cells stay in registers for all k tiles
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int p = min(g * ET_GROUP + r0 + 16 * i, N - 1);
          qv4[i] = __ldg(reinterpret_cast<const float4*>(PQ + (base + p) * 128 + 64 + 4 * kc));
        }
      }
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((live >> i) & 1) {
          h.x = hv[i].x + qv4[i].x; h.y = hv[i].y + qv4[i].y;
          h.z = hv[i].z + qv4[i].z; h.w = hv[i].w + qv4[i].w;
          h.x = h.x > 0.f ? h.x : 0.2f * h.x;
          h.y = h.y > 0.f ? h.y : 0.2f * h.y;
          h.z = h.z > 0.f ? h.z : 0.2f * h.z;
          h.w = h.w > 0.f ? h.w : 0.2f * h.w;
        }
        hv[i] = h;
      }
      if (lt == 0) EDGE_TR(4, u);
      if (u >= 2) tc::mbar_wait(&bar_full[st], ((u >> 1) - 1) & 1);  // stage's MMAs finished
      if (lt == 0) EDGE_TR(5, u);
      unsigned char* a_hi = smem + S::A_OFF + st * 2 * S::A_TILE;
      unsigned char* a_lo = a_hi + S::A_TILE;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int r = r0 + 16 * i;
        float4 hi, lo;
        tc::split4(hv[i], hi, lo);
        *reinterpret_cast<float4*>(a_hi + kc * LBO_A + r * 16) = hi;
        *reinterpret_cast<float4*>(a_lo + kc * LBO_A + r * 16) = lo;
      }
      if (u + 1 < U) {
        issue_gather();  // nb holds the indices of tile u + 1
        load_nb(u + 2);
      }
      tc::fence_async_smem();
      et_mbar_arrive(&bar_sfull[st]);
      if (lt == 0) EDGE_TR(6, u);
    }
  } else {
    // --------------------------- epilogue: thread = point row, all 64 channels ------------------
    const int row = 32 * w + lane;
    float mx[64];
    for (int u = 0; u < U; ++u) {
      const int st = u & 1;
      const int g = g_begin + u / k, j = u % k;
      if (j == 0) {
#pragma unroll
        for (int c = 0; c < 64; ++c) mx[c] = -INFINITY;
      }
      tc::mbar_wait(&bar_full[st], (u >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(32 * w) << 16) + (uint32_t)(st * 64 + 32 * hh), v);
        if (hh == 1) {  // both halves are in registers: the accumulator buffer is free
          tc::tc_fence_before();
          et_mbar_arrive(&bar_tfree[st]);
          if (tid == 0) EDGE_TR(7, u);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float y = fmaf(s_aff[32 * hh + c], v[c], s_aff[64 + 32 * hh + c]);
          y = fmaxf(y, 0.2f * y);  // LeakyReLU(0.2)
          mx[32 * hh + c] = fmaxf(mx[32 * hh + c], y);
        }
      }
      if (j == k - 1) {
        const int p = g * ET_GROUP + row;
        if (p < N) {
          float* y = Y + map(base + p) * (int64_t)ldy;
#pragma unroll
          for (int c = 0; c < 64; c += 4)
            *reinterpret_cast<float4*>(y + c) = make_float4(mx[c], mx[c + 1], mx[c + 2], mx[c + 3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem_d, 128);
}

int launch_edge_mlp_tc(const float* PQ, const int32_t* idx, const float* w2, const float* s2,
                       const float* t2, int64_t B, int N, int k, float* Y, int ldy, RowMap map,
                       cudaStream_t st) {
  if (k < 1 || k > 64) return R3DFS_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(edge_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       EdgeTcSmem::TOTAL);
  if (e != cudaSuccess) return (int)e;
  const int n_groups = (N + ET_GROUP - 1) / ET_GROUP;
  // several groups per CTA (W2 tile, TMEM allocation and pipeline prologue amortised) as long as
  // the grid still fills the machine a few times over
  int gpc = 2;
  while (gpc > 1 && (int64_t)B * ((n_groups + gpc - 1) / gpc) < 4 * 148) gpc >>= 1;
  dim3 grid((n_groups + gpc - 1) / gpc, (unsigned)B);
  edge_tc_kernel<<<grid, ET_THREADS, EdgeTcSmem::TOTAL, st>>>(PQ, idx, w2, s2, t2, N, k, gpc, Y,
                                                             ldy, map);
  R3DFS_CHECK_LAUNCH();
  return 0;
}
