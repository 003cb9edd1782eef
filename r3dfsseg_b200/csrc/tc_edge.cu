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
#define ET_THREADS 416  // warps 0-3 epilogue, 4-7 gather, 8-11 convert, 12 MMA issue
#define ET_GATHERERS 128
#define ET_GROUP 128  // points per group (= MMA rows)

#define ET_RAW 4  // gathered P tiles in flight (cp.async ring)
struct EdgeTcSmem {
  static constexpr int W_TILE = tc::tile_bytes(64, 16);
  static constexpr int RAW_TILE = 128 * 64 * 4;  // gathered P rows, FP32: [row][16 chunks], XOR-swizzled
  static constexpr int RAW_OFF = 0;              // ET_RAW stages
  static constexpr int W_OFF = ET_RAW * RAW_TILE;  // W2 hi, lo (the split A operand lives in TMEM)
  static constexpr int TOTAL = W_OFF + 2 * W_TILE + 64;
};
// TMEM columns: accumulators 0-63 | 64-127, A stage s at 128 + 128 s (hi 64 columns, lo 64 columns)
#define ET_COL_A 128

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
  __shared__ uint64_t bar_sfull[2];  // TMEM operand stage b written by the 128 converter threads
  __shared__ uint64_t bar_raw[ET_RAW];    // gathered tile landed (cp.async arrivals of the 128 gather threads)
  __shared__ uint64_t bar_rfree[ET_RAW];  // gathered tile read back by the 128 converter threads
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int64_t base = (int64_t)b * N;
  const int n_groups = (N + ET_GROUP - 1) / ET_GROUP;
  const int g_begin = blockIdx.x * groups_per_cta;
  const int g_end = min(n_groups, g_begin + groups_per_cta);
  const int U = (g_end - g_begin) * k;  // one tile per (group, neighbour slot)
  constexpr int LBO_W = tc::tile_lbo(64);
  constexpr uint32_t IDESC = tc::make_idesc_tf32(128, 64);

  if (tid == 0) {
    tc::mbar_init(&bar_full[0], 1);
    tc::mbar_init(&bar_full[1], 1);
    tc::mbar_init(&bar_tfree[0], 128);
    tc::mbar_init(&bar_tfree[1], 128);
    tc::mbar_init(&bar_sfull[0], 128);
    tc::mbar_init(&bar_sfull[1], 128);
    for (int i = 0; i < ET_RAW; ++i) {
      tc::mbar_init(&bar_raw[i], ET_GATHERERS);
      tc::mbar_init(&bar_rfree[i], 128);
    }
    tc::mbar_fence_init();
  }
  if (w == 0) tc::tmem_alloc(&tmem_base_s, 512);
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
    if (w == 0) tc::tmem_dealloc(tmem_d, 512);
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
      constexpr uint64_t KW = tc::desc_kstep(LBO_W);
      for (int u = 0; u < U; ++u) {
        const int st = u & 1;
        tc::mbar_wait(&bar_sfull[st], (u >> 1) & 1);                      // A stage written (TMEM)
        if (lane == 0) EDGE_TR(0, u);
        if (u >= 2) tc::mbar_wait(&bar_tfree[st], ((u >> 1) - 1) & 1);    // accumulator drained
        if (lane == 0) EDGE_TR(1, u);
        tc::tc_fence_after();
        const uint32_t ah = tmem_d + ET_COL_A + st * 128, al = ah + 64;
        const uint32_t d = tmem_d + st * 64;
        tc::mma_tf32_ts_e(d, al, dwh, IDESC, 0);
        tc::mma_tf32_ts_e(d, ah, dwl, IDESC, 1);
        tc::mma_tf32_ts_e(d, ah, dwh, IDESC, 1);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks) {
          tc::mma_tf32_ts_e(d, al + 8 * ks, dwh + ks * KW, IDESC, 1);
          tc::mma_tf32_ts_e(d, ah + 8 * ks, dwl + ks * KW, IDESC, 1);
          tc::mma_tf32_ts_e(d, ah + 8 * ks, dwh + ks * KW, IDESC, 1);
        }
        tc::mma_commit_elect(&bar_full[st]);
        if (lane == 0) EDGE_TR(2, u);
      }
    }
  } else if (w >= 4) {
    // --------------------------- producers -----------------------------------------------------
    // Two producer roles, decoupled by a ring of ET_RAW gathered tiles in shared memory:
    //  warps 4-7  GATHER: cp.async, up to ET_RAW tiles ahead, no registers.  A thread copies 16
    //             fixed (row, 16-byte chunk) cells of every tile — half a warp fetches one contiguous
    //             256-byte row of P — to chunk c of row r at r * 256 + ((c ^ (r & 15)) * 16
    //             (conflict-free for this mapping and for the converters') and lets the copies
    //             arrive on the tile's mbarrier; the neighbour indices run one tile ahead.
    //  warps 8-11 CONVERT: thread = point row (a warp owns TMEM lane quarter w % 4): reads its row
    //             back, adds Q (in registers for the whole group), LeakyReLU, TF32 hi / lo split,
    //             and writes both parts into the TMEM A stage (tcgen05.st).
    // The MMA reads only W2 from shared memory.  Before: every producer thread gathered into
    // registers one tile ahead, converted, and stored hi / lo operand tiles to shared memory (64 KB
    // written + 96 KB read by the MMAs per tile next to 48 KB of W2 reads): that serial chain per
    // tile — wait for the gather + add 900, stores 1200, next gather 500 cycles — set the pace at
    // 2850 cycles per tile with 1150 of MMAs.  Measured and dropped on the way: a lane-per-row
    // register gather (8x the L1 wavefronts), staging through shared memory behind a producer-wide
    // barrier, 12 / 16 producer warps and 8 gather warps with setmaxnreg (no gain: the copies of a
    // tile take ~1800 cycles however many warps issue them), one 256-byte cp.async.bulk per row (the
    // TMA engine needs ~23 cycles per small copy: 3000 cycles per tile), and all eight warps doing
    // both roles in lockstep (2400: the LSU idles while everybody converts and vice versa).
    if (w < 8) {
      const int gt = tid - 128;
      constexpr int NCH = 16;  // cells per thread: rows (gt >> 4) + 8 i
      const int kc = gt & 15;
      const int r0 = gt >> 4;
      // neighbour indices run TWO tiles ahead of their copies (two register sets): loaded one tile
      // ahead, their L2 round trip sat in front of every tile's copies
      int nbA[NCH], nbB[NCH];
      auto load_nb = [&](int (&nb)[NCH], int u2) {
        const int g2 = g_begin + u2 / k, j = u2 % k;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int p = g2 * ET_GROUP + r0 + 8 * i;
          nb[i] = (u2 < U && p < N) ? idx[(base + p) * k + j] : 0;  // rows beyond N are never stored
        }
      };
      auto copy_tile = [&](const int (&nb)[NCH], int u) {
        const int rs = u % ET_RAW;
        if (u >= ET_RAW) tc::mbar_wait(&bar_rfree[rs], ((u / ET_RAW) - 1) & 1);
        if (gt == 0) EDGE_TR(3, u);
        const uint32_t dst0 = tc::smem_u32(smem + S::RAW_OFF + rs * S::RAW_TILE);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int r = r0 + 8 * i;
          const float* src = PQ + (base + nb[i]) * 128 + 4 * kc;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                           dst0 + r * 256 + ((kc ^ (r & 15)) << 4)),
                       "l"(src)
                       : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                         tc::smem_u32(&bar_raw[rs]))
                     : "memory");
      };
      load_nb(nbA, 0);
      load_nb(nbB, 1);
      for (int u = 0; u < U; u += 2) {
        copy_tile(nbA, u);
        load_nb(nbA, u + 2);
        if (gt == 0) EDGE_TR(4, u);
        if (u + 1 < U) {
          copy_tile(nbB, u + 1);
          load_nb(nbB, u + 3);
          if (gt == 0) EDGE_TR(4, u + 1);
        }
      }
    } else {
      const int row = 32 * (w & 3) + lane;
      const uint32_t trow = tmem_d + ((uint32_t)(32 * (w & 3)) << 16) + ET_COL_A;
      float4 qv4[16];
      for (int u = 0; u < U; ++u) {
        const int st = u & 1, rs = u % ET_RAW;
        const int g = g_begin + u / k;
        if (u % k == 0) {  // new group: this row's Q values stay in registers for all k tiles
          const int p = min(g * ET_GROUP + row, N - 1);
          const float4* src = reinterpret_cast<const float4*>(PQ + (base + p) * 128 + 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) qv4[i] = __ldg(src + i);
        }
        tc::mbar_wait(&bar_raw[rs], (u / ET_RAW) & 1);  // the gathered tile has landed
        if (u >= 2) tc::mbar_wait(&bar_full[st], ((u >> 1) - 1) & 1);  // TMEM A stage: its MMAs finished
        tc::tc_fence_after();
        if (tid == 256) EDGE_TR(5, u);
        const unsigned char* raw = smem + S::RAW_OFF + rs * S::RAW_TILE + row * 256;
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // 16 channels at a time
          float hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = 4 * q + i;
            const float4 g4 = *reinterpret_cast<const float4*>(raw + ((c ^ (row & 15)) << 4));
            const float4 q4 = qv4[c];
            float4 h;
            h.x = g4.x + q4.x; h.y = g4.y + q4.y; h.z = g4.z + q4.z; h.w = g4.w + q4.w;
            h.x = h.x > 0.f ? h.x : 0.2f * h.x;
            h.y = h.y > 0.f ? h.y : 0.2f * h.y;
            h.z = h.z > 0.f ? h.z : 0.2f * h.z;
            h.w = h.w > 0.f ? h.w : 0.2f * h.w;
            float4 ph, pl;
            tc::split4(h, ph, pl);
            hi[4 * i] = ph.x; hi[4 * i + 1] = ph.y; hi[4 * i + 2] = ph.z; hi[4 * i + 3] = ph.w;
            lo[4 * i] = pl.x; lo[4 * i + 1] = pl.y; lo[4 * i + 2] = pl.z; lo[4 * i + 3] = pl.w;
          }
          tc::tmem_st16(trow + st * 128 + 16 * q, hi);
          tc::tmem_st16(trow + st * 128 + 64 + 16 * q, lo);
        }
        et_mbar_arrive(&bar_rfree[rs]);  // (the loads above have returned: their values were used)
        tc::tmem_st_wait();
        tc::tc_fence_before();
        et_mbar_arrive(&bar_sfull[st]);
        if (tid == 256) EDGE_TR(6, u);
      }
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
  if (w == 0) tc::tmem_dealloc(tmem_d, 512);
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
