// Farthest point sampling with an on-chip int8 filter (torch_cluster.fps as called at reference
// models/mpti.py:613; same picks, bit for bit, as fps_kernel in proto.cu).
//
// fps_kernel re-reads every FP32 row of a set for every pick (m sweeps of n x 768 B; 29 GB of HBM
// traffic per 25 episodes).  Only ~3 % of the rows change their running minimum in a pick, so here
// each row is kept as 192 unsigned bytes q (x ~ lo_d + q_d * step, one step per set) in the SHARED
// MEMORY of the set's thread-block cluster, with the exact norm eps_i of its quantisation error.
// Distances between quantised rows are exact integers (dp4a), so
//      |x_i - x_s|  >=  step * sqrt(|q_i - q_s|^2) - eps_i - eps_s
// is a rigorous lower bound; a row is re-read in FP32 from HBM only when that bound cannot prove
// that its running minimum stays (measured: 3-4 % of the rows per pick).  The FP32 arithmetic of
// the rows that are re-read is fps_kernel's (8 lanes per row, same fma chain and shuffle order),
// so the pick sequence is identical.
#include <cooperative_groups.h>

#include "common.cuh"
#include "proto.cuh"

namespace cg = cooperative_groups;

#define FQ_THREADS 512
#define FQ_WARPS (FQ_THREADS / 32)
#define FQ_D 192
#define FQ_QW (FQ_D / 4)    // 48 packed words per row
#define FQ_ROWB 208         // 192 B + 16 B: an odd number of 16-byte units -> conflict-free LDS.128
#define FQ_MAX_CL 16
#define FQ_CANDB 224        // candidate record: q row (208 B) + {value, index, eps, qn}

struct FqMeta {
  float v;
  int idx;
  float eps;
  int qn;
};

__device__ __forceinline__ int fq_target_count(int n, int k) {
  float ratio = (float)((double)k / (double)n);
  float prod = __fmul_rn((float)n, ratio);
  return (int)ceilf(prod);
}

// Row li of this CTA (global index i = li * CL + rank): resident rows live in shared memory,
// the rest in the global spill area (FQ_ROWB bytes per row: q | dist | eps | qn).
struct FqRows {
  uint8_t* s_q;
  float* s_dist;
  float* s_eps;
  int* s_qn;
  uint8_t* g_rows;  // spill area of this set (row index = global i)
  int r_res, CL, rank;
  __device__ __forceinline__ bool res(int li) const { return li < r_res; }
  __device__ __forceinline__ uint8_t* spill(int li) const {
    return g_rows + (int64_t)(li * CL + rank) * FQ_ROWB;
  }
  __device__ __forceinline__ const uint4* q(int li) const {
    return reinterpret_cast<const uint4*>(res(li) ? s_q + (size_t)li * FQ_ROWB : spill(li));
  }
  __device__ __forceinline__ uint32_t* qw(int li) const {
    return reinterpret_cast<uint32_t*>(res(li) ? s_q + (size_t)li * FQ_ROWB : spill(li));
  }
  __device__ __forceinline__ float* dist(int li) const {
    return res(li) ? s_dist + li : reinterpret_cast<float*>(spill(li) + FQ_D);
  }
  __device__ __forceinline__ float* eps(int li) const {
    return res(li) ? s_eps + li : reinterpret_cast<float*>(spill(li) + FQ_D + 4);
  }
  __device__ __forceinline__ int* qn(int li) const {
    return res(li) ? s_qn + li : reinterpret_cast<int*>(spill(li) + FQ_D + 8);
  }
};

__global__ __launch_bounds__(FQ_THREADS, 1) void fps_q8_kernel(
    const float* __restrict__ feat, const int32_t* __restrict__ set_off,
    const int32_t* __restrict__ set_n, int set_first, int set_per, int set_stride, int m_max,
    int k_for_count, int r_res, uint8_t* __restrict__ spill, int32_t* __restrict__ idx_out,
    int32_t* __restrict__ cnt_out) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = cluster.num_blocks();
  const int rank = cluster.block_rank();
  const int set = (blockIdx.y / set_per) * set_stride + set_first + blockIdx.y % set_per;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float s_lo[FQ_D];
  __shared__ float s_red[2][FQ_WARPS];
  __shared__ float s_wv[FQ_WARPS];
  __shared__ int s_wi[FQ_WARPS];
  __shared__ float s_step, s_absmax;

  const int n = set_n[set];
  const int64_t row0 = set_off[set];
  int m;
  if (k_for_count > 0)
    m = (n > k_for_count) ? fq_target_count(n, k_for_count) : 0;
  else
    m = min(m_max, n);
  m = min(m, m_max);
  if (rank == 0 && threadIdx.x == 0 && cnt_out) cnt_out[set] = m;
  int32_t* out = idx_out + (int64_t)set * m_max;
  if (m <= 0) return;  // uniform over the cluster: nobody reaches a barrier

  cluster.sync();  // every CTA of the cluster runs before anyone writes into a peer's shared memory
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  // shared-memory carve: resident rows | dist | eps | qn | candidate exchange [2][CL]
  FqRows R;
  R.s_q = smem;
  R.s_dist = reinterpret_cast<float*>(smem + (size_t)r_res * FQ_ROWB);
  R.s_eps = R.s_dist + r_res;
  R.s_qn = reinterpret_cast<int*>(R.s_eps + r_res);
  uint8_t* s_cand = reinterpret_cast<uint8_t*>(R.s_qn + r_res);
  R.g_rows = spill + row0 * FQ_ROWB;
  R.r_res = r_res;
  R.CL = CL;
  R.rank = rank;
  const int n_loc = (n - rank + CL - 1) / CL;  // rows i = li * CL + rank < n
  const float* fset = feat + row0 * (int64_t)FQ_D;

  // ---- phase 0a: dist_i = |x_i - x_0|^2 (the first sweep of fps_kernel) and per-dimension
  //      min / max of the slice ------------------------------------------------------------------
  float4 sfrag[6];
  {
    const float4* srow = reinterpret_cast<const float4*>(fset);
#pragma unroll
    for (int u = 0; u < 6; ++u) sfrag[u] = srow[sub + 8 * u];
  }
  float mn[24], mx[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    mn[j] = INFINITY;
    mx[j] = -INFINITY;
  }
  for (int lb = w * 4; lb < n_loc; lb += FQ_WARPS * 4) {
    const int li = lb + grp;
    const bool valid = li < n_loc;
    float acc = 0.f;
    if (valid) {
      const float4* xrow =
          reinterpret_cast<const float4*>(fset + (int64_t)(li * CL + rank) * FQ_D);
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        float4 x = xrow[sub + 8 * u];
        float d0 = x.x - sfrag[u].x, d1 = x.y - sfrag[u].y, d2 = x.z - sfrag[u].z,
              d3 = x.w - sfrag[u].w;
        acc = fmaf(d0, d0, acc);
        acc = fmaf(d1, d1, acc);
        acc = fmaf(d2, d2, acc);
        acc = fmaf(d3, d3, acc);
        mn[4 * u + 0] = fminf(mn[4 * u + 0], x.x);
        mx[4 * u + 0] = fmaxf(mx[4 * u + 0], x.x);
        mn[4 * u + 1] = fminf(mn[4 * u + 1], x.y);
        mx[4 * u + 1] = fmaxf(mx[4 * u + 1], x.y);
        mn[4 * u + 2] = fminf(mn[4 * u + 2], x.z);
        mx[4 * u + 2] = fmaxf(mx[4 * u + 2], x.z);
        mn[4 * u + 3] = fminf(mn[4 * u + 3], x.w);
        mx[4 * u + 3] = fmaxf(mx[4 * u + 3], x.w);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (valid && sub == 0) *R.dist(li) = acc;
  }
  // slice min/max: over the 4 row groups of a warp, the warps (shared memory), the cluster (DSMEM)
  float* s_part = reinterpret_cast<float*>(smem);                   // [FQ_WARPS][384], dead q area
  float* s_xch = s_part + FQ_WARPS * 2 * FQ_D;                      // [CL][384]
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], 8));
    mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], 16));
    mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 8));
    mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 16));
  }
  if (grp == 0) {
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const int d = 4 * (sub + 8 * (j >> 2)) + (j & 3);
      s_part[w * 2 * FQ_D + d] = mn[j];
      s_part[w * 2 * FQ_D + FQ_D + d] = mx[j];
    }
  }
  __syncthreads();
  if (tid < 2 * FQ_D) {
    float v = s_part[tid];
    for (int q = 1; q < FQ_WARPS; ++q) {
      float o = s_part[q * 2 * FQ_D + tid];
      v = tid < FQ_D ? fminf(v, o) : fmaxf(v, o);
    }
    for (int r = 0; r < CL; ++r) *cluster.map_shared_rank(&s_xch[rank * 2 * FQ_D + tid], r) = v;
  }
  cluster.sync();
  float rng = 0.f, ab = 0.f;
  if (tid < FQ_D) {
    float lo = s_xch[tid], hi = s_xch[FQ_D + tid];
    for (int r = 1; r < CL; ++r) {
      lo = fminf(lo, s_xch[r * 2 * FQ_D + tid]);
      hi = fmaxf(hi, s_xch[r * 2 * FQ_D + FQ_D + tid]);
    }
    s_lo[tid] = lo;
    rng = hi - lo;
    ab = fmaxf(fabsf(lo), fabsf(hi));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rng = fmaxf(rng, __shfl_xor_sync(0xffffffffu, rng, o));
    ab = fmaxf(ab, __shfl_xor_sync(0xffffffffu, ab, o));
  }
  if (lane == 0) {
    s_red[0][w] = rng;
    s_red[1][w] = ab;
  }
  __syncthreads();
  if (tid == 0) {
    float r = 0.f, a = 0.f;
    for (int q = 0; q < FQ_WARPS; ++q) {
      r = fmaxf(r, s_red[0][q]);
      a = fmaxf(a, s_red[1][q]);
    }
    // one step for the whole set, so that |q_i - q_s|^2 (an integer) times step^2 is the exact
    // squared distance of the quantised rows; a degenerate set (all rows equal) gets step 1
    float st = __fmul_rn(__fdiv_rn(r, 255.f), 1.000001f);
    s_step = (st > 0.f && st < INFINITY) ? st : 1.f;
    s_absmax = a;
  }
  __syncthreads();  // also: every thread is done with s_part / s_xch before q rows overwrite them
  const float step = s_step, inv_step = 1.f / step;
  // slack on eps: |x - (lo + q step)| as evaluated in FP32 vs exactly: each of the 192 differences
  // is off by <= 2^-24 (|x| + |lo + q step|), the sum of squares by < 2e-5 relative
  const float eps_slack = 2e-6f * s_absmax;

  // ---- phase 0c: quantise the slice ---------------------------------------------------------------
  {
    float lo[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) lo[j] = s_lo[4 * (sub + 8 * (j >> 2)) + (j & 3)];
    for (int lb = w * 4; lb < n_loc; lb += FQ_WARPS * 4) {
      const int li = lb + grp;
      const bool valid = li < n_loc;
      float acc = 0.f;
      int qq = 0;
      uint32_t pk[6];
      if (valid) {
        const float4* xrow =
            reinterpret_cast<const float4*>(fset + (int64_t)(li * CL + rank) * FQ_D);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          float4 x = xrow[sub + 8 * u];
          const float xv[4] = {x.x, x.y, x.z, x.w};
          uint32_t p = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            int q = __float2int_rn((xv[c] - lo[4 * u + c]) * inv_step);
            q = min(255, max(0, q));
            float t = fmaf((float)q, step, lo[4 * u + c]);
            float df = xv[c] - t;
            acc = fmaf(df, df, acc);
            qq += q * q;
            p |= (uint32_t)q << (8 * c);
          }
          pk[u] = p;
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      qq += __shfl_xor_sync(0xffffffffu, qq, 1);
      qq += __shfl_xor_sync(0xffffffffu, qq, 2);
      qq += __shfl_xor_sync(0xffffffffu, qq, 4);
      if (valid) {
        uint32_t* qrow = R.qw(li);
#pragma unroll
        for (int u = 0; u < 6; ++u) qrow[sub + 8 * u] = pk[u];
        if (sub == 0) {
          *R.eps(li) = fmaf(sqrtf(acc), 1.0001f, eps_slack);
          *R.qn(li) = qq;
        }
      }
    }
  }
  __syncthreads();

  // ---- picks ------------------------------------------------------------------------------------
  if (rank == 0 && tid == 0) out[0] = 0;
  for (int pick = 1; pick < m; ++pick) {
    // argmax of the running minima (ties -> lowest global index)
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int li = tid; li < n_loc; li += FQ_THREADS) {
      float v = *R.dist(li);
      if (v > bv) {  // li increases within a thread: strict > keeps the lowest index
        bv = v;
        bi = li * CL + rank;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_wv[w] = bv;
      s_wi[w] = bi;
    }
    __syncthreads();
    bv = lane < FQ_WARPS ? s_wv[lane] : -2.f;
    bi = lane < FQ_WARPS ? s_wi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    bv = __shfl_sync(0xffffffffu, bv, 0);
    bi = __shfl_sync(0xffffffffu, bi, 0);
    // publish this CTA's candidate (value, index, q row, eps, qn): warp r writes to peer r
    const int par = pick & 1;
    if (w < CL) {
      uint8_t* dst = cluster.map_shared_rank(s_cand + (size_t)(par * CL + rank) * FQ_CANDB, w);
      const bool have = bi != 0x7fffffff;
      const int lbest = have ? bi / CL : 0;
      if (lane < 12) {
        uint4 qv = have ? R.q(lbest)[lane] : make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(dst)[lane] = qv;
      } else if (lane == 12) {
        FqMeta mt;
        mt.v = bv;
        mt.idx = bi;
        mt.eps = have ? *R.eps(lbest) : 0.f;
        mt.qn = have ? *R.qn(lbest) : 0;
        *reinterpret_cast<FqMeta*>(dst + FQ_ROWB) = mt;
      }
    }
    cluster.sync();
    const uint8_t* cbase = s_cand + (size_t)par * CL * FQ_CANDB;
    int wr = 0;
    FqMeta win = *reinterpret_cast<const FqMeta*>(cbase + FQ_ROWB);
    for (int r = 1; r < CL; ++r) {
      FqMeta c = *reinterpret_cast<const FqMeta*>(cbase + (size_t)r * FQ_CANDB + FQ_ROWB);
      if (c.v > win.v || (c.v == win.v && c.idx < win.idx)) {
        win = c;
        wr = r;
      }
    }
    if (win.idx == 0x7fffffff) win.idx = 0;
    if (rank == 0 && tid == 0) out[pick] = win.idx;
    if (pick == m - 1) break;

    // new seed: FP32 fragments (prefetched; only the undecided rows need them) and its q row
    {
      const float4* srow = reinterpret_cast<const float4*>(fset + (int64_t)win.idx * FQ_D);
#pragma unroll
      for (int u = 0; u < 6; ++u) sfrag[u] = srow[sub + 8 * u];
    }
    uint32_t qs[FQ_QW];
    {
      const uint4* qsrc = reinterpret_cast<const uint4*>(cbase + (size_t)wr * FQ_CANDB);
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        uint4 t = qsrc[c];
        qs[4 * c + 0] = t.x;
        qs[4 * c + 1] = t.y;
        qs[4 * c + 2] = t.z;
        qs[4 * c + 3] = t.w;
      }
    }
    const float eps_s = win.eps;
    const int qn_s = win.qn;
    for (int base = 0; base < n_loc; base += FQ_THREADS) {
      const int li = base + tid;
      bool undecided = false;
      if (li < n_loc) {
        const uint4* qr = R.q(li);
        uint32_t dot = 0;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          uint4 t = qr[c];
          dot = __dp4a(t.x, qs[4 * c + 0], dot);
          dot = __dp4a(t.y, qs[4 * c + 1], dot);
          dot = __dp4a(t.z, qs[4 * c + 2], dot);
          dot = __dp4a(t.w, qs[4 * c + 3], dot);
        }
        const int I = *R.qn(li) + qn_s - 2 * (int)dot;  // |q_i - q_s|^2, exact
        const float lb = step * sqrtf((float)I) * 0.999999f - *R.eps(li) - eps_s;
        undecided = !(lb > 0.f && lb * lb * 0.9999f >= *R.dist(li));
      }
      unsigned mask = __ballot_sync(0xffffffffu, undecided);
      // the warp re-reads its undecided rows in FP32, four at a time (8 lanes per row)
      while (mask) {
        unsigned mm = mask;  // grp-th set bit
        for (int c = 0; c < grp; ++c) mm &= mm - 1;
        const int src = __ffs(mm) - 1;
        const bool valid = src >= 0;
        const int lr = base + (w << 5) + (valid ? src : 0);
        float acc = 0.f;
        if (valid) {
          const float4* xrow =
              reinterpret_cast<const float4*>(fset + (int64_t)(lr * CL + rank) * FQ_D);
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            float4 x = xrow[sub + 8 * u];
            float d0 = x.x - sfrag[u].x, d1 = x.y - sfrag[u].y, d2 = x.z - sfrag[u].z,
                  d3 = x.w - sfrag[u].w;
            acc = fmaf(d0, d0, acc);
            acc = fmaf(d1, d1, acc);
            acc = fmaf(d2, d2, acc);
            acc = fmaf(d3, d3, acc);
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (valid && sub == 0) {
          float* dp = R.dist(lr);
          *dp = fminf(*dp, acc);
        }
        // drop the (up to) four lowest set bits
#pragma unroll
        for (int c = 0; c < 4; ++c) mask &= mask - 1;
      }
      __syncwarp();
    }
  }
  // keep every CTA's shared memory alive until all remote writes/reads are done
  cluster.sync();
}

size_t fps_q8_spill_bytes(int64_t total_rows) { return (size_t)total_rows * FQ_ROWB; }

static size_t fq_smem_bytes(int r_res, int CL) {
  return (size_t)r_res * (FQ_ROWB + 12) + (size_t)2 * CL * FQ_CANDB;
}

// Largest number of resident rows per CTA (one CTA per SM).
static int fq_rows_per_cta(int CL) {
  const size_t budget = 227 * 1024 - 4096;  // static shared memory + slack
  int r = (int)((budget - (size_t)2 * CL * FQ_CANDB) / (FQ_ROWB + 12));
  return r & ~3;
}

// One class of sets: set index = (j / set_per) * set_stride + set_first + j % set_per, j < n_launch.
int launch_fps_q8(const float* feat, const int32_t* set_off, const int32_t* set_n, int n_launch,
                  int set_first, int set_per, int set_stride, int n_cap, int cl, int m_max,
                  int k_for_count, uint8_t* spill, int32_t* idx_out, int32_t* cnt_out,
                  cudaStream_t st) {
  if (!spill || n_launch <= 0 || n_launch > 65535) return R3DFS_E_BADARG;
  // default: the smallest cluster whose shared memory holds the largest possible set
  int CL = 1;
  if (cl > 0) {
    while (CL < cl && CL < FQ_MAX_CL) CL *= 2;
  } else {
    while (CL < FQ_MAX_CL && (int64_t)CL * fq_rows_per_cta(CL) < n_cap) CL *= 2;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cudaError_t e;
  for (;; CL = 8) {
    int r_res = min(fq_rows_per_cta(CL), ((n_cap + CL - 1) / CL + 3) & ~3);
    // phase 0 borrows the q area for (FQ_WARPS + CL) x 384 floats
    const int r_min = (int)(((size_t)(FQ_WARPS + CL) * 2 * FQ_D * sizeof(float) + FQ_ROWB - 1) /
                            FQ_ROWB);
    r_res = max(r_res, (r_min + 3) & ~3);
    const size_t smem = fq_smem_bytes(r_res, CL);
    // always the maximum: the attribute is per function, and concurrent host threads may launch
    // different set classes
    e = cudaFuncSetAttribute(fps_q8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024 - 4096);
    if (e != cudaSuccess) return (int)e;
    if (CL > 8) {
      e = cudaFuncSetAttribute(fps_q8_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return (int)e;
    }
    cfg = {};
    cfg.gridDim = dim3(CL, n_launch, 1);
    cfg.blockDim = dim3(FQ_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 1;
    if (CL > 8) {  // can the device co-schedule a 16-CTA cluster of this kernel at all?
      e = cudaOccupancyMaxActiveClusters(&n_clusters, fps_q8_kernel, &cfg);
      if (e != cudaSuccess) n_clusters = 0;
      (void)cudaGetLastError();
    }
    if (n_clusters >= 1) {
      e = cudaLaunchKernelEx(&cfg, fps_q8_kernel, feat, set_off, set_n, set_first, set_per,
                             set_stride, m_max, k_for_count, r_res, spill, idx_out, cnt_out);
      if (e != cudaSuccess) return (int)e;
      ++r3dfs_launches;
      return 0;
    }
    if (CL <= 8) return R3DFS_E_UNSUPPORTED;
  }
}
