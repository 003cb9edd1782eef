// Generic kernels of the meta-training path (reference models/mpti_learner.py:50-79 around
// MPTI_SelfAtten.forward(train=True)): strided FP32 GEMM for the weight/input gradients,
// batch-statistics BatchNorm forward/backward, the EdgeConv pieces on materialised edge activations,
// the attention map's softmax/dropout, and the fused Adam update.
//
// Unlike the inference path nothing here is fused away: training runs one episode (12 clouds) per
// step, the reference itself materialises every tensor autograd needs, and 180 GB of HBM hold them
// with room to spare.  What is kept from the inference design: point-major rows, the per-point
// first EdgeConv conv (W1 [x_j - x_i ; x_i] = W1a x_j + (W1b - W1a) x_i), tcgen05 forward GEMMs.
#include "train.cuh"

// ---------------------------------------------------------------------------------------------
// strided SGEMM, 128 x 64 x 16 tiles, 8 x 4 outputs per thread
// ---------------------------------------------------------------------------------------------
#define SG_BM 128
#define SG_BN 64
#define SG_BK 16

template <bool A_KFAST, bool B_NFAST>
__global__ __launch_bounds__(256) void sgemm_kernel(
    const float* __restrict__ A, int64_t sAm, int64_t sAk, int64_t bsA, const float* __restrict__ B,
    int64_t sBk, int64_t sBn, int64_t bsB, float* __restrict__ C, int64_t ldc, int64_t bsC, int M,
    int N, int K, int splits, int kchunk, float alpha, float beta, float* __restrict__ partial) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int z = blockIdx.z, b = z / splits, sp = z % splits;
  const int k_begin = sp * kchunk, k_end = min(K, k_begin + kchunk);
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;
  const float* Ab = A + (int64_t)b * bsA;
  const float* Bb = B + (int64_t)b * bsB;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += SG_BK) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + 256 * r;
      int mm, kk;
      if (A_KFAST) {
        kk = e & 15;
        mm = e >> 4;
      } else {
        mm = e & 127;
        kk = e >> 7;
      }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < k_end) ? Ab[gm * sAm + gk * sAk] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + 256 * r;
      int nn, kk;
      if (B_NFAST) {
        nn = e & 63;
        kk = e >> 6;
      } else {
        kk = e & 15;
        nn = e >> 4;
      }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < k_end) ? Bb[gk * sBk + gn * sBn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (splits == 1) {
        float* c = C + (int64_t)b * bsC + (int64_t)gm * ldc + gn;
        *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
      } else {
        partial[((int64_t)z * M + gm) * N + gn] = acc[i][j];
      }
    }
  }
}

__global__ void sgemm_reduce_kernel(const float* __restrict__ partial, int M, int N, int splits,
                                    float alpha, float beta, float* __restrict__ C, int64_t ldc,
                                    int64_t bsC) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * N) return;
  const int m = e / N, n = e % N;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += partial[((int64_t)(b * splits + sp) * M + m) * N + n];
  float* c = C + (int64_t)b * bsC + (int64_t)m * ldc + n;
  *c = alpha * s + (beta != 0.f ? beta * *c : 0.f);
}

int launch_sgemm_reduce(const float* partial, int M, int N, int splits, float alpha, float beta,
                        float* C, int64_t ldc, cudaStream_t st) {
  sgemm_reduce_kernel<<<dim3((M * N + 255) / 256, 1), 256, 0, st>>>(partial, M, N, splits, alpha,
                                                                   beta, C, ldc, 0);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

int sgemm_splits(int M, int N, int64_t K, int batch) {
  const int64_t tiles = (int64_t)((M + SG_BM - 1) / SG_BM) * ((N + SG_BN - 1) / SG_BN) * batch;
  if (tiles >= 148 || K < 2048) return 1;
  int64_t s = (4 * 148 + tiles - 1) / tiles;
  const int64_t by_k = (K + 511) / 512;
  if (s > by_k) s = by_k;
  if (s > 1024) s = 1024;
  return (int)(s < 1 ? 1 : s);
}

int launch_sgemm(const float* A, int64_t sAm, int64_t sAk, int64_t bsA, const float* B, int64_t sBk,
                 int64_t sBn, int64_t bsB, float* C, int64_t ldc, int64_t bsC, int M, int N, int K,
                 int batch, float alpha, float beta, int splits, float* partial, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return R3DFS_E_BADARG;
  if (splits < 1) splits = 1;
  if (splits > 1 && !partial) return R3DFS_E_BADARG;
  int kchunk = (K + splits - 1) / splits;
  kchunk = (kchunk + SG_BK - 1) / SG_BK * SG_BK;
  splits = (K + kchunk - 1) / kchunk;
  if ((int64_t)batch * splits > 65535) return R3DFS_E_UNSUPPORTED;
  dim3 grid((M + SG_BM - 1) / SG_BM, (N + SG_BN - 1) / SG_BN, batch * splits);
  const bool akf = sAk == 1, bnf = sBn == 1;
#define SG_LAUNCH(AK, BN_)                                                                       \
  sgemm_kernel<AK, BN_><<<grid, 256, 0, st>>>(A, sAm, sAk, bsA, B, sBk, sBn, bsB, C, ldc, bsC, M, \
                                              N, K, splits, kchunk, alpha, beta, partial)
  if (akf && bnf) SG_LAUNCH(true, true);
  else if (akf) SG_LAUNCH(true, false);
  else if (bnf) SG_LAUNCH(false, true);
  else SG_LAUNCH(false, false);
#undef SG_LAUNCH
  R3DFS_CHECK_LAUNCH();
  if (splits > 1) {
    sgemm_reduce_kernel<<<dim3((M * N + 255) / 256, batch), 256, 0, st>>>(partial, M, N, splits,
                                                                         alpha, beta, C, ldc, bsC);
    R3DFS_CHECK_LAUNCH();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// column reductions over (rows x C) matrices, C a multiple of 64: two sums per channel, FP64
// partials per row block, combined in block order (deterministic).
//   mode 0: (x, x^2)                 -> BatchNorm statistics
//   mode 1: (dz, dz * xhat)          -> BatchNorm backward sums, dz = dy * act'(gamma xhat + beta)
//   mode 2: (x, -)                   -> column sum
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == ACT_LRELU) return z > 0.f ? 1.f : 0.2f;
  return 1.f;
}

template <int MODE>
__global__ __launch_bounds__(256) void col_reduce_kernel(
    const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t ld_dy,
    int64_t rows, int C, int64_t rows_per_block, const float* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, int act,
    double* __restrict__ partial) {
  __shared__ double s0[256], s1[256];
  const int tid = threadIdx.x, c = blockIdx.x * 64 + (tid & 63), rl = tid >> 6;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  double a0 = 0.0, a1 = 0.0;
  float mean = 0.f, istd = 0.f, g = 0.f, bt = 0.f;
  if (MODE == 1) {
    mean = stats[c];
    istd = stats[C + c];
    g = gamma[c];
    bt = beta[c];
  }
  for (int64_t r = r0 + rl; r < r1; r += 4) {
    const float v = x[r * ldx + c];
    if (MODE == 0) {
      a0 += (double)v;
      a1 += (double)v * (double)v;
    } else if (MODE == 1) {
      const float xh = (v - mean) * istd;
      const float dz = dy[r * ld_dy + c] * act_grad(fmaf(g, xh, bt), act);
      a0 += (double)dz;
      a1 += (double)dz * (double)xh;
    } else {
      a0 += (double)v;
    }
  }
  s0[tid] = a0;
  s1[tid] = a1;
  __syncthreads();
  if (rl == 0) {
    a0 = s0[tid] + s0[tid + 64] + s0[tid + 128] + s0[tid + 192];
    a1 = s1[tid] + s1[tid + 64] + s1[tid + 128] + s1[tid + 192];
    partial[((int64_t)blockIdx.y * 2) * C + c] = a0;
    partial[((int64_t)blockIdx.y * 2 + 1) * C + c] = a1;
  }
}

static inline void col_reduce_grid(int64_t rows, int C, dim3& grid, int64_t& rpb) {
  int64_t nb = (rows + 63) / 64;
  const int64_t cap = BN_MAX_BLOCKS / (C / 64) > 0 ? BN_MAX_BLOCKS / (C / 64) : 1;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  rpb = (rows + nb - 1) / nb;
  nb = (rows + rpb - 1) / rpb;
  grid = dim3(C / 64, (unsigned)nb);
}

// Sums the per-row-block partials of one channel: 16 slices of blocks per channel (64 channels x 16
// slices = 1024 threads), slices combined in slice order -> deterministic.  Returns the totals to
// the slice-0 thread of every channel (other threads get zeros and `lead` = false).
__device__ __forceinline__ void finish_partials(const double* __restrict__ partial, int nb, int C,
                                                int c, double& s, double& q, bool& lead) {
  __shared__ double f0[1024], f1[1024];
  const int tid = threadIdx.x, sl = tid >> 6;
  const int per = (nb + 15) / 16;
  const int b0 = sl * per, b1 = min(nb, b0 + per);
  double a0 = 0.0, a1 = 0.0;
  for (int b = b0; b < b1; ++b) {
    a0 += partial[((int64_t)b * 2) * C + c];
    a1 += partial[((int64_t)b * 2 + 1) * C + c];
  }
  f0[tid] = a0;
  f1[tid] = a1;
  __syncthreads();
  lead = sl == 0;
  s = q = 0.0;
  if (lead) {
    for (int t = 0; t < 16; ++t) {
      s += f0[tid + 64 * t];
      q += f1[tid + 64 * t];
    }
  }
}

__global__ __launch_bounds__(1024) void bn_stats_finish_kernel(const double* __restrict__ partial, int nb, int C,
                                       int64_t rows, float eps, float momentum,
                                       float* __restrict__ running, float* __restrict__ stats) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  double s, q;
  bool lead;
  finish_partials(partial, nb, C, c, s, q, lead);
  if (!lead) return;
  const double mean = s / (double)rows;
  double var = q / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running) {
    const double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running[c] = (1.f - momentum) * running[c] + momentum * (float)mean;
    running[C + c] = (1.f - momentum) * running[C + c] + momentum * (float)unb;
  }
}

int launch_bn_stats(const float* x, int64_t ldx, int64_t rows, int C, float eps, float momentum,
                    float* running, float* stats, double* scratch, cudaStream_t st) {
  if (C % 64 != 0 || C > 512 || rows <= 0) return R3DFS_E_UNSUPPORTED;
  dim3 grid;
  int64_t rpb;
  col_reduce_grid(rows, C, grid, rpb);
  col_reduce_kernel<0><<<grid, 256, 0, st>>>(x, ldx, nullptr, 0, rows, C, rpb, nullptr, nullptr,
                                            nullptr, 0, scratch);
  R3DFS_CHECK_LAUNCH();
  bn_stats_finish_kernel<<<C / 64, 1024, 0, st>>>(scratch, (int)grid.y, C, rows, eps,
                                                         momentum, running, stats);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

__global__ void bn_act_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C,
                              const float* __restrict__ stats, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int act, float* __restrict__ y,
                              int64_t ldy) {
  const int c4n = C >> 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * c4n) return;
  const int64_t r = e / c4n;
  const int c = (int)(e % c4n) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
  const float in[4] = {v.x, v.y, v.z, v.w};
  float o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float xh = (in[q] - stats[c + q]) * stats[C + c + q];
    o[q] = apply_act(fmaf(gamma[c + q], xh, beta[c + q]), act);
  }
  *reinterpret_cast<float4*>(y + r * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
}

int launch_bn_act(const float* x, int64_t ldx, int64_t rows, int C, const float* stats,
                  const float* gamma, const float* beta, int act, float* y, int64_t ldy,
                  cudaStream_t st) {
  const int64_t total = rows * (C >> 2);
  bn_act_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ldx, rows, C, stats, gamma, beta,
                                                                act, y, ldy);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// sums -> coefficients (sum dz / rows, sum dz xhat / rows) in coef[2C]; dgamma += , dbeta +=
__global__ __launch_bounds__(1024) void bn_bwd_finish_kernel(
    const double* __restrict__ partial, int nb, int C, int64_t rows, float* __restrict__ coef,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  double s, q;
  bool lead;
  finish_partials(partial, nb, C, c, s, q, lead);
  if (!lead) return;
  coef[c] = (float)(s / (double)rows);
  coef[C + c] = (float)(q / (double)rows);
  if (dbeta) dbeta[c] += (float)s;
  if (dgamma) dgamma[c] += (float)q;
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, int64_t ld_dy,
                                    const float* __restrict__ x, int64_t ldx, int64_t rows, int C,
                                    const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, int act,
                                    const float* __restrict__ coef, float* __restrict__ dx,
                                    int64_t ld_dx) {
  const int c4n = C >> 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * c4n) return;
  const int64_t r = e / c4n;
  const int c = (int)(e % c4n) * 4;
  const float4 xv = *reinterpret_cast<const float4*>(x + r * ldx + c);
  const float4 dv = *reinterpret_cast<const float4*>(dy + r * ld_dy + c);
  const float xi[4] = {xv.x, xv.y, xv.z, xv.w};
  const float di[4] = {dv.x, dv.y, dv.z, dv.w};
  float o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float istd = stats[C + c + q], g = gamma[c + q];
    const float xh = (xi[q] - stats[c + q]) * istd;
    const float dz = di[q] * act_grad(fmaf(g, xh, beta[c + q]), act);
    o[q] = g * istd * (dz - coef[c + q] - xh * coef[C + c + q]);
  }
  *reinterpret_cast<float4*>(dx + r * ld_dx + c) = make_float4(o[0], o[1], o[2], o[3]);
}

int launch_bn_act_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ldx, int64_t rows,
                      int C, const float* stats, const float* gamma, const float* beta, int act,
                      float* dx, int64_t ld_dx, float* dgamma, float* dbeta, double* scratch,
                      cudaStream_t st) {
  if (C % 64 != 0 || C > 512 || rows <= 0) return R3DFS_E_UNSUPPORTED;
  dim3 grid;
  int64_t rpb;
  col_reduce_grid(rows, C, grid, rpb);
  col_reduce_kernel<1><<<grid, 256, 0, st>>>(x, ldx, dy, ld_dy, rows, C, rpb, stats, gamma, beta,
                                            act, scratch);
  R3DFS_CHECK_LAUNCH();
  float* coef = reinterpret_cast<float*>(scratch + (size_t)2 * 512 * BN_MAX_BLOCKS);
  bn_bwd_finish_kernel<<<C / 64, 1024, 0, st>>>(scratch, (int)grid.y, C, rows, coef, dgamma,
                                                       dbeta);
  R3DFS_CHECK_LAUNCH();
  const int64_t total = rows * (C >> 2);
  bn_bwd_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      dy, ld_dy, x, ldx, rows, C, stats, gamma, beta, act, coef, dx, ld_dx);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

__global__ __launch_bounds__(1024) void col_sum_finish_kernel(const double* __restrict__ partial,
                                                              int nb, int C, float* __restrict__ out) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  double s, q;
  bool lead;
  finish_partials(partial, nb, C, c, s, q, lead);
  if (lead) out[c] += (float)s;
}

int launch_col_sum_acc(const float* x, int64_t ldx, int64_t rows, int C, float* out, double* scratch,
                       cudaStream_t st) {
  if (C % 64 != 0 || C > 512 || rows <= 0) return R3DFS_E_UNSUPPORTED;
  dim3 grid;
  int64_t rpb;
  col_reduce_grid(rows, C, grid, rpb);
  col_reduce_kernel<2><<<grid, 256, 0, st>>>(x, ldx, nullptr, 0, rows, C, rpb, nullptr, nullptr,
                                            nullptr, 0, scratch);
  R3DFS_CHECK_LAUNCH();
  col_sum_finish_kernel<<<C / 64, 1024, 0, st>>>(scratch, (int)grid.y, C, out);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// EdgeConv on materialised edge rows (edge e = point * k + slot, 64 channels)
// ---------------------------------------------------------------------------------------------
// h1pre[e] = P[nbr] + Q[point]   (the first 1x1 conv of models/dgcnn.py:45-61 done per point)
__global__ void edge_pre_kernel(const float* __restrict__ PQ, const int32_t* __restrict__ idx, int N,
                                int k, int64_t total, float* __restrict__ h1) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c4 = (int)(t & 15);
  const int64_t e = t >> 4;
  const int64_t pt = e / k;
  const int64_t base = pt / N * N;
  const int64_t nb = base + idx[e];
  const float4 p = __ldg(reinterpret_cast<const float4*>(PQ + nb * 128) + c4);
  const float4 q = __ldg(reinterpret_cast<const float4*>(PQ + pt * 128 + 64) + c4);
  reinterpret_cast<float4*>(h1 + e * 64)[c4] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
}

int launch_edge_pre(const float* PQ, const int32_t* idx, int64_t B, int N, int k, float* h1pre,
                    cudaStream_t st) {
  const int64_t total = B * N * k * 16;
  edge_pre_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(PQ, idx, N, k, total, h1pre);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// y[point][c] = max over slots of LReLU(BN(h2pre)), first maximum's slot kept for the backward
__global__ void edge_max_kernel(const float* __restrict__ h2, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                int64_t M, int k, float* __restrict__ y, int64_t ldy,
                                uint8_t* __restrict__ arg) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * 64) return;
  const int c = (int)(t & 63);
  const int64_t pt = t >> 6;
  const float mean = stats[c], istd = stats[64 + c], g = gamma[c], bt = beta[c];
  float best = -INFINITY;
  int bj = 0;
  for (int j = 0; j < k; ++j) {
    float z = fmaf(g, (h2[(pt * k + j) * 64 + c] - mean) * istd, bt);
    z = z > 0.f ? z : 0.2f * z;
    if (z > best) {
      best = z;
      bj = j;
    }
  }
  y[pt * ldy + c] = best;
  arg[t] = (uint8_t)bj;
}

int launch_edge_max(const float* h2pre, const float* stats, const float* gamma, const float* beta,
                    int64_t M, int k, float* y, int64_t ldy, uint8_t* arg, cudaStream_t st) {
  edge_max_kernel<<<(unsigned)((M * 64 + 255) / 256), 256, 0, st>>>(h2pre, stats, gamma, beta, M, k,
                                                                   y, ldy, arg);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

__global__ void edge_max_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy,
                                    const uint8_t* __restrict__ arg, int64_t M, int k,
                                    float* __restrict__ dA2) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * k * 64) return;
  const int c = (int)(t & 63);
  const int64_t e = t >> 6;
  const int64_t pt = e / k;
  const int j = (int)(e - pt * k);
  dA2[t] = arg[pt * 64 + c] == j ? dy[pt * ld_dy + c] : 0.f;
}

int launch_edge_max_bwd(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t M, int k,
                        float* dA2, cudaStream_t st) {
  const int64_t total = M * k * 64;
  edge_max_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, ld_dy, arg, M, k, dA2);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// adjoint of edge_pre: dQ[point] = sum over its slots (ordered), dP[nbr] += (float atomics — the
// one place training is not bit-reproducible, as in the reference's CUDA index_select backward)
__global__ void edge_pre_bwd_kernel(const float* __restrict__ dh1, const int32_t* __restrict__ idx,
                                    int N, int k, int64_t M, float* __restrict__ dPQ) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * 64) return;
  const int c = (int)(t & 63);
  const int64_t pt = t >> 6;
  const int64_t base = pt / N * N;
  float q = 0.f;
  for (int j = 0; j < k; ++j) {
    const float d = dh1[(pt * k + j) * 64 + c];
    q += d;
    atomicAdd(dPQ + (base + idx[pt * k + j]) * 128 + c, d);
  }
  dPQ[pt * 128 + 64 + c] = q;
}

int launch_edge_pre_bwd(const float* dh1, const int32_t* idx, int64_t B, int N, int k, float* dPQ,
                        cudaStream_t st) {
  const int64_t M = B * N;
  cudaError_t e = cudaMemsetAsync(dPQ, 0, sizeof(float) * M * 128, st);
  if (e != cudaSuccess) return (int)e;
  edge_pre_bwd_kernel<<<(unsigned)((M * 64 + 255) / 256), 256, 0, st>>>(dh1, idx, N, k, M, dPQ);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// dW1 (64, 2C) += unfold of d[W1a ; W1b - W1a] (128, C):  dW1a = dP - dQ,  dW1b = dQ
__global__ void unfold_w1_grad_kernel(const float* __restrict__ dWf, int C, float* __restrict__ dW1) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 64 * C) return;
  const int o = e / C, kk = e % C;
  const float dp = dWf[o * C + kk], dq = dWf[(64 + o) * C + kk];
  dW1[o * 2 * C + kk] += dp - dq;
  dW1[o * 2 * C + C + kk] += dq;
}

int launch_unfold_w1_grad(const float* dWf, int C, float* dW1, cudaStream_t st) {
  unfold_w1_grad_kernel<<<(64 * C + 255) / 256, 256, 0, st>>>(dWf, C, dW1);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// attention map (models/attention.py:43-46): row softmax, dropout with a stored keep mask
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* s_red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();
  if (lane == 0) s_red[w] = v;
  __syncthreads();
  float r = s_red[0];
  for (int q = 1; q < (int)(blockDim.x >> 5); ++q) r = is_max ? fmaxf(r, s_red[q]) : r + s_red[q];
  return r;
}

__global__ __launch_bounds__(256) void softmax_rows_kernel(float* __restrict__ S, int n,
                                                          const uint8_t* __restrict__ mask, float p,
                                                          float* __restrict__ Pd) {
  __shared__ float s_red[8];
  const int64_t row = blockIdx.x;
  float* s = S + row * n;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, s[j]);
  mx = block_reduce(mx, true, s_red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float e = expf(s[j] - mx);
    s[j] = e;
    sum += e;
  }
  sum = block_reduce(sum, false, s_red);
  const float inv = 1.f / sum, keep_scale = 1.f / (1.f - p);
  for (int j = threadIdx.x; j < n; j += 256) {
    const float pv = s[j] * inv;
    s[j] = pv;
    if (mask) Pd[row * n + j] = mask[row * n + j] ? pv * keep_scale : 0.f;
  }
}

int launch_softmax_rows(float* S, int64_t rows, int n, const uint8_t* mask, float p, float* Pd,
                        cudaStream_t st) {
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(S, n, mask, p, Pd);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

__global__ __launch_bounds__(256) void softmax_rows_bwd_kernel(const float* __restrict__ P,
                                                              float* __restrict__ dPd, int n,
                                                              const uint8_t* __restrict__ mask,
                                                              float p) {
  __shared__ float s_red[8];
  const int64_t row = blockIdx.x;
  const float* pr = P + row * n;
  float* d = dPd + row * n;
  const float keep_scale = 1.f / (1.f - p);
  float dot = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) {
    float dp = d[j];
    if (mask) dp = mask[row * n + j] ? dp * keep_scale : 0.f;
    d[j] = dp;
    dot += dp * pr[j];
  }
  dot = block_reduce(dot, false, s_red);
  for (int j = threadIdx.x; j < n; j += 256) d[j] = pr[j] * (d[j] - dot);
}

int launch_softmax_rows_bwd(const float* P, float* dPd, int64_t rows, int n, const uint8_t* mask,
                            float p, cudaStream_t st) {
  softmax_rows_bwd_kernel<<<(unsigned)rows, 256, 0, st>>>(P, dPd, n, mask, p);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// counter-based keep mask: element i is kept iff u(seed, i) >= p, u from a splitmix64 hash
__global__ void dropout_mask_kernel(uint64_t seed, int64_t n, float p, uint8_t* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  mask[i] = u >= p ? 1 : 0;
}

int launch_dropout_mask(uint64_t seed, int64_t n, float p, uint8_t* mask, cudaStream_t st) {
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, n, p, mask);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
template <bool ADD>
__global__ void cols_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int ncols,
                            float* __restrict__ dst, int64_t ldd) {
  const int c4n = ncols >> 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * c4n) return;
  const int64_t r = e / c4n;
  const int c = (int)(e % c4n) * 4;
  const float4 s = *reinterpret_cast<const float4*>(src + r * lds + c);
  float4* d = reinterpret_cast<float4*>(dst + r * ldd + c);
  if (ADD) {
    float4 o = *d;
    o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
    *d = o;
  } else {
    *d = s;
  }
}

int launch_add_cols(const float* src, int64_t lds, int64_t rows, int ncols, float* dst, int64_t ldd,
                    cudaStream_t st) {
  const int64_t total = rows * (ncols >> 2);
  cols_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, lds, rows, ncols, dst, ldd);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

int launch_copy_cols_plain(const float* src, int64_t lds, int64_t rows, int ncols, float* dst,
                           int64_t ldd, cudaStream_t st) {
  const int64_t total = rows * (ncols >> 2);
  cols_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, lds, rows, ncols, dst, ldd);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols,
                                 float* __restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  const int r = e / cols, c = e % cols;
  dst[c * rows + r] = src[e];
}

int launch_transpose(const float* src, int rows, int cols, float* dst, cudaStream_t st) {
  transpose_kernel<<<(rows * cols + 255) / 256, 256, 0, st>>>(src, rows, cols, dst);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) p[e] = v;
}

int launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st) {
  fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults as used at models/mpti_learner.py:26-32): two learning-rate
// groups split at n_group0 (encoder | rest); grad_scale folds the data-parallel mean.
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, int64_t n_group0, float lr0, float lr1,
                            float beta1, float beta2, float eps, float bc1, float bc2,
                            float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = g[i] * grad_scale;
  const float mi = beta1 * m[i] + (1.f - beta1) * gr;
  const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
  m[i] = mi;
  v[i] = vi;
  const float lr = i < n_group0 ? lr0 : lr1;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_group0,
                float lr0, float lr1, float beta1, float beta2, float eps, float bc1, float bc2,
                float grad_scale, cudaStream_t st) {
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, n_group0, lr0, lr1, beta1,
                                                          beta2, eps, bc1, bc2, grad_scale);
  R3DFS_CHECK_LAUNCH();
  return 0;
}
