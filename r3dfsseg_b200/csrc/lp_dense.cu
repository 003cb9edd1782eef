// Dense cross-check of the label-propagation solve (reference models/mpti.py:758-776).
// The reference inverts the dense (I - alpha S); the product path solves the same system by
// conjugate gradients on the sparse graph (lp.cu).  This file solves it a third way, entirely on
// the GPU and in FP64: the merged symmetric rows are expanded to the dense SPD matrix
// M = I - alpha S, factored M = L L^T by a blocked right-looking Cholesky, and Z = M^-1 Y comes from
// two triangular solves.  It is not on the timed path: n^3/3 = 29 GFLOP at n = 4416 in plain FP64
// FMAs (a few ms per graph), meant for `r3dfs_lp_cholesky` — the in-library yardstick the CG solve
// is checked against (tests/test_gpu_parity.py::test_label_propagate_vs_dense_solve).
#include "common.cuh"
#include "lp.cuh"

#define CH_B 64  // block size of the factorisation

// M = I - alpha S from the merged rows (rows of invalid nodes: identity)
__global__ void dense_system_kernel(const int32_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ rowlen,
                                    const uint16_t* __restrict__ mcol,
                                    const float* __restrict__ mval, int nn, int k, float alpha,
                                    double* __restrict__ M) {
  const int g = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= nn) return;
  const int lane = threadIdx.x & 31;
  const int64_t vb = (int64_t)g * nn;
  double* row = M + (vb + i) * (int64_t)nn;
  if (lane == 0) row[i] = 1.0;
  const int L = rowlen[vb + i];
  const int64_t mb = vb * lp_rowcap(k) + rowptr[vb + i];
  // a column occurs at most once per merged row, and never on the diagonal
  for (int t = lane; t < L; t += 32) row[mcol[mb + t]] = -(double)alpha * (double)mval[mb + t];
}

// Cholesky of the diagonal block (j, j): one CTA, the block in shared memory
__global__ __launch_bounds__(CH_B) void chol_diag_kernel(double* __restrict__ M, int nn, int j0,
                                                         int* __restrict__ info) {
  __shared__ double a[CH_B][CH_B + 1];
  const int g = blockIdx.x, t = threadIdx.x;
  double* A = M + (int64_t)g * nn * nn;
  const int nb = min(CH_B, nn - j0);
  for (int c = 0; c < nb; ++c) a[t][c] = t < nb ? A[(int64_t)(j0 + t) * nn + j0 + c] : 0.0;
  __syncthreads();
  for (int c = 0; c < nb; ++c) {
    if (t == c) {
      const double d = a[c][c];
      if (!(d > 0.0)) atomicExch(&info[g], j0 + c + 1);  // not positive definite
      a[c][c] = sqrt(fmax(d, 1e-300));
    }
    __syncthreads();
    if (t > c && t < nb) a[t][c] /= a[c][c];
    __syncthreads();
    if (t > c && t < nb)
      for (int c2 = c + 1; c2 <= t; ++c2) a[t][c2] -= a[t][c] * a[c2][c];
    __syncthreads();
  }
  if (t < nb)
    for (int c = 0; c <= t; ++c) A[(int64_t)(j0 + t) * nn + j0 + c] = a[t][c];
}

// panel below the diagonal block: L_ij = A_ij L_jj^-T ; one thread per row of the panel
__global__ __launch_bounds__(CH_B) void chol_trsm_kernel(double* __restrict__ M, int nn, int j0) {
  __shared__ double l[CH_B][CH_B + 1];
  const int g = blockIdx.y, t = threadIdx.x;
  double* A = M + (int64_t)g * nn * nn;
  const int nb = min(CH_B, nn - j0);
  for (int c = 0; c < nb; ++c) l[t][c] = (t < nb && c <= t) ? A[(int64_t)(j0 + t) * nn + j0 + c] : 0.0;
  __syncthreads();
  const int i = j0 + nb + blockIdx.x * CH_B + t;
  if (i >= nn) return;
  double x[CH_B];
  double* row = A + (int64_t)i * nn + j0;
#pragma unroll 8
  for (int c = 0; c < CH_B; ++c) x[c] = c < nb ? row[c] : 0.0;
  for (int c = 0; c < nb; ++c) {
    double v = x[c];
    for (int q = 0; q < c; ++q) v -= x[q] * l[c][q];
    x[c] = v / l[c][c];
  }
  for (int c = 0; c < nb; ++c) row[c] = x[c];
}

// trailing update of the lower triangle: A_ik -= L_ij L_kj^T for block rows i >= k > j
__global__ __launch_bounds__(256) void chol_update_kernel(double* __restrict__ M, int nn, int j0) {
  __shared__ double li[CH_B][CH_B / 2 + 1], lk[CH_B][CH_B / 2 + 1];  // half of the panel depth at a time
  const int g = blockIdx.z;
  const int bi = blockIdx.y, bk = blockIdx.x;
  if (bk > bi) return;
  double* A = M + (int64_t)g * nn * nn;
  const int nb = min(CH_B, nn - j0);
  const int r0 = j0 + nb + bi * CH_B, c0 = j0 + nb + bk * CH_B;
  if (r0 >= nn || c0 >= nn) return;
  const int t = threadIdx.x;
  const int tr = (t / 16) * 4, tc = (t % 16) * 4;  // 4 x 4 outputs per thread
  double acc[4][4] = {};
  for (int h0 = 0; h0 < nb; h0 += CH_B / 2) {
    __syncthreads();
    for (int e = t; e < CH_B * (CH_B / 2); e += 256) {
      const int r = e / (CH_B / 2), c = e % (CH_B / 2);
      li[r][c] = (r0 + r < nn && h0 + c < nb) ? A[(int64_t)(r0 + r) * nn + j0 + h0 + c] : 0.0;
      lk[r][c] = (c0 + r < nn && h0 + c < nb) ? A[(int64_t)(c0 + r) * nn + j0 + h0 + c] : 0.0;
    }
    __syncthreads();
    for (int q = 0; q < CH_B / 2; ++q) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = li[tr + u][q];
        b[u] = lk[tc + u][q];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = r0 + tr + u, c = c0 + tc + v;
      if (r < nn && c < nn && c <= r) A[(int64_t)r * nn + c] -= acc[u][v];
    }
}

// Z = L^-T L^-1 Y for nc <= 8 right-hand sides: one CTA per graph, block by block
__global__ __launch_bounds__(1024) void chol_solve_kernel(const double* __restrict__ M, int nn,
                                                          const float* __restrict__ Y, int nc,
                                                          const uint8_t* __restrict__ valid,
                                                          double* __restrict__ B,
                                                          float* __restrict__ Z) {
  __shared__ double xs[CH_B][8];
  __shared__ double red[8][CH_B][8];
  const int g = blockIdx.x, t = threadIdx.x;
  const double* A = M + (int64_t)g * nn * nn;
  double* b = B + (int64_t)g * nn * 8;
  for (int e = t; e < nn * 8; e += 1024) {
    const int i = e >> 3, c = e & 7;
    b[e] = (c < nc && valid[(int64_t)g * nn + i]) ? (double)Y[((int64_t)g * nn + i) * nc + c] : 0.0;
  }
  __syncthreads();
  // forward: L y = b
  for (int j0 = 0; j0 < nn; j0 += CH_B) {
    const int nb = min(CH_B, nn - j0);
    if (t < 8) {  // thread c solves the block's rows for right-hand side c
      for (int r = 0; r < nb; ++r) {
        double v = b[(int64_t)(j0 + r) * 8 + t];
        for (int q = 0; q < r; ++q) v -= A[(int64_t)(j0 + r) * nn + j0 + q] * xs[q][t];
        v /= A[(int64_t)(j0 + r) * nn + j0 + r];
        xs[r][t] = v;
        b[(int64_t)(j0 + r) * 8 + t] = v;
      }
    }
    __syncthreads();
    for (int e = t; e < (nn - j0 - nb) * 8; e += 1024) {
      const int i = j0 + nb + (e >> 3), c = e & 7;
      double v = 0.0;
      const double* row = A + (int64_t)i * nn + j0;
      for (int q = 0; q < nb; ++q) v = fma(row[q], xs[q][c], v);
      b[(int64_t)i * 8 + c] -= v;
    }
    __syncthreads();
  }
  // backward: L^T z = y  (column access of L: rows below the block)
  for (int j0 = ((nn - 1) / CH_B) * CH_B; j0 >= 0; j0 -= CH_B) {
    const int nb = min(CH_B, nn - j0);
    // subtract the contribution of the already solved rows i >= j0 + nb: sum_i L[i][j0+q] z_i.
    // Thread (row group rg, column q) walks rows i = j0 + nb + rg, + 8, ... (a row's 64 columns are
    // contiguous, so a warp reads 256 contiguous bytes), then the 8 groups are added in order.
    {
      const int q = t & 63, rg = t >> 6;
      double acc[8] = {};
      if (rg < 8 && q < nb)
        for (int i = j0 + nb + rg; i < nn; i += 8) {
          const double l = A[(int64_t)i * nn + j0 + q];
          const double* z = b + (int64_t)i * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fma(l, z[c], acc[c]);
        }
      if (rg < 8)
#pragma unroll
        for (int c = 0; c < 8; ++c) red[rg][q][c] = acc[c];
      __syncthreads();
      for (int e = t; e < nb * 8; e += 1024) {
        const int q2 = e >> 3, c = e & 7;
        double v = 0.0;
        for (int r = 0; r < 8; ++r) v += red[r][q2][c];
        xs[q2][c] = b[(int64_t)(j0 + q2) * 8 + c] - v;
      }
    }
    __syncthreads();
    if (t < 8) {
      for (int r = nb - 1; r >= 0; --r) {
        double v = xs[r][t];
        for (int q = r + 1; q < nb; ++q) v -= A[(int64_t)(j0 + q) * nn + j0 + r] * xs[q][t];
        v /= A[(int64_t)(j0 + r) * nn + j0 + r];
        xs[r][t] = v;
        b[(int64_t)(j0 + r) * 8 + t] = v;
      }
    }
    __syncthreads();
  }
  for (int e = t; e < nn * nc; e += 1024) {
    const int i = e / nc, c = e % nc;
    Z[((int64_t)g * nn + i) * nc + c] = (float)b[(int64_t)i * 8 + c];
  }
}

size_t lp_cholesky_scratch_bytes(int G, int nn) {
  return align_up(sizeof(double) * (size_t)G * nn * nn, 256) +
         align_up(sizeof(double) * (size_t)G * nn * 8, 256) + align_up(sizeof(int) * (size_t)G, 256);
}

int launch_lp_cholesky_solve(const int32_t* rowptr, const int32_t* rowlen, const uint16_t* mcol,
                             const float* mval, const uint8_t* valid, int G, int nn, int k,
                             const float* Y, int nc, float alpha, float* Z, void* scratch,
                             int32_t* info_out, cudaStream_t st) {
  if (nc < 1 || nc > 8 || G > 65535) return R3DFS_E_UNSUPPORTED;
  unsigned char* sp = reinterpret_cast<unsigned char*>(scratch);
  double* M = reinterpret_cast<double*>(sp);
  double* B = reinterpret_cast<double*>(sp + align_up(sizeof(double) * (size_t)G * nn * nn, 256));
  int* info = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(B) +
                                     align_up(sizeof(double) * (size_t)G * nn * 8, 256));
  cudaError_t e = cudaMemsetAsync(M, 0, sizeof(double) * (size_t)G * nn * nn, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(info, 0, sizeof(int) * (size_t)G, st);
  if (e != cudaSuccess) return (int)e;
  dense_system_kernel<<<dim3((nn + 7) / 8, G), 256, 0, st>>>(rowptr, rowlen, mcol, mval, nn, k, alpha, M);
  R3DFS_CHECK_LAUNCH();
  for (int j0 = 0; j0 < nn; j0 += CH_B) {
    chol_diag_kernel<<<G, CH_B, 0, st>>>(M, nn, j0, info);
    R3DFS_CHECK_LAUNCH();
    const int rest = nn - j0 - CH_B;
    if (rest <= 0) break;
    const int nbk = (rest + CH_B - 1) / CH_B;
    chol_trsm_kernel<<<dim3(nbk, G), CH_B, 0, st>>>(M, nn, j0);
    R3DFS_CHECK_LAUNCH();
    chol_update_kernel<<<dim3(nbk, nbk, G), 256, 0, st>>>(M, nn, j0);
    R3DFS_CHECK_LAUNCH();
  }
  chol_solve_kernel<<<G, 1024, 0, st>>>(M, nn, Y, nc, valid, B, Z);
  R3DFS_CHECK_LAUNCH();
  if (info_out) {
    e = cudaMemcpyAsync(info_out, info, sizeof(int) * (size_t)G, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}
